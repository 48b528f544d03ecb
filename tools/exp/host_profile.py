"""Where the host time of a small transform goes: cProfile of 300 forward + inverse calls on one 512 x 512 image."""
import cProfile, pstats, sys, os, io
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
import dtcwt_b200
xf = dtcwt_b200.Transform2d("near_sym_b", "qshift_b")
X = torch.rand(512, 512, device="cuda")
for _ in range(50):
    Z = xf.inverse(xf.forward(X, 4))
torch.cuda.synchronize()
pr = cProfile.Profile()
pr.enable()
for _ in range(300):
    Z = xf.inverse(xf.forward(X, 4))
torch.cuda.synchronize()
pr.disable()
s = io.StringIO()
pstats.Stats(pr, stream=s).sort_stats("tottime").print_stats(22)
print(s.getvalue()[:6000])
