"""BASELINE configs[1]: one 512x512 image (mandrill size), 4 levels, near_sym_b + qshift_b: latency of forward+inverse
through the public API (device-resident input), CUDA events, median of 200."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np, torch
import dtcwt_b200
xf = dtcwt_b200.Transform2d("near_sym_b", "qshift_b")
try:
    X = torch.from_numpy(np.load(os.path.join(os.path.dirname(__file__), "..", "..", "tests", "golden", "inputs.npz"))["mandrill"]).cuda()
except Exception:
    X = torch.rand(512, 512, device="cuda")
for _ in range(20):
    Z = xf.inverse(xf.forward(X, 4))
torch.cuda.synchronize()
ts = []
for _ in range(200):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); Z = xf.inverse(xf.forward(X, 4)); e1.record(); torch.cuda.synchronize()
    ts.append(e0.elapsed_time(e1))
ts.sort()
print("config 2 (512x512, 4 levels, fwd+inv, 8 launches): median %.1f us, min %.1f us, recon err %.2e" % (
    1e3 * ts[100], 1e3 * ts[0], float((Z - X).abs().max())))

# the same call captured once in a CUDA graph and replayed (dtcwt_b200.graph.Graphed): the host work of the eight launches is gone
rt = dtcwt_b200.graph.Graphed(lambda x: xf.inverse(xf.forward(x, 4)), X)
for _ in range(20):
    Zg = rt(X)
torch.cuda.synchronize()
ts = []
for _ in range(200):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); Zg = rt(X); e1.record(); torch.cuda.synchronize()
    ts.append(e0.elapsed_time(e1))
ts.sort()
print("config 2, CUDA-graph replay: median %.1f us, min %.1f us, max |graph - eager| %.2e, recon err %.2e" % (
    1e3 * ts[100], 1e3 * ts[0], float((Zg - Z).abs().max()), float((Zg - X).abs().max())))
