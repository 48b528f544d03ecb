"""Experiment: does running two half-batches on two CUDA streams beat one full batch on one stream?"""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
import dtcwt_b200
xf = dtcwt_b200.Transform2d("near_sym_b", "qshift_b")
dev = torch.device("cuda", 0)
pool = [torch.rand((16, 4096, 4096), device=dev) for _ in range(3)]
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()

def one(i):
    p = xf.forward_channels(pool[i % 3], "nhw", nlevels=4)
    return xf.inverse_channels(p, "nhw")

def two(i, parts=2):
    X = pool[i % 3]
    main = torch.cuda.current_stream()
    outs = []
    n = X.shape[0] // parts
    streams = [s1, s2][:parts]
    for k, s in enumerate(streams):
        s.wait_stream(main)
        with torch.cuda.stream(s):
            p = xf.forward_channels(X[k * n:(k + 1) * n], "nhw", nlevels=4)
            outs.append(xf.inverse_channels(p, "nhw"))
    for s in streams:
        main.wait_stream(s)
    return outs

def staggered(i):
    """half A forward, then (half A inverse || half B forward), then half B inverse"""
    return two(i)

for name, fn in (("one stream, 16 images", one), ("two streams, 2 x 8 images", two)):
    for i in range(3): fn(i)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(10): fn(i + 3)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 10
    print("%-30s %.3f ms/step  %.1f Gpix/s" % (name, ms, 16 * 4096 * 4096 / ms / 1e6))
