"""CUDA-graph replay of a fixed-shape transform.

A small transform is bound by its launches, not by the GPU: config 2 of BASELINE.json (one 512 x 512 image, four
levels, forward + inverse) is eight kernels of a few microseconds each behind ~45 us of host work per launch
(Python, ctypes, the launch itself).  Capturing the call once and replaying the graph removes that host work for
every later call with the same shape: the C ABI launches on the stream it is given, so it is capturable as is; the
outputs and scratch buffers allocated during the capture stay alive in the graph's private memory pool.

    rt = dtcwt_b200.graph.Graphed(lambda x: xf.inverse(xf.forward(x, 4)), example)
    z = rt(image)            # copies `image` into the captured input, replays, returns the captured output

The returned tensors (or the tensors inside a returned Pyramid / tuple / list) are the graph's static outputs: they are
overwritten by the next call, so copy what must outlive it.  This is a convenience on top of the reference's API,
not part of it; nothing else in the package depends on it.
"""
from __future__ import annotations

import torch

__all__ = ["Graphed"]


class Graphed(object):
    """``fn`` captured for inputs shaped like ``example`` (a CUDA tensor); call it with tensors of that shape/dtype."""

    def __init__(self, fn, example, warmup=3):
        if not isinstance(example, torch.Tensor) or example.device.type != "cuda":
            raise ValueError("Graphed needs a CUDA tensor as the example input")
        self._fn = fn
        self.static_input = example.detach().clone()
        side = torch.cuda.Stream(device=example.device)
        side.wait_stream(torch.cuda.current_stream(example.device))
        with torch.cuda.stream(side):                 # warm-up outside the capture: lazy initialisation, allocator pools
            for _ in range(max(1, int(warmup))):
                fn(self.static_input)
        torch.cuda.current_stream(example.device).wait_stream(side)
        torch.cuda.synchronize(example.device)
        self.graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.graph):
            self.static_output = fn(self.static_input)

    def __call__(self, x):
        if tuple(x.shape) != tuple(self.static_input.shape) or x.dtype != self.static_input.dtype:
            raise ValueError("Graphed was captured for %s %s, got %s %s" % (
                tuple(self.static_input.shape), self.static_input.dtype, tuple(x.shape), x.dtype))
        self.static_input.copy_(x, non_blocking=True)
        self.graph.replay()
        return self.static_output
