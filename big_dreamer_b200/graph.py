"""CUDA-graph capture of a whole hot-path step (forward + backward).

The library only enqueues kernels on the current stream, never synchronises and takes all its
memory from PyTorch's allocator, so a complete actor-loss step -- imagine_ahead, the reward / value
heads, lambda_return, the loss and ``loss.backward()`` (~45 kernels at the default sizes) -- can be
captured once and replayed as ONE graph launch.  That removes the per-kernel launch gaps and all of
the host work (ctypes calls, autograd bookkeeping) from the step.

    step = bd.CapturedStep(fn, static_inputs)     # fn(*static_inputs) -> tensor or tuple of tensors
    out = step(*new_inputs)                       # copies into the static inputs, replays

Rules (those of torch.cuda.graph): shapes are fixed; ``fn`` must not synchronise or read values on
the host; random draws inside ``fn`` use torch's graph-safe generator; gradients produced by
``fn`` live in the graph's private pool and are overwritten by every replay (set ``p.grad = None``
inside ``fn`` so the backward allocates them during capture; return the ``.grad`` tensors from ``fn`` if
other code may rebind ``p.grad`` between replays).
"""
from __future__ import annotations

import os
from typing import Callable, Sequence

import torch


class CapturedStep:
    def __init__(self, fn: Callable, static_inputs: Sequence[torch.Tensor] = (), warmup: int = 3):
        self.fn = fn
        self.static_inputs = list(static_inputs)
        self.graph = torch.cuda.CUDAGraph()
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):           # warm-up off the default stream (allocator, caches)
            for _ in range(warmup):
                fn(*self.static_inputs)
        torch.cuda.current_stream().wait_stream(side)
        torch.cuda.synchronize()
        # capture_error_mode: torch's default ("global") also polices CUDA calls that are merely *potentially* unsafe
        # under capture (attribute / occupancy queries the library makes the first time it meets a kernel variant, the
        # allocator growing the private pool); after a long test session one of them invalidated the capture of the
        # acting path (cudaErrorStreamCaptureInvalidated, order-dependent).  "relaxed" permits them -- the library never
        # synchronises or touches host memory during a step, which is what actually matters for a capture.
        mode = os.environ.get("BD_GRAPH_CAPTURE_MODE", "relaxed")
        with torch.cuda.graph(self.graph, capture_error_mode=mode):
            self.outputs = fn(*self.static_inputs)

    def __call__(self, *inputs: torch.Tensor):
        for dst, src in zip(self.static_inputs, inputs):
            if src is not dst:
                dst.copy_(src, non_blocking=True)
        self.graph.replay()
        return self.outputs

    def replay(self):
        self.graph.replay()
        return self.outputs
