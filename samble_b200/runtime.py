"""CUDA-graph replay of a frozen forward: the whole model step (our kernels + the library GEMMs) becomes
one graph launch, which removes the ~250 per-step kernel-launch latencies that otherwise bound a
B=16 step on the host (SURVEY 7 hard part 7).  Every C-ABI entry point is capture-safe: no allocation,
no synchronisation, all work on the caller's stream."""
from __future__ import annotations

import torch

from . import blocks


class GraphedForward:
    """model must be in eval mode with frozen bin boundaries (models.freeze_boundaries): the dynamic
    boundary update is a stateful, rank-coupled training-time step and is not replayable."""

    def __init__(self, model: torch.nn.Module, *example_inputs: torch.Tensor, warmup: int = 2):
        if model.training:
            raise RuntimeError("GraphedForward needs model.eval()")
        for m in model.modules():
            if isinstance(m, blocks.DownSampleToken) and m.dynamic_boundaries_enable:
                raise RuntimeError("GraphedForward needs frozen boundaries (models.freeze_boundaries)")
        self.model = model
        self.static_in = [t.clone() for t in example_inputs]
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side), torch.no_grad():
            for _ in range(warmup):
                model(*self.static_in)
        torch.cuda.current_stream().wait_stream(side)
        self.graph = torch.cuda.CUDAGraph()
        with torch.no_grad(), torch.cuda.graph(self.graph):
            self.static_out = model(*self.static_in)

    def __call__(self, *inputs: torch.Tensor):
        for dst, src in zip(self.static_in, inputs):
            if dst.data_ptr() != src.data_ptr():
                dst.copy_(src, non_blocking=True)
        self.graph.replay()
        return self.static_out
