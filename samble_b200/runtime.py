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


class HostPipeline:
    """Serving loop around a GraphedForward for callers whose clouds and results live in (pinned) host memory.

    submit(x_host, ...) enqueues, on the current stream: the H2D copy of the inputs, the graph replay, and a
    device-to-device copy of the result into one of `depth` staging buffers; the D2H copy of that buffer runs on a
    side stream, so it overlaps the next step's compute.  The returned (host_tensor, event) pair is valid once the
    event has completed; wait_previous() / wait_all() make the current stream wait for earlier results (what a
    latency-bound caller, or a benchmark that must account for every copy, does)."""

    def __init__(self, graphed: GraphedForward, depth: int = 2):
        self.g = graphed
        self.depth = depth
        self.copy_stream = torch.cuda.Stream()
        out = graphed.static_out
        self.stage = [torch.empty_like(out) for _ in range(depth)]
        self.host = [torch.empty(out.shape, dtype=out.dtype).pin_memory() for _ in range(depth)]
        self.done = [torch.cuda.Event() for _ in range(depth)]
        self.count = 0
        for e in self.done:
            e.record()

    def submit(self, *host_inputs: torch.Tensor):
        s = self.count % self.depth
        main = torch.cuda.current_stream()
        for dst, src in zip(self.g.static_in, host_inputs):
            dst.copy_(src, non_blocking=True)
        self.g.graph.replay()
        main.wait_event(self.done[s])                       # staging buffer s: its previous D2H has finished
        self.stage[s].copy_(self.g.static_out, non_blocking=True)
        ready = torch.cuda.Event()
        ready.record(main)
        with torch.cuda.stream(self.copy_stream):
            self.copy_stream.wait_event(ready)
            self.host[s].copy_(self.stage[s], non_blocking=True)
            self.done[s].record(self.copy_stream)
        self.count += 1
        return self.host[s], self.done[s]

    def wait_previous(self):
        """current stream waits for the result of the submit BEFORE the latest one"""
        if self.count >= 2:
            torch.cuda.current_stream().wait_event(self.done[(self.count - 2) % self.depth])

    def wait_all(self):
        torch.cuda.current_stream().wait_stream(self.copy_stream)

