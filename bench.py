#!/usr/bin/env python
"""Benchmark of the SAMBLE hot path on B200 (contract: README of the task / DESIGN.md section 6).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--batch B] [--points N]
  python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...
  python bench.py --impl reference ...        # the reference's CPU algorithm (oracle port) on host cores

A "step" is one SAMBLE ShapeNetPart segmentation forward (BASELINE config 3: per-GPU batch of B=16
clouds, N=2048 points, k=32, downsample 2048->1024->512, eval mode, frozen bin boundaries, topk
sampling) on synthetic clouds with seeded weights.  `value` = clouds/s with inputs resident in HBM;
`e2e` = the same forward called with HOST (pinned) inputs and a host copy of the logits, copies timed.
The batch is sharded by cloud across ranks with no collective on the path (weak scaling).
One JSON line is printed by rank 0.
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import threading
import time

import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC, UNIT = "clouds/sec SAMBLE seg fwd (N=2048)", "clouds/s"


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--batch", type=int, default=16, help="clouds per GPU per step")
    ap.add_argument("--points", type=int, default=2048)
    ap.add_argument("--impl", default="native", choices=["native", "reference"])
    ap.add_argument("--cpu-sample-batch", type=int, default=4)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-graph", action="store_true", help="launch kernel by kernel instead of replaying a CUDA graph")
    return ap.parse_args()


def workload_config(args, world):
    return {"workload": f"SAMBLE ShapeNetPart seg forward, B={args.batch}/GPU N={args.points} k=32 "
                        f"M=[{args.points // 2},{args.points // 4}] nb=4, eval, frozen boundaries, sample_mode=topk",
            "global_batch": args.batch * world, "points": args.points, "parallelism": f"cloud-sharded x{world}",
            "l2": "256 MiB scratch write between timed steps (L2 flush)"}


# ------------------------------------------------------------------------------- CPU arm


def cpu_reference_rate(batch, points, steps, warmup):
    """The reference's algorithm (oracle port: dense cdist/topk/softmax on ATen CPU) on host cores."""
    from oracle import samble_oracle as O
    from samble_b200 import models
    from samble_b200.config import seg_config
    from samble_b200.testing import fill_state_dict_, synthetic_clouds

    # torchrun exports OMP_NUM_THREADS=1; the CPU arm is entitled to every host core it can use
    torch.set_num_threads(max(1, len(os.sched_getaffinity(0))))
    cfg = seg_config(M=(points // 2, points // 4))
    sd = fill_state_dict_(models.ShapeNetModel(cfg).state_dict(), seed=1, sharpen=4.0)
    x, cat = synthetic_clouds(batch, points, seed=2)
    states = [O.DSState(True), O.DSState(True)]
    times = []
    with torch.no_grad():
        O.seg_forward(sd, cfg, x, cat, states)          # calibration
        for s in states:
            s.dynamic = False
        for i in range(warmup + steps):
            t0 = time.perf_counter()
            O.seg_forward(sd, cfg, x, cat, states)
            if i >= warmup:
                times.append(time.perf_counter() - t0)
    dt = sum(times) / len(times)
    return batch / dt, dt


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    B = args.cpu_sample_batch
    steps, warmup = max(1, min(args.steps, 5)), max(1, min(args.warmup, 1))
    rate, dt = cpu_reference_rate(B, args.points, steps, warmup)
    cores = torch.get_num_threads()
    line = {"impl": "reference", "metric": METRIC, "value": rate, "unit": UNIT, "n_gpus": args.gpus, "steps": steps,
            "warmup": warmup, "ms_per_step": dt * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic", "config": workload_config(args, max(1, args.gpus)),
            "cpu_baseline": {"value": rate, "unit": UNIT, "cores": cores, "kind": "port",
                             "sample": f"seg forward of B={B} clouds N={args.points} per step, {steps} timed steps "
                                       f"(reference is pure Python/ATen; its CPU path = oracle port, bit-exact to it)"},
            "e2e": {"value": rate, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------- clocks


class ClockSampler(threading.Thread):
    def __init__(self, index):
        super().__init__(daemon=True)
        self.samples, self.reasons, self.max_mhz, self.stop_flag = [], set(), None, False
        try:
            import pynvml

            pynvml.nvmlInit()
            self.nv, self.h = pynvml, pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
        except Exception:
            self.nv = None

    NAMES = {0x8: "hw_slowdown", 0x40: "hw_thermal_slowdown", 0x20: "sw_thermal_slowdown", 0x4: "sw_power_cap",
             0x80: "hw_power_brake", 0x2: "applications_clocks_setting", 0x100: "display_clock_setting"}

    def run(self):
        if self.nv is None:
            return
        while not self.stop_flag:
            try:
                self.samples.append(self.nv.nvmlDeviceGetClockInfo(self.h, self.nv.NVML_CLOCK_SM))
                mask = self.nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)
                for bit, name in self.NAMES.items():
                    if mask & bit:
                        self.reasons.add(name)
            except Exception:
                pass
            time.sleep(0.05)

    def result(self):
        s = sorted(self.samples)
        return {"sm_mhz": s[len(s) // 2] if s else None, "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons),
                "samples": len(s)}


# ------------------------------------------------------------------------------- native arm


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            d = json.load(f)
        return d, "measured (MEASURED_PEAKS.json)"
    return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0}, "fallback (B200_PROFILING.md)"


def knn_census(B, N):
    """feature-space kNN calls of one seg forward: (Nq, Nr, C) x count."""
    return [(N, N, 64, 1), (N, N, 128, 3), (N // 2, N // 2, 128, 3), (N // 4, N // 4, 128, 1)]


def linear_census(model, x, cat):
    """Algorithmic bytes / FLOPs of the point-wise linear layers of one forward: every call of ops.linear /
    ops.linear_pool is intercepted once and its operands counted (activations in + weights + activations out
    (+ residual); 2*M*K*Nout FLOPs)."""
    from samble_b200 import ops

    tot = {"bytes": 0.0, "flops": 0.0, "calls": 0}
    real_linear, real_pool = ops.linear, ops.linear_pool

    def lin(xx, w, **kw):
        y = real_linear(xx, w, **kw)
        nout, k = w.shape[0], w[0].numel()
        m = xx.numel() // k
        res = kw.get("residual")
        tot["bytes"] += 4.0 * (xx.numel() + w.numel() + y.numel() + (res.numel() if res is not None else 0))
        tot["flops"] += 2.0 * m * k * nout
        tot["calls"] += 1
        return y

    def pool(xx, w, **kw):
        r = real_pool(xx, w, **kw)
        nout, k = w.shape[0], w[0].numel()
        m = xx.numel() // k
        tot["bytes"] += 4.0 * (xx.numel() + w.numel() + 2 * (m // 32) * nout)     # partial max/sum rows instead of y
        tot["flops"] += 2.0 * m * k * nout
        tot["calls"] += 1
        return r

    ops.linear, ops.linear_pool = lin, pool
    try:
        model(x, cat)
        torch.cuda.synchronize()
    finally:
        ops.linear, ops.linear_pool = real_linear, real_pool
    return tot


def downsample_block_ms(model, x, cat, reps=3):
    """Device time of the DownSampleToken blocks of one forward (their kNN, scoring, sampler and selected-row
    attention): BASELINE's secondary figure 'kNN+sampling us/batch' and the 'kNN+DownSample path' of SURVEY 8d."""
    ds_list = list(model.block.downsample_list)
    spans = []
    originals = [ds.forward for ds in ds_list]

    def wrap(fn):
        def timed(*a, **k):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            out = fn(*a, **k)
            e1.record()
            spans.append((e0, e1))
            return out
        return timed

    for ds, fn in zip(ds_list, originals):
        ds.forward = wrap(fn)
    try:
        for _ in range(reps):
            model(x, cat)
        torch.cuda.synchronize()
    finally:
        for ds in ds_list:
            del ds.forward                       # back to the class method
    return sum(a.elapsed_time(b) for a, b in spans) / reps


def ncu_traffic(kernel):
    """dram__bytes_read.sum + dram__bytes_write.sum per step of `kernel`, from the committed ncu --set full capture
    (profiles/r1_traffic.json, written by tools/summarize_ncu.py traffic); None if that kernel was not captured."""
    p = os.path.join(ROOT, "profiles", "r1_traffic.json")
    if not os.path.exists(p):
        return None
    with open(p) as f:
        return json.load(f).get("dram_bytes_per_step", {}).get(kernel)


def knn_sampling_block(ds_ms, B, N, pk):
    """BASELINE.json's secondary metric and SURVEY 8d's 'kNN + DownSample path': both DownSampleToken blocks of the
    step (feature kNN + q/k/v + QK^T row statistics + edge scores + bins / k / per-bin top-k + selected rows x V).
    Algorithmic FLOPs per cloud from SURVEY 8d: kNN 2*N^2*C plus DS 6*N*D^2 + 2*N*(N+nb)*D + 2*N*K*D + 2*M*(N+nb)*D, for
    N -> N/2 and N/2 -> N/4 (3.68 GFLOP at N=2048).  Peak = tf32 tensor rate (half the measured bf16 rate); the
    kernels spend 3 tf32 MMAs per fp32-class product (linear layers, QK^T) or 2 passes (kNN), stated so the fraction
    can be read either way."""
    def ds_flops(n, m, d=128, nb=4, k=32):
        return 2.0 * n * n * d + 6.0 * n * d * d + 2.0 * n * (n + nb) * d + 2.0 * n * k * d + 2.0 * m * (n + nb) * d

    flops = (ds_flops(N, N // 2) + ds_flops(N // 2, N // 4)) * B
    tf32_peak = pk["bf16_tflops_sustained"] / 2
    ach = flops / (ds_ms * 1e-3) / 1e12
    return {"us_per_batch": ds_ms * 1e3, "clouds": B, "algorithmic_gflop_per_batch": flops / 1e9, "achieved_tflops": ach,
            "tf32_peak_tflops": tf32_peak, "frac_of_tf32_peak": ach / tf32_peak,
            "note": "both DownSampleToken blocks (2048->1024, 1024->512) incl. their feature kNN, eager launches"}


def roofline_of(dom, ms, census, B, N, pk):
    """Roofline block for the dominant kernel family of the step (DESIGN.md section 5 states the per-unit figures)."""
    sec = ms * 1e-3
    if dom.startswith("linear"):
        # fp32 point-wise layers, K <= 1024: 2*K*Nout/(4*(K+Nout)) ~ 50-100 FLOP/B, i.e. memory-side at fp32-class
        # tensor throughput (3 tf32 MMAs per product); reported against HBM, with the tensor-pipe share beside it
        ach = census["bytes"] / sec / 1e9
        t = ncu_traffic(dom)
        return {"bound": "hbm", "achieved": ach, "peak": pk["hbm_gbs"], "unit": "GB/s", "frac": ach / pk["hbm_gbs"],
                "traffic": t, "algorithmic_bytes_per_step": census["bytes"], "calls_per_step": census["calls"],
                "tensor_tflops_3xtf32": 3 * census["flops"] / sec / 1e12,
                "note": "all linear_tma launches of the step; achieved = (X + W + Y [+ residual]) bytes / their summed duration"}
    if dom.startswith("knn_tc") or dom.startswith("knn_select") or dom.startswith("knn_rerank"):
        flops = sum(2.0 * nq * nr * (c + 8) * cnt for nq, nr, c, cnt in knn_census(B, N)) * B
        ach = flops / sec / 1e12
        return {"bound": "tensor", "achieved": ach, "peak": pk["bf16_tflops_sustained"] / 2, "unit": "TFLOP/s",
                "frac": ach / (pk["bf16_tflops_sustained"] / 2), "traffic": ncu_traffic(dom),
                "note": "tf32 distance GEMM of one pass; peak = half the measured bf16 rate (kind::tf32 issues at half the bf16 rate)"}
    return {"bound": "hbm", "achieved": None, "peak": pk["hbm_gbs"], "unit": "GB/s", "frac": None, "traffic": ncu_traffic(dom)}


def run_native(args):
    import torch.distributed as dist

    from samble_b200 import _lib as L
    from samble_b200 import models
    from samble_b200.config import seg_config
    from samble_b200.testing import fill_state_dict_, synthetic_clouds

    rank, world = int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the native path has no CPU fallback")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    torch.backends.cuda.matmul.allow_tf32 = False          # stock GEMMs stay true fp32
    torch.backends.cudnn.allow_tf32 = False
    B, N = args.batch, args.points
    cfg = seg_config(M=(N // 2, N // 4))
    model = models.ShapeNetModel(cfg)
    model.load_state_dict(fill_state_dict_(model.state_dict(), seed=1, sharpen=4.0))
    model = model.eval().to(dev)
    # every rank owns its own shard of the global batch: clouds [rank*B, (rank+1)*B)
    xh, cath = synthetic_clouds(B * world, N, seed=2)
    xh, cath = xh[rank * B:(rank + 1) * B].contiguous().pin_memory(), cath[rank * B:(rank + 1) * B].contiguous().pin_memory()
    x, cat = xh.to(dev), cath.to(dev)
    out_h = torch.empty(B, 50, N, dtype=torch.float32).pin_memory()
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    lib = L.lib()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(step_fn, steps, tail_fn=None):
        evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(steps)]
        barrier()
        lib.samble_reset_launch_count()
        for i, (s, e) in enumerate(evs):
            flush.zero_()                                  # evict L2 between timed steps
            s.record()
            step_fn()
            if tail_fn is not None and i == steps - 1:
                tail_fn()                                  # e.g. the last step's own D2H: nothing escapes the timed regions
            e.record()
        barrier()
        ms = sum(s.elapsed_time(e) for s, e in evs)
        t = torch.tensor([ms], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item()), int(lib.samble_launch_count())

    with torch.no_grad():
        model(x, cat)                                       # calibration batch (dynamic boundaries)
        models.freeze_boundaries(model)
        lib.samble_reset_launch_count()
        model(x, cat)
        launches_per_step = int(lib.samble_launch_count())
        run = model
        if not args.no_graph:
            from samble_b200.runtime import GraphedForward

            run = GraphedForward(model, x, cat)             # whole step = one graph launch

        def step_resident():
            run(x, cat)

        pipe = None
        if not args.no_graph:
            from samble_b200.runtime import HostPipeline

            pipe = HostPipeline(run)       # the serving loop a host-resident caller uses: D2H of step i overlaps step i+1

        def step_e2e():
            # every timed step contains: H2D of its inputs, the forward, and the wait for the PREVIOUS step's D2H (which
            # ran beside this step's compute); the last step also waits for its own (tail_fn) -- all copies are timed
            if pipe is None:
                y = model(xh.to(dev, non_blocking=True), cath.to(dev, non_blocking=True))
                out_h.copy_(y, non_blocking=True)
            else:
                pipe.submit(xh, cath)
                pipe.wait_previous()

        for _ in range(max(3, args.warmup)):
            step_resident()
        sampler = ClockSampler(local)
        sampler.start()
        ms, launches = timed(step_resident, args.steps)
        for _ in range(2):
            step_e2e()
        ms_e2e, _ = timed(step_e2e, args.steps, tail_fn=(pipe.wait_all if pipe is not None else None))
        sampler.stop_flag = True
        sampler.join()

        # per-kernel device time of one profiled forward (CUDA events around every launch of ours)
        L.profile(True)
        torch.cuda.synchronize()
        t0 = torch.cuda.Event(enable_timing=True)
        t1 = torch.cuda.Event(enable_timing=True)
        t0.record()
        for _ in range(3):
            model(x, cat)
        t1.record()
        torch.cuda.synchronize()
        prof = L.profile_report()
        L.profile(False)
        prof_total_ms = t0.elapsed_time(t1) / 3
        census = linear_census(model, x, cat)
        ds_ms = downsample_block_ms(model, x, cat)

    value = B * world * args.steps / (ms / 1e3)
    e2e_value = B * world * args.steps / (ms_e2e / 1e3)
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    pk, pk_src = peaks()
    native_ms = {k: v[1] / 3 for k, v in prof.items()}
    dom = max(native_ms, key=native_ms.get)
    step_ms = ms / args.steps                     # the graph-replayed step the kernels' event-timed durations are set against
    roof = {"kernel": dom, "ms_per_step": native_ms[dom], "share_of_step": native_ms[dom] / step_ms,
            "launches_per_step": prof[dom][0] // 3, "peak_source": pk_src}
    roof.update(roofline_of(dom, native_ms[dom], census, B, N, pk))
    line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": max(3, args.warmup),
            "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic", "config": workload_config(args, world), "clocks": sampler.result(),
            "e2e": {"value": e2e_value, "unit": UNIT, "ms_per_step": ms_e2e / args.steps,
                    "h2d_bytes_per_step": int(xh.numel() * 4 + cath.numel() * 4), "d2h_bytes_per_step": int(out_h.numel() * 4)},
            "gpu_launches": launches_per_step * args.steps, "launch_mode": "eager" if args.no_graph else "cuda_graph",
            "roofline": roof,
            "knn_plus_sampling": knn_sampling_block(ds_ms, B, N, pk),
            "kernels_ms_per_step": {k: round(v, 4) for k, v in sorted(native_ms.items(), key=lambda kv: -kv[1])},
            "native_share_of_step": min(1.0, sum(native_ms.values()) / step_ms)}
    if not args.no_cpu_baseline and world == 1:
        rate, dt = cpu_reference_rate(args.cpu_sample_batch, N, 2, 1)
        line["cpu_baseline"] = {"value": rate, "unit": UNIT, "cores": torch.get_num_threads(), "kind": "port",
                                "sample": f"oracle seg forward, B={args.cpu_sample_batch} N={N}, 2 timed steps after 1 warm-up"}
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    a = parse()
    if a.impl == "reference":
        run_reference(a)
    else:
        run_native(a)
