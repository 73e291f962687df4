"""Teacher-forced end-to-end parity harness.  TEST INFRASTRUCTURE ONLY (same rule as samble_oracle.py:
only tests/, __graft_entry__.smoke() and bench.py's CPU legs may import it).

The hot path makes two kinds of discrete decisions -- neighbour sets and sampled point indices -- and fp32
near-ties in them are broken differently by any two implementations (the reference itself changes its sampled
set when only the batch size changes, SURVEY 7 hard part 1).  Comparing two free-running forwards therefore
yields "overlap rates", which pin nothing.  This harness instead:

  1. runs the native model on the GPU and records every kNN result (in call order) and the sampled indices;
  2. runs the CPU oracle with those decisions forced in (oracle.Forcing), so both sides gather the same
     neighbours and the same rows, and every float downstream must agree within plain fp32 tolerance --
     asserted unconditionally;
  3. judges each decision where it was made, in fp64, on the very inputs the native kernel made it from (recorded
     from the GPU; their closeness to the oracle's is what step 2 checks): kNN rows with testing.knn_parity against
     the fp64 neighbour sets (near-tie band), sampled indices with testing.ds_parity (fp64 scores, fp32-class band);
     agreement with the oracle's own free choice is reported as a rate.
"""
from __future__ import annotations

import contextlib
from typing import Dict, List

import torch

from oracle import samble_oracle as O
from samble_b200 import blocks, ops
from samble_b200.testing import ds_parity, ds_scores_fp64, knn_parity


@contextlib.contextmanager
def record_decisions(log: List):
    """Record (signature, idx[, dist]) of every neighbour search the native path issues, in call order.  Every search
    goes through ops._knn_strided (knn, knn_indices, group, select_neighbors, select_neighbors_interpolate) except the
    fused 3-NN interpolation, which is asked separately for its neighbours."""
    real_knn, real_i3r, real_i3s = ops._knn_strided, ops.interpolate3_rows, ops.interpolate3_search

    inputs = getattr(log, "inputs", None)

    def _knn_strided(a, b, k, layout, want_dist, idx_dtype=torch.int64, ordered=True):
        dist, idx = real_knn(a, b, k, layout, want_dist, idx_dtype, ordered)
        pa, pb = (a.detach(), b.detach()) if layout == "bnc" else (a.detach().transpose(1, 2), b.detach().transpose(1, 2))
        entry = ((pa.shape[1], pb.shape[1], pa.shape[2], k), idx.detach().cpu().long())
        if want_dist:
            entry = entry + ((-dist).detach().cpu(),)
        log.append(entry)
        if inputs is not None:
            inputs.append((pa.cpu(), pb.cpu()))
        return dist, idx

    def interpolate3_rows(xyz_up, xyz_sel, feat_rows, out):
        # the fused 3-NN interpolation does not return its neighbours: ask the same search for them
        probe = torch.zeros(xyz_sel.shape[0], 4, xyz_sel.shape[2], device=xyz_sel.device)
        _, idx, dist = ops.interpolate3(xyz_up, xyz_sel, probe, want_idx=True)
        log.append(((xyz_up.shape[2], xyz_sel.shape[2], 3, 3), idx.detach().cpu().long(), dist.detach().cpu()))
        if inputs is not None:
            inputs.append((xyz_up.detach().transpose(1, 2).cpu(), xyz_sel.detach().transpose(1, 2).cpu()))
        return real_i3r(xyz_up, xyz_sel, feat_rows, out)

    def interpolate3_search(xyz_up, xyz_sel):
        probe = torch.zeros(xyz_sel.shape[0], 4, xyz_sel.shape[2], device=xyz_sel.device)
        _, idx, dist = ops.interpolate3(xyz_up, xyz_sel, probe, want_idx=True)
        log.append(((xyz_up.shape[2], xyz_sel.shape[2], 3, 3), idx.detach().cpu().long(), dist.detach().cpu()))
        if inputs is not None:
            inputs.append((xyz_up.detach().transpose(1, 2).cpu(), xyz_sel.detach().transpose(1, 2).cpu()))
        nn_idx, nn_w = real_i3s(xyz_up, xyz_sel)
        assert torch.equal(nn_idx.long(), idx)                     # the split search is the fused kernel's search
        return nn_idx, nn_w

    ops._knn_strided, ops.interpolate3_rows, ops.interpolate3_search = _knn_strided, interpolate3_rows, interpolate3_search
    try:
        yield log
    finally:
        ops._knn_strided, ops.interpolate3_rows, ops.interpolate3_search = real_knn, real_i3r, real_i3s


class _Log(list):
    """the kNN decision log; .inputs holds, per entry, the (a, b) point sets the native kernel searched"""

    def __init__(self):
        super().__init__()
        self.inputs: List = []


def close_frac(x, ref, atol=2e-4, rtol=2e-4) -> float:
    x, ref = x.detach().cpu().double(), ref.detach().cpu().double()
    return float(((x - ref).abs() <= atol + rtol * ref.abs()).double().mean())


def forward_parity(model, sd: Dict[str, torch.Tensor], cfg, x: torch.Tensor, cat=None, *, which: str = "seg",
                   device: str = "cuda:0", ulps: float = 4.0, ulps_oracle: float = 16.0) -> dict:
    """model: samble_b200.models.{ShapeNetModel,ModelNetModel} in eval mode with FROZEN boundaries (calibrated
    before); sd: its state_dict on the CPU; x (B,3,N) / cat (B,16,1) CPU tensors.  Returns the report dict; use
    `assert_report` for the pass/fail rules."""
    ds_list = list(model.block.downsample_list)
    for ds in ds_list:
        if ds.dynamic_boundaries_enable or ds.bin_boundaries is None:
            raise RuntimeError("forward_parity: calibrate one batch and freeze the boundaries first")
    log = _Log()
    ds_in: List = []
    hooks = [ds.register_forward_pre_hook(lambda mod, args: ds_in.append(args[0].detach().cpu())) for ds in ds_list]
    try:
        with torch.no_grad(), record_decisions(log):
            y = model(x.to(device), cat.to(device)) if which == "seg" else model(x.to(device))
    finally:
        for h in hooks:
            h.remove()
    torch.cuda.synchronize()
    mine_knn = list(zip([e[1] for e in log], log.inputs))            # before the oracle consumes the log
    mine_dist = [e[2] if len(e) > 2 else None for e in log]
    mine = [dict(idx=ds.idx.cpu(), mask=ds.bin_points_mask.cpu(), k=ds.k_point_to_choose.cpu(),
                 score=ds.attention_point_score.cpu(), w_raw=ds.bin_weights_beforerelu.cpu(),
                 tok=ds.attention_bins_beforesoftmax.cpu()) for ds in ds_list]
    states = [O.DSState(False, [t.detach().cpu().clone() for t in ds.bin_boundaries]) for ds in ds_list]
    rec: dict = {}
    with torch.no_grad(), O.forcing(O.Forcing(knn_log=log, ds_idx=[m["idx"] for m in mine])) as f:
        y_ref = O.seg_forward(sd, cfg, x, cat, states, rec) if which == "seg" else O.cls_forward(sd, cfg, x, states, rec)
    report = dict(knn=[], ds=[], unforced_knn_calls=len(f.knn_log))
    # ---- neighbour searches: the native choice against the fp64 neighbour sets of the native kernel's OWN inputs
    forced_calls = [c for c in f.knn_seen if c["forced"] is not None]
    report["oracle_knn_calls_without_native_counterpart"] = len(f.knn_seen) - len(forced_calls)
    for call, (mine_idx, (a, b)), fdist in zip(forced_calls, mine_knn, mine_dist):
        assert torch.equal(call["forced"], mine_idx)
        k = mine_idx.shape[-1]
        d = (a.double().pow(2).sum(-1, keepdim=True) + b.double().pow(2).sum(-1).unsqueeze(1)
             - 2 * a.double() @ b.double().transpose(1, 2))
        truth = d.topk(k, dim=-1, largest=False)[1]
        order = d.gather(2, mine_idx).argsort(dim=-1, stable=True)        # (any-order sets: put ours in distance order)
        rep = knn_parity(mine_idx.gather(2, order), truth, a, b)
        rep.update(sig=call["sig"], set_equal_to_oracle_rate=float((mine_idx.sort(-1)[0] == call["idx"].sort(-1)[0]).all(-1).float().mean()),
                   input_max_abs_diff_vs_oracle=float((a - call["a"]).abs().max()))
        if fdist is not None:
            # forced distances: judged here against fp64 in the reference's normalised units (ops.py:23-29).  The GEMM form
            # |a|^2 + |b|^2 - 2ab cancels, so the honest bound is on d^2: a few roundings of |a'|^2 + |b'|^2.
            mu = a.double().mean(1, keepdim=True)
            sg = (a.double() - mu).std(1, keepdim=True).mean(2, keepdim=True)
            an, bn = (a.double() - mu) / sg, (b.double() - mu) / sg
            bsel = torch.gather(bn.unsqueeze(1).expand(-1, a.shape[1], -1, -1), 2, mine_idx.unsqueeze(-1).expand(-1, -1, -1, 3))
            d2 = (an.unsqueeze(2) - bsel).pow(2).sum(-1)
            mag = an.pow(2).sum(-1, keepdim=True) + bsel.pow(2).sum(-1)
            rep["forced_dist_noise_ulps"] = float(((fdist.double().pow(2) - d2).abs() / (2.0 ** -24 * mag)).max())
            rep["forced_dist_max_diff_vs_oracle"] = float((fdist + call["dist"]).abs().max()) if torch.equal(call["idx"], mine_idx) else None
        del d
        report["knn"].append(rep)
    # ---- sampled indices
    for i, (ds, m) in enumerate(zip(ds_list, mine)):
        r = rec[f"ds{i}"]
        pre = f"block.downsample_list.{i}."
        x_mine = ds_in[i]                          # the DownSample input the native kernels saw
        C = x_mine.shape[1]
        # the DownSample's own kNN is the recorded search whose input is this layer's input
        knn_idx = next(idx for idx, (a, _) in mine_knn if a.shape[1] == x_mine.shape[2] and a.shape[2] == C
                       and torch.equal(a, x_mine.transpose(1, 2)))
        s64, amp = ds_scores_fp64(x_mine, sd[pre + "q_conv.weight"].view(C, C), sd[pre + "k_conv.weight"].view(C, C),
                                  sd[pre + "bin_tokens"][0], knn_idx)
        rep_in = close_frac(x_mine, r["x_in"])
        cuts = states[i].boundaries[0].reshape(-1)[1:]
        rep = ds_parity(s64, amp, cuts, m["idx"], m["mask"], m["k"], ulps=ulps)
        B, _, M = m["idx"].shape
        own = r["own_idx"]                        # what the oracle chose from ITS fp32 scores on the same inputs / neighbours
        rep["oracle_exact_rate"] = float((m["idx"] == own).float().mean())
        rep["oracle_set_overlap"] = float(sum(len(set(m["idx"][b, 0].tolist()) & set(own[b, 0].tolist())) for b in range(B)) / (B * M))
        rep["oracle_bin_mismatch"] = int((m["mask"] != r["mask"]).any(-1).sum())
        rep["k_equal"] = bool(torch.equal(m["k"].long(), r["k"].long()))
        rep["k_max_diff"] = int((m["k"].long() - r["k"].long()).abs().max())
        # k per bin is an exact function of (bin weights, bin sizes): the reference's own routine on the NATIVE bins
        counts = m["mask"][:, 0].sum(1)
        rep["k_consistent"] = bool(torch.equal(O.calculate_num_points_to_choose(torch.relu(m["w_raw"]), counts, M).long(), m["k"].long()))
        tl = r["token_logits"][:, 0].double()                       # (B,N,nb); bin weight = mean token logit over the bin
        w64 = (tl * m["mask"][:, 0]).sum(1) / (counts + 1e-8)
        rep["bin_weight_max_err"] = float((m["w_raw"].double() - w64).abs().max())
        # the oracle, judged by the same referee on ITS inputs: its decision must be explainable too (sanity of the band)
        s64o, ampo = ds_scores_fp64(r["x_in"], sd[pre + "q_conv.weight"].view(C, C), sd[pre + "k_conv.weight"].view(C, C),
                                    sd[pre + "bin_tokens"][0], knn_idx)
        rep["oracle_vs_fp64"] = ds_parity(s64o, ampo, cuts, own, r["mask"], r["k"], ulps=ulps_oracle)
        big = s64.unsqueeze(1) > 1e-30                 # (below that the fp32 probabilities underflow)
        rel = ((m["score"].double() - s64.unsqueeze(1)).abs() / s64.unsqueeze(1).abs().clamp_min(1e-300))[big]
        rel_o = ((r["score"].double() - s64.unsqueeze(1)).abs() / s64.unsqueeze(1).abs().clamp_min(1e-300))[big]
        rep["score_rel_err_vs_fp64"] = dict(native_max=float(rel.max()), native_median=float(rel.median()),
                                            oracle_max=float(rel_o.max()), oracle_median=float(rel_o.median()))
        rep["token_logits_close"] = close_frac(m["tok"], r["token_logits"])
        rep["input_close"] = rep_in
        report["ds"].append(rep)
    report["logits_close_frac"] = close_frac(y, y_ref)
    report["logits_max_abs_err"] = float((y.cpu().double() - y_ref.double()).abs().max())
    report["logits_scale"] = float(y_ref.abs().max())
    report["finite"] = bool(torch.isfinite(y).all())
    return report


def gradient_parity(model, sd: Dict[str, torch.Tensor], cfg, x: torch.Tensor, cat=None, *, which: str = "seg",
                    device: str = "cuda:0", seed: int = 0) -> dict:
    """one backward pass of the native differentiable path against the oracle's ATen autograd.

    The native model runs with gradients enabled (its own mode: eval -> running-statistics BatchNorm, train -> batch
    statistics, mirrored in the oracle via O.bn_training); its neighbour sets, 3-NN distances and sampled indices are
    forced into the oracle (decisions carry no gradient on either side); the scalar loss is sum(logits * probe) with a
    fixed random probe.  Returns per-tensor errors relative to the largest reference gradient entry of that tensor."""
    ds_list = list(model.block.downsample_list)
    for ds in ds_list:
        if ds.dynamic_boundaries_enable or ds.bin_boundaries is None:
            raise RuntimeError("gradient_parity: calibrate one batch and freeze the boundaries first")
    from samble_b200._precision import strict_fp32

    import torch.nn.functional as F

    log = _Log()
    masks: List[torch.Tensor] = []
    real_lrelu = F.leaky_relu

    def recording_lrelu(inp, negative_slope=0.01, inplace=False):
        masks.append((inp.detach() > 0).cpu())
        return real_lrelu(inp, negative_slope, inplace)

    model.zero_grad(set_to_none=True)
    xg = x.to(device).requires_grad_(True)
    F.leaky_relu = recording_lrelu
    try:
        with record_decisions(log):
            y = model(xg, cat.to(device)) if which == "seg" else model(xg)
    finally:
        F.leaky_relu = real_lrelu
    y = y[0] if isinstance(y, tuple) else y
    probe = torch.randn(y.shape, generator=torch.Generator().manual_seed(seed))
    with strict_fp32():          # cuDNN's default lets convolution BACKWARD passes use TF32; the comparison is fp32 vs fp32
        (y * probe.to(device)).sum().backward()
    torch.cuda.synchronize()
    mine = {n: p.grad.detach().cpu() for n, p in model.named_parameters() if p.grad is not None}
    no_grad_params = [n for n, p in model.named_parameters() if p.requires_grad and p.grad is None]
    mine_x = xg.grad.detach().cpu()
    ds_idx = [ds.idx.cpu() for ds in ds_list]

    def oracle_grads(dtype):
        """the oracle's autograd in fp32 (the reference's arithmetic) or fp64 (the referee), same forced decisions"""
        states = [O.DSState(False, [t.detach().cpu().to(dtype).clone() for t in ds.bin_boundaries]) for ds in ds_list]
        sdg = {k: (v.detach().to(dtype).requires_grad_(True) if (v.is_floating_point() and "running_" not in k)
                   else (v.detach().to(dtype) if v.is_floating_point() else v.detach().clone())) for k, v in sd.items()}
        xr = x.detach().to(dtype).requires_grad_(True)
        klog = [(e[0], e[1]) + ((e[2].to(dtype),) if len(e) > 2 else ()) for e in log]
        with O.forcing(O.Forcing(knn_log=klog, ds_idx=ds_idx, keep_inputs=False, lrelu_masks=masks)) as f, O.bn_training(model.training):
            y_ref = (O.seg_forward(sdg, cfg, xr, cat.to(dtype), states) if which == "seg" else O.cls_forward(sdg, cfg, xr, states))
        (y_ref * probe.to(dtype)).sum().backward()
        grads = {n: r.grad for n, r in sdg.items() if isinstance(r, torch.Tensor) and r.requires_grad and r.grad is not None}
        return y_ref.detach(), grads, xr.grad, len(f.knn_log), len(f.lrelu_masks), f.lrelu_flips

    y32, g32, x32, unforced, masks_left, flips32 = oracle_grads(torch.float32)
    y64, g64, x64, _, _, flips64 = oracle_grads(torch.float64)
    rep = dict(unforced_knn_calls=unforced, logits_close_frac=close_frac(y, y32), logits_close_frac_fp64=close_frac(y, y64),
               params={}, missing=no_grad_params, lrelu_calls=len(masks), lrelu_masks_unconsumed=masks_left,
               lrelu_elements=int(sum(m.numel() for m in masks)), lrelu_flips_vs_fp32=flips32, lrelu_flips_vs_fp64=flips64)
    # per tensor: error against the fp64 referee, in units of the largest gradient entry of the WHOLE model for that kind
    # of tensor would hide small tensors, and a tensor's own largest entry is meaningless where the true gradient is zero
    # (a bias in front of a batch-statistics BatchNorm) -- so both are reported: `rel` (own scale) and `abs`.
    scale_all = max(float(v.abs().max()) for v in g64.values())

    def errs(g, r64, r32):
        e_native = float((g.double() - r64).abs().max())
        e_oracle = float((r32.double() - r64).abs().max())
        own = float(r64.abs().max())
        return dict(native=e_native, oracle32=e_oracle, own_scale=own, rel=e_native / max(own, 1e-30),
                    rel_oracle32=e_oracle / max(own, 1e-30))

    for n, r in g64.items():
        if n not in mine:
            rep["missing"].append(n)
            continue
        rep["params"][n] = errs(mine[n], r, g32[n])
    rep["input"] = errs(mine_x, x64, x32)
    rep["model_grad_scale"] = scale_all
    rep["n_params"] = len(rep["params"])
    return rep


def assert_gradient_report(rep: dict, tol: float = 1e-4, noise: float = 8.0, floor: float = 1e-6) -> None:
    """Every gradient tensor agrees with the fp64 referee within `tol` of its own largest entry, OR within `noise` times the
    error the reference's own fp32 autograd makes on that tensor, OR -- tensors whose true gradient vanishes (a bias feeding a
    batch-statistics BatchNorm) -- within `floor` of the model's largest gradient entry."""
    assert rep["unforced_knn_calls"] == 0 and rep["lrelu_masks_unconsumed"] == 0
    assert rep["lrelu_flips_vs_fp64"] <= 1e-5 * rep["lrelu_elements"] + 4, "forced LeakyReLU sides must be near-zero pre-activations only"
    assert not rep["missing"], rep["missing"]
    bad = {}
    for n, e in list(rep["params"].items()) + [("<input>", rep["input"])]:
        ok = e["native"] <= tol * e["own_scale"] or e["native"] <= noise * e["oracle32"] or e["native"] <= floor * rep["model_grad_scale"]
        if not ok:
            bad[n] = e
    assert not bad, bad


def assert_report(rep: dict, *, logits_frac: float = 1.0, knn_exact: float = 0.995) -> None:
    """The pass/fail rules: nothing unexplained anywhere, floats within 2e-4 + 2e-4*|ref| everywhere."""
    assert rep["finite"]
    assert rep["unforced_knn_calls"] == 0, f"{rep['unforced_knn_calls']} recorded neighbour searches were never consumed by the oracle"
    assert rep["oracle_knn_calls_without_native_counterpart"] == 0
    for k in rep["knn"]:
        assert k["unexplained_rows"] == 0, k
        assert k.get("forced_dist_noise_ulps", 0.0) <= 16.0, k
        assert k["exact_rate"] >= knn_exact, k
    for i, d in enumerate(rep["ds"]):
        assert d["unexplained_bin_flips"] == 0 and d["unexplained_topk_swaps"] == 0, (i, d)
        assert d["chosen_outside_bin"] == 0 and d["duplicate_rows"] == 0, (i, d)
        assert d["k_consistent"] and d["bin_weight_max_err"] <= 1e-4, (i, d)
        o = d["oracle_vs_fp64"]
        assert o["unexplained_bin_flips"] == 0 and o["unexplained_topk_swaps"] == 0, ("band too tight even for the oracle", i, o)
        assert d["token_logits_close"] == 1.0, (i, d)
    assert rep["logits_close_frac"] >= logits_frac, (rep["logits_close_frac"], rep["logits_max_abs_err"], rep["logits_scale"])


def brief(rep: dict) -> str:
    knn = rep["knn"]
    parts = [f"kNN calls {len(knn)}: exact-rate min {min(k['exact_rate'] for k in knn):.5f}, unexplained rows "
             f"{sum(k['unexplained_rows'] for k in knn)}"]
    for i, d in enumerate(rep["ds"]):
        parts.append(f"DS{i}: vs oracle exact {d['oracle_exact_rate']:.4f} / set overlap {d['oracle_set_overlap']:.4f}, "
                     f"fp64 referee: bin flips {d['bin_flips']} (unexplained {d['unexplained_bin_flips']}), top-k swaps "
                     f"{d['topk_swaps']} (unexplained {d['unexplained_topk_swaps']}), k equal {d['k_equal']}, score rel err "
                     f"native {d['score_rel_err_vs_fp64']['native_max']:.1e} / oracle {d['score_rel_err_vs_fp64']['oracle_max']:.1e}")
    parts.append(f"logits within 2e-4+2e-4|ref|: {rep['logits_close_frac']:.6f} (max abs err {rep['logits_max_abs_err']:.2e} "
                 f"on scale {rep['logits_scale']:.1f})")
    return "; ".join(parts)
