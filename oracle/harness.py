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
  3. judges each decision where it was made, on the oracle's own inputs of that stage: kNN rows with
     testing.knn_parity (fp64, near-tie band), sampled indices with testing.ds_parity (fp64 scores, fp32-class
     band) and against the oracle's own choice (exact-match rate).
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
    """Record (signature, idx) of every neighbour search the native blocks issue, in call order."""
    real_knn, real_i3r = ops.knn_indices, ops.interpolate3_rows

    def knn_indices(pcd, K, idx_dtype=torch.int32, ordered=True):
        idx = real_knn(pcd, K, idx_dtype, ordered)
        log.append(((pcd.shape[2], pcd.shape[2], pcd.shape[1], K), idx.detach().cpu().long()))
        return idx

    def interpolate3_rows(xyz_up, xyz_sel, feat_rows, out):
        # the fused 3-NN interpolation does not return its neighbours: ask the same search for them
        probe = torch.zeros(xyz_sel.shape[0], 4, xyz_sel.shape[2], device=xyz_sel.device)
        _, idx, _ = ops.interpolate3(xyz_up, xyz_sel, probe, want_idx=True)
        log.append(((xyz_up.shape[2], xyz_sel.shape[2], 3, 3), idx.detach().cpu().long()))
        return real_i3r(xyz_up, xyz_sel, feat_rows, out)

    ops.knn_indices, ops.interpolate3_rows = knn_indices, interpolate3_rows
    try:
        yield log
    finally:
        ops.knn_indices, ops.interpolate3_rows = real_knn, real_i3r


def close_frac(x, ref, atol=2e-4, rtol=2e-4) -> float:
    x, ref = x.detach().cpu().double(), ref.detach().cpu().double()
    return float(((x - ref).abs() <= atol + rtol * ref.abs()).double().mean())


def forward_parity(model, sd: Dict[str, torch.Tensor], cfg, x: torch.Tensor, cat=None, *, which: str = "seg",
                   device: str = "cuda:0", ulps: float = 64.0) -> dict:
    """model: samble_b200.models.{ShapeNetModel,ModelNetModel} in eval mode with FROZEN boundaries (calibrated
    before); sd: its state_dict on the CPU; x (B,3,N) / cat (B,16,1) CPU tensors.  Returns the report dict; use
    `assert_report` for the pass/fail rules."""
    ds_list = list(model.block.downsample_list)
    for ds in ds_list:
        if ds.dynamic_boundaries_enable or ds.bin_boundaries is None:
            raise RuntimeError("forward_parity: calibrate one batch and freeze the boundaries first")
    log: List = []
    with torch.no_grad(), record_decisions(log):
        y = model(x.to(device), cat.to(device)) if which == "seg" else model(x.to(device))
    torch.cuda.synchronize()
    mine = [dict(idx=ds.idx.cpu(), mask=ds.bin_points_mask.cpu(), k=ds.k_point_to_choose.cpu(),
                 score=ds.attention_point_score.cpu(), w_raw=ds.bin_weights_beforerelu.cpu(),
                 tok=ds.attention_bins_beforesoftmax.cpu()) for ds in ds_list]
    states = [O.DSState(False, [t.detach().cpu().clone() for t in ds.bin_boundaries]) for ds in ds_list]
    rec: dict = {}
    with torch.no_grad(), O.forcing(O.Forcing(knn_log=log, ds_idx=[m["idx"] for m in mine])) as f:
        y_ref = O.seg_forward(sd, cfg, x, cat, states, rec) if which == "seg" else O.cls_forward(sd, cfg, x, states, rec)
    report = dict(knn=[], ds=[], unforced_knn_calls=len(f.knn_log))
    # ---- neighbour searches, each on the oracle's own inputs of that stage
    for call in f.knn_seen:
        if call["forced"] is None:
            report["knn"].append(dict(sig=call["sig"], forced=False))
            continue
        nq, nr, c, k = call["sig"]
        mine_idx, own = call["forced"], call["idx"]
        # any-order neighbour SETS: compare sorted by the oracle's distance order => position-wise after set alignment
        a, b = call["a"], call["b"]
        d = torch.cdist(a.double(), b.double())
        order = d.gather(2, mine_idx).argsort(dim=-1, stable=True)
        rep = knn_parity(mine_idx.gather(2, order), own, a, b)
        rep.update(sig=call["sig"], forced=True,
                   set_equal_rate=float((mine_idx.sort(-1)[0] == own.sort(-1)[0]).all(-1).float().mean()))
        report["knn"].append(rep)
    # ---- sampled indices
    for i, (ds, m) in enumerate(zip(ds_list, mine)):
        r = rec[f"ds{i}"]
        pre = f"block.downsample_list.{i}."
        C = r["x_in"].shape[1]
        # the DownSample's own kNN is the call whose inputs are this layer's input
        knn_idx = next(c["forced"] for c in f.knn_seen if c["forced"] is not None and c["a"].shape[1] == r["x_in"].shape[2]
                       and c["a"].shape[2] == C and torch.equal(c["a"], r["x_in"].transpose(1, 2)))
        s64, amp = ds_scores_fp64(r["x_in"], sd[pre + "q_conv.weight"].view(C, C), sd[pre + "k_conv.weight"].view(C, C),
                                  sd[pre + "bin_tokens"][0], knn_idx)
        cuts = states[i].boundaries[0].reshape(-1)[1:]
        rep = ds_parity(s64, amp, cuts, m["idx"], m["mask"], m["k"], ulps=ulps)
        B, _, M = m["idx"].shape
        own = r["own_idx"]                        # what the oracle chose from ITS fp32 scores on the same inputs / neighbours
        rep["oracle_exact_rate"] = float((m["idx"] == own).float().mean())
        rep["oracle_set_overlap"] = float(sum(len(set(m["idx"][b, 0].tolist()) & set(own[b, 0].tolist())) for b in range(B)) / (B * M))
        rep["oracle_bin_mismatch"] = int((m["mask"] != r["mask"]).any(-1).sum())
        rep["k_equal"] = bool(torch.equal(m["k"].long(), r["k"].long()))
        rep["k_max_diff"] = int((m["k"].long() - r["k"].long()).abs().max())
        # the oracle, judged by the same referee: its own decision must be explainable too (sanity of the band)
        rep["oracle_vs_fp64"] = ds_parity(s64, amp, cuts, own, r["mask"], r["k"], ulps=ulps)
        rel = ((m["score"].double() - s64.unsqueeze(1)).abs() / s64.unsqueeze(1).abs().clamp_min(1e-300))
        rel_o = ((r["score"].double() - s64.unsqueeze(1)).abs() / s64.unsqueeze(1).abs().clamp_min(1e-300))
        rep["score_rel_err_vs_fp64"] = dict(native_max=float(rel.max()), native_median=float(rel.median()),
                                            oracle_max=float(rel_o.max()), oracle_median=float(rel_o.median()))
        rep["token_logits_close"] = close_frac(m["tok"], r["token_logits"])
        report["ds"].append(rep)
    report["logits_close_frac"] = close_frac(y, y_ref)
    report["logits_max_abs_err"] = float((y.cpu().double() - y_ref.double()).abs().max())
    report["logits_scale"] = float(y_ref.abs().max())
    report["finite"] = bool(torch.isfinite(y).all())
    return report


def assert_report(rep: dict, *, logits_frac: float = 1.0, knn_exact: float = 0.995) -> None:
    """The pass/fail rules: nothing unexplained anywhere, floats within 2e-4 + 2e-4*|ref| everywhere."""
    assert rep["finite"]
    assert rep["unforced_knn_calls"] == 0, f"{rep['unforced_knn_calls']} recorded neighbour searches were never consumed by the oracle"
    for k in rep["knn"]:
        assert k["forced"], f"oracle kNN call {k['sig']} had no native counterpart"
        assert k["unexplained_rows"] == 0, k
        assert k["exact_rate"] >= knn_exact, k
    for i, d in enumerate(rep["ds"]):
        assert d["unexplained_bin_flips"] == 0 and d["unexplained_topk_swaps"] == 0, (i, d)
        assert d["chosen_outside_bin"] == 0 and d["duplicate_rows"] == 0, (i, d)
        assert d["k_max_diff"] <= 1, (i, d)
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
