"""Pin the oracle against the committed golden vectors (outputs of the real reference,
tests/golden/make_golden.py).  CPU only.  Same torch build => bit-exact for integer
outputs; float outputs get a tight tolerance so a different host CPU (other BLAS
kernels) does not flake."""
import os

import numpy as np
import pytest
import torch

from oracle import samble_oracle as O
from samble_b200.config import cls_config, seg_config
from samble_b200.testing import fill_state_dict_, knn_parity, synthetic_clouds, synthetic_features
from tests.golden import make_golden as G

GOLD = os.path.join(os.path.dirname(__file__), "golden")
ATOL = 2e-5


@pytest.fixture(scope="module")
def ops_gold():
    return np.load(os.path.join(GOLD, "ops_small.npz"))


def _t(a):
    return torch.from_numpy(np.asarray(a))


@pytest.mark.parametrize("case", G.KNN_CASES, ids=[c[0] for c in G.KNN_CASES])
def test_knn(ops_gold, case):
    name, B, Nq, Nr, C, k, seed = case
    a = synthetic_features(B, Nq, C, seed)
    b = a if name.endswith("self") or name.startswith("tiny") else synthetic_features(B, Nr, C, seed + 100)
    d, i = O.knn(a, b, k)
    rep = knn_parity(i, _t(ops_gold[f"knn.{name}.idx"]), a, b)
    assert rep["unexplained_rows"] == 0, rep
    assert rep["exact_rate"] > 0.999, rep
    np.testing.assert_allclose(d.numpy(), ops_gold[f"knn.{name}.dist"], atol=ATOL, rtol=1e-5)


@pytest.mark.parametrize("case", G.GROUP_CASES, ids=[c[0] for c in G.GROUP_CASES])
def test_group_mask_gather(ops_gold, case):
    name, B, C, N, K, seed = case
    x = synthetic_features(B, C, N, seed)
    for gt in ("neighbor", "diff", "center_neighbor", "center_diff"):
        g, i = O.group(x, K, gt)
        assert np.array_equal(i.numpy(), ops_gold[f"group.{name}.idx"])
        np.testing.assert_array_equal(g.contiguous().numpy(), ops_gold[f"group.{name}.{gt}"])
    m = np.packbits(O.neighbor_mask(x, K).numpy().astype(bool), axis=-1)
    np.testing.assert_array_equal(m, ops_gold[f"mask.{name}"])
    sel = torch.stack([torch.randperm(N, generator=torch.Generator().manual_seed(seed + b))[: N // 2]
                       for b in range(B)]).unsqueeze(1)
    np.testing.assert_array_equal(O.gather_by_idx(x, sel).numpy(), ops_gold[f"gather.{name}"])
    with pytest.raises(ValueError):
        O.group(x, K, "nope")


def test_interpolate_neighbors(ops_gold):
    unk, kn = synthetic_features(2, 3, 96, 41), synthetic_features(2, 3, 40, 42)
    feat = synthetic_features(2, 16, 40, 43)
    nbr, i, d = O.select_neighbors_interpolate(unk, kn, feat, 3)
    np.testing.assert_array_equal(i.numpy(), ops_gold["interp.idx"])
    np.testing.assert_allclose(d.numpy(), ops_gold["interp.d"], atol=ATOL)
    np.testing.assert_array_equal(nbr.contiguous().numpy(), ops_gold["interp.nbr"])


@pytest.mark.parametrize("case", G.KALLOC_CASES, ids=[c[0] for c in G.KALLOC_CASES])
def test_num_points_to_choose(ops_gold, case):
    name, B, nb, M, N, seed = case
    w, cnt = G.kalloc_inputs(B, nb, M, N, seed)
    k = O.calculate_num_points_to_choose(w, cnt, M)
    np.testing.assert_array_equal(k.numpy(), ops_gold[f"kalloc.{name}"])
    assert bool((k.sum(1) == M).all())


def test_bin_partition_and_topk(ops_gold):
    score = torch.rand(3, 1, 200, generator=torch.Generator().manual_seed(51)) * 1e-3
    bnd, mask = O.bin_partition(score, None, True, 0.99, 4)
    np.testing.assert_allclose(bnd[0].numpy(), ops_gold["bin.dyn_init.upper"], rtol=1e-6)
    np.testing.assert_array_equal(mask.numpy(), ops_gold["bin.dyn_init.mask"])
    score2 = torch.rand(3, 1, 200, generator=torch.Generator().manual_seed(52)) * 1e-3
    bnd2, mask2 = O.bin_partition(score2, [t.clone() for t in bnd], True, 0.99, 4)
    np.testing.assert_allclose(bnd2[0].numpy(), ops_gold["bin.dyn_ema.upper"], rtol=1e-6)
    np.testing.assert_array_equal(mask2.numpy(), ops_gold["bin.dyn_ema.mask"])
    _, mask3 = O.bin_partition(score2, bnd, False, 0.99, 4)
    np.testing.assert_array_equal(mask3.numpy(), ops_gold["bin.static.mask"])
    w = torch.rand(3, 4, generator=torch.Generator().manual_seed(53))
    kk = O.calculate_num_points_to_choose(w, mask3.squeeze(1).sum(1), 100)
    np.testing.assert_array_equal(kk.numpy(), ops_gold["bin.static.k"])
    idx = O.generating_downsampled_index(100, score2, mask3, "topk", 0.1, kk)
    np.testing.assert_array_equal(idx.numpy(), ops_gold["bin.static.idx"])
    # stochastic modes (utils/ops.py:507-613): distributions live on the right bins; draws respect bins and counts
    for mode, bt in (("uniform", 0.1), ("random", 0.1), ("random", "mode_1"), ("random", "mode_4")):
        p = O.sampling_probabilities(score2, mask3, mode, bt).reshape(3, 4, 200)
        inbin = mask3.squeeze(1).permute(0, 2, 1)
        nonempty = inbin.sum(-1, keepdim=True) > 0
        assert torch.all((p > 0) == torch.where(nonempty, inbin, torch.ones_like(inbin)))
        if mode == "random":
            assert torch.allclose(p.sum(-1)[nonempty.squeeze(-1)], torch.tensor(1.0), atol=1e-5)
        idx_r = O.generating_downsampled_index(100, score2, mask3, mode, bt, kk, generator=torch.Generator().manual_seed(7))
        assert tuple(idx_r.shape) == (3, 1, 100)
        for b in range(3):
            off = 0
            for j in range(4):
                seg = idx_r[b, 0, off:off + int(kk[b, j])]
                assert len(set(seg.tolist())) == len(seg) and bool(mask3[b, 0, seg, j].all())
                off += int(kk[b, j])
    with pytest.raises(ValueError):
        O.generating_downsampled_index(100, score2, mask3, "bogus", 0.1, kk)


def _block_sd():
    """state_dict names/shapes of the seg model without building the reference: taken from
    our own mirror model (samble_b200.models), which is state_dict-compatible."""
    from samble_b200.models import ShapeNetModel

    cfg = seg_config(M=(G.BLOCK_N // 2, G.BLOCK_N // 4))
    m = ShapeNetModel(cfg)
    return cfg, fill_state_dict_(m.state_dict(), seed=5, sharpen=4.0)


def test_blocks():
    gold = np.load(os.path.join(GOLD, "blocks_small.npz"))
    cfg, sd = _block_sd()
    N = G.BLOCK_N
    x3, x128 = synthetic_features(2, 3, N, 61), synthetic_features(2, 128, N, 62)
    with torch.no_grad():
        np.testing.assert_allclose(O.edgeconv(sd, "block.embedding_list.0.", x3, 32).numpy(), gold["edgeconv0"], atol=ATOL)
        np.testing.assert_allclose(O.edgeconv(sd, "block.embedding_list.1.", synthetic_features(2, 64, N, 63), 32).numpy(),
                                   gold["edgeconv1"], atol=ATOL)
        np.testing.assert_allclose(O.n2p_attention(sd, "block.feature_learning_layer_list.0.", x128, 32).numpy(),
                                   gold["n2p0"], atol=ATOL)
        st = O.DSState(True)
        for tag in ("calib", "frozen"):
            out = O.downsample_token(sd, "block.downsample_list.0.", x128, N // 2, 32, 4, st)
            np.testing.assert_array_equal(out["idx"].numpy(), gold[f"ds0.{tag}.idx"])
            np.testing.assert_array_equal(out["k"].numpy(), gold[f"ds0.{tag}.k"])
            np.testing.assert_allclose(out["score"].numpy(), gold[f"ds0.{tag}.score"], rtol=1e-5, atol=1e-9)
            np.testing.assert_allclose(out["x_ds"].numpy(), gold[f"ds0.{tag}.x_ds"], atol=ATOL)
            np.testing.assert_allclose(out["bin_weights_beforerelu"].numpy(), gold[f"ds0.{tag}.w"], atol=ATOL)
            np.testing.assert_allclose(out["boundaries"][0].numpy(), gold[f"ds0.{tag}.upper"], rtol=1e-6)
            st.dynamic = False
        xyz_up, xyz_dn = synthetic_features(2, 3, N, 64), synthetic_features(2, 3, N // 2, 65)
        dn = synthetic_features(2, 128, N // 2, 66)
        up = O.upsample_interpolation(sd, "block.upsample_list.0.", x128, dn, xyz_up, xyz_dn, 3)
        np.testing.assert_allclose(up.numpy(), gold["upsample0"], atol=ATOL)


@pytest.mark.parametrize("which", ["seg", "cls"])
def test_models(which):
    from samble_b200 import models

    c = G.MODEL_CASES[which]
    gold = np.load(os.path.join(GOLD, f"{which}_small.npz"))
    cfg = (seg_config if which == "seg" else cls_config)(M=c["M"])
    m = (models.ShapeNetModel if which == "seg" else models.ModelNetModel)(cfg)
    sd = fill_state_dict_(m.state_dict(), seed=c["wseed"], sharpen=4.0)
    x, cat = synthetic_clouds(c["B"], c["N"], c["xseed"])
    states = [O.DSState(True), O.DSState(True)]
    with torch.no_grad():
        for tag in ("calib", "frozen"):
            rec = {}
            y = O.seg_forward(sd, cfg, x, cat, states, rec) if which == "seg" else O.cls_forward(sd, cfg, x, states, rec)
            for i in range(2):
                np.testing.assert_array_equal(rec[f"ds{i}"]["idx"].numpy(), gold[f"{tag}.ds{i}.idx"])
                np.testing.assert_array_equal(rec[f"ds{i}"]["k"].numpy(), gold[f"{tag}.ds{i}.k"])
                np.testing.assert_allclose(rec[f"ds{i}"]["boundaries"][0].numpy(), gold[f"{tag}.ds{i}.upper"], rtol=1e-6)
            np.testing.assert_allclose(y.numpy(), gold[f"{tag}.logits"], atol=1e-4)
            for s in states:
                s.dynamic = False
