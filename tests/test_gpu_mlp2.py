"""Fused two-layer point-wise MLP (csrc/mlp2.cu) against the two `linear` launches it replaces and an fp64 referee.
models/attention.py:187-192 (feed-forward + residual + bn2), models/seg_model.py:205-214 (conv2 -> conv3)."""
import pytest
import torch

from samble_b200 import ops

pytestmark = pytest.mark.gpu


def _case(B, P, K1, Hd, N2, seed, per_cloud_shift1=False, scale1=True, residual=True, res_first=True, lrelu2=False):
    g = torch.Generator().manual_seed(seed)
    r = lambda *s: torch.randn(*s, generator=g)
    x = r(B, P, K1).cuda()
    w1 = (r(Hd, K1) / K1 ** 0.5).cuda()
    w2 = (r(N2, Hd) / Hd ** 0.5).cuda()
    s1 = (1 + 0.1 * r(Hd)).cuda() if scale1 else None
    h1 = (0.1 * r(B, Hd) if per_cloud_shift1 else 0.1 * r(Hd)).cuda()
    s2, h2 = (1 + 0.1 * r(N2)).cuda(), (0.1 * r(N2)).cuda()
    res = r(B, P, N2).cuda() if residual else None
    return x, w1, w2, s1, h1, s2, h2, res, res_first, lrelu2


def _two_launches(x, w1, w2, s1, h1, s2, h2, res, res_first, lrelu2):
    h = ops.linear(x, w1, scale=s1, shift=h1, lrelu=True)
    return ops.linear(h, w2, scale=s2, shift=h2, lrelu=lrelu2, residual=res, residual_first=res_first)


def _fp64(x, w1, w2, s1, h1, s2, h2, res, res_first, lrelu2):
    d = lambda t: None if t is None else t.double()
    x, w1, w2, s1, h1, s2, h2, res = map(d, (x, w1, w2, s1, h1, s2, h2, res))
    h = x @ w1.t()
    if s1 is not None:
        h = h * s1
    h = h + (h1[:, None, :] if h1.dim() == 2 else h1)
    h = torch.where(h > 0, h, 0.2 * h)
    y = h @ w2.t()
    if res is not None and res_first:
        y = y + res
    y = y * s2 + h2
    if lrelu2:
        y = torch.where(y > 0, y, 0.2 * y)
    if res is not None and not res_first:
        y = y + res
    return y


@pytest.mark.parametrize("B,P,K1,Hd", [(2, 2048, 128, 512), (3, 300, 128, 512), (1, 130, 64, 256), (16, 1024, 128, 512), (2, 128, 128, 128),
                                        (2, 10000, 100, 384),      # odd chunk count, several tiles per CTA: the acc1 buffers alternate across tiles
                                        (1, 19100, 36, 128)])     # one chunk per tile, 150 tiles
def test_mlp2_feed_forward_matches_two_linear_launches(B, P, K1, Hd):
    """N2 = 128: the same products as the two launches; the second layer accumulates in one chain instead of 8-K-block
    chains, so the results agree to fp32 rounding (and are identical when Hd <= 256 = one chain either way)."""
    args = _case(B, P, K1, Hd, 128, seed=P + Hd)
    x, w1, w2, s1, h1, s2, h2, res, rf, l2 = args
    y = ops.mlp2(x, w1, w2, scale1=s1, shift1=h1, scale2=s2, shift2=h2, residual=res, residual_first=rf, lrelu2=l2)
    torch.cuda.synchronize()
    ref2 = _two_launches(*args)
    ref64 = _fp64(*args)
    err = (y.double() - ref64).abs().max().item()
    err2 = (ref2.double() - ref64).abs().max().item()
    print(f"{K1}->{Hd}->128 M={B * P}: max|fused - fp64| {err:.2e}, max|two launches - fp64| {err2:.2e}, "
          f"identical: {torch.equal(y, ref2)}")
    assert torch.isfinite(y).all()
    assert err <= 2e-5 * max(1.0, ref64.abs().max().item())
    assert err <= 2.0 * err2 + 1e-6
    if Hd <= 256:
        assert torch.equal(y, ref2)


@pytest.mark.parametrize("B,P,Hd,per_cloud", [(2, 2048, 1024, True), (3, 384, 1024, True), (2, 200, 512, False), (2, 9728, 384, True)])
def test_mlp2_head_256_wide(B, P, Hd, per_cloud):
    """N2 = 256 (seg head conv2 -> conv3): one 32-K-block accumulation chain; fp32-class against fp64."""
    args = _case(B, P, 128, Hd, 256, seed=P, per_cloud_shift1=per_cloud, residual=False, lrelu2=True)
    x, w1, w2, s1, h1, s2, h2, res, rf, l2 = args
    y = ops.mlp2(x, w1, w2, scale1=s1, shift1=h1, scale2=s2, shift2=h2, lrelu2=l2)
    torch.cuda.synchronize()
    ref2 = _two_launches(*args)
    ref64 = _fp64(*args)
    err = (y.double() - ref64).abs().max().item()
    err2 = (ref2.double() - ref64).abs().max().item()
    print(f"128->{Hd}->256 M={B * P}: max|fused - fp64| {err:.2e}, max|two launches - fp64| {err2:.2e}")
    assert torch.isfinite(y).all()
    assert err <= 3e-5 * max(1.0, ref64.abs().max().item())


def test_mlp2_variants_no_scale_residual_last():
    args = _case(2, 512, 128, 512, 128, seed=5, scale1=False, res_first=False, lrelu2=True)
    x, w1, w2, s1, h1, s2, h2, res, rf, l2 = args
    y = ops.mlp2(x, w1, w2, scale1=None, shift1=h1, scale2=s2, shift2=h2, residual=res, residual_first=False, lrelu2=True)
    ref = _fp64(*args)
    assert (y.double() - ref).abs().max().item() <= 2e-5 * max(1.0, ref.abs().max().item())
    w1s, w2s = w1[:256].contiguous(), w2[:, :256].contiguous()
    y = ops.mlp2(x, w1s, w2s, lrelu1=True)                     # no scale / shift at all (the feed-forward as the model calls it)
    h = ops.linear(x, w1s, lrelu=True)
    assert torch.equal(y, ops.linear(h, w2s))
    y = ops.mlp2(x, w1, w2, lrelu1=True)
    ref = torch.nn.functional.leaky_relu(x.double() @ w1.double().t(), 0.2) @ w2.double().t()
    assert (y.double() - ref).abs().max().item() <= 2e-5 * max(1.0, ref.abs().max().item())


def test_mlp2_rejects_unsupported_widths():
    x = torch.randn(1, 128, 128, device="cuda")
    with pytest.raises(RuntimeError):
        ops.mlp2(x, torch.randn(200, 128, device="cuda"), torch.randn(128, 200, device="cuda"))
    with pytest.raises(RuntimeError):
        ops.mlp2(x, torch.randn(256, 128, device="cuda"), torch.randn(64, 256, device="cuda"))
