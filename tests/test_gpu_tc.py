"""tcgen05 plumbing self-test on real hardware (descriptors / swizzle / TMEM mapping)."""
import pytest
import torch

from samble_b200 import _lib as L

pytestmark = pytest.mark.gpu


def _tf32_trunc(t):
    return (t.view(torch.int32) & ~0x1FFF).view(torch.float32)


@pytest.mark.parametrize("K", [32, 64, 128])
def test_tc_gemm_selftest(K):
    g = torch.Generator().manual_seed(K)
    A = torch.randn(128, K, generator=g).cuda()
    B = torch.randn(128, K, generator=g).cuda()
    D = torch.zeros(128, 128, device="cuda")
    L.check(L.lib().samble_selftest_tc_gemm(L.ptr(A), L.ptr(B), K, L.ptr(D), None, None, L.stream()), "selftest")
    torch.cuda.synchronize()
    exact = A.double() @ B.double().t()
    trunc = _tf32_trunc(A).double() @ _tf32_trunc(B).double().t()
    err_exact = (D.double() - exact).abs().max().item()
    err_trunc = (D.double() - trunc).abs().max().item()
    print(f"K={K}: max|D-fp64| = {err_exact:.3e}, max|D-tf32trunc| = {err_trunc:.3e}")
    assert err_exact < 0.05 * (K ** 0.5), "tcgen05 result is not the GEMM: descriptor/swizzle/TMEM mapping is wrong"


def test_tc_gemm_selftest_nonswizzled_extra_k_step():
    """the compact SWIZZLE_NONE [128 x 32 B] slice that carries |b|^2 through the kNN GEMM."""
    g = torch.Generator().manual_seed(1)
    A, B = torch.randn(128, 64, generator=g).cuda(), torch.randn(128, 64, generator=g).cuda()
    Ax, Bx = torch.randn(128, 8, generator=g).cuda(), torch.randn(128, 8, generator=g).cuda()
    D = torch.zeros(128, 128, device="cuda")
    L.check(L.lib().samble_selftest_tc_gemm(L.ptr(A), L.ptr(B), 64, L.ptr(D), L.ptr(Ax), L.ptr(Bx), L.stream()), "selftest")
    torch.cuda.synchronize()
    ref = _tf32_trunc(A).double() @ _tf32_trunc(B).double().t() + _tf32_trunc(Ax).double() @ _tf32_trunc(Bx).double().t()
    err = (D.double() - ref).abs().max().item()
    print(f"ext: max|D-tf32trunc| = {err:.3e}")
    assert err < 1e-3


LIN_CASES = [  # (B, P, K, Nout, x_layout, out_layout, scale, shift, lrelu, residual, res_first)
    (2, 300, 128, 384, "rows", "rows", False, None, False, False, False),
    (2, 2048, 128, 512, "rows", "rows", True, "shared", True, False, False),
    (3, 257, 512, 128, "rows", "bcn", True, "shared", False, True, True),
    (2, 1000, 128, 1024, "bcn", "bcn", True, "cloud", True, False, False),
    (1, 130, 64, 128, "bcn", "rows", False, "shared", False, True, False),
    (2, 64, 1024, 256, "bcn", "bcn", True, "shared", True, False, False),
    (1, 77, 256, 50, "bcn", "bcn", False, None, False, False, False),
    (2, 96, 6, 64, "rows", "rows", False, "shared", False, False, False),
    (1, 128, 100, 200, "rows", "rows", True, None, True, True, False),
    (2, 515, 512, 128, "rows", "bcn", True, "shared", False, "rows", True),      # the N2P feed-forward tail
]


@pytest.mark.parametrize("case", LIN_CASES, ids=[f"B{c[0]}_P{c[1]}_K{c[2]}_N{c[3]}_{c[4]}_{c[5]}" for c in LIN_CASES])
def test_linear_3xtf32_matches_fp64(case):
    from samble_b200 import ops

    B, P, K, Nout, xl, ol, use_scale, shift_kind, lrelu, use_res, res_first = case
    g = torch.Generator().manual_seed(B * 1000 + P + K)
    x_rows = torch.randn(B, P, K, generator=g) * 2
    w = torch.randn(Nout, K, generator=g) / K ** 0.5
    scale = torch.rand(Nout, generator=g) + 0.5 if use_scale else None
    shift = None if shift_kind is None else (torch.randn(Nout, generator=g) if shift_kind == "shared" else torch.randn(B, Nout, generator=g))
    res_rows = torch.randn(B, P, Nout, generator=g) if use_res else None
    ref = x_rows.double() @ w.double().t()
    if use_res and res_first:
        ref = ref + res_rows.double()
    if scale is not None:
        ref = ref * scale.double()
    if shift is not None:
        ref = ref + (shift.double() if shift.dim() == 1 else shift.double().unsqueeze(1))
    if lrelu:
        ref = torch.where(ref > 0, ref, 0.2 * ref)
    if use_res and not res_first:
        ref = ref + res_rows.double()
    x = x_rows.cuda() if xl == "rows" else x_rows.transpose(1, 2).contiguous().cuda()
    rl = use_res if isinstance(use_res, str) else ol
    res = None if res_rows is None else (res_rows.cuda() if rl == "rows" else res_rows.transpose(1, 2).contiguous().cuda())
    y = ops.linear(x, w.cuda(), x_layout=xl, out_layout=ol, scale=None if scale is None else scale.cuda(),
                   shift=None if shift is None else shift.cuda(), lrelu=lrelu, residual=res, residual_first=res_first,
                   residual_layout=rl)
    torch.cuda.synchronize()
    y_rows = y if ol == "rows" else y.transpose(1, 2)
    assert tuple(y_rows.shape) == (B, P, Nout)
    err = (y_rows.double().cpu() - ref).abs().max().item()
    fp32 = (x_rows @ w.t()).double()                    # what a plain fp32 GEMM on the CPU gives, for scale
    print(f"max err vs fp64 {err:.2e}; an fp32 CPU GEMM is off by {(fp32 - x_rows.double() @ w.double().t()).abs().max().item():.2e}")
    assert err < 2e-5 * max(1.0, ref.abs().max().item())


POOL_CASES = [(2, 256, 128, 1024, True, "shared", True), (3, 96, 64, 200, False, None, False), (1, 2048, 128, 1024, True, "cloud", True),
              (4, 32, 36, 50, True, "shared", True), (2, 512, 512, 128, False, "shared", True)]


@pytest.mark.parametrize("case", POOL_CASES, ids=[f"B{c[0]}_P{c[1]}_K{c[2]}_N{c[3]}" for c in POOL_CASES])
def test_linear_pool_matches_fp64(case):
    """samble_linear_pool: max / mean over each cloud's points of the fused layer, activation never stored."""
    from samble_b200 import ops

    B, P, K, Nout, use_scale, shift_kind, lrelu = case
    g = torch.Generator().manual_seed(B * 77 + P + K)
    x = torch.randn(B, P, K, generator=g) * 2
    w = torch.randn(Nout, K, generator=g) / K ** 0.5
    scale = torch.rand(Nout, generator=g) + 0.5 if use_scale else None
    shift = None if shift_kind is None else (torch.randn(Nout, generator=g) if shift_kind == "shared" else torch.randn(B, Nout, generator=g))
    ref = x.double() @ w.double().t()
    if scale is not None:
        ref = ref * scale.double()
    if shift is not None:
        ref = ref + (shift.double() if shift.dim() == 1 else shift.double().unsqueeze(1))
    if lrelu:
        ref = torch.where(ref > 0, ref, 0.2 * ref)
    mx, mean = ops.linear_pool(x.cuda(), w.cuda(), scale=None if scale is None else scale.cuda(),
                               shift=None if shift is None else shift.cuda(), lrelu=lrelu)
    torch.cuda.synchronize()
    tol = 2e-5 * max(1.0, ref.abs().max().item())
    assert (mx.double().cpu() - ref.max(dim=1)[0]).abs().max().item() < tol
    assert (mean.double().cpu() - ref.mean(dim=1)).abs().max().item() < tol
    only_max, none = ops.linear_pool(x.cuda(), w.cuda(), scale=None if scale is None else scale.cuda(),
                                     shift=None if shift is None else shift.cuda(), lrelu=lrelu, want_mean=False)
    assert none is None and torch.equal(only_max, mx)
    # deterministic: fixed reduction order, no atomics
    mx2, mean2 = ops.linear_pool(x.cuda(), w.cuda(), scale=None if scale is None else scale.cuda(),
                                 shift=None if shift is None else shift.cuda(), lrelu=lrelu)
    assert torch.equal(mx, mx2) and torch.equal(mean, mean2)


@pytest.mark.parametrize("case", [(2, 256, 128, 640), (3, 128, 64, 100), (2, 512, 1024, 128)], ids=lambda c: f"B{c[0]}_R{c[1]}_K{c[2]}_N{c[3]}")
def test_cloud_matmul_matches_fp64(case):
    """samble_cloud_matmul: per-cloud weights, and the softmax-row epilogue with known row statistics."""
    from samble_b200 import ops

    B, R, K, Nout = case
    g = torch.Generator().manual_seed(B + R + K)
    x = torch.randn(B, R, K, generator=g)
    w = torch.randn(B, Nout, K, generator=g) / K ** 0.5
    ref = torch.matmul(x.double(), w.double().transpose(1, 2))
    y = ops.cloud_matmul(x.cuda(), w.cuda())
    assert (y.double().cpu() - ref).abs().max().item() < 2e-5 * max(1.0, ref.abs().max().item())
    div = 3.0
    logits = ref / div
    rmax = logits.max(dim=-1)[0]
    rsum = torch.exp(logits - rmax.unsqueeze(-1)).sum(-1)
    p = ops.cloud_matmul(x.cuda(), w.cuda(), row_max=rmax.float().cuda(), row_sum=rsum.float().cuda(), logit_div=div)
    pref = torch.softmax(logits, dim=-1)
    assert (p.double().cpu() - pref).abs().max().item() < 2e-6
    assert (p.double().cpu().sum(-1) - 1).abs().max().item() < 1e-5
