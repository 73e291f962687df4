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
