"""tcgen05.mma with the A operand in tensor memory: correctness against the smem-operand selftest and issue rate.
Run on the GPU box: python tools/probe_tmem_a.py"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from samble_b200 import _lib as L

lib = L.lib()
for K in (32, 128, 256):
    g = torch.Generator().manual_seed(K)
    A, B = torch.randn(128, K, generator=g).cuda(), torch.randn(128, K, generator=g).cuda()
    D = torch.zeros(128, 128, device="cuda")
    L.check(lib.samble_selftest_tc_gemm_ts(L.ptr(A), L.ptr(B), K, L.ptr(D), 0, 1, None, L.stream()), "ts")
    torch.cuda.synchronize()
    tr = lambda t: (t.view(torch.int32) & ~0x1FFF).view(torch.float32)
    ref = tr(A).double() @ tr(B).double().t()
    print(f"K={K}: max|D - tf32trunc ref| = {(D.double() - ref).abs().max().item():.3e}  (|ref| max {ref.abs().max().item():.2f})")
K = 128
A, B = torch.randn(128, K).cuda(), torch.randn(128, K).cuda()
D = torch.zeros(128, 128, device="cuda")
for ctas in (1, 148):
    out = torch.zeros(ctas, dtype=torch.int64, device="cuda")
    iters = 200
    L.check(lib.samble_selftest_tc_gemm_ts(L.ptr(A), L.ptr(B), K, L.ptr(D), iters, ctas, L.ptr(out), L.stream()), "ts")
    torch.cuda.synchronize()
    print(f"TMEM-A tf32 128x128x8: {out.double().mean().item() / (iters * 16):.1f} cycles per MMA on {ctas} CTAs")
    out2 = torch.zeros(ctas, dtype=torch.int64, device="cuda")
    L.check(lib.samble_selftest_mma_rate(128, iters * 4, ctas, L.ptr(out2), L.stream()), "rate")
    torch.cuda.synchronize()
    print(f"smem-A tf32 128x128x8: {out2.double().mean().item() / (iters * 16):.1f} cycles per MMA on {ctas} CTAs")

print("bf16 (kind::f16), A packed two K elements per TMEM column:")
for K in (64, 128):
    g = torch.Generator().manual_seed(K)
    A, B = torch.randn(128, K, generator=g).cuda(), torch.randn(128, K, generator=g).cuda()
    ref = A.bfloat16().double() @ B.bfloat16().double().t()
    for smem_a in (1, 0):
        D = torch.zeros(128, 128, device="cuda")
        L.check(lib.samble_selftest_tc_gemm_ts_bf16(L.ptr(A), L.ptr(B), K, L.ptr(D), 0, 1, None, smem_a, L.stream()), "ts bf16")
        torch.cuda.synchronize()
        print(f"K={K} A in {'smem' if smem_a else 'TMEM'}: max|D - bf16 ref| = {(D.double() - ref).abs().max().item():.3e}  (|ref| max {ref.abs().max().item():.2f})")
K = 128
D = torch.zeros(128, 128, device="cuda")
# the same instruction stream on RANDOM and on CONSTANT operands: the dense bf16 pipe is throttled by the data it multiplies
for label, A, B in (("random data", torch.randn(128, K).cuda(), torch.randn(128, K).cuda()), ("constant data", torch.ones(128, K).cuda(), torch.full((128, K), 0.5).cuda()),
                    ("zeros", torch.zeros(128, K).cuda(), torch.zeros(128, K).cuda())):
    for ctas in (1, 148):
        for smem_a in (1, 0):
            out = torch.zeros(ctas, dtype=torch.int64, device="cuda")
            iters = 400
            L.check(lib.samble_selftest_tc_gemm_ts_bf16(L.ptr(A), L.ptr(B), K, L.ptr(D), iters, ctas, L.ptr(out), smem_a, L.stream()), "ts bf16")
            torch.cuda.synchronize()
            print(f"bf16 128x128x16, {label}, A in {'smem' if smem_a else 'TMEM'}: {out.double().mean().item() / (iters * 8):.1f} cycles per MMA on {ctas} CTAs")
