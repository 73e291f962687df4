import os, sys, torch, time
torch.backends.cuda.matmul.allow_tf32 = False
M,K,N = 32768, 128, 512
a = torch.randn(M,K,device='cuda'); b = torch.randn(K,N,device='cuda')
ref = (a.double()@b.double())
def t(fn, n=20):
    for _ in range(3): fn()
    torch.cuda.synchronize(); s=torch.cuda.Event(enable_timing=True); e=torch.cuda.Event(enable_timing=True); s.record()
    for _ in range(n): fn()
    e.record(); torch.cuda.synchronize(); return s.elapsed_time(e)/n
ms = t(lambda: a@b); y = a@b
print(os.environ.get('CUBLAS_EMULATE_SINGLE_PRECISION'), os.environ.get('CUBLAS_EMULATION_STRATEGY'), 'ms %.4f'%ms, 'TF/s %.1f'%(2*M*K*N/ms/1e9), 'maxerr %.2e'%((y.double()-ref).abs().max().item()))
torch.backends.cuda.matmul.allow_tf32 = True
ms = t(lambda: a@b); y = a@b
print('tf32: ms %.4f TF/s %.1f maxerr %.2e'%(ms, 2*M*K*N/ms/1e9, (y.double()-ref).abs().max().item()))
print(torch.version.cuda, torch.backends.cuda.preferred_blas_library())
try:
    print('fp32_precision attr:', torch.backends.cuda.matmul.fp32_precision)
except Exception as ex: print('no fp32_precision', ex)
