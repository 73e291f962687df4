"""fp32 means fp32: the reference is evaluated with TF32 off (SURVEY 8c protocol step 1) and the
sampled indices depend on the last bits of the features, so every library GEMM/conv issued by our
blocks runs with TF32 disabled, whatever the process-wide default is."""
from __future__ import annotations

import contextlib
import functools

import torch


@contextlib.contextmanager
def strict_fp32():
    cudnn, mm = torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    try:
        yield
    finally:
        torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32 = cudnn, mm


def fp32_forward(fn):
    @functools.wraps(fn)
    def wrapped(*a, **k):
        with strict_fp32():
            return fn(*a, **k)

    return wrapped
