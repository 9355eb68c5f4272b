#!/usr/bin/env python
"""Time the memory-bound row kernels at the ViT-L/16 step's shapes (CUDA events, L2-exceeding working sets rotated).
usage (GPU box): python tools/microbench.py"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from jepa_b200 import kernels as K  # noqa: E402

dev = torch.device("cuda")
BF, F32 = torch.bfloat16, torch.float32


def timeit(fn, iters=20, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


def main():
    for name, T, D in (("target", 50176, 1024), ("ctx", 13056, 1024), ("pred", 76032, 384)):
        x = torch.randn(T, D, device=dev).to(BF)
        dy = torch.randn(T, D, device=dev).to(BF)
        dres = torch.randn(T, D, device=dev).to(BF)
        y, dx = torch.empty_like(x), torch.empty_like(x)
        g, b = torch.ones(D, device=dev), torch.zeros(D, device=dev)
        mean, rstd = torch.empty(T, device=dev), torch.empty(T, device=dev)
        dg, db = torch.zeros(D, device=dev), torch.zeros(D, device=dev)
        t_f = timeit(lambda: K.layernorm_fwd(x, y, g, b, 1e-6, mean, rstd))
        t_b = timeit(lambda: K.layernorm_bwd(dy, x, g, mean, rstd, dres, dx, dg, db))
        by = T * D * 2
        print(f"ln_fwd {name:7s} T={T} D={D}: {t_f * 1e3:7.1f} us  {2 * by / t_f * 1e-6:7.0f} GB/s | "
              f"ln_bwd: {t_b * 1e3:7.1f} us  {4 * by / t_b * 1e-6:7.0f} GB/s")
    for name, T, D in (("ctx_qkv", 13056, 3072), ("ctx_fc1", 13056, 4096), ("pred_fc1", 76032, 1536), ("pred_d", 76032, 384)):
        x = torch.randn(T, D, device=dev).to(BF)
        out = torch.zeros(D, device=dev)
        t = timeit(lambda: K.colsum(x, out))
        print(f"colsum {name:8s} T={T} D={D}: {t * 1e3:7.1f} us  {T * D * 2 / t * 1e-6:7.0f} GB/s")


if __name__ == "__main__":
    main()
