"""Tiny driver for ncu captures: a few 2^24 forward NTTs and 2^20 MSMs through the C ABI (no timing here -- numbers printed under a
profiler are never bench values).  usage: ncu ... python scripts/prof_kernels.py [ntt|msm|both]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch
import zkb200
from zkb200 import arithmetic as A

what = sys.argv[1] if len(sys.argv) > 1 else "both"
ctx = zkb200.default_context(0)
if what in ("ntt", "both"):
    log_n = 24
    w, wi = A.root_of_unity(log_n)
    data = A.random_fr_dev(1 << log_n, 1)
    for _ in range(3):
        A.best_fft_dev(data, w, log_n)
    torch.cuda.synchronize()
if what in ("msm", "both"):
    n = 1 << 20
    g_can = torch.tensor([[1, 0, 0, 0], [2, 0, 0, 0]], dtype=torch.int64, device="cuda")
    gen = A.field_unop_dev(A.FQ, A.UOP_TO_MONT, g_can).cpu().numpy().view(np.uint64).reshape(8)
    bases = A.g1_fixed_base_mul_dev(gen, A.random_fr_dev(n, 7))
    scal = A.random_fr_dev(n, 77)
    for _ in range(3):
        A.best_multiexp_dev(scal, bases)
    torch.cuda.synchronize()
print("done", ctx.launch_count)
