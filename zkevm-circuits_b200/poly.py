"""Host mirror of the polynomial helpers of halo2_proofs::arithmetic used by the prover (eval_polynomial,
kate_division) plus the scans behind the permutation / lookup arguments.  Device tensors: torch int64 (n,4)."""
import ctypes
import numpy as np

from .lib import check, default_context
from .arithmetic import _vp, _limbs_ptr, _cur_stream


def fr_powers_dev(base, n, device="cuda", ctx=None):
    import torch
    ctx = ctx or default_context()
    out = torch.empty((n, 4), dtype=torch.int64, device=device)
    bk, bp = _limbs_ptr(base)
    check(ctx.lib.zkb_fr_powers_dev(ctx.handle, bp, n, _vp(out.data_ptr()), _cur_stream()))
    return out


def eval_polynomial_dev(polys, x, ctx=None):
    """polys: list of device tensors with the same length; returns numpy uint64 (len(polys), 4)."""
    ctx = ctx or default_context(polys[0].device.index)
    n = polys[0].shape[0]
    ptrs = (ctypes.c_void_p * len(polys))(*[p.data_ptr() for p in polys])
    out = np.zeros((len(polys), 4), dtype=np.uint64)
    xk, xp = _limbs_ptr(x)
    check(ctx.lib.zkb_poly_eval_dev(ctx.handle, ctypes.cast(ptrs, _vp), len(polys), n, xp, _vp(out.ctypes.data), _cur_stream()))
    return out


def prefix_product_dev(a, init, ctx=None):
    import torch
    ctx = ctx or default_context(a.device.index)
    out = torch.empty_like(a)
    ik, ip = _limbs_ptr(init)
    check(ctx.lib.zkb_fr_prefix_product_dev(ctx.handle, _vp(a.data_ptr()), a.shape[0], ip, _vp(out.data_ptr()), _cur_stream()))
    return out


def prefix_sum_dev(a, init, ctx=None):
    import torch
    ctx = ctx or default_context(a.device.index)
    out = torch.empty_like(a)
    ik, ip = _limbs_ptr(init)
    check(ctx.lib.zkb_fr_prefix_sum_dev(ctx.handle, _vp(a.data_ptr()), a.shape[0], ip, _vp(out.data_ptr()), _cur_stream()))
    return out


def kate_division_dev(a, u, ctx=None):
    import torch
    ctx = ctx or default_context(a.device.index)
    out = torch.empty_like(a)
    uk, up = _limbs_ptr(u)
    check(ctx.lib.zkb_kate_division_dev(ctx.handle, _vp(a.data_ptr()), a.shape[0], up, _vp(out.data_ptr()), _cur_stream()))
    return out


def ntt_batch_dev(cols, omega, log_n, scale=None, coset_zeta=0, ctx=None):
    """In-place best_fft of several device columns (list of (n,4) int64 CUDA tensors) in one launch per pass; returns the list."""
    import ctypes
    import numpy as np
    from .arithmetic import _cur_stream, _limbs_ptr
    ctx = ctx or default_context(cols[0].device.index)
    ptrs = (ctypes.c_void_p * len(cols))(*[c.data_ptr() for c in cols])
    kw, wp = _limbs_ptr(omega)
    ks, sp = _limbs_ptr(scale) if scale is not None else (None, None)
    check(ctx.lib.zkb_ntt_fr_batch_dev(ctx.handle, ctypes.cast(ptrs, ctypes.c_void_p), len(cols), log_n, wp, sp, coset_zeta, _cur_stream()))
    return cols
