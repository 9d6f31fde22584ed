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
