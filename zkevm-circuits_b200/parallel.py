"""Multi-GPU host layer: one process per GPU (torchrun), `torch.distributed` for the plumbing (NCCL on GPUs, gloo in the
CPU tests).  Only the two exchange steps the path really has use a collective:

  * MSM shards by POINT RANGE: every rank reduces its own (scalars, bases) slice, the 64-byte partial sums are
    all-gathered and added on the host -> identical commitment on every rank.  No point ever crosses NVLink.
  * One huge NTT (k >= 26) shards the DOMAIN: n = P * M; rank r owns the cyclic subsequence x[r + P t]; local size-M
    NTT, twiddle by omega^(r t), ONE all-to-all, size-P transform across the received blocks.  Output layout = "strips":
    rank s holds X[k M + s M/P + t] for k < P, t < M/P.

The arithmetic is injected through an `ops` object so the same plumbing runs on the GPU (DeviceOps: CUDA kernels through
the C ABI) and, in tests/test_parallel_gloo.py, on CPU tensors with a stand-in backend (world_size = 2, gloo).
Independent units (columns, commitments of different columns) need no collective at all: they are dealt round-robin
(`owner_of`) -- that is the weak-scaling mode bench.py measures.
"""
import ctypes
import numpy as np

from .lib import check, load_library

_vp = ctypes.c_void_p


def owner_of(unit, world):
    """round-robin owner of an independent unit (advice column, lookup argument, quotient coset part)."""
    return unit % world


def shard_range(n, rank, world):
    """contiguous point range of `rank` for a point-range sharded MSM."""
    base, rem = divmod(n, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


class Deal:
    """Host mirror of prover.cu's `Deal`: `count` independent units (columns, lookup arguments, coset parts) are cut into `world`
    contiguous blocks of blk = ceil(count / world); rank r computes block r.  Results live in a slab of world * blk unit slots (the
    tail is padding) so that ONE in-place all-gather -- every rank contributes its own block -- completes the slab on every rank."""

    def __init__(self, count, rank, world):
        self.count, self.rank, self.world = count, rank, world
        self.on = world > 1 and count >= 2
        self.blk = (count + world - 1) // world if self.on else count

    def mine(self, i):
        return (not self.on) or i // self.blk == self.rank

    def padded(self):
        return self.world * self.blk if self.on else self.count

    def gather(self, slab, group=None):
        """slab: tensor with padded() rows; this rank's block is filled; after the call every row < count is filled on every rank"""
        import torch
        import torch.distributed as dist
        if not self.on:
            return slab
        parts = [slab[r * self.blk:(r + 1) * self.blk] for r in range(self.world)]   # views into the slab: an in-place all-gather
        dist.all_gather(parts, slab[self.rank * self.blk:(self.rank + 1) * self.blk].clone(), group=group)
        return slab


def g1_sum_affine(points):
    """points: numpy uint64 (m, 8) -> (affine uint64[8], compressed bytes); host-only C-ABI call."""
    pts = np.ascontiguousarray(points, dtype=np.uint64)
    out = np.zeros(8, dtype=np.uint64)
    comp = (ctypes.c_uint8 * 32)()
    check(load_library().zkb_g1_sum_affine_host(_vp(pts.ctypes.data), pts.shape[0], _vp(out.ctypes.data), ctypes.cast(comp, _vp)))
    return out, bytes(comp)


def combine_msm_partials(local_affine, group=None, device="cpu"):
    """all-gather the per-rank partial sums (uint64[8]) and add them; returns (affine, compressed) identical on all ranks."""
    import torch
    import torch.distributed as dist
    world = dist.get_world_size(group)
    mine = torch.from_numpy(np.ascontiguousarray(local_affine).view(np.int64).copy()).to(device)
    bufs = [torch.empty_like(mine) for _ in range(world)]
    dist.all_gather(bufs, mine, group=group)
    pts = np.stack([b.cpu().numpy().view(np.uint64) for b in bufs])
    return g1_sum_affine(pts)


def best_multiexp_sharded(coeffs_shard_dev, bases_shard_dev, group=None):
    """Point-range sharded MSM: this rank's slice of scalars/bases (device tensors) -> the full commitment on every rank."""
    from . import arithmetic as A
    part = A.best_multiexp_dev(coeffs_shard_dev, bases_shard_dev)
    return combine_msm_partials(part.affine, group=group, device=coeffs_shard_dev.device)


# ---------------------------------------------------------------------------------------------------- distributed NTT
class DeviceOps:
    """CUDA backend of ntt_distributed."""

    def __init__(self):
        from . import arithmetic as A, poly
        from .lib import default_context
        self.A, self.poly, self.ctx = A, poly, default_context()

    def pow_omega(self, omega, e):
        """omega^e as uint64[4] (exponentiation by squaring on a 1-element device tensor)."""
        import torch
        A = self.A
        base = torch.from_numpy(np.ascontiguousarray(omega).view(np.int64).reshape(1, 4).copy()).cuda()
        acc = None
        while e:
            if e & 1:
                acc = base if acc is None else A.field_binop_dev(A.FR, A.OP_MUL, acc, base)
            base = A.field_unop_dev(A.FR, A.UOP_SQR, base)
            e >>= 1
        if acc is None:
            one = torch.tensor([[1, 0, 0, 0]], dtype=torch.int64, device="cuda")
            acc = A.field_unop_dev(A.FR, A.UOP_TO_MONT, one)
        return acc.cpu().numpy().view(np.uint64)[0]

    def local_ntt(self, x, omega, log_m):
        return self.A.best_fft_dev(x, omega, log_m)

    def powers(self, base, n, like):
        return self.poly.fr_powers_dev(base, n, device=like.device)

    def mul(self, a, b):
        return self.A.field_binop_dev(self.A.FR, self.A.OP_MUL, a, b)

    def cross(self, blocks, p, omega_p):
        """blocks: (p * len, 4) tensor, p row-blocks -> size-p transform across the blocks."""
        import torch
        from .arithmetic import _cur_stream
        out = torch.empty_like(blocks)
        w = np.ascontiguousarray(omega_p)
        check(self.ctx.lib.zkb_ntt_cross_dev(self.ctx.handle, _vp(blocks.data_ptr()), _vp(out.data_ptr()), p, blocks.shape[0] // p,
                                             _vp(w.ctypes.data), _cur_stream()))
        return out


def ntt_distributed(local, log_n, omega, ops, group=None):
    """One NTT of size 2^log_n spread over the ranks of `group`.
    local: this rank's cyclic subsequence x[rank + P t] as an (M, 4) tensor (M = n / P).  Returns the (M, 4) strip tensor:
    row k * (M/P) + t  holds  X[k M + rank M/P + t]."""
    import torch
    import torch.distributed as dist
    P = dist.get_world_size(group)
    r = dist.get_rank(group)
    log_p = P.bit_length() - 1
    assert 1 << log_p == P, "world size must be a power of two"
    M = 1 << (log_n - log_p)
    assert local.shape[0] == M and M % P == 0
    omega_m = ops.pow_omega(omega, P)                   # order M
    z = ops.local_ntt(local, omega_m, log_n - log_p)    # Z[r, k2]
    if r:
        z = ops.mul(z, ops.powers(ops.pow_omega(omega, r), M, z))   # * omega^(r k2)
    recv = torch.empty_like(z)
    dist.all_to_all_single(recv, z.contiguous(), group=group)       # block s of z goes to rank s; recv block j came from rank j
    omega_p = ops.pow_omega(omega, M)                   # order P
    return ops.cross(recv, P, omega_p)


def cyclic_shard(x_full, rank, world):
    """test / ingest helper: the cyclic subsequence of a natural-order array."""
    return x_full[rank::world]


def strips_to_natural(strips, world):
    """strips: list (by rank) of (M, 4) arrays in the strip layout -> natural-order (n, 4) array."""
    M = strips[0].shape[0]
    blk = M // world
    out = np.empty((M * world, 4), dtype=strips[0].dtype)
    for s, st in enumerate(strips):
        for k in range(world):
            out[k * M + s * blk: k * M + (s + 1) * blk] = st[k * blk:(k + 1) * blk]
    return out


# ---------------------------------------------------------------------------------------------------- C-ABI sharded entry points
def ntt_sharded_dev(local, log_n, omega, direction=0, scale=None, ctx=None):
    """zkb_ntt_fr_sharded_dev: one size-2^log_n transform over the ranks of the context's communicator (Context.init_comm first).
    direction 0: `local` = x[rank + P t] (cyclic) -> strips;  direction 1: strips -> cyclic.  The twiddle + all-to-all run inside the
    transform kernels over NVLink peer memory (ZKB_SHARDED_EXCHANGE=nccl: ncclSend/ncclRecv baseline).  Returns a new tensor."""
    import torch
    from .lib import default_context
    from .arithmetic import _cur_stream
    ctx = ctx or default_context()
    out = torch.empty_like(local)
    w = np.ascontiguousarray(omega, dtype=np.uint64)
    sc = np.ascontiguousarray(scale, dtype=np.uint64) if scale is not None else None
    check(ctx.lib.zkb_ntt_fr_sharded_dev(ctx.handle, _vp(local.data_ptr()), _vp(out.data_ptr()), int(log_n), _vp(w.ctypes.data),
                                         _vp(sc.ctypes.data) if sc is not None else None, int(direction), _cur_stream()))
    return out


def msm_sharded_dev(coeffs_shard, bases_shard, ctx=None):
    """zkb_msm_g1_sharded_dev: point-range sharded MSM -> (affine uint64[8], compressed bytes), identical on every rank."""
    from .lib import default_context
    from .arithmetic import _cur_stream
    ctx = ctx or default_context()
    out = np.zeros(8, dtype=np.uint64)
    comp = (ctypes.c_uint8 * 32)()
    check(ctx.lib.zkb_msm_g1_sharded_dev(ctx.handle, _vp(coeffs_shard.data_ptr()), _vp(bases_shard.data_ptr()), int(coeffs_shard.shape[0]),
                                         _vp(out.ctypes.data), ctypes.cast(comp, _vp), _cur_stream()))
    return out, bytes(comp)
