"""CPU, world_size = 2, gloo: the multi-GPU host plumbing (zkb200.parallel) -- sharding maps, the all-gather + host
point-sum of a point-range sharded MSM, and the layout / collective usage of the domain-sharded NTT -- with the oracle
standing in for the per-rank CUDA kernels (which cannot run without a GPU)."""
import os
import sys
import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, fn, ret):
    sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "oracle")); sys.path.insert(0, os.path.join(ROOT, "tests"))
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        ret[rank] = fn(rank, world)
    finally:
        dist.destroy_process_group()


def _free_port():
    import socket
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def run_world(fn, world=2, port=None):
    port = port or _free_port()
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_worker, args=(world, port, fn, ret), nprocs=world, join=True)
    return [ret[r] for r in range(world)]


def _msm_case(rank, world):
    import oracle_lib
    from util import rand_field
    from zkb200 import parallel
    orc = oracle_lib.load()
    n = 777
    k = rand_field(n, 5)
    bases = orc.g1_fixed_base_mul(orc.g1_generator(), k)
    s = rand_field(n, 6)
    lo, hi = parallel.shard_range(n, rank, world)
    part = orc.g1_to_affine(orc.best_multiexp(s[lo:hi], bases[lo:hi]))      # stand-in for the per-rank GPU MSM
    aff, comp = parallel.combine_msm_partials(part)
    full = orc.g1_to_affine(orc.best_multiexp(s, bases))
    return bool((aff == full).all()) and comp == orc.g1_compress(full), (lo, hi)


def test_sharded_msm_combine_gloo():
    res = run_world(_msm_case, 2)
    assert all(ok for ok, _ in res)
    assert res[0][1] == (0, 389) and res[1][1] == (389, 777)


class OracleOps:
    """CPU stand-in for parallel.DeviceOps."""
    def __init__(self):
        import oracle_lib
        self.o = oracle_lib.load()

    def _np(self, t): return t.numpy().view(np.uint64)
    def _t(self, a): return torch.from_numpy(np.ascontiguousarray(a).view(np.int64))
    def pow_omega(self, omega, e): return self.o.fr_pow(np.ascontiguousarray(omega), e)
    def local_ntt(self, x, omega, log_m): return self._t(self.o.best_fft(self._np(x), omega, log_m))
    def powers(self, base, n, like): return self._t(self.o.fr_powers(np.ascontiguousarray(base), n))
    def mul(self, a, b): return self._t(self.o.fr_mul(self._np(a).copy(), self._np(b).copy()))

    def cross(self, blocks, p, omega_p):
        b = self._np(blocks).reshape(p, -1, 4)
        out = np.zeros_like(b)
        for k in range(p):
            acc = np.zeros_like(b[0])
            for j in range(p):
                w = self.o.fr_pow(np.ascontiguousarray(omega_p), (j * k) % p)
                acc = self.o.fr_add(acc, self.o.fr_mul(np.ascontiguousarray(b[j]), np.repeat(w[None], b.shape[1], axis=0)))
            out[k] = acc
        return self._t(out.reshape(-1, 4))


def _ntt_case(rank, world):
    from util import rand_field
    from zkb200 import parallel
    ops = OracleOps()
    log_n = 9
    x = rand_field(1 << log_n, 77)
    omega = ops.o.fr_omega(log_n)
    local = torch.from_numpy(parallel.cyclic_shard(x, rank, world).copy().view(np.int64))
    strip = parallel.ntt_distributed(local, log_n, omega, ops)
    return strip.numpy().view(np.uint64)


def test_distributed_ntt_layout_gloo():
    import oracle_lib
    from util import rand_field
    from zkb200 import parallel
    strips = run_world(_ntt_case, 2)
    orc = oracle_lib.load()
    x = rand_field(1 << 9, 77)
    full = orc.best_fft(x, orc.fr_omega(9), 9)
    assert (parallel.strips_to_natural(strips, 2) == full).all()


def test_owner_and_ranges():
    from zkb200 import parallel
    assert [parallel.owner_of(i, 4) for i in range(6)] == [0, 1, 2, 3, 0, 1]
    cover = []
    for r in range(8):
        lo, hi = parallel.shard_range(1003, r, 8)
        cover += list(range(lo, hi))
    assert cover == list(range(1003))


def test_g1_sum_host_matches_oracle(oracle):
    from util import rand_field
    from zkb200 import parallel
    pts = oracle.g1_fixed_base_mul(oracle.g1_generator(), rand_field(9, 3))
    pts[4] = 0                                                # identity among the partials
    aff, comp = parallel.g1_sum_affine(pts)
    one = oracle.fq_from_canonical(np.array([[1, 0, 0, 0]], dtype=np.uint64))[0]
    acc = None
    for p in pts:
        j = np.concatenate([p, one if p.any() else np.zeros(4, dtype=np.uint64)])
        acc = j if acc is None else oracle.g1_add(acc, j)
    exp = oracle.g1_to_affine(acc)
    assert (aff == exp).all() and comp == oracle.g1_compress(exp)


def _deal_case(rank, world):
    """the dealing of the multi-GPU create_proof (prover.cu `Deal`, mirrored by parallel.Deal): contiguous blocks, one in-place
    all-gather, padding rows never read"""
    from zkb200 import parallel
    ok = True
    for count in (1, 2, 3, 7, 8, 16, 49, 128):
        d = parallel.Deal(count, rank, world)
        owners = [sum(1 for r in range(world) if parallel.Deal(count, r, world).mine(i)) for i in range(count)]
        ok &= all(o == (1 if d.on else world) for o in owners)          # every unit has exactly one owner (or is computed everywhere)
        ok &= d.padded() >= count and (not d.on or d.padded() - count < world)
        slab = torch.full((d.padded(), 4), -1, dtype=torch.int64)
        for i in range(count):
            if d.mine(i):
                slab[i] = torch.tensor([i, 10 * i, rank, 7])
        d.gather(slab)
        exp_rank = [(i // d.blk if d.on else rank) for i in range(count)]
        ok &= all(slab[i].tolist() == [i, 10 * i, exp_rank[i], 7] for i in range(count))
    return ok


def test_block_dealing_all_gather_gloo():
    assert all(run_world(_deal_case, 2))
