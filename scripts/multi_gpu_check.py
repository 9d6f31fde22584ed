"""torchrun --nproc-per-node N scripts/multi_gpu_check.py [quick] : validates the multi-GPU paths on real GPUs:
  1. domain-sharded NTT through the C ABI (zkb_ntt_fr_sharded_dev: twiddle + all-to-all fused into the transform kernels over
     NVLink peer memory; and the ncclSend/ncclRecv baseline) == single-GPU NTT of the same input, forward and round trip,
  2. point-range sharded MSM (zkb_msm_g1_sharded_dev) == single-GPU MSM,
and times both (device events, max over ranks).  Prints one JSON line on rank 0; exit code 1 on any mismatch."""
import os, sys, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch, torch.distributed as dist


def run(sizes=(16, 22, 24), msm_log=20, reps=3, ctx=None):
    rank, world = dist.get_rank(), dist.get_world_size()
    import zkb200
    from zkb200 import arithmetic as A, parallel
    from zkb200.params import g1_generator
    if ctx is None:
        ctx = zkb200.default_context(torch.cuda.current_device())
        ctx.init_comm()
    out = {"world": world}

    def timed(fn):
        best = None
        for _ in range(reps):
            dist.barrier(); torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); r = fn(); e1.record(); torch.cuda.synchronize()
            ms = torch.tensor([e0.elapsed_time(e1)], device="cuda"); dist.all_reduce(ms, op=dist.ReduceOp.MAX)
            best = float(ms.item()) if best is None else min(best, float(ms.item()))
        return r, best
    for log_n in sizes:
        n = 1 << log_n
        omega, omega_inv = A.root_of_unity(log_n)
        n_can = torch.tensor([[n, 0, 0, 0]], dtype=torch.int64, device="cuda")
        ninv = A.field_unop_dev(A.FR, A.UOP_INV, A.field_unop_dev(A.FR, A.UOP_TO_MONT, n_can)).cpu().numpy().view(np.uint64)[0]
        full = A.random_fr_dev(n, 4242)                    # same seed on every rank -> same array everywhere
        local_x = full[rank::world].contiguous()
        ref = A.best_fft_dev(full.clone(), omega, log_n)
        M = n // world; blk = M // world
        exp = torch.cat([ref[k * M + rank * blk: k * M + (rank + 1) * blk] for k in range(world)])
        res = {}
        for mode in ("p2p", "nccl"):
            os.environ["ZKB_SHARDED_EXCHANGE"] = mode
            strip, ms_f = timed(lambda: parallel.ntt_sharded_dev(local_x, log_n, omega, 0, ctx=ctx))
            back, ms_b = timed(lambda: parallel.ntt_sharded_dev(strip, log_n, omega_inv, 1, scale=ninv, ctx=ctx))
            res[mode] = {"forward_ok": bool(torch.equal(strip, exp)), "round_trip_ok": bool(torch.equal(back, local_x)), "forward_ms": ms_f, "inverse_ms": ms_b}
        os.environ["ZKB_SHARDED_EXCHANGE"] = "p2p"
        t1 = torch.cuda.Event(enable_timing=True); t2 = torch.cuda.Event(enable_timing=True)
        t1.record(); A.best_fft_dev(full, omega, log_n); t2.record(); torch.cuda.synchronize()
        res["single_gpu_ms"] = t1.elapsed_time(t2)
        res["ok"] = all(v["forward_ok"] and v["round_trip_ok"] for v in res.values() if isinstance(v, dict))
        out[f"ntt_2^{log_n}"] = res
        del full, ref, exp, strip, back
        torch.cuda.empty_cache()
    n = 1 << msm_log
    gen = g1_generator()
    bases = A.g1_fixed_base_mul_dev(gen, A.random_fr_dev(n, 7))
    scal = A.random_fr_dev(n, 8)
    lo, hi = parallel.shard_range(n, rank, world)
    (aff, comp), ms = timed(lambda: parallel.msm_sharded_dev(scal[lo:hi].contiguous(), bases[lo:hi].contiguous(), ctx=ctx))
    single = A.best_multiexp_dev(scal, bases)
    out[f"msm_2^{msm_log}"] = {"ok": bool(comp == single.compressed), "sharded_ms": ms}
    oks = torch.tensor([int(all(v.get("ok", True) for v in out.values() if isinstance(v, dict)))], device="cuda")
    dist.all_reduce(oks, op=dist.ReduceOp.MIN)
    out["all_ranks_ok"] = bool(oks.item())
    return out


def main():
    rank, local, world = int(os.environ["RANK"]), int(os.environ["LOCAL_RANK"]), int(os.environ["WORLD_SIZE"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    quick = len(sys.argv) > 1 and sys.argv[1] == "quick"
    out = run(sizes=(12, 16) if quick else (16, 22, 24), msm_log=14 if quick else 20)
    if rank == 0:
        print(json.dumps(out))
    dist.destroy_process_group()
    sys.exit(0 if out["all_ranks_ok"] else 1)


if __name__ == "__main__":
    main()
