"""torchrun --nproc-per-node N scripts/multi_gpu_check.py : validates the multi-GPU paths on real GPUs over NCCL:
  1. domain-sharded NTT (one all-to-all) == single-GPU NTT of the same input,
  2. point-range sharded MSM == single-GPU MSM,
and times both (device events, max over ranks)."""
import os, sys, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch, torch.distributed as dist


def main():
    rank, local, world = int(os.environ["RANK"]), int(os.environ["LOCAL_RANK"]), int(os.environ["WORLD_SIZE"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    from zkb200 import arithmetic as A, parallel
    from zkb200.params import g1_generator
    ops = parallel.DeviceOps()
    out = {"world": world}
    # ---- NTT
    for log_n in (16, 24, 26):
        n = 1 << log_n
        omega, _ = A.root_of_unity(log_n)
        full = A.random_fr_dev(n, 4242)                    # same seed on every rank -> same array everywhere
        local_x = full[rank::world].contiguous()
        for it in range(3):
            dist.barrier(); torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            strip = parallel.ntt_distributed(local_x.clone(), log_n, omega, ops)
            e1.record(); torch.cuda.synchronize()
            ms = torch.tensor([e0.elapsed_time(e1)], device="cuda"); dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        ref = A.best_fft_dev(full.clone(), omega, log_n)
        M = n // world; blk = M // world
        exp = torch.cat([ref[k * M + rank * blk: k * M + (rank + 1) * blk] for k in range(world)])
        ok = torch.equal(strip, exp)
        t1 = torch.cuda.Event(enable_timing=True); t2 = torch.cuda.Event(enable_timing=True)
        t1.record(); A.best_fft_dev(full, omega, log_n); t2.record(); torch.cuda.synchronize()
        out[f"ntt_2^{log_n}"] = {"ok": bool(ok), "distributed_ms": float(ms.item()), "single_gpu_ms": t1.elapsed_time(t2)}
        del full, ref, strip, exp
        torch.cuda.empty_cache()
    # ---- MSM
    n = 1 << 20
    gen = g1_generator()
    bases = A.g1_fixed_base_mul_dev(gen, A.random_fr_dev(n, 7))
    scal = A.random_fr_dev(n, 8)
    lo, hi = parallel.shard_range(n, rank, world)
    for it in range(3):
        dist.barrier(); torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        aff, comp = parallel.best_multiexp_sharded(scal[lo:hi].contiguous(), bases[lo:hi].contiguous())
        e1.record(); torch.cuda.synchronize()
        ms = torch.tensor([e0.elapsed_time(e1)], device="cuda"); dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    single = A.best_multiexp_dev(scal, bases)
    out["msm_2^20"] = {"ok": bool(comp == single.compressed), "sharded_ms": float(ms.item())}
    oks = torch.tensor([int(all(v.get("ok", True) for v in out.values() if isinstance(v, dict)))], device="cuda")
    dist.all_reduce(oks, op=dist.ReduceOp.MIN)
    out["all_ranks_ok"] = bool(oks.item())
    if rank == 0:
        print(json.dumps(out))
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
