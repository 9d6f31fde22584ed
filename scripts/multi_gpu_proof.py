"""torchrun --nproc-per-node N scripts/multi_gpu_proof.py K N_GATES N_LOOKUPS N_PERM
Multi-GPU create_proof (commitment batches and quotient coset parts dealt across ranks over NCCL) must emit exactly the
single-GPU proof; reports the wall-clock of both (max over ranks)."""
import os, sys, json, time, hashlib
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np, torch, torch.distributed as dist


def run(k, ng, nl, npm, reps=3):
    """all ranks of an initialised NCCL process group call this; returns the result dict on rank 0 (None elsewhere)."""
    rank, local, world = int(os.environ["RANK"]), int(os.environ["LOCAL_RANK"]), int(os.environ["WORLD_SIZE"])
    import zkb200
    from zkb200 import plonk as Z
    from wide_circuit import WideCircuit
    from zkb200.params import ParamsKZG
    wc = WideCircuit(k, n_gates=ng, n_lookups=nl, n_perm=npm, two_phase=True, seed=3)
    params = ParamsKZG.unsafe_setup_with_s(k, 1234)
    keep = []

    def h(t):
        p = torch.empty(t.shape, dtype=t.dtype).pin_memory(); p.copy_(t); keep.append(p)
        return p.numpy().view(np.uint64)
    fixed = [h(t) for t in wc.fixed]; sigma = [h(t) for t in wc.sigma]
    g, gl = h(params.g), h(params.g_lagrange)
    cols0 = {c: h(t) for c, t in wc.synthesize_dev(0, {}).items()}

    def synth(phase, ch):
        return cols0 if phase == 0 else {c: h(t) for c, t in wc.synthesize_dev(phase, ch).items()}
    zb, pb, rp, tr = h(wc.z_blinds), h(wc.phi_blinds), h(wc.random_poly), h(wc.transcript_repr[None])[0]

    def timed(pk):
        best, proof = 1e9, None
        for _ in range(reps):
            dist.barrier(); torch.cuda.synchronize()
            t0 = time.perf_counter()
            proof = Z.create_proof(pk, tr, [], synth, zb, pb, rp)
            torch.cuda.synchronize()
            dt = torch.tensor([time.perf_counter() - t0], device="cuda"); dist.all_reduce(dt, op=dist.ReduceOp.MAX)
            best = min(best, float(dt.item()))
        return proof, best
    ctx1 = zkb200.Context(local)
    pk1 = Z.ProvingKey(wc.cs, fixed, sigma, g, gl, ctx=ctx1)
    proof1, t1 = timed(pk1)
    pk1.close()
    ctxn = zkb200.Context(local)
    ctxn.init_comm()
    pkn = Z.ProvingKey(wc.cs, fixed, sigma, g, gl, ctx=ctxn)
    proofn, tn = timed(pkn)
    same = torch.tensor([int(proofn == proof1)], device="cuda"); dist.all_reduce(same, op=dist.ReduceOp.MIN)
    pkn.close()
    ctxn.close()
    ctx1.close()
    if rank == 0:
        return {"k": k, "world": world, "advice_columns": wc.cs.num_advice, "single_gpu_seconds": t1, "multi_gpu_seconds": tn,
                "speedup": t1 / tn, "identical_proof_on_all_ranks": bool(same.item()), "proof_sha256": hashlib.sha256(proofn).hexdigest()[:16]}
    return None


def main():
    k, ng, nl, npm = [int(x) for x in sys.argv[1:5]]
    local = int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    res = run(k, ng, nl, npm)
    if res is not None:
        print(json.dumps(res))
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
