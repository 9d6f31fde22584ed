#!/usr/bin/env python
"""bench.py -- headline benchmark of the B200 Halo2/KZG hot path.

Workload at N=1 (BASELINE.json configs[1]): 2^24-point Fr NTT forward + inverse round trip, data resident in HBM.
A "step" is one round trip (forward NTT, then inverse NTT with the 1/n scaling fused).  N>1: every rank runs its own
2^24 round trip on its own GPU (independent columns of a proof shard across GPUs with no data-path collective: weak
scaling).  `value` = Fr butterflies per second over all ranks; `e2e` = the same through the host-pointer C-ABI entry
point (H2D + kernels + D2H inside the timed region).  Extras: 2^20-point G1 MSM (configs[0]) G1-adds/s.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
"""
import argparse
import json
import os
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

LOG_N = 24
N_PTS = 1 << LOG_N
BUTTERFLIES_PER_DIR = (N_PTS // 2) * LOG_N          # 201,326,592  
ALG_BYTES_PER_DIR = 2 * 32 * N_PTS                  # 1,073,741,824 B: read + write each element once per transform
METRIC = "NTT Fr-butterflies/s (2^24 fwd+inv round trip)"
UNIT = "butterflies/s"
MSM_LOG_N = 20
CONFIG = {"workload": "2^24-point Fr NTT forward+inverse round trip per GPU (BASELINE configs[1])", "log_n": LOG_N,
          "l2": "working set 512 MiB per transform > 126 MB L2 (no flush needed)", "per_rank": "independent transform per rank"}


def dist_env():
    return int(os.environ.get("RANK", 0)), int(os.environ.get("LOCAL_RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (recipe in B200_PROFILING.md)."""
    Q = "index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown," \
        "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap,utilization.gpu"

    def __init__(self, index):
        self.index = index
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "50", "-i", str(self.index)],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
        except Exception:
            self.proc = None

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            out, _ = self.proc.communicate(timeout=5)
        except Exception:
            self.proc.kill()
            out = ""
        sm, mx, reasons, pw = [], [], set(), []
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for line in out.strip().splitlines():
            f = [x.strip() for x in line.split(",")]
            if len(f) < 9:
                continue
            try:
                if float(f[8]) < 50:       # keep only samples taken under load (GPU utilisation >= 50 %)
                    continue
                sm.append(float(f[1])); mx.append(float(f[2])); pw.append(float(f[3]))
            except ValueError:
                continue
            for nm, v in zip(names, f[4:8]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "power_w_max": max(pw) if pw else None, "samples": len(sm), "reasons": sorted(reasons)}


def cpu_proof_model(orc, shape, threads_note=""):
    """CPU cost model of one create_proof on the host cores from MEASURED oracle primitives (the Rust prover cannot run here):
    time one best_multiexp and one best_fft of the circuit's sizes with the oracle (all cores) and multiply by the number of
    commitments / transforms halo2's create_proof performs for this shape (one MSM per committed polynomial, one size-n
    iFFT per committed column, one extended (k + 3) FFT per polynomial entering the quotient, one extended iFFT).  Quotient
    evaluation, permutation / lookup scans and witness generation are NOT included, so this is a lower bound of the CPU prover."""
    import numpy as np
    k, A, L, Pn = shape["k"], shape["advice_columns"], shape["lookup_arguments"], shape["permutation_columns"]
    nf, d = shape.get("fixed_columns", 3), shape.get("cs_degree", 9)
    n = 1 << k
    nsets = (Pn + (d - 2) - 1) // (d - 2)
    rng = np.random.default_rng(5)

    def rand_fr(m):
        a = rng.integers(0, 1 << 64, size=(m, 4), dtype=np.uint64)
        a[:, 3] = rng.integers(0, 0x30644E72E131A029, size=m, dtype=np.uint64)
        return a
    # bases: multiples of the generator by small scalars are enough for timing (bucket additions dominate, not the points' values)
    m_small = 1 << 12
    small = orc.g1_fixed_base_mul(orc.g1_generator(), rand_fr(m_small))
    bases = np.tile(small, (n // m_small, 1)) if n >= m_small else small[:n]
    s = rand_fr(n)
    orc.best_multiexp(s[:1 << 10], bases[:1 << 10])          # warm up the thread pool
    t0 = time.perf_counter(); orc.best_multiexp(s, bases); t_msm = time.perf_counter() - t0
    w = orc.fr_omega(k)
    a = rand_fr(n)
    orc.best_fft(a, w, k)
    t0 = time.perf_counter(); orc.best_fft(a, w, k); t_fft = time.perf_counter() - t0
    ek = k + 3
    t_fft_ext = t_fft * ((1 << ek) * ek) / (n * k)          # scaled n log n (an extended transform at k = 23 needs GBs per call)
    commits = A + 2 * L + nsets + 1 + (d - 1) + 2
    iffts = A + 2 * L + nsets
    ext_ffts = A + nf + Pn + nsets + 2 * L + 3
    total = commits * t_msm + iffts * t_fft + (ext_ffts + 1) * t_fft_ext
    return {"seconds_lower_bound": total, "msm_seconds": t_msm, "fft_seconds": t_fft, "extended_fft_seconds_scaled": t_fft_ext,
            "commitments": commits, "iffts": iffts, "extended_ffts": ext_ffts + 1, "cores": orc.num_threads(),
            "note": "measured oracle primitives x halo2 operation counts; excludes quotient evaluation, scans and witness generation"}


def run_reference(args):
    """--impl reference: the reference's CPU algorithm for the path (oracle best_fft restatement; the Rust crate cannot
    be built here: no cargo/rustc) on all host cores, same workload and metric."""
    rank, _, world = dist_env()
    if rank != 0:
        return
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import numpy as np
    import oracle_lib
    orc = oracle_lib.load()
    rng = np.random.default_rng(1)
    a = rng.integers(0, 1 << 64, size=(N_PTS, 4), dtype=np.uint64)
    a[:, 3] = rng.integers(0, 0x30644E72E131A029, size=N_PTS, dtype=np.uint64)
    w = orc.fr_omega(LOG_N)
    wi = orc.fr_inv(w[None])[0]
    import ctypes
    p = a.ctypes.data_as(ctypes.POINTER(ctypes.c_uint64))
    wp = w.ctypes.data_as(ctypes.POINTER(ctypes.c_uint64))
    wip = wi.ctypes.data_as(ctypes.POINTER(ctypes.c_uint64))

    def step():
        orc.lib.zko_best_fft(p, wp, ctypes.c_uint32(LOG_N))
        orc.lib.zko_best_fft(p, wip, ctypes.c_uint32(LOG_N))

    for _ in range(min(args.warmup, 1)):
        step()
    steps = max(1, min(args.steps, 5))
    t0 = time.perf_counter()
    for _ in range(steps):
        step()
    dt = (time.perf_counter() - t0) / steps
    val = 2 * BUTTERFLIES_PER_DIR / dt
    cores = orc.num_threads()
    line = {"impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus, "steps": steps, "warmup": min(args.warmup, 1),
            "ms_per_step": dt * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u256 (4x u64 Montgomery, BN254 Fr)",
            "data": "synthetic", "config": CONFIG,
            "cpu_baseline": {"value": val, "unit": UNIT, "cores": cores, "kind": "port",
                             "sample": f"{steps} full 2^24 fwd+inv round trips, oracle best_fft (OpenMP)"},
            "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=150)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-msm", action="store_true")
    ap.add_argument("--no-proof", action="store_true")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)

    import numpy as np
    import torch
    rank, local_rank, world = dist_env()
    assert torch.cuda.is_available(), "bench.py needs a CUDA device (no CPU fallback)"
    torch.cuda.set_device(local_rank)
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    import zkb200
    from zkb200 import arithmetic as A
    ctx = zkb200.default_context(local_rank)
    W = max(3, args.warmup)
    K = args.steps

    def barrier():
        if world > 1:
            import torch.distributed as dist
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(ms):
        if world > 1:
            import torch.distributed as dist
            t = torch.tensor([ms], dtype=torch.float64, device="cuda")
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            return float(t.item())
        return ms

    w, wi = A.root_of_unity(LOG_N)
    # n^-1 in Montgomery form, computed on the device (field kernels), no oracle involved
    n_can = torch.tensor([[N_PTS, 0, 0, 0]], dtype=torch.int64, device="cuda")
    ninv = A.field_unop_dev(A.FR, A.UOP_INV, A.field_unop_dev(A.FR, A.UOP_TO_MONT, n_can)).cpu().numpy().view(np.uint64)[0]
    data = A.random_fr_dev(N_PTS, 1000 + rank)
    orig = data.clone()

    def step_dev():
        A.best_fft_dev(data, w, LOG_N)
        A.best_fft_dev(data, wi, LOG_N, scale=ninv)

    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    for _ in range(W):
        step_dev()
    barrier()
    assert torch.equal(data, orig), "round trip does not restore the input"
    launches0 = ctx.launch_count
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record()
    for _ in range(K):
        step_dev()
    e1.record()
    barrier()
    ms_total = max_over_ranks(e0.elapsed_time(e1))
    launches = ctx.launch_count - launches0
    # nvidia-smi samples every 50 ms: a short timed region (small --steps) yields too few samples under load, so keep the same
    # kernels running (untimed) until the load window is >= 1 s before stopping the sampler
    extended = 0
    try:
        per_step_ms = max(ms_total / max(K, 1), 1e-3)
        if ms_total < 1000.0:
            extended = int(min(400, (1000.0 - ms_total) / per_step_ms + 1))
            for _ in range(extended):
                step_dev()
            torch.cuda.synchronize()
    except Exception:
        extended = -1
    barrier()
    clocks = sampler.stop() if rank == 0 else None
    if clocks is not None:
        clocks["load_window"] = "timed region" if extended == 0 else f"timed region + {extended} untimed identical steps"
    ms_step = ms_total / K
    value = world * 2 * BUTTERFLIES_PER_DIR / (ms_step * 1e-3)

    # end to end through the host-pointer C-ABI call: pinned host buffer, H2D + D2H inside the timed region
    host = torch.empty((N_PTS, 4), dtype=torch.int64).pin_memory()
    host.copy_(orig)
    Ke = max(2, min(K, 5))

    def step_e2e():
        A.best_fft_pinned(host, w, LOG_N)
        A.best_fft_pinned(host, wi, LOG_N, scale=ninv)

    step_e2e()
    barrier()
    t0 = time.perf_counter()
    for _ in range(Ke):
        step_e2e()
    torch.cuda.synchronize()
    e2e_ms = max_over_ranks((time.perf_counter() - t0) * 1e3 / Ke)
    barrier()
    assert torch.equal(host, orig.cpu()), "e2e round trip does not restore the input"
    e2e_val = world * 2 * BUTTERFLIES_PER_DIR / (e2e_ms * 1e-3)

    # roofline of the dominant kernel (ntt_pass_kernel): 2 launches per transform, each reads + writes the array once.
    # Algorithmic bytes of a transform = 2*32*n; one launch does half of a transform's passes.
    pass_launches = launches                    # ntt_tile_kernel launches inside the timed region (3 passes per direction at 2^24)
    passes_per_dir = pass_launches // (2 * K)
    avg_launch_s = (ms_total * 1e-3) / pass_launches
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    peak = float(peaks.get("hbm_gbs", 6650.0))
    achieved = (ALG_BYTES_PER_DIR / passes_per_dir) / avg_launch_s / 1e9
    roofline = {"bound": "hbm", "kernel": "ntt_tile_kernel", "launches_per_transform": passes_per_dir, "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                "peak_source": "MEASURED_PEAKS.json hbm_gbs (measured copy)" if peaks else "fallback 6650 GB/s",
                # dram__bytes_read.sum + dram__bytes_write.sum per ntt_pass_kernel launch at 2^24, averaged over the two passes of a
                # transform (ncu --set full, profiles/r01_ntt_pass_v3_final_ncu.txt): pass 1 = 1074 MB read (537 MB data + 537 MB
                # complete inter-pass twiddle table: HBM traffic deliberately traded for one multiply per element) + 507 MB
                # written, pass 2 = 537 MB + 489 MB.  Algorithmic: 536.9 MB per launch.
                "traffic": 1303.9e6,
                "note": "kernel is integer-multiply-pipe bound (1 Montgomery mul = 139 IMAD.WIDE per butterfly), see DESIGN.md; "
                        "alg bytes/launch = 2*32*2^24/2 passes"}

    extras = {}
    if not args.no_msm:
        n = 1 << MSM_LOG_N
        sys.path.insert(0, os.path.join(ROOT, "oracle"))
        gen = np.zeros(8, dtype=np.uint64)
        # generator (1, 2) in Montgomery form via the device kernels
        g_can = torch.tensor([[1, 0, 0, 0], [2, 0, 0, 0]], dtype=torch.int64, device="cuda")
        gen[:] = A.field_unop_dev(A.FQ, A.UOP_TO_MONT, g_can).cpu().numpy().view(np.uint64).reshape(8)
        bases = A.g1_fixed_base_mul_dev(gen, A.random_fr_dev(n, 7 + rank))
        scal = A.random_fr_dev(n, 77 + rank)
        for _ in range(2):
            A.best_multiexp_dev(scal, bases)
        barrier()
        m0, m1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        reps = 5
        m0.record()
        for _ in range(reps):
            r = A.best_multiexp_dev(scal, bases)
        m1.record()
        barrier()
        msm_ms = max_over_ranks(m0.elapsed_time(m1) / reps)
        adds = A.msm_last_adds(ctx)
        extras["msm_2^20"] = {"ms": msm_ms, "g1_adds": adds, "g1_adds_per_s": world * adds / (msm_ms * 1e-3),
                              "alg_bytes": n * 96, "achieved_gbs": n * 96 / (msm_ms * 1e-3) / 1e9,
                              "frac_of_hbm": n * 96 / (msm_ms * 1e-3) / 1e9 / peak, "commitment": r.compressed.hex()}

    if not args.no_proof and rank == 0:
        # BASELINE configs[2]/[3] stand-ins: Keccak-shaped k=17 and SuperCircuit-shaped k=20 synthetic circuits,
        # full create_proof through the C-ABI session (H2D of every advice column inside the timed region).
        sys.path.insert(0, os.path.join(ROOT, "scripts"))
        import proof_bench
        try:
            extras["proof_keccak_shape_k17"] = proof_bench.run(17, 64, 8, 16, reps=3)
            extras["proof_super_shape_k20"] = proof_bench.run(20, 64, 8, 16, reps=2)
        except Exception as e:  # keep the headline line even if the extra fails
            extras["proof_error"] = repr(e)
    barrier()
    if not args.no_proof and world > 1:
        # one proof spread over all ranks (commitment batches, coset parts, lookups dealt across the GPUs over NCCL): strong scaling of
        # BASELINE configs[3] (SuperCircuit-shaped k = 20); every rank participates, rank 0 reports
        sys.path.insert(0, os.path.join(ROOT, "scripts"))
        import multi_gpu_proof
        try:
            res = multi_gpu_proof.run(20, 64, 8, 16, reps=2)
            if rank == 0:
                extras["proof_super_shape_k20_multi_gpu"] = res
        except Exception as e:
            if rank == 0:
                extras["proof_multi_gpu_error"] = repr(e)
    barrier()

    cpu_baseline = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        sys.path.insert(0, os.path.join(ROOT, "oracle"))
        import oracle_lib
        orc = oracle_lib.load()
        a = orig.cpu().numpy().view(np.uint64).copy()
        t0 = time.perf_counter()
        f = orc.best_fft(a, w, LOG_N)
        b = orc.best_fft(f, wi, LOG_N)
        dt = time.perf_counter() - t0
        cpu_baseline = {"value": 2 * BUTTERFLIES_PER_DIR / dt, "unit": UNIT, "cores": orc.num_threads(), "kind": "port",
                        "sample": "one full 2^24 fwd+inv round trip (2 x 201,326,592 butterflies), oracle best_fft, OpenMP all cores",
                        "seconds": dt}
        # proof-level context: CPU lower bound for the k = 17 shape from measured primitives (about 10 s of CPU work)
        try:
            if "proof_keccak_shape_k17" in extras:
                extras["proof_keccak_shape_k17"]["cpu_model"] = cpu_proof_model(orc, extras["proof_keccak_shape_k17"])
        except Exception as e:
            extras["cpu_model_error"] = repr(e)

    if rank == 0:
        line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": K, "warmup": W, "ms_per_step": ms_step,
                "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
                "dtype": "u256 (8x u32 Montgomery limbs, BN254 Fr)", "data": "synthetic",
                "config": CONFIG,
                "e2e": {"value": e2e_val, "unit": UNIT, "ms_per_step": e2e_ms, "h2d_bytes_per_step": 2 * 32 * N_PTS, "d2h_bytes_per_step": 2 * 32 * N_PTS},
                "gpu_launches": launches, "roofline": roofline, "cpu_baseline": cpu_baseline, "clocks": clocks, "extras": extras}
        print(json.dumps(line), flush=True)
    if world > 1:
        import torch.distributed as dist
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
