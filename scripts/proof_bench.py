"""Time create_proof on a synthetic WideCircuit (stand-in shapes for BASELINE configs[2] / configs[3]).
usage: python scripts/proof_bench.py K N_GATES N_LOOKUPS N_PERM [REPS]"""
import os, sys, time, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np, torch


def run(k, n_gates, n_lookups, n_perm, reps=2, two_phase=True):
    import zkb200
    from zkb200 import plonk as Z
    from wide_circuit import WideCircuit
    from zkb200.params import ParamsKZG
    ctx = zkb200.default_context()
    t0 = time.perf_counter()
    wc = WideCircuit(k, n_gates=n_gates, n_lookups=n_lookups, n_perm=n_perm, two_phase=two_phase, seed=3)
    params = ParamsKZG.unsafe_setup_with_s(k, 1234)
    torch.cuda.synchronize()
    t_setup = time.perf_counter() - t0
    def h(t):
        """device tensor -> numpy view of a PINNED host copy (what a shim-allocated advice column would be)"""
        p = torch.empty(t.shape, dtype=t.dtype).pin_memory()
        p.copy_(t)
        keep.append(p)
        return p.numpy().view(np.uint64)
    keep = []
    fixed = [h(t) for t in wc.fixed]
    sigma = [h(t) for t in wc.sigma]
    t0 = time.perf_counter()
    pk = Z.ProvingKey(wc.cs, fixed, sigma, h(params.g), h(params.g_lagrange))
    t_pk = time.perf_counter() - t0
    # witness on the host (as Rust's synthesize would leave it): phase-0 columns precomputed, phase-1 computed in the callback
    cols0 = {c: h(t) for c, t in wc.synthesize_dev(0, {}).items()}

    def synth(phase, ch):
        if phase == 0: return cols0
        return {c: h(t) for c, t in wc.synthesize_dev(phase, ch).items()}
    zb, pb, rp, tr = h(wc.z_blinds), h(wc.phi_blinds), h(wc.random_poly), h(wc.transcript_repr[None])[0]
    times = []
    proof = None
    for r in range(reps + 1):
        l0 = ctx.launch_count
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        proof = Z.create_proof(pk, tr, [], synth, zb, pb, rp)
        torch.cuda.synchronize()
        times.append(time.perf_counter() - t0)
        launches = ctx.launch_count - l0
    na = wc.cs.num_advice
    res = {"k": k, "advice_columns": na, "fixed_columns": 3, "lookup_arguments": n_lookups, "lookup_input_sets": 3 * n_lookups, "permutation_columns": n_perm,
           "gates": len(wc.cs.gates), "cs_degree": wc.cs.degree, "phases": wc.cs.num_phases(), "proof_bytes": len(proof),
           "seconds_best": min(times[1:]), "seconds_median": sorted(times[1:])[len(times[1:]) // 2], "seconds_all": times, "host_buffers": "pinned", "kernel_launches": launches, "h2d_bytes": na * (1 << k) * 32,
           "setup_seconds": t_setup, "pk_upload_seconds": t_pk}
    pk.close()
    return res


if __name__ == "__main__":
    a = [int(x) for x in sys.argv[1:]]
    print(json.dumps(run(*a[:4], reps=a[4] if len(a) > 4 else 2)))
