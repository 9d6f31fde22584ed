"""Parity at the BASELINE shapes and sizes.

 * byte parity: the CUDA proving session vs the CPU oracle's restated halo2 prover on Keccak-shaped (56 rotations, 58 blinding
   factors, two-column tables, 2 phases) and SuperCircuit-shaped (3 phases, instance column in the permutation, condition *
   constraint gates) stand-ins at k = 11..13 -- sizes the oracle prover finishes in seconds;
 * BASELINE sizes (configs[2] k = 17, configs[3] k = 20): the oracle prover would need minutes to hours, so the proof the GPU
   produced is checked by the pinned oracle VERIFIER (it accepts the reference's own k = 25 proof, tests/test_fixture_proof.py),
   whose cost does not depend on k -- the acceptance criterion of the reference's own benches
   (circuit-benchmarks/src/packed_multi_keccak.rs:89-104, super_circuit.rs:134-154: prove, then verify_proof).  The vk commitments
   the verifier uses are cross-checked against the SRS trapdoor ([f(s)] G computed in the scalar field), so they do not depend
   on any MSM implementation.  Tampered proofs must be rejected.
Blinding-row conventions and RNG-facing inputs are boundary inputs here (prover bytes are unpinned without Rust)."""
import numpy as np
import pytest

import halo2_ref as H
import pyref as P
from test_gpu_prover_wide import to_oracle_cs

pytestmark = pytest.mark.gpu


def prove_gpu(sc, params, transcript="blake2b", pinned=False):
    from zkb200 import plonk as Z
    h = sc.host
    fixed = [h(t) for t in sc.fixed]
    sigma = [h(t) for t in sc.sigma]
    srs = params.load()
    pk = Z.ProvingKey(sc.cs, fixed, sigma, srs=srs)
    synth = lambda phase, ch: {c: h(t) for c, t in sc.synthesize_dev(phase, ch).items()}
    inst = [h(t) for t in sc.instances]
    proof = Z.create_proof(pk, h(sc.transcript_repr[None])[0], inst, synth, h(sc.z_blinds), h(sc.phi_blinds), h(sc.random_poly), transcript=transcript)
    return pk, proof, fixed, sigma, inst


@pytest.mark.parametrize("kind,k,kw", [("keccak", 12, dict(scale=0.25)), ("super", 11, dict(advice=48, scale=1.0, n_gates=90)),
                                       ("super", 13, dict(advice=64, scale=1.0, n_gates=120))])
def test_standin_proof_bytes_match_oracle(kind, k, kw):
    import standins
    from zkb200.params import ParamsKZG
    sc = standins.keccak_shape(k, seed=k, **kw) if kind == "keccak" else standins.super_shape(k, seed=k, **kw)
    if kind == "keccak":
        assert sc.bf == 58 and sc.shape["distinct_rotations"] >= 56     # keccak_packed_multi.rs:59-68: 59 unusable rows
    else:
        assert sc.cs.num_phases() == 3 and sc.cs.num_instance == 1
    cs = to_oracle_cs(sc.cs)
    ref = H.Ref(cs, 4321)
    assert ref.bf == sc.bf and ref.d == sc.cs.degree == 9
    assert cs.advice_queries == sc.cs.advice_queries and cs.fixed_queries == sc.cs.fixed_queries and cs.instance_queries == sc.cs.instance_queries
    params = ParamsKZG.unsafe_setup_with_s(k, 4321)
    pk, proof, fixed, sigma, inst = prove_gpu(sc, params)

    F = ref.F
    h = sc.host
    n, bf = sc.n, sc.bf
    pkr = {"fixed_values": fixed, "fixed_polys": [ref.lagrange_to_coeff(v) for v in fixed], "sigma_values": sigma,
           "sigma_polys": [ref.lagrange_to_coeff(v) for v in sigma]}
    l0 = np.zeros((n, 4), dtype=np.uint64); l0[0] = ref.w_arr(1)
    lb = np.zeros((n, 4), dtype=np.uint64); lb[n - bf:] = ref.w_arr(1)
    ll = np.zeros((n, 4), dtype=np.uint64); ll[n - bf - 1] = ref.w_arr(1)
    pkr["l0"], pkr["l_last"], pkr["l_blind"] = [ref.lagrange_to_coeff(v) for v in (l0, ll, lb)]
    pkr["fixed_commitments"] = [ref.commit_lagrange(v) for v in fixed]
    pkr["sigma_commitments"] = [ref.commit_lagrange(v) for v in sigma]
    zb, pb = h(sc.z_blinds), h(sc.phi_blinds)
    blinds = {"z": [F.ints(zb[i * bf:(i + 1) * bf]) for i in range(sc.nsets)], "phi": [F.ints(pb[i * bf:(i + 1) * bf]) for i in range(sc.L)],
              "random_poly": h(sc.random_poly)}
    trep = F.ints(h(sc.transcript_repr[None]))[0]
    inst_ints = [F.ints(a) for a in inst]

    def synth_ref(phase, ch):
        chm = {i: F.arr([v])[0] for i, v in ch.items()}
        return {c: h(t) for c, t in sc.synthesize_dev(phase, chm).items()}
    proof_ref, dbg = ref.create_proof(pkr, trep, inst_ints, synth_ref, blinds)
    assert all(v == 0 for v in dbg["phi_last"])
    assert ref.verify_proof(pkr, trep, inst_ints, proof_ref)
    assert proof == proof_ref
    # vk bytes (SerdeFormat::Processed) from the device == the oracle's commitments
    vk = pk.vk_bytes()
    exp = b"".join(bytes(ref.o.g1_compress(c)) for c in pkr["fixed_commitments"] + pkr["sigma_commitments"])
    assert vk[8:] == exp


def trapdoor_commit_lagrange(F, values_mont, k, s):
    """[f(s)] G for the polynomial with Lagrange values `values`: f(s) = sum_i v_i L_i(s), L_i(s) = w^i (s^n - 1) / (n (s - w^i)),
    all in the scalar field (python ints) -- independent of every MSM implementation."""
    n = 1 << k
    w = pow(P.FR_ROOT_OF_UNITY, 1 << (P.FR_S - k), P.R_MOD)
    wi = F.powers(w, n)
    den = F.scal(F.sub(F.full(s, n), wi), n)
    num = F.scal(wi, (pow(s, n, P.R_MOD) - 1) % P.R_MOD)
    terms = F.mul(F.mul(num, F.inv(den)), values_mont)
    c = F.o.fr_to_canonical(np.ascontiguousarray(terms))
    fs = sum(sum(c[:, i].tolist()) << (64 * i) for i in range(4)) % P.R_MOD
    return P.g1_mul(P.G1_GEN, fs)


def verify_gpu_proof(sc, pk_or_vk, proof, inst, s, fixed, sigma, spot=((0, "fixed"), (-1, "sigma"))):
    """oracle verifier over the device's vk (a ProvingKey, or its vk_bytes()); -> (accepted, tampered proof rejected, number of vk
    commitments cross-checked by the trapdoor)"""
    cs = to_oracle_cs(sc.cs)
    ref = H.Ref(cs, s, build_srs=False)
    F = ref.F
    vk = pk_or_vk if isinstance(pk_or_vk, (bytes, bytearray)) else pk_or_vk.vk_bytes()
    assert int.from_bytes(vk[:4], "big") == sc.k and int.from_bytes(vk[4:8], "big") == sc.cs.num_fixed
    pts = [P.g1_decompress(vk[8 + 32 * i: 40 + 32 * i]) for i in range((len(vk) - 8) // 32)]
    # an all-zero fixed column (an unused selector) commits to the identity: (0, 0) in halo2curves' affine layout
    to_aff = lambda pt: np.zeros(8, dtype=np.uint64) if pt is None else np.array(P.limbs(P.to_mont(pt[0], P.Q_MOD)) + P.limbs(P.to_mont(pt[1], P.Q_MOD)), dtype=np.uint64)
    nf = sc.cs.num_fixed
    pkr = {"fixed_commitments": [to_aff(p) for p in pts[:nf]], "sigma_commitments": [to_aff(p) for p in pts[nf:]]}
    checked = 0
    for idx, which in spot:
        vals, cm = (fixed[idx], pts[:nf][idx]) if which == "fixed" else (sigma[idx], pts[nf:][idx])
        assert trapdoor_commit_lagrange(F, vals, sc.k, s) == cm, f"vk commitment {which}[{idx}] is not [f(s)] G"
        checked += 1
    trep = F.ints(sc.host(sc.transcript_repr[None]))[0]
    inst_ints = [F.ints(a) for a in inst]
    ok = ref.verify_proof(pkr, trep, inst_ints, proof)
    bad = bytearray(proof)
    bad[len(bad) // 2] ^= 1          # an evaluation (or commitment) in the middle of the proof
    try:
        rejected = not ref.verify_proof(pkr, trep, inst_ints, bytes(bad))
    except Exception:
        rejected = True              # a flipped bit may make a point undecodable: also a rejection
    return ok, rejected, checked


def test_keccak_shape_k17_proof_verifies():
    """BASELINE configs[2] size: k = 17 Keccak-shaped stand-in (58 blinding factors, 46 656 - 78 125-row tables, 105 lookup input sets)."""
    import standins
    from zkb200.params import ParamsKZG
    s = 1234
    sc = standins.keccak_shape(17, seed=3)
    assert sc.bf == 58 and sc.shape["lookup_input_sets"] >= 100 and max(sc.shape["lookup_tables"]) == 78125
    params = ParamsKZG.unsafe_setup_with_s(17, s)
    pk, proof, fixed, sigma, inst = prove_gpu(sc, params)
    ok, rejected, checked = verify_gpu_proof(sc, pk, proof, inst, s, fixed, sigma)
    assert ok and rejected and checked == 2


def test_super_shape_k20_proof_verifies():
    """BASELINE configs[3] size: k = 20 SuperCircuit-shaped stand-in (3 phases, instance column, 128 advice columns)."""
    import standins
    from zkb200.params import ParamsKZG
    s = 1234   # zkevm-circuits/src/super_circuit/test.rs:74 uses the same toy trapdoor
    sc = standins.super_shape(20, advice=128, seed=5)
    assert sc.cs.num_phases() == 3 and sc.cs.num_instance == 1 and sc.shape["gates"] >= 600
    params = ParamsKZG.unsafe_setup_with_s(20, s)
    pk, proof, fixed, sigma, inst = prove_gpu(sc, params)
    ok, rejected, checked = verify_gpu_proof(sc, pk, proof, inst, s, fixed, sigma)
    assert ok and rejected and checked == 2
