"""CPU check of the stand-in generator (tests/standins.py) and of the oracle prover / verifier on its shapes: witness produced with
the oracle's field arithmetic (OracleOps), proof by the oracle's restated create_proof, accepted by the oracle verifier; a flipped
proof byte and a broken witness cell are rejected.  Covers what the GPU parity tests rely on without needing a GPU: 56-point
rotation sets in SHPLONK, three phases, an instance column inside the permutation, two-column lookup tables."""
import numpy as np
import pytest

import halo2_ref as H
import standins
from test_gpu_prover_wide import to_oracle_cs


def oracle_prove(sc, s=777):
    cs = to_oracle_cs(sc.cs)
    ref = H.Ref(cs, s)
    assert ref.bf == sc.bf and ref.d == sc.cs.degree
    assert cs.advice_queries == sc.cs.advice_queries and cs.fixed_queries == sc.cs.fixed_queries and cs.instance_queries == sc.cs.instance_queries
    F, h, n, bf = ref.F, sc.host, sc.n, sc.bf
    fixed = [h(t) for t in sc.fixed]
    sigma = [h(t) for t in sc.sigma]
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
    inst = [F.ints(h(t)) for t in sc.instances]

    def synth(phase, ch):
        chm = {i: F.arr([v])[0] for i, v in ch.items()}
        return {c: h(t) for c, t in sc.synthesize_dev(phase, chm).items()}
    proof, dbg = ref.create_proof(pkr, trep, inst, synth, blinds)
    return ref, pkr, trep, inst, proof, dbg


@pytest.mark.parametrize("kind,k,kw", [("keccak", 9, dict(scale=0.12)), ("super", 8, dict(advice=40, n_gates=60))])
def test_oracle_proves_and_verifies_standin(kind, k, kw):
    ops = standins.OracleOps()
    sc = standins.keccak_shape(k, seed=1, ops=ops, **kw) if kind == "keccak" else standins.super_shape(k, seed=2, ops=ops, **kw)
    if kind == "keccak":
        assert sc.bf == 58 and sc.shape["distinct_rotations"] >= 56
    else:
        assert sc.cs.num_phases() == 3 and sc.cs.num_instance == 1 and (H.INSTANCE, 0) in [tuple(c) for c in sc.cs.perm_columns]
    ref, pkr, trep, inst, proof, dbg = oracle_prove(sc)
    assert all(v == 0 for v in dbg["phi_last"])
    assert ref.verify_proof(pkr, trep, inst, proof)
    bad = bytearray(proof)
    bad[len(bad) // 2] ^= 1
    try:
        ok = ref.verify_proof(pkr, trep, inst, bytes(bad))
    except Exception:
        ok = False
    assert not ok


def test_unsatisfied_gate_is_rejected():
    """break one defined cell: the prover still runs (gates are not checked while proving), the verifier must refuse"""
    ops = standins.OracleOps()
    sc = standins.super_shape(8, advice=40, n_gates=60, seed=3, ops=ops)
    sc.adv0[sc.c_def0][5] = ops.rand(1, 99)[0]
    ref, pkr, trep, inst, proof, _ = oracle_prove(sc)
    assert not ref.verify_proof(pkr, trep, inst, proof)
