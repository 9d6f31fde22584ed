"""THE parity anchor for the whole-proof path (CPU): the oracle's restated halo2 verifier -- its own gate / permutation / logUp
formulas, evaluation and query order, SHPLONK algebra -- accepts the REFERENCE'S OWN proof: aggregator/data/batch-task.json
chunk_proofs[0] (a k = 25 thin-compression proof produced by Scroll's prover, used by aggregator/src/tests/aggregation.rs:160,244),
read with the restated Poseidon transcript of snark-verifier-sdk and decided by a real pairing against the production SRS element
PARAMS_G2_SECRET_POWER (prover/src/utils.rs:36).  The same verifier (with the Blake2b transcript the reference's benches use)
is what accepts the CUDA prover's proofs in tests/test_gpu_prover*.py."""
import re
import numpy as np
import pytest

import pyref as P
import halo2_ref as H
import pairing_ref as E
import poseidon_ref as PO
from circuits import ThinCompressionShape

R = P.R_MOD


PoseidonReader = H.Ref.PoseidonReader     # snark_verifier PoseidonTranscript<NativeLoader>: points absorbed as (x mod r, y mod r)


@pytest.fixture(scope="module")
def setup(golden):
    proof = bytes.fromhex(golden["proof_hex"])
    raw = bytes.fromhex(golden["instances_hex"])
    instances = [[int.from_bytes(raw[i: i + 32], "big") for i in range(0, len(raw), 32)]]
    nums = [int(x, 16) for x in re.findall(r"0x([0-9a-f]{64})", golden["params_g2_secret_power"])]
    s_g2 = (E.FQ2([nums[0], nums[1]]), E.FQ2([nums[2], nums[3]]))
    assert E.g2_is_on_curve(s_g2) and E.g2_is_on_curve(E.G2)
    cs = ThinCompressionShape.constraint_system(golden["domain"]["k"])          # the circuit the fixture's Protocol describes
    ref = H.Ref(cs, None, build_srs=False)
    pre = [np.array(p["x"] + p["y"], dtype=np.uint64) for p in golden["preprocessed"]]
    pk = {"fixed_commitments": pre[:4], "sigma_commitments": pre[4:]}
    tis = P.from_mont(P.from_limbs(golden["transcript_initial_state"]), R)     # vk.transcript_repr
    decide = lambda lhs, rhs: E.pairing_product_is_one([(lhs, E.G2), (P.g1_neg(rhs), s_g2)])   # e(lhs, g2) == e(rhs, [s]g2)
    return ref, pk, tis, instances, proof, PO.Spec(5, 8, 60), decide


def test_pairing_is_bilinear():
    a, b = 1234567, 987654321
    e1 = E.pairing(E.G2, P.G1_GEN)
    assert not (e1 == E.FQ12.one())
    assert E.pairing(E.g2_mul(E.G2, b), P.g1_mul(P.G1_GEN, a)) == e1 ** (a * b)
    assert E.g2_mul(E.G2, R) is None


def test_reference_fixture_proof_verifies(setup):
    ref, pk, tis, instances, proof, spec, decide = setup
    assert ref.bf == 6 and ref.d == 5 and ref.dom.qdeg == 4
    assert ref.verify_proof(pk, tis, instances, proof, reader=PoseidonReader(proof, spec), decide=decide)


def test_reference_fixture_rejects_tampering(setup):
    ref, pk, tis, instances, proof, spec, decide = setup
    bad = bytearray(proof)
    bad[32 * 12 + 5] ^= 1                                   # an advice evaluation
    assert not ref.verify_proof(pk, tis, instances, bytes(bad), reader=PoseidonReader(bytes(bad), spec), decide=decide)
    inst = [list(instances[0])]
    inst[0][20] ^= 1                                        # a public-input byte
    assert not ref.verify_proof(pk, tis, inst, proof, reader=PoseidonReader(proof, spec), decide=decide)
    assert not ref.verify_proof(pk, (tis + 1) % R, instances, proof, reader=PoseidonReader(proof, spec), decide=decide)
    # wrong Poseidon parameters break the Fiat-Shamir challenges
    assert not ref.verify_proof(pk, tis, instances, proof, reader=PoseidonReader(proof, PO.Spec(5, 8, 57)), decide=decide)


def test_numerator_formulas_match_snark_verifier_expression(setup, golden):
    """Row a4's constraint formulas, term by term: evaluate snark-verifier's OWN expression tree for the quotient numerator (stored
    in the fixture's Protocol) at the proof's evaluation point and challenges, and compare with the numerator the oracle computes
    from its restated halo2 formulas (custom gate, permutation argument, mv-lookup/logUp, y-Horner order)."""
    ref, pk, tis, instances, proof, spec, decide = setup
    dbg = {}
    assert ref.verify_proof(pk, tis, instances, proof, reader=PoseidonReader(proof, spec), decide=decide, dbg=dbg)
    n, x = ref.cs.n, dbg["x"]
    xn = pow(x, n, R)
    omega = ref.dom.omega
    chal = [dbg["theta"], dbg["beta"], dbg["gamma"], dbg["y"]]                 # ch0..ch3 of the Protocol
    ev = dbg["evals"]
    # snark-verifier polynomial numbering: p0-p3 fixed, p4-p6 sigma, p7 instance, p8 advice, p9 m, p10 z, p11 phi, p12 random
    poly_eval = {}
    for c in range(4): poly_eval[(c, 0)] = ev[(H.FIXED, c, 0)]
    for i, v in enumerate(dbg["sigma_evals"]): poly_eval[(4 + i, 0)] = v
    poly_eval[(7, 0)] = ev[(H.INSTANCE, 0, 0)]
    for r in range(4): poly_eval[(8, r)] = ev[(H.ADVICE, 0, r)]
    phi_x, phi_nx, m_x = dbg["lookup_evals"][0]
    poly_eval[(9, 0)] = m_x
    poly_eval[(10, 0)], poly_eval[(10, 1)] = dbg["z_evals"][0][0], dbg["z_evals"][0][1]
    poly_eval[(11, 0)], poly_eval[(11, 1)] = phi_x, phi_nx
    poly_eval[(12, 0)] = dbg["random_eval"]

    def lagrange(i):
        wi = pow(omega, i % n, R)
        return (xn - 1) * wi % R * pow(n * (x - wi) % R, -1, R) % R

    def ev_expr(e):
        (k, v), = e.items()
        if k == "Constant": return P.from_mont(P.from_limbs(v), R)
        if k == "Polynomial": return poly_eval[(v["poly"], v["rotation"])]
        if k == "Challenge": return chal[v]
        if k == "CommonPolynomial":
            if v == "Identity": return x
            return lagrange(v["Lagrange"])
        if k == "Negated": return (-ev_expr(v)) % R
        if k == "Sum": return (ev_expr(v[0]) + ev_expr(v[1])) % R
        if k == "Product": return ev_expr(v[0]) * ev_expr(v[1]) % R
        if k == "Scaled": return ev_expr(v[0]) * P.from_mont(P.from_limbs(v[1]), R) % R
        if k == "DistributePowers":
            exprs, base = v
            b, acc = ev_expr(base), 0
            for sub in exprs: acc = (acc * b + ev_expr(sub)) % R
            return acc
        raise ValueError(k)
    assert ev_expr(golden["quotient_numerator"]) == dbg["numerator"]


def test_fixture_instance_accumulator_passes_pairing(setup):
    """The reference's own sanity check (aggregator/src/core.rs:75-89,129-140): the KZG accumulator carried in the first 12 instance
    cells -- [lhs.x, lhs.y, rhs.x, rhs.y] as 3 limbs of 88 bits each (fe_to_limbs, LIMBS = 3, BITS = 88) -- satisfies
    e(lhs, g2) == e(rhs, [s]g2).  Independent of the transcript and of the PLONK verifier."""
    ref, pk, tis, instances, proof, spec, decide = setup
    limbs = instances[0][:12]
    assert all(v < (1 << 88) for v in limbs)
    coords = [limbs[3 * i] + (limbs[3 * i + 1] << 88) + (limbs[3 * i + 2] << 176) for i in range(4)]
    lhs, rhs = (coords[0], coords[1]), (coords[2], coords[3])
    assert all(c < P.Q_MOD for c in coords) and P.g1_is_on_curve(lhs) and P.g1_is_on_curve(rhs)
    assert decide(lhs, rhs)
    assert not decide(P.g1_add(lhs, P.G1_GEN), rhs)
