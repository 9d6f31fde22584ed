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
