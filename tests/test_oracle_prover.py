"""CPU: the oracle's restated create_proof produces proofs its own (upstream-equation) verifier accepts, and rejects
tampered proofs / unsatisfied witnesses.  This is the reference the CUDA prover is compared against byte for byte."""
import numpy as np
import pytest

import halo2_ref as H
from circuits import ToyCircuit, ThinCompressionShape, GatesOnlyCircuit


def prove(tc, srs_s=1234):
    ref = H.Ref(tc.cs, srs_s)
    F = ref.F
    pk = ref.keygen([F.arr(c) for c in tc.fixed_ints], tc.copies)
    blinds = {"z": tc.blinds_ints["z"], "phi": tc.blinds_ints["phi"], "random_poly": F.arr(tc.blinds_ints["random_poly"])}
    synth = lambda phase, ch: {c: F.arr(v) for c, v in tc.advice_ints(phase, ch).items()}
    proof, dbg = ref.create_proof(pk, tc.transcript_repr, tc.instances, synth, blinds)
    return ref, pk, proof, dbg


@pytest.mark.parametrize("k,kw", [(5, {}), (6, dict(two_phase=False)), (6, dict(lookups=False, extra_perm=False)), (7, {})])
def test_prove_verify(k, kw):
    tc = ToyCircuit(k, seed=k, **kw)
    ref, pk, proof, dbg = prove(tc)
    # the lookup grand sums close: phi[last usable] == 0
    assert all(v == 0 for v in dbg["phi_last"])
    # quotient has the right degree: numerator vanishes on the domain => h has n*(d-1) coefficients and we kept all
    assert ref.verify_proof(pk, tc.transcript_repr, tc.instances, proof)
    # structure: commitments + evals + 2 opening points
    cs = tc.cs
    nsets = len(dbg["zs"])
    npoints = cs.num_advice + 2 * len(cs.lookups) + nsets + 1 + ref.dom.qdeg + 2
    nevals = len(cs.advice_queries) + len(cs.fixed_queries) + 1 + len(cs.perm_columns) + (3 * nsets - 1) + 3 * len(cs.lookups)
    assert len(proof) == 32 * (npoints + nevals)


def test_reject_tampered_proof_and_instance():
    tc = ToyCircuit(5, seed=42)
    ref, pk, proof, _ = prove(tc)
    assert ref.verify_proof(pk, tc.transcript_repr, tc.instances, proof)
    bad = bytearray(proof)
    bad[32 * 20 + 3] ^= 1
    try:
        ok = ref.verify_proof(pk, tc.transcript_repr, tc.instances, bytes(bad))
    except AssertionError:
        ok = False
    assert not ok
    inst = [list(tc.instances[0])]
    inst[0][0] = (inst[0][0] + 1) % H.R
    assert not ref.verify_proof(pk, tc.transcript_repr, inst, proof)


def test_reject_unsatisfied_witness():
    tc = ToyCircuit(5, seed=43)
    tc.tamper()
    ref, pk, proof, _ = prove(tc)
    assert not ref.verify_proof(pk, tc.transcript_repr, tc.instances, proof)


def test_thin_compression_shape_matches_fixture_layout(golden):
    """Prove the reference's thin compression constraint system (from the fixture's Protocol) on a synthetic witness: the proof
    must have the fixture's layout -- 28 items: 1 advice, 1 m, z, phi, random, 4 h pieces, 17 evaluations, 2 opening points --
    and the evaluation order recorded in the fixture."""
    tc = ThinCompressionShape(7, seed=5)
    ref, pk, proof, dbg = prove(tc)
    assert ref.verify_proof(pk, tc.transcript_repr, tc.instances, proof)
    assert len(proof) == len(bytes.fromhex(golden["proof_hex"])) == 28 * 32
    assert ref.dom.qdeg == golden["quotient_num_chunk"] and ref.bf == 6
    # evaluation order: snark-verifier polynomial numbering p0-p3 fixed (table, constants, q_gate, q_lookup), p4-p6 sigma,
    # p8 advice, p9 m, p10 z, p11 phi, p12 random
    cs = tc.cs
    order = [(8, r) for (_, r) in cs.advice_queries] + [(c, r) for (c, r) in cs.fixed_queries] + [(12, 0)] + [(4, 0), (5, 0), (6, 0)] \
        + [(10, 0), (10, 1)] + [(11, 0), (11, 1), (9, 0)]
    assert order == [(e["poly"], e["rotation"]) for e in golden["evaluations"]]
    # commitments / evaluations decode as points / canonical scalars at the fixture's positions
    import pyref as P
    for i in list(range(9)) + [26, 27]:
        assert P.g1_is_on_curve(P.g1_decompress(proof[32 * i: 32 * i + 32]))
    for i in range(9, 26):
        assert int.from_bytes(proof[32 * i: 32 * i + 32], "little") < P.R_MOD


def test_gates_only_circuit():
    tc = GatesOnlyCircuit(5, seed=9)
    ref, pk, proof, _ = prove(tc)
    assert ref.verify_proof(pk, tc.transcript_repr, tc.instances, proof)


@pytest.mark.parametrize("kind,k", [("toy", 5), ("thin", 6)])
def test_prove_verify_with_poseidon_transcript(kind, k):
    """create_proof with the SDK's Poseidon transcript (what gen_snark_shplonk uses); the Poseidon restatement and the verifier are
    pinned by the reference's own proof in tests/test_fixture_proof.py."""
    tc = (ToyCircuit if kind == "toy" else ThinCompressionShape)(k, seed=70 + k)
    ref = H.Ref(tc.cs, 1234)
    F = ref.F
    pk = ref.keygen([F.arr(c) for c in tc.fixed_ints], tc.copies)
    blinds = {"z": tc.blinds_ints["z"], "phi": tc.blinds_ints["phi"], "random_poly": F.arr(tc.blinds_ints["random_poly"])}
    synth = lambda phase, ch: {c: F.arr(v) for c, v in tc.advice_ints(phase, ch).items()}
    proof, _ = ref.create_proof(pk, tc.transcript_repr, tc.instances, synth, blinds, transcript=H.Ref.PoseidonTranscript(ref))
    assert ref.verify_proof(pk, tc.transcript_repr, tc.instances, proof, reader=H.Ref.PoseidonReader(proof))
    assert not ref.verify_proof(pk, tc.transcript_repr, tc.instances, proof)          # a Blake2b reader must not accept it
    proof_b, _ = ref.create_proof(pk, tc.transcript_repr, tc.instances, synth, blinds)
    assert proof_b != proof and len(proof_b) == len(proof)


def test_keccak_and_evm_transcript():
    """Keccak-f[1600] pinned by hashlib.sha3_256 (same permutation) and by the reference's KECCAK_CODE_HASH_EMPTY; the oracle prover
    and verifier round-trip with snark-verifier's EVM transcript framing (uncompressed big-endian proof items)."""
    import hashlib, random
    import keccak_ref as K
    rnd = random.Random(4)
    for ln in (0, 1, 135, 136, 137, 300):
        data = bytes(rnd.randrange(256) for _ in range(ln))
        assert K.sha3_256(data) == hashlib.sha3_256(data).digest()
    # eth-types/src/lib.rs KECCAK_CODE_HASH_EMPTY
    assert K.keccak256(b"").hex() == "c5d2460186f7233c927e7db2dcc703c0e500b653ca82273b7bfad8045d85a470"
    tc = ThinCompressionShape(6, seed=31)
    ref = H.Ref(tc.cs, 1234)
    F = ref.F
    pk = ref.keygen([F.arr(c) for c in tc.fixed_ints], tc.copies)
    blinds = {"z": tc.blinds_ints["z"], "phi": tc.blinds_ints["phi"], "random_poly": F.arr(tc.blinds_ints["random_poly"])}
    synth = lambda phase, ch: {c: F.arr(v) for c, v in tc.advice_ints(phase, ch).items()}
    proof, _ = ref.create_proof(pk, tc.transcript_repr, tc.instances, synth, blinds, transcript=K.EvmTranscript(ref))
    assert len(proof) == 11 * 64 + 17 * 32                       # uncompressed points, 32-byte scalars
    assert ref.verify_proof(pk, tc.transcript_repr, tc.instances, proof, reader=K.EvmTranscript(proof=proof))
    bad = bytearray(proof); bad[700] ^= 1
    try:
        ok = ref.verify_proof(pk, tc.transcript_repr, tc.instances, bytes(bad), reader=K.EvmTranscript(proof=bytes(bad)))
    except AssertionError:
        ok = False
    assert not ok
