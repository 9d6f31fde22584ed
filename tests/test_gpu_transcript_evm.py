"""GPU parity for create_proof under snark-verifier's EVM transcript (gen_evm_proof_shplonk, prover/src/common/prover/evm.rs:67):
the CUDA session with transcript kind 2 must emit the same bytes as the oracle prover driven by oracle/keccak_ref.EvmTranscript
(uncompressed big-endian points, big-endian scalars, Keccak-256 challenges), and the oracle verifier must accept them."""
import numpy as np
import pytest

import halo2_ref as H
import keccak_ref as K
from circuits import ToyCircuit, ThinCompressionShape
from test_gpu_prover import to_product_cs

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("kind,k", [("toy", 6), ("thin", 8)])
def test_create_proof_evm_transcript_matches_oracle(kind, k):
    from zkb200 import plonk as Z
    tc = (ToyCircuit if kind == "toy" else ThinCompressionShape)(k, seed=300 + k)
    ref = H.Ref(tc.cs, 1234)
    F = ref.F
    fixed = [F.arr(c) for c in tc.fixed_ints]
    pkr = ref.keygen(fixed, tc.copies)
    rp = F.arr(tc.blinds_ints["random_poly"])
    blinds = {"z": tc.blinds_ints["z"], "phi": tc.blinds_ints["phi"], "random_poly": rp}
    synth_ref = lambda phase, ch: {c: F.arr(v) for c, v in tc.advice_ints(phase, ch).items()}
    proof_ref, _ = ref.create_proof(pkr, tc.transcript_repr, tc.instances, synth_ref, blinds, transcript=K.EvmTranscript(ref))
    pk = Z.ProvingKey(to_product_cs(tc.cs, ref.bf, ref.d), fixed, pkr["sigma_values"], ref.g, ref.g_lagrange)
    synth = lambda phase, ch: {c: F.arr(v) for c, v in tc.advice_ints(phase, {i: F.ints(v[None])[0] for i, v in ch.items()}).items()}
    zb = np.concatenate([F.arr(b) for b in tc.blinds_ints["z"]]); pb = np.concatenate([F.arr(b) for b in tc.blinds_ints["phi"]])
    proof = Z.create_proof(pk, F.arr([tc.transcript_repr])[0], [F.arr(c) for c in tc.instances], synth, zb, pb, rp, transcript="evm")
    assert len(proof) == len(proof_ref)
    diff = next((i // 32 for i in range(0, len(proof), 32) if proof[i: i + 32] != proof_ref[i: i + 32]), None)
    assert diff is None, f"first differing 32-byte word of the proof: {diff}"
    assert ref.verify_proof(pkr, tc.transcript_repr, tc.instances, proof, reader=K.EvmTranscript(proof=proof))
