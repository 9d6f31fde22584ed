"""GPU parity at SuperCircuit-like SHAPE (degree 9, three chunked lookup input sets, multi-chunk permutation, two phases):
the synthetic WideCircuit proven by the CUDA session vs the CPU oracle, byte for byte; ParamsKZG built on the GPU vs the
oracle's SRS."""
import numpy as np
import pytest

import halo2_ref as H

pytestmark = pytest.mark.gpu


def to_oracle_cs(z):
    from zkb200 import plonk as Z
    F = H.FA()

    def conv(e):
        op = e.op
        if op == Z.CONST: return H.const(F.ints(np.array([e.a], dtype=np.uint64))[0])
        if op == Z.FIXED: return H.fixed(e.a, e.b)
        if op == Z.ADVICE: return H.advice(e.a, e.b)
        if op == Z.INSTANCE: return H.instance(e.a, e.b)
        if op == Z.CHALLENGE: return H.challenge(e.a)
        if op == Z.NEG: return -conv(e.a)
        if op == Z.ADD: return conv(e.a) + conv(e.b)
        if op == Z.MUL: return conv(e.a) * conv(e.b)
        if op == Z.SCALED: return H.scaled(conv(e.a), F.ints(np.array([e.b], dtype=np.uint64))[0])
        raise ValueError
    cs = H.ConstraintSystem(z.k, z.num_fixed, z.num_advice, z.num_instance, z.advice_phase, z.challenge_phase)
    cs.gates = [conv(g) for g in z.gates]
    cs.lookups = [H.Lookup([[conv(e) for e in inp] for inp in ins], [conv(e) for e in tb]) for ins, tb in z.lookups]
    cs.perm_columns = list(z.perm_columns)
    cs.finalize()
    return cs


def test_params_setup_matches_oracle(oracle):
    from zkb200.params import ParamsKZG
    k, s = 8, 1234
    p = ParamsKZG.unsafe_setup_with_s(k, s)
    ref = H.Ref(H.ConstraintSystem(k, 0, 1, 0).finalize(), s)
    assert (p.g.cpu().numpy().view(np.uint64) == ref.g).all()
    assert (p.g_lagrange.cpu().numpy().view(np.uint64) == ref.g_lagrange).all()


@pytest.mark.parametrize("k,kw", [(8, dict(n_gates=6, n_lookups=1, n_perm=9)), (10, dict(n_gates=9, n_lookups=2, n_perm=10)),
                                  (9, dict(n_gates=4, n_lookups=1, n_perm=3, two_phase=False))])
def test_wide_circuit_matches_oracle(k, kw):
    from zkb200 import plonk as Z
    from wide_circuit import WideCircuit
    from zkb200.params import ParamsKZG
    wc = WideCircuit(k, seed=k, **kw)
    cs = to_oracle_cs(wc.cs)
    ref = H.Ref(cs, 4321)
    assert ref.bf == wc.bf and ref.d == wc.cs.degree
    assert cs.advice_queries == wc.cs.advice_queries and cs.fixed_queries == wc.cs.fixed_queries
    F = ref.F
    h = wc.host
    fixed = [h(t) for t in wc.fixed]
    sigma = [h(t) for t in wc.sigma]
    # the oracle's keygen must reproduce the generator's sigma from the copy constraints? -> it takes sigma as given:
    pkr = {"fixed_values": fixed, "fixed_polys": [ref.lagrange_to_coeff(v) for v in fixed], "sigma_values": sigma,
           "sigma_polys": [ref.lagrange_to_coeff(v) for v in sigma]}
    n, bf = wc.n, wc.bf
    l0 = np.zeros((n, 4), dtype=np.uint64); l0[0] = ref.w_arr(1)
    lb = np.zeros((n, 4), dtype=np.uint64); lb[n - bf:] = ref.w_arr(1)
    ll = np.zeros((n, 4), dtype=np.uint64); ll[n - bf - 1] = ref.w_arr(1)
    pkr["l0"], pkr["l_last"], pkr["l_blind"] = [ref.lagrange_to_coeff(v) for v in (l0, ll, lb)]
    pkr["fixed_commitments"] = [ref.commit_lagrange(v) for v in fixed]
    pkr["sigma_commitments"] = [ref.commit_lagrange(v) for v in sigma]
    nsets = (wc.P + ref.chunk - 1) // ref.chunk
    zb, pb = h(wc.z_blinds), h(wc.phi_blinds)
    blinds = {"z": [F.ints(zb[i * bf:(i + 1) * bf]) for i in range(nsets)], "phi": [F.ints(pb[i * bf:(i + 1) * bf]) for i in range(wc.L)],
              "random_poly": h(wc.random_poly)}
    trep = F.ints(h(wc.transcript_repr[None]))[0]

    def synth_ref(phase, ch):
        chm = {i: F.arr([v])[0] for i, v in ch.items()}
        return {c: h(t) for c, t in wc.synthesize_dev(phase, chm).items()}
    proof_ref, dbg = ref.create_proof(pkr, trep, [], synth_ref, blinds)
    assert all(v == 0 for v in dbg["phi_last"])
    assert ref.verify_proof(pkr, trep, [], proof_ref)

    params = ParamsKZG.unsafe_setup_with_s(k, 4321)
    pk = Z.ProvingKey(wc.cs, fixed, sigma, h(params.g), h(params.g_lagrange))
    synth = lambda phase, ch: {c: h(t) for c, t in wc.synthesize_dev(phase, ch).items()}
    proof = Z.create_proof(pk, h(wc.transcript_repr[None])[0], [], synth, zb, pb, h(wc.random_poly))
    assert proof == proof_ref
