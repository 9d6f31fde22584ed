"""Pin the CPU oracle: (1) against the reference's own fixture (tests/golden/thin_chunk_proof.json, extracted from
aggregator/data/batch-task.json by tests/golden/make_golden.py); (2) against an independent big-integer
implementation (oracle/pyref.py).  CPU only."""
import random
import numpy as np
import pyref as P


def mont_arr(vals, p):
    return np.array([P.limbs(P.to_mont(v, p)) for v in vals], dtype=np.uint64).reshape(-1, 4)


def unmont(limbs4, p):
    return P.from_mont(P.from_limbs(limbs4), p)


# ------------------------------------------------------------------ fixture-pinned facts
def test_fixture_montgomery_constants(oracle, golden):
    one, delta, delta2 = [np.array(c, dtype=np.uint64) for c in golden["numerator_constants_mont_limbs"]]
    got_one = oracle.fr_from_canonical(np.array([[1, 0, 0, 0]], dtype=np.uint64))[0]
    assert (got_one == one).all()
    seven = oracle.fr_from_canonical(np.array([[7, 0, 0, 0]], dtype=np.uint64))[0]
    d = oracle.fr_pow(seven, 1 << 28)
    assert (d == delta).all()
    assert (oracle.fr_mul(d[None], d[None])[0] == delta2).all()


def test_fixture_domain(oracle, golden):
    dom = golden["domain"]
    k = dom["k"]
    w = oracle.fr_omega(k)
    assert (w == np.array(dom["gen"], dtype=np.uint64)).all()
    winv = oracle.fr_inv(w[None])[0]
    assert (winv == np.array(dom["gen_inv"], dtype=np.uint64)).all()
    n = oracle.fr_from_canonical(np.array([[dom["n"], 0, 0, 0]], dtype=np.uint64))
    assert (oracle.fr_inv(n)[0] == np.array(dom["n_inv"], dtype=np.uint64)).all()
    # omega has exact order 2^k
    assert unmont(oracle.fr_pow(w, 1 << k), P.R_MOD) == 1
    assert unmont(oracle.fr_pow(w, 1 << (k - 1)), P.R_MOD) == P.R_MOD - 1


def test_fixture_point_compression(oracle, golden):
    vk = bytes.fromhex(golden["vk_hex"])
    assert int.from_bytes(vk[0:4], "big") == golden["domain"]["k"]
    nfixed = int.from_bytes(vk[4:8], "big")
    assert nfixed == 4 and len(golden["preprocessed"]) == 7
    for i, pp in enumerate(golden["preprocessed"]):
        aff = np.array(pp["x"] + pp["y"], dtype=np.uint64)
        assert oracle.g1_is_on_curve(aff)
        assert oracle.g1_compress(aff) == vk[8 + 32 * i: 40 + 32 * i]


def test_fixture_proof_layout(golden):
    proof = bytes.fromhex(golden["proof_hex"])
    assert len(proof) == 28 * 32
    nw = sum(golden["num_witness"]) + golden["quotient_num_chunk"]
    for i in list(range(nw)) + [26, 27]:
        pt = P.g1_decompress(proof[32 * i: 32 * i + 32])
        assert P.g1_is_on_curve(pt) and P.g1_compress(pt) == proof[32 * i: 32 * i + 32]
    for i in range(nw, nw + len(golden["evaluations"])):
        assert int.from_bytes(proof[32 * i: 32 * i + 32], "little") < P.R_MOD
    assert nw + len(golden["evaluations"]) + 2 == 28


# ------------------------------------------------------------------ big-integer cross-checks
def test_field_ops_vs_bigint(oracle):
    rnd = random.Random(1)
    for which, p in ((0, P.R_MOD), (1, P.Q_MOD)):
        a = [rnd.randrange(p) for _ in range(300)] + [0, 1, p - 1, p - 1]
        b = [rnd.randrange(p) for _ in range(300)] + [p - 1, p - 1, p - 1, 1]
        A, B = mont_arr(a, p), mont_arr(b, p)
        for op, f in ((0, lambda x, y: (x + y) % p), (1, lambda x, y: (x - y) % p), (2, lambda x, y: x * y % p)):
            assert (oracle.field_binop(which, op, A, B) == mont_arr([f(x, y) for x, y in zip(a, b)], p)).all()
        assert (oracle.field_unop(which, 0, A) == mont_arr([pow(x, -1, p) if x else 0 for x in a], p)).all()
        assert (oracle.field_unop(which, 2, A) == np.array([P.limbs(x) for x in a], dtype=np.uint64)).all()


def test_from_u512(oracle):
    rnd = random.Random(2)
    raw = bytes(rnd.randrange(256) for _ in range(64 * 50)) + b"\xff" * 64 + bytes(64)
    got = oracle.fr_from_u512(raw)
    exp = mont_arr([P.fr_from_u512(raw[64 * i: 64 * i + 64]) for i in range(len(raw) // 64)], P.R_MOD)
    assert (got == exp).all()


def test_best_fft_vs_definition(oracle):
    rnd = random.Random(3)
    for k in range(0, 11):
        n = 1 << k
        a = [rnd.randrange(P.R_MOD) for _ in range(n)]
        w = oracle.fr_omega(k)
        assert unmont(w, P.R_MOD) == P.omega(k)
        got = oracle.best_fft(mont_arr(a, P.R_MOD), w, k)
        exp = P.ntt_naive(a, P.omega(k)) if k <= 6 else P.ntt(a, P.omega(k))
        assert (got == mont_arr(exp, P.R_MOD)).all(), k


def test_best_fft_roundtrip(oracle):
    k = 14
    rng = np.random.default_rng(4)
    a = rng.integers(0, 1 << 60, size=(1 << k, 4), dtype=np.uint64)
    w = oracle.fr_omega(k)
    f = oracle.best_fft(a, w, k)
    b = oracle.best_fft(f, oracle.fr_inv(w[None])[0], k)
    ninv = oracle.fr_inv(oracle.fr_from_canonical(np.array([[1 << k, 0, 0, 0]], dtype=np.uint64)))
    b = oracle.fr_mul(b, np.repeat(ninv, 1 << k, axis=0))
    assert (a == b).all()


def test_best_multiexp_vs_bigint(oracle):
    rnd = random.Random(5)
    G = oracle.g1_generator()
    for n in (1, 2, 3, 5, 31, 32, 100, 257):
        ks = [rnd.randrange(P.R_MOD) for _ in range(n)]
        bases = oracle.g1_fixed_base_mul(G, mont_arr(ks, P.R_MOD))
        pts = [P.g1_mul(P.G1_GEN, k) for k in ks]
        for i in range(n):
            assert (unmont(bases[i, :4], P.Q_MOD), unmont(bases[i, 4:], P.Q_MOD)) == pts[i]
        sc = [rnd.randrange(P.R_MOD) for _ in range(n)]
        if n > 3:
            sc[0], sc[1], sc[2] = 0, 1, P.R_MOD - 1
        exp = P.msm_naive(sc, pts)
        for th in (1, 3, 8):
            aff = oracle.g1_to_affine(oracle.best_multiexp(mont_arr(sc, P.R_MOD), bases, th))
            assert (unmont(aff[:4], P.Q_MOD), unmont(aff[4:], P.Q_MOD)) == exp
            assert oracle.g1_compress(aff) == P.g1_compress(exp)


def test_multiexp_degenerate(oracle):
    G = oracle.g1_generator()
    n = 64
    bases = np.repeat(G[None], n, axis=0)          # repeated bases -> doubling path inside buckets
    ones = mont_arr([1] * n, P.R_MOD)
    aff = oracle.g1_to_affine(oracle.best_multiexp(ones, bases, 2))
    assert (unmont(aff[:4], P.Q_MOD), unmont(aff[4:], P.Q_MOD)) == P.g1_mul(P.G1_GEN, n)
    zeros = np.zeros((n, 4), dtype=np.uint64)
    assert oracle.g1_compress(oracle.g1_to_affine(oracle.best_multiexp(zeros, bases, 2))) == bytes(32)
    bases[::2] = 0                                   # identity bases = (0,0)
    aff = oracle.g1_to_affine(oracle.best_multiexp(ones, bases, 2))
    assert (unmont(aff[:4], P.Q_MOD), unmont(aff[4:], P.Q_MOD)) == P.g1_mul(P.G1_GEN, n // 2)
    # P + (-P) inside one bucket
    neg = G.copy()
    neg[4:] = mont_arr([P.Q_MOD - 2], P.Q_MOD)[0]
    b2 = np.stack([G, neg])
    assert oracle.g1_compress(oracle.g1_to_affine(oracle.best_multiexp(mont_arr([5, 5], P.R_MOD), b2, 1))) == bytes(32)
