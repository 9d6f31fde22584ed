"""GPU parity: the CUDA path through the C ABI vs the CPU oracle, bit-exact (integer arithmetic)."""
import numpy as np
import pytest

import pyref as P
from util import rand_field, to_dev, to_host

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def A():
    from zkb200 import arithmetic
    return arithmetic


def edge_values(p):
    vals = [0, 1, 2, p - 1, p - 2, (p - 1) // 2, (1 << 253) % p, (1 << 128) - 1]
    return np.array([P.limbs(P.to_mont(v, p)) for v in vals], dtype=np.uint64)


@pytest.mark.parametrize("field", [0, 1])
def test_field_ops(A, oracle, field):
    p = P.R_MOD if field == 0 else P.Q_MOD
    e = edge_values(p)
    a = np.concatenate([rand_field(100000, 11 + field), e, e[::-1]])
    b = np.concatenate([rand_field(100000, 13 + field), e, e])
    da, db = to_dev(a), to_dev(b)
    for op in (0, 1, 2):
        got = to_host(A.field_binop_dev(field, op, da, db))
        assert (got == oracle.field_binop(field, op, a, b)).all(), f"binop {op}"
    for uop in (1, 2, 3, 4):
        got = to_host(A.field_unop_dev(field, uop, da))
        exp = oracle.field_unop(field, uop, a) if uop != 1 else None
        if uop == 1:
            # canonical -> Montgomery needs canonical inputs < p : use the canonical forms
            can = oracle.field_unop(field, 2, a)
            got = to_host(A.field_unop_dev(field, 1, to_dev(can)))
            exp = a
        assert (got == exp).all(), f"unop {uop}"
    small = np.concatenate([a[:2000], e])
    got = to_host(A.field_unop_dev(field, 0, to_dev(small)))
    assert (got == oracle.field_unop(field, 0, small)).all()


def test_batch_invert(A, oracle):
    a = rand_field(100003, 21)
    a[::97] = 0
    a[5] = 0
    got = to_host(A.fr_batch_invert_dev(to_dev(a)))
    assert (got == oracle.fr_inv(a)).all()


@pytest.mark.parametrize("log_n", list(range(0, 15)) + [16, 17, 20, 21, 22, 23])
def test_ntt_vs_oracle(A, oracle, log_n):
    a = rand_field(1 << log_n, 100 + log_n)
    w, wi = A.root_of_unity(log_n)
    assert (w == oracle.fr_omega(log_n)).all()
    got = A.best_fft(a.copy(), w, log_n)
    assert (got == oracle.best_fft(a, w, log_n)).all()
    got = A.best_fft(a.copy(), wi, log_n)
    assert (got == oracle.best_fft(a, wi, log_n)).all()


@pytest.mark.parametrize("log_n", [3, 12, 13, 18])
def test_ntt_scale_and_coset(A, oracle, log_n):
    n = 1 << log_n
    a = rand_field(n, 7 + log_n)
    w, wi = A.root_of_unity(log_n)
    ninv = oracle.fr_inv(oracle.fr_from_canonical(np.array([[n, 0, 0, 0]], dtype=np.uint64)))[0]
    zeta = oracle.fr_from_canonical(np.array([P.limbs(P.FR_ZETA)], dtype=np.uint64))[0]
    zeta2 = oracle.fr_mul(zeta[None], zeta[None])[0]
    one = oracle.fr_from_canonical(np.array([[1, 0, 0, 0]], dtype=np.uint64))[0]
    # coeff_to_extended-style: input i scaled by ZETA^(i mod 3), forward transform
    zp = np.stack([one, zeta, zeta2])[np.arange(n) % 3]
    exp = oracle.best_fft(oracle.fr_mul(a, zp), w, log_n)
    assert (A.best_fft(a.copy(), w, log_n, coset_zeta=1) == exp).all()
    # extended_to_coeff-style: inverse transform, * n^-1, output k scaled by ZETA^(-(k mod 3))
    zm = np.stack([one, zeta2, zeta])[np.arange(n) % 3]
    exp = oracle.fr_mul(oracle.fr_mul(oracle.best_fft(a, wi, log_n), np.repeat(ninv[None], n, axis=0)), zm)
    assert (A.best_fft(a.copy(), wi, log_n, scale=ninv, coset_zeta=2) == exp).all()
    # plain lagrange_to_coeff: inverse * n^-1 ; round trip restores the input
    f = A.best_fft(a.copy(), w, log_n)
    assert (A.best_fft(f, wi, log_n, scale=ninv) == a).all()


def test_ntt_2_24_roundtrip_and_linearity(A, oracle):
    """BASELINE config #2 at full size: size-independent properties (round trip, linearity, spot values)."""
    import torch
    log_n = 24
    n = 1 << log_n
    w, wi = A.root_of_unity(log_n)
    ninv = oracle.fr_inv(oracle.fr_from_canonical(np.array([[n, 0, 0, 0]], dtype=np.uint64)))[0]
    a = A.random_fr_dev(n, 1234)
    b = A.random_fr_dev(n, 4321)
    fa = A.best_fft_dev(a.clone(), w, log_n)
    fb = A.best_fft_dev(b.clone(), w, log_n)
    fab = A.best_fft_dev(A.field_binop_dev(0, 0, a, b), w, log_n)
    assert torch.equal(fab, A.field_binop_dev(0, 0, fa, fb))
    back = A.best_fft_dev(fa.clone(), wi, log_n, scale=ninv)
    assert torch.equal(back, a)
    # spot check a few outputs against the definition  X[k] = sum_j a[j] w^(jk), evaluated by Horner on the oracle side
    # through a 2^24-point oracle transform of a sparse input (delta at j0): X[k] = a[j0] * w^(j0 k)
    j0 = 123457
    d = torch.zeros_like(a)
    d[j0] = a[j0]
    fd = to_host(A.best_fft_dev(d, w, log_n))
    aj = to_host(a[j0:j0 + 1])[0]
    for k in (0, 1, 2, 77777, n - 1):
        wk = oracle.fr_pow(w, (j0 * k) % n)
        assert (fd[k] == oracle.fr_mul(aj[None], wk[None])[0]).all()


@pytest.mark.parametrize("log_n", [24, 25, 26])
def test_ntt_full_size_vs_oracle(A, oracle, log_n):
    """BASELINE config #2 (2^24) and the sizes of the three-pass plan (2^25, 2^26: config #5's domain) compared ELEMENT BY ELEMENT
    with the oracle's best_fft restatement -- forward, and inverse with the fused 1/n."""
    import torch
    n = 1 << log_n
    w, wi = A.root_of_unity(log_n)
    ninv = oracle.fr_inv(oracle.fr_from_canonical(np.array([[n, 0, 0, 0]], dtype=np.uint64)))[0]
    a = A.random_fr_dev(n, 9000 + log_n)
    ha = to_host(a)
    f = A.best_fft_dev(a.clone(), w, log_n)
    exp = oracle.best_fft(ha, w, log_n)
    assert (to_host(f) == exp).all()
    del exp
    back = A.best_fft_dev(f, wi, log_n, scale=ninv)
    assert torch.equal(back, a)
    if log_n == 24:
        inv = A.best_fft_dev(a.clone(), wi, log_n, scale=ninv)
        exp = oracle.fr_mul(oracle.best_fft(ha, wi, log_n), np.repeat(ninv[None], n, axis=0))
        assert (to_host(inv) == exp).all()


def test_ntt_batch_columns(A, oracle):
    """several columns through one launch per pass (how the prover transforms a stage's columns), distinct buffers"""
    from zkb200 import poly as Pz
    log_n = 14
    w, _ = A.root_of_unity(log_n)
    cols = [rand_field(1 << log_n, 640 + i) for i in range(5)]
    got = Pz.ntt_batch_dev([to_dev(c) for c in cols], w, log_n)
    for c, g in zip(cols, got):
        assert (to_host(g) == oracle.best_fft(c, w, log_n)).all()


def make_bases(A, oracle, n, seed):
    """n distinct affine points [k_i] G computed on the GPU (checked against the oracle on a prefix)."""
    k = rand_field(n, seed)
    G = oracle.g1_generator()
    bases = to_host(A.g1_fixed_base_mul_dev(G, to_dev(k)))
    m = min(n, 64)
    assert (bases[:m] == oracle.g1_fixed_base_mul(G, k[:m])).all()
    return bases


@pytest.mark.parametrize("n", [0, 1, 2, 5, 33, 1000, 1 << 14, (1 << 16) + 3])
def test_msm_vs_oracle(A, oracle, n):
    if n == 0:
        r = A.best_multiexp(np.zeros((0, 4), dtype=np.uint64), np.zeros((0, 8), dtype=np.uint64))
        assert r.compressed == bytes(32)
        return
    bases = make_bases(A, oracle, n, 300 + n)
    s = rand_field(n, 500 + n)
    if n > 4:
        s[0] = 0
        s[1] = oracle.fr_from_canonical(np.array([[1, 0, 0, 0]], dtype=np.uint64))[0]
        s[2] = oracle.fr_from_canonical(np.array([P.limbs(P.R_MOD - 1)], dtype=np.uint64))[0]
        bases[3] = 0  # identity base
    r = A.best_multiexp(s, bases)
    exp_aff = oracle.g1_to_affine(oracle.best_multiexp(s, bases))
    assert (r.affine == exp_aff).all()
    assert r.compressed == oracle.g1_compress(exp_aff)
    assert oracle.g1_is_on_curve(r.affine)


def test_msm_degenerate(A, oracle):
    G = oracle.g1_generator()
    n = 4096
    bases = np.repeat(G[None], n, axis=0).copy()      # all points equal: doubling inside buckets
    ones = np.repeat(oracle.fr_from_canonical(np.array([[1, 0, 0, 0]], dtype=np.uint64)), n, axis=0)
    r = A.best_multiexp(ones, bases)
    assert (r.affine == oracle.g1_to_affine(oracle.best_multiexp(ones, bases))).all()
    assert A.best_multiexp(np.zeros((n, 4), dtype=np.uint64), bases).compressed == bytes(32)
    neg = bases.copy()
    neg[1::2, 4:] = oracle.field_unop(1, 4, bases[1::2, 4:].copy())   # P, -P, P, -P ... -> identity
    assert A.best_multiexp(ones, neg).compressed == bytes(32)


def test_msm_fixture_points_compress(A, oracle, golden):
    """1 * P for the fixture's preprocessed commitments must compress to the reference vk bytes."""
    vk = bytes.fromhex(golden["vk_hex"])
    one = oracle.fr_from_canonical(np.array([[1, 0, 0, 0]], dtype=np.uint64))
    for i, pp in enumerate(golden["preprocessed"]):
        base = np.array([pp["x"] + pp["y"]], dtype=np.uint64)
        r = A.best_multiexp(one, base)
        assert r.compressed == vk[8 + 32 * i: 40 + 32 * i]


def test_msm_2_20_vs_oracle(A, oracle):
    """BASELINE config #1: 2^20-point MSM, bit-exact compressed commitment vs the CPU oracle."""
    n = 1 << 20
    k = A.random_fr_dev(n, 99)
    bases_t = A.g1_fixed_base_mul_dev(oracle.g1_generator(), k)
    s_t = A.random_fr_dev(n, 100)
    r = A.best_multiexp_dev(s_t, bases_t)
    exp = oracle.g1_to_affine(oracle.best_multiexp(to_host(s_t), to_host(bases_t)))
    assert r.compressed == oracle.g1_compress(exp)
    assert (r.affine == exp).all()


def test_msm_batch_vs_single(A, oracle):
    """Batched multi-column MSM (how a phase's advice columns are committed) == one MSM per column, incl. skewed columns."""
    import torch
    n = 1 << 13
    bases = make_bases(A, oracle, n, 4242)
    cols = [rand_field(n, 900 + i) for i in range(5)]
    cols[1][:] = 0                                   # all-zero column -> identity commitment
    cols[2][:, 1:] = 0; cols[2][:, 0] &= np.uint64(3)   # tiny Montgomery limbs pattern (structured column)
    one = oracle.fr_from_canonical(np.array([[1, 0, 0, 0]], dtype=np.uint64))[0]
    cols[3][:] = one                                 # every scalar equal: one giant bucket per window (skew path)
    cols[3][::7] = 0
    bt = to_dev(bases)
    got = A.best_multiexp_batch_dev([to_dev(c) for c in cols], bt)
    for c, g in zip(cols, got):
        exp = oracle.g1_to_affine(oracle.best_multiexp(c, bases))
        assert (g == exp).all()
    # a batch larger than the internal per-pass limit
    many = [to_dev(rand_field(n, 2000 + i)) for i in range(70)]
    got = A.best_multiexp_batch_dev(many, bt)
    for i in (0, 33, 69):
        assert (got[i] == A.best_multiexp_dev(many[i], bt).affine).all()


@pytest.mark.parametrize("k,j", [(6, 3), (9, 5), (11, 9)])
def test_evaluation_domain_vs_oracle(oracle, k, j):
    """EvaluationDomain::{lagrange_to_coeff, coeff_to_extended, extended_to_coeff} mirror vs the oracle's restatement."""
    import halo2_ref as H
    from zkb200.domain import EvaluationDomain
    ref = H.Ref(H.ConstraintSystem(k, 0, 1, 0).finalize(), 99)
    ref.dom = H.Domain(k, j)
    dom = EvaluationDomain(j, k)
    assert dom.extended_k == ref.dom.extended_k
    a = rand_field(1 << k, 31 + k)
    coeff = dom.lagrange_to_coeff(to_dev(a))
    assert (to_host(coeff) == ref.lagrange_to_coeff(a)).all()
    ext = dom.coeff_to_extended(coeff)
    assert (to_host(ext) == ref.coeff_to_extended(to_host(coeff))).all()
    back = dom.extended_to_coeff(ext)
    exp = ref.extended_to_coeff(to_host(ext))
    assert (to_host(back) == exp).all()
    assert (to_host(back)[: 1 << k] == to_host(coeff)).all() and not to_host(back)[1 << k:].any()


def test_msm_2_23_linearity(A, oracle):
    """BASELINE config #5 shard size (2^26 / 8 GPUs = 2^23 points per GPU): size-independent property
    MSM(a + b) == MSM(a) + MSM(b), plus MSM(0) == identity and a spot value from a sparse scalar vector."""
    import torch
    from zkb200 import parallel
    n = 1 << 23
    bases = A.g1_fixed_base_mul_dev(oracle.g1_generator(), A.random_fr_dev(n, 5150))
    a, b = A.random_fr_dev(n, 1), A.random_fr_dev(n, 2)
    ra, rb = A.best_multiexp_dev(a, bases), A.best_multiexp_dev(b, bases)
    rab = A.best_multiexp_dev(A.field_binop_dev(0, 0, a, b), bases)
    s, comp = parallel.g1_sum_affine(np.stack([ra.affine, rb.affine]))
    assert (rab.affine == s).all() and rab.compressed == comp
    assert oracle.g1_is_on_curve(rab.affine)
    sparse = torch.zeros_like(a)
    idx = [3, 77777, n - 1]
    for i in idx: sparse[i] = a[i]
    r = A.best_multiexp_dev(sparse, bases)
    hb, ha = to_host(bases), to_host(a)
    exp = oracle.g1_to_affine(oracle.best_multiexp(np.stack([ha[i] for i in idx]), np.stack([hb[i] for i in idx])))
    assert (r.affine == exp).all()


def test_msm_reduction_levels_decided_on_device(A, oracle):
    """The reduction levels after the chunked level 0 are decided on the device (the host never learns the bucket sizes).  Random
    scalars: buckets hold ~32 entries (two or three 32-entry chunks), but the TOP window sees only the two leading bits of a
    254-bit scalar, so four buckets collect n/4 entries each -> 512 partials -> two 64-way levels at n = 2^16.  All-equal scalars
    put all n entries of every window into one bucket (2048 partials): also two levels.  n = 2^11 needs one level at most for
    either.  Results are compared with the oracle in every case."""
    from zkb200 import default_context
    ctx = default_context()
    levels = lambda: ctx.lib.zkb_msm_last_levels(ctx.handle)
    n = 1 << 16
    bases = make_bases(A, oracle, n, 777)
    bt = to_dev(bases)
    rnd = rand_field(n, 778)
    r = A.best_multiexp_dev(to_dev(rnd), bt)
    lv_random = levels()
    assert (r.affine == oracle.g1_to_affine(oracle.best_multiexp(rnd, bases))).all()
    same = np.repeat(rand_field(1, 779), n, axis=0)
    r2 = A.best_multiexp_dev(to_dev(same), bt)
    lv_skew = levels()
    assert (r2.affine == oracle.g1_to_affine(oracle.best_multiexp(same, bases))).all()
    assert lv_random == 2 and lv_skew == 2
    m = 1 << 11
    r3 = A.best_multiexp_dev(to_dev(same[:m]), bt[:m])
    assert levels() == 1     # 2048 entries in one bucket -> 64 partials -> one level
    assert (r3.affine == oracle.g1_to_affine(oracle.best_multiexp(same[:m], bases[:m]))).all()


@pytest.mark.parametrize("k", [6, 10])
def test_srs_handle_commit_and_downsize(A, oracle, k):
    """zkb_srs_*: commit / commit_lagrange against the loaded handle == oracle best_multiexp; ParamsKZG::downsize on the device
    (group iFFT of the truncated g) == the oracle's trapdoor construction [L_i(s)] G of the smaller SRS, point for point."""
    import halo2_ref as H
    from zkb200.params import ParamsKZG
    s = 31337
    ref = H.Ref(H.ConstraintSystem(k, 0, 1, 0).finalize(), s)
    params = ParamsKZG(k, ref.g, ref.g_lagrange)
    srs = params.load()
    assert srs.k == k
    v = rand_field(1 << k, 50 + k)
    assert (srs.commit_lagrange(to_dev(v)).affine == ref.commit_lagrange(v)).all()
    assert (srs.commit(to_dev(v)).affine == ref.commit(v)).all()
    short = rand_field((1 << k) - 5, 51 + k)
    assert (srs.commit(to_dev(short)).affine == ref.commit(short)).all()
    # g_lagrange derived on the device from g alone
    srs2 = ParamsKZG(k, ref.g, None).load(derive_lagrange=True)
    assert (srs2.read(1) == ref.g_lagrange).all()
    small = srs.downsize(k - 2)
    ref_small = H.Ref(H.ConstraintSystem(k - 2, 0, 1, 0).finalize(), s)
    assert (small.read(0) == ref_small.g).all()
    assert (small.read(1) == ref_small.g_lagrange).all()
    vs = rand_field(1 << (k - 2), 52 + k)
    assert (small.commit_lagrange(to_dev(vs)).affine == ref_small.commit_lagrange(vs)).all()
