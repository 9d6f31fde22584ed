"""Pure-Python big-integer micro-oracle for the BN254 hot path (TEST INFRASTRUCTURE ONLY).

This file is part of the parity oracle.  Only tests/, __graft_entry__.smoke() and
bench.py's cpu_baseline leg may import it.  It restates, with Python ints, the *mathematical
definitions* that halo2curves 0.1.0 (scroll-tech/halo2curves @ a495a7b, src/bn256/{fr,fq,curve}.rs)
and halo2_proofs 1.1.0 (scroll-tech/halo2 @ e5ddf67, src/arithmetic.rs) implement, so that the C
restatement (oracle/zk_oracle.c) and the CUDA product can both be checked against an independent
third implementation on small sizes.  Constants are those verified against the in-tree fixture
aggregator/data/batch-task.json.
"""
R_MOD = 21888242871839275222246405745257275088548364400416034343698204186575808495617  # Fr
Q_MOD = 21888242871839275222246405745257275088696311157297823662689037894645226208583  # Fq
MONT_R = 1 << 256
FR_S = 28
FR_GENERATOR = 7
FR_ROOT_OF_UNITY = pow(FR_GENERATOR, (R_MOD - 1) >> FR_S, R_MOD)
FR_DELTA = pow(FR_GENERATOR, 1 << FR_S, R_MOD)
FR_ZETA = pow(FR_GENERATOR, (R_MOD - 1) // 3, R_MOD)
G1_B = 3
G1_GEN = (1, 2)


def to_mont(a, p):
    return (a * MONT_R) % p


def from_mont(a, p):
    return (a * pow(MONT_R, -1, p)) % p


def limbs(a):
    """256-bit int -> 4 little-endian u64 limbs (halo2curves in-memory layout)."""
    return [(a >> (64 * i)) & 0xFFFFFFFFFFFFFFFF for i in range(4)]


def from_limbs(l):
    return sum(int(v) << (64 * i) for i, v in enumerate(l))


def omega(k):
    """Generator of the 2^k domain: ROOT_OF_UNITY^(2^(S-k)) (halo2 EvaluationDomain::new)."""
    return pow(FR_ROOT_OF_UNITY, 1 << (FR_S - k), R_MOD)


def ntt_naive(a, w):
    """out[k] = sum_j a[j] w^(jk)  -- definition of best_fft's result (O(n^2), small n only)."""
    n = len(a)
    return [sum(a[j] * pow(w, j * k, R_MOD) for j in range(n)) % R_MOD for k in range(n)]


def ntt(a, w):
    """Recursive radix-2 NTT, natural in / natural out."""
    n = len(a)
    if n == 1:
        return list(a)
    e = ntt(a[0::2], w * w % R_MOD)
    o = ntt(a[1::2], w * w % R_MOD)
    out = [0] * n
    t = 1
    for i in range(n // 2):
        x = o[i] * t % R_MOD
        out[i] = (e[i] + x) % R_MOD
        out[i + n // 2] = (e[i] - x) % R_MOD
        t = t * w % R_MOD
    return out


# ---- G1 (affine tuples, None = identity) ---------------------------------------------------
def g1_is_on_curve(P):
    if P is None:
        return True
    x, y = P
    return (y * y - x * x * x - G1_B) % Q_MOD == 0


def g1_add(P, Q):
    if P is None:
        return Q
    if Q is None:
        return P
    x1, y1 = P
    x2, y2 = Q
    if x1 == x2:
        if (y1 + y2) % Q_MOD == 0:
            return None
        lam = 3 * x1 * x1 * pow(2 * y1, -1, Q_MOD) % Q_MOD
    else:
        lam = (y2 - y1) * pow(x2 - x1, -1, Q_MOD) % Q_MOD
    x3 = (lam * lam - x1 - x2) % Q_MOD
    y3 = (lam * (x1 - x3) - y1) % Q_MOD
    return (x3, y3)


def g1_neg(P):
    return None if P is None else (P[0], (-P[1]) % Q_MOD)


def g1_mul(P, s):
    s %= R_MOD
    acc = None
    while s:
        if s & 1:
            acc = g1_add(acc, P)
        P = g1_add(P, P)
        s >>= 1
    return acc


def msm_naive(scalars, points):
    acc = None
    for s, P in zip(scalars, points):
        acc = g1_add(acc, g1_mul(P, s))
    return acc


def g1_compress(P):
    """halo2curves derive/curve.rs compressed form: 32 B LE x, bit 6 of byte 31 = y & 1;
    identity = 32 zero bytes.  Verified on the fixture's vk/proof points."""
    if P is None:
        return bytes(32)
    x, y = P
    b = bytearray(x.to_bytes(32, "little"))
    b[31] |= (y & 1) << 6
    return bytes(b)


def fq_sqrt(a):
    # q = 3 mod 4
    r = pow(a, (Q_MOD + 1) // 4, Q_MOD)
    return r if r * r % Q_MOD == a % Q_MOD else None


def g1_decompress(b):
    b = bytearray(b)
    sign = (b[31] >> 6) & 1
    assert b[31] >> 7 == 0
    b[31] &= 0x3F
    x = int.from_bytes(b, "little")
    if x == 0 and sign == 0:
        return None
    y = fq_sqrt((x * x * x + G1_B) % Q_MOD)
    assert y is not None, "not on curve"
    if (y & 1) != sign:
        y = Q_MOD - y
    return (x, y)


def fr_from_u512(b64):
    """Fr::from_uniform_bytes / from_u512: 64 LE bytes reduced mod r."""
    return int.from_bytes(b64, "little") % R_MOD
