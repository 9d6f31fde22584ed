"""Big-integer model of the bounds the lazy (< 2p) NTT butterflies rely on (csrc/ff.cuh fp_mul_t<REDUCE=false>, fp_add_lazy, fp_sub_lazy,
fp_cond_sub; csrc/ntt.cu ntt_round).  Word-serial CIOS with 32-bit words, R = 2^256:
   running value t < a + p after every word, pre-shift total < 2^288 (fits the 9-limb X and the 8-limb Y<<32 accumulators),
   result < a*b/R + p  ->  < 2p  for a < 4p (first operand, the one multiplied by the words of b) and b < p.
Checked on adversarial (maximal) and random operands for both BN254 fields; plus the butterfly invariants."""
import random

P_FR = 0x30644E72E131A029B85045B68181585D2833E84879B9709143E1F593F0000001
P_FQ = 0x30644E72E131A029B85045B68181585D97816A916871CA8D3C208C16D87CFD47
R = 1 << 256
W = 1 << 32


def cios_lazy(a, b, p):
    inv = (-pow(p, -1, W)) % W          # -p^-1 mod 2^32 (FrParams::INV / FqParams::INV)
    t = 0
    for i in range(8):
        bi = (b >> (32 * i)) & (W - 1)
        m = ((t + a * bi) % W) * inv % W
        pre = t + a * bi + m * p
        assert pre % W == 0 and pre < 1 << 288, "pre-shift total must fit X (9 limbs) + Y << 32 (8 limbs)"
        t = pre >> 32
        assert t < a + p, "running value bound"
    return t


def test_inv_constants_match_ff_cuh():
    assert (-pow(P_FR, -1, W)) % W == 0xEFFFFFFF and (-pow(P_FQ, -1, W)) % W == 0xE4866389


def test_headroom():
    for p in (P_FR, P_FQ):
        assert 4 * p < R and 5 * p < R          # a + p < 5p fits 256 bits for a < 4p


def test_lazy_product_stays_below_2p():
    rnd = random.Random(7)
    for p in (P_FR, P_FQ):
        cases = [(4 * p - 1, p - 1), (4 * p - 1, 1), (2 * p - 1, p - 1), (0, p - 1), (4 * p - 1, 0), (2 * p, p - 1)]
        cases += [(rnd.randrange(4 * p), rnd.randrange(p)) for _ in range(3000)]
        for a, b in cases:
            t = cios_lazy(a, b, p)
            assert t < 2 * p
            assert t % p == a * b * pow(R, -1, p) % p      # it IS the Montgomery product, up to one multiple of p


def test_butterfly_invariants():
    """x, y < 2p  ->  add_lazy < 2p ; sub_lazy in (0, 4p) without a conditional ; after the twiddle multiply (or the conditional
    subtraction of 2p when the twiddle is 1) everything is < 2p again; the final conditional subtraction of p lands in [0, p)"""
    rnd = random.Random(11)
    p = P_FR
    for _ in range(3000):
        x, y, w = rnd.randrange(2 * p), rnd.randrange(2 * p), rnd.randrange(p)
        s = x + y
        assert s < R
        s = s - 2 * p if s >= 2 * p else s
        assert s < 2 * p and s % p == (x + y) % p
        d = ((x - y) % R + 2 * p) % R           # wrap modulo 2^256, then + 2p: what the two carry chains compute
        assert d == x - y + 2 * p and 0 < d < 4 * p
        t = cios_lazy(d, w, p)
        assert t < 2 * p and t % p == (x - y) * w * pow(R, -1, p) % p
        d2 = d - 2 * p if d >= 2 * p else d
        assert d2 < 2 * p and d2 % p == (x - y) % p
        f = t - p if t >= p else t
        assert 0 <= f < p
