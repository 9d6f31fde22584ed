"""BN254 optimal-ate pairing in plain Python (TEST INFRASTRUCTURE ONLY; see oracle/pyref.py for the rules).

Restates the textbook construction used by the reference's verifier side (halo2curves 0.1.0 src/bn256/{fq2,fq6,fq12,engine}.rs
compute the same bilinear map; pairing values are unique, so any correct implementation agrees): Fq12 = Fq[w]/(w^12 - 18 w^6 + 82)
(so w^6 = 9 + i with i^2 = -1), sextic twist E'(Fq2): y^2 = x^3 + 3/(9+i), Miller loop over 6u+2 with the two Frobenius
corrections, final exponentiation by (q^12 - 1)/r.  Slow (seconds per pairing) -- used only to check KZG accumulators.
"""
import pyref as P

Q = P.Q_MOD
R = P.R_MOD
ATE_LOOP_COUNT = 29793968203157093288          # 6u + 2, u = 4965661367192848881
LOG_ATE = 63
FQ12_MOD = [82, 0, 0, 0, 0, 0, -18 % Q, 0, 0, 0, 0, 0]   # w^12 = 18 w^6 - 82


class FQP:
    """element of Fq[x]/(modulus), modulus monic given by its low coefficients"""
    deg = 0
    mod = None

    def __init__(self, c):
        assert len(c) == self.deg
        self.c = [int(v) % Q for v in c]

    @classmethod
    def one(cls): return cls([1] + [0] * (cls.deg - 1))
    @classmethod
    def zero(cls): return cls([0] * cls.deg)
    def __add__(self, o): return type(self)([a + b for a, b in zip(self.c, o.c)])
    def __sub__(self, o): return type(self)([a - b for a, b in zip(self.c, o.c)])
    def __neg__(self): return type(self)([-a for a in self.c])
    def __eq__(self, o): return self.c == o.c
    def is_zero(self): return not any(self.c)

    def __mul__(self, o):
        if isinstance(o, int):
            return type(self)([a * o for a in self.c])
        d = self.deg
        b = [0] * (2 * d - 1)
        for i, x in enumerate(self.c):
            if x:
                for j, y in enumerate(o.c):
                    b[i + j] += x * y
        for exp in range(2 * d - 2, d - 1, -1):
            top = b[exp] % Q
            if top:
                for i, m in enumerate(self.mod):
                    if m:
                        b[exp - d + i] -= top * m
            b[exp] = 0
        return type(self)(b[:d])

    def __pow__(self, e):
        out, base = type(self).one(), self
        while e:
            if e & 1: out = out * base
            base = base * base
            e >>= 1
        return out

    def inv(self):
        # extended Euclid on polynomials over Fq
        def deg(p):
            d = len(p) - 1
            while d and p[d] % Q == 0: d -= 1
            return d

        def poly_divmod(a, b):
            a = [x % Q for x in a]
            db = deg(b)
            o = [0] * max(1, len(a) - db)
            ib = pow(b[db], -1, Q)
            for i in range(deg(a) - db, -1, -1):
                qv = a[db + i] * ib % Q
                o[i] = qv
                for j in range(db + 1): a[j + i] = (a[j + i] - b[j] * qv) % Q
            return o, a[:max(db, 1)]
        d = self.deg
        lm, hm = [1] + [0] * d, [0] * (d + 1)
        low, high = self.c + [0], [m % Q for m in self.mod] + [1]
        while deg(low):
            r, _ = poly_divmod(high, low)
            r += [0] * (d + 1 - len(r))
            nm, new = list(hm), list(high)
            for i in range(d + 1):
                for j in range(d + 1 - i):
                    nm[i + j] -= lm[i] * r[j]
                    new[i + j] -= low[i] * r[j]
            nm = [x % Q for x in nm]; new = [x % Q for x in new]
            lm, low, hm, high = nm, new, lm, low
        iv = pow(low[0], -1, Q)
        return type(self)([x * iv for x in lm[:d]])

    def __truediv__(self, o): return self * o.inv()


class FQ2(FQP):
    deg = 2
    mod = [1, 0]          # i^2 = -1


class FQ12(FQP):
    deg = 12
    mod = FQ12_MOD


B2 = FQ2([3, 0]) / FQ2([9, 1])
G2 = (FQ2([10857046999023057135944570762232829481370756359578518086990519993285655852781,
           11559732032986387107991004021392285783925812861821192530917403151452391805634]),
      FQ2([8495653923123431417604973247489272438418190587263600148770280649306958101930,
           4082367875863433681332203403145435568316851327593401208105741076214120093531]))
W = FQ12([0, 1] + [0] * 10)


def g2_is_on_curve(pt):
    x, y = pt
    return y * y - x * x * x == B2


def _double(pt):
    x, y = pt
    lam = (x * x * 3) / (y * 2)
    nx = lam * lam - x * 2
    return nx, lam * (x - nx) - y


def _add(p1, p2):
    if p1 is None: return p2
    if p2 is None: return p1
    x1, y1 = p1; x2, y2 = p2
    if x1 == x2:
        return _double(p1) if y1 == y2 else None
    lam = (y2 - y1) / (x2 - x1)
    nx = lam * lam - x1 - x2
    return nx, lam * (x1 - nx) - y1


def g2_mul(pt, k):
    acc = None
    while k:
        if k & 1: acc = _add(acc, pt)
        pt = _double(pt)
        k >>= 1
    return acc


def _twist(pt):
    x, y = pt
    xc = [x.c[0] - x.c[1] * 9, x.c[1]]
    yc = [y.c[0] - y.c[1] * 9, y.c[1]]
    nx = FQ12([xc[0]] + [0] * 5 + [xc[1]] + [0] * 5)
    ny = FQ12([yc[0]] + [0] * 5 + [yc[1]] + [0] * 5)
    return nx * (W ** 2), ny * (W ** 3)


def _cast(pt):
    return FQ12([pt[0]] + [0] * 11), FQ12([pt[1]] + [0] * 11)


def _line(p1, p2, t):
    x1, y1 = p1; x2, y2 = p2; xt, yt = t
    if not (x1 == x2):
        m = (y2 - y1) / (x2 - x1)
        return m * (xt - x1) - (yt - y1)
    if y1 == y2:
        m = (x1 * x1 * 3) / (y1 * 2)
        return m * (xt - x1) - (yt - y1)
    return xt - x1


def miller_loop(q_g2, p_g1):
    """un-exponentiated Miller value f_{6u+2,Q}(P) with the Frobenius corrections; q_g2: affine FQ2 pair, p_g1: (x, y) ints."""
    if q_g2 is None or p_g1 is None: return FQ12.one()
    Qt, Pt = _twist(q_g2), _cast(p_g1)
    Rp, f = Qt, FQ12.one()
    for i in range(LOG_ATE, -1, -1):
        f = f * f * _line(Rp, Rp, Pt)
        Rp = _double(Rp)
        if ATE_LOOP_COUNT & (1 << i):
            f = f * _line(Rp, Qt, Pt)
            Rp = _add(Rp, Qt)
    Q1 = (Qt[0] ** Q, Qt[1] ** Q)
    nQ2 = (Q1[0] ** Q, -(Q1[1] ** Q))
    f = f * _line(Rp, Q1, Pt)
    Rp = _add(Rp, Q1)
    f = f * _line(Rp, nQ2, Pt)
    return f


def final_exponentiate(f):
    return f ** ((Q ** 12 - 1) // R)


def pairing(q_g2, p_g1):
    return final_exponentiate(miller_loop(q_g2, p_g1))


def pairing_product_is_one(pairs):
    """prod e(P_i, Q_i) == 1 for pairs [(p_g1, q_g2), ...] with a single final exponentiation."""
    f = FQ12.one()
    for p_g1, q_g2 in pairs:
        f = f * miller_loop(q_g2, p_g1)
    return final_exponentiate(f) == FQ12.one()
