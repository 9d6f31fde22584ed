"""GPU parity for the polynomial utilities (SURVEY 8a rows a5-a9 building blocks) vs big-integer definitions."""
import numpy as np
import pytest

import pyref as P
from util import rand_field, to_dev, to_host

pytestmark = pytest.mark.gpu
R = P.R_MOD


def ints(oracle, a):
    return [P.from_limbs(r) for r in oracle.fr_to_canonical(np.ascontiguousarray(a))]


def mont(oracle, vals):
    return oracle.fr_from_canonical(np.array([P.limbs(v % R) for v in vals], dtype=np.uint64).reshape(-1, 4))


@pytest.mark.parametrize("n", [1, 7, 8, 2047, 2048, 2049, 70000, (1 << 20) + 5])
def test_powers(oracle, n):
    from zkb200 import poly
    base = rand_field(1, 5)[0]
    got = to_host(poly.fr_powers_dev(base, n))
    assert (got == oracle.fr_powers(base, n)).all()


@pytest.mark.parametrize("n", [1, 5, 2048, 2049, 5000, 1 << 16, (1 << 20) + 3])
def test_eval_polynomial(oracle, n):
    from zkb200 import poly
    polys = [rand_field(n, 40 + i) for i in range(3)]
    x = rand_field(1, 77)[0]
    got = poly.eval_polynomial_dev([to_dev(p) for p in polys], x)
    xi = ints(oracle, x[None])[0]
    for p, g in zip(polys, got):
        acc = 0
        for c in reversed(ints(oracle, p)):
            acc = (acc * xi + c) % R
        assert ints(oracle, g[None])[0] == acc


@pytest.mark.parametrize("n", [1, 9, 2048, 2049, 600000, (1 << 20) + 1])
def test_prefix_scans(oracle, n):
    from zkb200 import poly
    a = rand_field(n, 3)
    init = rand_field(1, 4)[0]
    gp = to_host(poly.prefix_product_dev(to_dev(a), init))
    gs = to_host(poly.prefix_sum_dev(to_dev(a), init))
    ai = ints(oracle, a)
    cur_p = cur_s = ints(oracle, init[None])[0]
    ep, es = [], []
    for v in ai:
        ep.append(cur_p); es.append(cur_s)
        cur_p = cur_p * v % R
        cur_s = (cur_s + v) % R
    assert (gp == mont(oracle, ep)).all()
    assert (gs == mont(oracle, es)).all()


@pytest.mark.parametrize("n", [2, 9, 2048, 2049, 4097, 600001, (1 << 20)])
def test_kate_division(oracle, n):
    from zkb200 import poly
    a = rand_field(n, 8)
    u = rand_field(1, 9)[0]
    got = to_host(poly.kate_division_dev(to_dev(a), u))
    ai = ints(oracle, a)
    ui = ints(oracle, u[None])[0]
    q = [0] * n
    tmp = 0
    for i in range(n - 1, 0, -1):
        tmp = (ai[i] + tmp * ui) % R
        q[i - 1] = tmp
    assert (got == mont(oracle, q)).all()
