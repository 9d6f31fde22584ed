"""CPU check of the NTT kernel's index arithmetic (tests/ntt_tile_model.py mirrors csrc/ntt.cu) against the DFT definition:
single-, two- and three-pass plans, the fused scale / zeta-coset / input-scale variants."""
import random
import pytest

import ntt_tile_model as M


def sampled_check(vals, out, omega, ks, post=lambda k, v: v):
    n = len(vals)
    for k in ks:
        wk = pow(omega, k, M.R_MOD)
        acc, cur = 0, 1
        for j in range(n):
            acc = (acc + vals[j] * cur) % M.R_MOD
            cur = cur * wk % M.R_MOD
        assert out[k] == post(k, acc), f"output {k} differs"


@pytest.mark.parametrize("log_n", list(range(0, 14)))
def test_small_tiles_all_plan_shapes(log_n):
    # tile = 32 elements, final transforms <= 2^5, the others <= 2^3 (2^4 when needed): 1 pass up to 2^5, 2 up to 2^8, 3 above
    rnd = random.Random(log_n)
    n = 1 << log_n
    vals = [rnd.randrange(M.R_MOD) for _ in range(n)]
    w = M.omega_for(log_n)
    out = M.ntt(vals, log_n, w, tile_bits=5, max_bits=5, threads=4, pref_inner_bits=3, max_inner_bits=4)
    ks = range(n) if n <= 64 else [0, 1, n - 1] + [rnd.randrange(n) for _ in range(6)]
    sampled_check(vals, out, w, ks)


@pytest.mark.parametrize("log_n", [4, 9, 12])
def test_fused_variants(log_n):
    rnd = random.Random(100 + log_n)
    n = 1 << log_n
    vals = [rnd.randrange(M.R_MOD) for _ in range(n)]
    w = M.omega_for(log_n)
    zeta = pow(7, (M.R_MOD - 1) // 3, M.R_MOD)
    sc = rnd.randrange(M.R_MOD)
    ins = [rnd.randrange(M.R_MOD) for _ in range(n)]
    ks = [0, 1, 2, n - 1] + [rnd.randrange(n) for _ in range(4)]
    out = M.ntt(vals, log_n, w, 5, 5, 4, pref_inner_bits=3, max_inner_bits=4, scale=sc)
    sampled_check(vals, out, w, ks, lambda k, v: v * sc % M.R_MOD)
    out = M.ntt(vals, log_n, w, 5, 5, 4, pref_inner_bits=3, max_inner_bits=4, coset_zeta=1, zeta=zeta)
    sampled_check([v * pow(zeta, i % 3, M.R_MOD) % M.R_MOD for i, v in enumerate(vals)], out, w, ks)
    out = M.ntt(vals, log_n, w, 5, 5, 4, pref_inner_bits=3, max_inner_bits=4, coset_zeta=2, zeta=zeta, scale=sc)
    sampled_check(vals, out, w, ks, lambda k, v: v * sc % M.R_MOD * pow(zeta, (3 - k % 3) % 3, M.R_MOD) % M.R_MOD)
    out = M.ntt(vals, log_n, w, 5, 5, 4, pref_inner_bits=3, max_inner_bits=4, in_scale=ins)
    sampled_check([v * s % M.R_MOD for v, s in zip(vals, ins)], out, w, ks)


@pytest.mark.parametrize("log_n", [11, 12, 13])
def test_production_constants(log_n):
    # the kernel's own limits (2048-element tiles, 256 threads, inner factors <= 2^9): single pass at 2^11, two passes (6,6) / (7,6) above
    rnd = random.Random(7 + log_n)
    n = 1 << log_n
    vals = [rnd.randrange(M.R_MOD) for _ in range(n)]
    w = M.omega_for(log_n)
    out = M.ntt(vals, log_n, w, tile_bits=11, max_bits=11, threads=256, pref_inner_bits=9, max_inner_bits=9)
    sampled_check(vals, out, w, [0, 1, n // 2, n - 1, rnd.randrange(n), rnd.randrange(n)])


def test_plan_shapes_of_the_kernel_limits():
    # 2^24 (BASELINE config #2) = 8 + 8 + 8; 2^20 = 9 + 11; 2^28 (largest Fr domain) = 9 + 9 + 10; every inner factor <= 2^9
    shapes = {k: M.Plan(k, 11, 11, 9, 9).bits for k in range(0, 29)}
    assert shapes[24] == [8, 8, 8] and shapes[20] == [9, 11] and shapes[28] == [9, 9, 10] and shapes[11] == [11] and shapes[12] == [6, 6]
    assert shapes[21] == [7, 7, 7] and shapes[26] == [9, 9, 8]
    for k, b in shapes.items():
        assert sum(b) == k and (len(b) == 1 or all(x >= 6 for x in b))   # multi-pass factors >= 2^6: C * (A + 1) fits the padded tile
