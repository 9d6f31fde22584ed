import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "oracle")); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np, torch
import oracle_lib
from zkb200 import arithmetic as A
from util import rand_field, to_dev, to_host
orc = oracle_lib.load()
a = rand_field(8, 1)
got = to_host(A.field_unop_dev(0, 0, to_dev(a)))
exp = orc.fr_inv(a)
print("inv got", got[:2]); print("inv exp", exp[:2])
# check a * got == 1 ?
print("a*got", orc.fr_mul(a, got)[:2])
sq = to_host(A.field_unop_dev(0, 3, to_dev(a)))
print("sqr ok", (sq == orc.fr_mul(a, a)).all())
import pyref as P
def um(x): return P.from_mont(P.from_limbs(x), P.R_MOD)
av = [um(x) for x in a]
for op, f in ((5, lambda v: pow(v,3,P.R_MOD)), (6, lambda v: pow(v,65537,P.R_MOD)), (7, lambda v: pow(v,P.R_MOD-2,P.R_MOD)), (8, lambda v: v)):
    g = to_host(A.field_unop_dev(0, op, to_dev(a)))
    print("op", op, [um(x) == f(v) for x, v in zip(g, av)][:4], g[0])
