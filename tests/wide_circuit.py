"""(test / benchmark infrastructure, not product code -- moved out of the package after the round-1 review)
Synthetic, satisfiable, degree-9 constraint systems with vectorised witness generation (the small round-1 stand-in; the SURVEY-shaped
ones are in tests/standins.py).

The reference's circuits cannot be synthesised without Rust, so the benchmarks and large parity
tests use stand-ins with the same prover-relevant shape: many advice columns, custom gates with rotations, degree-9
mv-lookup arguments (three chunked input sets per table, as `chunk_lookups()` produces), a multi-chunk permutation, an
optional second phase driven by a challenge (zkevm-circuits/src/util.rs:120-133).  Witness columns are produced with
the product's own CUDA field kernels (in the real system they come from Rust `synthesize`, row a10 of SURVEY 8a).
"""
import numpy as np

from zkb200 import arithmetic as A
from zkb200 import poly
from zkb200.plonk import ConstraintSystem, Expression as E, ADVICE
from zkb200.params import fr_scalar_dev, fr_ints_to_dev, bcast, fr_pow2k_dev


class WideCircuit:
    """fixed: 0 q_gate, 1 q_lk, 2 table.   advice: [w_0 .. w_{G+1}] arithmetic chain, then 3*L lookup inputs, then P copy
    columns, then (optionally) one phase-1 column f.   gate i: q_gate * (w_i(rot r_i) * w_{i+1} + w_i - w_{i+2})."""
    ROTS = [0, 1, -1, 2]

    def __init__(self, k, n_gates=8, n_lookups=2, n_perm=9, two_phase=True, seed=1, table_bits=None):
        import torch
        self.k, self.n = k, 1 << k
        n = self.n
        G, L, Pn = n_gates, n_lookups, n_perm
        self.G, self.L, self.P, self.two_phase = G, L, Pn, two_phase
        n_arith = G + 2
        self.c_lk0 = n_arith
        self.c_perm0 = n_arith + 3 * L
        self.c_f = self.c_perm0 + Pn
        na = self.c_f + (1 if two_phase else 0)
        phases = [0] * na
        if two_phase: phases[self.c_f] = 1
        rots_used = {}
        gates = []
        for i in range(G):
            r = self.ROTS[i % 4]
            gates.append(E.Fixed(0) * (E.Advice(i, r) * E.Advice(i + 1) + E.Advice(i) + (-E.Advice(i + 2))))
        if two_phase:
            gates.append(E.Fixed(0) * (E.Advice(self.c_f) + (-(E.Challenge(0) * E.Advice(0) * E.Advice(1)))))
        lookups = []
        for l in range(L):
            ins = [[E.Fixed(1) * E.Advice(self.c_lk0 + 3 * l + j)] for j in range(3)]
            lookups.append((ins, [E.Fixed(2)]))
        # queries in order of first use; blinding factors = max(3, max distinct rotations per advice column) + 2
        aq, fq = [], []

        def collect(e):
            if e.op == ADVICE and (e.a, e.b) not in aq: aq.append((e.a, e.b))
            elif e.op == 1 and (e.a, e.b) not in fq: fq.append((e.a, e.b))
            elif e.op in (5, 8): collect(e.a)
            elif e.op in (6, 7): collect(e.a); collect(e.b)
        for g in gates: collect(g)
        for ins, tb in lookups:
            for inp in ins:
                for e in inp: collect(e)
            for e in tb: collect(e)
        for j in range(Pn):
            if (self.c_perm0 + j, 0) not in aq: aq.append((self.c_perm0 + j, 0))
        per_col = {}
        for c, _ in aq: per_col[c] = per_col.get(c, 0) + 1
        bf = max(3, max(per_col.values())) + 2
        degree = max(3, 2 + 3 * 2 + 1 if L else 3)           # 9 with lookups (three degree-2 input sets + table)
        cs = ConstraintSystem(k, 3, na, 0, phases, [0] if two_phase else [], bf, degree)
        cs.gates, cs.lookups = gates, lookups
        cs.perm_columns = [(ADVICE, self.c_perm0 + j) for j in range(Pn)]
        cs.advice_queries, cs.fixed_queries, cs.instance_queries = aq, fq, []
        self.cs, self.bf = cs, bf
        usable = n - (bf + 1)
        self.usable = usable
        dev = "cuda"
        gen = torch.Generator(device=dev)
        gen.manual_seed(seed)
        self.seed = seed
        one = fr_scalar_dev(1)
        zero4 = torch.zeros((n, 4), dtype=torch.int64, device=dev)
        rows = torch.arange(n, device=dev)
        is_usable = (rows < usable)
        # ---- fixed columns
        q = torch.where(is_usable[:, None], bcast(one, n), zero4)
        T = min(usable, 1 << (table_bits or max(1, k - 1)))
        tvals = torch.where(rows < T, rows, torch.zeros_like(rows))
        table = fr_ints_to_dev(tvals)
        self.fixed = [q, q.clone(), table]
        # ---- advice, phase 0
        def blind(col, s):
            rnd = A.random_fr_dev(bf + 1, seed * 1000 + s)
            col[usable:] = rnd
            return col
        adv = [None] * na
        adv[0] = A.random_fr_dev(n, seed * 7 + 1)
        adv[1] = A.random_fr_dev(n, seed * 7 + 2)
        for i in range(G):
            r = self.ROTS[i % 4]
            wi_rot = torch.roll(adv[i], -r, dims=0).contiguous()
            v = A.field_binop_dev(A.FR, A.OP_ADD, A.field_binop_dev(A.FR, A.OP_MUL, wi_rot, adv[i + 1]), adv[i])
            adv[i + 2] = blind(v, 10 + i)
        for l in range(L):
            for j in range(3):
                idx = torch.randint(0, T, (n,), device=dev, generator=gen)
                adv[self.c_lk0 + 3 * l + j] = blind(fr_ints_to_dev(idx), 100 + 3 * l + j)
        # ---- copy columns: c_j[r] = c_0[pi_j(r)] on usable rows; sigma links the cells holding the same c_0 cell in a cycle
        base = A.random_fr_dev(n, seed * 7 + 3)
        pis = [torch.arange(usable, device=dev)] + [torch.randperm(usable, device=dev, generator=gen) for _ in range(Pn - 1)]
        for j in range(Pn):
            col = base.clone()
            col[:usable] = base[pis[j]]
            adv[self.c_perm0 + j] = blind(col, 200 + j)
        self.adv0 = adv
        # sigma: cell (c_j, r) with s = pi_j(r) -> (c_{j+1}, inv_{j+1}(s))
        omega, _ = A.root_of_unity(k)
        W = poly.fr_powers_dev(omega, n)
        delta = fr_pow2k_dev(fr_scalar_dev(7), 28)
        dpow = [one]
        for _ in range(Pn): dpow.append(A.field_binop_dev(A.FR, A.OP_MUL, dpow[-1], delta))
        invs = []
        for j in range(Pn):
            inv = torch.empty(usable, dtype=torch.int64, device=dev)
            inv[pis[j]] = torch.arange(usable, device=dev)
            invs.append(inv)
        self.sigma = []
        for j in range(Pn):
            jn = (j + 1) % Pn
            tgt_rows = torch.arange(n, device=dev)
            tgt_rows[:usable] = invs[jn][pis[j]]
            col_scale = torch.where(is_usable[:, None], bcast(dpow[jn], n), bcast(dpow[j], n))
            self.sigma.append(A.field_binop_dev(A.FR, A.OP_MUL, W[tgt_rows].contiguous(), col_scale))
        # ---- blinding scalars, transcript_repr
        nsets = (Pn + (degree - 2) - 1) // (degree - 2)
        self.z_blinds = A.random_fr_dev(max(1, nsets * bf), seed * 7 + 4)[: nsets * bf]
        self.phi_blinds = A.random_fr_dev(max(1, L * bf), seed * 7 + 5)[: L * bf]
        self.random_poly = A.random_fr_dev(n, seed * 7 + 6)
        self.transcript_repr = A.random_fr_dev(1, seed * 7 + 7)[0]

    def synthesize_dev(self, phase, challenges):
        """-> {advice column: device tensor} for the columns of `phase` (challenges: {idx: numpy uint64[4]})."""
        import torch
        out = {}
        if phase == 0:
            for c, a in enumerate(self.adv0):
                if a is not None: out[c] = a
        elif phase == 1 and self.two_phase:
            ch = torch.from_numpy(np.ascontiguousarray(challenges[0]).view(np.int64)).cuda().reshape(1, 4)
            v = A.field_binop_dev(A.FR, A.OP_MUL, A.field_binop_dev(A.FR, A.OP_MUL, self.adv0[0], self.adv0[1]), bcast(ch, self.n))
            v[self.usable:] = A.random_fr_dev(self.bf + 1, self.seed * 1000 + 999)
            out[self.c_f] = v
        return out

    def host(self, t):
        return t.cpu().numpy().view(np.uint64)
