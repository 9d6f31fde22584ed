"""Synthetic, SATISFIABLE stand-ins with the prover-relevant SHAPE of the reference's benchmark circuits.

The reference's circuits cannot be synthesised without Rust, so the large parity tests and the proof benchmarks use constraint
systems with the same shape parameters, read off the reference:

  keccak_shape(k=17)   packed-multi Keccak (circuit-benchmarks/src/packed_multi_keccak.rs:72-87): one hot cell column queried at
                       the 56 rotations -48..+23 (keccak_circuit.rs:139,469,549,647; => 58 blinding factors, 59 unusable rows,
                       keccak_packed_multi.rs:59-68), five two-column lookup tables of 59 049 / 65 536 / 46 656 / 78 125 rows
                       (normalize_3/4/6, chi_base at k = 17: keccak_circuit/table.rs:20-29, util.rs:228-236), O(10^2) lookup
                       input sets chunked three per argument (cs.chunk_lookups(), degree 9), a second phase for the RLC columns,
                       no instance column.
  super_shape(k=20, A) SuperCircuit (circuit-benchmarks/src/super_circuit.rs:117-132): three phases (zkevm-circuits/src/
                       util.rs:120-133), A advice columns (EVM step width 154: evm_circuit/param.rs:10), O(10^3) gate
                       polynomials of the form condition * constraint, O(10^2) permutation columns, O(10^2) lookup input sets,
                       one instance column of 32 byte cells (pi_circuit.rs:1908-1921) that also takes part in the permutation.

Gate polynomials: base columns hold random values; a defined column D_s = X_s(rot r_s) * Y_s + Z_s is pinned by
q * (D_s - def_s); every further gate is q * h_j * (D_s - def_s) with a fresh "condition" h_j over other cells and rotations
(degree 5), the shape of the EVM circuit's  selector * constraint  products -- satisfied on every active row, not the zero
polynomial, and sharing the (D_s - def_s) sub-expression like a compiled circuit does.

Witness columns are produced with the product's own CUDA field kernels (in the real system they come from Rust `synthesize`,
row a10 of SURVEY 8a, and stay on the CPU).  This module is test / benchmark infrastructure, not product code.
"""
import numpy as np

from zkb200.plonk import ConstraintSystem, Expression as E, ADVICE, FIXED, INSTANCE, NEG, ADD, MUL, SCALED, CONST, CHALLENGE

R_MOD = 0x30644E72E131A029B85045B68181585D2833E84879B9709143E1F593F0000001
ROOT_OF_UNITY_28 = 0x03DDB9F5166D18B798865EA93DD31F743215CF6DD39329C8D34F1ED960C37C9C


def bcast(scalar_t, n):
    return scalar_t.expand(n, 4).contiguous()


class DeviceOps:
    """witness arithmetic on the GPU through the product's element-wise field kernels (tests and the GPU arm of bench.py)"""
    device = "cuda"

    def __init__(self):
        from zkb200 import arithmetic as A, poly, params
        self.A, self.poly, self.params = A, poly, params

    def rand(self, n, seed): return self.A.random_fr_dev(n, seed)
    def from_ints(self, v): return self.params.fr_ints_to_dev(v)
    def scalar(self, v): return self.params.fr_scalar_dev(v % R_MOD)
    def mul(self, a, b): return self.A.field_binop_dev(self.A.FR, self.A.OP_MUL, a, b)
    def add(self, a, b): return self.A.field_binop_dev(self.A.FR, self.A.OP_ADD, a, b)
    def powers(self, base_int, n): return self.poly.fr_powers_dev(self.scalar(base_int).cpu().numpy().view(np.uint64)[0], n)


class OracleOps:
    """the same arithmetic on the CPU through the oracle library: used by bench.py's reference arm, where nothing of the product
    may run (values differ from DeviceOps' generator; only the shape matters there)"""
    device = "cpu"

    def __init__(self):
        import oracle_lib
        self.o = oracle_lib.load()

    @staticmethod
    def _np(t): return np.ascontiguousarray(t.numpy()).view(np.uint64)
    @staticmethod
    def _t(a):
        import torch
        return torch.from_numpy(np.ascontiguousarray(a).view(np.int64))

    def rand(self, n, seed):
        rng = np.random.default_rng(seed)
        a = rng.integers(0, 1 << 64, size=(n, 4), dtype=np.uint64)
        a[:, 3] = rng.integers(0, 0x30644E72E131A029, size=n, dtype=np.uint64)
        return self._t(a)

    def from_ints(self, v):
        z = np.zeros((v.shape[0], 4), dtype=np.uint64)
        z[:, 0] = v.numpy().astype(np.uint64)
        return self._t(self.o.fr_from_canonical(z))

    def scalar(self, v):
        v %= R_MOD
        z = np.array([[(v >> (64 * i)) & 0xFFFFFFFFFFFFFFFF for i in range(4)]], dtype=np.uint64)
        return self._t(self.o.fr_from_canonical(z))

    def mul(self, a, b): return self._t(self.o.fr_mul(self._np(a), self._np(b)))
    def add(self, a, b): return self._t(self.o.fr_add(self._np(a), self._np(b)))
    def powers(self, base_int, n): return self._t(self.o.fr_powers(self._np(self.scalar(base_int))[0], n))

KECCAK_TABLE_ROWS = (59049, 65536, 46656, 78125, 59049)  # normalize_3, normalize_4, normalize_6, chi_base, (second normalize_3 use)


def _degree(e):
    if e.op in (CONST, CHALLENGE): return 0
    if e.op in (FIXED, ADVICE, INSTANCE): return 1
    if e.op in (NEG, SCALED): return _degree(e.a)
    if e.op == ADD: return max(_degree(e.a), _degree(e.b))
    return _degree(e.a) + _degree(e.b)


class _Lcg:
    """tiny deterministic generator for structural choices (column / rotation picks); not used for field values"""
    def __init__(self, seed): self.s = (seed * 2654435761 + 12345) & 0xFFFFFFFF
    def next(self, m):
        self.s = (self.s * 1664525 + 1013904223) & 0xFFFFFFFF
        return (self.s >> 8) % m


class ShapedCircuit:
    def __init__(self, k, *, n_base=8, n_defined=8, n_gates=32, hot_rots=(0, 1, -1, 2), cold_rots=(0, 1, -1, 2), table_rows=(), table_width=2,
                 n_lookup_args=0, pairs_per_table=2, lookup_rots=(0, 1, 2), n_perm=9, phases=2, instance_cells=0, seed=1, ops=None):
        import torch
        ops = ops or DeviceOps()
        self.ops = ops
        assert 1 <= phases <= 3 and n_defined >= 1 and n_base >= 3
        self.k, self.n = k, 1 << k
        n = self.n
        rng = _Lcg(seed)
        T = len(table_rows)
        W = table_width
        L = n_lookup_args if T else 0
        self.L, self.P, self.phases, self.seed = L, n_perm, phases, seed
        # ---- column plan
        # fixed: 0 q_gate, 1 q_lk, 2 q_pi, then T tables x W columns
        nf = 3 + T * W
        c_base0 = 0
        c_def0 = n_base
        c_pair0 = n_base + n_defined                       # T * pairs_per_table pairs of W columns (lookup inputs)
        n_pairs = T * pairs_per_table
        c_perm0 = c_pair0 + n_pairs * W
        c_pi = c_perm0 + n_perm                            # advice cell column constrained to the instance column
        c_f = c_pi + (1 if instance_cells else 0)          # phase-1 RLC column
        c_g = c_f + (1 if phases >= 2 else 0)              # phase-2 column
        na = c_g + (1 if phases >= 3 else 0)
        self.na = na
        adv_phase = [0] * na
        if phases >= 2: adv_phase[c_f] = 1
        if phases >= 3: adv_phase[c_g] = 2
        ch_phase = [0, 1][: phases - 1]
        self.c_def0, self.c_pair0, self.c_perm0, self.c_pi, self.c_f, self.c_g = c_def0, c_pair0, c_perm0, c_pi, c_f, c_g
        n_plain = n_base + n_defined                        # columns gate conditions may read
        q_gate = E.Fixed(0)
        # ---- definitions: D_s = X(rot) * Y + Z over columns with smaller index
        defs = []
        for s in range(n_defined):
            lim = n_base + s
            x, y, z = rng.next(lim), rng.next(lim), rng.next(lim)
            r = cold_rots[rng.next(len(cold_rots))]
            defs.append((x, r, y, z))
        def_expr = [E.Advice(c_def0 + s) + (-(E.Advice(x, r) * E.Advice(y) + E.Advice(z))) for s, (x, r, y, z) in enumerate(defs)]
        gates = [q_gate * def_expr[s] for s in range(n_defined)]
        # ---- condition gates: q * h_j * (D_s - def_s); the hot column 0 walks through every rotation of hot_rots
        hot_i = 0
        for j in range(max(0, n_gates - n_defined)):
            s = j % n_defined
            # the first len(hot_rots) conditions read the hot column (every rotation once), later ones every other gate
            u, v, w = 0 if (j < len(hot_rots) or j % 2 == 0) else rng.next(n_plain), rng.next(n_plain), rng.next(n_plain)
            if u == 0:
                ru = hot_rots[hot_i % len(hot_rots)]; hot_i += 1
            else:
                ru = cold_rots[rng.next(len(cold_rots))]
            rv = cold_rots[rng.next(len(cold_rots))] if v != 0 else 0
            rw = cold_rots[rng.next(len(cold_rots))] if w != 0 else 0
            h = E.Advice(u, ru) * E.Advice(v, rv) + E.Advice(w, rw)
            gates.append(q_gate * (h * def_expr[s]))
        if instance_cells:
            gates.append(E.Fixed(2) * (E.Advice(c_pi) + (-E.Instance(0))))
        if phases >= 2:
            gates.append(q_gate * (E.Advice(c_f) + (-(E.Challenge(0) * E.Advice(0) * E.Advice(1)))))
        if phases >= 3:
            gates.append(q_gate * (E.Advice(c_g) + (-(E.Challenge(1) * E.Advice(c_f) + E.Advice(2)))))
        # ---- lookups: argument l looks into table l mod T with three input sets (pairs of that table at a rotation)
        lookups = []
        self.lk_plan = []
        max_lrot = max([abs(r) for r in lookup_rots] + [0])
        for l in range(L):
            t = l % T
            sets = []
            for i in range(3):
                pr = t * pairs_per_table + ((l // T) * 3 + i) % pairs_per_table
                rot = lookup_rots[(l * 3 + i) % len(lookup_rots)]
                sets.append((pr, rot))
            self.lk_plan.append((t, sets))
            ins = [[E.Fixed(1) * E.Advice(c_pair0 + pr * W + cc, rot) for cc in range(W)] for pr, rot in sets]
            lookups.append((ins, [E.Fixed(3 + t * W + cc) for cc in range(W)]))
        # ---- queries in order of first use (halo2 ConstraintSystem::query_*), blinding factors, degree
        aq, fq, iq = [], [], []
        aset, fset, iset = set(), set(), set()
        seen = set()

        def collect(e):
            if id(e) in seen: return
            seen.add(id(e))
            if e.op == ADVICE:
                if (e.a, e.b) not in aset: aset.add((e.a, e.b)); aq.append((e.a, e.b))
            elif e.op == FIXED:
                if (e.a, e.b) not in fset: fset.add((e.a, e.b)); fq.append((e.a, e.b))
            elif e.op == INSTANCE:
                if (e.a, e.b) not in iset: iset.add((e.a, e.b)); iq.append((e.a, e.b))
            elif e.op in (NEG, SCALED): collect(e.a)
            elif e.op in (ADD, MUL): collect(e.a); collect(e.b)
        for g in gates: collect(g)
        for ins, tb in lookups:
            for inp in ins:
                for e in inp: collect(e)
            for e in tb: collect(e)
        perm_columns = [(ADVICE, c_perm0 + j) for j in range(n_perm)] + ([(INSTANCE, 0)] if instance_cells else [])
        for (t_, i_) in perm_columns:
            q, qs = {ADVICE: (aq, aset), FIXED: (fq, fset), INSTANCE: (iq, iset)}[t_]
            if (i_, 0) not in qs: qs.add((i_, 0)); q.append((i_, 0))
        per_col = {}
        for c, _ in aq: per_col[c] = per_col.get(c, 0) + 1
        bf = max(3, max(per_col.values())) + 2
        degree = 3
        for g in gates: degree = max(degree, _degree(g))
        for ins, tb in lookups:
            degree = max(degree, 4, 2 + sum(max(_degree(e) for e in inp) for inp in ins) + max(max(_degree(e) for e in tb), 1))
        cs = ConstraintSystem(k, nf, na, 1 if instance_cells else 0, adv_phase, ch_phase, bf, degree)
        cs.gates, cs.lookups, cs.perm_columns = gates, lookups, perm_columns
        cs.advice_queries, cs.fixed_queries, cs.instance_queries = aq, fq, iq
        self.cs, self.bf = cs, bf
        usable = n - (bf + 1)
        self.usable = usable
        assert usable > max(table_rows, default=0), "the largest table does not fit the usable rows"
        # ---- values (on ops.device)
        dev = ops.device
        gen = torch.Generator(device=dev)
        gen.manual_seed(seed)
        one = ops.scalar(1)
        zero4 = torch.zeros((n, 4), dtype=torch.int64, device=dev)
        rows = torch.arange(n, device=dev)
        is_usable = rows < usable
        q = torch.where(is_usable[:, None], bcast(one, n), zero4)
        q_lk = torch.where(((rows >= max_lrot) & (rows < usable - max_lrot))[:, None], bcast(one, n), zero4)
        q_pi = torch.where((rows < instance_cells)[:, None], bcast(one, n), zero4)
        self.fixed = [q, q_lk, q_pi]
        tbl_int = []
        for t in range(T):
            cols_t = []
            for cc in range(W):
                v = torch.where(rows < table_rows[t], rows * (7 * cc + 1) + cc * (rows > 0), torch.zeros_like(rows))  # row 0 = all zeros
                cols_t.append(v)
                self.fixed.append(ops.from_ints(v))
            tbl_int.append(cols_t)

        def blind(col, s):
            col[usable:] = ops.rand(bf + 1, seed * 1000 + s)
            return col
        mul, add = ops.mul, ops.add
        adv = [None] * na
        for c in range(n_base):
            adv[c] = ops.rand(n, seed * 7919 + c)
        for s, (x, r, y, z) in enumerate(defs):
            xr = torch.roll(adv[x], -r, dims=0).contiguous() if r else adv[x]
            adv[c_def0 + s] = blind(add(mul(xr, adv[y]), adv[z]), 10 + s)
        for pr in range(n_pairs):
            t = pr // pairs_per_table
            idx = torch.randint(0, table_rows[t], (n,), device=dev, generator=gen)
            for cc in range(W):
                adv[c_pair0 + pr * W + cc] = blind(ops.from_ints(tbl_int[t][cc][idx]), 5000 + pr * W + cc)
        # copy columns: c_j[r] = c_0[pi_j(r)] on usable rows; sigma links the cells holding the same c_0 cell in a cycle
        Pn = n_perm
        base = ops.rand(n, seed * 7 + 3)
        pis = [torch.arange(usable, device=dev)] + [torch.randperm(usable, device=dev, generator=gen) for _ in range(Pn - 1)]
        for j in range(Pn):
            col = base.clone()
            col[:usable] = base[pis[j]]
            adv[c_perm0 + j] = blind(col, 200 + j)
        self.instances = []
        if instance_cells:
            vals = torch.randint(0, 256, (instance_cells,), device=dev, generator=gen)
            inst = ops.from_ints(vals)
            self.instances = [inst]
            col = ops.rand(n, seed * 7 + 11)
            col[:instance_cells] = inst
            adv[c_pi] = blind(col, 300)
        self.adv0 = adv
        Wp = ops.powers(pow(ROOT_OF_UNITY_28, 1 << (28 - k), R_MOD), n)   # omega_k^i
        delta = ops.scalar(pow(7, 1 << 28, R_MOD))                      # DELTA = 7^(2^28)
        dpow = [one]
        for _ in range(len(perm_columns)): dpow.append(mul(dpow[-1], delta))
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
            self.sigma.append(mul(Wp[tgt_rows].contiguous(), col_scale))
        if instance_cells:   # the instance column takes part in the permutation argument with no copies: sigma = delta^P * omega^r
            self.sigma.append(mul(Wp, bcast(dpow[Pn], n)))
        nsets = (len(perm_columns) + (degree - 2) - 1) // (degree - 2)
        self.nsets = nsets
        self.z_blinds = ops.rand(max(1, nsets * bf), seed * 7 + 4)[: nsets * bf]
        self.phi_blinds = ops.rand(max(1, L * bf), seed * 7 + 5)[: L * bf]
        self.random_poly = ops.rand(n, seed * 7 + 6)
        self.transcript_repr = ops.rand(1, seed * 7 + 7)[0]
        self.shape = {"k": k, "advice_columns": na, "fixed_columns": nf, "instance_columns": 1 if instance_cells else 0, "phases": phases,
                      "gates": len(gates), "lookup_arguments": L, "lookup_input_sets": 3 * L, "lookup_tables": [int(x) for x in table_rows],
                      "lookup_width": W, "permutation_columns": len(perm_columns), "cs_degree": degree, "blinding_factors": bf,
                      "advice_queries": len(aq), "fixed_queries": len(fq), "distinct_rotations": len({r for _, r in aq})}

    def synthesize_dev(self, phase, challenges):
        """-> {advice column: device tensor} for the columns of `phase` (challenges: {idx: numpy uint64[4]})."""
        import torch
        ops = self.ops
        mul = ops.mul
        out = {}
        chd = lambda i: bcast(torch.from_numpy(np.ascontiguousarray(challenges[i]).view(np.int64)).to(ops.device).reshape(1, 4), self.n)
        if phase == 0:
            for c, a in enumerate(self.adv0):
                if a is not None: out[c] = a
        elif phase == 1:
            v = mul(mul(self.adv0[0], self.adv0[1]), chd(0))
            v[self.usable:] = ops.rand(self.bf + 1, self.seed * 1000 + 999)
            self._f = v
            out[self.c_f] = v
        elif phase == 2:
            v = ops.add(mul(self._f, chd(1)), self.adv0[2])
            v[self.usable:] = ops.rand(self.bf + 1, self.seed * 1000 + 998)
            out[self.c_g] = v
        return out

    def host(self, t):
        return t.cpu().numpy().view(np.uint64)


def keccak_shape(k=17, seed=3, scale=1.0, ops=None):
    """KeccakCircuit-like shape; `scale` < 1 shrinks the counts (not the rotation set) for the oracle-sized parity tests."""
    rots56 = tuple(range(-48, 8))                        # 56 distinct rotations on the hot column -> 58 blinding factors (59 unusable rows)
    tables = KECCAK_TABLE_ROWS if k >= 17 else tuple(min(r, (1 << k) // 3) for r in KECCAK_TABLE_ROWS)
    sc = lambda v, lo: max(lo, int(round(v * scale)))
    n_def = sc(40, 3)
    return ShapedCircuit(k, n_base=sc(24, 4), n_defined=n_def, n_gates=max(sc(220, 60), n_def + 64), hot_rots=rots56, cold_rots=(0, 1, -12, 2, -1, 5, -24, 7),
                         table_rows=tables, table_width=2, n_lookup_args=sc(35, 5), pairs_per_table=sc(4, 2), lookup_rots=(0, 1, 2, 11, 12, -12),
                         n_perm=sc(12, 8), phases=2, instance_cells=0, seed=seed, ops=ops)


def super_shape(k=20, advice=256, seed=5, scale=1.0, n_gates=None, ops=None):
    """SuperCircuit-like shape with about `advice` advice columns, three phases, an instance column of 32 byte cells."""
    sc = lambda v, lo: max(lo, int(round(v * scale)))
    tables = (256, 65536, 1 << 12, 50000) if k >= 17 else (min(256, (1 << k) // 4), (1 << k) // 3, (1 << k) // 5, (1 << k) // 4)
    n_perm = sc(max(16, advice * 3 // 8), 9)
    pairs = sc(max(2, advice // 32), 2)
    width = 2
    used = len(tables) * pairs * width + n_perm + 1 + 2
    rest = max(8, int(advice * scale) - used)
    n_base = max(4, rest * 2 // 5)
    n_def = max(3, rest - n_base)
    ng = n_gates if n_gates is not None else sc(max(256, advice * 5), 80)
    return ShapedCircuit(k, n_base=n_base, n_defined=n_def, n_gates=ng, hot_rots=(0, 1, -1, 2, 3, -2), cold_rots=(0, 1, -1, 2),
                         table_rows=tables, table_width=width, n_lookup_args=sc(max(8, advice // 8), 4), pairs_per_table=pairs,
                         lookup_rots=(0, 1, -1), n_perm=n_perm, phases=3, instance_cells=32, seed=seed, ops=ops)
