"""Pure-Python model of the INDEX ARITHMETIC of csrc/ntt.cu (tile decoding, TMA box coordinates, shared-memory layouts L0 / L1,
radix-2^R register rounds, tile-major boundary tables, digit-reversed stores), parametrised by the tile / transform limits so
that the 3-pass plans can be exercised at sizes a big-integer simulation finishes in seconds.  It mirrors the kernel
statement by statement (same names); tests/test_ntt_tile_model.py checks it against the definition of the DFT.  It is a CPU
check of the host-side planning logic and of the kernel's addressing -- not a product path."""

R_MOD = 0x30644E72E131A029B85045B68181585D2833E84879B9709143E1F593F0000001
ROOT_28 = 0x03DDB9F5166D18B798865EA93DD31F743215CF6DD39329C8D34F1ED960C37C9C


def omega_for(log_n):
    w = ROOT_28
    for _ in range(log_n, 28):
        w = w * w % R_MOD
    return w


def brev(x, bits):
    r = 0
    for _ in range(bits):
        r = (r << 1) | (x & 1)
        x >>= 1
    return r


def swz(i):
    return i ^ ((i >> 3) & 7)


class Plan:
    def __init__(self, log_n, tile_bits, max_bits, pref_inner_bits=None, max_inner_bits=None):
        """max_bits bounds the final pass (NTT_MAX_BITS), pref_inner_bits the others (NTT_PREF_INNER_BITS), relaxed to max_inner_bits
        (NTT_MAX_INNER_BITS) only when three passes would not reach log_n otherwise."""
        pi = max_bits if pref_inner_bits is None else pref_inner_bits
        mi = pi if max_inner_bits is None else max_inner_bits
        self.log_n, self.tile_bits, self.max_bits = log_n, tile_bits, max_bits
        if log_n <= max_bits:
            self.bits = [log_n]
        elif log_n <= pi + max_bits:
            b0 = min(pi, (log_n + 1) // 2)
            self.bits = [b0, log_n - b0]
        else:
            cap = mi if log_n > 2 * pi + max_bits else pi
            b0 = min(cap, (log_n + 2) // 3)
            b1 = min(cap, (log_n - b0 + 1) // 2)
            self.bits = [b0, b1, log_n - b0 - b1]
        assert all(b <= mi for b in self.bits[:-1]) and self.bits[-1] <= max_bits
        self.npass = len(self.bits)

    def geom(self, ps):
        consumed = sum(self.bits[:ps])
        a = self.bits[ps]
        log_inner = self.log_n - consumed - a
        is_final = ps == self.npass - 1
        lc = self.tile_bits - a
        cap = (self.bits[0] if self.npass > 1 else 0) if is_final else log_inner
        lc = min(lc, cap)
        return dict(a=a, shift=consumed, log_inner=log_inner, is_final=is_final, log_c=lc)


def boundary_table(plan, ps, omega, scale=1):
    g = plan.geom(ps)
    a, lc, li, sh = g["a"], g["log_c"], g["log_inner"], g["shift"]
    out = []
    for idx in range(1 << (a + li)):
        c = idx & ((1 << lc) - 1)
        q = (idx >> lc) & ((1 << a) - 1)
        cblk = idx >> (lc + a)
        j_in = (cblk << lc) + c
        k = brev(q, a)
        out.append(pow(omega, (j_in * k) << sh, R_MOD) * scale % R_MOD)
    return out


def run_pass(plan, ps, src, dst, omega, threads, coset_in=None, in_scale=None, scale=None, coset_out=None, tw=None, box_rows=4):
    """One launch of ntt_tile_kernel for one column.  src/dst: python lists (n).  threads: NTT_THREADS of the model."""
    g = plan.geom(ps)
    a, log_c, log_inner, is_final = g["a"], g["log_c"], g["log_inner"], g["is_final"]
    A, C = 1 << a, 1 << log_c
    elems = A << log_c
    n = 1 << plan.log_n
    a1 = plan.bits[0] if (is_final and plan.npass >= 2) else 0
    a2 = plan.bits[1] if (is_final and plan.npass == 3) else 0
    w_loc = pow(omega, 1 << (plan.log_n - a), R_MOD)
    loc = [pow(w_loc, i, R_MOD) for i in range(max(1, A >> 1))]
    tiles_per_col = n >> (a + log_c)
    per_thread = 8
    assert elems <= threads * per_thread
    for tau in range(tiles_per_col):
        lb = (a1 if is_final else log_inner) - log_c
        blk = tau & ((1 << lb) - 1)
        outer = tau >> lb
        c0 = blk << log_c
        # ---- issue_data: L0 buffer, element (c, r) at index c * A + r
        L0 = [None] * elems
        if not is_final:
            br = min(A, box_rows)
            for c in range(C):
                for r0 in range(0, A, br):
                    x_coord, y_coord = (c0 + c) * 4, (outer << a) + r0       # u64 units / rows
                    for rr in range(br):
                        # tensor view: dim0 = S * 4 u64, dim1 = n / S rows -> element index = row * S + x / 4
                        L0[(c << a) + r0 + rr] = src[((y_coord + rr) << log_inner) + x_coord // 4]
        else:
            for c in range(C):
                sub = ((c0 + c) << a2) + outer
                for r in range(A):
                    L0[(c << a) + r] = src[(sub << a) + r]
        L1 = [None] * (C * (A + 1))
        # ---- rounds
        def ntt_round(R, first, last, s):
            E = 1 << R
            NG = per_thread // E
            total_groups = elems >> R
            lgpc = a - R
            lq = 0 if last else a - s - R
            assert lq == a - s - R
            q = 1 << lq
            regs = {}
            for tid in range(threads):
                for u in range(NG):
                    G = (tid + u * threads) & (total_groups - 1)
                    if first or log_c < 3:
                        c, gg = G >> lgpc, G & ((1 << lgpc) - 1)
                        ploc, ghi = gg & (q - 1), gg >> lq
                    else:
                        c = G & ((1 << log_c) - 1)
                        rest, lgh = G >> log_c, lgpc - lq
                        ghi, ploc = rest & ((1 << lgh) - 1), rest >> lgh
                    rbase = (ghi << (lq + R)) + ploc
                    x = []
                    for m in range(E):
                        r = rbase + (m << lq)
                        if first:
                            v = L0[(c << a) + r]
                            idx = (r << log_inner) + c0 + c
                            if coset_in is not None:
                                v = v * coset_in[idx % 3] % R_MOD
                            if in_scale is not None:
                                v = v * in_scale[idx] % R_MOD
                        else:
                            v = L1[c * (A + 1) + swz(r)]
                        x.append(v)
                    regs[(tid, u)] = (c, ploc, rbase, x)
            # barrier; compute + store
            for (tid, u), (c, ploc, rbase, x) in regs.items():
                for t in range(R):
                    d = E >> (t + 1)
                    trivial = (d << lq) == 1
                    for m in range(E):
                        if (m & d) == 0:
                            uu, vv = x[m], x[m + d]
                            x[m] = (uu + vv) % R_MOD
                            dif = (uu - vv) % R_MOD
                            if last:
                                if (m & (d - 1)) != 0:
                                    dif = dif * loc[(m & (d - 1)) << (s + t)] % R_MOD
                            elif not trivial:
                                pos = ploc + ((m & (d - 1)) << lq)
                                dif = dif * loc[pos << (s + t)] % R_MOD
                            x[m + d] = dif
                for m in range(E):
                    r = rbase + (m << lq)
                    L1[c * (A + 1) + swz(r)] = x[m]
        if a == 0:
            ntt_round(0, True, False, 0)
        else:
            r0 = a % 3 if a % 3 else 3
            ntt_round(r0, True, False, 0)
            s = r0
            while a - s > 3:
                ntt_round(3, False, False, s); s += 3
            if a - s == 3:
                ntt_round(3, False, True, s)
        # ---- store phase (non-final: the boundary-table tile arrives in chunks of one store iteration through the ring)
        chunk_elems = min(elems, threads)
        nchunks = elems // chunk_elems
        for u in range(nchunks if not is_final else per_thread):
            for tid in range(threads):
                e = tid + u * threads
                if (tid >= chunk_elems) if not is_final else (e >= elems): continue
                c, q = e & (C - 1), e >> log_c
                k = brev(q, a)
                v = L1[c * (A + 1) + swz(q)]
                if not is_final:
                    v = v * tw[blk * elems + u * chunk_elems + tid] % R_MOD
                    oidx = ((((outer << a) + k)) << log_inner) + c0 + c
                else:
                    oidx = c0 + c + (outer << a1) + (k << (a1 + a2))
                    if scale is not None: v = v * scale % R_MOD
                    if coset_out is not None:
                        mm = oidx % 3
                        if mm: v = v * coset_out[3 - mm] % R_MOD
                dst[oidx] = v


def ntt(values, log_n, omega, tile_bits, max_bits, threads, scale=None, coset_zeta=0, zeta=None, in_scale=None, pref_inner_bits=None, max_inner_bits=None):
    """Mirror of ntt_fr_batch_device for one column."""
    plan = Plan(log_n, tile_bits, max_bits, pref_inner_bits, max_inner_bits)
    n = 1 << log_n
    scratch = [None] * n
    out = [None] * n
    cz_in = [1, zeta, zeta * zeta % R_MOD] if coset_zeta == 1 else None
    cz_out = [1, zeta, zeta * zeta % R_MOD] if coset_zeta == 2 else None
    for ps in range(plan.npass):
        g = plan.geom(ps)
        src = values if ps == 0 else scratch
        dst = out if g["is_final"] else scratch
        tw = None
        if not g["is_final"]:
            tw = boundary_table(plan, ps, omega, scale if (scale is not None and ps == plan.npass - 2) else 1)
        if src is dst:
            src = list(src)   # tiles read and write disjoint sets; the model just snapshots
        run_pass(plan, ps, src, dst, omega, threads, coset_in=cz_in if ps == 0 else None, in_scale=in_scale if ps == 0 else None,
                 scale=scale if plan.npass == 1 else None, coset_out=cz_out if g["is_final"] else None, tw=tw)
    return out


def dft(values, omega):
    n = len(values)
    return [sum(values[j] * pow(omega, j * k, R_MOD) for j in range(n)) % R_MOD for k in range(n)]
