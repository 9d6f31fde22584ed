/*
 * zk_oracle.c -- CPU oracle for the Halo2/KZG hot path (MSM, NTT, encodings).  TEST INFRASTRUCTURE ONLY.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may build, link or
 * call this library.  The product path (zkevm-circuits_b200/) must never route through it.
 *
 * The hot-path algorithms live in third-party crates absent from /root/reference:
 *   halo2_proofs 1.1.0  = scroll-tech/halo2 branch v1.1 @ e5ddf67e5ae16be38d6368ed355c7c41906272ab (Cargo.lock:2214-2216)
 *   halo2curves  0.1.0  = scroll-tech/halo2curves branch v0.1.0 @ a495a7b11ad13e5cd0cca7ca5d737b398cfaf1b7 (Cargo.lock:2239-2241)
 * Each function below restates the published algorithm of the named upstream function.  The restatement is pinned
 * (tests/test_oracle_golden.py) against the reference's fixture aggregator/data/batch-task.json (encodings, Montgomery
 * form, domain generators, DELTA) and against an independent big-integer implementation (oracle/pyref.py).
 * The reference's call sites for these functions: circuit-benchmarks/src/super_circuit.rs:117-132 (create_proof ->
 * commit_lagrange -> best_multiexp; lagrange_to_coeff -> best_fft).
 *
 * Build: make -C oracle   (gcc -O3 -march=native -fopenmp -shared -fPIC)  -> oracle/libzkoracle.so
 */
#include <stdio.h>
#include <stdlib.h>
#include <math.h>
#ifdef _OPENMP
#include <omp.h>
/* loops shorter than this run on the calling thread: a fork/join over 128 threads costs more than 8 k field operations (the
 * upstream prover's rayon `parallelize` makes the same kind of size-based decision) */
#define ZKO_PAR_MIN 16384
#endif
#include "zko_field.h"
#include "zko_curve.h"

#define FR (&ZKO_FR)
#define API __attribute__((visibility("default")))

/* explicit thread count: launchers (torchrun) export OMP_NUM_THREADS=1 to their workers, which would silently turn the CPU baseline
 * into a single-thread run; the harness sets the count it wants and reports what it got */
API void zko_set_num_threads(int n) {
    if (n > 0) omp_set_num_threads(n);
}
API int zko_num_threads(void) {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}

/* ------------------------------------------------------------------------------------------------
 * element-wise field helpers (for unit tests of the device field arithmetic); which: 0 = Fr, 1 = Fq
 * ---------------------------------------------------------------------------------------------- */
static const zko_field_params *pick(int which) { return which ? &ZKO_FQ : &ZKO_FR; }

API void zko_field_binop(int which, int op, const uint64_t *a, const uint64_t *b, uint64_t *out, uint64_t n) {
    const zko_field_params *F = pick(which);
#pragma omp parallel for schedule(static) if (n >= ZKO_PAR_MIN)
    for (int64_t i = 0; i < (int64_t)n; ++i) {
        const fe_t *x = (const fe_t *)(a + 4 * i), *y = (const fe_t *)(b + 4 * i);
        fe_t *o = (fe_t *)(out + 4 * i);
        switch (op) {
        case 0: fe_add(o, x, y, F); break;
        case 1: fe_sub(o, x, y, F); break;
        case 2: fe_mul(o, x, y, F); break;
        default: break;
        }
    }
}
API void zko_field_unop(int which, int op, const uint64_t *a, uint64_t *out, uint64_t n) {
    const zko_field_params *F = pick(which);
#pragma omp parallel for schedule(static) if (n >= ZKO_PAR_MIN)
    for (int64_t i = 0; i < (int64_t)n; ++i) {
        const fe_t *x = (const fe_t *)(a + 4 * i);
        fe_t *o = (fe_t *)(out + 4 * i);
        switch (op) {
        case 0: fe_inv(o, x, F); break;                       /* invert (0 -> 0) */
        case 1: fe_from_canonical(o, x->l, F); break;         /* canonical -> Montgomery */
        case 2: { uint64_t c[4]; fe_to_canonical(c, x, F); memcpy(o->l, c, 32); } break;
        case 3: fe_sqr(o, x, F); break;
        case 4: fe_neg(o, x, F); break;
        default: break;
        }
    }
}
API void zko_fr_from_u512(const uint8_t *bytes64, uint64_t *out, uint64_t n) {
    for (uint64_t i = 0; i < n; ++i) fe_from_u512((fe_t *)(out + 4 * i), bytes64 + 64 * i, FR);
}
API void zko_fr_pow(const uint64_t a[4], const uint64_t e[4], uint64_t out[4]) { fe_pow((fe_t *)out, (const fe_t *)a, e, FR); }

/* omega_k = ROOT_OF_UNITY^(2^(28-k)); halo2_proofs poly/domain.rs EvaluationDomain::new */
static const uint64_t FR_ROOT_OF_UNITY_CANON[4] = {0xd34f1ed960c37c9cULL, 0x3215cf6dd39329c8ULL, 0x98865ea93dd31f74ULL, 0x03ddb9f5166d18b7ULL};
API void zko_fr_omega(uint32_t k, uint64_t out[4]) {
    fe_t w;
    fe_from_canonical(&w, FR_ROOT_OF_UNITY_CANON, FR);
    for (uint32_t i = k; i < 28; ++i) fe_sqr(&w, &w, FR);
    memcpy(out, w.l, 32);
}

/* ------------------------------------------------------------------------------------------------
 * best_fft  (halo2_proofs src/arithmetic.rs `best_fft`): in-place radix-2, natural order in and out:
 *   1. swap a[k] <-> a[bitreverse(k)]   2. twiddles[i] = omega^i, i < n/2
 *   3. log_n butterfly layers (upstream recurses with rayon::join; the arithmetic per butterfly is
 *      t = b * twiddle; b = a - t; a = a + t, twiddle index stride n / chunk) -- layers are data-parallel,
 *      here an OpenMP loop per layer.  Result: a'[k] = sum_j a[j] omega^(jk).
 * ---------------------------------------------------------------------------------------------- */
static inline uint64_t bitrev(uint64_t k, uint32_t l) {
    uint64_t r = 0;
    for (uint32_t i = 0; i < l; ++i) { r = (r << 1) | (k & 1); k >>= 1; }
    return r;
}

API void zko_best_fft(uint64_t *data, const uint64_t omega[4], uint32_t log_n) {
    fe_t *a = (fe_t *)data;
    const uint64_t n = 1ULL << log_n;
    if (log_n == 0) return;
#pragma omp parallel for schedule(static) if (n >= ZKO_PAR_MIN)
    for (int64_t k = 0; k < (int64_t)n; ++k) {
        uint64_t rk = bitrev((uint64_t)k, log_n);
        if ((uint64_t)k < rk) { fe_t t = a[k]; a[k] = a[rk]; a[rk] = t; }
    }
    const uint64_t half = n / 2;
    fe_t *tw = (fe_t *)malloc(sizeof(fe_t) * (half ? half : 1));
    fe_t w;
    memcpy(w.l, omega, 32);
    /* twiddles: serial prefix in blocks, each block seeded with omega^(block start) */
    {
        int nt = zko_num_threads();
        uint64_t blk = (half + nt - 1) / nt;
        if (blk == 0) blk = 1;
#pragma omp parallel for schedule(static) if (n >= ZKO_PAR_MIN)
        for (int64_t b = 0; b < (int64_t)((half + blk - 1) / blk); ++b) {
            uint64_t s = (uint64_t)b * blk, e = s + blk > half ? half : s + blk;
            uint64_t ex[4] = {s, 0, 0, 0};
            fe_t cur;
            fe_pow(&cur, &w, ex, FR);
            for (uint64_t i = s; i < e; ++i) { tw[i] = cur; fe_mul(&cur, &cur, &w, FR); }
        }
    }
    /* layers: same butterflies as upstream; scheduled cache-blocked -- the first `local` layers only mix elements inside
     * blocks of 2^local entries, so each block runs them back to back while it is cache resident (one thread per block);
     * the remaining layers are data-parallel sweeps.  The arithmetic per butterfly is unchanged. */
    uint32_t local = log_n < 14 ? log_n : 14;
    if (n >> local < (uint64_t)zko_num_threads()) local = 0;   /* too few blocks to keep all threads busy */
    if (local) {
        const uint64_t bsz = 1ULL << local;
#pragma omp parallel for schedule(static) if (n >= ZKO_PAR_MIN)
        for (int64_t b = 0; b < (int64_t)(n >> local); ++b) {
            fe_t *blk = a + (uint64_t)b * bsz;
            uint64_t chunk = 2, twiddle_chunk = half;
            for (uint32_t layer = 0; layer < local; ++layer) {
                const uint64_t hc = chunk / 2;
                for (uint64_t s0 = 0; s0 < bsz; s0 += chunk) {
                    for (uint64_t i = 0; i < hc; ++i) {
                        fe_t *lo = &blk[s0 + i], *hi = lo + hc;
                        fe_t t;
                        if (i == 0) t = *hi; else fe_mul(&t, hi, &tw[i * twiddle_chunk], FR);
                        fe_sub(hi, lo, &t, FR);
                        fe_add(lo, lo, &t, FR);
                    }
                }
                chunk *= 2;
                twiddle_chunk /= 2;
            }
        }
    }
    uint64_t chunk = 2ULL << local, twiddle_chunk = half >> local;
    for (uint32_t layer = local; layer < log_n; ++layer) {
        const uint64_t hc = chunk / 2;
#pragma omp parallel for schedule(static) if (n >= ZKO_PAR_MIN)
        for (int64_t idx = 0; idx < (int64_t)half; ++idx) {
            uint64_t blk = (uint64_t)idx / hc, i = (uint64_t)idx % hc;
            fe_t *lo = &a[blk * chunk + i], *hi = lo + hc;
            fe_t t;
            if (i == 0) t = *hi; else fe_mul(&t, hi, &tw[i * twiddle_chunk], FR);
            fe_sub(hi, lo, &t, FR);
            fe_add(lo, lo, &t, FR);
        }
        chunk *= 2;
        twiddle_chunk /= 2;
    }
    free(tw);
}

/* ------------------------------------------------------------------------------------------------
 * best_multiexp (halo2_proofs src/arithmetic.rs `multiexp_serial` + `best_multiexp`):
 *   c = 1 if n<4, 3 if n<32, else ceil(ln n);  segments = 256/c + 1, processed high to low with c doublings
 *   between; per segment 2^c - 1 buckets (None / Affine / Projective), each point added to bucket[digit-1];
 *   summation by parts from the top bucket down (running_sum, acc += running_sum).
 *   best_multiexp: chunk = n / num_threads; one multiexp_serial per chunk; results folded by addition.
 * out = Jacobian (x,y,z) 12 limbs.
 * ---------------------------------------------------------------------------------------------- */
static inline uint64_t get_at(uint32_t segment, uint32_t c, const uint8_t bytes[32]) {
    uint32_t skip_bits = segment * c, skip_bytes = skip_bits / 8;
    if (skip_bytes >= 32) return 0;
    uint8_t v[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    for (uint32_t i = 0; i < 8 && skip_bytes + i < 32; ++i) v[i] = bytes[skip_bytes + i];
    uint64_t tmp;
    memcpy(&tmp, v, 8);
    tmp >>= skip_bits - skip_bytes * 8;
    return tmp % (1ULL << c);
}

typedef struct { int kind; g1a_t a; g1j_t j; } bucket_t; /* 0 none, 1 affine, 2 projective */

static void multiexp_serial(const uint64_t *scalars_mont, const g1a_t *bases, uint64_t n, g1j_t *acc) {
    uint8_t *reprs = (uint8_t *)malloc(32 * (n ? n : 1));
    for (uint64_t i = 0; i < n; ++i) {
        uint64_t c4[4];
        fe_to_canonical(c4, (const fe_t *)(scalars_mont + 4 * i), FR);
        memcpy(reprs + 32 * i, c4, 32);
    }
    uint32_t c;
    if (n < 4) c = 1; else if (n < 32) c = 3; else c = (uint32_t)ceil(log((double)(uint32_t)n));
    uint32_t segments = 256 / c + 1;
    uint64_t nb = (1ULL << c) - 1;
    bucket_t *buckets = (bucket_t *)malloc(sizeof(bucket_t) * nb);
    for (int32_t seg = (int32_t)segments - 1; seg >= 0; --seg) {
        for (uint32_t d = 0; d < c; ++d) g1j_double(acc, acc);
        for (uint64_t b = 0; b < nb; ++b) buckets[b].kind = 0;
        for (uint64_t i = 0; i < n; ++i) {
            uint64_t dg = get_at((uint32_t)seg, c, reprs + 32 * i);
            if (!dg) continue;
            bucket_t *bk = &buckets[dg - 1];
            if (bk->kind == 0) { bk->kind = 1; bk->a = bases[i]; }
            else if (bk->kind == 1) { g1j_t t; g1j_from_affine(&t, &bk->a); g1j_add_affine(&bk->j, &t, &bases[i]); bk->kind = 2; }
            else g1j_add_affine(&bk->j, &bk->j, &bases[i]);
        }
        g1j_t running;
        g1j_set_identity(&running);
        for (int64_t b = (int64_t)nb - 1; b >= 0; --b) {
            if (buckets[b].kind == 1) g1j_add_affine(&running, &running, &buckets[b].a);
            else if (buckets[b].kind == 2) g1j_add(&running, &running, &buckets[b].j);
            g1j_add(acc, acc, &running);
        }
    }
    free(buckets);
    free(reprs);
}

API void zko_best_multiexp(const uint64_t *scalars_mont, const uint64_t *bases_affine, uint64_t n, uint64_t out_jac[12], int threads) {
    const g1a_t *bases = (const g1a_t *)bases_affine;
    if (threads <= 0) threads = zko_num_threads();
    g1j_t total;
    g1j_set_identity(&total);
    if (n > (uint64_t)threads) {
        uint64_t chunk = n / threads;
        uint64_t nchunks = (n + chunk - 1) / chunk;
        g1j_t *res = (g1j_t *)malloc(sizeof(g1j_t) * nchunks);
#pragma omp parallel for schedule(dynamic, 1) num_threads(threads)
        for (int64_t ci = 0; ci < (int64_t)nchunks; ++ci) {
            uint64_t s = (uint64_t)ci * chunk, len = s + chunk > n ? n - s : chunk;
            g1j_set_identity(&res[ci]);
            multiexp_serial(scalars_mont + 4 * s, bases + s, len, &res[ci]);
        }
        for (uint64_t ci = 0; ci < nchunks; ++ci) g1j_add(&total, &total, &res[ci]);
        free(res);
    } else {
        multiexp_serial(scalars_mont, bases, n, &total);
    }
    memcpy(out_jac, &total, sizeof(g1j_t));
}

/* ------------------------------------------------------------------------------------------------ G1 helpers */
API void zko_g1_to_affine(const uint64_t jac[12], uint64_t aff[8]) { g1j_to_affine((g1a_t *)aff, (const g1j_t *)jac); }
API void zko_g1_compress(const uint64_t aff[8], uint8_t out[32]) { g1a_compress(out, (const g1a_t *)aff); }
API int zko_g1_is_on_curve(const uint64_t aff[8]) { return g1a_is_on_curve((const g1a_t *)aff); }
API void zko_g1_add(const uint64_t a[12], const uint64_t b[12], uint64_t out[12]) { g1j_add((g1j_t *)out, (const g1j_t *)a, (const g1j_t *)b); }

/* out[i] = [scalars[i]] * base  (affine out); scalars Montgomery Fr.  Used to build SRS-like base sets:
 * ParamsKZG::unsafe_setup_with_s builds g[i] = [s^i] G1 (halo2_proofs poly/kzg/commitment.rs). */
API void zko_g1_fixed_base_mul(const uint64_t base_aff[8], const uint64_t *scalars_mont, uint64_t n, uint64_t *out_aff) {
#pragma omp parallel for schedule(static)
    for (int64_t i = 0; i < (int64_t)n; ++i) {
        uint64_t c4[4];
        fe_to_canonical(c4, (const fe_t *)(scalars_mont + 4 * i), FR);
        g1j_t r;
        g1j_mul_canonical(&r, (const g1a_t *)base_aff, c4);
        g1j_to_affine((g1a_t *)(out_aff + 8 * i), &r);
    }
}

/* powers: out[i] = s^i (Montgomery) */
API void zko_fr_powers(const uint64_t s[4], uint64_t n, uint64_t *out) {
    fe_t cur;
    fe_one(&cur, FR);
    for (uint64_t i = 0; i < n; ++i) { memcpy(out + 4 * i, cur.l, 32); fe_mul(&cur, &cur, (const fe_t *)s, FR); }
}

API void zko_g1_generator(uint64_t aff[8]) {
    g1a_t g;
    fe_from_u64(&g.x, 1, FQ);
    fe_from_u64(&g.y, 2, FQ);
    memcpy(aff, &g, 64);
}
