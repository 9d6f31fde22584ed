/*
 * zko_field.h -- CPU oracle: BN254 Fr / Fq arithmetic.   TEST INFRASTRUCTURE ONLY.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
 * build, link or call anything under oracle/.  The product (zkevm-circuits_b200/) never does.
 *
 * Restates (from the published algorithm; source not under /root/reference):
 *   halo2curves 0.1.0 @ a495a7b  src/bn256/fr.rs, src/bn256/fq.rs  (field_arithmetic!/ field_common! macros:
 *   4 x u64 little-endian limbs, Montgomery form with R = 2^256, `montgomery_reduce`, `mul`, `add`, `sub`,
 *   `invert` = a^(p-2), `from_u512`, `to_repr`/`from_repr` canonical little-endian bytes).
 * Pinned against the reference fixture aggregator/data/batch-task.json (tests/golden/thin_chunk_proof.json):
 *   Montgomery(1), DELTA, DELTA^2, domain generator k=25, n_inv -- tests/test_oracle_golden.py.
 */
#ifndef ZKO_FIELD_H
#define ZKO_FIELD_H
#include <stdint.h>
#include <string.h>

typedef unsigned __int128 u128;
typedef struct { uint64_t l[4]; } fe_t; /* one field element, Montgomery form unless stated */

typedef struct {
    uint64_t p[4];   /* modulus */
    uint64_t inv;    /* -p^{-1} mod 2^64 */
    uint64_t r[4];   /* R   mod p  (Montgomery 1) */
    uint64_t r2[4];  /* R^2 mod p */
} zko_field_params;

static const zko_field_params ZKO_FR = {
    {0x43e1f593f0000001ULL, 0x2833e84879b97091ULL, 0xb85045b68181585dULL, 0x30644e72e131a029ULL},
    0xc2e1f593efffffffULL,
    {0xac96341c4ffffffbULL, 0x36fc76959f60cd29ULL, 0x666ea36f7879462eULL, 0x0e0a77c19a07df2fULL},
    {0x1bb8e645ae216da7ULL, 0x53fe3ab1e35c59e3ULL, 0x8c49833d53bb8085ULL, 0x0216d0b17f4e44a5ULL}};

static const zko_field_params ZKO_FQ = {
    {0x3c208c16d87cfd47ULL, 0x97816a916871ca8dULL, 0xb85045b68181585dULL, 0x30644e72e131a029ULL},
    0x87d20782e4866389ULL,
    {0xd35d438dc58f0d9dULL, 0x0a78eb28f5c70b3dULL, 0x666ea36f7879462cULL, 0x0e0a77c19a07df2fULL},
    {0xf32cfc5b538afa89ULL, 0xb5e71911d44501fbULL, 0x47ab1eff0a417ff6ULL, 0x06d89f71cab8351fULL}};

static inline int fe_is_zero(const fe_t *a) { return (a->l[0] | a->l[1] | a->l[2] | a->l[3]) == 0; }
static inline int fe_eq(const fe_t *a, const fe_t *b) { return memcmp(a, b, sizeof(fe_t)) == 0; }

/* a >= p ? */
static inline int fe_geq_p(const uint64_t a[4], const uint64_t p[4]) {
    for (int i = 3; i >= 0; --i) {
        if (a[i] > p[i]) return 1;
        if (a[i] < p[i]) return 0;
    }
    return 1;
}

static inline void fe_add(fe_t *o, const fe_t *a, const fe_t *b, const zko_field_params *F) {
    u128 c = 0;
    uint64_t t[4];
    for (int i = 0; i < 4; ++i) { c += (u128)a->l[i] + b->l[i]; t[i] = (uint64_t)c; c >>= 64; }
    /* p < 2^254 so no carry out of 256 bits */
    if (fe_geq_p(t, F->p)) {
        u128 br = 0;
        for (int i = 0; i < 4; ++i) { u128 d = (u128)t[i] - F->p[i] - (uint64_t)br; t[i] = (uint64_t)d; br = (d >> 64) & 1; }
    }
    memcpy(o->l, t, 32);
}

static inline void fe_sub(fe_t *o, const fe_t *a, const fe_t *b, const zko_field_params *F) {
    uint64_t t[4];
    u128 br = 0;
    for (int i = 0; i < 4; ++i) { u128 d = (u128)a->l[i] - b->l[i] - (uint64_t)br; t[i] = (uint64_t)d; br = (d >> 64) & 1; }
    if (br) {
        u128 c = 0;
        for (int i = 0; i < 4; ++i) { c += (u128)t[i] + F->p[i]; t[i] = (uint64_t)c; c >>= 64; }
    }
    memcpy(o->l, t, 32);
}

static inline void fe_neg(fe_t *o, const fe_t *a, const zko_field_params *F) {
    fe_t z = {{0, 0, 0, 0}};
    fe_sub(o, &z, a, F);
}

static inline void fe_dbl(fe_t *o, const fe_t *a, const zko_field_params *F) { fe_add(o, a, a, F); }

/* Montgomery multiplication (CIOS), o = a*b*R^{-1} mod p */
static inline void fe_mul(fe_t *o, const fe_t *a, const fe_t *b, const zko_field_params *F) {
    uint64_t t[6] = {0, 0, 0, 0, 0, 0};
    for (int i = 0; i < 4; ++i) {
        u128 c = 0;
        for (int j = 0; j < 4; ++j) {
            c += (u128)a->l[j] * b->l[i] + t[j];
            t[j] = (uint64_t)c;
            c >>= 64;
        }
        c += t[4];
        t[4] = (uint64_t)c;
        t[5] = (uint64_t)(c >> 64);
        uint64_t m = t[0] * F->inv;
        c = (u128)m * F->p[0] + t[0];
        c >>= 64;
        for (int j = 1; j < 4; ++j) {
            c += (u128)m * F->p[j] + t[j];
            t[j - 1] = (uint64_t)c;
            c >>= 64;
        }
        c += t[4];
        t[3] = (uint64_t)c;
        t[4] = t[5] + (uint64_t)(c >> 64);
    }
    if (t[4] || fe_geq_p(t, F->p)) {
        u128 br = 0;
        for (int i = 0; i < 4; ++i) { u128 d = (u128)t[i] - F->p[i] - (uint64_t)br; t[i] = (uint64_t)d; br = (d >> 64) & 1; }
    }
    memcpy(o->l, t, 32);
}

static inline void fe_sqr(fe_t *o, const fe_t *a, const zko_field_params *F) { fe_mul(o, a, a, F); }

static inline void fe_one(fe_t *o, const zko_field_params *F) { memcpy(o->l, F->r, 32); }
static inline void fe_zero(fe_t *o) { memset(o, 0, 32); }

/* canonical integer (4 LE limbs, < p) -> Montgomery */
static inline void fe_from_canonical(fe_t *o, const uint64_t c[4], const zko_field_params *F) {
    fe_t a, r2;
    memcpy(a.l, c, 32);
    memcpy(r2.l, F->r2, 32);
    fe_mul(o, &a, &r2, F);
}
/* Montgomery -> canonical integer */
static inline void fe_to_canonical(uint64_t c[4], const fe_t *a, const zko_field_params *F) {
    fe_t one = {{1, 0, 0, 0}}, t;
    fe_mul(&t, a, &one, F);
    memcpy(c, t.l, 32);
}
static inline void fe_from_u64(fe_t *o, uint64_t v, const zko_field_params *F) {
    uint64_t c[4] = {v, 0, 0, 0};
    fe_from_canonical(o, c, F);
}

/* o = a^e, e as 4 LE limbs */
static inline void fe_pow(fe_t *o, const fe_t *a, const uint64_t e[4], const zko_field_params *F) {
    fe_t acc, base = *a;
    fe_one(&acc, F);
    for (int i = 0; i < 256; ++i) {
        if ((e[i >> 6] >> (i & 63)) & 1) fe_mul(&acc, &acc, &base, F);
        fe_sqr(&base, &base, F);
    }
    *o = acc;
}

/* Fermat inversion a^(p-2); inverse of zero is zero (halo2's batch-invert convention treats 0 separately) */
static inline void fe_inv(fe_t *o, const fe_t *a, const zko_field_params *F) {
    uint64_t e[4];
    memcpy(e, F->p, 32);
    e[0] -= 2; /* p[0] >= 2, no borrow */
    fe_pow(o, a, e, F);
}

/* 64 little-endian bytes -> element (halo2curves `from_u512`: d0*R^2 + d1*R^3 in Montgomery terms == value mod p) */
static inline void fe_from_u512(fe_t *o, const uint8_t b[64], const zko_field_params *F) {
    /* lo + hi * 2^256 mod p, computed as mont(lo)*1 + mont(hi)*mont(2^256) */
    uint64_t lo[4], hi[4];
    memcpy(lo, b, 32);
    memcpy(hi, b + 32, 32);
    /* lo, hi may be >= p: fe_mul tolerates inputs < 2^256 when the other operand is < p (result < 2p handled) --
       to be safe reduce by repeated subtraction (at most 5 times since 2^256 / p < 6) */
    for (int k = 0; k < 6; ++k) {
        if (fe_geq_p(lo, F->p)) { u128 br = 0; for (int i = 0; i < 4; ++i) { u128 d = (u128)lo[i] - F->p[i] - (uint64_t)br; lo[i] = (uint64_t)d; br = (d >> 64) & 1; } }
        if (fe_geq_p(hi, F->p)) { u128 br = 0; for (int i = 0; i < 4; ++i) { u128 d = (u128)hi[i] - F->p[i] - (uint64_t)br; hi[i] = (uint64_t)d; br = (d >> 64) & 1; } }
    }
    fe_t a, bb, r2, t;
    fe_from_canonical(&a, lo, F);
    fe_from_canonical(&bb, hi, F);
    memcpy(r2.l, F->r2, 32);     /* r2 as a Montgomery element represents R mod p = 2^256 mod p */
    fe_mul(&t, &bb, &r2, F);     /* hi * 2^256 */
    fe_add(o, &a, &t, F);
}

#endif
