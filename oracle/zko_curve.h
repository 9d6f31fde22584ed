/*
 * zko_curve.h -- CPU oracle: BN254 G1 (y^2 = x^3 + 3 over Fq).   TEST INFRASTRUCTURE ONLY (see zko_field.h).
 *
 * Restates halo2curves 0.1.0 @ a495a7b src/bn256/curve.rs + src/derive/curve.rs (`new_curve_impl!`):
 *   G1Affine { x, y } with identity encoded as (0, 0); G1 { x, y, z } Jacobian with identity z = 0;
 *   `to_affine` (x/z^2, y/z^3), compressed encoding = 32 B little-endian canonical x with
 *   (y & 1) << 6 OR-ed into byte 31, identity = 32 zero bytes.
 * The group law is restated from the standard Jacobian formulas (dbl-2009-l, add-2007-bl, madd-2007-bl);
 * affine results are unique so any complete formula set yields identical bytes.
 * Pinned: tests/test_oracle_golden.py compresses the fixture's `preprocessed` points and compares with the vk bytes.
 */
#ifndef ZKO_CURVE_H
#define ZKO_CURVE_H
#include "zko_field.h"

typedef struct { fe_t x, y; } g1a_t;    /* affine, Montgomery coords; identity = (0,0) */
typedef struct { fe_t x, y, z; } g1j_t; /* Jacobian; identity: z = 0 */

#define FQ (&ZKO_FQ)

static inline int g1a_is_identity(const g1a_t *p) { return fe_is_zero(&p->x) && fe_is_zero(&p->y); }
static inline int g1j_is_identity(const g1j_t *p) { return fe_is_zero(&p->z); }
static inline void g1j_set_identity(g1j_t *p) { fe_zero(&p->x); fe_one(&p->y, FQ); fe_zero(&p->z); }
static inline void g1j_from_affine(g1j_t *o, const g1a_t *p) {
    if (g1a_is_identity(p)) { g1j_set_identity(o); return; }
    o->x = p->x; o->y = p->y; fe_one(&o->z, FQ);
}

static inline void g1j_double(g1j_t *o, const g1j_t *p) {
    if (g1j_is_identity(p)) { *o = *p; return; }
    fe_t a, b, c, d, e, f, t, x3, y3, z3;
    fe_sqr(&a, &p->x, FQ);            /* A = X^2 */
    fe_sqr(&b, &p->y, FQ);            /* B = Y^2 */
    fe_sqr(&c, &b, FQ);               /* C = B^2 */
    fe_add(&t, &p->x, &b, FQ);
    fe_sqr(&t, &t, FQ);
    fe_sub(&t, &t, &a, FQ);
    fe_sub(&t, &t, &c, FQ);
    fe_dbl(&d, &t, FQ);               /* D = 2((X+B)^2 - A - C) */
    fe_dbl(&e, &a, FQ);
    fe_add(&e, &e, &a, FQ);           /* E = 3A */
    fe_sqr(&f, &e, FQ);               /* F = E^2 */
    fe_mul(&z3, &p->y, &p->z, FQ);
    fe_dbl(&z3, &z3, FQ);             /* Z3 = 2YZ */
    fe_dbl(&t, &d, FQ);
    fe_sub(&x3, &f, &t, FQ);          /* X3 = F - 2D */
    fe_sub(&t, &d, &x3, FQ);
    fe_mul(&y3, &e, &t, FQ);
    fe_dbl(&c, &c, FQ); fe_dbl(&c, &c, FQ); fe_dbl(&c, &c, FQ);
    fe_sub(&y3, &y3, &c, FQ);         /* Y3 = E(D - X3) - 8C */
    o->x = x3; o->y = y3; o->z = z3;
}

static inline void g1j_add(g1j_t *o, const g1j_t *p, const g1j_t *q) {
    if (g1j_is_identity(p)) { *o = *q; return; }
    if (g1j_is_identity(q)) { *o = *p; return; }
    fe_t z1z1, z2z2, u1, u2, s1, s2, h, i, j, r, v, t, x3, y3, z3;
    fe_sqr(&z1z1, &p->z, FQ);
    fe_sqr(&z2z2, &q->z, FQ);
    fe_mul(&u1, &p->x, &z2z2, FQ);
    fe_mul(&u2, &q->x, &z1z1, FQ);
    fe_mul(&s1, &p->y, &q->z, FQ); fe_mul(&s1, &s1, &z2z2, FQ);
    fe_mul(&s2, &q->y, &p->z, FQ); fe_mul(&s2, &s2, &z1z1, FQ);
    if (fe_eq(&u1, &u2)) {
        if (fe_eq(&s1, &s2)) { g1j_double(o, p); return; }
        g1j_set_identity(o); return;
    }
    fe_sub(&h, &u2, &u1, FQ);
    fe_dbl(&i, &h, FQ); fe_sqr(&i, &i, FQ);     /* I = (2H)^2 */
    fe_mul(&j, &h, &i, FQ);                      /* J = H I */
    fe_sub(&r, &s2, &s1, FQ); fe_dbl(&r, &r, FQ);/* r = 2(S2-S1) */
    fe_mul(&v, &u1, &i, FQ);                     /* V = U1 I */
    fe_sqr(&x3, &r, FQ); fe_sub(&x3, &x3, &j, FQ); fe_dbl(&t, &v, FQ); fe_sub(&x3, &x3, &t, FQ);
    fe_sub(&t, &v, &x3, FQ); fe_mul(&y3, &r, &t, FQ);
    fe_mul(&t, &s1, &j, FQ); fe_dbl(&t, &t, FQ); fe_sub(&y3, &y3, &t, FQ);
    fe_add(&z3, &p->z, &q->z, FQ); fe_sqr(&z3, &z3, FQ); fe_sub(&z3, &z3, &z1z1, FQ); fe_sub(&z3, &z3, &z2z2, FQ);
    fe_mul(&z3, &z3, &h, FQ);
    o->x = x3; o->y = y3; o->z = z3;
}

static inline void g1j_add_affine(g1j_t *o, const g1j_t *p, const g1a_t *q) {
    if (g1a_is_identity(q)) { *o = *p; return; }
    if (g1j_is_identity(p)) { g1j_from_affine(o, q); return; }
    fe_t z1z1, u2, s2, h, hh, i, j, r, v, t, x3, y3, z3;
    fe_sqr(&z1z1, &p->z, FQ);
    fe_mul(&u2, &q->x, &z1z1, FQ);
    fe_mul(&s2, &q->y, &p->z, FQ); fe_mul(&s2, &s2, &z1z1, FQ);
    if (fe_eq(&p->x, &u2)) {
        if (fe_eq(&p->y, &s2)) { g1j_double(o, p); return; }
        g1j_set_identity(o); return;
    }
    fe_sub(&h, &u2, &p->x, FQ);
    fe_sqr(&hh, &h, FQ);
    fe_dbl(&i, &hh, FQ); fe_dbl(&i, &i, FQ);     /* I = 4HH */
    fe_mul(&j, &h, &i, FQ);
    fe_sub(&r, &s2, &p->y, FQ); fe_dbl(&r, &r, FQ);
    fe_mul(&v, &p->x, &i, FQ);
    fe_sqr(&x3, &r, FQ); fe_sub(&x3, &x3, &j, FQ); fe_dbl(&t, &v, FQ); fe_sub(&x3, &x3, &t, FQ);
    fe_sub(&t, &v, &x3, FQ); fe_mul(&y3, &r, &t, FQ);
    fe_mul(&t, &p->y, &j, FQ); fe_dbl(&t, &t, FQ); fe_sub(&y3, &y3, &t, FQ);
    fe_add(&z3, &p->z, &h, FQ); fe_sqr(&z3, &z3, FQ); fe_sub(&z3, &z3, &z1z1, FQ); fe_sub(&z3, &z3, &hh, FQ);
    o->x = x3; o->y = y3; o->z = z3;
}

static inline void g1j_to_affine(g1a_t *o, const g1j_t *p) {
    if (g1j_is_identity(p)) { fe_zero(&o->x); fe_zero(&o->y); return; }
    fe_t zi, zi2, zi3;
    fe_inv(&zi, &p->z, FQ);
    fe_sqr(&zi2, &zi, FQ);
    fe_mul(&zi3, &zi2, &zi, FQ);
    fe_mul(&o->x, &p->x, &zi2, FQ);
    fe_mul(&o->y, &p->y, &zi3, FQ);
}

static inline void g1a_neg(g1a_t *o, const g1a_t *p) {
    o->x = p->x;
    if (g1a_is_identity(p)) { o->y = p->y; return; }
    fe_neg(&o->y, &p->y, FQ);
}

/* o = [s] p, s canonical (non-Montgomery) 4 LE limbs */
static inline void g1j_mul_canonical(g1j_t *o, const g1a_t *p, const uint64_t s[4]) {
    g1j_t acc;
    g1j_set_identity(&acc);
    for (int i = 255; i >= 0; --i) {
        g1j_double(&acc, &acc);
        if ((s[i >> 6] >> (i & 63)) & 1) g1j_add_affine(&acc, &acc, p);
    }
    *o = acc;
}

static inline void g1a_compress(uint8_t out[32], const g1a_t *p) {
    if (g1a_is_identity(p)) { memset(out, 0, 32); return; }
    uint64_t x[4], y[4];
    fe_to_canonical(x, &p->x, FQ);
    fe_to_canonical(y, &p->y, FQ);
    memcpy(out, x, 32); /* little-endian host */
    out[31] |= (uint8_t)((y[0] & 1) << 6);
}

static inline int g1a_is_on_curve(const g1a_t *p) {
    if (g1a_is_identity(p)) return 1;
    fe_t y2, x3, b;
    fe_sqr(&y2, &p->y, FQ);
    fe_sqr(&x3, &p->x, FQ); fe_mul(&x3, &x3, &p->x, FQ);
    fe_from_u64(&b, 3, FQ);
    fe_add(&x3, &x3, &b, FQ);
    return fe_eq(&y2, &x3);
}

#endif
