/* oracle/mont.h — width-generic Montgomery field arithmetic on 64-bit limbs (TEST INFRASTRUCTURE).
 *
 * PARITY UNPINNED: the reference tree holds no field code; dusk-bls12_381 0.8
 * (pinned at /root/reference/Cargo.toml:20) is not on disk.  This restates its representation
 * — `Scalar` = 4×u64, R = 2^256; `Fp` = 6×u64, R = 2^384; every value kept fully reduced —
 * from the public definition of Montgomery arithmetic.  All derived constants (−m⁻¹ mod 2^64,
 * R, R²) are computed at start-up from the modulus alone and then checked in tests against
 * SURVEY.md Appendix A.
 *
 * Include with MF_NAME (prefix) and MF_N (limb count) defined; may be included more than once.
 */
#include <stdint.h>
#include <string.h>
#include "mont_asm.h"

#ifndef MF_CAT
#define MF_CAT_(a, b) a##_##b
#define MF_CAT(a, b) MF_CAT_(a, b)
typedef unsigned __int128 u128;
#endif

#define MF(x) MF_CAT(MF_NAME, x)

typedef struct { uint64_t l[MF_N]; } MF(t);

static MF(t) MF(MOD);      /* modulus, set by <name>_init */
static uint64_t MF(INV);   /* −MOD⁻¹ mod 2^64 */
static MF(t) MF(R1);       /* R mod MOD  (Montgomery one) */
static MF(t) MF(R2);       /* R² mod MOD */

static inline int MF(geq)(const MF(t) *a, const MF(t) *b) {
    for (int i = MF_N - 1; i >= 0; i--) {
        if (a->l[i] > b->l[i]) return 1;
        if (a->l[i] < b->l[i]) return 0;
    }
    return 1;
}
static inline int MF(eq)(const MF(t) *a, const MF(t) *b) { return memcmp(a, b, sizeof(*a)) == 0; }
static inline int MF(is_zero)(const MF(t) *a) {
    uint64_t o = 0;
    for (int i = 0; i < MF_N; i++) o |= a->l[i];
    return o == 0;
}
static inline uint64_t MF(raw_add)(MF(t) *o, const MF(t) *a, const MF(t) *b) {
    u128 c = 0;
    for (int i = 0; i < MF_N; i++) { c += (u128)a->l[i] + b->l[i]; o->l[i] = (uint64_t)c; c >>= 64; }
    return (uint64_t)c;
}
static inline uint64_t MF(raw_sub)(MF(t) *o, const MF(t) *a, const MF(t) *b) {
    uint64_t br = 0;
    for (int i = 0; i < MF_N; i++) {
        u128 d = (u128)a->l[i] - b->l[i] - br;
        o->l[i] = (uint64_t)d;
        br = (uint64_t)(d >> 64) & 1;
    }
    return br;
}
/* o = (hi:s) − MOD if that is non-negative, else s; branch-free (as upstream's constant-time code). */
static inline void MF(cond_sub_mod)(MF(t) *o, const MF(t) *s, uint64_t hi) {
    MF(t) d;
    uint64_t br = MF(raw_sub)(&d, s, &MF(MOD));
    uint64_t keep = (uint64_t)0 - (uint64_t)(hi < br); /* all-ones when the subtraction underflowed */
    for (int i = 0; i < MF_N; i++) o->l[i] = (s->l[i] & keep) | (d.l[i] & ~keep);
}
static inline void MF(add)(MF(t) *o, const MF(t) *a, const MF(t) *b) {
    MF(t) s;
    uint64_t c = MF(raw_add)(&s, a, b);
    MF(cond_sub_mod)(o, &s, c);
}
static inline void MF(sub)(MF(t) *o, const MF(t) *a, const MF(t) *b) {
    MF(t) s, m;
    uint64_t mask = (uint64_t)0 - MF(raw_sub)(&s, a, b);
    for (int i = 0; i < MF_N; i++) m.l[i] = MF(MOD).l[i] & mask;
    MF(raw_add)(o, &s, &m);
}
static inline void MF(neg)(MF(t) *o, const MF(t) *a) {
    if (MF(is_zero)(a)) { *o = *a; return; }
    MF(raw_sub)(o, &MF(MOD), a);
}
static inline void MF(dbl)(MF(t) *o, const MF(t) *a) { MF(add)(o, a, a); }

/* Montgomery product o = a·b·R⁻¹ mod MOD, fully reduced: schoolbook 2N-limb product followed by
 * N word-by-word reduction rounds (the product-then-`montgomery_reduce` shape dusk-bls12_381 uses). */
static inline void MF(mul)(MF(t) *o, const MF(t) *a, const MF(t) *b) {
#if defined(ORC_HAVE_MONT_ASM) && !defined(ORC_NO_ASM)
    if (MF_N == 4) { orc_mont_mul4_asm(o->l, a->l, b->l, MF(MOD).l, MF(INV)); return; }
    if (MF_N == 6) { orc_mont_mul6_asm(o->l, a->l, b->l, MF(MOD).l, MF(INV)); return; }
#endif
    uint64_t t[2 * MF_N];
    {
        u128 c = 0;
#pragma GCC unroll 8
        for (int j = 0; j < MF_N; j++) {
            c += (u128)a->l[j] * b->l[0];
            t[j] = (uint64_t)c;
            c >>= 64;
        }
        t[MF_N] = (uint64_t)c;
    }
#pragma GCC unroll 8
    for (int i = 1; i < MF_N; i++) {
        u128 c = 0;
#pragma GCC unroll 8
        for (int j = 0; j < MF_N; j++) {
            c += (u128)a->l[j] * b->l[i] + t[i + j];
            t[i + j] = (uint64_t)c;
            c >>= 64;
        }
        t[i + MF_N] = (uint64_t)c;
    }
    uint64_t carry2 = 0;
#pragma GCC unroll 8
    for (int i = 0; i < MF_N; i++) {
        uint64_t m = t[i] * MF(INV);
        u128 c = (u128)m * MF(MOD).l[0] + t[i];
        c >>= 64;
#pragma GCC unroll 8
        for (int j = 1; j < MF_N; j++) {
            c += (u128)m * MF(MOD).l[j] + t[i + j];
            t[i + j] = (uint64_t)c;
            c >>= 64;
        }
        c += (u128)t[i + MF_N] + carry2;
        t[i + MF_N] = (uint64_t)c;
        carry2 = (uint64_t)(c >> 64);
    }
    MF(t) r;
    memcpy(r.l, t + MF_N, sizeof(r.l));
    MF(cond_sub_mod)(o, &r, carry2);
}
static inline void MF(sqr)(MF(t) *o, const MF(t) *a) { MF(mul)(o, a, a); }

static inline void MF(to_mont)(MF(t) *o, const MF(t) *a) { MF(mul)(o, a, &MF(R2)); }
static inline void MF(from_mont)(MF(t) *o, const MF(t) *a) {
    MF(t) one;
    memset(&one, 0, sizeof(one));
    one.l[0] = 1;
    MF(mul)(o, a, &one);
}
/* a^e for a multi-limb exponent (plain integer, LE limbs); a and result in Montgomery form. */
static inline void MF(pow)(MF(t) *o, const MF(t) *a, const uint64_t *e, int e_limbs) {
    MF(t) r = MF(R1);
    for (int i = e_limbs * 64 - 1; i >= 0; i--) {
        MF(sqr)(&r, &r);
        if ((e[i / 64] >> (i % 64)) & 1) MF(mul)(&r, &r, a);
    }
    *o = r;
}
/* Fermat inversion a^(MOD−2); 0 ↦ 0. */
static inline void MF(inv)(MF(t) *o, const MF(t) *a) {
    MF(t) e = MF(MOD), two;
    memset(&two, 0, sizeof(two));
    two.l[0] = 2;
    MF(raw_sub)(&e, &e, &two);
    MF(pow)(o, a, e.l, MF_N);
}
static void MF(init)(const uint64_t *modulus) {
    memcpy(MF(MOD).l, modulus, sizeof(MF(MOD).l));
    uint64_t m0 = modulus[0], x = 1; /* Newton: x ← x(2 − m0·x) doubles the correct bits */
    for (int i = 0; i < 6; i++) x *= 2 - m0 * x;
    MF(INV) = (uint64_t)0 - x;
    /* R = 2^(64·N) mod MOD and R² by modular doubling from 1. */
    MF(t) v;
    memset(&v, 0, sizeof(v));
    v.l[0] = 1;
    for (int i = 0; i < 2 * 64 * MF_N; i++) {
        uint64_t c = MF(raw_add)(&v, &v, &v);
        if (c || MF(geq)(&v, &MF(MOD))) MF(raw_sub)(&v, &v, &MF(MOD));
        if (i == 64 * MF_N - 1) MF(R1) = v;
    }
    MF(R2) = v;
}
#undef MF
