/* oracle/oracle.c — CPU restatement of the MSM / NTT hot path (TEST INFRASTRUCTURE).
 *
 * This file is the checker the CUDA path is compared against and the CPU baseline bench.py
 * times.  It is never linked into, imported by, or called from the product library.
 *
 * PARITY UNPINNED.  /root/reference contains neither MSM nor FFT code, no tests and no golden
 * vectors; the arithmetic lives in crates that are not on disk:
 *   - dusk-bls12_381 "0.8"  (/root/reference/Cargo.toml:20): Scalar, Fp, G1Affine/G1Projective,
 *     multiscalar_mul::msm_variable_base
 *   - dusk-plonk "0.8.2"    (/root/reference/Cargo.toml:19): fft::EvaluationDomain
 * What is restated here is their *published algorithm* as recollected in SURVEY.md Appendix B
 * (B.1 msm_variable_base, B.2 EvaluationDomain / serial_fft / parallel_fft) so the timed CPU
 * baseline does the same work the Rust crates do.  Correctness is anchored independently on
 * oracle/model.py (plain big ints, no Montgomery) and on algebraic invariants (tests/).
 * The only reference call sites that reach this arithmetic are scalar-field ops
 * (/root/reference/src/zk/gadgets.rs:65-66,213,219,230,241-244).
 *
 * Build: see oracle/Makefile (gcc -O3 -march=native -shared).
 */
#define _GNU_SOURCE
#include <pthread.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define MF_NAME fr
#define MF_N 4
#include "mont.h"
#undef MF_NAME
#undef MF_N
#define MF_NAME fp
#define MF_N 6
#include "mont.h"
#undef MF_NAME
#undef MF_N

#define ORC_API __attribute__((visibility("default")))

/* ------------------------------------------------------------------------------------ constants */
static const uint64_t FR_MODULUS[4] = {0xffffffff00000001ull, 0x53bda402fffe5bfeull, 0x3339d80809a1d805ull,
                                       0x73eda753299d7d48ull};
static const uint64_t FP_MODULUS[6] = {0xb9feffffffffaaabull, 0x1eabfffeb153ffffull, 0x6730d2a0f6b0f624ull,
                                       0x64774b84f38512bfull, 0x4b1ba7b6434bacd7ull, 0x1a0111ea397fe69aull};
/* G1 generator, canonical (non-Montgomery) LE limbs — public BLS12-381 constant. */
static const uint64_t G1_GX[6] = {0xfb3af00adb22c6bbull, 0x6c55e83ff97a1aefull, 0xa14e3a3f171bac58ull,
                                  0xc3688c4f9774b905ull, 0x2695638c4fa9ac0full, 0x17f1d3a73197d794ull};
static const uint64_t G1_GY[6] = {0x0caa232946c5e7e1ull, 0xd03cc744a2888ae4ull, 0x00db18cb2c04b3edull,
                                  0xfcf5e095d5d00af6ull, 0xa09e30ed741d8ae4ull, 0x08b3f481e3aaa0f1ull};
#define TWO_ADICITY 32

static fr_t FR_ROOT_OF_UNITY; /* 7^((r−1)/2^32), Montgomery */
static fr_t FR_GENERATOR;     /* 7, Montgomery */
static fp_t FP_B3;            /* 3·b = 12, Montgomery */
static int g_ready = 0;

static void fr_from_u64(fr_t *o, uint64_t v) {
    fr_t t;
    memset(&t, 0, sizeof(t));
    t.l[0] = v;
    fr_to_mont(o, &t);
}
static void fp_from_u64(fp_t *o, uint64_t v) {
    fp_t t;
    memset(&t, 0, sizeof(t));
    t.l[0] = v;
    fp_to_mont(o, &t);
}

ORC_API void orc_init(void) {
    if (g_ready) return;
    fr_init(FR_MODULUS);
    fp_init(FP_MODULUS);
    fr_from_u64(&FR_GENERATOR, 7);
    /* exponent (r−1) >> 32 */
    uint64_t e[4];
    fr_t rm1 = fr_MOD;
    rm1.l[0] -= 1;
    for (int i = 0; i < 4; i++) e[i] = (rm1.l[i] >> 32) | (i < 3 ? rm1.l[i + 1] << 32 : 0);
    fr_pow(&FR_ROOT_OF_UNITY, &FR_GENERATOR, e, 4);
    fp_from_u64(&FP_B3, 12);
    g_ready = 1;
}

/* ------------------------------------------------------------------------------------ PRNG */
static inline uint64_t splitmix64(uint64_t *x) {
    uint64_t z = (*x += 0x9E3779B97F4A7C15ull);
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}
/* n values uniform in [0,r) as raw LE limbs (same stream as model.random_fr). */
ORC_API void orc_random_fr(uint64_t seed, size_t n, uint64_t *out) {
    orc_init();
    uint64_t st = seed;
    size_t k = 0;
    while (k < n) {
        fr_t v;
        for (int i = 0; i < 4; i++) v.l[i] = splitmix64(&st);
        v.l[3] &= 0x7fffffffffffffffull;
        if (!fr_geq(&v, &fr_MOD)) memcpy(out + 4 * k++, v.l, 32);
    }
}

/* ------------------------------------------------------------------------------------ field API */
ORC_API void orc_fr_consts(uint64_t *modulus, uint64_t *inv, uint64_t *r1, uint64_t *r2, uint64_t *root, uint64_t *gen) {
    orc_init();
    memcpy(modulus, fr_MOD.l, 32);
    *inv = fr_INV;
    memcpy(r1, fr_R1.l, 32);
    memcpy(r2, fr_R2.l, 32);
    memcpy(root, FR_ROOT_OF_UNITY.l, 32);
    memcpy(gen, FR_GENERATOR.l, 32);
}
ORC_API void orc_fp_consts(uint64_t *modulus, uint64_t *inv, uint64_t *r1, uint64_t *r2) {
    orc_init();
    memcpy(modulus, fp_MOD.l, 48);
    *inv = fp_INV;
    memcpy(r1, fp_R1.l, 48);
    memcpy(r2, fp_R2.l, 48);
}
#define VEC_OP(name, T, W, expr)                                                            \
    ORC_API void name(const uint64_t *a, const uint64_t *b, uint64_t *o, size_t n) {        \
        orc_init();                                                                         \
        for (size_t i = 0; i < n; i++) {                                                    \
            T x, y, z;                                                                      \
            memcpy(&x, a + W * i, 8 * W);                                                   \
            memcpy(&y, b + W * i, 8 * W);                                                   \
            expr;                                                                           \
            memcpy(o + W * i, &z, 8 * W);                                                   \
        }                                                                                   \
    }
VEC_OP(orc_fr_mul, fr_t, 4, fr_mul(&z, &x, &y))
VEC_OP(orc_fr_add, fr_t, 4, fr_add(&z, &x, &y))
VEC_OP(orc_fr_sub, fr_t, 4, fr_sub(&z, &x, &y))
VEC_OP(orc_fp_mul, fp_t, 6, fp_mul(&z, &x, &y))
VEC_OP(orc_fp_add, fp_t, 6, fp_add(&z, &x, &y))
VEC_OP(orc_fp_sub, fp_t, 6, fp_sub(&z, &x, &y))
#define VEC_OP1(name, T, W, expr)                                                           \
    ORC_API void name(const uint64_t *a, uint64_t *o, size_t n) {                           \
        orc_init();                                                                         \
        for (size_t i = 0; i < n; i++) {                                                    \
            T x, z;                                                                         \
            memcpy(&x, a + W * i, 8 * W);                                                   \
            expr;                                                                           \
            memcpy(o + W * i, &z, 8 * W);                                                   \
        }                                                                                   \
    }
VEC_OP1(orc_fr_to_mont, fr_t, 4, fr_to_mont(&z, &x))
VEC_OP1(orc_fr_from_mont, fr_t, 4, fr_from_mont(&z, &x))
VEC_OP1(orc_fr_inv, fr_t, 4, fr_inv(&z, &x))
VEC_OP1(orc_fp_to_mont, fp_t, 6, fp_to_mont(&z, &x))
VEC_OP1(orc_fp_from_mont, fp_t, 6, fp_from_mont(&z, &x))
VEC_OP1(orc_fp_inv, fp_t, 6, fp_inv(&z, &x))

/* ------------------------------------------------------------------------------------ G1 */
typedef struct { fp_t x, y; int inf; } g1a_t;  /* affine, Montgomery coordinates */
typedef struct { fp_t x, y, z; } g1p_t;        /* homogeneous projective, identity = (0,1,0) */

static void g1p_identity(g1p_t *o) {
    memset(o, 0, sizeof(*o));
    o->y = fp_R1;
}
static int g1p_is_identity(const g1p_t *a) { return fp_is_zero(&a->z); }
static void g1p_from_affine(g1p_t *o, const g1a_t *a) {
    if (a->inf) { g1p_identity(o); return; }
    o->x = a->x;
    o->y = a->y;
    o->z = fp_R1;
}
static void g1a_from_proj(g1a_t *o, const g1p_t *a) {
    if (g1p_is_identity(a)) { memset(o, 0, sizeof(*o)); o->inf = 1; return; }
    fp_t zi;
    fp_inv(&zi, &a->z);
    fp_mul(&o->x, &a->x, &zi);
    fp_mul(&o->y, &a->y, &zi);
    o->inf = 0;
}
/* Complete addition for short-Weierstrass a = 0 (Renes–Costello–Batina 2016, Alg. 7) — the
 * formula family dusk-bls12_381's G1Projective::add uses. */
static void g1p_add(g1p_t *o, const g1p_t *p, const g1p_t *q) {
    fp_t t0, t1, t2, t3, t4, x3, y3, z3;
    fp_mul(&t0, &p->x, &q->x);
    fp_mul(&t1, &p->y, &q->y);
    fp_mul(&t2, &p->z, &q->z);
    fp_add(&t3, &p->x, &p->y);
    fp_add(&t4, &q->x, &q->y);
    fp_mul(&t3, &t3, &t4);
    fp_add(&t4, &t0, &t1);
    fp_sub(&t3, &t3, &t4);
    fp_add(&t4, &p->y, &p->z);
    fp_add(&x3, &q->y, &q->z);
    fp_mul(&t4, &t4, &x3);
    fp_add(&x3, &t1, &t2);
    fp_sub(&t4, &t4, &x3);
    fp_add(&x3, &p->x, &p->z);
    fp_add(&y3, &q->x, &q->z);
    fp_mul(&x3, &x3, &y3);
    fp_add(&y3, &t0, &t2);
    fp_sub(&y3, &x3, &y3);
    fp_add(&x3, &t0, &t0);
    fp_add(&t0, &x3, &t0);
    fp_mul(&t2, &FP_B3, &t2);
    fp_add(&z3, &t1, &t2);
    fp_sub(&t1, &t1, &t2);
    fp_mul(&y3, &FP_B3, &y3);
    fp_mul(&x3, &t4, &y3);
    fp_mul(&t2, &t3, &t1);
    fp_sub(&x3, &t2, &x3);
    fp_mul(&y3, &y3, &t0);
    fp_mul(&t1, &t1, &z3);
    fp_add(&y3, &t1, &y3);
    fp_mul(&t0, &t0, &t3);
    fp_mul(&z3, &z3, &t4);
    fp_add(&z3, &z3, &t0);
    o->x = x3; o->y = y3; o->z = z3;
}
/* Mixed addition (RCB16 Alg. 8); an identity affine operand leaves p unchanged. */
static void g1p_add_mixed(g1p_t *o, const g1p_t *p, const g1a_t *q) {
    if (q->inf) { *o = *p; return; }
    fp_t t0, t1, t2, t3, t4, x3, y3, z3;
    fp_mul(&t0, &p->x, &q->x);
    fp_mul(&t1, &p->y, &q->y);
    fp_add(&t3, &q->x, &q->y);
    fp_add(&t4, &p->x, &p->y);
    fp_mul(&t3, &t3, &t4);
    fp_add(&t4, &t0, &t1);
    fp_sub(&t3, &t3, &t4);
    fp_mul(&t4, &q->y, &p->z);
    fp_add(&t4, &t4, &p->y);
    fp_mul(&y3, &q->x, &p->z);
    fp_add(&y3, &y3, &p->x);
    fp_add(&x3, &t0, &t0);
    fp_add(&t0, &x3, &t0);
    fp_mul(&t2, &FP_B3, &p->z);
    fp_add(&z3, &t1, &t2);
    fp_sub(&t1, &t1, &t2);
    fp_mul(&y3, &FP_B3, &y3);
    fp_mul(&x3, &t4, &y3);
    fp_mul(&t2, &t3, &t1);
    fp_sub(&x3, &t2, &x3);
    fp_mul(&y3, &y3, &t0);
    fp_mul(&t1, &t1, &z3);
    fp_add(&y3, &t1, &y3);
    fp_mul(&t0, &t0, &t3);
    fp_mul(&z3, &z3, &t4);
    fp_add(&z3, &z3, &t0);
    o->x = x3; o->y = y3; o->z = z3;
}
/* Doubling (RCB16 Alg. 9). */
static void g1p_double(g1p_t *o, const g1p_t *p) {
    fp_t t0, t1, t2, x3, y3, z3;
    fp_sqr(&t0, &p->y);
    fp_add(&z3, &t0, &t0);
    fp_add(&z3, &z3, &z3);
    fp_add(&z3, &z3, &z3);
    fp_mul(&t1, &p->y, &p->z);
    fp_sqr(&t2, &p->z);
    fp_mul(&t2, &FP_B3, &t2);
    fp_mul(&x3, &t2, &z3);
    fp_add(&y3, &t0, &t2);
    fp_mul(&z3, &t1, &z3);
    fp_add(&t1, &t2, &t2);
    fp_add(&t2, &t1, &t2);
    fp_sub(&t0, &t0, &t2);
    fp_mul(&y3, &t0, &y3);
    fp_add(&y3, &x3, &y3);
    fp_mul(&t1, &p->x, &p->y);
    fp_mul(&x3, &t0, &t1);
    fp_add(&x3, &x3, &x3);
    o->x = x3; o->y = y3; o->z = z3;
}
static void g1p_mul_u(g1p_t *o, const g1p_t *p, const uint64_t *k, int limbs) {
    g1p_t acc;
    g1p_identity(&acc);
    for (int i = limbs * 64 - 1; i >= 0; i--) {
        g1p_double(&acc, &acc);
        if ((k[i / 64] >> (i % 64)) & 1) g1p_add(&acc, &acc, p);
    }
    *o = acc;
}

/* Packed affine layout shared with the C ABI: x[6] ‖ y[6] u64 LE limbs, Montgomery. */
static void load_affine(g1a_t *o, const uint64_t *xy) {
    memcpy(o->x.l, xy, 48);
    memcpy(o->y.l, xy + 6, 48);
    o->inf = 0;
}

ORC_API void orc_g1_generator(uint64_t *xy_mont) {
    orc_init();
    fp_t x, y;
    memcpy(x.l, G1_GX, 48);
    memcpy(y.l, G1_GY, 48);
    fp_to_mont(&x, &x);
    fp_to_mont(&y, &y);
    memcpy(xy_mont, x.l, 48);
    memcpy(xy_mont + 6, y.l, 48);
}
/* 1 if y² = x³ + 4 for every packed affine point. */
ORC_API int orc_g1_on_curve(const uint64_t *xy, size_t n) {
    orc_init();
    fp_t b;
    fp_from_u64(&b, 4);
    for (size_t i = 0; i < n; i++) {
        g1a_t a;
        load_affine(&a, xy + 12 * i);
        fp_t l, r;
        fp_sqr(&l, &a.y);
        fp_sqr(&r, &a.x);
        fp_mul(&r, &r, &a.x);
        fp_add(&r, &r, &b);
        if (!fp_eq(&l, &r)) return 0;
    }
    return 1;
}
/* Projective (X,Y,Z Montgomery, 18 limbs) → canonical affine: out_xy = x‖y plain LE limbs
 * (non-Montgomery); returns 1 for the identity (out zeroed). */
ORC_API int orc_g1_proj_to_affine_canonical(const uint64_t *xyz, uint64_t *out_xy) {
    orc_init();
    g1p_t p;
    memcpy(p.x.l, xyz, 48);
    memcpy(p.y.l, xyz + 6, 48);
    memcpy(p.z.l, xyz + 12, 48);
    g1a_t a;
    g1a_from_proj(&a, &p);
    if (a.inf) { memset(out_xy, 0, 96); return 1; }
    fp_from_mont(&a.x, &a.x);
    fp_from_mont(&a.y, &a.y);
    memcpy(out_xy, a.x.l, 48);
    memcpy(out_xy + 6, a.y.l, 48);
    return 0;
}
/* Synthetic bases P_i = (a + i·d)·G, packed affine Montgomery. */
ORC_API void orc_synthetic_bases(size_t n, uint64_t a, uint64_t d, uint64_t *out_xy) {
    orc_init();
    uint64_t gxy[12];
    orc_g1_generator(gxy);
    g1a_t g;
    load_affine(&g, gxy);
    g1p_t gp, cur, step;
    g1p_from_affine(&gp, &g);
    g1p_mul_u(&cur, &gp, &a, 1);
    g1p_mul_u(&step, &gp, &d, 1);
    /* normalise in chunks with one inversion each (Montgomery's trick): (a + i·d) ≢ 0 mod r for the sizes used, so Z ≠ 0 */
    enum { CH = 1024 };
    g1p_t *pts = malloc(CH * sizeof(g1p_t));
    fp_t *pre = malloc(CH * sizeof(fp_t));
    for (size_t base = 0; base < n; base += CH) {
        const size_t m = n - base < CH ? n - base : CH;
        for (size_t i = 0; i < m; i++) {
            pts[i] = cur;
            if (i == 0) pre[0] = cur.z; else fp_mul(&pre[i], &pre[i - 1], &cur.z);
            g1p_add(&cur, &cur, &step);
        }
        fp_t inv;
        fp_inv(&inv, &pre[m - 1]);
        for (size_t i = m; i-- > 0;) {
            fp_t zi, x, y;
            if (i) { fp_mul(&zi, &inv, &pre[i - 1]); fp_mul(&inv, &inv, &pts[i].z); } else zi = inv;
            fp_mul(&x, &pts[i].x, &zi);
            fp_mul(&y, &pts[i].y, &zi);
            memcpy(out_xy + 12 * (base + i), x.l, 48);
            memcpy(out_xy + 12 * (base + i) + 6, y.l, 48);
        }
    }
    free(pts);
    free(pre);
}

/* ------------------------------------------------------------------------------------ MSM */
/* ⌊log2(a)·69/100⌋ — `ln_without_floats`. */
static unsigned ln_without_floats(size_t a) {
    unsigned lg = 0;
    while ((a >> (lg + 1)) != 0) lg++;
    return lg * 69 / 100;
}
typedef struct {
    const uint64_t *points, *scalars_canon;
    size_t n;
    unsigned c, n_windows;
    g1p_t *window_sums;
    volatile int next;
    pthread_mutex_t mu;
} msm_job_t;

static void msm_window(const msm_job_t *J, unsigned wi) {
    const unsigned c = J->c, w_start = wi * c;
    const size_t nb = ((size_t)1 << c) - 1;
    g1p_t res;
    g1p_identity(&res);
    g1p_t *buckets = malloc(nb * sizeof(g1p_t));
    for (size_t b = 0; b < nb; b++) g1p_identity(&buckets[b]);
    for (size_t i = 0; i < J->n; i++) {
        const uint64_t *s = J->scalars_canon + 4 * i;
        if ((s[0] | s[1] | s[2] | s[3]) == 0) continue;
        g1a_t pt;
        load_affine(&pt, J->points + 12 * i);
        if (s[0] == 1 && (s[1] | s[2] | s[3]) == 0) {
            if (w_start == 0) g1p_add_mixed(&res, &res, &pt);
            continue;
        }
        /* (scalar >> w_start) mod 2^c on the canonical limbs (`reduce()`, `divn`) */
        unsigned limb = w_start / 64, off = w_start % 64;
        uint64_t v = s[limb] >> off;
        if (off && limb + 1 < 4) v |= s[limb + 1] << (64 - off);
        uint64_t dgt = v & (((uint64_t)1 << c) - 1);
        if (dgt) g1p_add_mixed(&buckets[dgt - 1], &buckets[dgt - 1], &pt);
    }
    g1p_t running;
    g1p_identity(&running);
    for (size_t b = nb; b-- > 0;) {
        g1p_add(&running, &running, &buckets[b]);
        g1p_add(&res, &res, &running);
    }
    free(buckets);
    J->window_sums[wi] = res;
}
static void *msm_worker(void *arg) {
    msm_job_t *J = arg;
    for (;;) {
        pthread_mutex_lock(&J->mu);
        int w = J->next++;
        pthread_mutex_unlock(&J->mu);
        if (w >= (int)J->n_windows) break;
        msm_window(J, (unsigned)w);
    }
    return NULL;
}
/* msm_variable_base(points, scalars) → projective X‖Y‖Z (Montgomery, 18 limbs).
 * points: n × (x[6]‖y[6]) Montgomery; scalars: n × 4 limbs Montgomery (what `Scalar.0` holds).
 * threads = 1 is what /root/reference/Cargo.toml:19 builds; >1 models the rayon build. */
ORC_API void orc_msm_variable_base(const uint64_t *points, const uint64_t *scalars_mont, size_t n, uint64_t *out_xyz,
                                   int threads) {
    orc_init();
    msm_job_t J;
    memset(&J, 0, sizeof(J));
    uint64_t *canon = malloc(32 * (n ? n : 1));
    for (size_t i = 0; i < n; i++) {
        fr_t s;
        memcpy(s.l, scalars_mont + 4 * i, 32);
        fr_from_mont(&s, &s);
        memcpy(canon + 4 * i, s.l, 32);
    }
    J.points = points;
    J.scalars_canon = canon;
    J.n = n;
    J.c = n < 32 ? 3 : ln_without_floats(n) + 2;
    J.n_windows = (255 + J.c - 1) / J.c;
    J.window_sums = malloc(J.n_windows * sizeof(g1p_t));
    pthread_mutex_init(&J.mu, NULL);
    if (threads <= 1) {
        for (unsigned w = 0; w < J.n_windows; w++) msm_window(&J, w);
    } else {
        pthread_t *th = malloc(threads * sizeof(pthread_t));
        for (int t = 0; t < threads; t++) pthread_create(&th[t], NULL, msm_worker, &J);
        for (int t = 0; t < threads; t++) pthread_join(th[t], NULL);
        free(th);
    }
    g1p_t total;
    g1p_identity(&total);
    for (unsigned w = J.n_windows - 1; w >= 1; w--) {
        g1p_add(&total, &total, &J.window_sums[w]);
        for (unsigned k = 0; k < J.c; k++) g1p_double(&total, &total);
    }
    g1p_add(&total, &total, &J.window_sums[0]);
    memcpy(out_xyz, total.x.l, 48);
    memcpy(out_xyz + 6, total.y.l, 48);
    memcpy(out_xyz + 12, total.z.l, 48);
    free(J.window_sums);
    free(canon);
    pthread_mutex_destroy(&J.mu);
}
/* Σ sᵢ·Pᵢ by independent double-and-add (definition; O(255·n) group ops, small n only). */
ORC_API void orc_msm_naive(const uint64_t *points, const uint64_t *scalars_mont, size_t n, uint64_t *out_xyz) {
    orc_init();
    g1p_t total;
    g1p_identity(&total);
    for (size_t i = 0; i < n; i++) {
        fr_t s;
        memcpy(s.l, scalars_mont + 4 * i, 32);
        fr_from_mont(&s, &s);
        g1a_t a;
        load_affine(&a, points + 12 * i);
        g1p_t p, t;
        g1p_from_affine(&p, &a);
        g1p_mul_u(&t, &p, s.l, 4);
        g1p_add(&total, &total, &t);
    }
    memcpy(out_xyz, total.x.l, 48);
    memcpy(out_xyz + 6, total.y.l, 48);
    memcpy(out_xyz + 12, total.z.l, 48);
}

/* ------------------------------------------------------------------------------------ NTT */
static inline uint32_t bitrev32(uint32_t k, unsigned log) {
    uint32_t r = 0;
    for (unsigned i = 0; i < log; i++) { r = (r << 1) | (k & 1); k >>= 1; }
    return r;
}
static void fr_pow_u64(fr_t *o, const fr_t *a, uint64_t e) { fr_pow(o, a, &e, 1); }

/* serial_fft — SURVEY B.2: bit-reversal swaps, then log stages of radix-2 DIT butterflies with a
 * running twiddle (w *= w_m). */
static void serial_fft(fr_t *a, size_t n, const fr_t *omega, unsigned log) {
    for (size_t k = 0; k < n; k++) {
        size_t rk = bitrev32((uint32_t)k, log);
        if (k < rk) { fr_t t = a[k]; a[k] = a[rk]; a[rk] = t; }
    }
    size_t m = 1;
    for (unsigned s = 0; s < log; s++) {
        fr_t w_m;
        fr_pow_u64(&w_m, omega, n / (2 * m));
        for (size_t k = 0; k < n; k += 2 * m) {
            fr_t w = fr_R1;
            for (size_t j = 0; j < m; j++) {
                fr_t t;
                fr_mul(&t, &a[k + j + m], &w);
                fr_sub(&a[k + j + m], &a[k + j], &t);
                fr_add(&a[k + j], &a[k + j], &t);
                fr_mul(&w, &w, &w_m);
            }
        }
        m *= 2;
    }
}
/* parallel_fft — the `std`-feature path: 2^log_cpus interleaved sub-transforms of
 * size n / 2^log_cpus, each preceded by an O(n) twisting pass, then an un-shuffle. */
typedef struct {
    const fr_t *a;
    fr_t *tmp;
    const fr_t *omega;
    unsigned log_n, log_cpus;
    size_t j;
} pfft_arg_t;
static void *pfft_worker(void *vp) {
    pfft_arg_t *A = vp;
    const unsigned log_new_n = A->log_n - A->log_cpus;
    const size_t num_cpus = (size_t)1 << A->log_cpus, new_n = (size_t)1 << log_new_n, n = (size_t)1 << A->log_n;
    fr_t new_omega, omega_j, omega_step;
    fr_pow_u64(&new_omega, A->omega, num_cpus);
    fr_pow_u64(&omega_j, A->omega, A->j);
    fr_pow_u64(&omega_step, A->omega, (uint64_t)A->j << log_new_n);
    fr_t elt = fr_R1;
    fr_t *tmp = A->tmp;
    for (size_t i = 0; i < new_n; i++) {
        fr_t acc;
        memset(&acc, 0, sizeof(acc));
        for (size_t s = 0; s < num_cpus; s++) {
            size_t idx = (i + (s << log_new_n)) % n;
            fr_t t;
            fr_mul(&t, &A->a[idx], &elt);
            fr_add(&acc, &acc, &t);
            fr_mul(&elt, &elt, &omega_step);
        }
        tmp[i] = acc;
        fr_mul(&elt, &elt, &omega_j);
    }
    serial_fft(tmp, new_n, &new_omega, log_new_n);
    return NULL;
}
static void best_fft(fr_t *a, size_t n, const fr_t *omega, unsigned log, int threads) {
    unsigned log_cpus = 0;
    while ((2u << log_cpus) <= (unsigned)(threads < 1 ? 1 : threads)) log_cpus++;
    if (threads <= 1 || log <= log_cpus) { serial_fft(a, n, omega, log); return; }
    const size_t num_cpus = (size_t)1 << log_cpus, new_n = n >> log_cpus;
    fr_t *tmp = malloc(n * sizeof(fr_t));
    pfft_arg_t *args = malloc(num_cpus * sizeof(pfft_arg_t));
    pthread_t *th = malloc(num_cpus * sizeof(pthread_t));
    for (size_t j = 0; j < num_cpus; j++) {
        args[j] = (pfft_arg_t){a, tmp + j * new_n, omega, log, log_cpus, j};
        pthread_create(&th[j], NULL, pfft_worker, &args[j]);
    }
    for (size_t j = 0; j < num_cpus; j++) pthread_join(th[j], NULL);
    const size_t mask = num_cpus - 1;
    for (size_t idx = 0; idx < n; idx++) a[idx] = tmp[(idx & mask) * new_n + (idx >> log_cpus)];
    free(th);
    free(args);
    free(tmp);
}
static void distribute_powers(fr_t *a, size_t n, const fr_t *g) {
    fr_t p = fr_R1;
    for (size_t i = 0; i < n; i++) {
        fr_mul(&a[i], &a[i], &p);
        fr_mul(&p, &p, g);
    }
}
/* EvaluationDomain::{fft, ifft, coset_fft, coset_ifft} in place on 2^log_n Montgomery scalars.
 * Returns −1 when log_n ≥ 32 (EvaluationDomain::new's error).  threads as for the MSM. */
ORC_API int orc_ntt(uint64_t *data, uint32_t log_n, int inverse, int coset, int threads) {
    orc_init();
    if (log_n >= TWO_ADICITY) return -1;
    const size_t n = (size_t)1 << log_n;
    fr_t *a = (fr_t *)data;
    fr_t group_gen, group_gen_inv, size_inv, gen_inv, nn;
    fr_pow_u64(&group_gen, &FR_ROOT_OF_UNITY, (uint64_t)1 << (TWO_ADICITY - log_n));
    fr_inv(&group_gen_inv, &group_gen);
    fr_from_u64(&nn, n);
    fr_inv(&size_inv, &nn);
    fr_inv(&gen_inv, &FR_GENERATOR);
    if (!inverse) {
        if (coset) distribute_powers(a, n, &FR_GENERATOR);
        best_fft(a, n, &group_gen, log_n, threads);
    } else {
        best_fft(a, n, &group_gen_inv, log_n, threads);
        for (size_t i = 0; i < n; i++) fr_mul(&a[i], &a[i], &size_inv);
        if (coset) distribute_powers(a, n, &gen_inv);
    }
    return 0;
}
/* Domain constants (Montgomery) for the host-side mirror tests. */
ORC_API int orc_domain(uint32_t log_n, uint64_t *group_gen, uint64_t *group_gen_inv, uint64_t *size_inv,
                       uint64_t *generator_inv) {
    orc_init();
    if (log_n >= TWO_ADICITY) return -1;
    fr_t g, gi, nn, ni, seven_inv;
    fr_pow_u64(&g, &FR_ROOT_OF_UNITY, (uint64_t)1 << (TWO_ADICITY - log_n));
    fr_inv(&gi, &g);
    fr_from_u64(&nn, (uint64_t)1 << log_n);
    fr_inv(&ni, &nn);
    fr_inv(&seven_inv, &FR_GENERATOR);
    memcpy(group_gen, g.l, 32);
    memcpy(group_gen_inv, gi.l, 32);
    memcpy(size_inv, ni.l, 32);
    memcpy(generator_inv, seven_inv.l, 32);
    return 0;
}

/* ------------------------------------------------------------------------------------ PLONK prover restatement */
#include "plonk_oracle.inc"
