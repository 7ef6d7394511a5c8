// Host-emulation harness: runs the *device* field templates (csrc/field.cuh) on the CPU with the
// carry-flag primitives emulated (PB200_HOST_EMU), exposing them through a C ABI for pytest.
#define PB200_HOST_EMU 1
#include "../../plonk-prototype_b200/csrc/field.cuh"
#include <cstddef>
#include <cstring>
// x + p as a plain multi-limb addition (no reduction): lifts a canonical value into the upper half of the lazy range [0, 2p)
template <class F> struct ParamsOf;
template <> struct ParamsOf<Fr> { typedef FrParams T; };
template <> struct ParamsOf<Fp> { typedef FpParams T; };
template <class F> static F lift(const F &x) {
    F r;
    uint64_t c = 0;
    for (int i = 0; i < F::N; i++) {
        const uint64_t t = (uint64_t)x.l[i] + ParamsOf<F>::T::mod(i) + c;
        r.l[i] = (uint32_t)t;
        c = t >> 32;
    }
    return r;   // x < p ⇒ x + p < 2p < 2^(32N): no carry out
}
// is v < 2p ?  (v − 2p borrows)
template <class F> static bool below_2p(const F &v) {
    int64_t borrow = 0;
    for (int i = 0; i < F::N; i++) {
        const int64_t t = (int64_t)v.l[i] - (int64_t)F::mod2(i) - borrow;
        borrow = t < 0;
    }
    return borrow != 0;
}
template <class F> static F checked(const F &v) {
    if (!below_2p(v)) __builtin_trap();     // a lazy operation left its range
    return v.canonical();
}
template <class F> static void binop(int op, const uint32_t *a, const uint32_t *b, uint32_t *o, size_t n) {
    for (size_t i = 0; i < n; i++) {
        F x, y, z;
        memcpy(x.l, a + F::N * i, 4 * F::N);
        memcpy(y.l, b + F::N * i, 4 * F::N);
        switch (op) {
            case 0: z = x * y; break;
            case 1: z = x + y; break;
            case 2: z = x - y; break;
            case 3: z = x.from_mont(); break;
            case 4: z = x.to_mont(); break;
            case 5: z = x.inv(); break;
            case 6: z = x.neg(); break;
            case 7: z = x.sqr(); break;
            case 14: z = x.is_zero() ? x : x.neg_nonzero(); break;
            // lazy representation [0, 2p): operands taken from both halves of the range, results must stay inside it
            case 9: z = checked(F::mul_lazy(x, lift(y))); if (z != checked(F::mul_lazy(x, y))) __builtin_trap(); break;
            case 10: z = checked(F::add_lazy(lift(x), lift(y))); if (z != checked(F::add_lazy(x, lift(y))) || z != checked(F::add_lazy(lift(x), y)) || z != checked(F::add_lazy(x, y))) __builtin_trap(); break;
            case 12: z = checked(F::mul_unreduced(lift(x), lift(y))); if (z != checked(F::mul_unreduced(x, lift(y))) || z != checked(F::mul_unreduced(lift(x), y))) __builtin_trap(); break;   // both operands lazy (Fp only)
            case 13: z = checked(F::sqr_unreduced(lift(x))); if (z != checked(F::sqr_unreduced(x))) __builtin_trap(); break;
            case 11: z = checked(F::sub_lazy(lift(x), lift(y))); if (z != checked(F::sub_lazy(x, lift(y))) || z != checked(F::sub_lazy(lift(x), y)) || z != checked(F::sub_lazy(x, y))) __builtin_trap(); break;
            default: z = F::one();
        }
        memcpy(o + F::N * i, z.l, 4 * F::N);
    }
}
extern "C" {
void emu_fr(int op, const uint32_t *a, const uint32_t *b, uint32_t *o, size_t n) { binop<Fr>(op, a, b, o, n); }
void emu_fp(int op, const uint32_t *a, const uint32_t *b, uint32_t *o, size_t n) { binop<Fp>(op, a, b, o, n); }
}

// ---- G1 (csrc/g1.cuh) -------------------------------------------------------------------------------
#include "../../plonk-prototype_b200/csrc/g1.cuh"
static G1Affine ld_aff(const uint32_t *p) { G1Affine a; memcpy(a.x.l, p, 48); memcpy(a.y.l, p + 12, 48); return a; }
static void st_xyzz_affine(const G1Xyzz &v, uint32_t *o) {  // o: 24 words x‖y, then 1 word "is identity"
    G1Affine a;
    if (!g1_to_affine(v, a)) { memset(o, 0, 96); o[24] = 1; return; }
    memcpy(o, a.x.l, 48); memcpy(o + 12, a.y.l, 48); o[24] = 0;
}
extern "C" {
// op 0: P+Q via madd; 1: P+Q via add (both XYZZ, second one re-randomised by doubling twice and adding back);
// 2: 2P via dbl_affine; 3: 2P via dbl on XYZZ; 4: k·P small; 5: chain: ((P+Q)+Q)+(−Q) via madd;
// inputs are packed affine Montgomery (24 words each); output 25 words per element.
void emu_g1(int op, const uint32_t *p, const uint32_t *q, uint32_t *o, size_t n, uint64_t k) {
    for (size_t i = 0; i < n; i++) {
        G1Affine P = ld_aff(p + 24 * i), Q = ld_aff(q + 24 * i);
        G1Xyzz r;
        switch (op) {
            case 0: r = G1Xyzz::from_affine(P); g1_madd(r, Q); break;
            case 1: {
                G1Xyzz a = g1_dbl(g1_dbl_affine(P));           // 4P in non-trivial ZZ
                G1Xyzz nP = G1Xyzz::from_affine(P); nP.y = nP.y.neg();
                a = g1_add(a, nP); a = g1_add(a, nP); a = g1_add(a, nP);  // back to P with ZZ ≠ 1
                G1Xyzz b = g1_dbl_affine(Q); G1Xyzz nQ = G1Xyzz::from_affine(Q); nQ.y = nQ.y.neg();
                b = g1_add(b, nQ);                             // Q with ZZ ≠ 1
                r = g1_add(a, b);
                break;
            }
            case 2: r = g1_dbl_affine(P); break;
            case 3: r = g1_dbl(g1_dbl_affine(P)); break;
            case 4: r = g1_mul_small(G1Xyzz::from_affine(P), k); break;
            case 5: { r = G1Xyzz::from_affine(P); g1_madd(r, Q); g1_madd(r, Q); G1Affine nQ = Q; nQ.y = nQ.y.neg(); g1_madd(r, nQ); break; }
            case 7: {  // the lazy mixed addition: P, then Q three times, −Q, P again (doubling path), through g1_madd_lazy
                r = G1Xyzz::identity(); g1_madd_lazy(r, P); g1_madd_lazy(r, Q); g1_madd_lazy(r, Q); g1_madd_lazy(r, Q);
                G1Affine nQ = Q; nQ.y = nQ.y.neg(); g1_madd_lazy(r, nQ); g1_madd_lazy(r, P);
                if (!below_2p(r.x) || !below_2p(r.y) || !below_2p(r.zz) || !below_2p(r.zzz)) __builtin_trap();
                r = g1_canonical(r); break; }
            case 8: { r = G1Xyzz::identity(); g1_madd_lazy(r, P); g1_madd_lazy(r, P); G1Affine nP = P; nP.y = nP.y.neg(); g1_madd_lazy(r, nP); g1_madd_lazy(r, nP); g1_madd_lazy(r, Q); r = g1_canonical(r); break; }   // doubling, then cancellation to the identity, then Q
            case 6: { r = G1Xyzz::identity(); g1_madd(r, P); G1Affine nP = P; nP.y = nP.y.neg(); g1_madd(r, nP); g1_madd(r, Q); break; }
            default: r = G1Xyzz::identity();
        }
        st_xyzz_affine(r, o + 25 * i);
    }
}
}
