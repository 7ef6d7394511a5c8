// field.cuh — Montgomery prime-field arithmetic on 32-bit limbs for sm_100a.
//
// Replaces (on the GPU) the arithmetic of dusk-bls12_381 0.8's `Scalar` (Fr, 4×u64, R = 2^256) and
// `Fp` (6×u64, R = 2^384), pinned by /root/reference/Cargo.toml:20 and reached from the reference
// at /root/reference/src/zk/gadgets.rs:65-66,213,219.  The memory image is identical (little-endian
// limbs, Montgomery form, fully reduced), so 8 resp. 12 u32 limbs alias the Rust 4 resp. 6 u64.
//
// Multiplication is word-serial Montgomery (CIOS) with the partial products split into two
// accumulators — one collecting a[even]·b_i, one a[odd]·b_i one limb higher — so each row is two
// uninterrupted mad.lo.cc/madc.hi.cc carry chains (→ IMAD.WIDE.U32 with carry in SASS) and no
// carry ever has to be re-aligned.  See DESIGN.md §"Field arithmetic" for the derivation and the
// capacity argument (needs modulus < 2^(32N−1); Fr is 255 bits in 256, Fp 381 in 384).
#pragma once
#include "carry.cuh"

template <class P>
struct Field {
    static constexpr int N = P::N;
    uint32_t l[N];

    PB_HD static Field zero() {
        Field r;
#pragma unroll
        for (int i = 0; i < N; i++) r.l[i] = 0;
        return r;
    }
    PB_HD static Field one() {  // Montgomery 1 = R mod p
        Field r;
#pragma unroll
        for (int i = 0; i < N; i++) r.l[i] = P::r1(i);
        return r;
    }
    PB_HD static Field r2() {
        Field r;
#pragma unroll
        for (int i = 0; i < N; i++) r.l[i] = P::r2(i);
        return r;
    }
    PB_HD bool is_zero() const {
        uint32_t o = 0;
#pragma unroll
        for (int i = 0; i < N; i++) o |= l[i];
        return o == 0;
    }
    PB_HD bool operator==(const Field &b) const {
        uint32_t o = 0;
#pragma unroll
        for (int i = 0; i < N; i++) o |= l[i] ^ b.l[i];
        return o == 0;
    }
    PB_HD bool operator!=(const Field &b) const { return !(*this == b); }

    // r = a − p if a ≥ p else a   (a < 2p)
    PB_HD static Field reduce_once(const Field &a) {
        Field d;
        d.l[0] = cc::sub_cc(a.l[0], P::mod(0));
#pragma unroll
        for (int i = 1; i < N; i++) d.l[i] = cc::subc_cc(a.l[i], P::mod(i));
        uint32_t borrow = cc::subc(0, 0);  // 0xffffffff when a < p
        Field r;
#pragma unroll
        for (int i = 0; i < N; i++) r.l[i] = borrow ? a.l[i] : d.l[i];
        return r;
    }
    PB_HD friend Field operator+(const Field &a, const Field &b) {
        Field s;
        s.l[0] = cc::add_cc(a.l[0], b.l[0]);
#pragma unroll
        for (int i = 1; i < N - 1; i++) s.l[i] = cc::addc_cc(a.l[i], b.l[i]);
        s.l[N - 1] = cc::addc(a.l[N - 1], b.l[N - 1]);  // a+b < 2p < 2^(32N): no carry out
        return reduce_once(s);
    }
    PB_HD friend Field operator-(const Field &a, const Field &b) {
        Field d;
        d.l[0] = cc::sub_cc(a.l[0], b.l[0]);
#pragma unroll
        for (int i = 1; i < N; i++) d.l[i] = cc::subc_cc(a.l[i], b.l[i]);
        uint32_t mask = cc::subc(0, 0);  // all-ones when a < b
        Field r;
        r.l[0] = cc::add_cc(d.l[0], P::mod(0) & mask);
#pragma unroll
        for (int i = 1; i < N - 1; i++) r.l[i] = cc::addc_cc(d.l[i], P::mod(i) & mask);
        r.l[N - 1] = cc::addc(d.l[N - 1], P::mod(N - 1) & mask);
        return r;
    }
    PB_HD Field neg() const { return is_zero() ? *this : (zero() - *this); }
    PB_HD Field dbl() const { return *this + *this; }

    // acc[0..N) += Σ_{j even} x[j]·y·2^(32j); carry-out left in the flag.
    PB_HD static void chain_mad(uint32_t *acc, const uint32_t *x, uint32_t y) {
        acc[0] = cc::mad_lo_cc(x[0], y, acc[0]);
        acc[1] = cc::madc_hi_cc(x[0], y, acc[1]);
#pragma unroll
        for (int j = 2; j < N; j += 2) {
            acc[j] = cc::madc_lo_cc(x[j], y, acc[j]);
            acc[j + 1] = cc::madc_hi_cc(x[j], y, acc[j + 1]);
        }
    }
    // One Montgomery elimination step on T = E + O·2^32: adds m·p with m = −E[0]/p mod 2^32, so E[0]
    // becomes 0.  p's limbs are compile-time immediates.
    PB_HD static void redc_step(uint32_t *E, uint32_t *O) {
        uint32_t m = cc::mul_lo(E[0], P::INV32);
        O[0] = cc::mad_lo_cc(P::mod(1), m, O[0]);
        O[1] = cc::madc_hi_cc(P::mod(1), m, O[1]);
#pragma unroll
        for (int j = 2; j < N; j += 2) {
            O[j] = cc::madc_lo_cc(P::mod(j + 1), m, O[j]);
            O[j + 1] = cc::madc_hi_cc(P::mod(j + 1), m, O[j + 1]);
        }
        E[0] = cc::mad_lo_cc(P::mod(0), m, E[0]);
        E[1] = cc::madc_hi_cc(P::mod(0), m, E[1]);
#pragma unroll
        for (int j = 2; j < N; j += 2) {
            E[j] = cc::madc_lo_cc(P::mod(j), m, E[j]);
            E[j + 1] = cc::madc_hi_cc(P::mod(j), m, E[j + 1]);
        }
        O[N - 1] = cc::addc(O[N - 1], 0);
    }
    // Round i ≥ 1.  On entry the running value (already divided by 2^32) is  S[1..] + D, where S is
    // the array whose limb 0 was just zeroed.  On exit D plays S's role for the next round.
    PB_HD static void round(uint32_t *S, uint32_t *D, const uint32_t *a, uint32_t bi) {
        D[0] = cc::add_cc(D[0], S[1]);  // fold S[1] into the new even accumulator, carry → odd chain
#pragma unroll
        for (int j = 0; j < N - 2; j += 2) {  // new odd accumulator = S[2..] + a[odd]·bi, written over S
            S[j] = cc::madc_lo_cc(a[j + 1], bi, S[j + 2]);
            S[j + 1] = cc::madc_hi_cc(a[j + 1], bi, S[j + 3]);
        }
        S[N - 2] = cc::madc_lo_cc(a[N - 1], bi, 0);
        S[N - 1] = cc::madc_hi(a[N - 1], bi, 0);
        chain_mad(D, a, bi);  // even accumulator += a[even]·bi
        S[N - 1] = cc::addc(S[N - 1], 0);
        redc_step(D, S);
    }
    PB_HD friend Field operator*(const Field &a, const Field &b) {
        uint32_t A[N], B[N];  // A: even accumulator first, B: odd accumulator first
#pragma unroll
        for (int j = 0; j < N; j += 2) {
            A[j] = cc::mul_lo(a.l[j], b.l[0]);
            A[j + 1] = cc::mul_hi(a.l[j], b.l[0]);
            B[j] = cc::mul_lo(a.l[j + 1], b.l[0]);
            B[j + 1] = cc::mul_hi(a.l[j + 1], b.l[0]);
        }
        redc_step(A, B);
#pragma unroll
        for (int i = 1; i < N - 1; i += 2) {
            round(A, B, a.l, b.l[i]);
            round(B, A, a.l, b.l[i + 1]);
        }
        round(A, B, a.l, b.l[N - 1]);
        // the last round left B with limb 0 cleared: value = B[1..] + A  (< 2p)
        Field r;
        r.l[0] = cc::add_cc(A[0], B[1]);
#pragma unroll
        for (int k = 1; k < N - 1; k++) r.l[k] = cc::addc_cc(A[k], B[k + 1]);
        r.l[N - 1] = cc::addc(A[N - 1], 0);
        return reduce_once(r);
    }
    PB_HD Field sqr() const { return *this * *this; }

    // Montgomery form → canonical integer (multiply by 1): N elimination rounds, no products with b.
    PB_HD Field from_mont() const {
        Field one_raw = zero();
        one_raw.l[0] = 1;
        return *this * one_raw;
    }
    PB_HD Field to_mont() const { return *this * r2(); }

    // this^e, e a plain little-endian u32 exponent of `words` words.
    PB_HD Field pow(const uint32_t *e, int words) const {
        Field r = one();
        for (int i = words * 32 - 1; i >= 0; i--) {
            r = r.sqr();
            if ((e[i >> 5] >> (i & 31)) & 1) r = r * *this;
        }
        return r;
    }
    PB_HD Field pow_u64(uint64_t e) const {
        uint32_t w[2] = {(uint32_t)e, (uint32_t)(e >> 32)};
        return pow(w, 2);
    }
    // Fermat inverse; 0 ↦ 0.
    PB_HD Field inv() const {
        uint32_t e[N];
        e[0] = cc::sub_cc(P::mod(0), 2);
#pragma unroll
        for (int i = 1; i < N; i++) e[i] = cc::subc_cc(P::mod(i), 0);
        return pow(e, N);
    }
};

// ---------------------------------------------------------------------------------------------------
// BLS12-381 scalar field Fr.  r ≡ 1 (mod 2^32) ⇒ −r⁻¹ mod 2^32 = 0xffffffff.
struct FrParams {
    static constexpr int N = 8;
    static constexpr uint32_t INV32 = 0xffffffffu;
    PB_HD static constexpr uint32_t mod(int i) {
        constexpr uint32_t v[8] = {0x00000001u, 0xffffffffu, 0xfffe5bfeu, 0x53bda402u,
                                   0x09a1d805u, 0x3339d808u, 0x299d7d48u, 0x73eda753u};
        return v[i];
    }
    PB_HD static constexpr uint32_t r1(int i) {  // 2^256 mod r
        constexpr uint32_t v[8] = {0xfffffffeu, 0x00000001u, 0x00034802u, 0x5884b7fau,
                                   0xecbc4ff5u, 0x998c4fefu, 0xacc5056fu, 0x1824b159u};
        return v[i];
    }
    PB_HD static constexpr uint32_t r2(int i) {  // 2^512 mod r
        constexpr uint32_t v[8] = {0xf3f29c6du, 0xc999e990u, 0x87925c23u, 0x2b6cedcbu,
                                   0x7254398fu, 0x05d31496u, 0x9f59ff11u, 0x0748d9d9u};
        return v[i];
    }
};
// BLS12-381 base field Fp.
struct FpParams {
    static constexpr int N = 12;
    static constexpr uint32_t INV32 = 0xfffcfffdu;
    PB_HD static constexpr uint32_t mod(int i) {
        constexpr uint32_t v[12] = {0xffffaaabu, 0xb9feffffu, 0xb153ffffu, 0x1eabfffeu, 0xf6b0f624u, 0x6730d2a0u,
                                    0xf38512bfu, 0x64774b84u, 0x434bacd7u, 0x4b1ba7b6u, 0x397fe69au, 0x1a0111eau};
        return v[i];
    }
    PB_HD static constexpr uint32_t r1(int i) {  // 2^384 mod p
        constexpr uint32_t v[12] = {0x0002fffdu, 0x76090000u, 0xc40c0002u, 0xebf4000bu, 0x53c758bau, 0x5f489857u,
                                    0x70525745u, 0x77ce5853u, 0xa256ec6du, 0x5c071a97u, 0xfa80e493u, 0x15f65ec3u};
        return v[i];
    }
    PB_HD static constexpr uint32_t r2(int i) {  // 2^768 mod p
        constexpr uint32_t v[12] = {0x1c341746u, 0xf4df1f34u, 0x09d104f1u, 0x0a76e6a6u, 0x4c95b6d5u, 0x8de5476cu,
                                    0x939d83c0u, 0x67eb88a9u, 0xb519952du, 0x9a793e85u, 0x92cae3aau, 0x11988fe5u};
        return v[i];
    }
};
typedef Field<FrParams> Fr;
typedef Field<FpParams> Fp;
