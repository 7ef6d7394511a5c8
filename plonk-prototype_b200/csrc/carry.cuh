// carry.cuh — 32-bit carry-chain primitives.
//
// On the device every primitive is one PTX instruction.  The multiplier is written on 64-bit columns:
// mul.wide.u32 feeding add.cc.u64 / addc.cc.u64, which ptxas fuses into one IMAD.WIDE.U32.X (32×32+64
// with carry-in/out through a predicate) — the instruction the integer roofline in DESIGN.md counts.
// (Measured with cuobjdump: the 32-bit mad.lo.cc/madc.hi.cc pair form fuses only some of the time.)  On the host (PB200_HOST_EMU builds
// used by tests/test_host_emulation.py) the same names are emulated with an explicit carry flag so
// the multi-limb algorithms built on top can be checked on a machine without a GPU.  The host
// emulation is test scaffolding only: the shipped library never calls it.
#pragma once
#include <stdint.h>

#if defined(__CUDA_ARCH__)
#define PB_HD __device__ __forceinline__
namespace cc {
PB_HD uint32_t add_cc(uint32_t a, uint32_t b) { uint32_t r; asm volatile("add.cc.u32 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b)); return r; }
PB_HD uint32_t addc_cc(uint32_t a, uint32_t b) { uint32_t r; asm volatile("addc.cc.u32 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b)); return r; }
PB_HD uint32_t addc(uint32_t a, uint32_t b) { uint32_t r; asm volatile("addc.u32 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b)); return r; }
PB_HD uint32_t sub_cc(uint32_t a, uint32_t b) { uint32_t r; asm volatile("sub.cc.u32 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b)); return r; }
PB_HD uint32_t subc_cc(uint32_t a, uint32_t b) { uint32_t r; asm volatile("subc.cc.u32 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b)); return r; }
PB_HD uint32_t subc(uint32_t a, uint32_t b) { uint32_t r; asm volatile("subc.u32 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b)); return r; }
PB_HD uint32_t mul_lo(uint32_t a, uint32_t b) { uint32_t r; asm volatile("mul.lo.u32 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b)); return r; }
PB_HD uint32_t mul_hi(uint32_t a, uint32_t b) { uint32_t r; asm volatile("mul.hi.u32 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b)); return r; }
PB_HD uint32_t mad_lo_cc(uint32_t a, uint32_t b, uint32_t c) { uint32_t r; asm volatile("mad.lo.cc.u32 %0, %1, %2, %3;" : "=r"(r) : "r"(a), "r"(b), "r"(c)); return r; }
PB_HD uint32_t mad_hi_cc(uint32_t a, uint32_t b, uint32_t c) { uint32_t r; asm volatile("mad.hi.cc.u32 %0, %1, %2, %3;" : "=r"(r) : "r"(a), "r"(b), "r"(c)); return r; }
PB_HD uint32_t madc_lo_cc(uint32_t a, uint32_t b, uint32_t c) { uint32_t r; asm volatile("madc.lo.cc.u32 %0, %1, %2, %3;" : "=r"(r) : "r"(a), "r"(b), "r"(c)); return r; }
PB_HD uint32_t madc_hi_cc(uint32_t a, uint32_t b, uint32_t c) { uint32_t r; asm volatile("madc.hi.cc.u32 %0, %1, %2, %3;" : "=r"(r) : "r"(a), "r"(b), "r"(c)); return r; }
PB_HD uint32_t madc_lo(uint32_t a, uint32_t b, uint32_t c) { uint32_t r; asm volatile("madc.lo.u32 %0, %1, %2, %3;" : "=r"(r) : "r"(a), "r"(b), "r"(c)); return r; }
PB_HD uint32_t madc_hi(uint32_t a, uint32_t b, uint32_t c) { uint32_t r; asm volatile("madc.hi.u32 %0, %1, %2, %3;" : "=r"(r) : "r"(a), "r"(b), "r"(c)); return r; }
PB_HD uint64_t mul_wide(uint32_t a, uint32_t b) { uint64_t r; asm("mul.wide.u32 %0, %1, %2;" : "=l"(r) : "r"(a), "r"(b)); return r; }
PB_HD uint64_t add_cc64(uint64_t a, uint64_t b) { uint64_t r; asm volatile("add.cc.u64 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b)); return r; }
PB_HD uint64_t addc_cc64(uint64_t a, uint64_t b) { uint64_t r; asm volatile("addc.cc.u64 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b)); return r; }
PB_HD uint64_t addc64(uint64_t a, uint64_t b) { uint64_t r; asm volatile("addc.u64 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b)); return r; }
PB_HD uint32_t lo32(uint64_t x) { return (uint32_t)x; }
PB_HD uint32_t hi32(uint64_t x) { return (uint32_t)(x >> 32); }
PB_HD uint64_t pack64(uint32_t l, uint32_t h) { uint64_t r; asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "r"(l), "r"(h)); return r; }
}  // namespace cc
#elif defined(PB200_HOST_EMU)
#define PB_HD inline
namespace cc {
static thread_local uint32_t CF = 0;  // the emulated carry / borrow flag
PB_HD uint32_t add_cc(uint32_t a, uint32_t b) { uint64_t s = (uint64_t)a + b; CF = (uint32_t)(s >> 32); return (uint32_t)s; }
PB_HD uint32_t addc_cc(uint32_t a, uint32_t b) { uint64_t s = (uint64_t)a + b + CF; CF = (uint32_t)(s >> 32); return (uint32_t)s; }
PB_HD uint32_t addc(uint32_t a, uint32_t b) { return a + b + CF; }
PB_HD uint32_t sub_cc(uint32_t a, uint32_t b) { uint64_t s = (uint64_t)a - b; CF = (uint32_t)(s >> 63); return (uint32_t)s; }
PB_HD uint32_t subc_cc(uint32_t a, uint32_t b) { uint64_t s = (uint64_t)a - b - CF; CF = (uint32_t)(s >> 63); return (uint32_t)s; }
PB_HD uint32_t subc(uint32_t a, uint32_t b) { return a - b - CF; }
PB_HD uint32_t mul_lo(uint32_t a, uint32_t b) { return a * b; }
PB_HD uint32_t mul_hi(uint32_t a, uint32_t b) { return (uint32_t)(((uint64_t)a * b) >> 32); }
PB_HD uint32_t mad_lo_cc(uint32_t a, uint32_t b, uint32_t c) { return add_cc(mul_lo(a, b), c); }
PB_HD uint32_t mad_hi_cc(uint32_t a, uint32_t b, uint32_t c) { return add_cc(mul_hi(a, b), c); }
PB_HD uint32_t madc_lo_cc(uint32_t a, uint32_t b, uint32_t c) { return addc_cc(mul_lo(a, b), c); }
PB_HD uint32_t madc_hi_cc(uint32_t a, uint32_t b, uint32_t c) { return addc_cc(mul_hi(a, b), c); }
PB_HD uint32_t madc_lo(uint32_t a, uint32_t b, uint32_t c) { return addc(mul_lo(a, b), c); }
PB_HD uint32_t madc_hi(uint32_t a, uint32_t b, uint32_t c) { return addc(mul_hi(a, b), c); }
PB_HD uint64_t mul_wide(uint32_t a, uint32_t b) { return (uint64_t)a * b; }
PB_HD uint64_t add_cc64(uint64_t a, uint64_t b) { unsigned __int128 s = (unsigned __int128)a + b; CF = (uint32_t)(s >> 64); return (uint64_t)s; }
PB_HD uint64_t addc_cc64(uint64_t a, uint64_t b) { unsigned __int128 s = (unsigned __int128)a + b + CF; CF = (uint32_t)(s >> 64); return (uint64_t)s; }
PB_HD uint64_t addc64(uint64_t a, uint64_t b) { return a + b + CF; }
PB_HD uint32_t lo32(uint64_t x) { return (uint32_t)x; }
PB_HD uint32_t hi32(uint64_t x) { return (uint32_t)(x >> 32); }
PB_HD uint64_t pack64(uint32_t l, uint32_t h) { return ((uint64_t)h << 32) | l; }
}  // namespace cc
#else
// Host pass of nvcc over a .cu file: device functions are parsed but never executed.
#define PB_HD __device__ __forceinline__
namespace cc {
#define PB_STUB2(n) PB_HD uint32_t n(uint32_t, uint32_t) { return 0; }
#define PB_STUB3(n) PB_HD uint32_t n(uint32_t, uint32_t, uint32_t) { return 0; }
PB_STUB2(add_cc) PB_STUB2(addc_cc) PB_STUB2(addc) PB_STUB2(sub_cc) PB_STUB2(subc_cc) PB_STUB2(subc)
PB_STUB2(mul_lo) PB_STUB2(mul_hi)
PB_STUB3(mad_lo_cc) PB_STUB3(mad_hi_cc) PB_STUB3(madc_lo_cc) PB_STUB3(madc_hi_cc) PB_STUB3(madc_lo) PB_STUB3(madc_hi)
PB_HD uint64_t mul_wide(uint32_t, uint32_t) { return 0; }
PB_HD uint64_t add_cc64(uint64_t, uint64_t) { return 0; }
PB_HD uint64_t addc_cc64(uint64_t, uint64_t) { return 0; }
PB_HD uint64_t addc64(uint64_t, uint64_t) { return 0; }
PB_HD uint32_t lo32(uint64_t) { return 0; }
PB_HD uint32_t hi32(uint64_t) { return 0; }
PB_HD uint64_t pack64(uint32_t, uint32_t) { return 0; }
#undef PB_STUB2
#undef PB_STUB3
}  // namespace cc
#endif
