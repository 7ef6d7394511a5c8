// Host-emulation harness: runs the *device* field templates (csrc/field.cuh) on the CPU with the
// carry-flag primitives emulated (PB200_HOST_EMU), exposing them through a C ABI for pytest.
#define PB200_HOST_EMU 1
#include "../../plonk-prototype_b200/csrc/field.cuh"
#include <cstddef>
#include <cstring>
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
            default: z = F::one();
        }
        memcpy(o + F::N * i, z.l, 4 * F::N);
    }
}
extern "C" {
void emu_fr(int op, const uint32_t *a, const uint32_t *b, uint32_t *o, size_t n) { binop<Fr>(op, a, b, o, n); }
void emu_fp(int op, const uint32_t *a, const uint32_t *b, uint32_t *o, size_t n) { binop<Fp>(op, a, b, o, n); }
}
