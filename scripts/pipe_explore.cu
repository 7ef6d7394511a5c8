// pipe_explore.cu — what the dispatch of a B200 SM sustains for the instruction kinds a big-integer multiplier can be built
// from, alone and mixed: IMAD.WIDE.U32 (the shipped multipliers), IADD3 (their carry / correction work), DFMA (the 52-bit-limb
// floating-point route: hi = fma(a, b, 0) rz, lo = fma(a, b, −hi)).  Answers two questions left open by
// profiles/pipe_model_experiments_r02.txt: (1) how much ALU-pipe work rides free next to a stream of wide multiplies,
// (2) what DFMA sustains on this part and whether it overlaps with IMAD.WIDE.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o /tmp/pipe_explore scripts/pipe_explore.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

// per inner step: NW wide multiplies, NA 3-input adds, ND double FMAs, all on independent accumulators
template <int NW, int NA, int ND>
__global__ void k(uint64_t *sink, uint32_t iters, uint32_t b0) {
    uint64_t w[NW > 0 ? NW : 1];
    uint32_t a[NA > 0 ? NA : 1];
    double d[ND > 0 ? ND : 1];
    const uint32_t b = b0 | 1u;
    const double db = 1.0000001 + 1e-9 * (double)(b0 & 7);
#pragma unroll
    for (int c = 0; c < (NW > 0 ? NW : 1); c++) w[c] = (uint64_t)(threadIdx.x + 1) * 0x9E3779B97F4A7C15ull + c * 77 + blockIdx.x;
#pragma unroll
    for (int c = 0; c < (NA > 0 ? NA : 1); c++) a[c] = threadIdx.x * 2654435761u + c;
#pragma unroll
    for (int c = 0; c < (ND > 0 ? ND : 1); c++) d[c] = 1.0 + 1e-3 * (threadIdx.x + c);
    for (uint32_t i = 0; i < iters; i++) {
#pragma unroll
        for (int r = 0; r < 8; r++) {
#pragma unroll
            for (int c = 0; c < NW; c++) {
                uint32_t lo = (uint32_t)w[c];
                asm volatile("mad.wide.u32 %0, %1, %2, %0;" : "+l"(w[c]) : "r"(lo), "r"(b));
            }
#pragma unroll
            for (int c = 0; c < NA; c++) asm volatile("add.u32 %0, %0, %1;\n\txor.b32 %0, %0, %2;" : "+r"(a[c]) : "r"(b), "r"(i));
#pragma unroll
            for (int c = 0; c < ND; c++) asm volatile("fma.rz.f64 %0, %0, %1, %2;" : "+d"(d[c]) : "d"(db), "d"(1e-30));
        }
    }
    uint64_t s = 0;
#pragma unroll
    for (int c = 0; c < (NW > 0 ? NW : 1); c++) s ^= w[c];
#pragma unroll
    for (int c = 0; c < (NA > 0 ? NA : 1); c++) s ^= a[c];
#pragma unroll
    for (int c = 0; c < (ND > 0 ? ND : 1); c++) s ^= (uint64_t)__double_as_longlong(d[c]);
    sink[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int NW, int NA, int ND>
void run(int sms, uint64_t *sink, double clk_hz) {
    const int threads = 512, ctas_per_sm = 1;   // 16 warps per SM
    const uint32_t iters = 4096;
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    float best = 1e30f;
    for (int rep = 0; rep < 4; rep++) {
        cudaEventRecord(e0);
        k<NW, NA, ND><<<sms * ctas_per_sm, threads>>>(sink, iters, 12345u + rep);
        cudaEventRecord(e1);
        cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        if (rep && ms < best) best = ms;
    }
    const double steps = (double)sms * threads * iters * 8.0, sec = best * 1e-3, per = steps / sec / sms / clk_hz;   // inner steps per clk per SM (lanes)
    printf("wide=%d add=%d dfma=%d : %.3f ms | per clk per SM: %.1f IMAD.WIDE lanes, %.1f add lanes, %.1f DFMA lanes | cycles per warp-step per scheduler %.2f\n",
           NW, NA, ND, best, per * NW, per * NA * 2, per * ND, 32.0 * 4 / per);
}

int main() {
    cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
    int khz = 0; cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, 0);
    const double clk = 1.965e9;
    printf("%s, %d SMs, attr clock %d kHz (rates below assume 1965 MHz)\n", p.name, p.multiProcessorCount, khz);
    uint64_t *sink; cudaMalloc(&sink, (size_t)p.multiProcessorCount * 4 * 1024 * 8);
    const int s = p.multiProcessorCount;
    run<8, 0, 0>(s, sink, clk);   // wide multiplies only
    run<0, 8, 0>(s, sink, clk);   // adds only (2 per unit)
    run<0, 0, 8>(s, sink, clk);   // DFMA only
    run<8, 2, 0>(s, sink, clk);   // 8 wide + 4 adds
    run<8, 4, 0>(s, sink, clk);   // 8 wide + 8 adds
    run<8, 8, 0>(s, sink, clk);   // 8 wide + 16 adds
    run<8, 12, 0>(s, sink, clk);  // 8 wide + 24 adds
    run<8, 0, 4>(s, sink, clk);   // 8 wide + 4 DFMA
    run<8, 0, 8>(s, sink, clk);   // 8 wide + 8 DFMA
    run<8, 0, 16>(s, sink, clk);  // 8 wide + 16 DFMA
    run<4, 0, 16>(s, sink, clk);
    run<0, 8, 8>(s, sink, clk);   // adds + DFMA
    cudaError_t e = cudaDeviceSynchronize();
    printf("status: %s\n", cudaGetErrorString(e));
    return e != cudaSuccess;
}
