// imad_explore.cu — what does the integer pipe of a B200 SM sustain for IMAD.WIDE.U32 (plain and with carry)?
// Sweeps warps per SM and independent chains per thread.  Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

template <int CH, int MODE>
__global__ void k(uint64_t *sink, uint32_t iters, uint32_t b0) {
    uint64_t acc[CH];
    uint32_t b = b0 | 1u;
#pragma unroll
    for (int c = 0; c < CH; c++) acc[c] = (uint64_t)(threadIdx.x + 1) * 0x9E3779B97F4A7C15ull + c * 77 + blockIdx.x;
    for (uint32_t i = 0; i < iters; i++) {
#pragma unroll
        for (int r = 0; r < 16; r++) {
            if (MODE == 0) {  // independent chains: acc = lo(acc)*b + acc   (IMAD.WIDE.U32)
#pragma unroll
                for (int c = 0; c < CH; c++) {
                    uint32_t lo = (uint32_t)acc[c];
                    asm volatile("mad.wide.u32 %0, %1, %2, %0;" : "+l"(acc[c]) : "r"(lo), "r"(b));
                }
            } else {  // carry chains across the CH accumulators (IMAD.WIDE.U32.X): one add.cc + (CH-1) addc.cc
                uint64_t p[CH];
#pragma unroll
                for (int c = 0; c < CH; c++) asm("mul.wide.u32 %0, %1, %2;" : "=l"(p[c]) : "r"((uint32_t)acc[c]), "r"(b));
                asm volatile("add.cc.u64 %0, %0, %1;" : "+l"(acc[0]) : "l"(p[0]));
#pragma unroll
                for (int c = 1; c < CH; c++) asm volatile("addc.cc.u64 %0, %0, %1;" : "+l"(acc[c]) : "l"(p[c]));
            }
        }
    }
    uint64_t s = 0;
#pragma unroll
    for (int c = 0; c < CH; c++) s ^= acc[c];
    sink[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int CH, int MODE>
void run(int sms, int warps_per_sm, uint64_t *sink) {
    // one CTA per SM with warps_per_sm warps (≤ 32), or several CTAs per SM for more
    int ctas_per_sm = (warps_per_sm + 31) / 32, threads = warps_per_sm / ctas_per_sm * 32;
    uint32_t iters = 2048;
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    float best = 1e30f;
    for (int rep = 0; rep < 4; rep++) {
        cudaEventRecord(e0);
        k<CH, MODE><<<sms * ctas_per_sm, threads>>>(sink, iters, 12345u + rep);
        cudaEventRecord(e1);
        cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        if (rep && ms < best) best = ms;
    }
    double ops = (double)sms * ctas_per_sm * threads * iters * 16.0 * CH;
    printf("mode=%s chains=%d warps/SM=%2d : %.3f T lane-op/s  = %.2f per clk per SM @1965MHz (%.3f ms)\n",
           MODE ? "carry(.X)" : "plain", CH, warps_per_sm, ops / (best * 1e-3) / 1e12, ops / (best * 1e-3) / sms / 1.965e9, best);
}

int main() {
    cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
    int sms = p.multiProcessorCount;
    uint64_t *sink; cudaMalloc(&sink, (size_t)sms * 4 * 1024 * 8);
    printf("%s, %d SMs\n", p.name, sms);
    int ws[] = {4, 8, 12, 16, 32, 64};
    for (int w : ws) { run<1, 0>(sms, w, sink); run<2, 0>(sms, w, sink); run<4, 0>(sms, w, sink); run<8, 0>(sms, w, sink); }
    for (int w : ws) { run<2, 1>(sms, w, sink); run<4, 1>(sms, w, sink); run<6, 1>(sms, w, sink); run<8, 1>(sms, w, sink); }
    return 0;
}
