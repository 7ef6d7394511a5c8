// api.cu — context lifetime, device buffers, profiling accessors and the integer-pipe microbenchmark.
// The C ABI is declared in include/pb200.h; nothing here (or anywhere in csrc/) touches oracle/.
#include "common.cuh"
#include "field.cuh"

int ntt_module_init(pb200_ctx *ctx);   // ntt.cu
void ntt_free_plans(pb200_ctx *ctx);   // ntt.cu
int msm_module_init(pb200_ctx *ctx);   // msm.cu

int pb_ensure(pb200_ctx *ctx, void **buf, size_t *have, size_t need) {
    if (*have >= need) return 0;
    if (*buf) {
        PB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
        PB_CUDA(ctx, cudaFree(*buf));
        *buf = nullptr;
        *have = 0;
    }
    PB_CUDA(ctx, cudaMalloc(buf, need));
    *have = need;
    return 0;
}

extern "C" int pb200_init(pb200_ctx **out, int device_id) {
    if (!out) return PB200_ERR_ARG;
    *out = nullptr;
    int count = 0;
    if (cudaGetDeviceCount(&count) != cudaSuccess || count <= 0 || device_id < 0 || device_id >= count)
        return PB200_ERR_NO_DEVICE;  // no CPU fallback: without a GPU the library refuses to work
    cudaDeviceProp prop;
    if (cudaGetDeviceProperties(&prop, device_id) != cudaSuccess) return PB200_ERR_NO_DEVICE;
    if (prop.major < 10) return PB200_ERR_NO_DEVICE;  // kernels are built for sm_100a only
    pb200_ctx *ctx = new pb200_ctx();
    ctx->device = device_id;
    ctx->sm_count = prop.multiProcessorCount;
    *out = ctx;
    PB_CUDA(ctx, cudaSetDevice(device_id));
    PB_CUDA(ctx, cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking));
    {   // keep freed stream-ordered allocations in the pool instead of returning them to the driver
        cudaMemPool_t pool;
        if (cudaDeviceGetDefaultMemPool(&pool, device_id) == cudaSuccess) {
            uint64_t keep = ~0ull;
            cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep);
        }
    }
    ctx->pinned_bytes = 1 << 16;
    PB_CUDA(ctx, cudaMallocHost(&ctx->pinned, ctx->pinned_bytes));
    PB_TRY(ntt_module_init(ctx));
    PB_TRY(msm_module_init(ctx));
    return 0;
}
extern "C" void pb200_destroy(pb200_ctx *ctx) {
    if (!ctx) return;
    cudaSetDevice(ctx->device);
    if (ctx->stream) cudaStreamSynchronize(ctx->stream);
    pb200_comm_destroy(ctx);
    ntt_free_plans(ctx);
    if (ctx->msm_ws) cudaFree(ctx->msm_ws);
    if (ctx->stage) cudaFree(ctx->stage);
    if (ctx->pinned) cudaFreeHost(ctx->pinned);
    if (ctx->stream) cudaStreamDestroy(ctx->stream);
    delete ctx;
}
extern "C" const char *pb200_last_error(const pb200_ctx *ctx) { return ctx ? ctx->err.c_str() : "null context"; }
extern "C" void *pb200_stream(pb200_ctx *ctx) { return ctx ? (void *)ctx->stream : nullptr; }
extern "C" int pb200_sync(pb200_ctx *ctx) {
    if (!ctx) return PB200_ERR_ARG;
    PB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return 0;
}
extern "C" int pb200_malloc(pb200_ctx *ctx, void **dev_ptr, size_t bytes) {
    if (!ctx) return PB200_ERR_ARG;
    PB_ARG(ctx, dev_ptr != nullptr);
    PB_CUDA(ctx, cudaSetDevice(ctx->device));
    PB_CUDA(ctx, cudaMalloc(dev_ptr, bytes ? bytes : 1));
    return 0;
}
extern "C" int pb200_free(pb200_ctx *ctx, void *dev_ptr) {
    if (!ctx) return PB200_ERR_ARG;
    PB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    PB_CUDA(ctx, cudaFree(dev_ptr));
    return 0;
}
extern "C" int pb200_h2d(pb200_ctx *ctx, void *dev_dst, const void *host_src, size_t bytes) {
    if (!ctx) return PB200_ERR_ARG;
    PB_CUDA(ctx, cudaMemcpyAsync(dev_dst, host_src, bytes, cudaMemcpyHostToDevice, ctx->stream));
    PB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return 0;
}
extern "C" int pb200_d2h(pb200_ctx *ctx, void *host_dst, const void *dev_src, size_t bytes) {
    if (!ctx) return PB200_ERR_ARG;
    PB_CUDA(ctx, cudaMemcpyAsync(host_dst, dev_src, bytes, cudaMemcpyDeviceToHost, ctx->stream));
    PB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return 0;
}
extern "C" int pb200_profile_enable(pb200_ctx *ctx, int on) {
    if (!ctx) return PB200_ERR_ARG;
    ctx->profile = on != 0;
    return 0;
}
extern "C" int pb200_profile_ms(pb200_ctx *ctx, const char *name, float *ms) {
    if (!ctx) return PB200_ERR_ARG;
    PB_ARG(ctx, name && ms);
    auto it = ctx->prof_ms.find(name);
    PB_ARG(ctx, it != ctx->prof_ms.end());
    *ms = it->second;
    return 0;
}
extern "C" int pb200_profile_reset(pb200_ctx *ctx) {
    if (!ctx) return PB200_ERR_ARG;
    ctx->prof_sum.clear();
    ctx->prof_cnt.clear();
    return 0;
}
extern "C" int pb200_profile_sum_ms(pb200_ctx *ctx, const char *name, float *ms, uint32_t *count) {
    if (!ctx) return PB200_ERR_ARG;
    PB_ARG(ctx, name && ms);
    auto it = ctx->prof_sum.find(name);
    *ms = it != ctx->prof_sum.end() ? it->second : 0.0f;
    if (count) *count = it != ctx->prof_sum.end() ? ctx->prof_cnt[name] : 0;
    return 0;
}
extern "C" uint64_t pb200_launch_count(const pb200_ctx *ctx) { return ctx ? ctx->launches : 0; }

// ---------------------------------------------------------------------------------------------------
// Integer-pipe microbenchmark.  Each thread runs carry chains of mul.wide.u32 + add.cc.u64/addc.cc.u64
// (→ IMAD.WIDE.U32.X, the instruction both hot kernels are made of) over 8 accumulators whose low words
// feed the next multiplicand, so ptxas cannot hoist or strength-reduce anything (an earlier version with a
// loop-invariant product was folded away and reported a meaningless figure).  32 warps per SM.
// Measured on B200: 32 lane-ops/clk/SM — IMAD.WIDE is half-rate, i.e. 64 lo+hi "32-bit IMAD" ops/clk/SM,
// the planning figure of SURVEY.md §8d.  scripts/imad_explore.cu sweeps warps/SM and chain counts.
namespace {
__global__ void __launch_bounds__(256) imad_wide_kernel(uint64_t *sink, uint32_t iters, uint32_t b0) {
    uint64_t acc[8];
    const uint32_t b = b0 | 1u;
#pragma unroll
    for (int k = 0; k < 8; k++) acc[k] = (uint64_t)(threadIdx.x + 1) * 0x9E3779B97F4A7C15ull + k * 77 + blockIdx.x;
    for (uint32_t i = 0; i < iters; i++) {
#pragma unroll
        for (int r = 0; r < 16; r++) {
            uint64_t p[8];
#pragma unroll
            for (int k = 0; k < 8; k++) p[k] = cc::mul_wide((uint32_t)acc[k], b);
            acc[0] = cc::add_cc64(acc[0], p[0]);
#pragma unroll
            for (int k = 1; k < 8; k++) acc[k] = cc::addc_cc64(acc[k], p[k]);
        }
    }
    uint64_t s = 0;
#pragma unroll
    for (int k = 0; k < 8; k++) s ^= acc[k];
    sink[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
}  // namespace

extern "C" int pb200_imad_peak(pb200_ctx *ctx, double *wide_lane_ops_per_s, double *sm_clock_mhz_est) {
    if (!ctx) return PB200_ERR_ARG;
    PB_ARG(ctx, wide_lane_ops_per_s != nullptr);
    PB_CUDA(ctx, cudaSetDevice(ctx->device));
    const int blocks = ctx->sm_count * 4, threads = 256;
    const uint32_t iters = 2048;
    uint64_t *sink = nullptr;
    PB_CUDA(ctx, cudaMalloc(&sink, (size_t)blocks * threads * 8));
    cudaEvent_t e0, e1;
    PB_CUDA(ctx, cudaEventCreate(&e0));
    PB_CUDA(ctx, cudaEventCreate(&e1));
    float best = 1e30f;
    for (int rep = 0; rep < 5; rep++) {
        PB_CUDA(ctx, cudaEventRecord(e0, ctx->stream));
        imad_wide_kernel<<<blocks, threads, 0, ctx->stream>>>(sink, iters, 12345u + rep);
        PB_LAUNCHED(ctx);
        PB_CUDA(ctx, cudaEventRecord(e1, ctx->stream));
        PB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
        float ms = 0;
        PB_CUDA(ctx, cudaEventElapsedTime(&ms, e0, e1));
        if (rep > 0 && ms < best) best = ms;
    }
    const double ops = (double)blocks * threads * (double)iters * 16.0 * 8.0;
    *wide_lane_ops_per_s = ops / (best * 1e-3);
    // lane-ops per SM per second ÷ 32 lanes/clk (the measured issue rate) = the SM clock the run sustained
    if (sm_clock_mhz_est) *sm_clock_mhz_est = *wide_lane_ops_per_s / ctx->sm_count / 32.0 / 1e6;
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    cudaFree(sink);
    return 0;
}
