// ntt.cu — radix-2 NTT / iNTT / coset-NTT over the BLS12-381 scalar field for sm_100a.
//
// Replaces dusk-plonk 0.8.2 `fft::EvaluationDomain::{fft, ifft, coset_fft, coset_ifft}` (crate pinned
// at /root/reference/Cargo.toml:19; algorithm restated in SURVEY.md App. B.2 and oracle/oracle.c).
// Natural order in, natural order out, bit-exact (outputs are fully reduced field elements).
//
// Structure: a four-step decomposition n = n1·n2(·n3) with at most three passes
// over HBM.  Every pass stages a tile of 2^S points × C columns in shared memory (limb-planar,
// skewed against bank conflicts), runs S decimation-in-frequency stages as radix-8 register
// butterflies with one shared-memory exchange per three stages, and writes the tile back with the
// bit-reversal, the inter-pass twiddle ω_n^{j·k}, the coset powers 7^{±j} and the n⁻¹ scaling fused
// into the load / store.  Tiles are contiguous along the column axis, so global accesses are
// C·32-byte runs (one 32-byte sector per element at worst).
#include <cuda.h>

#include <algorithm>
#include <cstring>
#include "common.cuh"
#include "field.cuh"

namespace {

constexpr uint32_t kTileLogMax = 11;  // 2^11 elements × 32 B = 64 KiB of shared memory per CTA
constexpr uint32_t kTwoAdicity = 32;

struct NttPass {
    uint32_t S;          // log2 of the sub-transform length
    uint32_t logC;       // log2 of columns per CTA
    uint32_t type;       // 0: strided sub-transform, in place (column axis contiguous)
                         // 1: contiguous sub-transform, output transposed to natural order
    uint32_t ncol_log;   // type 0: log2(columns per row block)
    uint32_t nrows_log;  // type 1: log2(number of rows) = log n − S
    uint32_t n1_log;     // type 1: log2 of the first-pass length (row ↔ natural-index digit swap)
    uint32_t load_mode;  // 0 none | 1 × lo[e]·hi[e], e = natural input index | 2 × full[e]           (coset_fft)
                         // 3 × pow((col_offset + col)·x << mult_log)  (type 0: inverse four-step twiddle before the columns)
    uint32_t store_mode; // 0 none | 1 × constant | 2 × pow(col·k << mult_log) | 3 × pow(natural output index)
                         // 4 × full[(k << ncol_log) + col] (inter-pass twiddle, one lookup) | 5 × full[natural output index]
    uint32_t mult_log;
    uint32_t col_offset; // global index of local column 0 (sharded four-step: this rank owns a column range)
    uint32_t batch_log;  // blockIdx.y selects one of several independent vectors of 2^batch_log scalars each
    uint32_t l_B, s_B;   // split point of the two-level power tables
    const Fr *tw;        // ω_{2^S}^j, j < 2^(S−1)
    const Fr *l_lo, *l_hi, *s_lo, *s_hi;
    const Fr *s_const;
    const Fr *l_full, *s_full;  // single-lookup tables (one multiply instead of two)
    // type 0 only — fused exchange of the sharded four-step: when peer_log_rl != 0xffffffff the store goes straight
    // into the destination rank's row-layout buffer (peer memory over NVLink, CUDA IPC): row k belongs to rank
    // k >> peer_log_rl and lands at [(k mod rl)·m + col_offset + col], i.e. all-to-all and transpose are the store.
    uint32_t peer_log_rl, peer_log_m;
    Fr *peer[8];
};

__device__ __forceinline__ Fr g_load(const Fr *p) {
    const uint4 *q = reinterpret_cast<const uint4 *>(p);
    uint4 a = q[0], b = q[1];
    Fr r;
    r.l[0] = a.x; r.l[1] = a.y; r.l[2] = a.z; r.l[3] = a.w;
    r.l[4] = b.x; r.l[5] = b.y; r.l[6] = b.z; r.l[7] = b.w;
    return r;
}
__device__ __forceinline__ void g_store(Fr *p, const Fr &v) {
    uint4 *q = reinterpret_cast<uint4 *>(p);
    q[0] = make_uint4(v.l[0], v.l[1], v.l[2], v.l[3]);
    q[1] = make_uint4(v.l[4], v.l[5], v.l[6], v.l[7]);
}
// Shared-memory tile: 8 limb planes of PS words; element i lives at word i ^ ((i >> 3) & 31) of each plane.
// In a round on bits [b_lo, b_lo+2] the 32 lanes of a warp cover the low L0 = logC + b_lo index bits and, when
// L0 < 5, the bits from L0+3 upward; bank bit k = i_k ^ i_(k+3) is a bijection of the lane bits for every L0
// (checked case by case in DESIGN.md §4.2), so every access pattern of the kernel is bank-conflict free.
__device__ __forceinline__ uint32_t sm_slot(uint32_t i) { return i ^ ((i >> 3) & 31u); }
__device__ __forceinline__ Fr sm_load(const uint32_t *sm, uint32_t PS, uint32_t i) {
    uint32_t s = sm_slot(i);
    Fr r;
#pragma unroll
    for (int w = 0; w < 8; w++) r.l[w] = sm[w * PS + s];
    return r;
}
__device__ __forceinline__ void sm_store(uint32_t *sm, uint32_t PS, uint32_t i, const Fr &v) {
    uint32_t s = sm_slot(i);
#pragma unroll
    for (int w = 0; w < 8; w++) sm[w * PS + s] = v.l[w];
}
__device__ __forceinline__ Fr pow2level(const Fr *lo, const Fr *hi, uint32_t B, uint32_t e) {
    Fr a = g_load(lo + (e & ((1u << B) - 1)));
    Fr b = g_load(hi + (e >> B));
    return a * b;
}
// DIF butterfly: (a, b) ← (a + b, (a − b)·w)
__device__ __forceinline__ void bfly(Fr &a, Fr &b, const Fr &w) {
    Fr s = a + b;
    Fr d = a - b;
    a = s;
    b = d * w;
}
__device__ __forceinline__ void bfly1(Fr &a, Fr &b) {  // w = 1
    Fr s = a + b;
    b = a - b;
    a = s;
}

#include "ntt_tma.cuh"

// Three DIF stages on bits [b_lo, b_lo+2] (b_lo ≥ 1) of the eight points a thread holds; v = x mod 2^b_lo.
__device__ __forceinline__ void radix8_general(Fr (&a)[8], const Fr *tw, uint32_t v, uint32_t b_lo, uint32_t S) {
    {
        const uint32_t sh = S - 3 - b_lo;
#pragma unroll
        for (int e = 0; e < 4; e++) {
            Fr w = g_load(tw + ((((uint32_t)e << b_lo) | v) << sh));
            bfly(a[e], a[e + 4], w);
        }
    }
    {
        const uint32_t sh = S - 2 - b_lo;
#pragma unroll
        for (int e = 0; e < 2; e++) {
            Fr w = g_load(tw + ((((uint32_t)e << b_lo) | v) << sh));
            bfly(a[e], a[e + 2], w);
            bfly(a[e + 4], a[e + 6], w);
        }
    }
    {
        Fr w = g_load(tw + (v << (S - 1 - b_lo)));
#pragma unroll
        for (int e = 0; e < 8; e += 2) bfly(a[e], a[e + 1], w);
    }
}
// The last round, on bits [0, 2]: only stages b_top … 0 remain and their twiddles are the constants ω₈^k.
__device__ __forceinline__ void radix8_low(Fr (&a)[8], const Fr *tw, int b_top, uint32_t S) {
    if (b_top >= 2) {
        bfly1(a[0], a[4]);
#pragma unroll
        for (int e = 1; e < 4; e++) {
            Fr w = g_load(tw + ((uint32_t)e << (S - 3)));
            bfly(a[e], a[e + 4], w);
        }
    }
    if (b_top >= 1) {
        Fr w4 = g_load(tw + (1u << (S - 2)));
        bfly1(a[0], a[2]);
        bfly1(a[4], a[6]);
        bfly(a[1], a[3], w4);
        bfly(a[5], a[7], w4);
    }
#pragma unroll
    for (int e = 0; e < 8; e += 2) bfly1(a[e], a[e + 1]);
}

// One pass: stage the tile global → shared (coalesced, optional input scaling), run every round through shared
// memory, then shared → global with the bit reversal, twiddle and scalings.  (A variant that fed the first round
// straight from global memory and stored from the last round's registers measured 8-25 % slower on B200 — the extra
// live registers cost more than the two saved exchanges — and was dropped.)
__global__ void __launch_bounds__(256, 2) ntt_pass_kernel(const Fr *__restrict__ in, Fr *__restrict__ out, NttPass p) {
    extern __shared__ uint32_t sm[];
    in += (uint64_t)blockIdx.y << p.batch_log;
    out += (uint64_t)blockIdx.y << p.batch_log;
    const uint32_t S = p.S, logC = p.logC, C = 1u << logC;
    const uint32_t T = 1u << (S + logC), PS = max(T, 32u);
    const uint32_t tid = threadIdx.x, nthr = blockDim.x;
    uint64_t in_base = 0;
    uint32_t rowrev0 = 0, col0 = 0;
    if (p.type == 0) {
        uint32_t blocks_per_row_log = p.ncol_log - logC;
        uint32_t R = blockIdx.x >> blocks_per_row_log;
        col0 = (blockIdx.x & ((1u << blocks_per_row_log) - 1)) << logC;
        in_base = ((uint64_t)R << (S + p.ncol_log)) + col0;
    } else {
        rowrev0 = blockIdx.x << logC;
    }
    const uint32_t n2_log = p.nrows_log - p.n1_log;
    for (uint32_t i = tid; i < T; i += nthr) {
        uint32_t x, c;
        uint64_t addr;
        if (p.type == 0) {
            c = i & (C - 1);
            x = i >> logC;
            addr = in_base + ((uint64_t)x << p.ncol_log) + c;
        } else {
            x = i & ((1u << S) - 1);
            c = i >> S;
            uint32_t rr = rowrev0 + c;
            uint32_t row = ((rr & ((1u << p.n1_log) - 1)) << n2_log) + (rr >> p.n1_log);
            addr = ((uint64_t)row << S) + x;
        }
        Fr v = g_load(in + addr);
        if (p.load_mode == 1) v = v * pow2level(p.l_lo, p.l_hi, p.l_B, (uint32_t)addr);
        else if (p.load_mode == 2) v = v * g_load(p.l_full + addr);
        else if (p.load_mode == 3) v = v * pow2level(p.l_lo, p.l_hi, p.l_B, ((p.col_offset + col0 + c) * x) << p.mult_log);
        sm_store(sm, PS, (x << logC) + c, v);
    }
    __syncthreads();
    {
        const uint32_t c = tid & (C - 1), tx = tid >> logC;
        int b_top = (int)S - 1;
        Fr a[8];
        while (b_top >= 3) {
            const uint32_t b_lo = (uint32_t)b_top - 2;
            const uint32_t v = tx & ((1u << b_lo) - 1), u = tx >> b_lo;
            const uint32_t xbase = (u << (b_lo + 3)) | v;
#pragma unroll
            for (int e = 0; e < 8; e++) a[e] = sm_load(sm, PS, ((xbase | ((uint32_t)e << b_lo)) << logC) + c);
            radix8_general(a, p.tw, v, b_lo, S);
#pragma unroll
            for (int e = 0; e < 8; e++) sm_store(sm, PS, ((xbase | ((uint32_t)e << b_lo)) << logC) + c, a[e]);
            __syncthreads();
            b_top -= 3;
        }
#pragma unroll
        for (int e = 0; e < 8; e++) a[e] = sm_load(sm, PS, (((tx << 3) | (uint32_t)e) << logC) + c);
        radix8_low(a, p.tw, b_top, S);
#pragma unroll
        for (int e = 0; e < 8; e++) sm_store(sm, PS, (((tx << 3) | (uint32_t)e) << logC) + c, a[e]);
        __syncthreads();
    }
    for (uint32_t i = tid; i < T; i += nthr) {
        const uint32_t c = i & (C - 1), x = i >> logC;
        const uint32_t k = __brev(x) >> (32 - S);
        Fr v = sm_load(sm, PS, i);
        uint64_t addr;
        if (p.type == 0) {
            addr = in_base + ((uint64_t)k << p.ncol_log) + c;
            if (p.store_mode == 2) v = v * pow2level(p.s_lo, p.s_hi, p.s_B, ((p.col_offset + col0 + c) * k) << p.mult_log);
            else if (p.store_mode == 4) v = v * g_load(p.s_full + (((uint64_t)k << p.ncol_log) + col0 + c));
            if (p.peer_log_rl != 0xffffffffu) {
                Fr *dst = p.peer[k >> p.peer_log_rl];
                g_store(dst + (((uint64_t)(k & ((1u << p.peer_log_rl) - 1)) << p.peer_log_m) + p.col_offset + col0 + c), v);
                continue;
            }
        } else {
            addr = (uint64_t)(rowrev0 + c) + ((uint64_t)k << p.nrows_log);
            if (p.store_mode == 3) v = v * pow2level(p.s_lo, p.s_hi, p.s_B, (uint32_t)addr);
            else if (p.store_mode == 5) v = v * g_load(p.s_full + addr);
        }
        if (p.store_mode == 1) v = v * g_load(p.s_const);
        g_store(out + addr, v);
    }
}

// Domains of 1, 2 or 4 points: one thread evaluates the definition directly.
__global__ void ntt_tiny_kernel(Fr *data, uint32_t log_n, int inverse, int coset, const Fr *consts) {
    const uint32_t n = 1u << log_n;
    Fr w = g_load(consts + 0), ninv = g_load(consts + 1), g = g_load(consts + 2);
    Fr a[4], o[4];
    for (uint32_t j = 0; j < n; j++) a[j] = g_load(data + j);
    if (coset && !inverse) {
        Fr pw = Fr::one();
        for (uint32_t j = 0; j < n; j++) { a[j] = a[j] * pw; pw = pw * g; }
    }
    for (uint32_t i = 0; i < n; i++) {
        Fr acc = Fr::zero(), wi = w.pow_u64(i), pw = Fr::one();
        for (uint32_t j = 0; j < n; j++) { acc = acc + a[j] * pw; pw = pw * wi; }
        o[i] = acc;
    }
    if (inverse) {
        Fr pw = ninv;
        for (uint32_t i = 0; i < n; i++) { o[i] = o[i] * pw; if (coset) pw = pw * g; }
    }
    for (uint32_t i = 0; i < n; i++) g_store(data + i, o[i]);
}

// consts[0] = ω_n (or ω_n⁻¹ when inverse), [1] = n⁻¹, [2] = 7 (or 7⁻¹ when inverse), [3] = 1
__global__ void ntt_consts_kernel(Fr *consts, uint32_t log_n, int inverse) {
    Fr seven = Fr::zero();
    seven.l[0] = 7;
    seven = seven.to_mont();
    // (r − 1) >> 32
    const uint32_t e[7] = {0xffffffffu, 0xfffe5bfeu, 0x53bda402u, 0x09a1d805u, 0x3339d808u, 0x299d7d48u, 0x73eda753u};
    Fr root = seven.pow(e, 7);  // primitive 2^32-th root of unity
    Fr w = root;
    for (uint32_t i = log_n; i < kTwoAdicity; i++) w = w.sqr();
    Fr two = Fr::one().dbl();
    Fr ninv = two.inv().pow_u64(log_n);
    Fr g = seven;
    if (inverse) { w = w.inv(); g = g.inv(); }
    g_store(consts + 0, w);
    g_store(consts + 1, ninv);
    g_store(consts + 2, g);
    g_store(consts + 3, Fr::one());
}
// out[j] = pre · base^(j·stride)
__global__ void fill_powers_kernel(Fr *out, uint32_t count, const Fr *base, uint64_t stride, const Fr *pre) {
    uint32_t j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= count) return;
    Fr b = g_load(base), p = g_load(pre);
    g_store(out + j, p * b.pow_u64((uint64_t)j * stride));
}

// full[(k << ncol_log) + j] = lo/hi-composed base^((j·k) << mult_log), j < 2^ncol_log, k < 2^S
__global__ void fill_boundary_kernel(Fr *out, uint32_t S, uint32_t ncol_log, uint32_t mult_log, const Fr *lo, const Fr *hi,
                                     uint32_t B) {
    const uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
    if (i >= (1ull << (S + ncol_log))) return;
    const uint32_t j = (uint32_t)(i & ((1ull << ncol_log) - 1)), k = (uint32_t)(i >> ncol_log);
    g_store(out + i, pow2level(lo, hi, B, (j * k) << mult_log));
}
// full[e] = lo[e & mask]·hi[e >> B], e < 2^L
__global__ void fill_linear_kernel(Fr *out, uint32_t L, const Fr *lo, const Fr *hi, uint32_t B) {
    const uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
    if (i >= (1ull << L)) return;
    g_store(out + i, pow2level(lo, hi, B, (uint32_t)i));
}

// out[(r·B + b)·C + c] = in[(b·R + r)·C + c]: regroups the G received blocks of an all-to-all ([B][R][C] → [R][B][C]).
__global__ void block_transpose_kernel(Fr *__restrict__ out, const Fr *__restrict__ in, uint32_t B, uint32_t R, uint32_t C) {
    const uint64_t total = (uint64_t)B * R * C;
    for (uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; i < total; i += (uint64_t)gridDim.x * blockDim.x) {
        const uint32_t c = (uint32_t)(i % C);
        const uint64_t br = i / C;
        const uint32_t r = (uint32_t)(br % R), b = (uint32_t)(br / R);
        g_store(out + ((uint64_t)r * B + b) * C + c, g_load(in + i));
    }
}

}  // namespace

struct NttPlan {
    uint32_t log_n = 0;
    int inverse = 0, coset = 0;
    int n_pass = 0;
    NttPass pass[3];
    bool tma = false;        // every pass also has a TMA description (ntt_pass_tma_kernel); same tables, same tile shape
    NttPassTma tpass[3];
    Fr *consts = nullptr;
    std::vector<void *> allocs;
    size_t table_bytes = 0;
};

static uint32_t env_u32(const char *name, uint32_t dflt) {
    const char *v = getenv(name);
    return v ? (uint32_t)atoi(v) : dflt;
}

static int fill_powers(pb200_ctx *ctx, NttPlan *pl, Fr **out, uint32_t count, const Fr *base, uint64_t stride, const Fr *pre) {
    void *buf = nullptr;
    PB_CUDA(ctx, cudaMalloc(&buf, (size_t)count * sizeof(Fr)));
    pl->allocs.push_back(buf);
    fill_powers_kernel<<<(count + 127) / 128, 128, 0, ctx->stream>>>((Fr *)buf, count, base, stride, pre);
    PB_LAUNCHED(ctx);
    *out = (Fr *)buf;
    return 0;
}

static int ntt_build_plan(pb200_ctx *ctx, uint32_t L, int inverse, int coset, NttPlan **out) {
    NttPlan *pl = new NttPlan();
    pl->log_n = L;
    pl->inverse = inverse;
    pl->coset = coset;
    *out = pl;  // the caller caches it once it is complete (ntt_plan_discard on any failure)
    void *cbuf = nullptr;
    PB_CUDA(ctx, cudaMalloc(&cbuf, 4 * sizeof(Fr)));
    pl->allocs.push_back(cbuf);
    pl->consts = (Fr *)cbuf;
    ntt_consts_kernel<<<1, 1, 0, ctx->stream>>>(pl->consts, L, inverse);
    PB_LAUNCHED(ctx);
    const Fr *c_w = pl->consts + 0, *c_ninv = pl->consts + 1, *c_g = pl->consts + 2, *c_one = pl->consts + 3;
    if (L < 3) { pl->n_pass = 0; return 0; }

    // Plan shape.  "tma": tiles of 2^S points × 4 columns with S ≤ 9 (two passes up to 2^18, three up to 2^27), the shape
    // ntt_pass_tma_kernel is built for — 128-byte runs in global memory for every pass.  Otherwise (tiny and huge domains,
    // PB200_NTT_PLAN=legacy): as few passes as an 2^11-point tile allows, columns per tile from the tile size.
    const char *plan_env = getenv("PB200_NTT_PLAN");
    const bool tma_shape = !(plan_env && !strcmp(plan_env, "legacy")) && L >= env_u32("PB200_NTT_TMA_MIN_LOG", 12) && L <= 27;
    const int P = tma_shape ? (L <= 18 ? 2 : 3) : (L <= kTileLogMax ? 1 : (L <= 2 * kTileLogMax ? 2 : 3));
    uint32_t S[3] = {0, 0, 0};
    if (P == 1) S[0] = L;
    else if (P == 2) { S[0] = (L + 1) / 2; S[1] = L - S[0]; }
    else { S[0] = (L + 2) / 3; S[1] = (L - S[0] + 1) / 2; S[2] = L - S[0] - S[1]; }
    pl->n_pass = P;
    const uint32_t B = (L + 1) / 2;  // two-level tables: lo has 2^B entries, hi 2^(L−B)

    Fr *lo = nullptr, *hi_plain = nullptr, *hi_scaled = nullptr, *clo = nullptr, *chi = nullptr;
    if (P >= 2) {
        PB_TRY(fill_powers(ctx, pl, &lo, 1u << B, c_w, 1, c_one));
        PB_TRY(fill_powers(ctx, pl, &hi_plain, 1u << (L - B), c_w, 1ull << B, c_one));
        hi_scaled = hi_plain;
        if (inverse && !coset) PB_TRY(fill_powers(ctx, pl, &hi_scaled, 1u << (L - B), c_w, 1ull << B, c_ninv));
    }
    if (coset) {
        PB_TRY(fill_powers(ctx, pl, &clo, 1u << B, c_g, 1, c_one));
        PB_TRY(fill_powers(ctx, pl, &chi, 1u << (L - B), c_g, 1ull << B, inverse ? c_ninv : c_one));
    }
    // Single-lookup tables replace the two-level composition (one multiply per element instead of two) while
    // the context's table budget allows: a table is as long as the vector (32·n bytes).
    const uint32_t full_max_log = env_u32("PB200_NTT_FULL_TABLE_MAX_LOG", 26);
    const uint32_t tile_pref = std::min(kTileLogMax, std::max(3u, env_u32("PB200_NTT_TILE_LOG", 10)));
    // CTA tile: the preferred size, but small transforms get smaller tiles so that ≥ 512 CTAs exist
    auto tile_log_for = [&](uint32_t s_log) { return std::max(s_log, std::min(tile_pref, L >= 9 ? L - 9 : 0u)); };
    size_t budget_used = 0;
    for (auto &kv : ctx->ntt_plans) budget_used += kv.second->table_bytes;
    auto full_ok = [&](uint32_t log_entries) {
        return L <= full_max_log && budget_used + pl->table_bytes + ((size_t)32 << log_entries) <= ((size_t)16 << 30);
    };
    auto alloc_table = [&](uint32_t log_entries, Fr **out) -> int {
        void *buf = nullptr;
        PB_CUDA(ctx, cudaMalloc(&buf, (size_t)32 << log_entries));
        pl->allocs.push_back(buf);
        pl->table_bytes += (size_t)32 << log_entries;
        *out = (Fr *)buf;
        return 0;
    };
    uint32_t done = 0;  // bits already transformed by earlier passes
    for (int i = 0; i < P; i++) {
        NttPass &p = pl->pass[i];
        memset(&p, 0, sizeof(p));
        p.peer_log_rl = 0xffffffffu;
        p.S = S[i];
        Fr *tw = nullptr;
        PB_TRY(fill_powers(ctx, pl, &tw, 1u << (S[i] - 1), c_w, 1ull << (L - S[i]), c_one));
        p.tw = tw;
        const bool last = (i == P - 1);
        if (!last) {
            p.type = 0;
            p.ncol_log = L - done - S[i];
            p.logC = tma_shape ? 2 : std::min(tile_log_for(S[i]) - S[i], p.ncol_log);
            p.store_mode = 2;
            p.mult_log = done;  // ω_m^{j·k} with m = n / 2^done equals ω_n^{(j·k) << done}
            p.s_lo = lo;
            p.s_hi = (i == 0) ? hi_scaled : hi_plain;
            p.s_B = B;
            if (full_ok(S[i] + p.ncol_log)) {
                Fr *full = nullptr;
                PB_TRY(alloc_table(S[i] + p.ncol_log, &full));
                const uint64_t cnt = 1ull << (S[i] + p.ncol_log);
                fill_boundary_kernel<<<(uint32_t)((cnt + 255) / 256), 256, 0, ctx->stream>>>(full, S[i], p.ncol_log, p.mult_log,
                                                                                           p.s_lo, p.s_hi, B);
                PB_LAUNCHED(ctx);
                p.store_mode = 4;
                p.s_full = full;
            }
        } else {
            p.type = 1;
            p.nrows_log = L - S[i];
            p.n1_log = (P == 3) ? S[0] : p.nrows_log;
            p.logC = tma_shape ? 2 : std::min(tile_log_for(S[i]) - S[i], p.nrows_log);
            if (inverse && coset) {
                p.store_mode = 3; p.s_lo = clo; p.s_hi = chi; p.s_B = B;
                if (full_ok(L)) {
                    Fr *full = nullptr;
                    PB_TRY(alloc_table(L, &full));
                    fill_linear_kernel<<<(uint32_t)(((1ull << L) + 255) / 256), 256, 0, ctx->stream>>>(full, L, clo, chi, B);
                    PB_LAUNCHED(ctx);
                    p.store_mode = 5;
                    p.s_full = full;
                }
            }
            else if (inverse && P == 1) { p.store_mode = 1; p.s_const = c_ninv; }
        }
        if (i == 0 && coset && !inverse) {
            p.load_mode = 1; p.l_lo = clo; p.l_hi = chi; p.l_B = B;
            if (full_ok(L)) {
                Fr *full = nullptr;
                PB_TRY(alloc_table(L, &full));
                fill_linear_kernel<<<(uint32_t)(((1ull << L) + 255) / 256), 256, 0, ctx->stream>>>(full, L, clo, chi, B);
                PB_LAUNCHED(ctx);
                p.load_mode = 2;
                p.l_full = full;
            }
        }
        done += S[i];
    }
    // TMA description of the same passes: needs the single-lookup tables (one multiply per element) throughout
    if (tma_shape && pb_encode_tiled() != nullptr) {
        bool ok = true;
        for (int i = 0; i < P && ok; i++) {
            const NttPass &p = pl->pass[i];
            ok = p.S >= 6 && p.S <= 9 && p.logC == 2 && (p.load_mode == 0 || p.load_mode == 2) &&
                 (p.type == 0 ? p.store_mode == 4 : (p.store_mode == 0 || p.store_mode == 5));
        }
        for (int i = 0; i < P && ok; i++) {
            const NttPass &p = pl->pass[i];
            NttPassTma &t = pl->tpass[i];
            memset(&t, 0, sizeof(t));
            t.S = p.S; t.type = p.type; t.ncol_log = p.ncol_log; t.nrows_log = p.nrows_log; t.n1_log = p.n1_log;
            t.load_mode = p.load_mode; t.store_mode = p.store_mode; t.l_full = p.l_full; t.s_full = p.s_full;
            const uint32_t entries = ntt_tw_image_entries(p.S);
            void *img = nullptr;
            PB_CUDA(ctx, cudaMalloc(&img, (size_t)entries * sizeof(Fr)));
            pl->allocs.push_back(img);
            ntt_tw_image_kernel<<<(entries + 127) / 128, 128, 0, ctx->stream>>>((Fr *)img, p.tw, p.S);
            PB_LAUNCHED(ctx);
            t.tw_img = (const Fr *)img;
            t.tw_bytes = entries * (uint32_t)sizeof(Fr);
        }
        pl->tma = ok;
    }
    return 0;
}

// A plan whose construction failed half-way (e.g. cudaMalloc of a 32·n-byte table) must not stay cached: the next
// transform with the same key would launch on null tables.
static void ntt_plan_discard(pb200_ctx *ctx, NttPlan *pl) {
    if (!pl) return;
    cudaStreamSynchronize(ctx->stream);  // fill kernels may still be writing the tables
    for (void *a : pl->allocs) cudaFree(a);
    delete pl;
    (void)cudaGetLastError();
}

void ntt_free_plans(pb200_ctx *ctx) {
    for (auto &kv : ctx->ntt_plans) {
        for (void *a : kv.second->allocs) cudaFree(a);
        delete kv.second;
    }
    ctx->ntt_plans.clear();
    if (ctx->ntt_scratch) cudaFree(ctx->ntt_scratch);
    ctx->ntt_scratch = nullptr;
    ctx->ntt_scratch_bytes = 0;
}

// ntt_pass_tma_kernel instantiations: S ∈ 6…9 × resident-thread budget (512 threads/SM = 128 registers, 384 = 168).
typedef void (*NttTmaKernel)(const CUtensorMap, const CUtensorMap, const Fr *, const NttPassTma);
static NttTmaKernel ntt_tma_kernel_for(uint32_t S, bool wide_regs) {
    switch (S) {
        case 6: return wide_regs ? ntt_pass_tma_kernel<6, 384> : ntt_pass_tma_kernel<6, 512>;
        case 7: return wide_regs ? ntt_pass_tma_kernel<7, 384> : ntt_pass_tma_kernel<7, 512>;
        case 8: return wide_regs ? ntt_pass_tma_kernel<8, 384> : ntt_pass_tma_kernel<8, 512>;
        case 9: return wide_regs ? ntt_pass_tma_kernel<9, 384> : ntt_pass_tma_kernel<9, 512>;
    }
    return nullptr;
}
static size_t ntt_tma_smem_bytes(uint32_t S) {
    return 1024 + ((size_t)32 << (S + 2)) + (((size_t)ntt_tw_image_entries(S) * sizeof(Fr) + 15) & ~(size_t)15) + 16;
}
// One pass over `nb` vectors of 2^L scalars (src / dst point at the first of them).
static int ntt_launch_tma(pb200_ctx *ctx, const NttPassTma &tp, const Fr *src, Fr *dst, uint32_t L, uint32_t blocks, uint32_t nb) {
    // measured on B200 (profiles/ntt_sweep_r02.json): 512 threads per SM at 128 registers and 384 at ≤ 168 tie up to S = 8
    // (2^24: 3.40 ms both) and the wide variant loses at S = 9, where it leaves one CTA per SM (2^26: 17.1 vs 14.4 ms)
    static const char *regs_env = getenv("PB200_NTT_TMA_REGS");
    const bool wide_regs = regs_env != nullptr && !strcmp(regs_env, "168");
    const uint32_t box_rows = std::min(1u << tp.S, 256u);
    CUtensorMap map_in, map_out;
    if (tp.type == 0) {
        const uint64_t cols = 1ull << tp.ncol_log, rows = (uint64_t)nb << (L - tp.ncol_log);
        PB_TRY(pb_make_tile_map(ctx, &map_in, src, cols, rows, box_rows));
        PB_TRY(pb_make_tile_map(ctx, &map_out, dst, cols, rows, box_rows));
    } else {
        PB_TRY(pb_make_tile_map(ctx, &map_out, dst, 1ull << tp.nrows_log, (uint64_t)nb << tp.S, box_rows));
        map_in = map_out;   // rows are staged through registers: no input map
    }
    NttTmaKernel k = ntt_tma_kernel_for(tp.S, wide_regs);
    PB_ARG(ctx, k != nullptr);
    k<<<dim3(blocks, nb), 1u << (tp.S - 1), ntt_tma_smem_bytes(tp.S), ctx->stream>>>(map_in, map_out, src, tp);
    PB_LAUNCHED(ctx);
    return 0;
}

static int ntt_run(pb200_ctx *ctx, Fr *data, uint32_t L, int inverse, int coset, uint32_t batch = 1) {
    if (L == 0) return 0;  // a one-point domain: every variant is the identity map
    NttPlan *pl = nullptr;
    auto it = ctx->ntt_plans.find(L | ((uint32_t)inverse << 8) | ((uint32_t)coset << 9));
    if (it != ctx->ntt_plans.end()) {
        pl = it->second;
    } else {
        const int rc = ntt_build_plan(ctx, L, inverse, coset, &pl);
        if (rc) {
            ntt_plan_discard(ctx, pl);
            return rc;
        }
        ctx->ntt_plans[L | ((uint32_t)inverse << 8) | ((uint32_t)coset << 9)] = pl;
    }
    PbTimer timer(ctx, "ntt.total");
    if (pl->n_pass == 0) {
        for (uint32_t b = 0; b < batch; b++) {
            ntt_tiny_kernel<<<1, 1, 0, ctx->stream>>>(data + ((size_t)b << L), L, inverse, coset, pl->consts);
            PB_LAUNCHED(ctx);
        }
    } else {
        const size_t bytes = (sizeof(Fr) << L) * batch;
        Fr *scratch = nullptr;
        if (pl->n_pass > 1) {
            PB_TRY(pb_ensure(ctx, &ctx->ntt_scratch, &ctx->ntt_scratch_bytes, bytes));
            scratch = (Fr *)ctx->ntt_scratch;
        }
        for (int i = 0; i < pl->n_pass; i++) {
            const NttPass &p = pl->pass[i];
            const uint32_t tile_log = p.S + p.logC;
            const uint32_t T = 1u << tile_log;
            const size_t smem = (size_t)8 * std::max(T, 32u) * sizeof(uint32_t);
            const uint32_t threads = T >> 3, blocks = 1u << (L - tile_log);
            // pass 0 reads the caller's vector, the last pass writes it; intermediates live in scratch
            const Fr *src = (i == 0) ? data : scratch;
            Fr *dst = (i == pl->n_pass - 1) ? data : scratch;
            NttPass pb = p;
            pb.batch_log = L;
            static const bool force_legacy = getenv("PB200_NTT_KERNEL") && !strcmp(getenv("PB200_NTT_KERNEL"), "legacy");
            for (uint32_t b0 = 0; b0 < batch; b0 += 32768) {  // gridDim.y ≤ 65535
                const uint32_t nb = std::min<uint32_t>(32768, batch - b0);
                if (pl->tma && !force_legacy) {
                    NttPassTma tp = pl->tpass[i];
                    tp.batch_log = L;
                    tp.last = (i == pl->n_pass - 1);
                    PB_TRY(ntt_launch_tma(ctx, tp, src + ((size_t)b0 << L), dst + ((size_t)b0 << L), L, blocks, nb));
                    continue;
                }
                ntt_pass_kernel<<<dim3(blocks, nb), threads, smem, ctx->stream>>>(src + ((size_t)b0 << L), dst + ((size_t)b0 << L), pb);
                PB_LAUNCHED(ctx);
            }
        }
    }
    timer.stop();
    if (ctx->profile) {
        PB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
        timer.collect();
    }
    return 0;
}

int ntt_module_init(pb200_ctx *ctx) {
    PB_CUDA(ctx, cudaFuncSetAttribute(ntt_pass_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                      8 * (1 << kTileLogMax) * 4));
    for (uint32_t S = 6; S <= 9; S++)
        for (int w = 0; w < 2; w++)
            PB_CUDA(ctx, cudaFuncSetAttribute(ntt_tma_kernel_for(S, w != 0), cudaFuncAttributeMaxDynamicSharedMemorySize,
                                              (int)ntt_tma_smem_bytes(S)));
    return 0;
}

extern "C" int pb200_domain_log_size(size_t num_coeffs, uint32_t *log_n) {
    if (!log_n) return PB200_ERR_ARG;
    uint32_t l = 0;
    while (((size_t)1 << l) < num_coeffs) l++;
    if (l >= kTwoAdicity) return PB200_ERR_ARG;  // EvaluationDomain::new → Err(InvalidEvalDomainSize)
    *log_n = l;
    return 0;
}
extern "C" int pb200_ntt_dev(pb200_ctx *ctx, uint64_t *data_dev, uint32_t log_n, int inverse, int coset) {
    if (!ctx) return PB200_ERR_ARG;
    PB_ARG(ctx, data_dev != nullptr);
    PB_ARG(ctx, log_n < kTwoAdicity);
    PB_CUDA(ctx, cudaSetDevice(ctx->device));
    return ntt_run(ctx, (Fr *)data_dev, log_n, inverse ? 1 : 0, coset ? 1 : 0);
}
extern "C" int pb200_ntt_batch_dev(pb200_ctx *ctx, uint64_t *data_dev, uint32_t log_n, uint32_t batch, int inverse, int coset) {
    if (!ctx) return PB200_ERR_ARG;
    PB_ARG(ctx, data_dev != nullptr);
    PB_ARG(ctx, log_n < kTwoAdicity && batch >= 1);
    PB_CUDA(ctx, cudaSetDevice(ctx->device));
    return ntt_run(ctx, (Fr *)data_dev, log_n, inverse ? 1 : 0, coset ? 1 : 0, batch);
}
extern "C" int pb200_ntt(pb200_ctx *ctx, uint64_t *data_host, uint32_t log_n, int inverse, int coset) {
    if (!ctx) return PB200_ERR_ARG;
    PB_ARG(ctx, data_host != nullptr);
    PB_ARG(ctx, log_n < kTwoAdicity);
    PB_CUDA(ctx, cudaSetDevice(ctx->device));
    const size_t bytes = (size_t)32 << log_n;
    PB_TRY(pb_ensure(ctx, &ctx->stage, &ctx->stage_bytes, bytes));
    PB_CUDA(ctx, cudaMemcpyAsync(ctx->stage, data_host, bytes, cudaMemcpyHostToDevice, ctx->stream));
    PB_TRY(ntt_run(ctx, (Fr *)ctx->stage, log_n, inverse ? 1 : 0, coset ? 1 : 0));
    PB_CUDA(ctx, cudaMemcpyAsync(data_host, ctx->stage, bytes, cudaMemcpyDeviceToHost, ctx->stream));
    PB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return 0;
}

// Plan of the sharded four-step's column step (one pass, type 0): consts, two-level ω_n tables, ω_{n1} butterfly table.
static int ntt_build_columns_plan(pb200_ctx *ctx, uint32_t log_n, uint32_t log_n1, int inverse, NttPlan **out) {
    NttPlan *pl = new NttPlan();
    pl->log_n = log_n;
    pl->inverse = inverse;
    *out = pl;
    void *cbuf = nullptr;
    PB_CUDA(ctx, cudaMalloc(&cbuf, 5 * sizeof(Fr)));
    pl->allocs.push_back(cbuf);
    pl->consts = (Fr *)cbuf;
    ntt_consts_kernel<<<1, 1, 0, ctx->stream>>>(pl->consts, log_n, inverse);
    PB_LAUNCHED(ctx);
    const Fr *c_w = pl->consts + 0, *c_one = pl->consts + 3;
    // n1⁻¹ = (2⁻¹)^log_n1: consts[1] holds 2^−log_n; reuse the constants kernel for a 2^log_n1 domain
    Fr *c1 = nullptr;
    void *c1buf = nullptr;
    PB_CUDA(ctx, cudaMalloc(&c1buf, 4 * sizeof(Fr)));
    pl->allocs.push_back(c1buf);
    c1 = (Fr *)c1buf;
    ntt_consts_kernel<<<1, 1, 0, ctx->stream>>>(c1, log_n1, inverse);
    PB_LAUNCHED(ctx);
    const uint32_t B = (log_n + 1) / 2;
    Fr *lo = nullptr, *hi = nullptr, *tw = nullptr;
    PB_TRY(fill_powers(ctx, pl, &lo, 1u << B, c_w, 1, c_one));
    PB_TRY(fill_powers(ctx, pl, &hi, 1u << (log_n - B), c_w, 1ull << B, c_one));
    PB_TRY(fill_powers(ctx, pl, &tw, 1u << (log_n1 - 1), c_w, 1ull << (log_n - log_n1), c_one));
    NttPass &p = pl->pass[0];
    memset(&p, 0, sizeof(p));
    p.peer_log_rl = 0xffffffffu;
    p.S = log_n1;
    p.type = 0;
    p.tw = tw;
    if (!inverse) { p.store_mode = 2; p.s_lo = lo; p.s_hi = hi; p.s_B = B; }
    else { p.load_mode = 3; p.l_lo = lo; p.l_hi = hi; p.l_B = B; p.store_mode = 1; p.s_const = c1 + 1; }
    pl->n_pass = 1;
    return 0;
}
// ---- sharded four-step building blocks (orchestrated by plonk-prototype_b200/dist_ntt.py) ----------
// Column step of a 2^log_n transform split as n = n1·m over G ranks: this rank holds the n1 × cols matrix of its
// column range [col_offset, col_offset + cols) (row-major, cols = 2^log_cols).
//   forward: length-n1 NTT down every column, then × ω_n^{(col_offset + col)·k1}          (before the all-to-all)
//   inverse: × ω_n^{−(col_offset + col)·k1}, then length-n1 iNTT down every column, × n1⁻¹  (after the all-to-all back)
// One pass, in place; log_n1 ≤ 11.
static int ntt_columns_impl(pb200_ctx *ctx, uint64_t *data_dev, uint32_t log_n, uint32_t log_n1, uint32_t log_cols, uint32_t col_offset,
                            int inverse, uint32_t world, void *const *peer_rows) {
    if (!ctx) return PB200_ERR_ARG;
    PB_ARG(ctx, data_dev != nullptr);
    PB_ARG(ctx, log_n < kTwoAdicity && log_n1 >= 3 && log_n1 <= kTileLogMax && log_n1 + log_cols <= log_n);
    PB_CUDA(ctx, cudaSetDevice(ctx->device));
    inverse = inverse ? 1 : 0;
    const uint32_t key = 0x40000000u | log_n | (log_n1 << 8) | ((uint32_t)inverse << 16);
    NttPlan *pl = nullptr;
    auto it = ctx->ntt_plans.find(key);
    if (it != ctx->ntt_plans.end()) {
        pl = it->second;
    } else {
        const int rc = ntt_build_columns_plan(ctx, log_n, log_n1, inverse, &pl);
        if (rc) {
            ntt_plan_discard(ctx, pl);
            return rc;
        }
        ctx->ntt_plans[key] = pl;
    }
    NttPass p = pl->pass[0];
    p.ncol_log = log_cols;
    p.col_offset = col_offset;
    p.mult_log = 0;  // the twiddle is ω_n^{j'·k1} with j' the global column index in [0, n / n1)
    if (peer_rows) {
        uint32_t log_g = 0;
        while ((1u << log_g) < world) log_g++;
        p.peer_log_rl = log_n1 - log_g;
        p.peer_log_m = log_n - log_n1;
        for (uint32_t h = 0; h < world; h++) p.peer[h] = (Fr *)peer_rows[h];
    }
    const uint32_t tile_log = std::max(log_n1, std::min(10u, log_n1 + log_cols));
    p.logC = std::min(tile_log - log_n1, log_cols);
    const uint32_t T = 1u << (p.S + p.logC);
    const size_t smem = (size_t)8 * std::max(T, 32u) * sizeof(uint32_t);
    const uint32_t blocks = 1u << (log_cols - p.logC);
    ntt_pass_kernel<<<blocks, T >> 3, smem, ctx->stream>>>((const Fr *)data_dev, (Fr *)data_dev, p);
    PB_LAUNCHED(ctx);
    return 0;
}
extern "C" int pb200_ntt_columns_dev(pb200_ctx *ctx, uint64_t *data_dev, uint32_t log_n, uint32_t log_n1, uint32_t log_cols,
                                     uint32_t col_offset, int inverse) {
    return ntt_columns_impl(ctx, data_dev, log_n, log_n1, log_cols, col_offset, inverse, 0, nullptr);
}
extern "C" int pb200_ntt_columns_scatter_dev(pb200_ctx *ctx, uint64_t *data_dev, uint32_t log_n, uint32_t log_n1, uint32_t log_cols,
                                             uint32_t col_offset, uint32_t world, void *const *peer_row_bufs) {
    if (!ctx) return PB200_ERR_ARG;
    PB_ARG(ctx, peer_row_bufs != nullptr && world >= 1 && world <= 8 && (world & (world - 1)) == 0 && (1u << log_n1) >= world);
    for (uint32_t h = 0; h < world; h++) PB_ARG(ctx, peer_row_bufs[h] != nullptr && peer_row_bufs[h] != (void *)data_dev);
    return ntt_columns_impl(ctx, data_dev, log_n, log_n1, log_cols, col_offset, 0, world, peer_row_bufs);
}
// CUDA IPC plumbing so that one-process-per-GPU hosts can hand each other their receive buffers.
extern "C" int pb200_ipc_export(pb200_ctx *ctx, void *dev_ptr, unsigned char handle_out[64]) {
    if (!ctx) return PB200_ERR_ARG;
    PB_ARG(ctx, dev_ptr != nullptr && handle_out != nullptr);
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
    cudaIpcMemHandle_t h;
    PB_CUDA(ctx, cudaIpcGetMemHandle(&h, dev_ptr));
    memcpy(handle_out, &h, 64);
    return 0;
}
extern "C" int pb200_ipc_open(pb200_ctx *ctx, const unsigned char handle[64], void **dev_ptr_out) {
    if (!ctx) return PB200_ERR_ARG;
    PB_ARG(ctx, handle != nullptr && dev_ptr_out != nullptr);
    PB_CUDA(ctx, cudaSetDevice(ctx->device));
    cudaIpcMemHandle_t h;
    memcpy(&h, handle, 64);
    PB_CUDA(ctx, cudaIpcOpenMemHandle(dev_ptr_out, h, cudaIpcMemLazyEnablePeerAccess));
    return 0;
}
extern "C" int pb200_ipc_close(pb200_ctx *ctx, void *dev_ptr) {
    if (!ctx) return PB200_ERR_ARG;
    PB_CUDA(ctx, cudaIpcCloseMemHandle(dev_ptr));
    return 0;
}
extern "C" int pb200_block_transpose_dev(pb200_ctx *ctx, uint64_t *dst_dev, const uint64_t *src_dev, uint32_t blocks, uint32_t rows,
                                         uint32_t cols) {
    if (!ctx) return PB200_ERR_ARG;
    PB_ARG(ctx, dst_dev != nullptr && src_dev != nullptr && dst_dev != src_dev);
    PB_CUDA(ctx, cudaSetDevice(ctx->device));
    const uint64_t total = (uint64_t)blocks * rows * cols;
    if (total == 0) return 0;
    const uint32_t grid = (uint32_t)std::min<uint64_t>((total + 255) / 256, (uint64_t)ctx->sm_count * 32);
    block_transpose_kernel<<<grid, 256, 0, ctx->stream>>>((Fr *)dst_dev, (const Fr *)src_dev, blocks, rows, cols);
    PB_LAUNCHED(ctx);
    return 0;
}
