// msm_tail.cu — the latency-bound tail of the MSM pipeline: partial-slot levels, bucket reduction, window sums,
// window combine, and the set-up kernels (pre-doubled SRS copies, synthetic bases, rank-result sum).
// These kernels run few warps with long serial chains of group operations, so they are compiled with one shared,
// non-inlined field multiplier (PB_FIELD_NOINLINE_MUL): ~10× less code (cold instruction fetch was the
// dominant cost of the small launches) and far fewer registers (more CTAs per SM for the bucket reduction).
// The algorithms are documented in msm.cu / DESIGN.md §4.3.
#define PB_FIELD_NOINLINE_MUL 1
#include <algorithm>

#include "msm_common.cuh"

namespace {

// Levels ≥ 2: partial slots (bucket id or kInvalid, XYZZ point) → buckets / next-level slots.
__global__ void __launch_bounds__(128) msm_accumulate_slots_kernel(const uint32_t *__restrict__ in_gb, const G1Xyzz *__restrict__ in_pt,
                                                                   const uint32_t *__restrict__ n_in_ptr, uint32_t L,
                                                                   G1Xyzz *buckets, uint32_t *out_gb, G1Xyzz *out_pt,
                                                                   uint32_t *n_out_ptr) {
    const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
    const uint32_t n_in = *n_in_ptr;
    // thread u runs iff u·L + 1 < n_in, so exactly ceil((n_in − 1) / L) threads write their two output slots
    if (t == 0) *n_out_ptr = n_in > 1 ? 2 * (uint32_t)(((uint64_t)n_in - 1 + L - 1) / L) : 0;
    // Segments start at odd slots: the two partials of a bucket that straddled one boundary of the previous level
    // sit at slots (2u+1, 2u+2), so an odd-aligned segmentation never splits such a pair and one level finishes
    // every bucket that is not heavy.  Slot 0 (the head slot of thread 0) is always a hole.
    const uint64_t start64 = (uint64_t)t * L + 1;
    if (start64 >= n_in) return;
    const uint32_t start = (uint32_t)start64, end = (uint32_t)min((uint64_t)n_in, start64 + L);
    const uint32_t prev = in_gb[start - 1];
    const uint32_t next = end < n_in ? in_gb[end] : kInvalid;
    RunSink sink{buckets, out_gb, out_pt, t, kInvalid, kInvalid};

    uint32_t cur = kInvalid;
    bool have = false, tl = false;
    G1Xyzz acc = G1Xyzz::identity();
    for (uint32_t i = start; i < end; i++) {
        const uint32_t gb = in_gb[i];
        if (have && gb != cur) {  // a different bucket or a hole ends the run
            sink.flush(cur, acc, tl, false);
            have = false;
        }
        if (gb == kInvalid) continue;
        if (!have) {
            have = true;
            cur = gb;
            tl = (i == start && prev == gb);
            acc = G1Xyzz::identity();
        }
        acc = g1_add(acc, load_xyzz(in_pt + i));
    }
    if (have) sink.flush(cur, acc, tl, next == cur);
    sink.finish();
}

// ---------------------------------------------------------------------------------------- reduce
// One thread per chunk of K consecutive buckets of one window:
//   Σ_{j<K} (qK + j + 1)·B_{qK+j} = Σ_j (j+1)·B_j  (running sum)  +  (qK)·Σ_j B_j  (small scalar mul)
__global__ void __launch_bounds__(128) msm_reduce_chunks_kernel(const G1Xyzz *__restrict__ buckets, const uint32_t *__restrict__ count,
                                                                MsmCfg cfg, G1Xyzz *chunk_sums) {
    const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
    const uint32_t chunks_log = cfg.nb_log - cfg.K_log;
    if (t >= (cfg.W << chunks_log)) return;
    const uint32_t w = t >> chunks_log, q = t & ((1u << chunks_log) - 1), K = 1u << cfg.K_log;
    const uint32_t base = (w << cfg.nb_log) + (q << cfg.K_log);
    G1Xyzz run = G1Xyzz::identity(), acc = G1Xyzz::identity();
    for (int j = (int)K - 1; j >= 0; j--) {
        if (count[base + j]) run = g1_add(run, load_xyzz(buckets + base + j));
        acc = g1_add(acc, run);
    }
    if (q) acc = g1_add(acc, g1_mul_small(run, (uint64_t)q << cfg.K_log));
    store_xyzz(chunk_sums + t, acc);
}
// Tree sum of point arrays: CTA (g, part) adds the items [part·per, (part+1)·per) of group g into out[g·parts + part].
// Called twice: chunk sums → `parts` partial sums per window → one sum per window.
__global__ void __launch_bounds__(128) msm_sum_kernel(const G1Xyzz *__restrict__ in, uint32_t items_per_group, uint32_t parts,
                                                      G1Xyzz *out) {
    __shared__ uint4 sm[128 * 12];
    const uint32_t g = blockIdx.x / parts, part = blockIdx.x % parts, tid = threadIdx.x;
    const uint32_t per = (items_per_group + parts - 1) / parts;
    const uint32_t lo = part * per, hi = min(lo + per, items_per_group);
    G1Xyzz acc = G1Xyzz::identity();
    for (uint32_t i = lo + tid; i < hi; i += blockDim.x) acc = g1_add(acc, load_xyzz(in + (size_t)g * items_per_group + i));
    G1Xyzz *smp = reinterpret_cast<G1Xyzz *>(sm);
    store_xyzz(smp + tid, acc);
    __syncthreads();
    for (uint32_t s = blockDim.x >> 1; s > 0; s >>= 1) {
        if (tid < s) {
            acc = g1_add(load_xyzz(smp + tid), load_xyzz(smp + tid + s));
            store_xyzz(smp + tid, acc);
        }
        __syncthreads();
    }
    if (tid == 0) store_xyzz(out + blockIdx.x, acc);
}
// XYZZ (x = X/ZZ, y = Y/ZZZ) → homogeneous projective (X·ZZZ : Y·ZZ : ZZ·ZZZ), the coordinate system of upstream's
// G1Projective.  Like upstream's result the triple is not normalised (no inversion on the device: the caller's
// `G1Affine::from` does it, exactly as with the Rust implementation); the identity is (0, R, 0).
__device__ __forceinline__ void write_projective(const G1Xyzz &p, uint32_t *result) {
    Fp X = Fp::zero(), Y = Fp::one(), Z = Fp::zero();
    if (!p.is_identity()) {
        X = p.x * p.zzz;
        Y = p.y * p.zz;
        Z = p.zz * p.zzz;
    }
    for (int i = 0; i < 12; i++) {
        result[i] = X.l[i];
        result[12 + i] = Y.l[i];
        result[24 + i] = Z.l[i];
    }
}
// Horner over the windows (c doublings each), optional accumulation across pieces.
// result: 36 words — X[12] ‖ Y[12] ‖ Z[12] homogeneous projective, (0, R, 0) for the identity.
__global__ void msm_combine_kernel(const G1Xyzz *window_sums, MsmCfg cfg, G1Xyzz *running_total, int first_piece,
                                   int last_piece, uint32_t *result) {
    G1Xyzz total = G1Xyzz::identity();
    for (int w = (int)cfg.W - 1; w >= 0; w--) {
        if (!total.is_identity())
            for (uint32_t k = 0; k < cfg.c; k++) total = g1_dbl(total);
        total = g1_add(total, load_xyzz(window_sums + w));
    }
    if (!first_piece) total = g1_add(total, load_xyzz(running_total));
    store_xyzz(running_total, total);
    if (last_piece) write_projective(total, result);
}

// Batched MSM over pre-doubled bases: bucket set j already holds the whole sum of scalar vector j.
__global__ void msm_batch_results_kernel(const G1Xyzz *set_sums, uint32_t batch, uint32_t *results) {
    const uint32_t j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j < batch) write_projective(load_xyzz(set_sums + j), results + 36 * j);
}

// Σ of `count` projective (X:Y:Z) points given as 36-word records; one thread (count is a handful of ranks).
__global__ void g1_sum_kernel(const uint32_t *pts, uint32_t count, uint32_t *result) {
    G1Xyzz total = G1Xyzz::identity();
    for (uint32_t i = 0; i < count; i++) {
        Fp X, Y, Z;
        for (int k = 0; k < 12; k++) { X.l[k] = pts[36 * i + k]; Y.l[k] = pts[36 * i + 12 + k]; Z.l[k] = pts[36 * i + 24 + k]; }
        if (Z.is_zero()) continue;
        // homogeneous (X:Y:Z) → XYZZ with ZZ = Z², ZZZ = Z³:  x = X/Z = X·Z/ZZ, y = Y/Z = Y·Z²/ZZZ
        G1Xyzz p;
        p.zz = Z.sqr();
        p.zzz = p.zz * Z;
        p.x = X * Z;
        p.y = Y * p.zz;
        total = g1_add(total, p);
    }
    write_projective(total, result);
}

// Thread j adds record j of every rank's block: pts = [count][batch] records of 36 words (projective X ‖ Y ‖ Z) as an
// all-gather of the ranks' batched MSM results delivers them; results[j] = Σ_r pts[r][j].
__global__ void g1_sum_batch_kernel(const uint32_t *pts, uint32_t count, uint32_t batch, uint32_t *results) {
    const uint32_t j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= batch) return;
    G1Xyzz total = G1Xyzz::identity();
    for (uint32_t r = 0; r < count; r++) {
        const uint32_t *q = pts + 36 * ((size_t)r * batch + j);
        Fp X, Y, Z;
        for (int k = 0; k < 12; k++) { X.l[k] = q[k]; Y.l[k] = q[12 + k]; Z.l[k] = q[24 + k]; }
        if (Z.is_zero()) continue;
        G1Xyzz p;
        p.zz = Z.sqr();
        p.zzz = p.zz * Z;
        p.x = X * Z;
        p.y = Y * p.zz;
        total = g1_add(total, p);
    }
    write_projective(total, results + 36 * j);
}

// ------------------------------------------------------------------------------ synthetic bases
// bases[i] = (a + i·d)·G; each thread walks `per` consecutive points by repeated addition of d·G.
__global__ void __launch_bounds__(128) synthetic_bases_kernel(G1Affine *out, uint64_t n, uint64_t a, uint64_t d, uint32_t per) {
    const uint64_t t = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
    const uint64_t i0 = t * per;
    if (i0 >= n) return;
    // G1 generator, Montgomery form
    G1Affine g;
    {
        const uint32_t gx[12] = {0xfd530c16u, 0x5cb38790u, 0x9976fff5u, 0x7817fc67u, 0x143ba1c1u, 0x154f95c7u,
                                 0xf3d0e747u, 0xf0ae6acdu, 0x21dbf440u, 0xedce6eccu, 0x9e0bfb75u, 0x12017741u};
        const uint32_t gy[12] = {0x0ce72271u, 0xbaac93d5u, 0x7918fd8eu, 0x8c22631au, 0x570725ceu, 0xdd595f13u,
                                 0x50405194u, 0x51ac5829u, 0xad0059c0u, 0x0e1c8c3fu, 0x5008a26au, 0x0bbc3efcu};
        for (int k = 0; k < 12; k++) { g.x.l[k] = gx[k]; g.y.l[k] = gy[k]; }
    }
    const G1Xyzz G = G1Xyzz::from_affine(g);
    G1Xyzz cur = g1_mul_small(G, a + i0 * d);
    const G1Xyzz step = g1_mul_small(G, d);
    for (uint32_t k = 0; k < per && i0 + k < n; k++) {
        G1Affine af;
        g1_to_affine(cur, af);  // (a + i·d) is never ≡ 0 mod r for the sizes used
        store_fp2(reinterpret_cast<uint4 *>(out + i0 + k), af.x, af.y);
        cur = g1_add(cur, step);
    }
}

// pre[w·n + i] = 2^(c·w)·P_i in affine form (one thread per base; c doublings then a normalisation per window).
__global__ void __launch_bounds__(128) msm_precompute_kernel(const G1Affine *__restrict__ bases, uint32_t n, uint32_t c, uint32_t W,
                                                             G1Affine *pre) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    G1Affine p;
    load_fp2(reinterpret_cast<const uint4 *>(bases + i), p.x, p.y);
    store_fp2(reinterpret_cast<uint4 *>(pre + i), p.x, p.y);
    G1Xyzz cur = G1Xyzz::from_affine(p);
    for (uint32_t w = 1; w < W; w++) {
        for (uint32_t k = 0; k < c; k++) cur = g1_dbl(cur);
        G1Affine a;
        g1_to_affine(cur, a);  // a point of prime order never doubles to the identity
        store_fp2(reinterpret_cast<uint4 *>(pre + (size_t)w * n + i), a.x, a.y);
        cur = G1Xyzz::from_affine(a);
    }
}


// ------------------------------------------------------------------------------ projective bases (pippenger, iterator form)
// Homogeneous projective (X:Y:Z) bases → packed affine, Montgomery's trick per thread: `per` points share one inversion.
// A point with Z = 0 (the identity, which the affine layout cannot hold) becomes the generator with its scalar forced to
// zero in the copy of the scalar vector the MSM then reads.
__global__ void __launch_bounds__(128) g1_normalize_kernel(const uint32_t *__restrict__ xyz, const uint64_t *__restrict__ scalars_in,
                                                           uint64_t n, G1Affine *out, uint64_t *scalars_out) {
    constexpr uint32_t per = 8;
    const uint64_t i0 = (blockIdx.x * (uint64_t)blockDim.x + threadIdx.x) * per;
    if (i0 >= n) return;
    const uint32_t cnt = (uint32_t)min((uint64_t)per, n - i0);
    auto load_z = [&](uint32_t k) {
        Fp z;
        for (int w = 0; w < 12; w++) z.l[w] = xyz[36 * (i0 + k) + 24 + w];
        return z;
    };
    Fp prefix[per];   // prefix[k] = Π_{j<k, Z_j ≠ 0} Z_j
    Fp run = Fp::one();
    for (uint32_t k = 0; k < cnt; k++) {
        prefix[k] = run;
        const Fp z = load_z(k);
        if (!z.is_zero()) run = run * z;
    }
    Fp inv = run.inv();   // 1 / Π Z_j
    for (int k = (int)cnt - 1; k >= 0; k--) {
        const Fp z = load_z((uint32_t)k);
        G1Affine a;
        uint4 s0 = reinterpret_cast<const uint4 *>(scalars_in + 4 * (i0 + k))[0], s1 = reinterpret_cast<const uint4 *>(scalars_in + 4 * (i0 + k))[1];
        if (z.is_zero()) {
            const uint32_t gx[12] = {0xfd530c16u, 0x5cb38790u, 0x9976fff5u, 0x7817fc67u, 0x143ba1c1u, 0x154f95c7u,
                                     0xf3d0e747u, 0xf0ae6acdu, 0x21dbf440u, 0xedce6eccu, 0x9e0bfb75u, 0x12017741u};
            const uint32_t gy[12] = {0x0ce72271u, 0xbaac93d5u, 0x7918fd8eu, 0x8c22631au, 0x570725ceu, 0xdd595f13u,
                                     0x50405194u, 0x51ac5829u, 0xad0059c0u, 0x0e1c8c3fu, 0x5008a26au, 0x0bbc3efcu};
            for (int w = 0; w < 12; w++) { a.x.l[w] = gx[w]; a.y.l[w] = gy[w]; }
            s0 = make_uint4(0, 0, 0, 0);
            s1 = s0;
        } else {
            const Fp zi = inv * prefix[k];   // 1 / Z_k
            inv = inv * z;
            Fp X, Y;
            for (int w = 0; w < 12; w++) { X.l[w] = xyz[36 * (i0 + k) + w]; Y.l[w] = xyz[36 * (i0 + k) + 12 + w]; }
            a.x = X * zi;
            a.y = Y * zi;
        }
        store_fp2(reinterpret_cast<uint4 *>(out + i0 + k), a.x, a.y);
        reinterpret_cast<uint4 *>(scalars_out + 4 * (i0 + k))[0] = s0;
        reinterpret_cast<uint4 *>(scalars_out + 4 * (i0 + k))[1] = s1;
    }
}

}  // namespace

// ---- launchers (declared in msm_common.cuh) -----------------------------------------------------------
int tail_accumulate_slots(pb200_ctx *ctx, uint32_t grid_threads, const uint32_t *in_gb, const G1Xyzz *in_pt, const uint32_t *n_in_ptr,
                          uint32_t L, G1Xyzz *buckets, uint32_t *out_gb, G1Xyzz *out_pt, uint32_t *n_out_ptr) {
    const uint32_t block = grid_threads >= 128 ? 128 : 32;
    msm_accumulate_slots_kernel<<<(grid_threads + block - 1) / block, block, 0, ctx->stream>>>(in_gb, in_pt, n_in_ptr, L, buckets, out_gb,
                                                                                             out_pt, n_out_ptr);
    PB_LAUNCHED(ctx);
    return 0;
}
int tail_reduce_chunks(pb200_ctx *ctx, uint32_t n_chunks, const G1Xyzz *buckets, const uint32_t *count, MsmCfg cfg, G1Xyzz *chunk_sums) {
    msm_reduce_chunks_kernel<<<(n_chunks + 127) / 128, 128, 0, ctx->stream>>>(buckets, count, cfg, chunk_sums);
    PB_LAUNCHED(ctx);
    return 0;
}
int tail_sum(pb200_ctx *ctx, uint32_t groups, uint32_t parts, const G1Xyzz *in, uint32_t items_per_group, G1Xyzz *out) {
    msm_sum_kernel<<<groups * parts, 128, 0, ctx->stream>>>(in, items_per_group, parts, out);
    PB_LAUNCHED(ctx);
    return 0;
}
int tail_combine(pb200_ctx *ctx, const G1Xyzz *window_sums, MsmCfg cfg, G1Xyzz *running_total, int first_piece, int last_piece,
                 uint32_t *result) {
    msm_combine_kernel<<<1, 1, 0, ctx->stream>>>(window_sums, cfg, running_total, first_piece, last_piece, result);
    PB_LAUNCHED(ctx);
    return 0;
}
int tail_batch_results(pb200_ctx *ctx, const G1Xyzz *set_sums, uint32_t batch, uint32_t *results) {
    msm_batch_results_kernel<<<1, 32, 0, ctx->stream>>>(set_sums, batch, results);
    PB_LAUNCHED(ctx);
    return 0;
}
int tail_g1_sum(pb200_ctx *ctx, const uint32_t *pts, uint32_t count, uint32_t *result) {
    g1_sum_kernel<<<1, 1, 0, ctx->stream>>>(pts, count, result);
    PB_LAUNCHED(ctx);
    return 0;
}
int tail_g1_sum_batch(pb200_ctx *ctx, const uint32_t *pts, uint32_t count, uint32_t batch, uint32_t *results) {
    g1_sum_batch_kernel<<<(batch + 31) / 32, 32, 0, ctx->stream>>>(pts, count, batch, results);
    PB_LAUNCHED(ctx);
    return 0;
}
int tail_precompute(pb200_ctx *ctx, const G1Affine *bases, uint32_t n, uint32_t c, uint32_t W, G1Affine *pre) {
    msm_precompute_kernel<<<(n + 127) / 128, 128, 0, ctx->stream>>>(bases, n, c, W, pre);
    PB_LAUNCHED(ctx);
    return 0;
}
int tail_synthetic_bases(pb200_ctx *ctx, G1Affine *out, uint64_t n, uint64_t a, uint64_t d) {
    const uint32_t per = 32;
    const uint64_t threads = (n + per - 1) / per;
    synthetic_bases_kernel<<<(uint32_t)((threads + 127) / 128), 128, 0, ctx->stream>>>(out, n, a, d, per);
    PB_LAUNCHED(ctx);
    return 0;
}
int tail_g1_normalize(pb200_ctx *ctx, const uint32_t *xyz, const uint64_t *scalars_in, uint64_t n, G1Affine *out, uint64_t *scalars_out) {
    const uint64_t threads = (n + 7) / 8;
    g1_normalize_kernel<<<(uint32_t)((threads + 127) / 128), 128, 0, ctx->stream>>>(xyz, scalars_in, n, out, scalars_out);
    PB_LAUNCHED(ctx);
    return 0;
}
