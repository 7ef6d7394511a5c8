// kzg.cu — the KZG10 layer directly above the MSM: SRS generation and opening witnesses.
//
// Replaces, on the GPU, the bulk arithmetic of dusk-plonk 0.8.2 `commitment_scheme::kzg10`
// (crate pinned at /root/reference/Cargo.toml:19):
//   PublicParameters::setup            powers_of_g[i] = τ^i·G                      → pb200_srs_generate
//   CommitKey::commit                  one MSM over powers_of_g[..len]              → pb200_msm_g1(_dev)   (msm.cu)
//   CommitKey::compute_single_witness  (p(X) − p(z)) / (X − z) by Ruffini's rule    → pb200_kzg_witness_dev
//   CommitKey::compute_aggregate_witness  Σ vⁱ·pᵢ, then the single witness          → pb200_fr_horner_step_dev + witness
// Pairing checks, (de)serialisation of the parameters and the transcript stay on the host.
//
// Ruffini's recurrence q_{i−1} = a_i + z·q_i is sequential; here q_i = z^{−(i+1)} · Σ_{j>i} a_j z^j, i.e. an
// elementwise multiply by z^j, a suffix sum over Fr (three-kernel scan) and an elementwise multiply by z^{−(i+1)};
// p(z) is the total of the same scan.  z = 0 degenerates to a shift.
#define PB_FIELD_NOINLINE_MUL 1
#include <algorithm>
#include <cstring>

#include "fr_vec.cuh"
#include "msm_common.cuh"

namespace {

__device__ __forceinline__ G1Affine generator_affine() {  // G1 generator, Montgomery form
    const uint32_t gx[12] = {0xfd530c16u, 0x5cb38790u, 0x9976fff5u, 0x7817fc67u, 0x143ba1c1u, 0x154f95c7u,
                             0xf3d0e747u, 0xf0ae6acdu, 0x21dbf440u, 0xedce6eccu, 0x9e0bfb75u, 0x12017741u};
    const uint32_t gy[12] = {0x0ce72271u, 0xbaac93d5u, 0x7918fd8eu, 0x8c22631au, 0x570725ceu, 0xdd595f13u,
                             0x50405194u, 0x51ac5829u, 0xad0059c0u, 0x0e1c8c3fu, 0x5008a26au, 0x0bbc3efcu};
    G1Affine g;
    for (int k = 0; k < 12; k++) { g.x.l[k] = gx[k]; g.y.l[k] = gy[k]; }
    return g;
}

// ---- SRS generation -----------------------------------------------------------------------------------------
// table[w·255 + (d−1)] = d·2^(8w)·G, w < 32, d = 1..255 (affine): first the 32 window bases 2^(8w)·G (one thread
// each, 8w doublings), then one thread per table entry (a small scalar multiple and a normalisation).
__global__ void srs_window_bases_kernel(G1Affine *bases32) {
    const uint32_t w = blockIdx.x * blockDim.x + threadIdx.x;
    if (w >= 32) return;
    G1Xyzz base = G1Xyzz::from_affine(generator_affine());
    for (uint32_t k = 0; k < 8 * w; k++) base = g1_dbl(base);
    G1Affine a;
    g1_to_affine(base, a);
    store_fp2(reinterpret_cast<uint4 *>(bases32 + w), a.x, a.y);
}
__global__ void __launch_bounds__(128) srs_table_kernel(const G1Affine *bases32, G1Affine *table) {
    const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= 32 * 255) return;
    const uint32_t w = t / 255, d = t % 255 + 1;
    G1Affine b;
    load_fp2(reinterpret_cast<const uint4 *>(bases32 + w), b.x, b.y);
    G1Affine a;
    g1_to_affine(g1_mul_small(G1Xyzz::from_affine(b), d), a);
    store_fp2(reinterpret_cast<uint4 *>(table + t), a.x, a.y);
}
// out[i] = τ^i·G: τ^i by square-and-multiply, then one table lookup and mixed addition per scalar byte.
__global__ void __launch_bounds__(128) srs_powers_kernel(const G1Affine *__restrict__ table, const Fr *tau, uint32_t first, uint32_t n,
                                                         G1Affine *out) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const Fr s = ld_fr(tau).pow_u32(first + i).from_mont();
    G1Xyzz acc = G1Xyzz::identity();
    for (uint32_t w = 0; w < 32; w++) {
        const uint32_t d = (s.l[w >> 2] >> (8 * (w & 3))) & 0xff;
        if (d) {
            G1Affine t;
            load_fp2(reinterpret_cast<const uint4 *>(table + w * 255 + d - 1), t.x, t.y);
            g1_madd(acc, t);
        }
    }
    G1Affine a;
    if (!g1_to_affine(acc, a)) { a.x = Fp::zero(); a.y = Fp::zero(); }  // τ^i ≡ 0 cannot happen for τ ≠ 0
    store_fp2(reinterpret_cast<uint4 *>(out + i), a.x, a.y);
}

// ---- Fr vector helpers ---------------------------------------------------------------------------------------
// acc[j] = acc[j]·c + p[j]   (one Horner step of Σ vⁱ·pᵢ over whole polynomials; p shorter than acc is zero-padded)
__global__ void fr_horner_step_kernel(Fr *acc, const Fr *p, uint32_t n_acc, uint32_t n_p, const Fr *c) {
    const uint32_t j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= n_acc) return;
    Fr v = ld_fr(acc + j) * ld_fr(c);
    if (j < n_p) v = v + ld_fr(p + j);
    st_fr(acc + j, v);
}
// consts[0] = z, consts[1] = z⁻¹ (0 ↦ 0), consts[2] = 1
__global__ void witness_consts_kernel(const Fr *z, Fr *consts) {
    const Fr v = ld_fr(z);
    st_fr(consts + 0, v);
    st_fr(consts + 1, v.inv());
    st_fr(consts + 2, Fr::one());
}
// b[j] = a[j]·z^j, four consecutive j per thread (one power by square-and-multiply, then running products)
// (j_off: global index of a[0] when a / b are a rank's coefficient slice of a sharded polynomial)
__global__ void witness_scale_kernel(const Fr *a, Fr *b, uint32_t n, const Fr *consts, uint32_t j_off) {
    const uint32_t j0 = (blockIdx.x * blockDim.x + threadIdx.x) * 4;
    if (j0 >= n) return;
    const Fr z = ld_fr(consts + 0);
    Fr p = z.pow_u32(j_off + j0);
    for (uint32_t k = 0; k < 4 && j0 + k < n; k++) {
        st_fr(b + j0 + k, ld_fr(a + j0 + k) * p);
        p = p * z;
    }
}
// Suffix sums over Fr in three kernels (tiles of 1024 elements, 256 threads × 4).  S[i] = Σ_{j ≥ i} b[j].
constexpr uint32_t kFrScanTile = 1024;
__device__ __forceinline__ Fr block_suffix_scan(Fr v, Fr *sm, Fr *total) {  // inclusive suffix scan across 256 threads
    const uint32_t t = threadIdx.x;
    st_fr(sm + t, v);
    __syncthreads();
    for (uint32_t o = 1; o < 256; o <<= 1) {
        Fr add = Fr::zero();
        if (t + o < 256) add = ld_fr(sm + t + o);
        __syncthreads();
        v = v + add;
        st_fr(sm + t, v);
        __syncthreads();
    }
    if (total) *total = ld_fr(sm + 0);
    return v;
}
__global__ void __launch_bounds__(256) fr_scan_tile_sums_kernel(const Fr *b, uint32_t n, Fr *tile_sums) {
    __shared__ uint4 smraw[256 * 2];
    Fr *sm = reinterpret_cast<Fr *>(smraw);
    const uint32_t base = blockIdx.x * kFrScanTile + threadIdx.x * 4;
    Fr s = Fr::zero();
    for (uint32_t k = 0; k < 4; k++)
        if (base + k < n) s = s + ld_fr(b + base + k);
    Fr total;
    block_suffix_scan(s, sm, &total);
    if (threadIdx.x == 0) st_fr(tile_sums + blockIdx.x, total);
}
// tile_sums[t] ← Σ_{u > t} tile_sums[u] (exclusive suffix), single CTA; total of everything → *grand_total
__global__ void __launch_bounds__(256) fr_scan_top_kernel(Fr *tile_sums, uint32_t n_tiles, Fr *grand_total) {
    __shared__ uint4 smraw[256 * 2];
    Fr *sm = reinterpret_cast<Fr *>(smraw);
    const uint32_t per = (n_tiles + 255) / 256;
    const uint32_t lo = threadIdx.x * per, hi = min(lo + per, n_tiles);
    Fr s = Fr::zero();
    for (uint32_t i = lo; i < hi; i++) s = s + ld_fr(tile_sums + i);
    Fr total;
    Fr inc = block_suffix_scan(s, sm, &total);
    Fr run = inc - s;  // sum of the later threads' ranges
    for (uint32_t i = hi; i-- > lo;) {
        Fr v = ld_fr(tile_sums + i);
        st_fr(tile_sums + i, run);
        run = run + v;
    }
    if (threadIdx.x == 0) st_fr(grand_total, total);
}
// q[i] = z^{−(i+1)} · S[i+1]  with S the suffix sums of b (q[n−1] = 0)
// (sliced: indices are global index − j_off, and `carry` = Σ of b over every later slice, gathered from the other ranks)
__global__ void __launch_bounds__(256) witness_finish_kernel(const Fr *b, uint32_t n, const Fr *tile_sums, const Fr *consts, Fr *q,
                                                             uint32_t j_off, const Fr *carry) {
    __shared__ uint4 smraw[256 * 2];
    Fr *sm = reinterpret_cast<Fr *>(smraw);
    const uint32_t base = blockIdx.x * kFrScanTile + threadIdx.x * 4;
    Fr v[4], s = Fr::zero();
    for (int k = 3; k >= 0; k--) {
        v[k] = (base + k < n) ? ld_fr(b + base + k) : Fr::zero();
        s = s + v[k];
    }
    Fr inc = block_suffix_scan(s, sm, nullptr);
    Fr run = (inc - s) + ld_fr(tile_sums + blockIdx.x);  // Σ of everything after this thread's four elements
    if (carry) run = run + ld_fr(carry);
    const Fr zinv = ld_fr(consts + 1);
    Fr zi[4];  // z^{−(base+k+1)}
    zi[0] = zinv.pow_u32(j_off + base + 1);
    for (int k = 1; k < 4; k++) zi[k] = zi[k - 1] * zinv;
    for (int k = 3; k >= 0; k--) {
        // run = S[base+k+1]
        if (base + k < n) st_fr(q + base + k, run * zi[k]);
        run = run + v[k];
    }
}
// z = 0: q[i] = a[i+1], q[n−1] = 0, p(0) = a[0]
__global__ void witness_shift_kernel(const Fr *a, uint32_t n, Fr *q, Fr *eval) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i == 0) st_fr(eval, ld_fr(a));
    if (i >= n) return;
    st_fr(q + i, i + 1 < n ? ld_fr(a + i + 1) : Fr::zero());
}

}  // namespace

extern "C" int pb200_srs_generate(pb200_ctx *ctx, const uint64_t tau_mont[4], size_t n_points, pb200_srs **out) {
    return pb200_srs_generate_range(ctx, tau_mont, 0, n_points, out);
}
extern "C" int pb200_srs_generate_range(pb200_ctx *ctx, const uint64_t tau_mont[4], size_t first, size_t n_points, pb200_srs **out) {
    if (!ctx) return PB200_ERR_ARG;
    PB_ARG(ctx, tau_mont != nullptr && out != nullptr && n_points >= 1 && first + n_points < ((size_t)1 << 31));
    PB_ARG(ctx, (tau_mont[0] | tau_mont[1] | tau_mont[2] | tau_mont[3]) != 0);
    PB_CUDA(ctx, cudaSetDevice(ctx->device));
    void *table = nullptr, *tau = nullptr, *pts = nullptr;
    PB_CUDA(ctx, cudaMalloc(&table, (32 * 255 + 32) * sizeof(G1Affine)));
    cudaError_t e = cudaMalloc(&tau, sizeof(Fr));
    if (e == cudaSuccess) e = cudaMalloc(&pts, n_points * sizeof(G1Affine));
    if (e == cudaSuccess) e = cudaMemcpyAsync(tau, tau_mont, 32, cudaMemcpyHostToDevice, ctx->stream);
    if (e == cudaSuccess) {
        G1Affine *bases32 = (G1Affine *)table + 32 * 255;
        srs_window_bases_kernel<<<1, 32, 0, ctx->stream>>>(bases32);
        srs_table_kernel<<<(32 * 255 + 127) / 128, 128, 0, ctx->stream>>>(bases32, (G1Affine *)table);
        ctx->launches++;
        srs_powers_kernel<<<(uint32_t)((n_points + 127) / 128), 128, 0, ctx->stream>>>((const G1Affine *)table, (const Fr *)tau,
                                                                                     (uint32_t)first, (uint32_t)n_points, (G1Affine *)pts);
        ctx->launches += 2;
        e = cudaGetLastError();
    }
    if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
    cudaFree(table);
    cudaFree(tau);
    if (e != cudaSuccess) {
        cudaFree(pts);
        return pb_fail(ctx, PB200_ERR_CUDA, "srs generate", cudaGetErrorString(e), __FILE__, __LINE__);
    }
    pb200_srs *s = new pb200_srs();
    s->dev = (const uint64_t *)pts;
    s->n = n_points;
    s->owned = true;
    *out = s;
    return 0;
}
extern "C" const uint64_t *pb200_srs_dev_ptr(const pb200_srs *srs) { return srs ? srs->dev : nullptr; }

extern "C" int pb200_fr_horner_step_dev(pb200_ctx *ctx, uint64_t *acc_dev, size_t n_acc, const uint64_t *poly_dev, size_t n_poly,
                                        const uint64_t c_mont[4]) {
    if (!ctx) return PB200_ERR_ARG;
    PB_ARG(ctx, acc_dev != nullptr && c_mont != nullptr && (poly_dev != nullptr || n_poly == 0) && n_poly <= n_acc &&
                    n_acc < ((size_t)1 << 32));
    if (n_acc == 0) return 0;
    PB_CUDA(ctx, cudaSetDevice(ctx->device));
    void *c = nullptr;
    PB_CUDA(ctx, cudaMallocAsync(&c, sizeof(Fr), ctx->stream));
    PB_CUDA(ctx, cudaMemcpyAsync(c, c_mont, 32, cudaMemcpyHostToDevice, ctx->stream));
    fr_horner_step_kernel<<<(uint32_t)((n_acc + 255) / 256), 256, 0, ctx->stream>>>((Fr *)acc_dev, (const Fr *)poly_dev, (uint32_t)n_acc,
                                                                                  (uint32_t)n_poly, (const Fr *)c);
    PB_LAUNCHED(ctx);
    cudaFreeAsync(c, ctx->stream);
    return 0;
}
extern "C" int pb200_kzg_witness_dev(pb200_ctx *ctx, const uint64_t *poly_dev, size_t n, const uint64_t z_mont[4],
                                     uint64_t *quotient_dev, uint64_t eval_mont_out[4]) {
    if (!ctx) return PB200_ERR_ARG;
    PB_ARG(ctx, poly_dev != nullptr && quotient_dev != nullptr && z_mont != nullptr && eval_mont_out != nullptr);
    PB_ARG(ctx, n >= 1 && n < ((size_t)1 << 32) && poly_dev != quotient_dev);
    PB_CUDA(ctx, cudaSetDevice(ctx->device));
    const uint32_t n32 = (uint32_t)n, n_tiles = (n32 + kFrScanTile - 1) / kFrScanTile;
    void *small = nullptr;  // z | consts[3] | grand total | tile sums
    PB_CUDA(ctx, cudaMallocAsync(&small, (size_t)(5 + n_tiles) * sizeof(Fr), ctx->stream));
    Fr *zd = (Fr *)small, *consts = zd + 1, *total = zd + 4, *tiles = zd + 5;
    cudaStream_t st = ctx->stream;
    PB_CUDA(ctx, cudaMemcpyAsync(zd, z_mont, 32, cudaMemcpyHostToDevice, st));
    const bool z_is_zero = (z_mont[0] | z_mont[1] | z_mont[2] | z_mont[3]) == 0;
    Fr *q = (Fr *)quotient_dev;
    if (z_is_zero) {
        witness_shift_kernel<<<(n32 + 255) / 256, 256, 0, st>>>((const Fr *)poly_dev, n32, q, total);
        PB_LAUNCHED(ctx);
    } else {
        witness_consts_kernel<<<1, 1, 0, st>>>(zd, consts);
        PB_LAUNCHED(ctx);
        // b = a·z^j is staged in the quotient buffer, the finishing pass reads its own tile before overwriting it
        witness_scale_kernel<<<(n32 + 1023) / 1024, 256, 0, st>>>((const Fr *)poly_dev, q, n32, consts, 0);
        PB_LAUNCHED(ctx);
        fr_scan_tile_sums_kernel<<<n_tiles, 256, 0, st>>>(q, n32, tiles);
        PB_LAUNCHED(ctx);
        fr_scan_top_kernel<<<1, 256, 0, st>>>(tiles, n_tiles, total);
        PB_LAUNCHED(ctx);
        witness_finish_kernel<<<n_tiles, 256, 0, st>>>(q, n32, tiles, consts, q, 0, nullptr);
        PB_LAUNCHED(ctx);
    }
    cudaError_t e = cudaMemcpyAsync(ctx->pinned, total, 32, cudaMemcpyDeviceToHost, st);
    cudaError_t e2 = cudaStreamSynchronize(st);
    cudaFreeAsync(small, st);
    if (e != cudaSuccess || e2 != cudaSuccess)
        return pb_fail(ctx, PB200_ERR_CUDA, "kzg witness", cudaGetErrorString(e != cudaSuccess ? e : e2), __FILE__, __LINE__);
    memcpy(eval_mont_out, ctx->pinned, 32);
    return 0;
}

// ---- sliced witness for the sharded prover (one process per GPU): rank r holds coefficients [lo, lo + cnt) of p(X) and needs
// the same slice of q(X) = (p(X) − p(z)) / (X − z) for its share of the commitment.  q_i = z^{−(i+1)}·Σ_{j>i} a_j z^j splits into
// the sum inside the slice (local scan) plus the total of every later slice — 32 bytes per rank, exchanged by the caller between
// the two phases.  `work` must hold 5 + ⌈cnt / 1024⌉ scalars (z | consts[3] | slice total | tile sums) and is owned by the caller.
int kzg_witness_slice_phase1(pb200_ctx *ctx, const uint64_t *poly_slice_dev, uint32_t lo, uint32_t cnt, const uint64_t z_mont[4],
                             uint64_t *q_slice_dev, uint64_t *work_dev, uint64_t slice_total_out[4]) {
    const uint32_t n_tiles = (cnt + kFrScanTile - 1) / kFrScanTile;
    Fr *zd = (Fr *)work_dev, *consts = zd + 1, *total = zd + 4, *tiles = zd + 5;
    cudaStream_t st = ctx->stream;
    PB_CUDA(ctx, cudaMemcpyAsync(zd, z_mont, 32, cudaMemcpyHostToDevice, st));
    witness_consts_kernel<<<1, 1, 0, st>>>(zd, consts);
    PB_LAUNCHED(ctx);
    witness_scale_kernel<<<(cnt + 1023) / 1024, 256, 0, st>>>((const Fr *)poly_slice_dev, (Fr *)q_slice_dev, cnt, consts, lo);
    PB_LAUNCHED(ctx);
    fr_scan_tile_sums_kernel<<<n_tiles, 256, 0, st>>>((const Fr *)q_slice_dev, cnt, tiles);
    PB_LAUNCHED(ctx);
    fr_scan_top_kernel<<<1, 256, 0, st>>>(tiles, n_tiles, total);
    PB_LAUNCHED(ctx);
    PB_CUDA(ctx, cudaMemcpyAsync(ctx->pinned, total, 32, cudaMemcpyDeviceToHost, st));
    PB_CUDA(ctx, cudaStreamSynchronize(st));
    memcpy(slice_total_out, ctx->pinned, 32);
    return 0;
}
int kzg_witness_slice_phase2(pb200_ctx *ctx, uint32_t lo, uint32_t cnt, uint64_t *q_slice_dev, uint64_t *work_dev,
                             const uint64_t later_slices_total[4]) {
    const uint32_t n_tiles = (cnt + kFrScanTile - 1) / kFrScanTile;
    Fr *zd = (Fr *)work_dev, *consts = zd + 1, *total = zd + 4, *tiles = zd + 5;
    cudaStream_t st = ctx->stream;
    PB_CUDA(ctx, cudaMemcpyAsync(total, later_slices_total, 32, cudaMemcpyHostToDevice, st));   // the slot is free again: reuse it for the carry
    witness_finish_kernel<<<n_tiles, 256, 0, st>>>((const Fr *)q_slice_dev, cnt, tiles, consts, (Fr *)q_slice_dev, lo, total);
    PB_LAUNCHED(ctx);
    return 0;
}
size_t kzg_witness_slice_work_scalars(uint32_t cnt) { return 5 + (cnt + kFrScanTile - 1) / kFrScanTile; }
