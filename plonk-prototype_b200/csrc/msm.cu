// msm.cu — placeholder until the Pippenger pipeline lands (next commit).
#include "common.cuh"
int msm_module_init(pb200_ctx *) { return 0; }
extern "C" int pb200_srs_upload(pb200_ctx *ctx, const uint64_t *, size_t, pb200_srs **) { return pb_fail(ctx, PB200_ERR_ARG, "msm", "not built yet", __FILE__, __LINE__); }
extern "C" int pb200_srs_wrap_dev(pb200_ctx *ctx, const uint64_t *, size_t, pb200_srs **) { return pb_fail(ctx, PB200_ERR_ARG, "msm", "not built yet", __FILE__, __LINE__); }
extern "C" void pb200_srs_free(pb200_ctx *, pb200_srs *) {}
extern "C" size_t pb200_srs_len(const pb200_srs *) { return 0; }
extern "C" int pb200_msm_g1(pb200_ctx *ctx, const pb200_srs *, size_t, const uint64_t *, size_t, uint64_t *) { return pb_fail(ctx, PB200_ERR_ARG, "msm", "not built yet", __FILE__, __LINE__); }
extern "C" int pb200_msm_g1_dev(pb200_ctx *ctx, const pb200_srs *, size_t, const uint64_t *, size_t, uint64_t *) { return pb_fail(ctx, PB200_ERR_ARG, "msm", "not built yet", __FILE__, __LINE__); }
extern "C" uint32_t pb200_msm_window_bits(size_t) { return 0; }
extern "C" int pb200_synthetic_bases_dev(pb200_ctx *ctx, uint64_t *, size_t, uint64_t, uint64_t) { return pb_fail(ctx, PB200_ERR_ARG, "msm", "not built yet", __FILE__, __LINE__); }
