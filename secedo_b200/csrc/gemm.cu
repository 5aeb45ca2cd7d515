// K3 — int8 tcgen05/TMEM GEMM path (placeholder until the kernel lands in this round).
#include "common.cuh"

int sgpu_gemm_counts(sgpu_ctx *ctx, const sgpu_pileup *, const LinkResult &, sgpu_counts *, uint64_t *) {
    return sgpu_fail(ctx, SGPU_E_ARG, "GEMM path not built");
}
