// K4 — first-order read-pair counts by per-locus cross-cell pair scatter.
//
// For every locus and every pair of surviving entries (a, b) from different cells that are not both
// tail reads, one int32 counter is incremented: S[i][j] if the bases agree, D[i][j] otherwise, with
// i < j the two cells (only the upper triangle is kept). This is the per-locus incidence count
// sum_pairs x_s / x_d of similarity_matrix.cpp:221-229; pairs that overlap at several loci are
// completed by multilocus.cu. The same kernel, restricted to tail x tail pairs and with sign -1,
// removes from the GEMM path's result the pairs the reference never compares.
//
// One CTA per locus; the locus' entry codes are staged through shared memory in tiles so that a
// locus of any size (c = 25 ... 20 000) is handled; pairs are enumerated from a flattened index so
// all threads are busy for small loci too. The counters are RED.ADD (no return value) on L2/HBM.
#include "common.cuh"

#include <algorithm>

namespace {

constexpr int SC_THREADS = 256;
constexpr uint32_t SC_TILE = 2048;

__device__ __forceinline__ void tri_unrank(uint32_t k, uint32_t &i, uint32_t &j) {
    // k -> (i < j) with k = j (j - 1) / 2 + i
    j = static_cast<uint32_t>((1.f + sqrtf(1.f + 8.f * static_cast<float>(k))) * 0.5f);
    while (j * (j - 1) / 2 > k) {
        --j;
    }
    while ((j + 1) * j / 2 <= k) {
        ++j;
    }
    i = k - j * (j - 1) / 2;
}

__global__ void __launch_bounds__(SC_THREADS) scatter_pairs_kernel(
        const uint64_t *__restrict__ row_ptr, const uint32_t *__restrict__ code, uint64_t n_loci,
        int32_t *__restrict__ S, int32_t *__restrict__ D, uint32_t n_cells, int sign, int only_tail,
        const uint32_t *__restrict__ locus_list /* null: all loci */, unsigned long long *__restrict__ n_pairs) {
    __shared__ uint32_t sA[SC_TILE];
    __shared__ uint32_t sB[SC_TILE];
    unsigned long long local_pairs = 0;
    for (uint64_t li = blockIdx.x; li < n_loci; li += gridDim.x) {
        const uint64_t l = locus_list ? locus_list[li] : li;
        const uint64_t e0 = row_ptr[l];
        const uint64_t n = row_ptr[l + 1] - e0;
        const uint32_t n_tiles = static_cast<uint32_t>((n + SC_TILE - 1) / SC_TILE);
        for (uint32_t ta = 0; ta < n_tiles; ++ta) {
            const uint64_t a0 = e0 + static_cast<uint64_t>(ta) * SC_TILE;
            const uint32_t nA = static_cast<uint32_t>(min(static_cast<uint64_t>(SC_TILE), e0 + n - a0));
            __syncthreads();
            for (uint32_t t = threadIdx.x; t < nA; t += SC_THREADS) {
                sA[t] = code[a0 + t];
            }
            for (uint32_t tb = ta; tb < n_tiles; ++tb) {
                const uint64_t b0 = e0 + static_cast<uint64_t>(tb) * SC_TILE;
                const uint32_t nB = static_cast<uint32_t>(min(static_cast<uint64_t>(SC_TILE), e0 + n - b0));
                const bool diag = tb == ta;
                __syncthreads();
                if (!diag) {
                    for (uint32_t t = threadIdx.x; t < nB; t += SC_THREADS) {
                        sB[t] = code[b0 + t];
                    }
                    __syncthreads();
                }
                const uint32_t *pB = diag ? sA : sB;
                const uint32_t total = diag ? nA * (nA - 1) / 2 : nA * nB;
                for (uint32_t k = threadIdx.x + SC_THREADS * blockIdx.y; k < total; k += SC_THREADS * gridDim.y) {
                    uint32_t i, j;
                    if (diag) {
                        tri_unrank(k, i, j);
                    } else {
                        i = k / nB;
                        j = k - i * nB;
                    }
                    const uint32_t ca = sA[i], cb = pB[j];
                    if (ca == CODE_DROPPED || cb == CODE_DROPPED) {
                        continue;
                    }
                    const uint32_t both_tail = (ca & cb & 2u) >> 1;
                    if (both_tail != static_cast<uint32_t>(only_tail)) {
                        continue;
                    }
                    const uint32_t ia = code_cell(ca), ib = code_cell(cb);
                    if (ia == ib) {
                        continue; // reads of the same cell are never compared (similarity_matrix.cpp:215)
                    }
                    const uint32_t lo = min(ia, ib), hi = max(ia, ib);
                    int32_t *plane = ((ca ^ cb) & 0xCu) == 0 ? S : D;
                    atomicAdd(&plane[static_cast<uint64_t>(lo) * n_cells + hi], sign);
                    ++local_pairs;
                }
            }
        }
    }
    // block-level sum of the pair counter
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        local_pairs += __shfl_xor_sync(0xffffffffu, local_pairs, o);
    }
    if ((threadIdx.x & 31) == 0 && local_pairs) {
        atomicAdd(n_pairs, local_pairs);
    }
}

} // namespace

int sgpu_scatter_pairs(sgpu_ctx *ctx, const sgpu_pileup *p, const LinkResult &lr, sgpu_counts *c,
                       int sign, bool only_tail_pairs, uint64_t *n_pairs) {
    cudaStream_t st = ctx->stream;
    if (n_pairs) {
        *n_pairs = 0;
    }
    if (p->n_loci == 0) {
        return SGPU_OK;
    }
    DevBuf<unsigned long long> d_np;
    SGPU_CUDA(ctx, d_np.alloc(1, ctx));
    SGPU_CUDA(ctx, cudaMemsetAsync(d_np.p, 0, sizeof(unsigned long long), st));
    // tail x tail pairs only exist at the loci behind the cutoff of each chromosome; a ranged call counts its own loci
    const uint64_t n_loci = only_tail_pairs ? lr.n_tail_loci : lr.ranged ? lr.n_own : p->n_loci;
    if (n_loci == 0) {
        return SGPU_OK;
    }
    const unsigned gx = static_cast<unsigned>(std::min<uint64_t>(n_loci, static_cast<uint64_t>(ctx->sm_count) * 16));
    // few loci: split each locus' pair index space over blockIdx.y so that the whole GPU is busy
    const unsigned gy = gx >= 4u * ctx->sm_count ? 1u : std::min(64u, (4u * ctx->sm_count + gx - 1) / gx);
    const dim3 grid(gx, gy);
    SGPU_LAUNCH(ctx, (scatter_pairs_kernel<<<grid, SC_THREADS, 0, st>>>(p->d_row_ptr, lr.code.p, n_loci, c->i32 + PLANE_S * c->nn,
                                                     c->i32 + PLANE_D * c->nn, c->n, sign, only_tail_pairs ? 1 : 0,
                                                     only_tail_pairs ? lr.tail_loci.p : lr.ranged ? lr.own_loci.p : nullptr, d_np.p)));
    SGPU_CUDA(ctx, cudaGetLastError());
    if (n_pairs) {
        SGPU_CUDA(ctx, cudaMemcpyAsync(&ctx->h_scratch[0], d_np.p, sizeof(uint64_t), cudaMemcpyDeviceToHost, st));
        SGPU_CUDA(ctx, cudaStreamSynchronize(st));
        *n_pairs = ctx->h_scratch[0];
    }
    return SGPU_OK;
}
