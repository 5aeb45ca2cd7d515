// K5 — exact correction for read pairs that overlap at two or more loci.
//
// The reference scores a read pair once, with the log-likelihood of its WHOLE overlap (x_s, x_d)
// (similarity_matrix.cpp:221-241), which is not additive over loci. The first-order planes S, D
// count each shared locus separately; this kernel finds the pairs with x_s + x_d >= 2, accounts
// each at the first locus the two reads share, and records its overlap class:
//     x_s + x_d == 2  -> int32 planes H2[(2,0),(1,1),(0,2)]
//     x_s + x_d == 3  -> int32 planes H3[(3,0),(2,1),(1,2),(0,3)]
//     x_s + x_d >= 4  -> fp64 spill plane, += G(x_s, x_d) (rare; allocated on demand)
// plus a global class histogram. The epilogue then adds G(s,d) = F(s,d) - s F(1,0) - d F(0,1) per
// recorded pair, which together with F(1,0) S + F(0,1) D gives exactly sum_pairs F(x_s, x_d).
//
// Only entries of reads that keep >= 2 loci take part, and read linking (reads.cu) already lists
// them: they are among the "special" entries, numbered in entry order (hence grouped by locus). Each
// thread pairs one of them with the later ones of the same locus, merging the two reads' stored
// (locus, base) lists.
#include "common.cuh"

#include <algorithm>

namespace {

constexpr int TB = 256;

struct MultiArgs {
    uint64_t n_special;       // upper bound of the number of candidates (launch size)
    const uint64_t *n_me;     // device: number of candidates
    const uint32_t *me_code;  // cell << 4 | base << 2 | tail << 1 | 1
    const uint32_t *me_locus;
    const uint32_t *me_pos;   // element of the read's stored list that belongs to this locus
    const uint32_t *me_beg;   // stored list [me_beg, me_end) in g_list / g_base
    const uint32_t *me_end;
    const uint32_t *g_list;   // stored loci, ascending
    const uint8_t *g_base;    // stored bases
    int32_t *H2;              // 3 planes
    int32_t *H3;              // 4 planes
    double *spill;            // may be null
    const double *G;          // device table SGPU_MAX_CLASS^2 (only read when spill != null)
    unsigned long long *hist; // SGPU_MAX_CLASS^2
    uint32_t n_cells;
    uint64_t nn;
    uint32_t L;
    int spill_only;           // second pass: record only the classes with x_s + x_d >= 4
    unsigned int *max_order;
    unsigned long long *n_pairs;
    int *err;
};

__global__ void __launch_bounds__(TB) multilocus_kernel(MultiArgs a) {
    // the class histogram has a handful of hot bins: count per block, flush once
    __shared__ uint32_t s_hist[SGPU_MAX_CLASS * SGPU_MAX_CLASS];
    __shared__ unsigned int s_max_order;
    const uint64_t ia = static_cast<uint64_t>(blockIdx.x) * TB + threadIdx.x;
    const uint64_t n_me = *a.n_me;
    if (static_cast<uint64_t>(blockIdx.x) * TB >= n_me) {
        return; // whole block behind the last candidate
    }
    for (uint32_t i = threadIdx.x; i < SGPU_MAX_CLASS * SGPU_MAX_CLASS; i += TB) {
        s_hist[i] = 0;
    }
    if (threadIdx.x == 0) {
        s_max_order = 0;
    }
    __syncthreads();
    unsigned long long local_pairs = 0;
    unsigned int local_max = 0;
    if (ia < n_me) {
        const uint32_t loc = a.me_locus[ia];
        const uint32_t ca = a.me_code[ia];
        const uint32_t a0 = a.me_beg[ia], ap = a.me_pos[ia], a1 = a.me_end[ia];
        for (uint64_t ib = ia + 1; ib < n_me; ++ib) {
            if (a.me_locus[ib] != loc) {
                break;
            }
            const uint32_t cb = a.me_code[ib];
            if (code_cell(ca) == code_cell(cb) || (ca & cb & 2u)) {
                continue; // same cell, or both reads behind the cutoff K
            }
            const uint32_t b0 = a.me_beg[ib], bp = a.me_pos[ib], b1 = a.me_end[ib];
            // two-pointer merge over the stored loci (similarity_matrix.cpp:221-229), the part behind this
            // locus first: most pairs share nothing there and are single-locus overlaps (first order)
            uint32_t xs = 0, xd = 0;
            for (uint32_t i = ap + 1, j = bp + 1; i < a1 && j < b1;) {
                const uint32_t la = a.g_list[i], lb = a.g_list[j];
                if (la == lb) {
                    a.g_base[i] == a.g_base[j] ? ++xs : ++xd;
                    ++i;
                    ++j;
                } else if (la < lb) {
                    ++i;
                } else {
                    ++j;
                }
            }
            if (xs + xd == 0) {
                continue;
            }
            bool earlier = false; // a common locus before this one: the pair is accounted there
            for (uint32_t i = a0, j = b0; i < ap && j < bp;) {
                const uint32_t la = a.g_list[i], lb = a.g_list[j];
                if (la == lb) {
                    earlier = true;
                    break;
                }
                la < lb ? ++i : ++j;
            }
            if (earlier) {
                continue;
            }
            a.g_base[ap] == a.g_base[bp] ? ++xs : ++xd; // this locus
            if (xs >= SGPU_MAX_CLASS || xd >= SGPU_MAX_CLASS || xs >= a.L || xd >= a.L) {
                atomicExch(a.err, SGPU_E_CLASS_RANGE);
                continue;
            }
            const uint32_t order = xs + xd;
            const uint32_t lo = min(code_cell(ca), code_cell(cb)), hi = max(code_cell(ca), code_cell(cb));
            const uint64_t ij = static_cast<uint64_t>(lo) * a.n_cells + hi;
            if (!a.spill_only) {
                ++local_pairs;
                atomicAdd(&s_hist[xs * SGPU_MAX_CLASS + xd], 1u);
                if (order == 2) {
                    atomicAdd(&a.H2[static_cast<uint64_t>(xd) * a.nn + ij], 1);
                } else if (order == 3) {
                    atomicAdd(&a.H3[static_cast<uint64_t>(xd) * a.nn + ij], 1);
                }
                local_max = max(local_max, order);
            }
            if (order >= 4 && a.spill != nullptr) {
                atomicAdd(&a.spill[ij], a.G[xs * SGPU_MAX_CLASS + xd]);
            }
        }
    }
    if (local_max) {
        atomicMax(&s_max_order, local_max);
    }
    __syncthreads();
    for (uint32_t i = threadIdx.x; i < SGPU_MAX_CLASS * SGPU_MAX_CLASS; i += TB) {
        if (s_hist[i]) {
            atomicAdd(&a.hist[i], static_cast<unsigned long long>(s_hist[i]));
        }
    }
    if (threadIdx.x == 0 && s_max_order) {
        atomicMax(a.max_order, s_max_order);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        local_pairs += __shfl_xor_sync(0xffffffffu, local_pairs, o);
    }
    if ((threadIdx.x & 31) == 0 && local_pairs) {
        atomicAdd(a.n_pairs, local_pairs);
    }
}

unsigned blocks_for(uint64_t n) { return static_cast<unsigned>(ceil_div_u64(n ? n : 1, TB)); }

} // namespace

int sgpu_multilocus(sgpu_ctx *ctx, const sgpu_pileup *p, const LinkResult &lr, sgpu_counts *c,
                    uint32_t L, uint64_t *n_pairs_multi) {
    cudaStream_t st = ctx->stream;
    if (n_pairs_multi) {
        *n_pairs_multi = 0;
    }
    if (lr.n_multi == 0 || p->n_entries == 0 || lr.n_special == 0) {
        return SGPU_OK;
    }
    const uint64_t NS = lr.n_special;
    DevBuf<unsigned long long> d_np;
    DevBuf<unsigned int> d_max;
    DevBuf<int> d_err;
    DevBuf<double> d_G;
    SGPU_CUDA(ctx, d_np.alloc(1, ctx));
    SGPU_CUDA(ctx, d_max.alloc(1, ctx));
    SGPU_CUDA(ctx, d_err.alloc(1, ctx));
    SGPU_CUDA(ctx, cudaMemsetAsync(d_np.p, 0, sizeof(unsigned long long), st));
    SGPU_CUDA(ctx, cudaMemsetAsync(d_max.p, 0, sizeof(unsigned int), st));
    SGPU_CUDA(ctx, cudaMemsetAsync(d_err.p, 0, sizeof(int), st));
    if (c->spill != nullptr) {
        SGPU_CUDA(ctx, d_G.alloc(SGPU_MAX_CLASS * SGPU_MAX_CLASS, ctx));
        SGPU_TRY(sgpu_build_gtable(ctx, c->eps, c->h, c->theta, c->L, SGPU_MAX_CLASS, d_G.p, nullptr));
    }
    MultiArgs a;
    a.n_special = NS;
    a.n_me = lr.me_idx.p + NS;
    a.me_code = lr.me_code.p;
    a.me_locus = lr.me_locus.p;
    a.me_pos = lr.me_pos.p;
    a.me_beg = lr.me_beg.p;
    a.me_end = lr.me_end.p;
    a.g_list = lr.g_list.p;
    a.g_base = lr.g_base.p;
    a.H2 = c->i32 + PLANE_H2 * c->nn;
    a.H3 = c->i32 + PLANE_H3 * c->nn;
    a.spill = c->spill;
    a.G = d_G.p;
    a.hist = reinterpret_cast<unsigned long long *>(c->hist);
    a.n_cells = c->n;
    a.nn = c->nn;
    a.L = L;
    a.spill_only = 0;
    a.max_order = d_max.p;
    a.n_pairs = d_np.p;
    a.err = d_err.p;
    SGPU_LAUNCH(ctx, (multilocus_kernel<<<blocks_for(NS), TB, 0, st>>>(a)));
    SGPU_CUDA(ctx, cudaGetLastError());
    SGPU_CUDA(ctx, cudaMemcpyAsync(&ctx->h_scratch[0], d_np.p, sizeof(uint64_t), cudaMemcpyDeviceToHost, st));
    SGPU_CUDA(ctx, cudaMemcpyAsync(&ctx->h_scratch[1], d_max.p, sizeof(unsigned int), cudaMemcpyDeviceToHost, st));
    SGPU_CUDA(ctx, cudaMemcpyAsync(&ctx->h_scratch[2], d_err.p, sizeof(int), cudaMemcpyDeviceToHost, st));
    SGPU_CUDA(ctx, cudaStreamSynchronize(st));
    if (static_cast<int>(ctx->h_scratch[2] & 0xFFFFFFFFu) != 0) {
        return sgpu_fail(ctx, SGPU_E_CLASS_RANGE, "a read pair overlaps at >= %d loci (or >= max_fragment_length)", SGPU_MAX_CLASS);
    }
    if (n_pairs_multi) {
        *n_pairs_multi = ctx->h_scratch[0];
    }
    const unsigned int max_order = static_cast<unsigned int>(ctx->h_scratch[1] & 0xFFFFFFFFu);
    // planes that can be non-zero from now on: (2,0) (1,1) (0,2), and (3,0) .. (0,3) once an overlap of
    // three loci was seen (zeroing, the multi-GPU reduction and the epilogue skip the others)
    if (max_order >= 2) {
        c->planes_used = std::max(c->planes_used, max_order >= 3 ? static_cast<int>(N_PLANES) : static_cast<int>(PLANE_H3));
    }
    if (max_order >= 4 && c->spill == nullptr) {
        // first pair of order >= 4 seen by this counts object: create the spill plane and record
        // just those classes in a second pass
        SGPU_CUDA(ctx, cudaMalloc(&c->spill, c->nn * sizeof(double)));
        SGPU_CUDA(ctx, cudaMemsetAsync(c->spill, 0, c->nn * sizeof(double), st));
        SGPU_CUDA(ctx, d_G.alloc(SGPU_MAX_CLASS * SGPU_MAX_CLASS, ctx));
        SGPU_TRY(sgpu_build_gtable(ctx, c->eps, c->h, c->theta, c->L, SGPU_MAX_CLASS, d_G.p, nullptr));
        a.spill = c->spill;
        a.G = d_G.p;
        a.spill_only = 1;
        SGPU_LAUNCH(ctx, (multilocus_kernel<<<blocks_for(NS), TB, 0, st>>>(a)));
        SGPU_CUDA(ctx, cudaGetLastError());
        SGPU_CUDA(ctx, cudaStreamSynchronize(st));
    }
    return SGPU_OK;
}
