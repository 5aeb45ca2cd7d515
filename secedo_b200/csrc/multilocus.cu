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
// Only entries of reads that keep >= 2 loci take part (a small minority of the pileup): they are
// compacted into a list ordered by locus, and each thread pairs one of them with the later ones of
// the same locus, merging the two reads' stored (locus, base) lists.
#include "common.cuh"

namespace {

constexpr int TB = 256;

__global__ void __launch_bounds__(TB) multi_entry_flag_kernel(const uint32_t *__restrict__ code, uint64_t n_entries,
                                                              uint8_t *__restrict__ flag) {
    const uint64_t e = static_cast<uint64_t>(blockIdx.x) * TB + threadIdx.x;
    if (e < n_entries) {
        const uint32_t c = code[e];
        flag[e] = (c != CODE_DROPPED && (c & 1u)) ? 1 : 0;
    }
}

__global__ void __launch_bounds__(TB) multi_entry_compact_kernel(const uint8_t *__restrict__ flag,
                                                                 const uint64_t *__restrict__ idx, uint64_t n_entries,
                                                                 uint32_t *__restrict__ me) {
    const uint64_t e = static_cast<uint64_t>(blockIdx.x) * TB + threadIdx.x;
    if (e < n_entries && flag[e]) {
        me[idx[e]] = static_cast<uint32_t>(e);
    }
}

struct MultiArgs {
    const uint32_t *me;       // entries of multi-locus reads, ascending (hence grouped by locus)
    uint64_t n_me;
    const uint32_t *code;
    const uint32_t *eread;
    const uint32_t *eloc;
    const uint32_t *r_multi;
    const uint64_t *m_off;
    const uint32_t *m_locus;
    const uint8_t *m_base;
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
    const uint64_t ia = static_cast<uint64_t>(blockIdx.x) * TB + threadIdx.x;
    unsigned long long local_pairs = 0;
    if (ia < a.n_me) {
        const uint32_t ea = a.me[ia];
        const uint32_t loc = a.eloc[ea];
        const uint32_t ca = a.code[ea];
        const uint32_t ma = a.r_multi[a.eread[ea]];
        const uint64_t a0 = a.m_off[ma], a1 = a.m_off[ma + 1];
        for (uint64_t ib = ia + 1; ib < a.n_me; ++ib) {
            const uint32_t eb = a.me[ib];
            if (a.eloc[eb] != loc) {
                break;
            }
            const uint32_t cb = a.code[eb];
            if (code_cell(ca) == code_cell(cb) || (ca & cb & 2u)) {
                continue; // same cell, or both reads behind the cutoff K
            }
            const uint32_t mb = a.r_multi[a.eread[eb]];
            const uint64_t b0 = a.m_off[mb], b1 = a.m_off[mb + 1];
            // two-pointer merge over the stored loci (similarity_matrix.cpp:221-229)
            uint32_t xs = 0, xd = 0, first_common = 0xFFFFFFFFu;
            for (uint64_t i = a0, j = b0; i < a1 && j < b1;) {
                const uint32_t la = a.m_locus[i], lb = a.m_locus[j];
                if (la == lb) {
                    if (first_common == 0xFFFFFFFFu) {
                        first_common = la;
                    }
                    a.m_base[i] == a.m_base[j] ? ++xs : ++xd;
                    ++i;
                    ++j;
                } else if (la < lb) {
                    ++i;
                } else {
                    ++j;
                }
            }
            if (first_common != loc || xs + xd < 2) {
                continue; // accounted at an earlier shared locus, or a single-locus overlap
            }
            if (xs >= SGPU_MAX_CLASS || xd >= SGPU_MAX_CLASS || xs >= a.L || xd >= a.L) {
                atomicExch(a.err, SGPU_E_CLASS_RANGE);
                continue;
            }
            const uint32_t order = xs + xd;
            const uint32_t lo = min(code_cell(ca), code_cell(cb)), hi = max(code_cell(ca), code_cell(cb));
            const uint64_t ij = static_cast<uint64_t>(lo) * a.n_cells + hi;
            if (!a.spill_only) {
                ++local_pairs;
                atomicAdd(&a.hist[xs * SGPU_MAX_CLASS + xd], 1ull);
                if (order == 2) {
                    atomicAdd(&a.H2[static_cast<uint64_t>(xd) * a.nn + ij], 1);
                } else if (order == 3) {
                    atomicAdd(&a.H3[static_cast<uint64_t>(xd) * a.nn + ij], 1);
                } else {
                    atomicMax(a.max_order, order);
                }
            }
            if (order >= 4 && a.spill != nullptr) {
                atomicAdd(&a.spill[ij], a.G[xs * SGPU_MAX_CLASS + xd]);
            }
        }
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
    const uint64_t E = p->n_entries;
    if (lr.n_multi == 0 || E == 0) {
        return SGPU_OK;
    }
    DevBuf<uint8_t> flag;
    DevBuf<uint64_t> idx;
    DevBuf<uint32_t> me;
    DevBuf<unsigned long long> d_np;
    DevBuf<unsigned int> d_max;
    DevBuf<int> d_err;
    DevBuf<double> d_G;
    SGPU_CUDA(ctx, flag.alloc(E, st));
    SGPU_CUDA(ctx, idx.alloc(E + 1, st));
    SGPU_LAUNCH(ctx, (multi_entry_flag_kernel<<<blocks_for(E), TB, 0, st>>>(lr.code.p, E, flag.p)));
    SGPU_TRY(sgpu_scan_u8_u64(ctx, flag.p, idx.p, E));
    SGPU_CUDA(ctx, cudaMemcpyAsync(&ctx->h_scratch[0], idx.p + E, sizeof(uint64_t), cudaMemcpyDeviceToHost, st));
    SGPU_CUDA(ctx, cudaStreamSynchronize(st));
    const uint64_t NME = ctx->h_scratch[0];
    if (NME == 0) {
        return SGPU_OK;
    }
    SGPU_CUDA(ctx, me.alloc(NME, st));
    SGPU_LAUNCH(ctx, (multi_entry_compact_kernel<<<blocks_for(E), TB, 0, st>>>(flag.p, idx.p, E, me.p)));
    SGPU_CUDA(ctx, d_np.alloc(1, st));
    SGPU_CUDA(ctx, d_max.alloc(1, st));
    SGPU_CUDA(ctx, d_err.alloc(1, st));
    SGPU_CUDA(ctx, cudaMemsetAsync(d_np.p, 0, sizeof(unsigned long long), st));
    SGPU_CUDA(ctx, cudaMemsetAsync(d_max.p, 0, sizeof(unsigned int), st));
    SGPU_CUDA(ctx, cudaMemsetAsync(d_err.p, 0, sizeof(int), st));
    if (c->spill != nullptr) {
        SGPU_CUDA(ctx, d_G.alloc(SGPU_MAX_CLASS * SGPU_MAX_CLASS, st));
        SGPU_TRY(sgpu_build_gtable(ctx, c->eps, c->h, c->theta, c->L, SGPU_MAX_CLASS, d_G.p, nullptr));
    }
    c->planes_used = N_PLANES;

    MultiArgs a;
    a.me = me.p;
    a.n_me = NME;
    a.code = lr.code.p;
    a.eread = lr.eread.p;
    a.eloc = lr.eloc.p;
    a.r_multi = lr.r_multi.p;
    a.m_off = lr.m_off.p;
    a.m_locus = lr.m_locus.p;
    a.m_base = lr.m_base.p;
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
    SGPU_LAUNCH(ctx, (multilocus_kernel<<<blocks_for(NME), TB, 0, st>>>(a)));
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
    if (max_order >= 4 && c->spill == nullptr) {
        // first pair of order >= 4 seen by this counts object: create the spill plane and record
        // just those classes in a second pass
        SGPU_CUDA(ctx, cudaMalloc(&c->spill, c->nn * sizeof(double)));
        SGPU_CUDA(ctx, cudaMemsetAsync(c->spill, 0, c->nn * sizeof(double), st));
        SGPU_CUDA(ctx, d_G.alloc(SGPU_MAX_CLASS * SGPU_MAX_CLASS, st));
        SGPU_TRY(sgpu_build_gtable(ctx, c->eps, c->h, c->theta, c->L, SGPU_MAX_CLASS, d_G.p, nullptr));
        a.spill = c->spill;
        a.G = d_G.p;
        a.spill_only = 1;
        SGPU_LAUNCH(ctx, (multilocus_kernel<<<blocks_for(NME), TB, 0, st>>>(a)));
        SGPU_CUDA(ctx, cudaGetLastError());
        SGPU_CUDA(ctx, cudaStreamSynchronize(st));
    }
    return SGPU_OK;
}
