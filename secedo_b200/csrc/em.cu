// Expectation-maximisation refinement of a two-way split.
//
// Reference: expectation_maximization.cpp:131-160 (driver), cluster_center :19-40, maximization_step :59-92,
// expectation_step :109-129; called from divide_cluster, spectral_clustering.cpp:375-377, on the FILTERED pileup that
// already lives in HBM here. Per iteration the reference makes one pass over all entries: per locus the weighted base
// composition of both clusters ("centres", 2 x 4 sums over the locus' entries of the membership probability of the
// entry's cell), its clamped / renormalised logarithm, and then per entry ll_a[cell] += centre_a[base],
// ll_b[cell] += centre_b[base]; the E step is O(cells).
//
// Here: one WARP per locus (loci of the path have ~10^2..10^4 entries), 8 warps per CTA, grid-stride over the loci.
// The per-cell log likelihoods of one iteration are accumulated as 64-bit FIXED POINT (scale 2^32) with global
// integer reductions (RED.ADD.64, resolved in L2: 8 bytes per cell, always resident; only ll_b - ll_a is needed): integer addition is order
// independent, so the result is bit-reproducible, and no shared memory is needed, which leaves room for 40 warps per
// SM to hide the two dependent gathers per entry (probability of the entry's cell, position of its cell). A first
// version with per-CTA shared-memory accumulators was 10 x slower: 64-bit shared-memory atomics compile to a
// compare-and-swap loop (ATOMS.CAST.SPIN) and 128 KB of accumulators allow only 8 warps per SM. Quantisation 2^-33 per
// term, i.e. below the rounding of the reference's own running fp64 sums (ulp of ~10^7 is 2 10^-9); |term| < 2^6 and
// up to 2^25 entries of one cell per iteration fit in 63 bits. The E-step kernel converts and adds to the running
// log likelihoods. Every entry is read twice by the same warp (second time from L1/L2).
// Quirks kept: the probability is looked up with the GROUP id, the likelihood is accumulated at id_to_pos[group]
// (:24 vs :88-89); the log likelihoods are not reset between iterations (:136-137); sum == 0 -> log(1/4) (:29-31).
#include "common.cuh"

#include <cmath>
#include <cstring>

namespace {

constexpr int EM_WARPS = 8;
constexpr int EM_THREADS = EM_WARPS * 32;
constexpr double EM_FIX = 4294967296.0; // 2^32
constexpr int EM_ILP = 4;               // entries per lane in flight
constexpr int EM_REPLICAS = 32;         // copies of the per-cell accumulators: 8000 cells are only 500 L2 lines, and every
                                        // line serialises its reductions (integer sums: any split gives the same total)

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        v += __shfl_xor_sync(0xFFFFFFFFu, v, o);
    }
    return v;
}

// acc: [EM_REPLICAS][n_cells] int64, zero at launch: sum over the cell's entries of centre_b[base] - centre_a[base]. The E step
// only ever uses ll_b - ll_a (:120), so ONE reduction per entry carries everything
__global__ void __launch_bounds__(EM_THREADS)
em_mstep_kernel(const uint64_t *__restrict__ row_ptr, const uint16_t *__restrict__ gid_base, uint64_t n_loci,
                const double *__restrict__ prob_b, const uint32_t *__restrict__ id_to_pos, uint32_t n_cells, double theta,
                unsigned long long *__restrict__ acc) {
    const int lane = threadIdx.x & 31;
    const uint64_t warp0 = (static_cast<uint64_t>(blockIdx.x) * EM_THREADS + threadIdx.x) >> 5;
    acc += (warp0 % EM_REPLICAS) * n_cells;
    const uint64_t n_warps = (static_cast<uint64_t>(gridDim.x) * EM_THREADS) >> 5;
    for (uint64_t l = warp0; l < n_loci; l += n_warps) {
        const uint64_t e0 = row_ptr[l], e1 = row_ptr[l + 1];
        // weighted base counts of both clusters (a with 1 - p like the reference, :66-69)
        double sa[4] = { 0, 0, 0, 0 }, sb[4] = { 0, 0, 0, 0 };
        // four entries per lane in flight: the kernel is bound by the latency of the two dependent loads per entry
        // (ncu: 50 long-scoreboard stall cycles per issue with one entry per lane), not by bandwidth
        for (uint64_t e = e0 + lane; e < e1; e += 32 * EM_ILP) {
            uint32_t gb[EM_ILP];
            double pb[EM_ILP];
#pragma unroll
            for (int u = 0; u < EM_ILP; ++u) {
                gb[u] = e + 32 * u < e1 ? gid_base[e + 32 * u] : 0xFFFFFFFFu;
            }
#pragma unroll
            for (int u = 0; u < EM_ILP; ++u) {
                pb[u] = gb[u] != 0xFFFFFFFFu ? prob_b[gb[u] >> 2] : 0.0;
            }
#pragma unroll
            for (int u = 0; u < EM_ILP; ++u) { // same per-lane order of additions as one entry at a time
                if (gb[u] != 0xFFFFFFFFu) {
                    const double pa = 1.0 - pb[u];
#pragma unroll
                    for (int b = 0; b < 4; ++b) {
                        const bool hit = (gb[u] & 3u) == static_cast<uint32_t>(b);
                        sa[b] += hit ? pa : 0.0;
                        sb[b] += hit ? pb[u] : 0.0;
                    }
                }
            }
        }
        double ca[4], cb[4];
#pragma unroll
        for (int b = 0; b < 4; ++b) {
            ca[b] = warp_sum(sa[b]);
            cb[b] = warp_sum(sb[b]);
        }
        // lanes 0..3: centre a of base `lane`; lanes 4..7: centre b (cluster_center :27-38)
        const bool is_b = (lane & 4) != 0;
        const double c0 = is_b ? cb[0] : ca[0], c1 = is_b ? cb[1] : ca[1], c2 = is_b ? cb[2] : ca[2], c3 = is_b ? cb[3] : ca[3];
        const double s = c0 + c1 + c2 + c3;
        double lg;
        if (s == 0.0) {
            lg = log(0.25);
        } else {
            const double n0 = c0 / s > theta ? c0 / s : theta, n1 = c1 / s > theta ? c1 / s : theta;
            const double n2 = c2 / s > theta ? c2 / s : theta, n3 = c3 / s > theta ? c3 / s : theta;
            const double s2 = n0 + n1 + n2 + n3;
            const int bsel = lane & 3;
            const double v = bsel == 0 ? n0 : bsel == 1 ? n1 : bsel == 2 ? n2 : n3;
            lg = log(v / s2);
        }
        const long long fx = __double2ll_rn(lg * EM_FIX);
        long long td[4];
#pragma unroll
        for (int b = 0; b < 4; ++b) {
            td[b] = __shfl_sync(0xFFFFFFFFu, fx, 4 + b) - __shfl_sync(0xFFFFFFFFu, fx, b);
        }
        for (uint64_t e = e0 + lane; e < e1; e += 32 * EM_ILP) {
            uint32_t gb[EM_ILP], cell[EM_ILP];
#pragma unroll
            for (int u = 0; u < EM_ILP; ++u) {
                gb[u] = e + 32 * u < e1 ? gid_base[e + 32 * u] : 0xFFFFFFFFu;
            }
#pragma unroll
            for (int u = 0; u < EM_ILP; ++u) {
                cell[u] = gb[u] != 0xFFFFFFFFu ? id_to_pos[gb[u] >> 2] : 0u;
            }
#pragma unroll
            for (int u = 0; u < EM_ILP; ++u) {
                if (gb[u] != 0xFFFFFFFFu) {
                    const int b = gb[u] & 3;
                    const long long v = b == 0 ? td[0] : b == 1 ? td[1] : b == 2 ? td[2] : td[3];
                    atomicAdd(acc + cell[u], static_cast<unsigned long long>(v));
                }
            }
        }
    }
}

// ll += this iteration's fixed-point sums (which are cleared for the next one); expectation_step (:109-129) by ONE
// CTA of 1024 threads
__global__ void __launch_bounds__(1024) em_estep_kernel(unsigned long long *__restrict__ acc, uint32_t n_cells,
                                                        double *__restrict__ ll /* [n_cells]: ll_b - ll_a */, double *__restrict__ prob_b,
                                                        uint32_t *__restrict__ done_flag) {
    __shared__ double red[32];
    __shared__ int all_done;
    double psum = 0.0;
    for (uint32_t i = threadIdx.x; i < n_cells; i += 1024) {
        long long sum = 0;
        for (int r = 0; r < EM_REPLICAS; ++r) {
            sum += static_cast<long long>(acc[static_cast<size_t>(r) * n_cells + i]);
            acc[static_cast<size_t>(r) * n_cells + i] = 0ull;
        }
        ll[i] += static_cast<double>(sum) * (1.0 / EM_FIX);
        psum += prob_b[i];
    }
    psum = warp_sum(psum);
    if ((threadIdx.x & 31) == 0) {
        red[threadIdx.x >> 5] = psum;
    }
    if (threadIdx.x == 0) {
        all_done = 1;
    }
    __syncthreads();
    double total = 0.0;
    for (int i = 0; i < 32; ++i) {
        total += red[i];
    }
    const double prior_b = total / n_cells, prior_a = 1.0 - prior_b;
    bool done = true;
    for (uint32_t i = threadIdx.x; i < n_cells; i += 1024) {
        double d = ll[i];
        d = d < -100.0 ? -100.0 : (d > 100.0 ? 100.0 : d);
        const double odds = exp(d);
        const double p = 1.0 - 1.0 / (1.0 + odds * prior_b / prior_a);
        done = done && fabs(p - prob_b[i]) < 1e-2;
        prob_b[i] = p;
    }
    if (!done) {
        all_done = 0; // benign race: every writer stores 0
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        *done_flag = static_cast<uint32_t>(all_done);
    }
}

// max group id and max id_to_pos over the entries (range check: the reference reads out of range otherwise)
__global__ void em_range_kernel(const uint16_t *__restrict__ gid_base, uint64_t n_entries, const uint32_t *__restrict__ id_to_pos,
                                uint32_t n_groups, uint32_t n_cells, uint32_t *__restrict__ bad) {
    const uint64_t stride = static_cast<uint64_t>(gridDim.x) * blockDim.x;
    bool b = false;
    for (uint64_t e = static_cast<uint64_t>(blockIdx.x) * blockDim.x + threadIdx.x; e < n_entries; e += stride) {
        const uint32_t g = gid_base[e] >> 2;
        b = b || g >= n_cells || g >= n_groups || id_to_pos[g] >= n_cells;
    }
    if (b) {
        *bad = 1u;
    }
}

} // namespace

int sgpu_em_impl(sgpu_ctx *ctx, const sgpu_pileup *p, const uint32_t *h_id_to_pos, uint32_t n_groups, double theta,
                 double *h_prob_b, uint32_t n_cells, uint32_t max_iterations, uint32_t *iterations, float *ms_total) {
    cudaStream_t st = ctx->stream;
    if (n_cells == 0 || !h_prob_b) {
        return sgpu_fail(ctx, SGPU_E_ARG, "expectation_maximization: empty probability vector");
    }
    SGPU_WAIT_PILEUP(ctx, p);
    struct Events { // destroyed on every return path
        cudaEvent_t a = nullptr, b = nullptr;
        ~Events() {
            if (a) {
                cudaEventDestroy(a);
            }
            if (b) {
                cudaEventDestroy(b);
            }
        }
    } evs;
    SGPU_CUDA(ctx, cudaEventCreate(&evs.a));
    SGPU_CUDA(ctx, cudaEventCreate(&evs.b));
    const cudaEvent_t t0 = evs.a, t1 = evs.b;
    SGPU_CUDA(ctx, cudaEventRecord(t0, st));
    DevBuf<double> d_prob, d_ll;
    DevBuf<uint32_t> d_map, d_flags;
    const uint32_t map_n = std::max(n_groups, 1u);
    SGPU_CUDA(ctx, d_prob.alloc(n_cells, ctx));
    SGPU_CUDA(ctx, d_ll.alloc(n_cells, ctx));
    SGPU_CUDA(ctx, d_map.alloc(map_n, ctx));
    SGPU_CUDA(ctx, d_flags.alloc(2, ctx));
    SGPU_CUDA(ctx, cudaMemcpyAsync(d_prob.p, h_prob_b, n_cells * sizeof(double), cudaMemcpyHostToDevice, st));
    if (n_groups) {
        SGPU_CUDA(ctx, cudaMemcpyAsync(d_map.p, h_id_to_pos, n_groups * sizeof(uint32_t), cudaMemcpyHostToDevice, st));
    }
    SGPU_CUDA(ctx, cudaMemsetAsync(d_ll.p, 0, n_cells * sizeof(double), st));
    SGPU_CUDA(ctx, cudaMemsetAsync(d_flags.p, 0, 2 * sizeof(uint32_t), st));
    uint32_t h_flags[2] = { 0, 0 };
    if (p->n_entries) {
        SGPU_LAUNCH(ctx, (em_range_kernel<<<4 * ctx->sm_count, 256, 0, st>>>(p->d_gid_base, p->n_entries, d_map.p, n_groups, n_cells,
                                                                             d_flags.p + 1)));
        SGPU_CUDA(ctx, cudaMemcpyAsync(h_flags, d_flags.p, sizeof(h_flags), cudaMemcpyDeviceToHost, st));
        SGPU_CUDA(ctx, cudaStreamSynchronize(st));
        if (h_flags[1]) {
            // prob_cluster[group_id] / log_likelihood.at(id_to_pos[group_id]) out of range in the reference
            return sgpu_fail(ctx, SGPU_E_CELL_RANGE, "expectation_maximization: a group id or id_to_pos[group id] is >= %u cells", n_cells);
        }
    }
    DevBuf<unsigned long long> d_acc;
    SGPU_CUDA(ctx, d_acc.alloc(static_cast<size_t>(EM_REPLICAS) * n_cells, ctx));
    SGPU_CUDA(ctx, cudaMemsetAsync(d_acc.p, 0, static_cast<size_t>(EM_REPLICAS) * n_cells * sizeof(unsigned long long), st));
    const uint32_t n_cta = static_cast<uint32_t>(std::min<uint64_t>(5ull * ctx->sm_count, (p->n_loci + EM_WARPS - 1) / EM_WARPS));
    uint32_t it = 0;
    for (;;) {
        if (n_cta) {
            SGPU_LAUNCH(ctx, (em_mstep_kernel<<<n_cta, EM_THREADS, 0, st>>>(p->d_row_ptr, p->d_gid_base, p->n_loci, d_prob.p, d_map.p, n_cells,
                                                                           theta, d_acc.p)));
        }
        SGPU_LAUNCH(ctx, (em_estep_kernel<<<1, 1024, 0, st>>>(d_acc.p, n_cells, d_ll.p, d_prob.p, d_flags.p)));
        SGPU_CUDA(ctx, cudaGetLastError());
        SGPU_CUDA(ctx, cudaMemcpyAsync(h_flags, d_flags.p, sizeof(uint32_t), cudaMemcpyDeviceToHost, st));
        SGPU_CUDA(ctx, cudaStreamSynchronize(st));
        ++it;
        if (h_flags[0] || (max_iterations && it >= max_iterations)) {
            break;
        }
    }
    SGPU_CUDA(ctx, cudaMemcpyAsync(h_prob_b, d_prob.p, n_cells * sizeof(double), cudaMemcpyDeviceToHost, st));
    SGPU_CUDA(ctx, cudaEventRecord(t1, st));
    SGPU_CUDA(ctx, cudaStreamSynchronize(st));
    if (iterations) {
        *iterations = it;
    }
    if (ms_total) {
        SGPU_CUDA(ctx, cudaEventElapsedTime(ms_total, t0, t1));
    }
    return SGPU_OK;
}
