// K1 — the per-locus significance filter (reference: util/is_significant.cpp:78-193).
//
//   filter_count_kernel    one warp per locus: pooled A/C/G/T counts over the entries whose group is
//                          in the current sub-cluster, then the Bayesian test in fp64 (lane 0)
//   filter_compact_kernel  one warp per kept locus: ballot compaction of the in-cluster entries
//                          into the filtered CSR (order preserved, as Filter::filter does)
// Both are pure streaming kernels bounded by HBM: 2 B/entry read by the first, 6 B/entry read +
// 6 B/kept entry written by the second.
#include "common.cuh"

#include <cmath>
#include <algorithm>
#include <vector>

namespace {

// util/is_significant.cpp:11-36
__constant__ double c_Ks[5][20] = {
    { -1.64504967001201, -1.38868450353301, -1.38780664765677, -1.38779600211955, -1.3877952855556,
      -1.38779524274215, -1.38779524274142, -1.38779524274141, -1.3877952427414, -1.38779524274139,
      -1.38779524274138, -1.38779524274138, -1.38779524274138, -1.38779524274138, -1.38779524274139,
      -1.38780870444455, -1.38780870444455, -1.38780870444455, -1.38780870444455, -1.38780870444455 },
    { -1.56013904495168, -1.38819451352203, -1.38781438946096, -1.38779659244035, -1.38779537799054,
      -1.3877952484612, -1.3877952427842, -1.38779524274906, -1.3877952427457, -1.38779524274275,
      -1.38780870444458, -1.38780870444459, -1.38780870444459, -1.38780870444469, -1.38780870444459,
      -1.42736056742577, -1.42736056742575, -6.19144172018466, -6.19144172018466, -14.1885779508362 },
    { -1.47780038365618, -1.3885722463397, -1.38781428162649, -1.3877984410546, -1.38779548312685,
      -1.3877952855556, -1.38779524455204, -1.38779524331456, -1.38780873675669, -1.38780870687333,
      -1.42737804009806, -6.19144172131432, -14.1885779508648, -6.1914418045659, -30.1993093269287,
      -30.1993093268559, -30.1993093268539, -54.2154105288032, -62.2207775961199, -46.2100434614866 },
    { -1.47780038365618, -1.38868450353301, -1.38782829051844, -1.3877984410546, -1.38779625512927,
      -1.38779556321717, -1.38780972304588, -1.3878087226245, -6.21747860711653, -22.1939432034943,
      -14.1886670526002, -22.1939422903721, -46.2100434614866, -54.2154105288069, -70.2261446634366,
      -62.2207775961199, -86.2368787980699, -110.25298000002, -118.258347067337, -102.247612932703 },
    { -1.52859626647315, -1.38967447346712, -1.38787138908447, -1.38780282263764, -1.387805349423,
      -1.38882047800373, -1.49793700616569, -6.19975747800726, -22.197881249831, -38.2046765807324,
      -38.2046769835162, -70.2261446634383, -54.2154105303641, -78.2315117307532, -86.2368787980699,
      -118.258347067337, -126.263714134653, -134.26908120197, -158.28518240392, -158.28518240392 }
};

struct FilterParams {
    double theta;
    double log_theta_3;         // log(theta/3)
    double log_one_minus_theta; // log(1-theta)
    double log_1_4;             // log(1/4)
    double log_homo_prior;      // log(hetero_prior): quirk of util/is_significant.cpp:45, reproduced
    int cell_proportion;
};

// The significance test of util/is_significant.cpp:78-138 on the pooled counts of one locus (uint16 like the reference's
// base_count, which wraps). Four integer gates on the sorted counts, then a likelihood ratio in fp64: the log-probability
// of "every read shows the majority base" against the log of an evidence that is the sum of five genotype models, each a
// prior times powers of per-read probabilities. The products and the sum are formed in the reference's order (fp64
// multiplication and addition are not associative; the decision must be bit-identical).
struct GenotypeModel { // prior * x^nx * y^ny * z^nz
    double prior, x, y, z;
    uint32_t nx, ny, nz;
};
__device__ __forceinline__ double model_probability(const GenotypeModel &m) {
    double p = m.prior * pow(m.x, static_cast<double>(m.nx));
    if (m.y >= 0) {
        p = p * pow(m.y, static_cast<double>(m.ny));
    }
    if (m.z >= 0) {
        p = p * pow(m.z, static_cast<double>(m.nz));
    }
    return p;
}

__device__ bool is_significant_dev(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, const FilterParams &fp) {
    // ascending: n[3] = majority base, n[2] = runner-up (5-comparator sorting network)
    uint32_t n[4] = { c0 & 0xFFFFu, c1 & 0xFFFFu, c2 & 0xFFFFu, c3 & 0xFFFFu };
    const uint32_t depth = n[0] + n[1] + n[2] + n[3];
    if (depth < 2) {
        return false;
    }
    auto order = [&](int i, int j) {
        const uint32_t lo = min(n[i], n[j]), hi = max(n[i], n[j]);
        n[i] = lo;
        n[j] = hi;
    };
    order(0, 1);
    order(2, 3);
    order(0, 2);
    order(1, 3);
    order(1, 2);
    const uint32_t major = n[3], second = n[2], rest = n[0] + n[1];
    // gates (:92-103): a second base must exist, >= 5 non-majority reads, and the majority must lead 3 : 2
    if (second == 0 || second + rest < 5 || static_cast<double>(major) < 1.5 * static_cast<double>(second)) {
        return false;
    }
    // threshold column for the nearest multiple of 10 (round half to even, :67-70,106-107), clamped to the table
    const double column = fmin(fmax(rint(depth / 10.) - 1, 0.), 19.);

    const double theta = fp.theta, err3 = theta / 3;
    const double prior_het = 0.0005, prior_mut = 1e-6, prior_hom = 1 - prior_het - prior_mut;
    // log P(all reads from one homozygous genotype): the constants come from the host (Filter's constructor)
    double log_homozygous = major * fp.log_one_minus_theta + (depth - major) * fp.log_theta_3;
    log_homozygous += fp.log_1_4;
    log_homozygous += fp.log_homo_prior;
    // the five models of the evidence, summed in this order (:112-136); y / z < 0 marks an unused factor
    const GenotypeModel models[5] = {
        { prior_hom, 1 - theta, err3, -1., major, depth - major, 0 },                                  // homozygous, errors only
        { prior_het, 0.5 - err3, err3, -1., major + second, rest, 0 },                                 // germline heterozygous
        { prior_hom * prior_mut, 0.75 - 2 * theta / 3, 0.25, err3, major, second, rest },               // homozygous + somatic
        { prior_het * prior_mut, 0.5 - theta, 0.25, err3, major, n[1] + second, n[0] },                 // heterozygous + somatic
        { prior_het * prior_mut * prior_mut, 1 - theta, -1., -1., depth, 0, 0 },                        // two somatic mutations
    };
    double evidence = model_probability(models[0]);
#pragma unroll
    for (int k = 1; k < 5; ++k) {
        evidence = evidence + model_probability(models[k]);
    }
    return log_homozygous - log(evidence) < c_Ks[fp.cell_proportion][static_cast<uint32_t>(column)];
}

constexpr int FILTER_THREADS = 256;

// One WARP per locus (grid-stride over warps), 8 entries per 16-byte load and four loads in flight per lane (2 KB per
// warp, 128 KB per SM): no CTA-wide barrier per locus, so the load latency of one locus is hidden by the other warps
// (the CTA-per-locus version paid a barrier pair and an exposed load round trip per locus: 30 % of HBM).
template <typename GB>
__global__ void __launch_bounds__(FILTER_THREADS) filter_count_kernel(
        const uint64_t *__restrict__ row_ptr, const GB *__restrict__ gid_base, uint64_t n_loci,
        const uint32_t *__restrict__ in_mask /* bit per group id */, uint32_t n_groups,
        uint4 *__restrict__ counts /* pooled A, C, G, T per locus */, int *__restrict__ err) {
    extern __shared__ uint32_t s_mask[]; // one bit per group id, at least 512 words (all 14-bit ids); bits behind n_groups are 0
    constexpr int PER_VEC = 16 / sizeof(GB); // entries per 16-byte load
    const uint32_t mask_words = (n_groups + 31) / 32, smem_words = max(512u, mask_words);
    for (uint32_t i = threadIdx.x; i < smem_words; i += blockDim.x) {
        s_mask[i] = i < mask_words ? in_mask[i] : 0;
    }
    __syncthreads();
    const int lane = threadIdx.x & 31;
    const bool aligned = (reinterpret_cast<uintptr_t>(gid_base) & 15u) == 0;
    const uint64_t warps_total = static_cast<uint64_t>(gridDim.x) * (FILTER_THREADS / 32);
    uint32_t max_gid = 0;
    for (uint64_t l = static_cast<uint64_t>(blockIdx.x) * (FILTER_THREADS / 32) + (threadIdx.x >> 5); l < n_loci; l += warps_total) {
        const uint64_t e0 = row_ptr[l], e1 = row_ptr[l + 1];
        // per-lane counters: one byte per base packed in a word (one shift + add per entry), spilled into 32-bit
        // counters before a byte can overflow
        uint32_t c[4] = { 0, 0, 0, 0 };
        uint32_t packed = 0, since_flush = 0;
        auto flush = [&]() {
            c[0] += packed & 0xFFu;
            c[1] += (packed >> 8) & 0xFFu;
            c[2] += (packed >> 16) & 0xFFu;
            c[3] += packed >> 24;
            packed = 0;
            since_flush = 0;
        };
        auto count = [&](uint32_t gb) {
            uint32_t gid = gb >> 2;
            max_gid = max(max_gid, gid);
            if (sizeof(GB) == 4) {
                gid = min(gid, smem_words * 32 - 1); // a wide id behind the mask is an error (reported below), not a fault
            }
            // 14-bit ids: the shared-memory mask covers all 16 384 of them
            const uint32_t in = (s_mask[gid >> 5] >> (gid & 31)) & 1u;
            packed += in << (8u * (gb & 3u));
        };
        auto count8 = [&](const uint4 &q) {
            const uint32_t w[4] = { q.x, q.y, q.z, q.w };
            if (sizeof(GB) == 2) {
#pragma unroll
                for (int k = 0; k < 8; ++k) {
                    count((w[k >> 1] >> (16 * (k & 1))) & 0xFFFFu);
                }
            } else {
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    count(w[k]);
                }
            }
        };
        // 16-byte aligned middle part (none if the caller's array itself is not aligned)
        const uint64_t up = (e0 + PER_VEC - 1) & ~static_cast<uint64_t>(PER_VEC - 1), down = e1 & ~static_cast<uint64_t>(PER_VEC - 1);
        const uint64_t a0 = up < e1 ? up : e1, a1 = (aligned && down > a0) ? down : a0;
        for (uint64_t e = e0 + lane; e < a0; e += 32) { // < 8 entries
            count(gid_base[e]);
        }
        const uint4 *vec = reinterpret_cast<const uint4 *>(gid_base + a0);
        const uint64_t nv = (a1 - a0) / PER_VEC;
        for (uint64_t v = lane; v < nv; v += 128) {
            uint4 q[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                q[u] = v + 32 * u < nv ? vec[v + 32 * u] : make_uint4(0, 0, 0, 0);
            }
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                if (v + 32 * u < nv) {
                    count8(q[u]);
                }
            }
            since_flush += 4 * PER_VEC;
            if (since_flush > 200) {
                flush();
            }
        }
        for (uint64_t e = a1 + lane; e < e1; e += 32) { // < 8 entries, or all of an unaligned array
            count(gid_base[e]);
            if (++since_flush > 200) {
                flush();
            }
        }
        flush();
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
#pragma unroll
            for (int b = 0; b < 4; ++b) {
                c[b] += __shfl_xor_sync(0xffffffffu, c[b], o);
            }
        }
        if (lane == 0) {
            counts[l] = make_uint4(c[0], c[1], c[2], c[3]);
        }
    }
    if (max_gid >= n_groups) {
        atomicExch(err, SGPU_E_CELL_RANGE);
    }
}

// the Bayesian test, one thread per locus (fp64 pow / log: kept out of the counting CTAs)
__global__ void __launch_bounds__(FILTER_THREADS) filter_decide_kernel(const uint4 *__restrict__ counts, uint64_t n_loci,
                                                                      FilterParams fp, uint8_t *__restrict__ keep,
                                                                      uint32_t *__restrict__ kept_cnt,
                                                                      unsigned int *__restrict__ max_kept) {
    const uint64_t l = static_cast<uint64_t>(blockIdx.x) * FILTER_THREADS + threadIdx.x;
    unsigned int n = 0;
    if (l < n_loci) {
        const uint4 c = counts[l];
        const bool sig = is_significant_dev(c.x, c.y, c.z, c.w, fp);
        n = sig ? c.x + c.y + c.z + c.w : 0;
        keep[l] = sig ? 1 : 0;
        kept_cnt[l] = n;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        n = max(n, __shfl_xor_sync(0xffffffffu, n, o));
    }
    if ((threadIdx.x & 31) == 0 && n) {
        atomicMax(max_kept, n);
    }
}

// Lazily uploaded pileup: the read ids of the KEPT loci are pulled from the caller's mapped host memory into a device
// array with the same indexing. One warp per kept locus, 16 bytes per lane and four loads in flight (2 KB per warp):
// narrow 4-byte-per-lane reads of host memory reached only 27 GB/s over PCIe.
__global__ void __launch_bounds__(FILTER_THREADS) pull_read_ids_kernel(const uint64_t *__restrict__ row_ptr, uint64_t n_loci,
                                                                       const uint8_t *__restrict__ keep,
                                                                       const uint32_t *__restrict__ src, uint32_t *__restrict__ dst) {
    const int lane = threadIdx.x & 31;
    const uint64_t warps_total = static_cast<uint64_t>(gridDim.x) * (FILTER_THREADS / 32);
    for (uint64_t l = static_cast<uint64_t>(blockIdx.x) * (FILTER_THREADS / 32) + (threadIdx.x >> 5); l < n_loci; l += warps_total) {
        if (!keep[l]) {
            continue;
        }
        const uint64_t e0 = row_ptr[l], e1 = row_ptr[l + 1];
        // elements up to the first 16-byte boundary of BOTH arrays (they share the alignment if the bases do; if not,
        // everything goes through the scalar loop)
        const uint64_t mis_s = (reinterpret_cast<uintptr_t>(src + e0) >> 2) & 3u, mis_d = (reinterpret_cast<uintptr_t>(dst + e0) >> 2) & 3u;
        uint64_t b0 = e0 + ((4 - mis_s) & 3u);
        uint64_t b1 = b0 + ((e1 > b0 ? e1 - b0 : 0) & ~3ull);
        if (mis_s != mis_d || b0 > e1) {
            b0 = b1 = e1;
        }
        for (uint64_t e = e0 + lane; e < min(b0, e1); e += 32) {
            dst[e] = src[e];
        }
        const uint4 *s4 = reinterpret_cast<const uint4 *>(src + b0);
        uint4 *d4 = reinterpret_cast<uint4 *>(dst + b0);
        const uint64_t n4 = (b1 - b0) / 4;
        for (uint64_t i = lane; i < n4; i += 128) {
            uint4 v[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                if (i + 32 * u < n4) {
                    v[u] = s4[i + 32 * u];
                }
            }
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                if (i + 32 * u < n4) {
                    d4[i + 32 * u] = v[u];
                }
            }
        }
        for (uint64_t e = b1 + lane; e < e1; e += 32) {
            dst[e] = src[e];
        }
    }
}

// WITH_IDS = false: every group belongs to the sub-cluster, so a kept locus keeps all its entries in place: the read ids are
// not copied (4 of the 6 bytes per entry), the filtered pileup reads them through a view of the source (view_off[locus] = the
// locus' first source entry, sgpu_pileup::view_read_id)
template <typename GB, bool WITH_IDS>
__global__ void __launch_bounds__(FILTER_THREADS) filter_compact_kernel(
        const uint64_t *__restrict__ row_ptr, const uint32_t *__restrict__ position,
        const uint32_t *__restrict__ read_id, const GB *__restrict__ gid_base, uint64_t n_loci,
        const uint32_t *__restrict__ in_mask, uint32_t n_groups, const uint8_t *__restrict__ keep,
        const uint64_t *__restrict__ new_locus /* exclusive scan of keep */,
        const uint64_t *__restrict__ new_row /* exclusive scan of kept_cnt, n_loci + 1 */,
        uint64_t *__restrict__ out_row_ptr, uint32_t *__restrict__ out_position,
        uint32_t *__restrict__ out_read_id, GB *__restrict__ out_gid_base, uint64_t *__restrict__ view_off) {
    extern __shared__ uint32_t s_mask[];
    const uint32_t mask_words = (n_groups + 31) / 32;
    for (uint32_t i = threadIdx.x; i < mask_words; i += blockDim.x) {
        s_mask[i] = in_mask[i];
    }
    __syncthreads();
    const int lane = threadIdx.x & 31;
    const uint64_t warps_total = static_cast<uint64_t>(gridDim.x) * (FILTER_THREADS / 32);
    for (uint64_t l = static_cast<uint64_t>(blockIdx.x) * (FILTER_THREADS / 32) + (threadIdx.x >> 5); l < n_loci;
         l += warps_total) {
        if (!keep[l]) {
            continue;
        }
        const uint64_t nl = new_locus[l];
        uint64_t w = new_row[l];
        const uint64_t e0 = row_ptr[l], e1 = row_ptr[l + 1];
        if (lane == 0) {
            out_row_ptr[nl] = w;
            out_position[nl] = position[l];
            if (!WITH_IDS) {
                view_off[nl] = e0;
            }
        }
        if (!WITH_IDS) { // a plain copy of the locus' (group, base) entries
            for (uint64_t i = lane; i < e1 - e0; i += 256) {
                GB v[8];
#pragma unroll
                for (int u = 0; u < 8; ++u) {
                    v[u] = i + 32 * u < e1 - e0 ? gid_base[e0 + i + 32 * u] : GB(0);
                }
#pragma unroll
                for (int u = 0; u < 8; ++u) {
                    if (i + 32 * u < e1 - e0) {
                        out_gid_base[w + i + 32 * u] = v[u];
                    }
                }
            }
            continue;
        }
        for (uint64_t base = e0; base < e1; base += 128) { // four independent 32-entry groups in flight
            uint32_t gb[4], rid[4];
            bool in[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const uint64_t e = base + 32 * u + lane;
                gb[u] = e < e1 ? gid_base[e] : 0;
                rid[u] = e < e1 ? read_id[e] : 0;
                in[u] = e < e1;
            }
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const uint32_t gid = gb[u] >> 2;
                in[u] = in[u] && gid < n_groups && ((s_mask[gid >> 5] >> (gid & 31)) & 1u);
                const uint32_t ballot = __ballot_sync(0xffffffffu, in[u]);
                if (in[u]) {
                    const uint64_t dst = w + __popc(ballot & ((1u << lane) - 1u));
                    out_read_id[dst] = rid[u];
                    out_gid_base[dst] = static_cast<GB>(gb[u]);
                }
                w += __popc(ballot);
            }
        }
    }
}

__global__ void remap_chr_ptr_kernel(const uint64_t *__restrict__ chr_ptr, uint32_t n_chr,
                                     const uint64_t *__restrict__ new_locus, uint64_t *__restrict__ out_chr_ptr,
                                     uint64_t *__restrict__ out_row_ptr, const uint64_t *__restrict__ new_row,
                                     uint64_t n_loci) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i <= n_chr) {
        out_chr_ptr[i] = new_locus[chr_ptr[i]];
    }
    if (i == 0) {
        out_row_ptr[new_locus[n_loci]] = new_row[n_loci]; // closing offset
    }
}

__global__ void is_significant_kernel(const uint16_t *__restrict__ counts4, uint64_t n, FilterParams fp,
                                      uint8_t *__restrict__ out) {
    const uint64_t i = static_cast<uint64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
    if (i < n) {
        out[i] = is_significant_dev(counts4[4 * i], counts4[4 * i + 1], counts4[4 * i + 2], counts4[4 * i + 3], fp) ? 1 : 0;
    }
}

FilterParams make_params(double theta, int cell_proportion) {
    FilterParams fp;
    fp.theta = theta;
    fp.log_theta_3 = std::log(theta / 3);           // Filter::Filter, util/is_significant.cpp:48-52
    fp.log_one_minus_theta = std::log(1 - theta);
    fp.log_1_4 = std::log(1. / 4);                  // :38
    fp.log_homo_prior = std::log(0.0005);           // :45
    fp.cell_proportion = cell_proportion;
    return fp;
}

} // namespace

int sgpu_is_significant_impl(sgpu_ctx *ctx, const uint16_t *h_counts4, uint64_t n, double theta,
                             int cell_proportion, uint8_t *h_out) {
    if (cell_proportion < 0 || cell_proportion > 4) {
        return sgpu_fail(ctx, SGPU_E_ARG, "cell_proportion must be 0..4");
    }
    if (n == 0) {
        return SGPU_OK;
    }
    cudaStream_t st = ctx->stream;
    DevBuf<uint16_t> d_in;
    DevBuf<uint8_t> d_out;
    SGPU_CUDA(ctx, d_in.alloc(4 * n, ctx));
    SGPU_CUDA(ctx, d_out.alloc(n, ctx));
    SGPU_CUDA(ctx, cudaMemcpyAsync(d_in.p, h_counts4, 4 * n * sizeof(uint16_t), cudaMemcpyHostToDevice, st));
    SGPU_LAUNCH(ctx, (is_significant_kernel<<<static_cast<unsigned>(ceil_div_u64(n, 256)), 256, 0, st>>>(d_in.p, n, make_params(theta, cell_proportion), d_out.p)));
    SGPU_CUDA(ctx, cudaGetLastError());
    SGPU_CUDA(ctx, cudaMemcpyAsync(h_out, d_out.p, n, cudaMemcpyDeviceToHost, st));
    SGPU_CUDA(ctx, cudaStreamSynchronize(st));
    return SGPU_OK;
}

int sgpu_filter_impl(sgpu_ctx *ctx, const sgpu_pileup *in, const uint32_t *h_id_to_pos, uint32_t n_groups,
                     double theta, int cell_proportion, sgpu_pileup **filtered, double *avg_coverage) {
    if (cell_proportion < 0 || cell_proportion > 4) {
        return sgpu_fail(ctx, SGPU_E_ARG, "cell_proportion must be 0..4");
    }
    cudaStream_t st = ctx->stream;
    const uint64_t P = in->n_loci;
    if (in->view_read_id) { // the input is itself a filtered pileup that reads its ids through a view: give it its own first
        SGPU_TRY(sgpu_pileup_materialize(ctx, in));
    }
    // sub-cluster membership as a bit mask (id_to_pos[gid] != NO_POS, util/is_significant.cpp:169)
    const uint32_t mask_words = (n_groups + 31) / 32;
    std::vector<uint32_t> h_mask(mask_words ? mask_words : 1, 0);
    const uint32_t no_pos = in->wide ? SGPU_NO_POS_WIDE : SGPU_NO_POS;
    if (!in->wide && n_groups > 16384) {
        return sgpu_fail(ctx, SGPU_E_ARG, "more than 16 384 groups need a wide pileup (sgpu_pileup_upload_wide)");
    }
    bool all_in = true; // every group belongs to the sub-cluster (the root of the recursion): kept loci keep all their entries
    for (uint32_t g = 0; g < n_groups; ++g) {
        if (h_id_to_pos[g] != no_pos) {
            h_mask[g >> 5] |= 1u << (g & 31);
        } else {
            all_in = false;
        }
    }
    const size_t smem = std::max<size_t>(512, h_mask.size()) * sizeof(uint32_t);
    if (smem > 200 * 1024) {
        return sgpu_fail(ctx, SGPU_E_ARG, "n_groups too large for the shared-memory membership mask");
    }
    DevBuf<uint32_t> d_mask, d_cnt;
    DevBuf<uint8_t> d_keep;
    DevBuf<uint64_t> d_new_locus, d_new_row;
    DevBuf<int> d_err;
    DevBuf<unsigned int> d_maxk;
    SGPU_CUDA(ctx, d_maxk.alloc(1, ctx));
    SGPU_CUDA(ctx, cudaMemsetAsync(d_maxk.p, 0, sizeof(unsigned int), st));
    SGPU_CUDA(ctx, d_mask.alloc(h_mask.size(), ctx));
    SGPU_CUDA(ctx, d_cnt.alloc(P, ctx));
    SGPU_CUDA(ctx, d_keep.alloc(P, ctx));
    SGPU_CUDA(ctx, d_new_locus.alloc(P + 1, ctx));
    SGPU_CUDA(ctx, d_new_row.alloc(P + 1, ctx));
    SGPU_CUDA(ctx, d_err.alloc(1, ctx));
    SGPU_CUDA(ctx, cudaMemsetAsync(d_err.p, 0, sizeof(int), st));
    SGPU_CUDA(ctx, cudaMemcpyAsync(d_mask.p, h_mask.data(), h_mask.size() * sizeof(uint32_t), cudaMemcpyHostToDevice, st));

    const FilterParams fp = make_params(theta, cell_proportion);
    const unsigned grid = static_cast<unsigned>(std::min<uint64_t>(ceil_div_u64(P ? P : 1, FILTER_THREADS / 32),
                                                                  static_cast<uint64_t>(ctx->sm_count) * 32));
    if (smem > 48 * 1024) {
        SGPU_GB(in, (void)gid_base_; SGPU_CUDA(ctx, cudaFuncSetAttribute(filter_count_kernel<GB>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
                SGPU_CUDA(ctx, cudaFuncSetAttribute(filter_compact_kernel<GB, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
                SGPU_CUDA(ctx, cudaFuncSetAttribute(filter_compact_kernel<GB, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)));
    }
    if (P) {
        const unsigned cgrid = static_cast<unsigned>(std::min<uint64_t>(ceil_div_u64(P, FILTER_THREADS / 32), static_cast<uint64_t>(ctx->sm_count) * 8));
        DevBuf<uint4> d_counts;
        SGPU_CUDA(ctx, d_counts.alloc(P, ctx));
        SGPU_GB(in, SGPU_LAUNCH(ctx, (filter_count_kernel<GB><<<cgrid, FILTER_THREADS, smem, st>>>(in->d_row_ptr, gid_base_, P, d_mask.p, n_groups,
                                                                                               d_counts.p, d_err.p))));
        SGPU_LAUNCH(ctx, (filter_decide_kernel<<<static_cast<unsigned>(ceil_div_u64(P, FILTER_THREADS)), FILTER_THREADS, 0, st>>>(
                                 d_counts.p, P, fp, d_keep.p, d_cnt.p, d_maxk.p)));
        SGPU_CUDA(ctx, cudaGetLastError());
    }
    SGPU_TRY(sgpu_scan_u8_u64(ctx, d_keep.p, d_new_locus.p, P));
    SGPU_TRY(sgpu_scan_u32_u64(ctx, d_cnt.p, d_new_row.p, P));
    // totals -> host
    SGPU_CUDA(ctx, cudaMemcpyAsync(&ctx->h_scratch[0], d_new_locus.p + P, sizeof(uint64_t), cudaMemcpyDeviceToHost, st));
    SGPU_CUDA(ctx, cudaMemcpyAsync(&ctx->h_scratch[1], d_new_row.p + P, sizeof(uint64_t), cudaMemcpyDeviceToHost, st));
    SGPU_CUDA(ctx, cudaMemcpyAsync(&ctx->h_scratch[2], d_err.p, sizeof(int), cudaMemcpyDeviceToHost, st));
    SGPU_CUDA(ctx, cudaMemcpyAsync(&ctx->h_scratch[3], d_maxk.p, sizeof(unsigned int), cudaMemcpyDeviceToHost, st));
    SGPU_CUDA(ctx, cudaStreamSynchronize(st));
    if (static_cast<int>(ctx->h_scratch[2] & 0xFFFFFFFFu) != 0) {
        return sgpu_fail(ctx, SGPU_E_CELL_RANGE, "pileup holds a group id >= n_groups (%u)", n_groups);
    }
    const uint64_t Lk = ctx->h_scratch[0], Ek = ctx->h_scratch[1];

    sgpu_pileup *out = new sgpu_pileup();
    PileupOwner out_owner(ctx, out);
    out->n_chr = in->n_chr;
    out->n_loci = Lk;
    out->n_entries = Ek;
    out->max_row = static_cast<uint32_t>(ctx->h_scratch[3] & 0xFFFFFFFFu); // read linking sizes its table by it
    out->owns = true;
    out->wide = in->wide;
    out->h_chr_ptr = new uint64_t[in->n_chr + 1];
    SGPU_CUDA(ctx, sgpu_dev_alloc(ctx, reinterpret_cast<void **>(&out->d_chr_ptr), (in->n_chr + 1) * sizeof(uint64_t)));
    SGPU_CUDA(ctx, sgpu_dev_alloc(ctx, reinterpret_cast<void **>(&out->d_row_ptr), (Lk + 1) * sizeof(uint64_t)));
    SGPU_CUDA(ctx, sgpu_dev_alloc(ctx, reinterpret_cast<void **>(&out->d_position), (Lk ? Lk : 1) * sizeof(uint32_t)));
    // Read ids: copied, or - when nothing inside a kept locus is dropped and the source holds them on the device - left where
    // they are (sgpu_pileup::view_read_id; SECEDO_B200_FILTER_VIEW=0 always copies)
    const char *env_view = getenv("SECEDO_B200_FILTER_VIEW");
    const bool as_view = all_in && P && in->d_read_id != nullptr && !(env_view && env_view[0] == '0');
    if (as_view) {
        SGPU_CUDA(ctx, sgpu_dev_alloc(ctx, reinterpret_cast<void **>(&out->view_off), (Lk ? Lk : 1) * sizeof(uint64_t)));
    } else {
        SGPU_CUDA(ctx, sgpu_dev_alloc(ctx, reinterpret_cast<void **>(&out->d_read_id), (Ek ? Ek : 1) * sizeof(uint32_t)));
    }
    if (in->wide) {
        SGPU_CUDA(ctx, sgpu_dev_alloc(ctx, reinterpret_cast<void **>(&out->d_gid_base32), (Ek ? Ek : 1) * sizeof(uint32_t)));
    } else {
        SGPU_CUDA(ctx, sgpu_dev_alloc(ctx, reinterpret_cast<void **>(&out->d_gid_base), (Ek ? Ek : 1) * sizeof(uint16_t)));
    }
    if (P) {
        // lazily uploaded pileup: the read ids of the KEPT loci are pulled straight from pinned host memory (the
        // rejected loci's never cross PCIe)
        const uint32_t *rid_src = in->d_read_id;
        DevBuf<uint32_t> pulled;
        if (!rid_src) {
            SGPU_CUDA(ctx, pulled.alloc(in->n_entries ? in->n_entries : 1, ctx));
            SGPU_LAUNCH(ctx, (pull_read_ids_kernel<<<grid, FILTER_THREADS, 0, st>>>(in->d_row_ptr, P, d_keep.p, in->zc_read_id, pulled.p)));
            rid_src = pulled.p;
        }
        if (as_view) {
            SGPU_GB(in, SGPU_LAUNCH(ctx, (filter_compact_kernel<GB, false><<<grid, FILTER_THREADS, mask_words * sizeof(uint32_t), st>>>(
                                    in->d_row_ptr, in->d_position, rid_src, gid_base_, P, d_mask.p, n_groups, d_keep.p, d_new_locus.p,
                                    d_new_row.p, out->d_row_ptr, out->d_position, nullptr,
                                    const_cast<GB *>(in->wide ? reinterpret_cast<const GB *>(out->d_gid_base32)
                                                              : reinterpret_cast<const GB *>(out->d_gid_base)),
                                    out->view_off))));
            sgpu_pileup *src = const_cast<sgpu_pileup *>(in);
            src->view_refs.fetch_add(1);
            out->view_src = src;
            out->view_read_id = in->d_read_id;
        } else {
            SGPU_GB(in, SGPU_LAUNCH(ctx, (filter_compact_kernel<GB, true><<<grid, FILTER_THREADS, mask_words * sizeof(uint32_t), st>>>(
                                    in->d_row_ptr, in->d_position, rid_src, gid_base_, P, d_mask.p, n_groups, d_keep.p, d_new_locus.p,
                                    d_new_row.p, out->d_row_ptr, out->d_position, out->d_read_id,
                                    const_cast<GB *>(in->wide ? reinterpret_cast<const GB *>(out->d_gid_base32)
                                                              : reinterpret_cast<const GB *>(out->d_gid_base)),
                                    nullptr))));
        }
    }
    SGPU_LAUNCH(ctx, (remap_chr_ptr_kernel<<<(in->n_chr + 256) / 256, 256, 0, st>>>(in->d_chr_ptr, in->n_chr, d_new_locus.p, out->d_chr_ptr,
                                                                 out->d_row_ptr, d_new_row.p, P)));
    SGPU_CUDA(ctx, cudaGetLastError());
    SGPU_CUDA(ctx, cudaMemcpyAsync(out->h_chr_ptr, out->d_chr_ptr, (in->n_chr + 1) * sizeof(uint64_t), cudaMemcpyDeviceToHost, st));
    SGPU_CUDA(ctx, cudaStreamSynchronize(st));
    if (avg_coverage) {
        *avg_coverage = Lk == 0 ? 0.0 : static_cast<double>(Ek) / static_cast<double>(Lk); // :188, in 64 bits
    }
    *filtered = out_owner.release();
    return SGPU_OK;
}
