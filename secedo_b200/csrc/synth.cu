// Deterministic synthetic pileups generated directly in HBM (bench / large-scale test input).
//
// The headline configuration (8 000 cells at 0.5x) is ~24 KB of pileup per locus, so inputs of
// benchmark size cannot be produced on the host in reasonable time. Every quantity is a pure
// function of (seed, global locus index, slot), counter-based, so any sub-range can be downloaded
// and replayed through the CPU oracle. Model: loci ~`spacing` bp apart; classes
// noise / germline het / somatic het in one clone; reads per locus ~ N(lambda, lambda) rounded
// (normal approximation of Poisson), cells uniform; each base flipped with probability theta;
// a fraction p_multi of the fresh fragments of a locus also covers the next locus; a fraction
// p_mate of the entries is followed (at the end of the locus) by a second entry of the same read
// (overlapping mates), which disagrees with probability p_mate_mismatch.
#include "common.cuh"

#include <algorithm>

namespace {

constexpr int TB = 256;

__device__ __forceinline__ uint64_t splitmix(uint64_t x) {
    x += 0x9E3779B97F4A7C15ull;
    x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ull;
    x = (x ^ (x >> 27)) * 0x94D049BB133111EBull;
    return x ^ (x >> 31);
}
__device__ __forceinline__ uint64_t h3(uint64_t seed, uint64_t a, uint64_t b, uint64_t c) {
    return splitmix(splitmix(splitmix(seed ^ (a * 0xD1B54A32D192ED03ull)) ^ (b * 0x8CB92BA72F3D8DD7ull)) ^ c);
}
__device__ __forceinline__ float u01(uint64_t h) { return (h >> 40) * (1.0f / 16777216.0f); }

struct SynthP {
    uint32_t n_cells, n_chr, loci_per_chr, n_clones;
    float coverage, frac_somatic, frac_germline, theta, p_multi, p_mate, p_mate_mismatch;
    uint32_t spacing;
    uint64_t seed;
};

// fresh fragments starting at locus g
__device__ __forceinline__ uint32_t n_fresh(const SynthP &p, uint64_t g) {
    const float lambda = p.n_cells * p.coverage / (1.0f + p.p_multi);
    float z = -6.0f;
#pragma unroll
    for (int k = 0; k < 12; ++k) {
        z += u01(h3(p.seed, g, 0x100 + k, 1));
    }
    const float v = lambda + sqrtf(lambda) * z;
    return v < 1.0f ? 1u : static_cast<uint32_t>(v + 0.5f);
}
__device__ __forceinline__ uint32_t n_cont(const SynthP &p, uint64_t g) { // fragments continuing from locus g-1
    if (g % p.loci_per_chr == 0) {
        return 0;
    }
    return static_cast<uint32_t>(p.p_multi * n_fresh(p, g - 1));
}
__device__ __forceinline__ bool has_mate(const SynthP &p, uint64_t g, uint32_t slot) {
    return p.p_mate > 0.f && u01(h3(p.seed, g, slot, 2)) < p.p_mate;
}
__device__ __forceinline__ uint32_t cell_of_fresh(const SynthP &p, uint64_t g, uint32_t f) {
    return static_cast<uint32_t>(h3(p.seed, g, f, 3) % p.n_cells);
}
__device__ __forceinline__ uint32_t observed_base(const SynthP &p, uint64_t g, uint32_t slot, uint32_t cell) {
    const uint64_t hl = h3(p.seed, g, 0, 4);
    const float cls = u01(hl);
    const uint32_t ref = (hl >> 8) & 3u, alt = (ref + 1 + ((hl >> 10) % 3)) & 3u;
    const uint32_t som_clone = static_cast<uint32_t>((hl >> 16) % p.n_clones);
    const uint32_t clone = static_cast<uint32_t>(static_cast<uint64_t>(cell) * p.n_clones / p.n_cells);
    const bool het = cls < p.frac_somatic ? clone == som_clone : cls < p.frac_somatic + p.frac_germline;
    const uint64_t he = h3(p.seed, g, slot, 5);
    uint32_t base = (het && (he & 1)) ? alt : ref;
    if (u01(he) < p.theta) {
        base = (base + 1 + ((he >> 4) % 3)) & 3u;
    }
    return base;
}

__global__ void __launch_bounds__(TB) synth_count_kernel(SynthP p, uint64_t n_loci, uint32_t *__restrict__ cnt,
                                                         uint32_t *__restrict__ position) {
    const int lane = threadIdx.x & 31;
    const uint64_t g = (static_cast<uint64_t>(blockIdx.x) * TB + threadIdx.x) >> 5;
    if (g >= n_loci) {
        return;
    }
    const uint32_t slots = n_cont(p, g) + n_fresh(p, g);
    uint32_t dups = 0;
    for (uint32_t s = lane; s < slots; s += 32) {
        dups += has_mate(p, g, s) ? 1 : 0;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        dups += __shfl_xor_sync(0xffffffffu, dups, o);
    }
    if (lane == 0) {
        cnt[g] = slots + dups;
        const uint32_t idx = static_cast<uint32_t>(g % p.loci_per_chr);
        position[g] = 1000u + idx * p.spacing + static_cast<uint32_t>(h3(p.seed, g, 0, 6) % (p.spacing / 2 + 1));
    }
}

template <typename GB>
__global__ void __launch_bounds__(TB) synth_fill_kernel(SynthP p, uint64_t n_loci, const uint64_t *__restrict__ row_ptr,
                                                        uint32_t *__restrict__ read_id, GB *__restrict__ gid_base) {
    const int lane = threadIdx.x & 31;
    const uint64_t g = (static_cast<uint64_t>(blockIdx.x) * TB + threadIdx.x) >> 5;
    if (g >= n_loci) {
        return;
    }
    const uint64_t e0 = row_ptr[g];
    const uint32_t nc = n_cont(p, g), slots = nc + n_fresh(p, g);
    const uint64_t prev0 = nc ? row_ptr[g - 1] : 0;
    const uint32_t prev_nc = nc ? n_cont(p, g - 1) : 0;
    uint32_t dup_base = 0;
    for (uint32_t s0 = 0; s0 < slots; s0 += 32) {
        const uint32_t s = s0 + lane;
        bool dup = false;
        uint32_t rid = 0, cell = 0, base = 0;
        if (s < slots) {
            if (s < nc) { // continues the s-th fresh fragment of the previous locus
                rid = static_cast<uint32_t>(prev0 + prev_nc + s);
                cell = cell_of_fresh(p, g - 1, s);
            } else {
                rid = static_cast<uint32_t>(e0 + s);
                cell = cell_of_fresh(p, g, s - nc);
            }
            rid *= 2654435761u; // odd multiplier: a bijection on 32 bits, ids are neither sorted nor dense
            base = observed_base(p, g, s, cell);
            read_id[e0 + s] = rid;
            gid_base[e0 + s] = static_cast<GB>((cell << 2) | base);
            dup = has_mate(p, g, s);
        }
        const uint32_t ballot = __ballot_sync(0xffffffffu, dup);
        if (dup) {
            const uint64_t d = e0 + slots + dup_base + __popc(ballot & ((1u << lane) - 1u));
            const uint64_t hm = h3(p.seed, g, s, 7);
            const uint32_t b2 = u01(hm) < p.p_mate_mismatch ? (base + 1 + ((hm >> 4) % 3)) & 3u : base;
            read_id[d] = rid;
            gid_base[d] = static_cast<GB>((cell << 2) | b2);
        }
        dup_base += __popc(ballot);
    }
}

} // namespace

extern "C" int sgpu_synth_pileup(sgpu_ctx *ctx, const sgpu_synth_params *sp, sgpu_pileup **out) {
    SGPU_CUDA(ctx, cudaSetDevice(ctx->device));
    cudaStream_t st = ctx->stream;
    if (sp->n_cells == 0 || sp->n_cells >= (1u << 27) || sp->n_clones == 0 || sp->spacing < 2 || sp->n_chr == 0 || sp->n_chr > 255) {
        return sgpu_fail(ctx, SGPU_E_ARG, "invalid synthetic pileup parameters");
    }
    SynthP p;
    p.n_cells = sp->n_cells;
    p.n_chr = sp->n_chr;
    p.loci_per_chr = sp->loci_per_chr;
    p.n_clones = sp->n_clones;
    p.coverage = sp->coverage;
    p.frac_somatic = sp->frac_somatic;
    p.frac_germline = sp->frac_germline;
    p.theta = sp->theta;
    p.p_multi = sp->p_multi;
    p.p_mate = sp->p_mate;
    p.p_mate_mismatch = sp->p_mate_mismatch;
    p.spacing = sp->spacing;
    p.seed = sp->seed;
    const uint64_t P = static_cast<uint64_t>(sp->n_chr) * sp->loci_per_chr;
    sgpu_pileup *pl = new sgpu_pileup();
    PileupOwner pl_owner(ctx, pl);
    pl->n_chr = sp->n_chr;
    pl->n_loci = P;
    pl->owns = true;
    pl->h_chr_ptr = new uint64_t[sp->n_chr + 1];
    for (uint32_t c = 0; c <= sp->n_chr; ++c) {
        pl->h_chr_ptr[c] = static_cast<uint64_t>(c) * sp->loci_per_chr;
    }
    SGPU_CUDA(ctx, sgpu_dev_alloc(ctx, reinterpret_cast<void **>(&pl->d_chr_ptr), (sp->n_chr + 1) * sizeof(uint64_t)));
    SGPU_CUDA(ctx, cudaMemcpyAsync(pl->d_chr_ptr, pl->h_chr_ptr, (sp->n_chr + 1) * sizeof(uint64_t), cudaMemcpyHostToDevice, st));
    SGPU_CUDA(ctx, sgpu_dev_alloc(ctx, reinterpret_cast<void **>(&pl->d_row_ptr), (P + 1) * sizeof(uint64_t)));
    SGPU_CUDA(ctx, sgpu_dev_alloc(ctx, reinterpret_cast<void **>(&pl->d_position), std::max<uint64_t>(P, 1) * sizeof(uint32_t)));
    DevBuf<uint32_t> cnt;
    SGPU_CUDA(ctx, cnt.alloc(std::max<uint64_t>(P, 1), ctx));
    const unsigned grid = static_cast<unsigned>(ceil_div_u64(std::max<uint64_t>(P, 1) * 32, TB));
    SGPU_LAUNCH(ctx, (synth_count_kernel<<<grid, TB, 0, st>>>(p, P, cnt.p, pl->d_position)));
    SGPU_CUDA(ctx, cudaGetLastError());
    SGPU_TRY(sgpu_scan_u32_u64(ctx, cnt.p, pl->d_row_ptr, P));
    SGPU_CUDA(ctx, cudaMemcpyAsync(&ctx->h_scratch[0], pl->d_row_ptr + P, sizeof(uint64_t), cudaMemcpyDeviceToHost, st));
    SGPU_CUDA(ctx, cudaStreamSynchronize(st));
    pl->n_entries = ctx->h_scratch[0];
    if (pl->n_entries >= 0xFFFFFFF0ull) {
        return sgpu_fail(ctx, SGPU_E_ARG, "synthetic pileup too large for 32-bit read ids");
    }
    SGPU_CUDA(ctx, sgpu_dev_alloc(ctx, reinterpret_cast<void **>(&pl->d_read_id), std::max<uint64_t>(pl->n_entries, 1) * sizeof(uint32_t)));
    pl->wide = sp->n_cells > 16383; // beyond the reference's 14-bit group ids (cfg5)
    if (pl->wide) {
        SGPU_CUDA(ctx, sgpu_dev_alloc(ctx, reinterpret_cast<void **>(&pl->d_gid_base32), std::max<uint64_t>(pl->n_entries, 1) * sizeof(uint32_t)));
        SGPU_LAUNCH(ctx, (synth_fill_kernel<uint32_t><<<grid, TB, 0, st>>>(p, P, pl->d_row_ptr, pl->d_read_id, pl->d_gid_base32)));
    } else {
        SGPU_CUDA(ctx, sgpu_dev_alloc(ctx, reinterpret_cast<void **>(&pl->d_gid_base), std::max<uint64_t>(pl->n_entries, 1) * sizeof(uint16_t)));
        SGPU_LAUNCH(ctx, (synth_fill_kernel<uint16_t><<<grid, TB, 0, st>>>(p, P, pl->d_row_ptr, pl->d_read_id, pl->d_gid_base)));
    }
    SGPU_CUDA(ctx, cudaGetLastError());
    SGPU_CUDA(ctx, cudaStreamSynchronize(st));
    *out = pl_owner.release();
    return SGPU_OK;
}
