// K3 — first-order read-pair counts as a dense int8 GEMM on the 5th-generation tensor cores.
//
// Per locus l and cell i let c_b(i) be the number of surviving reads of cell i showing base b.
// The reference's per-locus contributions are  S_ij = sum_b c_b(i) c_b(j)  (same base) and
// T_ij = c(i) c(j), c = sum_b c_b  (all pairs), D = T - S. With the 4x4 Hadamard matrix H
// (H H^T = 4 I, first row all ones) and u = H c:
//        u_0(i) u_0(j)                       = T_ij
//        u_1(i) u_1(j) + u_2 u_2 + u_3 u_3   = 4 S_ij - T_ij
// so ONE K-dimension of 4 int8 values per (cell, locus) feeds two int32 accumulators — 4 MMA
// K-slices per 32 loci instead of the 5 (four base planes + a total plane) of the direct form.
//
//   stage_count_kernel   entries -> packed per-(locus, cell) base counts (4 x u8 in a u32), L2 atomics
//   transform_kernel     packed counts -> Hadamard planes, transposed into the K-major tile layout
//                        U[cell][k-block][plane][32 loci] (one 128-byte row per cell and k-block)
//   syrk_kernel          persistent, warp-specialised: warp 0 = TMA producer (cp.async.bulk.tensor,
//                        128B swizzle, 4-stage mbarrier ring), warp 1 = tcgen05.mma.kind::i8 issuer
//                        (M=128, N=256, K=32; accumulators Q = 4S - T and T in TMEM, 2 x 256
//                        columns), warps 2-5 = epilogue (tcgen05.ld, S = (Q+T)/4, D = T - S,
//                        RED.ADD into the int32 planes, upper triangle only)
// Only output tiles that intersect the upper triangle are computed; the K range of a panel is
// split across CTAs when there are fewer tiles than SMs.
#include "common.cuh"

#include <cuda.h>

#include <algorithm>
#include <cstdlib>
#include <vector>

namespace {

constexpr int BM = 128;            // rows (cells) of an output tile = UMMA M
constexpr int BN = 256;            // columns (cells) of an output tile = UMMA N
constexpr int KB_BYTES = 128;      // bytes of K per stage and row: 32 loci x 4 planes
constexpr int LOCI_PER_KB = 32;
constexpr int STAGES = 4;
constexpr int A_BYTES = BM * KB_BYTES;  // 16 KB
constexpr int B_BYTES = BN * KB_BYTES;  // 32 KB
constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + 1024 /* alignment slack */ + 256 /* barriers */;
constexpr int GEMM_THREADS = 192;  // 6 warps
constexpr uint32_t TMEM_COLS = 512;

// Panel layout. One 128-byte row per (cell, k-block of 32 loci). The rows of CK consecutive k-blocks
// form a chunk [chunk][cell][k-block in chunk]: the ~1 200 loci that are staged concurrently then touch
// a few tens of MB of address space (2 MB pages: a handful of TLB entries) instead of one row in each of
// the 8 192 cell stripes of a [cell][K] matrix, 128 KB apart, which made the staging atomics TLB bound.
// The GEMM reads the same buffer through a 3-D tensor map (bytes in chunk row, cell, chunk).
struct PanelLayout {
    uint32_t ck_shift;    // log2(k-blocks per chunk)
    uint32_t cell_words;  // 32-bit words between consecutive cells of a chunk: an odd number of 128-byte lines
    uint64_t chunk_words; // words per chunk = n_pad * cell_words
    __host__ __device__ uint64_t row(uint32_t cell, uint32_t kb) const { // first word of the row (cell, kb)
        const uint32_t ch = kb >> ck_shift, kbi = kb - (ch << ck_shift);
        return ch * chunk_words + static_cast<uint64_t>(cell) * cell_words + kbi * 32u;
    }
    __host__ __device__ uint64_t word(uint32_t cell, uint32_t col) const { // staged counts of (cell, locus col)
        return row(cell, col >> 5) + (col & 31u);
    }
};

// ------------------------------------------------------------------------------------------------
// staging
//
// The panel buffer holds one 128-byte row per (cell, k-block of 32 loci). While the counts are being
// accumulated, word j of the row is locus j's four base counts (4 x u8); transform_kernel then turns
// the row IN PLACE into the K-major operand layout [plane 0..3][32 loci] of Hadamard planes, so the
// staged counts cost one memset, one pass of atomics and one read + write of the panel.
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ void add_count(uint32_t *__restrict__ U, uint64_t word, uint32_t base, int *__restrict__ err) {
    const uint32_t sh = 8u * base;
    const uint32_t old = atomicAdd(&U[word], 1u << sh);
    if (((old >> sh) & 0xFFu) >= 127u) {
        atomicExch(err, SGPU_E_COUNT_RANGE); // > 127 reads of one cell at one locus: outside int8
    }
}

// entries that are the only entry of their read, loci [l0, l1): one block per locus. Four entries per
// thread are in flight (all loads first, then the atomics, then the range checks): the chain
// gid_base -> group map -> atomic is three dependent memory round trips.
template <bool RETURNING>
__global__ void __launch_bounds__(256) stage_main_kernel(const uint64_t *__restrict__ row_ptr,
                                                         const uint16_t *__restrict__ gid_base,
                                                         const uint32_t *__restrict__ sp_bits,
                                                         const uint32_t *__restrict__ gmap, uint32_t n_groups,
                                                         uint32_t n_cells, uint32_t l0, uint32_t l1, uint32_t col0,
                                                         PanelLayout pl, uint32_t *__restrict__ U,
                                                         int *__restrict__ err) {
    constexpr int UNROLL = 4;
    for (uint32_t l = l0 + blockIdx.x; l < l1; l += gridDim.x) {
        const uint64_t e0 = row_ptr[l], e1 = row_ptr[l + 1];
        const uint64_t col = pl.word(0, l - col0); // + cell * cell_words
        for (uint64_t eb = e0 + threadIdx.x; eb < e1; eb += 256 * UNROLL) {
            uint32_t gb[UNROLL], cell[UNROLL], old[UNROLL];
            bool live[UNROLL];
#pragma unroll
            for (int u = 0; u < UNROLL; ++u) {
                const uint64_t e = eb + 256 * u;
                live[u] = e < e1;
                gb[u] = live[u] ? gid_base[e] : 0;
                if (live[u] && ((sp_bits[e >> 5] >> (e & 31)) & 1u)) {
                    live[u] = false; // staged by stage_special_kernel
                }
            }
#pragma unroll
            for (int u = 0; u < UNROLL; ++u) {
                const uint32_t gid = gb[u] >> 2;
                cell[u] = 0;
                if (live[u]) {
                    if (gid >= n_groups || (cell[u] = gmap[gid]) >= n_cells) {
                        atomicExch(err, SGPU_E_CELL_RANGE);
                        live[u] = false;
                    }
                }
            }
#pragma unroll
            for (int u = 0; u < UNROLL; ++u) {
                old[u] = 0;
                if (live[u]) {
                    const uint32_t inc = 1u << (8u * (gb[u] & 3u));
                    uint32_t *w = &U[static_cast<uint64_t>(cell[u]) * pl.cell_words + col];
                    if (RETURNING) {
                        old[u] = atomicAdd(w, inc);
                    } else {
                        atomicAdd(w, inc); // RED
                    }
                }
            }
            if (RETURNING) {
#pragma unroll
                for (int u = 0; u < UNROLL; ++u) {
                    if (live[u] && ((old[u] >> (8u * (gb[u] & 3u))) & 0xFFu) >= 127u) {
                        atomicExch(err, SGPU_E_COUNT_RANGE); // > 127 reads of one cell at one locus: outside int8
                    }
                }
            }
        }
    }
}

// entries of reads with several entries that survived the mate rule; sp_locus is ascending, the loci
// [l0, l1) own a contiguous range of special entries that is found by bisection
__global__ void __launch_bounds__(256) stage_special_kernel(const uint32_t *__restrict__ sp_code,
                                                            const uint32_t *__restrict__ sp_locus, uint64_t n_special,
                                                            uint32_t l0, uint32_t l1, uint32_t col0, PanelLayout pl,
                                                            uint32_t *__restrict__ U, int *__restrict__ err) {
    uint64_t lo = 0, hi = n_special;
    while (lo < hi) {
        const uint64_t mid = (lo + hi) >> 1;
        if (sp_locus[mid] < l0) {
            lo = mid + 1;
        } else {
            hi = mid;
        }
    }
    for (uint64_t s = lo + static_cast<uint64_t>(blockIdx.x) * 256 + threadIdx.x; s < n_special; s += static_cast<uint64_t>(gridDim.x) * 256) {
        const uint32_t l = sp_locus[s];
        if (l >= l1) {
            break;
        }
        const uint32_t c = sp_code[s];
        if (c != CODE_DROPPED) {
            add_count(U, pl.word(code_cell(c), l - col0), code_base(c), err);
        }
    }
}

// tail correction panel: column b of the panel holds ONLY the tail reads of locus tail_loci[b]. A
// read that is the only entry... of its read was created at its own locus, i.e. behind the cutoff.
__global__ void __launch_bounds__(256) stage_tail_kernel(const uint64_t *__restrict__ row_ptr,
                                                         const uint16_t *__restrict__ gid_base,
                                                         const uint32_t *__restrict__ sp_bits,
                                                         const uint32_t *__restrict__ gmap, uint32_t n_groups,
                                                         uint32_t n_cells, const uint32_t *__restrict__ tail_loci,
                                                         PanelLayout pl, uint32_t *__restrict__ U,
                                                         int *__restrict__ err) {
    const uint32_t l = tail_loci[blockIdx.x];
    const uint64_t e0 = row_ptr[l], e1 = row_ptr[l + 1];
    for (uint64_t e = e0 + threadIdx.x; e < e1; e += 256) {
        if ((sp_bits[e >> 5] >> (e & 31)) & 1u) {
            continue;
        }
        const uint32_t gb = gid_base[e];
        const uint32_t gid = gb >> 2;
        uint32_t cell;
        if (gid >= n_groups || (cell = gmap[gid]) >= n_cells) {
            atomicExch(err, SGPU_E_CELL_RANGE);
            continue;
        }
        add_count(U, pl.word(cell, blockIdx.x), gb & 3u, err);
    }
}

__global__ void __launch_bounds__(256) stage_tail_special_kernel(const uint32_t *__restrict__ sp_code,
                                                                 const uint32_t *__restrict__ sp_locus,
                                                                 uint64_t n_special, const uint32_t *__restrict__ tail_loci,
                                                                 uint32_t n_tail /* columns of this panel */,
                                                                 PanelLayout pl, uint32_t *__restrict__ U,
                                                                 int *__restrict__ err) {
    const uint64_t s = static_cast<uint64_t>(blockIdx.x) * 256 + threadIdx.x;
    if (s >= n_special) {
        return;
    }
    const uint32_t c = sp_code[s], l = sp_locus[s];
    if (c == CODE_DROPPED || !code_tail(c)) {
        return;
    }
    uint32_t lo = 0, hi = n_tail; // tail_loci is ascending
    while (lo < hi) {
        const uint32_t mid = (lo + hi) >> 1;
        if (tail_loci[mid] < l) {
            lo = mid + 1;
        } else {
            hi = mid;
        }
    }
    if (lo < n_tail && tail_loci[lo] == l) {
        add_count(U, pl.word(code_cell(c), lo), code_base(c), err);
    }
}

// In-place Hadamard transform of the rows: 8 consecutive lanes own one 128-byte row (one cell, one
// k-block); each loads 4 loci x 4 base counts (16 B) and, after the whole warp has loaded, stores the
// four plane words of those loci.
__global__ void __launch_bounds__(256) transform_kernel(uint32_t *__restrict__ U, uint32_t kbs /* valid k-blocks */, uint32_t n_pad,
                                                        PanelLayout pl, int *__restrict__ err) {
    // consecutive 8-lane groups walk the k-blocks of a chunk row, then the cells of the chunk: contiguous
    // memory. The grid covers whole chunks (a multiple of 32 threads), rows behind kbs are skipped.
    const uint64_t t = static_cast<uint64_t>(blockIdx.x) * 256 + threadIdx.x;
    const uint32_t g = t & 7;
    const uint64_t rk = t >> 3;
    const uint32_t ck = 1u << pl.ck_shift;
    const uint64_t per_chunk = static_cast<uint64_t>(n_pad) << pl.ck_shift; // rows of a chunk
    const uint32_t ch = static_cast<uint32_t>(rk / per_chunk);
    const uint64_t in_chunk = rk - ch * per_chunk;
    const uint32_t cell = static_cast<uint32_t>(in_chunk >> pl.ck_shift);
    const uint32_t kb = (ch << pl.ck_shift) + static_cast<uint32_t>(in_chunk & (ck - 1));
    const bool valid = kb < kbs;
    uint32_t *r = U + pl.row(cell, valid ? kb : 0);
    uint4 v = make_uint4(0, 0, 0, 0);
    if (valid) {
        v = *reinterpret_cast<const uint4 *>(r + 4 * g);
    }
    const uint32_t in[4] = { v.x, v.y, v.z, v.w };
    uint32_t w[4] = { 0, 0, 0, 0 }; // per plane, 4 loci packed
    bool bad = false;
#pragma unroll
    for (int q = 0; q < 4; ++q) {
        const int c0 = in[q] & 0xFF, c1 = (in[q] >> 8) & 0xFF, c2 = (in[q] >> 16) & 0xFF, c3 = in[q] >> 24;
        const int u0 = c0 + c1 + c2 + c3;
        const int u1 = c0 - c1 + c2 - c3;
        const int u2 = c0 + c1 - c2 - c3;
        const int u3 = c0 - c1 - c2 + c3;
        bad |= u0 > 127;
        w[0] |= static_cast<uint32_t>(u0 & 0xFF) << (8 * q);
        w[1] |= static_cast<uint32_t>(u1 & 0xFF) << (8 * q);
        w[2] |= static_cast<uint32_t>(u2 & 0xFF) << (8 * q);
        w[3] |= static_cast<uint32_t>(u3 & 0xFF) << (8 * q);
    }
    __syncwarp(); // every lane of the row has its counts in registers before any plane word is written
    if (valid) {
#pragma unroll
        for (int pl_i = 0; pl_i < 4; ++pl_i) {
            r[pl_i * 8 + g] = w[pl_i];
        }
    }
    if (bad) {
        atomicExch(err, SGPU_E_COUNT_RANGE);
    }
}

// ------------------------------------------------------------------------------------------------
// PTX helpers
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return static_cast<uint32_t>(__cvta_generic_to_shared(p)); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    asm volatile(
            "{\n\t"
            ".reg .pred p;\n\t"
            "WAIT_LOOP:\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
            "@p bra WAIT_DONE;\n\t"
            "bra WAIT_LOOP;\n\t"
            "WAIT_DONE:\n\t"
            "}" ::"r"(bar),
            "r"(parity)
            : "memory");
}
__device__ __forceinline__ void tma_load_3d(uint32_t dst, const CUtensorMap *map, uint32_t bar, int32_t c0, int32_t c1, int32_t c2) {
    asm volatile(
            "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];" ::"r"(dst),
            "l"(map), "r"(bar), "r"(c0), "r"(c1), "r"(c2)
            : "memory");
}
__device__ __forceinline__ bool elect_one() {
    uint32_t pred;
    asm volatile(
            "{\n\t"
            ".reg .pred P;\n\t"
            "elect.sync _|P, 0xffffffff;\n\t"
            "selp.b32 %0, 1, 0, P;\n\t"
            "}"
            : "=r"(pred));
    return pred != 0;
}
// UMMA shared-memory descriptor: K-major, 128B swizzle, 8-row groups 1024 B apart (SBO), LBO = 1
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr) {
    uint64_t d = 0;
    d |= static_cast<uint64_t>((saddr & 0x3FFFFu) >> 4);      // start address, bits [0,14)
    d |= static_cast<uint64_t>(1) << 16;                      // leading byte offset (unused for SW128 K-major)
    d |= static_cast<uint64_t>(1024 >> 4) << 32;              // stride byte offset
    d |= static_cast<uint64_t>(1) << 46;                      // descriptor version (Blackwell)
    d |= static_cast<uint64_t>(2) << 61;                      // SWIZZLE_128B
    return d;
}
// instruction descriptor, kind::i8: D = S32, A = B = signed 8 bit, both K-major, N = 256, M = 128
constexpr uint32_t IDESC = (2u << 4) | (1u << 7) | (1u << 10) | ((BN >> 3) << 17) | ((BM >> 4) << 24);

__device__ __forceinline__ void umma_i8(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t accumulate) {
    asm volatile(
            "{\n\t"
            ".reg .pred p;\n\t"
            "setp.ne.b32 p, %4, 0;\n\t"
            "tcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, p;\n\t"
            "}" ::"r"(d_tmem),
            "l"(a_desc), "l"(b_desc), "r"(IDESC), "r"(accumulate)
            : "memory");
}
__device__ __forceinline__ void umma_commit(uint32_t bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&r)[32]) {
    asm volatile(
            "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
            "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
            "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
            : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
              "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
              "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
              "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
            : "r"(taddr));
}

struct WorkItem {
    uint32_t rb, cb;   // row block (x128), column block (x256)
    uint32_t k0, k1;   // k-block range
};

// work item w = (tile w % n_tiles, K split w / n_tiles)
struct WorkList {
    const uint2 *tiles; // (row block, column block), rasterised for L2 reuse
    uint32_t n_tiles;
    uint32_t n_work;    // n_tiles * splits
    uint32_t kbs;       // k-blocks of the panel
    uint32_t per;       // k-blocks per split
    uint32_t ck_shift;  // log2(k-blocks per chunk of the panel layout)
};
__device__ __forceinline__ WorkItem work_item(const WorkList &wl, uint32_t w) {
    const uint32_t split = w / wl.n_tiles;
    const uint2 t = wl.tiles[w - split * wl.n_tiles];
    const uint32_t k0 = split * wl.per;
    return WorkItem{ t.x, t.y, k0, min(wl.kbs, k0 + wl.per) };
}

// ------------------------------------------------------------------------------------------------
// the GEMM
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(GEMM_THREADS, 1) syrk_kernel(const __grid_constant__ CUtensorMap map_u,
                                                               const WorkList wl, int32_t *__restrict__ S,
                                                               int32_t *__restrict__ D, uint32_t n_cells, int sign) {
    const uint32_t n_work = wl.n_work;
    extern __shared__ uint8_t smem_raw[];
    // 128B swizzle needs 1024-byte aligned tiles
    uint8_t *smem = reinterpret_cast<uint8_t *>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~static_cast<uintptr_t>(1023));
    uint64_t *bars = reinterpret_cast<uint64_t *>(smem + STAGES * STAGE_BYTES);
    // bars[0..S) full, [S..2S) empty, [2S] tmem_full, [2S+1] tmem_empty, then the TMEM base address
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(bars + 2 * STAGES + 2);
    const uint32_t full0 = smem_u32(bars), empty0 = smem_u32(bars + STAGES);
    const uint32_t tmem_full = smem_u32(bars + 2 * STAGES), tmem_empty = smem_u32(bars + 2 * STAGES + 1);
    const uint32_t warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

    if (warp == 0 && lane == 0) {
        asm volatile("prefetch.tensormap [%0];" ::"l"(&map_u) : "memory");
        for (int s = 0; s < STAGES; ++s) {
            mbar_init(full0 + 8 * s, 1);
            mbar_init(empty0 + 8 * s, 1);
        }
        mbar_init(tmem_full, 1);
        mbar_init(tmem_empty, 128);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "n"(TMEM_COLS));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem_base = *tmem_slot;
    const uint32_t tmem_Q = tmem_base, tmem_T = tmem_base + BN; // columns [0,256) and [256,512)

    if (warp == 0) {
        // ===== TMA producer =====
        if (elect_one()) {
            uint32_t stage = 0, phase = 0;
            for (uint32_t w = blockIdx.x; w < n_work; w += gridDim.x) {
                const WorkItem wi = work_item(wl, w);
                for (uint32_t kb = wi.k0; kb < wi.k1; ++kb) {
                    mbar_wait(empty0 + 8 * stage, phase ^ 1);
                    const uint32_t sa = smem_u32(smem + stage * STAGE_BYTES), sb = sa + A_BYTES;
                    const uint32_t bar = full0 + 8 * stage;
                    mbar_expect_tx(bar, STAGE_BYTES);
                    const int32_t ch = static_cast<int32_t>(kb >> wl.ck_shift);
                    const int32_t kx = static_cast<int32_t>((kb - (static_cast<uint32_t>(ch) << wl.ck_shift)) * KB_BYTES);
                    tma_load_3d(sa, &map_u, bar, kx, wi.rb * BM, ch);
                    tma_load_3d(sb, &map_u, bar, kx, wi.cb * BN, ch);
                    tma_load_3d(sb + A_BYTES, &map_u, bar, kx, wi.cb * BN + 128, ch);
                    if (++stage == STAGES) {
                        stage = 0;
                        phase ^= 1;
                    }
                }
            }
        }
    } else if (warp == 1) {
        // ===== MMA issuer =====
        if (elect_one()) {
            uint32_t stage = 0, phase = 0, acc_phase = 0;
            for (uint32_t w = blockIdx.x; w < n_work; w += gridDim.x) {
                const WorkItem wi = work_item(wl, w);
                mbar_wait(tmem_empty, acc_phase ^ 1); // epilogue has drained the accumulators
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                for (uint32_t kb = wi.k0; kb < wi.k1; ++kb) {
                    mbar_wait(full0 + 8 * stage, phase);
                    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                    const uint32_t sa = smem_u32(smem + stage * STAGE_BYTES), sb = sa + A_BYTES;
                    const uint32_t first = kb == wi.k0 ? 0u : 1u;
                    // plane 0 (u_0 = reads per cell) -> T; planes 1..3 -> Q = 4S - T
                    umma_i8(tmem_T, make_desc(sa), make_desc(sb), first);
                    umma_i8(tmem_Q, make_desc(sa + 32), make_desc(sb + 32), first);
                    umma_i8(tmem_Q, make_desc(sa + 64), make_desc(sb + 64), 1u);
                    umma_i8(tmem_Q, make_desc(sa + 96), make_desc(sb + 96), 1u);
                    umma_commit(empty0 + 8 * stage); // frees the smem slot once these MMAs retire
                    if (++stage == STAGES) {
                        stage = 0;
                        phase ^= 1;
                    }
                }
                umma_commit(tmem_full); // accumulators complete
                acc_phase ^= 1;
            }
        }
    } else {
        // ===== epilogue: warps 2..5, TMEM lanes 32*(warp%4) .. +31 =====
        const uint32_t lane_base = 32 * (warp & 3);
        uint32_t acc_phase = 0;
        for (uint32_t w = blockIdx.x; w < n_work; w += gridDim.x) {
            const WorkItem wi = work_item(wl, w);
            mbar_wait(tmem_full, acc_phase);
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            const uint32_t row = wi.rb * BM + lane_base + lane;
            int32_t *Srow = S + static_cast<uint64_t>(row) * n_cells;
            int32_t *Drow = D + static_cast<uint64_t>(row) * n_cells;
#pragma unroll 1
            for (uint32_t cc = 0; cc < BN / 32; ++cc) {
                uint32_t q[32], t[32];
                tmem_ld32(tmem_Q + (lane_base << 16) + cc * 32, q);
                tmem_ld32(tmem_T + (lane_base << 16) + cc * 32, t);
                asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
                const uint32_t col0 = wi.cb * BN + cc * 32;
                if (row < n_cells && col0 + 31 > row) {
#pragma unroll
                    for (int k = 0; k < 32; ++k) {
                        const uint32_t col = col0 + k;
                        const int32_t tv = static_cast<int32_t>(t[k]);
                        const int32_t sv = (static_cast<int32_t>(q[k]) + tv) >> 2; // (4S - T + T) / 4
                        const int32_t dv = tv - sv;
                        if (col > row && col < n_cells) {
                            if (sv) {
                                atomicAdd(Srow + col, sign * sv);
                            }
                            if (dv) {
                                atomicAdd(Drow + col, sign * dv);
                            }
                        }
                    }
                }
            }
            asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
            mbar_arrive(tmem_empty);
            acc_phase ^= 1;
        }
    }
    __syncthreads();
    if (warp == 1) {
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(TMEM_COLS));
    }
}

} // namespace

// Upper-triangle output tiles in an order that keeps the tiles in flight (one per SM) inside a block
// of ~18 row blocks x 8 column blocks, so that a wave touches ~4 400 distinct operand rows instead of
// ~20 000: bands of 8 column blocks, row-block-major inside a band.
static int tile_list(sgpu_ctx *ctx, uint32_t N, uint32_t n_pad, const uint2 **d_tiles, uint32_t *n_tiles) {
    if (ctx->tile_cache && ctx->tile_cache_cells == N) {
        *d_tiles = static_cast<const uint2 *>(ctx->tile_cache);
        *n_tiles = ctx->tile_cache_n;
        return SGPU_OK;
    }
    constexpr uint32_t BAND = 8;
    std::vector<uint2> tiles;
    const uint32_t n_cb = n_pad / BN, n_rb = n_pad / BM;
    for (uint32_t band = 0; band < n_cb; band += BAND) {
        const uint32_t cb_end = std::min(n_cb, band + BAND);
        for (uint32_t rb = 0; rb < n_rb; ++rb) {
            for (uint32_t cb = band; cb < cb_end; ++cb) {
                // the tile intersects the upper triangle: its last column > its first row
                if (cb * BN + BN - 1 > rb * BM && rb * BM < N && cb * BN < N) {
                    tiles.push_back(make_uint2(rb, cb));
                }
            }
        }
    }
    if (ctx->tile_cache) {
        SGPU_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
        SGPU_CUDA(ctx, cudaFree(ctx->tile_cache));
        ctx->tile_cache = nullptr;
    }
    SGPU_CUDA(ctx, cudaMalloc(&ctx->tile_cache, std::max<size_t>(1, tiles.size()) * sizeof(uint2)));
    SGPU_CUDA(ctx, cudaMemcpy(ctx->tile_cache, tiles.data(), tiles.size() * sizeof(uint2), cudaMemcpyHostToDevice));
    ctx->tile_cache_cells = N;
    ctx->tile_cache_n = static_cast<uint32_t>(tiles.size());
    *d_tiles = static_cast<const uint2 *>(ctx->tile_cache);
    *n_tiles = ctx->tile_cache_n;
    return SGPU_OK;
}

int sgpu_gemm_counts(sgpu_ctx *ctx, const sgpu_pileup *p, const LinkResult &lr, sgpu_counts *c, uint64_t *n_pairs) {
    cudaStream_t st = ctx->stream;
    if (n_pairs) {
        *n_pairs = 0; // not enumerated on this path
    }
    const uint64_t P = p->n_loci;
    if (P == 0 || p->n_entries == 0 || c->n < 2) {
        return SGPU_OK;
    }
    const uint32_t N = c->n;
    const uint32_t n_pad = (N + BN - 1) / BN * BN;
    // panel: at most ~2 GB of staged counts / Hadamard planes
    uint64_t panel = (1ull << 31) / (4ull * n_pad) / LOCI_PER_KB * LOCI_PER_KB;
    panel = std::max<uint64_t>(panel, LOCI_PER_KB);
    panel = std::min<uint64_t>(panel, (P + LOCI_PER_KB - 1) / LOCI_PER_KB * LOCI_PER_KB);
    const uint64_t kbs_max = panel / LOCI_PER_KB;
    PanelLayout pl;
    pl.ck_shift = 5; // 32 k-blocks = 1 024 loci per chunk: 4 KB per cell, 33.8 MB per chunk at 8 192 cells
    if (const char *env = getenv("SECEDO_B200_CHUNK_SHIFT")) {
        pl.ck_shift = static_cast<uint32_t>(std::min(10, std::max(0, atoi(env))));
    }
    const uint32_t ck = 1u << pl.ck_shift;
    pl.cell_words = (ck | 1u) * 32u; // odd number of lines: consecutive cells do not alias in the L2 sets
    pl.chunk_words = static_cast<uint64_t>(n_pad) * pl.cell_words;
    const uint64_t n_chunks_max = (kbs_max + ck - 1) / ck;

    DevBuf<uint32_t> U;
    DevBuf<int> d_err;
    SGPU_CUDA(ctx, U.alloc(n_chunks_max * pl.chunk_words, ctx));
    SGPU_CUDA(ctx, d_err.alloc(2, ctx));
    SGPU_CUDA(ctx, cudaMemsetAsync(d_err.p, 0, 2 * sizeof(int), st));
    SGPU_TRACE(ctx, "gemm: alloc");

    // tensor map over U: [chunk][cell][ck x 128 bytes], box = 128 cells x 128 bytes, 128B swizzle
    CUtensorMap map;
    {
        const cuuint64_t gdim[3] = { static_cast<cuuint64_t>(ck) * KB_BYTES, n_pad, n_chunks_max };
        const cuuint64_t gstride[2] = { static_cast<cuuint64_t>(pl.cell_words) * 4, pl.chunk_words * 4 };
        const cuuint32_t box[3] = { KB_BYTES, 128, 1 };
        const cuuint32_t estr[3] = { 1, 1, 1 };
        // resolved through the runtime so that the library does not link libcuda.so (it must load, and
        // fail loudly in sgpu_init, on a machine without a driver)
        typedef CUresult (*encode_fn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *,
                                      const cuuint64_t *, const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave,
                                      CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
        void *fn = nullptr;
        cudaDriverEntryPointQueryResult qres;
        SGPU_CUDA(ctx, cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres));
        if (fn == nullptr || qres != cudaDriverEntryPointSuccess) {
            return sgpu_fail(ctx, SGPU_E_CUDA, "driver does not provide cuTensorMapEncodeTiled");
        }
        CUresult r = reinterpret_cast<encode_fn>(fn)(&map, CU_TENSOR_MAP_DATA_TYPE_UINT8, 3, U.p, gdim, gstride, box, estr,
                                            CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                                            CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (r != CUDA_SUCCESS) {
            return sgpu_fail(ctx, SGPU_E_CUDA, "cuTensorMapEncodeTiled failed (%d)", static_cast<int>(r));
        }
    }
    SGPU_CUDA(ctx, cudaFuncSetAttribute(syrk_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES));
    const uint2 *d_tiles = nullptr;
    uint32_t n_tiles = 0;
    SGPU_TRY(tile_list(ctx, N, n_pad, &d_tiles, &n_tiles));

    SGPU_TRACE(ctx, "gemm: tensor map + tiles");
    // CUDA events on the launching stream around the staging kernels and around the tcgen05 kernel
    std::vector<cudaEvent_t> evs;
    auto mark = [&]() {
        cudaEvent_t e;
        cudaEventCreateWithFlags(&e, cudaEventDefault);
        cudaEventRecord(e, st);
        evs.push_back(e);
    };
    const uint32_t sms = static_cast<uint32_t>(ctx->sm_count);
    const unsigned sp_blocks = static_cast<unsigned>(ceil_div_u64(lr.n_special ? lr.n_special : 1, 256));
    // zero / Hadamard-transform in place the chunks that hold the first kbs k-blocks
    auto clear_kbs = [&](uint32_t kbs) -> int {
        const uint64_t chunks = (kbs + ck - 1) / ck;
        SGPU_CUDA(ctx, cudaMemsetAsync(U.p, 0, chunks * pl.chunk_words * sizeof(uint32_t), st));
        return SGPU_OK;
    };
    auto transform_kbs = [&](uint32_t kbs) {
        const uint64_t chunks = (kbs + ck - 1) / ck;
        const uint64_t items = chunks * (static_cast<uint64_t>(n_pad) << pl.ck_shift) * 8; // a multiple of 256
        SGPU_LAUNCH(ctx, (transform_kernel<<<static_cast<unsigned>(items / 256), 256, 0, st>>>(U.p, kbs, n_pad, pl, d_err.p)));
    };
    // GEMM of the nl loci currently staged; sign -1 subtracts
    auto gemm_panel = [&](uint64_t nl, int sign) -> int {
        const uint32_t kbs = static_cast<uint32_t>((nl + LOCI_PER_KB - 1) / LOCI_PER_KB);
        // split K so that every SM has work even when there are few tiles
        uint32_t splits = 1;
        if (n_tiles < 2 * sms) {
            splits = static_cast<uint32_t>(std::min<uint64_t>(kbs, (2 * sms + n_tiles - 1) / n_tiles));
        }
        WorkList wl;
        wl.tiles = d_tiles;
        wl.n_tiles = n_tiles;
        wl.kbs = kbs;
        wl.ck_shift = pl.ck_shift;
        wl.per = (kbs + splits - 1) / splits;
        splits = (kbs + wl.per - 1) / wl.per; // no empty split
        wl.n_work = n_tiles * splits;
        const unsigned grid = static_cast<unsigned>(std::min<uint64_t>(wl.n_work, sms));
        mark(); // [3k+1] staging done, GEMM begins
        SGPU_LAUNCH(ctx, (syrk_kernel<<<grid, GEMM_THREADS, SMEM_BYTES, st>>>(map, wl, c->i32 + PLANE_S * c->nn, c->i32 + PLANE_D * c->nn, N, sign)));
        SGPU_CUDA(ctx, cudaGetLastError());
        mark(); // [3k+2] GEMM done
        ++ctx->n_syrk;
        return SGPU_OK;
    };
    const char *env_red = getenv("SECEDO_B200_STAGE_RED"); // experiment: fire-and-forget atomics (no range check)
    const bool stage_red = env_red && env_red[0] == '1';
    for (uint64_t l0 = 0; l0 < P; l0 += panel) {
        const uint64_t l1 = std::min<uint64_t>(P, l0 + panel);
        const uint64_t nl = l1 - l0;
        const uint32_t kbs = static_cast<uint32_t>((nl + LOCI_PER_KB - 1) / LOCI_PER_KB);
        mark(); // [3k] staging begins
        SGPU_TRY(clear_kbs(kbs));
        const unsigned grid = static_cast<unsigned>(std::min<uint64_t>(nl, static_cast<uint64_t>(sms) * 16));
        if (stage_red) {
            SGPU_LAUNCH(ctx, (stage_main_kernel<false><<<grid, 256, 0, st>>>(p->d_row_ptr, p->d_gid_base, lr.sp_bits.p, lr.gmap.p, lr.n_groups, N,
                                                                             static_cast<uint32_t>(l0), static_cast<uint32_t>(l1),
                                                                             static_cast<uint32_t>(l0), pl, U.p, d_err.p)));
        } else {
            SGPU_LAUNCH(ctx, (stage_main_kernel<true><<<grid, 256, 0, st>>>(p->d_row_ptr, p->d_gid_base, lr.sp_bits.p, lr.gmap.p, lr.n_groups, N,
                                                                            static_cast<uint32_t>(l0), static_cast<uint32_t>(l1),
                                                                            static_cast<uint32_t>(l0), pl, U.p, d_err.p)));
        }
        if (lr.n_special) {
            SGPU_LAUNCH(ctx, (stage_special_kernel<<<sp_blocks, 256, 0, st>>>(lr.sp_code.p, lr.sp_locus.p, lr.n_special, static_cast<uint32_t>(l0),
                                                                             static_cast<uint32_t>(l1), static_cast<uint32_t>(l0), pl, U.p, d_err.p)));
        }
        transform_kbs(kbs);
        SGPU_TRACE(ctx, "gemm: stage main");
        SGPU_TRY(gemm_panel(nl, +1));
    }
    SGPU_TRACE(ctx, "gemm: main panels");
    // The reference never compares two reads that both lie behind the per-chromosome cutoff K
    // (SURVEY F2). Those pairs only exist at the few loci behind the cutoff: subtract Z Z^T, Z = counts
    // of the tail reads alone at those loci, with the same kernel and sign -1.
    for (uint64_t t0 = 0; t0 < lr.n_tail_loci; t0 += panel) {
        const uint64_t nl = std::min<uint64_t>(lr.n_tail_loci - t0, panel);
        const uint32_t kbs = static_cast<uint32_t>((nl + LOCI_PER_KB - 1) / LOCI_PER_KB);
        mark();
        SGPU_TRY(clear_kbs(kbs));
        SGPU_LAUNCH(ctx, (stage_tail_kernel<<<static_cast<unsigned>(nl), 256, 0, st>>>(p->d_row_ptr, p->d_gid_base, lr.sp_bits.p, lr.gmap.p, lr.n_groups,
                                                                                       N, lr.tail_loci.p + t0, pl, U.p, d_err.p)));
        if (lr.n_special) {
            SGPU_LAUNCH(ctx, (stage_tail_special_kernel<<<sp_blocks, 256, 0, st>>>(lr.sp_code.p, lr.sp_locus.p, lr.n_special, lr.tail_loci.p + t0,
                                                                                  static_cast<uint32_t>(nl), pl, U.p, d_err.p)));
        }
        transform_kbs(kbs);
        SGPU_TRY(gemm_panel(nl, -1));
    }
    SGPU_TRACE(ctx, "gemm: tail panels");
    SGPU_CUDA(ctx, cudaMemcpyAsync(&ctx->h_scratch[2], d_err.p, sizeof(int), cudaMemcpyDeviceToHost, st));
    SGPU_CUDA(ctx, cudaStreamSynchronize(st));
    for (size_t k = 0; k + 2 < evs.size(); k += 3) {
        float a = 0.f, b = 0.f;
        cudaEventElapsedTime(&a, evs[k], evs[k + 1]);
        cudaEventElapsedTime(&b, evs[k + 1], evs[k + 2]);
        ctx->ms_stage += a;
        ctx->ms_syrk += b;
    }
    for (cudaEvent_t e : evs) {
        cudaEventDestroy(e);
    }
    const int err = static_cast<int>(ctx->h_scratch[2] & 0xFFFFFFFFu);
    if (err == SGPU_E_CELL_RANGE) {
        return sgpu_fail(ctx, err, "a group id is >= n_groups or maps to a cell >= num_cells (filter the pileup first)");
    }
    if (err != 0) {
        return sgpu_fail(ctx, SGPU_E_COUNT_RANGE, "more than 127 reads of one cell at one locus: outside the int8 GEMM path, use the scatter path");
    }
    return SGPU_OK;
}
