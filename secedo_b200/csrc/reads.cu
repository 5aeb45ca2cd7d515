// K2 — read linking, paired-end mate rule and tail cutoff.
//
// The reference walks the filtered loci sequentially and rebuilds reads in a hash map keyed by
// read id (similarity_matrix.cpp:342-403). What the similarity matrix needs from that walk is, per
// pileup entry: which read it belongs to, the read's index in order of first appearance, whether
// the entry survives the mate rule (:387-395), whether the read keeps >= 2 loci, and whether the
// read lies behind the per-chromosome cutoff K that the batching at :354-373 + the clear() at
// :407-408 imply. This file computes exactly that, data-parallel:
//
//   link_insert      open-addressing hash (chromosome, read id) -> smallest entry index
//   link_first       is this entry the first of its read?         (+ exclusive scan = read index)
//   link_reads       read index per entry, cell/start per read, extra-entry count per read
//   cand_fill/sort   entry lists of the reads with >= 2 entries (a small minority)
//   mate_rule        sequential replay of the insertion rules for those reads only
//   cutoff           one thread per chromosome, two-pointer sweep over loci
//   make_codes       packs (cell, base, tail, multi | dropped) per entry for the counting kernels
#include "common.cuh"

#include <algorithm>
#include <vector>

namespace {

constexpr uint64_t KEY_EMPTY = ~0ull;
constexpr int TB = 256;

__device__ __forceinline__ uint64_t mix64(uint64_t x) {
    x ^= x >> 33;
    x *= 0xff51afd7ed558ccdull;
    x ^= x >> 33;
    x *= 0xc4ceb9fe1a85ec53ull;
    x ^= x >> 33;
    return x;
}

__device__ __forceinline__ uint32_t chr_of_locus(const uint64_t *__restrict__ chr_ptr, uint32_t n_chr, uint64_t l) {
    uint32_t lo = 0, hi = n_chr; // largest c with chr_ptr[c] <= l
    while (hi - lo > 1) {
        uint32_t mid = (lo + hi) >> 1;
        if (chr_ptr[mid] <= l) {
            lo = mid;
        } else {
            hi = mid;
        }
    }
    return lo;
}

// eloc[e] = locus of entry e; lchr[l] = chromosome of locus l; checks strictly increasing positions
__global__ void __launch_bounds__(TB) locus_fill_kernel(const uint64_t *__restrict__ row_ptr,
                                                        const uint32_t *__restrict__ position,
                                                        const uint64_t *__restrict__ chr_ptr, uint32_t n_chr,
                                                        uint64_t n_loci, uint32_t *__restrict__ eloc,
                                                        uint8_t *__restrict__ lchr, int *__restrict__ err) {
    const int lane = threadIdx.x & 31;
    const uint64_t warps_total = static_cast<uint64_t>(gridDim.x) * (TB / 32);
    for (uint64_t l = static_cast<uint64_t>(blockIdx.x) * (TB / 32) + (threadIdx.x >> 5); l < n_loci; l += warps_total) {
        if (lane == 0) {
            const uint32_t c = chr_of_locus(chr_ptr, n_chr, l);
            lchr[l] = static_cast<uint8_t>(c);
            if (l > chr_ptr[c] && position[l] <= position[l - 1]) {
                atomicExch(err, SGPU_E_POSITIONS);
            }
        }
        const uint64_t e0 = row_ptr[l], e1 = row_ptr[l + 1];
        for (uint64_t e = e0 + lane; e < e1; e += 32) {
            eloc[e] = static_cast<uint32_t>(l);
        }
    }
}

__global__ void __launch_bounds__(TB) link_insert_kernel(const uint32_t *__restrict__ read_id,
                                                         const uint32_t *__restrict__ eloc,
                                                         const uint8_t *__restrict__ lchr, uint64_t n_entries,
                                                         uint64_t *__restrict__ keys, uint32_t *__restrict__ vals,
                                                         uint64_t mask, uint32_t *__restrict__ eslot) {
    const uint64_t e = static_cast<uint64_t>(blockIdx.x) * TB + threadIdx.x;
    if (e >= n_entries) {
        return;
    }
    const uint64_t key = (static_cast<uint64_t>(lchr[eloc[e]]) << 32) | read_id[e];
    uint64_t slot = mix64(key) & mask;
    for (;;) {
        uint64_t cur = keys[slot];
        if (cur == KEY_EMPTY) {
            cur = atomicCAS(reinterpret_cast<unsigned long long *>(&keys[slot]), KEY_EMPTY, key);
            if (cur == KEY_EMPTY) {
                break;
            }
        }
        if (cur == key) {
            break;
        }
        slot = (slot + 1) & mask;
    }
    atomicMin(&vals[slot], static_cast<uint32_t>(e));
    eslot[e] = static_cast<uint32_t>(slot);
}

// ---- windowed linking ----------------------------------------------------------------------------
// A fragment spans less than L bp, so all entries of a read lie in the loci within L bp after the
// locus of its first entry. One CTA per "owner" locus builds a hash table of that locus' read ids in
// shared memory (value = smallest entry index carrying the id) and streams the entries of the
// following loci inside the window against it; every entry ends up with efirst[e] = the smallest
// entry index of its read. No global-memory hash table: reads are 4 B/entry per pass, coalesced.
constexpr uint32_t ID_EMPTY = 0xFFFFFFFFu;
constexpr int WIN_THREADS = 512;
constexpr uint32_t WIN_MAX_SLOTS = 16384; // 128 KB of shared memory

__global__ void __launch_bounds__(TB) max_locus_size_kernel(const uint64_t *__restrict__ row_ptr, uint64_t n_loci,
                                                            unsigned int *__restrict__ out) {
    const uint64_t l = static_cast<uint64_t>(blockIdx.x) * TB + threadIdx.x;
    unsigned int n = 0;
    if (l < n_loci) {
        n = static_cast<unsigned int>(min(row_ptr[l + 1] - row_ptr[l], static_cast<uint64_t>(0xFFFFFFFFu)));
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        n = max(n, __shfl_xor_sync(0xffffffffu, n, o));
    }
    if ((threadIdx.x & 31) == 0 && n) {
        atomicMax(out, n);
    }
}

__global__ void __launch_bounds__(TB) iota_kernel(uint32_t *__restrict__ a, uint64_t n) {
    const uint64_t i = static_cast<uint64_t>(blockIdx.x) * TB + threadIdx.x;
    if (i < n) {
        a[i] = static_cast<uint32_t>(i);
    }
}

__device__ __forceinline__ uint32_t hash32(uint32_t x) {
    x ^= x >> 16;
    x *= 0x7feb352du;
    x ^= x >> 15;
    x *= 0x846ca68bu;
    x ^= x >> 16;
    return x;
}

__global__ void __launch_bounds__(WIN_THREADS) link_window_kernel(
        const uint64_t *__restrict__ row_ptr, const uint32_t *__restrict__ position, const uint32_t *__restrict__ read_id,
        const uint8_t *__restrict__ lchr, const uint64_t *__restrict__ chr_ptr, uint64_t n_loci, uint32_t L,
        uint32_t slots /* power of two */, uint32_t *__restrict__ efirst) {
    extern __shared__ uint32_t s_tab[]; // keys[slots], vals[slots]
    uint32_t *keys = s_tab, *vals = s_tab + slots;
    const uint32_t mask = slots - 1;
    for (uint64_t lo = blockIdx.x; lo < n_loci; lo += gridDim.x) {
        const uint64_t e0 = row_ptr[lo], e1 = row_ptr[lo + 1];
        const uint64_t chr_end = chr_ptr[lchr[lo] + 1];
        const uint32_t p0 = position[lo];
        __syncthreads(); // previous owner's probes are done
        for (uint32_t i = threadIdx.x; i < slots; i += WIN_THREADS) {
            keys[i] = ID_EMPTY;
            vals[i] = 0xFFFFFFFFu;
        }
        __syncthreads();
        // build: id -> smallest entry index at the owner locus
        for (uint64_t e = e0 + threadIdx.x; e < e1; e += WIN_THREADS) {
            const uint32_t id = read_id[e];
            uint32_t s = hash32(id) & mask;
            for (;;) {
                const uint32_t prev = atomicCAS(&keys[s], ID_EMPTY, id);
                if (prev == ID_EMPTY || prev == id) {
                    break;
                }
                s = (s + 1) & mask;
            }
            atomicMin(&vals[s], static_cast<uint32_t>(e));
        }
        __syncthreads();
        // the owner's own entries (two entries of one read at one locus: overlapping mates), then the
        // loci of the window
        for (uint64_t l = lo; l < chr_end; ++l) {
            if (l > lo && (position[l] - p0 >= L)) {
                break;
            }
            const uint64_t a0 = row_ptr[l], a1 = row_ptr[l + 1];
            for (uint64_t e = a0 + threadIdx.x; e < a1; e += WIN_THREADS) {
                const uint32_t id = read_id[e];
                uint32_t s = hash32(id) & mask;
                for (;;) {
                    const uint32_t k = keys[s];
                    if (k == id) {
                        const uint32_t m = vals[s];
                        if (m < static_cast<uint32_t>(e)) {
                            atomicMin(&efirst[e], m);
                        }
                        break;
                    }
                    if (k == ID_EMPTY) {
                        break;
                    }
                    s = (s + 1) & mask;
                }
            }
        }
    }
}

// isfirst[e] and the check that first-entry pointers are idempotent (a chain means a read id
// spanning >= L through intermediate loci)
__global__ void __launch_bounds__(TB) link_finish_kernel(const uint32_t *__restrict__ efirst, uint64_t n_entries,
                                                         uint8_t *__restrict__ isfirst, int *__restrict__ err) {
    const uint64_t e = static_cast<uint64_t>(blockIdx.x) * TB + threadIdx.x;
    if (e >= n_entries) {
        return;
    }
    const uint32_t f = efirst[e];
    isfirst[e] = f == static_cast<uint32_t>(e) ? 1 : 0;
    if (efirst[f] != f) {
        atomicExch(err, SGPU_E_FRAGMENT_SPAN);
    }
}

// efirst[e] (in place over eslot) = first entry of e's read; isfirst[e]
__global__ void __launch_bounds__(TB) link_first_kernel(uint32_t *__restrict__ eslot_efirst,
                                                        const uint32_t *__restrict__ vals, uint64_t n_entries,
                                                        uint8_t *__restrict__ isfirst) {
    const uint64_t e = static_cast<uint64_t>(blockIdx.x) * TB + threadIdx.x;
    if (e >= n_entries) {
        return;
    }
    const uint32_t f = vals[eslot_efirst[e]];
    eslot_efirst[e] = f;
    isfirst[e] = f == static_cast<uint32_t>(e) ? 1 : 0;
}

__global__ void __launch_bounds__(TB) link_reads_kernel(const uint32_t *__restrict__ efirst,
                                                        const uint64_t *__restrict__ rscan,
                                                        const uint16_t *__restrict__ gid_base,
                                                        const uint32_t *__restrict__ eloc,
                                                        const uint32_t *__restrict__ gmap, uint32_t n_groups,
                                                        uint32_t num_cells, uint64_t n_entries,
                                                        uint32_t *__restrict__ eread, uint32_t *__restrict__ r_cell,
                                                        uint32_t *__restrict__ r_startloc,
                                                        uint32_t *__restrict__ r_extra, int *__restrict__ err) {
    const uint64_t e = static_cast<uint64_t>(blockIdx.x) * TB + threadIdx.x;
    if (e >= n_entries) {
        return;
    }
    const uint32_t f = efirst[e];
    const uint32_t r = static_cast<uint32_t>(rscan[f]);
    eread[e] = r;
    if (f == static_cast<uint32_t>(e)) {
        // the read's cell is fixed by its first entry (similarity_matrix.cpp:379, :208-209)
        const uint32_t gid = gid_base[e] >> 2;
        uint32_t cell = 0;
        if (gid >= n_groups || (cell = gmap[gid]) >= num_cells) {
            atomicExch(err, SGPU_E_CELL_RANGE);
            cell = 0;
        }
        r_cell[r] = cell;
        r_startloc[r] = eloc[e];
    } else {
        atomicAdd(&r_extra[r], 1u);
    }
}

// cand[r] = read has >= 2 entries; clen[r] = its entry count (0 for non candidates)
__global__ void __launch_bounds__(TB) cand_flag_kernel(const uint32_t *__restrict__ r_extra, uint64_t n_reads,
                                                       uint8_t *__restrict__ cand, uint32_t *__restrict__ clen) {
    const uint64_t r = static_cast<uint64_t>(blockIdx.x) * TB + threadIdx.x;
    if (r >= n_reads) {
        return;
    }
    const uint32_t x = r_extra[r];
    cand[r] = x ? 1 : 0;
    clen[r] = x ? x + 1 : 0;
}

__global__ void __launch_bounds__(TB) cand_fill_kernel(const uint32_t *__restrict__ eread,
                                                       const uint8_t *__restrict__ cand,
                                                       const uint64_t *__restrict__ c_off /* per read, scan of clen */,
                                                       uint64_t n_entries, uint32_t *__restrict__ cursor /* per read */,
                                                       uint32_t *__restrict__ c_list) {
    const uint64_t e = static_cast<uint64_t>(blockIdx.x) * TB + threadIdx.x;
    if (e >= n_entries) {
        return;
    }
    const uint32_t r = eread[e];
    if (cand[r]) {
        const uint32_t k = atomicAdd(&cursor[r], 1u);
        c_list[c_off[r] + k] = static_cast<uint32_t>(e);
    }
}

// One thread per candidate read: sort its entries (entry order = the order in which the reference
// meets them), replay the insertion rules (similarity_matrix.cpp:383-402) and leave the stored
// (locus, base) list in place. nst[r] = number of loci the read keeps.
__global__ void __launch_bounds__(TB) mate_rule_kernel(const uint8_t *__restrict__ cand,
                                                       const uint64_t *__restrict__ c_off, uint64_t n_reads,
                                                       uint32_t *__restrict__ c_list /* in: entries, out: stored loci */,
                                                       uint8_t *__restrict__ c_base, const uint32_t *__restrict__ eloc,
                                                       const uint16_t *__restrict__ gid_base,
                                                       const uint32_t *__restrict__ position,
                                                       const uint32_t *__restrict__ r_startloc, uint32_t L,
                                                       uint8_t *__restrict__ edrop, uint32_t *__restrict__ nst,
                                                       int *__restrict__ err) {
    const uint64_t r = static_cast<uint64_t>(blockIdx.x) * TB + threadIdx.x;
    if (r >= n_reads || !cand[r]) {
        return;
    }
    const uint64_t o = c_off[r];
    const uint32_t n = static_cast<uint32_t>(c_off[r + 1] - o);
    uint32_t *lst = c_list + o;
    for (uint32_t i = 1; i < n; ++i) { // insertion sort, n is tiny
        const uint32_t v = lst[i];
        uint32_t j = i;
        while (j > 0 && lst[j - 1] > v) {
            lst[j] = lst[j - 1];
            --j;
        }
        lst[j] = v;
    }
    const uint32_t start_pos = position[r_startloc[r]];
    uint32_t n_st = 0;
    uint32_t last_entry = 0; // entry that produced the last stored element
    for (uint32_t i = 0; i < n; ++i) {
        const uint32_t e = lst[i];
        const uint32_t loc = eloc[e];
        const uint8_t base = gid_base[e] & 3;
        if (position[loc] - start_pos >= L) {
            atomicExch(err, SGPU_E_FRAGMENT_SPAN);
        }
        if (n_st > 0 && lst[n_st - 1] == loc) { // read.pos.back() == pd.position (:387)
            if (c_base[o + n_st - 1] != base) { // mates disagree: remove the stored base too (:390-393)
                edrop[last_entry] = 1;
                --n_st;
                // the element below (if any) was produced by an entry we no longer track; it can only
                // be popped by a further entry at ITS locus, which cannot come (loci are increasing)
            }
            edrop[e] = 1;
            continue;
        }
        // lst[0..n_st) is overwritten with stored loci; i >= n_st always holds
        lst[n_st] = loc;
        c_base[o + n_st] = base;
        last_entry = e;
        ++n_st;
    }
    nst[r] = n_st;
}

// readbase[l] = number of reads created before locus l (l = 0..n_loci)
__global__ void __launch_bounds__(TB) readbase_kernel(const uint64_t *__restrict__ row_ptr,
                                                      const uint64_t *__restrict__ rscan, uint64_t n_loci,
                                                      uint64_t *__restrict__ readbase) {
    const uint64_t l = static_cast<uint64_t>(blockIdx.x) * TB + threadIdx.x;
    if (l <= n_loci) {
        readbase[l] = rscan[row_ptr[l]];
    }
}

// SURVEY Appendix A.4. One CTA per chromosome: for a chunk of loci all threads compute
// u(l) = number of reads whose start + L <= position[l] (a binary search over the positions, reads are
// created locus by locus), then one thread replays the reference's batching rule over the chunk.
constexpr int CUT_THREADS = 256;
constexpr int CUT_CHUNK = 1024;
__global__ void __launch_bounds__(CUT_THREADS) cutoff_kernel(const uint64_t *__restrict__ chr_ptr, uint32_t n_chr,
                                                             const uint32_t *__restrict__ position,
                                                             const uint64_t *__restrict__ readbase, uint32_t L,
                                                             uint32_t num_threads, uint64_t *__restrict__ Kglob,
                                                             uint64_t *__restrict__ tail_locus) {
    __shared__ uint64_t s_u[CUT_CHUNK];
    __shared__ uint32_t s_j[CUT_CHUNK];
    const uint32_t c = blockIdx.x;
    if (c >= n_chr) {
        return;
    }
    const uint64_t l0 = chr_ptr[c], l1 = chr_ptr[c + 1];
    const uint64_t rb0 = readbase[l0];
    const uint64_t need = 4ull * num_threads; // BATCH_SIZE * num_threads (similarity_matrix.cpp:354-356)
    uint64_t front = 0, jk = l0;              // used by thread 0 only
    for (uint64_t base = l0; base < l1; base += CUT_CHUNK) {
        const uint32_t n = static_cast<uint32_t>(min(static_cast<uint64_t>(CUT_CHUNK), l1 - base));
        for (uint32_t i = threadIdx.x; i < n; i += CUT_THREADS) {
            const uint64_t l = base + i;
            const uint64_t p = position[l];
            // j = first locus in [l0, l] whose reads are NOT complete at p: position[j] + L > p
            uint64_t lo = l0, hi = l;
            while (lo < hi) {
                const uint64_t mid = (lo + hi) >> 1;
                if (static_cast<uint64_t>(position[mid]) + L > p) {
                    hi = mid;
                } else {
                    lo = mid + 1;
                }
            }
            s_u[i] = readbase[lo] - rb0; // reads with start + L <= p, all created before locus l
            s_j[i] = static_cast<uint32_t>(lo - l0);
        }
        __syncthreads();
        if (threadIdx.x == 0) {
            for (uint32_t i = 0; i < n; ++i) {
                const uint64_t u = s_u[i];
                if (u > front && u - front >= need) {
                    front = u;
                    jk = l0 + s_j[i];
                }
            }
        }
        __syncthreads();
    }
    if (threadIdx.x == 0) {
        Kglob[c] = rb0 + front;
        tail_locus[c] = jk; // reads with index >= K are exactly those created at loci >= jk
    }
}

__global__ void __launch_bounds__(TB) make_codes_kernel(
        const uint32_t *__restrict__ eread, const uint32_t *__restrict__ efirst, const uint32_t *__restrict__ eloc,
        const uint8_t *__restrict__ lchr, const uint16_t *__restrict__ gid_base, const uint32_t *__restrict__ r_cell,
        const uint8_t *__restrict__ cand, const uint32_t *__restrict__ nst, const uint8_t *__restrict__ edrop,
        const uint64_t *__restrict__ Kglob, uint64_t n_entries, uint32_t *__restrict__ code,
        unsigned long long *__restrict__ stats /* [0]=dropped [1]=multi entries */) {
    const uint64_t e = static_cast<uint64_t>(blockIdx.x) * TB + threadIdx.x;
    uint32_t dropped = 0, multi = 0;
    if (e < n_entries) {
        const uint32_t r = eread[e];
        if (edrop[e]) {
            code[e] = CODE_DROPPED;
            dropped = 1;
        } else {
            const uint32_t cell = r_cell[r];
            const uint32_t base = gid_base[e] & 3u;
            const uint32_t tail = r >= Kglob[lchr[eloc[e]]] ? 1u : 0u;
            multi = (cand[r] && nst[r] >= 2) ? 1u : 0u;
            code[e] = (cell << 4) | (base << 2) | (tail << 1) | multi;
        }
    }
    const uint32_t bd = __ballot_sync(0xffffffffu, dropped), bm = __ballot_sync(0xffffffffu, multi);
    if ((threadIdx.x & 31) == 0) {
        if (bd) {
            atomicAdd(&stats[0], static_cast<unsigned long long>(__popc(bd)));
        }
        if (bm) {
            atomicAdd(&stats[1], static_cast<unsigned long long>(__popc(bm)));
        }
    }
    (void)efirst;
}

// multi reads: flag + stored length
__global__ void __launch_bounds__(TB) multi_flag_kernel(const uint8_t *__restrict__ cand, const uint32_t *__restrict__ nst,
                                                        uint64_t n_reads, uint8_t *__restrict__ mflag,
                                                        uint32_t *__restrict__ mlen) {
    const uint64_t r = static_cast<uint64_t>(blockIdx.x) * TB + threadIdx.x;
    if (r >= n_reads) {
        return;
    }
    const bool m = cand[r] && nst[r] >= 2;
    mflag[r] = m ? 1 : 0;
    mlen[r] = m ? nst[r] : 0;
}

__global__ void __launch_bounds__(TB) multi_copy_kernel(const uint8_t *__restrict__ mflag,
                                                        const uint64_t *__restrict__ midx /* scan of mflag */,
                                                        const uint64_t *__restrict__ moff_r /* scan of mlen, per read */,
                                                        const uint64_t *__restrict__ c_off, const uint32_t *__restrict__ c_list,
                                                        const uint8_t *__restrict__ c_base, const uint32_t *__restrict__ nst,
                                                        uint64_t n_reads, uint32_t *__restrict__ r_multi,
                                                        uint64_t *__restrict__ m_off, uint32_t *__restrict__ m_locus,
                                                        uint8_t *__restrict__ m_base) {
    const uint64_t r = static_cast<uint64_t>(blockIdx.x) * TB + threadIdx.x;
    if (r > n_reads) {
        return;
    }
    if (r == n_reads) {
        m_off[midx[n_reads]] = moff_r[n_reads];
        return;
    }
    if (!mflag[r]) {
        r_multi[r] = 0xFFFFFFFFu;
        return;
    }
    const uint64_t m = midx[r], dst = moff_r[r], src = c_off[r];
    r_multi[r] = static_cast<uint32_t>(m);
    m_off[m] = dst;
    for (uint32_t i = 0; i < nst[r]; ++i) {
        m_locus[dst + i] = c_list[src + i];
        m_base[dst + i] = c_base[src + i];
    }
}

unsigned blocks_for(uint64_t n) { return static_cast<unsigned>(ceil_div_u64(n ? n : 1, TB)); }

} // namespace

int sgpu_link_reads(sgpu_ctx *ctx, const sgpu_pileup *p, uint32_t num_cells, uint32_t L,
                    const uint32_t *h_group_id_to_pos, uint32_t n_groups, uint32_t num_threads,
                    LinkResult *out) {
    cudaStream_t st = ctx->stream;
    const uint64_t E = p->n_entries, P = p->n_loci;
    if (E >= 0x7FFFFFF0ull || P >= 0x7FFFFFF0ull) {
        return sgpu_fail(ctx, SGPU_E_ARG, "more than 2^31 entries in one call: pass the pileup in batches of chromosomes");
    }
    if (p->n_chr > 255) {
        return sgpu_fail(ctx, SGPU_E_ARG, "more than 255 chromosomes in one call");
    }
    if (num_threads == 0) {
        return sgpu_fail(ctx, SGPU_E_ARG, "num_threads must be >= 1");
    }
    if (num_cells >= (1u << 27)) {
        return sgpu_fail(ctx, SGPU_E_ARG, "num_cells too large");
    }
    out->n_reads = out->n_multi = out->n_dropped = out->n_tail = out->n_multi_entries = 0;
    SGPU_CUDA(ctx, out->code.alloc(E, st));
    SGPU_CUDA(ctx, out->eread.alloc(E, st));
    SGPU_CUDA(ctx, out->eloc.alloc(E, st));
    if (E == 0) {
        SGPU_CUDA(ctx, out->m_off.alloc(1, st));
        SGPU_CUDA(ctx, cudaMemsetAsync(out->m_off.p, 0, sizeof(uint64_t), st));
        return SGPU_OK;
    }

    DevBuf<uint8_t> lchr, isfirst, edrop;
    DevBuf<uint32_t> gmap, efirst, vals;
    DevBuf<uint64_t> keys, rscan, readbase, Kglob, tail_locus;
    DevBuf<int> d_err;
    DevBuf<unsigned long long> d_stats;
    SGPU_CUDA(ctx, lchr.alloc(P, st));
    SGPU_CUDA(ctx, isfirst.alloc(E, st));
    SGPU_CUDA(ctx, edrop.alloc(E, st));
    SGPU_CUDA(ctx, gmap.alloc(n_groups ? n_groups : 1, st));
    SGPU_CUDA(ctx, efirst.alloc(E, st));
    SGPU_CUDA(ctx, rscan.alloc(E + 1, st));
    SGPU_CUDA(ctx, readbase.alloc(P + 1, st));
    SGPU_CUDA(ctx, Kglob.alloc(p->n_chr, st));
    SGPU_CUDA(ctx, tail_locus.alloc(p->n_chr, st));
    SGPU_CUDA(ctx, d_err.alloc(1, st));
    SGPU_CUDA(ctx, d_stats.alloc(2, st));
    SGPU_CUDA(ctx, cudaMemsetAsync(d_err.p, 0, sizeof(int), st));
    SGPU_CUDA(ctx, cudaMemsetAsync(d_stats.p, 0, 2 * sizeof(unsigned long long), st));
    SGPU_CUDA(ctx, cudaMemsetAsync(edrop.p, 0, E, st));
    SGPU_CUDA(ctx, cudaMemcpyAsync(gmap.p, h_group_id_to_pos, n_groups * sizeof(uint32_t), cudaMemcpyHostToDevice, st));

    const unsigned locus_grid = static_cast<unsigned>(std::min<uint64_t>(ceil_div_u64(P, TB / 32), static_cast<uint64_t>(ctx->sm_count) * 32));
    SGPU_LAUNCH(ctx, (locus_fill_kernel<<<locus_grid, TB, 0, st>>>(p->d_row_ptr, p->d_position, p->d_chr_ptr, p->n_chr, P, out->eloc.p, lchr.p, d_err.p)));

    // ---- first entry of every read: windowed shared-memory hashing, or (loci too large for shared
    // memory) a global open-addressing hash (chromosome, read id) -> first entry
    DevBuf<unsigned int> d_maxn;
    SGPU_CUDA(ctx, d_maxn.alloc(1, st));
    SGPU_CUDA(ctx, cudaMemsetAsync(d_maxn.p, 0, sizeof(unsigned int), st));
    SGPU_LAUNCH(ctx, (max_locus_size_kernel<<<blocks_for(P), TB, 0, st>>>(p->d_row_ptr, P, d_maxn.p)));
    SGPU_CUDA(ctx, cudaMemcpyAsync(&ctx->h_scratch[0], d_maxn.p, sizeof(unsigned int), cudaMemcpyDeviceToHost, st));
    SGPU_CUDA(ctx, cudaStreamSynchronize(st));
    const uint32_t max_n = static_cast<uint32_t>(ctx->h_scratch[0] & 0xFFFFFFFFu);
    uint32_t slots = 1024;
    while (slots < max_n + max_n / 2) {
        slots <<= 1;
    }
    if (slots <= WIN_MAX_SLOTS) {
        const size_t smem = static_cast<size_t>(slots) * 2 * sizeof(uint32_t);
        SGPU_CUDA(ctx, cudaFuncSetAttribute(link_window_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)));
        SGPU_LAUNCH(ctx, (iota_kernel<<<blocks_for(E), TB, 0, st>>>(efirst.p, E)));
        const unsigned per_sm = static_cast<unsigned>(std::max<size_t>(1, std::min<size_t>(4, (200 * 1024) / smem)));
        const unsigned wgrid = static_cast<unsigned>(std::min<uint64_t>(P, static_cast<uint64_t>(ctx->sm_count) * per_sm));
        SGPU_LAUNCH(ctx, (link_window_kernel<<<wgrid, WIN_THREADS, smem, st>>>(p->d_row_ptr, p->d_position, p->d_read_id, lchr.p,
                                                                              p->d_chr_ptr, P, L, slots, efirst.p)));
        SGPU_LAUNCH(ctx, (link_finish_kernel<<<blocks_for(E), TB, 0, st>>>(efirst.p, E, isfirst.p, d_err.p)));
    } else {
        uint64_t cap = 1024;
        while (cap < 2 * E) {
            cap <<= 1;
        }
        SGPU_CUDA(ctx, keys.alloc(cap, st));
        SGPU_CUDA(ctx, vals.alloc(cap, st));
        SGPU_CUDA(ctx, cudaMemsetAsync(keys.p, 0xFF, cap * sizeof(uint64_t), st));
        SGPU_CUDA(ctx, cudaMemsetAsync(vals.p, 0xFF, cap * sizeof(uint32_t), st));
        SGPU_LAUNCH(ctx, (link_insert_kernel<<<blocks_for(E), TB, 0, st>>>(p->d_read_id, out->eloc.p, lchr.p, E, keys.p, vals.p, cap - 1, efirst.p)));
        SGPU_LAUNCH(ctx, (link_first_kernel<<<blocks_for(E), TB, 0, st>>>(efirst.p, vals.p, E, isfirst.p)));
        keys.release();
        vals.release();
    }
    SGPU_CUDA(ctx, cudaGetLastError());
    SGPU_TRY(sgpu_scan_u8_u64(ctx, isfirst.p, rscan.p, E));
    SGPU_CUDA(ctx, cudaMemcpyAsync(&ctx->h_scratch[0], rscan.p + E, sizeof(uint64_t), cudaMemcpyDeviceToHost, st));
    SGPU_CUDA(ctx, cudaStreamSynchronize(st));
    const uint64_t R = ctx->h_scratch[0];
    out->n_reads = R;

    // ---- per-read tables
    DevBuf<uint32_t> r_cell, r_startloc, r_extra, clen, nst, cursor;
    DevBuf<uint8_t> cand;
    DevBuf<uint64_t> c_off;
    SGPU_CUDA(ctx, r_cell.alloc(R, st));
    SGPU_CUDA(ctx, r_startloc.alloc(R, st));
    SGPU_CUDA(ctx, r_extra.alloc(R, st));
    SGPU_CUDA(ctx, clen.alloc(R, st));
    SGPU_CUDA(ctx, nst.alloc(R, st));
    SGPU_CUDA(ctx, cursor.alloc(R, st));
    SGPU_CUDA(ctx, cand.alloc(R, st));
    SGPU_CUDA(ctx, c_off.alloc(R + 1, st));
    SGPU_CUDA(ctx, cudaMemsetAsync(r_extra.p, 0, R * sizeof(uint32_t), st));
    SGPU_CUDA(ctx, cudaMemsetAsync(cursor.p, 0, R * sizeof(uint32_t), st));
    SGPU_CUDA(ctx, cudaMemsetAsync(nst.p, 0, R * sizeof(uint32_t), st));
    SGPU_LAUNCH(ctx, (link_reads_kernel<<<blocks_for(E), TB, 0, st>>>(efirst.p, rscan.p, p->d_gid_base, out->eloc.p, gmap.p, n_groups, num_cells, E,
                                                   out->eread.p, r_cell.p, r_startloc.p, r_extra.p, d_err.p)));
    SGPU_LAUNCH(ctx, (cand_flag_kernel<<<blocks_for(R), TB, 0, st>>>(r_extra.p, R, cand.p, clen.p)));
    SGPU_CUDA(ctx, cudaGetLastError());
    SGPU_TRY(sgpu_scan_u32_u64(ctx, clen.p, c_off.p, R));
    SGPU_CUDA(ctx, cudaMemcpyAsync(&ctx->h_scratch[0], c_off.p + R, sizeof(uint64_t), cudaMemcpyDeviceToHost, st));
    SGPU_CUDA(ctx, cudaStreamSynchronize(st));
    const uint64_t CE = ctx->h_scratch[0];

    DevBuf<uint32_t> c_list;
    DevBuf<uint8_t> c_base;
    SGPU_CUDA(ctx, c_list.alloc(CE ? CE : 1, st));
    SGPU_CUDA(ctx, c_base.alloc(CE ? CE : 1, st));
    if (CE) {
        SGPU_LAUNCH(ctx, (cand_fill_kernel<<<blocks_for(E), TB, 0, st>>>(out->eread.p, cand.p, c_off.p, E, cursor.p, c_list.p)));
        SGPU_LAUNCH(ctx, (mate_rule_kernel<<<blocks_for(R), TB, 0, st>>>(cand.p, c_off.p, R, c_list.p, c_base.p, out->eloc.p, p->d_gid_base,
                                                      p->d_position, r_startloc.p, L, edrop.p, nst.p, d_err.p)));
    }
    SGPU_LAUNCH(ctx, (// ---- cutoff K per chromosome
    readbase_kernel<<<blocks_for(P + 1), TB, 0, st>>>(p->d_row_ptr, rscan.p, P, readbase.p)));
    SGPU_LAUNCH(ctx, (cutoff_kernel<<<p->n_chr, CUT_THREADS, 0, st>>>(p->d_chr_ptr, p->n_chr, p->d_position, readbase.p, L, num_threads, Kglob.p, tail_locus.p)));
    SGPU_LAUNCH(ctx, (// ---- codes
    make_codes_kernel<<<blocks_for(E), TB, 0, st>>>(out->eread.p, efirst.p, out->eloc.p, lchr.p, p->d_gid_base, r_cell.p, cand.p,
                                                   nst.p, edrop.p, Kglob.p, E, out->code.p, d_stats.p)));
    SGPU_CUDA(ctx, cudaGetLastError());

    // ---- multi-locus read tables
    DevBuf<uint8_t> mflag;
    DevBuf<uint32_t> mlen;
    DevBuf<uint64_t> midx, moff_r;
    SGPU_CUDA(ctx, mflag.alloc(R, st));
    SGPU_CUDA(ctx, mlen.alloc(R, st));
    SGPU_CUDA(ctx, midx.alloc(R + 1, st));
    SGPU_CUDA(ctx, moff_r.alloc(R + 1, st));
    SGPU_LAUNCH(ctx, (multi_flag_kernel<<<blocks_for(R), TB, 0, st>>>(cand.p, nst.p, R, mflag.p, mlen.p)));
    SGPU_TRY(sgpu_scan_u8_u64(ctx, mflag.p, midx.p, R));
    SGPU_TRY(sgpu_scan_u32_u64(ctx, mlen.p, moff_r.p, R));
    // scalars to the host: n_multi, total stored, error, stats, K per chromosome, reads per chromosome
    std::vector<uint64_t> h_K(p->n_chr), h_rb(p->n_chr + 1), h_tl(p->n_chr);
    SGPU_CUDA(ctx, cudaMemcpyAsync(h_tl.data(), tail_locus.p, p->n_chr * sizeof(uint64_t), cudaMemcpyDeviceToHost, st));
    SGPU_CUDA(ctx, cudaMemcpyAsync(&ctx->h_scratch[0], midx.p + R, sizeof(uint64_t), cudaMemcpyDeviceToHost, st));
    SGPU_CUDA(ctx, cudaMemcpyAsync(&ctx->h_scratch[1], moff_r.p + R, sizeof(uint64_t), cudaMemcpyDeviceToHost, st));
    SGPU_CUDA(ctx, cudaMemcpyAsync(&ctx->h_scratch[2], d_err.p, sizeof(int), cudaMemcpyDeviceToHost, st));
    SGPU_CUDA(ctx, cudaMemcpyAsync(&ctx->h_scratch[4], d_stats.p, 2 * sizeof(uint64_t), cudaMemcpyDeviceToHost, st));
    SGPU_CUDA(ctx, cudaMemcpyAsync(h_K.data(), Kglob.p, p->n_chr * sizeof(uint64_t), cudaMemcpyDeviceToHost, st));
    SGPU_CUDA(ctx, cudaStreamSynchronize(st));
    const int err = static_cast<int>(ctx->h_scratch[2] & 0xFFFFFFFFu);
    if (err == SGPU_E_CELL_RANGE) {
        return sgpu_fail(ctx, err, "a group id is >= n_groups or maps to a cell >= num_cells (filter the pileup first)");
    }
    if (err == SGPU_E_POSITIONS) {
        return sgpu_fail(ctx, err, "positions are not strictly increasing inside a chromosome");
    }
    if (err == SGPU_E_FRAGMENT_SPAN) {
        return sgpu_fail(ctx, err, "a read id spans >= max_fragment_length (%u): undefined in the reference", L);
    }
    const uint64_t NM = ctx->h_scratch[0], MS = ctx->h_scratch[1];
    out->n_multi = NM;
    out->n_dropped = ctx->h_scratch[4];
    out->n_multi_entries = ctx->h_scratch[5];
    // tail reads = reads of each chromosome with index >= K
    for (uint32_t c = 0; c < p->n_chr; ++c) {
        SGPU_CUDA(ctx, cudaMemcpyAsync(&h_rb[c + 1], readbase.p + p->h_chr_ptr[c + 1], sizeof(uint64_t), cudaMemcpyDeviceToHost, st));
    }
    SGPU_CUDA(ctx, cudaStreamSynchronize(st));
    std::vector<uint32_t> h_tail_loci;
    for (uint32_t c = 0; c < p->n_chr; ++c) {
        out->n_tail += h_rb[c + 1] - h_K[c];
        for (uint64_t l = h_tl[c]; l < p->h_chr_ptr[c + 1]; ++l) {
            h_tail_loci.push_back(static_cast<uint32_t>(l));
        }
    }
    out->n_tail_loci = h_tail_loci.size();
    SGPU_CUDA(ctx, out->tail_loci.alloc(h_tail_loci.size() ? h_tail_loci.size() : 1, st));
    if (!h_tail_loci.empty()) {
        SGPU_CUDA(ctx, cudaMemcpyAsync(out->tail_loci.p, h_tail_loci.data(), h_tail_loci.size() * sizeof(uint32_t), cudaMemcpyHostToDevice, st));
        SGPU_CUDA(ctx, cudaStreamSynchronize(st));
    }

    SGPU_CUDA(ctx, out->r_multi.alloc(R, st));
    SGPU_CUDA(ctx, out->m_off.alloc(NM + 1, st));
    SGPU_CUDA(ctx, out->m_locus.alloc(MS ? MS : 1, st));
    SGPU_CUDA(ctx, out->m_base.alloc(MS ? MS : 1, st));
    SGPU_LAUNCH(ctx, (multi_copy_kernel<<<blocks_for(R + 1), TB, 0, st>>>(mflag.p, midx.p, moff_r.p, c_off.p, c_list.p, c_base.p, nst.p, R,
                                                       out->r_multi.p, out->m_off.p, out->m_locus.p, out->m_base.p)));
    SGPU_CUDA(ctx, cudaGetLastError());
    SGPU_CUDA(ctx, cudaStreamSynchronize(st)); // temporaries are released by the destructors below
    return SGPU_OK;
}
