// Direct ingestion of SECEDO's binary pileup format into the device CSR.
//
// The reference reads a chromosome's `.bin` file record by record into std::vector<PosData>
// (util/pileup_reader.cpp:139-257; the writer is pileup.cpp:327-332):
//     u32 position | u16 coverage | u32 read_id[coverage] | u16 (cell_id << 2 | base)[coverage]
// and maps every cell id to its group (:206-215). Here the host only WALKS the records (6 bytes per
// locus: offsets, the max_coverage filter :195-197 and the optional position list :199-210), the file
// bytes go to the device as they are, and one kernel unpacks the kept records into read_id / gid_base
// with the group map applied. Records are only 2-byte aligned, so the 32-bit read ids are assembled from
// 16-bit halves.
#include "common.cuh"

#include <algorithm>
#include <cstring>
#include <vector>

namespace {

constexpr int TB = 256;

__global__ void __launch_bounds__(TB) bin_unpack_kernel(const uint16_t *__restrict__ bytes16 /* file bytes as u16 */,
                                                        const uint64_t *__restrict__ rec_off16 /* per locus: offset / 2 */,
                                                        const uint64_t *__restrict__ row_ptr, uint64_t l0, uint64_t n_loci,
                                                        const uint16_t *__restrict__ id_to_group, uint32_t n_ids,
                                                        uint32_t *__restrict__ read_id, uint16_t *__restrict__ gid_base,
                                                        unsigned int *__restrict__ max_ids /* [0] cell, [1] group */,
                                                        int *__restrict__ err) {
    const int lane = threadIdx.x & 31;
    const uint64_t warps_total = static_cast<uint64_t>(gridDim.x) * (TB / 32);
    unsigned int max_cell = 0, max_group = 0;
    for (uint64_t l = static_cast<uint64_t>(blockIdx.x) * (TB / 32) + (threadIdx.x >> 5); l < n_loci; l += warps_total) {
        const uint64_t e0 = row_ptr[l0 + l];
        const uint32_t cov = static_cast<uint32_t>(row_ptr[l0 + l + 1] - e0);
        const uint16_t *rec = bytes16 + rec_off16[l0 + l] + 3; // behind position and coverage
        const uint16_t *cb = rec + 2ull * cov;
        for (uint32_t i = lane; i < cov; i += 32) {
            read_id[e0 + i] = static_cast<uint32_t>(rec[2 * i]) | (static_cast<uint32_t>(rec[2 * i + 1]) << 16);
            const uint32_t v = cb[i];
            const uint32_t cell = v >> 2;
            uint32_t group = 0;
            if (cell >= n_ids) {
                atomicExch(err, SGPU_E_CELL_RANGE); // "Cell id is too large" + exit(1) in the reference
            } else {
                group = id_to_group[cell];
            }
            gid_base[e0 + i] = static_cast<uint16_t>((group << 2) | (v & 3u));
            max_cell = max(max_cell, cell);
            max_group = max(max_group, group);
        }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        max_cell = max(max_cell, __shfl_xor_sync(0xffffffffu, max_cell, o));
        max_group = max(max_group, __shfl_xor_sync(0xffffffffu, max_group, o));
    }
    if (lane == 0) {
        atomicMax(&max_ids[0], max_cell);
        atomicMax(&max_ids[1], max_group);
    }
}

} // namespace

extern "C" int sgpu_pileup_from_bin(sgpu_ctx *ctx, uint32_t n_chr, const void *const *file_bytes, const uint64_t *file_sizes,
                                    const uint16_t *id_to_group, uint32_t n_ids, uint32_t max_coverage,
                                    const uint32_t *const *positions, const uint64_t *n_positions, sgpu_pileup **out,
                                    uint32_t *n_cells, uint32_t *n_groups) {
    SGPU_CUDA(ctx, cudaSetDevice(ctx->device));
    cudaStream_t st = ctx->stream;
    // ---- host: walk the records (util/pileup_reader.cpp:165-211)
    std::vector<uint64_t> chr_ptr(n_chr + 1, 0), row_ptr(1, 0), rec_off16, file_base16(n_chr + 1, 0);
    std::vector<uint32_t> position;
    for (uint32_t c = 0; c < n_chr; ++c) {
        const uint8_t *b = static_cast<const uint8_t *>(file_bytes[c]);
        const uint64_t size = file_sizes[c];
        if (size & 1) {
            return sgpu_fail(ctx, SGPU_E_ARG, "chromosome %u: a binary pileup has an even number of bytes", c);
        }
        const uint32_t *pos_list = positions ? positions[c] : nullptr;
        const uint64_t n_pos = (positions && n_positions) ? n_positions[c] : 0;
        uint64_t pos_idx = 0;
        file_base16[c + 1] = file_base16[c] + size / 2;
        for (uint64_t off = 0; off + 6 <= size;) {
            uint32_t p;
            uint16_t cov;
            std::memcpy(&p, b + off, 4);
            std::memcpy(&cov, b + off + 4, 2);
            const uint64_t rec = off;
            off += 6 + 6ull * cov;
            if (off > size) {
                return sgpu_fail(ctx, SGPU_E_ARG, "chromosome %u: truncated record at byte %llu", c, (unsigned long long)rec);
            }
            if (cov > max_coverage) {
                continue;
            }
            if (pos_list && n_pos) {
                while (pos_idx < n_pos && pos_list[pos_idx] < p) {
                    ++pos_idx;
                }
                if (pos_idx == n_pos) {
                    break; // all valid positions were read
                }
                if (pos_list[pos_idx] > p) {
                    continue;
                }
            }
            position.push_back(p);
            rec_off16.push_back(file_base16[c] + rec / 2);
            row_ptr.push_back(row_ptr.back() + cov);
        }
        chr_ptr[c + 1] = position.size();
    }
    const uint64_t P = position.size(), E = row_ptr.back(), total16 = file_base16[n_chr];
    if (E >= 0xFFFFFFF0ull) {
        return sgpu_fail(ctx, SGPU_E_ARG, "more than 2^32 entries in one call: ingest the chromosomes in batches");
    }
    // ---- device: raw bytes + offsets up, unpack
    sgpu_pileup *p = new sgpu_pileup();
    PileupOwner p_owner(ctx, p);
    p->n_chr = n_chr;
    p->n_loci = P;
    p->n_entries = E;
    p->h_chr_ptr = new uint64_t[n_chr + 1];
    std::memcpy(p->h_chr_ptr, chr_ptr.data(), (n_chr + 1) * sizeof(uint64_t));
    SGPU_CUDA(ctx, sgpu_dev_alloc(ctx, reinterpret_cast<void **>(&p->d_chr_ptr), (n_chr + 1) * sizeof(uint64_t)));
    SGPU_CUDA(ctx, sgpu_dev_alloc(ctx, reinterpret_cast<void **>(&p->d_row_ptr), (P + 1) * sizeof(uint64_t)));
    SGPU_CUDA(ctx, sgpu_dev_alloc(ctx, reinterpret_cast<void **>(&p->d_position), (P ? P : 1) * sizeof(uint32_t)));
    SGPU_CUDA(ctx, sgpu_dev_alloc(ctx, reinterpret_cast<void **>(&p->d_read_id), (E ? E : 1) * sizeof(uint32_t)));
    SGPU_CUDA(ctx, sgpu_dev_alloc(ctx, reinterpret_cast<void **>(&p->d_gid_base), (E ? E : 1) * sizeof(uint16_t)));
    DevBuf<uint16_t> d_bytes, d_map;
    DevBuf<uint64_t> d_off;
    DevBuf<unsigned int> d_max;
    DevBuf<int> d_err;
    SGPU_CUDA(ctx, d_bytes.alloc(total16 ? total16 : 1, ctx));
    SGPU_CUDA(ctx, d_map.alloc(n_ids ? n_ids : 1, ctx));
    SGPU_CUDA(ctx, d_off.alloc(P ? P : 1, ctx));
    SGPU_CUDA(ctx, d_max.alloc(2, ctx));
    SGPU_CUDA(ctx, d_err.alloc(1, ctx));
    SGPU_CUDA(ctx, cudaMemsetAsync(d_max.p, 0, 2 * sizeof(unsigned int), st));
    SGPU_CUDA(ctx, cudaMemsetAsync(d_err.p, 0, sizeof(int), st));
    SGPU_CUDA(ctx, cudaMemcpyAsync(p->d_chr_ptr, p->h_chr_ptr, (n_chr + 1) * sizeof(uint64_t), cudaMemcpyHostToDevice, st));
    SGPU_CUDA(ctx, cudaMemcpyAsync(p->d_row_ptr, row_ptr.data(), (P + 1) * sizeof(uint64_t), cudaMemcpyHostToDevice, st));
    if (n_ids) {
        SGPU_CUDA(ctx, cudaMemcpyAsync(d_map.p, id_to_group, n_ids * sizeof(uint16_t), cudaMemcpyHostToDevice, st));
    }
    for (uint32_t c = 0; c < n_chr; ++c) {
        if (file_sizes[c]) {
            SGPU_CUDA(ctx, cudaMemcpyAsync(d_bytes.p + file_base16[c], file_bytes[c], file_sizes[c], cudaMemcpyHostToDevice, st));
        }
    }
    if (P) {
        SGPU_CUDA(ctx, cudaMemcpyAsync(p->d_position, position.data(), P * sizeof(uint32_t), cudaMemcpyHostToDevice, st));
        SGPU_CUDA(ctx, cudaMemcpyAsync(d_off.p, rec_off16.data(), P * sizeof(uint64_t), cudaMemcpyHostToDevice, st));
        const unsigned grid = static_cast<unsigned>(std::min<uint64_t>(ceil_div_u64(P, TB / 32), static_cast<uint64_t>(ctx->sm_count) * 32));
        SGPU_LAUNCH(ctx, (bin_unpack_kernel<<<grid, TB, 0, st>>>(d_bytes.p, d_off.p, p->d_row_ptr, 0, P, d_map.p, n_ids, p->d_read_id,
                                                                 p->d_gid_base, d_max.p, d_err.p)));
        SGPU_CUDA(ctx, cudaGetLastError());
    }
    SGPU_CUDA(ctx, cudaMemcpyAsync(&ctx->h_scratch[0], d_max.p, 2 * sizeof(unsigned int), cudaMemcpyDeviceToHost, st));
    SGPU_CUDA(ctx, cudaMemcpyAsync(&ctx->h_scratch[1], d_err.p, sizeof(int), cudaMemcpyDeviceToHost, st));
    SGPU_CUDA(ctx, cudaStreamSynchronize(st)); // the host vectors above are pageable and die here
    if (static_cast<int>(ctx->h_scratch[1] & 0xFFFFFFFFu) != 0) { // p_owner frees the half-built pileup
        return sgpu_fail(ctx, SGPU_E_CELL_RANGE, "Cell id is too large (>= %u). Increase --max_cell_count if using the default mapping, "
                         "or fix the mapping in --merge_file", n_ids);
    }
    // max_cell_id + 1, max_cell_id_grouped + 1 (util/pileup_reader.cpp:232-233)
    if (n_cells) {
        *n_cells = static_cast<uint32_t>(ctx->h_scratch[0] & 0xFFFFFFFFu) + 1;
    }
    if (n_groups) {
        *n_groups = static_cast<uint32_t>(ctx->h_scratch[0] >> 32) + 1;
    }
    *out = p_owner.release();
    return SGPU_OK;
}
