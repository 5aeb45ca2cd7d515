// Several GPUs behind ONE call of the C ABI (single process, one host thread per GPU): what a host program that is not
// an MPI / torchrun job — SECEDO's divide_cluster through the C++ shim — uses to spread computeSimilarityMatrix over all
// GPUs of the box (sgpu_init(devices, n_dev)).
//
//   1. the filtered loci are cut into one piece per GPU INSIDE the chromosomes, balanced by the work of the chosen path
//      (loci for the GEMM, squared coverage for the pair scatter), each piece with a halo of max_fragment_length bp;
//   2. the tail cutoff of every chromosome is decided from its end on GPU 0 (sgpu_chromosome_cutoff);
//   3. every GPU uploads its piece over its own PCIe link and accumulates it (sgpu_counts_accumulate_range);
//   4. the epilogue over peer memory: every GPU sums the planes of all GPUs over its share of the tiles through
//      NVLink (cudaDeviceEnablePeerAccess) and transforms them in the same kernel; the extrema are combined on the
//      host (two doubles per GPU); every GPU writes its share of the matrix straight into page-locked host memory.
// No NCCL: the only data that crosses between the GPUs are the count planes themselves, read in place.
#include "common.cuh"

#include <algorithm>
#include <condition_variable>
#include <cstring>
#include <mutex>
#include <thread>
#include <vector>

struct sgpu_multi {
    std::vector<sgpu_ctx *> ctx;
    std::vector<int> device;
    std::vector<sgpu_counts *> counts; // per GPU, for counts_cells cells
    uint32_t counts_cells = 0;
    double *staging = nullptr;         // page-locked, mapped, portable: counts_cells^2 doubles
    std::string error;
};

namespace {

struct Barrier { // std::barrier without C++20
    std::mutex m;
    std::condition_variable cv;
    int n, waiting = 0, generation = 0;
    explicit Barrier(int n_) : n(n_) {}
    void wait() {
        std::unique_lock<std::mutex> lk(m);
        const int g = generation;
        if (++waiting == n) {
            waiting = 0;
            ++generation;
            cv.notify_all();
        } else {
            cv.wait(lk, [&] { return g != generation; });
        }
    }
};

int mfail(sgpu_multi *m, int code, const std::string &msg) {
    if (m) {
        m->error = msg;
    }
    return code;
}

// one piece: per chromosome it touches, the loci [lo, hi) it has to hold and the positions it owns
struct PieceRange {
    uint32_t chrom;
    uint64_t lo, hi;
    uint32_t own_pos_begin, own_pos_end;
};

// the same cut as secedo_b200.dist.plan_pieces: runs of (nearly) equal total weight over the concatenated chromosomes
std::vector<std::vector<PieceRange>> plan_pieces(uint32_t n_chr, const uint64_t *chr_ptr, const uint32_t *position,
                                                 const std::vector<double> &weight, int n_pieces, uint32_t L) {
    std::vector<std::vector<PieceRange>> pieces(n_pieces);
    double total = 0;
    for (double w : weight) {
        total += w;
    }
    std::vector<double> cum(weight.size() + 1, 0.0); // weight in front of locus l
    for (size_t l = 0; l < weight.size(); ++l) {
        cum[l + 1] = cum[l] + weight[l];
    }
    for (uint32_t c = 0; c < n_chr; ++c) {
        const uint64_t l0 = chr_ptr[c], l1 = chr_ptr[c + 1];
        if (l1 == l0) {
            continue;
        }
        for (int k = 0; k < n_pieces; ++k) {
            const double w0 = total * k / n_pieces, w1 = total * (k + 1) / n_pieces;
            const uint64_t a = std::lower_bound(cum.begin() + l0, cum.begin() + l1, w0) - cum.begin();
            const uint64_t b = k == n_pieces - 1 ? l1 : std::lower_bound(cum.begin() + l0, cum.begin() + l1, w1) - cum.begin();
            if (b <= a) {
                continue;
            }
            PieceRange r;
            r.chrom = c;
            r.own_pos_begin = a > l0 ? position[a] : 0;
            r.own_pos_end = b < l1 ? position[b] : 0xFFFFFFFFu;
            // halo: positions > first owned - L and < last owned + L
            const int64_t first = position[a], last = position[b - 1];
            uint64_t lo = a, hi = b;
            while (lo > l0 && static_cast<int64_t>(position[lo - 1]) > first - static_cast<int64_t>(L)) {
                --lo;
            }
            while (hi < l1 && static_cast<int64_t>(position[hi]) < last + static_cast<int64_t>(L)) {
                ++hi;
            }
            r.lo = lo;
            r.hi = hi;
            pieces[k].push_back(r);
        }
    }
    return pieces;
}

// CSR of a set of locus ranges of a host pileup, rebased (one chromosome per range)
struct HostPiece {
    std::vector<uint64_t> chr_ptr, row_ptr;
    std::vector<uint32_t> position, read_id, own_lo, own_hi, tail;
    std::vector<uint16_t> gid_base;
};

HostPiece slice(const std::vector<PieceRange> &ranges, const uint64_t *row_ptr, const uint32_t *position, const uint32_t *read_id,
                const uint16_t *gid_base, const std::vector<uint32_t> &tail_position) {
    HostPiece h;
    h.chr_ptr.push_back(0);
    h.row_ptr.push_back(0);
    for (const PieceRange &r : ranges) {
        const uint64_t e0 = row_ptr[r.lo], e1 = row_ptr[r.hi];
        for (uint64_t l = r.lo; l < r.hi; ++l) {
            h.position.push_back(position[l]);
            h.row_ptr.push_back(h.row_ptr.back() + (row_ptr[l + 1] - row_ptr[l]));
        }
        h.read_id.insert(h.read_id.end(), read_id + e0, read_id + e1);
        h.gid_base.insert(h.gid_base.end(), gid_base + e0, gid_base + e1);
        h.chr_ptr.push_back(h.position.size());
        h.own_lo.push_back(r.own_pos_begin);
        h.own_hi.push_back(r.own_pos_end);
        h.tail.push_back(tail_position[r.chrom]);
    }
    return h;
}

} // namespace

extern "C" {

int sgpu_multi_init(const int *devices, int n_devices, sgpu_multi **out) {
    if (!out) {
        return SGPU_E_ARG;
    }
    sgpu_multi *m = new sgpu_multi();
    *out = m;
    int visible = 0;
    if (cudaGetDeviceCount(&visible) != cudaSuccess || visible == 0) {
        cudaGetLastError();
        return mfail(m, SGPU_E_CUDA, "no CUDA device visible (there is no CPU fallback)");
    }
    std::vector<int> devs;
    if (devices && n_devices > 0) {
        devs.assign(devices, devices + n_devices);
    } else {
        for (int d = 0; d < visible; ++d) {
            devs.push_back(d);
        }
    }
    if (devs.size() > SGPU_MAX_PEERS) {
        return mfail(m, SGPU_E_ARG, "at most " + std::to_string(SGPU_MAX_PEERS) + " GPUs");
    }
    for (int d : devs) {
        sgpu_ctx *c = nullptr;
        const int rc = sgpu_init(d, &c);
        if (rc != SGPU_OK) {
            const std::string msg = c ? sgpu_last_error(c) : "sgpu_init failed";
            sgpu_shutdown(c);
            return mfail(m, rc, msg);
        }
        m->ctx.push_back(c);
        m->device.push_back(d);
    }
    // every GPU reads every other GPU's count planes in place
    for (size_t i = 0; i < devs.size(); ++i) {
        cudaSetDevice(devs[i]);
        for (size_t j = 0; j < devs.size(); ++j) {
            if (i == j) {
                continue;
            }
            int can = 0;
            cudaDeviceCanAccessPeer(&can, devs[i], devs[j]);
            if (!can) {
                return mfail(m, SGPU_E_CUDA, "GPU " + std::to_string(devs[i]) + " cannot access GPU " + std::to_string(devs[j]) + " (no peer path)");
            }
            const cudaError_t e = cudaDeviceEnablePeerAccess(devs[j], 0);
            if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled) {
                return mfail(m, SGPU_E_CUDA, std::string("cudaDeviceEnablePeerAccess: ") + cudaGetErrorString(e));
            }
            cudaGetLastError();
        }
    }
    return SGPU_OK;
}

void sgpu_multi_shutdown(sgpu_multi *m) {
    if (!m) {
        return;
    }
    for (size_t i = 0; i < m->ctx.size(); ++i) {
        if (i < m->counts.size() && m->counts[i]) {
            sgpu_counts_free(m->ctx[i], m->counts[i]);
        }
    }
    if (m->staging) {
        cudaFreeHost(m->staging);
    }
    for (sgpu_ctx *c : m->ctx) {
        sgpu_shutdown(c);
    }
    delete m;
}

const char *sgpu_multi_last_error(const sgpu_multi *m) { return m ? m->error.c_str() : "null handle"; }
int sgpu_multi_size(const sgpu_multi *m) { return m ? static_cast<int>(m->ctx.size()) : 0; }
sgpu_ctx *sgpu_multi_ctx(sgpu_multi *m, int i) { return (m && i >= 0 && i < static_cast<int>(m->ctx.size())) ? m->ctx[i] : nullptr; }

int sgpu_multi_similarity(sgpu_multi *m, uint32_t n_chr, const uint64_t *chr_ptr, const uint64_t *row_ptr, const uint32_t *position,
                          const uint32_t *read_id, const uint16_t *gid_base, uint32_t num_cells, uint32_t max_fragment_length,
                          const uint32_t *group_id_to_pos, uint32_t n_groups, double mutation_rate, double homozygous_rate,
                          double seq_error_rate, uint32_t num_threads, int normalization, int path, double *out,
                          sgpu_stats *stats) {
    if (!m || m->ctx.empty()) {
        return SGPU_E_ARG;
    }
    if (normalization < 0 || normalization > 2) {
        return mfail(m, SGPU_E_ARG, "Invalid normalization: " + std::to_string(normalization));
    }
    const int G = static_cast<int>(m->ctx.size());
    const uint32_t L = max_fragment_length;
    const uint64_t P = chr_ptr[n_chr], E = P ? row_ptr[P] : 0;
    const uint64_t nn = static_cast<uint64_t>(num_cells) * num_cells;
    // ---- count planes and the host staging matrix, kept between calls of the same size
    if (m->counts_cells != num_cells || m->counts.size() != static_cast<size_t>(G)) {
        for (size_t i = 0; i < m->counts.size(); ++i) {
            sgpu_counts_free(m->ctx[i], m->counts[i]);
        }
        m->counts.assign(G, nullptr);
        if (m->staging) {
            cudaFreeHost(m->staging);
            m->staging = nullptr;
        }
        for (int g = 0; g < G; ++g) {
            const int rc = sgpu_counts_create(m->ctx[g], num_cells, &m->counts[g]);
            if (rc != SGPU_OK) {
                return mfail(m, rc, sgpu_last_error(m->ctx[g]));
            }
        }
        cudaSetDevice(m->device[0]);
        if (cudaHostAlloc(reinterpret_cast<void **>(&m->staging), std::max<uint64_t>(nn, 1) * sizeof(double),
                          cudaHostAllocPortable | cudaHostAllocMapped) != cudaSuccess) {
            cudaGetLastError();
            return mfail(m, SGPU_E_CUDA, "cudaHostAlloc of the result staging matrix failed");
        }
        m->counts_cells = num_cells;
    }
    // ---- path and piece weights: the whole pileup decides the path, so that every GPU takes the same one
    if (path == SGPU_PATH_AUTO) {
        sgpu_pileup probe;
        probe.n_loci = P;
        probe.n_entries = E;
        path = P == 0 ? SGPU_PATH_SCATTER : sgpu_choose_path(&probe, num_cells);
    }
    std::vector<double> weight(P);
    for (uint64_t l = 0; l < P; ++l) {
        const double c = static_cast<double>(row_ptr[l + 1] - row_ptr[l]);
        weight[l] = path == SGPU_PATH_SCATTER ? 1.0 + c * c : 1.0 + 1e-4 * c; // pairs per locus / one K column per locus
    }
    const std::vector<std::vector<PieceRange>> pieces = plan_pieces(n_chr, chr_ptr, position, weight, G, L);
    // ---- the cutoff of every chromosome from its end (GPU 0); longer suffixes until it is decided
    std::vector<uint32_t> tail_position(n_chr, 0xFFFFFFFFu);
    {
        std::vector<uint64_t> span(n_chr, 6ull * L);
        std::vector<uint8_t> done(n_chr, 0);
        for (uint32_t c = 0; c < n_chr; ++c) {
            done[c] = chr_ptr[c + 1] == chr_ptr[c];
        }
        for (;;) {
            std::vector<PieceRange> ends;
            std::vector<uint32_t> which;
            std::vector<uint8_t> whole;
            for (uint32_t c = 0; c < n_chr; ++c) {
                if (done[c]) {
                    continue;
                }
                const uint64_t l0 = chr_ptr[c], l1 = chr_ptr[c + 1];
                const int64_t from = static_cast<int64_t>(position[l1 - 1]) - static_cast<int64_t>(span[c]);
                uint64_t lo = std::lower_bound(position + l0, position + l1, static_cast<uint32_t>(std::max<int64_t>(from, 0))) - position;
                if (from <= 0) {
                    lo = l0;
                }
                ends.push_back(PieceRange{ c, lo, l1, 0, 0xFFFFFFFFu });
                which.push_back(c);
                whole.push_back(lo == l0 ? 1 : 0);
            }
            if (ends.empty()) {
                break;
            }
            const HostPiece h = slice(ends, row_ptr, position, read_id, gid_base, tail_position);
            sgpu_pileup *dev = nullptr;
            int rc = sgpu_pileup_upload(m->ctx[0], static_cast<uint32_t>(ends.size()), h.chr_ptr.data(), h.row_ptr.data(),
                                        h.position.data(), h.read_id.data(), h.gid_base.data(), &dev);
            std::vector<uint32_t> tp(ends.size());
            std::vector<uint8_t> ok(ends.size());
            if (rc == SGPU_OK) {
                rc = sgpu_chromosome_cutoff(m->ctx[0], dev, L, num_threads, whole.data(), tp.data(), ok.data());
            }
            sgpu_pileup_free(m->ctx[0], dev);
            if (rc != SGPU_OK) {
                return mfail(m, rc, sgpu_last_error(m->ctx[0]));
            }
            for (size_t i = 0; i < ends.size(); ++i) {
                if (ok[i]) {
                    tail_position[which[i]] = tp[i];
                    done[which[i]] = 1;
                } else {
                    span[which[i]] *= 4;
                }
            }
        }
    }
    // ---- one thread per GPU
    std::vector<int> rcs(G, SGPU_OK);
    std::vector<sgpu_stats> sts(G);
    std::vector<double> ext(2 * G, 0.0);
    std::vector<int> planes(G, 2), spills(G, 0);
    Barrier bar(G);
    std::vector<const int32_t *> peer_planes(G);
    std::vector<const double *> peer_spill(G, nullptr);
    auto worker = [&](int g) {
        sgpu_ctx *ctx = m->ctx[g];
        sgpu_counts *c = m->counts[g];
        int rc = sgpu_counts_zero(ctx, c);
        std::memset(&sts[g], 0, sizeof(sgpu_stats));
        if (rc == SGPU_OK && !pieces[g].empty()) {
            const HostPiece h = slice(pieces[g], row_ptr, position, read_id, gid_base, tail_position);
            sgpu_pileup *dev = nullptr;
            rc = sgpu_pileup_upload(ctx, static_cast<uint32_t>(pieces[g].size()), h.chr_ptr.data(), h.row_ptr.data(), h.position.data(),
                                    h.read_id.data(), h.gid_base.data(), &dev);
            if (rc == SGPU_OK) {
                rc = sgpu_counts_accumulate_range(ctx, c, dev, L, group_id_to_pos, n_groups, mutation_rate, homozygous_rate, seq_error_rate,
                                                  num_threads, h.own_lo.data(), h.own_hi.data(), h.tail.data(), path, &sts[g]);
            }
            sgpu_pileup_free(ctx, dev);
        }
        if (rc == SGPU_OK) {
            rc = sgpu_synchronize(ctx);
        }
        rcs[g] = rc;
        int32_t *i32 = nullptr;
        uint64_t n_i32 = 0, n_f64 = 0;
        double *f64 = nullptr;
        sgpu_counts_buffers(c, &i32, &n_i32, &f64, &n_f64, nullptr, nullptr);
        planes[g] = nn ? static_cast<int>(n_i32 / nn) : 2;
        spills[g] = n_f64 > 0;
        bar.wait(); // everybody has accumulated (or failed)
        bool failed = false;
        int max_planes = 2, any_spill = 0;
        for (int q = 0; q < G; ++q) {
            failed = failed || rcs[q] != SGPU_OK;
            max_planes = std::max(max_planes, planes[q]);
            any_spill = any_spill || spills[q];
        }
        if (failed) {
            return;
        }
        rc = sgpu_counts_set_layout(ctx, c, max_planes, any_spill);
        if (rc == SGPU_OK) {
            rc = sgpu_synchronize(ctx); // a freshly created spill plane is zero before anybody reads it
        }
        sgpu_counts_buffers(c, &i32, &n_i32, &f64, &n_f64, nullptr, nullptr);
        peer_planes[g] = i32;
        peer_spill[g] = f64;
        rcs[g] = rc;
        bar.wait(); // layouts agree, pointers published
        for (int q = 0; q < G; ++q) {
            failed = failed || rcs[q] != SGPU_OK;
        }
        if (failed) {
            return;
        }
        double *d_ext = nullptr;
        rc = sgpu_slab_raw(ctx, c, peer_planes.data(), any_spill ? peer_spill.data() : nullptr, G, g, G, L, mutation_rate, homozygous_rate,
                           seq_error_rate, &d_ext);
        if (rc == SGPU_OK) {
            rc = sgpu_synchronize(ctx);
        }
        if (rc == SGPU_OK && cudaMemcpy(&ext[2 * g], d_ext, 2 * sizeof(double), cudaMemcpyDeviceToHost) != cudaSuccess) {
            rc = SGPU_E_CUDA;
        }
        rcs[g] = rc;
        bar.wait(); // nobody reads a peer's planes any more; all extrema are on the host
        for (int q = 0; q < G; ++q) {
            failed = failed || rcs[q] != SGPU_OK;
        }
        if (failed) {
            return;
        }
        double mx[2] = { ext[0], ext[1] };
        for (int q = 1; q < G; ++q) {
            mx[0] = std::max(mx[0], ext[2 * q]);
            mx[1] = std::max(mx[1], ext[2 * q + 1]);
        }
        if (cudaMemcpy(d_ext, mx, 2 * sizeof(double), cudaMemcpyHostToDevice) != cudaSuccess) {
            rc = SGPU_E_CUDA;
        }
        if (rc == SGPU_OK) {
            void *alias = nullptr;
            if (cudaHostGetDevicePointer(&alias, m->staging, 0) != cudaSuccess) {
                rc = SGPU_E_CUDA;
            } else {
                rc = sgpu_slab_finalize(ctx, c, normalization, static_cast<double *>(alias), nullptr);
            }
        }
        if (rc == SGPU_OK) {
            rc = sgpu_synchronize(ctx);
        }
        rcs[g] = rc;
        bar.wait(); // the staging matrix is complete
        for (int q = 0; q < G; ++q) {
            failed = failed || rcs[q] != SGPU_OK;
        }
        if (failed || !out) {
            return;
        }
        // rows [g / G, (g + 1) / G) of the staging matrix -> the caller's (pageable) matrix
        const uint64_t r0 = static_cast<uint64_t>(num_cells) * g / G, r1 = static_cast<uint64_t>(num_cells) * (g + 1) / G;
        std::memcpy(out + r0 * num_cells, m->staging + r0 * num_cells, (r1 - r0) * num_cells * sizeof(double));
    };
    std::vector<std::thread> threads;
    for (int g = 1; g < G; ++g) {
        threads.emplace_back(worker, g);
    }
    worker(0);
    for (std::thread &t : threads) {
        t.join();
    }
    for (int g = 0; g < G; ++g) {
        if (rcs[g] != SGPU_OK) {
            return mfail(m, rcs[g], std::string("GPU ") + std::to_string(m->device[g]) + ": " + sgpu_last_error(m->ctx[g]));
        }
    }
    if (stats) {
        std::memset(stats, 0, sizeof(*stats));
        stats->path_used = path;
        for (int g = 0; g < G; ++g) { // additive numbers; the times of the slowest GPU
            stats->n_loci += sts[g].n_loci;
            stats->n_entries += sts[g].n_entries;
            stats->n_pairs_first += sts[g].n_pairs_first;
            stats->n_pairs_multi += sts[g].n_pairs_multi;
            stats->n_span_splits += sts[g].n_span_splits;
            stats->gemm_launches += sts[g].gemm_launches;
            stats->ms_link = std::max(stats->ms_link, sts[g].ms_link);
            stats->ms_first_order = std::max(stats->ms_first_order, sts[g].ms_first_order);
            stats->ms_multi = std::max(stats->ms_multi, sts[g].ms_multi);
            stats->ms_stage = std::max(stats->ms_stage, sts[g].ms_stage);
            stats->ms_gemm = std::max(stats->ms_gemm, sts[g].ms_gemm);
        }
    }
    return SGPU_OK;
}

} // extern "C"
