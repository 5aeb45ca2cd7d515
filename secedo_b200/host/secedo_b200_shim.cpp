// Drop-in host shim: defines the reference's own entry points for the hot path on top of the C ABI
// (include/secedo_b200.h), so that spectral_clustering.cpp / secedo_main.cpp / the reference's tests
// link unchanged.
//
//   Filter::Filter, Filter::is_significant (x2), Filter::filter, Filter::log_fact
//                                   replaces util/is_significant.cpp:48-193
//   computeSimilarityMatrix         replaces similarity_matrix.cpp:295-433
//   expectation_maximization        replaces expectation_maximization.cpp:131-160
//
// Build inside SECEDO:  add this file INSTEAD OF util/is_significant.cpp and similarity_matrix.cpp to
// the `util` / `similarity_matrix` targets (and INSTEAD OF expectation_maximization.cpp), add <repo>/include to the include path and link
// libsecedo_b200.so. Build stand-alone (tests): -Itests/compat.
//
// Same argument meaning and error behaviour as the reference: an unknown normalization throws
// std::logic_error("Invalid normalization: ...") (similarity_matrix.cpp:264); every other failure
// prints the message and exits with status 1 (the reference logs through spdlog and calls
// std::exit(1)). There is no CPU fallback: without a B200 the first call fails.
//
// Device residency across divide_cluster's recursion (spectral_clustering.cpp:336-433): every node of the recursion
// calls Filter::filter on the SAME whole pileup `pds` with another id_to_pos, then computeSimilarityMatrix (and
// optionally expectation_maximization) on what Filter::filter returned. The shim therefore keeps two device pileups:
// the last raw pileup it uploaded (recognised by the address of the caller's vector plus a fingerprint over every
// locus) and the last filter result (recognised by the fingerprint of the vector<vector<PosData>> it returned). The
// whole pileup crosses PCIe once per run instead of once per call, and the filtered one not at all.
// SECEDO_B200_NO_CACHE=1 switches both off. SECEDO_B200_DEVICES=0,1,... (default: all visible GPUs) selects the GPUs
// computeSimilarityMatrix spreads the loci over; SECEDO_B200_DEVICE=n the one everything else runs on.
#include "expectation_maximization.hpp"
#include "similarity_matrix.hpp"
#include "util/is_significant.hpp"
#include "util/logger.hpp"

#include "secedo_b200.h"

#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <stdexcept>

namespace {

sgpu_ctx *context() {
    static sgpu_ctx *ctx = nullptr;
    if (!ctx) {
        const char *dev = std::getenv("SECEDO_B200_DEVICE");
        const int rc = sgpu_init(dev ? std::atoi(dev) : 0, &ctx);
        if (rc != SGPU_OK) {
            std::fprintf(stderr, "secedo_b200: %s\n", sgpu_last_error(ctx));
            std::exit(1);
        }
    }
    return ctx;
}

void check(int rc) {
    if (rc != SGPU_OK) {
        std::fprintf(stderr, "secedo_b200: error %d: %s\n", rc, sgpu_last_error(context()));
        std::exit(1);
    }
}

// all GPUs computeSimilarityMatrix spreads the loci over (SECEDO_B200_DEVICES=0,1,...; default: every visible GPU), or
// null when there is only one
sgpu_multi *multi() {
    static bool tried = false;
    static sgpu_multi *m = nullptr;
    if (!tried) {
        tried = true;
        std::vector<int> devs;
        if (const char *e = std::getenv("SECEDO_B200_DEVICES")) {
            for (const char *p = e; *p;) {
                devs.push_back(std::atoi(p));
                while (*p && *p != ',') {
                    ++p;
                }
                if (*p == ',') {
                    ++p;
                }
            }
        }
        if (devs.size() == 1) {
            return nullptr;
        }
        const int rc = sgpu_multi_init(devs.empty() ? nullptr : devs.data(), static_cast<int>(devs.size()), &m);
        if (rc != SGPU_OK) {
            std::fprintf(stderr, "secedo_b200: %s\n", sgpu_multi_last_error(m));
            std::exit(1);
        }
        if (sgpu_multi_size(m) < 2) {
            sgpu_multi_shutdown(m);
            m = nullptr;
        }
    }
    return m;
}

bool cache_enabled() {
    static const bool on = [] {
        const char *e = std::getenv("SECEDO_B200_NO_CACHE");
        return !(e && e[0] == '1');
    }();
    return on;
}

// what identifies the CONTENT of a vector<vector<PosData>> cheaply: sizes, every position and entry count, and the
// first / last entry of every locus (O(loci), not O(entries))
struct Fingerprint {
    uint64_t n_chr = 0, n_loci = 0, n_entries = 0, hash = 0;
    bool operator==(const Fingerprint &o) const {
        return n_chr == o.n_chr && n_loci == o.n_loci && n_entries == o.n_entries && hash == o.hash;
    }
};

Fingerprint fingerprint(const std::vector<std::vector<PosData>> &pds) {
    Fingerprint f;
    f.n_chr = pds.size();
    uint64_t h = 0x9E3779B97F4A7C15ull;
    auto mix = [&h](uint64_t v) {
        h ^= v + 0x9E3779B97F4A7C15ull + (h << 6) + (h >> 2);
        h *= 0xBF58476D1CE4E5B9ull;
    };
    for (const auto &chr : pds) {
        f.n_loci += chr.size();
        mix(chr.size());
        for (const auto &pd : chr) {
            const size_t n = pd.read_ids.size();
            f.n_entries += n;
            mix((static_cast<uint64_t>(pd.position) << 32) | n);
            if (n) {
                mix((static_cast<uint64_t>(pd.read_ids.front()) << 32) | pd.read_ids.back());
                mix((static_cast<uint64_t>(pd.group_ids_bases.front()) << 16) | pd.group_ids_bases.back());
            }
        }
    }
    f.hash = h;
    return f;
}

// vector<vector<PosData>> -> the five flat CSR arrays of the ABI
struct Csr {
    std::vector<uint64_t> chr_ptr, row_ptr;
    std::vector<uint32_t> position, read_id;
    std::vector<uint16_t> gid_base;

    explicit Csr(const std::vector<std::vector<PosData>> &pds) {
        uint64_t n_loci = 0, n_entries = 0;
        for (const auto &chr : pds) {
            n_loci += chr.size();
            for (const auto &pd : chr) {
                n_entries += pd.read_ids.size();
            }
        }
        chr_ptr.reserve(pds.size() + 1);
        row_ptr.reserve(n_loci + 1);
        position.reserve(n_loci);
        read_id.reserve(n_entries);
        gid_base.reserve(n_entries);
        chr_ptr.push_back(0);
        row_ptr.push_back(0);
        for (const auto &chr : pds) {
            for (const auto &pd : chr) {
                position.push_back(pd.position);
                read_id.insert(read_id.end(), pd.read_ids.begin(), pd.read_ids.end());
                gid_base.insert(gid_base.end(), pd.group_ids_bases.begin(), pd.group_ids_bases.end());
                row_ptr.push_back(read_id.size());
            }
            chr_ptr.push_back(position.size());
        }
    }

    sgpu_pileup *upload() const {
        sgpu_pileup *p = nullptr;
        check(sgpu_pileup_upload(context(), static_cast<uint32_t>(chr_ptr.size() - 1), chr_ptr.data(), row_ptr.data(),
                                 position.data(), read_id.data(), gid_base.data(), &p));
        return p;
    }
};

// a device pileup that outlives the call that created it
struct Resident {
    const void *address = nullptr; // of the caller's outer vector (raw pileup only)
    Fingerprint fp;
    sgpu_pileup *dev = nullptr;
    void reset() {
        if (dev) {
            sgpu_pileup_free(context(), dev);
        }
        dev = nullptr;
        address = nullptr;
    }
};
Resident g_raw, g_filtered;
uint64_t g_uploads = 0; // pileups flattened + uploaded so far (tests)

// device pileup of pos_data; *owned = the caller frees it after use
sgpu_pileup *stage(const std::vector<std::vector<PosData>> &pos_data, bool *owned) {
    *owned = false;
    if (cache_enabled()) {
        const Fingerprint fp = fingerprint(pos_data);
        if (g_filtered.dev && g_filtered.fp == fp) {
            return g_filtered.dev; // what Filter::filter just returned: still on the device
        }
        if (g_raw.dev && g_raw.address == &pos_data && g_raw.fp == fp) {
            return g_raw.dev;      // the whole pileup of an earlier node of the recursion
        }
        // a pileup not seen before takes the raw slot (divide_cluster passes the same one at every node)
        g_raw.reset();
        g_raw.dev = Csr(pos_data).upload();
        g_raw.address = &pos_data;
        g_raw.fp = fp;
        ++g_uploads;
        return g_raw.dev;
    }
    *owned = true;
    ++g_uploads;
    return Csr(pos_data).upload();
}

int normalization_code(const std::string &normalization) {
    if (normalization == "ADD_MIN") {
        return SGPU_NORM_ADD_MIN;
    }
    if (normalization == "EXPONENTIATE") {
        return SGPU_NORM_EXPONENTIATE;
    }
    if (normalization == "SCALE_MAX_1") {
        return SGPU_NORM_SCALE_MAX_1;
    }
    throw std::logic_error("Invalid normalization: " + normalization);
}

} // namespace

// test hook: number of pileups the shim has flattened and uploaded
extern "C" uint64_t secedo_b200_shim_uploads() { return g_uploads; }

// ---- Filter -------------------------------------------------------------------------------------------

Filter::Filter(double theta, uint8_t cell_proportion)
    : theta(theta),
      cell_proportion(cell_proportion),
      log_theta_3(std::log(theta / 3)),
      log_one_minus_theta(std::log(1 - theta)) {
    // table used only by log_fact(), which is not on the hot path
    log_factorial.reserve(171);
    double f = 1;
    log_factorial.push_back(std::log(f));
    for (uint32_t i = 1; i < 171; ++i) {
        f *= i;
        log_factorial.push_back(std::log(f));
    }
}

double Filter::log_fact(uint32_t n) {
    return n > 170 ? 0.5 * std::log(2 * M_PI * n) + n * std::log(n / M_E) : log_factorial.at(n);
}

bool Filter::is_significant(std::array<uint16_t, 4> &base_count) {
    // the reference returns before sorting its in/out argument when the coverage is below 2 (util/is_significant.cpp:83-88)
    if (base_count[0] + base_count[1] + base_count[2] + base_count[3] < 2) {
        return false;
    }
    uint8_t out = 0;
    check(sgpu_is_significant(context(), base_count.data(), 1, theta, cell_proportion, &out));
    std::sort(base_count.begin(), base_count.end()); // the reference sorts its argument in place (:88)
    return out != 0;
}

bool Filter::is_significant(const PosData &pos_data, uint16_t *coverage) {
    std::array<uint16_t, 4> base_count = { 0, 0, 0, 0 };
    for (uint32_t i = 0; i < pos_data.size(); ++i) {
        base_count[pos_data.base(i)]++;
    }
    *coverage = base_count[0] + base_count[1] + base_count[2] + base_count[3];
    return is_significant(base_count);
}

std::pair<std::vector<std::vector<PosData>>, double> Filter::filter(const std::vector<std::vector<PosData>> &pos_data,
                                                                    const std::vector<uint32_t> &id_to_pos,
                                                                    const std::string &marker, uint32_t num_threads) {
    (void)num_threads; // the reference parallelises over chromosomes with it; the result does not depend on it
    sgpu_ctx *ctx = context();
    bool owned = false;
    sgpu_pileup *in = stage(pos_data, &owned), *kept = nullptr;
    double avg_coverage = 0;
    check(sgpu_filter(ctx, in, id_to_pos.data(), static_cast<uint32_t>(id_to_pos.size()), theta, cell_proportion, &kept,
                      &avg_coverage));
    if (owned) {
        sgpu_pileup_free(ctx, in);
    }
    uint32_t n_chr = 0;
    uint64_t n_loci = 0, n_entries = 0;
    check(sgpu_pileup_dims(kept, &n_chr, &n_loci, &n_entries));
    std::vector<uint64_t> chr_ptr(n_chr + 1), row_ptr(n_loci + 1);
    std::vector<uint32_t> position(n_loci), read_id(n_entries);
    std::vector<uint16_t> gid_base(n_entries);
    check(sgpu_pileup_download(ctx, kept, chr_ptr.data(), row_ptr.data(), position.data(), read_id.data(), gid_base.data()));
    std::vector<std::vector<PosData>> result(n_chr);
    for (uint32_t c = 0; c < n_chr; ++c) {
        result[c].reserve(chr_ptr[c + 1] - chr_ptr[c]);
        for (uint64_t l = chr_ptr[c]; l < chr_ptr[c + 1]; ++l) {
            result[c].emplace_back(position[l], std::vector<uint32_t>(read_id.begin() + row_ptr[l], read_id.begin() + row_ptr[l + 1]),
                                   std::vector<uint16_t>(gid_base.begin() + row_ptr[l], gid_base.begin() + row_ptr[l + 1]));
        }
    }
    // the filtered pileup stays on the device for the computeSimilarityMatrix / expectation_maximization calls that
    // divide_cluster makes with this very result (spectral_clustering.cpp:354-356, 375-383)
    g_filtered.reset();
    if (cache_enabled()) {
        g_filtered.dev = kept;
        g_filtered.fp = fingerprint(result);
    } else {
        sgpu_pileup_free(ctx, kept);
    }
    // util/is_significant.cpp:190-191
    logger()->trace("Avg coverage for cluster {}: {}. Total positions: {}", marker, avg_coverage, n_loci);
    return { std::move(result), avg_coverage };
}

// ---- similarity matrix ------------------------------------------------------------------------------------

Matd computeSimilarityMatrix(const std::vector<std::vector<PosData>> &pos_data, uint32_t num_cells,
                             uint32_t max_fragment_length, const std::vector<uint32_t> &group_id_to_pos,
                             double mutation_rate, double homozygous_rate, double seq_error_rate,
                             const uint32_t num_threads, const std::string &marker, const std::string &normalization) {
    (void)marker;
    const int norm = normalization_code(normalization); // throws like the reference, before any work
    Matd result(num_cells, num_cells);
    logger()->trace("Normalizing similarity matrix..."); // similarity_matrix.cpp:272 (the epilogue is part of the call)
    if (sgpu_multi *m = multi()) {
        // several GPUs: pieces of the chromosomes, one per GPU, uploaded over the GPUs' own PCIe links
        const Csr csr(pos_data);
        ++g_uploads;
        const int rc = sgpu_multi_similarity(m, static_cast<uint32_t>(csr.chr_ptr.size() - 1), csr.chr_ptr.data(), csr.row_ptr.data(),
                                             csr.position.data(), csr.read_id.data(), csr.gid_base.data(), num_cells, max_fragment_length,
                                             group_id_to_pos.data(), static_cast<uint32_t>(group_id_to_pos.size()), mutation_rate,
                                             homozygous_rate, seq_error_rate, num_threads, norm, SGPU_PATH_AUTO, result.data(), nullptr);
        if (rc != SGPU_OK) {
            std::fprintf(stderr, "secedo_b200: error %d: %s\n", rc, sgpu_multi_last_error(m));
            std::exit(1);
        }
        return result;
    }
    sgpu_ctx *ctx = context();
    bool owned = false;
    sgpu_pileup *p = stage(pos_data, &owned);
    check(sgpu_similarity(ctx, p, num_cells, max_fragment_length, group_id_to_pos.data(),
                          static_cast<uint32_t>(group_id_to_pos.size()), mutation_rate, homozygous_rate, seq_error_rate,
                          num_threads /* selects the reference's tail cutoff */, norm, SGPU_PATH_AUTO, result.data(), nullptr));
    if (owned) {
        sgpu_pileup_free(ctx, p);
    }
    return result;
}

// ---- expectation maximisation -------------------------------------------------------------------------------

void expectation_maximization(const std::vector<std::vector<PosData>> &pos_data, const std::vector<uint32_t> &cell_id_to_cell_pos,
                              uint32_t num_threads, double theta, std::vector<double> *prob_cluster_b) {
    (void)num_threads; // unused by the reference as well (expectation_maximization.cpp:133)
    sgpu_ctx *ctx = context();
    bool owned = false;
    sgpu_pileup *p = stage(pos_data, &owned);
    check(sgpu_expectation_maximization(ctx, p, cell_id_to_cell_pos.data(), static_cast<uint32_t>(cell_id_to_cell_pos.size()), theta,
                                        prob_cluster_b->data(), static_cast<uint32_t>(prob_cluster_b->size()), 0, nullptr, nullptr));
    if (owned) {
        sgpu_pileup_free(ctx, p);
    }
}
