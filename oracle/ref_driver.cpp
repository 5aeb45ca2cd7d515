// TEST INFRASTRUCTURE ONLY — never linked into, imported by or executed from the product path.
//
// Thin extern "C" driver around the UNMODIFIED reference implementation of the hot path.
// It is compiled together with the reference's own translation units *where they lie* under
// $SECEDO_REF (default /root/reference) by oracle/Makefile into oracle/_ref/libsecedo_ref.so.
// No reference source is copied into this repository; this file only converts flat CSR arrays
// into the reference's `std::vector<std::vector<PosData>>` and calls:
//   Filter::filter / Filter::is_significant      (util/is_significant.cpp:78-193)
//   computeSimilarityMatrix                      (similarity_matrix.cpp:295-433)
//   log_prob_same_genotype / log_prob_diff_genotype (similarity_matrix.cpp:117-170, file-local
//   helpers reached by including the reference .cpp as part of this TU)
//   read_pileup (text)                           (util/pileup_reader.cpp:12-137)
//
// CSR convention used everywhere in this repo (see include/secedo_b200.h):
//   chr_ptr[n_chr+1]  locus offsets per chromosome
//   row_ptr[n_loci+1] entry offsets per locus
//   position[n_loci], read_id[n_entries], gid_base[n_entries] (gid<<2|base, u16 as PosData)

#include "similarity_matrix.cpp" // the reference TU itself (path given with -I$SECEDO_REF)

#include "expectation_maximization.hpp"
#include "util/is_significant.hpp"
#include "util/pileup_reader.hpp"

#include <chrono>
#include <cstring>

namespace {

std::vector<std::vector<PosData>> from_csr(uint32_t n_chr,
                                           const uint64_t *chr_ptr,
                                           const uint64_t *row_ptr,
                                           const uint32_t *position,
                                           const uint32_t *read_id,
                                           const uint16_t *gid_base) {
    std::vector<std::vector<PosData>> pds(n_chr);
    for (uint32_t c = 0; c < n_chr; ++c) {
        pds[c].reserve(chr_ptr[c + 1] - chr_ptr[c]);
        for (uint64_t l = chr_ptr[c]; l < chr_ptr[c + 1]; ++l) {
            std::vector<uint32_t> ids(read_id + row_ptr[l], read_id + row_ptr[l + 1]);
            std::vector<uint16_t> gb(gid_base + row_ptr[l], gid_base + row_ptr[l + 1]);
            pds[c].emplace_back(position[l], std::move(ids), std::move(gb));
        }
    }
    return pds;
}

// holds the result of the last ref_filter() call until it is fetched
std::vector<std::vector<PosData>> g_filtered;

} // namespace

extern "C" {

/** Filter::is_significant on n count tuples (A,C,G,T order irrelevant — it sorts). */
void ref_is_significant(const uint16_t *counts4,
                        uint64_t n,
                        double theta,
                        int cell_proportion,
                        uint8_t *out) {
    Filter filter(theta, static_cast<uint8_t>(cell_proportion));
    for (uint64_t i = 0; i < n; ++i) {
        std::array<uint16_t, 4> bc
                = { counts4[4 * i], counts4[4 * i + 1], counts4[4 * i + 2], counts4[4 * i + 3] };
        out[i] = filter.is_significant(bc) ? 1 : 0;
    }
}

/**
 * Filter::filter. Result kept in a module-level buffer; sizes returned here, data fetched with
 * ref_filter_fetch(). Returns elapsed seconds of the reference call itself.
 */
double ref_filter(uint32_t n_chr,
                  const uint64_t *chr_ptr,
                  const uint64_t *row_ptr,
                  const uint32_t *position,
                  const uint32_t *read_id,
                  const uint16_t *gid_base,
                  const uint32_t *id_to_pos,
                  uint32_t n_groups,
                  double theta,
                  int cell_proportion,
                  uint32_t num_threads,
                  uint64_t *n_loci_out,
                  uint64_t *n_entries_out,
                  double *avg_coverage_out) {
    auto pds = from_csr(n_chr, chr_ptr, row_ptr, position, read_id, gid_base);
    std::vector<uint32_t> map(id_to_pos, id_to_pos + n_groups);
    Filter filter(theta, static_cast<uint8_t>(cell_proportion));
    auto t0 = std::chrono::steady_clock::now();
    auto [res, cov] = filter.filter(pds, map, "", num_threads);
    auto t1 = std::chrono::steady_clock::now();
    g_filtered = std::move(res);
    uint64_t nl = 0, ne = 0;
    for (const auto &chr : g_filtered) {
        nl += chr.size();
        for (const auto &pd : chr) {
            ne += pd.size();
        }
    }
    *n_loci_out = nl;
    *n_entries_out = ne;
    *avg_coverage_out = cov;
    return std::chrono::duration<double>(t1 - t0).count();
}

void ref_filter_fetch(uint64_t *chr_ptr,
                      uint64_t *row_ptr,
                      uint32_t *position,
                      uint32_t *read_id,
                      uint16_t *gid_base) {
    uint64_t l = 0, e = 0;
    chr_ptr[0] = 0;
    row_ptr[0] = 0;
    for (size_t c = 0; c < g_filtered.size(); ++c) {
        for (const auto &pd : g_filtered[c]) {
            position[l] = pd.position;
            std::memcpy(read_id + e, pd.read_ids.data(), pd.size() * sizeof(uint32_t));
            std::memcpy(gid_base + e, pd.group_ids_bases.data(), pd.size() * sizeof(uint16_t));
            e += pd.size();
            row_ptr[++l] = e;
        }
        chr_ptr[c + 1] = l;
    }
    g_filtered.clear();
}

/** computeSimilarityMatrix; `out` is num_cells*num_cells doubles. Returns elapsed seconds. */
double ref_similarity(uint32_t n_chr,
                      const uint64_t *chr_ptr,
                      const uint64_t *row_ptr,
                      const uint32_t *position,
                      const uint32_t *read_id,
                      const uint16_t *gid_base,
                      uint32_t num_cells,
                      uint32_t max_fragment_length,
                      const uint32_t *group_id_to_pos,
                      uint32_t n_groups,
                      double mutation_rate,
                      double homozygous_rate,
                      double seq_error_rate,
                      uint32_t num_threads,
                      const char *normalization,
                      double *out) {
    auto pds = from_csr(n_chr, chr_ptr, row_ptr, position, read_id, gid_base);
    std::vector<uint32_t> map(group_id_to_pos, group_id_to_pos + n_groups);
    auto t0 = std::chrono::steady_clock::now();
    Matd m = computeSimilarityMatrix(pds, num_cells, max_fragment_length, map, mutation_rate,
                                     homozygous_rate, seq_error_rate, num_threads, "",
                                     normalization);
    auto t1 = std::chrono::steady_clock::now();
    for (uint32_t i = 0; i < num_cells; ++i) {
        for (uint32_t j = 0; j < num_cells; ++j) {
            out[static_cast<uint64_t>(i) * num_cells + j] = m(i, j);
        }
    }
    return std::chrono::duration<double>(t1 - t0).count();
}

/** LS / LD tables of the reference for 0 <= x_s, x_d < n (n <= max_fragment_length). */
void ref_log_probs(double mutation_rate,
                   double homozygous_rate,
                   double seq_error_rate,
                   uint32_t max_fragment_length,
                   uint32_t n,
                   double *ls_out,
                   double *ld_out) {
    Cache cache(mutation_rate, homozygous_rate, seq_error_rate, max_fragment_length);
    Matd ls = Matd::fill(max_fragment_length, max_fragment_length,
                         std::numeric_limits<double>::max());
    Matd ld = Matd::fill(max_fragment_length, max_fragment_length,
                         std::numeric_limits<double>::max());
    for (uint32_t s = 0; s < n; ++s) {
        for (uint32_t d = 0; d < n; ++d) {
            ls_out[s * n + d] = log_prob_same_genotype(s, d, cache, ls);
            ld_out[s * n + d] = log_prob_diff_genotype(s, d, cache, ld);
        }
    }
}

/**
 * read_pileup on a text .pileup file with the identity grouping; two-call protocol (sizes, then
 * data) through the same module-level buffer as ref_filter.
 */
uint32_t ref_read_pileup_text(const char *fname,
                              uint32_t max_coverage,
                              uint64_t *n_loci_out,
                              uint64_t *n_entries_out) {
    std::vector<uint16_t> id_to_group = get_grouping(1, "", 10'000);
    auto [pds, n_cells, max_len] = read_pileup(fname, id_to_group, [](uint64_t) {}, max_coverage);
    std::ignore = n_cells;
    uint64_t ne = 0;
    for (const auto &pd : pds) {
        ne += pd.size();
    }
    *n_loci_out = pds.size();
    *n_entries_out = ne;
    g_filtered.clear();
    g_filtered.push_back(std::move(pds));
    return max_len;
}

/**
 * read_pileup (text or .bin by suffix) with an explicit grouping, max_coverage and an optional list of
 * positions (util/pileup_reader.cpp:259-271); same two-call protocol. Returns n_cells (max cell id + 1).
 */
uint32_t ref_read_pileup(const char *fname,
                         const uint16_t *id_to_group,
                         uint32_t n_ids,
                         uint32_t max_coverage,
                         const uint32_t *positions,
                         uint64_t n_positions,
                         uint64_t *n_loci_out,
                         uint64_t *n_entries_out,
                         uint32_t *max_len_out) {
    std::vector<uint16_t> grouping(id_to_group, id_to_group + n_ids);
    std::vector<uint32_t> pos(positions, positions + n_positions);
    auto [pds, n_cells, max_len] = read_pileup(fname, grouping, [](uint64_t) {}, max_coverage, pos, false);
    uint64_t ne = 0;
    for (const auto &pd : pds) {
        ne += pd.size();
    }
    *n_loci_out = pds.size();
    *n_entries_out = ne;
    *max_len_out = max_len;
    g_filtered.clear();
    g_filtered.push_back(std::move(pds));
    return n_cells;
}

/** expectation_maximization (expectation_maximization.cpp:131-160); prob_cluster_b updated in place. Seconds. */
double ref_expectation_maximization(uint32_t n_chr,
                                    const uint64_t *chr_ptr,
                                    const uint64_t *row_ptr,
                                    const uint32_t *position,
                                    const uint32_t *read_id,
                                    const uint16_t *gid_base,
                                    const uint32_t *id_to_pos,
                                    uint32_t n_groups,
                                    double theta,
                                    uint32_t n_cells,
                                    double *prob_cluster_b) {
    auto pds = from_csr(n_chr, chr_ptr, row_ptr, position, read_id, gid_base);
    std::vector<uint32_t> map(id_to_pos, id_to_pos + n_groups);
    std::vector<double> prob(prob_cluster_b, prob_cluster_b + n_cells);
    auto t0 = std::chrono::steady_clock::now();
    expectation_maximization(pds, map, 1, theta, &prob);
    auto t1 = std::chrono::steady_clock::now();
    std::memcpy(prob_cluster_b, prob.data(), n_cells * sizeof(double));
    return std::chrono::duration<double>(t1 - t0).count();
}

int ref_omp_max_threads() {
    return omp_get_max_threads();
}

} // extern "C"
