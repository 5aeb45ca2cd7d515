// End-to-end timing THROUGH THE C++ DROP-IN SHIM, the way an unmodified SECEDO calls it (spectral_clustering.cpp:336-356):
// a pageable std::vector<std::vector<PosData>> in, Filter::filter -> computeSimilarityMatrix -> Matd out. Everything the
// shim does on the host (flattening the vectors into CSR, pageable uploads, rebuilding the filtered vectors, copying the
// matrix into the Matd) is inside the timed calls. The pileup is generated here (cfg3 shape: cells x coverage, half of
// the loci carry a somatic variant in a quarter of the cells), not read from a file.
//   shim_bench <cells> <coverage> <chromosomes> <loci per chromosome> <threads> [repeats]
// prints one JSON line.
#include "similarity_matrix.hpp"
#include "util/is_significant.hpp"

#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <string>

extern "C" uint64_t secedo_b200_shim_uploads();

static inline uint64_t next(uint64_t &s) { // xorshift64*
    s ^= s >> 12;
    s ^= s << 25;
    s ^= s >> 27;
    return s * 0x2545F4914F6CDD1Dull;
}
static double now() {
    return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count();
}

int main(int argc, char **argv) {
    if (argc < 6) {
        return 2;
    }
    const uint32_t n_cells = std::atoi(argv[1]);
    const double coverage = std::atof(argv[2]);
    const uint32_t n_chr = std::atoi(argv[3]), loci_per_chr = std::atoi(argv[4]), threads = std::atoi(argv[5]);
    const int repeats = argc > 6 ? std::atoi(argv[6]) : 2;
    const double theta = 0.001, eps = 0.01, h = 0.15;
    const uint32_t L = 1000;
    // every (cell, locus) holds a read with probability 1 - exp(-coverage) (and a second one with the rest: enough for timing)
    const uint64_t p_read = static_cast<uint64_t>((1.0 - std::exp(-coverage)) * 4294967296.0);
    const uint64_t p_err = static_cast<uint64_t>(theta * 4294967296.0);
    uint64_t seed = 0x9E3779B97F4A7C15ull;
    uint32_t rid = 1;
    uint64_t entries = 0;
    std::vector<std::vector<PosData>> pds(n_chr);
    const double t_gen0 = now();
    for (uint32_t c = 0; c < n_chr; ++c) {
        pds[c].reserve(loci_per_chr);
        for (uint32_t l = 0; l < loci_per_chr; ++l) {
            const bool somatic = (next(seed) & 1u) != 0;
            const uint32_t ref = next(seed) & 3u, alt = (ref + 1 + next(seed) % 3) & 3u;
            std::vector<uint32_t> ids;
            std::vector<uint16_t> gbs;
            ids.reserve(static_cast<size_t>(n_cells * (1.0 - std::exp(-coverage)) * 1.05) + 16);
            gbs.reserve(ids.capacity());
            for (uint32_t cell = 0; cell < n_cells; ++cell) {
                const uint64_t r = next(seed);
                if ((r & 0xFFFFFFFFu) >= p_read) {
                    continue;
                }
                uint32_t base = (somatic && (cell & 3u) == 1u) ? alt : ref; // a clone of a quarter of the cells (a 50:50 split reads as a germline heterozygous locus and is filtered out)
                if (((r >> 32) & 0xFFFFFFFFu) < p_err) {
                    base = (base + 1 + (r >> 60) % 3) & 3u;
                }
                ids.push_back(rid++);
                gbs.push_back(static_cast<uint16_t>((cell << 2) | base));
            }
            entries += ids.size();
            pds[c].emplace_back(1000 + 400 * l, std::move(ids), std::move(gbs));
        }
    }
    const double t_gen = now() - t_gen0;
    std::vector<uint32_t> id_to_pos(n_cells);
    for (uint32_t i = 0; i < n_cells; ++i) {
        id_to_pos[i] = i;
    }
    Filter filter(theta);
    double best_filter = 1e30, best_sim = 1e30, first_filter = 0, first_sim = 0;
    uint64_t sig = 0, kept_entries = 0;
    double checksum = 0;
    for (int r = 0; r < repeats; ++r) {
        const double t0 = now();
        auto [filtered, cov] = filter.filter(pds, id_to_pos, "", threads);
        const double t1 = now();
        Matd m = computeSimilarityMatrix(filtered, n_cells, L, id_to_pos, eps, h, theta, threads, "", "ADD_MIN");
        const double t2 = now();
        sig = 0;
        kept_entries = 0;
        for (const auto &chr : filtered) {
            sig += chr.size();
            for (const auto &pd : chr) {
                kept_entries += pd.size();
            }
        }
        checksum = m(0, 1) + m(n_cells / 2, n_cells - 1) + cov;
        if (r == 0) {
            first_filter = t1 - t0;
            first_sim = t2 - t1;
        } else {
            best_filter = std::min(best_filter, t1 - t0);
            best_sim = std::min(best_sim, t2 - t1);
        }
    }
    if (repeats < 2) {
        best_filter = first_filter;
        best_sim = first_sim;
    }
    std::printf("{\"what\": \"Filter::filter + computeSimilarityMatrix through the C++ shim: pageable vector<vector<PosData>> in, Matd out\", "
                "\"n_cells\": %u, \"coverage\": %g, \"prefilter_loci\": %llu, \"entries\": %llu, \"significant_loci\": %llu, \"kept_entries\": %llu, "
                "\"first_call_s\": {\"filter\": %.4f, \"similarity\": %.4f}, \"repeat_call_s\": {\"filter\": %.4f, \"similarity\": %.4f}, "
                "\"loci_per_s_first\": %.1f, \"loci_per_s_repeat\": %.1f, \"pileup_uploads\": %llu, \"host_generation_s\": %.2f, \"checksum\": %.9g, "
                "\"note\": \"first call: flatten + pageable upload of the whole pileup inside Filter::filter; repeat call: the shim recognises the "
                "same pds (address + fingerprint) and the pileup is already resident, as in divide_cluster's recursion\"}\n",
                n_cells, coverage, static_cast<unsigned long long>(static_cast<uint64_t>(n_chr) * loci_per_chr),
                static_cast<unsigned long long>(entries), static_cast<unsigned long long>(sig), static_cast<unsigned long long>(kept_entries),
                first_filter, first_sim, best_filter, best_sim, sig / (first_filter + first_sim), sig / (best_filter + best_sim),
                static_cast<unsigned long long>(secedo_b200_shim_uploads()), t_gen, checksum);
    return 0;
}
