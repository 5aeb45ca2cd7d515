/*
 * TEST INFRASTRUCTURE ONLY — CPU restatement (plain C) of the reference algorithm for SECEDO's
 * similarity-matrix hot path. Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * --impl reference legs may load this library, and only as the checker / reported baseline. The
 * product path (secedo_b200/) never links, imports or executes anything from oracle/.
 *
 * Pinning status: the filter half is pinned by the reference's own known-answer tests
 * (tests/test_is_significant.cpp:46-216, replayed in tests/test_oracle_filter.py). The reference
 * has NO test that pins the similarity matrix numerically; that half is pinned
 * against outputs of the unmodified reference compiled here (oracle/_ref, see Makefile) and
 * against committed fixtures generated from it (tests/golden/, tests/golden/make_golden.py).
 *
 * All functions take the flat CSR pileup layout documented in include/secedo_b200.h.
 */
#ifndef SECEDO_ORACLE_H
#define SECEDO_ORACLE_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define ORC_NO_POS 16383u /* util/is_significant.hpp:11 */
#define ORC_MAX_CLASS 64  /* x_s, x_d < 64 (Pascal triangle in uint64 stays exact to row 67) */

enum { ORC_ADD_MIN = 0, ORC_EXPONENTIATE = 1, ORC_SCALE_MAX_1 = 2 };

/* util/is_significant.cpp:78-138 */
int orc_is_significant(const uint16_t counts[4], double theta, int cell_proportion);

/*
 * util/is_significant.cpp:149-193. keep_locus[n_loci] / keep_entry[n_entries] are 0/1 flags
 * (an entry is kept iff its group is in the sub-cluster AND its locus is significant); the caller
 * compacts. avg_coverage reproduces the reference's 32-bit accumulation (:156,187); avg_coverage64
 * is the same quantity without wrap-around.
 */
int orc_filter(uint32_t n_chr,
               const uint64_t *chr_ptr,
               const uint64_t *row_ptr,
               const uint32_t *read_id,
               const uint16_t *gid_base,
               const uint32_t *id_to_pos,
               uint32_t n_groups,
               double theta,
               int cell_proportion,
               uint8_t *keep_locus,
               uint8_t *keep_entry,
               uint64_t *n_kept_loci,
               uint64_t *n_kept_entries,
               double *avg_coverage,
               double *avg_coverage64);

/* similarity_matrix.cpp:38-104,117-170: LS/LD for 0 <= x_s,x_d < n, row-major n*n. */
int orc_log_probs(double mutation_rate,
                  double homozygous_rate,
                  double seq_error_rate,
                  uint32_t max_fragment_length,
                  uint32_t n,
                  double *ls_out,
                  double *ld_out);

/*
 * similarity_matrix.cpp:295-433 at a stated num_threads (the result depends on it).
 * out_M: num_cells^2 doubles (normalised, as the reference returns it).
 * Optional instrumentation (each may be NULL):
 *   S1, D1   num_cells^2 int32: per cell pair, number of (counted read pair, shared locus)
 *            incidences with equal / different base (= sum of x_s / x_d over counted pairs)
 *   H        3*num_cells^2 int32: number of counted read pairs in class (2,0), (1,1), (0,2)
 *   class_hist ORC_MAX_CLASS^2 uint64: number of counted read pairs per (x_s, x_d), i<j once
 *   K_out    n_chr uint64: per-chromosome cutoff (reads with index >= K are never "read 1")
 *   raw_M    num_cells^2 doubles: mat_diff - mat_same before normalisation
 * Returns 0, or <0 on inputs outside the reference's defined behaviour (fragment span >= L,
 * overlap class >= ORC_MAX_CLASS or >= L, cell index >= num_cells).
 */
int orc_similarity(uint32_t n_chr,
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
                   int normalization,
                   double *out_M,
                   int32_t *S1,
                   int32_t *D1,
                   int32_t *H,
                   uint64_t *class_hist,
                   uint64_t *K_out,
                   double *raw_M);

/* The same two restatements on WIDE pileups: 32-bit entries (group id << 2 | base, 30-bit group ids) and 0xFFFFFFFF as
 * "not in the cluster". The reference itself cannot hold more than 16 383 groups (sequenced_data.hpp:29-36), so for
 * BASELINE config 5 (20 000 cells) the oracle is this restatement; below 16 384 groups it is checked to agree with the
 * 16-bit entry points, which are pinned against the compiled reference (tests/test_oracle_similarity.py). */
#define ORC_NO_POS_WIDE 0xFFFFFFFFu
int orc_filter_wide(uint32_t n_chr, const uint64_t *chr_ptr, const uint64_t *row_ptr, const uint32_t *read_id,
                    const uint32_t *gid_base32, const uint32_t *id_to_pos, uint32_t n_groups, double theta, int cell_proportion,
                    uint8_t *keep_locus, uint8_t *keep_entry, uint64_t *n_kept_loci, uint64_t *n_kept_entries,
                    double *avg_coverage, double *avg_coverage64);
int orc_similarity_wide(uint32_t n_chr, const uint64_t *chr_ptr, const uint64_t *row_ptr, const uint32_t *position,
                        const uint32_t *read_id, const uint32_t *gid_base32, uint32_t num_cells, uint32_t max_fragment_length,
                        const uint32_t *group_id_to_pos, uint32_t n_groups, double mutation_rate, double homozygous_rate,
                        double seq_error_rate, uint32_t num_threads, int normalization, double *out_M, int32_t *S1, int32_t *D1,
                        int32_t *H, uint64_t *class_hist, uint64_t *K_out, double *raw_M);

/* similarity_matrix.cpp:271-293 on a row-major n*n matrix, in place. */
int orc_normalize(int normalization, uint32_t n, double *m);

/* expectation_maximization(), expectation_maximization.cpp:131-160 (with cluster_center :19-40, maximization_step
 * :59-92, expectation_step :109-129), on the CSR form of the filtered pileup. prob_cluster_b (n_cells) is updated in
 * place. max_iterations = 0: until convergence like the reference. Returns -1 where the reference reads out of range. */
int orc_expectation_maximization(uint32_t n_chr, const uint64_t *chr_ptr, const uint64_t *row_ptr, const uint16_t *gid_base,
                                 const uint32_t *id_to_pos, uint32_t n_groups, double theta, uint32_t n_cells,
                                 double *prob_cluster_b, uint32_t max_iterations, uint32_t *iterations);

/* laplacian(), spectral_clustering.cpp:33-52, on row-major n*n matrices. */
int orc_laplacian(uint32_t n, const double *a, double *out);

#ifdef __cplusplus
}
#endif
#endif
