/*
 * TEST INFRASTRUCTURE ONLY — see secedo_oracle.h for the rules on who may use this file.
 *
 * Plain-C restatement of the reference hot path. Every function cites the reference lines it
 * follows (paths relative to the reference root). It is deliberately a *direct* restatement
 * (reads are rebuilt, read pairs are enumerated, LS/LD are added pair by pair into
 * mat_same/mat_diff) and NOT the count-matrix decomposition the CUDA path uses, so agreement
 * between the two is evidence and not tautology. The integer matrices it can emit on the side
 * (S1, D1, H, class_hist) are instrumentation of that enumeration.
 *
 * Compile with -ffp-contract=off so that products are evaluated like the reference's -O3 x86 build.
 */
#include "secedo_oracle.h"

#include <fenv.h>
#include <float.h>
#include <math.h>
#include <stdlib.h>
#include <string.h>

/* ------------------------------------------------------------------------------------------ */
/* Filter                                                                                      */
/* ------------------------------------------------------------------------------------------ */

/* util/is_significant.cpp:11-36 — thresholds for coverage 10,20,...,200 per cell proportion */
static const double Ks[5][20] = {
    { -1.64504967001201, -1.38868450353301, -1.38780664765677, -1.38779600211955,
      -1.3877952855556,  -1.38779524274215, -1.38779524274142, -1.38779524274141,
      -1.3877952427414,  -1.38779524274139, -1.38779524274138, -1.38779524274138,
      -1.38779524274138, -1.38779524274138, -1.38779524274139, -1.38780870444455,
      -1.38780870444455, -1.38780870444455, -1.38780870444455, -1.38780870444455 },
    { -1.56013904495168, -1.38819451352203, -1.38781438946096, -1.38779659244035,
      -1.38779537799054, -1.3877952484612,  -1.3877952427842,  -1.38779524274906,
      -1.3877952427457,  -1.38779524274275, -1.38780870444458, -1.38780870444459,
      -1.38780870444459, -1.38780870444469, -1.38780870444459, -1.42736056742577,
      -1.42736056742575, -6.19144172018466, -6.19144172018466, -14.1885779508362 },
    { -1.47780038365618, -1.3885722463397,  -1.38781428162649, -1.3877984410546,
      -1.38779548312685, -1.3877952855556,  -1.38779524455204, -1.38779524331456,
      -1.38780873675669, -1.38780870687333, -1.42737804009806, -6.19144172131432,
      -14.1885779508648, -6.1914418045659,  -30.1993093269287, -30.1993093268559,
      -30.1993093268539, -54.2154105288032, -62.2207775961199, -46.2100434614866 },
    { -1.47780038365618, -1.38868450353301, -1.38782829051844, -1.3877984410546,
      -1.38779625512927, -1.38779556321717, -1.38780972304588, -1.3878087226245,
      -6.21747860711653, -22.1939432034943, -14.1886670526002, -22.1939422903721,
      -46.2100434614866, -54.2154105288069, -70.2261446634366, -62.2207775961199,
      -86.2368787980699, -110.25298000002,  -118.258347067337, -102.247612932703 },
    { -1.52859626647315, -1.38967447346712, -1.38787138908447, -1.38780282263764,
      -1.387805349423,   -1.38882047800373, -1.49793700616569, -6.19975747800726,
      -22.197881249831,  -38.2046765807324, -38.2046769835162, -70.2261446634383,
      -54.2154105303641, -78.2315117307532, -86.2368787980699, -118.258347067337,
      -126.263714134653, -134.26908120197,  -158.28518240392,  -158.28518240392 }
};

static void sort4(uint16_t b[4]) { /* ascending, util/is_significant.cpp:88 */
    for (int i = 1; i < 4; ++i) {
        uint16_t v = b[i];
        int j = i - 1;
        while (j >= 0 && b[j] > v) {
            b[j + 1] = b[j];
            --j;
        }
        b[j + 1] = v;
    }
}

int orc_is_significant(const uint16_t counts[4], double theta, int cell_proportion) {
    /* util/is_significant.cpp:78-138 */
    uint16_t b[4] = { counts[0], counts[1], counts[2], counts[3] };
    uint32_t coverage = (uint32_t)b[0] + b[1] + b[2] + b[3];
    if (coverage < 2) {
        return 0;
    }
    sort4(b);
    if (b[2] == 0) {
        return 0;
    }
    if (b[2] + b[1] + b[0] < 5) {
        return 0;
    }
    if (b[3] < 1.5 * b[2]) {
        return 0;
    }
    fesetround(FE_TONEAREST); /* :67-70 round half to even */
    double t = nearbyint(coverage / 10.) - 1;
    if (t < 0.) {
        t = 0.;
    }
    if (t > 19.) {
        t = 19.;
    }
    uint32_t threshold_idx = (uint32_t)t;

    /* :39-45 priors; note log_homo_prior = log(hetero_prior) (a quirk that is reproduced) */
    const double hetero_prior = 0.0005;
    const double mut_prior = 1e-6;
    const double homo_prior = 1 - hetero_prior - mut_prior;
    const double log_homo_prior = log(hetero_prior);
    const double log_1_4 = log(1. / 4);
    const double log_theta_3 = log(theta / 3);
    const double log_one_minus_theta = log(1 - theta);

    double log_prob_homozygous = b[3] * log_one_minus_theta + (coverage - b[3]) * log_theta_3;
    log_prob_homozygous += log_1_4;
    log_prob_homozygous += log_homo_prior;

    double prob_all_c1 = homo_prior * pow(1 - theta, b[3]) * pow(theta / 3, coverage - b[3]);
    double prob_hetero
            = hetero_prior * pow(0.5 - theta / 3, b[3] + b[2]) * pow(theta / 3, b[0] + b[1]);
    double prob_homo_som = homo_prior * mut_prior * pow(0.75 - 2 * theta / 3, b[3])
            * pow(0.25, b[2]) * pow(theta / 3, b[0] + b[1]);
    double prob_hetero_som = hetero_prior * mut_prior * pow(0.5 - theta, b[3])
            * pow(0.25, b[1] + b[2]) * pow(theta / 3, b[0]);
    double prob_two_somatic = hetero_prior * mut_prior * mut_prior * pow(1 - theta, coverage);
    double log_evidence = log(prob_all_c1 + prob_hetero + prob_homo_som + prob_hetero_som
                              + prob_two_somatic);
    return log_prob_homozygous - log_evidence < Ks[cell_proportion][threshold_idx];
}

/* group id << 2 | base of entry e: 16-bit entries (the reference's PosData) or, for pileups beyond its 14-bit group
 * ids (BASELINE config 5), 32-bit ones; exactly one of the two arrays is given */
static inline uint32_t entry_at(const uint16_t *gb16, const uint32_t *gb32, uint64_t e) {
    return gb32 ? gb32[e] : gb16[e];
}

static int filter_impl(uint32_t n_chr,
               const uint64_t *chr_ptr,
               const uint64_t *row_ptr,
               const uint32_t *read_id,
               const uint16_t *gid_base16,
               const uint32_t *gid_base32,
               uint32_t no_pos,
               const uint32_t *id_to_pos,
               uint32_t n_groups,
               double theta,
               int cell_proportion,
               uint8_t *keep_locus,
               uint8_t *keep_entry,
               uint64_t *n_kept_loci,
               uint64_t *n_kept_entries,
               double *avg_coverage,
               double *avg_coverage64) {
    /* util/is_significant.cpp:149-193 */
    (void)read_id;
    uint64_t loci64 = 0, entries64 = 0;
    int32_t total_coverage_i = 0; /* std::accumulate(..., 0) sums in int (:187) */
    uint32_t total_positions = 0; /* std::atomic<uint32_t> (:158) */
    for (uint32_t c = 0; c < n_chr; ++c) {
        uint32_t coverage_chr = 0; /* :156 */
        uint32_t kept_here = 0;
        for (uint64_t l = chr_ptr[c]; l < chr_ptr[c + 1]; ++l) {
            uint16_t bc[4] = { 0, 0, 0, 0 };
            uint64_t n_in = 0;
            for (uint64_t e = row_ptr[l]; e < row_ptr[l + 1]; ++e) {
                uint32_t gid = entry_at(gid_base16, gid_base32, e) >> 2;
                if (gid >= n_groups) {
                    return -1;
                }
                if (id_to_pos[gid] == no_pos) { /* :169 */
                    keep_entry[e] = 0;
                    continue;
                }
                keep_entry[e] = 1;
                bc[entry_at(gid_base16, gid_base32, e) & 3]++;
                n_in++;
            }
            int sig = orc_is_significant(bc, theta, cell_proportion);
            keep_locus[l] = (uint8_t)sig;
            if (sig) {
                coverage_chr += (uint32_t)n_in;
                kept_here++;
                loci64++;
                entries64 += n_in;
            } else {
                for (uint64_t e = row_ptr[l]; e < row_ptr[l + 1]; ++e) {
                    keep_entry[e] = 0;
                }
            }
        }
        total_positions += kept_here;
        total_coverage_i = (int32_t)((uint32_t)total_coverage_i + coverage_chr);
    }
    uint32_t total_coverage = (uint32_t)total_coverage_i;
    *n_kept_loci = loci64;
    *n_kept_entries = entries64;
    *avg_coverage = total_positions == 0 ? 0 : (double)total_coverage / total_positions; /* :188 */
    *avg_coverage64 = loci64 == 0 ? 0 : (double)entries64 / (double)loci64;
    return 0;
}

int orc_filter(uint32_t n_chr, const uint64_t *chr_ptr, const uint64_t *row_ptr, const uint32_t *read_id,
               const uint16_t *gid_base, const uint32_t *id_to_pos, uint32_t n_groups, double theta, int cell_proportion,
               uint8_t *keep_locus, uint8_t *keep_entry, uint64_t *n_kept_loci, uint64_t *n_kept_entries,
               double *avg_coverage, double *avg_coverage64) {
    return filter_impl(n_chr, chr_ptr, row_ptr, read_id, gid_base, NULL, ORC_NO_POS, id_to_pos, n_groups, theta, cell_proportion,
                       keep_locus, keep_entry, n_kept_loci, n_kept_entries, avg_coverage, avg_coverage64);
}

int orc_filter_wide(uint32_t n_chr, const uint64_t *chr_ptr, const uint64_t *row_ptr, const uint32_t *read_id,
                    const uint32_t *gid_base32, const uint32_t *id_to_pos, uint32_t n_groups, double theta, int cell_proportion,
                    uint8_t *keep_locus, uint8_t *keep_entry, uint64_t *n_kept_loci, uint64_t *n_kept_entries,
                    double *avg_coverage, double *avg_coverage64) {
    return filter_impl(n_chr, chr_ptr, row_ptr, read_id, NULL, gid_base32, ORC_NO_POS_WIDE, id_to_pos, n_groups, theta,
                       cell_proportion, keep_locus, keep_entry, n_kept_loci, n_kept_entries, avg_coverage, avg_coverage64);
}

/* ------------------------------------------------------------------------------------------ */
/* Log-probability tables                                                                      */
/* ------------------------------------------------------------------------------------------ */

#define TBL 132 /* power tables / Pascal rows kept: enough for x_s + x_d <= 2*(ORC_MAX_CLASS-1) */

typedef struct {
    /* similarity_matrix.cpp:38-67 */
    double pss[TBL], psd[TBL], pds[TBL], pdd[TBL];
    double p1he[TBL], p1he2[TBL], phe2[TBL], ph[TBL], pe[TBL], p05[TBL], psspds[TBL], psdpdd[TBL];
    uint64_t comb[TBL][TBL];
    uint32_t len; /* number of valid entries (= max_read_size, capped at TBL) */
    double ls[ORC_MAX_CLASS][ORC_MAX_CLASS];
    double ld[ORC_MAX_CLASS][ORC_MAX_CLASS];
    uint32_t L;
} cache_t;

static void extend(double *a, uint32_t n) { /* similarity_matrix.cpp:81: a[p] = a[p-1] * a[1] */
    for (uint32_t p = 2; p < n; ++p) {
        a[p] = a[p - 1] * a[1];
    }
}

static void cache_init(cache_t *c, double epsilon, double h, double theta, uint32_t max_read_size) {
    /* similarity_matrix.cpp:42-103 */
    const double theta2 = theta * theta;
    const double p_same_diff = 2 * theta * (1 - theta) + 2 * theta2 / 3;
    const double p_same_same = 1 - p_same_diff;
    const double p_diff_same = 2 * (1 - theta) * theta / 3 + 2 * theta2 / 9;
    const double p_diff_diff = 1 - p_diff_same;
    uint32_t n = max_read_size < 2 ? 2 : max_read_size;
    if (n > TBL) {
        n = TBL;
    }
    c->len = n;
    c->L = max_read_size;
#define INIT(arr, v) \
    do { \
        (arr)[0] = 1; \
        (arr)[1] = (v); \
        extend((arr), n); \
    } while (0)
    INIT(c->pss, p_same_same);
    INIT(c->psd, p_same_diff);
    INIT(c->pds, p_diff_same);
    INIT(c->pdd, p_diff_diff);
    INIT(c->p1he, 1 - epsilon - h);
    INIT(c->p1he2, 1 - epsilon * 0.5 - h);
    INIT(c->phe2, h + epsilon * 0.5);
    INIT(c->ph, h);
    INIT(c->pe, epsilon);
    INIT(c->p05, 0.5);
    INIT(c->psspds, p_same_same + p_diff_same);
    INIT(c->psdpdd, p_same_diff + p_diff_diff);
#undef INIT
    memset(c->comb, 0, sizeof(c->comb));
    c->comb[0][0] = 1;
    for (uint32_t r = 1; r < TBL; ++r) { /* Pascal triangle in uint64, wraps like the reference */
        c->comb[r][0] = 1;
        c->comb[r][r] = 1;
        for (uint32_t i = 1; i < r; ++i) {
            c->comb[r][i] = c->comb[r - 1][i - 1] + c->comb[r - 1][i];
        }
    }
    for (uint32_t s = 0; s < ORC_MAX_CLASS; ++s) {
        for (uint32_t d = 0; d < ORC_MAX_CLASS; ++d) {
            c->ls[s][d] = DBL_MAX; /* :314-317 sentinel */
            c->ld[s][d] = DBL_MAX;
        }
    }
}

static double log_prob_diff_genotype(uint32_t x_s, uint32_t x_d, cache_t *c) {
    /* similarity_matrix.cpp:117-141 */
    if (c->ld[x_s][x_d] != DBL_MAX) {
        return c->ld[x_s][x_d];
    }
    double prob = 0;
    for (uint32_t k = 0; k <= x_s; ++k) {
        for (uint32_t l = 0; l <= x_d; ++l) {
            for (uint32_t p = 0; p <= x_s - k; ++p) {
                for (uint32_t q = 0; q <= x_d - l; ++q) {
                    uint64_t cc = c->comb[x_s][k] * c->comb[x_d][l] * c->comb[x_s - k][p]
                            * c->comb[x_d - l][q];
                    uint32_t rest = x_s + x_d - k - l - p - q;
                    prob += cc * c->p1he[k + l] * 0.5
                            * (c->pss[k] * c->psd[l] + c->pds[k] * c->pdd[l]) * c->pe[rest]
                            * c->p05[rest] * c->psspds[x_s - k - p] * c->psdpdd[x_d - l - q]
                            * c->ph[p + q] * c->pss[p] * c->psd[q];
                }
            }
        }
    }
    prob *= c->comb[x_s + x_d][x_s];
    c->ld[x_s][x_d] = log(prob);
    return c->ld[x_s][x_d];
}

static double log_prob_same_genotype(uint32_t x_s, uint32_t x_d, cache_t *c) {
    /* similarity_matrix.cpp:153-170 */
    if (c->ls[x_s][x_d] != DBL_MAX) {
        return c->ls[x_s][x_d];
    }
    double p = 0;
    for (uint32_t k = 0; k <= x_s; ++k) {
        for (uint32_t l = 0; l <= x_d; ++l) {
            uint64_t cc = c->comb[x_s][k] * c->comb[x_d][l];
            p += cc * c->p1he2[k + l] * 0.5 * (c->pss[k] * c->psd[l] + c->pds[k] * c->pdd[l])
                    * c->phe2[x_s + x_d - k - l] * c->pss[x_s - k] * c->psd[x_d - l];
        }
    }
    p *= c->comb[x_s + x_d][x_s];
    c->ls[x_s][x_d] = log(p);
    return c->ls[x_s][x_d];
}

int orc_log_probs(double mutation_rate,
                  double homozygous_rate,
                  double seq_error_rate,
                  uint32_t max_fragment_length,
                  uint32_t n,
                  double *ls_out,
                  double *ld_out) {
    if (n > ORC_MAX_CLASS || n > max_fragment_length || 2 * n > TBL + 1) {
        return -1;
    }
    cache_t *c = (cache_t *)malloc(sizeof(cache_t));
    cache_init(c, mutation_rate, homozygous_rate, seq_error_rate, max_fragment_length);
    for (uint32_t s = 0; s < n; ++s) {
        for (uint32_t d = 0; d < n; ++d) {
            if (s + d >= c->len) { /* the reference would index past its power tables */
                ls_out[s * n + d] = NAN;
                ld_out[s * n + d] = NAN;
                continue;
            }
            ls_out[s * n + d] = log_prob_same_genotype(s, d, c);
            ld_out[s * n + d] = log_prob_diff_genotype(s, d, c);
        }
    }
    free(c);
    return 0;
}

/* ------------------------------------------------------------------------------------------ */
/* Normalisation                                                                               */
/* ------------------------------------------------------------------------------------------ */

int orc_normalize(int normalization, uint32_t n, double *m) {
    /* similarity_matrix.cpp:271-293; min/max scan the full matrix incl. diagonal, mat.hpp:35-53 */
    uint64_t nn = (uint64_t)n * n;
    if (nn == 0) {
        return 0;
    }
    switch (normalization) {
        case ORC_ADD_MIN: {
            for (uint64_t i = 0; i < nn; ++i) {
                m[i] *= -1;
            }
            double mn = m[0];
            for (uint64_t i = 1; i < nn; ++i) {
                if (m[i] < mn) {
                    mn = m[i];
                }
            }
            double a = fabs(mn);
            for (uint64_t i = 0; i < nn; ++i) {
                m[i] += a;
            }
            break;
        }
        case ORC_EXPONENTIATE:
            for (uint64_t i = 0; i < nn; ++i) {
                m[i] = 1. / (exp(m[i]) + 1);
            }
            break;
        case ORC_SCALE_MAX_1: {
            for (uint32_t i = 0; i < n; ++i) {
                m[(uint64_t)i * n + i] = 0;
            }
            double mx = m[0];
            for (uint64_t i = 1; i < nn; ++i) {
                if (m[i] > mx) {
                    mx = m[i];
                }
            }
            double f = 1. / mx;
            for (uint64_t i = 0; i < nn; ++i) {
                m[i] *= f;
            }
            break;
        }
        default:
            return -1; /* the reference throws std::logic_error (:264) */
    }
    for (uint32_t i = 0; i < n; ++i) {
        m[(uint64_t)i * n + i] = 0; /* :292 */
    }
    return 0;
}

/* ------------------------------------------------------------------------------------------ */
/* Similarity matrix                                                                           */
/* ------------------------------------------------------------------------------------------ */

typedef struct {
    uint32_t *keys; /* read_id + 1, 0 = empty */
    uint32_t *vals; /* read index */
    uint64_t mask;
} map_t;

static void map_init(map_t *m, uint64_t n) {
    uint64_t cap = 16;
    while (cap < 2 * n + 2) {
        cap <<= 1;
    }
    m->keys = (uint32_t *)calloc(cap, sizeof(uint32_t));
    m->vals = (uint32_t *)malloc(cap * sizeof(uint32_t));
    m->mask = cap - 1;
}

static void map_free(map_t *m) {
    free(m->keys);
    free(m->vals);
}

/* returns the slot of `key`, claiming an empty one (vals = UINT32_MAX) when absent; read ids
 * equal to UINT32_MAX would alias the empty marker and are rejected by the caller */
static uint64_t map_slot(map_t *m, uint32_t key) {
    uint64_t h = ((uint64_t)key * 0x9E3779B97F4A7C15ull) >> 20;
    for (;;) {
        h &= m->mask;
        if (m->keys[h] == key + 1) {
            return h;
        }
        if (m->keys[h] == 0) {
            m->keys[h] = key + 1;
            m->vals[h] = UINT32_MAX;
            return h;
        }
        ++h;
    }
}

typedef struct {
    uint32_t cell;  /* group_id_to_pos[group id of the read's FIRST entry] (:208-209,379) */
    uint32_t start; /* position of the first entry, kept even if that base is later removed */
    uint64_t off;   /* offset of this read's (pos, base) list in the pools */
    uint32_t cap;   /* number of pileup entries carrying this read id */
    uint32_t n;     /* number of currently stored (pos, base) */
} read_t;

static int similarity_impl(uint32_t n_chr,
                   const uint64_t *chr_ptr,
                   const uint64_t *row_ptr,
                   const uint32_t *position,
                   const uint32_t *read_id,
                   const uint16_t *gid_base16,
                   const uint32_t *gid_base32,
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
                   double *raw_M) {
    const uint64_t NN = (uint64_t)num_cells * num_cells;
    const uint32_t L = max_fragment_length;
    int rc = 0;
    if (normalization < 0 || normalization > 2) {
        return -1;
    }
    double *mat_same = (double *)calloc(NN ? NN : 1, sizeof(double)); /* :306-307 */
    double *mat_diff = (double *)calloc(NN ? NN : 1, sizeof(double));
    cache_t *cache = (cache_t *)malloc(sizeof(cache_t));
    cache_init(cache, mutation_rate, homozygous_rate, seq_error_rate, L); /* :330 */
    if (S1) {
        memset(S1, 0, NN * sizeof(int32_t));
    }
    if (D1) {
        memset(D1, 0, NN * sizeof(int32_t));
    }
    if (H) {
        memset(H, 0, 3 * NN * sizeof(int32_t));
    }
    if (class_hist) {
        memset(class_hist, 0, sizeof(uint64_t) * ORC_MAX_CLASS * ORC_MAX_CLASS);
    }

    for (uint32_t c = 0; c < n_chr && rc == 0; ++c) { /* :345; all state is per chromosome (:407-408) */
        const uint64_t l0 = chr_ptr[c], l1 = chr_ptr[c + 1];
        const uint64_t e0 = row_ptr[l0], e1 = row_ptr[l1];
        const uint64_t ne = e1 - e0;
        if (K_out) {
            K_out[c] = 0;
        }
        if (ne == 0) {
            continue;
        }
        /* ---- pass 1: read index = order of first appearance of the read id (:377-382) ---- */
        map_t map;
        map_init(&map, ne);
        uint32_t *eread = (uint32_t *)malloc(ne * sizeof(uint32_t));
        read_t *reads = (read_t *)malloc(ne * sizeof(read_t));
        uint32_t n_reads = 0;
        for (uint64_t l = l0; l < l1 && rc == 0; ++l) {
            if (l > l0 && position[l] <= position[l - 1]) {
                rc = -6; /* loci must be strictly increasing inside a chromosome (:398 assert) */
                break;
            }
            for (uint64_t e = row_ptr[l]; e < row_ptr[l + 1]; ++e) {
                if (read_id[e] == UINT32_MAX) {
                    rc = -5;
                    break;
                }
                uint64_t s = map_slot(&map, read_id[e]);
                if (map.vals[s] == UINT32_MAX) {
                    uint32_t gid = entry_at(gid_base16, gid_base32, e) >> 2;
                    if (gid >= n_groups || group_id_to_pos[gid] >= num_cells) {
                        rc = -2; /* the reference would index out of bounds (mat.hpp:119) */
                        break;
                    }
                    map.vals[s] = n_reads;
                    reads[n_reads].cell = group_id_to_pos[gid];
                    reads[n_reads].start = position[l];
                    reads[n_reads].cap = 0;
                    reads[n_reads].n = 0;
                    n_reads++;
                }
                eread[e - e0] = map.vals[s];
                reads[map.vals[s]].cap++;
            }
        }
        map_free(&map);
        if (rc) {
            free(eread);
            free(reads);
            break;
        }
        uint64_t acc = 0;
        for (uint32_t r = 0; r < n_reads; ++r) {
            reads[r].off = acc;
            acc += reads[r].cap;
        }
        uint32_t *rl_locus = (uint32_t *)malloc(ne * sizeof(uint32_t)); /* locus index in chr */
        uint8_t *rl_base = (uint8_t *)malloc(ne);
        /* ---- pass 2: replay the insertion rules incl. the paired-end mate handling (:383-402) */
        for (uint64_t l = l0; l < l1 && rc == 0; ++l) {
            for (uint64_t e = row_ptr[l]; e < row_ptr[l + 1]; ++e) {
                read_t *rd = &reads[eread[e - e0]];
                uint8_t base = entry_at(gid_base16, gid_base32, e) & 3;
                if (position[l] - rd->start >= L) {
                    rc = -3; /* fragment longer than max_fragment_length: the reference would
                                have retired the read and the outcome depends on batch timing */
                    break;
                }
                if (rd->n > 0 && position[l0 + rl_locus[rd->off + rd->n - 1]] == position[l]) {
                    /* two reads at the same position (:387-395) */
                    if (rl_base[rd->off + rd->n - 1] != base) {
                        rd->n--; /* mates disagree: drop the stored base as well */
                    }
                    continue;
                }
                rl_locus[rd->off + rd->n] = (uint32_t)(l - l0);
                rl_base[rd->off + rd->n] = base;
                rd->n++;
            }
        }
        /* ---- cutoff K (:348-373 and the tail drop at :407-408) ---- */
        uint64_t front = 0;
        {
            uint64_t created = 0; /* reads inserted before the current locus */
            uint64_t u = 0;       /* reads (by index) with start + L <= p */
            uint32_t *first_seen = (uint32_t *)calloc(n_reads ? n_reads : 1, sizeof(uint32_t));
            for (uint64_t l = l0; l < l1; ++l) {
                uint32_t p = position[l];
                while (u < created && (uint64_t)reads[u].start + L <= p) {
                    ++u;
                }
                uint64_t completed = u > front ? u - front : 0;
                if (completed >= 4ull * num_threads) { /* BATCH_SIZE * num_threads (:354-356) */
                    front = u;
                }
                for (uint64_t e = row_ptr[l]; e < row_ptr[l + 1]; ++e) {
                    uint32_t r = eread[e - e0];
                    if (!first_seen[r]) {
                        first_seen[r] = 1;
                        created++;
                    }
                }
            }
            free(first_seen);
        }
        const uint64_t K = front;
        if (K_out) {
            K_out[c] = K;
        }
        /* ---- per-locus lists of surviving (read, base), in read-index order ---- */
        const uint64_t nl = l1 - l0;
        uint64_t *lptr = (uint64_t *)calloc(nl + 1, sizeof(uint64_t));
        for (uint32_t r = 0; r < n_reads; ++r) {
            for (uint32_t i = 0; i < reads[r].n; ++i) {
                lptr[rl_locus[reads[r].off + i] + 1]++;
            }
        }
        for (uint64_t l = 0; l < nl; ++l) {
            lptr[l + 1] += lptr[l];
        }
        uint32_t *lread = (uint32_t *)malloc((lptr[nl] ? lptr[nl] : 1) * sizeof(uint32_t));
        uint8_t *lbase = (uint8_t *)malloc(lptr[nl] ? lptr[nl] : 1);
        uint64_t *cur = (uint64_t *)malloc((nl + 1) * sizeof(uint64_t));
        memcpy(cur, lptr, (nl + 1) * sizeof(uint64_t));
        for (uint32_t r = 0; r < n_reads; ++r) {
            for (uint32_t i = 0; i < reads[r].n; ++i) {
                uint64_t l = rl_locus[reads[r].off + i];
                lread[cur[l]] = r;
                lbase[cur[l]] = rl_base[reads[r].off + i];
                cur[l]++;
            }
        }
        free(cur);
        /* ---- read pairs (:189-243). Pair (r1 < r2) is evaluated when r1 is processed, which
         * happens iff r1 < K; it is accounted at the first locus the two reads share. ---- */
        for (uint64_t l = 0; l < nl && rc == 0; ++l) {
            for (uint64_t a = lptr[l]; a < lptr[l + 1] && rc == 0; ++a) {
                const uint32_t r1 = lread[a];
                if (r1 >= K) {
                    break; /* list is in read order: all later r1 are >= K too */
                }
                const read_t *R1 = &reads[r1];
                for (uint64_t b = a + 1; b < lptr[l + 1]; ++b) {
                    const uint32_t r2 = lread[b];
                    const read_t *R2 = &reads[r2];
                    if (R1->cell == R2->cell) { /* :215 */
                        continue;
                    }
                    const int same_here = lbase[a] == lbase[b];
                    const uint64_t ij = (uint64_t)R1->cell * num_cells + R2->cell;
                    const uint64_t ji = (uint64_t)R2->cell * num_cells + R1->cell;
                    if (S1 && same_here) {
                        S1[ij]++;
                        S1[ji]++;
                    }
                    if (D1 && !same_here) {
                        D1[ij]++;
                        D1[ji]++;
                    }
                    uint32_t x_s, x_d;
                    if (R1->n == 1 || R2->n == 1) {
                        x_s = same_here;
                        x_d = !same_here;
                    } else {
                        /* two-pointer merge over the sorted position lists (:221-229) */
                        x_s = x_d = 0;
                        uint32_t first_common = UINT32_MAX;
                        for (uint32_t i1 = 0, i2 = 0; i1 < R1->n && i2 < R2->n;) {
                            uint32_t p1 = rl_locus[R1->off + i1], p2 = rl_locus[R2->off + i2];
                            if (p1 == p2) {
                                if (first_common == UINT32_MAX) {
                                    first_common = p1;
                                }
                                rl_base[R1->off + i1] == rl_base[R2->off + i2] ? x_s++ : x_d++;
                                i1++;
                                i2++;
                            } else if (p1 < p2) {
                                i1++;
                            } else {
                                i2++;
                            }
                        }
                        if (first_common != (uint32_t)l) {
                            continue; /* this pair was accounted at an earlier shared locus */
                        }
                    }
                    if (x_s >= ORC_MAX_CLASS || x_d >= ORC_MAX_CLASS || x_s >= L || x_d >= L
                        || x_s + x_d >= cache->len) {
                        rc = -4;
                        break;
                    }
                    if (class_hist) {
                        class_hist[x_s * ORC_MAX_CLASS + x_d]++;
                    }
                    if (H && x_s + x_d == 2) {
                        int32_t *h = H + (uint64_t)x_d * NN; /* (2,0)->0, (1,1)->1, (0,2)->2 */
                        h[ij]++;
                        h[ji]++;
                    }
                    /* :238-241 then apply_updates (:246-254) */
                    double ls = log_prob_same_genotype(x_s, x_d, cache);
                    double ld = log_prob_diff_genotype(x_s, x_d, cache);
                    mat_same[ij] += ls;
                    mat_same[ji] = mat_same[ij];
                    mat_diff[ij] += ld;
                    mat_diff[ji] = mat_diff[ij];
                }
            }
        }
        free(lptr);
        free(lread);
        free(lbase);
        free(rl_locus);
        free(rl_base);
        free(eread);
        free(reads);
    }

    if (rc == 0) {
        for (uint64_t i = 0; i < NN; ++i) {
            mat_diff[i] -= mat_same[i]; /* :428 */
        }
        if (raw_M) {
            memcpy(raw_M, mat_diff, NN * sizeof(double));
        }
        rc = orc_normalize(normalization, num_cells, mat_diff); /* :430 */
        if (out_M) {
            memcpy(out_M, mat_diff, NN * sizeof(double));
        }
    }
    free(mat_same);
    free(mat_diff);
    free(cache);
    return rc;
}

int orc_similarity(uint32_t n_chr, const uint64_t *chr_ptr, const uint64_t *row_ptr, const uint32_t *position,
                   const uint32_t *read_id, const uint16_t *gid_base, uint32_t num_cells, uint32_t max_fragment_length,
                   const uint32_t *group_id_to_pos, uint32_t n_groups, double mutation_rate, double homozygous_rate,
                   double seq_error_rate, uint32_t num_threads, int normalization, double *out_M, int32_t *S1, int32_t *D1,
                   int32_t *H, uint64_t *class_hist, uint64_t *K_out, double *raw_M) {
    return similarity_impl(n_chr, chr_ptr, row_ptr, position, read_id, gid_base, NULL, num_cells, max_fragment_length,
                           group_id_to_pos, n_groups, mutation_rate, homozygous_rate, seq_error_rate, num_threads, normalization,
                           out_M, S1, D1, H, class_hist, K_out, raw_M);
}

/* the same restatement on a pileup with 30-bit group ids (beyond what the reference's PosData can hold) */
int orc_similarity_wide(uint32_t n_chr, const uint64_t *chr_ptr, const uint64_t *row_ptr, const uint32_t *position,
                        const uint32_t *read_id, const uint32_t *gid_base32, uint32_t num_cells, uint32_t max_fragment_length,
                        const uint32_t *group_id_to_pos, uint32_t n_groups, double mutation_rate, double homozygous_rate,
                        double seq_error_rate, uint32_t num_threads, int normalization, double *out_M, int32_t *S1, int32_t *D1,
                        int32_t *H, uint64_t *class_hist, uint64_t *K_out, double *raw_M) {
    return similarity_impl(n_chr, chr_ptr, row_ptr, position, read_id, NULL, gid_base32, num_cells, max_fragment_length,
                           group_id_to_pos, n_groups, mutation_rate, homozygous_rate, seq_error_rate, num_threads, normalization,
                           out_M, S1, D1, H, class_hist, K_out, raw_M);
}

/* ------------------------------------------------------------------------------------------ */
/* Graph Laplacian                                                         */
/* ------------------------------------------------------------------------------------------ */

int orc_laplacian(uint32_t n, const double *a, double *out) {
    /* spectral_clustering.cpp:33-52: degrees summed along the row in index order, 1/sqrt (0 stays 0),
     * result(r,c) = (r == c) - diag[r] * diag[c] * a(r,c), lower triangle mirrored */
    double *diag = (double *)calloc(n ? n : 1, sizeof(double));
    if (!diag) {
        return -1;
    }
    for (uint32_t r = 0; r < n; ++r) {
        for (uint32_t c = 0; c < n; ++c) {
            diag[r] += a[(uint64_t)r * n + c];
        }
    }
    for (uint32_t r = 0; r < n; ++r) {
        diag[r] = diag[r] == 0 ? 0 : 1 / sqrt(diag[r]);
    }
    for (uint32_t r = 0; r < n; ++r) {
        for (uint32_t c = 0; c <= r; ++c) {
            double v = (r == c ? 1 : 0) - diag[r] * diag[c] * a[(uint64_t)r * n + c];
            out[(uint64_t)r * n + c] = v;
            out[(uint64_t)c * n + r] = v;
        }
    }
    free(diag);
    return 0;
}

/* ------------------------------------------------------------------------------------------ */
/* Expectation maximisation                                                */
/* ------------------------------------------------------------------------------------------ */

static void orc_cluster_center(const uint16_t *gb, uint64_t n, const double *prob, double theta, double *center) {
    /* expectation_maximization.cpp:19-40; prob is indexed by the GROUP id (:24), not by id_to_pos */
    center[0] = center[1] = center[2] = center[3] = 0;
    for (uint64_t i = 0; i < n; ++i) {
        center[gb[i] & 3] += prob[gb[i] >> 2];
    }
    double s = center[0] + center[1] + center[2] + center[3];
    if (s == 0) {
        center[0] = center[1] = center[2] = center[3] = log(0.25);
        return;
    }
    for (int b = 0; b < 4; ++b) {
        center[b] = center[b] / s > theta ? center[b] / s : theta;
    }
    s = center[0] + center[1] + center[2] + center[3];
    for (int b = 0; b < 4; ++b) {
        center[b] = log(center[b] / s);
    }
}

int orc_expectation_maximization(uint32_t n_chr, const uint64_t *chr_ptr, const uint64_t *row_ptr, const uint16_t *gid_base,
                                 const uint32_t *id_to_pos, uint32_t n_groups, double theta, uint32_t n_cells,
                                 double *prob_cluster_b, uint32_t max_iterations, uint32_t *iterations) {
    /* expectation_maximization.cpp:131-160: the log likelihoods are NOT reset between iterations (:136-137 are
     * outside the loop), per-chromosome sums are added chromosome by chromosome (:146-151) */
    double *lla = (double *)calloc(n_cells ? n_cells : 1, sizeof(double));
    double *llb = (double *)calloc(n_cells ? n_cells : 1, sizeof(double));
    double *ca = (double *)calloc(n_cells ? n_cells : 1, sizeof(double));
    double *cb = (double *)calloc(n_cells ? n_cells : 1, sizeof(double));
    double *pa = (double *)calloc(n_cells ? n_cells : 1, sizeof(double));
    uint32_t it = 0;
    int rc = 0, done = 0;
    while (!done && rc == 0) {
        for (uint32_t c = 0; c < n_chr && rc == 0; ++c) {
            for (uint32_t i = 0; i < n_cells; ++i) { /* maximization_step :59-92 */
                pa[i] = 1 - prob_cluster_b[i];
                ca[i] = cb[i] = 0;
            }
            for (uint64_t l = chr_ptr[c]; l < chr_ptr[c + 1] && rc == 0; ++l) {
                const uint16_t *gb = gid_base + row_ptr[l];
                uint64_t n = row_ptr[l + 1] - row_ptr[l];
                for (uint64_t i = 0; i < n; ++i) {
                    uint32_t g = gb[i] >> 2;
                    if (g >= n_cells || g >= n_groups || id_to_pos[g] >= n_cells) {
                        rc = -1; /* out-of-range read / std::out_of_range in the reference */
                        break;
                    }
                }
                if (rc) {
                    break;
                }
                double center_a[4], center_b[4];
                orc_cluster_center(gb, n, pa, theta, center_a);
                orc_cluster_center(gb, n, prob_cluster_b, theta, center_b);
                for (uint64_t i = 0; i < n; ++i) {
                    ca[id_to_pos[gb[i] >> 2]] += center_a[gb[i] & 3];
                    cb[id_to_pos[gb[i] >> 2]] += center_b[gb[i] & 3];
                }
            }
            for (uint32_t i = 0; i < n_cells; ++i) {
                lla[i] += ca[i];
                llb[i] += cb[i];
            }
        }
        if (rc) {
            break;
        }
        /* expectation_step :109-129 */
        double sum = 0;
        for (uint32_t i = 0; i < n_cells; ++i) {
            sum += prob_cluster_b[i];
        }
        double prior_b = sum / n_cells, prior_a = 1 - prior_b;
        done = 1;
        for (uint32_t i = 0; i < n_cells; ++i) {
            double d = llb[i] - lla[i];
            d = d < -100. ? -100. : (d > 100. ? 100. : d);
            double odds = exp(d);
            double prob = 1 - 1 / (1 + odds * prior_b / prior_a);
            done &= fabs(prob - prob_cluster_b[i]) < 1e-2;
            prob_cluster_b[i] = prob;
        }
        ++it;
        if (max_iterations && it >= max_iterations) {
            break;
        }
    }
    if (iterations) {
        *iterations = it;
    }
    free(lla);
    free(llb);
    free(ca);
    free(cb);
    free(pa);
    return rc;
}
