/*
 * pp_oracle.c -- CPU restatement of the two PyPairs hot paths.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing under pypairs_b200/ may import, link or
 * execute this file; it is the checker for tests/, __graft_entry__.smoke()
 * and the cpu_baseline / --impl reference legs of bench.py.
 *
 * Parity status: PINNED.  The restatement is checked (tests/test_oracle.py)
 * against
 *   - the reference's own known-answer vector `min_ref`
 *     (pypairs/tests/test_sandbag.py:26-59),
 *   - fixtures produced by importing the reference's Numba kernels in the
 *     build container (oracle/make_golden.py -> tests/golden/ npz files),
 *   - MT19937 / shuffle known answers of numba 0.65.0
 *     (numba/cpython/randomimpl.py:109-171, 454-521, 1929-1946).
 *
 * Every function cites the reference lines it follows.  Loop structure is kept
 * faithful to the reference (one scalar pass per gene pair; per-cell reseeded
 * MT19937) so that timing it with OpenMP over the same axis the reference
 * parallelises (prange over g1; gufunc over cells) is a fair "port" baseline.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#define ORC_API __attribute__((visibility("default")))

ORC_API int orc_max_threads(void) {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}

/* ------------------------------------------------------------------------
 * sandbag: check_pairs  (pypairs/tools/sandbag.py:160-199)
 *
 * x      : [n_cells, n_genes] row-major float64 (the `raw_data_in` argument;
 *          the reference transposes it to gene-major at :165, we index it)
 * cats   : class index per cell, sample order           (sandbag.py:120)
 * thr    : ceil(n_k * fraction) per class               (sandbag.py:212-217)
 * result : [n_genes, n_genes] int64, pre-filled with -1 (sandbag.py:166)
 * Ties add to neither counter (:179-182); `>=` thresholds (:185,:188); the
 * last class satisfying wins (:186,:189); asymmetric if/elif (:192-197).
 * ---------------------------------------------------------------------- */
ORC_API void orc_check_pairs(const double *x, const int64_t *cats, const int64_t *thr,
                             int64_t n_cells, int64_t n_genes, int64_t n_classes,
                             int64_t *result, int n_threads) {
    /* gene-major copy, exactly the reference's ascontiguousarray(X.T) (:165) */
    double *xt = (double *)malloc(sizeof(double) * (size_t)n_cells * (size_t)n_genes);
    for (int64_t i = 0; i < n_cells; ++i)
        for (int64_t g = 0; g < n_genes; ++g) xt[g * n_cells + i] = x[i * n_genes + g];
    for (int64_t i = 0; i < n_genes * n_genes; ++i) result[i] = -1;
#ifdef _OPENMP
    if (n_threads <= 0) n_threads = omp_get_max_threads();
#pragma omp parallel for schedule(dynamic, 1) num_threads(n_threads)
#endif
    for (int64_t g1 = 0; g1 < n_genes; ++g1) {
        double *num_pos = (double *)malloc(sizeof(double) * (size_t)n_classes);
        double *num_neg = (double *)malloc(sizeof(double) * (size_t)n_classes);
        const double *r1 = xt + g1 * n_cells;
        for (int64_t g2 = g1 + 1; g2 < n_genes; ++g2) {
            const double *r2 = xt + g2 * n_cells;
            int64_t up = 0, up_idx = -1, down = 0, down_idx = -1;
            for (int64_t k = 0; k < n_classes; ++k) num_pos[k] = num_neg[k] = 0.0; /* :175-176 */
            for (int64_t i = 0; i < n_cells; ++i) {                                 /* :178-182 */
                if (r1[i] > r2[i]) num_pos[cats[i]] += 1.0;
                else if (r1[i] < r2[i]) num_neg[cats[i]] += 1.0;
            }
            for (int64_t k = 0; k < n_classes; ++k) {                               /* :184-190 */
                if (num_pos[k] >= (double)thr[k]) { up += 1; up_idx = k; }
                if (num_neg[k] >= (double)thr[k]) { down += 1; down_idx = k; }
            }
            if (up == 1) {                                                          /* :192-197 */
                if (down == n_classes - 1) result[g1 * n_genes + g2] = up_idx;
            } else if (down == 1) {
                if (up == n_classes - 1) result[g2 * n_genes + g1] = down_idx;
            }
        }
        free(num_pos);
        free(num_neg);
    }
    free(xt);
}

/* Per-class counts of selected pairs (the num_pos/num_neg of sandbag.py:175-182),
 * used to pin the GPU kernel's counters, not only its decisions.
 * pairs: [n_pairs,2] (g1,g2) ; pos/neg: [n_pairs, n_classes] int64. */
ORC_API void orc_pair_counts(const double *x, const int64_t *cats, int64_t n_cells, int64_t n_genes,
                             int64_t n_classes, const int64_t *pairs, int64_t n_pairs, int64_t *pos,
                             int64_t *neg) {
    memset(pos, 0, sizeof(int64_t) * (size_t)(n_pairs * n_classes));
    memset(neg, 0, sizeof(int64_t) * (size_t)(n_pairs * n_classes));
    for (int64_t p = 0; p < n_pairs; ++p) {
        int64_t g1 = pairs[2 * p], g2 = pairs[2 * p + 1];
        for (int64_t i = 0; i < n_cells; ++i) {
            double a = x[i * n_genes + g1], b = x[i * n_genes + g2];
            if (a > b) pos[p * n_classes + cats[i]] += 1;
            else if (a < b) neg[p * n_classes + cats[i]] += 1;
        }
    }
}

/* ------------------------------------------------------------------------
 * MT19937 exactly as Numba drives it for np.random.seed / np.random.shuffle
 *   seeding     numba_rnd_init == classic init_genrand (multiplier 1812433253)
 *               reached from np.random.seed (randomimpl.py:207-220)
 *   tempering   randomimpl.py:109-131
 *   randint(n)  masked rejection on the LOW nbits of one 32-bit draw,
 *               nbits = bit_length(n-1)  (randomimpl.py:149-171, 454-521)
 *   shuffle     Fisher-Yates descending i = n-1..1, j = randint(i+1)
 *               (randomimpl.py:1929-1946)
 * ---------------------------------------------------------------------- */
#define MT_N 624
#define MT_M 397
typedef struct { uint32_t mt[MT_N]; int idx; } orc_mt_t;

static void mt_seed(orc_mt_t *s, uint32_t seed) {
    s->mt[0] = seed;
    for (int i = 1; i < MT_N; ++i)
        s->mt[i] = 1812433253u * (s->mt[i - 1] ^ (s->mt[i - 1] >> 30)) + (uint32_t)i;
    s->idx = MT_N;
}
static void mt_refill(orc_mt_t *s) {
    uint32_t *mt = s->mt, y;
    int kk;
    for (kk = 0; kk < MT_N - MT_M; ++kk) {
        y = (mt[kk] & 0x80000000u) | (mt[kk + 1] & 0x7fffffffu);
        mt[kk] = mt[kk + MT_M] ^ (y >> 1) ^ ((y & 1u) ? 0x9908b0dfu : 0u);
    }
    for (; kk < MT_N - 1; ++kk) {
        y = (mt[kk] & 0x80000000u) | (mt[kk + 1] & 0x7fffffffu);
        mt[kk] = mt[kk + (MT_M - MT_N)] ^ (y >> 1) ^ ((y & 1u) ? 0x9908b0dfu : 0u);
    }
    y = (mt[MT_N - 1] & 0x80000000u) | (mt[0] & 0x7fffffffu);
    mt[MT_N - 1] = mt[MT_M - 1] ^ (y >> 1) ^ ((y & 1u) ? 0x9908b0dfu : 0u);
    s->idx = 0;
}
static inline uint32_t mt_next(orc_mt_t *s) {
    if (s->idx >= MT_N) mt_refill(s);
    uint32_t y = s->mt[s->idx++];
    y ^= (y >> 11);
    y ^= (y << 7) & 0x9d2c5680u;
    y ^= (y << 15) & 0xefc60000u;
    y ^= (y >> 18);
    return y;
}
static inline int bit_length_u64(uint64_t v) { return v ? 64 - __builtin_clzll(v) : 0; }
/* randint(0, n) for 2 <= n <= 2^32 */
static inline int64_t mt_randint(orc_mt_t *s, int64_t n) {
    if (n == 1) return 0;
    int nbits = bit_length_u64((uint64_t)(n - 1));
    uint32_t mask = 0xffffffffu >> (32 - nbits);
    int64_t r;
    do { r = (int64_t)(mt_next(s) & mask); } while (r >= n);
    return r;
}

ORC_API void orc_mt_outputs(uint32_t seed, int64_t n, uint32_t *out) {
    orc_mt_t s;
    mt_seed(&s, seed);
    for (int64_t i = 0; i < n; ++i) out[i] = mt_next(&s);
}

/* Cumulative permutation table: seed once, shuffle arange(n) `iters` times,
 * row k = state after k+1 shuffles.  The k-th null vector of any cell is then
 * v_used[table[k][.]]. */
ORC_API void orc_perm_table(uint32_t seed, int64_t n, int64_t iters, int32_t *table) {
    orc_mt_t s;
    mt_seed(&s, seed);
    int32_t *a = (int32_t *)malloc(sizeof(int32_t) * (size_t)(n > 0 ? n : 1));
    for (int64_t i = 0; i < n; ++i) a[i] = (int32_t)i;
    for (int64_t k = 0; k < iters; ++k) {
        for (int64_t i = n - 1; i > 0; --i) {
            int64_t j = mt_randint(&s, i + 1);
            int32_t t = a[i]; a[i] = a[j]; a[j] = t;
        }
        memcpy(table + k * n, a, sizeof(int32_t) * (size_t)n);
    }
    free(a);
}

/* ------------------------------------------------------------------------
 * cyclone: get_proportion  (pypairs/tools/cyclone.py:201-220)
 * ---------------------------------------------------------------------- */
static inline double get_proportion(const double *sample, int64_t min_pairs, const int64_t *pairs,
                                    int64_t n_pairs, int64_t *hits_out, int64_t *total_out) {
    int64_t hits = 0, total = 0;
    for (int64_t i = 0; i < n_pairs; ++i) {
        double a = sample[pairs[2 * i]], b = sample[pairs[2 * i + 1]];
        if (a != b) {
            if (a > b) hits += 1;
            total += 1;
        }
    }
    if (hits_out) *hits_out = hits;
    if (total_out) *total_out = total;
    if (hits < min_pairs || total == 0) return NAN;
    return (double)hits / (double)total;
}

/* ------------------------------------------------------------------------
 * cyclone: get_sample_score_guv + get_phase_scores (cyclone.py:223-257)
 * matrix [n_cells, n_genes] f64; used bool[n_genes]; pairs [n_pairs,2] into the
 * compacted used-gene vector.  fix_seed != 0 reseeds per cell (:226, :196-199).
 * obs_hits/obs_total (nullable) export the observed counters of :228.
 * ---------------------------------------------------------------------- */
ORC_API void orc_phase_scores(const double *matrix, int64_t n_cells, int64_t n_genes,
                              const uint8_t *used, const int64_t *pairs, int64_t n_pairs,
                              int64_t iterations, int64_t min_iter, int64_t min_pairs,
                              int fix_seed, uint32_t seed, double *scores, int64_t *obs_hits,
                              int64_t *obs_total, int n_threads) {
    int64_t n_used = 0;
    for (int64_t g = 0; g < n_genes; ++g) n_used += used[g] ? 1 : 0;
#ifdef _OPENMP
    if (n_threads <= 0) n_threads = omp_get_max_threads();
#pragma omp parallel num_threads(n_threads)
#endif
    {
        double *v = (double *)malloc(sizeof(double) * (size_t)(n_used > 0 ? n_used : 1));
        orc_mt_t *rng = (orc_mt_t *)malloc(sizeof(orc_mt_t));
        mt_seed(rng, seed);
#ifdef _OPENMP
#pragma omp for schedule(dynamic, 4)
#endif
        for (int64_t c = 0; c < n_cells; ++c) {
            const double *row = matrix + c * n_genes;
            if (fix_seed) mt_seed(rng, seed);                       /* :226 */
            int64_t u = 0;
            for (int64_t g = 0; g < n_genes; ++g)                   /* :227 sample[used] */
                if (used[g]) v[u++] = row[g];
            int64_t h0, t0;
            double cur = get_proportion(v, min_pairs, pairs, n_pairs, &h0, &t0); /* :228 */
            if (obs_hits) obs_hits[c] = h0;
            if (obs_total) obs_total[c] = t0;
            double score = NAN;                                      /* :230-231 (no return) */
            int64_t below = 0, total = 0;
            for (int64_t k = 0; k < iterations; ++k) {              /* :235-241 */
                for (int64_t i = n_used - 1; i > 0; --i) {          /* np.random.shuffle */
                    int64_t j = mt_randint(rng, i + 1);
                    double t = v[i]; v[i] = v[j]; v[j] = t;
                }
                double nw = get_proportion(v, min_pairs, pairs, n_pairs, NULL, NULL);
                if (nw != -1.0) {                                    /* always true, NaN != -1 */
                    if (nw < cur) below += 1;
                    total += 1;
                }
            }
            if (total == 0) score = NAN;                             /* :243-249 */
            if (total >= min_iter) score = (double)below / (double)total;
            else score = NAN;
            scores[c] = score;
        }
        free(v);
        free(rng);
    }
}

/* ------------------------------------------------------------------------
 * Philox4x32-10 (Salmon et al. 2011) and the Philox-mode permutation table.
 * This mode is OUR definition (the reference has no counter-based generator);
 * the GPU library's table generator must match this restatement bit for bit.
 *   key     = (seed_lo, seed_hi)
 *   counter = (draw_block, iteration, class_index, 0x50414952 "PAIR")
 * Iteration k starts from the identity and does a descending Fisher-Yates;
 * step i (n-1..1) takes 32-bit word (n-1-i) of the iteration's stream (4 words
 * per counter block) and maps it to j = mulhi32(word, i+1) (Lemire
 * multiply-shift without rejection; bias < 2^-20 for n <= 4096).
 * ---------------------------------------------------------------------- */
static inline void philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t k0,
                                 uint32_t k1, uint32_t out[4]) {
    for (int r = 0; r < 10; ++r) {
        uint64_t p0 = (uint64_t)0xD2511F53u * c0, p1 = (uint64_t)0xCD9E8D57u * c2;
        uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ k0, n1 = (uint32_t)p1;
        uint32_t n2 = (uint32_t)(p0 >> 32) ^ c3 ^ k1, n3 = (uint32_t)p0;
        c0 = n0; c1 = n1; c2 = n2; c3 = n3;
        k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
    }
    out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}
ORC_API void orc_philox(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t k0, uint32_t k1,
                        uint32_t *out) {
    philox4x32_10(c0, c1, c2, c3, k0, k1, out);
}
ORC_API void orc_perm_table_philox(uint64_t seed, uint32_t class_index, int64_t n, int64_t iters,
                                   int32_t *table) {
    uint32_t k0 = (uint32_t)seed, k1 = (uint32_t)(seed >> 32);
    for (int64_t k = 0; k < iters; ++k) {
        int32_t *a = table + k * n;
        for (int64_t i = 0; i < n; ++i) a[i] = (int32_t)i;
        uint32_t w[4] = {0, 0, 0, 0};
        for (int64_t i = n - 1; i > 0; --i) {
            int64_t d = n - 1 - i;
            if ((d & 3) == 0) philox4x32_10((uint32_t)(d >> 2), (uint32_t)k, class_index, 0x50414952u, k0, k1, w);
            uint32_t j = (uint32_t)(((uint64_t)w[d & 3] * (uint64_t)(i + 1)) >> 32);
            int32_t t = a[i]; a[i] = a[j]; a[j] = t;
        }
    }
}

/* Score every cell against a given permutation table (either generator):
 * the table-replay form of cyclone.py:223-249 used to check the GPU path in
 * both modes.  vec_k[i] = v[table[k][i]]. */
ORC_API void orc_phase_scores_table(const double *matrix, int64_t n_cells, int64_t n_genes,
                                    const uint8_t *used, const int64_t *pairs, int64_t n_pairs,
                                    const int32_t *table, int64_t iterations, int64_t min_iter,
                                    int64_t min_pairs, double *scores, int n_threads) {
    int64_t n_used = 0;
    for (int64_t g = 0; g < n_genes; ++g) n_used += used[g] ? 1 : 0;
#ifdef _OPENMP
    if (n_threads <= 0) n_threads = omp_get_max_threads();
#pragma omp parallel num_threads(n_threads)
#endif
    {
        double *v = (double *)malloc(sizeof(double) * (size_t)(n_used > 0 ? n_used : 1));
        double *w = (double *)malloc(sizeof(double) * (size_t)(n_used > 0 ? n_used : 1));
#ifdef _OPENMP
#pragma omp for schedule(dynamic, 4)
#endif
        for (int64_t c = 0; c < n_cells; ++c) {
            const double *row = matrix + c * n_genes;
            int64_t u = 0;
            for (int64_t g = 0; g < n_genes; ++g)
                if (used[g]) v[u++] = row[g];
            double cur = get_proportion(v, min_pairs, pairs, n_pairs, NULL, NULL);
            int64_t below = 0, total = 0;
            for (int64_t k = 0; k < iterations; ++k) {
                const int32_t *t = table + k * n_used;
                for (int64_t i = 0; i < n_used; ++i) w[i] = v[t[i]];
                double nw = get_proportion(w, min_pairs, pairs, n_pairs, NULL, NULL);
                if (nw < cur) below += 1;
                total += 1;
            }
            scores[c] = (total >= min_iter && total > 0) ? (double)below / (double)total : NAN;
        }
        free(v);
        free(w);
    }
}
