/*
 * rank_oracle.c -- TEST INFRASTRUCTURE ONLY (the checker, never the product path).
 *
 * Plain-C restatement of the two CMC/mAP evaluators on AGRL's test-time path:
 *
 *   oracle_eval_market1501  follows  torchreid/metrics/rank_cylib/rank_cy.pyx:154-241
 *                           (eval_market1501_cy) incl. function_cumsum :245-249
 *   oracle_eval_mars        follows  torchreid/metrics/rank.py:160-177 (evaluate_mars)
 *                           and      torchreid/metrics/rank.py:180-212 (Compute_AP)
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference leg may
 * load this file's library.  Parity is PINNED: tests/test_oracle_rank.py checks it against the
 * reference's own compiled rank_cy (oracle/_ref) and against golden vectors produced by the
 * reference's Python evaluate_mars (tests/golden/).
 *
 * Determinism the reference lacks: numpy's default argsort is unstable; like the north star
 * ("breaks ties stably by gallery index") the restatement orders a row by (distance, index),
 * NaN last, -0.0 == +0.0 -- i.e. np.argsort(kind='stable').
 *
 * Arithmetic notes (what has to be reproduced bit for bit):
 *   market1501: every accumulator is a C float; the AP term is evaluated in double and the sum
 *               rounded back to float at each step (rank_cy.pyx:219-225 after Cython lowering);
 *               the `cmc` scratch row is NOT cleared between queries (:177, :208-214), so a query
 *               whose kept gallery is shorter than max_rank inherits a stale tail.
 *   mars:       Python float (double) scalars; CMC = np.mean(axis=0) (row-by-row adds of 0/1),
 *               mAP = np.mean(ap) which is numpy's pairwise summation (restated below).
 *
 * Build: make -C oracle port   (gcc -O2 -ffp-contract=off, no -ffast-math)
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <math.h>

typedef struct { float d; int64_t j; } entry_t;

/* total order used by a stable ascending argsort: value, then index; NaN sorts last */
static int entry_less(const entry_t *a, const entry_t *b)
{
    int an = isnan(a->d), bn = isnan(b->d);
    if (an || bn) {
        if (an != bn) return bn;          /* non-NaN < NaN */
        return a->j < b->j;
    }
    if (a->d < b->d) return 1;
    if (a->d > b->d) return 0;
    return a->j < b->j;                   /* ties (incl. -0.0 vs +0.0) by gallery index */
}

static int entry_cmp(const void *pa, const void *pb)
{
    const entry_t *a = (const entry_t *)pa, *b = (const entry_t *)pb;
    if (entry_less(a, b)) return -1;
    if (entry_less(b, a)) return 1;
    return 0;
}

static void sort_row(const float *row, int64_t n, entry_t *buf)
{
    for (int64_t j = 0; j < n; ++j) { buf[j].d = row[j]; buf[j].j = j; }
    qsort(buf, (size_t)n, sizeof(entry_t), entry_cmp);
}

/* numpy's pairwise float64 summation (numpy/_core/src/umath/loops_utils.h.src, DOUBLE_pairwise_sum):
 * <8 sequential; <=128: eight interleaved partial sums then a fixed tree; else split at n/2 rounded
 * down to a multiple of 8.  This is what np.mean / np.add.reduce run on a contiguous 1-D double array. */
double oracle_pairwise_sum_f64(const double *a, int64_t n)
{
    if (n < 8) {
        double res = 0.0;
        for (int64_t i = 0; i < n; ++i) res += a[i];
        return res;
    } else if (n <= 128) {
        double r[8];
        for (int k = 0; k < 8; ++k) r[k] = a[k];
        int64_t i;
        for (i = 8; i < n - (n % 8); i += 8)
            for (int k = 0; k < 8; ++k) r[k] += a[i + k];
        double res = ((r[0] + r[1]) + (r[2] + r[3])) + ((r[4] + r[5]) + (r[6] + r[7]));
        for (; i < n; ++i) res += a[i];
        return res;
    } else {
        int64_t n2 = n / 2;
        n2 -= n2 % 8;
        return oracle_pairwise_sum_f64(a, n2) + oracle_pairwise_sum_f64(a + n2, n - n2);
    }
}

/*
 * market1501 metric.  out_cmc has room for max_rank floats; the number actually written is
 * min(max_rank, num_g) and is returned through *out_rank_len.  Returns 0, or 1 when no query is
 * valid (the reference raises AssertionError there), or -1 on allocation failure.
 */
int oracle_eval_market1501(const float *distmat,
                           const int64_t *q_pids, const int64_t *g_pids,
                           const int64_t *q_camids, const int64_t *g_camids,
                           int64_t num_q, int64_t num_g, int64_t max_rank,
                           float *out_cmc, float *out_mAP, float *out_all_AP,
                           int64_t *out_rank_len, int64_t *out_num_valid)
{
    if (num_g < max_rank) max_rank = num_g;
    entry_t *buf     = (entry_t *)malloc(sizeof(entry_t) * (size_t)(num_g > 0 ? num_g : 1));
    float   *hits    = (float *)calloc((size_t)(num_g > 0 ? num_g : 1), sizeof(float));
    float   *scratch = (float *)calloc((size_t)(num_g > 0 ? num_g : 1), sizeof(float)); /* never re-zeroed */
    float   *cum     = (float *)calloc((size_t)(num_g > 0 ? num_g : 1), sizeof(float));
    float   *all_cmc = (float *)calloc((size_t)(num_q * max_rank > 0 ? num_q * max_rank : 1), sizeof(float));
    float   *all_ap  = (float *)calloc((size_t)(num_q > 0 ? num_q : 1), sizeof(float));
    if (!buf || !hits || !scratch || !cum || !all_cmc || !all_ap) return -1;

    float num_valid = 0.f;
    for (int64_t q = 0; q < num_q; ++q) {
        const int64_t pid = q_pids[q], cam = q_camids[q];
        sort_row(distmat + q * num_g, num_g, buf);

        /* drop gallery entries of the same identity seen by the same camera */
        int64_t kept = 0;
        int any_hit = 0;
        for (int64_t r = 0; r < num_g; ++r) {
            const int64_t j = buf[r].j;
            if (g_pids[j] != pid || g_camids[j] != cam) {
                const float m = (g_pids[j] == pid) ? 1.f : 0.f;
                hits[kept++] = m;
                if (m > 1e-31f) any_hit = 1;
            }
        }
        if (!any_hit) continue;                 /* identity absent from the gallery: query skipped */

        /* CMC row: running count clipped at one, written into the persistent scratch */
        if (kept > 0) {
            scratch[0] = hits[0];
            for (int64_t r = 1; r < kept; ++r) scratch[r] = hits[r] + scratch[r - 1];
            for (int64_t r = 0; r < kept; ++r) if (scratch[r] > 1.f) scratch[r] = 1.f;
        }
        for (int64_t r = 0; r < max_rank; ++r) all_cmc[q * max_rank + r] = scratch[r];
        num_valid = (float)((double)num_valid + 1.0);

        /* average precision, float accumulators, term in double */
        cum[0] = hits[0];
        for (int64_t r = 1; r < kept; ++r) cum[r] = hits[r] + cum[r - 1];
        float n_rel = 0.f, acc = 0.f;
        for (int64_t r = 0; r < kept; ++r) {
            acc   = (float)((double)acc + ((double)cum[r] / ((double)r + 1.0)) * (double)hits[r]);
            n_rel = n_rel + hits[r];
        }
        all_ap[q] = acc / n_rel;
    }

    int rc = 0;
    if (!(num_valid > 0.f)) {
        rc = 1;
    } else {
        for (int64_t r = 0; r < max_rank; ++r) {
            float s = 0.f;
            for (int64_t q = 0; q < num_q; ++q) s += all_cmc[q * max_rank + r];
            out_cmc[r] = s / num_valid;
        }
        float m = 0.f;
        for (int64_t q = 0; q < num_q; ++q) m += all_ap[q];
        *out_mAP = m / num_valid;
    }
    if (out_all_AP)    memcpy(out_all_AP, all_ap, sizeof(float) * (size_t)num_q);
    if (out_rank_len)  *out_rank_len = max_rank;
    if (out_num_valid) *out_num_valid = (int64_t)num_valid;
    free(buf); free(hits); free(scratch); free(cum); free(all_cmc); free(all_ap);
    return rc;
}

/*
 * MARS metric (what the reference's test() really calls).  Requires num_g >= max_rank (the
 * reference's row assignment fails to broadcast otherwise).  Returns 0; 2 when some query has no
 * cross-camera match (the reference raises ZeroDivisionError); -1 on allocation failure.
 */
int oracle_eval_mars(const float *distmat,
                     const int64_t *q_pids, const int64_t *g_pids,
                     const int64_t *q_camids, const int64_t *g_camids,
                     int64_t num_q, int64_t num_g, int64_t max_rank,
                     double *out_cmc, double *out_mAP, double *out_ap)
{
    if (num_g < max_rank) return 3;
    entry_t *buf  = (entry_t *)malloc(sizeof(entry_t) * (size_t)(num_g > 0 ? num_g : 1));
    double  *row  = (double *)malloc(sizeof(double) * (size_t)(max_rank > 0 ? max_rank : 1));
    double  *ap   = (double *)calloc((size_t)(num_q > 0 ? num_q : 1), sizeof(double));
    double  *csum = (double *)calloc((size_t)(max_rank > 0 ? max_rank : 1), sizeof(double));
    if (!buf || !row || !ap || !csum) return -1;

    int rc = 0;
    for (int64_t q = 0; q < num_q && rc == 0; ++q) {
        const int64_t pid = q_pids[q], cam = q_camids[q];
        int64_t ngood = 0;
        for (int64_t j = 0; j < num_g; ++j)
            if (g_pids[j] == pid && g_camids[j] != cam) ++ngood;
        sort_row(distmat + q * num_g, num_g, buf);

        for (int64_t r = 0; r < max_rank; ++r) row[r] = 0.0;
        double old_recall = 0.0, old_precision = 1.0, acc = 0.0;
        int64_t inter = 0, j_eff = 0, good_now = 0, njunk = 0;
        for (int64_t n = 0; n < max_rank; ++n) {
            const int64_t g = buf[n].j;
            const int is_good = (g_pids[g] == pid) && (g_camids[g] != cam);
            const int is_junk = (g_pids[g] == -1) || ((g_pids[g] == pid) && (g_camids[g] == cam));
            if (is_good) {
                for (int64_t r = n - njunk; r < max_rank; ++r) row[r] = 1.0;
                ++good_now;
            }
            if (is_junk) { ++njunk; continue; }
            if (is_good) ++inter;
            if (ngood == 0) { rc = 2; break; }              /* ZeroDivisionError in the reference */
            const double recall    = (double)inter / (double)ngood;
            const double precision = (double)inter / (double)(j_eff + 1);
            acc = acc + ((recall - old_recall) * (old_precision + precision)) / 2.0;
            old_recall = recall;
            old_precision = precision;
            ++j_eff;
            if (good_now == ngood) break;
        }
        ap[q] = acc;
        for (int64_t r = 0; r < max_rank; ++r) csum[r] += row[r];   /* np.mean(axis=0): row-by-row adds */
    }
    if (rc == 0) {
        for (int64_t r = 0; r < max_rank; ++r) out_cmc[r] = csum[r] / (double)num_q;
        *out_mAP = oracle_pairwise_sum_f64(ap, num_q) / (double)num_q;
    }
    if (out_ap) memcpy(out_ap, ap, sizeof(double) * (size_t)num_q);
    free(buf); free(row); free(ap); free(csum);
    return rc;
}
