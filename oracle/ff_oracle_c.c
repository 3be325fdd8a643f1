/* TEST INFRASTRUCTURE — plain-C restatement of the arithmetic of the Fast-Forward
 * re-ranking hot path.  Not product code: only tests/, __graft_entry__.smoke() and
 * bench.py's CPU-baseline leg load the library built from this file.
 *
 * The reference (mrjleo/fast-forward-indexes v0.8.0) is pure Python; the arithmetic on
 * this path is done by numpy / pandas:
 *   - `ff_score = np.sum(q_reps * d_reps, axis=1)`      src/fast_forward/index/base.py:303
 *   - `groupby(["id","q_no"]).aggregate(max|mean|first)` src/fast_forward/index/base.py:306-312
 *   - `alpha * score + (1 - alpha) * score_other`        src/fast_forward/ranking.py:319
 *   - sort by score DESC (stable) + head(k)              src/fast_forward/ranking.py:115-117,285-291
 * This file writes those out as explicit scalar loops so the summation trees are visible
 * (they are what the CUDA kernel reproduces lane by lane) and so mid-size cases can be
 * checked quickly.  tests/test_oracle.py pins it bit-for-bit against numpy/pandas and
 * against the golden vectors produced by the unmodified reference.
 *
 * Build: see oracle/Makefile (-O2 -ffp-contract=off: products and sums must round
 * separately, never fuse into FMA).
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define FFO_MODE_PASSAGE 1
#define FFO_MODE_MAXP 2
#define FFO_MODE_FIRSTP 3
#define FFO_MODE_AVEP 4

/* numpy's fp32 pairwise summation of the element-wise products q[i]*d[i] (N1):
 *   n < 8    : sequential from 0
 *   n <= 128 : r[j] = x[j] (j<8); r[j] += x[i+j] for i = 8,16,..; then
 *              ((r0+r1)+(r2+r3)) + ((r4+r5)+(r6+r7)); then the n%8 tail sequentially
 *   else     : split at n2 = n/2 - (n/2)%8 and add the two halves
 */
static float pairwise_dot(const float *q, const float *d, int64_t n) {
    if (n < 8) {
        float res = 0.f;
        for (int64_t i = 0; i < n; i++) res += q[i] * d[i];
        return res;
    }
    if (n <= 128) {
        float r[8];
        for (int j = 0; j < 8; j++) r[j] = q[j] * d[j];
        int64_t i;
        for (i = 8; i < n - (n % 8); i += 8)
            for (int j = 0; j < 8; j++) r[j] += q[i + j] * d[i + j];
        float res = ((r[0] + r[1]) + (r[2] + r[3])) + ((r[4] + r[5]) + (r[6] + r[7]));
        for (; i < n; i++) res += q[i] * d[i];
        return res;
    }
    int64_t n2 = n / 2;
    n2 -= n2 % 8;
    return pairwise_dot(q, d, n2) + pairwise_dot(q + n2, d + n2, n - n2);
}

/* One (query, passage) dot product exactly as `np.sum(q * d, axis=1)` yields it:
 * the reduction starts from the identity 0 and adds the pairwise result. */
float ffo_dot_f32(const float *q, const float *d, int64_t n) {
    return 0.f + pairwise_dot(q, d, n);
}

/* Scores for n pairs.  Unit u covers rows unit_rows[unit_off[u] .. unit_off[u+1]) of the
 * row-major fp32 matrix `vectors` (mode-resolved by the caller: PASSAGE/FIRSTP units have
 * one row).  MAXP = max, AVEP = fp32 Kahan sum in row order / fp32(count) (N2), else first. */
void ffo_score_pairs(const float *vectors, int64_t dim, const int64_t *unit_off,
                     const int64_t *unit_rows, const int64_t *pair_q, const int64_t *pair_unit,
                     int64_t n_pairs, const float *qvecs, int mode, float *out) {
#pragma omp parallel for schedule(dynamic, 256)
    for (int64_t p = 0; p < n_pairs; p++) {
        const float *q = qvecs + pair_q[p] * dim;
        int64_t b = unit_off[pair_unit[p]], e = unit_off[pair_unit[p] + 1];
        float acc = 0.f, comp = 0.f;
        for (int64_t k = b; k < e; k++) {
            float s = ffo_dot_f32(q, vectors + unit_rows[k] * dim, dim);
            if (mode == FFO_MODE_MAXP) {
                if (k == b || s > acc) acc = s; /* NaN scores are outside the contract */
            } else if (mode == FFO_MODE_AVEP) {
                float y = s - comp;
                float t = acc + y;
                comp = (t - acc) - y;
                if (comp != comp) comp = 0.f;
                acc = t;
            } else if (k == b) {
                acc = s;
            }
        }
        if (mode == FFO_MODE_AVEP) acc = acc / (float)(e - b);
        out[p] = acc;
    }
}

/* ranking.py:319 — fl32(fl32(alpha)*s) + fl32(fl32(1-alpha)*f), 1-alpha formed in double. */
void ffo_interpolate(const float *score_self, const float *score_other, int64_t n, double alpha,
                     float *out) {
    const float a = (float)alpha, b = (float)(1.0 - alpha);
    for (int64_t i = 0; i < n; i++) {
        float x = a * score_self[i];
        float y = b * score_other[i];
        out[i] = x + y;
    }
}

/* Per query block [q_off[q], q_off[q+1]): stable order by score DESC, keep k.
 * Outputs are [nq, k], padded with (-inf, -1). */
typedef struct { float s; int32_t pos; } ffo_item;

static int item_cmp(const void *a, const void *b) {
    const ffo_item *x = (const ffo_item *)a, *y = (const ffo_item *)b;
    if (x->s > y->s) return -1;
    if (x->s < y->s) return 1;
    return (x->pos > y->pos) - (x->pos < y->pos);
}

void ffo_topk(const int64_t *q_off, int64_t nq, const float *scores, int64_t k, float *out_s,
              int32_t *out_pos) {
#pragma omp parallel for schedule(dynamic, 4)
    for (int64_t q = 0; q < nq; q++) {
        int64_t b = q_off[q], c = q_off[q + 1] - b;
        ffo_item *items = (ffo_item *)malloc(sizeof(ffo_item) * (size_t)(c > 0 ? c : 1));
        for (int64_t i = 0; i < c; i++) { items[i].s = scores[b + i]; items[i].pos = (int32_t)i; }
        qsort(items, (size_t)c, sizeof(ffo_item), item_cmp);
        for (int64_t i = 0; i < k; i++) {
            out_s[q * k + i] = i < c ? items[i].s : -INFINITY;
            out_pos[q * k + i] = i < c ? items[i].pos : -1;
        }
        free(items);
    }
}

/* nanopq PQ.decode followed by the reference dot: used to check ADC at mid sizes.
 * codes [n_rows, M] u8, codewords [M, Ks, Ds] f32 -> out [n_rows, M*Ds]. */
void ffo_pq_decode(const uint8_t *codes, int64_t n_rows, int64_t M, int64_t Ks, int64_t Ds,
                   const float *codewords, float *out) {
#pragma omp parallel for schedule(static)
    for (int64_t r = 0; r < n_rows; r++)
        for (int64_t m = 0; m < M; m++)
            memcpy(out + (r * M + m) * Ds, codewords + (m * Ks + codes[r * M + m]) * Ds,
                   sizeof(float) * (size_t)Ds);
}
