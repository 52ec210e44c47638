/*
 * ndb_oracle.c -- CPU oracle (TEST INFRASTRUCTURE ONLY; see ndb_oracle.h).
 *
 * Every function cites the reference range (relative to /root/reference/) it
 * restates.  The restatement works on flat in-memory arrays instead of
 * PostgreSQL buffers; arithmetic order, operand types, comparison operators
 * and loop bounds follow the cited code literally.
 */
#include "ndb_oracle.h"

#include <float.h>
#include <math.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

/* ======================================================================== */
/* Operator arithmetic: NeuronDB/src/vector/vector_distance.c               */
/* ======================================================================== */

/* check_dimensions NaN/Inf scan, vector_distance.c:55-73 */
int orc_check_vector(const float *v, int dim)
{
    for (int i = 0; i < dim; i++)
        if (isnan(v[i]) || isinf(v[i]))
            return i;
    return -1;
}

/* l2_distance, vector_distance.c:93-122: fp64 Kahan sum of squared diffs, sqrt, cast */
float orc_l2_distance(const float *a, const float *b, int dim)
{
    double c = 0.0, sum = 0.0;
    for (int i = 0; i < dim; i++) {
        double diff = (double) a[i] - (double) b[i];
        double y = (diff * diff) - c;
        double t = sum + y;
        c = (t - sum) - y;
        sum = t;
    }
    return (float) sqrt(sum);
}

/* inner_product_distance, vector_distance.c:145-157: returns (float)(-sum) */
float orc_inner_product_distance(const float *a, const float *b, int dim)
{
    double sum = 0.0;
    for (int i = 0; i < dim; i++)
        sum += (double) a[i] * (double) b[i];
    return (float) (-sum);
}

/* inner_product_simd scalar fall-through, vector_distance_simd.c:557:
 * `return -inner_product_distance(a, b)` -> +dot. */
float orc_inner_product_op(const float *a, const float *b, int dim)
{
    return -orc_inner_product_distance(a, b, dim);
}

/* cosine_distance, vector_distance.c:180-213 */
float orc_cosine_distance(const float *a, const float *b, int dim)
{
    double dot = 0.0, na = 0.0, nb = 0.0;
    for (int i = 0; i < dim; i++) {
        double va = (double) a[i], vb = (double) b[i];
        dot += va * vb;
        na += va * va;
        nb += vb * vb;
    }
    if (na == 0.0 || nb == 0.0)
        return 1.0;
    return (float) (1.0 - (dot / (sqrt(na) * sqrt(nb))));
}

/* ======================================================================== */
/* AVX bodies restated lane by lane: vector_distance_simd.c:85-137,159-392  */
/* ======================================================================== */

/* horizontal_sum_avx2 (:85-101): (v[0..3]+v[4..7]) -> s; s[0]+s[1], s[2]+s[3]; then add.
 * horizontal_sum_avx512 (:120-137): first folds 16 -> 8 by lo+hi, then the same tree. */
static float orc_hsum(const float *v, int lanes)
{
    float t8[8];
    if (lanes == 16) {
        for (int i = 0; i < 8; i++) t8[i] = v[i] + v[i + 8];
    } else {
        for (int i = 0; i < 8; i++) t8[i] = v[i];
    }
    float s0 = t8[0] + t8[4], s1 = t8[1] + t8[5], s2 = t8[2] + t8[6], s3 = t8[3] + t8[7];
    /* movehdup: shuf = (s1,s1,s3,s3); sums = s + shuf -> (s0+s1, ., s2+s3, .) */
    float p0 = s0 + s1, p2 = s2 + s3;
    /* movehl + add_ss: p0 + p2 */
    return p0 + p2;
}

/* l2_distance_avx2 / _avx512, :159-185 / :188-217 (sub, mul, add -- no fmadd) */
float orc_l2_avx(const float *a, const float *b, int dim, int lanes)
{
    float acc[16] = {0};
    int simd_end = (dim / lanes) * lanes, i;
    for (i = 0; i < simd_end; i += lanes)
        for (int l = 0; l < lanes; l++) {
            float diff = a[i + l] - b[i + l];
            float sq = diff * diff;
            acc[l] = acc[l] + sq;
        }
    float sum = orc_hsum(acc, lanes);
    for (i = simd_end; i < dim; i++) {
        float diff = a[i] - b[i];
        sum += diff * diff;
    }
    return sqrtf(sum);
}

/* inner_product_avx2 / _avx512, :233-258 / :261-291 (mul, add); returns +sum */
float orc_ip_avx(const float *a, const float *b, int dim, int lanes)
{
    float acc[16] = {0};
    int simd_end = (dim / lanes) * lanes, i;
    for (i = 0; i < simd_end; i += lanes)
        for (int l = 0; l < lanes; l++) {
            float prod = a[i + l] * b[i + l];
            acc[l] = acc[l] + prod;
        }
    float sum = orc_hsum(acc, lanes);
    for (i = simd_end; i < dim; i++)
        sum += a[i] * b[i];
    return sum;
}

/* cosine_distance_avx2 / _avx512, :300-345 / :348-392 (fmadd in the vector body,
 * plain mul+add in the scalar tail; the AVX build implies -mfma is NOT required for
 * the tail because the reference enables only -mavx2; _mm256_fmadd_ps is explicit). */
float orc_cosine_avx(const float *a, const float *b, int dim, int lanes)
{
    float d[16] = {0}, na[16] = {0}, nb[16] = {0};
    int simd_end = (dim / lanes) * lanes, i;
    for (i = 0; i < simd_end; i += lanes)
        for (int l = 0; l < lanes; l++) {
            float va = a[i + l], vb = b[i + l];
            d[l] = fmaf(va, vb, d[l]);
            na[l] = fmaf(va, va, na[l]);
            nb[l] = fmaf(vb, vb, nb[l]);
        }
    float dot = orc_hsum(d, lanes), norm_a = orc_hsum(na, lanes), norm_b = orc_hsum(nb, lanes);
    for (i = simd_end; i < dim; i++) {
        float va = a[i], vb = b[i];
        /* the AVX2 build needs -mfma for _mm256_fmadd_ps, and gcc (-ffp-contract=fast, its
         * default) then contracts this scalar tail too: pinned against gcc 13 -O2 -mavx2 -mfma */
        dot = fmaf(va, vb, dot);
        norm_a = fmaf(va, va, norm_a);
        norm_b = fmaf(vb, vb, norm_b);
    }
    if (norm_a == 0.0f || norm_b == 0.0f)
        return 1.0f;
    float similarity = dot / (sqrtf(norm_a) * sqrtf(norm_b));
    return 1.0f - similarity;
}

/* ======================================================================== */
/* Index-AM arithmetic                                                      */
/* ======================================================================== */

/* ivfComputeDistance, ivf_am.c:1550-1592 (f32 sequential; default -> L2) */
float orc_ivf_distance(const float *v1, const float *v2, int dim, int strategy)
{
    float sum = 0.0f, dot = 0.0f, n1 = 0.0f, n2 = 0.0f;
    if (strategy == 2) {
        for (int i = 0; i < dim; i++) {
            dot += v1[i] * v2[i];
            n1 += v1[i] * v1[i];
            n2 += v2[i] * v2[i];
        }
        n1 = sqrtf(n1);
        n2 = sqrtf(n2);
        if (n1 == 0.0f || n2 == 0.0f)
            return 1.0f;
        return 1.0f - (dot / (n1 * n2));
    }
    if (strategy == ORC_IP) {
        /* NOT in the reference (ivfComputeDistance has no IP case and falls to
         * L2).  Added for BASELINE config 4, flagged: -dot in f32 sequential, the sign
         * convention of hnsw_am.c:1334-1337. */
        for (int i = 0; i < dim; i++)
            dot += v1[i] * v2[i];
        return -dot;
    }
    for (int i = 0; i < dim; i++) {
        float diff = v1[i] - v2[i];
        sum += diff * diff;
    }
    return sqrtf(sum);
}

/* hnswComputeDistance, hnsw_am.c:1301-1345 (f32 op, f64 accumulate) */
float orc_hnsw_distance(const float *v1, const float *v2, int dim, int strategy)
{
    double sum = 0.0, dot = 0.0, n1 = 0.0, n2 = 0.0;
    switch (strategy) {
    case 2:
        for (int i = 0; i < dim; i++) {
            dot += v1[i] * v2[i];
            n1 += v1[i] * v1[i];
            n2 += v2[i] * v2[i];
        }
        n1 = sqrt(n1);
        n2 = sqrt(n2);
        if (n1 == 0.0 || n2 == 0.0)
            return 2.0f;
        return (float) (1.0f - (dot / (n1 * n2)));
    case 3:
        for (int i = 0; i < dim; i++)
            dot += v1[i] * v2[i];
        return (float) (-dot);
    default: /* 1 = L2 (other strategies ereport(ERROR) in the reference) */
        for (int i = 0; i < dim; i++) {
            double d = v1[i] - v2[i];
            sum += d * d;
        }
        return (float) sqrt(sum);
    }
}

/* vector_distance_l2 (squared, no sqrt), ivf_am.c:2255-2269 */
float orc_kmeans_l2sq(const float *v1, const float *v2, int dim)
{
    float sum = 0.0;
    for (int i = 0; i < dim; i++) {
        float diff = v1[i] - v2[i];
        sum += diff * diff;
    }
    return sum;
}

float orc_distance(const float *a, const float *b, int dim, int metric, int arith)
{
    switch (arith) {
    case ORC_ARITH_OP_F64:
        return metric == ORC_L2 ? orc_l2_distance(a, b, dim)
             : metric == ORC_COSINE ? orc_cosine_distance(a, b, dim)
             : orc_inner_product_op(a, b, dim);
    case ORC_ARITH_AVX2:
    case ORC_ARITH_AVX512: {
        /* dispatchers vector_distance_simd.c:467-509,516-558,571-613: the SIMD body only
         * when dim >= lanes, else the scalar functions */
        int lanes = arith == ORC_ARITH_AVX2 ? 8 : 16;
        if (dim < lanes) {
            if (arith == ORC_ARITH_AVX512 && dim >= 8)
                lanes = 8;      /* an AVX-512 build also defines __AVX2__ ... but caps == AVX512,
                                 * so the AVX2 branch is skipped (:486-494): falls to scalar */
            return orc_distance(a, b, dim, metric, ORC_ARITH_OP_F64);
        }
        return metric == ORC_L2 ? orc_l2_avx(a, b, dim, lanes)
             : metric == ORC_COSINE ? orc_cosine_avx(a, b, dim, lanes)
             : orc_ip_avx(a, b, dim, lanes);
    }
    case ORC_ARITH_HNSW:
        return orc_hnsw_distance(a, b, dim, metric);
    default:
        return orc_ivf_distance(a, b, dim, metric);
    }
}

void orc_distance_pairs(const float *A, const float *B, float *out, int64_t n, int dim,
                        int metric, int arith)
{
    for (int64_t i = 0; i < n; i++)
        out[i] = orc_distance(A + i * dim, B + i * dim, dim, metric, arith);
}

/* ======================================================================== */
/* Exact kNN: SeqScan + top-N sort.  Ties by (dist ASC, id ASC), the          */
/* reference's only explicit rule (src/util/distributed.c:425-438).           */
/* ======================================================================== */

typedef struct { float d; int64_t id; } OrcCand;

static inline int cand_less(float d1, int64_t i1, float d2, int64_t i2)
{
    return d1 < d2 || (d1 == d2 && i1 < i2);
}

/* keep the k smallest in a sorted array (insertion) */
static inline void topk_push(OrcCand *best, int *cnt, int k, float d, int64_t id)
{
    int n = *cnt;
    if (n == k && !cand_less(d, id, best[k - 1].d, best[k - 1].id))
        return;
    int pos = n < k ? n : k - 1;
    while (pos > 0 && cand_less(d, id, best[pos - 1].d, best[pos - 1].id)) {
        best[pos] = best[pos - 1];
        pos--;
    }
    best[pos].d = d;
    best[pos].id = id;
    if (n < k) *cnt = n + 1;
}

void orc_knn_exact(const float *X, const int64_t *ids, int64_t n, int dim,
                   const float *Q, int nq, int k, int metric, int arith,
                   float *out_dist, int64_t *out_ids, int nthreads)
{
#ifdef _OPENMP
#pragma omp parallel for schedule(dynamic, 1) num_threads(nthreads > 0 ? nthreads : 1)
#endif
    for (int q = 0; q < nq; q++) {
        OrcCand *best = (OrcCand *) malloc(sizeof(OrcCand) * (size_t) k);
        int cnt = 0;
        const float *qv = Q + (size_t) q * dim;
        for (int64_t i = 0; i < n; i++) {
            /* operator argument order: row <-> query (a = row, b = query) */
            float d = orc_distance(X + (size_t) i * dim, qv, dim, metric, arith);
            topk_push(best, &cnt, k, d, ids ? ids[i] : i);
        }
        for (int j = 0; j < k; j++) {
            out_dist[(size_t) q * k + j] = j < cnt ? best[j].d : INFINITY;
            out_ids[(size_t) q * k + j] = j < cnt ? best[j].id : -1;
        }
        free(best);
    }
}

/* ======================================================================== */
/* IVF k-means: ivf_am.c:2070-2294                                           */
/* ======================================================================== */

/* ivfbuild: maxSamples = Min(10000, nlists*100), ivf_am.c:580; the build callback keeps
 * the first maxSamples heap tuples (:485-495) */
int orc_ivf_train_samples(int64_t nrows, int nlists)
{
    int64_t m = (int64_t) nlists * 100;
    if (m > 10000) m = 10000;
    return (int) (nrows < m ? nrows : m);
}

/* find_nearest_centroid, ivf_am.c:2274-2294: squared L2, strict <, best starts at 0 */
static int orc_find_nearest_centroid(const float *v, const float *C, int k, int dim)
{
    int best = 0;
    float bestDist = FLT_MAX;
    for (int c = 0; c < k; c++) {
        float dist = orc_kmeans_l2sq(v, C + (size_t) c * dim, dim);
        if (dist < bestDist) {
            bestDist = dist;
            best = c;
        }
    }
    return best;
}

/* kmeans_assign, ivf_am.c:2164-2177 */
void orc_kmeans_assign(const float *X, int64_t n, int dim, const float *C, int k,
                       int *assign, int nthreads)
{
#ifdef _OPENMP
#pragma omp parallel for schedule(static) num_threads(nthreads > 0 ? nthreads : 1)
#endif
    for (int64_t i = 0; i < n; i++)
        assign[i] = orc_find_nearest_centroid(X + (size_t) i * dim, C, k, dim);
}

/* kmeans_update_centroids, ivf_am.c:2182-2213: zero, f32 sum in sample order, divide
 * by the int count; empty clusters stay at zero.  counts[] is recomputed here the way
 * kmeans_assign does (:2169-2176). */
void orc_kmeans_update(const float *X, const int *assign, int64_t n, int dim, int k,
                       float *C, int *counts)
{
    memset(counts, 0, sizeof(int) * (size_t) k);
    for (int64_t i = 0; i < n; i++)
        counts[assign[i]]++;
    for (size_t j = 0; j < (size_t) k * dim; j++)
        C[j] = 0.0;
    for (int64_t i = 0; i < n; i++) {
        float *c = C + (size_t) assign[i] * dim;
        const float *x = X + (size_t) i * dim;
        for (int j = 0; j < dim; j++)
            c[j] += x[j];
    }
    for (int c = 0; c < k; c++)
        if (counts[c] > 0)
            for (int j = 0; j < dim; j++)
                C[(size_t) c * dim + j] /= counts[c];
}

/* kmeans_init + kmeans_run, ivf_am.c:2070-2112, 2117-2159.
 * Returns the number of Lloyd iterations executed (iter+1 at the break, else max_iter). */
int orc_kmeans_train(const float *X, int n, int dim, int k, int max_iter, float threshold,
                     float *C, int *assign, int *counts, float *cost_out)
{
    /* kmeans_init: centroid i := sample i for i < n; palloc'd (nalloc zeroes) otherwise */
    for (int i = 0; i < k; i++)
        for (int j = 0; j < dim; j++)
            C[(size_t) i * dim + j] = i < n ? X[(size_t) i * dim + j] : 0.0f;

    float prevCost = FLT_MAX, cost = 0.0f;
    int iter;
    for (iter = 0; iter < max_iter; iter++) {
        orc_kmeans_assign(X, n, dim, C, k, assign, 1);
        orc_kmeans_update(X, assign, n, dim, k, C, counts);
        /* kmeans_compute_cost, :2218-2233: f32 running sum of squared L2 in sample order */
        cost = 0.0;
        for (int i = 0; i < n; i++)
            cost += orc_kmeans_l2sq(X + (size_t) i * dim, C + (size_t) assign[i] * dim, dim);
        if (fabs(prevCost - cost) < threshold) {
            iter++;
            break;
        }
        prevCost = cost;
    }
    if (cost_out) *cost_out = cost;
    return iter;
}

/* ======================================================================== */
/* ivfinsert assignment, ivf_am.c:906-935: sqrtf(sum f32 diff^2), strict <   */
/* ======================================================================== */
void orc_ivf_assign(const float *X, int64_t n, int dim, const float *C, int nlists,
                    int *out_list, int nthreads)
{
#ifdef _OPENMP
#pragma omp parallel for schedule(static) num_threads(nthreads > 0 ? nthreads : 1)
#endif
    for (int64_t r = 0; r < n; r++) {
        const float *v = X + (size_t) r * dim;
        int best = 0;           /* min_idx = 0, ivf_am.c:812 */
        float minDist = FLT_MAX;
        for (int i = 0; i < nlists; i++) {
            const float *c = C + (size_t) i * dim;
            float dist = 0.0f;
            for (int j = 0; j < dim; j++) {
                float diff = v[j] - c[j];
                dist += diff * diff;
            }
            dist = sqrtf(dist);
            if (dist < minDist) {
                minDist = dist;
                best = i;
            }
        }
        out_list[r] = best;
    }
}

/* ======================================================================== */
/* IVF search                                                                */
/* ======================================================================== */

/* ivfSelectClusters, ivf_am.c:1597-1717: L2 (strategy 1) to every centroid, then nprobe
 * passes of "smallest not yet selected", strict <, FLT_MAX start (so a centroid at
 * distance >= FLT_MAX is never selected -> -1). */
void orc_ivf_select_clusters(const float *q, int dim, const float *C, int nlists, int nprobe,
                             int *selected)
{
    if (nprobe > nlists) nprobe = nlists;
    float *cd = (float *) malloc(sizeof(float) * (size_t) nlists);
    for (int i = 0; i < nlists; i++)
        cd[i] = orc_ivf_distance(q, C + (size_t) i * dim, dim, 1);
    for (int i = 0; i < nprobe; i++) {
        int bestIdx = -1;
        float bestDist = FLT_MAX;
        for (int j = 0; j < nlists; j++) {
            int already = 0;
            for (int s = 0; s < i; s++)
                if (selected[s] == j) { already = 1; break; }
            if (!already && cd[j] < bestDist) {
                bestDist = cd[j];
                bestIdx = j;
            }
        }
        selected[i] = bestIdx;
    }
    free(cd);
}

/* ivfCollectCandidates, ivf_am.c:1722-1909 */
int orc_ivf_search_one(const float *X, const int64_t *ids, int dim,
                       const float *C, int nlists, const int64_t *list_off, const int64_t *list_rows,
                       const float *q, int nprobe, int k, int strategy, int literal,
                       float *out_dist, int64_t *out_ids)
{
    if (nprobe > nlists) nprobe = nlists;
    int *sel = (int *) malloc(sizeof(int) * (size_t) (nprobe > 0 ? nprobe : 1));
    orc_ivf_select_clusters(q, dim, C, nlists, nprobe, sel);
    int result = 0;

    if (literal) {
        int maxCand = k * 10;                                   /* :1743 */
        float *cdist = (float *) malloc(sizeof(float) * (size_t) maxCand);
        int64_t *cid = (int64_t *) malloc(sizeof(int64_t) * (size_t) maxCand);
        int cc = 0;
        for (int i = 0; i < nprobe && cc < maxCand; i++) {      /* :1764 */
            int cl = sel[i];
            if (cl < 0 || cl >= nlists) continue;
            for (int64_t p = list_off[cl]; p < list_off[cl + 1] && cc < maxCand; p++) { /* :1811 */
                int64_t row = list_rows[p];
                cdist[cc] = orc_ivf_distance(q, X + (size_t) row * dim, dim, strategy);
                cid[cc] = ids ? ids[row] : row;
                cc++;
            }
        }
        if (cc > 0) {
            /* selection sort over an index array with swaps, strict <, :1861-1881 */
            int actualK = k < cc ? k : cc;
            int *idx = (int *) malloc(sizeof(int) * (size_t) cc);
            for (int i = 0; i < cc; i++) idx[i] = i;
            for (int i = 0; i < actualK; i++) {
                int bestIdx = i;
                float bestDist = cdist[idx[i]];
                for (int j = i + 1; j < cc; j++)
                    if (cdist[idx[j]] < bestDist) {
                        bestDist = cdist[idx[j]];
                        bestIdx = j;
                    }
                if (bestIdx != i) {
                    int t = idx[i]; idx[i] = idx[bestIdx]; idx[bestIdx] = t;
                }
            }
            for (int i = 0; i < actualK; i++) {
                out_dist[i] = cdist[idx[i]];
                out_ids[i] = cid[idx[i]];
            }
            result = actualK;
            free(idx);
        }
        free(cdist);
        free(cid);
    } else {
        OrcCand *best = (OrcCand *) malloc(sizeof(OrcCand) * (size_t) k);
        int cnt = 0;
        for (int i = 0; i < nprobe; i++) {
            int cl = sel[i];
            if (cl < 0 || cl >= nlists) continue;
            for (int64_t p = list_off[cl]; p < list_off[cl + 1]; p++) {
                int64_t row = list_rows[p];
                float d = orc_ivf_distance(q, X + (size_t) row * dim, dim, strategy);
                topk_push(best, &cnt, k, d, ids ? ids[row] : row);
            }
        }
        for (int i = 0; i < cnt; i++) {
            out_dist[i] = best[i].d;
            out_ids[i] = best[i].id;
        }
        result = cnt;
        free(best);
    }
    for (int i = result; i < k; i++) {
        out_dist[i] = INFINITY;
        out_ids[i] = -1;
    }
    free(sel);
    return result;
}

void orc_ivf_search(const float *X, const int64_t *ids, int dim,
                    const float *C, int nlists, const int64_t *list_off, const int64_t *list_rows,
                    const float *Q, int nq, int nprobe, int k, int strategy, int literal,
                    float *out_dist, int64_t *out_ids, int *out_count, int nthreads)
{
#ifdef _OPENMP
#pragma omp parallel for schedule(dynamic, 4) num_threads(nthreads > 0 ? nthreads : 1)
#endif
    for (int q = 0; q < nq; q++) {
        int c = orc_ivf_search_one(X, ids, dim, C, nlists, list_off, list_rows,
                                   Q + (size_t) q * dim, nprobe, k, strategy, literal,
                                   out_dist + (size_t) q * k, out_ids + (size_t) q * k);
        if (out_count) out_count[q] = c;
    }
}

/* ======================================================================== */
/* recall@k, src/ml/ml_recall_metrics.c:65-126 (mean over queries)           */
/* ======================================================================== */
double orc_recall_at_k(const int64_t *found, const int64_t *truth, int nq, int k)
{
    double total = 0.0;
    for (int q = 0; q < nq; q++) {
        int hit = 0;
        for (int i = 0; i < k; i++)
            for (int j = 0; j < k; j++)
                if (found[(size_t) q * k + j] == truth[(size_t) q * k + i]) { hit++; break; }
        total += (double) hit / k;
    }
    return nq > 0 ? total / nq : 0.0;
}

/* merge_distributed_results ordering, src/util/distributed.c:425-438: (dist, id) */
void orc_merge_topk(const float *dist, const int64_t *ids, int nshards, int nq, int k,
                    float *out_dist, int64_t *out_ids)
{
    OrcCand *best = (OrcCand *) malloc(sizeof(OrcCand) * (size_t) k);
    for (int q = 0; q < nq; q++) {
        int cnt = 0;
        for (int s = 0; s < nshards; s++)
            for (int j = 0; j < k; j++) {
                size_t o = ((size_t) s * nq + q) * k + j;
                if (ids[o] < 0) continue;
                topk_push(best, &cnt, k, dist[o], ids[o]);
            }
        for (int j = 0; j < k; j++) {
            out_dist[(size_t) q * k + j] = j < cnt ? best[j].d : INFINITY;
            out_ids[(size_t) q * k + j] = j < cnt ? best[j].id : -1;
        }
    }
    free(best);
}

/* ---- index key extraction -------------------------------------------------------------- */
/* fp16_to_float, src/types/quantization.c:171-215: sign / 5-bit exponent / 10-bit mantissa;
 * exponent 31 -> Inf/NaN with the mantissa moved up; normal numbers re-biased by 127 - 15 (exact,
 * = IEEE).  QUIRK (kept, the reference's own result is the contract): a binary16 SUBNORMAL is
 * renormalised with the float exponent 127 - 15 - (10 - e), e = 1 - shifts, i.e. 103 - shifts,
 * where IEEE needs 113 - shifts: subnormal halves come out 2^-10 times their IEEE value. */
float orc_fp16_to_float(uint16_t h)
{
    uint32_t sign = ((uint32_t) h & 0x8000u) << 16;
    uint32_t e = ((uint32_t) h & 0x7c00u) >> 10;
    uint32_t man = (uint32_t) h & 0x03ffu;
    uint32_t f;
    if (e == 0) {
        if (man == 0) f = sign;
        else {
            /* shift the leading one up to bit 10, counting the shifts (:188-196) */
            int shifts = 0;
            while ((man & 0x0400u) == 0) { man <<= 1; shifts++; }
            man &= 0x03ffu;
            f = sign | ((uint32_t) (127 - 15 - (10 - (1 - shifts))) << 23) | (man << 13);
        }
    } else if (e == 0x1f) {
        f = sign | 0x7f800000u | (man << 13);
    } else {
        f = sign | ((e + 127 - 15) << 23) | (man << 13);
    }
    float r;
    memcpy(&r, &f, 4);
    return r;
}

/* halfvec branch of hnswExtractVectorData (hnsw_am.c:1435-1450) / ivfExtractVectorData (ivf_am.c:165-173) */
void orc_keys_from_halfvec(const uint16_t *h, int64_t n, int dim, float *rows)
{
    for (int64_t i = 0; i < n * dim; i++) rows[i] = orc_fp16_to_float(h[i]);
}

/* bit branch (hnsw_am.c:1480-1507, ivf_am.c:190-206): MSB-first inside each byte, 1 -> +1, 0 -> -1 */
void orc_keys_from_bits(const uint8_t *bits, int64_t n, int nbits, float *rows)
{
    const int row_bytes = (nbits + 7) / 8;
    for (int64_t r = 0; r < n; r++)
        for (int i = 0; i < nbits; i++) {
            int byte_idx = i / 8, bit_idx = i % 8;
            int v = (bits[r * row_bytes + byte_idx] >> (8 - 1 - bit_idx)) & 1;
            rows[r * nbits + i] = v ? 1.0f : -1.0f;
        }
}

/* sparsevec branch (hnsw_am.c:1451-1479, ivf_am.c:174-189): memset 0, then entries in order */
void orc_keys_from_sparse(const int64_t *indptr, const int32_t *indices, const float *values,
                          int64_t n, int total_dim, float *rows)
{
    for (int64_t r = 0; r < n; r++) {
        float *row = rows + r * total_dim;
        memset(row, 0, sizeof(float) * (size_t) total_dim);
        for (int64_t e = indptr[r]; e < indptr[r + 1]; e++)
            if (indices[e] >= 0 && indices[e] < total_dim) row[indices[e]] = values[e];
    }
}
