/*
 * ndb_oracle_ml.c -- CPU oracle for the SQL functions that sit on the same kernels as the index path
 * (SURVEY.md 8f-3): knn_classify / knn_regress and cluster_kmeans.
 *
 * TEST INFRASTRUCTURE ONLY (see ndb_oracle.h).  Restated over in-memory arrays from
 *   NeuronDB/src/ml/ml_knn.c:63-90 (compare_samples, euclidean_distance), :264-333 (classify), :504-557 (regress)
 *   NeuronDB/src/ml/ml_kmeans.c:45-139 (kmeanspp_init), :146-303 (cluster_kmeans)
 *   NeuronDB/src/util/neurondb_simd_impl.c:36-104 (neurondb_l2_distance_squared, the build without -mavx2)
 *
 * Pinning: PINNED against the reference's own euclidean_distance / compare_samples / kmeanspp_init /
 * neurondb_l2_distance_squared and the Lloyd loop of cluster_kmeans, cut out of those files by
 * oracle/extract_ref_leafs.py (tests/test_oracle.py, golden tests/golden/ml_paths.npz).
 */
#include <float.h>
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#include "ndb_oracle.h"

/* ml_knn.c:76-90: the difference is a float operation, its square and the sum are double */
double orc_ml_euclidean(const float *a, const float *b, int dim)
{
    double sum = 0.0;
    for (int i = 0; i < dim; i++) {
        double diff = a[i] - b[i];
        sum += diff * diff;
    }
    return sqrt(sum);
}

typedef struct { double distance; double label; int row; } OrcKnnSample;

/* ml_knn.c:63-74 orders by distance alone; equal distances keep row order here (what glibc's merge-sorting qsort
 * does for the reference; the C standard leaves it open, so the tests stay away from ties at the k-th place) */
static int orc_knn_cmp(const void *a, const void *b)
{
    const OrcKnnSample *x = (const OrcKnnSample *) a, *y = (const OrcKnnSample *) b;
    if (x->distance < y->distance) return -1;
    if (x->distance > y->distance) return 1;
    return x->row < y->row ? -1 : (x->row > y->row);
}

/* one query: class (:326-333; labels outside {0,1} do not vote, class 1 needs a strict majority), mean of the k
 * labels in sorted order (:555-557), and the k nearest rows */
void orc_knn_ml(const float *X, const double *labels, int n, int dim, const float *q, int k, int *cls, double *mean,
                int *rows_out)
{
    OrcKnnSample *s = (OrcKnnSample *) malloc(sizeof(OrcKnnSample) * (size_t) n);
    double votes[2] = {0.0, 0.0}, prediction = 0.0;
    for (int i = 0; i < n; i++) {
        s[i].distance = orc_ml_euclidean(q, X + (size_t) i * dim, dim);
        s[i].label = labels[i];
        s[i].row = i;
    }
    qsort(s, (size_t) n, sizeof(OrcKnnSample), orc_knn_cmp);
    for (int i = 0; i < k && i < n; i++) {
        int c = (int) s[i].label;
        if (c >= 0 && c < 2) votes[c] += 1.0;
        prediction += s[i].label;
        if (rows_out) rows_out[i] = s[i].row;
    }
    *cls = votes[1] > votes[0] ? 1 : 0;
    *mean = prediction / k;
    free(s);
}

/* neurondb_simd_impl.c:94-101 (scalar tail = the whole loop without AVX2/NEON): all double */
double orc_l2_distance_squared(const float *a, const float *b, int n)
{
    double sum = 0.0;
    for (int i = 0; i < n; i++) {
        double diff = (double) a[i] - (double) b[i];
        sum += diff * diff;
    }
    return sum;
}

/* ml_kmeans.c:45-139.  draws[c] = the c-th value rand() returned (the function calls it once per seed, nothing
 * else in between); rand_max = RAND_MAX of that libc.  D^2 weights: float difference, FLOAT square, double sum. */
int orc_kmeanspp_init(const float *X, int nvec, int dim, int k, const int *draws, int rand_max, int *centroids)
{
    unsigned char *selected = (unsigned char *) calloc((size_t) nvec, 1);
    double *dist = (double *) calloc((size_t) nvec, sizeof(double));
    centroids[0] = draws[0] % nvec;
    selected[centroids[0]] = 1;
    for (int i = 0; i < nvec; i++) {
        double acc = 0.0;
        for (int d = 0; d < dim; d++) {
            float diff = X[(size_t) i * dim + d] - X[(size_t) centroids[0] * dim + d];
            acc += diff * diff;
        }
        dist[i] = acc;
    }
    for (int c = 1; c < k; c++) {
        int picked = -1;
        double sum = 0.0, r;
        for (int i = 0; i < nvec; i++)
            if (!selected[i]) sum += dist[i];
        r = ((double) draws[c] / (double) rand_max) * sum;
        for (int i = 0; i < nvec; i++) {
            if (selected[i]) continue;
            r -= dist[i];
            if (r <= 0) { picked = i; break; }
        }
        if (picked < 0)
            for (int i = 0; i < nvec; i++)
                if (!selected[i]) { picked = i; break; }
        if (picked < 0) { free(dist); free(selected); return -1; }
        centroids[c] = picked;
        selected[picked] = 1;
        for (int i = 0; i < nvec; i++) {
            double acc = 0.0;
            for (int d = 0; d < dim; d++)
                acc += (X[(size_t) i * dim + d] - X[(size_t) picked * dim + d]) * (X[(size_t) i * dim + d] - X[(size_t) picked * dim + d]);
            if (acc < dist[i]) dist[i] = acc;
        }
    }
    free(dist);
    free(selected);
    return 0;
}

/* ml_kmeans.c:146-303: seeds -> Lloyd until no assignment changes or max_iters (< 1 -> 100); assignment in double
 * (strict <, lowest index wins), update = float sums in row order / count, empty clusters end at zero.
 * labels are 1-based (:286).  Returns the number of iterations run, < 0 on the argument errors (:170-186). */
int orc_cluster_kmeans(const float *X, int nvec, int dim, int k, int max_iters, const int *draws, int rand_max,
                       int *labels, float *centers_out, int *seeds_out)
{
    if (k <= 1) return -1;                 /* "number of clusters must be at least 2" */
    if (max_iters < 1) max_iters = 100;
    if (nvec < k) return -2;               /* "not enough vectors for cluster count" */
    int *idx = (int *) malloc(sizeof(int) * (size_t) k);
    int *assign = (int *) malloc(sizeof(int) * (size_t) nvec);
    int *counts = (int *) malloc(sizeof(int) * (size_t) k);
    float *centers = (float *) malloc(sizeof(float) * (size_t) k * dim);
    if (orc_kmeanspp_init(X, nvec, dim, k, draws, rand_max, idx) != 0) return -3;
    for (int c = 0; c < k; c++) memcpy(centers + (size_t) c * dim, X + (size_t) idx[c] * dim, sizeof(float) * (size_t) dim);
    if (seeds_out) memcpy(seeds_out, idx, sizeof(int) * (size_t) k);
    for (int i = 0; i < nvec; i++) assign[i] = -1;
    int changed = 1, iter;
    for (iter = 0; iter < max_iters && changed; iter++) {
        changed = 0;
        for (int i = 0; i < nvec; i++) {
            int best = -1;
            double min_dist = DBL_MAX;
            for (int c = 0; c < k; c++) {
                double dist = orc_l2_distance_squared(X + (size_t) i * dim, centers + (size_t) c * dim, dim);
                if (dist < min_dist) { min_dist = dist; best = c; }
            }
            if (assign[i] != best) { assign[i] = best; changed = 1; }
        }
        memset(centers, 0, sizeof(float) * (size_t) k * dim);
        memset(counts, 0, sizeof(int) * (size_t) k);
        for (int i = 0; i < nvec; i++) {
            int c = assign[i];
            for (int d = 0; d < dim; d++) centers[(size_t) c * dim + d] += X[(size_t) i * dim + d];
            counts[c]++;
        }
        for (int c = 0; c < k; c++)
            if (counts[c] > 0)
                for (int d = 0; d < dim; d++) centers[(size_t) c * dim + d] /= counts[c];
    }
    for (int i = 0; i < nvec; i++) labels[i] = assign[i] + 1;
    if (centers_out) memcpy(centers_out, centers, sizeof(float) * (size_t) k * dim);
    free(idx); free(assign); free(counts); free(centers);
    return iter;
}

/* ---- product quantisation: NeuronDB/src/ml/ml_product_quantization.c ------------------------------
 * train_subspace_kmeans :80-190, train_pq_codebook :195-415, pq_encode_vector :421-536, pq_asymmetric_distance :1003-1110.
 * Codebook layout as the bytea carries it: float centroids[m][ksub][dsub].  PINNED against the reference's own
 * train_subspace_kmeans and the text of the encode / distance loops (oracle/extract_ref_leafs.py). */

/* :80-190 on one subspace (rows of dsub floats, contiguous): seeds = rows draws[c] % nvec (duplicates allowed),
 * assignments start at 0 (palloc0), the loop leaves before the update when nothing changed; empty clusters end at zero */
void orc_pq_train_subspace(const float *S, int nvec, int dsub, int k, const int *draws, float *centroids, int max_iters)
{
    int *assign = (int *) calloc((size_t) nvec, sizeof(int));
    int *counts = (int *) malloc(sizeof(int) * (size_t) k);
    for (int c = 0; c < k; c++) memcpy(centroids + (size_t) c * dsub, S + (size_t) (draws[c] % nvec) * dsub, sizeof(float) * (size_t) dsub);
    for (int iter = 0; iter < max_iters; iter++) {
        int changed = 0;
        for (int i = 0; i < nvec; i++) {
            double min_dist = DBL_MAX;
            int best = -1;
            for (int c = 0; c < k; c++) {
                double dist = 0.0;
                for (int d = 0; d < dsub; d++) {
                    double diff = (double) S[(size_t) i * dsub + d] - (double) centroids[(size_t) c * dsub + d];
                    dist += diff * diff;
                }
                if (dist < min_dist) { min_dist = dist; best = c; }
            }
            if (assign[i] != best) { assign[i] = best; changed = 1; }
        }
        if (!changed) break;
        memset(counts, 0, sizeof(int) * (size_t) k);
        memset(centroids, 0, sizeof(float) * (size_t) k * dsub);
        for (int i = 0; i < nvec; i++) {
            int c = assign[i];
            for (int d = 0; d < dsub; d++) centroids[(size_t) c * dsub + d] += S[(size_t) i * dsub + d];
            counts[c]++;
        }
        for (int c = 0; c < k; c++)
            if (counts[c] > 0)
                for (int d = 0; d < dsub; d++) centroids[(size_t) c * dsub + d] /= counts[c];
    }
    free(assign); free(counts);
}

/* train_pq_codebook's loop over the subspaces (:303-352): draws[sub*ksub + c], 100 iterations */
int orc_pq_train(const float *X, int nvec, int dim, int m, int ksub, const int *draws, int max_iters, float *codebooks)
{
    if (m < 1 || m > 128 || ksub < 2 || ksub > 65536 || dim <= 0 || dim % m != 0 || nvec < 1) return -1;
    int dsub = dim / m;
    float *S = (float *) malloc(sizeof(float) * (size_t) nvec * dsub);
    for (int sub = 0; sub < m; sub++) {
        for (int i = 0; i < nvec; i++) memcpy(S + (size_t) i * dsub, X + (size_t) i * dim + (size_t) sub * dsub, sizeof(float) * (size_t) dsub);
        orc_pq_train_subspace(S, nvec, dsub, ksub, draws + (size_t) sub * ksub, codebooks + (size_t) sub * ksub * dsub, max_iters);
    }
    free(S);
    return 0;
}

/* pq_encode_vector :479-503 for n rows */
void orc_pq_encode(const float *X, int64_t n, int dim, const float *codebooks, int m, int ksub, int16_t *codes)
{
    int dsub = dim / m;
    for (int64_t i = 0; i < n; i++)
        for (int sub = 0; sub < m; sub++) {
            int start_dim = sub * dsub, best = -1;
            double min_dist = DBL_MAX;
            for (int c = 0; c < ksub; c++) {
                double dist = 0.0;
                for (int d = 0; d < dsub; d++) {
                    double diff = (double) X[(size_t) i * dim + start_dim + d] - (double) codebooks[((size_t) sub * ksub + c) * dsub + d];
                    dist += diff * diff;
                }
                if (dist < min_dist) { min_dist = dist; best = c; }
            }
            codes[(size_t) i * m + sub] = (int16_t) best;
        }
}

/* pq_asymmetric_distance :1063-1098: ONE double chain over all dimensions, then (float) sqrt; -1 on an invalid code */
float orc_pq_asymmetric_distance(const float *q, const int16_t *codes, const float *codebooks, int m, int ksub, int dsub)
{
    double total_dist = 0.0;
    for (int sub = 0; sub < m; sub++) {
        int start_dim = sub * dsub, code = codes[sub];
        if (code < 0 || code >= ksub) return -1.0f;
        for (int d = 0; d < dsub; d++) {
            double diff = (double) q[start_dim + d] - (double) codebooks[((size_t) sub * ksub + code) * dsub + d];
            total_dist += diff * diff;
        }
    }
    return (float) sqrt(total_dist);
}

/* ORDER BY pq_asymmetric_distance(q, codes, codebook) LIMIT k over n encoded rows: distances of every row (dist_all,
 * optional nq*n) and the k nearest by (distance, row) */
typedef struct { float d; int64_t row; } OrcPqHit;
static int orc_pq_cmp(const void *a, const void *b)
{
    const OrcPqHit *x = (const OrcPqHit *) a, *y = (const OrcPqHit *) b;
    if (x->d < y->d) return -1;
    if (x->d > y->d) return 1;
    return x->row < y->row ? -1 : (x->row > y->row);
}
void orc_pq_knn(const float *Q, int nq, const int16_t *codes, int64_t n, const float *codebooks, int dim, int m, int ksub, int k,
                float *dist, int64_t *rows, float *dist_all)
{
    int dsub = dim / m;
#pragma omp parallel for schedule(dynamic, 1)
    for (int j = 0; j < nq; j++) {
        OrcPqHit *h = (OrcPqHit *) malloc(sizeof(OrcPqHit) * (size_t) n);
        for (int64_t i = 0; i < n; i++) {
            h[i].d = orc_pq_asymmetric_distance(Q + (size_t) j * dim, codes + (size_t) i * m, codebooks, m, ksub, dsub);
            h[i].row = i;
            if (dist_all) dist_all[(size_t) j * n + i] = h[i].d;
        }
        qsort(h, (size_t) n, sizeof(OrcPqHit), orc_pq_cmp);
        for (int i = 0; i < k; i++) {
            dist[(size_t) j * k + i] = i < n ? h[i].d : INFINITY;
            rows[(size_t) j * k + i] = i < n ? h[i].row : -1;
        }
        free(h);
    }
}

/* ---- per-vector quantisers: NeuronDB/src/types/quantization.c ---------------------------------------
 * quantize_vector_i8 :42-86, float4_to_fp16 :141-168 / quantize_vector_f16 :220-236, quantize_vector_binary :284-312,
 * binary_hamming_distance :385-427, quantize_vector_uint8 :1354-1402, quantize_vector_ternary :1455-1503,
 * quantize_vector_int4 :1562-1641 (the CPU branch).  Output = the data[] bytes of the varlena, rows packed.
 * PINNED against those functions compiled from the reference source (oracle/extract_ref_leafs.py). */
enum { ORC_Q_INT8 = 1, ORC_Q_FP16 = 2, ORC_Q_BINARY = 3, ORC_Q_UINT8 = 4, ORC_Q_TERNARY = 5, ORC_Q_INT4 = 6 };

int64_t orc_quantized_row_bytes(int kind, int dim)
{
    switch (kind) {
    case ORC_Q_INT8: case ORC_Q_UINT8: return dim;
    case ORC_Q_FP16: return 2 * (int64_t) dim;
    case ORC_Q_BINARY: return (dim + 7) / 8;
    case ORC_Q_TERNARY: return (dim * 2 + 7) / 8;
    case ORC_Q_INT4: return (dim + 1) / 2;
    }
    return -1;
}

static uint16_t orc_float_to_fp16(float f)            /* :141-168: truncates the mantissa, flushes subnormals, NaN -> inf */
{
    uint32_t u;
    memcpy(&u, &f, 4);
    uint16_t sign = (u >> 16) & 0x8000;
    uint32_t mantissa = u & 0x7fffff;
    int16_t exp = (int16_t) (((u >> 23) & 0xff) - 127 + 15);
    if (exp <= 0) return sign;
    if (exp >= 31) return sign | 0x7c00;
    return (uint16_t) (sign | (exp << 10) | (mantissa >> 13));
}

void orc_quantize_row(int kind, const float *v, int dim, uint8_t *out)
{
    int i;
    memset(out, 0, (size_t) orc_quantized_row_bytes(kind, dim));
    if (kind == ORC_Q_INT8) {
        float max_abs = 0.0f;
        for (i = 0; i < dim; i++) { float a = fabsf(v[i]); if (a > max_abs) max_abs = a; }
        if (max_abs == 0.0f) return;
        float scale = 127.0f / max_abs;
        for (i = 0; i < dim; i++) {
            float val = v[i] * scale;
            if (val > 127.0f) val = 127.0f;
            if (val < -128.0f) val = -128.0f;
            ((int8_t *) out)[i] = (int8_t) rintf(val);
        }
    } else if (kind == ORC_Q_FP16) {
        for (i = 0; i < dim; i++) { uint16_t h = orc_float_to_fp16(v[i]); memcpy(out + 2 * i, &h, 2); }
    } else if (kind == ORC_Q_BINARY) {
        for (i = 0; i < dim; i++) if (v[i] > 0.0f) out[i / 8] |= (uint8_t) (1 << (i % 8));
    } else if (kind == ORC_Q_UINT8) {
        float mn = 0.0f, mx = 0.0f;
        for (i = 0; i < dim; i++) {
            if (i == 0) mn = mx = v[i];
            else { if (v[i] < mn) mn = v[i]; if (v[i] > mx) mx = v[i]; }
        }
        if (mx == mn) return;
        float scale = 255.0f / (mx - mn);
        for (i = 0; i < dim; i++) {
            float nv = (v[i] - mn) * scale;
            if (nv > 255.0f) nv = 255.0f;
            if (nv < 0.0f) nv = 0.0f;
            out[i] = (uint8_t) rintf(nv);
        }
    } else if (kind == ORC_Q_TERNARY) {
        float max_abs = 0.0f;
        for (i = 0; i < dim; i++) { float a = fabsf(v[i]); if (a > max_abs) max_abs = a; }
        float threshold = max_abs / 3.0f;
        for (i = 0; i < dim; i++) {
            uint8_t value = v[i] > threshold ? 2 : (v[i] < -threshold ? 1 : 0);
            out[(i * 2) / 8] |= (uint8_t) (value << ((i * 2) % 8));
        }
    } else if (kind == ORC_Q_INT4) {
        float max_abs = 0.0f;
        for (i = 0; i < dim; i++) { float a = fabsf(v[i]); if (a > max_abs) max_abs = a; }
        if (max_abs == 0.0f) return;
        float scale = 7.0f / max_abs;
        for (i = 0; i < dim; i++) {
            int8_t value;
            float scaled = v[i] * scale;
            if (scaled > 7.0f) value = 7;
            else if (scaled < -8.0f) value = -8;
            else value = (int8_t) rintf(scaled);
            uint8_t uvalue = (uint8_t) (8 + value);
            if (uvalue > 15) uvalue = 15;
            out[i / 2] |= (uint8_t) (uvalue << ((i % 2) * 4));
        }
    }
}

void orc_quantize_rows(int kind, const float *X, int64_t n, int dim, uint8_t *out)
{
    int64_t rb = orc_quantized_row_bytes(kind, dim);
    for (int64_t i = 0; i < n; i++) orc_quantize_row(kind, X + (size_t) i * dim, dim, out + (size_t) i * rb);
}

int orc_hamming(const uint8_t *a, const uint8_t *b, int nbits)                     /* :409-424 */
{
    int count = 0, nbytes = (nbits + 7) / 8;
    for (int i = 0; i < nbytes; i++) count += __builtin_popcount((uint8_t) (a[i] ^ b[i]));
    return count;
}

/* ORDER BY binary_hamming_distance(bits, q) LIMIT k: k nearest rows by (distance, row); -1 / -1 past the end */
typedef struct { int d; int64_t row; } OrcHamHit;
static int orc_ham_cmp(const void *a, const void *b)
{
    const OrcHamHit *x = (const OrcHamHit *) a, *y = (const OrcHamHit *) b;
    if (x->d != y->d) return x->d < y->d ? -1 : 1;
    return x->row < y->row ? -1 : (x->row > y->row);
}
void orc_hamming_knn(const uint8_t *rows, int64_t n, int nbits, const uint8_t *Q, int nq, int k, int32_t *dist, int64_t *ids)
{
    int nbytes = (nbits + 7) / 8;
#pragma omp parallel for schedule(dynamic, 1)
    for (int j = 0; j < nq; j++) {
        OrcHamHit *h = (OrcHamHit *) malloc(sizeof(OrcHamHit) * (size_t) n);
        for (int64_t i = 0; i < n; i++) { h[i].d = orc_hamming(rows + (size_t) i * nbytes, Q + (size_t) j * nbytes, nbits); h[i].row = i; }
        qsort(h, (size_t) n, sizeof(OrcHamHit), orc_ham_cmp);
        for (int i = 0; i < k; i++) {
            dist[(size_t) j * k + i] = i < n ? h[i].d : -1;
            ids[(size_t) j * k + i] = i < n ? h[i].row : -1;
        }
        free(h);
    }
}

/* ---- cluster_minibatch_kmeans (NeuronDB/src/ml/ml_minibatch_kmeans.c:67-198 minibatch_kmeans_pp_init, :206-449) ----------
 * draws = the values rand() returns, in call order; *consumed = how many the function took (the seeding stops early when
 * the remaining weights sum below 1e-10, so the count is data dependent).  All-double arithmetic; the centroid update is
 * sequential over the batch: count++, eta = 1/count, c = (float)((1 - eta) c + eta x).  labels 1-based.
 * Returns 0, -1 / -2 / -3 on the argument errors (:238-300), -4 when the draws run out.
 * PINNED against the reference's own minibatch_kmeans_pp_init and the text of the main loop (oracle/extract_ref_leafs.py). */
int orc_cluster_minibatch_kmeans(const float *X, int nvec, int dim, int k, int batch_size, int max_iters, const int *draws, int ndraws,
                                 int rand_max, int *consumed, int *labels, float *centers_out)
{
    if (k < 2) return -1;                          /* "num_clusters must be at least 2" */
    if (batch_size < 1) return -2;                 /* "batch_size must be at least 1" */
    if (max_iters < 1) max_iters = 100;
    if (nvec < k) return -3;                       /* "Not enough vectors (%d) for %d clusters" */
    if (batch_size > nvec) batch_size = nvec;
    int cur = 0;
#define ORC_NEXT_RAND() (cur < ndraws ? draws[cur++] : (cur++, -1))
    float *C = (float *) calloc((size_t) k * dim, sizeof(float));
    unsigned char *selected = (unsigned char *) calloc((size_t) nvec, 1);
    double *dist = (double *) calloc((size_t) nvec, sizeof(double));
    int *counts = (int *) calloc((size_t) k, sizeof(int));
    int *bidx = (int *) malloc(sizeof(int) * (size_t) batch_size), *bass = (int *) malloc(sizeof(int) * (size_t) batch_size);
    {   /* minibatch_kmeans_pp_init */
        int first = ORC_NEXT_RAND() % nvec;
        memcpy(C, X + (size_t) first * dim, sizeof(float) * (size_t) dim);
        selected[first] = 1;
        for (int i = 0; i < nvec; i++) {
            double acc = 0.0;
            for (int d = 0; d < dim; d++) { double diff = (double) X[(size_t) i * dim + d] - (double) C[d]; acc += diff * diff; }
            dist[i] = acc;
        }
        for (int c = 1; c < k; c++) {
            double sum = 0.0, r;
            int picked = -1;
            for (int i = 0; i < nvec; i++) if (!selected[i]) sum += dist[i];
            if (sum < 1e-10) break;
            r = ((double) ORC_NEXT_RAND() / rand_max) * sum;
            for (int i = 0; i < nvec; i++) {
                if (selected[i]) continue;
                r -= dist[i];
                if (r <= 0.0) { picked = i; break; }
            }
            if (picked < 0) for (int i = 0; i < nvec; i++) if (!selected[i]) { picked = i; break; }
            if (picked < 0) break;
            memcpy(C + (size_t) c * dim, X + (size_t) picked * dim, sizeof(float) * (size_t) dim);
            selected[picked] = 1;
            for (int i = 0; i < nvec; i++) {
                if (selected[i]) continue;
                double acc = 0.0;
                for (int d = 0; d < dim; d++) { double diff = (double) X[(size_t) i * dim + d] - (double) C[(size_t) c * dim + d]; acc += diff * diff; }
                if (acc < dist[i]) dist[i] = acc;
            }
        }
    }
    for (int iter = 0; iter < max_iters; iter++) {
        for (int i = 0; i < batch_size; i++) bidx[i] = ORC_NEXT_RAND() % nvec;
        for (int i = 0; i < batch_size; i++) {
            double min_dist = DBL_MAX;
            int best = 0;
            for (int c = 0; c < k; c++) {
                double d2 = 0.0;
                for (int d = 0; d < dim; d++) { double diff = (double) X[(size_t) bidx[i] * dim + d] - (double) C[(size_t) c * dim + d]; d2 += diff * diff; }
                if (d2 < min_dist) { min_dist = d2; best = c; }
            }
            bass[i] = best;
        }
        for (int i = 0; i < batch_size; i++) {
            int cl = bass[i];
            counts[cl]++;
            double lr = 1.0 / counts[cl];
            for (int d = 0; d < dim; d++)
                C[(size_t) cl * dim + d] = (float) ((1.0 - lr) * C[(size_t) cl * dim + d] + lr * X[(size_t) bidx[i] * dim + d]);
        }
    }
    for (int i = 0; i < nvec; i++) {
        double min_dist = DBL_MAX;
        int best = 0;
        for (int c = 0; c < k; c++) {
            double d2 = 0.0;
            for (int d = 0; d < dim; d++) { double diff = (double) X[(size_t) i * dim + d] - (double) C[(size_t) c * dim + d]; d2 += diff * diff; }
            if (d2 < min_dist) { min_dist = d2; best = c; }
        }
        labels[i] = best + 1;
    }
#undef ORC_NEXT_RAND
    if (centers_out) memcpy(centers_out, C, sizeof(float) * (size_t) k * dim);
    if (consumed) *consumed = cur;
    free(C); free(selected); free(dist); free(counts); free(bidx); free(bass);
    return cur > ndraws ? -4 : 0;
}
