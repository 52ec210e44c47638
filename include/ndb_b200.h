/*
 * ndb_b200.h -- C ABI of the B200-native vector-search hot path for NeuronDB.
 *
 * This is the drop-in boundary: a C-ABI shared library
 * (neurondb_b200/lib/libndb_b200.so) with plain pointers and sizes.  Every entry
 * point names the reference interface it replaces (paths relative to the reference
 * repository root).  INTEGRATION.md shows the reference-side binding: the
 * `ndb_gpu_backend` vtable instance and the index-AM call sites.
 *
 * Conventions (the reference's own, NeuronDB/include/neurondb_gpu_backend.h:24-27):
 *   - return 0 on success, a negative NDB_B200_E* code on failure; never throws, never
 *     ereport()s; the caller turns a failure into ereport(ERROR).  There is NO CPU
 *     fallback: a missing GPU is NDB_B200_ENOTINIT / NDB_B200_ECUDA.
 *   - all `const float *` / output arguments without a `_dev` suffix are HOST pointers,
 *     row-major fp32, caller-owned; they are copied in/out before the call returns and
 *     never retained.  `_dev` entry points take DEVICE pointers and a cudaStream_t
 *     (passed as void *) and return without synchronising.
 *   - handles are opaque, owned by the library, freed explicitly or at shutdown; one
 *     handle may be used by one thread at a time (one PostgreSQL backend = one thread).
 *   - k-nearest results are sorted by (distance ASC, id ASC), the reference's only
 *     explicit tie rule (NeuronDB/src/util/distributed.c:425-438); missing results are
 *     (+inf, -1).
 *   - "arith" selects which of the reference's arithmetic variants is reproduced
 *     bit for bit (see NDB_ARITH_*); NDB_ARITH_FAST / NDB_ARITH_TENSOR are the
 *     tolerance-bounded fast paths (1e-5 / 1e-3 relative, BASELINE.json north_star).
 */
#ifndef NDB_B200_H
#define NDB_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define NDB_B200_ABI_VERSION 2

/* error codes */
#define NDB_B200_OK        0
#define NDB_B200_EINVAL   (-1)   /* bad argument (NULL, n <= 0, k out of range ...)          */
#define NDB_B200_ECUDA    (-2)   /* CUDA runtime failure; see ndb_b200_last_error()          */
#define NDB_B200_ENOTINIT (-3)   /* ndb_b200_init() not called / no device                   */
#define NDB_B200_EVECTOR  (-4)   /* NaN/Inf in a vector (vector_distance.c:55-73 raises)     */
#define NDB_B200_EDIM     (-5)   /* dimension mismatch (vector_distance.c:42-47)             */
#define NDB_B200_ENOMEM   (-6)
#define NDB_B200_ESTATE   (-7)   /* index not trained / empty / wrong handle kind            */
#define NDB_B200_ERANGE   (-8)   /* result NaN/Inf (vector_distance.c:117-120)               */

/* metrics: the AMs' sk_strategy numbering (hnsw_am.c:1312-1337, ivf_am.c:1561-1591) */
#define NDB_L2      1
#define NDB_COSINE  2
#define NDB_IP      3

/* arithmetic variants of the reference that the kernels reproduce bit for bit */
#define NDB_ARITH_OP_F64   0  /* <->,<=>,<#> default build: fp64, Kahan for L2
                                 (src/vector/vector_distance.c:93-122,145-157,180-213);
                                 <#> yields +dot (vector_distance_simd.c:557)                  */
#define NDB_ARITH_AVX2     1  /* the operators as an AVX2 build of the reference computes them
                                 (vector_distance_simd.c:159-185,233-258,300-345): 8 f32 lane
                                 accumulators, fixed reduction tree (:85-101), scalar tail;
                                 ndb_b200_distance_pairs / _rows only                           */
#define NDB_ARITH_AVX512   2  /* the same with 16 lanes (:188-217,261-291,348-392,120-137)      */
#define NDB_ARITH_IVF_F32  3  /* ivfComputeDistance: f32 sequential (ivf_am.c:1550-1592);
                                 NDB_IP = -dot f32 sequential (not in the reference, Q5)     */
#define NDB_ARITH_HNSW     4  /* hnswComputeDistance: f32 op, f64 accumulate
                                 (hnsw_am.c:1301-1345)                                        */
#define NDB_ARITH_FAST     5  /* fp32 FFMA, several accumulators: <= 1e-5 relative          */
#define NDB_ARITH_TENSOR   6  /* bf16 tcgen05 tiles, fp32 accumulate; all three metrics, dim <= 2048,
                                 k <= 16 (ndb_b200_knn_exact) / k <= 32 (ndb_b200_ivf_search), anything else
                                 -> NDB_B200_EINVAL (no fallback).
                                 ndb_b200_knn_exact: distances <= 1e-3 relative.
                                 ndb_b200_ivf_search: the tensor cores only PROPOSE candidates (coarse
                                 quantiser and list scans); they are re-evaluated with NDB_ARITH_IVF_F32
                                 and the answer is certified or recomputed exactly (ndb_b200_ivf_cert_stats),
                                 so ids and distances equal the NDB_ARITH_IVF_F32 search's, bit for bit   */

/* IVF search modes */
#define NDB_IVF_FULL     0    /* scan every probed list completely                             */
#define NDB_IVF_LITERAL  1    /* ivfCollectCandidates as written: stop after k*10 candidates in
                                 probe order, selection-sort ties (ivf_am.c:1743,1861-1881)   */
/* HNSW search / build modes */
#define NDB_HNSW_BESTFIRST 1  /* search_layer with beam ef (what hnsw_scan.c intends)        */
#define NDB_HNSW_LITERAL   0  /* hnswSearch as written: BFS until ef candidates (Q12)        */

typedef struct ndb_b200_dataset ndb_b200_dataset;   /* rows resident in HBM (seq-scan target)  */
typedef struct ndb_b200_ivf     ndb_b200_ivf;       /* IVF index: centroids + inverted lists   */
typedef struct ndb_b200_hnsw    ndb_b200_hnsw;      /* HNSW graph                              */

/* ---- lifecycle: ndb_gpu_backend.init/.shutdown/.is_available/.device_count/.set_device
 *      (include/neurondb_gpu_backend.h:38-45; CUDA instance src/gpu/cuda/gpu_backend_cuda.c:162-260) */
int         ndb_b200_init(int device);
void        ndb_b200_shutdown(void);
int         ndb_b200_is_available(void);
int         ndb_b200_device_count(void);
int         ndb_b200_abi_version(void);
const char *ndb_b200_last_error(void);
/* device facts for NDBGpuDeviceInfo (include/neurondb_gpu_types.h) */
int         ndb_b200_device_info(int device, char *name, size_t name_len, size_t *total_mem,
                                 size_t *free_mem, int *cc_major, int *cc_minor, int *sm_count);
/* ndb_gpu_backend.mem_alloc/.mem_free/.memcpy_h2d/.memcpy_d2h (neurondb_gpu_backend.h:48-51) */
int         ndb_b200_mem_alloc(void **ptr, size_t bytes);
int         ndb_b200_mem_free(void *ptr);
int         ndb_b200_memcpy_h2d(void *dst, const void *src, size_t bytes);
int         ndb_b200_memcpy_d2h(void *dst, const void *src, size_t bytes);
int         ndb_b200_host_alloc_pinned(void **ptr, size_t bytes);
int         ndb_b200_host_free_pinned(void *ptr);
/* ndb_gpu_backend.stream_create/.stream_destroy/.stream_synchronize (:351-353) */
int         ndb_b200_stream_create(void **stream);
int         ndb_b200_stream_destroy(void *stream);
int         ndb_b200_stream_synchronize(void *stream);

/* ---- b2 vtable launchers: paired rows, out[i] = dist(A_i, B_i), host pointers
 *      .launch_l2_distance / .launch_cosine (neurondb_gpu_backend.h:54-65; replaces the
 *      per-pair cuBLAS path gpu_backend_cuda.c:387-537).  fp32, sqrtf; cosine clamps to
 *      [-1,1] and returns 1.0 on a non-positive norm, as the CUDA backend does (:518-529). */
int ndb_b200_launch_l2_distance(const float *A, const float *B, float *out, int n, int d, void *stream);
int ndb_b200_launch_cosine(const float *A, const float *B, float *out, int n, int d, void *stream);
/* .launch_kmeans_assign / .launch_kmeans_update (:66-79; gpu_kmeans_kernels.cu:53-155,
 * gpu_backend_cuda.c:677-710): argmin squared L2 with strict <, per-cluster mean */
int ndb_b200_launch_kmeans_assign(const float *X, const float *C, int *idx, int n, int d, int k, void *stream);
int ndb_b200_launch_kmeans_update(const float *X, const int *idx, float *C, int n, int d, int k, void *stream);

/* ---- operator backends: vector_l2_distance_op / vector_cosine_distance_op /
 *      vector_inner_product_distance_op (src/index/opclass.c:53-161) and the SQL batch
 *      functions vector_*_distance_batch / vector_*_distance_gpu (src/vector/vector_batch.c:37-160,
 *      src/gpu/common/gpu_sql.c:90-160), n pairs at once.  NDB_B200_EVECTOR on NaN/Inf input. */
int ndb_b200_distance_pairs(int metric, int arith, const float *A, const float *B, float *out,
                            int64_t n, int dim);
/* one query against n rows (ORDER BY v <-> q without LIMIT; neurondb_gpu_batch_l2_distance,
 * src/gpu/common/gpu_batch.c:55-81) */
int ndb_b200_distance_rows(int metric, int arith, const float *X, int64_t n, int dim,
                           const float *q, float *out);

/* vector_l2_distance_batch / vector_cosine_distance_batch / vector_inner_product_distance_batch
 * (src/vector/vector_batch.c:37-420): rows = n x dim floats, dims[i] = the dimension of array element i (0 = a NULL
 * element).  An element whose dimension differs from `dim` (the query's) yields nulls[i] = 1 (:123-150); the others
 * get l2_distance / cosine_distance / -inner_product_distance (= +dot) in the operators' fp64 arithmetic. */
int ndb_b200_vector_distance_batch(int metric, const float *rows, const int *dims, int64_t n, int dim, const float *query,
                                   float *out, uint8_t *nulls);

/* ---- dataset: the heap column a SeqScan reads ------------------------------- */
int ndb_b200_dataset_create(int dim, ndb_b200_dataset **out);
int ndb_b200_dataset_append(ndb_b200_dataset *ds, const float *rows, const int64_t *ids, int64_t n);
int ndb_b200_dataset_append_dev(ndb_b200_dataset *ds, const float *rows_dev, const int64_t *ids_dev,
                                int64_t n, void *stream);
int64_t ndb_b200_dataset_size(const ndb_b200_dataset *ds);
void ndb_b200_dataset_free(ndb_b200_dataset *ds);
/* exact kNN = SeqScan + top-N sort over `metric` evaluated with `arith`
 * (SELECT ... ORDER BY v <-> q LIMIT k; opclass.c:53-83 -> vector_distance.c:93-122) */
int ndb_b200_knn_exact(ndb_b200_dataset *ds, int metric, int arith, const float *Q, int nq, int k,
                       float *dist, int64_t *ids);
int ndb_b200_knn_exact_dev(ndb_b200_dataset *ds, int metric, int arith, const float *Q_dev, int nq,
                           int k, float *dist_dev, int64_t *ids_dev, void *stream);

/* knn_classify / knn_regress (src/ml/ml_knn.c:112-357, 363-569; brute-force kernels src/gpu/cuda/gpu_knn_kernels.cu):
 * euclidean_distance (:76-90: f32 difference, f64 sum) of every resident row to each query, the k nearest, then the
 * binary majority vote (labels other than 0 / 1 are ignored, class 1 needs a strict majority) or the mean of the
 * targets.  labels[i] / targets[i] belong to the i-th appended row.  k < 1 -> EINVAL, fewer than k rows -> ERANGE,
 * NaN / Inf in a query -> EVECTOR, as the SQL functions raise. */
int ndb_b200_knn_classify(ndb_b200_dataset *ds, const double *labels, const float *Q, int nq, int k, int *out_class);
int ndb_b200_knn_regress(ndb_b200_dataset *ds, const double *targets, const float *Q, int nq, int k, double *out);

/* cluster_kmeans (src/ml/ml_kmeans.c:146-303) over n rows of X (row-major, what neurondb_fetch_vectors_from_table
 * returns): k-means++ seeding (kmeanspp_init :45-139: float difference and square, double sums, the D^2-weighted walk
 * in row order) and Lloyd's iterations until no assignment changes or max_iters (< 1 -> 100, :175-176); assignment by
 * neurondb_l2_distance_squared (util/neurondb_simd_impl.c:36-104, all double; strict <, lowest index wins), update =
 * float sums in row order / count, empty clusters end at zero.  rand_draws = the k values rand() returns, in call
 * order (the reference draws exactly k and nothing else in between: the caller draws them, the stream stays where the
 * reference leaves it); rand_max = RAND_MAX.  labels[n] are 1-based as the SQL function returns them (:286);
 * centers[k*dim], iters, seeds[k] (the rows kmeanspp_init picked) are optional.  k <= 1 or n < k -> EINVAL with the
 * reference's messages; NaN / Inf in X -> EVECTOR. */
int ndb_b200_cluster_kmeans(const float *X, int n, int dim, int k, int max_iters, const int *rand_draws, int rand_max,
                            int *labels, float *centers, int *iters, int *seeds);

/* cluster_minibatch_kmeans (src/ml/ml_minibatch_kmeans.c:206-449; seeding minibatch_kmeans_pp_init :67-198, all-double
 * weights, stops early when the remaining weights sum below 1e-10 -- centroids it never reaches stay zero).  Per iteration
 * batch_size rows rand() % n, nearest centroid in double, then sequentially over the batch: count++, eta = 1 / count,
 * c = (float) ((1 - eta) c + eta x); finally every row's nearest centroid, labels 1-based.  How often the reference calls
 * rand() depends on the data, so the caller passes the function: next_rand(rand_state) is called exactly when and as
 * often as the reference calls rand() (the glue passes a wrapper around rand() itself).  batch_size > n -> n, max_iters < 1
 * -> 100; k < 2, batch_size < 1, n < k -> EINVAL with the reference's messages; NaN / Inf -> EVECTOR. */
typedef int (*ndb_b200_rand_fn)(void *state);
int ndb_b200_cluster_minibatch_kmeans(const float *X, int n, int dim, int k, int batch_size, int max_iters, ndb_b200_rand_fn next_rand,
                                      void *rand_state, int rand_max, int *labels, float *centers);

/* ---- product quantisation (src/ml/ml_product_quantization.c).  Codebooks are laid out as the reference's
 *      bytea carries them after its three ints: float centroids[m][ksub][dsub], dsub = dim / m.  Arithmetic is the SQL
 *      functions': double difference / square / sum, strict <, lowest code wins; results are bit-identical to them.
 *      Shape errors carry the reference's messages: m outside 1..128 or ksub outside 2..65536 -> EINVAL (:218-227),
 *      dim % m != 0 -> EDIM (:270-276), NaN / Inf -> EVECTOR. */
/* train_pq_codebook (:195-415): per subspace train_subspace_kmeans (:80-190) -- seeds = rows rand() % n, Lloyd until no
 * assignment changes or max_iters (the reference passes 100; < 0 -> 100).  rand_draws = the m*ksub values rand() returns,
 * in call order (subspace-major). */
int ndb_b200_pq_train(const float *X, int n, int dim, int m, int ksub, int max_iters, const int *rand_draws, float *codebooks);
/* pq_encode_vector (:421-536) for n rows: int2 codes [n][m] as the SQL function returns them (ksub <= 32768) */
int ndb_b200_pq_encode(const float *X, int64_t n, int dim, const float *codebooks, int m, int ksub, int16_t *codes);
/* ndb_gpu_backend.launch_pq_encode (include/neurondb_gpu_backend.h:106-113; CUDA instance gpu_backend_cuda.c:711-732 ->
 * gpu_pq_encode_batch, src/gpu/cuda/gpu_pq_kernels.cu:54-110, 163-213): the same signature, byte codes (ks <= 256) */
int ndb_b200_launch_pq_encode(const float *X, const float *codebooks, uint8_t *codes, int n, int d, int m, int ks, void *stream);
/* Resident encoded rows + the asymmetric-distance scan: ORDER BY pq_asymmetric_distance(q, codes, codebook) LIMIT k
 * (:1003-1110; gpu_pq_asymmetric_distance_batch, gpu_pq_kernels.cu:127-158, 215-269).  ksub <= 256 (byte codes on the
 * device), m * ksub * 8 bytes must fit shared memory.  add encodes and appends (codes_out optional, [n][m]); add_codes
 * appends codes made elsewhere (a code outside 0..ksub-1 -> ERANGE with the reference's message, nothing appended);
 * search returns the k nearest rows by (distance, row), +inf / -1 past the end; distances returns all nq * n of them
 * (rechecked, optional: how many fell back from the table sum to the reference's chain). */
typedef struct ndb_b200_pq ndb_b200_pq;
int ndb_b200_pq_create(int dim, int m, int ksub, const float *codebooks, ndb_b200_pq **out);
void ndb_b200_pq_free(ndb_b200_pq *pq);
int64_t ndb_b200_pq_size(const ndb_b200_pq *pq);
int ndb_b200_pq_add(ndb_b200_pq *pq, const float *X, int64_t n, int16_t *codes_out);
int ndb_b200_pq_add_codes(ndb_b200_pq *pq, const int16_t *codes, int64_t n);
int ndb_b200_pq_search(ndb_b200_pq *pq, const float *Q, int nq, int k, float *dist, int64_t *rows);
int ndb_b200_pq_search_dev(ndb_b200_pq *pq, const float *Q_dev, int nq, int k, float *dist_dev, int64_t *rows_dev, void *stream);
int ndb_b200_pq_distances(ndb_b200_pq *pq, const float *Q, int nq, float *dist, unsigned long long *rechecked);

/* ---- per-vector quantisers (src/types/quantization.c) for n rows at once, and the Hamming scan over binary rows.
 *      out receives the data[] bytes of the varlena each function builds, rows packed (ndb_b200_quantized_row_bytes each):
 *      INT8    quantize_vector_i8 :42-86 (rintf(x * 127 / max|x|), zero row -> zeros)
 *      FP16    quantize_vector_f16 :220-236 / float4_to_fp16 :141-168 (mantissa truncated, subnormals flushed, overflow -> inf)
 *      BINARY  quantize_vector_binary :284-312 (bit i%8 of byte i/8 = x > 0)
 *      UINT8   quantize_vector_uint8 :1354-1402 (rintf((x - min) * 255 / (max - min)), constant row -> zeros)
 *      TERNARY quantize_vector_ternary :1455-1503 (2 bits per dimension against max|x| / 3)
 *      INT4    quantize_vector_int4 :1562-1641 (nibble 8 + rintf(x * 7 / max|x|), low nibble first, zero row -> zero bytes)
 *      Bit-identical to those functions.  (The backend's launch_quant_* members are NOT bound to these: the reference's CUDA
 *      kernels behind them round where the SQL functions truncate, take a caller-supplied scale, and so define other results.) */
#define NDB_QUANT_INT8    1
#define NDB_QUANT_FP16    2
#define NDB_QUANT_BINARY  3
#define NDB_QUANT_UINT8   4
#define NDB_QUANT_TERNARY 5
#define NDB_QUANT_INT4    6
int64_t ndb_b200_quantized_row_bytes(int kind, int dim);
int ndb_b200_quantize_rows(int kind, const float *X, int64_t n, int dim, void *out);
/* ORDER BY binary_hamming_distance(bits, q) LIMIT k (:385-427): rows / Q are (nbits+7)/8 bytes each; the k nearest rows
 * by (distance, row index); -1 / -1 past the end. */
int ndb_b200_hamming_knn(const uint8_t *rows, int64_t n, int nbits, const uint8_t *Q, int nq, int k, int32_t *dist, int64_t *ids);

/* ---- IVF k-means: kmeans_init/run/assign/update_centroids/compute_cost
 *      (src/index/ivf_am.c:2070-2294).  Literal semantics: centroids := first k rows,
 *      <= max_iter Lloyd steps, stop when |prevCost - cost| < tol, f32 sequential sums.
 *      C is k*d, assign n, counts k (any may be NULL).  Returns iterations in *iters. */
int ndb_b200_kmeans_train(const float *X, int n, int d, int k, int max_iter, float tol,
                          float *C, int *assign, int *counts, int *iters, float *cost);

/* Row-sharded training (one process per GPU, rows split, centroids replicated): the local half of
 * one Lloyd iteration.  shard_step assigns the n local rows to the k centroids C_dev
 * (find_nearest_centroid, ivf_am.c:2274-2294) and writes the per-cluster f32 sums [k*d] and counts [k]
 * of the local members (kmeans_update_centroids :2182-2213 without the division); the caller
 * all-reduces sums and counts over the ranks, divides (empty cluster -> zeros), and calls shard_cost
 * with the new centroids for its share of kmeans_compute_cost (:2218-2233), all-reducing that too.
 * All pointers are device pointers.  With a single rank the sequence is bit-identical to kmeans_train as long as
 * n <= 2^20 rows; above that the per-cluster sums and the cost are accumulated by fixed trees (deterministic,
 * equal to the sequential loops to fp32 rounding -- which the multi-rank result is anyway). */
int ndb_b200_kmeans_shard_step_dev(const float *X_dev, int64_t n, int d, int k, const float *C_dev, int *assign_dev,
                                   float *sums_dev, int *counts_dev, void *stream);
int ndb_b200_kmeans_shard_cost_dev(const float *X_dev, int64_t n, int d, const float *C_dev, const int *assign_dev,
                                   float *cost_dev, void *stream);

/* ---- IVF index: ivfbuild / ivfinsert / ivfrescan+ivfgettuple (src/index/ivf_am.c) -------- */
int ndb_b200_ivf_create(int dim, int nlists, int metric, ndb_b200_ivf **out);
void ndb_b200_ivf_free(ndb_b200_ivf *ix);
/* ivfbuild (:501-745): k-means on the first min(10000, nlists*100) rows (:580) */
int ndb_b200_ivf_train(ndb_b200_ivf *ix, const float *rows, int64_t n);
int ndb_b200_ivf_set_centroids(ndb_b200_ivf *ix, const float *C);          /* nlists*dim */
int ndb_b200_ivf_get_centroids(const ndb_b200_ivf *ix, float *C);
/* ivfinsert (:797-1167) for n rows in order: nearest centroid by sqrtf(sum f32 diff^2),
 * strict < (:906-935), append to that list.  out_list (may be NULL) receives the list ids. */
int ndb_b200_ivf_insert(ndb_b200_ivf *ix, const float *rows, const int64_t *ids, int64_t n, int *out_list);
/* assignment only (the :906-935 loop), host in / host out */
int ndb_b200_ivf_assign(const ndb_b200_ivf *ix, const float *rows, int64_t n, int *out_list);
/* load an index relation image: block 0 = IvfMetaPageData, centroid page, list page chains
 * (layouts ivf_am.c:75-106,241-261; PostgreSQL page header/line pointers).  Pages are staged
 * through pinned memory and decoded on the device. */
int ndb_b200_ivf_load_relation(ndb_b200_ivf *ix, const void *blocks, uint32_t nblocks);
/* the tail of the build: list layout of the inserted rows and (NDB_ARITH_TENSOR) the blocked bf16 copy.
 * Searches do it lazily on first use; call it to keep it out of the first query (and inside build timing). */
int ndb_b200_ivf_prepare(ndb_b200_ivf *ix, int arith);
int64_t ndb_b200_ivf_size(const ndb_b200_ivf *ix);
int ndb_b200_ivf_dim(const ndb_b200_ivf *ix);
int ndb_b200_ivf_list_sizes(const ndb_b200_ivf *ix, int64_t *sizes /* nlists */);
/* ivfrescan + ivfgettuple for a batch (:1439-1545,1911-2027): ivfSelectClusters (:1597-1717,
 * always L2) then ivfCollectCandidates (:1722-1909) with the index metric.
 * mode NDB_IVF_FULL | NDB_IVF_LITERAL; arith NDB_ARITH_IVF_F32 | _FAST | _TENSOR. */
int ndb_b200_ivf_search(ndb_b200_ivf *ix, const float *Q, int nq, int nprobe, int k, int mode,
                        int arith, float *dist, int64_t *ids);
int ndb_b200_ivf_search_dev(ndb_b200_ivf *ix, const float *Q_dev, int nq, int nprobe, int k, int mode,
                            int arith, float *dist_dev, int64_t *ids_dev, void *stream);
/* Pipelined form of ndb_b200_ivf_search for a stream of batches (a scan feeding executor batches):
 * begin queues one batch and returns a ticket, end(ticket) blocks until that batch's dist / ids are
 * filled.  Up to two batches per index may be in flight; the copies of neighbouring batches run on
 * their own streams and overlap the kernels, which still run one batch at a time.  Q, dist and ids
 * must stay valid until end (pinned memory makes the copies true DMA).  A third begin without an end
 * -> NDB_B200_ESTATE; a NaN/Inf query is reported by end (NDB_B200_EVECTOR). */
int ndb_b200_ivf_search_begin(ndb_b200_ivf *ix, const float *Q, int nq, int nprobe, int k, int mode,
                              int arith, float *dist, int64_t *ids, int *ticket);
int ndb_b200_ivf_search_end(ndb_b200_ivf *ix, int ticket);
/* ivf_knn_search_gpu(index, query, k, nprobe) (src/gpu/common/gpu_sql.c:931-1456): the SQL function's argument
 * checks and its (id, distance) rows; *nresults <= k rows are valid.  k, nprobe <= 128 here. */
int ndb_b200_ivf_knn_search_gpu(ndb_b200_ivf *ix, const float *query, int dim, int k, int nprobe, int64_t *ids, float *dist,
                                int *nresults);
/* NDB_ARITH_TENSOR searches are certified (csrc/ivf_cert.cuh): the bf16 tensor-core scan only proposes
 * candidates; each query's answer is accepted when a rounding-error bound proves that no row outside the
 * re-evaluated candidates can enter the reference's top k, and recomputed exactly (every row of the probed
 * lists, ivfComputeDistance's arithmetic) otherwise -- so ids and distance bits equal NDB_ARITH_IVF_F32's.
 * Statistics of the last such search: out[0] queries the 32-candidate certificate rejected (list scan),
 * out[1] exact re-evaluations of list candidates, out[2] / out[3] the same for the coarse quantiser,
 * out[4] queries for which every row of the probed lists had to be evaluated, out[5] rows evaluated by the exact
 * kernel's segment rescans. */
int ndb_b200_ivf_cert_stats(ndb_b200_ivf *ix, int64_t *out /* 6 */);
/* ivfSelectClusters alone: probe lists per query, nq*nprobe ints, -1 = none (:1597-1717) */
int ndb_b200_ivf_select_clusters(ndb_b200_ivf *ix, const float *Q, int nq, int nprobe, int *probes);
/* multi-GPU: keep only the lists l with l % world == rank (lists partition across ranks,
 * centroids stay replicated) -- call before insert/load */
int ndb_b200_ivf_set_shard(ndb_b200_ivf *ix, int rank, int world);

/* ---- HNSW: hnswbuild/hnswinsert/hnswInsertNode, hnswrescan/hnswgettuple/hnswSearch
 *      (src/index/hnsw_am.c:343-538, 904-1056, 1545-2080, 2091-2670) ------------------------ */
int ndb_b200_hnsw_create(int dim, int m, int ef_construction, int ef_search, int metric,
                         ndb_b200_hnsw **out);
void ndb_b200_hnsw_free(ndb_b200_hnsw *h);
/* build from rows; node id = row index (the reference's BlockNumber - 1).  levels: optional
 * n ints (hnswGetRandomLevel draws, :1143-1161); NULL = draw with libc random() seeded by seed.
 * batch = insertion batch size (1 = the sequential algorithm, id-exact with the oracle). */
int ndb_b200_hnsw_build(ndb_b200_hnsw *h, const float *rows, const int64_t *ids, int64_t n,
                        const int *levels, unsigned seed, int batch);
/* load a prebuilt graph (flat arrays as exported by the oracle / decoded from node pages) */
int ndb_b200_hnsw_load_graph(ndb_b200_hnsw *h, const float *rows, const int64_t *ids, int64_t n,
                             const int *levels, const uint32_t *nbr0, const int16_t *cnt,
                             const int64_t *upper_off, const uint32_t *upper,
                             uint32_t entry_point, int entry_level);
int ndb_b200_hnsw_export_graph(const ndb_b200_hnsw *h, int *levels, uint32_t *nbr0, int16_t *cnt,
                               int64_t *upper_off, uint32_t *upper, int64_t upper_cap,
                               uint32_t *entry_point, int *entry_level);
int64_t ndb_b200_hnsw_size(const ndb_b200_hnsw *h);
/* Neighbour selection of ndb_b200_hnsw_build.  NDB_HNSW_SELECT_CLOSEST (default) is the reference:
 * the closest m candidates become the forward links, and a back-link is written only while the
 * neighbour has a free slot among its 2m (hnsw_am.c:2386-2424,2493-2513; the prune block after it,
 * :2515-2612, can never run).  On a million rows that leaves late nodes nearly without in-links and
 * recall@10 below 0.9 at any ef.  NDB_HNSW_SELECT_HEURISTIC is an extension for the recall target of
 * the benchmark: forward links chosen by the diversity heuristic of Malkov & Yashunin (Alg. 4), and
 * a full neighbour re-selects its 2m links among the current ones and the new node.  m <= 16.
 * Search and the stored layout are unchanged. */
#define NDB_HNSW_SELECT_CLOSEST    0
#define NDB_HNSW_SELECT_HEURISTIC  1
int ndb_b200_hnsw_set_select(ndb_b200_hnsw *h, int select_mode);
/* load node pages (one HnswNodeData per 8 KB page, hnsw_am.c:124-181) + the meta page */
int ndb_b200_hnsw_load_relation(ndb_b200_hnsw *h, const void *blocks, uint32_t nblocks);
/* hnswSearch for a batch: strategy = metric (the AMs always pass 1, Q4), ef, k;
 * out nodes are row ids (heap TIDs in the reference come from a second page read, :1025-1052) */
int ndb_b200_hnsw_search(ndb_b200_hnsw *h, const float *Q, int nq, int strategy, int ef, int k,
                         int mode, float *dist, int64_t *ids);
int ndb_b200_hnsw_search_dev(ndb_b200_hnsw *h, const float *Q_dev, int nq, int strategy, int ef, int k,
                             int mode, float *dist_dev, int64_t *ids_dev, void *stream);
/* replicas: the graph does not shard; the rank that built it broadcasts it (ncclBroadcast) and
 * every rank then answers its own share of the queries.  Collective over the communicator's ranks. */
int ndb_b200_hnsw_broadcast(ndb_b200_hnsw *h, int root);
/* hnsw_knn_search_gpu(index, query, k, ef_search) (src/gpu/common/gpu_sql.c:498-930): argument checks + result */
int ndb_b200_hnsw_knn_search_gpu(ndb_b200_hnsw *h, const float *query, int dim, int k, int ef_search, int64_t *ids, float *dist,
                                 int *nresults);
/* distance evaluations of the last search batch (for the bytes-per-query roofline) */
int64_t ndb_b200_hnsw_last_evals(const ndb_b200_hnsw *h);

/* ---- multi-GPU merge: merge_distributed_results (src/util/distributed.c:320-487) ---------
 * dist/ids are [nshards][nq][k] device arrays (the NCCL all-gather result); output [nq][k]
 * sorted by (dist, id). */
int ndb_b200_merge_topk_dev(const float *dist_dev, const int64_t *ids_dev, int nshards, int nq, int k,
                            float *out_dist_dev, int64_t *out_ids_dev, void *stream);
int ndb_b200_merge_topk(const float *dist, const int64_t *ids, int nshards, int nq, int k,
                        float *out_dist, int64_t *out_ids);

/* ---- multi-GPU: one process per GPU, NCCL over NVLink ----------------------
 * Replaces the reference's scale-out path, NeuronDB/src/util/distributed.c: a coordinator that runs the
 * query on every shard over libpq (:56-300) and merges the per-shard (id, distance) rows on the host by
 * (distance ASC, id ASC) (:323-487; order :425-438).  Here a shard is one GPU of the box.
 *   comm_unique_id  rank 0 draws the 128-byte NCCL id; the host side (torchrun env, a file, a
 *                   PostgreSQL shared-memory slot ...) hands it to the other ranks
 *   comm_init       ncclCommInitRank on the device given to ndb_b200_init; collective over all ranks
 *   comm_shutdown   destroys the communicator (also done by ndb_b200_shutdown)
 * NCCL is loaded at the first comm_* call (dlopen libnccl.so.2; NDB_B200_NCCL_LIB overrides the name);
 * a process that never calls them does not need it. */
#define NDB_B200_COMM_ID_BYTES 128
int ndb_b200_comm_unique_id(void *id, size_t len);
int ndb_b200_comm_init(int rank, int world, const void *id, size_t len);
int ndb_b200_comm_shutdown(void);
int ndb_b200_comm_rank(void);
int ndb_b200_comm_nranks(void);
int ndb_b200_comm_nccl_version(void);
/* 1 once the sharded searches exchange through peer-memory windows (NVLink stores by our own kernel), 0 = ncclAllGather */
int ndb_b200_comm_exchange_is_p2p(void);
/* raw collectives on device buffers, queued on `stream` (NULL = the library stream); with one rank they
 * are a device copy / a no-op.  type: 0 = f32, 1 = i32, 2 = f64, 3 = i64. */
int ndb_b200_comm_allgather_dev(const void *send_dev, void *recv_dev, size_t bytes_per_rank, void *stream);
int ndb_b200_comm_allreduce_sum_dev(void *buf_dev, size_t count, int type, void *stream);
int ndb_b200_comm_broadcast_dev(void *buf_dev, size_t bytes, int root, void *stream);
/* Sharded search = distributed_knn_search + merge_distributed_results (distributed.c:56-487): every rank
 * answers the (replicated) query batch against the rows its handle holds -- whole lists
 * (ndb_b200_ivf_set_shard) or a stripe of every list (the caller inserts rows i with i % world == rank,
 * with their global ids) -- writes its top-k as packed records (4-byte distance block + 8-byte id block
 * = 12 bytes per result) into an exchange window; a push kernel stores them into every peer's window over NVLink
 * (CUDA IPC peer memory) and raises a flag -- or, where IPC is unavailable / NDB_B200_EXCHANGE=nccl, ONE
 * ncclAllGather moves them -- and every rank merges the world's lists on the device by (dist, id).  All ranks must make the same call; every rank receives the full result.
 * Without a communicator (or world == 1) these are the plain searches. */
int ndb_b200_ivf_search_sharded_dev(ndb_b200_ivf *ix, const float *Q_dev, int nq, int nprobe, int k, int mode, int arith,
                                    float *dist_dev, int64_t *ids_dev, void *stream);
int ndb_b200_ivf_search_sharded(ndb_b200_ivf *ix, const float *Q, int nq, int nprobe, int k, int mode, int arith,
                                float *dist, int64_t *ids);
int ndb_b200_knn_exact_sharded_dev(ndb_b200_dataset *ds, int metric, int arith, const float *Q_dev, int nq, int k,
                                   float *dist_dev, int64_t *ids_dev, void *stream);
/* Row-sharded kmeans_run (ivf_am.c:2117-2159): n_local rows per rank, centroids replicated.  C_dev in: the
 * initial centroids (kmeans_init: the first k rows of the GLOBAL order, broadcast by the caller), out: the
 * trained ones.  Per Lloyd iteration one all-reduce of the k*d f32 sums, one of the k counts, one of the
 * cost.  One rank: bit-identical to ndb_b200_kmeans_train; several: equal to fp32 rounding (the order in
 * which the ranks' partial sums are added is the reduction's). */
int ndb_b200_kmeans_train_sharded_dev(const float *X_dev, int64_t n_local, int d, int k, int max_iter, float tol,
                                      float *C_dev, int *assign_dev, int *counts_dev, int *iters, float *cost,
                                      void *stream);

/* ---- index key extraction: ivfExtractVectorData (ivf_am.c:117-218), hnswExtractVectorData -----
 * (hnsw_am.c:1402-1519).  The access methods accept vector, halfvec, sparsevec and bit columns and
 * turn every key into float4[dim] first; these do it for a batch of n keys, writing n rows of dim
 * floats (row-major) that dataset_append / ivf_insert / hnsw_build take as they are.
 *   vector    : n detoasted Vector datums laid end to end (struct Vector, neurondb.h:35-41: int32
 *               vl_len_, int16 dim, int16 unused, float4 data[dim] = 8 + 4*dim bytes each); a datum
 *               of another dimension -> NDB_B200_EDIM (check_dimensions, vector_distance.c:55-73)
 *   halfvec   : n*dim IEEE binary16 values (VectorF16.data, neurondb.h:44-49) through
 *               fp16_to_float (src/types/quantization.c:171-215): IEEE for zeros, normals, Inf and
 *               NaN; subnormal halves come out 2^-10 times their IEEE value, as in the reference
 *   bit       : n rows of ceil(nbits/8) bytes (VARBITS), bit i = MSB-first -> 1.0f, else -1.0f
 *   sparsevec : CSR batch of VectorMap entries (neurondb_types.h:47-53,106-107): row r holds
 *               indices/values [indptr[r], indptr[r+1]); zero fill, entries applied in order (a
 *               repeated index keeps the last value), indices outside [0,total_dim) ignored
 * dim / nbits / total_dim must be 1..32767 (the reference's check), else NDB_B200_EINVAL. */
int ndb_b200_keys_from_vector(const void *datums, int64_t n, int dim, float *rows);
int ndb_b200_keys_from_halfvec(const uint16_t *h, int64_t n, int dim, float *rows);
int ndb_b200_keys_from_halfvec_dev(const uint16_t *h_dev, int64_t n, int dim, float *rows_dev, void *stream);
int ndb_b200_keys_from_bits(const uint8_t *bits, int64_t n, int nbits, float *rows);
int ndb_b200_keys_from_bits_dev(const uint8_t *bits_dev, int64_t n, int nbits, float *rows_dev, void *stream);
int ndb_b200_keys_from_sparse(const int64_t *indptr, const int32_t *indices, const float *values, int64_t n,
                              int total_dim, float *rows);

/* ---- instrumentation --------------------------------------------------------------------- */
/* kernels launched by this library since init (bench.py's gpu_launches) */
int64_t ndb_b200_launch_count(void);
/* device time (ms, CUDA events on the library's stream) of the dominant kernel of the last
 * search call and the algorithmic bytes it streamed; 0 if timing is disabled */
int ndb_b200_set_timing(int enabled);
int ndb_b200_last_kernel_stats(double *ms, double *algo_bytes, int64_t *distance_evals);

#ifdef __cplusplus
}
#endif
#endif /* NDB_B200_H */
