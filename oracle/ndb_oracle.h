/*
 * ndb_oracle.h -- CPU oracle for the NeuronDB vector-search hot path.
 *
 * TEST INFRASTRUCTURE ONLY.  This is a plain-C restatement, over in-memory
 * arrays, of the arithmetic and control flow of the reference's CPU path
 * (NeuronDB/src/vector/vector_distance{,_simd}.c, src/index/ivf_am.c,
 * src/index/hnsw_am.c).  It exists so that tests/, __graft_entry__.smoke()
 * and bench.py's cpu_baseline / --impl reference legs can check and time the
 * reference algorithm.  Nothing under neurondb_b200/ may link or call it.
 *
 * Pinning status (see DESIGN.md "Oracle"):
 *   - operator distances (orc_l2_distance, orc_inner_product_*, orc_cosine_distance
 *     and the AVX restatements): PINNED against the 8 known answers of
 *     NeuronDB/t/005_distances_comprehensive.t and against the reference's own
 *     vector_distance.c / vector_distance_simd.c compiled under oracle/pgshim
 *     (oracle/_ref/libndb_ref_distance*.so) on seeded random inputs, bit for bit.
 *   - index arithmetic (orc_ivf_distance, orc_hnsw_distance, orc_hnsw_random_level,
 *     orc_kmeans_train / _assign / _update): PINNED against the reference's own
 *     ivfComputeDistance, hnswComputeDistance, hnswGetRandomLevel and k-means
 *     block, cut out of ivf_am.c / hnsw_am.c by oracle/extract_ref_leafs.py and
 *     compiled with the reference's flags (oracle/_ref/libndb_ref_leafs.so), and
 *     against tests/golden/index_leafs.npz generated from it.
 *   - key extraction (orc_fp16_to_float): PINNED the same way (libndb_ref_fp16.so).
 *   - knn_classify / knn_regress / cluster_kmeans (ndb_oracle_ml.c): PINNED against the reference's own
 *     euclidean_distance, compare_samples, kmeanspp_init, neurondb_l2_distance_squared and the text of
 *     cluster_kmeans' Lloyd loop (libndb_ref_leafs.so; golden tests/golden/ml_paths.npz).
 *   - the control flow around them (ivfSelectClusters, ivfCollectCandidates, the
 *     ivfinsert assignment loop, hnswSearch, hnswInsertNode): welded to the buffer
 *     manager, no results asserted by the reference's tests -- "parity unpinned";
 *     restated line by line from the cited source ranges.
 *
 * Compile with -O2 -ffp-contract=off (the reference's default build has no
 * -march flag, so no FMA contraction can occur: NeuronDB/build.sh:712).
 */
#ifndef NDB_ORACLE_H
#define NDB_ORACLE_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define ORC_INVALID 0xFFFFFFFFu        /* InvalidBlockNumber */
#define ORC_HNSW_MAX_LEVEL 16          /* hnsw_am.c:84 */
#define ORC_HNSW_MAX_M2 256            /* 2 * m upper bound of the scratch arrays (m <= 128, hnsw_am.c:79) */

/* metrics / strategies (sk_strategy numbering of the AMs: 1=L2, 2=cosine, 3=-IP) */
enum { ORC_L2 = 1, ORC_COSINE = 2, ORC_IP = 3 };

/* arithmetic variants of the same metric */
enum {
    ORC_ARITH_OP_F64 = 0,   /* <->,<=>,<#> scalar default build: fp64 (Kahan for L2) */
    ORC_ARITH_AVX2 = 1,     /* vector_distance_simd.c AVX2 bodies (8 f32 lanes)      */
    ORC_ARITH_AVX512 = 2,   /* AVX-512 bodies (16 f32 lanes)                          */
    ORC_ARITH_IVF_F32 = 3,  /* ivfComputeDistance: f32 sequential                     */
    ORC_ARITH_HNSW = 4      /* hnswComputeDistance: f32 op -> f64 accumulate          */
};

/* ---- distances ---------------------------------------------------------- */
int   orc_check_vector(const float *v, int dim);     /* index of first NaN/Inf, or -1 */
float orc_l2_distance(const float *a, const float *b, int dim);
float orc_inner_product_distance(const float *a, const float *b, int dim); /* -dot */
float orc_inner_product_op(const float *a, const float *b, int dim);       /* <#> = +dot (Q3) */
float orc_cosine_distance(const float *a, const float *b, int dim);
float orc_l2_avx(const float *a, const float *b, int dim, int lanes);
float orc_ip_avx(const float *a, const float *b, int dim, int lanes);
float orc_cosine_avx(const float *a, const float *b, int dim, int lanes);
float orc_ivf_distance(const float *v1, const float *v2, int dim, int strategy);
float orc_hnsw_distance(const float *v1, const float *v2, int dim, int strategy);
float orc_kmeans_l2sq(const float *v1, const float *v2, int dim);
/* generic dispatcher used by the batch drivers */
float orc_distance(const float *a, const float *b, int dim, int metric, int arith);
void  orc_distance_pairs(const float *A, const float *B, float *out, int64_t n, int dim,
                         int metric, int arith);

/* ---- exact kNN (seq-scan + top-N sort, ties by (dist,id)) --------------- */
void orc_knn_exact(const float *X, const int64_t *ids, int64_t n, int dim,
                   const float *Q, int nq, int k, int metric, int arith,
                   float *out_dist, int64_t *out_ids, int nthreads);

/* ---- IVF k-means (ivf_am.c:2070-2294) ----------------------------------- */
int  orc_kmeans_train(const float *X, int n, int dim, int k, int max_iter, float threshold,
                      float *C, int *assign, int *counts, float *cost_out);
void orc_kmeans_assign(const float *X, int64_t n, int dim, const float *C, int k,
                       int *assign, int nthreads);
void orc_kmeans_update(const float *X, const int *assign, int64_t n, int dim, int k,
                       float *C, int *counts);
int  orc_ivf_train_samples(int64_t nrows, int nlists);   /* min(10000, lists*100), ivf_am.c:580 */

/* ---- IVF insert-time assignment (ivf_am.c:906-935) ---------------------- */
void orc_ivf_assign(const float *X, int64_t n, int dim, const float *C, int nlists,
                    int *out_list, int nthreads);

/* ---- IVF search (ivf_am.c:1597-1717, 1722-1909) -------------------------
 * Lists are CSR over insertion order: list_off[nlists+1], list_rows[] = row index into X.
 * literal != 0: k*10 candidate cap and selection-sort tie behaviour (Q9).
 * literal == 0: every probed list is scanned fully, ties by (dist,id).          */
void orc_ivf_select_clusters(const float *q, int dim, const float *C, int nlists, int nprobe,
                             int *selected);
int  orc_ivf_search_one(const float *X, const int64_t *ids, int dim,
                        const float *C, int nlists, const int64_t *list_off, const int64_t *list_rows,
                        const float *q, int nprobe, int k, int strategy, int literal,
                        float *out_dist, int64_t *out_ids);
void orc_ivf_search(const float *X, const int64_t *ids, int dim,
                    const float *C, int nlists, const int64_t *list_off, const int64_t *list_rows,
                    const float *Q, int nq, int nprobe, int k, int strategy, int literal,
                    float *out_dist, int64_t *out_ids, int *out_count, int nthreads);

/* ---- HNSW (hnsw_am.c:1143-1161, 1545-2080, 2091-2670) ------------------- */
typedef struct OrcHnsw OrcHnsw;
OrcHnsw *orc_hnsw_create(int dim, int m, int ef_construction, int ef_search, float ml, int64_t capacity);
void     orc_hnsw_free(OrcHnsw *g);
int      orc_hnsw_random_level(float ml);            /* libc random(), hnsw_am.c:1143-1161 */
/* mode: 0 = literal hnswInsertNode semantics, 1 = per-level search without self hits */
void     orc_hnsw_insert(OrcHnsw *g, const float *vec, int level, int mode);
void     orc_hnsw_build(OrcHnsw *g, const float *X, int64_t n, const int *levels, int mode);
/* search_mode: 0 = literal hnswSearch (BFS until ef candidates, Q12), 1 = best-first */
int      orc_hnsw_search_one(const OrcHnsw *g, const float *q, int strategy, int ef, int k,
                             int search_mode, uint32_t *out_nodes, float *out_dist);
void     orc_hnsw_search(const OrcHnsw *g, const float *Q, int nq, int strategy, int ef, int k,
                         int search_mode, uint32_t *out_nodes, float *out_dist, int *out_count,
                         int nthreads);
/* graph export (flat arrays, the same layout ndb_b200_hnsw_load_graph takes) */
int64_t  orc_hnsw_size(const OrcHnsw *g);
void     orc_hnsw_meta(const OrcHnsw *g, uint32_t *entry_point, int *entry_level, int *max_level);
void     orc_hnsw_export(const OrcHnsw *g, int *levels /*n*/, uint32_t *nbr0 /*n*2m*/,
                         int16_t *cnt /*n*16*/, int64_t *upper_off /*n+1*/, uint32_t *upper /*sum(level)*2m*/);
int64_t  orc_hnsw_upper_slots(const OrcHnsw *g);    /* sum over nodes of level * 2m */
int64_t  orc_hnsw_distance_evals(void);             /* counter of the last search batch */

/* ---- relation images in the reference's on-disk page layout (ndb_oracle_pages.c) ------- */
int64_t orc_ivf_encode_relation(const float *X, const int64_t *tids, int64_t n, int dim,
                                const float *C, int nlists, int nprobe, const int *assign,
                                uint8_t *blocks, int64_t cap_blocks, int multi_page_centroids);
void    orc_page_mark_dead(uint8_t *blocks, int64_t block, int offnum);
int64_t orc_hnsw_encode_relation(const OrcHnsw *g, const float *X, const int64_t *tids, int dim, int m, int efc,
                                 int efs, uint8_t *blocks, int64_t cap_blocks);

/* ---- recall@k (ml_recall_metrics.c:65-126) ------------------------------ */
double orc_recall_at_k(const int64_t *found, const int64_t *truth, int nq, int k);

/* ---- (dist,id) merge of per-shard top-k (distributed.c:425-438) ---------- */
void orc_merge_topk(const float *dist, const int64_t *ids, int nshards, int nq, int k,
                    float *out_dist, int64_t *out_ids);

/* ---- index key extraction (ivf_am.c:117-218, hnsw_am.c:1402-1519) -------- */
float orc_fp16_to_float(uint16_t h);                 /* src/types/quantization.c:171-215 */
void orc_keys_from_halfvec(const uint16_t *h, int64_t n, int dim, float *rows);
void orc_keys_from_bits(const uint8_t *bits, int64_t n, int nbits, float *rows);
void orc_keys_from_sparse(const int64_t *indptr, const int32_t *indices, const float *values,
                          int64_t n, int total_dim, float *rows);

/* ---- SQL functions on the same kernels (ndb_oracle_ml.c; ml_knn.c, ml_kmeans.c) ---- */
double orc_ml_euclidean(const float *a, const float *b, int dim);            /* ml_knn.c:76-90 */
void orc_knn_ml(const float *X, const double *labels, int n, int dim, const float *q, int k, int *cls, double *mean,
                int *rows_out);                                               /* ml_knn.c:264-333, 504-557 */
double orc_l2_distance_squared(const float *a, const float *b, int n);       /* neurondb_simd_impl.c:36-104 */
int orc_kmeanspp_init(const float *X, int nvec, int dim, int k, const int *draws, int rand_max, int *centroids);
int orc_cluster_kmeans(const float *X, int nvec, int dim, int k, int max_iters, const int *draws, int rand_max,
                       int *labels, float *centers_out, int *seeds_out);     /* ml_kmeans.c:45-139, 146-303 */

int orc_cluster_minibatch_kmeans(const float *X, int nvec, int dim, int k, int batch_size, int max_iters, const int *draws, int ndraws,
                                 int rand_max, int *consumed, int *labels, float *centers_out);   /* ml_minibatch_kmeans.c:67-198, 206-449 */
/* product quantisation (ml_product_quantization.c:80-190, 195-415, 421-536, 1003-1110); codebooks [m][ksub][dsub] */
void orc_pq_train_subspace(const float *S, int nvec, int dsub, int k, const int *draws, float *centroids, int max_iters);
int orc_pq_train(const float *X, int nvec, int dim, int m, int ksub, const int *draws, int max_iters, float *codebooks);
void orc_pq_encode(const float *X, int64_t n, int dim, const float *codebooks, int m, int ksub, int16_t *codes);
float orc_pq_asymmetric_distance(const float *q, const int16_t *codes, const float *codebooks, int m, int ksub, int dsub);
void orc_pq_knn(const float *Q, int nq, const int16_t *codes, int64_t n, const float *codebooks, int dim, int m, int ksub, int k,
                float *dist, int64_t *rows, float *dist_all);

/* per-vector quantisers and the Hamming scan (src/types/quantization.c); kinds 1 int8, 2 fp16, 3 binary, 4 uint8, 5 ternary, 6 int4 */
int64_t orc_quantized_row_bytes(int kind, int dim);
void orc_quantize_row(int kind, const float *v, int dim, uint8_t *out);
void orc_quantize_rows(int kind, const float *X, int64_t n, int dim, uint8_t *out);
int orc_hamming(const uint8_t *a, const uint8_t *b, int nbits);
void orc_hamming_knn(const uint8_t *rows, int64_t n, int nbits, const uint8_t *Q, int nq, int k, int32_t *dist, int64_t *ids);

/* sizes / field offsets used by the relation encoders (ndb_oracle_pages.c), in the order of the
 * reference-side ref_layout() built by oracle/extract_ref_leafs.py; returns the count */
int orc_page_layout(int64_t *out);

#ifdef __cplusplus
}
#endif
#endif
