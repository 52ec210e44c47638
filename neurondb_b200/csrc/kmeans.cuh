// kmeans.cuh -- shared declarations for the k-means / assignment kernels (kmeans.cu).
#pragma once
#include "layout.cuh"
#include "tc.cuh"

namespace ndb {

// nearest centroids on the tensor cores, certified (cert_common.cuh): blocked bf16 copy of the centroids, scratch,
// work items cached per batch shape, the queries sent to the exact kernel and the counters [0] of them, [1] exact evaluations
struct CentroidSearch {
    TcStore store;
    TcScratch scr;
    DevBuf fb_list, counters, probe, cdist;
    bool store_ok = false;
    int items_nq = -1;
    uint32_t items_tpr = 0, items_nranges = 0;
    int64_t items_ntiles = -1;
};

struct KMeansWork {
    DevBuf X, C, cstore, assign, keys_sorted, vals_in, vals_sorted, start, counts, cub_tmp, dcost, cost;
    ScanScratch scr;
    CentroidSearch cs;
};

// probe[nq][np] / cdist[nq][np] = the np nearest of the L centroids (C row-major, cstore its IL32 copy) of every row of
// Q_dev, ordered by (distance, index) in the reference's fp32 arithmetic: sqrtf'd L2 (ivfSelectClusters :1597-1717,
// ivfinsert :906-935) or, with `squared`, k-means' squared L2 (:2255-2294).  np <= 32.  The centroid store is rebuilt
// when cs.store_ok is false.  counters (2 x u64, device) accumulate [0] exact-kernel queries, [1] exact evaluations.
int nearest_centroids_tensor(CentroidSearch &cs, const float *C_rowmajor, const float *cstore_il32, int L, int dim, int dimp,
                             const float *Q_dev, int nq, int np, bool squared, uint32_t *probe, float *cdist,
                             unsigned long long *counters, cudaStream_t s);

void kmeans_at_shutdown();

int dense_scan(const float *store, const void *vnorm, int64_t nvec, int dim, int dimp, int metric, int arith,
               const float *Q_dev, int nq, int k, ScanScratch &scr, int *out_nparts, cudaStream_t s);
// w.C must hold the k*dim row-major centroids on the device
int kmeans_assign_dev(KMeansWork &w, const float *dX, int64_t n, int dim, int k, int metric, int *d_assign,
                      cudaStream_t s);
int group_by_cluster_dev(KMeansWork &w, const int *d_assign, int64_t n, int k, cudaStream_t s);
// sums_only: leave the per-cluster sums instead of the means (row-sharded training)
int kmeans_update_dev(KMeansWork &w, const float *dX, const int *d_assign, int64_t n, int dim, int k, float *dC,
                      int *d_counts, cudaStream_t s, bool sums_only = false);

int kmeans_run_dev(KMeansWork &w, int n, int d, int k, int max_iter, float tol, int *iters, float *cost_out,
                   cudaStream_t s);

// ---- fp64 Lloyd (cluster_kmeans.cu): the arithmetic of cluster_kmeans (ml_kmeans.c:226-278) and train_subspace_kmeans
// (ml_product_quantization.c:108-186): double difference / square / sum, strict <, lowest index wins; ordered float update.
// dXT = the rows transposed ([dim][n], row stride n).  w.C holds the k*dim centres, w.assign the current assignment.
int upload_rows_checked(DevBuf &dst, const float *X, size_t count, const char *who, cudaStream_t s);
int transpose_rows_dev(const float *dX, int64_t n, int dim, float *dXT, cudaStream_t s);
int gather_rows_dev(const float *dX, const int *rows_dev, int nrows, int dim, float *out, cudaStream_t s);
// (wide: scratch for the k*dim centres widened to double)
int nearest_f64_dev(const float *dXT, const float *dC, int64_t n, int dim, int k, int *assign, int *changed, DevBuf &wide, cudaStream_t s);
// stop_before_update: leave the loop after an assignment pass that changed nothing (train_subspace_kmeans); otherwise the
// update runs once more and the loop condition ends it (cluster_kmeans)
int lloyd_f64_dev(KMeansWork &w, const float *dX, const float *dXT, int64_t n, int dim, int k, int max_iters, bool stop_before_update,
                  int *dchanged, int *iters, cudaStream_t s);

}  // namespace ndb
