// kmeans.cuh -- shared declarations for the k-means / assignment kernels (kmeans.cu).
#pragma once
#include "layout.cuh"

namespace ndb {

struct KMeansWork {
    DevBuf X, C, cstore, assign, keys_sorted, vals_in, vals_sorted, start, counts, cub_tmp, dcost, cost;
    ScanScratch scr;
};

void kmeans_at_shutdown();

int dense_scan(const float *store, const void *vnorm, int64_t nvec, int dim, int dimp, int metric, int arith,
               const float *Q_dev, int nq, int k, ScanScratch &scr, int *out_nparts, cudaStream_t s);
// w.C must hold the k*dim row-major centroids on the device
int kmeans_assign_dev(KMeansWork &w, const float *dX, int64_t n, int dim, int k, int metric, int *d_assign,
                      cudaStream_t s);
// sums_only: leave the per-cluster sums instead of the means (row-sharded training)
int kmeans_update_dev(KMeansWork &w, const float *dX, const int *d_assign, int64_t n, int dim, int k, float *dC,
                      int *d_counts, cudaStream_t s, bool sums_only = false);

int kmeans_run_dev(KMeansWork &w, int n, int d, int k, int max_iter, float tol, int *iters, float *cost_out,
                   cudaStream_t s);

}  // namespace ndb
