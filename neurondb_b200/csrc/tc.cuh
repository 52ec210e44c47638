// tc.cuh -- host-side handles of the tensor-core (tcgen05) distance path (tc_knn.cu).
#pragma once
#include "layout.cuh"

namespace ndb {

// stored rows in the blocked bf16 layout + squared norms of the rounded rows
struct TcStore {
    DevBuf xb, xnorm;
    int64_t valid_for = -1;
    int64_t ntiles = 0;
    int nkc = 0;
};

// one unit of work of tc_knn_kernel: query tile `qtile` (128 queries) against stored tiles [t0, t1)
// (256 rows each); lane ql / column-half h writes its k results at partial slot
// out_base + ql * out_stride + h, for ql < nq
struct TcItem {
    uint32_t qtile, t0, t1, nq, out_base, out_stride;
};

struct TcScratch {
    DevBuf qb, qnorm, pdist, pslot, debug, items;
};

struct TcParams;
int tc_launch(const TcParams &p, int metric, int k, cudaStream_t s);

int tc_build_store(TcStore &st, const float *il32_store, int64_t n, int dim, int dimp, cudaStream_t s);
int tc_knn(const TcStore &st, TcScratch &sc, int dim, int metric, const float *Q_dev, int nq, int k, const int64_t *ids,
           float *dist_dev, int64_t *ids_dev, float *debug_d_dev, cudaStream_t s);

}  // namespace ndb
