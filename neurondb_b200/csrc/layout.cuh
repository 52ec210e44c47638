// layout.cuh -- cross-TU declarations: IL32 layout kernels, scan/merge launchers.
#pragma once
#include "common.cuh"

namespace ndb {

struct ScanParams;

// paired rows (core.cu)
int launch_pairs(int metric, int arith, const float *dA, const float *dB, float *dOut, int64_t n, int dim,
                 int64_t b_stride, cudaStream_t s);

// ---- layout.cu ----------------------------------------------------------------------------
// rows_dev [n][dim] row-major -> IL32 store; row i goes to slot slot_of_row[i] (or slot_base+i).
int il32_scatter(const float *rows_dev, int64_t n, int dim, int dimp, const uint32_t *slot_of_row,
                 uint32_t slot_base, float *store, cudaStream_t s);
// move vectors between IL32 stores: dst slot dst_slot[i] <- src slot src_slot[i]
int il32_gather(const float *src_store, const uint32_t *src_slot, float *dst_store, const uint32_t *dst_slot,
                int64_t n, int dimp, cudaStream_t s);
// IL32 store -> row-major rows (for export / tests)
int il32_to_rows(const float *store, int64_t n, int dim, int dimp, float *rows_dev, cudaStream_t s);
// per-vector norm accumulators in the arithmetic `arith` uses for cosine (float or double)
size_t norm_elem_size(int arith);
int slot_norms(int arith, const float *store, int64_t nslots, int dim, int dimp, void *out, cudaStream_t s);
int row_norms(int arith, const float *rows_dev, int64_t n, int dim, void *out, cudaStream_t s);

// ---- scan_launch.cu -----------------------------------------------------------------------
// queries per tile the scan kernel will use for (arith, dim, k); 0 if the shape is unsupported
int scan_pick_qt(int arith, int dim, int k);
int launch_scan(int metric, int arith, int qt, const ScanParams &prm, uint32_t items_upper, cudaStream_t s);
// merge [nq][nparts][k] (dist, slot) partials -> (dist, id) ; ids == nullptr => id = slot
int launch_merge_parts(const float *pdist, const uint32_t *pslot, const int64_t *ids, int nq, int nparts, int k,
                       float *out_dist, int64_t *out_ids, uint32_t *out_slot, cudaStream_t s);
int launch_merge_shards(const float *dist, const int64_t *ids, int nshards, int nq, int k, float *out_dist,
                        int64_t *out_ids, cudaStream_t s);

// scratch shared by the search entry points of one handle
struct ScanScratch {
    DevBuf pdist, pslot, counter, qnorm;
    int ensure(size_t partials, int k, size_t nq, size_t norm_bytes)
    {
        NDB_CHECK(pdist.reserve(partials * k * sizeof(float)));
        NDB_CHECK(pslot.reserve(partials * k * sizeof(uint32_t)));
        NDB_CHECK(counter.reserve(64));
        if (norm_bytes) NDB_CHECK(qnorm.reserve(nq * norm_bytes));
        return NDB_B200_OK;
    }
};

}  // namespace ndb
