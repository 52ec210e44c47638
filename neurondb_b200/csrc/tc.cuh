// tc.cuh -- host-side handles of the tensor-core (tcgen05) distance path (tc_knn.cu).
#pragma once
#include "layout.cuh"
#include <cuda_bf16.h>

namespace ndb {

constexpr int TC_M = 128;          // queries per tile  (TMEM lanes)
constexpr int TC_N = 256;          // stored rows per tile (TMEM columns per accumulator)
constexpr int TC_KC = 128;         // dims per K-chunk (one smem stage)
constexpr int TC_MAX_CHUNKS = 2;   // K-chunks of a query tile that stay resident in shared memory (dim <= 256)
constexpr int TC_MAX_DIM = 2048;   // beyond TC_MAX_CHUNKS chunks the query tile is streamed with the stored tiles
constexpr int TC_KMAX = 16;        // entries of a thread-local register list (dense kNN: k <= 16)
constexpr int TC_IVF_KMAX = 32;    // the certified IVF search proposes from 16-entry partial lists and re-evaluates up to 64
                                   // candidates per query: k <= 32
constexpr int TC_PACKED_MAX_TILES = 16;   // packed-key epilogue: 11 index bits = 16 tiles x 128 columns per half
constexpr int TC_IDX_BITS = 11;           // low mantissa bits of a packed key that carry the candidate's index within its item
constexpr uint32_t TC_IDX_MASK = (1u << TC_IDX_BITS) - 1;

// stored rows in the blocked bf16 layout + squared norms of the rounded rows
struct TcStore {
    DevBuf xb, xnorm, xrinv, xpad0; // blocked rows, squared norms of the rounded rows, 1 / norm (cosine), 0 | +inf (inner product)
    DevBuf stats;                   // 4 floats: rounding-error maxima over the stored rows (tc_row_norms_kernel)
    int64_t valid_for = -1, rinv_for = -2, pad0_for = -2;
    int64_t ntiles = 0;
    int nkc = 0;
};

// one unit of work of tc_knn_kernel: query tile `qtile` (128 queries) against stored tiles [t0, t1)
// (256 rows each); lane ql / column-half h writes its k results at partial slot
// out_base + ql * out_stride + h, for ql < nq
struct TcItem {
    uint32_t qtile, t0, t1, nq, out_base, out_stride;
    uint32_t rep;       // 4: every query sits on four consecutive tile positions and replica r scans only the r-th
                        // 32-column chunk of its column half (sparsely probed IVF lists); else 1
    uint32_t nrows;     // stored rows in [t0, t1): the columns beyond them (the last tile's padding) never rank
};

struct TcScratch {
    DevBuf qb, qnorm, qerr, pdist, pslot, debug, items, gthr;
};

struct TcParams {
    const __nv_bfloat16 *xb;       // blocked stored rows
    const float *xnorm;            // [ntiles * 256]
    const __nv_bfloat16 *qb;       // blocked query tiles
    const float *qnorm;            // [query tiles * 128]
    int nkc;                       // K-chunks (dimp / 128)
    int k;
    const TcItem *items;           // work items
    uint32_t nitems;
    const uint32_t *nitems_ptr;    // optional: the item count lives on the device (nitems is then only an upper bound)
    const uint32_t *item_lo_ptr;   // optional (device): this launch covers the items [*item_lo_ptr, *item_hi_ptr) only;
    const uint32_t *item_hi_ptr;   // null = from the first / to the last (two-phase IVF scans)
    float *pdist;                  // partial results, indexed through TcItem::out_base / out_stride
    uint32_t *pslot;
    const uint32_t *qmap;          // optional (list mode): tile position -> (query * nprobe + rank), INVALID_SLOT = empty
    uint32_t nprobe;
    float *gthr;                   // optional (list mode): per query, an upper bound of its k-th best candidate, shared
                                   // between the items of the query (atomic min; initialised to a huge finite value)
    const float *qerr;             // optional: squared norm of each tile position's bf16 rounding error (with cstats)
    const float *cstats;           // optional: TcStore::stats of the stored rows -> the shared bound is published RELAXED
                                   // (cert_bound.cuh), which is what makes the selection certifiable
    int dim;                       // row dimension (for the accumulation slack, and for kg_last)
    int nstages, xstride;          // set by tc_launch: depth of the stored-row ring and the byte stride of its stages
    int heavy;                     // set by tc_launch: takers per 32 columns above which a warp sorts instead of inserting
    int kg_last;                   // set by tc_launch: 8-element K groups of the last chunk that are copied and multiplied
    int packed;                    // 1: items span <= TC_PACKED_MAX_TILES tiles; (distance | index) keys, sorting-network epilogue
    float *debug_d;                // optional: raw accumulator of the first tile [128][256]
    unsigned long long *dbg_counters; // builds with -DNDB_TC_COUNTERS only: epilogue statistics (chunks, chunks with a taker, sorted chunks, insert rounds, takers)
    int kpub;                      // list mode: the shared bound is published from the kpub-th key of a full list (0: the last,
                                   // k-th; < 0: never -- k exceeds the list length, only ivf_tc_bound_kernel sets the bound)
    int debug_mode;                // NDB_TC_DEBUG: 1 = no epilogue, 2 = no MMA issue, 4 = no X bulk copies, 8 = epilogue reads TMEM only (bisection aid)
};

int tc_launch(const TcParams &p, int metric, int k, cudaStream_t s);
int tc_store_rinv(TcStore &st, const float **out, cudaStream_t s);
int tc_store_pad0(TcStore &st, const float **out, cudaStream_t s);
int tc_build_store_mapped(TcStore &st, const float *il32_store, const uint32_t *src_slot_dev, int64_t n, int dim, int dimp,
                          cudaStream_t s);
// npos_dev (optional): device count of tile positions actually in use (nqpad is then an upper bound)
// negate: the tiles hold -q (inner product: the accumulator is then the candidate -x.q itself)
int tc_block_queries(const float *Q_dev, const uint32_t *qmap_dev, uint32_t nprobe, int nq, int nqpad, int dim, int nkc,
                     __nv_bfloat16 *qb, float *qnorm, cudaStream_t s, const uint32_t *npos_dev = nullptr, float *qerr = nullptr,
                     bool negate = false);

int tc_build_store(TcStore &st, const float *il32_store, int64_t n, int dim, int dimp, cudaStream_t s);
int tc_knn(const TcStore &st, TcScratch &sc, int dim, int metric, const float *Q_dev, int nq, int k, const int64_t *ids,
           float *dist_dev, int64_t *ids_dev, uint32_t *slots_dev, float *debug_d_dev, bool packed, cudaStream_t s);

}  // namespace ndb
