// ivf.cu -- IVF index: ivfbuild (train), ivfinsert (assign + append), ivfrescan/ivfgettuple
// (ivfSelectClusters + ivfCollectCandidates) of NeuronDB/src/index/ivf_am.c, for query batches.
//
// Device layout.  Centroids: one IL32 store of nlists vectors (replicated on every rank).
// Inverted lists: one IL32 store; list l occupies the 32-vector blocks
// [list_blk[l], list_blk[l] + ceil(len/32)), entries sorted by id inside a list so that the
// (dist, slot) order the scan keeps equals the (dist, id) order inside one list.  ids[] maps
// slot -> heap id; lit_order[] maps (list, j-th inserted) -> slot for the literal mode, whose
// candidate cut-off depends on insertion (page/offset) order.
//
// Search = 4 launches on one stream, no host round trip:
//   1. dense scan queries x centroids, k = nprobe, always L2          (ivfSelectClusters :1597-1717)
//   2. bucket the (query, probed list) pairs by list -> work items     (build_items_*)
//   3. list-mode scan: each list block is fetched once per 8 queries   (ivfCollectCandidates :1722-1909)
//   4. merge the nprobe partial top-k lists per query by (dist, id)
#include "kmeans.cuh"
#include "scan.cuh"
#include "pages.cuh"

#include <algorithm>
#include <numeric>
#include <cub/block/block_scan.cuh>
#include <cub/device/device_radix_sort.cuh>
#include "tc.cuh"
#include "comm.cuh"

using namespace ndb;

struct ndb_b200_ivf {
    int dim = 0, dimp = 0, nlists = 0, metric = NDB_L2;
    bool trained = false;
    int rank = 0, world = 1;
    std::vector<float> C_host;
    KMeansWork kw;                       // kw.C = centroids row-major on device, kw.cstore = IL32
    // arena: rows kept by this shard, row-major, insertion order
    DevBuf arena;
    int64_t nrows = 0;
    int64_t inserted_total = 0;          // next default id: one past the largest id seen / handed out, over ALL rows passed in
    std::vector<int64_t> row_id;
    std::vector<int32_t> row_list;
    // laid-out lists
    bool dirty = true;
    VecStore store;
    DevBuf ids, vnorm_ivf, vnorm_fast, lit_order, d_list_len, d_list_blk, d_list_order;
    bool vnorm_ivf_ok = false, vnorm_fast_ok = false;
    std::vector<uint32_t> list_len, list_blk;
    // scratch
    ScanScratch cscr, scr;
    DevBuf probe, cnt, fill, qoff, item_off, qmap, pairpos, items, nitems, stats, tmp_rows, tmp_assign, tmp_keep;
    DevBuf qbuf, outd, outi, cdist;
    int64_t last_scanned = 0;
    // pipelined searches (search_begin / search_end): two batches in flight, each with its own staging
    struct Slot {
        DevBuf q, d, i, bad;
        cudaEvent_t h2d = nullptr, done = nullptr, d2h = nullptr;
        unsigned long long *h_bad = nullptr;      // pinned
        bool busy = false;
    } slot[2];
    ~ndb_b200_ivf()
    {
        for (Slot &sl : slot) {
            if (sl.h2d) cudaEventDestroy(sl.h2d);
            if (sl.done) cudaEventDestroy(sl.done);
            if (sl.d2h) cudaEventDestroy(sl.d2h);
            if (sl.h_bad) cudaFreeHost(sl.h_bad);
        }
    }
    // tensor-core copy (NDB_ARITH_TENSOR): every list padded to whole 256-row tiles of blocked bf16
    bool tc_ok = false;
    TcStore tc;                          // lists (the centroids' copy lives in kw.cs)
    TcScratch tcs;
    uint32_t tc_max_nseg = 1;            // longest list, in segments of the tensor scan
    uint64_t tc_sum_nseg = 0, tc_nonempty = 0;   // over the non-empty lists
    DevBuf tc_src, tc_row, d_ltile8;     // tensor row -> IL32 slot / arena row; first tile of each list (* 8, in 32-row blocks)
    // certified selection (ivf_cert.cuh): queries sent to the exact kernels, [0] lists [1] exact evaluations (lists)
    // [2] coarse [3] exact evaluations (coarse); items of the coarse scan, cached per batch shape
    DevBuf cert_counters, fb_list, fb_tau, cert_dbg;
    bool coarse_by_rank = false;         // set by the sharded entry point for the duration of one call: the coarse quantiser is
                                         // split by queries over the communicator's ranks and the probe lists all-gathered
    DevBuf d_start, d_sorted_list, d_row_of_slot;   // layout by-products kept for the tensor maps: first position of every list in
                                                    // the (list, id) order, the lists in that order, IL32 slot -> arena row
    uint64_t tc_tiles = 0;               // 256-row tiles of all lists
    double tc_rows_per_pair = 0.0;       // expected rows a (query, probed list) pair scans: the size-biased mean list length
};

namespace ndb {

// Long lists are cut into runs of `segb` 32-vector blocks so that one work item is bounded, and
// the items are emitted longest-list-first (`order` = lists by descending length, fixed at
// layout time): the persistent CTAs pick the heavy items up first and the short ones fill the
// tail, whatever the list-length skew.
__host__ __device__ inline uint32_t ivf_nseg(uint32_t len, uint32_t segb)
{
    const uint32_t blocks = (len + 31) / 32;
    return (blocks + segb - 1) / segb;
}
static uint32_t ivf_seg_blocks()
{
    static const uint32_t v = [] { const char *e = getenv("NDB_IVF_SEG_BLOCKS"); int x = e ? atoi(e) : 64; return (uint32_t) (x >= 1 ? x : 64); }();
    return v;
}

// Tensor path: a list probed by at most rep_max queries of the batch gets every query FOUR times, on
// consecutive tile positions (= the same lane of the four TMEM lane quarters, see
// tc_block_queries_kernel), and each replica's epilogue thread scans a quarter of its column half.
// Most tile-steps of a batch belong to long lists probed by a handful of queries; this cuts the
// serial epilogue work per tile of exactly those from 128 columns per thread to 32.
__host__ __device__ inline uint32_t ivf_rep(uint32_t c, uint32_t rep_max) { return c <= rep_max ? 4u : 1u; }

// Two-phase scans (tensor path, split = nlists): the pairs (query, its NEAREST list) are bucketed apart from the others, as
// "virtual list" l, the remaining pairs of list l as virtual list nlists + l.  The nearest lists are scanned first, their
// partial lists give every query a bound close to its final k-th distance, and the bulk of the work -- the other
// nprobe - 1 lists -- runs against that bound from its first tile on.  split = 0: one bucket per list.
__host__ __device__ inline uint32_t ivf_vlist(uint32_t l, int64_t pair, uint32_t nprobe, uint32_t split)
{
    return (split && pair % nprobe != 0) ? l + split : l;
}


}  // namespace ndb
#include "ivf_cert.cuh"
namespace ndb {

// ---- work-item construction -----------------------------------------------------------------
__global__ void ivf_hist_kernel(const uint32_t *__restrict__ probe, int64_t npairs, const uint32_t *__restrict__ list_len,
                                int nlists, uint32_t *__restrict__ cnt, uint32_t nprobe, uint32_t split)
{
    const int64_t p = (int64_t) blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= npairs) return;
    const uint32_t l = probe[p];
    if (l < (uint32_t) nlists && list_len[l] > 0) atomicAdd(&cnt[ivf_vlist(l, p, nprobe, split)], 1u);
}

// single CTA: exclusive scans, in `order`, of cnt (-> qoff) and of tiles * segments (-> item_off).
// split != 0: 2 * nlists virtual lists, phase 1 (the queries' nearest lists, in `order`) numbered before phase 2;
// nitems[2] = the number of phase-1 items.  scanned[0] = rows x queries of the whole batch, scanned[1] = of phase 2 alone.
__global__ void __launch_bounds__(1024) ivf_offsets_kernel(const uint32_t *__restrict__ cnt, const uint32_t *__restrict__ list_len,
                                                           const uint32_t *__restrict__ order, int nlists, int qt, uint32_t segb,
                                                           uint32_t qalign, uint32_t rep_max, uint32_t *__restrict__ qoff,
                                                           uint32_t *__restrict__ item_off, uint32_t *__restrict__ nitems,
                                                           unsigned long long *__restrict__ scanned, uint32_t split)
{
    typedef cub::BlockScan<uint32_t, 1024> Scan;
    __shared__ typename Scan::TempStorage tmp;
    __shared__ unsigned long long s_scanned, s_first;
    const int nv = split ? 2 * nlists : nlists;
    const int ipt = (nv + 1023) / 1024;
    const int b = threadIdx.x * ipt, e = min(nv, b + ipt);
    if (threadIdx.x == 0) { s_scanned = 0; s_first = 0; nitems[2] = 0; }
    uint32_t sq = 0, st = 0;
    unsigned long long sc = 0, sf = 0;
    for (int i = b; i < e; i++) {
        const uint32_t l = order[i < nlists ? i : i - nlists], v = i < nlists ? l : l + split;
        const uint32_t c = cnt[v], cr = c * ivf_rep(c, rep_max);
        sq += (cr + qalign - 1) / qalign * qalign;
        st += ((cr + qt - 1) / qt) * ivf_nseg(list_len[l], segb);
        sc += (unsigned long long) c * list_len[l];
        if (split && i < nlists) sf += (unsigned long long) c * list_len[l];
    }
    uint32_t oq, ot;
    Scan(tmp).ExclusiveSum(sq, oq);
    __syncthreads();
    Scan(tmp).ExclusiveSum(st, ot);
    __syncthreads();
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) { sc += __shfl_xor_sync(FULL, sc, o); sf += __shfl_xor_sync(FULL, sf, o); }
    if ((threadIdx.x & 31) == 0 && sc) atomicAdd(&s_scanned, sc);
    if ((threadIdx.x & 31) == 0 && sf) atomicAdd(&s_first, sf);
    for (int i = b; i < e; i++) {
        const uint32_t l = order[i < nlists ? i : i - nlists], v = i < nlists ? l : l + split;
        const uint32_t c = cnt[v], cr = c * ivf_rep(c, rep_max);
        if (split && i == nlists) nitems[2] = ot;               // everything before: phase 1
        qoff[v] = oq;
        item_off[v] = ot;
        oq += (cr + qalign - 1) / qalign * qalign;
        ot += ((cr + qt - 1) / qt) * ivf_nseg(list_len[l], segb);
    }
    if (threadIdx.x == 1023) { nitems[0] = ot; nitems[1] = oq; }   // last thread's running totals == grand totals
    __syncthreads();
    if (threadIdx.x == 0) { scanned[0] = s_scanned; scanned[1] = s_scanned - s_first; }
}

__global__ void ivf_scatter_kernel(const uint32_t *__restrict__ probe, int64_t npairs, const uint32_t *__restrict__ list_len,
                                   int nlists, const uint32_t *__restrict__ qoff, uint32_t *__restrict__ fill,
                                   uint32_t *__restrict__ qmap, uint32_t *__restrict__ pairpos,
                                   const uint32_t *__restrict__ cnt, uint32_t rep_max, uint32_t nprobe, uint32_t split)
{
    const int64_t p = (int64_t) blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= npairs) return;
    const uint32_t l = probe[p];
    if (l < (uint32_t) nlists && list_len[l] > 0) {
        const uint32_t v = ivf_vlist(l, p, nprobe, split);
        const uint32_t pos = atomicAdd(&fill[v], 1u);      // position of this query among the (virtual) list's queries
        const uint32_t r = cnt ? ivf_rep(cnt[v], rep_max) : 1u;
        for (uint32_t j = 0; j < r; j++) qmap[qoff[v] + pos * r + j] = (uint32_t) p;
        pairpos[p] = pos;
    }
}

__global__ void ivf_items_kernel(const uint32_t *__restrict__ cnt, const uint32_t *__restrict__ qoff,
                                 const uint32_t *__restrict__ item_off, const uint32_t *__restrict__ list_len,
                                 const uint32_t *__restrict__ list_blk, int nlists, int qt, uint32_t segb,
                                 WorkItem *__restrict__ items)
{
    // one warp per list, lanes over its (tile, segment) items
    const int l = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (l >= nlists) return;
    const uint32_t c = cnt[l];
    const uint32_t tiles = (c + qt - 1) / qt;
    const uint32_t len = list_len[l], nseg = ivf_nseg(len, segb);
    // item (tile t, segment sg) sits at item_off[l] + t * nseg + sg
    for (uint32_t i = threadIdx.x & 31; i < tiles * nseg; i += 32) {
        const uint32_t t = i / nseg, sg = i - t * nseg;
        WorkItem it;
        it.blk_begin = list_blk[l] + sg * segb;
        it.nvec = min(segb * 32, len - sg * segb * 32);
        it.qoff = qoff[l] + t * qt;
        it.nq = min((uint32_t) qt, c - t * qt);
        it.part = sg;
        items[item_off[l] + i] = it;
    }
}

// merge, per query, the partial top-k lists written by the list-mode scan: for each probed list
// the query sits at position pairpos[p] among that list's queries, i.e. in tile pos / tile and
// slot pos % tile of the items item_off[l] + (pos / tile) * nseg + sg.  (dist, id) order.
template <int KR>
__global__ void ivf_merge_kernel(const float *__restrict__ pdist, const uint32_t *__restrict__ pslot,
                                 const int64_t *__restrict__ ids, const uint32_t *__restrict__ probe,
                                 const uint32_t *__restrict__ pairpos, const uint32_t *__restrict__ item_off,
                                 const uint32_t *__restrict__ list_len, int nq, int nprobe, int nlists, int tile, uint32_t segb, int k,
                                 float *__restrict__ out_dist, int64_t *__restrict__ out_ids)
{
    const int lane = threadIdx.x & 31;
    const int q = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (q >= nq) return;
    WarpTopK<KR, int64_t> top;
    top.init();
    for (int r = 0; r < nprobe; r++) {
        const size_t p = (size_t) q * nprobe + r;
        const uint32_t l = probe[p];
        if (l >= (uint32_t) nlists) continue;
        const uint32_t len = list_len[l];
        if (len == 0) continue;
        const uint32_t pos = pairpos[p], nseg = ivf_nseg(len, segb);
        const uint32_t item0 = item_off[l] + (pos / tile) * nseg;
        const uint32_t it = pos % tile;
        for (uint32_t sg = 0; sg < nseg; sg++) {
            const size_t base = ((size_t) (item0 + sg) * tile + it) * k;
            for (int i = lane; i < round_up(k, 32); i += 32) {
                float cd = INFINITY;
                int64_t id = -1;
                bool valid = false;
                if (i < k) {
                    const uint32_t s = pslot[base + i];
                    if (s != INVALID_SLOT) { cd = pdist[base + i]; id = ids[s]; valid = true; }
                }
                top.offer(cd, id, valid, lane, k);
            }
        }
    }
#pragma unroll
    for (int r = 0; r < KR; r++) {
        const int e = r * 32 + lane;
        if (e < k) {
            const bool have = top.key[r] != KeyMax<int64_t>::v;
            out_dist[(size_t) q * k + e] = have ? top.d[r] : INFINITY;
            out_ids[(size_t) q * k + e] = have ? top.key[r] : -1;
        }
    }
}

// ---- tensor-core list scan (NDB_ARITH_TENSOR) ---------------------------------------------------
// Work items for tc_knn_kernel: (tile of 128 queries probing list l) x (segment of the list's
// 256-row tiles).  list_blk here is the list's first tile * 8, segb the segment length in 32-row
// blocks, so the bucketing kernels above are shared with the fp32 path.
__global__ void ivf_tc_items_kernel(const uint32_t *__restrict__ cnt, const uint32_t *__restrict__ qoff,
                                    const uint32_t *__restrict__ item_off, const uint32_t *__restrict__ list_len,
                                    const uint32_t *__restrict__ ltile8, int nlists, uint32_t segb, uint32_t rep_max,
                                    TcItem *__restrict__ items, uint32_t split, int seg_major)
{
    const int v = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;           // virtual list (ivf_vlist)
    if (v >= (split ? 2 * nlists : nlists)) return;
    const int l = v >= nlists ? v - nlists : v;
    const uint32_t rp = ivf_rep(cnt[v], rep_max);
    const uint32_t c = cnt[v] * rp;                            // tile positions in use
    const uint32_t tiles = (c + TC_M - 1) / TC_M;
    const uint32_t len = list_len[l], nseg = ivf_nseg(len, segb);
    for (uint32_t i = threadIdx.x & 31; i < tiles * nseg; i += 32) {
        const uint32_t t = i / nseg, sg = i - t * nseg;
        const uint32_t nvec = min(segb * 32, len - sg * segb * 32);
        TcItem it;
        it.qtile = qoff[v] / TC_M + t;
        it.t0 = (ltile8[l] + sg * segb) / 8;
        it.t1 = it.t0 + (nvec + TC_N - 1) / TC_N;
        it.nq = min((uint32_t) TC_M, c - t * TC_M);
        it.out_base = (item_off[v] + i) * (2 * TC_M);
        it.out_stride = 2;
        it.rep = rp;
        it.nrows = nvec;
        // Execution order within the list (the output slot above is not affected): segment-major puts the query tiles of
        // one segment on consecutive items, i.e. on different SMs at the same time, so the segment's rows come out of
        // HBM once and out of L2 for the other tiles; tile-major streams the whole list once per query tile.
        items[item_off[v] + (seg_major ? sg * tiles + t : i)] = it;
    }
}

// ---- literal ivfCollectCandidates (:1722-1909): warp per query ----------------------------------
// candidates = the first k*10 entries met walking the probed lists in probe order and each
// list in insertion order; then the reference's selection sort over an index array (strict <,
// swap), which fixes the order of exact ties.
template <class P>
__global__ void __launch_bounds__(128) ivf_literal_kernel(const float4 *__restrict__ vecs, const float *__restrict__ vnorm,
                                                          const float *__restrict__ Q, const uint32_t *__restrict__ probe,
                                                          const uint32_t *__restrict__ list_len, const uint32_t *__restrict__ list_blk,
                                                          const uint32_t *__restrict__ lit_order, const int64_t *__restrict__ ids,
                                                          int nq, int nprobe, int nlists, int dim, int dimp, int k,
                                                          float *__restrict__ out_dist, int64_t *__restrict__ out_ids)
{
    extern __shared__ unsigned char lit_smem[];
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const int q = blockIdx.x * 4 + w;
    const int maxc = k * 10;
    float *cd = reinterpret_cast<float *>(lit_smem) + (size_t) w * maxc;
    uint32_t *cs = reinterpret_cast<uint32_t *>(lit_smem + sizeof(float) * 4 * (size_t) maxc) + (size_t) w * maxc;
    int *idx = reinterpret_cast<int *>(lit_smem + 8 * 4 * (size_t) maxc) + (size_t) w * maxc;
    if (q >= nq) return;
    const float *qv = Q + (size_t) q * dim;
    // gather candidate slots
    int cc = 0;
    for (int i = 0; i < nprobe && cc < maxc; i++) {
        const uint32_t l = probe[(size_t) q * nprobe + i];
        if (l >= (uint32_t) nlists) continue;
        const uint32_t len = list_len[l], base = list_blk[l] * 32;
        const int take = min((int) len, maxc - cc);
        for (int j = lane; j < take; j += 32) cs[cc + j] = lit_order[base + j];
        cc += take;
    }
    __syncwarp();
    // distances: lane per candidate, sequential over the dimensions
    typename P::N qn = 0;
    if (P::NORMS) for (int j = 0; j < dim; j++) P::nstep(qn, qv[j]);
    for (int c = lane; c < cc; c += 32) {
        const uint32_t slot = cs[c];
        const float4 *vp = vecs + (size_t) (slot >> 5) * (8 * (size_t) dimp) + (slot & 31);
        typename P::Acc acc;
        P::init(acc);
        for (int j = 0; j < dim; j += 4) {
            const float4 x = vp[(size_t) (j >> 2) * 32];
            P::step(acc, x.x, qv[j]);
            if (j + 1 < dim) P::step(acc, x.y, qv[j + 1]);
            if (j + 2 < dim) P::step(acc, x.z, qv[j + 2]);
            if (j + 3 < dim) P::step(acc, x.w, qv[j + 3]);
        }
        cd[c] = P::finish(acc, P::NORMS ? vnorm[slot] : 0, qn);
        idx[c] = c;
    }
    __syncwarp();
    if (lane == 0) {
        const int actualK = min(k, cc);
        for (int i = 0; i < actualK; i++) {
            int bestIdx = i;
            float bestDist = cd[idx[i]];
            for (int j = i + 1; j < cc; j++)
                if (cd[idx[j]] < bestDist) { bestDist = cd[idx[j]]; bestIdx = j; }
            if (bestIdx != i) { const int t = idx[i]; idx[i] = idx[bestIdx]; idx[bestIdx] = t; }
        }
    }
    __syncwarp();
    for (int i = lane; i < k; i += 32) {
        const bool have = i < cc;
        out_dist[(size_t) q * k + i] = have ? cd[idx[i]] : INFINITY;
        out_ids[(size_t) q * k + i] = have ? ids[cs[idx[i]]] : -1;
    }
}

// relation loader: pull the float4[dim] payload of entry i out of the staged 8 KB pages
__global__ void extract_payload_kernel(const unsigned char *__restrict__ pages, const unsigned long long *__restrict__ src,
                                       int64_t n, int dim, float *__restrict__ arena)
{
    const int64_t t = (int64_t) blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n * dim) return;
    const int64_t i = t / dim;
    const int j = (int) (t - i * dim);
    arena[t] = reinterpret_cast<const float *>(pages + src[i])[j];
}

__global__ void copy_kept_rows_kernel(const float *__restrict__ rows, const uint32_t *__restrict__ keep, int64_t nkeep,
                                      int dim, float *__restrict__ arena)
{
    const int64_t t = (int64_t) blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= nkeep * dim) return;
    const int64_t i = t / dim;
    const int j = (int) (t - i * dim);
    arena[t] = rows[(size_t) keep[i] * dim + j];
}

static int ivf_sync_centroids(ndb_b200_ivf *ix, cudaStream_t s)
{
    NDB_CHECK(ix->kw.C.reserve((size_t) ix->nlists * ix->dim * 4));
    NDB_CUDA(cudaMemcpyAsync(ix->kw.C.p, ix->C_host.data(), (size_t) ix->nlists * ix->dim * 4, cudaMemcpyHostToDevice, s));
    NDB_CUDA(cudaStreamSynchronize(s));
    return NDB_B200_OK;
}

// ---- list layout on the device ----------------------------------------------------------------------------
// rows (insertion order) -> IL32 slots: list l occupies the 32-vector blocks [list_blk[l], list_blk[l] + ceil(len/32)),
// entries sorted by id inside a list (insertion order between equal ids).  Two stable radix sorts do it: by id, then
// by list; a third (by list alone) gives the insertion order inside each list for the literal mode.
__global__ void iota32_kernel(uint32_t *out, int64_t n)
{
    const int64_t i = (int64_t) blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = (uint32_t) i;
}
__global__ void gather_u32_kernel(const uint32_t *__restrict__ src, const uint32_t *__restrict__ idx, int64_t n, uint32_t *__restrict__ out)
{
    const int64_t i = (int64_t) blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = src[idx[i]];
}
// first position of every list in the list-sorted order: start[l] for l in [0, L], from the sorted list keys
__global__ void list_starts_kernel(const uint32_t *__restrict__ sorted_list, int64_t n, int L, uint32_t *__restrict__ start)
{
    const int64_t i = (int64_t) blockIdx.x * blockDim.x + threadIdx.x;
    if (i > n) return;
    const int cur = i < n ? (int) sorted_list[i] : L;
    const int prev = i > 0 ? (int) sorted_list[i - 1] : -1;
    for (int l = prev + 1; l <= cur && l <= L; l++) start[l] = (uint32_t) i;
}
// single CTA: list_len, list_blk (exclusive scan of the lists' block counts), ltile8 (of their 256-row tile counts, * 8)
__global__ void __launch_bounds__(1024) list_extents_kernel(const uint32_t *__restrict__ start, int L, uint32_t *__restrict__ list_len,
                                                             uint32_t *__restrict__ list_blk, uint32_t *__restrict__ ltile8,
                                                             unsigned long long *__restrict__ totals /* blocks, tiles */)
{
    typedef cub::BlockScan<unsigned long long, 1024> Scan;
    __shared__ typename Scan::TempStorage tmp;
    const int ipt = (L + 1023) / 1024;
    const int b = threadIdx.x * ipt, e = min(L, b + ipt);
    unsigned long long sb = 0, st = 0;
    for (int l = b; l < e; l++) {
        const uint32_t len = start[l + 1] - start[l];
        sb += (len + 31) / 32;
        st += (len + TC_N - 1) / TC_N;
    }
    unsigned long long ob, ot, tb, tt;
    Scan(tmp).ExclusiveSum(sb, ob, tb);
    __syncthreads();
    Scan(tmp).ExclusiveSum(st, ot, tt);
    for (int l = b; l < e; l++) {
        const uint32_t len = start[l + 1] - start[l];
        list_len[l] = len;
        list_blk[l] = (uint32_t) ob;
        ltile8[l] = (uint32_t) (ot * 8);
        ob += (len + 31) / 32;
        ot += (len + TC_N - 1) / TC_N;
    }
    if (threadIdx.x == 0) { totals[0] = tb; totals[1] = tt; }
}
// position p of the (list, id)-sorted order: row, list -> slot; fills slot_of_row, ids_by_slot, row_of_slot
__global__ void place_rows_kernel(const uint32_t *__restrict__ sorted_row, const uint32_t *__restrict__ sorted_list,
                                  const uint32_t *__restrict__ start, const uint32_t *__restrict__ list_blk,
                                  const int64_t *__restrict__ row_id, int64_t n, uint32_t *__restrict__ slot_of_row,
                                  int64_t *__restrict__ ids_by_slot, uint32_t *__restrict__ row_of_slot)
{
    const int64_t p = (int64_t) blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= n) return;
    const uint32_t row = sorted_row[p], l = sorted_list[p];
    const uint32_t slot = list_blk[l] * 32 + (uint32_t) (p - start[l]);
    slot_of_row[row] = slot;
    ids_by_slot[slot] = row_id[row];
    row_of_slot[slot] = row;
}
// position p of the list-sorted (insertion) order: the j-th inserted row of list l sits at lit[list_blk[l] * 32 + j]
__global__ void place_literal_kernel(const uint32_t *__restrict__ ins_row, const uint32_t *__restrict__ ins_list,
                                     const uint32_t *__restrict__ start, const uint32_t *__restrict__ list_blk,
                                     const uint32_t *__restrict__ slot_of_row, int64_t n, uint32_t *__restrict__ lit)
{
    const int64_t p = (int64_t) blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= n) return;
    const uint32_t l = ins_list[p];
    lit[list_blk[l] * 32 + (uint32_t) (p - start[l])] = slot_of_row[ins_row[p]];
}
// tensor layout maps: tensor row ltile8[l] * 32 + j  <->  IL32 slot list_blk[l] * 32 + j, arena row
__global__ void tensor_maps_kernel(const uint32_t *__restrict__ sorted_list, const uint32_t *__restrict__ start,
                                   const uint32_t *__restrict__ list_blk, const uint32_t *__restrict__ ltile8,
                                   const uint32_t *__restrict__ row_of_slot, int64_t n, uint32_t *__restrict__ tc_src,
                                   uint32_t *__restrict__ tc_row)
{
    const int64_t p = (int64_t) blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= n) return;
    const uint32_t l = sorted_list[p], j = (uint32_t) (p - start[l]);
    const uint32_t slot = list_blk[l] * 32 + j, trow = ltile8[l] * 32 + j;
    tc_src[trow] = slot;
    tc_row[trow] = row_of_slot[slot];
}

// (re)build the IL32 list store from the arena
static int ivf_layout(ndb_b200_ivf *ix, cudaStream_t s)
{
    if (!ix->dirty) return NDB_B200_OK;
    const int64_t n = ix->nrows;
    const int L = ix->nlists;
    NDB_REQUIRE(n < (int64_t) 0x7ffffff0ll, NDB_B200_EINVAL, "ivf: too many rows for 32-bit slot ids");
    const int64_t n1 = n ? n : 1;
    DevBuf d_list, d_id, k64a, k64b, v32a, v32b, l32a, l32b, cub_tmp, d_tot;
    NDB_CHECK(d_list.reserve((size_t) n1 * 4)); NDB_CHECK(d_id.reserve((size_t) n1 * 8));
    NDB_CHECK(k64b.reserve((size_t) n1 * 8));
    NDB_CHECK(v32a.reserve((size_t) n1 * 4)); NDB_CHECK(v32b.reserve((size_t) n1 * 4));
    NDB_CHECK(l32a.reserve((size_t) n1 * 4)); NDB_CHECK(l32b.reserve((size_t) n1 * 4));
    NDB_CHECK(d_tot.reserve(16));
    NDB_CHECK(ix->d_list_len.reserve((size_t) L * 4));
    NDB_CHECK(ix->d_list_blk.reserve((size_t) L * 4));
    NDB_CHECK(ix->d_list_order.reserve((size_t) L * 4));
    NDB_CHECK(ix->d_ltile8.reserve((size_t) L * 4));
    NDB_CHECK(ix->d_start.reserve((size_t) (L + 1) * 4));
    NDB_CHECK(ix->tmp_assign.reserve((size_t) n1 * 4));
    NDB_CHECK(ix->d_sorted_list.reserve((size_t) n1 * 4));
    const unsigned g = (unsigned) ((n1 + 255) / 256);
    if (n) {
        NDB_CUDA(cudaMemcpyAsync(d_list.p, ix->row_list.data(), (size_t) n * 4, cudaMemcpyHostToDevice, s));
        NDB_CUDA(cudaMemcpyAsync(d_id.p, ix->row_id.data(), (size_t) n * 8, cudaMemcpyHostToDevice, s));
        int lbits = 1;
        while ((1ll << lbits) < L) lbits++;
        size_t t1 = 0, t2 = 0;
        NDB_CUDA(cub::DeviceRadixSort::SortPairs(nullptr, t1, d_id.as<unsigned long long>(), k64b.as<unsigned long long>(), v32a.as<uint32_t>(),
                                                 v32b.as<uint32_t>(), (int) n, 0, 64, s));
        NDB_CUDA(cub::DeviceRadixSort::SortPairs(nullptr, t2, l32a.as<uint32_t>(), l32b.as<uint32_t>(), v32a.as<uint32_t>(), v32b.as<uint32_t>(),
                                                 (int) n, 0, lbits, s));
        NDB_CHECK(cub_tmp.reserve(std::max(t1, t2)));
        // (a) insertion order inside each list: stable sort of the row indices by list
        iota32_kernel<<<g, 256, 0, s>>>(v32a.as<uint32_t>(), n);
        size_t tb = cub_tmp.cap;
        NDB_CUDA(cub::DeviceRadixSort::SortPairs(cub_tmp.p, tb, d_list.as<uint32_t>(), l32b.as<uint32_t>(), v32a.as<uint32_t>(), v32b.as<uint32_t>(),
                                                 (int) n, 0, lbits, s));
        // l32b = lists in sorted order, v32b = rows in (list, insertion) order
        list_starts_kernel<<<(unsigned) ((n + 1 + 255) / 256), 256, 0, s>>>(l32b.as<uint32_t>(), n, L, ix->d_start.as<uint32_t>());
        list_extents_kernel<<<1, 1024, 0, s>>>(ix->d_start.as<uint32_t>(), L, ix->d_list_len.as<uint32_t>(), ix->d_list_blk.as<uint32_t>(),
                                               ix->d_ltile8.as<uint32_t>(), d_tot.as<unsigned long long>());
        count_launch(6);
        NDB_CUDA(cudaGetLastError());
    } else {
        NDB_CUDA(cudaMemsetAsync(ix->d_start.p, 0, (size_t) (L + 1) * 4, s));
        list_extents_kernel<<<1, 1024, 0, s>>>(ix->d_start.as<uint32_t>(), L, ix->d_list_len.as<uint32_t>(), ix->d_list_blk.as<uint32_t>(),
                                               ix->d_ltile8.as<uint32_t>(), d_tot.as<unsigned long long>());
        count_launch();
    }
    unsigned long long tot[2] = {0, 0};
    ix->list_len.resize(L);
    ix->list_blk.resize(L);
    NDB_CUDA(cudaMemcpyAsync(tot, d_tot.p, 16, cudaMemcpyDeviceToHost, s));
    NDB_CUDA(cudaMemcpyAsync(ix->list_len.data(), ix->d_list_len.p, (size_t) L * 4, cudaMemcpyDeviceToHost, s));
    NDB_CUDA(cudaMemcpyAsync(ix->list_blk.data(), ix->d_list_blk.p, (size_t) L * 4, cudaMemcpyDeviceToHost, s));
    NDB_CUDA(cudaStreamSynchronize(s));
    const uint64_t blk = tot[0];
    ix->tc_tiles = tot[1];
    NDB_REQUIRE(blk * 32 < 0xfffffff0ull, NDB_B200_EINVAL, "ivf: too many slots for 32-bit slot ids");
    const int64_t nslots = (int64_t) blk * 32, ns1 = nslots ? nslots : 1;
    ix->store.dim = ix->dim;
    ix->store.dimp = ix->dimp;
    ix->store.nblk = (int64_t) blk;
    const size_t bytes = ix->store.bytes_for((int64_t) blk);
    NDB_CHECK(ix->store.data.reserve(bytes ? bytes : 16));
    NDB_CHECK(ix->ids.reserve((size_t) ns1 * 8));
    NDB_CHECK(ix->lit_order.reserve((size_t) ns1 * 4));
    NDB_CHECK(ix->d_row_of_slot.reserve((size_t) ns1 * 4));
    if (bytes) NDB_CUDA(cudaMemsetAsync(ix->store.data.p, 0, bytes, s));
    NDB_CUDA(cudaMemsetAsync(ix->ids.p, 0xFF, (size_t) ns1 * 8, s));              // -1
    NDB_CUDA(cudaMemsetAsync(ix->lit_order.p, 0xFF, (size_t) ns1 * 4, s));        // INVALID_SLOT
    NDB_CUDA(cudaMemsetAsync(ix->d_row_of_slot.p, 0xFF, (size_t) ns1 * 4, s));
    if (n) {
        int lbits = 1;
        while ((1ll << lbits) < L) lbits++;
        // (b) (list, id) order: stable sort by id, then stable sort by list
        iota32_kernel<<<g, 256, 0, s>>>(v32a.as<uint32_t>(), n);
        size_t tb = cub_tmp.cap;
        NDB_CHECK(k64a.reserve((size_t) n * 8));
        // ids are ordered as SIGNED int64 by the reference's (dist, id) rule: flip the sign bit for the unsigned radix sort
        NDB_CUDA(cub::DeviceRadixSort::SortPairs(cub_tmp.p, tb, d_id.as<long long>(), k64b.as<long long>(), v32a.as<uint32_t>(),
                                                 l32a.as<uint32_t>(), (int) n, 0, 64, s));
        // l32a = rows by ascending id; their lists, then the second sort
        gather_u32_kernel<<<g, 256, 0, s>>>(d_list.as<uint32_t>(), l32a.as<uint32_t>(), n, v32a.as<uint32_t>());
        tb = cub_tmp.cap;
        NDB_CUDA(cub::DeviceRadixSort::SortPairs(cub_tmp.p, tb, v32a.as<uint32_t>(), ix->d_sorted_list.as<uint32_t>(), l32a.as<uint32_t>(),
                                                 k64a.as<uint32_t>(), (int) n, 0, lbits, s));
        // d_sorted_list = lists, k64a (as u32) = rows, both in (list, id) order
        place_rows_kernel<<<g, 256, 0, s>>>(k64a.as<uint32_t>(), ix->d_sorted_list.as<uint32_t>(), ix->d_start.as<uint32_t>(),
                                            ix->d_list_blk.as<uint32_t>(), d_id.as<int64_t>(), n, ix->tmp_assign.as<uint32_t>(),
                                            ix->ids.as<int64_t>(), ix->d_row_of_slot.as<uint32_t>());
        place_literal_kernel<<<g, 256, 0, s>>>(v32b.as<uint32_t>(), l32b.as<uint32_t>(), ix->d_start.as<uint32_t>(), ix->d_list_blk.as<uint32_t>(),
                                               ix->tmp_assign.as<uint32_t>(), n, ix->lit_order.as<uint32_t>());
        count_launch(8);
        NDB_CUDA(cudaGetLastError());
        NDB_CHECK(il32_scatter(ix->arena.as<float>(), n, ix->dim, ix->dimp, ix->tmp_assign.as<uint32_t>(), 0, ix->store.ptr(), s));
    }
    // lists by descending length (stable): the work items are emitted longest first
    std::vector<uint32_t> lorder(L);
    std::iota(lorder.begin(), lorder.end(), 0u);
    std::stable_sort(lorder.begin(), lorder.end(), [&](uint32_t a, uint32_t b) { return ix->list_len[a] > ix->list_len[b]; });
    NDB_CUDA(cudaMemcpyAsync(ix->d_list_order.p, lorder.data(), (size_t) L * 4, cudaMemcpyHostToDevice, s));
    NDB_CUDA(cudaStreamSynchronize(s));     // host vectors / temporaries above go out of scope
    ix->vnorm_ivf_ok = ix->vnorm_fast_ok = false;
    ix->tc_ok = false;
    ix->dirty = false;
    return NDB_B200_OK;
}

static int ivf_coarse(ndb_b200_ivf *ix, const float *Q_dev, int nq, int np, int arith, cudaStream_t s);

static uint32_t ivf_tc_seg_tiles()
{
    static const uint32_t v = [] { const char *e = getenv("NDB_IVF_TC_SEG_TILES"); int x = e ? atoi(e) : 16; return (uint32_t) (x >= 1 && x <= TC_PACKED_MAX_TILES ? x : 16); }();
    return v;
}

// blocked bf16 copies for the tensor path: lists (each padded to whole 256-row tiles) and centroids
static int ivf_tensor_ready(ndb_b200_ivf *ix, cudaStream_t s)
{
    const int L = ix->nlists;
    if (ix->tc_ok) return NDB_B200_OK;
    const uint64_t nt = ix->tc_tiles;                      // 256-row tiles of all lists (ivf_layout)
    NDB_REQUIRE(nt * TC_N < 0xfffffff0ull, NDB_B200_EINVAL, "ivf: too many rows for the tensor copy");
    const size_t nrow_t = (size_t) (nt ? nt : 1) * TC_N;
    NDB_CHECK(ix->tc_src.reserve(nrow_t * 4));
    NDB_CHECK(ix->tc_row.reserve(nrow_t * 4));
    NDB_CUDA(cudaMemsetAsync(ix->tc_src.p, 0xFF, nrow_t * 4, s));          // pad rows: INVALID_SLOT
    NDB_CUDA(cudaMemsetAsync(ix->tc_row.p, 0xFF, nrow_t * 4, s));
    if (ix->nrows) {
        tensor_maps_kernel<<<(unsigned) ((ix->nrows + 255) / 256), 256, 0, s>>>(ix->d_sorted_list.as<uint32_t>(), ix->d_start.as<uint32_t>(),
                                                                                ix->d_list_blk.as<uint32_t>(), ix->d_ltile8.as<uint32_t>(),
                                                                                ix->d_row_of_slot.as<uint32_t>(), ix->nrows,
                                                                                ix->tc_src.as<uint32_t>(), ix->tc_row.as<uint32_t>());
        count_launch();
        NDB_CUDA(cudaGetLastError());
    }
    NDB_CHECK(tc_build_store_mapped(ix->tc, ix->store.ptr(), ix->tc_src.as<uint32_t>(), (int64_t) (nt ? nt : 1) * TC_N, ix->dim,
                                    ix->dimp, s));
    NDB_CUDA(cudaStreamSynchronize(s));
    const uint32_t segb = ivf_tc_seg_tiles() * 8;
    ix->tc_max_nseg = 1;
    ix->tc_sum_nseg = ix->tc_nonempty = 0;
    double s1 = 0.0, s2 = 0.0;
    for (int l = 0; l < L; l++) { s1 += ix->list_len[l]; s2 += (double) ix->list_len[l] * ix->list_len[l]; }
    ix->tc_rows_per_pair = s1 > 0.0 ? s2 / s1 : 0.0;      // queries fall on lists in proportion to their length
    for (int l = 0; l < L; l++) {
        if (!ix->list_len[l]) continue;
        const uint32_t ns = ivf_nseg(ix->list_len[l], segb);
        ix->tc_max_nseg = std::max(ix->tc_max_nseg, ns);
        ix->tc_sum_nseg += ns;
        ix->tc_nonempty++;
    }
    ix->tc_ok = true;
    return NDB_B200_OK;
}

static int ivf_tc_margin()
{
    static const int v = [] { const char *e = getenv("NDB_IVF_TC_MARGIN"); int x = e ? atoi(e) : 6; return x >= 0 ? x : 6; }();
    return v;
}

// NDB_ARITH_TENSOR search: coarse quantizer and list scans on the tensor cores (bf16 products, fp32
// accumulation) select k + margin candidates per query, which are then re-ranked in fp32.
static int ivf_search_tensor(ndb_b200_ivf *ix, const float *Q_dev, int nq, int np, int k, float *dist_dev, int64_t *ids_dev,
                             cudaStream_t s)
{
    NDB_REQUIRE(k <= TC_IVF_KMAX, NDB_B200_EINVAL, "ivf tensor path: k=%d > %d", k, TC_IVF_KMAX);
    NDB_REQUIRE(ix->dim <= TC_MAX_DIM, NDB_B200_EINVAL, "ivf tensor path: dim %d > %d", ix->dim, TC_MAX_DIM);
    NDB_CHECK(ivf_tensor_ready(ix, s));
    const int L = ix->nlists;
    const int64_t npairs = (int64_t) nq * np;
    const int kc = std::min(TC_KMAX, k + ivf_tc_margin());

    // 1. coarse quantizer: k = nprobe nearest centroids (always L2, ivfSelectClusters)
    NDB_CHECK(ix->probe.reserve((size_t) npairs * 4));
    NDB_CHECK(ix->cdist.reserve((size_t) npairs * 4));
    NDB_CHECK(ix->cert_counters.reserve(64));
    NDB_CHECK(ix->fb_list.reserve((size_t) nq * 4));
    NDB_CHECK(ix->fb_tau.reserve((size_t) nq * 4));
    NDB_CUDA(cudaMemsetAsync(ix->cert_counters.p, 0, 64, s));
    if (np <= 32 && ix->dim <= TC_MAX_DIM) {
        const int world = ix->coarse_by_rank ? comm_nranks() : 1;
        if (world > 1) {
            // Sharded search: every rank would select the same probe lists for the same queries.  Each rank does it for
            // its slice of the batch and one all-gather hands every rank all of them (certified selection: the probe
            // lists do not depend on who computed them).
            const int slice = (nq + world - 1) / world, rank = comm_rank();
            const int q0 = std::min(nq, rank * slice), q1 = std::min(nq, q0 + slice);
            NDB_CHECK(ix->probe.reserve((size_t) world * slice * np * 4));
            NDB_CHECK(ix->cdist.reserve((size_t) slice * np * 4));
            uint32_t *mine = ix->probe.as<uint32_t>() + (size_t) rank * slice * np;
            if (q1 > q0)
                NDB_CHECK(nearest_centroids_tensor(ix->kw.cs, ix->kw.C.as<float>(), ix->kw.cstore.as<float>(), L, ix->dim, ix->dimp,
                                                   Q_dev + (size_t) q0 * ix->dim, q1 - q0, np, false, mine, ix->cdist.as<float>(),
                                                   ix->cert_counters.as<unsigned long long>() + 2, s));
            NDB_CHECK(comm_allgather(mine, ix->probe.p, (size_t) slice * np * 4, s));
        } else {
            NDB_CHECK(nearest_centroids_tensor(ix->kw.cs, ix->kw.C.as<float>(), ix->kw.cstore.as<float>(), L, ix->dim, ix->dimp, Q_dev, nq, np,
                                               false, ix->probe.as<uint32_t>(), ix->cdist.as<float>(),
                                               ix->cert_counters.as<unsigned long long>() + 2, s));
        }
    } else {
        NDB_CHECK(ivf_coarse(ix, Q_dev, nq, np, NDB_ARITH_IVF_F32, s));
    }

    // 2. bucket (query, list) pairs by list; query positions padded to whole 128-query tiles per list
    // Two phases when a (query, list) pair scans many rows (ivf_vlist): the bound the nearest lists establish then saves far
    // more epilogue work in the other lists than the extra, thinly filled query tiles of phase 1 cost.
    const char *phases_e = getenv("NDB_IVF_TC_PHASES");        // 1 / 2 force the choice (measurement switch)
    const int phases_env = phases_e ? atoi(phases_e) : 0;
    // (expected evaluations of the batch on this rank: below ~1e9 the scan is too short to repay two more launches and
    // the thinly filled query tiles of phase 1 -- C2, 0.35e9: 0.56 ms in two phases, 0.48 ms in one)
    // (k > 16: a 16-entry partial list cannot bound the k-th best by itself, so only the bound kernel of the two-phase
    // scan gives such a search a shared bound at all)
    const bool two_phase = np > 1 && (phases_env == 2 || (phases_env != 1 && (k > TC_KMAX || ix->tc_rows_per_pair * (double) npairs >= 1.0e9)));
    const uint32_t split = two_phase ? (uint32_t) L : 0u;
    const int NV = two_phase ? 2 * L : L;                   // virtual lists
    NDB_CHECK(ix->cnt.reserve((size_t) NV * 4 * 2));
    NDB_CHECK(ix->qoff.reserve((size_t) NV * 4));
    NDB_CHECK(ix->item_off.reserve((size_t) NV * 4));
    NDB_CHECK(ix->pairpos.reserve((size_t) npairs * 4));
    NDB_CHECK(ix->nitems.reserve(64));
    NDB_CHECK(ix->stats.reserve(64));
    uint32_t *cnt = ix->cnt.as<uint32_t>(), *fill = cnt + NV;
    const uint32_t segb = ivf_tc_seg_tiles() * 8;
    NDB_CUDA(cudaMemsetAsync(cnt, 0, (size_t) NV * 4 * 2, s));
    ivf_hist_kernel<<<(unsigned) ((npairs + 255) / 256), 256, 0, s>>>(ix->probe.as<uint32_t>(), npairs, ix->d_list_len.as<uint32_t>(), L, cnt,
                                                                     (uint32_t) np, split);
    static const uint32_t rep_max = [] { const char *e = getenv("NDB_IVF_TC_REP_MAX"); int x = e ? atoi(e) : 32; return (uint32_t) (x >= 0 && x <= 32 ? x : 32); }();
    ivf_offsets_kernel<<<1, 1024, 0, s>>>(cnt, ix->d_list_len.as<uint32_t>(), ix->d_list_order.as<uint32_t>(), L, TC_M, segb,
                                          (uint32_t) TC_M, rep_max, ix->qoff.as<uint32_t>(), ix->item_off.as<uint32_t>(),
                                          ix->nitems.as<uint32_t>(), ix->stats.as<unsigned long long>(), split);
    count_launch(2);
    NDB_CUDA(cudaGetLastError());
    // How many work items and query-tile positions this batch needs depends on how its probes fall on
    // the lists.  Both have host-side upper bounds (every list with queries wastes < 128 positions;
    // items <= max_nseg * pairs / 128 + one per segment of a non-empty list); when the buffers those
    // bounds ask for are affordable the kernels read the exact counts from device memory and the
    // search needs no host round trip at all.  Otherwise one 8-byte read-back sizes them exactly.
    const uint64_t vmul = two_phase ? 2 : 1;
    size_t n_items = (size_t) ix->tc_max_nseg * (size_t) ((npairs + TC_M - 1) / TC_M) + vmul * ix->tc_sum_nseg;
    size_t npos = (size_t) npairs + (size_t) (TC_M - 1) * std::min<uint64_t>(vmul * ix->tc_nonempty, (uint64_t) npairs);
    npos = (npos + TC_M - 1) / TC_M * TC_M;
    const bool exact_counts = getenv("NDB_IVF_TC_SYNC") != nullptr ||
                              n_items * 2 * TC_M * kc * 8 + npos * ix->tc.nkc * TC_KC * 2 > ((size_t) 1 << 30);
    if (exact_counts) {
        uint32_t *h_tot = reinterpret_cast<uint32_t *>(ctx().pinned);
        NDB_CUDA(cudaMemcpyAsync(h_tot, ix->nitems.p, 8, cudaMemcpyDeviceToHost, s));
        NDB_CUDA(cudaStreamSynchronize(s));
        n_items = h_tot[0];
        npos = h_tot[1];
    }
    if (getenv("NDB_IVF_DEBUG")) fprintf(stderr, "ivf tensor: %zu items, %zu query positions for %lld pairs (%s); %.0f rows per pair expected, %d phase(s)\n", n_items, npos, (long long) npairs, exact_counts ? "exact" : "bounds", ix->tc_rows_per_pair, two_phase ? 2 : 1);
    if (n_items == 0) n_items = 1;          // (nothing to scan: the kernels see the zero count on the device)
    if (npos == 0) npos = TC_M;
    NDB_CHECK(ix->qmap.reserve(npos * 4));
    NDB_CUDA(cudaMemsetAsync(ix->qmap.p, 0xFF, npos * 4, s));
    ivf_scatter_kernel<<<(unsigned) ((npairs + 255) / 256), 256, 0, s>>>(ix->probe.as<uint32_t>(), npairs, ix->d_list_len.as<uint32_t>(), L,
                                                                        ix->qoff.as<uint32_t>(), fill, ix->qmap.as<uint32_t>(),
                                                                        ix->pairpos.as<uint32_t>(), cnt, rep_max, (uint32_t) np, split);
    NDB_CHECK(ix->tcs.items.reserve(n_items * sizeof(TcItem)));
    ivf_tc_items_kernel<<<(unsigned) ((NV + 3) / 4), 128, 0, s>>>(cnt, ix->qoff.as<uint32_t>(), ix->item_off.as<uint32_t>(),
                                                                 ix->d_list_len.as<uint32_t>(), ix->d_ltile8.as<uint32_t>(), L, segb,
                                                                 rep_max, ix->tcs.items.as<TcItem>(), split,
                                                                 getenv("NDB_IVF_TC_TILE_MAJOR") ? 0 : 1);
    count_launch(2);
    NDB_CUDA(cudaGetLastError());
    if (getenv("NDB_IVF_DEBUG")) {
        // tile-steps by live queries per item: how full are the 128-query tiles the tensor cores work on
        uint32_t cnt_h[2];
        NDB_CUDA(cudaMemcpyAsync(cnt_h, ix->nitems.p, 8, cudaMemcpyDeviceToHost, s));
        NDB_CUDA(cudaStreamSynchronize(s));
        std::vector<TcItem> hi(cnt_h[0]);
        NDB_CUDA(cudaMemcpy(hi.data(), ix->tcs.items.p, hi.size() * sizeof(TcItem), cudaMemcpyDeviceToHost));
        unsigned long long steps[5] = {0, 0, 0, 0, 0}, tot = 0, live = 0;
        for (const TcItem &t : hi) {
            const unsigned ts = t.t1 - t.t0;
            steps[t.nq <= 16 ? 0 : t.nq <= 32 ? 1 : t.nq <= 64 ? 2 : t.nq <= 96 ? 3 : 4] += ts;
            tot += ts;
            live += (unsigned long long) ts * t.nq;
        }
        fprintf(stderr, "ivf tensor: %u items, %llu tile-steps; by live queries <=16:%llu <=32:%llu <=64:%llu <=96:%llu <=128:%llu; mean live %.1f\n",
                cnt_h[0], tot, steps[0], steps[1], steps[2], steps[3], steps[4], tot ? (double) live / tot : 0.0);
    }

    // 3. gather the query tiles (blocked bf16) and run the tensor kernel over the items
    const int nkc = ix->tc.nkc;
    NDB_CHECK(ix->tcs.qb.reserve(npos * nkc * TC_KC * 2));
    NDB_CHECK(ix->tcs.qnorm.reserve(npos * 4));
    NDB_CHECK(ix->tcs.qerr.reserve(npos * 4));
    NDB_CHECK(tc_block_queries(Q_dev, ix->qmap.as<uint32_t>(), (uint32_t) np, nq, (int) npos, ix->dim, nkc,
                               ix->tcs.qb.as<__nv_bfloat16>(), ix->tcs.qnorm.as<float>(), s, ix->nitems.as<uint32_t>() + 1,
                               ix->tcs.qerr.as<float>(), ix->metric == NDB_IP));
    const size_t nparts = n_items * 2 * TC_M;
    NDB_CHECK(ix->tcs.pdist.reserve(nparts * kc * 4));
    NDB_CHECK(ix->tcs.pslot.reserve(nparts * kc * 4));
    TcParams p;
    memset(&p, 0, sizeof(p));
    p.xb = ix->tc.xb.as<__nv_bfloat16>();
    p.xnorm = ix->tc.xnorm.as<float>();
    if (ix->metric == NDB_COSINE) NDB_CHECK(tc_store_rinv(ix->tc, &p.xnorm, s));
    if (ix->metric == NDB_IP) NDB_CHECK(tc_store_pad0(ix->tc, &p.xnorm, s));
    p.qb = ix->tcs.qb.as<__nv_bfloat16>();
    p.qnorm = ix->tcs.qnorm.as<float>();
    p.qerr = ix->tcs.qerr.as<float>();
    p.cstats = ix->tc.stats.as<float>();
    p.dim = ix->dim;
    p.nkc = nkc;
    p.k = kc;
    p.items = ix->tcs.items.as<TcItem>();
    p.nitems = (uint32_t) n_items;
    p.nitems_ptr = ix->nitems.as<uint32_t>();
    p.pdist = ix->tcs.pdist.as<float>();
    p.pslot = ix->tcs.pslot.as<uint32_t>();
    NDB_CHECK(ix->tcs.gthr.reserve((size_t) nq * 4));
    static int calls = 0;
    if (!getenv("NDB_IVF_TC_EXPERIMENT_KEEP_BOUNDS") || calls++ == 0)      // experiment only: start from the previous call's final bounds
        NDB_CUDA(cudaMemsetAsync(ix->tcs.gthr.p, 0x7f, (size_t) nq * 4, s));      // 3.4e38: "no bound yet"
    p.qmap = ix->qmap.as<uint32_t>();
    p.nprobe = (uint32_t) np;
    p.gthr = getenv("NDB_IVF_TC_NOSHARE") ? nullptr : ix->tcs.gthr.as<float>();
    p.packed = getenv("NDB_IVF_TC_UNPACKED") ? 0 : 1;
    p.kpub = k > kc ? -1 : (getenv("NDB_IVF_TC_KPUB_LAST") ? 0 : k);
    p.debug_mode = getenv("NDB_TC_DEBUG") ? atoi(getenv("NDB_TC_DEBUG")) : 0;
    unsigned long long *ctr = ix->cert_counters.as<unsigned long long>();
    if (two_phase) {
        // Phase 1: the queries' nearest lists.  Then the k-th best key over each query's partial lists there becomes its
        // bound (relaxed, so that the selection stays certifiable).  Phase 2: the other nprobe - 1 lists -- 97 % of the
        // rows, of which then only a handful per query pass the bound at all.
        uint32_t *nit = ix->nitems.as<uint32_t>();
        p.item_lo_ptr = nullptr;
        p.item_hi_ptr = nit + 2;
        NDB_CHECK(tc_launch(p, ix->metric, kc, s));
        const unsigned bgrid = (unsigned) ((nq + 3) / 4);
#define NDB_BND(M)                                                                                                   \
        ivf_tc_bound_kernel<M><<<bgrid, 128, 0, s>>>(ix->tcs.pdist.as<float>(), ix->tcs.pslot.as<uint32_t>(), Q_dev,  \
                                                     ix->probe.as<uint32_t>(), ix->pairpos.as<uint32_t>(),            \
                                                     ix->item_off.as<uint32_t>(), ix->d_list_len.as<uint32_t>(), cnt, \
                                                     rep_max, nq, np, L, segb, ix->dim, kc, k, ix->tc.stats.as<float>(), p.gthr)
        if (ix->metric == NDB_L2) NDB_BND(NDB_L2);
        else if (ix->metric == NDB_COSINE) NDB_BND(NDB_COSINE);
        else NDB_BND(NDB_IP);
#undef NDB_BND
        count_launch();
        NDB_CUDA(cudaGetLastError());
        p.item_lo_ptr = nit + 2;
        p.item_hi_ptr = nullptr;
        NDB_CHECK(tc_launch(p, ix->metric, kc, s));         // (the library's kernel timer reports this, the dominant launch)
    } else {
        NDB_CHECK(tc_launch(p, ix->metric, kc, s));
    }
    if (ctx().timing) {                 // evals resolved from ix->stats in last_kernel_stats
        Context &c = ctx();
        c.last_bytes = -1.0;
        c.last_evals = -1;
        c.stats_src = two_phase ? (const void *) (ix->stats.as<unsigned long long>() + 1) : ix->stats.p;   // the timed launch's share
        c.stats_dim = ix->dim;
    }

    // 4. merge + certified fp32 re-rank; the queries the certificate rejects are recomputed exactly
    const unsigned mgrid = (unsigned) ((nq + 3) / 4);
    const size_t fsm = (size_t) ix->dimp * 4 + FB_THREADS * 4 + FB_THREADS * 8;
    const unsigned fgrid = (unsigned) std::min<int>(nq, 2 * ctx().sm_count);
    NDB_REQUIRE(p.packed && p.gthr, NDB_B200_EINVAL, "ivf tensor path: the certified finish needs packed keys and the shared bound");
    float *cert_dbg = nullptr;           // NDB_CERT_DEBUG: (query, G, tau, bound(G), |S|, ||eq||, ||q||, max||ex||) of the first 16 full scans
    if (getenv("NDB_CERT_DEBUG")) {
        NDB_CHECK(ix->cert_dbg.reserve(16 * 8 * 4));
        NDB_CUDA(cudaMemsetAsync(ix->cert_dbg.p, 0, 16 * 8 * 4, s));
        cert_dbg = ix->cert_dbg.as<float>();
    }
#define NDB_FIN(M)                                                                                                   \
    do {                                                                                                             \
        auto fin = k <= TC_KMAX ? ivf_tc_finish_cert_kernel<Arith<M, NDB_ARITH_IVF_F32>, M, 1>                        \
                                : ivf_tc_finish_cert_kernel<Arith<M, NDB_ARITH_IVF_F32>, M, 2>;                       \
        fin<<<mgrid, 128, 0, s>>>(                                                                                    \
            ix->tcs.pdist.as<float>(), ix->tcs.pslot.as<uint32_t>(), ix->tc_src.as<uint32_t>(), ix->tc_row.as<uint32_t>(), \
            ix->arena.as<float>(), ix->ids.as<int64_t>(), Q_dev, ix->probe.as<uint32_t>(),                            \
            ix->pairpos.as<uint32_t>(), ix->item_off.as<uint32_t>(), ix->d_list_len.as<uint32_t>(), p.gthr, cnt, rep_max, split, \
            nq, np, L, segb, ix->dim, kc, k, ix->tc.stats.as<float>(), dist_dev, ids_dev, ix->fb_list.as<uint32_t>(),  \
            ix->fb_tau.as<float>(), ctr);                                                                             \
        ivf_exact_fallback_kernel<Arith<M, NDB_ARITH_IVF_F32>, M><<<fgrid, FB_THREADS, fsm, s>>>(                            \
            ix->fb_list.as<uint32_t>(), ix->fb_tau.as<float>(), ctr, ix->tcs.pdist.as<float>(), ix->tcs.pslot.as<uint32_t>(), ix->probe.as<uint32_t>(), \
            ix->pairpos.as<uint32_t>(), ix->item_off.as<uint32_t>(), ix->d_list_len.as<uint32_t>(),                   \
            ix->d_ltile8.as<uint32_t>(), p.gthr, cnt, rep_max, split, segb, kc, ix->tc_src.as<uint32_t>(),                    \
            ix->tc_row.as<uint32_t>(), ix->arena.as<float>(), ix->ids.as<int64_t>(), Q_dev, ix->tc.stats.as<float>(), \
            reinterpret_cast<const float4 *>(ix->store.ptr()), ix->d_list_blk.as<uint32_t>(), ix->dimp,                \
            np, L, ix->dim, k, dist_dev, ids_dev, cert_dbg);                                                          \
    } while (0)
    if (ix->metric == NDB_L2) NDB_FIN(NDB_L2);
    else if (ix->metric == NDB_COSINE) NDB_FIN(NDB_COSINE);
    else NDB_FIN(NDB_IP);
#undef NDB_FIN
    count_launch();
    count_launch();
    NDB_CUDA(cudaGetLastError());
    return NDB_B200_OK;
}

static int ivf_vnorm(ndb_b200_ivf *ix, int arith, const void **out, cudaStream_t s)
{
    DevBuf &b = arith == NDB_ARITH_FAST ? ix->vnorm_fast : ix->vnorm_ivf;
    bool &ok = arith == NDB_ARITH_FAST ? ix->vnorm_fast_ok : ix->vnorm_ivf_ok;
    const int64_t nslots = ix->store.nblk * 32;
    if (!ok) {
        NDB_CHECK(b.reserve((size_t) (nslots ? nslots : 1) * 4));
        NDB_CHECK(slot_norms(arith, ix->store.ptr(), nslots, ix->dim, ix->dimp, b.p, s));
        ok = true;
    }
    *out = b.p;
    return NDB_B200_OK;
}

// nearest centroid per row (ivfinsert :906-935): sqrtf'd L2, strict <, lowest index
static int ivf_assign_dev(ndb_b200_ivf *ix, const float *d_rows, int64_t n, int *d_out, cudaStream_t s)
{
    return kmeans_assign_dev(ix->kw, d_rows, n, ix->dim, ix->nlists, NDB_L2, d_out, s);      // tensor cores when the centroid set is large
}

// probe lists per query into ix->probe ([nq][np] uint32, INVALID_SLOT = none)
static int ivf_coarse(ndb_b200_ivf *ix, const float *Q_dev, int nq, int np, int arith, cudaStream_t s)
{
    const int carith = arith == NDB_ARITH_FAST ? NDB_ARITH_FAST : NDB_ARITH_IVF_F32;
    int nparts = 1;
    NDB_CHECK(dense_scan(ix->kw.cstore.as<float>(), nullptr, ix->nlists, ix->dim, ix->dimp, NDB_L2, carith, Q_dev, nq, np,
                         ix->cscr, &nparts, s));
    NDB_CHECK(ix->probe.reserve((size_t) nq * np * 4));
    NDB_CHECK(ix->cdist.reserve((size_t) nq * np * 4));
    return launch_merge_parts(ix->cscr.pdist.as<float>(), ix->cscr.pslot.as<uint32_t>(), nullptr, nq, nparts, np,
                              ix->cdist.as<float>(), nullptr, ix->probe.as<uint32_t>(), s);
}

static int ivf_ready(ndb_b200_ivf *ix, cudaStream_t s)
{
    NDB_REQUIRE(ix->trained, NDB_B200_ESTATE, "ivf: index has no centroids (train or set_centroids first)");
    NDB_CHECK(ivf_layout(ix, s));
    return NDB_B200_OK;
}

void ivf_set_coarse_by_rank(ndb_b200_ivf *ix, bool on) { ix->coarse_by_rank = on; }

}  // namespace ndb

extern "C" {

int ndb_b200_ivf_create(int dim, int nlists, int metric, ndb_b200_ivf **out)
{
    NDB_CHECK(require_init());
    NDB_REQUIRE(out && dim > 0 && dim <= 16000, NDB_B200_EINVAL, "ivf_create: dim must be 1..16000");
    NDB_REQUIRE(nlists >= 1 && nlists <= 1000000, NDB_B200_EINVAL, "ivf_create: lists out of range");
    NDB_REQUIRE(metric >= NDB_L2 && metric <= NDB_IP, NDB_B200_EINVAL, "ivf_create: unknown metric %d", metric);
    ndb_b200_ivf *ix = new ndb_b200_ivf();
    ix->dim = dim;
    ix->dimp = round_up(dim, 4);
    ix->nlists = nlists;
    ix->metric = metric;
    *out = ix;
    return NDB_B200_OK;
}

void ndb_b200_ivf_free(ndb_b200_ivf *ix)
{
    if (!ix) return;
    if (ctx().initialized) {
        cudaSetDevice(ctx().device);
        cudaStreamSynchronize(ctx().stream);
        cudaStreamSynchronize(ctx().d2h_stream);
    }
    delete ix;
}

int ndb_b200_ivf_set_shard(ndb_b200_ivf *ix, int rank, int world)
{
    NDB_REQUIRE(ix && world >= 1 && rank >= 0 && rank < world, NDB_B200_EINVAL, "ivf_set_shard: bad rank/world");
    NDB_REQUIRE(ix->nrows == 0, NDB_B200_ESTATE, "ivf_set_shard: must be called before rows are inserted");
    ix->rank = rank;
    ix->world = world;
    return NDB_B200_OK;
}

static int ivf_install_centroids(ndb_b200_ivf *ix, cudaStream_t s)
{
    // kw.C holds the row-major centroids; build the IL32 copy used by every scan
    const int64_t blocks = (ix->nlists + 31) / 32;
    const size_t bytes = (size_t) blocks * 32 * ix->dimp * 4;
    NDB_CHECK(ix->kw.cstore.reserve(bytes));
    NDB_CUDA(cudaMemsetAsync(ix->kw.cstore.p, 0, bytes, s));
    NDB_CHECK(il32_scatter(ix->kw.C.as<float>(), ix->nlists, ix->dim, ix->dimp, nullptr, 0, ix->kw.cstore.as<float>(), s));
    ix->trained = true;
    ix->kw.cs.store_ok = false;
    return NDB_B200_OK;
}

int ndb_b200_ivf_train(ndb_b200_ivf *ix, const float *rows, int64_t n)
{
    NDB_CHECK(require_init());
    NDB_REQUIRE(ix && rows && n > 0, NDB_B200_EINVAL, "ivf_train: NULL or empty input");
    // maxSamples = Min(10000, nlists * 100), first rows in heap order (:580, :485-495)
    int64_t ns = (int64_t) ix->nlists * 100;
    if (ns > 10000) ns = 10000;
    if (ns > n) ns = n;
    // "ivf: not enough sample vectors" (:596-603)
    NDB_REQUIRE(ns >= ix->nlists, NDB_B200_ESTATE, "ivf: not enough sample vectors (%lld < %d)", (long long) ns, ix->nlists);
    NDB_REQUIRE(find_nonfinite(rows, ns * ix->dim) < 0, NDB_B200_EVECTOR, "ivf_train: NaN/Inf in samples");
    cudaStream_t s = ctx().stream;
    NDB_CHECK(ix->kw.X.reserve((size_t) ns * ix->dim * 4));
    NDB_CUDA(cudaMemcpyAsync(ix->kw.X.p, rows, (size_t) ns * ix->dim * 4, cudaMemcpyHostToDevice, s));
    int iters = 0;
    float cost = 0.0f;
    // IVF_MAX_ITERATIONS 50, IVF_CONVERGENCE_THRESHOLD 0.001 (:56-57)
    NDB_CHECK(kmeans_run_dev(ix->kw, (int) ns, ix->dim, ix->nlists, 50, 0.001f, &iters, &cost, s));
    ix->C_host.resize((size_t) ix->nlists * ix->dim);
    NDB_CUDA(cudaMemcpyAsync(ix->C_host.data(), ix->kw.C.p, ix->C_host.size() * 4, cudaMemcpyDeviceToHost, s));
    NDB_CUDA(cudaStreamSynchronize(s));
    return ivf_install_centroids(ix, s);
}

int ndb_b200_ivf_set_centroids(ndb_b200_ivf *ix, const float *C)
{
    NDB_CHECK(require_init());
    NDB_REQUIRE(ix && C, NDB_B200_EINVAL, "ivf_set_centroids: NULL input");
    ix->C_host.assign(C, C + (size_t) ix->nlists * ix->dim);
    cudaStream_t s = ctx().stream;
    NDB_CHECK(ivf_sync_centroids(ix, s));
    return ivf_install_centroids(ix, s);
}

int ndb_b200_ivf_get_centroids(const ndb_b200_ivf *ix, float *C)
{
    NDB_REQUIRE(ix && C && ix->trained, NDB_B200_ESTATE, "ivf_get_centroids: index not trained");
    memcpy(C, ix->C_host.data(), ix->C_host.size() * 4);
    return NDB_B200_OK;
}

int ndb_b200_ivf_assign(const ndb_b200_ivf *cix, const float *rows, int64_t n, int *out_list)
{
    ndb_b200_ivf *ix = const_cast<ndb_b200_ivf *>(cix);
    NDB_CHECK(require_init());
    NDB_REQUIRE(ix && rows && out_list && n > 0, NDB_B200_EINVAL, "ivf_assign: NULL or empty input");
    NDB_REQUIRE(ix->trained, NDB_B200_ESTATE, "ivf_assign: index not trained");
    cudaStream_t s = ctx().stream;
    const int64_t chunk = 1 << 20;
    for (int64_t off = 0; off < n; off += chunk) {
        const int64_t m = n - off < chunk ? n - off : chunk;
        NDB_CHECK(ix->tmp_rows.reserve((size_t) m * ix->dim * 4));
        NDB_CHECK(ix->tmp_assign.reserve((size_t) m * 4));
        NDB_CUDA(cudaMemcpyAsync(ix->tmp_rows.p, rows + (size_t) off * ix->dim, (size_t) m * ix->dim * 4, cudaMemcpyHostToDevice, s));
        NDB_CHECK(ivf_assign_dev(ix, ix->tmp_rows.as<float>(), m, ix->tmp_assign.as<int>(), s));
        NDB_CUDA(cudaMemcpyAsync(out_list + off, ix->tmp_assign.p, (size_t) m * 4, cudaMemcpyDeviceToHost, s));
        NDB_CUDA(cudaStreamSynchronize(s));
    }
    return NDB_B200_OK;
}

int ndb_b200_ivf_insert(ndb_b200_ivf *ix, const float *rows, const int64_t *ids, int64_t n, int *out_list)
{
    NDB_CHECK(require_init());
    NDB_REQUIRE(ix && rows && n > 0, NDB_B200_EINVAL, "ivf_insert: NULL or empty input");
    NDB_REQUIRE(ix->trained, NDB_B200_ESTATE, "ivf_insert: index has no centroids block");
    cudaStream_t s = ctx().stream;
    const int64_t chunk = 1 << 20;
    std::vector<int> assign;
    std::vector<uint32_t> keep;
    // default ids number every row ever passed to this handle, kept or not: the ranks of a list-sharded index see the
    // same rows in the same order, so they agree on the ids whatever each of them keeps
    const int64_t next_default_id = ix->inserted_total;
    // NaN / Inf rows are rejected (vector_distance.c:55-73) for the whole batch: the check runs on the device, on the
    // staged chunks, and a failure rolls the handle back to where it was
    const int64_t nrows0 = ix->nrows;
    const size_t nid0 = ix->row_id.size();
    const bool all_mine = ix->world == 1;
    if (all_mine) NDB_CHECK(ix->arena.grow((size_t) (ix->nrows + n) * ix->dim * 4, (size_t) ix->nrows * ix->dim * 4, s));
    unsigned long long *h_bad = reinterpret_cast<unsigned long long *>(static_cast<char *>(ctx().pinned) + 512);
    for (int64_t off = 0; off < n; off += chunk) {
        const int64_t m = n - off < chunk ? n - off : chunk;
        NDB_CHECK(ix->tmp_assign.reserve((size_t) m * 4));
        // one rank holds every list: the chunk is staged straight into the arena
        float *d_rows;
        if (all_mine) {
            d_rows = ix->arena.as<float>() + (size_t) ix->nrows * ix->dim;
        } else {
            NDB_CHECK(ix->tmp_rows.reserve((size_t) m * ix->dim * 4));
            d_rows = ix->tmp_rows.as<float>();
        }
        NDB_CUDA(cudaMemcpyAsync(d_rows, rows + (size_t) off * ix->dim, (size_t) m * ix->dim * 4, cudaMemcpyHostToDevice, s));
        NDB_CHECK(validate_into(d_rows, m * ix->dim, ctx().d_badidx, s));
        NDB_CUDA(cudaMemcpyAsync(h_bad, ctx().d_badidx, 8, cudaMemcpyDeviceToHost, s));
        NDB_CHECK(ivf_assign_dev(ix, d_rows, m, ix->tmp_assign.as<int>(), s));
        assign.resize(m);
        NDB_CUDA(cudaMemcpyAsync(assign.data(), ix->tmp_assign.p, (size_t) m * 4, cudaMemcpyDeviceToHost, s));
        NDB_CUDA(cudaStreamSynchronize(s));
        if (*h_bad != ~0ull) {
            ix->nrows = nrows0;
            ix->row_id.resize(nid0);
            ix->row_list.resize(nid0);
            set_error("ivf_insert: NaN/Inf in row %lld", (long long) (off + (int64_t) (*h_bad / (unsigned long long) ix->dim)));
            return NDB_B200_EVECTOR;
        }
        if (out_list) memcpy(out_list + off, assign.data(), (size_t) m * 4);
        if (all_mine) {
            ix->row_list.insert(ix->row_list.end(), assign.begin(), assign.end());
            const size_t base = ix->row_id.size();
            ix->row_id.resize(base + m);
            if (ids) memcpy(ix->row_id.data() + base, ids + off, (size_t) m * 8);
            else for (int64_t i = 0; i < m; i++) ix->row_id[base + i] = next_default_id + off + i;
            ix->nrows += m;
            ix->dirty = true;
            continue;
        }
        keep.clear();
        for (int64_t i = 0; i < m; i++) {
            if (assign[i] % ix->world != ix->rank) continue;      // list owned by another rank
            keep.push_back((uint32_t) i);
            ix->row_list.push_back(assign[i]);
            ix->row_id.push_back(ids ? ids[off + i] : next_default_id + off + i);
        }
        const int64_t nk = (int64_t) keep.size();
        if (nk) {
            NDB_CHECK(ix->arena.grow((size_t) (ix->nrows + nk) * ix->dim * 4, (size_t) ix->nrows * ix->dim * 4, s));
            NDB_CHECK(ix->tmp_keep.reserve((size_t) nk * 4));
            NDB_CUDA(cudaMemcpyAsync(ix->tmp_keep.p, keep.data(), (size_t) nk * 4, cudaMemcpyHostToDevice, s));
            copy_kept_rows_kernel<<<(unsigned) ((nk * ix->dim + 255) / 256), 256, 0, s>>>(
                ix->tmp_rows.as<float>(), ix->tmp_keep.as<uint32_t>(), nk, ix->dim, ix->arena.as<float>() + (size_t) ix->nrows * ix->dim);
            count_launch();
            NDB_CUDA(cudaGetLastError());
            NDB_CUDA(cudaStreamSynchronize(s));
            ix->nrows += nk;
            ix->dirty = true;
        }
    }
    if (ids) { for (int64_t i = 0; i < n; i++) ix->inserted_total = std::max(ix->inserted_total, ids[i] + 1); }
    else ix->inserted_total += n;
    return NDB_B200_OK;
}

int64_t ndb_b200_ivf_size(const ndb_b200_ivf *ix) { return ix ? ix->nrows : 0; }
int ndb_b200_ivf_dim(const ndb_b200_ivf *ix) { return ix ? ix->dim : 0; }

int ndb_b200_ivf_list_sizes(const ndb_b200_ivf *ix, int64_t *sizes)
{
    NDB_REQUIRE(ix && sizes, NDB_B200_EINVAL, "ivf_list_sizes: NULL input");
    for (int l = 0; l < ix->nlists; l++) sizes[l] = 0;
    for (int64_t r = 0; r < ix->nrows; r++) sizes[ix->row_list[r]]++;
    return NDB_B200_OK;
}

// Finish the build: lay the inserted rows out as lists and, for NDB_ARITH_TENSOR, make the blocked bf16 copy.
// The searches do this lazily on their first call; ivfbuild-time accounting calls it explicitly so that
// "index build seconds" contains all of it.
int ndb_b200_ivf_prepare(ndb_b200_ivf *ix, int arith)
{
    NDB_CHECK(require_init());
    NDB_REQUIRE(ix, NDB_B200_EINVAL, "ivf_prepare: NULL index");
    cudaStream_t s = ctx().stream;
    NDB_CHECK(ivf_ready(ix, s));
    if (arith == NDB_ARITH_TENSOR && ix->dim <= TC_MAX_DIM) NDB_CHECK(ivf_tensor_ready(ix, s));
    NDB_CUDA(cudaStreamSynchronize(s));
    return NDB_B200_OK;
}

// ivf_knn_search_gpu(index_name, query, k, nprobe default 10) (src/gpu/common/gpu_sql.c:931-1456): the SQL function's
// argument checks (:972-991) and its result, (heap id, distance) rows in ascending distance, over the resident index.
// The reference's version evaluates one candidate per launch_l2_distance call (:1278-1293); here it is one search.
int ndb_b200_ivf_knn_search_gpu(ndb_b200_ivf *ix, const float *query, int dim, int k, int nprobe, int64_t *ids, float *dist,
                                int *nresults)
{
    NDB_CHECK(require_init());
    NDB_REQUIRE(ix, NDB_B200_EINVAL, "ivf_knn_search_gpu: index cannot be NULL");
    NDB_REQUIRE(query && ids && dist && nresults, NDB_B200_EINVAL, "ivf_knn_search_gpu: query vector cannot be NULL");
    NDB_REQUIRE(k > 0 && k <= 10000, NDB_B200_EINVAL, "ivf_knn_search_gpu: k must be between 1 and 10000");
    NDB_REQUIRE(nprobe > 0 && nprobe <= 1000, NDB_B200_EINVAL, "ivf_knn_search_gpu: nprobe must be between 1 and 1000");
    NDB_REQUIRE(dim == ix->dim, NDB_B200_EDIM, "ivf_knn_search_gpu: query dimension %d does not match index dimension %d", dim, ix->dim);
    NDB_REQUIRE(k <= 128, NDB_B200_EINVAL, "ivf_knn_search_gpu: k > 128 is not supported by this library (k = %d)", k);
    const int np = nprobe < ix->nlists ? nprobe : ix->nlists;
    NDB_REQUIRE(np <= 128, NDB_B200_EINVAL, "ivf_knn_search_gpu: nprobe > 128 is not supported by this library (nprobe = %d)", nprobe);
    NDB_CHECK(ndb_b200_ivf_search(ix, query, 1, np, k, NDB_IVF_FULL, NDB_ARITH_IVF_F32, dist, ids));
    int n = 0;
    while (n < k && ids[n] >= 0) n++;
    *nresults = n;
    return NDB_B200_OK;
}

// certified selection of the last NDB_ARITH_TENSOR search: out[0] = queries whose list scan went to the exact
// kernel, out[1] = exact re-evaluations of list candidates, out[2] / out[3] = the same for the coarse quantiser,
// out[4] = queries whose every probed row had to be evaluated, out[5] = rows evaluated by segment rescans
int ndb_b200_ivf_cert_stats(ndb_b200_ivf *ix, int64_t *out)
{
    NDB_CHECK(require_init());
    NDB_REQUIRE(ix && out, NDB_B200_EINVAL, "ivf_cert_stats: NULL argument");
    for (int i = 0; i < 6; i++) out[i] = 0;
    if (!ix->cert_counters.p) return NDB_B200_OK;
    unsigned long long h[6];
    NDB_CUDA(cudaStreamSynchronize(ctx().stream));
    NDB_CUDA(cudaMemcpy(h, ix->cert_counters.p, sizeof(h), cudaMemcpyDeviceToHost));
    for (int i = 0; i < 6; i++) out[i] = (int64_t) h[i];
    if (getenv("NDB_CERT_DEBUG") && ix->cert_dbg.p) {
        float r[16 * 8];
        NDB_CUDA(cudaMemcpy(r, ix->cert_dbg.p, sizeof(r), cudaMemcpyDeviceToHost));
        for (int i = 0; i < 16 && i < (int) h[4]; i++)
            fprintf(stderr, "cert full scan: q=%d R=%g tau=%g bound(R)=%g rescanned=%d ||eq||=%g ||q||=%g tau0=%g\n", (int) r[i * 8], r[i * 8 + 1],
                    r[i * 8 + 2], r[i * 8 + 3], (int) r[i * 8 + 4], r[i * 8 + 5], r[i * 8 + 6], r[i * 8 + 7]);
    }
    return NDB_B200_OK;
}

int ndb_b200_ivf_select_clusters(ndb_b200_ivf *ix, const float *Q, int nq, int nprobe, int *probes)
{
    NDB_CHECK(require_init());
    NDB_REQUIRE(ix && Q && probes && nq > 0 && nprobe > 0, NDB_B200_EINVAL, "ivf_select_clusters: bad argument");
    NDB_REQUIRE(ix->trained, NDB_B200_ESTATE, "ivf_select_clusters: index not trained");
    cudaStream_t s = ctx().stream;
    const int np = nprobe < ix->nlists ? nprobe : ix->nlists;
    NDB_REQUIRE(np <= 128, NDB_B200_EINVAL, "ivf: nprobe > 128 is not supported");
    NDB_CHECK(ix->qbuf.reserve((size_t) nq * ix->dim * 4));
    NDB_CUDA(cudaMemcpyAsync(ix->qbuf.p, Q, (size_t) nq * ix->dim * 4, cudaMemcpyHostToDevice, s));
    NDB_CHECK(ivf_coarse(ix, ix->qbuf.as<float>(), nq, np, NDB_ARITH_IVF_F32, s));
    std::vector<uint32_t> h((size_t) nq * np);
    NDB_CUDA(cudaMemcpyAsync(h.data(), ix->probe.p, h.size() * 4, cudaMemcpyDeviceToHost, s));
    NDB_CUDA(cudaStreamSynchronize(s));
    for (int q = 0; q < nq; q++)
        for (int i = 0; i < nprobe; i++)
            probes[(size_t) q * nprobe + i] = (i < np && h[(size_t) q * np + i] != INVALID_SLOT) ? (int) h[(size_t) q * np + i] : -1;
    return NDB_B200_OK;
}

int ndb_b200_ivf_search_dev(ndb_b200_ivf *ix, const float *Q_dev, int nq, int nprobe, int k, int mode, int arith,
                            float *dist_dev, int64_t *ids_dev, void *stream)
{
    NDB_CHECK(require_init());
    NDB_REQUIRE(ix && Q_dev && dist_dev && ids_dev && nq > 0 && nprobe > 0, NDB_B200_EINVAL, "ivf_search: bad argument");
    NDB_REQUIRE(k >= 1 && k <= 128, NDB_B200_EINVAL, "ivf_search: k=%d out of range 1..128", k);
    NDB_REQUIRE(arith == NDB_ARITH_IVF_F32 || arith == NDB_ARITH_FAST || arith == NDB_ARITH_TENSOR, NDB_B200_EINVAL,
                "ivf_search: arith %d unsupported", arith);
    cudaStream_t s = stream ? (cudaStream_t) stream : ctx().stream;
    NDB_CHECK(ivf_ready(ix, s));
    const int np = nprobe < ix->nlists ? nprobe : ix->nlists;
    NDB_REQUIRE(np <= 128, NDB_B200_EINVAL, "ivf: nprobe > 128 is not supported");
    if (arith == NDB_ARITH_TENSOR) {
        NDB_REQUIRE(mode != NDB_IVF_LITERAL, NDB_B200_EINVAL, "ivf_search: the literal mode has no tensor variant");
        return ivf_search_tensor(ix, Q_dev, nq, np, k, dist_dev, ids_dev, s);
    }
    const int L = ix->nlists;
    const int64_t npairs = (int64_t) nq * np;

    // 1. ivfSelectClusters
    NDB_CHECK(ivf_coarse(ix, Q_dev, nq, np, arith, s));
    const void *vnorm = nullptr;
    if (ix->metric == NDB_COSINE) NDB_CHECK(ivf_vnorm(ix, arith, &vnorm, s));

    if (mode == NDB_IVF_LITERAL) {
        NDB_REQUIRE(k * 10 <= 2560, NDB_B200_EINVAL, "ivf literal mode: k too large");
        NDB_REQUIRE(arith != NDB_ARITH_FAST, NDB_B200_EINVAL, "ivf literal mode reproduces ivfCollectCandidates: arith must be NDB_ARITH_IVF_F32");
        const size_t smem = (size_t) 12 * 4 * k * 10;                 // k = 128: 61 440 bytes, above the 48 KB default
        NDB_REQUIRE(smem <= ctx().smem_optin, NDB_B200_EINVAL, "ivf literal mode: k=%d needs %zu bytes of shared memory", k, smem);
        const unsigned grid = (unsigned) ((nq + 3) / 4);
#define NDB_LIT(M) do { \
        static uint64_t cfg = 0; \
        if (cfg != ctx().generation) { \
            NDB_CUDA(cudaFuncSetAttribute(ivf_literal_kernel<Arith<M, NDB_ARITH_IVF_F32>>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) (12 * 4 * 1280))); \
            cfg = ctx().generation; \
        } \
        ivf_literal_kernel<Arith<M, NDB_ARITH_IVF_F32>><<<grid, 128, smem, s>>>( \
        reinterpret_cast<const float4 *>(ix->store.ptr()), (const float *) vnorm, Q_dev, ix->probe.as<uint32_t>(), \
        ix->d_list_len.as<uint32_t>(), ix->d_list_blk.as<uint32_t>(), ix->lit_order.as<uint32_t>(), ix->ids.as<int64_t>(), \
        nq, np, L, ix->dim, ix->dimp, k, dist_dev, ids_dev); } while (0)
        if (ix->metric == NDB_L2) NDB_LIT(NDB_L2);
        else if (ix->metric == NDB_COSINE) NDB_LIT(NDB_COSINE);
        else NDB_LIT(NDB_IP);
#undef NDB_LIT
        count_launch();
        NDB_CUDA(cudaGetLastError());
        return NDB_B200_OK;
    }

    // 2. bucket pairs by list
    const int qt = scan_pick_qt(arith, ix->dim, k);
    NDB_REQUIRE(qt > 0, NDB_B200_EINVAL, "ivf_search: unsupported shape dim=%d k=%d", ix->dim, k);
    NDB_CHECK(ix->cnt.reserve((size_t) L * 4 * 2));
    NDB_CHECK(ix->qoff.reserve((size_t) L * 4));
    NDB_CHECK(ix->item_off.reserve((size_t) L * 4));
    NDB_CHECK(ix->qmap.reserve((size_t) npairs * 4));
    NDB_CHECK(ix->pairpos.reserve((size_t) npairs * 4));
    NDB_CHECK(ix->nitems.reserve(64));
    NDB_CHECK(ix->stats.reserve(64));
    uint32_t *cnt = ix->cnt.as<uint32_t>(), *fill = cnt + L;
    NDB_CUDA(cudaMemsetAsync(cnt, 0, (size_t) L * 4 * 2, s));
    ivf_hist_kernel<<<(unsigned) ((npairs + 255) / 256), 256, 0, s>>>(ix->probe.as<uint32_t>(), npairs, ix->d_list_len.as<uint32_t>(), L, cnt,
                                                                     (uint32_t) np, 0u);
    const uint32_t segb = ivf_seg_blocks();
    ivf_offsets_kernel<<<1, 1024, 0, s>>>(cnt, ix->d_list_len.as<uint32_t>(), ix->d_list_order.as<uint32_t>(), L, qt, segb, 1u, 0u,
                                          ix->qoff.as<uint32_t>(), ix->item_off.as<uint32_t>(), ix->nitems.as<uint32_t>(),
                                          ix->stats.as<unsigned long long>(), 0u);
    // the number of work items depends on how the batch's probes fall on long and short lists:
    // one 4-byte read-back sizes the item and partial-result buffers exactly
    uint32_t *h_nitems = reinterpret_cast<uint32_t *>(ctx().pinned);
    NDB_CUDA(cudaMemcpyAsync(h_nitems, ix->nitems.p, 4, cudaMemcpyDeviceToHost, s));
    ivf_scatter_kernel<<<(unsigned) ((npairs + 255) / 256), 256, 0, s>>>(ix->probe.as<uint32_t>(), npairs, ix->d_list_len.as<uint32_t>(), L,
                                                                        ix->qoff.as<uint32_t>(), fill, ix->qmap.as<uint32_t>(),
                                                                        ix->pairpos.as<uint32_t>(), nullptr, 0u, (uint32_t) np, 0u);
    count_launch(3);
    NDB_CUDA(cudaGetLastError());
    NDB_CUDA(cudaStreamSynchronize(s));
    const size_t n_items = *h_nitems;
    NDB_CHECK(ix->items.reserve((n_items + 1) * sizeof(WorkItem)));
    ivf_items_kernel<<<(unsigned) ((L + 3) / 4), 128, 0, s>>>(cnt, ix->qoff.as<uint32_t>(), ix->item_off.as<uint32_t>(),
                                                             ix->d_list_len.as<uint32_t>(), ix->d_list_blk.as<uint32_t>(), L, qt, segb,
                                                             ix->items.as<WorkItem>());
    count_launch();
    NDB_CUDA(cudaGetLastError());

    // 3. list scan with fused top-k
    NDB_CHECK(ix->scr.ensure((n_items + 1) * qt, k, nq, ix->metric == NDB_COSINE ? 4 : 0));
    if (ix->metric == NDB_COSINE) NDB_CHECK(row_norms(arith, Q_dev, nq, ix->dim, ix->scr.qnorm.p, s));
    ScanParams p;
    memset(&p, 0, sizeof(p));
    p.vecs = reinterpret_cast<const float4 *>(ix->store.ptr());
    p.vnorm = vnorm;
    p.Q = Q_dev;
    p.qnorm = ix->metric == NDB_COSINE ? ix->scr.qnorm.p : nullptr;
    p.dim = ix->dim; p.dimp = ix->dimp; p.k = k;
    p.items = ix->items.as<WorkItem>();
    p.n_items_ptr = ix->nitems.as<uint32_t>();
    p.qmap = ix->qmap.as<uint32_t>();
    p.nprobe = (uint32_t) np;
    p.counter = ix->scr.counter.as<uint32_t>();
    p.pdist = ix->scr.pdist.as<float>();
    p.pslot = ix->scr.pslot.as<uint32_t>();
    Context &c = ctx();
    if (c.timing) NDB_CUDA(cudaEventRecord(c.ev0, s));
    NDB_CHECK(launch_scan(ix->metric, arith, qt, p, (uint32_t) n_items, s));
    if (c.timing) {
        NDB_CUDA(cudaEventRecord(c.ev1, s));
        c.last_ms = -1.0;
        c.last_bytes = -1.0;          // resolved from ix->stats in last_kernel_stats
        c.last_evals = -1;
        c.stats_src = ix->stats.p;
        c.stats_dim = ix->dim;
    }

    // 4. merge the probed lists' partial results by (dist, id)
    const unsigned mgrid = (unsigned) ((nq + 3) / 4);
    if (k <= 32)
        ivf_merge_kernel<1><<<mgrid, 128, 0, s>>>(ix->scr.pdist.as<float>(), ix->scr.pslot.as<uint32_t>(), ix->ids.as<int64_t>(),
                                                  ix->probe.as<uint32_t>(), ix->pairpos.as<uint32_t>(), ix->item_off.as<uint32_t>(),
                                                  ix->d_list_len.as<uint32_t>(), nq, np, L, qt, segb, k, dist_dev, ids_dev);
    else
        ivf_merge_kernel<4><<<mgrid, 128, 0, s>>>(ix->scr.pdist.as<float>(), ix->scr.pslot.as<uint32_t>(), ix->ids.as<int64_t>(),
                                                  ix->probe.as<uint32_t>(), ix->pairpos.as<uint32_t>(), ix->item_off.as<uint32_t>(),
                                                  ix->d_list_len.as<uint32_t>(), nq, np, L, qt, segb, k, dist_dev, ids_dev);
    count_launch();
    NDB_CUDA(cudaGetLastError());
    return NDB_B200_OK;
}

int ndb_b200_ivf_search(ndb_b200_ivf *ix, const float *Q, int nq, int nprobe, int k, int mode, int arith, float *dist,
                        int64_t *ids)
{
    NDB_CHECK(require_init());
    NDB_REQUIRE(ix && Q && dist && ids && nq > 0, NDB_B200_EINVAL, "ivf_search: NULL or empty input");
    cudaStream_t s = ctx().stream;
    const size_t qb = (size_t) nq * ix->dim * 4, m = (size_t) nq * k;
    NDB_CHECK(ix->qbuf.reserve(qb));
    NDB_CHECK(ix->outd.reserve(m * 4));
    NDB_CHECK(ix->outi.reserve(m * 8));
    NDB_CUDA(cudaMemcpyAsync(ix->qbuf.p, Q, qb, cudaMemcpyHostToDevice, s));
    NDB_CHECK(validate_begin(ix->qbuf.as<float>(), (int64_t) nq * ix->dim, s));
    NDB_CHECK(ndb_b200_ivf_search_dev(ix, ix->qbuf.as<float>(), nq, nprobe, k, mode, arith, ix->outd.as<float>(),
                                      ix->outi.as<int64_t>(), s));
    NDB_CUDA(cudaMemcpyAsync(dist, ix->outd.p, m * 4, cudaMemcpyDeviceToHost, s));
    NDB_CUDA(cudaMemcpyAsync(ids, ix->outi.p, m * 8, cudaMemcpyDeviceToHost, s));
    NDB_CUDA(cudaStreamSynchronize(s));
    NDB_REQUIRE(validate_end() < 0, NDB_B200_EVECTOR, "ivf_search: NaN/Inf in query");
    return NDB_B200_OK;
}

// ---- pipelined form ----------------------------------------------------------------------------
// search_begin queues one batch and returns; search_end blocks until that batch's results are in the
// caller's buffers.  Up to two batches per index may be in flight: the H2D copy of batch i+1 (its own
// stream) and the D2H copy of batch i-1 (another) overlap the kernels of batch i, which still run one
// batch at a time on the library stream (the index's scratch belongs to one search at a time).
// Q, dist and ids must stay valid until search_end; pinned buffers make the copies true DMA.
int ndb_b200_ivf_search_begin(ndb_b200_ivf *ix, const float *Q, int nq, int nprobe, int k, int mode, int arith, float *dist,
                              int64_t *ids, int *ticket)
{
    NDB_CHECK(require_init());
    NDB_REQUIRE(ix && Q && dist && ids && ticket && nq > 0, NDB_B200_EINVAL, "ivf_search_begin: NULL or empty input");
    int t = -1;
    for (int i = 0; i < 2; i++) if (!ix->slot[i].busy) { t = i; break; }
    NDB_REQUIRE(t >= 0, NDB_B200_ESTATE, "ivf_search_begin: two batches already in flight (call ivf_search_end first)");
    ndb_b200_ivf::Slot &sl = ix->slot[t];
    Context &c = ctx();
    if (!sl.h2d) {
        NDB_CUDA(cudaEventCreateWithFlags(&sl.h2d, cudaEventDisableTiming));
        NDB_CUDA(cudaEventCreateWithFlags(&sl.done, cudaEventDisableTiming));
        NDB_CUDA(cudaEventCreateWithFlags(&sl.d2h, cudaEventDisableTiming));
        NDB_CUDA(cudaMallocHost(&sl.h_bad, 8));
        NDB_CHECK(sl.bad.reserve(8));
    }
    const size_t qb = (size_t) nq * ix->dim * 4, m = (size_t) nq * k;
    NDB_CHECK(sl.q.reserve(qb));
    NDB_CHECK(sl.d.reserve(m * 4));
    NDB_CHECK(sl.i.reserve(m * 8));
    *sl.h_bad = ~0ull;
    NDB_CUDA(cudaMemcpyAsync(sl.q.p, Q, qb, cudaMemcpyHostToDevice, c.h2d_stream));
    NDB_CUDA(cudaEventRecord(sl.h2d, c.h2d_stream));
    NDB_CUDA(cudaStreamWaitEvent(c.stream, sl.h2d, 0));
    NDB_CHECK(validate_into(sl.q.as<float>(), (int64_t) nq * ix->dim, sl.bad.as<unsigned long long>(), c.stream));
    NDB_CHECK(ndb_b200_ivf_search_dev(ix, sl.q.as<float>(), nq, nprobe, k, mode, arith, sl.d.as<float>(), sl.i.as<int64_t>(), c.stream));
    NDB_CUDA(cudaEventRecord(sl.done, c.stream));
    NDB_CUDA(cudaStreamWaitEvent(c.d2h_stream, sl.done, 0));
    NDB_CUDA(cudaMemcpyAsync(dist, sl.d.p, m * 4, cudaMemcpyDeviceToHost, c.d2h_stream));
    NDB_CUDA(cudaMemcpyAsync(ids, sl.i.p, m * 8, cudaMemcpyDeviceToHost, c.d2h_stream));
    NDB_CUDA(cudaMemcpyAsync(sl.h_bad, sl.bad.p, 8, cudaMemcpyDeviceToHost, c.d2h_stream));
    NDB_CUDA(cudaEventRecord(sl.d2h, c.d2h_stream));
    sl.busy = true;
    *ticket = t;
    return NDB_B200_OK;
}

int ndb_b200_ivf_search_end(ndb_b200_ivf *ix, int ticket)
{
    NDB_CHECK(require_init());
    NDB_REQUIRE(ix && ticket >= 0 && ticket < 2 && ix->slot[ticket].busy, NDB_B200_EINVAL, "ivf_search_end: no such batch in flight");
    ndb_b200_ivf::Slot &sl = ix->slot[ticket];
    sl.busy = false;
    NDB_CUDA(cudaEventSynchronize(sl.d2h));
    NDB_REQUIRE(*sl.h_bad == ~0ull, NDB_B200_EVECTOR, "ivf_search: NaN/Inf in query");
    return NDB_B200_OK;
}

// ---- relation loader: meta page + centroid page(s) + inverted-list page chains ----------------
// The host walks headers, line pointers and chain links (a few bytes per page); the vector
// payloads -- the bulk -- are pulled out of the staged pages on the device.
int ndb_b200_ivf_load_relation(ndb_b200_ivf *ix, const void *blocks, uint32_t nblocks)
{
    using namespace ndb::pg;
    NDB_CHECK(require_init());
    NDB_REQUIRE(ix && blocks && nblocks >= 2, NDB_B200_EINVAL, "ivf_load_relation: need at least the meta and centroid blocks");
    NDB_REQUIRE(ix->nrows == 0, NDB_B200_ESTATE, "ivf_load_relation: index already holds rows");
    IvfMeta meta;
    memcpy(&meta, page_at(blocks, 0) + PAGE_HEADER, sizeof(meta));
    NDB_REQUIRE(meta.magic == IVF_MAGIC, NDB_B200_EINVAL, "ivf_load_relation: bad magic 0x%08x in the meta page", meta.magic);
    NDB_REQUIRE(meta.nlists == ix->nlists, NDB_B200_EINVAL, "ivf_load_relation: relation has %d lists, handle has %d", meta.nlists, ix->nlists);
    NDB_REQUIRE(meta.dim == ix->dim || meta.insertedVectors == 0, NDB_B200_EDIM, "ivf_load_relation: relation dim %d, handle dim %d", meta.dim, ix->dim);
    NDB_REQUIRE(meta.centroidsBlock != INVALID_BLOCK && meta.centroidsBlock < nblocks, NDB_B200_ESTATE, "ivf_load_relation: no centroids block");

    // centroid items: IvfCentroidData + float4[dim]; one page in the reference (Q7), consecutive
    // pages accepted here
    const int L = ix->nlists, dim = ix->dim;
    std::vector<float> C((size_t) L * dim, 0.0f);
    std::vector<uint32_t> first(L, INVALID_BLOCK);
    std::vector<uint8_t> have_list(L, 0);
    int got = 0;
    for (uint32_t b = meta.centroidsBlock; b < nblocks && got < L; b++) {
        const uint8_t *page = page_at(blocks, b);
        if (special_offset(page) != BLCKSZ - 24) break;            // not a centroid page
        const int maxoff = max_offset(page);
        for (int off = 1; off <= maxoff && got < L; off++) {
            uint32_t lo = 0, len = 0;
            const int st = checked_item(page, off, 24 + (uint32_t) dim * 4, &lo, &len);
            if (st == 0) continue;
            NDB_REQUIRE(st == 1, NDB_B200_EINVAL, "ivf_load_relation: centroid item %d/%d lies outside its page or is too short", (int) b, off);
            IvfCentroidHdr h;
            memcpy(&h, page + lo, sizeof(h));
            NDB_REQUIRE(h.listId >= 0 && h.listId < L && h.dim == dim, NDB_B200_EINVAL,
                        "ivf_load_relation: centroid item %d/%d is inconsistent (list %d dim %d)", (int) b, off, h.listId, h.dim);
            NDB_REQUIRE(!have_list[h.listId], NDB_B200_EINVAL, "ivf_load_relation: list %d has two centroid items", h.listId);
            have_list[h.listId] = 1;
            memcpy(&C[(size_t) h.listId * dim], page + lo + 24, (size_t) dim * 4);
            first[h.listId] = h.firstBlock;
            got++;
        }
    }
    NDB_REQUIRE(got == L, NDB_B200_EINVAL, "ivf_load_relation: found %d of %d centroids", got, L);
    NDB_CHECK(ndb_b200_ivf_set_centroids(ix, C.data()));

    // walk every chain: page order = insertion order inside a list (ivf_am.c:1793-1840)
    std::vector<uint64_t> src;           // byte offset of each live entry's vector payload
    std::vector<int64_t> new_id;         // committed to the index only when the whole relation has been read
    std::vector<int32_t> new_list;
    std::vector<uint8_t> seen(nblocks, 0);
    for (int l = 0; l < L; l++) {
        if (l % ix->world != ix->rank) continue;
        for (uint32_t b = first[l]; b != INVALID_BLOCK;) {
            NDB_REQUIRE(b < nblocks && !seen[b], NDB_B200_EINVAL, "ivf_load_relation: list %d chain is corrupt at block %u", l, b);
            seen[b] = 1;
            const uint8_t *page = page_at(blocks, b);
            NDB_REQUIRE(special_offset(page) == BLCKSZ - 8, NDB_B200_EINVAL, "ivf_load_relation: block %u is not a list page", b);
            const int maxoff = max_offset(page);
            for (int off = 1; off <= maxoff; off++) {
                uint32_t lo = 0, len = 0;
                const int st = checked_item(page, off, IVF_ENTRY_HDR, &lo, &len);
                if (st == 0) continue;                               // unused or LP_DEAD (:1816)
                NDB_REQUIRE(st == 1, NDB_B200_EINVAL, "ivf_load_relation: entry %u/%d lies outside its page", b, off);
                int16_t edim;
                memcpy(&edim, page + lo + 6, 2);
                if (edim != dim) continue;                           // :1821-1822
                NDB_REQUIRE(len >= IVF_ENTRY_HDR + (uint32_t) dim * 4, NDB_B200_EINVAL, "ivf_load_relation: entry %u/%d is truncated", b, off);
                new_id.push_back(tid_unpack(page + lo));
                new_list.push_back(l);
                src.push_back((uint64_t) b * BLCKSZ + lo + IVF_ENTRY_HDR);
            }
            IvfListSpecial sp;
            memcpy(&sp, page + BLCKSZ - 8, sizeof(sp));
            b = sp.nextBlock;
        }
    }
    const int64_t n = (int64_t) src.size();
    if (n == 0) { ix->dirty = true; return NDB_B200_OK; }
    cudaStream_t s = ctx().stream;
    DevBuf d_pages, d_src;
    NDB_CHECK(d_pages.reserve((size_t) nblocks * BLCKSZ));
    NDB_CHECK(d_src.reserve((size_t) n * 8));
    NDB_CHECK(ix->arena.reserve((size_t) n * dim * 4));
    NDB_CUDA(cudaMemcpyAsync(d_pages.p, blocks, (size_t) nblocks * BLCKSZ, cudaMemcpyHostToDevice, s));
    NDB_CUDA(cudaMemcpyAsync(d_src.p, src.data(), (size_t) n * 8, cudaMemcpyHostToDevice, s));
    extract_payload_kernel<<<(unsigned) ((n * dim + 255) / 256), 256, 0, s>>>(d_pages.as<unsigned char>(), d_src.as<unsigned long long>(), n, dim,
                                                                             ix->arena.as<float>());
    count_launch();
    NDB_CUDA(cudaGetLastError());
    NDB_CUDA(cudaStreamSynchronize(s));
    for (int64_t v : new_id) ix->inserted_total = std::max(ix->inserted_total, v + 1);
    ix->row_id = std::move(new_id);
    ix->row_list = std::move(new_list);
    ix->nrows = n;
    ix->dirty = true;
    return NDB_B200_OK;
}

}  // extern "C"
