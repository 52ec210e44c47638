// scan.cuh -- the streaming distance scan with fused warp-level top-k.
//
// One kernel serves every "many queries x many stored vectors -> k smallest" step of the path:
//   * exact kNN over a dataset (SeqScan + top-N sort)                 dense mode
//   * ivfSelectClusters: queries x centroids, k = nprobe (ivf_am.c:1597-1717)      dense mode
//   * ivfinsert / kmeans_assign: rows x centroids, k = 1 (:906-935, :2164-2177)    dense mode
//   * ivfCollectCandidates: (query, probed list) pairs grouped by list (:1722-1909)  list mode
//
// Work item = (a run of 32-vector blocks, a tile of QT queries).  A CTA of NW warps takes an
// item; warp w streams blocks w, w+NW, ... with one 128-bit load per lane per 4 dimensions
// (IL32 layout, 512 contiguous bytes per warp-load).  Lane l owns vector l of the block and
// walks its dimensions in order with QT accumulators -- the query tile sits in shared memory
// and is read with broadcast LDS.128 -- so every distance is produced by exactly the
// reference's sequential loop (see arith.cuh) while each stored vector is fetched once per QT
// queries.  No distance matrix is written: each warp keeps a sorted top-k per query in
// registers (WarpTopK), filters new distances against the k-th best with one ballot, merges
// the NW per-warp lists through shared memory and writes k (dist, slot) pairs per
// (query, item-part).  A second small kernel merges the parts per query by (dist, id).
#pragma once
#include "arith.cuh"

namespace ndb {

struct WorkItem {
    uint32_t blk_begin;   // first 32-vector block
    uint32_t nvec;        // valid vectors from blk_begin*32 on
    uint32_t qoff;        // first entry of the query tile (index into qmap, or query index)
    uint32_t nq;          // queries in this tile (<= QT)
    uint32_t part;        // dense mode: which partial slot of the query this item fills
};

struct ScanParams {
    const float4 *vecs;        // IL32 store
    const void *vnorm;         // per-slot norm accumulators (Arith::N) or nullptr
    const float *Q;            // [nq_total][dim] row-major
    const void *qnorm;         // per-query norm accumulators or nullptr
    int dim, dimp, k;
    // list mode
    const WorkItem *items;     // nullptr => dense mode
    const uint32_t *n_items_ptr;
    const uint32_t *qmap;      // entry -> pair index p ; query = p / nprobe ; partial = p
    uint32_t nprobe;
    // dense mode: item i -> seg = i / ntiles, tile = i % ntiles
    uint32_t dense_items, dense_ntiles, dense_seg_blocks, dense_nq, dense_nparts;
    uint64_t dense_nvec;
    uint32_t *counter;         // persistent-CTA work counter (zeroed before launch)
    float *pdist;              // [partial][k]
    uint32_t *pslot;
};

constexpr int SCAN_NW = 8;     // warps per CTA

template <class QE> __device__ __forceinline__ void load_q4(const QE *q, QE &q0, QE &q1, QE &q2, QE &q3);
template <> __device__ __forceinline__ void load_q4<float>(const float *q, float &q0, float &q1, float &q2, float &q3)
{
    const float4 t = *reinterpret_cast<const float4 *>(q);
    q0 = t.x; q1 = t.y; q2 = t.z; q3 = t.w;
}
template <> __device__ __forceinline__ void load_q4<double>(const double *q, double &q0, double &q1, double &q2, double &q3)
{
    const double2 t0 = *reinterpret_cast<const double2 *>(q);
    const double2 t1 = *reinterpret_cast<const double2 *>(q + 2);
    q0 = t0.x; q1 = t0.y; q2 = t1.x; q3 = t1.y;
}

template <class P, int QT, int KR>
__global__ void __launch_bounds__(SCAN_NW * 32) scan_topk_kernel(const ScanParams prm)
{
    using QE = typename P::Q;
    using NT = typename P::N;
    using Acc = typename P::Acc;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    QE *qs = reinterpret_cast<QE *>(smem_raw);                       // [QT][dimp]
    // cross-warp merge area after the query tile: [NW][QT][KR*32] (dist, slot)
    float *md = reinterpret_cast<float *>(smem_raw + sizeof(QE) * (size_t) QT * prm.dimp);
    uint32_t *ms = reinterpret_cast<uint32_t *>(md + SCAN_NW * QT * KR * 32);
    __shared__ WorkItem s_item;
    __shared__ uint32_t s_qidx[QT];     // global query index per tile entry
    __shared__ uint32_t s_pidx[QT];     // partial-result index per tile entry
    __shared__ NT s_qnorm[QT];

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int dimp = prm.dimp, dim = prm.dim, k = prm.k;
    const int nfull = dim >> 2, rem = dim & 3;

    for (;;) {
        // ---- fetch the next work item ---------------------------------------------------
        if (tid == 0) {
            uint32_t idx = atomicAdd(prm.counter, 1u);
            WorkItem it;
            if (prm.items) {
                uint32_t n = *prm.n_items_ptr;
                if (idx < n) it = prm.items[idx];
                else it.nq = 0;
            } else if (idx < prm.dense_items) {
                uint32_t seg = idx / prm.dense_ntiles, tile = idx % prm.dense_ntiles;
                uint64_t v0 = (uint64_t) seg * prm.dense_seg_blocks * 32;
                uint64_t left = prm.dense_nvec - v0;
                uint64_t cap = (uint64_t) prm.dense_seg_blocks * 32;
                it.blk_begin = seg * prm.dense_seg_blocks;
                it.nvec = (uint32_t) (left < cap ? left : cap);
                it.qoff = tile * QT;
                uint32_t ql = prm.dense_nq - it.qoff;
                it.nq = ql < (uint32_t) QT ? ql : (uint32_t) QT;
                it.part = seg;
            } else {
                it.nq = 0;
            }
            s_item = it;
        }
        __syncthreads();
        const WorkItem item = s_item;
        if (item.nq == 0) break;

        // ---- stage the query tile ---------------------------------------------------------
        if (tid < QT) {
            uint32_t qi = 0, pi = 0;
            if (tid < (int) item.nq) {
                if (prm.qmap) {
                    uint32_t p = prm.qmap[item.qoff + tid];
                    qi = p / prm.nprobe;
                    pi = p;
                } else {
                    qi = item.qoff + tid;
                    pi = qi * prm.dense_nparts + item.part;
                }
            }
            s_qidx[tid] = qi;
            s_pidx[tid] = pi;
            if (P::NORMS) s_qnorm[tid] = tid < (int) item.nq ? reinterpret_cast<const NT *>(prm.qnorm)[qi] : NT(0);
        }
        __syncthreads();
        for (int i = tid; i < QT * dimp; i += SCAN_NW * 32) {
            int qi = i / dimp, dd = i - qi * dimp;
            float v = 0.0f;
            if (qi < (int) item.nq && dd < dim) v = prm.Q[(size_t) s_qidx[qi] * dim + dd];
            qs[i] = (QE) v;
        }
        __syncthreads();

        WarpTopK<KR, uint32_t> top[QT];
#pragma unroll
        for (int qi = 0; qi < QT; qi++) top[qi].init();

        // ---- stream the blocks ------------------------------------------------------------
        const uint32_t nblk = (item.nvec + 31) >> 5;
        for (uint32_t b = warp; b < nblk; b += SCAN_NW) {
            const uint32_t slot = (item.blk_begin + b) * 32 + lane;
            const bool valid = b * 32 + lane < item.nvec;
            const float4 *vp = prm.vecs + (size_t) (item.blk_begin + b) * (8 * (size_t) dimp) + lane;
            Acc acc[QT];
#pragma unroll
            for (int qi = 0; qi < QT; qi++) P::init(acc[qi]);
#pragma unroll 2
            for (int c = 0; c < nfull; c++) {
                const float4 x = __ldg(vp + (size_t) c * 32);
#pragma unroll
                for (int qi = 0; qi < QT; qi++) {
                    QE q0, q1, q2, q3;
                    load_q4<QE>(qs + (size_t) qi * dimp + 4 * c, q0, q1, q2, q3);
                    P::step(acc[qi], x.x, q0);
                    P::step(acc[qi], x.y, q1);
                    P::step(acc[qi], x.z, q2);
                    P::step(acc[qi], x.w, q3);
                }
            }
            if (rem) {
                // dim % 4 trailing elements: the pad is never fed to the accumulators (a zero
                // element is not a no-op for the Kahan recurrence)
                const float4 x = __ldg(vp + (size_t) nfull * 32);
#pragma unroll
                for (int qi = 0; qi < QT; qi++) {
                    QE q0, q1, q2, q3;
                    load_q4<QE>(qs + (size_t) qi * dimp + 4 * nfull, q0, q1, q2, q3);
                    P::step(acc[qi], x.x, q0);
                    if (rem > 1) P::step(acc[qi], x.y, q1);
                    if (rem > 2) P::step(acc[qi], x.z, q2);
                }
            }
            NT xn = NT(0);
            if (P::NORMS) xn = valid ? reinterpret_cast<const NT *>(prm.vnorm)[slot] : NT(0);
#pragma unroll
            for (int qi = 0; qi < QT; qi++) {
                const float dist = P::finish(acc[qi], xn, P::NORMS ? s_qnorm[qi] : NT(0));
                top[qi].offer(dist, slot, valid && qi < (int) item.nq, lane, k);
            }
        }

        // ---- merge the NW per-warp lists per query, write k results -----------------------
#pragma unroll
        for (int qi = 0; qi < QT; qi++) {
#pragma unroll
            for (int r = 0; r < KR; r++) {
                const int o = ((warp * QT + qi) * KR + r) * 32 + lane;
                md[o] = top[qi].d[r];
                ms[o] = top[qi].key[r];
            }
        }
        __syncthreads();
        for (int qi = warp; qi < (int) item.nq; qi += SCAN_NW) {
            WarpTopK<KR, uint32_t> fin;
            fin.init();
            for (int w = 0; w < SCAN_NW; w++) {
#pragma unroll
                for (int r = 0; r < KR; r++) {
                    const int o = ((w * QT + qi) * KR + r) * 32 + lane;
                    const float cd = md[o];
                    const uint32_t cs = ms[o];
                    fin.offer(cd, cs, cs != INVALID_SLOT && r * 32 + lane < k, lane, k);
                }
            }
            const size_t base = (size_t) s_pidx[qi] * k;
#pragma unroll
            for (int r = 0; r < KR; r++) {
                const int e = r * 32 + lane;
                if (e < k) {
                    prm.pdist[base + e] = fin.d[r];
                    prm.pslot[base + e] = fin.key[r];
                }
            }
        }
        __syncthreads();   // smem is reused by the next item
    }
}

// ---------------------------------------------------------------------------------------
// merge of per-part top-k lists: warp per query, (dist, id) order.
//   in : pdist/pslot [nq][nparts][k]   (slot -> id through ids[], or id = slot if ids == null)
//   out: dist [nq][k], ids [nq][k] (+inf, -1 when fewer than k)
// ---------------------------------------------------------------------------------------
template <int KR>
__global__ void merge_parts_kernel(const float *pdist, const uint32_t *pslot, const int64_t *ids,
                                   int nq, int nparts, int k, float *out_dist, int64_t *out_ids,
                                   uint32_t *out_slot)
{
    const int lane = threadIdx.x & 31;
    const int q = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (q >= nq) return;
    WarpTopK<KR, int64_t> top;
    top.init();
    const size_t total = (size_t) nparts * k;
    const size_t base = (size_t) q * total;
    for (size_t i = lane; i < round_up64(total, 32); i += 32) {
        float cd = INFINITY;
        int64_t id = -1;
        bool valid = false;
        if (i < total) {
            const uint32_t s = pslot[base + i];
            if (s != INVALID_SLOT) {
                cd = pdist[base + i];
                id = ids ? ids[s] : (int64_t) s;
                valid = true;
            }
        }
        top.offer(cd, id, valid, lane, k);
    }
#pragma unroll
    for (int r = 0; r < KR; r++) {
        const int e = r * 32 + lane;
        if (e < k) {
            const bool have = top.key[r] != KeyMax<int64_t>::v;
            out_dist[(size_t) q * k + e] = have ? top.d[r] : INFINITY;
            if (out_ids) out_ids[(size_t) q * k + e] = have ? top.key[r] : -1;
            if (out_slot) out_slot[(size_t) q * k + e] = have ? (uint32_t) top.key[r] : INVALID_SLOT;
        }
    }
}

// merge of [nshards][nq][k] (dist, id) lists from other ranks (distributed.c:425-438 order)
template <int KR>
__global__ void merge_shards_kernel(const float *dist, const int64_t *ids, int nshards, int nq, int k,
                                    float *out_dist, int64_t *out_ids)
{
    const int lane = threadIdx.x & 31;
    const int q = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (q >= nq) return;
    WarpTopK<KR, int64_t> top;
    top.init();
    for (int s = 0; s < nshards; s++) {
        const size_t base = ((size_t) s * nq + q) * k;
        for (int i = lane; i < round_up(k, 32); i += 32) {
            float cd = INFINITY;
            int64_t id = -1;
            if (i < k) { cd = dist[base + i]; id = ids[base + i]; }
            top.offer(cd, id, i < k && id >= 0, lane, k);
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

}  // namespace ndb
