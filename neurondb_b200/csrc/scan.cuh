// scan.cuh -- the streaming distance scan with fused warp-level top-k.
//
// One kernel serves every "many queries x many stored vectors -> k smallest" step of the path:
//   * exact kNN over a dataset (SeqScan + top-N sort)                 dense mode
//   * ivfSelectClusters: queries x centroids, k = nprobe (ivf_am.c:1597-1717)      dense mode
//   * ivfinsert / kmeans_assign: rows x centroids, k = 1 (:906-935, :2164-2177)    dense mode
//   * ivfCollectCandidates: (query, probed list) pairs grouped by list (:1722-1909)  list mode
//
// Work item = (a run of 32-vector blocks, a tile of NW*QT queries).  A persistent CTA takes items
// from an atomic counter.  Lane l owns vector l of a block and walks its dimensions in order with
// QT accumulators -- the query tile sits in shared memory and is read with broadcast LDS.128 -- so
// every distance is produced by exactly the reference's sequential loop (see arith.cuh) while each
// stored vector is fetched from HBM/L2 once per NW*QT queries.  The queries of a tile are dealt to
// the warps round-robin and every warp sees the whole run for its queries, so there is no
// cross-warp merge.  No distance matrix is written: a warp keeps a sorted top-k per query in
// registers (WarpTopK), filters new distances against the k-th best with one ballot, and falls
// back to a shuffle bitonic merge when many lanes pass; it writes k (dist, slot) pairs per
// (query, item-part), and a second small kernel merges the parts per query by (dist, id).
//
// Two ways of bringing the blocks in (profiles/r01_scan_variants.txt):
//   * scan_topk_kernel (list mode): a dedicated producer warp issues TMA bulk copies
//     (cp.async.bulk + mbarrier; an IL32 block is one contiguous run of bytes, so a 16 KB stage is
//     a single copy) into a SCAN_STAGES-deep ring with full/empty barriers; NW consumer warps.
//   * scan_topk_direct_kernel (dense mode): every warp streams the run itself with coalesced
//     LDG.128 (512-byte warp loads); the warps never wait for each other.
#pragma once
#include "arith.cuh"

namespace ndb {

struct WorkItem {
    uint32_t blk_begin;   // first 32-vector block
    uint32_t nvec;        // valid vectors from blk_begin*32 on
    uint32_t qoff;        // first entry of the query tile (index into qmap, or query index)
    uint32_t nq;          // queries in this tile (<= NW*QT)
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
    const uint32_t *qmap;      // entry -> pair index p ; query = p / nprobe
    uint32_t nprobe;           // partial index (list mode) = item index * tile + position in tile
    // dense mode: item i -> seg = i / ntiles, tile = i % ntiles
    uint32_t dense_items, dense_ntiles, dense_seg_blocks, dense_nq, dense_nparts;
    uint64_t dense_nvec;
    uint32_t *counter;         // persistent-CTA work counter (zeroed before launch)
    float *pdist;              // [partial][k]
    uint32_t *pslot;
};

constexpr int SCAN_STAGES = 3;        // TMA ring depth
constexpr int SCAN_CH = 32;           // float4 chunks (of 32 lanes) per stage: 32 * 512 B = 16 KB
constexpr int SCAN_STAGE_BYTES = SCAN_CH * 512;
constexpr int SCAN_MAX_TILE = 64;     // NW * QT upper bound

struct ScanShape {
    int qt, nw, kr;
    size_t smem;
    int tile() const { return qt * nw; }
};

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

// ---- mbarrier / TMA bulk-copy primitives (PTX ISA: mbarrier, cp.async.bulk) -------------------
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t) __cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_fence_init()
{
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t *bar, uint32_t bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t *bar)
{
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity)
{
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "WAIT_%=:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1, %2;\n\t"
        "@p bra DONE_%=;\n\t"
        "bra WAIT_%=;\n\t"
        "DONE_%=:\n\t"
        "}" ::"r"(smem_u32(bar)), "r"(parity), "r"(0x989680u) : "memory");      // suspend-time hint: sleep, do not spin
}
// global -> shared bulk copy completing on an mbarrier (SASS: UBLKCP)
__device__ __forceinline__ void tma_bulk_g2s(void *dst_smem, const void *src_gmem, uint32_t bytes, uint64_t *bar)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                     smem_u32(dst_smem)),
                 "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}
// the chunks of one staged block-part for the NQ live queries of a warp: lane l walks vector l;
// the queries of the warp sit `qstride` elements apart in shared memory
template <class P, int QT, int NQ>
__device__ __forceinline__ void consume_stage(typename P::Acc (&acc)[QT], const float4 *sb, const typename P::Q *myq,
                                              size_t qstride, int c0, int nf, int nch, int rem)
{
    using QE = typename P::Q;
    // One chunk per iteration: NQ * 12 FP instructions is already NQ independent chains, and a
    // body of ~100 instructions stays inside the 6 KB L0 instruction cache (an unroll of 4 does
    // not: ncu showed no_instruction as the top stall).  The next x is fetched one chunk ahead.
    float4 xn = nf > 0 ? sb[0] : make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll 1
    for (int c = 0; c < nf; c++) {
        const float4 x = xn;
        if (c + 1 < nf) xn = sb[(c + 1) * 32];
#pragma unroll
        for (int qi = 0; qi < NQ; qi++) {
            QE q0, q1, q2, q3;
            load_q4<QE>(myq + (size_t) qi * qstride + 4 * (c0 + c), q0, q1, q2, q3);
            P::step(acc[qi], x.x, q0);
            P::step(acc[qi], x.y, q1);
            P::step(acc[qi], x.z, q2);
            P::step(acc[qi], x.w, q3);
        }
    }
    if (rem && nf < nch) {
        // dim % 4 trailing elements: the pad is never fed to the accumulators (a zero element is
        // not a no-op for the Kahan recurrence)
        const float4 x = sb[nf * 32];
#pragma unroll
        for (int qi = 0; qi < NQ; qi++) {
            QE q0, q1, q2, q3;
            load_q4<QE>(myq + (size_t) qi * qstride + 4 * (c0 + nf), q0, q1, q2, q3);
            P::step(acc[qi], x.x, q0);
            if (rem > 1) P::step(acc[qi], x.y, q1);
            if (rem > 2) P::step(acc[qi], x.z, q2);
        }
    }
}

// ---------------------------------------------------------------------------------------
// direct variant: no shared-memory ring.  Every warp streams the item's blocks itself with
// coalesced 128-bit loads (one 512-byte warp-load per 4 dimensions, PF chunks in flight in
// registers); the NW warps of a CTA read the same blocks at about the same time, so all but
// the first reader hit L1.  Warps never wait for each other inside an item.
// ---------------------------------------------------------------------------------------
template <class P, int QT, int NQ>
__device__ __forceinline__ void consume_block_direct(typename P::Acc (&acc)[QT], const float4 *vp, const typename P::Q *myq,
                                                     size_t qstride, int nfull, int rem)
{
    using QE = typename P::Q;
    constexpr int PF = 4;
    const int ntot = nfull + (rem ? 1 : 0);
    float4 xb[PF];
#pragma unroll
    for (int i = 0; i < PF; i++) xb[i] = i < ntot ? __ldg(vp + (size_t) i * 32) : make_float4(0.f, 0.f, 0.f, 0.f);
    int c = 0;
    for (; c + PF <= nfull; c += PF) {
#pragma unroll
        for (int u = 0; u < PF; u++) {
            const float4 x = xb[u];
            if (c + u + PF < ntot) xb[u] = __ldg(vp + (size_t) (c + u + PF) * 32);
#pragma unroll
            for (int qi = 0; qi < NQ; qi++) {
                QE q0, q1, q2, q3;
                load_q4<QE>(myq + (size_t) qi * qstride + 4 * (c + u), q0, q1, q2, q3);
                P::step(acc[qi], x.x, q0);
                P::step(acc[qi], x.y, q1);
                P::step(acc[qi], x.z, q2);
                P::step(acc[qi], x.w, q3);
            }
        }
    }
#pragma unroll
    for (int u = 0; u < PF; u++) {
        const float4 x = xb[u];
        if (c + u < nfull) {
#pragma unroll
            for (int qi = 0; qi < NQ; qi++) {
                QE q0, q1, q2, q3;
                load_q4<QE>(myq + (size_t) qi * qstride + 4 * (c + u), q0, q1, q2, q3);
                P::step(acc[qi], x.x, q0);
                P::step(acc[qi], x.y, q1);
                P::step(acc[qi], x.z, q2);
                P::step(acc[qi], x.w, q3);
            }
        } else if (c + u == nfull && rem) {
            // dim % 4 trailing elements (the pad is never fed to the accumulators)
#pragma unroll
            for (int qi = 0; qi < NQ; qi++) {
                QE q0, q1, q2, q3;
                load_q4<QE>(myq + (size_t) qi * qstride + 4 * (c + u), q0, q1, q2, q3);
                P::step(acc[qi], x.x, q0);
                if (rem > 1) P::step(acc[qi], x.y, q1);
                if (rem > 2) P::step(acc[qi], x.z, q2);
            }
        }
    }
}

template <class P, int QT, int KR>
__global__ void __launch_bounds__(256, 2) scan_topk_direct_kernel(const ScanParams prm)
{
    using QE = typename P::Q;
    using NT = typename P::N;
    using Acc = typename P::Acc;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    QE *qs = reinterpret_cast<QE *>(smem_raw);       // [NW*QT][dimp]
    __shared__ WorkItem s_item;
    __shared__ uint32_t s_item_idx;
    __shared__ uint32_t s_qidx[SCAN_MAX_TILE];
    __shared__ uint32_t s_pidx[SCAN_MAX_TILE];
    __shared__ NT s_qnorm[SCAN_MAX_TILE];

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int nthreads = blockDim.x, NW = nthreads >> 5, TQ = NW * QT;
    const int dimp = prm.dimp, dim = prm.dim, k = prm.k;
    const int nfull = dim >> 2, rem = dim & 3;

    for (;;) {
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
                it.qoff = tile * TQ;
                uint32_t ql = prm.dense_nq - it.qoff;
                it.nq = ql < (uint32_t) TQ ? ql : (uint32_t) TQ;
                it.part = seg;
            } else {
                it.nq = 0;
            }
            s_item = it;
            s_item_idx = idx;
        }
        __syncthreads();
        const WorkItem item = s_item;
        if (item.nq == 0) break;
        const uint32_t nblk = (item.nvec + 31) >> 5;

        if (tid < TQ) {
            uint32_t qi = 0, pi = 0;
            if (tid < (int) item.nq) {
                if (prm.qmap) {
                    uint32_t p = prm.qmap[item.qoff + tid];
                    qi = p / prm.nprobe;
                    pi = s_item_idx * (uint32_t) TQ + (uint32_t) tid;
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
        for (int i = tid; i < TQ * dimp; i += nthreads) {
            int qi = i / dimp, dd = i - qi * dimp;
            float v = 0.0f;
            if (qi < (int) item.nq && dd < dim) v = prm.Q[(size_t) s_qidx[qi] * dim + dd];
            qs[i] = (QE) v;
        }
        __syncthreads();

        // live queries dealt round-robin: warp w owns tile positions w, w+NW, ...
        const int myn = warp < (int) item.nq ? ((int) item.nq - warp + NW - 1) / NW : 0;
        const QE *myq = qs + (size_t) warp * dimp;
        const size_t qstride = (size_t) NW * dimp;
        WarpTopK<KR, uint32_t> top[QT];
#pragma unroll
        for (int qi = 0; qi < QT; qi++) top[qi].init();

        if (myn > 0) {
            for (uint32_t b = 0; b < nblk; b++) {
                const float4 *vp = prm.vecs + (size_t) (item.blk_begin + b) * (8 * (size_t) dimp) + lane;
                Acc acc[QT];
#pragma unroll
                for (int qi = 0; qi < QT; qi++) P::init(acc[qi]);
                if (QT == 1) {
                    consume_block_direct<P, QT, 1>(acc, vp, myq, qstride, nfull, rem);
                } else {
                    switch (myn) {
                    case 1: consume_block_direct<P, QT, 1>(acc, vp, myq, qstride, nfull, rem); break;
                    case 2: consume_block_direct<P, QT, (QT >= 2 ? 2 : 1)>(acc, vp, myq, qstride, nfull, rem); break;
                    case 3: consume_block_direct<P, QT, (QT >= 3 ? 3 : 1)>(acc, vp, myq, qstride, nfull, rem); break;
                    case 4: consume_block_direct<P, QT, (QT >= 4 ? 4 : 1)>(acc, vp, myq, qstride, nfull, rem); break;
                    case 5: consume_block_direct<P, QT, (QT >= 5 ? 5 : 1)>(acc, vp, myq, qstride, nfull, rem); break;
                    case 6: consume_block_direct<P, QT, (QT >= 6 ? 6 : 1)>(acc, vp, myq, qstride, nfull, rem); break;
                    case 7: consume_block_direct<P, QT, (QT >= 7 ? 7 : 1)>(acc, vp, myq, qstride, nfull, rem); break;
                    default: consume_block_direct<P, QT, QT>(acc, vp, myq, qstride, nfull, rem); break;
                    }
                }
                const uint32_t slot = (item.blk_begin + b) * 32 + lane;
                const bool valid = b * 32 + lane < item.nvec;
                NT xn = NT(0);
                if (P::NORMS) xn = valid ? reinterpret_cast<const NT *>(prm.vnorm)[slot] : NT(0);
#pragma unroll
                for (int qi = 0; qi < QT; qi++) {
                    if (qi < myn) {
                        const float dist = P::finish(acc[qi], xn, P::NORMS ? s_qnorm[warp + qi * NW] : NT(0));
                        top[qi].offer(dist, slot, valid, lane, k);
                    }
                }
            }
#pragma unroll
            for (int qi = 0; qi < QT; qi++) {
                if (qi < myn) {
                    const size_t base = (size_t) s_pidx[warp + qi * NW] * k;
#pragma unroll
                    for (int r = 0; r < KR; r++) {
                        const int e = r * 32 + lane;
                        if (e < k) {
                            prm.pdist[base + e] = top[qi].d[r];
                            prm.pslot[base + e] = top[qi].key[r];
                        }
                    }
                }
            }
        }
        __syncthreads();   // the query tile and s_* are rewritten by the next item
    }
}

template <class P, int QT, int KR>
__global__ void __maxnreg__(112) scan_topk_kernel(const ScanParams prm)
{
    using QE = typename P::Q;
    using NT = typename P::N;
    using Acc = typename P::Acc;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    // [stage ring][query tile]
    float4 *ring = reinterpret_cast<float4 *>(smem_raw);
    QE *qs = reinterpret_cast<QE *>(smem_raw + SCAN_STAGES * SCAN_STAGE_BYTES);       // [NW*QT][dimp]
    __shared__ __align__(8) uint64_t full_bar[SCAN_STAGES];    // TMA bytes of the stage have landed
    __shared__ __align__(8) uint64_t empty_bar[SCAN_STAGES];   // all consumer warps are done with the stage
    __shared__ WorkItem s_item;
    __shared__ uint32_t s_item_idx;
    __shared__ uint32_t s_qidx[SCAN_MAX_TILE];     // global query index per tile entry
    __shared__ uint32_t s_pidx[SCAN_MAX_TILE];     // partial-result index per tile entry
    __shared__ NT s_qnorm[SCAN_MAX_TILE];

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    // warps 0..NW-1 consume, warp NW is the TMA producer
    const int nthreads = blockDim.x, NW = (nthreads >> 5) - 1, TQ = NW * QT;
    const int dimp = prm.dimp, dim = prm.dim, k = prm.k;
    const int nfull = dim >> 2, rem = dim & 3, nchunk = dimp >> 2;
    const int spb = (nchunk + SCAN_CH - 1) / SCAN_CH;          // stages per 32-vector block

    if (tid == 0) {
        for (int s = 0; s < SCAN_STAGES; s++) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], NW); }
        mbar_fence_init();
    }
    __syncthreads();
    uint32_t git = 0;       // stages consumed so far by this CTA (ring position and phase)

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
                it.qoff = tile * TQ;
                uint32_t ql = prm.dense_nq - it.qoff;
                it.nq = ql < (uint32_t) TQ ? ql : (uint32_t) TQ;
                it.part = seg;
            } else {
                it.nq = 0;
            }
            s_item = it;
            s_item_idx = idx;
        }
        __syncthreads();
        const WorkItem item = s_item;
        if (item.nq == 0) break;
        const uint32_t nblk = (item.nvec + 31) >> 5;
        const uint32_t total = nblk * spb;

        // ---- stage the query tile ---------------------------------------------------------
        if (tid < TQ) {
            uint32_t qi = 0, pi = 0;
            if (tid < (int) item.nq) {
                if (prm.qmap) {
                    uint32_t p = prm.qmap[item.qoff + tid];
                    qi = p / prm.nprobe;
                    pi = s_item_idx * (uint32_t) TQ + (uint32_t) tid;
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
        for (int i = tid; i < TQ * dimp; i += nthreads) {
            int qi = i / dimp, dd = i - qi * dimp;
            float v = 0.0f;
            if (qi < (int) item.nq && dd < dim) v = prm.Q[(size_t) s_qidx[qi] * dim + dd];
            qs[i] = (QE) v;
        }
        __syncthreads();

        // the live queries of the tile are dealt round-robin to the consumer warps (warp w owns tile
        // positions w, w+NW, ...), so a partial tile keeps every warp busy with fewer queries each
        const int myn = warp < NW && warp < (int) item.nq ? ((int) item.nq - warp + NW - 1) / NW : 0;
        const bool active = myn > 0;
        const QE *myq = qs + (size_t) warp * dimp;
        const size_t qstride = (size_t) NW * dimp;
        WarpTopK<KR, uint32_t> top[QT];
#pragma unroll
        for (int qi = 0; qi < QT; qi++) top[qi].init();
        Acc acc[QT];

        // ---- consume the stages ------------------------------------------------------------
        if (warp == NW) {
            // producer warp: one elected lane refills a ring slot as soon as every consumer warp
            // has released it; stage st = (block st / spb, chunk range st % spb)
            if (lane == 0) {
                for (uint32_t st = 0; st < total; st++) {
                    const uint32_t g = git + st, buf = g % SCAN_STAGES;
                    mbar_wait(&empty_bar[buf], ((g / SCAN_STAGES) & 1u) ^ 1u);
                    const uint32_t b = st / spb, r = st - b * spb;
                    const uint32_t c0 = r * SCAN_CH;
                    const uint32_t nch = min((uint32_t) SCAN_CH, (uint32_t) nchunk - c0);
                    const float4 *src = prm.vecs + (size_t) (item.blk_begin + b) * (8 * (size_t) dimp) + (size_t) c0 * 32;
                    mbar_arrive_expect_tx(&full_bar[buf], nch * 512u);
                    tma_bulk_g2s(ring + (size_t) buf * (SCAN_STAGE_BYTES / 16), src, nch * 512u, &full_bar[buf]);
                }
            }
        } else
        for (uint32_t st = 0; st < total; st++) {
            const uint32_t buf = (git + st) % SCAN_STAGES;
            // warps without a live query still follow the ring so that the phases stay in step
            mbar_wait(&full_bar[buf], ((git + st) / SCAN_STAGES) & 1u);
            if (active) {
                const uint32_t b = st / spb, r = st - b * spb;
                const int c0 = (int) r * SCAN_CH;
                const int nch = min(SCAN_CH, nchunk - c0);
                const int nf = max(0, min(nch, nfull - c0));          // chunks with 4 live elements
                const float4 *sb = ring + (size_t) buf * (SCAN_STAGE_BYTES / 16) + lane;
                if (r == 0) {
#pragma unroll
                    for (int qi = 0; qi < QT; qi++) P::init(acc[qi]);
                }
                if (QT == 1) {
                    consume_stage<P, QT, 1>(acc, sb, myq, qstride, c0, nf, nch, rem);
                } else {
                    switch (myn) {
                    case 1: consume_stage<P, QT, 1>(acc, sb, myq, qstride, c0, nf, nch, rem); break;
                    case 2: consume_stage<P, QT, (QT >= 2 ? 2 : 1)>(acc, sb, myq, qstride, c0, nf, nch, rem); break;
                    case 3: consume_stage<P, QT, (QT >= 3 ? 3 : 1)>(acc, sb, myq, qstride, c0, nf, nch, rem); break;
                    case 4: consume_stage<P, QT, (QT >= 4 ? 4 : 1)>(acc, sb, myq, qstride, c0, nf, nch, rem); break;
                    case 5: consume_stage<P, QT, (QT >= 5 ? 5 : 1)>(acc, sb, myq, qstride, c0, nf, nch, rem); break;
                    case 6: consume_stage<P, QT, (QT >= 6 ? 6 : 1)>(acc, sb, myq, qstride, c0, nf, nch, rem); break;
                    case 7: consume_stage<P, QT, (QT >= 7 ? 7 : 1)>(acc, sb, myq, qstride, c0, nf, nch, rem); break;
                    default: consume_stage<P, QT, QT>(acc, sb, myq, qstride, c0, nf, nch, rem); break;
                    }
                }
                if (r == (uint32_t) spb - 1) {
                    const uint32_t slot = (item.blk_begin + b) * 32 + lane;
                    const bool valid = b * 32 + lane < item.nvec;
                    NT xn = NT(0);
                    if (P::NORMS) xn = valid ? reinterpret_cast<const NT *>(prm.vnorm)[slot] : NT(0);
#pragma unroll
                    for (int qi = 0; qi < QT; qi++) {
                        if (qi < myn) {
                            const float dist = P::finish(acc[qi], xn, P::NORMS ? s_qnorm[warp + qi * NW] : NT(0));
                            top[qi].offer(dist, slot, valid, lane, k);
                        }
                    }
                }
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(&empty_bar[buf]);       // this warp is done with the stage
        }
        git += total;

        // ---- each warp writes the k best of its own queries --------------------------------
        if (active) {
#pragma unroll
            for (int qi = 0; qi < QT; qi++) {
                if (qi < myn) {
                    const size_t base = (size_t) s_pidx[warp + qi * NW] * k;
#pragma unroll
                    for (int r = 0; r < KR; r++) {
                        const int e = r * 32 + lane;
                        if (e < k) {
                            prm.pdist[base + e] = top[qi].d[r];
                            prm.pslot[base + e] = top[qi].key[r];
                        }
                    }
                }
            }
        }
        __syncthreads();   // the query tile and s_* are rewritten by the next item
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

// the same merge when a query's partial lists hold at most 32 entries in all (the tensor path's two
// column halves): one entry per lane and a single 15-stage bitonic sort by (dist, key)
template <class KeyT>
__global__ void merge_parts32_kernel(const float *pdist, const uint32_t *pslot, const int64_t *ids, int nq, int total, int k,
                                     float *out_dist, int64_t *out_ids, uint32_t *out_slot)
{
    const int lane = threadIdx.x & 31;
    const int q = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (q >= nq) return;
    float d = INFINITY;
    KeyT key = KeyMax<KeyT>::v;
    if (lane < total) {
        const uint32_t s = pslot[(size_t) q * total + lane];
        if (s != INVALID_SLOT) {
            d = pdist[(size_t) q * total + lane];
            key = ids ? (KeyT) ids[s] : (KeyT) s;
        }
    }
    warp_sort32<KeyT>(d, key, lane);
    if (lane < k) {
        const bool have = key != KeyMax<KeyT>::v;
        out_dist[(size_t) q * k + lane] = have ? d : INFINITY;
        if (out_ids) out_ids[(size_t) q * k + lane] = have ? (int64_t) key : -1;
        if (out_slot) out_slot[(size_t) q * k + lane] = have ? (uint32_t) key : INVALID_SLOT;
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
