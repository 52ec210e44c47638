// scan_launch.cu -- instantiation and launch of the fused scan + top-k kernel (scan.cuh).
#include "layout.cuh"
#include "scan.cuh"

namespace ndb {

static size_t qelem_size(int arith) { return arith == NDB_ARITH_OP_F64 ? 8 : 4; }

static int kr_for(int k) { return k <= 32 ? 1 : (k <= 128 ? 4 : 0); }

static size_t scan_smem_bytes(size_t qelem, int tile, int dimp)
{
    return (size_t) SCAN_STAGES * SCAN_STAGE_BYTES + qelem * (size_t) tile * dimp;
}

// tile shape for (arith, dim, k): QT queries per warp, NW warps per CTA.  Prefer 8 x 8 with two
// CTAs per SM; shrink the CTA, then the per-warp tile, until the query tile fits in shared memory.
static bool scan_shape(int arith, int dim, int k, ScanShape *out)
{
    const int kr = kr_for(k);
    if (kr == 0 || dim <= 0) return false;
    const int dimp = round_up(dim, 4);
    const size_t limit = (ctx().smem_optin ? ctx().smem_optin : 227 * 1024) - 2048;
    const size_t qe = qelem_size(arith);
    // consumer warps per CTA: 4 by default (measured on B200, profiles/r01_scan_variants.txt: CTAs
    // of 4 + 1 warps with three CTAs per SM beat 8 + 1 with two -- smaller groups of warps
    // moving in step through the ring).  NDB_SCAN_NW is a tuning knob for experiments only.
    static const int nw_max = [] { const char *e = getenv("NDB_SCAN_NW"); int v = e ? atoi(e) : 4; return (v == 1 || v == 2 || v == 4 || v == 8) ? v : 4; }();
    for (int pass = 0; pass < 2; pass++) {
        const size_t budget = pass == 0 ? limit / 2 : limit;
        for (int qt = 8; qt >= 1; qt = (qt > 1 ? 1 : 0)) {
            for (int nw = nw_max; nw >= 1; nw >>= 1) {
                if (pass == 0 && nw < 4 && nw_max >= 4) continue;
                const size_t smem = scan_smem_bytes(qe, qt * nw, dimp);
                if (smem <= budget) {
                    out->qt = qt; out->nw = nw; out->kr = kr; out->smem = smem;
                    return true;
                }
            }
            if (qt == 1) break;
        }
    }
    return false;
}

int scan_pick_qt(int arith, int dim, int k)
{
    ScanShape sh;
    return scan_shape(arith, dim, k, &sh) ? sh.tile() : 0;
}

// Streaming variant.  Default: the TMA ring for list mode (short runs re-read from L2 by many
// tiles: staging once per CTA wins) and direct per-warp 128-bit loads for dense mode (long runs,
// warps never wait for each other).  NDB_SCAN_MODE=direct|tma forces one (A/B knob for profiles/).
static bool scan_direct(bool list_mode)
{
    static const int v = [] { const char *e = getenv("NDB_SCAN_MODE"); return !e ? 0 : (strcmp(e, "tma") == 0 ? 2 : (strcmp(e, "direct") == 0 ? 1 : 0)); }();
    return v == 0 ? !list_mode : v == 1;
}

template <class P, int QT, int KR>
static int launch_one_direct(const ScanParams &prm, const ScanShape &sh, uint32_t items_upper, cudaStream_t s)
{
    auto kern = scan_topk_direct_kernel<P, QT, KR>;
    const size_t smem = sh.smem - (size_t) SCAN_STAGES * SCAN_STAGE_BYTES;      // query tile only
    static thread_local size_t configured = 0;
    static thread_local uint64_t cfg_gen = 0;
    if (cfg_gen != ctx().generation) { configured = 0; cfg_gen = ctx().generation; }     // new device / re-init
    if (smem > configured) {
        NDB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) smem));
        configured = smem;
    }
    int per_sm = 0;
    NDB_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, sh.nw * 32, smem));
    if (per_sm < 1) per_sm = 1;
    uint32_t grid = (uint32_t) ctx().sm_count * (uint32_t) per_sm;
    if (items_upper < grid) grid = items_upper;
    if (grid == 0) return NDB_B200_OK;
    NDB_CUDA(cudaMemsetAsync(prm.counter, 0, sizeof(uint32_t), s));
    kern<<<grid, sh.nw * 32, smem, s>>>(prm);
    count_launch();
    NDB_CUDA(cudaGetLastError());
    return NDB_B200_OK;
}

template <class P, int QT, int KR>
static int launch_one(const ScanParams &prm, const ScanShape &sh, uint32_t items_upper, cudaStream_t s)
{
    if (scan_direct(prm.items != nullptr)) return launch_one_direct<P, QT, KR>(prm, sh, items_upper, s);
    auto kern = scan_topk_kernel<P, QT, KR>;
    static thread_local size_t configured = 0;
    static thread_local uint64_t cfg_gen = 0;
    if (cfg_gen != ctx().generation) { configured = 0; cfg_gen = ctx().generation; }
    if (sh.smem > configured) {
        NDB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) sh.smem));
        configured = sh.smem;
    }
    int per_sm = 0;
    NDB_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, (sh.nw + 1) * 32, sh.smem));
    if (per_sm < 1) per_sm = 1;
    uint32_t grid = (uint32_t) ctx().sm_count * (uint32_t) per_sm;
    if (items_upper < grid) grid = items_upper;
    if (grid == 0) return NDB_B200_OK;
    NDB_CUDA(cudaMemsetAsync(prm.counter, 0, sizeof(uint32_t), s));
    kern<<<grid, (sh.nw + 1) * 32, sh.smem, s>>>(prm);      // nw consumer warps + 1 TMA producer warp
    count_launch();
    NDB_CUDA(cudaGetLastError());
    return NDB_B200_OK;
}

template <class P> static int launch_policy(const ScanShape &sh, const ScanParams &prm, uint32_t items_upper, cudaStream_t s)
{
    if (sh.qt == 8 && sh.kr == 1) return launch_one<P, 8, 1>(prm, sh, items_upper, s);
    if (sh.qt == 8 && sh.kr == 4) return launch_one<P, 8, 4>(prm, sh, items_upper, s);
    if (sh.qt == 1 && sh.kr == 1) return launch_one<P, 1, 1>(prm, sh, items_upper, s);
    if (sh.qt == 1 && sh.kr == 4) return launch_one<P, 1, 4>(prm, sh, items_upper, s);
    set_error("scan: unsupported tile/k combination qt=%d k=%d", sh.qt, prm.k);
    return NDB_B200_EINVAL;
}

int launch_scan(int metric, int arith, int tile, const ScanParams &prm, uint32_t items_upper, cudaStream_t s)
{
    ScanShape sh;
    if (!scan_shape(arith, prm.dim, prm.k, &sh) || sh.tile() != tile) {
        set_error("scan: tile %d does not match the shape for dim=%d k=%d", tile, prm.dim, prm.k);
        return NDB_B200_EINVAL;
    }
#define NDB_SCAN_CASE(M, A) \
    if (metric == M && arith == A) return launch_policy<Arith<M, A>>(sh, prm, items_upper, s);
    NDB_SCAN_CASE(NDB_L2, NDB_ARITH_IVF_F32)
    NDB_SCAN_CASE(METRIC_L2SQ, NDB_ARITH_IVF_F32)
    NDB_SCAN_CASE(NDB_COSINE, NDB_ARITH_IVF_F32)
    NDB_SCAN_CASE(NDB_IP, NDB_ARITH_IVF_F32)
    NDB_SCAN_CASE(NDB_L2, NDB_ARITH_OP_F64)
    NDB_SCAN_CASE(NDB_COSINE, NDB_ARITH_OP_F64)
    NDB_SCAN_CASE(NDB_IP, NDB_ARITH_OP_F64)
    NDB_SCAN_CASE(NDB_L2, NDB_ARITH_HNSW)          // f32 difference, f64 sum: also ml_knn.c's euclidean_distance
    NDB_SCAN_CASE(NDB_L2, NDB_ARITH_FAST)
    NDB_SCAN_CASE(METRIC_L2SQ, NDB_ARITH_FAST)
    NDB_SCAN_CASE(NDB_COSINE, NDB_ARITH_FAST)
    NDB_SCAN_CASE(NDB_IP, NDB_ARITH_FAST)
#undef NDB_SCAN_CASE
    set_error("scan: unsupported metric/arith %d/%d", metric, arith);
    return NDB_B200_EINVAL;
}

int launch_merge_parts(const float *pdist, const uint32_t *pslot, const int64_t *ids, int nq, int nparts, int k,
                       float *out_dist, int64_t *out_ids, uint32_t *out_slot, cudaStream_t s)
{
    if (nq <= 0) return NDB_B200_OK;
    const int kr = kr_for(k);
    const unsigned grid = (unsigned) ((nq + 3) / 4);
    if (nparts * k <= 32) {
        if (ids) merge_parts32_kernel<int64_t><<<grid, 128, 0, s>>>(pdist, pslot, ids, nq, nparts * k, k, out_dist, out_ids, out_slot);
        else merge_parts32_kernel<uint32_t><<<grid, 128, 0, s>>>(pdist, pslot, ids, nq, nparts * k, k, out_dist, out_ids, out_slot);
    } else if (kr == 1) merge_parts_kernel<1><<<grid, 128, 0, s>>>(pdist, pslot, ids, nq, nparts, k, out_dist, out_ids, out_slot);
    else if (kr == 4) merge_parts_kernel<4><<<grid, 128, 0, s>>>(pdist, pslot, ids, nq, nparts, k, out_dist, out_ids, out_slot);
    else { set_error("merge: k=%d out of range (1..128)", k); return NDB_B200_EINVAL; }
    count_launch();
    NDB_CUDA(cudaGetLastError());
    return NDB_B200_OK;
}

int launch_merge_shards(const float *dist, const int64_t *ids, int nshards, int nq, int k, float *out_dist,
                        int64_t *out_ids, cudaStream_t s)
{
    if (nq <= 0) return NDB_B200_OK;
    const int kr = kr_for(k);
    const unsigned grid = (unsigned) ((nq + 3) / 4);
    if (kr == 1) merge_shards_kernel<1><<<grid, 128, 0, s>>>(dist, ids, nshards, nq, k, out_dist, out_ids);
    else if (kr == 4) merge_shards_kernel<4><<<grid, 128, 0, s>>>(dist, ids, nshards, nq, k, out_dist, out_ids);
    else { set_error("merge: k=%d out of range (1..128)", k); return NDB_B200_EINVAL; }
    count_launch();
    NDB_CUDA(cudaGetLastError());
    return NDB_B200_OK;
}

}  // namespace ndb

using namespace ndb;

extern "C" {

int ndb_b200_merge_topk_dev(const float *dist_dev, const int64_t *ids_dev, int nshards, int nq, int k,
                            float *out_dist_dev, int64_t *out_ids_dev, void *stream)
{
    NDB_CHECK(require_init());
    NDB_REQUIRE(dist_dev && ids_dev && out_dist_dev && out_ids_dev && nshards > 0 && nq > 0 && k > 0 && k <= 128,
                NDB_B200_EINVAL, "merge_topk: bad argument");
    return launch_merge_shards(dist_dev, ids_dev, nshards, nq, k, out_dist_dev, out_ids_dev,
                               stream ? (cudaStream_t) stream : ctx().stream);
}

int ndb_b200_merge_topk(const float *dist, const int64_t *ids, int nshards, int nq, int k, float *out_dist,
                        int64_t *out_ids)
{
    NDB_CHECK(require_init());
    NDB_REQUIRE(dist && ids && out_dist && out_ids && nshards > 0 && nq > 0 && k > 0 && k <= 128, NDB_B200_EINVAL,
                "merge_topk: bad argument");
    const size_t n = (size_t) nshards * nq * k, m = (size_t) nq * k;
    DevBuf dd, di, od, oi;
    NDB_CHECK(dd.reserve(n * 4)); NDB_CHECK(di.reserve(n * 8)); NDB_CHECK(od.reserve(m * 4)); NDB_CHECK(oi.reserve(m * 8));
    cudaStream_t s = ctx().stream;
    NDB_CUDA(cudaMemcpyAsync(dd.p, dist, n * 4, cudaMemcpyHostToDevice, s));
    NDB_CUDA(cudaMemcpyAsync(di.p, ids, n * 8, cudaMemcpyHostToDevice, s));
    NDB_CHECK(launch_merge_shards(dd.as<float>(), di.as<int64_t>(), nshards, nq, k, od.as<float>(), oi.as<int64_t>(), s));
    NDB_CUDA(cudaMemcpyAsync(out_dist, od.p, m * 4, cudaMemcpyDeviceToHost, s));
    NDB_CUDA(cudaMemcpyAsync(out_ids, oi.p, m * 8, cudaMemcpyDeviceToHost, s));
    NDB_CUDA(cudaStreamSynchronize(s));
    return NDB_B200_OK;
}

}  // extern "C"
