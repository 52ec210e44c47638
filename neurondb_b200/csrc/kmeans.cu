// kmeans.cu -- IVF k-means training (kmeans_init/run/assign/update_centroids/compute_cost,
// NeuronDB/src/index/ivf_am.c:2070-2294) and the ndb_gpu_backend k-means launchers.
//
// Literal semantics, reproduced bit for bit:
//   assign : argmin_c sum_f32 (x-c)^2, strict <, lowest index wins        -> fused scan, k = 1
//   update : per cluster, f32 running sum over its members IN SAMPLE ORDER, divided by the
//            int count; empty clusters stay at zero                        -> segmented sequential sum
//   cost   : f32 running sum over samples IN ORDER of the squared L2        -> one sequential chain
// Order matters for f32 sums, so the update groups samples by cluster with a stable sort and
// each (cluster, dimension) pair is one thread walking its members in ascending sample index.
#include "kmeans.cuh"
#include "scan.cuh"
#include "cert_common.cuh"

#include <cub/device/device_radix_sort.cuh>
#include <cfloat>
#include <algorithm>

namespace ndb {

// scratch of the row-sharded entry points, kept across Lloyd iterations and released by ndb_b200_shutdown
static KMeansWork g_shard_work;
static DevBuf g_shard_dcost;
void kmeans_at_shutdown()
{
    KMeansWork &w = g_shard_work;
    for (DevBuf *b : {&w.X, &w.C, &w.cstore, &w.assign, &w.keys_sorted, &w.vals_in, &w.vals_sorted, &w.start, &w.counts,
                      &w.cub_tmp, &w.dcost, &w.cost, &w.scr.pdist, &w.scr.pslot, &w.scr.counter, &w.scr.qnorm, &w.cs.store.xb,
                      &w.cs.store.xnorm, &w.cs.store.xrinv, &w.cs.store.xpad0, &w.cs.store.stats, &w.cs.scr.qb, &w.cs.scr.qnorm,
                      &w.cs.scr.qerr, &w.cs.scr.pdist, &w.cs.scr.pslot, &w.cs.scr.items, &w.cs.fb_list, &w.cs.counters, &w.cs.probe,
                      &w.cs.cdist})
        b->release();
    g_shard_dcost.release();
    w.cs.store_ok = false;
    w.cs.items_nq = -1;
}

__global__ void iota_u32_kernel(uint32_t *out, int64_t n)
{
    const int64_t i = (int64_t) blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = (uint32_t) i;
}

__global__ void slot_to_assign_kernel(const uint32_t *slot, int *assign, int64_t n)
{
    const int64_t i = (int64_t) blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) assign[i] = slot[i] == INVALID_SLOT ? 0 : (int) slot[i];   // best starts at 0 (:2277)
}

// segment boundaries of the sorted cluster keys: start[c] = first position with key >= c
__global__ void segment_starts_kernel(const int *sorted_keys, int64_t n, int k, int *start)
{
    const int64_t i = (int64_t) blockIdx.x * blockDim.x + threadIdx.x;
    if (i > n) return;
    const int cur = i < n ? sorted_keys[i] : k;
    const int prev = i > 0 ? sorted_keys[i - 1] : -1;
    for (int c = prev + 1; c <= cur && c <= k; c++) start[c] = (int) i;
}

// one thread per (cluster, dimension): sequential f32 sum over the cluster's members in
// ascending sample order, then sum / (float) count  (ivf_am.c:2189-2212)
__global__ void kmeans_update_kernel(const float *__restrict__ X, const uint32_t *__restrict__ members,
                                     const int *__restrict__ start, int dim, int k, float *__restrict__ C,
                                     int *__restrict__ counts, bool sums_only)
{
    const int c = blockIdx.x;
    const int b = start[c], e = start[c + 1];
    if (threadIdx.x == 0 && blockIdx.y == 0 && counts) counts[c] = e - b;
    const int j = blockIdx.y * blockDim.x + threadIdx.x;
    if (j >= dim) return;
    float sum = 0.0f;
    for (int t = b; t < e; t++) sum = __fadd_rn(sum, X[(size_t) members[t] * dim + j]);
    if (e > b && !sums_only) sum = __fdiv_rn(sum, (float) (e - b));
    C[(size_t) c * dim + j] = sum;
}

// The same per-cluster sums for LARGE row sets (row-sharded training over millions of rows): eight independent
// accumulators per (cluster, dimension) instead of one dependent chain.  The additions happen in a different order
// than kmeans_update_centroids' loop, so the result agrees with it to fp32 rounding, not bit for bit -- as the
// row-sharded result already does through the order of the all-reduce.
__global__ void kmeans_update_fast_kernel(const float *__restrict__ X, const uint32_t *__restrict__ members,
                                          const int *__restrict__ start, int dim, int k, float *__restrict__ C,
                                          int *__restrict__ counts, bool sums_only)
{
    const int c = blockIdx.x;
    const int b = start[c], e = start[c + 1];
    if (threadIdx.x == 0 && blockIdx.y == 0 && counts) counts[c] = e - b;
    const int j = blockIdx.y * blockDim.x + threadIdx.x;
    if (j >= dim) return;
    float acc[8];
#pragma unroll
    for (int u = 0; u < 8; u++) acc[u] = 0.0f;
    int t = b;
    for (; t + 8 <= e; t += 8) {
#pragma unroll
        for (int u = 0; u < 8; u++) acc[u] += X[(size_t) members[t + u] * dim + j];
    }
    for (; t < e; t++) acc[0] += X[(size_t) members[t] * dim + j];
    float sum = ((acc[0] + acc[1]) + (acc[2] + acc[3])) + ((acc[4] + acc[5]) + (acc[6] + acc[7]));
    if (e > b && !sums_only) sum = __fdiv_rn(sum, (float) (e - b));
    C[(size_t) c * dim + j] = sum;
}

// sum of n floats, two levels of fixed trees (deterministic, not the sequential chain)
__global__ void __launch_bounds__(256) tree_sum_partial_kernel(const float *__restrict__ v, int64_t n, float *__restrict__ part)
{
    __shared__ float sh[256];
    float acc = 0.0f;
    const int64_t per = (n + gridDim.x - 1) / gridDim.x, lo = blockIdx.x * per, hi = lo + per < n ? lo + per : n;
    for (int64_t i = lo + threadIdx.x; i < hi; i += 256) acc += v[i];
    sh[threadIdx.x] = acc;
    __syncthreads();
    for (int o = 128; o > 0; o >>= 1) {
        if ((int) threadIdx.x < o) sh[threadIdx.x] += sh[threadIdx.x + o];
        __syncthreads();
    }
    if (threadIdx.x == 0) part[blockIdx.x] = sh[0];
}
__global__ void __launch_bounds__(256) tree_sum_final_kernel(const float *__restrict__ part, int np, float *__restrict__ out)
{
    __shared__ float sh[256];
    float acc = 0.0f;
    for (int i = threadIdx.x; i < np; i += 256) acc += part[i];
    sh[threadIdx.x] = acc;
    __syncthreads();
    for (int o = 128; o > 0; o >>= 1) {
        if ((int) threadIdx.x < o) sh[threadIdx.x] += sh[threadIdx.x + o];
        __syncthreads();
    }
    if (threadIdx.x == 0) *out = sh[0];
}

// row sets above this size take the reordered (rounding-equivalent) sums in the row-sharded entry points
static const int64_t KMEANS_FAST_ROWS = getenv("NDB_KMEANS_EXACT_SUMS") ? ((int64_t) 1 << 40) : ((int64_t) 1 << 20);     // (measurement switch)

// squared L2 of each sample to its centroid, lane-per-sample through a transposed smem tile
__global__ void __launch_bounds__(128) kmeans_sample_cost_kernel(const float *__restrict__ X, const float *__restrict__ C,
                                                                  const int *__restrict__ assign, int64_t n, int dim,
                                                                  float *__restrict__ out)
{
    __shared__ float ta[4][32][33];
    __shared__ float tb[4][32][33];
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const int64_t p0 = ((int64_t) blockIdx.x * 4 + w) * 32;
    if (p0 >= n) return;
    const int np = (int) (n - p0 < 32 ? n - p0 : 32);
    float acc = 0.0f;
    for (int c0 = 0; c0 < dim; c0 += 32) {
        const int cw = dim - c0 < 32 ? dim - c0 : 32;
        for (int r = 0; r < np; r++) {
            if (lane < cw) {
                ta[w][r][lane] = X[(size_t) (p0 + r) * dim + c0 + lane];
                tb[w][r][lane] = C[(size_t) assign[p0 + r] * dim + c0 + lane];
            }
        }
        __syncwarp();
        if (lane < np)
            for (int j = 0; j < cw; j++) {
                const float diff = __fsub_rn(ta[w][lane][j], tb[w][lane][j]);
                acc = __fadd_rn(acc, __fmul_rn(diff, diff));
            }
        __syncwarp();
    }
    if (lane < np) out[p0 + lane] = acc;
}

// cost += d_i in sample order: one dependent f32 chain (ivf_am.c:2218-2233); values are
// staged through shared memory so the chain never waits on DRAM
__global__ void sequential_sum_kernel(const float *__restrict__ v, int64_t n, float *out)
{
    __shared__ float buf[1024];
    float acc = 0.0f;
    for (int64_t base = 0; base < n; base += 1024) {
        const int m = (int) (n - base < 1024 ? n - base : 1024);
        if ((int) threadIdx.x < m) buf[threadIdx.x] = v[base + threadIdx.x];
        __syncthreads();
        if (threadIdx.x == 0)
            for (int i = 0; i < m; i++) acc = __fadd_rn(acc, buf[i]);
        __syncthreads();
    }
    if (threadIdx.x == 0) *out = acc;
}

static int centroids_to_store(KMeansWork &w, int k, int dim, int dimp, cudaStream_t s)
{
    const int64_t blocks = (k + 31) / 32;
    const size_t bytes = (size_t) blocks * 32 * dimp * sizeof(float);
    NDB_CHECK(w.cstore.reserve(bytes));
    NDB_CUDA(cudaMemsetAsync(w.cstore.p, 0, bytes, s));
    return il32_scatter(w.C.as<float>(), k, dim, dimp, nullptr, 0, w.cstore.as<float>(), s);
}

// ---- nearest centroids on the tensor cores, certified -------------------------------------------------------
// The centroid store is scanned in `nranges` tile ranges; each (row, range, column half) leaves its kc best bf16 keys
// and ivf_coarse_cert_kernel re-evaluates the best of their union in the reference's arithmetic, certifying the
// answer or sending the row to the exact kernel (cert_common.cuh).
__global__ void probe_to_assign_kernel(const uint32_t *__restrict__ probe, int *__restrict__ assign, int64_t n)
{
    const int64_t i = (int64_t) blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) assign[i] = probe[i] == INVALID_SLOT ? 0 : (int) probe[i];
}

int nearest_centroids_tensor(CentroidSearch &cs, const float *C_rowmajor, const float *cstore_il32, int L, int dim, int dimp,
                             const float *Q_dev, int nq, int np, bool squared, uint32_t *probe, float *cdist,
                             unsigned long long *counters, cudaStream_t s)
{
    NDB_REQUIRE(np >= 1 && np <= 32 && dim <= TC_MAX_DIM, NDB_B200_EINVAL, "nearest_centroids_tensor: np %d / dim %d out of range", np, dim);
    if (!cs.store_ok) {
        NDB_CHECK(tc_build_store(cs.store, cstore_il32, L, dim, dimp, s));
        cs.store_ok = true;
    }
    const TcStore &st = cs.store;
    const int nkc = st.nkc, kc = TC_KMAX;
    const uint32_t nqt = (uint32_t) ((nq + TC_M - 1) / TC_M);
    const int nqpad = (int) nqt * TC_M;
    NDB_CHECK(cs.scr.qb.reserve((size_t) nqpad * nkc * TC_KC * 2));
    NDB_CHECK(cs.scr.qnorm.reserve((size_t) nqpad * 4));
    NDB_CHECK(tc_block_queries(Q_dev, nullptr, 0, nq, nqpad, dim, nkc, cs.scr.qb.as<__nv_bfloat16>(), cs.scr.qnorm.as<float>(), s));
    // ranges: enough partial lists that their union holds np + a margin of complete entries (np <= 16: the two column
    // halves of one range; else four ranges), and about two items per SM when the store is long enough
    const uint32_t sms = (uint32_t) ctx().sm_count;
    uint32_t nranges = std::max<uint32_t>(np <= 16 ? 1u : 4u, nqt ? (2 * sms) / nqt : 1u);
    nranges = std::max<uint32_t>(1u, std::min<uint32_t>(nranges, (uint32_t) st.ntiles));
    uint32_t tpr = (uint32_t) ((st.ntiles + nranges - 1) / nranges);
    if (tpr > (uint32_t) TC_PACKED_MAX_TILES) tpr = TC_PACKED_MAX_TILES;
    nranges = (uint32_t) ((st.ntiles + tpr - 1) / tpr);
    const uint32_t nitems = nqt * nranges;
    if (cs.items_nq != nq || cs.items_tpr != tpr || cs.items_nranges != nranges || cs.items_ntiles != st.ntiles) {
        std::vector<TcItem> items(nitems);
        for (uint32_t i = 0; i < nitems; i++) {
            const uint32_t qt = i % nqt, xr = i / nqt;
            TcItem &it = items[i];
            it.qtile = qt;
            it.t0 = xr * tpr;
            it.t1 = std::min<uint32_t>((uint32_t) st.ntiles, it.t0 + tpr);
            it.nq = (uint32_t) std::min<int>(TC_M, nq - (int) qt * TC_M);
            it.out_base = qt * TC_M * nranges * 2 + xr * 2;
            it.out_stride = nranges * 2;
            it.rep = 1;
            it.nrows = (uint32_t) (std::min<int64_t>(st.valid_for, (int64_t) it.t1 * TC_N) - (int64_t) it.t0 * TC_N);
        }
        NDB_CHECK(cs.scr.items.reserve((size_t) nitems * sizeof(TcItem)));
        NDB_CUDA(cudaMemcpyAsync(cs.scr.items.p, items.data(), (size_t) nitems * sizeof(TcItem), cudaMemcpyHostToDevice, s));
        NDB_CUDA(cudaStreamSynchronize(s));        // `items` is a host temporary; once per batch shape
        cs.items_nq = nq; cs.items_tpr = tpr; cs.items_nranges = nranges; cs.items_ntiles = st.ntiles;
    }
    const int nparts = (int) nranges * 2;
    NDB_CHECK(cs.scr.pdist.reserve((size_t) nqpad * nparts * kc * 4));
    NDB_CHECK(cs.scr.pslot.reserve((size_t) nqpad * nparts * kc * 4));
    NDB_CHECK(cs.fb_list.reserve((size_t) nq * 4));
    TcParams p;
    memset(&p, 0, sizeof(p));
    p.xb = st.xb.as<__nv_bfloat16>();
    p.xnorm = st.xnorm.as<float>();
    p.qb = cs.scr.qb.as<__nv_bfloat16>();
    p.qnorm = cs.scr.qnorm.as<float>();
    p.nkc = nkc;
    p.dim = dim;
    p.k = kc;
    p.items = cs.scr.items.as<TcItem>();
    p.nitems = nitems;
    p.pdist = cs.scr.pdist.as<float>();
    p.pslot = cs.scr.pslot.as<uint32_t>();
    p.packed = 1;
    NDB_CHECK(tc_launch(p, NDB_L2, kc, s));
    const unsigned grid = (unsigned) ((nq + 3) / 4);
    const size_t fsm = (size_t) dimp * 4 + 256 * 4 + 256 * 8;
    const unsigned fgrid = (unsigned) std::min<int>(nq, 2 * (int) sms);
    using PS = Arith<NDB_L2, NDB_ARITH_IVF_F32>;
    using PQ = Arith<METRIC_L2SQ, NDB_ARITH_IVF_F32>;
#define NDB_CC(KRC, P)                                                                                                      \
    do {                                                                                                                    \
        ivf_coarse_cert_kernel<KRC, P><<<grid, 128, 0, s>>>(p.pdist, p.pslot, nparts, kc, C_rowmajor, Q_dev, nq, L, dim, np,   \
                                                            st.stats.as<float>(), probe, cdist, cs.fb_list.as<uint32_t>(), counters); \
        ivf_coarse_fallback_kernel<P><<<fgrid, 256, fsm, s>>>(cs.fb_list.as<uint32_t>(), counters, C_rowmajor, Q_dev, L, dim, np, probe, cdist); \
    } while (0)
    if (np <= 16) { if (squared) NDB_CC(1, PQ); else NDB_CC(1, PS); }
    else { if (squared) NDB_CC(2, PQ); else NDB_CC(2, PS); }
#undef NDB_CC
    count_launch(2);
    NDB_CUDA(cudaGetLastError());
    return NDB_B200_OK;
}

// assign[i] = nearest centroid of row i (dX row-major); metric selects the squared (k-means,
// :2274-2294) or the sqrtf'd (ivfinsert, :906-935) comparison
int kmeans_assign_dev(KMeansWork &w, const float *dX, int64_t n, int dim, int k, int metric, int *d_assign,
                      cudaStream_t s)
{
    const int dimp = round_up(dim, 4);
    NDB_CHECK(centroids_to_store(w, k, dim, dimp, s));
    // Large centroid sets: GEMM-form distances on the tensor cores, certified per row (cert_common.cuh) -- the
    // assignment is the reference loop's (strict <, lowest index), only found faster.  NDB_ASSIGN_FP32=1 keeps the
    // fp32 scan (measurement switch).
    static const bool force_fp32 = getenv("NDB_ASSIGN_FP32") != nullptr;
    if (k >= 256 && dim <= TC_MAX_DIM && !force_fp32) {
        w.cs.store_ok = false;                                  // the centroids change between calls
        NDB_CHECK(w.cs.counters.reserve(64));
        NDB_CUDA(cudaMemsetAsync(w.cs.counters.p, 0, 64, s));
        const int64_t tchunk = 1 << 20;
        for (int64_t off = 0; off < n; off += tchunk) {
            const int m = (int) (n - off < tchunk ? n - off : tchunk);
            NDB_CHECK(w.cs.probe.reserve((size_t) m * 4));
            NDB_CHECK(w.cs.cdist.reserve((size_t) m * 4));
            NDB_CHECK(nearest_centroids_tensor(w.cs, w.C.as<float>(), w.cstore.as<float>(), k, dim, dimp, dX + (size_t) off * dim, m, 1,
                                               metric == METRIC_L2SQ, w.cs.probe.as<uint32_t>(), w.cs.cdist.as<float>(),
                                               w.cs.counters.as<unsigned long long>(), s));
            probe_to_assign_kernel<<<(unsigned) ((m + 255) / 256), 256, 0, s>>>(w.cs.probe.as<uint32_t>(), d_assign + off, m);
            count_launch();
            NDB_CUDA(cudaGetLastError());
        }
        return NDB_B200_OK;
    }
    const int64_t chunk = 1 << 22;
    for (int64_t off = 0; off < n; off += chunk) {
        const int m = (int) (n - off < chunk ? n - off : chunk);
        int nparts = 1;
        NDB_CHECK(dense_scan(w.cstore.as<float>(), nullptr, k, dim, dimp, metric, NDB_ARITH_IVF_F32,
                             dX + (size_t) off * dim, m, 1, w.scr, &nparts, s));
        if (nparts == 1) {
            slot_to_assign_kernel<<<(unsigned) ((m + 255) / 256), 256, 0, s>>>(w.scr.pslot.as<uint32_t>(), d_assign + off, m);
            count_launch();
        } else {
            // more than one segment of centroids: merge by (dist, index) first
            DevBuf md, msl;
            NDB_CHECK(md.reserve((size_t) m * 4)); NDB_CHECK(msl.reserve((size_t) m * 4));
            NDB_CHECK(launch_merge_parts(w.scr.pdist.as<float>(), w.scr.pslot.as<uint32_t>(), nullptr, m, nparts, 1,
                                         md.as<float>(), nullptr, msl.as<uint32_t>(), s));
            slot_to_assign_kernel<<<(unsigned) ((m + 255) / 256), 256, 0, s>>>(msl.as<uint32_t>(), d_assign + off, m);
            count_launch();
            NDB_CUDA(cudaStreamSynchronize(s));
        }
        NDB_CUDA(cudaGetLastError());
    }
    return NDB_B200_OK;
}

// members of every cluster in ascending row order: w.vals_sorted = the row indices grouped by cluster (stable radix
// sort), w.start[c] .. w.start[c + 1] = the group of cluster c
int group_by_cluster_dev(KMeansWork &w, const int *d_assign, int64_t n, int k, cudaStream_t s)
{
    NDB_REQUIRE(n < (int64_t) 0x7fffffff, NDB_B200_EINVAL, "kmeans_update: n too large");
    NDB_CHECK(w.keys_sorted.reserve((size_t) n * 4));
    NDB_CHECK(w.vals_in.reserve((size_t) n * 4));
    NDB_CHECK(w.vals_sorted.reserve((size_t) n * 4));
    NDB_CHECK(w.start.reserve((size_t) (k + 2) * 4));
    iota_u32_kernel<<<(unsigned) ((n + 255) / 256), 256, 0, s>>>(w.vals_in.as<uint32_t>(), n);
    count_launch();
    int bits = 1;
    while ((1 << bits) < k) bits++;
    size_t tmp_bytes = 0;
    NDB_CUDA(cub::DeviceRadixSort::SortPairs(nullptr, tmp_bytes, d_assign, w.keys_sorted.as<int>(), w.vals_in.as<uint32_t>(),
                                             w.vals_sorted.as<uint32_t>(), (int) n, 0, bits, s));
    NDB_CHECK(w.cub_tmp.reserve(tmp_bytes));
    // LSD radix sort is stable: members of a cluster stay in ascending sample order
    NDB_CUDA(cub::DeviceRadixSort::SortPairs(w.cub_tmp.p, tmp_bytes, d_assign, w.keys_sorted.as<int>(), w.vals_in.as<uint32_t>(),
                                             w.vals_sorted.as<uint32_t>(), (int) n, 0, bits, s));
    count_launch(3);
    segment_starts_kernel<<<(unsigned) ((n + 1 + 255) / 256), 256, 0, s>>>(w.keys_sorted.as<int>(), n, k, w.start.as<int>());
    count_launch();
    NDB_CUDA(cudaGetLastError());
    return NDB_B200_OK;
}

int kmeans_update_dev(KMeansWork &w, const float *dX, const int *d_assign, int64_t n, int dim, int k, float *dC,
                      int *d_counts, cudaStream_t s, bool sums_only)
{
    NDB_CHECK(group_by_cluster_dev(w, d_assign, n, k, s));
    dim3 grid((unsigned) k, (unsigned) ((dim + 127) / 128));
    if (sums_only && n > KMEANS_FAST_ROWS)
        kmeans_update_fast_kernel<<<grid, 128, 0, s>>>(dX, w.vals_sorted.as<uint32_t>(), w.start.as<int>(), dim, k, dC, d_counts, sums_only);
    else
        kmeans_update_kernel<<<grid, 128, 0, s>>>(dX, w.vals_sorted.as<uint32_t>(), w.start.as<int>(), dim, k, dC, d_counts, sums_only);
    count_launch();
    NDB_CUDA(cudaGetLastError());
    return NDB_B200_OK;
}

// kmeans_init + kmeans_run (:2070-2159) on w.X (n*d row-major, already on the device).
// Leaves centroids in w.C, assignments in w.assign, counts in w.counts.
int kmeans_run_dev(KMeansWork &w, int n, int d, int k, int max_iter, float tol, int *iters, float *cost_out,
                   cudaStream_t s)
{
    const size_t cb = (size_t) k * d * 4;
    NDB_CHECK(w.C.reserve(cb)); NDB_CHECK(w.assign.reserve((size_t) n * 4));
    NDB_CHECK(w.counts.reserve((size_t) k * 4)); NDB_CHECK(w.dcost.reserve((size_t) n * 4)); NDB_CHECK(w.cost.reserve(16));
    // kmeans_init: centroid i := sample i for i < n, zero otherwise
    NDB_CUDA(cudaMemsetAsync(w.C.p, 0, cb, s));
    NDB_CUDA(cudaMemcpyAsync(w.C.p, w.X.p, (size_t) (k < n ? k : n) * d * 4, cudaMemcpyDeviceToDevice, s));
    NDB_CUDA(cudaMemsetAsync(w.counts.p, 0, (size_t) k * 4, s));
    float prevCost = FLT_MAX, cost = 0.0f;
    int iter;
    for (iter = 0; iter < max_iter; iter++) {
        NDB_CHECK(kmeans_assign_dev(w, w.X.as<float>(), n, d, k, METRIC_L2SQ, w.assign.as<int>(), s));
        NDB_CHECK(kmeans_update_dev(w, w.X.as<float>(), w.assign.as<int>(), n, d, k, w.C.as<float>(), w.counts.as<int>(), s));
        kmeans_sample_cost_kernel<<<(unsigned) ((n + 127) / 128), 128, 0, s>>>(w.X.as<float>(), w.C.as<float>(),
                                                                              w.assign.as<int>(), n, d, w.dcost.as<float>());
        sequential_sum_kernel<<<1, 1024, 0, s>>>(w.dcost.as<float>(), n, w.cost.as<float>());
        count_launch(2);
        NDB_CUDA(cudaMemcpyAsync(&cost, w.cost.p, 4, cudaMemcpyDeviceToHost, s));
        NDB_CUDA(cudaStreamSynchronize(s));
        // if (fabs(prevCost - cost) < state->threshold) break;   (:2141)
        if (fabs(prevCost - cost) < tol) { iter++; break; }
        prevCost = cost;
    }
    *iters = iter;
    *cost_out = cost;
    return NDB_B200_OK;
}

}  // namespace ndb

using namespace ndb;

extern "C" {

int ndb_b200_kmeans_train(const float *X, int n, int d, int k, int max_iter, float tol, float *C, int *assign,
                          int *counts, int *iters, float *cost_out)
{
    NDB_CHECK(require_init());
    NDB_REQUIRE(X && n > 0 && d > 0 && k > 0 && max_iter >= 0, NDB_B200_EINVAL, "kmeans_train: bad argument");
    NDB_REQUIRE(find_nonfinite(X, (int64_t) n * d) < 0, NDB_B200_EVECTOR, "kmeans_train: NaN/Inf in samples");
    cudaStream_t s = ctx().stream;
    KMeansWork w;
    const size_t xb = (size_t) n * d * 4, cb = (size_t) k * d * 4;
    NDB_CHECK(w.X.reserve(xb));
    NDB_CUDA(cudaMemcpyAsync(w.X.p, X, xb, cudaMemcpyHostToDevice, s));
    int iter = 0;
    float cost = 0.0f;
    NDB_CHECK(kmeans_run_dev(w, n, d, k, max_iter, tol, &iter, &cost, s));
    if (C) NDB_CUDA(cudaMemcpyAsync(C, w.C.p, cb, cudaMemcpyDeviceToHost, s));
    if (assign) NDB_CUDA(cudaMemcpyAsync(assign, w.assign.p, (size_t) n * 4, cudaMemcpyDeviceToHost, s));
    if (counts) NDB_CUDA(cudaMemcpyAsync(counts, w.counts.p, (size_t) k * 4, cudaMemcpyDeviceToHost, s));
    NDB_CUDA(cudaStreamSynchronize(s));
    if (iters) *iters = iter;
    if (cost_out) *cost_out = cost;
    return NDB_B200_OK;
}

int ndb_b200_launch_kmeans_assign(const float *X, const float *C, int *idx, int n, int d, int k, void *stream)
{
    (void) stream;
    NDB_CHECK(require_init());
    NDB_REQUIRE(X && C && idx && n > 0 && d > 0 && k > 0, NDB_B200_EINVAL, "kmeans_assign: bad argument");
    cudaStream_t s = ctx().stream;
    KMeansWork w;
    NDB_CHECK(w.X.reserve((size_t) n * d * 4)); NDB_CHECK(w.C.reserve((size_t) k * d * 4)); NDB_CHECK(w.assign.reserve((size_t) n * 4));
    NDB_CUDA(cudaMemcpyAsync(w.X.p, X, (size_t) n * d * 4, cudaMemcpyHostToDevice, s));
    NDB_CUDA(cudaMemcpyAsync(w.C.p, C, (size_t) k * d * 4, cudaMemcpyHostToDevice, s));
    NDB_CHECK(kmeans_assign_dev(w, w.X.as<float>(), n, d, k, METRIC_L2SQ, w.assign.as<int>(), s));
    NDB_CUDA(cudaMemcpyAsync(idx, w.assign.p, (size_t) n * 4, cudaMemcpyDeviceToHost, s));
    NDB_CUDA(cudaStreamSynchronize(s));
    return NDB_B200_OK;
}

int ndb_b200_launch_kmeans_update(const float *X, const int *idx, float *C, int n, int d, int k, void *stream)
{
    (void) stream;
    NDB_CHECK(require_init());
    NDB_REQUIRE(X && C && idx && n > 0 && d > 0 && k > 0, NDB_B200_EINVAL, "kmeans_update: bad argument");
    for (int i = 0; i < n; i++)
        NDB_REQUIRE(idx[i] >= 0 && idx[i] < k, NDB_B200_EINVAL, "kmeans_update: assignment %d out of range at %d", idx[i], i);
    cudaStream_t s = ctx().stream;
    KMeansWork w;
    NDB_CHECK(w.X.reserve((size_t) n * d * 4)); NDB_CHECK(w.C.reserve((size_t) k * d * 4)); NDB_CHECK(w.assign.reserve((size_t) n * 4));
    NDB_CUDA(cudaMemcpyAsync(w.X.p, X, (size_t) n * d * 4, cudaMemcpyHostToDevice, s));
    NDB_CUDA(cudaMemcpyAsync(w.assign.p, idx, (size_t) n * 4, cudaMemcpyHostToDevice, s));
    NDB_CHECK(kmeans_update_dev(w, w.X.as<float>(), w.assign.as<int>(), n, d, k, w.C.as<float>(), nullptr, s));
    NDB_CUDA(cudaMemcpyAsync(C, w.C.p, (size_t) k * d * 4, cudaMemcpyDeviceToHost, s));
    NDB_CUDA(cudaStreamSynchronize(s));
    return NDB_B200_OK;
}

// ---- row-sharded training: one Lloyd iteration's local half -----------------------
// Every rank holds a slice of the rows and all k centroids.  shard_step assigns the local rows
// (kmeans_assign, :2274-2294) and leaves the per-cluster f32 SUMS (not means) and counts of the
// local members; the caller all-reduces both (NCCL), divides, and calls shard_cost with the new
// centroids for its part of kmeans_compute_cost (:2218-2233).  With one rank the sequence is
// bit-identical to kmeans_train.
int ndb_b200_kmeans_shard_step_dev(const float *X_dev, int64_t n, int d, int k, const float *C_dev, int *assign_dev,
                                   float *sums_dev, int *counts_dev, void *stream)
{
    NDB_CHECK(require_init());
    NDB_REQUIRE(X_dev && C_dev && assign_dev && sums_dev && counts_dev && n > 0 && d > 0 && k > 0, NDB_B200_EINVAL,
                "kmeans_shard_step: bad argument");
    cudaStream_t s = stream ? (cudaStream_t) stream : ctx().stream;
    KMeansWork &w = g_shard_work;
    NDB_CHECK(w.C.reserve((size_t) k * d * 4));
    NDB_CUDA(cudaMemcpyAsync(w.C.p, C_dev, (size_t) k * d * 4, cudaMemcpyDeviceToDevice, s));
    NDB_CHECK(kmeans_assign_dev(w, X_dev, n, d, k, METRIC_L2SQ, assign_dev, s));
    return kmeans_update_dev(w, X_dev, assign_dev, n, d, k, sums_dev, counts_dev, s, true);
}

int ndb_b200_kmeans_shard_cost_dev(const float *X_dev, int64_t n, int d, const float *C_dev, const int *assign_dev,
                                   float *cost_dev, void *stream)
{
    NDB_CHECK(require_init());
    NDB_REQUIRE(X_dev && C_dev && assign_dev && cost_dev && n > 0 && d > 0, NDB_B200_EINVAL, "kmeans_shard_cost: bad argument");
    cudaStream_t s = stream ? (cudaStream_t) stream : ctx().stream;
    DevBuf &dcost = g_shard_dcost;
    NDB_CHECK(dcost.reserve((size_t) n * 4));
    kmeans_sample_cost_kernel<<<(unsigned) ((n + 127) / 128), 128, 0, s>>>(X_dev, C_dev, assign_dev, n, d, dcost.as<float>());
    if (n > KMEANS_FAST_ROWS) {
        // millions of rows: the cost is a convergence test, its single f32 chain (6 M dependent adds) is replaced by trees
        NDB_CHECK(g_shard_work.cost.reserve(1024 * 4));
        tree_sum_partial_kernel<<<1024, 256, 0, s>>>(dcost.as<float>(), n, g_shard_work.cost.as<float>());
        tree_sum_final_kernel<<<1, 256, 0, s>>>(g_shard_work.cost.as<float>(), 1024, cost_dev);
        count_launch();
    } else {
        sequential_sum_kernel<<<1, 1024, 0, s>>>(dcost.as<float>(), n, cost_dev);
    }
    count_launch(2);
    NDB_CUDA(cudaGetLastError());
    return NDB_B200_OK;
}

}  // extern "C"
