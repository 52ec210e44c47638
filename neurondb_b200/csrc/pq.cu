// pq.cu -- product quantisation (NeuronDB/src/ml/ml_product_quantization.c) on the device: codebook training
// (train_pq_codebook :195-415, train_subspace_kmeans :80-190), encoding (pq_encode_vector :421-536; the backend's
// launch_pq_encode, include/neurondb_gpu_backend.h:106-113, CUDA instance src/gpu/cuda/gpu_pq_kernels.cu:54-110) and the
// asymmetric-distance scan (pq_asymmetric_distance :1003-1110; gpu_pq_kernels.cu:127-158) with a fused top-k.
//
// Arithmetic is the SQL functions' (what a default build of the reference executes): double difference, double square,
// double sum, strict < with the lowest code winning; the asymmetric distance is ONE double chain over all dimensions,
// then (float) sqrt.
//
// The scan is the classic ADC: per query a table T[sub][code] = sum_d (q - c)^2 is built in shared memory (each entry the
// reference's chain over that subspace) and a row costs m byte loads + m table reads -- HBM-bound on the code bytes
// (m bytes per row and query batch, SURVEY 8d's per-unit figure for a quantised scan).  Adding m table entries is not
// the reference's association (it runs one chain across the subspace boundaries), so the sum differs from the
// reference's in the last bits of the DOUBLE; the float it is cast to is the same unless the double lies within the
// rounding error of a float rounding boundary.  Every distance is therefore certified -- (float)(d(1-e)) == (float)(d(1+e))
// with e above both sums' error bound -- and the few that are not (about dim * 2^-27 of them) are re-evaluated with the
// reference's chain.  Result: float distances bit-identical to pq_asymmetric_distance, top-k by (distance, row).
#include <cfloat>
#include <cstdlib>

#include "kmeans.cuh"
#include "scan.cuh"
#include "block_topk.cuh"

namespace ndb {

struct PqIndex {
    int dim = 0, m = 0, ksub = 0, dsub = 0;
    int64_t n = 0;
    DevBuf codebooks;        // float [m][ksub][dsub]
    DevBuf codes;            // uint8 [n][m]   (ksub <= 256)
    DevBuf qbuf, pdist, pslot, outd, outi, all;
};

// codes of n rows: nearest codeword per subspace.  dXT = rows transposed [dim][n].
static int pq_encode_dev(const float *dXT, int64_t n, int dim, const float *dcb, int m, int ksub, int *dassign /* n */, int *dchanged,
                         uint8_t *codes8 /* [n][m] or null */, int16_t *codes16 /* [n][m] or null */, cudaStream_t s);

__global__ void pq_store_codes_kernel(const int *__restrict__ assign, int64_t n, int m, int sub, uint8_t *__restrict__ c8, int16_t *__restrict__ c16)
{
    const int64_t i = (int64_t) blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    if (c8) c8[(size_t) i * m + sub] = (uint8_t) assign[i];
    if (c16) c16[(size_t) i * m + sub] = (int16_t) assign[i];
}

static int pq_encode_dev(const float *dXT, int64_t n, int dim, const float *dcb, int m, int ksub, int *dassign, int *dchanged,
                         uint8_t *codes8, int16_t *codes16, cudaStream_t s)
{
    const int dsub = dim / m;
    DevBuf wide;
    for (int sub = 0; sub < m; sub++) {
        NDB_CHECK(nearest_f64_dev(dXT + (size_t) sub * dsub * n, dcb + (size_t) sub * ksub * dsub, n, dsub, ksub, dassign, dchanged, wide, s));
        pq_store_codes_kernel<<<(unsigned) ((n + 255) / 256), 256, 0, s>>>(dassign, n, m, sub, codes8, codes16);
        count_launch();
    }
    NDB_CUDA(cudaGetLastError());
    return NDB_B200_OK;
}

__global__ void pq_widen_codes_kernel(const int16_t *__restrict__ in, int64_t total, int ksub, uint8_t *__restrict__ out, int *__restrict__ bad)
{
    const int64_t i = (int64_t) blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= total) return;
    const int c = in[i];
    if (c < 0 || c >= ksub) atomicMin(bad, (int) (i < 0x7fffffff ? i : 0x7ffffffe));
    out[i] = (uint8_t) c;
}

// the reference's chain for one (query, row): pq_asymmetric_distance :1063-1098
__device__ __forceinline__ float pq_exact_distance(const float *__restrict__ q, const uint8_t *__restrict__ code, const float *__restrict__ cb,
                                                   int m, int ksub, int dsub)
{
    double total = 0.0;
    for (int sub = 0; sub < m; sub++) {
        const float *cw = cb + ((size_t) sub * ksub + code[sub]) * dsub;
        for (int d = 0; d < dsub; d++) {
            const double diff = __dsub_rn((double) q[sub * dsub + d], (double) cw[d]);
            total = __dadd_rn(total, __dmul_rn(diff, diff));
        }
    }
    return __double2float_rn(__dsqrt_rn(total));
}

// sum over the subspaces of table[sub][code[sub]], in subspace order
__device__ __forceinline__ double pq_add(double a, double b) { return __dadd_rn(a, b); }
__device__ __forceinline__ float pq_add(float a, float b) { return __fadd_rn(a, b); }
// (KSUB = 256: the usual 8-bit codebook, table rows at a constant stride -- the index arithmetic folds into the loads)
template <class T, int KSUB>
__device__ __forceinline__ T pq_table_sum(const T *__restrict__ table, const uint8_t *__restrict__ code, int m, int ksub_rt)
{
    const int ksub = KSUB > 0 ? KSUB : ksub_rt;
    T total = (T) 0;
    if ((m & 3) == 0) {
        const uint32_t *c4 = reinterpret_cast<const uint32_t *>(code);
        for (int sub = 0; sub < m; sub += 4) {
            const uint32_t w = c4[sub >> 2];
            total = pq_add(total, table[(sub + 0) * ksub + (w & 0xff)]);
            total = pq_add(total, table[(sub + 1) * ksub + ((w >> 8) & 0xff)]);
            total = pq_add(total, table[(sub + 2) * ksub + ((w >> 16) & 0xff)]);
            total = pq_add(total, table[(sub + 3) * ksub + (w >> 24)]);
        }
    } else {
        for (int sub = 0; sub < m; sub++) total = pq_add(total, table[sub * ksub + code[sub]]);
    }
    return total;
}

// grid (nparts, nq); every block builds its query's table, its warps walk the block's row range 32 rows at a time.
// KR > 0: partial top-k lists pdist / pslot [nq][nparts][k]; KR == 0: all distances to out_all [nq][n].
// FILTER (top-k only): a float copy of the table screens the rows.  Its sum t32 bounds the true sum from below
// (t32 (1 - delta), delta = (m + 8) 2^-23 covering the entries' and the additions' roundings), and the smallest k-th
// distance any warp of the block holds bounds what can still enter the block's result, so a row with
// t32 (1 - delta) > bound^2 (1 + 2^-22) is dropped after m 4-byte reads; everything else takes the exact path below.  The
// 4-byte reads cost about half the bank conflicts of the 8-byte ones, and the fp64 work disappears from the common case.
template <int KR, bool FILTER, int KSUB>
__global__ void __launch_bounds__(256) pq_adc_kernel(const float *__restrict__ Q, const uint8_t *__restrict__ codes, const float *__restrict__ cb,
                                                     int64_t n, int dim, int m, int ksub, int k, int64_t rows_per_part, double eps,
                                                     float *__restrict__ pdist, uint32_t *__restrict__ pslot, float *__restrict__ out_all,
                                                     unsigned long long *__restrict__ recheck_count)
{
    extern __shared__ double lut[];                    // [m][ksub] doubles, then (FILTER) [m][ksub] floats
    __shared__ unsigned int s_bound;                   // float bits of the block's smallest k-th distance
    float *lut32 = reinterpret_cast<float *>(lut + (size_t) m * ksub);
    const int dsub = dim / m;
    const int qi = blockIdx.y, part = blockIdx.x;
    const float *q = Q + (size_t) qi * dim;
    for (int e = threadIdx.x; e < m * ksub; e += blockDim.x) {
        const int sub = e / ksub;
        const float *cw = cb + (size_t) e * dsub;
        double t = 0.0;
        for (int d = 0; d < dsub; d++) {
            const double diff = __dsub_rn((double) q[sub * dsub + d], (double) cw[d]);
            t = __dadd_rn(t, __dmul_rn(diff, diff));
        }
        lut[e] = t;
        if (FILTER) lut32[e] = __double2float_rn(t);
    }
    if (threadIdx.x == 0) s_bound = 0x7f800000u;       // +inf
    __syncthreads();
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
    const int64_t r0 = (int64_t) part * rows_per_part;
    const int64_t r1 = r0 + rows_per_part < n ? r0 + rows_per_part : n;
    const float one_minus_delta = 1.0f - (float) (m + 8) * 1.1920928955078125e-07f;       // exact: a multiple of 2^-23 below 1
    unsigned int bound_seen = 0x7f800000u;
    float limit = INFINITY;                            // bound^2 (1 + 2^-22), rounded up
    WarpTopK<(KR > 0 ? KR : 1), uint32_t> top;
    if (KR > 0) top.init();
    for (int64_t base = r0 + (int64_t) warp * 32; base < r1; base += (int64_t) nwarps * 32) {
        const int64_t row = base + lane;
        const bool valid = row < r1;
        const uint8_t *code = codes + (size_t) (valid ? row : r0) * m;
        bool pass = valid;
        if (FILTER) {
            const unsigned int b = *(volatile unsigned int *) &s_bound;
            if (b != bound_seen) {
                bound_seen = b;
                const double bd = (double) __uint_as_float(b);
                limit = __double2float_ru(bd * bd * (1.0 + 2.384185791015625e-07));
            }
            const float t32 = pq_table_sum<float, KSUB>(lut32, code, m, ksub);
            pass = valid && (__fmul_rd(t32, one_minus_delta) <= limit || t32 < 1e-30f);      // (below: float subnormals carry no bound)
        }
        float f = INFINITY;
        if (pass) {
            const double total = pq_table_sum<double, KSUB>(lut, code, m, ksub);
            const double d = __dsqrt_rn(total);
            const float lo = __double2float_rn(d * (1.0 - eps)), hi = __double2float_rn(d * (1.0 + eps));
            if (lo == hi) f = lo;
            else {
                f = pq_exact_distance(q, code, cb, m, ksub, dsub);
                if (recheck_count) atomicAdd(recheck_count, 1ull);
            }
            if (KR == 0) out_all[(size_t) qi * n + row] = f;
        }
        if (KR > 0) {
            const float td_before = top.td;
            top.offer(f, (uint32_t) row, pass, lane, k);
            if (FILTER && lane == 0 && top.td < td_before) atomicMin(&s_bound, __float_as_uint(top.td));
        }
    }
    if (KR > 0) {
        __syncthreads();                               // the tables are no longer needed: their memory carries the merge
        block_topk_write<(KR > 0 ? KR : 1)>(top, lut, warp, lane, nwarps, k, pdist, pslot, ((size_t) qi * gridDim.x + part) * k);
    }
}

static int pq_scan(PqIndex *pq, const float *Q_dev, int nq, int k, float *dist_dev, int64_t *rows_dev, float *all_dev,
                   unsigned long long *recheck_dev, cudaStream_t s)
{
    NDB_REQUIRE(nq <= 65535, NDB_B200_EINVAL, "pq: at most 65535 queries per call (got %d); split the batch", nq);     // grid.y
    const size_t lut_bytes = (size_t) pq->m * pq->ksub * sizeof(double);
    const int kr = k <= 0 ? 0 : (k <= 32 ? 1 : (k <= 128 ? 4 : -1));
    size_t smem = lut_bytes;
    // the float screen needs half as much again; without it (table too large, or NDB_PQ_NO_FILTER for comparison) every row
    // takes the exact path
    const bool filter = kr > 0 && lut_bytes + lut_bytes / 2 <= ctx().smem_optin && !getenv("NDB_PQ_NO_FILTER");
    if (filter) smem += lut_bytes / 2;
    if (kr > 0 && (size_t) 8 * kr * 32 * 8 > smem) smem = (size_t) 8 * kr * 32 * 8;
    NDB_REQUIRE(smem <= ctx().smem_optin, NDB_B200_EINVAL, "pq: m * ksub = %d table entries do not fit shared memory", pq->m * pq->ksub);
    // rows per block: enough to amortise the table (ksub * dim * 3 fp64 operations) over the rows (m reads each), and about
    // two waves of blocks over the SMs
    int64_t rows_per_part = 65536;
    const int64_t want_blocks = (int64_t) 2 * ctx().sm_count;
    while (rows_per_part > 4096 && (pq->n + rows_per_part - 1) / rows_per_part * nq < want_blocks) rows_per_part >>= 1;
    const int nparts = (int) ((pq->n + rows_per_part - 1) / rows_per_part);
    // error bound of either summation: (dim + m) additions and dim products of non-negative terms, each within 2^-53
    // relative; sqrt halves it and adds one rounding.  eps is well above that.
    double eps = (double) (pq->dim + pq->m + 8) * 4.0 * 1.1102230246251565e-16;
    if (const char *e = getenv("NDB_PQ_EPS")) eps = atof(e) > eps ? atof(e) : eps;      // (test switch: widen the band to exercise the re-evaluation)
    dim3 grid((unsigned) nparts, (unsigned) nq);
#define NDB_PQ_LAUNCH2(KR, F, KS)                                                                                                      \
    do {                                                                                                                               \
        NDB_CUDA(cudaFuncSetAttribute(pq_adc_kernel<KR, F, KS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) smem));             \
        pq_adc_kernel<KR, F, KS><<<grid, 256, smem, s>>>(Q_dev, pq->codes.as<uint8_t>(), pq->codebooks.as<float>(), pq->n, pq->dim,     \
                                                         pq->m, pq->ksub, k, rows_per_part, eps, pq->pdist.as<float>(),                \
                                                         pq->pslot.as<uint32_t>(), all_dev, recheck_dev);                              \
    } while (0)
#define NDB_PQ_LAUNCH1(KR, F)                                                                                                          \
    do {                                                                                                                               \
        if (pq->ksub == 256) NDB_PQ_LAUNCH2(KR, F, 256);                                                                               \
        else NDB_PQ_LAUNCH2(KR, F, 0);                                                                                                 \
    } while (0)
#define NDB_PQ_LAUNCH(KR)                                                                                                              \
    do {                                                                                                                               \
        if (filter) NDB_PQ_LAUNCH1(KR, true);                                                                                          \
        else NDB_PQ_LAUNCH1(KR, false);                                                                                                \
    } while (0)
    Context &c = ctx();
    if (c.timing) NDB_CUDA(cudaEventRecord(c.ev0, s));
    if (kr == 0) NDB_PQ_LAUNCH1(0, false);
    else {
        NDB_CHECK(pq->pdist.reserve((size_t) nq * nparts * k * sizeof(float)));
        NDB_CHECK(pq->pslot.reserve((size_t) nq * nparts * k * sizeof(uint32_t)));
        if (kr == 1) NDB_PQ_LAUNCH(1);
        else if (kr == 4) NDB_PQ_LAUNCH(4);
        else { set_error("pq_search: k=%d out of range (1..128)", k); return NDB_B200_EINVAL; }
    }
#undef NDB_PQ_LAUNCH
#undef NDB_PQ_LAUNCH1
#undef NDB_PQ_LAUNCH2
    if (c.timing) {                                    // ndb_b200_last_kernel_stats: the scan kernel alone
        NDB_CUDA(cudaEventRecord(c.ev1, s));
        c.last_bytes = (double) pq->n * pq->m * nq;    // m code bytes per (row, query)
        c.last_evals = pq->n * (int64_t) nq;
        c.stats_src = nullptr;
        c.last_ms = -1.0;
    }
    count_launch();
    NDB_CUDA(cudaGetLastError());
    if (kr > 0)
        NDB_CHECK(launch_merge_parts(pq->pdist.as<float>(), pq->pslot.as<uint32_t>(), nullptr, nq, nparts, k, dist_dev, rows_dev, nullptr, s));
    return NDB_B200_OK;
}

}  // namespace ndb

using namespace ndb;

struct ndb_b200_pq : PqIndex {};

static int pq_check_shape(int dim, int m, int ksub, const char *who)
{
    NDB_REQUIRE(m >= 1 && m <= 128, NDB_B200_EINVAL, "m (number of subspaces) must be between 1 and 128");                       // :218-222
    NDB_REQUIRE(ksub >= 2 && ksub <= 65536, NDB_B200_EINVAL, "ksub (centroids per subspace) must be between 2 and 65536");      // :223-227
    NDB_REQUIRE(dim > 0, NDB_B200_EINVAL, "Invalid vector dimension: %d", dim);
    NDB_REQUIRE(dim % m == 0, NDB_B200_EDIM, "Vector dimension %d must be divisible by number of subspaces m=%d", dim, m);       // :270-276
    (void) who;
    return NDB_B200_OK;
}

extern "C" {

int ndb_b200_pq_train(const float *X, int n, int dim, int m, int ksub, int max_iters, const int *rand_draws, float *codebooks)
{
    NDB_CHECK(require_init());
    NDB_REQUIRE(X && rand_draws && codebooks, NDB_B200_EINVAL, "pq_train: NULL argument");
    NDB_CHECK(pq_check_shape(dim, m, ksub, "pq_train"));
    NDB_REQUIRE(n > 0, NDB_B200_EINVAL, "No training vectors found");                                                          // :243-249
    if (max_iters < 0) max_iters = 100;                                                                                        // the reference passes 100 (:340-346)
    const int dsub = dim / m;
    cudaStream_t s = ctx().stream;
    KMeansWork w;
    DevBuf X_all, XT, dseeds, dchanged, dcb;
    const size_t xb = (size_t) n * dim * 4, sb = (size_t) n * dsub * 4;
    NDB_CHECK(upload_rows_checked(X_all, X, (size_t) n * dim, "pq_train", s));
    NDB_CHECK(XT.reserve(xb)); NDB_CHECK(w.X.reserve(sb)); NDB_CHECK(w.C.reserve((size_t) ksub * dsub * 4));
    NDB_CHECK(w.assign.reserve((size_t) n * 4)); NDB_CHECK(w.counts.reserve((size_t) ksub * 4));
    NDB_CHECK(dseeds.reserve((size_t) ksub * 4)); NDB_CHECK(dchanged.reserve(4)); NDB_CHECK(dcb.reserve((size_t) m * ksub * dsub * 4));
    NDB_CHECK(transpose_rows_dev(X_all.as<float>(), n, dim, XT.as<float>(), s));
    std::vector<int> seeds(ksub);
    for (int sub = 0; sub < m; sub++) {
        // the subspace's rows, contiguous (:329-338), and its seeds: rows rand() % nvec, duplicates allowed (:98-103)
        NDB_CUDA(cudaMemcpy2DAsync(w.X.p, (size_t) dsub * 4, X_all.as<float>() + (size_t) sub * dsub, (size_t) dim * 4, (size_t) dsub * 4,
                                   (size_t) n, cudaMemcpyDeviceToDevice, s));
        for (int c = 0; c < ksub; c++) {
            NDB_REQUIRE(rand_draws[(size_t) sub * ksub + c] >= 0, NDB_B200_EINVAL, "pq_train: rand() values are non-negative");
            seeds[c] = rand_draws[(size_t) sub * ksub + c] % n;
        }
        NDB_CUDA(cudaMemcpyAsync(dseeds.p, seeds.data(), (size_t) ksub * 4, cudaMemcpyHostToDevice, s));
        NDB_CUDA(cudaStreamSynchronize(s));                                    // (seeds is reused by the next subspace)
        NDB_CHECK(gather_rows_dev(w.X.as<float>(), dseeds.as<int>(), ksub, dsub, w.C.as<float>(), s));
        NDB_CUDA(cudaMemsetAsync(w.assign.p, 0, (size_t) n * 4, s));           // palloc0 (:96)
        int iters = 0;
        NDB_CHECK(lloyd_f64_dev(w, w.X.as<float>(), XT.as<float>() + (size_t) sub * dsub * n, n, dsub, ksub, max_iters, true, dchanged.as<int>(),
                                &iters, s));
        NDB_CUDA(cudaMemcpyAsync(dcb.as<float>() + (size_t) sub * ksub * dsub, w.C.p, (size_t) ksub * dsub * 4, cudaMemcpyDeviceToDevice, s));
    }
    NDB_CUDA(cudaMemcpyAsync(codebooks, dcb.p, (size_t) m * ksub * dsub * 4, cudaMemcpyDeviceToHost, s));
    NDB_CUDA(cudaStreamSynchronize(s));
    return NDB_B200_OK;
}

static int pq_encode_host(const float *X, int64_t n, int dim, const float *codebooks, int m, int ksub, uint8_t *codes8, int16_t *codes16)
{
    NDB_CHECK(require_init());
    NDB_REQUIRE(X && codebooks && (codes8 || codes16) && n > 0, NDB_B200_EINVAL, "pq_encode: NULL or empty argument");
    NDB_CHECK(pq_check_shape(dim, m, ksub, "pq_encode"));
    NDB_REQUIRE(!codes8 || ksub <= 256, NDB_B200_EINVAL, "pq_encode: byte codes need ksub <= 256");
    NDB_REQUIRE(!codes16 || ksub <= 32768, NDB_B200_EINVAL, "pq_encode: int2 codes need ksub <= 32768");
    cudaStream_t s = ctx().stream;
    DevBuf dX, dXT, dcb, dassign, dchanged, dcodes;
    const size_t xb = (size_t) n * dim * 4, cbb = (size_t) m * ksub * (dim / m) * 4, ob = (size_t) n * m * (codes8 ? 1 : 2);
    NDB_CHECK(upload_rows_checked(dX, X, (size_t) n * dim, "pq_encode", s));
    NDB_CHECK(dXT.reserve(xb)); NDB_CHECK(dcb.reserve(cbb)); NDB_CHECK(dassign.reserve((size_t) n * 4));
    NDB_CHECK(dchanged.reserve(4)); NDB_CHECK(dcodes.reserve(ob));
    NDB_CUDA(cudaMemcpyAsync(dcb.p, codebooks, cbb, cudaMemcpyHostToDevice, s));
    NDB_CHECK(transpose_rows_dev(dX.as<float>(), n, dim, dXT.as<float>(), s));
    NDB_CUDA(cudaMemsetAsync(dassign.p, 0xff, (size_t) n * 4, s));
    NDB_CHECK(pq_encode_dev(dXT.as<float>(), n, dim, dcb.as<float>(), m, ksub, dassign.as<int>(), dchanged.as<int>(),
                            codes8 ? dcodes.as<uint8_t>() : nullptr, codes16 ? dcodes.as<int16_t>() : nullptr, s));
    NDB_CUDA(cudaMemcpyAsync(codes8 ? (void *) codes8 : (void *) codes16, dcodes.p, ob, cudaMemcpyDeviceToHost, s));
    NDB_CUDA(cudaStreamSynchronize(s));
    return NDB_B200_OK;
}

int ndb_b200_pq_encode(const float *X, int64_t n, int dim, const float *codebooks, int m, int ksub, int16_t *codes)
{
    return pq_encode_host(X, n, dim, codebooks, m, ksub, nullptr, codes);
}

int ndb_b200_launch_pq_encode(const float *X, const float *codebooks, uint8_t *codes, int n, int d, int m, int ks, void *stream)
{
    (void) stream;
    return pq_encode_host(X, n, d, codebooks, m, ks, codes, nullptr);
}

int ndb_b200_pq_create(int dim, int m, int ksub, const float *codebooks, ndb_b200_pq **out)
{
    NDB_CHECK(require_init());
    NDB_REQUIRE(codebooks && out, NDB_B200_EINVAL, "pq_create: NULL argument");
    NDB_CHECK(pq_check_shape(dim, m, ksub, "pq_create"));
    NDB_REQUIRE(ksub <= 256, NDB_B200_EINVAL, "pq_create: the resident scan keeps byte codes (ksub <= 256), got ksub=%d", ksub);
    NDB_REQUIRE((size_t) m * ksub * 8 <= ctx().smem_optin, NDB_B200_EINVAL, "pq_create: m * ksub = %d table entries do not fit shared memory", m * ksub);
    const size_t cbb = (size_t) m * ksub * (dim / m) * 4;
    NDB_REQUIRE(find_nonfinite(codebooks, (int64_t) (cbb / 4)) < 0, NDB_B200_EVECTOR, "pq_create: NaN/Inf in the codebook");
    ndb_b200_pq *pq = new ndb_b200_pq();
    pq->dim = dim; pq->m = m; pq->ksub = ksub; pq->dsub = dim / m;
    int rc = pq->codebooks.reserve(cbb);
    if (rc == NDB_B200_OK && cudaMemcpy(pq->codebooks.p, codebooks, cbb, cudaMemcpyHostToDevice) != cudaSuccess) {
        set_error("pq_create: codebook upload failed");
        rc = NDB_B200_ECUDA;
    }
    if (rc != NDB_B200_OK) { delete pq; return rc; }
    *out = pq;
    return NDB_B200_OK;
}

void ndb_b200_pq_free(ndb_b200_pq *pq) { delete pq; }

int64_t ndb_b200_pq_size(const ndb_b200_pq *pq) { return pq ? pq->n : 0; }

int ndb_b200_pq_add(ndb_b200_pq *pq, const float *X, int64_t n, int16_t *codes_out)
{
    NDB_CHECK(require_init());
    NDB_REQUIRE(pq && X && n > 0, NDB_B200_EINVAL, "pq_add: NULL or empty argument");
    NDB_REQUIRE(pq->n + n < (int64_t) 0xfffffff0ll, NDB_B200_EINVAL, "pq_add: too many rows for 32-bit slots");
    cudaStream_t s = ctx().stream;
    DevBuf dX, dXT, dassign, dchanged, d16;
    const size_t xb = (size_t) n * pq->dim * 4;
    NDB_CHECK(upload_rows_checked(dX, X, (size_t) n * pq->dim, "pq_add", s));
    NDB_CHECK(dXT.reserve(xb)); NDB_CHECK(dassign.reserve((size_t) n * 4)); NDB_CHECK(dchanged.reserve(4));
    if (codes_out) NDB_CHECK(d16.reserve((size_t) n * pq->m * 2));
    NDB_CHECK(pq->codes.grow((size_t) (pq->n + n) * pq->m, (size_t) pq->n * pq->m, s));
    NDB_CHECK(transpose_rows_dev(dX.as<float>(), n, pq->dim, dXT.as<float>(), s));
    NDB_CUDA(cudaMemsetAsync(dassign.p, 0xff, (size_t) n * 4, s));
    NDB_CHECK(pq_encode_dev(dXT.as<float>(), n, pq->dim, pq->codebooks.as<float>(), pq->m, pq->ksub, dassign.as<int>(), dchanged.as<int>(),
                            pq->codes.as<uint8_t>() + (size_t) pq->n * pq->m, codes_out ? d16.as<int16_t>() : nullptr, s));
    if (codes_out) NDB_CUDA(cudaMemcpyAsync(codes_out, d16.p, (size_t) n * pq->m * 2, cudaMemcpyDeviceToHost, s));
    NDB_CUDA(cudaStreamSynchronize(s));
    pq->n += n;
    return NDB_B200_OK;
}

int ndb_b200_pq_add_codes(ndb_b200_pq *pq, const int16_t *codes, int64_t n)
{
    NDB_CHECK(require_init());
    NDB_REQUIRE(pq && codes && n > 0, NDB_B200_EINVAL, "pq_add_codes: NULL or empty argument");
    NDB_REQUIRE(pq->n + n < (int64_t) 0xfffffff0ll, NDB_B200_EINVAL, "pq_add_codes: too many rows for 32-bit slots");
    cudaStream_t s = ctx().stream;
    DevBuf d16, dbad;
    const int64_t total = n * pq->m;
    NDB_CHECK(d16.reserve((size_t) total * 2)); NDB_CHECK(dbad.reserve(4));
    NDB_CHECK(pq->codes.grow((size_t) (pq->n + n) * pq->m, (size_t) pq->n * pq->m, s));
    NDB_CUDA(cudaMemcpyAsync(d16.p, codes, (size_t) total * 2, cudaMemcpyHostToDevice, s));
    NDB_CUDA(cudaMemsetAsync(dbad.p, 0x7f, 4, s));
    pq_widen_codes_kernel<<<(unsigned) ((total + 255) / 256), 256, 0, s>>>(d16.as<int16_t>(), total, pq->ksub,
                                                                          pq->codes.as<uint8_t>() + (size_t) pq->n * pq->m, dbad.as<int>());
    count_launch();
    int bad = 0;
    NDB_CUDA(cudaMemcpyAsync(&bad, dbad.p, 4, cudaMemcpyDeviceToHost, s));
    NDB_CUDA(cudaStreamSynchronize(s));
    // "Invalid PQ code %d at subspace %d (valid: 0-%d)" (:1072-1078); the rows are not appended
    NDB_REQUIRE(bad == 0x7f7f7f7f, NDB_B200_ERANGE, "Invalid PQ code %d at subspace %d (valid: 0-%d)", (int) codes[bad], bad % pq->m, pq->ksub - 1);
    pq->n += n;
    return NDB_B200_OK;
}

static int pq_stage_queries(ndb_b200_pq *pq, const float *Q, int nq, const char *who, cudaStream_t s)
{
    NDB_REQUIRE(pq && Q && nq > 0, NDB_B200_EINVAL, "%s: NULL or empty input", who);
    NDB_REQUIRE(pq->n > 0, NDB_B200_ESTATE, "%s: no encoded rows", who);
    const size_t qb = (size_t) nq * pq->dim * sizeof(float);
    NDB_CHECK(pq->qbuf.reserve(qb));
    NDB_CUDA(cudaMemcpyAsync(pq->qbuf.p, Q, qb, cudaMemcpyHostToDevice, s));
    NDB_CHECK(validate_begin(pq->qbuf.as<float>(), (int64_t) nq * pq->dim, s));
    return NDB_B200_OK;
}

int ndb_b200_pq_search(ndb_b200_pq *pq, const float *Q, int nq, int k, float *dist, int64_t *rows)
{
    NDB_CHECK(require_init());
    NDB_REQUIRE(dist && rows && k >= 1 && k <= 128, NDB_B200_EINVAL, "pq_search: bad argument (k 1..128)");
    cudaStream_t s = ctx().stream;
    NDB_CHECK(pq_stage_queries(pq, Q, nq, "pq_search", s));
    const size_t m = (size_t) nq * k;
    NDB_CHECK(pq->outd.reserve(m * sizeof(float))); NDB_CHECK(pq->outi.reserve(m * sizeof(int64_t)));
    NDB_CHECK(pq_scan(pq, pq->qbuf.as<float>(), nq, k, pq->outd.as<float>(), pq->outi.as<int64_t>(), nullptr, nullptr, s));
    NDB_CUDA(cudaMemcpyAsync(dist, pq->outd.p, m * sizeof(float), cudaMemcpyDeviceToHost, s));
    NDB_CUDA(cudaMemcpyAsync(rows, pq->outi.p, m * sizeof(int64_t), cudaMemcpyDeviceToHost, s));
    NDB_CUDA(cudaStreamSynchronize(s));
    const int64_t bad = validate_end();
    NDB_REQUIRE(bad < 0, NDB_B200_EVECTOR, "vector contains NaN or Infinity at index %lld", (long long) (bad % pq->dim));
    return NDB_B200_OK;
}

int ndb_b200_pq_search_dev(ndb_b200_pq *pq, const float *Q_dev, int nq, int k, float *dist_dev, int64_t *rows_dev, void *stream)
{
    NDB_CHECK(require_init());
    NDB_REQUIRE(pq && Q_dev && dist_dev && rows_dev && nq > 0 && k >= 1 && k <= 128, NDB_B200_EINVAL, "pq_search_dev: bad argument (k 1..128)");
    NDB_REQUIRE(pq->n > 0, NDB_B200_ESTATE, "pq_search_dev: no encoded rows");
    return pq_scan(pq, Q_dev, nq, k, dist_dev, rows_dev, nullptr, nullptr, stream ? (cudaStream_t) stream : ctx().stream);
}

int ndb_b200_pq_distances(ndb_b200_pq *pq, const float *Q, int nq, float *dist, unsigned long long *rechecked)
{
    NDB_CHECK(require_init());
    NDB_REQUIRE(dist, NDB_B200_EINVAL, "pq_distances: NULL output");
    cudaStream_t s = ctx().stream;
    NDB_CHECK(pq_stage_queries(pq, Q, nq, "pq_distances", s));
    const size_t bytes = (size_t) nq * pq->n * sizeof(float), cnt_off = (bytes + 7) & ~(size_t) 7;
    NDB_CHECK(pq->all.reserve(cnt_off + 8));
    unsigned long long *cnt = reinterpret_cast<unsigned long long *>(pq->all.as<char>() + cnt_off);
    NDB_CUDA(cudaMemsetAsync(cnt, 0, 8, s));
    NDB_CHECK(pq_scan(pq, pq->qbuf.as<float>(), nq, 0, nullptr, nullptr, pq->all.as<float>(), cnt, s));
    NDB_CUDA(cudaMemcpyAsync(dist, pq->all.p, bytes, cudaMemcpyDeviceToHost, s));
    unsigned long long c = 0;
    NDB_CUDA(cudaMemcpyAsync(&c, cnt, 8, cudaMemcpyDeviceToHost, s));
    NDB_CUDA(cudaStreamSynchronize(s));
    if (rechecked) *rechecked = c;
    const int64_t bad = validate_end();
    NDB_REQUIRE(bad < 0, NDB_B200_EVECTOR, "vector contains NaN or Infinity at index %lld", (long long) (bad % pq->dim));
    return NDB_B200_OK;
}

}  // extern "C"
