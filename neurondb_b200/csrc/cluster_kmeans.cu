// cluster_kmeans.cu -- the SQL function cluster_kmeans (NeuronDB/src/ml/ml_kmeans.c:146-303) on the device:
// k-means++ seeding (kmeanspp_init, :45-139) followed by Lloyd's iterations until no assignment changes.
//
// Literal semantics, reproduced bit for bit (oracle: orc_cluster_kmeans, pinned to the reference's own functions):
//   seeding : D^2 weight of a row = sum_f64 of (float)((x-c)*(x-c)) with a FLOAT difference and a FLOAT square (:64-74,
//             :126-134); the next seed is the first unselected row at which r = rand()/RAND_MAX * sum, reduced by the
//             weights IN ROW ORDER, reaches <= 0 (:85-101).  Both the sum and the walk are sequential fp64 chains whose
//             rounding decides the row.  They are NOT run sequentially in the common case: a parallel prefix sum of the
//             weights gives every row's r within a proven distance (16 n 2^-53 of the total) of the value the
//             sequential walk holds there, and the one row whose predecessor is above +eps and which itself is below
//             -eps is the reference's pick (the walk's r never increases, so no earlier row can have stopped it).  Only
//             when no row is that clear (r = 0, a total of 0, a crossing inside the band) does one thread walk the tiles
//             the block stages in shared memory -- the literal loop.  The weights themselves (n*dim work per seed) are
//             one thread per row.
//   assign  : argmin_c neurondb_l2_distance_squared (util/neurondb_simd_impl.c:36-104, the build without AVX2: double
//             difference, double square, double sum), strict <, lowest index wins (:236-259)
//   update  : float sums over the members in row order / count, empty clusters end at zero (:261-281) -- this is
//             kmeans_update_dev of the IVF trainer, which has the same semantics
// The caller supplies the k values rand() returned (the function draws exactly k, nothing else in between), so the
// process's rand() stream stays where the reference leaves it.
//
// Layout: the rows live twice on the device, row-major (update: threads along the dimensions of one member) and
// transposed [dim][n] (weights and assignment: threads along the rows, every load coalesced, each thread walking the
// dimensions of its row in order).  Roofline of the assignment: 3 rounded fp64 instructions per (row, cluster,
// dimension) -- fp64-issue bound like the operator scan (C1), not HBM bound (the transposed rows are re-read from L1/L2).
#include <cfloat>
#include <cstdlib>
#include <cub/cub.cuh>

#include "kmeans.cuh"

namespace ndb {

__global__ void __launch_bounds__(256) ckm_transpose_kernel(const float *__restrict__ X, int64_t n, int dim, float *__restrict__ XT)
{
    __shared__ float tile[32][33];
    const int64_t r0 = (int64_t) blockIdx.x * 32;
    const int d0 = blockIdx.y * 32;
    for (int j = threadIdx.y; j < 32; j += 8) {
        const int64_t r = r0 + j;
        const int d = d0 + threadIdx.x;
        if (r < n && d < dim) tile[j][threadIdx.x] = X[(size_t) r * dim + d];
    }
    __syncthreads();
    for (int j = threadIdx.y; j < 32; j += 8) {
        const int d = d0 + j;
        const int64_t r = r0 + threadIdx.x;
        if (r < n && d < dim) XT[(size_t) d * n + r] = tile[threadIdx.x][j];
    }
}

// D^2 weight against seed c (already in seeds[c]): dist[i] = acc for the first seed, min(dist[i], acc) afterwards.
// ALL_DOUBLE: minibatch_kmeans_pp_init's arithmetic (ml_minibatch_kmeans.c:92-103, 156-170: double difference and square)
// instead of kmeanspp_init's (float difference, float square).
template <bool ALL_DOUBLE>
__global__ void __launch_bounds__(256) ckm_weight_kernel(const float *__restrict__ XT, const float *__restrict__ X, int64_t n, int dim,
                                                         const int *__restrict__ seeds, int c, double *__restrict__ dist)
{
    const int64_t i = (int64_t) blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float *crow = X + (size_t) seeds[c] * dim;
    double acc = 0.0;
    for (int d = 0; d < dim; d++) {
        if (ALL_DOUBLE) {
            const double diff = __dsub_rn((double) XT[(size_t) d * n + i], (double) __ldg(crow + d));
            acc = __dadd_rn(acc, __dmul_rn(diff, diff));
        } else {
            const float diff = __fsub_rn(XT[(size_t) d * n + i], __ldg(crow + d));
            acc = __dadd_rn(acc, (double) __fmul_rn(diff, diff));
        }
    }
    if (c == 0 || acc < dist[i]) dist[i] = acc;
}

constexpr int CKM_NONE = 0x7f7f7f7f;

__global__ void ckm_mask_kernel(const double *__restrict__ dist, const unsigned char *__restrict__ selected, int64_t n, double *__restrict__ w)
{
    const int64_t i = (int64_t) blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) w[i] = selected[i] ? 0.0 : dist[i];
}

// P = inclusive prefix sums of the masked weights (any association).  The sequential walk's r at row j and
// q_j = r0 - P[j] differ by less than 8 n u T (u = 2^-53, T = the total), so a row with q_{i-1} >= eps and
// q_i <= -eps, eps = 16 (n + 16) u T, is where the walk stops.  At most one row qualifies.
__global__ void ckm_pick_cert_kernel(const double *__restrict__ P, const unsigned char *__restrict__ selected, int64_t n,
                                     int draw, double rand_max, int c, double rel_eps, int *__restrict__ cert)
{
    const int64_t i = (int64_t) blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n || selected[i]) return;
    const double T = P[n - 1];
    const double eps = rel_eps * T;
    if (!(eps > 0.0)) return;
    const double r0 = __dmul_rn(__ddiv_rn((double) draw, rand_max), T);
    const double before = i > 0 ? P[i - 1] : 0.0;
    if (r0 - before >= eps && P[i] - r0 >= eps) atomicMin(cert + c, (int) i);
}

// seed c: c == 0 -> draws[0] % n (:59); otherwise the D^2-weighted draw (:78-119): the certified row if there is one,
// else the literal walk.  One block.
constexpr int CKM_TILE = 4096;
__global__ void __launch_bounds__(1024) ckm_pick_kernel(const double *__restrict__ dist, unsigned char *__restrict__ selected, int64_t n,
                                                         int draw, double rand_max, int c, int *__restrict__ seeds,
                                                         const int *__restrict__ cert, unsigned long long *__restrict__ walked)
{
    __shared__ double buf[CKM_TILE];
    __shared__ unsigned char sel[CKM_TILE];
    __shared__ long long s_picked;
    __shared__ unsigned long long s_first;
    if (c == 0) {
        if (threadIdx.x == 0) {
            const int first = draw % (int) n;
            seeds[0] = first;
            selected[first] = 1;
        }
        return;
    }
    if (cert && cert[c] != CKM_NONE) {
        if (threadIdx.x == 0) {
            seeds[c] = cert[c];
            selected[cert[c]] = 1;
        }
        return;
    }
    if (threadIdx.x == 0 && walked) atomicAdd(walked, 1ull);
    // sum of the unselected weights in row order (adding 0.0 for a selected row leaves a non-negative sum unchanged)
    double sum = 0.0;
    for (int64_t base = 0; base < n; base += CKM_TILE) {
        const int m = (int) (n - base < CKM_TILE ? n - base : CKM_TILE);
        for (int t = threadIdx.x; t < m; t += blockDim.x) buf[t] = selected[base + t] ? 0.0 : dist[base + t];
        __syncthreads();
        if (threadIdx.x == 0) {
#pragma unroll 8
            for (int t = 0; t < m; t++) sum = __dadd_rn(sum, buf[t]);
        }
        __syncthreads();
    }
    if (threadIdx.x == 0) { s_picked = -1; s_first = ~0ull; }
    double r = __dmul_rn(__ddiv_rn((double) draw, rand_max), sum);
    __syncthreads();
    for (int64_t base = 0; base < n; base += CKM_TILE) {
        const int m = (int) (n - base < CKM_TILE ? n - base : CKM_TILE);
        for (int t = threadIdx.x; t < m; t += blockDim.x) { buf[t] = dist[base + t]; sel[t] = selected[base + t]; }
        __syncthreads();
        if (threadIdx.x == 0) {
            for (int t = 0; t < m; t++) {
                if (sel[t]) continue;
                r = __dsub_rn(r, buf[t]);
                if (r <= 0) { s_picked = base + t; break; }
            }
        }
        __syncthreads();
        if (s_picked >= 0) break;
    }
    if (s_picked < 0) {                                   // rounding left r > 0 after the last row: first unselected row (:103-113)
        unsigned long long mine = ~0ull;
        for (int64_t i = threadIdx.x; i < n; i += blockDim.x)
            if (!selected[i]) { mine = (unsigned long long) i; break; }
        atomicMin(&s_first, mine);
        __syncthreads();
        if (threadIdx.x == 0) s_picked = (long long) s_first;
        __syncthreads();
    }
    if (threadIdx.x == 0) {
        seeds[c] = (int) s_picked;
        selected[s_picked] = 1;
    }
}

__global__ void ckm_gather_centers_kernel(const float *__restrict__ X, const int *__restrict__ seeds, int dim, float *__restrict__ C)
{
    const int c = blockIdx.x;
    for (int d = threadIdx.x; d < dim; d += blockDim.x) C[(size_t) c * dim + d] = X[(size_t) seeds[c] * dim + d];
}

__global__ void ckm_widen_kernel(const float *__restrict__ in, int64_t n, double *__restrict__ out)
{
    const int64_t i = (int64_t) blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = (double) in[i];
}

// nearest center of each row in fp64, CPT centers per pass over the row (independent chains hide the fp64 latency,
// the row element is converted once for all of them; the centres arrive already widened -- the conversion of a centre
// element is the same for every row, and F2F shares the slow pipe with the fp64 arithmetic)
template <int CPT>
__global__ void __launch_bounds__(128) ckm_assign_kernel(const float *__restrict__ XT, const double *__restrict__ C, int64_t n, int dim, int k,
                                                         int *__restrict__ assign, int *__restrict__ changed)
{
    const int64_t i = (int64_t) blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t il = i < n ? i : n - 1;
    int best = -1;
    double min_dist = DBL_MAX;
    for (int c0 = 0; c0 < k; c0 += CPT) {
        double acc[CPT];
        const double *crow[CPT];
#pragma unroll
        for (int j = 0; j < CPT; j++) {
            acc[j] = 0.0;
            crow[j] = C + (size_t) (c0 + j < k ? c0 + j : k - 1) * dim;
        }
        for (int d = 0; d < dim; d++) {
            const double x = (double) XT[(size_t) d * n + il];
#pragma unroll
            for (int j = 0; j < CPT; j++) {
                const double diff = __dsub_rn(x, __ldg(crow[j] + d));
                acc[j] = __dadd_rn(acc[j], __dmul_rn(diff, diff));
            }
        }
#pragma unroll
        for (int j = 0; j < CPT; j++)
            if (c0 + j < k && acc[j] < min_dist) { min_dist = acc[j]; best = c0 + j; }
    }
    if (i < n && assign[i] != best) {
        assign[i] = best;
        *changed = 1;
    }
}

// the reference's `for (i...) if (!selected[i]) sum += dist[i]` as written: one thread, tiles staged by the block
__global__ void __launch_bounds__(1024) ckm_seq_sum_kernel(const double *__restrict__ dist, const unsigned char *__restrict__ selected, int64_t n,
                                                            double *__restrict__ out)
{
    __shared__ double buf[CKM_TILE];
    double sum = 0.0;
    for (int64_t base = 0; base < n; base += CKM_TILE) {
        const int m = (int) (n - base < CKM_TILE ? n - base : CKM_TILE);
        for (int t = threadIdx.x; t < m; t += blockDim.x) buf[t] = selected[base + t] ? 0.0 : dist[base + t];
        __syncthreads();
        if (threadIdx.x == 0)
            for (int t = 0; t < m; t++) sum = __dadd_rn(sum, buf[t]);
        __syncthreads();
    }
    if (threadIdx.x == 0) *out = sum;
}

// cluster_minibatch_kmeans' centroid update (ml_minibatch_kmeans.c:383-403) for one batch: the batch members of a cluster
// in batch order (members / start from group_by_cluster_dev), one thread per (cluster, dimension):
// count++, eta = 1 / count, c = (float) ((1 - eta) c + eta x) in double.  counts[] is advanced by mbk_counts_kernel afterwards.
__global__ void mbk_update_kernel(const float *__restrict__ B, const uint32_t *__restrict__ members, const int *__restrict__ start, int dim,
                                  float *__restrict__ C, const int *__restrict__ counts)
{
    const int c = blockIdx.x;
    const int b = start[c], e = start[c + 1];
    const int j = blockIdx.y * blockDim.x + threadIdx.x;
    if (b == e || j >= dim) return;
    int cnt = counts[c];
    float cv = C[(size_t) c * dim + j];
    for (int t = b; t < e; t++) {
        cnt++;
        const double lr = __ddiv_rn(1.0, (double) cnt);
        cv = __double2float_rn(__dadd_rn(__dmul_rn(__dsub_rn(1.0, lr), (double) cv), __dmul_rn(lr, (double) B[(size_t) members[t] * dim + j])));
    }
    C[(size_t) c * dim + j] = cv;
}

__global__ void mbk_counts_kernel(const int *__restrict__ start, int k, int *__restrict__ counts)
{
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c < k) counts[c] += start[c + 1] - start[c];
}

// host rows -> device, NaN / Inf scan on the device (the 2 * dim tests per row of the reference's check, off the host)
int upload_rows_checked(DevBuf &dst, const float *X, size_t count, const char *who, cudaStream_t s)
{
    NDB_CHECK(dst.reserve(count * sizeof(float)));
    NDB_CUDA(cudaMemcpyAsync(dst.p, X, count * sizeof(float), cudaMemcpyHostToDevice, s));
    NDB_CHECK(validate_begin(dst.as<float>(), (int64_t) count, s));
    NDB_CUDA(cudaStreamSynchronize(s));
    NDB_REQUIRE(validate_end() < 0, NDB_B200_EVECTOR, "%s: NaN/Inf in the vectors", who);
    return NDB_B200_OK;
}

int transpose_rows_dev(const float *dX, int64_t n, int dim, float *dXT, cudaStream_t s)
{
    ckm_transpose_kernel<<<dim3((unsigned) ((n + 31) / 32), (unsigned) ((dim + 31) / 32)), dim3(32, 8), 0, s>>>(dX, n, dim, dXT);
    count_launch();
    NDB_CUDA(cudaGetLastError());
    return NDB_B200_OK;
}

int gather_rows_dev(const float *dX, const int *rows_dev, int nrows, int dim, float *out, cudaStream_t s)
{
    ckm_gather_centers_kernel<<<nrows, 128, 0, s>>>(dX, rows_dev, dim, out);
    count_launch();
    NDB_CUDA(cudaGetLastError());
    return NDB_B200_OK;
}

int nearest_f64_dev(const float *dXT, const float *dC, int64_t n, int dim, int k, int *assign, int *changed, DevBuf &wide, cudaStream_t s)
{
    const int64_t cn = (int64_t) k * dim;
    NDB_CHECK(wide.reserve((size_t) cn * sizeof(double)));
    ckm_widen_kernel<<<(unsigned) ((cn + 255) / 256), 256, 0, s>>>(dC, cn, wide.as<double>());
    ckm_assign_kernel<4><<<(unsigned) ((n + 127) / 128), 128, 0, s>>>(dXT, wide.as<double>(), n, dim, k, assign, changed);
    count_launch(2);
    NDB_CUDA(cudaGetLastError());
    return NDB_B200_OK;
}

int lloyd_f64_dev(KMeansWork &w, const float *dX, const float *dXT, int64_t n, int dim, int k, int max_iters, bool stop_before_update,
                  int *dchanged, int *iters, cudaStream_t s)
{
    int changed = 1, iter = 0;
    for (iter = 0; iter < max_iters && (stop_before_update || changed); iter++) {
        NDB_CUDA(cudaMemsetAsync(dchanged, 0, 4, s));
        NDB_CHECK(nearest_f64_dev(dXT, w.C.as<float>(), n, dim, k, w.assign.as<int>(), dchanged, w.dcost, s));
        NDB_CUDA(cudaMemcpyAsync(&changed, dchanged, 4, cudaMemcpyDeviceToHost, s));
        if (stop_before_update) {
            NDB_CUDA(cudaStreamSynchronize(s));
            if (!changed) break;
        }
        NDB_CHECK(kmeans_update_dev(w, dX, w.assign.as<int>(), n, dim, k, w.C.as<float>(), w.counts.as<int>(), s));
        NDB_CUDA(cudaStreamSynchronize(s));
    }
    if (iters) *iters = iter;
    return NDB_B200_OK;
}

}  // namespace ndb

using namespace ndb;

extern "C" {

int ndb_b200_cluster_kmeans(const float *X, int n, int dim, int k, int max_iters, const int *rand_draws, int rand_max,
                            int *labels, float *centers, int *iters, int *seeds)
{
    NDB_CHECK(require_init());
    NDB_REQUIRE(X && labels && rand_draws && n > 0 && dim > 0 && rand_max > 0, NDB_B200_EINVAL, "cluster_kmeans: bad argument");
    NDB_REQUIRE(k > 1, NDB_B200_EINVAL, "number of clusters must be at least 2");                                        // :170-174
    NDB_REQUIRE(n >= k, NDB_B200_EINVAL, "not enough vectors for cluster count (need >= %d, have %d)", k, n);           // :180-186
    for (int c = 0; c < k; c++) NDB_REQUIRE(rand_draws[c] >= 0, NDB_B200_EINVAL, "cluster_kmeans: rand() values are non-negative");
    if (max_iters < 1) max_iters = 100;                                                                                 // :175-176
    cudaStream_t s = ctx().stream;
    KMeansWork w;
    DevBuf XT, dist, selected, dseeds, dchanged;
    const size_t xb = (size_t) n * dim * 4, cb = (size_t) k * dim * 4;
    NDB_CHECK(upload_rows_checked(w.X, X, (size_t) n * dim, "cluster_kmeans", s));
    NDB_CHECK(XT.reserve(xb)); NDB_CHECK(w.C.reserve(cb));
    NDB_CHECK(w.assign.reserve((size_t) n * 4)); NDB_CHECK(w.counts.reserve((size_t) k * 4));
    NDB_CHECK(dist.reserve((size_t) n * 8)); NDB_CHECK(selected.reserve((size_t) n));
    NDB_CHECK(dseeds.reserve((size_t) k * 4)); NDB_CHECK(dchanged.reserve(4));
    NDB_CUDA(cudaMemsetAsync(selected.p, 0, (size_t) n, s));
    const float *dX = w.X.as<float>();
    NDB_CHECK(transpose_rows_dev(dX, n, dim, XT.as<float>(), s));
    const unsigned rb = (unsigned) ((n + 255) / 256);
    const bool literal_walk = getenv("NDB_CKM_SEQUENTIAL") != nullptr;       // (test / measurement switch: always the literal loop)
    DevBuf masked, prefix, cert, cub_tmp, walked;
    size_t tmp_bytes = 0;
    NDB_CHECK(walked.reserve(8));
    NDB_CUDA(cudaMemsetAsync(walked.p, 0, 8, s));
    if (!literal_walk) {
        NDB_CHECK(masked.reserve((size_t) n * 8)); NDB_CHECK(prefix.reserve((size_t) n * 8)); NDB_CHECK(cert.reserve((size_t) k * 4));
        NDB_CUDA(cub::DeviceScan::InclusiveSum(nullptr, tmp_bytes, masked.as<double>(), prefix.as<double>(), n, s));
        NDB_CHECK(cub_tmp.reserve(tmp_bytes));
        NDB_CUDA(cudaMemsetAsync(cert.p, 0x7f, (size_t) k * 4, s));
    }
    const double rel_eps = 16.0 * (double) (n + 16) * 1.1102230246251565e-16;
    for (int c = 0; c < k; c++) {
        if (c > 0 && !literal_walk) {
            ckm_mask_kernel<<<rb, 256, 0, s>>>(dist.as<double>(), selected.as<unsigned char>(), n, masked.as<double>());
            NDB_CUDA(cub::DeviceScan::InclusiveSum(cub_tmp.p, tmp_bytes, masked.as<double>(), prefix.as<double>(), n, s));
            ckm_pick_cert_kernel<<<rb, 256, 0, s>>>(prefix.as<double>(), selected.as<unsigned char>(), n, rand_draws[c], (double) rand_max, c,
                                                  rel_eps, cert.as<int>());
            count_launch(4);
        }
        ckm_pick_kernel<<<1, 1024, 0, s>>>(dist.as<double>(), selected.as<unsigned char>(), n, rand_draws[c], (double) rand_max, c,
                                          dseeds.as<int>(), literal_walk ? nullptr : cert.as<int>(), walked.as<unsigned long long>());
        if (c + 1 < k)            // the weights against the last seed are never read (:121-135 computes them all the same)
            ckm_weight_kernel<false><<<rb, 256, 0, s>>>(XT.as<float>(), dX, n, dim, dseeds.as<int>(), c, dist.as<double>());
        count_launch(2);
    }
    unsigned long long n_walked = 0;
    NDB_CUDA(cudaMemcpyAsync(&n_walked, walked.p, 8, cudaMemcpyDeviceToHost, s));
    NDB_CHECK(gather_rows_dev(dX, dseeds.as<int>(), k, dim, w.C.as<float>(), s));
    NDB_CUDA(cudaMemsetAsync(w.assign.p, 0xff, (size_t) n * 4, s));                                                    // assignments[i] = -1 (:206-207)
    int iter = 0;
    NDB_CHECK(lloyd_f64_dev(w, dX, XT.as<float>(), n, dim, k, max_iters, false, dchanged.as<int>(), &iter, s));           // :226-278
    NDB_CUDA(cudaMemcpyAsync(labels, w.assign.p, (size_t) n * 4, cudaMemcpyDeviceToHost, s));
    if (centers) NDB_CUDA(cudaMemcpyAsync(centers, w.C.p, cb, cudaMemcpyDeviceToHost, s));
    if (seeds) NDB_CUDA(cudaMemcpyAsync(seeds, dseeds.p, (size_t) k * 4, cudaMemcpyDeviceToHost, s));
    NDB_CUDA(cudaStreamSynchronize(s));
    ctx().last_ms = 0.0; ctx().last_bytes = 0.0; ctx().stats_src = nullptr;
    ctx().last_evals = (int64_t) n_walked;                 // ndb_b200_last_kernel_stats: seeds that needed the literal walk
    for (int i = 0; i < n; i++) labels[i] += 1;                                                                        // 1-based labels (:286)
    if (iters) *iters = iter;
    return NDB_B200_OK;
}

// cluster_minibatch_kmeans (ml_minibatch_kmeans.c:206-449).  The number of rand() calls depends on the data (the seeding
// stops when the remaining weights sum below 1e-10), so the caller hands over a function that IS its rand(): it is called
// exactly when and as often as the reference calls rand().
int ndb_b200_cluster_minibatch_kmeans(const float *X, int n, int dim, int k, int batch_size, int max_iters, ndb_b200_rand_fn next_rand,
                                      void *rand_state, int rand_max, int *labels, float *centers)
{
    NDB_CHECK(require_init());
    NDB_REQUIRE(X && labels && next_rand && rand_max > 0, NDB_B200_EINVAL, "cluster_minibatch_kmeans: bad argument");
    NDB_REQUIRE(k >= 2, NDB_B200_EINVAL, "num_clusters must be at least 2");                                             // :238-241
    NDB_REQUIRE(batch_size >= 1, NDB_B200_EINVAL, "batch_size must be at least 1");                                      // :242-245
    if (max_iters < 1) max_iters = 100;                                                                                 // :246-247
    NDB_REQUIRE(n > 0, NDB_B200_EINVAL, "No vectors found");
    NDB_REQUIRE(dim > 0, NDB_B200_EINVAL, "Invalid vector dimension: %d", dim);
    NDB_REQUIRE(n >= k, NDB_B200_EINVAL, "Not enough vectors (%d) for %d clusters", n, k);                               // :296-300
    if (batch_size > n) batch_size = n;                                                                                 // :303-304
    cudaStream_t s = ctx().stream;
    KMeansWork w;
    DevBuf XT, dist, selected, dseeds, dchanged, masked, prefix, cert, cub_tmp, walked, dsum, didx, B, BT, bassign;
    const size_t xb = (size_t) n * dim * 4, cb = (size_t) k * dim * 4;
    NDB_CHECK(upload_rows_checked(w.X, X, (size_t) n * dim, "cluster_minibatch_kmeans", s));
    NDB_CHECK(XT.reserve(xb)); NDB_CHECK(w.C.reserve(cb)); NDB_CHECK(w.assign.reserve((size_t) n * 4)); NDB_CHECK(w.counts.reserve((size_t) k * 4));
    NDB_CHECK(dist.reserve((size_t) n * 8)); NDB_CHECK(selected.reserve((size_t) n)); NDB_CHECK(dseeds.reserve((size_t) k * 4));
    NDB_CHECK(dchanged.reserve(4)); NDB_CHECK(masked.reserve((size_t) n * 8)); NDB_CHECK(prefix.reserve((size_t) n * 8));
    NDB_CHECK(cert.reserve((size_t) k * 4)); NDB_CHECK(walked.reserve(8)); NDB_CHECK(dsum.reserve(8));
    NDB_CHECK(didx.reserve((size_t) batch_size * 4)); NDB_CHECK(B.reserve((size_t) batch_size * dim * 4));
    NDB_CHECK(BT.reserve((size_t) batch_size * dim * 4)); NDB_CHECK(bassign.reserve((size_t) batch_size * 4));
    size_t tmp_bytes = 0;
    NDB_CUDA(cub::DeviceScan::InclusiveSum(nullptr, tmp_bytes, masked.as<double>(), prefix.as<double>(), n, s));
    NDB_CHECK(cub_tmp.reserve(tmp_bytes));
    NDB_CUDA(cudaMemsetAsync(cert.p, 0x7f, (size_t) k * 4, s));
    NDB_CUDA(cudaMemsetAsync(walked.p, 0, 8, s));
    NDB_CUDA(cudaMemsetAsync(selected.p, 0, (size_t) n, s));
    NDB_CUDA(cudaMemsetAsync(w.C.p, 0, cb, s));                                 // centroids the seeding never reaches stay zero
    NDB_CUDA(cudaMemsetAsync(w.counts.p, 0, (size_t) k * 4, s));
    const float *dX = w.X.as<float>();
    NDB_CHECK(transpose_rows_dev(dX, n, dim, XT.as<float>(), s));
    const unsigned rb = (unsigned) ((n + 255) / 256);
    const double rel_eps = 16.0 * (double) (n + 16) * 1.1102230246251565e-16;
    // ---- minibatch_kmeans_pp_init (:67-198)
    for (int c = 0; c < k; c++) {
        if (c > 0) {
            ckm_mask_kernel<<<rb, 256, 0, s>>>(dist.as<double>(), selected.as<unsigned char>(), n, masked.as<double>());
            NDB_CUDA(cub::DeviceScan::InclusiveSum(cub_tmp.p, tmp_bytes, masked.as<double>(), prefix.as<double>(), n, s));
            count_launch(3);
            double total = 0.0;
            NDB_CUDA(cudaMemcpyAsync(&total, prefix.as<double>() + (n - 1), 8, cudaMemcpyDeviceToHost, s));
            NDB_CUDA(cudaStreamSynchronize(s));
            // `if (sum < 1e-10) break;` (:114-115) on the sequential sum: the parallel total decides unless it lies within the
            // two summations' distance of the threshold, in which case the loop is run as written
            const double band = rel_eps * total;
            if (total + band >= 1e-10 && total - band < 1e-10) {
                ckm_seq_sum_kernel<<<1, 1024, 0, s>>>(dist.as<double>(), selected.as<unsigned char>(), n, dsum.as<double>());
                count_launch();
                NDB_CUDA(cudaMemcpyAsync(&total, dsum.p, 8, cudaMemcpyDeviceToHost, s));
                NDB_CUDA(cudaStreamSynchronize(s));
            }
            if (total < 1e-10) break;
        }
        const int draw = next_rand(rand_state);
        NDB_REQUIRE(draw >= 0, NDB_B200_EINVAL, "cluster_minibatch_kmeans: rand() values are non-negative");
        if (c > 0) {
            ckm_pick_cert_kernel<<<rb, 256, 0, s>>>(prefix.as<double>(), selected.as<unsigned char>(), n, draw, (double) rand_max, c, rel_eps,
                                                  cert.as<int>());
            count_launch();
        }
        ckm_pick_kernel<<<1, 1024, 0, s>>>(dist.as<double>(), selected.as<unsigned char>(), n, draw, (double) rand_max, c, dseeds.as<int>(),
                                          cert.as<int>(), walked.as<unsigned long long>());
        count_launch();
        NDB_CHECK(gather_rows_dev(dX, dseeds.as<int>() + c, 1, dim, w.C.as<float>() + (size_t) c * dim, s));
        if (c + 1 < k) {
            ckm_weight_kernel<true><<<rb, 256, 0, s>>>(XT.as<float>(), dX, n, dim, dseeds.as<int>(), c, dist.as<double>());
            count_launch();
        }
    }
    NDB_CUDA(cudaGetLastError());
    // ---- the mini-batch loop (:347-410)
    std::vector<int> idx((size_t) batch_size);
    const dim3 ugrid((unsigned) k, (unsigned) ((dim + 127) / 128));
    for (int iter = 0; iter < max_iters; iter++) {
        for (int i = 0; i < batch_size; i++) {
            const int draw = next_rand(rand_state);
            NDB_REQUIRE(draw >= 0, NDB_B200_EINVAL, "cluster_minibatch_kmeans: rand() values are non-negative");
            idx[i] = draw % n;
        }
        NDB_CUDA(cudaMemcpyAsync(didx.p, idx.data(), (size_t) batch_size * 4, cudaMemcpyHostToDevice, s));
        NDB_CUDA(cudaStreamSynchronize(s));                                    // (idx is refilled by the next iteration)
        NDB_CHECK(gather_rows_dev(dX, didx.as<int>(), batch_size, dim, B.as<float>(), s));
        NDB_CHECK(transpose_rows_dev(B.as<float>(), batch_size, dim, BT.as<float>(), s));
        NDB_CUDA(cudaMemsetAsync(bassign.p, 0xff, (size_t) batch_size * 4, s));
        NDB_CHECK(nearest_f64_dev(BT.as<float>(), w.C.as<float>(), batch_size, dim, k, bassign.as<int>(), dchanged.as<int>(), w.dcost, s));
        NDB_CHECK(group_by_cluster_dev(w, bassign.as<int>(), batch_size, k, s));
        mbk_update_kernel<<<ugrid, 128, 0, s>>>(B.as<float>(), w.vals_sorted.as<uint32_t>(), w.start.as<int>(), dim, w.C.as<float>(),
                                              w.counts.as<int>());
        mbk_counts_kernel<<<(unsigned) ((k + 255) / 256), 256, 0, s>>>(w.start.as<int>(), k, w.counts.as<int>());
        count_launch(2);
    }
    NDB_CUDA(cudaGetLastError());
    // ---- final assignment of every row (:412-437)
    NDB_CUDA(cudaMemsetAsync(w.assign.p, 0xff, (size_t) n * 4, s));
    NDB_CHECK(nearest_f64_dev(XT.as<float>(), w.C.as<float>(), n, dim, k, w.assign.as<int>(), dchanged.as<int>(), w.dcost, s));
    NDB_CUDA(cudaMemcpyAsync(labels, w.assign.p, (size_t) n * 4, cudaMemcpyDeviceToHost, s));
    if (centers) NDB_CUDA(cudaMemcpyAsync(centers, w.C.p, cb, cudaMemcpyDeviceToHost, s));
    NDB_CUDA(cudaStreamSynchronize(s));
    for (int i = 0; i < n; i++) labels[i] += 1;                                                                        // 1-based (:441-442)
    return NDB_B200_OK;
}

}  // extern "C"
