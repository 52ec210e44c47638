// core.cu -- context, memory helpers, paired-row distance kernels and the ndb_gpu_backend
// vtable launchers (neurondb_gpu_backend.h:38-79) of libndb_b200.so.
#include "common.cuh"
#include "arith.cuh"
#include "layout.cuh"
#include "comm.cuh"
#include "kmeans.cuh"

#include <cstdarg>
#include <unistd.h>

namespace ndb {

thread_local char g_last_error[512] = "";

void set_error(const char *fmt, ...)
{
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_last_error, sizeof(g_last_error), fmt, ap);
    va_end(ap);
}

static Context g_ctx;
static uint64_t g_generation = 0;
static pid_t g_ctx_pid = 0;
Context &ctx() { return g_ctx; }

int require_init()
{
    // a fork()ed PostgreSQL backend must not reuse the parent's CUDA state
    // (the reference does the same check: gpu_backend_cuda.c:167-184)
    if (g_ctx.initialized && g_ctx_pid != getpid()) {
        g_ctx = Context();
    }
    NDB_REQUIRE(g_ctx.initialized, NDB_B200_ENOTINIT, "ndb_b200_init() has not been called in this process");
    cudaError_t e = cudaSetDevice(g_ctx.device);
    NDB_REQUIRE(e == cudaSuccess, NDB_B200_ECUDA, "cudaSetDevice(%d): %s", g_ctx.device, cudaGetErrorString(e));
    return NDB_B200_OK;
}

int64_t find_nonfinite(const float *v, int64_t n)
{
    // blocks of 4096 without an early exit (vectorises), then locate inside the offending block
    for (int64_t b = 0; b < n; b += 4096) {
        const int64_t e = b + 4096 < n ? b + 4096 : n;
        uint32_t any = 0;
        for (int64_t i = b; i < e; i++) {
            uint32_t u;
            memcpy(&u, v + i, 4);
            any |= (u & 0x7f800000u) == 0x7f800000u;
        }
        if (any)
            for (int64_t i = b; i < e; i++) {
                uint32_t u;
                memcpy(&u, v + i, 4);
                if ((u & 0x7f800000u) == 0x7f800000u) return i;
            }
    }
    return -1;
}

__global__ void nonfinite_kernel(const float *__restrict__ v, int64_t n, unsigned long long *__restrict__ bad)
{
    unsigned long long first = ~0ull;
    for (int64_t i = (int64_t) blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t) gridDim.x * blockDim.x)
        if ((__float_as_uint(v[i]) & 0x7f800000u) == 0x7f800000u && (unsigned long long) i < first) first = (unsigned long long) i;
    if (first != ~0ull) atomicMin(bad, first);
}

int validate_begin(const float *v_dev, int64_t n, cudaStream_t s)
{
    Context &c = ctx();
    unsigned long long *h = reinterpret_cast<unsigned long long *>(static_cast<char *>(c.pinned) + 256);
    *h = ~0ull;
    NDB_CUDA(cudaMemsetAsync(c.d_badidx, 0xFF, 8, s));
    const int64_t blocks = (n + 255) / 256;
    nonfinite_kernel<<<(unsigned) (blocks < 4 * c.sm_count ? blocks : 4 * c.sm_count), 256, 0, s>>>(v_dev, n, c.d_badidx);
    count_launch();
    NDB_CUDA(cudaGetLastError());
    NDB_CUDA(cudaMemcpyAsync(h, c.d_badidx, 8, cudaMemcpyDeviceToHost, s));
    return NDB_B200_OK;
}

int validate_into(const float *v_dev, int64_t n, unsigned long long *cell_dev, cudaStream_t s)
{
    Context &c = ctx();
    NDB_CUDA(cudaMemsetAsync(cell_dev, 0xFF, 8, s));
    const int64_t blocks = (n + 255) / 256;
    nonfinite_kernel<<<(unsigned) (blocks < 4 * c.sm_count ? blocks : 4 * c.sm_count), 256, 0, s>>>(v_dev, n, cell_dev);
    count_launch();
    NDB_CUDA(cudaGetLastError());
    return NDB_B200_OK;
}

int64_t validate_end()
{
    const unsigned long long v = *reinterpret_cast<unsigned long long *>(static_cast<char *>(ctx().pinned) + 256);
    return v == ~0ull ? -1 : (int64_t) v;
}

// ---------------------------------------------------------------------------------------
// paired rows: out[i] = dist(A_i, B_i).  One warp owns 32 pairs; a [32 rows][32 dims] tile of
// A and of B is read coalesced (128 B per row) into padded shared memory, then lane l walks
// the dimensions of pair l in order -- the reference's sequential accumulation, bit for bit.
// b_stride = 0 evaluates one query (B) against every row of A.
// ---------------------------------------------------------------------------------------
template <class P>
__global__ void __launch_bounds__(128) pairs_kernel(const float *__restrict__ A, const float *__restrict__ B,
                                                     float *__restrict__ out, int64_t n, int dim,
                                                     int64_t b_stride)
{
    using Acc = typename P::Acc;
    using NT = typename P::N;
    using QE = typename P::Q;
    __shared__ float ta[4][32][33];
    __shared__ float tb[4][32][33];
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const int64_t p0 = ((int64_t) blockIdx.x * 4 + w) * 32;
    if (p0 >= n) return;
    const int np = (int) (n - p0 < 32 ? n - p0 : 32);
    Acc acc;
    P::init(acc);
    NT na = NT(0), nb = NT(0);
    for (int c0 = 0; c0 < dim; c0 += 32) {
        const int cw = dim - c0 < 32 ? dim - c0 : 32;
        for (int r = 0; r < np; r++) {
            if (lane < cw) {
                ta[w][r][lane] = A[(size_t) (p0 + r) * dim + c0 + lane];
                tb[w][r][lane] = B[(size_t) (p0 + r) * b_stride + c0 + lane];
            }
        }
        __syncwarp();
        if (lane < np) {
            for (int j = 0; j < cw; j++) {
                const float x = ta[w][lane][j], q = tb[w][lane][j];
                P::step(acc, x, (QE) q);
                if (P::NORMS) { P::nstep(na, x); P::nstep(nb, q); }
            }
        }
        __syncwarp();
    }
    if (lane < np) out[p0 + lane] = P::finish(acc, na, nb);
}

// ---------------------------------------------------------------------------------------
// The operators as an AVX build of the reference computes them (SURVEY 8a row a5;
// vector_distance_simd.c:159-392 with the horizontal sums of :85-137): LANES (8 = AVX2, 16 =
// AVX-512) f32 lane accumulators over the first dim / LANES * LANES elements, a fixed reduction
// tree, then a scalar tail.  L2 and inner product multiply and add separately, cosine uses fmadd
// (and, built with -mfma as the AVX2 build must be, gcc contracts its scalar tail too).  Element
// i goes to lane accumulator i % LANES, which is static here because a tile column block starts
// at a multiple of 32.
// ---------------------------------------------------------------------------------------
template <int LANES> __device__ __forceinline__ float avx_hsum(const float (&v)[LANES])
{
    float t[8];
#pragma unroll
    for (int i = 0; i < 8; i++) t[i] = LANES == 16 ? __fadd_rn(v[i], v[(i + 8) % LANES]) : v[i];
    const float s0 = __fadd_rn(t[0], t[4]), s1 = __fadd_rn(t[1], t[5]), s2 = __fadd_rn(t[2], t[6]), s3 = __fadd_rn(t[3], t[7]);
    return __fadd_rn(__fadd_rn(s0, s1), __fadd_rn(s2, s3));
}

template <int METRIC, int LANES>
__global__ void __launch_bounds__(128) pairs_avx_kernel(const float *__restrict__ A, const float *__restrict__ B,
                                                         float *__restrict__ out, int64_t n, int dim, int64_t b_stride)
{
    __shared__ float ta[4][32][33];
    __shared__ float tb[4][32][33];
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const int64_t p0 = ((int64_t) blockIdx.x * 4 + w) * 32;
    if (p0 >= n) return;
    const int np = (int) (n - p0 < 32 ? n - p0 : 32);
    const int simd_end = dim / LANES * LANES;
    float d[LANES], na[LANES], nb[LANES];
#pragma unroll
    for (int l = 0; l < LANES; l++) { d[l] = 0.0f; na[l] = 0.0f; nb[l] = 0.0f; }
    float sum = 0.0f, suma = 0.0f, sumb = 0.0f;          // after the reduction tree: the scalar tail runs on these
    for (int c0 = 0; c0 < dim; c0 += 32) {
        const int cw = dim - c0 < 32 ? dim - c0 : 32;
        for (int r = 0; r < np; r++) {
            if (lane < cw) {
                ta[w][r][lane] = A[(size_t) (p0 + r) * dim + c0 + lane];
                tb[w][r][lane] = B[(size_t) (p0 + r) * b_stride + c0 + lane];
            }
        }
        __syncwarp();
        if (lane < np) {
            const int vec_cw = min(cw, simd_end - c0);               // elements of this block inside the vector body
            for (int j0 = 0; j0 + LANES <= vec_cw; j0 += LANES) {
#pragma unroll
                for (int l = 0; l < LANES; l++) {
                    const float x = ta[w][lane][j0 + l], q = tb[w][lane][j0 + l];
                    if (METRIC == NDB_L2) {
                        const float df = __fsub_rn(x, q);
                        d[l] = __fadd_rn(d[l], __fmul_rn(df, df));
                    } else if (METRIC == NDB_IP) {
                        d[l] = __fadd_rn(d[l], __fmul_rn(x, q));
                    } else {
                        d[l] = __fmaf_rn(x, q, d[l]);
                        na[l] = __fmaf_rn(x, x, na[l]);
                        nb[l] = __fmaf_rn(q, q, nb[l]);
                    }
                }
            }
            if (c0 + cw == dim) {                                    // last block: reduce, then the scalar tail
                sum = avx_hsum<LANES>(d);
                if (METRIC == NDB_COSINE) { suma = avx_hsum<LANES>(na); sumb = avx_hsum<LANES>(nb); }
                for (int j = max(simd_end - c0, 0); j < cw; j++) {
                    const float x = ta[w][lane][j], q = tb[w][lane][j];
                    if (METRIC == NDB_L2) {
                        const float df = __fsub_rn(x, q);
                        sum = __fadd_rn(sum, __fmul_rn(df, df));
                    } else if (METRIC == NDB_IP) {
                        sum = __fadd_rn(sum, __fmul_rn(x, q));
                    } else {
                        sum = __fmaf_rn(x, q, sum);
                        suma = __fmaf_rn(x, x, suma);
                        sumb = __fmaf_rn(q, q, sumb);
                    }
                }
            }
        }
        __syncwarp();
    }
    if (lane < np) {
        float r;
        if (METRIC == NDB_L2) r = __fsqrt_rn(sum);
        else if (METRIC == NDB_IP) r = sum;                           // inner_product_simd's +dot (Q3)
        else r = (suma == 0.0f || sumb == 0.0f) ? 1.0f : __fsub_rn(1.0f, __fdiv_rn(sum, __fmul_rn(__fsqrt_rn(suma), __fsqrt_rn(sumb))));
        out[p0 + lane] = r;
    }
}

// the CUDA backend's cosine (gpu_backend_cuda.c:459-537): Sdot / (Snrm2 * Snrm2), clamped to
// [-1,1], 1.0 when a norm is <= 0.  cuBLAS level-1 summation order is unspecified, so this is
// the fp32 tolerance path.
struct CosBackend {
    using Q = float; using N = float; using Acc = float;
    static constexpr bool NORMS = true;
    __device__ static void init(Acc &a) { a = 0.0f; }
    __device__ static void step(Acc &a, float x, Q q) { a = fmaf(x, q, a); }
    __device__ static void nstep(N &n, float v) { n = fmaf(v, v, n); }
    __device__ static float finish(const Acc &dot, N xn, N qn)
    {
        const float na = sqrtf(xn), nb = sqrtf(qn);
        if (na <= 0.0f || nb <= 0.0f) return 1.0f;
        float c = dot / (na * nb);
        c = c < -1.0f ? -1.0f : (c > 1.0f ? 1.0f : c);
        return 1.0f - c;
    }
};

template <class P>
static int run_pairs(const float *dA, const float *dB, float *dOut, int64_t n, int dim, int64_t b_stride,
                     cudaStream_t s)
{
    const int64_t blocks = (n + 127) / 128;
    pairs_kernel<P><<<(unsigned) blocks, 128, 0, s>>>(dA, dB, dOut, n, dim, b_stride);
    count_launch();
    NDB_CUDA(cudaGetLastError());
    return NDB_B200_OK;
}

int launch_pairs(int metric, int arith, const float *dA, const float *dB, float *dOut, int64_t n, int dim,
                 int64_t b_stride, cudaStream_t s)
{
    // the *_simd dispatchers (vector_distance_simd.c:467-509,516-558,571-613) take the SIMD body only when
    // dim >= lanes (and an AVX-512 build skips its AVX2 branch): shorter vectors go to the scalar functions
    if ((arith == NDB_ARITH_AVX2 && dim < 8) || (arith == NDB_ARITH_AVX512 && dim < 16)) arith = NDB_ARITH_OP_F64;
#define NDB_PAIRS_CASE(M, A) \
    if (metric == M && arith == A) return run_pairs<Arith<M, A>>(dA, dB, dOut, n, dim, b_stride, s);
    NDB_PAIRS_CASE(NDB_L2, NDB_ARITH_OP_F64)
    NDB_PAIRS_CASE(NDB_COSINE, NDB_ARITH_OP_F64)
    NDB_PAIRS_CASE(NDB_IP, NDB_ARITH_OP_F64)
    NDB_PAIRS_CASE(NDB_L2, NDB_ARITH_IVF_F32)
    NDB_PAIRS_CASE(NDB_COSINE, NDB_ARITH_IVF_F32)
    NDB_PAIRS_CASE(NDB_IP, NDB_ARITH_IVF_F32)
    NDB_PAIRS_CASE(NDB_L2, NDB_ARITH_HNSW)
    NDB_PAIRS_CASE(NDB_COSINE, NDB_ARITH_HNSW)
    NDB_PAIRS_CASE(NDB_IP, NDB_ARITH_HNSW)
    NDB_PAIRS_CASE(NDB_L2, NDB_ARITH_FAST)
    NDB_PAIRS_CASE(NDB_COSINE, NDB_ARITH_FAST)
    NDB_PAIRS_CASE(NDB_IP, NDB_ARITH_FAST)
#undef NDB_PAIRS_CASE
    if (arith == NDB_ARITH_AVX2 || arith == NDB_ARITH_AVX512) {
        const unsigned blocks = (unsigned) ((n + 127) / 128);
#define NDB_AVX_CASE(M)                                                                                              \
        if (metric == M) {                                                                                           \
            if (arith == NDB_ARITH_AVX2) pairs_avx_kernel<M, 8><<<blocks, 128, 0, s>>>(dA, dB, dOut, n, dim, b_stride); \
            else pairs_avx_kernel<M, 16><<<blocks, 128, 0, s>>>(dA, dB, dOut, n, dim, b_stride);                     \
            count_launch();                                                                                          \
            NDB_CUDA(cudaGetLastError());                                                                            \
            return NDB_B200_OK;                                                                                      \
        }
        NDB_AVX_CASE(NDB_L2)
        NDB_AVX_CASE(NDB_COSINE)
        NDB_AVX_CASE(NDB_IP)
#undef NDB_AVX_CASE
    }
    set_error("unsupported metric/arith combination %d/%d", metric, arith);
    return NDB_B200_EINVAL;
}

// host-pointer driver shared by the pair / row entry points
static int pairs_host(int metric, int arith, bool backend_cos, const float *A, const float *B, float *out,
                      int64_t n, int dim, bool b_is_query, bool validate)
{
    NDB_CHECK(require_init());
    NDB_REQUIRE(A && B && out && n > 0 && dim > 0, NDB_B200_EINVAL, "NULL pointer or non-positive size");
    if (validate) {
        int64_t bad = find_nonfinite(A, n * dim);
        if (bad < 0) bad = find_nonfinite(B, b_is_query ? dim : n * dim);
        NDB_REQUIRE(bad < 0, NDB_B200_EVECTOR, "vector contains NaN or Infinity at index %lld",
                    (long long) (bad % dim));
    }
    Context &c = ctx();
    const size_t abytes = (size_t) n * dim * sizeof(float);
    const size_t bbytes = b_is_query ? (size_t) dim * sizeof(float) : abytes;
    float *dA = nullptr, *dB = nullptr, *dO = nullptr;
    int rc = NDB_B200_OK;
    cudaError_t e;
    if ((e = cudaMalloc(&dA, abytes)) != cudaSuccess || (e = cudaMalloc(&dB, bbytes)) != cudaSuccess ||
        (e = cudaMalloc(&dO, (size_t) n * sizeof(float))) != cudaSuccess) {
        set_error("cudaMalloc failed: %s", cudaGetErrorString(e));
        rc = NDB_B200_ENOMEM;
    }
    if (rc == NDB_B200_OK) {
        e = cudaMemcpyAsync(dA, A, abytes, cudaMemcpyHostToDevice, c.stream);
        if (e == cudaSuccess) e = cudaMemcpyAsync(dB, B, bbytes, cudaMemcpyHostToDevice, c.stream);
        if (e != cudaSuccess) { set_error("H2D failed: %s", cudaGetErrorString(e)); rc = NDB_B200_ECUDA; }
    }
    if (rc == NDB_B200_OK) {
        rc = backend_cos ? run_pairs<CosBackend>(dA, dB, dO, n, dim, b_is_query ? 0 : dim, c.stream)
                         : launch_pairs(metric, arith, dA, dB, dO, n, dim, b_is_query ? 0 : dim, c.stream);
    }
    if (rc == NDB_B200_OK) {
        e = cudaMemcpyAsync(out, dO, (size_t) n * sizeof(float), cudaMemcpyDeviceToHost, c.stream);
        if (e == cudaSuccess) e = cudaStreamSynchronize(c.stream);
        if (e != cudaSuccess) { set_error("D2H failed: %s", cudaGetErrorString(e)); rc = NDB_B200_ECUDA; }
    }
    cudaFree(dA); cudaFree(dB); cudaFree(dO);
    if (rc == NDB_B200_OK && validate && arith == NDB_ARITH_OP_F64 && metric != NDB_IP) {
        // l2_distance / cosine_distance raise on a NaN/Inf result (vector_distance.c:117-120,207-210)
        int64_t bad = find_nonfinite(out, n);
        NDB_REQUIRE(bad < 0, NDB_B200_ERANGE, "distance calculation resulted in NaN or Infinity (pair %lld)",
                    (long long) bad);
    }
    return rc;
}

}  // namespace ndb

using namespace ndb;

extern "C" {

int ndb_b200_abi_version(void) { return NDB_B200_ABI_VERSION; }
const char *ndb_b200_last_error(void) { return g_last_error; }

int ndb_b200_device_count(void)
{
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; }
    return n;
}

int ndb_b200_is_available(void) { return ndb_b200_device_count() > 0 ? 1 : 0; }

int ndb_b200_init(int device)
{
    if (g_ctx.initialized && g_ctx_pid == getpid() && g_ctx.device == device) return NDB_B200_OK;
    if (g_ctx.initialized && g_ctx_pid == getpid()) ndb_b200_shutdown();
    g_ctx = Context();
    g_ctx.generation = ++g_generation;
    int n = ndb_b200_device_count();
    NDB_REQUIRE(n > 0, NDB_B200_ENOTINIT, "no CUDA device visible (this library has no CPU fallback)");
    NDB_REQUIRE(device >= 0 && device < n, NDB_B200_EINVAL, "device %d out of range [0,%d)", device, n);
    NDB_CUDA(cudaSetDevice(device));
    cudaDeviceProp prop;
    NDB_CUDA(cudaGetDeviceProperties(&prop, device));
    NDB_REQUIRE(prop.major == 10, NDB_B200_ENOTINIT,
                "device %d is sm_%d%d; this library carries sm_100a code only", device, prop.major, prop.minor);
    g_ctx.device = device;
    g_ctx.sm_count = prop.multiProcessorCount;
    g_ctx.smem_optin = prop.sharedMemPerBlockOptin;
    NDB_CUDA(cudaStreamCreateWithFlags(&g_ctx.stream, cudaStreamNonBlocking));
    NDB_CUDA(cudaStreamCreateWithFlags(&g_ctx.h2d_stream, cudaStreamNonBlocking));
    NDB_CUDA(cudaStreamCreateWithFlags(&g_ctx.d2h_stream, cudaStreamNonBlocking));
    NDB_CUDA(cudaEventCreate(&g_ctx.ev0));
    NDB_CUDA(cudaEventCreate(&g_ctx.ev1));
    g_ctx.pinned_bytes = 1 << 20;
    NDB_CUDA(cudaMallocHost(&g_ctx.pinned, g_ctx.pinned_bytes));
    NDB_CUDA(cudaMalloc(&g_ctx.d_badidx, 8));
    g_ctx.initialized = true;
    g_ctx_pid = getpid();
    return NDB_B200_OK;
}

void ndb_b200_shutdown(void)
{
    if (!g_ctx.initialized || g_ctx_pid != getpid()) { g_ctx = Context(); return; }
    cudaSetDevice(g_ctx.device);
    cudaStreamSynchronize(g_ctx.stream);
    comm_at_shutdown();
    kmeans_at_shutdown();
    if (g_ctx.pinned) cudaFreeHost(g_ctx.pinned);
    if (g_ctx.d_badidx) cudaFree(g_ctx.d_badidx);
    g_ctx.d_badidx = nullptr;
    if (g_ctx.ev0) cudaEventDestroy(g_ctx.ev0);
    if (g_ctx.ev1) cudaEventDestroy(g_ctx.ev1);
    if (g_ctx.stream) cudaStreamDestroy(g_ctx.stream);
    if (g_ctx.h2d_stream) cudaStreamDestroy(g_ctx.h2d_stream);
    if (g_ctx.d2h_stream) cudaStreamDestroy(g_ctx.d2h_stream);
    g_ctx = Context();
}

int ndb_b200_device_info(int device, char *name, size_t name_len, size_t *total_mem, size_t *free_mem,
                         int *cc_major, int *cc_minor, int *sm_count)
{
    cudaDeviceProp prop;
    NDB_CUDA(cudaGetDeviceProperties(&prop, device));
    if (name && name_len) snprintf(name, name_len, "%s", prop.name);
    if (total_mem) *total_mem = prop.totalGlobalMem;
    if (free_mem) {
        size_t f = 0, t = 0;
        int cur = 0;
        cudaGetDevice(&cur);
        cudaSetDevice(device);
        cudaMemGetInfo(&f, &t);
        cudaSetDevice(cur);
        *free_mem = f;
    }
    if (cc_major) *cc_major = prop.major;
    if (cc_minor) *cc_minor = prop.minor;
    if (sm_count) *sm_count = prop.multiProcessorCount;
    return NDB_B200_OK;
}

int ndb_b200_mem_alloc(void **ptr, size_t bytes)
{
    NDB_CHECK(require_init());
    NDB_REQUIRE(ptr && bytes, NDB_B200_EINVAL, "mem_alloc: NULL or zero size");
    NDB_CUDA(cudaMalloc(ptr, bytes));
    return NDB_B200_OK;
}
int ndb_b200_mem_free(void *ptr)
{
    NDB_CHECK(require_init());
    NDB_CUDA(cudaFree(ptr));
    return NDB_B200_OK;
}
int ndb_b200_memcpy_h2d(void *dst, const void *src, size_t bytes)
{
    NDB_CHECK(require_init());
    NDB_CUDA(cudaMemcpy(dst, src, bytes, cudaMemcpyHostToDevice));
    return NDB_B200_OK;
}
int ndb_b200_memcpy_d2h(void *dst, const void *src, size_t bytes)
{
    NDB_CHECK(require_init());
    NDB_CUDA(cudaMemcpy(dst, src, bytes, cudaMemcpyDeviceToHost));
    return NDB_B200_OK;
}
int ndb_b200_host_alloc_pinned(void **ptr, size_t bytes)
{
    NDB_CHECK(require_init());
    NDB_CUDA(cudaMallocHost(ptr, bytes));
    return NDB_B200_OK;
}
int ndb_b200_host_free_pinned(void *ptr)
{
    NDB_CHECK(require_init());
    NDB_CUDA(cudaFreeHost(ptr));
    return NDB_B200_OK;
}
int ndb_b200_stream_create(void **stream)
{
    NDB_CHECK(require_init());
    cudaStream_t s;
    NDB_CUDA(cudaStreamCreateWithFlags(&s, cudaStreamNonBlocking));
    *stream = (void *) s;
    return NDB_B200_OK;
}
int ndb_b200_stream_destroy(void *stream)
{
    NDB_CHECK(require_init());
    NDB_CUDA(cudaStreamDestroy((cudaStream_t) stream));
    return NDB_B200_OK;
}
int ndb_b200_stream_synchronize(void *stream)
{
    NDB_CHECK(require_init());
    NDB_CUDA(cudaStreamSynchronize(stream ? (cudaStream_t) stream : ctx().stream));
    return NDB_B200_OK;
}

int64_t ndb_b200_launch_count(void) { return ctx().launches; }
int ndb_b200_set_timing(int enabled) { ctx().timing = enabled != 0; return NDB_B200_OK; }
int ndb_b200_last_kernel_stats(double *ms, double *algo_bytes, int64_t *evals)
{
    Context &c = ctx();
    if (c.initialized && c.last_ms < 0.0) {
        float t = 0.0f;
        NDB_CUDA(cudaEventSynchronize(c.ev1));
        NDB_CUDA(cudaEventElapsedTime(&t, c.ev0, c.ev1));
        c.last_ms = t;
    }
    if (c.initialized && c.last_bytes < 0.0 && c.stats_src) {
        unsigned long long scanned = 0;
        NDB_CUDA(cudaMemcpy(&scanned, c.stats_src, sizeof(scanned), cudaMemcpyDeviceToHost));
        // sum over (query, probed list) of len * d * sizeof(float) + len * 8 (ids)
        c.last_bytes = (double) scanned * (c.stats_dim * 4.0 + 8.0);
        c.last_evals = (int64_t) scanned;
    }
    if (ms) *ms = ctx().last_ms;
    if (algo_bytes) *algo_bytes = ctx().last_bytes;
    if (evals) *evals = ctx().last_evals;
    return NDB_B200_OK;
}

// ---- operator backends -------------------------------------------------------------------
int ndb_b200_distance_pairs(int metric, int arith, const float *A, const float *B, float *out, int64_t n, int dim)
{
    return pairs_host(metric, arith, false, A, B, out, n, dim, false, true);
}
int ndb_b200_distance_rows(int metric, int arith, const float *X, int64_t n, int dim, const float *q, float *out)
{
    return pairs_host(metric, arith, false, X, q, out, n, dim, true, true);
}

// vector_l2_distance_batch / vector_cosine_distance_batch / vector_inner_product_distance_batch
// (src/vector/vector_batch.c:37-420): an array of vectors against one query, in the operators' fp64 arithmetic
// (the functions call l2_distance / cosine_distance / -inner_product_distance, :152,278,404).  An element that is
// NULL, or whose dimension is invalid or differs from the query's, yields NULL (:123-150); an empty array and an
// invalid query dimension are errors (:85-88,69-73).
int ndb_b200_vector_distance_batch(int metric, const float *rows, const int *dims, int64_t n, int dim, const float *query,
                                   float *out, uint8_t *nulls)
{
    NDB_CHECK(require_init());
    NDB_REQUIRE(rows && dims && query && out && nulls, NDB_B200_EINVAL, "vector array and query vector must not be NULL");
    NDB_REQUIRE(n > 0, NDB_B200_EINVAL, "vector array must not be empty");
    NDB_REQUIRE(dim > 0 && dim <= 16000, NDB_B200_EINVAL, "invalid query vector dimension: %d", dim);
    NDB_REQUIRE(metric >= NDB_L2 && metric <= NDB_IP, NDB_B200_EINVAL, "distance_batch: unknown metric %d", metric);
    // the valid elements are packed, evaluated in one launch and scattered back
    std::vector<int64_t> valid;
    valid.reserve((size_t) n);
    for (int64_t i = 0; i < n; i++) {
        nulls[i] = dims[i] == dim ? 0 : 1;
        out[i] = 0.0f;
        if (!nulls[i]) valid.push_back(i);
    }
    if (valid.empty()) return NDB_B200_OK;
    if ((int64_t) valid.size() == n) return pairs_host(metric, NDB_ARITH_OP_F64, false, rows, query, out, n, dim, true, true);
    std::vector<float> packed(valid.size() * (size_t) dim), res(valid.size());
    for (size_t j = 0; j < valid.size(); j++) memcpy(&packed[j * dim], rows + (size_t) valid[j] * dim, (size_t) dim * 4);
    NDB_CHECK(pairs_host(metric, NDB_ARITH_OP_F64, false, packed.data(), query, res.data(), (int64_t) valid.size(), dim, true, true));
    for (size_t j = 0; j < valid.size(); j++) out[valid[j]] = res[j];
    return NDB_B200_OK;
}

// ---- ndb_gpu_backend launchers (stream == NULL => synchronous, out valid on return) ----------
int ndb_b200_launch_l2_distance(const float *A, const float *B, float *out, int n, int d, void *stream)
{
    (void) stream;     // host pointers: the result must be in `out` on return either way
    // ||A_i - B_i||_2 in fp32 (cublasSnrm2 of the difference, gpu_backend_cuda.c:430-437)
    return pairs_host(NDB_L2, NDB_ARITH_IVF_F32, false, A, B, out, n, d, false, false);
}
int ndb_b200_launch_cosine(const float *A, const float *B, float *out, int n, int d, void *stream)
{
    (void) stream;
    return pairs_host(NDB_COSINE, NDB_ARITH_FAST, true, A, B, out, n, d, false, false);
}

}  // extern "C"
