// quant.cu -- the per-vector quantisers of NeuronDB/src/types/quantization.c for whole row sets, and the Hamming scan
// over binary rows (quantised list formats).
//
//   kind                reference                              output row (the varlena's data[] bytes)
//   NDB_QUANT_INT8      quantize_vector_i8      :42-86         dim   int8   rintf(x * 127 / max|x|), clamped; zero row -> zeros
//   NDB_QUANT_FP16      quantize_vector_f16     :220-236       2 dim        float4_to_fp16 :141-168: mantissa TRUNCATED, subnormal
//                                                                           results flushed to zero, NaN / overflow -> inf
//   NDB_QUANT_BINARY    quantize_vector_binary  :284-312       (dim+7)/8    bit i%8 of byte i/8 = x > 0
//   NDB_QUANT_UINT8     quantize_vector_uint8   :1354-1402     dim   uint8  rintf((x - min) * 255 / (max - min)); constant row -> zeros
//   NDB_QUANT_TERNARY   quantize_vector_ternary :1455-1503     (2dim+7)/8   2 bits: 2 if x > max|x|/3, 1 if x < -max|x|/3, else 0
//   NDB_QUANT_INT4      quantize_vector_int4    :1562-1641     (dim+1)/2    nibble 8 + clamp(rintf(x * 7 / max|x|)), low nibble first;
//                                                                           zero row -> zero bytes (not nibble 8)
//   binary_hamming_distance :385-427: popcount of the XOR over (dim+7)/8 bytes
//
// All of it is byte work bound by HBM: one pass for the row statistics (a warp per row), one pass writing one output
// byte (or half) per thread -- 2 x 4 dim bytes read (the second mostly from L2), the row's bytes written.  f32 operations
// are the explicit round-to-nearest intrinsics, rintf = the default rounding mode's round-half-even on both sides.
// The Hamming scan keeps a block's queries in shared memory and reads each row's words once per query block; its top-k
// is the scan kernels' (distance as an exactly representable float, row as key).
#include <cstdlib>
#include "layout.cuh"
#include "arith.cuh"
#include "block_topk.cuh"

namespace ndb {

struct RowStats { float max_abs, mn, mx; };

__global__ void __launch_bounds__(256) quant_stats_kernel(const float *__restrict__ X, int64_t n, int dim, RowStats *__restrict__ st)
{
    const int lane = threadIdx.x & 31;
    const int64_t row = (int64_t) blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (row >= n) return;
    const float *v = X + (size_t) row * dim;
    float ma = 0.0f, mn = v[0], mx = v[0];
    for (int i = lane; i < dim; i += 32) {
        const float x = v[i];
        ma = fmaxf(ma, fabsf(x));
        mn = fminf(mn, x);
        mx = fmaxf(mx, x);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        ma = fmaxf(ma, __shfl_xor_sync(FULL, ma, o));
        mn = fminf(mn, __shfl_xor_sync(FULL, mn, o));
        mx = fmaxf(mx, __shfl_xor_sync(FULL, mx, o));
    }
    if (lane == 0) { st[row].max_abs = ma; st[row].mn = mn; st[row].mx = mx; }
}

__device__ __forceinline__ uint16_t float_to_fp16_trunc(float f)            // float4_to_fp16 :141-168
{
    const uint32_t u = __float_as_uint(f);
    const uint16_t sign = (u >> 16) & 0x8000;
    const uint32_t mantissa = u & 0x7fffff;
    const int exp = (int) ((u >> 23) & 0xff) - 127 + 15;
    if (exp <= 0) return sign;
    if (exp >= 31) return sign | 0x7c00;
    return (uint16_t) (sign | (exp << 10) | (mantissa >> 13));
}

// one thread per output byte (per half for FP16)
template <int KIND>
__global__ void __launch_bounds__(256) quant_rows_kernel(const float *__restrict__ X, int64_t n, int dim, int row_units,
                                                         const RowStats *__restrict__ st, uint8_t *__restrict__ out)
{
    const int64_t t = (int64_t) blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n * row_units) return;
    const int64_t row = t / row_units;
    const int u = (int) (t % row_units);
    const float *v = X + (size_t) row * dim;
    if (KIND == NDB_QUANT_FP16) {
        reinterpret_cast<uint16_t *>(out)[t] = float_to_fp16_trunc(v[u]);
    } else if (KIND == NDB_QUANT_INT8) {
        const float max_abs = st[row].max_abs;
        int8_t r = 0;
        if (max_abs != 0.0f) {
            float val = __fmul_rn(v[u], __fdiv_rn(127.0f, max_abs));
            if (val > 127.0f) val = 127.0f;
            if (val < -128.0f) val = -128.0f;
            r = (int8_t) rintf(val);
        }
        out[t] = (uint8_t) r;
    } else if (KIND == NDB_QUANT_UINT8) {
        const float mn = st[row].mn, mx = st[row].mx;
        uint8_t r = 0;
        if (mx != mn) {
            float nv = __fmul_rn(__fsub_rn(v[u], mn), __fdiv_rn(255.0f, __fsub_rn(mx, mn)));
            if (nv > 255.0f) nv = 255.0f;
            if (nv < 0.0f) nv = 0.0f;
            r = (uint8_t) rintf(nv);
        }
        out[t] = r;
    } else if (KIND == NDB_QUANT_BINARY) {
        uint8_t b = 0;
#pragma unroll
        for (int j = 0; j < 8; j++) {
            const int i = u * 8 + j;
            if (i < dim && v[i] > 0.0f) b |= (uint8_t) (1u << j);
        }
        out[t] = b;
    } else if (KIND == NDB_QUANT_TERNARY) {
        const float threshold = __fdiv_rn(st[row].max_abs, 3.0f);
        uint8_t b = 0;
#pragma unroll
        for (int j = 0; j < 4; j++) {
            const int i = u * 4 + j;
            if (i < dim) {
                const float x = v[i];
                const uint8_t value = x > threshold ? 2 : (x < -threshold ? 1 : 0);
                b |= (uint8_t) (value << (2 * j));
            }
        }
        out[t] = b;
    } else if (KIND == NDB_QUANT_INT4) {
        const float max_abs = st[row].max_abs;
        uint8_t b = 0;
        if (max_abs != 0.0f) {
            const float scale = __fdiv_rn(7.0f, max_abs);
#pragma unroll
            for (int j = 0; j < 2; j++) {
                const int i = u * 2 + j;
                if (i < dim) {
                    const float scaled = __fmul_rn(v[i], scale);
                    int value;
                    if (scaled > 7.0f) value = 7;
                    else if (scaled < -8.0f) value = -8;
                    else value = (int) rintf(scaled);
                    int uvalue = 8 + value;
                    if (uvalue > 15) uvalue = 15;
                    b |= (uint8_t) (uvalue << (4 * j));
                }
            }
        }
        out[t] = b;
    }
}

// ---- Hamming scan: grid (nparts, ceil(nq / QB)); a block keeps QB queries in shared memory, a lane owns a row ---------
constexpr int HAM_QB = 8;
template <int KR>
__global__ void __launch_bounds__(256) hamming_topk_kernel(const uint8_t *__restrict__ rows, int64_t n, int nbytes, const uint8_t *__restrict__ Q,
                                                           int nq, int k, int64_t rows_per_part, float *__restrict__ pdist,
                                                           uint32_t *__restrict__ pslot)
{
    extern __shared__ __align__(16) unsigned char hsm[];          // QB queries of nbytes (padded to 4), later the merge
    const int nwords = (nbytes + 3) >> 2;
    uint32_t *qs = reinterpret_cast<uint32_t *>(hsm);
    const int q0 = blockIdx.y * HAM_QB, part = blockIdx.x;
    for (int e = threadIdx.x; e < HAM_QB * nwords * 4; e += blockDim.x) {
        const int qq = e / (nwords * 4), b = e % (nwords * 4);
        hsm[e] = (q0 + qq < nq && b < nbytes) ? Q[(size_t) (q0 + qq) * nbytes + b] : (unsigned char) 0;
    }
    __syncthreads();
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
    const int64_t r0 = (int64_t) part * rows_per_part;
    const int64_t r1 = r0 + rows_per_part < n ? r0 + rows_per_part : n;
    const bool words = (nbytes & 3) == 0;
    // warp w serves query q0 + w (HAM_QB == warps per block): every warp walks the whole row range for its query, the
    // rows' bytes come from L1 / L2 after the first warp touched them
    WarpTopK<KR, uint32_t> top;
    top.init();
    const uint32_t *myq = qs + warp * nwords;
    const bool live = q0 + warp < nq;
    for (int64_t base = r0; base < r1; base += 32) {
        const int64_t row = base + lane;
        const bool valid = live && row < r1;
        int cnt = 0;
        if (valid) {
            if (words) {
                const uint32_t *rw = reinterpret_cast<const uint32_t *>(rows + (size_t) row * nbytes);
                for (int w = 0; w < nwords; w++) cnt += __popc(rw[w] ^ myq[w]);
            } else {
                const uint8_t *rb = rows + (size_t) row * nbytes;
                const uint8_t *qb = reinterpret_cast<const uint8_t *>(myq);
                for (int b = 0; b < nbytes; b++) cnt += __popc((uint32_t) (rb[b] ^ qb[b]));
            }
        }
        top.offer((float) cnt, (uint32_t) row, valid, lane, k);
    }
    if (!live) return;
    const size_t ob = ((size_t) (q0 + warp) * gridDim.x + part) * k;
#pragma unroll
    for (int r = 0; r < KR; r++) {
        const int e = r * 32 + lane;
        if (e < k) { pdist[ob + e] = top.d[r]; pslot[ob + e] = top.key[r]; }
    }
    (void) nwarps;
}

// k <= 32: a lane loads its row ONCE and scores it against all HAM_QB queries of the block (their words broadcast from
// shared memory), one WarpTopK list per query in registers; the warps split the block's rows and their lists are merged
// per query at the end.  L2 -> SM traffic is the row bytes once per block of 8 queries instead of once per query.
__global__ void __launch_bounds__(256) hamming_topk_mq_kernel(const uint8_t *__restrict__ rows, int64_t n, int nbytes, const uint8_t *__restrict__ Q,
                                                              int nq, int k, int64_t rows_per_part, float *__restrict__ pdist,
                                                              uint32_t *__restrict__ pslot)
{
    extern __shared__ __align__(16) unsigned char hsm[];          // HAM_QB queries of nwords words, then the merge area (2 KB)
    const int nwords = (nbytes + 3) >> 2;
    const uint32_t *qs = reinterpret_cast<const uint32_t *>(hsm);
    void *merge = hsm + (((size_t) HAM_QB * nwords * 4 + 15) & ~(size_t) 15);
    const int q0 = blockIdx.y * HAM_QB, part = blockIdx.x;
    for (int e = threadIdx.x; e < HAM_QB * nwords * 4; e += blockDim.x) {
        const int qq = e / (nwords * 4), b = e % (nwords * 4);
        hsm[e] = (q0 + qq < nq && b < nbytes) ? Q[(size_t) (q0 + qq) * nbytes + b] : (unsigned char) 0;
    }
    __syncthreads();
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
    const int64_t r0 = (int64_t) part * rows_per_part;
    const int64_t r1 = r0 + rows_per_part < n ? r0 + rows_per_part : n;
    const bool words = (nbytes & 3) == 0;
    WarpTopK<1, uint32_t> top[HAM_QB];
#pragma unroll
    for (int j = 0; j < HAM_QB; j++) top[j].init();
    for (int64_t base = r0 + (int64_t) warp * 32; base < r1; base += (int64_t) nwarps * 32) {
        const int64_t row = base + lane;
        const bool valid = row < r1;
        int cnt[HAM_QB];
#pragma unroll
        for (int j = 0; j < HAM_QB; j++) cnt[j] = 0;
        if (valid) {
            if (words) {
                const uint32_t *rw = reinterpret_cast<const uint32_t *>(rows + (size_t) row * nbytes);
                for (int w = 0; w < nwords; w++) {
                    const uint32_t x = rw[w];
#pragma unroll
                    for (int j = 0; j < HAM_QB; j++) cnt[j] += __popc(x ^ qs[j * nwords + w]);
                }
            } else {
                const uint8_t *rb = rows + (size_t) row * nbytes;
                const uint8_t *qb = reinterpret_cast<const uint8_t *>(qs);
                for (int b = 0; b < nbytes; b++) {
                    const uint32_t x = rb[b];
#pragma unroll
                    for (int j = 0; j < HAM_QB; j++) cnt[j] += __popc(x ^ (uint32_t) qb[j * nwords * 4 + b]);
                }
            }
        }
#pragma unroll
        for (int j = 0; j < HAM_QB; j++) top[j].offer((float) cnt[j], (uint32_t) row, valid, lane, k);
    }
#pragma unroll
    for (int j = 0; j < HAM_QB; j++) {
        __syncthreads();
        if (q0 + j < nq) block_topk_write<1>(top[j], merge, warp, lane, nwarps, k, pdist, pslot, ((size_t) (q0 + j) * gridDim.x + part) * k);
    }
}

__global__ void hamming_finish_kernel(const float *__restrict__ d, const int64_t *__restrict__ ids, int64_t total, int32_t *__restrict__ out)
{
    const int64_t i = (int64_t) blockIdx.x * blockDim.x + threadIdx.x;
    if (i < total) out[i] = ids[i] < 0 ? -1 : (int32_t) d[i];
}

}  // namespace ndb

using namespace ndb;

extern "C" {

int64_t ndb_b200_quantized_row_bytes(int kind, int dim)
{
    if (dim <= 0) return -1;
    switch (kind) {
    case NDB_QUANT_INT8: case NDB_QUANT_UINT8: return dim;
    case NDB_QUANT_FP16: return 2 * (int64_t) dim;
    case NDB_QUANT_BINARY: return (dim + 7) / 8;
    case NDB_QUANT_TERNARY: return ((int64_t) dim * 2 + 7) / 8;
    case NDB_QUANT_INT4: return (dim + 1) / 2;
    }
    return -1;
}

int ndb_b200_quantize_rows(int kind, const float *X, int64_t n, int dim, void *out)
{
    NDB_CHECK(require_init());
    NDB_REQUIRE(X && out && n > 0, NDB_B200_EINVAL, "quantize_rows: NULL or empty argument");
    NDB_REQUIRE(dim >= 1 && dim <= 32767, NDB_B200_EINVAL, "quantize_rows: dim %d out of range 1..32767", dim);      // int16 dim of the varlena
    const int64_t rb = ndb_b200_quantized_row_bytes(kind, dim);
    NDB_REQUIRE(rb > 0, NDB_B200_EINVAL, "quantize_rows: unknown kind %d", kind);
    cudaStream_t s = ctx().stream;
    DevBuf dX, dst, dout;
    const size_t count = (size_t) n * dim;
    NDB_CHECK(dX.reserve(count * 4)); NDB_CHECK(dst.reserve((size_t) n * sizeof(RowStats))); NDB_CHECK(dout.reserve((size_t) n * rb));
    NDB_CUDA(cudaMemcpyAsync(dX.p, X, count * 4, cudaMemcpyHostToDevice, s));
    NDB_CHECK(validate_begin(dX.as<float>(), (int64_t) count, s));
    if (kind != NDB_QUANT_FP16 && kind != NDB_QUANT_BINARY) {
        quant_stats_kernel<<<(unsigned) ((n + 7) / 8), 256, 0, s>>>(dX.as<float>(), n, dim, dst.as<RowStats>());
        count_launch();
    }
    const int units = (int) (kind == NDB_QUANT_FP16 ? dim : rb);
    const unsigned grid = (unsigned) ((n * units + 255) / 256);
#define NDB_Q_LAUNCH(K) quant_rows_kernel<K><<<grid, 256, 0, s>>>(dX.as<float>(), n, dim, units, dst.as<RowStats>(), dout.as<uint8_t>())
    switch (kind) {
    case NDB_QUANT_INT8: NDB_Q_LAUNCH(NDB_QUANT_INT8); break;
    case NDB_QUANT_FP16: NDB_Q_LAUNCH(NDB_QUANT_FP16); break;
    case NDB_QUANT_BINARY: NDB_Q_LAUNCH(NDB_QUANT_BINARY); break;
    case NDB_QUANT_UINT8: NDB_Q_LAUNCH(NDB_QUANT_UINT8); break;
    case NDB_QUANT_TERNARY: NDB_Q_LAUNCH(NDB_QUANT_TERNARY); break;
    default: NDB_Q_LAUNCH(NDB_QUANT_INT4); break;
    }
#undef NDB_Q_LAUNCH
    count_launch();
    NDB_CUDA(cudaGetLastError());
    NDB_CUDA(cudaMemcpyAsync(out, dout.p, (size_t) n * rb, cudaMemcpyDeviceToHost, s));
    NDB_CUDA(cudaStreamSynchronize(s));
    // the reference's vector type cannot hold NaN / Inf (vector_in rejects them); uint8's min / max scan has no defined answer for them
    NDB_REQUIRE(validate_end() < 0, NDB_B200_EVECTOR, "quantize_rows: NaN/Inf in the vectors");
    return NDB_B200_OK;
}

int ndb_b200_hamming_knn(const uint8_t *rows, int64_t n, int nbits, const uint8_t *Q, int nq, int k, int32_t *dist, int64_t *ids)
{
    NDB_CHECK(require_init());
    NDB_REQUIRE(rows && Q && dist && ids && n > 0 && nq > 0, NDB_B200_EINVAL, "hamming_knn: NULL or empty argument");
    NDB_REQUIRE(nbits >= 1 && nbits <= 32767, NDB_B200_EINVAL, "hamming_knn: nbits %d out of range 1..32767", nbits);
    NDB_REQUIRE(k >= 1 && k <= 128, NDB_B200_EINVAL, "hamming_knn: k=%d out of range 1..128", k);
    NDB_REQUIRE(n < (int64_t) 0xfffffff0ll, NDB_B200_EINVAL, "hamming_knn: too many rows for 32-bit slots");
    NDB_REQUIRE(nq <= 65535 * HAM_QB, NDB_B200_EINVAL, "hamming_knn: at most %d queries per call (got %d); split the batch", 65535 * HAM_QB, nq);
    const int nbytes = (nbits + 7) / 8, nwords = (nbytes + 3) / 4;
    cudaStream_t s = ctx().stream;
    DevBuf drows, dq, pdist, pslot, od, oi, o32;
    NDB_CHECK(drows.reserve((size_t) n * nbytes + 4)); NDB_CHECK(dq.reserve((size_t) nq * nbytes));
    NDB_CUDA(cudaMemcpyAsync(drows.p, rows, (size_t) n * nbytes, cudaMemcpyHostToDevice, s));
    NDB_CUDA(cudaMemcpyAsync(dq.p, Q, (size_t) nq * nbytes, cudaMemcpyHostToDevice, s));
    int64_t rows_per_part = 1 << 20;
    const int qblocks = (nq + HAM_QB - 1) / HAM_QB;
    while (rows_per_part > 8192 && (n + rows_per_part - 1) / rows_per_part * qblocks < 2 * ctx().sm_count) rows_per_part >>= 1;
    const int nparts = (int) ((n + rows_per_part - 1) / rows_per_part);
    const size_t m = (size_t) nq * k;
    NDB_CHECK(pdist.reserve(m * nparts * 4)); NDB_CHECK(pslot.reserve(m * nparts * 4));
    NDB_CHECK(od.reserve(m * 4)); NDB_CHECK(oi.reserve(m * 8)); NDB_CHECK(o32.reserve(m * 4));
    const size_t smem = (size_t) HAM_QB * nwords * 4;
    const size_t smem_mq = ((smem + 15) & ~(size_t) 15) + 8 * 32 * 8;
    dim3 grid((unsigned) nparts, (unsigned) qblocks);
    if (k <= 32 && !getenv("NDB_HAMMING_PER_QUERY"))      // (switch: the warp-per-query kernel, for comparison)
        hamming_topk_mq_kernel<<<grid, 256, smem_mq, s>>>(drows.as<uint8_t>(), n, nbytes, dq.as<uint8_t>(), nq, k, rows_per_part, pdist.as<float>(), pslot.as<uint32_t>());
    else if (k <= 32) hamming_topk_kernel<1><<<grid, 256, smem, s>>>(drows.as<uint8_t>(), n, nbytes, dq.as<uint8_t>(), nq, k, rows_per_part, pdist.as<float>(), pslot.as<uint32_t>());
    else hamming_topk_kernel<4><<<grid, 256, smem, s>>>(drows.as<uint8_t>(), n, nbytes, dq.as<uint8_t>(), nq, k, rows_per_part, pdist.as<float>(), pslot.as<uint32_t>());
    count_launch();
    NDB_CUDA(cudaGetLastError());
    NDB_CHECK(launch_merge_parts(pdist.as<float>(), pslot.as<uint32_t>(), nullptr, nq, nparts, k, od.as<float>(), oi.as<int64_t>(), nullptr, s));
    hamming_finish_kernel<<<(unsigned) ((m + 255) / 256), 256, 0, s>>>(od.as<float>(), oi.as<int64_t>(), (int64_t) m, o32.as<int32_t>());
    count_launch();
    NDB_CUDA(cudaMemcpyAsync(dist, o32.p, m * 4, cudaMemcpyDeviceToHost, s));
    NDB_CUDA(cudaMemcpyAsync(ids, oi.p, m * 8, cudaMemcpyDeviceToHost, s));
    NDB_CUDA(cudaStreamSynchronize(s));
    return NDB_B200_OK;
}

}  // extern "C"
