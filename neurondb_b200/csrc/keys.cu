// keys.cu -- index key extraction: the access methods turn whatever the indexed
// column holds into float4[dim] before anything else happens (ivfExtractVectorData, ivf_am.c:117-218;
// hnswExtractVectorData, hnsw_am.c:1402-1519):
//   vector     -> copied out of the varlena (struct Vector, neurondb.h:35-41: int32 vl_len_, int16 dim,
//                 int16 unused, float4 data[dim]); a datum whose dim differs is an error
//   halfvec    -> fp16_to_float per element (src/types/quantization.c:171-215; IEEE except that
//                 subnormal halves come out 2^-10 too small -- the reference's result is kept)
//   sparsevec  -> zero-filled row, result[indices[i]] = values[i] in entry order, out-of-range
//                 indices ignored (so a repeated index keeps the LAST value)
//   bit        -> bit i (MSB first within each byte) ? 1.0f : -1.0f
// Batched here: n keys in, n rows of `dim` floats out, ready for dataset_append / ivf_insert /
// hnsw_build.  Pure byte shuffling: one pass, bound by HBM.
#include "common.cuh"


namespace ndb {

__global__ void keys_halfvec_kernel(const uint16_t *__restrict__ h, int64_t total, float *__restrict__ out)
{
    for (int64_t i = (int64_t) blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t) gridDim.x * blockDim.x)
    {
        // fp16_to_float (quantization.c:171-215) in integer arithmetic, so that NaN payloads come out as
        // the reference leaves them (a hardware conversion would quiet signalling NaNs).  IEEE for
        // zeros, normals, Inf and NaN; SUBNORMAL halves are renormalised with an exponent that is 10
        // too small (:186-197) -- they come out 2^-10 times their IEEE value, and that is kept.
        const uint32_t b = h[i];
        const uint32_t sign = (b & 0x8000u) << 16, e = (b & 0x7c00u) >> 10, man = b & 0x03ffu;
        uint32_t f;
        if (e == 0) {
            if (man == 0) f = sign;
            else {
                const int shifts = __clz((int) man) - 21;             // bring the leading one up to bit 10
                f = sign | ((uint32_t) (103 - shifts) << 23) | (((man << shifts) & 0x03ffu) << 13);
            }
        } else if (e == 0x1f) f = sign | 0x7f800000u | (man << 13);
        else f = sign | ((e + 112u) << 23) | (man << 13);
        out[i] = __uint_as_float(f);
    }
}

__global__ void keys_bits_kernel(const uint8_t *__restrict__ bits, int64_t n, int nbits, int row_bytes, float *__restrict__ out)
{
    const int64_t total = n * nbits;
    for (int64_t t = (int64_t) blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (int64_t) gridDim.x * blockDim.x) {
        const int64_t r = t / nbits;
        const int i = (int) (t - r * nbits);
        const int v = (bits[r * row_bytes + (i >> 3)] >> (7 - (i & 7))) & 1;
        out[t] = v ? 1.0f : -1.0f;
    }
}

// one thread per row: the entries are applied in order, which is what makes "last one wins"
__global__ void keys_sparse_kernel(const int64_t *__restrict__ indptr, const int32_t *__restrict__ indices,
                                   const float *__restrict__ values, int64_t n, int total_dim, float *__restrict__ out)
{
    const int64_t r = (int64_t) blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= n) return;
    float *row = out + r * total_dim;
    for (int64_t e = indptr[r]; e < indptr[r + 1]; e++) {
        const int32_t j = indices[e];
        if (j >= 0 && j < total_dim) row[j] = values[e];
    }
}

// detoasted Vector datums laid end to end, each 8 + 4 * dim bytes; bad[0] = first datum whose header
// disagrees with dim (or ~0)
__global__ void keys_vector_kernel(const unsigned char *__restrict__ datums, int64_t n, int dim, float *__restrict__ out,
                                   unsigned long long *__restrict__ bad)
{
    const int64_t total = n * dim;
    const size_t stride = 8 + (size_t) dim * 4;
    for (int64_t t = (int64_t) blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (int64_t) gridDim.x * blockDim.x) {
        const int64_t r = t / dim;
        const int j = (int) (t - r * dim);
        const unsigned char *d = datums + (size_t) r * stride;
        if (j == 0 && *reinterpret_cast<const int16_t *>(d + 4) != (int16_t) dim) atomicMin(bad, (unsigned long long) r);
        out[t] = *reinterpret_cast<const float *>(d + 8 + (size_t) j * 4);
    }
}

static unsigned grid_for(int64_t total)
{
    const int64_t b = (total + 255) / 256, cap = (int64_t) ctx().sm_count * 16;
    return (unsigned) (b < cap ? (b > 0 ? b : 1) : cap);
}

// stage `bytes` of host input on the device, run `launch(dev_in, dev_out)`, copy n*dim floats back
template <class F> static int staged(const void *in, size_t in_bytes, float *rows_out, size_t out_floats, F launch)
{
    cudaStream_t s = ctx().stream;
    DevBuf din, dout;
    NDB_CHECK(din.reserve(in_bytes ? in_bytes : 16));
    NDB_CHECK(dout.reserve(out_floats * 4 ? out_floats * 4 : 16));
    NDB_CUDA(cudaMemcpyAsync(din.p, in, in_bytes, cudaMemcpyHostToDevice, s));
    int rc = launch(din.p, dout.as<float>(), s);
    if (rc == NDB_B200_OK) {
        cudaError_t e = cudaMemcpyAsync(rows_out, dout.p, out_floats * 4, cudaMemcpyDeviceToHost, s);
        if (e == cudaSuccess) e = cudaStreamSynchronize(s);
        if (e != cudaSuccess) { set_error("keys: copy back failed: %s", cudaGetErrorString(e)); rc = NDB_B200_ECUDA; }
    }
    din.release();
    dout.release();
    return rc;
}

}  // namespace ndb

using namespace ndb;

extern "C" {

int ndb_b200_keys_from_vector(const void *datums, int64_t n, int dim, float *rows)
{
    NDB_CHECK(require_init());
    NDB_REQUIRE(datums && rows && n >= 0, NDB_B200_EINVAL, "keys_from_vector: bad argument");
    NDB_REQUIRE(dim > 0 && dim <= 32767, NDB_B200_EINVAL, "invalid vector dimension %d", dim);
    if (n == 0) return NDB_B200_OK;
    unsigned long long first_bad = ~0ull;
    int rc = staged(datums, (size_t) n * (8 + (size_t) dim * 4), rows, (size_t) n * dim, [&](void *din, float *dout, cudaStream_t s) {
        Context &c = ctx();
        NDB_CUDA(cudaMemsetAsync(c.d_badidx, 0xFF, 8, s));
        keys_vector_kernel<<<grid_for(n * dim), 256, 0, s>>>(static_cast<const unsigned char *>(din), n, dim, dout, c.d_badidx);
        count_launch();
        NDB_CUDA(cudaGetLastError());
        NDB_CUDA(cudaMemcpyAsync(&first_bad, c.d_badidx, 8, cudaMemcpyDeviceToHost, s));
        return (int) NDB_B200_OK;
    });
    NDB_CHECK(rc);
    NDB_REQUIRE(first_bad == ~0ull, NDB_B200_EDIM, "vector dimensions must match: datum %llu is not of dimension %d", first_bad, dim);
    return NDB_B200_OK;
}

int ndb_b200_keys_from_halfvec_dev(const uint16_t *h_dev, int64_t n, int dim, float *rows_dev, void *stream)
{
    NDB_CHECK(require_init());
    NDB_REQUIRE(h_dev && rows_dev && n >= 0, NDB_B200_EINVAL, "keys_from_halfvec: bad argument");
    NDB_REQUIRE(dim > 0 && dim <= 32767, NDB_B200_EINVAL, "invalid halfvec dimension %d", dim);
    if (n == 0) return NDB_B200_OK;
    cudaStream_t s = stream ? (cudaStream_t) stream : ctx().stream;
    keys_halfvec_kernel<<<grid_for(n * dim), 256, 0, s>>>(h_dev, n * dim, rows_dev);
    count_launch();
    NDB_CUDA(cudaGetLastError());
    return NDB_B200_OK;
}

int ndb_b200_keys_from_halfvec(const uint16_t *h, int64_t n, int dim, float *rows)
{
    NDB_CHECK(require_init());
    NDB_REQUIRE(h && rows && n >= 0, NDB_B200_EINVAL, "keys_from_halfvec: bad argument");
    NDB_REQUIRE(dim > 0 && dim <= 32767, NDB_B200_EINVAL, "invalid halfvec dimension %d", dim);
    if (n == 0) return NDB_B200_OK;
    return staged(h, (size_t) n * dim * 2, rows, (size_t) n * dim, [&](void *din, float *dout, cudaStream_t s) {
        return ndb_b200_keys_from_halfvec_dev(static_cast<const uint16_t *>(din), n, dim, dout, s);
    });
}

int ndb_b200_keys_from_bits_dev(const uint8_t *bits_dev, int64_t n, int nbits, float *rows_dev, void *stream)
{
    NDB_CHECK(require_init());
    NDB_REQUIRE(bits_dev && rows_dev && n >= 0, NDB_B200_EINVAL, "keys_from_bits: bad argument");
    NDB_REQUIRE(nbits > 0 && nbits <= 32767, NDB_B200_EINVAL, "invalid bit vector length %d", nbits);
    if (n == 0) return NDB_B200_OK;
    cudaStream_t s = stream ? (cudaStream_t) stream : ctx().stream;
    keys_bits_kernel<<<grid_for(n * nbits), 256, 0, s>>>(bits_dev, n, nbits, (nbits + 7) / 8, rows_dev);
    count_launch();
    NDB_CUDA(cudaGetLastError());
    return NDB_B200_OK;
}

int ndb_b200_keys_from_bits(const uint8_t *bits, int64_t n, int nbits, float *rows)
{
    NDB_CHECK(require_init());
    NDB_REQUIRE(bits && rows && n >= 0, NDB_B200_EINVAL, "keys_from_bits: bad argument");
    NDB_REQUIRE(nbits > 0 && nbits <= 32767, NDB_B200_EINVAL, "invalid bit vector length %d", nbits);
    if (n == 0) return NDB_B200_OK;
    return staged(bits, (size_t) n * ((nbits + 7) / 8), rows, (size_t) n * nbits, [&](void *din, float *dout, cudaStream_t s) {
        return ndb_b200_keys_from_bits_dev(static_cast<const uint8_t *>(din), n, nbits, dout, s);
    });
}

int ndb_b200_keys_from_sparse(const int64_t *indptr, const int32_t *indices, const float *values, int64_t n,
                              int total_dim, float *rows)
{
    NDB_CHECK(require_init());
    NDB_REQUIRE(indptr && rows && n >= 0, NDB_B200_EINVAL, "keys_from_sparse: bad argument");
    NDB_REQUIRE(total_dim > 0 && total_dim <= 32767, NDB_B200_EINVAL, "invalid sparsevec total_dim %d", total_dim);
    if (n == 0) return NDB_B200_OK;
    const int64_t nnz = indptr[n];
    NDB_REQUIRE(indptr[0] == 0 && nnz >= 0 && (nnz == 0 || (indices && values)), NDB_B200_EINVAL, "keys_from_sparse: bad CSR arrays");
    for (int64_t r = 0; r < n; r++)
        NDB_REQUIRE(indptr[r] <= indptr[r + 1], NDB_B200_EINVAL, "keys_from_sparse: indptr not monotone at row %lld", (long long) r);
    cudaStream_t s = ctx().stream;
    DevBuf dp, di, dv, dout;
    int rc = NDB_B200_OK;
    const size_t out_bytes = (size_t) n * total_dim * 4;
    if ((rc = dp.reserve((size_t) (n + 1) * 8)) == NDB_B200_OK && (rc = di.reserve(nnz ? (size_t) nnz * 4 : 16)) == NDB_B200_OK &&
        (rc = dv.reserve(nnz ? (size_t) nnz * 4 : 16)) == NDB_B200_OK && (rc = dout.reserve(out_bytes)) == NDB_B200_OK) {
        cudaError_t e = cudaMemcpyAsync(dp.p, indptr, (size_t) (n + 1) * 8, cudaMemcpyHostToDevice, s);
        if (e == cudaSuccess && nnz) e = cudaMemcpyAsync(di.p, indices, (size_t) nnz * 4, cudaMemcpyHostToDevice, s);
        if (e == cudaSuccess && nnz) e = cudaMemcpyAsync(dv.p, values, (size_t) nnz * 4, cudaMemcpyHostToDevice, s);
        if (e == cudaSuccess) e = cudaMemsetAsync(dout.p, 0, out_bytes, s);
        if (e == cudaSuccess) {
            keys_sparse_kernel<<<(unsigned) ((n + 127) / 128), 128, 0, s>>>(dp.as<int64_t>(), di.as<int32_t>(), dv.as<float>(), n,
                                                                           total_dim, dout.as<float>());
            count_launch();
            e = cudaGetLastError();
        }
        if (e == cudaSuccess) e = cudaMemcpyAsync(rows, dout.p, out_bytes, cudaMemcpyDeviceToHost, s);
        if (e == cudaSuccess) e = cudaStreamSynchronize(s);
        if (e != cudaSuccess) { set_error("keys_from_sparse: %s", cudaGetErrorString(e)); rc = NDB_B200_ECUDA; }
    }
    dp.release(); di.release(); dv.release(); dout.release();
    return rc;
}

}  // extern "C"
