// layout.cu -- IL32 layout conversion and per-vector norm kernels.
#include "layout.cuh"
#include "arith.cuh"

namespace ndb {

// one thread per (row, float4 chunk): coalesced reads along the row, 16-byte scattered writes
__global__ void il32_scatter_kernel(const float *__restrict__ rows, int64_t n, int dim, int dimp,
                                    const uint32_t *__restrict__ slot_of_row, uint32_t slot_base,
                                    float4 *__restrict__ store)
{
    const int nchunk = dimp >> 2;
    const int64_t t = (int64_t) blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n * nchunk) return;
    const int64_t row = t / nchunk;
    const int c = (int) (t - row * nchunk);
    const float *src = rows + (size_t) row * dim + 4 * c;
    float4 v;
    v.x = 4 * c + 0 < dim ? src[0] : 0.0f;
    v.y = 4 * c + 1 < dim ? src[1] : 0.0f;
    v.z = 4 * c + 2 < dim ? src[2] : 0.0f;
    v.w = 4 * c + 3 < dim ? src[3] : 0.0f;
    const uint32_t slot = slot_of_row ? slot_of_row[row] : slot_base + (uint32_t) row;
    if (slot == INVALID_SLOT) return;
    store[(size_t) (slot >> 5) * (8 * (size_t) dimp) + (size_t) c * 32 + (slot & 31)] = v;
}

__global__ void il32_gather_kernel(const float4 *__restrict__ src, const uint32_t *__restrict__ src_slot,
                                   float4 *__restrict__ dst, const uint32_t *__restrict__ dst_slot, int64_t n,
                                   int dimp)
{
    const int nchunk = dimp >> 2;
    const int64_t t = (int64_t) blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n * nchunk) return;
    const int64_t i = t / nchunk;
    const int c = (int) (t - i * nchunk);
    const uint32_t s = src_slot[i], d = dst_slot[i];
    dst[(size_t) (d >> 5) * (8 * (size_t) dimp) + (size_t) c * 32 + (d & 31)] =
        src[(size_t) (s >> 5) * (8 * (size_t) dimp) + (size_t) c * 32 + (s & 31)];
}

__global__ void il32_to_rows_kernel(const float4 *__restrict__ store, int64_t n, int dim, int dimp,
                                    float *__restrict__ rows)
{
    const int nchunk = dimp >> 2;
    const int64_t t = (int64_t) blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n * nchunk) return;
    const int64_t slot = t / nchunk;
    const int c = (int) (t - slot * nchunk);
    const float4 v = store[(size_t) (slot >> 5) * (8 * (size_t) dimp) + (size_t) c * 32 + (slot & 31)];
    float *dst = rows + (size_t) slot * dim + 4 * c;
    if (4 * c + 0 < dim) dst[0] = v.x;
    if (4 * c + 1 < dim) dst[1] = v.y;
    if (4 * c + 2 < dim) dst[2] = v.z;
    if (4 * c + 3 < dim) dst[3] = v.w;
}

static unsigned grid_for(int64_t work, int block) { return (unsigned) ((work + block - 1) / block); }

int il32_scatter(const float *rows_dev, int64_t n, int dim, int dimp, const uint32_t *slot_of_row,
                 uint32_t slot_base, float *store, cudaStream_t s)
{
    if (n <= 0) return NDB_B200_OK;
    il32_scatter_kernel<<<grid_for(n * (dimp >> 2), 256), 256, 0, s>>>(rows_dev, n, dim, dimp, slot_of_row, slot_base,
                                                                      reinterpret_cast<float4 *>(store));
    count_launch();
    NDB_CUDA(cudaGetLastError());
    return NDB_B200_OK;
}

int il32_gather(const float *src_store, const uint32_t *src_slot, float *dst_store, const uint32_t *dst_slot,
                int64_t n, int dimp, cudaStream_t s)
{
    if (n <= 0) return NDB_B200_OK;
    il32_gather_kernel<<<grid_for(n * (dimp >> 2), 256), 256, 0, s>>>(
        reinterpret_cast<const float4 *>(src_store), src_slot, reinterpret_cast<float4 *>(dst_store), dst_slot, n, dimp);
    count_launch();
    NDB_CUDA(cudaGetLastError());
    return NDB_B200_OK;
}

int il32_to_rows(const float *store, int64_t n, int dim, int dimp, float *rows_dev, cudaStream_t s)
{
    if (n <= 0) return NDB_B200_OK;
    il32_to_rows_kernel<<<grid_for(n * (dimp >> 2), 256), 256, 0, s>>>(reinterpret_cast<const float4 *>(store), n, dim,
                                                                      dimp, rows_dev);
    count_launch();
    NDB_CUDA(cudaGetLastError());
    return NDB_B200_OK;
}

// ---- norms: the cosine loops accumulate norm1/norm2 independently of the dot product, in
// dimension order; doing it once per vector yields the same bits as doing it per pair ---------
template <class P> __global__ void slot_norms_kernel(const float4 *__restrict__ store, int64_t nslots, int dim,
                                                      int dimp, typename P::N *__restrict__ out)
{
    const int64_t slot = (int64_t) blockIdx.x * blockDim.x + threadIdx.x;
    if (slot >= nslots) return;
    const float4 *vp = store + (size_t) (slot >> 5) * (8 * (size_t) dimp) + (slot & 31);
    typename P::N n = 0;
    const int nfull = dim >> 2, rem = dim & 3;
    for (int c = 0; c < nfull; c++) {
        const float4 x = vp[(size_t) c * 32];
        P::nstep(n, x.x); P::nstep(n, x.y); P::nstep(n, x.z); P::nstep(n, x.w);
    }
    if (rem) {
        const float4 x = vp[(size_t) nfull * 32];
        P::nstep(n, x.x);
        if (rem > 1) P::nstep(n, x.y);
        if (rem > 2) P::nstep(n, x.z);
    }
    out[slot] = n;
}

template <class P> __global__ void row_norms_kernel(const float *__restrict__ rows, int64_t n, int dim,
                                                     typename P::N *__restrict__ out)
{
    const int64_t r = (int64_t) blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= n) return;
    typename P::N acc = 0;
    for (int j = 0; j < dim; j++) P::nstep(acc, rows[(size_t) r * dim + j]);
    out[r] = acc;
}

size_t norm_elem_size(int arith) { return (arith == NDB_ARITH_OP_F64 || arith == NDB_ARITH_HNSW) ? 8 : 4; }

int slot_norms(int arith, const float *store, int64_t nslots, int dim, int dimp, void *out, cudaStream_t s)
{
    if (nslots <= 0) return NDB_B200_OK;
    const float4 *st = reinterpret_cast<const float4 *>(store);
    const unsigned g = grid_for(nslots, 128);
    switch (arith) {
    case NDB_ARITH_OP_F64: slot_norms_kernel<Arith<NDB_COSINE, NDB_ARITH_OP_F64>><<<g, 128, 0, s>>>(st, nslots, dim, dimp, (double *) out); break;
    case NDB_ARITH_HNSW: slot_norms_kernel<Arith<NDB_COSINE, NDB_ARITH_HNSW>><<<g, 128, 0, s>>>(st, nslots, dim, dimp, (double *) out); break;
    case NDB_ARITH_IVF_F32: slot_norms_kernel<Arith<NDB_COSINE, NDB_ARITH_IVF_F32>><<<g, 128, 0, s>>>(st, nslots, dim, dimp, (float *) out); break;
    case NDB_ARITH_FAST: slot_norms_kernel<Arith<NDB_COSINE, NDB_ARITH_FAST>><<<g, 128, 0, s>>>(st, nslots, dim, dimp, (float *) out); break;
    default: set_error("slot_norms: unsupported arith %d", arith); return NDB_B200_EINVAL;
    }
    count_launch();
    NDB_CUDA(cudaGetLastError());
    return NDB_B200_OK;
}

int row_norms(int arith, const float *rows_dev, int64_t n, int dim, void *out, cudaStream_t s)
{
    if (n <= 0) return NDB_B200_OK;
    const unsigned g = grid_for(n, 128);
    switch (arith) {
    case NDB_ARITH_OP_F64: row_norms_kernel<Arith<NDB_COSINE, NDB_ARITH_OP_F64>><<<g, 128, 0, s>>>(rows_dev, n, dim, (double *) out); break;
    case NDB_ARITH_HNSW: row_norms_kernel<Arith<NDB_COSINE, NDB_ARITH_HNSW>><<<g, 128, 0, s>>>(rows_dev, n, dim, (double *) out); break;
    case NDB_ARITH_IVF_F32: row_norms_kernel<Arith<NDB_COSINE, NDB_ARITH_IVF_F32>><<<g, 128, 0, s>>>(rows_dev, n, dim, (float *) out); break;
    case NDB_ARITH_FAST: row_norms_kernel<Arith<NDB_COSINE, NDB_ARITH_FAST>><<<g, 128, 0, s>>>(rows_dev, n, dim, (float *) out); break;
    default: set_error("row_norms: unsupported arith %d", arith); return NDB_B200_EINVAL;
    }
    count_launch();
    NDB_CUDA(cudaGetLastError());
    return NDB_B200_OK;
}

}  // namespace ndb
