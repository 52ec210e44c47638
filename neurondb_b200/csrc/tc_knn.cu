// tc_knn.cu -- GEMM-form distances on the 5th-generation tensor cores (tcgen05 + TMEM) with a
// fused per-query top-k, for the genuinely dense contractions of the path: brute-force kNN and
// (k = 1) nearest-centroid assignment over bf16 data (BASELINE config 5; tolerance path,
// |rel err| <= 1e-3, north_star).
//
//   ||x - q||^2 = ||x||^2 - 2 x.q + ||q||^2            (L2; ranked squared, sqrt on output)
//   -x.q                                                (inner product, HNSW sign convention; the query tiles hold -q,
//                                                        so the accumulator is the candidate itself)
// The same kernel scans the IVF lists of a query batch (list mode, packed keys: csrc/ivf.cu) and the centroid store
// (coarse quantiser, list assignment, k-means assignment: csrc/kmeans.cu); there its candidates are only PROPOSALS that a
// certified fp32 re-evaluation turns into the reference's exact answer (ivf_cert.cuh, cert_common.cuh).
//
// Layout.  Stored rows and queries are kept as bf16 in the UMMA *canonical K-major no-swizzle*
// layout, blocked so that every operand tile is one contiguous run of bytes:
//   X: [tile of 256 rows][K-chunk of 128 dims][16 k-groups][32 row-groups][8 rows][8 elems]  = 64 KB per (tile, chunk)
//   Q: [tile of 128 queries][K-chunk][16 k-groups][16 row-groups][8][8]                     = 32 KB per (tile, chunk)
// i.e. 8x8 "core matrices" of 128 contiguous bytes; consecutive core matrices along the
// row direction are 128 B apart (SBO), along K they are 32*128 B (X) / 16*128 B (Q) apart (LBO).
// A (tile, chunk) is fetched with ONE cp.async.bulk (no tensor map, no swizzle) and is directly
// consumable by tcgen05.mma through a shared-memory matrix descriptor.
//
// Kernel (persistent, 384 threads):
//   warp 0    TMA producer: Q tile once per work item (double-buffered), then X (tile, chunk) stages into a ring of
//             2-5 stages (144 KB; only the K groups that carry data are copied: 48 KB stages at dim 96)
//   warp 1    MMA issuer: one elected thread issues up to 8 x tcgen05.mma.kind::f16 (M=128, N=256, K=16) per
//             stage into one of two 256-column TMEM accumulators; tcgen05.commit frees the smem
//             stage and, after the last chunk, publishes the accumulator
//   warp 2    TMEM allocator (512 columns)
//   warps 4-11 epilogue: thread = (query = TMEM lane, half of the 256 columns); tcgen05.ld 32 columns
//             at a time; candidates (L2: fma(-2, dot, ||x||^2); inner product: the accumulator) are min-reduced
//             with 3-input minima and compared against the thread's k-th best once per 32 (inner product: per 64);
//             the top-k is a thread-local sorted list -- no cross-lane traffic at all.  The work items of a query
//             share an upper bound of its k-th best through a global float (p.gthr), in list and in dense mode.
// Work item = (query tile, range of X tiles); per-item top-k lists are merged by (dist, id) by
// merge_parts_kernel (scan.cuh) or, in list mode, by the certified finish.
#include "layout.cuh"
#include "scan.cuh"
#include "tc.cuh"
#include "cert_bound.cuh"

#include <cuda_bf16.h>
#include <algorithm>
#include <cfloat>

namespace ndb {

constexpr int TC_XSTAGE_BYTES = TC_N * TC_KC * 2;      // 64 KB: one (tile, K-chunk) of stored rows
// The ring of stored-row stages: 144 KB.  Rows of more than one K-chunk use it as 2 stages of 64 KB.  With one chunk
// (dim <= 128) only the K groups that carry data are copied, a stage is kg_last * 4 KB and the ring holds
// min(5, 144 KB / stage) of them -- dim 96: three 48 KB stages, dim 64: four of 32 KB.  The ring is latency-bound (a
// stage is refilled ~2 us after the MMAs that read it retire), so tiles in flight are what sets the pace.
constexpr int TC_XRING_BYTES = 144 * 1024;
constexpr int TC_MAX_STAGES = 5;                       // < TC_NORM_RING - 2: the norms of a tile stay until its epilogue
constexpr int TC_QCHUNK_BYTES = TC_M * TC_KC * 2;      // 32 KB
constexpr int TC_NORM_RING = 8;

// ---- layout conversion -------------------------------------------------------------------
// IL32 fp32 store -> blocked bf16 + per-row squared norm of the rounded values (fp32)
// src_slot (optional): tensor row r is IL32 slot src_slot[r] (INVALID_SLOT = pad row); identity if null
__global__ void tc_block_rows_kernel(const float4 *__restrict__ store, const uint32_t *__restrict__ src_slot, int64_t n,
                                     int64_t npad, int dim, int dimp, int nkc, __nv_bfloat16 *__restrict__ xb,
                                     float *__restrict__ xnorm)
{
    const int groups = nkc * (TC_KC / 8);                      // 8-element groups per row (padded dims)
    const int64_t t = (int64_t) blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= npad * groups) return;
    const int64_t row = t / groups;
    const int g = (int) (t - row * groups);
    float v[8];
#pragma unroll
    for (int i = 0; i < 8; i++) v[i] = 0.0f;
    const int64_t srow = src_slot ? (row < npad && src_slot[row] != INVALID_SLOT ? (int64_t) src_slot[row] : -1) : (row < n ? row : -1);
    if (srow >= 0) {
#pragma unroll
        for (int h = 0; h < 2; h++) {
            const int c = 2 * g + h;                           // float4 chunk of the row
            if (4 * c < dimp) {
                const float4 x = store[(size_t) (srow >> 5) * (8 * (size_t) dimp) + (size_t) c * 32 + (srow & 31)];
                v[4 * h + 0] = 4 * c + 0 < dim ? x.x : 0.0f;
                v[4 * h + 1] = 4 * c + 1 < dim ? x.y : 0.0f;
                v[4 * h + 2] = 4 * c + 2 < dim ? x.z : 0.0f;
                v[4 * h + 3] = 4 * c + 3 < dim ? x.w : 0.0f;
            }
        }
    }
    const int64_t tile = row / TC_N;
    const int rr = (int) (row - tile * TC_N);
    const int chunk = g / (TC_KC / 8), kc = g % (TC_KC / 8);
    const size_t off = ((((size_t) (tile * nkc + chunk) * (TC_KC / 8) + kc) * (TC_N / 8) + (rr >> 3)) * 8 + (rr & 7)) * 8;
    __nv_bfloat16 o[8];
#pragma unroll
    for (int i = 0; i < 8; i++) o[i] = __float2bfloat16_rn(v[i]);
    *reinterpret_cast<uint4 *>(xb + off) = *reinterpret_cast<const uint4 *>(o);
    (void) xnorm;
}

// one thread per row: squared norm of the bf16-rounded row; +inf for pad rows so they never rank.
// stats (4 floats, zeroed by the caller) receives, over all stored rows, the maxima the certified
// selection needs to bound what bf16 rounding can do to a distance (ivf.cu, "certified selection"):
//   [0] max ||x - bf16(x)||^2   [1] max ||x||^2   [2] max ||x - bf16(x)||^2 / ||bf16(x)||^2   [3] max ||x||^2 / ||bf16(x)||^2
__global__ void tc_row_norms_kernel(const float4 *__restrict__ store, const uint32_t *__restrict__ src_slot, int64_t n,
                                    int64_t npad, int dim, int dimp, float *__restrict__ xnorm, float *__restrict__ stats)
{
    const int64_t row = (int64_t) blockIdx.x * blockDim.x + threadIdx.x;
    float e2 = 0.0f, x2 = 0.0f, acc = 0.0f;
    bool real = false;
    if (row < npad) {
        const int64_t srow = src_slot ? (src_slot[row] != INVALID_SLOT ? (int64_t) src_slot[row] : -1) : (row < n ? row : -1);
        if (srow < 0) {
            xnorm[row] = INFINITY;
        } else {
            real = true;
            const float4 *vp = store + (size_t) (srow >> 5) * (8 * (size_t) dimp) + (srow & 31);
            for (int c = 0; 4 * c < dimp; c++) {
                const float4 x = vp[(size_t) c * 32];
                const float xs[4] = {4 * c + 0 < dim ? x.x : 0.0f, 4 * c + 1 < dim ? x.y : 0.0f, 4 * c + 2 < dim ? x.z : 0.0f,
                                     4 * c + 3 < dim ? x.w : 0.0f};
#pragma unroll
                for (int i = 0; i < 4; i++) {
                    const float r = __bfloat162float(__float2bfloat16_rn(xs[i]));
                    acc = fmaf(r, r, acc);
                    x2 = fmaf(xs[i], xs[i], x2);
                    e2 = fmaf(xs[i] - r, xs[i] - r, e2);
                }
            }
            xnorm[row] = acc;
        }
    }
    if (!stats) return;
    // rounded up a little: these feed upper bounds, and the fmaf chains above round to nearest
    float v0 = real ? e2 * 1.0001f : 0.0f, v1 = real ? x2 * 1.0001f : 0.0f;
    float v2 = (real && acc > 0.0f) ? e2 / acc * 1.0002f : 0.0f, v3 = (real && acc > 0.0f) ? x2 / acc * 1.0002f : 0.0f;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        v0 = fmaxf(v0, __shfl_xor_sync(FULL, v0, o));
        v1 = fmaxf(v1, __shfl_xor_sync(FULL, v1, o));
        v2 = fmaxf(v2, __shfl_xor_sync(FULL, v2, o));
        v3 = fmaxf(v3, __shfl_xor_sync(FULL, v3, o));
    }
    if ((threadIdx.x & 31) == 0) {        // non-negative floats order like their bit patterns
        atomicMax(reinterpret_cast<int *>(stats) + 0, __float_as_int(v0));
        atomicMax(reinterpret_cast<int *>(stats) + 1, __float_as_int(v1));
        atomicMax(reinterpret_cast<int *>(stats) + 2, __float_as_int(v2));
        atomicMax(reinterpret_cast<int *>(stats) + 3, __float_as_int(v3));
    }
}

// row-major fp32 queries -> blocked bf16 query tiles + squared norms
// qmap (optional): tile position q holds query qmap[q] / nprobe (INVALID_SLOT = empty position)
// One CTA per (query tile, K-chunk): the 32 KB blocked tile-chunk is assembled in shared memory (16 adjacent lanes = the
// sixteen 8-element groups of one position, so the norm is a fixed xor tree whatever position a query sits on) and
// written out as one contiguous, coalesced run -- the direct version wrote 16 bytes per thread at a 2 KB stride.
__global__ void __launch_bounds__(256) tc_block_queries_kernel(const float *__restrict__ Q, const uint32_t *__restrict__ qmap, uint32_t nprobe,
                                                               int nq, int nqpad, const uint32_t *__restrict__ npos, int dim, int nkc,
                                                               __nv_bfloat16 *__restrict__ qb, float *__restrict__ qnorm,
                                                               float *__restrict__ qerr, bool negate)
{
    __shared__ __align__(16) __nv_bfloat16 tile_s[TC_M * TC_KC];
    const int tile = blockIdx.x / nkc, chunk = blockIdx.x % nkc;
    if (npos) nqpad = min(nqpad, (int) *npos);          // both are multiples of 128
    if (tile * TC_M >= nqpad) return;
    constexpr int G = TC_KC / 8;                        // 16 groups per position and chunk
    for (int piece = threadIdx.x; piece < TC_M * G; piece += 256) {
        const int qi_ = piece / G, kc = piece % G;
        const int q = tile * TC_M + qi_;
        const int g = chunk * G + kc;
        const int64_t src = qmap ? (qmap[q] != INVALID_SLOT ? (int64_t) (qmap[q] / nprobe) : -1) : (q < nq ? q : -1);
        __nv_bfloat16 o[8];
        float part = 0.0f, perr = 0.0f;
        if (src >= 0 && g * 8 + 8 <= dim && (dim & 3) == 0) {
            const float4 v0 = *reinterpret_cast<const float4 *>(Q + (size_t) src * dim + g * 8);
            const float4 v1 = *reinterpret_cast<const float4 *>(Q + (size_t) src * dim + g * 8 + 4);
            const float vv[8] = {v0.x, v0.y, v0.z, v0.w, v1.x, v1.y, v1.z, v1.w};
#pragma unroll
            for (int i = 0; i < 8; i++) {
                o[i] = __float2bfloat16_rn(vv[i]);
                const float r = __bfloat162float(o[i]);
                part = fmaf(r, r, part);
                perr = fmaf(vv[i] - r, vv[i] - r, perr);
                if (negate) o[i] = __float2bfloat16_rn(-vv[i]);      // (rounding is symmetric: exactly -o[i])
            }
        } else {
#pragma unroll
            for (int i = 0; i < 8; i++) {
                const int d = g * 8 + i;
                const float v = (src >= 0 && d < dim) ? Q[(size_t) src * dim + d] : 0.0f;
                o[i] = __float2bfloat16_rn(v);
                const float r = __bfloat162float(o[i]);
                part = fmaf(r, r, part);
                perr = fmaf(v - r, v - r, perr);
                if (negate) o[i] = __float2bfloat16_rn(-v);
            }
        }
        // Query i of a tile sits on TMEM lane (i % 4) * 32 + i / 4: the tile's queries are dealt round-robin
        // to the four lane quarters, each of which only one pair of epilogue warps can read.  A tile with 16
        // live queries (most tile-steps of an IVF batch: long lists probed by a few queries) then keeps all
        // eight epilogue warps busy with 4 queries each instead of two warps with 16.
        const int rr = (qi_ & 3) * 32 + (qi_ >> 2);
        const size_t soff = (((size_t) kc * (TC_M / 8) + (rr >> 3)) * 8 + (rr & 7)) * 8;
        *reinterpret_cast<uint4 *>(tile_s + soff) = *reinterpret_cast<const uint4 *>(o);
        // squared norm of the rounded query (one K-chunk: dim <= 128; more chunks: tc_query_norms_kernel)
        if (nkc == 1) {
#pragma unroll
            for (int o2 = G >> 1; o2 > 0; o2 >>= 1) {
                part += __shfl_xor_sync(FULL, part, o2);
                perr += __shfl_xor_sync(FULL, perr, o2);
            }
            if (kc == 0) { qnorm[q] = part; if (qerr) qerr[q] = perr; }   // qerr: squared norm of the query's bf16 rounding error
        }
    }
    __syncthreads();
    uint4 *dst = reinterpret_cast<uint4 *>(qb + (size_t) (tile * nkc + chunk) * TC_M * TC_KC);
    const uint4 *srcp = reinterpret_cast<const uint4 *>(tile_s);
    for (int i = threadIdx.x; i < TC_M * TC_KC / 8; i += 256) dst[i] = srcp[i];
}

// squared norm of the bf16-rounded query, one thread per tile position (rows of more than 256 dims)
__global__ void tc_query_norms_kernel(const float *__restrict__ Q, const uint32_t *__restrict__ qmap, uint32_t nprobe, int nq,
                                      int nqpad, const uint32_t *__restrict__ npos, int dim, float *__restrict__ qnorm,
                                      float *__restrict__ qerr)
{
    const int q = blockIdx.x * blockDim.x + threadIdx.x;
    if (npos) nqpad = min(nqpad, (int) *npos);
    if (q >= nqpad) return;
    const int64_t src = qmap ? (qmap[q] != INVALID_SLOT ? (int64_t) (qmap[q] / nprobe) : -1) : (q < nq ? q : -1);
    float acc = 0.0f, err = 0.0f;
    if (src >= 0)
        for (int d = 0; d < dim; d++) {
            const float v = Q[(size_t) src * dim + d];
            const float r = __bfloat162float(__float2bfloat16_rn(v));
            acc = fmaf(r, r, acc);
            err = fmaf(v - r, v - r, err);
        }
    qnorm[q] = acc;
    if (qerr) qerr[q] = err;
}

// ---- tcgen05 / TMEM primitives -------------------------------------------------------------
__device__ __forceinline__ void tmem_alloc(uint32_t *holder_smem, uint32_t ncols)
{
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(holder_smem)), "r"(ncols) : "memory");
}
__device__ __forceinline__ void tmem_relinquish() { asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols)
{
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
// D[tmem] (+)= A[smem desc] * B[smem desc]^T, bf16 in, fp32 accumulate, issued by one thread
__device__ __forceinline__ void umma_bf16(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate)
{
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t"
        "}\n" ::"r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate) : "memory");
}
// arrives on an mbarrier once every previously issued tcgen05.mma of this thread has completed
__device__ __forceinline__ void umma_commit(uint64_t *bar)
{
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&r)[32])
{
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
          "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
          "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(taddr)
        : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// shared-memory matrix descriptor, K-major, no swizzle (cute::UMMA::SmemDescriptor bit layout):
//   [0,14) start address >> 4 | [16,30) LBO >> 4 (K-direction core-matrix stride) |
//   [32,46) SBO >> 4 (row-direction 8-row-group stride) | [46,48) version = 1 | [61,64) layout = 0
__device__ __forceinline__ uint64_t umma_desc(uint32_t smem_addr, uint32_t lbo_bytes, uint32_t sbo_bytes)
{
    return (uint64_t) ((smem_addr >> 4) & 0x3fffu) | ((uint64_t) ((lbo_bytes >> 4) & 0x3fffu) << 16) |
           ((uint64_t) ((sbo_bytes >> 4) & 0x3fffu) << 32) | (1ull << 46);
}
// instruction descriptor (cute::UMMA::InstrDescriptor): D = F32 (bits 4-5 = 1), A = B = BF16
// (bits 7-9, 10-12 = 1), both K-major (bits 15, 16 = 0), N >> 3 at bit 17, M >> 4 at bit 24
__host__ __device__ constexpr uint32_t umma_idesc_bf16(int m, int n)
{
    return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t) (n >> 3) << 17) | ((uint32_t) (m >> 4) << 24);
}

// thread-local sorted top-KT list in registers: branch-free insertion of (d, id).
//   L'[j] = L[j]            if L[j] precedes the new pair
//         = new             if L[j-1] precedes it (or j == 0) but L[j] does not
//         = L[j-1]          otherwise (shifted down; the last entry falls off)
// One copy of this code sits in a loop over the (rare) passing candidates of a 32-column chunk:
// inlining it per candidate blows the instruction cache (ncu: no_instruction 17.8 cycles/issue),
// and a called version with the list in local memory serialises on LDL/STL latency.
template <int KT>
__device__ __forceinline__ void tk_insert(float (&bd)[KT], uint32_t (&bi)[KT], float d, uint32_t id)
{
    bool lt[KT];
#pragma unroll
    for (int j = 0; j < KT; j++) lt[j] = bd[j] < d || (bd[j] == d && bi[j] < id);
#pragma unroll
    for (int j = KT - 1; j >= 0; j--) {
        const bool take_new = j == 0 ? !lt[0] : (lt[j - 1] && !lt[j]);
        const bool take_prev = j > 0 && !lt[j - 1];
        const float pd = j > 0 ? bd[j - 1] : d;
        const uint32_t pi = j > 0 ? bi[j - 1] : id;
        bd[j] = take_new ? d : (take_prev ? pd : bd[j]);
        bi[j] = take_new ? id : (take_prev ? pi : bi[j]);
    }
}

// c[i] for a run-time i with c held in registers: 5-level select tree on the bits of i
__device__ __forceinline__ float pick32(const float (&c)[32], int i)
{
    float t16[16], t8[8], t4[4], t2[2];
#pragma unroll
    for (int j = 0; j < 16; j++) t16[j] = (i & 1) ? c[2 * j + 1] : c[2 * j];
#pragma unroll
    for (int j = 0; j < 8; j++) t8[j] = (i & 2) ? t16[2 * j + 1] : t16[2 * j];
#pragma unroll
    for (int j = 0; j < 4; j++) t4[j] = (i & 4) ? t8[2 * j + 1] : t8[2 * j];
#pragma unroll
    for (int j = 0; j < 2; j++) t2[j] = (i & 8) ? t4[2 * j + 1] : t4[2 * j];
    return (i & 16) ? t2[1] : t2[0];
}

// spin (no suspend-time hint): every tile hand-off between producer, MMA issuer and epilogue sits on
// the critical path, and a suspended waiter wakes up late
__device__ __forceinline__ void mbar_spin(uint64_t *bar, uint32_t parity)
{
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "SPIN_%=:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@p bra OUT_%=;\n\t"
        "bra SPIN_%=;\n\t"
        "OUT_%=:\n\t"
        "}" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}

// ---- packed keys (short items: IVF lists, coarse quantizer) --------------------------------------
// Where an item holds only a few hundred candidates per query, most of the epilogue's time would go
// into list insertions (a thread's k-th best tightens slowly, and a warp pays for its slowest
// lane).  For such items the candidate distance and its index within the item are packed into ONE
// float: the low TC_IDX_BITS mantissa bits carry the index, so that a compare-exchange is two
// FMNMX and whole groups of 16 candidates can go through a sorting network instead of being
// inserted one by one.  The 12 remaining mantissa bits (2.4e-4) are finer than the bf16 products
// that produced the value.
static_assert(TC_PACKED_MAX_TILES == 1 << (TC_IDX_BITS - 7), "index bits");
constexpr float TC_KEY_BIG = 1.7014118346046923e38f;             // 2^127 (index bits all zero): keys at or above are pad
                                                                 // rows / empty slots, also once an index is packed in
constexpr int TC_HEAVY = 6;                                      // passes per 32 columns above which the warp sorts

template <int N> __device__ __forceinline__ void oem_sort(float (&a)[N])     // Batcher odd-even merge sort, ascending
{
#pragma unroll
    for (int p = 1; p < N; p <<= 1)
#pragma unroll
        for (int k = p; k >= 1; k >>= 1)
#pragma unroll
            for (int j = k % p; j <= N - 1 - k; j += 2 * k)
#pragma unroll
                for (int i = 0; i <= (k - 1 < N - j - k - 1 ? k - 1 : N - j - k - 1); i++)
                    if ((i + j) / (2 * p) == (i + j + k) / (2 * p)) {
                        const float lo = fminf(a[i + j], a[i + j + k]), hi = fmaxf(a[i + j], a[i + j + k]);
                        a[i + j] = lo;
                        a[i + j + k] = hi;
                    }
}
template <int N> __device__ __forceinline__ void bitonic_merge(float (&m)[N])   // bitonic -> ascending
{
#pragma unroll
    for (int s = N / 2; s >= 1; s >>= 1)
#pragma unroll
        for (int i = 0; i < N; i++)
            if (!(i & s)) {
                const float lo = fminf(m[i], m[i + s]), hi = fmaxf(m[i], m[i + s]);
                m[i] = lo;
                m[i + s] = hi;
            }
}
__device__ __forceinline__ float tc_pack(float c, uint32_t idx)
{
    return __uint_as_float((__float_as_uint(fminf(c, TC_KEY_BIG)) & ~TC_IDX_MASK) | idx);
}
// L (ascending, 16) <- the 16 smallest of L and the 16 keys g
__device__ __forceinline__ void tc_sort_merge16(float (&L)[16], float (&g)[16])
{
    oem_sort<16>(g);
#pragma unroll
    for (int i = 0; i < 16; i++) L[i] = fminf(L[i], g[15 - i]);
    bitonic_merge<16>(L);
}
// branch-free insertion of key x into the ascending list L[0..KT)
template <int KT> __device__ __forceinline__ void tc_key_insert(float (&L)[16], float x)
{
#pragma unroll
    for (int j = KT - 1; j >= 1; j--) L[j] = fmaxf(L[j - 1], fminf(L[j], x));
    L[0] = fminf(L[0], x);
}

// next representable float above a finite g
__device__ __forceinline__ float float_next_up(float g)
{
    const uint32_t b = __float_as_uint(g + 0.0f);              // -0 -> +0
    return __uint_as_float(g >= 0.0f ? b + 1u : b - 1u);
}

// float atomic min for values of either sign (the cell starts at a positive value)
__device__ __forceinline__ void atomic_min_f32(float *a, float v)
{
    if (v >= 0.0f) atomicMin(reinterpret_cast<int *>(a), __float_as_int(v));
    else atomicMax(reinterpret_cast<unsigned *>(a), __float_as_uint(v));
}

// Items are dealt to the persistent CTAs in boustrophedon order: round r goes 0..G-1 when r is even and
// G-1..0 when odd.  The callers emit items longest first, and a plain round-robin would hand CTA 0 the
// longest item of every round.
__device__ __forceinline__ uint32_t tc_item_of(uint32_t round, uint32_t cta, uint32_t ncta)
{
    return round * ncta + ((round & 1u) ? ncta - 1u - cta : cta);
}

template <int KT, int METRIC, bool PACKED>
__global__ void __launch_bounds__(384, 1) tc_knn_kernel(const TcParams p)
{
    extern __shared__ __align__(1024) unsigned char tsm[];
    // [Q tile: nkc * 32 KB][X ring: 2 * 64 KB][norm ring: 8 * 1 KB]
    unsigned char *q_smem = tsm;
    unsigned char *x_smem = tsm + (size_t) TC_MAX_CHUNKS * TC_QCHUNK_BYTES;
    float *n_smem = reinterpret_cast<float *>(x_smem + (size_t) TC_XRING_BYTES);
    __shared__ __align__(8) uint64_t full_bar[TC_MAX_STAGES], empty_bar[TC_MAX_STAGES], q_full[2], q_empty[2], acc_full[2], acc_empty[2];
    const uint32_t nstages = (uint32_t) p.nstages, xstride = (uint32_t) p.xstride;
    __shared__ uint32_t tmem_holder;

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    // this launch's items: [item_lo, item_lo + nitems)
    const uint32_t item_lo = p.item_lo_ptr ? *p.item_lo_ptr : 0u;
    const uint32_t nitems = (p.item_hi_ptr ? *p.item_hi_ptr : (p.nitems_ptr ? *p.nitems_ptr : p.nitems)) - item_lo;

    if (tid == 0) {
        for (int s = 0; s < TC_MAX_STAGES; s++) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], 1); }
        for (int b = 0; b < 2; b++) { mbar_init(&q_full[b], 1); mbar_init(&q_empty[b], 1); }
        for (int a = 0; a < 2; a++) { mbar_init(&acc_full[a], 1); mbar_init(&acc_empty[a], 256); }
        mbar_fence_init();
    }
    if (warp == 2) { tmem_alloc(&tmem_holder, 512); tmem_relinquish(); }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = tmem_holder;
    // the Q region holds TC_MAX_CHUNKS 32 KB chunks: with one K-chunk (dim <= 128) it is used as two
    // buffers, so that the next item's query tile loads while this item's MMAs run
    const uint32_t nqbuf = p.nkc == 1 ? 2u : 1u;
    // More than TC_MAX_CHUNKS K-chunks (dim > 256): the query tile no longer fits next to the X ring, so
    // its chunks are streamed with the X chunks instead -- stage s holds X chunk c in the ring and Q chunk
    // c in slot s of the Q region, both under full_bar[s] / empty_bar[s].  The Q chunk is re-read (from
    // L2) for every stored tile: +50 % shared-memory fill traffic, same HBM traffic.
    const bool q_streamed = p.nkc > TC_MAX_CHUNKS;

    if (warp == 0) {
        // ===== TMA producer =====
        if (lane == 0) {
            uint32_t tile_it = 0, item_it = 0, s = 0, sph = 0;              // ring slot and its phase
            // the last K-chunk of a row holds kg_last 8-element groups that are not all padding; in the blocked layout the
            // groups of a (tile, chunk) are contiguous in K order, so only that prefix is copied (and multiplied)
            const uint32_t x_last = (uint32_t) p.kg_last * (TC_N / 8) * 128u, q_last = (uint32_t) p.kg_last * (TC_M / 8) * 128u;
            for (uint32_t item = tc_item_of(0, blockIdx.x, gridDim.x); item < nitems; item = tc_item_of(++item_it, blockIdx.x, gridDim.x)) {
                const TcItem it = p.items[item_lo + item];
                const uint32_t qt = it.qtile;
                const uint32_t qb_i = item_it % nqbuf, qn_i = item_it / nqbuf;
                if (!q_streamed) {
                    mbar_spin(&q_empty[qb_i], (qn_i & 1u) ^ 1u);       // MMA finished with this buffer's previous tile
                    mbar_arrive_expect_tx(&q_full[qb_i], (uint32_t) (p.nkc - 1) * TC_QCHUNK_BYTES + q_last);
                    for (int c = 0; c < p.nkc; c++)
                        tma_bulk_g2s(q_smem + (size_t) (qb_i * p.nkc + c) * TC_QCHUNK_BYTES,
                                     reinterpret_cast<const unsigned char *>(p.qb) + ((size_t) qt * p.nkc + c) * TC_QCHUNK_BYTES,
                                     c == p.nkc - 1 ? q_last : (uint32_t) TC_QCHUNK_BYTES, &q_full[qb_i]);
                }
                const uint32_t t0 = it.t0, t1 = it.t1;
                for (uint32_t t = t0; t < t1; t++, tile_it++) {
                    for (int c = 0; c < p.nkc; c++) {
                        mbar_spin(&empty_bar[s], sph ^ 1u);
                        const bool skip_x = (p.debug_mode & 4) != 0;
                        const bool lastc = c == p.nkc - 1;
                        const uint32_t xbytes = lastc ? x_last : (uint32_t) TC_XSTAGE_BYTES, qbytes = lastc ? q_last : (uint32_t) TC_QCHUNK_BYTES;
                        const uint32_t bytes = (skip_x ? 0 : xbytes) + (c == 0 ? TC_N * 4 : 0) + (q_streamed ? qbytes : 0);
                        mbar_arrive_expect_tx(&full_bar[s], bytes);
                        if (q_streamed)
                            tma_bulk_g2s(q_smem + (size_t) s * TC_QCHUNK_BYTES,
                                         reinterpret_cast<const unsigned char *>(p.qb) + ((size_t) qt * p.nkc + c) * TC_QCHUNK_BYTES,
                                         qbytes, &full_bar[s]);
                        if (!skip_x) tma_bulk_g2s(x_smem + (size_t) s * xstride,
                                     reinterpret_cast<const unsigned char *>(p.xb) + ((size_t) t * p.nkc + c) * TC_XSTAGE_BYTES,
                                     xbytes, &full_bar[s]);
                        if (c == 0)
                            tma_bulk_g2s(n_smem + (size_t) (tile_it % TC_NORM_RING) * TC_N, p.xnorm + (size_t) t * TC_N, TC_N * 4,
                                         &full_bar[s]);
                        if (++s == nstages) { s = 0; sph ^= 1u; }
                    }
                }
            }
        }
    } else if (warp == 1) {
        // ===== MMA issuer =====
        if (lane == 0) {
            const uint32_t idesc = umma_idesc_bf16(TC_M, TC_N);
            uint32_t tile_it = 0, item_it = 0, s = 0, sph = 0;
            for (uint32_t item = tc_item_of(0, blockIdx.x, gridDim.x); item < nitems; item = tc_item_of(++item_it, blockIdx.x, gridDim.x)) {
                const TcItem it = p.items[item_lo + item];
                const uint32_t t0 = it.t0, t1 = it.t1;
                const uint32_t qb_i = item_it % nqbuf, qn_i = item_it / nqbuf;
                if (!q_streamed) {
                    mbar_spin(&q_full[qb_i], qn_i & 1u);
                    tc_fence_after();
                }
                for (uint32_t t = t0; t < t1; t++, tile_it++) {
                    const uint32_t a = tile_it & 1u;
                    mbar_spin(&acc_empty[a], ((tile_it >> 1) & 1u) ^ 1u);      // epilogue drained this accumulator
                    tc_fence_after();
                    const uint32_t tmem_d = tmem_base + a * TC_N;
                    for (int c = 0; c < p.nkc; c++) {
                        mbar_spin(&full_bar[s], sph);
                        tc_fence_after();
                        const uint32_t qa = smem_u32(q_smem + (size_t) (q_streamed ? s : qb_i * p.nkc + c) * TC_QCHUNK_BYTES);
                        const uint32_t xa = smem_u32(x_smem + (size_t) s * xstride);
                        const int nks = (c == p.nkc - 1 ? p.kg_last : TC_KC / 8) / 2;
#pragma unroll
                        for (int ks = 0; ks < TC_KC / 16; ks++) {
                            // one MMA consumes K = 16 = two 8-element core matrices along K
                            const uint64_t da = umma_desc(qa + ks * 2 * (TC_M / 8) * 128, (TC_M / 8) * 128, 128);
                            const uint64_t db = umma_desc(xa + ks * 2 * (TC_N / 8) * 128, (TC_N / 8) * 128, 128);
                            if (ks < nks && !(p.debug_mode & 2)) umma_bf16(tmem_d, da, db, idesc, (c > 0 || ks > 0) ? 1u : 0u);
                        }
                        umma_commit(&empty_bar[s]);                     // smem stage reusable once these MMAs retire
                        if (++s == nstages) { s = 0; sph ^= 1u; }
                    }
                    umma_commit(&acc_full[a]);                          // accumulator complete
                }
                if (!q_streamed) umma_commit(&q_empty[qb_i]);           // Q buffer reusable
            }
        }
    } else if (warp >= 4) {
        // ===== epilogue: 8 warps; thread = (query lane, column half) =====
        // warps 4-7 read accumulator columns [0,128), warps 8-11 columns [128,256) of the same 128
        // TMEM lanes (a warp may only touch the lane quarter 32*(warp%4)); each thread keeps its own
        // top-KT list and the two halves are merged with the other parts afterwards.
        const int ql = lane * 4 + (warp & 3);                           // query within the tile; it sits on TMEM lane
                                                                        // (warp & 3) * 32 + lane (tc_block_queries_kernel)
        const int half = (warp - 4) >> 2;
        const uint32_t lane_addr = (uint32_t) ((warp & 3) * 32) << 16;
        uint32_t tile_it = 0;
        uint32_t item_it = 0;
#ifdef NDB_TC_COUNTERS
        uint32_t dc_chunks = 0, dc_any = 0, dc_heavy = 0, dc_iters = 0, dc_takers = 0;
#endif
        for (uint32_t item = tc_item_of(0, blockIdx.x, gridDim.x); item < nitems; item = tc_item_of(++item_it, blockIdx.x, gridDim.x)) {
            const TcItem it = p.items[item_lo + item];
            const uint32_t t0 = it.t0, t1 = it.t1;
            const float qn = p.qnorm[(size_t) it.qtile * TC_M + ql];
            const bool live = (uint32_t) ql < it.nq;
            float bd[PACKED ? 16 : KT];                                 // PACKED: keys (distance | index), ascending
            uint32_t bi[PACKED ? 1 : KT];
#pragma unroll
            for (int j = 0; j < (PACKED ? 16 : KT); j++) bd[j] = PACKED ? FLT_MAX : INFINITY;
#pragma unroll
            for (int j = 0; j < (PACKED ? 1 : KT); j++) bi[j] = INVALID_SLOT;
            // list mode: the items of one query (its probed lists, their segments, the two column
            // halves) share an upper bound of the query's KT-th best candidate.  Whatever order the
            // items run in, the bound only ever filters candidates that cannot be in the merged
            // top-KT, so the merged result does not depend on the schedule.  Lanes without a query
            // (ql >= nq) never take a candidate.
            float *gcell = nullptr;
            float gcap = live ? INFINITY : -INFINITY, published = INFINITY;
            // certified selection (cert_bound.cuh): the bound this list publishes is relaxed by what bf16 rounding can
            // do to a distance, so that every non-full list of the query stays complete below it
            CertQ cq;
            if (p.cstats) {
                const float eq = sqrtf(p.qerr[(size_t) it.qtile * TC_M + ql]) * 1.0002f, qr = sqrtf(qn);
                cq.eq = eq; cq.qnr = qr * 1.0002f; cq.qn = (qr + eq) * 1.0002f; cq.qn_lo = fmaxf(qr - eq, 0.0f) * 0.9998f;
            }
            if (p.gthr && live) {
                if (p.qmap) {
                    const uint32_t pr = p.qmap[(size_t) it.qtile * TC_M + ql];
                    if (pr != INVALID_SLOT) gcell = p.gthr + pr / p.nprobe;
                } else {
                    gcell = p.gthr + (size_t) it.qtile * TC_M + ql;         // dense mode: one cell per tile position
                }
            }
            float thr = gcap;                                           // KT-th best so far, in "candidate" units
            // candidate units: L2 -> ||x||^2 - 2 x.q (PACKED: + ||q||^2, i.e. the squared distance);
            // IP -> -x.q.  Pad rows carry +inf norms and never rank.
            // PACKED keys and the shared bound are in units of (candidate + cadd); the tile loop compares raw
            // candidates against thr - cadd
            const float cadd = (PACKED && METRIC == NDB_L2) ? qn : 0.0f;
            // the bound is read one tile ahead (the load's latency hides behind the wait for the
            // accumulator or the previous tile's work); a slightly stale bound is still a bound
            float gnext = FLT_MAX;
            if (gcell) gnext = *reinterpret_cast<volatile float *>(gcell);
            for (uint32_t t = t0; t < t1; t++, tile_it++) {
                const uint32_t a = tile_it & 1u;
                mbar_spin(&acc_full[a], (tile_it >> 1) & 1u);
                tc_fence_after();
                if (gcell) {
                    // strictly above the shared bound: ties with the bound itself must survive
                    gcap = fminf(gcap, float_next_up(gnext));
                    thr = fminf(thr, gcap);
                    if (t + 1 < t1) gnext = *reinterpret_cast<volatile float *>(gcell);
                }
                const float *xn = n_smem + (size_t) (tile_it % TC_NORM_RING) * TC_N;
                const uint32_t tile_rows = min((uint32_t) TC_N, it.nrows - (t - t0) * TC_N);     // stored rows of this tile
                // PACKED (short items, latency-bound epilogue): the TMEM read of chunk j + 1 is in flight
                // while chunk j is processed.  The dense kernel is ALU-bound and keeps the plain order.
                // a warp whose 32 query lanes are all beyond the item's query count has nothing to read: it
                // only keeps the accumulator hand-shake in step (most tile-steps of an IVF batch belong to
                // long lists probed by a handful of queries)
                const bool warp_dead = (uint32_t) (warp & 3) >= it.nq;          // (its first lane's query is the lowest)
                // replicated queries (it.rep == 4): this warp's lanes are replica (warp & 3) and scan chunk (warp & 3) only
                const int jlo = it.rep == 4 ? (warp & 3) : 0;
                const int nchunk = ((p.debug_mode & 1) || warp_dead) ? 0 : (it.rep == 4 ? jlo + 1 : TC_N / 64);
                const uint32_t acc_addr = tmem_base + lane_addr + a * TC_N + half * (TC_N / 2);
                bool changed = false;                                   // did this tile put anything into the list
                // candidates of one 32-column chunk (its accumulator values are in registers) and their minimum
                constexpr int NN4 = METRIC == NDB_IP ? 1 : 8;
                auto candidates = [&](const uint32_t (&v)[32], const float4 (&n4s)[NN4], float (&c)[32]) -> float {
                    if (METRIC == NDB_IP) {
                        // the query tile holds -q: the accumulator is the candidate itself, no norms
#pragma unroll
                        for (int i = 0; i < 32; i++) c[i] = __uint_as_float(v[i]);
                    } else {
#pragma unroll
                        for (int i4 = 0; i4 < 8; i4++) {
                            const float4 n4 = n4s[i4 < NN4 ? i4 : 0];
                            if (METRIC == NDB_L2) {
                                c[4 * i4 + 0] = fmaf(-2.0f, __uint_as_float(v[4 * i4 + 0]), n4.x);
                                c[4 * i4 + 1] = fmaf(-2.0f, __uint_as_float(v[4 * i4 + 1]), n4.y);
                                c[4 * i4 + 2] = fmaf(-2.0f, __uint_as_float(v[4 * i4 + 2]), n4.z);
                                c[4 * i4 + 3] = fmaf(-2.0f, __uint_as_float(v[4 * i4 + 3]), n4.w);
                            } else {
                                // p.xnorm holds 1 / ||x|| here (0 for a zero row, +inf for a pad row): -x.q / ||x||
                                // orders the rows like the cosine distance, whose 1 / ||q|| is applied on output.
                                // A pad row is all zeros, its dot product is exactly 0 and 0 * inf = NaN, which
                                // loses every comparison below (and packs as an empty key): one FMUL per column.
                                c[4 * i4 + 0] = -__uint_as_float(v[4 * i4 + 0]) * n4.x;
                                c[4 * i4 + 1] = -__uint_as_float(v[4 * i4 + 1]) * n4.y;
                                c[4 * i4 + 2] = -__uint_as_float(v[4 * i4 + 2]) * n4.z;
                                c[4 * i4 + 3] = -__uint_as_float(v[4 * i4 + 3]) * n4.w;
                            }
                        }
                    }
                    // minimum of the 32 candidates through 3-input minima (FMNMX3): 16 instructions
                    float m[11];
#pragma unroll
                    for (int i = 0; i < 10; i++) m[i] = fminf(fminf(c[3 * i], c[3 * i + 1]), c[3 * i + 2]);
                    m[10] = fminf(c[30], c[31]);
                    m[0] = fminf(fminf(m[0], m[1]), m[2]);
                    m[3] = fminf(fminf(m[3], m[4]), m[5]);
                    m[6] = fminf(fminf(m[6], m[7]), m[8]);
                    m[9] = fminf(m[9], m[10]);
                    return fminf(fminf(m[0], m[3]), fminf(m[6], m[9]));
                };
                // the rare part: some lane of the warp holds a candidate below its threshold in chunk j
                auto takers = [&](const float (&c)[32], const int j, const float cmin) {
                    const int col0 = half * (TC_N / 2) + j * 32;
                    // columns that may rank: stored rows only (inner product: the padding of a list's last tile is zero
                    // rows with candidate 0; the other metrics give pad rows +inf / NaN candidates), live queries only
                    const uint32_t ncol = tile_rows > (uint32_t) col0 ? tile_rows - (uint32_t) col0 : 0u;
                    const uint32_t okmask = !live ? 0u : (ncol >= 32u ? 0xffffffffu : (1u << ncol) - 1u);
                    changed = true;
                    if constexpr (PACKED) {
                        uint32_t mask = 0;
                        float thc = thr - cadd;                         // threshold in raw candidate units
                        if (cmin < thc) {
#pragma unroll
                            for (int i = 0; i < 32; i++) mask |= (c[i] < thc ? 1u : 0u) << i;
                            mask &= okmask;
                        }
                        const uint32_t ibase = (t - t0) * (TC_N / 2) + j * 32;
#ifdef NDB_TC_COUNTERS
                        dc_any += __any_sync(FULL, mask != 0) ? 1 : 0;
                        dc_takers += __popc(mask);
                        if (__any_sync(FULL, __popc(mask) > p.heavy)) dc_heavy++;
                        else { int mx = __popc(mask); for (int o = 16; o > 0; o >>= 1) mx = max(mx, __shfl_xor_sync(FULL, mx, o)); dc_iters += mx; }
#endif
                        if (__any_sync(FULL, __popc(mask) > p.heavy)) {
                            // some lane has many takers (its list is still filling): every lane sorts
                            // its 32 keys through the network -- a fixed cost, no lane-by-lane tail
                            float g[16];
#pragma unroll
                            for (int i = 0; i < 16; i++) g[i] = (okmask >> i) & 1u ? tc_pack(c[i] + cadd, ibase + i) : FLT_MAX;
                            tc_sort_merge16(bd, g);
#pragma unroll
                            for (int i = 0; i < 16; i++) g[i] = (okmask >> (16 + i)) & 1u ? tc_pack(c[16 + i] + cadd, ibase + 16 + i) : FLT_MAX;
                            tc_sort_merge16(bd, g);
                            thr = fminf(bd[KT - 1], gcap);
                        } else {
                            while (mask) {
                                const int i = __ffs(mask) - 1;
                                mask &= mask - 1;
                                const float cand = pick32(c, i);
                                if (cand < thc) {
                                    tc_key_insert<KT>(bd, tc_pack(cand + cadd, ibase + i));
                                    thr = fminf(bd[KT - 1], gcap);
                                    thc = thr - cadd;
                                }
                            }
                        }
                    } else {
                        if (cmin < thr) {
                            uint32_t mask = 0;
#pragma unroll
                            for (int i = 0; i < 32; i++) mask |= (c[i] < thr ? 1u : 0u) << i;
                            mask &= okmask;
                            while (mask) {
                                const int i = __ffs(mask) - 1;
                                mask &= mask - 1;
                                const float cand = pick32(c, i);
                                if (cand < thr) {
                                    tk_insert<KT>(bd, bi, cand, t * TC_N + col0 + i);
                                    thr = fminf(bd[KT - 1], gcap);
                                }
                            }
                        }
                    }
                };
                // Two chunks per round: both TMEM reads in flight together, two independent min trees, ONE vote that
                // decides the common case (no lane of the warp has a taker in either chunk).  With two epilogue warps
                // per scheduler the loop is bound by its dependent latencies, not by issue slots.
                // (Inner product only: the other metrics keep candidates apart from the accumulator values, and two chunks of
                // both do not fit the register file -- measured slower.)
                constexpr int CPR = METRIC == NDB_IP ? 2 : 1;           // chunks per round
#pragma unroll 1
                for (int j = jlo; j < nchunk; j += CPR) {
                    const bool two = CPR == 2 && j + 1 < nchunk;
                    uint32_t va[32], vb[32];
                    tmem_ld32(acc_addr + j * 32, va);
                    // (an odd last chunk is read twice rather than leaving vb conditionally defined: the compiler then keeps
                    // the candidates in the registers the TMEM read delivered them to instead of copying them)
                    if constexpr (CPR == 2) tmem_ld32(acc_addr + (two ? j + 1 : j) * 32, vb);
                    // the chunk's row norms meanwhile: their shared-memory latency overlaps the TMEM wait
                    float4 na[NN4], nb[NN4];
                    if (METRIC != NDB_IP) {
                        const float *xc = xn + half * (TC_N / 2) + j * 32;
#pragma unroll
                        for (int i4 = 0; i4 < NN4; i4++) na[i4] = *reinterpret_cast<const float4 *>(xc + 4 * i4);
                        if (two) {
#pragma unroll
                            for (int i4 = 0; i4 < NN4; i4++) nb[i4] = *reinterpret_cast<const float4 *>(xc + 32 + 4 * i4);
                        }
                    }
                    tmem_ld_wait();
                    if (p.debug_mode & 8) {                    // bisection: TMEM reads only
                        if (__uint_as_float(va[0]) == 1.2345e-30f) bd[0] = 0.0f;
                        continue;
                    }
                    if (p.debug_d && item == 0 && t == t0) {   // first item's first tile (ndbdbg_tc_gemm)
                        const int col0 = half * (TC_N / 2) + j * 32;
#pragma unroll
                        for (int i = 0; i < 32; i++) p.debug_d[(size_t) ql * TC_N + col0 + i] = __uint_as_float(va[i]);
                        if (two) {
#pragma unroll
                            for (int i = 0; i < 32; i++) p.debug_d[(size_t) ql * TC_N + col0 + 32 + i] = __uint_as_float(vb[i]);
                        }
                    }
#ifdef NDB_TC_COUNTERS
                    dc_chunks += two ? 2 : 1;
#endif
                    float ca[32], cb[32];
                    const float ma = candidates(va, na, ca);
                    float mb = INFINITY;
                    if constexpr (CPR == 2) {
                        mb = candidates(vb, nb, cb);
                        mb = two ? mb : INFINITY;
                    }
                    if constexpr (PACKED) {
                        // (the packed path sorts warp-wide: its lanes enter together)
                        const float lim = thr - cadd;
                        if (!__any_sync(FULL, fminf(ma, mb) < lim)) continue;
                        if (__any_sync(FULL, ma < lim)) takers(ca, j, ma);
                        if (two && __any_sync(FULL, mb < thr - cadd)) takers(cb, j + 1, mb);
                    } else {
                        if (ma < thr) takers(ca, j, ma);               // lane by lane: no vote on the dense kernel's hot path
                        if (two && mb < thr) takers(cb, j + 1, mb);
                    }
                }
                tc_fence_before();
                mbar_arrive(&acc_empty[a]);                             // 256 arrivals release the accumulator
                // the key the bound is taken from: any list that holds kpub entries proves that the query's kpub-th best is
                // at most its kpub-th key (kpub = the caller's k; the list itself keeps KT >= kpub entries)
                // (only a tile that put something into the list can improve the bound)
                float kth = changed ? bd[KT - 1] : published;
                // (a warp-uniform switch: written as a loop of selects the compiler turns it into an indexed load and moves
                // the list to local memory)
#define NDB_KTH(I) case (I) + 1: kth = bd[(I) < KT ? (I) : KT - 1]; break;
                if (__any_sync(FULL, changed))
                switch (p.kpub) {
                    NDB_KTH(0) NDB_KTH(1) NDB_KTH(2) NDB_KTH(3) NDB_KTH(4) NDB_KTH(5) NDB_KTH(6) NDB_KTH(7)
                    NDB_KTH(8) NDB_KTH(9) NDB_KTH(10) NDB_KTH(11) NDB_KTH(12) NDB_KTH(13) NDB_KTH(14)
                    default: break;
                }
#undef NDB_KTH
                if (gcell && p.kpub >= 0 && kth < published) {          // own list full enough and improved
                    published = kth;
                    float pub = published;
                    if (PACKED)     // an upper bound of the value the key stands for
                        pub = __uint_as_float(pub >= 0.0f ? (__float_as_uint(pub) | TC_IDX_MASK) : (__float_as_uint(pub) & ~TC_IDX_MASK));
                    if (!PACKED || pub < TC_KEY_BIG) {
                        if (p.cstats) pub = cert_relax<METRIC>(pub, p.cstats, cq, p.dim);
                        atomic_min_f32(gcell, pub);
                    }
                }
            }
            if (live) {
                const size_t base = ((size_t) it.out_base + (size_t) ql * it.out_stride + half) * p.k;
                float od[KT];
                uint32_t os[KT];
#pragma unroll
                for (int j = 0; j < KT; j++) {
                    {
                        float d = bd[j];
                        uint32_t slot;
                        if (PACKED) {
                            const uint32_t bits = __float_as_uint(d), idx = bits & TC_IDX_MASK;
                            const bool have = d < TC_KEY_BIG;
                            slot = have ? (t0 + (idx >> 7)) * TC_N + half * (TC_N / 2) + (idx & 127u) : INVALID_SLOT;
                            // the key's value: squared distance (L2) / -dot (IP), comparable with p.gthr
                            d = have ? __uint_as_float(bits & ~TC_IDX_MASK) : INFINITY;
                        } else {
                            slot = bi[j];
                            if (METRIC == NDB_L2 && slot != INVALID_SLOT) d = sqrtf(fmaxf(d + qn, 0.0f));
                            if (METRIC == NDB_COSINE && slot != INVALID_SLOT) d = qn > 0.0f ? fmaf(d, 1.0f / sqrtf(qn), 1.0f) : 1.0f;
                        }
                        od[j] = d;
                        os[j] = slot;
                    }
                }
                if (KT % 4 == 0 && p.k == KT) {
                    // a full list is KT * 4 bytes, 16-byte aligned: four 128-bit stores per array (scalar stores reuse
                    // their address registers and wait for one another)
#pragma unroll
                    for (int j = 0; j < KT; j += 4) {
                        *reinterpret_cast<float4 *>(p.pdist + base + j) = make_float4(od[j], od[j + 1], od[j + 2], od[j + 3]);
                        *reinterpret_cast<uint4 *>(p.pslot + base + j) = make_uint4(os[j], os[j + 1], os[j + 2], os[j + 3]);
                    }
                } else {
#pragma unroll
                    for (int j = 0; j < KT; j++)
                        if (j < p.k) { p.pdist[base + j] = od[j]; p.pslot[base + j] = os[j]; }
                }
            }
        }
#ifdef NDB_TC_COUNTERS
        if (p.dbg_counters) {
            // warp-level counts from lane 0, takers from every lane
            unsigned long long *dc = p.dbg_counters + (p.item_lo_ptr ? 8 : 0);      // second phase of a two-phase scan: [8..13)
            if (lane == 0) {
                atomicAdd(dc + 0, (unsigned long long) dc_chunks);
                atomicAdd(dc + 1, (unsigned long long) dc_any);
                atomicAdd(dc + 2, (unsigned long long) dc_heavy);
                atomicAdd(dc + 3, (unsigned long long) dc_iters);
            }
            atomicAdd(dc + 4, (unsigned long long) dc_takers);
        }
#endif
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 2) tmem_dealloc(tmem_base, 512);
}

// ---- host side ---------------------------------------------------------------------------------
static size_t tc_smem_bytes() { return (size_t) TC_MAX_CHUNKS * TC_QCHUNK_BYTES + (size_t) TC_XRING_BYTES + (size_t) TC_NORM_RING * TC_N * 4; }

int tc_build_store(TcStore &st, const float *il32_store, int64_t n, int dim, int dimp, cudaStream_t s)
{
    return tc_build_store_mapped(st, il32_store, nullptr, n, dim, dimp, s);
}

// src_slot_dev (optional, length n rounded up to 256): tensor row -> IL32 slot, INVALID_SLOT = pad
int tc_build_store_mapped(TcStore &st, const float *il32_store, const uint32_t *src_slot_dev, int64_t n, int dim, int dimp,
                          cudaStream_t s)
{
    NDB_REQUIRE(dim <= TC_MAX_DIM, NDB_B200_EINVAL, "tensor path: dim %d > %d is not supported", dim, TC_MAX_DIM);
    const int nkc = (dim + TC_KC - 1) / TC_KC;
    const int64_t ntiles = (n + TC_N - 1) / TC_N, npad = ntiles * TC_N;
    NDB_CHECK(st.xb.reserve((size_t) npad * nkc * TC_KC * 2));
    NDB_CHECK(st.xnorm.reserve((size_t) npad * 4));
    const int groups = nkc * (TC_KC / 8);
    tc_block_rows_kernel<<<(unsigned) ((npad * groups + 255) / 256), 256, 0, s>>>(reinterpret_cast<const float4 *>(il32_store), src_slot_dev, n, npad, dim,
                                                                                 dimp, nkc, st.xb.as<__nv_bfloat16>(), st.xnorm.as<float>());
    NDB_CHECK(st.stats.reserve(16));
    NDB_CUDA(cudaMemsetAsync(st.stats.p, 0, 16, s));
    tc_row_norms_kernel<<<(unsigned) ((npad + 255) / 256), 256, 0, s>>>(reinterpret_cast<const float4 *>(il32_store), src_slot_dev, n, npad, dim, dimp,
                                                                        st.xnorm.as<float>(), st.stats.as<float>());
    count_launch(2);
    NDB_CUDA(cudaGetLastError());
    st.valid_for = n;
    st.rinv_for = -2;
    st.pad0_for = -2;
    st.ntiles = ntiles;
    st.nkc = nkc;
    return NDB_B200_OK;
}

__global__ void tc_rinv_kernel(const float *__restrict__ xnorm, int64_t n, float *__restrict__ out)
{
    const int64_t i = (int64_t) blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float v = xnorm[i];
    out[i] = v == INFINITY ? INFINITY : (v > 0.0f ? 1.0f / sqrtf(v) : 0.0f);
}

__global__ void tc_pad0_kernel(const float *__restrict__ xnorm, int64_t n, float *__restrict__ out)
{
    const int64_t i = (int64_t) blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = xnorm[i] == INFINITY ? INFINITY : 0.0f;
}

// per-row additive term of the inner-product metric: 0 for a stored row, +inf for a pad row
int tc_store_pad0(TcStore &st, const float **out, cudaStream_t s)
{
    const int64_t n = st.ntiles * TC_N;
    if (st.pad0_for != st.valid_for || st.xpad0.cap < (size_t) n * 4) {
        NDB_CHECK(st.xpad0.reserve((size_t) n * 4));
        tc_pad0_kernel<<<(unsigned) ((n + 255) / 256), 256, 0, s>>>(st.xnorm.as<float>(), n, st.xpad0.as<float>());
        count_launch();
        NDB_CUDA(cudaGetLastError());
        st.pad0_for = st.valid_for;
    }
    *out = st.xpad0.as<float>();
    return NDB_B200_OK;
}

// per-row 1 / ||x|| for the cosine metric (built on first use, kept with the store)
int tc_store_rinv(TcStore &st, const float **out, cudaStream_t s)
{
    const int64_t n = st.ntiles * TC_N;
    if (st.rinv_for != st.valid_for || st.xrinv.cap < (size_t) n * 4) {
        NDB_CHECK(st.xrinv.reserve((size_t) n * 4));
        tc_rinv_kernel<<<(unsigned) ((n + 255) / 256), 256, 0, s>>>(st.xnorm.as<float>(), n, st.xrinv.as<float>());
        count_launch();
        NDB_CUDA(cudaGetLastError());
        st.rinv_for = st.valid_for;
    }
    *out = st.xrinv.as<float>();
    return NDB_B200_OK;
}

// blocked bf16 query tiles + squared norms; qmap_dev (optional) gathers the tile positions
int tc_block_queries(const float *Q_dev, const uint32_t *qmap_dev, uint32_t nprobe, int nq, int nqpad, int dim, int nkc,
                     __nv_bfloat16 *qb, float *qnorm, cudaStream_t s, const uint32_t *npos_dev, float *qerr, bool negate)
{
    tc_block_queries_kernel<<<(unsigned) ((nqpad / TC_M) * nkc), 256, 0, s>>>(Q_dev, qmap_dev, nprobe, nq, nqpad, npos_dev, dim, nkc, qb, qnorm, qerr,
                                                                              negate);
    count_launch();
    if (nkc > 1) {
        tc_query_norms_kernel<<<(unsigned) ((nqpad + 127) / 128), 128, 0, s>>>(Q_dev, qmap_dev, nprobe, nq, nqpad, npos_dev, dim, qnorm, qerr);
        count_launch();
    }
    NDB_CUDA(cudaGetLastError());
    return NDB_B200_OK;
}

#ifdef NDB_TC_COUNTERS
static unsigned long long *g_tc_dbg_counters = nullptr;
#endif
// launch the persistent kernel over p.items (grid = min(items, SMs)); records timing events
int tc_launch(const TcParams &p, int metric, int k, cudaStream_t s)
{
    NDB_REQUIRE(k >= 1 && k <= TC_KMAX, NDB_B200_EINVAL, "tensor path: k must be 1..%d", TC_KMAX);
    NDB_REQUIRE(metric == NDB_L2 || metric == NDB_IP || metric == NDB_COSINE, NDB_B200_EINVAL, "tensor path: metric %d not supported", metric);
    if (p.nitems == 0) return NDB_B200_OK;
    const size_t smem = tc_smem_bytes();
    const uint32_t sms = (uint32_t) ctx().sm_count;
    const uint32_t grid = (p.nitems < sms && !p.nitems_ptr) ? p.nitems : sms;
    Context &c = ctx();
    {   // 8-element K groups of the last chunk that carry data, rounded up to whole MMAs (K = 16)
        TcParams &pm = const_cast<TcParams &>(p);
        const int rest = p.dim - (p.nkc - 1) * TC_KC;
        pm.kg_last = (p.dim > 0 && rest > 0 && rest <= TC_KC && !getenv("NDB_TC_FULL_K")) ? ((rest + 15) / 16) * 2 : TC_KC / 8;
        const char *se = getenv("NDB_TC_STAGES"), *he = getenv("NDB_TC_HEAVY");            // measurement switches
        pm.heavy = he ? atoi(he) : TC_HEAVY;
        if (p.nkc == 1) {
            pm.xstride = pm.kg_last * (TC_N / 8) * 128;
            pm.nstages = std::min(TC_MAX_STAGES, TC_XRING_BYTES / pm.xstride);
            if (se && atoi(se) >= 1 && atoi(se) < pm.nstages) pm.nstages = atoi(se);
        } else {
            pm.xstride = TC_XSTAGE_BYTES;
            pm.nstages = 2;
        }
    }
#ifdef NDB_TC_COUNTERS
    static DevBuf dbgc;
    if (!dbgc.p) { NDB_CHECK(dbgc.reserve(128)); NDB_CUDA(cudaMemsetAsync(dbgc.p, 0, 128, s)); }
    const_cast<TcParams &>(p).dbg_counters = p.gthr ? dbgc.as<unsigned long long>() : nullptr;      // list mode only
    g_tc_dbg_counters = dbgc.as<unsigned long long>();
#endif
    if (c.timing) NDB_CUDA(cudaEventRecord(c.ev0, s));
    // list length: the smallest of {1, 10, 16} that holds k (a shorter list = a tighter threshold)
#define NDB_TC_LAUNCH(KT, M, PK)                                                                              \
    do {                                                                                                      \
        static uint64_t cfg = 0;                                                                              \
        if (cfg != ctx().generation) {                                                                        \
            NDB_CUDA(cudaFuncSetAttribute(tc_knn_kernel<KT, M, PK>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) smem)); \
            cfg = ctx().generation;                                                                           \
        }                                                                                                     \
        tc_knn_kernel<KT, M, PK><<<grid, 384, smem, s>>>(p);                                                  \
    } while (0)
#define NDB_TC_PICK(M, PK)                                                                                    \
    do {                                                                                                      \
        if (k == 1) NDB_TC_LAUNCH(1, M, PK);                                                                  \
        else if (k <= 10) NDB_TC_LAUNCH(10, M, PK);                                                           \
        else NDB_TC_LAUNCH(TC_KMAX, M, PK);                                                                   \
    } while (0)
    // packed: every item spans at most TC_PACKED_MAX_TILES tiles (the caller's contract)
    if (metric == NDB_L2) {
        if (p.packed) NDB_TC_PICK(NDB_L2, true);
        else NDB_TC_PICK(NDB_L2, false);
    } else if (metric == NDB_COSINE) {
        if (p.packed) NDB_TC_PICK(NDB_COSINE, true);
        else NDB_TC_PICK(NDB_COSINE, false);
    } else {
        if (p.packed) NDB_TC_PICK(NDB_IP, true);
        else NDB_TC_PICK(NDB_IP, false);
    }
#undef NDB_TC_PICK
#undef NDB_TC_LAUNCH
    count_launch();
    NDB_CUDA(cudaGetLastError());
    if (c.timing) {
        NDB_CUDA(cudaEventRecord(c.ev1, s));
        c.last_ms = -1.0;
        c.last_bytes = 0.0;
        c.last_evals = 0;
        c.stats_src = nullptr;
    }
    return NDB_B200_OK;
}

// top-k of nq row-major fp32 queries against a TcStore; writes (dist, id) like the scan path
int tc_knn(const TcStore &st, TcScratch &sc, int dim, int metric, const float *Q_dev, int nq, int k, const int64_t *ids,
           float *dist_dev, int64_t *ids_dev, uint32_t *slots_dev, float *debug_d_dev, bool packed, cudaStream_t s)
{
    NDB_REQUIRE(k >= 1 && k <= TC_KMAX, NDB_B200_EINVAL, "tensor path: k must be 1..%d", TC_KMAX);
    NDB_REQUIRE(metric == NDB_L2 || metric == NDB_IP || metric == NDB_COSINE, NDB_B200_EINVAL, "tensor path: metric %d not supported", metric);
    NDB_REQUIRE(st.ntiles > 0, NDB_B200_ESTATE, "tensor path: empty store");
    const int nkc = st.nkc;
    const uint32_t nqt = (uint32_t) ((nq + TC_M - 1) / TC_M);
    const int nqpad = (int) nqt * TC_M;
    NDB_CHECK(sc.qb.reserve((size_t) nqpad * nkc * TC_KC * 2));
    NDB_CHECK(sc.qnorm.reserve((size_t) nqpad * 4));
    NDB_CHECK(tc_block_queries(Q_dev, nullptr, 0, nq, nqpad, dim, nkc, sc.qb.as<__nv_bfloat16>(), sc.qnorm.as<float>(), s, nullptr, nullptr,
                               metric == NDB_IP));
    // Split the stored tiles into ranges so that every SM gets ~8 work items (of at least 32 tiles): the persistent CTAs
    // take whole items, and with ~2 per SM the last round ran a fraction of the grid (C5: 316 items on 148 SMs).  The
    // items of a query share an upper bound of its k-th best candidate (p.gthr, as in list mode), so a range does not
    // pay for filling its lists from nothing.
    const uint32_t sms = (uint32_t) ctx().sm_count;
    static const uint32_t per_sm = [] { const char *e = getenv("NDB_TC_ITEMS_PER_SM"); int x = e ? atoi(e) : 8; return (uint32_t) (x >= 1 ? x : 8); }();
    uint32_t nranges = (per_sm * sms + nqt - 1) / nqt;
    if (nranges > (uint32_t) (st.ntiles / 32)) nranges = (uint32_t) (st.ntiles / 32);
    if (nranges < 1) nranges = 1;
    if (nranges > (uint32_t) st.ntiles) nranges = (uint32_t) st.ntiles;
    uint32_t tpr = (uint32_t) ((st.ntiles + nranges - 1) / nranges);
    if (packed) {
        // short stores (the IVF centroids): an item's first tile pays for filling the top-k lists, so
        // prefer one item per query tile over many one-tile items as soon as half the SMs have work
        if (tpr > (uint32_t) TC_PACKED_MAX_TILES) tpr = TC_PACKED_MAX_TILES;
        else if (2 * nqt >= sms && st.ntiles <= TC_PACKED_MAX_TILES) tpr = (uint32_t) st.ntiles;
    }
    nranges = (uint32_t) ((st.ntiles + tpr - 1) / tpr);
    NDB_CHECK(sc.pdist.reserve((size_t) nq * nranges * 2 * k * 4));
    NDB_CHECK(sc.pslot.reserve((size_t) nq * nranges * 2 * k * 4));
    // dense work items: (query tile, range of stored tiles); partial slot of (query q, part) is
    // q * nparts + part with nparts = 2 * nranges (two column halves per range)
    const uint32_t nitems = nqt * nranges;
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
    NDB_CHECK(sc.items.reserve((size_t) nitems * sizeof(TcItem)));
    NDB_CUDA(cudaMemcpyAsync(sc.items.p, items.data(), (size_t) nitems * sizeof(TcItem), cudaMemcpyHostToDevice, s));
    NDB_CUDA(cudaStreamSynchronize(s));        // `items` is a host temporary
    TcParams p;
    memset(&p, 0, sizeof(p));
    p.xb = st.xb.as<__nv_bfloat16>();
    p.xnorm = st.xnorm.as<float>();
    if (metric == NDB_COSINE) NDB_CHECK(tc_store_rinv(const_cast<TcStore &>(st), &p.xnorm, s));
    if (metric == NDB_IP) NDB_CHECK(tc_store_pad0(const_cast<TcStore &>(st), &p.xnorm, s));
    p.qb = sc.qb.as<__nv_bfloat16>();
    p.qnorm = sc.qnorm.as<float>();
    p.nkc = nkc; p.k = k;
    p.dim = dim;
    p.items = sc.items.as<TcItem>();
    p.nitems = nitems;
    p.pdist = sc.pdist.as<float>();
    p.pslot = sc.pslot.as<uint32_t>();
    p.packed = packed ? 1 : 0;
    if (!packed && !getenv("NDB_TC_DENSE_NOSHARE")) {
        NDB_CHECK(sc.gthr.reserve((size_t) nqpad * 4));
        NDB_CUDA(cudaMemsetAsync(sc.gthr.p, 0x7f, (size_t) nqpad * 4, s));       // 3.4e38: "no bound yet"
        p.gthr = sc.gthr.as<float>();
    }
    p.debug_d = debug_d_dev;
    p.debug_mode = getenv("NDB_TC_DEBUG") ? atoi(getenv("NDB_TC_DEBUG")) : 0;
    const size_t smem = tc_smem_bytes();
    NDB_CHECK(tc_launch(p, metric, k, s));
    if (ctx().timing) {
        ctx().last_bytes = (double) st.ntiles * TC_N * nkc * TC_KC * 2.0;          // stored bf16 bytes, read once per launch
        ctx().last_evals = (int64_t) st.valid_for * nq;
    }
    return launch_merge_parts(sc.pdist.as<float>(), sc.pslot.as<uint32_t>(), ids, nq, (int) nranges * 2, k, dist_dev, ids_dev, slots_dev, s);
}

}  // namespace ndb

#ifdef NDB_TC_COUNTERS
// development hook: read and reset the epilogue statistics of the list-mode launches since the last call
extern "C" int ndbdbg_tc_counters(unsigned long long *out /* [16] */)
{
    using namespace ndb;
    NDB_CHECK(require_init());
    for (int i = 0; i < 16; i++) out[i] = 0;
    if (!g_tc_dbg_counters) return NDB_B200_OK;
    NDB_CUDA(cudaDeviceSynchronize());
    NDB_CUDA(cudaMemcpy(out, g_tc_dbg_counters, 128, cudaMemcpyDeviceToHost));
    NDB_CUDA(cudaMemset(g_tc_dbg_counters, 0, 128));
    return NDB_B200_OK;
}
#endif
// development hook (not part of the ABI in include/ndb_b200.h): raw accumulator tile D = Q X^T of the
// first (query tile, row tile) -- used by tests/test_gpu_tensor.py to validate the UMMA descriptors
extern "C" int ndbdbg_tc_gemm(const float *Q, int nq, const float *X, int n, int dim, float *D /* [128][256] */)
{
    using namespace ndb;
    NDB_CHECK(require_init());
    NDB_REQUIRE(Q && X && D && nq >= 1 && nq <= TC_M && n >= 1 && n <= TC_N && dim >= 1, NDB_B200_EINVAL, "dbg_tc_gemm: bad shape");
    cudaStream_t s = ctx().stream;
    const int dimp = round_up(dim, 4);
    DevBuf dq, dx, store, dd, od, oi;
    TcStore st;
    TcScratch sc;
    NDB_CHECK(dq.reserve((size_t) nq * dim * 4)); NDB_CHECK(dx.reserve((size_t) n * dim * 4));
    const int64_t blocks = (n + 31) / 32;
    NDB_CHECK(store.reserve((size_t) blocks * 32 * dimp * 4));
    NDB_CHECK(dd.reserve((size_t) TC_M * TC_N * 4)); NDB_CHECK(od.reserve((size_t) nq * 4)); NDB_CHECK(oi.reserve((size_t) nq * 8));
    NDB_CUDA(cudaMemcpyAsync(dq.p, Q, (size_t) nq * dim * 4, cudaMemcpyHostToDevice, s));
    NDB_CUDA(cudaMemcpyAsync(dx.p, X, (size_t) n * dim * 4, cudaMemcpyHostToDevice, s));
    NDB_CUDA(cudaMemsetAsync(store.p, 0, (size_t) blocks * 32 * dimp * 4, s));
    NDB_CUDA(cudaMemsetAsync(dd.p, 0, (size_t) TC_M * TC_N * 4, s));
    NDB_CHECK(il32_scatter(dx.as<float>(), n, dim, dimp, nullptr, 0, store.as<float>(), s));
    NDB_CHECK(tc_build_store(st, store.as<float>(), n, dim, dimp, s));
    NDB_CHECK(tc_knn(st, sc, dim, NDB_L2, dq.as<float>(), nq, 1, nullptr, od.as<float>(), oi.as<int64_t>(), nullptr, dd.as<float>(), false, s));
    NDB_CUDA(cudaMemcpyAsync(D, dd.p, (size_t) TC_M * TC_N * 4, cudaMemcpyDeviceToHost, s));
    NDB_CUDA(cudaStreamSynchronize(s));
    return NDB_B200_OK;
}
