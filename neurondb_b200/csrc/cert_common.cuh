// cert_common.cuh -- device helpers of the certified tensor-core selection (see ivf_cert.cuh for the scheme) and
// the certified nearest-centroid kernels shared by the IVF coarse quantiser, ivfinsert's list assignment and the
// k-means assignment step.
#pragma once
#include "arith.cuh"
#include "tc.cuh"
#include "cert_bound.cuh"

namespace ndb {

__device__ __forceinline__ CertQ cert_query(const float *__restrict__ qv, int dim, int lane)
{
    float e2 = 0.0f, q2 = 0.0f, r2 = 0.0f;
    for (int j = lane; j < dim; j += 32) {
        const float v = qv[j], r = __bfloat162float(__float2bfloat16_rn(v));
        e2 = fmaf(v - r, v - r, e2);
        q2 = fmaf(v, v, q2);
        r2 = fmaf(r, r, r2);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        e2 += __shfl_xor_sync(FULL, e2, o);
        q2 += __shfl_xor_sync(FULL, q2, o);
        r2 += __shfl_xor_sync(FULL, r2, o);
    }
    CertQ c;
    c.eq = sqrtf(e2) * 1.0002f;
    c.qn = sqrtf(q2) * 1.0002f;
    c.qnr = sqrtf(r2) * 1.0002f;
    c.qn_lo = sqrtf(q2) * 0.9998f;
    return c;
}

// the reference's distance between query qv and the row-major row xv (policy P = Arith<metric, IVF_F32>)
template <class P>
__device__ __forceinline__ float cert_exact(const float *__restrict__ qv, const float *__restrict__ xv, int dim)
{
    typename P::Acc acc;
    typename P::N vn = 0, qn = 0;
    P::init(acc);
    if ((dim & 3) == 0) {
#pragma unroll 8
        for (int j = 0; j < dim; j += 4) {
            const float4 x = *reinterpret_cast<const float4 *>(xv + j);
            const float4 qq = *reinterpret_cast<const float4 *>(qv + j);       // (rows and queries are 16-byte aligned: dim % 4 == 0)
            P::step(acc, x.x, qq.x);
            P::step(acc, x.y, qq.y);
            P::step(acc, x.z, qq.z);
            P::step(acc, x.w, qq.w);
            if (P::NORMS) {
                P::nstep(vn, x.x); P::nstep(vn, x.y); P::nstep(vn, x.z); P::nstep(vn, x.w);
                P::nstep(qn, qq.x); P::nstep(qn, qq.y); P::nstep(qn, qq.z); P::nstep(qn, qq.w);
            }
        }
    } else {
        for (int j = 0; j < dim; j++) {
            P::step(acc, xv[j], qv[j]);
            if (P::NORMS) { P::nstep(vn, xv[j]); P::nstep(qn, qv[j]); }
        }
    }
    return P::finish(acc, vn, qn);
}

// the same for IL32 slot `slot` of a list store (lane-per-slot: a warp's loads are 512 contiguous bytes)
template <class P>
__device__ __forceinline__ float cert_exact_il32(const float *__restrict__ qv, const float4 *__restrict__ vecs, uint32_t slot, int dim,
                                                 int dimp)
{
    const float4 *vp = vecs + (size_t) (slot >> 5) * (8 * (size_t) dimp) + (slot & 31);
    typename P::Acc acc;
    typename P::N vn = 0, qn = 0;
    P::init(acc);
#pragma unroll 8
    for (int j = 0; j < dim; j += 4) {
        const float4 x = vp[(size_t) (j >> 2) * 32];
        P::step(acc, x.x, qv[j]);
        if (P::NORMS) { P::nstep(vn, x.x); P::nstep(qn, qv[j]); }
        if (j + 1 < dim) { P::step(acc, x.y, qv[j + 1]); if (P::NORMS) { P::nstep(vn, x.y); P::nstep(qn, qv[j + 1]); } }
        if (j + 2 < dim) { P::step(acc, x.z, qv[j + 2]); if (P::NORMS) { P::nstep(vn, x.z); P::nstep(qn, qv[j + 2]); } }
        if (j + 3 < dim) { P::step(acc, x.w, qv[j + 3]); if (P::NORMS) { P::nstep(vn, x.w); P::nstep(qn, qv[j + 3]); } }
    }
    return P::finish(acc, vn, qn);
}

// the certificate compares distances; k-means' policy ranks by the SQUARED distance (ivf_am.c:2255-2269)
template <class P> struct CertTau { __device__ static float f(float t) { return t; } };
template <> struct CertTau<Arith<METRIC_L2SQ, NDB_ARITH_IVF_F32>> { __device__ static float f(float t) { return sqrtf(t) * 1.000002f; } };

// Re-evaluate the (up to 32 * KRC) candidates `cand` holds in ascending key order (entry e = register e / 32,
// lane e % 32; key = row-major row index, INVALID_SLOT = none), 16 at a time, best first, and certify.  (Measured on C2:
// 16 at a time re-evaluates 25 rows per query on average and is faster than 32 at a time, which always fetches 32.)
// `g_rest` = the smallest key any row outside `cand` can have; `complete` = there is no such row.
// Returns the exact top-k in `top` and whether it is certified.
template <class P, int METRIC, int KRC, class RowOf, class KeyOf>
__device__ __forceinline__ bool cert_rerank(const float *__restrict__ qv, const float *__restrict__ rows, int dim,
                                            const WarpTopK<KRC, uint32_t> &cand, float g_rest, bool complete,
                                            const float *__restrict__ st, const CertQ &cq, int k, int lane, RowOf row_of, KeyOf key_of,
                                            WarpTopK<1, int64_t> &top, unsigned long long *__restrict__ n_exact)
{
    top.init();
    bool certified = false, done = false;
#pragma unroll
    for (int c = 0; c < 2 * KRC; c++) {
        if (done) break;
        const int r = c >> 1, half = c & 1;
        const uint32_t ck = cand.key[r];
        const bool mine = ck != INVALID_SLOT && (lane >> 4) == half;
        float ed = INFINITY;
        int64_t id = -1;
        if (mine) { ed = cert_exact<P>(qv, rows + (size_t) row_of(ck) * dim, dim); id = key_of(ck); }
        const unsigned m = __ballot_sync(FULL, mine);
        top.offer(ed, id, mine, lane, k);
        if (n_exact && lane == 0 && m) atomicAdd(n_exact, (unsigned long long) __popc(m));
        // smallest key among the rows not yet re-evaluated: entry 16 (c + 1) of cand, else g_rest
        float g = g_rest;
        bool more = false;
        if (c + 1 < 2 * KRC) {
            const int r2 = (c + 1) >> 1, l2 = ((c + 1) & 1) * 16;
            const float kn = __shfl_sync(FULL, cand.d[r2], l2);
            const uint32_t sn = __shfl_sync(FULL, cand.key[r2], l2);
            more = sn != INVALID_SLOT;
            if (more) g = fminf(kn, g_rest);
        }
        certified = (complete && !more) || cert_lower_bound<METRIC>(g, st, cq, dim) > CertTau<P>::f(top.td);
        done = certified || !more;
    }
    return certified;
}

// ---- coarse quantiser: ivfSelectClusters (:1597-1717) certified the same way -----------------------------
// pdist / pslot: per query `nparts` partial lists of kc (squared-distance key, centroid) entries from the tensor
// scan of the centroid store (no shared bound there: G = the smallest kc-th key of a full list).  Writes the
// nprobe nearest centroids by (fp32 L2 of policy P -- sqrtf'd for ivfSelectClusters / ivfinsert, squared for k-means --,
// index): the reference's order, repeated scan with strict <.
template <int KRC, class P>
__global__ void __launch_bounds__(128) ivf_coarse_cert_kernel(
    const float *__restrict__ pdist, const uint32_t *__restrict__ pslot, int nparts, int kc, const float *__restrict__ C,
    const float *__restrict__ Q, int nq, int nlists, int dim, int np, const float *__restrict__ stats,
    uint32_t *__restrict__ probe, float *__restrict__ cdist, uint32_t *__restrict__ fb_list,
    unsigned long long *__restrict__ counters)
{
    const int lane = threadIdx.x & 31;
    const int q = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (q >= nq) return;
    const size_t base = (size_t) q * nparts * kc;
    // G = min over the full partial lists of their last key
    float G = INFINITY;
    for (int p = lane; p < nparts; p += 32)
        if (pslot[base + (size_t) p * kc + kc - 1] != INVALID_SLOT) G = fminf(G, pdist[base + (size_t) p * kc + kc - 1]);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) G = fminf(G, __shfl_xor_sync(FULL, G, o));
    WarpTopK<KRC, uint32_t> cand;
    cand.init();
    uint32_t n_in = 0;
    const int total = nparts * kc;
    for (int i0 = 0; i0 < total; i0 += 32) {
        const int i = i0 + lane;
        float d = INFINITY;
        uint32_t sl = INVALID_SLOT;
        if (i < total) { sl = pslot[base + i]; d = pdist[base + i]; }
        const bool ok = sl != INVALID_SLOT && sl < (uint32_t) nlists && d <= G;
        const unsigned m = __ballot_sync(FULL, ok);
        if (m) { n_in += __popc(m); cand.offer(d, sl, ok, lane, 32 * KRC); }
    }
    const float *qv = Q + (size_t) q * dim;
    const CertQ cq = cert_query(qv, dim, lane);
    const float keylast = __shfl_sync(FULL, cand.d[KRC - 1], 31);
    const float g_rest = n_in > 32 * KRC ? keylast : G;
    const bool complete = G == INFINITY && n_in <= 32 * KRC;
    WarpTopK<1, int64_t> top;
    const bool ok = cert_rerank<P, NDB_L2, KRC>(qv, C, dim, cand, g_rest, complete, stats, cq, np, lane,
                                                                                [&](uint32_t c) { return c; }, [&](uint32_t c) { return (int64_t) c; },
                                                                                top, counters ? counters + 1 : nullptr);
    if (!ok) {
        if (lane == 0) fb_list[atomicAdd(counters, 1ull)] = (uint32_t) q;
        return;
    }
    if (lane < np) {
        const bool got = top.key[0] != KeyMax<int64_t>::v;
        probe[(size_t) q * np + lane] = got ? (uint32_t) top.key[0] : INVALID_SLOT;
        cdist[(size_t) q * np + lane] = got ? top.d[0] : INFINITY;
    }
}

// exact ivfSelectClusters for the queries the certificate rejected: one CTA per query over all centroids
template <class P>
__global__ void __launch_bounds__(256) ivf_coarse_fallback_kernel(
    const uint32_t *__restrict__ fb_list, const unsigned long long *__restrict__ counters, const float *__restrict__ C,
    const float *__restrict__ Q, int nlists, int dim, int np, uint32_t *__restrict__ probe, float *__restrict__ cdist)
{
    extern __shared__ float fb_smem[];
    float *qs = fb_smem;
    float *md = fb_smem + ((dim + 3) & ~3);
    int64_t *mi = reinterpret_cast<int64_t *>(md + 256);
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const unsigned nfb = (unsigned) counters[0];
    for (unsigned f = blockIdx.x; f < nfb; f += gridDim.x) {
        const uint32_t q = fb_list[f];
        __syncthreads();
        for (int j = threadIdx.x; j < dim; j += blockDim.x) qs[j] = Q[(size_t) q * dim + j];
        __syncthreads();
        WarpTopK<1, int64_t> top;
        top.init();
        for (int j0 = w * 32; j0 < nlists; j0 += 256) {
            const int j = j0 + lane;
            const bool valid = j < nlists;
            const float ed = valid ? cert_exact<P>(qs, C + (size_t) j * dim, dim) : INFINITY;
            top.offer(ed, (int64_t) j, valid, lane, np);
        }
        md[w * 32 + lane] = top.d[0];
        mi[w * 32 + lane] = top.key[0];
        __syncthreads();
        if (w == 0) {
            WarpTopK<1, int64_t> fin;
            fin.init();
            for (int ww = 0; ww < 8; ww++) {
                const float d = md[ww * 32 + lane];
                const int64_t id = mi[ww * 32 + lane];
                fin.offer(d, id, id != KeyMax<int64_t>::v, lane, np);
            }
            if (lane < np) {
                const bool got = fin.key[0] != KeyMax<int64_t>::v;
                probe[(size_t) q * np + lane] = got ? (uint32_t) fin.key[0] : INVALID_SLOT;
                cdist[(size_t) q * np + lane] = got ? fin.d[0] : INFINITY;
            }
        }
    }
}

}  // namespace ndb
