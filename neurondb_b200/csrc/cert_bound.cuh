// cert_bound.cuh -- the rounding-error bounds of the certified tensor-core selection (ivf_cert.cuh), shared by
// the scan epilogue (tc_knn.cu: how far the shared per-query bound must be relaxed) and the finish kernels.
//
// Keys are computed from x~ = bf16(x), q~ = bf16(q); ex = x~ - x, eq = q~ - q; bf16 x bf16 products are exact in
// fp32, so the only approximations are the roundings of the inputs, the fp32 accumulation of `dim` terms
// (<= (dim + 8) 2^-23 of the magnitudes) and the 12-bit packing of the keys (a key differs from the value it stands
// for by less than 2^-12 relative; 2^-11 is charged, in both directions):
//     L2      key = ||x~ - q~||^2          | ||x - q|| - sqrt(key) |           <= ||ex|| + ||eq||
//     IP      key = -x~.q~                 | -x.q - key |                      <= ||ex|| ||q~|| + ||x|| ||eq||
//     cosine  key = -x~.q~ / ||x~||        | -x.q / ||x|| - key |              <= rho (||q~|| + ||q||) + kappa ||eq||
// st[] = maxima over the stored rows (TcStore::stats): ||ex||^2, ||x||^2, rho^2 = (||ex|| / ||x~||)^2,
// kappa^2 = (||x|| / ||x~||)^2.
#pragma once
#include "common.cuh"

namespace ndb {

// norms of a query and of its bf16 rounding error, each rounded up a little (they feed upper bounds)
struct CertQ { float eq, qn, qnr, qn_lo; };

// additive slack E of a key (IP, cosine) or of its square root (L2), and the accumulation slack D in key units
template <int METRIC>
__device__ __forceinline__ void cert_slack(const float *__restrict__ st, const CertQ &c, int dim, float &E, float &D, float &gam)
{
    const float exm = sqrtf(st[0]) * 1.0002f, xm = sqrtf(st[1]) * 1.0002f;
    gam = (float) (dim + 8) * 1.1920929e-7f;                        // (dim + 8) * 2^-23
    const float xr = xm + exm;                                      // >= ||bf16(x)||
    if (METRIC == NDB_L2) {
        E = exm + c.eq;
        D = gam * (xr + c.qnr) * (xr + c.qnr);
    } else if (METRIC == NDB_IP) {
        E = exm * c.qnr + xm * c.eq + gam * xm * c.qn;
        D = gam * xr * c.qnr;
    } else {
        const float rho = sqrtf(st[2]) * 1.0002f, kap = sqrtf(st[3]) * 1.0002f;
        E = rho * (c.qnr + c.qn) + kap * c.eq;
        D = gam * c.qnr;
    }
}

// a lower bound of the reference-arithmetic value of every row whose key compared >= `key`
template <int METRIC>
__device__ __forceinline__ float cert_lower_bound(float key, const float *__restrict__ st, const CertQ &c, int dim)
{
    float E, D, gam;
    cert_slack<METRIC>(st, c, dim, E, D, gam);
    const float kv = key - fabsf(key) * 4.8828125e-4f - D;          // 2^-11: keys carry 12 mantissa bits
    if (METRIC == NDB_L2) {
        const float r = sqrtf(fmaxf(kv, 0.0f)) - E;
        return r - fabsf(r) * gam;
    }
    if (METRIC == NDB_IP) return kv - E;
    if (!(c.qn_lo > 0.0f)) return -INFINITY;                        // zero query: every distance is 1.0f, ties by id
    const float u = kv - E;
    return 1.0f + (u < 0.0f ? u / c.qn_lo : u / c.qn) - 8.0f * gam;
}

// an upper bound of the reference-arithmetic value of a row whose key compared <= `key`
template <int METRIC>
__device__ __forceinline__ float cert_upper_bound(float key, const float *__restrict__ st, const CertQ &c, int dim)
{
    float E, D, gam;
    cert_slack<METRIC>(st, c, dim, E, D, gam);
    const float kv = key + fabsf(key) * 4.8828125e-4f + D;
    if (METRIC == NDB_L2) {
        const float r = sqrtf(fmaxf(kv, 0.0f)) + E;
        return r + fabsf(r) * gam;
    }
    if (METRIC == NDB_IP) return kv + E;
    if (!(c.qn_lo > 0.0f)) return INFINITY;
    const float u = kv + E;
    return 1.0f + (u > 0.0f ? u / c.qn_lo : u / c.qn) + 8.0f * gam;
}

// The relaxed bound: the smallest key R (rounded up generously) with cert_lower_bound(R) > cert_upper_bound(key).
// A partial list that ends with `key` as its kc-th entry proves that the final k-th exact value is at most
// cert_upper_bound(key); rows whose key is >= R can then never enter the exact top k.  The scan publishes R, not
// `key`, as the shared per-query bound, so that every NON-full partial list is complete below R.
template <int METRIC>
__device__ __forceinline__ float cert_relax(float key, const float *__restrict__ st, const CertQ &c, int dim)
{
    float E, D, gam;
    cert_slack<METRIC>(st, c, dim, E, D, gam);
    const float ub = cert_upper_bound<METRIC>(key, st, c, dim);
    float R;
    if (METRIC == NDB_L2) {
        const float t = (ub + fabsf(ub) * 2.0f * gam) + E;          // sqrt(kv) must reach this
        R = t * t + D;
    } else if (METRIC == NDB_IP) {
        R = ub + E + D;
    } else {
        if (!(c.qn_lo > 0.0f)) return INFINITY;
        const float v = ub - 1.0f + 16.0f * gam;                    // (u / ||q||) must reach this
        const float u = v > 0.0f ? v * c.qn : v * c.qn_lo;
        R = u + E + D;
    }
    R += fabsf(R) * 9.765625e-4f + 1e-30f;                           // 2^-10: the lower bound's own packing slack (2^-11) + strictness
    return R;
}

}  // namespace ndb
