// arith.cuh -- the reference's arithmetic variants as device policies.
//
// Each policy reproduces one CPU loop of the reference bit for bit: same operand types,
// same operation order, explicit round-to-nearest intrinsics so that nvcc can never
// contract a multiply and an add into an FMA where the reference (built without -mfma,
// NeuronDB/build.sh:712) rounds twice.  Where an FMA is used (fp64 accumulation of an
// exactly representable product) the result is identical by construction; each such
// place says why.
//
//   Arith<M, NDB_ARITH_OP_F64>   src/vector/vector_distance.c:93-122,145-157,180-213
//   Arith<M, NDB_ARITH_IVF_F32>  src/index/ivf_am.c:1550-1592 (and :906-935, :2255-2269)
//   Arith<M, NDB_ARITH_HNSW>     src/index/hnsw_am.c:1301-1345
//   Arith<M, NDB_ARITH_FAST>     fp32 FFMA (tolerance path, 1e-5 relative)
//
// METRIC_L2SQ (0) is k-means' squared L2 without the sqrt (ivf_am.c:2255-2269).
#pragma once
#include "common.cuh"

namespace ndb {

constexpr int METRIC_L2SQ = 0;

template <int METRIC, int ARITH> struct Arith;

// ---------------------------------------------------------------------------------------
// ivfComputeDistance: f32 sequential.  L2: sum += (a-b)*(a-b); sqrtf.
// ---------------------------------------------------------------------------------------
template <> struct Arith<NDB_L2, NDB_ARITH_IVF_F32> {
    using Q = float; using N = float; using Acc = float;
    static constexpr bool NORMS = false;
    __device__ static void init(Acc &a) { a = 0.0f; }
    __device__ static void step(Acc &a, float x, Q q)
    {
        float diff = __fsub_rn(q, x);                 // vec1 = query, vec2 = entry
        a = __fadd_rn(a, __fmul_rn(diff, diff));
    }
    __device__ static float finish(const Acc &a, N, N) { return __fsqrt_rn(a); }
    __device__ static void nstep(N &, float) {}
};
template <> struct Arith<METRIC_L2SQ, NDB_ARITH_IVF_F32> : Arith<NDB_L2, NDB_ARITH_IVF_F32> {
    __device__ static float finish(const Acc &a, N, N) { return a; }
};
// cosine: dot, norm1, norm2 are three independent f32 running sums; the two norms depend on
// one vector only, so they are accumulated once per vector (same order => same bits).
template <> struct Arith<NDB_COSINE, NDB_ARITH_IVF_F32> {
    using Q = float; using N = float; using Acc = float;
    static constexpr bool NORMS = true;
    __device__ static void init(Acc &a) { a = 0.0f; }
    __device__ static void step(Acc &a, float x, Q q) { a = __fadd_rn(a, __fmul_rn(q, x)); }
    __device__ static void nstep(N &n, float v) { n = __fadd_rn(n, __fmul_rn(v, v)); }
    __device__ static float finish(const Acc &dot, N xn, N qn)
    {
        float n1 = __fsqrt_rn(qn), n2 = __fsqrt_rn(xn);
        if (n1 == 0.0f || n2 == 0.0f) return 1.0f;
        return __fsub_rn(1.0f, __fdiv_rn(dot, __fmul_rn(n1, n2)));
    }
};
// -dot in f32 sequential: NOT in the reference's IVF; sign as hnsw_am.c:1334-1337
template <> struct Arith<NDB_IP, NDB_ARITH_IVF_F32> {
    using Q = float; using N = float; using Acc = float;
    static constexpr bool NORMS = false;
    __device__ static void init(Acc &a) { a = 0.0f; }
    __device__ static void step(Acc &a, float x, Q q) { a = __fadd_rn(a, __fmul_rn(q, x)); }
    __device__ static void nstep(N &, float) {}
    __device__ static float finish(const Acc &dot, N, N) { return -dot; }
};

// ---------------------------------------------------------------------------------------
// operator arithmetic: fp64.  l2_distance is Kahan-compensated.
// ---------------------------------------------------------------------------------------
struct KahanAcc { double sum, c; };
template <> struct Arith<NDB_L2, NDB_ARITH_OP_F64> {
    using Q = double; using N = double; using Acc = KahanAcc;
    static constexpr bool NORMS = false;
    __device__ static void init(Acc &a) { a.sum = 0.0; a.c = 0.0; }
    __device__ static void step(Acc &a, float x, Q q)
    {
        double diff = __dsub_rn((double) x, q);       // a = row, b = query
        double y = __dsub_rn(__dmul_rn(diff, diff), a.c);
        double t = __dadd_rn(a.sum, y);
        a.c = __dsub_rn(__dsub_rn(t, a.sum), y);
        a.sum = t;
    }
    __device__ static void nstep(N &, float) {}
    __device__ static float finish(const Acc &a, N, N) { return __double2float_rn(__dsqrt_rn(a.sum)); }
};
// <#>: inner_product_simd falls through to -inner_product_distance = -((float)(-sum)) = +dot.
// (double)a*(double)b is exact (24x24 bits), so fma(a,b,sum) == sum + a*b rounded once.
template <> struct Arith<NDB_IP, NDB_ARITH_OP_F64> {
    using Q = double; using N = double; using Acc = double;
    static constexpr bool NORMS = false;
    __device__ static void init(Acc &a) { a = 0.0; }
    __device__ static void step(Acc &a, float x, Q q) { a = __fma_rn((double) x, q, a); }
    __device__ static void nstep(N &, float) {}
    __device__ static float finish(const Acc &a, N, N) { return -__double2float_rn(-a); }
};
template <> struct Arith<NDB_COSINE, NDB_ARITH_OP_F64> {
    using Q = double; using N = double; using Acc = double;
    static constexpr bool NORMS = true;
    __device__ static void init(Acc &a) { a = 0.0; }
    __device__ static void step(Acc &a, float x, Q q) { a = __fma_rn((double) x, q, a); }   // exact product
    __device__ static void nstep(N &n, float v) { n = __fma_rn((double) v, (double) v, n); }
    __device__ static float finish(const Acc &dot, N xn, N qn)
    {
        if (xn == 0.0 || qn == 0.0) return 1.0f;
        // 1.0 - dot / (sqrt(norm_a) * sqrt(norm_b)); a = row, b = query
        double den = __dmul_rn(__dsqrt_rn(xn), __dsqrt_rn(qn));
        return __double2float_rn(__dsub_rn(1.0, __ddiv_rn(dot, den)));
    }
};

// ---------------------------------------------------------------------------------------
// hnswComputeDistance: the f32 operation is rounded to f32, then accumulated in f64.
// ---------------------------------------------------------------------------------------
template <> struct Arith<NDB_L2, NDB_ARITH_HNSW> {
    using Q = float; using N = double; using Acc = double;
    static constexpr bool NORMS = false;
    __device__ static void init(Acc &a) { a = 0.0; }
    __device__ static void step(Acc &a, float x, Q q)
    {
        double d = (double) __fsub_rn(q, x);          // double d = vec1[i] - vec2[i]  (f32 subtract)
        a = __fma_rn(d, d, a);                        // d*d exact in f64 => one rounding, as sum += d*d
    }
    __device__ static void nstep(N &, float) {}
    __device__ static float finish(const Acc &a, N, N) { return __double2float_rn(__dsqrt_rn(a)); }
};
template <> struct Arith<NDB_COSINE, NDB_ARITH_HNSW> {
    using Q = float; using N = double; using Acc = double;
    static constexpr bool NORMS = true;
    __device__ static void init(Acc &a) { a = 0.0; }
    __device__ static void step(Acc &a, float x, Q q) { a = __dadd_rn(a, (double) __fmul_rn(q, x)); }
    __device__ static void nstep(N &n, float v) { n = __dadd_rn(n, (double) __fmul_rn(v, v)); }
    __device__ static float finish(const Acc &dot, N xn, N qn)
    {
        double n1 = __dsqrt_rn(qn), n2 = __dsqrt_rn(xn);
        if (n1 == 0.0 || n2 == 0.0) return 2.0f;
        return __double2float_rn(__dsub_rn(1.0, __ddiv_rn(dot, __dmul_rn(n1, n2))));
    }
};
template <> struct Arith<NDB_IP, NDB_ARITH_HNSW> {
    using Q = float; using N = double; using Acc = double;
    static constexpr bool NORMS = false;
    __device__ static void init(Acc &a) { a = 0.0; }
    __device__ static void step(Acc &a, float x, Q q) { a = __dadd_rn(a, (double) __fmul_rn(q, x)); }
    __device__ static void nstep(N &, float) {}
    __device__ static float finish(const Acc &a, N, N) { return __double2float_rn(-a); }
};

// ---------------------------------------------------------------------------------------
// FAST: fp32 FFMA.  Not bit-identical to any reference loop; |rel err| <= 1e-5 contract.
// ---------------------------------------------------------------------------------------
template <> struct Arith<NDB_L2, NDB_ARITH_FAST> {
    using Q = float; using N = float; using Acc = float;
    static constexpr bool NORMS = false;
    __device__ static void init(Acc &a) { a = 0.0f; }
    __device__ static void step(Acc &a, float x, Q q) { float d = x - q; a = fmaf(d, d, a); }
    __device__ static void nstep(N &, float) {}
    __device__ static float finish(const Acc &a, N, N) { return sqrtf(a); }
};
template <> struct Arith<METRIC_L2SQ, NDB_ARITH_FAST> : Arith<NDB_L2, NDB_ARITH_FAST> {
    __device__ static float finish(const Acc &a, N, N) { return a; }
};
template <> struct Arith<NDB_IP, NDB_ARITH_FAST> {
    using Q = float; using N = float; using Acc = float;
    static constexpr bool NORMS = false;
    __device__ static void init(Acc &a) { a = 0.0f; }
    __device__ static void step(Acc &a, float x, Q q) { a = fmaf(x, q, a); }
    __device__ static void nstep(N &, float) {}
    __device__ static float finish(const Acc &a, N, N) { return -a; }
};
template <> struct Arith<NDB_COSINE, NDB_ARITH_FAST> {
    using Q = float; using N = float; using Acc = float;
    static constexpr bool NORMS = true;
    __device__ static void init(Acc &a) { a = 0.0f; }
    __device__ static void step(Acc &a, float x, Q q) { a = fmaf(x, q, a); }
    __device__ static void nstep(N &n, float v) { n = fmaf(v, v, n); }
    __device__ static float finish(const Acc &dot, N xn, N qn)
    {
        if (xn == 0.0f || qn == 0.0f) return 1.0f;
        return 1.0f - dot / (sqrtf(xn) * sqrtf(qn));
    }
};

// ---------------------------------------------------------------------------------------
// warp-distributed sorted top-K: entry e = r*32 + lane, ascending by (dist, key).
// ---------------------------------------------------------------------------------------
template <class KeyT> __device__ __forceinline__ bool pair_less(float d1, KeyT k1, float d2, KeyT k2)
{
    return d1 < d2 || (d1 == d2 && k1 < k2);
}

template <class KeyT> __device__ __forceinline__ KeyT shfl_key(KeyT v, int src);
template <> __device__ __forceinline__ uint32_t shfl_key<uint32_t>(uint32_t v, int src) { return __shfl_sync(FULL, v, src); }
template <> __device__ __forceinline__ int64_t shfl_key<int64_t>(int64_t v, int src) { return __shfl_sync(FULL, v, src); }
template <class KeyT> __device__ __forceinline__ KeyT shfl_up_key(KeyT v);
template <> __device__ __forceinline__ uint32_t shfl_up_key<uint32_t>(uint32_t v) { return __shfl_up_sync(FULL, v, 1); }
template <> __device__ __forceinline__ int64_t shfl_up_key<int64_t>(int64_t v) { return __shfl_up_sync(FULL, v, 1); }

template <class KeyT> __device__ __forceinline__ KeyT shfl_xor_key(KeyT v, int m);
template <> __device__ __forceinline__ uint32_t shfl_xor_key<uint32_t>(uint32_t v, int m) { return __shfl_xor_sync(FULL, v, m); }
template <> __device__ __forceinline__ int64_t shfl_xor_key<int64_t>(int64_t v, int m) { return __shfl_xor_sync(FULL, v, m); }

template <class KeyT> struct KeyMax;
template <> struct KeyMax<uint32_t> { static constexpr uint32_t v = 0xffffffffu; };
template <> struct KeyMax<int64_t> { static constexpr int64_t v = 0x7fffffffffffffffll; };

// full 32-lane bitonic sort, ascending by (d, key): 15 shuffle stages
template <class KeyT> __device__ __forceinline__ void warp_sort32(float &d, KeyT &key, int lane)
{
#pragma unroll
    for (int kk = 2; kk <= 32; kk <<= 1) {
#pragma unroll
        for (int j = kk >> 1; j > 0; j >>= 1) {
            const float pd = __shfl_xor_sync(FULL, d, j);
            const KeyT pk = shfl_xor_key<KeyT>(key, j);
            const bool want_min = ((lane & j) == 0) == ((lane & kk) == 0);
            const bool take = want_min ? pair_less<KeyT>(pd, pk, d, key) : pair_less<KeyT>(d, key, pd, pk);
            if (take) { d = pd; key = pk; }
        }
    }
}

template <int KR, class KeyT> struct WarpTopK {
    float d[KR];
    KeyT key[KR];
    float td;       // threshold = entry k-1 (warp-uniform)
    KeyT tk;

    __device__ __forceinline__ void init()
    {
#pragma unroll
        for (int r = 0; r < KR; r++) { d[r] = INFINITY; key[r] = KeyMax<KeyT>::v; }
        td = INFINITY;
        tk = KeyMax<KeyT>::v;
    }
    __device__ __forceinline__ void refresh_threshold(int k)
    {
        const int tr = (k - 1) >> 5, tl = (k - 1) & 31;
        float v = d[0];
        KeyT kk = key[0];
#pragma unroll
        for (int r = 1; r < KR; r++) if (r == tr) { v = d[r]; kk = key[r]; }
        td = __shfl_sync(FULL, v, tl);
        tk = shfl_key<KeyT>(kk, tl);
    }
    // warp-uniform insert of one candidate that already passed the threshold test
    __device__ __forceinline__ void insert(float nd, KeyT nk, int lane, int k)
    {
        int pos = 0;
#pragma unroll
        for (int r = 0; r < KR; r++) pos += __popc(__ballot_sync(FULL, pair_less<KeyT>(d[r], key[r], nd, nk)));
#pragma unroll
        for (int r = KR - 1; r >= 0; r--) {
            float ud = __shfl_up_sync(FULL, d[r], 1);
            KeyT uk = shfl_up_key<KeyT>(key[r]);
            if (r > 0) {
                float cd = __shfl_sync(FULL, d[r - 1], 31);
                KeyT ck = shfl_key<KeyT>(key[r - 1], 31);
                if (lane == 0) { ud = cd; uk = ck; }
            }
            const int e = r * 32 + lane;
            if (e == pos) { d[r] = nd; key[r] = nk; }
            else if (e > pos) { d[r] = ud; key[r] = uk; }
        }
        refresh_threshold(k);
    }
    // KR == 1 only: merge up to 32 candidates (one per lane) in one go.  Bitonic-sort the
    // candidates across the lanes, take the element-wise minimum with the reversed current list
    // (which leaves the 32 smallest of the union as a bitonic sequence) and finish with one
    // bitonic merge: 21 shuffle stages regardless of how many lanes carry a candidate.
    __device__ __forceinline__ void merge32(float cd, KeyT ck, bool pass, int lane, int k)
    {
        float nd = pass ? cd : INFINITY;
        KeyT nk = pass ? ck : KeyMax<KeyT>::v;
#pragma unroll
        for (int kk = 2; kk <= 32; kk <<= 1) {
#pragma unroll
            for (int j = kk >> 1; j > 0; j >>= 1) {
                const float pd = __shfl_xor_sync(FULL, nd, j);
                const KeyT pk = shfl_xor_key<KeyT>(nk, j);
                const bool want_min = ((lane & j) == 0) == ((lane & kk) == 0);
                const bool take = want_min ? pair_less<KeyT>(pd, pk, nd, nk) : pair_less<KeyT>(nd, nk, pd, pk);
                if (take) { nd = pd; nk = pk; }
            }
        }
        const float rd = __shfl_sync(FULL, nd, 31 - lane);
        const KeyT rk = shfl_key<KeyT>(nk, 31 - lane);
        if (pair_less<KeyT>(rd, rk, d[0], key[0])) { d[0] = rd; key[0] = rk; }
#pragma unroll
        for (int j = 16; j > 0; j >>= 1) {
            const float pd = __shfl_xor_sync(FULL, d[0], j);
            const KeyT pk = shfl_xor_key<KeyT>(key[0], j);
            const bool want_min = (lane & j) == 0;
            const bool take = want_min ? pair_less<KeyT>(pd, pk, d[0], key[0]) : pair_less<KeyT>(d[0], key[0], pd, pk);
            if (take) { d[0] = pd; key[0] = pk; }
        }
        refresh_threshold(k);
    }
    // every lane offers (cd, ck) if valid; all 32 lanes must call
    __device__ __forceinline__ void offer(float cd, KeyT ck, bool valid, int lane, int k)
    {
        const bool pass = valid && pair_less<KeyT>(cd, ck, td, tk);
        unsigned m = __ballot_sync(FULL, pass);
        if (KR == 1 && __popc(m) >= 8) {       // serial insertion costs ~20 instr per candidate
            merge32(cd, ck, pass, lane, k);
            return;
        }
        while (m) {
            const int src = __ffs(m) - 1;
            m &= m - 1;
            const float nd = __shfl_sync(FULL, cd, src);
            const KeyT nk = shfl_key<KeyT>(ck, src);
            if (pair_less<KeyT>(nd, nk, td, tk)) insert(nd, nk, lane, k);
        }
    }
};

}  // namespace ndb
