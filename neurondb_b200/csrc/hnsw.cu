// hnsw.cu -- HNSW graph search and build (NeuronDB/src/index/hnsw_am.c: hnswSearch :1545-2080,
// hnswInsertNode :2091-2670, hnswComputeDistance :1301-1345).
//
// Device graph (all flat arrays, node id = insertion index = the reference's BlockNumber - 1):
//   vec   [n][dimp] f32 row-major (rows padded to 4 floats so float4 loads are aligned)
//   level [n] i32, cnt [n][16] i16 (neighborCount per level), nbr0 [n][2m] u32 (level 0),
//   upper_off [n+1] i64 + upper [...] u32: levels 1..level of node i at upper_off[i] + (lev-1)*2m
// 0xFFFFFFFF is InvalidBlockNumber (slots are memset 0xFF, hnsw_am.c:2149).
//
// One warp serves one query.  Expanding a node evaluates its <= 2m neighbours with ONE LANE
// PER NEIGHBOUR: lane j walks the dimensions of neighbour j in order, accumulating exactly as
// hnswComputeDistance does (f32 operation, f64 running sum), so distances are bit-identical to
// the reference; the query sits in shared memory (broadcast reads).  Visited marks live in a
// per-warp bitset in global memory (n bits, L2-resident) that is cleared through the list of
// touched nodes.  Search modes:
//   NDB_HNSW_LITERAL   greedy descent + level-0 BFS that stops at ef candidates
//   NDB_HNSW_BESTFIRST search_layer with beam ef; W is one array sorted by (dist,id) in shared
//                      memory and the nearest unexpanded entry is expanded until none is left.
#include "layout.cuh"
#include "arith.cuh"
#include "pages.cuh"
#include "comm.cuh"

#include <algorithm>
#include <cub/device/device_radix_sort.cuh>
#include <cstdlib>

using namespace ndb;

namespace ndb {

constexpr uint32_t EXPANDED = 0x80000000u;
constexpr int HNSW_MAX_LEVEL = 16;

struct HnswGraph {
    const float *vec;
    const int64_t *ids;
    const int *level;
    int16_t *cnt;
    uint32_t *nbr0;
    const int64_t *upper_off;
    uint32_t *upper;
    const double *vnorm;
    int64_t n;
    int dim, dimp, m2;
};

__device__ __forceinline__ uint32_t *nbr_slots(const HnswGraph &g, uint32_t node, int lev)
{
    if (lev == 0) return g.nbr0 + (size_t) node * g.m2;
    if (lev > g.level[node]) return nullptr;
    return g.upper + g.upper_off[node] + (size_t) (lev - 1) * g.m2;
}

// hnswValidateNeighborCount (:1167-1187)
__device__ __forceinline__ int nbr_count(const HnswGraph &g, uint32_t node, int lev)
{
    int c = g.cnt[(size_t) node * HNSW_MAX_LEVEL + lev];
    return c < 0 ? 0 : (c > g.m2 ? g.m2 : c);
}

template <class P>
__device__ __forceinline__ float node_distance(const HnswGraph &g, const float *qs, typename P::N qn, uint32_t node)
{
    const float4 *vp = reinterpret_cast<const float4 *>(g.vec + (size_t) node * g.dimp);
    typename P::Acc acc;
    P::init(acc);
    const int nfull = g.dim >> 2, rem = g.dim & 3;
#pragma unroll 4
    for (int c = 0; c < nfull; c++) {
        const float4 x = __ldg(vp + c);
        const float4 q = *reinterpret_cast<const float4 *>(qs + 4 * c);
        P::step(acc, x.x, q.x); P::step(acc, x.y, q.y); P::step(acc, x.z, q.z); P::step(acc, x.w, q.w);
    }
    if (rem) {
        const float4 x = __ldg(vp + nfull);
        const float *q = qs + 4 * nfull;
        P::step(acc, x.x, q[0]);
        if (rem > 1) P::step(acc, x.y, q[1]);
        if (rem > 2) P::step(acc, x.z, q[2]);
    }
    return P::finish(acc, P::NORMS ? (typename P::N) g.vnorm[node] : (typename P::N) 0, qn);
}

struct WarpCtx {
    float *qs;            // [dimp] query (shared)
    float *wd;            // [ef] distances (shared)
    uint32_t *wid;        // [ef] node ids, top bit = expanded (shared)
    int *widx;            // [ef] scratch index array for the literal selection sort (shared)
    // Visited set.  Searches keep it in SHARED memory: an exact open-addressing hash set of node ids (hset, hcap slots,
    // linear probing, EMPTY = 0xFFFFFFFF), one per warp -- a query at ef_search = 40 touches ~1 400 of a million nodes,
    // and a bitset over all nodes (125 KB per warp at 1 M nodes, 600 MB for the grid) turns every test into a DRAM
    // access.  When the set passes 7/8 of its slots (large ef) its content moves to the per-warp bitset in global memory
    // and the search carries on there; the graph build (ef_construction >= 64, every level) uses the bitset throughout.
    uint32_t *hset;       // shared; nullptr = bitset only
    int hcap;             // power of two
    bool hashed;          // the set currently lives in hset
    uint32_t *bits;       // per-warp visited bitset, (n+31)/32 words (global)
    uint32_t *vlist;      // nodes whose bit was set (global)
    int vcap, nvis;
    long long evals;
};

constexpr uint32_t HSET_EMPTY = 0xFFFFFFFFu;
__device__ __forceinline__ uint32_t hset_hash(uint32_t e) { return e * 2654435761u; }

__device__ __forceinline__ bool test_bit(const uint32_t *bits, uint32_t e) { return (bits[e >> 5] >> (e & 31)) & 1u; }

// has node e been visited (any lane, any e < n; no side effect)
__device__ __forceinline__ bool is_visited(const WarpCtx &w, uint32_t e)
{
    if (!w.hashed) return test_bit(w.bits, e);
    const uint32_t mask = (uint32_t) w.hcap - 1u;
    for (uint32_t h = (hset_hash(e) >> 8) & mask;; h = (h + 1u) & mask) {
        const uint32_t v = w.hset[h];
        if (v == e) return true;
        if (v == HSET_EMPTY) return false;
    }
}

// marks `e` visited for every lane with fresh == true (the fresh lanes hold distinct nodes) and records it for the later
// clear; moves the set to the global bitset when the shared one fills up
__device__ __forceinline__ void mark_visited(WarpCtx &w, uint32_t e, bool fresh, int lane)
{
    const unsigned fm = __ballot_sync(FULL, fresh);
    if (w.hashed) {
        if (fresh) {
            const uint32_t mask = (uint32_t) w.hcap - 1u;
            for (uint32_t h = (hset_hash(e) >> 8) & mask;; h = (h + 1u) & mask)
                if (atomicCAS(&w.hset[h], HSET_EMPTY, e) == HSET_EMPTY) break;
        }
        w.nvis += __popc(fm);
        __syncwarp();
        if (w.nvis > w.hcap - w.hcap / 8) {
            int cnt = 0;
            for (int base = 0; base < w.hcap; base += 32) {
                const uint32_t v = w.hset[base + lane];
                const bool ok = v != HSET_EMPTY;
                const unsigned m = __ballot_sync(FULL, ok);
                if (ok) {
                    atomicOr(&w.bits[v >> 5], 1u << (v & 31));
                    const int pos = cnt + __popc(m & ((1u << lane) - 1));
                    if (pos < w.vcap) w.vlist[pos] = v;
                }
                cnt += __popc(m);
            }
            w.nvis = cnt;
            w.hashed = false;
            __syncwarp();
        }
        return;
    }
    if (fresh) {
        atomicOr(&w.bits[e >> 5], 1u << (e & 31));
        const int pos = w.nvis + __popc(fm & ((1u << lane) - 1));
        if (pos < w.vcap) w.vlist[pos] = e;
    }
    w.nvis += __popc(fm);
    __syncwarp();
}

__device__ __forceinline__ void clear_visited(WarpCtx &w, int64_t n, int lane)
{
    __syncwarp();
    if (w.hset) {
        for (int i = lane; i < w.hcap; i += 32) w.hset[i] = HSET_EMPTY;
        if (w.hashed) { w.nvis = 0; __syncwarp(); return; }
        w.hashed = true;            // (the bitset is cleared below; the next query starts in shared memory again)
    }
    if (w.nvis <= w.vcap) {
        for (int i = lane; i < w.nvis; i += 32) w.bits[w.vlist[i] >> 5] = 0u;
    } else {
        const int64_t words = (n + 31) >> 5;
        for (int64_t i = lane; i < words; i += 32) w.bits[i] = 0u;
    }
    w.nvis = 0;
    __syncwarp();
}

template <class P>
__device__ __forceinline__ float eval_one(const HnswGraph &g, WarpCtx &w, typename P::N qn, uint32_t node, int lane)
{
    float d = 0.0f;
    if (lane == 0) d = node_distance<P>(g, w.qs, qn, node);
    w.evals++;
    return __shfl_sync(FULL, d, 0);
}

// greedy descent on levels from_level .. to_level_exclusive+1 (hnsw_am.c:1638-1750)
template <class P>
__device__ uint32_t greedy_descent(const HnswGraph &g, WarpCtx &w, typename P::N qn, uint32_t cur, int from_level,
                                   int to_level_exclusive, int lane)
{
    for (int level = from_level; level > to_level_exclusive; level--) {
        bool better;
        do {
            better = false;
            if (cur == INVALID_SLOT || cur >= (uint32_t) g.n) break;
            float bestd = eval_one<P>(g, w, qn, cur, lane);
            uint32_t beste = cur;
            if (g.level[cur] >= level) {
                const uint32_t *slots = nbr_slots(g, cur, level);
                const int nc = nbr_count(g, cur, level);
                for (int base = 0; base < nc; base += 32) {
                    const int j = base + lane;
                    const uint32_t e = j < nc ? slots[j] : INVALID_SLOT;
                    const bool valid = e != INVALID_SLOT && e < (uint32_t) g.n;
                    float d = INFINITY;
                    if (valid) d = node_distance<P>(g, w.qs, qn, e);
                    w.evals += __popc(__ballot_sync(FULL, valid));
                    // running strict-< minimum in slot order == first occurrence of the minimum
                    float md = d;
                    int ml = valid ? lane : 64;
#pragma unroll
                    for (int o = 16; o > 0; o >>= 1) {
                        const float od = __shfl_xor_sync(FULL, md, o);
                        const int ol = __shfl_xor_sync(FULL, ml, o);
                        if (od < md || (od == md && ol < ml)) { md = od; ml = ol; }
                    }
                    if (ml < 32 && md < bestd) {
                        bestd = md;
                        beste = __shfl_sync(FULL, e, ml);
                    }
                }
            }
            if (beste != cur) { cur = beste; better = true; }
        } while (better);
    }
    return cur;
}

// best-first search_layer at `lev`; returns the number of entries in W (sorted by (dist,id))
template <class P>
__device__ int search_layer(const HnswGraph &g, WarpCtx &w, typename P::N qn, uint32_t ep, int lev, int ef, int lane)
{
    mark_visited(w, ep, lane == 0 && !is_visited(w, ep), lane);
    const float d0 = eval_one<P>(g, w, qn, ep, lane);
    if (lane == 0) { w.wd[0] = d0; w.wid[0] = ep; }
    __syncwarp();
    int wn = 1;
    for (;;) {
        int ci = 0x7fffffff;
        for (int i = lane; i < wn; i += 32)
            if (!(w.wid[i] & EXPANDED)) { ci = i; break; }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) ci = min(ci, __shfl_xor_sync(FULL, ci, o));
        if (ci == 0x7fffffff) break;
        const uint32_t c = w.wid[ci];
        __syncwarp();
        if (lane == 0) w.wid[ci] = c | EXPANDED;
        __syncwarp();
        const uint32_t *slots = nbr_slots(g, c, lev);
        const int nc = slots ? nbr_count(g, c, lev) : 0;
        for (int base = 0; base < nc; base += 32) {
            const int j = base + lane;
            const uint32_t e = j < nc ? slots[j] : INVALID_SLOT;
            const bool valid = e != INVALID_SLOT && e < (uint32_t) g.n;
            const unsigned same = __match_any_sync(FULL, e);
            const bool fresh = valid && (__ffs(same) - 1 == lane) && !is_visited(w, e);
            mark_visited(w, e, fresh, lane);
            float d = INFINITY;
            if (fresh) d = node_distance<P>(g, w.qs, qn, e);
            unsigned em = __ballot_sync(FULL, fresh);
            w.evals += __popc(em);
            while (em) {
                const int src = __ffs(em) - 1;
                em &= em - 1;
                const float nd = __shfl_sync(FULL, d, src);
                const uint32_t ne = __shfl_sync(FULL, e, src);
                if (wn == ef && !pair_less<uint32_t>(nd, ne, w.wd[wn - 1], w.wid[wn - 1] & ~EXPANDED)) continue;
                int pos = 0;
                for (int i = lane; i < wn; i += 32) pos += pair_less<uint32_t>(w.wd[i], w.wid[i] & ~EXPANDED, nd, ne) ? 1 : 0;
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) pos += __shfl_xor_sync(FULL, pos, o);
                const int last = wn < ef ? wn : ef - 1;
                for (int hi = last; hi > pos; hi -= 32) {
                    const int i = hi - lane;
                    const bool mv = i > pos;
                    float td = 0.0f;
                    uint32_t ti = 0;
                    if (mv) { td = w.wd[i - 1]; ti = w.wid[i - 1]; }
                    __syncwarp();
                    if (mv) { w.wd[i] = td; w.wid[i] = ti; }
                    __syncwarp();
                }
                if (lane == 0) { w.wd[pos] = nd; w.wid[pos] = ne; }
                __syncwarp();
                if (wn < ef) wn++;
            }
        }
    }
    return wn;
}

// level-0 phase of hnswSearch as written (:1752-2013): BFS over candidates[] in array order,
// stops expanding once ef candidates exist; then the selection sort with swaps.
template <class P>
__device__ int level0_literal(const HnswGraph &g, WarpCtx &w, typename P::N qn, uint32_t cur, int ef, int k, int lane)
{
    mark_visited(w, cur, lane == 0, lane);
    const float d0 = eval_one<P>(g, w, qn, cur, lane);
    if (lane == 0) { w.wd[0] = d0; w.wid[0] = cur; }
    __syncwarp();
    int cc = 1;
    for (int i = 0; i < cc && cc < ef; i++) {
        const uint32_t c = w.wid[i];
        if (c == INVALID_SLOT || c >= (uint32_t) g.n) continue;
        const uint32_t *slots = nbr_slots(g, c, 0);
        const int nc = nbr_count(g, c, 0);
        for (int base = 0; base < nc; base += 32) {
            const int j = base + lane;
            const uint32_t e = j < nc ? slots[j] : INVALID_SLOT;
            const bool valid = e != INVALID_SLOT && e < (uint32_t) g.n;
            const unsigned same = __match_any_sync(FULL, e);
            const bool fresh = valid && (__ffs(same) - 1 == lane) && !is_visited(w, e);
            mark_visited(w, e, fresh, lane);
            float d = INFINITY;
            if (fresh) d = node_distance<P>(g, w.qs, qn, e);
            unsigned em = __ballot_sync(FULL, fresh);
            w.evals += __popc(em);
            while (em) {
                const int src = __ffs(em) - 1;
                em &= em - 1;
                const float nd = __shfl_sync(FULL, d, src);
                const uint32_t ne = __shfl_sync(FULL, e, src);
                if (cc < ef) {
                    if (lane == 0) { w.wd[cc] = nd; w.wid[cc] = ne; }
                    cc++;
                } else {
                    // worst = first occurrence of the maximum (:1951-1961)
                    float wdm = -INFINITY;
                    int wi = 0x7fffffff;
                    for (int l = lane; l < cc; l += 32) {
                        const float v = w.wd[l];
                        if (v > wdm || wi == 0x7fffffff) { wdm = v; wi = l; }
                    }
#pragma unroll
                    for (int o = 16; o > 0; o >>= 1) {
                        const float od = __shfl_xor_sync(FULL, wdm, o);
                        const int oi = __shfl_xor_sync(FULL, wi, o);
                        if (oi != 0x7fffffff && (wi == 0x7fffffff || od > wdm || (od == wdm && oi < wi))) { wdm = od; wi = oi; }
                    }
                    if (nd < wdm && lane == 0) { w.wd[wi] = nd; w.wid[wi] = ne; }
                }
                __syncwarp();
            }
        }
    }
    // selection sort over an index array with swaps, strict < (:1984-2013)
    for (int i = lane; i < cc; i += 32) w.widx[i] = i;
    __syncwarp();
    if (lane == 0) {
        for (int i = 0; i < k && i < cc; i++) {
            int minIdx = i;
            float minDist = w.wd[w.widx[i]];
            for (int j = i + 1; j < cc; j++)
                if (w.wd[w.widx[j]] < minDist) { minDist = w.wd[w.widx[j]]; minIdx = j; }
            if (minIdx != i) { const int t = w.widx[i]; w.widx[i] = w.widx[minIdx]; w.widx[minIdx] = t; }
        }
    }
    __syncwarp();
    return cc;
}

struct HnswSearchArgs {
    HnswGraph g;
    const float *Q;
    int nq, ef, k, mode;
    uint32_t entry;
    int entry_level;
    uint32_t *bits;          // [total_warps][words]
    uint32_t *vlist;         // [total_warps][vcap]
    int64_t words;
    int vcap;
    int hcap;                // slots of the per-warp visited hash set in shared memory (0: bitset only)
    float *out_dist;
    int64_t *out_ids;
    unsigned long long *evals;
};

template <class P>
__global__ void hnsw_search_kernel(const HnswSearchArgs a)
{
    extern __shared__ __align__(16) unsigned char hs[];
    const int wpb = blockDim.x >> 5, lane = threadIdx.x & 31, wl = threadIdx.x >> 5;
    const size_t per_warp = (size_t) a.g.dimp * 4 + (size_t) a.ef * 12 + (size_t) a.hcap * 4;
    unsigned char *base = hs + per_warp * wl;
    WarpCtx w;
    w.qs = reinterpret_cast<float *>(base);
    w.wd = reinterpret_cast<float *>(base + (size_t) a.g.dimp * 4);
    w.wid = reinterpret_cast<uint32_t *>(w.wd + a.ef);
    w.widx = reinterpret_cast<int *>(w.wid + a.ef);
    w.hset = a.hcap ? reinterpret_cast<uint32_t *>(w.widx + a.ef) : nullptr;
    w.hcap = a.hcap;
    w.hashed = a.hcap != 0;
    for (int i = lane; i < a.hcap; i += 32) w.hset[i] = HSET_EMPTY;
    __syncwarp();
    const int gw = blockIdx.x * wpb + wl, total = gridDim.x * wpb;
    w.bits = a.bits + (size_t) gw * a.words;
    w.vlist = a.vlist + (size_t) gw * a.vcap;
    w.vcap = a.vcap;
    w.nvis = 0;
    w.evals = 0;
    for (int q = gw; q < a.nq; q += total) {
        for (int i = lane; i < a.g.dimp; i += 32) w.qs[i] = i < a.g.dim ? a.Q[(size_t) q * a.g.dim + i] : 0.0f;
        __syncwarp();
        typename P::N qn = 0;
        if (P::NORMS) for (int i = 0; i < a.g.dim; i++) P::nstep(qn, w.qs[i]);
        int count = 0;
        if (a.entry != INVALID_SLOT && a.g.n > 0) {
            int cl = a.entry_level;
            if (cl < 0 || cl >= HNSW_MAX_LEVEL) cl = 0;
            const uint32_t cur = greedy_descent<P>(a.g, w, qn, a.entry, cl, 0, lane);
            if (a.mode == NDB_HNSW_LITERAL) {
                const int cc = level0_literal<P>(a.g, w, qn, cur, a.ef, a.k, lane);
                count = min(a.k, cc);
                for (int i = lane; i < count; i += 32) {
                    a.out_dist[(size_t) q * a.k + i] = w.wd[w.widx[i]];
                    a.out_ids[(size_t) q * a.k + i] = a.g.ids[w.wid[w.widx[i]]];
                }
            } else {
                const int wn = search_layer<P>(a.g, w, qn, cur, 0, a.ef, lane);
                count = min(a.k, wn);
                for (int i = lane; i < count; i += 32) {
                    a.out_dist[(size_t) q * a.k + i] = w.wd[i];
                    a.out_ids[(size_t) q * a.k + i] = a.g.ids[w.wid[i] & ~EXPANDED];
                }
            }
            clear_visited(w, a.g.n, lane);
        }
        for (int i = count + lane; i < a.k; i += 32) {
            a.out_dist[(size_t) q * a.k + i] = INFINITY;
            a.out_ids[(size_t) q * a.k + i] = -1;
        }
        __syncwarp();
    }
    if (lane == 0 && w.evals) atomicAdd(a.evals, (unsigned long long) w.evals);
}

// ---- build: batched hnswInsertNode ---------------------------------------------------------
struct HnswBuildArgs {
    HnswGraph g;
    uint32_t first, count;       // nodes [first, first+count) are inserted by this launch
    int m, efc;
    uint32_t entry;
    int entry_level;
    uint32_t *bits;
    uint32_t *vlist;
    int64_t words;
    int vcap;
    unsigned long long *req;     // back-link requests: ((u*16+lev) << 32) | v
    unsigned int *nreq;
    unsigned long long *evals;
    int heuristic;               // NDB_HNSW_SELECT_HEURISTIC: diversity selection instead of closest-m
};

// Neighbour selection by the diversity heuristic (Malkov & Yashunin, Alg. 4; hnswlib's
// getNeighborsByHeuristic2), an EXTENSION over the reference's closest-m rule: candidates are taken in
// ascending (distance to the base, id) order and one is kept if it is closer to the base than to every
// neighbour kept so far.  cand(i, &id, &d) yields the i-th candidate (warp-uniform).  The kept
// neighbours live one per lane (lane j < kept holds the j-th), so a candidate is checked against all
// of them at once: lane j walks the dimensions of (candidate, kept_j) in order -- the same sequential
// arithmetic as everywhere else, hence the same graph as the CPU oracle when built one node at a time.
// cs: [dimp] shared-memory staging of the candidate's vector.  Returns the number kept (<= want <= 32).
template <class P, class CandFn>
__device__ int select_heuristic_warp(const HnswGraph &g, WarpCtx &w, float *cs, int ncand, int want, int lane, CandFn cand,
                                     uint32_t &my_id, float &my_d)
{
    int kept = 0;
    my_id = INVALID_SLOT;
    my_d = INFINITY;
    for (int i = 0; i < ncand && kept < want; i++) {
        uint32_t c;
        float dc;
        cand(i, c, dc);
        __syncwarp();
        for (int t = lane; t < g.dimp; t += 32) cs[t] = g.vec[(size_t) c * g.dimp + t];
        __syncwarp();
        bool closer_to_kept = false;
        if (lane < kept) closer_to_kept = node_distance<P>(g, cs, 0.0, my_id) < dc;
        w.evals += kept;
        if (!__any_sync(FULL, closer_to_kept)) {
            if (lane == kept) { my_id = c; my_d = dc; }
            kept++;
        }
    }
    __syncwarp();
    return kept;
}

// per new node: search the graph as it stood before this batch, keep the closest m per level
// (hnsw_am.c:2365-2424), write the forward links (:2452-2458), queue the back-links
__global__ void hnsw_build_search_kernel(const HnswBuildArgs a)
{
    using P = Arith<NDB_L2, NDB_ARITH_HNSW>;          // hnswInsertNode always searches with L2 (:2375-2384)
    extern __shared__ __align__(16) unsigned char hs[];
    const int wpb = blockDim.x >> 5, lane = threadIdx.x & 31, wl = threadIdx.x >> 5;
    const size_t per_warp = (size_t) a.g.dimp * 4 * (a.heuristic ? 2 : 1) + (size_t) a.efc * 12;
    unsigned char *base = hs + per_warp * wl;
    WarpCtx w;
    w.qs = reinterpret_cast<float *>(base);
    w.wd = reinterpret_cast<float *>(base + (size_t) a.g.dimp * 4);
    w.wid = reinterpret_cast<uint32_t *>(w.wd + a.efc);
    w.widx = reinterpret_cast<int *>(w.wid + a.efc);
    float *cs = reinterpret_cast<float *>(w.widx + a.efc);          // heuristic only: candidate staging
    w.hset = nullptr;                                               // the build keeps its visited set in the global bitset
    w.hcap = 0;
    w.hashed = false;
    const int gw = blockIdx.x * wpb + wl, total = gridDim.x * wpb;
    w.bits = a.bits + (size_t) gw * a.words;
    w.vlist = a.vlist + (size_t) gw * a.vcap;
    w.vcap = a.vcap;
    w.nvis = 0;
    w.evals = 0;
    HnswGraph g = a.g;
    g.n = a.first;                                    // only nodes inserted before this batch are visible
    for (uint32_t b = gw; b < a.count; b += total) {
        const uint32_t v = a.first + b;
        for (int i = lane; i < g.dimp; i += 32) w.qs[i] = g.vec[(size_t) v * g.dimp + i];
        __syncwarp();
        if (a.entry == INVALID_SLOT || a.entry_level < 0) continue;
        const int level = g.level[v];
        const int maxLevel = min(level, a.entry_level);
        uint32_t ep = greedy_descent<P>(g, w, 0.0, a.entry, a.entry_level, maxLevel, lane);
        for (int lev = maxLevel; lev >= 0; lev--) {
            const int wn = search_layer<P>(g, w, 0.0, ep, lev, a.efc, lane);
            clear_visited(w, g.n, lane);
            int sel = min(a.m, wn);
            uint32_t *mine = lev == 0 ? a.g.nbr0 + (size_t) v * g.m2
                                      : a.g.upper + a.g.upper_off[v] + (size_t) (lev - 1) * g.m2;
            uint32_t hid = INVALID_SLOT;
            float hd = INFINITY;
            if (a.heuristic)
                sel = select_heuristic_warp<P>(g, w, cs, wn, a.m, lane,
                                               [&](int i, uint32_t &c, float &dc) { c = w.wid[i] & ~EXPANDED; dc = w.wd[i]; }, hid, hd);
            unsigned int rbase = 0;
            if (lane == 0 && sel > 0) rbase = atomicAdd(a.nreq, (unsigned int) sel);
            rbase = __shfl_sync(FULL, rbase, 0);
            for (int i = lane; i < sel; i += 32) {
                const uint32_t u = a.heuristic ? hid : (w.wid[i] & ~EXPANDED);
                mine[i] = u;
                a.req[rbase + i] = ((unsigned long long) ((unsigned long long) u * 16 + lev) << 32) | v;
            }
            if (lane == 0) a.g.cnt[(size_t) v * HNSW_MAX_LEVEL + lev] = (int16_t) sel;
            if (wn > 0) ep = w.wid[0] & ~EXPANDED;
            __syncwarp();
        }
    }
    if (lane == 0 && w.evals) atomicAdd(a.evals, (unsigned long long) w.evals);
}

// sorted requests: append v to (u, lev) while a slot below 2m is free (:2493-2513)
__global__ void hnsw_backlink_write_kernel(const unsigned long long *req, unsigned int nreq, HnswGraph g)
{
    const unsigned int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= nreq) return;
    const unsigned long long r = req[i];
    const unsigned long long key = r >> 32;
    int rank = 0;
    while (rank < g.m2 && i >= (unsigned) (rank + 1) && (req[i - rank - 1] >> 32) == key) rank++;
    if (rank >= g.m2) return;
    const uint32_t u = (uint32_t) (key >> 4), v = (uint32_t) (r & 0xffffffffu);
    const int lev = (int) (key & 15);
    uint32_t *slots = nbr_slots(g, u, lev);
    if (!slots) return;
    const int pos = nbr_count(g, u, lev) + rank;
    if (pos < g.m2) slots[pos] = v;
}

// Heuristic mode: one warp per (node u, level) segment of the sorted requests.  The new nodes v of the
// segment are taken in ascending order: appended while u has a free slot, otherwise u's neighbours are
// re-selected among the current ones and v (hnswlib's mutuallyConnectNewElement).
__global__ void hnsw_backlink_select_kernel(const unsigned long long *req, unsigned int nreq, HnswGraph g,
                                            unsigned long long *evals)
{
    using P = Arith<NDB_L2, NDB_ARITH_HNSW>;
    extern __shared__ __align__(16) unsigned char hs[];
    const int lane = threadIdx.x & 31, wl = threadIdx.x >> 5;
    const unsigned int i0 = blockIdx.x * (blockDim.x >> 5) + wl;
    if (i0 >= nreq) return;
    const unsigned long long key = req[i0] >> 32;
    if (i0 > 0 && (req[i0 - 1] >> 32) == key) return;               // not the first request of its segment
    float *us = reinterpret_cast<float *>(hs) + (size_t) wl * 2 * g.dimp;     // u's vector
    float *cs = us + g.dimp;                                        // candidate staging
    const uint32_t u = (uint32_t) (key >> 4);
    const int lev = (int) (key & 15);
    uint32_t *slots = nbr_slots(g, u, lev);
    if (!slots) return;
    WarpCtx w;
    w.qs = us;
    w.evals = 0;
    for (int t = lane; t < g.dimp; t += 32) us[t] = g.vec[(size_t) u * g.dimp + t];
    __syncwarp();
    int nc = nbr_count(g, u, lev);
    for (unsigned int i = i0; i < nreq && (req[i] >> 32) == key; i++) {
        const uint32_t v = (uint32_t) (req[i] & 0xffffffffu);
        if (nc < g.m2) {
            if (lane == 0) slots[nc] = v;
            nc++;
            continue;
        }
        // full: (distance to u, id) of the current neighbours, one per lane, and of v
        const uint32_t e = lane < nc ? slots[lane] : INVALID_SLOT;
        float d = INFINITY;
        uint32_t id = 0xffffffffu;
        if (e != INVALID_SLOT) { d = node_distance<P>(g, us, 0.0, e); id = e; }
        float dv = 0.0f;
        if (lane == 0) dv = node_distance<P>(g, us, 0.0, v);
        dv = __shfl_sync(FULL, dv, 0);
        w.evals += nc + 1;
        warp_sort32<uint32_t>(d, id, lane);                         // ascending; empty lanes (INF, ~0) last
        const int rank = __popc(__ballot_sync(FULL, id != 0xffffffffu && pair_less<uint32_t>(d, id, dv, v)));
        uint32_t kid;
        float kd;
        const int kept = select_heuristic_warp<P>(g, w, cs, nc + 1, g.m2, lane,
                                                  [&](int j, uint32_t &c, float &dc) {
                                                      const int src = j < rank ? j : j - 1;
                                                      const uint32_t sc = __shfl_sync(FULL, id, src & 31);
                                                      const float sd = __shfl_sync(FULL, d, src & 31);
                                                      c = j == rank ? v : sc;
                                                      dc = j == rank ? dv : sd;
                                                  }, kid, kd);
        __syncwarp();
        if (lane < g.m2) slots[lane] = lane < kept ? kid : INVALID_SLOT;
        __syncwarp();
        nc = kept;
    }
    if (lane == 0) {
        g.cnt[(size_t) u * HNSW_MAX_LEVEL + lev] = (int16_t) nc;
        if (w.evals) atomicAdd(evals, (unsigned long long) w.evals);
    }
}

__global__ void hnsw_backlink_count_kernel(const unsigned long long *req, unsigned int nreq, HnswGraph g)
{
    const unsigned int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= nreq) return;
    const unsigned long long key = req[i] >> 32;
    if (i + 1 < nreq && (req[i + 1] >> 32) == key) return;      // not the last request of its segment
    int seg = 1;
    while (seg <= g.m2 && i >= (unsigned) seg && (req[i - seg] >> 32) == key) seg++;
    const uint32_t u = (uint32_t) (key >> 4);
    const int lev = (int) (key & 15);
    if (!nbr_slots(g, u, lev)) return;
    const int c = nbr_count(g, u, lev) + seg;
    g.cnt[(size_t) u * HNSW_MAX_LEVEL + lev] = (int16_t) (c > g.m2 ? g.m2 : c);
}

__global__ void pad_rows_kernel(const float *__restrict__ rows, int64_t n, int dim, int dimp, float *__restrict__ out)
{
    const int64_t t = (int64_t) blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n * dimp) return;
    const int64_t r = t / dimp;
    const int j = (int) (t - r * dimp);
    out[t] = j < dim ? rows[(size_t) r * dim + j] : 0.0f;
}

__global__ void hnsw_iota_kernel(int64_t *out, int64_t n)
{
    const int64_t i = (int64_t) blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = i;
}

// per-node sum of squares in hnswComputeDistance's cosine arithmetic (f32 product, f64 sum)
__global__ void hnsw_norms_kernel(const float *__restrict__ vec, int64_t n, int dim, int dimp, double *__restrict__ out)
{
    const int64_t r = (int64_t) blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= n) return;
    double acc = 0.0;
    for (int j = 0; j < dim; j++) Arith<NDB_COSINE, NDB_ARITH_HNSW>::nstep(acc, vec[(size_t) r * dimp + j]);
    out[r] = acc;
}

}  // namespace ndb

struct ndb_b200_hnsw {
    int dim = 0, dimp = 0, m = 16, efc = 200, efs = 64, metric = NDB_L2;
    int64_t n = 0;
    uint32_t entry = INVALID_SLOT;
    int entry_level = -1;
    DevBuf vec, ids, level, cnt, nbr0, upper_off, upper, vnorm;
    bool vnorm_ok = false;
    std::vector<int> h_level;
    std::vector<int64_t> h_upper_off;
    DevBuf bits, vlist, evals, req, req_sorted, nreq, cub_tmp, qbuf, outd, outi;
    int64_t last_evals = 0;
    int64_t bits_words = 0;
    int bits_warps = 0;
    int select_mode = NDB_HNSW_SELECT_CLOSEST;
};

namespace ndb {

static HnswGraph graph_of(const ndb_b200_hnsw *h)
{
    HnswGraph g;
    g.vec = h->vec.as<float>();
    g.ids = h->ids.as<int64_t>();
    g.level = h->level.as<int>();
    g.cnt = h->cnt.as<int16_t>();
    g.nbr0 = h->nbr0.as<uint32_t>();
    g.upper_off = h->upper_off.as<int64_t>();
    g.upper = h->upper.as<uint32_t>();
    g.vnorm = h->vnorm.as<double>();
    g.n = h->n;
    g.dim = h->dim;
    g.dimp = h->dimp;
    g.m2 = 2 * h->m;
    return g;
}

// shared-memory budget -> warps per CTA and grid for the warp-per-query kernels
// visited hash set of a search: ~34 * ef nodes are touched at M = 16 (C3: 1 364 at ef = 40); sized for a load of 7/8 at
// most, 4096 slots (16 KB per warp) at most -- beyond that the search moves to the global bitset on the fly
static int hnsw_hset_slots(int ef)
{
    static const int off = [] { const char *e = getenv("NDB_HNSW_VISITED_GLOBAL"); return e && atoi(e) ? 1 : 0; }();
    if (off) return 0;
    int cap = 1024;
    while (cap < 4096 && cap * 7 / 8 < ef * 36) cap <<= 1;
    return cap;
}

static int hnsw_launch_shape(const ndb_b200_hnsw *h, int ef, int *wpb, int *grid, size_t *smem, bool staging = false, int hcap = 0)
{
    const size_t per_warp = (size_t) h->dimp * 4 * (staging ? 2 : 1) + (size_t) ef * 12 + (size_t) hcap * 4;
    const size_t limit = ctx().smem_optin ? ctx().smem_optin - 2048 : 200 * 1024;
    NDB_REQUIRE(per_warp <= limit, NDB_B200_EINVAL, "hnsw: dim=%d with ef=%d does not fit in shared memory", h->dim, ef);
    int w = (int) std::min<size_t>(8, (limit / 2) / per_warp);     // aim for >= 2 CTAs per SM
    if (w < 1) w = 1;
    *wpb = w;
    *smem = per_warp * w;
    int ctas_per_sm = (int) std::min<size_t>(8, limit / *smem);
    if (ctas_per_sm * w > 32) ctas_per_sm = std::max(1, 32 / w);
    *grid = ctx().sm_count * ctas_per_sm;
    return NDB_B200_OK;
}

static int hnsw_scratch(ndb_b200_hnsw *h, int total_warps, cudaStream_t s)
{
    const int64_t words = (h->n + 31) / 32 + 1;
    const int vcap = 16384;
    if (words != h->bits_words || total_warps > h->bits_warps) {
        NDB_CHECK(h->bits.reserve((size_t) total_warps * words * 4));
        NDB_CHECK(h->vlist.reserve((size_t) total_warps * vcap * 4));
        NDB_CUDA(cudaMemsetAsync(h->bits.p, 0, (size_t) total_warps * words * 4, s));
        h->bits_words = words;
        h->bits_warps = total_warps;
    }
    NDB_CHECK(h->evals.reserve(64));
    return NDB_B200_OK;
}

static int hnsw_alloc_nodes(ndb_b200_hnsw *h, const float *rows, const int64_t *ids, int64_t n, const int *levels,
                            cudaStream_t s)
{
    NDB_REQUIRE(n < (int64_t) 0x7ffffff0, NDB_B200_EINVAL, "hnsw: too many nodes");
    h->n = n;
    h->h_level.assign(levels, levels + n);
    h->h_upper_off.assign(n + 1, 0);
    for (int64_t i = 0; i < n; i++) {
        NDB_REQUIRE(levels[i] >= 0 && levels[i] < HNSW_MAX_LEVEL, NDB_B200_EINVAL, "hnsw: level %d out of range", levels[i]);
        h->h_upper_off[i + 1] = h->h_upper_off[i] + (int64_t) levels[i] * 2 * h->m;
    }
    const int64_t up = h->h_upper_off[n];
    DevBuf tmp;
    NDB_CHECK(tmp.reserve((size_t) n * h->dim * 4));
    NDB_CHECK(h->vec.reserve((size_t) n * h->dimp * 4));
    NDB_CHECK(h->ids.reserve((size_t) n * 8));
    NDB_CHECK(h->level.reserve((size_t) n * 4));
    NDB_CHECK(h->cnt.reserve((size_t) n * HNSW_MAX_LEVEL * 2));
    NDB_CHECK(h->nbr0.reserve((size_t) n * 2 * h->m * 4));
    NDB_CHECK(h->upper_off.reserve((size_t) (n + 1) * 8));
    NDB_CHECK(h->upper.reserve((size_t) (up ? up : 1) * 4));
    NDB_CUDA(cudaMemcpyAsync(tmp.p, rows, (size_t) n * h->dim * 4, cudaMemcpyHostToDevice, s));
    pad_rows_kernel<<<(unsigned) ((n * h->dimp + 255) / 256), 256, 0, s>>>(tmp.as<float>(), n, h->dim, h->dimp, h->vec.as<float>());
    count_launch();
    if (ids) NDB_CUDA(cudaMemcpyAsync(h->ids.p, ids, (size_t) n * 8, cudaMemcpyHostToDevice, s));
    else { hnsw_iota_kernel<<<(unsigned) ((n + 255) / 256), 256, 0, s>>>(h->ids.as<int64_t>(), n); count_launch(); }
    NDB_CUDA(cudaMemcpyAsync(h->level.p, levels, (size_t) n * 4, cudaMemcpyHostToDevice, s));
    NDB_CUDA(cudaMemcpyAsync(h->upper_off.p, h->h_upper_off.data(), (size_t) (n + 1) * 8, cudaMemcpyHostToDevice, s));
    NDB_CUDA(cudaMemsetAsync(h->cnt.p, 0, (size_t) n * HNSW_MAX_LEVEL * 2, s));
    NDB_CUDA(cudaMemsetAsync(h->nbr0.p, 0xff, (size_t) n * 2 * h->m * 4, s));
    NDB_CUDA(cudaMemsetAsync(h->upper.p, 0xff, (size_t) (up ? up : 1) * 4, s));
    NDB_CUDA(cudaStreamSynchronize(s));
    h->vnorm_ok = false;
    h->bits_words = 0;
    return NDB_B200_OK;
}

static int hnsw_norms(ndb_b200_hnsw *h, cudaStream_t s)
{
    if (h->vnorm_ok) return NDB_B200_OK;
    NDB_CHECK(h->vnorm.reserve((size_t) (h->n ? h->n : 1) * 8));
    if (h->n) {
        hnsw_norms_kernel<<<(unsigned) ((h->n + 127) / 128), 128, 0, s>>>(h->vec.as<float>(), h->n, h->dim, h->dimp, h->vnorm.as<double>());
        count_launch();
        NDB_CUDA(cudaGetLastError());
    }
    h->vnorm_ok = true;
    return NDB_B200_OK;
}

template <class P> static int run_search(const HnswSearchArgs &a, int grid, int wpb, size_t smem, cudaStream_t s)
{
    auto kern = hnsw_search_kernel<P>;
    NDB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) smem));
    kern<<<grid, wpb * 32, smem, s>>>(a);
    count_launch();
    NDB_CUDA(cudaGetLastError());
    return NDB_B200_OK;
}

}  // namespace ndb

extern "C" {

int ndb_b200_hnsw_create(int dim, int m, int ef_construction, int ef_search, int metric, ndb_b200_hnsw **out)
{
    NDB_CHECK(require_init());
    NDB_REQUIRE(out && dim > 0 && dim <= 16000, NDB_B200_EINVAL, "hnsw_create: dim must be 1..16000");
    // hnswoptions validation (hnsw_am.c:805-864): m 2..128, ef 4..10000, ef >= m
    NDB_REQUIRE(m >= 2 && m <= 128, NDB_B200_EINVAL, "hnsw: m must be between 2 and 128");
    NDB_REQUIRE(ef_construction >= 4 && ef_construction <= 10000 && ef_construction >= m, NDB_B200_EINVAL,
                "hnsw: ef_construction must be between 4 and 10000 and >= m");
    NDB_REQUIRE(ef_search >= 4 && ef_search <= 10000 && ef_search >= m, NDB_B200_EINVAL,
                "hnsw: ef_search must be between 4 and 10000 and >= m");
    NDB_REQUIRE(metric >= NDB_L2 && metric <= NDB_IP, NDB_B200_EINVAL, "hnsw_create: unknown metric %d", metric);
    ndb_b200_hnsw *h = new ndb_b200_hnsw();
    h->dim = dim;
    h->dimp = round_up(dim, 4);
    h->m = m;
    h->efc = ef_construction;
    h->efs = ef_search;
    h->metric = metric;
    *out = h;
    return NDB_B200_OK;
}

void ndb_b200_hnsw_free(ndb_b200_hnsw *h)
{
    if (!h) return;
    if (ctx().initialized) { cudaSetDevice(ctx().device); cudaStreamSynchronize(ctx().stream); }
    delete h;
}

int64_t ndb_b200_hnsw_size(const ndb_b200_hnsw *h) { return h ? h->n : 0; }
int64_t ndb_b200_hnsw_last_evals(const ndb_b200_hnsw *ch)
{
    ndb_b200_hnsw *h = const_cast<ndb_b200_hnsw *>(ch);
    if (!h || !h->evals.p) return 0;
    unsigned long long v = 0;
    if (cudaMemcpy(&v, h->evals.p, 8, cudaMemcpyDeviceToHost) != cudaSuccess) return -1;
    return (int64_t) v;
}

int ndb_b200_hnsw_set_select(ndb_b200_hnsw *h, int select_mode)
{
    NDB_REQUIRE(h && (select_mode == NDB_HNSW_SELECT_CLOSEST || select_mode == NDB_HNSW_SELECT_HEURISTIC), NDB_B200_EINVAL,
                "hnsw_set_select: bad argument");
    h->select_mode = select_mode;
    return NDB_B200_OK;
}

int ndb_b200_hnsw_load_graph(ndb_b200_hnsw *h, const float *rows, const int64_t *ids, int64_t n, const int *levels,
                             const uint32_t *nbr0, const int16_t *cnt, const int64_t *upper_off, const uint32_t *upper,
                             uint32_t entry_point, int entry_level)
{
    NDB_CHECK(require_init());
    NDB_REQUIRE(h && rows && levels && nbr0 && cnt && upper_off && n > 0, NDB_B200_EINVAL, "hnsw_load_graph: NULL or empty input");
    NDB_REQUIRE(find_nonfinite(rows, n * h->dim) < 0, NDB_B200_EVECTOR, "hnsw_load_graph: NaN/Inf in rows");
    cudaStream_t s = ctx().stream;
    NDB_CHECK(hnsw_alloc_nodes(h, rows, ids, n, levels, s));
    for (int64_t i = 0; i <= n; i++)
        NDB_REQUIRE(upper_off[i] == h->h_upper_off[i], NDB_B200_EINVAL, "hnsw_load_graph: upper_off[%lld] does not match the levels", (long long) i);
    NDB_CUDA(cudaMemcpyAsync(h->nbr0.p, nbr0, (size_t) n * 2 * h->m * 4, cudaMemcpyHostToDevice, s));
    NDB_CUDA(cudaMemcpyAsync(h->cnt.p, cnt, (size_t) n * HNSW_MAX_LEVEL * 2, cudaMemcpyHostToDevice, s));
    if (upper_off[n] > 0) {
        NDB_REQUIRE(upper, NDB_B200_EINVAL, "hnsw_load_graph: upper slots missing");
        NDB_CUDA(cudaMemcpyAsync(h->upper.p, upper, (size_t) upper_off[n] * 4, cudaMemcpyHostToDevice, s));
    }
    NDB_CUDA(cudaStreamSynchronize(s));
    h->entry = entry_point;
    h->entry_level = entry_level;
    return NDB_B200_OK;
}

int ndb_b200_hnsw_export_graph(const ndb_b200_hnsw *h, int *levels, uint32_t *nbr0, int16_t *cnt, int64_t *upper_off,
                               uint32_t *upper, int64_t upper_cap, uint32_t *entry_point, int *entry_level)
{
    NDB_CHECK(require_init());
    NDB_REQUIRE(h && h->n > 0, NDB_B200_ESTATE, "hnsw_export_graph: empty index");
    const int64_t n = h->n, up = h->h_upper_off[n];
    NDB_REQUIRE(!upper || upper_cap >= up, NDB_B200_EINVAL, "hnsw_export_graph: upper buffer too small (%lld < %lld)",
                (long long) upper_cap, (long long) up);
    NDB_CUDA(cudaStreamSynchronize(ctx().stream));
    if (levels) memcpy(levels, h->h_level.data(), (size_t) n * 4);
    if (upper_off) memcpy(upper_off, h->h_upper_off.data(), (size_t) (n + 1) * 8);
    if (nbr0) NDB_CUDA(cudaMemcpy(nbr0, h->nbr0.p, (size_t) n * 2 * h->m * 4, cudaMemcpyDeviceToHost));
    if (cnt) NDB_CUDA(cudaMemcpy(cnt, h->cnt.p, (size_t) n * HNSW_MAX_LEVEL * 2, cudaMemcpyDeviceToHost));
    if (upper && up) NDB_CUDA(cudaMemcpy(upper, h->upper.p, (size_t) up * 4, cudaMemcpyDeviceToHost));
    if (entry_point) *entry_point = h->entry;
    if (entry_level) *entry_level = h->entry_level;
    return NDB_B200_OK;
}

// the graph does not shard, so one rank builds (hnswbuild is serial in the reference too,
// hnsw_am.c:343-415) and the finished graph -- node vectors, levels, neighbour slots -- reaches the
// replicas with ncclBroadcast over NVLink (C3: ~3.2 GB).  Collective: every rank of the communicator calls it.
int ndb_b200_hnsw_broadcast(ndb_b200_hnsw *h, int root)
{
    NDB_CHECK(require_init());
    NDB_REQUIRE(h && root >= 0 && root < comm_nranks(), NDB_B200_EINVAL, "hnsw_broadcast: bad argument");
    if (comm_nranks() == 1) return NDB_B200_OK;
    cudaStream_t s = ctx().stream;
    const bool am_root = comm_rank() == root;
    NDB_REQUIRE(!am_root || h->n > 0, NDB_B200_ESTATE, "hnsw_broadcast: the root holds no graph");
    struct Hdr { int64_t n, up; uint32_t entry; int entry_level, dim, m; } hdr = {0, 0, INVALID_SLOT, -1, h->dim, h->m};
    if (am_root) { hdr.n = h->n; hdr.up = h->h_upper_off[h->n]; hdr.entry = h->entry; hdr.entry_level = h->entry_level; }
    DevBuf dh;
    NDB_CHECK(dh.reserve(sizeof(Hdr)));
    NDB_CUDA(cudaMemcpyAsync(dh.p, &hdr, sizeof(Hdr), cudaMemcpyHostToDevice, s));
    NDB_CHECK(comm_broadcast(dh.p, sizeof(Hdr), root, s));
    Hdr got;
    NDB_CUDA(cudaMemcpyAsync(&got, dh.p, sizeof(Hdr), cudaMemcpyDeviceToHost, s));
    NDB_CUDA(cudaStreamSynchronize(s));
    NDB_REQUIRE(got.dim == h->dim && got.m == h->m, NDB_B200_EDIM, "hnsw_broadcast: the root's index has dim %d m %d, this handle dim %d m %d",
                got.dim, got.m, h->dim, h->m);
    const int64_t n = got.n, up = got.up;
    if (!am_root) {
        h->n = n;
        NDB_CHECK(h->vec.reserve((size_t) n * h->dimp * 4));
        NDB_CHECK(h->ids.reserve((size_t) n * 8));
        NDB_CHECK(h->level.reserve((size_t) n * 4));
        NDB_CHECK(h->cnt.reserve((size_t) n * HNSW_MAX_LEVEL * 2));
        NDB_CHECK(h->nbr0.reserve((size_t) n * 2 * h->m * 4));
        NDB_CHECK(h->upper_off.reserve((size_t) (n + 1) * 8));
        NDB_CHECK(h->upper.reserve((size_t) (up ? up : 1) * 4));
    }
    NDB_CHECK(comm_broadcast(h->vec.p, (size_t) n * h->dimp * 4, root, s));
    NDB_CHECK(comm_broadcast(h->ids.p, (size_t) n * 8, root, s));
    NDB_CHECK(comm_broadcast(h->level.p, (size_t) n * 4, root, s));
    NDB_CHECK(comm_broadcast(h->cnt.p, (size_t) n * HNSW_MAX_LEVEL * 2, root, s));
    NDB_CHECK(comm_broadcast(h->nbr0.p, (size_t) n * 2 * h->m * 4, root, s));
    NDB_CHECK(comm_broadcast(h->upper_off.p, (size_t) (n + 1) * 8, root, s));
    if (up) NDB_CHECK(comm_broadcast(h->upper.p, (size_t) up * 4, root, s));
    if (!am_root) {
        h->h_level.resize(n);
        h->h_upper_off.resize(n + 1);
        NDB_CUDA(cudaMemcpyAsync(h->h_level.data(), h->level.p, (size_t) n * 4, cudaMemcpyDeviceToHost, s));
        NDB_CUDA(cudaMemcpyAsync(h->h_upper_off.data(), h->upper_off.p, (size_t) (n + 1) * 8, cudaMemcpyDeviceToHost, s));
        h->entry = got.entry;
        h->entry_level = got.entry_level;
        h->vnorm_ok = false;
        h->bits_words = 0;
    }
    NDB_CUDA(cudaStreamSynchronize(s));
    return NDB_B200_OK;
}

int ndb_b200_hnsw_build(ndb_b200_hnsw *h, const float *rows, const int64_t *ids, int64_t n, const int *levels,
                        unsigned seed, int batch)
{
    NDB_CHECK(require_init());
    NDB_REQUIRE(h && rows && n > 0, NDB_B200_EINVAL, "hnsw_build: NULL or empty input");
    NDB_REQUIRE(find_nonfinite(rows, n * h->dim) < 0, NDB_B200_EVECTOR, "hnsw_build: NaN/Inf in rows");
    NDB_REQUIRE(h->efc <= 2048, NDB_B200_EINVAL, "hnsw_build: ef_construction > 2048 is not supported");
    cudaStream_t s = ctx().stream;
    std::vector<int> drawn;
    if (!levels) {
        // hnswGetRandomLevel (:1143-1161): libc random(), ml = 0.36, capped at 15
        drawn.resize(n);
        srandom(seed);
        for (int64_t i = 0; i < n; i++) {
            double r = (double) random() / (double) RAND_MAX;
            while (r == 0.0) r = (double) random() / (double) RAND_MAX;
            int lv = (int) (-log(r) * 0.36f);
            drawn[i] = lv > HNSW_MAX_LEVEL - 1 ? HNSW_MAX_LEVEL - 1 : (lv < 0 ? 0 : lv);
        }
        levels = drawn.data();
    }
    NDB_CHECK(hnsw_alloc_nodes(h, rows, ids, n, levels, s));
    int wpb, grid;
    size_t smem;
    const bool heuristic = h->select_mode == NDB_HNSW_SELECT_HEURISTIC;
    NDB_REQUIRE(!heuristic || 2 * h->m <= 32, NDB_B200_EINVAL, "hnsw_build: heuristic selection supports m <= 16 (m = %d)", h->m);
    NDB_CHECK(hnsw_launch_shape(h, h->efc, &wpb, &grid, &smem, heuristic));
    NDB_CHECK(hnsw_scratch(h, grid * wpb, s));
    NDB_CUDA(cudaFuncSetAttribute(hnsw_build_search_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) smem));
    const size_t sel_smem = (size_t) 4 * 2 * h->dimp * 4;             // hnsw_backlink_select_kernel: 4 warps per CTA
    if (heuristic)
        NDB_CUDA(cudaFuncSetAttribute(hnsw_backlink_select_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) sel_smem));
    const int64_t max_batch = batch > 0 ? batch : 4096;
    const size_t req_cap = (size_t) max_batch * HNSW_MAX_LEVEL * h->m;
    NDB_CHECK(h->req.reserve(req_cap * 8));
    NDB_CHECK(h->req_sorted.reserve(req_cap * 8));
    NDB_CHECK(h->nreq.reserve(64));
    NDB_CUDA(cudaMemsetAsync(h->evals.p, 0, 8, s));
    size_t tmp_bytes = 0;
    NDB_CUDA(cub::DeviceRadixSort::SortKeys(nullptr, tmp_bytes, h->req.as<unsigned long long>(), h->req_sorted.as<unsigned long long>(),
                                            (int) req_cap, 0, 64, s));
    NDB_CHECK(h->cub_tmp.reserve(tmp_bytes));

    uint32_t entry = INVALID_SLOT;
    int entry_level = -1;
    int64_t done = 0;
    while (done < n) {
        // batch policy: batch == 0 grows with the graph so that new nodes missing each other stay
        // a small fraction of it; batch == 1 is the sequential algorithm
        int64_t b = batch > 0 ? batch : std::max<int64_t>(1, std::min<int64_t>(max_batch, done / 16));
        if (b > n - done) b = n - done;
        HnswBuildArgs a;
        a.g = graph_of(h);
        a.first = (uint32_t) done;
        a.count = (uint32_t) b;
        a.m = h->m;
        a.efc = h->efc;
        a.entry = entry;
        a.entry_level = entry_level;
        a.bits = h->bits.as<uint32_t>();
        a.vlist = h->vlist.as<uint32_t>();
        a.words = h->bits_words;
        a.vcap = 16384;
        a.req = h->req.as<unsigned long long>();
        a.nreq = h->nreq.as<unsigned int>();
        a.evals = h->evals.as<unsigned long long>();
        a.heuristic = heuristic ? 1 : 0;
        unsigned int nreq = 0;
        if (entry != INVALID_SLOT) {
            NDB_CUDA(cudaMemsetAsync(h->nreq.p, 0, 4, s));
            const int g = (int) std::min<int64_t>(grid, (b + wpb - 1) / wpb);
            hnsw_build_search_kernel<<<g, wpb * 32, smem, s>>>(a);
            count_launch();
            NDB_CUDA(cudaGetLastError());
            NDB_CUDA(cudaMemcpyAsync(&nreq, h->nreq.p, 4, cudaMemcpyDeviceToHost, s));
            NDB_CUDA(cudaStreamSynchronize(s));
        }
        if (nreq) {
            size_t tb = tmp_bytes;
            NDB_CUDA(cub::DeviceRadixSort::SortKeys(h->cub_tmp.p, tb, h->req.as<unsigned long long>(),
                                                    h->req_sorted.as<unsigned long long>(), (int) nreq, 0, 64, s));
            const unsigned gb = (nreq + 255) / 256;
            if (heuristic) {
                hnsw_backlink_select_kernel<<<(nreq + 3) / 4, 128, sel_smem, s>>>(h->req_sorted.as<unsigned long long>(), nreq, a.g,
                                                                                  h->evals.as<unsigned long long>());
            } else {
                hnsw_backlink_write_kernel<<<gb, 256, 0, s>>>(h->req_sorted.as<unsigned long long>(), nreq, a.g);
                hnsw_backlink_count_kernel<<<gb, 256, 0, s>>>(h->req_sorted.as<unsigned long long>(), nreq, a.g);
            }
            count_launch(5);
            NDB_CUDA(cudaGetLastError());
        }
        // Step 6 (:2649-2666) for the nodes of the batch, in order
        for (int64_t i = done; i < done + b; i++)
            if (entry == INVALID_SLOT || levels[i] > entry_level) { entry = (uint32_t) i; entry_level = levels[i]; }
        done += b;
    }
    NDB_CUDA(cudaStreamSynchronize(s));
    h->entry = entry;
    h->entry_level = entry_level;
    return NDB_B200_OK;
}

int ndb_b200_hnsw_search_dev(ndb_b200_hnsw *h, const float *Q_dev, int nq, int strategy, int ef, int k, int mode,
                             float *dist_dev, int64_t *ids_dev, void *stream)
{
    NDB_CHECK(require_init());
    NDB_REQUIRE(h && Q_dev && dist_dev && ids_dev && nq > 0, NDB_B200_EINVAL, "hnsw_search: NULL or empty input");
    NDB_REQUIRE(k >= 1 && ef >= 1 && ef <= 4096, NDB_B200_EINVAL, "hnsw_search: ef must be 1..4096 and k >= 1");
    NDB_REQUIRE(strategy >= NDB_L2 && strategy <= NDB_IP, NDB_B200_EINVAL, "hnsw: unsupported distance strategy %d", strategy);
    NDB_REQUIRE(mode == NDB_HNSW_LITERAL || mode == NDB_HNSW_BESTFIRST, NDB_B200_EINVAL, "hnsw_search: unknown mode %d", mode);
    cudaStream_t s = stream ? (cudaStream_t) stream : ctx().stream;
    int wpb, grid;
    size_t smem;
    const int hcap = hnsw_hset_slots(ef);
    NDB_CHECK(hnsw_launch_shape(h, ef, &wpb, &grid, &smem, false, hcap));
    if ((int64_t) grid * wpb > nq) grid = (nq + wpb - 1) / wpb;
    NDB_CHECK(hnsw_scratch(h, ctx().sm_count * 32, s));
    if (strategy == NDB_COSINE) NDB_CHECK(hnsw_norms(h, s));
    NDB_CUDA(cudaMemsetAsync(h->evals.p, 0, 8, s));
    HnswSearchArgs a;
    a.g = graph_of(h);
    a.Q = Q_dev;
    a.nq = nq; a.ef = ef; a.k = k; a.mode = mode;
    a.entry = h->entry;
    a.entry_level = h->entry_level;
    a.bits = h->bits.as<uint32_t>();
    a.vlist = h->vlist.as<uint32_t>();
    a.words = h->bits_words;
    a.vcap = 16384;
    a.hcap = hcap;
    a.out_dist = dist_dev;
    a.out_ids = ids_dev;
    a.evals = h->evals.as<unsigned long long>();
    Context &c = ctx();
    if (c.timing) NDB_CUDA(cudaEventRecord(c.ev0, s));
    int rc;
    if (strategy == NDB_L2) rc = run_search<Arith<NDB_L2, NDB_ARITH_HNSW>>(a, grid, wpb, smem, s);
    else if (strategy == NDB_COSINE) rc = run_search<Arith<NDB_COSINE, NDB_ARITH_HNSW>>(a, grid, wpb, smem, s);
    else rc = run_search<Arith<NDB_IP, NDB_ARITH_HNSW>>(a, grid, wpb, smem, s);
    if (c.timing && rc == NDB_B200_OK) {
        NDB_CUDA(cudaEventRecord(c.ev1, s));
        c.last_ms = -1.0;
        c.last_bytes = 0.0;
        c.last_evals = 0;
        c.stats_src = nullptr;
    }
    return rc;
}

int ndb_b200_hnsw_search(ndb_b200_hnsw *h, const float *Q, int nq, int strategy, int ef, int k, int mode, float *dist,
                         int64_t *ids)
{
    NDB_CHECK(require_init());
    NDB_REQUIRE(h && Q && dist && ids && nq > 0, NDB_B200_EINVAL, "hnsw_search: NULL or empty input");
    cudaStream_t s = ctx().stream;
    const size_t qb = (size_t) nq * h->dim * 4, m = (size_t) nq * k;
    NDB_CHECK(h->qbuf.reserve(qb));
    NDB_CHECK(h->outd.reserve(m * 4));
    NDB_CHECK(h->outi.reserve(m * 8));
    NDB_CUDA(cudaMemcpyAsync(h->qbuf.p, Q, qb, cudaMemcpyHostToDevice, s));
    NDB_CHECK(validate_begin(h->qbuf.as<float>(), (int64_t) nq * h->dim, s));
    NDB_CHECK(ndb_b200_hnsw_search_dev(h, h->qbuf.as<float>(), nq, strategy, ef, k, mode, h->outd.as<float>(),
                                       h->outi.as<int64_t>(), s));
    NDB_CUDA(cudaMemcpyAsync(dist, h->outd.p, m * 4, cudaMemcpyDeviceToHost, s));
    NDB_CUDA(cudaMemcpyAsync(ids, h->outi.p, m * 8, cudaMemcpyDeviceToHost, s));
    NDB_CUDA(cudaStreamSynchronize(s));
    NDB_REQUIRE(validate_end() < 0, NDB_B200_EVECTOR, "hnsw_search: NaN/Inf in query");
    return NDB_B200_OK;
}

// ---- relation loader: meta page + one HnswNodeData item per 8 KB page (hnsw_am.c:108-181) -----
// node id = block - 1; neighbour slots hold block numbers and are rebased the same way.
// hnsw_knn_search_gpu(index_name, query, k, ef_search default 100) (src/gpu/common/gpu_sql.c:498-930): the SQL
// function's argument checks (:556-580) and its result over the resident graph; strategy 1 as the access method.
int ndb_b200_hnsw_knn_search_gpu(ndb_b200_hnsw *h, const float *query, int dim, int k, int ef_search, int64_t *ids, float *dist,
                                 int *nresults)
{
    NDB_CHECK(require_init());
    NDB_REQUIRE(h, NDB_B200_EINVAL, "hnsw_knn_search_gpu: index cannot be NULL");
    NDB_REQUIRE(query && ids && dist && nresults, NDB_B200_EINVAL, "hnsw_knn_search_gpu: query vector cannot be NULL");
    NDB_REQUIRE(k > 0 && k <= 10000, NDB_B200_EINVAL, "hnsw_knn_search_gpu: k must be between 1 and 10000");
    NDB_REQUIRE(ef_search > 0 && ef_search <= 10000, NDB_B200_EINVAL, "hnsw_knn_search_gpu: ef_search must be between 1 and 10000");
    NDB_REQUIRE(dim > 0 && dim == h->dim, NDB_B200_EDIM, "hnsw_knn_search_gpu: invalid query dimension %d", dim);
    NDB_REQUIRE(k <= 128, NDB_B200_EINVAL, "hnsw_knn_search_gpu: k > 128 is not supported by this library (k = %d)", k);
    const int ef = ef_search < k ? k : ef_search;
    NDB_CHECK(ndb_b200_hnsw_search(h, query, 1, 1, ef, k, NDB_HNSW_BESTFIRST, dist, ids));
    int n = 0;
    while (n < k && ids[n] >= 0) n++;
    *nresults = n;
    return NDB_B200_OK;
}

int ndb_b200_hnsw_load_relation(ndb_b200_hnsw *h, const void *blocks, uint32_t nblocks)
{
    using namespace ndb::pg;
    NDB_CHECK(require_init());
    NDB_REQUIRE(h && blocks && nblocks >= 2, NDB_B200_EINVAL, "hnsw_load_relation: need the meta block and at least one node");
    HnswMeta meta;
    memcpy(&meta, page_at(blocks, 0) + PAGE_HEADER, sizeof(meta));
    NDB_REQUIRE(meta.magic == HNSW_MAGIC, NDB_B200_EINVAL, "hnsw_load_relation: bad magic 0x%08x in the meta page", meta.magic);
    NDB_REQUIRE(meta.m == h->m, NDB_B200_EINVAL, "hnsw_load_relation: relation m=%d, handle m=%d", (int) meta.m, h->m);
    const int64_t n = (int64_t) nblocks - 1;
    const int dim = h->dim, m2 = 2 * h->m;
    std::vector<float> rows((size_t) n * dim);
    std::vector<int64_t> ids(n);
    std::vector<int> levels(n);
    std::vector<uint32_t> nbr0((size_t) n * m2, INVALID_SLOT);
    std::vector<int16_t> cnt((size_t) n * HNSW_MAX_LEVEL, 0);
    std::vector<int64_t> uoff(n + 1, 0);
    std::vector<uint32_t> upper;
    auto rebase = [&](uint32_t blk) { return (blk == INVALID_BLOCK || blk == 0 || blk >= nblocks) ? INVALID_SLOT : blk - 1; };
    for (int64_t i = 0; i < n; i++) {
        const uint8_t *page = page_at(blocks, (uint32_t) (i + 1));
        NDB_REQUIRE(max_offset(page) >= 1, NDB_B200_EINVAL, "hnsw_load_relation: block %lld holds no node", (long long) (i + 1));
        uint32_t lo = 0, len = 0;
        // always FirstOffsetNumber (:70-80); the header is only read once the item is known to lie inside the page
        NDB_REQUIRE(checked_item(page, 1, HNSW_NODE_HDR + (uint32_t) dim * 4, &lo, &len) == 1, NDB_B200_EINVAL,
                    "hnsw_load_relation: block %lld has no valid node item (line pointer outside the page or too short)", (long long) (i + 1));
        const uint8_t *node = page + lo;
        int32_t level;
        int16_t ndim;
        memcpy(&level, node + 8, 4);
        memcpy(&ndim, node + 12, 2);
        NDB_REQUIRE(level >= 0 && level < HNSW_MAX_LEVEL, NDB_B200_EINVAL, "hnsw_load_relation: invalid node level %d at block %lld", level, (long long) (i + 1));
        NDB_REQUIRE(ndim == dim, NDB_B200_EDIM, "hnsw_load_relation: node dim %d, handle dim %d", (int) ndim, dim);
        NDB_REQUIRE(len >= HNSW_NODE_HDR + (uint32_t) dim * 4 + (uint32_t) (level + 1) * m2 * 4, NDB_B200_EINVAL,
                    "hnsw_load_relation: node item at block %lld is truncated", (long long) (i + 1));
        ids[i] = tid_unpack(node);
        levels[i] = level;
        memcpy(&cnt[(size_t) i * HNSW_MAX_LEVEL], node + 14, 2 * HNSW_MAX_LEVEL);
        memcpy(&rows[(size_t) i * dim], node + HNSW_NODE_HDR, (size_t) dim * 4);
        const uint8_t *nb = node + HNSW_NODE_HDR + (size_t) dim * 4;
        uoff[i + 1] = uoff[i] + (int64_t) level * m2;
        for (int l = 0; l <= level; l++)
            for (int j = 0; j < m2; j++) {
                uint32_t v;
                memcpy(&v, nb + ((size_t) l * m2 + j) * 4, 4);
                if (l == 0) nbr0[(size_t) i * m2 + j] = rebase(v);
                else upper.push_back(rebase(v));
            }
    }
    if (upper.empty()) upper.push_back(INVALID_SLOT);
    const uint32_t entry = rebase(meta.entryPoint);
    return ndb_b200_hnsw_load_graph(h, rows.data(), ids.data(), n, levels.data(), nbr0.data(), cnt.data(), uoff.data(),
                                    upper.data(), entry, meta.entryLevel);
}

}  // extern "C"
