// ivf_cert.cuh -- certified selection for the tensor-core IVF search (NDB_ARITH_TENSOR).
//
// The tensor cores rank candidates by bf16 products; the reference ranks them by ivfComputeDistance
// (NeuronDB/src/index/ivf_am.c:1550-1592) in sequential fp32.  north_star's contract: ids identical to the
// reference wherever the distance gap exceeds the tolerance.  These kernels give more: the ids and distance
// bits of the fp32 path, always, because every answer is either CERTIFIED or recomputed exactly.
//
// What the scan leaves behind (tc_knn.cu).  Per query, partial lists of at most kc (key, row) entries, one per
// (probed list, segment, column half, replica), and the final value G of the shared per-query bound: the
// smallest kc-th key any full partial list ended with.  A candidate was dropped only (i) by the threshold
// test, at a moment when the threshold was >= G, or (ii) out of the tail of a full list, whose kc-th key is
// >= G.  Hence the entries with key <= G are the COMPLETE set of rows whose key is below G (up to the 12-bit
// packing of the keys, which the bound below absorbs).
//
// Certificate.  Take the best 32 entries by key; re-evaluate them with the reference's arithmetic;
// tau = the k-th smallest exact value so far.  Every row not re-evaluated has a key >= g, with
// g = the key of the first entry not re-evaluated (or G when there is none).  With x~ = bf16(x), q~ = bf16(q),
// ex = x~ - x, eq = q~ - q (the only approximation in the key; bf16 x bf16 products are exact in fp32):
//     L2      key = ||x~ - q~||^2          ||x - q||       >= sqrt(key) - ||ex|| - ||eq||
//     IP      key = -x~.q~                 -x.q            >= key - ||ex|| ||q~|| - ||x|| ||eq||
//     cosine  key = -x~.q~ / ||x~||        -x.q / ||x||    >= key - rho (||q~|| + ||q||) - kappa ||eq||,
//                                          rho = max ||ex|| / ||x~||, kappa = max ||x|| / ||x~||
// with the maxima of ||ex||, ||x||, rho, kappa over the stored rows taken at build time (TcStore::stats),
// ||eq||, ||q||, ||q~|| computed here, plus slack for fp32 accumulation ((dim + 8) 2^-23 of the magnitudes) and
// for the key packing (2^-10 relative).  If that lower bound exceeds tau the answer is certified: no row
// outside the re-evaluated set can enter the reference's top k.  Otherwise the query goes to the exact
// kernel below, which evaluates every row of its probed lists in the reference's arithmetic.
#pragma once
#include "cert_common.cuh"

namespace ndb {

// ---- lists: merge + certified re-rank (replaces the fixed k + 6 margin) ------------------------------
// gthr[q] = R, the RELAXED shared bound (cert_bound.cuh): every candidate the threshold test dropped had a key
// >= R; a full partial list also dropped candidates above its own kc-th key.  Hence every row that is in no
// partial list has a key >= min(R, smallest kc-th key of a full list) =: g_lists.
// KRC: candidates re-evaluated per query / 32 (1 for k <= 16; 2 for k <= 32).
template <class P, int METRIC, int KRC>
__global__ void __launch_bounds__(128) ivf_tc_finish_cert_kernel(
    const float *__restrict__ pdist, const uint32_t *__restrict__ pslot, const uint32_t *__restrict__ tc_src,
    const uint32_t *__restrict__ tc_row, const float *__restrict__ arena, const int64_t *__restrict__ ids,
    const float *__restrict__ Q, const uint32_t *__restrict__ probe, const uint32_t *__restrict__ pairpos,
    const uint32_t *__restrict__ item_off, const uint32_t *__restrict__ list_len, const float *__restrict__ gthr,
    const uint32_t *__restrict__ cnt, uint32_t rep_max, uint32_t split, int nq, int nprobe, int nlists, uint32_t segb, int dim, int kc, int k,
    const float *__restrict__ stats, float *__restrict__ out_dist, int64_t *__restrict__ out_ids,
    uint32_t *__restrict__ fb_list, float *__restrict__ fb_tau,
    unsigned long long *__restrict__ counters /* [0] queries sent to the exact kernel, [1] exact evaluations */)
{
    const int lane = threadIdx.x & 31;
    const int q = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (q >= nq) return;
    WarpTopK<KRC, uint32_t> cand;
    cand.init();
    const float R = gthr[q];
    const bool bounded = R < 1.0e38f;                  // "no bound": no partial list of this query ever filled
    float own_min = INFINITY;                           // smallest kc-th key of a full partial list
    uint32_t n_in = 0;                                  // entries met (warp-uniform)
    for (int r0 = 0; r0 < nprobe; r0 += 32) {
        uint32_t my_first = 0, my_nseg = 0, my_rep = 1;
        if (r0 + lane < nprobe) {
            const size_t p = (size_t) q * nprobe + r0 + lane;
            const uint32_t l = probe[p];
            if (l < (uint32_t) nlists) {
                const uint32_t len = list_len[l];
                if (len) {
                    const uint32_t v = ivf_vlist(l, (int64_t) p, (uint32_t) nprobe, split);
                    my_rep = ivf_rep(cnt[v], rep_max);
                    const uint32_t pos = pairpos[p] * my_rep;
                    my_nseg = ivf_nseg(len, segb);
                    my_first = (item_off[v] + (pos / TC_M) * my_nseg) * (2 * TC_M) + (pos % TC_M) * 2;
                }
            }
        }
        const int nr = min(32, nprobe - r0);
        for (int r = 0; r < nr; r++) {
            const uint32_t nseg = __shfl_sync(FULL, my_nseg, r), first = __shfl_sync(FULL, my_first, r);
            const int nent = (int) __shfl_sync(FULL, my_rep, r) * 2 * kc;
            // A (probe, segment) unit holds nent = 32 entries (two column halves) or 128 (replicated queries).  The walk is
            // a latency chain -- a C4 query has ~290 units, one L2 / DRAM round trip each -- so units of 32 entries are
            // fetched four at a time (consecutive segments), units of 64 two at a time.
            const int per = nent <= 32 ? 4 : (nent <= 64 ? 2 : 1);
            for (uint32_t sg = 0; sg < nseg; sg += per) {
                float cdv[4];
                uint32_t slv[4];
                int eidx[4];
#pragma unroll
                for (int u = 0; u < 4; u++) {
                    const uint32_t unit = sg + (per == 4 ? u : per == 2 ? (u >> 1) : 0);
                    const int e = (per == 4 ? 0 : per == 2 ? (u & 1) * 32 : u * 32) + lane;
                    eidx[u] = e;
                    cdv[u] = INFINITY;
                    slv[u] = INVALID_SLOT;
                    if (unit < nseg && e < nent) {
                        const size_t base = ((size_t) first + (size_t) unit * (2 * TC_M)) * kc;
                        slv[u] = pslot[base + e];
                        cdv[u] = pdist[base + e];
                    }
                }
#pragma unroll
                for (int u = 0; u < 4; u++) {
                    const bool ok = slv[u] != INVALID_SLOT;
                    if (ok && eidx[u] % kc == kc - 1) own_min = fminf(own_min, cdv[u]);      // last entry of a full list
                    const unsigned m = __ballot_sync(FULL, ok);
                    if (m) { n_in += __popc(m); cand.offer(cdv[u], slv[u], ok, lane, 32 * KRC); }
                }
            }
        }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) own_min = fminf(own_min, __shfl_xor_sync(FULL, own_min, o));
    const float *qv = Q + (size_t) q * dim;
    const CertQ cq = cert_query(qv, dim, lane);
    // register r, lane e holds candidate 32 r + e (ascending key).  Rows outside the 32 KRC: in a partial list -> key >= the
    // last candidate's key; in none -> key >= min(R, own_min).  With no full list, no bound and no overflow the candidates
    // are everything.
    const float keylast = __shfl_sync(FULL, cand.d[KRC - 1], 31);
    const float g_rest = fminf(n_in > 32 * KRC ? keylast : INFINITY, fminf(R, own_min));
    const bool complete = !bounded && own_min == INFINITY && n_in <= 32 * KRC;
    WarpTopK<1, int64_t> top;
    const bool ok = cert_rerank<P, METRIC, KRC>(qv, arena, dim, cand, g_rest, complete, stats, cq, k, lane,
                                              [&](uint32_t ts) { return tc_row[ts]; }, [&](uint32_t ts) { return ids[tc_src[ts]]; }, top,
                                              counters ? counters + 1 : nullptr);
    if (!ok) {
        if (lane == 0) {
            const unsigned long long f = atomicAdd(counters, 1ull);
            fb_list[f] = (uint32_t) q;
            fb_tau[f] = top.td;           // an upper bound of the final k-th exact value (+inf: fewer than k so far)
        }
        return;                           // the exact kernel writes this query's result
    }
    if (lane < k) {
        const bool got = top.key[0] != KeyMax<int64_t>::v;
        out_dist[(size_t) q * k + lane] = got ? top.d[0] : INFINITY;
        out_ids[(size_t) q * k + lane] = got ? top.key[0] : -1;
    }
}

// ---- two-phase scans: the bound phase 1 leaves behind ----------------------------------------------------------
// After the scan of every query's NEAREST list (phase 1) its partial lists -- one per (segment, column half, replica) of
// that list -- each hold at most kc keys.  The k-th smallest key K over their union is an upper bound of the query's final
// k-th key (the union holds k rows at or below it), far tighter than the kc-th key of any single partial list; its
// relaxation (cert_bound.cuh) goes into the shared bound gthr[q] that phase 2 filters against.  Warp per query.
template <int METRIC>
__global__ void __launch_bounds__(128) ivf_tc_bound_kernel(
    const float *__restrict__ pdist, const uint32_t *__restrict__ pslot, const float *__restrict__ Q,
    const uint32_t *__restrict__ probe, const uint32_t *__restrict__ pairpos, const uint32_t *__restrict__ item_off,
    const uint32_t *__restrict__ list_len, const uint32_t *__restrict__ cnt, uint32_t rep_max, int nq, int nprobe, int nlists,
    uint32_t segb, int dim, int kc, int k, const float *__restrict__ stats, float *__restrict__ gthr)
{
    const int lane = threadIdx.x & 31;
    const int q = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (q >= nq) return;
    const size_t p = (size_t) q * nprobe;                    // rank 0: virtual list = the list itself
    const uint32_t l = probe[p];
    if (l >= (uint32_t) nlists) return;
    const uint32_t len = list_len[l];
    if (!len) return;
    const uint32_t rep = ivf_rep(cnt[l], rep_max), pos = pairpos[p] * rep, nseg = ivf_nseg(len, segb);
    const size_t first = (size_t) (item_off[l] + (pos / TC_M) * nseg) * (2 * TC_M) + (pos % TC_M) * 2;
    const int nent = (int) rep * 2 * kc;
    WarpTopK<1, uint32_t> best;
    best.init();
    uint32_t n_in = 0;
    // the nseg * nent entries as one flat range, four loads per lane in flight (the walk is a latency chain otherwise)
    const uint32_t total = nseg * (uint32_t) nent;
    for (uint32_t e0 = 0; e0 < total; e0 += 128) {
        uint32_t sl[4];
        float kd[4];
#pragma unroll
        for (int u = 0; u < 4; u++) {
            const uint32_t e = e0 + u * 32 + lane;
            sl[u] = INVALID_SLOT;
            kd[u] = INFINITY;
            if (e < total) {
                const uint32_t sg = e / (uint32_t) nent, w = e - sg * (uint32_t) nent;
                const size_t at = (first + (size_t) sg * (2 * TC_M)) * kc + w;
                sl[u] = pslot[at];
                kd[u] = pdist[at];
            }
        }
#pragma unroll
        for (int u = 0; u < 4; u++) {
            const bool ok = sl[u] != INVALID_SLOT;
            const unsigned m = __ballot_sync(FULL, ok);
            if (m) { n_in += __popc(m); best.offer(kd[u], sl[u], ok, lane, k); }
        }
    }
    if (n_in < (uint32_t) k) return;                          // fewer than k rows met: no bound
    float K = best.td;                                       // the k-th smallest key (its index bits masked off)
    // an upper bound of the value the packed key stood for (tc_knn_kernel publishes the same way)
    K = K >= 0.0f ? __uint_as_float(__float_as_uint(K) | TC_IDX_MASK) : K;
    const CertQ cq = cert_query(Q + (size_t) q * dim, dim, lane);
    if (lane == 0 && K < 1.0e38f) {
        const float R = cert_relax<METRIC>(K, stats, cq, dim);
        if (R < gthr[q]) gthr[q] = R;                         // (nothing else writes the bounds between the two scans)
    }
}

// ---- lists: the queries the 32-candidate certificate rejected ---------------------------------------------
// One CTA (32 warps) per such query.  tau0 = the k-th exact value the finish kernel reached, an upper bound of the
// final one.  Per (probed list, segment) unit:
//   - if a full partial list of the unit ends with a key m that cannot be certified against tau0
//     (cert_lower_bound(m) <= tau0), that list may have dropped a row of the answer: every row of the unit (at most
//     16 tiles = 4096 rows) is evaluated in the reference's arithmetic;
//   - otherwise the unit's entries are evaluated: whatever it dropped is certified out (full lists by their own
//     kc-th key, the others by R, for which cert_lower_bound(R) > tau0 holds by construction of the relaxation).
// Level 3, if cert_lower_bound(R) > tau fails after all (a zero query under cosine, fewer than k rows ...): every
// row of the probed lists -- ivfCollectCandidates without the cut-off.  Rows are reached through the tensor
// layout's row map (list l = tensor rows [ltile8[l] * 32, + len)).
constexpr int FB_THREADS = 1024, FB_WARPS = FB_THREADS / 32;

template <class P, int METRIC>
__global__ void __launch_bounds__(FB_THREADS) ivf_exact_fallback_kernel(
    const uint32_t *__restrict__ fb_list, const float *__restrict__ fb_tau, unsigned long long *__restrict__ counters,
    const float *__restrict__ pdist, const uint32_t *__restrict__ pslot, const uint32_t *__restrict__ probe,
    const uint32_t *__restrict__ pairpos, const uint32_t *__restrict__ item_off, const uint32_t *__restrict__ list_len,
    const uint32_t *__restrict__ ltile8, const float *__restrict__ gthr, const uint32_t *__restrict__ cnt, uint32_t rep_max,
    uint32_t split, uint32_t segb, int kc, const uint32_t *__restrict__ tc_src, const uint32_t *__restrict__ tc_row,
    const float *__restrict__ arena, const int64_t *__restrict__ ids, const float *__restrict__ Q,
    const float *__restrict__ stats, const float4 *__restrict__ vecs, const uint32_t *__restrict__ list_blk, int dimp,
    int nprobe, int nlists, int dim, int k, float *__restrict__ out_dist, int64_t *__restrict__ out_ids, float *__restrict__ dbg)
{
    extern __shared__ float fb_smem[];                   // [dim] query, then FB_THREADS (dist, id) pairs
    float *qs = fb_smem;
    float *md = fb_smem + ((dim + 3) & ~3);
    int64_t *mi = reinterpret_cast<int64_t *>(md + FB_THREADS);
    constexpr int MAXU = 64;
    __shared__ unsigned s_rows, s_nu;
    __shared__ float s_tau, s_res_d[32];
    __shared__ int64_t s_res_i[32];
    __shared__ uint32_t s_first[128], s_nseg[128], s_nent[128], s_l[128], s_len[128], s_ubase[129];     // nprobe <= 128
    __shared__ uint32_t s_slot0[MAXU], s_nrow[MAXU];      // units to rescan: first IL32 slot, rows
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const unsigned nfb = (unsigned) counters[0];
    for (unsigned f = blockIdx.x; f < nfb; f += gridDim.x) {
        const uint32_t q = fb_list[f];
        const float tau0 = fb_tau[f];
        __syncthreads();
        if (threadIdx.x == 0) { s_rows = 0; s_nu = 0; }
        for (int j = threadIdx.x; j < dim; j += blockDim.x) qs[j] = Q[(size_t) q * dim + j];
        __syncthreads();
        const CertQ cq = cert_query(qs, dim, lane);
        const float R = gthr[q];
        WarpTopK<1, int64_t> top;
        // merge of the 32 per-warp lists by warp 0: the k best of the CTA stay in s_res_*, their k-th value is returned
        auto merge_lists = [&]() -> float {
            md[w * 32 + lane] = top.d[0];
            mi[w * 32 + lane] = top.key[0];
            __syncthreads();
            if (w == 0) {
                WarpTopK<1, int64_t> fin;
                fin.init();
                for (int ww = 0; ww < FB_WARPS; ww++) {
                    const float d = md[ww * 32 + lane];
                    const int64_t id = mi[ww * 32 + lane];
                    fin.offer(d, id, id != KeyMax<int64_t>::v, lane, k);
                }
                s_res_d[lane] = fin.d[0];
                s_res_i[lane] = fin.key[0];
                if (lane == 0) s_tau = fin.td;
            }
            __syncthreads();
            return s_tau;
        };
        auto write_result = [&]() {
            if (w == 0 && lane < k) {
                const bool got = s_res_i[lane] != KeyMax<int64_t>::v;
                out_dist[(size_t) q * k + lane] = got ? s_res_d[lane] : INFINITY;
                out_ids[(size_t) q * k + lane] = got ? s_res_i[lane] : -1;
            }
        };
        top.init();
        // the (probed list, segment) units of the query, dealt round-robin to the warps
        if (threadIdx.x < 128) {
            uint32_t first = 0, nseg = 0, nent = 0, l = 0, len = 0;
            if ((int) threadIdx.x < nprobe) {
                const size_t p = (size_t) q * nprobe + threadIdx.x;
                l = probe[p];
                if (l < (uint32_t) nlists) {
                    len = list_len[l];
                    if (len) {
                        const uint32_t v = ivf_vlist(l, (int64_t) p, (uint32_t) nprobe, split);
                        const uint32_t rep = ivf_rep(cnt[v], rep_max), pos = pairpos[p] * rep;
                        nseg = ivf_nseg(len, segb);
                        first = (item_off[v] + (pos / TC_M) * nseg) * (2 * TC_M) + (pos % TC_M) * 2;
                        nent = rep * 2 * kc;
                    }
                }
            }
            s_first[threadIdx.x] = first; s_nseg[threadIdx.x] = nseg; s_nent[threadIdx.x] = nent; s_l[threadIdx.x] = l; s_len[threadIdx.x] = len;
        }
        __syncthreads();
        if (threadIdx.x == 0) {
            uint32_t acc = 0;
            for (int r = 0; r < nprobe; r++) { s_ubase[r] = acc; acc += s_nseg[r]; }
            s_ubase[nprobe] = acc;
        }
        __syncthreads();
        const uint32_t nunits = s_ubase[nprobe];
        int r = 0;
        for (uint32_t un = w; un < nunits; un += FB_WARPS) {
            while (s_ubase[r + 1] <= un) r++;
            {
                const uint32_t sg = un - s_ubase[r], first = s_first[r], l = s_l[r], len = s_len[r];
                const int nent = (int) s_nent[r];
                const size_t base = ((size_t) first + (size_t) sg * (2 * TC_M)) * kc;
                // smallest kc-th key of a full partial list of this unit
                float m = INFINITY;
                for (int i = lane; i < nent; i += 32)
                    if (i % kc == kc - 1 && pslot[base + i] != INVALID_SLOT) m = fminf(m, pdist[base + i]);
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) m = fminf(m, __shfl_xor_sync(FULL, m, o));
                const bool rescan = m < INFINITY && !(cert_lower_bound<METRIC>(m, stats, cq, dim) > tau0);
                bool queued = false;
                if (rescan) {
                    // all 8 warps rescan the unit together afterwards; its rows are IL32 slots list_blk[l] * 32 + ...
                    unsigned u = 0;
                    if (lane == 0) u = atomicAdd(&s_nu, 1u);
                    u = __shfl_sync(FULL, u, 0);
                    if (u < (unsigned) MAXU) {
                        if (lane == 0) {
                            s_slot0[u] = list_blk[l] * 32 + sg * segb * 32;
                            s_nrow[u] = min(segb * 32, len - sg * segb * 32);
                        }
                        queued = true;
                    }
                }
                if (rescan && !queued) {                  // (more than MAXU units: this warp does it alone)
                    const uint32_t slot0 = list_blk[l] * 32 + sg * segb * 32, nrow = min(segb * 32, len - sg * segb * 32);
                    if (lane == 0) atomicAdd(&s_rows, nrow);
                    for (uint32_t j0 = 0; j0 < nrow; j0 += 32) {
                        const uint32_t j = j0 + lane;
                        const bool valid = j < nrow;
                        float ed = INFINITY;
                        int64_t id = -1;
                        if (valid) { ed = cert_exact_il32<P>(qs, vecs, slot0 + j, dim, dimp); id = ids[slot0 + j]; }
                        top.offer(ed, id, valid, lane, k);
                    }
                } else if (rescan) {
                    // queued
                } else {
                    for (int i0 = 0; i0 < nent; i0 += 32) {
                        const int i = i0 + lane;
                        const uint32_t ts = i < nent ? pslot[base + i] : INVALID_SLOT;
                        // an entry whose own key is certified out against tau0 cannot be in the answer
                        const bool ok = ts != INVALID_SLOT && !(cert_lower_bound<METRIC>(pdist[base + i], stats, cq, dim) > tau0);
                        if (!__any_sync(FULL, ok)) continue;
                        float ed = INFINITY;
                        int64_t id = -1;
                        if (ok) { ed = cert_exact<P>(qs, arena + (size_t) tc_row[ts] * dim, dim); id = ids[tc_src[ts]]; }
                        top.offer(ed, id, ok, lane, k);
                    }
                }
            }
        }
        __syncthreads();
        {
            const unsigned nu = min(s_nu, (unsigned) MAXU);
            for (unsigned u = 0; u < nu; u++) {
                const uint32_t slot0 = s_slot0[u], nrow = s_nrow[u];
                if (threadIdx.x == 0) s_rows += nrow;
                for (uint32_t j0 = w * 32; j0 < nrow; j0 += FB_THREADS) {
                    const uint32_t j = j0 + lane;
                    const bool valid = j < nrow;
                    float ed = INFINITY;
                    int64_t id = -1;
                    if (valid) { ed = cert_exact_il32<P>(qs, vecs, slot0 + j, dim, dimp); id = ids[slot0 + j]; }
                    top.offer(ed, id, valid, lane, k);
                }
            }
        }
        const float tau = merge_lists();
        const bool certified = (tau <= tau0 || tau0 == INFINITY) && cert_lower_bound<METRIC>(R, stats, cq, dim) > tau;
        if (threadIdx.x == 0) {
            atomicAdd(counters + 5, (unsigned long long) s_rows);
            if (!certified) {
                const unsigned long long n3 = atomicAdd(counters + 4, 1ull);
                if (dbg && n3 < 16) {
                    float *rr = dbg + n3 * 8;
                    rr[0] = (float) q; rr[1] = R; rr[2] = tau; rr[3] = cert_lower_bound<METRIC>(R, stats, cq, dim); rr[4] = (float) s_rows;
                    rr[5] = cq.eq; rr[6] = cq.qn; rr[7] = tau0;
                }
            }
        }
        if (certified) {
            write_result();
            continue;
        }
        // ---- level 3: everything
        top.init();
        for (int r = 0; r < nprobe; r++) {
            const uint32_t l = probe[(size_t) q * nprobe + r];
            if (l >= (uint32_t) nlists) continue;
            const uint32_t len = list_len[l], base = ltile8[l] * 32;
            for (uint32_t j0 = w * 32; j0 < len; j0 += FB_THREADS) {
                const uint32_t j = j0 + lane;
                const bool valid = j < len;
                float ed = INFINITY;
                int64_t id = -1;
                if (valid) {
                    ed = cert_exact<P>(qs, arena + (size_t) tc_row[base + j] * dim, dim);
                    id = ids[tc_src[base + j]];
                }
                top.offer(ed, id, valid, lane, k);
            }
        }
        merge_lists();
        write_result();
    }
}

}  // namespace ndb
