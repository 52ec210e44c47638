"""CPU: the rounding-error bounds of the certified tensor-core selection (neurondb_b200/csrc/cert_bound.cuh), restated in
numpy float32 and attacked with random rows and queries.  For every (query, row) the key the tensor path would hold --
bf16-rounded inputs, fp32 accumulation (emulated both rounding to nearest and truncating, the worst a tensor core does),
the 12-bit packing at both ends of its range -- must satisfy

    cert_lower_bound(key) <= reference distance (ivfComputeDistance's f32 loop, from the oracle) <= cert_upper_bound(key)

and cert_relax(key) must be a key whose lower bound lies strictly above the upper bound of `key`.  These are the
inequalities the certified finish relies on; a violation here would be a wrong answer on the GPU."""
import numpy as np
import pytest

import oracle_lib as O

F = np.float32
IDX_MASK = np.uint32((1 << 11) - 1)


def bf16(a):
    """round-to-nearest-even to bfloat16, back as float32"""
    u = np.ascontiguousarray(a, np.float32).view(np.uint32).astype(np.uint64)
    r = ((u + 0x7FFF + ((u >> 16) & 1)) >> 16) << 16
    return r.astype(np.uint32).view(np.float32)


def acc32(terms, truncate):
    """fp32 accumulation of exact products in the given order; truncate = round toward zero after every addition"""
    s = F(0.0)
    for t in terms:
        e = float(s) + float(t)
        r = F(e)
        if truncate and abs(float(r)) > abs(e):
            r = np.nextafter(r, F(0.0))
        s = r
    return s


def pack_ends(key):
    """tc_pack: the low 11 mantissa bits are replaced by an index -- both extremes"""
    u = np.array([key], np.float32).view(np.uint32)
    lo = (u & ~IDX_MASK).view(np.float32)[0]
    hi = (u | IDX_MASK).view(np.float32)[0]
    return (lo, hi) if key >= 0 else (hi, lo)


class CertQ:
    def __init__(self, q):
        r = bf16(q)
        q64, r64 = q.astype(np.float64), r.astype(np.float64)
        self.eq = F(np.sqrt(F(np.sum((q64 - r64) ** 2)))) * F(1.0002)
        self.qn = F(np.sqrt(F(np.sum(q64 ** 2)))) * F(1.0002)
        self.qnr = F(np.sqrt(F(np.sum(r64 ** 2)))) * F(1.0002)
        self.qn_lo = F(np.sqrt(F(np.sum(q64 ** 2)))) * F(0.9998)


def stats_of(X):
    R = bf16(X).astype(np.float64)
    X64 = X.astype(np.float64)
    e2, x2, r2 = ((X64 - R) ** 2).sum(1), (X64 ** 2).sum(1), (R ** 2).sum(1)
    ok = r2 > 0
    return np.array([F(e2.max()) * F(1.0001), F(x2.max()) * F(1.0001),
                     F((e2[ok] / r2[ok]).max() if ok.any() else 0.0) * F(1.0002),
                     F((x2[ok] / r2[ok]).max() if ok.any() else 0.0) * F(1.0002)], np.float32)


def slack(metric, st, c, dim):
    exm, xm = F(np.sqrt(st[0])) * F(1.0002), F(np.sqrt(st[1])) * F(1.0002)
    gam = F(dim + 8) * F(1.1920929e-7)
    xr = xm + exm
    if metric == 1:
        return exm + c.eq, gam * (xr + c.qnr) * (xr + c.qnr), gam
    if metric == 3:
        return exm * c.qnr + xm * c.eq + gam * xm * c.qn, gam * xr * c.qnr, gam
    rho, kap = F(np.sqrt(st[2])) * F(1.0002), F(np.sqrt(st[3])) * F(1.0002)
    return rho * (c.qnr + c.qn) + kap * c.eq, gam * c.qnr, gam


def lower_bound(metric, key, st, c, dim):
    E, D, gam = slack(metric, st, c, dim)
    kv = key - abs(key) * F(4.8828125e-4) - D
    if metric == 1:
        r = F(np.sqrt(max(kv, F(0.0)))) - E
        return r - abs(r) * gam
    if metric == 3:
        return kv - E
    if not c.qn_lo > 0:
        return F(-np.inf)
    u = kv - E
    return F(1.0) + (u / c.qn_lo if u < 0 else u / c.qn) - F(8.0) * gam


def upper_bound(metric, key, st, c, dim):
    E, D, gam = slack(metric, st, c, dim)
    kv = key + abs(key) * F(4.8828125e-4) + D
    if metric == 1:
        r = F(np.sqrt(max(kv, F(0.0)))) + E
        return r + abs(r) * gam
    if metric == 3:
        return kv + E
    if not c.qn_lo > 0:
        return F(np.inf)
    u = kv + E
    return F(1.0) + (u / c.qn_lo if u > 0 else u / c.qn) + F(8.0) * gam


def relax(metric, key, st, c, dim):
    E, D, gam = slack(metric, st, c, dim)
    ub = upper_bound(metric, key, st, c, dim)
    if metric == 1:
        t = (ub + abs(ub) * F(2.0) * gam) + E
        R = t * t + D
    elif metric == 3:
        R = ub + E + D
    else:
        if not c.qn_lo > 0:
            return F(np.inf)
        v = ub - F(1.0) + F(16.0) * gam
        u = v * c.qn if v > 0 else v * c.qn_lo
        R = u + E + D
    return R + abs(R) * F(9.765625e-4) + F(1e-30)


def keys_of(metric, x, q, truncate):
    """the candidate value the epilogue forms for one (query, row) pair (tc_knn.cu: candidates())"""
    xr, qr = bf16(x).astype(np.float64), bf16(q).astype(np.float64)
    dot = acc32(xr * qr, truncate)                                    # bf16 x bf16 products are exact
    xn = acc32(xr * xr, False)                                        # tc_row_norms_kernel: an fmaf chain
    qn = F(np.sum((qr * qr).astype(np.float32)))                      # a fixed tree in the kernel; any order is within the bound
    if metric == 1:
        return F(F(F(-2.0) * dot + xn) + qn)                          # fmaf(-2, dot, ||x~||^2) + ||q~||^2
    if metric == 3:
        return -dot
    rinv = F(1.0) / F(np.sqrt(xn)) if xn > 0 else F(0.0)
    return -dot * rinv


@pytest.mark.parametrize("metric", [1, 2, 3])
def test_bounds_enclose_the_reference_distance(metric):
    rng = np.random.default_rng(100 + metric)
    with np.errstate(over="ignore", invalid="ignore"):
        for trial in range(140):
            dim = int(rng.choice([1, 3, 8, 32, 96, 128, 200, 768]))
            scale = float(10.0 ** rng.integers(-3, 4))
            n = 24
            X = (rng.standard_normal((n, dim)) * scale).astype(np.float32)
            q = (rng.standard_normal(dim) * scale).astype(np.float32)
            if trial % 4 == 0:
                X[:6] = q + (rng.standard_normal((6, dim)) * scale * 1e-3).astype(np.float32)      # near the query: tiny distances
            if trial % 5 == 0:
                X[6:9] *= np.float32(1e-4)                                                         # short rows (cosine: small norms)
            if trial % 7 == 0:
                X[9] = 0.0
            st = stats_of(X)
            c = CertQ(q)
            ref = O.distance_pairs(np.repeat(q[None, :], n, 0), X, metric, O.ARITH_IVF_F32)        # vec1 = query, vec2 = row
            for i in range(n):
                if metric == 2 and not np.any(bf16(X[i])):
                    continue                                            # a zero row packs as an empty key in the kernel
                for truncate in (False, True):
                    key = keys_of(metric, X[i], q, truncate)
                    for pk in pack_ends(key):
                        lo, hi = lower_bound(metric, pk, st, c, dim), upper_bound(metric, pk, st, c, dim)
                        assert lo <= ref[i] <= hi, (trial, dim, scale, i, truncate, float(pk), float(lo), float(ref[i]), float(hi))
                        R = relax(metric, pk, st, c, dim)
                        if np.isfinite(R) and np.isfinite(hi):
                            assert lower_bound(metric, R, st, c, dim) > hi, (trial, dim, i, float(pk), float(R))


# ---- the certified selection built on those bounds (cert_common.cuh: ivf_coarse_cert_kernel + cert_rerank) ----------
def _certified_nearest(keys, exact, st, c, dim, np_, nparts, kc, cap):
    """The decision logic of the certified coarse stage, restated: per-range partial lists of the kc smallest keys, the
    `cap` smallest of their union as candidates, exact re-evaluation 16 at a time in key order, certificate after every
    chunk.  Returns the certified top-np_ (list of row indices) or None (the kernel would send the query to the exact path)."""
    L = len(keys)
    bounds = np.linspace(0, L, nparts + 1).astype(int)
    G = np.inf
    union = []
    for p in range(nparts):
        idx = np.arange(bounds[p], bounds[p + 1])
        order = idx[np.lexsort((idx, keys[idx]))][:kc]
        if len(order) == kc:
            G = min(G, float(keys[order[-1]]))                       # a full list: rows it may have dropped have keys >= its last one
        union.extend(order.tolist())
    union = np.array(union)
    union = union[keys[union] <= G]                                  # entries above G are outranked by rows nobody kept
    union = union[np.lexsort((union, keys[union]))]
    cand = union[:cap]
    g_rest = float(keys[cand[-1]]) if len(union) > cap else G
    complete = G == np.inf and len(union) <= cap
    top = []                                                         # (exact distance, row), ascending
    for c0 in range(0, len(cand), 16):
        for r in cand[c0:c0 + 16]:
            top.append((float(exact[r]), int(r)))
        top = sorted(top)[:np_]
        more = c0 + 16 < len(cand)
        g = min(float(keys[cand[c0 + 16]]), g_rest) if more else g_rest
        td = top[np_ - 1][0] if len(top) >= np_ else np.inf
        certified = (complete and not more) or (np.isfinite(g) and lower_bound(1, F(g), st, c, dim) > F(td))
        if certified:
            return [r for _, r in top]
        if not more:
            return None
    return None


def test_certified_nearest_centroids_are_the_exact_ones():
    """Whenever the restated certificate accepts, the answer is the reference's (ivfSelectClusters: the np nearest centroids
    by (f32 L2, index)) -- on clustered centroids, near-duplicates and exact ties -- and it accepts most of the time.
    (The `keys <= G` filter of the candidate set is load-bearing: without it this test finds a clump of 75 near-identical
    centroids where a dropped row outranks an accepted answer.)"""
    rng = np.random.default_rng(77)
    accepted = total = 0
    with np.errstate(over="ignore", invalid="ignore"):
        for trial in range(60):
            dim = int(rng.choice([8, 32, 128]))
            L = int(rng.choice([64, 300, 1024]))
            C = rng.standard_normal((L, dim)).astype(np.float32)
            if trial % 3 == 0:
                C[: L // 4] = C[0] + (rng.standard_normal((L // 4, dim)) * 1e-3).astype(np.float32)    # a tight clump
            if trial % 4 == 0:
                C[5] = C[3]                                                                               # exact duplicates: ties by index
            q = (C[rng.integers(0, L)] + rng.standard_normal(dim).astype(np.float32) * np.float32(0.3)).astype(np.float32)
            st, cq = stats_of(C), CertQ(q)
            exact = O.distance_pairs(np.repeat(q[None, :], L, 0), C, 1, O.ARITH_IVF_F32)
            keys = np.array([pack_ends(keys_of(1, C[i], q, bool(i & 1)))[i >> 1 & 1] for i in range(L)], np.float32)
            want = np.lexsort((np.arange(L), exact))
            for np_, nparts, kc, cap in ((1, 4, 7, 32), (8, 8, 14, 32), (16, 2, 16, 128), (32, 16, 16, 128)):
                got = _certified_nearest(keys, exact, st, cq, dim, np_, nparts, kc, cap)
                total += 1
                if got is not None:
                    accepted += 1
                    assert got == want[:np_].tolist(), (trial, dim, L, np_, nparts, kc, cap)
    assert accepted > 0.5 * total


# ---- the tensor IVF search as a whole: scan with the relaxed shared bound, then the certified finish --------------------
# A model of what tc_knn.cu's list mode leaves behind and of ivf_tc_finish_cert_kernel's decision (ivf_cert.cuh), per query:
#   * the probed rows are split into units (probed list x segment x column half); a unit keeps its kc smallest packed keys;
#   * a row enters a unit's list only if its key is below the unit's threshold = min(own kc-th key, next_up(shared bound as last
#     read -- possibly stale, i.e. larger));
#   * a unit whose list holds >= k entries publishes R = cert_relax(upper end of its k-th key) into the shared bound (min);
#   * finish: candidates = the 32 KRC smallest entries over all units; rows elsewhere have keys >= g_rest =
#     min(last candidate key if the entries overflow, R, smallest kc-th key of a full list); re-evaluate 16 at a time; accept when
#     lower_bound(g) > tau.
# Property: an accepted answer is the reference's top k by (ivfComputeDistance, id) over ALL probed rows.
BIG = F(1.7014118346046923e38)


def _model_scan(metric, keys_raw, units, kc, k, st, cq, dim, rng):
    shared, history = F(np.inf), [F(np.inf)]
    lists = [[] for _ in units]
    published = [F(np.inf)] * len(units)
    gcap = [F(np.inf)] * len(units)
    sched = [(u, c0) for u, rows in enumerate(units) for c0 in range(0, len(rows), 32)]
    order = rng.permutation(len(sched))
    order = sorted(order, key=lambda i: (sched[i][1] + rng.integers(0, 96), sched[i][0]))     # roughly in step, like concurrent CTAs
    for i in order:
        u, c0 = sched[i]
        seen = history[max(0, len(history) - 1 - int(rng.integers(0, 3)))]                   # a stale read is a larger bound
        gcap[u] = min(gcap[u], np.nextafter(seen, F(np.inf)))
        L = lists[u]
        changed = False
        for j, row in enumerate(units[u][c0:c0 + 32]):
            thr = min(L[-1][0] if len(L) == kc else F(np.inf), gcap[u])
            if keys_raw[row] < thr:
                u32 = np.array([min(keys_raw[row], BIG)], np.float32).view(np.uint32)
                packed = ((u32 & ~IDX_MASK) | np.uint32((c0 + j) & 0x7FF)).view(np.float32)[0]
                L.append((packed, int(row)))
                L.sort()
                del L[kc:]
                changed = True
        if changed and len(L) >= k and L[k - 1][0] < published[u]:
            published[u] = L[k - 1][0]
            pub = pack_ends(published[u])[1]                        # upper end of the value the packed key stands for
            if pub < BIG:
                shared = min(shared, relax(metric, pub, st, cq, dim))
                history.append(shared)
    return lists, shared


def _model_finish(metric, lists, R, exact, kc, k, cap, st, cq, dim):
    entries = sorted(e for L in lists for e in L)
    own_min = min([L[-1][0] for L in lists if len(L) == kc], default=F(np.inf))
    cand = entries[:cap]
    g_rest = min(cand[-1][0] if len(entries) > cap else F(np.inf), R, own_min)
    complete = not (R < F(1.0e38)) and own_min == np.inf and len(entries) <= cap
    top = []
    for c0 in range(0, len(cand), 16):
        top = sorted(top + [(float(exact[r]), r) for _, r in cand[c0:c0 + 16]])[:k]
        more = c0 + 16 < len(cand)
        g = min(cand[c0 + 16][0], g_rest) if more else g_rest
        tau = top[k - 1][0] if len(top) >= k else np.inf
        if (complete and not more) or (np.isfinite(g) and lower_bound(metric, F(g), st, cq, dim) > F(tau)):
            return [r for _, r in top]
        if not more:
            return None
    return None if cand else ([] if complete else None)


@pytest.mark.parametrize("metric", [1, 2, 3])
def test_model_of_the_tensor_ivf_search_accepts_only_exact_answers(metric):
    rng = np.random.default_rng(500 + metric)
    accepted = total = 0
    with np.errstate(over="ignore", invalid="ignore"):
        for trial in range(36):
            dim = int(rng.choice([16, 96, 128]))
            n = int(rng.choice([200, 900, 2500]))
            centre = rng.standard_normal(dim).astype(np.float32)
            X = (centre + rng.standard_normal((n, dim)).astype(np.float32) * np.float32(rng.choice([0.05, 0.3, 1.0]))).astype(np.float32)
            if trial % 3 == 0:
                X[: n // 5] = X[0] + (rng.standard_normal((n // 5, dim)) * 1e-3).astype(np.float32)      # near-duplicates: keys cannot order them
            if trial % 4 == 0:
                X[7] = X[3]                                                                                # an exact tie, decided by id
            q = (X[rng.integers(0, n)] + rng.standard_normal(dim).astype(np.float32) * np.float32(0.2)).astype(np.float32)
            st, cq = stats_of(X), CertQ(q)
            exact = O.distance_pairs(np.repeat(q[None, :], n, 0), X, metric, O.ARITH_IVF_F32)
            keys_raw = np.array([keys_of(metric, X[i], q, bool(i & 1)) for i in range(n)], np.float32)
            want = np.lexsort((np.arange(n), exact))
            # units of uneven size, as probed lists of uneven length cut into segments and column halves
            cuts = np.sort(rng.choice(np.arange(1, n), size=min(n - 1, int(rng.integers(3, 24))), replace=False))
            units = [u for u in np.split(rng.permutation(n), cuts) if len(u)]
            for k, kc, cap in ((1, 7, 32), (10, 16, 32), (24, 16, 64)):
                lists, R = _model_scan(metric, keys_raw, units, kc, min(k, kc), st, cq, dim, rng)
                got = _model_finish(metric, lists, R, exact, kc, k, cap, st, cq, dim)
                total += 1
                if got is not None:
                    accepted += 1
                    assert got == want[:k].tolist(), (trial, metric, dim, n, k, kc, cap)
    assert accepted > 0.3 * total
