"""CPU: the rounding-error certificates of the SURVEY 8f kernels, restated in numpy and attacked with random and adversarial
inputs.  Python floats are IEEE doubles rounded to nearest without contraction and np.float32 arithmetic rounds to float
after every operation, so these are the device's operations (csrc/cluster_kmeans.cu: ckm_pick_cert_kernel;
csrc/pq.cu: pq_adc_kernel).  The properties:

  1. cluster_kmeans' D^2-weighted draw: whenever the certificate on a parallel prefix sum names a row, the reference's
     sequential walk (ml_kmeans.c:78-101) stops at that row -- for any association of the parallel sum.
  2. PQ scan: whenever (float)(d(1-e)) == (float)(d(1+e)) for the table sum, the float equals the one the reference's single
     chain over all dimensions yields (ml_product_quantization.c:1063-1098).
  3. PQ float screen: t32 (1 - delta) never exceeds the reference's double sum, so a screened-out row is really farther
     than the bound."""
import numpy as np

U = 2.0 ** -53
RAND_MAX = 2147483647


def sequential_pick(w, sel, draw):
    s = 0.0
    for i in range(len(w)):
        if not sel[i]:
            s += float(w[i])
    r = (float(draw) / float(RAND_MAX)) * s
    for i in range(len(w)):
        if sel[i]:
            continue
        r -= float(w[i])
        if r <= 0:
            return i
    for i in range(len(w)):
        if not sel[i]:
            return i
    return -1


def parallel_prefix(masked, block):
    """inclusive prefix sums with another association: per-block running sums plus a carry of pairwise block totals"""
    n = len(masked)
    out = np.empty(n)
    carry = 0.0
    for b in range(0, n, block):
        seg = masked[b:b + block]
        out[b:b + block] = carry + np.cumsum(seg)
        carry = carry + float(np.sum(seg))                 # numpy's pairwise sum: not the sequential order
    return out


def certified_pick(w, sel, draw, block):
    n = len(w)
    masked = np.where(sel, 0.0, w)
    P = parallel_prefix(masked, block)
    T = float(P[-1])
    eps = 16.0 * (n + 16) * U * T
    if not eps > 0.0:
        return None
    r0 = (float(draw) / float(RAND_MAX)) * T
    before = np.concatenate([[0.0], P[:-1]])
    ok = (~sel) & (r0 - before >= eps) & (P - r0 >= eps)
    idx = np.nonzero(ok)[0]
    assert len(idx) <= 1, "the certificate must name at most one row"
    return int(idx[0]) if len(idx) else None


def test_certified_weighted_draw_is_the_sequential_walk():
    rng = np.random.default_rng(20261017)
    certified = total = 0
    for trial in range(400):
        n = int(rng.integers(2, 3000))
        kind = trial % 5
        if kind == 0:
            w = rng.random(n) ** 8 * 10.0 ** rng.integers(-20, 20)               # wide dynamic range
        elif kind == 1:
            w = np.full(n, float(rng.random()) + 0.1)                            # all equal: crossings land ON prefix values
        elif kind == 2:
            w = rng.random(n) * (rng.random(n) < 0.1)                            # mostly zeros (duplicates of chosen seeds)
        elif kind == 3:
            w = np.float32(rng.standard_normal(n) ** 2).astype(np.float64)       # float-valued weights
        else:
            w = 10.0 ** rng.uniform(-300, 300, n)                                # near the ends of the double range
        sel = rng.random(n) < 0.05
        sel[rng.integers(0, n)] = True
        if sel.all():
            sel[0] = False
        for draw in (0, 1, RAND_MAX, RAND_MAX - 1, int(rng.integers(0, RAND_MAX)), int(rng.integers(0, RAND_MAX))):
            want = sequential_pick(w, sel, draw)
            for block in (1, 7, 256):
                with np.errstate(over="ignore", invalid="ignore"):
                    got = certified_pick(w, sel, draw, block)
                total += 1
                if got is not None:
                    certified += 1
                    assert got == want, (trial, kind, n, draw, block)
    assert certified > 0.5 * total                        # (draws of 0 / RAND_MAX and the degenerate kinds go to the literal walk)


def _pq_case(rng, m, dsub, ksub, scale):
    q = (rng.standard_normal(m * dsub) * scale).astype(np.float32)
    cb = (rng.standard_normal((m, ksub, dsub)) * scale).astype(np.float32)
    return q, cb


def _chain(q, cb, code):
    """pq_asymmetric_distance: one double chain over all dimensions"""
    m, ksub, dsub = cb.shape
    t = 0.0
    for s in range(m):
        for d in range(dsub):
            diff = float(q[s * dsub + d]) - float(cb[s, code[s], d])
            t += diff * diff
    return t


def _table(q, cb):
    m, ksub, dsub = cb.shape
    T = np.zeros((m, ksub))
    for s in range(m):
        for c in range(ksub):
            t = 0.0
            for d in range(dsub):
                diff = float(q[s * dsub + d]) - float(cb[s, c, d])
                t += diff * diff
            T[s, c] = t
    return T


def test_pq_table_sum_certificate_and_float_screen():
    rng = np.random.default_rng(7)
    checked = certified = 0
    for trial in range(60):
        m, dsub, ksub = int(rng.choice([1, 2, 4, 8, 16])), int(rng.choice([1, 2, 3, 8])), int(rng.choice([2, 16, 64]))
        scale = float(10.0 ** rng.integers(-6, 7))
        q, cb = _pq_case(rng, m, dsub, ksub, scale)
        if trial % 6 == 0:                                 # small integers: every operation exact, distances ON float boundaries
            q = np.round(q / scale * 4).astype(np.float32)
            cb = np.round(cb / scale * 4).astype(np.float32)
        T = _table(q, cb)
        T32 = T.astype(np.float32)
        dim = m * dsub
        eps = (dim + m + 8) * 4.0 * 1.1102230246251565e-16
        one_minus_delta = np.float32(1.0) - np.float32(m + 8) * np.float32(2.0 ** -23)
        for _ in range(200):
            code = rng.integers(0, ksub, m)
            ref_total = _chain(q, cb, code)
            ref_f = np.float32(np.sqrt(ref_total))
            total = 0.0
            t32 = np.float32(0.0)
            for s in range(m):
                total += float(T[s, code[s]])
                t32 = np.float32(t32 + T32[s, code[s]])
            d = float(np.sqrt(total))
            lo, hi = np.float32(d * (1.0 - eps)), np.float32(d * (1.0 + eps))
            checked += dim >= 4
            if lo == hi:
                certified += dim >= 4
                assert lo == ref_f, (trial, m, dsub, ksub, code)
            # the screen's lower bound (float subnormals excluded, as in the kernel: t32 < 1e-30 always passes)
            if t32 >= np.float32(1e-30) and np.isfinite(t32):
                assert float(np.float32(t32 * one_minus_delta)) <= ref_total * (1.0 + 1e-15), (trial, m, t32, ref_total)
    # (one-dimensional vectors are the exception: |q - c| of two floats is often a 25-bit number, i.e. exactly halfway between
    # two floats -- those go down the reference's chain, as they must)
    assert certified > 0.99 * checked


def test_knn_double_order_certificate():
    """knn_classify / knn_regress (csrc/knn.cu: knn_neighbours): candidates are the kq smallest by (float distance, row); after
    re-sorting them by (double distance, row), the first k are the reference's neighbours whenever the float distance of the last
    candidate is strictly above that of the k-th -- rounding to float is monotonic."""
    rng = np.random.default_rng(11)
    certified = total = 0
    for trial in range(300):
        n, k = int(rng.integers(20, 2000)), int(rng.integers(1, 12))
        d = np.abs(rng.standard_normal(n)) + 1.0
        if trial % 3 == 0:                                   # clumps of doubles that share one float
            base = np.float32(d[rng.integers(0, n)])
            idx = rng.integers(0, n, int(rng.integers(2, 30)))
            d[idx] = float(base) * (1.0 + rng.uniform(-2e-8, 2e-8, len(idx)))
        f = d.astype(np.float32)
        kq = min(n, k + 8)
        cand = np.lexsort((np.arange(n), f))[:kq]
        order = cand[np.lexsort((cand, d[cand]))]
        want = np.lexsort((np.arange(n), d))[:k]
        total += 1
        if kq >= n or f[cand[-1]] > np.float32(d[order[k - 1]]):
            certified += 1
            assert np.array_equal(order[:k], want), trial
    assert certified > 0.8 * total
