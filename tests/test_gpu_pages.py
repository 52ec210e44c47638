"""GPU: relation images in the reference's page layout are decoded into the device structures."""
import numpy as np
import pytest

import workloads as W

pytestmark = pytest.mark.gpu
BITS = lambda a: np.ascontiguousarray(a, np.float32).view(np.uint32)


def tid(i):
    return ((i // 100 + 1) << 16) | (i % 100 + 1)          # (heap block, offset) packed


@pytest.mark.parametrize("dim,lists,metric", [(128, 12, 1), (20, 7, 2), (33, 5, 3)])
def test_ivf_load_relation_matches_inserted_index(ndb, orc, dim, lists, metric):
    n = 4000
    X = W.mixture(n, dim, lists, 100 + dim)
    Q = W.mixture(64, dim, lists, 101 + dim, centers_seed=100 + dim)
    tids = np.array([tid(i) for i in range(n)], np.int64)
    C, _, _, _, _ = orc.kmeans_train(X[: min(n, lists * 100)], lists)
    assign = orc.ivf_assign(X, C)
    nb, blocks = orc.ivf_encode_relation(X, C, assign, tids)
    assert nb > 0
    ix = ndb.IvfIndex(dim, lists, metric)
    ix.load_relation(blocks)
    assert len(ix) == n
    assert np.array_equal(BITS(ix.centroids()), BITS(C))
    off, rows = orc.lists_from_assignment(assign, lists)
    assert np.array_equal(ix.list_sizes(), np.diff(off))
    for mode in (ndb.IVF_FULL, ndb.IVF_LITERAL):
        d, i = ix.search(Q, 4, 10, mode)
        od, oi, _ = orc.ivf_search(X, C, off, rows, Q, 4, 10, strategy=metric, literal=mode == ndb.IVF_LITERAL, ids=tids)
        assert np.array_equal(i, oi) and np.array_equal(BITS(d), BITS(od))


def test_ivf_load_relation_skips_dead_items(ndb, orc):
    """LP_DEAD line pointers (what bulkdelete leaves) are skipped, as in ivf_am.c:1816."""
    n, dim, lists = 1500, 16, 4
    X = W.mixture(n, dim, lists, 7)
    C, _, _, _, _ = orc.kmeans_train(X[:400], lists)
    assign = orc.ivf_assign(X, C)
    nb, blocks = orc.ivf_encode_relation(X, C, assign)
    # kill the first three items of the first list page
    for off in (1, 2, 3):
        orc.page_mark_dead(blocks, 2, off)
    ix = ndb.IvfIndex(dim, lists)
    ix.load_relation(blocks)
    assert len(ix) == n - 3
    first_list = assign[0]                              # block 2 is the first page of row 0's list
    dead = np.flatnonzero(assign == first_list)[:3]
    keep = np.ones(n, bool)
    keep[dead] = False
    off, rows = orc.lists_from_assignment(assign[keep], lists)
    d, i = ix.search(X[:50], lists, 5)
    od, oi, _ = orc.ivf_search(X[keep], C, off, rows, X[:50], lists, 5, ids=np.flatnonzero(keep).astype(np.int64))
    assert np.array_equal(i, oi) and np.array_equal(BITS(d), BITS(od))


def test_ivf_load_relation_rejects_garbage(ndb):
    ix = ndb.IvfIndex(8, 4)
    with pytest.raises(ndb.NdbError):
        ix.load_relation(np.zeros((3, 8192), np.uint8))            # bad magic


def test_hnsw_load_relation_search_parity(ndb, orc):
    n, dim, m = 1200, 48, 8
    X = W.normalised(n, dim, 5)
    Q = W.normalised(80, dim, 6)
    levels = orc.hnsw_levels(n, seed=11)
    g = orc.Hnsw(dim, m, 32, 32, capacity=n)
    g.build(X, levels, 0)                                           # the literal reference build
    tids = np.array([tid(i) for i in range(n)], np.int64)
    nb, blocks = orc.hnsw_encode_relation(g, X, tids, efc=32, efs=32)
    assert nb == n + 1
    h = ndb.HnswIndex(dim, m, 32, 32)
    h.load_relation(blocks)
    assert len(h) == n
    for mode, smode in ((ndb.HNSW_LITERAL, 0), (ndb.HNSW_BESTFIRST, 1)):
        d, i = h.search(Q, 32, 10, 1, mode)
        od, on, _ = g.search(Q, 32, 10, 1, smode)
        want = np.where(on == 0xFFFFFFFF, -1, tids[np.minimum(on, n - 1)])
        assert np.array_equal(i, want) and np.array_equal(BITS(d), BITS(od))


def _set_line_pointer(blocks, block, offnum, lp_off, lp_len, flags=1):
    b = blocks.reshape(-1, 8192)
    lp = (lp_off & 0x7fff) | ((flags & 3) << 15) | ((lp_len & 0x7fff) << 17)
    b[block, 24 + 4 * (offnum - 1):24 + 4 * offnum] = np.frombuffer(np.uint32(lp).tobytes(), np.uint8)


def test_loaders_reject_line_pointers_that_leave_the_page(ndb, orc):
    """A torn or corrupt page: lp_off / lp_len are 15-bit fields and may point past the 8 KB block.  The loaders must
    refuse the relation, and a refused IVF relation must leave the handle empty and usable."""
    n, dim, lists = 1500, 16, 4
    X = W.mixture(n, dim, lists, 7)
    C, _, _, _, _ = orc.kmeans_train(X[:400], lists)
    assign = orc.ivf_assign(X, C)
    nb, blocks = orc.ivf_encode_relation(X, C, assign)
    for lp_off, lp_len in ((32000, 72), (8100, 500), (8, 72)):        # beyond the block; runs over pd_special; inside the header
        bad = blocks.copy()
        _set_line_pointer(bad, 2, 2, lp_off, lp_len)
        ix = ndb.IvfIndex(dim, lists)
        with pytest.raises(ndb.NdbError):
            ix.load_relation(bad)
        assert len(ix) == 0
        ix.ivfinsert(X[:100])                    # the failed load left no stale row ids behind
        assert len(ix) == 100 and np.array_equal(np.sort(ix.search(X[:5], lists, 1)[1][:, 0]), np.arange(5))
    # a second centroid item for the same list
    bad = blocks.copy().reshape(-1, 8192)
    hdr = bad[1, 24:28].view(np.uint32)[0] & 0x7fff      # first centroid item's offset
    second = bad[1, 28:32].view(np.uint32)[0] & 0x7fff
    bad[1, second:second + 4] = bad[1, hdr:hdr + 4]      # duplicate listId
    with pytest.raises(ndb.NdbError):
        ndb.IvfIndex(dim, lists).load_relation(bad.reshape(-1))
    # HNSW: node item pointing outside its page
    Xh = W.gaussian(300, 12, 5)
    levels = orc.hnsw_levels(300, seed=3)
    g = orc.Hnsw(12, 8, 32, 32, capacity=300)
    g.build(Xh, levels, 1)
    nbh, hb = orc.hnsw_encode_relation(g, Xh)
    badh = hb.copy()
    _set_line_pointer(badh, 5, 1, 8000, 3000)
    with pytest.raises(ndb.NdbError):
        ndb.HnswIndex(12, 8, 32, 32).load_relation(badh)
