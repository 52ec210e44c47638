"""CPU: the oracle's relation images follow the reference's page layout (ivf_am.c / hnsw_am.c)."""
import os
import struct

import numpy as np
import pytest

import oracle_lib as O
import workloads as W


def header(page):
    lsn, csum, flags, lower, upper, special, psv, prune = struct.unpack_from("<QHHHHHHI", page.tobytes(), 0)
    return dict(lower=lower, upper=upper, special=special, psv=psv)


def test_ivf_relation_layout():
    X = W.mixture(900, 128, 6, 3)
    C, _, _, _, _ = O.kmeans_train(X[:600], 6)
    assign = O.ivf_assign(X, C)
    nb, blocks = O.ivf_encode_relation(X, C, assign)
    assert nb > 2
    magic, version, nlists, nprobe, dim, cblk, inserted = struct.unpack_from("<IIiiiIq", blocks[0].tobytes(), 24)
    assert (magic, version, nlists, dim, cblk, inserted) == (0x49564646, 1, 6, 128, 1, 900)
    h = header(blocks[1])
    assert h["special"] == 8192 - 24 and h["psv"] == 8192 | 4           # PageInit(..., sizeof(IvfCentroidData))
    assert (h["lower"] - 24) // 4 == 6                                   # six centroid items
    # 128-d entries are 8 + 512 B: 15 per 8 KB list page
    counts = []
    for b in range(2, nb):
        hh = header(blocks[b])
        assert hh["special"] == 8192 - 8
        nxt, cnt = struct.unpack_from("<Ii", blocks[b].tobytes(), 8192 - 8)
        assert cnt == (hh["lower"] - 24) // 4 <= 15
        counts.append(cnt)
    assert sum(counts) == 900 and max(counts) == 15


def test_ivf_single_centroid_page_overflows_like_the_reference():
    """dim 128: 536-byte centroid items, ~15 per page; lists=64 cannot be written."""
    X = W.gaussian(200, 128, 1)
    C = X[:64].copy()
    rc, _ = O.ivf_encode_relation(X, C, np.zeros(200, np.int32), multi_page_centroids=False)
    assert rc == -1
    rc, _ = O.ivf_encode_relation(X, C[:15], np.zeros(200, np.int32), multi_page_centroids=False)
    assert rc > 0


def test_hnsw_relation_layout():
    X = W.gaussian(300, 24, 2)
    levels = O.hnsw_levels(300, seed=2)
    g = O.Hnsw(24, 8, 32, 32, capacity=300)
    g.build(X, levels, 1)
    nb, blocks = O.hnsw_encode_relation(g, X, efc=32, efs=32)
    assert nb == 301
    e = g.export()
    magic, version, ep, el, ml, m, efc, efs, pad, mlf, inserted = struct.unpack_from("<IIIiihhhhfq", blocks[0].tobytes(), 24)
    assert magic == 0x48534E57 and m == 8 and inserted == 300 and ep == e["entry_point"] + 1 and el == e["entry_level"]
    for i in (0, 17, 299):
        page = blocks[i + 1].tobytes()
        lp, = struct.unpack_from("<I", page, 24)
        off, ln = lp & 0x7fff, lp >> 17
        level, dim = struct.unpack_from("<ih", page, off + 8)
        assert level == levels[i] and dim == 24
        assert ln == (48 + 24 * 4 + (level + 1) * 16 * 4 + 7) // 8 * 8      # HnswNodeSizeWithM (hnsw_am.c:159-165)
        vec = np.frombuffer(page, np.float32, 24, off + 48)
        assert np.array_equal(vec, X[i])
        nb0 = np.frombuffer(page, np.uint32, 16, off + 48 + 96)
        want = np.where(e["nbr0"][i] == 0xFFFFFFFF, 0xFFFFFFFF, e["nbr0"][i] + 1)
        assert np.array_equal(nb0, want)


# ---- the page layouts themselves, against the reference's struct definitions ----------------------------
@pytest.mark.skipif(O.ref_leafs_lib() is None, reason="oracle/_ref not built (reference tree absent)")
def test_page_layout_equals_the_reference_structs():
    """Sizes and field offsets of IvfMetaPageData, IvfCentroidData, IvfListPageHeader, IvfListEntryData,
    HnswMetaPageData, HnswNodeData, HnswNodeSizeWithM and HnswGetNeighborsSafe, read back from the reference's
    own definitions (oracle/extract_ref_leafs.py), equal what the oracle's relation encoders write -- and the
    GPU loaders are tested against those relations (tests/test_gpu_pages.py)."""
    want, got = O.ref_page_layout(), O.page_layout()
    assert np.array_equal(got, want), (got, want)
    assert want[0] == 32 and want[9] == 24 and want[22] == 40 and want[34] == 48 and want[40] == 3248


def test_page_layout_golden():
    g = np.load(os.path.join(os.path.dirname(__file__), "golden", "index_leafs.npz"))
    assert np.array_equal(O.page_layout(), g["page_layout"])
