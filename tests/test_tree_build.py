"""Product tree build (libdpq: dpq_tree_from_edges host half, dpq_tree_build with the GPU edge
search) against files written by the UNMODIFIED reference (tests/golden, canonical stable-sort
build) and against the oracle.  Everything here is bit-exact."""
import numpy as np
import pytest

import datagen as dg
import deltapq_b200 as dpq
from oracle import pyoracle as po


def _check_against_reference_files(t, g):
    n = int(g["n"])
    assert np.array_equal(t["vec_id"], g["vec_id"])
    assert np.array_equal(t["payload"], g["payload"])                       # ..._compressed_codes_opt body
    qn = t["qnodes"].reshape(n + 1, 60)
    assert np.array_equal(qn[:n, 4:], g["qnode_tail"])                      # ..._TreeNodesDFS records
    assert np.array_equal(qn[:n, 0:4].copy().view(np.uint32).ravel(), g["vec_id"])
    tail = np.zeros(60, np.uint8)
    tail[16] = 1
    assert np.array_equal(qn[n], tail)
    # lossless: the stream decodes to the codes in DFS order
    codes, depth, parent = po.decode(t["payload"], n, 8)
    assert np.array_equal(codes, g["codes"][t["vec_id"]])
    assert np.array_equal(depth, t["depth"])
    assert np.array_equal(parent[1:].astype(np.uint32), t["parent_pos"][1:])


@pytest.mark.parametrize("fixture", ["golden4000", "golden1501"])
def test_host_layout_and_stream_match_reference_files(request, fixture):
    g = request.getfixturevalue(fixture)
    t = dpq.tree_from_edges(g["codes"], g["cw"], g["edges"], int(g["root"]))
    _check_against_reference_files(t, g)
    assert t["n_diffs"] == len(g["payload"]) - 8 - (3 * (int(g["n"]) - 1) + 1) // 2


def test_host_layout_m16_extension_matches_oracle(golden_m16):
    g = golden_m16
    edges, root = po.find_edges(g["codes"][:300], 256, 1, 1)
    t = dpq.tree_from_edges(g["codes"][:300], g["cw"], edges, root)
    lay = po.layout(g["codes"][:300], edges, root, po.centroid_tables(g["cw"]))
    for k in ("vec_id", "parent_pos", "child_num", "depth", "max_dist", "max_dist2p"):
        assert np.array_equal(t[k], lay[k]), k
    assert np.array_equal(t["payload"], po.stream(g["codes"][:300], lay))
    assert "qnodes" not in t


def test_host_layout_rejects_bad_edges(golden1501):
    g = golden1501
    bad = g["edges"].copy()
    bad[5, 1] = bad[6, 1]                       # duplicate child: not a tree
    with pytest.raises(dpq.DpqError):
        dpq.tree_from_edges(g["codes"], g["cw"], bad, int(g["root"]))
    with pytest.raises(dpq.DpqError):
        dpq.tree_from_edges(g["codes"], g["cw"], g["edges"], 10 ** 6)


def test_single_node_tree():
    cw = np.zeros((8, 256, 2), np.float32)
    t = dpq.tree_from_edges(np.arange(8, dtype=np.uint8)[None], cw, np.zeros((0, 2), np.uint32), 0)
    assert np.array_equal(t["payload"], np.arange(8, dtype=np.uint8))


# ------------------------------------------------------------------------------ GPU ----
@pytest.mark.gpu
@pytest.mark.parametrize("fixture", ["golden4000", "golden1501"])
def test_gpu_find_edges_matches_reference_edges_file(request, fixture):
    g = request.getfixturevalue(fixture)
    edges, root = dpq.find_edges(g["codes"], 256, 1, 1)
    assert root == int(g["root"])
    assert np.array_equal(edges, g["edges"])                                # ..._Approx_Edges body
    t = dpq.tree_build(g["codes"], g["cw"])
    _check_against_reference_files(t, g)


@pytest.mark.gpu
@pytest.mark.parametrize("method", [1, 2])
@pytest.mark.parametrize("shape", [(3000, 8, 256), (700, 16, 256), (500, 8, 16), (400, 12, 100), (1, 8, 256), (2, 8, 256)])
def test_gpu_find_edges_matches_oracle(shape, method):
    n, M, K = shape
    rng = np.random.default_rng(n * 31 + M)
    # few distinct values per subspace: many duplicates, long runs, deep stars
    codes = rng.integers(0, min(K, 6), size=(n, M)).astype(np.uint8)
    codes[rng.random(n) < 0.3] = codes[0]
    oe, oroot = po.find_edges(codes, K, 1, method)
    ge, groot = dpq.find_edges(codes, K, 1, method)
    assert groot == oroot
    assert np.array_equal(ge, oe)


@pytest.mark.gpu
def test_gpu_find_edges_sift_20k_and_diffs():
    base = dg.sift_like(20000, 128, seed=5)
    cw = dg.roundtrip_codebook(dg.kmeans_codebook(dg.sift_like(4000, 128, seed=6), 8, 256, iters=3))
    codes = dpq.encode(cw, base)
    assert np.array_equal(codes, po.encode(cw, base))
    oe, oroot = po.find_edges(codes, 256, 1, 1)
    t = dpq.tree_build(codes, cw)
    assert t["root_id"] == oroot and np.array_equal(t["edges"], oe)
    lay = po.layout(codes, oe, oroot, po.centroid_tables(cw))
    assert np.array_equal(t["payload"], po.stream(codes, lay))
    bm, nd = dpq.edge_diffs(codes, t["edges"])
    assert nd == t["n_diffs"]


@pytest.mark.gpu
@pytest.mark.parametrize("shape", [(5000, 8, 256, 6), (1200, 16, 256, 4), (900, 12, 100, 3), (3000, 8, 256, 200),
                                   (1, 8, 256, 2), (2, 8, 256, 2), (3, 16, 256, 2)])
def test_gpu_layout_and_stream_equal_host_layout(shape):
    """dpq_tree_build lays the tree out and writes the stream on the device (layout.cu);
    dpq_tree_from_edges is the sequential host walk of the same reference stage
    (DCAT.h:1334-1487, 1765-1842).  Every array must be identical."""
    n, M, K, distinct = shape
    rng = np.random.default_rng(n * 7 + M)
    codes = rng.integers(0, min(K, distinct), size=(n, M)).astype(np.uint8)
    codes[rng.random(n) < 0.2] = codes[0]
    cw = rng.normal(size=(M, K, 4)).astype(np.float32)
    g = dpq.tree_build(codes, cw)
    h = dpq.tree_from_edges(codes, cw, g["edges"], g["root_id"])
    for k in ("vec_id", "parent_pos", "child_num", "depth", "max_dist", "max_dist2p", "codes_by_pos", "payload"):
        assert np.array_equal(g[k], h[k]), k
    assert g["n_diffs"] == h["n_diffs"]
    if M == 8:
        assert np.array_equal(g["qnodes"], h["qnodes"])


@pytest.mark.gpu
def test_gpu_layout_sift_200k_equals_host_layout():
    base = dg.sift_like(200000, 128, seed=11)
    cw = dg.roundtrip_codebook(dg.kmeans_codebook(dg.sift_like(4000, 128, seed=6), 8, 256, iters=3))
    codes = dpq.encode(cw, base)
    g = dpq.tree_build(codes, cw)
    h = dpq.tree_from_edges(codes, cw, g["edges"], g["root_id"])
    for k in ("vec_id", "parent_pos", "child_num", "depth", "max_dist", "max_dist2p", "codes_by_pos", "payload"):
        assert np.array_equal(g[k], h[k]), k


def _planted_codes(n, M, K, rng, n_seeds, max_changes, min_changes=0):
    """Sparse codes (uniform over K^M) with planted near-duplicates: every code is a seed code with
    min_changes..max_changes random subspaces redrawn, so merges happen at many different diff
    levels (min_changes = 0 also plants exact duplicates)."""
    seeds = rng.integers(0, K, size=(n_seeds, M))
    codes = seeds[rng.integers(0, n_seeds, n)].copy()
    for i in range(n):
        ch = rng.integers(min_changes, max_changes + 1)
        idx = rng.choice(M, size=ch, replace=False)
        codes[i, idx] = rng.integers(0, K, size=ch)
    return codes.astype(np.uint8)


@pytest.mark.gpu
@pytest.mark.parametrize("min_changes", [0, 1])
@pytest.mark.parametrize("shape", [(2000, 12, 256, 300, 8, 1), (1500, 16, 256, 200, 10, 1), (1800, 13, 50, 250, 6, 2)])
def test_gpu_find_edges_futile_pass_prefilter_is_exact(shape, min_changes, monkeypatch):
    """M >= 12: passes whose dropped set contains no pair's changed-subspace mask are skipped
    (edges.cu futile_pass_prefilter).  Edges must equal the oracle's and the unfiltered run's."""
    n, M, K, n_seeds, max_changes, method = shape
    rng = np.random.default_rng(n + M)
    codes = _planted_codes(n, M, K, rng, n_seeds, max_changes, min_changes)
    monkeypatch.delenv("DPQ_NO_PREFILTER", raising=False)
    ge, groot = dpq.find_edges(codes, K, 1, method)
    oe, oroot = po.find_edges(codes, K, 1, method)
    assert groot == oroot and np.array_equal(ge, oe)
    monkeypatch.setenv("DPQ_NO_PREFILTER", "1")
    ue, uroot = dpq.find_edges(codes, K, 1, method)
    assert uroot == groot and np.array_equal(ue, ge)


@pytest.mark.gpu
def test_gpu_find_edges_prefilter_sift_m12_equals_unfiltered(monkeypatch):
    base = dg.sift_like(30000, 120, seed=15)
    cw = dg.roundtrip_codebook(dg.kmeans_codebook(dg.sift_like(4000, 120, seed=16), 12, 256, iters=3))
    codes = dpq.encode(cw, base)
    monkeypatch.delenv("DPQ_NO_PREFILTER", raising=False)
    ge, groot = dpq.find_edges(codes, 256, 1, 1)
    monkeypatch.setenv("DPQ_NO_PREFILTER", "1")
    ue, uroot = dpq.find_edges(codes, 256, 1, 1)
    assert uroot == groot and np.array_equal(ue, ge)
