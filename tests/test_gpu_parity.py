"""GPU: the CUDA path (through the C ABI, libdpq.so) against the oracle and the reference
fixtures.  Integer/byte results bit-exact; distances bit-exact where the reference's double
accumulation is exact, else within 1e-5 relative; ids modulo ties at 1e-5."""
import os

import numpy as np
import pytest

import datagen as dg
import deltapq_b200 as dpq
from helpers import assert_topk_equal, REL_TOL, gt_fixture, assert_gt_equals_reference_text
from oracle import pyoracle as po

pytestmark = pytest.mark.gpu


@pytest.fixture(params=["v2", "gen1"], autouse=True)
def engine(request, monkeypatch):
    """Every test runs on both scan engines: the second-generation fixed-record kernel
    (scan2.cu, default for M <= 8) and the first-generation op-program kernel (kernels.cu,
    DPQ_ENGINE=1; also what M = 16 uses)."""
    if request.param == "gen1":
        monkeypatch.setenv("DPQ_ENGINE", "1")
    else:
        monkeypatch.delenv("DPQ_ENGINE", raising=False)
    return request.param


def _open(g, **opts):
    ix = dpq.DeltaTreeIndex(g["payload"], int(g["n"]), int(g["M"]), int(g["K"]), pos2id=g["vec_id"])
    ix.set_codebook(g["cw"])
    for k, v in opts.items():
        ix.set_option(k, v)
    return ix


def test_adc_tables_bit_exact(golden4000):
    g = golden4000
    lut = dpq.adc_tables(g["cw"], g["queries"])
    for i in range(len(g["ref_lut"])):
        assert np.array_equal(lut[i], g["ref_lut"][i])          # vs the reference itself
    for i, q in enumerate(g["queries"]):
        assert np.array_equal(lut[i], po.lut(g["cw"], q))        # vs the oracle


@pytest.mark.parametrize("pack", [1, 2])
@pytest.mark.parametrize("fixture", ["golden4000", "golden1501"])
def test_search_matches_reference_fixture(request, fixture, pack):
    g = request.getfixturevalue(fixture)
    n, k = int(g["n"]), int(g["topk"])
    ix = _open(g, pack=pack)
    pos, ids, dist = ix.search(g["queries"], k)
    ref_pos = np.where(g["ref_pos"] == n, n - 1, g["ref_pos"])
    assert np.array_equal(dist, g["ref_dist"])                   # bit-exact distances
    assert_topk_equal(pos, dist, ref_pos, g["ref_dist"])
    assert np.array_equal(ids, g["vec_id"][pos])
    ix.close()


@pytest.mark.parametrize("pack,slices,warps", [(1, 1, 16), (1, 5, 3), (2, 3, 8), (2, 0, 16)])
def test_search_vs_oracle_all_nodes(golden4000, pack, slices, warps):
    """k = 64: deep into the candidate lists; every reported (pos, dist) is checked against
    the oracle's per-node distance, and the sets against the oracle's top-k."""
    g = golden4000
    n = int(g["n"])
    ix = _open(g, pack=pack, slices=slices, warps=warps)
    k = 64
    pos, _, dist = ix.search(g["queries"], k)
    for i, q in enumerate(g["queries"]):
        opos, odist, nd = po.scan(g["payload"], n, g["cw"], q, k, want_node_dist=True)
        assert np.array_equal(dist[i], odist)
        assert_topk_equal(pos[i], dist[i], opos, odist, node_dist=nd)
        assert len(set(pos[i])) == k
    ix.close()


def test_engine_selection(golden4000, golden_m16, engine):
    ix = _open(golden4000)
    assert ix.stat("engine") == (2 if engine == "v2" else 1)
    ix.close()


@pytest.mark.parametrize("opts", [dict(slices=1), dict(slices=7), dict(epoch=4, trigger=20), dict(epoch=1),
                                  dict(ramp=0), dict(ramp=0, epoch=256)])
def test_v2_options_and_overflow_path(golden4000, engine, opts):
    """Epoch / trigger / slice settings change how candidates are collected, never the result;
    ramp=0 floods the candidate buffers in the first epoch, which must divert the affected
    queries to the exact fallback instead of losing candidates."""
    if engine != "v2":
        pytest.skip("v2 options")
    g = golden4000
    n = int(g["n"])
    ix = _open(g, **opts)
    for k in (10, 100):
        pos, _, dist = ix.search(g["queries"], k)
        for i, q in enumerate(g["queries"]):
            opos, odist, nd = po.scan(g["payload"], n, g["cw"], q, k, want_node_dist=True)
            assert np.array_equal(dist[i], odist)
            assert_topk_equal(pos[i], dist[i], opos, odist, node_dist=nd)
    if opts.get("ramp") == 0:
        assert ix.stat("last_fallback") > 0
    ix.close()


@pytest.mark.parametrize("opts", [dict(coarse=1), dict(coarse=1, sample=4), dict(coarse=1, sample=1),
                                  dict(coarse=1, bcap8=32), dict(coarse=1, slices=3, warps8=24),
                                  dict(coarse=1, seed=1), dict(coarse=1, seed=1, presample=64, levels8=40),
                                  dict(coarse=1, seed=1, bcap8=32)])
def test_coarse_search_is_exact(golden4000, engine, opts):
    """Three-phase search (sample scan -> 8-bit coarse scan -> exact re-score, scan8.cu) gives
    the same answers as the oracle; a tiny candidate buffer (bcap8=32) overflows and must divert
    the query to the exact fallback instead of losing candidates."""
    if engine != "v2":
        pytest.skip("coarse search belongs to the v2 engine")
    g = golden4000
    n = int(g["n"])
    rng = np.random.default_rng(23)
    queries = np.clip(g["queries"][rng.integers(0, len(g["queries"]), 130)] + rng.integers(-12, 13, (130, 128)), 0, 255).astype(np.float32)
    ix = _open(g, **opts)
    for k in (1, 10, 40):
        pos, ids, dist = ix.search(queries, k)
        assert ix.stat("last_coarse") == 1
        for i in range(0, 130, 7):
            opos, odist, nd = po.scan(g["payload"], n, g["cw"], queries[i], k, want_node_dist=True)
            assert np.array_equal(dist[i], odist)
            assert_topk_equal(pos[i], dist[i], opos, odist, node_dist=nd)
            assert np.array_equal(ids[i], g["vec_id"][pos[i]])
    if opts.get("bcap8") == 32:
        assert ix.stat("last_fallback") > 0
    ix.close()


def test_coarse_search_duplicates_and_exact_matches(engine):
    """Queries that ARE database vectors (k-th distance 0 for duplicates) and heavy ties."""
    if engine != "v2":
        pytest.skip("coarse search belongs to the v2 engine")
    rng = np.random.default_rng(71)
    M, K, n = 8, 256, 6000
    cw = dg.roundtrip_codebook((rng.random((M, K, 4)) * 50).astype(np.float32))
    uniq = rng.integers(0, K, (300, M)).astype(np.uint8)
    codes = uniq[rng.integers(0, 300, n)]
    _, _, lay, payload = po.build_tree(codes, cw)
    ix = dpq.DeltaTreeIndex(payload, n, M, K)
    ix.set_codebook(cw)
    ix.set_option("coarse", 1)
    # queries = exact reconstructions of some codes: distance 0 to ~20 duplicates each
    queries = np.stack([np.concatenate([cw[m, uniq[i, m]] for m in range(M)]) for i in range(40)]).astype(np.float32)
    for k in (5, 30):
        pos, _, dist = ix.search(queries, k)
        for i in range(0, 40, 3):
            opos, odist, nd = po.scan(payload, n, cw, queries[i], k, want_node_dist=True)
            assert np.array_equal(dist[i], odist)
            assert_topk_equal(pos[i], dist[i], opos, odist, node_dist=nd)
    ix.close()


def test_large_batch_is_split(golden1501, engine):
    """More than 32768 queries in one dpq_index_search call are processed in sub-batches."""
    g = golden1501
    rng = np.random.default_rng(3)
    Q = 40000
    queries = np.clip(g["queries"][rng.integers(0, len(g["queries"]), Q)] + rng.integers(-20, 21, (Q, 128)), 0, 255).astype(np.float32)
    ix = _open(g)
    pos, ids, dist = ix.search(queries, 5)
    for i in (0, 32767, 32768, 39999):
        opos, odist = po.scan(g["payload"], int(g["n"]), g["cw"], queries[i], 5)
        assert np.array_equal(dist[i], odist)
    ix.close()


def test_many_queries_ragged_groups(golden4000, engine):
    """Q not a multiple of the group size (56 / 48..52): padding lanes must stay silent."""
    g = golden4000
    n = int(g["n"])
    rng = np.random.default_rng(17)
    base_q = g["queries"]
    queries = np.clip(base_q[rng.integers(0, len(base_q), 300)] + rng.integers(-9, 10, (300, 128)), 0, 255).astype(np.float32)
    ix = _open(g)
    for Q in (1, 55, 57, 113, 300):
        pos, _, dist = ix.search(queries[:Q], 10)
        for i in range(0, Q, max(1, Q // 9)):
            opos, odist, nd = po.scan(g["payload"], n, g["cw"], queries[i], 10, want_node_dist=True)
            assert np.array_equal(dist[i], odist)
            assert_topk_equal(pos[i], dist[i], opos, odist, node_dist=nd)
    ix.close()


def test_forced_exact_fallback(golden4000):
    g = golden4000
    a = _open(g)
    b = _open(g, force_fallback=1)
    pa = a.search(g["queries"], 10)
    pb = b.search(g["queries"], 10)
    assert b.stat("last_fallback") == len(g["queries"]) and a.stat("last_fallback") == 0
    assert np.array_equal(pa[2], pb[2]) and np.array_equal(pa[0], pb[0])
    a.close(), b.close()


def test_duplicates_and_ties():
    """Heavy exact ties: many duplicate codes.  Distances must match the oracle exactly and
    ids modulo ties; own tie rule = lower position first."""
    rng = np.random.default_rng(7)
    M, K, n = 8, 256, 3000
    uniq = rng.integers(0, K, (40, M)).astype(np.uint8)
    codes = uniq[rng.integers(0, 40, n)]
    cw = dg.roundtrip_codebook((rng.random((M, K, 4)) * 50).astype(np.float32))
    _, _, lay, payload = po.build_tree(codes, cw)
    ix = dpq.DeltaTreeIndex(payload, n, M, K)
    ix.set_codebook(cw)
    queries = (rng.random((9, M * 4)) * 50).astype(np.float32)
    for pack in (1, 2):
        ix.set_option("pack", pack)
        pos, _, dist = ix.search(queries, 20)
        for i, q in enumerate(queries):
            opos, odist, nd = po.scan(payload, n, cw, q, 20, want_node_dist=True)
            assert np.array_equal(dist[i], odist)
            assert_topk_equal(pos[i], dist[i], opos, odist, node_dist=nd)
            # deterministic tie rule: (distance, position) ascending
            order = np.lexsort((pos[i], dist[i]))
            assert np.array_equal(order, np.arange(20))
    ix.close()


def test_tiny_trees():
    rng = np.random.default_rng(9)
    M, K = 8, 256
    cw = dg.roundtrip_codebook((rng.random((M, K, 2)) * 9).astype(np.float32))
    for n in (1, 2, 3, 7):
        codes = rng.integers(0, K, (n, M)).astype(np.uint8)
        if n > 1:
            _, _, lay, payload = po.build_tree(codes, cw)
        else:
            payload = codes[0].copy()
        ix = dpq.DeltaTreeIndex(payload, n, M, K)
        ix.set_codebook(cw)
        q = (rng.random((3, M * 2)) * 9).astype(np.float32)
        pos, _, dist = ix.search(q, 4)
        for i in range(3):
            opos, odist = po.scan(payload, n, cw, q[i], min(4, n))
            assert np.array_equal(dist[i][: min(4, n)], odist)
            assert np.all(pos[i][min(4, n):] == 0xFFFFFFFF)
        ix.close()


def test_sharded_search_merges_to_whole(golden4000):
    """SURVEY 8e on one GPU: 4 shards opened as 4 indexes, local top-k lists merged by the
    merge kernel == the unsharded result (positions are global)."""
    g = golden4000
    n, k, R = int(g["n"]), 10, 4
    Q = len(g["queries"])
    whole = _open(g)
    wpos, _, wdist = whole.search(g["queries"], k)
    dq = dpq.DeviceBuffer(g["queries"].nbytes).upload(g["queries"])
    dk = dpq.DeviceBuffer(R * Q * k * 8)
    do = dpq.DeviceBuffer(Q * k * 8)
    shards = []
    for r in range(R):
        ix = dpq.DeltaTreeIndex(g["payload"], n, 8, 256, rank=r, n_ranks=R)
        ix.set_codebook(g["cw"])
        ix.search_device(dq.ptr, Q, k, dk.ptr.value + r * Q * k * 8)
        ix.sync()
        shards.append(ix)
    assert sum(s.stat("n_local") for s in shards) == n
    shards[0].merge_device(dk.ptr, R, Q, k, do.ptr)
    shards[0].sync()
    pos, dist = dpq.unpack_keys(do.download(np.uint64, (Q, k)))
    assert np.array_equal(dist, wdist) and np.array_equal(pos, wpos)
    for s in shards:
        s.close()
    whole.close()


def test_forest_parts_merge_to_plain_adc(engine):
    """Config C5 layout on one GPU: the code set cut into 3 parts by vector id, one DeltaTree per
    part (dpq_tree_build), opened with dpq_index_open_part; the merged top-k must be the top-k
    of plain ADC over ALL codes (oracle tables: float entries, double sum), ids global."""
    if engine == "gen1":  # the first-generation program hard-codes the root at position 0
        with pytest.raises(dpq.DpqError):
            dpq.DeltaTreeIndex(np.arange(8, dtype=np.uint8), 1, 8, 256, first_pos=5)
        return
    base = dg.sift_like(9000, 128, seed=21)
    cw = dg.roundtrip_codebook(dg.kmeans_codebook(dg.sift_like(3000, 128, seed=22), 8, 256, iters=3))
    queries = dg.sift_like(40, 128, seed=23)
    codes = dpq.encode(cw, base)
    k, Q = 10, len(queries)
    cuts = [0, 2500, 6001, 9000]
    dq = dpq.DeviceBuffer(queries.nbytes).upload(queries)
    dk = dpq.DeviceBuffer(3 * Q * k * 8)
    do = dpq.DeviceBuffer(Q * k * 8)
    parts, id_of_pos = [], np.zeros(9000, np.int64)
    for p in range(3):
        a, b = cuts[p], cuts[p + 1]
        t = dpq.tree_build(codes[a:b], cw, want=("payload", "vec_id"))
        ix = dpq.DeltaTreeIndex(t["payload"], b - a, 8, 256, pos2id=t["vec_id"], first_pos=a)
        ix.set_codebook(cw)
        ix.search_device(dq.ptr, Q, k, dk.ptr.value + p * Q * k * 8)
        ix.sync()
        id_of_pos[a:b] = t["vec_id"].astype(np.int64) + a
        # the host-buffer path reports the same shifted positions and the part-local ids
        hpos, hid, hdist = ix.search(queries, k)
        lp, ld = dpq.unpack_keys(dk.download(np.uint64, (3, Q, k))[p])
        assert np.array_equal(hpos, lp) and np.array_equal(hdist, ld)
        assert hpos.min() >= a and hpos.max() < b
        assert np.array_equal(hid, t["vec_id"][hpos - a])
        parts.append(ix)
    parts[0].merge_device(dk.ptr, 3, Q, k, do.ptr)
    parts[0].sync()
    pos, dist = dpq.unpack_keys(do.download(np.uint64, (Q, k)))
    ids = id_of_pos[pos]
    for i, q in enumerate(queries):
        tab = po.lut(cw, q).astype(np.float64)                      # [M][K]
        d_all = tab[np.arange(8)[None, :], codes].sum(axis=1).astype(np.float32)
        order = np.lexsort((np.arange(9000), d_all))[:k]
        assert_topk_equal(ids[i], dist[i], order, d_all[order], node_dist=d_all)
    for ix in parts:
        ix.close()


def test_m16_extension_search(golden_m16):
    """Configs 3/4 (M=16): no reference tree oracle exists; parity is
    against the generalised restatement, top-100."""
    g = golden_m16
    codes, cw = g["codes"], g["cw"]
    _, _, lay, payload = po.build_tree(codes, cw)
    ix = dpq.DeltaTreeIndex(payload, len(codes), 16, 256, pos2id=lay["vec_id"])
    ix.set_codebook(cw)
    for pack in (1, 2):
        ix.set_option("pack", pack)
        pos, ids, dist = ix.search(g["queries"], 100)
        for i, q in enumerate(g["queries"]):
            opos, odist, nd = po.scan(payload, len(codes), cw, q, 100, want_node_dist=True)
            np.testing.assert_allclose(dist[i], odist, rtol=REL_TOL)
            assert_topk_equal(pos[i], dist[i], opos, odist, node_dist=nd)
    ix.close()


@pytest.mark.parametrize("M,K,Ds", [(4, 256, 4), (8, 100, 3), (3, 16, 5), (12, 256, 2), (16, 200, 2), (9, 64, 4)])
def test_search_other_shapes(M, K, Ds, engine):
    """M < 8 (narrow records padded with the zero row), 8 < M <= 16 (wide records), K < 256."""
    rng = np.random.default_rng(M * 131 + K)
    n, k = 2500, 12
    cw = dg.roundtrip_codebook((rng.random((M, K, Ds)) * 40).astype(np.float32))
    base = (rng.random((n, M * Ds)) * 40).astype(np.float32)
    codes = po.encode(cw, base)
    codes[rng.random(n) < 0.2] = codes[0]  # duplicates
    _, _, lay, payload = po.build_tree(codes, cw)
    ix = dpq.DeltaTreeIndex(payload, n, M, K, pos2id=lay["vec_id"])
    ix.set_codebook(cw)
    assert ix.stat("engine") == (2 if engine == "v2" else 1)
    queries = (rng.random((70, M * Ds)) * 40).astype(np.float32)
    for coarse in ((0, 1) if engine == "v2" else (0,)):  # 1: forced three-phase coarse search (narrow and wide shapes)
        ix.set_option("coarse", coarse)
        pos, ids, dist = ix.search(queries, k)
        if engine == "v2":
            assert ix.stat("last_coarse") == coarse
        for i in range(0, 70, 6):
            opos, odist, nd = po.scan(payload, n, cw, queries[i], k, want_node_dist=True)
            np.testing.assert_allclose(dist[i], odist, rtol=REL_TOL)
            assert_topk_equal(pos[i], dist[i], opos, odist, node_dist=nd)
            assert np.array_equal(ids[i], lay["vec_id"][pos[i]])
    ix.close()


@pytest.mark.parametrize("opts", [dict(coarse=1), dict(coarse=1, sample=2), dict(coarse=1, bcap8=32),
                                  dict(coarse=1, levels8=123), dict(coarse=1, seed=1)])
def test_coarse_search_wide_shape_is_exact(golden_m16, engine, opts):
    """The wide coarse scan (M = 16: 48 queries per CTA, entries saturating at 15, sixteen reads per
    node) + exact re-score == the oracle, for top-10 and top-100 (configs C3 / C4)."""
    if engine != "v2":
        pytest.skip("coarse search belongs to the v2 engine")
    g = golden_m16
    codes, cw = g["codes"], g["cw"]
    n = len(codes)
    _, _, lay, payload = po.build_tree(codes, cw)
    ix = dpq.DeltaTreeIndex(payload, n, 16, 256, pos2id=lay["vec_id"])
    ix.set_codebook(cw)
    for name, v in opts.items():
        ix.set_option(name, v)
    rng = np.random.default_rng(5)
    queries = np.concatenate([g["queries"], np.clip(g["queries"] + rng.integers(-9, 10, g["queries"].shape), 0, 255)]).astype(np.float32)
    for k in (1, 10, 100, 128):
        pos, ids, dist = ix.search(queries, k)
        assert ix.stat("last_coarse") == 1
        for i in range(0, len(queries), 3):
            opos, odist, nd = po.scan(payload, n, cw, queries[i], k, want_node_dist=True)
            np.testing.assert_allclose(dist[i], odist, rtol=REL_TOL)
            assert_topk_equal(pos[i], dist[i], opos, odist, node_dist=nd)
    if opts.get("bcap8") == 32:
        assert ix.stat("last_fallback") > 0
    ix.close()


def test_coarse_search_narrow_top100(golden4000, engine):
    """topk in (64, 128] now also takes the coarse path on the narrow shape."""
    if engine != "v2":
        pytest.skip("coarse search belongs to the v2 engine")
    g = golden4000
    ix = _open(g, coarse=1)
    pos, ids, dist = ix.search(g["queries"], 100)
    assert ix.stat("last_coarse") == 1
    for i in range(0, len(g["queries"]), 5):
        opos, odist, nd = po.scan(g["payload"], int(g["n"]), g["cw"], g["queries"][i], 100, want_node_dist=True)
        assert np.array_equal(dist[i], odist)
        assert_topk_equal(pos[i], dist[i], opos, odist, node_dist=nd)
    ix.close()


def test_encode_bit_exact(golden4000, golden_m16):
    for g in (golden4000, golden_m16):
        assert np.array_equal(dpq.encode(g["cw"], g["base_head"]), g["codes"][:256])
    rng = np.random.default_rng(2)
    cw = dg.roundtrip_codebook(rng.random((16, 256, 60)).astype(np.float32))  # GIST shape
    x = rng.random((700, 960)).astype(np.float32)
    assert np.array_equal(dpq.encode(cw, x), po.encode(cw, x))
    cw = dg.roundtrip_codebook(rng.random((4, 100, 7)).astype(np.float32))   # odd Ds, padding
    x = rng.random((33, 27)).astype(np.float32)
    assert np.array_equal(dpq.encode(cw, x), po.encode(cw, x))


def test_encode_tensor_core_filter_is_exact():
    """encode_tc.cu: the tcgen05 scores only FILTER; the code written is the reference loop's
    (pq_tree.cpp:215-237) on inputs built to break a filter: duplicated centroids (ties go to the
    lowest id), vectors equal to centroids, midpoints of centroid pairs (exact ties in real
    arithmetic), near-ties a few ulps apart, magnitudes from 1e-30 to 1e18, zero vectors, zero
    padding (D < M * Ds), K < 256, ragged n, NaN / inf components."""
    rng = np.random.default_rng(11)
    for M, K, Ds, n in ((8, 256, 16, 3001), (16, 256, 8, 1500), (4, 100, 16, 700), (8, 7, 4, 513), (3, 33, 8, 129)):
        cw = rng.normal(size=(M, K, Ds)).astype(np.float32) * 40 + 60
        cw[:, K // 2] = cw[:, 1]                      # duplicate centroid
        if K > 6:
            cw[:, 5] = np.nextafter(cw[:, 4], np.float32(np.inf))   # one ulp apart
        D = M * Ds
        x = rng.normal(size=(n, D)).astype(np.float32) * 40 + 60
        xs = x.reshape(n, M, Ds)
        pick = rng.integers(0, K, size=(n, M))
        own = cw[np.arange(M)[None, :], pick]         # [n][M][Ds]
        xs[0:200] = own[0:200]                        # exactly a centroid (possibly the duplicated one)
        other = cw[np.arange(M)[None, :], (pick + 1) % K]
        xs[200:400] = (own[200:400] + other[200:400]) * np.float32(0.5)   # midpoints
        xs[400:420] = 0.0
        x = xs.reshape(n, D)
        for scale in (1.0, 1e-3, 1e4):
            cws, xq = (cw * np.float32(scale)).astype(np.float32), (x * np.float32(scale)).astype(np.float32)
            got = dpq.encode(cws, xq)
            assert dpq.encode_stat("tc") == 1
            assert np.array_equal(got, po.encode(cws, xq)), (M, K, Ds, scale)
        # the bound cannot be trusted out here: every centroid becomes a candidate, still the reference's answer
        for cs, xsc in ((1e-30, 1e-30), (1e18, 1e18), (1e-30, 1e6), (1e10, 1e-20)):
            cws, xq = (cw * np.float32(cs)).astype(np.float32), (x[:300] * np.float32(xsc)).astype(np.float32)
            assert np.array_equal(dpq.encode(cws, xq), po.encode(cws, xq)), (M, K, Ds, cs, xsc)
        xq = x[:300].copy()
        xq[3, 1] = np.nan
        xq[7, 0] = np.inf
        xq[9, D - 1] = -np.inf
        cwn = cw.copy()
        cwn[0, 2, 0] = np.nan
        assert np.array_equal(dpq.encode(cw, xq), po.encode(cw, xq))
        assert np.array_equal(dpq.encode(cwn, xq), po.encode(cwn, xq))
        # zero padding: D < M * Ds
        for cut in (3, 4):
            xc = x[:257, :D - cut].copy()
            assert np.array_equal(dpq.encode(cw, xc), po.encode(cw, xc))
    # SIFT-shaped integers (every component exact in bf16 hi + lo), and the SIMT kernel on the same input
    cw = dg.roundtrip_codebook((rng.random((8, 256, 16)) * 140).astype(np.float32))
    x = rng.integers(0, 256, size=(40000, 128)).astype(np.float32)
    got = dpq.encode(cw, x)
    assert dpq.encode_stat("tc") == 1 and dpq.encode_stat("kernel_us") >= 0
    os.environ["DPQ_ENCODE_TC"] = "0"
    try:
        simt = dpq.encode(cw, x)
        assert dpq.encode_stat("tc") == 0
    finally:
        del os.environ["DPQ_ENCODE_TC"]
    assert np.array_equal(got, simt)
    assert np.array_equal(got[:4000], po.encode(cw, x[:4000]))


def test_search_into_page_locked_result_arrays(golden4000):
    """dpq_index_search: caller arrays that are page-locked are written by the unpack kernel itself;
    pageable ones go through the staging buffer -- same content either way, also mixed."""
    g = golden4000
    ix = _open(g)
    q = g["queries"]
    Q, k = len(q), 10
    ref = ix.search(q, k)
    pinned = tuple(dpq.pinned_array((Q, k), dt) for dt in (np.uint32, np.uint32, np.float32))
    for a in pinned:
        a[...] = 0
    got = ix.search(q, k, out=pinned)
    for a, b in zip(ref, got):
        assert np.array_equal(a, b)
    mixed = (np.zeros((Q, k), np.uint32), dpq.pinned_array((Q, k), np.uint32), np.zeros((Q, k), np.float32))
    got = ix.search(q, k, out=mixed)
    for a, b in zip(ref, got):
        assert np.array_equal(a, b)
    ix.close()


def test_edge_diffs(golden4000):
    g = golden4000
    bm, nd = dpq.edge_diffs(g["codes"], g["edges"])
    a, b = g["codes"][g["edges"][:, 0]], g["codes"][g["edges"][:, 1]]
    want = ((a != b) * (1 << np.arange(8))).sum(1).astype(np.uint32)
    assert np.array_equal(bm, want)
    assert nd == int((a != b).sum()) == len(g["payload"]) - 8 - (3 * (int(g["n"]) - 1) + 1) // 2


def test_groundtruth_exact():
    base = dg.sift_like(20000, 128, seed=4)
    qs = dg.sift_like(12, 128, seed=5)
    ids, dist = dpq.groundtruth(base, qs, 10, chunk=7000)
    oid, odist = po.groundtruth(base, qs, 10, chunk=7000)
    assert np.array_equal(dist, odist)
    assert_topk_equal(ids, dist, oid, odist)
    g = dg.gist_like(3000, 960, seed=6)
    gq = dg.gist_like(5, 960, seed=7)
    ids, dist = dpq.groundtruth(g, gq, 5)
    oid, odist = po.groundtruth(g, gq, 5)
    assert np.array_equal(dist, odist)
    assert_topk_equal(ids, dist, oid, odist)


def test_medium_tree_end_to_end():
    """50K codes: encode on the GPU (bit-exact), oracle-built tree, search vs oracle."""
    base = dg.sift_like(50000, 128, seed=21)
    cw = dg.roundtrip_codebook(dg.kmeans_codebook(dg.sift_like(8000, 128, seed=22), 8, 256, iters=4))
    codes = dpq.encode(cw, base)
    assert np.array_equal(codes[:2000], po.encode(cw, base[:2000]))
    _, _, lay, payload = po.build_tree(codes, cw)
    ix = dpq.DeltaTreeIndex(payload, len(codes), 8, 256, pos2id=lay["vec_id"])
    ix.set_codebook(cw)
    queries = dg.sift_like(64, 128, seed=23)
    for pack in (1, 2):
        ix.set_option("pack", pack)
        ix.set_option("coarse", pack - 1)  # second round: the three-phase coarse search
        pos, ids, dist = ix.search(queries, 10)
        # (forcing the coarse search on a 50K tree gives a loose cap from a 3K-node sample: a few
        # candidate buffers may overflow into the exact fallback; results are checked either way)
        assert ix.stat("last_fallback") <= (2 if pack == 1 else 16)
        for i in range(0, 64, 7):
            opos, odist, nd = po.scan(payload, len(codes), cw, queries[i], 10, want_node_dist=True)
            assert np.array_equal(dist[i], odist)
            assert_topk_equal(pos[i], dist[i], opos, odist, node_dist=nd)
    ix.close()


@pytest.mark.parametrize("shape", [
    # n, D, Q, topk, generator
    (30000, 128, 130, 10, "sift"),      # Q not a multiple of 128, n not a multiple of 256
    (9000, 960, 40, 5, "gist"),         # non-integer values: the bf16 hi + lo split and its error bound
    (12345, 100, 300, 64, "ints"),      # D not a multiple of 64
    (7000, 30, 17, 3, "gauss"),         # D not a multiple of 4, signed values
    (3000, 128, 9, 10, "sift"),         # fewer vectors than the dense seed
    (40000, 128, 64, 10, "dupes"),      # heavy exact ties
])
def test_groundtruth_tensor_core_filter_equals_plain_kernels(shape, monkeypatch):
    """dpq_groundtruth_*: the tcgen05 filter + exact re-score path must give exactly what the
    plain exact kernels (DPQ_GT_TC=0) and the oracle (pmain:138-166 arithmetic) give."""
    n, D, Q, k, gen = shape
    rng = np.random.default_rng(n + D)
    if gen == "sift":
        base, qs = dg.sift_like(n, D, seed=31), dg.sift_like(Q, D, seed=32)
    elif gen == "gist":
        base, qs = dg.gist_like(n, D, seed=33), dg.gist_like(Q, D, seed=34)
    elif gen == "ints":
        base = rng.integers(0, 256, size=(n, D)).astype(np.float32)
        qs = rng.integers(0, 256, size=(Q, D)).astype(np.float32)
    elif gen == "gauss":
        base, qs = rng.normal(size=(n, D)).astype(np.float32), rng.normal(size=(Q, D)).astype(np.float32)
    else:
        pool = dg.sift_like(500, D, seed=35)
        base, qs = pool[rng.integers(0, 500, n)], dg.sift_like(Q, D, seed=36)
    monkeypatch.delenv("DPQ_GT_TC", raising=False)
    st = {}
    ids, dist = dpq.groundtruth(base, qs, k, chunk=17000, stats=st)
    assert st["tc"] == 2                              # pipelined form (TMA + warp-specialised roles)
    monkeypatch.setenv("DPQ_GT_TC", "1")              # synchronous form of the same filter
    st1 = {}
    sid, sdist = dpq.groundtruth(base, qs, k, chunk=17000, stats=st1)
    assert st1["tc"] == 1 and np.array_equal(sid, ids) and np.array_equal(sdist, dist)
    seed = min(8192, max(4096, 64 * k))
    if n > seed:
        assert st["tc_vectors"] == n - seed          # everything after the dense seed went through the filter
        if gen != "dupes":
            assert st["tc_flagged"] < Q              # ... and the candidate lists held
    monkeypatch.setenv("DPQ_GT_TC", "0")
    st0 = {}
    pid, pdist = dpq.groundtruth(base, qs, k, chunk=17000, stats=st0)
    assert st0["tc"] == 0
    assert np.array_equal(dist, pdist) and np.array_equal(ids, pid)
    oid, odist = po.groundtruth(base[:, :], qs[: min(Q, 24)], k, chunk=17000)
    assert np.array_equal(dist[: min(Q, 24)], odist)
    assert_topk_equal(ids[: min(Q, 24)], dist[: min(Q, 24)], oid, odist)


def test_encode_u8_bvecs_records_equal_float_encode(tmp_path):
    """dpq_encode_u8: raw .bvecs records (4-byte header + D bytes) converted on the device ==
    dpq_encode of the same components as floats == the oracle (pq_tree.cpp:192-253)."""
    base = dg.sift_like(5000, 128, seed=41)
    cw = dg.roundtrip_codebook(dg.kmeans_codebook(dg.sift_like(3000, 128, seed=42), 8, 256, iters=3))
    path = str(tmp_path / "base.bvecs")
    dg.write_vecs(path, base, ext="bvecs")
    raw = np.fromfile(path, np.uint8)
    assert raw.size == 5000 * 132
    codes = dpq.encode_u8(cw, raw, 5000, 128, 132, 4)
    assert np.array_equal(codes, dpq.encode(cw, base))
    assert np.array_equal(codes[:1500], po.encode(cw, base[:1500]))


@pytest.mark.parametrize("shape", [(120000, 8, 10), (9000, 8, 64), (5000, 16, 100), (1, 8, 1), (65, 8, 5)])
def test_index_from_tree_arrays_equals_index_from_stream(shape, engine):
    """dpq_index_open_tree compiles the scan program on the GPU from the layout arrays; the index
    must answer exactly like the one compiled on the host from the byte stream (program.cpp)."""
    n, M, k = shape
    base = dg.sift_like(n, 128, seed=51)
    cw = dg.roundtrip_codebook(dg.kmeans_codebook(dg.sift_like(3000, 128, seed=52), M, 256, iters=3))
    queries = dg.sift_like(150, 128, seed=53)
    codes = dpq.encode(cw, base)
    # the first-generation program (DPQ_ENGINE=1) cannot shift positions; open_tree then goes through the stream
    shift = 7 if engine == "v2" else 0
    t = dpq.tree_build(codes, cw, want=("payload", "vec_id"), open_index_at=shift)
    a = t["index"]
    b = dpq.DeltaTreeIndex(t["payload"], n, M, 256, pos2id=t["vec_id"], first_pos=shift if shift else None)
    for ix in (a, b):
        ix.set_codebook(cw)
    apos, aid, adist = a.search(queries, min(k, 256))
    bpos, bid, bdist = b.search(queries, min(k, 256))
    assert np.array_equal(apos, bpos) and np.array_equal(aid, bid) and np.array_equal(adist, bdist)
    assert apos[apos != 0xFFFFFFFF].min() >= shift
    for name in ("n_codes", "n_bytes", "n_local", "n_diffs", "n_chunks", "v2_delta_nodes", "engine"):
        assert a.stat(name) == b.stat(name), name
    a.close()
    b.close()


@pytest.mark.parametrize("n,M", [(30000, 8), (7000, 16), (5000, 4)])
def test_device_resident_tree_and_its_shards(n, M, engine):
    """dpq_tree_build_device leaves the tree in HBM (the 10^9-code layout): same stream, vec_id and
    codes as the host-resident build, and dpq_index_open_tree_shard deals the same depth-1-subtree
    shards as the stream reader does (dpq_index_open(payload, rank, n_ranks))."""
    if engine == "gen1":
        pytest.skip("a device-resident tree needs the code-array engine")
    base = dg.sift_like(n, 128, seed=61)
    cw = dg.roundtrip_codebook(dg.kmeans_codebook(dg.sift_like(3000, 128, seed=62), M, 256, iters=3))
    queries = dg.sift_like(200, 128, seed=63)
    codes = dpq.encode(cw, base)
    host = dpq.tree_build(codes, cw, want=("payload", "vec_id", "codes_by_pos", "depth"))
    dcodes = dpq.DeviceBuffer(codes.nbytes).upload(codes)
    dt = dpq.DeviceTree(dcodes.ptr.value, n, M, cw)          # codes passed as a DEVICE pointer
    assert dt.stat("on_device") == 1 and dt.stat("n_diffs") == host["n_diffs"] and dt.stat("root_id") == host["root_id"]
    assert np.array_equal(dt.fetch("payload", np.uint8), host["payload"])
    assert np.array_equal(dt.fetch("vec_id", np.uint32), host["vec_id"])
    assert np.array_equal(dt.fetch("depth", np.uint8), host["depth"])
    assert np.array_equal(dt.fetch("codes_by_pos", np.uint8).reshape(-1, M), host["codes_by_pos"])
    assert sum(dt.stat(f"depth_hist_{d}") for d in range(17)) == n
    k = 10
    for R in (1, 3):
        for r in range(R):
            a = dt.shard(r, R)
            b = dpq.DeltaTreeIndex(host["payload"], n, M, 256, pos2id=host["vec_id"], rank=r, n_ranks=R)
            for name in ("n_codes", "n_local", "base_pos", "n_diffs", "n_chunks", "engine"):
                assert a.stat(name) == b.stat(name), (name, r, R)
            assert abs(a.stat("n_bytes") - b.stat("n_bytes")) <= 1  # the shard's depth nibbles round differently
            for ix in (a, b):
                ix.set_codebook(cw)
            apos, aid, adist = a.search(queries, k)
            bpos, bid, bdist = b.search(queries, k)
            assert np.array_equal(apos, bpos) and np.array_equal(aid, bid) and np.array_equal(adist, bdist)
            a.close()
            b.close()
    dt.free()
    dcodes.free()


@pytest.mark.parametrize("name", ["gist_gt_n3000_d960", "gist_gt_n5000_d96"])
@pytest.mark.parametrize("tc", ["1", "0"])
def test_groundtruth_vs_reference_binary(name, tc, monkeypatch):
    """dpq_groundtruth_* (tensor-core filter + exact re-score, and the plain exact kernels) against
    what the unmodified reference's `pqtree -task groundtruth` wrote for GIST-shaped non-integer
    floats (tests/golden/make_golden_gt.py): same ids, same printed distances."""
    monkeypatch.setenv("DPQ_GT_TC", tc)
    base, queries, k, ref_ids, ref_text = gt_fixture(name)
    ids, dist = dpq.groundtruth(base, queries, k, chunk=1700)
    assert_gt_equals_reference_text(ids, dist, ref_ids, ref_text)
    oid, odist = po.groundtruth(base, queries, k)
    assert np.array_equal(dist, odist)


def test_gpu_stream_decoder_equals_host_decoder(golden4000, golden1501, golden_m16, monkeypatch, engine):
    """dpq_index_open decodes the on-disk stream on the GPU (program_dev.cu: speculative record
    boundaries, per-level parent scans, level-order codes); DPQ_HOST_DECODE=1 keeps the sequential
    host decoder (program.cpp).  Same shards, same statistics, same answers; malformed streams are
    refused by both."""
    if engine == "gen1":
        pytest.skip("the first-generation program is compiled on the host")
    cases = [(golden4000["payload"], 4000, 8, golden4000["cw"], golden4000["queries"], golden4000["vec_id"]),
             (golden1501["payload"], 1501, 8, golden1501["cw"], golden1501["queries"], golden1501["vec_id"])]
    g = golden_m16
    _, _, lay, payload16 = po.build_tree(g["codes"], g["cw"])
    cases.append((payload16, len(g["codes"]), 16, g["cw"], g["queries"], lay["vec_id"]))
    # a bigger tree: many 4 KB stream blocks, every depth
    base = dg.sift_like(60000, 128, seed=71)
    cw = dg.roundtrip_codebook(dg.kmeans_codebook(dg.sift_like(3000, 128, seed=72), 8, 256, iters=3))
    t = dpq.tree_build(dpq.encode(cw, base), cw, want=("payload", "vec_id"))
    cases.append((t["payload"], 60000, 8, cw, dg.sift_like(100, 128, seed=73), t["vec_id"]))
    for payload, n, M, cw_, queries, vec_id in cases:
        for R in (1, 3):
            for r in range(R):
                monkeypatch.delenv("DPQ_HOST_DECODE", raising=False)
                a = dpq.DeltaTreeIndex(payload, n, M, 256, pos2id=vec_id, rank=r, n_ranks=R)
                monkeypatch.setenv("DPQ_HOST_DECODE", "1")
                b = dpq.DeltaTreeIndex(payload, n, M, 256, pos2id=vec_id, rank=r, n_ranks=R)
                for name in ("n_codes", "n_local", "base_pos", "n_diffs", "n_chunks", "engine"):
                    assert a.stat(name) == b.stat(name), (name, n, r, R)
                assert abs(a.stat("n_bytes") - b.stat("n_bytes")) <= 1
                assert [a.stat(f"depth_hist_{d}") for d in range(16)] == [b.stat(f"depth_hist_{d}") for d in range(16)]
                for ix in (a, b):
                    ix.set_codebook(cw_)
                apos, aid, adist = a.search(queries, 10)
                bpos, bid, bdist = b.search(queries, 10)
                assert np.array_equal(apos, bpos) and np.array_equal(aid, bid) and np.array_equal(adist, bdist)
                a.close()
                b.close()
    monkeypatch.delenv("DPQ_HOST_DECODE", raising=False)
    payload, n = golden1501["payload"], 1501
    with pytest.raises(dpq.DpqError, match="trunc|mismatch|missing"):
        dpq.DeltaTreeIndex(payload[:-3], n, 8, 256)
    with pytest.raises(dpq.DpqError, match="trunc|mismatch|missing"):
        dpq.DeltaTreeIndex(np.concatenate([payload, np.zeros(5, np.uint8)]), n, 8, 256)
    bad = payload.copy()
    bad[8] = 0x77  # depth jump 0 -> 7
    with pytest.raises(dpq.DpqError, match="depth"):
        dpq.DeltaTreeIndex(bad, n, 8, 256)
    with pytest.raises(dpq.DpqError, match="centroid"):
        dpq.DeltaTreeIndex(payload, n, 8, 3)   # K = 3: the stream's centroid ids do not fit
    # a one-node tree and a two-node tree
    one = dpq.DeltaTreeIndex(np.arange(8, dtype=np.uint8), 1, 8, 256)
    assert one.stat("n_local") == 1
    one.close()


def test_multi_index_single_gpu_and_probe(golden4000, tmp_path, engine):
    """dpq_multi_* through ctypes with one GPU (the path `deltapq -task query -gpus N` and
    bench.py's multi_cpp block use; with N > 1 it all-gathers over NCCL): equals the plain index."""
    import json
    import subprocess
    import sys
    import os
    g = golden4000
    n = int(g["n"])
    tree, qnode = str(tmp_path / "tree.bin"), str(tmp_path / "qnodes.bin")
    with open(tree, "wb") as f:
        f.write(np.array([n, len(g["payload"])], np.int64).tobytes())
        f.write(g["payload"].tobytes())
    qn = np.zeros((n + 1, 60), np.uint8)
    qn[:n, 0:4] = g["vec_id"].astype(np.uint32).view(np.uint8).reshape(-1, 4)
    qn.tofile(qnode)
    mx = dpq.MultiIndex(tree, 8, 256, 1, qnode_path=qnode)
    mx.set_codebook(g["cw"])
    pos, ids, dist = mx.search(g["queries"], 10)
    assert mx.stat(0, "n_local") == n
    mx.close()
    ix = _open(g)
    opos, oids, odist = ix.search(g["queries"], 10)
    ix.close()
    assert np.array_equal(pos, opos) and np.array_equal(ids, oids) and np.array_equal(dist, odist)
    if engine == "v2":
        np.savez(str(tmp_path / "multi_probe.npz"), cw=g["cw"], queries=g["queries"], topk=10, n_codes=n)
        root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
        r = subprocess.run([sys.executable, os.path.join(root, "tools", "multi_probe.py"), str(tmp_path), "1", "2"],
                           capture_output=True, text=True, timeout=300)
        assert r.returncode == 0, r.stderr[-500:]
        out = json.loads([ln for ln in r.stdout.splitlines() if ln.startswith("{")][-1])
        assert out["equals_single_gpu"] and out["shard_nodes"] == [n]


@pytest.mark.parametrize("scale", [1e-22, 3e-20, 1.0, 2e18, 1.8e19])
def test_adc_tables_extreme_ranges(scale, engine):
    """The ADC-table kernels round their accumulator to float inside the double domain; partial sums
    in the float-subnormal range and from 2^127 up take the conversion path instead.  Both must equal
    the oracle bit for bit (dpq_adc_tables = adc_entry; a search = lut2_kernel / lut_small_kernel), so
    the data is scaled into those ranges."""
    rng = np.random.default_rng(17)
    M, K, Ds, n = 8, 256, 4, 5000
    cw = (rng.random((M, K, Ds)).astype(np.float32) * np.float32(scale)).astype(np.float32)
    queries = (rng.random((40, M * Ds)).astype(np.float32) * np.float32(scale)).astype(np.float32)
    if scale > 1e19:  # one term per entry, up to 3.2e38: the top binade of float without overflowing the sum
        cw[:, :, 1:] = 0
        queries.reshape(40, M, Ds)[:, :, 1:] = 0
    queries[0] = cw[:, 3, :].reshape(-1)          # exact zeros in every subspace
    lut = dpq.adc_tables(cw, queries)
    for i in range(0, 40, 7):
        assert np.array_equal(lut[i], po.lut(cw, queries[i])), (scale, i)
    if scale > 1e19:
        assert np.isfinite(lut).all() and (lut > 1.8e38).any()
        return  # a node's eight entries would overflow float: tables only
    codes = rng.integers(0, K, (n, M)).astype(np.uint8)
    _, _, lay, payload = po.build_tree(codes, cw)
    ix = dpq.DeltaTreeIndex(payload, n, M, K, pos2id=lay["vec_id"])
    ix.set_codebook(cw)
    for qs in (queries, queries[:3]):              # batched tables and the latency-mode tables
        pos, ids, dist = ix.search(qs, 5)
        for i in range(0, len(qs), 9):
            opos, odist, nd = po.scan(payload, n, cw, qs[i], 5, want_node_dist=True)
            assert np.array_equal(dist[i], odist), (scale, i, dist[i], odist)
    ix.close()
