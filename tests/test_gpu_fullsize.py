"""GPU parity at the sizes and on the code path the benchmarks time: the AUTOMATIC search path
(no `coarse=` override: three-phase coarse search from 100K nodes per shard, sample stride 8 / 16 /
32 by tree size) against the oracle scan (DCAT.h:3731-3892 restated in oracle/dpq_oracle.c) on
the same trees, for BASELINE configs C2 (1M codes, M=8, top-10), C3 shape (M=16, top-100) and C4
shape (960-d floats, M=16), plus a forest of parts above the coarse threshold.

Bars: distances bit-equal where the oracle's double accumulation is exact (integer SIFT-shaped
data), else within 1e-5 relative; ids modulo ties at 1e-5 (helpers.assert_topk_equal); every
reported (position, distance) is checked against the oracle's per-node distance."""
import numpy as np
import pytest

import datagen as dg
import deltapq_b200 as dpq
from helpers import assert_topk_equal, REL_TOL
from oracle import pyoracle as po

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def sift1m():
    """bench.py's workload: same generator, seeds and codebook recipe (bench.synth)."""
    base = dg.sift_like(1_000_000, 128, seed=1)
    cw = dg.roundtrip_codebook(dg.kmeans_codebook(dg.sift_like(20000, 128, seed=3), 8, 256, iters=6))
    queries = dg.sift_like(10_000, 128, seed=2)
    codes = dpq.encode(cw, base)
    return codes, cw, queries


def _check_vs_oracle(ix, payload, n, cw, queries, k, which, exact_dist=True):
    pos, ids, dist = ix.search(queries, k)
    assert np.all(np.diff(dist.astype(np.float64), axis=1) >= 0)
    for i in which:
        opos, odist, nd = po.scan(payload, n, cw, queries[i], k, want_node_dist=True)
        if exact_dist:
            assert np.array_equal(dist[i], odist), (i, dist[i], odist)
        assert_topk_equal(pos[i], dist[i], opos, odist, node_dist=nd)
    return pos, ids, dist


@pytest.mark.parametrize("n,stride", [(120_000, 8), (400_000, 16), (1_000_000, 32)])
def test_auto_path_vs_oracle_c2(sift1m, n, stride):
    """C2 and two smaller trees, every sample stride the automatic path picks; 10K queries per
    call as in bench.py (90 query groups), 30 of them compared with the oracle."""
    codes, cw, queries = sift1m
    t = dpq.tree_build(codes[:n], cw, want=("payload", "vec_id"), open_index_at=0)
    ix = t["index"]
    ix.set_codebook(cw)
    which = list(range(0, 10_000, 345))[:30]
    pos, ids, dist = _check_vs_oracle(ix, t["payload"], n, cw, queries, 10, which)
    assert ix.stat("last_coarse") == 1 and ix.stat("engine") == 2
    assert np.array_equal(ids, t["vec_id"][pos])
    # the index opened from the byte stream answers identically (host decode path)
    ix2 = dpq.DeltaTreeIndex(t["payload"], n, 8, 256, pos2id=t["vec_id"])
    ix2.set_codebook(cw)
    pos2, _, dist2 = ix2.search(queries[:2000], 10)
    assert np.array_equal(pos2, pos[:2000]) and np.array_equal(dist2, dist[:2000])
    # top-64 (deepest list the automatic narrow path serves) on a smaller batch
    _check_vs_oracle(ix, t["payload"], n, cw, queries[:300], 64, range(0, 300, 43))
    assert ix.stat("last_coarse") == 1
    ix.close()
    ix2.close()


def test_auto_path_sharded_1m(sift1m):
    """SURVEY 8e at C2 size: 4 depth-1-subtree shards (250K nodes each, automatic coarse path per
    shard) merged on the device == the oracle over the whole tree."""
    codes, cw, queries = sift1m
    n, k, R, Q = 1_000_000, 10, 4, 2000
    t = dpq.tree_build(codes, cw, want=("payload", "vec_id"))
    q = np.ascontiguousarray(queries[:Q])
    dq = dpq.DeviceBuffer(q.nbytes).upload(q)
    dk = dpq.DeviceBuffer(R * Q * k * 8)
    do = dpq.DeviceBuffer(Q * k * 8)
    shards = []
    for r in range(R):
        ix = dpq.DeltaTreeIndex(t["payload"], n, 8, 256, rank=r, n_ranks=R)
        ix.set_codebook(cw)
        ix.search_device(dq.ptr, Q, k, dk.ptr.value + r * Q * k * 8)
        ix.sync()
        assert ix.stat("last_coarse") == 1
        shards.append(ix)
    assert sum(s.stat("n_local") for s in shards) == n
    shards[0].merge_device(dk.ptr, R, Q, k, do.ptr)
    shards[0].sync()
    pos, dist = dpq.unpack_keys(do.download(np.uint64, (Q, k)))
    for i in range(0, Q, 97):
        opos, odist, nd = po.scan(t["payload"], n, cw, q[i], k, want_node_dist=True)
        assert np.array_equal(dist[i], odist)
        assert_topk_equal(pos[i], dist[i], opos, odist, node_dist=nd)
    for s in shards:
        s.close()


def test_auto_path_c3_shape():
    """C3 shape: M = 16, K = 256 (Ds = 8), top-100, 250K codes: wide coarse search (4-bit tables,
    112 queries per CTA) on the automatic path, vs the extension oracle (no reference tree format
    exists at M = 16)."""
    n, k = 250_000, 100
    base = dg.sift_like(n, 128, seed=1)
    cw = dg.roundtrip_codebook(dg.kmeans_codebook(dg.sift_like(8000, 128, seed=3), 16, 256, iters=4))
    queries = dg.sift_like(1000, 128, seed=2)
    codes = dpq.encode(cw, base)
    assert np.array_equal(codes[:3000], po.encode(cw, base[:3000]))
    t = dpq.tree_build(codes, cw, want=("payload", "vec_id"), open_index_at=0)
    ix = t["index"]
    ix.set_codebook(cw)
    _check_vs_oracle(ix, t["payload"], n, cw, queries, k, range(0, 1000, 67))
    assert ix.stat("last_coarse") == 1
    _check_vs_oracle(ix, t["payload"], n, cw, queries[:224], 10, range(0, 224, 31))
    ix.set_option("coarse", 0)  # the 15-bit wide scan alone
    _check_vs_oracle(ix, t["payload"], n, cw, queries[:96], k, range(0, 96, 19))
    ix.close()


def test_auto_path_c4_shape():
    """C4 shape: 960-d non-integer floats, M = 16 (Ds = 60): encode bit-exact vs the oracle,
    search on the automatic path, distances within 1e-5 of the oracle's double accumulation."""
    n, k = 120_000, 100
    base = dg.gist_like(n, 960, seed=1)
    cw = dg.roundtrip_codebook(dg.kmeans_codebook(dg.gist_like(6000, 960, seed=3), 16, 256, iters=3))
    queries = dg.gist_like(300, 960, seed=2)
    codes = dpq.encode(cw, base)
    assert np.array_equal(codes[:1500], po.encode(cw, base[:1500]))
    t = dpq.tree_build(codes, cw, want=("payload", "vec_id"), open_index_at=0)
    ix = t["index"]
    ix.set_codebook(cw)
    pos, ids, dist = ix.search(queries, k)
    assert ix.stat("last_coarse") == 1
    for i in range(0, 300, 23):
        opos, odist, nd = po.scan(t["payload"], n, cw, queries[i], k, want_node_dist=True)
        np.testing.assert_allclose(dist[i], odist, rtol=REL_TOL)
        assert_topk_equal(pos[i], dist[i], opos, odist, node_dist=nd)
    ix.close()


def test_auto_path_forest_of_large_parts(sift1m):
    """C5 layout with parts above the coarse threshold: 3 parts of 130K codes, each its own
    DeltaTree, merged top-10 == plain ADC over all codes (oracle tables: float entries, double sum)."""
    codes, cw, queries = sift1m
    P, n_part, k, Q = 3, 130_000, 10, 500
    q = np.ascontiguousarray(queries[:Q])
    dq = dpq.DeviceBuffer(q.nbytes).upload(q)
    dk = dpq.DeviceBuffer(P * Q * k * 8)
    do = dpq.DeviceBuffer(Q * k * 8)
    parts, id_of_pos = [], np.zeros(P * n_part, np.int64)
    for p in range(P):
        a = p * n_part
        t = dpq.tree_build(codes[a:a + n_part], cw, want=("payload", "vec_id"), open_index_at=a)
        ix = t["index"]
        ix.set_codebook(cw)
        ix.search_device(dq.ptr, Q, k, dk.ptr.value + p * Q * k * 8)
        ix.sync()
        assert ix.stat("last_coarse") == 1
        id_of_pos[a:a + n_part] = t["vec_id"].astype(np.int64) + a
        parts.append(ix)
    parts[0].merge_device(dk.ptr, P, Q, k, do.ptr)
    parts[0].sync()
    pos, dist = dpq.unpack_keys(do.download(np.uint64, (Q, k)))
    ids = id_of_pos[pos]
    sub = codes[:P * n_part]
    for i in range(0, Q, 41):
        tab = po.lut(cw, q[i]).astype(np.float64)
        d_all = tab[np.arange(8)[None, :], sub].sum(axis=1).astype(np.float32)
        order = np.lexsort((np.arange(len(sub)), d_all))[:k]
        assert np.array_equal(dist[i], d_all[order])
        assert_topk_equal(ids[i], dist[i], order, d_all[order], node_dist=d_all)
    for ix in parts:
        ix.close()


def test_massive_ties_take_the_exact_fallback():
    """thousands of nodes tying at the k-th distance (duplicate-heavy data) used to
    overflow the fallback's 2048-slot buffer and fail the whole batch.  The fallback now keeps a
    running top-k on the full (distance, position) key: ties resolve by lower position."""
    rng = np.random.default_rng(5)
    n, M, K = 150_000, 8, 256
    cw = rng.random((M, K, 4)).astype(np.float32) * 100
    distinct = rng.integers(0, K, (40, M)).astype(np.uint8)
    codes = distinct[rng.integers(0, 40, n)]          # 40 distinct codes, ~3750 copies each
    codes[::997] = rng.integers(0, K, (len(codes[::997]), M))
    t = dpq.tree_build(codes, cw, want=("payload", "vec_id"), open_index_at=0)
    ix = t["index"]
    ix.set_codebook(cw)
    queries = rng.random((64, M * 4)).astype(np.float32) * 100
    for k in (10, 100):
        pos, ids, dist = ix.search(queries, k)
        for i in range(0, 64, 9):
            opos, odist, nd = po.scan(t["payload"], n, cw, queries[i], k, want_node_dist=True)
            np.testing.assert_allclose(dist[i], odist, rtol=REL_TOL)
            assert_topk_equal(pos[i], dist[i], opos, odist, node_dist=nd)
            # own tie rule: (distance, position) ascending -> the k smallest keys exactly
            order = np.lexsort((np.arange(n), nd))[:k]
            assert np.array_equal(pos[i], order)
    ix.set_option("force_fallback", 1)   # every query through the exact fallback (Q > 4096 used to fail)
    ix.set_option("coarse", 0)
    many = np.ascontiguousarray(np.tile(queries, (80, 1)))  # 5120 queries
    pos, ids, dist = ix.search(many, 10)
    assert ix.stat("last_fallback") == len(many)
    assert np.array_equal(pos[:64], pos[64:128]) and np.array_equal(pos[:64], pos[-64:])
    ix.close()


@pytest.mark.parametrize("n", [120_000, 1_000_000])
def test_latency_mode_vs_oracle(sift1m, n):
    """Latency mode (scan1.cu: lanes = nodes, two queries per pass of the code array) is what a call
    with a handful of queries takes automatically; same answers as the oracle and as the batched
    path, for odd / even query counts and the presample-only and sampled-pass variants."""
    codes, cw, queries = sift1m
    t = dpq.tree_build(codes[:n], cw, want=("payload", "vec_id"), open_index_at=0)
    ix = t["index"]
    ix.set_codebook(cw)
    for Q, k in ((1, 10), (2, 10), (5, 32), (16, 1), (7, 10)):
        q = np.ascontiguousarray(queries[100:100 + Q])
        ix.set_option("latency", -1)
        pos, ids, dist = ix.search(q, k)
        assert ix.stat("last_latency") == 1 and ix.stat("last_coarse") == 0
        assert np.array_equal(ids, t["vec_id"][pos])
        for i in range(Q):
            opos, odist, nd = po.scan(t["payload"], n, cw, q[i], k, want_node_dist=True)
            assert np.array_equal(dist[i], odist), (Q, k, i)
            assert_topk_equal(pos[i], dist[i], opos, odist, node_dist=nd)
        ix.set_option("latency", 0)
        bpos, _, bdist = ix.search(q, k)
        assert ix.stat("last_latency") == 0
        assert np.array_equal(bpos, pos) and np.array_equal(bdist, dist)
    ix.close()


def test_latency_mode_small_and_odd_trees(sift1m):
    """Forced latency mode on trees smaller than one chunk, odd node counts, shards and duplicates."""
    codes, cw, queries = sift1m
    for n in (1, 2, 65, 2049, 5001):
        t = dpq.tree_build(codes[:n], cw, want=("payload", "vec_id"), open_index_at=0)
        ix = t["index"]
        ix.set_codebook(cw)
        ix.set_option("latency", 1)
        k = min(10, 32)
        q = np.ascontiguousarray(queries[:3])
        pos, ids, dist = ix.search(q, k)
        assert ix.stat("last_latency") == 1
        for i in range(3):
            # (the reference's size-k heap leaves its unfilled slots in FRONT when k > n; compare with the
            # oracle's per-node distances instead: the k smallest by (distance, position))
            opos, odist, nd = po.scan(t["payload"], n, cw, q[i], min(n, k), want_node_dist=True)
            m = min(n, k)
            order = np.lexsort((np.arange(n), nd))[:m]
            assert np.array_equal(dist[i][:m], nd[order]) and np.array_equal(dist[i][:m], odist[:m])
            assert_topk_equal(pos[i][:m], dist[i][:m], order, nd[order], node_dist=nd)
            assert np.all(pos[i][m:] == 0xFFFFFFFF)
        ix.close()
    # a shard that does not start at position 0, and heavy duplicates (cap = many ties)
    n = 50_000
    dup = codes[:n].copy()
    dup[1000:30000] = dup[7]
    t = dpq.tree_build(dup, cw, want=("payload", "vec_id"))
    for r in range(3):
        ix = dpq.DeltaTreeIndex(t["payload"], n, 8, 256, pos2id=t["vec_id"], rank=r, n_ranks=3)
        ix.set_codebook(cw)
        q = np.ascontiguousarray(queries[:4])
        ix.set_option("latency", 1)
        a = ix.search(q, 10)
        assert ix.stat("last_latency") == 1
        ix.set_option("latency", 0)
        b = ix.search(q, 10)
        assert np.array_equal(a[0], b[0]) and np.array_equal(a[2], b[2])
        ix.close()


def test_two_level_sample_on_a_large_shard():
    """Shards of tens of millions of nodes and more (the 10^9-code tree over 1..8 GPUs): the seeded pipeline
    samples in two levels (~64K nodes under the presample cap, then every 64th batch under that cap) before
    the full pass.  Same answers as the 15-bit sample pass (`seed=0`) and as plain ADC over the codes."""
    n, M, K, Q, k = 20_000_000, 8, 256, 300, 10
    rng = np.random.default_rng(5)
    codes = rng.integers(0, K, size=(n, M), dtype=np.uint8)
    cw = dg.roundtrip_codebook((rng.random((M, K, 16)) * 120).astype(np.float32))
    queries = (rng.random((Q, 128)) * 120).astype(np.float32)
    dcodes = dpq.DeviceBuffer(codes.nbytes).upload(codes)
    dt = dpq.DeviceTree(dcodes.ptr.value, n, M, cw)
    ix = dt.shard(0, 1)
    ix.set_codebook(cw)
    pos, ids, dist = ix.search(queries, k)
    assert ix.stat("last_coarse") == 1 and ix.stat("last_fallback") == 0
    assert ix.stat("last_sample_stride") >= 128 and ix.stat("last_refine_stride") == 64
    ix.set_option("seed", 0)
    pos0, ids0, dist0 = ix.search(queries, k)
    assert ix.stat("last_sample_stride") == 64 and ix.stat("last_refine_stride") == 0
    assert np.array_equal(dist, dist0) and np.array_equal(pos, pos0) and np.array_equal(ids, ids0)
    # plain ADC over the codes for a few queries: distance = float(double sum of the float table entries)
    lut = dpq.adc_tables(cw, queries[:3])
    for i in range(3):
        d = np.zeros(n, np.float64)
        for m in range(M):
            d += lut[i, m][codes[:, m]].astype(np.float64)
        d32 = d.astype(np.float32)
        best = np.sort(d32, kind="stable")[:k]
        assert np.array_equal(dist[i], best)
        assert np.array_equal(d32[ids[i]], dist[i])
    ix.close()
    dt.free()
    dcodes.free()
