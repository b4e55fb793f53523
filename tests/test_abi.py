"""CPU: the C-ABI library loads, exports every symbol include/dpq.h declares, fails loudly
without a GPU, and its host-side tree compiler is correct (interpreted on the CPU)."""
import ctypes as C
import os
import re

import numpy as np
import pytest

import deltapq_b200 as dpq
from helpers import interpret_any, interpret_program, interpret_program2
from oracle import pyoracle as po

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_exports_every_declared_symbol():
    hdr = open(os.path.join(ROOT, "include", "dpq.h")).read()
    declared = set(re.findall(r"\b(dpq_[a-z0-9_]+)\s*\(", hdr))
    assert declared, "no declarations parsed"
    L = dpq.lib()
    for name in sorted(declared):
        assert hasattr(L, name), f"libdpq.so does not export {name}"
    assert declared == set(dpq.SYMBOLS)
    assert L.dpq_version() >= 100


def test_fails_loudly_without_gpu(golden1501):
    if dpq.device_count() > 0:
        pytest.skip("a GPU is present")
    g = golden1501
    with pytest.raises(dpq.DpqError, match="no CUDA device"):
        dpq.DeltaTreeIndex(g["payload"], int(g["n"]), 8, 256)
    with pytest.raises(dpq.DpqError, match="no CUDA device"):
        dpq.encode(g["cw"], g["base_head"])
    with pytest.raises(dpq.DpqError, match="no CUDA device"):
        dpq.adc_tables(g["cw"], g["queries"])
    with pytest.raises(dpq.DpqError, match="no CUDA device"):
        dpq.DeltaTreeIndex(g["payload"], int(g["n"]), 8, 256, first_pos=1000)          # dpq_index_open_part
    with pytest.raises(dpq.DpqError, match="no CUDA device"):
        dpq.encode_u8(g["cw"], np.zeros(132 * 4, np.uint8), 4, 128, 132, 4)
    with pytest.raises(dpq.DpqError, match="no CUDA device"):
        dpq.groundtruth(g["base_head"], g["queries"], 3)
    with pytest.raises(dpq.DpqError, match="no CUDA device"):
        dpq.tree_build(g["codes"], g["cw"])
    # the host half of the build needs no GPU; opening its tree as an index does
    L = dpq.lib()
    codes = np.ascontiguousarray(g["codes"], np.uint8)
    cw = np.ascontiguousarray(g["cw"], np.float32)
    edges = np.ascontiguousarray(g["edges"], np.uint32)
    t = C.c_void_p()
    assert L.dpq_tree_from_edges(codes.ctypes.data_as(C.c_void_p), len(codes), 8, 256, cw.ctypes.data_as(C.c_void_p),
                                 cw.shape[2], edges.ctypes.data_as(C.c_void_p), int(g["root"]), C.byref(t)) == 0
    ix = C.c_void_p()
    assert L.dpq_index_open_tree(t, 0, C.byref(ix)) == -2 and b"no CUDA device" in L.dpq_last_error()
    L.dpq_tree_free(t)


def test_rejects_malformed_stream(golden1501):
    g = golden1501
    n = int(g["n"])
    with pytest.raises(dpq.DpqError, match="trunc|mismatch|missing"):
        dpq.compile_program(g["payload"][:-3], n, 8, 256)
    bad = g["payload"].copy()
    bad[8] = 0x77  # depth jump 0 -> 7
    with pytest.raises(dpq.DpqError, match="depth"):
        dpq.compile_program(bad, n, 8, 256)
    with pytest.raises(dpq.DpqError):
        dpq.compile_program(g["payload"], n, 33, 256)


@pytest.mark.parametrize("engine", [0, 1])
@pytest.mark.parametrize("chunk_nodes", [4, 37, 256, 100000])
def test_program_reproduces_every_node_distance(golden4000, chunk_nodes, engine):
    """The compiled op program, interpreted on the CPU with an integer table, reproduces the
    oracle's per-node sums exactly, for any chunking."""
    g = golden4000
    n = int(g["n"])
    prog = dpq.compile_program(g["payload"], n, 8, 256, chunk_nodes=chunk_nodes, engine=engine)
    assert prog["v2"] == (engine == 0)
    assert prog["n_local"] == n and prog["base_pos"] == 0
    assert prog["n_bytes"] == len(g["payload"])
    assert np.array_equal(prog["codes"], g["codes"][g["vec_id"]])
    rng = np.random.default_rng(3)
    table = rng.integers(0, 1 << 20, 8 * 256).astype(np.int64)
    pos, d = interpret_any(prog, table)
    assert np.array_equal(np.sort(pos), np.arange(n))
    want = table.reshape(8, 256)[np.arange(8)[None, :], prog["codes"]].sum(1)
    assert np.array_equal(d[np.argsort(pos)], want)


def test_v2_program_is_the_code_array(golden4000):
    """The code-array program is the decoded codes by DFS position (what the oracle's decoder
    gives), padded to the scan's word stride: 8 bytes per node at M = 8."""
    g = golden4000
    n = int(g["n"])
    prog = dpq.compile_program(g["payload"], n, 8, 256)
    codes, depth, parent = po.decode(g["payload"], n, 8)
    assert prog["cstride"] == 8 and prog["codes_padded"].shape == (n, 8)
    assert np.array_equal(prog["codes"], codes)
    assert prog["n_diffs"] == int((codes[1:] != codes[parent[1:]]).sum())


def test_deep_streams(golden4000):
    """the reader's depth limit follows the FORMAT (nibble & 7 when M <= 8, 4 bits in the
    M > 8 extension), not the table size: an M = 16, K = 128 tree of depth 9 opens; an M <= 8 tree
    deeper than 7 cannot be written (the reference reader would alias its depths)."""
    rng = np.random.default_rng(11)
    # a chain: node i differs from node i-1 in one subspace -> depth grows by one per node
    for M, K, depth_ok in ((16, 128, 12), (8, 256, 7)):
        n = depth_ok + 1
        codes = np.zeros((n, M), np.uint8)
        for i in range(1, n):
            codes[i] = codes[i - 1]
            codes[i, i % M] = 1 + i % (K - 1)
        edges = np.array([[i - 1, i] for i in range(1, n)], np.uint32)
        cw = rng.random((M, K, 2)).astype(np.float32)
        t = dpq.tree_from_edges(codes, cw, edges, 0)
        assert int(t["depth"].max()) == depth_ok
        prog = dpq.compile_program(t["payload"], n, M, K)
        assert prog["v2"] and np.array_equal(prog["codes"], codes[t["vec_id"]])
    # one level deeper than the M <= 8 format can hold: refused by the writer
    n = 9
    codes = np.zeros((n, 8), np.uint8)
    for i in range(1, n):
        codes[i] = codes[i - 1]
        codes[i, i % 8] = i
    edges = np.array([[i - 1, i] for i in range(1, n)], np.uint32)
    with pytest.raises(dpq.DpqError, match="depth"):
        dpq.tree_from_edges(codes, rng.random((8, 256, 2)).astype(np.float32), edges, 0)


@pytest.mark.parametrize("engine", [0, 1])
@pytest.mark.parametrize("n_ranks", [2, 3, 8])
def test_shards_partition_the_tree(golden4000, n_ranks, engine):
    g = golden4000
    n = int(g["n"])
    codes_dfs = g["codes"][g["vec_id"]]
    table = np.arange(8 * 256, dtype=np.int64) * 7 + 1
    want = table.reshape(8, 256)[np.arange(8)[None, :], codes_dfs].sum(1)
    seen = np.zeros(n, np.int32)
    total_bytes = 0
    next_base = 0
    for r in range(n_ranks):
        prog = dpq.compile_program(g["payload"], n, 8, 256, rank=r, n_ranks=n_ranks, chunk_nodes=64, engine=engine)
        if prog["n_local"] == 0:
            continue
        assert prog["base_pos"] == next_base  # contiguous position ranges
        next_base = prog["base_pos"] + prog["n_local"]
        assert np.array_equal(prog["codes"], codes_dfs[prog["base_pos"]:next_base])
        pos, d = interpret_any(prog, table)
        seen[pos] += 1
        assert np.array_equal(d, want[pos])
        total_bytes += prog["n_bytes"]
    assert next_base == n and np.all(seen == 1)
    # every shard repeats the root code and may round its depth nibbles up
    assert len(g["payload"]) <= total_bytes <= len(g["payload"]) + n_ranks * 9


def test_m16_program(golden_m16):
    g = golden_m16
    codes, cw = g["codes"], g["cw"]
    _, _, lay, payload = po.build_tree(codes, cw)
    table = np.random.default_rng(5).integers(0, 1 << 18, 16 * 256).astype(np.int64)
    want = table.reshape(16, 256)[np.arange(16)[None, :], codes[lay["vec_id"]]].sum(1)
    prog = dpq.compile_program(payload, len(codes), 16, 256, chunk_nodes=50, engine=1)
    assert prog["rb"] == 12 and prog["levels"] == 16 and not prog["v2"]
    pos, d = interpret_program(prog, table)
    assert np.array_equal(d[np.argsort(pos)], want)
    # wide code-array program: 16 code bytes per node, 3 x 16-byte table rows in the 15-bit scan
    prog = dpq.compile_program(payload, len(codes), 16, 256)
    assert prog["v2"] and prog["v2_nf"] == 16 and prog["v2_lpg"] == 3 and prog["codes_padded"].shape[1] == 16
    pos, d = interpret_program2(prog, table)
    assert np.array_equal(d[np.argsort(pos)], want)


@pytest.mark.parametrize("M,K", [(4, 256), (8, 100), (3, 16), (12, 256), (16, 200), (9, 64)])
def test_v2_program_other_shapes(M, K):
    """M < 8 / M < 16: pad code bytes are 0 and read the all-zero rows of the unused subspaces."""
    rng = np.random.default_rng(M * 1000 + K)
    n = 700
    codes = rng.integers(0, min(K, 5), (n, M)).astype(np.uint8)
    cw = rng.random((M, K, 2)).astype(np.float32)
    _, _, lay, payload = po.build_tree(codes, cw)
    prog = dpq.compile_program(payload, n, M, K)
    assert prog["v2"] and prog["v2_nf"] == (8 if M <= 8 else 16)
    table = rng.integers(0, 1 << 16, M * K).astype(np.int64)
    pos, d = interpret_program2(prog, table)
    want = table.reshape(M, K)[np.arange(M)[None, :], codes[lay["vec_id"]]].sum(1)
    assert np.array_equal(np.sort(pos), np.arange(n)) and np.array_equal(d[np.argsort(pos)], want)


def test_single_node_tree():
    payload = np.arange(8, dtype=np.uint8)
    prog = dpq.compile_program(payload, 1, 8, 256)
    pos, d = interpret_any(prog, np.arange(2048, dtype=np.int64))
    assert list(pos) == [0] and d[0] == sum(m * 256 + m for m in range(8))
