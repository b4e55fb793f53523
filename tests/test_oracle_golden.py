"""CPU: the oracle restatement against fixtures produced by the UNMODIFIED reference
(tests/golden/make_golden.py).  Bit-exact for codes, edges, QNode records, stream bytes,
ADC tables and distances; positions modulo exact ties."""
import os

import numpy as np
import pytest

import datagen as dg
from helpers import assert_topk_equal, gt_fixture, assert_gt_equals_reference_text
from oracle import pyoracle as po


def _check_tree(g):
    codes, cw = g["codes"], g["cw"]
    edges, root, lay, payload = po.build_tree(codes, cw)
    assert root == int(g["root"])
    assert np.array_equal(edges, g["edges"])
    assert np.array_equal(lay["vec_id"], g["vec_id"])
    qn = po.qnodes8(codes, lay).reshape(-1, 60)
    assert np.array_equal(qn[: len(codes), 4:], g["qnode_tail"])
    assert np.array_equal(payload, g["payload"])


def test_tree_build_even_n(golden4000):
    _check_tree(golden4000)


def test_tree_build_odd_n(golden1501):
    _check_tree(golden1501)


def test_encode(golden4000, golden_m16):
    for g in (golden4000, golden_m16):
        got = po.encode(g["cw"], g["base_head"])
        assert np.array_equal(got, g["codes"][: len(got)])


def test_lut_bit_exact(golden4000):
    g = golden4000
    for i in range(len(g["ref_lut"])):
        assert np.array_equal(po.lut(g["cw"], g["queries"][i]), g["ref_lut"][i])


def _check_scan(g):
    n, k = int(g["n"]), int(g["topk"])
    for i, q in enumerate(g["queries"]):
        pos, dist, nd = po.scan(g["payload"], n, g["cw"], q, k, want_node_dist=True)
        assert np.array_equal(dist, g["ref_dist"][i])  # bit-exact distances
        ref_pos = np.where(g["ref_pos"][i] == n, n - 1, g["ref_pos"][i])  # App. C.1 quirk
        assert_topk_equal(pos, dist, ref_pos, g["ref_dist"][i], node_dist=nd)


def test_scan_even_n(golden4000):
    _check_scan(golden4000)


def test_scan_odd_n(golden1501):
    _check_scan(golden1501)


def test_decode_lossless(golden4000):
    g = golden4000
    codes, depth, parent = po.decode(g["payload"], int(g["n"]), 8)
    assert np.array_equal(codes, g["codes"][g["vec_id"]])
    assert depth[0] == 0 and parent[0] == -1
    assert np.array_equal(depth[1:], depth[parent[1:]] + 1)


def test_scan_equals_plain_adc(golden4000):
    """DeltaTree distances == plain ADC distances of the decoded codes."""
    g = golden4000
    n = int(g["n"])
    codes = g["codes"][g["vec_id"]]
    for q in g["queries"][:4]:
        lut = po.lut(g["cw"], q).astype(np.float64)
        plain = lut[np.arange(8)[None, :], codes].sum(1).astype(np.float32)
        _, _, nd = po.scan(g["payload"], n, g["cw"], q, 10, want_node_dist=True)
        assert np.array_equal(plain, nd)


def test_m16_extension_roundtrip(golden_m16):
    """M=16 has no reference tree oracle: the generalised restatement
    must at least be lossless and agree with plain ADC."""
    g = golden_m16
    codes, cw = g["codes"], g["cw"]
    edges, root, lay, payload = po.build_tree(codes, cw)
    dec, depth, parent = po.decode(payload, len(codes), 16)
    assert np.array_equal(dec, codes[lay["vec_id"]])
    q = g["queries"][0]
    lut = po.lut(cw, q).astype(np.float64)
    plain = lut[np.arange(16)[None, :], dec].sum(1).astype(np.float32)
    _, _, nd = po.scan(payload, len(codes), cw, q, 5, want_node_dist=True)
    np.testing.assert_allclose(nd, plain, rtol=1e-6)


def test_groundtruth_small():
    rng = np.random.default_rng(0)
    base = rng.integers(0, 256, (500, 32)).astype(np.float32)
    qs = rng.integers(0, 256, (5, 32)).astype(np.float32)
    ids, dist = po.groundtruth(base, qs, 7, chunk=128)
    d = ((base[None].astype(np.float64) - qs[:, None]) ** 2).sum(2)
    for i in range(5):
        order = np.argsort(d[i], kind="stable")[:7]
        assert np.array_equal(dist[i], d[i][order].astype(np.float32))


@pytest.mark.parametrize("name", ["gist_gt_n3000_d960", "gist_gt_n5000_d96"])
def test_groundtruth_oracle_vs_reference_binary(name):
    """The ground-truth port (oracle/dpq_oracle.c, float product / double sum, main.cpp:150-156)
    against what `pqtree -task groundtruth` of the unmodified reference wrote for GIST-shaped
    non-integer floats (tests/golden/make_golden_gt.py)."""
    base, queries, k, ref_ids, ref_text = gt_fixture(name)
    ids, dist = po.groundtruth(base, queries, k, chunk=1000)
    assert_gt_equals_reference_text(ids, dist, ref_ids, ref_text)
