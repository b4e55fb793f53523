"""CPU, build container only: the oracle against the live reference build (oracle/_ref).
Skipped where oracle/_ref is absent."""
import os
import shutil
import subprocess
import tempfile

import numpy as np
import pytest

import datagen as dg
from oracle import pyoracle as po

pytestmark = pytest.mark.skipif(not po.have_ref() or not os.path.exists(po.REF_DIR + "/deltapq_canon"),
                                reason="oracle/_ref not built")


@pytest.mark.parametrize("n,seed", [(3000, 5), (2501, 6)])
def test_pipeline_matches_reference(n, seed):
    tmp = tempfile.mkdtemp(prefix="dpq_ref_")
    try:
        base, queries, cw = dg.make_dataset(tmp, n, 6, seed=seed)
        subprocess.run([po.REF_DIR + "/pqtree", "-dataset", tmp, "-task", "encode", "-m", "8", "-k", "256",
                        "-N", str(n), "-ext", "fvecs"], capture_output=True, check=True)
        codes = po.encode(cw, base)
        assert np.array_equal(codes, dg.read_codes(f"{tmp}/codes.bin.plain.M8K256N{n}", 8))
        subprocess.run([po.REF_DIR + "/deltapq_canon", "-dataset", tmp, "-task", "approx_tree", "-m", "8",
                        "-k", "256", "-h", "1", "-diff", "8", "-N", str(n)], capture_output=True, check=True)
        edges, root, lay, payload = po.build_tree(codes, cw)
        e = np.fromfile(f"{tmp}/M8K256H1_Approx_Edges_N{n}", dtype=np.uint32)
        assert e[0] == root and np.array_equal(e[1:].reshape(-1, 2), edges)
        assert np.array_equal(np.fromfile(f"{tmp}/M8K256_Approx_TreeNodesDFS_N{n}", dtype=np.uint8),
                              po.qnodes8(codes, lay))
        nc, ref_payload = dg.read_dtc(f"{tmp}/M8K256_Approx_compressed_codes_opt_N{n}")
        assert nc == n and np.array_equal(ref_payload, payload)
        rpos, rdist, _, rlut = po.ref_scan(payload, n, cw, queries, 10, want_lut=True)
        for i, q in enumerate(queries):
            assert np.array_equal(po.lut(cw, q), rlut[i])
            _, dist = po.scan(payload, n, cw, q, 10)
            assert np.array_equal(dist, rdist[i])
    finally:
        shutil.rmtree(tmp, ignore_errors=True)


def test_ref_encode_harness():
    rng = np.random.default_rng(1)
    cw = dg.roundtrip_codebook(rng.random((8, 256, 16)).astype(np.float32) * 100)
    x = (rng.random((300, 128)) * 100).astype(np.float32)
    got, _ = po.ref_encode(cw, x)
    assert np.array_equal(got, po.encode(cw, x))
