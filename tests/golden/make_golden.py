"""Generates tests/golden/*.npz by running the UNMODIFIED reference (oracle/_ref, built from
/root/reference by oracle/Makefile) on small deterministic inputs.  Run in the build
container only (the GPU box has no /root/reference):

    python tests/golden/make_golden.py

Each fixture holds the inputs (codebook, vectors, queries) and what the reference produced:
codes (pqtree -task encode), Edges / QNode / compressed-tree files (deltapq_canon -task
approx_tree; canonical stable-sort build) and the in-memory scan's
top-k + ADC tables (DCAT.h:3731 through oracle/ref_harness.cpp).
"""
import os
import shutil
import subprocess
import sys
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import datagen as dg  # noqa: E402
from oracle import pyoracle as po  # noqa: E402


def run(cmd):
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError(f"{cmd} failed rc={r.returncode}\n{r.stdout[-2000:]}\n{r.stderr[-2000:]}")


def make(name, n, n_query, M, K, d, seed, topk):
    po.build(ref=True)
    tmp = tempfile.mkdtemp(prefix="dpq_golden_")
    try:
        base, queries, cw = dg.make_dataset(tmp, n, n_query, M=M, K=K, d=d, seed=seed)
        run([po.REF_DIR + "/pqtree", "-dataset", tmp, "-task", "encode", "-m", str(M), "-k", str(K),
             "-N", str(n), "-ext", "fvecs"])
        codes = dg.read_codes(f"{tmp}/codes.bin.plain.M{M}K{K}N{n}", M)
        out = dict(M=M, K=K, n=n, cw=cw, queries=queries, codes=codes, topk=topk,
                   base_head=base[:256].copy())
        if M == 8:
            run([po.REF_DIR + "/deltapq_canon", "-dataset", tmp, "-task", "approx_tree", "-m", str(M),
                 "-k", str(K), "-h", "1", "-diff", str(M), "-N", str(n)])
            e = np.fromfile(f"{tmp}/M{M}K{K}H1_Approx_Edges_N{n}", dtype=np.uint32)
            qn = np.fromfile(f"{tmp}/M{M}K{K}_Approx_TreeNodesDFS_N{n}", dtype=np.uint8).reshape(n + 1, 60)
            nc, payload = dg.read_dtc(f"{tmp}/M{M}K{K}_Approx_compressed_codes_opt_N{n}")
            assert nc == n
            pos, dist, _, lut = po.ref_scan(payload, n, cw, queries, topk, want_lut=True)
            out.update(root=e[0], edges=e[1:].reshape(-1, 2), payload=payload,
                       vec_id=qn[:n, 0:4].copy().view(np.uint32).ravel(),
                       qnode_tail=qn[:n, 4:].copy(),  # everything but vec_id, for byte-exact checks
                       ref_pos=pos, ref_dist=dist, ref_lut=lut[:4].copy())
        gt_ids, gt_dist = None, None
        np.savez_compressed(os.path.join(ROOT, "tests", "golden", name + ".npz"), **out)
        print(name, "ok", {k: getattr(v, "shape", v) for k, v in out.items()})
    finally:
        shutil.rmtree(tmp, ignore_errors=True)


if __name__ == "__main__":
    make("sift_n4000_m8", 4000, 24, 8, 256, 128, seed=11, topk=10)   # even N: trailing node
    make("sift_n1501_m8", 1501, 8, 8, 256, 128, seed=12, topk=5)    # odd N
    make("sift_n600_m16", 600, 4, 16, 256, 128, seed=13, topk=5)    # M=16: codes only (no ref tree)
