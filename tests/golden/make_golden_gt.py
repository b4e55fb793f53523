"""Generates tests/golden/gist_gt_*.npz: the reference's OWN ground truth (`pqtree -task
groundtruth`, main.cpp:569-669 -> batch_partial_topk_queries main.cpp:138-166, text writer
pqbase.cpp:294-312) on GIST-shaped non-integer floats, where the float-product / double-sum order
matters.  Run in the build container only (needs oracle/_ref):

    python tests/golden/make_golden_gt.py

The fixture holds the generator parameters (the inputs are re-created from tests/datagen.py with
the same seeds; a checksum guards against generator drift), the queries, and what the reference
wrote: ids and the 6-significant-digit distances of its text file."""
import hashlib
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


def make(name, n, d, n_query, topk, seed, kind):
    po.build(ref=True)
    gen = dg.gist_like if kind == "gist" else dg.sift_like
    base = gen(n, d, seed=seed)
    queries = gen(n_query, d, seed=seed + 1)
    tmp = tempfile.mkdtemp(prefix="dpq_golden_gt_")
    try:
        dg.write_vecs(f"{tmp}/base.fvecs", base)
        dg.write_vecs(f"{tmp}/query.fvecs", queries)
        os.makedirs(f"{tmp}/groundtruth")
        r = subprocess.run([po.REF_DIR + "/pqtree", "-dataset", tmp, "-task", "groundtruth", "-ext", "fvecs",
                            "-N", str(n), "-query_size", str(n_query), "-topk", str(topk)], capture_output=True, text=True)
        assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
        lines = open(f"{tmp}/groundtruth/N{n}Top{topk}.txt").read().splitlines()
        assert lines[0].strip().rstrip(",") == f"{n_query},{topk}", lines[0]
        ids = np.zeros((n_query, topk), np.uint32)
        dist = np.zeros((n_query, topk), np.float64)
        for i in range(n_query):
            f = [x for x in lines[1 + i].split(",") if x.strip()]
            ids[i] = [int(v) for v in f[0::2]]
            dist[i] = [float(v) for v in f[1::2]]
        np.savez_compressed(os.path.join(ROOT, "tests", "golden", name + ".npz"), n=n, d=d, n_query=n_query, topk=topk,
                            seed=seed, kind=kind, queries=queries, ref_ids=ids, ref_dist_text=dist,
                            base_sha1=hashlib.sha1(base.tobytes()).hexdigest())
        print(name, "ok", ids.shape, dist[0, :3])
    finally:
        shutil.rmtree(tmp, ignore_errors=True)


if __name__ == "__main__":
    make("gist_gt_n3000_d960", 3000, 960, 12, 10, seed=31, kind="gist")
    make("gist_gt_n5000_d96", 5000, 96, 16, 20, seed=33, kind="gist")
