"""Developer probe: wide-shape (M=16) coarse search knobs on a 1M-code tree.
Usage: python tools/probe_wide_coarse.py [N] [Q]"""
import os, sys, time, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import datagen as dg
import deltapq_b200 as dpq
N = int(sys.argv[1]) if len(sys.argv) > 1 else 1000000
Q = int(sys.argv[2]) if len(sys.argv) > 2 else 2000
M, K = 16, 256
base = dg.sift_like(N, 128, seed=1)
cw = dg.roundtrip_codebook(dg.kmeans_codebook(dg.sift_like(20000, 128, seed=3), M, K, iters=6))
queries = dg.sift_like(Q, 128, seed=2)
codes = dpq.encode(cw, base)
t = dpq.tree_build(codes, cw, want=("payload", "vec_id"), open_index_at=0)
ix = t["index"]; ix.set_codebook(cw)
ref = {}
for topk in (10, 100):
    ix.set_option("coarse", 0)
    pos, ids, dist = ix.search(queries, topk); pos, ids, dist = ix.search(queries, topk)
    ref[topk] = dist
    print(json.dumps(dict(topk=topk, mode="15-bit", total_ms=ix.stat("last_total_us") / 1e3, scan_ms=ix.stat("last_scan_us") / 1e3)), flush=True)
    for levels in (80, 119):
        for bcap in (512, 2048, 8192):
            for sample in (0, 4):
                ix.set_option("coarse", 1); ix.set_option("levels8", levels); ix.set_option("bcap8", bcap); ix.set_option("sample", sample)
                pos, ids, d2 = ix.search(queries, topk); pos, ids, d2 = ix.search(queries, topk)
                print(json.dumps(dict(topk=topk, levels=levels, bcap8=bcap, sample=sample, total_ms=ix.stat("last_total_us") / 1e3,
                                      scan_ms=ix.stat("last_scan_us") / 1e3, scan8_ms=ix.stat("last_scan8_us") / 1e3,
                                      fallback=ix.stat("last_fallback"), survivors_per_query=round(ix.stat("cand8_total") / Q),
                                      same=bool(np.array_equal(d2, ref[topk])))), flush=True)
