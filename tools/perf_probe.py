"""Developer probe (NOT bench.py, not a product path): builds a 1M-code tree with the oracle
builder and times dpq search configurations.  Usage: python tools/perf_probe.py [N] [Q]"""
import os, sys, time, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import datagen as dg
import deltapq_b200 as dpq
from oracle import pyoracle as po

N = int(sys.argv[1]) if len(sys.argv) > 1 else 1000000
Q = int(sys.argv[2]) if len(sys.argv) > 2 else 10000
configs = sys.argv[3:] or ["pack=1", "pack=2"]
t = time.time()
base = dg.sift_like(N, 128, seed=1)
cw = dg.roundtrip_codebook(dg.kmeans_codebook(dg.sift_like(20000, 128, seed=3), 8, 256, iters=6))
queries = dg.sift_like(Q, 128, seed=2)
print("gen %.1fs" % (time.time() - t), flush=True)
t = time.time(); codes = dpq.encode(cw, base); print("gpu encode %.2fs" % (time.time() - t), flush=True)
t = time.time(); _, _, lay, payload = po.build_tree(codes, cw); print("oracle tree %.1fs" % (time.time() - t), flush=True)
n_bytes = len(payload)
print("n_bytes", n_bytes, "mean diffs", (n_bytes - 8 - (3 * (N - 1) + 1) // 2) / (N - 1))
ix = dpq.DeltaTreeIndex(payload, N, 8, 256, pos2id=lay["vec_id"])
ix.set_codebook(cw)
print("ops_bytes", ix.stat("ops_bytes"), "chunks", ix.stat("n_chunks"))
ref = None
for cfg in configs:
    for kv in cfg.split(","):
        k, v = kv.split("="); ix.set_option(k, int(v))
    for it in range(3):
        t = time.time(); pos, ids, dist = ix.search(queries, 10); wall = time.time() - t
        scan_us = ix.stat("last_scan_us"); tot = ix.stat("last_total_us"); lut = ix.stat("last_lut_us")
        eff = Q * n_bytes / (scan_us * 1e-6) / 1e9
        print(json.dumps(dict(cfg=cfg, it=it, wall_ms=round(wall * 1e3, 2), scan_ms=scan_us / 1e3, lut_ms=lut / 1e3,
                              total_ms=tot / 1e3, qps_scan=round(Q / (scan_us * 1e-6)), eff_GBs=round(eff, 1),
                              frac_hbm=round(eff / 6553, 3), fallback=ix.stat("last_fallback"))), flush=True)
    if ref is None:
        ref = (pos, dist)
    else:
        print("same dist as first cfg:", np.array_equal(ref[1], dist), "same pos:", np.array_equal(ref[0], pos))
for i in range(0, Q, max(1, Q // 5)):
    opos, odist = po.scan(payload, N, cw, queries[i], 10)
    print("q", i, "dist eq oracle", np.array_equal(odist, ref[1][i]), "pos eq", np.array_equal(opos, ref[0][i].astype(np.int32)))
