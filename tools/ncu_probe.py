"""Developer probe for ncu: small fixed workload (N codes, Q queries), tree built by the
oracle builder, cached in gpurun_out/.  Usage: python tools/ncu_probe.py N Q [opt=val,...]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import datagen as dg
import deltapq_b200 as dpq
from oracle import pyoracle as po
N, Q = int(sys.argv[1]), int(sys.argv[2])
cache = os.path.join(ROOT, "gpurun_out", f"probe_tree_{N}.npz")
if os.path.exists(cache):
    z = np.load(cache); payload, cw = z["payload"], z["cw"]
else:
    base = dg.sift_like(N, 128, seed=1)
    cw = dg.roundtrip_codebook(dg.kmeans_codebook(dg.sift_like(20000, 128, seed=3), 8, 256, iters=6))
    codes = dpq.encode(cw, base)
    _, _, lay, payload = po.build_tree(codes, cw)
    os.makedirs(os.path.dirname(cache), exist_ok=True)
    np.savez(cache, payload=payload, cw=cw)
queries = dg.sift_like(Q, 128, seed=2)
ix = dpq.DeltaTreeIndex(payload, N, 8, 256)
ix.set_codebook(cw)
for kv in (sys.argv[3].split(",") if len(sys.argv) > 3 else []):
    k, v = kv.split("="); ix.set_option(k, int(v))
for it in range(2):
    ix.search(queries, 10)
    print("scan_us", ix.stat("last_scan_us"), "Q", Q, "N", N, flush=True)
