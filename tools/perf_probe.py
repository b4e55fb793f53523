"""Developer probe (NOT bench.py, not a product path): builds the bench tree once with libdpq
and times search configurations.  Usage: python tools/perf_probe.py N Q cfg [cfg ...]
cfg = comma separated option=value pairs for dpq_index_set_option."""
import os, sys, time, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import datagen as dg
import deltapq_b200 as dpq

N = int(sys.argv[1]) if len(sys.argv) > 1 else 1000000
Q = int(sys.argv[2]) if len(sys.argv) > 2 else 10000
configs = sys.argv[3:] or ["epoch=32"]
base = dg.sift_like(N, 128, seed=1)
cw = dg.roundtrip_codebook(dg.kmeans_codebook(dg.sift_like(20000, 128, seed=3), 8, 256, iters=6))
queries = dg.sift_like(Q, 128, seed=2)
codes = dpq.encode(cw, base)
t = time.time(); tree = dpq.tree_build(codes, cw); print("tree build %.2fs" % (time.time() - t), flush=True)
payload = tree["payload"]
n_bytes = len(payload)
ix = dpq.DeltaTreeIndex(payload, N, 8, 256, pos2id=tree["vec_id"])
ix.set_codebook(cw)
print("engine", ix.stat("engine"), "prog_bytes", ix.stat("ops_bytes"), "chunks", ix.stat("n_chunks"),
      "delta_nodes", ix.stat("v2_delta_nodes"), flush=True)
ref = None
for cfg in configs:
    for kv in cfg.split(","):
        k, v = kv.split("="); ix.set_option(k, int(v))
    best = None
    for it in range(4):
        pos, ids, dist = ix.search(queries, 10)
        scan_us = ix.stat("last_scan_us"); tot = ix.stat("last_total_us"); lut = ix.stat("last_lut_us")
        if best is None or scan_us < best[0]:
            best = (scan_us, tot, lut)
    scan_us, tot, lut = best
    eff = Q * n_bytes / (scan_us * 1e-6) / 1e9
    print(json.dumps(dict(cfg=cfg, scan_ms=scan_us / 1e3, lut_ms=lut / 1e3, total_ms=tot / 1e3,
                          qps_total=round(Q / (tot * 1e-6)), eff_GBs=round(eff, 1),
                          fallback=ix.stat("last_fallback"), cand8_per_query=round(ix.stat("cand8_total") / max(ix.stat("last_device_queries"), 1), 1))), flush=True)
    if ref is None:
        ref = (pos, dist)
    else:
        print("  same as first cfg:", np.array_equal(ref[1], dist) and np.array_equal(ref[0], pos), flush=True)
