"""Developer probe: BASELINE configs[2] shape (SIFT1M-shaped, M=16 K=256, top-100) or, with a
third argument "gist", configs[3] shape (GIST-shaped 960-d floats in [0,1), M=16 K=256).
Usage: python tools/perf_probe_m16.py [N=1000000] [Q=2000] [gist]"""
import os, sys, time, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import datagen as dg
import deltapq_b200 as dpq
from oracle import pyoracle as po
N = int(sys.argv[1]) if len(sys.argv) > 1 else 1000000
Q = int(sys.argv[2]) if len(sys.argv) > 2 else 2000
M = 16; K = 256; topk = 100
gist = len(sys.argv) > 3 and sys.argv[3] == "gist"
gen, D = (dg.gist_like, 960) if gist else (dg.sift_like, 128)
base = gen(N, D, seed=1)
cw = dg.roundtrip_codebook(dg.kmeans_codebook(gen(8000 if gist else 20000, D, seed=3), M, K, iters=4 if gist else 6))
queries = gen(Q, D, seed=2)
if gist:
    assert np.array_equal(dpq.encode(cw, base[:3000]), po.encode(cw, base[:3000])), "GIST-shaped encode not bit-exact"
t = time.time(); codes = dpq.encode(cw, base); print("encode %.2fs" % (time.time() - t), flush=True)
t = time.time(); tree = dpq.tree_build(codes, cw); print("tree build %.2fs" % (time.time() - t), "n_bytes", len(tree["payload"]), "n_diffs", tree["n_diffs"], flush=True)
payload = tree["payload"]
ix = dpq.DeltaTreeIndex(payload, N, M, K, pos2id=tree["vec_id"]); ix.set_codebook(cw)
print("engine", ix.stat("engine"), "prog_bytes", ix.stat("ops_bytes"), flush=True)
for it in range(3):
    pos, ids, dist = ix.search(queries, topk)
    print(json.dumps(dict(scan_ms=ix.stat("last_scan_us") / 1e3, lut_ms=ix.stat("last_lut_us") / 1e3, total_ms=ix.stat("last_total_us") / 1e3,
                          qps=round(Q / (ix.stat("last_total_us") * 1e-6)), fallback=ix.stat("last_fallback"),
                          coarse=ix.stat("last_coarse"), scan8_ms=ix.stat("last_scan8_us") / 1e3,
                          cand8_per_query=round(ix.stat("cand8_total") / max(ix.stat("last_device_queries"), 1), 1))), flush=True)
for cfg in sys.argv[4:] if gist else sys.argv[3:]:
    for kv in cfg.split(","):
        kk, v = kv.split("="); ix.set_option(kk, int(v))
    for kk in (100, 10):
        best = None
        for it in range(3):
            p2, i2, d2 = ix.search(queries, kk)
            tot = ix.stat("last_total_us")
            if best is None or tot < best[0]:
                best = (tot, ix.stat("last_scan_us"), ix.stat("last_lut_us"), ix.stat("last_scan8_us"))
        print(json.dumps(dict(cfg=cfg, topk=kk, total_ms=best[0] / 1e3, scan_ms=best[1] / 1e3, lut_ms=best[2] / 1e3,
                              scan8_ms=best[3] / 1e3, qps=round(Q / (best[0] * 1e-6)), coarse=ix.stat("last_coarse"),
                              fallback=ix.stat("last_fallback"),
                              cand8_per_query=round(ix.stat("cand8_total") / max(ix.stat("last_device_queries"), 1), 1) if ix.stat("last_coarse") else None,
                              same_as_default=bool(kk != 100 or (np.array_equal(p2, pos) and np.array_equal(d2, dist))))), flush=True)
for i in (0, Q // 2):
    opos, odist = po.scan(payload, N, cw, queries[i], topk)
    print("q", i, "dist allclose", bool(np.allclose(odist, dist[i], rtol=1e-5)), "exact", bool(np.array_equal(odist, dist[i])))
