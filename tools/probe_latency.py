"""Developer probe: latency of small query batches (Q = 1 .. 32) on a 1M-code tree and on a
125M-code shard, latency mode (scan1.cu, lanes = nodes) against the batched path, with the
physical HBM rate of the scan kernel (8 B/node / kernel time).
Usage: python tools/probe_latency.py [N_BIG]      (N_BIG defaults to 125000000)"""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch  # noqa: E402
import bench as B  # noqa: E402
import datagen as dg  # noqa: E402
import deltapq_b200 as dpq  # noqa: E402

n_big = int(sys.argv[1]) if len(sys.argv) > 1 else 125_000_000
only_big = len(sys.argv) > 2 and sys.argv[2] == "only_big"  # profiling runs: the big shard, Q = 2, latency mode only
dev = torch.device("cuda", 0)
cw = dg.roundtrip_codebook(dg.kmeans_codebook(dg.sift_like(20000, 128, seed=3), 8, 256, iters=6))
queries = dg.sift_like(64, 128, seed=2)
for n in ((n_big,) if only_big else (1_000_000, n_big)):
    codes = torch.empty((n, 8), dtype=torch.uint8, device=dev)
    B.gen_codes_device(torch, dpq, dev, cw, n, 1000, codes)
    tree = dpq.DeviceTree(codes.data_ptr(), n, 8, cw)
    del codes
    torch.cuda.empty_cache()
    ix = tree.shard(0, 1)
    tree.free()
    ix.set_codebook(cw)
    d_key = torch.empty((64, 10), dtype=torch.int64, device=dev)
    ref = {}
    for mode in ((1,) if only_big else (1, 0)):
        ix.set_option("latency", mode)
        for Q in ((2,) if only_big else (1, 2, 4, 8, 16, 32)):
            if mode == 1 and Q > 16:
                continue
            d_q = torch.from_numpy(np.ascontiguousarray(queries[:Q])).to(dev)
            best_dev, best_wall, scan_us = None, None, None
            for it in range(6):
                torch.cuda.synchronize()
                t0 = time.perf_counter()
                ix.search_device(d_q.data_ptr(), Q, 10, d_key.data_ptr())
                ix.sync()
                wall = (time.perf_counter() - t0) * 1e6
                tot = ix.stat("last_total_us")
                if best_dev is None or tot < best_dev:
                    best_dev, scan_us = tot, ix.stat("last_scan8_us")
                best_wall = wall if best_wall is None else min(best_wall, wall)
            keys = d_key[:Q].cpu().numpy().copy()
            same = None
            if Q in ref:
                same = bool(np.array_equal(ref[Q], keys))
            else:
                ref[Q] = keys
            line = dict(n=n, mode="latency" if ix.stat("last_latency") else "batched", Q=Q, device_us=best_dev, wall_us=round(best_wall, 1),
                        scan_kernel_us=scan_us, fallback=ix.stat("last_fallback"), same_as_latency_mode=same)
            if ix.stat("last_latency") and scan_us and scan_us > 0:
                line["scan_hbm_gbs"] = round(n * 8 / (scan_us * 1e-6) / 1e9, 1)
                line["scan_hbm_frac_of_6553"] = round(n * 8 / (scan_us * 1e-6) / 1e9 / 6553.0, 3)
            print(json.dumps(line), flush=True)
    ix.close()
