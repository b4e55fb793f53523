"""Developer probe: time to open a DeltaTree from its on-disk byte stream (dpq_index_open: upload + GPU
decode, program_dev.cu) against the sequential host decoder (DPQ_HOST_DECODE=1, program.cpp), at N codes.
Usage: python tools/probe_open.py [N]      (default 125000000)"""
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

n = int(sys.argv[1]) if len(sys.argv) > 1 else 125_000_000
dev = torch.device("cuda", 0)
cw = dg.roundtrip_codebook(dg.kmeans_codebook(dg.sift_like(20000, 128, seed=3), 8, 256, iters=6))
codes = torch.empty((n, 8), dtype=torch.uint8, device=dev)
B.gen_codes_device(torch, dpq, dev, cw, n, 1000, codes)
tree = dpq.DeviceTree(codes.data_ptr(), n, 8, cw)
del codes
torch.cuda.empty_cache()
payload = tree.fetch("payload", np.uint8)
vec_id = tree.fetch("vec_id", np.uint32)
ref = tree.shard(0, 1)
tree.free()
ref.set_codebook(cw)
q = dg.sift_like(64, 128, seed=2)
want = ref.search(q, 10)
ref.close()
out = {"n_codes": n, "stream_bytes": int(len(payload))}
for name, env in (("gpu_decode", None), ("host_decode", "1")):
    if env:
        os.environ["DPQ_HOST_DECODE"] = env
    else:
        os.environ.pop("DPQ_HOST_DECODE", None)
    best = None
    for _ in range(2 if env is None else 1):
        t = time.perf_counter()
        ix = dpq.DeltaTreeIndex(payload, n, 8, 256, pos2id=vec_id)
        dt = time.perf_counter() - t
        best = dt if best is None else min(best, dt)
        ix.set_codebook(cw)
        got = ix.search(q, 10)
        assert all(np.array_equal(a, b) for a, b in zip(got, want)), name
        ix.close()
    out[name + "_open_s"] = round(best, 3)
print(json.dumps(out))
