"""The C++ multi-GPU path (dpq_multi_*: one process, one subtree shard per GPU, NCCL all-gather bound
at run time, device merge -- what `deltapq -task query -gpus N` runs) on the tree and queries bench.py
saved, checked against the single-GPU search of the same queries.  bench.py runs this in a
SUBPROCESS with a timeout so that nothing here can stall the timed bench.
Usage: python tools/multi_probe.py DIR N_GPUS [STEPS]   -> one JSON line"""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import deltapq_b200 as dpq  # noqa: E402


def main():
    d, n_gpus = sys.argv[1], int(sys.argv[2])
    steps = int(sys.argv[3]) if len(sys.argv) > 3 else 5
    z = np.load(os.path.join(d, "multi_probe.npz"))
    cw, queries, k, n = z["cw"], z["queries"], int(z["topk"]), int(z["n_codes"])
    tree, qnode = os.path.join(d, "tree.bin"), os.path.join(d, "qnodes.bin")
    t0 = time.perf_counter()
    mx = dpq.MultiIndex(tree, 8, 256, n_gpus, qnode_path=qnode)
    mx.set_codebook(cw)
    open_s = time.perf_counter() - t0
    pos, ids, dist = mx.search(queries, k)  # warm-up (scratch, NCCL channels)
    t0 = time.perf_counter()
    for _ in range(steps):
        pos, ids, dist = mx.search(queries, k)
    dt = (time.perf_counter() - t0) / steps
    shard_nodes = [mx.stat(r, "n_local") for r in range(n_gpus)]
    open_parts = {"shards_open_s": round(mx.stat(0, "multi_open_us") / 1e6, 2), "nccl_comm_init_s": round(mx.stat(0, "multi_nccl_init_us") / 1e6, 2)}
    mx.close()
    dpq.set_device(0)
    one = dpq.DeltaTreeIndex.from_file(tree, 8, 256, qnode_path=qnode)
    one.set_codebook(cw)
    opos, oids, odist = one.search(queries, k)
    one.close()
    same = bool(np.array_equal(pos, opos) and np.array_equal(dist, odist) and np.array_equal(ids, oids))
    print(json.dumps({"n_gpus": n_gpus, "queries": int(len(queries)), "topk": k, "n_codes": n, "open_s": round(open_s, 2), "open_breakdown": open_parts,
                      "ms_per_call_host_buffers": dt * 1e3, "queries_per_s": len(queries) / dt,
                      "shard_nodes": shard_nodes, "equals_single_gpu": same,
                      "path": "dpq_multi_open_file / dpq_multi_search (C++ host, NCCL via dlopen, no torch)"}))
    return 0 if same else 1


if __name__ == "__main__":
    sys.exit(main())
