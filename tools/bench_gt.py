"""Ground truth (pqtree -task groundtruth, pmain:569-669) at BASELINE configs[1] shape with the
base vectors already on the device: tensor-core filter path vs the plain exact kernels.
Usage: python tools/bench_gt.py [N] [Q] [D] [topk]      (prints one JSON line)"""
import ctypes as C, json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import torch
import datagen as dg
import deltapq_b200 as dpq

N = int(sys.argv[1]) if len(sys.argv) > 1 else 1000000
Q = int(sys.argv[2]) if len(sys.argv) > 2 else 10000
D = int(sys.argv[3]) if len(sys.argv) > 3 else 128
K = int(sys.argv[4]) if len(sys.argv) > 4 else 10
gen = dg.gist_like if D == 960 else dg.sift_like
base = torch.from_numpy(gen(N, D, seed=1)).cuda()
queries = np.ascontiguousarray(gen(Q, D, seed=2))
L = dpq.lib()


def run(tc, nq):
    os.environ["DPQ_GT_TC"] = str(int(tc))
    st = C.c_void_p()
    q = np.ascontiguousarray(queries[:nq])
    dpq._check(L.dpq_groundtruth_begin(dpq._ptr(q), nq, D, K, C.byref(st)))
    torch.cuda.synchronize()
    t = time.perf_counter()
    dpq._check(L.dpq_groundtruth_chunk(st, C.c_void_p(base.data_ptr()), N, 0))
    dt = time.perf_counter() - t
    stats = {k: int(L.dpq_groundtruth_stat(st, k.encode())) for k in ("tc", "tc_vectors", "tc_flagged", "tc_filter_us", "tc_rescore_us")}
    ids = np.empty((nq, K), np.uint32); dist = np.empty((nq, K), np.float32)
    dpq._check(L.dpq_groundtruth_finish(st, dpq._ptr(ids), dpq._ptr(dist)))
    return dt, ids, dist, stats


run(2, min(Q, 256))  # warm-up: context, allocations
t_tc, ids, dist, stats = run(2, Q)
t_sync, sid, sdist, sstats = run(1, Q)
nq_plain = min(Q, 1000)  # the plain path on a bounded sample of the queries
t_plain, pid, pdist, _ = run(0, nq_plain)
same = bool(np.array_equal(ids[:nq_plain], pid) and np.array_equal(dist[:nq_plain], pdist))
flops = 2.0 * N * Q * D
print(json.dumps(dict(N=N, Q=Q, D=D, topk=K, tc_s=round(t_tc, 4), tc_queries_per_s=round(Q / t_tc),
                      tc_effective_tflops=round(flops / t_tc / 1e12, 2), tc_mma_tflops=round(3 * flops / t_tc / 1e12, 2),
                      filter_kernel_mma_tflops=round(3 * flops * (stats["tc_vectors"] / N) / max(stats["tc_filter_us"], 1) / 1e6, 1),
                      plain_s=round(t_plain, 4), plain_queries=nq_plain, plain_queries_per_s=round(nq_plain / t_plain),
                      speedup=round((nq_plain / t_plain) and (Q / t_tc) / (nq_plain / t_plain), 1),
                      identical_to_plain=same, sync_form_s=round(t_sync, 4), sync_form_filter_us=sstats["tc_filter_us"],
                      sync_form_identical=bool(np.array_equal(sid, ids) and np.array_equal(sdist, dist)), **stats)))
