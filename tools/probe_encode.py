"""Developer probe: tensor-core encoder vs SIMT encoder, timing and equality (prints as it goes)."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import deltapq_b200 as dpq

rng = np.random.default_rng(0)
for M, Ds in ((8, 16), (16, 8), (8, 4)):
    cw = (rng.random((M, 256, Ds)) * 140).astype(np.float32)
    for n in (256, 100000, 1000000):
        x = rng.integers(0, 256, size=(n, M * Ds)).astype(np.float32)
        out = {}
        for tc in ("1", "0"):
            os.environ["DPQ_ENCODE_TC"] = tc
            t = time.perf_counter()
            codes = dpq.encode(cw, x)
            dt = time.perf_counter() - t
            out[tc] = codes
            print(dict(M=M, Ds=Ds, n=n, tc=dpq.encode_stat("tc"), wall_ms=round(dt * 1e3, 2), kernel_us=dpq.encode_stat("kernel_us")), flush=True)
        print("   equal:", bool(np.array_equal(out["1"], out["0"])), flush=True)
