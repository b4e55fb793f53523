"""Secondary kernels vs the UNMODIFIED reference on the same box (developer measurement, not
bench.py): encode, approx_tree, groundtruth at BASELINE configs[1] shape.  Prints one JSON line
per stage.  The reference legs run the binaries in oracle/_ref (all host threads, OpenMP)."""
import json, os, subprocess, sys, tempfile, time, shutil
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import datagen as dg
import deltapq_b200 as dpq
from oracle import pyoracle as po

N = int(sys.argv[1]) if len(sys.argv) > 1 else 1000000
NQ = int(sys.argv[2]) if len(sys.argv) > 2 else 100
M = int(sys.argv[3]) if len(sys.argv) > 3 else 8
run_ref = os.path.exists(os.path.join(po.REF_DIR, "pqtree")) and "--no-ref" not in sys.argv
d = tempfile.mkdtemp(prefix="dpq_sec_")
try:
    base, queries, cw = dg.make_dataset(d, N, NQ, M=M, K=256, d=128, seed=0, n_learn=20000)
    os.makedirs(d + "/groundtruth")
    BIN = os.path.join(ROOT, "deltapq_b200", "bin")
    common = ["-dataset", d, "-m", str(M), "-k", "256", "-N", str(N), "-ext", "fvecs"]

    def timed(cmd):
        t = time.perf_counter()
        r = subprocess.run(cmd, capture_output=True, text=True)
        dt = time.perf_counter() - t
        if r.returncode != 0:
            raise RuntimeError(f"{cmd}: rc={r.returncode}\n{r.stdout[-1500:]}\n{r.stderr[-1500:]}")
        return dt

    # in-process kernel-level times (host buffers in, host buffers out)
    dpq.encode(cw, base[:1000])
    t = time.perf_counter(); codes = dpq.encode(cw, base); t_enc = time.perf_counter() - t
    # device-resident input (what the 10^9-code pipeline feeds the encoder): kernel rate against the FP32
    # issue peak (the reference's arithmetic is FSUB / FMUL / FADD per dimension, no FMA: 3 ops per term)
    dx = dpq.DeviceBuffer(base.nbytes).upload(base)
    dc = dpq.DeviceBuffer(N * M)
    dpq.encode_device(cw, dx.ptr.value, min(N, 4096), base.shape[1], dc.ptr.value)
    t_dev = None
    for _ in range(3):
        t = time.perf_counter(); dpq.encode_device(cw, dx.ptr.value, N, base.shape[1], dc.ptr.value); dt_ = time.perf_counter() - t
        t_dev = dt_ if t_dev is None else min(t_dev, dt_)
    enc_tc, enc_kernel_us = dpq.encode_stat("tc"), dpq.encode_stat("kernel_us")
    assert np.array_equal(dc.download(np.uint8, (N, M)), codes)
    # the SIMT encoder on the same device-resident input
    os.environ["DPQ_ENCODE_TC"] = "0"
    t_simt = None
    for _ in range(2):
        t = time.perf_counter(); dpq.encode_device(cw, dx.ptr.value, N, base.shape[1], dc.ptr.value); dt_ = time.perf_counter() - t
        t_simt = dt_ if t_simt is None else min(t_simt, dt_)
    simt_kernel_us = dpq.encode_stat("kernel_us")
    del os.environ["DPQ_ENCODE_TC"]
    assert np.array_equal(dc.download(np.uint8, (N, M)), codes)
    dx.free(); dc.free()
    Ds = cw.shape[2]
    fp32_ops = 3.0 * N * M * 256 * Ds
    t = time.perf_counter(); tree = dpq.tree_build(codes, cw); t_tree = time.perf_counter() - t
    t = time.perf_counter(); tree = dpq.tree_build(codes, cw); t_tree = min(t_tree, time.perf_counter() - t)  # second call: warm
    t = time.perf_counter(); ge, gr = dpq.find_edges(codes, 256, 1, 1); t_edges = time.perf_counter() - t
    t = time.perf_counter(); gid, gd = dpq.groundtruth(base, queries, 10); t_gt = time.perf_counter() - t
    out = dict(N=N, M=M, encode_s=round(t_enc, 3), encode_vec_per_s=round(N / t_enc),
               encode_device_resident_s=round(t_dev, 4), encode_device_resident_vec_per_s=round(N / t_dev),
               encode_tensor_core_path=enc_tc, encode_kernel_us=enc_kernel_us,
               encode_kernel_vec_per_s=round(N / max(enc_kernel_us, 1) * 1e6),
               encode_useful_tflops=round(2.0 * N * 256 * M * Ds / max(enc_kernel_us, 1) / 1e6, 1),   # 2 N K D
               encode_mma_issued_tflops=round(2.0 * N * M * 256 * 64 / max(enc_kernel_us, 1) / 1e6, 1),  # hi.hi + hi.lo + lo.hi + norm, K padded to 64
               encode_simt_device_resident_s=round(t_simt, 4), encode_simt_kernel_us=simt_kernel_us,
               encode_simt_fp32_issue_frac=round(fp32_ops / max(simt_kernel_us, 1) * 1e6 / (148 * 128 * 1.965e9), 3),
               find_edges_s=round(t_edges, 3), tree_build_s=round(t_tree, 3),
               groundtruth_s=round(t_gt, 3), groundtruth_ms_per_query=round(t_gt / NQ * 1e3, 2),
               tree_edge_stage_s=round(tree["edge_us"] / 1e6, 3), tree_layout_stage_s=round(tree["layout_us"] / 1e6, 3),
               n_bytes=int(len(tree["payload"])), n_diffs=tree["n_diffs"])
    # CLI wall times (file I/O included, like the reference's own timers)
    out["cli_encode_s"] = round(timed([BIN + "/pqtree", "-task", "encode"] + common), 2)
    out["cli_approx_tree_s"] = round(timed([BIN + "/deltapq", "-task", "approx_tree", "-h", "1", "-diff", str(M)] + common), 2)
    out["cli_groundtruth_s"] = round(timed([BIN + "/pqtree", "-task", "groundtruth", "-query_size", str(NQ), "-topk", "10"] + common), 2)
    if run_ref:
        r = tempfile.mkdtemp(prefix="dpq_secref_")
        try:
            for f in ("base.fvecs", "query.fvecs", f"M{M}K256codewords.txt"):
                os.symlink(os.path.join(d, f), os.path.join(r, f))
            os.makedirs(r + "/groundtruth")
            rc = ["-dataset", r] + common[2:]
            out["ref_encode_s"] = round(timed([po.REF_DIR + "/pqtree", "-task", "encode"] + rc), 2)
            if M == 8:
                out["ref_approx_tree_s"] = round(timed([po.REF_DIR + "/deltapq_canon", "-task", "approx_tree", "-h", "1", "-diff", "8"] + rc), 2)
            out["ref_groundtruth_s"] = round(timed([po.REF_DIR + "/pqtree", "-task", "groundtruth", "-query_size", str(NQ), "-topk", "10"] + rc), 2)
            out["ref_threads"] = len(os.sched_getaffinity(0))
            same = np.array_equal(np.fromfile(f"{d}/codes.bin.plain.M{M}K256N{N}", np.uint8), np.fromfile(f"{r}/codes.bin.plain.M{M}K256N{N}", np.uint8))
            out["codes_identical"] = bool(same)
            if M == 8:
                nm = f"M8K256_Approx_compressed_codes_opt_N{N}"
                out["tree_identical"] = bool(np.array_equal(np.fromfile(f"{d}/{nm}", np.uint8), np.fromfile(f"{r}/{nm}", np.uint8)))
        finally:
            shutil.rmtree(r, ignore_errors=True)
    print(json.dumps(out))
finally:
    shutil.rmtree(d, ignore_errors=True)
