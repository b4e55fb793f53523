#!/usr/bin/env python
"""bench_c5.py -- queries/sec on the SIFT1B-shaped workload (BASELINE.json configs[4]): the
code set is cut into parts by vector id, one part (or several) per GPU; every GPU generates its
part's bvecs-shaped vectors, encodes them (dpq_encode), builds the part's DeltaTree
(dpq_tree_build: GPU edge search + GPU layout) and opens it with dpq_index_open_part.  One step
= the same batch of queries scanned over every part, local top-k per part, ONE NCCL all-gather of
the key lists, device k-way merge.  Nothing under oracle/ is touched except the
bounded cpu_baseline sample on rank 0.

    python tools/bench_c5.py --codes-per-part 125000000                  # 1 GPU, one eighth of C5
    torchrun --nproc-per-node 8 tools/bench_c5.py --codes-per-part 125000000      # 1B codes, 8 GPUs
    python tools/bench_c5.py --codes-per-part 125000000 --parts-per-gpu 8         # 1B codes, 1 GPU

Prints one JSON line shaped like bench.py's.  The merged result is checked inside the run against
plain ADC over ALL codes for a handful of queries (float table entries, double sum = the
reference's arithmetic), computed with torch on the device from the encoder's output.
"""
import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import bench as B  # noqa: E402  (clock sampler, timed loop, peak)

DIM, PQ_M, PQ_K, TOPK = 128, 8, 256, 10


def gen_part_codes(torch, dpq, dev, cw, n, seed, chunk=1 << 21):
    """SIFT-like bvecs-shaped vectors (tests/datagen.sift_like's mixture, drawn with torch on the
    device: integer components 0..255) -> codes [n][8] on the device, never storing the vectors."""
    n_clusters, sigma, nb, w = 256, 14.0, DIM // 16, 16
    crng = np.random.default_rng(1234567)  # the centres of datagen.sift_like
    centres = torch.from_numpy(np.clip(crng.gamma(2.0, 22.0, size=(nb, n_clusters, w)), 0, 255)).to(dev, torch.float32)
    pop = 1.0 / np.arange(1, n_clusters + 1) ** 0.7
    pop = torch.from_numpy(pop / pop.sum()).to(dev, torch.float32)
    g = torch.Generator(device=dev)
    g.manual_seed(seed)
    codes = torch.empty((n, PQ_M), dtype=torch.uint8, device=dev)
    for s in range(0, n, chunk):
        c = min(chunk, n - s)
        x = torch.empty((c, DIM), dtype=torch.float32, device=dev)
        for b in range(nb):
            cid = torch.multinomial(pop, c, replacement=True, generator=g)
            blk = centres[b][cid] + sigma * torch.randn((c, w), device=dev, generator=g)
            x[:, b * w:(b + 1) * w] = blk.round_().clamp_(0, 255)
        dpq.encode_device(cw, x.data_ptr(), c, DIM, codes[s:].data_ptr())
        del x
    torch.cuda.synchronize()
    return codes


def plain_adc_topk(torch, dpq, cw, codes_dev, queries, k, id0):
    """Exact top-k of plain ADC over codes_dev for a few queries: (dist float32, global id)."""
    tabs = dpq.adc_tables(cw, queries)  # [q][M][K] float32, the reference's arithmetic
    out = []
    for t in tabs:
        tt = torch.from_numpy(t).to(codes_dev.device, torch.float64)
        d = torch.zeros(codes_dev.shape[0], dtype=torch.float64, device=codes_dev.device)
        for m in range(PQ_M):
            d += tt[m][codes_dev[:, m].long()]
        d32 = d.to(torch.float32)
        v, i = torch.topk(d32, k, largest=False, sorted=True)
        out.append((v.cpu().numpy(), i.cpu().numpy().astype(np.int64) + id0))
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--codes-per-part", type=int, default=125_000_000)
    ap.add_argument("--parts-per-gpu", type=int, default=1)
    ap.add_argument("--queries", type=int, default=10_000)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--check-queries", type=int, default=4)
    ap.add_argument("--cpu-baseline-queries", type=int, default=4)
    ap.add_argument("--opts", default="")
    ap.add_argument("--open-from-stream", action="store_true")
    args = ap.parse_args()

    import torch
    import deltapq_b200 as dpq
    import datagen as dg

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if dpq.device_count() < 1 or not torch.cuda.is_available():
        raise SystemExit("bench_c5.py needs a CUDA device (libdpq has no CPU fallback)")
    torch.cuda.set_device(local)
    dpq.set_device(local)
    dev = torch.device("cuda", local)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=dev)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(*vals):
        if world == 1:
            return [float(v) for v in vals]
        t = torch.tensor(list(vals), dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return [float(x) for x in t.tolist()]

    n_part, P, Q, k = args.codes_per_part, args.parts_per_gpu, args.queries, TOPK
    n_parts_total = world * P
    n_total = n_part * n_parts_total
    assert n_total < 2 ** 32 - 1, "positions are 32-bit"
    learn = dg.sift_like(20000, DIM, seed=3)
    cw = dg.roundtrip_codebook(dg.kmeans_codebook(learn, PQ_M, PQ_K, iters=6))
    queries = dg.sift_like(Q, DIM, seed=2)
    stream = torch.cuda.current_stream()

    parts, setup = [], {"gen_encode_s": 0.0, "edge_search_s": 0.0, "layout_s": 0.0, "fetch_s": 0.0, "open_s": 0.0}
    n_bytes_local, n_diffs_local = 0, 0
    checks = []
    payload0 = None
    nq_chk = min(args.check_queries, Q)
    for p in range(P):
        gp = rank * P + p  # global part index
        t0 = time.perf_counter()
        d_codes = gen_part_codes(torch, dpq, dev, cw, n_part, seed=1000 + gp)
        t1 = time.perf_counter()
        if nq_chk:
            checks.append(plain_adc_topk(torch, dpq, cw, d_codes, queries[:nq_chk], k, gp * n_part))
        codes = d_codes.cpu().numpy()
        del d_codes
        torch.cuda.empty_cache()
        t2 = time.perf_counter()
        if args.open_from_stream:  # the file path: host decode of the byte stream (program.cpp)
            tree = dpq.tree_build(codes, cw, h=1, method=1, want=("payload", "vec_id"))
            t3 = time.perf_counter()
            ix = dpq.DeltaTreeIndex(tree["payload"], n_part, PQ_M, PQ_K, pos2id=tree["vec_id"], first_pos=gp * n_part)
        else:  # dpq_index_open_tree: the scan program compiled on the GPU from the layout arrays
            tree = dpq.tree_build(codes, cw, h=1, method=1, want=("payload", "vec_id"), open_index_at=gp * n_part)
            t3 = time.perf_counter()
            ix = tree["index"]
        del codes
        ix.set_codebook(cw)
        for kv in (args.opts.split(",") if args.opts else []):
            kk, v = kv.split("=")
            ix.set_option(kk, int(v))
        ix.set_stream(stream.cuda_stream)
        t4 = time.perf_counter()
        setup["gen_encode_s"] += t1 - t0
        setup["edge_search_s"] += tree["edge_us"] / 1e6
        setup["layout_s"] += tree["layout_us"] / 1e6
        setup["fetch_s"] += (t3 - t2) - (tree["edge_us"] + tree["layout_us"]) / 1e6
        setup["open_s"] += t4 - t3
        n_bytes_local += len(tree["payload"])
        n_diffs_local += tree["n_diffs"]
        # keep what the id translation of the final check needs, drop the rest
        parts.append({"ix": ix, "vec_id": tree["vec_id"], "first_pos": gp * n_part})
        if rank == 0 and p == 0 and args.cpu_baseline_queries > 0:
            payload0 = tree["payload"]
        del tree

    d_q = torch.from_numpy(queries).to(dev)
    d_loc = torch.empty((P, Q, k), dtype=torch.int64, device=dev)
    d_one = torch.empty((Q, k), dtype=torch.int64, device=dev)
    d_all = torch.empty((world, Q, k), dtype=torch.int64, device=dev)
    d_mrg = torch.empty((Q, k), dtype=torch.int64, device=dev)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    ix0 = parts[0]["ix"]

    def step():
        for p, part in enumerate(parts):
            part["ix"].search_device(d_q.data_ptr(), Q, k, d_loc[p].data_ptr())
        local_keys = d_loc[0]
        if P > 1:
            ix0.merge_device(d_loc.data_ptr(), P, Q, k, d_one.data_ptr())
            local_keys = d_one
        if world > 1:
            dist.all_gather_into_tensor(d_all.view(-1), local_keys.reshape(-1))
            ix0.merge_device(d_all.data_ptr(), world, Q, k, d_mrg.data_ptr())
        else:
            d_mrg.copy_(local_keys)

    sampler = B.ClockSampler(local)
    if rank == 0:
        sampler.start()
        time.sleep(0.3)
    for _ in range(args.warmup):
        flush.zero_()
        step()
    barrier()
    for part in parts:
        part["ix"].set_option("timing_reset", 1)
    total_ms = B.timed_steps(torch, stream, flush, args.steps, step, barrier)
    clocks = sampler.stop() if rank == 0 else None
    scan_ns = sum(p_["ix"].stat("sum_scan_ns") for p_ in parts)
    scan8_ns = sum(p_["ix"].stat("sum_scan8_ns") for p_ in parts)
    lut_ns = sum(p_["ix"].stat("sum_lut_ns") for p_ in parts)
    calls = max(ix0.stat("timed_calls"), 1)
    coarse = ix0.stat("last_coarse") == 1
    for part in parts:
        part["ix"].sync()
    fallback = sum(p_["ix"].stat("last_fallback") for p_ in parts)
    launches = sum(p_["ix"].stat("last_launches") for p_ in parts) + (1 if P > 1 else 0) + (1 if world > 1 else 0)
    total_ms, scan_ms_max = max_over_ranks(total_ms, scan_ns / 1e6 / calls)

    # ---- e2e: host buffers through dpq_index_search on every part + host-side merge of the lists
    h_q = dpq.pinned_array(queries.shape, np.float32)
    h_q[...] = queries
    barrier()
    t0 = time.perf_counter()
    e2e_steps = max(1, min(args.steps, 2))
    for _ in range(e2e_steps):
        lists = [part["ix"].search(h_q, k) for part in parts]
        if P > 1 or world > 1:  # merge of the parts' lists (+ the NCCL gather of the per-rank keys)
            keys = np.stack([(l[2].view(np.uint32).astype(np.uint64) << np.uint64(32)) | l[0] for l in lists])
            d_keys = torch.from_numpy(keys.view(np.int64)).to(dev)
            ix0.merge_device(d_keys.data_ptr(), P, Q, k, d_one.data_ptr())
            if world > 1:
                dist.all_gather_into_tensor(d_all.view(-1), d_one.view(-1))
                ix0.merge_device(d_all.data_ptr(), world, Q, k, d_mrg.data_ptr())
                _ = d_mrg.cpu()
            else:
                _ = d_one.cpu()
    barrier()
    e2e_s = (time.perf_counter() - t0) / e2e_steps
    (e2e_s,) = max_over_ranks(e2e_s)

    # ---- check: merged top-k == plain ADC over ALL codes, for the first few queries
    step()
    torch.cuda.synchronize()
    mpos, mdist = dpq.unpack_keys(d_mrg.cpu().numpy().view(np.uint64))
    assert np.all(np.diff(mdist.astype(np.float64), axis=1) >= 0), "merged top-k not ascending"
    check = None
    if nq_chk:
        loc_d = np.stack([np.stack([c[q][0] for c in checks]) for q in range(nq_chk)])  # [q][P][k]
        loc_i = np.stack([np.stack([c[q][1] for c in checks]) for q in range(nq_chk)])
        if world > 1:
            td = torch.from_numpy(loc_d).to(dev)
            ti = torch.from_numpy(loc_i).to(dev)
            gd = [torch.empty_like(td) for _ in range(world)]
            gi = [torch.empty_like(ti) for _ in range(world)]
            dist.all_gather(gd, td)
            dist.all_gather(gi, ti)
            loc_d = np.concatenate([x.cpu().numpy() for x in gd], axis=1)
            loc_i = np.concatenate([x.cpu().numpy() for x in gi], axis=1)
        # positions -> global ids needs every part's vec_id: each rank translates its own hits
        ids = np.full(mpos[:nq_chk].shape, -1, np.int64)
        for part in parts:
            a = part["first_pos"]
            sel = (mpos[:nq_chk] >= a) & (mpos[:nq_chk] < a + n_part)
            ids[sel] = part["vec_id"][(mpos[:nq_chk][sel] - a).astype(np.int64)].astype(np.int64) + a
        if world > 1:
            ti = torch.from_numpy(ids).to(dev)
            dist.all_reduce(ti, op=dist.ReduceOp.MAX)
            ids = ti.cpu().numpy()
        ok = True
        for q in range(nq_chk):
            d_all_q, i_all_q = loc_d[q].ravel(), loc_i[q].ravel()
            order = np.lexsort((i_all_q, d_all_q))[:k]
            ok &= bool(np.array_equal(d_all_q[order], mdist[q]))
            kth = d_all_q[order][-1]
            ok &= {int(x) for x, d in zip(ids[q], mdist[q]) if d < kth} == {int(i_all_q[o]) for o in order if d_all_q[o] < kth}
        check = {"queries": nq_chk, "equals_plain_adc_over_all_codes": ok}
        assert ok, "merged top-k differs from plain ADC over all codes"

    n_bytes_tot, n_diffs_tot = n_bytes_local, n_diffs_local
    if world > 1:
        t = torch.tensor([n_bytes_local, n_diffs_local], dtype=torch.int64, device=dev)
        dist.all_reduce(t)
        n_bytes_tot, n_diffs_tot = int(t[0]), int(t[1])
        ts = torch.tensor([setup[k_] for k_ in sorted(setup)], dtype=torch.float64, device=dev)
        dist.all_reduce(ts, op=dist.ReduceOp.MAX)
        setup = {k_: float(v) for k_, v in zip(sorted(setup), ts.tolist())}

    if rank == 0:
        peak, peak_src = B.measured_peak()
        alg_bytes = Q * (n_bytes_local + P * (4 * DIM + 8 * k))  # per GPU per step (all its parts)
        dom_s = (scan8_ns if coarse else scan_ns) / 1e9 / calls
        achieved = alg_bytes / dom_s / 1e9
        qps = Q * args.steps / (total_ms / 1e3)
        cpu_base = None
        if payload0 is not None:
            from oracle import pyoracle as po
            nq = args.cpu_baseline_queries
            kind = "reference" if po.have_ref() else "port"
            t0 = time.perf_counter()
            if kind == "reference":
                _, _, secs = po.ref_scan(payload0, n_part, cw, np.ascontiguousarray(queries[:nq]), k)
            else:
                for qv in queries[:nq]:
                    po.scan(payload0, n_part, cw, qv, k)
                secs = time.perf_counter() - t0
            cpu_base = {"value": nq / secs / n_parts_total, "unit": B.UNIT, "cores": 1, "kind": kind,
                        "sample": f"{nq} queries scanned over ONE of the {n_parts_total} parts ({n_part} codes, in-memory scan "
                                  f"DCAT.h:3731, 1 thread, {secs:.1f} s), divided by the number of parts"}
        line = {
            "metric": f"queries/sec @top10 on {n_total / 1e9:.3g}B-code DeltaTree forest (SIFT1B-shaped synthetic, M=8 K=256)",
            "value": qps, "unit": B.UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": total_ms / args.steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "u8 coarse filter + u16 fixed-point sample + f64 exact re-score",
            "data": "synthetic",
            "config": {"workload": f"SIFT1B-shaped synthetic bvecs: {n_total} x 128-d, M=8 K=256 h=1, {Q} queries per step, top-10 "
                                   f"(BASELINE configs[4]{'' if n_total >= 10**9 else ', reduced: ' + str(n_parts_total) + ' part(s) of 1/8'})",
                       "n_codes": n_total, "codes_per_part": n_part, "parts_per_gpu": P, "queries_per_step": Q, "topk": k,
                       "n_bytes": n_bytes_tot, "mean_diffs_per_node": round(n_diffs_tot / max(n_total - n_parts_total, 1), 3),
                       "sharding": f"forest: {n_parts_total} independently built DeltaTrees (parts by vector id), {P} per GPU; "
                                   f"same {Q} queries on every part; NCCL all-gather of {Q * k * 8} B/rank + device merge",
                       "l2": "256 MiB buffer written before every timed step; each part's program (16 B/node) exceeds L2",
                       "tree": "built in the run by libdpq on the scanning GPU (encode + edge search + layout)",
                       "setup_s_max_over_ranks": {k_: round(v, 2) for k_, v in setup.items()},
                       "open": "host decode of the byte stream (open_s)" if args.open_from_stream else
                               "scan program compiled on the GPU from the layout arrays (inside fetch_s)",
                       "opts": args.opts or "default"},
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "traffic": None, "kernel": "scan8_kernel" if coarse else "scan2_kernel",
                         "algorithmic_bytes_per_launch": alg_bytes / P, "kernel_ms_per_launch": dom_s * 1e3 / P,
                         "peak_source": peak_src,
                         "note": "per GPU (rank 0); effective bandwidth over the on-disk stream bytes, one pass serves 112 queries; "
                                 f"physical floor = {Q // 112 + 1} passes x 16 B/node program"},
            "cpu_baseline": cpu_base,
            "e2e": {"value": Q / e2e_s, "unit": B.UNIT, "h2d_bytes_per_step": world * P * Q * DIM * 4,
                    "d2h_bytes_per_step": world * P * Q * k * 8},
            "gpu_launches": int(launches * args.steps),
            "clocks": clocks,
            "breakdown_ms_per_step": {"lut": lut_ns / 1e6 / calls, "all_scan_phases": scan_ns / 1e6 / calls,
                                      "coarse_scan_kernel": scan8_ns / 1e6 / calls if coarse else None,
                                      "scan_max_over_ranks": scan_ms_max, "exact_fallback_queries": fallback},
            "check": check,
        }
        print(json.dumps(line), flush=True)
    for part in parts:
        part["ix"].close()
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    _holder, _print = [], print
    with B.StdoutToStderr():  # library banners (NCCL) go to stderr; stdout carries the JSON line only
        import builtins
        builtins.print = lambda *a, **k: _holder.append(" ".join(str(x) for x in a))
        try:
            _rc = main()
        finally:
            builtins.print = _print
    for _l in _holder:
        _print(_l, flush=True)
    sys.exit(_rc)
