"""N > 1 host logic on CPU: two ranks over gloo (world_size 2, 127.0.0.1).  Each rank compiles
ITS shard of the tree with the product's tree compiler (libdpq, host side), scores it with the
test interpreter of the device program, keeps a local top-k with GLOBAL positions, all-gathers
the uint64 keys exactly as bench.py does over NCCL, and merges; the merged list must equal the
unsharded top-k of the oracle.  (The device merge kernel and the sharded GPU scan are covered by
tests/test_gpu_parity.py::test_sharded_search_merges_to_whole.)"""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _rank_main(rank, world, port, k, out_dir):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import deltapq_b200 as dpq
    from helpers import interpret_any
    from oracle import pyoracle as po
    z = np.load(os.path.join(ROOT, "tests", "golden", "sift_n4000_m8.npz"))
    n = int(z["n"])
    prog = dpq.compile_program(z["payload"], n, 8, 256, rank=rank, n_ranks=world)
    Q = 6
    keys = np.full((Q, k), np.iinfo(np.int64).max, np.int64)
    for qi in range(Q):
        table = po.lut(z["cw"], z["queries"][qi]).ravel()
        pos, d = interpret_any(prog, table.astype(np.float64))
        d32 = d.astype(np.float32)
        key = (d32.view(np.uint32).astype(np.uint64) << np.uint64(32)) | pos.astype(np.uint64)
        key.sort()
        keys[qi, :min(k, len(key))] = key[:k].view(np.int64)
    gathered = [torch.empty((Q, k), dtype=torch.int64) for _ in range(world)]
    dist.all_gather(gathered, torch.from_numpy(keys))
    allk = torch.stack(gathered).numpy().view(np.uint64)            # [world][Q][k], like d_all in bench.py
    merged = np.sort(allk.transpose(1, 0, 2).reshape(Q, -1), axis=1)[:, :k]
    # shards partition the tree
    cnt = torch.tensor([prog["n_local"]])
    dist.all_reduce(cnt)
    assert int(cnt) == n
    if rank == 0:
        np.save(os.path.join(out_dir, "merged.npy"), merged)
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.timeout(300)
def test_two_rank_sharded_topk_equals_unsharded(tmp_path):
    import deltapq_b200 as dpq
    from helpers import assert_topk_equal
    from oracle import pyoracle as po
    k, world = 10, 2
    mp.spawn(_rank_main, args=(world, _free_port(), k, str(tmp_path)), nprocs=world, join=True)
    merged = np.load(tmp_path / "merged.npy")
    pos, distv = dpq.unpack_keys(merged)
    z = np.load(os.path.join(ROOT, "tests", "golden", "sift_n4000_m8.npz"))
    for qi in range(merged.shape[0]):
        opos, odist = po.scan(z["payload"], int(z["n"]), z["cw"], z["queries"][qi], k)
        np.testing.assert_allclose(distv[qi], odist, rtol=1e-5)
        assert_topk_equal(pos[qi], distv[qi], opos, odist)
