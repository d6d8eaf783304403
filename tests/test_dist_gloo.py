"""CPU, world_size 2 over gloo: the result gather reassembles per-stream records in global stream
order from unequal shards (the N>1 host path of bench.py / ShardedBank without GPUs)."""
import os
import socket
import sys

import numpy as np
import pytest

from conftest import REPO


def _worker(rank, world, port, n_total, out_dir):
    sys.path.insert(0, REPO)
    import torch
    import torch.distributed as dist
    from easywakeword_b200.dist import ResultGather, shard_range
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    g = ResultGather(n_total, world, rank, device="cpu")
    a, b = shard_range(n_total, world, rank)
    # what K2/K3 would have written for the local streams: score = global id / 7, flags = id * 3 + 1
    ids = np.arange(a, b)
    rec = np.zeros((g.pad, 2), dtype=np.int32)
    rec[: b - a, 0] = (ids / 7.0).astype(np.float32).view(np.int32)
    rec[: b - a, 1] = (ids * 3 + 1).astype(np.uint32).view(np.int32)
    g.local.copy_(torch.from_numpy(rec))
    for _ in range(3):                       # the gather is per step: repeat it
        out = g.gather()
    np.save(os.path.join(out_dir, f"r{rank}.npy"), out.numpy())
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("n_total", [10, 7])
def test_result_gather_world2(tmp_path, n_total):
    import torch.multiprocessing as mp
    from easywakeword_b200.dist import ResultGather
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    mp.spawn(_worker, args=(2, port, n_total, str(tmp_path)), nprocs=2, join=True)
    ids = np.arange(n_total)
    for r in range(2):
        got = np.load(os.path.join(str(tmp_path), f"r{r}.npy"))
        assert got.shape == (n_total, 2)
        import torch
        dec = ResultGather.decode(torch.from_numpy(got))
        assert np.array_equal(dec["score"], (ids / 7.0).astype(np.float32))
        assert np.array_equal(dec["flags"], (ids * 3 + 1).astype(np.uint32))
