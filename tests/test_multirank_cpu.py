"""CPU, world_size 2 over gloo: the N>1 launch style (one process per GPU, static partition, no data-path collective).
Each rank takes ``partition_windows(n, world)[rank]``, runs a stand-in for the device step on its shard and reports a
time; rank 0 checks that the shards tile the job in window order and that the job time is the max over ranks — the
same plumbing bench.py --gpus N uses with NCCL."""
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


def _worker(rank, world, port, n_windows, out_dir):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from whisper_aries_b200 import partition_windows
    from whisper_aries_b200.scheduler import ChunkScheduler
    start, stop = partition_windows(n_windows, world)[rank]
    windows = np.arange(n_windows * 4, dtype=np.float32).reshape(n_windows, 4)     # every rank sees the job description
    local = np.zeros((n_windows, 1), np.float32)

    def fake_device(w, a, b, o):
        o[a:b, 0] = w[a:b].sum(axis=1)

    res = ChunkScheduler([fake_device]).run(windows[start:stop], local[start:stop])
    assert all(r.success for r in res)
    dist.barrier()
    t = torch.tensor([1.0 + rank], dtype=torch.float64)            # pretend rank 1 was slower
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    gathered = [None] * world
    dist.all_gather_object(gathered, (start, stop, local[start:stop].copy()))      # test-only gather, not the data path
    if rank == 0:
        full = np.concatenate([g[2] for g in gathered])
        np.save(os.path.join(out_dir, "full.npy"), full)
        np.save(os.path.join(out_dir, "meta.npy"), np.array([float(t.item())] + [x for g in gathered for x in g[:2]]))
    dist.destroy_process_group()


@pytest.mark.parametrize("n_windows", [120, 7])
def test_two_ranks_tile_the_job(tmp_path, n_windows):
    world = 2
    mp.spawn(_worker, args=(world, _free_port(), n_windows, str(tmp_path)), nprocs=world, join=True)
    full = np.load(tmp_path / "full.npy")
    meta = np.load(tmp_path / "meta.npy")
    ref = np.arange(n_windows * 4, dtype=np.float32).reshape(n_windows, 4).sum(axis=1, keepdims=True)
    assert np.array_equal(full, ref)                                # shards concatenate in window order, nothing lost
    assert meta[0] == 2.0                                           # job time = max over ranks
    assert meta[1:].tolist() == [0, -(-n_windows // 2), -(-n_windows // 2), n_windows]
