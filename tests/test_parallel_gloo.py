"""Multi-rank host logic on CPU: world_size-2 gloo (the GPU path uses the same code with NCCL)."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from loco_mujoco_b200.parallel import shard_range, gather_rollout, aggregate_throughput, RolloutGather


def test_shard_range_partitions_the_env_axis():
    for n, w in [(4096, 8), (4097, 8), (10, 3), (5, 8)]:
        spans = [shard_range(n, w, r) for r in range(w)]
        assert spans[0][0] == 0 and sum(c for _, c in spans) == n
        for (o1, c1), (o2, _) in zip(spans, spans[1:]):
            assert o1 + c1 == o2
        assert max(c for _, c in spans) - min(c for _, c in spans) <= 1


def _worker(rank, world, port, n_total, out_dir):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    off, cnt = shard_range(n_total, world, rank)
    D = 5
    ids = torch.arange(off, off + cnt, dtype=torch.float32)
    obs = ids[:, None] * 10 + torch.arange(D, dtype=torch.float32)[None, :]
    rew = ids * 0.5
    done = (ids.long() % 3) == 0
    g_obs, g_rew, g_done = gather_rollout(obs, rew, done)
    thr, tmax = aggregate_throughput(local_units=cnt * 7, local_seconds=1.0 + rank)
    # chunked rollout gather: 7 steps, chunk 3 -> two complete chunks gathered, the last one returned by finish()
    rec = 37
    g = RolloutGather(rec, 3, "cpu")
    n_g = 0
    for k in range(7):
        g.slot().copy_(torch.full((rec,), 16 * rank + k, dtype=torch.uint8))
        n_g += int(g.advance())
    chunk = g.finish()
    np.savez(os.path.join(out_dir, "r%d.npz" % rank), obs=g_obs.numpy(), rew=g_rew.numpy(), done=g_done.numpy(),
             thr=thr, tmax=tmax, chunk=chunk.numpy(), n_g=n_g, stride=g.stride)
    dist.destroy_process_group()


def test_gather_and_throughput_world2(tmp_path):
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    n_total, world = 11, 2
    mp.spawn(_worker, args=(world, port, n_total, str(tmp_path)), nprocs=world, join=True)
    ids = np.arange(n_total, dtype=np.float32)
    for r in range(world):
        d = np.load(tmp_path / ("r%d.npz" % r))
        assert np.array_equal(d["obs"], ids[:, None] * 10 + np.arange(5, dtype=np.float32)[None, :])
        assert np.array_equal(d["rew"], ids * 0.5)
        assert np.array_equal(d["done"], (ids.astype(np.int64) % 3) == 0)
        assert d["tmax"] == pytest.approx(2.0) and d["thr"] == pytest.approx(n_total * 7 / 2.0)
        assert int(d["n_g"]) == 2 and int(d["stride"]) == 48 and d["chunk"].shape == (world, 3, 37)
        for src in range(world):                 # second chunk = steps 3, 4, 5 of every rank, in rank order
            for t in range(3):
                assert (d["chunk"][src, t] == 16 * src + 3 + t).all()
