"""Host-side multi-GPU logic on CPU: world_size-2 gloo process group (the N>1 bench path uses the same functions with
NCCL).  No GPU compute here."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from sac_rcbf_b200 import sharding


def test_shard_range_partitions_exactly():
    for n in (0, 1, 7, 512, 1 << 20, (1 << 20) + 3):
        for w in (1, 2, 3, 8):
            spans = [sharding.shard_range(n, r, w) for r in range(w)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            sizes = [hi - lo for lo, hi in spans]
            assert max(sizes) - min(sizes) <= 1


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, n_total, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    lo, hi = sharding.shard_range(n_total, rank, world)
    g = torch.Generator().manual_seed(1234)
    reward = torch.randn(n_total, generator=g)[lo:hi]
    cost = torch.rand(n_total, generator=g)[lo:hi]
    done = (torch.rand(n_total, generator=g) > 0.9)[lo:hi]
    counters = torch.arange(8, dtype=torch.int64) * (rank + 1)
    stats = sharding.reduce_rollout_stats(sharding.local_rollout_stats(reward, cost, done, done, counters))
    q.put((rank, stats))
    dist.barrier()
    dist.destroy_process_group()


def test_world_size_2_stats_reduction_gloo():
    world, n_total = 2, 1001
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, n_total, q)) for r in range(world)]
    for p in procs:
        p.start()
    got = dict(q.get(timeout=120) for _ in range(world))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    g = torch.Generator().manual_seed(1234)
    reward = torch.randn(n_total, generator=g)
    cost = torch.rand(n_total, generator=g)
    done = torch.rand(n_total, generator=g) > 0.9
    for r in range(world):
        s = got[r]
        assert s["instances"] == n_total
        assert abs(s["sum_reward"] - reward.double().sum().item()) < 1e-4
        assert abs(s["sum_cost"] - cost.double().sum().item()) < 1e-4
        assert s["n_done"] == int(done.sum()) and s["n_goal_met"] == int(done.sum())
        assert s["qp_uncertified"] == 1 * (1 + 2) and s["qp_trivial"] == 3 * (1 + 2) and s["qp_fallback"] == 5 * (1 + 2)


def test_single_process_reduction_is_identity():
    r = torch.ones(5)
    s = sharding.reduce_rollout_stats(sharding.local_rollout_stats(r, r * 0.1, r > 2))
    assert s["instances"] == 5 and abs(s["sum_cost"] - 0.5) < 1e-6 and s["n_done"] == 0
