"""Replay ring kernels on the GPU: latency of the reference's call shapes (sample(256), batch_push of one rollout batch)
and bandwidth of large draws, next to the eager-torch formulation (randperm over the ring + 7 index ops) they replace."""
import os, sys, time
sys.path.insert(0, os.getcwd())
import torch
from sac_rcbf_b200.replay_memory import DeviceReplayMemory

dev = torch.device("cuda")


def wall(fn, reps):
    for _ in range(5):
        fn()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(reps):
        fn()
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / reps * 1e6


def events(fn, reps):
    for _ in range(3):
        fn()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps * 1e3


cap = 1 << 22
mem = DeviceReplayMemory(cap, seed=0, obs_dim=7, action_dim=2)
n = cap
g = torch.Generator(device=dev); g.manual_seed(0)
src = [torch.randn(n, 7, device=dev), torch.randn(n, 2, device=dev), torch.randn(n, device=dev), torch.randn(n, 7, device=dev),
       torch.ones(n, device=dev), torch.rand(n, device=dev), torch.rand(n, device=dev)]
mem.batch_push(*src)
row_bytes = 4 * (7 + 2 + 1 + 7 + 3)


def eager_sample(b):
    idx = torch.randperm(mem.size, device=dev)[:b]
    return (mem.state[idx], mem.action[idx], mem.reward[idx], mem.next_state[idx], mem.mask[idx], mem.t[idx], mem.next_t[idx])


print("ring: %d rows x %d B" % (cap, row_bytes))
for b in (256, 4096, 65536, 1 << 20, 1 << 22):
    us = events(lambda: mem.sample(b), 20)
    us_w = wall(lambda: mem.sample(b), 200 if b <= 65536 else 20)
    us_e = wall(lambda: eager_sample(b), 20)
    print("sample(%8d): kernel path %9.1f us device / %9.1f us wall  (%.0f GB/s read+write)   eager torch %9.1f us wall"
          % (b, us, us_w, 2 * b * row_bytes / us / 1e3, us_e))
for b in (25, 5000, 1 << 20, 1 << 22):
    part = [s[:b] for s in src]
    us = events(lambda: mem.batch_push(*part), 20)
    us_w = wall(lambda: mem.batch_push(*part), 100 if b <= 5000 else 20)
    print("batch_push(%8d): %9.1f us device / %9.1f us wall  (%.0f GB/s read+write)" % (b, us, us_w, 2 * b * row_bytes / us / 1e3))
