"""Profiling target: a few 4 Mi-row draws and pushes on the replay ring (ncu -k regex:k_replay)."""
import os, sys
sys.path.insert(0, os.getcwd())
import torch
from sac_rcbf_b200.replay_memory import DeviceReplayMemory
dev = torch.device("cuda")
n = 1 << 22
mem = DeviceReplayMemory(n, seed=0, obs_dim=7, action_dim=2)
src = [torch.randn(n, 7, device=dev), torch.randn(n, 2, device=dev), torch.randn(n, device=dev), torch.randn(n, 7, device=dev),
       torch.ones(n, device=dev), torch.rand(n, device=dev), torch.rand(n, device=dev)]
for _ in range(4):
    mem.batch_push(*src)
    mem.sample(n)
torch.cuda.synchronize()
print("done")
