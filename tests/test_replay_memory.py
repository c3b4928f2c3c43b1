"""DeviceReplayMemory semantics vs rcbf_sac/replay_memory.py (restated inline), on CPU tensors (the class is plain
torch indexing; the GPU test repeats a subset on the device)."""
import numpy as np
import torch

from sac_rcbf_b200.replay_memory import DeviceReplayMemory


class _RefMemory:                      # rcbf_sac/replay_memory.py:4-35, minus sampling
    def __init__(self, capacity):
        self.capacity, self.buffer, self.position = capacity, [], 0

    def push(self, *item):
        if len(self.buffer) < self.capacity:
            self.buffer.append(None)
        self.buffer[self.position] = item
        self.position = (self.position + 1) % self.capacity


def test_ring_semantics_match_reference():
    cap, od, ad = 37, 7, 2
    mem = DeviceReplayMemory(cap, seed=0, obs_dim=od, action_dim=ad, device="cpu", dtype=torch.float64)
    ref = _RefMemory(cap)
    rng = np.random.default_rng(0)
    for n in (5, 1, 20, 30, 3, 80, 0, 11):          # wraps, exact fill, a batch larger than the capacity, empty
        s, a, r = rng.normal(size=(n, od)), rng.normal(size=(n, ad)), rng.normal(size=n)
        s2, m, t = rng.normal(size=(n, od)), (rng.random(n) > 0.2), rng.random(n)
        mem.batch_push(s, a, r, s2, m, t, t + 0.02)
        for i in range(n):
            ref.push(s[i], a[i], r[i], s2[i], m[i], t[i], t[i] + 0.02)
        assert len(mem) == len(ref.buffer) and mem.position == ref.position
        for slot, item in enumerate(ref.buffer):
            np.testing.assert_array_equal(mem.state[slot].numpy(), item[0])
            np.testing.assert_array_equal(mem.action[slot].numpy(), item[1])
            assert mem.reward[slot].item() == item[2] and mem.mask[slot].item() == float(item[4])
            assert mem.t[slot].item() == item[5] and mem.next_t[slot].item() == item[6]
    mem.push(np.ones(od), np.ones(ad), 1.0, np.ones(od), 1.0, t=0.5, next_t=0.52)
    assert mem.reward[(mem.position - 1) % cap].item() == 1.0


def test_sample_without_replacement_and_shapes():
    mem = DeviceReplayMemory(100, seed=1, obs_dim=3, action_dim=1, device="cpu")
    x = torch.arange(60, dtype=torch.float32)
    mem.batch_push(x[:, None].expand(60, 3), x[:, None], x, x[:, None].expand(60, 3), torch.ones(60), x, x + 1)
    s, a, r, s2, m, t, nt = mem.sample(60)
    assert s.shape == (60, 3) and a.shape == (60, 1) and r.shape == (60,)
    assert sorted(r.tolist()) == list(range(60))            # every stored item exactly once
    assert torch.equal(s[:, 0], r) and torch.equal(nt, t + 1)
    try:
        mem.sample(61)
        assert False
    except ValueError:
        pass
