"""DeviceReplayMemory semantics vs the reference's ReplayMemory (rcbf_sac/replay_memory.py:4-35).

The ring content after every batch of a push schedule (wraps, exact fill, a batch larger than the capacity, an empty
batch, a single push) comes from the REAL class: tests/golden/replay_memory.npz is written by oracle/make_golden.py from
the unmodified reference source, and where /root/reference is present the real class is also driven live, side by side.
The same check runs on CPU tensors here and on the device under `-m gpu`."""
import numpy as np
import pytest
import torch

from oracle import make_golden, ref_loader
from sac_rcbf_b200.replay_memory import DeviceReplayMemory


def _rows(mem):
    n = len(mem)
    cols = [mem.state[:n], mem.action[:n], mem.reward[:n, None], mem.next_state[:n], mem.mask[:n, None], mem.t[:n, None],
            mem.next_t[:n, None]]
    return torch.cat([c.double().cpu() for c in cols], 1).numpy()


def _check_against_golden(g, device):
    cap = int(g["capacity"])
    mem = DeviceReplayMemory(cap, seed=0, obs_dim=7, action_dim=2, device=device, dtype=torch.float64)
    for k, b in enumerate(make_golden.replay_inputs(cap)):
        mem.batch_push(*b)
        assert len(mem) == g["after_%d" % k].shape[0] and mem.position == int(g["position_%d" % k])
        np.testing.assert_array_equal(_rows(mem), g["after_%d" % k])
    mem.push(np.ones(7), np.ones(2), 1.0, np.ones(7), 1.0, t=0.5, next_t=0.52)
    assert mem.position == int(g["position_push"])
    np.testing.assert_array_equal(_rows(mem), g["after_push"])
    smp = mem.sample(16)
    assert [x.dim() for x in smp] == list(g["sample_shapes"])
    return mem


def test_ring_semantics_match_reference_golden(golden):
    _check_against_golden(golden("replay_memory.npz"), "cpu")


@pytest.mark.skipif(not ref_loader.reference_available(), reason="reference source not mounted")
def test_ring_semantics_match_live_reference():
    ref = ref_loader.load_reference()
    cap = 23
    real = ref.ReplayMemory(cap, 0)
    mem = DeviceReplayMemory(cap, seed=0, obs_dim=7, action_dim=2, device="cpu", dtype=torch.float64)
    for b in make_golden.replay_inputs(cap, seed=4):
        real.batch_push(*b)
        mem.batch_push(*b)
        assert len(mem) == len(real) and mem.position == real.position
        want = np.stack([np.concatenate([np.ravel(x) for x in it]) for it in real.buffer]) if len(real) else np.zeros((0, 20))
        np.testing.assert_array_equal(_rows(mem), want)
    # sampling: same return structure (7 stacked arrays, batch first), every drawn row is a stored transition
    out_real, out_mine = real.sample(8), mem.sample(8)
    assert [np.asarray(x).shape for x in out_real] == [tuple(x.shape) for x in out_mine]
    stored = {tuple(np.round(r, 12)) for r in _rows(mem)}
    mine = torch.cat([out_mine[0], out_mine[1], out_mine[2][:, None], out_mine[3], out_mine[4][:, None],
                      out_mine[5][:, None], out_mine[6][:, None]], 1).numpy()
    assert all(tuple(np.round(r, 12)) in stored for r in mine)
    with pytest.raises(ValueError):
        real.sample(cap + 1)
    with pytest.raises(ValueError):
        mem.sample(cap + 1)


@pytest.mark.gpu
def test_ring_semantics_match_reference_golden_on_device(golden):
    mem = _check_against_golden(golden("replay_memory.npz"), "cuda")
    assert mem.state.is_cuda and mem.sample(4)[0].is_cuda


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
