"""Closed-loop parity (size-independent property of the path as a SYSTEM): T consecutive fused safe steps on the device
(float32 state carried on the GPU) against the oracle loop -- reference assembly + exact QP + numpy float64 env -- fed
with the same RL actions and disturbance estimates.  One-step tests cannot see an error that only matters once it is
fed back through the dynamics (e.g. a biased correction that slowly walks an instance into a hazard)."""
import types

import numpy as np
import pytest
import torch

from oracle import rcbf_oracle as O

pytestmark = pytest.mark.gpu


def test_unicycle_closed_loop_tracks_oracle_for_60_steps():
    import sac_rcbf_b200 as S
    B, T = 192, 60
    rng = np.random.default_rng(2024)
    # hazard-heavy start: look-ahead point 0.75 .. 1.3 from a hazard centre (just outside r_c = 0.72), any heading
    hz = O.UNICYCLE["hazards_locations"]
    idx = rng.integers(0, len(hz), B)
    r, phi = rng.uniform(0.78, 1.3, B), rng.uniform(-np.pi, np.pi, B)
    st0 = np.stack([hz[idx, 0] + r * np.cos(phi), hz[idx, 1] + r * np.sin(phi), rng.uniform(-np.pi, np.pi, B)], 1)
    env = S.UnicycleEnv(num_envs=B)
    layer = S.CBFQPLayer(env, types.SimpleNamespace(cuda=True), gamma_b=20, k_d=3.0, l_p=0.03)
    env.reset()
    env.state = torch.as_tensor(st0, dtype=torch.float32).cuda()
    st = env.state.cpu().numpy().astype(np.float64)          # the oracle starts from the SAME float32 numbers
    step = np.zeros(B, np.int64)
    last = O.unicycle_goal_dist(st)
    mu = np.zeros((B, 3), np.float32)
    sg = np.full((B, 3), 0.2, np.float32)                    # prior MAX_STD (dynamics.py:24)
    tt = torch.from_numpy
    alive = np.ones(B, bool)
    worst = 0.0
    for k in range(T):
        u = rng.uniform(-1, 1, (B, 2)).astype(np.float32)
        us_dev, obs, rew, done, info = env.safe_step(layer, tt(u).cuda(), tt(mu).cuda(), tt(sg).cuda())
        us_ora = O.safe_action("Unicycle", tt(st.astype(np.float32)), tt(u), tt(mu), tt(sg), solver="exact",
                               gamma_b=20.0).numpy()
        out = O.unicycle_env_step(st, us_ora.astype(np.float64), step, last)
        st, step, last = out["state"], out["episode_step"], out["last_goal_dist"]
        alive &= ~out["done"]                                 # goal reached -> the reference would reset; stop comparing
        dev_st = env.state.cpu().numpy().astype(np.float64)
        err = np.abs(dev_st - st).max(1)
        worst = max(worst, float(np.median(err[alive])))
        assert np.isfinite(dev_st).all()
    # kinks (active-set switches, the [-1, 1] action clip of unicycle_env.py:62) can split a few trajectories for good;
    # everything else must still be on top of the oracle after 60 fed-back steps
    close = err[alive] < 2e-3
    assert close.mean() > 0.97, (close.mean(), np.sort(err[alive])[-5:])
    assert worst < 1e-4, worst
    # and the safety property itself: nobody driven by the layer ends inside a hazard's cost radius
    d2 = ((dev_st[:, None, :2] - hz[None]) ** 2).sum(-1)
    assert (d2.min(1) >= 0.6 ** 2).mean() > 0.97
    stats = layer.solver_stats()
    assert stats["nan"] == 0 and stats["uncertified"] == 0


def test_cars_closed_loop_tracks_oracle_for_40_steps():
    """Same for SimulatedCars (positions ~ 100 m in float32, so the comparison is relative to the gap scale; instances
    that come within 1e-3 of a `< 6 / < 13` braking switch are dropped -- a discontinuity of the reference model)."""
    import sac_rcbf_b200 as S
    B, T = 160, 40
    st0, _, _, sg, t0 = O.synth_cars(B, seed=77)
    env = S.SimulatedCarsEnv(num_envs=B)
    layer = S.CBFQPLayer(env, types.SimpleNamespace(cuda=True), gamma_b=20, k_d=3.0, l_p=0.03)
    env.reset()
    env.state = torch.as_tensor(st0).cuda()
    env._t.copy_(torch.as_tensor(t0).cuda())
    st = st0.astype(np.float64)
    t = t0.astype(np.float64)
    step = np.zeros(B, np.int64)
    mu = np.zeros((B, 10), np.float32)
    rng = np.random.default_rng(5)
    tt = torch.from_numpy
    keep = np.ones(B, bool)
    for k in range(T):
        keep &= O.cars_threshold_margin(st) > 1e-3
        u = rng.uniform(-1, 1, (B, 1)).astype(np.float32)
        env.safe_step(layer, tt(u).cuda(), tt(sg).cuda())
        us = O.safe_action("SimulatedCars", tt(st.astype(np.float32)), tt(u), tt(mu), tt(sg), solver="exact",
                           gamma_b=20.0).numpy()
        out = O.cars_env_step(st, us.astype(np.float64), t, step)
        st, t, step = out["state"], out["t"], out["episode_step"]
    dev_st = env.state.cpu().numpy().astype(np.float64)
    assert keep.sum() > 0.8 * B
    err_pos = np.abs(dev_st[:, 0::2] - st[:, 0::2]).max(1)[keep]
    err_vel = np.abs(dev_st[:, 1::2] - st[:, 1::2]).max(1)[keep]
    # float32 positions near 200 m carry 1.5e-5 of rounding per step; velocities feed back through kp = 4, k_brake = 20
    assert np.median(err_pos) < 2e-3 and np.median(err_vel) < 2e-3, (np.median(err_pos), np.median(err_vel))
    assert (err_pos < 2e-2).mean() > 0.97 and (err_vel < 5e-2).mean() > 0.97
    stats = layer.solver_stats()
    assert stats["nan"] == 0 and stats["uncertified"] == 0


def test_fused_step_replays_from_a_cuda_graph():
    """The C entry points allocate nothing and only enqueue work on the caller's stream, so a rollout loop can be
    captured once and replayed (launch-bound small-batch loops: SURVEY 'CUDA streams and graphs').  Three replayed
    steps must equal three eager steps bit for bit."""
    import sac_rcbf_b200 as S
    B = 4096
    st, ac, mu, sg = O.synth_unicycle(B, seed=99)
    dev = lambda a: torch.from_numpy(a).cuda()  # noqa: E731
    layer_args = types.SimpleNamespace(cuda=True)
    outs = []
    for use_graph in (False, True):
        env = S.UnicycleEnv(num_envs=B)
        layer = S.CBFQPLayer(env, layer_args, gamma_b=20, k_d=3.0, l_p=0.03)
        env.reset()
        env.state = dev(st)
        a_, m_, s_ = dev(ac), dev(mu), dev(sg)
        if use_graph:
            side = torch.cuda.Stream()
            side.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(side):
                env.safe_step(layer, a_, m_, s_)              # warm-up on the capture stream (buffers, attributes)
            torch.cuda.current_stream().wait_stream(side)
            env.reset()
            env.state = dev(st)
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g):
                env.safe_step(layer, a_, m_, s_)
            env.reset()
            env.state = dev(st)
            for _ in range(3):
                g.replay()
        else:
            for _ in range(3):
                env.safe_step(layer, a_, m_, s_)
        torch.cuda.synchronize()
        outs.append((env.state.clone(), env._safe_action.clone(), env._obs.clone(), env._reward.clone()))
    for a, b in zip(*outs):
        assert torch.equal(a, b)
