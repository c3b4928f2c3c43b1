"""generate_model_rollouts -- drop-in for rcbf_sac/generate_rollouts.py:6-81 with the per-transition arithmetic
(get_state, prior step + GP mean, Gaussian sample, observation rebuild, reward, done) in ONE kernel per horizon step
(`rcbf_*_rollout_step_*`), SURVEY.md 8f row 2.

Works with the reference's host-side ReplayMemory (numpy in / numpy out, float64 kernel = the reference's numpy
arithmetic) and with DeviceReplayMemory (everything stays on the GPU, float32 kernel).  The Gaussian draw is made with
the caller's generator (numpy global RNG on the host path like generate_rollouts.py:31, torch generator on the device
path) and handed to the kernel as standard-normal `eps`, so a transition is a deterministic function of its inputs.

Reference quirk kept: `reward_goal` is added twice when the goal is reached (:50,:53).
Deliberate deviation for k_horizon > 1 (identical for the reference's k_horizon = 1 call, main.py:53-57): the reference
hands `predict_next_state` the ORIGINAL `t_batch` at every horizon step (:30) and drops done rows from the observation
batch only (:78-79), so after the first dropped row its time and observation batches no longer have the same length;
here the time batch is rolled (t <- next_t) and shrunk together with the observations, which is what the pushed
`t_batch_` / `next_t_batch_` of :71-72 intend.  A memory that stores no times (t = None, Unicycle) is passed through.
"""
from copy import deepcopy

import numpy as np
import torch

from . import _lib


def rollout_transition(env, dynamics_model, obs, action, t, eps):
    """One model transition for a batch.  obs (B,n_o), action (B,n_u), t (B,), eps (B,n_s) standard normal (or None for
    the mean).  ndarray in -> ndarray out (float64 kernel); device tensors in -> device tensors out (their dtype).
    Returns next_obs, reward, done, next_t."""
    lib = _lib.load()
    host = not torch.is_tensor(obs)
    dev = dynamics_model.device
    dt_ = torch.float64 if host else (obs.dtype if obs.dtype in (torch.float32, torch.float64) else torch.float32)
    suf = "f64" if dt_ == torch.float64 else "f32"
    T = lambda x, shape: None if x is None else torch.as_tensor(  # noqa: E731
        np.asarray(x, np.float64) if not torch.is_tensor(x) else x).to(dev, dt_).reshape(shape).contiguous()
    B = int(obs.shape[0])
    mode = env.dynamics_mode
    n_s = dynamics_model.n_s
    o = T(obs, (B, -1))
    a = T(action, (B, -1))
    state = dynamics_model.get_state(o)
    mean, std = dynamics_model.predict_disturbance(state)       # GP (or prior) mean / std at the current state
    mean, std = T(mean, (B, n_s)), T(std, (B, n_s))
    e = T(eps, (B, n_s))
    next_obs = torch.empty_like(o)
    reward = torch.empty((B,), dtype=dt_, device=dev)
    done = torch.empty((B,), dtype=torch.uint8, device=dev)
    with torch.cuda.device(dev):
        if mode == 'Unicycle':
            gp = env.unwrapped.goal_pos
            rc = getattr(lib, "rcbf_unicycle_rollout_step_" + suf)(
                _lib.ptr(o), _lib.ptr(a), _lib.ptr(mean), _lib.ptr(std), _lib.ptr(e), B, float(env.dt), float(gp[0]),
                float(gp[1]), _lib.ptr(next_obs), _lib.ptr(reward), _lib.ptr(done), _lib.stream_ptr(dev))
            next_t = None if t is None else t + env.dt
        elif mode == 'SimulatedCars':
            tt = T(t, (B,))
            next_t = torch.empty_like(tt)
            rc = getattr(lib, "rcbf_cars_rollout_step_" + suf)(
                _lib.ptr(o), _lib.ptr(a), _lib.ptr(tt), _lib.ptr(mean), _lib.ptr(std), _lib.ptr(e), B, float(env.dt),
                float(getattr(env, 'kp', 4.0)), float(getattr(env, 'k_brake', 20.0)), int(env.max_episode_steps),
                _lib.ptr(next_obs), _lib.ptr(reward), _lib.ptr(done), _lib.ptr(next_t), _lib.stream_ptr(dev))
        else:
            raise Exception('Environment/Dynamics mode {} not Recognized!'.format(mode))
    _lib.check(rc, "rcbf_rollout_step")
    done = done.bool()
    if host:
        return (next_obs.cpu().numpy(), reward.cpu().numpy(), done.cpu().numpy(),
                next_t if (next_t is None or not torch.is_tensor(next_t)) else next_t.cpu().numpy())
    return next_obs, reward, done, next_t


def generate_model_rollouts(env, memory_model, memory, agent, dynamics_model, k_horizon=1, batch_size=20,
                            warmup=False, generator=None):
    """Same signature and side effects as the reference (fills `memory_model`, returns it)."""

    def policy(observation):
        if warmup and env.action_space:
            return agent.select_action(observation, dynamics_model, warmup=True)
        return agent.select_action(observation, dynamics_model, evaluate=False)

    obs_batch, action_batch, reward_batch, next_obs_batch, mask_batch, t_batch, next_t_batch = \
        memory.sample(batch_size=batch_size)
    on_device = torch.is_tensor(obs_batch)
    obs_batch_ = obs_batch.clone() if on_device else deepcopy(obs_batch)
    t_batch_ = None if t_batch is None else (t_batch.clone() if on_device and torch.is_tensor(t_batch) else deepcopy(t_batch))
    n_s = dynamics_model.n_s

    for k in range(k_horizon):
        B = obs_batch_.shape[0]
        if B == 0:
            break
        action_batch_ = policy(obs_batch_ if on_device else np.asarray(obs_batch_))
        if on_device:
            action_batch_ = torch.as_tensor(action_batch_).to(obs_batch_.device, obs_batch_.dtype)
            eps = torch.randn((B, n_s), generator=generator, device=obs_batch_.device, dtype=obs_batch_.dtype)
        else:
            eps = np.random.normal(0.0, 1.0, (B, n_s))        # global numpy RNG like generate_rollouts.py:31
        next_obs_batch_, reward_batch_, done_batch_, next_t_batch_ = rollout_transition(
            env, dynamics_model, obs_batch_, action_batch_, t_batch_, eps)
        mask_batch_ = ~done_batch_ if on_device else np.invert(done_batch_)
        memory_model.batch_push(obs_batch_, action_batch_, reward_batch_, next_obs_batch_, mask_batch_, t_batch_,
                                next_t_batch_)
        keep = ~done_batch_
        obs_batch_ = next_obs_batch_[keep]
        t_batch_ = None if next_t_batch_ is None else next_t_batch_[keep]
    return memory_model
