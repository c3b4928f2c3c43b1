"""Caller contract: the three reference call sites of the path, restated here with OUR classes
plugged in -- RCBF_SAC.get_safe_action (rcbf_sac/sac_cbf.py:218-238), select_action (:59-90) and
generate_model_rollouts (rcbf_sac/generate_rollouts.py:6-81).  The reference source is not available on the GPU box,
so the few lines of each caller are restated next to their file:line; what is under test is that our objects accept
exactly the calls those lines make (names, argument order, array kinds, shapes, dtypes) and produce the oracle's
numbers."""
import types

import numpy as np
import pytest
import torch

from oracle import rcbf_oracle as O

pytestmark = pytest.mark.gpu


class _Agent:
    """The safety-relevant part of RCBF_SAC with a fixed 'policy' (the RL nets are out of scope)."""

    def __init__(self, env, args, S):
        self.device = torch.device("cuda")
        self.action_space = env.action_space
        self.cbf_layer = S.CBFQPLayer(env, args, args.gamma_b, args.k_d, args.l_p)      # sac_cbf.py:50
        self.rng = np.random.default_rng(0)

    def get_safe_action(self, obs_batch, action_batch, dynamics_model):                  # sac_cbf.py:218-238
        state_batch = dynamics_model.get_state(obs_batch)
        mean_pred_batch, sigma_pred_batch = dynamics_model.predict_disturbance(state_batch)
        return self.cbf_layer.get_safe_action(state_batch, action_batch, mean_pred_batch, sigma_pred_batch)

    def select_action(self, state, dynamics_model, evaluate=False, warmup=False):        # sac_cbf.py:59-90
        state = torch.from_numpy(state).type(torch.FloatTensor).to(self.device)
        expand_dim = len(state.shape) == 1
        if expand_dim:
            state = state.unsqueeze(0)
        batch_size = state.shape[0]
        action = torch.zeros((batch_size, self.action_space.shape[0])).to(self.device)
        for i in range(batch_size):
            action[i] = torch.from_numpy(self.action_space.sample()).to(self.device)
        self.last_nominal = action.clone()
        safe_action = self.get_safe_action(state, action, dynamics_model)
        return safe_action.detach().cpu().numpy()[0] if expand_dim else safe_action.detach().cpu().numpy()


def _args():
    return types.SimpleNamespace(cuda=True, gp_model_size=2000, l_p=0.03, gamma_b=20, k_d=3.0)


@pytest.fixture(scope="module")
def S():
    import sac_rcbf_b200 as S_
    return S_


@pytest.mark.parametrize("env_name", ["Unicycle", "SimulatedCars"])
def test_training_loop_step_contract(S, env_name):
    """main.py:47-113 for a few steps: obs -> get_state -> select_action -> env.step -> append_transition."""
    args = _args()
    args.env_name = env_name
    env = S.build_env(args)
    env.seed(0)
    agent = _Agent(env, args, S)
    dm = S.DynamicsModel(env, args)
    obs = env.reset()
    mode = env.dynamics_mode
    for k in range(20):
        state = dm.get_state(obs)                                                        # main.py:50
        assert isinstance(state, np.ndarray) and state.shape == (dm.n_s,)
        action = agent.select_action(obs, dm, warmup=True)                               # main.py:93
        assert isinstance(action, np.ndarray) and action.shape == env.action_space.shape
        # oracle for the safe action of this very state / nominal action
        st = torch.from_numpy(state[None]).float()
        mu, sg = O.prior_disturbance(mode, 1)
        ref = O.safe_action(mode, st, agent.last_nominal.cpu(), torch.from_numpy(mu).float(),
                            torch.from_numpy(sg).float(), solver="exact", gamma_b=20.0)
        assert np.abs(action - ref.numpy()[0]).max() < 1e-4
        obs2, reward, done, info = env.step(action)                                      # main.py:95
        assert isinstance(obs2, np.ndarray) and isinstance(reward, float) and isinstance(done, bool)
        next_state = dm.get_state(obs2)
        t = None if mode == "Unicycle" else np.array([env.t - env.dt])
        dm.append_transition(state, action, next_state, t_batch=t)                       # main.py:111-113
        obs = obs2
    assert dm.history_counter == 20
    # the disturbance the prior cannot explain is what the reference's GP would learn: Unicycle drag term
    d = dm.disturbance_history["disturbance"][:20]
    assert np.isfinite(d).all()
    if mode == "Unicycle":
        assert np.abs(d[:, 2]).max() < 1e-9            # heading has no unmodelled term (unicycle_env.py:86-87)


@pytest.mark.parametrize("env_name", ["Unicycle", "SimulatedCars"])
def test_generate_model_rollouts_contract(S, env_name):
    """generate_rollouts.py:17-66 with batch 25 (main.py:56,257) against the oracle's numbers."""
    args = _args()
    args.env_name = env_name
    env = S.build_env(args)
    agent = _Agent(env, args, S)
    dm = S.DynamicsModel(env, args)
    mode = env.dynamics_mode
    B = 25
    rng = np.random.default_rng(1)
    if mode == "Unicycle":
        st, _, _, _ = O.synth_unicycle(B, seed=21)
        st = st.astype(np.float64)
        obs_batch = O.unicycle_obs(st)
        t_batch = np.zeros(B)
    else:
        st, _, _, _, t = O.synth_cars(B, seed=21)
        st = st.astype(np.float64)
        obs_batch = O.cars_obs(st)
        t_batch = t.astype(np.float64)
    action_batch_ = agent.select_action(obs_batch, dm, warmup=True)                      # :28
    assert action_batch_.shape == (B, env.action_space.shape[0])
    state_batch_ = dm.get_state(obs_batch)                                               # :29
    nxt_mu, nxt_std, nxt_t = dm.predict_next_state(state_batch_, action_batch_, t_batch=t_batch)   # :30
    ref_mu, ref_std, ref_t = O.predict_next_state(mode, state_batch_, action_batch_.astype(np.float64), t_batch)
    np.testing.assert_allclose(nxt_mu, ref_mu, rtol=1e-12, atol=1e-11)
    np.testing.assert_allclose(nxt_std, ref_std, atol=1e-15)
    np.testing.assert_allclose(nxt_t, ref_t)
    next_state_batch_ = rng.normal(nxt_mu, nxt_std)                                      # :31
    next_obs_batch_ = dm.get_obs(next_state_batch_)                                      # :32
    np.testing.assert_allclose(next_obs_batch_, O.get_obs(mode, next_state_batch_), atol=1e-14)
    if mode == "Unicycle":                                                               # :34-55
        goal_rel = env.unwrapped.goal_pos[:2] - next_obs_batch_[:, :2]
        dist2goal = np.linalg.norm(goal_rel, axis=1)
        assert dist2goal.shape == (B,)
        assert next_obs_batch_.shape == (B, 4)
    else:                                                                                # :57-66
        reward = -5.0 * np.abs(action_batch_.squeeze() ** 2) / env.max_episode_steps
        done = nxt_t >= env.max_episode_steps * env.dt
        assert reward.shape == (B,) and done.shape == (B,)


# ----------------------------------------------------------------------------------------------------- next rows (8f)
def test_rollout_transition_vs_golden_and_oracle(S, golden):
    """the fused rollout-transition kernel vs the reference's own generate_model_rollouts output
    (float64 host path) and vs the oracle at float32 on the device."""
    g = golden("model_rollouts.npz")
    for env_name, k in (("Unicycle", "unicycle"), ("SimulatedCars", "simulatedcars")):
        args = _args(); args.env_name = env_name
        env = S.build_env(args)
        dm = S.DynamicsModel(env, args)
        nobs, rew, done, nt = S.rollout_transition(env, dm, g[k + "_obs"], g[k + "_action"], g[k + "_t"], g[k + "_eps"])
        assert isinstance(nobs, np.ndarray) and nobs.dtype == np.float64
        np.testing.assert_allclose(nobs, g[k + "_next_obs"], rtol=1e-12, atol=1e-12)
        np.testing.assert_allclose(rew, g[k + "_reward"], rtol=1e-12, atol=1e-12)
        np.testing.assert_array_equal(~done, g[k + "_mask"].astype(bool))
        np.testing.assert_allclose(nt, g[k + "_next_t"], atol=1e-14)
        # float32, device resident, 200k instances vs the oracle
        B = 200000
        if env_name == "Unicycle":
            st, ac, _, _ = O.synth_unicycle(B, seed=6); obs = O.unicycle_obs(st.astype(np.float64)); t = np.zeros(B)
        else:
            st, ac, _, _, t = O.synth_cars(B, seed=6); obs = O.cars_obs(st.astype(np.float64)); t = t.astype(np.float64)
        eps = np.random.default_rng(2).normal(size=(B, dm.n_s))
        ro, rr, rd, rt = O.rollout_step(env_name, obs, ac.astype(np.float64), t, eps)
        dev = lambda a: torch.from_numpy(np.ascontiguousarray(a)).float().cuda()  # noqa: E731
        o32, r32, d32, t32 = S.rollout_transition(env, dm, dev(obs), dev(ac), dev(t), dev(eps))
        assert o32.is_cuda and o32.dtype == torch.float32
        keep = np.ones(B, bool) if env_name == "Unicycle" else O.cars_threshold_margin(st) > 1e-3
        err = np.abs(o32.cpu().numpy() - ro) / np.maximum(np.abs(ro), 1.0)
        if env_name == "Unicycle":      # compass ~ 1/dist conditioning, like the env observation
            dist = -np.log(ro[:, 6])
            err[:, 4:6] *= np.minimum(dist, 1.0)[:, None]
        assert err[keep].max() < 3e-6, err[keep].max()
        assert np.abs(r32.cpu().numpy() - rr)[keep].max() < 3e-6
        near = np.abs(-np.log(np.maximum(ro[:, 6], 1e-30)) - 0.3) < 1e-5 if env_name == "Unicycle" else np.zeros(B, bool)
        assert ((d32.cpu().numpy() == rd) | near).all()


def test_generate_model_rollouts_host_and_device_memories(S):
    """Same call as main.py:53-57, once with a host-side reference-style memory and once with DeviceReplayMemory."""
    args = _args(); args.env_name = "Unicycle"
    env = S.build_env(args)
    agent = _Agent(env, args, S)
    dm = S.DynamicsModel(env, args)
    B = 25
    st, _, _, _ = O.synth_unicycle(200, seed=33)
    obs = O.unicycle_obs(st.astype(np.float64))

    class HostMem:                               # rcbf_sac/replay_memory.py shape of things
        def __init__(self):
            self.rows = []

        def sample(self, batch_size):
            idx = np.arange(batch_size)
            return obs[idx], None, None, None, None, np.zeros(batch_size), None

        def batch_push(self, s, a, r, s2, m, t, t2):
            for i in range(s.shape[0]):
                self.rows.append((s[i], a[i], r[i], s2[i], m[i], t[i], t2[i]))

        def __len__(self):
            return len(self.rows)

    np.random.seed(3)
    mm = S.generate_model_rollouts(env, HostMem(), HostMem(), agent, dm, k_horizon=2, batch_size=B, warmup=True)
    assert B <= len(mm) <= 2 * B
    s, a, r, s2, m, t, t2 = mm.rows[0]
    assert s.shape == (7,) and a.shape == (2,) and s2.shape == (7,) and abs(t2 - t - env.dt) < 1e-12
    # device memories
    mem = S.DeviceReplayMemory(1000, seed=0, obs_dim=7, action_dim=2)
    x = torch.from_numpy(obs).float().cuda()
    mem.batch_push(x, torch.zeros(200, 2), torch.zeros(200), x, torch.ones(200), torch.zeros(200), torch.zeros(200))
    mem_model = S.DeviceReplayMemory(1000, seed=1, obs_dim=7, action_dim=2)

    class DevAgent(_Agent):
        def select_action(self, state, dynamics_model, evaluate=False, warmup=False):
            if torch.is_tensor(state):
                action = 2 * torch.rand((state.shape[0], 2), device=state.device) - 1
                return self.get_safe_action(state, action, dynamics_model)
            return super().select_action(state, dynamics_model, evaluate, warmup)

    out = S.generate_model_rollouts(env, mem_model, mem, DevAgent(env, args, S), dm, k_horizon=1, batch_size=B)
    assert out is mem_model and len(mem_model) == B
    s, a, r, s2, m, t, t2 = mem_model.sample(B)
    assert s.is_cuda and s2.shape == (B, 7) and torch.isfinite(s2).all() and torch.isfinite(r).all()
    assert torch.allclose(t2, t + env.dt)
