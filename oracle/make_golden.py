"""TEST INFRASTRUCTURE ONLY -- generate tests/golden/*.npz from the UNMODIFIED reference source.

Run in the development container (needs /root/reference):

    python -m oracle.make_golden

What comes from where:
  * env trajectories, prior dynamics, obs<->state maps, prior disturbance, constraint assembly (P,q,G,h) and the
    CascadeCBFLayer assembly are outputs of the reference's own code (imported via oracle/ref_loader.py).
  * `safe_action` / `grad_action` are outputs of the reference's own CBFQPLayer.get_safe_action with
    ``qpth.qp.QPFunction`` bound to oracle/qpth_pdipm.py (qpth itself is not installed anywhere in this image, so
    the solve step is a restatement -- parity for it is unpinned, see that file's header).
  * `x_exact` is oracle/exact_qp.py on the reference-assembled, reference-normalised (G~, h~).
Seeds: 12345 (the reference's default --seed, main.py:223).
"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from oracle import exact_qp, ref_loader, rcbf_oracle as O  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden")
tt = torch.from_numpy


HAZARDS_7 = np.array([[0.0, 0.0], [-1.5, 1.5], [-1.5, -1.5], [1.5, -1.5], [1.5, 1.5], [0.0, 1.8], [1.9, 0.2]])
HAZARDS_11 = np.concatenate([HAZARDS_7, np.array([[-2.0, 0.1], [0.3, -2.1], [2.4, 2.2], [-0.8, 0.7]])], 0)


def layer_fixture(ref, mode, B, gamma_b, hazards=None):
    args = ref_loader.make_args()
    env = ref.UnicycleEnv() if mode == "Unicycle" else ref.SimulatedCarsEnv()
    if hazards is not None:                      # the layer sizes itself from len(env.hazards_locations), :35
        env.hazards_locations = np.array(hazards, dtype=np.float64)
    layer = ref.CBFQPLayer(env, args, gamma_b=gamma_b, k_d=3.0, l_p=0.03)
    if mode == "Unicycle":
        st, ac, mu, sg = O.synth_unicycle(B, seed=12345)
        if hazards is not None:                  # put the hazard-heavy stratum around THESE hazards
            rng = np.random.default_rng(99)
            nh = B // 3
            hz = np.asarray(hazards)[rng.integers(0, len(hazards), nh)]
            r, phi = rng.uniform(0.3, 1.1, nh), rng.uniform(-np.pi, np.pi, nh)
            st[:nh, 0] = hz[:, 0] + r * np.cos(phi)
            st[:nh, 1] = hz[:, 1] + r * np.sin(phi)
        t = np.zeros(B, np.float32)
    else:
        st, ac, mu, sg, t = O.synth_cars(B, seed=12345)
    P, q, G, h = layer.get_cbf_qp_constraints(tt(st), tt(ac), tt(mu), tt(sg))
    # reference normalisation lines replayed verbatim in spirit (diff_cbf_qp.py:103-106)
    Gh = torch.cat((G, h.unsqueeze(2)), -1)
    n = torch.max(torch.abs(Gh), dim=2, keepdim=True)[0]
    Gn, hn = G / n, h / n.squeeze(-1)
    a = tt(ac).clone().requires_grad_(True)
    final = layer.get_safe_action(tt(st), a, tt(mu), tt(sg))
    # a fixed, seeded upstream gradient
    w = torch.from_numpy(np.random.default_rng(777).normal(size=final.shape).astype(np.float32))
    (final * w).sum().backward()
    xe, lam, act, viol = exact_qp.solve_exact(P.double().numpy(), q.double().numpy(), Gn.double().numpy(),
                                              hn.double().numpy())
    # 1-D call contract (diff_cbf_qp.py:64-69,79)
    final_1d = layer.get_safe_action(tt(st[0]), tt(ac[0]), tt(mu[0]), tt(sg[0]))
    extra = {} if hazards is None else {"hazards": np.asarray(hazards, np.float64)}
    return dict(state=st, action=ac, mean=mu, sigma=sg, t=t, gamma_b=np.float32(gamma_b), **extra,
                P=P.numpy(), q=q.numpy(), G=G.numpy(), h=h.numpy(), Gn=Gn.numpy(), hn=hn.numpy(),
                safe_action=final.detach().numpy(), grad_w=w.numpy(), grad_action=a.grad.numpy(),
                x_exact=xe, lam_exact=lam, active_exact=act, viol_exact=viol, safe_action_1d=final_1d.numpy())


def unicycle_traj(ref, T=900):
    env = ref.UnicycleEnv()
    rng = np.random.default_rng(12345)
    obs0 = env.reset()
    acts = rng.uniform(-1.5, 1.5, (T, 2))      # beyond [-1,1] on purpose: exercises the clip (unicycle_env.py:62)
    # steer through a hazard and to the goal so cost / goal_met / done all occur
    obs, rew, done, cost, goal, states = [], [], [], [], [], []
    for k in range(T):
        a = acts[k].copy()
        if k >= 50:                            # P-controller toward the goal, like the reference's demo controller
            o = obs[-1]
            a = np.array([1.0, 2.0 * np.arctan2(o[5], o[4])]) + 0.1 * acts[k]
            acts[k] = a
        o, r, d, info = env.step(a)
        obs.append(o.copy()); rew.append(r); done.append(d); cost.append(info.get("cost", 0.0))
        goal.append(bool(info.get("goal_met", False))); states.append(env.state.copy())
        if d:
            acts = acts[:k + 1]
            break
    return dict(obs0=obs0, actions=acts, obs=np.array(obs), reward=np.array(rew), done=np.array(done),
                cost=np.array(cost), goal_met=np.array(goal), state=np.array(states))


def cars_traj(ref, T=300):
    np.random.seed(0)
    env = ref.SimulatedCarsEnv()
    np.random.seed(0)
    obs0 = env.reset()
    v_noise = env.state[1] - 30.0
    rng = np.random.default_rng(12345)
    acts = rng.uniform(-1.0, 1.0, (T, 1)) * np.linspace(0.2, 3.0, T)[:, None]
    obs, rew, done, cost, states, ts = [], [], [], [], [], []
    for k in range(T):
        o, r, d, info = env.step(acts[k].copy())
        obs.append(o.copy()); rew.append(r); done.append(d); cost.append(info["cost"]); states.append(env.state.copy())
        ts.append(env.t)
    return dict(obs0=obs0, v_noise=np.float64(v_noise), actions=acts, obs=np.array(obs), reward=np.array(rew),
                done=np.array(done), cost=np.array(cost), state=np.array(states), t=np.array(ts))


def dynamics_fixture(ref, B=64):
    args = ref_loader.make_args()
    out = {}
    for mode, envc in (("Unicycle", ref.UnicycleEnv), ("SimulatedCars", ref.SimulatedCarsEnv)):
        env = envc()
        dm = ref.DynamicsModel(env, args)
        if mode == "Unicycle":
            st, ac, _, _ = O.synth_unicycle(B, seed=4242)
            t = None
        else:
            st, ac, _, _, t = O.synth_cars(B, seed=4242)
            t = t.astype(np.float64)
        st = st.astype(np.float64); ac = ac.astype(np.float64)
        nxt, std, tn = dm.predict_next_state(st, ac, t_batch=t)
        obs = dm.get_obs(st)
        st_back = dm.get_state(obs)
        mean, fstd = dm.predict_disturbance(st)
        k = mode.lower()
        out.update({k + "_state": st, k + "_action": ac, k + "_next": nxt, k + "_std_dt": std, k + "_obs": obs,
                    k + "_state_from_obs": st_back, k + "_dist_mean": mean, k + "_dist_std": fstd})
        if t is not None:
            out[k + "_t"] = t
            out[k + "_t_next"] = tn
    return out


def cascade_fixture(ref, B=16):
    env = ref.UnicycleEnv()
    lay = ref.CascadeCBFLayer(env, gamma_b=100, k_d=1.5, l_p=0.03)
    st, ac, mu, sg = (a.astype(np.float64) for a in O.synth_unicycle(B, seed=99))
    Gs, hs, Ps = [], [], []
    for i in range(B):
        P, q, G, h = lay.get_cbf_qp_constraints(ac[i], st[i], mu[i], sg[i])
        Gs.append(G); hs.append(h); Ps.append(P)
    envc = ref.SimulatedCarsEnv()
    layc = ref.CascadeCBFLayer(envc, gamma_b=100, k_d=1.5)
    stc, acc, muc, sgc, _ = (a.astype(np.float64) for a in O.synth_cars(B, seed=99))
    Gc, hc = [], []
    for i in range(B):
        P, q, G, h = layc.get_cbf_qp_constraints(acc[i], stc[i], muc[i], sgc[i])
        Gc.append(G); hc.append(np.asarray(h, np.float64))
    return dict(state=st, action=ac, mean=mu, sigma=sg, G=np.array(Gs), h=np.array(hs), P=np.array(Ps),
                cars_state=stc, cars_action=acc, cars_sigma=sgc, cars_G=np.array(Gc), cars_h=np.array(hc))


def rollouts_fixture(ref, B=25):
    """The reference's own generate_model_rollouts (rcbf_sac/generate_rollouts.py) driven with a stub agent / memory and
    a recorded Gaussian draw; what it pushes to memory_model is the golden output."""
    out = {}
    args = ref_loader.make_args()
    for mode, envc in (("Unicycle", ref.UnicycleEnv), ("SimulatedCars", ref.SimulatedCarsEnv)):
        env = envc()
        dm = ref.DynamicsModel(env, args)
        rng = np.random.default_rng(4321)
        if mode == "Unicycle":
            st, ac, _, _ = O.synth_unicycle(B, seed=4321)
            st = st.astype(np.float64)
            st[:3, :2] = env.goal_pos + np.array([[0.05, -0.1], [0.2, 0.1], [-0.15, 0.05]])   # some reach the goal
            obs = O.unicycle_obs(st)
            t = np.arange(B) * 0.02
            n_s, n_u = 3, 2
        else:
            st, ac, _, _, t = O.synth_cars(B, seed=4321)
            st = st.astype(np.float64)
            obs = O.cars_obs(st)
            t = t.astype(np.float64)
            t[:4] = 300 * 0.02 - 0.02 * np.array([0.5, 1.0, 1.5, 3.0])     # straddle the time horizon
            n_s, n_u = 10, 1
        actions = rng.uniform(-1, 1, (B, n_u)) * (2.5 if mode == "Unicycle" else 10.0)
        eps = rng.normal(size=(B, n_s))

        class Mem:
            def __init__(self):
                self.items = None

            def sample(self, batch_size):
                return obs.copy(), None, None, None, None, t.copy(), None

            def batch_push(self, *a):
                self.items = [np.array(x, copy=True) for x in a]

        class Agent:
            def select_action(self, o, dmodel, evaluate=False, warmup=False):
                return actions.copy()

        mm = Mem()
        orig = np.random.normal
        np.random.normal = lambda mu, sd, *a, **k: mu + sd * eps        # generate_rollouts.py:31
        try:
            ref.generate_model_rollouts(env, mm, Mem(), Agent(), dm, k_horizon=1, batch_size=B)
        finally:
            np.random.normal = orig
        k = mode.lower()
        names = ("obs", "action", "reward", "next_obs", "mask", "t", "next_t")
        for nme, val in zip(names, mm.items):
            out[k + "_" + nme] = np.asarray(val, np.float64) if val.dtype != bool else val
        out[k + "_eps"] = eps
    return out


def gp_fixture(ref, n_uni=120, n_cars=96, n_test=40):
    """Disturbance history = output of the REFERENCE's own envs + DynamicsModel.append_transition (random actions,
    GP refits disabled: gpytorch is absent); hyper-parameters / posterior = oracle/gp_oracle.py (parity unpinned)."""
    from oracle import gp_oracle
    args = ref_loader.make_args()
    args.gp_model_size = 100000          # never reaches the refit trigger (dynamics.py:303)
    out = {}
    rng = np.random.default_rng(777)
    for mode, envc, n in (("Unicycle", ref.UnicycleEnv, n_uni), ("SimulatedCars", ref.SimulatedCarsEnv, n_cars)):
        np.random.seed(12345)
        env = envc()
        dm = ref.DynamicsModel(env, args)
        env.reset()
        for _ in range(n):
            state = np.copy(env.state if mode == "Unicycle" else env.state)
            t = None if mode == "Unicycle" else np.array([env.t])
            a = rng.uniform(-1, 1, env.action_space.shape[0])
            _, _, done, _ = env.step(a)
            if mode == "Unicycle":
                a_used = np.clip(a, -1.0, 1.0)   # unicycle_env.py:62 clips before stepping
                dm.append_transition(state, a_used, np.copy(env.state))
            else:
                dm.append_transition(state, a, np.copy(env.state), t_batch=t)
            if done:
                env.reset()
        train_x = dm.disturbance_history['state'][:dm.history_counter].copy()
        train_y = dm.disturbance_history['disturbance'][:dm.history_counter].copy()
        gps = gp_oracle.DisturbanceGPs(train_x, train_y, ref.MAX_STD[mode], training_iter=70)
        lo, hi = train_x.min(0), train_x.max(0)
        test_x = rng.uniform(lo - 0.2 * (hi - lo + 1e-3), hi + 0.2 * (hi - lo + 1e-3), (n_test, train_x.shape[1]))
        mean, std = gps.predict_disturbance(test_x)
        k = mode.lower()
        out.update({k + "_train_x": train_x, k + "_train_y": train_y, k + "_raw": np.stack([g.raw for g in gps.gps]),
                    k + "_test_x": test_x, k + "_mean": mean, k + "_std": std})
    return out


REPLAY_SCHEDULE = (5, 1, 20, 30, 3, 80, 0, 11)   # wraps, exact fill, a batch larger than the capacity, an empty batch


def replay_inputs(cap=37, od=7, ad=2, seed=0):
    """The push schedule of the replay fixture (regenerated identically by the tests)."""
    rng = np.random.default_rng(seed)
    out = []
    for n in REPLAY_SCHEDULE:
        s, a, r = rng.normal(size=(n, od)), rng.normal(size=(n, ad)), rng.normal(size=n)
        s2, m, t = rng.normal(size=(n, od)), (rng.random(n) > 0.2).astype(np.float64), rng.random(n)
        out.append((s, a, r, s2, m, t, t + 0.02))
    return out


def replay_fixture(ref, cap=37):
    """rcbf_sac/replay_memory.py:4-35 driven through batch_push / push; the ring content after every batch."""
    mem = ref.ReplayMemory(cap, 0)
    out = {"capacity": np.int64(cap)}
    for k, b in enumerate(replay_inputs(cap)):
        mem.batch_push(*b)
        items = [np.concatenate([np.ravel(x) for x in it]) for it in mem.buffer]
        out["after_%d" % k] = np.stack(items) if items else np.zeros((0, 20))
        out["position_%d" % k] = np.int64(mem.position)
    mem.push(np.ones(7), np.ones(2), 1.0, np.ones(7), 1.0, 0.5, 0.52)
    out["after_push"] = np.stack([np.concatenate([np.ravel(x) for x in it]) for it in mem.buffer])
    out["position_push"] = np.int64(mem.position)
    smp = mem.sample(16)                                    # shapes / dtypes of what sample returns (:28-33)
    out["sample_shapes"] = np.array([np.asarray(x).ndim for x in smp])
    return out


def main():
    ref = ref_loader.load_reference()
    if "--hazards-only" in sys.argv:
        np.savez_compressed(os.path.join(OUT, "unicycle_layer_7haz_b256.npz"), **layer_fixture(ref, "Unicycle", 256, 20.0, HAZARDS_7))
        np.savez_compressed(os.path.join(OUT, "unicycle_layer_11haz_b128.npz"), **layer_fixture(ref, "Unicycle", 128, 20.0, HAZARDS_11))
        return
    if "--replay-only" in sys.argv:
        np.savez_compressed(os.path.join(OUT, "replay_memory.npz"), **replay_fixture(ref))
        return
    if "--gp-only" in sys.argv:
        np.savez_compressed(os.path.join(OUT, "gp_disturbance.npz"), **gp_fixture(ref))
        return
    os.makedirs(OUT, exist_ok=True)
    torch.manual_seed(12345)
    np.savez_compressed(os.path.join(OUT, "unicycle_layer_b256.npz"), **layer_fixture(ref, "Unicycle", 256, 20.0))
    np.savez_compressed(os.path.join(OUT, "cars_layer_b512.npz"), **layer_fixture(ref, "SimulatedCars", 512, 20.0))
    torch.manual_seed(12345)
    np.savez_compressed(os.path.join(OUT, "unicycle_layer_7haz_b256.npz"), **layer_fixture(ref, "Unicycle", 256, 20.0, HAZARDS_7))
    np.savez_compressed(os.path.join(OUT, "unicycle_layer_11haz_b128.npz"), **layer_fixture(ref, "Unicycle", 128, 20.0, HAZARDS_11))
    np.savez_compressed(os.path.join(OUT, "unicycle_env_traj.npz"), **unicycle_traj(ref))
    np.savez_compressed(os.path.join(OUT, "cars_env_traj.npz"), **cars_traj(ref))
    np.savez_compressed(os.path.join(OUT, "dynamics_prior.npz"), **dynamics_fixture(ref))
    np.savez_compressed(os.path.join(OUT, "cascade_layer.npz"), **cascade_fixture(ref))
    np.savez_compressed(os.path.join(OUT, "model_rollouts.npz"), **rollouts_fixture(ref))
    np.savez_compressed(os.path.join(OUT, "gp_disturbance.npz"), **gp_fixture(ref))
    np.savez_compressed(os.path.join(OUT, "replay_memory.npz"), **replay_fixture(ref))
    for f in sorted(os.listdir(OUT)):
        print(f, os.path.getsize(os.path.join(OUT, f)))


if __name__ == "__main__":
    main()
