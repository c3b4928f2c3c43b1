"""CPU tests that pin the oracle (oracle/*.py) before anything is compared with it.

1. oracle vs the committed golden vectors (generated from the unmodified reference source by
   oracle/make_golden.py) -- runs everywhere;
2. oracle vs the live reference source (skipped when /root/reference is absent, e.g. on the GPU box);
3. the restated qpth PDIPM vs the exact active-set solver (the QP solve is unpinned by the reference).
"""
import numpy as np
import pytest
import torch

from oracle import exact_qp, qpth_pdipm, ref_loader, rcbf_oracle as O

tt = torch.from_numpy


# ----------------------------------------------------------------------------- survey KATs
def test_kat_unicycle_reset_and_step():
    r = O.unicycle_reset(1)
    np.testing.assert_allclose(r["obs"][0], [-2.5, -2.5, 1, 0, 0.70700679533, 0.70700679533, 8.4932570472e-4],
                               rtol=0, atol=1e-10)
    out = O.unicycle_env_step(r["state"], np.array([[1.0, 0.5]]), r["episode_step"], r["last_goal_dist"])
    np.testing.assert_allclose(out["state"][0], [-2.4819998, -2.5000199987, 0.01], atol=1e-9)
    assert abs(out["reward"][0] - 0.012702420761572242) < 1e-14
    assert not out["done"][0] and out["cost"][0] == 0.0


def test_kat_unicycle_prior():
    nxt, std, _ = O.predict_next_state("Unicycle", np.array([[-2.5, -2.5, 0.0]]), np.array([[1.0, 0.5]]))
    np.testing.assert_allclose(nxt[0], [-2.48, -2.5, 0.01], atol=1e-14)
    np.testing.assert_allclose(std[0], [0.004] * 3, atol=1e-15)


def test_kat_assembly():
    st = torch.tensor([[-2.5, -2.5, 0.0]]); ac = torch.tensor([[1.0, 0.5]])
    mu = torch.zeros(1, 3); sg = torch.full((1, 3), 0.2)
    P, q, G, h = O.assemble_unicycle(st, ac, mu, sg, gamma_b=20.0, l_p=0.03)
    np.testing.assert_allclose(G[0, :5].numpy(), [[2.47, .075, -1], [.97, .12, -1], [.97, .03, -1], [3.97, .03, -1],
                                                  [3.97, .12, -1]], atol=1e-6)
    np.testing.assert_allclose(h[0].numpy(), [4138.09766, 11070.7627, 5.81109619, 10707.7080, 76233.3828, 1.5, 3.5,
                                              2.0, 3.0], rtol=2e-7)
    np.random.seed(0)
    v = np.random.normal(0, 0.5)
    rs = O.cars_reset(1, [v])
    P, q, G, h = O.assemble_cars(tt(rs["state"]).float(), torch.tensor([[0.3]]), torch.zeros(1, 10),
                                 tt(np.array([O.MAX_STD["SimulatedCars"]])).float(), gamma_b=20.0)
    np.testing.assert_allclose(G[0].numpy(), [[300, -200], [-300, -200], [1, 0], [-1, 0]], atol=1e-4)
    np.testing.assert_allclose(h[0].numpy(), [3665.07543945, 7304.04003906, 9.7, 10.3], rtol=3e-7)


# ----------------------------------------------------------------------------- oracle vs golden fixtures
def test_env_unicycle_vs_golden(golden):
    g = golden("unicycle_env_traj.npz")
    r = O.unicycle_reset(1)
    np.testing.assert_allclose(r["obs"][0], g["obs0"], atol=1e-15)
    st, step, last = r["state"], r["episode_step"], r["last_goal_dist"]
    for k in range(len(g["reward"])):
        out = O.unicycle_env_step(st, g["actions"][k:k + 1], step, last)
        np.testing.assert_allclose(out["state"][0], g["state"][k], atol=1e-12)
        np.testing.assert_allclose(out["obs"][0], g["obs"][k], atol=1e-12)
        assert abs(out["reward"][0] - g["reward"][k]) < 1e-12
        assert bool(out["done"][0]) == bool(g["done"][k]) and bool(out["goal_met"][0]) == bool(g["goal_met"][k])
        assert abs(out["cost"][0] - g["cost"][k]) < 1e-15
        st, step, last = out["state"], out["episode_step"], out["last_goal_dist"]
    assert g["done"][-1] and g["goal_met"][-1] and (g["cost"] > 0).any()


def test_env_cars_vs_golden(golden):
    g = golden("cars_env_traj.npz")
    r = O.cars_reset(1, [float(g["v_noise"])])
    np.testing.assert_allclose(r["obs"][0], g["obs0"], atol=1e-15)
    st, t, step = r["state"], r["t"], r["episode_step"]
    for k in range(len(g["reward"])):
        out = O.cars_env_step(st, g["actions"][k:k + 1], t, step)
        np.testing.assert_allclose(out["state"][0], g["state"][k], rtol=1e-13, atol=1e-11)
        np.testing.assert_allclose(out["obs"][0], g["obs"][k], rtol=1e-13, atol=1e-12)
        assert abs(out["reward"][0] - g["reward"][k]) < 1e-15 and abs(out["cost"][0] - g["cost"][k]) < 1e-15
        assert bool(out["done"][0]) == bool(g["done"][k])
        st, t, step = out["state"], out["t"], out["episode_step"]
    assert g["done"][-1] and (g["cost"] != 0).any()


def test_dynamics_vs_golden(golden):
    g = golden("dynamics_prior.npz")
    for mode, k in (("Unicycle", "unicycle"), ("SimulatedCars", "simulatedcars")):
        t = g.get(k + "_t")
        nxt, std, tn = O.predict_next_state(mode, g[k + "_state"], g[k + "_action"], t)
        np.testing.assert_allclose(nxt, g[k + "_next"], rtol=1e-14, atol=1e-12)
        np.testing.assert_allclose(std, g[k + "_std_dt"], atol=1e-15)
        np.testing.assert_allclose(O.get_obs(mode, g[k + "_state"]), g[k + "_obs"], atol=1e-14)
        np.testing.assert_allclose(O.get_state(mode, g[k + "_obs"]), g[k + "_state_from_obs"], atol=1e-12)
        m, s = O.prior_disturbance(mode, g[k + "_state"].shape[0])
        np.testing.assert_allclose(m, g[k + "_dist_mean"]); np.testing.assert_allclose(s, g[k + "_dist_std"])
        if t is not None:
            np.testing.assert_allclose(tn, g[k + "_t_next"])


@pytest.mark.parametrize("mode,name", [("Unicycle", "unicycle_layer_b256.npz"), ("SimulatedCars", "cars_layer_b512.npz")])
def test_layer_vs_golden(golden, mode, name):
    g = golden(name)
    st, ac, mu, sg = (tt(g[k]) for k in ("state", "action", "mean", "sigma"))
    P, q, G, h = O.ASSEMBLE[mode](st, ac, mu, sg, gamma_b=float(g["gamma_b"]))
    np.testing.assert_allclose(P.numpy(), g["P"], rtol=0, atol=0)
    np.testing.assert_allclose(G.numpy(), g["G"], rtol=1e-6, atol=1e-6)
    np.testing.assert_allclose(h.numpy(), g["h"], rtol=2e-6, atol=1e-4)
    a = ac.clone().requires_grad_(True)
    fin = O.safe_action(mode, st, a, mu, sg, gamma_b=float(g["gamma_b"]))
    assert np.abs(fin.detach().numpy() - g["safe_action"]).max() < 2e-5
    (fin * tt(g["grad_w"])).sum().backward()
    gr, gg = a.grad.numpy(), g["grad_action"]
    assert np.linalg.norm(gr - gg) / np.linalg.norm(gg) < 1e-3
    # exact leg
    fe = O.safe_action(mode, st, ac, mu, sg, solver="exact", gamma_b=float(g["gamma_b"]))
    assert np.abs(fe.numpy() - g["safe_action"]).max() < 2e-5


def test_cascade_vs_golden(golden):
    g = golden("cascade_layer.npz")
    for i in range(g["state"].shape[0]):
        P, q, G, h = O.assemble_unicycle_cascade(g["action"][i], g["state"][i], g["mean"][i], g["sigma"][i])
        np.testing.assert_allclose(G, g["G"][i], atol=1e-13); np.testing.assert_allclose(h, g["h"][i], rtol=1e-12)
        np.testing.assert_allclose(P, g["P"][i])
        P, q, G, h = O.assemble_cars_cascade(g["cars_action"][i], g["cars_state"][i], None, g["cars_sigma"][i])
        np.testing.assert_allclose(G, g["cars_G"][i], atol=1e-10); np.testing.assert_allclose(h, g["cars_h"][i], rtol=1e-12)


# ----------------------------------------------------------------------------- restated qpth vs exact optimum
@pytest.mark.parametrize("mode,B", [("Unicycle", 256), ("SimulatedCars", 512)])
def test_qpth_restatement_vs_exact(mode, B):
    if mode == "Unicycle":
        st, ac, mu, sg = O.synth_unicycle(B, seed=7)
    else:
        st, ac, mu, sg, _ = O.synth_cars(B, seed=7)
    P, q, G, h = O.ASSEMBLE[mode](tt(st).double(), tt(ac).double(), tt(mu).double(), tt(sg).double(), gamma_b=20.0)
    Gn, hn, n = O.normalise_rows(G, h)
    x, lam, s, info = qpth_pdipm.pdipm_forward(P, q, Gn, hn, eps=1e-4, notImprovedLim=10, maxIter=100000)
    xe, lame, act, viol = exact_qp.solve_exact(P.numpy(), q.numpy(), Gn.numpy(), hn.numpy())
    assert viol.max() < 1e-9
    nu = ac.shape[1]
    assert np.abs(x.numpy()[:, :nu] - xe[:, :nu]).max() < 1e-5      # batch-global stop over-converges
    # every normalised row satisfied by the exact optimum
    assert (hn.numpy() - np.einsum("bmj,bj->bm", Gn.numpy(), xe)).min() > -1e-9
    # and the KKT multipliers agree where they matter
    assert np.abs(lam.numpy() - lame).max() < 1e-2 * max(1.0, np.abs(lame).max())


def test_qpth_backward_vs_finite_differences():
    st, ac, mu, sg = O.synth_unicycle(64, seed=11)
    st, mu, sg = tt(st).double(), tt(mu).double(), tt(sg).double()
    a = tt(ac).double().requires_grad_(True)
    fin = O.safe_action("Unicycle", st, a, mu, sg, assembly_dtype=torch.float64, eps=1e-10, gamma_b=20.0)
    w = torch.from_numpy(np.random.default_rng(3).normal(size=fin.shape))
    (fin * w).sum().backward()
    g = a.grad.numpy()
    eps = 1e-6
    fd = np.zeros_like(g)
    for j in range(2):
        d = torch.zeros_like(a); d[:, j] = eps
        fp = O.safe_action("Unicycle", st, (a + d).detach(), mu, sg, solver="exact", assembly_dtype=torch.float64, gamma_b=20.0)
        fm = O.safe_action("Unicycle", st, (a - d).detach(), mu, sg, solver="exact", assembly_dtype=torch.float64, gamma_b=20.0)
        fd[:, j] = ((fp - fm) * w).sum(1).numpy() / (2 * eps)
    err = np.abs(g - fd).max(1)
    # active-set changes inside the FD stencil give isolated outliers; the bulk must agree tightly
    assert np.median(err) < 1e-6 and (err < 1e-3).mean() > 0.9


# ----------------------------------------------------------------------------- oracle vs LIVE reference source
needs_ref = pytest.mark.skipif(not ref_loader.reference_available(), reason="/root/reference not present")


@needs_ref
@pytest.mark.parametrize("mode", ["Unicycle", "SimulatedCars"])
def test_oracle_vs_live_reference_layer(mode):
    ref = ref_loader.load_reference()
    env = ref.UnicycleEnv() if mode == "Unicycle" else ref.SimulatedCarsEnv()
    layer = ref.CBFQPLayer(env, ref_loader.make_args(), gamma_b=20, k_d=3.0, l_p=0.03)
    if mode == "Unicycle":
        st, ac, mu, sg = O.synth_unicycle(300, seed=5)
    else:
        st, ac, mu, sg, _ = O.synth_cars(300, seed=5)
    P, q, G, h = layer.get_cbf_qp_constraints(tt(st), tt(ac), tt(mu), tt(sg))
    P2, q2, G2, h2 = O.ASSEMBLE[mode](tt(st), tt(ac), tt(mu), tt(sg), gamma_b=20.0)
    assert (P - P2).abs().max() == 0 and (G - G2).abs().max() < 1e-5
    assert ((h - h2).abs() / (h.abs() + 1)).max() < 2e-6
    fa = layer.get_safe_action(tt(st), tt(ac), tt(mu), tt(sg))
    fo = O.safe_action(mode, tt(st), tt(ac), tt(mu), tt(sg), gamma_b=20.0)
    assert (fa - fo).abs().max() < 1e-5


@needs_ref
def test_oracle_vs_live_reference_envs():
    ref = ref_loader.load_reference()
    env = ref.UnicycleEnv()
    env.reset()
    r = O.unicycle_reset(1)
    st, step, last = r["state"], r["episode_step"], r["last_goal_dist"]
    rng = np.random.default_rng(1)
    for k in range(50):
        a = rng.uniform(-1.2, 1.2, 2)
        o, rew, d, info = env.step(a)
        out = O.unicycle_env_step(st, a[None], step, last)
        np.testing.assert_allclose(out["obs"][0], o, atol=1e-13)
        assert abs(out["reward"][0] - rew) < 1e-13
        st, step, last = out["state"], out["episode_step"], out["last_goal_dist"]


def test_rollout_step_vs_golden(golden):
    """oracle restatement of generate_model_rollouts' transition vs what the reference's own function pushed."""
    g = golden("model_rollouts.npz")
    for mode, k in (("Unicycle", "unicycle"), ("SimulatedCars", "simulatedcars")):
        nobs, rew, done, nt = O.rollout_step(mode, g[k + "_obs"], g[k + "_action"], g[k + "_t"], g[k + "_eps"])
        np.testing.assert_allclose(nobs, g[k + "_next_obs"], rtol=1e-13, atol=1e-13)
        np.testing.assert_allclose(rew, g[k + "_reward"], rtol=1e-13, atol=1e-13)
        np.testing.assert_array_equal(~done, g[k + "_mask"].astype(bool))
        np.testing.assert_allclose(nt, g[k + "_next_t"], rtol=0, atol=1e-15)
        assert (~g[k + "_mask"].astype(bool)).sum() >= 2          # the fixture exercises done
    # the reference's double reward_goal (generate_rollouts.py:50,53)
    d = ~g["unicycle_mask"].astype(bool)
    assert (g["unicycle_reward"][d] > 1.9).all()
