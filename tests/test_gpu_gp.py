"""GPU parity of the disturbance-GP posterior kernel: `rcbf_gp_predict_{f32,f64}` through the
C ABI (sac_rcbf_b200.gp_model) against oracle/gp_oracle.py (float64 dense Cholesky; parity unpinned vs gpytorch, see
that file's header) on the golden history produced by the reference's own envs + append_transition.

Tolerances: float64 I/O 1e-7 relative (the kernel accumulates in float64; outputscale - |w|^2 cancels up to
n * outputscale / noise ~ 1e4..1e6); float32 I/O 2e-6 relative (one rounding of the inputs and of the outputs)."""
import os
import types

import numpy as np
import pytest
import torch

from oracle import gp_oracle as G

pytestmark = pytest.mark.gpu

GOLD = np.load(os.path.join(os.path.dirname(__file__), "golden", "gp_disturbance.npz"))
MAX_STD = {"unicycle": [0.2] * 3, "simulatedcars": [0, 0.2] * 5}


def _bank(mode):
    from sac_rcbf_b200.gp_model import DisturbanceGPBank
    tx, ty = GOLD[mode + "_train_x"], GOLD[mode + "_train_y"]
    xs, ys = tx.std(0), ty.std(0)
    return DisturbanceGPBank(tx / (xs + 1e-8), ty / (ys + 1e-8), MAX_STD[mode], x_scale=xs, y_scale=ys + 1e-8)


@pytest.mark.parametrize("mode", ["unicycle", "simulatedcars"])
def test_fit_and_predict_match_oracle_on_golden_history(mode):
    bank = _bank(mode)
    bank.train(70)
    assert np.allclose(bank.raw.cpu().numpy(), GOLD[mode + "_raw"], rtol=1e-6, atol=1e-6)
    x = torch.as_tensor(GOLD[mode + "_test_x"]).cuda()
    scale = np.abs(GOLD[mode + "_train_y"]).max(0) + 1e-300
    mean, std = bank.predict(x)
    assert mean.dtype == torch.float64 and mean.is_cuda
    assert np.all(np.abs(mean.cpu().numpy() - GOLD[mode + "_mean"]) <= 1e-7 * scale)
    assert np.allclose(std.cpu().numpy(), GOLD[mode + "_std"], rtol=1e-7, atol=0)
    mean32, std32 = bank.predict(x.float())
    assert mean32.dtype == torch.float32
    assert np.all(np.abs(mean32.cpu().numpy() - GOLD[mode + "_mean"]) <= 2e-6 * scale)
    assert np.allclose(std32.cpu().numpy(), GOLD[mode + "_std"], rtol=2e-6, atol=0)


@pytest.mark.parametrize("mode", ["unicycle", "simulatedcars"])
def test_far_field_path_agrees_with_exact_kernels_and_falls_back_per_point(mode):
    """At the reference's lengthscale the bank qualifies for the far-field polynomial (rcbf_gp_posterior::ff_coef).
    (i) it must agree with the exact low-rank kernel (far_field = False) far below the parity tolerance;
    (ii) points outside its validity bound (here: 50x farther out, interleaved with ordinary points so both kinds share
    warps) are finished exactly inside the same launch."""
    bank = _bank(mode)
    bank.raw = torch.as_tensor(GOLD[mode + "_raw"]).cuda().clone()
    bank.build_posterior()
    assert bank.far_field_active and bank._post[0].ff_coef
    exact = _bank(mode)
    exact.raw = bank.raw.clone()
    exact.far_field = False
    exact.build_posterior()
    assert not exact.far_field_active and not exact._post[0].ff_coef
    test = np.repeat(GOLD[mode + "_test_x"], 3, axis=0)
    test[1::3] *= 50.0
    test[2::7] *= -400.0
    x = torch.as_tensor(test).cuda()
    (m_ff, s_ff), (m_ex, s_ex) = bank.predict(x), exact.predict(x)
    scale = torch.as_tensor(np.abs(GOLD[mode + "_train_y"]).max(0) + 1e-300).cuda()
    assert float(((m_ff - m_ex).abs() / scale).max()) <= 2e-9
    assert float(((s_ff - s_ex).abs() / s_ex).max()) <= 2e-9
    ora = G.DisturbanceGPs(GOLD[mode + "_train_x"], GOLD[mode + "_train_y"], MAX_STD[mode], training_iter=0)
    for g, raw in zip(ora.gps, GOLD[mode + "_raw"]):
        g.raw = raw.copy()
    # oracle check on the ordinary points only: hundreds of data-widths away the posterior extrapolates the (tiny)
    # linear / quadratic kernel components and the dense float64 Cholesky of the oracle is itself ill-conditioned there
    plain = np.array([i for i in range(len(test)) if i % 3 == 0 and i % 7 != 2])
    om, os_ = ora.predict_disturbance(test[plain])
    assert np.all(np.abs(m_ff.cpu().numpy()[plain] - om) <= 1e-7 * scale.cpu().numpy())
    assert np.allclose(s_ff.cpu().numpy()[plain], os_, rtol=1e-7, atol=0)


@pytest.mark.parametrize("n_test", [1, 31, 33, 1000])
def test_ragged_batches_and_unpadded_training_sets(n_test):
    """n = 117 training points (not a multiple of the 32-point chunk), batches that do not fill a 32-point tile."""
    rng = np.random.default_rng(n_test)
    n = 117
    x = rng.uniform(-3, 3, (n, 3))
    y = np.stack([0.2 * np.sin(x[:, 0]) + 0.05 * rng.standard_normal(n), 0.1 * x[:, 1] ** 2], 1)
    from sac_rcbf_b200.gp_model import DisturbanceGPBank
    bank = DisturbanceGPBank(x, y, [0.2, 0.2])
    bank.set_hyperparameters(lengthscale=[2.0, 50.0], outputscale=[0.5, 0.3], noise=[0.02, 0.001])
    test = rng.uniform(-3.5, 3.5, (n_test, 3))
    mean, std = bank.predict(torch.as_tensor(test).cuda())
    for g in range(2):
        gp = G.ExactGP(x, y[:, g], 0.2)
        gp.raw = bank.raw[g].cpu().numpy().copy()
        p = gp.predict(test)
        assert np.allclose(mean[:, g].cpu().numpy(), p["mean"], rtol=0, atol=1e-7 * np.abs(y[:, g]).max())
        assert np.allclose(std[:, g].cpu().numpy() ** 2, p["f_var"], rtol=1e-7)


def test_strided_test_points_are_read_in_place():
    """`env._state4[:, :3]` (row stride 4) must give the same answer as a dense copy, for both kernel families."""
    bank = _bank("unicycle")
    bank.raw = torch.as_tensor(GOLD["unicycle_raw"]).cuda().clone()
    st4 = torch.zeros(77, 4, device="cuda")
    st4[:, :3] = torch.as_tensor(np.resize(GOLD["unicycle_test_x"], (77, 3)), dtype=torch.float32).cuda()
    st4[:, 3] = 123.0
    for ff in (True, False):
        bank.far_field = ff
        bank.build_posterior()
        m0, s0 = bank.predict(st4[:, :3].contiguous())
        m1, s1 = bank.predict(st4[:, :3])
        assert torch.equal(m0, m1) and torch.equal(s0, s1)


def test_full_rank_factor_path():
    """lengthscale ~ data spread in 3-D: the kernel matrix is far from low rank, the factor keeps (nearly) all n rows
    and the kernel walks several 64-row tiles."""
    rng = np.random.default_rng(5)
    n = 200
    x = rng.uniform(-3, 3, (n, 3))
    y = (np.sin(x[:, 0]) * np.cos(x[:, 1]) + 0.05 * rng.standard_normal(n))[:, None]
    from sac_rcbf_b200.gp_model import DisturbanceGPBank
    bank = DisturbanceGPBank(x, y, [0.2])
    bank.set_hyperparameters(lengthscale=[0.8], outputscale=[1.0], noise=[0.01])
    bank.build_posterior()
    assert bank._post[0].tile_rows == 64 and bank.ranks[0] > 64
    test = rng.uniform(-3, 3, (257, 3))
    mean, std = bank.predict(torch.as_tensor(test).cuda())
    gp = G.ExactGP(x, y[:, 0], 0.2)
    gp.raw = bank.raw[0].cpu().numpy().copy()
    p = gp.predict(test)
    assert np.allclose(mean[:, 0].cpu().numpy(), p["mean"], rtol=0, atol=1e-7)
    assert np.allclose(std[:, 0].cpu().numpy() ** 2, p["f_var"], rtol=1e-7)


def test_estimator_facade_matches_reference_predict_contract():
    """gp_model.py:86-114: ndarray in -> dict of ndarrays (mean, f_var incl. noise, f_covar, lower/upper ci = 2 std)."""
    from sac_rcbf_b200.gp_model import GPyDisturbanceEstimator
    rng = np.random.default_rng(1)
    x = rng.uniform(-1, 1, (60, 3))
    y = 0.5 + 0.1 * rng.standard_normal(60)
    est = GPyDisturbanceEstimator(x, y, 0.2, device=torch.device("cuda"))
    est.train(10)
    gp = G.ExactGP(x, y, 0.2)
    gp.train(10)
    test = rng.uniform(-1, 1, (9, 3))
    out = est.predict(test)
    p = gp.predict(test.astype(np.float32))
    assert isinstance(out["mean"], np.ndarray) and out["mean"].shape == (9,)
    assert np.allclose(out["mean"], p["mean"], rtol=2e-6) and np.allclose(out["f_var"], p["f_var"], rtol=4e-6)
    assert np.allclose(out["upper_ci"] - out["lower_ci"], 4 * np.sqrt(out["f_var"]), rtol=1e-5)
    assert np.allclose(np.diag(out["f_covar"]), out["f_var"], rtol=1e-4)
    tout = est.predict(torch.as_tensor(test, dtype=torch.float32))
    assert torch.is_tensor(tout["mean"]) and tout["mean"].device.type == "cpu"
    sd = est.model.state_dict()                                  # dynamics.py:418
    assert "covar_module.base_kernel.raw_lengthscale" in sd and "likelihood.noise_covar.raw_noise" in sd


@pytest.mark.parametrize("mode", ["Unicycle", "SimulatedCars"])
def test_dynamics_model_fits_and_predicts_on_device(mode, tmp_path):
    """append_transition -> (automatic) fit_gp_model -> predict_disturbance, the reference's call sequence
    (main.py:151-153 / sac_cbf.py:230), against the oracle fitted on the same transitions; then save / load."""
    import sac_rcbf_b200 as S
    key = mode.lower()
    tx, ty = GOLD[key + "_train_x"], GOLD[key + "_train_y"]
    env = S.build_env(types.SimpleNamespace(env_name=mode))
    args = types.SimpleNamespace(cuda=True, gp_model_size=10 * len(tx), l_p=0.03)
    dm = S.DynamicsModel(env, args)
    # feed the golden history through append_transition: next_state = prior_next + dt * disturbance
    t = None if mode == "Unicycle" else np.zeros(len(tx))
    rng = np.random.default_rng(3)
    u = rng.uniform(-1, 1, (len(tx), dm.n_u))
    prior_next, _, _ = dm.predict_next_state(tx, u, t, use_gps=False)
    assert not dm.disturb_estimators
    dm.append_transition(tx, u, prior_next + env.dt * ty, t_batch=t)      # len(tx) == gp_model_size / 10 -> one refit
    assert dm.disturb_estimators and len(dm.disturb_estimators) == dm.n_s
    assert np.allclose(dm.train_y, ty, rtol=0, atol=1e-9 * (1 + np.abs(ty).max()) / env.dt)
    ora = G.DisturbanceGPs(dm.train_x, dm.train_y, S.MAX_STD[mode], training_iter=70)
    test = GOLD[key + "_test_x"]
    om, os_ = ora.predict_disturbance(test)
    scale = np.abs(dm.train_y).max(0) + 1e-300
    m, s = dm.predict_disturbance(test)                                   # ndarray in -> ndarray out
    assert isinstance(m, np.ndarray) and np.all(np.abs(m - om) <= 1e-6 * scale) and np.allclose(s, os_, rtol=1e-6)
    xt = torch.as_tensor(test, dtype=torch.float32).cuda()
    mt, st = dm.predict_disturbance(xt)                                   # tensor in -> device tensors out
    assert mt.is_cuda and mt.dtype == torch.float32 and mt.shape == xt.shape
    assert np.all(np.abs(mt.cpu().numpy() - om) <= 3e-6 * scale) and np.allclose(st.cpu().numpy(), os_, rtol=3e-6)
    m1, s1 = dm.predict_disturbance(test[0])                              # 1-D in -> 1-D out (dynamics.py:364-366)
    assert m1.shape == (dm.n_s,) and np.allclose(m1, m[0]) and np.allclose(s1, s[0])
    # predict_next_state now adds dt * GP mean and returns dt * GP std (dynamics.py:92,96)
    nxt, std, _ = dm.predict_next_state(test, u[:len(test)], None if t is None else t[:len(test)])
    pn, _, _ = dm.predict_next_state(test, u[:len(test)], None if t is None else t[:len(test)], use_gps=False)
    assert np.allclose(nxt, pn + env.dt * m, rtol=0, atol=1e-12 * (1 + np.abs(pn).max()))
    assert np.allclose(std, env.dt * s)
    dm.save_disturbance_models(str(tmp_path))
    dm2 = S.DynamicsModel(env, args)
    dm2.load_disturbance_models(str(tmp_path))
    m2, s2 = dm2.predict_disturbance(test)
    assert np.array_equal(m2, m) and np.array_equal(s2, s)


def test_model_rollout_transition_uses_the_fitted_gps_on_the_device():
    """generate_rollouts.py:29-31 with fitted GPs: get_state -> predict_next_state (prior + dt * GP mean, dt * GP std)
    -> Gaussian sample.  Device tensors in, device tensors out, against the oracle transition fed with the ORACLE's GP
    posterior at the same states."""
    import sac_rcbf_b200 as S
    from oracle import rcbf_oracle as O
    tx, ty = GOLD["unicycle_train_x"], GOLD["unicycle_train_y"]
    env = S.build_env(types.SimpleNamespace(env_name="Unicycle"))
    dm = S.DynamicsModel(env, types.SimpleNamespace(cuda=True, gp_model_size=3000, l_p=0.03))
    dm._install_gp_bank(tx, ty)
    dm._gp_bank.raw = torch.as_tensor(GOLD["unicycle_raw"]).cuda().clone()
    dm._gp_bank.build_posterior()
    rng = np.random.default_rng(11)
    B = 300
    state = GOLD["unicycle_test_x"][rng.integers(0, 40, B)] + 0.01 * rng.standard_normal((B, 3))
    dist = np.linalg.norm(np.array([2.5, 2.5]) - state[:, :2], axis=1)
    obs = np.concatenate([state[:, :2], np.cos(state[:, 2:3]), np.sin(state[:, 2:3]), np.zeros((B, 2)),
                          np.exp(-dist)[:, None]], 1)
    act = rng.uniform(-1, 1, (B, 2))
    eps = rng.standard_normal((B, 3))
    ora = G.DisturbanceGPs(tx, ty, MAX_STD["unicycle"], training_iter=0)
    for g, raw in zip(ora.gps, GOLD["unicycle_raw"]):
        g.raw = raw.copy()
    om, os_ = ora.predict_disturbance(O.get_state("Unicycle", obs))
    want = O.rollout_step("Unicycle", obs, act, None, eps, mean=om, std=os_)
    dev = lambda a: torch.as_tensor(a, dtype=torch.float64).cuda()  # noqa: E731
    got = S.rollout_transition(env, dm, dev(obs), dev(act), None, dev(eps))
    assert got[0].is_cuda
    assert np.allclose(got[0].cpu().numpy(), want[0], rtol=0, atol=1e-9)
    assert np.allclose(got[1].cpu().numpy(), want[1], rtol=0, atol=1e-9)
    assert np.array_equal(got[2].cpu().numpy(), want[2])


def test_reference_demo_shapes_one_dimensional_inputs_and_empty_batches():
    """gp_model.py:117-134 (__main__ demo): 1-D train_x / test_x tensors, prior_std = 0, 50 training iterations."""
    from sac_rcbf_b200.gp_model import GPyDisturbanceEstimator
    gen = torch.Generator().manual_seed(0)
    train_x = torch.linspace(0, 1, 100)
    train_y = torch.sin(train_x * (2 * np.pi)) + torch.randn(train_x.size(), generator=gen) * 0.2
    est = GPyDisturbanceEstimator(train_x, train_y, 0.0, device=torch.device("cuda"))
    est.train(50)
    test_x = torch.linspace(0, 1, 51)
    pred = est.predict(test_x)
    gp = G.ExactGP(train_x.numpy().astype(np.float64)[:, None], train_y.numpy().astype(np.float64), 0.0)
    gp.train(50)
    p = gp.predict(test_x.numpy().astype(np.float64)[:, None])
    assert pred["mean"].shape == (51,) and pred["mean"].dtype == torch.float32
    assert np.allclose(pred["mean"].numpy(), p["mean"], rtol=0, atol=2e-6 * np.abs(train_y.numpy()).max())
    assert np.allclose(pred["f_var"].numpy(), p["f_var"], rtol=4e-6)
    assert pred["f_covar"].shape == (51, 51)
    empty = est.predict(torch.zeros(0))
    assert empty["mean"].shape == (0,)
    m, s = est.bank.predict(torch.zeros(0, 1, device="cuda"))
    assert m.shape == (0, 1) and s.shape == (0, 1)
