"""CPU tests of the disturbance-GP row: the oracle against its known answers and the golden
fixture, and the host side of sac_rcbf_b200.gp_model (training, factor packing, rank truncation) against the oracle.
The CUDA predict kernel itself is covered by tests/test_gpu_gp.py."""
import os

import numpy as np
import pytest
import torch

from oracle import gp_oracle as G
from tests.gp_sim import eval_far_field, eval_packed

GOLD = np.load(os.path.join(os.path.dirname(__file__), "golden", "gp_disturbance.npz"))
MAX_STD = {"unicycle": [0.2] * 3, "simulatedcars": [0, 0.2] * 5}


def _toy(n=80, d=3, seed=0):
    rng = np.random.default_rng(seed)
    x = rng.uniform(-3, 3, (n, d))
    y = 0.3 + 0.1 * x[:, 0] + 0.05 * rng.standard_normal(n)
    return x / (x.std(0) + 1e-8), y / (y.std() + 1e-8), rng


def test_oracle_known_answer_rank_one():
    """At the reference's pinned lengthscale (1e5) the kernel matrix is os * 1 1^T to ~1e-10: Sherman-Morrison."""
    x, y, rng = _toy()
    gp = G.ExactGP(x, y, 0.2)
    gp.raw[2] = G.inv_softplus(0.05 - G.NOISE_LOWER)
    p = gp.predict(rng.uniform(-2, 2, (7, 3)))
    m, v = G.rank_one_closed_form(y, gp.outputscale, gp.noise)
    assert np.allclose(p["mean"], m, rtol=0, atol=2e-6 * abs(m))
    assert np.allclose(p["f_var"], v, rtol=1e-6)


def test_oracle_gradient_matches_finite_differences():
    x, y, _ = _toy(40)
    gp = G.ExactGP(x, y, 0.2)
    gp.raw[0] = G.inv_softplus(1.5)      # a lengthscale at which d/d lengthscale is not swamped by the 1e5 scale
    _, g = gp.loss_and_grad(priors=False)
    for k in range(3):
        h = 1e-6
        gp.raw[k] += h
        lp, _ = gp.loss_and_grad(priors=False)
        gp.raw[k] -= 2 * h
        lm, _ = gp.loss_and_grad(priors=False)
        gp.raw[k] += h
        fd = (lp - lm) / (2 * h)
        assert abs(fd - g[k]) <= 1e-6 * max(1.0, abs(fd)), (k, fd, g[k])


@pytest.mark.parametrize("mode", ["unicycle", "simulatedcars"])
def test_oracle_reproduces_golden(mode):
    gps = G.DisturbanceGPs(GOLD[mode + "_train_x"], GOLD[mode + "_train_y"], MAX_STD[mode], training_iter=70)
    assert np.allclose(np.stack([g.raw for g in gps.gps]), GOLD[mode + "_raw"], rtol=1e-9, atol=1e-9)
    mean, std = gps.predict_disturbance(GOLD[mode + "_test_x"])
    assert np.allclose(mean, GOLD[mode + "_mean"], rtol=1e-9, atol=1e-15)
    assert np.allclose(std, GOLD[mode + "_std"], rtol=1e-9, atol=1e-18)


def _bank(mode, train=True):
    from sac_rcbf_b200.gp_model import DisturbanceGPBank
    tx, ty = GOLD[mode + "_train_x"], GOLD[mode + "_train_y"]
    xs, ys = tx.std(0), ty.std(0)
    bank = DisturbanceGPBank(tx / (xs + 1e-8), ty / (ys + 1e-8), MAX_STD[mode], device="cpu", x_scale=xs,
                             y_scale=ys + 1e-8)
    if train:
        bank.train(70)
    return bank


@pytest.mark.parametrize("mode", ["unicycle", "simulatedcars"])
def test_bank_training_matches_oracle(mode):
    bank = _bank(mode)
    assert np.allclose(bank.raw.numpy(), GOLD[mode + "_raw"], rtol=1e-7, atol=1e-7)


@pytest.mark.parametrize("mode", ["unicycle", "simulatedcars"])
def test_packed_posterior_matches_oracle(mode):
    """The factor the kernel reads (eigen-truncated to the numerical rank, tile-major, zero padded) reproduces the
    dense Cholesky posterior of the oracle."""
    bank = _bank(mode, train=False)
    bank.raw = torch.as_tensor(GOLD[mode + "_raw"]).clone()
    bank.build_posterior()
    assert max(bank.ranks) <= 16                       # lengthscale 1e5: numerically low rank
    mean, std = eval_packed(bank, GOLD[mode + "_test_x"])
    scale = np.abs(GOLD[mode + "_train_y"]).max(0) + 1e-300
    assert np.all(np.abs(mean - GOLD[mode + "_mean"]) <= 1e-7 * scale)
    assert np.allclose(std, GOLD[mode + "_std"], rtol=1e-7, atol=0)


@pytest.mark.parametrize("mode", ["unicycle", "simulatedcars"])
def test_far_field_tables_match_oracle_and_gate_distant_points(mode):
    """The second-order far-field polynomial the host tabulates (what k_gp_farfield evaluates) reproduces the oracle on
    the golden test points, all of which pass the validity bound; points 400x farther out must be refused by it."""
    bank = _bank(mode, train=False)
    bank.raw = torch.as_tensor(GOLD[mode + "_raw"]).clone()
    bank.build_posterior()
    assert bank.far_field_active
    mean, std, ok = eval_far_field(bank, GOLD[mode + "_test_x"])
    assert ok.all()
    scale = np.abs(GOLD[mode + "_train_y"]).max(0) + 1e-300
    assert np.all(np.abs(mean - GOLD[mode + "_mean"]) <= 1e-8 * scale)
    assert np.allclose(std, GOLD[mode + "_std"], rtol=1e-9, atol=0)
    _, _, ok_far = eval_far_field(bank, 400.0 * GOLD[mode + "_test_x"])
    assert not ok_far[:, -1].any()      # (GPs with a ~zero outputscale keep a much wider bound: checked per GP)
    bank.far_field = False
    bank.build_posterior()
    assert not bank.far_field_active and not bank._post[0].ff_coef


def test_packed_posterior_moderate_lengthscale_keeps_enough_rank():
    """A kernel that is NOT low rank (lengthscale ~ data spread): the validated truncation must keep what the dense
    solve needs (here it ends up using most or all of the n rows) and still match the oracle."""
    x, y, rng = _toy(150)
    from sac_rcbf_b200.gp_model import DisturbanceGPBank
    bank = DisturbanceGPBank(x, y[:, None], [0.2], device="cpu")
    bank.set_hyperparameters(lengthscale=[0.7], outputscale=[0.8], noise=[0.01])
    bank.build_posterior()
    gp = G.ExactGP(x, y, 0.2)
    gp.raw = bank.raw[0].numpy().copy()
    test = rng.uniform(-2.5, 2.5, (60, 3))
    p = gp.predict(test)
    mean, std = eval_packed(bank, test)
    assert bank._post[0].tile_rows == 64 and bank.ranks[0] > 48 and not bank.far_field_active
    assert np.allclose(mean[:, 0], p["mean"], rtol=0, atol=1e-7 * np.abs(y).max())
    assert np.allclose(std[:, 0] ** 2, p["f_var"], rtol=1e-7)


def test_predict_without_cuda_fails_loudly():
    from sac_rcbf_b200 import RcbfLibraryError
    bank = _bank("unicycle", train=False)
    with pytest.raises(RcbfLibraryError):
        bank.predict(torch.zeros(4, 3))


def test_analytic_training_gradient_equals_autograd():
    """train() uses the closed-form gradient (one Cholesky inverse per step); it must be the autograd gradient."""
    from sac_rcbf_b200.gp_model import _sq_dists
    bank = _bank("unicycle", train=False)
    d2 = _sq_dists(bank.train_x, bank.train_x)
    raw = bank.raw.clone()
    raw[:, 0] = torch.log(torch.expm1(torch.tensor(1.7, dtype=torch.float64)))
    raw[:, 2] = -1.0
    raw.requires_grad_(True)
    loss = bank._neg_mll(raw, d2)
    loss.sum().backward()
    with torch.no_grad():
        loss2, grad2 = bank._neg_mll_and_grad(raw, d2)
    assert torch.allclose(loss, loss2, rtol=1e-13, atol=0)
    assert torch.allclose(raw.grad, grad2, rtol=1e-10, atol=0)
