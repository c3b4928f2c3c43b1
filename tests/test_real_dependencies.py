"""Pin hooks for the reference's unvendored dependencies: qpth (rcbf_sac/diff_cbf_qp.py:7,139), quadprog
(rcbf_sac/cbf_qp.py:276) and gpytorch (rcbf_sac/gp_model.py:12-114) are absent from this image, so the QP solve and the GP
posterior are pinned to restatements ("parity unpinned").  Each test below probes for the real package at run time and,
when an image ships it, compares the restatement with it -- the moment one of them appears the oracle is pinned (or the
difference is flagged) without anyone having to remember."""
import numpy as np
import pytest
import torch

from oracle import exact_qp, qpth_pdipm, rcbf_oracle as O

tt = torch.from_numpy


def _real(name):
    """import the REAL package or skip (oracle/ref_loader.py may have put a stand-in module under the same name)"""
    mod = pytest.importorskip(name)
    if getattr(mod, "__rcbf_stub__", False):
        pytest.skip("%s is not installed (only the ref_loader stand-in is present)" % name)
    return mod


def test_oracle_reports_its_qp_backend():
    assert "qpth" in O.QP_BACKEND


def test_restated_pdipm_vs_real_qpth():
    _real("qpth")
    from qpth.qp import QPFunction

    for mode, synth in (("Unicycle", O.synth_unicycle), ("SimulatedCars", O.synth_cars)):
        d = synth(256, seed=5)
        st, ac, mu, sg = (tt(np.asarray(a)) for a in d[:4])
        P, q, G, h = O.ASSEMBLE[mode](st, ac, mu, sg, gamma_b=20.0)
        Gn, hn, _ = O.normalise_rows(G, h)
        e = torch.empty(0, dtype=torch.float64)
        args = (P.double(), q.double(), Gn.double().requires_grad_(True), hn.double().requires_grad_(True), e, e)
        x_real = QPFunction(verbose=0, check_Q_spd=False, maxIter=100000, notImprovedLim=10, eps=1e-4)(*args)
        g_real = torch.autograd.grad(x_real.sum(), args[3])[0]
        args2 = (P.double(), q.double(), Gn.double().requires_grad_(True), hn.double().requires_grad_(True), e, e)
        x_mine = qpth_pdipm.QPFunction(verbose=0, check_Q_spd=False, maxIter=100000, notImprovedLim=10, eps=1e-4)(*args2)
        g_mine = torch.autograd.grad(x_mine.sum(), args2[3])[0]
        # same algorithm, same stopping rule: iterates agree to rounding; both within eps of the exact optimum
        assert float((x_real - x_mine).abs().max()) < 1e-6, mode
        assert float((g_real - g_mine).norm() / g_real.norm()) < 1e-5, mode
        # and the end-to-end oracle, which prefers the real package, agrees with its forced restatement
        a = O.safe_action(mode, st, ac, mu, sg, gamma_b=20.0)
        b = O.safe_action(mode, st, ac, mu, sg, gamma_b=20.0, force_restatement=True)
        assert float((a - b).abs().max()) < 1e-5, mode


def test_exact_enumerator_vs_real_quadprog():
    quadprog = _real("quadprog")
    rng = np.random.default_rng(3)
    for _ in range(200):
        G = rng.standard_normal((9, 3))
        h = rng.uniform(-0.2, 2.0, 9)
        Pm = np.diag([10.0, 1e-4, 1e7])                                         # cbf_qp.py:146
        x_real = quadprog.solve_qp(Pm, np.zeros(3), -G.T, -h)[0]                # cbf_qp.py:276 convention: C'x >= b
        x_mine = exact_qp.solve_exact(Pm[None], np.zeros((1, 3)), G[None], h[None])[0][0]
        assert np.abs(x_real - x_mine).max() < 1e-7 * max(1.0, np.abs(x_real).max())


def test_gp_oracle_vs_real_gpytorch():
    gpytorch = _real("gpytorch")
    from oracle import gp_oracle

    rng = np.random.default_rng(0)
    x = rng.uniform(-1, 1, (200, 3))
    y = 0.3 * np.sin(x[:, 0]) + 0.05 * rng.standard_normal(200)
    xt = rng.uniform(-1, 1, (50, 3))
    prior_std = 0.2

    class BaseGPy(gpytorch.models.ExactGP):                                      # rcbf_sac/gp_model.py:11-27
        def __init__(self, tx, ty, lik):
            super().__init__(tx, ty, lik)
            self.mean_module = gpytorch.means.ZeroMean()
            self.covar_module = gpytorch.kernels.ScaleKernel(
                gpytorch.kernels.RBFKernel(lengthscale_prior=gpytorch.priors.NormalPrior(1e5, 1e-5)),
                outputscale_prior=gpytorch.priors.NormalPrior(prior_std + 1e-6, 1e-5))
            self.covar_module.base_kernel.lengthscale = 1e5
            self.covar_module.outputscale = prior_std + 1e-6

        def forward(self, z):
            return gpytorch.distributions.MultivariateNormal(self.mean_module(z), self.covar_module(z))

    lik = gpytorch.likelihoods.GaussianLikelihood().double()
    model = BaseGPy(torch.as_tensor(x), torch.as_tensor(y), lik).double()
    g = gp_oracle.ExactGP(x, y, prior_std)
    # same initial hyper-parameters and the same loss (exact MLL + the two Normal priors, per data point) at the start
    assert abs(g.lengthscale - float(model.covar_module.base_kernel.lengthscale)) < 1e-3
    assert abs(g.outputscale - float(model.covar_module.outputscale)) < 1e-9
    assert abs(g.noise - float(lik.noise)) < 1e-9
    model.train(); lik.train()
    mll = gpytorch.mlls.ExactMarginalLogLikelihood(lik, model)
    loss = -mll(model(torch.as_tensor(x)), torch.as_tensor(y))                   # gp_model.py:63-82
    assert abs(float(loss) - g.loss_and_grad()[0]) < 1e-6 * max(1.0, abs(float(loss)))
    model.eval(); lik.eval()
    with torch.no_grad():                                                        # (exact posterior: no fast_pred_var)
        pred = lik(model(torch.as_tensor(xt)))
    out = g.predict(xt, min_variance=0.0)
    assert np.abs(out["mean"] - pred.mean.numpy()).max() < 1e-6
    assert np.abs(out["f_var"] - pred.variance.numpy()).max() < 1e-6
