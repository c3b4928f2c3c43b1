"""TEST INFRASTRUCTURE ONLY -- CPU oracle for the disturbance-GP posterior.

PARITY UNPINNED: gpytorch is not in this image (and not under /root/reference), so the reference's
`GPyDisturbanceEstimator` (rcbf_sac/gp_model.py:11-114) cannot be run here.  This file restates, in float64 numpy,
the published exact-GP equations that class evaluates through gpytorch (version unpinned by the reference):

  model       gp_model.py:13-27   ExactGP, ZeroMean, ScaleKernel(RBFKernel):  k(a,b) = os * exp(-|a-b|^2 / (2 l^2))
  parameters  gpytorch Positive constraint = softplus(raw); GaussianLikelihood noise = softplus(raw_noise) + 1e-4
  init        gp_model.py:21-22   lengthscale = 1e5, outputscale = prior_std + 1e-6; raw_noise = 0
  priors      gp_model.py:18-20   Normal(1e5, 1e-5) on the lengthscale, Normal(prior_std + 1e-6, 1e-5) on the outputscale
  loss        gp_model.py:63-82   -ExactMarginalLogLikelihood = -(log N(y; 0, K + s2 I) + sum log-priors) / n, Adam(lr=0.1)
  predict     gp_model.py:86-114  likelihood(model(x)):  mean = K*^T (K + s2 I)^-1 y,
                                  'f_var' = os - diag(K*^T (K + s2 I)^-1 K*) + s2   (likelihood noise INCLUDED),
                                  clamped from below at gpytorch's float32 `min_variance` (1e-6)
  wrapper     dynamics.py:306-340,371-379  x / (std_x + 1e-8) at fit time but x / std_x at predict time;
                                  outputs scaled back by (std_y + 1e-8)

`fast_pred_var()` (LOVE) approximates the same variance with a Lanczos low-rank factor and gpytorch switches to CG above
800 training points, so the reference's own float32 output is itself only an approximation of these equations.
Known answers used to pin this file instead: the Sherman-Morrison closed form for the rank-one kernel the reference's
pinned lengthscale (1e5) produces, and finite differences of the marginal likelihood (tests/test_gp_host.py).
"""
import numpy as np

NOISE_LOWER = 1e-4          # gpytorch GaussianLikelihood default noise constraint GreaterThan(1e-4)
MIN_VARIANCE_F32 = 1e-6     # gpytorch.settings.min_variance for float32 tensors
LENGTHSCALE_INIT = 1e5      # gp_model.py:21
PRIOR_SIGMA = 1e-5          # gp_model.py:18,20


def softplus(x):
    return np.logaddexp(0.0, x)


def inv_softplus(v):
    return v + np.log(-np.expm1(-v))


def sigmoid(x):
    return 0.5 * (1.0 + np.tanh(0.5 * x))


def sq_dists(a, b):
    a = np.asarray(a, np.float64)
    b = np.asarray(b, np.float64)
    d = a[:, None, :] - b[None, :, :]
    return np.einsum("ijk,ijk->ij", d, d)


def kernel(a, b, lengthscale, outputscale):
    return outputscale * np.exp(-sq_dists(a, b) / (2.0 * lengthscale * lengthscale))


class ExactGP:
    """One output dimension.  `train_x` (n, d) and `train_y` (n,) are the already-normalised arrays the reference hands
    to GPyDisturbanceEstimator (dynamics.py:333)."""

    def __init__(self, train_x, train_y, prior_std):
        self.x = np.asarray(train_x, np.float64).reshape(len(train_y), -1)
        self.y = np.asarray(train_y, np.float64)
        self.prior_os = prior_std + 1e-6
        self.raw = np.array([inv_softplus(LENGTHSCALE_INIT), inv_softplus(self.prior_os), 0.0])

    # ---- hyper-parameters
    @property
    def lengthscale(self):
        return float(softplus(self.raw[0]))

    @property
    def outputscale(self):
        return float(softplus(self.raw[1]))

    @property
    def noise(self):
        return float(softplus(self.raw[2]) + NOISE_LOWER)

    # ---- marginal likelihood and its gradient w.r.t. the raw parameters
    def loss_and_grad(self, priors=True):
        """priors=False drops the two Normal priors (used only by the finite-difference self-check: at any
        lengthscale but 1e5 the prior term is ~1e19 and hides everything else in float64)."""
        pw = 1.0 if priors else 0.0
        n = len(self.y)
        l, os_, s2 = self.lengthscale, self.outputscale, self.noise
        d2 = sq_dists(self.x, self.x)
        e = np.exp(-d2 / (2.0 * l * l))
        khat = os_ * e + s2 * np.eye(n)
        chol = np.linalg.cholesky(khat)
        alpha = np.linalg.solve(chol.T, np.linalg.solve(chol, self.y))
        logdet = 2.0 * np.log(np.diag(chol)).sum()
        logp = -0.5 * self.y @ alpha - 0.5 * logdet - 0.5 * n * np.log(2.0 * np.pi)
        lp = lambda v, mu: -0.5 * ((v - mu) / PRIOR_SIGMA) ** 2 - np.log(PRIOR_SIGMA * np.sqrt(2.0 * np.pi))
        total = (logp + pw * (lp(l, LENGTHSCALE_INIT) + lp(os_, self.prior_os))) / n
        kinv = np.linalg.solve(chol.T, np.linalg.solve(chol, np.eye(n)))
        w = np.outer(alpha, alpha) - kinv                       # d logp / d Khat = w / 2
        g_l = 0.5 * np.sum(w * (os_ * e * d2 / l ** 3)) - pw * (l - LENGTHSCALE_INIT) / PRIOR_SIGMA ** 2
        g_os = 0.5 * np.sum(w * e) - pw * (os_ - self.prior_os) / PRIOR_SIGMA ** 2
        g_s2 = 0.5 * np.trace(w)
        grad = np.array([g_l, g_os, g_s2]) * sigmoid(self.raw) / n
        return -total, -grad

    def train(self, training_iter, lr=0.1, betas=(0.9, 0.999), eps=1e-8):
        """torch.optim.Adam defaults (gp_model.py:66)."""
        m = np.zeros(3)
        v = np.zeros(3)
        losses = []
        for it in range(1, training_iter + 1):
            loss, g = self.loss_and_grad()
            losses.append(loss)
            m = betas[0] * m + (1 - betas[0]) * g
            v = betas[1] * v + (1 - betas[1]) * g * g
            mh = m / (1 - betas[0] ** it)
            vh = v / (1 - betas[1] ** it)
            self.raw = self.raw - lr * mh / (np.sqrt(vh) + eps)
        return losses

    # ---- posterior
    def predict(self, test_x, min_variance=MIN_VARIANCE_F32):
        """{'mean', 'f_var'} as gp_model.py:99-101 returns them ('f_var' includes the likelihood noise)."""
        xs = np.asarray(test_x, np.float64).reshape(-1, self.x.shape[1])
        l, os_, s2 = self.lengthscale, self.outputscale, self.noise
        khat = kernel(self.x, self.x, l, os_) + s2 * np.eye(len(self.y))
        chol = np.linalg.cholesky(khat)
        ks = kernel(self.x, xs, l, os_)                          # (n, B)
        v = np.linalg.solve(chol, ks)
        mean = v.T @ np.linalg.solve(chol, self.y)
        var = os_ - np.einsum("ij,ij->j", v, v) + s2
        return {"mean": mean, "f_var": np.maximum(var, min_variance)}


def rank_one_closed_form(y, outputscale, noise):
    """Known answer when exp(-d^2 / 2l^2) == 1 for all pairs (what float32 sees at l = 1e5): K = os 1 1^T, and by
    Sherman-Morrison  mean = os * sum(y) / (s2 + n os),  var = os * s2 / (s2 + n os) + s2."""
    n = len(y)
    return outputscale * np.sum(y) / (noise + n * outputscale), outputscale * noise / (noise + n * outputscale) + noise


class DisturbanceGPs:
    """dynamics.py:306-340 (fit_gp_model) + :371-379 (fitted branch of predict_disturbance)."""

    def __init__(self, train_x, train_y, max_std, training_iter=70):
        self.train_x = np.asarray(train_x, np.float64)
        self.train_y = np.asarray(train_y, np.float64)
        xn = self.train_x / (np.std(self.train_x, axis=0) + 1e-8)          # dynamics.py:327-328
        yn = self.train_y / (np.std(self.train_y, axis=0) + 1e-8)          # dynamics.py:329-330
        self.gps = []
        for i in range(self.train_y.shape[1]):
            gp = ExactGP(xn, yn[:, i], max_std[i])
            gp.train(training_iter)
            self.gps.append(gp)

    def predict_disturbance(self, test_x):
        test_x = np.asarray(test_x, np.float64)
        x_std = np.std(self.train_x, axis=0)
        y_std = np.std(self.train_y, axis=0)
        xs = test_x / x_std                                                # dynamics.py:375 (no +1e-8 here)
        means = np.zeros(test_x.shape)
        f_std = np.zeros(test_x.shape)
        for i, gp in enumerate(self.gps):
            p = gp.predict(xs)
            means[:, i] = p["mean"] * (y_std[i] + 1e-8)                    # dynamics.py:378
            f_std[:, i] = np.sqrt(p["f_var"]) * (y_std[i] + 1e-8)          # dynamics.py:379
        return means, f_std
