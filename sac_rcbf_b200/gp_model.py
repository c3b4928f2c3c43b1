"""Disturbance GPs on the device -- rcbf_sac/gp_model.py behind the same class name, plus the batched bank that
`DynamicsModel.predict_disturbance` uses.

What runs where:
  * `predict` (every env / rollout / SAC-update step): ONE launch of the hand-written CUDA kernel
    `rcbf_gp_predict_{f32,f64}` (csrc/rcbf_gp.cu) for all output dimensions; test points and results never leave the
    device (the reference goes tensor -> numpy -> tensor -> cpu per call, dynamics.py:359-362, gp_model.py:102-107).
  * `train` / posterior factorisation (once per refit, every gp_model_size/10 transitions, dynamics.py:303-304): dense
    float64 linear algebra through torch.linalg on the device -- library code, off the per-step path.

Model (gp_model.py:11-27): ExactGP, ZeroMean, ScaleKernel(RBFKernel); softplus-constrained raw parameters, noise lower
bound 1e-4; lengthscale initialised at 1e5 with a Normal(1e5, 1e-5) prior, outputscale at prior_std + 1e-6 with a
Normal(prior_std + 1e-6, 1e-5) prior; loss = -(exact log marginal likelihood + log priors) / n, Adam(lr = 0.1).
Differences from the reference, on purpose: (i) float64 exact Cholesky everywhere instead of gpytorch's float32
CG / Lanczos (LOVE) approximations -- deterministic, and what those approximate; (ii) all output dimensions are
trained as one batch.
"""
import ctypes as C
import math

import numpy as np
import torch

from . import _lib
from . import _params as P

NOISE_LOWER = 1e-4           # gpytorch GaussianLikelihood: GreaterThan(1e-4)
MIN_VARIANCE = 1e-6          # gpytorch.settings.min_variance for float32 (the dtype the reference predicts in)
LENGTHSCALE_INIT = 1e5       # gp_model.py:21
PRIOR_SIGMA = 1e-5           # gp_model.py:18,20
_CHUNK = 64                  # kLrChunk in csrc/rcbf_gp.cu (a multiple of kGpChunk)


def _inv_softplus(v):
    v = torch.as_tensor(v, dtype=torch.float64)
    return v + torch.log(-torch.expm1(-v))


def _sq_dists(a, b):
    return (a[:, None, :] - b[None, :, :]).square().sum(-1)


class DisturbanceGPBank:
    """n_gp independent exact GPs that share their (normalised) training inputs.

    train_x (n, d), train_y (n, n_gp): the arrays the reference passes to GPyDisturbanceEstimator, i.e. already divided
    by (std + 1e-8) (dynamics.py:327-330).  x_scale / y_scale are applied inside the predict kernel: test points are
    DIVIDED by x_scale (dynamics.py:375) and the outputs MULTIPLIED by y_scale (dynamics.py:378-379)."""

    def __init__(self, train_x, train_y, prior_std, device=None, x_scale=None, y_scale=None):
        if device is None:
            _lib.require_cuda()
            device = torch.device("cuda", torch.cuda.current_device())
        self.device = torch.device(device)   # "cpu" is accepted for fitting / factorisation only; predict() needs CUDA
        f64 = dict(dtype=torch.float64, device=self.device)
        self.train_x = torch.as_tensor(np.asarray(train_x, np.float64) if not torch.is_tensor(train_x) else train_x)
        self.train_x = self.train_x.to(**f64).reshape(self.train_x.shape[0], -1).contiguous()
        ty = torch.as_tensor(np.asarray(train_y, np.float64) if not torch.is_tensor(train_y) else train_y).to(**f64)
        self.train_y = ty.reshape(self.train_x.shape[0], -1).contiguous()
        self.n, self.d = self.train_x.shape
        self.n_gp = self.train_y.shape[1]
        if self.d > 16:
            raise ValueError("the GP kernel supports at most 16 input dimensions")
        prior_std = torch.as_tensor(np.broadcast_to(np.asarray(prior_std, np.float64), (self.n_gp,)).copy(), **f64)
        self.prior_outputscale = prior_std + 1e-6
        # raw parameters, one row per GP: [lengthscale, outputscale, noise] before softplus
        self.raw = torch.stack([_inv_softplus(torch.full((self.n_gp,), LENGTHSCALE_INIT)).to(**f64),
                                _inv_softplus(self.prior_outputscale.cpu()).to(**f64),
                                torch.zeros(self.n_gp, **f64)], 1).contiguous()
        self.x_scale = torch.ones(self.d, **f64) if x_scale is None else torch.as_tensor(x_scale).to(**f64)
        self.y_scale = torch.ones(self.n_gp, **f64) if y_scale is None else torch.as_tensor(y_scale).to(**f64)
        self.include_noise = True
        self.far_field = True       # allow the far-field polynomial path when the bank qualifies (build_posterior)
        self.far_field_active = False
        self._post = None

    # ---------------------------------------------------------------------------------------------- hyper-parameters
    @property
    def lengthscale(self):
        return torch.nn.functional.softplus(self.raw[:, 0])

    @property
    def outputscale(self):
        return torch.nn.functional.softplus(self.raw[:, 1])

    @property
    def noise(self):
        return torch.nn.functional.softplus(self.raw[:, 2]) + NOISE_LOWER

    def set_hyperparameters(self, lengthscale=None, outputscale=None, noise=None):
        f64 = dict(dtype=torch.float64, device=self.device)
        if lengthscale is not None:
            self.raw[:, 0] = _inv_softplus(torch.as_tensor(lengthscale, dtype=torch.float64).cpu()).to(**f64)
        if outputscale is not None:
            self.raw[:, 1] = _inv_softplus(torch.as_tensor(outputscale, dtype=torch.float64).cpu()).to(**f64)
        if noise is not None:
            self.raw[:, 2] = _inv_softplus(torch.as_tensor(noise, dtype=torch.float64).cpu() - NOISE_LOWER).to(**f64)
        self._post = None

    # ---------------------------------------------------------------------------------------------- training
    def _neg_mll(self, raw, d2):
        """-(log N(y; 0, K + s2 I) + log priors) / n for every GP (gp_model.py:69-79), shape (n_gp,)."""
        n = self.n
        sp = torch.nn.functional.softplus
        ls, os_, s2 = sp(raw[:, 0]), sp(raw[:, 1]), sp(raw[:, 2]) + NOISE_LOWER
        khat = os_[:, None, None] * torch.exp(-d2[None] / (2.0 * ls * ls)[:, None, None])
        khat = khat + s2[:, None, None] * torch.eye(n, dtype=d2.dtype, device=d2.device)
        chol = torch.linalg.cholesky(khat)
        y = self.train_y.t().unsqueeze(-1)                                    # (n_gp, n, 1)
        alpha = torch.cholesky_solve(y, chol)
        logp = -0.5 * (y * alpha).sum((1, 2)) - torch.log(torch.diagonal(chol, dim1=1, dim2=2)).sum(1) \
            - 0.5 * n * math.log(2.0 * math.pi)
        lp = lambda v, mu: -0.5 * ((v - mu) / PRIOR_SIGMA) ** 2 - math.log(PRIOR_SIGMA * math.sqrt(2.0 * math.pi))
        return -(logp + lp(ls, LENGTHSCALE_INIT) + lp(os_, self.prior_outputscale)) / n

    def _neg_mll_and_grad(self, raw, d2):
        """Same loss with its analytic gradient w.r.t. the raw parameters,
        d logp / d theta = 1/2 tr((alpha alpha^T - Khat^-1) dKhat/dtheta): one Cholesky + one Cholesky inverse per
        step instead of autograd's backward through the factorisation (about 3x less float64 work)."""
        n = self.n
        sp = torch.nn.functional.softplus
        ls, os_, s2 = sp(raw[:, 0]), sp(raw[:, 1]), sp(raw[:, 2]) + NOISE_LOWER
        e = torch.exp(-d2[None] / (2.0 * ls * ls)[:, None, None])
        khat = os_[:, None, None] * e
        khat.diagonal(dim1=1, dim2=2).add_(s2[:, None])
        chol = torch.linalg.cholesky(khat)
        kinv = torch.cholesky_inverse(chol)
        y = self.train_y.t()                                                   # (n_gp, n)
        alpha = torch.einsum("gij,gj->gi", kinv, y)
        logp = -0.5 * (y * alpha).sum(1) - torch.log(torch.diagonal(chol, dim1=1, dim2=2)).sum(1) \
            - 0.5 * n * math.log(2.0 * math.pi)
        lp = lambda v, mu: -0.5 * ((v - mu) / PRIOR_SIGMA) ** 2 - math.log(PRIOR_SIGMA * math.sqrt(2.0 * math.pi))
        loss = -(logp + lp(ls, LENGTHSCALE_INIT) + lp(os_, self.prior_outputscale)) / n
        w = alpha[:, :, None] * alpha[:, None, :] - kinv                       # 2 dlogp / dKhat
        we = w * e
        g_ls = 0.5 * os_ / ls ** 3 * (we * d2[None]).sum((1, 2)) - (ls - LENGTHSCALE_INIT) / PRIOR_SIGMA ** 2
        g_os = 0.5 * we.sum((1, 2)) - (os_ - self.prior_outputscale) / PRIOR_SIGMA ** 2
        g_s2 = 0.5 * torch.diagonal(w, dim1=1, dim2=2).sum(1)
        grad = -torch.stack([g_ls, g_os, g_s2], 1) * torch.sigmoid(raw) / n
        return loss, grad

    def train(self, training_iter, verbose=False):
        d2 = _sq_dists(self.train_x, self.train_x)
        raw = self.raw.clone().requires_grad_(True)
        opt = torch.optim.Adam([raw], lr=0.1)                                  # gp_model.py:66
        for i in range(training_iter):
            opt.zero_grad()
            with torch.no_grad():
                loss, raw.grad = self._neg_mll_and_grad(raw, d2)              # GPs are independent: per-GP Adam steps
            if verbose:
                sp = torch.nn.functional.softplus
                print('\tIter %d/%d - Loss: %s   lengthscale: %s   noise: %s' % (
                    i + 1, training_iter, loss.detach().cpu().numpy(), sp(raw[:, 0]).detach().cpu().numpy(),
                    (sp(raw[:, 2]) + NOISE_LOWER).detach().cpu().numpy()))
            opt.step()
        self.raw = raw.detach().contiguous()
        self._post = None

    # ---------------------------------------------------------------------------------------------- posterior cache
    def _dense_reference(self, zs, g, kmat, chol):
        """float64 dense posterior of GP g at normalised points zs (validation of a truncated factor only)."""
        os_, ls = self.outputscale[g], self.lengthscale[g]
        ks = os_ * torch.exp(-_sq_dists(self.train_x, zs) / (2.0 * ls * ls))   # (n, B)
        v = torch.linalg.solve_triangular(chol, ks, upper=False)
        mean = (v * torch.linalg.solve_triangular(chol, self.train_y[:, g:g + 1], upper=False)).sum(0)
        return mean, os_ - (v * v).sum(0)

    def build_posterior(self, rank_tol=1e-8):
        """Factor (K_g + s2_g I)^-1 = F_g^T F_g for every GP from a float64 eigendecomposition of K_g, keep only the
        rows the numerical rank needs and CHECK the truncation against the dense solve on probe points (training
        points, jittered and far-away copies); any GP whose truncated factor misses `rank_tol` keeps all n rows."""
        n, dev = self.n, self.device
        f64 = dict(dtype=torch.float64, device=dev)
        n_pad = (n + _CHUNK - 1) // _CHUNK * _CHUNK
        gen = torch.Generator(device="cpu").manual_seed(1234)
        idx = torch.randperm(n, generator=gen)[:min(n, 96)].to(dev)
        base = self.train_x[idx]
        spread = self.train_x.std(0, unbiased=False) + 1e-12
        probes = torch.cat([base, base + 0.5 * spread * torch.randn(base.shape, generator=gen).to(**f64),
                            3.0 * base + spread])
        d2 = _sq_dists(self.train_x, self.train_x)
        factors, projs = [], []
        for g in range(self.n_gp):
            os_, ls, s2 = self.outputscale[g], self.lengthscale[g], self.noise[g]
            kmat = os_ * torch.exp(-d2 / (2.0 * ls * ls))
            lam, q = torch.linalg.eigh(kmat)
            lam, q = lam.flip(0), q.flip(1)                                    # descending
            chol = torch.linalg.cholesky(kmat + s2 * torch.eye(n, **f64))
            m_ref, v_ref = self._dense_reference(probes, g, kmat, chol)
            ks = os_ * torch.exp(-_sq_dists(self.train_x, probes) / (2.0 * ls * ls))
            y = self.train_y[:, g]
            chosen = n
            for thr in (1e-10, 1e-13):
                r = max(1, int((lam > thr * lam[0]).sum()))
                if r >= n // 2:
                    break
                f = q[:, :r].t() / torch.sqrt(lam[:r].clamp_min(0.0) + s2)[:, None]
                w = f @ ks
                m_err = ((w * (f @ y)[:, None]).sum(0) - m_ref).abs().max()
                v_err = ((os_ - (w * w).sum(0)) - v_ref).abs().max()
                if m_err <= rank_tol * (y.abs().max() + 1e-300) and v_err <= rank_tol * (s2 + v_ref.abs().max()):
                    chosen = r
                    break
            f = q[:, :chosen].t() / torch.sqrt(lam[:chosen].clamp_min(0.0) + s2)[:, None]
            factors.append(f)
            projs.append(f @ y)
        self.ranks = [f.shape[0] for f in factors]
        max_rank = max(self.ranks)
        tile_rows = 4 if max_rank <= 4 else 8 if max_rank <= 8 else 16 if max_rank <= 48 else 64
        max_tiles = (max_rank + tile_rows - 1) // tile_rows
        factor = torch.zeros(self.n_gp, max_tiles, n_pad, tile_rows, **f64)
        proj_y = torch.zeros(self.n_gp, max_tiles * tile_rows, **f64)
        r_tiles = torch.zeros(self.n_gp, dtype=torch.int32, device=dev)
        for g, (f, py) in enumerate(zip(factors, projs)):
            r = f.shape[0]
            tiles = (r + tile_rows - 1) // tile_rows
            fp = torch.zeros(tiles * tile_rows, n_pad, **f64)
            fp[:r, :n] = f
            factor[g, :tiles] = fp.reshape(tiles, tile_rows, n_pad).permute(0, 2, 1)
            proj_y[g, :r] = py
            r_tiles[g] = tiles
        dim_pad = (self.d + 3) // 4 * 4
        train_z = torch.zeros(n_pad, dim_pad, **f64)
        train_z[:n, :self.d] = self.train_x
        inv_x = torch.zeros(dim_pad, **f64)
        inv_x[:self.d] = 1.0 / self.x_scale
        hyp = torch.stack([1.0 / (2.0 * self.lengthscale ** 2), self.outputscale, self.noise, self.y_scale], 1)
        ff_coef, ff_amax, ff_zmax = self._far_field_tables(train_z, factor, proj_y, hyp, max_tiles, tile_rows, dim_pad)
        keep = (train_z, inv_x, hyp.contiguous(), r_tiles, factor.contiguous(), proj_y, ff_coef, ff_amax)
        post = P.GpPosterior(ff_coef=None if ff_coef is None else ff_coef.data_ptr(),
                             ff_amax=None if ff_amax is None else ff_amax.data_ptr(), ff_zmax=ff_zmax,
                             train_z=keep[0].data_ptr(), inv_x_scale=keep[1].data_ptr(), hyp=keep[2].data_ptr(),
                             r_tiles=keep[3].data_ptr(), factor=keep[4].data_ptr(), proj_y=keep[5].data_ptr(),
                             n_pad=n_pad, n_in=self.d, dim_pad=dim_pad, n_gp=self.n_gp, max_tiles=max_tiles,
                             tile_rows=tile_rows, include_noise=int(self.include_noise), min_variance=MIN_VARIANCE)
        self._post = (post, keep)
        return self

    def _far_field_tables(self, train_z, factor, proj_y, hyp, max_tiles, tile_rows, dim_pad, rel_tol=1e-9):
        """Coefficients of the second-order far-field expansion (include/rcbf_b200.h, rcbf_gp_posterior::ff_coef) and
        the per-GP validity bound.  With a_j = inv_2l2 (s - 2 z*.z_j + t_j), s = |z*|^2, t_j = |z_j|^2:
            w_r = os sum_j F_rj (1 - a_j + a_j^2 / 2),  remainder <= os |F_r|_1 a_max^3 / 6.
        Returns (None, None, 0) when the fast path is off or not even the training points would qualify."""
        self.far_field_active = False
        if not self.far_field or max_tiles != 1 or tile_rows > 16:
            return None, None, 0.0
        f64 = dict(dtype=torch.float64, device=self.device)
        z = train_z                                            # (n_pad, dim_pad), zero rows beyond n
        t = (z * z).sum(1)
        zmax = float(t.max().sqrt())
        nc = 3 + 2 * dim_pad + dim_pad * (dim_pad + 1) // 2
        coef = torch.zeros(self.n_gp, tile_rows, nc, **f64)
        amax = torch.zeros(self.n_gp, **f64)
        iu = torch.triu_indices(dim_pad, dim_pad).to(self.device)
        offdiag = (iu[0] != iu[1]).to(torch.float64) + 1.0     # symmetric matrix folded onto k <= l
        for g in range(self.n_gp):
            i2, os_, noise = hyp[g, 0], hyp[g, 1], hyp[g, 2]
            f = factor[g, 0].t()                               # (tile_rows, n_pad)
            s0, s1, s2 = f.sum(1), f @ t, f @ (t * t)
            v0, v1 = f @ z, f @ (z * t[:, None])
            mm = torch.einsum("rj,jk,jl->rkl", f, z, z)
            coef[g, :, 0] = os_ * (s0 - i2 * s1 + 0.5 * i2 * i2 * s2)
            coef[g, :, 1] = os_ * (-i2 * s0 + i2 * i2 * s1)
            coef[g, :, 2] = os_ * 0.5 * i2 * i2 * s0
            coef[g, :, 3:3 + dim_pad] = os_[None, None] * (2.0 * i2 * v0 - 2.0 * i2 * i2 * v1)
            coef[g, :, 3 + dim_pad:3 + 2 * dim_pad] = os_ * (-2.0 * i2 * i2) * v0
            coef[g, :, 3 + 2 * dim_pad:] = os_ * 2.0 * i2 * i2 * mm[:, iu[0], iu[1]] * offdiag
            rows = max(1, self.ranks[g])
            e_g = os_ * f.abs().sum(1).max()
            y = self.train_y[:, g]
            tol_w = rel_tol * torch.minimum(noise / (2.0 * os_.sqrt() * rows),
                                            (y.abs().max() + 1e-300) / (proj_y[g].abs().sum() + 1e-300))
            amax[g] = (6.0 * tol_w / (e_g + 1e-300)) ** (1.0 / 3.0)
        a_train = (2.0 * zmax) ** 2 * hyp[:, 0]
        if bool((a_train > amax).any()):
            return None, None, 0.0
        self.far_field_active = True
        return coef.contiguous(), amax, zmax

    # ---------------------------------------------------------------------------------------------- prediction
    def predict(self, test_x):
        """test_x (B, d) float32 / float64 device tensor -> (mean, std), each (B, n_gp), same dtype, on the device.
        A row-strided view with unit column stride (e.g. `env._state4[:, :3]`) is read in place."""
        if self.device.type != "cuda":
            raise _lib.RcbfLibraryError("GP prediction runs in the CUDA kernel only (no CPU fallback); bank is on %s"
                                        % self.device)
        lib = _lib.load()
        if self._post is None:
            self.build_posterior()
        post, _keep = self._post
        x = test_x.detach()
        if x.dtype not in (torch.float32, torch.float64):
            x = x.float()
        x = x.to(self.device)
        if not (x.dim() == 2 and x.shape[1] == self.d and x.stride(1) == 1 and x.stride(0) >= self.d):
            x = x.reshape(-1, self.d).contiguous()
        post.test_stride = x.stride(0) if x.shape[0] > 1 else self.d
        mean = torch.empty(x.shape[0], self.n_gp, dtype=x.dtype, device=self.device)
        std = torch.empty_like(mean)
        fn = lib.rcbf_gp_predict_f64 if x.dtype == torch.float64 else lib.rcbf_gp_predict_f32
        with torch.cuda.device(self.device):
            rc = fn(_lib.ptr(x), x.shape[0], C.byref(post), _lib.ptr(mean), _lib.ptr(std), _lib.stream_ptr(self.device))
        _lib.check(rc, "rcbf_gp_predict")
        return mean, std

    # ---------------------------------------------------------------------------------------------- (de)serialisation
    def state_dict(self):
        return {"raw": self.raw.cpu(), "train_x": self.train_x.cpu(), "train_y": self.train_y.cpu(),
                "x_scale": self.x_scale.cpu(), "y_scale": self.y_scale.cpu(),
                "prior_outputscale": self.prior_outputscale.cpu()}

    def load_state_dict(self, sd):
        self.raw = sd["raw"].to(self.device, torch.float64).contiguous()
        self._post = None


_SD_KEYS = ("covar_module.base_kernel.raw_lengthscale", "covar_module.raw_outputscale",
            "likelihood.noise_covar.raw_noise")   # gpytorch names of the three trainable tensors, bank column order


class BankMember:
    """One GP of a bank behind the attribute names reference code touches on a GPyDisturbanceEstimator:
    `.model.state_dict()` / `.model.load_state_dict()` (dynamics.py:404,418) and `.predict()`."""

    def __init__(self, bank, index):
        self.bank, self.index = bank, index
        self.device = bank.device
        self.model = self
        self.likelihood = self

    def state_dict(self):
        """The key set `BaseGPy.state_dict()` has under gpytorch (gp_model.py:11-27: the three raw parameters plus the
        constraint bounds and prior buffers gpytorch registers next to them), so that the reference's strict
        `model.load_state_dict(weights[i])` (dynamics.py:404) finds every key and no unexpected one.  gpytorch is absent
        from this image: the list follows gpytorch 1.x and is unverified here.  The float64 raw parameters (the bank
        trains in float64) travel separately, see DynamicsModel.save_disturbance_models."""
        r = self.bank.raw[self.index].cpu()
        f = lambda v: torch.tensor(float(v))  # noqa: E731
        pos_inf, prior_os = float("inf"), float(self.bank.prior_outputscale[self.index])
        return {
            "likelihood.noise_covar.raw_noise": r[2].reshape(1).float(),
            "likelihood.noise_covar.raw_noise_constraint.lower_bound": f(1e-4),
            "likelihood.noise_covar.raw_noise_constraint.upper_bound": f(pos_inf),
            "covar_module.raw_outputscale": r[1].reshape(()).float(),
            "covar_module.base_kernel.raw_lengthscale": r[0].reshape(1, 1).float(),
            "covar_module.base_kernel.lengthscale_prior.loc": f(1e5),
            "covar_module.base_kernel.lengthscale_prior.scale": f(1e-5),
            "covar_module.base_kernel.raw_lengthscale_constraint.lower_bound": f(0.0),
            "covar_module.base_kernel.raw_lengthscale_constraint.upper_bound": f(pos_inf),
            "covar_module.outputscale_prior.loc": f(prior_os),
            "covar_module.outputscale_prior.scale": f(1e-5),
            "covar_module.raw_outputscale_constraint.lower_bound": f(0.0),
            "covar_module.raw_outputscale_constraint.upper_bound": f(pos_inf),
        }

    def load_state_dict(self, sd, strict=False):
        if "raw_float64" in sd:     # files written by round-1 builds
            raw = sd["raw_float64"].to(torch.float64)
        else:
            raw = torch.stack([sd[k].reshape(-1)[0].to(torch.float64) for k in _SD_KEYS])
        self.bank.raw[self.index] = raw.to(self.bank.device)
        self.bank._post = None

    def train(self, training_iter, verbose=False):
        self.bank.train(training_iter, verbose)

    def predict(self, test_x):
        """gp_model.py:86-114 for this output dimension: test_x is already normalised; results come back on the CPU
        (ndarray in -> ndarrays out).  The bank's x / y scalings are NOT applied here, like the reference's member."""
        is_tensor = torch.is_tensor(test_x)
        x = test_x if is_tensor else torch.as_tensor(np.asarray(test_x, np.float32))   # to_tensor(.., FloatTensor)
        x = x.to(self.device).reshape(-1, self.bank.d)
        b = self.bank
        mean, std = b.predict(x * b.x_scale.to(x.dtype))
        ys = b.y_scale[self.index].to(x.dtype)
        mean, std = (mean[:, self.index] / ys).cpu(), (std[:, self.index] / ys).cpu()
        out = _Prediction(self, x, not is_tensor)
        out.update(mean=mean, f_var=std ** 2, lower_ci=mean - 2.0 * std, upper_ci=mean + 2.0 * std)  # gp_model.py:99-105
        if not is_tensor:
            for k in list(out):
                out[k] = out[k].numpy()
        return out

    def _full_covariance(self, x):
        b, g = self.bank, self.index
        os_, ls, s2 = b.outputscale[g], b.lengthscale[g], b.noise[g]
        z = x.to(torch.float64)
        chol = torch.linalg.cholesky(os_ * torch.exp(-_sq_dists(b.train_x, b.train_x) / (2 * ls * ls)) +
                                     s2 * torch.eye(b.n, dtype=torch.float64, device=b.device))
        v = torch.linalg.solve_triangular(chol, os_ * torch.exp(-_sq_dists(b.train_x, z) / (2 * ls * ls)), upper=False)
        cov = os_ * torch.exp(-_sq_dists(z, z) / (2 * ls * ls)) - v.t() @ v
        return (cov + s2 * torch.eye(z.shape[0], dtype=torch.float64, device=b.device)).to(x.dtype)


class _Prediction(dict):
    """dict returned by GPyDisturbanceEstimator.predict; 'f_covar' (B x B, gp_model.py:101) is built on first access."""

    def __init__(self, est, test_x, as_numpy):
        super().__init__()
        self._est, self._x, self._np = est, test_x, as_numpy

    def __missing__(self, key):
        if key != "f_covar":
            raise KeyError(key)
        cov = self._est._full_covariance(self._x).cpu()
        self[key] = cov.numpy() if self._np else cov
        return self[key]


class GPyDisturbanceEstimator(BankMember):
    """Same constructor / train / predict as rcbf_sac/gp_model.py:30-114 (one output dimension = a bank of one)."""

    def __init__(self, train_x, train_y, prior_std, likelihood=None, device=None):
        if likelihood is not None:
            raise NotImplementedError("only the default GaussianLikelihood (gp_model.py:50-51) is supported")
        ty = np.asarray(train_y.cpu() if torch.is_tensor(train_y) else train_y, np.float64).reshape(-1, 1)
        super().__init__(DisturbanceGPBank(train_x, ty, prior_std, device=device), 0)
        self.train_x = self.bank.train_x
        self.train_y = self.bank.train_y[:, 0]
