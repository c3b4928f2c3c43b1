"""DynamicsModel -- prior-dynamics part of rcbf_sac/dynamics.py behind the same method names.

SURVEY.md section 8a rows D3-D5: `predict_next_state` (prior f, g + dt * disturbance mean), `get_state`, `get_obs`,
the zero-mean / MAX_STD prior branch of `predict_disturbance`, `append_transition` (history ring buffer).
Section 8f row 1: `fit_gp_model` / the fitted branch of `predict_disturbance` / save + load run on the device through
`gp_model.DisturbanceGPBank` (one CUDA launch for all state dimensions, no host round trip for tensor inputs).
`disturbance_fn` still overrides everything when given.
"""
import numpy as np
import torch

from . import _lib

DYNAMICS_MODE = {'Unicycle': {'n_s': 3, 'n_u': 2},          # dynamics.py:22
                 'SimulatedCars': {'n_s': 10, 'n_u': 1}}    # dynamics.py:23
MAX_STD = {'Unicycle': [2e-1, 2e-1, 2e-1], 'SimulatedCars': [0, 0.2, 0, 0.2, 0, 0.2, 0, 0.2, 0, 0.2]}  # dynamics.py:24


class DynamicsModel:

    def __init__(self, env, args, disturbance_fn=None):
        """env: needs dynamics_mode, dt (+ kp, k_brake for SimulatedCars); args: gp_model_size, l_p (optional), cuda.
        disturbance_fn(state_batch ndarray) -> (mean, std) replaces the fitted-GP branch when given."""
        _lib.require_cuda()
        self._lib = _lib.load()
        self.env = env
        if env.dynamics_mode not in DYNAMICS_MODE:
            raise Exception('Unknown Dynamics mode.')
        self.n_s = DYNAMICS_MODE[self.env.dynamics_mode]['n_s']
        self.n_u = DYNAMICS_MODE[self.env.dynamics_mode]['n_u']
        self.disturb_estimators = None
        self.disturbance_fn = disturbance_fn
        self.disturbance_history = dict()
        self.history_counter = 0
        self.max_history_count = getattr(args, 'gp_model_size', 2000)
        self.disturbance_history['state'] = np.zeros((self.max_history_count, self.n_s))
        self.disturbance_history['disturbance'] = np.zeros((self.max_history_count, self.n_s))
        self.train_x = None
        self.train_y = None
        if hasattr(args, 'l_p'):
            self.l_p = args.l_p
        dev_num = getattr(args, "device_num", None)
        self.device = torch.device("cuda", torch.cuda.current_device() if dev_num is None else int(dev_num))

    # ------------------------------------------------------------------------------------------------ prior step
    def _predict_next_device(self, st, u, t, mean):
        """st (B,n_s), u (B,n_u), t (B,) or None, mean (B,n_s) or None: device tensors of one dtype (f32 or f64)."""
        suf = "f64" if st.dtype == torch.float64 else "f32"
        n = st.shape[0]
        nxt = torch.empty_like(st)
        dev = self.device
        with torch.cuda.device(dev):
            if self.env.dynamics_mode == 'Unicycle':
                fn = getattr(self._lib, "rcbf_unicycle_predict_next_" + suf)
                rc = fn(_lib.ptr(st), _lib.ptr(u), _lib.ptr(mean), n, float(self.env.dt), _lib.ptr(nxt),
                        _lib.stream_ptr(dev))
            else:
                fn = getattr(self._lib, "rcbf_cars_predict_next_" + suf)
                rc = fn(_lib.ptr(st), _lib.ptr(u), _lib.ptr(t), _lib.ptr(mean), n, float(self.env.dt),
                        float(getattr(self.env, 'kp', 4.0)), float(getattr(self.env, 'k_brake', 20.0)), _lib.ptr(nxt),
                        _lib.stream_ptr(dev))
        _lib.check(rc, "rcbf_predict_next")
        return nxt

    def predict_next_state(self, state_batch, u_batch, t_batch=None, use_gps=True):
        """next = s + dt (f(s,t) + g(s) u) [+ dt * disturbance mean]; returns (next, dt * std, t + dt)
        (dynamics.py:60-105).  ndarray in -> ndarray out (float64, like the reference); tensors stay on the device."""
        is_tensor = torch.is_tensor(state_batch)
        expand_dims = len(state_batch.shape) == 1
        if is_tensor:
            dt_ = state_batch.dtype if state_batch.dtype in (torch.float32, torch.float64) else torch.float32
            st = state_batch.detach().to(self.device, dt_)
            u = torch.as_tensor(u_batch).detach().to(self.device, dt_)
            t = None if t_batch is None else torch.as_tensor(t_batch).detach().to(self.device, dt_)
        else:
            dt_ = torch.float64
            st = torch.as_tensor(np.asarray(state_batch, np.float64)).to(self.device)
            u = torch.as_tensor(np.asarray(u_batch, np.float64)).to(self.device)
            t = None if t_batch is None else torch.as_tensor(np.asarray(t_batch, np.float64)).to(self.device)
        if expand_dims:
            st = st.unsqueeze(0)
        u = u.reshape(st.shape[0], self.n_u).contiguous()
        st = st.contiguous()
        if self.env.dynamics_mode == 'SimulatedCars':
            if t is None:   # get_f / get_g have no default for t_batch (dynamics.py:158,164)
                raise TypeError("predict_next_state for SimulatedCars needs t_batch")
            t = t.reshape(st.shape[0]).contiguous()
        if use_gps:
            mean, std = self.predict_disturbance(st)
            mean, std = mean.contiguous(), std
        else:
            mean, std = None, torch.zeros_like(st)
        nxt = self._predict_next_device(st, u, t, mean)
        std = self.env.dt * std
        if expand_dims:
            nxt, std = nxt.squeeze(0), std.squeeze(0)
        t_out = t_batch
        if t_batch is not None:
            t_out = t_batch + self.env.dt
        if is_tensor:
            return nxt, std, t_out
        return nxt.cpu().numpy(), std.cpu().numpy(), t_out

    def predict_next_obs(self, state, u):
        next_state, _, _ = self.predict_next_state(state, u)
        return self.get_obs(next_state)

    # ------------------------------------------------------------------------------------------------ obs <-> state
    def get_state(self, obs):
        """dynamics.py:190-232.  Same kind / dtype / device out as in; tensors are NOT round-tripped through numpy."""
        expand_dims = len(obs.shape) == 1
        xp = torch if torch.is_tensor(obs) else np
        if expand_dims:
            obs = obs[None]
        if self.env.dynamics_mode == 'Unicycle':
            theta = xp.atan2(obs[:, 3], obs[:, 2]) if xp is torch else np.arctan2(obs[:, 3], obs[:, 2])
            state_batch = xp.stack((obs[:, 0], obs[:, 1], theta), 1) if xp is torch else \
                np.stack((obs[:, 0], obs[:, 1], theta), 1).astype(np.float64)
        elif self.env.dynamics_mode == 'SimulatedCars':
            state_batch = obs.clone() if xp is torch else np.copy(obs)
            state_batch[:, ::2] *= 100.0
            state_batch[:, 1::2] *= 30.0
        else:
            raise Exception('Unknown dynamics')
        if expand_dims:
            state_batch = state_batch[0]
        return state_batch

    def get_obs(self, state_batch):
        """dynamics.py:234-261 (Unicycle returns the 4 state-derived entries; rollouts rebuild the other 3)."""
        xp = torch if torch.is_tensor(state_batch) else np
        if self.env.dynamics_mode == 'Unicycle':
            obs = xp.stack((state_batch[:, 0], state_batch[:, 1], xp.cos(state_batch[:, 2]),
                            xp.sin(state_batch[:, 2])), 1)
        elif self.env.dynamics_mode == 'SimulatedCars':
            obs = state_batch.clone() if xp is torch else np.copy(state_batch)
            obs[:, ::2] /= 100.0
            obs[:, 1::2] /= 30.0
        else:
            raise Exception('Unknown dynamics')
        return obs

    # ------------------------------------------------------------------------------------------------ disturbance
    def predict_disturbance(self, test_x):
        """(mean, std) of the additive disturbance.  Prior branch of dynamics.py:381-390: zero mean, MAX_STD."""
        is_tensor = torch.is_tensor(test_x)
        expand_dims = len(test_x.shape) == 1
        if expand_dims:
            test_x = test_x[None]
        if self.disturbance_fn is not None:
            # documented contract: disturbance_fn(state_batch ndarray) -> (mean, std).  It is called with an ndarray
            # whatever the caller passed, and its outputs come back in the caller's kind (tensor -> same device / dtype)
            xs = test_x.detach().cpu().numpy() if is_tensor else np.asarray(test_x)
            means, f_std = self.disturbance_fn(xs)
            if is_tensor:
                means = torch.as_tensor(means).to(test_x.device, test_x.dtype)
                f_std = torch.as_tensor(f_std).to(test_x.device, test_x.dtype)
            else:
                means = means.detach().cpu().numpy() if torch.is_tensor(means) else np.asarray(means)
                f_std = f_std.detach().cpu().numpy() if torch.is_tensor(f_std) else np.asarray(f_std)
        elif self.disturb_estimators:
            # dynamics.py:371-379: x / std_x, every GP, mean * (std_y + 1e-8), sqrt(f_var) * (std_y + 1e-8) -- the
            # scalings live in the bank and are applied inside the kernel
            if is_tensor:
                dt_ = test_x.dtype if test_x.dtype in (torch.float32, torch.float64) else torch.float32
                means, f_std = self._gp_bank.predict(test_x.detach().to(self.device, dt_))
                means, f_std = means.to(test_x.device, test_x.dtype), f_std.to(test_x.device, test_x.dtype)
            else:
                xs = torch.as_tensor(np.asarray(test_x, np.float64)).to(self.device)
                means, f_std = (v.cpu().numpy() for v in self._gp_bank.predict(xs))
        else:
            max_std = MAX_STD[self.env.dynamics_mode]
            if is_tensor:
                means = torch.zeros_like(test_x)
                f_std = torch.as_tensor(max_std, dtype=test_x.dtype, device=test_x.device).expand_as(test_x).clone()
            else:
                means = np.zeros(test_x.shape)
                f_std = np.ones(test_x.shape) * np.asarray(max_std)
        if expand_dims:
            means, f_std = means[0], f_std[0]
        return means, f_std

    def append_transition(self, state_batch, u_batch, next_state_batch, t_batch=None):
        """Record (state, estimated disturbance) in the ring buffer (dynamics.py:263-304) and refit the GPs every
        max_history_count / 10 points, like the reference."""
        expand_dims = len(state_batch.shape) == 1
        if expand_dims:
            state_batch = np.expand_dims(state_batch, 0)
            next_state_batch = np.expand_dims(next_state_batch, 0)
            u_batch = np.expand_dims(u_batch, 0)
        if self.env.dynamics_mode == 'SimulatedCars' and t_batch is None:
            t_batch = np.zeros(state_batch.shape[0])
        prior_next, _, _ = self.predict_next_state(np.asarray(state_batch, np.float64), u_batch, t_batch, use_gps=False)
        disturbance_batch = (next_state_batch - prior_next) / self.env.dt
        for i in range(state_batch.shape[0]):
            k = self.history_counter % self.max_history_count
            self.disturbance_history['state'][k] = state_batch[i]
            self.disturbance_history['disturbance'][k] = disturbance_batch[i]
            self.history_counter += 1
            if self.history_counter % (self.max_history_count / 10) == 0:
                self.fit_gp_model()

    def fit_gp_model(self, training_iter=70):
        """dynamics.py:306-340: normalise the history by its std (+1e-8), one GP per state dimension with the MAX_STD
        prior outputscale, `training_iter` Adam steps on the exact marginal likelihood -- all dimensions as one batch
        on the device -- then cache the posterior factors the predict kernel reads."""
        if self.history_counter < self.max_history_count:
            train_x = self.disturbance_history['state'][:self.history_counter]
            train_y = self.disturbance_history['disturbance'][:self.history_counter]
        else:
            train_x = self.disturbance_history['state']
            train_y = self.disturbance_history['disturbance']
        self._install_gp_bank(np.array(train_x), np.array(train_y))
        self._gp_bank.train(training_iter)
        self._gp_bank.build_posterior()

    def _install_gp_bank(self, train_x, train_y):
        from .gp_model import DisturbanceGPBank, BankMember
        train_x_std = np.std(train_x, axis=0)
        train_y_std = np.std(train_y, axis=0)
        with np.errstate(divide='ignore'):
            x_scale = train_x_std + 0.0                      # predict divides by std_x WITHOUT the 1e-8 (dynamics.py:375)
        self._gp_bank = DisturbanceGPBank(train_x / (train_x_std + 1e-8), train_y / (train_y_std + 1e-8),
                                          MAX_STD[self.env.dynamics_mode], device=self.device, x_scale=x_scale,
                                          y_scale=train_y_std + 1e-8)
        self.disturb_estimators = [BankMember(self._gp_bank, i) for i in range(self.n_s)]
        self.train_x = train_x
        self.train_y = train_y

    def load_disturbance_models(self, output):
        """dynamics.py:393-410.  Reads this class's own files and the reference's (a list of gpytorch state dicts).
        Deliberate difference: the reference rebuilds its GPs on the RAW saved history here (:403) although it fitted
        them on the normalised one (:326-334) and keeps normalising the queries (:375), so its loaded models are not the
        models it saved; this loader restores what was saved (normalised history, as at fit time)."""
        if output is None:
            return
        try:
            weights = torch.load('{}/gp_models.pkl'.format(output), map_location='cpu')
            train_x = torch.load('{}/gp_models_train_x.pkl'.format(output), weights_only=False)
            train_y = torch.load('{}/gp_models_train_y.pkl'.format(output), weights_only=False)
            self._install_gp_bank(np.asarray(train_x, np.float64), np.asarray(train_y, np.float64))
            import os
            raw64 = '{}/gp_models_raw_f64.pkl'.format(output)
            raw64 = torch.load(raw64, map_location='cpu') if os.path.exists(raw64) else None
            for i in range(self.n_s):
                self.disturb_estimators[i].load_state_dict(weights[i] if raw64 is None else {"raw_float64": raw64[i]})
            self._gp_bank.build_posterior()
        except Exception:
            raise Exception('Could not load GP models from {}'.format(output))

    def save_disturbance_models(self, output):
        """dynamics.py:412-423: the same three files; gp_models.pkl holds one gpytorch-keyed state dict per GP (float32,
        see BankMember.state_dict).  A fourth file, gp_models_raw_f64.pkl, keeps the float64 raw parameters this class
        trains in; the reference ignores it, this class prefers it on load."""
        if not self.disturb_estimators or self.train_x is None or self.train_y is None:
            return
        torch.save([est.state_dict() for est in self.disturb_estimators], '{}/gp_models.pkl'.format(output))
        torch.save(self._gp_bank.raw.detach().cpu().clone(), '{}/gp_models_raw_f64.pkl'.format(output))
        torch.save(self.train_x, '{}/gp_models_train_x.pkl'.format(output))
        torch.save(self.train_y, '{}/gp_models_train_y.pkl'.format(output))

    def seed(self, s):
        torch.manual_seed(s)
        if torch.cuda.is_available():
            torch.cuda.manual_seed(s)
