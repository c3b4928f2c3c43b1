"""DynamicsModel -- prior-dynamics part of rcbf_sac/dynamics.py behind the same method names.

In scope: `predict_next_state` (prior f, g + dt * disturbance mean), `get_state`,
`get_obs`, the zero-mean / MAX_STD prior branch of `predict_disturbance`, `append_transition` (history ring buffer).
Out of scope: fitting / evaluating the GPyTorch disturbance GPs (dynamics.py:306-340,371-379) -- the kernels consume
the (mean, std) tensors a GP produces; plug one in through `disturbance_fn`.
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
            means, f_std = self.disturbance_fn(test_x)
        elif self.disturb_estimators:
            raise NotImplementedError("fitted-GP prediction (dynamics.py:371-379) is out of scope; pass disturbance_fn")
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
        """Record (state, estimated disturbance) in the ring buffer (dynamics.py:263-304).  GP refits are delegated
        to `fit_gp_model`, a no-op unless a subclass / hook provides one."""
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
        """GP fitting (dynamics.py:306-340) is outside the hot-path scope: no-op."""
        return None

    def seed(self, s):
        torch.manual_seed(s)
        if torch.cuda.is_available():
            torch.cuda.manual_seed(s)
