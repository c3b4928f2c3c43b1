"""Batched SimulatedCarsEnv on the GPU -- drop-in for envs/simulated_cars_env.py (gym old 4-tuple API).

Front <- Car 1 <- Car 2 <- Car 3 <- Car 4 (controlled) <- Car 5.  num_envs == 1 reproduces the reference's
single-instance contract on the float64 kernel (reset draws its velocity noise from numpy's global RNG exactly like
simulated_cars_env.py:118); num_envs > 1 is device-resident float32.
"""
import numpy as np
import torch

from .. import _lib, _params
from ..diff_cbf_qp import _f32c
from ..spaces import Box


class SimulatedCarsEnv:
    metadata = {'render.modes': ['human']}

    def __init__(self, num_envs=1, device=None, precision=None, auto_reset=False, seed=None):
        _lib.require_cuda()
        self._lib = _lib.load()
        self.num_envs = int(num_envs)
        self.device = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
        if self.device.type == "cuda" and self.device.index is None:      # "cuda" -> the current device, explicitly
            self.device = torch.device("cuda", torch.cuda.current_device())
        self.precision = precision or ("f64" if self.num_envs == 1 else "f32")
        self._dtype = torch.float64 if self.precision == "f64" else torch.float32
        self.auto_reset = bool(auto_reset)

        self.dynamics_mode = 'SimulatedCars'
        self.action_space = Box(low=-1.0, high=1.0, shape=(1,))               # simulated_cars_env.py:18
        self.safe_action_space = Box(low=-10.0, high=10.0, shape=(1,))        # :19
        self.observation_space = Box(low=-1e10, high=1e10, shape=(10,))       # :20
        self.max_episode_steps = 300
        self.dt = 0.02
        self.kp = 4.0
        self.k_brake = 20.0
        self.disturb_mean = np.zeros((1,))
        self.disturb_covar = np.diag([0.2 ** 2])

        n = self.num_envs
        self._state = torch.zeros((n, 10), dtype=self._dtype, device=self.device)
        self._t = torch.zeros((n,), dtype=self._dtype, device=self.device)
        self._step = torch.zeros((n,), dtype=torch.int32, device=self.device)
        # step outputs: typed views into ONE device buffer (single-env step = one device->host copy)
        # (single-instance float64 env: pinned HOST buffers the kernel reads / writes directly, like UnicycleEnv)
        isz = 8 if self._dtype == torch.float64 else 4
        self._mapped = (n == 1 and self.precision == "f64")
        self._outbuf = torch.zeros(((12 * isz + 1) * n + 15) // 16 * 16, dtype=torch.uint8, device=self.device)
        if self._mapped:
            self._outbuf = torch.zeros_like(self._outbuf, device="cpu").pin_memory()
            self._out_np = self._outbuf.numpy()
            self._vals_np = self._out_np[:12 * isz].view(np.float64)
            self._act_pin = torch.zeros((1,), dtype=torch.float64).pin_memory()
            self._act_np = self._act_pin.numpy()
        typed = self._outbuf[:12 * isz * n].view(self._dtype)
        self._obs = typed[:10 * n].view(n, 10)
        self._reward = typed[10 * n:11 * n]
        self._cost = typed[11 * n:12 * n]
        self._done = self._outbuf[12 * isz * n:12 * isz * n + n]
        self._gen = torch.Generator(device=self.device)
        if seed is not None:
            self._gen.manual_seed(int(seed))
        self.reset()

    @property
    def unwrapped(self):
        return self

    def _env_params(self):
        key = (self.dt, self.kp, self.k_brake, self.max_episode_steps, self.auto_reset)
        cache = getattr(self, "_ep_cache", None)
        if cache is None or cache[0] != key:
            cache = self._ep_cache = (key, _params.cars_env_params(*key))
        return cache[1]

    def _enter_device(self):
        prev = torch.cuda.current_device()
        if prev != self.device.index:
            torch.cuda.set_device(self.device)
            return prev
        return None

    def _wait(self):
        _lib.check(self._lib.rcbf_stream_synchronize(_lib.stream_ptr(self.device)), "cudaStreamSynchronize")

    def _fn(self, name):
        return getattr(self._lib, "rcbf_cars_env_%s_%s" % (name, self.precision))

    @property
    def state(self):
        return self._state[0].double().cpu().numpy() if self.num_envs == 1 else self._state

    @state.setter
    def state(self, value):
        v = torch.as_tensor(np.asarray(value) if not torch.is_tensor(value) else value).to(self.device, self._dtype)
        self._state[:] = v.reshape(-1, 10)

    @property
    def t(self):
        return float(self._t[0].item()) if self.num_envs == 1 else self._t

    @property
    def episode_step(self):
        return int(self._step[0].item()) if self.num_envs == 1 else self._step

    def seed(self, s=None):
        self.action_space.seed(s)
        if s is not None:
            self._gen.manual_seed(int(s))
        return [s]

    def close(self):
        pass

    def render(self, mode='human', close=False):
        print('Ep_step = {}, \tState = {}'.format(self.episode_step, self.state))

    def reset(self, mask=None, v_noise=None):
        """All cars at (34, 28, 22, 16, 10), velocities 30 + ONE shared N(0, 0.5) draw per instance, car 4 at 35."""
        n = self.num_envs
        if v_noise is None:
            if n == 1:
                v_noise = torch.tensor([np.random.normal(0, 0.5)], dtype=self._dtype)   # global numpy RNG, like :118
            else:
                v_noise = 0.5 * torch.randn((n,), generator=self._gen, device=self.device, dtype=self._dtype)
        vn = torch.as_tensor(v_noise).to(self.device, self._dtype).reshape(n).contiguous()
        m = None if mask is None else mask.to(self.device, torch.uint8).contiguous()
        with torch.cuda.device(self.device):
            rc = self._fn("reset")(_lib.ptr(self._state), _lib.ptr(self._t), _lib.ptr(self._step), _lib.ptr(vn),
                                   _lib.ptr(m), n, _lib.ptr(self._obs), _lib.stream_ptr(self.device))
        _lib.check(rc, "rcbf_cars_env_reset")
        if self._mapped:
            self._wait()
            return self._vals_np[:10].copy()
        return self._obs[0].double().cpu().numpy() if n == 1 else self._obs.clone()

    def step(self, action):
        if self._mapped:
            if torch.is_tensor(action):
                action = action.detach().cpu().numpy()
            self._act_np[0] = np.asarray(action, np.float64).reshape(())
            a = self._act_pin
        elif torch.is_tensor(action):
            a = action.detach().to(self.device, self._dtype).reshape(self.num_envs).contiguous()
        else:
            a = torch.as_tensor(np.asarray(action, np.float64).reshape(self.num_envs)).to(self.device, self._dtype)
        prev = self._enter_device()
        rc = self._fn("step")(_lib.ptr(self._state), _lib.ptr(self._t), _lib.ptr(self._step), _lib.ptr(a),
                              self.num_envs, self._env_params(), _lib.ptr(self._obs), _lib.ptr(self._reward),
                              _lib.ptr(self._done), _lib.ptr(self._cost), _lib.stream_ptr(self.device))
        if prev is not None:
            torch.cuda.set_device(prev)
        _lib.check(rc, "rcbf_cars_env_step")
        if self.num_envs == 1:
            isz = 8 if self._dtype == torch.float64 else 4
            if self._mapped:
                self._wait()                               # the kernel wrote straight into the pinned host buffer
                vals, flags = self._vals_np, self._out_np
            else:
                host = self._outbuf.cpu()                  # ONE device->host copy (and the only synchronisation)
                vals, flags = host[:12 * isz].view(self._dtype).double().numpy(), host.numpy()
            info = {'cost': float(vals[11]), 'goal_met': False}                 # simulated_cars_env.py:85
            return vals[:10].copy(), float(vals[10]), bool(flags[12 * isz]), info
        info = {'cost': self._cost.clone(), 'goal_met': torch.zeros_like(self._done, dtype=torch.bool)}
        return self._obs.clone(), self._reward.clone(), self._done.bool(), info

    def safe_step(self, cbf_layer, action_rl, sigma_pred, want_status=False):
        """Fused K5 (float32 layout): get_safe_action + step in one launch.  Returns (safe_action, obs, reward, done, info)."""
        if self.precision != "f32":
            raise ValueError("safe_step runs on the float32 env layout (precision='f32')")
        dev = self.device
        n = self.num_envs
        ac, sg = _f32c(action_rl, dev), _f32c(sigma_pred, dev)
        if getattr(self, "_safe_action", None) is None:
            self._safe_action = torch.empty((n, 1), dtype=torch.float32, device=dev)
        if getattr(self, "_counters", None) is None:
            self._counters = torch.zeros(_params.WS_WORDS, dtype=torch.int64, device=dev)
        if getattr(self, "_own_key", None) != (id(self._safe_action), id(self._counters)):
            self._own_key = (id(self._safe_action), id(self._counters))
            self._own_ptrs = tuple(t.data_ptr() for t in (self._state, self._t, self._step, self._safe_action, self._obs,
                                                          self._reward, self._done, self._cost, self._counters))
        status = torch.empty((n,), dtype=torch.int32, device=dev) if want_status else None
        o = self._own_ptrs
        prev = self._enter_device()
        lp = cbf_layer._params()
        tok = cbf_layer._publish_arm(self._counters, lp) if cbf_layer.check_nan else None   # kernel publishes the counters
        rc = self._lib.rcbf_cars_safe_step(o[0], o[1], o[2], ac.data_ptr(), sg.data_ptr(), n, lp,
                                           self._env_params(), o[3], o[4], o[5], o[6], o[7], _lib.ptr(status), o[8],
                                           _lib.stream_ptr(dev))
        cbf_layer._publish_disarm(lp)
        if prev is not None:
            torch.cuda.set_device(prev)
        _lib.check(rc, "rcbf_cars_safe_step")
        cbf_layer._last_counters = self._counters      # layer.solver_stats() also covers fused steps (cumulative)
        cbf_layer._last_stats = None
        cbf_layer._check_fused_step(self, self._counters, tok)   # raises 'QP Failed to solve' like the reference (check_nan)
        info = {'cost': self._cost, 'status': status}
        return self._safe_action, self._obs, self._reward, self._done, info

    def safe_step_host(self, cbf_layer, action_rl, sigma_pred, out=None, chunks=8):
        """Fused safe step with HOST tensors in and out (see UnicycleEnv.safe_step_host)."""
        import ctypes as C
        if self.precision != "f32":
            raise ValueError("safe_step_host runs on the float32 env layout (precision='f32')")
        n = self.num_envs
        if out is None:
            mk = lambda shape, dt: torch.empty(shape, dtype=dt).pin_memory()  # noqa: E731
            out = dict(safe_action=mk((n, 1), torch.float32), obs=mk((n, 10), torch.float32),
                       reward=mk((n,), torch.float32), done=mk((n,), torch.uint8), cost=mk((n,), torch.float32))
        for t in (action_rl, sigma_pred):
            if t.is_cuda or t.dtype != torch.float32 or not t.is_contiguous():
                raise ValueError("safe_step_host takes contiguous float32 CPU tensors")
        nf = C.c_int32(0)
        torch.cuda.current_stream(self.device).synchronize()
        rc = self._lib.rcbf_cars_safe_step_host(
            _lib.ptr(self._state), _lib.ptr(self._t), _lib.ptr(self._step), _lib.ptr(action_rl), _lib.ptr(sigma_pred), n,
            cbf_layer._params(), self._env_params(), _lib.ptr(out["safe_action"]), _lib.ptr(out["obs"]),
            _lib.ptr(out["reward"]), _lib.ptr(out["done"]), _lib.ptr(out["cost"]), C.byref(nf), self.device.index or 0,
            int(chunks))
        _lib.check(rc, "rcbf_cars_safe_step_host")
        if cbf_layer.check_nan and nf.value > 0:
            raise Exception('QP Failed to solve')
        return out
