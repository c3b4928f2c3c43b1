"""Batched UnicycleEnv on the GPU -- drop-in for envs/unicycle_env.py (gym old 4-tuple API).

num_envs == 1 reproduces the single-instance contract of the reference exactly (ndarray obs, python float reward, bool
done, info dict with 'goal_met' only when met and 'cost' only inside a hazard) and runs the float64 kernel, which
follows the numpy arithmetic of the reference operation by operation.  num_envs > 1 keeps everything on the device
(torch tensors in / out) and defaults to the float32 layout: state = one float4 (x, y, theta, last_goal_dist).
Rendering (unicycle_env.py:145-213) is out of scope.
"""
import numpy as np
import torch

from .. import _lib, _params
from ..diff_cbf_qp import _f32c
from ..spaces import Box


class UnicycleEnv:
    metadata = {'render.modes': ['human']}

    def __init__(self, num_envs=1, device=None, precision=None, auto_reset=False):
        _lib.require_cuda()
        self._lib = _lib.load()
        self.num_envs = int(num_envs)
        self.device = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
        if self.device.type == "cuda" and self.device.index is None:      # "cuda" -> the current device, explicitly
            self.device = torch.device("cuda", torch.cuda.current_device())
        self.precision = precision or ("f64" if self.num_envs == 1 else "f32")
        self._dtype = torch.float64 if self.precision == "f64" else torch.float32
        self.auto_reset = bool(auto_reset)

        self.dynamics_mode = 'Unicycle'
        self.action_space = Box(low=-1.0, high=1.0, shape=(2,))               # unicycle_env.py:21
        self.safe_action_space = Box(low=-2.5, high=2.5, shape=(2,))          # :22
        self.observation_space = Box(low=-1e10, high=1e10, shape=(7,))        # :23
        self.bds = np.array([[-3., -3.], [3., 3.]])                           # :24
        self.hazards_radius = 0.6                                             # :25
        self.hazards_locations = np.array([[0., 0.], [-1., 1.], [-1., -1.], [1., -1.], [1., 1.]]) * 1.5   # :26
        self.dt = 0.02
        self.max_episode_steps = 1000
        self.reward_goal = 1.0
        self.goal_size = 0.3
        self.goal_pos = np.array([2.5, 2.5])
        self.initial_state = np.array([-2.5, -2.5, 0.])                       # :137
        self.viewer = None

        n = self.num_envs
        self._state4 = torch.zeros((n, 4), dtype=self._dtype, device=self.device)
        self._step = torch.zeros((n,), dtype=torch.int32, device=self.device)
        # step outputs: typed views into ONE device buffer, so that the single-env gym contract (obs ndarray, float
        # reward, bool done, info dict -- main.py:93-95) costs one device->host copy per step instead of five
        # The single-instance float64 env (the reference's training loop, main.py:93-95) goes one step further: its
        # action / output buffers are PINNED HOST memory the kernel reads and writes directly (page-locked memory is
        # device-accessible under unified addressing), so a step is one launch + one stream wait, no copy at all.
        isz = 8 if self._dtype == torch.float64 else 4
        self._mapped = (n == 1 and self.precision == "f64")
        self._outbuf = torch.zeros(((9 * isz + 2) * n + 15) // 16 * 16, dtype=torch.uint8, device=self.device)
        if self._mapped:
            self._outbuf = torch.zeros_like(self._outbuf, device="cpu").pin_memory()
            self._out_np = self._outbuf.numpy()
            self._vals_np = self._out_np[:9 * isz].view(np.float64)
            self._act_pin = torch.zeros((1, 2), dtype=torch.float64).pin_memory()
            self._act_np = self._act_pin.numpy()
        typed = self._outbuf[:9 * isz * n].view(self._dtype)
        self._obs = typed[:7 * n].view(n, 7)
        self._reward = typed[7 * n:8 * n]
        self._cost = typed[8 * n:9 * n]
        self._done = self._outbuf[9 * isz * n:9 * isz * n + n]
        self._goal = self._outbuf[9 * isz * n + n:9 * isz * n + 2 * n]
        self.reset()

    # -------------------------------------------------------------------------------------------- helpers
    @property
    def unwrapped(self):
        return self

    def _env_params(self):
        """C parameter struct of the env, rebuilt only when an attribute it mirrors was re-assigned (identity of the
        arrays, value of the scalars); an IN-PLACE edit of an array attribute is caught by a content comparison every
        256th call."""
        quick = (id(self.hazards_locations), self.hazards_radius, self.dt, id(self.goal_pos), self.goal_size,
                 self.reward_goal, id(self.initial_state), self.max_episode_steps, self.auto_reset)
        cache = getattr(self, "_ep_cache", None)
        self._ep_calls = getattr(self, "_ep_calls", 0) + 1
        if cache is not None and cache[0] == quick and self._ep_calls & 255:
            return cache[2]
        content = (np.asarray(self.hazards_locations, np.float64).tobytes(), np.asarray(self.goal_pos, np.float64).tobytes(),
                   np.asarray(self.initial_state, np.float64).tobytes())
        if cache is not None and cache[0] == quick and cache[1] == content:
            return cache[2]
        e = _params.unicycle_env_params(self.hazards_locations, self.hazards_radius, self.dt, self.goal_pos,
                                        self.goal_size, self.reward_goal, self.initial_state,
                                        self.max_episode_steps, self.auto_reset)
        self._ep_cache = (quick, content, e)
        return e

    def _enter_device(self):
        """Make the env's device current for a launch; returns the previous one when a switch was needed."""
        prev = torch.cuda.current_device()
        if prev != self.device.index:
            torch.cuda.set_device(self.device)
            return prev
        return None

    def _wait(self):
        """Block until the launches on the current stream have finished (the mapped single-instance buffers are read by
        the host right after)."""
        _lib.check(self._lib.rcbf_stream_synchronize(_lib.stream_ptr(self.device)), "cudaStreamSynchronize")

    def _fn(self, name):
        return getattr(self._lib, "rcbf_unicycle_env_%s_%s" % (name, self.precision))

    @property
    def state(self):
        """(3,) float64 ndarray for num_envs == 1 (like the reference attribute), else a (N,3) device tensor view."""
        if self.num_envs == 1:
            return self._state4[0, :3].double().cpu().numpy()
        return self._state4[:, :3]

    @state.setter
    def state(self, value):
        v = torch.as_tensor(np.asarray(value) if not torch.is_tensor(value) else value).to(self.device, self._dtype)
        self._state4[:, :3] = v.reshape(-1, 3)
        gp = torch.as_tensor(self.goal_pos, dtype=self._dtype, device=self.device)
        self._state4[:, 3] = torch.linalg.norm(gp - self._state4[:, :2], dim=1)

    @property
    def episode_step(self):
        return int(self._step[0].item()) if self.num_envs == 1 else self._step

    @property
    def last_goal_dist(self):
        return float(self._state4[0, 3].item()) if self.num_envs == 1 else self._state4[:, 3]

    def seed(self, s=None):
        self.action_space.seed(s)
        return [s]

    def close(self):
        pass

    def render(self, mode='human', close=False):
        raise NotImplementedError("rendering (envs/unicycle_env.py:145-213) is outside the hot-path scope")

    # -------------------------------------------------------------------------------------------- API
    def reset(self, mask=None):
        """Reset all instances (or those where mask != 0) to (-2.5, -2.5, 0); returns the observation."""
        m = None if mask is None else mask.to(self.device, torch.uint8).contiguous()
        e = self._env_params()
        with torch.cuda.device(self.device):
            rc = self._fn("reset")(_lib.ptr(self._state4), _lib.ptr(self._step), _lib.ptr(m), self.num_envs, e,
                                   _lib.ptr(self._obs), _lib.stream_ptr(self.device))
        _lib.check(rc, "rcbf_unicycle_env_reset")
        return self.get_obs_from_buffer()

    def get_obs_from_buffer(self):
        if self._mapped:
            self._wait()
            return self._vals_np[:7].copy()
        if self.num_envs == 1:
            return self._obs[0].double().cpu().numpy()
        return self._obs.clone()

    def get_obs(self):
        """Observation of the current state (unicycle_env.py:215-231)."""
        e = self._env_params()
        zero = torch.zeros((self.num_envs,), dtype=torch.uint8, device=self.device)
        with torch.cuda.device(self.device):
            rc = self._fn("reset")(_lib.ptr(self._state4), _lib.ptr(self._step), _lib.ptr(zero), self.num_envs, e,
                                   _lib.ptr(self._obs), _lib.stream_ptr(self.device))
        _lib.check(rc, "rcbf_unicycle_env_reset(obs only)")
        return self.get_obs_from_buffer()

    def step(self, action):
        """(obs, reward, done, info).  The action is clipped to [-1, 1] inside the kernel (unicycle_env.py:62)."""
        if self._mapped:
            if torch.is_tensor(action):
                action = action.detach().cpu().numpy()
            self._act_np[0] = np.asarray(action, np.float64).reshape(2)
            a = self._act_pin
        elif torch.is_tensor(action):
            a = action.detach().to(self.device, self._dtype).reshape(self.num_envs, 2).contiguous()
        else:
            a = torch.as_tensor(np.asarray(action, np.float64).reshape(self.num_envs, 2)).to(self.device, self._dtype)
        e = self._env_params()
        prev = self._enter_device()
        rc = self._fn("step")(_lib.ptr(self._state4), _lib.ptr(self._step), _lib.ptr(a), self.num_envs, e,
                              _lib.ptr(self._obs), _lib.ptr(self._reward), _lib.ptr(self._done),
                              _lib.ptr(self._cost), _lib.ptr(self._goal), _lib.stream_ptr(self.device))
        if prev is not None:
            torch.cuda.set_device(prev)
        _lib.check(rc, "rcbf_unicycle_env_step")
        return self._pack_step_outputs()

    def _pack_step_outputs(self):
        if self.num_envs == 1:
            isz = 8 if self._dtype == torch.float64 else 4
            if self._mapped:
                self._wait()                               # the kernel wrote straight into the pinned host buffer
                vals, flags = self._vals_np, self._out_np
            else:
                host = self._outbuf.cpu()                  # ONE device->host copy (and the only synchronisation)
                vals, flags = host[:9 * isz].view(self._dtype).double().numpy(), host.numpy()
            done, goal = bool(flags[9 * isz]), bool(flags[9 * isz + 1])
            info = dict()
            if goal:
                info['goal_met'] = True                    # only present when met (unicycle_env.py:98)
            c = float(vals[8])
            if c != 0.0:
                info['cost'] = c                           # only present inside a hazard (:106-110)
            return vals[:7].copy(), float(vals[7]), done, info
        info = {'cost': self._cost.clone(), 'goal_met': self._goal.bool()}
        return self._obs.clone(), self._reward.clone(), self._done.bool(), info

    def safe_step(self, cbf_layer, action_rl, mean_pred, sigma_pred, want_status=False):
        """Fused K5: CBFQPLayer.get_safe_action + step in ONE kernel launch (float32 layout only).
        Returns (safe_action, obs, reward, done, info)."""
        if self.precision != "f32":
            raise ValueError("safe_step runs on the float32 env layout (precision='f32')")
        lp = cbf_layer._params()
        if cbf_layer._general_hz is not None:
            raise ValueError("the fused step is specialised for up to 5 hazards (unicycle_env.py:26); with %d use "
                             "cbf_layer.get_safe_action + env.step" % cbf_layer._general_hz.shape[0])
        dev = self.device
        ac, mu, sg = _f32c(action_rl, dev), _f32c(mean_pred, dev), _f32c(sigma_pred, dev)
        n = self.num_envs
        if getattr(self, "_safe_action", None) is None:
            self._safe_action = torch.empty((n, 2), dtype=torch.float32, device=dev)
        if getattr(self, "_counters", None) is None:
            self._counters = torch.zeros(_params.WS_WORDS, dtype=torch.int64, device=dev)
        if getattr(self, "_own_key", None) != (id(self._safe_action), id(self._counters)):
            # (pointers of the env-owned buffers only change when a buffer object is replaced: looked up once)
            self._own_key = (id(self._safe_action), id(self._counters))
            self._own_ptrs = tuple(t.data_ptr() for t in (self._state4, self._step, self._safe_action, self._obs,
                                                          self._reward, self._done, self._cost, self._goal,
                                                          self._counters))
        status = torch.empty((n,), dtype=torch.int32, device=dev) if want_status else None
        o = self._own_ptrs
        prev = self._enter_device()
        tok = cbf_layer._publish_arm(self._counters, lp) if cbf_layer.check_nan else None   # kernel publishes the counters
        rc = self._lib.rcbf_unicycle_safe_step(o[0], o[1], ac.data_ptr(), mu.data_ptr(), sg.data_ptr(), n,
                                               lp, self._env_params(), o[2], o[3], o[4], o[5], o[6],
                                               o[7], _lib.ptr(status), o[8], _lib.stream_ptr(dev))
        cbf_layer._publish_disarm(lp)
        if prev is not None:
            torch.cuda.set_device(prev)
        _lib.check(rc, "rcbf_unicycle_safe_step")
        cbf_layer._last_counters = self._counters      # layer.solver_stats() also covers fused steps (cumulative)
        cbf_layer._last_stats = None
        cbf_layer._check_fused_step(self, self._counters, tok)   # raises 'QP Failed to solve' like the reference (check_nan)
        info = {'cost': self._cost, 'goal_met': self._goal, 'status': status}
        return self._safe_action, self._obs, self._reward, self._done, info

    HOST_OUTPUTS = ("safe_action", "obs", "reward", "done", "cost", "goal_met")

    def safe_step_host(self, cbf_layer, action_rl, mean_pred=None, sigma_pred=None, out=None, chunks=8, outputs=None,
                       gp=None):
        """Fused safe step with HOST tensors in and out (the end-to-end path): `action_rl (n,2)`, `mean_pred (n,3)`,
        `sigma_pred (n,3)` float32 CPU tensors (pinned memory recommended); the env state stays on the GPU.  The C
        library pipelines H2D / kernel / D2H over `chunks` slices on its own streams and returns when `out` is valid.

        `outputs`: the subset of HOST_OUTPUTS the caller wants on the host (default: all six); the others are produced
        on the device but not copied back.  `gp`: a fitted `DisturbanceGPBank` (3 inputs, 3 outputs) -- the disturbance
        mean / std are then evaluated ON THE DEVICE from the resident state (rcbf_sac/sac_cbf.py:230-236) and
        `mean_pred` / `sigma_pred` must be None: the only host input is the action.
        Returns a dict of pinned CPU tensors (reused when passed back as `out`)."""
        import ctypes as C
        if self.precision != "f32":
            raise ValueError("safe_step_host runs on the float32 env layout (precision='f32')")
        n = self.num_envs
        names = self.HOST_OUTPUTS if outputs is None else tuple(outputs)
        for k in names:
            if k not in self.HOST_OUTPUTS:
                raise ValueError("unknown host output %r" % (k,))
        shapes = dict(safe_action=((n, 2), torch.float32), obs=((n, 7), torch.float32), reward=((n,), torch.float32),
                      done=((n,), torch.uint8), cost=((n,), torch.float32), goal_met=((n,), torch.uint8))
        if out is None:
            out = {}
        for k in names:
            if k not in out:
                out[k] = torch.empty(shapes[k][0], dtype=shapes[k][1]).pin_memory()
        ins = (action_rl,) if gp is not None else (action_rl, mean_pred, sigma_pred)
        if gp is not None and (mean_pred is not None or sigma_pred is not None):
            raise ValueError("with gp= the disturbance is evaluated on the device: pass mean_pred = sigma_pred = None")
        for t in ins:
            if t is None or t.is_cuda or t.dtype != torch.float32 or not t.is_contiguous():
                raise ValueError("safe_step_host takes contiguous float32 CPU tensors")
        nf = C.c_int32(0)
        torch.cuda.current_stream(self.device).synchronize()      # the library uses its own streams
        o = [_lib.ptr(out[k]) if k in names else None for k in self.HOST_OUTPUTS]
        if gp is None:
            rc = self._lib.rcbf_unicycle_safe_step_host(
                _lib.ptr(self._state4), _lib.ptr(self._step), _lib.ptr(action_rl), _lib.ptr(mean_pred),
                _lib.ptr(sigma_pred), n, cbf_layer._params(), self._env_params(), o[0], o[1], o[2], o[3], o[4], o[5],
                C.byref(nf), self.device.index or 0, int(chunks))
        else:
            if gp._post is None:
                gp.build_posterior()
            rc = self._lib.rcbf_unicycle_safe_step_host_gp(
                _lib.ptr(self._state4), _lib.ptr(self._step), _lib.ptr(action_rl), C.byref(gp._post[0]), n,
                cbf_layer._params(), self._env_params(), o[0], o[1], o[2], o[3], o[4], o[5], C.byref(nf),
                self.device.index or 0, int(chunks))
        _lib.check(rc, "rcbf_unicycle_safe_step_host")
        if cbf_layer.check_nan and nf.value > 0:
            raise Exception('QP Failed to solve')
        return out

    # kept for API parity with envs/unicycle_env.py:113,260
    def goal_met(self):
        d = torch.linalg.norm(torch.as_tensor(self.goal_pos, dtype=self._dtype, device=self.device)
                              - self._state4[:, :2], dim=1) <= self.goal_size
        return bool(d[0].item()) if self.num_envs == 1 else d

    def obs_compass(self):
        o = self.get_obs()
        return o[4:6] if self.num_envs == 1 else o[:, 4:6]
