"""CBFQPLayer -- drop-in for rcbf_sac/diff_cbf_qp.py:10-395 backed by the sm_100a kernels.

Same constructor, attributes, method names, argument order, shapes, dtypes and exceptions as the reference class;
`get_safe_action` is differentiable w.r.t. `action_batch` (the only input that carries grad in the reference, see
rcbf_sac/sac_cbf.py:233-236).  Compute always runs on the CUDA device; tensors that live elsewhere are copied in and
the result is returned on the caller's device.  There is no CPU path.
"""
import weakref

import numpy as np
import torch

from . import _lib, _params

DYNAMICS_MODE = {"Unicycle": {"n_s": 3, "n_u": 2}, "SimulatedCars": {"n_s": 10, "n_u": 1}}  # dynamics.py:22-23


def _f32c(t, device):
    """float32, contiguous, on `device`, detached -- returned as is when it already is all of that (the common case
    costs no torch dispatch)."""
    if t.dtype == torch.float32 and t.device == device and not t.requires_grad and t.is_contiguous():
        return t
    return t.detach().to(device=device, dtype=torch.float32).contiguous()


class _SafeActionFn(torch.autograd.Function):
    """final = clamp(a + QP(a)) with the implicit-KKT backward kernel (K4).  Only entered when the action carries grad;
    the forward saves one int32 per instance (status + active set), the backward rebuilds the rest."""

    @staticmethod
    def forward(ctx, layer, state, action, mean, sigma):
        dev = layer.device
        st, ac, sg = _f32c(state, dev), _f32c(action, dev), _f32c(sigma, dev)
        mu = _f32c(mean, dev)
        out, meta = layer._forward_meta(st, ac, mu, sg)
        ctx.layer = layer
        ctx.in_device = action.device
        ctx.in_dtype = action.dtype
        ctx.save_for_backward(st, ac, mu, sg, meta)
        return out.to(device=action.device, dtype=action.dtype) if (action.device != dev or action.dtype != torch.float32) else out

    @staticmethod
    def backward(ctx, grad_out):
        layer = ctx.layer
        st, ac, mu, sg, meta = ctx.saved_tensors
        go = _f32c(grad_out, layer.device)
        ga = layer._backward_meta(st, ac, mu, sg, meta, go)
        return None, None, ga.to(device=ctx.in_device, dtype=ctx.in_dtype), None, None


_QP_ROWS = {3: (9, 12, 16), 2: (4,)}     # row counts rcbf_qp_solve is instantiated for, per number of variables


class _QPFn(torch.autograd.Function):
    """Generic batched QP (cbf_layer API): float64 in, float64 out, qpth-style gradients for Q, p, G, h."""

    @staticmethod
    def forward(ctx, layer, Q, p, G, h):
        lib = _lib.load()
        dev = layer.device
        n, m_in, nz = G.shape
        Qd, pd, Gd, hd = (t.detach().to(device=dev, dtype=torch.float64).contiguous() for t in (Q, p, G, h))
        # the kernel is instantiated for a few row counts; a smaller system (a layer on K != 5 hazards: m = K + 4) is
        # padded with copies of its first row -- a duplicate constraint changes neither the feasible set nor the optimum
        m = min([s_ for s_ in _QP_ROWS.get(nz, ()) if s_ >= m_in], default=m_in)
        if m > m_in:
            Gd = torch.cat((Gd, Gd[:, :1].expand(n, m - m_in, nz)), 1).contiguous()
            hd = torch.cat((hd, hd[:, :1].expand(n, m - m_in)), 1).contiguous()
        x = torch.empty((n, nz), dtype=torch.float64, device=dev)
        lam = torch.empty((n, m), dtype=torch.float64, device=dev)
        slack = torch.empty((n, m), dtype=torch.float64, device=dev)
        counters = torch.zeros(8, dtype=torch.int64, device=dev)
        with torch.cuda.device(dev):
            rc = lib.rcbf_qp_solve(_lib.ptr(Qd), _lib.ptr(pd), _lib.ptr(Gd), _lib.ptr(hd), n, nz, m, _lib.ptr(x),
                                   _lib.ptr(lam), _lib.ptr(slack), None, None, _lib.ptr(counters),
                                   _lib.stream_ptr(dev))
        if rc == -1:
            raise NotImplementedError("generic QP kernel is instantiated for nz = 3 with up to 16 rows and (nz, m) = "
                                      "(2, 4), got (%d, %d)" % (nz, m_in))
        _lib.check(rc, "rcbf_qp_solve")
        layer._last_counters = counters
        layer._last_stats = None
        ctx.layer = layer
        ctx.save_for_backward(Qd, Gd, x, lam, slack)
        ctx.in_devices = (Q.device, p.device, G.device, h.device)
        ctx.m_in = m_in
        return x.to(G.device)

    @staticmethod
    def backward(ctx, gx):
        lib = _lib.load()
        layer = ctx.layer
        dev = layer.device
        Qd, Gd, x, lam, slack = ctx.saved_tensors
        n, m, nz = Gd.shape
        g = gx.detach().to(device=dev, dtype=torch.float64).contiguous()
        dQ = torch.empty_like(Qd)
        dp = torch.empty((n, nz), dtype=torch.float64, device=dev)
        dG = torch.empty_like(Gd)
        dh = torch.empty((n, m), dtype=torch.float64, device=dev)
        with torch.cuda.device(dev):
            rc = lib.rcbf_qp_solve_bwd(_lib.ptr(Qd), _lib.ptr(Gd), _lib.ptr(x), _lib.ptr(lam), _lib.ptr(slack),
                                       _lib.ptr(g), n, nz, m, _lib.ptr(dQ), _lib.ptr(dp), _lib.ptr(dG), _lib.ptr(dh),
                                       _lib.stream_ptr(dev))
        _lib.check(rc, "rcbf_qp_solve_bwd")
        d = ctx.in_devices
        if ctx.m_in < m:     # padding rows were copies of row 0: their (inactive-duplicate) gradient belongs to it
            dG[:, 0] += dG[:, ctx.m_in:].sum(1)
            dh[:, 0] += dh[:, ctx.m_in:].sum(1)
            dG, dh = dG[:, :ctx.m_in].contiguous(), dh[:, :ctx.m_in].contiguous()
        return None, dQ.to(d[0]), dp.to(d[1]), dG.to(d[2]), dh.to(d[3])


class CBFQPLayer:

    def __init__(self, env, args, gamma_b=100, k_d=1.5, l_p=0.03):
        """Constructor of CBFLayer (signature of rcbf_sac/diff_cbf_qp.py:12).

        Parameters
        ----------
        env : gym.env-like
            must expose dynamics_mode, safe_action_space, action_space and, per mode, hazards_locations /
            hazards_radius (Unicycle) or kp / k_brake (SimulatedCars)  (diff_cbf_qp.py:27-41,205-206,286)
        args : namespace with `.cuda` (diff_cbf_qp.py:25); optional `.device_num`
        gamma_b, k_d, l_p : as in the reference (k_d is stored but unused by this layer, like diff_cbf_qp.py:36/:261)
        """
        _lib.require_cuda()
        _lib.load()
        dev_num = getattr(args, "device_num", None)
        self.device = torch.device("cuda", torch.cuda.current_device() if dev_num is None else int(dev_num))
        # (the reference computes on the CPU when args.cuda is False; here the QP always runs on the GPU -- there is no
        #  CPU path -- and results are returned on the device / dtype of the caller's action tensor)

        self.env = env
        self.u_min, self.u_max = self.get_control_bounds()
        self.gamma_b = gamma_b

        if self.env.dynamics_mode not in DYNAMICS_MODE:
            raise Exception('Dynamics mode not supported.')

        if self.env.dynamics_mode == 'Unicycle':
            self.num_cbfs = len(env.hazards_locations)
            self.k_d = k_d
            self.l_p = l_p
        elif self.env.dynamics_mode == 'SimulatedCars':
            self.num_cbfs = 2

        self.action_dim = env.action_space.shape[0]
        self.num_ineq_constraints = self.num_cbfs + 2 * self.action_dim
        self.check_nan = True       # reference behaviour: sync + raise on NaN (diff_cbf_qp.py:141-143)
        # "presolve": greedy active-set guess + float64 KKT certificate, interior point only as fallback (default);
        # "pdipm": every non-trivial QP runs the primal-dual interior point (the reference's algorithm family)
        self.solver = "presolve"
        self._last_counters = None  # workspace of the last fused env step (device tensor, cumulative counters)
        self._last_stats = None     # per-call counter increments of the last synchronised launch
        self._params_cache = None
        self._general_hz = None     # (K, 2) float32 host table when the layer has more than 5 hazards

    def _workspace(self):
        """RCBF_WS_WORDS-word solver workspace (counters + fallback queue), zero-initialised ONCE.  The counters
        accumulate on the device; per-call numbers are differences taken on the host (`_sync_counters`), so a call
        costs no extra launch."""
        ws = getattr(self, "_ws", None)
        if ws is None or ws.device != self.device:
            ws = self._ws = torch.zeros(_params.WS_WORDS, dtype=torch.int64, device=self.device)
            self._ws_base = [0] * 8
            self._ws_mirror = None
        return ws

    def _publish_read(self, ws):
        """The 8 cumulative counters of workspace `ws` once everything enqueued so far on the current stream has finished.
        Low-latency form: a one-warp kernel (rcbf_counters_publish) copies the counters into a pinned host mirror and then
        stores a token there; this thread polls the token -- no cudaMemcpy, no cudaStreamSynchronize on the common
        path (a stream that stays busy for more than ~20 ms falls back to the blocking wait)."""
        mir = getattr(self, "_ws_mirror", None)
        if mir is None:
            self._ws_mirror_t = torch.zeros(9, dtype=torch.int64).pin_memory()
            mir = self._ws_mirror = self._ws_mirror_t.numpy()
            self._ws_token = 0
        self._ws_token += 1
        token = self._ws_token
        dev = self.device
        prev = torch.cuda.current_device()
        lib, stream = self._launch_ctx()
        rc = lib.rcbf_counters_publish(ws.data_ptr(), self._ws_mirror_t.data_ptr(), token, stream)
        if prev != dev.index:
            torch.cuda.set_device(prev)
        _lib.check(rc, "rcbf_counters_publish")
        spins = 0
        while mir[0] != token:
            spins += 1
            if spins > 100000:
                _lib.check(lib.rcbf_stream_synchronize(stream), "cudaStreamSynchronize")
                spins = 0
        return mir[1:9].tolist()

    # In-kernel publication (include/rcbf_b200.h: RCBF_SOLVER_PUBLISH): the last block of the step / layer kernel itself
    # copies the counters into a pinned host mirror bound to the workspace and stores this call's token there, so the
    # reference's per-call NaN test costs ONE launch + a poll instead of two launches.
    def _publish_arm(self, ws, p):
        """Request the publication for the launch that follows: -> token, or None when that launch cannot publish
        (interior-point mode, the general-hazard kernels, CUDA-graph capture) and `_publish_read` has to be used."""
        if self.solver != "presolve" or self._general_hz is not None or torch.cuda.is_current_stream_capturing():
            return None
        mirrors = self.__dict__.setdefault("_mirrors", {})
        key = ws.data_ptr()
        m = mirrors.get(key)
        if m is None or m[3]() is not ws:      # new workspace (possibly a fresh allocation at a recycled address)
            if len(mirrors) > 64:
                for k_ in [k_ for k_, v_ in mirrors.items() if v_[3]() is None]:
                    del mirrors[k_]
            t = torch.zeros(9, dtype=torch.int64).pin_memory()
            m = mirrors[key] = [t, t.numpy(), 0, weakref.ref(ws)]
            lib, stream = self._launch_ctx()
            _lib.check(lib.rcbf_counters_bind_mirror(key, t.data_ptr(), stream), "rcbf_counters_bind_mirror")
        m[2] = tok = (m[2] % 0x3fffff) + 1
        p.solver_mode = (p.solver_mode & 0xff) | 0x100 | (tok << 9)
        return tok

    @staticmethod
    def _publish_disarm(p):
        p.solver_mode &= 0xff

    def _publish_wait(self, ws, tok):
        """The 8 cumulative counters once the armed launch has published them (falls back to `_publish_read`)."""
        m = self._mirrors[ws.data_ptr()]
        mir = m[1]
        spins = 0
        while mir[0] != tok:
            spins += 1
            if spins > 200000:      # not published (e.g. the workspace was re-created at the same address): re-bind
                del self._mirrors[ws.data_ptr()]
                return self._publish_read(ws)
        return mir[1:9].tolist()

    def _sync_counters(self, tok=None):
        """Read the layer's own workspace (synchronises); returns this call's increments of the 8 counters."""
        cur = self._publish_read(self._ws) if tok is None else self._publish_wait(self._ws, tok)
        delta = [c - b for c, b in zip(cur, self._ws_base)]
        self._ws_base = cur
        self._last_stats = delta
        self._last_counters = None
        return delta

    def _check_fused_step(self, env, counters, tok=None):
        """NaN test of a fused env step (the reference raises on any NaN safe action, diff_cbf_qp.py:141-143): reads the
        env's workspace when `check_nan` is on (one host wait per step, like the reference's `.any()`); skipped while the
        stream is being captured into a CUDA graph and when `check_nan` is False (read `solver_stats()` instead)."""
        if not self.check_nan or torch.cuda.is_current_stream_capturing():
            return
        nan_now = (self._publish_read(counters) if tok is None else self._publish_wait(counters, tok))[0]
        seen = getattr(env, "_nan_seen", 0)
        if nan_now < seen:          # the workspace was re-zeroed by its owner
            env._nan_seen = seen = nan_now
        if nan_now > seen:
            env._nan_seen = nan_now
            print('\033[91m QP Failed to solve - result is nan == True!\033[00m')
            raise Exception('QP Failed to solve')

    def _before_launch(self):
        """Counters advanced by launches that were not read back (check_nan = False) must not be blamed on this call."""
        if self.check_nan and getattr(self, "_unread", False):
            self._sync_counters()
            self._unread = False

    def _after_launch(self, tok=None):
        self._last_stats = None
        self._last_counters = None
        if not self.check_nan:
            self._unread = True
            return
        if self._sync_counters(tok)[0] > 0:
            print('\033[91m QP Failed to solve - result is nan == True!\033[00m')
            raise Exception('QP Failed to solve')

    # ------------------------------------------------------------------------------------------------------ params
    def _solver_mode(self):
        if self.solver not in ("presolve", "pdipm"):
            raise ValueError("solver must be 'presolve' or 'pdipm', got %r" % (self.solver,))
        return 0 if self.solver == "presolve" else 1

    def _bounds_host(self):
        """Host copies of u_min / u_max, refreshed only when the tensors are replaced or written (no per-call sync)."""
        tag = (self.u_min.data_ptr(), self.u_min._version, self.u_max.data_ptr(), self.u_max._version)
        if getattr(self, "_bounds_tag", None) != tag:
            self._bounds_np = (self.u_min.detach().cpu().numpy().copy(), self.u_max.detach().cpu().numpy().copy())
            self._bounds_tag = tag
        return self._bounds_np

    def _params(self):
        """C parameter struct, rebuilt when a public attribute the reference reads at call time has changed."""
        env = self.env
        # fast path: nothing the struct depends on was re-assigned (identity / value of the scalars, identity of the
        # arrays; an IN-PLACE edit of env.hazards_locations is caught by the bytes comparison every 256th call)
        quick = (self.solver, self.gamma_b, getattr(self, "l_p", None), id(getattr(env, "hazards_locations", None)),
                 getattr(env, "hazards_radius", None), getattr(env, "kp", None), getattr(env, "k_brake", None),
                 self.u_min.data_ptr(), self.u_min._version, self.u_max.data_ptr(), self.u_max._version)
        self._params_calls = getattr(self, "_params_calls", 0) + 1
        if self._params_cache is not None and getattr(self, "_params_quick", None) == quick and self._params_calls & 255:
            return self._params_cache[1]
        self._params_quick = quick
        lo, hi = self._bounds_host()
        if env.dynamics_mode == 'Unicycle':
            hz = np.asarray(env.hazards_locations, np.float64)
            key = ('U', self.solver, float(self.gamma_b), float(self.l_p), float(env.hazards_radius), hz.tobytes(),
                   self._bounds_tag)
            if self._params_cache is None or self._params_cache[0] != key:
                # up to 5 hazards: the specialised hot kernels (fewer are padded with inert ones); 6 .. 12: the general
                # per-instance kernels of rcbf_general.cu, which take the hazard table as a separate host array
                general = hz.ndim == 2 and hz.shape[0] > _params.UNI_HAZ
                if general and (hz.shape[1] != 2 or hz.shape[0] > _params.MAX_HAZARDS):
                    raise ValueError("the sm_100a kernels support 1..%d hazards of shape (K, 2), got %r"
                                     % (_params.MAX_HAZARDS, hz.shape))
                if general and self.solver != "presolve":
                    raise ValueError("solver = 'pdipm' is only built for the 5-hazard kernels")
                p = _params.unicycle_params(None if general else hz, env.hazards_radius, float(self.gamma_b),
                                            float(self.l_p), lo, hi, solver_mode=self._solver_mode())
                self._general_hz = np.ascontiguousarray(hz, np.float32) if general else None
                self._params_cache = (key, p)
        else:
            key = ('C', self.solver, float(self.gamma_b), float(env.kp), float(env.k_brake), self._bounds_tag)
            if self._params_cache is None or self._params_cache[0] != key:
                p = _params.cars_params(float(self.gamma_b), float(env.kp), float(env.k_brake), float(lo[0]),
                                        float(hi[0]), solver_mode=self._solver_mode())
                self._params_cache = (key, p)
        return self._params_cache[1]

    # ------------------------------------------------------------------------------------------------- raw launches
    def _launch_ctx(self):
        """(lib, stream pointer) with the layer's device current (switching only when it is not already)."""
        dev = self.device
        if torch.cuda.current_device() != dev.index:
            torch.cuda.set_device(dev)
        return _lib.load(), _lib.stream_ptr(dev)

    def _forward_raw(self, st, ac, mu, sg, save=False, want_status=False):
        """Forward launch with the optional DENSE saved tensors x / lam / slack (diagnostics and the legacy backward
        entry); the autograd path uses `_forward_meta`."""
        dev = self.device
        n = st.shape[0]
        mode = self.env.dynamics_mode
        nz, m, nu = (3, 9, 2) if mode == 'Unicycle' else (2, 4, 1)
        out = torch.empty((n, nu), dtype=torch.float32, device=dev)
        x = lam = slack = status = iters = None
        if save:
            x = torch.empty((n, nz), dtype=torch.float32, device=dev)
            lam = torch.empty((n, m), dtype=torch.float32, device=dev)
            slack = torch.empty((n, m), dtype=torch.float32, device=dev)
        if want_status:
            status = torch.empty((n,), dtype=torch.int32, device=dev)
            iters = torch.empty((n,), dtype=torch.int32, device=dev)
        counters = self._workspace()
        self._before_launch()
        p = self._params()
        prev = torch.cuda.current_device()
        lib, stream = self._launch_ctx()
        tok = self._publish_arm(counters, p) if self.check_nan else None
        ghz = self._general_hz
        if ghz is not None:
            if save:
                raise NotImplementedError("dense saved tensors are only produced by the 5-hazard kernels")
            rc = lib.rcbf_unicycle_safe_action_general(_lib.ptr(st), _lib.ptr(ac), _lib.ptr(mu), _lib.ptr(sg), n, p,
                                                       ghz.ctypes.data, ghz.shape[0], _lib.ptr(out), None,
                                                       _lib.ptr(status), _lib.ptr(counters), stream)
        elif mode == 'Unicycle':
            rc = lib.rcbf_unicycle_safe_action(_lib.ptr(st), _lib.ptr(ac), _lib.ptr(mu), _lib.ptr(sg), n, p,
                                               _lib.ptr(out), _lib.ptr(x), _lib.ptr(lam), _lib.ptr(slack),
                                               _lib.ptr(status), _lib.ptr(iters), _lib.ptr(counters), stream)
        else:
            rc = lib.rcbf_cars_safe_action(_lib.ptr(st), _lib.ptr(ac), _lib.ptr(sg), n, p, _lib.ptr(out),
                                           _lib.ptr(x), _lib.ptr(lam), _lib.ptr(slack), _lib.ptr(status),
                                           _lib.ptr(iters), _lib.ptr(counters), stream)
        if prev != dev.index:
            torch.cuda.set_device(prev)
        self._publish_disarm(p)
        _lib.check(rc, "rcbf_%s_safe_action" % mode)
        self._last_status, self._last_iters = status, iters
        self._after_launch(tok)
        return out, x, lam, slack

    def _forward_meta(self, st, ac, mu, sg):
        """Forward launch of the differentiable path: safe action + one int32 per instance (status << 16 | active set)."""
        dev = self.device
        n = st.shape[0]
        mode = self.env.dynamics_mode
        out = torch.empty((n, 2 if mode == 'Unicycle' else 1), dtype=torch.float32, device=dev)
        meta = torch.empty((n,), dtype=torch.int32, device=dev)
        counters = self._workspace()
        self._before_launch()
        p = self._params()
        prev = torch.cuda.current_device()
        lib, stream = self._launch_ctx()
        tok = self._publish_arm(counters, p) if self.check_nan else None
        ghz = self._general_hz
        if ghz is not None:
            rc = lib.rcbf_unicycle_safe_action_general(_lib.ptr(st), _lib.ptr(ac), _lib.ptr(mu), _lib.ptr(sg), n, p,
                                                       ghz.ctypes.data, ghz.shape[0], _lib.ptr(out), _lib.ptr(meta), None,
                                                       _lib.ptr(counters), stream)
        elif mode == 'Unicycle':
            rc = lib.rcbf_unicycle_safe_action_saved(_lib.ptr(st), _lib.ptr(ac), _lib.ptr(mu), _lib.ptr(sg), n, p,
                                                     _lib.ptr(out), _lib.ptr(meta), _lib.ptr(counters), stream)
        else:
            rc = lib.rcbf_cars_safe_action_saved(_lib.ptr(st), _lib.ptr(ac), _lib.ptr(sg), n, p, _lib.ptr(out),
                                                 _lib.ptr(meta), _lib.ptr(counters), stream)
        if prev != dev.index:
            torch.cuda.set_device(prev)
        self._publish_disarm(p)
        _lib.check(rc, "rcbf_%s_safe_action_saved" % mode)
        self._after_launch(tok)
        return out, meta

    def _backward_meta(self, st, ac, mu, sg, meta, go):
        n = st.shape[0]
        ga = torch.empty_like(ac)
        p = self._params()
        prev = torch.cuda.current_device()
        lib, stream = self._launch_ctx()
        ghz = self._general_hz
        if ghz is not None:
            rc = lib.rcbf_unicycle_safe_action_bwd_general(_lib.ptr(st), _lib.ptr(ac), _lib.ptr(mu), _lib.ptr(sg),
                                                           _lib.ptr(meta), _lib.ptr(go), n, p, ghz.ctypes.data,
                                                           ghz.shape[0], _lib.ptr(ga), stream)
        elif self.env.dynamics_mode == 'Unicycle':
            rc = lib.rcbf_unicycle_safe_action_bwd_meta(_lib.ptr(st), _lib.ptr(ac), _lib.ptr(mu), _lib.ptr(sg),
                                                        _lib.ptr(meta), _lib.ptr(go), n, p, _lib.ptr(ga), stream)
        else:
            rc = lib.rcbf_cars_safe_action_bwd_meta(_lib.ptr(st), _lib.ptr(ac), _lib.ptr(sg), _lib.ptr(meta),
                                                    _lib.ptr(go), n, p, _lib.ptr(ga), stream)
        if prev != self.device.index:
            torch.cuda.set_device(prev)
        _lib.check(rc, "rcbf_safe_action_bwd_meta")
        return ga

    def _backward_raw(self, st, ac, mu, sg, x, lam, slack, go):
        """Legacy backward entry on the dense saved tensors of `_forward_raw(save=True)`."""
        lib = _lib.load()
        dev = self.device
        n = st.shape[0]
        ga = torch.empty_like(ac)
        p = self._params()
        with torch.cuda.device(dev):
            if self.env.dynamics_mode == 'Unicycle':
                rc = lib.rcbf_unicycle_safe_action_bwd(_lib.ptr(st), _lib.ptr(ac), _lib.ptr(mu), _lib.ptr(sg),
                                                       _lib.ptr(x), _lib.ptr(lam), _lib.ptr(slack), _lib.ptr(go), n, p,
                                                       _lib.ptr(ga), _lib.stream_ptr(dev))
            else:
                rc = lib.rcbf_cars_safe_action_bwd(_lib.ptr(st), _lib.ptr(ac), _lib.ptr(sg), _lib.ptr(x),
                                                   _lib.ptr(lam), _lib.ptr(slack), _lib.ptr(go), n, p, _lib.ptr(ga),
                                                   _lib.stream_ptr(dev))
        _lib.check(rc, "rcbf_safe_action_bwd")
        return ga

    # ------------------------------------------------------------------------------------------------- public API
    def get_safe_action(self, state_batch, action_batch, mean_pred_batch, sigma_batch):
        """Safe action = clamp(action + u_cbf) (diff_cbf_qp.py:44-79).  1-D inputs are accepted and return 1-D."""
        expand_dims = len(state_batch.shape) == 1
        if expand_dims:
            action_batch = action_batch.unsqueeze(0)
            state_batch = state_batch.unsqueeze(0)
            mean_pred_batch = mean_pred_batch.unsqueeze(0)
            sigma_batch = sigma_batch.unsqueeze(0)
        assert len(state_batch.shape) == 2 and len(action_batch.shape) == 2 and len(mean_pred_batch.shape) == 2 and \
            len(sigma_batch.shape) == 2, print(state_batch.shape, action_batch.shape, mean_pred_batch.shape,
                                               sigma_batch.shape)
        if torch.is_grad_enabled() and action_batch.requires_grad:
            final_action = _SafeActionFn.apply(self, state_batch, action_batch, mean_pred_batch, sigma_batch)
        else:  # nothing to differentiate (select_action, --no_diff_qp): straight to the kernel, no autograd node
            dev = self.device
            out, _, _, _ = self._forward_raw(_f32c(state_batch, dev), _f32c(action_batch, dev),
                                             _f32c(mean_pred_batch, dev), _f32c(sigma_batch, dev))
            final_action = out if (action_batch.device == dev and action_batch.dtype == torch.float32) else \
                out.to(device=action_batch.device, dtype=action_batch.dtype)
        return final_action if not expand_dims else final_action.squeeze(0)

    def solve_qp(self, Ps, qs, Gs, hs):
        """Row-normalise [G|h] and solve; returns x[:, :-1] (diff_cbf_qp.py:81-109).  Like the reference, Gs is
        normalised IN PLACE."""
        Ghs = torch.cat((Gs, hs.unsqueeze(2)), -1)
        Ghs_norm = torch.max(torch.abs(Ghs), dim=2, keepdim=True)[0]
        Gs /= Ghs_norm
        hs = hs / Ghs_norm.squeeze(-1)
        sol = self.cbf_layer(Ps, qs, Gs, hs,
                             solver_args={"check_Q_spd": False, "maxIter": 100000, "notImprovedLim": 10, "eps": 1e-4})
        safe_action_batch = sol[:, :-1]
        return safe_action_batch

    def cbf_layer(self, Qs, ps, Gs, hs, As=None, bs=None, solver_args=None):
        """Batched QP  min 1/2 x'Qx + p'x  s.t. Gx <= h  -> float32 (B, nz)  (diff_cbf_qp.py:111-144).
        solver_args are accepted for signature compatibility; the kernel converges every QP to its KKT point."""
        if As is not None and bs is not None and (As.numel() > 0 or bs.numel() > 0):
            raise NotImplementedError("equality constraints are never used by the reference (diff_cbf_qp.py:135-137)")
        result = _QPFn.apply(self, Qs, ps, Gs, hs).float()
        if torch.any(torch.isnan(result)):
            print('\033[91m QP Failed to solve - result is nan == True!\033[00m')
            raise Exception('QP Failed to solve')
        return result

    def get_cbf_qp_constraints(self, state_batch, action_batch, mean_pred_batch, sigma_pred_batch):
        """P (B,nz,nz), q (B,nz), G (B,m,nz), h (B,m) as float32 (diff_cbf_qp.py:146-379)."""
        assert len(state_batch.shape) == 2 and len(action_batch.shape) == 2 and len(mean_pred_batch.shape) == 2 and len(
            sigma_pred_batch.shape) == 2, print(state_batch.shape, action_batch.shape, mean_pred_batch.shape,
                                                sigma_pred_batch.shape)
        lib = _lib.load()
        dev = self.device
        out_dev = state_batch.device
        st, ac, sg = _f32c(state_batch, dev), _f32c(action_batch, dev), _f32c(sigma_pred_batch, dev)
        mu = _f32c(mean_pred_batch, dev)
        n = st.shape[0]
        mode = self.env.dynamics_mode
        p = self._params()
        with torch.cuda.device(dev):
            if mode == 'Unicycle' and self._general_hz is not None:
                ghz = self._general_hz
                G = torch.empty((n, ghz.shape[0] + 4, 3), dtype=torch.float32, device=dev)
                h = torch.empty((n, ghz.shape[0] + 4), dtype=torch.float32, device=dev)
                rc = lib.rcbf_unicycle_assemble_general(_lib.ptr(st), _lib.ptr(ac), _lib.ptr(mu), _lib.ptr(sg), n, p,
                                                        ghz.ctypes.data, ghz.shape[0], _lib.ptr(G), _lib.ptr(h),
                                                        _lib.stream_ptr(dev))
                P = torch.diag(torch.tensor([1.e0, 1.e-2, 1e5])).repeat(n, 1, 1).to(out_dev)
            elif mode == 'Unicycle':
                G = torch.empty((n, 9, 3), dtype=torch.float32, device=dev)
                h = torch.empty((n, 9), dtype=torch.float32, device=dev)
                rc = lib.rcbf_unicycle_assemble(_lib.ptr(st), _lib.ptr(ac), _lib.ptr(mu), _lib.ptr(sg), n, p,
                                                _lib.ptr(G), _lib.ptr(h), _lib.stream_ptr(dev))
                if self.num_cbfs < 5:   # drop the inert rows of the padding hazards (see _params._check_hazards)
                    rows = list(range(self.num_cbfs)) + [5, 6, 7, 8]
                    G, h = G[:, rows, :].contiguous(), h[:, rows].contiguous()
                P = torch.diag(torch.tensor([1.e0, 1.e-2, 1e5])).repeat(n, 1, 1).to(out_dev)
            elif mode == 'SimulatedCars':
                G = torch.empty((n, 4, 2), dtype=torch.float32, device=dev)
                h = torch.empty((n, 4), dtype=torch.float32, device=dev)
                rc = lib.rcbf_cars_assemble(_lib.ptr(st), _lib.ptr(ac), _lib.ptr(sg), n, p, _lib.ptr(G), _lib.ptr(h),
                                            _lib.stream_ptr(dev))
                P = torch.diag(torch.tensor([0.1, 1e1])).repeat(n, 1, 1).to(out_dev)
            else:
                raise Exception('Dynamics mode unknown!')
        _lib.check(rc, "rcbf_assemble")
        q = torch.zeros((n, self.action_dim + 1), device=out_dev)
        return P, q, G.to(out_dev), h.to(out_dev)

    def get_control_bounds(self):
        """u_min, u_max tensors on the device (diff_cbf_qp.py:381-395)."""
        u_min = torch.tensor(self.env.safe_action_space.low).to(self.device)
        u_max = torch.tensor(self.env.safe_action_space.high).to(self.device)
        return u_min, u_max

    # ---------------------------------------------------------------------------------------------- diagnostics
    def solver_stats(self):
        """Counters of the last launch: dict(nan, uncertified, f64_passes, trivial, sum_iters, fallback, ...).  After a
        call made with `check_nan = False` (no sync) or a fused env step the numbers are read here: the increments
        since the previous read of that workspace."""
        if getattr(self, "_last_stats", None) is not None:
            c = self._last_stats
        elif self._last_counters is not None:           # workspace of a fused env step (cumulative)
            c = self._last_counters[:8].cpu().tolist()
        elif getattr(self, "_ws", None) is not None:
            c = self._sync_counters()
        else:
            return None
        return dict(nan=c[0], uncertified=c[1], f64_passes=c[2], trivial=c[3], sum_iters=c[4], fallback=c[5],
                    fallback_iters=c[6])


# north_star names the class DiffCBFLayer; the reference only has CBFQPLayer (SURVEY.md "Naming note")
DiffCBFLayer = CBFQPLayer
