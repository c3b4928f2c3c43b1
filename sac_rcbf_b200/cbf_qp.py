"""CascadeCBFLayer -- API-compatible shim for the single-instance numpy layer rcbf_sac/cbf_qp.py:5-358.

Same kernels as CBFQPLayer with that file's constants: k_d applied to the sigma term (cbf_qp.py:141), signed sigma map
(:119), P = diag(10, 1e-4, 1e7) (:146), SimulatedCars ignores sigma (:210-211).  `get_u_safe` returns the unclamped
correction only, like the reference (:46-53, :277).  The reference works in numpy float64, so this layer does too:
float64 assembly + row normalisation kernel (rcbf_*_assemble_f64), then the generic float64 QP kernel (rcbf_qp_solve);
the reference's per-solve print (:278) is dropped.
"""
import numpy as np
import torch

from . import _lib, _params
from .diff_cbf_qp import DYNAMICS_MODE


class CascadeCBFLayer:

    def __init__(self, env, gamma_b=100, k_d=1.5, l_p=0.03):
        _lib.require_cuda()
        self._lib = _lib.load()
        self.device = torch.device("cuda", torch.cuda.current_device())
        self.env = env
        self.u_min, self.u_max = self.get_control_bounds()
        self.gamma_b = gamma_b
        self.k_d = k_d
        self.l_p = l_p
        if self.env.dynamics_mode not in DYNAMICS_MODE:
            raise Exception('Dynamics mode not supported.')

    def _params(self):
        if self.env.dynamics_mode == 'Unicycle':
            return _params.unicycle_params(self.env.hazards_locations, self.env.hazards_radius, float(self.gamma_b),
                                           float(self.l_p), self.u_min, self.u_max, p_diag=(1.e1, 1.e-4, 1e7),
                                           sigma_scale=float(self.k_d), abs_sigma_map=False)
        return _params.cars_params(float(self.gamma_b), float(self.env.kp), float(self.env.k_brake),
                                   float(self.u_min[0]), float(self.u_max[0]), sigma_scale=0.0)

    def _dev(self, a, width):
        return torch.as_tensor(np.asarray(a, np.float64).reshape(1, width)).to(self.device)

    def _assemble(self, u_nom, state, mean_pred, sigma_pred, normalise):
        """float64 G (1,m,nz), h (1,m) on the device, optionally row-normalised like cbf_qp.py:262-265."""
        mode = self.env.dynamics_mode
        dev = self.device
        p = self._params()
        with torch.cuda.device(dev):
            if mode == 'Unicycle':
                G = torch.empty((1, 9, 3), dtype=torch.float64, device=dev)
                h = torch.empty((1, 9), dtype=torch.float64, device=dev)
                ts = (self._dev(state, 3), self._dev(u_nom, 2), self._dev(mean_pred, 3), self._dev(sigma_pred, 3))
                rc = self._lib.rcbf_unicycle_assemble_f64(_lib.ptr(ts[0]), _lib.ptr(ts[1]), _lib.ptr(ts[2]),
                                                          _lib.ptr(ts[3]), 1, p, int(normalise), _lib.ptr(G), _lib.ptr(h),
                                                          _lib.stream_ptr(dev))
                P = np.diag([1.e1, 1.e-4, 1e7])                                 # cbf_qp.py:146
            else:
                G = torch.empty((1, 4, 2), dtype=torch.float64, device=dev)
                h = torch.empty((1, 4), dtype=torch.float64, device=dev)
                ts = (self._dev(state, 10), self._dev(u_nom, 1), self._dev(sigma_pred, 10))
                rc = self._lib.rcbf_cars_assemble_f64(_lib.ptr(ts[0]), _lib.ptr(ts[1]), _lib.ptr(ts[2]), 1, p,
                                                      int(normalise), _lib.ptr(G), _lib.ptr(h), _lib.stream_ptr(dev))
                P = np.diag([0.1, 1e1])                                         # :218
        _lib.check(rc, "rcbf_assemble_f64")
        return P, G, h

    def get_u_safe(self, u_nom, s, mean_pred, sigma):
        """Correction to add to u_nom (cbf_qp.py:29-53).  NOTE the argument order (u_nom first).  float64 throughout,
        like the reference: float64 assembly + normalisation kernel, then the generic float64 QP kernel."""
        dev = self.device
        P, G, h = self._assemble(u_nom, s, mean_pred, sigma, normalise=True)
        nz, m = G.shape[2], G.shape[1]
        Q = torch.as_tensor(P, dtype=torch.float64, device=dev).reshape(1, nz, nz).contiguous()
        q = torch.zeros((1, nz), dtype=torch.float64, device=dev)
        x = torch.empty((1, nz), dtype=torch.float64, device=dev)
        lam = torch.empty((1, m), dtype=torch.float64, device=dev)
        slack = torch.empty((1, m), dtype=torch.float64, device=dev)
        with torch.cuda.device(dev):
            rc = self._lib.rcbf_qp_solve(_lib.ptr(Q), _lib.ptr(q), _lib.ptr(G), _lib.ptr(h), 1, nz, m, _lib.ptr(x),
                                         _lib.ptr(lam), _lib.ptr(slack), None, None, None, _lib.stream_ptr(dev))
        _lib.check(rc, "rcbf_qp_solve")
        xs = x[0].cpu().numpy()
        if np.any(np.isnan(xs)):
            raise ValueError("constraints are inconsistent, no solution")   # what quadprog raises (cbf_qp.py:279-281)
        if np.abs(xs[-1]) > 1e-1:
            print('CBF indicates constraint violation might occur. epsilon = {}'.format(xs[-1]))   # :283-284
        return xs[:-1]

    def get_cbf_qp_constraints(self, u_nom, state, mean_pred, sigma_pred):
        """P, q, G, h as float64 ndarrays (cbf_qp.py:55-240)."""
        P, G, h = self._assemble(u_nom, state, mean_pred, sigma_pred, normalise=False)
        return P, np.zeros(P.shape[0]), G[0].cpu().numpy(), h[0].cpu().numpy()

    def get_cbfs(self, hazards_locations, hazards_radius):
        """h(x) and dh/dx closures with the +0.07 buffer of cbf_qp.py:288-325 (host-side helper, not on the QP path)."""
        hazards_locations = np.array(hazards_locations)
        collision_radius = hazards_radius + 0.07

        def lookahead(state):
            if self.env.dynamics_mode in ('SafetyGym_point', 'Unicycle'):
                return np.array([state[0] + self.l_p * np.cos(state[2]), state[1] + self.l_p * np.sin(state[2])])
            return state

        def get_h(state):
            return 0.5 * (np.sum((lookahead(state) - hazards_locations) ** 2, axis=1) - collision_radius ** 2)

        def get_dhdx(state):
            return lookahead(state) - hazards_locations

        return get_h, get_dhdx

    def get_control_bounds(self):
        return self.env.safe_action_space.low, self.env.safe_action_space.high

    def get_min_h_val(self, state):
        get_h, _ = self.get_cbfs(self.env.hazards_locations, self.env.hazards_radius)
        return np.min(get_h(state))
