"""TEST INFRASTRUCTURE ONLY -- loader for the *unmodified* reference source.

The reference (yemam3/SAC-RCBF, mounted read-only at /root/reference in the
development container) is pure Python but depends on wheels that are not in
this image (gym, gpytorch, qpth, quadprog).  This module injects minimal stub
modules for those names into ``sys.modules`` and then imports the reference
packages as they are, so that

  * ``envs.unicycle_env.UnicycleEnv`` / ``envs.simulated_cars_env.SimulatedCarsEnv``
  * ``rcbf_sac.dynamics.DynamicsModel`` (prior paths only)
  * ``rcbf_sac.diff_cbf_qp.CBFQPLayer``  (``qpth.qp.QPFunction`` -> oracle/qpth_pdipm.py)
  * ``rcbf_sac.cbf_qp.CascadeCBFLayer``  (``quadprog.solve_qp``  -> oracle/exact_qp.py)

run from the reference's own files.  It is used ONLY by
``oracle/make_golden.py`` (to generate tests/golden/*.npz) and by the
"oracle vs reference" tests that are skipped when /root/reference is absent.
/root/reference does not exist on the GPU box, so nothing on the `-m gpu`
path, in smoke() or in bench.py imports this file.

Nothing under oracle/ is importable from the product package sac_rcbf_b200.
"""
import os
import sys
import types

import numpy as np

REFERENCE_ROOT = os.environ.get("RCBF_REFERENCE_ROOT", "/root/reference")


def reference_available() -> bool:
    return os.path.isfile(os.path.join(REFERENCE_ROOT, "rcbf_sac", "diff_cbf_qp.py"))


class _Box:
    """Minimal gym.spaces.Box (unicycle_env.py:21-23, simulated_cars_env.py:18-20)."""

    def __init__(self, low, high, shape=None, dtype=np.float32):
        self.shape = tuple(shape) if shape is not None else np.shape(low)
        self.dtype = np.dtype(dtype)
        self.low = np.full(self.shape, low, dtype=self.dtype)
        self.high = np.full(self.shape, high, dtype=self.dtype)
        self._rng = np.random.RandomState()

    def seed(self, s=None):
        self._rng = np.random.RandomState(s)
        return [s]

    def sample(self):
        return self._rng.uniform(self.low, self.high).astype(self.dtype)

    def contains(self, x):
        x = np.asarray(x)
        return x.shape == self.shape and bool(np.all(x >= self.low) and np.all(x <= self.high))


class _Env:
    metadata = {}

    @property
    def unwrapped(self):
        return self

    def seed(self, s=None):
        return [s]

    def close(self):
        pass


def _missing(name):
    """True when `name` is neither imported nor installed: only then is a stub put in its place (a real package -- should
    a later image ship qpth / gpytorch / quadprog / gym -- is always preferred)."""
    import importlib.util

    if name in sys.modules:
        return False
    try:
        return importlib.util.find_spec(name) is None
    except (ImportError, ValueError):
        return True


def is_stub(module) -> bool:
    return bool(getattr(module, "__rcbf_stub__", False))


def _install_stubs():
    _install_stub_modules()
    for name in ("gym", "gym.spaces", "gym.error", "gpytorch", "gpytorch.models", "qpth", "qpth.qp", "quadprog"):
        m = sys.modules.get(name)
        if m is not None and getattr(m, "__file__", None) is None and getattr(m, "__path__", None) is None:
            m.__rcbf_stub__ = True          # a types.ModuleType made below, not a package found on disk


def _install_stub_modules():
    if _missing("gym"):
        gym = types.ModuleType("gym")
        spaces = types.ModuleType("gym.spaces")
        error = types.ModuleType("gym.error")
        spaces.Box = _Box
        gym.Env = _Env
        gym.spaces = spaces
        gym.error = error
        sys.modules["gym"] = gym
        sys.modules["gym.spaces"] = spaces
        sys.modules["gym.error"] = error
    if _missing("gpytorch"):
        gpytorch = types.ModuleType("gpytorch")
        models = types.ModuleType("gpytorch.models")

        class ExactGP:  # base class only needed at import time (gp_model.py:12)
            pass

        models.ExactGP = ExactGP
        gpytorch.models = models
        sys.modules["gpytorch"] = gpytorch
        sys.modules["gpytorch.models"] = models
    if _missing("qpth"):
        from oracle import qpth_pdipm

        qpth = types.ModuleType("qpth")
        qp = types.ModuleType("qpth.qp")
        qp.QPFunction = qpth_pdipm.QPFunction
        qpth.qp = qp
        sys.modules["qpth"] = qpth
        sys.modules["qpth.qp"] = qp
    if _missing("quadprog"):
        from oracle import exact_qp

        quadprog = types.ModuleType("quadprog")
        quadprog.solve_qp = exact_qp.quadprog_solve_qp
        sys.modules["quadprog"] = quadprog


_loaded = None


def load_reference():
    """Import the reference packages; returns a namespace of the classes used."""
    global _loaded
    if _loaded is not None:
        return _loaded
    if not reference_available():
        raise RuntimeError("reference source not found at %s" % REFERENCE_ROOT)
    _install_stubs()
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)
    from envs.unicycle_env import UnicycleEnv
    from envs.simulated_cars_env import SimulatedCarsEnv
    from rcbf_sac.dynamics import DynamicsModel, DYNAMICS_MODE, MAX_STD
    from rcbf_sac.diff_cbf_qp import CBFQPLayer
    from rcbf_sac.cbf_qp import CascadeCBFLayer
    from rcbf_sac import generate_rollouts
    from rcbf_sac.replay_memory import ReplayMemory

    ns = types.SimpleNamespace(
        UnicycleEnv=UnicycleEnv,
        SimulatedCarsEnv=SimulatedCarsEnv,
        DynamicsModel=DynamicsModel,
        DYNAMICS_MODE=DYNAMICS_MODE,
        MAX_STD=MAX_STD,
        CBFQPLayer=CBFQPLayer,
        CascadeCBFLayer=CascadeCBFLayer,
        generate_model_rollouts=generate_rollouts.generate_model_rollouts,
        ReplayMemory=ReplayMemory,
    )
    _loaded = ns
    return ns


def make_args(cuda=False, gp_model_size=2000, l_p=0.03):
    """argparse-like namespace with the three fields the hot path reads
    (diff_cbf_qp.py:25, dynamics.py:48,55-58)."""
    return types.SimpleNamespace(cuda=cuda, gp_model_size=gp_model_size, l_p=l_p)
