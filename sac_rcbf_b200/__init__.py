"""sac_rcbf_b200 -- B200-native (sm_100a) implementation of SAC-RCBF's per-step safety hot path.

Public surface (mirrors the reference's names):
    CBFQPLayer / DiffCBFLayer   rcbf_sac/diff_cbf_qp.py
    CascadeCBFLayer             rcbf_sac/cbf_qp.py
    DynamicsModel               rcbf_sac/dynamics.py (prior paths + device disturbance GPs)
    GPyDisturbanceEstimator, DisturbanceGPBank      rcbf_sac/gp_model.py (exact-GP fit + CUDA posterior kernel)
    UnicycleEnv, SimulatedCarsEnv, build_env        envs/*.py, build_env.py
    generate_model_rollouts, DeviceReplayMemory     rcbf_sac/generate_rollouts.py, replay_memory.py (device-resident)
The compute lives in librcbf_b200.so (hand-written CUDA, C ABI in include/rcbf_b200.h).  No CPU fallback.
"""
from ._lib import RcbfLibraryError, load as load_library  # noqa: F401


def __getattr__(name):  # lazy: importing the package must work on a box without CUDA (build / ABI checks)
    if name in ("CBFQPLayer", "DiffCBFLayer"):
        from . import diff_cbf_qp
        return getattr(diff_cbf_qp, name)
    if name == "CascadeCBFLayer":
        from .cbf_qp import CascadeCBFLayer
        return CascadeCBFLayer
    if name in ("DynamicsModel", "DYNAMICS_MODE", "MAX_STD"):
        from . import dynamics
        return getattr(dynamics, name)
    if name in ("UnicycleEnv", "SimulatedCarsEnv"):
        from . import envs
        return getattr(envs, name)
    if name in ("GPyDisturbanceEstimator", "DisturbanceGPBank"):
        from . import gp_model
        return getattr(gp_model, name)
    if name == "DeviceReplayMemory":
        from .replay_memory import DeviceReplayMemory
        return DeviceReplayMemory
    if name in ("generate_model_rollouts", "rollout_transition"):
        from . import generate_rollouts
        return getattr(generate_rollouts, name)
    if name == "build_env":
        from ._build_env import build_env
        return build_env
    raise AttributeError(name)
