"""Throughput of the auxiliary kernels (env.step alone, prior predict_next_state, model rollout transition) at 4 Mi
instances: they are plain streaming kernels and should sit near the HBM roofline."""
import os, sys, types
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import sac_rcbf_b200 as S
from sac_rcbf_b200 import workloads

dev = torch.device("cuda")
n = 1 << 22


def timeit(fn, it=10):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(it):
        fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / it


args = types.SimpleNamespace(cuda=True, gp_model_size=3000, l_p=0.03)
for mode in ("Unicycle", "SimulatedCars"):
    if mode == "Unicycle":
        st, ac, mu, sg = (torch.from_numpy(a).to(dev) for a in workloads.synth_unicycle(n, seed=1))
        env = S.UnicycleEnv(num_envs=n, device=dev, auto_reset=True); env.reset(); env.state = st
        t = None
        bytes_step = 16 + 4 + 8 + 16 + 4 + 28 + 4 + 1 + 4 + 1
        n_o = 7
    else:
        st, ac, mu, sg, t = (torch.from_numpy(a).to(dev) for a in workloads.synth_cars(n, seed=1))
        env = S.SimulatedCarsEnv(num_envs=n, device=dev, auto_reset=True); env.reset(); env.state = st; env._t.copy_(t)
        bytes_step = 40 + 4 + 4 + 4 + 40 + 40 + 4 + 4 + 4 + 1 + 4
        n_o = 10
    dm = S.DynamicsModel(env, args)
    ms = timeit(lambda: env.step(ac))
    print("%-14s env.step            %.3f ms  %.3e steps/s  ~%.2f TB/s" % (mode, ms, n / ms * 1e3, n * bytes_step / ms / 1e9))
    ms = timeit(lambda: dm.predict_next_state(st, ac, t, use_gps=True))
    ns = st.shape[1]
    print("%-14s predict_next_state  %.3f ms  %.3e /s  (python wrapper incl. prior-sigma tensors)" % (mode, ms, n / ms * 1e3))
    obs = dm.get_obs(st) if mode == "SimulatedCars" else torch.cat(
        [dm.get_obs(st), torch.zeros(n, 2, device=dev), torch.full((n, 1), 0.1, device=dev)], 1)
    eps = torch.randn(n, ns, device=dev)
    ms = timeit(lambda: S.rollout_transition(env, dm, obs, ac, t, eps))
    print("%-14s rollout_transition  %.3f ms  %.3e /s  (get_state + predict_disturbance + kernel)" % (mode, ms, n / ms * 1e3))
