"""Differentiable path at scale (BASELINE config 3 semantics at 4 Mi instances): forward(meta) + backward(meta) launches
for ncu (scripts/gpu_prof_bwd.sh) and a CUDA-event timing of both."""
import os, sys, types
import numpy as np, torch
sys.path.insert(0, os.getcwd())
import sac_rcbf_b200 as S
from oracle import rcbf_oracle as O

B = int(sys.argv[1]) if len(sys.argv) > 1 else 1 << 22
env_name = sys.argv[2] if len(sys.argv) > 2 else "Unicycle"
args = types.SimpleNamespace(cuda=True, gp_model_size=2000, l_p=0.03)
dev = torch.device("cuda")
if env_name == "Unicycle":
    st, ac, mu, sg = (torch.from_numpy(a).to(dev) for a in O.synth_unicycle(B, seed=12345))
    env = S.UnicycleEnv(num_envs=8)
else:
    st, ac, mu, sg, _ = (torch.from_numpy(np.asarray(a)).to(dev) for a in O.synth_cars(B, seed=12345))
    env = S.SimulatedCarsEnv(num_envs=8)
layer = S.CBFQPLayer(env, args, gamma_b=20, k_d=3.0, l_p=0.03)
layer.check_nan = False
go = torch.ones_like(ac)


def ev(fn, it=20):
    for _ in range(3): fn()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(it): fn()
    e1.record(); torch.cuda.synchronize(); return e0.elapsed_time(e1) / it


out, meta = layer._forward_meta(st, ac, mu, sg)
t_f = ev(lambda: layer._forward_meta(st, ac, mu, sg))
t_b = ev(lambda: layer._backward_meta(st, ac, mu, sg, meta, go))
heavy = float(((meta >> 16) == 1).float().mean())
print("%s B=%d: forward(meta) %.4f ms  backward(meta) %.4f ms  fwd+bwd %.3e /s  non-trivial %.3f" % (env_name, B, t_f, t_b, B / (t_f + t_b) * 1e3, heavy))
