"""SimulatedCars get_safe_action (assembly + QP + clamp), 4 Mi instances, a few launches (A/B target)."""
import sys, types, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import sac_rcbf_b200 as S
from sac_rcbf_b200 import workloads as O
dev = torch.device("cuda")
nc = 1 << 22
stc, acc, muc, sgc, tc = (torch.from_numpy(a).to(dev) for a in O.synth_cars(nc, seed=12345))
env = S.SimulatedCarsEnv(num_envs=nc, device=dev)
lay = S.CBFQPLayer(env, types.SimpleNamespace(cuda=True), gamma_b=20, k_d=3.0, l_p=0.03)
lay.check_nan = False
for _ in range(6):
    lay._forward_raw(stc, acc, muc, sgc)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(10):
    lay._forward_raw(stc, acc, muc, sgc)
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 10
print("cars get_safe_action: %.4f ms -> %.3e QP/s" % (ms, nc / ms * 1e3))
