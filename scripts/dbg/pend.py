import os, sys, types
import numpy as np, torch
sys.path.insert(0, os.getcwd())
import sac_rcbf_b200 as S
from tests.test_gpu_parity import _extreme_unicycle, _cuda
args = types.SimpleNamespace(cuda=True, gp_model_size=2000, l_p=0.03)
env = S.UnicycleEnv(num_envs=8)
layer = S.CBFQPLayer(env, args, gamma_b=20, k_d=3.0, l_p=0.03)
for B in (1 << 14, 1 << 18, 1 << 22):
    st, ac, mu, sg = _extreme_unicycle(B, 6)
    out, _, _, _ = layer._forward_raw(_cuda(st), _cuda(ac), _cuda(mu), _cuda(sg))
    print(B, os.environ.get("RCBF_NO_SAFE2"), layer.solver_stats(), "nan", int(torch.isnan(out).sum()), "sum", float(out.double().sum()))
