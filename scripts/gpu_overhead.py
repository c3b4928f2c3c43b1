"""Where the ~11 us of a small-batch get_safe_action go (wall clock per call, tight loops)."""
import os, sys, time, types
import numpy as np, torch
sys.path.insert(0, os.getcwd())
import sac_rcbf_b200 as S
from sac_rcbf_b200 import _lib
from sac_rcbf_b200.diff_cbf_qp import _f32c
from oracle import rcbf_oracle as O

args = types.SimpleNamespace(cuda=True, gp_model_size=2000, l_p=0.03)
dev = torch.device("cuda", 0)
B = 512
st, ac, mu, sg = (torch.from_numpy(a).to(dev) for a in O.synth_unicycle(B, seed=3))
env = S.UnicycleEnv(num_envs=B, precision="f32")
layer = S.CBFQPLayer(env, args, gamma_b=20, k_d=3.0, l_p=0.03)
layer.check_nan = False


def t(fn, iters=20000, warm=500):
    for _ in range(warm): fn()
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(iters): fn()
    torch.cuda.synchronize(); return (time.perf_counter() - t0) / iters * 1e6


lib = _lib.load()
out = torch.empty((B, 2), device=dev)
ws = layer._workspace()
p = layer._params()
a_ = (st.data_ptr(), ac.data_ptr(), mu.data_ptr(), sg.data_ptr(), B, p, out.data_ptr(), None, None, None, None, None,
      ws.data_ptr(), None)
print("full get_safe_action            %.2f us" % t(lambda: layer.get_safe_action(st, ac, mu, sg)))
print("  _forward_raw                  %.2f us" % t(lambda: layer._forward_raw(st, ac, mu, sg)))
print("  raw ctypes launch (prebuilt)  %.2f us" % t(lambda: lib.rcbf_unicycle_safe_action(*a_)))
print("  torch.empty((B,2))            %.2f us" % t(lambda: torch.empty((B, 2), dtype=torch.float32, device=dev)))
print("  layer._params()               %.2f us" % t(lambda: layer._params()))
print("  4 x _f32c                     %.2f us" % t(lambda: (_f32c(st, dev), _f32c(ac, dev), _f32c(mu, dev), _f32c(sg, dev))))
print("  _launch_ctx                   %.2f us" % t(lambda: layer._launch_ctx()))
print("  5 x data_ptr                  %.2f us" % t(lambda: (st.data_ptr(), ac.data_ptr(), mu.data_ptr(), sg.data_ptr(), out.data_ptr())))
print("  current_device                %.2f us" % t(lambda: torch.cuda.current_device()))
print("  is_grad_enabled+requires_grad %.2f us" % t(lambda: torch.is_grad_enabled() and ac.requires_grad))
