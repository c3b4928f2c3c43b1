"""Small-batch latency of the reference's real call shapes (main.py:93-95, sac_cbf.py:59-90,131-149): get_safe_action
(no grad), forward+backward, fused safe step, single-env gym step -- B = 1 / 25 / 256 / 512 -- plus fwd/bwd throughput."""
import os, sys, time, types
import numpy as np, torch
sys.path.insert(0, os.getcwd())
import sac_rcbf_b200 as S
from oracle import rcbf_oracle as O

args = types.SimpleNamespace(cuda=True, gp_model_size=2000, l_p=0.03)
dev = torch.device("cuda")


def timeit(fn, iters=300, warm=30):
    for _ in range(warm): fn()
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(iters): fn()
    torch.cuda.synchronize(); return (time.perf_counter() - t0) / iters * 1e6


class _Identity(torch.autograd.Function):
    """autograd floor: a custom Function whose forward / backward launch one elementwise kernel each"""
    @staticmethod
    def forward(ctx, a):
        return a * 1.0

    @staticmethod
    def backward(ctx, g):
        return g * 1.0


_a = torch.zeros(512, 2, device=dev, requires_grad=True)
def _floor():
    _a.grad = None
    _Identity.apply(_a).sum().backward()
print("torch autograd floor (identity Function + .sum().backward(), B = 512): %.1f us" % timeit(_floor))
_e = torch.empty(512, 2, device=dev)
print("torch.empty + one elementwise launch: %.1f us" % timeit(lambda: torch.empty(512, 2, device=dev).copy_(_e)))

for check in (True, False):
    print("== check_nan =", check, "(True: one D2H sync per call, like the reference's NaN test)")
    for B in (1, 25, 256, 512, 4096):
        st, ac, mu, sg = (torch.from_numpy(a).to(dev) for a in O.synth_unicycle(B, seed=3))
        env = S.UnicycleEnv(num_envs=B, precision="f32")
        layer = S.CBFQPLayer(env, args, gamma_b=20, k_d=3.0, l_p=0.03)
        layer.check_nan = check
        t_fwd = timeit(lambda: layer.get_safe_action(st, ac, mu, sg))
        a = ac.clone().requires_grad_(True)
        def fb():
            a.grad = None
            layer.get_safe_action(st, a, mu, sg).sum().backward()
        t_fb = timeit(fb)
        env.state = st
        t_step = timeit(lambda: env.safe_step(layer, ac, mu, sg))
        print("B=%5d  get_safe_action %6.1f us   fwd+bwd %6.1f us   fused safe_step %6.1f us" % (B, t_fwd, t_fb, t_step))
# single-env gym loop (num_envs = 1, float64 layout): step() = launch + ONE D2H
env1 = S.UnicycleEnv()
t1 = timeit(lambda: env1.step(np.array([0.3, 0.1])), iters=200)
print("single-env UnicycleEnv.step (f64, gym contract): %.1f us" % t1)
# throughput of the differentiable path at scale
B = 1 << 22
st, ac, mu, sg = (torch.from_numpy(a).to(dev) for a in O.synth_unicycle(B, seed=12345))
env = S.UnicycleEnv(num_envs=8)
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
t_p = ev(lambda: layer._forward_raw(st, ac, mu, sg))
print("4Mi: forward(plain) %.4f ms  forward(meta) %.4f ms  backward(meta) %.4f ms  -> fwd+bwd %.3e /s" % (t_p, t_f, t_b, B / (t_f + t_b) * 1e3))
o2, x, lam, slack = layer._forward_raw(st, ac, mu, sg, save=True)
t_fo = ev(lambda: layer._forward_raw(st, ac, mu, sg, save=True))
t_bo = ev(lambda: layer._backward_raw(st, ac, mu, sg, x, lam, slack, go))
ga_new = layer._backward_meta(st, ac, mu, sg, meta, go)
ga_old = layer._backward_raw(st, ac, mu, sg, x, lam, slack, go)
d = (ga_new - ga_old).abs()
print("legacy dense path: forward(saved) %.4f ms backward %.4f ms; grad diff new vs legacy: max %.2e, rel-norm %.2e, n>1e-3: %d, out equal %s"
      % (t_fo, t_bo, float(d.max()), float(d.norm() / ga_old.norm()), int((d > 1e-3).sum()), bool(torch.equal(out, o2))))
