"""Soak the one-launch pending-queue protocol: thousands of fused steps on a stress distribution (about 1 % of the
instances pending per step), fresh random actions every step, auto-reset on; afterwards: no NaN, nothing uncertified,
queue bookkeeping and slots all zero, state finite.  Also alternates batch sizes so grids of different sizes reuse the
same workspace."""
import os, sys, types, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import sac_rcbf_b200 as S

dev = torch.device("cuda")
g = torch.Generator(device=dev); g.manual_seed(7)
hz = torch.tensor([[0., 0.], [-1.5, 1.5], [-1.5, -1.5], [1.5, -1.5], [1.5, 1.5]], device=dev)
total_pending = 0
t0 = time.time()
for n in (1 << 20, 4096 + 17, 1 << 18):
    env = S.UnicycleEnv(num_envs=n, device=dev, auto_reset=True)
    layer = S.CBFQPLayer(env, types.SimpleNamespace(cuda=True), gamma_b=20, k_d=3.0, l_p=0.03)
    env.reset()
    steps = 1500 if n >= (1 << 18) else 3000
    for k in range(steps):
        if k % 50 == 0:   # re-seed the states around the hazards (the layer pushes everybody away otherwise)
            idx = torch.randint(0, 5, (n,), generator=g, device=dev)
            r = 0.3 + 0.9 * torch.rand(n, generator=g, device=dev)
            phi = (2 * torch.rand(n, generator=g, device=dev) - 1) * np.pi
            env.state = torch.stack([hz[idx, 0] + r * torch.cos(phi), hz[idx, 1] + r * torch.sin(phi),
                                     (2 * torch.rand(n, generator=g, device=dev) - 1) * np.pi], 1)
        u = 5.0 * torch.rand((n, 2), generator=g, device=dev) - 2.5
        mu = torch.rand((n, 3), generator=g, device=dev) - 0.5
        sg = 3.0 * torch.rand((n, 3), generator=g, device=dev)
        us, obs, rew, done, info = env.safe_step(layer, u, mu, sg)
    torch.cuda.synchronize()
    c = env._counters[:16].cpu().tolist()
    slots = int(env._counters[16:].abs().sum())
    assert c[0] == 0 and c[1] == 0, c
    assert c[8] == 0 and c[9] == 0 and c[10] == 0 and slots == 0, (c, slots)
    assert torch.isfinite(env.state).all() and torch.isfinite(us).all() and torch.isfinite(obs).all()
    total_pending += c[5]
    print("n=%d steps=%d pending handled=%d trivial=%d" % (n, steps, c[5], c[3]))
print("soak ok: %d pending instances drained in-kernel, %.1f s" % (total_pending, time.time() - t0))
