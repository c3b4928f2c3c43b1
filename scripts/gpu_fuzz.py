"""Fuzz the Unicycle / Cars layers on extreme inputs against the exact oracle (dev tool; summarised in DESIGN.md)."""
import sys, types
sys.path.insert(0, '.')
import numpy as np, torch
import sac_rcbf_b200 as S
from oracle import rcbf_oracle as O
tt = torch.from_numpy
args = types.SimpleNamespace(cuda=True, gp_model_size=2000, l_p=0.03)
rng = np.random.default_rng(2025)
B = 60000
hz = O.UNICYCLE["hazards_locations"]
# strata: at hazard centres, on the collision circle, far outside the arena, huge headings, huge sigma / actions
st = np.zeros((B, 3)); k = B // 6
c = hz[rng.integers(0, 5, B)]
r = np.concatenate([rng.uniform(0, 1e-3, k), 0.72 + rng.normal(0, 1e-4, k), rng.uniform(0, 0.72, k), rng.uniform(5, 50, k),
                    rng.uniform(0.3, 1.2, k), rng.uniform(0.3, 1.2, B - 5 * k)])
phi = rng.uniform(-np.pi, np.pi, B)
st[:, 0] = c[:, 0] + r * np.cos(phi); st[:, 1] = c[:, 1] + r * np.sin(phi)
st[:, 2] = rng.uniform(-np.pi, np.pi, B); st[4 * k:5 * k, 2] = rng.uniform(-2000, 2000, k)
ac = rng.uniform(-1, 1, (B, 2)); ac[5 * k:] = rng.uniform(-2.5, 2.5, (B - 5 * k, 2))
mu = rng.uniform(-0.5, 0.5, (B, 3)); sg = rng.uniform(0, 0.2, (B, 3)); sg[5 * k:] = rng.uniform(0, 3.0, (B - 5 * k, 3))
st, ac, mu, sg = (a.astype(np.float32) for a in (st, ac, mu, sg))
env = S.UnicycleEnv(); layer = S.CBFQPLayer(env, args, gamma_b=20, k_d=3.0, l_p=0.03)
for solver in ("presolve", "pdipm"):
    layer.solver = solver
    d = [tt(a).cuda() for a in (st, ac, mu, sg)]
    out, x, lam, slack = layer._forward_raw(*d, save=True, want_status=True)
    stats = layer.solver_stats(); status = layer._last_status.cpu().numpy()
    fe = O.safe_action("Unicycle", tt(st), tt(ac), tt(mu), tt(sg), solver="exact", gamma_b=20.0).numpy()
    f64 = O.safe_action("Unicycle", tt(st), tt(ac), tt(mu), tt(sg), solver="exact", assembly_dtype=torch.float64, gamma_b=20.0).numpy()
    ill = np.abs(fe - f64).max(1) > 2e-5
    err = np.abs(out.cpu().numpy() - fe).max(1)
    print(solver, stats, 'status hist', np.bincount(status, minlength=6), 'ill-cond frac %.4f' % ill.mean(),
          'max err well-cond %.2e' % err[~ill].max(), 'max err all %.2e' % err.max(), 'min slack %.2e' % float(slack.min()),
          'per-stratum max err', [float('%.1e' % err[i * k:(i + 1) * k][~ill[i * k:(i + 1) * k]].max()) for i in range(6)])
# Cars extremes
stc, acc, muc, sgc, t = O.synth_cars(B, seed=7)
stc[:k, 6] = stc[:k, 4] - rng.uniform(0, 4, k).astype(np.float32)      # car 4 within / inside the 3.5 m radius of car 3
stc[k:2 * k, 8] = stc[k:2 * k, 6] - rng.uniform(0, 4, k).astype(np.float32)
acc[2 * k:3 * k] = rng.uniform(-10, 10, (k, 1)).astype(np.float32)
sgc[3 * k:4 * k, 1::2] = rng.uniform(0, 5, (k, 5)).astype(np.float32)
envc = S.SimulatedCarsEnv(); layc = S.CBFQPLayer(envc, args, gamma_b=20, k_d=3.0, l_p=0.03)
for solver in ("presolve", "pdipm"):
    layc.solver = solver
    out, x, lam, slack = layc._forward_raw(*[tt(a).cuda() for a in (stc, acc, muc, sgc)], save=True, want_status=True)
    fe = O.safe_action("SimulatedCars", tt(stc), tt(acc), tt(muc), tt(sgc), solver="exact", gamma_b=20.0).numpy()
    keep = O.cars_threshold_margin(stc) > 1e-4
    err = np.abs(out.cpu().numpy() - fe).max(1)
    print('cars', solver, layc.solver_stats(), 'max err %.2e' % err[keep].max(), 'min slack %.2e' % float(slack.min()))
