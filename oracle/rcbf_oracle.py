"""TEST INFRASTRUCTURE ONLY -- CPU restatement of the reference's per-step safety path.

Every function cites the reference file:line it restates (paths relative to
/root/reference).  It is validated against the *unmodified* reference source
(imported through oracle/ref_loader.py in the development container) by
oracle/make_golden.py and tests/test_oracle.py, and against the committed
fixtures tests/golden/*.npz everywhere else.  The QP solve itself goes through
oracle/qpth_pdipm.py (parity for that step is unpinned by the reference, see its
header) and oracle/exact_qp.py.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl
reference legs may import this module; the product package never does.

Arithmetic: torch on CPU, float64 by default (the reference mixes numpy f64 for
envs/dynamics, torch f32 for assembly and f64 for the qpth solve; pass
dtype=torch.float32 to mimic the f32 assembly exactly).
"""
import math

import numpy as np
import torch

from oracle import exact_qp, qpth_pdipm

# The reference's own QP solver (qpth, unvendored / unpinned dependency: rcbf_sac/diff_cbf_qp.py:7,139) is preferred
# whenever the image ships it; otherwise its restatement oracle/qpth_pdipm.py stands in ("parity unpinned").
try:  # pragma: no cover -- absent from this image
    import qpth as _real_qpth
    if getattr(_real_qpth, "__rcbf_stub__", False):      # oracle/ref_loader.py's stand-in, not the package
        raise ImportError("stub")
    from qpth.qp import QPFunction as _RealQPFunction
    QP_BACKEND = "qpth %s (the real package)" % getattr(_real_qpth, "__version__", "?")
except Exception:  # noqa: BLE001
    _RealQPFunction = None
    QP_BACKEND = "oracle/qpth_pdipm.py (restated qpth; the package is absent from the image)"

# envs/unicycle_env.py:24-32
UNICYCLE = dict(
    hazards_locations=np.array([[0.0, 0.0], [-1.0, 1.0], [-1.0, -1.0], [1.0, -1.0], [1.0, 1.0]]) * 1.5,
    hazards_radius=0.6,
    dt=0.02,
    max_episode_steps=1000,
    reward_goal=1.0,
    goal_size=0.3,
    goal_pos=np.array([2.5, 2.5]),
    u_min=np.array([-2.5, -2.5]),
    u_max=np.array([2.5, 2.5]),
    init_state=np.array([-2.5, -2.5, 0.0]),
)
# envs/simulated_cars_env.py:18-26,117-123
CARS = dict(
    dt=0.02, max_episode_steps=300, kp=4.0, k_brake=20.0, u_min=np.array([-10.0]), u_max=np.array([10.0]),
    init_pos=np.array([34.0, 28.0, 22.0, 16.0, 10.0]),
)
# rcbf_sac/dynamics.py:24
MAX_STD = {"Unicycle": [2e-1, 2e-1, 2e-1], "SimulatedCars": [0, 0.2, 0, 0.2, 0, 0.2, 0, 0.2, 0, 0.2]}


# --------------------------------------------------------------------------------------
# Q2-U  constraint assembly, Unicycle (torch layer)            rcbf_sac/diff_cbf_qp.py:202-266
# --------------------------------------------------------------------------------------
def assemble_unicycle(state, action, mean, sigma, gamma_b=20.0, l_p=0.03, hazards=None, hazards_radius=0.6,
                      u_min=None, u_max=None):
    """state (B,3) [x,y,theta], action (B,2), mean (B,3), sigma (B,3) -> P (B,3,3), q (B,3), G (B,9,3), h (B,9)."""
    dt_ = state.dtype
    B = state.shape[0]
    hz = torch.as_tensor(UNICYCLE["hazards_locations"] if hazards is None else hazards).to(dt_)
    # diff_cbf_qp.py:206 casts the hazard table through float32 first
    hz = hz.float().to(dt_)
    u_lo = torch.as_tensor(UNICYCLE["u_min"] if u_min is None else u_min).to(dt_)
    u_hi = torch.as_tensor(UNICYCLE["u_max"] if u_max is None else u_max).to(dt_)
    r_c = 1.2 * hazards_radius                                              # :207
    th = state[:, 2]
    c, s = torch.cos(th), torch.sin(th)                                     # :211-212
    px = state[:, 0] + l_p * c                                              # :216
    py = state[:, 1] + l_p * s                                              # :217
    # g_p = R diag(1, l_p) = [[c, -l_p s],[s, l_p c]]                         :225-233
    g00, g01, g10, g11 = c, -l_p * s, s, l_p * c
    # mu_p = g_p [0, mu_th]' + mu_xy                                          :236-238
    mpx = g01 * mean[:, 2] + mean[:, 0]
    mpy = g11 * mean[:, 2] + mean[:, 1]
    # sigma_p = |g_p| [0, sg_th]' + sg_xy                                     :239-241
    spx = g01.abs() * sigma[:, 2] + sigma[:, 0]
    spy = g11.abs() * sigma[:, 2] + sigma[:, 1]
    dx = px[:, None] - hz[None, :, 0]                                       # :248  (B,K)
    dy = py[:, None] - hz[None, :, 1]
    hcbf = 0.5 * (dx * dx + dy * dy - r_c ** 2)                             # :246
    K = hz.shape[0]
    m = K + 4
    # Lg = dhdp' g_p                                                         :259
    Lg0 = dx * g00[:, None] + dy * g10[:, None]
    Lg1 = dx * g01[:, None] + dy * g11[:, None]
    G = torch.zeros(B, m, 3, dtype=dt_)
    h = torch.zeros(B, m, dtype=dt_)
    G[:, :K, 0] = -Lg0
    G[:, :K, 1] = -Lg1
    G[:, :K, 2] = -1.0                                                      # :260
    # association as in :261 :  gamma*h^3 + ((dhdp.mu_p - |dhdp|.sigma_p) + Lg.u)      (k_d NOT applied)
    h[:, :K] = gamma_b * (hcbf * hcbf * hcbf) + (((dx * mpx[:, None] + dy * mpy[:, None])
                                                  - (dx.abs() * spx[:, None] + dy.abs() * spy[:, None]))
                                                 + (Lg0 * action[:, 0:1] + Lg1 * action[:, 1:2]))
    r = K
    for cidx in range(2):                                                   # :365-377
        G[:, r, cidx] = 1.0
        h[:, r] = u_hi[cidx] - action[:, cidx]
        r += 1
        G[:, r, cidx] = -1.0
        h[:, r] = -u_lo[cidx] + action[:, cidx]
        r += 1
    # :265 builds the diagonal in float32 before moving it to the working dtype
    P = torch.diag(torch.tensor([1.0, 1e-2, 1e5])).to(dt_).repeat(B, 1, 1)
    q = torch.zeros(B, 3, dtype=dt_)
    return P, q, G, h


# --------------------------------------------------------------------------------------
# Q2-C  constraint assembly, SimulatedCars (torch layer)       rcbf_sac/diff_cbf_qp.py:268-357
# --------------------------------------------------------------------------------------
def cars_prior_accels(pos, vel, kp, k_brake, v_des0=None):
    """Shared by Q2-C (:284-290), D3 (dynamics.py:171-177) and D2 (simulated_cars_env.py:58-64).
    pos, vel (B,5).  v_des0 (B,) optional lead-car desired velocity (30 - 10 sin(0.2 t))."""
    v_des = torch.full_like(vel, 30.0)
    if v_des0 is not None:
        v_des[:, 0] = v_des0
    acc = kp * (v_des - vel)
    d01 = pos[:, 0] - pos[:, 1]
    d12 = pos[:, 1] - pos[:, 2]
    d24 = pos[:, 2] - pos[:, 4]
    acc[:, 1] = acc[:, 1] - k_brake * d01 * (d01 < 6.0)
    acc[:, 2] = acc[:, 2] - k_brake * d12 * (d12 < 6.0)
    acc[:, 4] = acc[:, 4] - k_brake * d24 * (d24 < 13.0)
    return acc


def assemble_cars(state, action, mean, sigma, gamma_b=20.0, kp=4.0, k_brake=20.0, u_min=None, u_max=None):
    """state (B,10) [p0,v0,...,p4,v4], action (B,1), mean (B,10) (ignored, as in the reference), sigma (B,10)
    -> P (B,2,2), q (B,2), G (B,4,2), h (B,4)."""
    dt_ = state.dtype
    B = state.shape[0]
    u_lo = torch.as_tensor(CARS["u_min"] if u_min is None else u_min).to(dt_)
    u_hi = torch.as_tensor(CARS["u_max"] if u_max is None else u_max).to(dt_)
    r_c = 3.5                                                               # :272
    pos, vel = state[:, 0::2], state[:, 1::2]                               # :280-281
    acc = cars_prior_accels(pos, vel, kp, k_brake)                          # :284-290 (no lead-car sinusoid, :285)
    acc[:, 3] = 0.0                                                         # :289
    sg = sigma[:, 1::2]                                                     # :299 (velocity entries only)
    h13 = 0.5 * ((pos[:, 2] - pos[:, 3]) ** 2 - r_c ** 2)                   # :306
    h15 = 0.5 * ((pos[:, 4] - pos[:, 3]) ** 2 - r_c ** 2)                   # :307
    h13d = (pos[:, 3] - pos[:, 2]) * (vel[:, 3] - vel[:, 2])                # :310
    h15d = (pos[:, 3] - pos[:, 4]) * (vel[:, 3] - vel[:, 4])                # :311
    # grad(Lf h13) entries at state idx 4..7, grad(Lf h15) at idx 6..9        :314-326
    a4, a5, a6, a7 = vel[:, 2] - vel[:, 3], pos[:, 2] - pos[:, 3], vel[:, 3] - vel[:, 2], pos[:, 3] - pos[:, 2]
    b8, b9, b6, b7 = vel[:, 4] - vel[:, 3], pos[:, 4] - pos[:, 3], vel[:, 3] - vel[:, 4], pos[:, 3] - pos[:, 4]
    # f = (v0,a0,...,v4,a4): idx4=v2, idx5=a2, idx6=v3, idx7=a3, idx8=v4, idx9=a4
    Lff13 = a4 * vel[:, 2] + a5 * acc[:, 2] + a6 * vel[:, 3] + a7 * acc[:, 3]        # :319
    LfD13 = a5.abs() * sg[:, 2] + a7.abs() * sg[:, 3]                                  # :320
    Lff15 = b6 * vel[:, 3] + b7 * acc[:, 3] + b8 * vel[:, 4] + b9 * acc[:, 4]        # :327 (state-index order)
    LfD15 = b7.abs() * sg[:, 3] + b9.abs() * sg[:, 4]                                  # :328
    Lg13 = 50.0 * a7                                                                   # :331 (g = 50 e_7, :303)
    Lg15 = 50.0 * b7                                                                   # :332
    u = action[:, 0]
    G = torch.zeros(B, 4, 2, dtype=dt_)
    h = torch.zeros(B, 4, dtype=dt_)
    h[:, 0] = Lff13 - LfD13 + (gamma_b + gamma_b) * h13d + gamma_b * gamma_b * h13 + Lg13 * u   # :348
    h[:, 1] = Lff15 - LfD15 + (gamma_b + gamma_b) * h15d + gamma_b * gamma_b * h15 + Lg15 * u   # :349
    G[:, 0, 0] = -Lg13                                                      # :350
    G[:, 1, 0] = -Lg15                                                      # :351
    G[:, :2, 1] = -2e2                                                      # :352
    G[:, 2, 0] = 1.0                                                        # :369-370
    h[:, 2] = u_hi[0] - u
    G[:, 3, 0] = -1.0                                                       # :375-376
    h[:, 3] = -u_lo[0] + u
    P = torch.diag(torch.tensor([0.1, 1e1])).to(dt_).repeat(B, 1, 1)        # :356 (f32 literal)
    q = torch.zeros(B, 2, dtype=dt_)
    return P, q, G, h


# --------------------------------------------------------------------------------------
# Q3  row normalisation                                         rcbf_sac/diff_cbf_qp.py:103-106
# --------------------------------------------------------------------------------------
def normalise_rows(G, h):
    n = torch.max(torch.cat((G, h.unsqueeze(2)), -1).abs(), dim=2, keepdim=True)[0]
    return G / n, h / n.squeeze(-1), n.squeeze(-1)


ASSEMBLE = {"Unicycle": assemble_unicycle, "SimulatedCars": assemble_cars}


# --------------------------------------------------------------------------------------
# Q1/Q3/Q4  get_safe_action                                     rcbf_sac/diff_cbf_qp.py:44-144
# --------------------------------------------------------------------------------------
def safe_action(mode, state, action, mean, sigma, solver="qpth", assembly_dtype=torch.float32, eps=1e-4,
                return_aux=False, **kw):
    """Restates CBFQPLayer.get_safe_action.  `action` may require grad (float32 leaf, like the reference).

    solver = "qpth"  : oracle/qpth_pdipm.py, float64, batch-global stop, eps=1e-4, notImprovedLim=10 (:107,:139)
    solver = "exact" : oracle/exact_qp.py (no autograd)
    assembly_dtype   : torch.float32 reproduces the reference (assembly f32, solve f64, result .float());
                       torch.float64 is the "ideal arithmetic" variant.
    """
    force_restatement = kw.pop("force_restatement", False)   # tests: compare the restatement with the real package
    st, ac, mu, sg = (t.to(assembly_dtype) for t in (state, action, mean, sigma))
    P, q, G, h = ASSEMBLE[mode](st, ac, mu, sg, **kw)
    Gn, hn, n = normalise_rows(G, h)
    if solver == "qpth" and _RealQPFunction is not None and not force_restatement:
        e = torch.empty(0, dtype=torch.float64)                             # :107,:139 verbatim
        x = _RealQPFunction(verbose=0, check_Q_spd=False, maxIter=100000, notImprovedLim=10, eps=eps)(
            P.double(), q.double(), Gn.double(), hn.double(), e, e).to(assembly_dtype)
        aux = dict(info=None)
    elif solver == "qpth":
        fn = qpth_pdipm.QPFunction(verbose=0, check_Q_spd=False, maxIter=100000, notImprovedLim=10, eps=eps)
        e = torch.empty(0, dtype=torch.float64)
        x = fn(P.double(), q.double(), Gn.double(), hn.double(), e, e).to(assembly_dtype)
        aux = dict(info=qpth_pdipm._QPFn.last_info)
    elif solver == "exact":
        xe, lam, act, viol = exact_qp.solve_exact(P.double().detach().numpy(), q.double().detach().numpy(),
                                                  Gn.double().detach().numpy(), hn.double().detach().numpy())
        x = torch.from_numpy(xe).to(assembly_dtype)
        aux = dict(lam=lam, active=act, viol=viol)
    else:
        raise ValueError(solver)
    if torch.any(torch.isnan(x)):                                           # :141-143
        raise Exception("QP Failed to solve")
    n_u = ac.shape[1]
    lo = torch.as_tensor(kw.get("u_min", (UNICYCLE if mode == "Unicycle" else CARS)["u_min"])).to(assembly_dtype)
    hi = torch.as_tensor(kw.get("u_max", (UNICYCLE if mode == "Unicycle" else CARS)["u_max"])).to(assembly_dtype)
    final = torch.clamp(ac + x[:, :n_u], lo.repeat(ac.shape[0], 1), hi.repeat(ac.shape[0], 1))   # :77
    if return_aux:
        aux.update(x=x, Gn=Gn, hn=hn, n=n, P=P)
        return final, aux
    return final


# --------------------------------------------------------------------------------------
# N1  CascadeCBFLayer (numpy layer) assembly                    rcbf_sac/cbf_qp.py:55-240
# --------------------------------------------------------------------------------------
def assemble_unicycle_cascade(u_nom, state, mean, sigma, gamma_b=100.0, k_d=1.5, l_p=0.03):
    """Single instance, numpy f64.  Differences from the torch layer: k_d applied (:141), sigma_p without |.|
    (:119), P = diag(10, 1e-4, 1e7) (:146)."""
    hz = UNICYCLE["hazards_locations"]
    r_c = 1.2 * UNICYCLE["hazards_radius"]
    c, s = math.cos(state[2]), math.sin(state[2])
    p = np.array([state[0] + l_p * c, state[1] + l_p * s])                  # :94
    g_p = np.array([[c, -l_p * s], [s, l_p * c]])                           # :100-105
    hs = 0.5 * (np.sum((p - hz) ** 2, axis=1) - r_c ** 2)                   # :108
    dh = p - hz                                                             # :111
    mean_p = mean[:2] + l_p * np.array([-s, c]) * mean[2]                   # :117
    sigma_p = sigma[:2] + l_p * np.array([-s, c]) * sigma[2]                # :119
    K = hz.shape[0]
    G = np.zeros((K + 4, 3))
    h = np.zeros(K + 4)
    for i in range(K):
        G[i, :2] = -dh[i] @ g_p
        G[i, 2] = -1
        h[i] = gamma_b * hs[i] ** 3 + dh[i] @ mean_p + (dh[i] @ g_p) @ u_nom - k_d * np.abs(dh[i]) @ sigma_p  # :138-141
    r = K
    for cidx in range(2):                                                   # :226-238
        G[r, cidx] = 1
        h[r] = UNICYCLE["u_max"][cidx] - u_nom[cidx]
        r += 1
        G[r, cidx] = -1
        h[r] = -UNICYCLE["u_min"][cidx] + u_nom[cidx]
        r += 1
    return np.diag([1.0e1, 1.0e-4, 1e7]), np.zeros(3), G, h


def assemble_cars_cascade(u_nom, state, mean, sigma, gamma_b=100.0, kp=4.0, k_brake=20.0):
    """cbf_qp.py:149-221: as the torch layer but WITHOUT the sigma term (:210-211)."""
    st = torch.as_tensor(np.asarray(state, np.float64))[None]
    ac = torch.as_tensor(np.asarray(u_nom, np.float64))[None]
    z = torch.zeros_like(st)
    P, q, G, h = assemble_cars(st, ac, z, z, gamma_b=gamma_b, kp=kp, k_brake=k_brake)
    return P[0].numpy(), q[0].numpy(), G[0].numpy(), h[0].numpy()


def cascade_u_safe(mode, u_nom, state, mean, sigma, **kw):
    """CascadeCBFLayer.get_u_safe (cbf_qp.py:29-53,242-286): returns the correction only, unclamped."""
    fn = assemble_unicycle_cascade if mode == "Unicycle" else assemble_cars_cascade
    P, q, G, h = fn(np.asarray(u_nom, np.float64), np.asarray(state, np.float64), np.asarray(mean, np.float64),
                    np.asarray(sigma, np.float64), **kw)
    n = np.max(np.abs(np.concatenate((G, h[:, None]), 1)), axis=1)          # :271-274
    x, lam, act, viol = exact_qp.solve_exact(P[None], q[None], (G / n[:, None])[None], (h / n)[None])
    return x[0, :-1], x[0, -1]


# --------------------------------------------------------------------------------------
# D1  UnicycleEnv                                               envs/unicycle_env.py:46-143,215-280
# --------------------------------------------------------------------------------------
def unicycle_obs(state):
    """state (B,3) f64 -> obs (B,7)  (:215-231, compass :260-277)."""
    x, y, th = state[:, 0], state[:, 1], state[:, 2]
    c, s = np.cos(th), np.sin(th)
    vx, vy = UNICYCLE["goal_pos"][0] - x, UNICYCLE["goal_pos"][1] - y
    dist = np.sqrt(vx * vx + vy * vy)
    cx = vx * c + vy * s                     # row-vector times R(theta)      :272-274
    cy = -vx * s + vy * c
    nrm = np.sqrt(cx * cx + cy * cy) + 0.001                                # :276
    return np.stack([x, y, c, s, cx / nrm, cy / nrm, np.exp(-dist)], axis=1)


def unicycle_goal_dist(state):
    return np.sqrt((UNICYCLE["goal_pos"][0] - state[:, 0]) ** 2 + (UNICYCLE["goal_pos"][1] - state[:, 1]) ** 2)


def unicycle_env_step(state, action, episode_step, last_goal_dist):
    """Batched restatement of UnicycleEnv.step (:46-111).  All inputs numpy f64; returns a dict with
    state, obs, reward, done, goal_met, cost (0.1 or 0.0; the reference omits the key when 0), episode_step,
    last_goal_dist."""
    dt = UNICYCLE["dt"]
    a = np.clip(action, -1.0, 1.0)                                          # :62
    st = np.array(state, np.float64, copy=True)
    c, s = np.cos(st[:, 2]), np.sin(st[:, 2])
    st[:, 0] += dt * c * a[:, 0]                                            # :86
    st[:, 1] += dt * s * a[:, 0]
    st[:, 2] += dt * a[:, 1]
    c2, s2 = np.cos(st[:, 2]), np.sin(st[:, 2])                             # :87 uses the UPDATED theta
    st[:, 0] -= dt * 0.1 * c2 * c2
    st[:, 1] -= dt * 0.1 * s2 * c2
    step = episode_step + 1                                                 # :89
    dist = unicycle_goal_dist(st)
    reward = last_goal_dist - dist                                          # :93-95
    goal = dist <= UNICYCLE["goal_size"]                                    # :97,113-123
    reward = reward + goal * UNICYCLE["reward_goal"]
    done = goal | (step >= UNICYCLE["max_episode_steps"])                   # :100-102
    hz = UNICYCLE["hazards_locations"]
    d2 = (st[:, None, 0] - hz[None, :, 0]) ** 2 + (st[:, None, 1] - hz[None, :, 1]) ** 2
    cost = 0.1 * np.any(d2 < UNICYCLE["hazards_radius"] ** 2, axis=1)       # :106-110
    return dict(state=st, obs=unicycle_obs(st), reward=reward, done=done, goal_met=goal, cost=cost,
                episode_step=step, last_goal_dist=dist)


def unicycle_reset(B):
    st = np.tile(UNICYCLE["init_state"], (B, 1))                            # :134-140
    return dict(state=st, obs=unicycle_obs(st), episode_step=np.zeros(B, np.int64),
                last_goal_dist=unicycle_goal_dist(st))


# --------------------------------------------------------------------------------------
# D2  SimulatedCarsEnv                                          envs/simulated_cars_env.py:38-158
# --------------------------------------------------------------------------------------
def cars_obs(state):
    obs = np.array(state, np.float64, copy=True)                            # :155-158
    obs[:, 0::2] /= 100.0
    obs[:, 1::2] /= 30.0
    return obs


def cars_env_step(state, action, t, episode_step):
    """Batched restatement of SimulatedCarsEnv.step (:38-106).  state (B,10), action (B,1), t (B,), step (B,)."""
    dt, kp, kb = CARS["dt"], CARS["kp"], CARS["k_brake"]
    st = np.array(state, np.float64, copy=True)
    pos, vel = torch.from_numpy(st[:, 0::2].copy()), torch.from_numpy(st[:, 1::2].copy())
    v0 = torch.from_numpy(30.0 - 10.0 * np.sin(0.2 * np.asarray(t, np.float64)))    # :59-60
    acc = cars_prior_accels(pos, vel, kp, kb, v_des0=v0).numpy()            # :61-64 (car 4 keeps its P-term)
    acc = acc * 1.1                                                         # :67
    f = np.zeros_like(st)
    f[:, 0::2] = st[:, 1::2]                                                # :73
    f[:, 1::2] = acc                                                        # :74
    f[:, 7] += 50.0 * action[:, 0]                                          # :75-77
    st = st + dt * f
    t_new = t + dt                                                          # :79
    step = episode_step + 1
    done = step >= CARS["max_episode_steps"]                                # :83
    cost = -0.1 * ((st[:, 4] - st[:, 6]) < 2.99) - 0.1 * ((st[:, 6] - st[:, 8]) < 2.99)   # :95-106 (negative)
    reward = -5.0 * np.abs(action[:, 0] ** 2) / CARS["max_episode_steps"]   # :89-93
    return dict(state=st, obs=cars_obs(st), reward=reward, done=done, cost=cost, t=t_new, episode_step=step)


def cars_reset(B, v_noise):
    """:108-125.  v_noise (B,) is the single N(0,0.5) draw shared by all cars of an instance."""
    st = np.zeros((B, 10))
    st[:, 0::2] = CARS["init_pos"]
    st[:, 1::2] = 30.0 + np.asarray(v_noise, np.float64)[:, None]
    st[:, 7] = 35.0
    return dict(state=st, obs=cars_obs(st), t=np.zeros(B), episode_step=np.zeros(B, np.int64))


# --------------------------------------------------------------------------------------
# D3-D5  DynamicsModel prior paths                              rcbf_sac/dynamics.py:60-105,125-261,381-390
# --------------------------------------------------------------------------------------
def predict_next_state(mode, state, action, t=None, mean=None, std=None):
    """Prior step next = s + dt (f + g u) + dt*mean; returns (next, dt*std, t+dt)  (:86-105)."""
    st = np.asarray(state, np.float64)
    u = np.asarray(action, np.float64)
    B = st.shape[0]
    if mode == "Unicycle":
        dt = UNICYCLE["dt"]
        f = np.zeros_like(st)                                               # :141-143
        gu = np.stack([np.cos(st[:, 2]) * u[:, 0], np.sin(st[:, 2]) * u[:, 0], u[:, 1]], axis=1)   # :145-151
    else:
        dt = CARS["dt"]
        pos, vel = torch.from_numpy(st[:, 0::2].copy()), torch.from_numpy(st[:, 1::2].copy())
        v0 = torch.from_numpy(30.0 - 10.0 * np.sin(0.2 * np.asarray(t, np.float64)))            # :172
        acc = cars_prior_accels(pos, vel, CARS["kp"], CARS["k_brake"], v_des0=v0).numpy()
        acc[:, 3] = 0.0                                                     # :176 (and no x1.1)
        f = np.zeros_like(st)
        f[:, 0::2] = st[:, 1::2]
        f[:, 1::2] = acc
        gu = np.zeros_like(st)
        gu[:, 7] = 50.0 * u[:, 0]                                           # :158-162
    nxt = st + dt * (f + gu)
    if mean is None:
        mean, std = prior_disturbance(mode, B)
    nxt = nxt + dt * mean                                                   # :92
    return nxt, dt * std, (None if t is None else t + dt)


def prior_disturbance(mode, B):
    """zero mean, MAX_STD prior (:381-384)."""
    std = np.tile(np.asarray(MAX_STD[mode], np.float64), (B, 1))
    return np.zeros_like(std), std


def get_state(mode, obs):
    obs = np.asarray(obs, np.float64)
    if mode == "Unicycle":                                                  # :216-221
        return np.stack([obs[:, 0], obs[:, 1], np.arctan2(obs[:, 3], obs[:, 2])], axis=1)
    st = obs.copy()                                                         # :222-225
    st[:, 0::2] *= 100.0
    st[:, 1::2] *= 30.0
    return st


def get_obs(mode, state):
    st = np.asarray(state, np.float64)
    if mode == "Unicycle":                                                  # :249-254
        return np.stack([st[:, 0], st[:, 1], np.cos(st[:, 2]), np.sin(st[:, 2])], axis=1)
    return cars_obs(st)                                                     # :255-258


# --------------------------------------------------------------------------------------
# model-rollout transition                                     rcbf_sac/generate_rollouts.py:29-66
# --------------------------------------------------------------------------------------
def rollout_step(mode, obs, action, t, eps, mean=None, std=None):
    """One transition of generate_model_rollouts for a batch (numpy f64).  eps = standard-normal draw such that
    np.random.normal(mu, sd) == mu + sd * eps (:31).  Returns next_obs, reward, done, next_t."""
    obs = np.asarray(obs, np.float64)
    action = np.asarray(action, np.float64)
    B = obs.shape[0]
    state = get_state(mode, obs)                                            # :29
    mu, sd, next_t = predict_next_state(mode, state, action, t, mean=mean, std=std)   # :30
    nxt = mu + sd * np.asarray(eps, np.float64)                             # :31
    nobs = get_obs(mode, nxt)                                               # :32
    if mode == "Unicycle":
        dist_prev = -np.log(obs[:, -1])                                     # :37
        goal_rel = UNICYCLE["goal_pos"][None, :] - nobs[:, :2]              # :38
        dist = np.linalg.norm(goal_rel, axis=1)                             # :39
        c, s = np.cos(nxt[:, 2]), np.sin(nxt[:, 2])
        comp = np.stack([goal_rel[:, 0] * c + goal_rel[:, 1] * s, -goal_rel[:, 0] * s + goal_rel[:, 1] * c], 1)   # :42
        comp = comp / (np.sqrt(np.sum(comp ** 2, axis=1, keepdims=True)) + 0.001)                                 # :43
        nobs = np.hstack((nobs, comp, np.exp(-dist)[:, None]))              # :44
        reached = dist <= 0.3
        reward = (dist_prev - dist) * 1.0 + reached * 1.0                   # :50
        reward = reward + 1.0 * reached                                     # :53 (added twice)
        done = reached                                                      # :54
    else:
        reward = -5.0 * np.abs(action[:, 0] ** 2) / CARS["max_episode_steps"]   # :61
        done = next_t >= CARS["max_episode_steps"] * CARS["dt"]                 # :64
    return nobs, reward, done, next_t


# --------------------------------------------------------------------------------------
# Synthetic inputs
# --------------------------------------------------------------------------------------
# The synthetic workloads are defined once, in the product package (they are workload definitions, not reference
# arithmetic); re-exported here because every parity test draws its inputs through this module.
from sac_rcbf_b200.workloads import synth_cars, synth_unicycle  # noqa: E402,F401


def cars_threshold_margin(state):
    """min distance of an instance to one of the `< 6 / < 13` braking switches (SURVEY 'Discontinuities')."""
    st = np.asarray(state, np.float64)
    p = st[:, 0::2]
    return np.minimum(np.minimum(np.abs(p[:, 0] - p[:, 1] - 6.0), np.abs(p[:, 1] - p[:, 2] - 6.0)),
                      np.abs(p[:, 2] - p[:, 4] - 13.0))
