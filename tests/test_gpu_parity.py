"""GPU parity tests (`-m gpu`): the CUDA path, called through the Python mirror of the reference API (which goes through
the C ABI of include/rcbf_b200.h), against the CPU oracle and the golden fixtures generated from the reference source.

Tolerances are the ones BASELINE.json's north_star states:
    dynamics        1e-6 relative (fp32 vs the float64 oracle)
    safe actions    1e-4 absolute vs the reference QP solve, every normalised row satisfied to -1e-6
    gradients       1e-3 relative (norm-wise per batch)
"""
import types

import numpy as np
import pytest
import torch

from oracle import exact_qp, rcbf_oracle as O

pytestmark = pytest.mark.gpu

ACT_TOL = 1e-4
ROW_TOL = -1e-6
GRAD_TOL = 1e-3
DYN_RTOL = 1e-6

tt = torch.from_numpy


def _cuda(a):
    return torch.from_numpy(np.ascontiguousarray(a)).cuda()


@pytest.fixture(scope="module")
def S():
    import sac_rcbf_b200 as S_
    S_.load_library()
    return S_


def _args():
    return types.SimpleNamespace(cuda=True, gp_model_size=2000, l_p=0.03)


@pytest.fixture(scope="module")
def uni(S):
    env = S.UnicycleEnv()
    return env, S.CBFQPLayer(env, _args(), gamma_b=20, k_d=3.0, l_p=0.03)


@pytest.fixture(scope="module")
def cars(S):
    env = S.SimulatedCarsEnv()
    return env, S.CBFQPLayer(env, _args(), gamma_b=20, k_d=3.0, l_p=0.03)


def _row_slack(Gn, hn, x):
    return (hn.astype(np.float64) - np.einsum("bmj,bj->bm", Gn.astype(np.float64), x.astype(np.float64))).min()


# ----------------------------------------------------------------------------------------------------- assembly
@pytest.mark.parametrize("which,name", [("uni", "unicycle_layer_b256.npz"), ("cars", "cars_layer_b512.npz")])
def test_assembly_vs_golden(request, golden, which, name):
    env, layer = request.getfixturevalue(which)
    g = golden(name)
    P, q, G, h = layer.get_cbf_qp_constraints(*(_cuda(g[k]) for k in ("state", "action", "mean", "sigma")))
    assert P.shape == g["P"].shape and q.shape == g["q"].shape and G.shape == g["G"].shape and h.shape == g["h"].shape
    np.testing.assert_array_equal(P.cpu().numpy(), g["P"])
    np.testing.assert_array_equal(q.cpu().numpy(), g["q"])
    n = np.maximum(np.abs(g["G"]).max(2), np.abs(g["h"]))
    # same op order as the reference: differences are the last ulp of cos/sin only
    assert (np.abs(G.cpu().numpy() - g["G"]) / n[:, :, None]).max() < 5e-7
    assert (np.abs(h.cpu().numpy() - g["h"]) / n).max() < 3e-6
    if which == "cars":     # no transcendental on this path: bit exact
        np.testing.assert_array_equal(G.cpu().numpy(), g["G"])
        np.testing.assert_array_equal(h.cpu().numpy(), g["h"])


# ----------------------------------------------------------------------------------------------------- forward
@pytest.mark.parametrize("which,name", [("uni", "unicycle_layer_b256.npz"), ("cars", "cars_layer_b512.npz")])
def test_safe_action_vs_golden(request, golden, which, name):
    env, layer = request.getfixturevalue(which)
    g = golden(name)
    out = layer.get_safe_action(*(_cuda(g[k]) for k in ("state", "action", "mean", "sigma")))
    assert out.dtype == torch.float32 and out.is_cuda and out.shape == g["safe_action"].shape
    err = np.abs(out.cpu().numpy() - g["safe_action"])
    assert err.max() < ACT_TOL, err.max()
    # vs the exact optimum of the reference-assembled QP
    nu = g["action"].shape[1]
    lo, hi = env.safe_action_space.low, env.safe_action_space.high
    ex = np.clip(g["action"] + g["x_exact"][:, :nu].astype(np.float32), lo, hi)
    assert np.abs(out.cpu().numpy() - ex).max() < ACT_TOL
    st = layer.solver_stats()
    assert st["nan"] == 0 and st["uncertified"] == 0
    # 1-D contract (diff_cbf_qp.py:64-69,79)
    o1 = layer.get_safe_action(*(_cuda(g[k][0]) for k in ("state", "action", "mean", "sigma")))
    assert o1.shape == (nu,) and torch.equal(o1, out[0])
    # the golden 1-D value comes from qpth stopping on ITS residual test at B=1 (eps=1e-4, diff_cbf_qp.py:107), which
    # leaves it up to ~1e-2 from the optimum in the weakly weighted omega; we return the optimum
    assert np.abs(o1.cpu().numpy() - g["safe_action_1d"]).max() < 2e-2


def _forward_with_aux(layer, st, ac, mu, sg):
    out, x, lam, slack = layer._forward_raw(_cuda(st), _cuda(ac), _cuda(mu), _cuda(sg), save=True, want_status=True)
    return (out.cpu().numpy(), x.cpu().numpy(), lam.cpu().numpy(), slack.cpu().numpy(),
            layer._last_status.cpu().numpy(), layer._last_iters.cpu().numpy())


@pytest.mark.parametrize("mode", ["Unicycle", "SimulatedCars"])
def test_safe_action_vs_oracle_synthetic(request, mode):
    env, layer = request.getfixturevalue("uni" if mode == "Unicycle" else "cars")
    B = 20000
    if mode == "Unicycle":
        st, ac, mu, sg = O.synth_unicycle(B, seed=2024)
        keep = np.ones(B, bool)
    else:
        st, ac, mu, sg, _ = O.synth_cars(B, seed=2024)
        keep = O.cars_threshold_margin(st) > 1e-4          # braking switches (SURVEY 'Discontinuities')
    out, x, lam, slack, status, iters = _forward_with_aux(layer, st, ac, mu, sg)
    assert not np.isnan(out).any()
    assert (status <= 2).all(), np.bincount(status)
    # oracle: reference-order f32 assembly + exact solve; and the same with f64 assembly to flag instances whose
    # REFERENCE answer is not determined to 1e-4 by its own float32 data (last-ulp cos/sin sensitivity)
    fe, aux = O.safe_action(mode, tt(st), tt(ac), tt(mu), tt(sg), solver="exact", return_aux=True, gamma_b=20.0)
    f64 = O.safe_action(mode, tt(st), tt(ac), tt(mu), tt(sg), solver="exact", assembly_dtype=torch.float64, gamma_b=20.0)
    illcond = (np.abs(fe.numpy() - f64.numpy()).max(1) > 2e-5)
    ok = keep & ~illcond
    err = np.abs(out - fe.numpy()).max(1)
    assert illcond.mean() < 1e-3 and (~keep).mean() < 1e-3
    assert err[ok].max() < ACT_TOL, (err[ok].max(), int(np.argmax(err * ok)))
    assert err.max() < 5e-3                                  # even the ill-conditioned ones stay close
    # every normalised row satisfied by OUR (u, eps) on OUR normalised data
    Gn, hn = aux["Gn"].numpy(), aux["hn"].numpy()
    assert _row_slack(Gn[ok], hn[ok], x[ok]) > ROW_TOL
    # duals: non-negative, complementary
    assert lam.min() >= 0 and np.abs(lam * slack).max() < 1e-5 * max(1.0, lam.max())


@pytest.mark.parametrize("mode", ["Unicycle", "SimulatedCars"])
def test_pdipm_solver_mode_matches_presolve_mode(request, mode):
    """The interior-point-only mode (north_star's solver) and the default presolve mode end in the same float64 KKT
    certificate, so they must agree to rounding of the certificate inputs (here: bit for bit)."""
    env, layer = request.getfixturevalue("uni" if mode == "Unicycle" else "cars")
    B = 1 << 18
    arrs = O.synth_unicycle(B, seed=77) if mode == "Unicycle" else O.synth_cars(B, seed=77)
    st, ac, mu, sg = arrs[:4]
    ref = _forward_with_aux(layer, st, ac, mu, sg)
    stats_ref = layer.solver_stats()
    layer.solver = "pdipm"
    try:
        alt = _forward_with_aux(layer, st, ac, mu, sg)
        stats = layer.solver_stats()
    finally:
        layer.solver = "presolve"
    assert (alt[4] <= 2).all() and stats["nan"] == 0 and stats["uncertified"] == 0
    same = (alt[4] == 1) & (ref[4] == 1)          # both certified: identical active set -> identical numbers
    assert same.mean() > 0.2
    np.testing.assert_array_equal(alt[0][same | (ref[4] == 0)], ref[0][same | (ref[4] == 0)])
    assert np.abs(alt[0] - ref[0]).max() < 1e-5
    # the interior point really ran: iterations were spent, the presolve spent at most nz rounds
    assert stats["sum_iters"] > stats_ref["sum_iters"]


def _extreme_unicycle(B, seed):
    """Instances around the hazards with actions up to the actuator limits and disturbance std up to 3 (15x MAX_STD):
    about 1 % of them need a constraint dropped, which the greedy presolve does not do."""
    rng = np.random.default_rng(seed)
    hz = O.UNICYCLE["hazards_locations"]
    idx = rng.integers(0, len(hz), B)
    r, phi = rng.uniform(0.3, 1.2, B), rng.uniform(-np.pi, np.pi, B)
    st = np.stack([hz[idx, 0] + r * np.cos(phi), hz[idx, 1] + r * np.sin(phi), rng.uniform(-np.pi, np.pi, B)], 1)
    ac = rng.uniform(-2.5, 2.5, (B, 2))
    mu = rng.uniform(-0.5, 0.5, (B, 3))
    sg = rng.uniform(0, 3.0, (B, 3))
    return tuple(a.astype(np.float32) for a in (st, ac, mu, sg))


def _unicycle_without_workspace(layer, st, ac, mu, sg):
    from sac_rcbf_b200 import _lib
    lib = _lib.load()
    d = [_cuda(a) for a in (st, ac, mu, sg)]
    out = torch.empty((st.shape[0], 2), dtype=torch.float32, device="cuda")
    rc = lib.rcbf_unicycle_safe_action(_lib.ptr(d[0]), _lib.ptr(d[1]), _lib.ptr(d[2]), _lib.ptr(d[3]), st.shape[0],
                                       layer._params(), _lib.ptr(out), None, None, None, None, None, None,
                                       _lib.stream_ptr(layer.device))
    assert rc == 0
    return out.cpu().numpy()


def test_pending_instances_with_and_without_workspace(uni, cars):
    """Pending Unicycle instances are queued and finished by exhaustive enumeration -- by the draining warps of the same
    kernel when a workspace is given, by the pass-2 kernel scanning for the pending sentinel when workspace = NULL.
    Both must give identical results (and the oracle's).  SimulatedCars (10 candidate active sets) enumerates inline
    and must leave nothing pending at all."""
    from sac_rcbf_b200 import _lib
    env, layer = uni
    B = 1 << 18
    st, ac, mu, sg = _extreme_unicycle(B, 5)
    for _ in range(2):      # twice: the queue must come back empty for the next call
        ref = _forward_with_aux(layer, st, ac, mu, sg)
        stats = layer.solver_stats()
        assert stats["fallback"] > 200 and stats["uncertified"] == 0 and stats["nan"] == 0 and (ref[4] <= 2).all()
        assert not np.isnan(ref[0]).any()
        ws = layer._ws
        assert int(ws[8:11].abs().sum()) == 0 and int(ws[16:].abs().sum()) == 0   # bookkeeping reset, slots cleared
    np.testing.assert_array_equal(_unicycle_without_workspace(layer, st, ac, mu, sg), ref[0])
    pend = ref[5] == 4                                    # iters == nz + 1 marks "enumerated"
    assert pend.sum() == stats["fallback"]
    k = np.flatnonzero(pend)[:300]
    fe = O.safe_action("Unicycle", tt(st[k]), tt(ac[k]), tt(mu[k]), tt(sg[k]), solver="exact", gamma_b=20.0).numpy()
    f64 = O.safe_action("Unicycle", tt(st[k]), tt(ac[k]), tt(mu[k]), tt(sg[k]), solver="exact",
                        assembly_dtype=torch.float64, gamma_b=20.0).numpy()
    well = np.abs(fe - f64).max(1) <= 2e-5
    assert well.mean() > 0.5 and np.abs(ref[0][k] - fe)[well].max() < 1e-4
    # SimulatedCars: nothing is left pending
    lib = _lib.load()
    envc, layc = cars
    Bc = 1 << 20
    stc, acc, muc, sgc, _ = O.synth_cars(Bc, seed=5)
    refc = _forward_with_aux(layc, stc, acc, muc, sgc)
    sc = layc.solver_stats()
    assert sc["fallback"] == 0 and sc["uncertified"] == 0 and (refc[4] <= 2).all() and (refc[5] == 3).sum() > 100
    outc = torch.empty((Bc, 1), dtype=torch.float32, device="cuda")
    dc = [_cuda(a) for a in (stc, acc, sgc)]
    rc = lib.rcbf_cars_safe_action(_lib.ptr(dc[0]), _lib.ptr(dc[1]), _lib.ptr(dc[2]), Bc, layc._params(),
                                   _lib.ptr(outc), None, None, None, None, None, None, _lib.stream_ptr(layc.device))
    assert rc == 0
    np.testing.assert_array_equal(outc.cpu().numpy(), refc[0])


def test_pending_queue_overflow_falls_back_to_the_sentinel_scan(uni):
    """More pending instances in one call than the 32 752-slot queue holds: the stored part is drained as usual, the
    rest is found by the last block scanning for the sentinel.  Same results as the workspace-free path, queue clean."""
    env, layer = uni
    B = 1 << 22
    st, ac, mu, sg = _extreme_unicycle(B, 6)
    out, _, _, _ = layer._forward_raw(_cuda(st), _cuda(ac), _cuda(mu), _cuda(sg))
    stats = layer.solver_stats()
    assert stats["fallback"] > 32752 and stats["uncertified"] == 0 and stats["nan"] == 0
    ws = layer._ws
    assert int(ws[8:11].abs().sum()) == 0 and int(ws[16:].abs().sum()) == 0
    got = out.cpu().numpy()
    assert not np.isnan(got).any()
    np.testing.assert_array_equal(_unicycle_without_workspace(layer, st, ac, mu, sg), got)
    out2, _, _, _ = layer._forward_raw(_cuda(st[:4096]), _cuda(ac[:4096]), _cuda(mu[:4096]), _cuda(sg[:4096]))
    np.testing.assert_array_equal(out2.cpu().numpy(), got[:4096])     # the next call starts from a clean queue


@pytest.mark.parametrize("mode", ["Unicycle", "SimulatedCars"])
def test_ragged_empty_and_unaligned_batches(request, mode):
    """Edge cases of the tiled kernel: empty batch, sizes around the 32-instance tile, and row slices whose base
    address is not 16-byte aligned (the TMA bulk-copy staging must fall back to plain loads).  All must equal the
    corresponding rows of one big aligned launch bit for bit."""
    env, layer = request.getfixturevalue("uni" if mode == "Unicycle" else "cars")
    B = 4099
    arrs = O.synth_unicycle(B, seed=3, hazard_frac=0.6) if mode == "Unicycle" else O.synth_cars(B, seed=3)
    d = [_cuda(a) for a in arrs[:4]]
    ref = layer.get_safe_action(*d)
    for n in (0, 1, 31, 32, 33, 63, 64, 65, 1000):
        out = layer.get_safe_action(*(t[:n] for t in d))
        assert out.shape == ref[:n].shape and torch.equal(out, ref[:n])
    for off in (1, 2, 3, 5, 7):                 # 12/8/40-byte rows: most offsets break 16-byte alignment
        sl = [t[off:off + 777] for t in d]
        assert any(t.data_ptr() % 16 for t in sl)
        out = layer.get_safe_action(*sl)
        assert torch.equal(out, ref[off:off + 777])
    # non-contiguous / float64 inputs are accepted like torch ops would
    out = layer.get_safe_action(*(t.double() for t in d))
    assert out.dtype == torch.float64 and torch.equal(out.float(), ref)


def test_one_nan_instance_in_a_large_batch_raises_and_is_flagged(uni):
    env, layer = uni
    B = 100000
    st, ac, mu, sg = O.synth_unicycle(B, seed=8)
    st[54321, 1] = np.nan
    d = [_cuda(a) for a in (st, ac, mu, sg)]
    with pytest.raises(Exception, match="QP Failed to solve"):
        layer.get_safe_action(*d)
    layer.check_nan = False
    try:
        out, x, lam, slack = layer._forward_raw(*d, save=False, want_status=True)
    finally:
        layer.check_nan = True
    status = layer._last_status.cpu().numpy()
    assert status[54321] == 4 and (np.delete(status, 54321) <= 2).all()
    assert torch.isnan(out[54321]).all() and not torch.isnan(out).sum().item() > 2


@pytest.mark.parametrize("mode,gamma_b,l_p", [("Unicycle", 100.0, 0.03), ("Unicycle", 5.0, 0.1), ("SimulatedCars", 100.0, 0.03),
                                              ("SimulatedCars", 3.0, 0.03)])
def test_other_layer_parameters_vs_oracle(S, mode, gamma_b, l_p):
    """The layer's constructor defaults (gamma_b=100, diff_cbf_qp.py:12) and other gamma_b / l_p / hazard layouts /
    actuator bounds, not only the README's gamma_b=20."""
    import types as _t
    if mode == "Unicycle":
        env = S.UnicycleEnv()
        env.hazards_locations = np.array([[0.3, -0.2], [-1.2, 1.4], [-1.7, -0.9], [1.1, -1.6], [2.0, 0.7]])
        env.hazards_radius = 0.45
        env.safe_action_space = _t.SimpleNamespace(low=np.array([-1.5, -3.0], np.float32), high=np.array([2.0, 3.0], np.float32))
        st, ac, mu, sg = O.synth_unicycle(20000, seed=17)
        kw = dict(gamma_b=gamma_b, l_p=l_p, hazards=env.hazards_locations, hazards_radius=0.45,
                  u_min=np.array([-1.5, -3.0]), u_max=np.array([2.0, 3.0]))
    else:
        env = S.SimulatedCarsEnv()
        env.kp, env.k_brake = 3.0, 15.0
        env.safe_action_space = _t.SimpleNamespace(low=np.array([-6.0], np.float32), high=np.array([8.0], np.float32))
        st, ac, mu, sg, _ = O.synth_cars(20000, seed=17)
        kw = dict(gamma_b=gamma_b, kp=3.0, k_brake=15.0, u_min=np.array([-6.0]), u_max=np.array([8.0]))
    layer = S.CBFQPLayer(env, _args(), gamma_b=gamma_b, k_d=1.5, l_p=l_p)
    out = layer.get_safe_action(_cuda(st), _cuda(ac), _cuda(mu), _cuda(sg)).cpu().numpy()
    assert layer.solver_stats()["uncertified"] == 0
    fe = O.safe_action(mode, tt(st), tt(ac), tt(mu), tt(sg), solver="exact", **kw).numpy()
    f64 = O.safe_action(mode, tt(st), tt(ac), tt(mu), tt(sg), solver="exact", assembly_dtype=torch.float64, **kw).numpy()
    ok = np.abs(fe - f64).max(1) <= 2e-5
    if mode == "SimulatedCars":
        ok &= O.cars_threshold_margin(st) > 1e-4
    assert ok.mean() > 0.995
    assert np.abs(out - fe)[ok].max() < ACT_TOL


def test_fewer_than_five_hazards(S):
    """The reference layer sizes itself from len(env.hazards_locations) (diff_cbf_qp.py:35); 1..5 hazards are supported
    by padding with inert far-away hazards."""
    env = S.UnicycleEnv()
    env.hazards_locations = np.array([[0.4, 0.1], [-1.0, 1.2], [1.3, -0.8]])
    layer = S.CBFQPLayer(env, _args(), gamma_b=20, k_d=3.0, l_p=0.03)
    assert layer.num_cbfs == 3 and layer.num_ineq_constraints == 7
    st, ac, mu, sg = O.synth_unicycle(20000, seed=23)
    P, q, G, h = layer.get_cbf_qp_constraints(_cuda(st), _cuda(ac), _cuda(mu), _cuda(sg))
    kw = dict(gamma_b=20.0, hazards=env.hazards_locations)
    P2, q2, G2, h2 = O.assemble_unicycle(tt(st), tt(ac), tt(mu), tt(sg), **kw)
    assert G.shape == (20000, 7, 3) and h.shape == (20000, 7)
    n = np.maximum(np.abs(G2.numpy()).max(2), np.abs(h2.numpy()))
    assert (np.abs(G.cpu().numpy() - G2.numpy()) / n[:, :, None]).max() < 5e-7
    assert (np.abs(h.cpu().numpy() - h2.numpy()) / n).max() < 3e-6
    out = layer.get_safe_action(_cuda(st), _cuda(ac), _cuda(mu), _cuda(sg)).cpu().numpy()
    fe = O.safe_action("Unicycle", tt(st), tt(ac), tt(mu), tt(sg), solver="exact", **kw).numpy()
    f64 = O.safe_action("Unicycle", tt(st), tt(ac), tt(mu), tt(sg), solver="exact", assembly_dtype=torch.float64, **kw).numpy()
    ok = np.abs(fe - f64).max(1) <= 2e-5
    assert ok.mean() > 0.995 and np.abs(out - fe)[ok].max() < ACT_TOL
    # the 7-row system of get_cbf_qp_constraints goes through solve_qp too (padded to 9 rows inside)
    xq = layer.solve_qp(P[:512], q[:512], G[:512].clone(), h[:512]).cpu().numpy()
    assert np.abs(np.clip(ac[:512] + xq, -2.5, 2.5) - fe[:512])[ok[:512]].max() < ACT_TOL


def test_trivial_instances_pass_through(uni):
    env, layer = uni
    B = 4096
    st = np.tile(np.array([[2.9, -0.1, 0.3]], np.float32), (B, 1))     # far from every hazard
    rng = np.random.default_rng(0)
    ac = rng.uniform(-1, 1, (B, 2)).astype(np.float32)
    z = np.zeros((B, 3), np.float32)
    out, x, lam, slack, status, iters = _forward_with_aux(layer, st, ac, z, z)
    assert (status == 0).all() and (iters == 0).all()
    np.testing.assert_array_equal(out, ac)


def test_nan_raises_like_reference(uni):
    env, layer = uni
    st = torch.tensor([[float("nan"), 0.0, 0.0]], device="cuda")
    with pytest.raises(Exception, match="QP Failed to solve"):
        layer.get_safe_action(st, torch.zeros(1, 2, device="cuda"), torch.zeros(1, 3, device="cuda"),
                              torch.zeros(1, 3, device="cuda"))
    with pytest.raises(AssertionError):
        layer.get_cbf_qp_constraints(torch.zeros(3, device="cuda"), torch.zeros(2, device="cuda"),
                                     torch.zeros(3, device="cuda"), torch.zeros(3, device="cuda"))


@pytest.mark.parametrize("B", [4, 5000])
def test_nan_raises_on_the_fused_step_too(S, uni, cars, B):
    """The reference raises on ANY NaN safe action (diff_cbf_qp.py:141-143); so does the fused env step while
    `check_nan` is on -- once per offending step, not for ever after -- and with `check_nan = False` the step goes
    through and the counter is there to be read."""
    _, layer = uni
    st, ac, mu, sg = O.synth_unicycle(B, seed=12)
    env = S.UnicycleEnv(num_envs=B)
    env.state = _cuda(st)
    bad = _cuda(mu).clone()
    bad[B // 2, 1] = float("nan")
    env.safe_step(layer, _cuda(ac), _cuda(mu), _cuda(sg))                       # clean step: no exception
    with pytest.raises(Exception, match="QP Failed to solve"):
        env.safe_step(layer, _cuda(ac), bad, _cuda(sg))
    env.state = _cuda(st)
    env.safe_step(layer, _cuda(ac), _cuda(mu), _cuda(sg))                       # clean again: the old NaN is not re-reported
    layer.check_nan = False
    try:
        us, *_ = env.safe_step(layer, _cuda(ac), bad, _cuda(sg))
        assert torch.isnan(us[B // 2]).all() and layer.solver_stats()["nan"] >= 1
    finally:
        layer.check_nan = True
    _, layer_c = cars
    stc, acc, muc, sgc, t = O.synth_cars(B, seed=12)
    envc = S.SimulatedCarsEnv(num_envs=B)
    envc.state = _cuda(stc)
    badc = _cuda(sgc).clone()
    badc[B // 2, 7] = float("nan")
    with pytest.raises(Exception, match="QP Failed to solve"):
        envc.safe_step(layer_c, _cuda(acc), badc)


def test_unknown_dynamics_mode_raises(S):
    env = types.SimpleNamespace(dynamics_mode="Quadrotor", safe_action_space=types.SimpleNamespace(
        low=np.zeros(2, np.float32), high=np.ones(2, np.float32)), action_space=types.SimpleNamespace(shape=(2,)))
    with pytest.raises(Exception, match="Dynamics mode not supported."):
        S.CBFQPLayer(env, _args())


def test_inputs_not_mutated_and_cpu_tensors_accepted(uni, golden):
    env, layer = uni
    g = golden("unicycle_layer_b256.npz")
    ins = [tt(g[k].copy()) for k in ("state", "action", "mean", "sigma")]      # CPU tensors
    copies = [t.clone() for t in ins]
    out = layer.get_safe_action(*ins)
    assert out.device.type == "cpu"
    for a, b in zip(ins, copies):
        assert torch.equal(a, b)
    assert np.abs(out.numpy() - g["safe_action"]).max() < ACT_TOL


# ----------------------------------------------------------------------------------------------------- backward
@pytest.mark.parametrize("which,name", [("uni", "unicycle_layer_b256.npz"), ("cars", "cars_layer_b512.npz")])
def test_gradient_vs_golden(request, golden, which, name):
    env, layer = request.getfixturevalue(which)
    g = golden(name)
    a = _cuda(g["action"]).requires_grad_(True)
    out = layer.get_safe_action(_cuda(g["state"]), a, _cuda(g["mean"]), _cuda(g["sigma"]))
    (out * _cuda(g["grad_w"])).sum().backward()
    gr = a.grad.cpu().numpy()
    ref = g["grad_action"]
    rel = np.linalg.norm(gr - ref) / max(np.linalg.norm(ref), 1e-6)
    assert rel < GRAD_TOL, rel
    # per-instance: all but clamp-saturated / weakly active rows agree tightly
    e = np.abs(gr - ref).max(1)
    assert np.quantile(e, 0.95) < 1e-3 * max(1.0, np.abs(ref).max())


@pytest.mark.parametrize("mode", ["Unicycle", "SimulatedCars"])
def test_gradient_vs_oracle_autograd(request, mode):
    env, layer = request.getfixturevalue("uni" if mode == "Unicycle" else "cars")
    B = 512
    if mode == "Unicycle":
        st, ac, mu, sg = O.synth_unicycle(B, seed=31, hazard_frac=0.5)
    else:
        st, ac, mu, sg, _ = O.synth_cars(B, seed=31)
    w = np.random.default_rng(5).normal(size=ac.shape).astype(np.float32)
    a_ref = tt(ac).clone().requires_grad_(True)
    fo = O.safe_action(mode, tt(st), a_ref, tt(mu), tt(sg), gamma_b=20.0)       # reference f32 assembly + qpth (f64)
    (fo * tt(w)).sum().backward()
    a = _cuda(ac).requires_grad_(True)
    out = layer.get_safe_action(_cuda(st), a, _cuda(mu), _cuda(sg))
    (out * _cuda(w)).sum().backward()
    gr, ref = a.grad.cpu().numpy(), a_ref.grad.numpy()
    rel = np.linalg.norm(gr - ref) / max(np.linalg.norm(ref), 1e-6)
    assert rel < GRAD_TOL, rel
    # no-grad path allocates no saved tensors and matches the grad path bit for bit
    with torch.no_grad():
        out2 = layer.get_safe_action(_cuda(st), _cuda(ac), _cuda(mu), _cuda(sg))
    assert torch.equal(out2, out.detach())


@pytest.mark.parametrize("mode", ["Unicycle", "SimulatedCars"])
def test_gradient_tile_kernel_paths(request, mode):
    """The compact backward at a size that takes the TMA tile kernel (full 512-instance tiles) plus a ragged tail:
    bit-identical to the generic one-instance-per-thread kernel (reached through a 4-byte-misaligned view of the same
    data), within GRAD_TOL of the dense qpth-clamp backward on saved x / lam / slack, and exactly the clamp mask on the
    trivial instances."""
    env, layer = request.getfixturevalue("uni" if mode == "Unicycle" else "cars")
    B = 512 * 9 + 77
    if mode == "Unicycle":
        st, ac, mu, sg = O.synth_unicycle(B, seed=77, hazard_frac=0.5)
    else:
        st, ac, mu, sg, _ = O.synth_cars(B, seed=77)
    w = np.random.default_rng(9).normal(size=ac.shape).astype(np.float32)
    st, ac, mu, sg, w = (_cuda(a) for a in (st, ac, mu, sg, w))
    out, meta = layer._forward_meta(st, ac, mu, sg)
    ga = layer._backward_meta(st, ac, mu, sg, meta, w)

    def shifted(t):  # same values at an address that is 4 bytes off a 16-byte boundary: the tile kernel does not qualify
        buf = torch.empty(t.numel() + 1, dtype=t.dtype, device=t.device)
        v = buf[1:].view(t.shape)
        v.copy_(t)
        assert v.data_ptr() % 16 != 0
        return v
    gb = layer._backward_meta(shifted(st), ac, mu, sg, meta, w)
    assert torch.equal(ga, gb)
    out2, x, lam, slack = layer._forward_raw(st, ac, mu, sg, save=True)
    assert torch.equal(out, out2)
    gd = layer._backward_raw(st, ac, mu, sg, x, lam, slack, w)
    rel = float((ga - gd).norm() / gd.norm())
    assert rel < GRAD_TOL, rel
    triv = (meta >> 16) == 0
    assert 0.2 < float(triv.float().mean()) < 0.95
    lo, hi = layer.u_min.to(ac.device), layer.u_max.to(ac.device)
    mask = ((ac >= lo) & (ac <= hi)).float()
    assert torch.equal(ga[triv], (w * mask)[triv])


# ------------------------------------------------------------------------------------- arbitrary hazard sets (6 .. 12)
@pytest.mark.parametrize("name", ["unicycle_layer_7haz_b256.npz", "unicycle_layer_11haz_b128.npz"])
def test_general_hazard_count_vs_golden(S, golden, name):
    """The reference sizes its layer from len(env.hazards_locations) (diff_cbf_qp.py:35,243-261).  Fixtures: the
    reference's OWN CBFQPLayer on a 7- and an 11-hazard env (oracle/make_golden.py): constraint shapes and values, safe
    actions vs the golden outputs and vs the exact optimum, gradients, the 1-D contract, NaN behaviour."""
    g = golden(name)
    env = S.UnicycleEnv(num_envs=4, precision="f32")
    env.hazards_locations = g["hazards"]
    layer = S.CBFQPLayer(env, _args(), gamma_b=float(g["gamma_b"]), k_d=3.0, l_p=0.03)
    K = g["hazards"].shape[0]
    assert layer.num_cbfs == K and layer.num_ineq_constraints == K + 4
    ins = [_cuda(g[k]) for k in ("state", "action", "mean", "sigma")]
    P, q, G, h = layer.get_cbf_qp_constraints(*ins)
    assert G.shape == g["G"].shape == (g["state"].shape[0], K + 4, 3) and h.shape == g["h"].shape
    n = np.maximum(np.abs(g["G"]).max(2), np.abs(g["h"]))
    assert (np.abs(G.cpu().numpy() - g["G"]) / n[:, :, None]).max() < 5e-7
    assert (np.abs(h.cpu().numpy() - g["h"]) / n).max() < 3e-6
    np.testing.assert_array_equal(P.cpu().numpy(), g["P"])
    out = layer.get_safe_action(*ins)
    assert np.abs(out.cpu().numpy() - g["safe_action"]).max() < ACT_TOL
    ex = np.clip(g["action"] + g["x_exact"][:, :2].astype(np.float32), -2.5, 2.5)
    assert np.abs(out.cpu().numpy() - ex).max() < ACT_TOL
    st = layer.solver_stats()
    assert st["nan"] == 0 and st["uncertified"] == 0 and 0 < st["trivial"] < g["state"].shape[0]
    a = _cuda(g["action"]).requires_grad_(True)
    o2 = layer.get_safe_action(ins[0], a, ins[2], ins[3])
    assert torch.equal(o2.detach(), out)
    (o2 * _cuda(g["grad_w"])).sum().backward()
    rel = np.linalg.norm(a.grad.cpu().numpy() - g["grad_action"]) / np.linalg.norm(g["grad_action"])
    assert rel < GRAD_TOL, rel
    o1 = layer.get_safe_action(*(t[0] for t in ins))
    assert o1.shape == (2,) and torch.equal(o1, out[0])
    # solve_qp / cbf_layer on the (K + 4)-row system (padded to the next instantiated row count inside)
    xq = layer.solve_qp(_cuda(g["P"]), _cuda(g["q"]), _cuda(g["G"].copy()), _cuda(g["h"]))
    assert xq.shape == (g["G"].shape[0], 2) and np.abs(xq.cpu().numpy() - g["x_exact"][:, :2]).max() < ACT_TOL
    hq = _cuda(g["hn"]).double().requires_grad_(True)
    xs = layer.cbf_layer(_cuda(g["P"]).double(), _cuda(g["q"]).double(), _cuda(g["Gn"]).double(), hq)
    xs.sum().backward()
    assert hq.grad.shape == hq.shape and torch.isfinite(hq.grad).all()
    bad = ins[0].clone()
    bad[3, 0] = float("nan")
    with pytest.raises(Exception, match="QP Failed to solve"):
        layer.get_safe_action(bad, *ins[1:])
    # the fused env step is built for the reference env's 5 hazards: a clear error, not a wrong answer
    with pytest.raises(ValueError):
        env.safe_step(layer, ins[1][:4], ins[2][:4], ins[3][:4])


def test_more_than_twelve_hazards_raises(S):
    env = S.UnicycleEnv(num_envs=2, precision="f32")
    env.hazards_locations = np.random.default_rng(0).uniform(-3, 3, (13, 2))
    layer = S.CBFQPLayer(env, _args(), gamma_b=20, k_d=3.0, l_p=0.03)
    z = torch.zeros(2, 3, device="cuda")
    with pytest.raises(ValueError):
        layer.get_safe_action(z, torch.zeros(2, 2, device="cuda"), z, z)


# ----------------------------------------------------------------------------------------------------- generic QP API
@pytest.mark.parametrize("which,name", [("uni", "unicycle_layer_b256.npz"), ("cars", "cars_layer_b512.npz")])
def test_solve_qp_api(request, golden, which, name):
    env, layer = request.getfixturevalue(which)
    g = golden(name)
    nu = g["action"].shape[1]
    Gs = _cuda(g["G"].copy())
    x = layer.solve_qp(_cuda(g["P"]), _cuda(g["q"]), Gs, _cuda(g["h"]))
    assert x.shape == (g["G"].shape[0], nu) and x.dtype == torch.float32
    assert np.abs(x.cpu().numpy() - g["x_exact"][:, :nu]).max() < ACT_TOL
    # like the reference, Gs was normalised in place (diff_cbf_qp.py:105)
    assert np.abs(Gs.cpu().numpy() - g["Gn"]).max() < 1e-6
    # cbf_layer with gradients w.r.t. G and h vs the restated qpth backward
    from oracle import qpth_pdipm
    Gn = tt(g["Gn"]).double().requires_grad_(True); hn = tt(g["hn"]).double().requires_grad_(True)
    P, q = tt(g["P"]).double(), tt(g["q"]).double()
    e = torch.empty(0, dtype=torch.float64)
    xr = qpth_pdipm.QPFunction(eps=1e-10, notImprovedLim=10, maxIter=100)(P, q, Gn, hn, e, e)
    w = torch.from_numpy(np.random.default_rng(9).normal(size=tuple(xr.shape)))
    (xr * w).sum().backward()
    Gc = _cuda(g["Gn"]).double().requires_grad_(True); hc = _cuda(g["hn"]).double().requires_grad_(True)
    xc = layer.cbf_layer(P.cuda(), q.cuda(), Gc, hc)
    (xc.double() * w.cuda()).sum().backward()
    assert np.abs(xc.detach().cpu().numpy() - xr.detach().numpy()).max() < 1e-5
    for mine, ref in ((hc.grad.cpu(), hn.grad), (Gc.grad.cpu(), Gn.grad)):
        rel = (mine - ref).norm() / ref.norm().clamp_min(1e-9)
        assert rel < GRAD_TOL, rel


@pytest.mark.parametrize("nz,m", [(3, 9), (2, 4)])
def test_generic_qp_dense_Q_and_p_vs_exact_and_qpth_backward(uni, nz, m):
    """cbf_layer on QPs the layer itself never builds: dense SPD Q, non-zero p, dense G (the Cholesky change of
    variables and the dQ / dp gradients of the generic kernel)."""
    from oracle import qpth_pdipm
    env, layer = uni
    B = 4000
    rng = np.random.default_rng(nz * 100 + m)
    L = rng.normal(size=(B, nz, nz)) * 0.5 + np.eye(nz)[None] * 1.5
    Q = L @ L.transpose(0, 2, 1) + 0.1 * np.eye(nz)[None]
    p = rng.normal(size=(B, nz))
    G = rng.normal(size=(B, m, nz))
    x0 = rng.normal(size=(B, nz))
    h = np.einsum("bmj,bj->bm", G, x0) + rng.uniform(0.05, 2.0, size=(B, m))      # strictly feasible at x0
    xe, lam, act, viol = exact_qp.solve_exact(Q, p, G, h)
    assert viol.max() < 1e-8
    Qc, pc, Gc, hc = (torch.from_numpy(a).cuda().requires_grad_(True) for a in (Q, p, G, h))
    x = layer.cbf_layer(Qc, pc, Gc, hc)
    assert x.dtype == torch.float32 and np.abs(x.detach().cpu().numpy() - xe).max() < 2e-5 * max(1.0, np.abs(xe).max())
    assert layer.solver_stats()["nan"] == 0
    # gradients of a random linear functional vs the restated qpth backward evaluated at the exact solution
    w = rng.normal(size=(B, nz))
    (x.double() * torch.from_numpy(w).cuda()).sum().backward()
    slack = h - np.einsum("bmj,bj->bm", G, xe)
    dQ, dp, dG, dh = qpth_pdipm.pdipm_backward(tt(Q), tt(G), tt(xe), tt(lam), tt(np.maximum(slack, 0.0)), tt(w))
    for mine, ref in ((Qc.grad, dQ), (pc.grad, dp), (Gc.grad, dG), (hc.grad, dh)):
        rel = (mine.cpu() - ref).norm() / ref.norm().clamp_min(1e-9)
        assert rel < GRAD_TOL, rel
    with pytest.raises(NotImplementedError):
        layer.cbf_layer(torch.eye(4)[None].cuda().double(), torch.zeros(1, 4).cuda().double(),
                        torch.zeros(1, 5, 4).cuda().double(), torch.ones(1, 5).cuda().double())


# ----------------------------------------------------------------------------------------------------- dynamics
def test_unicycle_env_f64_vs_golden_trajectory(S, golden):
    g = golden("unicycle_env_traj.npz")
    env = S.UnicycleEnv()
    obs = env.reset()
    assert isinstance(obs, np.ndarray) and obs.shape == (7,)
    np.testing.assert_allclose(obs, g["obs0"], rtol=0, atol=1e-14)
    for k in range(len(g["reward"])):
        obs, r, d, info = env.step(g["actions"][k])
        assert isinstance(r, float) and isinstance(d, bool) and isinstance(info, dict)
        np.testing.assert_allclose(obs, g["obs"][k], rtol=1e-12, atol=1e-12)
        np.testing.assert_allclose(env.state, g["state"][k], rtol=1e-12, atol=1e-12)
        assert abs(r - g["reward"][k]) < 1e-11
        assert d == bool(g["done"][k])
        assert ("goal_met" in info) == bool(g["goal_met"][k])
        assert ("cost" in info) == (g["cost"][k] > 0) and abs(info.get("cost", 0.0) - g["cost"][k]) < 1e-15
    assert env.episode_step == len(g["reward"])


def test_cars_env_f64_vs_golden_trajectory(S, golden):
    g = golden("cars_env_traj.npz")
    np.random.seed(0)
    env = S.SimulatedCarsEnv()
    np.random.seed(0)
    obs = env.reset()
    np.testing.assert_allclose(obs, g["obs0"], rtol=0, atol=1e-15)
    for k in range(len(g["reward"])):
        obs, r, d, info = env.step(g["actions"][k])
        np.testing.assert_allclose(obs, g["obs"][k], rtol=1e-12, atol=1e-12)
        assert abs(r - g["reward"][k]) < 1e-15 and d == bool(g["done"][k])
        assert abs(info["cost"] - g["cost"][k]) < 1e-15 and info["goal_met"] is False
    assert abs(env.t - g["t"][-1]) < 1e-12


def test_env_f32_single_step_vs_oracle(S):
    B = 65536
    rng = np.random.default_rng(77)
    # Unicycle
    env = S.UnicycleEnv(num_envs=B)
    st = np.stack([rng.uniform(-3, 3, B), rng.uniform(-3, 3, B), rng.uniform(-np.pi, np.pi, B)], 1).astype(np.float32)
    env.state = _cuda(st)
    a = rng.uniform(-1.5, 1.5, (B, 2)).astype(np.float32)
    last = O.unicycle_goal_dist(st.astype(np.float64))
    ref = O.unicycle_env_step(st.astype(np.float64), a.astype(np.float64), np.zeros(B, np.int64), last)
    obs, rew, done, info = env.step(_cuda(a))
    new = env.state.cpu().numpy()
    scale = np.maximum(np.abs(ref["state"]), 1.0)
    assert (np.abs(new - ref["state"]) / scale).max() < DYN_RTOL
    oerr = np.abs(obs.cpu().numpy() - ref["obs"])
    assert oerr[:, [0, 1, 2, 3, 6]].max() < 2e-6
    # the compass is (goal - p)/|goal - p| rotated: float32 positions (ulp 2.4e-7 at |p| ~ 3) make it ill-conditioned
    # like 1/dist close to the goal
    assert (oerr[:, 4:6] * np.maximum(ref["last_goal_dist"], 1e-3)[:, None]).max() < 2e-6
    # reward is a difference of two O(5) distances: float32 carries it to ~1e-6 absolute
    assert np.abs(rew.cpu().numpy() - ref["reward"]).max() < 2e-6
    near_goal = np.abs(ref["last_goal_dist"] - 0.3) < 1e-5
    assert ((done.cpu().numpy() == ref["done"]) | near_goal).all()
    d2 = ((ref["state"][:, None, :2] - O.UNICYCLE["hazards_locations"][None]) ** 2).sum(2)
    near_hz = (np.abs(d2 - 0.36) < 1e-5).any(1)
    assert ((np.abs(info["cost"].cpu().numpy() - ref["cost"]) < 1e-7) | near_hz).all()
    # Cars
    envc = S.SimulatedCarsEnv(num_envs=B)
    stc, acc, _, _, t = O.synth_cars(B, seed=3)
    envc.state = _cuda(stc)
    envc._t.copy_(_cuda(t))
    refc = O.cars_env_step(stc.astype(np.float64), acc.astype(np.float64), t.astype(np.float64), np.zeros(B, np.int64))
    obs, rew, done, info = envc.step(_cuda(acc))
    keep = O.cars_threshold_margin(stc) > 1e-4
    newc = envc.state.cpu().numpy()
    assert (np.abs(newc - refc["state"])[keep] / np.maximum(np.abs(refc["state"][keep]), 1.0)).max() < DYN_RTOL
    assert (np.abs(obs.cpu().numpy() - refc["obs"])[keep] / np.maximum(np.abs(refc["obs"][keep]), 1.0)).max() < DYN_RTOL
    assert np.abs(rew.cpu().numpy() - refc["reward"]).max() < 1e-8


def test_predict_next_state_vs_golden(S, golden):
    g = golden("dynamics_prior.npz")
    for mode, envc, k in (("Unicycle", S.UnicycleEnv, "unicycle"), ("SimulatedCars", S.SimulatedCarsEnv, "simulatedcars")):
        dm = S.DynamicsModel(envc(), _args())
        t = g.get(k + "_t")
        nxt, std, tn = dm.predict_next_state(g[k + "_state"], g[k + "_action"], t_batch=t)
        assert isinstance(nxt, np.ndarray) and nxt.dtype == np.float64
        np.testing.assert_allclose(nxt, g[k + "_next"], rtol=1e-13, atol=1e-12)
        np.testing.assert_allclose(std, g[k + "_std_dt"], atol=1e-15)
        if t is not None:
            np.testing.assert_allclose(tn, g[k + "_t_next"])
        np.testing.assert_allclose(dm.get_obs(g[k + "_state"]), g[k + "_obs"], atol=1e-14)
        np.testing.assert_allclose(dm.get_state(g[k + "_obs"]), g[k + "_state_from_obs"], atol=1e-12)
        m, s = dm.predict_disturbance(g[k + "_state"])
        np.testing.assert_allclose(m, g[k + "_dist_mean"]); np.testing.assert_allclose(s, g[k + "_dist_std"])
        # float32 device tensors stay on the device and meet the 1e-6 relative bar
        st32 = _cuda(g[k + "_state"].astype(np.float32)); u32 = _cuda(g[k + "_action"].astype(np.float32))
        t32 = None if t is None else _cuda(t.astype(np.float32))
        n32, s32, _ = dm.predict_next_state(st32, u32, t_batch=t32)
        assert n32.is_cuda and n32.dtype == torch.float32
        keep = np.ones(len(nxt), bool) if t is None else O.cars_threshold_margin(g[k + "_state"]) > 1e-4
        rel = np.abs(n32.cpu().numpy() - g[k + "_next"]) / np.maximum(np.abs(g[k + "_next"]), 1.0)
        assert rel[keep].max() < 2 * DYN_RTOL


# ----------------------------------------------------------------------------------------------------- fused step
@pytest.mark.parametrize("B", [32768, 1000, 31, 400017])
def test_fused_safe_step_equals_layer_then_env(S, uni, cars, B):
    """One fused launch == get_safe_action followed by env.step, bit for bit, also for ragged sizes (1000 = 31 bulk-staged
    tiles + a partial one; 31 = no full tile at all; 400017 = every persistent warp walks several tiles, the steady state
    of the merged finish / inline solve) -- the fused SimulatedCars kernel solves inline and finishes full tiles through
    its coalesced warp-collective path (the layer kernel goes through the problem ring: an independent path), everything
    else through the per-lane path."""
    env_u, layer_u = uni
    st, ac, mu, sg = O.synth_unicycle(B, seed=9)
    a = S.UnicycleEnv(num_envs=B); b = S.UnicycleEnv(num_envs=B)
    a.state = _cuda(st); b.state = _cuda(st)
    us = layer_u.get_safe_action(_cuda(st), _cuda(ac), _cuda(mu), _cuda(sg))
    o1, r1, d1, i1 = a.step(us)
    us2, o2, r2, d2, i2 = b.safe_step(layer_u, _cuda(ac), _cuda(mu), _cuda(sg))
    assert torch.equal(us, us2) and torch.equal(o1, o2) and torch.equal(r1, r2) and torch.equal(d1, d2.bool())
    assert torch.equal(a.state, b.state) and torch.equal(i1["cost"], i2["cost"])
    # the status array is optional (the two-per-lane kernel keeps no per-instance class bytes without it): same results
    # with it, and the classes are those of the layer kernel
    c = S.UnicycleEnv(num_envs=B)
    c.state = _cuda(st)
    us3, o3, r3, d3, i3 = c.safe_step(layer_u, _cuda(ac), _cuda(mu), _cuda(sg), want_status=True)
    assert torch.equal(us3, us2) and torch.equal(o3, o2) and torch.equal(c.state, b.state)
    layer_u._forward_raw(_cuda(st), _cuda(ac), _cuda(mu), _cuda(sg), want_status=True)
    assert torch.equal(i3["status"], layer_u._last_status) and int((i3["status"] == 1).sum()) > 0
    env_c, layer_c = cars
    stc, acc, muc, sgc, t = O.synth_cars(B, seed=9)
    a = S.SimulatedCarsEnv(num_envs=B); b = S.SimulatedCarsEnv(num_envs=B)
    for e in (a, b):
        e.state = _cuda(stc); e._t.copy_(_cuda(t))
    us = layer_c.get_safe_action(_cuda(stc), _cuda(acc), _cuda(muc), _cuda(sgc))
    o1, r1, d1, i1 = a.step(us)
    us2, o2, r2, d2, i2 = b.safe_step(layer_c, _cuda(acc), _cuda(sgc), want_status=True)
    assert torch.equal(us, us2) and torch.equal(o1, o2) and torch.equal(r1, r2) and torch.equal(a.state, b.state)
    # (B >= 4096: the ring-compacted kernel k_cars2 on the full tiles, k_safe on the ragged rest)
    assert torch.equal(d1, d2.bool()) and torch.equal(i1["cost"], i2["cost"]) and torch.equal(a._t, b._t)
    assert torch.equal(a._step, b._step)
    layer_c._forward_raw(_cuda(stc), _cuda(acc), _cuda(muc), _cuda(sgc), want_status=True)
    assert torch.equal(i2["status"], layer_c._last_status) and int((i2["status"] == 1).sum()) > 0
    # a second step from the stepped state (the in-place state rows the bulk stores wrote are what the next launch reads)
    us = layer_c.get_safe_action(a.state, _cuda(acc), _cuda(muc), _cuda(sgc))
    o1, r1, d1, i1 = a.step(us)
    us2, o2, r2, d2, i2 = b.safe_step(layer_c, _cuda(acc), _cuda(sgc))
    assert torch.equal(us, us2) and torch.equal(o1, o2) and torch.equal(a.state, b.state) and torch.equal(a._t, b._t)


def test_fused_step_merged_finish_structured_need_patterns(S, uni):
    """The Unicycle presolve kernel finishes a solved instance on a lane that is idle in a LATER tile (finish ring +
    cp.async hand-over, `k_safe` merged finish).  Stress the hand-over with need/trivial patterns per tile ROUND (a
    persistent warp walks tiles w, w + 2368, w + 2*2368 ...): rounds where every instance needs a solve, rounds where
    none does, mixtures, and a ragged tail; then (i) the fused step must commute with a permutation of the instances
    (another lane / ring slot finishes every instance), (ii) equal get_safe_action + env.step, (iii) agree with the
    interior-point mode, whose finish path is the separate one."""
    env, layer = uni
    pool = 1 << 21
    st, ac, mu, sg = O.synth_unicycle(pool, seed=4242)
    status = _forward_with_aux(layer, st, ac, mu, sg)[4]
    need_idx, triv_idx = np.flatnonzero(status != 0), np.flatnonzero(status == 0)
    rng = np.random.default_rng(7)
    rnd = 2368 * 32                                    # instances per round of the persistent grid (148 x 4 x 4 warps)
    order, un, ut = [], 0, 0
    for frac in (1.0, 1.0, 0.0, 0.97, 0.0, 0.5, 0.9, 0.0, 1.0, 0.35):
        k = int(round(frac * rnd)) if frac < 1.0 else rnd
        pick = np.concatenate([need_idx[un:un + k], triv_idx[ut:ut + rnd - k]])
        un, ut = un + k, ut + rnd - k
        order.append(rng.permutation(pick) if 0.0 < frac < 1.0 else pick)
    order.append(np.concatenate([need_idx[un:un + 700], triv_idx[ut:ut + 301]]))   # ragged tail: 1001 instances
    order = np.concatenate(order)
    assert len(np.unique(order)) == len(order)
    st, ac, mu, sg = st[order], ac[order], mu[order], sg[order]
    B = len(order)

    def fused(perm=None, solver="presolve"):
        e = S.UnicycleEnv(num_envs=B)
        sel = (lambda a: a) if perm is None else (lambda a: a[perm])
        e.state = _cuda(sel(st))
        layer.solver = solver
        try:
            us, obs, rew, done, info = e.safe_step(layer, _cuda(sel(ac)), _cuda(sel(mu)), _cuda(sel(sg)))
            torch.cuda.synchronize()
        finally:
            layer.solver = "presolve"
        return [t.cpu().numpy().copy() for t in (us, obs, rew, done, info["cost"], e.state)]

    base = fused()
    assert not np.isnan(base[0]).any() and np.isfinite(base[5]).all()
    perm = rng.permutation(B)
    for a, b in zip(fused(perm), base):
        np.testing.assert_array_equal(a, b[perm])
    e2 = S.UnicycleEnv(num_envs=B)
    e2.state = _cuda(st)
    us = layer.get_safe_action(_cuda(st), _cuda(ac), _cuda(mu), _cuda(sg))
    o1, r1, d1, i1 = e2.step(us)
    np.testing.assert_array_equal(us.cpu().numpy(), base[0])
    np.testing.assert_array_equal(o1.cpu().numpy(), base[1])
    np.testing.assert_array_equal(e2.state.cpu().numpy(), base[5])
    alt = fused(solver="pdipm")
    assert np.abs(alt[0] - base[0]).max() < 1e-5 and np.abs(alt[5] - base[5]).max() < 1e-5
    assert (alt[0] == base[0]).all(1).mean() > 0.99


# ----------------------------------------------------------------------------------------------------- full size
def test_full_size_properties(uni):
    """BASELINE config 4 size (1M instances): size-independent properties instead of an oracle run."""
    env, layer = uni
    B = 1 << 20
    st, ac, mu, sg = O.synth_unicycle(B, seed=12345)
    out, x, lam, slack, status, iters = _forward_with_aux(layer, st, ac, mu, sg)
    assert not np.isnan(out).any() and (status <= 2).all()
    assert np.abs(out).max() <= 2.5
    # slack the kernel certified is the row residual of its own normalised data: every row satisfied
    assert slack.min() > ROW_TOL and lam.min() >= 0
    # trivial instances are passed through untouched
    triv = status == 0
    np.testing.assert_array_equal(out[triv], ac[triv])
    # permutation equivariance + determinism (no cross-instance coupling, unlike qpth's batch-global stop)
    perm = np.random.default_rng(1).permutation(B)
    out_p = _forward_with_aux(layer, st[perm], ac[perm], mu[perm], sg[perm])[0]
    np.testing.assert_array_equal(out_p, out[perm])
    # spot-check 4096 of them against the exact oracle
    idx = np.random.default_rng(2).choice(B, 4096, replace=False)
    fe = O.safe_action("Unicycle", tt(st[idx]), tt(ac[idx]), tt(mu[idx]), tt(sg[idx]), solver="exact", gamma_b=20.0)
    f64 = O.safe_action("Unicycle", tt(st[idx]), tt(ac[idx]), tt(mu[idx]), tt(sg[idx]), solver="exact",
                        assembly_dtype=torch.float64, gamma_b=20.0)
    ok = np.abs(fe.numpy() - f64.numpy()).max(1) <= 2e-5
    assert np.abs(out[idx] - fe.numpy())[ok].max() < ACT_TOL


def test_config5_cars_sweep_top_size_properties(cars):
    """BASELINE config 5's largest size (16 Mi SimulatedCars QPs on one GPU): every row feasible, duals valid, and a
    seeded sub-sample equal to the same rows solved in a small launch."""
    env, layer = cars
    B = 1 << 24
    g = torch.Generator(device="cuda").manual_seed(5)
    t = 6 * torch.rand(B, generator=g, device="cuda")
    st = torch.zeros(B, 10, device="cuda")
    st[:, 0::2] = torch.tensor([34., 28., 22., 16., 10.], device="cuda") + 30 * t[:, None] + 1.5 * torch.randn(B, 5, generator=g, device="cuda")
    st[:, 1::2] = 30 + 2 * torch.randn(B, 5, generator=g, device="cuda")
    st[:, 7] += 3
    ac = 2 * torch.rand(B, 1, generator=g, device="cuda") - 1
    sg = torch.zeros(B, 10, device="cuda")
    sg[:, 1::2] = 0.2 * torch.rand(B, 5, generator=g, device="cuda")
    mu = torch.zeros(B, 10, device="cuda")
    out, x, lam, slack = layer._forward_raw(st, ac, mu, sg, save=True, want_status=True)
    stats = layer.solver_stats()
    assert stats["nan"] == 0 and stats["uncertified"] == 0
    assert float(slack.min()) > ROW_TOL and float(lam.min()) >= 0 and float(out.abs().max()) <= 10.0
    idx = torch.randperm(B, generator=g, device="cuda")[:5000]
    small = layer.get_safe_action(st[idx], ac[idx], mu[idx], sg[idx])
    assert torch.equal(small, out[idx])
    fe = O.safe_action("SimulatedCars", st[idx].cpu(), ac[idx].cpu(), mu[idx].cpu(), sg[idx].cpu(), solver="exact", gamma_b=20.0)
    keep = O.cars_threshold_margin(st[idx].cpu().numpy()) > 1e-4
    assert np.abs(small.cpu().numpy() - fe.numpy())[keep].max() < ACT_TOL


def test_host_buffer_entry_matches_device_entry(S, uni, cars):
    import ctypes as C
    from sac_rcbf_b200 import _lib
    lib = _lib.load()
    for (env, layer), synth in ((uni, O.synth_unicycle), (cars, O.synth_cars)):
        B = 100003     # ragged on purpose
        arrs = synth(B, seed=4)
        st, ac, mu, sg = arrs[:4]
        ref = layer.get_safe_action(_cuda(st), _cuda(ac), _cuda(mu), _cuda(sg)).cpu().numpy()
        pin = [torch.from_numpy(a).pin_memory() for a in (st, ac, mu, sg)]
        out = torch.empty(ref.shape, dtype=torch.float32).pin_memory()
        nf = C.c_int32(-1)
        if env.dynamics_mode == "Unicycle":
            rc = lib.rcbf_unicycle_safe_action_host(_lib.ptr(pin[0]), _lib.ptr(pin[1]), _lib.ptr(pin[2]), _lib.ptr(pin[3]),
                                                    B, layer._params(), _lib.ptr(out), C.byref(nf), 0, 7)
        else:
            rc = lib.rcbf_cars_safe_action_host(_lib.ptr(pin[0]), _lib.ptr(pin[1]), _lib.ptr(pin[3]), B, layer._params(),
                                                _lib.ptr(out), C.byref(nf), 0, 7)
        assert rc == 0 and nf.value == 0
        np.testing.assert_array_equal(out.numpy(), ref)


def test_fused_step_host_entry_matches_device_entry(S, uni, cars):
    """rcbf_*_safe_step_host (pinned host tensors in/out, env state on the GPU) == the device-resident safe_step."""
    B = 100003
    env_u, layer_u = uni
    st, ac, mu, sg = O.synth_unicycle(B, seed=12)
    a = S.UnicycleEnv(num_envs=B); b = S.UnicycleEnv(num_envs=B)
    a.state = _cuda(st); b.state = _cuda(st)
    us, obs, rew, done, info = a.safe_step(layer_u, _cuda(ac), _cuda(mu), _cuda(sg))
    pin = lambda x: torch.from_numpy(x).pin_memory()  # noqa: E731
    out = b.safe_step_host(layer_u, pin(ac), pin(mu), pin(sg), chunks=5)
    assert torch.equal(out["safe_action"], us.cpu()) and torch.equal(out["obs"], obs.cpu())
    assert torch.equal(out["reward"], rew.cpu()) and torch.equal(out["done"], done.cpu())
    assert torch.equal(out["cost"], info["cost"].cpu()) and torch.equal(a.state, b.state)
    env_c, layer_c = cars
    stc, acc, muc, sgc, t = O.synth_cars(B, seed=12)
    a = S.SimulatedCarsEnv(num_envs=B); b = S.SimulatedCarsEnv(num_envs=B)
    for e in (a, b):
        e.state = _cuda(stc); e._t.copy_(_cuda(t))
    us, obs, rew, done, info = a.safe_step(layer_c, _cuda(acc), _cuda(sgc))
    out = b.safe_step_host(layer_c, pin(acc), pin(sgc), chunks=3)
    assert torch.equal(out["safe_action"], us.cpu()) and torch.equal(out["obs"], obs.cpu())
    assert torch.equal(out["reward"], rew.cpu()) and torch.equal(a.state, b.state)


def test_envs_accept_an_unindexed_cuda_device(S, uni, cars):
    """`device="cuda"` / torch.device("cuda") (no index) means the current device, like everywhere in torch."""
    for dev in ("cuda", torch.device("cuda")):
        e = S.UnicycleEnv(num_envs=8, device=dev, auto_reset=True)
        assert e.device.index == torch.cuda.current_device()
        z = torch.zeros(8, 3, device="cuda")
        e.safe_step(uni[1], torch.zeros(8, 2, device="cuda"), z, z)
        e.step(torch.zeros(8, 2, device="cuda"))
        c = S.SimulatedCarsEnv(num_envs=8, device=dev)
        c.safe_step(cars[1], torch.zeros(8, 1, device="cuda"), torch.zeros(8, 10, device="cuda"))
        c.step(torch.zeros(8, device="cuda"))
    torch.cuda.synchronize()


# ----------------------------------------------------------------------------------------------------- numpy layer shim
def test_cascade_layer_vs_oracle(S, golden):
    g = golden("cascade_layer.npz")
    env = S.UnicycleEnv()
    lay = S.CascadeCBFLayer(env, gamma_b=100, k_d=1.5, l_p=0.03)
    for i in range(g["state"].shape[0]):
        P, q, G, h = lay.get_cbf_qp_constraints(g["action"][i], g["state"][i], g["mean"][i], g["sigma"][i])
        n = np.maximum(np.abs(g["G"][i]).max(1), np.abs(g["h"][i]))
        assert (np.abs(G - g["G"][i]) / n[:, None]).max() < 1e-7 and (np.abs(h - g["h"][i]) / n).max() < 1e-7   # float64
        np.testing.assert_allclose(P, g["P"][i])
        u = lay.get_u_safe(g["action"][i], g["state"][i], g["mean"][i], g["sigma"][i])
        ue, _ = O.cascade_u_safe("Unicycle", g["action"][i], g["state"][i], g["mean"][i], g["sigma"][i])
        assert u.shape == (2,) and np.abs(u - ue).max() < 1e-4 * max(1.0, np.abs(ue).max())   # north_star tolerance
    envc = S.SimulatedCarsEnv()
    layc = S.CascadeCBFLayer(envc, gamma_b=100, k_d=1.5)
    for i in range(g["cars_state"].shape[0]):
        u = layc.get_u_safe(g["cars_action"][i], g["cars_state"][i], np.zeros(10), g["cars_sigma"][i])
        ue, _ = O.cascade_u_safe("SimulatedCars", g["cars_action"][i], g["cars_state"][i], np.zeros(10), g["cars_sigma"][i])
        assert u.shape == (1,) and np.abs(u - ue).max() < 1e-4 * max(1.0, np.abs(ue).max())


@pytest.mark.parametrize("B", [1, 512, 4096, 4096 + 17, 65536 + 40])
def test_counters_are_published_by_the_kernel_itself(S, uni, cars, B):
    """check_nan = True (the reference's per-call NaN test): the last block of the call's last kernel writes the counters
    and the call's token into the pinned mirror bound to the workspace -- one launch, no publish kernel.  Sizes cover
    k_safe alone, k_safe2 / k_cars2 alone and either followed by a ragged rest (then the rest's kernel publishes)."""
    env_u, layer_u = uni
    st, ac, mu, sg = O.synth_unicycle(B, seed=21)
    layer_u.check_nan = True
    for _ in range(3):
        out = layer_u.get_safe_action(_cuda(st), _cuda(ac), _cuda(mu), _cuda(sg))
        m = layer_u._mirrors[layer_u._ws.data_ptr()]            # (a fallback to the publish kernel would have dropped it)
        assert int(m[1][0]) == m[2] and layer_u._last_stats[0] == 0
        assert sum(layer_u._last_stats[3:4]) + 0 <= B and m[1][1:9].tolist() == layer_u._ws[:8].tolist()
    e = S.UnicycleEnv(num_envs=B, precision="f32")
    e.state = _cuda(st)
    for _ in range(2):
        e.safe_step(layer_u, _cuda(ac), _cuda(mu), _cuda(sg))
        m = layer_u._mirrors[e._counters.data_ptr()]
        assert int(m[1][0]) == m[2] and m[1][1:9].tolist() == e._counters[:8].tolist()
    ref = layer_u._forward_raw(_cuda(st), _cuda(ac), _cuda(mu), _cuda(sg))[0]
    assert torch.equal(out.reshape(ref.shape), ref)
    env_c, layer_c = cars
    stc, acc, muc, sgc, t = O.synth_cars(B, seed=21)
    layer_c.check_nan = True
    ec = S.SimulatedCarsEnv(num_envs=B, precision="f32")
    ec.state = _cuda(stc); ec._t.copy_(_cuda(t))
    for _ in range(2):
        ec.safe_step(layer_c, _cuda(acc), _cuda(sgc))
        m = layer_c._mirrors[ec._counters.data_ptr()]
        assert int(m[1][0]) == m[2] and m[1][1:9].tolist() == ec._counters[:8].tolist()
    # a NaN input still raises, through the published counter
    bad = ac.copy(); bad[B // 2, 0] = np.nan
    with pytest.raises(Exception, match="QP Failed to solve"):
        layer_u.get_safe_action(_cuda(st), _cuda(bad), _cuda(mu), _cuda(sg))


@pytest.mark.parametrize("B", [4096, 50000 + 19])
def test_cars_layer_ring_kernel_equals_the_one_tile_kernel(cars, B):
    """get_safe_action for SimulatedCars at n >= 4096 runs k_cars2<false> (problem ring + finish lag) on the full tiles;
    the same rows through arrays that are NOT 16-byte aligned take k_safe for every tile.  Safe action, status and the
    meta word the backward reads must agree bit for bit, and so must the gradients computed from them."""
    env_c, layer_c = cars
    stc, acc, muc, sgc, _ = O.synth_cars(B + 1, seed=33)
    d = [_cuda(x) for x in (stc, acc, muc, sgc)]
    al = [x[1:].clone() for x in d]                  # fresh allocations: 16-byte aligned
    un = [x[1:] for x in d]                          # views one row in: 40 / 4 bytes off
    assert al[0].data_ptr() % 16 == 0 and un[0].data_ptr() % 16 != 0
    out_a = layer_c._forward_raw(*al, want_status=True)[0]
    st_a = layer_c._last_status.clone()
    out_u = layer_c._forward_raw(*un, want_status=True)[0]
    assert torch.equal(out_a, out_u) and torch.equal(st_a, layer_c._last_status) and int((st_a == 1).sum()) > B // 10
    oa, ma = layer_c._forward_meta(*al)
    ou, mu_ = layer_c._forward_meta(*un)
    assert torch.equal(oa, out_a) and torch.equal(ou, out_a) and torch.equal(ma, mu_)
    go = torch.randn_like(oa)
    ga = layer_c._backward_meta(*al, ma, go)
    gu = layer_c._backward_meta(*un, mu_, go)
    assert torch.equal(ga, gu)
    k = slice(0, 2000)
    fe = O.safe_action("SimulatedCars", tt(stc[1:][k]), tt(acc[1:][k]), tt(muc[1:][k]), tt(sgc[1:][k]), solver="exact",
                       gamma_b=20.0).numpy()
    keep = O.cars_threshold_margin(stc[1:][k]) > 1e-4
    assert np.abs(out_a[k].cpu().numpy() - fe)[keep].max() < 1e-4


def test_cars_fused_step_full_size_properties(S, cars):
    """BASELINE's 4 Mi-instance size for the fused SimulatedCars step (k_cars2): the result of an instance does not
    depend on where it sits (permutation equivariance: other tile, other ring order, other B-step companions), and a
    seeded sub-sample equals the same rows stepped in a launch small enough to take the one-tile-at-a-time kernel."""
    env_c, layer_c = cars
    B = 1 << 22
    stc, acc, muc, sgc, t = O.synth_cars(B, seed=44)
    d = [_cuda(x) for x in (stc, acc, sgc, t)]
    g = torch.Generator(device="cuda").manual_seed(9)
    perm = torch.randperm(B, generator=g, device="cuda")
    outs = []
    for sel in (None, perm):
        st_, ac_, sg_, t_ = d if sel is None else [x[sel].contiguous() for x in d]
        e = S.SimulatedCarsEnv(num_envs=B)
        e.state = st_; e._t.copy_(t_)
        us, obs, rew, done, info = e.safe_step(layer_c, ac_, sg_)
        outs.append([x.clone() for x in (us, obs, rew, info["cost"], e.state, e._t)])
    for a, b in zip(*outs):
        assert torch.equal(a[perm], b)
    assert float(outs[0][0].abs().max()) <= 10.0 and not bool(torch.isnan(outs[0][1]).any())
    idx = torch.randperm(B, generator=g, device="cuda")[:3000]
    e = S.SimulatedCarsEnv(num_envs=3000)
    e.state = d[0][idx].contiguous(); e._t.copy_(d[3][idx])
    us, obs, rew, done, info = e.safe_step(layer_c, d[1][idx].contiguous(), d[2][idx].contiguous())
    for a, b in zip(outs[0], (us, obs, rew, info["cost"], e.state, e._t)):
        assert torch.equal(a[idx], b)
