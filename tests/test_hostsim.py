"""CPU check of the kernels' per-instance SOURCE (sac_rcbf_b200/csrc/rcbf_core.cuh) compiled for the host by
tests/hostsim -- test infrastructure, not a product path -- against the oracle.  This pins the numerics of the solver
design (float32 interior point + float64 KKT certificate + float64 straggler pass) without a GPU; the `-m gpu` tests
repeat the comparison on the real kernels."""
import numpy as np
import pytest
import torch

from oracle import exact_qp, rcbf_oracle as O
from sac_rcbf_b200 import _params as PR
from tests.hostsim import sim

tt = torch.from_numpy


@pytest.fixture(scope="module", autouse=True)
def _build():
    sim.build()


def test_unicycle_core_vs_oracle():
    B = 20000
    st, ac, mu, sg = O.synth_unicycle(B, seed=1)
    o = sim.unicycle_safe_action(st, ac, mu, sg, PR.unicycle_params(gamma_b=20.0))
    assert (o["status"] <= 2).all()
    # assembly: reference op order -> identical up to the ulp of libm's vs torch's cos/sin
    P, q, G, h = O.assemble_unicycle(tt(st), tt(ac), tt(mu), tt(sg), gamma_b=20.0)
    n = np.maximum(np.abs(G.numpy()).max(2), np.abs(h.numpy()))
    assert (np.abs(o["G"] - G.numpy()) / n[:, :, None]).max() < 5e-7 and (np.abs(o["h"] - h.numpy()) / n).max() < 3e-6
    assert (o["G"] == G.numpy()).mean() > 0.85   # the rest differ by the last ulp of sin/cos
    # solver vs the exact optimum of its own data
    Pd = np.tile(np.diag([1.0, 1e-2, 1e5]), (B, 1, 1))
    xe, lam, act, viol = exact_qp.solve_exact(Pd, np.zeros((B, 3)), o["Gn"].astype(np.float64), o["hn"].astype(np.float64))
    assert np.abs(o["x"] - xe).max() < 1e-6
    xf = o["x"].astype(np.float32).astype(np.float64)
    assert (o["hn"].astype(np.float64) - np.einsum("bmj,bj->bm", o["Gn"].astype(np.float64), xf)).min() > -1e-6
    assert o["lam"].min() >= 0 and np.abs(o["lam"] - lam).max() < 1e-6 * max(1.0, lam.max())
    # end to end vs the oracle on reference-assembled data (ill-conditioned instances flagged like in the GPU test)
    fe = O.safe_action("Unicycle", tt(st), tt(ac), tt(mu), tt(sg), solver="exact", gamma_b=20.0).numpy()
    f64 = O.safe_action("Unicycle", tt(st), tt(ac), tt(mu), tt(sg), solver="exact", assembly_dtype=torch.float64,
                        gamma_b=20.0).numpy()
    ok = np.abs(fe - f64).max(1) <= 2e-5
    assert ok.mean() > 0.999 and np.abs(o["out"] - fe)[ok].max() < 1e-4
    # iteration statistics the roofline accounting relies on
    it = np.where(o["iters"] >= 100, o["iters"] - 100, o["iters"])
    assert it.mean() < 2.0 and (o["iters"] >= 100).mean() < 1e-3


def test_cars_core_vs_oracle():
    B = 20000
    st, ac, mu, sg, _ = O.synth_cars(B, seed=1)
    o = sim.cars_safe_action(st, ac, sg, PR.cars_params(gamma_b=20.0))
    assert (o["status"] <= 2).all()
    P, q, G, h = O.assemble_cars(tt(st), tt(ac), tt(mu), tt(sg), gamma_b=20.0)
    np.testing.assert_array_equal(o["G"], G.numpy())      # no transcendental: bit exact with the reference order
    np.testing.assert_array_equal(o["h"], h.numpy())
    fe = O.safe_action("SimulatedCars", tt(st), tt(ac), tt(mu), tt(sg), solver="exact", gamma_b=20.0).numpy()
    assert np.abs(o["out"] - fe).max() < 1e-4
    xf = o["x"].astype(np.float32).astype(np.float64)
    assert (o["hn"].astype(np.float64) - np.einsum("bmj,bj->bm", o["Gn"].astype(np.float64), xf)).min() > -1e-6


def test_core_on_golden_batches(golden):
    g = golden("unicycle_layer_b256.npz")
    o = sim.unicycle_safe_action(g["state"], g["action"], g["mean"], g["sigma"], PR.unicycle_params(gamma_b=20.0))
    assert np.abs(o["out"] - g["safe_action"]).max() < 1e-4
    g = golden("cars_layer_b512.npz")
    o = sim.cars_safe_action(g["state"], g["action"], g["sigma"], PR.cars_params(gamma_b=20.0))
    assert np.abs(o["out"] - g["safe_action"]).max() < 1e-4


def test_core_kkt_conditions_on_extreme_inputs():
    """Host build of the kernel source on the stress distribution of the GPU suite (actions up to the actuator limits,
    disturbance std up to 15x MAX_STD, all instances around the hazards): whatever status the solver reports as solved
    must BE the optimum of the normalised QP it was given -- stationarity, primal / dual feasibility, complementarity
    in float64 -- including the instances the greedy presolve leaves pending (a drop is needed), which the host driver
    hands to the full chain (float32 interior point + certificate, float64 interior point) like pass 2 does."""
    B = 6000
    rng = np.random.default_rng(11)
    hz = O.UNICYCLE["hazards_locations"]
    idx = rng.integers(0, len(hz), B)
    r, phi = rng.uniform(0.3, 1.2, B), rng.uniform(-np.pi, np.pi, B)
    st = np.stack([hz[idx, 0] + r * np.cos(phi), hz[idx, 1] + r * np.sin(phi), rng.uniform(-np.pi, np.pi, B)], 1)
    ac, mu, sg = rng.uniform(-2.5, 2.5, (B, 2)), rng.uniform(-0.5, 0.5, (B, 3)), rng.uniform(0, 3.0, (B, 3))
    params = PR.unicycle_params(gamma_b=20.0)
    pdiag = np.array([1.0, 1e-2, 1e5])

    def kkt(o, sel):
        Gn, hn = o["Gn"][sel].astype(np.float64), o["hn"][sel].astype(np.float64)
        x, lam = o["x"][sel], o["lam"][sel]
        slack = hn - np.einsum("bmj,bj->bm", Gn, x)
        stat = pdiag[None] * x + np.einsum("bmj,bm->bj", Gn, lam)
        scale = 1.0 + np.abs(pdiag[None] * x).max(1)
        return slack.min(), lam.min(), (np.abs(lam * slack) / (1.0 + lam)).max(), (np.abs(stat).max(1) / scale).max()

    fast = sim.unicycle_safe_action(st, ac, mu, sg, params, mode=0)
    assert (fast["status"] <= 2).all()
    chained = fast["iters"] >= 1000           # the host driver ran the fallback chain for these (pending after pass 1)
    assert 10 < chained.sum() < 0.03 * B
    smin, lmin, comp, stat = kkt(fast, ~chained)
    assert smin > -1e-8 and lmin >= 0.0 and comp < 1e-8 and stat < 1e-7, (smin, lmin, comp, stat)   # multipliers reach 1e5
    smin, lmin, comp, stat = kkt(fast, chained)
    assert smin > -1e-7 and lmin >= -1e-9 and comp < 1e-5 and stat < 1e-5, (smin, lmin, comp, stat)
    # and the interior-point-only mode agrees with the presolve mode
    ipm = sim.unicycle_safe_action(st, ac, mu, sg, params, mode=1)
    assert (ipm["status"] <= 2).all() and np.abs(ipm["out"] - fast["out"]).max() < 1e-5


def test_backward_forms_on_host(golden):
    """The two implicit-KKT backward forms of csrc/rcbf_backward.cuh on the host build: the dense qpth-clamp form on saved
    x / lam / slack and the exact active-set form the compact kernels run (from the certified mask alone) agree with
    each other and with the golden gradients (reference CBFQPLayer + restated qpth backward)."""
    for mode, name in (("uni", "unicycle_layer_b256.npz"), ("cars", "cars_layer_b512.npz")):
        g = golden(name)
        gb = float(g["gamma_b"])
        if mode == "uni":
            gd, ga, status = sim.unicycle_bwd(g["state"], g["action"], g["mean"], g["sigma"], g["grad_w"],
                                              PR.unicycle_params(gamma_b=gb))
        else:
            gd, ga, status = sim.cars_bwd(g["state"], g["action"], g["sigma"], g["grad_w"], PR.cars_params(gamma_b=gb))
        ref = g["grad_action"]
        assert (status <= 1).all()
        assert np.linalg.norm(gd - ref) / np.linalg.norm(ref) < 1e-5
        assert np.linalg.norm(ga - ref) / np.linalg.norm(ref) < 1e-5
    B = 100000
    st, ac, mu, sg = O.synth_unicycle(B, seed=4)
    go = np.random.default_rng(4).standard_normal((B, 2)).astype(np.float32)
    gd, ga, status = sim.unicycle_bwd(st, ac, mu, sg, go, PR.unicycle_params(gamma_b=20.0))
    assert np.linalg.norm(gd - ga) / np.linalg.norm(gd) < 1e-4      # qpth's 1e-8 clamps vs their limit
    triv = status == 0
    assert 0.5 < triv.mean() < 0.8 and np.array_equal(ga[triv], go[triv])   # |u_rl| <= 1 < 2.5: the clamp mask is all-pass


@pytest.mark.parametrize("name", ["unicycle_layer_7haz_b256.npz", "unicycle_layer_11haz_b128.npz"])
def test_general_hazard_count_on_host(golden, name):
    """The row-count-templated source (assemble_unicycle_n, raw-row presolve, certificate, exact active-set backward) at
    NH = 8 / 12 CBF rows, composed like rcbf_general.cu does per thread, against the reference-generated fixtures."""
    g = golden(name)
    out, ga, status = sim.unicycle_general(g["state"], g["action"], g["mean"], g["sigma"], g["grad_w"],
                                           PR.unicycle_params(gamma_b=float(g["gamma_b"])), g["hazards"])
    assert (status <= 1).all() and 0 < (status == 1).sum() < len(status)
    assert np.abs(out - g["safe_action"]).max() < 1e-4
    ex = np.clip(g["action"] + g["x_exact"][:, :2].astype(np.float32), -2.5, 2.5)
    assert np.abs(out - ex).max() < 1e-4
    assert np.linalg.norm(ga - g["grad_action"]) / np.linalg.norm(g["grad_action"]) < 1e-3
