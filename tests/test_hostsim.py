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
