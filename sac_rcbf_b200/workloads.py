"""Synthetic workloads of SURVEY.md section 8(d): the input distributions every benchmark line and parity test draws
from.  Pure numpy, importable without CUDA.  (Workload definitions only -- nothing here evaluates the reference path.)

Unicycle: x, y ~ U(-3, 3) (the arena, unicycle_env.py:24), theta ~ U(-pi, pi), u_RL ~ U(-1, 1)^2 (:21),
mean ~ U(-0.1, 0.1)^3, sigma ~ U(0, 0.2)^3 (<= MAX_STD, dynamics.py:24), plus a hazard-heavy stratum (20 %) placed
0.3 .. 1.1 from a hazard centre so that slack and multi-active cases occur.
SimulatedCars: t ~ U(0, 6), pos = (34, 28, 22, 16, 10) + 30 t + N(0, 1.5^2), vel ~ N(30, 2^2) (car 4: +3),
u_RL ~ U(-1, 1), mean = 0, sigma = U(0, 0.2) on the velocity entries."""
import math

import numpy as np

UNICYCLE = {"hazards_locations": 1.5 * np.array([[0.0, 0.0], [-1.0, 1.0], [-1.0, -1.0], [1.0, -1.0], [1.0, 1.0]])}  # unicycle_env.py:26
CARS = {"init_pos": np.array([34.0, 28.0, 22.0, 16.0, 10.0])}  # simulated_cars_env.py:117


def synth_unicycle(B, seed=12345, hazard_frac=0.2):
    g = np.random.default_rng(seed)
    st = np.stack([g.uniform(-3, 3, B), g.uniform(-3, 3, B), g.uniform(-math.pi, math.pi, B)], axis=1)
    nh = int(B * hazard_frac)
    if nh > 0:
        hz = UNICYCLE["hazards_locations"][g.integers(0, 5, nh)]
        r = g.uniform(0.3, 1.1, nh)
        phi = g.uniform(-math.pi, math.pi, nh)
        st[:nh, 0] = hz[:, 0] + r * np.cos(phi)
        st[:nh, 1] = hz[:, 1] + r * np.sin(phi)
    ac = g.uniform(-1, 1, (B, 2))
    mu = g.uniform(-0.1, 0.1, (B, 3))
    sg = g.uniform(0, 0.2, (B, 3))
    perm = g.permutation(B)
    return tuple(a[perm].astype(np.float32) for a in (st, ac, mu, sg))


def synth_cars(B, seed=12345):
    g = np.random.default_rng(seed)
    t = g.uniform(0, 6, B)
    pos = CARS["init_pos"][None, :] + 30.0 * t[:, None] + g.normal(0, 1.5, (B, 5))
    vel = g.normal(30, 2, (B, 5))
    vel[:, 3] += 3.0
    st = np.zeros((B, 10))
    st[:, 0::2] = pos
    st[:, 1::2] = vel
    ac = g.uniform(-1, 1, (B, 1))
    mu = np.zeros((B, 10))
    sg = np.zeros((B, 10))
    sg[:, 1::2] = g.uniform(0, 0.2, (B, 5))
    return tuple(a.astype(np.float32) for a in (st, ac, mu, sg)) + (t.astype(np.float32),)


def bench_unicycle(n, seed=12345, sets=2, first=None):
    """The bench workload (BASELINE config 4 semantics) as HOST arrays, prefix-stable: instance i's data do not depend on
    n -- every array is drawn from its own Philox stream in instance order -- so `first=m` returns exactly the first m
    instances of the n-instance workload.  bench.py's GPU arm uploads all n; its CPU arm (`--impl reference`,
    `cpu_baseline`) runs on the leading instances of the SAME arrays.
    Returns (state (m,3), [(u_rl (m,2), mean (m,3), sigma (m,3))] * sets), float32."""
    m = n if first is None else min(int(first), n)

    def U(k, cols):
        return np.random.Generator(np.random.Philox(key=[seed, k])).random((m, cols), dtype=np.float32)

    a = U(0, 3)
    st = np.stack([-3 + 6 * a[:, 0], -3 + 6 * a[:, 1], (2 * a[:, 2] - 1) * np.float32(math.pi)], 1).astype(np.float32)
    b = U(1, 4)
    heavy = b[:, 0] < 0.2                                   # hazard-heavy stratum: 0.3 .. 1.1 m from a hazard centre
    hz = UNICYCLE["hazards_locations"][np.minimum((5 * b[:, 1]).astype(np.int64), 4)]
    r, phi = 0.3 + 0.8 * b[:, 2], (2 * b[:, 3] - 1) * np.float32(math.pi)
    st[heavy, 0] = (hz[:, 0] + r * np.cos(phi))[heavy]
    st[heavy, 1] = (hz[:, 1] + r * np.sin(phi))[heavy]
    batches = []
    for q in range(sets):
        batches.append(((2 * U(10 + 3 * q, 2) - 1).astype(np.float32),
                        (0.2 * U(11 + 3 * q, 3) - 0.1).astype(np.float32),
                        (0.2 * U(12 + 3 * q, 3)).astype(np.float32)))
    return st, batches
