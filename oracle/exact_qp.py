"""TEST INFRASTRUCTURE ONLY -- exact solver for the tiny strictly convex QPs of the path.

    minimise 1/2 x'Qx + p'x   s.t.  Gx <= h       (nz <= 3 variables, m <= 9 rows)

Ground truth that is independent of any interior-point method: enumerate every
active set A with |A| <= nz (1+9+36+84 = 130 subsets for m=9, nz=3), solve the
equality-constrained problem in float64

    (G_A Q^-1 G_A') lam = -(h_A + G_A Q^-1 p),   x = -Q^-1 (p + G_A' lam)

and accept the subset whose (x, lam) satisfies the KKT conditions (lam >= 0,
Gx <= h); for a strictly convex QP the KKT point is the unique optimum.  Used to
pin oracle/qpth_pdipm.py (the reference's solver, qpth, is absent from the image;
see that file's header) and as the "vs exact" leg of the parity tests.

Also provides ``quadprog_solve_qp`` with quadprog's call shape, used only by
oracle/ref_loader.py so that the reference's CascadeCBFLayer (rcbf_sac/cbf_qp.py:276,
``solve_qp(P, q, -G.T, -h)``) can run from its own source.
"""
from itertools import combinations

import numpy as np


def solve_exact(Q, p, G, h, tol=1e-9):
    """Batched exact solve.  Q (B,nz,nz), p (B,nz), G (B,m,nz), h (B,m) float64.

    Returns x (B,nz), lam (B,m), active (B,m) bool, viol (B,) = max KKT violation of the
    accepted subset (should be ~1e-12; large values flag a degenerate/ill-posed instance).
    """
    Q = np.asarray(Q, np.float64)
    p = np.asarray(p, np.float64)
    G = np.asarray(G, np.float64)
    h = np.asarray(h, np.float64)
    B, m, nz = G.shape
    Qinv = np.linalg.inv(Q)
    best_x = np.zeros((B, nz))
    best_lam = np.zeros((B, m))
    best_act = np.zeros((B, m), bool)
    best_v = np.full((B,), np.inf)

    x0 = -np.einsum("bij,bj->bi", Qinv, p)
    for k in range(0, nz + 1):
        for A in combinations(range(m), k):
            A = list(A)
            if k == 0:
                x = x0
                lamA = np.zeros((B, 0))
                ok = np.ones(B, bool)
            else:
                GA = G[:, A, :]                                   # (B,k,nz)
                GQ = np.einsum("bki,bij->bkj", GA, Qinv)          # (B,k,nz)
                M = np.einsum("bkj,blj->bkl", GQ, GA)             # (B,k,k)
                rhs = -(h[:, A] + np.einsum("bkj,bj->bk", GQ, p))
                det = np.linalg.det(M)
                scale = np.prod(np.maximum(np.einsum("bkk->bk", M), 1e-300), axis=1)
                ok = np.abs(det) > 1e-12 * scale                  # skip (near-)dependent rows
                Ms = np.where(ok[:, None, None], M, np.eye(k)[None])
                lamA = np.linalg.solve(Ms, rhs[..., None])[..., 0]
                x = x0 - np.einsum("bkj,bk->bj", GQ, lamA)
            slack = h - np.einsum("bmj,bj->bm", G, x)
            v = np.maximum(0.0, -slack.min(axis=1))
            if k > 0:
                v = np.maximum(v, np.maximum(0.0, -lamA.min(axis=1)))
            v = np.where(ok, v, np.inf)
            better = v < best_v - 1e-15
            if better.any():
                best_v = np.where(better, v, best_v)
                best_x = np.where(better[:, None], x, best_x)
                lam_full = np.zeros((B, m))
                if k > 0:
                    lam_full[:, A] = lamA
                best_lam = np.where(better[:, None], lam_full, best_lam)
                act = np.zeros((B, m), bool)
                act[:, A] = True
                best_act = np.where(better[:, None], act, best_act)
    return best_x, best_lam, best_act, best_v


def quadprog_solve_qp(Gq, a, C=None, b=None, meq=0):
    """quadprog.solve_qp call shape: min 1/2 x'Gq x - a'x  s.t.  C'x >= b."""
    if meq != 0:
        raise NotImplementedError
    Gq = np.asarray(Gq, np.float64)
    a = np.asarray(a, np.float64)
    n = a.shape[0]
    if C is None:
        G = np.zeros((1, 0, n))
        h = np.zeros((1, 0))
    else:
        G = -np.asarray(C, np.float64).T[None]
        h = -np.asarray(b, np.float64)[None]
    x, lam, act, v = solve_exact(Gq[None], -a[None], G, h)
    if not np.isfinite(v[0]) or v[0] > 1e-6:
        raise ValueError("constraints are inconsistent, no solution")
    f = 0.5 * x[0] @ Gq @ x[0] - a @ x[0]
    xu = np.linalg.solve(Gq, a)
    return x[0], f, xu, np.array([0, 0]), lam[0], np.nonzero(act[0])[0] + 1
