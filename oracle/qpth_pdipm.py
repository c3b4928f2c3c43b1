"""TEST INFRASTRUCTURE ONLY -- CPU restatement of the QP solver the reference calls.

PARITY STATUS: **unpinned by the reference.**  The reference solves its QPs with
the third-party package ``qpth`` (``from qpth.qp import QPFunction``,
rcbf_sac/diff_cbf_qp.py:7, called at :139 with
``QPFunction(verbose=0, check_Q_spd=False, maxIter=100000, notImprovedLim=10,
eps=1e-4)``, :107).  qpth is not vendored in /root/reference, is not pinned by
any requirements/lock file there (none exists), and is not installed in this
image, and the reference ships no tests or golden vectors for the solve.  This
file therefore restates qpth's *published* algorithm (Amos & Kolter, "OptNet",
ICML 2017, sec. 3 + appendix: batched primal-dual interior point with Mehrotra
predictor-corrector, dual Schur-complement LU; the call signature used by the
reference implies the factory API of qpth >= 0.0.15) and is pinned instead
against an independent mathematical ground truth, oracle/exact_qp.py (active-set
enumeration + KKT check), in tests/test_oracle.py.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl
reference legs may import this module.  The product package never does.

Algorithm (inequality-only, neq = 0; the reference always passes empty A, b,
diff_cbf_qp.py:135-137):

    minimise 1/2 x'Qx + p'x   s.t.  Gx <= h          (x in R^nz, m rows)

  init      solve K(D=I) [x;s;z] = -[p;0;-h];  if min(s)<0: s -= min(s)-1; same for z
  loop      rx = G'z + Qx + p ; rs = z ; rz = Gx + s - h ; mu = |s'z|/m
            resid = |rz|_2 + |rx|_2 + m*mu ; keep the per-QP best iterate by resid
            stop when   no QP of the batch improved for notImprovedLim iterations
                     or max_batch(best resid) < eps   or   min(mu) > 1e32
            d = z/s ; factor S = G Q^-1 G' + diag(1/d)           (batched LU, m x m)
            affine    K [dx;ds;dz] = -[rx;rs;rz]
            alpha   = min(1, step(z,dz_aff), step(s,ds_aff))
            sigma   = ((s+alpha ds_aff)'(z+alpha dz_aff) / s'z)^3
            corrector K [dx;ds;dz] = -[0; (-mu sigma + ds_aff*dz_aff)/s; 0]
            alpha   = min(1, 0.999 min(step(z,dz), step(s,ds)))   ; x,s,z += alpha * (aff+cor)
  where K = [[Q,0,G'],[0,D,I],[G,I,0]], D = diag(d) and
  step(v,dv) = min_i(-v_i/dv_i) with entries having dv_i>0 replaced by
  max(1, max over the whole batch of -v/dv)  (batch-global, harmless after min(.,1)).

  backward  d = clamp(lam,1e-8)/clamp(s,1e-8); solve K [dx;.;dlam] = -[dl/dx;0;0];
            dl/dp = dx ; dl/dh = -dlam ; dl/dG = dlam x' + lam dx' ; dl/dQ = 1/2 (dx x' + x dx')
"""
import torch


def _lu(M):
    return torch.linalg.lu_factor(M)


def _lu_solve(LU, rhs):
    # rhs (B, n) -> (B, n)
    return torch.linalg.lu_solve(LU[0], LU[1], rhs.unsqueeze(-1)).squeeze(-1)


def _bmv(M, v):
    return torch.bmm(M, v.unsqueeze(-1)).squeeze(-1)


def _factor_kkt(R, d):
    """S = R + diag(1/d), R = G Q^-1 G' (neq = 0)."""
    S = R + torch.diag_embed(1.0 / d)
    return _lu(S)


def _solve_kkt(Q_LU, d, G, S_LU, rx, rs, rz):
    """Solve K [dx;ds;dz] = -[rx;rs;rz] by block elimination on the dual Schur complement."""
    invQ_rx = _lu_solve(Q_LU, rx)
    hh = _bmv(G, invQ_rx) + rs / d - rz
    w = -_lu_solve(S_LU, hh)
    g1 = -rx - _bmv(G.transpose(1, 2), w)
    g2 = -rs - w
    dx = _lu_solve(Q_LU, g1)
    ds = g2 / d
    dz = w
    return dx, ds, dz


def _get_step(v, dv):
    a = -v / dv
    amax = a.max()
    # python `max(1.0, a.max())` semantics: a NaN maximum compares False and yields 1.0
    amax = amax if bool(amax > 1.0) else torch.tensor(1.0, dtype=a.dtype)
    a = torch.where(dv > 0, amax.expand_as(a), a)
    return a.min(dim=1)[0]


def pdipm_forward(Q, p, G, h, eps=1e-12, notImprovedLim=3, maxIter=20, per_qp_stop=False):
    """Batched PDIPM.  Q (B,nz,nz), p (B,nz), G (B,m,nz), h (B,m), all float64.

    Returns x, z(lams), s(slacks), info dict {iters, resid (B,), iter_hist (B,) first
    iteration index at which each QP's best resid dropped below eps}.
    ``per_qp_stop`` is NOT qpth behaviour; it is used only by the iteration-count probes.
    """
    B, m, nz = G.shape
    Q_LU = _lu(Q)
    R = torch.bmm(G, torch.linalg.lu_solve(Q_LU[0], Q_LU[1], G.transpose(1, 2)))

    d = torch.ones(B, m, dtype=Q.dtype)
    S_LU = _factor_kkt(R, d)
    x, s, z = _solve_kkt(Q_LU, d, G, S_LU, p, torch.zeros(B, m, dtype=Q.dtype), -h)

    Mn = s.min(dim=1, keepdim=True)[0]
    s = torch.where(Mn < 0, s - (Mn - 1), s)
    Mn = z.min(dim=1, keepdim=True)[0]
    z = torch.where(Mn < 0, z - (Mn - 1), z)

    best = None
    nNotImproved = 0
    first_below = torch.full((B,), -1, dtype=torch.long)
    iters = 0
    for i in range(maxIter):
        iters = i
        rx = _bmv(G.transpose(1, 2), z) + _bmv(Q, x) + p
        rs = z
        rz = _bmv(G, x) + s - h
        mu = torch.abs((s * z).sum(1) / m)
        resids = torch.linalg.norm(rz, dim=1) + torch.linalg.norm(rx, dim=1) + m * mu
        d = z / s
        try:
            S_LU = _factor_kkt(R, d)
        except RuntimeError:
            break
        if best is None:
            best = {"resids": resids.clone(), "x": x.clone(), "z": z.clone(), "s": s.clone()}
            nNotImproved = 0
        else:
            I = resids < best["resids"]
            if I.sum() > 0:
                nNotImproved = 0
            else:
                nNotImproved += 1
            best["resids"] = torch.where(I, resids, best["resids"])
            Ic = I.unsqueeze(1)
            best["x"] = torch.where(Ic, x, best["x"])
            best["z"] = torch.where(Ic, z, best["z"])
            best["s"] = torch.where(Ic, s, best["s"])
        newly = (best["resids"] < eps) & (first_below < 0)
        first_below[newly] = i
        if nNotImproved == notImprovedLim or best["resids"].max() < eps or mu.min() > 1e32:
            break

        dx_aff, ds_aff, dz_aff = _solve_kkt(Q_LU, d, G, S_LU, rx, rs, rz)
        alpha = torch.clamp(torch.minimum(_get_step(z, dz_aff), _get_step(s, ds_aff)), max=1.0)
        a = alpha.unsqueeze(1)
        t3 = ((s + a * ds_aff) * (z + a * dz_aff)).sum(1)
        t4 = (s * z).sum(1)
        sig = (t3 / t4) ** 3
        rs_c = ((-mu * sig).unsqueeze(1) + ds_aff * dz_aff) / s
        zx = torch.zeros_like(rx)
        zz = torch.zeros_like(rz)
        dx_cor, ds_cor, dz_cor = _solve_kkt(Q_LU, d, G, S_LU, zx, rs_c, zz)
        dx = dx_aff + dx_cor
        ds = ds_aff + ds_cor
        dz = dz_aff + dz_cor
        alpha = torch.clamp(0.999 * torch.minimum(_get_step(z, dz), _get_step(s, ds)), max=1.0)
        a = alpha.unsqueeze(1)
        x = x + a * dx
        s = s + a * ds
        z = z + a * dz

    info = {"iters": iters, "resid": best["resids"], "first_below_eps": first_below}
    return best["x"], best["z"], best["s"], info


def pdipm_backward(Q, G, x, lams, slacks, dl_dx):
    """Implicit-KKT gradient, qpth style (clamped d).  Returns dQ, dp, dG, dh."""
    B, m, nz = G.shape
    d = torch.clamp(lams, min=1e-8) / torch.clamp(slacks, min=1e-8)
    Q_LU = _lu(Q)
    R = torch.bmm(G, torch.linalg.lu_solve(Q_LU[0], Q_LU[1], G.transpose(1, 2)))
    S_LU = _factor_kkt(R, d)
    zero_m = torch.zeros(B, m, dtype=Q.dtype)
    dx, _, dlam = _solve_kkt(Q_LU, d, G, S_LU, dl_dx, zero_m, zero_m)
    dps = dx
    dGs = torch.einsum("bi,bj->bij", dlam, x) + torch.einsum("bi,bj->bij", lams, dx)
    dhs = -dlam
    dQs = 0.5 * (torch.einsum("bi,bj->bij", dx, x) + torch.einsum("bi,bj->bij", x, dx))
    return dQs, dps, dGs, dhs


class _QPFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, Q, p, G, h, A, b, eps, notImprovedLim, maxIter):
        if A.numel() > 0 or b.numel() > 0:
            raise NotImplementedError("oracle restates the inequality-only path (neq=0) the reference uses")
        x, lams, slacks, info = pdipm_forward(Q, p, G, h, eps=eps, notImprovedLim=notImprovedLim, maxIter=maxIter)
        ctx.save_for_backward(x, Q, G, lams, slacks)
        ctx.info = info
        _QPFn.last_info = info
        return x

    @staticmethod
    def backward(ctx, dl_dx):
        x, Q, G, lams, slacks = ctx.saved_tensors
        dQ, dp, dG, dh = pdipm_backward(Q, G, x, lams, slacks, dl_dx)
        return dQ, dp, dG, dh, None, None, None, None, None


def QPFunction(eps=1e-12, verbose=0, notImprovedLim=3, maxIter=20, solver=None, check_Q_spd=True):
    """Factory with qpth's call shape: ``QPFunction(**kw)(Q, p, G, h, A, b)``."""

    def apply(Q, p, G, h, A, b):
        return _QPFn.apply(Q, p, G, h, A, b, eps, notImprovedLim, maxIter)

    return apply
