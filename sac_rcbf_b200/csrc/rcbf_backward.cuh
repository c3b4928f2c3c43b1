// rcbf_backward.cuh -- implicit-KKT gradient of get_safe_action w.r.t. the nominal action.
//
// Reference semantics (what autograd + qpth compute):
//   final = clamp(a + x*[:n_u], u_min, u_max)                                   rcbf_sac/diff_cbf_qp.py:77
//   x*    = argmin 1/2 x'Px  s.t.  G~x <= h~ ,  [G~|h~] = [G|h]/n,  n_i = max_j(|G_ij|,|h_i|)     :103-107
//   h_i   = c_i + r_i'a   (r_i = Lg_i on CBF rows :261/:348-349, -/+e_c on the actuator rows :370/:376);  G is
//           independent of a.  Only `a` carries grad into the layer (sac_cbf.py:233-236 cut the graph elsewhere).
//   qpth backward: d = clamp(lam,1e-8)/clamp(s,1e-8);  K [dx;.;dlam] = -[dl/dx;0;0];
//                  dl/dh~ = -dlam,  dl/dG~ = dlam x' + lam dx'.
//
// The KKT system is solved in float64 by a mixed elimination that is stable for the extreme d this produces
// (1e-8/s on inactive rows, lam/1e-8 on active rows): inactive rows are folded into the primal block
// Q' = P + G_I' D_I G_I (well conditioned), the <= NZ active rows go through the dual Schur complement
// S = G_A Q'^-1 G_A' + D_A^-1.  This is algebraically the same K as qpth's.
#pragma once

#include "rcbf_core.cuh"

namespace rcbf {

// r[i][c] = d h_i / d a_c ; NU = number of controls ; rows [NCBF, M) are the actuator rows (+e_c, -e_c per control)
template <int NZ, int M, int NU>
RCBF_HD void safe_action_bwd(const Normalised<NZ, M>& nrm, const float rawG[M][NZ], const float rawh[M],
                             const float r[M][NU], const float p_diag[NZ], const float xs[NZ], const float lams[M],
                             const float slacks[M], const float a[NU], const float u_min[NU], const float u_max[NU],
                             const float gout[NU], float grad_a[NU]) {
  // clamp mask (torch.clamp passes grad where min <= v <= max)
  double g[NZ];
  RCBF_UNROLL
  for (int j = 0; j < NZ; ++j) g[j] = 0.0;
  RCBF_UNROLL
  for (int c = 0; c < NU; ++c) {
    const float v = a[c] + xs[c];
    g[c] = (v >= u_min[c] && v <= u_max[c]) ? (double)gout[c] : 0.0;
  }
  double d[M];
  uint32_t act = 0;
  int nact = 0;
  RCBF_UNROLL
  for (int i = 0; i < M; ++i) {
    const double l = fmax((double)lams[i], 1e-8), s = fmax((double)slacks[i], 1e-8);
    d[i] = l / s;
    const bool is_act = (d[i] > 1.0) && (nact < NZ);
    act |= is_act ? (1u << i) : 0u;
    nact += is_act ? 1 : 0;
  }
  // Q' = P + sum_{i not active} d_i g_i g_i'
  double Qp[NZ][NZ];
  RCBF_UNROLL
  for (int j = 0; j < NZ; ++j) {
    RCBF_UNROLL
    for (int k = 0; k <= j; ++k) {
      double acc = (j == k) ? (double)p_diag[j] : 0.0;
      RCBF_UNROLL
      for (int i = 0; i < M; ++i) {
        const double w = ((act >> i) & 1u) ? 0.0 : d[i];
        acc = fma((double)nrm.Gn[i][j] * w, (double)nrm.Gn[i][k], acc);
      }
      Qp[j][k] = acc;
    }
  }
  Chol<double, NZ> cq;
  cq.factor(Qp);
  // gather active rows
  double R[NZ][NZ], dinv[NZ];
  RCBF_UNROLL
  for (int k = 0; k < NZ; ++k) {
    RCBF_UNROLL
    for (int j = 0; j < NZ; ++j) R[k][j] = 0.0;
    dinv[k] = 1.0;
  }
  int cnt = 0;
  RCBF_UNROLL
  for (int i = 0; i < M; ++i) {
    const bool is_act = (act >> i) & 1u;
    RCBF_UNROLL
    for (int k = 0; k < NZ; ++k) {
      const bool put = is_act && (cnt == k);
      RCBF_UNROLL
      for (int j = 0; j < NZ; ++j) R[k][j] = put ? (double)nrm.Gn[i][j] : R[k][j];
      dinv[k] = put ? 1.0 / d[i] : dinv[k];
    }
    cnt += is_act ? 1 : 0;
  }
  // W_k = Q'^-1 R_k' ; S = R W + diag(dinv) ; t = Q'^-1 g
  double W[NZ][NZ], t[NZ];
  RCBF_UNROLL
  for (int k = 0; k < NZ; ++k) cq.solve(R[k], W[k]);
  cq.solve(g, t);
  double S[NZ][NZ], hh[NZ], wv[NZ];
  RCBF_UNROLL
  for (int k = 0; k < NZ; ++k) {
    RCBF_UNROLL
    for (int l = 0; l <= k; ++l) {
      double acc = (k == l) ? dinv[k] : 0.0;
      RCBF_UNROLL
      for (int j = 0; j < NZ; ++j) acc = fma(R[k][j], W[l][j], acc);
      S[k][l] = acc;
    }
    double acc = 0.0;
    RCBF_UNROLL
    for (int j = 0; j < NZ; ++j) acc = fma(R[k][j], t[j], acc);
    hh[k] = acc;  // G_A Q'^-1 rx  (rx = dl/dx)
  }
  Chol<double, NZ> cs;
  cs.factor(S);
  cs.solve(hh, wv);  // w_A = -S^-1 hh  -> keep +S^-1 hh and flip signs below
  // dx = Q'^-1 (-rx - G_A' w_A) = -t + W' (S^-1 hh)
  double dx[NZ];
  RCBF_UNROLL
  for (int j = 0; j < NZ; ++j) {
    double acc = -t[j];
    RCBF_UNROLL
    for (int k = 0; k < NZ; ++k) acc = fma(W[k][j], wv[k], acc);
    dx[j] = acc;
  }
  // dlam: active rows = w_A = -wv ; inactive rows = d_i g_i'dx.   Then chain through the row normalisation.
  int c2 = 0;
  double ga[NU];
  RCBF_UNROLL
  for (int c = 0; c < NU; ++c) ga[c] = g[c];  // identity path  a -> a + x
  RCBF_UNROLL
  for (int i = 0; i < M; ++i) {
    const bool is_act = (act >> i) & 1u;
    double gdx = 0.0, gx = 0.0;
    RCBF_UNROLL
    for (int j = 0; j < NZ; ++j) {
      gdx = fma((double)nrm.Gn[i][j], dx[j], gdx);
      gx = fma((double)nrm.Gn[i][j], (double)xs[j], gx);
    }
    double dl = d[i] * gdx;
    RCBF_UNROLL
    for (int k = 0; k < NZ; ++k) dl = (is_act && c2 == k) ? -wv[k] : dl;
    c2 += is_act ? 1 : 0;
    const double dht = -dl;  // dl/dh~_i
    // sum_j dl/dG~_ij G~_ij = dlam_i (g_i'x) + lam_i (g_i'dx)
    const double dGG = dl * gx + (double)lams[i] * gdx;
    const double n = (double)nrm.n[i];
    double dh = dht / n;  // direct path h~ = h/n
    if ((nrm.h_is_max >> i) & 1u) {
      const double dn = -(dGG + dht * (double)nrm.hn[i]) / n;  // dL/dn_i
      dh += dn * (rawh[i] >= 0.0f ? 1.0 : -1.0);
    }
    RCBF_UNROLL
    for (int c = 0; c < NU; ++c) ga[c] = fma(dh, (double)r[i][c], ga[c]);
  }
  RCBF_UNROLL
  for (int c = 0; c < NU; ++c) grad_a[c] = (float)ga[c];
}

// ---- exact active-set form (what the compact backward kernels run) ---------------------------------------------------
// The forward certificate established WHICH rows S (|S| <= NZ) are active at the optimum, so the solution is the closed
// form  y = A_S' Gm^-1 b_S  (A = G~ P^-1/2, b = h~, Gm = A_S A_S', x = P^-1/2 y) and its derivative is that of the
// equality-constrained problem -- the limit of qpth's backward above as its clamps 1e-8 -> 0 (d -> inf on the active rows,
// d -> 0 on the inactive ones; the terms dropped are O(1e-8 / slack_i) and O(1e-8 / lam_k), below the convergence noise
// of qpth's own iterate):
//     dL/db_S = v = Gm^-1 A_S gy          (gy = P^-1/2 dL/dx),      dL/dA_k = mu_k (gy - A_S'v) - v_k y,   mu = Gm^-1 b_S.
// Chain through the row normalisation (diff_cbf_qp.py:103-106): b_k = h_k / n_k, A_k = G_k P^-1/2 / n_k with G independent
// of the action.  If n_k does not depend on h_k:  dL/dh_k = v_k / n_k.  If n_k = |h_k| (h is the row maximum): b_k = +-1
// is constant and A_k = G_k P^-1/2 / |h_k|, so dL/dh_k = -(sgn h_k / n_k) dL/dA_k . A_k = (sgn h_k / n_k) v_k b_k
// (A_k is orthogonal to gy - A_S'v and A_k . y = b_k) = v_k / n_k again.  Hence, for every active row,
//     dL/da_c = g_c + sum_{k in S} (v_k / n_k) r_k[c],                                  inactive rows contribute nothing:
// only the <= NZ active rows are gathered and normalised (the same float32 quotients the forward certified), one
// NZ x NZ float64 Cholesky serves both the re-solve for x (clamp mask of diff_cbf_qp.py:77) and the gradient.
// pis = P^-1/2 in float64 (pis_of).
template <typename Pat, int NZ, int M, int NU>
RCBF_HD void safe_action_bwd_active(const float G[M][NZ], const float h[M], const float r[M][NU], const double pis[NZ],
                                    uint32_t mask, const float a[NU], const float u_min[NU], const float u_max[NU],
                                    const float gout[NU], float grad_a[NU]) {
  float g[NZ][NZ], hh[NZ], rr[NZ][NU];
  RCBF_UNROLL
  for (int k = 0; k < NZ; ++k) {
    RCBF_UNROLL
    for (int j = 0; j < NZ; ++j) g[k][j] = 0.f;
    RCBF_UNROLL
    for (int c = 0; c < NU; ++c) rr[k][c] = 0.f;
    hh[k] = 0.f;
  }
  int cnt = 0;
  RCBF_UNROLL
  for (int i = 0; i < M; ++i) {
    const bool act = (mask >> i) & 1u;
    RCBF_UNROLL
    for (int k = 0; k < NZ; ++k) {
      const bool put = act && (cnt == k);
      RCBF_UNROLL
      for (int j = 0; j < NZ; ++j)
        if (Pat::nz(i, j)) g[k][j] = put ? G[i][j] : g[k][j];
      RCBF_UNROLL
      for (int c = 0; c < NU; ++c) rr[k][c] = put ? r[i][c] : rr[k][c];
      hh[k] = put ? h[i] : hh[k];
    }
    cnt += act ? 1 : 0;
  }
  double R[NZ][NZ], rb[NZ], ninv[NZ];
  RCBF_UNROLL
  for (int k = 0; k < NZ; ++k) {  // normalise the gathered row exactly like the forward (lnp_greedy_raw / normalise_rows)
    float gm = 0.f;
    RCBF_UNROLL
    for (int j = 0; j < NZ; ++j) gm = fmaxf(gm, fabsf(g[k][j]));
    const float ha = fabsf(hh[k]);
    const float n = (k < cnt) ? ((ha != ha) ? ha : fmaxf(gm, ha)) : 1.f;
    const float rn = rcp_refined(n);
    RCBF_UNROLL
    for (int j = 0; j < NZ; ++j) R[k][j] = (double)div_by(g[k][j], n, rn) * pis[j];
    rb[k] = (double)div_by(hh[k], n, rn);
    ninv[k] = (double)rn;
  }
  double Gm[NZ][NZ];
  RCBF_UNROLL
  for (int k = 0; k < NZ; ++k) {
    RCBF_UNROLL
    for (int l = 0; l <= k; ++l) {
      double acc = 0.0;
      RCBF_UNROLL
      for (int j = 0; j < NZ; ++j) acc = fma(R[k][j], R[l][j], acc);
      Gm[k][l] = acc;
    }
    Gm[k][k] = (k < cnt) ? Gm[k][k] : 1.0;
  }
  Chol<double, NZ> ch;
  ch.factor(Gm);
  double mu[NZ];
  ch.solve(rb, mu);
  // x_c = P_c^-1/2 sum_k A_kc mu_k ; clamp mask (torch.clamp passes grad where min <= v <= max)
  double gy[NU], gc[NU];
  RCBF_UNROLL
  for (int c = 0; c < NU; ++c) {
    double y = 0.0;
    RCBF_UNROLL
    for (int k = 0; k < NZ; ++k) y = fma(R[k][c], mu[k], y);
    const float v = a[c] + (float)(y * pis[c]);
    gc[c] = (v >= u_min[c] && v <= u_max[c]) ? (double)gout[c] : 0.0;
    gy[c] = gc[c] * pis[c];
  }
  double w[NZ], v[NZ];
  RCBF_UNROLL
  for (int k = 0; k < NZ; ++k) {
    double acc = 0.0;
    RCBF_UNROLL
    for (int c = 0; c < NU; ++c) acc = fma(R[k][c], gy[c], acc);
    w[k] = acc;
  }
  ch.solve(w, v);
  RCBF_UNROLL
  for (int c = 0; c < NU; ++c) {
    double acc = gc[c];
    RCBF_UNROLL
    for (int k = 0; k < NZ; ++k) acc = fma(v[k] * ninv[k], (double)rr[k][c], acc);
    grad_a[c] = (float)acc;
  }
}

}  // namespace rcbf
