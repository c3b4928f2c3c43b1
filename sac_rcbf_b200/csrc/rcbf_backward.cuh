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

}  // namespace rcbf
