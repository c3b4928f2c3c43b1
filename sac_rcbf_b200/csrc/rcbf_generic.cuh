// rcbf_generic.cuh -- the generic small-QP entry (CBFQPLayer.cbf_layer / solve_qp API, rcbf_sac/diff_cbf_qp.py:81-144)
//
//     minimise 1/2 x'Qx + p'x   s.t.  Gx <= h        Q SPD (dense), float64 like the tensors qpth receives (:139)
//
// Forward: with Q = LL', y = L'x + L^-1 p turns the problem into the least-norm form of rcbf_core.cuh
// (A = G L^-T, b = h + A L^-1 p); duals and slacks are unchanged.  Backward: qpth's formulas on the dual Schur
// complement S = G Q^-1 G' + D^-1 (M x M, float64).  Not the hot path (the fused kernels are), kept simple.
#pragma once

#include "rcbf_core.cuh"

namespace rcbf {

template <int NZ, int M>
struct DirectCert {
  const LnpProblem<double, NZ, M>& P;
  RCBF_HD double a(int i, int j) const { return P.A[i][j]; }
  RCBF_HD double b(int i) const { return P.b[i]; }
};

template <int NZ, int M>
RCBF_HD void generic_qp_solve(const double* Q, const double* p, const double* G, const double* h, double x[NZ],
                              double lam[M], double s[M], int& status, int& iters) {
  double Qm[NZ][NZ];
  RCBF_UNROLL
  for (int j = 0; j < NZ; ++j) {
    RCBF_UNROLL
    for (int k = 0; k < NZ; ++k) Qm[j][k] = Q[j * NZ + k];
  }
  Chol<double, NZ> cq;
  cq.factor(Qm);
  double pv[NZ], v[NZ];
  RCBF_UNROLL
  for (int j = 0; j < NZ; ++j) pv[j] = p[j];
  cq.fwd(pv, v);  // v = L^-1 p
  LnpProblem<double, NZ, M> P;
  RCBF_UNROLL
  for (int i = 0; i < M; ++i) {
    double g[NZ];
    RCBF_UNROLL
    for (int j = 0; j < NZ; ++j) g[j] = G[i * NZ + j];
    cq.fwd(g, P.A[i]);  // row of G L^-T
    double acc = h[i];
    RCBF_UNROLL
    for (int j = 0; j < NZ; ++j) acc = fma(P.A[i][j], v[j], acc);
    P.b[i] = acc;
  }
  LnpSolution<double, NZ, M> sol;
  const DirectCert<NZ, M> cp{P};
  lnp_solve<double, double, DirectCert<NZ, M>, DensePat, NZ, M>(P, cp, sol, kTolSlack, kTolDual);
  double yv[NZ];
  RCBF_UNROLL
  for (int j = 0; j < NZ; ++j) yv[j] = sol.y[j] - v[j];
  cq.bwd(yv, x);  // x = L^-T (y - v)
  RCBF_UNROLL
  for (int i = 0; i < M; ++i) {
    lam[i] = sol.lam[i];
    s[i] = sol.s[i];
  }
  status = sol.status;
  iters = sol.iters;
}

template <int NZ, int M>
RCBF_HD void generic_qp_bwd(const double* Q, const double* G, const double* x, const double* lam, const double* slack,
                            const double* gx, double* dQ, double* dp, double* dG, double* dh) {
  double Qm[NZ][NZ];
  RCBF_UNROLL
  for (int j = 0; j < NZ; ++j) {
    RCBF_UNROLL
    for (int k = 0; k < NZ; ++k) Qm[j][k] = Q[j * NZ + k];
  }
  Chol<double, NZ> cq;
  cq.factor(Qm);
  double T[M][NZ], S[M][M], hh[M], g[NZ], t[NZ];
  RCBF_UNROLL
  for (int j = 0; j < NZ; ++j) g[j] = gx[j];
  cq.solve(g, t);  // Q^-1 dl/dx
  RCBF_UNROLL
  for (int i = 0; i < M; ++i) {
    double gi[NZ];
    RCBF_UNROLL
    for (int j = 0; j < NZ; ++j) gi[j] = G[i * NZ + j];
    cq.solve(gi, T[i]);
  }
  RCBF_UNROLL
  for (int i = 0; i < M; ++i) {
    const double d = fmax(lam[i], 1e-8) / fmax(slack[i], 1e-8);
    RCBF_UNROLL
    for (int k = 0; k <= i; ++k) {
      double acc = (i == k) ? 1.0 / d : 0.0;
      RCBF_UNROLL
      for (int j = 0; j < NZ; ++j) acc = fma(G[i * NZ + j], T[k][j], acc);
      S[i][k] = acc;
    }
    double acc = 0.0;
    RCBF_UNROLL
    for (int j = 0; j < NZ; ++j) acc = fma(G[i * NZ + j], t[j], acc);
    hh[i] = acc;
  }
  // in-place Cholesky of S (lower), then w = -S^-1 hh
  for (int i = 0; i < M; ++i) {
    for (int k = 0; k <= i; ++k) {
      double acc = S[i][k];
      for (int l = 0; l < k; ++l) acc -= S[i][l] * S[k][l];
      S[i][k] = (i == k) ? sqrt(acc) : acc / S[k][k];
    }
  }
  double w[M];
  for (int i = 0; i < M; ++i) {
    double acc = -hh[i];
    for (int l = 0; l < i; ++l) acc -= S[i][l] * w[l];
    w[i] = acc / S[i][i];
  }
  for (int i = M - 1; i >= 0; --i) {
    double acc = w[i];
    for (int l = i + 1; l < M; ++l) acc -= S[l][i] * w[l];
    w[i] = acc / S[i][i];
  }
  double r[NZ], dx[NZ];
  RCBF_UNROLL
  for (int j = 0; j < NZ; ++j) {
    double acc = -g[j];
    RCBF_UNROLL
    for (int i = 0; i < M; ++i) acc = fma(-G[i * NZ + j], w[i], acc);
    r[j] = acc;
  }
  cq.solve(r, dx);
  RCBF_UNROLL
  for (int j = 0; j < NZ; ++j) {
    dp[j] = dx[j];
    RCBF_UNROLL
    for (int k = 0; k < NZ; ++k) dQ[j * NZ + k] = 0.5 * (dx[j] * x[k] + x[j] * dx[k]);
  }
  RCBF_UNROLL
  for (int i = 0; i < M; ++i) {
    dh[i] = -w[i];
    RCBF_UNROLL
    for (int j = 0; j < NZ; ++j) dG[i * NZ + j] = w[i] * x[j] + lam[i] * dx[j];
  }
}

}  // namespace rcbf
