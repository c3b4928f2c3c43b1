// TEST INFRASTRUCTURE ONLY.
// Host (g++) build of sac_rcbf_b200/csrc/rcbf_core.cuh so the numerics of the exact per-instance source the CUDA
// kernels run can be checked against the oracle on a CPU-only box (tests/test_hostsim.py).  It is NOT a CPU
// fallback: nothing in the product package loads this library, and the product fails loudly without the CUDA one.
#include <stdint.h>
#include "../../sac_rcbf_b200/csrc/rcbf_core.cuh"
#include "../../sac_rcbf_b200/csrc/rcbf_backward.cuh"

using namespace rcbf;

// The layer on an arbitrary hazard set (rcbf_general.cu composes the same templated pieces per thread): NH CBF rows,
// hazards beyond K padded 1e4 m away.  Forward + exact active-set gradient.
template <int NH>
static void general_one(const UnicycleParams& p, const float* hz_xy, int K, const float* st, const float* ac,
                        const float* mu, const float* sg, const float* gout, float* out, float* ga, int* status) {
  constexpr int M = NH + 4;
  using Pat = CbfPat<NH, 2>;
  float hz[NH][2];
  for (int k = 0; k < NH; ++k) {
    hz[k][0] = k < K ? hz_xy[2 * k] : 1.0e4f;
    hz[k][1] = k < K ? hz_xy[2 * k + 1] : 1.0e4f + 10.0f * (float)(k - K);
  }
  float sn, cs, Lg[NH][2], h[M], G[M][3];
  sincos_t(st[2], &sn, &cs);
  assemble_unicycle_n<float, NH>(p, hz, st, sn, cs, ac, mu, sg, Lg, h);
  for (int k = 0; k < NH; ++k) { G[k][0] = -Lg[k][0]; G[k][1] = -Lg[k][1]; G[k][2] = -1.0f; }
  for (int cc = 0; cc < 2; ++cc)
    for (int j = 0; j < 3; ++j) { G[NH + 2 * cc][j] = (j == cc) ? 1.0f : 0.0f; G[NH + 2 * cc + 1][j] = (j == cc) ? -1.0f : 0.0f; }
  bool triv, nan;
  classify_raw<M>(h, triv, nan);
  NormSolution<3, M> sol;
  sol.x[0] = sol.x[1] = sol.x[2] = 0.0;
  sol.status = nan ? RCBF_NAN : RCBF_OK_TRIVIAL;
  sol.mask = 0u;
  if (!triv && !nan) {
    solve_raw_fast<Pat, 3, M>(G, h, p.p_diag, false, sol);
    if (sol.status == RCBF_PENDING) {
      Normalised<3, M> nrm;
      normalise_rows<Pat, 3, M>(G, h, nrm);
      solve_normalised_full<Pat, 3, M>(nrm, p.p_diag, false, sol);
    }
  }
  *status = sol.status;
  for (int c = 0; c < 2; ++c) out[c] = clampf(ac[c] + (float)sol.x[c], p.u_min[c], p.u_max[c]);
  if (sol.status == RCBF_OK_TRIVIAL) {
    for (int c = 0; c < 2; ++c) ga[c] = (ac[c] >= p.u_min[c] && ac[c] <= p.u_max[c]) ? gout[c] : 0.f;
    return;
  }
  float r[M][2];
  for (int k = 0; k < NH; ++k) { r[k][0] = Lg[k][0]; r[k][1] = Lg[k][1]; }
  for (int c = 0; c < 2; ++c) {
    r[NH + 2 * c][0] = (c == 0) ? -1.f : 0.f; r[NH + 2 * c][1] = (c == 1) ? -1.f : 0.f;
    r[NH + 2 * c + 1][0] = (c == 0) ? 1.f : 0.f; r[NH + 2 * c + 1][1] = (c == 1) ? 1.f : 0.f;
  }
  double pisd[3]; float pisf[3];
  pis_of<3, M>(p.p_diag, pisd, pisf);
  safe_action_bwd_active<Pat, 3, M, 2>(G, h, r, pisd, sol.mask == kMaskUnknown ? 0u : sol.mask, ac, p.u_min, p.u_max, gout, ga);
}


extern "C" {

void hs_unicycle_safe_action(int64_t n, const float* st, const float* ac, const float* mu, const float* sg,
                             const UnicycleParams* p, float* out, double* x, double* lam, double* s, int* status,
                             int* iters, float* Gn, float* hn, float* G, float* h, int mode) {
#pragma omp parallel for schedule(static)
  for (int64_t i = 0; i < n; ++i) {
    UniSolve w;
    if (mode == 0) unicycle_safe_action<0>(*p, st + 3 * i, ac + 2 * i, mu + 3 * i, sg + 3 * i, w, out + 2 * i);
    else unicycle_safe_action<1>(*p, st + 3 * i, ac + 2 * i, mu + 3 * i, sg + 3 * i, w, out + 2 * i);
    if (w.sol.status == RCBF_PENDING) {  // what the fallback kernel does
      const int it0 = w.sol.iters;
      if (mode == 0) unicycle_safe_action<2>(*p, st + 3 * i, ac + 2 * i, mu + 3 * i, sg + 3 * i, w, out + 2 * i);
      else unicycle_safe_action<3>(*p, st + 3 * i, ac + 2 * i, mu + 3 * i, sg + 3 * i, w, out + 2 * i);
      w.sol.iters += 1000 + it0;  // +1000 flags the fallback pass
    }
    for (int j = 0; j < 3; ++j) x[3 * i + j] = w.sol.x[j];
    for (int r = 0; r < kUniM; ++r) {
      lam[kUniM * i + r] = w.sol.lam[r];
      s[kUniM * i + r] = w.sol.s[r];
      hn[kUniM * i + r] = w.nrm.hn[r];
      h[kUniM * i + r] = w.raw.h[r];
      for (int j = 0; j < 3; ++j) {
        Gn[(kUniM * i + r) * 3 + j] = w.nrm.Gn[r][j];
        G[(kUniM * i + r) * 3 + j] = w.raw.G[r][j];
      }
    }
    status[i] = w.sol.status;
    iters[i] = w.sol.iters;
  }
}

void hs_cars_safe_action(int64_t n, const float* st, const float* ac, const float* sg, const CarsParams* p, float* out,
                         double* x, double* lam, double* s, int* status, int* iters, float* Gn, float* hn, float* G,
                         float* h, int mode) {
#pragma omp parallel for schedule(static)
  for (int64_t i = 0; i < n; ++i) {
    CarsSolve w;
    if (mode == 0) cars_safe_action<0>(*p, st + 10 * i, ac[i], sg + 10 * i, w, out + i);
    else cars_safe_action<1>(*p, st + 10 * i, ac[i], sg + 10 * i, w, out + i);
    if (w.sol.status == RCBF_PENDING) {
      const int it0 = w.sol.iters;
      if (mode == 0) cars_safe_action<2>(*p, st + 10 * i, ac[i], sg + 10 * i, w, out + i);
      else cars_safe_action<3>(*p, st + 10 * i, ac[i], sg + 10 * i, w, out + i);
      w.sol.iters += 1000 + it0;
    }
    for (int j = 0; j < 2; ++j) x[2 * i + j] = w.sol.x[j];
    for (int r = 0; r < kCarsM; ++r) {
      lam[kCarsM * i + r] = w.sol.lam[r];
      s[kCarsM * i + r] = w.sol.s[r];
      hn[kCarsM * i + r] = w.nrm.hn[r];
      h[kCarsM * i + r] = w.raw.h[r];
      for (int j = 0; j < 2; ++j) {
        Gn[(kCarsM * i + r) * 2 + j] = w.nrm.Gn[r][j];
        G[(kCarsM * i + r) * 2 + j] = w.raw.G[r][j];
      }
    }
    status[i] = w.sol.status;
    iters[i] = w.sol.iters;
  }
}

// gradient of get_safe_action w.r.t. the nominal action: the qpth-clamp form on the dense saved tensors (ga_dense) and the
// exact active-set form on the certified mask alone (ga_active), both from the same forward solve
void hs_unicycle_bwd(int64_t n, const float* st, const float* ac, const float* mu, const float* sg, const float* gout,
                     const UnicycleParams* p, float* ga_dense, float* ga_active, int* status) {
#pragma omp parallel for schedule(static)
  for (int64_t i = 0; i < n; ++i) {
    UniSolve w;
    float out[2];
    unicycle_safe_action<0>(*p, st + 3 * i, ac + 2 * i, mu + 3 * i, sg + 3 * i, w, out);
    if (w.sol.status == RCBF_PENDING) unicycle_safe_action<2>(*p, st + 3 * i, ac + 2 * i, mu + 3 * i, sg + 3 * i, w, out);
    status[i] = (w.sol.status == RCBF_OK_CERTIFIED && w.sol.mask == kMaskUnknown) ? RCBF_OK_IPM : w.sol.status;
    float r[kUniM][2], xs[3], ls[kUniM], ss[kUniM];
    for (int k = 0; k < kUniHaz; ++k) { r[k][0] = w.raw.Lg[k][0]; r[k][1] = w.raw.Lg[k][1]; }
    for (int c = 0; c < 2; ++c) {
      r[kUniHaz + 2 * c][0] = (c == 0) ? -1.f : 0.f; r[kUniHaz + 2 * c][1] = (c == 1) ? -1.f : 0.f;
      r[kUniHaz + 2 * c + 1][0] = (c == 0) ? 1.f : 0.f; r[kUniHaz + 2 * c + 1][1] = (c == 1) ? 1.f : 0.f;
    }
    for (int j = 0; j < 3; ++j) xs[j] = (float)w.sol.x[j];
    for (int k = 0; k < kUniM; ++k) { ls[k] = (float)w.sol.lam[k]; ss[k] = (float)w.sol.s[k]; }
    safe_action_bwd<kUniNZ, kUniM, 2>(w.nrm, w.raw.G, w.raw.h, r, p->p_diag, xs, ls, ss, ac + 2 * i, p->u_min, p->u_max,
                                      gout + 2 * i, ga_dense + 2 * i);
    double pisd[3]; float pisf[3];
    pis_of<kUniNZ, kUniM>(p->p_diag, pisd, pisf);
    safe_action_bwd_active<UniPat, kUniNZ, kUniM, 2>(w.raw.G, w.raw.h, r, pisd, w.sol.mask == kMaskUnknown ? 0u : w.sol.mask,
                                                     ac + 2 * i, p->u_min, p->u_max, gout + 2 * i, ga_active + 2 * i);
  }
}

void hs_cars_bwd(int64_t n, const float* st, const float* ac, const float* sg, const float* gout, const CarsParams* p,
                 float* ga_dense, float* ga_active, int* status) {
#pragma omp parallel for schedule(static)
  for (int64_t i = 0; i < n; ++i) {
    CarsSolve w;
    float out;
    cars_safe_action<0>(*p, st + 10 * i, ac[i], sg + 10 * i, w, &out);
    if (w.sol.status == RCBF_PENDING) cars_safe_action<2>(*p, st + 10 * i, ac[i], sg + 10 * i, w, &out);
    status[i] = (w.sol.status == RCBF_OK_CERTIFIED && w.sol.mask == kMaskUnknown) ? RCBF_OK_IPM : w.sol.status;
    float r[kCarsM][1] = {{w.raw.Lg[0]}, {w.raw.Lg[1]}, {-1.f}, {1.f}};
    float xs[2], ls[kCarsM], ss[kCarsM];
    for (int j = 0; j < 2; ++j) xs[j] = (float)w.sol.x[j];
    for (int k = 0; k < kCarsM; ++k) { ls[k] = (float)w.sol.lam[k]; ss[k] = (float)w.sol.s[k]; }
    const float lo[1] = {p->u_min}, hi[1] = {p->u_max};
    safe_action_bwd<kCarsNZ, kCarsM, 1>(w.nrm, w.raw.G, w.raw.h, r, p->p_diag, xs, ls, ss, ac + i, lo, hi, gout + i,
                                        ga_dense + i);
    double pisd[2]; float pisf[2];
    pis_of<kCarsNZ, kCarsM>(p->p_diag, pisd, pisf);
    safe_action_bwd_active<CarsPat, kCarsNZ, kCarsM, 1>(w.raw.G, w.raw.h, r, pisd, w.sol.mask == kMaskUnknown ? 0u : w.sol.mask,
                                                        ac + i, lo, hi, gout + i, ga_active + i);
  }
}

void hs_unicycle_general(int64_t n, const float* st, const float* ac, const float* mu, const float* sg, const float* gout,
                         const UnicycleParams* p, const float* hz_xy, int K, float* out, float* ga, int* status) {
#pragma omp parallel for schedule(static)
  for (int64_t i = 0; i < n; ++i) {
    if (K <= 8) general_one<8>(*p, hz_xy, K, st + 3 * i, ac + 2 * i, mu + 3 * i, sg + 3 * i, gout + 2 * i, out + 2 * i, ga + 2 * i, status + i);
    else general_one<12>(*p, hz_xy, K, st + 3 * i, ac + 2 * i, mu + 3 * i, sg + 3 * i, gout + 2 * i, out + 2 * i, ga + 2 * i, status + i);
  }
}

int hs_sizeof_unicycle_params() { return (int)sizeof(UnicycleParams); }
int hs_sizeof_cars_params() { return (int)sizeof(CarsParams); }
}
