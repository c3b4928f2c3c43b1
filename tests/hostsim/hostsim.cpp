// TEST INFRASTRUCTURE ONLY.
// Host (g++) build of sac_rcbf_b200/csrc/rcbf_core.cuh so the numerics of the exact per-instance source the CUDA
// kernels run can be checked against the oracle on a CPU-only box (tests/test_hostsim.py).  It is NOT a CPU
// fallback: nothing in the product package loads this library, and the product fails loudly without the CUDA one.
#include <stdint.h>
#include "../../sac_rcbf_b200/csrc/rcbf_core.cuh"

using namespace rcbf;

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

int hs_sizeof_unicycle_params() { return (int)sizeof(UnicycleParams); }
int hs_sizeof_cars_params() { return (int)sizeof(CarsParams); }
}
