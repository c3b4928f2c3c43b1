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
                             int* iters, float* Gn, float* hn, float* G, float* h) {
#pragma omp parallel for schedule(static)
  for (int64_t i = 0; i < n; ++i) {
    UniSolve w;
    unicycle_safe_action(*p, st + 3 * i, ac + 2 * i, mu + 3 * i, sg + 3 * i, w, out + 2 * i);
    for (int j = 0; j < 3; ++j) x[3 * i + j] = w.x[j];
    for (int r = 0; r < kUniM; ++r) {
      lam[kUniM * i + r] = w.lam[r];
      s[kUniM * i + r] = w.s[r];
      hn[kUniM * i + r] = w.nrm.hn[r];
      h[kUniM * i + r] = w.raw.h[r];
      for (int j = 0; j < 3; ++j) {
        Gn[(kUniM * i + r) * 3 + j] = w.nrm.Gn[r][j];
        G[(kUniM * i + r) * 3 + j] = w.raw.G[r][j];
      }
    }
    status[i] = w.status;
    iters[i] = w.iters;
  }
}

void hs_cars_safe_action(int64_t n, const float* st, const float* ac, const float* sg, const CarsParams* p, float* out,
                         double* x, double* lam, double* s, int* status, int* iters, float* Gn, float* hn, float* G,
                         float* h) {
#pragma omp parallel for schedule(static)
  for (int64_t i = 0; i < n; ++i) {
    CarsSolve w;
    cars_safe_action(*p, st + 10 * i, ac[i], sg + 10 * i, w, out + i);
    for (int j = 0; j < 2; ++j) x[2 * i + j] = w.x[j];
    for (int r = 0; r < kCarsM; ++r) {
      lam[kCarsM * i + r] = w.lam[r];
      s[kCarsM * i + r] = w.s[r];
      hn[kCarsM * i + r] = w.nrm.hn[r];
      h[kCarsM * i + r] = w.raw.h[r];
      for (int j = 0; j < 2; ++j) {
        Gn[(kCarsM * i + r) * 2 + j] = w.nrm.Gn[r][j];
        G[(kCarsM * i + r) * 2 + j] = w.raw.G[r][j];
      }
    }
    status[i] = w.status;
    iters[i] = w.iters;
  }
}

int hs_sizeof_unicycle_params() { return (int)sizeof(UnicycleParams); }
int hs_sizeof_cars_params() { return (int)sizeof(CarsParams); }
}
