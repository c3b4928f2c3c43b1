// rcbf_core.cuh -- per-instance arithmetic of the SAC-RCBF safety hot path.
//
// Everything here is a small, fully unrolled, register-resident function of ONE
// environment instance / ONE QP.  The __global__ kernels in rcbf_kernels.cu map one
// CUDA thread to one instance and call these.  The functions are also compilable
// for the host (RCBF_HD expands to nothing under g++) so the numerics of the exact
// same source can be exercised on a CPU-only box by tests/hostsim -- that build is
// test infrastructure and is never loaded by the product package.
//
// Reference being re-implemented (yemam3/SAC-RCBF):
//   constraint assembly   rcbf_sac/diff_cbf_qp.py:202-266 (Unicycle), :268-357 (SimulatedCars), :362-377 (box rows)
//   row normalisation     rcbf_sac/diff_cbf_qp.py:103-106
//   QP                    qpth PDIPM via rcbf_sac/diff_cbf_qp.py:107,139   (see DESIGN.md for the solver design)
//   clamp                 rcbf_sac/diff_cbf_qp.py:77
//   env steps             envs/unicycle_env.py:46-111,215-280 ; envs/simulated_cars_env.py:38-158
//   prior dynamics        rcbf_sac/dynamics.py:60-105,125-188
#pragma once

#include <math.h>
#include <stdint.h>

#include "../../include/rcbf_b200.h"

#if defined(__CUDACC__)
#define RCBF_HD __host__ __device__ __forceinline__
#define RCBF_HDC __host__ __device__
#define RCBF_UNROLL _Pragma("unroll")
#else
#define RCBF_HD inline __attribute__((always_inline))
#define RCBF_HDC
#define RCBF_UNROLL _Pragma("GCC unroll 16")
#endif

namespace rcbf {

// ------------------------------------------------------------------------------------------------
// scalar helpers
// ------------------------------------------------------------------------------------------------
template <typename T> RCBF_HD T t_abs(T x) { return x < T(0) ? -x : x; }
template <typename T> RCBF_HD T t_max(T a, T b) { return a > b ? a : b; }
template <typename T> RCBF_HD T t_min(T a, T b) { return a < b ? a : b; }
RCBF_HD float t_fma(float a, float b, float c) { return fmaf(a, b, c); }
RCBF_HD double t_fma(double a, double b, double c) { return fma(a, b, c); }
RCBF_HD float t_sqrt(float a) { return sqrtf(a); }
RCBF_HD double t_sqrt(double a) { return sqrt(a); }
// reciprocal square root: one MUFU (+ Newton for double) instead of an IEEE sqrt and an IEEE division; the tiny
// Cholesky factors below only ever need 1/l_kk
RCBF_HD float t_rsqrt(float a) {
#if defined(__CUDA_ARCH__)
  // the bare MUFU.RSQ: rsqrtf() wraps the same instruction in a denormal-input rescue (~6 more instructions) that the
  // arguments here -- squared norms of normalised, P^-1/2-scaled rows and Cholesky pivots, >= 1e-5 -- never need
  float r;
  asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(a));
  return r;
#else
  return 1.0f / sqrtf(a);
#endif
}
RCBF_HD double t_rsqrt(double a) {
#if defined(__CUDA_ARCH__)
  return rsqrt(a);
#else
  return 1.0 / sqrt(a);
#endif
}

// single-rounding float ops that the compiler may not contract into FMAs (reference-order assembly)
// mul_rn flushes denormal products to zero (.ftz) while add_rn / sub_rn do not: ptxas 12.9 contracts a PACKED
// `mul.rn.f32x2` feeding an `add.rn.f32x2` into one FFMA2 (single rounding!) despite the explicit .rn, unless the two
// differ in their ftz mode -- and the scalar instruction uses the same mode so that T = float and T = f2 agree bit for
// bit.  Products below 1.2e-38 do not occur in this path's data (metres, radians, m/s).
RCBF_HD float mul_rn(float a, float b) {
#if defined(__CUDA_ARCH__) && !defined(RCBF_EXP_CONTRACT)  // (experiment switch: let nvcc contract the assembly)
  float r;
  asm("mul.rn.ftz.f32 %0, %1, %2;" : "=f"(r) : "f"(a), "f"(b));
  return r;
#else
  return a * b;  // host build uses -ffp-contract=off
#endif
}
RCBF_HD float add_rn(float a, float b) {
#if defined(__CUDA_ARCH__) && !defined(RCBF_EXP_CONTRACT)
  return __fadd_rn(a, b);
#else
  return a + b;
#endif
}
RCBF_HD float sub_rn(float a, float b) {
#if defined(__CUDA_ARCH__) && !defined(RCBF_EXP_CONTRACT)
  return __fsub_rn(a, b);
#else
  return a - b;
#endif
}

// fast reciprocal used INSIDE the interior-point iteration only (search directions tolerate 1-2 ulp);
// the certificate / outputs use IEEE division.
RCBF_HD float t_rcp_fast(float a) {
#if defined(__CUDA_ARCH__)
  return __frcp_rn(a);
#else
  return 1.0f / a;
#endif
}
RCBF_HD double t_rcp_fast(double a) { return 1.0 / a; }

// a / n with r ~ 1/n: one multiply + one residual correction (the fast path of IEEE division; correctly rounded
// whenever r is within an ulp of 1/n, i.e. outside the denormal/overflow corners that normalised rows never reach)
RCBF_HD float div_by(float a, float n, float r) {
#if defined(__CUDA_ARCH__)
  const float q = a * r;
  return fmaf(fmaf(-n, q, a), r, q);
#else
  (void)r;
  return a / n;
#endif
}
RCBF_HD float rcp_refined(float n) {
#if defined(__CUDA_ARCH__)
  float r;  // MUFU.RCP is within one ulp: enough for the single-correction quotient in div_by to round correctly
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(n));
  return r;
#else
  return 1.0f / n;
#endif
}

}  // namespace rcbf
#include "rcbf_f2.cuh"  // packed pairs of float32 + the op set shared by T = float and T = f2
namespace rcbf {

// per-instance status codes: RCBF_OK_TRIVIAL ... RCBF_NAN, see include/rcbf_b200.h

// ------------------------------------------------------------------------------------------------
// Least-norm point in a polytope:   minimise 1/2 |y|^2   s.t.  A y <= b        (NZ vars, M rows)
//
// Every QP of the path  (min 1/2 x'Px s.t. G~x <= h~, P diagonal, q = 0)  is brought to this form with
// y = sqrt(P) x, A = G~ P^-1/2 (column equilibration).  A general SPD Q / non-zero p go through the Cholesky
// factor of Q (rcbf_generic.cuh).  Duals and slacks are invariant under the change of variables.
// ------------------------------------------------------------------------------------------------
template <typename T, int NZ, int M>
struct LnpProblem {
  T A[M][NZ];
  T b[M];
};

// Compile-time sparsity pattern of A: the actuator rows of both layers touch a single control and never the slack
// (diff_cbf_qp.py:365-377).  IEEE rules forbid the compiler from folding `0 * y`, so the structure is spelled out:
// every A[i][j] product below is guarded by Pat::nz(i, j), a constexpr that the unroller resolves.
struct DensePat {
  RCBF_HDC static constexpr bool nz(int, int) { return true; }
};
template <int NCBF, int NU>
struct CbfPat {  // rows [0, NCBF) dense; row NCBF + 2c and NCBF + 2c + 1 touch column c only
  RCBF_HDC static constexpr bool nz(int i, int j) { return i < NCBF || j == (i - NCBF) / 2; }
};

template <typename T, int NZ, int M>
struct LnpSolution {
  T y[NZ];
  T lam[M];
  T s[M];
  int status;
  int iters;
  uint32_t mask;  // active set of a certified solution (0: trivial); kMaskUnknown when the point is not a vertex guess
};
constexpr uint32_t kMaskUnknown = 0xffffu;

// --- tiny SPD solves (closed-form Cholesky), NZ in {1,2,3} -------------------------------------------
template <typename T>
struct Sym3 {  // symmetric 3x3, lower part
  T a00, a10, a11, a20, a21, a22;
};

template <typename T, int NZ>
struct Chol {
  // L (lower) with the diagonal stored as reciprocals
  T i00, l10, i11, l20, l21, i22;
  RCBF_HD void factor(const T S[NZ][NZ]) {
    i00 = t_rsqrt(S[0][0]);
    if (NZ > 1) {
      l10 = S[1][0] * i00;
      i11 = t_rsqrt(S[1][1] - l10 * l10);
    }
    if (NZ > 2) {
      l20 = S[2][0] * i00;
      l21 = (S[2][1] - l20 * l10) * i11;
      i22 = t_rsqrt(S[2][2] - l20 * l20 - l21 * l21);
    }
  }
  RCBF_HD void fwd(const T r[NZ], T w[NZ]) const {  // L w = r
    w[0] = r[0] * i00;
    if (NZ > 1) w[1] = (r[1] - l10 * w[0]) * i11;
    if (NZ > 2) w[2] = (r[2] - l20 * w[0] - l21 * w[1]) * i22;
  }
  RCBF_HD void bwd(const T w[NZ], T x[NZ]) const {  // L' x = w
    if (NZ > 2) {
      x[2] = w[2] * i22;
      x[1] = (w[1] - l21 * x[2]) * i11;
      x[0] = (w[0] - l10 * x[1] - l20 * x[2]) * i00;
    } else if (NZ > 1) {
      x[1] = w[1] * i11;
      x[0] = (w[0] - l10 * x[1]) * i00;
    } else {
      x[0] = w[0] * i00;
    }
  }
  RCBF_HD void solve(const T r[NZ], T x[NZ]) const {
    T w[NZ];
    fwd(r, w);
    bwd(w, x);
  }
};

// --- KKT certificate ----------------------------------------------------------------------------------
// Given a guessed active set (bit mask, at most NZ rows), solve the equality-constrained problem
//     (A_S A_S') lam_S = -b_S,   y = -A_S' lam_S
// in precision C and accept iff  lam_S >= -tol_l,  b - A y >= -tol_s on every row and |b_S - A_S y| <= tol_s.
// For a strictly convex QP a point passing this test IS the optimum (to tol), independent of how the guess
// was obtained -- this is what lets the float32 interior-point iteration stop as soon as the active set is
// identified instead of having to converge numerically.
// CP supplies the problem data in precision C through a(i,j) / b(i) (compile-time indices after unrolling), so that
// the certificate sees the *unrounded* scaled data even when the iteration runs on a float32 copy.
template <typename C, typename CP, typename Pat, int NZ, int M>
RCBF_HD bool lnp_certify(const CP& P, uint32_t mask, C tol_s, C tol_l, C y[NZ], C lam[M], C s[M]) {
  // gather up to NZ active rows (unused slots: zero row, b = 0  ->  lam = 0 through the unit diagonal below)
  C R[NZ][NZ];
  C rb[NZ];
  RCBF_UNROLL
  for (int k = 0; k < NZ; ++k) {
    RCBF_UNROLL
    for (int j = 0; j < NZ; ++j) R[k][j] = C(0);
    rb[k] = C(0);
  }
  int cnt = 0;
  RCBF_UNROLL
  for (int i = 0; i < M; ++i) {
    const bool act = (mask >> i) & 1u;
    RCBF_UNROLL
    for (int k = 0; k < NZ; ++k) {
      const bool put = act && (cnt == k);
      RCBF_UNROLL
      for (int j = 0; j < NZ; ++j) R[k][j] = put ? (Pat::nz(i, j) ? P.a(i, j) : C(0)) : R[k][j];
      rb[k] = put ? P.b(i) : rb[k];
    }
    cnt += act ? 1 : 0;
  }
  if (cnt > NZ) return false;
  // Gram matrix (unit diagonal on unused slots)
  C Gm[NZ][NZ];
  RCBF_UNROLL
  for (int k = 0; k < NZ; ++k) {
    RCBF_UNROLL
    for (int l = 0; l <= k; ++l) {
      C acc = C(0);
      RCBF_UNROLL
      for (int j = 0; j < NZ; ++j) acc = t_fma(R[k][j], R[l][j], acc);
      Gm[k][l] = acc;
    }
    Gm[k][k] = (k < cnt) ? Gm[k][k] : C(1);
  }
  Chol<C, NZ> ch;
  ch.factor(Gm);
  C nrb[NZ], lk[NZ];
  RCBF_UNROLL
  for (int k = 0; k < NZ; ++k) nrb[k] = -rb[k];
  ch.solve(nrb, lk);
  bool ok = true;
  RCBF_UNROLL
  for (int k = 0; k < NZ; ++k) ok = ok && (lk[k] >= -tol_l);  // NaN (dependent rows) compares false
  RCBF_UNROLL
  for (int j = 0; j < NZ; ++j) {
    C acc = C(0);
    RCBF_UNROLL
    for (int k = 0; k < NZ; ++k) acc = t_fma(-R[k][j], lk[k], acc);
    y[j] = acc;
  }
  int c2 = 0;
  RCBF_UNROLL
  for (int i = 0; i < M; ++i) {
    C acc = P.b(i);
    RCBF_UNROLL
    for (int j = 0; j < NZ; ++j)
      if (Pat::nz(i, j)) acc = t_fma(-P.a(i, j), y[j], acc);
    const bool act = (mask >> i) & 1u;
    C li = C(0);
    RCBF_UNROLL
    for (int k = 0; k < NZ; ++k) li = (act && c2 == k) ? lk[k] : li;
    c2 += act ? 1 : 0;
    lam[i] = li;
    s[i] = acc;
    ok = ok && (acc >= -tol_s) && (!act || acc <= tol_s);
  }
  return ok;
}

// (R R') lam = -rb on the first K of NZ gathered rows, lam = 0 on the rest: the same arithmetic (same expressions,
// same order) as factoring the NZ x NZ Gram matrix with a unit diagonal on the unused slots -- whose extra entries are
// exact zeros -- minus the multiplications by those zeros.
template <int NZ, int K>
RCBF_HD void greedy_solve_rows(const float R[NZ][NZ], const float rbg[NZ], float lk[NZ]) {
  float S[K][K];
  RCBF_UNROLL
  for (int k = 0; k < K; ++k) {
    RCBF_UNROLL
    for (int l = 0; l <= k; ++l) {
      float acc = 0.f;
      RCBF_UNROLL
      for (int j = 0; j < NZ; ++j) acc = fmaf(R[k][j], R[l][j], acc);
      S[k][l] = acc;
    }
  }
  RCBF_UNROLL
  for (int k = 0; k < NZ; ++k) lk[k] = 0.f;
  const float i00 = t_rsqrt(S[0][0]);
  const float w0 = -rbg[0] * i00;
  if (K == 1) {
    lk[0] = w0 * i00;
  } else if (K == 2) {
    const float l10 = S[K > 1 ? 1 : 0][0] * i00;
    const float i11 = t_rsqrt(S[K > 1 ? 1 : 0][K > 1 ? 1 : 0] - l10 * l10);
    const float w1 = (-rbg[1] - l10 * w0) * i11;
    lk[1] = w1 * i11;
    lk[0] = (w0 - l10 * lk[1]) * i00;
  } else {
    const float l10 = S[K > 1 ? 1 : 0][0] * i00;
    const float i11 = t_rsqrt(S[K > 1 ? 1 : 0][K > 1 ? 1 : 0] - l10 * l10);
    const float l20 = S[K > 2 ? 2 : 0][0] * i00;
    const float l21 = (S[K > 2 ? 2 : 0][K > 1 ? 1 : 0] - l20 * l10) * i11;
    const float i22 = t_rsqrt(S[K > 2 ? 2 : 0][K > 2 ? 2 : 0] - l20 * l20 - l21 * l21);
    const float w1 = (-rbg[1] - l10 * w0) * i11;
    const float w2 = (-rbg[K > 2 ? 2 : 0] - l20 * w0 - l21 * w1) * i22;
    lk[K > 2 ? 2 : 0] = w2 * i22;
    lk[1] = (w1 - l21 * lk[K > 2 ? 2 : 0]) * i11;
    lk[0] = (w0 - l10 * lk[1] - l20 * lk[K > 2 ? 2 : 0]) * i00;
  }
}

// multipliers / slacks in dense form, only when a caller wants them saved (backward pass, diagnostics)
template <typename Pat, int NZ, int M>
RCBF_HD void lnp_expand_aux(const float Gn[M][NZ], const float hn[M], const double pis[NZ], const float Rg[NZ][NZ],
                            const float rbg[NZ], uint32_t mask, const double y[NZ], const double lk[NZ], double lam[M],
                            double s[M]) {
  RCBF_UNROLL
  for (int i = 0; i < M; ++i) {
    double acc = (double)hn[i];
    RCBF_UNROLL
    for (int j = 0; j < NZ; ++j)
      if (Pat::nz(i, j)) acc = fma(-((double)Gn[i][j] * pis[j]), y[j], acc);
    s[i] = acc;
    // slot of row i: the gathered row that equals it (selection order is not the bit order)
    double li = 0.0;
    if ((mask >> i) & 1u) {
      RCBF_UNROLL
      for (int k = 0; k < NZ; ++k) {
        bool same = (rbg[k] == hn[i]);
        RCBF_UNROLL
        for (int j = 0; j < NZ; ++j) same = same && (Rg[k][j] == (Pat::nz(i, j) ? Gn[i][j] : 0.f));
        li = same ? lk[k] : li;
      }
    }
    lam[i] = li;
  }
}

template <typename T> struct LnpTol;
template <> struct LnpTol<float> {
  static constexpr int kMaxIter = 16;
  static constexpr int kFirstCert = 1;   // first iteration at which the certificate may be tried (if the mask repeats)
  static constexpr int kAlwaysCert = 2;  // from here on it is tried every iteration
  static constexpr float kResidTol = 0.0f;  // float32 never accepts on the IPM residual: certificate or straggler
};
template <> struct LnpTol<double> {
  static constexpr int kMaxIter = 60;
  static constexpr int kFirstCert = 1;
  static constexpr int kAlwaysCert = 2;
  static constexpr double kResidTol = 1e-11;
};

// ---- the interior point as a resumable state machine: ipm_init once, then ipm_step until it returns a final
// status.  (The persistent "pdipm" kernel refills a lane with a new problem as soon as its QP is done, so the
// iteration cannot be a closed loop.)
constexpr int IPM_CONTINUE = -1;

template <typename T, int NZ, int M>
struct IpmState {
  T y[NZ], s[M], z[M];
  T best_res;
  T by[NZ], bs[M], bz[M];  // best iterate by residual (only maintained when kTrackBest)
  uint32_t last_mask;
  int it;
};

template <typename T, typename C, int NZ, int M>
RCBF_HD int ipm_finish_best(const IpmState<T, NZ, M>& st, LnpSolution<C, NZ, M>& out, int status) {
  // not certified: hand back the best iterate by residual (what qpth returns)
  RCBF_UNROLL
  for (int j = 0; j < NZ; ++j) out.y[j] = C(st.by[j]);
  RCBF_UNROLL
  for (int i = 0; i < M; ++i) {
    out.lam[i] = C(st.bz[i]);
    out.s[i] = C(st.bs[i]);
  }
  out.status = (st.best_res < T(1e30)) ? status : RCBF_NAN;
  out.iters = st.it;
  out.mask = kMaskUnknown;
  return out.status;
}

template <typename T, typename Pat, int NZ, int M>
RCBF_HD void ipm_init(const LnpProblem<T, NZ, M>& P, IpmState<T, NZ, M>& st) {
  // initial point (qpth): (I + A'A) y = A'b ; s = b - A y ; z = -s ; shift both to >= 1 if needed
  {
    T S[NZ][NZ], r[NZ];
    RCBF_UNROLL
    for (int j = 0; j < NZ; ++j) {
      RCBF_UNROLL
      for (int k = 0; k <= j; ++k) {
        T acc = (j == k) ? T(1) : T(0);
        RCBF_UNROLL
        for (int i = 0; i < M; ++i)
          if (Pat::nz(i, j) && Pat::nz(i, k)) acc = t_fma(P.A[i][j], P.A[i][k], acc);
        S[j][k] = acc;
      }
      T acc = T(0);
      RCBF_UNROLL
      for (int i = 0; i < M; ++i)
        if (Pat::nz(i, j)) acc = t_fma(P.A[i][j], P.b[i], acc);
      r[j] = acc;
    }
    Chol<T, NZ> ch;
    ch.factor(S);
    ch.solve(r, st.y);
    T smin = T(1e30), zmin = T(1e30);
    RCBF_UNROLL
    for (int i = 0; i < M; ++i) {
      T acc = P.b[i];
      RCBF_UNROLL
      for (int j = 0; j < NZ; ++j)
        if (Pat::nz(i, j)) acc = t_fma(-P.A[i][j], st.y[j], acc);
      st.s[i] = acc;
      st.z[i] = -acc;
      smin = t_min(smin, st.s[i]);
      zmin = t_min(zmin, st.z[i]);
    }
    const T sshift = smin < T(0) ? (T(1) - smin) : T(0);
    const T zshift = zmin < T(0) ? (T(1) - zmin) : T(0);
    RCBF_UNROLL
    for (int i = 0; i < M; ++i) {
      st.s[i] += sshift;
      st.z[i] += zshift;
    }
  }

  st.best_res = T(1e30);
  RCBF_UNROLL
  for (int j = 0; j < NZ; ++j) st.by[j] = st.y[j];
  RCBF_UNROLL
  for (int i = 0; i < M; ++i) {
    st.bs[i] = st.s[i];
    st.bz[i] = st.z[i];
  }
  st.last_mask = 0xffffffffu;
  st.it = 0;
}

// one Mehrotra predictor-corrector iteration (+ the certificate attempt).  Returns IPM_CONTINUE or the final status
// (out is filled in the latter case).
template <typename T, typename C, typename CP, typename Pat, int NZ, int M, bool kTrackBest>
RCBF_HD int ipm_step(const LnpProblem<T, NZ, M>& P, const CP& cp, IpmState<T, NZ, M>& st, LnpSolution<C, NZ, M>& out,
                     C tol_s, C tol_l) {
  if (st.it >= LnpTol<T>::kMaxIter) return ipm_finish_best<T, C, NZ, M>(st, out, RCBF_MAXITER);
  {
  // residuals
  T rx[NZ], rz[M];
  T mu = T(0), nrx = T(0), nrz = T(0);
  RCBF_UNROLL
  for (int j = 0; j < NZ; ++j) {
    T acc = st.y[j];
    RCBF_UNROLL
    for (int i = 0; i < M; ++i)
      if (Pat::nz(i, j)) acc = t_fma(P.A[i][j], st.z[i], acc);
    rx[j] = acc;
    nrx = t_fma(acc, acc, nrx);
  }
  uint32_t mask = 0;
  RCBF_UNROLL
  for (int i = 0; i < M; ++i) {
    T acc = st.s[i] - P.b[i];
    RCBF_UNROLL
    for (int j = 0; j < NZ; ++j)
      if (Pat::nz(i, j)) acc = t_fma(P.A[i][j], st.y[j], acc);
    rz[i] = acc;
    nrz = t_fma(acc, acc, nrz);
    mu = t_fma(st.s[i], st.z[i], mu);
    mask |= (st.z[i] > st.s[i]) ? (1u << i) : 0u;
  }
  const T res = t_sqrt(nrx) + t_sqrt(nrz) + mu;  // qpth's resid: |rx| + |rz| + m*mu
  if (!(res == res)) return ipm_finish_best<T, C, NZ, M>(st, out, RCBF_MAXITER);  // overflow / breakdown
  if (kTrackBest && res < st.best_res) {
    st.best_res = res;
    RCBF_UNROLL
    for (int j = 0; j < NZ; ++j) st.by[j] = st.y[j];
    RCBF_UNROLL
    for (int i = 0; i < M; ++i) {
      st.bs[i] = st.s[i];
      st.bz[i] = st.z[i];
    }
  }
  // certified early exit: try the predicted active set whenever it is a candidate
  if (st.it >= LnpTol<T>::kFirstCert && (mask == st.last_mask || st.it >= LnpTol<T>::kAlwaysCert)) {
    if (lnp_certify<C, CP, Pat, NZ, M>(cp, mask, tol_s, tol_l, out.y, out.lam, out.s)) {
      out.status = RCBF_OK_CERTIFIED;
      out.iters = st.it;
      out.mask = mask;
      return RCBF_OK_CERTIFIED;
    }
  }
  st.last_mask = mask;
  if (res < LnpTol<T>::kResidTol) return ipm_finish_best<T, C, NZ, M>(st, out, RCBF_OK_IPM);
  mu *= T(1.0 / M);

  // scaling, normal matrix S = I + A' D A
  T w[M], d[M];
  RCBF_UNROLL
  for (int i = 0; i < M; ++i) {
    w[i] = t_rcp_fast(st.s[i] * st.z[i]);   // 1/(s z): 1/s = w z, 1/z = w s  (one reciprocal per row)
    d[i] = st.z[i] * st.z[i] * w[i];
  }
  T S[NZ][NZ];
  RCBF_UNROLL
  for (int j = 0; j < NZ; ++j) {
    RCBF_UNROLL
    for (int k = 0; k <= j; ++k) {
      T acc = (j == k) ? T(1) : T(0);
      RCBF_UNROLL
      for (int i = 0; i < M; ++i)
        if (Pat::nz(i, j) && Pat::nz(i, k)) acc = t_fma(P.A[i][j] * d[i], P.A[i][k], acc);
      S[j][k] = acc;
    }
  }
  Chol<T, NZ> ch;
  ch.factor(S);

  // affine direction: S dy = -rx + A'(z - d rz)
  T r[NZ], dy[NZ], ds[M], dz[M];
  RCBF_UNROLL
  for (int j = 0; j < NZ; ++j) {
    T acc = -rx[j];
    RCBF_UNROLL
    for (int i = 0; i < M; ++i)
      if (Pat::nz(i, j)) acc = t_fma(P.A[i][j], t_fma(-d[i], rz[i], st.z[i]), acc);
    r[j] = acc;
  }
  ch.solve(r, dy);
  T rho = T(0);
  RCBF_UNROLL
  for (int i = 0; i < M; ++i) {
    T acc = -rz[i];
    RCBF_UNROLL
    for (int j = 0; j < NZ; ++j)
      if (Pat::nz(i, j)) acc = t_fma(-P.A[i][j], dy[j], acc);
    ds[i] = acc;
    dz[i] = t_fma(-d[i], acc, -st.z[i]);
    rho = t_max(rho, -ds[i] * (w[i] * st.z[i]));
    rho = t_max(rho, -dz[i] * (w[i] * st.s[i]));
  }
  T alpha = rho > T(1) ? T(1) / rho : T(1);
  T t3 = T(0), t4 = T(0);
  RCBF_UNROLL
  for (int i = 0; i < M; ++i) {
    t3 = t_fma(t_fma(alpha, ds[i], st.s[i]), t_fma(alpha, dz[i], st.z[i]), t3);
    t4 = t_fma(st.s[i], st.z[i], t4);
  }
  T sig = t3 / t4;
  sig = sig * sig * sig;
  const T musig = mu * sig;
  // corrector folded into one combined solve: rs_tot = z + (-mu sig + ds_aff dz_aff)/s
  T rsc[M];
  RCBF_UNROLL
  for (int i = 0; i < M; ++i) rsc[i] = (t_fma(ds[i], dz[i], -musig)) * (w[i] * st.z[i]);
  RCBF_UNROLL
  for (int j = 0; j < NZ; ++j) {
    T acc = r[j];
    RCBF_UNROLL
    for (int i = 0; i < M; ++i)
      if (Pat::nz(i, j)) acc = t_fma(P.A[i][j], rsc[i], acc);
    r[j] = acc;
  }
  ch.solve(r, dy);
  rho = T(0);
  RCBF_UNROLL
  for (int i = 0; i < M; ++i) {
    T acc = -rz[i];
    RCBF_UNROLL
    for (int j = 0; j < NZ; ++j)
      if (Pat::nz(i, j)) acc = t_fma(-P.A[i][j], dy[j], acc);
    ds[i] = acc;
    dz[i] = t_fma(-d[i], acc, -(st.z[i] + rsc[i]));
    rho = t_max(rho, -ds[i] * (w[i] * st.z[i]));
    rho = t_max(rho, -dz[i] * (w[i] * st.s[i]));
  }
  alpha = rho > T(0.999) ? T(0.999) / rho : T(1);
  RCBF_UNROLL
  for (int j = 0; j < NZ; ++j) st.y[j] = t_fma(alpha, dy[j], st.y[j]);
  RCBF_UNROLL
  for (int i = 0; i < M; ++i) {
    st.s[i] = t_fma(alpha, ds[i], st.s[i]);
    st.z[i] = t_fma(alpha, dz[i], st.z[i]);
  }
  }
  st.it += 1;
  return IPM_CONTINUE;
}

// --- primal-dual interior point (Mehrotra predictor-corrector) with certified early exit -------------------
// T : precision of the iteration (float on the main path, double in the straggler pass)
// C : precision of the certificate (double: ~100 DFMA, B200 runs FP64 at 1:2)
template <typename T, typename C, typename CP, typename Pat, int NZ, int M>
RCBF_HD void lnp_solve(const LnpProblem<T, NZ, M>& P, const CP& cp, LnpSolution<C, NZ, M>& out, C tol_s, C tol_l) {
  // 0. trivial certificate: y = 0 is feasible  <=>  b >= 0 on every row
  {
    bool triv = true;
    bool nan = false;
    RCBF_UNROLL
    for (int i = 0; i < M; ++i) {
      triv = triv && (P.b[i] >= T(0));
      nan = nan || (P.b[i] != P.b[i]);
      RCBF_UNROLL
      for (int j = 0; j < NZ; ++j) nan = nan || (P.A[i][j] != P.A[i][j]);
    }
    if (triv || nan) {
      RCBF_UNROLL
      for (int j = 0; j < NZ; ++j) out.y[j] = nan ? C(NAN) : C(0);
      RCBF_UNROLL
      for (int i = 0; i < M; ++i) {
        out.lam[i] = C(0);
        out.s[i] = C(P.b[i]);
      }
      out.status = nan ? RCBF_NAN : RCBF_OK_TRIVIAL;
      out.iters = 0;
      out.mask = 0u;
      return;
    }
  }

  IpmState<T, NZ, M> st;
  ipm_init<T, Pat, NZ, M>(P, st);
  while (ipm_step<T, C, CP, Pat, NZ, M, true>(P, cp, st, out, tol_s, tol_l) == IPM_CONTINUE) {
  }
}

// ------------------------------------------------------------------------------------------------
// Parameters (plain C structs; mirrored by include/rcbf_b200.h)
// ------------------------------------------------------------------------------------------------
constexpr int kUniHaz = 5;  // envs/unicycle_env.py:26 (the layer reads len(env.hazards_locations), diff_cbf_qp.py:35)
constexpr int kUniM = kUniHaz + 4;
constexpr int kUniNZ = 3;
constexpr int kCarsM = 4;
constexpr int kCarsNZ = 2;
using UniPat = CbfPat<kUniHaz, 2>;
using CarsPat = CbfPat<2, 1>;

using UnicycleParams = rcbf_unicycle_params;  // include/rcbf_b200.h
using CarsParams = rcbf_cars_params;

// ------------------------------------------------------------------------------------------------
// Assembly: raw (G, h) exactly as get_cbf_qp_constraints returns them (before normalisation)
// ------------------------------------------------------------------------------------------------
struct UniRaw {
  float G[kUniM][kUniNZ];
  float h[kUniM];
  float Lg[kUniHaz][2];  // d h_i / d action  (= -G[i][:2]); kept for the backward chain
};

// sin and cos of a float32 angle, <= ~1.5 ulp for |x| < 1e5 (env headings stay below ~25 rad: theta drifts by at
// most dt * 1 rad/s per step over 1000 steps).  Cody-Waite reduction by pi/2 in three float32 pieces + the classic
// degree-7 / degree-8 minimax polynomials on [-pi/4, pi/4].  ~30 instructions, no slow path, no local memory (the
// library sincosf carries a Payne-Hanek branch that costs code size and registers in every kernel that calls it).
template <typename T>
RCBF_HD void sincos_v(T x, T* sn, T* cs) {  // T = float or f2 (two instances through the packed FP32 instructions)
  const T j = t_rint(mul_rn(x, T(0.636619772f)));       // nearest multiple of pi/2
  T r = t_fma(j, T(-1.57079601e+00f), x);
  r = t_fma(j, T(-3.13916473e-07f), r);
  r = t_fma(j, T(-5.39030253e-15f), r);
  const typename VecOf<T>::ivec q = t_toint(j);
  const T z = mul_rn(r, r);
  T ps = t_fma(z, T(-1.9515295891e-4f), T(8.3321608736e-3f));
  ps = t_fma(ps, z, T(-1.6666654611e-1f));
  ps = t_fma(mul_rn(ps, z), r, r);                      // sin(r)
  T pc = t_fma(z, T(2.443315711809948e-5f), T(-1.388731625493765e-3f));
  pc = t_fma(pc, z, T(4.166664568298827e-2f));
  pc = t_fma(mul_rn(pc, z), z, t_fma(z, T(-0.5f), T(1.0f)));  // cos(r)
  const typename VecOf<T>::mask swap = t_bit(q, 1);
  T sv = t_sel(swap, pc, ps);
  T cv = t_sel(swap, ps, pc);
  sv = t_sel(t_bit(q, 2), t_neg(sv), sv);
  cv = t_sel(t_bit(t_iadd(q, 1), 2), t_neg(cv), cv);
  *sn = sv;
  *cs = cv;
}
RCBF_HD void sincos_t(float x, float* sn, float* cs) { sincos_v<float>(x, sn, cs); }

// The assembly mirrors the reference's float32 operation ORDER (one rounding per torch op, no FMA contraction, the
// k-ordered accumulation of torch.bmm's small-matrix path) so that the QP data agree with the reference's to the
// last bit up to the ulp of cos/sin: near-degenerate instances amplify 1e-7 data noise into >1e-4 action noise.
// Written once for T = float (one instance) and T = f2 (two instances per lane, FMUL2 / FADD2): the packed
// instructions round each half exactly like the scalar ones, so both instantiations give the same bits.
// Output: Lg[i] = d h_i / d action (G[i][:2] = -Lg[i]) and the 9 right-hand sides.
// NH = number of hazards (CBF rows); hz = their centres.  The hot kernels use NH = kUniHaz with p.hazards (the
// reference env, unicycle_env.py:26); rcbf_general.cu instantiates larger NH for layers built on other hazard sets
// (the reference sizes the layer from len(env.hazards_locations), diff_cbf_qp.py:35,243-261).
template <typename T, int NH>
RCBF_HD void assemble_unicycle_n(const UnicycleParams& p, const float (*hz)[2], const T st[3], T s, T c, const T u[2],
                                 const T mu[3], const T sg[3], T Lg[NH][2], T h[NH + 4]) {
  // s, c = sin / cos of st[2]                              // :211-212
  const T lp = T(p.l_p);
  const T px = add_rn(st[0], mul_rn(lp, c));               // :216
  const T py = add_rn(st[1], mul_rn(lp, s));               // :217
  const T g01 = t_neg(mul_rn(s, lp)), g11 = mul_rn(c, lp); // g_p = R diag(1,l_p) = [[c, g01],[s, g11]]  :225-233
  const T mpx = add_rn(mul_rn(g01, mu[2]), mu[0]);         // :236-238
  const T mpy = add_rn(mul_rn(g11, mu[2]), mu[1]);
  const T a01 = p.abs_sigma_map ? t_fabs(g01) : g01;
  const T a11 = p.abs_sigma_map ? t_fabs(g11) : g11;
  const T spx = mul_rn(T(p.sigma_scale), add_rn(mul_rn(a01, sg[2]), sg[0]));  // :239-241 (scale = 1 in this layer)
  const T spy = mul_rn(T(p.sigma_scale), add_rn(mul_rn(a11, sg[2]), sg[1]));
  RCBF_UNROLL
  for (int i = 0; i < NH; ++i) {
    const T dx = sub_rn(px, T(hz[i][0]));                  // :248
    const T dy = sub_rn(py, T(hz[i][1]));
    const T hc = mul_rn(T(0.5f), sub_rn(add_rn(mul_rn(dx, dx), mul_rn(dy, dy)), T(p.collision_radius_sq)));  // :246
    const T L0 = add_rn(mul_rn(dx, c), mul_rn(dy, s));     // dhdp' g_p   :259
    const T L1 = add_rn(mul_rn(dx, g01), mul_rn(dy, g11));
    Lg[i][0] = L0;
    Lg[i][1] = L1;
    const T t1 = add_rn(mul_rn(dx, mpx), mul_rn(dy, mpy));
    const T t2 = add_rn(mul_rn(t_fabs(dx), spx), mul_rn(t_fabs(dy), spy));
    const T t3 = add_rn(mul_rn(L0, u[0]), mul_rn(L1, u[1]));
    h[i] = add_rn(mul_rn(T(p.gamma_b), mul_rn(mul_rn(hc, hc), hc)), add_rn(sub_rn(t1, t2), t3));  // :261
  }
  RCBF_UNROLL
  for (int cc = 0; cc < 2; ++cc) {  // :365-377
    const int r = NH + 2 * cc;
    h[r] = sub_rn(T(p.u_max[cc]), u[cc]);
    h[r + 1] = add_rn(T(-p.u_min[cc]), u[cc]);
  }
}

template <typename T>
RCBF_HD void assemble_unicycle_v(const UnicycleParams& p, const T st[3], T s, T c, const T u[2], const T mu[3],
                                 const T sg[3], T Lg[kUniHaz][2], T h[kUniM]) {
  assemble_unicycle_n<T, kUniHaz>(p, p.hazards, st, s, c, u, mu, sg, Lg, h);
}

RCBF_HD void assemble_unicycle_sc(const UnicycleParams& p, const float st[3], float s, float c, const float u[2],
                                  const float mu[3], const float sg[3], UniRaw& o) {
  assemble_unicycle_v<float>(p, st, s, c, u, mu, sg, o.Lg, o.h);
  RCBF_UNROLL
  for (int i = 0; i < kUniHaz; ++i) {
    o.G[i][0] = -o.Lg[i][0];
    o.G[i][1] = -o.Lg[i][1];
    o.G[i][2] = -1.0f;                                     // :260
  }
  RCBF_UNROLL
  for (int cc = 0; cc < 2; ++cc) {  // :365-377
    const int r = kUniHaz + 2 * cc;
    RCBF_UNROLL
    for (int j = 0; j < 3; ++j) {
      o.G[r][j] = (j == cc) ? 1.0f : 0.0f;
      o.G[r + 1][j] = (j == cc) ? -1.0f : 0.0f;
    }
  }
}

RCBF_HD void assemble_unicycle(const UnicycleParams& p, const float st[3], const float u[2], const float mu[3],
                               const float sg[3], UniRaw& o) {
  float s, c;
  sincos_t(st[2], &s, &c);
  assemble_unicycle_sc(p, st, s, c, u, mu, sg, o);
}

struct CarsRaw {
  float G[kCarsM][kCarsNZ];
  float h[kCarsM];
  float Lg[2];
};

// prior accelerations shared by the layer (:284-290), the prior model (dynamics.py:171-177) and the env
// (simulated_cars_env.py:58-64).  v0des = desired velocity of the lead car.  T = float (layer) or double (env/prior).
template <typename T>
RCBF_HD void cars_accels(T kp, T kb, const T pos[5], const T vel[5], T v0des, T acc[5]) {
  acc[0] = kp * (v0des - vel[0]);
  RCBF_UNROLL
  for (int i = 1; i < 5; ++i) acc[i] = kp * (T(30) - vel[i]);
  const T d01 = pos[0] - pos[1], d12 = pos[1] - pos[2], d24 = pos[2] - pos[4];
  acc[1] -= (d01 < T(6)) ? kb * d01 : T(0);
  acc[2] -= (d12 < T(6)) ? kb * d12 : T(0);
  acc[4] -= (d24 < T(13)) ? kb * d24 : T(0);
}

RCBF_HD void assemble_cars(const CarsParams& p, const float st[10], float u, const float sg[10], CarsRaw& o) {
  float pos[5], vel[5], acc[5];
  RCBF_UNROLL
  for (int i = 0; i < 5; ++i) {
    pos[i] = st[2 * i];
    vel[i] = st[2 * i + 1];
  }
  // :284-290 (no lead-car sinusoid in the layer, :285); each op rounded once like the torch expression
  RCBF_UNROLL
  for (int i = 0; i < 5; ++i) acc[i] = mul_rn(p.kp, sub_rn(30.0f, vel[i]));
  const float d01 = sub_rn(pos[0], pos[1]), d12 = sub_rn(pos[1], pos[2]), d24 = sub_rn(pos[2], pos[4]);
  acc[1] = (d01 < 6.0f) ? sub_rn(acc[1], mul_rn(p.k_brake, d01)) : acc[1];
  acc[2] = (d12 < 6.0f) ? sub_rn(acc[2], mul_rn(p.k_brake, d12)) : acc[2];
  acc[4] = (d24 < 13.0f) ? sub_rn(acc[4], mul_rn(p.k_brake, d24)) : acc[4];
  // acc[3] = 0 (:289): its products vanish below
  const float d23 = sub_rn(pos[2], pos[3]), d43 = sub_rn(pos[4], pos[3]);
  const float h13 = mul_rn(0.5f, sub_rn(mul_rn(d23, d23), p.collision_radius_sq));  // :306
  const float h15 = mul_rn(0.5f, sub_rn(mul_rn(d43, d43), p.collision_radius_sq));  // :307
  const float a7 = sub_rn(pos[3], pos[2]), b7 = sub_rn(pos[3], pos[4]);
  const float a6 = sub_rn(vel[3], vel[2]), b6 = sub_rn(vel[3], vel[4]);
  const float h13d = mul_rn(a7, a6);  // :310
  const float h15d = mul_rn(b7, b6);  // :311
  const float a4 = sub_rn(vel[2], vel[3]), a5 = d23;  // :314-318
  const float b8 = sub_rn(vel[4], vel[3]), b9 = d43;  // :322-326
  // bmm accumulates in state-index order; the a7*acc3 / b7*acc3 terms are exact zeros           :319,:327
  const float Lff13 = add_rn(add_rn(mul_rn(a4, vel[2]), mul_rn(a5, acc[2])), mul_rn(a6, vel[3]));
  const float Lff15 = add_rn(add_rn(mul_rn(b6, vel[3]), mul_rn(b8, vel[4])), mul_rn(b9, acc[4]));
  const float ss = p.sigma_scale;
  const float LfD13 = mul_rn(ss, add_rn(mul_rn(fabsf(a5), sg[5]), mul_rn(fabsf(a7), sg[7])));  // :320 (:299 odd entries)
  const float LfD15 = mul_rn(ss, add_rn(mul_rn(fabsf(b7), sg[7]), mul_rn(fabsf(b9), sg[9])));  // :328
  const float Lg13 = mul_rn(a7, 50.0f), Lg15 = mul_rn(b7, 50.0f);                              // :331-332 (g = 50 e_7)
  o.Lg[0] = Lg13;
  o.Lg[1] = Lg15;
  // :348-349, evaluated left to right
  o.h[0] = add_rn(add_rn(add_rn(sub_rn(Lff13, LfD13), mul_rn(p.gamma_2, h13d)), mul_rn(p.gamma_sq, h13)), mul_rn(Lg13, u));
  o.h[1] = add_rn(add_rn(add_rn(sub_rn(Lff15, LfD15), mul_rn(p.gamma_2, h15d)), mul_rn(p.gamma_sq, h15)), mul_rn(Lg15, u));
  o.G[0][0] = -Lg13;  // :350
  o.G[1][0] = -Lg15;  // :351
  o.G[0][1] = -p.slack_coeff;  // :352
  o.G[1][1] = -p.slack_coeff;
  o.G[2][0] = 1.0f;  // :369-370
  o.G[2][1] = 0.0f;
  o.h[2] = sub_rn(p.u_max, u);
  o.G[3][0] = -1.0f;  // :375-376
  o.G[3][1] = 0.0f;
  o.h[3] = add_rn(-p.u_min, u);
}

// ------------------------------------------------------------------------------------------------
// Row normalisation (diff_cbf_qp.py:103-106) + column equilibration -> least-norm problem
// ------------------------------------------------------------------------------------------------
template <int NZ, int M>
struct Normalised {
  float Gn[M][NZ];
  float hn[M];
  float n[M];
  uint32_t h_is_max;  // bit i: |h_i| is the (strict) row maximum  -> routes d n_i / d h_i in the backward
};

template <typename Pat, int NZ, int M>
RCBF_HD void normalise_rows(const float G[M][NZ], const float h[M], Normalised<NZ, M>& o) {
  o.h_is_max = 0;
  RCBF_UNROLL
  for (int i = 0; i < M; ++i) {
    float gm = 0.0f;
    RCBF_UNROLL
    for (int j = 0; j < NZ; ++j)
      if (Pat::nz(i, j)) gm = fmaxf(gm, fabsf(G[i][j]));
    const float ha = fabsf(h[i]);
    // torch.max returns the FIRST maximal index on ties; h is the last column (diff_cbf_qp.py:103-104)
    if (ha > gm) o.h_is_max |= (1u << i);
    const float n = (ha != ha) ? ha : fmaxf(gm, ha);  // propagate NaN like torch.max
    o.n[i] = n;
    const float r = rcp_refined(n);
    RCBF_UNROLL
    for (int j = 0; j < NZ; ++j) o.Gn[i][j] = Pat::nz(i, j) ? div_by(G[i][j], n, r) : 0.0f;
    o.hn[i] = div_by(h[i], n, r);
  }
}

template <typename T, typename Pat, int NZ, int M>
RCBF_HD void to_lnp(const Normalised<NZ, M>& nrm, const T pis[NZ] /* P^-1/2 */, LnpProblem<T, NZ, M>& P) {
  RCBF_UNROLL
  for (int i = 0; i < M; ++i) {
    RCBF_UNROLL
    for (int j = 0; j < NZ; ++j) P.A[i][j] = Pat::nz(i, j) ? T(nrm.Gn[i][j]) * pis[j] : T(0);
    P.b[i] = T(nrm.hn[i]);
  }
}

// certificate data of a normalised, column-equilibrated QP in float64: a(i,j) = G~_ij * P_j^-1/2 (exact product of
// the float32 datum the reference also sees and the float64 scale), b(i) = h~_i
template <int NZ, int M>
struct NormCert {
  const Normalised<NZ, M>& nrm;
  const double* pis;
  RCBF_HD double a(int i, int j) const { return (double)nrm.Gn[i][j] * pis[j]; }
  RCBF_HD double b(int i) const { return (double)nrm.hn[i]; }
};

// certificate tolerances (normalised rows have |entries| <= 1)
constexpr double kTolSlack = 1e-9;
constexpr double kTolDual = 1e-9;

#define RCBF_PENDING 5  /* internal: not certified by the fast path, queued for the interior-point pass */

template <int NZ, int M>
struct NormSolution {
  double x[NZ], lam[M], s[M];
  int status, iters;
  uint32_t mask;  // active set the certificate accepted (0: trivial; kMaskUnknown: interior-point iterate)
};

template <int NZ, int M>
RCBF_HD void pis_of(const float p_diag[NZ], double pisd[NZ], float pisf[NZ]) {
  RCBF_UNROLL
  for (int j = 0; j < NZ; ++j) {
    pisd[j] = 1.0 / sqrt((double)p_diag[j]);
    pisf[j] = (float)pisd[j];
  }
}

// Fast path of one normalised QP in "pdipm" solver mode: the float32 interior point with the float64 KKT certificate
// tried on its active-set prediction.  Leaves status = RCBF_PENDING when it cannot certify: the fallback pass takes
// over.  (Presolve mode: solve_raw_fast below.)
template <typename Pat, int NZ, int M>
RCBF_HD void solve_normalised_ipm(const Normalised<NZ, M>& nrm, const float p_diag[NZ], NormSolution<NZ, M>& o) {
  double pisd[NZ];
  float pisf[NZ];
  pis_of<NZ, M>(p_diag, pisd, pisf);
  const NormCert<NZ, M> cp{nrm, pisd};
  LnpProblem<float, NZ, M> Pf;
  to_lnp<float, Pat, NZ, M>(nrm, pisf, Pf);
  LnpSolution<double, NZ, M> sol;
  lnp_solve<float, double, NormCert<NZ, M>, Pat, NZ, M>(Pf, cp, sol, kTolSlack, kTolDual);
  o.status = (sol.status >= RCBF_MAXITER) ? RCBF_PENDING : sol.status;
  o.iters = sol.iters;
  o.mask = sol.mask;
  RCBF_UNROLL
  for (int i = 0; i < M; ++i) {
    o.lam[i] = sol.lam[i];
    o.s[i] = sol.s[i];
  }
  RCBF_UNROLL
  for (int j = 0; j < NZ; ++j) o.x[j] = sol.y[j] * pisd[j];
}

// ---- presolve on the rows AS ASSEMBLED ---------------------------------------------------------------------------
// The violation of row i in units of its norm, (h_i - A_i y) / |A_i| with A = G P^-1/2, does not depend on the row
// normalisation of diff_cbf_qp.py:103-106 (n_i cancels), so the greedy pass scores the raw rows and only the <= NZ rows
// that enter the active set are normalised (correctly rounded quotients, the float32 data the reference's QP sees).
// The float64 certificate then solves on those rows exactly as before; the rows OUTSIDE the active set are screened in
// float32 first: the sign of a slack does not depend on n_i > 0 either, so a raw slack that is positive by more than
// its own rounding-error bound is feasible, and only a row inside that band is normalised and checked in float64.
// Decisions and results are those of the all-rows float64 test; a QP costs ~150 instructions less.
RCBF_HD float key_of(float v, int i) {  // v with the row index in the 4 low mantissa bits: min over keys = arg min
#if defined(__CUDA_ARCH__)
  return __uint_as_float((__float_as_uint(v) & 0xfffffff0u) | (uint32_t)i);
#else
  union { float f; uint32_t u; } c;
  c.f = v;
  c.u = (c.u & 0xfffffff0u) | (uint32_t)i;
  return c.f;
#endif
}
RCBF_HD int key_index(float k) {
#if defined(__CUDA_ARCH__)
  return (int)(__float_as_uint(k) & 15u);
#else
  union { float f; uint32_t u; } c;
  c.f = k;
  return (int)(c.u & 15u);
#endif
}
RCBF_HD float min3f(float a, float b, float c) { return fminf(fminf(a, b), c); }  // NaN operands are ignored

// --- greedy dual active-set presolve -----------------------------------------------------------------------
// Starting from y = 0 (the unconstrained optimum), repeatedly add the most violated row (violation measured in
// units of the row norm) and re-solve the equality-constrained least-norm problem on the chosen rows, at most NZ
// times.  This is Goldfarb-Idnani without constraint dropping: it stops with `false` as soon as a multiplier turns
// negative (a drop would be needed) or NZ rows do not make the point feasible -- those instances go to the
// interior-point solver.  The returned mask is only a GUESS; the float64 certificate decides.
// (Rows are scored as assembled, see above; a row that entered the active set is never picked again.)
// How the presolve gets hold of the raw row it has just picked (index wi, a run-time value): by default a chain of
// selects over the register-resident rows; a caller whose problem also sits in addressable memory (the shared-memory ring
// of k_safe2) passes a fetcher that reads the row there instead -- same values, a fraction of the instructions.
template <typename Pat, int NZ, int M>
struct SelectRowFetch {
  RCBF_HD void operator()(int wi, const float G[M][NZ], const float h[M], float g[NZ], float& hh) const {
    RCBF_UNROLL
    for (int j = 0; j < NZ; ++j) g[j] = 0.f;
    hh = 0.f;
    RCBF_UNROLL
    for (int i = 0; i < M; ++i) {
      const bool put = (i == wi);
      RCBF_UNROLL
      for (int j = 0; j < NZ; ++j)
        if (Pat::nz(i, j)) g[j] = put ? G[i][j] : g[j];
      hh = put ? h[i] : hh;
    }
  }
};

template <typename Pat, int NZ, int M, typename Fetch = SelectRowFetch<Pat, NZ, M>>
RCBF_HD bool lnp_greedy_raw(const float G[M][NZ], const float h[M], const float pis[NZ], float A[M][NZ],
                            float Rg[NZ][NZ], float rbg[NZ], uint32_t& mask_out, int& rounds,
                            const Fetch& fetch = Fetch()) {
  static_assert(M <= 16, "the row index travels in 4 mantissa bits");
  float inv_norm[M];
  RCBF_UNROLL
  for (int i = 0; i < M; ++i) {
    float acc = 0.f;
    RCBF_UNROLL
    for (int j = 0; j < NZ; ++j) {
      A[i][j] = Pat::nz(i, j) ? G[i][j] * pis[j] : 0.f;
      if (Pat::nz(i, j)) acc = fmaf(A[i][j], A[i][j], acc);
    }
    inv_norm[i] = t_rsqrt(acc);
  }
  float y[NZ], R[NZ][NZ];
  RCBF_UNROLL
  for (int j = 0; j < NZ; ++j) {
    y[j] = 0.f;
    rbg[j] = 0.f;
    RCBF_UNROLL
    for (int k = 0; k < NZ; ++k) {
      R[j][k] = 0.f;
      Rg[j][k] = 0.f;
    }
  }
  uint32_t mask = 0;
  bool ok = true, feasible = false;
  int r = 0;
  RCBF_UNROLL
  for (; r <= NZ; ++r) {
    // most violated row at the current y: a tree of 3-input minima over (violation, index) keys
    float key[M];
    RCBF_UNROLL
    for (int i = 0; i < M; ++i) {
      float acc = h[i];
      if (r > 0) {  // (round 0 starts from y = 0: the products are exact zeros)
        RCBF_UNROLL
        for (int j = 0; j < NZ; ++j)
          if (Pat::nz(i, j)) acc = fmaf(-A[i][j], y[j], acc);
      }
      key[i] = key_of(acc * inv_norm[i], i);
    }
    float worst = -1e-6f;
    RCBF_UNROLL
    for (int i = 0; i + 2 < M; i += 3) worst = fminf(worst, min3f(key[i], key[i + 1], key[i + 2]));
    RCBF_UNROLL
    for (int i = M - M % 3; i < M; ++i) worst = fminf(worst, key[i]);
    if (!(worst < -1e-6f)) {
      feasible = true;
      break;
    }
    if (r == NZ) break;  // NZ rows and still infeasible
    const int wi = key_index(worst);
    mask |= 1u << wi;
    // gather the raw row and normalise it (diff_cbf_qp.py:103-106): these are the QP data of the reference
    float g[NZ], hh;
    fetch(wi, G, h, g, hh);
    RCBF_UNROLL
    for (int i = 0; i < M; ++i)               // a row of the active set is never picked again (its raw slack is only
      inv_norm[i] = (i == wi) ? 0.f : inv_norm[i];  // zero up to the rounding of the normalised quotients)
    {
      float gm = 0.f;
      RCBF_UNROLL
      for (int j = 0; j < NZ; ++j) gm = fmaxf(gm, fabsf(g[j]));
      const float ha = fabsf(hh);
      const float n = (ha != ha) ? ha : fmaxf(gm, ha);
      const float rn = rcp_refined(n);
      RCBF_UNROLL
      for (int j = 0; j < NZ; ++j) Rg[r][j] = div_by(g[j], n, rn);
      rbg[r] = div_by(hh, n, rn);
    }
    RCBF_UNROLL
    for (int j = 0; j < NZ; ++j) R[r][j] = Rg[r][j] * pis[j];
    // (R R') lam = -rb on the first r+1 rows (r is a compile-time constant once the loop is unrolled)
    float lk[NZ];
    if (r == 0) greedy_solve_rows<NZ, 1>(R, rbg, lk);
    else if (r == 1 || NZ == 2) greedy_solve_rows<NZ, (NZ < 2 ? NZ : 2)>(R, rbg, lk);
    else greedy_solve_rows<NZ, NZ>(R, rbg, lk);
    RCBF_UNROLL
    for (int k = 0; k < NZ; ++k) ok = ok && (lk[k] >= -1e-6f);  // NaN (dependent rows) -> false
    if (!ok) break;
    RCBF_UNROLL
    for (int j = 0; j < NZ; ++j) {
      float acc = 0.f;
      RCBF_UNROLL
      for (int k = 0; k < NZ; ++k)
        if (k <= r) acc = fmaf(-R[k][j], lk[k], acc);  // (rows beyond r are zero and carry lam = 0)
      y[j] = acc;
    }
  }
  mask_out = mask;
  rounds = r;
  return ok && feasible;
}

// row i of the normalised problem in float64 (the exact slack of the certificate): h~_i - sum_j (G~_ij P_j^-1/2) y_j
template <typename Pat, int NZ, int M>
RCBF_HD double exact_slack_of_raw_row(const float g[NZ], float hh, const double pis[NZ], const double y[NZ]) {
  float gm = 0.f;
  RCBF_UNROLL
  for (int j = 0; j < NZ; ++j) gm = fmaxf(gm, fabsf(g[j]));
  const float ha = fabsf(hh);
  const float n = (ha != ha) ? ha : fmaxf(gm, ha);
  const float rn = rcp_refined(n);
  double acc = (double)div_by(hh, n, rn);
  RCBF_UNROLL
  for (int j = 0; j < NZ; ++j) acc = fma(-((double)div_by(g[j], n, rn) * pis[j]), y[j], acc);
  return acc;
}

template <typename Pat, int NZ, int M>
RCBF_HD bool lnp_certify_raw(const float G[M][NZ], const float h[M], const float A[M][NZ], const double pis[NZ],
                             const float Rg[NZ][NZ], const float rbg[NZ], int cnt, uint32_t mask, double tol_s,
                             double tol_l, double y[NZ], double lk[NZ]) {
  double R[NZ][NZ], Gm[NZ][NZ], nrb[NZ];
  RCBF_UNROLL
  for (int k = 0; k < NZ; ++k) {
    RCBF_UNROLL
    for (int j = 0; j < NZ; ++j) R[k][j] = (double)Rg[k][j] * pis[j];
    nrb[k] = -(double)rbg[k];
  }
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
  ch.solve(nrb, lk);
  bool ok = true;
  RCBF_UNROLL
  for (int k = 0; k < NZ; ++k) ok = ok && (lk[k] >= -tol_l);
  RCBF_UNROLL
  for (int j = 0; j < NZ; ++j) {
    double acc = 0.0;
    RCBF_UNROLL
    for (int k = 0; k < NZ; ++k) acc = fma(-R[k][j], lk[k], acc);
    y[j] = acc;
  }
  // rows of the active set (the gathered, normalised slots): |slack| <= tol_s in float64
  RCBF_UNROLL
  for (int k = 0; k < NZ; ++k) {
    double acc = (double)rbg[k];
    RCBF_UNROLL
    for (int j = 0; j < NZ; ++j) acc = fma(-R[k][j], y[j], acc);
    ok = ok && (k >= cnt || (acc >= -tol_s && acc <= tol_s));
  }
  // the other rows: float32 screen on the raw rows, float64 only inside the rounding-error band
  float yf[NZ], ya[NZ];
  RCBF_UNROLL
  for (int j = 0; j < NZ; ++j) {
    yf[j] = (float)y[j];
    ya[j] = fabsf(yf[j]);
  }
  uint32_t amb = 0;
  RCBF_UNROLL
  for (int i = 0; i < M; ++i) {
    float sl = h[i], mag = fabsf(h[i]);
    RCBF_UNROLL
    for (int j = 0; j < NZ; ++j)
      if (Pat::nz(i, j)) {
        sl = fmaf(-A[i][j], yf[j], sl);
        mag = fmaf(fabsf(A[i][j]), ya[j], mag);
      }
    const bool act = (mask >> i) & 1u;
    // |float32 error of sl| <= (NZ + 2) * 2^-24 * mag (+ the float rounding of y and A: 3 * 2^-24 * mag): band = 1e-5 mag
    if (!act && !(sl > 1e-5f * mag)) amb |= 1u << i;
  }
#if defined(__CUDA_ARCH__)
#pragma unroll 1
#endif
  for (; amb != 0u; amb &= amb - 1u) {  // (rare) one copy of the exact test: gather the row like the presolve does
#if defined(__CUDA_ARCH__)
    const int wi = __ffs((int)amb) - 1;
#else
    const int wi = __builtin_ffs((int)amb) - 1;
#endif
    float g[NZ], hh = 0.f;
    RCBF_UNROLL
    for (int j = 0; j < NZ; ++j) g[j] = 0.f;
    RCBF_UNROLL
    for (int i = 0; i < M; ++i) {
      const bool put = (i == wi);
      RCBF_UNROLL
      for (int j = 0; j < NZ; ++j)
        if (Pat::nz(i, j)) g[j] = put ? G[i][j] : g[j];
      hh = put ? h[i] : hh;
    }
    ok = ok && (exact_slack_of_raw_row<Pat, NZ, M>(g, hh, pis, y) >= -tol_s);
  }
  return ok;
}

// Fast path of one QP given its rows as assembled (presolve mode): greedy guess, float64 certificate, and -- for the
// callers that save them -- the dense multipliers / slacks of the normalised problem.
template <typename Pat, int NZ, int M, typename Fetch = SelectRowFetch<Pat, NZ, M>>
RCBF_HD void solve_raw_fast(const float G[M][NZ], const float h[M], const float p_diag[NZ], bool want_aux,
                            NormSolution<NZ, M>& o, const Fetch& fetch = Fetch()) {
  double pisd[NZ];
  float pisf[NZ];
  pis_of<NZ, M>(p_diag, pisd, pisf);
  double y[NZ], lk[NZ];
  uint32_t mask;
  int rounds;
  float A[M][NZ], Rg[NZ][NZ], rbg[NZ];
  const bool guess = lnp_greedy_raw<Pat, NZ, M, Fetch>(G, h, pisf, A, Rg, rbg, mask, rounds, fetch);
  o.status = RCBF_PENDING;
  o.iters = rounds;
  o.mask = mask;
  RCBF_UNROLL
  for (int j = 0; j < NZ; ++j) y[j] = 0.0;
  if (guess && lnp_certify_raw<Pat, NZ, M>(G, h, A, pisd, Rg, rbg, rounds, mask, kTolSlack, kTolDual, y, lk)) {
    o.status = RCBF_OK_CERTIFIED;
    if (want_aux) {
      Normalised<NZ, M> nrm;
      normalise_rows<Pat, NZ, M>(G, h, nrm);
      lnp_expand_aux<Pat, NZ, M>(nrm.Gn, nrm.hn, pisd, Rg, rbg, mask, y, lk, o.lam, o.s);
    }
  }
  if (M <= 4 && o.status == RCBF_PENDING) {
    // Few rows (SimulatedCars: 10 candidate active sets): when the greedy guess is not certified, enumerate right
    // here with the all-rows float64 certificate instead of queueing for pass 2 (about 1 instance in 1000).
    Normalised<NZ, M> nrm;
    normalise_rows<Pat, NZ, M>(G, h, nrm);
    const NormCert<NZ, M> cp{nrm, pisd};
    double lam[M], sl[M];
#ifdef __CUDA_ARCH__
#pragma unroll 1
#endif
    for (uint32_t m = 1; m < (1u << M); ++m) {
      int pc = 0;
      RCBF_UNROLL
      for (int i = 0; i < M; ++i) pc += (m >> i) & 1u;
      if (pc > NZ) continue;
      if (lnp_certify<double, NormCert<NZ, M>, Pat, NZ, M>(cp, m, kTolSlack, kTolDual, y, lam, sl)) {
        o.status = RCBF_OK_CERTIFIED;
        o.iters = NZ + 1;  // marks "enumerated" in the iteration histogram
        o.mask = m;
        RCBF_UNROLL
        for (int i = 0; i < M; ++i) {
          o.lam[i] = lam[i];
          o.s[i] = sl[i];
        }
        break;
      }
    }
  }
  RCBF_UNROLL
  for (int j = 0; j < NZ; ++j) o.x[j] = y[j] * pisd[j];
}

// NaN-PROPAGATING 3-input min / max (one FMNMX3.NAN on sm_100): the result is NaN iff any input is
RCBF_HD float nanmin3(float a, float b, float c) {
#if defined(__CUDA_ARCH__)
  float r;
  asm("min.NaN.f32 %0, %1, %2, %3;" : "=f"(r) : "f"(a), "f"(b), "f"(c));
  return r;
#else
  return (a != a || b != b || c != c) ? NAN : fminf(fminf(a, b), c);
#endif
}
RCBF_HD float nanmax3(float a, float b, float c) {
#if defined(__CUDA_ARCH__)
  float r;
  asm("max.NaN.f32 %0, %1, %2, %3;" : "=f"(r) : "f"(a), "f"(b), "f"(c));
  return r;
#else
  return (a != a || b != b || c != c) ? NAN : fmaxf(fmaxf(a, b), c);
#endif
}
// "is any of these K words NaN", K >= 1: a NaN-propagating max chain, (K + 1) / 2 instructions + one compare
template <int K>
RCBF_HD bool any_nan(const float* w) {
  float m = w[0];
  RCBF_UNROLL
  for (int k = 1; k + 1 < K; k += 2) m = nanmax3(m, w[k], w[k + 1]);
  if ((K & 1) == 0) m = nanmax3(m, w[K - 1], w[K - 1]);
  return m != m;
}

// trivial test on the RAW rows (n_i > 0, so h~_i >= 0 <=> h_i >= 0): x = 0 is optimal iff every h_i >= 0.
// One NaN-propagating min over the rows answers both questions: NaN -> some h_i is NaN, else trivial iff min >= 0.
template <int M>
RCBF_HD void classify_raw(const float h[M], bool& triv, bool& nan) {
  float m = h[0];
  RCBF_UNROLL
  for (int i = 1; i + 1 < M; i += 2) m = nanmin3(m, h[i], h[i + 1]);
  if ((M & 1) == 0) m = nanmin3(m, h[M - 1], h[M - 1]);
  nan = (m != m);
  triv = (m >= 0.f);
}

// Fallback pass for the (rare) instances the fast path left pending: float32 interior point + certificate, then the
// float64 interior point (which may also accept on its residual test, like qpth).  iters += 100 flags the f64 pass.
template <typename Pat, int NZ, int M>
RCBF_HD void solve_normalised_full(const Normalised<NZ, M>& nrm, const float p_diag[NZ], bool skip_f32,
                                   NormSolution<NZ, M>& o) {
  double pisd[NZ];
  float pisf[NZ];
  pis_of<NZ, M>(p_diag, pisd, pisf);
  LnpSolution<double, NZ, M> sol;
  sol.status = RCBF_MAXITER;
  sol.iters = 0;
  const NormCert<NZ, M> cp{nrm, pisd};
  if (!skip_f32) {
    LnpProblem<float, NZ, M> Pf;
    to_lnp<float, Pat, NZ, M>(nrm, pisf, Pf);
    lnp_solve<float, double, NormCert<NZ, M>, Pat, NZ, M>(Pf, cp, sol, kTolSlack, kTolDual);
  }
  if (sol.status >= RCBF_MAXITER) {
    const int it0 = sol.iters;
    LnpProblem<double, NZ, M> Pd;
    to_lnp<double, Pat, NZ, M>(nrm, pisd, Pd);
    lnp_solve<double, double, NormCert<NZ, M>, Pat, NZ, M>(Pd, cp, sol, kTolSlack, kTolDual);
    sol.iters += it0 + 100;
  }
  RCBF_UNROLL
  for (int j = 0; j < NZ; ++j) o.x[j] = sol.y[j] * pisd[j];
  RCBF_UNROLL
  for (int i = 0; i < M; ++i) {
    o.lam[i] = sol.lam[i];
    o.s[i] = sol.s[i];
  }
  o.status = sol.status;
  o.iters = sol.iters;
  o.mask = sol.mask;
}

// ------------------------------------------------------------------------------------------------
// get_safe_action for one instance (diff_cbf_qp.py:44-79): assemble -> normalise -> solve -> clamp
// ------------------------------------------------------------------------------------------------
// torch.clamp propagates NaN (fminf/fmaxf would return the bound)
RCBF_HD float clampf(float v, float lo, float hi) { return (v != v) ? v : fminf(fmaxf(v, lo), hi); }

struct UniSolve {
  UniRaw raw;
  Normalised<kUniNZ, kUniM> nrm;
  NormSolution<kUniNZ, kUniM> sol;
};

template <typename Pat, int NZ, int M>
RCBF_HD void trivial_solution(const Normalised<NZ, M>& nrm, bool nan, NormSolution<NZ, M>& o) {
  RCBF_UNROLL
  for (int j = 0; j < NZ; ++j) o.x[j] = nan ? (double)NAN : 0.0;
  RCBF_UNROLL
  for (int i = 0; i < M; ++i) {
    o.lam[i] = 0.0;
    o.s[i] = (double)nrm.hn[i];
  }
  o.status = nan ? RCBF_NAN : RCBF_OK_TRIVIAL;
  o.iters = 0;
  o.mask = 0u;
}

// kMode: 0 = fast path with presolve, 1 = fast path with inline float32 IPM ("pdipm" mode), 2 = fallback pass,
//        3 = fallback pass that skips the float32 IPM (it already failed in mode 1)
// (reference driver used by the host simulation and the backward kernels; the hot kernels in rcbf_safe_kernels.cuh
//  call the same pieces but interleave them with the compaction ring)
template <int kMode>
RCBF_HD void unicycle_safe_action(const UnicycleParams& p, const float st[3], const float u[2], const float mu[3],
                                  const float sg[3], UniSolve& w, float u_safe[2]) {
  assemble_unicycle(p, st, u, mu, sg, w.raw);
  normalise_rows<UniPat, kUniNZ, kUniM>(w.raw.G, w.raw.h, w.nrm);
  bool triv, nan;
  classify_raw<kUniM>(w.raw.h, triv, nan);
  RCBF_UNROLL
  for (int i = 0; i < kUniHaz; ++i) nan = nan || (w.raw.G[i][0] != w.raw.G[i][0]) || (w.raw.G[i][1] != w.raw.G[i][1]);
  if (triv || nan) {
    trivial_solution<UniPat, kUniNZ, kUniM>(w.nrm, nan, w.sol);
  } else {
    if (kMode == 0) solve_raw_fast<UniPat, kUniNZ, kUniM>(w.raw.G, w.raw.h, p.p_diag, true, w.sol);
    if (kMode == 1) solve_normalised_ipm<UniPat, kUniNZ, kUniM>(w.nrm, p.p_diag, w.sol);
    if (kMode >= 2) solve_normalised_full<UniPat, kUniNZ, kUniM>(w.nrm, p.p_diag, kMode == 3, w.sol);
  }
  RCBF_UNROLL
  for (int c = 0; c < 2; ++c) u_safe[c] = clampf(u[c] + (float)w.sol.x[c], p.u_min[c], p.u_max[c]);  // :77
}

struct CarsSolve {
  CarsRaw raw;
  Normalised<kCarsNZ, kCarsM> nrm;
  NormSolution<kCarsNZ, kCarsM> sol;
};

template <int kMode>
RCBF_HD void cars_safe_action(const CarsParams& p, const float st[10], float u, const float sg[10], CarsSolve& w,
                              float* u_safe) {
  assemble_cars(p, st, u, sg, w.raw);
  normalise_rows<CarsPat, kCarsNZ, kCarsM>(w.raw.G, w.raw.h, w.nrm);
  bool triv, nan;
  classify_raw<kCarsM>(w.raw.h, triv, nan);
  nan = nan || (w.raw.G[0][0] != w.raw.G[0][0]) || (w.raw.G[1][0] != w.raw.G[1][0]);
  if (triv || nan) {
    trivial_solution<CarsPat, kCarsNZ, kCarsM>(w.nrm, nan, w.sol);
  } else {
    if (kMode == 0) solve_raw_fast<CarsPat, kCarsNZ, kCarsM>(w.raw.G, w.raw.h, p.p_diag, true, w.sol);
    if (kMode == 1) solve_normalised_ipm<CarsPat, kCarsNZ, kCarsM>(w.nrm, p.p_diag, w.sol);
    if (kMode >= 2) solve_normalised_full<CarsPat, kCarsNZ, kCarsM>(w.nrm, p.p_diag, kMode == 3, w.sol);
  }
  *u_safe = clampf(u + (float)w.sol.x[0], p.u_min, p.u_max);  // :77
}

}  // namespace rcbf
