// rcbf_kernels.cu -- sm_100a kernels + the C ABI of include/rcbf_b200.h.
//
// Mapping: one CUDA thread = one environment instance / one QP; everything of an instance (27 G~ + 9 h~ + 18
// interior-point iterates + temporaries) lives in registers, the inner loops are fully unrolled (rcbf_core.cuh).
// The work is FP32-pipe bound (~1 kflop per interior-point iteration against 64-128 B of HBM traffic per
// instance), so there is no shared-memory tiling: loads are plain coalesced row-major reads issued up front, the
// Unicycle env state is a single 16-byte load/store.  Blocks of 128 threads; see DESIGN.md for the roofline.
#include <cuda_runtime.h>
#include <stdint.h>

#include <mutex>

#include "rcbf_backward.cuh"
#include "rcbf_core.cuh"
#include "rcbf_dynamics.cuh"
#include "rcbf_generic.cuh"
#include "rcbf_tma.cuh"

using namespace rcbf;

namespace {

constexpr int kThreads = 128;

inline int grid_for(int64_t n) { return (int)((n + kThreads - 1) / kThreads); }
[[maybe_unused]] inline int sm_count() {  // SMs of the current device (cached per device)
  static int cached[64] = {};
  int dev = 0;
  cudaGetDevice(&dev);
  int& c = cached[dev & 63];
  if (c == 0) cudaDeviceGetAttribute(&c, cudaDevAttrMultiProcessorCount, dev);
  return c > 0 ? c : 148;
}
#define RCBF_LAUNCH_CHECK()                 \
  do {                                      \
    cudaError_t e_ = cudaGetLastError();    \
    if (e_ != cudaSuccess) return (int)e_;  \
  } while (0)

template <int K>
__device__ __forceinline__ void load_row(const float* __restrict__ base, int64_t i, float out[K]) {
#pragma unroll
  for (int j = 0; j < K; ++j) out[j] = __ldg(base + i * K + j);
}
template <int K>
__device__ __forceinline__ void load_row(const double* __restrict__ base, int64_t i, double out[K]) {
#pragma unroll
  for (int j = 0; j < K; ++j) out[j] = __ldg(base + i * K + j);
}
template <int K, typename T>
__device__ __forceinline__ void store_row(T* __restrict__ base, int64_t i, const T in[K]) {
#pragma unroll
  for (int j = 0; j < K; ++j) base[i * K + j] = in[j];
}

// Rows of K words per instance, one instance per thread.  A lane-per-row access pattern makes each of a warp's K
// load / store instructions touch EVERY 32-byte sector of the warp's span (K-fold request amplification on the L1 -> L2
// path, the limiter of these streaming kernels on B200); staging the block's rows through shared memory makes every
// global access fully coalesced.  All threads of the block must call (rows beyond n read as zero / are not written).
template <typename T, int K>
struct BlockRows {
  T* buf;  // kThreads * K words of shared memory, reused by successive calls
  __device__ explicit BlockRows(T* b) : buf(b) {}
  __device__ __forceinline__ void load(const T* __restrict__ base, int64_t i0, int64_t n, T row[K]) {
    const int cnt = (int)((n - i0) < (int64_t)kThreads ? (n - i0) : (int64_t)kThreads);
    const T* __restrict__ src = base + i0 * K;
    __syncthreads();
    for (int w = threadIdx.x; w < cnt * K; w += kThreads) buf[w] = __ldg(src + w);
    __syncthreads();
#pragma unroll
    for (int k = 0; k < K; ++k) row[k] = ((int)threadIdx.x < cnt) ? buf[threadIdx.x * K + k] : T(0);
  }
  __device__ __forceinline__ void store(T* __restrict__ base, int64_t i0, int64_t n, const T row[K]) {
    const int cnt = (int)((n - i0) < (int64_t)kThreads ? (n - i0) : (int64_t)kThreads);
    T* __restrict__ dst = base + i0 * K;
    __syncthreads();
    if ((int)threadIdx.x < cnt) {
#pragma unroll
      for (int k = 0; k < K; ++k) buf[threadIdx.x * K + k] = row[k];
    }
    __syncthreads();
    for (int w = threadIdx.x; w < cnt * K; w += kThreads) dst[w] = buf[w];
  }
};

// block-level accumulation of the solver counters (rare events -> rare atomics)
__device__ __forceinline__ void accumulate_counters(rcbf_counters_t* counters, bool valid, int status, int iters) {
  if (counters == nullptr) return;
  const int n_nan = __syncthreads_count(valid && status == RCBF_NAN);
  const int n_max = __syncthreads_count(valid && status == RCBF_MAXITER);
  const int n_f64 = __syncthreads_count(valid && iters >= 100);
  const int n_triv = __syncthreads_count(valid && status == RCBF_OK_TRIVIAL);
  const int n_pend = __syncthreads_count(valid && status == RCBF_PENDING);
  __shared__ int s_it;
  if (threadIdx.x == 0) s_it = 0;
  __syncthreads();
  int it = valid ? (iters >= 100 ? iters - 100 : iters) : 0;
  it = __reduce_add_sync(0xffffffffu, it);
  if ((threadIdx.x & 31) == 0 && it) atomicAdd(&s_it, it);
  __syncthreads();
  if (threadIdx.x == 0) {
    if (n_nan) atomicAdd(&counters[0], (unsigned long long)n_nan);
    if (n_max) atomicAdd(&counters[1], (unsigned long long)n_max);
    if (n_f64) atomicAdd(&counters[2], (unsigned long long)n_f64);
    if (n_triv) atomicAdd(&counters[3], (unsigned long long)n_triv);
    if (s_it) atomicAdd(&counters[4], (unsigned long long)s_it);
    if (n_pend) atomicAdd(&counters[5], (unsigned long long)n_pend);
  }
}


// ------------------------------------------------------------------------------------------------------------
// K2: constraint assembly (raw G, h of get_cbf_qp_constraints)
// ------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kThreads) k_unicycle_assemble(const float* __restrict__ st, const float* __restrict__ ac,
                                                                const float* __restrict__ mu, const float* __restrict__ sg,
                                                                int64_t n, UnicycleParams p, float* __restrict__ G,
                                                                float* __restrict__ h) {
  const int64_t i = (int64_t)blockIdx.x * kThreads + threadIdx.x;
  if (i >= n) return;
  float s[3], u[2], m[3], g[3];
  load_row<3>(st, i, s);
  load_row<2>(ac, i, u);
  load_row<3>(mu, i, m);
  load_row<3>(sg, i, g);
  UniRaw raw;
  assemble_unicycle(p, s, u, m, g, raw);
#pragma unroll
  for (int r = 0; r < kUniM; ++r) {
#pragma unroll
    for (int j = 0; j < kUniNZ; ++j) G[(i * kUniM + r) * kUniNZ + j] = raw.G[r][j];
    h[i * kUniM + r] = raw.h[r];
  }
}

__global__ void __launch_bounds__(kThreads) k_cars_assemble(const float* __restrict__ st, const float* __restrict__ ac,
                                                            const float* __restrict__ sg, int64_t n, CarsParams p,
                                                            float* __restrict__ G, float* __restrict__ h) {
  const int64_t i = (int64_t)blockIdx.x * kThreads + threadIdx.x;
  if (i >= n) return;
  float s[10], g[10];
  load_row<10>(st, i, s);
  load_row<10>(sg, i, g);
  CarsRaw raw;
  assemble_cars(p, s, __ldg(ac + i), g, raw);
#pragma unroll
  for (int r = 0; r < kCarsM; ++r) {
#pragma unroll
    for (int j = 0; j < kCarsNZ; ++j) G[(i * kCarsM + r) * kCarsNZ + j] = raw.G[r][j];
    h[i * kCarsM + r] = raw.h[r];
  }
}

// float64 assembly (+ optional row normalisation) -- the single-instance numpy layer CascadeCBFLayer works in float64
// (rcbf_sac/cbf_qp.py:84-240, normalisation :262-265); its QP then goes through the generic float64 kernel (rcbf_qp_solve).
// Same formulas as assemble_unicycle_v / assemble_cars, evaluated in double from double inputs (the float parameters of
// the structs are exact for every reference constant except l_p = 0.03 and r_c^2, which carry a 1e-8 relative rounding).
__device__ __forceinline__ void normalise_row_f64(double* g, int nz, double& h) {
  double n = fabs(h);
  for (int j = 0; j < nz; ++j) n = fmax(n, fabs(g[j]));
  for (int j = 0; j < nz; ++j) g[j] /= n;
  h /= n;
}

__global__ void __launch_bounds__(kThreads)
k_unicycle_assemble_f64(const double* __restrict__ st, const double* __restrict__ ac, const double* __restrict__ mu,
                        const double* __restrict__ sg, int64_t n, UnicycleParams p, int normalise,
                        double* __restrict__ G, double* __restrict__ h) {
  const int64_t i = (int64_t)blockIdx.x * kThreads + threadIdx.x;
  if (i >= n) return;
  double s[3], u[2], m[3], g[3];
  load_row<3>(st, i, s);
  load_row<2>(ac, i, u);
  load_row<3>(mu, i, m);
  load_row<3>(sg, i, g);
  const double lp = (double)p.l_p, c = cos(s[2]), sn = sin(s[2]);
  const double px = s[0] + lp * c, py = s[1] + lp * sn;                       // cbf_qp.py:94
  const double g01 = -lp * sn, g11 = lp * c;                                  // g_p = R diag(1, l_p)   :100-105
  const double mpx = m[0] + g01 * m[2], mpy = m[1] + g11 * m[2];              // :117
  const double a01 = p.abs_sigma_map ? fabs(g01) : g01, a11 = p.abs_sigma_map ? fabs(g11) : g11;
  const double spx = (double)p.sigma_scale * (g[0] + a01 * g[2]);             // :119 (signed), k_d of :141
  const double spy = (double)p.sigma_scale * (g[1] + a11 * g[2]);
  double* Gi = G + i * kUniM * kUniNZ;
  double* hi = h + i * kUniM;
#pragma unroll
  for (int k = 0; k < kUniHaz; ++k) {
    const double dx = px - (double)p.hazards[k][0], dy = py - (double)p.hazards[k][1];   // :111
    const double hc = 0.5 * (dx * dx + dy * dy - (double)p.collision_radius_sq);          // :108
    const double L0 = dx * c + dy * sn, L1 = dx * g01 + dy * g11;
    double row[3] = {-L0, -L1, -1.0};                                          // :136-137
    double hh = (double)p.gamma_b * hc * hc * hc + (dx * mpx + dy * mpy) + (L0 * u[0] + L1 * u[1]) -
                (fabs(dx) * spx + fabs(dy) * spy);                             // :138-141
    if (normalise) normalise_row_f64(row, 3, hh);
    Gi[3 * k] = row[0]; Gi[3 * k + 1] = row[1]; Gi[3 * k + 2] = row[2];
    hi[k] = hh;
  }
#pragma unroll
  for (int cc = 0; cc < 2; ++cc) {                                             // :226-238
    double r0[3] = {cc == 0 ? 1.0 : 0.0, cc == 1 ? 1.0 : 0.0, 0.0}, h0 = (double)p.u_max[cc] - u[cc];
    double r1[3] = {cc == 0 ? -1.0 : 0.0, cc == 1 ? -1.0 : 0.0, 0.0}, h1 = -(double)p.u_min[cc] + u[cc];
    if (normalise) {
      normalise_row_f64(r0, 3, h0);
      normalise_row_f64(r1, 3, h1);
    }
    const int r = kUniHaz + 2 * cc;
    for (int j = 0; j < 3; ++j) {
      Gi[3 * r + j] = r0[j];
      Gi[3 * (r + 1) + j] = r1[j];
    }
    hi[r] = h0;
    hi[r + 1] = h1;
  }
}

__global__ void __launch_bounds__(kThreads)
k_cars_assemble_f64(const double* __restrict__ st, const double* __restrict__ ac, const double* __restrict__ sg, int64_t n,
                    CarsParams p, int normalise, double* __restrict__ G, double* __restrict__ h) {
  const int64_t i = (int64_t)blockIdx.x * kThreads + threadIdx.x;
  if (i >= n) return;
  double s[10], g[10];
  load_row<10>(st, i, s);
  load_row<10>(sg, i, g);
  const double u = __ldg(ac + i);
  double pos[5], vel[5], acc[5];
#pragma unroll
  for (int k = 0; k < 5; ++k) {
    pos[k] = s[2 * k];
    vel[k] = s[2 * k + 1];
    acc[k] = (double)p.kp * (30.0 - vel[k]);                                   // cbf_qp.py:162-163
  }
  const double d01 = pos[0] - pos[1], d12 = pos[1] - pos[2], d24 = pos[2] - pos[4];
  acc[1] -= (d01 < 6.0) ? (double)p.k_brake * d01 : 0.0;                        // :164-167
  acc[2] -= (d12 < 6.0) ? (double)p.k_brake * d12 : 0.0;
  acc[4] -= (d24 < 13.0) ? (double)p.k_brake * d24 : 0.0;
  const double d23 = pos[2] - pos[3], d43 = pos[4] - pos[3];
  const double r2 = (double)p.collision_radius_sq;
  const double h13 = 0.5 * (d23 * d23 - r2), h15 = 0.5 * (d43 * d43 - r2);    // :178-179
  const double a7 = pos[3] - pos[2], b7 = pos[3] - pos[4], a6 = vel[3] - vel[2], b6 = vel[3] - vel[4];
  const double h13d = a7 * a6, h15d = b7 * b6;                                 // :182-183
  const double Lff13 = (vel[2] - vel[3]) * vel[2] + d23 * acc[2] + a6 * vel[3];        // :186-191 (acc[3] = 0)
  const double Lff15 = b6 * vel[3] + (vel[4] - vel[3]) * vel[4] + d43 * acc[4];        // :194-199
  const double ss = (double)p.sigma_scale;
  const double LfD13 = ss * (fabs(d23) * g[5] + fabs(a7) * g[7]);              // (0 in the cascade layer: :210-211)
  const double LfD15 = ss * (fabs(b7) * g[7] + fabs(d43) * g[9]);
  const double Lg13 = a7 * 50.0, Lg15 = b7 * 50.0;                             // :202-203
  double rows[4][2] = {{-Lg13, -(double)p.slack_coeff}, {-Lg15, -(double)p.slack_coeff}, {1.0, 0.0}, {-1.0, 0.0}};
  double hh[4] = {Lff13 - LfD13 + (double)p.gamma_2 * h13d + (double)p.gamma_sq * h13 + Lg13 * u,
                  Lff15 - LfD15 + (double)p.gamma_2 * h15d + (double)p.gamma_sq * h15 + Lg15 * u,
                  (double)p.u_max - u, -(double)p.u_min + u};
#pragma unroll
  for (int r = 0; r < 4; ++r) {
    if (normalise) normalise_row_f64(rows[r], 2, hh[r]);
    G[(i * 4 + r) * 2] = rows[r][0];
    G[(i * 4 + r) * 2 + 1] = rows[r][1];
    h[i * 4 + r] = hh[r];
  }
}

// ------------------------------------------------------------------------------------------------------------
// K4: backward (recomputes the cheap assembly, reads the saved x / lam / slack)
// ------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kThreads)
k_unicycle_safe_action_bwd(const float* __restrict__ st, const float* __restrict__ ac, const float* __restrict__ mu,
                           const float* __restrict__ sg, const float* __restrict__ x, const float* __restrict__ lam,
                           const float* __restrict__ slack, const float* __restrict__ gout, int64_t n, UnicycleParams p,
                           float* __restrict__ grad_a) {
  const int64_t i = (int64_t)blockIdx.x * kThreads + threadIdx.x;
  if (i >= n) return;
  float s[3], u[2], m[3], g[3], xs[3], ls[kUniM], ss[kUniM], go[2];
  load_row<3>(st, i, s);
  load_row<2>(ac, i, u);
  load_row<3>(mu, i, m);
  load_row<3>(sg, i, g);
  load_row<3>(x, i, xs);
  load_row<kUniM>(lam, i, ls);
  load_row<kUniM>(slack, i, ss);
  load_row<2>(gout, i, go);
  UniRaw raw;
  assemble_unicycle(p, s, u, m, g, raw);
  Normalised<kUniNZ, kUniM> nrm;
  normalise_rows<UniPat, kUniNZ, kUniM>(raw.G, raw.h, nrm);
  float r[kUniM][2];
#pragma unroll
  for (int k = 0; k < kUniHaz; ++k) {
    r[k][0] = raw.Lg[k][0];
    r[k][1] = raw.Lg[k][1];
  }
#pragma unroll
  for (int c = 0; c < 2; ++c) {  // h = u_max - a_c ; h = -u_min + a_c
    r[kUniHaz + 2 * c][0] = (c == 0) ? -1.f : 0.f;
    r[kUniHaz + 2 * c][1] = (c == 1) ? -1.f : 0.f;
    r[kUniHaz + 2 * c + 1][0] = (c == 0) ? 1.f : 0.f;
    r[kUniHaz + 2 * c + 1][1] = (c == 1) ? 1.f : 0.f;
  }
  float ga[2];
  safe_action_bwd<kUniNZ, kUniM, 2>(nrm, raw.G, raw.h, r, p.p_diag, xs, ls, ss, u, p.u_min, p.u_max, go, ga);
  store_row<2>(grad_a, i, ga);
}

__global__ void __launch_bounds__(kThreads)
k_cars_safe_action_bwd(const float* __restrict__ st, const float* __restrict__ ac, const float* __restrict__ sg,
                       const float* __restrict__ x, const float* __restrict__ lam, const float* __restrict__ slack,
                       const float* __restrict__ gout, int64_t n, CarsParams p, float* __restrict__ grad_a) {
  const int64_t i = (int64_t)blockIdx.x * kThreads + threadIdx.x;
  if (i >= n) return;
  float s[10], g[10], xs[2], ls[kCarsM], ss[kCarsM];
  load_row<10>(st, i, s);
  load_row<10>(sg, i, g);
  load_row<2>(x, i, xs);
  load_row<kCarsM>(lam, i, ls);
  load_row<kCarsM>(slack, i, ss);
  const float u = __ldg(ac + i);
  const float go = __ldg(gout + i);
  CarsRaw raw;
  assemble_cars(p, s, u, g, raw);
  Normalised<kCarsNZ, kCarsM> nrm;
  normalise_rows<CarsPat, kCarsNZ, kCarsM>(raw.G, raw.h, nrm);
  float r[kCarsM][1] = {{raw.Lg[0]}, {raw.Lg[1]}, {-1.f}, {1.f}};
  float ga[1];
  const float uu[1] = {u}, lo[1] = {p.u_min}, hi[1] = {p.u_max}, gg[1] = {go};
  safe_action_bwd<kCarsNZ, kCarsM, 1>(nrm, raw.G, raw.h, r, p.p_diag, xs, ls, ss, uu, lo, hi, gg, ga);
  grad_a[i] = ga[0];
}

// ------------------------------------------------------------------------------------------------------------
// K4, compact form: the forward saved ONE word per instance, (status << 16) | active set.  A trivial instance
// (x = 0, no active row: ~2/3 of them) needs nothing but the clamp mask of diff_cbf_qp.py:77 -- 28 bytes of traffic,
// a handful of instructions.  The others are compacted inside the block (so that the second phase runs with full
// warps), their constraint rows re-assembled, and the gradient follows from the <= NZ ACTIVE rows alone through the
// exact active-set form of the implicit-KKT backward (safe_action_bwd_active, rcbf_backward.cuh).
// An instance the forward could not pin to a vertex (interior-point result accepted on its residual: status
// RCBF_OK_IPM / RCBF_MAXITER, mask 0xffff -- not observed on any test distribution) re-runs the fallback solve and the
// dense qpth-clamp backward in a separate, non-inlined function, so that its ~250 live values do not set the register
// allocation of the path every other instance takes.
// ------------------------------------------------------------------------------------------------------------
constexpr int kBwdThreads = 128;

template <typename Pat, int NZ, int M>
__device__ __forceinline__ void rebuild_saved(const Normalised<NZ, M>& nrm, const float p_diag[NZ], float xs[NZ],
                                              float ls[M], float ss[M]) {
  NormSolution<NZ, M> sol;
  solve_normalised_full<Pat, NZ, M>(nrm, p_diag, false, sol);
#pragma unroll
  for (int j = 0; j < NZ; ++j) xs[j] = (float)sol.x[j];
#pragma unroll
  for (int i = 0; i < M; ++i) {
    ls[i] = (float)sol.lam[i];
    ss[i] = (float)sol.s[i];
  }
}

// phase 1 of both kernels: trivial / NaN instances are finished, the others are listed (dense) in shared memory;
// P^-1/2 in float64 (two IEEE double operations per entry) is computed once per block
template <int NU, int NZ>
__device__ __forceinline__ int bwd_classify(const float* __restrict__ ac, const int32_t* __restrict__ meta,
                                            const float* __restrict__ gout, int64_t n, const float* u_min,
                                            const float* u_max, const float* p_diag, float* __restrict__ grad_a,
                                            int* list, double* s_pis) {
  __shared__ int s_count;
  if (threadIdx.x == 0) s_count = 0;
  if (threadIdx.x >= 32 && threadIdx.x < 32 + NZ) s_pis[threadIdx.x - 32] = 1.0 / sqrt((double)p_diag[threadIdx.x - 32]);
  __syncthreads();
  const int64_t i = (int64_t)blockIdx.x * kBwdThreads + threadIdx.x;
  bool heavy = false;
  if (i < n) {
    const int status = __ldg(meta + i) >> 16;
    if (status == RCBF_OK_TRIVIAL || status == RCBF_NAN) {
      float v[NU], g[NU];
      if (NU == 2) {
        const float2 a2 = __ldg(reinterpret_cast<const float2*>(ac) + i), g2 = __ldg(reinterpret_cast<const float2*>(gout) + i);
        v[0] = a2.x; v[NU - 1] = a2.y; g[0] = g2.x; g[NU - 1] = g2.y;
      } else {
#pragma unroll
        for (int c = 0; c < NU; ++c) { v[c] = __ldg(ac + i * NU + c); g[c] = __ldg(gout + i * NU + c); }
      }
#pragma unroll
      for (int c = 0; c < NU; ++c) {
        const float vv = v[c] + 0.f;
        g[c] = (status == RCBF_NAN) ? NAN : ((vv >= u_min[c] && vv <= u_max[c]) ? g[c] : 0.f);
      }
      if (NU == 2) reinterpret_cast<float2*>(grad_a)[i] = make_float2(g[0], g[NU - 1]);
      else grad_a[i] = g[0];
    } else {
      heavy = true;
    }
  }
  const unsigned b = __ballot_sync(0xffffffffu, heavy);
  int base = 0;
  if ((threadIdx.x & 31) == 0 && b) base = atomicAdd(&s_count, __popc(b));
  base = __shfl_sync(0xffffffffu, base, 0);
  if (heavy) list[base + __popc(b & ((1u << (threadIdx.x & 31)) - 1u))] = threadIdx.x;
  __syncthreads();
  return s_count;
}

// per-environment pieces of the compact backward: row widths, the lean (active-set) gradient of one instance from its
// input rows (global or shared memory) and the out-of-line dense fallback from global memory
#ifndef RCBF_BWD_UNI_TILE
#define RCBF_BWD_UNI_TILE 512
#endif
#ifndef RCBF_BWD_UNI_MINB
#define RCBF_BWD_UNI_MINB 7
#endif
#ifndef RCBF_BWD_CARS_TILE
#define RCBF_BWD_CARS_TILE 256   // (A/B: 512 x 4 blocks 0.0650 ms, 256 x 7 0.0636, 256 x 8 0.0641, 128 x 8 0.0718 per 4 Mi:
#endif                           //  96 bytes per instance, i.e. 6.3 TB/s -- this one runs at the HBM roofline)
#ifndef RCBF_BWD_CARS_MINB
#define RCBF_BWD_CARS_MINB 7
#endif
struct UniBwd {
  using Params = UnicycleParams;
  static constexpr int NU = 2, NZ = kUniNZ, SF = 3, MF = 3;  // action / state / mean row widths (sigma row = state row)
  // tile kernel: instances per block and resident blocks per SM (A/B on B200, 4 Mi instances: 512 x 4 / 5 / 6 / 7 blocks
  // = 0.0862 / 0.0782 / 0.0739 / 0.0700 ms, 256 x 8 / 10 = 0.0745 / 0.0761 -- the lean path fits 72 registers, and more
  // blocks in flight hide both the tile fetch and the phase barriers)
  static constexpr int kTile = RCBF_BWD_UNI_TILE, kMinBlocks = RCBF_BWD_UNI_MINB;
  static __device__ __forceinline__ const float* u_min(const Params& p) { return p.u_min; }
  static __device__ __forceinline__ const float* u_max(const Params& p) { return p.u_max; }
  static __device__ __forceinline__ void lean(const float* __restrict__ sr, const float* __restrict__ ur,
                                              const float* __restrict__ mr, const float* __restrict__ gr,
                                              const float* __restrict__ gor, uint32_t mask, const Params& p,
                                              const double* pis_s, float* ga) {
    const float s[3] = {sr[0], sr[1], sr[2]}, u[2] = {ur[0], ur[1]}, m[3] = {mr[0], mr[1], mr[2]};
    const float g[3] = {gr[0], gr[1], gr[2]}, go[2] = {gor[0], gor[1]};
    UniRaw raw;
    assemble_unicycle(p, s, u, m, g, raw);
    float r[kUniM][2];
#pragma unroll
    for (int k = 0; k < kUniHaz; ++k) {
      r[k][0] = raw.Lg[k][0];
      r[k][1] = raw.Lg[k][1];
    }
#pragma unroll
    for (int c = 0; c < 2; ++c) {  // h = u_max - a_c ; h = -u_min + a_c
      r[kUniHaz + 2 * c][0] = (c == 0) ? -1.f : 0.f;
      r[kUniHaz + 2 * c][1] = (c == 1) ? -1.f : 0.f;
      r[kUniHaz + 2 * c + 1][0] = (c == 0) ? 1.f : 0.f;
      r[kUniHaz + 2 * c + 1][1] = (c == 1) ? 1.f : 0.f;
    }
    const double pis[kUniNZ] = {pis_s[0], pis_s[1], pis_s[2]};
    safe_action_bwd_active<UniPat, kUniNZ, kUniM, 2>(raw.G, raw.h, r, pis, mask, u, p.u_min, p.u_max, go, ga);
  }
  static __device__ __noinline__ void dense(const float* __restrict__ st, const float* __restrict__ ac,
                                            const float* __restrict__ mu, const float* __restrict__ sg,
                                            const float* __restrict__ gout, int64_t i, const Params* pp, float* ga) {
    const Params& p = *pp;
    float s[3], u[2], m[3], g[3], xs[3], ls[kUniM], ss[kUniM], go[2];
    load_row<3>(st, i, s);
    load_row<2>(ac, i, u);
    load_row<3>(mu, i, m);
    load_row<3>(sg, i, g);
    load_row<2>(gout, i, go);
    UniRaw raw;
    assemble_unicycle(p, s, u, m, g, raw);
    Normalised<kUniNZ, kUniM> nrm;
    normalise_rows<UniPat, kUniNZ, kUniM>(raw.G, raw.h, nrm);
    rebuild_saved<UniPat, kUniNZ, kUniM>(nrm, p.p_diag, xs, ls, ss);
    float r[kUniM][2];
#pragma unroll
    for (int k = 0; k < kUniHaz; ++k) {
      r[k][0] = raw.Lg[k][0];
      r[k][1] = raw.Lg[k][1];
    }
#pragma unroll
    for (int c = 0; c < 2; ++c) {
      r[kUniHaz + 2 * c][0] = (c == 0) ? -1.f : 0.f;
      r[kUniHaz + 2 * c][1] = (c == 1) ? -1.f : 0.f;
      r[kUniHaz + 2 * c + 1][0] = (c == 0) ? 1.f : 0.f;
      r[kUniHaz + 2 * c + 1][1] = (c == 1) ? 1.f : 0.f;
    }
    safe_action_bwd<kUniNZ, kUniM, 2>(nrm, raw.G, raw.h, r, p.p_diag, xs, ls, ss, u, p.u_min, p.u_max, go, ga);
  }
};

struct CarsBwd {
  using Params = CarsParams;
  static constexpr int NU = 1, NZ = kCarsNZ, SF = 10, MF = 0;  // (the layer does not read the disturbance mean, :299)
  static constexpr int kTile = RCBF_BWD_CARS_TILE, kMinBlocks = RCBF_BWD_CARS_MINB;
  static __device__ __forceinline__ const float* u_min(const Params& p) { return &p.u_min; }
  static __device__ __forceinline__ const float* u_max(const Params& p) { return &p.u_max; }
  static __device__ __forceinline__ void lean(const float* __restrict__ sr, const float* __restrict__ ur,
                                              const float* __restrict__, const float* __restrict__ gr,
                                              const float* __restrict__ gor, uint32_t mask, const Params& p,
                                              const double* pis_s, float* ga) {
    float s[10], g[10];
#pragma unroll
    for (int j = 0; j < 10; ++j) {
      s[j] = sr[j];
      g[j] = gr[j];
    }
    const float u = ur[0];
    CarsRaw raw;
    assemble_cars(p, s, u, g, raw);
    const float r[kCarsM][1] = {{raw.Lg[0]}, {raw.Lg[1]}, {-1.f}, {1.f}};
    const double pis[kCarsNZ] = {pis_s[0], pis_s[1]};
    const float uu[1] = {u}, gg[1] = {gor[0]}, lo[1] = {p.u_min}, hi[1] = {p.u_max};
    safe_action_bwd_active<CarsPat, kCarsNZ, kCarsM, 1>(raw.G, raw.h, r, pis, mask, uu, lo, hi, gg, ga);
  }
  static __device__ __noinline__ void dense(const float* __restrict__ st, const float* __restrict__ ac,
                                            const float* __restrict__, const float* __restrict__ sg,
                                            const float* __restrict__ gout, int64_t i, const Params* pp, float* ga) {
    const Params& p = *pp;
    float s[10], g[10], xs[2], ls[kCarsM], ss[kCarsM];
    load_row<10>(st, i, s);
    load_row<10>(sg, i, g);
    const float u = __ldg(ac + i);
    const float go = __ldg(gout + i);
    CarsRaw raw;
    assemble_cars(p, s, u, g, raw);
    Normalised<kCarsNZ, kCarsM> nrm;
    normalise_rows<CarsPat, kCarsNZ, kCarsM>(raw.G, raw.h, nrm);
    rebuild_saved<CarsPat, kCarsNZ, kCarsM>(nrm, p.p_diag, xs, ls, ss);
    const float r[kCarsM][1] = {{raw.Lg[0]}, {raw.Lg[1]}, {-1.f}, {1.f}};
    const float uu[1] = {u}, gg[1] = {go}, lo[1] = {p.u_min}, hi[1] = {p.u_max};
    safe_action_bwd<kCarsNZ, kCarsM, 1>(nrm, raw.G, raw.h, r, p.p_diag, xs, ls, ss, uu, lo, hi, gg, ga);
  }
};

__device__ __forceinline__ bool bwd_is_vertex(int mt) {
  return (mt >> 16) == RCBF_OK_CERTIFIED && (mt & 0xffff) != (int)kMaskUnknown;
}

// generic form (any alignment, any n): one instance per thread, rows read from global memory
template <class B>
__global__ void __launch_bounds__(kBwdThreads, 4)
k_safe_action_bwd_meta(const float* __restrict__ st, const float* __restrict__ ac, const float* __restrict__ mu,
                       const float* __restrict__ sg, const int32_t* __restrict__ meta, const float* __restrict__ gout,
                       int64_t n, const __grid_constant__ typename B::Params p, float* __restrict__ grad_a) {
  __shared__ int list[kBwdThreads];
  __shared__ double s_pis[B::NZ];
  const int count = bwd_classify<B::NU, B::NZ>(ac, meta, gout, n, B::u_min(p), B::u_max(p), p.p_diag, grad_a, list, s_pis);
  if ((int)threadIdx.x >= count) return;
  const int64_t i = (int64_t)blockIdx.x * kBwdThreads + list[threadIdx.x];
  const int mt = __ldg(meta + i);
  float ga[B::NU];
  if (bwd_is_vertex(mt))
    B::lean(st + i * B::SF, ac + i * B::NU, mu + i * B::MF, sg + i * B::SF, gout + i * B::NU, (uint32_t)mt & 0xffffu, p, s_pis, ga);
  else
    B::dense(st, ac, mu, sg, gout, i, &p, ga);
  store_row<B::NU>(grad_a, i, ga);
}

// tile form (every array base 16-byte aligned; full tiles of B::kTile instances): the rows of a tile are contiguous
// spans, so ONE thread fetches all of them with six cp.async.bulk copies (TMA, mbarrier completion) and nobody issues
// a global load; several resident blocks per SM overlap one tile's fetch with another tile's arithmetic.  Phase 1
// finishes the trivial instances in shared memory and lists the others; phase 2 runs the listed instances in chunks of
// 32 per warp from their shared-memory rows; the gradient tile (written in place of grad_out's) leaves by coalesced
// 128-bit stores.

template <class B>
struct alignas(16) BwdTileSmem {
  uint64_t bar;
  int count;
  int pad;
  double pis[4];
  int meta[B::kTile];
  int list[B::kTile];
  float ac[B::kTile * B::NU];
  float go[B::kTile * B::NU];
  float st[B::kTile * B::SF];
  float sg[B::kTile * B::SF];
  float mu[B::kTile * (B::MF ? B::MF : 1)];
};

template <class B>
__global__ void __launch_bounds__(kBwdThreads, B::kMinBlocks)
k_safe_action_bwd_tile(const float* __restrict__ st, const float* __restrict__ ac, const float* __restrict__ mu,
                       const float* __restrict__ sg, const int32_t* __restrict__ meta, const float* __restrict__ gout,
                       const __grid_constant__ typename B::Params p, float* __restrict__ grad_a) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  BwdTileSmem<B>& sh = *reinterpret_cast<BwdTileSmem<B>*>(smem_raw);
  constexpr int NU = B::NU;
  const int tid = threadIdx.x, lane = tid & 31;
  const int64_t i0 = (int64_t)blockIdx.x * B::kTile;
  if (tid == 0) {
    mbar_init(&sh.bar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    constexpr uint32_t kBytes = B::kTile * 4u * (1 + 2 * NU + 2 * B::SF + B::MF);
    mbar_expect_tx(&sh.bar, kBytes);
    bulk_g2s(sh.meta, meta + i0, B::kTile * 4, &sh.bar);
    bulk_g2s(sh.ac, ac + i0 * NU, B::kTile * 4 * NU, &sh.bar);
    bulk_g2s(sh.go, gout + i0 * NU, B::kTile * 4 * NU, &sh.bar);
    bulk_g2s(sh.st, st + i0 * B::SF, B::kTile * 4 * B::SF, &sh.bar);
    bulk_g2s(sh.sg, sg + i0 * B::SF, B::kTile * 4 * B::SF, &sh.bar);
    if (B::MF) bulk_g2s(sh.mu, mu + i0 * B::MF, B::kTile * 4 * B::MF, &sh.bar);
    sh.count = 0;
  }
  if (tid >= 32 && tid < 32 + B::NZ) sh.pis[tid - 32] = 1.0 / sqrt((double)p.p_diag[tid - 32]);
  __syncthreads();
  mbar_wait(&sh.bar, 0);
  const float* lo = B::u_min(p);
  const float* hi = B::u_max(p);
  // phase 1
#pragma unroll
  for (int q = 0; q < B::kTile / kBwdThreads; ++q) {
    const int idx = q * kBwdThreads + tid;
    const int status = sh.meta[idx] >> 16;
    const bool heavy = !(status == RCBF_OK_TRIVIAL || status == RCBF_NAN);
    if (!heavy) {
#pragma unroll
      for (int c = 0; c < NU; ++c) {
        const float vv = sh.ac[idx * NU + c] + 0.f;
        const float g = sh.go[idx * NU + c];
        sh.go[idx * NU + c] = (status == RCBF_NAN) ? NAN : ((vv >= lo[c] && vv <= hi[c]) ? g : 0.f);
      }
    }
    const unsigned b = __ballot_sync(0xffffffffu, heavy);
    int base = 0;
    if (lane == 0 && b) base = atomicAdd(&sh.count, __popc(b));
    base = __shfl_sync(0xffffffffu, base, 0);
    if (heavy) sh.list[base + __popc(b & ((1u << lane) - 1u))] = idx;
  }
  __syncthreads();
  // phase 2
  const int count = sh.count;
  for (int k = tid; k < count; k += kBwdThreads) {
    const int idx = sh.list[k];
    const int mt = sh.meta[idx];
    float ga[NU];
    if (bwd_is_vertex(mt))
      B::lean(sh.st + idx * B::SF, sh.ac + idx * NU, sh.mu + idx * B::MF, sh.sg + idx * B::SF, sh.go + idx * NU,
              (uint32_t)mt & 0xffffu, p, sh.pis, ga);
    else
      B::dense(st, ac, mu, sg, gout, i0 + idx, &p, ga);
#pragma unroll
    for (int c = 0; c < NU; ++c) sh.go[idx * NU + c] = ga[c];
  }
  __syncthreads();
  float4* dst = reinterpret_cast<float4*>(grad_a + i0 * NU);
  const float4* src = reinterpret_cast<const float4*>(sh.go);
#pragma unroll
  for (int w = tid; w < B::kTile * NU / 4; w += kBwdThreads) dst[w] = src[w];
}

template <class B>
int launch_bwd_meta(const float* st, const float* ac, const float* mu, const float* sg, const int32_t* meta,
                    const float* gout, int64_t n, const typename B::Params& p, float* grad_a, cudaStream_t s) {
  auto ok16 = [](const void* q) { return (reinterpret_cast<uintptr_t>(q) & 15) == 0; };
  int64_t done = 0;
  if (n >= 4 * B::kTile && ok16(st) && ok16(ac) && (B::MF == 0 || ok16(mu)) && ok16(sg) && ok16(meta) && ok16(gout) &&
      ok16(grad_a)) {
    int dev = 0;
    cudaGetDevice(&dev);
    static bool attr_set[64] = {};  // the opt-in to > 48 KB of dynamic shared memory is per function AND per device
    if (!attr_set[dev & 63]) {
      cudaFuncSetAttribute(k_safe_action_bwd_tile<B>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(BwdTileSmem<B>));
      attr_set[dev & 63] = true;
    }
    const int64_t ntiles = n / B::kTile;
    k_safe_action_bwd_tile<B><<<(unsigned)ntiles, kBwdThreads, sizeof(BwdTileSmem<B>), s>>>(st, ac, mu, sg, meta, gout, p, grad_a);
    done = ntiles * B::kTile;
  }
  if (done < n) {
    const int64_t m = n - done;
    k_safe_action_bwd_meta<B><<<(unsigned)((m + kBwdThreads - 1) / kBwdThreads), kBwdThreads, 0, s>>>(
        st + done * B::SF, ac + done * B::NU, B::MF ? mu + done * B::MF : mu, sg + done * B::SF, meta + done,
        gout + done * B::NU, m, p, grad_a + done * B::NU);
  }
  cudaError_t e = cudaGetLastError();
  return e == cudaSuccess ? 0 : (int)e;
}

// ------------------------------------------------------------------------------------------------------------
// K1: environment steps / resets / prior model
// ------------------------------------------------------------------------------------------------------------
template <typename T> struct Vec4;
template <> struct Vec4<float> { using type = float4; };
template <> struct Vec4<double> { using type = double4; };

template <typename T>
__device__ __forceinline__ void load_state4(const T* __restrict__ state4, int64_t i, T v[4]) {
  if constexpr (sizeof(T) == 4) {
    const float4 q = reinterpret_cast<const float4*>(state4)[i];
    v[0] = q.x; v[1] = q.y; v[2] = q.z; v[3] = q.w;
  } else {
    const double2 a = reinterpret_cast<const double2*>(state4)[2 * i];
    const double2 b = reinterpret_cast<const double2*>(state4)[2 * i + 1];
    v[0] = a.x; v[1] = a.y; v[2] = b.x; v[3] = b.y;
  }
}
template <typename T>
__device__ __forceinline__ void store_state4(T* __restrict__ state4, int64_t i, const T v[4]) {
  if constexpr (sizeof(T) == 4) {
    reinterpret_cast<float4*>(state4)[i] = make_float4(v[0], v[1], v[2], v[3]);
  } else {
    reinterpret_cast<double2*>(state4)[2 * i] = make_double2(v[0], v[1]);
    reinterpret_cast<double2*>(state4)[2 * i + 1] = make_double2(v[2], v[3]);
  }
}

template <typename T>
__global__ void __launch_bounds__(kThreads)
k_unicycle_env_step(T* __restrict__ state4, int32_t* __restrict__ step, const T* __restrict__ action, int64_t n,
                    UnicycleEnvParams e, T* __restrict__ obs, T* __restrict__ reward, uint8_t* __restrict__ done,
                    T* __restrict__ cost, uint8_t* __restrict__ goal_met) {
  __shared__ T s_rows[kThreads * 7];
  BlockRows<T, 7> rows(s_rows);
  const int64_t i0 = (int64_t)blockIdx.x * kThreads;
  const int64_t i = i0 + threadIdx.x;
  const bool valid = i < n;
  UniEnvOut<T> o;
#pragma unroll
  for (int k = 0; k < 7; ++k) o.obs[k] = T(0);
  if (valid) {
    T v[4], a[2];
    load_state4<T>(state4, i, v);
    load_row<2>(action, i, a);
    int stp = step[i];
    if constexpr (sizeof(T) == 4) {
      float s0, c0;
      sincos_t(v[2], &s0, &c0);
      unicycle_env_step_sc(e, v, v[3], stp, a, s0, c0, o);
    } else {
      unicycle_env_step<T>(e, v, v[3], stp, a, o);
    }
    reward[i] = o.reward;
    done[i] = (uint8_t)o.done;
    cost[i] = o.cost;
    goal_met[i] = (uint8_t)o.goal_met;
    if (e.auto_reset && o.done) unicycle_reset<T>(e, v, v[3], stp);
    store_state4<T>(state4, i, v);
    step[i] = stp;
  }
  rows.store(obs, i0, n, o.obs);
}

template <typename T>
__global__ void __launch_bounds__(kThreads)
k_unicycle_env_reset(T* __restrict__ state4, int32_t* __restrict__ step, const uint8_t* __restrict__ mask, int64_t n,
                     UnicycleEnvParams e, T* __restrict__ obs) {
  const int64_t i = (int64_t)blockIdx.x * kThreads + threadIdx.x;
  if (i >= n) return;
  T v[4];
  int stp;
  const bool doit = (mask == nullptr) || mask[i];
  if (doit) {
    unicycle_reset<T>(e, v, v[3], stp);
    store_state4<T>(state4, i, v);
    step[i] = stp;
  } else {
    load_state4<T>(state4, i, v);
  }
  if (obs != nullptr) {
    T s, c, ob[7];
    sincos_t(v[2], &s, &c);
    unicycle_obs<T>(e, v, c, s, unicycle_goal_dist<T>(e, v), ob);
    store_row<7>(obs, i, ob);
  }
}

template <typename T>
__global__ void __launch_bounds__(kThreads)
k_cars_env_step(T* __restrict__ state, T* __restrict__ t, int32_t* __restrict__ step, const T* __restrict__ action,
                int64_t n, CarsEnvParams e, T* __restrict__ obs, T* __restrict__ reward, uint8_t* __restrict__ done,
                T* __restrict__ cost) {
  __shared__ T s_rows[kThreads * 10];
  BlockRows<T, 10> rows(s_rows);
  const int64_t i0 = (int64_t)blockIdx.x * kThreads;
  const int64_t i = i0 + threadIdx.x;
  const bool valid = i < n;
  T s[10];
  rows.load(state, i0, n, s);
  CarsEnvOut<T> o;
#pragma unroll
  for (int k = 0; k < 10; ++k) o.obs[k] = T(0);
  if (valid) {
    T tt = t[i];
    int stp = step[i];
    cars_env_step<T>(e, s, tt, stp, action[i], o);
    reward[i] = o.reward;
    done[i] = (uint8_t)o.done;
    cost[i] = o.cost;
    t[i] = tt;
    step[i] = stp;
  }
  rows.store(obs, i0, n, o.obs);
  rows.store(state, i0, n, s);
}

template <typename T>
__global__ void __launch_bounds__(kThreads)
k_cars_env_reset(T* __restrict__ state, T* __restrict__ t, int32_t* __restrict__ step, const T* __restrict__ v_noise,
                 const uint8_t* __restrict__ mask, int64_t n, T* __restrict__ obs) {
  const int64_t i = (int64_t)blockIdx.x * kThreads + threadIdx.x;
  if (i >= n) return;
  T s[10];
  const bool doit = (mask == nullptr) || mask[i];
  if (doit) {
    T tt;
    int stp;
    cars_reset<T>(s, tt, stp, v_noise[i]);
    store_row<10>(state, i, s);
    t[i] = tt;
    step[i] = stp;
  } else {
    load_row<10>(state, i, s);
  }
  if (obs != nullptr) {
    T ob[10];
    cars_obs<T>(s, ob);
    store_row<10>(obs, i, ob);
  }
}

template <typename T>
__global__ void __launch_bounds__(kThreads)
k_unicycle_predict_next(const T* __restrict__ state, const T* __restrict__ action, const T* __restrict__ mean, int64_t n,
                        T dt, T* __restrict__ next) {
  const int64_t i = (int64_t)blockIdx.x * kThreads + threadIdx.x;  // (3-word rows: staging through shared memory
  if (i >= n) return;                                              //  measured slower than the direct accesses)
  T s[3], u[2], m[3] = {T(0), T(0), T(0)}, nx[3];
  load_row<3>(state, i, s);
  load_row<2>(action, i, u);
  if (mean != nullptr) load_row<3>(mean, i, m);
  unicycle_prior_next<T>(dt, s, u, m, nx);
  store_row<3>(next, i, nx);
}

template <typename T>
__global__ void __launch_bounds__(kThreads)
k_cars_predict_next(const T* __restrict__ state, const T* __restrict__ action, const T* __restrict__ t,
                    const T* __restrict__ mean, int64_t n, T dt, T kp, T kb, T* __restrict__ next) {
  __shared__ T s_rows[kThreads * 10];
  BlockRows<T, 10> rows(s_rows);
  const int64_t i0 = (int64_t)blockIdx.x * kThreads;
  const int64_t i = i0 + threadIdx.x;
  T s[10], m[10], nx[10];
  rows.load(state, i0, n, s);
#pragma unroll
  for (int j = 0; j < 10; ++j) m[j] = T(0);
  if (mean != nullptr) rows.load(mean, i0, n, m);
  const T a = (i < n) ? action[i] : T(0), tt = (i < n) ? t[i] : T(0);
  cars_prior_next<T>(dt, kp, kb, s, a, tt, m, nx);
  rows.store(next, i0, n, nx);
}

// ------------------------------------------------------------------------------------------------------------
// model-rollout transition (generate_model_rollouts), one thread per instance
// ------------------------------------------------------------------------------------------------------------
template <typename T>
__global__ void __launch_bounds__(kThreads)
k_unicycle_rollout_step(const T* __restrict__ obs, const T* __restrict__ action, const T* __restrict__ mean,
                        const T* __restrict__ std, const T* __restrict__ eps, int64_t n, T dt, T gx, T gy,
                        T* __restrict__ next_obs, T* __restrict__ reward, uint8_t* __restrict__ done) {
  const int64_t i = (int64_t)blockIdx.x * kThreads + threadIdx.x;  // (short rows: direct accesses measured faster)
  if (i >= n) return;
  T o[7], a[2], m[3] = {T(0), T(0), T(0)}, sd[3] = {T(0), T(0), T(0)}, e[3] = {T(0), T(0), T(0)}, no[7], r;
  int d;
  load_row<7>(obs, i, o);
  load_row<2>(action, i, a);
  if (mean != nullptr) load_row<3>(mean, i, m);
  if (std != nullptr && eps != nullptr) {
    load_row<3>(std, i, sd);
    load_row<3>(eps, i, e);
  }
  unicycle_rollout_step<T>(dt, gx, gy, o, a, m, sd, e, no, r, d);
  store_row<7>(next_obs, i, no);
  reward[i] = r;
  done[i] = (uint8_t)d;
}

template <typename T>
__global__ void __launch_bounds__(kThreads)
k_cars_rollout_step(const T* __restrict__ obs, const T* __restrict__ action, const T* __restrict__ t,
                    const T* __restrict__ mean, const T* __restrict__ std, const T* __restrict__ eps, int64_t n, T dt,
                    T kp, T kb, int max_steps, T* __restrict__ next_obs, T* __restrict__ reward,
                    uint8_t* __restrict__ done, T* __restrict__ next_t) {
  __shared__ T s_rows[kThreads * 10];
  BlockRows<T, 10> rows(s_rows);
  const int64_t i0 = (int64_t)blockIdx.x * kThreads;
  const int64_t i = i0 + threadIdx.x;
  T o[10], m[10], sd[10], e[10], no[10], r, nt;
  int d;
  rows.load(obs, i0, n, o);
#pragma unroll
  for (int j = 0; j < 10; ++j) m[j] = sd[j] = e[j] = T(0);
  if (mean != nullptr) rows.load(mean, i0, n, m);
  if (std != nullptr && eps != nullptr) {
    rows.load(std, i0, n, sd);
    rows.load(eps, i0, n, e);
  }
  const T a = (i < n) ? action[i] : T(0), tt = (i < n) ? t[i] : T(0);
  cars_rollout_step<T>(dt, kp, kb, max_steps, o, a, tt, m, sd, e, no, r, d, nt);
  rows.store(next_obs, i0, n, no);
  if (i < n) {
    reward[i] = r;
    done[i] = (uint8_t)d;
    next_t[i] = nt;
  }
}

// ------------------------------------------------------------------------------------------------------------
// generic QP (cbf_layer / solve_qp API), float64
// ------------------------------------------------------------------------------------------------------------
template <int NZ, int M>
__global__ void __launch_bounds__(kThreads)
k_qp_solve(const double* __restrict__ Q, const double* __restrict__ p, const double* __restrict__ G,
           const double* __restrict__ h, int64_t n, double* __restrict__ x, double* __restrict__ lam,
           double* __restrict__ slack, int32_t* __restrict__ status, int32_t* __restrict__ iters,
           rcbf_counters_t* counters) {
  const int64_t i0 = (int64_t)blockIdx.x * kThreads + threadIdx.x;
  const bool valid = i0 < n;
  const int64_t i = valid ? i0 : n - 1;
  double xo[NZ], lo[M], so[M];
  int st, it;
  generic_qp_solve<NZ, M>(Q + i * NZ * NZ, p + i * NZ, G + i * M * NZ, h + i * M, xo, lo, so, st, it);
  if (valid) {
    store_row<NZ>(x, i, xo);
    if (lam != nullptr) store_row<M>(lam, i, lo);
    if (slack != nullptr) store_row<M>(slack, i, so);
    if (status != nullptr) status[i] = st;
    if (iters != nullptr) iters[i] = it;
  }
  accumulate_counters(counters, valid, st, it);
}

template <int NZ, int M>
__global__ void __launch_bounds__(kThreads)
k_qp_solve_bwd(const double* __restrict__ Q, const double* __restrict__ G, const double* __restrict__ x,
               const double* __restrict__ lam, const double* __restrict__ slack, const double* __restrict__ gx, int64_t n,
               double* __restrict__ dQ, double* __restrict__ dp, double* __restrict__ dG, double* __restrict__ dh) {
  const int64_t i = (int64_t)blockIdx.x * kThreads + threadIdx.x;
  if (i >= n) return;
  generic_qp_bwd<NZ, M>(Q + i * NZ * NZ, G + i * M * NZ, x + i * NZ, lam + i * M, slack + i * M, gx + i * NZ,
                        dQ + i * NZ * NZ, dp + i * NZ, dG + i * M * NZ, dh + i * M);
}

// ------------------------------------------------------------------------------------------------------------
// FP32 FMA probe (roofline denominator for an FP32-pipe-bound path): 8 independent chains per thread
// ------------------------------------------------------------------------------------------------------------
// counters -> a host-visible mirror, then the caller's token: a host thread that polls mirror[0] learns that everything
// enqueued before this launch has finished AND reads the counters, without a cudaMemcpy / cudaStreamSynchronize pair
__global__ void k_counters_publish(const rcbf_counters_t* __restrict__ ws, volatile unsigned long long* mirror,
                                   unsigned long long token) {
  if (threadIdx.x < 8) mirror[1 + threadIdx.x] = ws[threadIdx.x];
  __syncwarp();
  __threadfence_system();
  if (threadIdx.x == 0) mirror[0] = token;
}

__global__ void k_bind_mirror(rcbf_counters_t* ws, unsigned long long mirror) { ws[11] = mirror; }  // kWsMirror (rcbf_safe_kernels.cuh)

__global__ void k_fp32_fma_probe(float* sink, int iters) {
  float a0 = threadIdx.x * 1e-3f, a1 = a0 + 1.f, a2 = a0 + 2.f, a3 = a0 + 3.f, a4 = a0 + 4.f, a5 = a0 + 5.f,
        a6 = a0 + 6.f, a7 = a0 + 7.f;
  const float b = 0.999f + blockIdx.x * 1e-9f, c = 1e-3f;
#pragma unroll 4
  for (int k = 0; k < iters; ++k) {
    a0 = fmaf(a0, b, c); a1 = fmaf(a1, b, c); a2 = fmaf(a2, b, c); a3 = fmaf(a3, b, c);
    a4 = fmaf(a4, b, c); a5 = fmaf(a5, b, c); a6 = fmaf(a6, b, c); a7 = fmaf(a7, b, c);
  }
  const float r = a0 + a1 + a2 + a3 + a4 + a5 + a6 + a7;
  if (r == 123.456f) sink[0] = r;  // never true; keeps the chains alive
}

}  // namespace

// ================================================================================================================
// C ABI
// ================================================================================================================
extern "C" {

const char* rcbf_version(void) { return "rcbf_b200 0.2 (sm_100a)"; }

int rcbf_unicycle_assemble(const float* state, const float* action, const float* mean, const float* sigma, int64_t n,
                           const rcbf_unicycle_params* p, float* G, float* h, void* stream) {
  if (n <= 0) return 0;
  k_unicycle_assemble<<<grid_for(n), kThreads, 0, (cudaStream_t)stream>>>(state, action, mean, sigma, n, *p, G, h);
  RCBF_LAUNCH_CHECK();
  return 0;
}

int rcbf_unicycle_assemble_f64(const double* state, const double* action, const double* mean, const double* sigma,
                               int64_t n, const rcbf_unicycle_params* p, int normalise, double* G, double* h, void* stream) {
  if (n <= 0) return 0;
  k_unicycle_assemble_f64<<<grid_for(n), kThreads, 0, (cudaStream_t)stream>>>(state, action, mean, sigma, n, *p, normalise,
                                                                              G, h);
  RCBF_LAUNCH_CHECK();
  return 0;
}

int rcbf_cars_assemble_f64(const double* state, const double* action, const double* sigma, int64_t n,
                           const rcbf_cars_params* p, int normalise, double* G, double* h, void* stream) {
  if (n <= 0) return 0;
  k_cars_assemble_f64<<<grid_for(n), kThreads, 0, (cudaStream_t)stream>>>(state, action, sigma, n, *p, normalise, G, h);
  RCBF_LAUNCH_CHECK();
  return 0;
}

int rcbf_cars_assemble(const float* state, const float* action, const float* sigma, int64_t n, const rcbf_cars_params* p,
                       float* G, float* h, void* stream) {
  if (n <= 0) return 0;
  k_cars_assemble<<<grid_for(n), kThreads, 0, (cudaStream_t)stream>>>(state, action, sigma, n, *p, G, h);
  RCBF_LAUNCH_CHECK();
  return 0;
}



int rcbf_unicycle_safe_action_bwd(const float* state, const float* action, const float* mean, const float* sigma,
                                  const float* x, const float* lam, const float* slack, const float* grad_out, int64_t n,
                                  const rcbf_unicycle_params* p, float* grad_action, void* stream) {
  if (n <= 0) return 0;
  k_unicycle_safe_action_bwd<<<grid_for(n), kThreads, 0, (cudaStream_t)stream>>>(state, action, mean, sigma, x, lam,
                                                                                 slack, grad_out, n, *p, grad_action);
  RCBF_LAUNCH_CHECK();
  return 0;
}

int rcbf_cars_safe_action_bwd(const float* state, const float* action, const float* sigma, const float* x,
                              const float* lam, const float* slack, const float* grad_out, int64_t n,
                              const rcbf_cars_params* p, float* grad_action, void* stream) {
  if (n <= 0) return 0;
  k_cars_safe_action_bwd<<<grid_for(n), kThreads, 0, (cudaStream_t)stream>>>(state, action, sigma, x, lam, slack,
                                                                             grad_out, n, *p, grad_action);
  RCBF_LAUNCH_CHECK();
  return 0;
}

int rcbf_unicycle_safe_action_bwd_meta(const float* state, const float* action, const float* mean, const float* sigma,
                                       const int32_t* meta, const float* grad_out, int64_t n,
                                       const rcbf_unicycle_params* p, float* grad_action, void* stream) {
  if (n <= 0) return 0;
  return launch_bwd_meta<UniBwd>(state, action, mean, sigma, meta, grad_out, n, *p, grad_action, (cudaStream_t)stream);
}

int rcbf_cars_safe_action_bwd_meta(const float* state, const float* action, const float* sigma, const int32_t* meta,
                                   const float* grad_out, int64_t n, const rcbf_cars_params* p, float* grad_action,
                                   void* stream) {
  if (n <= 0) return 0;
  return launch_bwd_meta<CarsBwd>(state, action, nullptr, sigma, meta, grad_out, n, *p, grad_action, (cudaStream_t)stream);
}

int rcbf_qp_solve(const double* Q, const double* p, const double* G, const double* h, int64_t n, int nz, int m,
                  double* x, double* lam, double* slack, int32_t* status, int32_t* iters, rcbf_counters_t* counters,
                  void* stream) {
  if (n <= 0) return 0;
  cudaStream_t s = (cudaStream_t)stream;
  if (nz == 3 && m == 9)
    k_qp_solve<3, 9><<<grid_for(n), kThreads, 0, s>>>(Q, p, G, h, n, x, lam, slack, status, iters, counters);
  else if (nz == 3 && m == 12)   // layers on 6 .. 8 hazards (the caller pads smaller systems with duplicate rows)
    k_qp_solve<3, 12><<<grid_for(n), kThreads, 0, s>>>(Q, p, G, h, n, x, lam, slack, status, iters, counters);
  else if (nz == 3 && m == 16)   // ... 9 .. 12 hazards
    k_qp_solve<3, 16><<<grid_for(n), kThreads, 0, s>>>(Q, p, G, h, n, x, lam, slack, status, iters, counters);
  else if (nz == 2 && m == 4)
    k_qp_solve<2, 4><<<grid_for(n), kThreads, 0, s>>>(Q, p, G, h, n, x, lam, slack, status, iters, counters);
  else
    return -1;  // unsupported shape: the Python layer raises NotImplementedError
  RCBF_LAUNCH_CHECK();
  return 0;
}

int rcbf_qp_solve_bwd(const double* Q, const double* G, const double* x, const double* lam, const double* slack,
                      const double* grad_x, int64_t n, int nz, int m, double* dQ, double* dp, double* dG, double* dh,
                      void* stream) {
  if (n <= 0) return 0;
  cudaStream_t s = (cudaStream_t)stream;
  if (nz == 3 && m == 9)
    k_qp_solve_bwd<3, 9><<<grid_for(n), kThreads, 0, s>>>(Q, G, x, lam, slack, grad_x, n, dQ, dp, dG, dh);
  else if (nz == 3 && m == 12)
    k_qp_solve_bwd<3, 12><<<grid_for(n), kThreads, 0, s>>>(Q, G, x, lam, slack, grad_x, n, dQ, dp, dG, dh);
  else if (nz == 3 && m == 16)
    k_qp_solve_bwd<3, 16><<<grid_for(n), kThreads, 0, s>>>(Q, G, x, lam, slack, grad_x, n, dQ, dp, dG, dh);
  else if (nz == 2 && m == 4)
    k_qp_solve_bwd<2, 4><<<grid_for(n), kThreads, 0, s>>>(Q, G, x, lam, slack, grad_x, n, dQ, dp, dG, dh);
  else
    return -1;
  RCBF_LAUNCH_CHECK();
  return 0;
}

#define RCBF_ENV_FUNCS(SUF, T)                                                                                          \
  int rcbf_unicycle_env_reset_##SUF(T* state4, int32_t* step, const uint8_t* mask, int64_t n,                          \
                                    const rcbf_unicycle_env_params* e, T* obs, void* stream) {                         \
    if (n <= 0) return 0;                                                                                               \
    k_unicycle_env_reset<T><<<grid_for(n), kThreads, 0, (cudaStream_t)stream>>>(state4, step, mask, n, *e, obs);       \
    RCBF_LAUNCH_CHECK();                                                                                                \
    return 0;                                                                                                           \
  }                                                                                                                     \
  int rcbf_unicycle_env_step_##SUF(T* state4, int32_t* step, const T* action, int64_t n,                               \
                                   const rcbf_unicycle_env_params* e, T* obs, T* reward, uint8_t* done, T* cost,       \
                                   uint8_t* goal_met, void* stream) {                                                   \
    if (n <= 0) return 0;                                                                                               \
    k_unicycle_env_step<T><<<grid_for(n), kThreads, 0, (cudaStream_t)stream>>>(state4, step, action, n, *e, obs,       \
                                                                               reward, done, cost, goal_met);          \
    RCBF_LAUNCH_CHECK();                                                                                                \
    return 0;                                                                                                           \
  }                                                                                                                     \
  int rcbf_cars_env_reset_##SUF(T* state, T* t, int32_t* step, const T* v_noise, const uint8_t* mask, int64_t n,       \
                                T* obs, void* stream) {                                                                 \
    if (n <= 0) return 0;                                                                                               \
    k_cars_env_reset<T><<<grid_for(n), kThreads, 0, (cudaStream_t)stream>>>(state, t, step, v_noise, mask, n, obs);    \
    RCBF_LAUNCH_CHECK();                                                                                                \
    return 0;                                                                                                           \
  }                                                                                                                     \
  int rcbf_cars_env_step_##SUF(T* state, T* t, int32_t* step, const T* action, int64_t n,                              \
                               const rcbf_cars_env_params* e, T* obs, T* reward, uint8_t* done, T* cost,               \
                               void* stream) {                                                                          \
    if (n <= 0) return 0;                                                                                               \
    k_cars_env_step<T><<<grid_for(n), kThreads, 0, (cudaStream_t)stream>>>(state, t, step, action, n, *e, obs, reward, \
                                                                           done, cost);                                \
    RCBF_LAUNCH_CHECK();                                                                                                \
    return 0;                                                                                                           \
  }                                                                                                                     \
  int rcbf_unicycle_predict_next_##SUF(const T* state, const T* action, const T* mean, int64_t n, double dt, T* next,  \
                                       void* stream) {                                                                  \
    if (n <= 0) return 0;                                                                                               \
    k_unicycle_predict_next<T><<<grid_for(n), kThreads, 0, (cudaStream_t)stream>>>(state, action, mean, n, (T)dt,      \
                                                                                   next);                              \
    RCBF_LAUNCH_CHECK();                                                                                                \
    return 0;                                                                                                           \
  }                                                                                                                     \
  int rcbf_cars_predict_next_##SUF(const T* state, const T* action, const T* t, const T* mean, int64_t n, double dt,   \
                                   double kp, double k_brake, T* next, void* stream) {                                  \
    if (n <= 0) return 0;                                                                                               \
    k_cars_predict_next<T><<<grid_for(n), kThreads, 0, (cudaStream_t)stream>>>(state, action, t, mean, n, (T)dt,       \
                                                                               (T)kp, (T)k_brake, next);               \
    RCBF_LAUNCH_CHECK();                                                                                                \
    return 0;                                                                                                           \
  }

RCBF_ENV_FUNCS(f32, float)
RCBF_ENV_FUNCS(f64, double)

#define RCBF_ROLLOUT_FUNCS(SUF, T)                                                                                      \
  int rcbf_unicycle_rollout_step_##SUF(const T* obs, const T* action, const T* mean, const T* std, const T* eps,       \
                                       int64_t n, double dt, double goal_x, double goal_y, T* next_obs, T* reward,     \
                                       uint8_t* done, void* stream) {                                                  \
    if (n <= 0) return 0;                                                                                               \
    k_unicycle_rollout_step<T><<<grid_for(n), kThreads, 0, (cudaStream_t)stream>>>(                                    \
        obs, action, mean, std, eps, n, (T)dt, (T)goal_x, (T)goal_y, next_obs, reward, done);                          \
    RCBF_LAUNCH_CHECK();                                                                                                \
    return 0;                                                                                                           \
  }                                                                                                                     \
  int rcbf_cars_rollout_step_##SUF(const T* obs, const T* action, const T* t, const T* mean, const T* std,             \
                                   const T* eps, int64_t n, double dt, double kp, double k_brake, int max_steps,       \
                                   T* next_obs, T* reward, uint8_t* done, T* next_t, void* stream) {                   \
    if (n <= 0) return 0;                                                                                               \
    k_cars_rollout_step<T><<<grid_for(n), kThreads, 0, (cudaStream_t)stream>>>(                                        \
        obs, action, t, mean, std, eps, n, (T)dt, (T)kp, (T)k_brake, max_steps, next_obs, reward, done, next_t);       \
    RCBF_LAUNCH_CHECK();                                                                                                \
    return 0;                                                                                                           \
  }

RCBF_ROLLOUT_FUNCS(f32, float)
RCBF_ROLLOUT_FUNCS(f64, double)



int rcbf_counters_publish(const rcbf_counters_t* workspace, uint64_t* host_mirror, uint64_t token, void* stream) {
  k_counters_publish<<<1, 32, 0, (cudaStream_t)stream>>>(workspace, reinterpret_cast<volatile unsigned long long*>(host_mirror),
                                                        (unsigned long long)token);
  RCBF_LAUNCH_CHECK();
  return 0;
}

int rcbf_counters_bind_mirror(rcbf_counters_t* workspace, uint64_t* host_mirror, void* stream) {
  k_bind_mirror<<<1, 1, 0, (cudaStream_t)stream>>>(workspace, (unsigned long long)reinterpret_cast<uintptr_t>(host_mirror));
  RCBF_LAUNCH_CHECK();
  return 0;
}

int rcbf_stream_synchronize(void* stream) { return (int)cudaStreamSynchronize((cudaStream_t)stream); }

int rcbf_fp32_fma_probe(float* sink, int blocks, int threads, int iters, void* stream) {
  k_fp32_fma_probe<<<blocks, threads, 0, (cudaStream_t)stream>>>(sink, iters);
  RCBF_LAUNCH_CHECK();
  return 0;
}

// ---------------------------------------------------------------------------------------------------------------
// host-buffer entry points (e2e path): chunked H2D -> kernel -> D2H pipeline over three internal streams
// ---------------------------------------------------------------------------------------------------------------
}  // extern "C"

namespace {
// Host-buffer pipelines.  State is kept PER DEVICE (streams, staging scratch and solver workspaces belong to the device
// they were created on) and every entry point holds one process-wide lock: two Python threads may call the *_host
// functions concurrently, they are simply serialised.
struct HostPipe {
  bool ready = false;
  cudaStream_t streams[3] = {nullptr, nullptr, nullptr};
  char* scratch = nullptr;
  size_t scratch_bytes = 0;
  rcbf_counters_t* counters = nullptr;  // 3 workspaces of RCBF_WS_WORDS words, one per stream
  rcbf_counters_t* failed_host = nullptr;  // pinned: the 3 NaN counters come back with the last chunk, no extra sync
  int ensure(int dev, size_t bytes) {
    cudaError_t e;
    if ((e = cudaSetDevice(dev)) != cudaSuccess) return (int)e;
    if (!ready) {
      for (auto& s : streams)
        if ((e = cudaStreamCreateWithFlags(&s, cudaStreamNonBlocking)) != cudaSuccess) return (int)e;
      const size_t wsb = 3 * RCBF_WS_WORDS * sizeof(rcbf_counters_t);
      if ((e = cudaMalloc(&counters, wsb)) != cudaSuccess) return (int)e;
      // "zero-initialised once" (rcbf_b200.h): queue bookkeeping and slots must start at zero; the kernels leave them
      // at zero, the per-call code below only clears the 8 counter words
      if ((e = cudaMemset(counters, 0, wsb)) != cudaSuccess) return (int)e;
      if ((e = cudaMallocHost(&failed_host, 3 * sizeof(rcbf_counters_t))) != cudaSuccess) return (int)e;
      ready = true;
    }
    if (bytes > scratch_bytes) {
      for (auto& s : streams) cudaStreamSynchronize(s);
      if (scratch) cudaFree(scratch);
      scratch = nullptr;
      scratch_bytes = 0;
      if ((e = cudaMalloc(&scratch, bytes)) != cudaSuccess) return (int)e;
      scratch_bytes = bytes;
    }
    return 0;
  }
};
constexpr int kMaxDevices = 64;
HostPipe g_pipes[kMaxDevices];
std::mutex g_pipe_mutex;

// Generic pipeline: `n_in` HOST input arrays, `n_tmp` device-only intermediates and `n_out` output arrays of the given
// BYTE widths per instance.  An output whose host pointer is NULL is still produced on the device (the kernels write
// every array) but is not copied back.  launch(dev_in[], dev_tmp[], dev_out[], lo, cnt, ws, stream) enqueues the work
// of one slice.  Slices are multiples of 64 instances (tiles of the two-per-lane kernel, 16-byte aligned arrays).
template <typename LaunchFn>
int run_pipe(const void* const* in_host, const int* in_bytes, int n_in, const int* tmp_bytes, int n_tmp,
             void* const* out_host, const int* out_bytes, int n_out, int64_t n, int device, int chunks,
             int32_t* n_failed_host, LaunchFn launch) {
  if (n <= 0) {
    if (n_failed_host) *n_failed_host = 0;
    return 0;
  }
  if (device < 0 || device >= kMaxDevices) return (int)cudaErrorInvalidDevice;
  std::lock_guard<std::mutex> lock(g_pipe_mutex);
  HostPipe& pp = g_pipes[device];
  size_t total = 0;
  size_t off_in[8], off_tmp[8], off_out[8];
  auto place = [&](size_t* off, const int* w, int cnt_) {
    for (int k = 0; k < cnt_; ++k) {
      off[k] = total;
      total += (((size_t)n * w[k]) + 255) & ~(size_t)255;
    }
  };
  place(off_in, in_bytes, n_in);
  place(off_tmp, tmp_bytes, n_tmp);
  place(off_out, out_bytes, n_out);
  int rc = pp.ensure(device, total);
  if (rc) return rc;
  if (chunks < 1) chunks = 1;
  for (int k = 0; k < 3; ++k)
    cudaMemsetAsync(pp.counters + k * RCBF_WS_WORDS, 0, 8 * sizeof(rcbf_counters_t), pp.streams[k]);
  void* dev_in[8];
  void* dev_tmp[8];
  void* dev_out[8];
  const int64_t per = (((n + chunks - 1) / chunks) + 63) & ~(int64_t)63;
  for (int c = 0; c < chunks; ++c) {
    const int64_t lo = (int64_t)c * per;
    const int64_t cnt = (lo + per <= n) ? per : (n - lo);
    if (cnt <= 0) break;
    cudaStream_t s = pp.streams[c % 3];
    for (int k = 0; k < n_in; ++k) {
      dev_in[k] = pp.scratch + off_in[k] + (size_t)lo * in_bytes[k];
      cudaMemcpyAsync(dev_in[k], static_cast<const char*>(in_host[k]) + (size_t)lo * in_bytes[k], (size_t)cnt * in_bytes[k],
                      cudaMemcpyHostToDevice, s);
    }
    for (int k = 0; k < n_tmp; ++k) dev_tmp[k] = pp.scratch + off_tmp[k] + (size_t)lo * tmp_bytes[k];
    for (int k = 0; k < n_out; ++k) dev_out[k] = pp.scratch + off_out[k] + (size_t)lo * out_bytes[k];
    rc = launch(dev_in, dev_tmp, dev_out, lo, cnt, pp.counters + (c % 3) * RCBF_WS_WORDS, s);
    if (rc) return rc;
    for (int k = 0; k < n_out; ++k)
      if (out_host[k] != nullptr)
        cudaMemcpyAsync(static_cast<char*>(out_host[k]) + (size_t)lo * out_bytes[k], dev_out[k], (size_t)cnt * out_bytes[k],
                        cudaMemcpyDeviceToHost, s);
  }
  for (int k = 0; k < 3; ++k)
    cudaMemcpyAsync(pp.failed_host + k, pp.counters + k * RCBF_WS_WORDS, sizeof(rcbf_counters_t), cudaMemcpyDeviceToHost,
                    pp.streams[k]);
  for (auto& s : pp.streams) {
    cudaError_t e = cudaStreamSynchronize(s);
    if (e != cudaSuccess) return (int)e;
  }
  if (n_failed_host) *n_failed_host = (int32_t)(pp.failed_host[0] + pp.failed_host[1] + pp.failed_host[2]);
  return (int)cudaGetLastError();
}
}  // namespace

extern "C" {

int rcbf_unicycle_safe_step_host(float* state4, int32_t* step, const float* action_host, const float* mean_host,
                                 const float* sigma_host, int64_t n, const rcbf_unicycle_params* p,
                                 const rcbf_unicycle_env_params* e, float* safe_action_host, float* obs_host,
                                 float* reward_host, uint8_t* done_host, float* cost_host, uint8_t* goal_met_host,
                                 int32_t* n_failed_host, int device, int chunks) {
  const void* in[3] = {action_host, mean_host, sigma_host};
  const int inb[3] = {8, 12, 12};
  void* out[6] = {safe_action_host, obs_host, reward_host, done_host, cost_host, goal_met_host};
  const int outb[6] = {8, 28, 4, 1, 4, 1};
  return run_pipe(in, inb, 3, nullptr, 0, out, outb, 6, n, device, chunks, n_failed_host,
                  [&](void** di, void**, void** d_o, int64_t lo, int64_t cnt, rcbf_counters_t* ws, cudaStream_t s) {
                    return rcbf_unicycle_safe_step(state4 + lo * 4, step + lo, (const float*)di[0], (const float*)di[1],
                                                   (const float*)di[2], cnt, p, e, (float*)d_o[0], (float*)d_o[1],
                                                   (float*)d_o[2], (uint8_t*)d_o[3], (float*)d_o[4], (uint8_t*)d_o[5],
                                                   nullptr, ws, (void*)s);
                  });
}

int rcbf_unicycle_safe_step_host_gp(float* state4, int32_t* step, const float* action_host,
                                    const rcbf_gp_posterior* post, int64_t n, const rcbf_unicycle_params* p,
                                    const rcbf_unicycle_env_params* e, float* safe_action_host, float* obs_host,
                                    float* reward_host, uint8_t* done_host, float* cost_host, uint8_t* goal_met_host,
                                    int32_t* n_failed_host, int device, int chunks) {
  if (post == nullptr || post->n_in != 3 || post->n_gp != 3) return (int)cudaErrorInvalidValue;
  rcbf_gp_posterior pg = *post;
  pg.test_stride = 4;  // the resident float4 state (x, y, theta, last_goal_dist) is the test point, read in place
  const void* in[1] = {action_host};
  const int inb[1] = {8};
  const int tmpb[2] = {12, 12};  // disturbance mean / std: produced and consumed on the device
  void* out[6] = {safe_action_host, obs_host, reward_host, done_host, cost_host, goal_met_host};
  const int outb[6] = {8, 28, 4, 1, 4, 1};
  return run_pipe(in, inb, 1, tmpb, 2, out, outb, 6, n, device, chunks, n_failed_host,
                  [&](void** di, void** dt, void** d_o, int64_t lo, int64_t cnt, rcbf_counters_t* ws, cudaStream_t s) {
                    int rc = rcbf_gp_predict_f32(state4 + lo * 4, cnt, &pg, (float*)dt[0], (float*)dt[1], (void*)s);
                    if (rc) return rc;
                    return rcbf_unicycle_safe_step(state4 + lo * 4, step + lo, (const float*)di[0], (const float*)dt[0],
                                                   (const float*)dt[1], cnt, p, e, (float*)d_o[0], (float*)d_o[1],
                                                   (float*)d_o[2], (uint8_t*)d_o[3], (float*)d_o[4], (uint8_t*)d_o[5],
                                                   nullptr, ws, (void*)s);
                  });
}

int rcbf_cars_safe_step_host(float* state, float* t, int32_t* step, const float* action_host, const float* sigma_host,
                             int64_t n, const rcbf_cars_params* p, const rcbf_cars_env_params* e,
                             float* safe_action_host, float* obs_host, float* reward_host, uint8_t* done_host,
                             float* cost_host, int32_t* n_failed_host, int device, int chunks) {
  const void* in[2] = {action_host, sigma_host};
  const int inb[2] = {4, 40};
  void* out[5] = {safe_action_host, obs_host, reward_host, done_host, cost_host};
  const int outb[5] = {4, 40, 4, 1, 4};
  return run_pipe(in, inb, 2, nullptr, 0, out, outb, 5, n, device, chunks, n_failed_host,
                  [&](void** di, void**, void** d_o, int64_t lo, int64_t cnt, rcbf_counters_t* ws, cudaStream_t s) {
                    return rcbf_cars_safe_step(state + lo * 10, t + lo, step + lo, (const float*)di[0],
                                               (const float*)di[1], cnt, p, e, (float*)d_o[0], (float*)d_o[1],
                                               (float*)d_o[2], (uint8_t*)d_o[3], (float*)d_o[4], nullptr, ws, (void*)s);
                  });
}

int rcbf_unicycle_safe_action_host(const float* state_host, const float* action_host, const float* mean_host,
                                   const float* sigma_host, int64_t n, const rcbf_unicycle_params* p,
                                   float* safe_action_host, int32_t* n_failed_host, int device, int chunks) {
  const void* in[4] = {state_host, action_host, mean_host, sigma_host};
  const int inb[4] = {12, 8, 12, 12};
  void* out[1] = {safe_action_host};
  const int outb[1] = {8};
  return run_pipe(in, inb, 4, nullptr, 0, out, outb, 1, n, device, chunks, n_failed_host,
                  [&](void** d, void**, void** d_o, int64_t, int64_t cnt, rcbf_counters_t* ctr, cudaStream_t s) {
                    return rcbf_unicycle_safe_action((const float*)d[0], (const float*)d[1], (const float*)d[2],
                                                     (const float*)d[3], cnt, p, (float*)d_o[0], nullptr, nullptr,
                                                     nullptr, nullptr, nullptr, ctr, (void*)s);
                  });
}

int rcbf_cars_safe_action_host(const float* state_host, const float* action_host, const float* sigma_host, int64_t n,
                               const rcbf_cars_params* p, float* safe_action_host, int32_t* n_failed_host, int device,
                               int chunks) {
  const void* in[3] = {state_host, action_host, sigma_host};
  const int inb[3] = {40, 4, 40};
  void* out[1] = {safe_action_host};
  const int outb[1] = {4};
  return run_pipe(in, inb, 3, nullptr, 0, out, outb, 1, n, device, chunks, n_failed_host,
                  [&](void** d, void**, void** d_o, int64_t, int64_t cnt, rcbf_counters_t* ctr, cudaStream_t s) {
                    return rcbf_cars_safe_action((const float*)d[0], (const float*)d[1], (const float*)d[2], cnt, p,
                                                 (float*)d_o[0], nullptr, nullptr, nullptr, nullptr, nullptr, ctr,
                                                 (void*)s);
                  });
}

}  // extern "C"
