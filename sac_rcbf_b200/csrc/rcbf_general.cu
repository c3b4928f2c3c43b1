// rcbf_general.cu -- the Unicycle safety layer for an ARBITRARY number of hazards (1 .. RCBF_MAX_HAZARDS).
//
// The reference sizes its layer from len(env.hazards_locations) (rcbf_sac/diff_cbf_qp.py:35,42; CBF row loop :243-261);
// the hot kernels (rcbf_safe_kernels.cuh, rcbf_safe2.cuh) are specialised for the 5 hazards of the reference env
// (unicycle_env.py:26).  A layer built on another hazard set runs here: the same per-instance source (assembly in
// reference op order, raw-row greedy presolve, float64 KKT certificate, interior-point fallback, exact active-set
// backward -- rcbf_core.cuh / rcbf_backward.cuh are templated on the row count) instantiated for NH = 8 and NH = 12 CBF
// rows, one instance per thread.  Fewer hazards than NH are padded with inert ones 1e4 m away (h ~ 1e24 > 0: such a row
// can never be active and is sliced out of the assembled constraints).
#include <cuda_runtime.h>
#include <stdint.h>

#include "rcbf_backward.cuh"
#include "rcbf_core.cuh"

using namespace rcbf;

namespace {

constexpr int kThreads = 128;
inline unsigned grid_for(int64_t n) { return (unsigned)((n + kThreads - 1) / kThreads); }

template <int NH>
struct GenHaz {
  float xy[NH][2];
};

template <int NH>
struct GenRaw {
  float G[NH + 4][3];
  float h[NH + 4];
  float Lg[NH][2];
};

template <int NH>
__device__ __forceinline__ void gen_assemble(const UnicycleParams& p, const GenHaz<NH>& hz, const float* __restrict__ st,
                                             const float* __restrict__ ac, const float* __restrict__ mu,
                                             const float* __restrict__ sg, int64_t i, float u[2], GenRaw<NH>& o) {
  float s[3], m[3], g[3];
#pragma unroll
  for (int j = 0; j < 3; ++j) {
    s[j] = __ldg(st + i * 3 + j);
    m[j] = __ldg(mu + i * 3 + j);
    g[j] = __ldg(sg + i * 3 + j);
  }
  u[0] = __ldg(ac + i * 2);
  u[1] = __ldg(ac + i * 2 + 1);
  float sn, cs;
  sincos_t(s[2], &sn, &cs);
  assemble_unicycle_n<float, NH>(p, hz.xy, s, sn, cs, u, m, g, o.Lg, o.h);
#pragma unroll
  for (int k = 0; k < NH; ++k) {
    o.G[k][0] = -o.Lg[k][0];
    o.G[k][1] = -o.Lg[k][1];
    o.G[k][2] = -1.0f;                                     // diff_cbf_qp.py:260
  }
#pragma unroll
  for (int cc = 0; cc < 2; ++cc) {                         // :365-377
    const int r = NH + 2 * cc;
#pragma unroll
    for (int j = 0; j < 3; ++j) {
      o.G[r][j] = (j == cc) ? 1.0f : 0.0f;
      o.G[r + 1][j] = (j == cc) ? -1.0f : 0.0f;
    }
  }
}

template <int NH>
__device__ __forceinline__ void gen_r(const GenRaw<NH>& raw, float r[NH + 4][2]) {
#pragma unroll
  for (int k = 0; k < NH; ++k) {
    r[k][0] = raw.Lg[k][0];
    r[k][1] = raw.Lg[k][1];
  }
#pragma unroll
  for (int c = 0; c < 2; ++c) {  // h = u_max - a_c ; h = -u_min + a_c
    r[NH + 2 * c][0] = (c == 0) ? -1.f : 0.f;
    r[NH + 2 * c][1] = (c == 1) ? -1.f : 0.f;
    r[NH + 2 * c + 1][0] = (c == 0) ? 1.f : 0.f;
    r[NH + 2 * c + 1][1] = (c == 1) ? 1.f : 0.f;
  }
}

// the instances the greedy presolve cannot certify (a drop is needed, degeneracy): float32 interior point + certificate,
// then the float64 interior point -- out of line, so that its state does not set the register allocation of the kernel
template <int NH>
__device__ __noinline__ void gen_solve_full(const GenRaw<NH>* raw, const float* p_diag, NormSolution<3, NH + 4>* sol) {
  Normalised<3, NH + 4> nrm;
  normalise_rows<CbfPat<NH, 2>, 3, NH + 4>(raw->G, raw->h, nrm);
  solve_normalised_full<CbfPat<NH, 2>, 3, NH + 4>(nrm, p_diag, false, *sol);
}

template <int NH>
__global__ void __launch_bounds__(kThreads)
k_general_safe_action(const float* __restrict__ st, const float* __restrict__ ac, const float* __restrict__ mu,
                      const float* __restrict__ sg, int64_t n, const __grid_constant__ UnicycleParams p,
                      const __grid_constant__ GenHaz<NH> hz, float* __restrict__ out, int32_t* __restrict__ meta,
                      int32_t* __restrict__ status, rcbf_counters_t* counters) {
  constexpr int M = NH + 4;
  using Pat = CbfPat<NH, 2>;
  const int64_t i = (int64_t)blockIdx.x * kThreads + threadIdx.x;
  int stt = -1, rounds = 0;
  bool fell_back = false;
  if (i < n) {
    GenRaw<NH> raw;
    float u[2];
    gen_assemble<NH>(p, hz, st, ac, mu, sg, i, u, raw);
    bool triv, nan;
    classify_raw<M>(raw.h, triv, nan);
#pragma unroll
    for (int k = 0; k < NH; ++k) nan = nan || (raw.Lg[k][0] != raw.Lg[k][0]) || (raw.Lg[k][1] != raw.Lg[k][1]);
    NormSolution<3, M> sol;
    sol.x[0] = sol.x[1] = sol.x[2] = nan ? (double)NAN : 0.0;
    sol.status = nan ? RCBF_NAN : RCBF_OK_TRIVIAL;
    sol.mask = 0u;
    sol.iters = 0;
    if (!triv && !nan) {
      solve_raw_fast<Pat, 3, M>(raw.G, raw.h, p.p_diag, false, sol);
      if (sol.status == RCBF_PENDING) {
        gen_solve_full<NH>(&raw, p.p_diag, &sol);
        fell_back = true;
      }
    }
    stt = sol.status;
    rounds = sol.iters >= 100 ? sol.iters - 100 : sol.iters;
    out[i * 2] = clampf(u[0] + (float)sol.x[0], p.u_min[0], p.u_max[0]);        // diff_cbf_qp.py:77
    out[i * 2 + 1] = clampf(u[1] + (float)sol.x[1], p.u_min[1], p.u_max[1]);
    if (meta != nullptr) meta[i] = (sol.status << 16) | (int)(sol.mask & 0xffffu);
    if (status != nullptr) status[i] = sol.status;
  }
  if (counters != nullptr) {
    const int n_nan = __syncthreads_count(stt == RCBF_NAN);
    const int n_unc = __syncthreads_count(stt == RCBF_MAXITER);
    const int n_triv = __syncthreads_count(stt == RCBF_OK_TRIVIAL);
    const int n_fb = __syncthreads_count(fell_back);
    __shared__ int s_it;
    if (threadIdx.x == 0) s_it = 0;
    __syncthreads();
    const int it = __reduce_add_sync(0xffffffffu, fell_back ? 0 : rounds);
    if ((threadIdx.x & 31) == 0 && it) atomicAdd(&s_it, it);
    __syncthreads();
    if (threadIdx.x == 0) {
      if (n_nan) atomicAdd(&counters[0], (unsigned long long)n_nan);
      if (n_unc) atomicAdd(&counters[1], (unsigned long long)n_unc);
      if (n_triv) atomicAdd(&counters[3], (unsigned long long)n_triv);
      if (s_it) atomicAdd(&counters[4], (unsigned long long)s_it);
      if (n_fb) atomicAdd(&counters[5], (unsigned long long)n_fb);
    }
  }
}

// dense qpth-clamp backward for an instance without a certified vertex (never observed): re-solves, out of line
template <int NH>
__device__ __noinline__ void gen_bwd_dense(const GenRaw<NH>* raw, const UnicycleParams* p, const float* u, const float* go,
                                           float* ga) {
  constexpr int M = NH + 4;
  Normalised<3, M> nrm;
  normalise_rows<CbfPat<NH, 2>, 3, M>(raw->G, raw->h, nrm);
  NormSolution<3, M> sol;
  solve_normalised_full<CbfPat<NH, 2>, 3, M>(nrm, p->p_diag, false, sol);
  float xs[3], ls[M], ss[M], r[M][2];
#pragma unroll
  for (int j = 0; j < 3; ++j) xs[j] = (float)sol.x[j];
#pragma unroll
  for (int k = 0; k < M; ++k) {
    ls[k] = (float)sol.lam[k];
    ss[k] = (float)sol.s[k];
  }
  gen_r<NH>(*raw, r);
  safe_action_bwd<3, M, 2>(nrm, raw->G, raw->h, r, p->p_diag, xs, ls, ss, u, p->u_min, p->u_max, go, ga);
}

template <int NH>
__global__ void __launch_bounds__(kThreads)
k_general_bwd(const float* __restrict__ st, const float* __restrict__ ac, const float* __restrict__ mu,
              const float* __restrict__ sg, const int32_t* __restrict__ meta, const float* __restrict__ gout, int64_t n,
              const __grid_constant__ UnicycleParams p, const __grid_constant__ GenHaz<NH> hz,
              float* __restrict__ grad_a) {
  constexpr int M = NH + 4;
  const int64_t i = (int64_t)blockIdx.x * kThreads + threadIdx.x;
  if (i >= n) return;
  const int mt = __ldg(meta + i);
  const int status = mt >> 16;
  const float go[2] = {__ldg(gout + i * 2), __ldg(gout + i * 2 + 1)};
  float ga[2];
  if (status == RCBF_OK_TRIVIAL || status == RCBF_NAN) {
#pragma unroll
    for (int c = 0; c < 2; ++c) {
      const float v = __ldg(ac + i * 2 + c) + 0.f;
      ga[c] = (status == RCBF_NAN) ? NAN : ((v >= p.u_min[c] && v <= p.u_max[c]) ? go[c] : 0.f);
    }
  } else {
    GenRaw<NH> raw;
    float u[2];
    gen_assemble<NH>(p, hz, st, ac, mu, sg, i, u, raw);
    if (status == RCBF_OK_CERTIFIED && (mt & 0xffff) != (int)kMaskUnknown) {
      float r[M][2];
      gen_r<NH>(raw, r);
      double pisd[3];
      float pisf[3];
      pis_of<3, M>(p.p_diag, pisd, pisf);
      safe_action_bwd_active<CbfPat<NH, 2>, 3, M, 2>(raw.G, raw.h, r, pisd, (uint32_t)mt & 0xffffu, u, p.u_min, p.u_max, go,
                                                     ga);
    } else {
      gen_bwd_dense<NH>(&raw, &p, u, go, ga);
    }
  }
  grad_a[i * 2] = ga[0];
  grad_a[i * 2 + 1] = ga[1];
}

// raw constraints as get_cbf_qp_constraints returns them: (n, K + 4, 3) and (n, K + 4), padding rows sliced out
template <int NH>
__global__ void __launch_bounds__(kThreads)
k_general_assemble(const float* __restrict__ st, const float* __restrict__ ac, const float* __restrict__ mu,
                   const float* __restrict__ sg, int64_t n, const __grid_constant__ UnicycleParams p,
                   const __grid_constant__ GenHaz<NH> hz, int K, float* __restrict__ G, float* __restrict__ h) {
  const int64_t i = (int64_t)blockIdx.x * kThreads + threadIdx.x;
  if (i >= n) return;
  GenRaw<NH> raw;
  float u[2];
  gen_assemble<NH>(p, hz, st, ac, mu, sg, i, u, raw);
  const int m_out = K + 4;
#pragma unroll
  for (int r = 0; r < NH + 4; ++r) {
    const int ro = r < NH ? r : r - NH + K;
    if (r < K || r >= NH) {
#pragma unroll
      for (int j = 0; j < 3; ++j) G[(i * m_out + ro) * 3 + j] = raw.G[r][j];
      h[i * m_out + ro] = raw.h[r];
    }
  }
}

template <int NH>
GenHaz<NH> pad_hazards(const float* xy, int K) {
  GenHaz<NH> hz;
  for (int k = 0; k < NH; ++k) {
    hz.xy[k][0] = k < K ? xy[2 * k] : 1.0e4f;
    hz.xy[k][1] = k < K ? xy[2 * k + 1] : 1.0e4f + 10.0f * (float)(k - K);
  }
  return hz;
}

}  // namespace

extern "C" {

int rcbf_unicycle_safe_action_general(const float* state, const float* action, const float* mean, const float* sigma,
                                      int64_t n, const rcbf_unicycle_params* p, const float* hazards_xy_host,
                                      int n_hazards, float* safe_action, int32_t* meta, int32_t* status,
                                      rcbf_counters_t* counters, void* stream) {
  if (n_hazards < 1 || n_hazards > RCBF_MAX_HAZARDS) return -1;
  if (n <= 0) return 0;
  cudaStream_t s = (cudaStream_t)stream;
  if (n_hazards <= 8)
    k_general_safe_action<8><<<grid_for(n), kThreads, 0, s>>>(state, action, mean, sigma, n, *p,
                                                              pad_hazards<8>(hazards_xy_host, n_hazards), safe_action,
                                                              meta, status, counters);
  else
    k_general_safe_action<12><<<grid_for(n), kThreads, 0, s>>>(state, action, mean, sigma, n, *p,
                                                               pad_hazards<12>(hazards_xy_host, n_hazards), safe_action,
                                                               meta, status, counters);
  return (int)cudaGetLastError();
}

int rcbf_unicycle_safe_action_bwd_general(const float* state, const float* action, const float* mean, const float* sigma,
                                          const int32_t* meta, const float* grad_out, int64_t n,
                                          const rcbf_unicycle_params* p, const float* hazards_xy_host, int n_hazards,
                                          float* grad_action, void* stream) {
  if (n_hazards < 1 || n_hazards > RCBF_MAX_HAZARDS) return -1;
  if (n <= 0) return 0;
  cudaStream_t s = (cudaStream_t)stream;
  if (n_hazards <= 8)
    k_general_bwd<8><<<grid_for(n), kThreads, 0, s>>>(state, action, mean, sigma, meta, grad_out, n, *p,
                                                      pad_hazards<8>(hazards_xy_host, n_hazards), grad_action);
  else
    k_general_bwd<12><<<grid_for(n), kThreads, 0, s>>>(state, action, mean, sigma, meta, grad_out, n, *p,
                                                       pad_hazards<12>(hazards_xy_host, n_hazards), grad_action);
  return (int)cudaGetLastError();
}

int rcbf_unicycle_assemble_general(const float* state, const float* action, const float* mean, const float* sigma,
                                   int64_t n, const rcbf_unicycle_params* p, const float* hazards_xy_host, int n_hazards,
                                   float* G, float* h, void* stream) {
  if (n_hazards < 1 || n_hazards > RCBF_MAX_HAZARDS) return -1;
  if (n <= 0) return 0;
  cudaStream_t s = (cudaStream_t)stream;
  if (n_hazards <= 8)
    k_general_assemble<8><<<grid_for(n), kThreads, 0, s>>>(state, action, mean, sigma, n, *p,
                                                           pad_hazards<8>(hazards_xy_host, n_hazards), n_hazards, G, h);
  else
    k_general_assemble<12><<<grid_for(n), kThreads, 0, s>>>(state, action, mean, sigma, n, *p,
                                                            pad_hazards<12>(hazards_xy_host, n_hazards), n_hazards, G, h);
  return (int)cudaGetLastError();
}

}  // extern "C"
