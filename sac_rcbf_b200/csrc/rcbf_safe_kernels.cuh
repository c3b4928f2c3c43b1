// rcbf_safe_kernels.cuh -- the hot kernels: get_safe_action forward (K2+K3) and the fused safe step (K5).
//
// One CUDA thread owns one instance, 256 threads per block, three phases per block:
//
//   A  every thread: coalesced loads, constraint assembly in the reference's float32 op order, row normalisation,
//      and the trivial test (h~ >= 0 on every row  <=>  x = 0 is optimal; ~2/3 of the synthetic instances).
//      Threads whose instance needs a solve pack its 28 (Unicycle) / 10 (SimulatedCars) words into shared memory at a
//      slot handed out with one ballot + one shared atomic per warp.
//   B  the first `count` threads of the block solve the packed problems (greedy active-set presolve + float64 KKT
//      certificate, or the float32 interior point in "pdipm" mode) and hand the correction back through shared memory.
//      Compaction is what keeps the warps of this phase full: without it they run at ~35 % lane utilisation.
//   C  every thread: clamp, and for the fused kernel env.step + all env outputs (one float4 state store).
//
// Instances phase B cannot certify get a tagged-NaN sentinel in safe_action[i][0] and are queued (workspace words
// [16..), see include/rcbf_b200.h); k_safe_fallback re-solves them with the interior-point chain and finishes them.
// If the caller passes no workspace, or the queue overflows, the fallback scans for the sentinel instead.
#pragma once

#include <cuda_runtime.h>

#include "rcbf_core.cuh"
#include "rcbf_dynamics.cuh"

namespace rcbf {

constexpr int kBlock = 256;
constexpr uint32_t kPendingBits = 0x7fc0dead;  // quiet NaN with a payload no arithmetic produces
constexpr int kWsCounters = 8;                 // workspace words [0, 8): counters
constexpr int kWsQueueCount = 8;               // [8]: number of queued instances
constexpr int kWsBlocksDone = 9;               // [9]: fallback blocks finished (last one resets the queue)
constexpr int kWsQueueBase = 16;               // [16, RCBF_WS_WORDS): queued instance indices
constexpr int kWsQueueCap = RCBF_WS_WORDS - kWsQueueBase;

template <int K>
__device__ __forceinline__ void ld_row(const float* __restrict__ base, int64_t i, float out[K]) {
#pragma unroll
  for (int j = 0; j < K; ++j) out[j] = __ldg(base + i * K + j);
}
template <int K>
__device__ __forceinline__ void st_row(float* __restrict__ base, int64_t i, const float in[K]) {
#pragma unroll
  for (int j = 0; j < K; ++j) base[i * K + j] = in[j];
}

// ---------------------------------------------------------------------------------------------------------------
// per-environment traits: argument block, loads, assembly, finishing
// ---------------------------------------------------------------------------------------------------------------
struct UniArgs {
  // inputs
  const float* st;   // (n,3) state          [safe_action]   -- unused by the fused kernel
  float* state4;     // (n,4) x,y,th,last    [safe_step]
  int32_t* step;     //                      [safe_step]
  const float* ac;   // (n,2) nominal action
  const float* mu;   // (n,3)
  const float* sg;   // (n,3)
  // outputs
  float* out;        // (n,2) safe action
  float* x;          // (n,3) nullable
  float* lam;        // (n,9) nullable
  float* slack;      // (n,9) nullable
  int32_t* status;   // nullable
  int32_t* iters;    // nullable
  float* obs;        // (n,7)  [safe_step]
  float* reward;
  uint8_t* done;
  float* cost;
  uint8_t* goal_met;
};

template <bool kFused>
struct UniEnv {
  static constexpr int NZ = kUniNZ, M = kUniM, NU = 2, NW = 28;
  using Pat = UniPat;
  using Args = UniArgs;
  using Params = UnicycleParams;
  using EnvParams = UnicycleEnvParams;
  struct Inst {
    float v[4];  // x, y, theta, last_goal_dist (fused) / unused
    float u[2];
    int stp;
  };
  __device__ static __forceinline__ void assemble(const Args& a, const Params& p, int64_t i, Inst& in,
                                                  Normalised<NZ, M>& nrm) {
    float m[3], g[3];
    if (kFused) {
      const float4 q = reinterpret_cast<const float4*>(a.state4)[i];
      in.v[0] = q.x; in.v[1] = q.y; in.v[2] = q.z; in.v[3] = q.w;
      in.stp = a.step[i];
    } else {
      ld_row<3>(a.st, i, in.v);
      in.v[3] = 0.f;
      in.stp = 0;
    }
    ld_row<2>(a.ac, i, in.u);
    ld_row<3>(a.mu, i, m);
    ld_row<3>(a.sg, i, g);
    UniRaw raw;
    assemble_unicycle(p, in.v, in.u, m, g, raw);
    normalise_rows<Pat, NZ, M>(raw.G, raw.h, nrm);
  }
  __device__ static __forceinline__ void finish(const Args& a, const Params& p, const EnvParams& e, int64_t i, Inst& in,
                                                const float xs[NU], int status) {
    float us[2];
#pragma unroll
    for (int c = 0; c < 2; ++c) us[c] = clampf(in.u[c] + xs[c], p.u_min[c], p.u_max[c]);  // diff_cbf_qp.py:77
    st_row<2>(a.out, i, us);
    if (a.status != nullptr) a.status[i] = status;
    if (kFused) {
      UniEnvOut<float> o;
      unicycle_env_step<float>(e, in.v, in.v[3], in.stp, us, o);
      st_row<7>(a.obs, i, o.obs);
      a.reward[i] = o.reward;
      a.done[i] = (uint8_t)o.done;
      a.cost[i] = o.cost;
      a.goal_met[i] = (uint8_t)o.goal_met;
      if (e.auto_reset && o.done) unicycle_reset<float>(e, in.v, in.v[3], in.stp);
      reinterpret_cast<float4*>(a.state4)[i] = make_float4(in.v[0], in.v[1], in.v[2], in.v[3]);
      a.step[i] = in.stp;
    }
  }
};

struct CarsArgs {
  const float* st;  // (n,10) [safe_action]
  float* state;     // (n,10) [safe_step] in/out
  float* t;         //        [safe_step]
  int32_t* step;
  const float* ac;  // (n,1)
  const float* sg;  // (n,10)
  float* out;       // (n,1)
  float* x;         // (n,2) nullable
  float* lam;       // (n,4) nullable
  float* slack;     // (n,4) nullable
  int32_t* status;
  int32_t* iters;
  float* obs;       // (n,10)
  float* reward;
  uint8_t* done;
  float* cost;
};

template <bool kFused>
struct CarsEnv {
  static constexpr int NZ = kCarsNZ, M = kCarsM, NU = 1, NW = 10;
  using Pat = CarsPat;
  using Args = CarsArgs;
  using Params = CarsParams;
  using EnvParams = CarsEnvParams;
  struct Inst {
    float u[1];
    float tt;
    int stp;
  };
  __device__ static __forceinline__ void assemble(const Args& a, const Params& p, int64_t i, Inst& in,
                                                  Normalised<NZ, M>& nrm) {
    float s[10], g[10];
    const float2* sp = reinterpret_cast<const float2*>(kFused ? a.state : a.st) + i * 5;  // rows are 40 B: 8-aligned
    const float2* gp = reinterpret_cast<const float2*>(a.sg) + i * 5;
#pragma unroll
    for (int k = 0; k < 5; ++k) {
      const float2 q = kFused ? sp[k] : __ldg(sp + k);
      const float2 r = __ldg(gp + k);
      s[2 * k] = q.x; s[2 * k + 1] = q.y;
      g[2 * k] = r.x; g[2 * k + 1] = r.y;
    }
    in.u[0] = __ldg(a.ac + i);
    if (kFused) {
      in.tt = a.t[i];
      in.stp = a.step[i];
    } else {
      in.tt = 0.f;
      in.stp = 0;
    }
    CarsRaw raw;
    assemble_cars(p, s, in.u[0], g, raw);
    normalise_rows<Pat, NZ, M>(raw.G, raw.h, nrm);
  }
  __device__ static __forceinline__ void finish(const Args& a, const Params& p, const EnvParams& e, int64_t i, Inst& in,
                                                const float xs[NU], int status) {
    const float us = clampf(in.u[0] + xs[0], p.u_min, p.u_max);  // diff_cbf_qp.py:77
    a.out[i] = us;
    if (a.status != nullptr) a.status[i] = status;
    if (kFused) {
      // the state is re-read here instead of being kept in 10 registers across the solve phase (L1/L2 hit)
      float s[10];
      float2* sp = reinterpret_cast<float2*>(a.state) + i * 5;
#pragma unroll
      for (int k = 0; k < 5; ++k) {
        const float2 q = sp[k];
        s[2 * k] = q.x; s[2 * k + 1] = q.y;
      }
      CarsEnvOut<float> o;
      cars_env_step<float>(e, s, in.tt, in.stp, us, o);
      float2* op = reinterpret_cast<float2*>(a.obs) + i * 5;
#pragma unroll
      for (int k = 0; k < 5; ++k) {
        op[k] = make_float2(o.obs[2 * k], o.obs[2 * k + 1]);
        sp[k] = make_float2(s[2 * k], s[2 * k + 1]);
      }
      a.reward[i] = o.reward;
      a.done[i] = (uint8_t)o.done;
      a.cost[i] = o.cost;
      a.t[i] = in.tt;
      a.step[i] = in.stp;
    }
  }
};

// ---------------------------------------------------------------------------------------------------------------
// shared helpers
// ---------------------------------------------------------------------------------------------------------------
template <class E>
__device__ __forceinline__ void pack_problem(const Normalised<E::NZ, E::M>& nrm, float w[E::NW]) {
  int k = 0;
#pragma unroll
  for (int i = 0; i < E::M; ++i) {
#pragma unroll
    for (int j = 0; j < E::NZ; ++j)
      if (E::Pat::nz(i, j)) w[k++] = nrm.Gn[i][j];
  }
#pragma unroll
  for (int i = 0; i < E::M; ++i) w[k++] = nrm.hn[i];
}

template <class E>
__device__ __forceinline__ void unpack_problem(const float w[E::NW], Normalised<E::NZ, E::M>& nrm) {
  int k = 0;
#pragma unroll
  for (int i = 0; i < E::M; ++i) {
#pragma unroll
    for (int j = 0; j < E::NZ; ++j) nrm.Gn[i][j] = E::Pat::nz(i, j) ? w[k++] : 0.f;
  }
#pragma unroll
  for (int i = 0; i < E::M; ++i) nrm.hn[i] = w[k++];
}

template <class E>
__device__ __forceinline__ void write_saved(const typename E::Args& a, int64_t i, const NormSolution<E::NZ, E::M>& sol) {
  if (a.x != nullptr) {
#pragma unroll
    for (int j = 0; j < E::NZ; ++j) a.x[i * E::NZ + j] = (float)sol.x[j];
  }
  if (a.lam != nullptr) {
#pragma unroll
    for (int r = 0; r < E::M; ++r) a.lam[i * E::M + r] = (float)sol.lam[r];
  }
  if (a.slack != nullptr) {
#pragma unroll
    for (int r = 0; r < E::M; ++r) a.slack[i * E::M + r] = (float)sol.s[r];
  }
  if (a.iters != nullptr) a.iters[i] = sol.iters;
}

__device__ __forceinline__ void block_counters(rcbf_counters_t* ws, bool valid, int status, int iters) {
  if (ws == nullptr) return;
  const int n_nan = __syncthreads_count(valid && status == RCBF_NAN);
  const int n_triv = __syncthreads_count(valid && status == RCBF_OK_TRIVIAL);
  const int n_pend = __syncthreads_count(valid && status == RCBF_PENDING);
  __shared__ int s_it;
  if (threadIdx.x == 0) s_it = 0;
  __syncthreads();
  int it = (valid && status != RCBF_PENDING) ? iters : 0;
  it = __reduce_add_sync(0xffffffffu, it);
  if ((threadIdx.x & 31) == 0 && it) atomicAdd(&s_it, it);
  __syncthreads();
  if (threadIdx.x == 0) {
    if (n_nan) atomicAdd(&ws[0], (unsigned long long)n_nan);
    if (n_triv) atomicAdd(&ws[3], (unsigned long long)n_triv);
    if (s_it) atomicAdd(&ws[4], (unsigned long long)s_it);
    if (n_pend) atomicAdd(&ws[5], (unsigned long long)n_pend);
  }
}

// ---------------------------------------------------------------------------------------------------------------
// pass 1
// ---------------------------------------------------------------------------------------------------------------
template <class E, int kMode /* 0 presolve, 1 pdipm */>
__global__ void __launch_bounds__(kBlock, kMode == 0 ? 2 : 1)
k_safe(typename E::Args a, int64_t n, typename E::Params p, typename E::EnvParams e, rcbf_counters_t* ws) {
  constexpr int NZ = E::NZ, M = E::M, NU = E::NU, NW = E::NW;
  __shared__ float s_prob[NW][kBlock];
  __shared__ float s_res[NU][kBlock];
  __shared__ int s_stat[kBlock];
  __shared__ unsigned short s_owner[kBlock];
  __shared__ int s_count;

  const int tid = threadIdx.x;
  const int64_t base = (int64_t)blockIdx.x * kBlock;
  const bool valid = base + tid < n;
  const int64_t i = valid ? base + tid : n - 1;
  if (tid == 0) s_count = 0;

  // ---- phase A
  typename E::Inst in;
  bool triv = true, nan = false;
  {
    Normalised<NZ, M> nrm;
    E::assemble(a, p, i, in, nrm);
#pragma unroll
    for (int r = 0; r < M; ++r) {
      triv = triv && (nrm.hn[r] >= 0.f);
      nan = nan || (nrm.hn[r] != nrm.hn[r]);
#pragma unroll
      for (int j = 0; j < NZ; ++j)
        if (E::Pat::nz(r, j)) nan = nan || (nrm.Gn[r][j] != nrm.Gn[r][j]);
    }
    const bool need = valid && !triv && !nan;
    __syncthreads();  // s_count = 0 visible
    const unsigned ballot = __ballot_sync(0xffffffffu, need);
    int wbase = 0;
    if ((tid & 31) == 0 && ballot) wbase = atomicAdd(&s_count, __popc(ballot));
    wbase = __shfl_sync(0xffffffffu, wbase, 0);
    if (need) {
      const int slot = wbase + __popc(ballot & ((1u << (tid & 31)) - 1u));
      float w[NW];
      pack_problem<E>(nrm, w);
#pragma unroll
      for (int k = 0; k < NW; ++k) s_prob[k][slot] = w[k];
      s_owner[slot] = (unsigned short)tid;
    }
    if (valid && !need && (a.x != nullptr || a.lam != nullptr || a.slack != nullptr || a.iters != nullptr)) {
      NormSolution<NZ, M> sol;  // trivial / NaN instance: x = 0 (NaN), lam = 0, slack = h~
#pragma unroll
      for (int j = 0; j < NZ; ++j) sol.x[j] = nan ? (double)NAN : 0.0;
#pragma unroll
      for (int r = 0; r < M; ++r) {
        sol.lam[r] = 0.0;
        sol.s[r] = (double)nrm.hn[r];
      }
      sol.iters = 0;
      write_saved<E>(a, i, sol);
    }
  }
  __syncthreads();

  // ---- phase B: compacted solve
  if (tid < s_count) {
    float w[NW];
#pragma unroll
    for (int k = 0; k < NW; ++k) w[k] = s_prob[k][tid];
    Normalised<NZ, M> nrm;
    unpack_problem<E>(w, nrm);
    NormSolution<NZ, M> sol;
    solve_normalised_fast<typename E::Pat, NZ, M, kMode == 0>(nrm, p.p_diag, sol);
    const int owner = s_owner[tid];
#pragma unroll
    for (int c = 0; c < NU; ++c) s_res[c][owner] = (float)sol.x[c];
    s_stat[owner] = sol.status | (sol.iters << 8);
    if (sol.status != RCBF_PENDING) write_saved<E>(a, base + owner, sol);
  }
  __syncthreads();

  // ---- phase C
  float xs[NU];
  int status = nan ? RCBF_NAN : RCBF_OK_TRIVIAL, iters = 0;
  if (valid && !triv && !nan) {
#pragma unroll
    for (int c = 0; c < NU; ++c) xs[c] = s_res[c][tid];
    status = s_stat[tid] & 255;
    iters = s_stat[tid] >> 8;
  } else {
#pragma unroll
    for (int c = 0; c < NU; ++c) xs[c] = nan ? NAN : 0.f;
  }
  if (valid) {
    if (status == RCBF_PENDING) {
      a.out[i * NU] = __uint_as_float(kPendingBits);
      if (ws != nullptr) {
        const unsigned long long slot = atomicAdd(&ws[kWsQueueCount], 1ULL);
        if (slot < (unsigned long long)kWsQueueCap) ws[kWsQueueBase + slot] = (unsigned long long)i;
      }
    } else {
      E::finish(a, p, e, i, in, xs, status);
    }
  }
  block_counters(ws, valid, status, iters);
}

// ---------------------------------------------------------------------------------------------------------------
// pass 2: interior-point fallback for the queued (or sentinel-marked) instances
// ---------------------------------------------------------------------------------------------------------------
template <class E, bool kSkipF32>
__device__ __forceinline__ void fallback_one(const typename E::Args& a, int64_t i, const typename E::Params& p,
                                             const typename E::EnvParams& e, rcbf_counters_t* ws) {
  constexpr int NZ = E::NZ, M = E::M, NU = E::NU;
  typename E::Inst in;
  Normalised<NZ, M> nrm;
  E::assemble(a, p, i, in, nrm);
  NormSolution<NZ, M> sol;
  solve_normalised_full<typename E::Pat, NZ, M>(nrm, p.p_diag, kSkipF32, sol);
  float xs[NU];
#pragma unroll
  for (int c = 0; c < NU; ++c) xs[c] = (float)sol.x[c];
  write_saved<E>(a, i, sol);
  E::finish(a, p, e, i, in, xs, sol.status);
  if (ws != nullptr) {
    if (sol.status == RCBF_NAN) atomicAdd(&ws[0], 1ULL);
    if (sol.status == RCBF_MAXITER) atomicAdd(&ws[1], 1ULL);
    if (sol.iters >= 100) atomicAdd(&ws[2], 1ULL);
    atomicAdd(&ws[6], (unsigned long long)(sol.iters >= 100 ? sol.iters - 100 : sol.iters));
  }
}

template <class E, bool kSkipF32>
__global__ void __launch_bounds__(128)
k_safe_fallback(typename E::Args a, int64_t n, typename E::Params p, typename E::EnvParams e, rcbf_counters_t* ws) {
  const int64_t gtid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int64_t gsz = (int64_t)gridDim.x * blockDim.x;
  bool scan = (ws == nullptr);
  if (!scan) {
    const unsigned long long cnt = ws[kWsQueueCount];
    if (cnt > (unsigned long long)kWsQueueCap) {
      scan = true;  // overflow: every pending instance still carries the sentinel
    } else {
      for (int64_t q = gtid; q < (int64_t)cnt; q += gsz) fallback_one<E, kSkipF32>(a, (int64_t)ws[kWsQueueBase + q], p, e, ws);
    }
  }
  if (scan) {
    for (int64_t i = gtid; i < n; i += gsz)
      if (__float_as_uint(__ldcg(a.out + i * E::NU)) == kPendingBits) fallback_one<E, kSkipF32>(a, i, p, e, ws);
  }
  if (ws != nullptr) {  // last block resets the queue for the next call
    __syncthreads();
    if (threadIdx.x == 0) {
      __threadfence();
      const unsigned long long d = atomicAdd(&ws[kWsBlocksDone], 1ULL);
      if (d == gridDim.x - 1) {
        ws[kWsQueueCount] = 0ULL;
        ws[kWsBlocksDone] = 0ULL;
        __threadfence();
      }
    }
  }
}

template <class E>
inline int launch_safe(const typename E::Args& a, int64_t n, const typename E::Params& p, const typename E::EnvParams& e,
                       rcbf_counters_t* ws, cudaStream_t s) {
  if (n <= 0) return 0;
  const int grid = (int)((n + kBlock - 1) / kBlock);
  // fallback grid: the queue holds at most kWsQueueCap entries; without a workspace it must scan all n
  const int fgrid = (ws != nullptr) ? 16 : (int)((n + 127) / 128 < 148 * 8 ? (n + 127) / 128 : 148 * 8);
  if (p.solver_mode == 0) {
    k_safe<E, 0><<<grid, kBlock, 0, s>>>(a, n, p, e, ws);
    k_safe_fallback<E, false><<<fgrid, 128, 0, s>>>(a, n, p, e, ws);
  } else {
    k_safe<E, 1><<<grid, kBlock, 0, s>>>(a, n, p, e, ws);
    k_safe_fallback<E, true><<<fgrid, 128, 0, s>>>(a, n, p, e, ws);
  }
  return (int)cudaGetLastError();
}

}  // namespace rcbf
