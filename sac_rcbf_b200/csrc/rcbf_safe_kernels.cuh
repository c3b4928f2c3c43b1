// rcbf_safe_kernels.cuh -- the hot kernels: get_safe_action forward (K2+K3) and the fused safe step (K5).
//
// One lane owns one instance; warps are persistent and walk 32-instance tiles (coalesced row-major loads).
// The inputs of the NEXT tile arrive by TMA bulk copies (cp.async.bulk + mbarrier, double buffered) while the warp
// works.  Per tile ("A-step") every lane assembles the constraints in the reference's float32 op order and runs the
// trivial test on the raw rows (h >= 0 on every row  <=>  x = 0 is optimal, ~2/3 of the synthetic instances); trivial
// lanes finish right there (clamp, env.step, outputs).  Lanes whose instance needs a solve push its raw rows (19 words
// Unicycle / 6 words SimulatedCars, the constant entries are not stored, + the instance index) into a WARP-PRIVATE
// ring in shared memory.  Whenever the ring holds >= 32 problems the warp runs a "B-step": all 32 lanes normalise and
// solve one problem each (greedy active-set presolve + float64 KKT certificate, or the float32 interior point in
// "pdipm" mode).  Unicycle, presolve mode ("merged finish"): the solving lane puts (index, correction, status, sin, cos)
// into a warp-private finish ring and starts cp.async copies of the instance's 28 bytes of finish inputs into it; in
// the finish pass of a LATER tile every lane whose own instance went to the problem ring takes one entry, so the
// finish code runs with (nearly) full warps too.  The SimulatedCars layer kernel and pdipm mode keep the earlier
// scheme: the lanes reload their instance after the next A-step and run a second finish pass.  So the expensive phase
// always runs with full warps (without compaction it ran at ~35 % lane utilisation), and there is no block-wide barrier
// anywhere: warps never wait for each other (until the very end of the kernel).
//
// SimulatedCars fused step: no ring at all.  Its 2 x 4 QP is cheap and the kernel is bound by memory requests, so the
// lanes that need a solve do it inline in the A-step (partial lane utilisation) and the whole tile -- every instance of
// it -- is finished through the coalesced warp-collective path (`finish_tile`); measured +37 % over the ring version,
// whose solved third came back in a second finish pass with lane-per-row 40-byte accesses.
//
// Instances a B-step cannot certify (a constraint would have to be dropped, borderline degeneracy: ~1.5e-5 of the
// Unicycle instances; SimulatedCars enumerates its 10 candidate active sets inline and leaves none) get a tagged-NaN
// sentinel in safe_action[i][0] and are queued in the caller's workspace.  "presolve" mode with a workspace: a warp
// that has run out of tiles drains that queue -- one WARP per instance enumerates every active set of size <= nz (129
// candidates) with the same float64 certificate, the interior point being the last resort -- so the step is ONE
// kernel launch; the last block to finish resets the queue (and, should it ever overflow, scans for the sentinel).
// Pass 2 (k_safe_fallback) remains for the other cases: without a workspace it scans for the sentinel, and in "pdipm"
// mode one thread per queued instance runs the float64 interior point.
#pragma once

#include <cstdlib>

#include <cuda_runtime.h>

#include "rcbf_core.cuh"
#include "rcbf_dynamics.cuh"
#include "rcbf_tma.cuh"

namespace rcbf {

#ifndef RCBF_MINB
#define RCBF_MINB 4  // resident blocks per SM the presolve-mode kernel is compiled for (128 registers; A/B on B200: 4 > 3 > 5)
#endif
#ifndef RCBF_MINB_PDIPM
#define RCBF_MINB_PDIPM 2  // interior-point ("pdipm") mode: the iteration state wants the registers
#endif
#ifndef RCBF_MINB_CARS
#define RCBF_MINB_CARS 4  // SimulatedCars fused step (inline solve, whole-tile finish): 4 (128 registers) > 5 > 3 >> 6 on B200
#endif                    // (with the earlier problem-ring version of that kernel 3 was best: it was request-bound)


// A/B switches (defaults = what measured best on B200; the alternatives are kept so the comparison can be re-run)
#ifndef RCBF_MERGE_FINISH
#define RCBF_MERGE_FINISH 1      // Unicycle presolve: solved instances are finished by idle lanes of later tiles
#endif
#ifndef RCBF_RING_VEC
#define RCBF_RING_VEC 1          // slot-major float4 problem ring (0: word-major scalar ring)
#endif
#ifndef RCBF_CARS_TILE_FINISH
#define RCBF_CARS_TILE_FINISH 1  // fused SimulatedCars step: coalesced whole-tile finish
#endif
#ifndef RCBF_CARS_INLINE
#define RCBF_CARS_INLINE 1       // fused SimulatedCars step: solve inline in the A-step (no problem ring)
#endif
#ifndef RCBF_CARS_INLINE_ALL
#define RCBF_CARS_INLINE_ALL 0   // ... also in the SimulatedCars layer-only kernel (measured: -27 %)
#endif

constexpr uint32_t kPendingBits = 0x7fc0dead;  // quiet NaN with a payload no arithmetic produces
constexpr int kWsCounters = 8;                 // workspace words [0, 8): counters
constexpr int kWsQueueCount = 8;               // [8]: number of queued instances
constexpr int kWsBlocksDone = 9;               // [9]: blocks finished (the last one resets the queue)
constexpr int kWsClaim = 10;                   // [10]: queue entries claimed by a draining warp (presolve mode)
constexpr int kWsMirror = 11;                  // [11]: host mirror bound by rcbf_counters_bind_mirror (0: none)
constexpr int kWsQueueBase = 16;               // [16, RCBF_WS_WORDS): queued instance indices + 1 (0 = empty slot)
constexpr int kWsQueueCap = RCBF_WS_WORDS - kWsQueueBase;

template <int K>
__device__ __forceinline__ void ld_row(const float* __restrict__ base, int64_t i, float out[K]) {
#pragma unroll
  for (int j = 0; j < K; ++j) out[j] = __ldg(base + i * K + j);
}
__device__ __forceinline__ void prefetch_l1(const void* p) { asm volatile("prefetch.global.L1 [%0];" ::"l"(p)); }

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
  int32_t* meta;     // nullable: (status << 16) | active-set mask -- all the backward pass needs saved
  float* obs;        // (n,7)  [safe_step]
  float* reward;
  uint8_t* done;
  float* cost;
  uint8_t* goal_met;
};

template <bool kFused>
struct UniEnv {
  static constexpr int NZ = kUniNZ, M = kUniM, NU = 2;
  static constexpr bool kTileFinish = false;
  static constexpr int kMinBlocks = RCBF_MINB;
  static constexpr bool kPdlPass1 = true;   // see launch_pdl: the presolve kernel fills the register file exactly
  using Pat = UniPat;
  using Args = UniArgs;
  using Params = UnicycleParams;
  using EnvParams = UnicycleEnvParams;
  struct Inst {
    float v[4];  // x, y, theta, last_goal_dist (fused) / unused
    float u[2];
    float sn, cs;  // sin / cos of theta: computed once, used by the assembly AND by env.step
    int stp;
  };
  __device__ static __forceinline__ void load_inst(const Args& a, int64_t i, Inst& in) {
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
    sincos_t(in.v[2], &in.sn, &in.cs);
  }
  // Merged finish (presolve mode): a solved instance is finished by a lane that is idle in a LATER tile's finish pass
  // (a lane whose own instance went to the ring), so the finish code runs once per tile instead of 1.35 times.  The
  // lane that solved it starts cp.async copies of the instance's finish inputs into the finish ring (they land during
  // the next A-step: no register is held and nobody waits for the L2 round trip); the sin / cos of the heading travel
  // with the problem through the rings instead of being recomputed.
  static constexpr bool kMergeFinish = RCBF_MERGE_FINISH;
  static constexpr bool kInlineSolve = false;
  static constexpr int NSC = 2;
  __device__ static __forceinline__ void stash(const Inst& in, float s[NSC]) {
    s[0] = in.sn;
    s[1] = in.cs;
  }
  template <int K>
  struct alignas(16) FinRing {   // one entry = 3 x 128 bits + the step counter
    float4 st[kFused ? K : 1];   // x, y, theta, last goal distance        (cp.async)
    float4 res[K];               // correction x0, x1, instance index, status
    float4 acsc[K];              // nominal action (cp.async), sin, cos
    int step[kFused ? K : 1];    //                                        (cp.async)
  };
  template <int K, bool kAc8>
  __device__ static __forceinline__ void fin_fetch(const Args& a, int i, FinRing<K>& f, int fs, const float s[NSC],
                                                   const float xs[NU], int status) {
    if (kFused) {
      cp_async<16>(&f.st[fs], a.state4 + (int64_t)i * 4);
      cp_async<4>(&f.step[fs], a.step + i);
    }
    float* acsc = reinterpret_cast<float*>(&f.acsc[fs]);
    if (kAc8) {
      cp_async<8>(acsc, a.ac + (int64_t)i * 2);
    } else {
      cp_async<4>(acsc, a.ac + (int64_t)i * 2);
      cp_async<4>(acsc + 1, a.ac + (int64_t)i * 2 + 1);
    }
    reinterpret_cast<float2*>(acsc)[1] = make_float2(s[0], s[1]);
    f.res[fs] = make_float4(xs[0], xs[1], __int_as_float(i), __int_as_float(status));
  }
  template <int K>
  __device__ static __forceinline__ void fin_read(const FinRing<K>& f, int fs, Inst& in, float xs[NU], int64_t& i,
                                                  int& status) {
    if (kFused) {
      const float4 q = f.st[fs];
      in.v[0] = q.x; in.v[1] = q.y; in.v[2] = q.z; in.v[3] = q.w;
      in.stp = f.step[fs];
    } else {
      in.v[0] = in.v[1] = in.v[2] = in.v[3] = 0.f;  // the layer's finish only clamps action + correction
      in.stp = 0;
    }
    const float4 r = f.res[fs], c = f.acsc[fs];
    xs[0] = r.x; xs[1] = r.y;
    i = __float_as_int(r.z);
    status = __float_as_int(r.w);
    in.u[0] = c.x; in.u[1] = c.y;
    in.sn = c.z;
    in.cs = c.w;
  }
  // raw rows in NWR = 19 words: (G[i][0], G[i][1]) of the 5 CBF rows + the 9 h; the rest of G is constant
  static constexpr int NWR = 19;
  struct Aux {  // the inputs only the assembly reads (loaded one tile ahead by the persistent loop)
    float m[3], g[3];
  };
  __device__ static __forceinline__ void load_aux(const Args& a, int64_t i, Aux& x) {
    ld_row<3>(a.mu, i, x.m);
    ld_row<3>(a.sg, i, x.g);
  }
  struct alignas(16) Stage {  // one tile of inputs as the bulk copies land them (same row-major layout as HBM)
    float st[32 * 4];      // fused: (x,y,th,last) x 32 ; plain: (x,y,th) x 32 in the first 96 words
    float ac[32 * 2];
    float mu[32 * 3];
    float sg[32 * 3];
    int step[32];
  };
  static constexpr uint32_t kStageBytes = (kFused ? 512 + 128 : 384) + 256 + 384 + 384;
  __host__ __device__ static __forceinline__ bool aligned(const Args& a) {
    auto ok = [](const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; };
    return ok(kFused ? (const void*)a.state4 : (const void*)a.st) && ok(a.ac) && ok(a.mu) && ok(a.sg) &&
           (!kFused || ok(a.step)) && (reinterpret_cast<uintptr_t>(a.out) & 7) == 0;
  }
  __device__ static __forceinline__ void issue(const Args& a, int64_t tile, Stage& sg_, uint64_t* bar) {
    const int64_t i0 = tile << 5;
    mbar_expect_tx(bar, kStageBytes);
    if (kFused) {
      bulk_g2s(sg_.st, a.state4 + i0 * 4, 512, bar);
      bulk_g2s(sg_.step, a.step + i0, 128, bar);
    } else {
      bulk_g2s(sg_.st, a.st + i0 * 3, 384, bar);
    }
    bulk_g2s(sg_.ac, a.ac + i0 * 2, 256, bar);
    bulk_g2s(sg_.mu, a.mu + i0 * 3, 384, bar);
    bulk_g2s(sg_.sg, a.sg + i0 * 3, 384, bar);
  }
  __device__ static __forceinline__ void read_stage(const Stage& sg_, int lane, Inst& in, Aux& x) {
    if (kFused) {
      const float4 q = reinterpret_cast<const float4*>(sg_.st)[lane];
      in.v[0] = q.x; in.v[1] = q.y; in.v[2] = q.z; in.v[3] = q.w;
      in.stp = sg_.step[lane];
    } else {
#pragma unroll
      for (int j = 0; j < 3; ++j) in.v[j] = sg_.st[lane * 3 + j];
      in.v[3] = 0.f;
      in.stp = 0;
    }
    const float2 u2 = reinterpret_cast<const float2*>(sg_.ac)[lane];
    in.u[0] = u2.x; in.u[1] = u2.y;
    sincos_t(in.v[2], &in.sn, &in.cs);
#pragma unroll
    for (int j = 0; j < 3; ++j) {
      x.m[j] = sg_.mu[lane * 3 + j];
      x.g[j] = sg_.sg[lane * 3 + j];
    }
  }
  __device__ static __forceinline__ void prefetch(const Args& a, int64_t i) {
    if (kFused) {
      prefetch_l1(a.state4 + i * 4);
      prefetch_l1(a.step + i);
    } else {
      prefetch_l1(a.st + i * 3);
    }
    prefetch_l1(a.ac + i * 2);
    prefetch_l1(a.mu + i * 3);
    prefetch_l1(a.sg + i * 3);
  }
  __device__ static __forceinline__ void assemble_raw(const Params& p, const Inst& in, const Aux& x, float w[NWR],
                                                      bool& triv, bool& nan) {
    UniRaw raw;
    assemble_unicycle_sc(p, in.v, in.sn, in.cs, in.u, x.m, x.g, raw);
    classify_raw<M>(raw.h, triv, nan);
#pragma unroll
    for (int r = 0; r < kUniHaz; ++r) {
      w[2 * r] = raw.G[r][0];
      w[2 * r + 1] = raw.G[r][1];
    }
    nan = nan || any_nan<2 * kUniHaz>(w);
#pragma unroll
    for (int r = 0; r < M; ++r) w[10 + r] = raw.h[r];
  }
  __device__ static __forceinline__ void unpack_raw(const float w[NWR], const Params&, float G[M][NZ], float h[M]) {
#pragma unroll
    for (int r = 0; r < kUniHaz; ++r) {
      G[r][0] = w[2 * r];
      G[r][1] = w[2 * r + 1];
      G[r][2] = -1.0f;  // diff_cbf_qp.py:260
    }
#pragma unroll
    for (int c = 0; c < 2; ++c) {  // diff_cbf_qp.py:365-377
#pragma unroll
      for (int j = 0; j < 3; ++j) {
        G[kUniHaz + 2 * c][j] = (j == c) ? 1.0f : 0.0f;
        G[kUniHaz + 2 * c + 1][j] = (j == c) ? -1.0f : 0.0f;
      }
    }
#pragma unroll
    for (int r = 0; r < M; ++r) h[r] = w[10 + r];
  }
  __device__ static __forceinline__ void normalise_packed(const float w[NWR], const Params& p, Normalised<NZ, M>& nrm) {
    float G[M][NZ], h[M];
    unpack_raw(w, p, G, h);
    normalise_rows<Pat, NZ, M>(G, h, nrm);
  }
  __device__ static __forceinline__ void assemble(const Args& a, const Params& p, int64_t i, const Inst& in,
                                                  Normalised<NZ, M>& nrm) {
    float w[NWR];
    bool triv, nan;
    Aux x;
    load_aux(a, i, x);
    assemble_raw(p, in, x, w, triv, nan);
    normalise_packed(w, p, nrm);
  }
  template <bool kOut8 = false>  // kOut8: safe_action rows are 8-byte aligned (part of `aligned`), one 64-bit store
  __device__ static __forceinline__ void finish(const Args& a, const Params& p, const EnvParams& e, int64_t i, Inst& in,
                                                const float xs[NU], int status) {
    float us[2];
#pragma unroll
    for (int c = 0; c < 2; ++c) us[c] = clampf(in.u[c] + xs[c], p.u_min[c], p.u_max[c]);  // diff_cbf_qp.py:77
    if (kOut8) reinterpret_cast<float2*>(a.out)[i] = make_float2(us[0], us[1]);
    else st_row<2>(a.out, i, us);
    if (a.status != nullptr) a.status[i] = status;
    if (kFused) {
      UniEnvOut<float> o;
      unicycle_env_step_sc(e, in.v, in.v[3], in.stp, us, in.sn, in.cs, o);
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
  int32_t* meta;    // nullable: (status << 16) | active-set mask
  float* obs;       // (n,10)
  float* reward;
  uint8_t* done;
  float* cost;
};

template <bool kFused>
struct CarsEnv {
  static constexpr int NZ = kCarsNZ, M = kCarsM, NU = 1;
  static constexpr bool kTileFinish = kFused && RCBF_CARS_TILE_FINISH;  // see finish_tile
  static constexpr int kMinBlocks = kFused ? RCBF_MINB_CARS : RCBF_MINB;  // get_safe_action alone: 4 is faster (A/B)
  static constexpr bool kPdlPass1 = false;  // measured: -16 % when pass 1 is launched as a dependent
  using Pat = CarsPat;
  using Args = CarsArgs;
  using Params = CarsParams;
  using EnvParams = CarsEnvParams;
  struct Inst {
    float u[1];
    float tt;
    int stp;
  };
  __device__ static __forceinline__ void load_inst(const Args& a, int64_t i, Inst& in) {
    in.u[0] = __ldg(a.ac + i);
    if (kFused) {
      in.tt = a.t[i];
      in.stp = a.step[i];
    } else {
      in.tt = 0.f;
      in.stp = 0;
    }
  }
  static constexpr bool kMergeFinish = false;  // the fused step finishes whole tiles (finish_tile)
  static constexpr bool kInlineSolve = (kFused || RCBF_CARS_INLINE_ALL) && RCBF_CARS_INLINE;  // see k_safe: solve in the A-step, no ring
  static constexpr int NSC = 1;
  __device__ static __forceinline__ void stash(const Inst&, float s[NSC]) { s[0] = 0.f; }
  template <int K>
  struct FinRing {
    int unused;
  };
  template <int K, bool kAc8>
  __device__ static __forceinline__ void fin_fetch(const Args&, int, FinRing<K>&, int, const float[NSC], const float[NU],
                                                   int) {}
  template <int K>
  __device__ static __forceinline__ void fin_read(const FinRing<K>&, int, Inst&, float[NU], int64_t&, int&) {}
  static constexpr int NWR = 6;  // G[0][0], G[1][0] + the 4 h; the slack column and the actuator rows are constant
  struct Aux {
    float s[10], g[10];
  };
  __device__ static __forceinline__ void load_aux(const Args& a, int64_t i, Aux& x) {
    const float2* sp = reinterpret_cast<const float2*>(kFused ? a.state : a.st) + i * 5;  // rows are 40 B: 8-aligned
    const float2* gp = reinterpret_cast<const float2*>(a.sg) + i * 5;
#pragma unroll
    for (int k = 0; k < 5; ++k) {
      const float2 q = kFused ? sp[k] : __ldg(sp + k);
      const float2 r = __ldg(gp + k);
      x.s[2 * k] = q.x; x.s[2 * k + 1] = q.y;
      x.g[2 * k] = r.x; x.g[2 * k + 1] = r.y;
    }
  }
  struct alignas(16) Stage {
    float st[32 * 10];
    float sg[32 * 10];
    float ac[32];
    float t[32];
    int step[32];
  };
  static constexpr uint32_t kStageBytes = 1280 + 1280 + 128 + (kFused ? 256 : 0);
  __host__ __device__ static __forceinline__ bool aligned(const Args& a) {
    auto ok = [](const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; };
    return ok(kFused ? (const void*)a.state : (const void*)a.st) && ok(a.ac) && ok(a.sg) &&
           (!kFused || (ok(a.t) && ok(a.step)));
  }
  __device__ static __forceinline__ void issue(const Args& a, int64_t tile, Stage& sg_, uint64_t* bar) {
    const int64_t i0 = tile << 5;
    mbar_expect_tx(bar, kStageBytes);
    bulk_g2s(sg_.st, (kFused ? a.state : a.st) + i0 * 10, 1280, bar);
    bulk_g2s(sg_.sg, a.sg + i0 * 10, 1280, bar);
    bulk_g2s(sg_.ac, a.ac + i0, 128, bar);
    if (kFused) {
      bulk_g2s(sg_.t, a.t + i0, 128, bar);
      bulk_g2s(sg_.step, a.step + i0, 128, bar);
    }
  }
  __device__ static __forceinline__ void read_stage(const Stage& sg_, int lane, Inst& in, Aux& x) {
    const float2* sp = reinterpret_cast<const float2*>(sg_.st) + lane * 5;
    const float2* gp = reinterpret_cast<const float2*>(sg_.sg) + lane * 5;
#pragma unroll
    for (int k = 0; k < 5; ++k) {
      const float2 q = sp[k], r = gp[k];
      x.s[2 * k] = q.x; x.s[2 * k + 1] = q.y;
      x.g[2 * k] = r.x; x.g[2 * k + 1] = r.y;
    }
    in.u[0] = sg_.ac[lane];
    in.tt = kFused ? sg_.t[lane] : 0.f;
    in.stp = kFused ? sg_.step[lane] : 0;
  }
  __device__ static __forceinline__ void prefetch(const Args& a, int64_t i) {
    const float* sp = (kFused ? a.state : a.st) + i * 10;
    prefetch_l1(sp);
    prefetch_l1(sp + 8);
    prefetch_l1(a.sg + i * 10);
    prefetch_l1(a.sg + i * 10 + 8);
    prefetch_l1(a.ac + i);
    if (kFused) {
      prefetch_l1(a.t + i);
      prefetch_l1(a.step + i);
    }
  }
  __device__ static __forceinline__ void assemble_raw(const Params& p, const Inst& in, const Aux& x, float w[NWR],
                                                      bool& triv, bool& nan) {
    CarsRaw raw;
    assemble_cars(p, x.s, in.u[0], x.g, raw);
    classify_raw<M>(raw.h, triv, nan);
    w[0] = raw.G[0][0];
    w[1] = raw.G[1][0];
    nan = nan || any_nan<2>(w);
#pragma unroll
    for (int r = 0; r < M; ++r) w[2 + r] = raw.h[r];
  }
  __device__ static __forceinline__ void unpack_raw(const float w[NWR], const Params& p, float G[M][NZ], float h[M]) {
    G[0][0] = w[0];
    G[1][0] = w[1];
    G[0][1] = -p.slack_coeff;  // diff_cbf_qp.py:352
    G[1][1] = -p.slack_coeff;
    G[2][0] = 1.0f;  G[2][1] = 0.0f;   // :369-370
    G[3][0] = -1.0f; G[3][1] = 0.0f;   // :375-376
#pragma unroll
    for (int r = 0; r < M; ++r) h[r] = w[2 + r];
  }
  __device__ static __forceinline__ void normalise_packed(const float w[NWR], const Params& p, Normalised<NZ, M>& nrm) {
    float G[M][NZ], h[M];
    unpack_raw(w, p, G, h);
    normalise_rows<Pat, NZ, M>(G, h, nrm);
  }
  __device__ static __forceinline__ void assemble(const Args& a, const Params& p, int64_t i, const Inst& in,
                                                  Normalised<NZ, M>& nrm) {
    float w[NWR];
    bool triv, nan;
    Aux x;
    load_aux(a, i, x);
    assemble_raw(p, in, x, w, triv, nan);
    normalise_packed(w, p, nrm);
  }
  template <bool kOut8 = false>
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
  // Warp-collective finish of the lanes that finish their OWN instance of a full, bulk-staged tile (job A).  The rows of
  // `state` and `obs` are 40 bytes: a lane-per-row access pattern touches every 32-byte sector of the 1280-byte tile
  // span five times (five 8-byte accesses at a 40-byte stride), and that request count -- not bytes, not latency --
  // is what bounds this kernel.  Here the input row comes from the TMA landing buffer, the new state / observation
  // rows go back into that buffer (the sigma half is free by now) and leave through ten fully coalesced, owner-masked
  // 128-byte stores each: 1/5 of the sector requests.  `on` lanes only; all 32 lanes must call.
  __device__ static __forceinline__ void finish_tile(const Args& a, const Params& p, const EnvParams& e, int64_t i0,
                                                     int lane, bool on, Inst& in, const float xs[NU], int status,
                                                     Stage& sg_) {
    float s[10];
    CarsEnvOut<float> o;
    const int64_t i = i0 + lane;
    if (on) {
      const float us = clampf(in.u[0] + xs[0], p.u_min, p.u_max);  // diff_cbf_qp.py:77
      a.out[i] = us;
      if (a.status != nullptr) a.status[i] = status;
      const float2* sp = reinterpret_cast<const float2*>(sg_.st) + lane * 5;
#pragma unroll
      for (int k = 0; k < 5; ++k) {
        const float2 q = sp[k];
        s[2 * k] = q.x; s[2 * k + 1] = q.y;
      }
      cars_env_step<float>(e, s, in.tt, in.stp, us, o);
      a.reward[i] = o.reward;
      a.done[i] = (uint8_t)o.done;
      a.cost[i] = o.cost;
      a.t[i] = in.tt;
      a.step[i] = in.stp;
      float2* wp = reinterpret_cast<float2*>(sg_.st) + lane * 5;   // own row: nobody else reads or writes it
      float2* op = reinterpret_cast<float2*>(sg_.sg) + lane * 5;
#pragma unroll
      for (int k = 0; k < 5; ++k) {
        wp[k] = make_float2(s[2 * k], s[2 * k + 1]);
        op[k] = make_float2(o.obs[2 * k], o.obs[2 * k + 1]);
      }
    }
    const unsigned owners = __ballot_sync(0xffffffffu, on);
    __syncwarp();
    float* gs = a.state + i0 * 10;
    float* go = a.obs + i0 * 10;
#pragma unroll
    for (int r = 0; r < 10; ++r) {
      const int w = r * 32 + lane;
      const int owner = (w * 205) >> 11;  // w / 10 for w < 1024
      if ((owners >> owner) & 1u) {
        gs[w] = sg_.st[w];
        go[w] = sg_.sg[w];
      }
    }
    // the buffer is handed back to the async proxy (next bulk copy) only after these generic-proxy accesses
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    __syncwarp();
  }
};

// ---------------------------------------------------------------------------------------------------------------
// shared helpers
// ---------------------------------------------------------------------------------------------------------------
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
  if (a.meta != nullptr) a.meta[i] = (sol.status << 16) | (int)(sol.mask & 0xffffu);
}

// ---------------------------------------------------------------------------------------------------------------
// pass 1: persistent warps, warp-private compaction ring
// ---------------------------------------------------------------------------------------------------------------
constexpr int kWarps = 4;             // warps per block
constexpr int kThreadsW = 32 * kWarps;
// Per-warp shared memory.  Problem ring: raw rows + instance index of the QPs waiting for a solve.  Finish ring:
// (index, correction, status) of solved instances waiting for clamp / env.step.  Stage: two TMA landing buffers.
template <class E, int kMode>
struct WarpShared {
  static constexpr int kRing = kMode == 0 ? 64 : 128;  // presolve: <= 31 left over + 32 new; pdipm: engine starts at 64
  static constexpr bool kMerge = kMode == 0 && E::kMergeFinish;
  // finish ring: pdipm <= 31 left over + one solve phase's output; merged presolve: 64 (a full-warp pass drains it
  // whenever a B-step's output might not fit, see k_safe); plain presolve: unused
  static constexpr int kFin = kMode == 0 ? (kMerge ? 64 : 32) : 256;
  static constexpr int kScRing = kMerge ? kRing : 1;
#if RCBF_RING_VEC
  // problem ring, slot-major: NWR raw words + the instance index, padded to float4s and moved by 128-bit shared-memory
  // accesses (Unicycle: 5 per problem instead of 20 scalar ones; the 80-byte slot stride is conflict-free per
  // quarter warp)
  static constexpr int kW4 = (E::NWR + 1 + 3) / 4;
  float4 w4[kRing][kW4];
  __device__ __forceinline__ void push(int slot, const float (&w)[E::NWR], int idx) {
    float v[kW4 * 4];
#pragma unroll
    for (int k = 0; k < kW4 * 4; ++k) v[k] = k < E::NWR ? w[k] : 0.f;
    v[E::NWR] = __int_as_float(idx);
#pragma unroll
    for (int q = 0; q < kW4; ++q) w4[slot][q] = make_float4(v[4 * q], v[4 * q + 1], v[4 * q + 2], v[4 * q + 3]);
  }
  __device__ __forceinline__ void pop(int slot, float (&w)[E::NWR], int& idx) const {
    float v[kW4 * 4];
#pragma unroll
    for (int q = 0; q < kW4; ++q) {
      const float4 t = w4[slot][q];
      v[4 * q] = t.x; v[4 * q + 1] = t.y; v[4 * q + 2] = t.z; v[4 * q + 3] = t.w;
    }
#pragma unroll
    for (int k = 0; k < E::NWR; ++k) w[k] = v[k];
    idx = __float_as_int(v[E::NWR]);
  }
  __device__ __forceinline__ unsigned short* scratch() { return reinterpret_cast<unsigned short*>(&w4[0][0]); }
#else
  float w[E::NWR][kRing];
  int idx[kRing];
  __device__ __forceinline__ void push(int slot, const float (&w_)[E::NWR], int idx_) {
#pragma unroll
    for (int k = 0; k < E::NWR; ++k) w[k][slot] = w_[k];
    idx[slot] = idx_;
  }
  __device__ __forceinline__ void pop(int slot, float (&w_)[E::NWR], int& idx_) const {
#pragma unroll
    for (int k = 0; k < E::NWR; ++k) w_[k] = w[k][slot];
    idx_ = idx[slot];
  }
  __device__ __forceinline__ unsigned short* scratch() { return reinterpret_cast<unsigned short*>(&w[0][0]); }
#endif
  float sc[E::NSC][kScRing];   // merged finish: words that travel with the problem (Unicycle: sin, cos)
  static constexpr int kFinPlain = kMerge ? 1 : kFin;   // (index, correction, status) rings of the other modes
  float fx[E::NU][kFinPlain];
  int fidx[kFinPlain];
  int fst[kFinPlain];
  typename E::template FinRing<kMerge ? kFin : 1> fin;   // merged finish: the instance's finish inputs (cp.async)
  typename E::Stage stage[2];
  uint64_t bar[2];
};

// Programmatic dependent launch (PDL): a kernel launched with the programmatic-stream-serialization attribute has its
// launch latency and block scheduling overlapped with the tail of the kernel before it.  `pdl_wait` blocks until the
// preceding grid in the stream has completed and flushed (a no-op for an ordinary launch) and must come before the
// first read of anything an earlier kernel wrote; `pdl_launch_dependents` lets the dependent grid start being
// scheduled.  Pass 2 is always launched this way.  Pass 1 (as a dependent of the PREVIOUS call's pass 2) only where
// A/B runs on B200 showed a gain: the Unicycle presolve kernel (128 registers x 4 blocks = the whole register file,
// so its blocks can only be placed evenly as the old ones drain; +1.8 %).  The SimulatedCars kernel leaves room for
// its blocks to be placed early and unevenly next to pass-2 blocks and the persistent grid then runs imbalanced
// (-16 %), so it is launched normally.
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }

// solver mode proper (bits 0..7 of params.solver_mode; the rest is the publication request, include/rcbf_b200.h)
template <class P>
__host__ __device__ __forceinline__ int solver_mode_of(const P& p) { return p.solver_mode & 0xff; }

// Last block of a call's last kernel, one thread, after every block's counter atomics (each block fences before it
// arrives on ws[kWsBlocksDone]): counters -> the bound host mirror, then the caller's token.
__device__ __forceinline__ void publish_counters(rcbf_counters_t* ws, int solver_mode) {
  if ((solver_mode & RCBF_SOLVER_PUBLISH) == 0) return;
  volatile rcbf_counters_t* vws = ws;
  volatile unsigned long long* mirror = reinterpret_cast<volatile unsigned long long*>(vws[kWsMirror]);
  if (mirror == nullptr) return;
#pragma unroll
  for (int c = 0; c < 8; ++c) mirror[1 + c] = vws[c];
  __threadfence_system();
  mirror[0] = (unsigned long long)((unsigned)solver_mode >> RCBF_SOLVER_TOKEN_SHIFT);
}

template <class E>
__device__ __forceinline__ void mark_pending(const typename E::Args& a, int64_t i, rcbf_counters_t* ws) {
  a.out[i * E::NU] = __uint_as_float(kPendingBits);
  if (ws != nullptr) {
    __threadfence();  // the sentinel must be visible before whoever drains the queue writes the real result over it
    const unsigned long long slot = atomicAdd(&ws[kWsQueueCount], 1ULL);
    if (slot < (unsigned long long)kWsQueueCap) ws[kWsQueueBase + slot] = (unsigned long long)i + 1ULL;
  }
}

// ---------------------------------------------------------------------------------------------------------------
// pass 2: the queued (or sentinel-marked) instances
// ---------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ void fallback_counters(rcbf_counters_t* ws, int status, int iters) {
  if (ws == nullptr) return;
  if (status == RCBF_NAN) atomicAdd(&ws[0], 1ULL);
  if (status == RCBF_MAXITER) atomicAdd(&ws[1], 1ULL);
  if (iters >= 100) atomicAdd(&ws[2], 1ULL);
  atomicAdd(&ws[6], (unsigned long long)(iters >= 100 ? iters - 100 : iters));
}

// thread-per-instance interior-point chain (float32 unless it already failed in pass 1, then float64)
template <class E, bool kSkipF32>
__device__ __forceinline__ void fallback_ipm(const typename E::Args& a, int64_t i, const typename E::Params& p,
                                             const typename E::EnvParams& e, rcbf_counters_t* ws) {
  constexpr int NZ = E::NZ, M = E::M, NU = E::NU;
  typename E::Inst in;
  E::load_inst(a, i, in);
  Normalised<NZ, M> nrm;
  E::assemble(a, p, i, in, nrm);
  NormSolution<NZ, M> sol;
  solve_normalised_full<typename E::Pat, NZ, M>(nrm, p.p_diag, kSkipF32, sol);
  float xs[NU];
#pragma unroll
  for (int c = 0; c < NU; ++c) xs[c] = (float)sol.x[c];
  write_saved<E>(a, i, sol);
  E::finish(a, p, e, i, in, xs, sol.status);
  fallback_counters(ws, sol.status, sol.iters);
}

// warp-per-instance exhaustive active-set enumeration with the float64 certificate ("presolve" mode)
template <class E>
__device__ __forceinline__ void fallback_enum(const typename E::Args& a, int64_t i, const typename E::Params& p,
                                              const typename E::EnvParams& e, rcbf_counters_t* ws,
                                              const unsigned short* table, int ntable, int lane) {
  constexpr int NZ = E::NZ, M = E::M, NU = E::NU;
  typename E::Inst in;
  E::load_inst(a, i, in);
  Normalised<NZ, M> nrm;
  E::assemble(a, p, i, in, nrm);  // every lane assembles the same instance: 32x redundant, a few hundred flops
  double pisd[NZ];
  float pisf[NZ];
  pis_of<NZ, M>(p.p_diag, pisd, pisf);
  const NormCert<NZ, M> cp{nrm, pisd};
  NormSolution<NZ, M> sol;
  bool found = false;
  for (int base = 0; base < ntable && !__any_sync(0xffffffffu, found); base += 32) {
    const int t = base + lane;
    double y[NZ], lam[M], s[M];
    const uint32_t mask = (t < ntable) ? table[t] : 0u;
    const bool ok = (t < ntable) && lnp_certify<double, NormCert<NZ, M>, typename E::Pat, NZ, M>(cp, mask, kTolSlack,
                                                                                                  kTolDual, y, lam, s);
    if (ok && !found) {
      found = true;
#pragma unroll
      for (int j = 0; j < NZ; ++j) sol.x[j] = y[j] * pisd[j];
#pragma unroll
      for (int r = 0; r < M; ++r) {
        sol.lam[r] = lam[r];
        sol.s[r] = s[r];
      }
      sol.status = RCBF_OK_CERTIFIED;
      sol.iters = NZ + 1;  // marks "enumerated" in the iteration histogram
      sol.mask = mask;
    }
  }
  const unsigned winners = __ballot_sync(0xffffffffu, found);
  if (winners == 0u) {
    if (lane == 0) fallback_ipm<E, false>(a, i, p, e, ws);  // degenerate to working precision: interior point
    return;
  }
  if (lane == __ffs(winners) - 1) {
    float xs[NU];
#pragma unroll
    for (int c = 0; c < NU; ++c) xs[c] = (float)sol.x[c];
    write_saved<E>(a, i, sol);
    E::finish(a, p, e, i, in, xs, sol.status);
    fallback_counters(ws, sol.status, 0);
  }
}


// table of every active set with 1..NZ rows (129 for the Unicycle, 10 for SimulatedCars), lane-parallel build
template <int NZ, int M>
__device__ __forceinline__ int build_enum_table(unsigned short* table, int lane) {
  int ntable = 0;
  for (int m0 = 0; m0 < (1 << M); m0 += 32) {
    const int m = m0 + lane;
    const int pc = __popc(m);
    const bool keep = (m < (1 << M)) && pc >= 1 && pc <= NZ;
    const unsigned b = __ballot_sync(0xffffffffu, keep);
    if (keep) table[ntable + __popc(b & ((1u << lane) - 1u))] = (unsigned short)m;
    ntable += __popc(b);
  }
  __syncwarp();
  return ntable;
}

// Tail of the presolve-mode kernel: a warp that has run out of tiles drains the queue of pending instances (one warp
// per instance, exhaustive enumeration).  Entries are claimed with a CAS on ws[kWsClaim]; an entry is published by its
// producer right after the counter increment, so the claimer spins the few cycles until the slot turns non-zero and
// clears it again (the queue is all-zero between calls).  Whatever a warp enqueued during its own tiles exists before
// that warp drains, so nothing is left when the last block finishes -- no second kernel launch is needed.
template <class E>
__device__ __noinline__ void tail_drain(const typename E::Args a, const typename E::Params p,
                                        const typename E::EnvParams e, rcbf_counters_t* ws, unsigned short* table,
                                        int lane) {  // by VALUE: taking the address of a kernel parameter would move it
                                                     // (and every access in the hot loop) from the constant bank to the stack
  volatile rcbf_counters_t* vws = ws;
  int ntable = -1;
  for (;;) {
    unsigned long long v = 0ULL;
    if (lane == 0) {
      for (;;) {
        const unsigned long long c = vws[kWsClaim];
        unsigned long long cnt = vws[kWsQueueCount];
        cnt = cnt < (unsigned long long)kWsQueueCap ? cnt : (unsigned long long)kWsQueueCap;
        if (c >= cnt) break;
        if (atomicCAS(&ws[kWsClaim], c, c + 1ULL) == c) {
          do {
            v = vws[kWsQueueBase + c];
          } while (v == 0ULL);
          vws[kWsQueueBase + c] = 0ULL;
          break;
        }
      }
    }
    v = __shfl_sync(0xffffffffu, v, 0);
    if (v == 0ULL) break;
    if (ntable < 0) ntable = build_enum_table<E::NZ, E::M>(table, lane);
    fallback_enum<E>(a, (int64_t)(v - 1ULL), p, e, ws, table, ntable, lane);
  }
}

// Queue overflow (more than kWsQueueCap pending instances in one call: pathological inputs): the LAST block scans the
// output for the pending sentinel.  Slow, but only there to stay correct.
template <class E>
__device__ __noinline__ void tail_scan(const typename E::Args a, int64_t n, const typename E::Params p,
                                       const typename E::EnvParams e, rcbf_counters_t* ws, unsigned short* table,
                                       int lane, int warp, int nwarps = kWarps) {
  const int ntable = build_enum_table<E::NZ, E::M>(table, lane);
  for (int64_t i0 = (int64_t)warp * 32; i0 < n; i0 += nwarps * 32) {
    const int64_t i = i0 + lane;
    const bool pend = (i < n) && __float_as_uint(__ldcg(a.out + i * E::NU)) == kPendingBits;
    unsigned b = __ballot_sync(0xffffffffu, pend);
    while (b) {
      const int src = __ffs(b) - 1;
      b &= b - 1;
      fallback_enum<E>(a, i0 + src, p, e, ws, table, ntable, lane);
    }
  }
}

template <class E, int kMode /* 0 presolve, 1 pdipm */, bool kBulk /* TMA bulk-copy input staging */,
          bool kSaved /* also write x / lam / slack / iters (backward pass, diagnostics) */>
__global__ void __launch_bounds__(kThreadsW, kMode == 0 ? E::kMinBlocks : RCBF_MINB_PDIPM)
k_safe(typename E::Args a, int64_t n, typename E::Params p, typename E::EnvParams e, rcbf_counters_t* ws) {
  constexpr int NZ = E::NZ, M = E::M, NU = E::NU, NWR = E::NWR;
  using Inst = typename E::Inst;
  using WS = WarpShared<E, kMode>;
  constexpr int kRing = WS::kRing, kFin = WS::kFin;
  constexpr bool kMerge = WS::kMerge;
  constexpr bool kInline = kMode == 0 && E::kInlineSolve;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  WS& sh = reinterpret_cast<WS*>(smem_raw)[warp];
  const unsigned lt_mask = (1u << lane) - 1u;
  // 32-bit tile arithmetic (launch_safe refuses n > 2^31 - 1): the full-tile test is `tile < nfull`
  const int ntiles = (int)((n + 31) >> 5);
  const int nfull = (int)(n >> 5);
  const int nw = (int)gridDim.x * kWarps;
  int tile = (int)blockIdx.x * kWarps + warp;
  int head = 0, qn = 0;     // problem ring
  int fhead = 0, fn = 0;    // finish ring
  int c_nan = 0, c_triv = 0, c_pend = 0, c_iters = 0;
  int nbulk = 0;  // bulk-staged tiles consumed so far by this warp (buffer = nbulk & 1, mbarrier parity = (nbulk >> 1) & 1)
  // presolve mode: job B = the instance this lane solved in the previous B-step, finished after the next A-step so
  // that only (index, correction, status) is live across the register-hungry solve.  (pdipm mode: finish ring.)
  bool onB = false;
  int iB = 0, stB = RCBF_OK_CERTIFIED;
  float xsB[NU];
#pragma unroll
  for (int c = 0; c < NU; ++c) xsB[c] = 0.f;
  pdl_wait();
  if (kBulk) {
    if (lane == 0) {
      mbar_init(&sh.bar[0], 1);
      mbar_init(&sh.bar[1], 1);
      asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncwarp();
    if (lane == 0 && tile < nfull) E::issue(a, tile, sh.stage[0], &sh.bar[0]);
  }

  for (;;) {
    const bool have_tile = tile < ntiles;
    // job A: this lane's own instance when it needs no solve
    bool onA = false;
    int64_t iA = 0;
    Inst inA;
    float xsA[NU];
    int stA = RCBF_OK_TRIVIAL;
    int stageA = -1;  // landing buffer that still holds this iteration's tile (bulk-staged full tiles only)

    if (have_tile) {  // ---- A-step: assemble, classify, queue
      const int64_t i0 = ((int64_t)tile << 5) + lane;
      const bool valid = i0 < n;
      iA = valid ? i0 : n - 1;
      typename E::Aux aux;
      if (kBulk && tile < nfull) {
        // this tile's inputs were bulk-copied into shared memory one iteration ago; wait, read, then put the NEXT
        // tile in flight into the other buffer (its last reader finished before the previous __syncwarp)
        const int b = nbulk & 1;
        stageA = b;
        mbar_wait(&sh.bar[b], (nbulk >> 1) & 1);
        E::read_stage(sh.stage[b], lane, inA, aux);
        __syncwarp();
        if (lane == 0 && tile + nw < nfull) E::issue(a, tile + nw, sh.stage[b ^ 1], &sh.bar[b ^ 1]);
        ++nbulk;
      } else {
        if (!kBulk && tile + nw < ntiles) {  // pull the NEXT tile's input lines towards L1 meanwhile
          const int64_t j0 = ((int64_t)(tile + nw) << 5) + lane;
          E::prefetch(a, j0 < n ? j0 : n - 1);
        }
        E::load_inst(a, iA, inA);
        E::load_aux(a, iA, aux);
      }
      float w[NWR];
      bool triv, nan;
      E::assemble_raw(p, inA, aux, w, triv, nan);
      bool need = valid && !triv && !nan;
      bool solvedA = false;
      NormSolution<NZ, M> solA;
      if constexpr (kInline) {
        // inline solve (SimulatedCars fused step): the 2 x 4 QP is cheap and the kernel is bound by memory requests,
        // so the lanes that need a solve do it right here at partial lane utilisation; every instance of the tile is
        // then finished by the coalesced whole-tile path (no second pass with lane-per-row 40-byte accesses)
        if (need) {
          float Gr[M][NZ], hr[M];
          E::unpack_raw(w, p, Gr, hr);
          solve_raw_fast<typename E::Pat, NZ, M>(Gr, hr, p.p_diag, kSaved, solA);
          if (solA.status == RCBF_PENDING) {
            mark_pending<E>(a, iA, ws);
            c_pend += 1;
          } else {
            solvedA = true;
            if (kSaved) write_saved<E>(a, iA, solA);
            c_iters += solA.iters;
          }
        }
        need = false;
      }
      const unsigned ballot = __ballot_sync(0xffffffffu, need);
      if (need) {
        const int slot = (head + qn + __popc(ballot & lt_mask)) & (kRing - 1);
        sh.push(slot, w, (int)iA);
        if constexpr (kMerge) {
          float sc[E::NSC];
          E::stash(inA, sc);
#pragma unroll
          for (int k = 0; k < E::NSC; ++k) sh.sc[k][slot] = sc[k];
        }
      }
      onA = kInline ? (valid && (triv || nan || solvedA)) : (valid && !need);
      stA = nan ? RCBF_NAN : RCBF_OK_TRIVIAL;
#pragma unroll
      for (int c = 0; c < NU; ++c) xsA[c] = nan ? NAN : 0.f;
      if constexpr (kInline) {
        if (solvedA) {
          stA = solA.status;
#pragma unroll
          for (int c = 0; c < NU; ++c) xsA[c] = (float)solA.x[c];
        }
      }
      if (kSaved && onA && !solvedA) {  // trivial / NaN instance: x = 0 (NaN), lam = 0, slack = h~
        Normalised<NZ, M> nrm;
        E::normalise_packed(w, p, nrm);
        NormSolution<NZ, M> sol;
        trivial_solution<typename E::Pat, NZ, M>(nrm, nan, sol);
        write_saved<E>(a, iA, sol);
      }
      c_nan += (onA && nan) ? 1 : 0;
      c_triv += (onA && !nan && !solvedA) ? 1 : 0;
      qn += __popc(ballot);
      tile += nw;
    }
    if constexpr (kMerge) cp_async_wait_all();  // the finish inputs the previous B-step asked for have landed
    __syncwarp();

    // ---- finish (clamp, env.step, outputs): ONE copy of the code.  Pass 0 = job A; later passes = job B of the
    // previous iteration (presolve mode) or full warps popped from the finish ring (pdipm mode)
#pragma unroll 1
    for (int j = 0;; ++j) {
      bool on;
      int64_t i;
      Inst in;
      float xs[NU];
      int stv;
      if constexpr (kMerge) {
        // merged finish: pass 0 = job A, its idle lanes (own instance queued, or past the end) each take one solved
        // instance from the finish ring.  A further pass with every lane idle runs while the ring could not take the
        // next B-step's 32 results (such a pass is a full warp of useful work, so it costs no efficiency).
        if (j > 0) {
          if (fn <= kFin - 32) break;
          onA = false;
        }
        on = onA;
        i = iA;
        in = inA;
#pragma unroll
        for (int c = 0; c < NU; ++c) xs[c] = xsA[c];
        stv = stA;
        if (fn > 0) {
          const unsigned idle = __ballot_sync(0xffffffffu, !onA);
          const int r = __popc(idle & lt_mask);
          if (!onA && r < fn) {
            const int fs = (fhead + r) & (kFin - 1);
            E::fin_read(sh.fin, fs, in, xs, i, stv);
            on = true;
          }
          const int took = min(fn, __popc(idle));
          fhead = (fhead + took) & (kFin - 1);
          fn -= took;
        }
      } else if (j == 0) {
        on = onA;
        i = iA;
        in = inA;
#pragma unroll
        for (int c = 0; c < NU; ++c) xs[c] = xsA[c];
        stv = stA;
      } else if (kMode == 0) {
        if (j > 1) break;
        on = onB;
        i = iB;
#pragma unroll
        for (int c = 0; c < NU; ++c) xs[c] = xsB[c];
        stv = stB;
        if (on) E::load_inst(a, i, in);
        onB = false;
      } else {
        if (fn == 0 || (fn < 32 && (have_tile || qn > 0))) break;
        const int take = fn < 32 ? fn : 32;
        on = lane < take;
        const int fs = (fhead + lane) & (kFin - 1);
        i = on ? sh.fidx[fs] : 0;
#pragma unroll
        for (int c = 0; c < NU; ++c) xs[c] = sh.fx[c][fs];
        stv = sh.fst[fs];
        if (on) E::load_inst(a, i, in);
        fhead = (fhead + take) & (kFin - 1);
        fn -= take;
      }
      if constexpr (E::kTileFinish) {
        if (j == 0 && stageA >= 0) {
          E::finish_tile(a, p, e, iA - lane, lane, on, in, xs, stv, sh.stage[stageA]);
          continue;
        }
      }
      if (on) E::template finish<kBulk>(a, p, e, i, in, xs, stv);
      __syncwarp();
    }

    if (kMode == 0) {
      // ---- B-step: a full warp of queued problems (or whatever is left once the tiles are exhausted)
      const int take = (qn >= 32) ? 32 : (have_tile ? 0 : qn);
      if (take > 0) {
#ifndef RCBF_EXP_NOSOLVE  // (experiment switch: phase-1-only cost model, results are WRONG when defined)
        const int slot = (head + lane) & (kRing - 1);
        if (lane < take) {
          float w[NWR];
          sh.pop(slot, w, iB);
          float Gr[M][NZ], hr[M];
          E::unpack_raw(w, p, Gr, hr);
          NormSolution<NZ, M> sol;
          solve_raw_fast<typename E::Pat, NZ, M>(Gr, hr, p.p_diag, kSaved, sol);
          if (sol.status == RCBF_PENDING) {
            mark_pending<E>(a, iB, ws);
            c_pend += 1;
          } else {
            onB = true;
            stB = sol.status;
#pragma unroll
            for (int c = 0; c < NU; ++c) xsB[c] = (float)sol.x[c];
            if (kSaved) write_saved<E>(a, iB, sol);
            c_iters += sol.iters;
          }
        }
        if constexpr (kMerge) {  // hand the solved instances to the finish ring (consumed by idle lanes of later finish passes)
          const unsigned fb = __ballot_sync(0xffffffffu, onB);
          if (onB) {
            const int fs = (fhead + fn + __popc(fb & lt_mask)) & (kFin - 1);
            float sc[E::NSC];
#pragma unroll
            for (int k = 0; k < E::NSC; ++k) sc[k] = sh.sc[k][slot];
            E::template fin_fetch<kFin, kBulk>(a, iB, sh.fin, fs, sc, xsB, stB);
          }
          fn += __popc(fb);
          onB = false;
        }
#endif
        head = (head + take) & (kRing - 1);
        qn -= take;
        __syncwarp();
      }
    } else {
      // ---- interior-point engine: every lane holds one QP; a lane whose QP is done (certified, or given up ->
      // pending) immediately takes the next problem from the ring, so the warp stays full although the iteration
      // counts differ widely (1..16, mean ~2).  Runs when >= 64 problems wait (or at the end) and drains the ring.
      if (qn >= 64 || (!have_tile && qn > 0)) {
        int remaining = qn;
        bool active = false;
        int my_idx = 0;
        Normalised<NZ, M> nrm;
        LnpProblem<float, NZ, M> Pf;
        IpmState<float, NZ, M> st;
        double pisd[NZ];
        float pisf[NZ];
        pis_of<NZ, M>(p.p_diag, pisd, pisf);
        for (;;) {
          const unsigned wb = __ballot_sync(0xffffffffu, !active);
          const int rank = __popc(wb & lt_mask);
          if (!active && rank < remaining) {
            const int slot = (head + (qn - remaining) + rank) & (kRing - 1);
            float w[NWR];
            sh.pop(slot, w, my_idx);
            E::normalise_packed(w, p, nrm);
            to_lnp<float, typename E::Pat, NZ, M>(nrm, pisf, Pf);
            ipm_init<float, typename E::Pat, NZ, M>(Pf, st);
            active = true;
          }
          remaining -= min(remaining, __popc(wb));
          if (!__any_sync(0xffffffffu, active)) break;
          int status = IPM_CONTINUE;
          LnpSolution<double, NZ, M> lsol;
          if (active) {
            const NormCert<NZ, M> cp{nrm, pisd};
            status = ipm_step<float, double, NormCert<NZ, M>, typename E::Pat, NZ, M, false>(Pf, cp, st, lsol, kTolSlack,
                                                                                             kTolDual);
          }
          const bool fin = active && status != IPM_CONTINUE;
          const bool ok = fin && status < RCBF_MAXITER;
          if (fin && !ok) {
            mark_pending<E>(a, my_idx, ws);
            c_pend += 1;
          }
          const unsigned fb = __ballot_sync(0xffffffffu, ok);
          if (ok) {
            const int fs = (fhead + fn + __popc(fb & lt_mask)) & (kFin - 1);
            NormSolution<NZ, M> sol;
#pragma unroll
            for (int j = 0; j < NZ; ++j) sol.x[j] = lsol.y[j] * pisd[j];
#pragma unroll
            for (int c = 0; c < NU; ++c) sh.fx[c][fs] = (float)sol.x[c];
            sh.fidx[fs] = my_idx;
            sh.fst[fs] = status;
            if (kSaved) {
#pragma unroll
              for (int r = 0; r < M; ++r) {
                sol.lam[r] = lsol.lam[r];
                sol.s[r] = lsol.s[r];
              }
              sol.iters = lsol.iters;
              sol.status = status;
              sol.mask = lsol.mask;
              write_saved<E>(a, my_idx, sol);
            }
            c_iters += lsol.iters;
          }
          fn += __popc(fb);
          if (fin) active = false;
        }
        head = (head + qn) & (kRing - 1);
        qn = 0;
        __syncwarp();
      }
    }

    if (!have_tile && qn == 0 && fn == 0 && !__any_sync(0xffffffffu, onB)) break;
  }
  pdl_launch_dependents();  // the next kernel may be scheduled as soon as every block of this grid got here
  const bool own_tail = (kMode == 0) && (ws != nullptr);  // presolve mode with a workspace: no pass-2 kernel
  if (own_tail) tail_drain<E>(a, p, e, ws, sh.scratch(), lane);

  if (ws != nullptr) {
    c_nan = __reduce_add_sync(0xffffffffu, c_nan);
    c_triv = __reduce_add_sync(0xffffffffu, c_triv);
    c_pend = __reduce_add_sync(0xffffffffu, c_pend);
    c_iters = __reduce_add_sync(0xffffffffu, c_iters);
    if (lane == 0) {
      if (c_nan) atomicAdd(&ws[0], (unsigned long long)c_nan);
      if (c_triv) atomicAdd(&ws[3], (unsigned long long)c_triv);
      if (c_iters) atomicAdd(&ws[4], (unsigned long long)c_iters);
      if (c_pend) atomicAdd(&ws[5], (unsigned long long)c_pend);
    }
  }
  if (own_tail) {  // the last block to get here handles a queue overflow and resets the queue for the next call
    __shared__ int s_last;
    __syncthreads();
    if (threadIdx.x == 0) {
      __threadfence();
      s_last = (atomicAdd(&ws[kWsBlocksDone], 1ULL) == (unsigned long long)gridDim.x - 1ULL) ? 1 : 0;
    }
    __syncthreads();
    if (s_last) {
      __threadfence();
      const unsigned long long cnt = *(volatile rcbf_counters_t*)&ws[kWsQueueCount];
      if (cnt > (unsigned long long)kWsQueueCap)
        tail_scan<E>(a, n, p, e, ws, sh.scratch(), lane, warp);
      __syncthreads();
      if (threadIdx.x == 0) {
        ws[kWsQueueCount] = 0ULL;
        ws[kWsClaim] = 0ULL;
        ws[kWsBlocksDone] = 0ULL;
        __threadfence();
        publish_counters(ws, p.solver_mode);
      }
    }
  }
}

// ---------------------------------------------------------------------------------------------------------------
// pass 2 kernel: sentinel scan when the caller gave no workspace, and the queue of the interior-point ("pdipm") mode
// ---------------------------------------------------------------------------------------------------------------
template <class E, int kMode>
__global__ void __launch_bounds__(128)
k_safe_fallback(typename E::Args a, int64_t n, typename E::Params p, typename E::EnvParams e, rcbf_counters_t* ws) {
  constexpr int NZ = E::NZ, M = E::M;
  const int lane = threadIdx.x & 31;
  const int64_t gtid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x, gsz = (int64_t)gridDim.x * blockDim.x;
  const int64_t gwarp = gtid >> 5, nwarp = gsz >> 5;
  pdl_wait();
  pdl_launch_dependents();
  unsigned long long cnt = (ws != nullptr) ? ws[kWsQueueCount] : ~0ULL;
  const bool scan = cnt > (unsigned long long)kWsQueueCap;  // no workspace, or overflow: sentinel scan
  if (!scan && cnt == 0ULL) goto done;

  if (kMode == 0) {
    __shared__ unsigned short s_table[4][160];
    unsigned short* table = s_table[threadIdx.x >> 5];
    const int ntable = build_enum_table<NZ, M>(table, lane);
    if (!scan) {
      for (int64_t q = gwarp; q < (int64_t)cnt; q += nwarp) {
        const unsigned long long v = ws[kWsQueueBase + q];
        __syncwarp();
        if (lane == 0) ws[kWsQueueBase + q] = 0ULL;
        fallback_enum<E>(a, (int64_t)(v - 1ULL), p, e, ws, table, ntable, lane);
      }
    } else {
      for (int64_t i0 = gwarp * 32; i0 < n; i0 += nwarp * 32) {
        const int64_t i = i0 + lane;
        const bool pend = (i < n) && __float_as_uint(__ldcg(a.out + i * E::NU)) == kPendingBits;
        unsigned b = __ballot_sync(0xffffffffu, pend);
        while (b) {
          const int src = __ffs(b) - 1;
          b &= b - 1;
          fallback_enum<E>(a, i0 + src, p, e, ws, table, ntable, lane);
        }
      }
    }
  } else {
    if (!scan) {
      for (int64_t q = gtid; q < (int64_t)cnt; q += gsz) {
        const unsigned long long v = ws[kWsQueueBase + q];
        ws[kWsQueueBase + q] = 0ULL;
        fallback_ipm<E, true>(a, (int64_t)(v - 1ULL), p, e, ws);
      }
    } else {
      for (int64_t i = gtid; i < n; i += gsz)
        if (__float_as_uint(__ldcg(a.out + i * E::NU)) == kPendingBits) fallback_ipm<E, true>(a, i, p, e, ws);
    }
  }
  if (scan && ws != nullptr)  // overflow: the stored part of the queue was not consumed entry by entry; clear it
    for (int64_t q = gtid; q < kWsQueueCap; q += gsz) ws[kWsQueueBase + q] = 0ULL;
done:
  if (ws != nullptr) {  // the last block to finish resets the queue for the next call
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

// kernel launch, optionally with the PDL attribute (RCBF_NO_PDL=1 in the environment forces plain launches)
template <typename... KArgs, typename... Args>
inline cudaError_t launch_pdl(bool pdl, void (*kern)(KArgs...), int grid, int block, size_t smem, cudaStream_t s,
                              Args... args) {
  static const bool env_off = [] {
    const char* v = getenv("RCBF_NO_PDL");
    return v != nullptr && v[0] == '1';
  }();
  const bool no_pdl = env_off || !pdl;
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3((unsigned)grid);
  cfg.blockDim = dim3((unsigned)block);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = s;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = no_pdl ? 0 : 1;
  return cudaLaunchKernelEx(&cfg, kern, KArgs(args)...);
}

// SMs of the current device (cached per device; the persistent grids are sized from it)
inline int device_sm_count() {
  static int cached[64] = {};
  int dev = 0;
  cudaGetDevice(&dev);
  int& c = cached[dev & 63];
  if (c == 0) cudaDeviceGetAttribute(&c, cudaDevAttrMultiProcessorCount, dev);
  return c > 0 ? c : 148;
}

template <class E>
inline int launch_safe(const typename E::Args& a, int64_t n, const typename E::Params& p, const typename E::EnvParams& e,
                       rcbf_counters_t* ws, cudaStream_t s) {
  if (n <= 0) return 0;
  if (n > 0x7fffffffLL) return -2;  // ring indices are 32-bit
  const int64_t ntiles = (n + 31) / 32;
  const int64_t want = (ntiles + kWarps - 1) / kWarps;
  const int sms = device_sm_count();
  const int resident = sms * (solver_mode_of(p) == 0 ? E::kMinBlocks : RCBF_MINB_PDIPM);  // persistent: one wave of resident blocks
  const int grid = (int)(want < resident ? want : resident);
  const int64_t fb = (n + 127) / 128;
  const int fgrid = (int)(fb < sms * 4 ? fb : sms * 4);
  // TMA bulk staging needs 16-byte aligned array bases (row spans of a 32-instance tile are then 16-byte multiples)
  const bool bulk = E::aligned(a) && n >= 32;
  const bool saved = (a.x != nullptr || a.lam != nullptr || a.slack != nullptr || a.iters != nullptr || a.meta != nullptr);
#define RCBF_LAUNCH_ONE(MODE, BULK, SAVED)                                                           \
  do {                                                                                               \
    constexpr size_t smem = sizeof(WarpShared<E, MODE>) * kWarps;                                    \
    static bool configured[64] = {}; /* > 48 KB of dynamic shared memory needs the opt-in, once per device */ \
    int dev_ = 0;                                                                                    \
    cudaGetDevice(&dev_);                                                                            \
    if (!configured[dev_ & 63]) {                                                                    \
      cudaFuncSetAttribute(k_safe<E, MODE, BULK, SAVED>, cudaFuncAttributeMaxDynamicSharedMemorySize, \
                           (int)smem);                                                               \
      configured[dev_ & 63] = true;                                                                  \
    }                                                                                                \
    launch_pdl(E::kPdlPass1 && MODE == 0 && !SAVED, k_safe<E, MODE, BULK, SAVED>, grid, kThreadsW, smem, s, a, n, \
               p, e, ws);                                                                            \
  } while (0)
#define RCBF_LAUNCH_SAFE(MODE)                                                                       \
  do {                                                                                               \
    if (bulk && saved) RCBF_LAUNCH_ONE(MODE, true, true);                                            \
    else if (bulk) RCBF_LAUNCH_ONE(MODE, true, false);                                               \
    else if (saved) RCBF_LAUNCH_ONE(MODE, false, true);                                              \
    else RCBF_LAUNCH_ONE(MODE, false, false);                                                        \
    if (ws == nullptr || MODE == 1) launch_pdl(true, k_safe_fallback<E, MODE>, fgrid, 128, 0, s, a, n, p, e, ws); \
  } while (0)
  if (solver_mode_of(p) == 0) RCBF_LAUNCH_SAFE(0);
  else RCBF_LAUNCH_SAFE(1);
#undef RCBF_LAUNCH_SAFE
#undef RCBF_LAUNCH_ONE
  return (int)cudaGetLastError();
}

}  // namespace rcbf
