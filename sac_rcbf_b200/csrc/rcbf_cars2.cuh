// rcbf_cars2.cuh -- k_cars2: the fused SimulatedCars safe step (get_safe_action, diff_cbf_qp.py:44-79 + :268-357, and
// SimulatedCarsEnv.step, simulated_cars_env.py:38-106) as a persistent kernel with a warp-private PROBLEM RING and a
// two-iteration finish lag, TMA in and out.
//
// k_safe<CarsEnv<true>> (rcbf_safe_kernels.cuh) solves a tile's QPs inline: the ~9 of 32 lanes whose instance needs a
// solve run the active-set presolve + float64 certificate while the other lanes idle (29 % lane utilisation on ~300 of
// the kernel's 964 warp instructions per tile).  Here, per persistent-loop iteration k of a warp (tile = 32 instances,
// one per lane):
//
//   A-step(tile k)   the tile's state / action / t / step rows arrived by cp.async.bulk one iteration ahead (sigma in
//                    a single buffer that is refilled as soon as it has been read); assembly, trivial test; a trivial
//                    / NaN instance gets its clamped action written over the nominal one in the landing slot, an
//                    instance that needs a solve pushes 6 words (the two Lgf, the four right-hand sides) + a tag into
//                    the warp's ring.
//   B-step           whenever >= 32 problems wait, or -- if the tile finished below still has problems in the ring --
//                    everything that waits: one problem per lane (~27 of 32 lanes busy instead of 9), the clamped safe
//                    action goes back into the owning tile's slot.
//   finish(tile k-2) every lane finishes its own instance of that tile from the slot: env.step, reward, cost, done;
//                    the new state rows replace the old ones in the slot, the observation rows are staged next to it,
//                    and both leave as ONE bulk store each (shared -> global, 1280 contiguous bytes) instead of twenty
//                    40-byte-stride vector stores per lane; the lane-contiguous arrays are stored directly.
//
// Instances the B-step cannot certify keep their OLD state / t / step, carry the pending sentinel in safe_action[i] and
// are queued AFTER their tile's bulk stores completed; the kernel's tail (tail_drain) redoes them -- one launch per step.
// Results are bit-identical to k_safe<CarsEnv<true>> (same per-instance functions, same operation order).
//
// k_cars2<false> is the same loop for get_safe_action alone (read-only state; the finish stores the safe action and the
// optional status / meta words): 0.100 -> see DESIGN 4.2b.
//
// Launch conditions (launch_cars2): solver_mode 0, no dense saved tensors, every array base 16-byte aligned; n is split
// into full 32-instance tiles for this kernel and a ragged tail (< 32) for k_safe.
#pragma once

#include "rcbf_safe_kernels.cuh"

namespace rcbf {

// A/B on B200 (4 Mi instances, ms per step): 2 blocks x 8 warps 0.1385 (steadiest), 1 x 16 0.138-0.143, 1 x 12 0.151,
// 1 x 20 (96 registers, spills) 0.160
#ifndef RCBF_C2_WARPS
#define RCBF_C2_WARPS 8
#endif
#ifndef RCBF_C2_MINB
#define RCBF_C2_MINB 2
#endif
#ifndef RCBF_C2_MIN_N
#define RCBF_C2_MIN_N 4096   // below this the few tiles spread over more warps with k_safe's 4-warp blocks
#endif
constexpr int kC2Warps = RCBF_C2_WARPS;
constexpr int kC2Threads = 32 * kC2Warps;
constexpr int kC2Ring = 64;  // the B-steps leave < 32 problems, a tile adds <= 32

struct alignas(16) C2Warp {
  struct alignas(16) In {   // TMA landing slot of one tile; `ac` turns into the clamped SAFE action (A- / B-steps), `st`
    float st[320];          // into the NEW state rows (finish), which leave from here
    float ac[32];
    float t[32];
    int step[32];
  };
  In in[4];                 // tiles k-2 (finishing), k-1, k, k+1 (in flight)
  float sg[320];            // sigma rows: only the assembly reads them (single buffer)
  float obs[320];           // observation rows of the tile being finished (bulk-stored from here)
  float4 ring[kC2Ring][2];  // problem ring: Lgf[2], h[4], tag, -
  uint8_t cls[3][32];       // per instance: RCBF_OK_TRIVIAL / RCBF_OK_CERTIFIED / RCBF_NAN / RCBF_PENDING
  uint16_t amask[3][32];    // layer-only kernel: active set of the certified solution (saved for the backward)
  uint64_t bar[2];
};
static_assert(sizeof(C2Warp) * kC2Warps * RCBF_C2_MINB + 1024 * RCBF_C2_MINB <= 233472, "shared memory of an SM");

// kFused: the safe step (state in place, env outputs); !kFused: get_safe_action alone (state read-only, safe action +
// optional status / meta out) -- same loop, the finish is then three lane-contiguous stores.
template <bool kFused>
__global__ void __launch_bounds__(kC2Threads, RCBF_C2_MINB)
k_cars2(CarsArgs a, int64_t n /* multiple of 32 */, CarsParams p, CarsEnvParams e, rcbf_counters_t* ws) {
  using E = CarsEnv<kFused>;
  constexpr int NZ = kCarsNZ, M = kCarsM, NWR = E::NWR;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  C2Warp& sh = reinterpret_cast<C2Warp*>(smem_raw)[warp];
  const unsigned lt_mask = (1u << lane) - 1u;
  const int ntiles = (int)(n >> 5);
  const int nw = (int)gridDim.x * kC2Warps;
  int tile = (int)blockIdx.x * kC2Warps + warp;  // tile of iteration k
  int tile1 = tile + nw;                         // k + 1
  int tm1 = 0, tm2 = 0;                          // k - 1, k - 2
  int head = 0, qn = 0;                          // problem ring
  int pk = 0, pk1 = 0;                           // problems pushed by tiles k, k-1 (the newest entries of the ring)
  int r3 = 0;                                    // k % 3
  int c_nan = 0, c_pend = 0, c_iters = 0, n_solve = 0, n_tiles = 0;
  constexpr uint32_t kInBytes = 1280 + 128 + (kFused ? 128 + 128 : 0) + 1280;

  auto issue = [&](int t, int k) {  // lane 0: tile t in flight into slot k & 3 (+ the sigma buffer)
    C2Warp::In& si = sh.in[k & 3];
    uint64_t* bar = &sh.bar[k & 1];
    const int64_t i0 = (int64_t)t << 5;
    mbar_expect_tx(bar, kInBytes);
    bulk_g2s(si.st, (kFused ? a.state : a.st) + i0 * 10, 1280, bar);
    bulk_g2s(si.ac, a.ac + i0, 128, bar);
    if (kFused) {
      bulk_g2s(si.t, a.t + i0, 128, bar);
      bulk_g2s(si.step, a.step + i0, 128, bar);
    }
    bulk_g2s(sh.sg, a.sg + i0 * 10, 1280, bar);
  };

  pdl_wait();
  if (lane == 0) {
    mbar_init(&sh.bar[0], 1);
    mbar_init(&sh.bar[1], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncwarp();
  if (lane == 0 && tile < ntiles) issue(tile, 0);

  int after = 0;  // iterations past this warp's last tile (the last two tiles are finished then)
#pragma unroll 1
  for (int k = 0;; ++k) {
    const bool have_tile = tile < ntiles;
    pk1 = pk;
    pk = 0;
    if (have_tile) {  // ---------------------------------------------------------------- A-step
      ++n_tiles;
      C2Warp::In& si = sh.in[k & 3];
      mbar_wait(&sh.bar[k & 1], (k >> 1) & 1);
      typename E::Inst in;
      typename E::Aux aux;
      {
        const float2* sp = reinterpret_cast<const float2*>(si.st) + lane * 5;
        const float2* gp = reinterpret_cast<const float2*>(sh.sg) + lane * 5;
#pragma unroll
        for (int q = 0; q < 5; ++q) {
          const float2 s2 = sp[q], g2 = gp[q];
          aux.s[2 * q] = s2.x; aux.s[2 * q + 1] = s2.y;
          aux.g[2 * q] = g2.x; aux.g[2 * q + 1] = g2.y;
        }
        in.u[0] = si.ac[lane];
        in.tt = 0.f;
        in.stp = 0;
      }
      __syncwarp();
      // slot (k+1) & 3 belonged to the tile finished in the previous iteration: its bulk stores (state rows from the
      // slot, observation rows from sh.obs) were issued then and have long read their sources
      if (lane == 0 && tile1 < ntiles) {
        bulk_wait_read0();
        issue(tile1, k + 1);
      }
      float w[NWR];
      bool triv, nan;
      E::assemble_raw(p, in, aux, w, triv, nan);
      const bool need = !triv && !nan;
      if (!need) si.ac[lane] = clampf(in.u[0] + (nan ? NAN : 0.f), p.u_min, p.u_max);   // diff_cbf_qp.py:77 (x = 0; NaN propagates)
      sh.cls[r3][lane] = (uint8_t)(nan ? RCBF_NAN : (need ? RCBF_OK_CERTIFIED : RCBF_OK_TRIVIAL));
      if (!kFused) sh.amask[r3][lane] = 0;
      const unsigned b = __ballot_sync(0xffffffffu, need);
      pk = __popc(b);
      if (need) {
        int slot = head + qn + __popc(b & lt_mask);
        slot -= slot >= kC2Ring ? kC2Ring : 0;
        const int tag = ((k & 3) << 8) | (r3 << 6) | lane;
        sh.ring[slot][0] = make_float4(w[0], w[1], w[2], w[3]);
        sh.ring[slot][1] = make_float4(w[4], w[5], __int_as_float(tag), 0.f);
      }
      qn += pk;
      n_solve += pk;
      if (__any_sync(0xffffffffu, nan)) c_nan += nan ? 1 : 0;
    } else {
      ++after;
    }
    __syncwarp();

    // ---------------------------------------------------------------- B-steps
    const bool fin_due = (k >= 2) && (after <= 2);
#pragma unroll 1
    for (;;) {
      // problems of tile k-2 (finished below) still waiting: the oldest entries.  (pk / pk1 count what tiles k / k-1
      // PUSHED; if some of those are solved already the expression is <= 0, and so is the true number)
      const int old = qn - pk - pk1;
      const int take = qn >= 32 ? 32 : ((fin_due && old > 0) ? qn : 0);
      if (take == 0) break;
      if (lane < take) {
        int slot = head + lane;
        slot -= slot >= kC2Ring ? kC2Ring : 0;
        const float4 q0 = sh.ring[slot][0], q1 = sh.ring[slot][1];
        const float w[NWR] = {q0.x, q0.y, q0.z, q0.w, q1.x, q1.y};
        const int tag = __float_as_int(q1.z);
        const int pos = tag & 31;
        float* up = sh.in[tag >> 8].ac + pos;
        float Gr[M][NZ], hr[M];
        E::unpack_raw(w, p, Gr, hr);
        NormSolution<NZ, M> sol;
        solve_raw_fast<CarsPat, NZ, M>(Gr, hr, p.p_diag, false, sol);
#ifdef RCBF_C2_FORCE_PENDING  // test build (scripts/gpu_cars_pending.sh): ~1 solved instance in 16 takes the pending path
        if ((((unsigned)tag * 2654435761u + (unsigned)slot * 40503u) >> 28) == 0u) sol.status = RCBF_PENDING;
#endif
        if (sol.status == RCBF_PENDING) {
          *up = __uint_as_float(kPendingBits);
          sh.cls[(tag >> 6) & 3][pos] = (uint8_t)RCBF_PENDING;
          c_pend += 1;
        } else {
          *up = clampf(*up + (float)sol.x[0], p.u_min, p.u_max);   // :77
          sh.cls[(tag >> 6) & 3][pos] = (uint8_t)sol.status;
          if (!kFused) sh.amask[(tag >> 6) & 3][pos] = (uint16_t)sol.mask;
          c_iters += sol.iters;
        }
      }
      head += take;
      head -= head >= kC2Ring ? kC2Ring : 0;
      qn -= take;
      __syncwarp();
    }

    // ---------------------------------------------------------------- finish(tile k-2): in place, bulk stores out
    if (fin_due) {
      C2Warp::In& sf = sh.in[(k + 2) & 3];
      const int rf = r3 == 2 ? 0 : r3 + 1;   // (k - 2) % 3
      const int64_t i0 = (int64_t)tm2 << 5;
      const int64_t i = i0 + lane;
      const float us = sf.ac[lane];
      const bool pend = __float_as_uint(us) == kPendingBits;
      a.out[i] = us;
      if (a.status != nullptr) a.status[i] = (int)sh.cls[rf][lane];
      if (!kFused && a.meta != nullptr)  // (status << 16) | active set: what the backward kernel needs
        a.meta[i] = ((int)sh.cls[rf][lane] << 16) | (int)sh.amask[rf][lane];
      if constexpr (kFused) {
        float s[10];
        {
          const float2* sp = reinterpret_cast<const float2*>(sf.st) + lane * 5;
#pragma unroll
          for (int q = 0; q < 5; ++q) {
            const float2 s2 = sp[q];
            s[2 * q] = s2.x; s[2 * q + 1] = s2.y;
          }
        }
        float tt = sf.t[lane];
        int stp = sf.step[lane];
        CarsEnvOut<float> o;
        cars_env_step<float>(e, s, tt, stp, us, o);
        if (!pend) {   // a pending instance keeps its old state / t / step: the kernel's tail redoes it from scratch
          a.reward[i] = o.reward;
          a.done[i] = (uint8_t)o.done;
          a.cost[i] = o.cost;
          a.t[i] = tt;
          a.step[i] = stp;
        }
        if (lane == 0) bulk_wait_read0();   // the previous tile's observation rows have left sh.obs
        __syncwarp();
        {
          float2* wp = reinterpret_cast<float2*>(sf.st) + lane * 5;   // own row: nobody else reads or writes it
          float2* op = reinterpret_cast<float2*>(sh.obs) + lane * 5;
#pragma unroll
          for (int q = 0; q < 5; ++q) {
            if (!pend) wp[q] = make_float2(s[2 * q], s[2 * q + 1]);
            op[q] = make_float2(o.obs[2 * q], o.obs[2 * q + 1]);
          }
        }
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // generic-proxy writes -> visible to the bulk copies
        __syncwarp();
        if (lane == 0) {
          bulk_s2g(a.state + i0 * 10, sf.st, 1280);
          bulk_s2g(a.obs + i0 * 10, sh.obs, 1280);
          bulk_commit();
        }
      }
      if (__any_sync(0xffffffffu, pend)) {  // (rare) queue them once the tile's stores are complete
        if (kFused) {
          if (lane == 0) bulk_wait0();
          __syncwarp();
        }
        if (ws != nullptr && pend) {
          __threadfence();
          const unsigned long long slot = atomicAdd(&ws[kWsQueueCount], 1ULL);
          if (slot < (unsigned long long)kWsQueueCap) ws[kWsQueueBase + slot] = (unsigned long long)i + 1ULL;
        }
      }
    }
    if (after >= 2) break;
    r3 = r3 == 2 ? 0 : r3 + 1;
    tm2 = tm1;
    tm1 = tile;
    tile = tile1;
    tile1 += nw;
  }
  if (lane == 0) bulk_wait0();  // shared memory stays valid until the last bulk store has read it
  __syncwarp();
  pdl_launch_dependents();
  const bool own_tail = ws != nullptr;
  if (ws != nullptr) {
    c_nan = __reduce_add_sync(0xffffffffu, c_nan);
    const int c_triv = 32 * n_tiles - n_solve - c_nan;
    c_pend = __reduce_add_sync(0xffffffffu, c_pend);
    c_iters = __reduce_add_sync(0xffffffffu, c_iters);
    if (lane == 0) {
      if (c_nan) atomicAdd(&ws[0], (unsigned long long)c_nan);
      if (c_triv) atomicAdd(&ws[3], (unsigned long long)c_triv);
      if (c_iters) atomicAdd(&ws[4], (unsigned long long)c_iters);
      if (c_pend) atomicAdd(&ws[5], (unsigned long long)c_pend);
    }
  }
  if (own_tail) tail_drain<E>(a, p, e, ws, reinterpret_cast<unsigned short*>(&sh.ring[0][0]), lane);
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
        tail_scan<E>(a, n, p, e, ws, reinterpret_cast<unsigned short*>(&sh.ring[0][0]), lane, warp, kC2Warps);
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

// Launch k_cars2<kFused> on the full 32-instance tiles of the call.  *handled = number of leading instances it covers (0: the
// call does not qualify); the caller runs k_safe (launch_safe) on the ragged rest.
template <bool kFused>
inline int launch_cars2_tiles(const CarsArgs& a, int64_t n, const CarsParams& p, const CarsEnvParams& e,
                              rcbf_counters_t* ws, cudaStream_t s, int64_t* handled) {
  using E = CarsEnv<kFused>;
  *handled = 0;
  static const bool env_off = [] {
    const char* v = getenv("RCBF_NO_CARS2");
    return v != nullptr && v[0] == '1';
  }();
  if (env_off || n < RCBF_C2_MIN_N || n > 0x7fffffffLL || solver_mode_of(p) != 0) return 0;
  auto ok16 = [](const void* q) { return (reinterpret_cast<uintptr_t>(q) & 15) == 0; };
  if (a.x != nullptr || a.lam != nullptr || a.slack != nullptr || a.iters != nullptr) return 0;
  if (kFused) {
    if (a.meta != nullptr || !(ok16(a.state) && ok16(a.t) && ok16(a.step) && ok16(a.ac) && ok16(a.sg) && ok16(a.obs)))
      return 0;
  } else if (!(ok16(a.st) && ok16(a.ac) && ok16(a.sg))) {
    return 0;
  }
  const int64_t n2 = n & ~(int64_t)31;
  const int64_t ntiles = n2 >> 5;
  int dev = 0;
  cudaGetDevice(&dev);
  static bool configured[64] = {};
  if (!configured[dev & 63]) {
    cudaFuncSetAttribute(k_cars2<kFused>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(sizeof(C2Warp) * kC2Warps));
    configured[dev & 63] = true;
  }
  const int sms = device_sm_count();
  const int64_t want = (ntiles + kC2Warps - 1) / kC2Warps;
  const int resident = sms * RCBF_C2_MINB;
  const int grid = (int)(want < resident ? want : resident);
  CarsParams pk = p;
  if (n2 != n) pk.solver_mode = solver_mode_of(p);  // a ragged rest follows: that kernel publishes the counters
  cudaError_t err = launch_pdl(true, k_cars2<kFused>, grid, kC2Threads, sizeof(C2Warp) * kC2Warps, s, a, n2, pk, e, ws);
  if (err != cudaSuccess) return (int)err;
  if (ws == nullptr) {  // no workspace: a second kernel scans safe_action for the pending sentinel
    const int64_t fb = (n2 + 127) / 128;
    err = launch_pdl(true, k_safe_fallback<E, 0>, (int)(fb < sms * 4 ? fb : sms * 4), 128, 0, s, a, n2, p, e, ws);
    if (err != cudaSuccess) return (int)err;
  }
  *handled = n2;
  return 0;
}

}  // namespace rcbf
