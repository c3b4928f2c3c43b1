// rcbf_safe2.cuh -- k_safe2: the Unicycle hot kernel, TWO instances per lane through the Blackwell packed-FP32
// instructions (FMUL2 / FADD2 / FFMA2, rcbf_f2.cuh), TMA in AND out, whole-tile in-place finish.
//
// get_safe_action (diff_cbf_qp.py:44-79) and, fused, UnicycleEnv.step (unicycle_env.py:46-111) for 64-instance
// tiles; a lane owns the adjacent instances (2*lane, 2*lane+1) of its warp's tile.  Per persistent-loop iteration k:
//
//   A-step(tile k)   inputs arrive by cp.async.bulk (TMA, mbarrier completion) one iteration ahead; sin/cos, the
//                    reference-order constraint assembly and the trivial test run PACKED (one instruction = both
//                    instances, each half rounded exactly like the scalar instruction); trivial / NaN instances get
//                    their clamped action written in place of the nominal one in the tile's input slot; instances that
//                    need a solve push 15 words (Lg 5x2, h 5; the actuator rows are rebuilt from u) + a tag into the
//                    warp-private problem ring.
//   B-step           whenever >= 32 problems wait: one problem per lane, greedy active-set presolve + float64 KKT
//                    certificate (rcbf_core.cuh), the clamped safe action goes back into the owning tile's slot.
//   finish(tile k-2) two iterations later every problem of that tile is solved (FIFO ring; if the stream is too
//                    sparse in problems the oldest ones are flushed by a partial B-step), so the OWNING lanes finish
//                    the whole tile in place: clamp result, packed env.step, observation, reward, cost, done.  All
//                    outputs of the tile are staged in shared memory and leave through cp.async.bulk stores
//                    (shared -> global, bulk_group completion): perfectly coalesced, a handful of instructions.
//
// Compared with k_safe (one instance per lane, rcbf_safe_kernels.cuh) this halves the issue slots of the arithmetic,
// amortises the loop / ring / staging glue over 64 instances, removes the finish ring with its 28-byte re-read of
// every solved instance, and turns 14 strided stores per instance pair into 8 bulk copies per tile.
//
// Instances the B-step cannot certify (~1.5e-5) keep their OLD state, get the pending sentinel in safe_action[i][0] and
// are queued AFTER their tile's bulk stores completed; the kernel's tail (tail_drain of rcbf_safe_kernels.cuh: one
// warp per instance, exhaustive enumeration) finishes them -- still one launch per step.
//
// Launch conditions (launch_safe2): solver_mode 0, no saved tensors, every array base 16-byte aligned; n is split
// into full 64-instance tiles for this kernel and a ragged tail (< 64) for k_safe.
#pragma once

#include "rcbf_safe_kernels.cuh"

namespace rcbf {

#ifndef RCBF_S2_MINB
#define RCBF_S2_MINB 1     // resident blocks per SM (A/B on B200: 1 x 12 warps > 3 x 4 warps > 2 x 6 warps; at 16 warps
                           // per SM: 1 x 16 0.1272 ms, 2 x 8 (ring of 82 entries so that both fit) 0.1320 ms)
#endif
#ifndef RCBF_S2_WARPS
#define RCBF_S2_WARPS 16   // warps per block
#endif
#ifndef RCBF_S2_CARRY_SINCOS
#define RCBF_S2_CARRY_SINCOS 0  // 1: keep sin / cos of the heading from the assembly for env.step (1.5 KB of shared memory
#endif                          //    per warp); 0: recompute them in the finish (same function, same input: same bits)
#ifndef RCBF_S2_RING
#define RCBF_S2_RING 86    // problem ring entries
#endif
#ifndef RCBF_S2_MIN_N
#define RCBF_S2_MIN_N 4096  // below this the one-per-lane kernel spreads the few tiles over more warps
#endif
constexpr int kS2Warps = RCBF_S2_WARPS;
constexpr int kS2Threads = 32 * kS2Warps;
constexpr int kS2Ring = RCBF_S2_RING;  // a tile pushes <= 64 problems on top of the < 32 the B-steps left over; what does
                                       // not fit (<= 31 + 64 - kS2Ring) is parked in the landing slot of the next tile
constexpr bool kS2CarrySinCos = RCBF_S2_CARRY_SINCOS != 0;
static_assert(kS2Ring >= 75 && kS2Ring <= 95, "31 + 64 - kS2Ring parked problems must fit a landing slot (20 entries)");

// The presolve's row fetch for a problem that sits in the ring (rcbf_core.cuh: SelectRowFetch): a CBF row is read back
// from the slot (G[i][:2] = -Lg[i], G[i][2] = -1), an actuator row is rebuilt from its index.  Same values as the
// register-resident copy the default fetch would pick through a 9-way select chain.
struct S2RowFetch {
  const float* slot;  // Lg[5][2], h[5], tag
  __device__ __forceinline__ void operator()(int wi, const float G[kUniM][kUniNZ], const float h[kUniM], float g[kUniNZ],
                                             float& hh) const {
    if (wi < kUniHaz) {
      const float2 l = *reinterpret_cast<const float2*>(slot + 2 * wi);
      g[0] = -l.x;
      g[1] = -l.y;
      g[2] = -1.0f;
      hh = slot[2 * kUniHaz + wi];
    } else {  // rows 5..8 = +e_0, -e_0, +e_1, -e_1 on the controls (diff_cbf_qp.py:365-377)
      const int a = wi - kUniHaz;
      g[0] = a == 0 ? 1.0f : (a == 1 ? -1.0f : 0.0f);
      g[1] = a == 2 ? 1.0f : (a == 3 ? -1.0f : 0.0f);
      g[2] = 0.0f;
      hh = a == 0 ? h[kUniHaz] : (a == 1 ? h[kUniHaz + 1] : (a == 2 ? h[kUniHaz + 2] : h[kUniHaz + 3]));
    }
  }
};

template <bool kFused>
struct alignas(16) S2Warp {
  struct alignas(16) In {   // TMA landing slot of one tile; after the A-/B-steps `ac` holds the clamped SAFE action,
    float st[kFused ? 256 : 192];  // after the finish `st` / `step` hold the new state: all three leave from here
    int step[kFused ? 64 : 4];
    float ac[128];
  };
  In in[4];                 // tiles k-2 (finishing), k-1, k, k+1 (in flight)
  float mu[192], sg[192];   // only the assembly reads them: single buffer, refilled right after the A-step's read
  float sn[kS2CarrySinCos && kFused ? 3 : 1][kS2CarrySinCos && kFused ? 64 : 4];  // sin / cos of the heading, kept from
  float cs[kS2CarrySinCos && kFused ? 3 : 1][kS2CarrySinCos && kFused ? 64 : 4];  // the assembly for env.step (optional)
  float4 ring[kS2Ring][4];  // problem ring: Lg[5][2], h[5], tag
  // (the observation rows of the tile being finished, 64 x 28 B = exactly one In slot, are staged in that tile's own
  //  slot once its state / step / action have been read, and leave from there by one bulk store)
  uint8_t cls[3][64];       // per instance: RCBF_OK_TRIVIAL / RCBF_OK_CERTIFIED / RCBF_NAN / RCBF_PENDING
  uint16_t amask[kFused ? 1 : 3][kFused ? 2 : 64];  // layer-only kernel: active set of the certified solution (saved for the backward)
  uint64_t bar[2];
};

static_assert(sizeof(S2Warp<true>::In) == 64 * 7 * 4, "a fused input slot doubles as the staging of the tile's observation rows");
static_assert(sizeof(S2Warp<true>) * kS2Warps * RCBF_S2_MINB + 1024 * RCBF_S2_MINB <= 233472,
              "shared memory of the resident blocks (+ 1 KB reserved per block) must fit the 228 KB of an SM");
static_assert(sizeof(S2Warp<true>) * kS2Warps + 256 <= 232448, "dynamic + static shared memory of one block <= 227 KB");

#ifdef RCBF_S2_MAXNREG
#define RCBF_S2_BOUNDS __maxnreg__(RCBF_S2_MAXNREG)
#else
#define RCBF_S2_BOUNDS __launch_bounds__(kS2Threads, RCBF_S2_MINB)
#endif
template <bool kFused>
__global__ void RCBF_S2_BOUNDS
k_safe2(UniArgs a, int64_t n /* multiple of 64 */, UnicycleParams p, UnicycleEnvParams e, UniEnvF ef /* = make_env_f(e) */,
        rcbf_counters_t* ws) {
  using E = UniEnv<kFused>;
  using WS = S2Warp<kFused>;
  constexpr int NZ = kUniNZ, M = kUniM;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  WS& sh = reinterpret_cast<WS*>(smem_raw)[warp];
  const unsigned lt_mask = (1u << lane) - 1u;
  const int ntiles = (int)(n >> 6);
  const int nw = (int)gridDim.x * kS2Warps;
  // Tiles are dealt round robin over the resident warps.  (Tried on B200: claiming every tile after the first from an
  // atomic ticket in the workspace to even out the tail.  65 536 same-address atomics per launch serialise in L2 for
  // about as long as the kernel runs: 0.153 ms against 0.137 ms with the static deal.  Dealing warp-major -- tile =
  // warp * gridDim.x + blockIdx.x, so that the last, partial round spreads over all SMs instead of 100 of 148 -- loses
  // 2.4 %: the warps of a block then stream from 16 far-apart regions instead of one contiguous span; dealing only the
  // last round that way loses 1.7 %, and dealing the partial round FIRST and warp-major (initial values only, nothing added
  // to the loop) still loses 1 %: the tile-count imbalance between SMs is not what the tail of this kernel waits for.)
  int tile = (int)blockIdx.x * kS2Warps + warp;   // tile of iteration k
  int tile1 = tile + nw;                          // tile of iteration k + 1
  int tm1 = 0, tm2 = 0;                           // tiles of iterations k - 1, k - 2
  int head = 0, qn = 0;          // problem ring
  int pk = 0, pk1 = 0;           // problems pushed by tiles k, k-1 (the newest pk + pk1 entries of the ring)
  int r3 = 0;                    // k % 3
  int c_nan = 0, c_triv = 0, c_pend = 0, c_iters = 0;
  int n_solve = 0, n_tiles = 0;  // (warp-uniform) problems pushed / tiles assembled by this warp
  constexpr uint32_t kInBytes = (kFused ? 1024 + 256 : 768) + 512 + 768 + 768;

  // lane 0 puts tile t in flight into slot k & 3 (+ the mu / sigma buffer).  cp.async.bulk is a uniform-datapath
  // instruction (UBLKCP): one copy per execution, so per-lane descriptors would only turn into a loop over lanes.
  auto issue = [&](int t, int k) {
    typename WS::In& si = sh.in[k & 3];
    uint64_t* bar = &sh.bar[k & 1];
    const int64_t i0 = (int64_t)t << 6;
    mbar_expect_tx(bar, kInBytes);
    if (kFused) {
      bulk_g2s(si.st, a.state4 + i0 * 4, 1024, bar);
      bulk_g2s(si.step, a.step + i0, 256, bar);
    } else {
      bulk_g2s(si.st, a.st + i0 * 3, 768, bar);
    }
    bulk_g2s(si.ac, a.ac + i0 * 2, 512, bar);
    bulk_g2s(sh.mu, a.mu + i0 * 3, 768, bar);
    bulk_g2s(sh.sg, a.sg + i0 * 3, 768, bar);
  };

  pdl_wait();
  if (lane == 0) {
    mbar_init(&sh.bar[0], 1);
    mbar_init(&sh.bar[1], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncwarp();
  if (lane == 0 && tile < ntiles) issue(tile, 0);

  const float reset_dist = unicycle_reset_dist(ef);
  // per-instance class bytes: only the layer kernel (meta) and callers that want a status array need them
  const bool want_cls = !kFused || a.status != nullptr;
  int after = 0;  // iterations past this warp's last tile (the last two tiles are finished then)
#pragma unroll 1
  for (int k = 0;; ++k) {
    const bool have_tile = tile < ntiles;
    pk1 = pk;
    pk = 0;
    int n_aux = 0;  // problems of this tile parked outside the (full) ring
    if (have_tile) {  // ---------------------------------------------------------------- A-step
      ++n_tiles;
      typename WS::In& si = sh.in[k & 3];
      mbar_wait(&sh.bar[k & 1], (k >> 1) & 1);
      f2 st[3], u[2], mu[3], sg[3];
      if (kFused) {
        const float4 qa = reinterpret_cast<const float4*>(si.st)[2 * lane];
        const float4 qb = reinterpret_cast<const float4*>(si.st)[2 * lane + 1];
        st[0] = f2(qa.x, qb.x); st[1] = f2(qa.y, qb.y); st[2] = f2_pin(qa.z, qb.z);
      } else {
        const float2* sp = reinterpret_cast<const float2*>(si.st) + 3 * lane;
        const float2 s0 = sp[0], s1 = sp[1], s2 = sp[2];
        st[0] = f2(s0.x, s1.y); st[1] = f2(s0.y, s2.x); st[2] = f2_pin(s1.x, s2.y);
      }
      {
        const float4 uu = reinterpret_cast<const float4*>(si.ac)[lane];
        u[0] = f2_pin(uu.x, uu.z); u[1] = f2_pin(uu.y, uu.w);
        const float2* mp = reinterpret_cast<const float2*>(sh.mu) + 3 * lane;
        const float2 m0 = mp[0], m1 = mp[1], m2 = mp[2];
        mu[0] = f2(m0.x, m1.y); mu[1] = f2(m0.y, m2.x); mu[2] = f2_pin(m1.x, m2.y);
        const float2* gp = reinterpret_cast<const float2*>(sh.sg) + 3 * lane;
        const float2 g0 = gp[0], g1 = gp[1], g2 = gp[2];
        sg[0] = f2(g0.x, g1.y); sg[1] = f2(g0.y, g2.x); sg[2] = f2_pin(g1.x, g2.y);
      }
      __syncwarp();
      f2 sn, cs;
      sincos_v<f2>(st[2], &sn, &cs);
      if (kFused && kS2CarrySinCos) {
        reinterpret_cast<float2*>(sh.sn[r3])[lane] = make_float2(sn.lo(), sn.hi());
        reinterpret_cast<float2*>(sh.cs[r3])[lane] = make_float2(cs.lo(), cs.hi());
      }
      f2 Lg[kUniHaz][2], h[M];
      assemble_unicycle_v<f2>(p, st, sn, cs, u, mu, sg, Lg, h);
      // trivial test on the raw rows (h >= 0 on every row <=> x = 0 optimal) and NaN screen, per half
      // (a NaN in Lg[i] reaches h[i] through the Lg . u term whatever u is, so screening the 9 right-hand sides covers
      //  the 10 coefficients as well)
      bool triv[2], nan[2];
#pragma unroll
      for (int hh = 0; hh < 2; ++hh) {
        float hv[M];
#pragma unroll
        for (int r = 0; r < M; ++r) hv[r] = hh ? h[r].hi() : h[r].lo();
        classify_raw<M>(hv, triv[hh], nan[hh]);
      }
      const bool need0 = !triv[0] && !nan[0], need1 = !triv[1] && !nan[1];
      // trivial / NaN instances: the clamped action replaces the nominal one in the slot (diff_cbf_qp.py:77)
      {
        // clamp, then NaN for a NaN instance (torch.clamp propagates it; a NaN action makes its rows NaN, i.e. nan[hh])
        float2* acp = reinterpret_cast<float2*>(si.ac) + 2 * lane;
        if (!need0) {
          const float a0 = fminf(fmaxf(u[0].lo(), p.u_min[0]), p.u_max[0]), a1 = fminf(fmaxf(u[1].lo(), p.u_min[1]), p.u_max[1]);
          acp[0] = make_float2(nan[0] ? NAN : a0, nan[0] ? NAN : a1);
        }
        if (!need1) {
          const float a0 = fminf(fmaxf(u[0].hi(), p.u_min[0]), p.u_max[0]), a1 = fminf(fmaxf(u[1].hi(), p.u_min[1]), p.u_max[1]);
          acp[1] = make_float2(nan[1] ? NAN : a0, nan[1] ? NAN : a1);
        }
        if (want_cls) {
          uchar2 c2;
          c2.x = (unsigned char)(nan[0] ? RCBF_NAN : (need0 ? RCBF_OK_CERTIFIED : RCBF_OK_TRIVIAL));
          c2.y = (unsigned char)(nan[1] ? RCBF_NAN : (need1 ? RCBF_OK_CERTIFIED : RCBF_OK_TRIVIAL));
          reinterpret_cast<uchar2*>(sh.cls[r3])[lane] = c2;
        }
        if (!kFused) reinterpret_cast<uint32_t*>(sh.amask[r3])[lane] = 0u;
      }
      const unsigned b0 = __ballot_sync(0xffffffffu, need0), b1 = __ballot_sync(0xffffffffu, need1);
      const int n0 = __popc(b0);
      pk = n0 + __popc(b1);
      // The ring holds kS2Ring problems; the B-step loop leaves < 32, so a tile with more than kS2Ring - 31 problems may
      // not fit (rare: > 55 of 64 instances need a solve).  The excess (<= 9) is parked in the landing slot of tile k + 1
      // -- whose prefetch is then delayed until the first B-step below has made room -- and enters the ring there.
      const int cap = kS2Ring - qn;
      n_aux = pk > cap ? pk - cap : 0;
      // slot (k+1) & 3 belonged to the tile finished in the previous iteration; the bulk store of its observation rows
      // (staged in that slot) was issued then and has long read it
      if (lane == 0) {
        if (kFused && (n_aux != 0 || tile1 < ntiles)) bulk_wait_read0();
        if (n_aux == 0 && tile1 < ntiles) issue(tile1, k + 1);
      }
      if (n_aux != 0) __syncwarp();
      float4(*aux)[4] = reinterpret_cast<float4(*)[4]>(&sh.in[(k + 1) & 3]);
      const int tagbase = ((k & 3) << 8) | (r3 << 6) | (2 * lane);
#pragma unroll
      for (int hh = 0; hh < 2; ++hh) {
        if (hh ? need1 : need0) {
          const int idx = hh ? n0 + __popc(b1 & lt_mask) : __popc(b0 & lt_mask);
          int slot = head + qn + idx;                        // head < kS2Ring and qn + idx < kS2Ring when it is used
          slot -= slot >= kS2Ring ? kS2Ring : 0;
          float4* dst = idx < cap ? sh.ring[slot] : aux[idx - cap];
          float v[16];
#pragma unroll
          for (int r = 0; r < kUniHaz; ++r) {
            v[2 * r] = hh ? Lg[r][0].hi() : Lg[r][0].lo();
            v[2 * r + 1] = hh ? Lg[r][1].hi() : Lg[r][1].lo();
            v[10 + r] = hh ? h[r].hi() : h[r].lo();
          }
          v[15] = __int_as_float(tagbase + hh);
          // (the four 128-bit stores want adjacent registers: 16 register moves per problem.  Sixteen 32-bit stores straight
          //  from the halves of the packed registers instead: 0.1265 against 0.1263 ms, three interleaved runs -- no gain)
#pragma unroll
          for (int q = 0; q < 4; ++q) dst[q] = make_float4(v[4 * q], v[4 * q + 1], v[4 * q + 2], v[4 * q + 3]);
        }
      }
      qn += pk - n_aux;
      if (__any_sync(0xffffffffu, nan[0] || nan[1])) c_nan += (nan[0] ? 1 : 0) + (nan[1] ? 1 : 0);   // (rare)
      n_solve += pk;   // warp-uniform: trivial instances = 64 * tiles - solved - NaN, taken at the end
    } else {
      ++after;
    }
    __syncwarp();

    // ---------------------------------------------------------------- B-steps
    // full warps while >= 32 problems wait; then, if tile k-2 (finished below) still has problems in the ring (they are
    // the oldest entries), flush exactly those
    const bool fin_due = (k >= 2) && (after <= 2);
#pragma unroll 1
    for (;;) {
      const int old = qn - pk - pk1;
      // (a forced flush that takes the whole ring instead of only tile k-2's leftovers saves 1.4 % of the B-steps in a
      //  simulation of this deal and nothing measurable on B200: 0.12400 against 0.12394 ms, three interleaved runs each)
      const int take = qn >= 32 ? 32 : ((fin_due && old > 0) ? old : 0);
      if (take == 0) break;
      if (lane < take) {
        int slot = head + lane;
        slot -= slot >= kS2Ring ? kS2Ring : 0;
        float v[16];
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          const float4 t = sh.ring[slot][q];
          v[4 * q] = t.x; v[4 * q + 1] = t.y; v[4 * q + 2] = t.z; v[4 * q + 3] = t.w;
        }
        const int tag = __float_as_int(v[15]);
        const int pos = tag & 63;
        float2* up = reinterpret_cast<float2*>(sh.in[tag >> 8].ac) + pos;
        const float2 uu = *up;
        float w[E::NWR];
#pragma unroll
        for (int r = 0; r < 2 * kUniHaz; ++r) w[r] = -v[r];  // G[i][:2] = -Lg[i]   diff_cbf_qp.py:259
#pragma unroll
        for (int r = 0; r < kUniHaz; ++r) w[10 + r] = v[10 + r];
        w[15] = sub_rn(p.u_max[0], uu.x);                    // actuator rows, :365-377 (same expressions as the assembly)
        w[16] = add_rn(-p.u_min[0], uu.x);
        w[17] = sub_rn(p.u_max[1], uu.y);
        w[18] = add_rn(-p.u_min[1], uu.y);
        float Gr[M][NZ], hr[M];
        E::unpack_raw(w, p, Gr, hr);
        NormSolution<NZ, M> sol;
        solve_raw_fast<UniPat, NZ, M, S2RowFetch>(Gr, hr, p.p_diag, false, sol,
                                                  S2RowFetch{reinterpret_cast<const float*>(sh.ring[slot])});
        if (sol.status == RCBF_PENDING) {
          *up = make_float2(__uint_as_float(kPendingBits), 0.f);
          if (want_cls) sh.cls[(tag >> 6) & 3][pos] = (unsigned char)RCBF_PENDING;
          c_pend += 1;
        } else {
          *up = make_float2(clampf(uu.x + (float)sol.x[0], p.u_min[0], p.u_max[0]),
                            clampf(uu.y + (float)sol.x[1], p.u_min[1], p.u_max[1]));  // :77
          if (!kFused) sh.amask[(tag >> 6) & 3][pos] = (uint16_t)sol.mask;
          c_iters += sol.iters;
        }
      }
      head += take;
      head -= head >= kS2Ring ? kS2Ring : 0;
      qn -= take;
      __syncwarp();
      if (n_aux != 0) {  // (rare) the parked problems enter the ring, then the delayed prefetch of tile k + 1 goes out
        const float4(*aux)[4] = reinterpret_cast<const float4(*)[4]>(&sh.in[(k + 1) & 3]);
        if (lane < n_aux) {
          int slot = head + qn + lane;
          slot -= slot >= kS2Ring ? kS2Ring : 0;
#pragma unroll
          for (int q = 0; q < 4; ++q) sh.ring[slot][q] = aux[lane][q];
        }
        qn += n_aux;
        n_aux = 0;
        __syncwarp();
        if (lane == 0 && tile1 < ntiles) issue(tile1, k + 1);
      }
    }

    // ---------------------------------------------------------------- finish(tile k-2): whole tile, in place, TMA out
    if (fin_due) {
      const int ft = tm2;
      typename WS::In& sf = sh.in[(k + 2) & 3];
      const int rf = r3 == 2 ? 0 : r3 + 1;                 // (k - 2) % 3 == (k + 1) % 3
      const int64_t i0 = (int64_t)ft << 6;
      uchar2 cl = make_uchar2(0, 0);
      bool pend0, pend1;
      if (want_cls) {
        cl = reinterpret_cast<const uchar2*>(sh.cls[rf])[lane];
        pend0 = cl.x == RCBF_PENDING;
        pend1 = cl.y == RCBF_PENDING;
      } else {  // fused step without a status output: a pending instance is recognised by its sentinel in the action slot
        const float4 t4 = reinterpret_cast<const float4*>(sf.ac)[lane];
        pend0 = __float_as_uint(t4.x) == kPendingBits;
        pend1 = __float_as_uint(t4.z) == kPendingBits;
      }
      if (a.status != nullptr) reinterpret_cast<int2*>(a.status + i0)[lane] = make_int2(cl.x, cl.y);
      if (!kFused && a.meta != nullptr) {  // (status << 16) | active set: what the backward kernel needs
        const uint32_t am = reinterpret_cast<const uint32_t*>(sh.amask[rf])[lane];
        reinterpret_cast<int2*>(a.meta + i0)[lane] =
            make_int2(((int)cl.x << 16) | (int)(am & 0xffffu), ((int)cl.y << 16) | (int)(am >> 16));
      }
      if (kFused) {
        const float4 qa = reinterpret_cast<const float4*>(sf.st)[2 * lane];
        const float4 qb = reinterpret_cast<const float4*>(sf.st)[2 * lane + 1];
        const int2 sp = reinterpret_cast<const int2*>(sf.step)[lane];
        const float4 us4 = reinterpret_cast<const float4*>(sf.ac)[lane];
        f2 v[3] = {f2(qa.x, qb.x), f2(qa.y, qb.y), f2(qa.z, qb.z)};
        f2 snf, csf;
        if (kS2CarrySinCos) {
          const float2 sn2 = reinterpret_cast<const float2*>(sh.sn[rf])[lane];
          const float2 cs2 = reinterpret_cast<const float2*>(sh.cs[rf])[lane];
          snf = f2_pin(sn2.x, sn2.y);
          csf = f2_pin(cs2.x, cs2.y);
        } else {
          sincos_v<f2>(f2_pin(qa.z, qb.z), &snf, &csf);   // what the A-step computed from the same heading
        }
        f2 last(qa.w, qb.w);
        typename VecOf<f2>::ivec stp = {sp.x, sp.y};
        const f2 us[2] = {f2(us4.x, us4.z), f2(us4.y, us4.w)};
        UniEnvOutV<f2> o;
        unicycle_env_step_v<f2>(ef, v, last, stp, us, snf, csf, o);
        if (ef.auto_reset && __any_sync(0xffffffffu, o.done.x || o.done.y)) {  // (a finished episode is a rare event)
          v[0] = t_sel(o.done, f2(ef.init_x), v[0]);
          v[1] = t_sel(o.done, f2(ef.init_y), v[1]);
          v[2] = t_sel(o.done, f2(ef.init_th), v[2]);
          last = t_sel(o.done, f2(reset_dist), last);
          stp.x = o.done.x ? 0 : stp.x;
          stp.y = o.done.y ? 0 : stp.y;
        }
        // Outputs: every array except the 28-byte observation rows is a contiguous span per tile that the lanes cover
        // with one (two for state4) fully coalesced vector store; a pending instance keeps its OLD state (the kernel's
        // tail redoes it from scratch) and whatever else is written for it here is overwritten there.
        reinterpret_cast<float4*>(a.out + i0 * 2)[lane] = us4;
        float4* gs = reinterpret_cast<float4*>(a.state4 + i0 * 4) + 2 * lane;
        gs[0] = pend0 ? qa : make_float4(v[0].lo(), v[1].lo(), v[2].lo(), last.lo());
        gs[1] = pend1 ? qb : make_float4(v[0].hi(), v[1].hi(), v[2].hi(), last.hi());
        reinterpret_cast<int2*>(a.step + i0)[lane] = make_int2(pend0 ? sp.x : stp.x, pend1 ? sp.y : stp.y);
        reinterpret_cast<float2*>(a.reward + i0)[lane] = make_float2(o.reward.lo(), o.reward.hi());
        reinterpret_cast<float2*>(a.cost + i0)[lane] = make_float2(o.cost.lo(), o.cost.hi());
        uchar2 d2, g2;
        d2.x = o.done.x; d2.y = o.done.y;
        g2.x = o.goal_met.x; g2.y = o.goal_met.y;
        reinterpret_cast<uchar2*>(a.done + i0)[lane] = d2;
        reinterpret_cast<uchar2*>(a.goal_met + i0)[lane] = g2;
        __syncwarp();  // every lane has read its part of the slot: it now stages the observation rows
        float2* ob = reinterpret_cast<float2*>(&sf) + 7 * lane;
        ob[0] = make_float2(o.obs[0].lo(), o.obs[1].lo());
        ob[1] = make_float2(o.obs[2].lo(), o.obs[3].lo());
        ob[2] = make_float2(o.obs[4].lo(), o.obs[5].lo());
        ob[3] = make_float2(o.obs[6].lo(), o.obs[0].hi());
        ob[4] = make_float2(o.obs[1].hi(), o.obs[2].hi());
        ob[5] = make_float2(o.obs[3].hi(), o.obs[4].hi());
        ob[6] = make_float2(o.obs[5].hi(), o.obs[6].hi());
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // generic-proxy writes -> visible to the bulk copy
        __syncwarp();
        if (lane == 0) {
          bulk_s2g(a.obs + i0 * 7, &sf, 1792);
          bulk_commit();
        }
      } else {
        reinterpret_cast<float4*>(a.out + i0 * 2)[lane] = reinterpret_cast<const float4*>(sf.ac)[lane];
      }
      if (__any_sync(0xffffffffu, pend0 || pend1)) {  // (rare) queue them once the tile's stores are complete
        if (kFused) {
          if (lane == 0) bulk_wait0();
          __syncwarp();
        }
        if (ws != nullptr) {
#pragma unroll
          for (int hh = 0; hh < 2; ++hh) {
            if (hh ? pend1 : pend0) {
              __threadfence();
              const unsigned long long slot = atomicAdd(&ws[kWsQueueCount], 1ULL);
              if (slot < (unsigned long long)kWsQueueCap)
                ws[kWsQueueBase + slot] = (unsigned long long)(i0 + 2 * lane + hh) + 1ULL;
            }
          }
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
  if (kFused && lane == 0) bulk_wait0();  // shared memory stays valid until the last bulk store has read it
  __syncwarp();
  pdl_launch_dependents();
  const bool own_tail = ws != nullptr;
  // (the counters go out BEFORE the tail: nothing of this function's state is live across the call)
  if (ws != nullptr) {
    c_nan = __reduce_add_sync(0xffffffffu, c_nan);
    c_triv = 64 * n_tiles - n_solve - c_nan;
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
        tail_scan<E>(a, n, p, e, ws, reinterpret_cast<unsigned short*>(&sh.ring[0][0]), lane, warp, kS2Warps);
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

// Launch the two-per-lane kernel on the full 64-instance tiles of the call.  *handled = number of leading instances it
// covers (0: the call does not qualify); the caller runs the one-per-lane kernel (launch_safe) on the ragged rest.
template <bool kFused>
inline int launch_safe2_tiles(const UniArgs& a, int64_t n, const UnicycleParams& p, const UnicycleEnvParams& e,
                              rcbf_counters_t* ws, cudaStream_t s, int64_t* handled) {
  using E = UniEnv<kFused>;
  *handled = 0;
  static const bool env_off = [] {
    const char* v = getenv("RCBF_NO_SAFE2");
    return v != nullptr && v[0] == '1';
  }();
  if (env_off || n < RCBF_S2_MIN_N || n > 0x7fffffffLL || solver_mode_of(p) != 0) return 0;
  auto ok16 = [](const void* q) { return (reinterpret_cast<uintptr_t>(q) & 15) == 0; };
  if (a.x != nullptr || a.lam != nullptr || a.slack != nullptr || a.iters != nullptr) return 0;
  if (a.meta != nullptr && (kFused || !ok16(a.meta))) return 0;
  if (!(ok16(a.ac) && ok16(a.mu) && ok16(a.sg) && ok16(a.out) && (a.status == nullptr || ok16(a.status)))) return 0;
  if (kFused) {
    if (!(ok16(a.state4) && ok16(a.step) && ok16(a.obs) && ok16(a.reward) && ok16(a.cost) && ok16(a.done) &&
          ok16(a.goal_met)))
      return 0;
  } else if (!ok16(a.st)) {
    return 0;
  }
  const int64_t n2 = n & ~(int64_t)63;
  const int64_t ntiles = n2 >> 6;
  int dev = 0;
  cudaGetDevice(&dev);
  static int sm_count[64] = {};
  if (sm_count[dev & 63] == 0) {
    cudaFuncSetAttribute(k_safe2<kFused>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                         (int)(sizeof(S2Warp<kFused>) * kS2Warps));
    cudaDeviceGetAttribute(&sm_count[dev & 63], cudaDevAttrMultiProcessorCount, dev);
  }
  const int sms = sm_count[dev & 63];
  const int64_t want = (ntiles + kS2Warps - 1) / kS2Warps;
  const int resident = sms * RCBF_S2_MINB;
  const int grid = (int)(want < resident ? want : resident);
  UnicycleParams pk = p;
  if (n2 != n) pk.solver_mode = solver_mode_of(p);  // a ragged rest follows: that kernel publishes the counters
  cudaError_t err =
      launch_pdl(true, k_safe2<kFused>, grid, kS2Threads, sizeof(S2Warp<kFused>) * kS2Warps, s, a, n2, pk, e, make_env_f(e), ws);
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
