// rcbf_gp.cu -- posterior of the disturbance GPs (mean, std for every state dimension) in one launch.
//
// Replaces GPyDisturbanceEstimator.predict (rcbf_sac/gp_model.py:86-114) and the fitted branch of
// DynamicsModel.predict_disturbance (rcbf_sac/dynamics.py:371-379).  The math is in include/rcbf_b200.h next to
// rcbf_gp_posterior; the host side (sac_rcbf_b200/gp_model.py) builds the factor F with F^T F = (K + noise I)^-1.
//
// Work decomposition: block (x, y) owns 32 test points of GP y.  For each 32-point chunk of the training set
//   1. TMA (cp.async.bulk + mbarrier, double buffered) brings the chunk's normalised inputs and the matching
//      [32][tile_rows] slab of F^T into shared memory while the previous chunk is being consumed;
//   2. the 128 threads evaluate the 32 x 32 kernel block k(z_c, z*_t) in float64 (one exp each, 8 per thread) into
//      shared memory -- K* is never written to HBM;
//   3. a register-tiled float64 rank-32 update  W[r][t] += F^T[c][r] * k[c][t]  ((tile_rows / 16) x 4 per thread).
// After the last chunk the block reduces |W[:, t]|^2 and W[:, t] . proj_y and writes mean / std of its 32 test points.
// That kernel serves factors with 64-row tiles (kernels that are not numerically low rank); k_gp_predict_lowrank below
// serves 4 / 8 / 16-row tiles.  Factors with more rows than one tile are walked tile by tile (the kernel block is
// recomputed per tile: with 64 rows
// per tile the exps cost about as much as the FMAs, so nothing is gained by spilling K* to HBM).
#include <cstdint>

#include <cuda_runtime.h>

#include "../../include/rcbf_b200.h"
#include "rcbf_tma.cuh"

namespace rcbf {
namespace {

constexpr int kGpTile = 32;     // test points per block
constexpr int kGpChunk = 32;    // training points per pipeline stage
constexpr int kGpThreads = 128;

template <int RT, int DP>
struct GpSmem {
  double fac[2][kGpChunk * RT];   // F^T slab   [c][r]
  double z[2][kGpChunk * DP];     // train inputs [c][k]
  double k[kGpChunk * kGpTile];   // kernel block [c][t]
  double red[2][4][kGpTile];      // cross-warp reduction of |w|^2 and w . proj_y
  uint64_t bar[2];
};

template <int RT, int DP, typename T>
__global__ void __launch_bounds__(kGpThreads) k_gp_predict(rcbf_gp_posterior p, const T* __restrict__ test_x,
                                                           int64_t n_test, T* __restrict__ mean,
                                                           T* __restrict__ sd) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  GpSmem<RT, DP>& sh = *reinterpret_cast<GpSmem<RT, DP>*>(smem_raw);
  constexpr int RM = RT / 16;  // factor rows per thread
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int gp = blockIdx.y;
  const int64_t t0 = (int64_t)blockIdx.x * kGpTile;
  const int64_t xs = p.test_stride ? p.test_stride : p.n_in;

  const double inv_2l2 = p.hyp[gp * 4 + 0], os = p.hyp[gp * 4 + 1], noise = p.hyp[gp * 4 + 2],
               y_scale = p.hyp[gp * 4 + 3];
  const int n_chunks = p.n_pad / kGpChunk;
  const int r_tiles = p.r_tiles[gp];
  const int total = r_tiles * n_chunks;
  const double* fac_gp = p.factor + (size_t)gp * p.max_tiles * p.n_pad * RT;
  const double* py_gp = p.proj_y + (size_t)gp * p.max_tiles * RT;

  if (tid == 0) {
    mbar_init(&sh.bar[0], 1);
    mbar_init(&sh.bar[1], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();

  auto issue = [&](int it) {  // thread 0 only
    const int rt = it / n_chunks, ch = it - rt * n_chunks, s = it & 1;
    constexpr uint32_t fac_bytes = kGpChunk * RT * 8, z_bytes = kGpChunk * DP * 8;
    mbar_expect_tx(&sh.bar[s], fac_bytes + z_bytes);
    bulk_g2s(sh.fac[s], fac_gp + ((size_t)rt * p.n_pad + (size_t)ch * kGpChunk) * RT, fac_bytes, &sh.bar[s]);
    bulk_g2s(sh.z[s], p.train_z + (size_t)ch * kGpChunk * DP, z_bytes, &sh.bar[s]);
  };
  if (tid == 0 && total > 0) issue(0);

  // this lane's test point, normalised (dynamics.py:375), zero padded to DP coordinates
  double zt[DP];
  {
    const int64_t t = t0 + lane;
#pragma unroll
    for (int k = 0; k < DP; ++k)
      zt[k] = (k < p.n_in && t < n_test) ? (double)test_x[t * xs + k] * p.inv_x_scale[k] : 0.0;
  }

  const int tr = tid >> 3, tt = tid & 7;
  double q[4] = {0, 0, 0, 0}, m[4] = {0, 0, 0, 0};
  double acc[RM][4];

  for (int it = 0; it < total; ++it) {
    const int rt = it / n_chunks, ch = it - rt * n_chunks, s = it & 1;
    if (ch == 0) {
#pragma unroll
      for (int a = 0; a < RM; ++a)
#pragma unroll
        for (int b = 0; b < 4; ++b) acc[a][b] = 0.0;
    }
    if (tid == 0 && it + 1 < total) issue(it + 1);  // stage (it+1)&1 was released by the barrier ending it-1
    mbar_wait(&sh.bar[s], (it >> 1) & 1);

    // kernel block: warp w evaluates training points c = w, w+4, ...; lane = test point
#pragma unroll
    for (int j = 0; j < kGpChunk / 4; ++j) {
      const int c = warp + 4 * j;
      const double* zc = &sh.z[s][c * DP];
      double d2 = 0.0;
#pragma unroll
      for (int k = 0; k < DP; ++k) {
        const double d = zc[k] - zt[k];
        d2 = fma(d, d, d2);
      }
      sh.k[c * kGpTile + lane] = os * exp(-d2 * inv_2l2);
    }
    __syncthreads();

    const double* fs = sh.fac[s];
#pragma unroll 8
    for (int c = 0; c < kGpChunk; ++c) {
      double fv[RM], kv[4];
#pragma unroll
      for (int a = 0; a < RM; ++a) fv[a] = fs[c * RT + tr * RM + a];
#pragma unroll
      for (int b = 0; b < 4; ++b) kv[b] = sh.k[c * kGpTile + tt * 4 + b];
#pragma unroll
      for (int a = 0; a < RM; ++a)
#pragma unroll
        for (int b = 0; b < 4; ++b) acc[a][b] = fma(fv[a], kv[b], acc[a][b]);
    }
    __syncthreads();

    if (ch == n_chunks - 1) {  // row tile finished: fold it into |w|^2 and w . proj_y
#pragma unroll
      for (int a = 0; a < RM; ++a) {
        const double py = py_gp[rt * RT + tr * RM + a];
#pragma unroll
        for (int b = 0; b < 4; ++b) {
          q[b] = fma(acc[a][b], acc[a][b], q[b]);
          m[b] = fma(acc[a][b], py, m[b]);
        }
      }
    }
  }

  // reduce over the 16 row groups: lanes with equal tt inside the warp (xor 8, 16), then across the 4 warps
#pragma unroll
  for (int b = 0; b < 4; ++b) {
    q[b] += __shfl_xor_sync(0xffffffffu, q[b], 8);
    q[b] += __shfl_xor_sync(0xffffffffu, q[b], 16);
    m[b] += __shfl_xor_sync(0xffffffffu, m[b], 8);
    m[b] += __shfl_xor_sync(0xffffffffu, m[b], 16);
  }
  if (lane < 8) {
#pragma unroll
    for (int b = 0; b < 4; ++b) {
      sh.red[0][warp][tt * 4 + b] = q[b];
      sh.red[1][warp][tt * 4 + b] = m[b];
    }
  }
  __syncthreads();
  if (tid < kGpTile) {
    const int64_t t = t0 + tid;
    if (t < n_test) {
      const double qq = (sh.red[0][0][tid] + sh.red[0][1][tid]) + (sh.red[0][2][tid] + sh.red[0][3][tid]);
      const double mm = (sh.red[1][0][tid] + sh.red[1][1][tid]) + (sh.red[1][2][tid] + sh.red[1][3][tid]);
      double var = os - qq + (p.include_noise ? noise : 0.0);
      var = var > p.min_variance ? var : p.min_variance;  // NaN -> min_variance is NOT wanted: keep NaN visible
      if (qq != qq) var = qq;
      mean[t * p.n_gp + gp] = (T)(mm * y_scale);
      sd[t * p.n_gp + gp] = (T)(sqrt(var) * y_scale);
    }
  }
}


// ---------------------------------------------------------------------------------------------------------------
// Low-rank variant (factor row tiles of R = 4 / 8 / 16 rows -- the reference's pinned lengthscale 1e5 gives rank <= 4
// for the Unicycle and ~11 for SimulatedCars): no K* staging at all.  Block = 32 test points x 8 warps; each warp
// takes 8 of the 64 training points of a chunk, a lane evaluates k(z_c, z*_lane) and immediately folds it into its
// R private accumulators  w[r] += F^T[c][r] * k  (F^T[c][.] is a shared-memory broadcast).  One barrier per chunk
// (stage release); the 8 partial w's are summed through shared memory once per row tile.
// ---------------------------------------------------------------------------------------------------------------
constexpr int kLrWarps = 8;
constexpr int kLrChunk = 64;
constexpr int kLrPerWarp = kLrChunk / kLrWarps;

template <int R, int DP>
struct GpLrSmem {
  double fac[2][kLrChunk * R];
  double z[2][kLrChunk * DP];
  double red[kLrWarps][R][kGpTile];
  uint64_t bar[2];
};

template <int R, int DP, typename T>
__global__ void __launch_bounds__(kLrWarps * 32) k_gp_predict_lowrank(rcbf_gp_posterior p, const T* __restrict__ test_x,
                                                                      int64_t n_test, T* __restrict__ mean,
                                                                      T* __restrict__ sd) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  GpLrSmem<R, DP>& sh = *reinterpret_cast<GpLrSmem<R, DP>*>(smem_raw);
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int gp = blockIdx.y;
  const int64_t t0 = (int64_t)blockIdx.x * kGpTile;
  const int64_t xs = p.test_stride ? p.test_stride : p.n_in;
  const double inv_2l2 = p.hyp[gp * 4 + 0], os = p.hyp[gp * 4 + 1], noise = p.hyp[gp * 4 + 2],
               y_scale = p.hyp[gp * 4 + 3];
  const int n_chunks = p.n_pad / kLrChunk;
  const int r_tiles = p.r_tiles[gp];
  const int total = r_tiles * n_chunks;
  const double* fac_gp = p.factor + (size_t)gp * p.max_tiles * p.n_pad * R;
  const double* py_gp = p.proj_y + (size_t)gp * p.max_tiles * R;

  if (tid == 0) {
    mbar_init(&sh.bar[0], 1);
    mbar_init(&sh.bar[1], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  auto issue = [&](int it) {  // thread 0 only
    const int rt = it / n_chunks, ch = it - rt * n_chunks, s = it & 1;
    constexpr uint32_t fac_bytes = kLrChunk * R * 8, z_bytes = kLrChunk * DP * 8;
    mbar_expect_tx(&sh.bar[s], fac_bytes + z_bytes);
    bulk_g2s(sh.fac[s], fac_gp + ((size_t)rt * p.n_pad + (size_t)ch * kLrChunk) * R, fac_bytes, &sh.bar[s]);
    bulk_g2s(sh.z[s], p.train_z + (size_t)ch * kLrChunk * DP, z_bytes, &sh.bar[s]);
  };
  if (tid == 0 && total > 0) issue(0);

  double zt[DP];
  {
    const int64_t t = t0 + lane;
#pragma unroll
    for (int k = 0; k < DP; ++k)
      zt[k] = (k < p.n_in && t < n_test) ? (double)test_x[t * xs + k] * p.inv_x_scale[k] : 0.0;
  }
  double w[R];
  double q = 0.0, m = 0.0;  // meaningful in warp 0 only

  for (int it = 0; it < total; ++it) {
    const int rt = it / n_chunks, ch = it - rt * n_chunks, s = it & 1;
    if (ch == 0) {
#pragma unroll
      for (int r = 0; r < R; ++r) w[r] = 0.0;
    }
    if (tid == 0 && it + 1 < total) issue(it + 1);
    mbar_wait(&sh.bar[s], (it >> 1) & 1);
#pragma unroll
    for (int j = 0; j < kLrPerWarp; ++j) {
      const int c = warp * kLrPerWarp + j;
      const double* zc = &sh.z[s][c * DP];
      double d2 = 0.0;
#pragma unroll
      for (int k = 0; k < DP; ++k) {
        const double d = zc[k] - zt[k];
        d2 = fma(d, d, d2);
      }
      const double kv = os * exp(-d2 * inv_2l2);
      const double* fc = &sh.fac[s][c * R];
#pragma unroll
      for (int r = 0; r < R; ++r) w[r] = fma(fc[r], kv, w[r]);
    }
    if (ch == n_chunks - 1) {  // row tile finished: sum the 8 partial w's, fold into |w|^2 and w . proj_y
#pragma unroll
      for (int r = 0; r < R; ++r) sh.red[warp][r][lane] = w[r];
      __syncthreads();
      if (warp == 0) {
#pragma unroll
        for (int r = 0; r < R; ++r) {
          double ws = 0.0;
#pragma unroll
          for (int v = 0; v < kLrWarps; ++v) ws += sh.red[v][r][lane];
          q = fma(ws, ws, q);
          m = fma(ws, py_gp[rt * R + r], m);
        }
      }
    }
    __syncthreads();  // releases stage s (and sh.red) for reuse
  }
  if (warp == 0) {
    const int64_t t = t0 + lane;
    if (t < n_test) {
      double var = os - q + (p.include_noise ? noise : 0.0);
      var = var > p.min_variance ? var : p.min_variance;
      if (q != q) var = q;  // keep a NaN visible instead of clamping it away
      mean[t * p.n_gp + gp] = (T)(m * y_scale);
      sd[t * p.n_gp + gp] = (T)(sqrt(var) * y_scale);
    }
  }
}


// ---------------------------------------------------------------------------------------------------------------
// Far-field fast path (see rcbf_gp_posterior::ff_coef): one lane per test point, no loop over the training set.
// Lanes whose point fails the validity bound are finished exactly by their warp: the 32 lanes split the training
// set, evaluate the kernel and F k* for that one point and shuffle-reduce -- same float64 arithmetic as the kernels
// above, so a bank whose test points are all "near" is merely slower, never wrong.
// ---------------------------------------------------------------------------------------------------------------
template <int R, int DP, typename T>
__global__ void __launch_bounds__(128) k_gp_farfield(rcbf_gp_posterior p, const T* __restrict__ test_x, int64_t n_test,
                                                     T* __restrict__ mean, T* __restrict__ sd) {
  constexpr int NC = 3 + 2 * DP + DP * (DP + 1) / 2;
  constexpr int P = DP <= 4 ? 4 : 2;  // test points per lane: every coefficient load feeds P FMAs
  constexpr unsigned kFull = 0xffffffffu;
  const int lane = threadIdx.x & 31;
  const int64_t n_tiles = (n_test + 32 * P - 1) / (32 * P);
  const int64_t xs = p.test_stride ? p.test_stride : p.n_in;
  const int64_t warps = ((int64_t)gridDim.x * blockDim.x) >> 5;
  for (int64_t tile = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5; tile < n_tiles; tile += warps) {
    int64_t t[P];
    bool valid[P];
    double zt[P][DP], s[P], reach[P];
#pragma unroll
    for (int i = 0; i < P; ++i) {
      t[i] = (tile * P + i) * 32 + lane;
      valid[i] = t[i] < n_test;
      s[i] = 0.0;
#pragma unroll
      for (int k = 0; k < DP; ++k) {
        zt[i][k] = (k < p.n_in && valid[i]) ? (double)test_x[t[i] * xs + k] * __ldg(p.inv_x_scale + k) : 0.0;
        s[i] = fma(zt[i][k], zt[i][k], s[i]);
      }
      reach[i] = sqrt(s[i]) + p.ff_zmax;
    }
    for (int g = 0; g < p.n_gp; ++g) {
      const double inv_2l2 = __ldg(p.hyp + g * 4 + 0), os = __ldg(p.hyp + g * 4 + 1), noise = __ldg(p.hyp + g * 4 + 2),
                   y_scale = __ldg(p.hyp + g * 4 + 3), amax = __ldg(p.ff_amax + g);
      const double* py = p.proj_y + (size_t)g * R;
      const double* cf = p.ff_coef + (size_t)g * R * NC;
      double q[P], m[P];
#pragma unroll
      for (int i = 0; i < P; ++i) q[i] = m[i] = 0.0;
#pragma unroll(R <= 4 ? R : 2)
      for (int r = 0; r < R; ++r) {
        const double* c = cf + r * NC;
        double w[P];
        {
          const double c0 = __ldg(c), c1 = __ldg(c + 1), c2 = __ldg(c + 2);
#pragma unroll
          for (int i = 0; i < P; ++i) w[i] = fma(s[i], fma(s[i], c2, c1), c0);
        }
#pragma unroll
        for (int k = 0; k < DP; ++k) {
          const double lin = __ldg(c + 3 + k), slin = __ldg(c + 3 + DP + k);
#pragma unroll
          for (int i = 0; i < P; ++i) w[i] = fma(zt[i][k], fma(s[i], slin, lin), w[i]);
        }
        int idx = 3 + 2 * DP;
#pragma unroll
        for (int k = 0; k < DP; ++k) {
          double row[P];
#pragma unroll
          for (int i = 0; i < P; ++i) row[i] = 0.0;
#pragma unroll
          for (int l = k; l < DP; ++l) {
            const double cq = __ldg(c + idx++);
#pragma unroll
            for (int i = 0; i < P; ++i) row[i] = fma(zt[i][l], cq, row[i]);
          }
#pragma unroll
          for (int i = 0; i < P; ++i) w[i] = fma(zt[i][k], row[i], w[i]);
        }
        const double pyr = __ldg(py + r);
#pragma unroll
        for (int i = 0; i < P; ++i) {
          q[i] = fma(w[i], w[i], q[i]);
          m[i] = fma(w[i], pyr, m[i]);
        }
      }
#pragma unroll
      for (int i = 0; i < P; ++i) {
        unsigned near = __ballot_sync(kFull, valid[i] && !(reach[i] * reach[i] * inv_2l2 <= amax));
        while (near) {  // exact evaluation of one point by the whole warp
          const int src = __ffs(near) - 1;
          near &= near - 1;
          double zs[DP];
#pragma unroll
          for (int k = 0; k < DP; ++k) zs[k] = __shfl_sync(kFull, zt[i][k], src);
          double w[R];
#pragma unroll
          for (int r = 0; r < R; ++r) w[r] = 0.0;
          const double* fac = p.factor + (size_t)g * p.max_tiles * p.n_pad * R;
          for (int j = lane; j < p.n_pad; j += 32) {
            double d2 = 0.0;
#pragma unroll
            for (int k = 0; k < DP; ++k) {
              const double d = __ldg(p.train_z + (size_t)j * DP + k) - zs[k];
              d2 = fma(d, d, d2);
            }
            const double kv = os * exp(-d2 * inv_2l2);
#pragma unroll
            for (int r = 0; r < R; ++r) w[r] = fma(__ldg(fac + (size_t)j * R + r), kv, w[r]);
          }
          double qe = 0.0, me = 0.0;
#pragma unroll
          for (int r = 0; r < R; ++r) {
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) w[r] += __shfl_xor_sync(kFull, w[r], o);
            qe = fma(w[r], w[r], qe);
            me = fma(w[r], __ldg(py + r), me);
          }
          if (lane == src) {
            q[i] = qe;
            m[i] = me;
          }
        }
        if (valid[i]) {
          double var = os - q[i] + (p.include_noise ? noise : 0.0);
          var = var > p.min_variance ? var : p.min_variance;
          if (q[i] != q[i]) var = q[i];
          mean[t[i] * p.n_gp + g] = (T)(m[i] * y_scale);
          sd[t[i] * p.n_gp + g] = (T)(sqrt(var) * y_scale);
        }
      }
    }
  }
}

template <int R, int DP, typename T>
int launch_gp_farfield(const rcbf_gp_posterior& p, const T* test_x, int64_t n_test, T* mean, T* sd, cudaStream_t s) {
  constexpr int P = DP <= 4 ? 4 : 2;
  const int64_t tiles = (n_test + 32 * P - 1) / (32 * P);
  const int64_t want = (tiles + 3) / 4;
  static int sms_cached[64] = {};
  int dev = 0;
  cudaGetDevice(&dev);
  int& sms = sms_cached[dev & 63];
  if (sms == 0) cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  const int cap = (sms > 0 ? sms : 148) * 16;
  const int grid = (int)(want < cap ? want : cap);
  k_gp_farfield<R, DP, T><<<grid, 128, 0, s>>>(p, test_x, n_test, mean, sd);
  return (int)cudaGetLastError();
}

template <int RT, int DP, typename T>
int launch_gp_one(const rcbf_gp_posterior& p, const T* test_x, int64_t n_test, T* mean, T* sd, cudaStream_t s) {
  static bool attr_done[64] = {};
  int dev = 0;
  cudaGetDevice(&dev);
  constexpr bool kLowRank = RT <= 16;
  if constexpr (kLowRank) {
    if (p.ff_coef && p.ff_amax && p.max_tiles == 1) return launch_gp_farfield<RT, DP, T>(p, test_x, n_test, mean, sd, s);
  }
  const int smem = kLowRank ? (int)sizeof(GpLrSmem<kLowRank ? RT : 16, DP>) : (int)sizeof(GpSmem<RT, DP>);
  if (dev >= 0 && dev < 64 && !attr_done[dev]) {
    cudaError_t e;
    if constexpr (kLowRank)
      e = cudaFuncSetAttribute(k_gp_predict_lowrank<RT, DP, T>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    else
      e = cudaFuncSetAttribute(k_gp_predict<RT, DP, T>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    if (e != cudaSuccess) return (int)e;
    attr_done[dev] = true;
  }
  const int64_t tiles = (n_test + kGpTile - 1) / kGpTile;
  int64_t done = 0;
  while (done < tiles) {  // grid.x limit 2^31-1 is far away; chunk anyway so grid.x * 32 stays in int64 math
    const int64_t now = tiles - done < (1 << 24) ? tiles - done : (1 << 24);
    dim3 grid((unsigned)now, (unsigned)p.n_gp);
    const int64_t off = done * kGpTile;
    if constexpr (kLowRank)
      k_gp_predict_lowrank<RT, DP, T><<<grid, kLrWarps * 32, smem, s>>>(p, test_x + off * (p.test_stride ? p.test_stride : p.n_in), n_test - off,
                                                                       mean + off * p.n_gp, sd + off * p.n_gp);
    else
      k_gp_predict<RT, DP, T><<<grid, kGpThreads, smem, s>>>(p, test_x + off * (p.test_stride ? p.test_stride : p.n_in), n_test - off,
                                                            mean + off * p.n_gp, sd + off * p.n_gp);
    done += now;
  }
  return (int)cudaGetLastError();
}

template <typename T>
int launch_gp(const rcbf_gp_posterior* ph, const T* test_x, int64_t n_test, T* mean, T* sd, void* stream) {
  if (!ph || n_test < 0) return (int)cudaErrorInvalidValue;
  const rcbf_gp_posterior& p = *ph;
  if (p.n_pad <= 0 || p.n_pad % kLrChunk || p.n_in <= 0 || p.n_in > p.dim_pad || p.n_gp <= 0 || p.n_gp > 65535 ||
      p.max_tiles <= 0 || (p.test_stride != 0 && p.test_stride < p.n_in))
    return (int)cudaErrorInvalidValue;
  if (n_test == 0) return 0;
  cudaStream_t s = (cudaStream_t)stream;
#define RCBF_GP_CASE(RT, DP) \
  if (p.tile_rows == RT && p.dim_pad == DP) return launch_gp_one<RT, DP, T>(p, test_x, n_test, mean, sd, s);
  RCBF_GP_CASE(4, 4)
  RCBF_GP_CASE(4, 8)
  RCBF_GP_CASE(4, 12)
  RCBF_GP_CASE(4, 16)
  RCBF_GP_CASE(8, 4)
  RCBF_GP_CASE(8, 8)
  RCBF_GP_CASE(8, 12)
  RCBF_GP_CASE(8, 16)
  RCBF_GP_CASE(16, 4)
  RCBF_GP_CASE(16, 8)
  RCBF_GP_CASE(16, 12)
  RCBF_GP_CASE(16, 16)
  RCBF_GP_CASE(64, 4)
  RCBF_GP_CASE(64, 8)
  RCBF_GP_CASE(64, 12)
  RCBF_GP_CASE(64, 16)
#undef RCBF_GP_CASE
  return (int)cudaErrorInvalidValue;
}

__global__ void k_fp64_fma_probe(double* sink, int iters) {
  double a0 = threadIdx.x * 1e-3, a1 = a0 + 1., a2 = a0 + 2., a3 = a0 + 3., a4 = a0 + 4., a5 = a0 + 5., a6 = a0 + 6.,
         a7 = a0 + 7.;
  const double b = 0.999 + blockIdx.x * 1e-9, c = 1e-3;
#pragma unroll 4
  for (int k = 0; k < iters; ++k) {
    a0 = fma(a0, b, c); a1 = fma(a1, b, c); a2 = fma(a2, b, c); a3 = fma(a3, b, c);
    a4 = fma(a4, b, c); a5 = fma(a5, b, c); a6 = fma(a6, b, c); a7 = fma(a7, b, c);
  }
  const double r = a0 + a1 + a2 + a3 + a4 + a5 + a6 + a7;
  if (r == 123.456) sink[0] = r;  // never true; keeps the chains alive
}

}  // namespace
}  // namespace rcbf

extern "C" {

int rcbf_gp_predict_f32(const float* test_x, int64_t n_test, const rcbf_gp_posterior* post_host, float* mean,
                        float* std, void* stream) {
  return rcbf::launch_gp<float>(post_host, test_x, n_test, mean, std, stream);
}
int rcbf_gp_predict_f64(const double* test_x, int64_t n_test, const rcbf_gp_posterior* post_host, double* mean,
                        double* std, void* stream) {
  return rcbf::launch_gp<double>(post_host, test_x, n_test, mean, std, stream);
}
int rcbf_fp64_fma_probe(double* sink, int blocks, int threads, int iters, void* stream) {
  rcbf::k_fp64_fma_probe<<<blocks, threads, 0, (cudaStream_t)stream>>>(sink, iters);
  return (int)cudaGetLastError();
}

}  // extern "C"
