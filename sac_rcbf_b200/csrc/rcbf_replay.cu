// rcbf_replay.cu -- device-resident replay ring (rcbf_sac/replay_memory.py:4-35).
//
// The reference keeps a Python list of (state, action, reward, next_state, mask, t, next_t) tuples, pushes one item per
// loop iteration (replay_memory.py:20-26) and draws `random.sample(buffer, batch_size)` (:30) followed by seven
// np.stack calls (:31).  Here the ring is seven preallocated device arrays (one per tuple field, row-major) and both
// operations are ONE launch each:
//
//   k_replay_push    rows i = 0..n-1 of the seven source arrays go to ring row (position + i) % capacity
//                    (what n successive push() calls do, :14-18; the caller has already dropped all but the newest
//                    `capacity` rows).
//   k_replay_sample  output row i comes from ring row perm_key(i): a keyed bijection of [0, size) (Feistel network on
//                    ceil(log2 size) bits, cycle-walked into the range), so the first `batch` values are `batch`
//                    distinct indices -- sampling WITHOUT replacement like random.sample, in O(batch) work and with no
//                    index array in HBM.  Index draw and the gather of all seven fields are fused: the block computes
//                    the indices of its 64 rows into shared memory, then every thread moves 4-byte words, consecutive
//                    threads writing consecutive words of the (contiguous) output arrays.
//
// Both kernels are HBM / latency bound byte movers; rows are handled as 4-byte words (a float64 ring has twice the words
// per row).  Layout decides the traffic of a random gather: with one array per field a drawn row touches nine 32-byte
// sectors for 80 bytes of payload, with the fields of a transition adjacent (ROW-MAJOR ring: field[f] points into one
// (capacity, row_stride) matrix, stride padded to a sector multiple) it touches three.  For that layout the tiled
// kernels (k_replay_*_rows) move whole rows as 16-byte vectors between the ring and a shared-memory tile of 64 rows
// and convert between row-major (ring) and field-major (the caller's batch arrays) inside the tile, so both sides are
// fully coalesced.  Any other layout (separate arrays, odd strides) takes the word-granular kernels.
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/rcbf_b200.h"

namespace rcbf {

constexpr int kRpFields = RCBF_REPLAY_FIELDS;
constexpr int kRpRows = 64;      // sample rows per block
constexpr int kRpThreads = 256;

struct RpDesc {
  uint32_t* ring[kRpFields];
  uint32_t* io[kRpFields];  // sources (push) or outputs (sample); a null entry is skipped
  int w[kRpFields];         // 4-byte words per row
  int64_t rs[kRpFields];    // 4-byte words between consecutive ring rows of the field
  int w0[kRpFields + 1];    // prefix sums: word j of the concatenated row belongs to field f with w0[f] <= j < w0[f+1]
  uint32_t mw[kRpFields];   // floor(2^32 / w[f]) + 1: q / w[f] == __umulhi(q, mw[f]) for the q < 2^16 of a tile
  uint32_t mv4;             // the same for the row stride in 16-byte vectors
};

__device__ __forceinline__ int rp_div(int q, uint32_t magic, int d) { return d == 1 ? q : (int)__umulhi((uint32_t)q, magic); }

__device__ __forceinline__ uint32_t rp_mix(uint32_t x) {  // murmur3 finaliser
  x ^= x >> 16;
  x *= 0x85ebca6bu;
  x ^= x >> 13;
  x *= 0xc2b2ae35u;
  x ^= x >> 16;
  return x;
}

// Keyed bijection of [0, size): balanced Feistel network over 2 * hb bits (2^(2 hb) >= size), 6 rounds, cycle-walked.
// i < size, so walking the cycle that starts at i always comes back into the range.
__host__ __device__ __forceinline__ uint64_t rp_rotation(uint64_t key, int64_t size) {
  uint64_t o = key * 0x9e3779b97f4a7c15ull;
  o ^= o >> 29;
  o *= 0xbf58476d1ce4e5b9ull;
  o ^= o >> 32;
  return o % (uint64_t)size;
}

__device__ __forceinline__ int64_t rp_perm(int64_t i, int64_t size, int hb, uint64_t key, uint64_t rot) {
  const uint32_t hm = (hb >= 32) ? 0xffffffffu : ((1u << hb) - 1u);
  const uint32_t k0 = (uint32_t)key, k1 = (uint32_t)(key >> 32);
  uint64_t v = (uint64_t)i;
  do {
    uint32_t l = (uint32_t)(v >> hb) & hm, r = (uint32_t)v & hm;
#pragma unroll
    for (int rd = 0; rd < 6; ++rd) {
      const uint32_t f = rp_mix(r + k0 * (2u * rd + 1u) + rp_mix(k1 + 0x9e3779b9u * (uint32_t)(rd + 1))) & hm;
      const uint32_t nl = r;
      r = l ^ f;
      l = nl;
    }
    v = ((uint64_t)l << hb) | (uint64_t)r;
  } while (v >= (uint64_t)size);
  // a key-derived rotation on top (rot = rp_rotation(key, size), computed once on the host): whatever structure a
  // narrow network leaves (2-bit halves for a 7-row ring), every output slot is exactly uniform over the rows
  v += rot;
  v -= v >= (uint64_t)size ? (uint64_t)size : 0ull;
  return (int64_t)v;
}

__device__ __forceinline__ int rp_field(const RpDesc& d, int j) {
  int f = 0;
#pragma unroll
  for (int q = 1; q < kRpFields; ++q) f += (j >= d.w0[q]) ? 1 : 0;
  return f;
}

__global__ void __launch_bounds__(kRpThreads)
k_replay_push(RpDesc d, int64_t capacity, int64_t position, int64_t n) {
  const int W = d.w0[kRpFields];
  const int64_t total = n * W;
  for (int64_t g = (int64_t)blockIdx.x * kRpThreads + threadIdx.x; g < total; g += (int64_t)gridDim.x * kRpThreads) {
    const int64_t i = g / W;
    const int j = (int)(g - i * W);
    const int f = rp_field(d, j);
    if (d.io[f] == nullptr) continue;
    int64_t row = position + i;
    row -= row >= capacity ? capacity : 0;   // position < capacity and n <= capacity
    const int c = j - d.w0[f];
    d.ring[f][row * d.rs[f] + c] = d.io[f][i * d.w[f] + c];
  }
}

__global__ void __launch_bounds__(kRpThreads)
k_replay_sample(RpDesc d, int64_t size, int64_t batch, int hb, uint64_t key, uint64_t rot, int64_t* idx_out) {
  __shared__ int64_t s_idx[kRpRows];
  const int W = d.w0[kRpFields];
  for (int64_t r0 = (int64_t)blockIdx.x * kRpRows; r0 < batch; r0 += (int64_t)gridDim.x * kRpRows) {
    const int rows = (int)(batch - r0 < kRpRows ? batch - r0 : kRpRows);
    if ((int)threadIdx.x < rows) {
      const int64_t ix = rp_perm(r0 + threadIdx.x, size, hb, key, rot);
      s_idx[threadIdx.x] = ix;
      if (idx_out != nullptr) idx_out[r0 + threadIdx.x] = ix;
    }
    __syncthreads();
    // field-major over the block's rows: consecutive threads write consecutive words of one output array
    for (int f = 0; f < kRpFields; ++f) {
      if (d.io[f] == nullptr) continue;
      const int w = d.w[f];
      const int cnt = rows * w;
      uint32_t* out = d.io[f] + r0 * w;
      const uint32_t* ring = d.ring[f];
      const int64_t rs = d.rs[f];
      for (int q = threadIdx.x; q < cnt; q += kRpThreads) {
        const int rr = q / w, c = q - rr * w;
        out[q] = __ldg(ring + s_idx[rr] * rs + c);
      }
    }
    __syncthreads();
  }
}


// ---- row-major ring: whole rows through a shared-memory tile ---------------------------------------------------------
constexpr int kRpMaxStride = 64;    // words (256 B) per ring row: the widest row the tile takes
constexpr int kRpSampleTileWords = 4096;  // 16 KB: 168 rows of the Unicycle's 96-byte stride (64 of the widest one)
constexpr int kRpPushTileWords = 6144;    // 24 KB: 256 rows
constexpr int kRpMaxRows = 256;
// Bytes in flight per barrier are what bounds these kernels.  4 Mi-row draw + gather on B200: 319 us with 64-row tiles,
// 214 us with 168 rows, 270 us with 256 rows (the 24 KB tile leaves the random row fetches 20 KB of L1); 4 Mi-row push:
// 317 / 200 / 179 us.
// (Tried on B200: a barrier-free form -- G lanes per row, one 16-byte vector of the ring row per lane, the four words
//  scattered to / gathered from the field arrays by 4-byte accesses.  440 us per 4 Mi rows against 319 us: the
//  sector-partial accesses to seven arrays cost more than the barriers they remove.)

// ring rows (sw words apart, 16-byte aligned) of the drawn indices -> tile -> the seven field-major outputs
__global__ void __launch_bounds__(kRpThreads)
k_replay_sample_rows(RpDesc d, const uint4* __restrict__ rows_base, int sw, int R /* rows per tile */, int64_t size,
                     int64_t batch, int hb, uint64_t key, uint64_t rot, int64_t* idx_out) {
  __shared__ int64_t s_idx[kRpMaxRows];
  __shared__ __align__(16) uint32_t tile[kRpSampleTileWords];
  const int v4 = sw >> 2;
  for (int64_t r0 = (int64_t)blockIdx.x * R; r0 < batch; r0 += (int64_t)gridDim.x * R) {
    const int rows = (int)(batch - r0 < R ? batch - r0 : R);
    if ((int)threadIdx.x < rows) {
      const int64_t ix = rp_perm(r0 + threadIdx.x, size, hb, key, rot);
      s_idx[threadIdx.x] = ix;
      if (idx_out != nullptr) idx_out[r0 + threadIdx.x] = ix;
    }
    __syncthreads();
    for (int q = threadIdx.x; q < rows * v4; q += kRpThreads) {
      const int rr = rp_div(q, d.mv4, v4), c = q - rr * v4;
      reinterpret_cast<uint4*>(tile)[q] = __ldg(rows_base + s_idx[rr] * v4 + c);
    }
    __syncthreads();
#pragma unroll
    for (int f = 0; f < kRpFields; ++f) {
      if (d.io[f] == nullptr) continue;
      const int w = d.w[f], off = d.w0[f];
      uint32_t* out = d.io[f] + r0 * w;   // (R is a multiple of 4: r0 * w words = a multiple of 16 bytes past the base)
      const int cnt = rows * w;
      const int nv = (reinterpret_cast<uintptr_t>(d.io[f]) & 15) == 0 ? cnt >> 2 : 0;   // 16-byte stores
      for (int q = threadIdx.x; q < nv; q += kRpThreads) {
        int rr = rp_div(4 * q, d.mw[f], w), c = 4 * q - rr * w;
        uint32_t v[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          v[j] = tile[rr * sw + off + c];
          if (++c == w) c = 0, ++rr;
        }
        reinterpret_cast<uint4*>(out)[q] = make_uint4(v[0], v[1], v[2], v[3]);
      }
      for (int q = 4 * nv + threadIdx.x; q < cnt; q += kRpThreads) {
        const int rr = rp_div(q, d.mw[f], w), c = q - rr * w;
        out[q] = tile[rr * sw + off + c];
      }
    }
    __syncthreads();
  }
}

// the seven field-major sources (all present) -> tile -> ring rows (position + i) % capacity
__global__ void __launch_bounds__(kRpThreads)
k_replay_push_rows(RpDesc d, uint4* __restrict__ rows_base, int sw, int R /* rows per tile */, int64_t capacity,
                   int64_t position, int64_t n) {
  __shared__ __align__(16) uint32_t tile[kRpPushTileWords];
  const int v4 = sw >> 2;
  const int W = d.w0[kRpFields];
  for (int64_t r0 = (int64_t)blockIdx.x * R; r0 < n; r0 += (int64_t)gridDim.x * R) {
    const int rows = (int)(n - r0 < R ? n - r0 : R);
#pragma unroll
    for (int f = 0; f < kRpFields; ++f) {
      const int w = d.w[f], off = d.w0[f];
      const uint32_t* src = d.io[f] + r0 * w;
      const int cnt = rows * w;
      const int nv = (reinterpret_cast<uintptr_t>(d.io[f]) & 15) == 0 ? cnt >> 2 : 0;   // 16-byte loads
      for (int q = threadIdx.x; q < nv; q += kRpThreads) {
        const uint4 t = __ldg(reinterpret_cast<const uint4*>(src) + q);
        const uint32_t v[4] = {t.x, t.y, t.z, t.w};
        int rr = rp_div(4 * q, d.mw[f], w), c = 4 * q - rr * w;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          tile[rr * sw + off + c] = v[j];
          if (++c == w) c = 0, ++rr;
        }
      }
      for (int q = 4 * nv + threadIdx.x; q < cnt; q += kRpThreads) {
        const int rr = rp_div(q, d.mw[f], w), c = q - rr * w;
        tile[rr * sw + off + c] = __ldg(src + q);
      }
    }
    for (int q = threadIdx.x; q < rows * (sw - W); q += kRpThreads) {  // the padding words of a row stay zero
      const int rr = q / (sw - W), c = q - rr * (sw - W);
      tile[rr * sw + W + c] = 0u;
    }
    __syncthreads();
    for (int q = threadIdx.x; q < rows * v4; q += kRpThreads) {
      const int rr = rp_div(q, d.mv4, v4), c = q - rr * v4;
      int64_t row = position + r0 + rr;
      row -= row >= capacity ? capacity : 0;
      rows_base[row * v4 + c] = reinterpret_cast<const uint4*>(tile)[q];
    }
    __syncthreads();
  }
}

// rows per tile for a stride of sw words: as many as the tile holds, a multiple of 4 (16-byte aligned field chunks)
inline int rp_tile_rows(int sw, int tile_words) {
  int r = tile_words / sw;
  r = r > kRpMaxRows ? kRpMaxRows : r;
  return r & ~3;
}

// Is the ring one row-major matrix the tiled kernels can move?  -> stride in words (0: no)
inline int rp_row_major(const RpDesc& d) {
  const int64_t sw = d.rs[0];
  if (sw <= 0 || sw > kRpMaxStride || (sw & 3) != 0 || sw < d.w0[kRpFields]) return 0;
  if ((reinterpret_cast<uintptr_t>(d.ring[0]) & 15) != 0) return 0;
  for (int f = 0; f < kRpFields; ++f)
    if (d.rs[f] != sw || d.ring[f] != d.ring[0] + d.w0[f]) return 0;
  return (int)sw;
}

inline int rp_fill(RpDesc& d, const rcbf_replay_ring* r, void* const io[]) {
  const int e = r->elem_bytes / 4;
  const int obs_words = r->obs_dim * e, act_words = r->action_dim * e, scalar_words = e;
  void* const* ring = r->field;
  const int w[kRpFields] = {obs_words, act_words, scalar_words, obs_words, scalar_words, scalar_words, scalar_words};
  d.w0[0] = 0;
  for (int f = 0; f < kRpFields; ++f) {
    if (ring[f] == nullptr || w[f] <= 0) return cudaErrorInvalidValue;
    d.ring[f] = static_cast<uint32_t*>(ring[f]);
    d.io[f] = static_cast<uint32_t*>(io[f]);
    d.w[f] = w[f];
    d.rs[f] = r->row_stride[f] > 0 ? r->row_stride[f] * e : w[f];
    if (d.rs[f] < w[f]) return cudaErrorInvalidValue;
    d.w0[f + 1] = d.w0[f] + w[f];
    d.mw[f] = (uint32_t)(0x100000000ull / (uint64_t)w[f]) + 1u;
  }
  const int64_t sw4 = d.rs[0] >> 2;
  d.mv4 = sw4 > 0 ? (uint32_t)(0x100000000ull / (uint64_t)sw4) + 1u : 0u;
  return 0;
}

inline int rp_sms() {
  int dev = 0, sms = 0;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  return sms > 0 ? sms : 148;
}

}  // namespace rcbf

extern "C" {

int rcbf_replay_push(const rcbf_replay_ring* ring, int64_t position, const void* const src[RCBF_REPLAY_FIELDS],
                     int64_t n, void* stream) {
  using namespace rcbf;
  if (ring == nullptr || n < 0 || n > ring->capacity || position < 0 || position >= ring->capacity ||
      (ring->elem_bytes != 4 && ring->elem_bytes != 8))
    return cudaErrorInvalidValue;
  if (n == 0) return 0;
  RpDesc d;
  if (int rc = rp_fill(d, ring, const_cast<void* const*>(src))) return rc;
  const int64_t cap = (int64_t)rp_sms() * 8;
  bool all_src = true;
  for (int f = 0; f < kRpFields; ++f) all_src = all_src && src[f] != nullptr;
  if (const int sw = all_src ? rp_row_major(d) : 0) {
    const int R = rp_tile_rows(sw, kRpPushTileWords);
    const int64_t want = (n + R - 1) / R;
    k_replay_push_rows<<<(int)(want < cap ? want : cap), kRpThreads, 0, static_cast<cudaStream_t>(stream)>>>(
        d, reinterpret_cast<uint4*>(d.ring[0]), sw, R, ring->capacity, position, n);
    return (int)cudaGetLastError();
  }
  const int64_t total = n * d.w0[kRpFields];
  const int64_t want = (total + kRpThreads - 1) / kRpThreads;
  k_replay_push<<<(int)(want < cap ? want : cap), kRpThreads, 0, static_cast<cudaStream_t>(stream)>>>(
      d, ring->capacity, position, n);
  return (int)cudaGetLastError();
}

int rcbf_replay_sample(const rcbf_replay_ring* ring, int64_t size, int64_t batch, uint64_t key,
                       void* const out[RCBF_REPLAY_FIELDS], int64_t* idx_out, void* stream) {
  using namespace rcbf;
  if (ring == nullptr || size <= 0 || size > ring->capacity || batch < 0 || batch > size ||
      (ring->elem_bytes != 4 && ring->elem_bytes != 8))
    return cudaErrorInvalidValue;
  if (batch == 0) return 0;
  RpDesc d;
  if (int rc = rp_fill(d, ring, out)) return rc;
  int bits = 2;
  while (bits < 62 && ((int64_t)1 << bits) < size) ++bits;
  bits += bits & 1;  // balanced halves
  const int64_t want = (batch + kRpRows - 1) / kRpRows;
  const int64_t cap = (int64_t)rp_sms() * 8;
  if (const int sw = rp_row_major(d)) {
    const int R = rp_tile_rows(sw, kRpSampleTileWords);
    const int64_t wantr = (batch + R - 1) / R;
    k_replay_sample_rows<<<(int)(wantr < cap ? wantr : cap), kRpThreads, 0, static_cast<cudaStream_t>(stream)>>>(
        d, reinterpret_cast<const uint4*>(d.ring[0]), sw, R, size, batch, bits / 2, key, rp_rotation(key, size), idx_out);
  }
  else
    k_replay_sample<<<(int)(want < cap ? want : cap), kRpThreads, 0, static_cast<cudaStream_t>(stream)>>>(
        d, size, batch, bits / 2, key, rp_rotation(key, size), idx_out);
  return (int)cudaGetLastError();
}

}  // extern "C"
