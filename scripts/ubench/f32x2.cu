// Microbenchmark: issue cost of packed fp32 (FFMA2 / FMUL2 / FADD2) vs scalar FFMA on sm_100a, alone and mixed with
// ALU (integer / logic) instructions.  Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o f32x2 f32x2.cu
#include <cstdio>
#include <cuda_runtime.h>
typedef unsigned long long u64;
__device__ __forceinline__ u64 pk(float a, float b) { u64 r; asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(a), "f"(b)); return r; }
__device__ __forceinline__ void upk(u64 v, float& a, float& b) { asm("mov.b64 {%0, %1}, %2;" : "=f"(a), "=f"(b) : "l"(v)); }
__device__ __forceinline__ u64 fma2(u64 a, u64 b, u64 c) { u64 r; asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c)); return r; }
__device__ __forceinline__ u64 mul2(u64 a, u64 b) { u64 r; asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b)); return r; }
__device__ __forceinline__ u64 add2(u64 a, u64 b) { u64 r; asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b)); return r; }

template <int MODE>
__global__ void k(float* out, int iters, float s) {
  float acc[16];
#pragma unroll
  for (int i = 0; i < 16; ++i) acc[i] = threadIdx.x * 0.001f + i;
  u64 p[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) p[i] = pk(acc[2 * i], acc[2 * i + 1]);
  const u64 ss = pk(s, s), cc = pk(0.5f, 0.25f);
  unsigned z[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) z[i] = threadIdx.x + i;
  for (int it = 0; it < iters; ++it) {
    if (MODE == 0) {  // 16 scalar FFMA
#pragma unroll
      for (int i = 0; i < 16; ++i) acc[i] = fmaf(acc[i], s, 0.5f);
    } else if (MODE == 1) {  // 8 FFMA2 (same flops)
#pragma unroll
      for (int i = 0; i < 8; ++i) p[i] = fma2(p[i], ss, cc);
    } else if (MODE == 2) {  // 16 scalar FFMA + 8 ALU (xor/add chain)
#pragma unroll
      for (int i = 0; i < 16; ++i) acc[i] = fmaf(acc[i], s, 0.5f);
#pragma unroll
      for (int i = 0; i < 8; ++i) z[i] = (z[i] ^ (z[i] >> 3)) + 0x9e3779b9u;
    } else if (MODE == 3) {  // 8 FFMA2 + 8x(2 ALU)
#pragma unroll
      for (int i = 0; i < 8; ++i) p[i] = fma2(p[i], ss, cc);
#pragma unroll
      for (int i = 0; i < 8; ++i) z[i] = (z[i] ^ (z[i] >> 3)) + 0x9e3779b9u;
    } else if (MODE == 4) {  // 8 FMUL2 + 8 FADD2
#pragma unroll
      for (int i = 0; i < 8; ++i) p[i] = add2(mul2(p[i], ss), cc);
    } else if (MODE == 5) {  // 16 FMUL + 16 FADD scalar
#pragma unroll
      for (int i = 0; i < 16; ++i) acc[i] = __fadd_rn(__fmul_rn(acc[i], s), 0.5f);
    } else if (MODE == 6) {  // dependent chain latency: 1 FFMA2 chain
      p[0] = fma2(p[0], ss, cc);
    } else if (MODE == 7) {  // dependent chain latency: 1 FFMA chain
      acc[0] = fmaf(acc[0], s, 0.5f);
    }
  }
  float r = 0.f;
#pragma unroll
  for (int i = 0; i < 8; ++i) { float a, b; upk(p[i], a, b); r += a + b + acc[2 * i] + acc[2 * i + 1] + (float)z[i]; }
  out[blockIdx.x * blockDim.x + threadIdx.x] = r;
}

template <int MODE>
void run(const char* name, float* out, int blocks, int threads, double flops_per_iter_thread) {
  const int iters = 20000;
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  k<MODE><<<blocks, threads>>>(out, 100, 0.999f);
  cudaDeviceSynchronize();
  cudaEventRecord(e0);
  k<MODE><<<blocks, threads>>>(out, iters, 0.999f);
  cudaEventRecord(e1);
  cudaDeviceSynchronize();
  float ms; cudaEventElapsedTime(&ms, e0, e1);
  const double th = (double)blocks * threads;
  printf("%-34s blocks=%d thr=%d  %.3f ms  %.2f TFLOP/s  %.3f ns/iter/warp-slot\n", name, blocks, threads, ms,
         flops_per_iter_thread * iters * th / ms / 1e9, ms * 1e6 / iters);
}

int main() {
  float* out; cudaMalloc(&out, 148 * 16 * 1024 * 4);
  for (int warps : {4, 8, 16}) {   // warps per SM (one block per SM)
    const int thr = warps * 32;
    printf("--- %d warps / SM\n", warps);
    run<0>("16 FFMA", out, 148, thr, 32);
    run<1>("8 FFMA2", out, 148, thr, 32);
    run<2>("16 FFMA + 16 ALU", out, 148, thr, 32);
    run<3>("8 FFMA2 + 16 ALU", out, 148, thr, 32);
    run<4>("8 FMUL2 + 8 FADD2", out, 148, thr, 32);
    run<5>("16 FMUL + 16 FADD", out, 148, thr, 32);
  }
  run<6>("FFMA2 dependent chain, 1 warp/SM", out, 148, 32, 4);
  run<7>("FFMA dependent chain, 1 warp/SM", out, 148, 32, 2);
  return 0;
}
