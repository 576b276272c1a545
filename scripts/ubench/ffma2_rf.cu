// microbenchmark: does FFMA2 sustain 1 per 2 cycles when all three 64-bit operands differ from instruction to instruction (depthwise inner loop)?
#include <cstdio>
#include <cuda_runtime.h>
typedef unsigned long long u64;
__device__ __forceinline__ void ffma2(float2& d, const float2& a, const float2& b) {
  asm volatile("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(reinterpret_cast<u64&>(d)) : "l"(reinterpret_cast<const u64&>(a)), "l"(reinterpret_cast<const u64&>(b)));
}
template <int MODE>
__global__ void k(float* out, int iters, long long* cyc) {
  float2 acc[5][4], x[8], w[25];
  for (int i = 0; i < 20; ++i) acc[i / 4][i % 4] = make_float2(threadIdx.x * 1e-3f + i, i * 0.5f);
  for (int i = 0; i < 8; ++i) x[i] = make_float2(1.0f + i * 1e-3f + threadIdx.x * 1e-6f, 0.999f - i * 1e-3f);
  for (int i = 0; i < 25; ++i) w[i] = make_float2(1e-3f * i, -1e-3f * i + threadIdx.x * 1e-7f);
  long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
    if (MODE == 0) {          // conv order: for q, kx, o
#pragma unroll
      for (int q = 0; q < 5; ++q)
#pragma unroll
        for (int kx = 0; kx < 5; ++kx)
#pragma unroll
          for (int o = 0; o < 4; ++o) ffma2(acc[q][o], x[o + kx], w[q * 5 + kx]);
    } else {                  // same count, two operands fixed (reuse cache friendly)
#pragma unroll
      for (int q = 0; q < 5; ++q)
#pragma unroll
        for (int kx = 0; kx < 5; ++kx)
#pragma unroll
          for (int o = 0; o < 4; ++o) ffma2(acc[q][o], x[0], w[0]);
    }
  }
  long long t1 = clock64();
  float s = 0;
  for (int i = 0; i < 20; ++i) s += acc[i / 4][i % 4].x + acc[i / 4][i % 4].y;
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
}
int main() {
  float* out; long long* cyc; cudaMalloc(&out, 1 << 24); cudaMalloc(&cyc, 8);
  const int iters = 4000;
  for (int mode = 0; mode < 2; ++mode)
    for (int warps : {4, 8, 16}) {
      long long h;
      for (int rep = 0; rep < 2; ++rep) {
        if (mode == 0) k<0><<<148, warps * 32>>>(out, iters, cyc); else k<1><<<148, warps * 32>>>(out, iters, cyc);
        cudaDeviceSynchronize();
      }
      cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
      double inst = (double)iters * 100 * (warps / 4.0);
      printf("%s warps/SMSP %d: %.3f cycles per FFMA2 per SMSP\n", mode == 0 ? "conv-pattern " : "fixed-operand", warps / 4, h / inst);
    }
  return 0;
}
