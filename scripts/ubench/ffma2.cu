// microbenchmark: issue rate of fma.rn.f32x2 (FFMA2) vs fma.rn.f32 (FFMA) per SM sub-partition, by resident warps
#include <cstdio>
#include <cuda_runtime.h>
typedef unsigned long long u64;
template <int MODE>
__global__ void k(float* out, int iters, long long* cyc) {
  float2 a[16];
  for (int i = 0; i < 16; ++i) a[i] = make_float2(threadIdx.x * 1e-3f + i, i * 0.5f);
  float2 b = make_float2(1.0001f, 0.9999f), c = make_float2(1e-3f, -1e-3f);
  long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 16; ++i) {
      if (MODE == 0) asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(reinterpret_cast<u64&>(a[i])) : "l"(reinterpret_cast<u64&>(b)), "l"(reinterpret_cast<u64&>(c)));
      else { asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(a[i].x) : "f"(b.x), "f"(c.x)); asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(a[i].y) : "f"(b.y), "f"(c.y)); }
    }
  }
  long long t1 = clock64();
  float s = 0;
  for (int i = 0; i < 16; ++i) s += a[i].x + a[i].y;
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
}
int main() {
  float* out; long long* cyc; cudaMalloc(&out, 1 << 24); cudaMalloc(&cyc, 8);
  const int iters = 20000;
  for (int mode = 0; mode < 2; ++mode)
    for (int warps : {4, 8, 16, 32}) {       // warps per SM (1 block per SM)
      long long h;
      for (int rep = 0; rep < 2; ++rep) {
        if (mode == 0) k<0><<<148, warps * 32>>>(out, iters, cyc); else k<1><<<148, warps * 32>>>(out, iters, cyc);
        cudaDeviceSynchronize();
      }
      cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
      double per_smsp_warps = warps / 4.0;
      double inst = (double)iters * 16 * (mode == 0 ? 1 : 2) * per_smsp_warps;     // warp-instructions per SMSP
      printf("%s warps/SMSP %.0f: %lld cycles, %.3f cycles per warp-instruction per SMSP, %.1f FMA lanes/clk/SMSP\n", mode == 0 ? "FFMA2" : "FFMA ", per_smsp_warps, h,
             h / inst, inst * (mode == 0 ? 64 : 32) / h);
    }
  return 0;
}
