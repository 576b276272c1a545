// Depthwise k3/k5 s1/s2 and stem 3x3 s2 convolutions (forward + backward), NHWC bf16, explicit static pads.
//
// Replaces the cuDNN/ATen kernels behind Conv2dStaticSamePadding (efficient_net_custom_utils.py:248-276) for
//   MBConvBlock._depthwise_conv (efficientnet_custom.py:66-73,109) and EfficientNet._conv_stem (:174-176,273),
// fused with what surrounds them in the reference graph:
//   * the producer's train-mode BatchNorm + swish is applied WHILE LOADING the input tile (y -> swish(a*y+b));
//     zero padding is applied after it, exactly like ZeroPad2d on the activated tensor (utils:268-275)
//   * per-channel sum / sum-of-squares partials of the bf16-rounded output (statistics of the following BatchNorm)
//   * backward: data gradient, weight gradient partials, and the reduction terms of the INPUT BatchNorm's backward
//     (sum dv, sum dv*yhat) are produced from one pass over (dY, Y_in).
// These are HBM/FMA-bound stencils: CUDA cores, one warp = 64 channels (2 per lane) of one 2x4 output patch,
// weights in registers, input tile staged in shared memory after the BN+swish transform.
#include "common.cuh"
#include "mclip_internal.h"
#include <cuda_fp16.h>
#include <algorithm>
using std::max;

#define DW_THREADS 128
#define DW_WARPS 4
#define DW_CCH 64          // channels per CTA tile: one warp lane = 2 channels (one packed f32x2)

struct DwDev {
  int N, H, W, C, Ho, Wo;
  int pl, pt;                 // left / top zero padding (right / bottom follow from the output size)
  int tiles_x, tiles_y, n_chunks, slots;
  const bf16* in;             // [N,H,W,C] pre-BN conv output (or a materialised activation when scale == nullptr)
  const float* scale;         // [C] a = gamma*invstd   (nullptr: no transform)
  const float* shift;         // [C] b = beta - mean*a
  int act;                    // 1: swish after the affine
  const float* w;             // [C, K*K] fp32 (reference layout [C,1,K,K])
  bf16* out;                  // [N,Ho,Wo,C]
  float* stats;               // [slots][2][C] or nullptr
  // backward only
  const bf16* dy;             // [N,Ho,Wo,C] gradient w.r.t. this conv's output
  bf16* dx;                   // [N,H,W,C]   gradient w.r.t. the PRE-activation input (dv = dA * swish'(a*y+b)) or w.r.t. x
  float* dw_part;             // [slots][K*K][C]
  float* bn_part;             // [slots][2][C]: sum dv, sum dv*yhat   (nullptr when the input has no BN)
  const float* mean;          // [C] batch mean / invstd of the input BN (for yhat)
  const float* invstd;
};

// ---- packed fp32x2 helpers (FFMA2 on sm_100: one issue slot per two FMAs) ---------------------------------------------------
typedef unsigned long long u64;
__device__ __forceinline__ float2 bf2_to_f2(uint32_t u) { return make_float2(__uint_as_float(u << 16), __uint_as_float(u & 0xffff0000u)); }
__device__ __forceinline__ void ffma2(float2& d, const float2& a, const float2& b) {
  u64& dd = reinterpret_cast<u64&>(d);
  asm("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(dd) : "l"(reinterpret_cast<const u64&>(a)), "l"(reinterpret_cast<const u64&>(b)));
}
__device__ __forceinline__ float2 ffma2r(const float2& a, const float2& b, const float2& c) {
  float2 d;
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(reinterpret_cast<u64&>(d)) : "l"(reinterpret_cast<const u64&>(a)), "l"(reinterpret_cast<const u64&>(b)),
      "l"(reinterpret_cast<const u64&>(c)));
  return d;
}
__device__ __forceinline__ float2 fmul2(const float2& a, const float2& b) {
  float2 d;
  asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(reinterpret_cast<u64&>(d)) : "l"(reinterpret_cast<const u64&>(a)), "l"(reinterpret_cast<const u64&>(b)));
  return d;
}

// ---- tile loader: global NHWC bf16 -> smem [rows][cols][64ch] bf16 with optional affine+swish, zero padded -------------
// swish(t) = t*sigma(t) = h + h*tanh(h) with h = t/2 : the 1/2 is folded into the affine, one MUFU per element.
// `pix` is a per-kernel table (ry << 8 | rx) of the tile's pixels so that no division runs per tile.
__device__ __forceinline__ void dw_make_pixtab(uint16_t* pix, int rows, int cols) {
  for (int p = threadIdx.x; p < rows * cols; p += DW_THREADS) pix[p] = (uint16_t)(((p / cols) << 8) | (p % cols));
}

// In-place BN-affine + swish of a TMA-staged tile [npix][64ch] (bf16) by the DW_THREADS compute threads.
// Pixels outside the image were zero-filled by TMA and must STAY zero (ZeroPad2d acts on the activated tensor).
__device__ __forceinline__ void dw_transform_tile(bf16* __restrict__ tile, const uint16_t* __restrict__ pix, int y0, int x0, int rows, int cols,
                                                  int H, int W, int C, int c0, const float* sc, const float* sh, int act) {
  const int v = threadIdx.x & 7;
  const int c = c0 + v * 8;
  if (c >= C) return;                                    // channels past C: zero-filled, never stored
  const float f = act ? 0.5f : 1.0f;
  float2 a[4], b[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    a[i] = make_float2(f * sc[c + 2 * i], f * sc[c + 2 * i + 1]);
    b[i] = make_float2(f * sh[c + 2 * i], f * sh[c + 2 * i + 1]);
  }
  const int npix = rows * cols;
  const bool interior = y0 >= 0 && x0 >= 0 && y0 + rows <= H && x0 + cols <= W;
  bf16* base = tile + v * 8;
  for (int p = threadIdx.x >> 3; p < npix; p += DW_THREADS / 8) {
    if (!interior) {
      const uint32_t pr = pix[p];
      const int y = y0 + (int)(pr >> 8), x = x0 + (int)(pr & 255u);
      if ((unsigned)y >= (unsigned)H || (unsigned)x >= (unsigned)W) continue;
    }
    bf16x8 o = *reinterpret_cast<const bf16x8*>(base + p * DW_CCH);
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      float2 h = ffma2r(bf2_to_f2(o.w[i]), a[i], b[i]);
      if (act) h = ffma2r(h, make_float2(fast_tanh(h.x), fast_tanh(h.y)), h);
      o.w[i] = pack_bf16(h.x, h.y);
    }
    *reinterpret_cast<bf16x8*>(base + p * DW_CCH) = o;
  }
}

// One warp computes a PRO x 4 patch of outputs for its 64 channels from a staged tile (row stride IW pixels):
// acc[oy][ox] += sum_{ky,kx} tile[(py+oy)*S+ky][(px+ox)*S+kx] * w[ky*K+kx]
template <int K, int S, int PRO>
__device__ __forceinline__ void dw_patch(const bf16* __restrict__ tile, int IW, int ty, int tx, int lane, const float2* __restrict__ w,
                                         float2 (&acc)[PRO][4]) {
  constexpr int PR = (PRO - 1) * S + K, PC = 3 * S + K;
#pragma unroll
  for (int i = 0; i < PRO; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = make_float2(0.f, 0.f);
  const uint32_t* base = reinterpret_cast<const uint32_t*>(tile) + (ty * IW + tx) * (DW_CCH / 2) + lane;
#pragma unroll
  for (int iy = 0; iy < PR; ++iy) {
    float2 r[PC];
    const uint32_t* rowp = base + iy * IW * (DW_CCH / 2);
#pragma unroll
    for (int ix = 0; ix < PC; ++ix) r[ix] = bf2_to_f2(rowp[ix * (DW_CCH / 2)]);
#pragma unroll
    for (int oy = 0; oy < PRO; ++oy) {
      const int ky = iy - oy * S;
      if (ky < 0 || ky >= K) continue;
#pragma unroll
      for (int kx = 0; kx < K; ++kx)
#pragma unroll
        for (int ox = 0; ox < 4; ++ox) ffma2(acc[oy][ox], r[ox * S + kx], w[ky * K + kx]);
    }
  }
}

// =====================================================================================================
// forward
// =====================================================================================================
// The halo tile of the NEXT work item is fetched by TMA (issued by one thread as soon as every warp has released the other
// stage) while the current one is transformed (BN+swish in place) and convolved.  TMA zero-fills the static padding and
// the channel tail, so the kernel has no load address arithmetic at all.
template <int K, int S, int TH, int TW, int DW_STAGES>
__global__ void __launch_bounds__(DW_THREADS, 4) mclip_dwconv_fwd_kernel(const __grid_constant__ CUtensorMap tmIn, const DwDev p) {
  constexpr int IH = (TH - 1) * S + K, IW = (TW - 1) * S + K;
  constexpr int PRO = (S == 1 && K == 3) ? 4 : 2;                 // output rows per warp patch
  constexpr uint32_t TILE_BYTES = IH * IW * DW_CCH * 2;
  extern __shared__ __align__(128) uint8_t smem_dw[];
  uint8_t* sbase = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_dw) + 127) & ~uintptr_t(127));
  uint16_t* pix = reinterpret_cast<uint16_t*>(sbase + DW_STAGES * TILE_BYTES);
  __shared__ float red[DW_WARPS][4][32];
  __shared__ __align__(8) uint64_t full[DW_STAGES], empty[DW_STAGES];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int chunk = blockIdx.x % p.n_chunks, slot = blockIdx.x / p.n_chunks;
  const int c0 = chunk * DW_CCH;
  if (threadIdx.x == 0) {
    tma_prefetch_desc(&tmIn);
    for (int s = 0; s < DW_STAGES; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], DW_WARPS); }
    fence_mbar_init();
  }
  dw_make_pixtab(pix, IH, IW);
  __syncthreads();
  const int tiles_per_img = p.tiles_x * p.tiles_y;
  const int total_tiles = p.N * tiles_per_img;
  const int my_tiles = total_tiles > slot ? (total_tiles - slot + p.slots - 1) / p.slots : 0;
  auto issue = [&](int it) {      // thread 0 only: fetch work item `it` into its stage once all warps released it
    const int st = it % DW_STAGES;
    const uint32_t ph = (uint32_t)(it / DW_STAGES) & 1u;
    const int t = slot + it * p.slots;
    const int n = t / tiles_per_img, tr = t % tiles_per_img;
    const int oy0 = (tr / p.tiles_x) * TH, ox0 = (tr % p.tiles_x) * TW;
    mbar_wait(&empty[st], ph ^ 1);
    mbar_expect_tx(&full[st], TILE_BYTES);
    tma_load_4d(sbase + (size_t)st * TILE_BYTES, &tmIn, &full[st], c0, ox0 * S - p.pl, oy0 * S - p.pt, n);
  };
  if (threadIdx.x == 0)
    for (int i = 0; i < DW_STAGES - 1 && i < my_tiles; ++i) issue(i);
  const int c = c0 + lane * 2;
  const bool cvalid = c < p.C;
  float2 w[K * K];
#pragma unroll
  for (int t = 0; t < K * K; ++t) w[t] = cvalid ? make_float2(p.w[(size_t)c * K * K + t], p.w[(size_t)(c + 1) * K * K + t]) : make_float2(0.f, 0.f);
  float2 s_sum = make_float2(0.f, 0.f), s_sq = make_float2(0.f, 0.f);
  const float2 one = make_float2(1.f, 1.f);
  const int cw = p.C >> 1;                                         // channel pairs per pixel (row stride in 32-bit words)
  int it = 0;
  for (int t = slot; t < total_tiles; t += p.slots, ++it) {
    const int st = it % DW_STAGES;
    const uint32_t ph = (uint32_t)(it / DW_STAGES) & 1u;
    const int n = t / tiles_per_img, tr = t % tiles_per_img;
    const int oy0 = (tr / p.tiles_x) * TH, ox0 = (tr % p.tiles_x) * TW;
    bf16* tile = reinterpret_cast<bf16*>(sbase + (size_t)st * TILE_BYTES);
    uint32_t* outw = reinterpret_cast<uint32_t*>(p.out + (size_t)n * p.Ho * p.Wo * p.C) + (c >> 1);
    if (threadIdx.x == 0 && it + DW_STAGES - 1 < my_tiles) issue(it + DW_STAGES - 1);
    __syncwarp();
    mbar_wait(&full[st], ph);
    if (p.scale) {
      dw_transform_tile(tile, pix, oy0 * S - p.pt, ox0 * S - p.pl, IH, IW, p.H, p.W, p.C, c0, p.scale, p.shift, p.act);
      named_bar_sync(1, DW_THREADS);
    }
    // patch loops stay ROLLED: their trip counts are small constants and nvcc would unroll them into several thousand SASS
    // instructions that every warp walks through once per tile (instruction-fetch stalls, cf. the GEMM epilogue)
#pragma unroll 1
    for (int pa = warp; pa < (TH / PRO) * (TW / 4); pa += DW_WARPS) {
      const int py = (pa / (TW / 4)) * PRO, px = (pa % (TW / 4)) * 4;
      const int y0 = oy0 + py, x0 = ox0 + px;
      if (y0 >= p.Ho || x0 >= p.Wo) continue;                      // warp-uniform
      float2 acc[PRO][4];
      dw_patch<K, S, PRO>(tile, IW, py * S, px * S, lane, w, acc);
      if (!cvalid) continue;
      uint32_t* op = outw + (y0 * p.Wo + x0) * cw;
      const bool inner = y0 + PRO <= p.Ho && x0 + 4 <= p.Wo;       // interior patch: no per-pixel checks
#pragma unroll
      for (int oy = 0; oy < PRO; ++oy)
#pragma unroll
        for (int ox = 0; ox < 4; ++ox) {
          if (inner || (y0 + oy < p.Ho && x0 + ox < p.Wo)) {
            const uint32_t pk = pack_bf16(acc[oy][ox].x, acc[oy][ox].y);
            op[(oy * p.Wo + ox) * cw] = pk;
            const float2 a = bf2_to_f2(pk);
            ffma2(s_sum, a, one);
            ffma2(s_sq, a, a);
          }
        }
    }
    // this warp is done with the stage (generic-proxy writes of the transform must be ordered before the next TMA write)
    fence_proxy_async_smem();
    __syncwarp();
    if (lane == 0) mbar_arrive(&empty[st]);
  }
  if (p.stats) {
    red[warp][0][lane] = s_sum.x; red[warp][1][lane] = s_sum.y; red[warp][2][lane] = s_sq.x; red[warp][3][lane] = s_sq.y;
    named_bar_sync(1, DW_THREADS);
    if (warp == 0 && cvalid) {
      float a = 0.f, b = 0.f, cc = 0.f, d = 0.f;
#pragma unroll
      for (int w2 = 0; w2 < DW_WARPS; ++w2) { a += red[w2][0][lane]; b += red[w2][1][lane]; cc += red[w2][2][lane]; d += red[w2][3][lane]; }
      float* stp = p.stats + (size_t)slot * 2 * p.C;
      stp[c] = a; stp[c + 1] = b; stp[p.C + c] = cc; stp[p.C + c + 1] = d;
    }
  }
}

// =====================================================================================================
// backward: weight gradient, data gradient (+ swish' of the input activation), input-BN reduction terms
// =====================================================================================================
// Tile = TH x TW OUTPUT pixels and the S*TH x S*TW INPUT pixels they own.  Two staged windows:
//   atile : activated input window feeding the owned outputs          (weight gradient)
//   gtile : dY window feeding the owned inputs; it contains the owned outputs' dY as a sub-window
// Per 2x4 patch a warp accumulates dW (registers, for the whole kernel) and computes the data gradient as a
// register-window correlation of gtile with the flipped kernel (stride 1) or per parity class (stride 2).
template <int K, int S, int TH, int TW, int DW_STAGES>
__global__ void __launch_bounds__(DW_THREADS, (K == 3) ? 4 : 3)
mclip_dwconv_bwd_kernel(const __grid_constant__ CUtensorMap tmIn, const __grid_constant__ CUtensorMap tmDy, const DwDev p) {
  constexpr int IH = (TH - 1) * S + K, IW = (TW - 1) * S + K;     // activation window (wgrad)
  constexpr int GH = (S == 1) ? TH + K - 1 : TH + (K + 1) / 2;    // dY window (dgrad)
  constexpr int GW = (S == 1) ? TW + K - 1 : TW + (K + 1) / 2;
  constexpr uint32_t A_BYTES = IH * IW * DW_CCH * 2, G_BYTES = GH * GW * DW_CCH * 2, STAGE_BYTES = A_BYTES + G_BYTES;
  extern __shared__ __align__(128) uint8_t smem_dw[];
  uint8_t* sbase = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_dw) + 127) & ~uintptr_t(127));
  float* wsm = reinterpret_cast<float*>(sbase + DW_STAGES * STAGE_BYTES);          // [K*K][64] weights
  uint16_t* apix = reinterpret_cast<uint16_t*>(wsm + K * K * DW_CCH);
  __shared__ float red[DW_WARPS][4][32];
  __shared__ __align__(8) uint64_t full[DW_STAGES], empty[DW_STAGES];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int chunk = blockIdx.x % p.n_chunks, slot = blockIdx.x / p.n_chunks;
  const int c0 = chunk * DW_CCH, c = c0 + lane * 2;
  const bool cvalid = c < p.C;
  if (threadIdx.x == 0) {
    tma_prefetch_desc(&tmIn); tma_prefetch_desc(&tmDy);
    for (int s = 0; s < DW_STAGES; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], DW_WARPS); }
    fence_mbar_init();
  }
  for (int i = threadIdx.x; i < K * K * DW_CCH; i += DW_THREADS) {
    const int t = i / DW_CCH, ch = i % DW_CCH;
    wsm[i] = (c0 + ch < p.C) ? p.w[(size_t)(c0 + ch) * K * K + t] : 0.f;
  }
  dw_make_pixtab(apix, IH, IW);
  __syncthreads();
  const int tiles_per_img = p.tiles_x * p.tiles_y;
  const int total_tiles = p.N * tiles_per_img;
  const int my_tiles = total_tiles > slot ? (total_tiles - slot + p.slots - 1) / p.slots : 0;
  auto tile_origin = [&](int t, int& n, int& oy0, int& ox0, int& gy0, int& gx0) {
    n = t / tiles_per_img;
    const int tr = t % tiles_per_img;
    oy0 = (tr / p.tiles_x) * TH; ox0 = (tr % p.tiles_x) * TW;
    const int gnum_y = oy0 * S + p.pt - (K - 1), gnum_x = ox0 * S + p.pl - (K - 1);
    gy0 = (S == 1) ? gnum_y : (gnum_y >= 0 ? (gnum_y + 1) / 2 : -((-gnum_y) / 2));   // ceil(gnum / S)
    gx0 = (S == 1) ? gnum_x : (gnum_x >= 0 ? (gnum_x + 1) / 2 : -((-gnum_x) / 2));
  };
  auto issue = [&](int it) {      // thread 0 only
    const int st = it % DW_STAGES;
    const uint32_t ph = (uint32_t)(it / DW_STAGES) & 1u;
    int n, oy0, ox0, gy0, gx0;
    tile_origin(slot + it * p.slots, n, oy0, ox0, gy0, gx0);
    mbar_wait(&empty[st], ph ^ 1);
    mbar_expect_tx(&full[st], STAGE_BYTES);
    uint8_t* dst = sbase + (size_t)st * STAGE_BYTES;
    tma_load_4d(dst, &tmIn, &full[st], c0, ox0 * S - p.pl, oy0 * S - p.pt, n);
    tma_load_4d(dst + A_BYTES, &tmDy, &full[st], c0, gx0, gy0, n);
  };
  if (threadIdx.x == 0)
    for (int i = 0; i < (DW_STAGES > 1 ? DW_STAGES - 1 : 1) && i < my_tiles; ++i) issue(i);
  float2 ab = make_float2(1.f, 1.f), bb = make_float2(0.f, 0.f), mu = make_float2(0.f, 0.f), is = make_float2(1.f, 1.f);
  if (p.scale && cvalid) { ab = make_float2(p.scale[c], p.scale[c + 1]); bb = make_float2(p.shift[c], p.shift[c + 1]); }
  if (p.bn_part && cvalid) { mu = make_float2(p.mean[c], p.mean[c + 1]); is = make_float2(p.invstd[c], p.invstd[c + 1]); }
  float2 bs = make_float2(0.f, 0.f), bq = make_float2(0.f, 0.f);
  const float2 abh = make_float2(0.5f * ab.x, 0.5f * ab.y), bbh = make_float2(0.5f * bb.x, 0.5f * bb.y);
  const float2 nmis = make_float2(-mu.x * is.x, -mu.y * is.y);
  const float2 half2 = make_float2(0.5f, 0.5f), one2 = make_float2(1.f, 1.f), two2 = make_float2(2.f, 2.f), neg1 = make_float2(-1.f, -1.f);
  float2 dw[K * K];
#pragma unroll
  for (int q = 0; q < K * K; ++q) dw[q] = make_float2(0.f, 0.f);
  float2 wf[(S == 1) ? K * K : 1];                                 // flipped kernel (stride-1 data gradient)
  if constexpr (S == 1) {
#pragma unroll
    for (int q = 0; q < K * K; ++q) wf[q] = *reinterpret_cast<const float2*>(wsm + (K * K - 1 - q) * DW_CCH + lane * 2);
  }
  const int cw = p.C >> 1;
  int it = 0;
  for (int t = slot; t < total_tiles; t += p.slots, ++it) {
    const int st = it % DW_STAGES;
    const uint32_t ph = (uint32_t)(it / DW_STAGES) & 1u;
    int n, oy0, ox0, gy0, gx0;
    tile_origin(t, n, oy0, ox0, gy0, gx0);
    const int iy0 = oy0 * S, ix0 = ox0 * S;                        // first owned input pixel
    bf16* atile = reinterpret_cast<bf16*>(sbase + (size_t)st * STAGE_BYTES);
    bf16* gtile = reinterpret_cast<bf16*>(sbase + (size_t)st * STAGE_BYTES + A_BYTES);
    const uint32_t* inw = reinterpret_cast<const uint32_t*>(p.in + (size_t)n * p.H * p.W * p.C) + (c >> 1);
    uint32_t* dxw = reinterpret_cast<uint32_t*>(p.dx + (size_t)n * p.H * p.W * p.C) + (c >> 1);
    if (DW_STAGES > 1 && threadIdx.x == 0 && it + DW_STAGES - 1 < my_tiles) issue(it + DW_STAGES - 1);
    __syncwarp();
    mbar_wait(&full[st], ph);
    if (p.scale) {
      dw_transform_tile(atile, apix, iy0 - p.pt, ix0 - p.pl, IH, IW, p.H, p.W, p.C, c0, p.scale, p.shift, p.act);
      named_bar_sync(1, DW_THREADS);
    }

    // dv = dA*swish'(a*y+b) (y = pre-BN input), BN-backward partials, store
    auto finish = [&](uint32_t yu, int off, float2 d) {
      if (p.scale) {
        const float2 yv = bf2_to_f2(yu);
        if (p.act) {
          // swish'(v) = s (1 + v (1 - s)), s = sigma(v) = 0.5 + 0.5 tanh(v/2); packed math, one MUFU per element
          const float2 hv = ffma2r(yv, abh, bbh);                                   // v / 2
          const float2 sg = ffma2r(make_float2(fast_tanh(hv.x), fast_tanh(hv.y)), half2, half2);
          const float2 om = ffma2r(sg, neg1, one2);                                 // 1 - s
          const float2 q = ffma2r(fmul2(hv, om), two2, one2);                       // 1 + v (1 - s)
          d = fmul2(d, fmul2(sg, q));
        }
        const uint32_t pk = pack_bf16(d.x, d.y);
        dxw[off] = pk;
        d = bf2_to_f2(pk);
        ffma2(bs, d, one2);
        ffma2(bq, d, ffma2r(yv, is, nmis));                                         // yhat = (y - mean) * invstd
      } else {
        dxw[off] = pack_bf16(d.x, d.y);
      }
    };

    // ---- weight gradient over the owned outputs: dW[ky,kx] += dY[oy,ox] * A[oy*S+ky, ox*S+kx] ----
#pragma unroll 1
    for (int pa = warp; pa < (TH / 2) * (TW / 4); pa += DW_WARPS) {
      const int py = (pa / (TW / 4)) * 2, px = (pa % (TW / 4)) * 4;
      if (oy0 + py >= p.Ho || ox0 + px >= p.Wo) continue;
      constexpr int PR = S + K, PC = 3 * S + K;
      float2 g[2][4];
      {
        // owned outputs sit in gtile at (oy - gy0, ox - gx0); outputs past Ho/Wo were zero-filled by the loader
        const uint32_t* gp = reinterpret_cast<const uint32_t*>(gtile) + ((oy0 + py - gy0) * GW + (ox0 + px - gx0)) * (DW_CCH / 2) + lane;
#pragma unroll
        for (int oy = 0; oy < 2; ++oy)
#pragma unroll
          for (int ox = 0; ox < 4; ++ox) g[oy][ox] = bf2_to_f2(gp[(oy * GW + ox) * (DW_CCH / 2)]);
      }
      const uint32_t* abase = reinterpret_cast<const uint32_t*>(atile) + (py * S * IW + px * S) * (DW_CCH / 2) + lane;
#pragma unroll
      for (int iy = 0; iy < PR; ++iy) {
        float2 r[PC];
#pragma unroll
        for (int ix = 0; ix < PC; ++ix) r[ix] = bf2_to_f2(abase[(iy * IW + ix) * (DW_CCH / 2)]);
#pragma unroll
        for (int oy = 0; oy < 2; ++oy) {
          const int ky = iy - oy * S;
          if (ky < 0 || ky >= K) continue;
#pragma unroll
          for (int kx = 0; kx < K; ++kx)
#pragma unroll
            for (int ox = 0; ox < 4; ++ox) ffma2(dw[ky * K + kx], g[oy][ox], r[ox * S + kx]);
        }
      }
    }

    // ---- data gradient over the owned inputs ----
    if constexpr (S == 1) {
#pragma unroll 1
      for (int pa = warp; pa < (TH / 2) * (TW / 4); pa += DW_WARPS) {
        const int ly = (pa / (TW / 4)) * 2, lx = (pa % (TW / 4)) * 4;
        const int y0 = iy0 + ly, x0 = ix0 + lx;
        if (y0 >= p.H || x0 >= p.W || !cvalid) continue;
        const int off0 = (y0 * p.W + x0) * cw;
        uint32_t yu[2][4];
#pragma unroll
        for (int oy = 0; oy < 2; ++oy)
#pragma unroll
          for (int ox = 0; ox < 4; ++ox)
            yu[oy][ox] = (p.scale && y0 + oy < p.H && x0 + ox < p.W) ? __ldg(inw + off0 + (oy * p.W + ox) * cw) : 0u;
        float2 acc[2][4];
        dw_patch<K, 1, 2>(gtile, GW, ly, lx, lane, wf, acc);
#pragma unroll
        for (int oy = 0; oy < 2; ++oy)
#pragma unroll
          for (int ox = 0; ox < 4; ++ox)
            if (y0 + oy < p.H && x0 + ox < p.W) finish(yu[oy][ox], off0 + (oy * p.W + ox) * cw, acc[oy][ox]);
      }
    } else {
      // stride 2: input pixel (y,x) receives taps ky = cy + 2m, kx = cx + 2n with cy = (y+pt)&1, cx = (x+pl)&1.
      // work item = one input row, one x parity, 4 same-parity pixels (consecutive dY columns).
      constexpr int NT = (K + 1) / 2;                                // taps per axis per parity class (max)
      constexpr int OWN_H = TH * 2, OWN_W = TW * 2;
#pragma unroll 1
      for (int it = warp; it < OWN_H * 2 * (OWN_W / 8); it += DW_WARPS) {
        const int ly = it / (2 * (OWN_W / 8)), rem = it % (2 * (OWN_W / 8));
        const int par = rem / (OWN_W / 8), strip = rem % (OWN_W / 8);
        const int y = iy0 + ly, xb = ix0 + strip * 8 + par;          // pixels xb, xb+2, xb+4, xb+6
        if (y >= p.H || xb >= p.W || !cvalid) continue;
        const int off0 = (y * p.W + xb) * cw;
        uint32_t yu[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) yu[j] = (p.scale && xb + 2 * j < p.W) ? __ldg(inw + off0 + 2 * j * cw) : 0u;
        const int cy = (y + p.pt) & 1, cx = (xb + p.pl) & 1;
        const int obx = ((xb + p.pl - cx) >> 1) - gx0;               // dY column (tile coords) of pixel j=0 for n=0
        float2 acc[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[j] = make_float2(0.f, 0.f);
#pragma unroll
        for (int m = 0; m < NT; ++m) {
          const int ky = cy + 2 * m;
          if (ky >= K) continue;
          const int gy = ((y + p.pt - ky) >> 1) - gy0;               // numerator is even by construction
          if (gy < 0 || gy >= GH) continue;
          float2 r[NT + 3];
#pragma unroll
          for (int q = 0; q < NT + 3; ++q) {
            const int gx = obx - (NT - 1) + q;
            uint32_t u = 0u;
            if (gx >= 0 && gx < GW) u = reinterpret_cast<const uint32_t*>(gtile)[(gy * GW + gx) * (DW_CCH / 2) + lane];
            r[q] = bf2_to_f2(u);
          }
#pragma unroll
          for (int nn = 0; nn < NT; ++nn) {
            const int kx = cx + 2 * nn;
            if (kx >= K) continue;
            const float2 wv = *reinterpret_cast<const float2*>(wsm + (ky * K + kx) * DW_CCH + lane * 2);
#pragma unroll
            for (int j = 0; j < 4; ++j) ffma2(acc[j], r[j - nn + NT - 1], wv);
          }
        }
#pragma unroll
        for (int j = 0; j < 4; ++j)
          if (xb + 2 * j < p.W) finish(yu[j], off0 + 2 * j * cw, acc[j]);
      }
    }
    fence_proxy_async_smem();
    __syncwarp();
    if (lane == 0) mbar_arrive(&empty[st]);
    if (DW_STAGES == 1 && threadIdx.x == 0 && it + 1 < my_tiles) issue(it + 1);
  }
  // ---- flush partials (all TMA loads have been consumed: the stage buffers are free) ----
  named_bar_sync(1, DW_THREADS);
  float* wred = reinterpret_cast<float*>(sbase);                   // reuse the tiles: [warps][K*K][64]
#pragma unroll
  for (int q = 0; q < K * K; ++q) *reinterpret_cast<float2*>(wred + ((size_t)warp * K * K + q) * DW_CCH + lane * 2) = dw[q];
  red[warp][0][lane] = bs.x; red[warp][1][lane] = bs.y; red[warp][2][lane] = bq.x; red[warp][3][lane] = bq.y;
  named_bar_sync(1, DW_THREADS);
  for (int i = threadIdx.x; i < K * K * DW_CCH; i += DW_THREADS) {
    const int t = i / DW_CCH, ch = i % DW_CCH;
    if (c0 + ch < p.C) {
      float s2 = 0.f;
#pragma unroll
      for (int w2 = 0; w2 < DW_WARPS; ++w2) s2 += wred[((size_t)w2 * K * K + t) * DW_CCH + ch];
      p.dw_part[((size_t)slot * K * K + t) * p.C + c0 + ch] = s2;
    }
  }
  if (p.bn_part && warp == 0 && cvalid) {
    float a = 0.f, b = 0.f, cc = 0.f, d = 0.f;
#pragma unroll
    for (int w2 = 0; w2 < DW_WARPS; ++w2) { a += red[w2][0][lane]; b += red[w2][1][lane]; cc += red[w2][2][lane]; d += red[w2][3][lane]; }
    float* st = p.bn_part + (size_t)slot * 2 * p.C;
    st[c] = a; st[c + 1] = b; st[p.C + c] = cc; st[p.C + c + 1] = d;
  }
}

// dW[c, t] = sum_slots part[slot][t][c]      (fixed order)
__global__ void mclip_dw_wgrad_reduce_kernel(const float* __restrict__ part, float* __restrict__ dw, int slots, int KK, int C, int accumulate) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= KK * C) return;
  const int t = i / C, c = i % C;
  float s = 0.f;
  for (int k = 0; k < slots; ++k) s += part[((size_t)k * KK + t) * C + c];
  float* o = dw + (size_t)c * KK + t;
  *o = accumulate ? *o + s : s;
}

// ------------------------------------------------------------------------------------------------
// host: depthwise
// ------------------------------------------------------------------------------------------------
template <int K, int S>
struct DwCfg {
  // forward tile (outputs).  Stride 2 halves the tile height: an 8x16 tile needs a 17x33 (k3) / 19x35 (k5) input window =
  // 72 / 85 KB per stage, i.e. ONE or two 4-warp CTAs per SM; 4x16 keeps the window at 38 / 49 KB and 4 CTAs resident.
  static constexpr int TH = (S == 1) ? 8 : 4;
  static constexpr int TW = 16;
  static constexpr int BTH = (S == 1) ? 8 : 4;      // backward tile height
  static constexpr int FST = (K == 3 && S == 1) ? 2 : 1;   // TMA stages, forward (k5 / stride 2: one stage keeps 4 CTAs per SM)
  static constexpr int BST = 1;                     // backward stages two tensors per tile: one stage, occupancy hides the TMA latency
};

static int dw_slots(int n_chunks, int total_tiles, int per_sm) {
  int s = (mclip_num_sms() * per_sm) / n_chunks;
  if (s < 1) s = 1;
  if (s > total_tiles) s = total_tiles;
  return s;
}

// 4-D tensor map over an NHWC bf16 tensor: dims {C, W, H, N}, box {64 channels, box_w, box_h, 1}, no swizzle.
static int dw_tmap(CUtensorMap* m, const void* ptr, int N, int H, int W, int C, int box_w, int box_h) {
  const unsigned long long dims[4] = {(unsigned long long)C, (unsigned long long)W, (unsigned long long)H, (unsigned long long)N};
  const unsigned long long strides[3] = {(unsigned long long)C * 2, (unsigned long long)W * C * 2, (unsigned long long)H * W * C * 2};
  const unsigned box[4] = {DW_CCH, (unsigned)box_w, (unsigned)box_h, 1};
  return mclip_tmap_encode_bf16(m, ptr, 4, dims, strides, box, 0);
}

template <int K, int S>
static int dw_launch_fwd(DwDev& p, cudaStream_t stream, int slots_given) {
  constexpr int TH = DwCfg<K, S>::TH, TW = DwCfg<K, S>::TW;
  constexpr int IH = (TH - 1) * S + K, IW = (TW - 1) * S + K;
  const int smem = DwCfg<K, S>::FST * IH * IW * DW_CCH * 2 + IH * IW * 2 + 256;
  p.tiles_x = ceil_div(p.Wo, TW); p.tiles_y = ceil_div(p.Ho, TH);
  p.n_chunks = ceil_div(p.C, DW_CCH);
  p.slots = slots_given;
  CUtensorMap tm;
  int rc = dw_tmap(&tm, p.in, p.N, p.H, p.W, p.C, IW, IH);
  if (rc) return rc;
  auto kern = mclip_dwconv_fwd_kernel<K, S, TH, TW, DwCfg<K, S>::FST>;
  static bool attr = false;
  if (!attr) { MCLIP_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem)); attr = true; }
  kern<<<p.n_chunks * p.slots, DW_THREADS, smem, stream>>>(tm, p);
  MCLIP_CHECK_LAUNCH();
  return MCLIP_OK;
}

template <int K, int S>
static int dw_launch_bwd(DwDev& p, cudaStream_t stream, int slots_given) {
  constexpr int TH = DwCfg<K, S>::BTH, TW = DwCfg<K, S>::TW;
  constexpr int IH = (TH - 1) * S + K, IW = (TW - 1) * S + K;
  constexpr int GH = (S == 1) ? TH + K - 1 : TH + (K + 1) / 2, GW = (S == 1) ? TW + K - 1 : TW + (K + 1) / 2;
  int smem = DwCfg<K, S>::BST * (IH * IW + GH * GW) * DW_CCH * 2 + K * K * DW_CCH * 4 + IH * IW * 2 + 256;
  const int red_bytes = DW_WARPS * K * K * DW_CCH * 4 + 256;          // the final dW reduction reuses the stage buffers
  if (smem < red_bytes) smem = red_bytes;
  // the tile grid must cover every OUTPUT pixel (weight gradient) and every INPUT pixel (data gradient): with the
  // reference's static pads, S*Ho can be smaller than H (e.g. H=33, k3 s2 pads (0,1) -> Ho=16 but input row 32 is read)
  p.tiles_x = ceil_div(max(p.Wo, ceil_div(p.W, S)), TW); p.tiles_y = ceil_div(max(p.Ho, ceil_div(p.H, S)), TH);
  p.n_chunks = ceil_div(p.C, DW_CCH);
  p.slots = slots_given;
  CUtensorMap tmIn, tmDy;
  int rc = dw_tmap(&tmIn, p.in, p.N, p.H, p.W, p.C, IW, IH);
  if (rc) return rc;
  if ((rc = dw_tmap(&tmDy, p.dy, p.N, p.Ho, p.Wo, p.C, GW, GH))) return rc;
  auto kern = mclip_dwconv_bwd_kernel<K, S, TH, TW, DwCfg<K, S>::BST>;
  static bool attr = false;
  if (!attr) { MCLIP_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem)); attr = true; }
  kern<<<p.n_chunks * p.slots, DW_THREADS, smem, stream>>>(tmIn, tmDy, p);
  MCLIP_CHECK_LAUNCH();
  return MCLIP_OK;
}

static int dw_tiles(const mclip_dwconv_args* a, bool bwd) {
  const int TH = a->stride == 1 ? 8 : 4, TW = 16;
  if (!bwd) return a->n * ceil_div(a->ho, TH) * ceil_div(a->wo, TW);
  const int hy = max(a->ho, ceil_div(a->h, a->stride)), wx = max(a->wo, ceil_div(a->w, a->stride));
  return a->n * ceil_div(hy, TH) * ceil_div(wx, TW);
}

template <int K, int S>
static int dw_blocks_per_sm(bool bwd) {
  static int cache[2] = {0, 0};
  if (cache[bwd]) return cache[bwd];
  int n = 0, smem;
  if (!bwd) {
    constexpr int TH = DwCfg<K, S>::TH, TW = DwCfg<K, S>::TW, IH = (TH - 1) * S + K, IW = (TW - 1) * S + K;
    smem = DwCfg<K, S>::FST * IH * IW * DW_CCH * 2 + IH * IW * 2 + 256;
    auto kern = mclip_dwconv_fwd_kernel<K, S, TH, TW, DwCfg<K, S>::FST>;
    cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, kern, DW_THREADS, smem);
  } else {
    constexpr int TH = DwCfg<K, S>::BTH, TW = DwCfg<K, S>::TW, IH = (TH - 1) * S + K, IW = (TW - 1) * S + K;
    constexpr int GH = (S == 1) ? TH + K - 1 : TH + (K + 1) / 2, GW = (S == 1) ? TW + K - 1 : TW + (K + 1) / 2;
    smem = DwCfg<K, S>::BST * (IH * IW + GH * GW) * DW_CCH * 2 + K * K * DW_CCH * 4 + IH * IW * 2 + 256;
    auto kern = mclip_dwconv_bwd_kernel<K, S, TH, TW, DwCfg<K, S>::BST>;
    cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, kern, DW_THREADS, smem);
  }
  if (n < 1) n = 1;
  cache[bwd] = n;
  return n;
}

extern "C" int mclip_dwconv_slots(const mclip_dwconv_args* a, int backward) {
  if (!a || a->c <= 0) return -1;
  if (mclip_dws_covers(a, backward)) return mclip_dws_slots(a, backward);
  const bool b = backward != 0;
  int per_sm;
  if (a->k == 3 && a->stride == 1) per_sm = dw_blocks_per_sm<3, 1>(b);
  else if (a->k == 3) per_sm = dw_blocks_per_sm<3, 2>(b);
  else if (a->stride == 1) per_sm = dw_blocks_per_sm<5, 1>(b);
  else per_sm = dw_blocks_per_sm<5, 2>(b);
  return dw_slots(ceil_div(a->c, DW_CCH), dw_tiles(a, b), per_sm);
}

static int dw_fill(const mclip_dwconv_args* a, DwDev& p) {
  MCLIP_REQUIRE(a && a->in && a->weight, "mclip_dwconv: null operand");
  MCLIP_REQUIRE((a->k == 3 || a->k == 5) && (a->stride == 1 || a->stride == 2), "mclip_dwconv: k=%d stride=%d unsupported", a->k, a->stride);
  MCLIP_REQUIRE(a->c % 8 == 0, "mclip_dwconv: C=%d must be a multiple of 8", a->c);
  MCLIP_REQUIRE(a->ho == (a->h + a->pad_top + a->pad_bottom - a->k) / a->stride + 1 && a->wo == (a->w + a->pad_left + a->pad_right - a->k) / a->stride + 1,
                "mclip_dwconv: output size %dx%d inconsistent with input %dx%d, k=%d s=%d pads (%d,%d,%d,%d)", a->ho, a->wo, a->h, a->w, a->k,
                a->stride, a->pad_left, a->pad_right, a->pad_top, a->pad_bottom);
  memset(&p, 0, sizeof(p));
  p.N = a->n; p.H = a->h; p.W = a->w; p.C = a->c; p.Ho = a->ho; p.Wo = a->wo; p.pl = a->pad_left; p.pt = a->pad_top;
  p.in = (const bf16*)a->in; p.scale = a->in_scale; p.shift = a->in_shift; p.act = a->in_act; p.w = a->weight;
  return MCLIP_OK;
}

extern "C" int mclip_dwconv_forward(const mclip_dwconv_args* a, void* stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  DwDev p;
  int rc = dw_fill(a, p);
  if (rc) return rc;
  MCLIP_REQUIRE(a->out, "mclip_dwconv_forward: null output");
  if (mclip_dws_covers(a, 0)) return mclip_dws_forward(a, stream_);
  p.out = (bf16*)a->out; p.stats = a->stats;
  const int slots = mclip_dwconv_slots(a, 0);
  if (a->stats) MCLIP_REQUIRE(a->stat_slots == slots, "mclip_dwconv_forward: stat_slots=%d, expected %d", a->stat_slots, slots);
  if (a->k == 3 && a->stride == 1) return dw_launch_fwd<3, 1>(p, stream, slots);
  if (a->k == 3 && a->stride == 2) return dw_launch_fwd<3, 2>(p, stream, slots);
  if (a->k == 5 && a->stride == 1) return dw_launch_fwd<5, 1>(p, stream, slots);
  return dw_launch_fwd<5, 2>(p, stream, slots);
}

extern "C" int mclip_dwconv_backward(const mclip_dwconv_args* a, void* stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  DwDev p;
  int rc = dw_fill(a, p);
  if (rc) return rc;
  MCLIP_REQUIRE(a->dy && a->dx && a->dweight && a->dw_partials, "mclip_dwconv_backward: null operand");
  if (mclip_dws_covers(a, 1)) return mclip_dws_backward(a, stream_);
  const int slots = mclip_dwconv_slots(a, 1);
  MCLIP_REQUIRE(a->stat_slots == slots, "mclip_dwconv_backward: stat_slots=%d, expected %d", a->stat_slots, slots);
  p.dy = (const bf16*)a->dy; p.dx = (bf16*)a->dx; p.dw_part = a->dw_partials;
  p.bn_part = a->in_scale ? a->bn_partials : nullptr; p.mean = a->in_mean; p.invstd = a->in_invstd;
  if (p.bn_part) MCLIP_REQUIRE(p.mean && p.invstd, "mclip_dwconv_backward: input BN statistics missing");
  if (a->k == 3 && a->stride == 1) rc = dw_launch_bwd<3, 1>(p, stream, slots);
  else if (a->k == 3 && a->stride == 2) rc = dw_launch_bwd<3, 2>(p, stream, slots);
  else if (a->k == 5 && a->stride == 1) rc = dw_launch_bwd<5, 1>(p, stream, slots);
  else rc = dw_launch_bwd<5, 2>(p, stream, slots);
  if (rc) return rc;
  const int KK = a->k * a->k;
  mclip_dw_wgrad_reduce_kernel<<<ceil_div(KK * a->c, 256), 256, 0, stream>>>(a->dw_partials, a->dweight, slots, KK, a->c, a->accumulate);
  MCLIP_CHECK_LAUNCH();
  return MCLIP_OK;
}

// =====================================================================================================
// stem: dense 3x3 stride-2 convolution, 3 input channels (EfficientNet._conv_stem, efficientnet_custom.py:174-176,273)
// as im2col + tcgen05 GEMM: this kernel gathers every output pixel's 27-tap patch (static pads zero-filled) from the fp32
// image (arbitrary element strides: the trainer hands NCHW-shaped NHWC memory, trainer_ddp.py:288-291) into a bf16 row of
// 32 (K padded for the MMA), t = ci*9 + ky*3 + kx like the OIHW weight.  Forward = mclip_gemm_tn(patches, W[c,32]) with
// the BN-statistics epilogue; weight gradient = mclip_gemm_wgrad(dY, patches).  (bf16 patches = what autocast feeds the conv.)
// =====================================================================================================
template <typename T> __device__ __forceinline__ float stem_ld(const T* p);
template <> __device__ __forceinline__ float stem_ld<float>(const float* p) { return __ldg(p); }
template <> __device__ __forceinline__ float stem_ld<__half>(const __half* p) { return __half2float(__ldg(p)); }
template <> __device__ __forceinline__ float stem_ld<bf16>(const bf16* p) { return __bfloat162float(__ldg(p)); }
template <> __device__ __forceinline__ float stem_ld<uint8_t>(const uint8_t* p) { return (float)__ldg(p); }

// CIN == 3: the trainer's [N,3,H,W] tensor; CIN == 1: one channel, replicated into the three tap groups (bit-identical to
// three identical channels).  NORM (uint8 input): ((u - min) / range - mean) / std per image, the reference's op order.
template <typename T, int CIN, bool NORM>
__global__ void __launch_bounds__(256) mclip_stem_im2col_kernel(const T* __restrict__ in, long long sn, long long sc, long long sh, long long sw,
                                                                bf16* __restrict__ out, int N, int H, int W, int Ho, int Wo, int pl, int pt,
                                                                const bf16* __restrict__ lut) {
  const long long npix = (long long)N * Ho * Wo;
  for (long long px = (long long)blockIdx.x * blockDim.x + threadIdx.x; px < npix; px += (long long)gridDim.x * blockDim.x) {
    const int ox = (int)(px % Wo);
    const int oy = (int)((px / Wo) % Ho);
    const int n = (int)(px / ((long long)Wo * Ho));
    float v[32];
#pragma unroll
    for (int t = 0; t < 32; ++t) v[t] = 0.f;
    const T* base = in + n * sn;
#pragma unroll
    for (int ky = 0; ky < 3; ++ky) {
      const int y = oy * 2 - pt + ky;
      if ((unsigned)y >= (unsigned)H) continue;
#pragma unroll
      for (int kx = 0; kx < 3; ++kx) {
        const int x = ox * 2 - pl + kx;
        if ((unsigned)x >= (unsigned)W) continue;
        const T* ip = base + y * sh + x * sw;
        if (CIN == 3) {
#pragma unroll
          for (int ci = 0; ci < 3; ++ci) v[ci * 9 + ky * 3 + kx] = stem_ld<T>(ip + ci * sc);
        } else {
          float t;
          if (NORM) t = __bfloat162float(lut[n * 256 + (int)__ldg(reinterpret_cast<const uint8_t*>(ip))]);      // 512-byte table per image: L1 hits
          else t = stem_ld<T>(ip);
          v[ky * 3 + kx] = t; v[9 + ky * 3 + kx] = t; v[18 + ky * 3 + kx] = t;
        }
      }
    }
    bf16* op = out + (size_t)px * 32;
#pragma unroll
    for (int g = 0; g < 4; ++g) stg_bf16x8(op + g * 8, pack8(v + g * 8));
  }
}

// per-image min / (max - min) of a uint8 image: one CTA per image, 16-byte loads
__global__ void __launch_bounds__(1024) mclip_image_minmax_u8_kernel(const uint8_t* __restrict__ in, long long hw, float* __restrict__ out,
                                                                     bf16* __restrict__ lut, float nmean, float nstd) {
  __shared__ int smin[32], smax[32];
  __shared__ float s_mn, s_rng;
  const uint8_t* p = in + (size_t)blockIdx.x * hw;
  int mn = 255, mx = 0;
  const long long nvec = (((uintptr_t)p & 15) == 0) ? hw / 16 : 0;
  for (long long i = threadIdx.x; i < nvec; i += blockDim.x) {
    const uint4 q = __ldg(reinterpret_cast<const uint4*>(p) + i);
    const uint32_t w4[4] = {q.x, q.y, q.z, q.w};
#pragma unroll
    for (int k = 0; k < 4; ++k)
#pragma unroll
      for (int b = 0; b < 4; ++b) { const int u = (w4[k] >> (8 * b)) & 255; mn = min(mn, u); mx = max(mx, u); }
  }
  for (long long i = nvec * 16 + threadIdx.x; i < hw; i += blockDim.x) { const int u = p[i]; mn = min(mn, u); mx = max(mx, u); }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) { mn = min(mn, __shfl_xor_sync(0xffffffffu, mn, o)); mx = max(mx, __shfl_xor_sync(0xffffffffu, mx, o)); }
  if ((threadIdx.x & 31) == 0) { smin[threadIdx.x >> 5] = mn; smax[threadIdx.x >> 5] = mx; }
  __syncthreads();
  if (threadIdx.x == 0) {
    for (int w = 0; w < (int)(blockDim.x >> 5); ++w) { mn = min(mn, smin[w]); mx = max(mx, smax[w]); }
    out[2 * blockIdx.x] = (float)mn;
    out[2 * blockIdx.x + 1] = (float)(mx - mn);       // image -= image.min(); image /= image.max()   (imagetext.py:130-131)
    s_mn = (float)mn; s_rng = (float)(mx - mn);
  }
  __syncthreads();
  if (lut && threadIdx.x < 256)                       // ((u - min) / range - mean) / std in fp32 (IEEE division), then the patch's bf16
    lut[blockIdx.x * 256 + threadIdx.x] = __float2bfloat16_rn(__fdiv_rn(__fsub_rn(__fdiv_rn(__fsub_rn((float)threadIdx.x, s_mn), s_rng), nmean), nstd));
}

extern "C" int mclip_image_norm_lut_u8(const void* in, int n, long long hw, float mean, float std, float* minmax, void* lut, void* stream) {
  MCLIP_REQUIRE(in && minmax && n > 0 && hw > 0 && (!lut || std != 0.f), "mclip_image_norm_lut_u8: bad arguments");
  mclip_image_minmax_u8_kernel<<<n, 1024, 0, (cudaStream_t)stream>>>((const uint8_t*)in, hw, minmax, (bf16*)lut, mean, std);
  MCLIP_CHECK_LAUNCH();
  return MCLIP_OK;
}

extern "C" int mclip_stem_im2col(const mclip_stem_args* a, void* stream_) {
  MCLIP_REQUIRE(a && a->in && a->out, "mclip_stem_im2col: null operand");
  MCLIP_REQUIRE(a->ho == (a->h + a->pad_top + a->pad_bottom - 3) / 2 + 1 && a->wo == (a->w + a->pad_left + a->pad_right - 3) / 2 + 1,
                "mclip_stem_im2col: output size inconsistent with the static padding");
  MCLIP_REQUIRE(a->in_channels == 0 || a->in_channels == 1 || a->in_channels == 3, "mclip_stem_im2col: in_channels=%d", a->in_channels);
  MCLIP_REQUIRE(a->in_dtype >= 0 && a->in_dtype <= 3, "mclip_stem_im2col: in_dtype=%d", a->in_dtype);
  const int cin = a->in_channels == 1 ? 1 : 3;
  MCLIP_REQUIRE(cin == 1 || a->in_dtype == 0, "mclip_stem_im2col: the 3-channel input is fp32 (the reference trainer's tensor)");
  MCLIP_REQUIRE(!a->norm_lut || (cin == 1 && a->in_dtype == 3), "mclip_stem_im2col: normalisation on load needs a 1-channel uint8 input");
  const long long npix = (long long)a->n * a->ho * a->wo;
  long long grid = (npix + 255) / 256;
  if (grid > (long long)mclip_num_sms() * 16) grid = (long long)mclip_num_sms() * 16;
  cudaStream_t st = (cudaStream_t)stream_;
#define STEM_LAUNCH(T, CIN, NORM)                                                                                                              \
  mclip_stem_im2col_kernel<T, CIN, NORM><<<(int)grid, 256, 0, st>>>((const T*)a->in, a->stride_n, a->stride_c, a->stride_h, a->stride_w, (bf16*)a->out, \
                                                                    a->n, a->h, a->w, a->ho, a->wo, a->pad_left, a->pad_top, (const bf16*)a->norm_lut)
  if (cin == 3) STEM_LAUNCH(float, 3, false);
  else if (a->in_dtype == 0) STEM_LAUNCH(float, 1, false);
  else if (a->in_dtype == 1) STEM_LAUNCH(__half, 1, false);
  else if (a->in_dtype == 2) STEM_LAUNCH(bf16, 1, false);
  else if (a->norm_lut) STEM_LAUNCH(uint8_t, 1, true);
  else STEM_LAUNCH(uint8_t, 1, false);
#undef STEM_LAUNCH
  MCLIP_CHECK_LAUNCH();
  return MCLIP_OK;
}
