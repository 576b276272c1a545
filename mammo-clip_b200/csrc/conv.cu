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
#include <algorithm>
using std::max;

#define DW_THREADS 256
#define DW_WARPS 8
#define DW_CCH 64          // channels per CTA tile: one warp lane = 2 channels

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

// ---- tile loader: global NHWC bf16 -> smem [rows][cols][64ch] bf16 with optional affine+swish, zero padded -------------
template <bool kTransform>
__device__ __forceinline__ void dw_load_tile(bf16* __restrict__ tile, const bf16* __restrict__ src, int n, int y0, int x0, int rows,
                                             int cols, int H, int W, int C, int c0, const float* sc, const float* sh, int act) {
  const int v = threadIdx.x & 7;                 // 8-channel vector inside the 64-channel chunk (fixed per thread)
  const int c = c0 + v * 8;
  const bool cvalid = c < C;
  float a[8], b[8];
  if (kTransform && cvalid) {
#pragma unroll
    for (int i = 0; i < 8; ++i) { a[i] = sc[c + i]; b[i] = sh[c + i]; }
  }
  const int npix = rows * cols;
  for (int p = threadIdx.x >> 3; p < npix; p += DW_THREADS / 8) {
    const int ry = p / cols, rx = p - ry * cols;
    const int y = y0 + ry, x = x0 + rx;
    bf16x8 val;
    val.w[0] = val.w[1] = val.w[2] = val.w[3] = 0u;
    if (cvalid && y >= 0 && y < H && x >= 0 && x < W) {
      val = ldg_bf16x8(src + (((size_t)n * H + y) * W + x) * C + c);
      if (kTransform) {
        float f[8];
        unpack8(val, f);
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          float t = fmaf(f[i], a[i], b[i]);
          f[i] = act ? swish_f(t) : t;
        }
        val = pack8(f);
      }
    }
    *reinterpret_cast<bf16x8*>(tile + ((size_t)p * DW_CCH + v * 8)) = val;
  }
}

// =====================================================================================================
// forward
// =====================================================================================================
template <int K, int S, int TH, int TW>
__global__ void __launch_bounds__(DW_THREADS) mclip_dwconv_fwd_kernel(const DwDev p) {
  constexpr int IH = (TH - 1) * S + K, IW = (TW - 1) * S + K;
  constexpr int PR = (2 - 1) * S + K, PC = (4 - 1) * S + K;     // input window of a 2x4 output patch
  extern __shared__ __align__(16) uint8_t smem_dw[];
  bf16* tile = reinterpret_cast<bf16*>(smem_dw);
  __shared__ float red[DW_WARPS][4][32];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int chunk = blockIdx.x % p.n_chunks, slot = blockIdx.x / p.n_chunks;
  const int c0 = chunk * DW_CCH, c = c0 + lane * 2;
  const bool cvalid = c < p.C;
  float w0[K * K], w1[K * K];
#pragma unroll
  for (int t = 0; t < K * K; ++t) {
    w0[t] = cvalid ? p.w[(size_t)c * K * K + t] : 0.f;
    w1[t] = cvalid ? p.w[(size_t)(c + 1) * K * K + t] : 0.f;
  }
  float s_sum0 = 0.f, s_sum1 = 0.f, s_sq0 = 0.f, s_sq1 = 0.f;
  const int tiles_per_img = p.tiles_x * p.tiles_y;
  const int total_tiles = p.N * tiles_per_img;
  for (int t = slot; t < total_tiles; t += p.slots) {
    const int n = t / tiles_per_img, tr = t % tiles_per_img;
    const int oy0 = (tr / p.tiles_x) * TH, ox0 = (tr % p.tiles_x) * TW;
    __syncthreads();
    if (p.scale) dw_load_tile<true>(tile, p.in, n, oy0 * S - p.pt, ox0 * S - p.pl, IH, IW, p.H, p.W, p.C, c0, p.scale, p.shift, p.act);
    else dw_load_tile<false>(tile, p.in, n, oy0 * S - p.pt, ox0 * S - p.pl, IH, IW, p.H, p.W, p.C, c0, nullptr, nullptr, 0);
    __syncthreads();
    for (int pa = warp; pa < (TH / 2) * (TW / 4); pa += DW_WARPS) {
      const int py = (pa / (TW / 4)) * 2, px = (pa % (TW / 4)) * 4;
      if (oy0 + py >= p.Ho || ox0 + px >= p.Wo) continue;        // warp-uniform
      float acc[2][4][2];
#pragma unroll
      for (int i = 0; i < 2; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j][0] = acc[i][j][1] = 0.f;
#pragma unroll
      for (int iy = 0; iy < PR; ++iy) {
        float r0[PC], r1[PC];
        const uint32_t* rowp = reinterpret_cast<const uint32_t*>(tile + ((size_t)((py * S + iy) * IW + px * S) * DW_CCH)) + lane;
#pragma unroll
        for (int ix = 0; ix < PC; ++ix) {
          uint32_t u = rowp[ix * (DW_CCH / 2)];
          r0[ix] = bf16_lo(u);
          r1[ix] = bf16_hi(u);
        }
#pragma unroll
        for (int oy = 0; oy < 2; ++oy) {
          const int ky = iy - oy * S;
          if (ky < 0 || ky >= K) continue;
#pragma unroll
          for (int kx = 0; kx < K; ++kx)
#pragma unroll
            for (int ox = 0; ox < 4; ++ox) {
              acc[oy][ox][0] = fmaf(r0[ox * S + kx], w0[ky * K + kx], acc[oy][ox][0]);
              acc[oy][ox][1] = fmaf(r1[ox * S + kx], w1[ky * K + kx], acc[oy][ox][1]);
            }
        }
      }
#pragma unroll
      for (int oy = 0; oy < 2; ++oy)
#pragma unroll
        for (int ox = 0; ox < 4; ++ox) {
          const int y = oy0 + py + oy, x = ox0 + px + ox;
          if (y < p.Ho && x < p.Wo && cvalid) {
            uint32_t pk = pack_bf16(acc[oy][ox][0], acc[oy][ox][1]);
            *reinterpret_cast<uint32_t*>(p.out + (((size_t)n * p.Ho + y) * p.Wo + x) * p.C + c) = pk;
            float a0 = bf16_lo(pk), a1 = bf16_hi(pk);
            s_sum0 += a0; s_sum1 += a1; s_sq0 = fmaf(a0, a0, s_sq0); s_sq1 = fmaf(a1, a1, s_sq1);
          }
        }
    }
  }
  if (p.stats) {
    red[warp][0][lane] = s_sum0; red[warp][1][lane] = s_sum1; red[warp][2][lane] = s_sq0; red[warp][3][lane] = s_sq1;
    __syncthreads();
    if (warp == 0 && cvalid) {
      float a = 0.f, b = 0.f, cc = 0.f, d = 0.f;
#pragma unroll
      for (int w = 0; w < DW_WARPS; ++w) { a += red[w][0][lane]; b += red[w][1][lane]; cc += red[w][2][lane]; d += red[w][3][lane]; }
      float* st = p.stats + (size_t)slot * 2 * p.C;
      st[c] = a; st[c + 1] = b; st[p.C + c] = cc; st[p.C + c + 1] = d;
    }
  }
}

// =====================================================================================================
// backward: data gradient (+ swish' of the input activation, + input-BN reduction terms) and weight gradient
// =====================================================================================================
// Tile = TH x TW OUTPUT pixels and the S*TH x S*TW INPUT pixels they own.
template <int K, int S, int TH, int TW>
__global__ void __launch_bounds__(DW_THREADS) mclip_dwconv_bwd_kernel(const DwDev p) {
  constexpr int IH = (TH - 1) * S + K, IW = (TW - 1) * S + K;     // activation window feeding the owned outputs (wgrad)
  constexpr int GH = (S == 1) ? TH + K - 1 : TH + (K + 1) / 2;    // dY window feeding the owned inputs (dgrad)
  constexpr int GW = (S == 1) ? TW + K - 1 : TW + (K + 1) / 2;
  extern __shared__ __align__(16) uint8_t smem_dw[];
  bf16* atile = reinterpret_cast<bf16*>(smem_dw);                 // [IH][IW][64] activated input
  bf16* gtile = atile + (size_t)IH * IW * DW_CCH;                 // [GH][GW][64] dY
  float* wsm = reinterpret_cast<float*>(gtile + (size_t)GH * GW * DW_CCH);   // [K*K][64] weights
  __shared__ float red[DW_WARPS][4][32];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int chunk = blockIdx.x % p.n_chunks, slot = blockIdx.x / p.n_chunks;
  const int c0 = chunk * DW_CCH, c = c0 + lane * 2;
  const bool cvalid = c < p.C;
  for (int i = threadIdx.x; i < K * K * DW_CCH; i += DW_THREADS) {
    int t = i / DW_CCH, ch = i % DW_CCH;
    wsm[i] = (c0 + ch < p.C) ? p.w[(size_t)(c0 + ch) * K * K + t] : 0.f;
  }
  float dw0[K * K], dw1[K * K];
#pragma unroll
  for (int t = 0; t < K * K; ++t) dw0[t] = dw1[t] = 0.f;
  float a0 = 1.f, b0 = 0.f, a1 = 1.f, b1 = 0.f, m0 = 0.f, m1 = 0.f, is0 = 1.f, is1 = 1.f;
  if (p.scale && cvalid) { a0 = p.scale[c]; a1 = p.scale[c + 1]; b0 = p.shift[c]; b1 = p.shift[c + 1]; }
  if (p.bn_part && cvalid) { m0 = p.mean[c]; m1 = p.mean[c + 1]; is0 = p.invstd[c]; is1 = p.invstd[c + 1]; }
  float bs0 = 0.f, bs1 = 0.f, bq0 = 0.f, bq1 = 0.f;
  // first dY row/col needed by the owned inputs:  oy >= ceil((y + pt - (K-1)) / S)
  const int tiles_per_img = p.tiles_x * p.tiles_y;
  const int total_tiles = p.N * tiles_per_img;
  for (int t = slot; t < total_tiles; t += p.slots) {
    const int n = t / tiles_per_img, tr = t % tiles_per_img;
    const int oy0 = (tr / p.tiles_x) * TH, ox0 = (tr % p.tiles_x) * TW;
    const int iy0 = oy0 * S, ix0 = ox0 * S;                        // first owned input pixel
    // floor division for possibly negative numerators
    const int gnum_y = iy0 + p.pt - (K - 1), gnum_x = ix0 + p.pl - (K - 1);
    const int gy0 = (S == 1) ? gnum_y : (gnum_y >= 0 ? (gnum_y + 1) / 2 : -((-gnum_y) / 2));
    const int gx0 = (S == 1) ? gnum_x : (gnum_x >= 0 ? (gnum_x + 1) / 2 : -((-gnum_x) / 2));
    __syncthreads();
    if (p.scale) dw_load_tile<true>(atile, p.in, n, iy0 - p.pt, ix0 - p.pl, IH, IW, p.H, p.W, p.C, c0, p.scale, p.shift, p.act);
    else dw_load_tile<false>(atile, p.in, n, iy0 - p.pt, ix0 - p.pl, IH, IW, p.H, p.W, p.C, c0, nullptr, nullptr, 0);
    dw_load_tile<false>(gtile, p.dy, n, gy0, gx0, GH, GW, p.Ho, p.Wo, p.C, c0, nullptr, nullptr, 0);
    __syncthreads();

    // ---- weight gradient over the owned outputs: dW[ky,kx] += dY[oy,ox] * A[oy*S+ky, ox*S+kx] ----
    for (int pa = warp; pa < (TH / 2) * (TW / 4); pa += DW_WARPS) {
      const int py = (pa / (TW / 4)) * 2, px = (pa % (TW / 4)) * 4;
      if (oy0 + py >= p.Ho || ox0 + px >= p.Wo) continue;
      constexpr int PR = (2 - 1) * S + K, PC = (4 - 1) * S + K;
      float g0[2][4], g1[2][4];
#pragma unroll
      for (int oy = 0; oy < 2; ++oy)
#pragma unroll
        for (int ox = 0; ox < 4; ++ox) {
          // dY of the owned outputs sits in gtile at (oy - gy0, ox - gx0); outputs past Ho/Wo were zero-filled
          const int gy = oy0 + py + oy - gy0, gx = ox0 + px + ox - gx0;
          uint32_t u = reinterpret_cast<const uint32_t*>(gtile + ((size_t)(gy * GW + gx) * DW_CCH))[lane];
          g0[oy][ox] = bf16_lo(u); g1[oy][ox] = bf16_hi(u);
        }
#pragma unroll
      for (int iy = 0; iy < PR; ++iy) {
        float r0[PC], r1[PC];
        const uint32_t* rowp = reinterpret_cast<const uint32_t*>(atile + ((size_t)((py * S + iy) * IW + px * S) * DW_CCH)) + lane;
#pragma unroll
        for (int ix = 0; ix < PC; ++ix) { uint32_t u = rowp[ix * (DW_CCH / 2)]; r0[ix] = bf16_lo(u); r1[ix] = bf16_hi(u); }
#pragma unroll
        for (int oy = 0; oy < 2; ++oy) {
          const int ky = iy - oy * S;
          if (ky < 0 || ky >= K) continue;
#pragma unroll
          for (int kx = 0; kx < K; ++kx)
#pragma unroll
            for (int ox = 0; ox < 4; ++ox) {
              dw0[ky * K + kx] = fmaf(g0[oy][ox], r0[ox * S + kx], dw0[ky * K + kx]);
              dw1[ky * K + kx] = fmaf(g1[oy][ox], r1[ox * S + kx], dw1[ky * K + kx]);
            }
        }
      }
    }

    // ---- data gradient over the owned inputs: dA[y,x] = sum_{ky,kx : (y+pt-ky)%S==0} dY[(y+pt-ky)/S, (x+pl-kx)/S] * w[ky,kx] ----
    constexpr int OWN_H = TH * S, OWN_W = TW * S;
    for (int pa = warp; pa < OWN_H * (OWN_W / 4); pa += DW_WARPS) {
      const int ly = pa / (OWN_W / 4), lx = (pa % (OWN_W / 4)) * 4;
      const int y = iy0 + ly;
      if (y >= p.H || ix0 + lx >= p.W) continue;
      float d0[4] = {0.f, 0.f, 0.f, 0.f}, d1[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
      for (int ky = 0; ky < K; ++ky) {
        const int ny = y + p.pt - ky;
        if (S == 2 && (ny & 1)) continue;                          // warp-uniform
        const int gy = (S == 1 ? ny : ny >> 1) - gy0;              // ny >= gnum_y*... inside the window by construction
        if (gy < 0 || gy >= GH) continue;
#pragma unroll
        for (int kx = 0; kx < K; ++kx) {
          const float2 wv = *reinterpret_cast<const float2*>(wsm + (ky * K + kx) * DW_CCH + lane * 2);
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const int nx = ix0 + lx + j + p.pl - kx;
            if (S == 2 && (nx & 1)) continue;
            const int gx = (S == 1 ? nx : nx >> 1) - gx0;
            if (gx < 0 || gx >= GW) continue;
            uint32_t u = reinterpret_cast<const uint32_t*>(gtile + ((size_t)(gy * GW + gx) * DW_CCH))[lane];
            d0[j] = fmaf(bf16_lo(u), wv.x, d0[j]);
            d1[j] = fmaf(bf16_hi(u), wv.y, d1[j]);
          }
        }
      }
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int x = ix0 + lx + j;
        if (x < p.W && cvalid) {
          const size_t off = (((size_t)n * p.H + y) * p.W + x) * p.C + c;
          float v0 = d0[j], v1 = d1[j];
          if (p.scale) {
            // dv = dA * swish'(a*y+b); yhat = (y-mean)*invstd
            uint32_t yu = *reinterpret_cast<const uint32_t*>(p.in + off);
            float y0 = bf16_lo(yu), y1 = bf16_hi(yu);
            if (p.act) { v0 *= swish_grad_f(fmaf(y0, a0, b0)); v1 *= swish_grad_f(fmaf(y1, a1, b1)); }
            uint32_t pk = pack_bf16(v0, v1);
            *reinterpret_cast<uint32_t*>(p.dx + off) = pk;
            v0 = bf16_lo(pk); v1 = bf16_hi(pk);
            bs0 += v0; bs1 += v1;
            bq0 = fmaf(v0, (y0 - m0) * is0, bq0); bq1 = fmaf(v1, (y1 - m1) * is1, bq1);
          } else {
            *reinterpret_cast<uint32_t*>(p.dx + off) = pack_bf16(v0, v1);
          }
        }
      }
    }
  }
  // ---- flush partials ----
  __syncthreads();
  float* wred = reinterpret_cast<float*>(atile);                   // reuse: [warps][K*K][64]
#pragma unroll
  for (int t = 0; t < K * K; ++t) {
    wred[((size_t)warp * K * K + t) * DW_CCH + lane * 2] = dw0[t];
    wred[((size_t)warp * K * K + t) * DW_CCH + lane * 2 + 1] = dw1[t];
  }
  red[warp][0][lane] = bs0; red[warp][1][lane] = bs1; red[warp][2][lane] = bq0; red[warp][3][lane] = bq1;
  __syncthreads();
  for (int i = threadIdx.x; i < K * K * DW_CCH; i += DW_THREADS) {
    const int t = i / DW_CCH, ch = i % DW_CCH;
    if (c0 + ch < p.C) {
      float s = 0.f;
#pragma unroll
      for (int w = 0; w < DW_WARPS; ++w) s += wred[((size_t)w * K * K + t) * DW_CCH + ch];
      p.dw_part[((size_t)slot * K * K + t) * p.C + c0 + ch] = s;
    }
  }
  if (p.bn_part && warp == 0 && cvalid) {
    float a = 0.f, b = 0.f, cc = 0.f, d = 0.f;
#pragma unroll
    for (int w = 0; w < DW_WARPS; ++w) { a += red[w][0][lane]; b += red[w][1][lane]; cc += red[w][2][lane]; d += red[w][3][lane]; }
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
  static constexpr int TH = (S == 1) ? 16 : 8;
  static constexpr int TW = 16;
};

static int dw_slots(int n_chunks, int total_tiles, int per_sm) {
  int s = (mclip_num_sms() * per_sm) / n_chunks;
  if (s < 1) s = 1;
  if (s > total_tiles) s = total_tiles;
  return s;
}

template <int K, int S>
static int dw_launch_fwd(DwDev& p, cudaStream_t stream, int slots_given) {
  constexpr int TH = DwCfg<K, S>::TH, TW = DwCfg<K, S>::TW;
  constexpr int IH = (TH - 1) * S + K, IW = (TW - 1) * S + K;
  const int smem = IH * IW * DW_CCH * 2;
  p.tiles_x = ceil_div(p.Wo, TW); p.tiles_y = ceil_div(p.Ho, TH);
  p.n_chunks = ceil_div(p.C, DW_CCH);
  p.slots = slots_given;
  auto kern = mclip_dwconv_fwd_kernel<K, S, TH, TW>;
  static bool attr = false;
  if (!attr) { MCLIP_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem)); attr = true; }
  kern<<<p.n_chunks * p.slots, DW_THREADS, smem, stream>>>(p);
  MCLIP_CHECK_LAUNCH();
  return MCLIP_OK;
}

template <int K, int S>
static int dw_launch_bwd(DwDev& p, cudaStream_t stream, int slots_given) {
  constexpr int TH = DwCfg<K, S>::TH / 2 * 1, TW = DwCfg<K, S>::TW;   // half-height tiles: two staged tensors
  constexpr int IH = (TH - 1) * S + K, IW = (TW - 1) * S + K;
  constexpr int GH = (S == 1) ? TH + K - 1 : TH + (K + 1) / 2, GW = (S == 1) ? TW + K - 1 : TW + (K + 1) / 2;
  int smem = (IH * IW + GH * GW) * DW_CCH * 2 + K * K * DW_CCH * 4;
  const int red_bytes = DW_WARPS * K * K * DW_CCH * 4;                // flush reuses the front of the buffer
  if (smem < red_bytes) smem = red_bytes;
  // the tile grid must cover every OUTPUT pixel (weight gradient) and every INPUT pixel (data gradient): with the
  // reference's static pads, S*Ho can be smaller than H (e.g. H=33, k3 s2 pads (0,1) -> Ho=16 but input row 32 is read)
  p.tiles_x = ceil_div(max(p.Wo, ceil_div(p.W, S)), TW); p.tiles_y = ceil_div(max(p.Ho, ceil_div(p.H, S)), TH);
  p.n_chunks = ceil_div(p.C, DW_CCH);
  p.slots = slots_given;
  auto kern = mclip_dwconv_bwd_kernel<K, S, TH, TW>;
  static bool attr = false;
  if (!attr) { MCLIP_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem)); attr = true; }
  kern<<<p.n_chunks * p.slots, DW_THREADS, smem, stream>>>(p);
  MCLIP_CHECK_LAUNCH();
  return MCLIP_OK;
}

static int dw_tiles(const mclip_dwconv_args* a, bool bwd) {
  int TH = a->stride == 1 ? 16 : 8, TW = 16;
  if (!bwd) return a->n * ceil_div(a->ho, TH) * ceil_div(a->wo, TW);
  TH /= 2;
  const int hy = max(a->ho, ceil_div(a->h, a->stride)), wx = max(a->wo, ceil_div(a->w, a->stride));
  return a->n * ceil_div(hy, TH) * ceil_div(wx, TW);
}

extern "C" int mclip_dwconv_slots(const mclip_dwconv_args* a, int backward) {
  if (!a || a->c <= 0) return -1;
  return dw_slots(ceil_div(a->c, DW_CCH), dw_tiles(a, backward != 0), backward ? 2 : 3);
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
// stem: dense 3x3 stride-2 convolution, fp32 input with arbitrary strides (the trainer hands NCHW-shaped NHWC memory,
// trainer_ddp.py:288-291), 3 input channels, Cout <= 64, bf16 NHWC output + BN statistics partials
// =====================================================================================================
#define STEM_THREADS 256
#define STEM_MAXC 64

struct StemDev {
  int N, H, W, Ho, Wo, C, pl, pt, slots;
  long long sn, sc, sh, sw;     // element strides of the fp32 input
  const float* in; const float* w;   // w: [C,3,3,3] (OIHW)
  bf16* out; float* stats;
  // backward
  const bf16* dy; float* dw_part;    // [slots][27][C]
};

__global__ void __launch_bounds__(STEM_THREADS) mclip_stem_fwd_kernel(const StemDev p) {
  __shared__ float ws[27][STEM_MAXC];
  __shared__ float red[2][STEM_MAXC];
  for (int i = threadIdx.x; i < 27 * STEM_MAXC; i += STEM_THREADS) {
    int t = i / STEM_MAXC, c = i % STEM_MAXC;     // t = ci*9 + ky*3 + kx
    ws[t][c] = c < p.C ? p.w[(size_t)c * 27 + t] : 0.f;
  }
  if (threadIdx.x < 2 * STEM_MAXC) red[threadIdx.x / STEM_MAXC][threadIdx.x % STEM_MAXC] = 0.f;
  __syncthreads();
  // thread = (pixel, group of 16 output channels): 4 channel groups cover 64 channels
  const int cg = threadIdx.x & 3;
  const int ngroups = (p.C + 15) / 16;
  float ssum[16], ssq[16];
#pragma unroll
  for (int i = 0; i < 16; ++i) ssum[i] = ssq[i] = 0.f;
  const long long npix = (long long)p.N * p.Ho * p.Wo;
  for (long long px = (long long)blockIdx.x * (STEM_THREADS / 4) + (threadIdx.x >> 2); px < npix; px += (long long)gridDim.x * (STEM_THREADS / 4)) {
    if (cg >= ngroups) continue;
    const int ox = (int)(px % p.Wo);
    const int oy = (int)((px / p.Wo) % p.Ho);
    const int n = (int)(px / ((long long)p.Wo * p.Ho));
    float acc[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) acc[i] = 0.f;
#pragma unroll
    for (int ky = 0; ky < 3; ++ky) {
      const int y = oy * 2 - p.pt + ky;
      if (y < 0 || y >= p.H) continue;
#pragma unroll
      for (int kx = 0; kx < 3; ++kx) {
        const int x = ox * 2 - p.pl + kx;
        if (x < 0 || x >= p.W) continue;
        const float* ip = p.in + n * p.sn + y * p.sh + x * p.sw;
#pragma unroll
        for (int ci = 0; ci < 3; ++ci) {
          const float v = __ldg(ip + ci * p.sc);
          const float4* wr = reinterpret_cast<const float4*>(&ws[ci * 9 + ky * 3 + kx][cg * 16]);
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            float4 w4 = wr[q];
            acc[q * 4 + 0] = fmaf(v, w4.x, acc[q * 4 + 0]); acc[q * 4 + 1] = fmaf(v, w4.y, acc[q * 4 + 1]);
            acc[q * 4 + 2] = fmaf(v, w4.z, acc[q * 4 + 2]); acc[q * 4 + 3] = fmaf(v, w4.w, acc[q * 4 + 3]);
          }
        }
      }
    }
    bf16* op = p.out + (size_t)px * p.C + cg * 16;
#pragma unroll
    for (int g = 0; g < 2; ++g) {
      if (cg * 16 + g * 8 < p.C) {
        bf16x8 pk = pack8(acc + g * 8);
        stg_bf16x8(op + g * 8, pk);
        float f[8];
        unpack8(pk, f);
#pragma unroll
        for (int i = 0; i < 8; ++i) { ssum[g * 8 + i] += f[i]; ssq[g * 8 + i] = fmaf(f[i], f[i], ssq[g * 8 + i]); }
      }
    }
  }
  if (p.stats) {
    // threads with equal cg are lanes {cg, cg+4, ...}: reduce over xor 4,8,16 then one shared atomic per warp
#pragma unroll
    for (int i = 0; i < 16; ++i) {
      float a = ssum[i], b = ssq[i];
#pragma unroll
      for (int o = 4; o < 32; o <<= 1) { a += __shfl_xor_sync(0xffffffffu, a, o); b += __shfl_xor_sync(0xffffffffu, b, o); }
      if ((threadIdx.x & 31) < 4) { atomicAdd(&red[0][cg * 16 + i], a); atomicAdd(&red[1][cg * 16 + i], b); }
    }
    __syncthreads();
    if (threadIdx.x < p.C) {
      p.stats[(size_t)blockIdx.x * 2 * p.C + threadIdx.x] = red[0][threadIdx.x];
      p.stats[(size_t)blockIdx.x * 2 * p.C + p.C + threadIdx.x] = red[1][threadIdx.x];
    }
  }
}

// weight gradient of the stem: dW[c, t] = sum_pixels dY[pix, c] * patch[pix, t],  t in 0..26
// CTA stages 64 pixels of (patch[27], dY[C]) in smem; thread owns (tap t, 4 channels) pairs.
__global__ void __launch_bounds__(STEM_THREADS) mclip_stem_wgrad_kernel(const StemDev p) {
  __shared__ float patch[64][28];
  __shared__ __align__(16) float dys[64][STEM_MAXC + 4];
  // work item w = t*16 + cq  (27 taps x 16 channel-quads = 432 items, 2 per thread max)
  float acc[2][4];
#pragma unroll
  for (int i = 0; i < 2; ++i) acc[i][0] = acc[i][1] = acc[i][2] = acc[i][3] = 0.f;
  const long long npix = (long long)p.N * p.Ho * p.Wo;
  for (long long base = (long long)blockIdx.x * 64; base < npix; base += (long long)gridDim.x * 64) {
    __syncthreads();
    for (int i = threadIdx.x; i < 64 * 27; i += STEM_THREADS) {
      const int lp = i / 27, t = i % 27;
      const long long px = base + lp;
      float v = 0.f;
      if (px < npix) {
        const int ox = (int)(px % p.Wo), oy = (int)((px / p.Wo) % p.Ho), n = (int)(px / ((long long)p.Wo * p.Ho));
        const int ci = t / 9, ky = (t % 9) / 3, kx = t % 3;
        const int y = oy * 2 - p.pt + ky, x = ox * 2 - p.pl + kx;
        if (y >= 0 && y < p.H && x >= 0 && x < p.W) v = __ldg(p.in + n * p.sn + ci * p.sc + y * p.sh + x * p.sw);
      }
      patch[lp][t] = v;
    }
    for (int i = threadIdx.x; i < 64 * STEM_MAXC; i += STEM_THREADS) {
      const int lp = i / STEM_MAXC, c = i % STEM_MAXC;
      const long long px = base + lp;
      dys[lp][c] = (px < npix && c < p.C) ? __bfloat162float(p.dy[(size_t)px * p.C + c]) : 0.f;
    }
    __syncthreads();
#pragma unroll
    for (int it = 0; it < 2; ++it) {
      const int wi = threadIdx.x + it * STEM_THREADS;
      if (wi < 27 * 16) {
        const int t = wi >> 4, cq = (wi & 15) * 4;
        for (int lp = 0; lp < 64; ++lp) {
          const float pv = patch[lp][t];
          const float4 d4 = *reinterpret_cast<const float4*>(&dys[lp][cq]);
          acc[it][0] = fmaf(pv, d4.x, acc[it][0]); acc[it][1] = fmaf(pv, d4.y, acc[it][1]);
          acc[it][2] = fmaf(pv, d4.z, acc[it][2]); acc[it][3] = fmaf(pv, d4.w, acc[it][3]);
        }
      }
    }
  }
#pragma unroll
  for (int it = 0; it < 2; ++it) {
    const int wi = threadIdx.x + it * STEM_THREADS;
    if (wi < 27 * 16) {
      const int t = wi >> 4, cq = (wi & 15) * 4;
#pragma unroll
      for (int e = 0; e < 4; ++e)
        if (cq + e < p.C) p.dw_part[((size_t)blockIdx.x * 27 + t) * p.C + cq + e] = acc[it][e];
    }
  }
}

extern "C" int mclip_stem_slots(int n, int ho, int wo) {
  long long npix = (long long)n * ho * wo;
  int g = mclip_num_sms() * 4;
  long long need = (npix + 63) / 64;
  return (int)(need < g ? need : g);
}

static int stem_fill(const mclip_stem_args* a, StemDev& p) {
  MCLIP_REQUIRE(a && a->in && a->weight, "mclip_stem: null operand");
  MCLIP_REQUIRE(a->c % 8 == 0 && a->c <= STEM_MAXC, "mclip_stem: Cout=%d must be a multiple of 8 and <= %d", a->c, STEM_MAXC);
  MCLIP_REQUIRE(a->ho == (a->h + a->pad_top + a->pad_bottom - 3) / 2 + 1 && a->wo == (a->w + a->pad_left + a->pad_right - 3) / 2 + 1,
                "mclip_stem: output size inconsistent with the static padding");
  memset(&p, 0, sizeof(p));
  p.N = a->n; p.H = a->h; p.W = a->w; p.Ho = a->ho; p.Wo = a->wo; p.C = a->c; p.pl = a->pad_left; p.pt = a->pad_top;
  p.sn = a->stride_n; p.sc = a->stride_c; p.sh = a->stride_h; p.sw = a->stride_w;
  p.in = a->in; p.w = a->weight;
  p.slots = mclip_stem_slots(a->n, a->ho, a->wo);
  return MCLIP_OK;
}

extern "C" int mclip_stem_forward(const mclip_stem_args* a, void* stream_) {
  StemDev p;
  int rc = stem_fill(a, p);
  if (rc) return rc;
  MCLIP_REQUIRE(a->out, "mclip_stem_forward: null output");
  if (a->stats) MCLIP_REQUIRE(a->stat_slots == p.slots, "mclip_stem_forward: stat_slots=%d, expected %d", a->stat_slots, p.slots);
  p.out = (bf16*)a->out; p.stats = a->stats;
  mclip_stem_fwd_kernel<<<p.slots, STEM_THREADS, 0, (cudaStream_t)stream_>>>(p);
  MCLIP_CHECK_LAUNCH();
  return MCLIP_OK;
}

extern "C" int mclip_stem_wgrad(const mclip_stem_args* a, void* stream_) {
  StemDev p;
  int rc = stem_fill(a, p);
  if (rc) return rc;
  MCLIP_REQUIRE(a->dy && a->dweight && a->dw_partials, "mclip_stem_wgrad: null operand");
  MCLIP_REQUIRE(a->stat_slots == p.slots, "mclip_stem_wgrad: stat_slots=%d, expected %d", a->stat_slots, p.slots);
  p.dy = (const bf16*)a->dy; p.dw_part = a->dw_partials;
  mclip_stem_wgrad_kernel<<<p.slots, STEM_THREADS, 0, (cudaStream_t)stream_>>>(p);
  MCLIP_CHECK_LAUNCH();
  // dweight is [C,27] (OIHW flattened) which is exactly the [c][t] layout of the depthwise reducer
  mclip_dw_wgrad_reduce_kernel<<<ceil_div(27 * a->c, 256), 256, 0, (cudaStream_t)stream_>>>(a->dw_partials, a->dweight, p.slots, 27, a->c, a->accumulate);
  MCLIP_CHECK_LAUNCH();
  return MCLIP_OK;
}
