// Depthwise convolutions as ROW-STREAMING stencils (replaces the halo-tile kernels of conv.cu for the shapes it covers).
//
// Same contract as conv.cu (MBConvBlock._depthwise_conv, efficientnet_custom.py:66-73,109; static pads of
// Conv2dStaticSamePadding, efficient_net_custom_utils.py:248-276; producer BN+swish applied on load, BN statistics of the
// output, backward = dX (+swish'), dW, input-BN reduction terms), different execution model:
//
//   * a CTA owns a vertical strip (16 output columns x 64 channels) of one image and walks DOWN it.  TMA streams blocks of K
//     input rows into a ring of shared-memory slots (mbarrier full/empty pairs, one elected producer thread), so every input
//     row is fetched once per strip: no vertical halo re-reads, no per-tile pipeline bubble.
//   * a warp owns 4 output columns x 64 channels (lane = 2 channels, packed fp32x2 math).  Per input row it loads the 4+K-1
//     pixels it needs, applies BN+swish IN REGISTERS (fp32, one MUFU per element) and issues K*K*4 FFMA2 into K rolling
//     accumulator rows (output row oy lives in accumulator slot (oy + const) mod K; the loop is unrolled K times so every
//     role is static).  No transform pass over shared memory, no bf16 re-rounding of the activated input, no CTA barrier.
//   * shared memory is addressed through 32-bit shared-window addresses (ld.shared), weights stay in registers.
// Instruction mix per input row (k5): 8 LDS + 16 unpack + 8+8 FFMA2 (affine, swish) + 16 MUFU + 100 FFMA2.
//
// Variants (all measured on the EN-B5 c3 geometry, profiles/r02*):
//   * backward = two warp ROLES (weight gradient / data gradient) over one ring of Y_in + dY rows.  k3: K-times unrolled bodies with
//     static accumulator roles; k5: ONE step body per role and register ROTATION of the rolling rows (the unrolled k5 bodies were
//     5568 SASS instructions and the two roles thrashed the instruction caches), 4-warp CTAs whose roles alternate with the CTA
//     parity so that every SM sub-partition runs both;
//   * CG = 16: the narrow k3 s1 layers (C <= 48) run four 8-lane groups per warp, each on its own 5-column sub-strip of the same 16
//     channels (5 * 32 B apart => four different bank groups), instead of leaving 25-62 % of the lanes of a 64-channel chunk idle;
//   * stride 2: forward with ceil(K/S) rolling rows, backward as a gather over a register window of dY rows.
#include "common.cuh"
#include "mclip_internal.h"
#include <algorithm>
#include <stdlib.h>
#include <type_traits>

typedef unsigned long long u64;

#ifdef DWS_TIMING
__device__ unsigned long long dws_timing[8];
extern "C" int mclip_dws_timing(unsigned long long* out8, int reset) {
  if (reset) { unsigned long long z[8] = {0}; cudaMemcpyToSymbol(dws_timing, z, sizeof(z)); return 0; }
  cudaDeviceSynchronize();
  cudaMemcpyFromSymbol(out8, dws_timing, 8 * sizeof(unsigned long long));
  return 0;
}
#define DWS_T0() const long long _t0 = clock64()
#define DWS_T1(i) _tacc[i] += clock64() - _t0
#else
#define DWS_T0()
#define DWS_T1(i)
#endif

namespace {

struct DwsDev {
  int N, H, W, C, Ho, Wo, pl, pt;
  int strips_x, segs, seg_rows, n_chunks, slots, items;      // items = N * strips_x * segs (per channel chunk)
  const bf16* in; const float* scale; const float* shift; int act;
  const float* w;
  bf16* out; float* stats;
  // backward
  const bf16* dy; bf16* dx; float* dw_part; float* bn_part; const float* mean; const float* invstd;
};

__device__ __forceinline__ float2 ffma2r(const float2& a, const float2& b, const float2& c) {
  float2 d;
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(reinterpret_cast<u64&>(d)) : "l"(reinterpret_cast<const u64&>(a)), "l"(reinterpret_cast<const u64&>(b)),
      "l"(reinterpret_cast<const u64&>(c)));
  return d;
}
__device__ __forceinline__ void ffma2(float2& d, const float2& a, const float2& b) { d = ffma2r(a, b, d); }
__device__ __forceinline__ float2 fmul2(const float2& a, const float2& b) {
  float2 d;
  asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(reinterpret_cast<u64&>(d)) : "l"(reinterpret_cast<const u64&>(a)), "l"(reinterpret_cast<const u64&>(b)));
  return d;
}
// bf16x2 word -> two fp32: byte permute (ALU pipe) for the low half, mask for the high half; keeps the FMA pipe for FFMA2
__device__ __forceinline__ float2 bf2_to_f2(uint32_t u) {
  return make_float2(__uint_as_float(__byte_perm(u, 0u, 0x1044)), __uint_as_float(u & 0xffff0000u));
}
// predicated 32-bit global store without a branch (the compiler turns `if (p) *ptr = v` into a BSSY/BRA/BSYNC region per store)
__device__ __forceinline__ void stg32_if(void* ptr, uint32_t v, bool pred) {
  asm volatile("{\n\t.reg .pred q;\n\tsetp.ne.u32 q, %2, 0;\n\t@q st.global.b32 [%0], %1;\n\t}" ::"l"(ptr), "r"(v), "r"((uint32_t)pred) : "memory");
}
// arrival counter of a ring slot: acq_rel at CTA scope, so the warp that observes the last arrival also observes that every
// other warp has finished reading the slot (and may hand it back to TMA)
__device__ __forceinline__ uint32_t atom_add_acqrel_smem(uint32_t* p, uint32_t v) {
  uint32_t old;
  asm volatile("atom.acq_rel.cta.shared::cta.add.u32 %0, [%1], %2;" : "=r"(old) : "r"(smem_u32(p)), "r"(v) : "memory");
  return old;
}
__device__ __forceinline__ uint32_t lds32(uint32_t addr) {
  uint32_t v;
  asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(addr));
  return v;
}

// -------------------------------------------------------------------------------------------------------------------
// forward, stride 1
// -------------------------------------------------------------------------------------------------------------------
#ifndef DWS_K3_CTAS
#define DWS_K3_CTAS 4
#endif
#ifndef DWS_K5_CTAS
#define DWS_K5_CTAS 3
#endif
#ifndef DWS_K3_NSLOT
#define DWS_K3_NSLOT 3
#endif
#ifndef DWS_K5_NSLOT
#define DWS_K5_NSLOT 3
#endif
#ifndef DWS_K3_REP
#define DWS_K3_REP 2
#endif
#ifndef DWS_K5_REP
#define DWS_K5_REP 1
#endif

// Walks the (item, block) sequence of one CTA.  Kept in EVERY thread (all values are warp-uniform) so that the TMA issuer role
// can rotate over the warps: no warp carries the producer's latency on its critical path every block.
struct DwsIter {
  int item, blk, nblk, n, x0, r0, rows;
};

// compile-time configuration of the forward kernel for (kernel size, stride)
// CG = channels per lane group: 64 (a warp = 32 lanes x 2 channels of ONE 4-column strip) or 16 (a warp = 4 groups of 8 lanes, each
// group a different 5-column sub-strip of the same 16 channels: the C <= 48 layers of blocks 0-2 would leave 25-62 % of the lanes idle
// with 64-channel chunks).  A pixel is CG*2 bytes in shared memory; 5-column sub-strips start 5*32 bytes apart, so the four groups of a
// warp hit four different 32-byte bank groups on every load (4-column strips would all hit the same one).
template <int K, int S, int CG = 64>
struct FwdCfg {
  static constexpr int SUBS = 64 / CG, LPG = CG / 2, PIXB = CG * 2;      // sub-strips per warp, lanes per group, bytes per pixel
  static constexpr int SW = CG == 64 ? 4 : 5, NW = 4, TW = SW * SUBS * NW;
  static constexpr int NA = (K + S - 1) / S;         // live accumulator rows: output row oy lives in slot oy % NA
  static constexpr int G = S * NA;                   // input rows per fully unrolled group (all slot roles static)
  static constexpr int REP = (S == 1 && K == 3) ? DWS_K3_REP : (S == 1 ? DWS_K5_REP : 1);
  static constexpr int RB = G * REP;                 // input rows per ring slot
  static constexpr int IW = (TW - 1) * S + K, PC = (SW - 1) * S + K;
  static constexpr int NSLOT = (S == 1) ? (K == 3 ? DWS_K3_NSLOT : DWS_K5_NSLOT) : (K == 3 ? 3 : 2);
  static constexpr int CTAS = (S == 1) ? (K == 3 ? DWS_K3_CTAS : DWS_K5_CTAS) : (K == 3 ? 4 : 3);
  static constexpr int SMEM = NSLOT * RB * IW * PIXB;
  static_assert((RB * IW * PIXB) % 128 == 0, "ring slots must stay 128-byte aligned for TMA");
};

template <int K, int S, bool ACT, int CG>
__global__ void __launch_bounds__(128, FwdCfg<K, S, CG>::CTAS) mclip_dws_fwd_kernel(const __grid_constant__ CUtensorMap tmIn, const DwsDev p) {
  using Cfg = FwdCfg<K, S, CG>;
  constexpr int SW = Cfg::SW, NW = Cfg::NW, TW = Cfg::TW, IW = Cfg::IW, PC = Cfg::PC, RB = Cfg::RB, NA = Cfg::NA, G = Cfg::G, REP = Cfg::REP, NSLOT = Cfg::NSLOT;
  constexpr int SUBS = Cfg::SUBS, LPG = Cfg::LPG;
  constexpr uint32_t PIXB = Cfg::PIXB, ROW_BYTES = IW * PIXB, SLOT_BYTES = RB * ROW_BYTES;
  extern __shared__ __align__(1024) uint8_t dws_smem[];
  __shared__ float red[NW][4][32];
  __shared__ __align__(8) uint64_t full[NSLOT];
  __shared__ uint32_t arrivals[NSLOT];               // monotonic: arrival number a of a slot is the last of its round iff a % NW == NW-1
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int chunk = blockIdx.x % p.n_chunks, slot = blockIdx.x / p.n_chunks;
  const int sub = lane / LPG, gl = lane % LPG;        // lane group (sub-strip) and lane within the group
  const int c0 = chunk * CG, c = c0 + gl * 2;
  const bool cvalid = c < p.C;
  if (threadIdx.x == 0) {
    tma_prefetch_desc(&tmIn);
    for (int s = 0; s < NSLOT; ++s) { mbar_init(&full[s], 1); arrivals[s] = 0; }
    fence_mbar_init();
  }
  __syncthreads();
  const uint32_t ring = smem_u32(dws_smem);
  const int my_items = p.items > slot ? (p.items - slot + p.slots - 1) / p.slots : 0;
  const int per_img = p.strips_x * p.segs;

  auto load_item = [&](DwsIter& it) {            // geometry of item `it.item` of this CTA (two divisions per item)
    const int g = slot + it.item * p.slots;
    it.n = g / per_img;
    const int rem = g - it.n * per_img;
    const int sy = rem / p.strips_x;
    it.x0 = (rem - sy * p.strips_x) * TW;
    it.r0 = sy * p.seg_rows;
    it.rows = min(p.Ho, it.r0 + p.seg_rows) - it.r0;
    it.nblk = (S * (it.rows - 1) + K + RB - 1) / RB;      // local steps j = 0 .. S*(rows-1)+K-1 : input row S*r0 - pt + j
    it.blk = 0;
  };
  // ---- producer ----
  // Every warp carries the iterator `pi` of the block that is NSLOT blocks ahead of the one it is consuming (warp-uniform
  // integer work, once per block).  The warp whose arrival is the LAST on a slot refills that slot at once: nobody ever waits
  // to issue, and the issue latency is on no warp's critical path as long as the ring has a block of slack.
  DwsIter pi;
  pi.item = 0;
  auto advance = [&]() { if (++pi.blk == pi.nblk) { if (++pi.item < my_items) load_item(pi); } };
  auto issue = [&](int s) {                          // one thread: block `pi` -> slot s
    mbar_expect_tx(&full[s], SLOT_BYTES);
    tma_load_4d(dws_smem + (size_t)s * SLOT_BYTES, &tmIn, &full[s], c0, S * pi.x0 - p.pl, S * pi.r0 - p.pt + pi.blk * RB, pi.n);
  };
  if (my_items > 0) {
    load_item(pi);
#pragma unroll 1
    for (int i = 0; i < NSLOT; ++i) {
      if (pi.item < my_items) {
        if (threadIdx.x == 0) issue(i);
        advance();
      }
    }
  }

  // ---- consumer state ----
  float2 w[K * K];
#pragma unroll
  for (int t = 0; t < K * K; ++t) w[t] = cvalid ? make_float2(p.w[(size_t)c * K * K + t], p.w[(size_t)(c + 1) * K * K + t]) : make_float2(0.f, 0.f);
  const float f = ACT ? 0.5f : 1.0f;                // swish(t) = h + h*tanh(h), h = t/2
  float2 a2 = make_float2(f, f), b2 = make_float2(0.f, 0.f);
  if (p.scale && cvalid) { a2 = make_float2(f * p.scale[c], f * p.scale[c + 1]); b2 = make_float2(f * p.shift[c], f * p.shift[c + 1]); }
  float2 s_sum = make_float2(0.f, 0.f), s_sq = make_float2(0.f, 0.f);
  const float2 one = make_float2(1.f, 1.f);
  const int cw = p.C >> 1;
  const int rstride = p.Wo * cw;                     // output row stride in 32-bit words
  const unsigned rstride_b = (unsigned)rstride * 4u, pix_b = (unsigned)p.C * 2u;     // ... and row / pixel strides in bytes
  DwsIter ci;
  ci.item = 0;
  int count = 0;                                     // blocks consumed so far (ring position)
#ifdef DWS_TIMING
  long long _tacc[4] = {0, 0, 0, 0};
  const long long _tk = clock64();
#endif
#pragma unroll 1
  for (; ci.item < my_items; ++ci.item) {
    load_item(ci);
    const int wx = ci.x0 + (warp * SUBS + sub) * SW;  // first output column of this lane group
    const bool wactive = wx < p.Wo;
    // validity masks of the warp's PC input columns / SW output columns
    uint32_t inmask = 0, outmask = 0;
#pragma unroll
    for (int ix = 0; ix < PC; ++ix) { const int gx = S * wx - p.pl + ix; inmask |= (gx >= 0 && gx < p.W) ? (1u << ix) : 0u; }
#pragma unroll
    for (int j = 0; j < SW; ++j) outmask |= (wx + j < p.Wo && cvalid) ? (1u << j) : 0u;
    const bool edge = inmask != ((1u << PC) - 1u);
    // this lane's word of output pixel (r0, wx); row oj of the item starts rstride words further per row
    char* const obase = reinterpret_cast<char*>(p.out) + (((size_t)ci.n * p.Ho + ci.r0) * (size_t)rstride + (size_t)wx * cw + (c >> 1)) * 4;
    const int iy0 = S * ci.r0 - p.pt;                // input row of local step 0
    // Local step j = input row iy0 + j feeds, with tap row ky, the output row (j - ky) / S of the item (when S divides j - ky);
    // that output row lives in accumulator slot ((j - ky) / S) % NA, is initialised by its ky == 0 tap and complete after its
    // ky == K-1 tap.  Groups of G = S*NA steps start at multiples of G, so every role below is a compile-time constant.
    float2 acc[NA][SW];
#pragma unroll
    for (int j = 0; j < NA; ++j)
#pragma unroll
      for (int o = 0; o < SW; ++o) acc[j][o] = make_float2(0.f, 0.f);

    // one input row (static position s within the group): FAST = row inside the image, no column masks, all SW outputs stored
    auto step = [&](auto fast_c, auto s_c, uint32_t base, int j0) {
      constexpr bool FAST = decltype(fast_c)::value;
      constexpr int s = decltype(s_c)::value;
      if (FAST || (unsigned)(iy0 + j0 + s) < (unsigned)p.H) {
        float2 x[PC];
#pragma unroll
        for (int ix = 0; ix < PC; ++ix) {
          float2 h = ffma2r(bf2_to_f2(lds32(base + s * ROW_BYTES + ix * PIXB)), a2, b2);
          if (ACT) h = ffma2r(h, make_float2(fast_tanh(h.x), fast_tanh(h.y)), h);
          x[ix] = h;
        }
        if (!FAST && edge) {                       // ZeroPad2d acts on the activated tensor: columns outside the image are 0
#pragma unroll
          for (int ix = 0; ix < PC; ++ix)
            if (!((inmask >> ix) & 1u)) x[ix] = make_float2(0.f, 0.f);
        }
#pragma unroll
        for (int ky = 0; ky < K; ++ky) {
          if ((s - ky + K * S) % S != 0) continue;             // this input row is not on output row (j-ky)/S's tap row ky
          constexpr int dummy = 0; (void)dummy;
          const int q = (((s - ky + K * S) / S - K) % NA + NA) % NA;       // floor((s-ky)/S) mod NA
#pragma unroll
          for (int kx = 0; kx < K; ++kx)
#pragma unroll
            for (int o = 0; o < SW; ++o) {
              if (ky == 0 && kx == 0) acc[q][o] = fmul2(x[S * o], w[0]);
              else ffma2(acc[q][o], x[S * o + kx], w[ky * K + kx]);
            }
        }
      } else if (s % S == 0) {                     // padding row: the slot that this step would have initialised starts at zero
        constexpr int qz = (s / S) % NA;
#pragma unroll
        for (int o = 0; o < SW; ++o) acc[qz][o] = make_float2(0.f, 0.f);
      }
      if ((s - (K - 1) + K * S) % S == 0) {        // an output row received its last tap row
        constexpr int qc = (((s - (K - 1) + K * S) / S - K) % NA + NA) % NA;
        const int oj = (j0 + s - (K - 1)) / S;     // exact: S divides j0 (multiple of G) and s-(K-1)
        if (FAST || (j0 + s >= K - 1 && oj < ci.rows)) {
          char* op = obase + (size_t)((unsigned)oj * rstride_b);
#pragma unroll
          for (int o = 0; o < SW; ++o) {
            // lanes past C hold zero weights, hence zero accumulators: only their stores need a predicate
            const bool st = FAST ? cvalid : (((outmask >> o) & 1u) != 0);
            stg32_if(op + (unsigned)o * pix_b, pack_bf16(acc[qc][o].x, acc[qc][o].y), st);
            if (FAST || st) {
              ffma2(s_sum, acc[qc][o], one);       // statistics from the fp32 accumulators (bf16 rounding of the stored value is unbiased)
              ffma2(s_sq, acc[qc][o], acc[qc][o]);
            }
          }
        }
      }
    };
    const bool warp_fast = !edge && wx + SW <= p.Wo;
#pragma unroll 1
    for (int blk = 0; blk < ci.nblk; ++blk, ++count) {
      const int sl = count % NSLOT;
      { DWS_T0(); mbar_wait(&full[sl], (uint32_t)(count / NSLOT) & 1u); DWS_T1(0); }
#ifdef DWS_TIMING
      const long long _tc = clock64();
#endif
#ifdef DWS_NOCOMPUTE
      if (false) {
#else
      if (wactive) {
#endif
#pragma unroll 1
        for (int rep = 0; rep < REP; ++rep) {
          const uint32_t base = ring + (uint32_t)sl * SLOT_BYTES + (uint32_t)(rep * G) * ROW_BYTES + (uint32_t)((warp * SUBS + sub) * SW * S) * PIXB + (uint32_t)gl * 4u;
          const int j0 = blk * RB + rep * G;         // local step of position s = 0
          // all G input rows inside the image and every output row completing in this group inside the segment?
          // (first completion: step >= K-1; last completing output: (j0 + G - 1 - (K-1)) / S rounded down to a completion step)
          const bool fast = warp_fast && iy0 + j0 >= 0 && iy0 + j0 + G <= p.H && j0 >= K - 1 && (j0 + G - 1 - (K - 1)) / S < ci.rows;
          if (fast) {
            step(std::true_type{}, std::integral_constant<int, 0>{}, base, j0);
            step(std::true_type{}, std::integral_constant<int, 1>{}, base, j0);
            step(std::true_type{}, std::integral_constant<int, 2>{}, base, j0);
            if constexpr (G > 3) step(std::true_type{}, std::integral_constant<int, 3>{}, base, j0);
            if constexpr (G > 4) step(std::true_type{}, std::integral_constant<int, 4>{}, base, j0);
            if constexpr (G > 5) step(std::true_type{}, std::integral_constant<int, 5>{}, base, j0);
          } else {
            step(std::false_type{}, std::integral_constant<int, 0>{}, base, j0);
            step(std::false_type{}, std::integral_constant<int, 1>{}, base, j0);
            step(std::false_type{}, std::integral_constant<int, 2>{}, base, j0);
            if constexpr (G > 3) step(std::false_type{}, std::integral_constant<int, 3>{}, base, j0);
            if constexpr (G > 4) step(std::false_type{}, std::integral_constant<int, 4>{}, base, j0);
            if constexpr (G > 5) step(std::false_type{}, std::integral_constant<int, 5>{}, base, j0);
          }
        }
      }
      __syncwarp();
#ifdef DWS_TIMING
      _tacc[1] += clock64() - _tc; _tacc[3] += 1;
#endif
      { DWS_T0();
      if (pi.item < my_items) {                      // block count + NSLOT exists: the last warp to leave slot sl fetches it
        if (lane == 0 && (atom_add_acqrel_smem(&arrivals[sl], 1u) % NW) == NW - 1) {
          fence_proxy_async_smem();                  // generic-proxy reads of the slot before the async-proxy (TMA) overwrite
          issue(sl);
        }
        advance();
      }
      DWS_T1(2); }
    }
  }
#ifdef DWS_TIMING
  if (lane == 0) {
    for (int i = 0; i < 4; ++i) atomicAdd(&dws_timing[i], (unsigned long long)_tacc[i]);
    atomicAdd(&dws_timing[4], (unsigned long long)(clock64() - _tk));
    atomicAdd(&dws_timing[5], 1ull);
  }
#endif
  if (p.stats) {
    red[warp][0][lane] = s_sum.x; red[warp][1][lane] = s_sum.y; red[warp][2][lane] = s_sq.x; red[warp][3][lane] = s_sq.y;
    __syncthreads();
    if (warp == 0 && cvalid && sub == 0) {
      float a = 0.f, b = 0.f, cc = 0.f, d = 0.f;
#pragma unroll
      for (int w2 = 0; w2 < NW; ++w2)
#pragma unroll
        for (int s2 = 0; s2 < SUBS; ++s2) {
          const int l2 = s2 * LPG + gl;
          a += red[w2][0][l2]; b += red[w2][1][l2]; cc += red[w2][2][l2]; d += red[w2][3][l2];
        }
      float* stp = p.stats + (size_t)slot * 2 * p.C;
      stp[c] = a; stp[c + 1] = b; stp[p.C + c] = cc; stp[p.C + c + 1] = d;
    }
  }
}

// -------------------------------------------------------------------------------------------------------------------
// backward, stride 1 (H == Ho, W == Wo): one pass over (Y_in, dY) produces dX (times swish'), dW partials and the input
// BatchNorm's reduction terms.  A CTA = 4 column strips x 2 ROLES (8 warps): for every strip one warp accumulates the weight
// gradient (K*K register accumulators for the whole kernel, a rolling window of K dY rows, the activated input row) and one
// warp computes the data gradient (the forward's rolling-accumulator scheme on dY with the flipped kernel, then swish' and the
// BN sums on the completed row).  Splitting the roles halves the register state per warp (25 weights OR 25 dW accumulators for
// k5) and lets both streams share one ring: slot = K*REP rows of Y_in + the K*REP rows of dY shifted by pad_top.
// -------------------------------------------------------------------------------------------------------------------
#ifndef DWS_BWD_NSLOT
#define DWS_BWD_NSLOT 3
#endif
template <int K, int NSLOT, int REP, bool BN, int NS, bool ROT, int CG>
__global__ void __launch_bounds__(NS * 64, NS == 4 ? 2 : 3) mclip_dws_bwd_s1_kernel(const __grid_constant__ CUtensorMap tmIn, const __grid_constant__ CUtensorMap tmDy, const DwsDev p) {
  constexpr int SUBS = 64 / CG, LPG = CG / 2;          // lane groups per warp / lanes per group (see FwdCfg)
  constexpr int SW = CG == 64 ? 4 : 5, NW = 2 * NS, TW = SW * SUBS * NS, IW = TW + K - 1, PC = SW + K - 1, RB = K * REP, KK = K;
  constexpr uint32_t PIXB = CG * 2, ROW_BYTES = IW * PIXB, PART_BYTES = RB * ROW_BYTES, SLOT_BYTES = 2 * PART_BYTES;
  static_assert(PART_BYTES % 128 == 0, "ring parts must stay 128-byte aligned for TMA");
  extern __shared__ __align__(1024) uint8_t dws_smem[];
  __shared__ float red[NS][4][32];
  __shared__ __align__(8) uint64_t full[NSLOT];
  __shared__ uint32_t arrivals[NSLOT];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  // NS == 4: first NS warps weight gradient, last NS warps data gradient (every SM sub-partition gets one of each).  NS == 2 (four
  // warps, one per sub-partition): alternate the roles with the CTA parity so that no sub-partition runs only the heavier role
  const int strip = NS == 2 ? (warp >> 1) : (warp % NS);
  const bool wrole = NS == 2 ? (((warp ^ blockIdx.x) & 1) == 0) : (warp < NS);
  const int chunk = blockIdx.x % p.n_chunks, slot = blockIdx.x / p.n_chunks;
  const int sub = lane / LPG, gl = lane % LPG;
  const int c0 = chunk * CG, c = c0 + gl * 2;
  const bool cvalid = c < p.C;
  if (threadIdx.x == 0) {
    tma_prefetch_desc(&tmIn); tma_prefetch_desc(&tmDy);
    for (int s = 0; s < NSLOT; ++s) { mbar_init(&full[s], 1); arrivals[s] = 0; }
    fence_mbar_init();
  }
  __syncthreads();
  const uint32_t ring = smem_u32(dws_smem);
  const int my_items = p.items > slot ? (p.items - slot + p.slots - 1) / p.slots : 0;
  const int per_img = p.strips_x * p.segs;

  auto load_item = [&](DwsIter& it) {
    const int g = slot + it.item * p.slots;
    it.n = g / per_img;
    const int rem = g - it.n * per_img;
    const int sy = rem / p.strips_x;
    it.x0 = (rem - sy * p.strips_x) * TW;
    it.r0 = sy * p.seg_rows;
    it.rows = min(p.H, it.r0 + p.seg_rows) - it.r0;
    // steps t = r0-(K-1) .. r1+K-2-pt : the data gradient of row r0 starts K-1 dY rows early, the weight gradient of dY row
    // r1-1 ends at input row r1-1-pt+K-1
    it.nblk = (it.rows + 2 * K - 2 - p.pt + RB - 1) / RB;
    it.blk = 0;
  };
  DwsIter pi;
  pi.item = 0;
  auto advance = [&]() { if (++pi.blk == pi.nblk) { if (++pi.item < my_items) load_item(pi); } };
  auto issue = [&](int s) {
    const int t = pi.r0 - (K - 1) + pi.blk * RB;
    uint8_t* dst = dws_smem + (size_t)s * SLOT_BYTES;
    mbar_expect_tx(&full[s], SLOT_BYTES);
    tma_load_4d(dst, &tmIn, &full[s], c0, pi.x0 - p.pl, t, pi.n);
    tma_load_4d(dst + PART_BYTES, &tmDy, &full[s], c0, pi.x0 + p.pl - (K - 1), t + p.pt, pi.n);
  };
  if (my_items > 0) {
    load_item(pi);
#pragma unroll 1
    for (int i = 0; i < NSLOT; ++i) {
      if (pi.item < my_items) {
        if (threadIdx.x == 0) issue(i);
        advance();
      }
    }
  }

  // ---- per-role register state ----
  // wgrad role: wv = dW accumulators; dgrad role: wv = the weights (wv[ky*K+kx] = w[ky][kx], the flip is in the tap indexing)
  float2 wv[K * K];
#pragma unroll
  for (int t = 0; t < K * K; ++t)
    wv[t] = (!wrole && cvalid) ? make_float2(p.w[(size_t)c * K * K + t], p.w[(size_t)(c + 1) * K * K + t]) : make_float2(0.f, 0.f);
  const float f = BN ? 0.5f : 1.0f;                  // BN <=> the input is a pre-BN tensor followed by swish
  float2 a2 = make_float2(f, f), b2 = make_float2(0.f, 0.f), mu_is = make_float2(1.f, 1.f), nmis = make_float2(0.f, 0.f);
  if (BN && cvalid) {
    a2 = make_float2(f * p.scale[c], f * p.scale[c + 1]); b2 = make_float2(f * p.shift[c], f * p.shift[c + 1]);
    mu_is = make_float2(p.invstd[c], p.invstd[c + 1]);
    nmis = make_float2(-p.mean[c] * mu_is.x, -p.mean[c + 1] * mu_is.y);
  }
  float2 bs = make_float2(0.f, 0.f), bq = make_float2(0.f, 0.f);
  const float2 one = make_float2(1.f, 1.f), half2 = make_float2(0.5f, 0.5f), two2 = make_float2(2.f, 2.f), neg1 = make_float2(-1.f, -1.f);
  const int cw = p.C >> 1;
  const unsigned rstride_b = (unsigned)(p.W * cw) * 4u, pix_b = (unsigned)p.C * 2u;
  DwsIter ci;
  ci.item = 0;
  int count = 0;
#pragma unroll 1
  for (; ci.item < my_items; ++ci.item) {
    load_item(ci);
    const int wx = ci.x0 + (strip * SUBS + sub) * SW;
    const bool wactive = wx < p.W;
    uint32_t inmask = 0, outmask = 0;
#pragma unroll
    for (int ix = 0; ix < PC; ++ix) { const int gx = wx - p.pl + ix; inmask |= (gx >= 0 && gx < p.W) ? (1u << ix) : 0u; }
#pragma unroll
    for (int j = 0; j < SW; ++j) outmask |= (wx + j < p.W && cvalid) ? (1u << j) : 0u;
    const bool edge = inmask != ((1u << PC) - 1u);
    const bool warp_fast = wrole ? !edge : (wx + SW <= p.W);
    char* const obase = reinterpret_cast<char*>(p.dx) + (((size_t)ci.n * p.H) * (size_t)(p.W * cw) + (size_t)wx * cw + (c >> 1)) * 4;
    const int t0 = ci.r0 - (K - 1), r1 = ci.r0 + ci.rows;
    // dgrad role: acc[q] = partial data gradient of the input row completing (K + q - r) % K steps ahead
    // wgrad role: acc[q] = dY row (4 pixels) of the step congruent to q: the rolling window of the last K dY rows
    float2 acc[K][SW];
#pragma unroll
    for (int j = 0; j < K; ++j)
#pragma unroll
      for (int o = 0; o < SW; ++o) acc[j][o] = make_float2(0.f, 0.f);

    auto wstep = [&](auto fast_c, auto r_c, uint32_t ybase, uint32_t gbase, int t) {       // weight-gradient role, input row t
      constexpr bool FAST = decltype(fast_c)::value;
      constexpr int r = decltype(r_c)::value;
      // newest dY row t+pt enters the window (only rows of this segment count: every (dY row, tap) pair is summed once)
      const int oy = t + p.pt;
      if (FAST || (oy >= ci.r0 && oy < r1)) {
#pragma unroll
        for (int o = 0; o < SW; ++o) acc[r][o] = bf2_to_f2(lds32(gbase + r * ROW_BYTES + (K - 1 + o) * PIXB - p.pl * PIXB));
      } else {
#pragma unroll
        for (int o = 0; o < SW; ++o) acc[r][o] = make_float2(0.f, 0.f);
      }
      if (FAST || (unsigned)t < (unsigned)p.H) {
        float2 x[PC];
#pragma unroll
        for (int ix = 0; ix < PC; ++ix) {
          float2 h = ffma2r(bf2_to_f2(lds32(ybase + r * ROW_BYTES + ix * PIXB)), a2, b2);
          if (BN) h = ffma2r(h, make_float2(fast_tanh(h.x), fast_tanh(h.y)), h);
          x[ix] = h;
        }
        if (!FAST && edge) {
#pragma unroll
          for (int ix = 0; ix < PC; ++ix)
            if (!((inmask >> ix) & 1u)) x[ix] = make_float2(0.f, 0.f);
        }
#pragma unroll
        for (int ky = 0; ky < K; ++ky) {             // dY row t+pt-ky sits in window slot (r - ky) mod K
          constexpr int dummy = 0; (void)dummy;
          const int q = (r - ky + KK) % KK;
#pragma unroll
          for (int kx = 0; kx < K; ++kx)
#pragma unroll
            for (int o = 0; o < SW; ++o) ffma2(wv[ky * K + kx], acc[q][o], x[o + kx]);
        }
      }
    };
    auto dstep = [&](auto fast_c, auto r_c, uint32_t ybase, uint32_t gbase, int t) {       // data-gradient role, dY row t+pt
      constexpr bool FAST = decltype(fast_c)::value;
      constexpr int r = decltype(r_c)::value;
      float2 x[PC];                                    // dY[t+pt][wx + pl - (K-1) + j]; TMA zero-fills rows/columns outside dY
#pragma unroll
      for (int j = 0; j < PC; ++j) x[j] = bf2_to_f2(lds32(gbase + r * ROW_BYTES + j * PIXB));
#pragma unroll
      for (int q = 0; q < K; ++q) {
        const int ky = (q - r + KK) % KK;              // slot q completes ky steps ahead: input row t+ky takes tap row ky
#pragma unroll
        for (int kx = 0; kx < K; ++kx)
#pragma unroll
          for (int o = 0; o < SW; ++o) {
            if (ky == K - 1 && kx == 0) acc[q][o] = fmul2(x[o + K - 1], wv[ky * K]);
            else ffma2(acc[q][o], x[o + K - 1 - kx], wv[ky * K + kx]);
          }
      }
      if (FAST || (t >= ci.r0 && t < r1)) {            // input row t is complete (slot r)
        char* op = obase + (size_t)((unsigned)t * rstride_b);
#pragma unroll
        for (int o = 0; o < SW; ++o) {
          const bool st = FAST ? cvalid : (((outmask >> o) & 1u) != 0);
          float2 d = acc[r][o];
          if (BN) {
            // dv = dA * swish'(v), v = a*y+b; swish'(v) = s + s(1-s)v with s = sigma(v) = 0.5 + 0.5 tanh(v/2)
            const float2 yv = bf2_to_f2(lds32(ybase + r * ROW_BYTES + (o * PIXB) + p.pl * PIXB));
            const float2 hv = ffma2r(yv, a2, b2);                                          // v / 2
            const float2 sg = ffma2r(make_float2(fast_tanh(hv.x), fast_tanh(hv.y)), half2, half2);
            const float2 om = ffma2r(sg, neg1, one);                                       // 1 - s
            const float2 qq = ffma2r(fmul2(hv, om), two2, one);                            // 1 + v (1 - s)
            d = fmul2(d, fmul2(sg, qq));
            if (!FAST && !st) d = make_float2(0.f, 0.f);
            ffma2(bs, d, one);                         // BN-backward sums from the fp32 values (the stored bf16 rounding is unbiased)
            ffma2(bq, d, ffma2r(yv, mu_is, nmis));     // yhat = (y - mean) * invstd
          }
          stg32_if(op + (unsigned)o * pix_b, pack_bf16(d.x, d.y), st);
        }
      }
    };
#pragma unroll 1
    for (int blk = 0; blk < ci.nblk; ++blk, ++count) {
      const int s = count % NSLOT;
      mbar_wait(&full[s], (uint32_t)(count / NSLOT) & 1u);
      if (wactive && ROT) {
        // ROT: ONE step body per role (the K-times unrolled static-role bodies of k5 are 5500 SASS instructions, and the two roles
        // thrash the instruction caches: 38 % of the stall samples were no_instruction); the rolling window / accumulator rows
        // are rotated with register moves instead (16 float2 per 100 FFMA2)
        const uint32_t sbase = ring + (uint32_t)s * SLOT_BYTES + (uint32_t)gl * 4u + (uint32_t)((strip * SUBS + sub) * SW) * PIXB;
        const int tb = t0 + blk * RB;                  // input row of the block's first step
        // one loop per (role, fast/edge) so that the rotation's register moves are not doubled by branch joins inside the loop
        auto wloop = [&](auto fast_c) {
#pragma unroll 1
          for (int rr = 0; rr < RB; ++rr) {
            const uint32_t ybase = sbase + (uint32_t)rr * ROW_BYTES;
            wstep(fast_c, std::integral_constant<int, 0>{}, ybase, ybase + PART_BYTES, tb + rr);
            // next step: the row that is j steps old must sit in window slot K-j
#pragma unroll
            for (int j = 1; j < K; ++j)
#pragma unroll
              for (int o = 0; o < SW; ++o) acc[j][o] = acc[(j + 1) % K][o];
          }
        };
        auto dloop = [&](auto fast_c) {
#pragma unroll 1
          for (int rr = 0; rr < RB; ++rr) {
            const uint32_t ybase = sbase + (uint32_t)rr * ROW_BYTES;
            dstep(fast_c, std::integral_constant<int, 0>{}, ybase, ybase + PART_BYTES, tb + rr);
            // slot q completes q steps ahead: everything moves one step closer (slot K-1 is re-initialised by the next step)
#pragma unroll
            for (int j = 0; j < K - 1; ++j)
#pragma unroll
              for (int o = 0; o < SW; ++o) acc[j][o] = acc[j + 1][o];
          }
        };
        if (wrole) {
          if (warp_fast && tb >= 0 && tb + RB <= p.H && tb + p.pt >= ci.r0 && tb + p.pt + RB <= r1) wloop(std::true_type{});
          else wloop(std::false_type{});
        } else {
          if (warp_fast && tb >= ci.r0 && tb + RB <= r1) dloop(std::true_type{});
          else dloop(std::false_type{});
        }
      } else if (wactive) {
#pragma unroll 1
        for (int rep = 0; rep < REP; ++rep) {
          const uint32_t sbase = ring + (uint32_t)s * SLOT_BYTES + (uint32_t)(rep * K) * ROW_BYTES + (uint32_t)gl * 4u;
          const uint32_t ybase = sbase + (uint32_t)((strip * SUBS + sub) * SW) * PIXB, gbase = ybase + PART_BYTES;
          const int t = t0 + blk * RB + rep * K;       // input row of step r = 0 of this group
          if (wrole) {
            const bool fast = warp_fast && t >= 0 && t + K <= p.H && t + p.pt >= ci.r0 && t + p.pt + K <= r1;
            if (fast) {
              wstep(std::true_type{}, std::integral_constant<int, 0>{}, ybase, gbase, t);
              wstep(std::true_type{}, std::integral_constant<int, 1>{}, ybase, gbase, t + 1);
              wstep(std::true_type{}, std::integral_constant<int, 2>{}, ybase, gbase, t + 2);
              if constexpr (K == 5) {
                wstep(std::true_type{}, std::integral_constant<int, 3>{}, ybase, gbase, t + 3);
                wstep(std::true_type{}, std::integral_constant<int, 4>{}, ybase, gbase, t + 4);
              }
            } else {
              wstep(std::false_type{}, std::integral_constant<int, 0>{}, ybase, gbase, t);
              wstep(std::false_type{}, std::integral_constant<int, 1>{}, ybase, gbase, t + 1);
              wstep(std::false_type{}, std::integral_constant<int, 2>{}, ybase, gbase, t + 2);
              if constexpr (K == 5) {
                wstep(std::false_type{}, std::integral_constant<int, 3>{}, ybase, gbase, t + 3);
                wstep(std::false_type{}, std::integral_constant<int, 4>{}, ybase, gbase, t + 4);
              }
            }
          } else {
            const bool fast = warp_fast && t >= ci.r0 && t + K <= r1;
            if (fast) {
              dstep(std::true_type{}, std::integral_constant<int, 0>{}, ybase, gbase, t);
              dstep(std::true_type{}, std::integral_constant<int, 1>{}, ybase, gbase, t + 1);
              dstep(std::true_type{}, std::integral_constant<int, 2>{}, ybase, gbase, t + 2);
              if constexpr (K == 5) {
                dstep(std::true_type{}, std::integral_constant<int, 3>{}, ybase, gbase, t + 3);
                dstep(std::true_type{}, std::integral_constant<int, 4>{}, ybase, gbase, t + 4);
              }
            } else {
              dstep(std::false_type{}, std::integral_constant<int, 0>{}, ybase, gbase, t);
              dstep(std::false_type{}, std::integral_constant<int, 1>{}, ybase, gbase, t + 1);
              dstep(std::false_type{}, std::integral_constant<int, 2>{}, ybase, gbase, t + 2);
              if constexpr (K == 5) {
                dstep(std::false_type{}, std::integral_constant<int, 3>{}, ybase, gbase, t + 3);
                dstep(std::false_type{}, std::integral_constant<int, 4>{}, ybase, gbase, t + 4);
              }
            }
          }
        }
      }
      __syncwarp();
      if (pi.item < my_items) {
        if (lane == 0 && (atom_add_acqrel_smem(&arrivals[s], 1u) % NW) == NW - 1) {
          fence_proxy_async_smem();
          issue(s);
        }
        advance();
      }
    }
  }
  // ---- flush: dW partials (wgrad warps) through the ring memory, BN partials (dgrad warps) ----
  __syncthreads();                                     // every TMA load has been consumed
  float* wred = reinterpret_cast<float*>(dws_smem);    // [NS][K*K][64]
  if (wrole) {
#pragma unroll
    for (int q = 0; q < K * K; ++q) *reinterpret_cast<float2*>(wred + ((size_t)strip * K * K + q) * 64 + lane * 2) = wv[q];
  } else {
    red[strip][0][lane] = bs.x; red[strip][1][lane] = bs.y; red[strip][2][lane] = bq.x; red[strip][3][lane] = bq.y;
  }
  __syncthreads();
  for (int i = threadIdx.x; i < K * K * CG; i += NS * 64) {
    const int t = i / CG, ch = i % CG;
    if (c0 + ch < p.C) {
      float s2 = 0.f;
#pragma unroll
      for (int w2 = 0; w2 < NS; ++w2)
#pragma unroll
        for (int g2 = 0; g2 < SUBS; ++g2) s2 += wred[((size_t)w2 * K * K + t) * 64 + g2 * CG + ch];      // lane l of a warp wrote floats 2l, 2l+1
      p.dw_part[((size_t)slot * K * K + t) * p.C + c0 + ch] = s2;
    }
  }
  if (BN && p.bn_part && warp == 0 && cvalid && sub == 0) {
    float a = 0.f, b = 0.f, cc = 0.f, d = 0.f;
#pragma unroll
    for (int w2 = 0; w2 < NS; ++w2)
#pragma unroll
      for (int g2 = 0; g2 < SUBS; ++g2) {
        const int l2 = g2 * LPG + gl;
        a += red[w2][0][l2]; b += red[w2][1][l2]; cc += red[w2][2][l2]; d += red[w2][3][l2];
      }
    float* st = p.bn_part + (size_t)slot * 2 * p.C;
    st[c] = a; st[c + 1] = b; st[p.C + c] = cc; st[p.C + c + 1] = d;
  }
}


// -------------------------------------------------------------------------------------------------------------------
// backward, stride 2.  Same two roles as stride 1, but the data gradient is a GATHER: input pixel (v, x) receives
//   dA = sum over ky = (v+pt) mod 2 (+2,+4), kx = (x+pl) mod 2 (+2,+4) of dY[(v+pt-ky)/2][(x+pl-kx)/2] * w[ky][kx]
// so the dgrad warp keeps a register window of the last NA = (K+1)/2 dY rows (4 + NA-1 pixels each) and produces the 8 input
// pixels of one row per step (then swish', store, BN sums).  Ownership is shifted by the pads so that every parity is a
// compile-time constant: a warp owns input columns 2*wx - pl + i (i < 8), an item owns input rows 2*r0 - pt + j (j < 2*rows);
// local step jj = j + 2(NA-1) >= 0 (the dgrad window needs NA-1 earlier dY rows), groups of G = 2*NA steps.
// Ring slot = G input rows of Y_in (33/35 px wide) + NA rows of dY (17/18 px wide).
// -------------------------------------------------------------------------------------------------------------------
template <int K, bool BN, bool ROT>
__global__ void __launch_bounds__(256, 2) mclip_dws_bwd_s2_kernel(const __grid_constant__ CUtensorMap tmIn, const __grid_constant__ CUtensorMap tmDy, const DwsDev p) {
  constexpr int SW = 4, NS = 4, NW = 8, TW = SW * NS, NA = (K + 1) / 2, G = 2 * NA, RB = G, NSLOT = 3;
  constexpr int IW = (TW - 1) * 2 + K, PC = (SW - 1) * 2 + K, IWG = TW + NA - 1, PCG = SW + NA - 1, OWN = 2 * SW;
  constexpr uint32_t ROWY = IW * 128, ROWG = IWG * 128, PARTY = RB * ROWY, SLOT_BYTES = PARTY + NA * ROWG;
  constexpr int JMIN = -2 * (NA - 1);
  extern __shared__ __align__(1024) uint8_t dws_smem[];
  __shared__ float red[NS][4][32];
  __shared__ __align__(8) uint64_t full[NSLOT];
  __shared__ uint32_t arrivals[NSLOT];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int strip = warp & 3;
  const bool wrole = warp < NS;
  const int chunk = blockIdx.x % p.n_chunks, slot = blockIdx.x / p.n_chunks;
  const int c0 = chunk * 64, c = c0 + lane * 2;
  const bool cvalid = c < p.C;
  if (threadIdx.x == 0) {
    tma_prefetch_desc(&tmIn); tma_prefetch_desc(&tmDy);
    for (int s = 0; s < NSLOT; ++s) { mbar_init(&full[s], 1); arrivals[s] = 0; }
    fence_mbar_init();
  }
  __syncthreads();
  const uint32_t ring = smem_u32(dws_smem);
  const int my_items = p.items > slot ? (p.items - slot + p.slots - 1) / p.slots : 0;
  const int per_img = p.strips_x * p.segs;
  const int rows_total = p.seg_rows * p.segs;      // >= max(Ho, ceil((H+pt)/2)): the row-pair grid (host plan)

  auto load_item = [&](DwsIter& it) {
    const int g = slot + it.item * p.slots;
    it.n = g / per_img;
    const int rem = g - it.n * per_img;
    const int sy = rem / p.strips_x;
    it.x0 = (rem - sy * p.strips_x) * TW;
    it.r0 = sy * p.seg_rows;
    it.rows = min(rows_total, it.r0 + p.seg_rows) - it.r0;
    it.nblk = (2 * (it.rows - 1) + K + 2 * (NA - 1) + RB - 1) / RB;
    it.blk = 0;
  };
  DwsIter pi;
  pi.item = 0;
  auto advance = [&]() { if (++pi.blk == pi.nblk) { if (++pi.item < my_items) load_item(pi); } };
  auto issue = [&](int s) {
    uint8_t* dst = dws_smem + (size_t)s * SLOT_BYTES;
    mbar_expect_tx(&full[s], SLOT_BYTES);
    tma_load_4d(dst, &tmIn, &full[s], c0, 2 * pi.x0 - p.pl, 2 * pi.r0 - p.pt + JMIN + pi.blk * RB, pi.n);
    tma_load_4d(dst + PARTY, &tmDy, &full[s], c0, pi.x0 - (NA - 1), pi.r0 - (NA - 1) + pi.blk * NA, pi.n);
  };
  if (my_items > 0) {
    load_item(pi);
#pragma unroll 1
    for (int i = 0; i < NSLOT; ++i) {
      if (pi.item < my_items) {
        if (threadIdx.x == 0) issue(i);
        advance();
      }
    }
  }

  float2 wv[K * K];                                    // wgrad role: dW accumulators; dgrad role: the weights
#pragma unroll
  for (int t = 0; t < K * K; ++t)
    wv[t] = (!wrole && cvalid) ? make_float2(p.w[(size_t)c * K * K + t], p.w[(size_t)(c + 1) * K * K + t]) : make_float2(0.f, 0.f);
  const float f = BN ? 0.5f : 1.0f;
  float2 a2 = make_float2(f, f), b2 = make_float2(0.f, 0.f), mu_is = make_float2(1.f, 1.f), nmis = make_float2(0.f, 0.f);
  if (BN && cvalid) {
    a2 = make_float2(f * p.scale[c], f * p.scale[c + 1]); b2 = make_float2(f * p.shift[c], f * p.shift[c + 1]);
    mu_is = make_float2(p.invstd[c], p.invstd[c + 1]);
    nmis = make_float2(-p.mean[c] * mu_is.x, -p.mean[c + 1] * mu_is.y);
  }
  float2 bs = make_float2(0.f, 0.f), bq = make_float2(0.f, 0.f);
  const float2 one = make_float2(1.f, 1.f), half2 = make_float2(0.5f, 0.5f), two2 = make_float2(2.f, 2.f), neg1 = make_float2(-1.f, -1.f);
  const int cw = p.C >> 1;
  const unsigned rstride_b = (unsigned)(p.W * cw) * 4u, pix_b = (unsigned)p.C * 2u;
  DwsIter ci;
  ci.item = 0;
  int count = 0;
#pragma unroll 1
  for (; ci.item < my_items; ++ci.item) {
    load_item(ci);
    const int wx = ci.x0 + strip * SW;               // first dY column of this warp's strip
    const int xo = 2 * wx - p.pl;                    // first owned input column (may be negative)
    const bool wactive = xo < p.W;
    uint32_t inmask = 0, ownmask = 0;
#pragma unroll
    for (int ix = 0; ix < PC; ++ix) { const int gx = xo + ix; inmask |= (gx >= 0 && gx < p.W) ? (1u << ix) : 0u; }
#pragma unroll
    for (int i = 0; i < OWN; ++i) { const int gx = xo + i; ownmask |= (gx >= 0 && gx < p.W && cvalid) ? (1u << i) : 0u; }
    const bool edge = inmask != ((1u << PC) - 1u);
    const bool warp_fast = wrole ? !edge : (xo >= 0 && xo + OWN <= p.W);
    const int v0 = 2 * ci.r0 - p.pt;                 // input row of local step j = 0
    // dx word of (row v0, column xo) for this lane; rows / columns advance by rstride_b / pix_b
    char* const obase = reinterpret_cast<char*>(p.dx) + ((long long)ci.n * p.H * (long long)(p.W * cw) + (long long)(c >> 1)) * 4 +
                        (long long)v0 * (long long)rstride_b + (long long)xo * (long long)pix_b;
    // window of the last NA dY rows: the dgrad role keeps PCG pixels per row, the wgrad role SW (zero outside its segment)
    float2 win[NA][PCG];
#pragma unroll
    for (int a = 0; a < NA; ++a)
#pragma unroll
      for (int b = 0; b < PCG; ++b) win[a][b] = make_float2(0.f, 0.f);

    // slot of dY row (j - ky) / 2 with j = jj + JMIN, jj = s (mod G):  ((s - ky)/2 - (NA-1)) mod NA
    auto wstep = [&](auto fast_c, auto s_c, uint32_t ybase, uint32_t gbase, int j) {
      constexpr bool FAST = decltype(fast_c)::value;
      constexpr int s = decltype(s_c)::value;
      if (s % 2 == 0) {                              // dY row j/2 enters the window
        constexpr int qn = ((s / 2 - (NA - 1)) % NA + NA) % NA;
        const int orel = (j - (s % 2)) / 2;
        if (FAST || (j >= 0 && orel < ci.rows && ci.r0 + orel < p.Ho)) {
#pragma unroll
          for (int o = 0; o < SW; ++o) win[qn][o] = bf2_to_f2(lds32(gbase + (s / 2) * ROWG + (NA - 1 + o) * 128));
        } else {
#pragma unroll
          for (int o = 0; o < SW; ++o) win[qn][o] = make_float2(0.f, 0.f);
        }
      }
      if (FAST || (unsigned)(v0 + j) < (unsigned)p.H) {
        float2 x[PC];
#pragma unroll
        for (int ix = 0; ix < PC; ++ix) {
          float2 h = ffma2r(bf2_to_f2(lds32(ybase + s * ROWY + ix * 128)), a2, b2);
          if (BN) h = ffma2r(h, make_float2(fast_tanh(h.x), fast_tanh(h.y)), h);
          x[ix] = h;
        }
        if (!FAST && edge) {
#pragma unroll
          for (int ix = 0; ix < PC; ++ix)
            if (!((inmask >> ix) & 1u)) x[ix] = make_float2(0.f, 0.f);
        }
#pragma unroll
        for (int ky = s % 2; ky < K; ky += 2) {
          constexpr int dummy = 0; (void)dummy;
          const int q = (((s - ky) / 2 - (NA - 1)) % NA + 2 * NA) % NA;
#pragma unroll
          for (int kx = 0; kx < K; ++kx)
#pragma unroll
            for (int o = 0; o < SW; ++o) ffma2(wv[ky * K + kx], win[q][o], x[2 * o + kx]);
        }
      }
    };
    auto dstep = [&](auto fast_c, auto s_c, uint32_t ybase, uint32_t gbase, int j) {
      constexpr bool FAST = decltype(fast_c)::value;
      constexpr int s = decltype(s_c)::value;
      if (s % 2 == 0) {                              // dY row j/2 enters the window (TMA zero-fills outside dY)
        constexpr int qn = ((s / 2 - (NA - 1)) % NA + NA) % NA;
#pragma unroll
        for (int b = 0; b < PCG; ++b) win[qn][b] = bf2_to_f2(lds32(gbase + (s / 2) * ROWG + b * 128));
      }
      // owned input row v0 + j (owned: 0 <= j < 2*rows)
      if (FAST || (j >= 0 && j < 2 * ci.rows && (unsigned)(v0 + j) < (unsigned)p.H)) {
        float2 acc[OWN];
#pragma unroll
        for (int i = 0; i < OWN; ++i) {
          bool first = true;
#pragma unroll
          for (int ky = s % 2; ky < K; ky += 2) {
            const int q = (((s - ky) / 2 - (NA - 1)) % NA + 2 * NA) % NA;
#pragma unroll
            for (int kx = i % 2; kx < K; kx += 2) {
              const int b = (i - kx) / 2 + (NA - 1);   // exact: i - kx is even
              if (first) { acc[i] = fmul2(win[q][b], wv[ky * K + kx]); first = false; }
              else ffma2(acc[i], win[q][b], wv[ky * K + kx]);
            }
          }
        }
        char* op = obase + (long long)j * (long long)rstride_b;
#pragma unroll
        for (int i = 0; i < OWN; ++i) {
          const bool st = FAST ? cvalid : (((ownmask >> i) & 1u) != 0);
          float2 d = acc[i];
          if (BN) {
            const float2 yv = bf2_to_f2(lds32(ybase + s * ROWY + i * 128));
            const float2 hv = ffma2r(yv, a2, b2);
            const float2 sg = ffma2r(make_float2(fast_tanh(hv.x), fast_tanh(hv.y)), half2, half2);
            const float2 om = ffma2r(sg, neg1, one);
            const float2 qq = ffma2r(fmul2(hv, om), two2, one);
            d = fmul2(d, fmul2(sg, qq));
            if (!FAST && !st) d = make_float2(0.f, 0.f);
            ffma2(bs, d, one);
            ffma2(bq, d, ffma2r(yv, mu_is, nmis));
          }
          stg32_if(op + (unsigned)i * pix_b, pack_bf16(d.x, d.y), st);
        }
      }
    };
#pragma unroll 1
    for (int blk = 0; blk < ci.nblk; ++blk, ++count) {
      const int sl = count % NSLOT;
      mbar_wait(&full[sl], (uint32_t)(count / NSLOT) & 1u);
      if (wactive) {
        const uint32_t sbase = ring + (uint32_t)sl * SLOT_BYTES + (uint32_t)lane * 4u;
        const uint32_t ybase = sbase + (uint32_t)(strip * OWN) * 128u, gbase = sbase + PARTY + (uint32_t)(strip * SW) * 128u;
        const int j = JMIN + blk * RB;               // local step j of group position s = 0
        if constexpr (ROT) {
          // ROT: one PAIR of steps (s = 0, 1: the newest dY row always enters window slot 1, the row of age a sits in slot (1 - a) mod NA)
          // per loop trip and a register rotation of the window, instead of G = 2*NA statically unrolled steps per role: a third of
          // the code (the k5 kernel was 6440 SASS instructions for two roles)
          auto rotate = [&]() {
#pragma unroll
            for (int a = NA - 1; a >= 1; --a)
#pragma unroll
              for (int b = 0; b < PCG; ++b) win[(1 - a + NA) % NA][b] = win[(2 - a + NA) % NA][b];
          };
          auto wloop = [&](auto fast_c) {
#pragma unroll 1
            for (int pp = 0; pp < NA; ++pp) {
              const uint32_t yb = ybase + (uint32_t)(2 * pp) * ROWY, gb = gbase + (uint32_t)pp * ROWG;
              wstep(fast_c, std::integral_constant<int, 0>{}, yb, gb, j + 2 * pp);
              wstep(fast_c, std::integral_constant<int, 1>{}, yb, gb, j + 2 * pp + 1);
              rotate();
            }
          };
          auto dloop = [&](auto fast_c) {
#pragma unroll 1
            for (int pp = 0; pp < NA; ++pp) {
              const uint32_t yb = ybase + (uint32_t)(2 * pp) * ROWY, gb = gbase + (uint32_t)pp * ROWG;
              dstep(fast_c, std::integral_constant<int, 0>{}, yb, gb, j + 2 * pp);
              dstep(fast_c, std::integral_constant<int, 1>{}, yb, gb, j + 2 * pp + 1);
              rotate();
            }
          };
          if (wrole) {
            const bool fast = warp_fast && v0 + j >= 0 && v0 + j + G <= p.H && j >= 0 && (j + G - 2) / 2 < ci.rows && ci.r0 + (j + G - 2) / 2 < p.Ho;
            if (fast) wloop(std::true_type{}); else wloop(std::false_type{});
          } else {
            const bool fast = warp_fast && j >= 0 && j + G <= 2 * ci.rows && v0 + j >= 0 && v0 + j + G <= p.H;
            if (fast) dloop(std::true_type{}); else dloop(std::false_type{});
          }
        } else if (wrole) {
          // fast: every A row inside the image, every dY row entering the window owned by this segment and inside dY
          const bool fast = warp_fast && v0 + j >= 0 && v0 + j + G <= p.H && j >= 0 && (j + G - 2) / 2 < ci.rows && ci.r0 + (j + G - 2) / 2 < p.Ho;
          if (fast) {
            wstep(std::true_type{}, std::integral_constant<int, 0>{}, ybase, gbase, j);
            wstep(std::true_type{}, std::integral_constant<int, 1>{}, ybase, gbase, j + 1);
            wstep(std::true_type{}, std::integral_constant<int, 2>{}, ybase, gbase, j + 2);
            wstep(std::true_type{}, std::integral_constant<int, 3>{}, ybase, gbase, j + 3);
            if constexpr (G > 4) {
              wstep(std::true_type{}, std::integral_constant<int, 4>{}, ybase, gbase, j + 4);
              wstep(std::true_type{}, std::integral_constant<int, 5>{}, ybase, gbase, j + 5);
            }
          } else {
            wstep(std::false_type{}, std::integral_constant<int, 0>{}, ybase, gbase, j);
            wstep(std::false_type{}, std::integral_constant<int, 1>{}, ybase, gbase, j + 1);
            wstep(std::false_type{}, std::integral_constant<int, 2>{}, ybase, gbase, j + 2);
            wstep(std::false_type{}, std::integral_constant<int, 3>{}, ybase, gbase, j + 3);
            if constexpr (G > 4) {
              wstep(std::false_type{}, std::integral_constant<int, 4>{}, ybase, gbase, j + 4);
              wstep(std::false_type{}, std::integral_constant<int, 5>{}, ybase, gbase, j + 5);
            }
          }
        } else {
          const bool fast = warp_fast && j >= 0 && j + G <= 2 * ci.rows && v0 + j >= 0 && v0 + j + G <= p.H;
          if (fast) {
            dstep(std::true_type{}, std::integral_constant<int, 0>{}, ybase, gbase, j);
            dstep(std::true_type{}, std::integral_constant<int, 1>{}, ybase, gbase, j + 1);
            dstep(std::true_type{}, std::integral_constant<int, 2>{}, ybase, gbase, j + 2);
            dstep(std::true_type{}, std::integral_constant<int, 3>{}, ybase, gbase, j + 3);
            if constexpr (G > 4) {
              dstep(std::true_type{}, std::integral_constant<int, 4>{}, ybase, gbase, j + 4);
              dstep(std::true_type{}, std::integral_constant<int, 5>{}, ybase, gbase, j + 5);
            }
          } else {
            dstep(std::false_type{}, std::integral_constant<int, 0>{}, ybase, gbase, j);
            dstep(std::false_type{}, std::integral_constant<int, 1>{}, ybase, gbase, j + 1);
            dstep(std::false_type{}, std::integral_constant<int, 2>{}, ybase, gbase, j + 2);
            dstep(std::false_type{}, std::integral_constant<int, 3>{}, ybase, gbase, j + 3);
            if constexpr (G > 4) {
              dstep(std::false_type{}, std::integral_constant<int, 4>{}, ybase, gbase, j + 4);
              dstep(std::false_type{}, std::integral_constant<int, 5>{}, ybase, gbase, j + 5);
            }
          }
        }
      }
      __syncwarp();
      if (pi.item < my_items) {
        if (lane == 0 && (atom_add_acqrel_smem(&arrivals[sl], 1u) % NW) == NW - 1) {
          fence_proxy_async_smem();
          issue(sl);
        }
        advance();
      }
    }
  }
  __syncthreads();
  float* wred = reinterpret_cast<float*>(dws_smem);
  if (wrole) {
#pragma unroll
    for (int q = 0; q < K * K; ++q) *reinterpret_cast<float2*>(wred + ((size_t)strip * K * K + q) * 64 + lane * 2) = wv[q];
  } else {
    red[strip][0][lane] = bs.x; red[strip][1][lane] = bs.y; red[strip][2][lane] = bq.x; red[strip][3][lane] = bq.y;
  }
  __syncthreads();
  for (int i = threadIdx.x; i < K * K * 64; i += 256) {
    const int t = i / 64, ch = i % 64;
    if (c0 + ch < p.C) {
      float s2 = 0.f;
#pragma unroll
      for (int w2 = 0; w2 < NS; ++w2) s2 += wred[((size_t)w2 * K * K + t) * 64 + ch];
      p.dw_part[((size_t)slot * K * K + t) * p.C + c0 + ch] = s2;
    }
  }
  if (BN && p.bn_part && warp == 0 && cvalid) {
    float a = 0.f, b = 0.f, cc = 0.f, d = 0.f;
#pragma unroll
    for (int w2 = 0; w2 < NS; ++w2) { a += red[w2][0][lane]; b += red[w2][1][lane]; cc += red[w2][2][lane]; d += red[w2][3][lane]; }
    float* st = p.bn_part + (size_t)slot * 2 * p.C;
    st[c] = a; st[c + 1] = b; st[p.C + c] = cc; st[p.C + c + 1] = d;
  }
}

// 4-D tensor map over an NHWC bf16 tensor: dims {C, W, H, N}, box {64 channels, box_w, box_h, 1}, no swizzle.
int dws_tmap(CUtensorMap* m, const void* ptr, int N, int H, int W, int C, int box_w, int box_h, int cg = 64) {
  const unsigned long long dims[4] = {(unsigned long long)C, (unsigned long long)W, (unsigned long long)H, (unsigned long long)N};
  const unsigned long long strides[3] = {(unsigned long long)C * 2, (unsigned long long)W * C * 2, (unsigned long long)H * W * C * 2};
  const unsigned box[4] = {(unsigned)cg, (unsigned)box_w, (unsigned)box_h, 1u};
#ifndef DWS_L2PROMO
#define DWS_L2PROMO 2
#endif
  return mclip_tmap_encode_bf16(m, ptr, 4, dims, strides, box, DWS_L2PROMO);
}

// work decomposition shared by mclip_dws_slots and the launchers
void dws_plan(const mclip_dwconv_args* a, int ctas_per_sm, int rows_total, DwsDev& p, int TW = 16, int cg = 64) {
  p.n_chunks = ceil_div(a->c, cg);
  p.strips_x = ceil_div(a->wo, TW);
  int ctas = (mclip_num_sms() * ctas_per_sm) / p.n_chunks;
  if (ctas < 1) ctas = 1;
  // split the strip of `rows_total` rows into segments until every CTA gets >= ~6 items (segment overhead: K-1 rows)
  const long long cols = (long long)a->n * p.strips_x;
  int segs = 1;
  while (cols * segs < 6LL * ctas && rows_total / (segs * 2) >= 24) segs *= 2;
  p.seg_rows = ceil_div(rows_total, segs);
  p.segs = ceil_div(rows_total, p.seg_rows);
  p.items = (int)(cols * p.segs);
  p.slots = std::min(ctas, p.items);
}

// 16-channel lane groups: the narrow k3 s1 layers (EN-B5 blocks 0-2: C = 48, 24, 24; EN-B2 blocks 0-1: 32, 16)
// (measured, EN-B5 blocks 0-2 at B = 64: C = 24 forward 1.24 -> 0.77 ms, backward 2.55 -> 1.63 ms; C = 48 backward 2.56 -> 2.33 ms but
// forward 1.24 -> 1.49 ms: three 16-channel chunks cost more than one 64-channel chunk with a quarter of its lanes idle)
static int dws_cg(const mclip_dwconv_args* a, bool backward) {
  static int enabled = -1;
  if (enabled < 0) { const char* e = getenv("MCLIP_DW_CG16"); enabled = e ? atoi(e) : 1; }
  return (enabled && a->k == 3 && a->stride == 1 && a->c <= (backward ? 48 : 32) && a->c % 2 == 0) ? 16 : 64;
}

template <int K, int S, int CG = 64>
int dws_launch_fwd(const mclip_dwconv_args* a, DwsDev& p, cudaStream_t stream) {
  using Cfg = FwdCfg<K, S, CG>;
  CUtensorMap tm;
  int rc = dws_tmap(&tm, p.in, p.N, p.H, p.W, p.C, Cfg::IW, Cfg::RB, CG);
  if (rc) return rc;
  auto kern = p.act ? mclip_dws_fwd_kernel<K, S, true, CG> : mclip_dws_fwd_kernel<K, S, false, CG>;
  static bool attr[2] = {false, false};
  if (!attr[p.act]) { MCLIP_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM)); attr[p.act] = true; }
  kern<<<p.n_chunks * p.slots, 128, Cfg::SMEM, stream>>>(tm, p);
  MCLIP_CHECK_LAUNCH();
  return MCLIP_OK;
}

static int dws_fwd_ctas(const mclip_dwconv_args* a) {
  if (dws_cg(a, false) == 16) return FwdCfg<3, 1, 16>::CTAS;
  if (a->stride == 1) return a->k == 3 ? FwdCfg<3, 1>::CTAS : FwdCfg<5, 1>::CTAS;
  return a->k == 3 ? FwdCfg<3, 2>::CTAS : FwdCfg<5, 2>::CTAS;
}

#ifndef DWS_BWD_K5_NS
#define DWS_BWD_K5_NS 2
#endif
#ifndef DWS_BWD_K3_ROT
#define DWS_BWD_K3_ROT 0
#endif
#ifndef DWS_BWD_K5_ROT
#define DWS_BWD_K5_ROT 1
#endif
#ifndef DWS_BWD2_K5_ROT
#define DWS_BWD2_K5_ROT 1
#endif
template <int K, int CG = 64>
struct BwdCfg {
  static constexpr int NSLOT = DWS_BWD_NSLOT;
  static constexpr int REP = (K == 3) ? 2 : 1;
  static constexpr int NS = (K == 3) ? 4 : DWS_BWD_K5_NS;      // column strips per CTA (k5: 2 -> 128 threads, 3 CTAs/SM at 168 registers, no spills)
  static constexpr int CTAS = NS == 4 ? 2 : 3;
  static constexpr bool ROT = (K == 3) ? (DWS_BWD_K3_ROT != 0) : (DWS_BWD_K5_ROT != 0);
  static constexpr int TW = (CG == 64 ? 4 : 5) * (64 / CG) * NS, IW = TW + K - 1;
  static constexpr int RING = NSLOT * 2 * K * REP * IW * CG * 2;
  static constexpr int SMEM = RING > NS * K * K * 64 * 4 ? RING : NS * K * K * 64 * 4;
};

template <int K, int CG = 64>
int dws_launch_bwd_s1(const mclip_dwconv_args* a, DwsDev& p, cudaStream_t stream) {
  using Cfg = BwdCfg<K, CG>;
  CUtensorMap tmIn, tmDy;
  int rc = dws_tmap(&tmIn, p.in, p.N, p.H, p.W, p.C, Cfg::IW, K * Cfg::REP, CG);
  if (rc) return rc;
  if ((rc = dws_tmap(&tmDy, p.dy, p.N, p.Ho, p.Wo, p.C, Cfg::IW, K * Cfg::REP, CG))) return rc;
  const bool bn = p.scale != nullptr;
  auto kern = bn ? mclip_dws_bwd_s1_kernel<K, Cfg::NSLOT, Cfg::REP, true, Cfg::NS, Cfg::ROT, CG>
                 : mclip_dws_bwd_s1_kernel<K, Cfg::NSLOT, Cfg::REP, false, Cfg::NS, Cfg::ROT, CG>;
  static bool attr[2] = {false, false};
  if (!attr[bn]) { MCLIP_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM)); attr[bn] = true; }
  kern<<<p.n_chunks * p.slots, Cfg::NS * 64, Cfg::SMEM, stream>>>(tmIn, tmDy, p);
  MCLIP_CHECK_LAUNCH();
  return MCLIP_OK;
}

template <int K>
struct Bwd2Cfg {
  static constexpr int NA = (K + 1) / 2, G = 2 * NA, IW = 30 + K, IWG = 16 + NA - 1;
  static constexpr int RING = 3 * (G * IW + NA * IWG) * 128;
  static constexpr int SMEM = RING > 4 * K * K * 64 * 4 ? RING : 4 * K * K * 64 * 4;
};

template <int K>
int dws_launch_bwd_s2(const mclip_dwconv_args* a, DwsDev& p, cudaStream_t stream) {
  using Cfg = Bwd2Cfg<K>;
  CUtensorMap tmIn, tmDy;
  int rc = dws_tmap(&tmIn, p.in, p.N, p.H, p.W, p.C, Cfg::IW, Cfg::G);
  if (rc) return rc;
  if ((rc = dws_tmap(&tmDy, p.dy, p.N, p.Ho, p.Wo, p.C, Cfg::IWG, Cfg::NA))) return rc;
  const bool bn = p.scale != nullptr;
  constexpr bool ROT = K == 5 && DWS_BWD2_K5_ROT != 0;
  auto kern = bn ? mclip_dws_bwd_s2_kernel<K, true, ROT> : mclip_dws_bwd_s2_kernel<K, false, ROT>;
  static bool attr[2] = {false, false};
  if (!attr[bn]) { MCLIP_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM)); attr[bn] = true; }
  kern<<<p.n_chunks * p.slots, 256, Cfg::SMEM, stream>>>(tmIn, tmDy, p);
  MCLIP_CHECK_LAUNCH();
  return MCLIP_OK;
}

static void dws_plan_bwd1(const mclip_dwconv_args* a, DwsDev& p) {
  if (dws_cg(a, true) == 16) dws_plan(a, BwdCfg<3, 16>::CTAS, a->h, p, BwdCfg<3, 16>::TW, 16);
  else if (a->k == 3) dws_plan(a, BwdCfg<3>::CTAS, a->h, p, BwdCfg<3>::TW);
  else dws_plan(a, BwdCfg<5>::CTAS, a->h, p, BwdCfg<5>::TW);
}

// stride-2 backward: the grid of (dY row, dY column) PAIRS must also cover every input pixel, with ownership shifted by the pads
static void dws_plan_bwd2(const mclip_dwconv_args* a, DwsDev& p) {
  const int rows = std::max(a->ho, ceil_div(a->h + a->pad_top, 2)), cols = std::max(a->wo, ceil_div(a->w + a->pad_left, 2));
  mclip_dwconv_args t = *a;
  t.wo = cols;
  dws_plan(&t, 2, rows, p);
}

void dws_fill(const mclip_dwconv_args* a, DwsDev& p) {
  memset(&p, 0, sizeof(p));
  p.N = a->n; p.H = a->h; p.W = a->w; p.C = a->c; p.Ho = a->ho; p.Wo = a->wo; p.pl = a->pad_left; p.pt = a->pad_top;
  p.in = (const bf16*)a->in; p.scale = a->in_scale; p.shift = a->in_shift; p.act = a->in_scale ? a->in_act : 0; p.w = a->weight;
}

}  // namespace

// Which shapes the streaming kernels cover (the rest stays on conv.cu): stride 1 forward for now.
bool mclip_dws_covers(const mclip_dwconv_args* a, int backward) {
  static int enabled = -1;
  if (enabled < 0) { const char* e = getenv("MCLIP_DW_STREAM"); enabled = e ? atoi(e) : 15; }     // bit 0: forward, bit 1: backward (stride 1), bit 2: stride-2 forward, bit 3: stride-2 backward
  if (!enabled) return false;
  if ((a->stride != 1 && a->stride != 2) || (a->k != 3 && a->k != 5)) return false;
  if (backward && a->stride == 2) return (enabled & 8) != 0 && (a->in_scale == nullptr || a->in_act == 1);
  if (backward) return (enabled & 2) != 0 && a->ho == a->h && a->wo == a->w && (a->in_scale == nullptr || a->in_act == 1);
  return (enabled & 1) != 0 && (a->stride == 1 || (enabled & 4) != 0);
}

int mclip_dws_slots(const mclip_dwconv_args* a, int backward) {
  DwsDev p;
  dws_fill(a, p);
  if (backward && a->stride == 2) dws_plan_bwd2(a, p);
  else if (backward) dws_plan_bwd1(a, p);
  else if (dws_cg(a, false) == 16) dws_plan(a, dws_fwd_ctas(a), a->ho, p, FwdCfg<3, 1, 16>::TW, 16);
  else dws_plan(a, dws_fwd_ctas(a), a->ho, p);
  return p.slots;
}

// dW[c, t] = sum_slots part[slot][t][c]      (fixed order)
__global__ void mclip_dws_wgrad_reduce_kernel(const float* __restrict__ part, float* __restrict__ dw, int slots, int KK, int C, int accumulate) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= KK * C) return;
  const int t = i / C, c = i % C;
  float s = 0.f;
  for (int k = 0; k < slots; ++k) s += part[((size_t)k * KK + t) * C + c];
  float* o = dw + (size_t)c * KK + t;
  *o = accumulate ? *o + s : s;
}

int mclip_dws_backward(const mclip_dwconv_args* a, void* stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  DwsDev p;
  dws_fill(a, p);
  if (a->stride == 2) dws_plan_bwd2(a, p); else dws_plan_bwd1(a, p);
  MCLIP_REQUIRE(a->stat_slots == p.slots, "mclip_dwconv_backward: stat_slots=%d, expected %d", a->stat_slots, p.slots);
  p.dy = (const bf16*)a->dy; p.dx = (bf16*)a->dx; p.dw_part = a->dw_partials;
  p.bn_part = a->in_scale ? a->bn_partials : nullptr; p.mean = a->in_mean; p.invstd = a->in_invstd;
  if (p.bn_part) MCLIP_REQUIRE(p.mean && p.invstd, "mclip_dwconv_backward: input BN statistics missing");
  int rc;
  if (a->stride == 2) rc = a->k == 3 ? dws_launch_bwd_s2<3>(a, p, stream) : dws_launch_bwd_s2<5>(a, p, stream);
  else if (dws_cg(a, true) == 16) rc = dws_launch_bwd_s1<3, 16>(a, p, stream);
  else rc = a->k == 3 ? dws_launch_bwd_s1<3>(a, p, stream) : dws_launch_bwd_s1<5>(a, p, stream);
  if (rc) return rc;
  const int KK = a->k * a->k;
  mclip_dws_wgrad_reduce_kernel<<<ceil_div(KK * a->c, 256), 256, 0, stream>>>(a->dw_partials, a->dweight, p.slots, KK, a->c, a->accumulate);
  MCLIP_CHECK_LAUNCH();
  return MCLIP_OK;
}

int mclip_dws_forward(const mclip_dwconv_args* a, void* stream_) {
  DwsDev p;
  dws_fill(a, p);
  const int cg = dws_cg(a, false);
  if (cg == 16) dws_plan(a, dws_fwd_ctas(a), a->ho, p, FwdCfg<3, 1, 16>::TW, 16);
  else dws_plan(a, dws_fwd_ctas(a), a->ho, p);
  p.out = (bf16*)a->out; p.stats = a->stats;
  if (a->stats) MCLIP_REQUIRE(a->stat_slots == p.slots, "mclip_dwconv_forward: stat_slots=%d, expected %d", a->stat_slots, p.slots);
  cudaStream_t st = (cudaStream_t)stream_;
  if (cg == 16) return dws_launch_fwd<3, 1, 16>(a, p, st);
  if (a->stride == 1) return a->k == 3 ? dws_launch_fwd<3, 1>(a, p, st) : dws_launch_fwd<5, 1>(a, p, st);
  return a->k == 3 ? dws_launch_fwd<3, 2>(a, p, st) : dws_launch_fwd<5, 2>(a, p, st);
}
