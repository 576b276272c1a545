// BatchNorm (train + eval), swish, squeeze-excite, pooling, residual / drop-connect: the HBM-bound passes between the
// convolutions of MBConvBlock (efficientnet_custom.py:91-132) and of the stem / head (:273,283,309-312), forward and
// backward.  All of them stream NHWC bf16 tensors with 16-byte vectors; a thread owns one 8-channel vector for the whole
// kernel (per-channel parameters and accumulators live in registers) and strides over pixels.
//
//   reference op (file:line)                               kernel here
//   nn.BatchNorm2d train: batch stats, running update      producer epilogue partials -> mclip_bn_finalize
//     (efficientnet_custom.py:53-54,64,74,88,177,205)
//   bn(x) -> swish -> adaptive_avg_pool2d (:110-115)       mclip_act_pool   (also writes the activated tensor U)
//   _se_reduce -> swish -> _se_expand -> sigmoid (:116-119) mclip_se_fc      (+ mclip_se_scale_weights: gate folded into W_proj)
//   _bn2, drop_connect, x + inputs (:123-131)               mclip_bn_apply
//   swish -> _avg_pooling -> flatten -> dropout (:283,309-312)  mclip_act_pool + mclip_pool_finalize
//   autograd of all of the above                            mclip_bn_bwd_reduce / _finalize / _apply, mclip_se_bwd_pass1, mclip_se_fc_bwd
#include "common.cuh"
#include "mclip_internal.h"
#include <stdlib.h>

#define EW_MAX_THREADS 256
#define EW_CPT 4            // channels per thread (8-byte vectors; a warp still covers 256 contiguous bytes)
#define EW_UNR 4            // pixels in flight per thread
#define EW_ASYNC_DEFAULT 15  // see mclip_ew_backward: which passes stage their inputs through the cp.async ring

typedef unsigned long long u64;
__device__ __forceinline__ float2 bf2_to_f2(uint32_t u) { return make_float2(__uint_as_float(u << 16), __uint_as_float(u & 0xffff0000u)); }
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
__device__ __forceinline__ uint2 ldg_b64(const bf16* p) {
  uint2 r;
  asm volatile("ld.global.nc.L1::no_allocate.v2.u32 {%0,%1}, [%2];" : "=r"(r.x), "=r"(r.y) : "l"(p));
  return r;
}
__device__ __forceinline__ void stg_b64(bf16* p, uint2 v) {
  asm volatile("st.global.L1::no_allocate.v2.u32 [%0], {%1,%2};" ::"l"(p), "r"(v.x), "r"(v.y) : "memory");
}
// ---- per-thread cp.async ring: the bytes in flight live in shared memory instead of registers, so a register-heavy pass
//      (SE pass 1: 5 accumulators + 10 per-channel constants) still keeps ~100 KB per SM outstanding towards HBM.
//      A thread only ever reads the slots it filled itself: cp.async.wait_group is the only synchronisation needed.
#define EW_RING 8            // slots per thread (EW_RING-1 pixels in flight); slot = 16 B: y (8 B) | dU or residual (8 B)
__device__ __forceinline__ void cp_async8(uint32_t dst, const void* src) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(dst), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }
__device__ __forceinline__ uint4 lds128(uint32_t addr) {
  uint4 v;
  asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(addr));
  return v;
}

// sigma(t) via one MUFU; returns (sigma(t.x), sigma(t.y))
__device__ __forceinline__ float2 sigmoid2(const float2& t) {
  const float2 h = make_float2(fast_tanh(0.5f * t.x), fast_tanh(0.5f * t.y));
  return ffma2r(h, make_float2(0.5f, 0.5f), make_float2(0.5f, 0.5f));
}

// Which streaming passes stage their inputs through the cp.async ring: bit m = backward mode m, bit 3 = forward pass.
// Initialised from MCLIP_EW_ASYNC (default EW_ASYNC_DEFAULT); mclip_set_ew_async() overrides it (tests, tuning).
static int g_ew_async = -1;
static int ew_async_mask() {
  if (g_ew_async < 0) { const char* e = getenv("MCLIP_EW_ASYNC"); g_ew_async = e ? atoi(e) : EW_ASYNC_DEFAULT; }
  return g_ew_async;
}
extern "C" int mclip_set_ew_async(int mask) { const int old = ew_async_mask(); g_ew_async = mask & 15; return old; }

struct EwGeom {
  int N, HW, C, G, CG, TPP, PL, chunks, pix_per_chunk;
};

// Channel groups of <= 1024 channels (256 threads x 4); a thread keeps its 4 channels for the whole kernel.
static int ew_geom(int n, int hw, int c, EwGeom* g, int* threads) {
  if (c % 8 != 0) { mclip_set_error("elementwise: C=%d must be a multiple of 8", c); return MCLIP_ERR_INVALID; }
  int G = (c + EW_MAX_THREADS * EW_CPT - 1) / (EW_MAX_THREADS * EW_CPT);
  while (c % (G * EW_CPT) != 0) ++G;
  g->N = n; g->HW = hw; g->C = c; g->G = G; g->CG = c / G; g->TPP = g->CG / EW_CPT;
  g->PL = EW_MAX_THREADS / g->TPP; if (g->PL < 1) g->PL = 1;
  *threads = g->TPP * g->PL;
  long long want = (long long)mclip_num_sms() * 16;
  int chunks = (int)((want + (long long)n * G - 1) / ((long long)n * G));
  int max_chunks = (hw + g->PL * EW_UNR - 1) / (g->PL * EW_UNR);
  if (chunks > max_chunks) chunks = max_chunks;
  if (chunks < 1) chunks = 1;
  g->pix_per_chunk = (hw + chunks - 1) / chunks;
  g->chunks = (hw + g->pix_per_chunk - 1) / g->pix_per_chunk;
  return MCLIP_OK;
}

extern "C" int mclip_ew_chunks(int n, int hw, int c) {
  EwGeom g; int t;
  if (ew_geom(n, hw, c, &g, &t)) return -1;
  return g.chunks;
}

// block reduction over the PL pixel lanes: dst[cv*4+i] = sum over lanes of v[i]   (smem: [PL][CG])
__device__ __forceinline__ void ew_block_reduce4(float* smem, const float2* v, int cv, int pl, int CG, int PL, float* dst) {
  __syncthreads();
  *reinterpret_cast<float4*>(smem + (size_t)pl * CG + cv * 4) = make_float4(v[0].x, v[0].y, v[1].x, v[1].y);
  __syncthreads();
  for (int i = threadIdx.x; i < CG; i += blockDim.x) {
    float s = 0.f;
    for (int l = 0; l < PL; ++l) s += smem[(size_t)l * CG + i];
    dst[i] = s;
  }
}

// ------------------------------------------------------------------------------------------------
// BatchNorm statistics -> affine (a = gamma*invstd, b = beta - mean*a), running-stat update
// ------------------------------------------------------------------------------------------------
// 1024 threads = 8 channels x 128 slot lanes (a warp reads 4 slot rows x 32 B): narrow layers still get C/8 CTAs and a lane
// strides over at most ~20 of the up to N*chunks ~ 2400 partial slots; fp64 combine in smem, fixed order
#define BNF_CH 8
#define BNF_LANES 128
__device__ __forceinline__ void bn_reduce_slots(const float* __restrict__ partials, int slots, int C, int c, int lane, double& s, double& q,
                                                double (*sm)[BNF_LANES][BNF_CH]) {
  s = 0.0; q = 0.0;
  if (c < C)
    for (int k = lane; k < slots; k += BNF_LANES) { s += (double)partials[((size_t)k * 2 + 0) * C + c]; q += (double)partials[((size_t)k * 2 + 1) * C + c]; }
  const int cl = threadIdx.x % BNF_CH;
  sm[0][lane][cl] = s; sm[1][lane][cl] = q;
  __syncthreads();
  if (lane == 0) {
    s = 0.0; q = 0.0;
#pragma unroll
    for (int l = 0; l < BNF_LANES; ++l) { s += sm[0][l][cl]; q += sm[1][l][cl]; }      // fixed order => deterministic
  }
}

__global__ void __launch_bounds__(BNF_CH * BNF_LANES) mclip_bn_finalize_kernel(
    const float* __restrict__ partials, int slots, int C, double count, const float* __restrict__ gamma, const float* __restrict__ beta,
    float* running_mean, float* running_var, long long* num_batches, float momentum, float eps, int training, float* scale, float* shift,
    float* mean_out, float* invstd_out) {
  __shared__ double sm[2][BNF_LANES][BNF_CH];
  const int c = blockIdx.x * BNF_CH + threadIdx.x % BNF_CH, lane = threadIdx.x / BNF_CH;
  if (blockIdx.x == 0 && threadIdx.x == 0 && training && num_batches) *num_batches += 1;
  float mean, invstd;
  if (training) {
    double s, q;
    bn_reduce_slots(partials, slots, C, c, lane, s, q, sm);
    if (lane != 0 || c >= C) return;
    double m = s / count;
    double var = q / count - m * m;
    if (var < 0.0) var = 0.0;
    mean = (float)m;
    invstd = (float)(1.0 / sqrt(var + (double)eps));
    if (running_mean) {
      double unbiased = count > 1.0 ? var * count / (count - 1.0) : var;
      running_mean[c] = (1.f - momentum) * running_mean[c] + momentum * mean;
      running_var[c] = (1.f - momentum) * running_var[c] + momentum * (float)unbiased;
    }
  } else {
    if (lane != 0 || c >= C) return;
    mean = running_mean[c];
    invstd = rsqrtf(running_var[c] + eps);
  }
  const float a = gamma[c] * invstd;
  scale[c] = a;
  shift[c] = beta[c] - mean * a;
  mean_out[c] = mean;
  invstd_out[c] = invstd;
}

extern "C" int mclip_bn_finalize(const mclip_bn_args* a, void* stream) {
  MCLIP_REQUIRE(a && a->gamma && a->beta && a->scale && a->shift && a->mean && a->invstd, "mclip_bn_finalize: null operand");
  MCLIP_REQUIRE(a->training ? (a->partials != nullptr && a->slots > 0 && a->count > 0) : (a->running_mean && a->running_var),
                "mclip_bn_finalize: %s", a->training ? "training needs partials and a positive count" : "eval needs running statistics");
  mclip_bn_finalize_kernel<<<ceil_div(a->c, BNF_CH), BNF_CH * BNF_LANES, 0, (cudaStream_t)stream>>>(a->partials, a->slots, a->c, (double)a->count, a->gamma, a->beta,
                                                                               a->running_mean, a->running_var, a->num_batches_tracked,
                                                                               a->momentum, a->eps, a->training, a->scale, a->shift, a->mean, a->invstd);
  MCLIP_CHECK_LAUNCH();
  return MCLIP_OK;
}

// ------------------------------------------------------------------------------------------------
// generic streaming pass (forward):  v = a*y+b ; u = act ? swish(v) : v ; u *= rowscale[n] ; u += residual
//   optional: write u (bf16), accumulate per-(n,chunk) channel sums of u (pooling partials)
// ------------------------------------------------------------------------------------------------
struct EwFwdDev {
  EwGeom g;
  const bf16* y; const float* scale; const float* shift; int act;
  const float* rowscale; const bf16* residual;
  bf16* out; float* pool_part;     // pool_part: [N][chunks][C]
};

template <bool ASYNC>
__global__ void __launch_bounds__(EW_MAX_THREADS, 4) mclip_ew_fwd_kernel(const EwFwdDev p) {
  extern __shared__ float ew_smem[];
  const EwGeom& g = p.g;
  const int cv = threadIdx.x % g.TPP, pl = threadIdx.x / g.TPP;
  const int grp = blockIdx.x % g.G, bc = blockIdx.x / g.G;
  const int n = bc / g.chunks, chunk = bc % g.chunks;
  const int p0 = chunk * g.pix_per_chunk, p1 = min(g.HW, p0 + g.pix_per_chunk);
  const int c = grp * g.CG + cv * EW_CPT;
  const float f = p.act ? 0.5f : 1.0f;                    // swish(t) = h + h*tanh(h), h = t/2
  float2 a[2], b[2], acc[2];
#pragma unroll
  for (int i = 0; i < 2; ++i) {
    a[i] = p.scale ? make_float2(f * p.scale[c + 2 * i], f * p.scale[c + 2 * i + 1]) : make_float2(f, f);
    b[i] = p.shift ? make_float2(f * p.shift[c + 2 * i], f * p.shift[c + 2 * i + 1]) : make_float2(0.f, 0.f);
    acc[i] = make_float2(0.f, 0.f);
  }
  const float rsv = p.rowscale ? p.rowscale[n] : 1.f;
  const float2 rs = make_float2(rsv, rsv), one = make_float2(1.f, 1.f);
  const size_t base = (size_t)n * g.HW * g.C + c;
  auto pixel = [&](int px, const uint32_t (&yw)[2], const uint32_t (&rw)[2]) {
    uint32_t ow[2];
#pragma unroll
    for (int i = 0; i < 2; ++i) {
      float2 h = ffma2r(bf2_to_f2(yw[i]), a[i], b[i]);
      if (p.act) h = ffma2r(h, make_float2(fast_tanh(h.x), fast_tanh(h.y)), h);
      if (p.rowscale) h = fmul2(h, rs);
      if (p.residual) h = ffma2r(bf2_to_f2(rw[i]), one, h);
      ow[i] = pack_bf16(h.x, h.y);
      if (p.pool_part) ffma2(acc[i], bf2_to_f2(ow[i]), one);     // pool what the consumer reads (bf16-rounded)
    }
    if (p.out) stg_b64(p.out + base + (size_t)px * g.C, make_uint2(ow[0], ow[1]));
  };
  if (ASYNC) {
    const uint32_t slot_stride = blockDim.x * 16u;
    const uint32_t my = smem_u32(ew_smem) + (uint32_t)(g.PL * g.CG) * 4u + threadIdx.x * 16u;
    const bool has_res = p.residual != nullptr;
    int px_issue = p0 + pl;
    auto issue = [&](int slot) {
      if (px_issue < p1) {
        const size_t off = base + (size_t)px_issue * g.C;
        cp_async8(my + slot * slot_stride, p.y + off);
        if (has_res) cp_async8(my + slot * slot_stride + 8u, p.residual + off);
      }
      cp_async_commit();
      px_issue += g.PL;
    };
#pragma unroll
    for (int s = 0; s < EW_RING - 1; ++s) issue(s);
    int slot = 0;
    for (int px = p0 + pl; px < p1; px += g.PL) {
      issue(slot == 0 ? EW_RING - 1 : slot - 1);
      cp_async_wait<EW_RING - 1>();
      const uint4 v = lds128(my + slot * slot_stride);
      const uint32_t yw[2] = {v.x, v.y}, rw[2] = {v.z, v.w};
      pixel(px, yw, rw);
      slot = slot == EW_RING - 1 ? 0 : slot + 1;
    }
    cp_async_wait<0>();
  } else {
    for (int px0 = p0 + pl; px0 < p1; px0 += g.PL * EW_UNR) {
      uint2 yv[EW_UNR], rv[EW_UNR];
#pragma unroll
      for (int u = 0; u < EW_UNR; ++u) {
        const int px = px0 + u * g.PL;
        if (px < p1) {
          const size_t off = base + (size_t)px * g.C;
          yv[u] = ldg_b64(p.y + off);
          if (p.residual) rv[u] = ldg_b64(p.residual + off);
        }
      }
#pragma unroll
      for (int u = 0; u < EW_UNR; ++u) {
        const int px = px0 + u * g.PL;
        if (px >= p1) break;
        const uint32_t yw[2] = {yv[u].x, yv[u].y}, rw[2] = {rv[u].x, rv[u].y};
        pixel(px, yw, rw);
      }
    }
  }
  if (p.pool_part) ew_block_reduce4(ew_smem, acc, cv, pl, g.CG, g.PL, p.pool_part + ((size_t)n * g.chunks + chunk) * g.C + grp * g.CG);
}

extern "C" int mclip_ew_forward(const mclip_ew_args* a, void* stream) {
  MCLIP_REQUIRE(a && a->y, "mclip_ew_forward: null input");
  EwFwdDev p; int threads;
  int rc = ew_geom(a->n, a->hw, a->c, &p.g, &threads);
  if (rc) return rc;
  if (a->pool_partials) MCLIP_REQUIRE(a->chunks == p.g.chunks, "mclip_ew_forward: chunks=%d, expected %d", a->chunks, p.g.chunks);
  p.y = (const bf16*)a->y; p.scale = a->scale; p.shift = a->shift; p.act = a->act; p.rowscale = a->rowscale;
  p.residual = (const bf16*)a->residual; p.out = (bf16*)a->out; p.pool_part = a->pool_partials;
  const int smem = p.g.PL * p.g.CG * 4;
  static int carveout_set = 0;
  if (!carveout_set) {
    MCLIP_CHECK_CUDA(cudaFuncSetAttribute(mclip_ew_fwd_kernel<true>, cudaFuncAttributePreferredSharedMemoryCarveout, 70));
    carveout_set = 1;
  }
  if ((ew_async_mask() >> 3) & 1) mclip_ew_fwd_kernel<true><<<a->n * p.g.chunks * p.g.G, threads, smem + EW_RING * threads * 16, (cudaStream_t)stream>>>(p);
  else mclip_ew_fwd_kernel<false><<<a->n * p.g.chunks * p.g.G, threads, smem, (cudaStream_t)stream>>>(p);
  MCLIP_CHECK_LAUNCH();
  return MCLIP_OK;
}

// pooled[n,c] = mult[n,c] * (sum_chunks part[n,chunk,c]) / HW          (head average pool + dropout mask)
__global__ void mclip_pool_finalize_kernel(const float* __restrict__ part, int N, int chunks, int C, float inv_hw, const float* __restrict__ mult,
                                           float* __restrict__ out) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= N * C) return;
  const int n = i / C, c = i % C;
  float s = 0.f;
  for (int k = 0; k < chunks; ++k) s += part[((size_t)n * chunks + k) * C + c];
  s *= inv_hw;
  out[i] = mult ? s * mult[i] : s;
}

extern "C" int mclip_pool_finalize(const float* partials, int n, int chunks, int c, int hw, const float* mult, float* out, void* stream) {
  MCLIP_REQUIRE(partials && out && n > 0 && c > 0 && hw > 0, "mclip_pool_finalize: bad arguments");
  mclip_pool_finalize_kernel<<<ceil_div((long long)n * c, 256), 256, 0, (cudaStream_t)stream>>>(partials, n, chunks, c, 1.0f / (float)hw, mult, out);
  MCLIP_CHECK_LAUNCH();
  return MCLIP_OK;
}

// ------------------------------------------------------------------------------------------------
// squeeze-excite FC stack, one CTA per sample
//   s = pooled mean ; z1 = W1 s + b1 ; h = swish(z1) ; z2 = W2 h + b2 ; g = sigmoid(z2)
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(1024) mclip_se_fc_kernel(const float* __restrict__ part, int chunks, int C, int Cse, float inv_hw,
                                                          const float* __restrict__ W1, const float* __restrict__ b1, const float* __restrict__ W2,
                                                          const float* __restrict__ b2, float* __restrict__ s_out, float* __restrict__ z1_out,
                                                          float* __restrict__ gate) {
  extern __shared__ float se_smem[];
  float* s = se_smem;            // [C]
  float* h = se_smem + C;        // [Cse]
  const int n = blockIdx.x;
  for (int c = threadIdx.x; c < C; c += blockDim.x) {
    float v = 0.f;
    for (int k = 0; k < chunks; ++k) v += part[((size_t)n * chunks + k) * C + c];
    v *= inv_hw;
    s[c] = v;
    s_out[(size_t)n * C + c] = v;
  }
  __syncthreads();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nw = blockDim.x >> 5;
  for (int j = warp; j < Cse; j += nw) {
    float v = 0.f;
    for (int c = lane; c < C; c += 32) v = fmaf(W1[(size_t)j * C + c], s[c], v);
    v = warp_sum(v);
    if (lane == 0) {
      v += b1[j];
      z1_out[(size_t)n * Cse + j] = v;
      h[j] = v * sigmoid_precise(v);
    }
  }
  __syncthreads();
  for (int c = threadIdx.x; c < C; c += blockDim.x) {
    float v = b2[c];
    for (int j = 0; j < Cse; ++j) v = fmaf(W2[(size_t)c * Cse + j], h[j], v);
    gate[(size_t)n * C + c] = sigmoid_precise(v);
  }
}

extern "C" int mclip_se_fc(const mclip_se_args* a, void* stream) {
  MCLIP_REQUIRE(a && a->pool_partials && a->w1 && a->b1 && a->w2 && a->b2 && a->pooled && a->z1 && a->gate, "mclip_se_fc: null operand");
  const int smem = (a->c + a->cse) * 4;
  // one CTA per sample: the FC stack is latency bound (a few hundred dependent L2 loads per thread at 256 threads for the
  // 1824/3072-channel blocks), so wide layers get more threads
  const int se_threads = a->c >= 1024 ? 1024 : a->c >= 512 ? 512 : 256;
  mclip_se_fc_kernel<<<a->n, se_threads, smem, (cudaStream_t)stream>>>(a->pool_partials, a->chunks, a->c, a->cse, 1.0f / (float)a->hw, a->w1, a->b1, a->w2, a->b2,
                                                                 a->pooled, a->z1, a->gate);
  MCLIP_CHECK_LAUNCH();
  return MCLIP_OK;
}

// Wg[n,co,ce] = bf16( W[co,ce] * gate[n,ce] ) : the SE gate folded into per-sample project-conv weights
__global__ void mclip_se_scale_weights_kernel(const float* __restrict__ W, const float* __restrict__ gate, bf16* __restrict__ out, int N, int Cout, int Cexp) {
  const long long per = (long long)Cout * Cexp / 8;
  const long long total = per * N;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int n = (int)(i / per);
    const long long r = i % per;
    const int ce = (int)((r * 8) % Cexp);
    const float4* wp = reinterpret_cast<const float4*>(W + r * 8);
    const float4* gp = reinterpret_cast<const float4*>(gate + (size_t)n * Cexp + ce);
    float4 w0 = wp[0], w1 = wp[1], g0 = gp[0], g1 = gp[1];
    float f[8] = {w0.x * g0.x, w0.y * g0.y, w0.z * g0.z, w0.w * g0.w, w1.x * g1.x, w1.y * g1.y, w1.z * g1.z, w1.w * g1.w};
    *reinterpret_cast<bf16x8*>(out + i * 8) = pack8(f);
  }
}

extern "C" int mclip_se_scale_weights(const float* w, const float* gate, void* out, int n, int cout, int cexp, void* stream) {
  MCLIP_REQUIRE(w && gate && out && cexp % 8 == 0, "mclip_se_scale_weights: bad arguments");
  long long total = (long long)n * cout * cexp / 8;
  int grid = (int)((total + 255) / 256);
  if (grid > mclip_num_sms() * 8) grid = mclip_num_sms() * 8;
  mclip_se_scale_weights_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(w, gate, (bf16*)out, n, cout, cexp);
  MCLIP_CHECK_LAUNCH();
  return MCLIP_OK;
}

// ------------------------------------------------------------------------------------------------
// backward streaming passes.
//   upstream gradient of the activation output u:   du = dU[n,px,c] (bf16)  or  dvec[n,c] (broadcast, head pool)
//                                                   du = du * gate[n,c] + dpool[n,c]     (SE)      du *= rowscale[n] (drop-connect)
//   dv = act ? du * swish'(a*y+b) : du      (mode dv_given: dU already holds dv)
//   reduce : partial sums of dv and dv*yhat per channel                       (BatchNorm backward, two-pass)
//   apply  : dY = a * (dv - c1 - yhat*c2)   with a = gamma*invstd, c1 = mean(dv), c2 = mean(dv*yhat)
//   se1    : A2 = gate * u (bf16, operand of the project wgrad) and dgate partials sum_px dU*u
// ------------------------------------------------------------------------------------------------
struct EwBwdDev {
  EwGeom g;
  const bf16* y; const float* scale; const float* shift; int act; int dv_given;
  const bf16* dU; const float* dvec; const float* gate; const float* dpool; const float* rowscale;
  const float* mean; const float* invstd;
  const float* c1; const float* c2;       // apply
  float* part;                            // reduce: [N*chunks][2][C]; se1: [N][chunks][C]
  bf16* out;                              // apply: dY ; se1: A2
};

template <int MODE, bool ASYNC>   // MODE 0 reduce, 1 apply, 2 se pass 1; ASYNC: inputs staged through the per-thread cp.async ring
__global__ void __launch_bounds__(EW_MAX_THREADS, 3) mclip_ew_bwd_kernel(const EwBwdDev p) {
  extern __shared__ float ew_smem[];
  const EwGeom& g = p.g;
  const int cv = threadIdx.x % g.TPP, pl = threadIdx.x / g.TPP;
  const int grp = blockIdx.x % g.G, bc = blockIdx.x / g.G;
  const int n = bc / g.chunks, chunk = bc % g.chunks;
  const int p0 = chunk * g.pix_per_chunk, p1 = min(g.HW, p0 + g.pix_per_chunk);
  const int c = grp * g.CG + cv * EW_CPT;
  constexpr int NACC = (MODE == 2) ? 5 : 2;
  float2 a[2], b[2], mu[2], is[2], gt[2], dp[2], dvv[2], k1[2], k2[2], acc[NACC][2];
  const float rsv = p.rowscale ? p.rowscale[n] : 1.f;
#pragma unroll
  for (int i = 0; i < 2; ++i) {
    const int ci = c + 2 * i;
    auto ld2 = [&](const float* q, float dflt) { return q ? make_float2(q[ci], q[ci + 1]) : make_float2(dflt, dflt); };
    auto ld2n = [&](const float* q, float dflt) { return q ? make_float2(q[(size_t)n * g.C + ci], q[(size_t)n * g.C + ci + 1]) : make_float2(dflt, dflt); };
    a[i] = ld2(p.scale, 1.f); b[i] = ld2(p.shift, 0.f); mu[i] = ld2(p.mean, 0.f); is[i] = ld2(p.invstd, 1.f);
    gt[i] = ld2n(p.gate, 1.f); dp[i] = ld2n(p.dpool, 0.f); dvv[i] = ld2n(p.dvec, 0.f);
    if (MODE != 2) { gt[i].x *= rsv; gt[i].y *= rsv; dp[i].x *= rsv; dp[i].y *= rsv; }    // drop-connect row scale folded in
    k1[i] = (MODE == 1) ? ld2(p.c1, 0.f) : make_float2(0.f, 0.f);
    k2[i] = (MODE == 1) ? ld2(p.c2, 0.f) : make_float2(0.f, 0.f);
#pragma unroll
    for (int q = 0; q < NACC; ++q) acc[q][i] = make_float2(0.f, 0.f);
  }
  const float2 one = make_float2(1.f, 1.f);
  const size_t base = (size_t)n * g.HW * g.C + c;
  // one pixel of this thread's 4 channels: yw = y (2 x bf16x2), dw_ = dU (ignored when p.dU == nullptr)
  auto pixel = [&](int px, const uint32_t (&yw)[2], const uint32_t (&dw_)[2]) {
    uint32_t ow[2];
#pragma unroll
    for (int i = 0; i < 2; ++i) {
      const float2 y = bf2_to_f2(yw[i]);
      const float2 du = p.dU ? bf2_to_f2(dw_[i]) : dvv[i];
      const float2 yh = fmul2(make_float2(y.x - mu[i].x, y.y - mu[i].y), is[i]);
      if (MODE == 2) {
        const float2 v = ffma2r(y, a[i], b[i]);
        const float2 sg = sigmoid2(v);
        const float2 u_ = fmul2(v, sg);                                             // swish(v)
        const float2 sp = fmul2(sg, ffma2r(v, make_float2(1.f - sg.x, 1.f - sg.y), one));   // swish'(v)
        const float2 dsp = fmul2(du, sp);
        ffma2(acc[0][i], du, u_);          // d gate
        ffma2(acc[1][i], dsp, one);        // sum dU s'
        ffma2(acc[2][i], sp, one);         // sum s'
        ffma2(acc[3][i], dsp, yh);         // sum dU s' yhat
        ffma2(acc[4][i], sp, yh);          // sum s' yhat
        const float2 o = fmul2(u_, gt[i]);
        ow[i] = pack_bf16(o.x, o.y);
      } else {
        float2 dv;
        if (p.dv_given) dv = du;
        else {
          dv = ffma2r(du, gt[i], dp[i]);
          if (p.act) {
            const float2 v = ffma2r(y, a[i], b[i]);
            const float2 sg = sigmoid2(v);
            dv = fmul2(dv, fmul2(sg, ffma2r(v, make_float2(1.f - sg.x, 1.f - sg.y), one)));
          }
        }
        if (MODE == 0) { ffma2(acc[0][i], dv, one); ffma2(acc[1][i], dv, yh); }
        else {
          const float2 t = ffma2r(yh, make_float2(-k2[i].x, -k2[i].y), make_float2(dv.x - k1[i].x, dv.y - k1[i].y));
          const float2 o = fmul2(a[i], t);
          ow[i] = pack_bf16(o.x, o.y);
        }
      }
    }
    if (MODE != 0) stg_b64(p.out + base + (size_t)px * g.C, make_uint2(ow[0], ow[1]));
  };
  if (ASYNC) {
    // ring slots follow the block-reduction scratch ([PL][CG] floats); slot s of thread t at (s*blockDim + t)*16
    const uint32_t slot_stride = blockDim.x * 16u;
    const uint32_t my = smem_u32(ew_smem) + (uint32_t)(g.PL * g.CG) * 4u + threadIdx.x * 16u;
    const bool has_du = p.dU != nullptr;
    int px_issue = p0 + pl;
    auto issue = [&](int slot) {
      if (px_issue < p1) {
        const size_t off = base + (size_t)px_issue * g.C;
        cp_async8(my + slot * slot_stride, p.y + off);
        if (has_du) cp_async8(my + slot * slot_stride + 8u, p.dU + off);
      }
      cp_async_commit();                   // always commit: the wait below counts groups
      px_issue += g.PL;
    };
#pragma unroll
    for (int s = 0; s < EW_RING - 1; ++s) issue(s);
    int slot = 0;
    for (int px = p0 + pl; px < p1; px += g.PL) {
      issue(slot == 0 ? EW_RING - 1 : slot - 1);          // refill the slot consumed by the previous iteration
      cp_async_wait<EW_RING - 1>();
      const uint4 v = lds128(my + slot * slot_stride);
      const uint32_t yw[2] = {v.x, v.y}, dw_[2] = {v.z, v.w};
      pixel(px, yw, dw_);
      slot = slot == EW_RING - 1 ? 0 : slot + 1;
    }
    cp_async_wait<0>();
  } else {
    constexpr int UNR = (MODE == 2) ? 2 : EW_UNR;
    for (int px0 = p0 + pl; px0 < p1; px0 += g.PL * UNR) {
      uint2 yv[UNR], dv_[UNR];
#pragma unroll
      for (int u = 0; u < UNR; ++u) {
        const int px = px0 + u * g.PL;
        if (px < p1) {
          const size_t off = base + (size_t)px * g.C;
          yv[u] = ldg_b64(p.y + off);
          if (p.dU) dv_[u] = ldg_b64(p.dU + off);
        }
      }
#pragma unroll
      for (int u = 0; u < UNR; ++u) {
        const int px = px0 + u * g.PL;
        if (px >= p1) break;
        const uint32_t yw[2] = {yv[u].x, yv[u].y}, dw_[2] = {dv_[u].x, dv_[u].y};
        pixel(px, yw, dw_);
      }
    }
  }
  if (MODE == 0) {
    float* dst = p.part + (size_t)bc * 2 * g.C + grp * g.CG;
    ew_block_reduce4(ew_smem, acc[0], cv, pl, g.CG, g.PL, dst);
    ew_block_reduce4(ew_smem, acc[1], cv, pl, g.CG, g.PL, dst + g.C);
  } else if (MODE == 2) {
    float* dst = p.part + (size_t)bc * 5 * g.C + grp * g.CG;
#pragma unroll
    for (int q = 0; q < 5; ++q) ew_block_reduce4(ew_smem, acc[q], cv, pl, g.CG, g.PL, dst + (size_t)q * g.C);
  }
}

// BN1-backward sums from the SE pass-1 partials: with dv = (dU*gate + dpool) * s'(v),
//   sum dv      = gate * sum(dU s')      + dpool * sum(s')
//   sum dv*yhat = gate * sum(dU s' yhat) + dpool * sum(s' yhat)        per (sample, channel)  ->  out [N][2][C]
__global__ void mclip_se_bn_combine_kernel(const float* __restrict__ part, int chunks, int C, const float* __restrict__ gate,
                                           const float* __restrict__ dpool, float* __restrict__ out, int N) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= N * C) return;
  const int n = i / C, c = i % C;
  float s1 = 0.f, s2 = 0.f, s3 = 0.f, s4 = 0.f;
  for (int k = 0; k < chunks; ++k) {
    const float* q = part + ((size_t)n * chunks + k) * 5 * C + c;
    s1 += q[C]; s2 += q[2 * C]; s3 += q[3 * C]; s4 += q[4 * C];
  }
  const float g = gate[i], d = dpool[i];
  out[((size_t)n * 2 + 0) * C + c] = g * s1 + d * s2;
  out[((size_t)n * 2 + 1) * C + c] = g * s3 + d * s4;
}

extern "C" int mclip_se_bn_combine(const float* partials, int n, int chunks, int c, const float* gate, const float* dpool, float* out, void* stream) {
  MCLIP_REQUIRE(partials && gate && dpool && out && n > 0 && c > 0, "mclip_se_bn_combine: bad arguments");
  mclip_se_bn_combine_kernel<<<ceil_div((long long)n * c, 256), 256, 0, (cudaStream_t)stream>>>(partials, chunks, c, gate, dpool, out, n);
  MCLIP_CHECK_LAUNCH();
  return MCLIP_OK;
}

extern "C" int mclip_ew_backward(const mclip_ew_bwd_args* a, void* stream) {
  MCLIP_REQUIRE(a && a->y && (a->du || a->dvec), "mclip_ew_backward: null input");
  MCLIP_REQUIRE(a->mode >= 0 && a->mode <= 2, "mclip_ew_backward: mode %d", a->mode);
  EwBwdDev p; int threads;
  int rc = ew_geom(a->n, a->hw, a->c, &p.g, &threads);
  if (rc) return rc;
  if (a->mode != 1) MCLIP_REQUIRE(a->partials && a->chunks == p.g.chunks, "mclip_ew_backward: chunks=%d, expected %d", a->chunks, p.g.chunks);
  if (a->mode != 0) MCLIP_REQUIRE(a->out, "mclip_ew_backward: null output");
  p.y = (const bf16*)a->y; p.scale = a->scale; p.shift = a->shift; p.act = a->act; p.dv_given = a->dv_given;
  p.dU = (const bf16*)a->du; p.dvec = a->dvec; p.gate = a->gate; p.dpool = a->dpool; p.rowscale = a->rowscale;
  p.mean = a->mean; p.invstd = a->invstd; p.c1 = a->c1; p.c2 = a->c2; p.part = a->partials; p.out = (bf16*)a->out;
  const int red_smem = p.g.PL * p.g.CG * 4;                               // multiple of 16 bytes (CG % 4 == 0)
  const int grid = a->n * p.g.chunks * p.g.G;
  cudaStream_t st = (cudaStream_t)stream;
  // MCLIP_EW_ASYNC: bit m selects the cp.async ring for backward mode m (bit 3: forward pass); default: SE pass 1 only
  const int async_mask = ew_async_mask();
  const int ring_smem = red_smem + EW_RING * threads * 16;
  static int carveout_set = 0;
  if (!carveout_set) {       // three 36 KB blocks per SM: ask for a shared-memory carveout that holds them
    MCLIP_CHECK_CUDA(cudaFuncSetAttribute(mclip_ew_bwd_kernel<0, true>, cudaFuncAttributePreferredSharedMemoryCarveout, 60));
    MCLIP_CHECK_CUDA(cudaFuncSetAttribute(mclip_ew_bwd_kernel<1, true>, cudaFuncAttributePreferredSharedMemoryCarveout, 60));
    MCLIP_CHECK_CUDA(cudaFuncSetAttribute(mclip_ew_bwd_kernel<2, true>, cudaFuncAttributePreferredSharedMemoryCarveout, 60));
    carveout_set = 1;
  }
  if ((async_mask >> a->mode) & 1) {
    if (a->mode == 0) mclip_ew_bwd_kernel<0, true><<<grid, threads, ring_smem, st>>>(p);
    else if (a->mode == 1) mclip_ew_bwd_kernel<1, true><<<grid, threads, ring_smem, st>>>(p);
    else mclip_ew_bwd_kernel<2, true><<<grid, threads, ring_smem, st>>>(p);
  } else {
    if (a->mode == 0) mclip_ew_bwd_kernel<0, false><<<grid, threads, red_smem, st>>>(p);
    else if (a->mode == 1) mclip_ew_bwd_kernel<1, false><<<grid, threads, red_smem, st>>>(p);
    else mclip_ew_bwd_kernel<2, false><<<grid, threads, red_smem, st>>>(p);
  }
  MCLIP_CHECK_LAUNCH();
  return MCLIP_OK;
}

// BatchNorm backward, second half: dgamma, dbeta, and the two means used by the apply pass
__global__ void __launch_bounds__(BNF_CH * BNF_LANES) mclip_bn_bwd_finalize_kernel(const float* __restrict__ partials, int slots, int C, double count,
                                                                                   int training, float* dgamma, float* dbeta, int accumulate, float* c1,
                                                                                   float* c2) {
  __shared__ double sm[2][BNF_LANES][BNF_CH];
  const int c = blockIdx.x * BNF_CH + threadIdx.x % BNF_CH, lane = threadIdx.x / BNF_CH;
  double s, q;
  bn_reduce_slots(partials, slots, C, c, lane, s, q, sm);
  if (lane != 0 || c >= C) return;
  if (dgamma) dgamma[c] = accumulate ? dgamma[c] + (float)q : (float)q;
  if (dbeta) dbeta[c] = accumulate ? dbeta[c] + (float)s : (float)s;
  c1[c] = training ? (float)(s / count) : 0.f;
  c2[c] = training ? (float)(q / count) : 0.f;
}

extern "C" int mclip_bn_bwd_finalize(const float* partials, int slots, int c, long long count, int training, float* dgamma, float* dbeta, int accumulate,
                                     float* c1, float* c2, void* stream) {
  MCLIP_REQUIRE(partials && c1 && c2 && slots > 0 && count > 0, "mclip_bn_bwd_finalize: bad arguments");
  mclip_bn_bwd_finalize_kernel<<<ceil_div(c, BNF_CH), BNF_CH * BNF_LANES, 0, (cudaStream_t)stream>>>(partials, slots, c, (double)count, training, dgamma, dbeta, accumulate, c1, c2);
  MCLIP_CHECK_LAUNCH();
  return MCLIP_OK;
}

// ------------------------------------------------------------------------------------------------
// squeeze-excite backward
//   k1 (per sample): dz2 = dg*g*(1-g) ; dh = W2^T dz2 ; dz1 = dh*swish'(z1) ; ds = W1^T dz1 ; dpool = ds/HW
//   k2 (per output element): dW2 = sum_n dz2 h^T ; db2 = sum_n dz2 ; dW1 = sum_n dz1 s^T ; db1 = sum_n dz1
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(1024) mclip_se_bwd1_kernel(const float* __restrict__ dg_part, int chunks, int cstride, int C, int Cse, float inv_hw,
                                                            const float* __restrict__ W1, const float* __restrict__ W2, const float* __restrict__ z1,
                                                            const float* __restrict__ gate, float* __restrict__ dz2_out, float* __restrict__ dz1_out,
                                                            float* __restrict__ dpool) {
  extern __shared__ float se_smem[];
  float* dz2 = se_smem;          // [C]
  float* dz1 = se_smem + C;      // [Cse]
  const int n = blockIdx.x;
  for (int c = threadIdx.x; c < C; c += blockDim.x) {
    float dg = 0.f;
    for (int k = 0; k < chunks; ++k) dg += dg_part[((size_t)n * chunks + k) * cstride + c];
    const float g = gate[(size_t)n * C + c];
    const float v = dg * g * (1.f - g);
    dz2[c] = v;
    dz2_out[(size_t)n * C + c] = v;
  }
  __syncthreads();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nw = blockDim.x >> 5;
  for (int j = warp; j < Cse; j += nw) {
    float v = 0.f;
    for (int c = lane; c < C; c += 32) v = fmaf(W2[(size_t)c * Cse + j], dz2[c], v);
    v = warp_sum(v);
    if (lane == 0) {
      const float z = z1[(size_t)n * Cse + j];
      const float s = sigmoid_precise(z);
      v *= s * (1.f + z * (1.f - s));
      dz1[j] = v;
      dz1_out[(size_t)n * Cse + j] = v;
    }
  }
  __syncthreads();
  for (int c = threadIdx.x; c < C; c += blockDim.x) {
    float v = 0.f;
    for (int j = 0; j < Cse; ++j) v = fmaf(W1[(size_t)j * C + c], dz1[j], v);
    dpool[(size_t)n * C + c] = v * inv_hw;
  }
}

__global__ void mclip_se_bwd2_kernel(int N, int C, int Cse, const float* __restrict__ dz2, const float* __restrict__ dz1, const float* __restrict__ z1,
                                     const float* __restrict__ s, float* dW1, float* db1, float* dW2, float* db2, int accumulate) {
  const int total = 2 * C * Cse + C + Cse;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
    float v = 0.f;
    float* dst;
    if (i < C * Cse) {                       // dW2[c][j] = sum_n dz2[n,c] * h[n,j]
      const int c = i / Cse, j = i % Cse;
      for (int n = 0; n < N; ++n) { const float z = z1[(size_t)n * Cse + j]; v = fmaf(dz2[(size_t)n * C + c], z * sigmoid_precise(z), v); }
      dst = dW2 + i;
    } else if (i < 2 * C * Cse) {            // dW1[j][c] = sum_n dz1[n,j] * s[n,c]
      const int k = i - C * Cse, j = k / C, c = k % C;
      for (int n = 0; n < N; ++n) v = fmaf(dz1[(size_t)n * Cse + j], s[(size_t)n * C + c], v);
      dst = dW1 + k;
    } else if (i < 2 * C * Cse + C) {
      const int c = i - 2 * C * Cse;
      for (int n = 0; n < N; ++n) v += dz2[(size_t)n * C + c];
      dst = db2 + c;
    } else {
      const int j = i - 2 * C * Cse - C;
      for (int n = 0; n < N; ++n) v += dz1[(size_t)n * Cse + j];
      dst = db1 + j;
    }
    *dst = accumulate ? *dst + v : v;
  }
}

extern "C" int mclip_se_fc_backward(const mclip_se_args* a, void* stream) {
  MCLIP_REQUIRE(a && a->dgate_partials && a->w1 && a->w2 && a->z1 && a->gate && a->pooled && a->dz2 && a->dz1 && a->dpool && a->dw1 && a->db1 && a->dw2 && a->db2,
                "mclip_se_fc_backward: null operand");
  const int smem = (a->c + a->cse) * 4;
  const int se_threads = a->c >= 1024 ? 1024 : a->c >= 512 ? 512 : 256;
  mclip_se_bwd1_kernel<<<a->n, se_threads, smem, (cudaStream_t)stream>>>(a->dgate_partials, a->chunks, a->dgate_chunk_stride > 0 ? a->dgate_chunk_stride : a->c, a->c, a->cse, 1.0f / (float)a->hw, a->w1, a->w2, a->z1, a->gate,
                                                                   a->dz2, a->dz1, a->dpool);
  MCLIP_CHECK_LAUNCH();
  const int total = 2 * a->c * a->cse + a->c + a->cse;
  int grid = ceil_div(total, 256);
  mclip_se_bwd2_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(a->n, a->c, a->cse, a->dz2, a->dz1, a->z1, a->pooled, a->dw1, a->db1, a->dw2, a->db2, a->accumulate);
  MCLIP_CHECK_LAUNCH();
  return MCLIP_OK;
}

// ------------------------------------------------------------------------------------------------
// fp32 master weights -> bf16 operands (straight and transposed), table driven: one launch for a whole tower
// ------------------------------------------------------------------------------------------------
// 32x32 tiles through shared memory: coalesced fp32 reads, coalesced bf16 writes of both the straight and the transposed copy
__global__ void __launch_bounds__(256) mclip_weight_prep_kernel(const mclip_prep_entry* __restrict__ table, int n_entries) {
  __shared__ bf16 tile[32][33];
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;          // 32 x 8
  for (int e = blockIdx.y; e < n_entries; e += gridDim.y) {
    const mclip_prep_entry t = table[e];
    const float* src = (const float*)t.src;
    bf16* dst = (bf16*)t.dst;
    bf16* dstT = (bf16*)t.dst_t;
    const int ld = t.dst_ld > 0 ? t.dst_ld : t.cols, ldt = t.dst_t_ld > 0 ? t.dst_t_ld : t.rows;
    const int tiles_c = (t.cols + 31) >> 5, tiles_r = (t.rows + 31) >> 5;
    for (int tl = blockIdx.x; tl < tiles_c * tiles_r; tl += gridDim.x) {
      const int r0 = (tl / tiles_c) << 5, c0 = (tl % tiles_c) << 5;
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const int r = r0 + ty + 8 * i, c = c0 + tx;
        if (r < t.rows && c < t.cols) {
          const bf16 v = __float2bfloat16_rn(src[(size_t)r * t.cols + c]);
          if (dst) dst[(size_t)r * ld + c] = v;
          tile[ty + 8 * i][tx] = v;
        }
      }
      if (dstT) {
        __syncthreads();
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const int c = c0 + ty + 8 * i, r = r0 + tx;              // transposed copy: row c, column r
          if (r < t.rows && c < t.cols) dstT[(size_t)c * ldt + r] = tile[tx][ty + 8 * i];
        }
        __syncthreads();
      }
    }
  }
}

extern "C" int mclip_weight_prep(const void* table_dev, int n_entries, void* stream) {
  MCLIP_REQUIRE(table_dev && n_entries > 0, "mclip_weight_prep: empty table");
  dim3 grid(32, n_entries < 1024 ? n_entries : 1024);
  mclip_weight_prep_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>((const mclip_prep_entry*)table_dev, n_entries);
  MCLIP_CHECK_LAUNCH();
  return MCLIP_OK;
}


// ------------------------------------------------------------------------------------------------
// folded BatchNorm backward of the expand convolution (see include/mclip.h): small per-block weight-space kernels
// ------------------------------------------------------------------------------------------------
// phase 0: 32x32 tiles of We [cexp, cin]: twe = bf16(t We) (straight), wcat[j, k] = bf16(a[k] We[k, j]) (transposed)
__global__ void __launch_bounds__(256) mclip_bn0_fold_prep_kernel(const mclip_bn0_fold_args a) {
  __shared__ float tile[32][33];
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  const int tiles_j = (a.cin + 31) >> 5, tiles_k = (a.k1pad + 31) >> 5;
  bf16* wcat = (bf16*)a.wcat; bf16* twe = (bf16*)a.twe;
  for (int tl = blockIdx.x; tl < tiles_j * tiles_k; tl += gridDim.x) {
    const int k0 = (tl / tiles_j) << 5, j0 = (tl % tiles_j) << 5;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int k = k0 + ty + 8 * i, j = j0 + tx;
      float av = 0.f;
      if (k < a.cexp && j < a.cin) {
        const float w = a.we[(size_t)k * a.cin + j], sc = a.scale[k];
        av = sc * w;
        twe[(size_t)k * a.cin + j] = __float2bfloat16_rn(-sc * a.c2[k] * a.invstd[k] * w);
      }
      tile[ty + 8 * i][tx] = av;
    }
    __syncthreads();
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int j = j0 + ty + 8 * i, k = k0 + tx;                 // transposed write: row j, column k (zeros in [cexp, k1pad))
      if (j < a.cin && k < a.k1pad) wcat[(size_t)j * a.ldw + k] = __float2bfloat16_rn(tile[tx][ty + 8 * i]);
    }
    __syncthreads();
  }
}
// Column reductions over the weight matrices: CTA = 32 columns (lane = column, coalesced rows) x 32 warps striding the reduction
// index, fixed-order cross-warp sum (deterministic).  (A thread per column walking all Cexp <= 3072 rows alone took 190 us per
// launch: 10 ms per c3 step for two "tiny" kernels.)
#define FOLD_WARPS 32
__device__ __forceinline__ float fold_block_sum(float v, float (*red)[33]) {      // result valid in warp 0
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  red[warp][lane] = v;
  __syncthreads();
  float s = 0.f;
  if (warp == 0)
#pragma unroll
    for (int w = 0; w < FOLD_WARPS; ++w) s += red[w][lane];
  return s;
}
// bias[j] = - sum_k a[k] c1[k] We[k, j]
__global__ void __launch_bounds__(32 * FOLD_WARPS) mclip_bn0_fold_bias_kernel(const mclip_bn0_fold_args a) {
  __shared__ float red[FOLD_WARPS][33];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int j = blockIdx.x * 32 + lane;
  float s = 0.f;
  if (j < a.cin)
    for (int k = warp; k < a.cexp; k += FOLD_WARPS) s = fmaf(a.scale[k] * a.c1[k], a.we[(size_t)k * a.cin + j], s);
  s = fold_block_sum(s, red);
  if (warp == 0 && j < a.cin) a.bias[j] = -s;
}
// phase 1: wcat[j, k1pad + i] = bf16(G[i, j]);  bias[j] -= sum_i xbar[i] * bf16(G[i, j])   (the SAME rounded G the GEMM multiplies X by)
__global__ void __launch_bounds__(32 * FOLD_WARPS) mclip_bn0_fold_g_kernel(const mclip_bn0_fold_args a) {
  __shared__ float red[FOLD_WARPS][33];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int j = blockIdx.x * 32 + lane;
  bf16* wcat = (bf16*)a.wcat;
  const float inv = (float)(1.0 / a.count);
  float s = 0.f;
  if (j < a.cin)
    for (int i = warp; i < a.cin; i += FOLD_WARPS) {
      const bf16 g = __float2bfloat16_rn(a.g[(size_t)i * a.cin + j]);
      wcat[(size_t)j * a.ldw + a.k1pad + i] = g;
      s = fmaf(a.sumx[i] * inv, __bfloat162float(g), s);
    }
  s = fold_block_sum(s, red);
  if (warp == 0 && j < a.cin) a.bias[j] -= s;
}
// phase 2: gc[j, i] = bf16(XtX[i, j] - sumx[i] sumx[j] / count)     (symmetric; stored as the [N, K] operand of mclip_gemm_tn)
__global__ void mclip_bn0_fold_center_kernel(const mclip_bn0_fold_args a) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= a.cin * a.cin) return;
  const int j = idx / a.cin, i = idx % a.cin;
  ((bf16*)a.gc)[idx] = __float2bfloat16_rn(a.g[(size_t)i * a.cin + j] - (float)((double)a.sumx[i] * (double)a.sumx[j] / a.count));
}
// phase 3: dwe[k, j] = a[k] (dwe[k, j] - c1[k] sumx[j]) + t[k] q[k, j]
__global__ void mclip_bn0_fold_dwe_kernel(const mclip_bn0_fold_args a) {
  const size_t total = (size_t)a.cexp * a.cin;
  for (size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (size_t)gridDim.x * blockDim.x) {
    const int k = (int)(idx / a.cin), j = (int)(idx % a.cin);
    const float sc = a.scale[k], t = -sc * a.c2[k] * a.invstd[k];
    a.dwe[idx] = sc * (a.dwe[idx] - a.c1[k] * a.sumx[j]) + t * __bfloat162float(((const bf16*)a.q)[idx]);
  }
}

extern "C" int mclip_bn0_fold(const mclip_bn0_fold_args* a, int phase, void* stream_) {
  cudaStream_t st = (cudaStream_t)stream_;
  MCLIP_REQUIRE(a && a->cexp > 0 && a->cin > 0 && a->count > 0, "mclip_bn0_fold: bad arguments");
  if (phase == 0) {
    MCLIP_REQUIRE(a->we && a->scale && a->invstd && a->c1 && a->c2 && a->wcat && a->twe && a->bias && a->k1pad >= a->cexp && a->ldw >= a->k1pad + a->cin,
                  "mclip_bn0_fold phase 0: null operand / bad layout");
    const int tiles = ((a->cin + 31) >> 5) * ((a->k1pad + 31) >> 5);
    mclip_bn0_fold_prep_kernel<<<tiles < 1184 ? tiles : 1184, 256, 0, st>>>(*a);
    mclip_bn0_fold_bias_kernel<<<ceil_div(a->cin, 32), 32 * FOLD_WARPS, 0, st>>>(*a);
  } else if (phase == 1) {
    MCLIP_REQUIRE(a->g && a->sumx && a->wcat && a->bias, "mclip_bn0_fold phase 1: null operand");
    mclip_bn0_fold_g_kernel<<<ceil_div(a->cin, 32), 32 * FOLD_WARPS, 0, st>>>(*a);
  } else if (phase == 2) {
    MCLIP_REQUIRE(a->g && a->sumx && a->gc, "mclip_bn0_fold phase 2: null operand");
    mclip_bn0_fold_center_kernel<<<ceil_div((long long)a->cin * a->cin, 256), 256, 0, st>>>(*a);
  } else if (phase == 3) {
    MCLIP_REQUIRE(a->dwe && a->q && a->sumx && a->scale && a->invstd && a->c1 && a->c2, "mclip_bn0_fold phase 3: null operand");
    const long long total = (long long)a->cexp * a->cin;
    int grid = (int)((total + 255) / 256); if (grid > mclip_num_sms() * 8) grid = mclip_num_sms() * 8;
    mclip_bn0_fold_dwe_kernel<<<grid, 256, 0, st>>>(*a);
  } else {
    MCLIP_REQUIRE(false, "mclip_bn0_fold: phase %d", phase);
  }
  MCLIP_CHECK_LAUNCH();
  return MCLIP_OK;
}

// ------------------------------------------------------------------------------------------------
// small row-wise ops of the CLIP head: fp32 -> bf16 cast, L2 normalisation fwd/bwd (clip.py:90-91), bias gradient
// ------------------------------------------------------------------------------------------------
__global__ void mclip_cast_bf16_kernel(const float* __restrict__ in, bf16* __restrict__ out, long long n) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) out[i] = __float2bfloat16_rn(in[i]);
}
extern "C" int mclip_cast_bf16(const float* in, void* out, long long n, void* stream) {
  MCLIP_REQUIRE(in && out && n > 0, "mclip_cast_bf16: bad arguments");
  int grid = (int)((n + 255) / 256); if (grid > mclip_num_sms() * 8) grid = mclip_num_sms() * 8;
  mclip_cast_bf16_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(in, (bf16*)out, n);
  MCLIP_CHECK_LAUNCH();
  return MCLIP_OK;
}

// one warp per row: e = x / ||x||   (no epsilon, clip.py:90-91); x is bf16 (projection GEMM output)
__global__ void mclip_l2norm_fwd_kernel(const bf16* __restrict__ x, float* __restrict__ e, float* __restrict__ norm, int rows, int D) {
  const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  if (row >= rows) return;
  float ss = 0.f;
  for (int d = lane; d < D; d += 32) { float v = __bfloat162float(x[(size_t)row * D + d]); ss = fmaf(v, v, ss); }
  ss = warp_sum(ss);
  const float nrm = sqrtf(ss), inv = 1.0f / nrm;
  for (int d = lane; d < D; d += 32) e[(size_t)row * D + d] = __bfloat162float(x[(size_t)row * D + d]) * inv;
  if (lane == 0) norm[row] = nrm;
}
// dx = (de - e * <e,de>) / ||x||  -> bf16 (operand of the projection dgrad/wgrad GEMMs)
__global__ void mclip_l2norm_bwd_kernel(const float* __restrict__ e, const float* __restrict__ de, const float* __restrict__ norm, bf16* __restrict__ dx,
                                        int rows, int D) {
  const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  if (row >= rows) return;
  float dot = 0.f;
  for (int d = lane; d < D; d += 32) dot = fmaf(e[(size_t)row * D + d], de[(size_t)row * D + d], dot);
  dot = warp_sum(dot);
  const float inv = 1.0f / norm[row];
  for (int d = lane; d < D; d += 32) dx[(size_t)row * D + d] = __float2bfloat16_rn((de[(size_t)row * D + d] - e[(size_t)row * D + d] * dot) * inv);
}
extern "C" int mclip_l2norm_forward(const void* x, float* e, float* norm, int rows, int d, void* stream) {
  MCLIP_REQUIRE(x && e && norm && rows > 0 && d > 0, "mclip_l2norm_forward: bad arguments");
  mclip_l2norm_fwd_kernel<<<ceil_div(rows, 4), 128, 0, (cudaStream_t)stream>>>((const bf16*)x, e, norm, rows, d);
  MCLIP_CHECK_LAUNCH();
  return MCLIP_OK;
}
extern "C" int mclip_l2norm_backward(const float* e, const float* de, const float* norm, void* dx, int rows, int d, void* stream) {
  MCLIP_REQUIRE(e && de && norm && dx && rows > 0 && d > 0, "mclip_l2norm_backward: bad arguments");
  mclip_l2norm_bwd_kernel<<<ceil_div(rows, 4), 128, 0, (cudaStream_t)stream>>>(e, de, norm, (bf16*)dx, rows, d);
  MCLIP_CHECK_LAUNCH();
  return MCLIP_OK;
}

// column sums of a bf16 [rows, cols] matrix -> fp32 (bias gradients of Linear layers)
__global__ void mclip_colsum_kernel(const bf16* __restrict__ x, float* __restrict__ out, int rows, int cols, long long ld, int accumulate) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= cols) return;
  float s = 0.f;
  for (int r = 0; r < rows; ++r) s += __bfloat162float(x[(size_t)r * ld + c]);
  out[c] = accumulate ? out[c] + s : s;
}
// Stage 1: CTA (column chunk of 64, row split): 8 warps stride the split's rows, one bf16x2 per lane, fixed-order smem
// reduction -> partial[split][cols].  Stage 2: out[c] (+)= sum_splits partial (fixed order => deterministic).
#define COLSUM_WARPS 8
#define COLSUM_MAX_SPLITS 64
__global__ void __launch_bounds__(32 * COLSUM_WARPS) mclip_colsum_part_kernel(const bf16* __restrict__ x, float* __restrict__ partial, int rows, int cols,
                                                                              long long ld, int rows_per_split) {
  __shared__ float red[COLSUM_WARPS][64];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int c = blockIdx.x * 64 + lane * 2;
  const int r0 = blockIdx.y * rows_per_split, r1 = min(rows, r0 + rows_per_split);
  float s0 = 0.f, s1 = 0.f;
  if (c < cols) {
#pragma unroll 4
    for (int r = r0 + warp; r < r1; r += COLSUM_WARPS) {
      const uint32_t w = *reinterpret_cast<const uint32_t*>(x + (size_t)r * ld + c);
      s0 += bf16_lo(w); s1 += bf16_hi(w);
    }
  }
  red[warp][lane * 2] = s0; red[warp][lane * 2 + 1] = s1;
  __syncthreads();
  if (threadIdx.x < 64 && blockIdx.x * 64 + threadIdx.x < cols) {
    float a = 0.f;
#pragma unroll
    for (int w = 0; w < COLSUM_WARPS; ++w) a += red[w][threadIdx.x];
    partial[(size_t)blockIdx.y * cols + blockIdx.x * 64 + threadIdx.x] = a;
  }
}
__global__ void mclip_colsum_final_kernel(const float* __restrict__ partial, float* __restrict__ out, int splits, int cols, int accumulate) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= cols) return;
  float a = 0.f;
  for (int s = 0; s < splits; ++s) a += partial[(size_t)s * cols + c];
  out[c] = accumulate ? out[c] + a : a;
}
static int colsum_splits(int rows) {
  int s = ceil_div(rows, 64);                      // >= 64 rows (8 per warp) per split
  return s < 1 ? 1 : (s > COLSUM_MAX_SPLITS ? COLSUM_MAX_SPLITS : s);
}
extern "C" long long mclip_colsum_workspace_bytes(int rows, int cols) { return (long long)colsum_splits(rows) * cols * 4; }
extern "C" int mclip_colsum(const void* x, float* out, int rows, int cols, long long ld, int accumulate, void* workspace, long long workspace_bytes,
                            void* stream) {
  MCLIP_REQUIRE(x && out && rows > 0 && cols > 0, "mclip_colsum: bad arguments");
  if (cols % 2 == 0 && ld % 2 == 0 && ((uintptr_t)x & 3) == 0 && workspace) {
    const int splits = colsum_splits(rows);
    MCLIP_REQUIRE(workspace_bytes >= (long long)splits * cols * 4, "mclip_colsum: workspace too small (%lld < %lld)", workspace_bytes,
                  (long long)splits * cols * 4);
    dim3 grid(ceil_div(cols, 64), splits);
    mclip_colsum_part_kernel<<<grid, 32 * COLSUM_WARPS, 0, (cudaStream_t)stream>>>((const bf16*)x, (float*)workspace, rows, cols, ld, ceil_div(rows, splits));
    MCLIP_CHECK_LAUNCH();
    mclip_colsum_final_kernel<<<ceil_div(cols, 256), 256, 0, (cudaStream_t)stream>>>((const float*)workspace, out, splits, cols, accumulate);
  } else {
    mclip_colsum_kernel<<<ceil_div(cols, 128), 128, 0, (cudaStream_t)stream>>>((const bf16*)x, out, rows, cols, ld, accumulate);
  }
  MCLIP_CHECK_LAUNCH();
  return MCLIP_OK;
}

// ------------------------------------------------------------------------------------------------
// AdamW over one flat fp32 parameter buffer (torch.optim.AdamW semantics: decoupled weight decay,
// bias-corrected moments; breastclip/optimizer/__init__.py:23-31 builds AdamW(lr, weight_decay) on all params)
// ------------------------------------------------------------------------------------------------
__global__ void mclip_adamw_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m, float* __restrict__ v, long long n, float lr,
                                   float beta1, float beta2, float eps, float wd, float bc1, float bc2_sqrt, float grad_scale) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const float gi = g[i] * grad_scale;
    float pi = p[i] * (1.f - lr * wd);
    const float mi = beta1 * m[i] + (1.f - beta1) * gi;
    const float vi = beta2 * v[i] + (1.f - beta2) * gi * gi;
    m[i] = mi; v[i] = vi;
    const float denom = sqrtf(vi) / bc2_sqrt + eps;
    pi -= (lr / bc1) * (mi / denom);
    p[i] = pi;
  }
}
extern "C" int mclip_adamw_step(float* param, const float* grad, float* exp_avg, float* exp_avg_sq, long long n, float lr, float beta1, float beta2,
                                float eps, float weight_decay, long long step, float grad_scale, void* stream) {
  MCLIP_REQUIRE(param && grad && exp_avg && exp_avg_sq && n > 0 && step >= 1, "mclip_adamw_step: bad arguments");
  const float bc1 = 1.f - powf(beta1, (float)step);
  const float bc2 = sqrtf(1.f - powf(beta2, (float)step));
  int grid = (int)((n + 255) / 256); if (grid > mclip_num_sms() * 8) grid = mclip_num_sms() * 8;
  mclip_adamw_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(param, grad, exp_avg, exp_avg_sq, n, lr, beta1, beta2, eps, weight_decay, bc1, bc2, grad_scale);
  MCLIP_CHECK_LAUNCH();
  return MCLIP_OK;
}
