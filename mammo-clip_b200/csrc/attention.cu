// BERT self-attention on the 5th-generation tensor cores (tcgen05 / TMEM / TMA), sequence length <= 256, head_dim 64.
//
// Replaces what transformers' BertSelfAttention (SDPA path, modeling_bert.py; called through text_encoder.py:47-49) computes
// per (sample, head):  S = Q K^T / 8 + key mask,  P = softmax(S),  O = dropout(P) V   and its backward
//   dV = (P o M)^T dO,  dP = (dO V^T) o M,  dS = P o (dP - delta),  delta_r = sum_c P_rc dP_rc,  dQ = dS K / 8,  dK = dS^T Q / 8
// (M = keep-mask * 1/(1-p)).  The SIMT kernels of bert.cu remain for longer sequences.
//
// Layout: every operand tile is a TMA box of the packed [tokens, 3H] qkv tensor (or the [tokens, H] dO tensor): 128 rows x 64
// columns of bf16 = 128-byte rows, 128B-swizzled.  Such a tile serves BOTH as a K-major operand (contraction over the 64
// head dimensions: Q K^T, dO V^T) and as an MN-major operand (contraction over the tokens: P V, dS K, dS^T Q, (P o M)^T dO) -- only
// the shared-memory descriptor changes.  Scores live in TMEM; one thread owns one query row (TMEM lane), reads it with
// tcgen05.ld, does the softmax arithmetic in fp32 registers and writes the bf16 probabilities back to shared memory in the
// same swizzled layout as the next MMA's A operand.  delta is the exact fp32 row sum of P o dP (not <dO, O> of the bf16-rounded
// output, which loses the cancellation in dP - delta at near-uniform attention).
#include "common.cuh"
#include "mclip_internal.h"
#include <stdlib.h>
#include <math.h>

namespace {

constexpr int AT_TILE_BYTES = 128 * 128;      // 128 rows x 64 bf16

struct AttTcDev {
  int B, L, heads, H;
  int Lp32;            // forward: keys rounded up to 32 (N of the score MMA, K of the PV MMA)
  int nkb;             // forward: 64-key blocks staged in shared memory
  int nt;              // backward: 128-token tiles
  uint32_t tmem_cols, ocol;
  const long long* amask; const uint8_t* dropmask; float drop_scale;
  bf16* out; float* lse;                 // forward
  const float* lse_in; bf16* dqkv;       // backward
};

__device__ __forceinline__ uint32_t sw128_off(int row, int chunk) { return (uint32_t)row * 128u + (uint32_t)((chunk ^ (row & 7)) << 4); }
__device__ __forceinline__ uint64_t desc_kmajor(const uint8_t* tile, int k16) { return umma_smem_desc_sw128(smem_u32(tile) + k16 * 32, 16, 1024); }
// MN-major: 64-element (128-byte) slabs along M/N `lbo` bytes apart, 8-row groups along K 1024 bytes apart, 16 K-rows per MMA
__device__ __forceinline__ uint64_t desc_mnmajor(const uint8_t* tile, int k16, uint32_t lbo) { return umma_smem_desc_sw128(smem_u32(tile) + k16 * 2048, lbo, 1024); }

// 32 fp32 values -> bf16, written as 4 swizzled 16-byte chunks of row `row` starting at column c0 of a [128 x 64]-slab tile
__device__ __forceinline__ void store_row32(uint8_t* tile, int row, int c0, const float* v) {
  uint8_t* slab = tile + (size_t)(c0 >> 6) * AT_TILE_BYTES;
  const int ch0 = (c0 & 63) >> 3;
#pragma unroll
  for (int g = 0; g < 4; ++g) {
    uint4 pk;
    pk.x = pack_bf16(v[g * 8 + 0], v[g * 8 + 1]); pk.y = pack_bf16(v[g * 8 + 2], v[g * 8 + 3]);
    pk.z = pack_bf16(v[g * 8 + 4], v[g * 8 + 5]); pk.w = pack_bf16(v[g * 8 + 6], v[g * 8 + 7]);
    *reinterpret_cast<uint4*>(slab + sw128_off(row, ch0 + g)) = pk;
  }
}

// keep-mask multipliers of 32 consecutive keys starting at kj0 (row pointer dm; nullptr: no dropout): two 16-byte loads when the
// row is 16-byte aligned (L % 16 == 0 and an aligned mask tensor), byte loads otherwise
__device__ __forceinline__ void load_mask32(const uint8_t* dm, int kj0, int L, float scale, float* mk) {
  if (!dm) {
#pragma unroll
    for (int e = 0; e < 32; ++e) mk[e] = 1.0f;
  } else if ((L & 15) == 0 && kj0 + 32 <= L && (reinterpret_cast<uintptr_t>(dm + kj0) & 15) == 0) {
    const uint4 a = *reinterpret_cast<const uint4*>(dm + kj0), b = *reinterpret_cast<const uint4*>(dm + kj0 + 16);
    const uint32_t w[8] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w};
#pragma unroll
    for (int e = 0; e < 32; ++e) mk[e] = ((w[e >> 2] >> ((e & 3) * 8)) & 0xffu) ? scale : 0.f;
  } else {
#pragma unroll
    for (int e = 0; e < 32; ++e) mk[e] = (kj0 + e < L && dm[kj0 + e]) ? scale : 0.f;
  }
}

// one accumulator row (64 fp32 columns at TMEM address taddr) * scale -> 64 bf16 at dst
__device__ __forceinline__ void store_acc_row(uint32_t taddr, float scale, bf16* dst, bool valid) {
  uint32_t o[64];
  tmem_ld64(taddr, o);
  tmem_ld_wait();
  if (valid) {
#pragma unroll
    for (int g = 0; g < 8; ++g) {
      float f[8];
#pragma unroll
      for (int e = 0; e < 8; ++e) f[e] = __uint_as_float(o[g * 8 + e]) * scale;
      stg_bf16x8(dst + g * 8, pack8(f));
    }
  }
}

// ---------------------------------------------------------------------------------------------------------------------
// forward: CTA = (128-query block, head, sample), 128 threads
// ---------------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128) mclip_att_tc_fwd_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmKV, const AttTcDev p) {
  extern __shared__ uint8_t att_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(att_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* Qs = smem;
  uint8_t* Ks = Qs + AT_TILE_BYTES;                      // [nkb*64 keys][128 B]
  uint8_t* Vs = Ks + (size_t)p.nkb * 8192;
  uint8_t* Ps = Vs + (size_t)p.nkb * 8192;               // nkb slabs of [128 queries][64 keys]
  __shared__ __align__(8) uint64_t bar_ld, bar_mma;
  __shared__ uint32_t tmem_slot;
  __shared__ float kbias[256];
  const int qb = blockIdx.x, head = blockIdx.y, b = blockIdx.z;
  const int tid = threadIdx.x, warp = tid >> 5;
  if (tid == 0) {
    tma_prefetch_desc(&tmQ); tma_prefetch_desc(&tmKV);
    mbar_init(&bar_ld, 1); mbar_init(&bar_mma, 1);
    fence_mbar_init();
  }
  for (int j = tid; j < 256; j += 128) kbias[j] = (j < p.L && p.amask[(size_t)b * p.L + j] != 0) ? 0.f : -INFINITY;
  if (warp == 0) { tmem_alloc(&tmem_slot, p.tmem_cols); tmem_relinquish(); }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = tmem_slot;
  if (tid == 0) {
    const uint32_t kv_bytes = (uint32_t)p.nkb * 8192u;
    mbar_expect_tx(&bar_ld, AT_TILE_BYTES + 2 * kv_bytes);
    tma_load_3d(Qs, &tmQ, &bar_ld, head * 64, qb * 128, b);
    tma_load_3d(Ks, &tmKV, &bar_ld, p.H + head * 64, 0, b);
    tma_load_3d(Vs, &tmKV, &bar_ld, 2 * p.H + head * 64, 0, b);
    mbar_wait(&bar_ld, 0);
    tc_fence_after();
    const uint32_t idesc_s = umma_idesc_bf16(128, p.Lp32, 0, 0);
#pragma unroll
    for (int kk = 0; kk < 4; ++kk) umma_bf16(tmem, desc_kmajor(Qs, kk), desc_kmajor(Ks, kk), idesc_s, kk != 0);
    umma_commit(&bar_mma);
  }
  mbar_wait(&bar_mma, 0);
  tc_fence_after();

  const int qi = qb * 128 + tid;
  const bool qvalid = qi < p.L;
  const uint32_t trow = tmem + ((uint32_t)(warp * 32) << 16);
  float m = -INFINITY;
  for (int c0 = 0; c0 < p.Lp32; c0 += 32) {
    uint32_t r[32];
    tmem_ld32(trow + (uint32_t)c0, r);
    tmem_ld_wait();
#pragma unroll
    for (int i = 0; i < 32; ++i) m = fmaxf(m, fmaf(__uint_as_float(r[i]), 0.125f, kbias[c0 + i]));
  }
  float lsum = 0.f;
  const uint8_t* dm = (p.dropmask && qvalid) ? p.dropmask + (((size_t)(b * p.heads + head)) * p.L + qi) * p.L : nullptr;
  for (int c0 = 0; c0 < p.Lp32; c0 += 32) {
    uint32_t r[32];
    tmem_ld32(trow + (uint32_t)c0, r);
    tmem_ld_wait();
    float v[32], mk[32];
    load_mask32(dm, c0, p.L, p.drop_scale, mk);
#pragma unroll
    for (int i = 0; i < 32; ++i) {
      const float s = fmaf(__uint_as_float(r[i]), 0.125f, kbias[c0 + i]);
      const float pv = (m == -INFINITY) ? 0.f : __expf(s - m);
      lsum += pv;
      v[i] = qvalid ? pv * mk[i] : 0.f;
    }
    store_row32(Ps, tid, c0, v);
  }
  fence_proxy_async_smem();
  tc_fence_before();
  __syncthreads();
  if (tid == 0) {
    tc_fence_after();
    const uint32_t idesc_o = umma_idesc_bf16(128, 64, 0, 1);
    const int ksteps = p.Lp32 >> 4;
    for (int ks = 0; ks < ksteps; ++ks) {
      const int kb = ks >> 2, kk = ks & 3;
      umma_bf16(tmem + p.ocol, desc_kmajor(Ps + (size_t)kb * AT_TILE_BYTES, kk), desc_mnmajor(Vs + (size_t)kb * 8192, kk, 8192), idesc_o, ks != 0);
    }
    umma_commit(&bar_mma);
  }
  mbar_wait(&bar_mma, 1);
  tc_fence_after();
  const float inv = lsum > 0.f ? 1.0f / lsum : 0.f;
  store_acc_row(trow + p.ocol, inv, p.out + ((size_t)b * p.L + (qvalid ? qi : 0)) * p.H + head * 64, qvalid);
  // log-sum-exp of the scaled, masked scores (saved for the backward pass); +inf for a fully masked row => p = 0 there
  if (p.lse && qvalid) p.lse[((size_t)(b * p.heads + head)) * p.L + qi] = lsum > 0.f ? m + __logf(lsum) : INFINITY;
  tc_fence_before();
  __syncthreads();
  if (warp == 0) { tc_fence_after(); tmem_dealloc(tmem, p.tmem_cols); }
}

// ---------------------------------------------------------------------------------------------------------------------
// backward: CTA = (head, sample), 128 threads; nt = ceil(L / 128) token tiles (1 or 2)
//   TMEM columns: S [0,128)  dPM [128,256)  dQ_i [256 + 64 i, +64)  dK_j [384,448)  dV_j [448,512)
// ---------------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128) mclip_att_tc_bwd_kernel(const __grid_constant__ CUtensorMap tmQKV, const __grid_constant__ CUtensorMap tmDO, const AttTcDev p) {
  extern __shared__ uint8_t att_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(att_raw) + 1023) & ~uintptr_t(1023));
  const int nt = p.nt;
  uint8_t* Qs = smem;                                    // [nt] tiles each
  uint8_t* Ks = Qs + (size_t)nt * AT_TILE_BYTES;
  uint8_t* Vs = Ks + (size_t)nt * AT_TILE_BYTES;
  uint8_t* Gs = Vs + (size_t)nt * AT_TILE_BYTES;         // dO
  uint8_t* PMs = Gs + (size_t)nt * AT_TILE_BYTES;        // (P o M): two 64-key slabs of [128 queries][128 B]
  uint8_t* dSs = PMs + 2 * AT_TILE_BYTES;                // dS / 8, same layout
  __shared__ __align__(8) uint64_t bar_ld, bar_mma;
  __shared__ uint32_t tmem_slot;
  __shared__ float kbias[256];
  const int head = blockIdx.x, b = blockIdx.y;
  const int tid = threadIdx.x, warp = tid >> 5;
  const int L = p.L, H = p.H;
  if (tid == 0) {
    tma_prefetch_desc(&tmQKV); tma_prefetch_desc(&tmDO);
    mbar_init(&bar_ld, 1); mbar_init(&bar_mma, 1);
    fence_mbar_init();
  }
  for (int j = tid; j < 256; j += 128) kbias[j] = (j < L && p.amask[(size_t)b * L + j] != 0) ? 0.f : -INFINITY;
  if (warp == 0) { tmem_alloc(&tmem_slot, 512); tmem_relinquish(); }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = tmem_slot;
  if (tid == 0) {
    mbar_expect_tx(&bar_ld, (uint32_t)(4 * nt) * AT_TILE_BYTES);
    for (int t = 0; t < nt; ++t) {
      tma_load_3d(Qs + (size_t)t * AT_TILE_BYTES, &tmQKV, &bar_ld, head * 64, t * 128, b);
      tma_load_3d(Ks + (size_t)t * AT_TILE_BYTES, &tmQKV, &bar_ld, H + head * 64, t * 128, b);
      tma_load_3d(Vs + (size_t)t * AT_TILE_BYTES, &tmQKV, &bar_ld, 2 * H + head * 64, t * 128, b);
      tma_load_3d(Gs + (size_t)t * AT_TILE_BYTES, &tmDO, &bar_ld, head * 64, t * 128, b);
    }
    mbar_wait(&bar_ld, 0);
  }
  uint32_t mma_phase = 0;
  const uint32_t trow = tmem + ((uint32_t)(warp * 32) << 16);
  const size_t rowbase = ((size_t)(b * p.heads + head)) * L;
  float lse_r[2], delta_r[2] = {0.f, 0.f};
#pragma unroll
  for (int i = 0; i < 2; ++i) { const int qi = i * 128 + tid; lse_r[i] = (i < nt && qi < L) ? p.lse_in[rowbase + qi] : INFINITY; }

  // scores S(i,j) = Q_i K_j^T and dPM(i,j) = dO_i V_j^T into TMEM; every thread returns once they are readable
  auto score_mmas = [&](int i, int j, int kt32) {
    if (tid == 0) {
      tc_fence_after();
      const uint32_t idesc = umma_idesc_bf16(128, kt32, 0, 0);
#pragma unroll
      for (int kk = 0; kk < 4; ++kk)
        umma_bf16(tmem, desc_kmajor(Qs + (size_t)i * AT_TILE_BYTES, kk), desc_kmajor(Ks + (size_t)j * AT_TILE_BYTES, kk), idesc, kk != 0);
#pragma unroll
      for (int kk = 0; kk < 4; ++kk)
        umma_bf16(tmem + 128, desc_kmajor(Gs + (size_t)i * AT_TILE_BYTES, kk), desc_kmajor(Vs + (size_t)j * AT_TILE_BYTES, kk), idesc, kk != 0);
      umma_commit(&bar_mma);
    }
    mbar_wait(&bar_mma, mma_phase);
    mma_phase ^= 1;
    tc_fence_after();
  };

  // ---- phase A: delta_i = sum over all keys of (P o M) o dPM ----
  for (int i = 0; i < nt; ++i) {
    const int qi = i * 128 + tid;
    const uint8_t* dm = (p.dropmask && qi < L) ? p.dropmask + (rowbase + qi) * L : nullptr;
    float acc = 0.f;
    for (int j = 0; j < nt; ++j) {
      const int kt = min(128, L - j * 128), kt32 = (kt + 31) & ~31;
      score_mmas(i, j, kt32);
      for (int c0 = 0; c0 < kt32; c0 += 32) {
        uint32_t rs[32], rg[32];
        tmem_ld32(trow + (uint32_t)c0, rs);
        tmem_ld32(trow + 128u + (uint32_t)c0, rg);
        float mk[32];
        load_mask32(dm, j * 128 + c0, L, p.drop_scale, mk);
        tmem_ld_wait();
#pragma unroll
        for (int e = 0; e < 32; ++e) {
          const int kj = j * 128 + c0 + e;
          const float pr = __expf(fmaf(__uint_as_float(rs[e]), 0.125f, kbias[kj]) - lse_r[i]);
          acc = fmaf(pr * mk[e], __uint_as_float(rg[e]), acc);
        }
      }
      tc_fence_before();
      __syncthreads();                                 // every row has been read: the next MMAs may overwrite S / dPM
    }
    delta_r[i] = acc;
  }

  // ---- phase B: key tile j outer (dK_j, dV_j accumulate over the query tiles), dQ_i accumulate over j ----
  const uint32_t idesc_tt = umma_idesc_bf16(128, 64, 1, 1), idesc_kt = umma_idesc_bf16(128, 64, 0, 1);
  for (int j = 0; j < nt; ++j) {
    const int kt = min(128, L - j * 128), kt32 = (kt + 31) & ~31;
    for (int i = 0; i < nt; ++i) {
      const int qi = i * 128 + tid;
      const uint8_t* dm = (p.dropmask && qi < L) ? p.dropmask + (rowbase + qi) * L : nullptr;
      score_mmas(i, j, kt32);
      for (int c0 = 0; c0 < 128; c0 += 32) {
        float pm[32], ds[32];
        if (c0 < kt32) {
          uint32_t rs[32], rg[32];
          tmem_ld32(trow + (uint32_t)c0, rs);
          tmem_ld32(trow + 128u + (uint32_t)c0, rg);
          float mk[32];
          load_mask32(dm, j * 128 + c0, L, p.drop_scale, mk);
          tmem_ld_wait();
#pragma unroll
          for (int e = 0; e < 32; ++e) {
            const int kj = j * 128 + c0 + e;
            const float pr = __expf(fmaf(__uint_as_float(rs[e]), 0.125f, kbias[kj]) - lse_r[i]);
            pm[e] = pr * mk[e];
            ds[e] = pr * (__uint_as_float(rg[e]) * mk[e] - delta_r[i]) * 0.125f;
          }
        } else {                                         // key columns past this tile: the token-contraction MMAs read all 128
#pragma unroll
          for (int e = 0; e < 32; ++e) { pm[e] = 0.f; ds[e] = 0.f; }
        }
        store_row32(PMs, tid, c0, pm);
        store_row32(dSs, tid, c0, ds);
      }
      fence_proxy_async_smem();
      tc_fence_before();
      __syncthreads();
      if (tid == 0) {
        tc_fence_after();
        const uint8_t* Qi = Qs + (size_t)i * AT_TILE_BYTES; const uint8_t* Gi = Gs + (size_t)i * AT_TILE_BYTES;
        const uint8_t* Kj = Ks + (size_t)j * AT_TILE_BYTES;
#pragma unroll
        for (int kk = 0; kk < 8; ++kk) {                 // contraction over the 128 queries of tile i
          umma_bf16(tmem + 448, desc_mnmajor(PMs, kk, AT_TILE_BYTES), desc_mnmajor(Gi, kk, 8192), idesc_tt, (i | kk) != 0);
          umma_bf16(tmem + 384, desc_mnmajor(dSs, kk, AT_TILE_BYTES), desc_mnmajor(Qi, kk, 8192), idesc_tt, (i | kk) != 0);
        }
#pragma unroll
        for (int ks = 0; ks < 8; ++ks) {                 // contraction over the 128 keys of tile j (zero past kt)
          const int kb = ks >> 2, kk = ks & 3;
          umma_bf16(tmem + 256 + 64 * i, desc_kmajor(dSs + (size_t)kb * AT_TILE_BYTES, kk), desc_mnmajor(Kj + (size_t)kb * 8192, kk, 8192), idesc_kt,
                    (j | ks) != 0);
        }
        umma_commit(&bar_mma);
      }
      mbar_wait(&bar_mma, mma_phase);                    // PMs / dSs may be rewritten, S / dPM overwritten
      mma_phase ^= 1;
      tc_fence_after();
    }
    {
      const int kj = j * 128 + tid;
      bf16* base = p.dqkv + ((size_t)b * L + (kj < L ? kj : 0)) * 3 * H + head * 64;
      store_acc_row(trow + 384, 1.0f, base + H, kj < L);
      store_acc_row(trow + 448, 1.0f, base + 2 * H, kj < L);
    }
    tc_fence_before();
    __syncthreads();
  }
  for (int i = 0; i < nt; ++i) {
    const int qi = i * 128 + tid;
    store_acc_row(trow + 256 + 64 * i, 1.0f, p.dqkv + ((size_t)b * L + (qi < L ? qi : 0)) * 3 * H + head * 64, qi < L);
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) { tc_fence_after(); tmem_dealloc(tmem, 512); }
}

int att_tmap(CUtensorMap* m, const void* ptr, int width, int L, int B, int box_rows) {
  const unsigned long long dims[3] = {(unsigned long long)width, (unsigned long long)L, (unsigned long long)B};
  const unsigned long long strides[2] = {(unsigned long long)width * 2, (unsigned long long)L * width * 2};
  const unsigned box[3] = {64u, (unsigned)box_rows, 1u};
  return mclip_tmap_encode_bf16(m, ptr, 3, dims, strides, box, 1);
}

}  // namespace

bool mclip_att_tc_covers(int seq_len, int heads, int head_dim) {
  static int enabled = -1;
  if (enabled < 0) { const char* e = getenv("MCLIP_ATT_TC"); enabled = e ? atoi(e) : 1; }
  return enabled && head_dim == 64 && seq_len >= 1 && seq_len <= 256 && heads >= 1;
}

int mclip_att_tc_forward(const void* qkv, const void* amask, const void* dropmask, float drop_scale, void* out, float* lse, int batch, int seq_len,
                         int heads, void* stream) {
  AttTcDev p;
  memset(&p, 0, sizeof(p));
  p.B = batch; p.L = seq_len; p.heads = heads; p.H = heads * 64;
  p.Lp32 = (seq_len + 31) & ~31;
  p.nkb = ceil_div(seq_len, 64);
  p.ocol = (uint32_t)((p.Lp32 + 63) & ~63);
  const uint32_t need = p.ocol + 64;
  p.tmem_cols = need <= 128 ? 128 : need <= 256 ? 256 : 512;
  p.amask = (const long long*)amask; p.dropmask = (const uint8_t*)dropmask; p.drop_scale = drop_scale;
  p.out = (bf16*)out; p.lse = lse;
  CUtensorMap tmQ, tmKV;
  int rc = att_tmap(&tmQ, qkv, 3 * p.H, seq_len, batch, 128);
  if (rc) return rc;
  if ((rc = att_tmap(&tmKV, qkv, 3 * p.H, seq_len, batch, p.nkb * 64))) return rc;
  const int smem = 1024 + AT_TILE_BYTES + p.nkb * (2 * 8192 + AT_TILE_BYTES);
  static int attr = 0;
  if (!attr) { MCLIP_CHECK_CUDA(cudaFuncSetAttribute(mclip_att_tc_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 1024 + AT_TILE_BYTES + 4 * (2 * 8192 + AT_TILE_BYTES))); attr = 1; }
  dim3 grid(ceil_div(seq_len, 128), heads, batch);
  mclip_att_tc_fwd_kernel<<<grid, 128, smem, (cudaStream_t)stream>>>(tmQ, tmKV, p);
  MCLIP_CHECK_LAUNCH();
  return MCLIP_OK;
}

int mclip_att_tc_backward(const void* qkv, const void* d_out, const float* lse, const void* amask, const void* dropmask, float drop_scale, void* dqkv,
                          int batch, int seq_len, int heads, void* stream) {
  AttTcDev p;
  memset(&p, 0, sizeof(p));
  p.B = batch; p.L = seq_len; p.heads = heads; p.H = heads * 64;
  p.nt = ceil_div(seq_len, 128);
  p.amask = (const long long*)amask; p.dropmask = (const uint8_t*)dropmask; p.drop_scale = drop_scale;
  p.lse_in = lse; p.dqkv = (bf16*)dqkv;
  CUtensorMap tmQKV, tmDO;
  int rc = att_tmap(&tmQKV, qkv, 3 * p.H, seq_len, batch, 128);
  if (rc) return rc;
  if ((rc = att_tmap(&tmDO, d_out, p.H, seq_len, batch, 128))) return rc;
  const int smem = 1024 + (4 * p.nt + 4) * AT_TILE_BYTES;
  static int attr = 0;
  if (!attr) { MCLIP_CHECK_CUDA(cudaFuncSetAttribute(mclip_att_tc_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 1024 + 12 * AT_TILE_BYTES)); attr = 1; }
  dim3 grid(heads, batch);
  mclip_att_tc_bwd_kernel<<<grid, 128, smem, (cudaStream_t)stream>>>(tmQKV, tmDO, p);
  MCLIP_CHECK_LAUNCH();
  return MCLIP_OK;
}
