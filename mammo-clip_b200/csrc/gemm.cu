// tcgen05 / TMEM / TMA GEMMs for sm_100a.
//
//  mclip_gemm_tn   : D[b,m,n] = epi( sum_k A[b,m,k] * B[b|0,n,k] )     both operands K-major (row-major [rows,K]).
//                    The 1x1 "pointwise" convolutions of MBConv (efficientnet_custom.py:105,122,283) in NHWC are exactly
//                    this GEMM with m = pixel, k = Cin, n = Cout; their data-gradients are the same GEMM with the
//                    transposed weight; BERT's Linear layers (transformers modeling_bert.py) likewise.
//                    Epilogue: +bias[n], erf-GELU, +residual[m,n], bf16 store via swizzled smem + TMA, and per-column
//                    sum / sum-of-squares partials (train-mode BatchNorm statistics, efficientnet_custom.py:106,123).
//  mclip_gemm_wgrad: D[i,j] = sum_r A[r,i] * B[r,j]  both operands MN-major (the reduction runs over pixels r):
//                    weight gradients of the 1x1 convolutions / Linear layers, split over r across CTAs, fp32 partials.
//
// Structure (both): persistent CTAs, warp 0 = TMA producer, warp 1 = TMEM allocator + single-thread MMA issuer,
// warps 2.. = epilogue (tcgen05.ld -> registers -> swizzled smem -> TMA store), smem ring of `stages` operand tiles
// guarded by full/empty mbarriers, two TMEM accumulator stages guarded by tmem_full/tmem_empty mbarriers.
#include "common.cuh"
#include "mclip_internal.h"
#include <cudaTypedefs.h>
#include <stdlib.h>

#define GEMM_EPI_WARPS 8              // two 64-column slabs per warp and tile, double-buffered staging
#define GEMM_THREADS (64 + 32 * GEMM_EPI_WARPS)
#define GEMM_STAGING_BYTES (GEMM_EPI_WARPS * 2 * 4096)
#define GEMM_BM 128
#define GEMM_BK 64
#define GEMM_SLAB_BYTES 4096          // 32 rows x 64 bf16, 128B-swizzled
#define GEMM_SMEM_LIMIT (227 * 1024)

struct GemmDev {
  int M, N, K, batches, b_batched;
  int m_blocks, n_blocks, block_n, k_blocks, stages, m_tiles_total;
  uint32_t idesc;
  const float* bias;
  const bf16* residual;
  long long res_ld, res_bs;
  int act;
  float* stats;     // [(gridDim.x / n_blocks) * 4][2][N]
  const uint8_t* dropmask; float drop_scale;
  bf16* aux; long long aux_ld;   // optional bf16 copy of the pre-activation value (after the bias), [batches*M, aux_ld]
  int K1, k1_blocks;   // K-concatenated A operand: columns [0,K1) come from tmA, the rest (K - K1pad... see host) from tmA2; k1_blocks = ceil(K1/64)
  int K2;              // columns of the second A operand (0: single operand, K1 == K)
  int small_k;      // K <= 64 and shared B: B panel resident in smem, A tiles staged with cp.async by warp 0
  const bf16* a_ptr; long long lda, a_bs;
  int debug;        // MCLIP_GEMM_DEBUG bitmask (experiments only): 1 no stats, 2 no TMA store, 4 no TMEM load/convert
};

// ------------------------------------------------------------------------------------------------
// host: tensor maps
// ------------------------------------------------------------------------------------------------
static PFN_cuTensorMapEncodeTiled_v12000 get_encode() {
  static PFN_cuTensorMapEncodeTiled_v12000 fn = nullptr;
  if (!fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess)
      fn = (PFN_cuTensorMapEncodeTiled_v12000)p;
  }
  return fn;
}

// bf16 tensor, dims[0] innermost.  strides in ELEMENTS for dims 1 and 2.
static int make_tmap_bf16_3d(CUtensorMap* m, const void* ptr, unsigned long long d0, unsigned long long d1, unsigned long long d2,
                             unsigned long long s1, unsigned long long s2, unsigned b0, unsigned b1, unsigned b2) {
  PFN_cuTensorMapEncodeTiled_v12000 enc = get_encode();
  if (!enc) { mclip_set_error("cuTensorMapEncodeTiled not available from the driver"); return MCLIP_ERR_CUDA; }
  if (((uintptr_t)ptr & 15) || (s1 * 2) % 16 || (d2 > 1 && (s2 * 2) % 16)) {
    mclip_set_error("TMA operand must be 16B aligned with 16B-multiple strides (ptr=%p s1=%llu s2=%llu)", ptr, s1, s2);
    return MCLIP_ERR_INVALID;
  }
  cuuint64_t dims[3] = {d0, d1, d2};
  cuuint64_t strides[2] = {s1 * 2, (d2 > 1 ? s2 : d1 * s1) * 2};
  cuuint32_t box[3] = {b0, b1, b2};
  cuuint32_t es[3] = {1, 1, 1};
  CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, const_cast<void*>(ptr), dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                   CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    mclip_set_error("cuTensorMapEncodeTiled failed (%d): dims %llu,%llu,%llu strides %llu,%llu box %u,%u,%u", (int)r, d0, d1, d2, s1, s2, b0, b1, b2);
    return MCLIP_ERR_CUDA;
  }
  return MCLIP_OK;
}

// ------------------------------------------------------------------------------------------------
// device: K-major GEMM
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ float gelu_erf(float v) { return 0.5f * v * (1.0f + erff(v * 0.70710678118654752f)); }

// Epilogue instantiations.  MODE 0: plain (the MBConv forward convs and most data gradients); 1: + residual (the
// data gradient of a block with a skip connection); 2: everything (bias, pre-activation copy, erf-GELU, dropout mask,
// residual: the BERT / head Linears).  STATS: BatchNorm (sum, sum sq) partials.  The epilogue is straight-line code that
// every warp runs once per tile, so its SIZE is its cost: the all-in-one, fully unrolled version (~5000 SASS
// instructions) stalled on instruction fetch (smsp "no_instruction" ~0.9 per issue, profiles/r01c_ncu_gemm_expand_blk4)
// and ran the small-K convolutions at 2-3 TB/s; compiling the unused paths out gives 3.5-5.  (Rolling the column loop into two
// 32-column trips shrinks the code further but pays a second TMEM-load wait per slab: measured slower, not kept.)
template <int MODE, bool STATS>
__global__ void __launch_bounds__(GEMM_THREADS, 1)
mclip_gemm_tn_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                     const __grid_constant__ CUtensorMap tmD, const __grid_constant__ CUtensorMap tmA2, const GemmDev p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  const uint32_t a_bytes = GEMM_BM * GEMM_BK * 2, b_bytes = (uint32_t)p.block_n * GEMM_BK * 2;
  // normal mode: every stage holds an A and a B tile; small-K mode: stages hold A only, one resident B panel follows them
  const uint32_t stage_bytes = p.small_k ? a_bytes : a_bytes + b_bytes;
  uint8_t* bpanel = smem + (size_t)p.stages * stage_bytes;
  uint8_t* dstage = bpanel + (p.small_k ? b_bytes : 0);
  uint64_t* bars = reinterpret_cast<uint64_t*>(dstage + GEMM_STAGING_BYTES);

  uint64_t* full = bars;
  uint64_t* empty = bars + p.stages;
  uint64_t* tfull = bars + 2 * p.stages;
  uint64_t* tempty = tfull + 4;
  uint64_t* bfull = tempty + 4;
  // TMEM accumulator stages: 2 x 256 columns; 4 x 128 when a tile is one 64-column slab (block_n <= 64).  Such tiles are tiny
  // (K <= 64 x N <= 64): with two stages the MMA -> commit -> epilogue -> release round trip (~1 us) bounded the whole kernel.
  const int nst = p.block_n <= 64 ? 4 : 2;
  const uint32_t st_cols = 512u / (uint32_t)nst;
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(bfull + 1);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmA); tma_prefetch_desc(&tmB); tma_prefetch_desc(&tmD);
    for (int s = 0; s < p.stages; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], 1); }
    // one 64-column slab per tile (block_n <= 64): the two epilogue warp groups take ALTERNATE tiles (= alternate TMEM stages)
    // instead of one group idling, so a stage is released by 4 warps
    for (int a = 0; a < 4; ++a) { mbar_init(&tfull[a], 1); mbar_init(&tempty[a], p.block_n <= 64 ? GEMM_EPI_WARPS / 2 : GEMM_EPI_WARPS); }
    mbar_init(bfull, 1);
    fence_mbar_init();
  }
  if (warp == 1) { tmem_alloc(tmem_ptr, 512); tmem_relinquish(); }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr;

  // tile assignment: n block fixed per CTA, m tiles strided (keeps the BN partial accumulators column-stable)
  const int n_blk = blockIdx.x % p.n_blocks;
  const int mt0 = blockIdx.x / p.n_blocks, mt_step = gridDim.x / p.n_blocks;
  const int n0 = n_blk * p.block_n;

  if (warp == 0 && p.small_k) {
    // ---- small-K producer (whole warp): B panel once by TMA; A tiles by 16-byte cp.async into the 128B-swizzled K-major
    //      layout (TMA would issue one 48..128-byte request per row, which costs more than the data itself) ----
    if (lane == 0) {
      mbar_expect_tx(bfull, b_bytes);
      tma_load_3d(bpanel, &tmB, bfull, 0, n0, 0);
    }
    const int kch = (p.K * 2) >> 4;                       // valid 16-byte chunks per row (K % 8 == 0)
    const int rch = ((p.K + 15) >> 4) * 2;                // chunks the MMAs read (K rounded up to 16)
    for (int st = 0; st < p.stages; ++st)                 // zero the K padding once; cp.async never touches it again
      for (int idx = lane; idx < GEMM_BM * (rch - kch); idx += 32) {
        const int r = idx / (rch - kch), c = kch + idx % (rch - kch);
        *reinterpret_cast<uint4*>(smem + (size_t)st * stage_bytes + r * 128 + ((c ^ (r & 7)) << 4)) = make_uint4(0u, 0u, 0u, 0u);
      }
    fence_proxy_async_smem();
    __syncwarp();
    constexpr int LOOK = 4;                               // tiles in flight behind the one being issued (stages >= LOOK + 2)
    // This warp's instruction stream IS the critical path of the tiny-K GEMMs (measured: 291 instructions per tile at ~5.5 cycles
    // each = the whole 1 us tile time, MMA and epilogue warps idle): everything tile-invariant is hoisted -- per-lane row offsets
    // (global and shared), the swizzle term ((r & 7) == (lane & 7) for all four rows of a lane), the (batch, m block) walk without
    // divisions -- and full tiles use the unpredicated copy.
    const uint32_t swz = (uint32_t)(lane & 7);
    const size_t row_bytes = (size_t)p.lda * 2;
    const char* const a_bytes_ptr = reinterpret_cast<const char*>(p.a_ptr);
    size_t roff[GEMM_BM / 32];
    uint32_t doff[GEMM_BM / 32];
#pragma unroll
    for (int rr = 0; rr < GEMM_BM / 32; ++rr) { roff[rr] = (size_t)(lane + rr * 32) * row_bytes; doff[rr] = (uint32_t)(lane + rr * 32) * 128u; }
    const uint32_t smem0 = smem_u32(smem);
    int b = mt0 / p.m_blocks, mb = mt0 % p.m_blocks;      // one division per kernel
    int it = 0, st = 0;
    uint32_t ph = 0;
    for (int mt = mt0; mt < p.m_tiles_total; mt += mt_step, ++it) {
      mbar_wait(&empty[st], ph ^ 1);
      const int m0 = mb * GEMM_BM;
      const char* tile = a_bytes_ptr + ((size_t)b * p.a_bs + (size_t)m0 * p.lda) * 2;
      const uint32_t sdst = smem0 + (uint32_t)st * stage_bytes;
      const int rows = p.M - m0;
      if (rows >= GEMM_BM) {
#pragma unroll
        for (int rr = 0; rr < GEMM_BM / 32; ++rr) {
          const char* src = tile + roff[rr];
          const uint32_t drow = sdst + doff[rr];
          for (int c = 0; c < kch; ++c)
            asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(drow + (((uint32_t)c ^ swz) << 4)), "l"(src + c * 16) : "memory");
        }
      } else {
#pragma unroll
        for (int rr = 0; rr < GEMM_BM / 32; ++rr) {
          const int r = lane + rr * 32;
          const uint32_t nbytes = r < rows ? 16u : 0u;    // rows past M are zero-filled
          const char* src = tile + (r < rows ? roff[rr] : 0);
          const uint32_t drow = sdst + doff[rr];
          for (int c = 0; c < kch; ++c)
            asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(drow + (((uint32_t)c ^ swz) << 4)), "l"(src + c * 16), "r"(nbytes) : "memory");
        }
      }
      asm volatile("cp.async.commit_group;" ::: "memory");
      if (it >= LOOK) {
        asm volatile("cp.async.wait_group %0;" ::"n"(LOOK) : "memory");
        fence_proxy_async_smem();
        __syncwarp();
        int fs = st - LOOK; if (fs < 0) fs += p.stages;
        if (lane == 0) mbar_arrive(&full[fs]);
      }
      if (++st == p.stages) { st = 0; ph ^= 1; }
      mb += mt_step;
      while (mb >= p.m_blocks) { mb -= p.m_blocks; ++b; }
    }
    // drain
    for (int d = (it < LOOK ? it : LOOK); d > 0; --d) {
      if (d == 4) asm volatile("cp.async.wait_group 3;" ::: "memory");
      else if (d == 3) asm volatile("cp.async.wait_group 2;" ::: "memory");
      else if (d == 2) asm volatile("cp.async.wait_group 1;" ::: "memory");
      else asm volatile("cp.async.wait_group 0;" ::: "memory");
      fence_proxy_async_smem();
      __syncwarp();
      if (lane == 0) mbar_arrive(&full[(it - d) % p.stages]);
    }
  } else if (warp == 0) {
    if (lane == 0) {
      int stage = 0; uint32_t phase = 0;
      for (int mt = mt0; mt < p.m_tiles_total; mt += mt_step) {
        const int b = mt / p.m_blocks, m0 = (mt % p.m_blocks) * GEMM_BM;
        for (int kb = 0; kb < p.k_blocks; ++kb) {
          mbar_wait(&empty[stage], phase ^ 1);
          uint8_t* sa = smem + (size_t)stage * stage_bytes;
          mbar_expect_tx(&full[stage], stage_bytes);
          if (kb < p.k1_blocks) tma_load_3d(sa, &tmA, &full[stage], kb * GEMM_BK, m0, b);
          else tma_load_3d(sa, &tmA2, &full[stage], (kb - p.k1_blocks) * GEMM_BK, m0, b);     // second operand of a K-concatenated A
          tma_load_3d(sa + a_bytes, &tmB, &full[stage], kb * GEMM_BK, n0, p.b_batched ? b : 0);
          if (++stage == p.stages) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      int stage = 0; uint32_t phase = 0; int as = 0; uint32_t aphase = 0;
      if (p.small_k) mbar_wait(bfull, 0);
      for (int mt = mt0; mt < p.m_tiles_total; mt += mt_step) {
        mbar_wait(&tempty[as], aphase ^ 1);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + (uint32_t)as * st_cols;
        for (int kb = 0; kb < p.k_blocks; ++kb) {
          mbar_wait(&full[stage], phase);
          tc_fence_after();
          const uint32_t sa = smem_u32(smem + (size_t)stage * stage_bytes), sb = p.small_k ? smem_u32(bpanel) : sa + a_bytes;
          const int krem = kb < p.k1_blocks ? p.K1 - kb * GEMM_BK : p.K2 - (kb - p.k1_blocks) * GEMM_BK;
          const int nk = krem >= GEMM_BK ? 4 : (krem + 15) >> 4;
          for (int kk = 0; kk < nk; ++kk)
            umma_bf16(d_tmem, umma_smem_desc_sw128(sa + kk * 32, 16, 1024), umma_smem_desc_sw128(sb + kk * 32, 16, 1024), p.idesc,
                      (kb | kk) != 0);
          umma_commit(&empty[stage]);
          if (++stage == p.stages) { stage = 0; phase ^= 1; }
        }
        umma_commit(&tfull[as]);
        if (++as == nst) { as = 0; aphase ^= 1; }
      }
    }
  } else {
    // ---------------- epilogue warps ----------------
    const int ew = warp - 2;
    const int q = warp & 3;                 // TMEM lane quarter this warp may read
    const int h = ew >> 2;                  // which half of the 64-column slabs
    const int nslabs = (p.block_n + 63) >> 6;
    uint8_t* my_buf = dstage + (size_t)ew * 2 * GEMM_SLAB_BYTES;
    float st_sum[2][2] = {{0.f, 0.f}, {0.f, 0.f}}, st_sq[2][2] = {{0.f, 0.f}, {0.f, 0.f}};
    const bool split = nslabs == 1;
    int buf = 0, it = 0;
    int b = mt0 / p.m_blocks, mb = mt0 % p.m_blocks - mt_step;          // (batch, m block) walk without per-tile divisions
    for (int mt = mt0; mt < p.m_tiles_total; mt += mt_step, ++it) {
      mb += mt_step;
      while (mb >= p.m_blocks) { mb -= p.m_blocks; ++b; }
      if (split && (it & 1) != h) continue;
      const int as = it & (nst - 1);                                    // nst is 2 or 4
      const uint32_t aphase = (uint32_t)(it >> (nst == 4 ? 2 : 1)) & 1u;
      const int m0 = mb * GEMM_BM;
      const int row = m0 + q * 32 + lane;
      const int nvalid = min(32, max(0, p.M - (m0 + q * 32)));
      mbar_wait(&tfull[as], aphase);
      tc_fence_after();
#pragma unroll
      for (int si = 0; si < 2; ++si) {
        const int s = split ? si : h + 2 * si;
        if (s >= nslabs) break;
        const int c0 = n0 + s * 64;
        if (c0 >= p.N) break;
        uint8_t* sb = my_buf + (size_t)buf * GEMM_SLAB_BYTES;
        if (lane == 0) tma_store_wait_read1();     // the store issued two slabs ago (same buffer) has drained
        __syncwarp();
        const uint32_t taddr = tmem_base + (uint32_t)as * st_cols + (uint32_t)(s * 64) + ((uint32_t)(q * 32) << 16);
        uint32_t r[64];                                       // one 64-column TMEM load and ONE wait per slab
        if (!(p.debug & 4)) tmem_ld64(taddr, r);
        tmem_ld_wait();
#pragma unroll
        for (int ch = 0; ch < 4; ++ch) {
          if (s * 64 + ch * 16 >= p.block_n) break;           // warp-uniform: columns past block_n hold no result
          float v[16];
#pragma unroll
          for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(r[ch * 16 + i]);
          const int cc = c0 + ch * 16;
          if (MODE == 2) {
            if (p.bias) {
#pragma unroll
              for (int i = 0; i < 16; ++i) if (cc + i < p.N) v[i] += __ldg(p.bias + cc + i);
            }
            if (p.aux && row < p.M) {
              bf16* ap = p.aux + ((size_t)b * p.M + row) * p.aux_ld + cc;
#pragma unroll
              for (int g = 0; g < 2; ++g)
                if (cc + g * 8 < p.N) stg_bf16x8(ap + g * 8, pack8(v + g * 8));
            }
            if (p.act == 1) {
#pragma unroll
              for (int i = 0; i < 16; ++i) v[i] = gelu_erf(v[i]);
            }
            if (p.dropmask && row < p.M && cc < p.N) {
              const uint4 mk = *reinterpret_cast<const uint4*>(p.dropmask + ((size_t)b * p.M + row) * p.N + cc);   // N % 16 == 0 required
              const uint32_t mw[4] = {mk.x, mk.y, mk.z, mk.w};
#pragma unroll
              for (int i = 0; i < 16; ++i) v[i] *= ((mw[i >> 2] >> ((i & 3) * 8)) & 0xffu) ? p.drop_scale : 0.f;
            }
          }
          if (MODE == 3) {                          // lean "+ bias (+ residual)" epilogue: the folded BN-backward data gradient
#pragma unroll
            for (int i = 0; i < 16; ++i) if (cc + i < p.N) v[i] += __ldg(p.bias + cc + i);
          }
          if (MODE >= 1 && p.residual && row < p.M) {
            const bf16* rp = p.residual + (size_t)b * p.res_bs + (size_t)row * p.res_ld + cc;
#pragma unroll
            for (int g = 0; g < 2; ++g) {
              if (cc + g * 8 < p.N) {                // N is a multiple of 8
                bf16x8 rv = ldg_bf16x8(rp + g * 8);
                float f[8]; unpack8(rv, f);
#pragma unroll
                for (int i = 0; i < 8; ++i) v[g * 8 + i] += f[i];
              }
            }
          }
#pragma unroll
          for (int g = 0; g < 2; ++g) {
            const int chunk = ch * 2 + g;
            uint4 pk;
            pk.x = pack_bf16(v[g * 8 + 0], v[g * 8 + 1]); pk.y = pack_bf16(v[g * 8 + 2], v[g * 8 + 3]);
            pk.z = pack_bf16(v[g * 8 + 4], v[g * 8 + 5]); pk.w = pack_bf16(v[g * 8 + 6], v[g * 8 + 7]);
            *reinterpret_cast<uint4*>(sb + lane * 128 + ((chunk ^ (lane & 7)) << 4)) = pk;
          }
        }
        fence_proxy_async_smem();
        __syncwarp();
        if (lane == 0 && !(p.debug & 2)) { tma_store_3d(&tmD, sb, c0, m0 + q * 32, b); tma_store_commit(); }
        if (STATS && (MODE != 2 || p.stats)) {
          // column sums over this warp's valid rows, read back from the swizzled slab (conflict-free); rows in order
          float s0 = 0.f, s1 = 0.f, q0 = 0.f, q1 = 0.f;
          const uint8_t* colp = sb + (lane & 3) * 4;
          const int cq = lane >> 2;
          if (nvalid == 32) {
#pragma unroll
            for (int r2 = 0; r2 < 32; ++r2) {
              uint32_t w = *reinterpret_cast<const uint32_t*>(colp + r2 * 128 + ((cq ^ (r2 & 7)) << 4));
              float a0 = bf16_lo(w), a1 = bf16_hi(w);
              s0 += a0; s1 += a1; q0 = fmaf(a0, a0, q0); q1 = fmaf(a1, a1, q1);
            }
          } else {
            for (int r2 = 0; r2 < nvalid; ++r2) {
              uint32_t w = *reinterpret_cast<const uint32_t*>(colp + r2 * 128 + ((cq ^ (r2 & 7)) << 4));
              float a0 = bf16_lo(w), a1 = bf16_hi(w);
              s0 += a0; s1 += a1; q0 = fmaf(a0, a0, q0); q1 = fmaf(a1, a1, q1);
            }
          }
          st_sum[si][0] += s0; st_sum[si][1] += s1; st_sq[si][0] += q0; st_sq[si][1] += q1;
        }
        buf ^= 1;
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&tempty[as]);
    }
    if (lane == 0) tma_store_wait_all0();
    if (STATS && (MODE != 2 || p.stats)) {
      if (split) {                     // the h = 1 group hands its partial sums of the same columns to the h = 0 group (fixed order)
        float* xch = reinterpret_cast<float*>(dstage) + (size_t)q * 128;
        named_bar_sync(1, 32 * GEMM_EPI_WARPS);                                 // staging is idle: EVERY warp's TMA stores have drained
        if (h == 1) { xch[lane * 4 + 0] = st_sum[0][0]; xch[lane * 4 + 1] = st_sum[0][1]; xch[lane * 4 + 2] = st_sq[0][0]; xch[lane * 4 + 3] = st_sq[0][1]; }
        named_bar_sync(1, 32 * GEMM_EPI_WARPS);
        if (h == 0) { st_sum[0][0] += xch[lane * 4 + 0]; st_sum[0][1] += xch[lane * 4 + 1]; st_sq[0][0] += xch[lane * 4 + 2]; st_sq[0][1] += xch[lane * 4 + 3]; }
      }
      const int slot = (blockIdx.x / p.n_blocks) * 4 + q;
#pragma unroll
      for (int si = 0; si < 2; ++si) {
        const int s = split ? (h == 0 ? si : 2) : h + 2 * si;
        if (s >= nslabs) break;
        const int col = n0 + s * 64 + lane * 2;
        // columns of this n block that belong to a later n block (block_n not a multiple of 64) are skipped
        if (s * 64 + lane * 2 < p.block_n) {
          if (col < p.N) { p.stats[((size_t)slot * 2 + 0) * p.N + col] = st_sum[si][0]; p.stats[((size_t)slot * 2 + 1) * p.N + col] = st_sq[si][0]; }
          if (col + 1 < p.N) { p.stats[((size_t)slot * 2 + 0) * p.N + col + 1] = st_sum[si][1]; p.stats[((size_t)slot * 2 + 1) * p.N + col + 1] = st_sq[si][1]; }
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) { tc_fence_after(); tmem_dealloc(tmem_base, 512); }
}

// ------------------------------------------------------------------------------------------------
// host: K-major GEMM
// ------------------------------------------------------------------------------------------------
// One n block of round16(N) columns when N <= 256; otherwise full 64-column slabs only (a partial slab of one n block
// would be stored over the next block's columns), picking the width with the least padding.
static int pick_block_n(int N, int* n_blocks) {
  if (N <= 256) { *n_blocks = 1; return (N + 15) / 16 * 16; }
  int best = 256, best_pad = ceil_div(N, 256) * 256;
  for (int bn = 192; bn >= 128; bn -= 64) {
    int pad = ceil_div(N, bn) * bn;
    if (pad < best_pad) { best = bn; best_pad = pad; }
  }
  *n_blocks = ceil_div(N, best);
  return best;
}

extern "C" int mclip_gemm_tn_stat_slots(int M, int N, int batches) {
  int nb; pick_block_n(N, &nb);
  int m_tiles = batches * ceil_div(M, GEMM_BM);
  int cap = mclip_num_sms() / nb * nb;
  if (cap < nb) cap = nb;
  int grid = m_tiles * nb < cap ? m_tiles * nb : cap;
  return grid / nb * 4;
}

extern "C" int mclip_gemm_tn(const mclip_gemm_args* g, void* stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  MCLIP_REQUIRE(g && g->a && g->b && g->d, "mclip_gemm_tn: null operand");
  MCLIP_REQUIRE(g->m > 0 && g->n > 0 && g->k > 0 && g->batches > 0, "mclip_gemm_tn: empty problem %d x %d x %d x %d", g->batches, g->m, g->n, g->k);
  MCLIP_REQUIRE(g->n % 8 == 0 && g->k % 8 == 0, "mclip_gemm_tn: N=%d and K=%d must be multiples of 8", g->n, g->k);
  MCLIP_REQUIRE(g->lda >= g->k && g->ldd >= g->n, "mclip_gemm_tn: leading dimensions too small");
  GemmDev p;
  memset(&p, 0, sizeof(p));
  p.M = g->m; p.N = g->n; p.K = g->k; p.batches = g->batches; p.b_batched = g->b_batch_stride != 0;
  p.block_n = pick_block_n(p.N, &p.n_blocks);
  p.m_blocks = ceil_div(p.M, GEMM_BM);
  p.K1 = p.K; p.k1_blocks = ceil_div(p.K, GEMM_BK);
  p.K2 = g->a2 ? g->k2 : 0;
  if (p.K2 > 0) {     // A = [a | a2] along K; b = [n, k1_blocks*64 + k2] with zeros in columns [k, k1_blocks*64)
    MCLIP_REQUIRE(g->k2 % 8 == 0 && g->lda2 >= g->k2 && ((uintptr_t)g->a2 & 15) == 0 && g->lda2 % 8 == 0, "mclip_gemm_tn: bad second A operand (k2=%d)", g->k2);
    MCLIP_REQUIRE(g->ldb >= p.k1_blocks * GEMM_BK + p.K2, "mclip_gemm_tn: ldb too small for the K-concatenated weights");
  } else MCLIP_REQUIRE(g->ldb >= g->k, "mclip_gemm_tn: leading dimensions too small");
  p.k_blocks = p.k1_blocks + ceil_div(p.K2, GEMM_BK);
  p.m_tiles_total = p.batches * p.m_blocks;
  p.idesc = umma_idesc_bf16(GEMM_BM, p.block_n, 0, 0);
  p.bias = g->bias; p.residual = (const bf16*)g->residual; p.res_ld = g->ldr; p.res_bs = g->r_batch_stride; p.act = g->act;
  p.stats = g->stats;
  { const char* e = getenv("MCLIP_GEMM_DEBUG"); p.debug = e ? atoi(e) : 0; if (p.debug & 1) p.stats = nullptr; }
  p.dropmask = (const uint8_t*)g->dropmask; p.drop_scale = g->drop_scale;
  if (g->dropmask) MCLIP_REQUIRE(g->n % 16 == 0, "mclip_gemm_tn: dropout mask needs N %% 16 == 0");
  p.aux = (bf16*)g->aux_pre; p.aux_ld = g->ld_aux;
  if (g->aux_pre) MCLIP_REQUIRE(g->ld_aux >= g->n && g->ld_aux % 8 == 0 && ((uintptr_t)g->aux_pre & 15) == 0, "mclip_gemm_tn: aux_pre needs a 16-byte aligned row layout");
  p.small_k = (p.k_blocks == 1 && p.K2 == 0 && !p.b_batched && g->lda % 8 == 0 && g->a_batch_stride % 8 == 0 && ((uintptr_t)g->a & 15) == 0) ? 1 : 0;
  if (getenv("MCLIP_GEMM_NO_SMALLK")) p.small_k = 0;
  p.a_ptr = (const bf16*)g->a; p.lda = g->lda; p.a_bs = g->a_batch_stride;
  const int stage_bytes = p.small_k ? GEMM_BM * GEMM_BK * 2 : GEMM_BM * GEMM_BK * 2 + p.block_n * GEMM_BK * 2;
  const int fixed = GEMM_STAGING_BYTES + 256 + 1024 + (p.small_k ? p.block_n * GEMM_BK * 2 : 0);
  p.stages = (GEMM_SMEM_LIMIT - fixed) / stage_bytes;
  if (p.stages > 8) p.stages = 8;
  if (p.small_k) MCLIP_REQUIRE(p.stages >= 6, "mclip_gemm_tn: small-K mode needs 6 stages");   // always true for block_n <= 256
  MCLIP_REQUIRE(p.stages >= 2, "mclip_gemm_tn: tile does not fit in shared memory");
  const int smem = p.stages * stage_bytes + fixed;
  CUtensorMap tmA, tmB, tmD, tmA2;
  int rc;
  if ((rc = make_tmap_bf16_3d(&tmA, g->a, p.K1, p.M, p.batches, g->lda, g->a_batch_stride, GEMM_BK, GEMM_BM, 1))) return rc;
  if (p.K2 > 0) {
    if ((rc = make_tmap_bf16_3d(&tmA2, g->a2, p.K2, p.M, p.batches, g->lda2, g->a2_batch_stride, GEMM_BK, GEMM_BM, 1))) return rc;
  } else tmA2 = tmA;
  // B holds the K-concatenated weights: columns [0, k1_blocks*64) pair with A (zero beyond K1), the next K2 columns with A2
  const int kb_total = p.K2 > 0 ? p.k1_blocks * GEMM_BK + p.K2 : p.K;
  if ((rc = make_tmap_bf16_3d(&tmB, g->b, kb_total, p.N, p.b_batched ? p.batches : 1, g->ldb, g->b_batch_stride, GEMM_BK, p.block_n, 1))) return rc;
  if ((rc = make_tmap_bf16_3d(&tmD, g->d, p.N, p.M, p.batches, g->ldd, g->d_batch_stride, 64, 32, 1))) return rc;
  int cap = mclip_num_sms() / p.n_blocks * p.n_blocks;
  if (cap < p.n_blocks) cap = p.n_blocks;
  long long tiles = (long long)p.m_tiles_total * p.n_blocks;
  int grid = tiles < cap ? (int)tiles : cap;
  if (g->stats) MCLIP_REQUIRE(g->stat_slots == grid / p.n_blocks * 4, "mclip_gemm_tn: stat_slots=%d, expected %d", g->stat_slots, grid / p.n_blocks * 4);
  typedef void (*kern_t)(const CUtensorMap, const CUtensorMap, const CUtensorMap, const CUtensorMap, const GemmDev);
  static const kern_t kerns[4][2] = {{mclip_gemm_tn_kernel<0, false>, mclip_gemm_tn_kernel<0, true>},
                                     {mclip_gemm_tn_kernel<1, false>, mclip_gemm_tn_kernel<1, true>},
                                     {mclip_gemm_tn_kernel<2, false>, mclip_gemm_tn_kernel<2, true>},
                                     {mclip_gemm_tn_kernel<3, false>, mclip_gemm_tn_kernel<3, true>}};
  static int attr_set = 0;
  if (!attr_set) {
    for (int m = 0; m < 4; ++m)
      for (int t = 0; t < 2; ++t) MCLIP_CHECK_CUDA(cudaFuncSetAttribute(kerns[m][t], cudaFuncAttributeMaxDynamicSharedMemorySize, GEMM_SMEM_LIMIT));
    attr_set = 1;
  }
  int mode = (p.dropmask || p.aux || p.act != 0) ? 2 : p.bias ? 3 : (p.residual ? 1 : 0);
  if (getenv("MCLIP_GEMM_GENERIC")) mode = 2;                 // experiments: force the all-in-one epilogue
  // the generic epilogue is instantiated once (STATS checked at run time there: its no-statistics build spills)
  kerns[mode][(p.stats || mode == 2) ? 1 : 0]<<<grid, GEMM_THREADS, smem, stream>>>(tmA, tmB, tmD, tmA2, p);
  MCLIP_CHECK_LAUNCH();
  return MCLIP_OK;
}

// ------------------------------------------------------------------------------------------------
// weight-gradient GEMM: D[i,j] = sum_r A[r,i] * B[r,j]   (A: [R, I] , B: [R, J], both row-major, bf16)
// ------------------------------------------------------------------------------------------------
#define WG_THREADS 192       // warp 0 TMA, warp 1 MMA, warps 2-5 epilogue
#define WG_BK 64             // rows of the reduction per stage

struct WgradDev {
  int R, I, J;
  int i_blocks, j_blocks, block_j, splits, rows_per_split, stages;
  uint32_t idesc;
  float* partial;     // [splits][I][J] fp32
};

__global__ void __launch_bounds__(WG_THREADS, 1)
mclip_gemm_wgrad_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, const WgradDev p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  const int jslabs = p.block_j >> 6;
  const uint32_t a_bytes = 2 * 8192, b_bytes = (uint32_t)jslabs * 8192, stage_bytes = a_bytes + b_bytes;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + (size_t)p.stages * stage_bytes);
  uint64_t* full = bars;
  uint64_t* empty = bars + p.stages;
  uint64_t* tfull = bars + 2 * p.stages;
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(tfull + 1);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmA); tma_prefetch_desc(&tmB);
    for (int s = 0; s < p.stages; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], 1); }
    mbar_init(tfull, 1);
    fence_mbar_init();
  }
  if (warp == 1) { tmem_alloc(tmem_ptr, 256); tmem_relinquish(); }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr;

  // unit = (split, i block, j block)
  const int tiles = p.i_blocks * p.j_blocks;
  const int split = blockIdx.x / tiles, t = blockIdx.x % tiles;
  const int i0 = (t / p.j_blocks) * GEMM_BM, j0 = (t % p.j_blocks) * p.block_j;
  const int r_begin = split * p.rows_per_split;
  const int r_end = min(p.R, r_begin + p.rows_per_split);
  const int kbs = r_end > r_begin ? (r_end - r_begin + WG_BK - 1) / WG_BK : 0;

  if (warp == 0) {
    // A 64-column slab that lies entirely past I (I <= 64: the Cin / Cout side of every early layer) is zeroed ONCE and never
    // loaded: TMA fetches the full box width from L2 even when all of it is out of bounds (measured on R x 24 operands: 5.2 L2
    // sectors requested per useful one, the L2 -> SM fabric at 7 TB/s while DRAM idled at 19 %).
    const bool a_hi = i0 + 64 < p.I;
    if (!a_hi) {
      for (int st = 0; st < p.stages; ++st)
        for (int idx = lane; idx < 8192 / 16; idx += 32)
          *reinterpret_cast<uint4*>(smem + (size_t)st * stage_bytes + 8192 + idx * 16) = make_uint4(0u, 0u, 0u, 0u);
      fence_proxy_async_smem();
      __syncwarp();
    }
    if (lane == 0) {
      int stage = 0; uint32_t phase = 0;
      for (int kb = 0; kb < kbs; ++kb) {
        mbar_wait(&empty[stage], phase ^ 1);
        uint8_t* sa = smem + (size_t)stage * stage_bytes;
        mbar_expect_tx(&full[stage], a_hi ? stage_bytes : stage_bytes - 8192);
        // rows_per_split is a multiple of WG_BK, so a block never straddles two splits; rows >= R are zero-filled by TMA
        const int r0 = r_begin + kb * WG_BK;
        tma_load_3d(sa, &tmA, &full[stage], i0, r0, 0);
        if (a_hi) tma_load_3d(sa + 8192, &tmA, &full[stage], i0 + 64, r0, 0);
        for (int js = 0; js < jslabs; ++js) tma_load_3d(sa + a_bytes + js * 8192, &tmB, &full[stage], j0 + js * 64, r0, 0);
        if (++stage == p.stages) { stage = 0; phase ^= 1; }
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      int stage = 0; uint32_t phase = 0;
      for (int kb = 0; kb < kbs; ++kb) {
        mbar_wait(&full[stage], phase);
        tc_fence_after();
        const uint32_t sa = smem_u32(smem + (size_t)stage * stage_bytes), sb = sa + a_bytes;
#pragma unroll
        for (int kk = 0; kk < 4; ++kk)     // 16 reduction rows per MMA = 2048 B further into every slab
          umma_bf16(tmem_base, umma_smem_desc_sw128(sa + kk * 2048, 8192, 1024), umma_smem_desc_sw128(sb + kk * 2048, 8192, 1024),
                    p.idesc, (kb | kk) != 0);
        umma_commit(&empty[stage]);
        if (++stage == p.stages) { stage = 0; phase ^= 1; }
      }
      umma_commit(tfull);
    }
  } else {
    const int q = warp & 3;
    const int i = i0 + q * 32 + lane;
    float* out = p.partial + ((size_t)split * p.I + i) * p.J;
    if (kbs > 0) {
      mbar_wait(tfull, 0);
      tc_fence_after();
    }
    for (int c = 0; c < p.block_j; c += 16) {
      float v[16];
      if (kbs > 0) {
        uint32_t r[16];
        tmem_ld16(tmem_base + (uint32_t)c + ((uint32_t)(q * 32) << 16), r);
        tmem_ld_wait();
#pragma unroll
        for (int e = 0; e < 16; ++e) v[e] = __uint_as_float(r[e]);
      } else {
#pragma unroll
        for (int e = 0; e < 16; ++e) v[e] = 0.f;
      }
      if (i < p.I) {
#pragma unroll
        for (int e = 0; e < 16; e += 4) {
          const int j = j0 + c + e;
          if (j + 3 < p.J) *reinterpret_cast<float4*>(out + j) = make_float4(v[e], v[e + 1], v[e + 2], v[e + 3]);
          else
            for (int x = 0; x < 4; ++x) if (j + x < p.J) out[j + x] = v[e + x];
        }
      }
    }
    tc_fence_before();
  }
  __syncthreads();
  if (warp == 1) { tc_fence_after(); tmem_dealloc(tmem_base, 256); }
}

// out[i,j] (+)= sum_s partial[s,i,j]   (fixed order => deterministic)
__global__ void mclip_wgrad_reduce_kernel(const float* __restrict__ partial, float* __restrict__ out, int splits, long long n,
                                          long long ld_out, int J, int accumulate, float scale) {
  for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < n; e += (long long)gridDim.x * blockDim.x) {
    float s = 0.f;
    for (int k = 0; k < splits; ++k) s += partial[(size_t)k * n + e];
    long long i = e / J, j = e % J;
    float* o = out + i * ld_out + j;
    *o = accumulate ? (*o + s * scale) : s * scale;
  }
}

static void wgrad_plan(int R, int I, int J, int* block_j, int* i_blocks, int* j_blocks, int* splits, int* rows_per_split) {
  int jb = ceil_div(J, 256);
  int bj = ceil_div(ceil_div(J, jb), 64) * 64;
  *block_j = bj;
  *j_blocks = ceil_div(J, bj);
  *i_blocks = ceil_div(I, GEMM_BM);
  int tiles = *i_blocks * *j_blocks;
  int sp = mclip_num_sms() / tiles;
  if (sp < 1) sp = 1;
  int kbs = ceil_div(R, WG_BK);
  if (sp > kbs) sp = kbs;
  int rps = ceil_div(kbs, sp) * WG_BK;
  *splits = ceil_div(R, rps);
  *rows_per_split = rps;
}

extern "C" long long mclip_gemm_wgrad_workspace_bytes(int R, int I, int J) {
  int bj, ib, jb, sp, rps;
  wgrad_plan(R, I, J, &bj, &ib, &jb, &sp, &rps);
  return (long long)sp * I * J * 4 + 256;
}

extern "C" int mclip_gemm_wgrad(const mclip_wgrad_args* g, void* stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  MCLIP_REQUIRE(g && g->a && g->b && g->out && g->workspace, "mclip_gemm_wgrad: null operand");
  MCLIP_REQUIRE(g->r > 0 && g->i > 0 && g->j > 0, "mclip_gemm_wgrad: empty problem");
  MCLIP_REQUIRE(g->i % 8 == 0 && g->j % 8 == 0 && g->lda % 8 == 0 && g->ldb % 8 == 0, "mclip_gemm_wgrad: I, J and leading dims must be multiples of 8");
  WgradDev p;
  memset(&p, 0, sizeof(p));
  p.R = g->r; p.I = g->i; p.J = g->j;
  wgrad_plan(p.R, p.I, p.J, &p.block_j, &p.i_blocks, &p.j_blocks, &p.splits, &p.rows_per_split);
  MCLIP_REQUIRE(g->workspace_bytes >= (long long)p.splits * p.I * p.J * 4, "mclip_gemm_wgrad: workspace too small");
  p.idesc = umma_idesc_bf16(GEMM_BM, p.block_j, 1, 1);
  p.partial = (float*)g->workspace;
  const int stage_bytes = 2 * 8192 + (p.block_j / 64) * 8192;
  p.stages = (GEMM_SMEM_LIMIT - 2048) / stage_bytes;
  if (p.stages > 8) p.stages = 8;
  const int smem = p.stages * stage_bytes + 2048;
  CUtensorMap tmA, tmB;
  int rc;
  if ((rc = make_tmap_bf16_3d(&tmA, g->a, p.I, p.R, 1, g->lda, 0, 64, WG_BK, 1))) return rc;
  if ((rc = make_tmap_bf16_3d(&tmB, g->b, p.J, p.R, 1, g->ldb, 0, 64, WG_BK, 1))) return rc;
  static int attr_set = 0;
  if (!attr_set) {
    MCLIP_CHECK_CUDA(cudaFuncSetAttribute(mclip_gemm_wgrad_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, GEMM_SMEM_LIMIT));
    attr_set = 1;
  }
  mclip_gemm_wgrad_kernel<<<p.splits * p.i_blocks * p.j_blocks, WG_THREADS, smem, stream>>>(tmA, tmB, p);
  MCLIP_CHECK_LAUNCH();
  const long long n = (long long)p.I * p.J;
  int rgrid = (int)((n + 255) / 256);
  if (rgrid > 4 * mclip_num_sms()) rgrid = 4 * mclip_num_sms();
  mclip_wgrad_reduce_kernel<<<rgrid, 256, 0, stream>>>(p.partial, g->out, p.splits, n, g->ldo, p.J, g->accumulate, 1.0f);
  MCLIP_CHECK_LAUNCH();
  return MCLIP_OK;
}
