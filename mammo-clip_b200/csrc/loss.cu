// Fused in-batch InfoNCE: (optional) NVLink push all-gather of the embeddings + logits GEMMs + row/column
// log-sum-exp + label-smoothed cross-entropy + the complete backward (dE for the local rows of every embedding
// tensor and d(logit_scale)), as ONE cooperative kernel.
//
// Replaces, in the reference:  loss/breast_clip_contrastive.py:28-59, loss/breast_clip.py:29-127 and the
// all_gather / reduce_scatter pair of util/dist_autograd.py:4-26.
//
// Formulation.  A "pair" (a, b) is the score matrix S = scale * E_a @ E_b^T over ALL W*B samples.  The reference's
// `logit_scale * x_a @ all_b.T` cross-entropy is the row-wise CE of S restricted to this rank's rows, and
// `logit_scale * x_b @ all_a.T` is the column-wise CE of S restricted to this rank's columns.  With q the smoothed
// one-hot target, the gradient that the reference obtains through reduce_scatter(SUM) of every rank's loss is
//     G_ij = [ w_row (softmax_row(S)_ij - q_ij) + w_col (softmax_col(S)_ij - q_ij) ] / B
//     dE_a[i] = scale * sum_j G_ij E_b[j]      dE_b[j] = scale * sum_i G_ij E_a[i]
// for local i (resp. j) and all j (resp. i): once every rank holds all embeddings it can form these without any
// further communication (verified against the reference on 2 gloo ranks).  d(scale) stays the
// rank-local dL_r/dscale, as in the reference (DDP averages it).
//
// Phases (grid barriers in between; every phase is a grid-stride loop over 32x64 score tiles, fp32 SIMT):
//   0  push: every CTA copies a slice of the local [K,B,D] slab into all W ranks' gather buffers (peer pointers),
//      then releases a per-source arrival counter on each peer (red.release.sys).  Tiles acquire only the sources
//      they read, so local tiles start while remote slabs are still in flight.
//   1  all tiles of S for every pair -> per-tile (max, sum exp) partials for rows and columns.
//   2  local rows x all columns and local columns x all rows: recompute the tile, build G, emit partial dE
//      (G @ E) and the scalar partials (loss terms, d scale).
//   3  deterministic fixed-order reduction of the partials into dE_k, loss, per-pair components, d scale.
//
// LSE exchange (world > 1, large global batches; ABI 5).  Phase 2 needs the row LSE of the local rows, the column LSE of the local
// columns AND the LSEs of every other rank's rows / columns.  Without an exchange every rank evaluates the whole (W*B)^2 matrix in
// phase 1 -- W/2 times more fp32 work than the "cross" (local rows x all columns + all rows x local columns) that yields its own
// LSEs.  With lse_all set, phase 1 covers only the cross, phase 1b pushes the 2*P*B local LSEs to every peer (second arrival on
// the same counters) and phase 2 reads the others' from the exchanged table.
#include "common.cuh"
#include "mclip_internal.h"
#include <stdlib.h>
#include <cooperative_groups.h>
#include <math.h>
namespace cg = cooperative_groups;

#define LT_M 32
#define LT_N 64
#define LT_K 16
#define LOSS_THREADS 256

struct LossDev {
  int W, rank, B, D, K, P;
  int WB, nI32, nJ64;          // tiles of the full score matrix: rows in 32s, columns in 64s
  int nLB32;                   // 32-row blocks of the local B rows
  float scale;
  const float* scale_dev;          // device scalar (takes precedence over `scale`)
  const float* local[MCLIP_LOSS_MAX_TENSORS];      // this rank's rows [B,D]
  const float* all[MCLIP_LOSS_MAX_TENSORS];        // gathered [W*B,D] (== local when W==1)
  float* const* peer_all;      // device table [W][K] of peers' gather buffers (W>1)
  unsigned int* const* peer_flags;  // device table [W] of peers' arrival counters (each [W] u32)
  const unsigned int* my_flags;     // this rank's arrival counters [W]
  unsigned int flag_target;         // arrivals (per source rank) that complete the embedding push of this call
  unsigned int flag_target2;        // ... and the LSE push (exchange mode)
  float* lse_all;                   // [W*B][2*MCLIP_LOSS_MAX_PAIRS]: (row LSE, col LSE) per pair of every global row/column, or nullptr
  float* const* peer_lse;           // device table [W] of the peers' lse_all
  int pa[MCLIP_LOSS_MAX_PAIRS], pb[MCLIP_LOSS_MAX_PAIRS];
  float w_row[MCLIP_LOSS_MAX_PAIRS], w_col[MCLIP_LOSS_MAX_PAIRS], eps[MCLIP_LOSS_MAX_PAIRS];
  // workspace
  float2* rowpart;   // [P][WB][nJ64]
  float2* colpart;   // [P][WB][nI32]
  float* gpart;      // [P][2][nJ64][B][D]
  float* spart;      // [P][2][nLB32][nJ64][2]   (loss, dscale) partials
  // outputs
  float* dE[MCLIP_LOSS_MAX_TENSORS];   // [B,D] each
  float* out;        // [2 + 2P]: loss, dscale, then per pair (row CE mean, col CE mean)
  int* status;       // device int or nullptr: 1 + rank of a peer that never arrived
  long long timeout_cycles;
  int window;        // record the transfer window (mclip_loss_win)
};

// Transfer-window instrumentation (bench.py --workload loss-sweep; SURVEY 8d "first remote store to last flag observed"):
// [0] = earliest %globaltimer at which a CTA of this rank started pushing; [1 + s] = EARLIEST %globaltimer at which any tile of this
// rank saw the arrival counter of remote source s complete (the first tile that needs a source polls from the start of phase 1, so
// this is the arrival as seen here; later observers only find it already set).  Window = max_s [1 + s] - [0].  Off unless
// mclip_loss_window(.., enable) switched it on.
__device__ unsigned long long mclip_loss_win[1 + MCLIP_LOSS_MAX_WORLD];
static int g_loss_window = 0;
__device__ __forceinline__ unsigned long long globaltimer_ns() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
  return t;
}
extern "C" int mclip_loss_window(unsigned long long* out2, int enable) {
  g_loss_window = enable;
  unsigned long long w[1 + MCLIP_LOSS_MAX_WORLD];
  if (out2) {
    MCLIP_CHECK_CUDA(cudaDeviceSynchronize());
    MCLIP_CHECK_CUDA(cudaMemcpyFromSymbol(w, mclip_loss_win, sizeof(w)));
    out2[0] = w[0];
    out2[1] = 0;
    for (int s2 = 0; s2 < MCLIP_LOSS_MAX_WORLD; ++s2)
      if (w[1 + s2] != ~0ull && w[1 + s2] > out2[1]) out2[1] = w[1 + s2];
  }
  for (int s2 = 0; s2 < 1 + MCLIP_LOSS_MAX_WORLD; ++s2) w[s2] = ~0ull;
  MCLIP_CHECK_CUDA(cudaMemcpyToSymbol(mclip_loss_win, w, sizeof(w)));
  return MCLIP_OK;
}

__device__ __forceinline__ unsigned int ld_acquire_sys(const unsigned int* p) {
  unsigned int v;
  asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void red_release_sys_add(unsigned int* p, unsigned int v) {
  asm volatile("red.release.sys.global.add.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}

// Wait until the slabs of every source rank that owns rows [r0, r1) have landed in this rank's gather buffer.
__device__ __forceinline__ void wait_sources(const LossDev& p, int r0, int r1, unsigned int target) {
  if (p.W == 1) return;
  if (threadIdx.x == 0) {
    int s0 = r0 / p.B, s1 = (min(r1, p.WB) - 1) / p.B;
    for (int s = s0; s <= s1; ++s) {
      if (s == p.rank) continue;   // own slab is written by this very grid before the phase barrier
      long long t0 = clock64();
      while ((int)(ld_acquire_sys(p.my_flags + s) - target) < 0) {
        if (clock64() - t0 > p.timeout_cycles) {   // a peer never arrived: report and finish (no trap: the context survives)
          if (p.status && atomicCAS(p.status, 0, 1 + s) == 0)
            printf("mclip loss: rank %d timed out waiting for the embeddings of rank %d\n", p.rank, s);
          break;
        }
        __nanosleep(200);
      }
      if (p.window && target == p.flag_target) atomicMin(&mclip_loss_win[1 + s], globaltimer_ns());
    }
  }
  __syncthreads();
}

// S tile = scale * X[x0:x0+32] @ Y[y0:y0+64]^T ; thread (ty,tx) holds rows ty*2+{0,1}, cols tx*4+{0..3}
__device__ __forceinline__ void score_tile(const float* __restrict__ X, int x0, int nx, const float* __restrict__ Y, int y0,
                                           int ny, int D, float scale, float (*Xs)[LT_M + 4], float (*Ys)[LT_N + 4],
                                           float acc[2][4]) {
  const int tid = threadIdx.x;
  const int ty = tid >> 4, tx = tid & 15;
#pragma unroll
  for (int i = 0; i < 2; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
  // loader mapping: float4 along D.  X: threads 0..127 -> row = t>>2, quad = t&3 ; Y: all 256 -> row = t>>2
  const int lrow = tid >> 2, lq = tid & 3;
  float4 xr = make_float4(0, 0, 0, 0), yr = make_float4(0, 0, 0, 0);
  auto fetch = [&](int k0) {
    if (lrow < LT_M) {
      int r = x0 + lrow;
      xr = (r < nx) ? *reinterpret_cast<const float4*>(X + (size_t)r * D + k0 + lq * 4) : make_float4(0, 0, 0, 0);
    }
    int r = y0 + lrow;
    yr = (r < ny) ? *reinterpret_cast<const float4*>(Y + (size_t)r * D + k0 + lq * 4) : make_float4(0, 0, 0, 0);
  };
  fetch(0);
  for (int k0 = 0; k0 < D; k0 += LT_K) {
    __syncthreads();
    if (lrow < LT_M) {
      Xs[lq * 4 + 0][lrow] = xr.x; Xs[lq * 4 + 1][lrow] = xr.y; Xs[lq * 4 + 2][lrow] = xr.z; Xs[lq * 4 + 3][lrow] = xr.w;
    }
    Ys[lq * 4 + 0][lrow] = yr.x; Ys[lq * 4 + 1][lrow] = yr.y; Ys[lq * 4 + 2][lrow] = yr.z; Ys[lq * 4 + 3][lrow] = yr.w;
    __syncthreads();
    if (k0 + LT_K < D) fetch(k0 + LT_K);
#pragma unroll
    for (int kk = 0; kk < LT_K; ++kk) {
      float2 a = *reinterpret_cast<const float2*>(&Xs[kk][ty * 2]);
      float4 b = *reinterpret_cast<const float4*>(&Ys[kk][tx * 4]);
      acc[0][0] = fmaf(a.x, b.x, acc[0][0]); acc[0][1] = fmaf(a.x, b.y, acc[0][1]);
      acc[0][2] = fmaf(a.x, b.z, acc[0][2]); acc[0][3] = fmaf(a.x, b.w, acc[0][3]);
      acc[1][0] = fmaf(a.y, b.x, acc[1][0]); acc[1][1] = fmaf(a.y, b.y, acc[1][1]);
      acc[1][2] = fmaf(a.y, b.z, acc[1][2]); acc[1][3] = fmaf(a.y, b.w, acc[1][3]);
    }
  }
#pragma unroll
  for (int i = 0; i < 2; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] *= scale;
}

__device__ __forceinline__ float lse_combine(const float2* parts, int n) {
  float m = -INFINITY;
  for (int t = 0; t < n; ++t) m = fmaxf(m, parts[t].x);
  float l = 0.f;
  for (int t = 0; t < n; ++t) l += parts[t].y * __expf(parts[t].x - m);
  return m + logf(l);
}

__global__ void __launch_bounds__(LOSS_THREADS) mclip_loss_kernel(const LossDev p_in) {
  LossDev p = p_in;
  if (p.scale_dev) p.scale = *p.scale_dev;
  cg::grid_group grid = cg::this_grid();
  __shared__ __align__(16) float Xs[LT_K][LT_M + 4];
  __shared__ __align__(16) float Ys[LT_K][LT_N + 4];
  __shared__ __align__(16) float Gt[LT_N][LT_M + 4];   // S tile, later G transposed: [col][row]
  __shared__ float lseX[LT_M], lseY[LT_N];
  __shared__ float red[LOSS_THREADS / 32][2];
  const int tid = threadIdx.x, ty = tid >> 4, tx = tid & 15;
  const int WB = p.WB, D = p.D, B = p.B;

  // ---------------- phase 0: push the local slab to every rank's gather buffer ----------------
  if (p.W > 1) {
    if (p.window && tid == 0) atomicMin(&mclip_loss_win[0], globaltimer_ns());
    const int vecs = B * D / 4;   // float4 per tensor
    for (int dst = 0; dst < p.W; ++dst) {
      int peer = (p.rank + dst) % p.W;   // stagger destinations across ranks
      for (int k = 0; k < p.K; ++k) {
        float4* out = reinterpret_cast<float4*>(p.peer_all[peer * p.K + k] + (size_t)p.rank * B * D);
        const float4* in = reinterpret_cast<const float4*>(p.local[k]);
        for (int v = blockIdx.x * LOSS_THREADS + tid; v < vecs; v += gridDim.x * LOSS_THREADS) out[v] = in[v];
      }
    }
    __threadfence_system();
    __syncthreads();
    if (tid == 0)
      for (int peer = 0; peer < p.W; ++peer)
        if (peer != p.rank) red_release_sys_add(p.peer_flags[peer] + p.rank, 1u);
    grid.sync();   // own slab is complete in the local gather buffer
  }

  // ---------------- phase 1: LSE partials over the full score matrix (or, exchange mode, over this rank's cross) ----------------
  {
    const bool xch = p.lse_all != nullptr;
    // tile ranges that contain the local rows / columns
    const int it0 = (p.rank * B) / LT_M, it1 = ((p.rank + 1) * B - 1) / LT_M, nItL = it1 - it0 + 1;
    const int jt0 = (p.rank * B) / LT_N, jt1 = ((p.rank + 1) * B - 1) / LT_N, nJtL = jt1 - jt0 + 1;
    const int unitsA = nItL * p.nJ64, unitsB = (p.nI32 - nItL) * nJtL;
    const int per_pair = xch ? unitsA + unitsB : p.nI32 * p.nJ64;
    for (int u = blockIdx.x; u < p.P * per_pair; u += gridDim.x) {
      int pr = u / per_pair, r = u % per_pair, it, jt;
      if (!xch) { it = r / p.nJ64; jt = r % p.nJ64; }
      else if (r < unitsA) { it = it0 + r / p.nJ64; jt = r % p.nJ64; }                    // local rows x all columns
      else { r -= unitsA; it = r / nJtL; if (it >= it0) it += nItL; jt = jt0 + r % nJtL; }   // other rows x local columns
      int x0 = it * LT_M, y0 = jt * LT_N;
      wait_sources(p, x0, x0 + LT_M, p.flag_target);
      wait_sources(p, y0, y0 + LT_N, p.flag_target);
      float acc[2][4];
      score_tile(p.all[p.pa[pr]], x0, WB, p.all[p.pb[pr]], y0, WB, D, p.scale, Xs, Ys, acc);
#pragma unroll
      for (int i = 0; i < 2; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) Gt[tx * 4 + j][ty * 2 + i] = acc[i][j];
      __syncthreads();
      if (tid < LT_M) {                 // row partial over this tile's valid columns
        int gi = x0 + tid;
        if (gi < WB) {
          int nc = min(LT_N, WB - y0);
          float m = -INFINITY;
          for (int c = 0; c < nc; ++c) m = fmaxf(m, Gt[c][tid]);
          float l = 0.f;
          for (int c = 0; c < nc; ++c) l += __expf(Gt[c][tid] - m);
          p.rowpart[((size_t)pr * WB + gi) * p.nJ64 + jt] = make_float2(m, l);
        }
      } else if (tid >= 64 && tid < 64 + LT_N) {   // column partial over this tile's valid rows
        int c = tid - 64, gj = y0 + c;
        if (gj < WB) {
          int nr = min(LT_M, WB - x0);
          float m = -INFINITY;
          for (int r2 = 0; r2 < nr; ++r2) m = fmaxf(m, Gt[c][r2]);
          float l = 0.f;
          for (int r2 = 0; r2 < nr; ++r2) l += __expf(Gt[c][r2] - m);
          p.colpart[((size_t)pr * WB + gj) * p.nI32 + it] = make_float2(m, l);
        }
      }
      __syncthreads();
    }
  }
  if (p.lse_all) {
    // ---------------- phase 1b: combine the local LSEs and push them to every rank ----------------
    __threadfence();
    grid.sync();
    for (int e = blockIdx.x * LOSS_THREADS + tid; e < p.P * B; e += gridDim.x * LOSS_THREADS) {
      const int pr = e / B, gi = p.rank * B + e % B;
      const float lr = lse_combine(p.rowpart + ((size_t)pr * WB + gi) * p.nJ64, p.nJ64);
      const float lc = lse_combine(p.colpart + ((size_t)pr * WB + gi) * p.nI32, p.nI32);
      for (int dst = 0; dst < p.W; ++dst) {
        float* o = p.peer_lse[(p.rank + dst) % p.W] + (size_t)gi * (2 * MCLIP_LOSS_MAX_PAIRS) + 2 * pr;
        *reinterpret_cast<float2*>(o) = make_float2(lr, lc);
      }
    }
    __threadfence_system();
    __syncthreads();
    if (tid == 0)
      for (int peer = 0; peer < p.W; ++peer)
        if (peer != p.rank) red_release_sys_add(p.peer_flags[peer] + p.rank, 1u);
  }
  __threadfence();
  grid.sync();

  // ---------------- phase 2: gradient tiles for the local rows (side 0) and local columns (side 1) ----------------
  {
    const int per_side = p.nLB32 * p.nJ64;
    for (int u = blockIdx.x; u < p.P * 2 * per_side; u += gridDim.x) {
      int pr = u / (2 * per_side), r = u % (2 * per_side), side = r / per_side;
      r %= per_side;
      int lb = r / p.nJ64, jt = r % p.nJ64;
      // side 0: X = E_a (local rows), Y = E_b (all);  side 1: X = E_b (local columns of S), Y = E_a (all rows of S)
      const int tx_id = side == 0 ? p.pa[pr] : p.pb[pr], ty_id = side == 0 ? p.pb[pr] : p.pa[pr];
      const float wX = side == 0 ? p.w_row[pr] : p.w_col[pr], wY = side == 0 ? p.w_col[pr] : p.w_row[pr];
      const float2* partX = side == 0 ? p.rowpart : p.colpart;   // LSE partials along the X-indexed dimension
      const float2* partY = side == 0 ? p.colpart : p.rowpart;
      const int nPX = side == 0 ? p.nJ64 : p.nI32, nPY = side == 0 ? p.nI32 : p.nJ64;
      const float eps = p.eps[pr];
      const int lx0 = lb * LT_M;                   // first local row of the tile
      const int gx0 = p.rank * B + lx0;            // its global index
      const int y0 = jt * LT_N;
      const int nxv = min(LT_M, B - lx0), nyv = min(LT_N, WB - y0);
      wait_sources(p, y0, y0 + LT_N, p.lse_all ? p.flag_target2 : p.flag_target);
      if (p.lse_all) {                 // exchanged table: [global index][pair][row LSE, col LSE]; X takes side's own kind, Y the other
        if (tid < LT_M) lseX[tid] = (tid < nxv) ? p.lse_all[(size_t)(gx0 + tid) * (2 * MCLIP_LOSS_MAX_PAIRS) + 2 * pr + side] : 0.f;
        else if (tid >= 64 && tid < 64 + LT_N) {
          int c = tid - 64;
          lseY[c] = (c < nyv) ? p.lse_all[(size_t)(y0 + c) * (2 * MCLIP_LOSS_MAX_PAIRS) + 2 * pr + (1 - side)] : 0.f;
        }
      } else if (tid < LT_M) lseX[tid] = (tid < nxv) ? lse_combine(partX + ((size_t)pr * WB + gx0 + tid) * nPX, nPX) : 0.f;
      else if (tid >= 64 && tid < 64 + LT_N) {
        int c = tid - 64;
        lseY[c] = (c < nyv) ? lse_combine(partY + ((size_t)pr * WB + y0 + c) * nPY, nPY) : 0.f;
      }
      float acc[2][4];
      // rows of X come from the gathered copy so both sides read identical bits on every rank
      score_tile(p.all[tx_id] + (size_t)p.rank * B * D, lx0, B, p.all[ty_id], y0, WB, D, p.scale, Xs, Ys, acc);
      __syncthreads();
      const float invB = 1.0f / (float)B, qoff = eps / (float)WB;
      float loss_part = 0.f, ds_part = 0.f;
#pragma unroll
      for (int i = 0; i < 2; ++i) {
        int lr = ty * 2 + i;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          int lc = tx * 4 + j;
          float g = 0.f;
          if (lr < nxv && lc < nyv) {
            float s = acc[i][j];
            float q = qoff + ((gx0 + lr) == (y0 + lc) ? (1.0f - eps) : 0.f);
            float pX = __expf(s - lseX[lr]), pY = __expf(s - lseY[lc]);
            g = invB * (wX * (pX - q) + wY * (pY - q));
            loss_part -= wX * invB * q * s;                // -(1-eps) S_xx - (eps/WB) sum_y S_xy
            ds_part += wX * invB * (pX - q) * s;
          }
          Gt[lc][lr] = g;
        }
      }
      if (jt == 0 && tid < nxv) loss_part += wX * invB * lseX[tid];
      __syncthreads();
      // partial dX[32, D] = scale * G[32,64] @ Y[64, D]
      {
        const float* Yp = p.all[ty_id];
        float* gp = p.gpart + ((((size_t)pr * 2 + side) * p.nJ64 + jt) * B + lx0) * D;
        const int half = tid >> 7, t7 = tid & 127;       // rows half*16.., 4 columns per thread per pass
        for (int d0 = t7 * 4; d0 < D; d0 += 512) {
          float o[16][4];
#pragma unroll
          for (int i = 0; i < 16; ++i) o[i][0] = o[i][1] = o[i][2] = o[i][3] = 0.f;
          for (int c = 0; c < nyv; ++c) {
            float4 y = *reinterpret_cast<const float4*>(Yp + (size_t)(y0 + c) * D + d0);
#pragma unroll
            for (int i4 = 0; i4 < 4; ++i4) {
              float4 g4 = *reinterpret_cast<const float4*>(&Gt[c][half * 16 + i4 * 4]);
              float gg[4] = {g4.x, g4.y, g4.z, g4.w};
#pragma unroll
              for (int e = 0; e < 4; ++e) {
                o[i4 * 4 + e][0] = fmaf(gg[e], y.x, o[i4 * 4 + e][0]);
                o[i4 * 4 + e][1] = fmaf(gg[e], y.y, o[i4 * 4 + e][1]);
                o[i4 * 4 + e][2] = fmaf(gg[e], y.z, o[i4 * 4 + e][2]);
                o[i4 * 4 + e][3] = fmaf(gg[e], y.w, o[i4 * 4 + e][3]);
              }
            }
          }
#pragma unroll
          for (int i = 0; i < 16; ++i) {
            int lr = half * 16 + i;
            if (lr < nxv)
              *reinterpret_cast<float4*>(gp + (size_t)lr * D + d0) =
                  make_float4(o[i][0] * p.scale, o[i][1] * p.scale, o[i][2] * p.scale, o[i][3] * p.scale);
          }
        }
      }
      // scalar partials, fixed-order block reduction
      loss_part = warp_sum(loss_part);
      ds_part = warp_sum(ds_part);
      if ((tid & 31) == 0) { red[tid >> 5][0] = loss_part; red[tid >> 5][1] = ds_part; }
      __syncthreads();
      if (tid == 0) {
        float a = 0.f, b = 0.f;
        for (int w = 0; w < LOSS_THREADS / 32; ++w) { a += red[w][0]; b += red[w][1]; }
        float* sp = p.spart + ((((size_t)pr * 2 + side) * p.nLB32 + lb) * p.nJ64 + jt) * 2;
        sp[0] = a;
        sp[1] = b / p.scale;
      }
      __syncthreads();
    }
  }
  __threadfence();
  grid.sync();

  // ---------------- phase 3: fixed-order reductions ----------------
  {
    const size_t slab = (size_t)B * D;
    for (int k = 0; k < p.K; ++k) {
      for (size_t e = (size_t)blockIdx.x * LOSS_THREADS + tid; e < slab; e += (size_t)gridDim.x * LOSS_THREADS) {
        float s = 0.f;
        for (int pr = 0; pr < p.P; ++pr)
          for (int side = 0; side < 2; ++side) {
            if ((side == 0 ? p.pa[pr] : p.pb[pr]) != k) continue;
            const float* gp = p.gpart + (((size_t)pr * 2 + side) * p.nJ64) * slab + e;
            for (int jt = 0; jt < p.nJ64; ++jt) s += gp[(size_t)jt * slab];
          }
        p.dE[k][e] = s;
      }
    }
    if (blockIdx.x == 0 && tid == 0) {
      float loss = 0.f, ds = 0.f;
      for (int pr = 0; pr < p.P; ++pr)
        for (int side = 0; side < 2; ++side) {
          float a = 0.f;
          const float* sp = p.spart + (((size_t)pr * 2 + side) * p.nLB32) * p.nJ64 * 2;
          for (int t = 0; t < p.nLB32 * p.nJ64; ++t) { a += sp[2 * t]; ds += sp[2 * t + 1]; }
          loss += a;
          float w = side == 0 ? p.w_row[pr] : p.w_col[pr];
          p.out[2 + 2 * pr + side] = (w != 0.f) ? a / w : 0.f;   // the unweighted CE mean of this direction
        }
      p.out[0] = (p.status && *reinterpret_cast<volatile int*>(p.status) != 0) ? __int_as_float(0x7fc00000) : loss;
      p.out[1] = ds;
    }
  }
}

// ------------------------------------------------------------------------------------------------
// host entry points
// ------------------------------------------------------------------------------------------------
static int loss_tiles(int W, int B, int* nI32, int* nJ64, int* nLB32) {
  int WB = W * B;
  *nI32 = ceil_div(WB, LT_M);
  *nJ64 = ceil_div(WB, LT_N);
  *nLB32 = ceil_div(B, LT_M);
  return WB;
}

extern "C" long long mclip_loss_workspace_bytes(int W, int B, int D, int P) {
  int nI, nJ, nL;
  long long WB = loss_tiles(W, B, &nI, &nJ, &nL);
  long long bytes = 0;
  bytes += (long long)P * WB * nJ * 8;            // rowpart
  bytes += (long long)P * WB * nI * 8;            // colpart
  bytes += (long long)P * 2 * nJ * B * D * 4;     // gpart
  bytes += (long long)P * 2 * nL * nJ * 2 * 4;    // spart
  return bytes + 1024;
}

extern "C" int mclip_loss_grid(int W, int B, int P) {
  int nI, nJ, nL;
  loss_tiles(W, B, &nI, &nJ, &nL);
  int occ = 0;
  if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, mclip_loss_kernel, LOSS_THREADS, 0) != cudaSuccess || occ < 1) return -1;
  int cap = mclip_num_sms() * (occ > 2 ? 2 : occ);
  int units = P * nI * nJ;
  int u2 = P * 2 * nL * nJ;
  if (u2 > units) units = u2;
  if (units < 1) units = 1;
  // all ranks must agree on the grid (the arrival counters count CTAs): it depends only on (W,B,P) and the SM count
  return units < cap ? units : cap;
}

extern "C" int mclip_contrastive_loss(const mclip_loss_args* a, void* stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  MCLIP_REQUIRE(a != nullptr, "mclip_contrastive_loss: null args");
  MCLIP_REQUIRE(a->world >= 1 && a->world <= MCLIP_LOSS_MAX_WORLD, "world=%d out of range", a->world);
  MCLIP_REQUIRE(a->rank >= 0 && a->rank < a->world, "rank=%d out of range", a->rank);
  MCLIP_REQUIRE(a->batch >= 1, "batch must be >= 1");
  MCLIP_REQUIRE(a->dim >= 16 && a->dim % 16 == 0, "embedding dim %d must be a multiple of 16", a->dim);
  MCLIP_REQUIRE(a->n_tensors >= 1 && a->n_tensors <= MCLIP_LOSS_MAX_TENSORS, "n_tensors=%d out of range", a->n_tensors);
  MCLIP_REQUIRE(a->n_pairs >= 1 && a->n_pairs <= MCLIP_LOSS_MAX_PAIRS, "n_pairs=%d out of range", a->n_pairs);
  MCLIP_REQUIRE(a->workspace && a->out, "workspace/out must be device pointers");
  MCLIP_REQUIRE(a->workspace_bytes >= mclip_loss_workspace_bytes(a->world, a->batch, a->dim, a->n_pairs),
                "workspace too small: %lld", a->workspace_bytes);
  LossDev p;
  memset(&p, 0, sizeof(p));
  p.W = a->world; p.rank = a->rank; p.B = a->batch; p.D = a->dim; p.K = a->n_tensors; p.P = a->n_pairs;
  p.WB = loss_tiles(p.W, p.B, &p.nI32, &p.nJ64, &p.nLB32);
  p.scale = a->logit_scale;
  p.scale_dev = a->logit_scale_dev;
  for (int k = 0; k < p.K; ++k) {
    MCLIP_REQUIRE(a->local[k] && a->grad[k], "tensor %d: null pointer", k);
    p.local[k] = a->local[k];
    p.all[k] = (p.W == 1) ? a->local[k] : a->gathered[k];
    MCLIP_REQUIRE(p.all[k], "tensor %d: gathered buffer missing", k);
    p.dE[k] = a->grad[k];
  }
  for (int i = 0; i < p.P; ++i) {
    MCLIP_REQUIRE(a->pair_a[i] >= 0 && a->pair_a[i] < p.K && a->pair_b[i] >= 0 && a->pair_b[i] < p.K, "pair %d: bad tensor index", i);
    p.pa[i] = a->pair_a[i]; p.pb[i] = a->pair_b[i];
    p.w_row[i] = a->w_row[i]; p.w_col[i] = a->w_col[i]; p.eps[i] = a->label_smoothing[i];
  }
  int grid = mclip_loss_grid(p.W, p.B, p.P);
  MCLIP_REQUIRE(grid >= 1, "could not size the cooperative grid");
  if (p.W > 1) {
    MCLIP_REQUIRE(a->peer_gathered && a->peer_flags && a->my_flags, "world>1 needs the peer pointer tables");
    MCLIP_REQUIRE(a->epoch >= 1, "epoch must start at 1 and increase by 1 per call");
    MCLIP_REQUIRE((a->batch * a->dim) % 4 == 0, "slab must be float4 divisible");
    p.peer_all = (float* const*)a->peer_gathered;
    p.peer_flags = (unsigned int* const*)a->peer_flags;
    p.my_flags = (const unsigned int*)a->my_flags;
    p.lse_all = a->lse_all; p.peer_lse = (float* const*)a->peer_lse;
    if (p.lse_all) MCLIP_REQUIRE(p.peer_lse, "lse_all needs the peer_lse table");
    const unsigned int per_call = p.lse_all ? 2u : 1u;       // arrivals per CTA and call on each peer's counter
    p.flag_target = ((unsigned int)a->epoch * per_call - (per_call - 1u)) * (unsigned int)grid;
    p.flag_target2 = (unsigned int)a->epoch * per_call * (unsigned int)grid;
  }
  char* ws = (char*)a->workspace;
  p.rowpart = (float2*)ws; ws += (size_t)p.P * p.WB * p.nJ64 * 8;
  p.colpart = (float2*)ws; ws += (size_t)p.P * p.WB * p.nI32 * 8;
  p.gpart = (float*)ws;    ws += (size_t)p.P * 2 * p.nJ64 * p.B * p.D * 4;
  p.spart = (float*)ws;
  p.out = a->out;
  p.status = a->status;
  {
    double secs = a->peer_timeout_s;
    if (secs <= 0.0) { const char* e = getenv("MCLIP_PEER_TIMEOUT_S"); secs = e ? atof(e) : 600.0; }
    if (secs <= 0.0) secs = 600.0;
    static int khz = 0;                              // cudaDevAttrClockRate is a slow driver query (~1 ms): once per process
    if (khz == 0) {
        int dev = 0, v = 0;
        cudaGetDevice(&dev);
        khz = (cudaDeviceGetAttribute(&v, cudaDevAttrClockRate, dev) == cudaSuccess && v > 0) ? v : 2000000;
    }
    p.timeout_cycles = (long long)(secs * (double)khz * 1e3);
  }
  p.window = g_loss_window;
  void* kargs[] = {(void*)&p};
  MCLIP_CHECK_CUDA(cudaLaunchCooperativeKernel((void*)mclip_loss_kernel, dim3(grid), dim3(LOSS_THREADS), kargs, 0, stream));
  return MCLIP_OK;
}
