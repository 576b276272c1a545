// BERT encoder forward pieces that are not GEMMs (the Linear layers run on mclip_gemm_tn with fused
// bias / GELU / dropout / residual epilogues): embeddings + LayerNorm, LayerNorm, masked softmax attention.
// Replaces what `HuggingfaceTextEncoder.forward` (text_encoder.py:47-49) executes inside transformers' BertModel
// (modeling_bert.py: BertEmbeddings, BertSelfAttention via SDPA, BertSelfOutput/BertOutput LayerNorms), post-LN BERT,
// LayerNorm eps 1e-12, attention scale 1/sqrt(64), additive padding mask, dropout p on embeddings / probs / sub-layer outputs.
#include "common.cuh"
#include "mclip_internal.h"
#include <math.h>

// ---- one warp per token: out = dropout(LN(word[id] + pos[l] + type[tt])) ---------------------------------------------
template <int H>
__global__ void __launch_bounds__(128) mclip_bert_embed_ln_kernel(const long long* __restrict__ ids, const long long* __restrict__ tts,
                                                                  const float* __restrict__ word, const float* __restrict__ pos,
                                                                  const float* __restrict__ type, const float* __restrict__ gamma,
                                                                  const float* __restrict__ beta, float eps, const uint8_t* __restrict__ dropmask,
                                                                  float drop_scale, bf16* __restrict__ out, int tokens, int L, int vocab) {
  constexpr int PER = H / 32;
  const int tok = blockIdx.x * 4 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  if (tok >= tokens) return;
  long long id = ids[tok];
  if (id < 0 || id >= vocab) id = 0;
  const long long tt = tts ? tts[tok] : 0;
  const int l = tok % L;
  float v[PER];
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < PER; ++i) {
    const int h = lane + i * 32;
    v[i] = word[(size_t)id * H + h] + pos[(size_t)l * H + h] + type[(size_t)tt * H + h];
    s += v[i];
  }
  const float mean = warp_sum(s) * (1.0f / H);
  float q = 0.f;
#pragma unroll
  for (int i = 0; i < PER; ++i) { const float d = v[i] - mean; q = fmaf(d, d, q); }
  const float rstd = rsqrtf(warp_sum(q) * (1.0f / H) + eps);
#pragma unroll
  for (int i = 0; i < PER; ++i) {
    const int h = lane + i * 32;
    float o = (v[i] - mean) * rstd * gamma[h] + beta[h];
    if (dropmask) o *= dropmask[(size_t)tok * H + h] ? drop_scale : 0.f;
    out[(size_t)tok * H + h] = __float2bfloat16_rn(o);
  }
}

// ---- one warp per row LayerNorm over bf16 input ----------------------------------------------------------------------
template <int H>
__global__ void __launch_bounds__(128) mclip_layernorm_kernel(const bf16* __restrict__ x, const float* __restrict__ gamma, const float* __restrict__ beta,
                                                              float eps, bf16* __restrict__ out, int rows) {
  constexpr int PER = H / 32;
  const int row = blockIdx.x * 4 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  if (row >= rows) return;
  float v[PER];
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < PER; ++i) { v[i] = __bfloat162float(x[(size_t)row * H + lane + i * 32]); s += v[i]; }
  const float mean = warp_sum(s) * (1.0f / H);
  float q = 0.f;
#pragma unroll
  for (int i = 0; i < PER; ++i) { const float d = v[i] - mean; q = fmaf(d, d, q); }
  const float rstd = rsqrtf(warp_sum(q) * (1.0f / H) + eps);
#pragma unroll
  for (int i = 0; i < PER; ++i) {
    const int h = lane + i * 32;
    out[(size_t)row * H + h] = __float2bfloat16_rn((v[i] - mean) * rstd * gamma[h] + beta[h]);
  }
}

// ---- attention: CTA = (batch, head, 64-query block); 4 threads per query row, online softmax over 64-key blocks -----
#define ATT_D 64
#define ATT_BQ 64
#define ATT_BK 64
__global__ void __launch_bounds__(256) mclip_bert_attention_kernel(const bf16* __restrict__ qkv, const long long* __restrict__ amask,
                                                                   const uint8_t* __restrict__ dropmask, float drop_scale, bf16* __restrict__ out,
                                                                   int B, int L, int heads) {
  __shared__ float Ks[ATT_BK][ATT_D + 1];
  __shared__ float Vs[ATT_BK][ATT_D + 1];
  __shared__ int kvalid[ATT_BK];
  const int H = heads * ATT_D;
  const int qb = blockIdx.x, head = blockIdx.y, b = blockIdx.z;
  const int row = threadIdx.x >> 2, part = threadIdx.x & 3;
  const int qi = qb * ATT_BQ + row;
  const bool qvalid = qi < L;
  float q[ATT_D];
  {
    const bf16* qp = qkv + ((size_t)(b * L + (qvalid ? qi : 0))) * 3 * H + head * ATT_D;
#pragma unroll
    for (int d = 0; d < ATT_D; d += 8) {
      float f[8];
      unpack8(*reinterpret_cast<const bf16x8*>(qp + d), f);
#pragma unroll
      for (int i = 0; i < 8; ++i) q[d + i] = f[i] * 0.125f;     // 1/sqrt(64)
    }
  }
  float m = -INFINITY, lsum = 0.f, acc[16];
#pragma unroll
  for (int i = 0; i < 16; ++i) acc[i] = 0.f;
  for (int k0 = 0; k0 < L; k0 += ATT_BK) {
    __syncthreads();
    for (int i = threadIdx.x; i < ATT_BK * (ATT_D / 8); i += 256) {
      const int kr = i / (ATT_D / 8), dv = (i % (ATT_D / 8)) * 8;
      const int kj = k0 + kr;
      float fk[8], fv[8];
      if (kj < L) {
        const bf16* base = qkv + ((size_t)(b * L + kj)) * 3 * H + head * ATT_D + dv;
        unpack8(*reinterpret_cast<const bf16x8*>(base + H), fk);
        unpack8(*reinterpret_cast<const bf16x8*>(base + 2 * H), fv);
      } else {
#pragma unroll
        for (int e = 0; e < 8; ++e) fk[e] = fv[e] = 0.f;
      }
#pragma unroll
      for (int e = 0; e < 8; ++e) { Ks[kr][dv + e] = fk[e]; Vs[kr][dv + e] = fv[e]; }
    }
    if (threadIdx.x < ATT_BK) { const int kj = k0 + threadIdx.x; kvalid[threadIdx.x] = (kj < L) && (amask[(size_t)b * L + kj] != 0); }
    __syncthreads();
    // scores of this thread's 16 keys
    float s[16];
    float bm = -INFINITY;
#pragma unroll
    for (int j = 0; j < 16; ++j) {
      const int kr = part * 16 + j;
      float a = 0.f;
#pragma unroll
      for (int d = 0; d < ATT_D; ++d) a = fmaf(q[d], Ks[kr][d], a);
      s[j] = kvalid[kr] ? a : -INFINITY;
      bm = fmaxf(bm, s[j]);
    }
    bm = fmaxf(bm, __shfl_xor_sync(0xffffffffu, bm, 1));
    bm = fmaxf(bm, __shfl_xor_sync(0xffffffffu, bm, 2));
    const float mnew = fmaxf(m, bm);
    if (mnew == -INFINITY) continue;                       // whole block masked for this row (quad-uniform)
    const float corr = __expf(m - mnew);
    float ps = 0.f;
#pragma unroll
    for (int j = 0; j < 16; ++j) { s[j] = __expf(s[j] - mnew); ps += s[j]; }
    ps += __shfl_xor_sync(0xffffffffu, ps, 1);
    ps += __shfl_xor_sync(0xffffffffu, ps, 2);
    lsum = lsum * corr + ps;
    m = mnew;
    if (dropmask && qvalid) {
      const uint8_t* dm = dropmask + (((size_t)(b * heads + head)) * L + qi) * L + k0 + part * 16;
#pragma unroll
      for (int j = 0; j < 16; ++j) if (k0 + part * 16 + j < L) s[j] *= dm[j] ? drop_scale : 0.f;
    }
#pragma unroll
    for (int i = 0; i < 16; ++i) acc[i] *= corr;
    // acc[d in my 16 dims] += sum over all 64 keys p * V ; probabilities of the other parts come by quad shuffles
#pragma unroll
    for (int src = 0; src < 4; ++src) {
#pragma unroll
      for (int j = 0; j < 16; ++j) {
        const float pj = __shfl_sync(0xffffffffu, s[j], (threadIdx.x & 28) | src, 32);
        const int kr = src * 16 + j;
#pragma unroll
        for (int i = 0; i < 16; ++i) acc[i] = fmaf(pj, Vs[kr][part * 16 + i], acc[i]);
      }
    }
  }
  if (qvalid) {
    const float inv = lsum > 0.f ? 1.0f / lsum : 0.f;
    float o[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) o[i] = acc[i] * inv;
    bf16* op = out + ((size_t)(b * L + qi)) * H + head * ATT_D + part * 16;
    *reinterpret_cast<bf16x8*>(op) = pack8(o);
    *reinterpret_cast<bf16x8*>(op + 8) = pack8(o + 8);
  }
}

extern "C" int mclip_bert_embed_ln(const mclip_bert_embed_args* a, void* stream) {
  MCLIP_REQUIRE(a && a->input_ids && a->word && a->pos && a->type && a->gamma && a->beta && a->out, "mclip_bert_embed_ln: null operand");
  MCLIP_REQUIRE(a->hidden == 768, "mclip_bert_embed_ln: hidden size %d not built (768 only)", a->hidden);
  MCLIP_REQUIRE(a->seq_len <= a->max_positions, "mclip_bert_embed_ln: seq_len %d exceeds position table %d", a->seq_len, a->max_positions);
  const int tokens = a->batch * a->seq_len;
  mclip_bert_embed_ln_kernel<768><<<ceil_div(tokens, 4), 128, 0, (cudaStream_t)stream>>>(
      (const long long*)a->input_ids, (const long long*)a->token_type_ids, a->word, a->pos, a->type, a->gamma, a->beta, a->eps,
      (const uint8_t*)a->dropmask, a->drop_scale, (bf16*)a->out, tokens, a->seq_len, a->vocab);
  MCLIP_CHECK_LAUNCH();
  return MCLIP_OK;
}

extern "C" int mclip_layernorm(const void* x, const float* gamma, const float* beta, float eps, void* out, int rows, int hidden, void* stream) {
  MCLIP_REQUIRE(x && gamma && beta && out && rows > 0, "mclip_layernorm: bad arguments");
  if (hidden == 768) mclip_layernorm_kernel<768><<<ceil_div(rows, 4), 128, 0, (cudaStream_t)stream>>>((const bf16*)x, gamma, beta, eps, (bf16*)out, rows);
  else if (hidden == 512) mclip_layernorm_kernel<512><<<ceil_div(rows, 4), 128, 0, (cudaStream_t)stream>>>((const bf16*)x, gamma, beta, eps, (bf16*)out, rows);
  else { mclip_set_error("mclip_layernorm: hidden size %d not built (768, 512)", hidden); return MCLIP_ERR_INVALID; }
  MCLIP_CHECK_LAUNCH();
  return MCLIP_OK;
}

extern "C" int mclip_bert_attention(const void* qkv, const void* attention_mask, const void* dropmask, float drop_scale, void* out, int batch,
                                    int seq_len, int heads, int head_dim, void* stream) {
  MCLIP_REQUIRE(qkv && attention_mask && out && batch > 0 && seq_len > 0, "mclip_bert_attention: bad arguments");
  MCLIP_REQUIRE(head_dim == ATT_D, "mclip_bert_attention: head_dim %d not built (64 only)", head_dim);
  dim3 grid(ceil_div(seq_len, ATT_BQ), heads, batch);
  mclip_bert_attention_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>((const bf16*)qkv, (const long long*)attention_mask, (const uint8_t*)dropmask, drop_scale,
                                                                      (bf16*)out, batch, seq_len, heads);
  MCLIP_CHECK_LAUNCH();
  return MCLIP_OK;
}
