// Internal: the public C ABI plus shared host declarations.
#pragma once
#include "../../include/mclip.h"
#define MCLIP_ABI_VERSION 5

// row-streaming depthwise kernels (dwstream.cu); conv.cu dispatches to them for the shapes they cover
bool mclip_dws_covers(const mclip_dwconv_args* a, int backward);
int mclip_dws_slots(const mclip_dwconv_args* a, int backward);
int mclip_dws_forward(const mclip_dwconv_args* a, void* stream);
int mclip_dws_backward(const mclip_dwconv_args* a, void* stream);

// tensor-core self-attention (attention.cu); bert.cu dispatches to it for seq_len <= 256, head_dim 64 (MCLIP_ATT_TC=0 keeps the SIMT kernels)
bool mclip_att_tc_covers(int seq_len, int heads, int head_dim);
int mclip_att_tc_forward(const void* qkv, const void* amask, const void* dropmask, float drop_scale, void* out, float* lse, int batch, int seq_len,
                         int heads, void* stream);
int mclip_att_tc_backward(const void* qkv, const void* d_out, const float* lse, const void* amask, const void* dropmask, float drop_scale, void* dqkv,
                          int batch, int seq_len, int heads, void* stream);
