// Internal: the public C ABI plus shared host declarations.
#pragma once
#include "../../include/mclip.h"
#define MCLIP_ABI_VERSION 4

// row-streaming depthwise kernels (dwstream.cu); conv.cu dispatches to them for the shapes they cover
bool mclip_dws_covers(const mclip_dwconv_args* a, int backward);
int mclip_dws_slots(const mclip_dwconv_args* a, int backward);
int mclip_dws_forward(const mclip_dwconv_args* a, void* stream);
int mclip_dws_backward(const mclip_dwconv_args* a, void* stream);
