// Internal: the public C ABI plus shared host declarations.
#pragma once
#include "../../include/mclip.h"
#define MCLIP_ABI_VERSION 3
