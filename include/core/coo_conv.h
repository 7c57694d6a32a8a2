/* Compatibility shim: the reference ships its ABI as src/core/coo_conv.h; here every
 * declaration lives in ../spgpu.h (see the citations there). */
#pragma once
#include <string.h>
#include "../spgpu.h"
