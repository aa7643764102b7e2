/* CPU shim (see cuda_runtime.h in this directory) — TEST INFRASTRUCTURE ONLY. */
#include "cuda_runtime.h"
