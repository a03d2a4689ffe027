// Host-side helpers shared by the C-ABI translation units: error slot, driver entry points
// (cuTensorMapEncodeTiled is fetched at run time so the library links against cudart only),
// device property cache.
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdarg.h>
#include <stdint.h>
#include <stdio.h>

#include "../../include/pmoe_b200.h"

namespace pmoe {

void set_error(const char* fmt, ...);
int num_sms();

// Encode a tiled tensor map (rank <= 5). dims/box in elements (innermost first), strides in BYTES
// for dims 1..rank-1. Returns 0 or a PMOE_ERR_* code.
int encode_tmap(CUtensorMap* out, CUtensorMapDataType dtype, int rank, void* base, const uint64_t* dims,
                const uint64_t* strides_bytes, const uint32_t* box, CUtensorMapSwizzle swz);

inline int check_launch(const char* what) {
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) {
    set_error("%s: launch failed: %s", what, cudaGetErrorString(e));
    return PMOE_ERR_LAUNCH;
  }
  return PMOE_OK;
}

}  // namespace pmoe
