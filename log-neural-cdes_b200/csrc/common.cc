// Version / status strings of the C ABI (include/lncde_b200.h).
#include <cuda_runtime_api.h>

#include "lncde_b200.h"

extern "C" {

const char* lncde_version(void) { return "lncde_b200 0.2 (sm_100a)"; }

const char* lncde_strerror(int status) {
  switch (status) {
    case LNCDE_OK: return "ok";
    case LNCDE_ERR_NULL: return "required pointer is NULL";
    case LNCDE_ERR_SHAPE: return "invalid or inconsistent shape";
    case LNCDE_ERR_UNSUPPORTED: return "configuration outside the compiled range";
    case LNCDE_ERR_WORKSPACE: return "workspace too small";
    case LNCDE_ERR_ALIGN: return "pointer not 16-byte aligned";
    default: break;
  }
  if (status > 0) return cudaGetErrorString((cudaError_t)status);
  return "unknown lncde status";
}

}  // extern "C"
