// TEMPORARY: entry points not written yet report "unsupported" (never a CPU fallback).
#include "lncde_b200.h"
extern "C" {
int lncde_linear_scan_fwd(void*, const float*, const float*, const float*, float*, int, int, int, int, int, int) { return LNCDE_ERR_UNSUPPORTED; }
size_t lncde_linear_scan_bwd_workspace_bytes(int, int, int, int, int) { return 0; }
int lncde_linear_scan_bwd(void*, const float*, const float*, const float*, const float*, float*, float*, int, int, int, int, int, int, void*, size_t) { return LNCDE_ERR_UNSUPPORTED; }
}
