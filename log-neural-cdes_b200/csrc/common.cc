// Version / status strings of the C ABI (include/lncde_b200.h).
#include <cuda_runtime_api.h>

#include <atomic>
#include <cstdlib>
#include <cstring>

#include "lncde_b200.h"
#include "tuning.h"

namespace lncde {
namespace {
const char* const kTuneNames[kTuneCount] = {"k1_tma", "k3_fused", "k2_fused_d", "k1_grid_cap", "k2_de_v2", "k3_stream",
                                            "k2_pscan"};
const char* const kTuneEnv[kTuneCount] = {"LNCDE_TUNE_K1_TMA", "LNCDE_TUNE_K3_FUSED", "LNCDE_TUNE_K2_FUSED_D",
                                          "LNCDE_TUNE_K1_GRID_CAP", "LNCDE_TUNE_K2_DE_V2", "LNCDE_TUNE_K3_STREAM",
                                          "LNCDE_TUNE_K2_PSCAN"};
std::atomic<int> g_tune[kTuneCount] = {{-1}, {-1}, {-1}, {-1}, {-1}, {-1}, {-1}};  // -1: not set yet -> environment, else 0
}  // namespace

int tuning_value(int key) {
  if (key < 0 || key >= kTuneCount) return 0;
  int v = g_tune[key].load(std::memory_order_relaxed);
  if (v < 0) {
    const char* e = std::getenv(kTuneEnv[key]);
    v = (e && e[0] >= '0' && e[0] <= '9') ? std::atoi(e) : 0;
    int expected = -1;
    g_tune[key].compare_exchange_strong(expected, v);  // a concurrent lncde_set_tuning wins
    v = g_tune[key].load(std::memory_order_relaxed);
  }
  return v;
}
}  // namespace lncde

extern "C" {

int lncde_set_tuning(const char* key, int value) {
  if (!key) return LNCDE_ERR_NULL;
  for (int k = 0; k < lncde::kTuneCount; ++k)
    if (std::strcmp(key, lncde::kTuneNames[k]) == 0) {
      lncde::g_tune[k].store(value < 0 ? 0 : value);
      return LNCDE_OK;
    }
  return LNCDE_ERR_UNSUPPORTED;
}

int lncde_get_tuning(const char* key) {
  if (!key) return LNCDE_ERR_NULL;
  for (int k = 0; k < lncde::kTuneCount; ++k)
    if (std::strcmp(key, lncde::kTuneNames[k]) == 0) return lncde::tuning_value(k);
  return LNCDE_ERR_UNSUPPORTED;
}

const char* lncde_version(void) { return "lncde_b200 0.1 (sm_100a)"; }

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
