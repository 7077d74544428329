// Measurement helper: FP32 FFMA issue-rate probe.  MEASURED_PEAKS.json carries HBM and bf16 tensor peaks only; the
// Log-NCDE solve kernels (K3 / K3b) are FP32-issue / latency bound, so bench.py quotes them against the FFMA rate
// this kernel reaches on the same device at the same clocks ("measure a pure-FFMA loop once").
#include <cuda_runtime.h>

#include "lncde_b200.h"

namespace lncde {
constexpr int kProbeChains = 8;  // independent accumulators per thread: hides the 4-cycle FFMA latency at 8 warps / SMSP

__global__ void __launch_bounds__(256) ffma_probe_kernel(float* __restrict__ out, int iters, float a, float b) {
  float x[kProbeChains];
#pragma unroll
  for (int k = 0; k < kProbeChains; ++k) x[k] = 1e-3f * (float)(threadIdx.x + k);
  for (int i = 0; i < iters; ++i) {
#pragma unroll
    for (int u = 0; u < 4; ++u)
#pragma unroll
      for (int k = 0; k < kProbeChains; ++k) x[k] = fmaf(x[k], a, b);
  }
  float s = 0.f;
#pragma unroll
  for (int k = 0; k < kProbeChains; ++k) s += x[k];
  out[blockIdx.x * (size_t)blockDim.x + threadIdx.x] = s;  // keeps the chains live
}
}  // namespace lncde

extern "C" int64_t lncde_probe_ffma_flops(int blocks, int iters) {
  if (blocks < 1 || iters < 1) return LNCDE_ERR_SHAPE;
  return (int64_t)blocks * 256 * (int64_t)iters * 4 * lncde::kProbeChains * 2;
}

extern "C" int lncde_probe_ffma(void* stream, float* out, int blocks, int iters) {
  if (!out) return LNCDE_ERR_NULL;
  if (blocks < 1 || iters < 1) return LNCDE_ERR_SHAPE;
  // a, b chosen so the chains stay finite and the compiler cannot fold them (x -> 0.999 x + 0.001 converges to 1)
  lncde::ffma_probe_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(out, iters, 0.999f, 0.001f);
  return (int)cudaGetLastError();
}
