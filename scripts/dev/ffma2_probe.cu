// Dev probe: issue rate of FFMA vs FFMA2 (fma.rn.f32x2, sm_100) on one B200, independent chains.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o /tmp/ffma2_probe scripts/dev/ffma2_probe.cu && /tmp/ffma2_probe
#include <cstdio>
#include <cuda_runtime.h>

template <int MODE>
__global__ void __launch_bounds__(256) probe(float* out, int iters) {
  float2 a[8];
  const float2 m = make_float2(0.999f, 0.998f), c = make_float2(0.001f, 0.002f);
#pragma unroll
  for (int k = 0; k < 8; ++k) a[k] = make_float2(threadIdx.x * 1e-3f + k, k);
  for (int i = 0; i < iters; ++i) {
#pragma unroll
    for (int r = 0; r < 4; ++r) {
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        if (MODE == 0) {
          a[k].x = fmaf(a[k].x, m.x, c.x);
          a[k].y = fmaf(a[k].y, m.y, c.y);
        } else {
          a[k] = __ffma2_rn(a[k], m, c);
        }
      }
    }
  }
  float s = 0.f;
#pragma unroll
  for (int k = 0; k < 8; ++k) s += a[k].x + a[k].y;
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

int main() {
  float* out;
  const int blocks = 148 * 8, iters = 4096;
  cudaMalloc(&out, blocks * 256 * sizeof(float));
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  for (int mode = 0; mode < 2; ++mode) {
    for (int rep = 0; rep < 3; ++rep) {
      cudaEventRecord(e0);
      if (mode == 0) probe<0><<<blocks, 256>>>(out, iters);
      else probe<1><<<blocks, 256>>>(out, iters);
      cudaEventRecord(e1);
      cudaEventSynchronize(e1);
      float ms;
      cudaEventElapsedTime(&ms, e0, e1);
      const double flop = 2.0 * 2 * 8 * 4 * (double)iters * blocks * 256;
      printf("%s rep %d: %.3f ms  %.1f TFLOP/s\n", mode ? "FFMA2" : "FFMA ", rep, ms, flop / ms * 1e-9);
    }
  }
  printf("%s\n", cudaGetErrorString(cudaGetLastError()));
  return 0;
}
