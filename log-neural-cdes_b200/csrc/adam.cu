// K4 – fused Adam update on the flat parameter buffer (optax.adam defaults as used at
// train.py:183: b1 0.9, b2 0.999, eps 1e-8, bias-corrected, eps outside the sqrt-hat).
#include <cuda_runtime.h>

#include "lncde_b200.h"

namespace lncde {
__global__ void adam_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m,
                            float* __restrict__ v, long long n, float lr, float b1, float b2, float eps,
                            float c1, float c2, float gs) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n;
       i += (long long)gridDim.x * blockDim.x) {
    const float gi = g[i] * gs;
    const float mi = b1 * m[i] + (1.f - b1) * gi;
    const float vi = b2 * v[i] + (1.f - b2) * gi * gi;
    m[i] = mi;
    v[i] = vi;
    // optax scale_by_adam: mu_hat / (sqrt(nu_hat) + eps)
    p[i] -= lr * (mi / c1) / (sqrtf(vi / c2) + eps);
  }
}

// Graph-capturable variant: the step count lives on the device so a captured launch stays valid.
__global__ void adam_dev_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m,
                                float* __restrict__ v, long long n, float lr, float b1, float b2, float eps,
                                const int* __restrict__ step_counter, float gs) {
  const float step = (float)(*step_counter + 1);
  const float c1 = 1.f - powf(b1, step), c2 = 1.f - powf(b2, step);
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n;
       i += (long long)gridDim.x * blockDim.x) {
    const float gi = g[i] * gs;
    const float mi = b1 * m[i] + (1.f - b1) * gi;
    const float vi = b2 * v[i] + (1.f - b2) * gi * gi;
    m[i] = mi;
    v[i] = vi;
    p[i] -= lr * (mi / c1) / (sqrtf(vi / c2) + eps);
  }
}
__global__ void bump_counter_kernel(int* c) { *c += 1; }
}  // namespace lncde

extern "C" int lncde_adam_step(void* stream, float* params, const float* grads, float* m, float* v,
                               int64_t n, float lr, float beta1, float beta2, float eps, int step,
                               float grad_scale) {
  if (!params || !grads || !m || !v) return LNCDE_ERR_NULL;
  if (n < 1 || step < 1) return LNCDE_ERR_SHAPE;
  const float c1 = 1.f - powf(beta1, (float)step);
  const float c2 = 1.f - powf(beta2, (float)step);
  const int threads = 256;
  long long blocks = (n + threads - 1) / threads;
  if (blocks > 148 * 8) blocks = 148 * 8;
  lncde::adam_kernel<<<(int)blocks, threads, 0, (cudaStream_t)stream>>>(params, grads, m, v, n, lr, beta1,
                                                                        beta2, eps, c1, c2, grad_scale);
  return (int)cudaGetLastError();
}

extern "C" int lncde_adam_step_dev(void* stream, float* params, const float* grads, float* m, float* v, int64_t n,
                                   float lr, float beta1, float beta2, float eps, int* step_counter,
                                   float grad_scale) {
  if (!params || !grads || !m || !v || !step_counter) return LNCDE_ERR_NULL;
  if (n < 1) return LNCDE_ERR_SHAPE;
  const int threads = 256;
  long long blocks = (n + threads - 1) / threads;
  if (blocks > 148 * 8) blocks = 148 * 8;
  cudaStream_t cs = (cudaStream_t)stream;
  lncde::adam_dev_kernel<<<(int)blocks, threads, 0, cs>>>(params, grads, m, v, n, lr, beta1, beta2, eps,
                                                          step_counter, grad_scale);
  lncde::bump_counter_kernel<<<1, 1, 0, cs>>>(step_counter);
  return (int)cudaGetLastError();
}
