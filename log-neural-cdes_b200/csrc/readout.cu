// K5: the pieces of a classification train step around the solve - initial state, read-out + cross-entropy, Lip(2)
// regulariser - with their gradients, as a handful of small kernels instead of ~100 eager torch / cuBLAS launches.
//
//   y0   = linear1(x0)                                   LogNeuralCDEs.py:139  /  init_layer, LinearNeuralCDEs.py:303
//   pred = softmax(linear2(h))                           LogNeuralCDEs.py:159-160 (h = y(t1)); LinearNeuralCDEs.py:388-391
//   loss = mean_b(-sum_c y log(pred + 1e-8))             train.py:95
//        + lambd * sum_layers (mean_rows ||W_row|| + ||b||)   train.py:79-93 (weights-only entries for the LNCDE's vf_A)
//
// Everything is tiny (B x H = 32 x 128, <= 8 layers of <= 37 k parameters) and sits between two long kernels: the
// point is launch count and determinism (fixed summation orders, no atomics), not bandwidth.
#include <cuda_runtime.h>

#include <cstdint>

#include "lncde_b200.h"

namespace lncde {

// ---------------------------------------------------------------------------------------------- affine layer
__global__ void __launch_bounds__(256) affine_fwd_kernel(const float* __restrict__ x, const float* __restrict__ W,
                                                          const float* __restrict__ b, float* __restrict__ y, int B, int n_in,
                                                          int n_out) {
  const long long total = (long long)B * n_out;
  for (long long e = blockIdx.x * (long long)blockDim.x + threadIdx.x; e < total; e += (long long)gridDim.x * blockDim.x) {
    const int o = (int)(e % n_out);
    const long long s = e / n_out;
    float acc = b[o];
    for (int c = 0; c < n_in; ++c) acc = fmaf(x[s * n_in + c], W[(size_t)o * n_in + c], acc);
    y[e] = acc;
  }
}

// Sums over the batch are split into chunks of kChunk samples (grid.y), each chunk summed in ascending sample order into
// its own partial, the partials added in ascending chunk order by chunk_reduce_kernel: deterministic for any B, and a
// batch of 2 048 does not leave one thread with 2 048 dependent loads.  A single chunk writes its result directly.
constexpr int kChunk = 32;

__global__ void __launch_bounds__(256) chunk_reduce_kernel(const float* __restrict__ part, float* __restrict__ out, int n,
                                                            int stride, int chunks, float scale) {
  for (int e = blockIdx.x * blockDim.x + threadIdx.x; e < n; e += gridDim.x * blockDim.x) {
    float acc = 0.f;
    for (int c = 0; c < chunks; ++c) acc += part[(size_t)c * stride + e];
    out[e] = acc * scale;
  }
}

// g_W[o][c] = sum_b g_y[b][o] x[b][c], g_b[o] = sum_b g_y[b][o]; one thread per (o, c) entry (+ one per bias) and chunk.
// dst: [chunks][n_out * n_in + n_out] partials, or (one chunk) g_W directly followed by g_b at dst + n_out * n_in.
__global__ void __launch_bounds__(256) affine_bwd_kernel(const float* __restrict__ x, const float* __restrict__ g_y,
                                                          float* __restrict__ dst, int B, int n_in, int n_out) {
  const int nW = n_out * n_in, nE = nW + n_out;
  const int s0 = blockIdx.y * kChunk, s1 = min(B, s0 + kChunk);
  float* out = dst + (size_t)blockIdx.y * nE;
  for (int e = blockIdx.x * blockDim.x + threadIdx.x; e < nE; e += gridDim.x * blockDim.x) {
    float acc = 0.f;
    if (e < nW) {
      const int o = e / n_in, c = e - o * n_in;
      for (int s = s0; s < s1; ++s) acc = fmaf(g_y[(size_t)s * n_out + o], x[(size_t)s * n_in + c], acc);
    } else {
      const int o = e - nW;
      for (int s = s0; s < s1; ++s) acc += g_y[(size_t)s * n_out + o];
    }
    out[e] = acc;
  }
}

// ---------------------------------------------------------------------------------------------- read-out + CE
constexpr int kMaxLabels = 32;

// one warp per sample: logits, softmax, loss term, d loss / d logits, d loss / d h
__global__ void __launch_bounds__(128) readout_ce_sample_kernel(const float* __restrict__ h, long long h_stride,
                                                                 const float* __restrict__ W, const float* __restrict__ b,
                                                                 const float* __restrict__ labels, float* __restrict__ pred,
                                                                 float* __restrict__ g_h, long long g_h_stride,
                                                                 float* __restrict__ g_z, float* __restrict__ loss_b, int B,
                                                                 int H, int L) {
  const int lane = threadIdx.x & 31;
  const int s = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (s >= B) return;
  const float* hs = h + (size_t)s * h_stride;
  float z[kMaxLabels];
#pragma unroll 1
  for (int l = 0; l < L; ++l) {
    float acc = 0.f;
    for (int k = lane; k < H; k += 32) acc = fmaf(W[(size_t)l * H + k], hs[k], acc);
#pragma unroll
    for (int m = 16; m >= 1; m >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, m);
    z[l] = acc + b[l];
  }
  float mx = z[0];
  for (int l = 1; l < L; ++l) mx = fmaxf(mx, z[l]);
  float den = 0.f;
  for (int l = 0; l < L; ++l) {
    z[l] = expf(z[l] - mx);
    den += z[l];
  }
  const float inv = 1.f / den, invB = 1.f / (float)B;
  float lossv = 0.f, dot = 0.f;  // dot = sum_k (dloss / dp_k) p_k
  for (int l = 0; l < L; ++l) {
    const float p = z[l] * inv, yl = labels[(size_t)s * L + l];
    z[l] = p;
    lossv -= yl * logf(p + 1e-8f);
    dot += (-yl / (p + 1e-8f)) * p;
  }
  for (int l = 0; l < L; ++l) {
    const float p = z[l], yl = labels[(size_t)s * L + l];
    const float gz = p * ((-yl / (p + 1e-8f)) - dot) * invB;  // softmax Jacobian applied to d loss / d pred
    if (lane == 0) {
      if (pred) pred[(size_t)s * L + l] = p;
      g_z[(size_t)s * L + l] = gz;
    }
    z[l] = gz;
  }
  if (lane == 0) loss_b[s] = lossv;
  float* gs = g_h + (size_t)s * g_h_stride;
  for (int k = lane; k < H; k += 32) {
    float acc = 0.f;
    for (int l = 0; l < L; ++l) acc = fmaf(z[l], W[(size_t)l * H + k], acc);
    gs[k] = acc;
  }
}

// per chunk of samples: [g_W (L x H) | g_b (L) | sum of loss_b] partials (or, one chunk, the results themselves with the
// loss divided by B): g_W = g_z^T h, g_b = sum_b g_z, sample order ascending
__global__ void __launch_bounds__(256) readout_ce_reduce_kernel(const float* __restrict__ h, long long h_stride,
                                                                 const float* __restrict__ g_z, const float* __restrict__ loss_b,
                                                                 float* __restrict__ dst, int B, int H, int L, float loss_scale) {
  const int nE = L * H + L + 1;
  const int s0 = blockIdx.y * kChunk, s1 = min(B, s0 + kChunk);
  float* out = dst + (size_t)blockIdx.y * nE;
  for (int e = blockIdx.x * blockDim.x + threadIdx.x; e < nE; e += gridDim.x * blockDim.x) {
    float acc = 0.f;
    if (e < L * H) {
      const int l = e / H, k = e - l * H;
      for (int s = s0; s < s1; ++s) acc = fmaf(g_z[(size_t)s * L + l], h[(size_t)s * h_stride + k], acc);
    } else if (e < L * H + L) {
      const int l = e - L * H;
      for (int s = s0; s < s1; ++s) acc += g_z[(size_t)s * L + l];
    } else {
      for (int s = s0; s < s1; ++s) acc += loss_b[s];
      acc *= loss_scale;
    }
    out[e] = acc;
  }
}

// ---------------------------------------------------------------------------------------------- Lip(2) regulariser
constexpr int kMaxLipLayers = 8;
struct LipParams {
  int n_layers;
  int w_off[kMaxLipLayers], b_off[kMaxLipLayers];  // b_off < 0: weights only
  int rows[kMaxLipLayers], cols[kMaxLipLayers];
  float lambd;
};

// one CTA, one warp per row in turn; loss += lambd * sum_l (mean_r ||W_r|| + ||b||), g += its gradient
__global__ void __launch_bounds__(1024) lip2_kernel(const float* __restrict__ params, float* __restrict__ loss,
                                                    float* __restrict__ g_params, const LipParams p) {
  __shared__ float warp_sum[32];
  __shared__ float total_s, row_norm;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
  if (threadIdx.x == 0) total_s = 0.f;
  __syncthreads();
  for (int l = 0; l < p.n_layers; ++l) {
    const float* Wl = params + p.w_off[l];
    float* gW = g_params + p.w_off[l];
    const int R = p.rows[l], Cn = p.cols[l];
    float mine = 0.f;  // sum of the row norms this warp handled
    if (Cn >= 1024) {
      // few long rows (the LNCDE generators: C rows of nb * bs * bs entries): the whole CTA per row, partials added in
      // warp order - a warp per row would leave all but R warps idle
      for (int r = 0; r < R; ++r) {
        const float* row = Wl + (size_t)r * Cn;
        float ss = 0.f;
#pragma unroll 4
        for (int c = threadIdx.x; c < Cn; c += blockDim.x) ss = fmaf(row[c], row[c], ss);
#pragma unroll
        for (int m = 16; m >= 1; m >>= 1) ss += __shfl_xor_sync(0xffffffffu, ss, m);
        __syncthreads();  // warp_sum / row_norm of the previous row have been read
        if (lane == 0) warp_sum[warp] = ss;
        __syncthreads();
        if (threadIdx.x == 0) {
          float t = 0.f;
          for (int w = 0; w < nw; ++w) t += warp_sum[w];
          row_norm = sqrtf(t);
        }
        __syncthreads();
        const float nrm = row_norm;
        if (threadIdx.x == 0) mine += nrm;
        const float sc = nrm > 0.f ? p.lambd / ((float)R * nrm) : 0.f;
#pragma unroll 4
        for (int c = threadIdx.x; c < Cn; c += blockDim.x) gW[(size_t)r * Cn + c] += sc * row[c];
      }
      __syncthreads();
    } else
    for (int r = warp; r < R; r += nw) {
      float ss = 0.f;
      for (int c = lane; c < Cn; c += 32) ss = fmaf(Wl[(size_t)r * Cn + c], Wl[(size_t)r * Cn + c], ss);
#pragma unroll
      for (int m = 16; m >= 1; m >>= 1) ss += __shfl_xor_sync(0xffffffffu, ss, m);
      const float nrm = sqrtf(ss);
      mine += nrm;
      const float sc = nrm > 0.f ? p.lambd / ((float)R * nrm) : 0.f;
      for (int c = lane; c < Cn; c += 32) gW[(size_t)r * Cn + c] += sc * Wl[(size_t)r * Cn + c];
    }
    if (lane == 0) warp_sum[warp] = mine;
    __syncthreads();
    if (warp == 0) {
      float layer = 0.f;
      if (lane == 0) {
        for (int w = 0; w < nw; ++w) layer += warp_sum[w];
        layer /= (float)R;
      }
      if (p.b_off[l] >= 0) {  // ||b||_2 of the whole bias vector (torch.linalg.norm(bias, dim=-1) of a 1-D tensor)
        const float* bl = params + p.b_off[l];
        float ss = 0.f;
        for (int r = lane; r < R; r += 32) ss = fmaf(bl[r], bl[r], ss);
#pragma unroll
        for (int m = 16; m >= 1; m >>= 1) ss += __shfl_xor_sync(0xffffffffu, ss, m);
        const float nrm = sqrtf(ss);
        const float sc = nrm > 0.f ? p.lambd / nrm : 0.f;
        for (int r = lane; r < R; r += 32) g_params[p.b_off[l] + r] += sc * bl[r];
        layer += nrm;
      }
      if (lane == 0) total_s += layer;
    }
    __syncthreads();
  }
  if (threadIdx.x == 0) *loss += p.lambd * total_s;
}

// ---------------------------------------------------------------------------------------------- mean over time
// out[b][h] = mean_t ys[b][t][h] (LinearNeuralCDEs.py:389: the LNCDE read-out averages the hidden states).  One CTA per
// sample; thread (g, h) sums the steps t = g (mod G), the G partials are added in ascending g: deterministic.
__global__ void __launch_bounds__(1024) time_mean_fwd_kernel(const float* __restrict__ ys, float* __restrict__ out, int T,
                                                              int H, int Hp) {
  extern __shared__ float part[];  // [G][Hp]
  const int G = blockDim.x / Hp, g = threadIdx.x / Hp, h = threadIdx.x % Hp;
  const float* src = ys + (size_t)blockIdx.x * T * H;
  float acc = 0.f;
  if (h < H && g < G)
    for (int t = g; t < T; t += G) acc += src[(size_t)t * H + h];
  if (g < G) part[g * Hp + h] = acc;
  __syncthreads();
  if (g == 0 && h < H) {
    float s = 0.f;
    for (int k = 0; k < G; ++k) s += part[k * Hp + h];
    out[(size_t)blockIdx.x * H + h] = s / (float)T;
  }
}

// g_ys[b][t][h] = g_mean[b][h] / T
__global__ void __launch_bounds__(256) time_mean_bwd_kernel(const float* __restrict__ g_mean, float* __restrict__ g_ys, int B,
                                                             int T, int H) {
  const long long total = (long long)B * T * H;
  const float inv = 1.f / (float)T;
  for (long long e = blockIdx.x * (long long)blockDim.x + threadIdx.x; e < total; e += (long long)gridDim.x * blockDim.x) {
    const int h = (int)(e % H);
    const long long b = e / ((long long)T * H);
    g_ys[e] = g_mean[b * H + h] * inv;
  }
}

}  // namespace lncde

extern "C" {

int lncde_time_mean_fwd(void* stream, const float* ys, float* out, int B, int T, int H) {
  using namespace lncde;
  if (!ys || !out) return LNCDE_ERR_NULL;
  if (B < 1 || T < 1 || H < 1) return LNCDE_ERR_SHAPE;
  if (H > 1024) return LNCDE_ERR_UNSUPPORTED;
  int Hp = 32;
  while (Hp < H) Hp <<= 1;
  time_mean_fwd_kernel<<<B, 1024, 1024 * sizeof(float), (cudaStream_t)stream>>>(ys, out, T, H, Hp);
  return (int)cudaGetLastError();
}

int lncde_time_mean_bwd(void* stream, const float* g_mean, float* g_ys, int B, int T, int H) {
  using namespace lncde;
  if (!g_mean || !g_ys) return LNCDE_ERR_NULL;
  if (B < 1 || T < 1 || H < 1) return LNCDE_ERR_SHAPE;
  const long long total = (long long)B * T * H;
  const long long want = (total + 255) / 256;
  time_mean_bwd_kernel<<<(int)(want < 148 * 16 ? want : 148 * 16), 256, 0, (cudaStream_t)stream>>>(g_mean, g_ys, B, T, H);
  return (int)cudaGetLastError();
}

int lncde_affine_fwd(void* stream, const float* x, const float* W, const float* b, float* y, int B, int n_in, int n_out) {
  using namespace lncde;
  if (!x || !W || !b || !y) return LNCDE_ERR_NULL;
  if (B < 1 || n_in < 1 || n_out < 1) return LNCDE_ERR_SHAPE;
  const long long total = (long long)B * n_out;
  const int grid = (int)((total + 255) / 256 < 148 * 8 ? (total + 255) / 256 : 148 * 8);
  affine_fwd_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(x, W, b, y, B, n_in, n_out);
  return (int)cudaGetLastError();
}

size_t lncde_affine_bwd_workspace_bytes(int B, int n_in, int n_out) {
  if (B < 1 || n_in < 1 || n_out < 1) return 0;
  const size_t chunks = ((size_t)B + lncde::kChunk - 1) / lncde::kChunk;
  return chunks > 1 ? chunks * (size_t)n_out * (n_in + 1) * sizeof(float) : 0;
}

int lncde_affine_bwd(void* stream, const float* x, const float* g_y, float* g_Wb, int B, int n_in, int n_out,
                     void* workspace, size_t workspace_bytes) {
  using namespace lncde;
  if (!x || !g_y || !g_Wb) return LNCDE_ERR_NULL;
  if (B < 1 || n_in < 1 || n_out < 1) return LNCDE_ERR_SHAPE;
  const int nE = n_out * (n_in + 1), chunks = (B + kChunk - 1) / kChunk;
  if (chunks > 1 && (!workspace || workspace_bytes < lncde_affine_bwd_workspace_bytes(B, n_in, n_out))) return LNCDE_ERR_WORKSPACE;
  cudaStream_t cs = (cudaStream_t)stream;
  float* dst = chunks > 1 ? (float*)workspace : g_Wb;
  affine_bwd_kernel<<<dim3((nE + 255) / 256, chunks), 256, 0, cs>>>(x, g_y, dst, B, n_in, n_out);
  int st = (int)cudaGetLastError();
  if (st != 0 || chunks == 1) return st;
  chunk_reduce_kernel<<<(nE + 255) / 256, 256, 0, cs>>>(dst, g_Wb, nE, nE, chunks, 1.f);
  return (int)cudaGetLastError();
}

size_t lncde_readout_ce_workspace_bytes(int B, int H, int L) {
  if (B < 1 || H < 1 || L < 1) return 0;
  const size_t chunks = ((size_t)B + lncde::kChunk - 1) / lncde::kChunk;
  return ((size_t)B * L + B + (chunks > 1 ? chunks * ((size_t)L * H + L + 1) : 0)) * sizeof(float);
}

int lncde_readout_ce(void* stream, const float* h, long long h_stride, const float* W, const float* b,
                     const float* labels, float* pred, float* g_h, long long g_h_stride, float* g_Wb_loss, int B, int H,
                     int L, void* workspace, size_t workspace_bytes) {
  using namespace lncde;
  if (!h || !W || !b || !labels || !g_h || !g_Wb_loss || !workspace) return LNCDE_ERR_NULL;
  if (B < 1 || H < 1 || L < 1 || h_stride < H || g_h_stride < H) return LNCDE_ERR_SHAPE;
  if (L > kMaxLabels) return LNCDE_ERR_UNSUPPORTED;
  if (workspace_bytes < lncde_readout_ce_workspace_bytes(B, H, L)) return LNCDE_ERR_WORKSPACE;
  float* g_z = (float*)workspace;
  float* loss_b = g_z + (size_t)B * L;
  cudaStream_t cs = (cudaStream_t)stream;
  readout_ce_sample_kernel<<<(B + 3) / 4, 128, 0, cs>>>(h, h_stride, W, b, labels, pred, g_h, g_h_stride, g_z, loss_b, B, H, L);
  int st = (int)cudaGetLastError();
  if (st != 0) return st;
  const int nE = L * H + L + 1, chunks = (B + kChunk - 1) / kChunk;
  float* dst = chunks > 1 ? loss_b + B : g_Wb_loss;
  readout_ce_reduce_kernel<<<dim3((nE + 255) / 256, chunks), 256, 0, cs>>>(h, h_stride, g_z, loss_b, dst, B, H, L,
                                                                            chunks > 1 ? 1.f : 1.f / (float)B);
  st = (int)cudaGetLastError();
  if (st != 0 || chunks == 1) return st;
  // the loss entry (last) is a plain sum of the chunk sums: scale it by 1 / B in a second, single-element pass
  chunk_reduce_kernel<<<(nE - 1 + 255) / 256, 256, 0, cs>>>(dst, g_Wb_loss, nE - 1, nE, chunks, 1.f);
  chunk_reduce_kernel<<<1, 32, 0, cs>>>(dst + (nE - 1), g_Wb_loss + (nE - 1), 1, nE, chunks, 1.f / (float)B);
  return (int)cudaGetLastError();
}

int lncde_lip2(void* stream, const float* params, float* loss, float* g_params, int n_layers, const int32_t* w_off,
               const int32_t* b_off, const int32_t* rows, const int32_t* cols, float lambd) {
  using namespace lncde;
  if (!params || !loss || !g_params || !w_off || !b_off || !rows || !cols) return LNCDE_ERR_NULL;
  if (n_layers < 1 || n_layers > kMaxLipLayers) return LNCDE_ERR_UNSUPPORTED;
  LipParams p;
  p.n_layers = n_layers;
  p.lambd = lambd;
  for (int l = 0; l < n_layers; ++l) {
    if (rows[l] < 1 || cols[l] < 1 || w_off[l] < 0) return LNCDE_ERR_SHAPE;
    p.w_off[l] = w_off[l]; p.b_off[l] = b_off[l]; p.rows[l] = rows[l]; p.cols[l] = cols[l];
  }
  lip2_kernel<<<1, 1024, 0, (cudaStream_t)stream>>>(params, loss, g_params, p);
  return (int)cudaGetLastError();
}

}  // extern "C"
