// K3n – Neural RDE Heun solve (forward) and its exact discrete adjoint (backward).
//
// Replaces NeuralRDE.__call__ + func, models/NeuralCDEs.py:227-266 of the reference (the
// reference's own comparison baseline of the Log-NCDE, same (ts, logsig, x0) input and the same diffrax
// Heun / ConstantStepSize loop):
//
//   field(t, y) = reshape(mlp_linear(vf(y)), (H, Dm)) @ logsig[row(t)][1:] / interval_length
//   vf = eqx.nn.MLP(H -> width, width, depth = n_vf - 1, relu hidden, tanh final)   (NeuralCDEs.py:209-216)
//   mlp_linear: width -> H * Dm, no activation                                       (:217-219), Dm = D - 1
//
// No JVPs: one stage is n_vf small mat-vecs plus ONE bilinear form per state row,
//   G[h] = sum_{d,k} Wm[h][d][k] l_d a_k + sum_d bm[h][d] l_d,
// i.e. a mat-vec of Wm seen as [H][Dm*width] with the rank-1 vector P = l (x) a.  Wm is the bulk of the
// parameters (EigenWorms NRDE config: 344 KB) and is re-read every stage: as many of its rows as fit stay in
// shared memory next to the hidden layers, the rest streams from L2 with coalesced 128-bit loads (one warp per
// 4 rows, all lanes busy because a row is Dm*width long).
// One persistent CTA per sample; stage -> (logsig row, dt / interval) from the same host table as K3
// (models/solver_grid.py).  The adjoint sweep recomputes the two stages of a step from the checkpoint y_n,
// does the transposed products, and defers every weight gradient as (left, right) vector pairs contracted
// afterwards by outer_gemm_kernel on all SMs - deterministic.
#include <cuda_runtime.h>

#include <cstdint>

#include "lncde_b200.h"
#include "logncde_common.cuh"
#include "outer_gemm.cuh"

namespace lncde {

constexpr int kNrdeMaxVf = 4;
constexpr int kNrdeKsplit = 48;        // mlp_linear gradient: tiles x 48 slices
constexpr int kNrdeKsplitSmall = 288;  // single-tile launches (hidden layers, mlp_linear bias)

struct NrdeDesc {
  int H, Dm, Dp, width, n_vf, NW;  // Dp = Dm rounded up to 4, NW = Dm * width
  int in_dim[kNrdeMaxVf], w_off[kNrdeMaxVf], b_off[kNrdeMaxVf];
  int stride[kNrdeMaxVf], sm_off[kNrdeMaxVf], kq_log2[kNrdeMaxVf];
  int wm_off, bm_off, n_params;
  int hidden_floats;  // padded hidden weights in shared memory
  int h_res;          // rows of Wm resident in shared memory (multiple of 4, or H)
  int set_floats;     // one set of stage intermediates
};

static int nrde_ilog2(int v) {
  int l = 0;
  while ((2 << l) <= v) ++l;
  return l;
}

// per-stage intermediates: a0 [H] | a [n_vf][width] | P [NW] | ell [Dp]
__host__ __device__ inline int nrde_set_floats(int H, int width, int n_vf, int NW, int Dp) {
  return ((H + 3) & ~3) + n_vf * width + NW + Dp;
}

static size_t nrde_fixed_floats(const NrdeDesc& d, bool bwd) {
  const size_t Hp = (size_t)((d.H + 3) & ~3);  // H-sized arrays are padded so that everything stays 16-byte aligned
  size_t n = (size_t)d.hidden_floats + (bwd ? 2 : 1) * d.set_floats + 3 * Hp;  // G, ycur/lam, k1/Gbar
  if (bwd) n += Hp + d.NW + d.width + (Hp > (size_t)d.width ? Hp : d.width) + kSolveThreads + d.n_vf * d.width;
  return n + 16;
}

static int make_nrde_desc(NrdeDesc* d, int H, int Dm, int width, int n_vf, bool bwd) {
  if (H < 1 || Dm < 1 || width < 1 || n_vf < 1) return LNCDE_ERR_SHAPE;
  if (n_vf > kNrdeMaxVf || width % 4 != 0 || H > kSolveThreads || width > kSolveThreads) return LNCDE_ERR_UNSUPPORTED;
  d->H = H; d->Dm = Dm; d->Dp = (Dm + 3) & ~3; d->width = width; d->n_vf = n_vf; d->NW = Dm * width;
  int off = 0, sm = 0;
  for (int l = 0; l < n_vf; ++l) {
    d->in_dim[l] = l == 0 ? H : width;
    d->w_off[l] = off;
    off += width * d->in_dim[l];
    d->b_off[l] = off;
    off += width;
    int kq = 0;
    if (width < kSolveThreads) kq = nrde_ilog2(kSolveThreads / width);
    const int cap = nrde_ilog2(d->in_dim[l] > 1 ? d->in_dim[l] / 2 : 1);
    if (kq > cap) kq = cap;
    if (kq > 5) kq = 5;
    d->kq_log2[l] = kq;
    const int KQ = 1 << kq;
    const int pad = (((KQ - d->in_dim[l]) % 32) + 32) % 32;  // row stride = KQ (mod 32): conflict-free row groups
    d->stride[l] = d->in_dim[l] + (KQ == 32 ? 0 : pad);
    d->sm_off[l] = sm;
    sm += width * d->stride[l];
    sm = (sm + 3) & ~3;
  }
  d->wm_off = off;
  off += H * d->NW;
  d->bm_off = off;
  off += H * Dm;
  d->n_params = off;
  d->hidden_floats = sm;
  d->set_floats = nrde_set_floats(H, width, n_vf, d->NW, d->Dp);
  d->h_res = 0;
  const size_t budget = (size_t)216 * 1024 / 4;
  const size_t fixed = nrde_fixed_floats(*d, bwd);
  if (fixed > budget) return LNCDE_ERR_UNSUPPORTED;
  size_t rows = (budget - fixed) / (size_t)d->NW;
  if (rows >= (size_t)H) rows = H;
  else rows &= ~(size_t)3;
  d->h_res = (int)rows;
  return LNCDE_OK;
}

static size_t nrde_smem_bytes(const NrdeDesc& d, bool bwd) {
  return (nrde_fixed_floats(d, bwd) + (size_t)d.h_res * d.NW) * 4;
}

struct NrdeSet {
  float *a0, *a, *P, *ell;
};
__device__ __forceinline__ NrdeSet nrde_set(const NrdeDesc& d, float* base) {
  NrdeSet s;
  s.a0 = base;
  s.a = s.a0 + ((d.H + 3) & ~3);
  s.P = s.a + d.n_vf * d.width;
  s.ell = s.P + d.NW;
  return s;
}

struct NrdeCtx {
  const float* params;
  const float* wb[kNrdeMaxVf];  // padded hidden weights in shared memory
  const float* swm;             // resident rows of Wm
};

__device__ __forceinline__ void nrde_stage_weights(const NrdeDesc& d, const float* __restrict__ params, float* sw,
                                                   float* swm, NrdeCtx& cx) {
  cx.params = params;
  for (int l = 0; l < d.n_vf; ++l) {
    float* dst = sw + d.sm_off[l];
    const float* src = params + d.w_off[l];
    const int n = d.width * d.in_dim[l];
    for (int t = threadIdx.x; t < n; t += kSolveThreads) {
      const int r = t / d.in_dim[l], c = t - r * d.in_dim[l];
      dst[r * d.stride[l] + c] = src[t];
    }
    cx.wb[l] = dst;
  }
  const int n = d.h_res * d.NW;
  const float* src = params + d.wm_off;
  for (int t = threadIdx.x; t < n; t += kSolveThreads) swm[t] = src[t];
  cx.swm = swm;
}

// 4 rows of Wm against P; RES: rows live in shared memory, else global (L2) through the read-only path
template <bool RES>
__device__ __forceinline__ void nrde_rows4(const float* __restrict__ w0, int NW, const float* __restrict__ P, int nrows,
                                           float acc[4]) {
  const int lane = threadIdx.x & 31, nq = NW >> 2;
#pragma unroll 2
  for (int q = lane; q < nq; q += 32) {
    const float4 p4 = *reinterpret_cast<const float4*>(P + 4 * q);
#pragma unroll
    for (int r = 0; r < 4; ++r) {
      if (r < nrows) {
        const float* wr = w0 + (size_t)r * NW + 4 * q;
        const float4 w4 = RES ? *reinterpret_cast<const float4*>(wr) : ldg4(wr);
        acc[r] = fmaf(w4.x, p4.x, acc[r]);
        acc[r] = fmaf(w4.y, p4.y, acc[r]);
        acc[r] = fmaf(w4.z, p4.z, acc[r]);
        acc[r] = fmaf(w4.w, p4.w, acc[r]);
      }
    }
  }
}

// Field at S.a0 with logsig row `ls_row`.  Leaves a_1..a_n, P, ell in S; G[h] (unscaled) if NEED_G.
template <bool NEED_G>
__device__ __noinline__ void nrde_stage_forward(const NrdeDesc& d, const NrdeCtx& cx, const NrdeSet& S,
                                                const float* __restrict__ ls_row, float* G) {
  const int Wd = d.width, H = d.H, NW = d.NW;
  for (int t = threadIdx.x; t < d.Dp; t += kSolveThreads) S.ell[t] = t < d.Dm ? ls_row[1 + t] : 0.f;
  const float* in = S.a0;
  for (int l = 0; l < d.n_vf; ++l) {
    const float* bias = cx.params + d.b_off[l];
    float* out = S.a + l * Wd;
    const bool last = l == d.n_vf - 1;
    matvec_rows<1>(cx.wb[l], d.stride[l], Wd, d.in_dim[l], d.kq_log2[l], in, 0, [&](int row, const float* acc) {
      const float z = acc[0] + bias[row];
      out[row] = last ? tanhf(z) : fmaxf(z, 0.f);
    });
    __syncthreads();
    in = out;
  }
  for (int t = threadIdx.x; t < NW; t += kSolveThreads) {
    const int dd = t / Wd, k = t - dd * Wd;
    S.P[t] = S.ell[dd] * in[k];
  }
  __syncthreads();
  if (!NEED_G) return;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = kSolveThreads >> 5;
  const float* bm = cx.params + d.bm_off;
  for (int r0 = warp * 4; r0 < H; r0 += nw * 4) {
    float acc[4] = {0.f, 0.f, 0.f, 0.f};
    const int nrows = H - r0 < 4 ? H - r0 : 4;
    if (r0 < d.h_res) nrde_rows4<true>(cx.swm + (size_t)r0 * NW, NW, S.P, nrows, acc);
    else nrde_rows4<false>(cx.params + d.wm_off + (size_t)r0 * NW, NW, S.P, nrows, acc);
    for (int dd = lane; dd < d.Dm; dd += 32) {
      const float lv = S.ell[dd];
#pragma unroll
      for (int r = 0; r < 4; ++r)
        if (r < nrows) acc[r] = fmaf(__ldg(bm + (size_t)(r0 + r) * d.Dm + dd), lv, acc[r]);
    }
#pragma unroll
    for (int m = 16; m >= 1; m >>= 1) {
#pragma unroll
      for (int r = 0; r < 4; ++r) acc[r] += __shfl_xor_sync(0xffffffffu, acc[r], m);
    }
    if (lane < nrows) G[r0 + lane] = lane == 0 ? acc[0] : lane == 1 ? acc[1] : lane == 2 ? acc[2] : acc[3];
  }
  __syncthreads();
}

struct NrdeFwdParams {
  NrdeDesc d;
  const float* params;
  const float* logsig;
  const float* y0;
  const int32_t* stage_row;
  const float* stage_scale;
  float* ys;
  int B, K, D, n_steps;
};

__global__ void __launch_bounds__(kSolveThreads, 1) nrde_fwd_kernel(const NrdeFwdParams p) {
  extern __shared__ __align__(16) float smem[];
  const NrdeDesc& d = p.d;
  const int H = d.H;
  float* sw = smem;
  float* swm = sw + d.hidden_floats;
  float* set0 = swm + (size_t)d.h_res * d.NW;
  const int Hp = (H + 3) & ~3;
  float* G = set0 + d.set_floats;
  float* ycur = G + Hp;
  float* k1 = ycur + Hp;
  NrdeCtx cx;
  nrde_stage_weights(d, p.params, sw, swm, cx);
  const NrdeSet S = nrde_set(d, set0);
  const int b = blockIdx.x;
  const float* ls = p.logsig + (size_t)b * p.K * p.D;
  float* ys = p.ys + (size_t)b * (p.n_steps + 1) * H;
  for (int h = threadIdx.x; h < H; h += kSolveThreads) {
    const float v = p.y0[(size_t)b * H + h];
    ycur[h] = v;
    ys[h] = v;
  }
  __syncthreads();
  for (int n = 0; n < p.n_steps; ++n) {
    const int r1 = p.stage_row[2 * n], r2 = p.stage_row[2 * n + 1];
    const float s1 = p.stage_scale[2 * n], s2 = p.stage_scale[2 * n + 1];
    for (int h = threadIdx.x; h < H; h += kSolveThreads) S.a0[h] = ycur[h];
    __syncthreads();
    nrde_stage_forward<true>(d, cx, S, ls + (size_t)r1 * p.D, G);
    for (int h = threadIdx.x; h < H; h += kSolveThreads) {
      const float kk = s1 * G[h];
      k1[h] = kk;
      S.a0[h] = ycur[h] + kk;
    }
    __syncthreads();
    nrde_stage_forward<true>(d, cx, S, ls + (size_t)r2 * p.D, G);
    for (int h = threadIdx.x; h < H; h += kSolveThreads) {
      const float yn = ycur[h] + 0.5f * (k1[h] + s2 * G[h]);
      ycur[h] = yn;
      ys[(size_t)(n + 1) * H + h] = yn;
    }
    __syncthreads();
  }
}

// ------------------------------------------------------------------ workspace of the adjoint
struct NrdeWs {  // float offsets; arrays are [Kt][dim], Kt = B * 2 * n_steps
  size_t G, P, L, Z[kNrdeMaxVf], A[kNrdeMaxVf], bias_part, gemm_part, total;
};

static NrdeWs make_nrde_ws(const NrdeDesc& d, int B, int n_steps) {
  NrdeWs w;
  const size_t Kt = (size_t)B * 2 * n_steps;
  size_t off = 0;
  auto take = [&](size_t n) {
    const size_t o = off;
    off += (n + 3) & ~(size_t)3;
    return o;
  };
  w.G = take(Kt * d.H);
  w.P = take(Kt * d.NW);
  w.L = take(Kt * d.Dp);
  for (int l = 0; l < d.n_vf; ++l) {
    w.Z[l] = take(Kt * d.width);
    w.A[l] = take(Kt * d.in_dim[l]);
  }
  w.bias_part = take((size_t)B * d.n_vf * d.width);
  size_t part = (size_t)d.H * d.NW * kNrdeKsplit;
  const size_t small = (size_t)d.width * (d.H > d.width ? d.H : d.width) * kNrdeKsplitSmall;
  if (small > part) part = small;
  const size_t bm = (size_t)d.H * d.Dm * kNrdeKsplitSmall;
  if (bm > part) part = bm;
  w.gemm_part = take(part);
  w.total = off;
  return w;
}

struct NrdeBwdParams {
  NrdeDesc d;
  NrdeWs ws;
  const float* params;
  const float* logsig;
  const int32_t* stage_row;
  const float* stage_scale;
  const float* ys;
  const float* g_ys;
  float* g_y0;
  float* wsp;
  int B, K, D, n_steps;
};

struct NrdeRev {
  float *Pbar, *zbar, *abar, *scratch, *bacc;
};

// Reverse of one stage.  Gbar [H]: cotangent of the unscaled field.  On exit R.abar[0..H) holds the cotangent
// of the stage input; the weight-gradient operands of this stage are stored at row `kidx` of the workspace.
__device__ __noinline__ void nrde_stage_reverse(const NrdeDesc& d, const NrdeCtx& cx, const NrdeSet& S,
                                                const float* Gbar, const NrdeRev& R, const NrdeWs& ws, float* wsp,
                                                size_t kidx) {
  const int Wd = d.width, H = d.H, NW = d.NW;
  // (i) operands of the mlp_linear gradient: dWm[h][:] += Gbar[h] * P, dbm[h][:] += Gbar[h] * ell
  for (int t = threadIdx.x; t < H; t += kSolveThreads) wsp[ws.G + kidx * H + t] = Gbar[t];
  {
    float4* dst = reinterpret_cast<float4*>(wsp + ws.P + kidx * NW);
    const float4* src = reinterpret_cast<const float4*>(S.P);
    for (int t = threadIdx.x; t < (NW >> 2); t += kSolveThreads) dst[t] = src[t];
  }
  for (int t = threadIdx.x; t < d.Dp; t += kSolveThreads) wsp[ws.L + kidx * d.Dp + t] = S.ell[t];
  // (ii) Pbar = Wm^T Gbar: one thread per 4 columns, rows streamed (coalesced 128-bit loads across the CTA)
  {
    const int nq = NW >> 2;
    const float* wg = cx.params + d.wm_off;
    for (int q = threadIdx.x; q < nq; q += kSolveThreads) {
      float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
      int h = 0;
      for (; h < d.h_res; ++h) {
        const float4 w4 = *reinterpret_cast<const float4*>(cx.swm + (size_t)h * NW + 4 * q);
        const float g = Gbar[h];
        acc.x = fmaf(w4.x, g, acc.x); acc.y = fmaf(w4.y, g, acc.y);
        acc.z = fmaf(w4.z, g, acc.z); acc.w = fmaf(w4.w, g, acc.w);
      }
#pragma unroll 8
      for (; h < H; ++h) {
        const float4 w4 = ldg4(wg + (size_t)h * NW + 4 * q);
        const float g = Gbar[h];
        acc.x = fmaf(w4.x, g, acc.x); acc.y = fmaf(w4.y, g, acc.y);
        acc.z = fmaf(w4.z, g, acc.z); acc.w = fmaf(w4.w, g, acc.w);
      }
      *reinterpret_cast<float4*>(R.Pbar + 4 * q) = acc;
    }
  }
  __syncthreads();
  // (iii) abar[k] = sum_d ell[d] Pbar[d][k]; through tanh
  {
    const float* alast = S.a + (d.n_vf - 1) * Wd;
    for (int k = threadIdx.x; k < Wd; k += kSolveThreads) {
      float s = 0.f;
      for (int dd = 0; dd < d.Dm; ++dd) s = fmaf(S.ell[dd], R.Pbar[dd * Wd + k], s);
      const float a = alast[k];
      R.zbar[k] = s * (1.f - a * a);
    }
  }
  __syncthreads();
  // (iv) hidden layers, last to first
  for (int l = d.n_vf - 1; l >= 0; --l) {
    const float* in = l == 0 ? S.a0 : S.a + (l - 1) * Wd;
    const int in_dim = d.in_dim[l];
    for (int t = threadIdx.x; t < Wd; t += kSolveThreads) {
      const float z = R.zbar[t];
      wsp[ws.Z[l] + kidx * Wd + t] = z;
      R.bacc[l * Wd + t] += z;
    }
    for (int t = threadIdx.x; t < in_dim; t += kSolveThreads) wsp[ws.A[l] + kidx * in_dim + t] = in[t];
    matvec_cols<1, false>(cx.wb[l], d.stride[l], Wd, in_dim, R.zbar, 0, R.abar, 0, R.scratch);
    if (l > 0) {
      for (int c = threadIdx.x; c < Wd; c += kSolveThreads) R.zbar[c] = in[c] > 0.f ? R.abar[c] : 0.f;
      __syncthreads();
    }
  }
}

__global__ void __launch_bounds__(kSolveThreads, 1) nrde_bwd_kernel(const NrdeBwdParams p) {
  extern __shared__ __align__(16) float smem[];
  const NrdeDesc& d = p.d;
  const int H = d.H, Wd = d.width;
  float* sw = smem;
  float* swm = sw + d.hidden_floats;
  float* set0 = swm + (size_t)d.h_res * d.NW;
  float* set1 = set0 + d.set_floats;
  const int Hp = (H + 3) & ~3;
  float* G = set1 + d.set_floats;
  float* lam = G + Hp;
  float* Gbar = lam + Hp;
  float* ybB = Gbar + Hp;
  NrdeRev R;
  R.Pbar = ybB + Hp;
  R.zbar = R.Pbar + d.NW;
  R.abar = R.zbar + Wd;
  R.scratch = R.abar + (Hp > Wd ? Hp : Wd);
  R.bacc = R.scratch + kSolveThreads;
  NrdeCtx cx;
  nrde_stage_weights(d, p.params, sw, swm, cx);
  const NrdeSet SA = nrde_set(d, set0), SB = nrde_set(d, set1);
  const int b = blockIdx.x;
  const float* ls = p.logsig + (size_t)b * p.K * p.D;
  const float* ys = p.ys + (size_t)b * (p.n_steps + 1) * H;
  const float* gy = p.g_ys + (size_t)b * (p.n_steps + 1) * H;
  for (int h = threadIdx.x; h < H; h += kSolveThreads) lam[h] = gy[(size_t)p.n_steps * H + h];
  for (int t = threadIdx.x; t < d.n_vf * Wd; t += kSolveThreads) R.bacc[t] = 0.f;
  __syncthreads();
  for (int n = p.n_steps - 1; n >= 0; --n) {
    const int r1 = p.stage_row[2 * n], r2 = p.stage_row[2 * n + 1];
    const float s1 = p.stage_scale[2 * n], s2 = p.stage_scale[2 * n + 1];
    const size_t kidx = ((size_t)b * p.n_steps + n) * 2;
    // recompute the two stages of step n from the checkpoint y_n
    for (int h = threadIdx.x; h < H; h += kSolveThreads) SA.a0[h] = ys[(size_t)n * H + h];
    __syncthreads();
    nrde_stage_forward<true>(d, cx, SA, ls + (size_t)r1 * p.D, G);
    for (int h = threadIdx.x; h < H; h += kSolveThreads) {
      SB.a0[h] = SA.a0[h] + s1 * G[h];
      Gbar[h] = 0.5f * s2 * lam[h];
    }
    __syncthreads();
    nrde_stage_forward<false>(d, cx, SB, ls + (size_t)r2 * p.D, G);
    // stage 2: y_{n+1} = y_n + (k1 + k2) / 2, k2 = s2 G2(y_n + k1)
    nrde_stage_reverse(d, cx, SB, Gbar, R, p.ws, p.wsp, kidx + 1);
    for (int h = threadIdx.x; h < H; h += kSolveThreads) {
      const float yb = R.abar[h];  // cotangent of y_n + k1
      ybB[h] = yb;
      Gbar[h] = s1 * (0.5f * lam[h] + yb);  // k1 = s1 G1(y_n) feeds y_{n+1} (weight 1/2) and the second stage
    }
    __syncthreads();
    nrde_stage_reverse(d, cx, SA, Gbar, R, p.ws, p.wsp, kidx);
    for (int h = threadIdx.x; h < H; h += kSolveThreads)
      lam[h] = lam[h] + ybB[h] + R.abar[h] + gy[(size_t)n * H + h];
    __syncthreads();
  }
  for (int h = threadIdx.x; h < H; h += kSolveThreads) p.g_y0[(size_t)b * H + h] = lam[h];
  for (int t = threadIdx.x; t < d.n_vf * Wd; t += kSolveThreads)
    p.wsp[p.ws.bias_part + (size_t)b * d.n_vf * Wd + t] = R.bacc[t];
}

// out[e] = sum_{k < count} part[k * stride + e]
__global__ void nrde_reduce_kernel(const float* __restrict__ part, float* __restrict__ out, int n, int count,
                                   size_t stride) {
  for (int e = blockIdx.x * blockDim.x + threadIdx.x; e < n; e += gridDim.x * blockDim.x) {
    float s = 0.f;
    for (int k = 0; k < count; ++k) s += part[(size_t)k * stride + e];
    out[e] = s;
  }
}

static int nrde_gemm(cudaStream_t cs, const float* left, const float* right, int M, int N, long long ldl, long long ldr,
                     long long K, int ksplit, float* part, float* out) {
  GemmParams g;
  g.ksplit = ksplit;
  g.part = part;
  g.n_tasks = 1;
  GemmTask& t = g.task[0];
  t.M = M; t.N = N; t.out_off = 0; t.part_off = 0; t.ksplit = ksplit;
  t.seg[0] = {left, right, ldl, ldr, K};
  t.seg[1] = {nullptr, nullptr, 0, 0, 0};
  launch_gemm(g, 0, 1, cs);
  int st = (int)cudaGetLastError();
  if (st != 0) return st;
  const int n = M * N;
  int grid = (n + 255) / 256;
  if (grid > 148 * 4) grid = 148 * 4;
  nrde_reduce_kernel<<<grid, 256, 0, cs>>>(part, out, n, ksplit, (size_t)n);
  return (int)cudaGetLastError();
}

static int check_nrde_args(int B, int K, int D, int H, int width, int n_vf, int n_steps) {
  if (B < 1 || K < 1 || D < 2 || H < 1 || width < 1 || n_vf < 1 || n_steps < 1) return LNCDE_ERR_SHAPE;
  return LNCDE_OK;
}

}  // namespace lncde

extern "C" {

int lncde_nrde_param_count(int H, int Dm, int width, int n_vf) {
  lncde::NrdeDesc d;
  if (lncde::make_nrde_desc(&d, H, Dm, width, n_vf, false) != LNCDE_OK) return -1;
  return d.n_params;
}

int lncde_nrde_fwd(void* stream, const float* params, const float* logsig, const float* y0, const int32_t* stage_row,
                   const float* stage_scale, float* ys, int B, int K, int D, int H, int width, int n_vf, int n_steps) {
  using namespace lncde;
  if (!params || !logsig || !y0 || !stage_row || !stage_scale || !ys) return LNCDE_ERR_NULL;
  int st = check_nrde_args(B, K, D, H, width, n_vf, n_steps);
  if (st != LNCDE_OK) return st;
  if (((uintptr_t)params & 15) != 0) return LNCDE_ERR_ALIGN;
  NrdeFwdParams p;
  st = make_nrde_desc(&p.d, H, D - 1, width, n_vf, false);
  if (st != LNCDE_OK) return st;
  p.params = params; p.logsig = logsig; p.y0 = y0; p.stage_row = stage_row; p.stage_scale = stage_scale; p.ys = ys;
  p.B = B; p.K = K; p.D = D; p.n_steps = n_steps;
  const size_t smem = nrde_smem_bytes(p.d, false);
  cudaFuncSetAttribute(nrde_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  nrde_fwd_kernel<<<B, kSolveThreads, smem, (cudaStream_t)stream>>>(p);
  return (int)cudaGetLastError();
}

size_t lncde_nrde_bwd_workspace_bytes(int B, int H, int Dm, int width, int n_vf, int n_steps) {
  using namespace lncde;
  NrdeDesc d;
  if (B < 1 || n_steps < 1 || make_nrde_desc(&d, H, Dm, width, n_vf, true) != LNCDE_OK) return 0;
  return make_nrde_ws(d, B, n_steps).total * 4;
}

int lncde_nrde_bwd(void* stream, const float* params, const float* logsig, const int32_t* stage_row,
                   const float* stage_scale, const float* ys, const float* g_ys, float* g_params, float* g_y0, int B,
                   int K, int D, int H, int width, int n_vf, int n_steps, void* workspace, size_t workspace_bytes) {
  using namespace lncde;
  if (!params || !logsig || !stage_row || !stage_scale || !ys || !g_ys || !g_params || !g_y0 || !workspace)
    return LNCDE_ERR_NULL;
  int st = check_nrde_args(B, K, D, H, width, n_vf, n_steps);
  if (st != LNCDE_OK) return st;
  if (((uintptr_t)params & 15) != 0 || ((uintptr_t)workspace & 15) != 0) return LNCDE_ERR_ALIGN;
  NrdeBwdParams p;
  st = make_nrde_desc(&p.d, H, D - 1, width, n_vf, true);
  if (st != LNCDE_OK) return st;
  const NrdeDesc& d = p.d;
  p.ws = make_nrde_ws(d, B, n_steps);
  if (workspace_bytes < p.ws.total * 4) return LNCDE_ERR_WORKSPACE;
  cudaStream_t cs = (cudaStream_t)stream;
  float* wsp = (float*)workspace;
  p.params = params; p.logsig = logsig; p.stage_row = stage_row; p.stage_scale = stage_scale; p.ys = ys;
  p.g_ys = g_ys; p.g_y0 = g_y0; p.wsp = wsp; p.B = B; p.K = K; p.D = D; p.n_steps = n_steps;
  const size_t smem = nrde_smem_bytes(d, true);
  cudaFuncSetAttribute(nrde_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  nrde_bwd_kernel<<<B, kSolveThreads, smem, cs>>>(p);
  st = (int)cudaGetLastError();
  if (st != 0) return st;
  // deferred weight gradients: sums of outer products over all (sample, step, stage) rows
  const long long Kt = (long long)B * 2 * n_steps;
  float* part = wsp + p.ws.gemm_part;
  for (int l = 0; l < d.n_vf; ++l) {
    st = nrde_gemm(cs, wsp + p.ws.Z[l], wsp + p.ws.A[l], d.width, d.in_dim[l], d.width, d.in_dim[l], Kt,
                   kNrdeKsplitSmall, part, g_params + d.w_off[l]);
    if (st != 0) return st;
  }
  st = nrde_gemm(cs, wsp + p.ws.G, wsp + p.ws.P, d.H, d.NW, d.H, d.NW, Kt, kNrdeKsplit, part, g_params + d.wm_off);
  if (st != 0) return st;
  st = nrde_gemm(cs, wsp + p.ws.G, wsp + p.ws.L, d.H, d.Dm, d.H, d.Dp, Kt, kNrdeKsplitSmall, part, g_params + d.bm_off);
  if (st != 0) return st;
  // hidden biases: per-sample partial sums, fixed order over the batch
  for (int l = 0; l < d.n_vf; ++l) {
    nrde_reduce_kernel<<<1, 256, 0, cs>>>(wsp + p.ws.bias_part + (size_t)l * d.width, g_params + d.b_off[l], d.width, B,
                                          (size_t)d.n_vf * d.width);
    st = (int)cudaGetLastError();
    if (st != 0) return st;
  }
  return LNCDE_OK;
}

}  // extern "C"
