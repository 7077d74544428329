// Specialised Log-NCDE solve kernels for the EigenWorms model shape
// (hidden H = 128, vector-field width 32, 2 hidden layers; C <= 8 channels).
//
// Same algorithm and workspace layout as the generic kernels in logncde.cu, but laid out
// so that the instruction count per Heun stage is close to the FMA floor:
//   * the last MLP layer (C*128 x 32, 80 % of the weights) lives in REGISTERS: each of the
//     C*64 "row threads" owns an 8 x 8 tile (8 consecutive output rows x 8 inputs).  The
//     same tile serves W a (forward: reduce over the 4 column groups = lane bits 0-1) and
//     W^T v (adjoint: reduce over the 8 row groups = lane bits 2-4) through shuffle
//     "transpose-reductions" that halve the live values at every step.
//   * the small layers (32 x 128, 32 x 32) sit in shared memory in the exact order the
//     threads read them (one conflict-free 128-bit load per 4 weights), each thread
//     applying its slice to the primal or to all C tangents.
//   * all loops are compile-time; one __syncthreads per dependent phase.
#pragma once
#include <cuda_runtime.h>

#include <cstdint>

namespace lncde {
namespace fast {

constexpr int H = 128, W = 32, NT = 512;

__device__ __forceinline__ float sigm(float z) { return 1.f / (1.f + expf(-z)); }

// One halving step of a transpose-reduction across lanes lane^mask: v[0..HALF) keeps the
// lower half for lanes with upper == false and the upper half otherwise.
template <int HALF>
__device__ __forceinline__ void xstep(float* v, bool upper, int mask) {
#pragma unroll
  for (int k = 0; k < HALF; ++k) {
    const float send = upper ? v[k] : v[k + HALF];
    const float keep = upper ? v[k + HALF] : v[k];
    v[k] = keep + __shfl_xor_sync(0xffffffffu, send, mask);
  }
}

__device__ __forceinline__ float dot8(const float4& a0, const float4& a1, const float4& b0, const float4& b1) {
  float s = a0.x * b0.x;
  s = fmaf(a0.y, b0.y, s);
  s = fmaf(a0.z, b0.z, s);
  s = fmaf(a0.w, b0.w, s);
  float t = a1.x * b1.x;
  t = fmaf(a1.y, b1.y, t);
  t = fmaf(a1.z, b1.z, t);
  t = fmaf(a1.w, b1.w, t);
  return s + t;
}

// shared-memory plan (float offsets)
template <int C>
struct Smem {
  static constexpr int CH = C * H, CW = C * W;
  // weights
  static constexpr int W1 = 0;               // [32][128] row-major
  static constexpr int W2 = W1 + W * H;      // [32][32]
  static constexpr int W1T = W2 + W * W;     // [128][32]   (backward only)
  static constexpr int W2T = W1T + H * W;    // [32][32]    (backward only)
  static constexpr int WEND_FWD = W1T;
  static constexpr int WEND_BWD = W2T + W * W;
  // stage buffer (relative)
  static constexpr int yin = 0, z1 = yin + H, a1 = z1 + W, s1 = a1 + W, z2 = s1 + W, a2 = z2 + W, s2 = a2 + W;
  static constexpr int o = s2 + W, w = o + CH, u1 = w + CH, p1 = u1 + CW, u2 = p1 + CW, p2 = u2 + CW;
  static constexpr int STAGE = p2 + CW;
  // common (relative to the common base)
  static constexpr int lam = 0;              // [8][8]
  static constexpr int lam1 = lam + 64;      // [8]
  static constexpr int dbuf = lam1 + 8;      // [CH]  last-layer tangent outputs / row staging
  static constexpr int ycur = dbuf + CH, k1 = ycur + H;
  static constexpr int COMMON_FWD = k1 + H;
  // backward extras
  static constexpr int g = COMMON_FWD, ybar = g + H, ab = ybar + H, yout = ab + H;
  static constexpr int red = yout + H;       // [16][32]
  static constexpr int wbar = red + 16 * W;  // [C][128]
  static constexpr int ub2 = wbar + CH, ub1 = ub2 + CW, sp2 = ub1 + CW, sp1 = sp2 + CW;  // [C][32] each
  static constexpr int sb2 = sp1 + CW, sb1 = sb2 + W, zb2 = sb1 + W, zb1 = zb2 + W;
  static constexpr int bacc = zb1 + W;       // [64] bias accumulators of the two hidden layers
  static constexpr int COMMON_BWD = bacc + 2 * W;
};

struct Regs {
  float Wt[8][8];  // last-layer tile: rows r0..r0+7, inputs cg*8..cg*8+7
  float b3[2];     // bias of the two rows this lane finishes
  float b1, b2;    // hidden-layer biases of output o = tid / 16
};

// rows finished by this lane after the forward transpose-reduce
__device__ __forceinline__ int lane_row0(int tid) {
  const int lane = tid & 31, warp = tid >> 5;
  const int rg = lane >> 2;
  return (warp * 8 + rg) * 8 + (lane & 1) * 4 + ((lane >> 1) & 1) * 2;
}

template <int C>
__device__ __forceinline__ void load_regs(Regs& R, const float* __restrict__ params, const MlpDesc& d) {
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int o = tid >> 4;
  R.b1 = params[d.b_off[0] + o];
  R.b2 = params[d.b_off[1] + o];
  if (tid < C * 64) {
    const int r0 = (warp * 8 + (lane >> 2)) * 8, cg = lane & 3;
#pragma unroll
    for (int e = 0; e < 8; ++e)
#pragma unroll
      for (int k = 0; k < 8; ++k) R.Wt[e][k] = params[d.w_off[2] + (size_t)(r0 + e) * W + cg * 8 + k];
    const int row0 = lane_row0(tid);
    R.b3[0] = params[d.b_off[2] + row0];
    R.b3[1] = params[d.b_off[2] + row0 + 1];
  }
}

// y[r0+e] = sum_k Wt[e][k] * x[cg*8+k], reduced over the 4 column groups; the lane ends
// with the two complete sums of rows lane_row0, lane_row0+1.
__device__ __forceinline__ void tile_rows(const Regs& R, const float* __restrict__ x8, int lane, float& r_a,
                                          float& r_b) {
  const float4 x0 = *reinterpret_cast<const float4*>(x8);
  const float4 x1 = *reinterpret_cast<const float4*>(x8 + 4);
  float v[8];
#pragma unroll
  for (int e = 0; e < 8; ++e) {
    float s = R.Wt[e][0] * x0.x;
    s = fmaf(R.Wt[e][1], x0.y, s);
    s = fmaf(R.Wt[e][2], x0.z, s);
    s = fmaf(R.Wt[e][3], x0.w, s);
    s = fmaf(R.Wt[e][4], x1.x, s);
    s = fmaf(R.Wt[e][5], x1.y, s);
    s = fmaf(R.Wt[e][6], x1.z, s);
    s = fmaf(R.Wt[e][7], x1.w, s);
    v[e] = s;
  }
  xstep<4>(v, lane & 1, 1);
  xstep<2>(v, lane & 2, 2);
  r_a = v[0];
  r_b = v[1];
}

// col[cg*8+k] = sum_e Wt[e][k] * x[r0+e], reduced over the 8 row groups of the warp; the lane
// ends with the warp-partial of column cg*8 + (bit2*4 + bit3*2 + bit4).
__device__ __forceinline__ float tile_cols(const Regs& R, const float* __restrict__ x8, int lane) {
  const float4 x0 = *reinterpret_cast<const float4*>(x8);
  const float4 x1 = *reinterpret_cast<const float4*>(x8 + 4);
  float v[8];
#pragma unroll
  for (int k = 0; k < 8; ++k) {
    float s = R.Wt[0][k] * x0.x;
    s = fmaf(R.Wt[1][k], x0.y, s);
    s = fmaf(R.Wt[2][k], x0.z, s);
    s = fmaf(R.Wt[3][k], x0.w, s);
    s = fmaf(R.Wt[4][k], x1.x, s);
    s = fmaf(R.Wt[5][k], x1.y, s);
    s = fmaf(R.Wt[6][k], x1.z, s);
    s = fmaf(R.Wt[7][k], x1.w, s);
    v[k] = s;
  }
  xstep<4>(v, lane & 4, 4);
  xstep<2>(v, lane & 8, 8);
  xstep<1>(v, lane & 16, 16);
  return v[0];
}
__device__ __forceinline__ int lane_col(int lane) {
  return (lane & 3) * 8 + ((lane >> 2) & 1) * 4 + ((lane >> 3) & 1) * 2 + ((lane >> 4) & 1);
}

// C vectors reduced over 16 lanes (bits 0-3): on exit lanes with even q hold the total of
// vector q >> 1 in v[0].
__device__ __forceinline__ void reduce16_vec8(float* v, int q) {
  xstep<4>(v, q & 8, 8);
  xstep<2>(v, q & 4, 4);
  xstep<1>(v, q & 2, 2);
  v[0] += __shfl_xor_sync(0xffffffffu, v[0], 1);
}
__device__ __forceinline__ int vec_of_q(int q) { return ((q >> 3) & 1) * 4 + ((q >> 2) & 1) * 2 + ((q >> 1) & 1); }

template <int C>
__device__ __forceinline__ void load_lambda(float* cm, const float* __restrict__ row, bool d2) {
  using S = Smem<C>;
  const int tid = threadIdx.x;
  if (tid < 64) {
    const int i = tid >> 3, j = tid & 7;
    float v = 0.f;
    if (d2 && i < C && j < C && i != j) {
      const int a = i < j ? i : j, b = i < j ? j : i;
      v = row[1 + C + a * C - (a * (a + 1)) / 2 + (b - a - 1)];
      if (i > j) v = -v;
    }
    cm[S::lam + tid] = v;
  } else if (tid < 72) {
    const int i = tid - 64;
    cm[S::lam1 + i] = i < C ? row[1 + i] : 0.f;
  }
}

// Forward evaluation of the field at sb[yin].  Leaves every intermediate in `sb`, the
// last-layer outputs of this lane in o_r / u3_r and, after the trailing barrier, the
// per-row tangent contributions in cm[dbuf].  The caller finishes
//   G[h] = sum_i lam1[i] o[i][h] + sum_j dbuf[j][h].
template <int C, bool D2>
__device__ __forceinline__ void stage_forward(const float* __restrict__ sw, float* sb, float* cm, const Regs& R,
                                              const float* __restrict__ ls_row, float (&o_r)[2], float (&u3_r)[2]) {
  using S = Smem<C>;
  const int tid = threadIdx.x, lane = tid & 31;
  const int o = tid >> 4, q = tid & 15;
  load_lambda<C>(cm, ls_row, D2);
  // P1: z1 = W1 y + b1
  const float4 w1a = *reinterpret_cast<const float4*>(sw + S::W1 + tid * 8);
  const float4 w1b = *reinterpret_cast<const float4*>(sw + S::W1 + tid * 8 + 4);
  {
    const float4 ya = *reinterpret_cast<const float4*>(sb + S::yin + q * 8);
    const float4 yb = *reinterpret_cast<const float4*>(sb + S::yin + q * 8 + 4);
    float acc = dot8(w1a, w1b, ya, yb);
    acc += __shfl_xor_sync(0xffffffffu, acc, 8);
    acc += __shfl_xor_sync(0xffffffffu, acc, 4);
    acc += __shfl_xor_sync(0xffffffffu, acc, 2);
    acc += __shfl_xor_sync(0xffffffffu, acc, 1);
    if (q == 0) {
      const float z = acc + R.b1, sg = sigm(z);
      sb[S::z1 + o] = z;
      sb[S::a1 + o] = z * sg;
      sb[S::s1 + o] = sg * (1.f + z * (1.f - sg));
    }
  }
  __syncthreads();
  // P2: z2 = W2 a1 + b2
  const float2 w2 = *reinterpret_cast<const float2*>(sw + S::W2 + tid * 2);
  {
    const float2 a = *reinterpret_cast<const float2*>(sb + S::a1 + q * 2);
    float acc = fmaf(w2.y, a.y, w2.x * a.x);
    acc += __shfl_xor_sync(0xffffffffu, acc, 8);
    acc += __shfl_xor_sync(0xffffffffu, acc, 4);
    acc += __shfl_xor_sync(0xffffffffu, acc, 2);
    acc += __shfl_xor_sync(0xffffffffu, acc, 1);
    if (q == 0) {
      const float z = acc + R.b2, sg = sigm(z);
      sb[S::z2 + o] = z;
      sb[S::a2 + o] = z * sg;
      sb[S::s2 + o] = sg * (1.f + z * (1.f - sg));
    }
  }
  __syncthreads();
  // P3: o = tanh(W3 a2 + b3)
  const int row0 = lane_row0(tid);
  if (tid < C * 64) {
    float ra, rb;
    tile_rows(R, sb + S::a2 + (lane & 3) * 8, lane, ra, rb);
    o_r[0] = tanhf(ra + R.b3[0]);
    o_r[1] = tanhf(rb + R.b3[1]);
    *reinterpret_cast<float2*>(sb + S::o + row0) = make_float2(o_r[0], o_r[1]);
  }
  __syncthreads();
  if (!D2) return;
  // P4: w_j = sum_i Lam_ij o_i
  if (tid < C * 64) {
    const int j = tid >> 6, hh = tid & 63;
    float wa = 0.f, wb = 0.f;
#pragma unroll
    for (int i = 0; i < C; ++i) {
      const float l = cm[S::lam + i * 8 + j];
      wa = fmaf(l, sb[S::o + i * H + hh], wa);
      wb = fmaf(l, sb[S::o + i * H + hh + 64], wb);
    }
    sb[S::w + j * H + hh] = wa;
    sb[S::w + j * H + hh + 64] = wb;
  }
  __syncthreads();
  // P5: u1_j = W1 w_j, p1_j = s1 * u1_j
  {
    float v[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      if (j < C) {
        const float4 xa = *reinterpret_cast<const float4*>(sb + S::w + j * H + q * 8);
        const float4 xb = *reinterpret_cast<const float4*>(sb + S::w + j * H + q * 8 + 4);
        v[j] = dot8(w1a, w1b, xa, xb);
      } else {
        v[j] = 0.f;
      }
    }
    reduce16_vec8(v, q);
    const int jv = vec_of_q(q);
    if (!(q & 1) && jv < C) {
      sb[S::u1 + jv * W + o] = v[0];
      sb[S::p1 + jv * W + o] = sb[S::s1 + o] * v[0];
    }
  }
  __syncthreads();
  // P6: u2_j = W2 p1_j, p2_j = s2 * u2_j
  {
    float v[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      if (j < C) {
        const float2 x = *reinterpret_cast<const float2*>(sb + S::p1 + j * W + q * 2);
        v[j] = fmaf(w2.y, x.y, w2.x * x.x);
      } else {
        v[j] = 0.f;
      }
    }
    reduce16_vec8(v, q);
    const int jv = vec_of_q(q);
    if (!(q & 1) && jv < C) {
      sb[S::u2 + jv * W + o] = v[0];
      sb[S::p2 + jv * W + o] = sb[S::s2 + o] * v[0];
    }
  }
  __syncthreads();
  // P7: u3 = W3[block j] p2_j ; d = (1 - o^2) u3
  if (tid < C * 64) {
    const int j = tid >> 6;
    tile_rows(R, sb + S::p2 + j * W + (lane & 3) * 8, lane, u3_r[0], u3_r[1]);
    *reinterpret_cast<float2*>(cm + S::dbuf + row0) =
        make_float2((1.f - o_r[0] * o_r[0]) * u3_r[0], (1.f - o_r[1] * o_r[1]) * u3_r[1]);
  }
  __syncthreads();
}

template <int C, bool D2>
__device__ __forceinline__ float field_value(const float* sb, const float* cm, int h) {
  using S = Smem<C>;
  float gsum = 0.f;
#pragma unroll
  for (int i = 0; i < C; ++i) gsum = fmaf(cm[S::lam1 + i], sb[S::o + i * H + h], gsum);
  if (D2) {
#pragma unroll
    for (int j = 0; j < C; ++j) gsum += cm[S::dbuf + j * H + h];
  }
  return gsum;
}

template <int C>
__device__ __forceinline__ void stage_weights_fast(const MlpDesc& d, const float* __restrict__ params, float* sw,
                                                   bool bwd) {
  using S = Smem<C>;
  for (int t = threadIdx.x; t < W * H; t += NT) {
    const float v = params[d.w_off[0] + t];
    sw[S::W1 + t] = v;
    if (bwd) {
      const int c = t / H, h = t - c * H;
      sw[S::W1T + h * W + c] = v;
    }
  }
  for (int t = threadIdx.x; t < W * W; t += NT) {
    const float v = params[d.w_off[1] + t];
    sw[S::W2 + t] = v;
    if (bwd) {
      const int c = t / W, c2 = t - c * W;
      sw[S::W2T + c2 * W + c] = v;
    }
  }
}

// ------------------------------------------------------------------------------ forward
template <int C, bool D2>
__global__ void __launch_bounds__(NT, 1) logncde_fwd_fast_kernel(const FwdParams p) {
  using S = Smem<C>;
  extern __shared__ __align__(16) float smem[];
  float* sw = smem;
  float* sb = smem + S::WEND_FWD;
  float* cm = sb + S::STAGE;
  const int tid = threadIdx.x;
  Regs R;
  load_regs<C>(R, p.params, p.d);
  stage_weights_fast<C>(p.d, p.params, sw, false);
  const int b = blockIdx.x;
  const float* ls = p.logsig + (size_t)b * p.K * p.D;
  float* ys = p.ys + (size_t)b * (p.n_steps + 1) * H;
  if (tid < H) {
    const float v = p.y0[(size_t)b * H + tid];
    cm[S::ycur + tid] = v;
    sb[S::yin + tid] = v;
    ys[tid] = v;
  }
  __syncthreads();
  float o_r[2], u3_r[2];
  for (int n = 0; n < p.n_steps; ++n) {
    const int r1 = p.stage_row[2 * n], r2 = p.stage_row[2 * n + 1];
    const float s1 = p.stage_scale[2 * n], s2 = p.stage_scale[2 * n + 1];
    stage_forward<C, D2>(sw, sb, cm, R, ls + (size_t)r1 * p.D, o_r, u3_r);
    if (tid < H) {
      const float kk = s1 * field_value<C, D2>(sb, cm, tid);
      cm[S::k1 + tid] = kk;
      sb[S::yin + tid] = cm[S::ycur + tid] + kk;
    }
    __syncthreads();
    stage_forward<C, D2>(sw, sb, cm, R, ls + (size_t)r2 * p.D, o_r, u3_r);
    if (tid < H) {
      const float yn = cm[S::ycur + tid] + 0.5f * (cm[S::k1 + tid] + s2 * field_value<C, D2>(sb, cm, tid));
      cm[S::ycur + tid] = yn;
      sb[S::yin + tid] = yn;
      ys[(size_t)(n + 1) * H + tid] = yn;
    }
    __syncthreads();
  }
}

// ------------------------------------------------------------------------------ backward
// Reverse sweep of one stage.  cm[g] holds dLoss/dG times the stage scale; on exit (after the
// trailing barrier) cm[yout] holds dLoss/d(stage input).
template <int C, bool D2>
__device__ __forceinline__ void stage_reverse(const float* __restrict__ sw, float* sb, float* cm, const Regs& R,
                                              const float (&o_r)[2], const float (&u3_r)[2], float (&b3acc)[2],
                                              const WsLayout& ws, float* __restrict__ wsp, size_t slot) {
  using S = Smem<C>;
  constexpr int CH = C * H, CW = C * W;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int o = tid >> 4, q = tid & 15;
  const int row0 = lane_row0(tid);
  // right-hand vectors of the weight-gradient outer products
  if (tid < H) wsp[ws.Y + slot * H + tid] = sb[S::yin + tid];
  if (tid < W) {
    wsp[ws.A[0] + slot * W + tid] = sb[S::a1 + tid];
    wsp[ws.A[1] + slot * W + tid] = sb[S::a2 + tid];
  }
  if (D2) {
    for (int t = tid; t < CH; t += NT) wsp[ws.Wt + slot * CH + t] = sb[S::w + t];
    if (tid < CW) {
      wsp[ws.P[0] + slot * CW + tid] = sb[S::p1 + tid];
      wsp[ws.P[1] + slot * CW + tid] = sb[S::p2 + tid];
    }
  }
  // Q1: cotangents at the field output; p2-bar partials
  float ob[2] = {0.f, 0.f}, tp[2] = {0.f, 0.f};
  if (tid < C * 64) {
    const int j = tid >> 6, h0 = row0 & (H - 1);
    const float2 gv = *reinterpret_cast<const float2*>(cm + S::g + h0);
    const float l1 = cm[S::lam1 + j];
    tp[0] = 1.f - o_r[0] * o_r[0];
    tp[1] = 1.f - o_r[1] * o_r[1];
    ob[0] = l1 * gv.x;
    ob[1] = l1 * gv.y;
    if (D2) {
      const float2 ub = make_float2(tp[0] * gv.x, tp[1] * gv.y);
      ob[0] = fmaf(-2.f * o_r[0], u3_r[0] * gv.x, ob[0]);
      ob[1] = fmaf(-2.f * o_r[1], u3_r[1] * gv.y, ob[1]);
      *reinterpret_cast<float2*>(wsp + ws.UL + slot * CH + row0) = ub;
      *reinterpret_cast<float2*>(cm + S::dbuf + row0) = ub;
      __syncwarp();
      const float part = tile_cols(R, cm + S::dbuf + (warp * 8 + (lane >> 2)) * 8, lane);
      cm[S::red + warp * W + lane_col(lane)] = part;
    }
  }
  if (D2) {
    __syncthreads();
    // Q2: u2-bar = s2 * p2-bar ; s2-bar partials
    if (tid < CW) {
      const int j = tid >> 5, c = tid & 31;
      const float pb = cm[S::red + (2 * j) * W + c] + cm[S::red + (2 * j + 1) * W + c];
      const float ub = sb[S::s2 + c] * pb;
      cm[S::ub2 + tid] = ub;
      cm[S::sp2 + tid] = sb[S::u2 + tid] * pb;
      wsp[ws.U[1] + slot * CW + tid] = ub;
    }
    __syncthreads();
    // Q3: p1-bar_j = W2^T u2-bar_j ; u1-bar = s1 * p1-bar ; s1-bar partials ; s2-bar
    {
      const float2 w2t = *reinterpret_cast<const float2*>(sw + S::W2T + tid * 2);
      float v[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        if (j < C) {
          const float2 x = *reinterpret_cast<const float2*>(cm + S::ub2 + j * W + q * 2);
          v[j] = fmaf(w2t.y, x.y, w2t.x * x.x);
        } else {
          v[j] = 0.f;
        }
      }
      reduce16_vec8(v, q);
      const int jv = vec_of_q(q);
      if (!(q & 1) && jv < C) {
        const float ub = sb[S::s1 + o] * v[0];
        cm[S::ub1 + jv * W + o] = ub;
        cm[S::sp1 + jv * W + o] = sb[S::u1 + jv * W + o] * v[0];
        wsp[ws.U[0] + slot * CW + jv * W + o] = ub;
      }
      if (tid < W) {
        float s = 0.f;
#pragma unroll
        for (int j = 0; j < C; ++j) s += cm[S::sp2 + j * W + tid];
        cm[S::sb2 + tid] = s;
      }
    }
    __syncthreads();
    // Q4: w-bar_j = W1^T u1-bar_j ; s1-bar
    {
      const int h = tid >> 2, q4 = tid & 3;
      const float4 wa = *reinterpret_cast<const float4*>(sw + S::W1T + tid * 8);
      const float4 wb = *reinterpret_cast<const float4*>(sw + S::W1T + tid * 8 + 4);
      float v[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        if (j < C) {
          const float4 xa = *reinterpret_cast<const float4*>(cm + S::ub1 + j * W + q4 * 8);
          const float4 xb = *reinterpret_cast<const float4*>(cm + S::ub1 + j * W + q4 * 8 + 4);
          v[j] = dot8(wa, wb, xa, xb);
        } else {
          v[j] = 0.f;
        }
      }
      xstep<4>(v, q4 & 1, 1);
      xstep<2>(v, q4 & 2, 2);
      const int jv = (q4 & 1) * 4 + ((q4 >> 1) & 1) * 2;
      if (jv < C) cm[S::wbar + jv * H + h] = v[0];
      if (jv + 1 < C) cm[S::wbar + (jv + 1) * H + h] = v[1];
      if (tid < W) {
        float s = 0.f;
#pragma unroll
        for (int j = 0; j < C; ++j) s += cm[S::sp1 + j * W + tid];
        cm[S::sb1 + tid] = s;
      }
    }
    __syncthreads();
  }
  // Q5: o-bar += Lam w-bar ; z3-bar = (1 - o^2) o-bar ; a2-bar partials
  if (tid < C * 64) {
    const int j = tid >> 6, h0 = row0 & (H - 1);
    if (D2) {
#pragma unroll
      for (int jj = 0; jj < C; ++jj) {
        const float l = cm[S::lam + j * 8 + jj];
        const float2 wv = *reinterpret_cast<const float2*>(cm + S::wbar + jj * H + h0);
        ob[0] = fmaf(l, wv.x, ob[0]);
        ob[1] = fmaf(l, wv.y, ob[1]);
      }
    }
    const float2 zb = make_float2(tp[0] * ob[0], tp[1] * ob[1]);
    b3acc[0] += zb.x;
    b3acc[1] += zb.y;
    *reinterpret_cast<float2*>(wsp + ws.ZL + slot * CH + row0) = zb;
    *reinterpret_cast<float2*>(cm + S::dbuf + row0) = zb;
    __syncwarp();
    const float part = tile_cols(R, cm + S::dbuf + (warp * 8 + (lane >> 2)) * 8, lane);
    cm[S::red + warp * W + lane_col(lane)] = part;
  }
  __syncthreads();
  // Q6: a2-bar ; z2-bar
  if (tid < W) {
    float ab2 = 0.f;
#pragma unroll
    for (int wq = 0; wq < 2 * C; ++wq) ab2 += cm[S::red + wq * W + tid];
    float zb = sb[S::s2 + tid] * ab2;
    if (D2) {
      const float z = sb[S::z2 + tid], sg = sigm(z);
      zb = fmaf(sg * (1.f - sg) * (2.f + z * (1.f - 2.f * sg)), cm[S::sb2 + tid], zb);
    }
    cm[S::zb2 + tid] = zb;
    cm[S::bacc + W + tid] += zb;
    wsp[ws.Z[1] + slot * W + tid] = zb;
  }
  __syncthreads();
  // Q7: a1-bar = W2^T z2-bar ; z1-bar
  {
    const float2 w2t = *reinterpret_cast<const float2*>(sw + S::W2T + tid * 2);
    const float2 x = *reinterpret_cast<const float2*>(cm + S::zb2 + q * 2);
    float acc = fmaf(w2t.y, x.y, w2t.x * x.x);
    acc += __shfl_xor_sync(0xffffffffu, acc, 8);
    acc += __shfl_xor_sync(0xffffffffu, acc, 4);
    acc += __shfl_xor_sync(0xffffffffu, acc, 2);
    acc += __shfl_xor_sync(0xffffffffu, acc, 1);
    if (q == 0) {
      float zb = sb[S::s1 + o] * acc;
      if (D2) {
        const float z = sb[S::z1 + o], sg = sigm(z);
        zb = fmaf(sg * (1.f - sg) * (2.f + z * (1.f - 2.f * sg)), cm[S::sb1 + o], zb);
      }
      cm[S::zb1 + o] = zb;
      cm[S::bacc + o] += zb;
      wsp[ws.Z[0] + slot * W + o] = zb;
    }
  }
  __syncthreads();
  // Q8: y-bar = W1^T z1-bar
  {
    const int h = tid >> 2, q4 = tid & 3;
    const float4 wa = *reinterpret_cast<const float4*>(sw + S::W1T + tid * 8);
    const float4 wb = *reinterpret_cast<const float4*>(sw + S::W1T + tid * 8 + 4);
    const float4 xa = *reinterpret_cast<const float4*>(cm + S::zb1 + q4 * 8);
    const float4 xb = *reinterpret_cast<const float4*>(cm + S::zb1 + q4 * 8 + 4);
    float acc = dot8(wa, wb, xa, xb);
    acc += __shfl_xor_sync(0xffffffffu, acc, 1);
    acc += __shfl_xor_sync(0xffffffffu, acc, 2);
    if (q4 == 0) cm[S::yout + h] = acc;
  }
  __syncthreads();
}

template <int C, bool D2>
__global__ void __launch_bounds__(NT, 1) logncde_bwd_fast_kernel(const BwdParams p) {
  using S = Smem<C>;
  extern __shared__ __align__(16) float smem[];
  float* sw = smem;
  float* sbA = smem + S::WEND_BWD;
  float* sbB = sbA + S::STAGE;
  float* cm = sbB + S::STAGE;
  const int tid = threadIdx.x;
  Regs R;
  load_regs<C>(R, p.params, p.d);
  stage_weights_fast<C>(p.d, p.params, sw, true);
  const int b = blockIdx.x;
  const float* ls = p.logsig + (size_t)b * p.K * p.D;
  const float* ys = p.ys + (size_t)b * (p.n_steps + 1) * H;
  const float* gys = p.g_ys + (size_t)b * (p.n_steps + 1) * H;
  if (tid < 2 * W) cm[S::bacc + tid] = 0.f;
  if (tid < H) cm[S::ybar + tid] = gys[(size_t)p.n_steps * H + tid];
  float b3acc[2] = {0.f, 0.f};
  float oA[2], uA[2], oB[2], uB[2];
  __syncthreads();
  for (int n = p.n_steps - 1; n >= 0; --n) {
    const int r1 = p.stage_row[2 * n], r2 = p.stage_row[2 * n + 1];
    const float s1 = p.stage_scale[2 * n], s2 = p.stage_scale[2 * n + 1];
    const size_t slot = ((size_t)b * p.n_steps + n) * 2;
    if (tid < H) sbA[S::yin + tid] = ys[(size_t)n * H + tid];
    __syncthreads();
    stage_forward<C, D2>(sw, sbA, cm, R, ls + (size_t)r1 * p.D, oA, uA);
    if (tid < H) sbB[S::yin + tid] = sbA[S::yin + tid] + s1 * field_value<C, D2>(sbA, cm, tid);
    __syncthreads();
    stage_forward<C, D2>(sw, sbB, cm, R, ls + (size_t)r2 * p.D, oB, uB);  // leaves Lambda of row r2
    if (tid < H) cm[S::g + tid] = 0.5f * s2 * cm[S::ybar + tid];
    __syncthreads();
    stage_reverse<C, D2>(sw, sbB, cm, R, oB, uB, b3acc, p.ws, p.wsp, slot + 1);
    if (tid < H) {
      const float a = cm[S::yout + tid];
      cm[S::ab + tid] = a;
      cm[S::g + tid] = s1 * (0.5f * cm[S::ybar + tid] + a);
    }
    load_lambda<C>(cm, ls + (size_t)r1 * p.D, D2);
    __syncthreads();
    stage_reverse<C, D2>(sw, sbA, cm, R, oA, uA, b3acc, p.ws, p.wsp, slot);
    if (tid < H) cm[S::ybar + tid] += cm[S::ab + tid] + cm[S::yout + tid] + gys[(size_t)n * H + tid];
    __syncthreads();
  }
  if (tid < H) p.g_y0[(size_t)b * H + tid] = cm[S::ybar + tid];
  const int n_bias = 2 * W + C * H;
  float* bp = p.wsp + p.ws.bias_part + (size_t)b * n_bias;
  if (tid < 2 * W) bp[tid] = cm[S::bacc + tid];
  if (tid < C * 64) {
    const int row0 = lane_row0(tid);
    bp[2 * W + row0] = b3acc[0];
    bp[2 * W + row0 + 1] = b3acc[1];
  }
}

template <int C>
constexpr size_t fwd_smem_bytes() {
  return (size_t)(Smem<C>::WEND_FWD + Smem<C>::STAGE + Smem<C>::COMMON_FWD) * 4;
}
template <int C>
constexpr size_t bwd_smem_bytes() {
  return (size_t)(Smem<C>::WEND_BWD + 2 * Smem<C>::STAGE + Smem<C>::COMMON_BWD) * 4;
}

template <int C>
static int launch_fwd_fast(cudaStream_t st, const FwdParams& p) {
  const size_t smem = fwd_smem_bytes<C>();
  if (p.d.depth == 2) {
    auto k = logncde_fwd_fast_kernel<C, true>;
    cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    k<<<p.B, NT, smem, st>>>(p);
  } else {
    auto k = logncde_fwd_fast_kernel<C, false>;
    cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    k<<<p.B, NT, smem, st>>>(p);
  }
  return (int)cudaGetLastError();
}
template <int C>
static int launch_bwd_fast(cudaStream_t st, const BwdParams& p) {
  const size_t smem = bwd_smem_bytes<C>();
  if (p.d.depth == 2) {
    auto k = logncde_bwd_fast_kernel<C, true>;
    cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    k<<<p.B, NT, smem, st>>>(p);
  } else {
    auto k = logncde_bwd_fast_kernel<C, false>;
    cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    k<<<p.B, NT, smem, st>>>(p);
  }
  return (int)cudaGetLastError();
}

inline bool shape_supported(int Hh, int C, int width, int n_hidden) {
  return Hh == H && width == W && n_hidden == 2 && (C == 4 || C == 6 || C == 7);
}

}  // namespace fast
}  // namespace lncde
