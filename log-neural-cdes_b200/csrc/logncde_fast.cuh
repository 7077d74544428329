// Specialised Log-NCDE solve kernels for the EigenWorms model shape
// (hidden H = 128, vector-field width 32, 2 hidden layers; C in {4, 6, 7} channels).
//
// Same algorithm and workspace layout as the generic kernels in logncde.cu, organised so
// that the instruction count and the shared-memory wavefronts per Heun stage are close to
// the floor (ncu of the first version: 3 860 smem wavefronts / 9 300 warp-instructions per
// stage, LSU-bound; see profiles/).  A warp-specialised variant (hidden layers on single
// warps, 5 barriers per stage instead of 12) was measured SLOWER (10.5 vs 7.2 ms forward):
// with 64+ registers pinned by the last-layer tile a lone warp cannot keep enough loads in
// flight, so the K-sliced all-warp form below is kept:
//   * last MLP layer (C*128 x 32, 80 % of the weights) in REGISTERS: each of the C*64 "row
//     threads" owns an 8 x 8 tile (8 consecutive output rows x 8 inputs).  The same tile
//     serves W a (forward: reduce over the 4 column groups = lane bits 0-1) and W^T v
//     (adjoint: reduce over the 8 row groups = lane bits 2-4) through shuffle
//     transpose-reductions that halve the live values at every step.
//   * hidden layers (32 x 128, 32 x 32): lane = output row, warp = K-slice.  Every lane of
//     a warp reads the SAME input slice (one broadcast wavefront per 128-bit load), weights
//     sit in registers, K-slices are combined through shared memory - no shuffles.
//   * the antisymmetric level-2 mixing w_j = sum_i Lam_ij f_i is moved through the first
//     layer by linearity:  W1 w_j = sum_i Lam_ij (W1 f_i), so it acts on 32-vectors instead
//     of 128-vectors (forward) and  Lam^T-mixing of u1-bar gives v_i with
//     o-bar_i += W1^T v_i,  dW1 += sum_i v_i (x) f_i  (adjoint).
//   * all loops compile-time; one __syncthreads per dependent phase.
#pragma once
#include <cuda_runtime.h>

#include <cstdint>

#include "ptx.cuh"

namespace lncde {
namespace fast {

constexpr int H = 128, W = 32, NT = 512;

// Block barrier of the solve kernels.  A -DLNCDE_PHASE_TIMING build (dev tool, scripts/phase_timing.py) records
// the clock at which every warp of CTA 0 ARRIVES at every barrier of a few stages, keyed by source line, so
// that the time between consecutive barriers (= one dependent phase) can be read off per phase.
#ifdef LNCDE_PHASE_TIMING
constexpr int kPtStages = 4, kPtLines = 1280, kPtWarps = NT / 32, kPtStage0 = 200;
__device__ long long g_phase_clock[kPtStages * kPtLines * kPtWarps];
__device__ __forceinline__ void pt_sync(int slot, int line) {
  if (slot >= 0 && blockIdx.x == 0 && (threadIdx.x & 31) == 0 && line < kPtLines)
    g_phase_clock[(slot * kPtLines + line) * kPtWarps + (threadIdx.x >> 5)] = clock64();
  __syncthreads();
}
__device__ __forceinline__ int pt_slot_of(int stage) {
  return (stage >= kPtStage0 && stage < kPtStage0 + kPtStages) ? stage - kPtStage0 : -1;
}
#define LNCDE_SYNC() pt_sync(pt_slot, __LINE__)
#else
__device__ __forceinline__ int pt_slot_of(int) { return -1; }
#define LNCDE_SYNC() __syncthreads()
#endif

// logistic function with the fast exp / reciprocal units (no cancellation here: |error| ~ 2 ulp)
__device__ __forceinline__ float sigm(float z) { return __frcp_rn(1.f + __expf(-z)); }

// One halving step of a transpose-reduction across lanes lane^mask.
template <int HALF>
__device__ __forceinline__ void xstep(float* v, bool upper, int mask) {
#pragma unroll
  for (int k = 0; k < HALF; ++k) {
    const float send = upper ? v[k] : v[k + HALF];
    const float keep = upper ? v[k + HALF] : v[k];
    v[k] = keep + __shfl_xor_sync(0xffffffffu, send, mask);
  }
}

__device__ __forceinline__ float dot4(const float* w, const float4& x) {
  float s = w[0] * x.x;
  s = fmaf(w[1], x.y, s);
  s = fmaf(w[2], x.z, s);
  return fmaf(w[3], x.w, s);
}
__device__ __forceinline__ float dot8(const float* w, const float4& x0, const float4& x1) {
  return dot4(w, x0) + dot4(w + 4, x1);
}

template <int C>
struct Smem {
  static constexpr int CH = C * H, CW = C * W;
  // weight region (backward only): transposed small layers in per-thread order
  static constexpr int W1T = 0;              // [16 warps][2 halves][32 lanes][4]
  static constexpr int W2T = W1T + H * W;    // [8 slices][32 lanes][4]
  static constexpr int WEND_FWD = 0;
  static constexpr int WEND_BWD = W2T + W * W;
  // stage buffer (relative)
  static constexpr int yin = 0, z1 = yin + H, a1 = z1 + W, s1 = a1 + W, z2 = s1 + W, a2 = z2 + W, s2 = a2 + W;
  static constexpr int o = s2 + W, u1 = o + CH, p1 = u1 + CW, u2 = p1 + CW, p2 = u2 + CW;
  static constexpr int STAGE = p2 + CW;
  // common (relative to the common base)
  static constexpr int lam = 0;              // [8][8]
  static constexpr int lam1 = lam + 64;      // [8]
  static constexpr int dbuf = lam1 + 8;      // [CH]  last-layer tangent outputs / row staging
  static constexpr int part = dbuf + CH;     // [4096] K-slice partials
  static constexpr int qv = part + 4096;     // [C][32]
  static constexpr int ycur = qv + CW, k1 = ycur + H;
  static constexpr int COMMON_FWD = k1 + H;
  // backward extras
  static constexpr int g = COMMON_FWD, ybar = g + H, ab = ybar + H;
  static constexpr int red = ab + H;         // [16][32]
  static constexpr int ub2 = red + 16 * W, ub1 = ub2 + CW, sp2 = ub1 + CW, sp1 = sp2 + CW, vb = sp1 + CW;
  static constexpr int sb2 = vb + CW, sb1 = sb2 + W, zb2 = sb1 + W, zb1 = zb2 + W;
  static constexpr int bacc = zb1 + W;       // [64] bias accumulators of the two hidden layers
  static constexpr int COMMON_BWD = bacc + 2 * W;
};

struct Regs {
  float2 Wt[8][4];  // last-layer tile: rows r0..r0+7, inputs cg*8..cg*8+7 as (even, odd) pairs for FFMA2
  float b3[2];     // bias of the two rows this lane finishes
  float w1[8];     // W1[lane][warp*8 .. +7]
  float w2[4];     // W2[lane][(warp&7)*4 .. +3]
  float b1, b2;    // hidden-layer biases of output row = lane
};

__device__ __forceinline__ int lane_row0(int tid) {
  const int lane = tid & 31, warp = tid >> 5;
  return (warp * 8 + (lane >> 2)) * 8 + (lane & 1) * 4 + ((lane >> 1) & 1) * 2;
}

template <int C>
__device__ __forceinline__ void load_regs(Regs& R, const float* __restrict__ params, const MlpDesc& d) {
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  R.b1 = params[d.b_off[0] + lane];
  R.b2 = params[d.b_off[1] + lane];
#pragma unroll
  for (int e = 0; e < 8; ++e) R.w1[e] = params[d.w_off[0] + lane * H + warp * 8 + e];
#pragma unroll
  for (int e = 0; e < 4; ++e) R.w2[e] = params[d.w_off[1] + lane * W + (warp & 7) * 4 + e];
  if (tid < C * 64) {
    const int r0 = (warp * 8 + (lane >> 2)) * 8, cg = lane & 3;
#pragma unroll
    for (int e = 0; e < 8; ++e)
#pragma unroll
      for (int q = 0; q < 4; ++q)
        R.Wt[e][q] = *reinterpret_cast<const float2*>(params + d.w_off[2] + (size_t)(r0 + e) * W + cg * 8 + 2 * q);
    const int row0 = lane_row0(tid);
    R.b3[0] = params[d.b_off[2] + row0];
    R.b3[1] = params[d.b_off[2] + row0 + 1];
  }
}

// The 8x8 register tile runs on packed fp32 pairs (FFMA2: two fused multiply-adds per issue slot, each half rounded
// like a scalar FFMA).  tile_rows pairs the inputs (even, odd) and adds the two halves at the end; tile_cols pairs
// the outputs and needs the input duplicated into both halves.
__device__ __forceinline__ void tile_rows(const Regs& R, const float* __restrict__ x8, int lane, float& r_a,
                                          float& r_b) {
  const float4 x0 = *reinterpret_cast<const float4*>(x8);
  const float4 x1 = *reinterpret_cast<const float4*>(x8 + 4);
  const float2 xa = make_float2(x0.x, x0.y), xb = make_float2(x0.z, x0.w);
  const float2 xc = make_float2(x1.x, x1.y), xd = make_float2(x1.z, x1.w);
  float v[8];
#pragma unroll
  for (int e = 0; e < 8; ++e) {
    float2 s = __fmul2_rn(R.Wt[e][0], xa);
    s = __ffma2_rn(R.Wt[e][1], xb, s);
    s = __ffma2_rn(R.Wt[e][2], xc, s);
    s = __ffma2_rn(R.Wt[e][3], xd, s);
    v[e] = s.x + s.y;
  }
  xstep<4>(v, lane & 1, 1);
  xstep<2>(v, lane & 2, 2);
  r_a = v[0];
  r_b = v[1];
}

__device__ __forceinline__ float tile_cols(const Regs& R, const float* __restrict__ x8, int lane) {
  const float4 x0 = *reinterpret_cast<const float4*>(x8);
  const float4 x1 = *reinterpret_cast<const float4*>(x8 + 4);
  const float xe[8] = {x0.x, x0.y, x0.z, x0.w, x1.x, x1.y, x1.z, x1.w};
  float2 s[4];
#pragma unroll
  for (int q = 0; q < 4; ++q) s[q] = __fmul2_rn(R.Wt[0][q], make_float2(xe[0], xe[0]));
#pragma unroll
  for (int e = 1; e < 8; ++e) {
    const float2 xx = make_float2(xe[e], xe[e]);
#pragma unroll
    for (int q = 0; q < 4; ++q) s[q] = __ffma2_rn(R.Wt[e][q], xx, s[q]);
  }
  float v[8];
#pragma unroll
  for (int q = 0; q < 4; ++q) {
    v[2 * q] = s[q].x;
    v[2 * q + 1] = s[q].y;
  }
  xstep<4>(v, lane & 4, 4);
  xstep<2>(v, lane & 8, 8);
  xstep<1>(v, lane & 16, 16);
  return v[0];
}
__device__ __forceinline__ int lane_col(int lane) {
  return (lane & 3) * 8 + ((lane >> 2) & 1) * 4 + ((lane >> 3) & 1) * 2 + ((lane >> 4) & 1);
}

// Level-1 / level-2 coordinates of one logsig row.  Threads 0..63 own one entry of the
// antisymmetric matrix Lam, threads 64..71 one level-1 coefficient.  The global load is
// issued one stage AHEAD (lambda_fetch) so its latency is off the dependent chain.
template <int C>
__device__ __forceinline__ float lambda_fetch(const float* __restrict__ row, bool d2) {
  const int tid = threadIdx.x;
  float v = 0.f;
  if (tid < 64) {
    const int i = tid >> 3, j = tid & 7;
    if (d2 && i < C && j < C && i != j) {
      const int a = i < j ? i : j, b = i < j ? j : i;
      v = __ldg(row + 1 + C + a * C - (a * (a + 1)) / 2 + (b - a - 1));  // sign applied at store time
    }
  } else if (tid < 72) {
    const int i = tid - 64;
    if (i < C) v = __ldg(row + 1 + i);
  }
  return v;
}
template <int C>
__device__ __forceinline__ void lambda_store(float* cm, float v) {
  using S = Smem<C>;
  const int tid = threadIdx.x;
  if (tid < 72) cm[S::lam + tid] = (tid < 64 && (tid >> 3) > (tid & 7)) ? -v : v;  // lam1 follows lam
}

// z = b + sum_q part[q*32 + lane] over NQ K-slices, then SiLU and its derivative.
template <int NQ>
__device__ __forceinline__ void hidden_finish(const float* part, float bias, int lane, float* z_out, float* a_out,
                                              float* s_out) {
  float acc[4] = {bias, 0.f, 0.f, 0.f};
#pragma unroll
  for (int qq = 0; qq < NQ; ++qq) acc[qq & 3] += part[qq * 32 + lane];
  const float z = (acc[0] + acc[1]) + (acc[2] + acc[3]);
  const float sg = sigm(z);
  z_out[lane] = z;
  a_out[lane] = z * sg;
  s_out[lane] = sg * (1.f + z * (1.f - sg));
}

// acc[q % NACC] += dot(Wq[(q*32 + lane)*4 .. +3], x[4q .. 4q+3]), q = 0..7 in order: one 32 x 32 mat-vec (or its
// transpose, depending on the layout of Wq) done by ONE warp, lane = output index, x read as broadcast
// 128-bit loads.  The partial products and the order they are summed in are those of the K-sliced path.
template <int NACC>
__device__ __forceinline__ void warp_matvec32(const float* __restrict__ Wq, const float* __restrict__ x, int lane,
                                              float (&acc)[NACC]) {
#pragma unroll
  for (int q = 0; q < 8; ++q) {
    const float4 wt = *reinterpret_cast<const float4*>(Wq + (q * 32 + lane) * 4);
    const float w4[4] = {wt.x, wt.y, wt.z, wt.w};
    acc[q % NACC] += dot4(w4, *reinterpret_cast<const float4*>(x + q * 4));
  }
}

// W2 in the order warp_matvec32 reads it for the FORWARD product: Wq[(q*32 + row)*4 + e] = W2[row][4q + e]
__device__ __forceinline__ void stage_w2r(const MlpDesc& d, const float* __restrict__ params, float* w2r) {
  for (int t = threadIdx.x; t < W * W; t += NT) {
    const int e = t & 3, row = (t >> 2) & 31, q = t >> 7;
    w2r[t] = params[d.w_off[1] + row * W + q * 4 + e];
  }
}

// Forward evaluation of the field at sb[yin].  Leaves every intermediate in `sb`, the
// last-layer outputs of this lane in o_r / u3_r and, after the trailing barrier, the per-row
// tangent contributions in cm[dbuf].  The caller finishes
//   G[h] = sum_i lam1[i] o[i][h] + sum_j dbuf[j][h].
// (Round 2 removed two variants that fused the 32 x 32 layer phases onto single warps - 8 / 7 block barriers per
// stage instead of 12 / 11: they measured 19.9 / 20.2 ms per step against 19.1 ms on B200.  The clock trace of
// logncde_v3.cuh shows why: a block barrier costs ~20 cycles here, the time is in the instruction streams.)
template <int C, bool D2>
__device__ __forceinline__ void stage_forward(float* sb, float* cm, const Regs& R, float lam_v,
                                              float (&o_r)[2], float (&u3_r)[2], int pt_slot = -1) {
  using S = Smem<C>;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  float* part = cm + S::part;
  lambda_store<C>(cm, lam_v);
  // P1: K-slice partials of z1 = W1 y
  {
    const float4 ya = *reinterpret_cast<const float4*>(sb + S::yin + warp * 8);
    const float4 yb = *reinterpret_cast<const float4*>(sb + S::yin + warp * 8 + 4);
    part[warp * 32 + lane] = dot8(R.w1, ya, yb);
  }
  LNCDE_SYNC();
  if (warp == 0) hidden_finish<16>(part, R.b1, lane, sb + S::z1, sb + S::a1, sb + S::s1);
  LNCDE_SYNC();
  // P2: z2 = W2 a1
  if (warp < 8) {
    const float4 a = *reinterpret_cast<const float4*>(sb + S::a1 + warp * 4);
    part[warp * 32 + lane] = dot4(R.w2, a);
  }
  LNCDE_SYNC();
  if (warp == 0) hidden_finish<8>(part, R.b2, lane, sb + S::z2, sb + S::a2, sb + S::s2);
  LNCDE_SYNC();
  // P3: o = tanh(W3 a2 + b3)
  const int row0 = lane_row0(tid);
  if (tid < C * 64) {
    float ra, rb;
    tile_rows(R, sb + S::a2 + (lane & 3) * 8, lane, ra, rb);
    o_r[0] = tanhf(ra + R.b3[0]);
    o_r[1] = tanhf(rb + R.b3[1]);
    *reinterpret_cast<float2*>(sb + S::o + row0) = make_float2(o_r[0], o_r[1]);
  }
  LNCDE_SYNC();
  if (!D2) return;
  // P5: q_i = W1 o_i, K-slice partials
#pragma unroll
  for (int i = 0; i < C; ++i) {
    const float4 xa = *reinterpret_cast<const float4*>(sb + S::o + i * H + warp * 8);
    const float4 xb = *reinterpret_cast<const float4*>(sb + S::o + i * H + warp * 8 + 4);
    part[(warp * 8 + i) * 32 + lane] = dot8(R.w1, xa, xb);
  }
  LNCDE_SYNC();
  if (warp < C) {
    float acc[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
    for (int qq = 0; qq < 16; ++qq) acc[qq & 3] += part[(qq * 8 + warp) * 32 + lane];
    cm[S::qv + warp * W + lane] = (acc[0] + acc[1]) + (acc[2] + acc[3]);
  }
  LNCDE_SYNC();
  // u1_j = sum_i Lam_ij q_i ; p1_j = s1 * u1_j
  if (warp < C) {
    float u = 0.f;
#pragma unroll
    for (int i = 0; i < C; ++i) u = fmaf(cm[S::lam + i * 8 + warp], cm[S::qv + i * W + lane], u);
    sb[S::u1 + warp * W + lane] = u;
    sb[S::p1 + warp * W + lane] = sb[S::s1 + lane] * u;
  }
  LNCDE_SYNC();
  // P6: u2_j = W2 p1_j (8 K-slices x 2 vector halves)
  {
    const int q = warp & 7, j0 = (warp >> 3) * 4;
#pragma unroll
    for (int jj = 0; jj < 4; ++jj) {
      const int j = j0 + jj;
      if (j < C) {
        const float4 x = *reinterpret_cast<const float4*>(sb + S::p1 + j * W + q * 4);
        part[(q * 8 + j) * 32 + lane] = dot4(R.w2, x);
      }
    }
  }
  LNCDE_SYNC();
  if (warp < C) {
    float acc[2] = {0.f, 0.f};
#pragma unroll
    for (int qq = 0; qq < 8; ++qq) acc[qq & 1] += part[(qq * 8 + warp) * 32 + lane];
    const float u = acc[0] + acc[1];
    sb[S::u2 + warp * W + lane] = u;
    sb[S::p2 + warp * W + lane] = sb[S::s2 + lane] * u;
  }
  LNCDE_SYNC();
  // P7: u3 = W3[block j] p2_j ; d = (1 - o^2) u3
  if (tid < C * 64) {
    const int j = tid >> 6;
    tile_rows(R, sb + S::p2 + j * W + (lane & 3) * 8, lane, u3_r[0], u3_r[1]);
    *reinterpret_cast<float2*>(cm + S::dbuf + row0) =
        make_float2((1.f - o_r[0] * o_r[0]) * u3_r[0], (1.f - o_r[1] * o_r[1]) * u3_r[1]);
  }
  LNCDE_SYNC();
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

// transposed small layers, stored in the order the backward threads read them
template <int C>
__device__ __forceinline__ void stage_weights_bwd(const MlpDesc& d, const float* __restrict__ params, float* sw) {
  using S = Smem<C>;
  // W1T: thread (lane, hq = warp&3, ks = warp>>2) handles h = hq*32 + lane, inputs c = ks*8 + half*4 + e
  for (int t = threadIdx.x; t < W * H; t += NT) {
    const int e = t & 3, lane = (t >> 2) & 31, half = (t >> 7) & 1, warp = t >> 8;
    const int h = (warp & 3) * 32 + lane, c = (warp >> 2) * 8 + half * 4 + e;
    sw[S::W1T + t] = params[d.w_off[0] + c * H + h];
  }
  // W2T: thread (lane = c', q) : inputs c = q*4 + e
  for (int t = threadIdx.x; t < W * W; t += NT) {
    const int e = t & 3, lane = (t >> 2) & 31, q = t >> 7;
    sw[S::W2T + t] = params[d.w_off[1] + (q * 4 + e) * W + lane];
  }
}

// ------------------------------------------------------------------------------ forward
// Training-mode forward: keep what the adjoint sweep needs of one stage in the workspace slot
// (the right-hand vectors of the weight-gradient pairs in their final place, plus z/s of the
// hidden layers and the tangent pre-activations).  Must run before sb[yin] is overwritten.
template <int C, bool D2>
__device__ __forceinline__ void save_stage(const float* sb, const float (&o_r)[2], const float (&u3_r)[2],
                                           const WsLayout& ws, float* __restrict__ wsp, size_t slot) {
  using S = Smem<C>;
  constexpr int CH = C * H, CW = C * W;
  const int tid = threadIdx.x;
  if (tid < H) {
    wsp[ws.Y + slot * H + tid] = sb[S::yin + tid];
    const int l = tid >> 5, c = tid & 31;  // ZS = [z1 | s1 | z2 | s2]
    const int src = (l == 0) ? S::z1 : (l == 1) ? S::s1 : (l == 2) ? S::z2 : S::s2;
    wsp[ws.ZS + slot * 4 * W + tid] = sb[src + c];
  }
  if (tid < W) {
    wsp[ws.A[0] + slot * W + tid] = sb[S::a1 + tid];
    wsp[ws.A[1] + slot * W + tid] = sb[S::a2 + tid];
  }
  if (tid < C * 64) {
    const int row0 = lane_row0(tid);
    *reinterpret_cast<float2*>(wsp + ws.Wt + slot * CH + row0) = make_float2(o_r[0], o_r[1]);
    if (D2) *reinterpret_cast<float2*>(wsp + ws.U3 + slot * CH + row0) = make_float2(u3_r[0], u3_r[1]);
  }
  if (D2 && tid < CW) {
    wsp[ws.P[0] + slot * CW + tid] = sb[S::p1 + tid];
    wsp[ws.P[1] + slot * CW + tid] = sb[S::p2 + tid];
    wsp[ws.US + slot * 2 * CW + tid] = sb[S::u1 + tid];
    wsp[ws.US + slot * 2 * CW + CW + tid] = sb[S::u2 + tid];
  }
}

template <int C, bool D2, bool SAVE>
__global__ void __launch_bounds__(NT, 1) logncde_fwd_fast_kernel(const FwdParams p) {
  using S = Smem<C>;
  extern __shared__ __align__(16) float smem[];
  float* sb = smem + S::WEND_FWD;
  float* cm = sb + S::STAGE;
  const int tid = threadIdx.x;
  Regs R;
  load_regs<C>(R, p.params, p.d);
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
  // ONE inlined stage body, looped over the 2*n_steps stages (keeps the loop inside the 32 KB
  // instruction cache); stage table entries and logsig rows are fetched one stage ahead
  const int n_stage = 2 * p.n_steps;
  float sc = p.stage_scale[0];
  float lam_c = lambda_fetch<C>(ls + (size_t)p.stage_row[0] * p.D, D2);
  for (int sidx = 0; sidx < n_stage; ++sidx) {
    const int nx = sidx + 1 < n_stage ? sidx + 1 : sidx;
    const float sc_n = p.stage_scale[nx];
    const float lam_n = lambda_fetch<C>(ls + (size_t)p.stage_row[nx] * p.D, D2);
    const int pt_slot = pt_slot_of(sidx);
    stage_forward<C, D2>(sb, cm, R, lam_c, o_r, u3_r, pt_slot);
    if (SAVE) save_stage<C, D2>(sb, o_r, u3_r, p.ws, p.wsp, (size_t)b * n_stage + sidx);
    if (tid < H) {
      const float gk = sc * field_value<C, D2>(sb, cm, tid);
      if (!(sidx & 1)) {  // first Heun stage: k1, then evaluate at y + k1
        cm[S::k1 + tid] = gk;
        sb[S::yin + tid] = cm[S::ycur + tid] + gk;
      } else {            // second stage: y_{n+1} = y_n + (k1 + k2) / 2
        const float yn = cm[S::ycur + tid] + 0.5f * (cm[S::k1 + tid] + gk);
        cm[S::ycur + tid] = yn;
        sb[S::yin + tid] = yn;
        ys[(size_t)((sidx >> 1) + 1) * H + tid] = yn;
      }
    }
    LNCDE_SYNC();
    sc = sc_n;
    lam_c = lam_n;
  }
}

// ------------------------------------------------------------------------------ backward
// Reverse sweep of one stage.  cm[g] holds dLoss/dG times the stage scale.  On exit (after the
// trailing barrier) cm[part + ks*128 + h], ks < 4, hold the K-slice partials of
// dLoss/d(stage input); the caller sums them.
template <int C, bool D2, bool RIGHTS>
__device__ __forceinline__ void stage_reverse(const float* __restrict__ sw, float* sb, float* cm, const Regs& R,
                                              const float (&o_r)[2], const float (&u3_r)[2], float (&b3acc)[2],
                                              const WsLayout& ws, float* __restrict__ wsp, size_t slot,
                                              int pt_slot = -1) {
  using S = Smem<C>;
  constexpr int CH = C * H, CW = C * W;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int row0 = lane_row0(tid);
  float* part = cm + S::part;
  // right-hand vectors of the weight-gradient outer products (already in place when the forward saved them)
  if (RIGHTS && tid < H) wsp[ws.Y + slot * H + tid] = sb[S::yin + tid];
  if (RIGHTS && tid < W) {
    wsp[ws.A[0] + slot * W + tid] = sb[S::a1 + tid];
    wsp[ws.A[1] + slot * W + tid] = sb[S::a2 + tid];
  }
  if (RIGHTS && D2) {
    for (int t = tid; t < CH; t += NT) wsp[ws.Wt + slot * CH + t] = sb[S::o + t];  // pairs with v_i (see header)
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
      const float pr = tile_cols(R, cm + S::dbuf + (warp * 8 + (lane >> 2)) * 8, lane);
      cm[S::red + warp * W + lane_col(lane)] = pr;
    }
  }
  if (D2) {
    LNCDE_SYNC();
    // Q2: u2-bar = s2 * p2-bar ; s2-bar partials
    if (warp < C) {
      const float pb = cm[S::red + (2 * warp) * W + lane] + cm[S::red + (2 * warp + 1) * W + lane];
      const float ub = sb[S::s2 + lane] * pb;
      cm[S::ub2 + tid] = ub;
      cm[S::sp2 + tid] = sb[S::u2 + tid] * pb;
      wsp[ws.U[1] + slot * CW + tid] = ub;
    }
    LNCDE_SYNC();
    // Q3: p1-bar_j = W2^T u2-bar_j  (8 K-slices x 2 vector halves) ; s2-bar
    {
      const int q = warp & 7, j0 = (warp >> 3) * 4;
      const float4 wt = *reinterpret_cast<const float4*>(sw + S::W2T + (q * 32 + lane) * 4);
      const float w4[4] = {wt.x, wt.y, wt.z, wt.w};
#pragma unroll
      for (int jj = 0; jj < 4; ++jj) {
        const int j = j0 + jj;
        if (j < C) {
          const float4 x = *reinterpret_cast<const float4*>(cm + S::ub2 + j * W + q * 4);
          part[(q * 8 + j) * 32 + lane] = dot4(w4, x);
        }
      }
      if (warp == 15) {
        float s = 0.f;
#pragma unroll
        for (int j = 0; j < C; ++j) s += cm[S::sp2 + j * W + lane];
        cm[S::sb2 + lane] = s;
      }
    }
    LNCDE_SYNC();
    // u1-bar = s1 * p1-bar ; s1-bar partials
    if (warp < C) {
      float acc[2] = {0.f, 0.f};
#pragma unroll
      for (int qq = 0; qq < 8; ++qq) acc[qq & 1] += part[(qq * 8 + warp) * 32 + lane];
      const float pb = acc[0] + acc[1];
      cm[S::ub1 + tid] = sb[S::s1 + lane] * pb;
      cm[S::sp1 + tid] = sb[S::u1 + tid] * pb;
    }
    LNCDE_SYNC();
    // v_i = sum_j Lam_ij u1-bar_j ; s1-bar
    if (warp < C) {
      float v = 0.f;
#pragma unroll
      for (int j = 0; j < C; ++j) v = fmaf(cm[S::lam + warp * 8 + j], cm[S::ub1 + j * W + lane], v);
      cm[S::vb + tid] = v;
      wsp[ws.U[0] + slot * CW + tid] = v;
    } else if (warp == 15) {
      float s = 0.f;
#pragma unroll
      for (int j = 0; j < C; ++j) s += cm[S::sp1 + j * W + lane];
      cm[S::sb1 + lane] = s;
    }
    LNCDE_SYNC();
    // Q4: o-bar_i[h] += sum_c W1[c][h] v_i[c]   (4 K-slices x 4 h-quarters)
    {
      const int ks = warp >> 2, h = (warp & 3) * 32 + lane;
      const float4 wa = *reinterpret_cast<const float4*>(sw + S::W1T + ((warp * 2 + 0) * 32 + lane) * 4);
      const float4 wb = *reinterpret_cast<const float4*>(sw + S::W1T + ((warp * 2 + 1) * 32 + lane) * 4);
      const float w8[8] = {wa.x, wa.y, wa.z, wa.w, wb.x, wb.y, wb.z, wb.w};
#pragma unroll
      for (int i = 0; i < C; ++i) {
        const float4 xa = *reinterpret_cast<const float4*>(cm + S::vb + i * W + ks * 8);
        const float4 xb = *reinterpret_cast<const float4*>(cm + S::vb + i * W + ks * 8 + 4);
        part[(ks * 8 + i) * H + h] = dot8(w8, xa, xb);
      }
      LNCDE_SYNC();
    }
  }
  // Q5: o-bar complete ; z3-bar = (1 - o^2) o-bar ; a2-bar partials
  if (tid < C * 64) {
    const int j = tid >> 6, h0 = row0 & (H - 1);
    if (D2) {
#pragma unroll
      for (int ks = 0; ks < 4; ++ks) {
        const float2 pv = *reinterpret_cast<const float2*>(part + (ks * 8 + j) * H + h0);
        ob[0] += pv.x;
        ob[1] += pv.y;
      }
    }
    const float2 zb = make_float2(tp[0] * ob[0], tp[1] * ob[1]);
    b3acc[0] += zb.x;
    b3acc[1] += zb.y;
    *reinterpret_cast<float2*>(wsp + ws.ZL + slot * CH + row0) = zb;
    *reinterpret_cast<float2*>(cm + S::dbuf + row0) = zb;
    __syncwarp();
    const float pr = tile_cols(R, cm + S::dbuf + (warp * 8 + (lane >> 2)) * 8, lane);
    cm[S::red + warp * W + lane_col(lane)] = pr;
  }
  LNCDE_SYNC();
  // Q6: a2-bar ; z2-bar
  if (warp == 0) {
    float acc[2] = {0.f, 0.f};
#pragma unroll
    for (int wq = 0; wq < 2 * C; ++wq) acc[wq & 1] += cm[S::red + wq * W + lane];
    float zb = sb[S::s2 + lane] * (acc[0] + acc[1]);
    if (D2) {
      const float z = sb[S::z2 + lane], sg = sigm(z);
      zb = fmaf(sg * (1.f - sg) * (2.f + z * (1.f - 2.f * sg)), cm[S::sb2 + lane], zb);
    }
    cm[S::zb2 + lane] = zb;
    cm[S::bacc + W + lane] += zb;
    wsp[ws.Z[1] + slot * W + lane] = zb;
  }
  LNCDE_SYNC();
  // Q7: a1-bar = W2^T z2-bar (8 K-slices)
  if (warp < 8) {
    const float4 wt = *reinterpret_cast<const float4*>(sw + S::W2T + (warp * 32 + lane) * 4);
    const float w4[4] = {wt.x, wt.y, wt.z, wt.w};
    const float4 x = *reinterpret_cast<const float4*>(cm + S::zb2 + warp * 4);
    part[warp * 32 + lane] = dot4(w4, x);
  }
  LNCDE_SYNC();
  if (warp == 0) {
    float acc[2] = {0.f, 0.f};
#pragma unroll
    for (int qq = 0; qq < 8; ++qq) acc[qq & 1] += part[qq * 32 + lane];
    float zb = sb[S::s1 + lane] * (acc[0] + acc[1]);
    if (D2) {
      const float z = sb[S::z1 + lane], sg = sigm(z);
      zb = fmaf(sg * (1.f - sg) * (2.f + z * (1.f - 2.f * sg)), cm[S::sb1 + lane], zb);
    }
    cm[S::zb1 + lane] = zb;
    cm[S::bacc + lane] += zb;
    wsp[ws.Z[0] + slot * W + lane] = zb;
  }
  LNCDE_SYNC();
  // Q8: y-bar = W1^T z1-bar, K-slice partials (summed by the caller)
  {
    const int ks = warp >> 2, h = (warp & 3) * 32 + lane;
    const float4 wa = *reinterpret_cast<const float4*>(sw + S::W1T + ((warp * 2 + 0) * 32 + lane) * 4);
    const float4 wb = *reinterpret_cast<const float4*>(sw + S::W1T + ((warp * 2 + 1) * 32 + lane) * 4);
    const float w8[8] = {wa.x, wa.y, wa.z, wa.w, wb.x, wb.y, wb.z, wb.w};
    const float4 xa = *reinterpret_cast<const float4*>(cm + S::zb1 + ks * 8);
    const float4 xb = *reinterpret_cast<const float4*>(cm + S::zb1 + ks * 8 + 4);
    part[ks * H + h] = dot8(w8, xa, xb);
  }
  LNCDE_SYNC();
}

template <int C>
__device__ __forceinline__ float yout_value(const float* cm, int h) {
  using S = Smem<C>;
  const float* part = cm + S::part;
  return (part[h] + part[H + h]) + (part[2 * H + h] + part[3 * H + h]);
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
  stage_weights_bwd<C>(p.d, p.params, sw);
  const int b = blockIdx.x;
  const float* ls = p.logsig + (size_t)b * p.K * p.D;
  const float* ys = p.ys + (size_t)b * (p.n_steps + 1) * H;
  const float* gys = p.g_ys + (size_t)b * (p.n_steps + 1) * H;
  if (tid < 2 * W) cm[S::bacc + tid] = 0.f;
  if (tid < H) cm[S::ybar + tid] = gys[(size_t)p.n_steps * H + tid];
  float b3acc[2] = {0.f, 0.f};
  float oA[2], uA[2], oB[2], uB[2];
  // software pipeline: everything the next step reads from global memory is fetched one
  // step ahead (stage table, checkpoint y_n, output cotangent, the two logsig rows)
  const int n0 = p.n_steps - 1;
  float s1 = p.stage_scale[2 * n0], s2 = p.stage_scale[2 * n0 + 1];
  float lam_a = lambda_fetch<C>(ls + (size_t)p.stage_row[2 * n0] * p.D, D2);
  float lam_b = lambda_fetch<C>(ls + (size_t)p.stage_row[2 * n0 + 1] * p.D, D2);
  float y_n = tid < H ? ys[(size_t)n0 * H + tid] : 0.f;
  float g_n = tid < H ? gys[(size_t)n0 * H + tid] : 0.f;
  __syncthreads();
  for (int n = p.n_steps - 1; n >= 0; --n) {
    const size_t slot = ((size_t)b * p.n_steps + n) * 2;
    const int nn = n > 0 ? n - 1 : 0;
    const int r1n = p.stage_row[2 * nn], r2n = p.stage_row[2 * nn + 1];
    const float s1n = p.stage_scale[2 * nn], s2n = p.stage_scale[2 * nn + 1];
    const float y_nn = tid < H ? ys[(size_t)nn * H + tid] : 0.f;
    const float g_nn = tid < H ? gys[(size_t)nn * H + tid] : 0.f;
    if (tid < H) sbA[S::yin + tid] = y_n;
    __syncthreads();
    stage_forward<C, D2>(sbA, cm, R, lam_a, oA, uA);
    if (tid < H) sbB[S::yin + tid] = sbA[S::yin + tid] + s1 * field_value<C, D2>(sbA, cm, tid);
    __syncthreads();
    stage_forward<C, D2>(sbB, cm, R, lam_b, oB, uB);  // leaves Lambda of row r2
    if (tid < H) cm[S::g + tid] = 0.5f * s2 * cm[S::ybar + tid];
    __syncthreads();
    stage_reverse<C, D2, true>(sw, sbB, cm, R, oB, uB, b3acc, p.ws, p.wsp, slot + 1);
    if (tid < H) {
      const float a = yout_value<C>(cm, tid);
      cm[S::ab + tid] = a;
      cm[S::g + tid] = s1 * (0.5f * cm[S::ybar + tid] + a);
    }
    lambda_store<C>(cm, lam_a);
    const float lam_an = lambda_fetch<C>(ls + (size_t)r1n * p.D, D2);
    const float lam_bn = lambda_fetch<C>(ls + (size_t)r2n * p.D, D2);
    __syncthreads();
    stage_reverse<C, D2, true>(sw, sbA, cm, R, oA, uA, b3acc, p.ws, p.wsp, slot);
    if (tid < H) cm[S::ybar + tid] += cm[S::ab + tid] + yout_value<C>(cm, tid) + g_n;
    __syncthreads();
    s1 = s1n; s2 = s2n;
    lam_a = lam_an; lam_b = lam_bn; y_n = y_nn; g_n = g_nn;
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


// Adjoint sweep over SAVED forward intermediates (training-mode forward ran with a workspace):
// no recomputation - one reverse stage per Heun stage, walking the slots backwards.  The z/s/u
// vectors of the next slot are staged with cp.async into the other half of a double buffer and
// the per-lane o / u3 values are prefetched into registers while the current stage runs.
template <int C, bool D2>
__device__ __forceinline__ void fetch_saved(float* sb, const WsLayout& ws, const float* __restrict__ wsp,
                                            size_t slot) {
  using S = Smem<C>;
  constexpr int CW = C * W;
  const int tid = threadIdx.x;
  if (tid < H) {
    const int l = tid >> 5, c = tid & 31;
    const int dst = (l == 0) ? S::z1 : (l == 1) ? S::s1 : (l == 2) ? S::z2 : S::s2;
    cp_async4(sb + dst + c, wsp + ws.ZS + slot * 4 * W + tid);
  }
  if (D2 && tid < CW) {
    cp_async4(sb + S::u1 + tid, wsp + ws.US + slot * 2 * CW + tid);
    cp_async4(sb + S::u2 + tid, wsp + ws.US + slot * 2 * CW + CW + tid);
  }
  cp_async_commit();
}

template <int C, bool D2>
__global__ void __launch_bounds__(NT, 1) logncde_bwd_saved_kernel(const BwdParams p) {
  using S = Smem<C>;
  constexpr int CH = C * H;
  extern __shared__ __align__(16) float smem[];
  float* sw = smem;
  float* sbuf[2] = {smem + S::WEND_BWD, smem + S::WEND_BWD + S::STAGE};
  float* cm = sbuf[1] + S::STAGE;
  const int tid = threadIdx.x;
  Regs R;
  load_regs<C>(R, p.params, p.d);
  stage_weights_bwd<C>(p.d, p.params, sw);
  const int b = blockIdx.x;
  const float* ls = p.logsig + (size_t)b * p.K * p.D;
  const float* gys = p.g_ys + (size_t)b * (p.n_steps + 1) * H;
  const float* wsp = p.wsp;
  if (tid < 2 * W) cm[S::bacc + tid] = 0.f;
  if (tid < H) cm[S::ybar + tid] = gys[(size_t)p.n_steps * H + tid];
  float b3acc[2] = {0.f, 0.f};
  const int n_stage = 2 * p.n_steps;
  const size_t slot0 = (size_t)b * n_stage;
  const int row0 = lane_row0(tid);
  const bool rowt = tid < C * 64;
  // prologue: everything of the last stage
  int sidx = n_stage - 1;
  float lam_c = lambda_fetch<C>(ls + (size_t)p.stage_row[sidx] * p.D, D2);
  float sc_c = p.stage_scale[sidx];
  float2 o_c = make_float2(0.f, 0.f), u_c = o_c;
  if (rowt) {
    o_c = *reinterpret_cast<const float2*>(wsp + p.ws.Wt + (slot0 + sidx) * CH + row0);
    if (D2) u_c = *reinterpret_cast<const float2*>(wsp + p.ws.U3 + (slot0 + sidx) * CH + row0);
  }
  float g_c = tid < H ? gys[(size_t)(p.n_steps - 1) * H + tid] : 0.f;  // cotangent of y_n, n = sidx / 2
  fetch_saved<C, D2>(sbuf[sidx & 1], p.ws, wsp, slot0 + sidx);
  cp_async_wait_all();
  __syncthreads();
  for (; sidx >= 0; --sidx) {
    const int nx = sidx > 0 ? sidx - 1 : 0;
    // prefetch the next (earlier) stage
    const float lam_n = lambda_fetch<C>(ls + (size_t)p.stage_row[nx] * p.D, D2);
    const float sc_n = p.stage_scale[nx];
    float2 o_n = make_float2(0.f, 0.f), u_n = o_n;
    if (rowt) {
      o_n = *reinterpret_cast<const float2*>(wsp + p.ws.Wt + (slot0 + nx) * CH + row0);
      if (D2) u_n = *reinterpret_cast<const float2*>(wsp + p.ws.U3 + (slot0 + nx) * CH + row0);
    }
    const float g_n = tid < H ? gys[(size_t)(nx >> 1) * H + tid] : 0.f;
    fetch_saved<C, D2>(sbuf[nx & 1], p.ws, wsp, slot0 + nx);
    // cotangent of this stage's field value
    const bool second = sidx & 1;  // stage 2 of its Heun step
    const int pt_slot = pt_slot_of(n_stage - 1 - sidx);  // slots count stages in sweep order
    if (tid < H)
      cm[S::g + tid] = second ? 0.5f * sc_c * cm[S::ybar + tid] : sc_c * (0.5f * cm[S::ybar + tid] + cm[S::ab + tid]);
    lambda_store<C>(cm, lam_c);
    LNCDE_SYNC();
    const float oc[2] = {o_c.x, o_c.y}, uc[2] = {u_c.x, u_c.y};
    stage_reverse<C, D2, false>(sw, sbuf[sidx & 1], cm, R, oc, uc, b3acc, p.ws, p.wsp, slot0 + sidx, pt_slot);
    if (tid < H) {
      const float yo = yout_value<C>(cm, tid);
      if (second) cm[S::ab + tid] = yo;
      else cm[S::ybar + tid] += cm[S::ab + tid] + yo + g_c;
    }
    cp_async_wait_all();
    LNCDE_SYNC();
    lam_c = lam_n; sc_c = sc_n; o_c = o_n; u_c = u_n;
    if (!second) g_c = g_n;  // g_n was fetched for step (nx >> 1); it changes only when the step changes
  }
  if (tid < H) p.g_y0[(size_t)b * H + tid] = cm[S::ybar + tid];
  const int n_bias = 2 * W + C * H;
  float* bp = p.wsp + p.ws.bias_part + (size_t)b * n_bias;
  if (tid < 2 * W) bp[tid] = cm[S::bacc + tid];
  if (rowt) {
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
  auto go = [&](auto k) {
    cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    k<<<p.B, NT, smem, st>>>(p);
  };
  if (p.d.depth == 2) {
    if (p.save) go(logncde_fwd_fast_kernel<C, true, true>);
    else go(logncde_fwd_fast_kernel<C, true, false>);
  } else {
    if (p.save) go(logncde_fwd_fast_kernel<C, false, true>);
    else go(logncde_fwd_fast_kernel<C, false, false>);
  }
  return (int)cudaGetLastError();
}
template <int C>
static int launch_bwd_fast(cudaStream_t st, const BwdParams& p) {
  const size_t smem = bwd_smem_bytes<C>();
  auto go = [&](auto k) {
    cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    k<<<p.B, NT, smem, st>>>(p);
  };
  if (p.d.depth == 2) {
    if (p.saved) go(logncde_bwd_saved_kernel<C, true>);
    else go(logncde_bwd_fast_kernel<C, true>);
  } else {
    if (p.saved) go(logncde_bwd_saved_kernel<C, false>);
    else go(logncde_bwd_fast_kernel<C, false>);
  }
  return (int)cudaGetLastError();
}

inline bool shape_supported(int Hh, int C, int width, int n_hidden) {
  return Hh == H && width == W && n_hidden == 2 && (C == 4 || C == 6 || C == 7);
}

}  // namespace fast
}  // namespace lncde
