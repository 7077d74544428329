// Log-NCDE solve kernels for the EigenWorms model shape (H = 128, width 32, 2 hidden layers), third generation:
// SIX block barriers per Heun stage instead of thirteen.
//
// The kernels of logncde_fast.cuh spend a stage in ~12 phases of "block barrier + shared-memory round trip + short
// FMA chain" (clock64 harness, profiles/README.md: 240 - 850 cycles per phase, 5 200 per stage; ncu: issue slots 42 %
// busy, stalled_barrier 3.5 per issue).  Their register tiles and K-slicing are kept - what changes is WHO finishes a
// reduction and where its result lives, so that most hand-offs stay inside a warp:
//   * the K-slice partials of a hidden layer are summed by EVERY warp that needs the result (lane = row: 16 or 8
//     conflict-free loads), so z / a = SiLU(z) / s = SiLU'(z) live in registers of all warps - the "finish" phases
//     of a single warp and their barriers are gone;
//   * z2 = W2 a1 takes its operand from those registers by shuffles (warp q < 8 owns K-slice q: 4 SHFL + 4 FFMA);
//   * the antisymmetric Lambda-mix u1_j = sum_i Lam_ij q_i is applied to the K-slice PARTIALS of q_i = W1 o_i
//     (linear, so mixing commutes with the sum over slices): the partials go to shared memory already mixed and
//     tangent warp j reduces its 16 of them, multiplies by s1 (its own registers), carries the 32-vector through
//     W2 on its own (weights in per-thread order in shared memory) and multiplies by s2;
//   * the Heun update is done by the threads that need its result: warp w owns y[8w .. 8w+7] (the K-slice of the
//     first layer it multiplies with), sums the C field contributions of its 8 rows (4 lane groups x conflict-free
//     loads + 2 shuffles), keeps y_n / k1 in registers and hands y to its lanes by shuffles - no barrier between the
//     end of a stage and the first-layer partials of the next one.
// Phases of a stage (| = __syncthreads):  A update + z1 partials | B1 z1, a1, s1; z2 partials | B2 z2, a2, s2; last
// layer + tanh | C q_i partials, mixed | D tangent warps: u1, p1, u2, p2 | E last-layer tangents, field rows |
// Depth 1 has no tangents: A | B1 | B2.  Same arithmetic as logncde_fast.cuh up to the association of three sums
// (mix before slice-sum, field rows summed per block first): agreement to fp32 rounding, not bit for bit.
#pragma once
#include <cuda_runtime.h>

#include <cstdint>

#include "logncde_fast.cuh"

namespace lncde {
namespace v3 {

// -DLNCDE_PHASE_TIMING (dev tool, scripts/phase_timing.py): lane 0 of every warp of CTA 0 records the clock at numbered
// points of ONE stage (before / after every block barrier), so the critical path of each phase and the cost of the
// barriers themselves can be read off.  Compiled out of the product library.
#ifdef LNCDE_PHASE_TIMING
constexpr int kTrPoints = 32, kTrStage = 200;
__device__ long long g_v3_trace[2][kTrPoints][16];
#define V3_TRACE(kern, id)                                                               \
  do {                                                                                   \
    if (trace_on && (threadIdx.x & 31) == 0) g_v3_trace[kern][id][threadIdx.x >> 5] = clock64(); \
  } while (0)
#define V3_SYNC(kern, id)    \
  do {                       \
    V3_TRACE(kern, id);      \
    __syncthreads();         \
    V3_TRACE(kern, (id) + 1); \
  } while (0)
#else
#define V3_TRACE(kern, id) \
  do {                     \
  } while (0)
#define V3_SYNC(kern, id) __syncthreads()
#endif

using fast::H;
using fast::NT;
using fast::Regs;
using fast::W;

template <int C>
struct Sm {
  static constexpr int CH = C * H, CW = C * W;
  static constexpr int GP = H + 8;                 // row pitch of the field rows: 4 lane groups hit 4 x 8 distinct banks
  static constexpr int part1 = 0;                  // [16][32]   K-slice partials of z1
  static constexpr int part2 = part1 + 16 * W;     // [8][32]    K-slice partials of z2
  static constexpr int o = part2 + 8 * W;          // [C][H]     last-layer outputs
  static constexpr int gb = o + CH;                // [C][GP]    field rows  l_j o_j + (1 - o_j^2) u3_j
  static constexpr int up = gb + C * GP;           // [16][C][32] mixed K-slice partials of u1
  static constexpr int p2 = up + 16 * CW;          // [C][32]
  static constexpr int wslot = p2 + CW;            // [16][32]   per-warp broadcast slot (a2, then p1_j)
  static constexpr int lam = wslot + 16 * W;       // [8][8] Lambda, then [8] level-1 coefficients (fast::lambda_store)
  static constexpr int w2r = lam + 72;             // [8][32][4] W2 in per-thread order (fast::stage_w2r)
  static constexpr int bias = w2r + W * W;         // b1 [32] | b2 [32] | b3 [C*H]  (read once per stage: not worth registers)
  static constexpr int FWD_TOTAL = bias + 2 * W + CH;
  static_assert(o % 4 == 0 && gb % 4 == 0 && up % 4 == 0 && p2 % 4 == 0 && wslot % 4 == 0 && lam % 4 == 0 && w2r % 4 == 0 &&
                    bias % 4 == 0, "128-bit shared-memory accesses need 16-byte aligned regions");
};

__device__ __forceinline__ float dot8r(const float* w, const float (&x)[8]) {
  float s0 = w[0] * x[0];
  s0 = fmaf(w[1], x[1], s0);
  s0 = fmaf(w[2], x[2], s0);
  s0 = fmaf(w[3], x[3], s0);
  float s1 = w[4] * x[4];
  s1 = fmaf(w[5], x[5], s1);
  s1 = fmaf(w[6], x[6], s1);
  s1 = fmaf(w[7], x[7], s1);
  return s0 + s1;
}

// G[h], h = 8 warp + (lane & 7): lane group g = lane >> 3 sums field rows g, g + 4; the groups are combined by shuffles
// (every lane ends up with the value of its h)
template <int C>
__device__ __forceinline__ float field_sum(const float* sm, int warp, int lane) {
  using S = Sm<C>;
  const int grp = lane >> 3, h = warp * 8 + (lane & 7);
  float g = 0.f;
#pragma unroll
  for (int t0 = 0; t0 < C; t0 += 4) {
    const int t = t0 + grp;
    if (t < C) g += sm[S::gb + t * S::GP + h];
  }
  g += __shfl_xor_sync(0xffffffffu, g, 8);
  g += __shfl_xor_sync(0xffffffffu, g, 16);
  return g;
}

// The stage table (logsig row and dt / interval length per Heun stage) is read through a register window: lane l of
// every warp holds entry base + l, windows of 32 stages, the next window loaded when the current one is entered - a
// stage reads its entries with a shuffle.  A plain `p.stage_scale[sidx + 1]` prefetch does NOT work: the value is
// warp-uniform, ptxas moves it to a uniform register right behind the load (LDG -> R2UR) and every stage waits for
// two global-memory round trips at its top (~1 000 of ~5 500 cycles per stage, measured with the clock trace).
struct StageWindow {
  float sc;
  int row;
};
__device__ __forceinline__ StageWindow window_load(const float* __restrict__ scale, const int32_t* __restrict__ rows,
                                                   int n_stage, int base, int lane) {
  int e = base + lane;
  e = e < 0 ? 0 : (e >= n_stage ? n_stage - 1 : e);
  StageWindow w;
  w.sc = __ldg(scale + e);
  w.row = __ldg(rows + e);
  return w;
}

struct HeunState {
  float ycur, k1;  // y_n and the first Heun increment of the row h = 8 warp + (lane & 7)
};

// End of stage `done` (0-based): the stage's field value -> input of the next stage; writes y_{n+1} after second stages.
template <int C>
__device__ __forceinline__ float heun_advance(const float* sm, HeunState& hs, float sc_done, int done, float* __restrict__ ys,
                                              int warp, int lane) {
  const float gk = sc_done * field_sum<C>(sm, warp, lane);
  if (!(done & 1)) {  // first Heun stage: k1, evaluate next at y + k1
    hs.k1 = gk;
    return hs.ycur + gk;
  }
  const float yn = hs.ycur + 0.5f * (hs.k1 + gk);  // y_{n+1} = y_n + (k1 + k2) / 2
  hs.ycur = yn;
  if (lane < 8) ys[(size_t)((done >> 1) + 1) * H + warp * 8 + lane] = yn;
  return yn;
}

template <int C, bool D2, bool SAVE>
__global__ void __launch_bounds__(NT, 1) logncde_fwd_v3_kernel(const FwdParams p) {
  using S = Sm<C>;
  constexpr int CH = C * H, CW = C * W;
  extern __shared__ __align__(16) float smem[];
  float* sm = smem;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  Regs R;  // only the last-layer tile and the first-layer K-slice stay in registers (the loop runs at the 128-register
  fast::load_regs<C>(R, p.params, p.d);  // ceiling of a 512-thread CTA); b1, b2, b3 and the W2 K-slice are read from shared memory
  fast::stage_w2r(p.d, p.params, sm + S::w2r);
  for (int t = tid; t < 2 * W + CH; t += NT)
    sm[S::bias + t] = p.params[t < W ? p.d.b_off[0] + t : t < 2 * W ? p.d.b_off[1] + t - W : p.d.b_off[2] + t - 2 * W];
  const int b = blockIdx.x;
  const float* ls = p.logsig + (size_t)b * p.K * p.D;
  float* ys = p.ys + (size_t)b * (p.n_steps + 1) * H;
  HeunState hs;
  hs.ycur = p.y0[(size_t)b * H + warp * 8 + (lane & 7)];
  hs.k1 = 0.f;
  if (lane < 8) ys[warp * 8 + lane] = hs.ycur;
  const bool rowt = tid < C * 64;
  const int row0 = fast::lane_row0(tid);  // first of the two last-layer rows this lane finishes (row threads)
  const int jblk = tid >> 6, h0 = row0 & (H - 1);
  const int n_stage = 2 * p.n_steps;
  StageWindow w0 = window_load(p.stage_scale, p.stage_row, n_stage, 0, lane);
  StageWindow w1 = window_load(p.stage_scale, p.stage_row, n_stage, 32, lane);
  float sc_prev = 0.f;
  float lam_c = fast::lambda_fetch<C>(ls + (size_t)__shfl_sync(0xffffffffu, w0.row, 0) * p.D, D2);
  __syncthreads();  // w2r staged
  for (int sidx = 0; sidx < n_stage; ++sidx) {
    if ((sidx & 31) == 0 && sidx > 0) {  // entering a new window: the one after it is requested now, read in 32 stages
      w0 = w1;
      w1 = window_load(p.stage_scale, p.stage_row, n_stage, sidx + 32, lane);
    }
    const float sc = __shfl_sync(0xffffffffu, w0.sc, sidx & 31);
    const int e_n = (sidx + 1) & 31;  // logsig row of the NEXT stage, fetched one stage ahead
    const int row_a = __shfl_sync(0xffffffffu, w0.row, e_n), row_b = __shfl_sync(0xffffffffu, w1.row, 0);
    const float lam_n = fast::lambda_fetch<C>(ls + (size_t)(e_n ? row_a : row_b) * p.D, D2);
    const size_t slot = (size_t)b * n_stage + sidx;
#ifdef LNCDE_PHASE_TIMING
    const bool trace_on = blockIdx.x == 0 && sidx == kTrStage;
#endif
    V3_TRACE(0, 12);  // top of the stage
    // ---- A: input of this stage (Heun update of the previous one), first-layer K-slice partials
    const float yin = sidx == 0 ? hs.ycur : heun_advance<C>(sm, hs, sc_prev, sidx - 1, ys, warp, lane);
    V3_TRACE(0, 13);  // Heun update done
    if (SAVE && lane < 8) p.wsp[p.ws.Y + slot * H + warp * 8 + lane] = yin;
    {
      float y8[8];
#pragma unroll
      for (int e = 0; e < 8; ++e) y8[e] = __shfl_sync(0xffffffffu, yin, e);
      sm[S::part1 + warp * 32 + lane] = dot8r(R.w1, y8);
    }
    V3_SYNC(0, 0);
    // ---- B1: z1, a1, s1 in every warp (lane = row); K-slice partials of z2 from registers
    float s1v, s2v;
    {
      float acc[4] = {sm[S::bias + lane], 0.f, 0.f, 0.f};
#pragma unroll
      for (int q = 0; q < 16; ++q) acc[q & 3] += sm[S::part1 + q * 32 + lane];
      const float z = (acc[0] + acc[1]) + (acc[2] + acc[3]);
      const float sg = fast::sigm(z);
      const float a1 = z * sg;
      s1v = sg * (1.f + z * (1.f - sg));
      if (warp < 8) {
        const float4 w2 = *reinterpret_cast<const float4*>(sm + S::w2r + (warp * 32 + lane) * 4);  // W2[lane][4 warp ..]
        float t = w2.x * __shfl_sync(0xffffffffu, a1, warp * 4);
        t = fmaf(w2.y, __shfl_sync(0xffffffffu, a1, warp * 4 + 1), t);
        t = fmaf(w2.z, __shfl_sync(0xffffffffu, a1, warp * 4 + 2), t);
        t = fmaf(w2.w, __shfl_sync(0xffffffffu, a1, warp * 4 + 3), t);
        sm[S::part2 + warp * 32 + lane] = t;
      }
      if (SAVE && warp == 15) {
        p.wsp[p.ws.ZS + slot * 4 * W + lane] = z;
        p.wsp[p.ws.ZS + slot * 4 * W + W + lane] = s1v;
        p.wsp[p.ws.A[0] + slot * W + lane] = a1;
      }
    }
    fast::lambda_store<C>(sm + S::lam, lam_c);  // last readers of the previous row: phases C / E / A, all behind us
    V3_SYNC(0, 2);
    // ---- B2: z2, a2, s2 in every warp; last layer o = tanh(W3 a2 + b3) on the register tiles
    float o_r[2] = {0.f, 0.f};
    {
      float acc[4] = {sm[S::bias + W + lane], 0.f, 0.f, 0.f};
#pragma unroll
      for (int q = 0; q < 8; ++q) acc[q & 3] += sm[S::part2 + q * 32 + lane];
      const float z = (acc[0] + acc[1]) + (acc[2] + acc[3]);
      const float sg = fast::sigm(z);
      const float a2 = z * sg;
      s2v = sg * (1.f + z * (1.f - sg));
      sm[S::wslot + warp * 32 + lane] = a2;
      if (SAVE && warp == 15) {
        p.wsp[p.ws.ZS + slot * 4 * W + 2 * W + lane] = z;
        p.wsp[p.ws.ZS + slot * 4 * W + 3 * W + lane] = s2v;
        p.wsp[p.ws.A[1] + slot * W + lane] = a2;
      }
      __syncwarp();
      if (rowt) {
        float ra, rb;
        fast::tile_rows(R, sm + S::wslot + warp * 32 + (lane & 3) * 8, lane, ra, rb);
        const float2 b3 = *reinterpret_cast<const float2*>(sm + S::bias + 2 * W + row0);
        o_r[0] = tanhf(ra + b3.x);
        o_r[1] = tanhf(rb + b3.y);
        if (D2) {
          *reinterpret_cast<float2*>(sm + S::o + row0) = make_float2(o_r[0], o_r[1]);
        } else {
          const float l1 = sm[S::lam + 64 + jblk];
          *reinterpret_cast<float2*>(sm + S::gb + jblk * S::GP + h0) = make_float2(l1 * o_r[0], l1 * o_r[1]);
          if (SAVE) *reinterpret_cast<float2*>(p.wsp + p.ws.Wt + slot * CH + row0) = make_float2(o_r[0], o_r[1]);
        }
      }
    }
    V3_SYNC(0, 4);
    if (D2) {
      // ---- C: K-slice partials of q_i = W1 o_i, Lambda-mixed before they leave the thread
      {
        float q[C];
#pragma unroll
        for (int i = 0; i < C; ++i) {
          const float4 xa = *reinterpret_cast<const float4*>(sm + S::o + i * H + warp * 8);
          const float4 xb = *reinterpret_cast<const float4*>(sm + S::o + i * H + warp * 8 + 4);
          q[i] = fast::dot8(R.w1, xa, xb);
        }
        float u[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
#pragma unroll
        for (int i = 0; i < C; ++i) {
          const float4 la = *reinterpret_cast<const float4*>(sm + S::lam + i * 8);
          const float4 lb = *reinterpret_cast<const float4*>(sm + S::lam + i * 8 + 4);
          const float l8[8] = {la.x, la.y, la.z, la.w, lb.x, lb.y, lb.z, lb.w};
#pragma unroll
          for (int j = 0; j < C; ++j) u[j] = fmaf(l8[j], q[i], u[j]);
        }
#pragma unroll
        for (int j = 0; j < C; ++j) sm[S::up + (warp * C + j) * 32 + lane] = u[j];
      }
      V3_SYNC(0, 6);
      // ---- D: tangent warp j: u1_j, p1_j = s1 u1_j, u2_j = W2 p1_j, p2_j = s2 u2_j
      if (warp < C) {
        float acc[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
        for (int q = 0; q < 16; ++q) acc[q & 3] += sm[S::up + (q * C + warp) * 32 + lane];
        const float u1 = (acc[0] + acc[1]) + (acc[2] + acc[3]);
        const float p1 = s1v * u1;
        sm[S::wslot + warp * 32 + lane] = p1;
        __syncwarp();
        float a2c[2] = {0.f, 0.f};
        fast::warp_matvec32<2>(sm + S::w2r, sm + S::wslot + warp * 32, lane, a2c);
        const float u2 = a2c[0] + a2c[1];
        const float p2v = s2v * u2;
        sm[S::p2 + warp * W + lane] = p2v;
        if (SAVE) {
          p.wsp[p.ws.P[0] + slot * CW + tid] = p1;
          p.wsp[p.ws.P[1] + slot * CW + tid] = p2v;
          p.wsp[p.ws.US + slot * 2 * CW + tid] = u1;
          p.wsp[p.ws.US + slot * 2 * CW + CW + tid] = u2;
        }
      }
      V3_SYNC(0, 8);
      // ---- E: u3_j = W3[block j] p2_j; field rows l_j o_j + (1 - o_j^2) u3_j
      if (rowt) {
        float ua, ub;
        fast::tile_rows(R, sm + S::p2 + jblk * W + (lane & 3) * 8, lane, ua, ub);
        const float l1 = sm[S::lam + 64 + jblk];
        *reinterpret_cast<float2*>(sm + S::gb + jblk * S::GP + h0) =
            make_float2(fmaf(l1, o_r[0], (1.f - o_r[0] * o_r[0]) * ua), fmaf(l1, o_r[1], (1.f - o_r[1] * o_r[1]) * ub));
        if (SAVE) {
          *reinterpret_cast<float2*>(p.wsp + p.ws.Wt + slot * CH + row0) = make_float2(o_r[0], o_r[1]);
          *reinterpret_cast<float2*>(p.wsp + p.ws.U3 + slot * CH + row0) = make_float2(ua, ub);
        }
      }
      V3_SYNC(0, 10);
    }
    sc_prev = sc;
    lam_c = lam_n;
  }
  heun_advance<C>(sm, hs, sc_prev, n_stage - 1, ys, warp, lane);  // y after the last step
}

template <int C>
constexpr size_t fwd_smem_bytes() {
  return (size_t)Sm<C>::FWD_TOTAL * 4;
}

template <int C>
static int launch_fwd_v3(cudaStream_t st, const FwdParams& p) {
  const size_t smem = fwd_smem_bytes<C>();
  auto go = [&](auto k) {
    cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    k<<<p.B, NT, smem, st>>>(p);
  };
  if (p.d.depth == 2) {
    if (p.save) go(logncde_fwd_v3_kernel<C, true, true>);
    else go(logncde_fwd_v3_kernel<C, true, false>);
  } else {
    if (p.save) go(logncde_fwd_v3_kernel<C, false, true>);
    else go(logncde_fwd_v3_kernel<C, false, false>);
  }
  return (int)cudaGetLastError();
}

// =====================================================================================================================
// Adjoint sweep over the saved forward intermediates, six block barriers per reverse stage (thirteen in
// fast::logncde_bwd_saved_kernel).  Same ideas as the forward kernel, transposed:
//   * tangent warp j takes its cotangent through s2, W2^T and s1 alone (R2);
//   * the Lambda-mix of the first-layer cotangents is applied AFTER W1^T, on the K-slice partials
//     (o-bar_i += sum_j Lam_ij W1^T u1-bar_j), so it needs no phase of its own; v_i = sum_j Lam_ij u1-bar_j, the left
//     vector of the dW1 tangent pairs, is formed on the side by warp i (R3);
//   * z2-bar is finished by every warp that multiplies with it and handed over by shuffles (R5), likewise z1-bar (R6);
//   * the stage-input cotangent is summed by the row threads that start the next reverse stage: each keeps y-bar /
//     a-bar of ITS two rows in registers (replicated over the C blocks) - no barrier between two reverse stages.
// Phases (| = __syncthreads):  R1 field cotangents, W3^T | R2 tangent warps | R3 W1^T + mix | R4 z3-bar, W3^T |
// R5 z2-bar, W2^T partials | R6 z1-bar, W1^T partials |        depth 1: R4 | R5 | R6.
// Workspace slots, weight-gradient pairs and the deferred GEMMs are those of logncde.cu / logncde_fast.cuh.
template <int C>
struct SmB {
  static constexpr int CH = C * H, CW = C * W;
  static constexpr int W1T = 0;                    // [16][2][32][4] (fast::stage_weights_bwd)
  static constexpr int W2T = W1T + H * W;          // [8][32][4]
  static constexpr int lam = W2T + W * W;          // [8][8] + [8]
  static constexpr int red = lam + 72;             // [16][32]  W3^T partials per row warp
  static constexpr int dbuf = red + 16 * W;        // [C*H]     row values staged for tile_cols
  static constexpr int ub1 = dbuf + CH;            // [C][32]   u1-bar_j
  static constexpr int sp1 = ub1 + CW;             // [C][32]   u1_j * p1-bar_j  (s1-bar contributions)
  static constexpr int sp2 = sp1 + CW;             // [C][32]
  static constexpr int part = sp2 + CW;            // [4][8][H] mixed K-slice partials of o-bar
  static constexpr int part7 = part + 32 * H;      // [8][32]   K-slice partials of a1-bar
  static constexpr int ypart = part7 + 8 * W;      // [4][H]    K-slice partials of the stage-input cotangent
  static constexpr int wslot = ypart + 4 * H;      // [16][32]
  static constexpr int sv = wslot + 16 * W;        // 2 x [z1 | s1 | z2 | s2 | u1 (CW) | u2 (CW)]  saved stage data
  static constexpr int SV = 4 * W + 2 * CW;
  static constexpr int TOTAL = sv + 2 * SV;
  static_assert(W2T % 4 == 0 && lam % 4 == 0 && red % 4 == 0 && dbuf % 4 == 0 && ub1 % 4 == 0 && part % 4 == 0 &&
                    ypart % 4 == 0 && wslot % 4 == 0 && sv % 4 == 0 && SV % 4 == 0,
                "128-bit shared-memory accesses need 16-byte aligned regions");
};

// Lambda (sign applied) and the level-1 coefficients of one logsig row, stored by threads t0 .. t0 + 71
template <int C>
__device__ __forceinline__ float lambda_fetch_at(const float* __restrict__ row, bool d2, int t) {
  float v = 0.f;
  if (t >= 0 && t < 64) {
    const int i = t >> 3, j = t & 7;
    if (d2 && i < C && j < C && i != j) {
      const int a = i < j ? i : j, bb = i < j ? j : i;
      v = __ldg(row + 1 + C + a * C - (a * (a + 1)) / 2 + (bb - a - 1));
    }
  } else if (t >= 64 && t < 72) {
    if (t - 64 < C) v = __ldg(row + 1 + (t - 64));
  }
  return v;
}
__device__ __forceinline__ void lambda_store_at(float* lam, float v, int t) {
  if (t >= 0 && t < 72) lam[t] = (t < 64 && (t >> 3) > (t & 7)) ? -v : v;
}

__device__ __forceinline__ float silu_d2(float z) {  // SiLU''(z)
  const float sg = fast::sigm(z);
  return sg * (1.f - sg) * (2.f + z * (1.f - 2.f * sg));
}

template <int C, bool D2>
__device__ __forceinline__ void fetch_saved_v3(float* dst, const WsLayout& ws, const float* __restrict__ wsp, size_t slot) {
  constexpr int CW = C * W;
  const int tid = threadIdx.x;
  if (tid < 4 * W) cp_async4(dst + tid, wsp + ws.ZS + slot * 4 * W + tid);
  if (D2 && tid < CW) {
    cp_async4(dst + 4 * W + tid, wsp + ws.US + slot * 2 * CW + tid);
    cp_async4(dst + 4 * W + CW + tid, wsp + ws.US + slot * 2 * CW + CW + tid);
  }
  cp_async_commit();
}

template <int C, bool D2>
__global__ void __launch_bounds__(NT, 1) logncde_bwd_v3_kernel(const BwdParams p) {
  using S = SmB<C>;
  constexpr int CH = C * H, CW = C * W;
  extern __shared__ __align__(16) float smem[];
  float* sm = smem;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  Regs R;  // only R.Wt is used (the compiler drops the rest)
  fast::load_regs<C>(R, p.params, p.d);
  fast::stage_weights_bwd<C>(p.d, p.params, sm);  // W1T | W2T at offsets 0 / H * W, as fast::Smem
  const int b = blockIdx.x;
  const float* ls = p.logsig + (size_t)b * p.K * p.D;
  const float* gys = p.g_ys + (size_t)b * (p.n_steps + 1) * H;
  const float* wsp = p.wsp;
  float* wso = p.wsp;
  const WsLayout& ws = p.ws;
  const int n_stage = 2 * p.n_steps;
  const size_t slot0 = (size_t)b * n_stage;
  const bool rowt = tid < C * 64;
  const int row0 = fast::lane_row0(tid), jblk = tid >> 6, h0 = row0 & (H - 1);
  const int ks = warp >> 2, hq = warp & 3;  // K-slice / row quarter of the W1^T products
  const int lt = tid - 288;                 // threads 288 .. 359 (warps 9 - 11) stage Lambda of the next row
  float b3acc[2] = {0.f, 0.f}, bacc = 0.f;  // bias gradients: last layer (row threads), hidden layers (warp 8: b2, warp 15: b1)
  float2 ybar = make_float2(0.f, 0.f), ab = make_float2(0.f, 0.f);
  if (rowt) ybar = *reinterpret_cast<const float2*>(gys + (size_t)p.n_steps * H + h0);
  // prologue: everything of the last stage
  int sidx = n_stage - 1;
  StageWindow w0 = window_load(p.stage_scale, p.stage_row, n_stage, sidx & ~31, lane);        // window of the last stage
  StageWindow w1 = window_load(p.stage_scale, p.stage_row, n_stage, (sidx & ~31) - 32, lane);  // the one below it
  lambda_store_at(sm + S::lam, lambda_fetch_at<C>(ls + (size_t)__shfl_sync(0xffffffffu, w0.row, sidx & 31) * p.D, D2, lt), lt);
  float2 o_c = make_float2(0.f, 0.f), u_c = o_c, g_c = o_c;
  if (rowt) {
    o_c = *reinterpret_cast<const float2*>(wsp + ws.Wt + (slot0 + sidx) * CH + row0);
    if (D2) u_c = *reinterpret_cast<const float2*>(wsp + ws.U3 + (slot0 + sidx) * CH + row0);
    g_c = *reinterpret_cast<const float2*>(gys + (size_t)(p.n_steps - 1) * H + h0);  // cotangent of y_n, n = sidx / 2
  }
  fetch_saved_v3<C, D2>(sm + S::sv + (sidx & 1) * S::SV, ws, wsp, slot0 + sidx);
  cp_async_wait_all();
  __syncthreads();
  bool first = true;
  for (; sidx >= 0; --sidx) {
    const int nx = sidx > 0 ? sidx - 1 : 0;
    const size_t slot = slot0 + sidx;
    const bool second = sidx & 1;  // stage 2 of its Heun step
#ifdef LNCDE_PHASE_TIMING
    const bool trace_on = blockIdx.x == 0 && sidx == kTrStage;
#endif
    V3_TRACE(1, 12);  // top of the reverse stage
    const float* sv = sm + S::sv + (sidx & 1) * S::SV;
    // prefetch the next (earlier) stage
    if ((sidx & 31) == 31 && !first) {  // entering the window below: request the one below that (stage-table register window)
      w0 = w1;
      w1 = window_load(p.stage_scale, p.stage_row, n_stage, (sidx & ~31) - 32, lane);
    }
    const float sc_c = __shfl_sync(0xffffffffu, w0.sc, sidx & 31);
    const int row_a = __shfl_sync(0xffffffffu, w0.row, nx & 31), row_b = __shfl_sync(0xffffffffu, w1.row, 31);
    const float lam_n = lambda_fetch_at<C>(ls + (size_t)((nx >> 5) == (sidx >> 5) ? row_a : row_b) * p.D, D2, lt);
    float2 o_n = make_float2(0.f, 0.f), u_n = o_n, g_n = o_n;
    if (rowt) {
      o_n = *reinterpret_cast<const float2*>(wsp + ws.Wt + (slot0 + nx) * CH + row0);
      if (D2) u_n = *reinterpret_cast<const float2*>(wsp + ws.U3 + (slot0 + nx) * CH + row0);
      g_n = *reinterpret_cast<const float2*>(gys + (size_t)(nx >> 1) * H + h0);
    }
    fetch_saved_v3<C, D2>(sm + S::sv + (nx & 1) * S::SV, ws, wsp, slot0 + nx);
    // ---- RA + R1 (row threads): finish the previous reverse stage, cotangent of this stage's field value
    float2 ob = make_float2(0.f, 0.f), tp = ob;
    if (rowt) {
      if (!first) {
        float2 yo = make_float2(0.f, 0.f);
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          const float2 pv = *reinterpret_cast<const float2*>(sm + S::ypart + k * H + h0);
          yo.x += pv.x;
          yo.y += pv.y;
        }
        if (!second) {  // the previous reverse stage was a second Heun stage: its input cotangent is a-bar
          ab = yo;
        } else {        // it was a first stage: step n + 1 is done, y-bar_n complete
          ybar.x += ab.x + yo.x + g_c.x;
          ybar.y += ab.y + yo.y + g_c.y;
          g_c = g_n;  // now inside step sidx / 2: its y-cotangent is this iteration's prefetch (gys[(sidx - 1) >> 1])
        }
      }
      const float2 gv = second ? make_float2(0.5f * sc_c * ybar.x, 0.5f * sc_c * ybar.y)
                               : make_float2(sc_c * (0.5f * ybar.x + ab.x), sc_c * (0.5f * ybar.y + ab.y));
      V3_TRACE(1, 13);  // stage-input cotangent combined
      const float l1 = sm[S::lam + 64 + jblk];
      tp = make_float2(1.f - o_c.x * o_c.x, 1.f - o_c.y * o_c.y);
      ob = make_float2(l1 * gv.x, l1 * gv.y);
      if (D2) {
        const float2 ub = make_float2(tp.x * gv.x, tp.y * gv.y);
        ob.x = fmaf(-2.f * o_c.x, u_c.x * gv.x, ob.x);
        ob.y = fmaf(-2.f * o_c.y, u_c.y * gv.y, ob.y);
        *reinterpret_cast<float2*>(wso + ws.UL + slot * CH + row0) = ub;
        *reinterpret_cast<float2*>(sm + S::dbuf + row0) = ub;
        __syncwarp();
        V3_TRACE(1, 14);  // cotangents formed (o / u3 of this stage consumed)
        const float pr = fast::tile_cols(R, sm + S::dbuf + (warp * 8 + (lane >> 2)) * 8, lane);
        sm[S::red + warp * W + fast::lane_col(lane)] = pr;
      }
    }
    first = false;
    if (D2) {
      V3_SYNC(1, 0);
      // ---- R2 (tangent warp j): p2-bar_j -> u2-bar_j -> p1-bar_j = W2^T u2-bar_j -> u1-bar_j; s-bar contributions
      if (warp < C) {
        const float pb = sm[S::red + (2 * warp) * W + lane] + sm[S::red + (2 * warp + 1) * W + lane];
        const float ub2 = sv[3 * W + lane] * pb;
        sm[S::sp2 + tid] = sv[4 * W + CW + tid] * pb;
        wso[ws.U[1] + slot * CW + tid] = ub2;
        sm[S::wslot + tid] = ub2;
        __syncwarp();
        float acc[2] = {0.f, 0.f};
        fast::warp_matvec32<2>(sm + S::W2T, sm + S::wslot + warp * W, lane, acc);
        const float pb1 = acc[0] + acc[1];
        sm[S::ub1 + tid] = sv[W + lane] * pb1;
        sm[S::sp1 + tid] = sv[4 * W + tid] * pb1;
      }
      V3_SYNC(1, 2);
      // ---- R3 (all warps): K-slice partials of W1^T u1-bar_j, Lambda-mixed to o-bar_i partials; v_i on the side
      {
        const float4 wa = *reinterpret_cast<const float4*>(sm + S::W1T + ((warp * 2 + 0) * 32 + lane) * 4);
        const float4 wb = *reinterpret_cast<const float4*>(sm + S::W1T + ((warp * 2 + 1) * 32 + lane) * 4);
        const float w8[8] = {wa.x, wa.y, wa.z, wa.w, wb.x, wb.y, wb.z, wb.w};
        float t[C];
#pragma unroll
        for (int j = 0; j < C; ++j) {
          const float4 xa = *reinterpret_cast<const float4*>(sm + S::ub1 + j * W + ks * 8);
          const float4 xb = *reinterpret_cast<const float4*>(sm + S::ub1 + j * W + ks * 8 + 4);
          t[j] = fast::dot8(w8, xa, xb);
        }
#pragma unroll
        for (int i = 0; i < C; ++i) {
          const float4 la = *reinterpret_cast<const float4*>(sm + S::lam + i * 8);
          const float4 lb = *reinterpret_cast<const float4*>(sm + S::lam + i * 8 + 4);
          const float l8[8] = {la.x, la.y, la.z, la.w, lb.x, lb.y, lb.z, lb.w};
          float a = 0.f;
#pragma unroll
          for (int j = 0; j < C; ++j) a = fmaf(l8[j], t[j], a);
          sm[S::part + (ks * 8 + i) * H + hq * 32 + lane] = a;
        }
        if (warp < C) {
          float v = 0.f;
#pragma unroll
          for (int j = 0; j < C; ++j) v = fmaf(sm[S::lam + warp * 8 + j], sm[S::ub1 + j * W + lane], v);
          wso[ws.U[0] + slot * CW + tid] = v;
        }
      }
      V3_SYNC(1, 4);
    }
    // ---- R4 (row threads): o-bar complete, z3-bar = (1 - o^2) o-bar, a2-bar partials = W3^T z3-bar
    if (rowt) {
      if (D2) {
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          const float2 pv = *reinterpret_cast<const float2*>(sm + S::part + (k * 8 + jblk) * H + h0);
          ob.x += pv.x;
          ob.y += pv.y;
        }
      }
      const float2 zb = make_float2(tp.x * ob.x, tp.y * ob.y);
      b3acc[0] += zb.x;
      b3acc[1] += zb.y;
      *reinterpret_cast<float2*>(wso + ws.ZL + slot * CH + row0) = zb;
      *reinterpret_cast<float2*>(sm + S::dbuf + row0) = zb;
      __syncwarp();
      const float pr = fast::tile_cols(R, sm + S::dbuf + (warp * 8 + (lane >> 2)) * 8, lane);
      sm[S::red + warp * W + fast::lane_col(lane)] = pr;
    }
    V3_SYNC(1, 6);
    // ---- R5: z2-bar in warps 0 - 8, K-slice partials of a1-bar = W2^T z2-bar from registers; Lambda of the next row
    if (warp <= 8) {
      float acc[2] = {0.f, 0.f};
#pragma unroll
      for (int wq = 0; wq < 2 * C; ++wq) acc[wq & 1] += sm[S::red + wq * W + lane];
      float zb2 = sv[3 * W + lane] * (acc[0] + acc[1]);
      if (D2) {
        float s2b = 0.f;
#pragma unroll
        for (int j = 0; j < C; ++j) s2b += sm[S::sp2 + j * W + lane];
        zb2 = fmaf(silu_d2(sv[2 * W + lane]), s2b, zb2);
      }
      if (warp < 8) {
        const float4 wt = *reinterpret_cast<const float4*>(sm + S::W2T + (warp * 32 + lane) * 4);  // W2[4 warp + e][lane]
        float a = wt.x * __shfl_sync(0xffffffffu, zb2, warp * 4);
        a = fmaf(wt.y, __shfl_sync(0xffffffffu, zb2, warp * 4 + 1), a);
        a = fmaf(wt.z, __shfl_sync(0xffffffffu, zb2, warp * 4 + 2), a);
        a = fmaf(wt.w, __shfl_sync(0xffffffffu, zb2, warp * 4 + 3), a);
        sm[S::part7 + warp * W + lane] = a;
      } else {
        bacc += zb2;
        wso[ws.Z[1] + slot * W + lane] = zb2;
      }
    }
    if (sidx > 0) lambda_store_at(sm + S::lam, lam_n, lt);  // this row's Lambda was last read in R3, its lam1 in R1
    V3_SYNC(1, 8);
    // ---- R6 (all warps): z1-bar, K-slice partials of the stage-input cotangent W1^T z1-bar
    {
      float acc[2] = {0.f, 0.f};
#pragma unroll
      for (int q = 0; q < 8; ++q) acc[q & 1] += sm[S::part7 + q * W + lane];
      float zb1 = sv[W + lane] * (acc[0] + acc[1]);
      if (D2) {
        float s1b = 0.f;
#pragma unroll
        for (int j = 0; j < C; ++j) s1b += sm[S::sp1 + j * W + lane];
        zb1 = fmaf(silu_d2(sv[lane]), s1b, zb1);
      }
      if (warp == 15) {
        bacc += zb1;
        wso[ws.Z[0] + slot * W + lane] = zb1;
      }
      const float4 wa = *reinterpret_cast<const float4*>(sm + S::W1T + ((warp * 2 + 0) * 32 + lane) * 4);
      const float4 wb = *reinterpret_cast<const float4*>(sm + S::W1T + ((warp * 2 + 1) * 32 + lane) * 4);
      float a = wa.x * __shfl_sync(0xffffffffu, zb1, ks * 8);
      a = fmaf(wa.y, __shfl_sync(0xffffffffu, zb1, ks * 8 + 1), a);
      a = fmaf(wa.z, __shfl_sync(0xffffffffu, zb1, ks * 8 + 2), a);
      a = fmaf(wa.w, __shfl_sync(0xffffffffu, zb1, ks * 8 + 3), a);
      float a2 = wb.x * __shfl_sync(0xffffffffu, zb1, ks * 8 + 4);
      a2 = fmaf(wb.y, __shfl_sync(0xffffffffu, zb1, ks * 8 + 5), a2);
      a2 = fmaf(wb.z, __shfl_sync(0xffffffffu, zb1, ks * 8 + 6), a2);
      a2 = fmaf(wb.w, __shfl_sync(0xffffffffu, zb1, ks * 8 + 7), a2);
      sm[S::ypart + ks * H + hq * 32 + lane] = a + a2;
    }
    cp_async_wait_all();
    V3_SYNC(1, 10);
    o_c = o_n;
    u_c = u_n;
  }
  // the first stage of step 0 has just been swept: y-bar_0 = y-bar + a-bar + its input cotangent + g_ys[0]
  if (rowt) {
    float2 yo = make_float2(0.f, 0.f);
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const float2 pv = *reinterpret_cast<const float2*>(sm + S::ypart + k * H + h0);
      yo.x += pv.x;
      yo.y += pv.y;
    }
    if (tid < 64)
      *reinterpret_cast<float2*>(p.g_y0 + (size_t)b * H + h0) =
          make_float2(ybar.x + ab.x + yo.x + g_c.x, ybar.y + ab.y + yo.y + g_c.y);
  }
  const int n_bias = 2 * W + CH;
  float* bp = p.wsp + ws.bias_part + (size_t)b * n_bias;
  if (warp == 15) bp[lane] = bacc;
  if (warp == 8) bp[W + lane] = bacc;
  if (rowt) {
    bp[2 * W + row0] = b3acc[0];
    bp[2 * W + row0 + 1] = b3acc[1];
  }
}

template <int C>
constexpr size_t bwd_smem_bytes() {
  return (size_t)SmB<C>::TOTAL * 4;
}

template <int C>
static int launch_bwd_v3(cudaStream_t st, const BwdParams& p) {
  const size_t smem = bwd_smem_bytes<C>();
  auto go = [&](auto k) {
    cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    k<<<p.B, NT, smem, st>>>(p);
  };
  if (p.d.depth == 2) go(logncde_bwd_v3_kernel<C, true>);
  else go(logncde_bwd_v3_kernel<C, false>);
  return (int)cudaGetLastError();
}

}  // namespace v3
}  // namespace lncde
