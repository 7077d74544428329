// K2, dense blocks: the two sequential sweeps of the DE-LNCDE scan spread over a thread-block CLUSTER per series.
//
// Problem: the single-CTA sweeps (scan_fwd_tma_kernel, scan_bwd_tma_lam_kernel) stream one series' flows through ONE SM.
// EigenWorms DE has 32 series of 1499 flows of 64 KB: 32 of 148 SMs pull 64 GB/s each and HBM idles at 25-30 %.
// Here CS CTAs of a cluster share a series.  CTA `rank` owns the rows [rank * H/CS, (rank+1) * H/CS) of every flow (a
// contiguous slice of the row-major flow, fetched by 1-D TMA bulk copies through a ring of NS stages) and the same
// entries of the state.  What crosses CTAs per step is only the H-vector:
//   forward   y_t[h] = sum_j F_t[h][j] y_{t-1}[j]: the owner of row h sends y_t[h] to every CTA;
//   reverse   lam_{t-1}[j] = g_{t-1}[j] + sum_h F_t[h][j] lam_t[h]: every CTA sends its partial column sums to the
//             owner of column j, which adds them.
// The exchange is one-sided: st.async into the peer's shared memory, completion counted in bytes on the peer's mbarrier
// (the same mechanism a TMA load uses), so a step costs one DSMEM store latency and no cluster barrier.  The receive
// buffers are double buffered by step parity; a CTA can only run one step ahead of its peers because every step needs
// every CTA's contribution, which is also what makes the stage ring and the buffers safe to reuse without a block barrier.
//
// Reference: the recurrence is the scan of /root/reference/models/LinearNeuralCDEs.py:120-160 (y_{i+1} = flow_i y_i).
#pragma once
#include <cuda_runtime.h>

#include <map>
#include <mutex>

#include "ptx.cuh"

namespace lncde {

constexpr int kClusterThreads = 256;

struct ClusterShape {
  int cs = 0;      // CTAs per series (0 = shape not handled)
  int ns = 0;      // ring stages
  size_t smem = 0;
};

// rows of one CTA: RH = H / CS, a multiple of 4 that divides the block size (a slice never straddles two blocks)
__host__ __device__ inline bool cluster_rows_ok(int H, int bs, int cs) {
  if (H % (4 * cs) != 0) return false;
  const int rh = H / cs;
  return rh <= bs && bs % rh == 0 && kClusterThreads % rh == 0;
}

// HH / BS: the state and block size at compile time (0 = take the run-time arguments): with the EigenWorms sizes fixed
// the row and column loops have one trip and the step is a straight line of ~70 instructions per warp.
template <int CS, int HH, int BS>
__global__ void __launch_bounds__(kClusterThreads, 1)
    scan_fwd_cluster_kernel(const float* __restrict__ F, const float* __restrict__ y0, float* __restrict__ ys, int T,
                            int H_rt, int bs_rt, int NS) {
  extern __shared__ __align__(128) unsigned char smraw[];
  const int H = HH ? HH : H_rt, bs = BS ? BS : bs_rt;
  const int RH = H / CS, SN = RH * bs;
  const unsigned slice_bytes = (unsigned)SN * 4, y_bytes = (unsigned)H * 4;
  float* tiles = reinterpret_cast<float*>(smraw);  // [NS][SN]
  float* ybuf = tiles + (size_t)NS * SN;           // [2][H]  y_t lives in half t & 1
  unsigned long long* fbar = reinterpret_cast<unsigned long long*>(ybuf + 2 * H);  // [NS] flow stages
  unsigned long long* ybar = fbar + NS;                                            // [2]  state halves
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nw = kClusterThreads / 32;
  const int rank = (int)cluster_ctarank(), b = blockIdx.x / CS;
  const int N = H * bs;
  const float* Fb = F + (size_t)b * T * N + (size_t)rank * SN;
  float* yb = ys + (size_t)b * T * H;
  if (tid == 0) {
    for (int s = 0; s < NS; ++s) mbar_init(fbar + s, 1);
    mbar_init(ybar + 0, 1);
    mbar_init(ybar + 1, 1);
    fence_mbar_init();
    // y_1 arrives in half 1, y_2 in half 0: both halves start armed (y_0 is loaded locally)
    if (T > 1) mbar_expect_tx(ybar + 1, y_bytes);
    if (T > 2) mbar_expect_tx(ybar + 0, y_bytes);
    for (int u = 0; u < NS && 1 + u < T; ++u) {
      mbar_expect_tx(fbar + u, slice_bytes);
      tma_load_1d(tiles + (size_t)u * SN, Fb + (size_t)(1 + u) * N, slice_bytes, fbar + u);
    }
  }
  for (int h = tid; h < H; h += kClusterThreads) {
    const float v = y0[(size_t)b * H + h];
    ybuf[h] = v;
    if (h / RH == rank) yb[h] = v;
  }
  __syncthreads();
  cluster_sync_all();  // every peer's barriers exist before the first remote store
  // the slice lies inside one block (RH | bs): its state columns start here
  const int ycol0 = ((rank * RH) / bs) * bs;
  int s = 0, sprev = NS - 1;      // stage of this step / of the previous one
  unsigned fphase = 0;            // parity of the current pass over the ring
  for (int t = 1; t < T; ++t) {
    if (t >= 2) {
      // y_{t-1} complete here <=> every warp of every CTA has finished step t-1.  (Plain wait: the payload is shared
      // memory written through the async proxy and counted on this barrier, as for a TMA load; an acquire at cluster
      // scope would add an L1 invalidate per step.)
      mbar_wait(ybar + ((t - 1) & 1), (unsigned)(((t - 2) >> 1) & 1));
    }
    mbar_wait(fbar + s, fphase);
    const float* Ft = tiles + (size_t)s * SN;
    const float* yv = ybuf + ((t - 1) & 1) * H + ycol0;
    float* yn = ybuf + (t & 1) * H;
    // warp per group of 4 rows, lanes across the columns
    for (int r0 = warp * 4; r0 < RH; r0 += nw * 4) {
      const int h0 = rank * RH + r0;
      float v[4] = {0.f, 0.f, 0.f, 0.f};
      for (int j4 = lane; j4 < bs / 4; j4 += 32) {
        const float4 y4 = *reinterpret_cast<const float4*>(yv + j4 * 4);
#pragma unroll
        for (int r = 0; r < 4; ++r) {
          const float4 f4 = *reinterpret_cast<const float4*>(Ft + (size_t)(r0 + r) * bs + j4 * 4);
          v[r] = fmaf(f4.x, y4.x, v[r]);
          v[r] = fmaf(f4.y, y4.y, v[r]);
          v[r] = fmaf(f4.z, y4.z, v[r]);
          v[r] = fmaf(f4.w, y4.w, v[r]);
        }
      }
      // transpose-reduce 4 rows over 32 lanes: afterwards all 8 lanes of group g = lane >> 3 hold row 2 (g & 1) + (g >> 1)
#pragma unroll
      for (int k = 0; k < 2; ++k) {
        const bool up = lane & 16;
        const float send = up ? v[k] : v[k + 2], keep = up ? v[k + 2] : v[k];
        v[k] = keep + __shfl_xor_sync(0xffffffffu, send, 16);
      }
      {
        const bool up = lane & 8;
        const float send = up ? v[0] : v[1], keep = up ? v[1] : v[0];
        v[0] = keep + __shfl_xor_sync(0xffffffffu, send, 8);
      }
      v[0] += __shfl_xor_sync(0xffffffffu, v[0], 4);
      v[0] += __shfl_xor_sync(0xffffffffu, v[0], 2);
      v[0] += __shfl_xor_sync(0xffffffffu, v[0], 1);
      const int h = h0 + ((lane >> 4) & 1) * 2 + ((lane >> 3) & 1), sub = lane & 7;
      // lane `sub` of the group delivers the value to CTA `sub` (its own CTA included); 4 bytes counted per delivery
      for (int dst = sub; dst < CS; dst += 8) st_async_f32(yn + h, v[0], ybar + (t & 1), (unsigned)dst);
      if (sub == 7) yb[(size_t)t * H + h] = v[0];
    }
    // housekeeping AFTER the send - every peer waits on this warp's rows, nobody on this.  y_{t-1} was complete at the
    // top of the step, so the stage of step t-1 is free (refilled NS steps ahead) and the half of y_{t+1} can be armed
    // (bytes that arrive before the arming are counted all the same).
    if (tid == 0 && t >= 2) {
      if (t + 1 < T) mbar_expect_tx(ybar + ((t + 1) & 1), y_bytes);
      const int tn = t - 1 + NS;
      if (tn < T) {
        fence_async_proxy();
        mbar_expect_tx(fbar + sprev, slice_bytes);
        tma_load_1d(tiles + (size_t)sprev * SN, Fb + (size_t)tn * N, slice_bytes, fbar + sprev);
      }
    }
    sprev = s;
    if (++s == NS) {
      s = 0;
      fphase ^= 1u;
    }
  }
  // nothing may be in flight towards a CTA that has exited: take delivery of the last state, then leave together
  if (T > 1) mbar_wait(ybar + ((T - 1) & 1), (unsigned)(((T - 2) >> 1) & 1));
  cluster_sync_all();
}

// lam_{T-1} = g_{T-1};  for t = T-1 .. 1:  lam_out[b,t,:] = lam_t;  lam_{t-1} = F_t^T lam_t + g_{t-1};  g_y0 = lam_0
// (same contract as scan_bwd_tma_lam_kernel).  CTA `rank` keeps lam[rank * RH .. +RH) and forms, for each of the bs
// columns of its block, the partial sum over its RH rows - split over RQ = 256 / bs thread groups - and sends it to
// the owner of the column; an owner receives 256 / RH partials per column.
template <int CS, int HH, int BS>
__global__ void __launch_bounds__(kClusterThreads, 1)
    scan_lam_cluster_kernel(const float* __restrict__ F, const float* __restrict__ g_ys, float* __restrict__ lam_out,
                            float* __restrict__ g_y0, int T, int H_rt, int bs_rt, int NS) {
  extern __shared__ __align__(128) unsigned char smraw[];
  const int H = HH ? HH : H_rt, bs = BS ? BS : bs_rt;
  const int RH = H / CS, SN = RH * bs, NP = kClusterThreads / RH;  // NP partials per owned column
  const unsigned slice_bytes = (unsigned)SN * 4, rx_bytes = (unsigned)kClusterThreads * 4;
  float* tiles = reinterpret_cast<float*>(smraw);  // [NS][SN]
  float* lamb = tiles + (size_t)NS * SN;           // [2][RH]
  float* rx = lamb + 2 * RH;                          // [2][NP][RH]  partials of step t in half t & 1
  unsigned long long* fbar = reinterpret_cast<unsigned long long*>(rx + 2 * kClusterThreads);  // [NS]
  unsigned long long* rbar = fbar + NS;                                                        // [2]
  const int tid = threadIdx.x;
  const int rank = (int)cluster_ctarank(), b = blockIdx.x / CS;
  const int N = H * bs;
  const float* Fb = F + (size_t)b * T * N + (size_t)rank * SN;
  const float* gb = g_ys + (size_t)b * T * H + rank * RH;
  float* lo = lam_out + (size_t)b * T * H + rank * RH;
  if (tid == 0) {
    for (int s = 0; s < NS; ++s) mbar_init(fbar + s, 1);
    mbar_init(rbar + 0, 1);
    mbar_init(rbar + 1, 1);
    fence_mbar_init();
    if (T > 1) mbar_expect_tx(rbar + ((T - 1) & 1), rx_bytes);
    if (T > 2) mbar_expect_tx(rbar + ((T - 2) & 1), rx_bytes);
    for (int u = 0; u < NS && T - 1 - u >= 1; ++u) {
      mbar_expect_tx(fbar + u, slice_bytes);
      tma_load_1d(tiles + (size_t)u * SN, Fb + (size_t)(T - 1 - u) * N, slice_bytes, fbar + u);
    }
  }
  if (tid < RH) lamb[tid] = gb[(size_t)(T - 1) * H + tid];
  __syncthreads();
  cluster_sync_all();
  // thread (jj, rq): column jj of the block this slice lies in, rows rq, rq + RQ, ...
  const int RQ = kClusterThreads / bs, jj = tid % bs, rq = tid / bs;
  const int row0 = rank * RH, blk = row0 / bs;
  const int col = blk * bs + jj;                            // global column
  const int owner = col / RH, oc = col - owner * RH;        // owning CTA, its local index
  const int src = (rank - blk * (bs / RH)) * RQ + rq;       // which of the owner's NP partials this is
  int s = 0, sprev = 0, cur = 0;   // ring stage of this step / the previous one, half of lamb that holds lam_t
  unsigned fphase = 0;
  float g_ahead = (tid < RH && T >= 2) ? gb[(size_t)(T - 2) * H + tid] : 0.f;
  for (int t = T - 1; t >= 1; --t) {
    const int it = T - 1 - t;
    const float* lam = lamb + cur * RH;
    // g_{t-1} is needed after the exchange: loaded one step ahead (its L2 latency is about a whole step)
    const float g_prev = g_ahead;
    g_ahead = (tid < RH && t >= 2) ? gb[(size_t)(t - 2) * H + tid] : 0.f;
    if (tid < RH) lo[(size_t)t * H + tid] = lam[tid];
    mbar_wait(fbar + s, fphase);
    const float* Ft = tiles + (size_t)s * SN + jj;
    float acc = 0.f;
#pragma unroll 16
    for (int i = rq; i < RH; i += RQ) acc = fmaf(Ft[(size_t)i * bs], lam[i], acc);
    st_async_f32(rx + ((t & 1) * NP + src) * RH + oc, acc, rbar + (t & 1), (unsigned)owner);
    if (tid == 0 && t + 1 - NS >= 1 && t + 1 <= T - 1) {
      // housekeeping after the send: the stage of step t+1 was released by the block barrier that closed it
      fence_async_proxy();
      mbar_expect_tx(fbar + sprev, slice_bytes);
      tma_load_1d(tiles + (size_t)sprev * SN, Fb + (size_t)(t + 1 - NS) * N, slice_bytes, fbar + sprev);
    }
    // all partials of step t are here <=> every thread of every CTA sharing this block has read lam_t and its stage
    mbar_wait(rbar + (t & 1), (unsigned)((it >> 1) & 1));
    if (tid < RH) {
      float sum = g_prev;
      for (int q = 0; q < NP; ++q) sum += rx[((t & 1) * NP + q) * RH + tid];
      lamb[(cur ^ 1) * RH + tid] = sum;  // the other half: lam_t may still be read by a slower warp's `lo` store
    }
    if (tid == 0 && t - 2 >= 1) mbar_expect_tx(rbar + (t & 1), rx_bytes);
    __syncthreads();  // lam_{t-1} visible; every thread of THIS CTA is past its reads of the stage
    cur ^= 1;
    sprev = s;
    if (++s == NS) {
      s = 0;
      fphase ^= 1u;
    }
  }
  const float* lamb_end = lamb + cur * RH;
  if (tid < RH) g_y0[(size_t)b * H + rank * RH + tid] = lamb_end[tid];
  cluster_sync_all();
}

inline size_t cluster_smem_bytes(bool bwd, int H, int bs, int cs, int ns) {
  const size_t slice = (size_t)(H / cs) * bs * 4;
  const size_t vec = bwd ? (size_t)(2 * (H / cs) + 2 * kClusterThreads) * 4 : (size_t)2 * H * 4;
  return ns * slice + vec + (size_t)(ns + 2) * 8 + 64;
}

// a refused attribute / launch is not an error of the call: the one-CTA sweep takes over (clear the runtime's last error)
static inline int cluster_fail() {
  (void)cudaGetLastError();
  return 0;
}

template <int CS, int HH, int BS>
static int launch_cluster_scan(bool bwd, cudaStream_t cs, const float* F, const float* a, float* out, float* g_y0, int B,
                               int T, int H, int bs, int ns, size_t smem, bool probe_only) {
#if LNCDE_EMU
  // the host emulator runs CTAs one at a time: it executes the kernels as clusters of one CTA (CS = 1)
  if (probe_only) return 1 << 30;
  if (!bwd) {
    cudaFuncSetAttribute(scan_fwd_cluster_kernel<CS, HH, BS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    scan_fwd_cluster_kernel<CS, HH, BS><<<B * CS, kClusterThreads, smem, cs>>>(F, a, out, T, H, bs, ns);
  } else {
    cudaFuncSetAttribute(scan_lam_cluster_kernel<CS, HH, BS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    scan_lam_cluster_kernel<CS, HH, BS><<<B * CS, kClusterThreads, smem, cs>>>(F, a, out, g_y0, T, H, bs, ns);
  }
  return 1 << 30;
#else
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3((unsigned)(B * CS));
  cfg.blockDim = dim3(kClusterThreads);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = cs;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = CS;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  int resident = 0;
  if (!bwd) {
    auto k = scan_fwd_cluster_kernel<CS, HH, BS>;
    if (cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess) return cluster_fail();
    if (CS > 8 && cudaFuncSetAttribute(k, cudaFuncAttributeNonPortableClusterSizeAllowed, 1) != cudaSuccess) return cluster_fail();
    if (cudaOccupancyMaxActiveClusters(&resident, k, &cfg) != cudaSuccess) return cluster_fail();
    if (!probe_only && cudaLaunchKernelEx(&cfg, k, F, a, out, T, H, bs, ns) != cudaSuccess) return cluster_fail();
  } else {
    auto k = scan_lam_cluster_kernel<CS, HH, BS>;
    if (cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess) return cluster_fail();
    if (CS > 8 && cudaFuncSetAttribute(k, cudaFuncAttributeNonPortableClusterSizeAllowed, 1) != cudaSuccess) return cluster_fail();
    if (cudaOccupancyMaxActiveClusters(&resident, k, &cfg) != cudaSuccess) return cluster_fail();
    if (!probe_only && cudaLaunchKernelEx(&cfg, k, F, a, out, g_y0, T, H, bs, ns) != cudaSuccess) return cluster_fail();
  }
  return resident;
#endif
}

// Shape of the sweep with CS CTAs per series (cs = 0: not handled)
inline ClusterShape cluster_shape(bool bwd, int H, int bs, int cs) {
  ClusterShape r;
  if (bs % 4 != 0 || bs > kClusterThreads || kClusterThreads % bs != 0 || !cluster_rows_ok(H, bs, cs)) return r;
  const size_t slice = (size_t)(H / cs) * bs * 4;
  int ns = (int)((200 * 1024) / slice);
  if (ns > 8) ns = 8;
  if (ns < 2) return r;
  r.cs = cs;
  r.ns = ns;
  r.smem = cluster_smem_bytes(bwd, H, bs, cs, ns);
  return r;
}

template <int CS>
static int cluster_try(bool bwd, cudaStream_t cs, const float* F, const float* a, float* out, float* g_y0, int B, int T,
                       int H, int bs) {
  const ClusterShape sh = cluster_shape(bwd, H, bs, CS);
  if (sh.cs == 0) return 0;
  // co-resident clusters of this (kernel, shared-memory size): asked once, the answer is a property of the device
  static std::mutex mu;
  static std::map<size_t, int> seen[2];
  int resident;
  {
    std::lock_guard<std::mutex> lock(mu);
    auto it = seen[bwd].find(sh.smem);
    if (it == seen[bwd].end()) {
      resident = launch_cluster_scan<CS, 0, 0>(bwd, cs, F, a, out, g_y0, B, T, H, bs, sh.ns, sh.smem, true);
      seen[bwd][sh.smem] = resident;
    } else {
      resident = it->second;
    }
  }
  if (resident < B) return 0;  // a second wave would double the sweep: a narrower cluster (or one CTA) is better
  if (H == 128 && bs == 128)  // the EigenWorms dense model: sizes folded at compile time
    return launch_cluster_scan<CS, 128, 128>(bwd, cs, F, a, out, g_y0, B, T, H, bs, sh.ns, sh.smem, false);
  return launch_cluster_scan<CS, 0, 0>(bwd, cs, F, a, out, g_y0, B, T, H, bs, sh.ns, sh.smem, false);
}

// true if the sweep was launched on clusters: the widest of 8 / 4 / 2 CTAs per series whose clusters are all resident
// at once; with 148 series or more one CTA per series already covers the SMs.
static inline bool scan_cluster_dispatch(bool bwd, cudaStream_t cs, const float* F, const float* a, float* out,
                                         float* g_y0, int B, int T, int H, int bs) {
#if LNCDE_EMU
  return cluster_try<1>(bwd, cs, F, a, out, g_y0, B, T, H, bs) > 0;
#else
  if (B * 2 > 148) return false;
  if (B * 8 <= 148 && cluster_try<8>(bwd, cs, F, a, out, g_y0, B, T, H, bs) > 0) return true;
  if (B * 4 <= 148 && cluster_try<4>(bwd, cs, F, a, out, g_y0, B, T, H, bs) > 0) return true;
  return cluster_try<2>(bwd, cs, F, a, out, g_y0, B, T, H, bs) > 0;
#endif
}

}  // namespace lncde
