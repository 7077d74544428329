/*
 * lncde_b200.h – C ABI of the B200-native Log-ODE hot path of log-neural-cdes.
 *
 * This is the drop-in boundary.  The reference (pure Python/JAX) has no FFI of
 * its own; each entry point below replaces the XLA-lowered body of the reference
 * function cited beside it (paths relative to the reference repository root).
 * A host binding (ctypes here, jax.ffi where JAX exists – see INTEGRATION.md)
 * passes raw DEVICE pointers, a CUDA stream and integer shapes.
 *
 * Conventions
 *  - all tensors row-major, contiguous, fp32 unless noted; int tables are int32
 *  - `stream` is a cudaStream_t passed as void* (0 = legacy default stream)
 *  - nothing allocates, synchronises or creates streams; all work is enqueued
 *    on `stream`; outputs are caller-allocated and fully overwritten
 *  - return value: 0 OK; <0 argument errors (checked before any launch);
 *    >0 a cudaError_t from the launch
 *  - re-entrant and thread-safe; the only process-wide state is an immutable,
 *    mutex-guarded host cache of Hall tables keyed by (width, depth).  Which kernel
 *    serves a call is a function of its shapes and of its `flags` argument alone.
 */
#ifndef LNCDE_B200_H_
#define LNCDE_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define LNCDE_OK 0
#define LNCDE_ERR_NULL (-1)        /* required pointer is NULL                 */
#define LNCDE_ERR_SHAPE (-2)       /* non-positive / inconsistent dimension    */
#define LNCDE_ERR_UNSUPPORTED (-3) /* depth / width outside the compiled range */
#define LNCDE_ERR_WORKSPACE (-4)   /* workspace too small                      */
#define LNCDE_ERR_ALIGN (-5)       /* pointer not 16-byte aligned              */

/* flags (per call; entry points ignore bits that do not concern them) */
#define LNCDE_FLAG_COMPAT 1u      /* reproduce reference quirks */
#define LNCDE_FLAG_SAVED_FWD 2u   /* lncde_logncde_bwd: workspace was filled by lncde_logncde_fwd */
#define LNCDE_FLAG_K1_TMA 4u      /* lncde_logsig_fwd: force the TMA bulk-copy kernel (default: chosen by windows per path) */
#define LNCDE_FLAG_GENERIC 8u     /* baseline kernels only - classic K1 staging, sequential D / BD scans, FFMA dense flow
                                     build with materialised dF, unstreamed generic Log-NCDE mat-vecs: the kernels every
                                     specialised one is checked against (tests, scripts/variant_check.py) */
#define LNCDE_FLAG_K2_MMA_SYNC 16u /* dense LNCDE: flow build and bracket gradient on the warp-level tensor pipe (mma.sync,
                                      3xTF32) instead of tcgen05 - the path of block sizes that are not multiples of 128 */
#define LNCDE_FLAG_K3_V3 32u      /* Log-NCDE, H = 128 / width 32 model: six-barrier stage kernels (logncde_v3.cuh); measured
                                     slower than the default on B200, kept for its clock-trace instrumentation */
#define LNCDE_FLAG_K2_ONE_CTA 64u /* dense LNCDE: one CTA per series for the two sweeps instead of a thread-block cluster
                                     per series (linear_cluster.cuh) - the path of batches that fill the SMs anyway */

const char* lncde_version(void);
/* Human-readable text for a status returned by any entry point. */
const char* lncde_strerror(int status);

/* ------------------------------------------------------------------ K0 ---
 * Hall basis and tensor->Lie projection tables (host, cached).
 * Replaces data_dir/hall_set.py:60-119 (HallSet.__init__/grow_up),
 * :132-161 (product), :191-208 (rbracket), :234-256 (t2l_matrix).           */

/* len(HallSet(width, depth).data), i.e. 1 + dim Lie^{<=depth}; <0 on error. */
int lncde_hall_size(int width, int depth);
/* pairs[n][2] = HallSet.data (row 0 = (0,0), letters (0,i)).                */
int lncde_hall_basis(int width, int depth, int32_t* pairs);
/* 1 + w + ... + w^depth (hall_set.py:28-33).                                */
int64_t lncde_tensor_dim(int width, int depth);
/* Number of non-zeros of t2l_matrix(depth).                                 */
int lncde_hall_t2l_nnz(int width, int depth);
/* t2l_matrix(depth) in CSR over its len(data) rows; column = flat tensor
 * index INCLUDING the scalar slot 0 (so it is exactly hall_set.py's column).
 * rowptr[n+1], cols[nnz], vals[nnz].                                        */
int lncde_hall_t2l_csr(int width, int depth, int32_t* rowptr, int32_t* cols, float* vals);

/* ------------------------------------------------------------------ K1 ---
 * Batched per-interval truncated log-signature in the Hall basis.
 * Replaces data_dir/generate_paths.py:22-77 (calc_paths) including
 * hall_basis_logsig (:14-19) and the signax signature/log/flatten it calls.
 *
 *   data   [B, L, C]   control path samples
 *   logsig [B, K, D]   K = lncde_logsig_num_windows(L, stepsize),
 *                      D = lncde_hall_size(C, depth); column 0 is always 0
 *   chunk  : with LNCDE_FLAG_COMPAT, the remainder window of sample b gets the
 *            last point of sample b-1 prepended unless b % chunk == 0
 *            (generate_paths.py:59; chunk = B for one calc_paths call, 128 for
 *            datasets.py:43-71 batch_calc_paths).  Ignored without the flag.
 *   t2l_*  : device CSR from lncde_hall_t2l_csr; may be NULL when depth <= 2
 *            and C <= 8 (closed-form Levy-area path).
 *   flags  : LNCDE_FLAG_COMPAT; LNCDE_FLAG_K1_TMA / LNCDE_FLAG_GENERIC force one of the two closed-form kernels
 *            (bit-identical results; by default paths of >= 128 windows take the TMA kernel).
 * depth in [1, 4].                                                          */
int lncde_logsig_num_windows(int L, int stepsize);
int lncde_logsig_fwd(void* stream, const float* data, float* logsig, int B, int L, int C,
                     int stepsize, int depth, unsigned flags, int chunk,
                     const int32_t* t2l_rowptr, const int32_t* t2l_cols, const float* t2l_vals);

/* ------------------------------------------------------------------ K3 ---
 * Log-NCDE: fixed-step Heun solve of the Lie-bracket Log-ODE field.
 * Replaces models/LogNeuralCDEs.py:107-162 (LogNeuralCDE.__call__, inner func)
 * and the diffrax Heun/ConstantStepSize loop it drives, batched over samples
 * (train.py:52-68 jax.vmap).
 *
 *   params : flat fp32 buffer [W0|b0|W1|b1|...|Wn|bn] of the vector-field MLP
 *            (eqx.nn.MLP layout: W_l[out,in] row-major, y = W x + b),
 *            layer sizes  H -> width (x n_hidden) -> C*H
 *   logsig [B, K, D]; y0 [B, H]
 *   stage_row[n_steps*2] int32: logsig row read by each Heun stage
 *            (searchsorted(intervals, t) - 1, wrapped), stage_scale[n_steps*2]:
 *            dt / (intervals[idx] - intervals[idx-1])   (LogNeuralCDEs.py:114-138)
 *   ys [B, n_steps+1, H]: the state after every step (ys[:,0] = y0) – the
 *            checkpoints the backward sweep and the regression read-out need
 * depth in {1, 2}.                                                          */
int lncde_logncde_param_count(int H, int C, int width, int n_hidden);
size_t lncde_logncde_fwd_smem_bytes(int H, int C, int width, int n_hidden);
/* workspace: NULL for inference.  In training pass lncde_logncde_bwd_workspace_bytes(...) bytes;
 * the forward may keep per-stage intermediates there so that the adjoint sweep need not
 * recompute them (then call lncde_logncde_bwd with the same buffer and LNCDE_FLAG_SAVED_FWD). */
int lncde_logncde_fwd(void* stream, const float* params, const float* logsig, const float* y0,
                      const int32_t* stage_row, const float* stage_scale, float* ys, int B, int K,
                      int D, int H, int C, int width, int n_hidden, int depth, int n_steps,
                      void* workspace, size_t workspace_bytes, unsigned flags);

/* Backward sweep (discretise-then-optimise, what diffrax's default adjoint
 * returns).  g_ys [B, n_steps+1, H] is dLoss/d ys (zero where unused).
 * Outputs: g_params [P] (summed over the batch), g_y0 [B, H].
 * workspace: lncde_logncde_bwd_workspace_bytes(...) bytes of device scratch. */
size_t lncde_logncde_bwd_workspace_bytes(int B, int H, int C, int width, int n_hidden, int depth,
                                         int n_steps);
int lncde_logncde_bwd(void* stream, const float* params, const float* logsig,
                      const int32_t* stage_row, const float* stage_scale, const float* ys,
                      const float* g_ys, float* g_params, float* g_y0, int B, int K, int D, int H,
                      int C, int width, int n_hidden, int depth, int n_steps, void* workspace,
                      size_t workspace_bytes, unsigned flags);

/* ----------------------------------------------------------------- K3n ---
 * Neural RDE solve (Heun, constant step) and its exact discrete adjoint.
 * Replaces NeuralRDE.__call__ + func, models/NeuralCDEs.py:227-266 (the reference's comparison
 * baseline on the same (ts, logsig, x0) input):
 *   field = reshape(mlp_linear(vf(y)), (H, D-1)) @ logsig[row][1:] * stage_scale
 *   vf = n_vf linear layers H -> width -> ... -> width, relu between, tanh after the last (NeuralCDEs.py:209-216,
 *        n_vf = the reference's vf_depth); mlp_linear: width -> H*(D-1) (:217-219).
 * params: flat fp32 [W0|b0|...|W_{n_vf-1}|b_{n_vf-1}|Wm|bm], eqx layout W[out,in]; 16-byte aligned.
 * width % 4 == 0, H <= 512, width <= 512, n_vf <= 4.  stage_row / stage_scale / ys / g_ys as for K3.       */
int lncde_nrde_param_count(int H, int Dm, int width, int n_vf); /* Dm = D - 1; < 0 if unsupported */
int lncde_nrde_fwd(void* stream, const float* params, const float* logsig, const float* y0,
                   const int32_t* stage_row, const float* stage_scale, float* ys, int B, int K, int D,
                   int H, int width, int n_vf, int n_steps);
size_t lncde_nrde_bwd_workspace_bytes(int B, int H, int Dm, int width, int n_vf, int n_steps);
int lncde_nrde_bwd(void* stream, const float* params, const float* logsig, const int32_t* stage_row,
                   const float* stage_scale, const float* ys, const float* g_ys, float* g_params,
                   float* g_y0, int B, int K, int D, int H, int width, int n_vf, int n_steps,
                   void* workspace, size_t workspace_bytes);

/* ------------------------------------------------------------------ K2 ---
 * Structured linear CDE (D / BD / DE-LNCDE): flows F_k = I + sum_l logsig[k,l] A_l
 * and the recurrence h_k = F_k h_{k-1}, k = 1..T-1 (flow 0 is never applied).
 * Replaces models/LinearNeuralCDEs.py:302-316 (diagonal), :318-353 (block),
 * :355-386 (scanner).
 *
 *   lie [nb, bs, bs, n_basis]  bracket matrices from log_ode (:192-232);
 *        for bs == 1 this is vf_A^T laid out [H, n_coeff]
 *   logsig [B, T, D]; uses columns 1 .. n_basis
 *   y0 [B, H];  ys [B, T, H] (ys[:,0] = y0)                                  */
/* Scratch shared by forward and backward: the flows F [B,T,nb,bs,bs] plus split-K partials (plus lam [B,T,H] and the
 * packed tensor-core operands for block sizes >= 64, chunk products for the time-parallel scans of block sizes <= 8).
 * The backward call must receive the SAME workspace the forward call of the same (lie, logsig) filled, untouched in
 * between, and the same `flags` (it reads F_t; the generic kernels overwrite it with dF_t).
 * Kernels by block size: bs <= 8 and T > 65: scans parallel in time (chunk products, boundary recurrence over the chunks,
 * per-chunk sweeps); bs | 32: warp-shuffle scans; bs >= 64: flows from tcgen05 (bs % 128 == 0) or mma.sync tensor
 * cores with 3xTF32 split precision; the two sequential sweeps on thread-block clusters of 8 / 4 / 2 CTAs per series
 * (batches of fewer than 75 series; TMA ring of flow slices, state exchanged with st.async onto the peers' mbarriers) or
 * one CTA per series; lam-only reverse sweep + triple-product bracket gradient on tcgen05 (any n_basis) or mma.sync
 * (n_basis <= 36).                                                                                                     */
size_t lncde_linear_scan_workspace_bytes(int B, int T, int nb, int bs, int n_basis);
int lncde_linear_scan_fwd(void* stream, const float* lie, const float* logsig, const float* y0,
                          float* ys, int B, int T, int D, int nb, int bs, int n_basis,
                          void* workspace, size_t workspace_bytes, unsigned flags);
/* g_ys [B,T,H] -> g_lie [nb,bs,bs,n_basis] (summed over batch), g_y0 [B,H].               */
int lncde_linear_scan_bwd(void* stream, const float* lie, const float* logsig, const float* ys,
                          const float* g_ys, float* g_lie, float* g_y0, int B, int T, int D, int nb,
                          int bs, int n_basis, void* workspace, size_t workspace_bytes, unsigned flags);

/* ----------------------------------------------------------------- K2a ---
 * Bracket matrices of the structured linear CDE and their gradient.  Replaces LogLinearCDE.log_ode,
 * models/LinearNeuralCDEs.py:192-232 (A_[u,v] = A_v A_u - A_u A_v level by level, stacked on the last axis) and its
 * reverse-mode derivative.
 *   vf    [C, nb, bs, bs]          the letters' matrices (vf_A reshaped, :318)
 *   pairs [n_basis][2]  HOST int32: HallSet(C, depth).data[1:] - (left, right) element numbers, 1-based, left == 0
 *                       for the letters (data_dir/hall_set.py:60-119; the order roughpy's lie_basis enumerates, :111-116)
 *   lie   [nb, bs, bs, n_basis]    what lncde_linear_scan_fwd takes
 * n_basis <= 160, degree <= 6.  workspace: lncde_lie_brackets_workspace_bytes bytes; the forward call leaves every A_k
 * there ([n_basis, nb, bs, bs]) and the backward call reads it (`workspace`) next to a scratch of the same size
 * (`workspace_g`).  g_vf [C, nb, bs, bs] is fully overwritten; deterministic (gather per parent, no atomics).       */
size_t lncde_lie_brackets_workspace_bytes(int n_basis, int nb, int bs);
int lncde_lie_brackets_fwd(void* stream, const float* vf, const int32_t* pairs, float* lie, int C, int n_basis, int nb,
                           int bs, void* workspace, size_t workspace_bytes);
int lncde_lie_brackets_bwd(void* stream, const float* g_lie, const int32_t* pairs, float* g_vf, int C, int n_basis, int nb,
                           int bs, const void* workspace, void* workspace_g, size_t workspace_bytes);

/* ----------------------------------------------------------------- K2d ---
 * Fused diagonal LNCDE (block size 1): f_t = 1 + sum_c logsig[t, 1 + c] vf_A[c, :], y_t = f_t * y_{t-1},
 * t = 1..T-1, WITHOUT materialising the flows (models/LinearNeuralCDEs.py:302-316 + scanner :355-386).
 * Same results as lncde_linear_scan_fwd/bwd with bs = 1, n_basis = C; a third of their HBM traffic.
 *   vf_At [H, C] (= vf_A transposed, the `lie` layout of the generic call for bs = 1); C <= 8
 *   logsig [B, T, D] with D >= 1 + C; y0 [B, H]; ys [B, T, H] (ys[:,0] = y0)
 *   g_ys [B, T, H] -> g_vf_At [H, C] (summed over the batch), g_y0 [B, H]
 * workspace (backward only): lncde_dlncde_bwd_workspace_bytes(B, H, C) bytes.                              */
size_t lncde_dlncde_bwd_workspace_bytes(int B, int H, int C);
int lncde_dlncde_fwd(void* stream, const float* vf_At, const float* logsig, const float* y0, float* ys,
                     int B, int T, int D, int H, int C);
int lncde_dlncde_bwd(void* stream, const float* vf_At, const float* logsig, const float* ys,
                     const float* g_ys, float* g_vf_At, float* g_y0, int B, int T, int D, int H, int C,
                     void* workspace, size_t workspace_bytes);
/* The same pair for the classification read-out, which only sees the mean of the states over time
 * (LinearNeuralCDEs.py:389): the forward also returns y_mean [B, H] = mean_t ys, the backward takes g_mean [B, H] =
 * dLoss/dy_mean and uses g_ys[b, t, :] = g_mean[b, :] / T for every t - no [B, T, H] cotangent is written or read. */
int lncde_dlncde_fwd_mean(void* stream, const float* vf_At, const float* logsig, const float* y0, float* ys,
                          float* y_mean, int B, int T, int D, int H, int C);
int lncde_dlncde_bwd_mean(void* stream, const float* vf_At, const float* logsig, const float* ys,
                          const float* g_mean, float* g_vf_At, float* g_y0, int B, int T, int D, int H, int C,
                          void* workspace, size_t workspace_bytes);

/* ------------------------------------------------------------------ K4 ---
 * Fused Adam update on a flat parameter buffer (optax.adam defaults,
 * train.py:134-135,183): m,v moments updated in place, bias-corrected.
 * grad_scale multiplies the gradient first (1/world after an all-reduce sum). */
int lncde_adam_step(void* stream, float* params, const float* grads, float* m, float* v, int64_t n,
                    float lr, float beta1, float beta2, float eps, int step, float grad_scale);

/* Same update with the step count held in device memory (*step_counter = steps done so far, incremented
 * by the call): a launch captured in a CUDA graph stays valid across replays.                          */
int lncde_adam_step_dev(void* stream, float* params, const float* grads, float* m, float* v, int64_t n,
                        float lr, float beta1, float beta2, float eps, int* step_counter, float grad_scale);

/* ------------------------------------------------------------------ K5 ---
 * The classification train step around the solve, as kernels: initial state, read-out + cross-entropy and the Lip(2)
 * regulariser, each with its gradient (fixed summation orders, no atomics).  Replaces the XLA-lowered bodies of
 *   y0 = linear1(x0)                          models/LogNeuralCDEs.py:139 (init_layer: models/LinearNeuralCDEs.py:303)
 *   pred = softmax(linear2(h))                models/LogNeuralCDEs.py:159-160, models/LinearNeuralCDEs.py:388-391
 *   mean_b(-sum_c y log(pred + 1e-8))         train.py:95 (classification_loss)
 *   lambd * sum_layers mean(||W_row|| + ||b||) train.py:79-93
 * and of their reverse-mode derivatives under eqx.filter_value_and_grad (train.py:71-72).                       */

/* y [B, n_out] = x [B, n_in] W^T + b   (eqx.nn.Linear layout W [n_out, n_in]) */
int lncde_affine_fwd(void* stream, const float* x, const float* W, const float* b, float* y, int B, int n_in, int n_out);
/* g_Wb = [g_W (n_out x n_in) | g_b (n_out)], the layout of an eqx.nn.Linear inside the flat parameter vector:
 * g_W = g_y^T x, g_b = sum_b g_y (fully overwritten).  Sums over the batch run in chunks of 32 samples, combined in a
 * fixed order; workspace: lncde_affine_bwd_workspace_bytes (0 bytes / may be NULL for B <= 32).                    */
size_t lncde_affine_bwd_workspace_bytes(int B, int n_in, int n_out);
int lncde_affine_bwd(void* stream, const float* x, const float* g_y, float* g_Wb, int B, int n_in, int n_out,
                     void* workspace, size_t workspace_bytes);

/* h: B rows of H floats, row s at h + s * h_stride (e.g. the last state of ys [B, n_steps + 1, H]); labels [B, L] one-hot
 * (or soft); L <= 32.  Outputs, all overwritten: pred [B, L] (may be NULL), g_h rows (stride g_h_stride) = d loss / d h,
 * g_Wb_loss = [g_W (L x H) | g_b (L) | loss (1)] - the linear layer's gradient in flat-parameter layout followed by the
 * mean cross-entropy.  workspace: lncde_readout_ce_workspace_bytes.                                                 */
size_t lncde_readout_ce_workspace_bytes(int B, int H, int L);
int lncde_readout_ce(void* stream, const float* h, long long h_stride, const float* W, const float* b,
                     const float* labels, float* pred, float* g_h, long long g_h_stride, float* g_Wb_loss, int B, int H,
                     int L, void* workspace, size_t workspace_bytes);

/* out [B, H] = mean over t of ys [B, T, H] (the LNCDE read-out, models/LinearNeuralCDEs.py:389) and its gradient
 * g_ys [B, T, H] = g_mean [B, H] / T (fully overwritten).  H <= 1024.                                                */
int lncde_time_mean_fwd(void* stream, const float* ys, float* out, int B, int T, int H);
int lncde_time_mean_bwd(void* stream, const float* g_mean, float* g_ys, int B, int T, int H);

/* loss [1] += lambd * sum_l (mean_rows ||W_l[row, :]|| + ||b_l||), g_params += its gradient.  Layer l: W_l at
 * params + w_off[l] as [rows[l], cols[l]], b_l at params + b_off[l] as [rows[l]] (b_off[l] < 0: no bias term - the
 * LNCDE's vf_A); the four index arrays are HOST arrays of n_layers <= 8 entries.                                     */
int lncde_lip2(void* stream, const float* params, float* loss, float* g_params, int n_layers, const int32_t* w_off,
               const int32_t* b_off, const int32_t* rows, const int32_t* cols, float lambd);

/* ------------------------------------------------------- measurement ---
 * FP32 FFMA issue-rate probe (no reference counterpart; SURVEY.md 8d asks for a measured FP32 peak next to the
 * HBM / tensor peaks of MEASURED_PEAKS.json).  Launches `blocks` CTAs of 256 threads, each thread running 8
 * independent FFMA chains for 4 * iters rounds; out [blocks * 256] floats.  lncde_probe_ffma_flops returns the
 * floating-point operations of one such launch (2 per FFMA).                                             */
int64_t lncde_probe_ffma_flops(int blocks, int iters);
int lncde_probe_ffma(void* stream, float* out, int blocks, int iters);

#ifdef __cplusplus
}
#endif
#endif /* LNCDE_B200_H_ */
