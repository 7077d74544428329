// jax.ffi (XLA FFI) handlers over the C ABI of include/lncde_b200.h.
//
// Compiled ONLY where the jaxlib headers exist (xla/ffi/api/ffi.h ships inside jaxlib:
// `python -c "import jax; print(jax.ffi.include_dir())"`).  JAX is not installed in the build
// image or on the GPU boxes, so this file is shipped as source and is UNTESTED here; the
// tested binding is ctypes + torch (lncde/_lib.py).  INTEGRATION.md shows the Python side.
#if defined(__has_include)
#if __has_include("xla/ffi/api/ffi.h")
#define LNCDE_HAVE_XLA_FFI 1
#endif
#endif

#ifdef LNCDE_HAVE_XLA_FFI
#include <cuda_runtime_api.h>

#include <string>

#include "lncde_b200.h"
#include "xla/ffi/api/ffi.h"

namespace ffi = xla::ffi;

static ffi::Error status_to_error(int st, const char* what) {
  if (st == 0) return ffi::Error::Success();
  return ffi::Error(ffi::ErrorCode::kInternal, std::string(what) + ": " + lncde_strerror(st));
}

// logsig = calc_paths(data, stepsize, depth)   (data_dir/generate_paths.py:22)
static ffi::Error LogsigImpl(cudaStream_t stream, ffi::Buffer<ffi::F32> data, ffi::Buffer<ffi::S32> rowptr,
                             ffi::Buffer<ffi::S32> cols, ffi::Buffer<ffi::F32> vals, int32_t stepsize, int32_t depth,
                             int32_t compat, int32_t chunk, ffi::ResultBuffer<ffi::F32> out) {
  auto d = data.dimensions();
  return status_to_error(
      lncde_logsig_fwd(stream, data.typed_data(), out->typed_data(), (int)d[0], (int)d[1], (int)d[2], stepsize, depth,
                       compat ? LNCDE_FLAG_COMPAT : 0u, chunk, rowptr.typed_data(), cols.typed_data(),
                       vals.typed_data()),
      "lncde_logsig_fwd");
}
XLA_FFI_DEFINE_HANDLER_SYMBOL(lncde_xla_logsig, LogsigImpl,
                              ffi::Ffi::Bind()
                                  .Ctx<ffi::PlatformStream<cudaStream_t>>()
                                  .Arg<ffi::Buffer<ffi::F32>>()
                                  .Arg<ffi::Buffer<ffi::S32>>()
                                  .Arg<ffi::Buffer<ffi::S32>>()
                                  .Arg<ffi::Buffer<ffi::F32>>()
                                  .Attr<int32_t>("stepsize")
                                  .Attr<int32_t>("depth")
                                  .Attr<int32_t>("compat")
                                  .Attr<int32_t>("chunk")
                                  .Ret<ffi::Buffer<ffi::F32>>());

// ys = solve(params, logsig, y0; stage table)   (models/LogNeuralCDEs.py:147-157)
static ffi::Error SolveFwdImpl(cudaStream_t stream, ffi::Buffer<ffi::F32> params, ffi::Buffer<ffi::F32> logsig,
                               ffi::Buffer<ffi::F32> y0, ffi::Buffer<ffi::S32> rows, ffi::Buffer<ffi::F32> scale,
                               int32_t C, int32_t width, int32_t n_hidden, int32_t depth,
                               ffi::ResultBuffer<ffi::F32> ys) {
  auto l = logsig.dimensions();
  auto y = y0.dimensions();
  const int n_steps = (int)rows.dimensions()[0];
  return status_to_error(
      lncde_logncde_fwd(stream, params.typed_data(), logsig.typed_data(), y0.typed_data(), rows.typed_data(),
                        scale.typed_data(), ys->typed_data(), (int)l[0], (int)l[1], (int)l[2], (int)y[1], C, width,
                        n_hidden, depth, n_steps, nullptr, 0, 0u),
      "lncde_logncde_fwd");
}
XLA_FFI_DEFINE_HANDLER_SYMBOL(lncde_xla_logncde_fwd, SolveFwdImpl,
                              ffi::Ffi::Bind()
                                  .Ctx<ffi::PlatformStream<cudaStream_t>>()
                                  .Arg<ffi::Buffer<ffi::F32>>()
                                  .Arg<ffi::Buffer<ffi::F32>>()
                                  .Arg<ffi::Buffer<ffi::F32>>()
                                  .Arg<ffi::Buffer<ffi::S32>>()
                                  .Arg<ffi::Buffer<ffi::F32>>()
                                  .Attr<int32_t>("C")
                                  .Attr<int32_t>("width")
                                  .Attr<int32_t>("n_hidden")
                                  .Attr<int32_t>("depth")
                                  .Ret<ffi::Buffer<ffi::F32>>());

// (g_params, g_y0) = solve_bwd(...); scratch is a declared extra result so XLA owns it.
static ffi::Error SolveBwdImpl(cudaStream_t stream, ffi::Buffer<ffi::F32> params, ffi::Buffer<ffi::F32> logsig,
                               ffi::Buffer<ffi::S32> rows, ffi::Buffer<ffi::F32> scale, ffi::Buffer<ffi::F32> ys,
                               ffi::Buffer<ffi::F32> g_ys, int32_t C, int32_t width, int32_t n_hidden, int32_t depth,
                               ffi::ResultBuffer<ffi::F32> g_params, ffi::ResultBuffer<ffi::F32> g_y0,
                               ffi::ResultBuffer<ffi::U8> scratch) {
  auto l = logsig.dimensions();
  auto y = ys.dimensions();
  return status_to_error(
      lncde_logncde_bwd(stream, params.typed_data(), logsig.typed_data(), rows.typed_data(), scale.typed_data(),
                        ys.typed_data(), g_ys.typed_data(), g_params->typed_data(), g_y0->typed_data(), (int)l[0],
                        (int)l[1], (int)l[2], (int)y[2], C, width, n_hidden, depth, (int)y[1] - 1,
                        scratch->typed_data(), scratch->size_bytes(), 0u),
      "lncde_logncde_bwd");
}
XLA_FFI_DEFINE_HANDLER_SYMBOL(lncde_xla_logncde_bwd, SolveBwdImpl,
                              ffi::Ffi::Bind()
                                  .Ctx<ffi::PlatformStream<cudaStream_t>>()
                                  .Arg<ffi::Buffer<ffi::F32>>()
                                  .Arg<ffi::Buffer<ffi::F32>>()
                                  .Arg<ffi::Buffer<ffi::S32>>()
                                  .Arg<ffi::Buffer<ffi::F32>>()
                                  .Arg<ffi::Buffer<ffi::F32>>()
                                  .Arg<ffi::Buffer<ffi::F32>>()
                                  .Attr<int32_t>("C")
                                  .Attr<int32_t>("width")
                                  .Attr<int32_t>("n_hidden")
                                  .Attr<int32_t>("depth")
                                  .Ret<ffi::Buffer<ffi::F32>>()
                                  .Ret<ffi::Buffer<ffi::F32>>()
                                  .Ret<ffi::Buffer<ffi::U8>>());

// ---- K3n: Neural RDE (NeuralCDEs.py:227-266).  params = [vf layers | mlp_linear]; attrs: width, n_vf.
static ffi::Error NrdeFwdImpl(cudaStream_t stream, ffi::Buffer<ffi::F32> params, ffi::Buffer<ffi::F32> logsig,
                              ffi::Buffer<ffi::F32> y0, ffi::Buffer<ffi::S32> rows, ffi::Buffer<ffi::F32> scale,
                              int32_t width, int32_t n_vf, ffi::ResultBuffer<ffi::F32> ys) {
  auto l = logsig.dimensions();
  auto y = y0.dimensions();
  const int n_steps = (int)rows.dimensions()[0];
  return status_to_error(
      lncde_nrde_fwd(stream, params.typed_data(), logsig.typed_data(), y0.typed_data(), rows.typed_data(),
                     scale.typed_data(), ys->typed_data(), (int)l[0], (int)l[1], (int)l[2], (int)y[1], width, n_vf, n_steps),
      "lncde_nrde_fwd");
}
XLA_FFI_DEFINE_HANDLER_SYMBOL(lncde_xla_nrde_fwd, NrdeFwdImpl,
                              ffi::Ffi::Bind()
                                  .Ctx<ffi::PlatformStream<cudaStream_t>>()
                                  .Arg<ffi::Buffer<ffi::F32>>()
                                  .Arg<ffi::Buffer<ffi::F32>>()
                                  .Arg<ffi::Buffer<ffi::F32>>()
                                  .Arg<ffi::Buffer<ffi::S32>>()
                                  .Arg<ffi::Buffer<ffi::F32>>()
                                  .Attr<int32_t>("width")
                                  .Attr<int32_t>("n_vf")
                                  .Ret<ffi::Buffer<ffi::F32>>());

static ffi::Error NrdeBwdImpl(cudaStream_t stream, ffi::Buffer<ffi::F32> params, ffi::Buffer<ffi::F32> logsig,
                              ffi::Buffer<ffi::S32> rows, ffi::Buffer<ffi::F32> scale, ffi::Buffer<ffi::F32> ys,
                              ffi::Buffer<ffi::F32> g_ys, int32_t width, int32_t n_vf,
                              ffi::ResultBuffer<ffi::F32> g_params, ffi::ResultBuffer<ffi::F32> g_y0,
                              ffi::ResultBuffer<ffi::U8> scratch) {
  auto l = logsig.dimensions();
  auto y = ys.dimensions();
  return status_to_error(
      lncde_nrde_bwd(stream, params.typed_data(), logsig.typed_data(), rows.typed_data(), scale.typed_data(),
                     ys.typed_data(), g_ys.typed_data(), g_params->typed_data(), g_y0->typed_data(), (int)l[0], (int)l[1],
                     (int)l[2], (int)y[2], width, n_vf, (int)y[1] - 1, scratch->typed_data(), scratch->size_bytes()),
      "lncde_nrde_bwd");
}
XLA_FFI_DEFINE_HANDLER_SYMBOL(lncde_xla_nrde_bwd, NrdeBwdImpl,
                              ffi::Ffi::Bind()
                                  .Ctx<ffi::PlatformStream<cudaStream_t>>()
                                  .Arg<ffi::Buffer<ffi::F32>>()
                                  .Arg<ffi::Buffer<ffi::F32>>()
                                  .Arg<ffi::Buffer<ffi::S32>>()
                                  .Arg<ffi::Buffer<ffi::F32>>()
                                  .Arg<ffi::Buffer<ffi::F32>>()
                                  .Arg<ffi::Buffer<ffi::F32>>()
                                  .Attr<int32_t>("width")
                                  .Attr<int32_t>("n_vf")
                                  .Ret<ffi::Buffer<ffi::F32>>()
                                  .Ret<ffi::Buffer<ffi::F32>>()
                                  .Ret<ffi::Buffer<ffi::U8>>());
// ---- K2: D / BD / DE-LNCDE (LinearNeuralCDEs.py:302-386).  lie [nb*bs*bs, n_basis] are the bracket matrices of log_ode
// (parameters only: stays in JAX), logsig [B, T, D], y0 [B, H]; attrs: nb, bs.  The scratch (flows) is a declared
// extra result so that XLA owns it.
static ffi::Error LinearScanFwdImpl(cudaStream_t stream, ffi::Buffer<ffi::F32> lie, ffi::Buffer<ffi::F32> logsig,
                                    ffi::Buffer<ffi::F32> y0, int32_t nb, int32_t bs, ffi::ResultBuffer<ffi::F32> ys,
                                    ffi::ResultBuffer<ffi::U8> scratch) {
  auto l = logsig.dimensions();
  const int n_basis = (int)lie.dimensions()[1];
  return status_to_error(
      lncde_linear_scan_fwd(stream, lie.typed_data(), logsig.typed_data(), y0.typed_data(), ys->typed_data(), (int)l[0],
                            (int)l[1], (int)l[2], nb, bs, n_basis, scratch->typed_data(), scratch->size_bytes(), 0u),
      "lncde_linear_scan_fwd");
}
XLA_FFI_DEFINE_HANDLER_SYMBOL(lncde_xla_linear_scan_fwd, LinearScanFwdImpl,
                              ffi::Ffi::Bind()
                                  .Ctx<ffi::PlatformStream<cudaStream_t>>()
                                  .Arg<ffi::Buffer<ffi::F32>>()
                                  .Arg<ffi::Buffer<ffi::F32>>()
                                  .Arg<ffi::Buffer<ffi::F32>>()
                                  .Attr<int32_t>("nb")
                                  .Attr<int32_t>("bs")
                                  .Ret<ffi::Buffer<ffi::F32>>()
                                  .Ret<ffi::Buffer<ffi::U8>>());

// (g_lie, g_y0) = linear_scan_bwd(...).  The C entry point needs the flows its forward call left in the workspace; XLA
// hands inputs over as const buffers, so this handler REBUILDS them (one more forward into its own scratch and a
// throw-away ys) instead of aliasing the forward's scratch - a maintainer who wants to save that pass registers the
// call with input_output_aliases on the scratch operand and drops the re-run.
static ffi::Error LinearScanBwdImpl(cudaStream_t stream, ffi::Buffer<ffi::F32> lie, ffi::Buffer<ffi::F32> logsig,
                                    ffi::Buffer<ffi::F32> y0, ffi::Buffer<ffi::F32> g_ys, int32_t nb, int32_t bs,
                                    ffi::ResultBuffer<ffi::F32> g_lie, ffi::ResultBuffer<ffi::F32> g_y0,
                                    ffi::ResultBuffer<ffi::F32> ys_scratch, ffi::ResultBuffer<ffi::U8> scratch) {
  auto l = logsig.dimensions();
  const int B = (int)l[0], T = (int)l[1], D = (int)l[2], n_basis = (int)lie.dimensions()[1];
  int st = lncde_linear_scan_fwd(stream, lie.typed_data(), logsig.typed_data(), y0.typed_data(), ys_scratch->typed_data(),
                                 B, T, D, nb, bs, n_basis, scratch->typed_data(), scratch->size_bytes(), 0u);
  if (st != 0) return status_to_error(st, "lncde_linear_scan_fwd (rebuild)");
  return status_to_error(
      lncde_linear_scan_bwd(stream, lie.typed_data(), logsig.typed_data(), ys_scratch->typed_data(), g_ys.typed_data(),
                            g_lie->typed_data(), g_y0->typed_data(), B, T, D, nb, bs, n_basis, scratch->typed_data(),
                            scratch->size_bytes(), 0u),
      "lncde_linear_scan_bwd");
}
XLA_FFI_DEFINE_HANDLER_SYMBOL(lncde_xla_linear_scan_bwd, LinearScanBwdImpl,
                              ffi::Ffi::Bind()
                                  .Ctx<ffi::PlatformStream<cudaStream_t>>()
                                  .Arg<ffi::Buffer<ffi::F32>>()
                                  .Arg<ffi::Buffer<ffi::F32>>()
                                  .Arg<ffi::Buffer<ffi::F32>>()
                                  .Arg<ffi::Buffer<ffi::F32>>()
                                  .Attr<int32_t>("nb")
                                  .Attr<int32_t>("bs")
                                  .Ret<ffi::Buffer<ffi::F32>>()
                                  .Ret<ffi::Buffer<ffi::F32>>()
                                  .Ret<ffi::Buffer<ffi::F32>>()
                                  .Ret<ffi::Buffer<ffi::U8>>());
#endif  // LNCDE_HAVE_XLA_FFI
