// XLA typed-FFI handlers for idoc_b200: the thin shim between JAX and the C ABI of include/idoc_b200.h.
//
// Boundary being replaced (the reference has no FFI of its own; its boundary is the Python call on pytrees):
//   idoc_lqr_fwd_ffi / idoc_lqr_vjp_ffi        <- idoc.lqr.build(T).direct / .implicit      idoc/lqr.py:167-180
//   idoc_lqr_backward_ffi                      <- idoc.lqr.backward(..., return_expected_change=True)   idoc/lqr.py:62-107
//   idoc_lqr_kkt_ffi                           <- idoc.lqr.kkt                                idoc/lqr.py:128-150
//   idoc_tilqr_fwd_ffi / idoc_tilqr_vjp_ffi    <- idoc.tilqr.build(T).direct / .implicit      idoc/tilqr.py:35-104
//   idoc_ilqr_solve_ffi / idoc_ilqr_vjp_ffi    <- idoc.ilqr.build(...).direct / .implicit     idoc/ilqr.py:179-278
//   idoc_ilqr_solve_box_ffi / idoc_ilqr_vjp_box_ffi   control-limited extension (no reference counterpart)
//   idoc_ilqr_linearize_ffi                    <- make_lqr_approx                             idoc/ilqr.py:127-158
//   idoc_blqr_fwd_ffi / idoc_blqr_vjp_ffi      <- idoc.blqr.build(T).direct / .implicit       idoc/blqr.py:245-261
//   idoc_bilqr_solve_box_ffi / idoc_bilqr_vjp_box_ffi  <- idoc.bilqr.build(...).direct / .implicit (idoc/bilqr.py:120-221) as one
//                                              kernel, with optional limits on the shared control (has_limits; no reference counterpart)
//
// Every handler only unpacks XLA buffers (dense row-major, leading batch axes flattened into B), checks dtypes / ranks,
// and forwards to the C ABI on the stream XLA provides: no allocation, no synchronisation, no global state, so the calls
// are CUDA-graph capturable and thread-safe per stream.  Scratch (the residual workspace `ws` that idoc_lqr_fwd writes and
// idoc_lqr_vjp reads, and the VJP workspace `ws2`) arrives as extra uint8 result buffers sized from Python with the
// idoc_*_ws_bytes functions; `ws` is a residual of the custom_vjp (idoc_b200/jax_api.py).
//
// Build (on a machine with jaxlib; the headers ship inside it, `python -c "import jax.ffi; print(jax.ffi.include_dir())"`):
//   g++ -O2 -std=c++17 -fPIC -shared -I$(python -c "import jax.ffi as f; print(f.include_dir())") -I/usr/local/cuda/include
//       -Iinclude ffi/idoc_b200_ffi.cc -Lidoc_b200 -l:libidoc_b200.so -Wl,-rpath,'$ORIGIN/../idoc_b200' -o ffi/libidoc_b200_ffi.so
// jaxlib is not installed in the build image of this repository (no network), so this file is compiled against the real
// headers only on a JAX-equipped machine; tests/test_ffi_shim.py compiles it here against a minimal stand-in of the header
// surface it uses (ffi/mock/) to keep it free of syntax and signature errors.
#include <cuda_runtime_api.h>

#include <cstdint>
#include <string>

#include "idoc_b200.h"
#include "xla/ffi/api/ffi.h"

namespace ffi = xla::ffi;

namespace {

struct Dims {
  int64_t B;
  int T, n, m;
};

// dtype code of the C ABI, or -1
int DType(const ffi::AnyBuffer& b) {
  switch (b.element_type()) {
    case ffi::F32: return IDOC_F32;
    case ffi::F64: return IDOC_F64;
    default: return -1;
  }
}

// A[..., T, n, n] and Bm[..., T, n, m] (time-varying) -> B = product of the leading axes
ffi::ErrorOr<Dims> DimsFromAB(const ffi::AnyBuffer& A, const ffi::AnyBuffer& Bm, bool time_varying) {
  const auto da = A.dimensions(), db = Bm.dimensions();
  const size_t tail = time_varying ? 3 : 2;
  if (da.size() < tail || db.size() != da.size()) return ffi::Unexpected(ffi::Error::InvalidArgument("A / B rank mismatch"));
  Dims d;
  d.B = 1;
  for (size_t i = 0; i + tail < da.size(); i++) d.B *= da[i];
  d.T = time_varying ? static_cast<int>(da[da.size() - 3]) : 0;
  d.n = static_cast<int>(da[da.size() - 1]);
  d.m = static_cast<int>(db[db.size() - 1]);
  if (da[da.size() - 2] != d.n || db[db.size() - 2] != d.n) return ffi::Unexpected(ffi::Error::InvalidArgument("A must be n x n, B n x m"));
  return d;
}

ffi::Error Status(int rc, const char* what) {
  if (rc == IDOC_OK) return ffi::Error::Success();
  const std::string msg = std::string(what) + ": " + idoc_last_error();
  if (rc == IDOC_ERR_INVALID || rc == IDOC_ERR_UNSUPPORTED || rc == IDOC_ERR_WORKSPACE) return ffi::Error::InvalidArgument(msg);
  return ffi::Error::Internal(msg);
}
ffi::Error IlqrStatus(int rc, const char* what) {
  if (rc == IDOC_OK) return ffi::Error::Success();
  const std::string msg = std::string(what) + ": " + idoc_ilqr_last_error();
  if (rc == IDOC_ERR_INVALID || rc == IDOC_ERR_UNSUPPORTED || rc == IDOC_ERR_WORKSPACE) return ffi::Error::InvalidArgument(msg);
  return ffi::Error::Internal(msg);
}

#define IDOC_SAME_DTYPE(ref, buf)                                                                \
  if ((buf).element_type() != (ref).element_type())                                              \
    return ffi::Error::InvalidArgument("every floating-point buffer of a call must have the same dtype");

// ------------------------------------------------------------------------------------------------ lqr
ffi::Error LqrFwd(cudaStream_t stream, ffi::AnyBuffer Q, ffi::AnyBuffer q, ffi::AnyBuffer Qf, ffi::AnyBuffer qf, ffi::AnyBuffer M,
                  ffi::AnyBuffer R, ffi::AnyBuffer r, ffi::AnyBuffer A, ffi::AnyBuffer Bm, ffi::AnyBuffer d, ffi::AnyBuffer x0,
                  ffi::Result<ffi::AnyBuffer> X, ffi::Result<ffi::AnyBuffer> U, ffi::Result<ffi::AnyBuffer> Nu,
                  ffi::Result<ffi::AnyBuffer> K, ffi::Result<ffi::AnyBuffer> k, ffi::Result<ffi::AnyBuffer> dCdc,
                  ffi::Result<ffi::AnyBuffer> status, ffi::Result<ffi::AnyBuffer> ws) {
  const int dt = DType(A);
  if (dt < 0) return ffi::Error::InvalidArgument("idoc_lqr_fwd: dtype must be float32 or float64");
  IDOC_SAME_DTYPE(A, Q) IDOC_SAME_DTYPE(A, q) IDOC_SAME_DTYPE(A, Qf) IDOC_SAME_DTYPE(A, qf) IDOC_SAME_DTYPE(A, M)
  IDOC_SAME_DTYPE(A, R) IDOC_SAME_DTYPE(A, r) IDOC_SAME_DTYPE(A, Bm) IDOC_SAME_DTYPE(A, d) IDOC_SAME_DTYPE(A, x0)
  auto dims = DimsFromAB(A, Bm, true);
  if (dims.has_error()) return dims.error();
  const Dims D = dims.value();
  return Status(idoc_lqr_fwd(stream, dt, D.B, D.T, D.n, D.m, Q.untyped_data(), q.untyped_data(), Qf.untyped_data(), qf.untyped_data(),
                             M.untyped_data(), R.untyped_data(), r.untyped_data(), A.untyped_data(), Bm.untyped_data(),
                             d.untyped_data(), x0.untyped_data(), X->untyped_data(), U->untyped_data(), Nu->untyped_data(),
                             K->untyped_data(), k->untyped_data(), dCdc->untyped_data(),
                             static_cast<int32_t*>(status->untyped_data()), ws->untyped_data(), ws->size_bytes()),
                "idoc_lqr_fwd");
}

ffi::Error LqrBackward(cudaStream_t stream, ffi::AnyBuffer Q, ffi::AnyBuffer q, ffi::AnyBuffer Qf, ffi::AnyBuffer qf, ffi::AnyBuffer M,
                       ffi::AnyBuffer R, ffi::AnyBuffer r, ffi::AnyBuffer A, ffi::AnyBuffer Bm, ffi::AnyBuffer d,
                       ffi::Result<ffi::AnyBuffer> K, ffi::Result<ffi::AnyBuffer> k, ffi::Result<ffi::AnyBuffer> dCdc,
                       ffi::Result<ffi::AnyBuffer> status, ffi::Result<ffi::AnyBuffer> ws) {
  const int dt = DType(A);
  if (dt < 0) return ffi::Error::InvalidArgument("idoc_lqr_backward: dtype must be float32 or float64");
  auto dims = DimsFromAB(A, Bm, true);
  if (dims.has_error()) return dims.error();
  const Dims D = dims.value();
  return Status(idoc_lqr_backward(stream, dt, D.B, D.T, D.n, D.m, Q.untyped_data(), q.untyped_data(), Qf.untyped_data(),
                                  qf.untyped_data(), M.untyped_data(), R.untyped_data(), r.untyped_data(), A.untyped_data(),
                                  Bm.untyped_data(), d.untyped_data(), K->untyped_data(), k->untyped_data(), dCdc->untyped_data(),
                                  static_cast<int32_t*>(status->untyped_data()), ws->untyped_data(), ws->size_bytes()),
                "idoc_lqr_backward");
}

// has_nub = 0: Nub is a 1-element placeholder and the Nu cotangent is taken as zero (every reference test: losses use X, U only)
ffi::Error LqrVjp(cudaStream_t stream, ffi::AnyBuffer M, ffi::AnyBuffer A, ffi::AnyBuffer Bm, ffi::AnyBuffer x0, ffi::AnyBuffer X,
                  ffi::AnyBuffer U, ffi::AnyBuffer Nu, ffi::AnyBuffer ws, ffi::AnyBuffer Xb, ffi::AnyBuffer Ub, ffi::AnyBuffer Nub,
                  ffi::Result<ffi::AnyBuffer> gQ, ffi::Result<ffi::AnyBuffer> gq, ffi::Result<ffi::AnyBuffer> gQf,
                  ffi::Result<ffi::AnyBuffer> gqf, ffi::Result<ffi::AnyBuffer> gM, ffi::Result<ffi::AnyBuffer> gR,
                  ffi::Result<ffi::AnyBuffer> gr, ffi::Result<ffi::AnyBuffer> gA, ffi::Result<ffi::AnyBuffer> gB,
                  ffi::Result<ffi::AnyBuffer> gd, ffi::Result<ffi::AnyBuffer> gx0, ffi::Result<ffi::AnyBuffer> ws2, bool has_nub,
                  bool reduce_batch) {
  const int dt = DType(A);
  if (dt < 0) return ffi::Error::InvalidArgument("idoc_lqr_vjp: dtype must be float32 or float64");
  IDOC_SAME_DTYPE(A, Xb) IDOC_SAME_DTYPE(A, Ub) IDOC_SAME_DTYPE(A, X) IDOC_SAME_DTYPE(A, U) IDOC_SAME_DTYPE(A, Nu)
  auto dims = DimsFromAB(A, Bm, true);
  if (dims.has_error()) return dims.error();
  const Dims D = dims.value();
  return Status(idoc_lqr_vjp(stream, dt, D.B, D.T, D.n, D.m, M.untyped_data(), A.untyped_data(), Bm.untyped_data(), x0.untyped_data(),
                             X.untyped_data(), U.untyped_data(), Nu.untyped_data(), ws.untyped_data(), Xb.untyped_data(),
                             Ub.untyped_data(), has_nub ? Nub.untyped_data() : nullptr, gQ->untyped_data(), gq->untyped_data(),
                             gQf->untyped_data(), gqf->untyped_data(), gM->untyped_data(), gR->untyped_data(), gr->untyped_data(),
                             gA->untyped_data(), gB->untyped_data(), gd->untyped_data(), gx0->untyped_data(), ws2->untyped_data(),
                             ws2->size_bytes(), reduce_batch ? 1 : 0),
                "idoc_lqr_vjp");
}

ffi::Error LqrKkt(cudaStream_t stream, ffi::AnyBuffer Q, ffi::AnyBuffer q, ffi::AnyBuffer Qf, ffi::AnyBuffer qf, ffi::AnyBuffer M,
                  ffi::AnyBuffer R, ffi::AnyBuffer r, ffi::AnyBuffer A, ffi::AnyBuffer Bm, ffi::AnyBuffer d, ffi::AnyBuffer x0,
                  ffi::AnyBuffer X, ffi::AnyBuffer U, ffi::AnyBuffer Nu, ffi::Result<ffi::AnyBuffer> rX,
                  ffi::Result<ffi::AnyBuffer> rU, ffi::Result<ffi::AnyBuffer> rNu) {
  const int dt = DType(A);
  if (dt < 0) return ffi::Error::InvalidArgument("idoc_lqr_kkt: dtype must be float32 or float64");
  auto dims = DimsFromAB(A, Bm, true);
  if (dims.has_error()) return dims.error();
  const Dims D = dims.value();
  return Status(idoc_lqr_kkt(stream, dt, D.B, D.T, D.n, D.m, Q.untyped_data(), q.untyped_data(), Qf.untyped_data(), qf.untyped_data(),
                             M.untyped_data(), R.untyped_data(), r.untyped_data(), A.untyped_data(), Bm.untyped_data(),
                             d.untyped_data(), x0.untyped_data(), X.untyped_data(), U.untyped_data(), Nu.untyped_data(),
                             rX->untyped_data(), rU->untyped_data(), rNu->untyped_data()),
                "idoc_lqr_kkt");
}

// ------------------------------------------------------------------------------------------------ tilqr
ffi::Error TilqrFwd(cudaStream_t stream, ffi::AnyBuffer Q, ffi::AnyBuffer Qf, ffi::AnyBuffer R, ffi::AnyBuffer A, ffi::AnyBuffer Bm,
                    ffi::AnyBuffer x0, ffi::Result<ffi::AnyBuffer> X, ffi::Result<ffi::AnyBuffer> U, ffi::Result<ffi::AnyBuffer> Nu,
                    ffi::Result<ffi::AnyBuffer> status, ffi::Result<ffi::AnyBuffer> ws, int64_t horizon) {
  const int dt = DType(A);
  if (dt < 0) return ffi::Error::InvalidArgument("idoc_tilqr_fwd: dtype must be float32 or float64");
  IDOC_SAME_DTYPE(A, Q) IDOC_SAME_DTYPE(A, Qf) IDOC_SAME_DTYPE(A, R) IDOC_SAME_DTYPE(A, Bm) IDOC_SAME_DTYPE(A, x0)
  auto dims = DimsFromAB(A, Bm, false);
  if (dims.has_error()) return dims.error();
  const Dims D = dims.value();
  return Status(idoc_tilqr_fwd(stream, dt, D.B, static_cast<int>(horizon), D.n, D.m, Q.untyped_data(), Qf.untyped_data(),
                               R.untyped_data(), A.untyped_data(), Bm.untyped_data(), x0.untyped_data(), X->untyped_data(),
                               U->untyped_data(), Nu->untyped_data(), static_cast<int32_t*>(status->untyped_data()),
                               ws->untyped_data(), ws->size_bytes()),
                "idoc_tilqr_fwd");
}

ffi::Error TilqrVjp(cudaStream_t stream, ffi::AnyBuffer A, ffi::AnyBuffer Bm, ffi::AnyBuffer x0, ffi::AnyBuffer X, ffi::AnyBuffer U,
                    ffi::AnyBuffer Nu, ffi::AnyBuffer ws, ffi::AnyBuffer Xb, ffi::AnyBuffer Ub, ffi::AnyBuffer Nub,
                    ffi::Result<ffi::AnyBuffer> gQ, ffi::Result<ffi::AnyBuffer> gQf, ffi::Result<ffi::AnyBuffer> gR,
                    ffi::Result<ffi::AnyBuffer> gA, ffi::Result<ffi::AnyBuffer> gB, ffi::Result<ffi::AnyBuffer> gx0,
                    ffi::Result<ffi::AnyBuffer> ws2, int64_t horizon, bool has_nub) {
  const int dt = DType(A);
  if (dt < 0) return ffi::Error::InvalidArgument("idoc_tilqr_vjp: dtype must be float32 or float64");
  auto dims = DimsFromAB(A, Bm, false);
  if (dims.has_error()) return dims.error();
  const Dims D = dims.value();
  return Status(idoc_tilqr_vjp(stream, dt, D.B, static_cast<int>(horizon), D.n, D.m, A.untyped_data(), Bm.untyped_data(),
                               x0.untyped_data(), X.untyped_data(), U.untyped_data(), Nu.untyped_data(), ws.untyped_data(),
                               Xb.untyped_data(), Ub.untyped_data(), has_nub ? Nub.untyped_data() : nullptr, gQ->untyped_data(),
                               gQf->untyped_data(), gR->untyped_data(), gA->untyped_data(), gB->untyped_data(), gx0->untyped_data(),
                               ws2->untyped_data(), ws2->size_bytes()),
                "idoc_tilqr_vjp");
}

// ------------------------------------------------------------------------------------------------ ilqr
// theta[B, theta_dim], x0[B, systems * n]; X0 / U0: the initial trajectory (XLA inputs are immutable, so it is copied into the
// X / U results, which idoc_ilqr_solve then iterates in place).  No host synchronisation (allow_sync = 0): every problem runs
// until its own convergence flag or `maxiter`.
struct IlqrDims {
  int64_t B;
  int T, N, m, theta_dim;
};
ffi::ErrorOr<IlqrDims> IlqrShapes(const ffi::AnyBuffer& theta, const ffi::AnyBuffer& X, const ffi::AnyBuffer& U) {
  const auto dx = X.dimensions(), du = U.dimensions(), dth = theta.dimensions();
  if (dx.size() < 2 || du.size() != dx.size() || dth.size() + 1 != dx.size())
    return ffi::Unexpected(ffi::Error::InvalidArgument("ilqr: expected theta[..., P], X[..., T, N], U[..., T, m]"));
  IlqrDims d;
  d.B = 1;
  for (size_t i = 0; i + 2 < dx.size(); i++) d.B *= dx[i];
  d.T = static_cast<int>(dx[dx.size() - 2]);
  d.N = static_cast<int>(dx[dx.size() - 1]);
  d.m = static_cast<int>(du[du.size() - 1]);
  d.theta_dim = static_cast<int>(dth[dth.size() - 1]);
  return d;
}

ffi::Error IlqrSolve(cudaStream_t stream, ffi::AnyBuffer theta, ffi::AnyBuffer x0, ffi::AnyBuffer X0, ffi::AnyBuffer U0,
                     ffi::Result<ffi::AnyBuffer> X, ffi::Result<ffi::AnyBuffer> U, ffi::Result<ffi::AnyBuffer> Nu,
                     ffi::Result<ffi::AnyBuffer> iters, ffi::Result<ffi::AnyBuffer> ls_failed, ffi::Result<ffi::AnyBuffer> ws,
                     int64_t problem, int64_t systems, int64_t maxiter, double thres, bool use_line_search, double ls_lower,
                     double ls_upper, double ls_rho, int64_t ls_maxiter) {
  const int dt = DType(X0);
  if (dt < 0) return ffi::Error::InvalidArgument("idoc_ilqr_solve: dtype must be float32 or float64");
  IDOC_SAME_DTYPE(X0, theta) IDOC_SAME_DTYPE(X0, x0) IDOC_SAME_DTYPE(X0, U0)
  auto dims = IlqrShapes(theta, X0, U0);
  if (dims.has_error()) return dims.error();
  const IlqrDims D = dims.value();
  if (D.N % systems != 0) return ffi::Error::InvalidArgument("idoc_ilqr_solve: state width is not a multiple of `systems`");
  if (cudaMemcpyAsync(X->untyped_data(), X0.untyped_data(), X0.size_bytes(), cudaMemcpyDeviceToDevice, stream) != cudaSuccess ||
      cudaMemcpyAsync(U->untyped_data(), U0.untyped_data(), U0.size_bytes(), cudaMemcpyDeviceToDevice, stream) != cudaSuccess)
    return ffi::Error::Internal("idoc_ilqr_solve: copying the initial trajectory failed");
  return IlqrStatus(idoc_ilqr_solve(stream, dt, static_cast<int>(problem), D.B, D.T, static_cast<int>(D.N / systems), D.m,
                                    static_cast<int>(systems), theta.untyped_data(), D.theta_dim, x0.untyped_data(),
                                    X->untyped_data(), U->untyped_data(), Nu->untyped_data(), static_cast<int>(maxiter), thres,
                                    use_line_search ? 1 : 0, ls_lower, ls_upper, ls_rho, static_cast<int>(ls_maxiter),
                                    static_cast<int32_t*>(iters->untyped_data()), static_cast<int32_t*>(ls_failed->untyped_data()),
                                    ws->untyped_data(), ws->size_bytes(), /*allow_sync=*/0),
                    "idoc_ilqr_solve");
}

ffi::Error IlqrSolveBox(cudaStream_t stream, ffi::AnyBuffer theta, ffi::AnyBuffer x0, ffi::AnyBuffer u_lo, ffi::AnyBuffer u_hi,
                        ffi::AnyBuffer X0, ffi::AnyBuffer U0, ffi::Result<ffi::AnyBuffer> X, ffi::Result<ffi::AnyBuffer> U,
                        ffi::Result<ffi::AnyBuffer> Nu, ffi::Result<ffi::AnyBuffer> iters, ffi::Result<ffi::AnyBuffer> ls_failed,
                        ffi::Result<ffi::AnyBuffer> ws, int64_t problem, int64_t maxiter, double thres, bool use_line_search,
                        double ls_lower, double ls_upper, double ls_rho, int64_t ls_maxiter) {
  const int dt = DType(X0);
  if (dt < 0) return ffi::Error::InvalidArgument("idoc_ilqr_solve_box: dtype must be float32 or float64");
  IDOC_SAME_DTYPE(X0, theta) IDOC_SAME_DTYPE(X0, x0) IDOC_SAME_DTYPE(X0, U0) IDOC_SAME_DTYPE(X0, u_lo) IDOC_SAME_DTYPE(X0, u_hi)
  auto dims = IlqrShapes(theta, X0, U0);
  if (dims.has_error()) return dims.error();
  const IlqrDims D = dims.value();
  if (cudaMemcpyAsync(X->untyped_data(), X0.untyped_data(), X0.size_bytes(), cudaMemcpyDeviceToDevice, stream) != cudaSuccess ||
      cudaMemcpyAsync(U->untyped_data(), U0.untyped_data(), U0.size_bytes(), cudaMemcpyDeviceToDevice, stream) != cudaSuccess)
    return ffi::Error::Internal("idoc_ilqr_solve_box: copying the initial trajectory failed");
  return IlqrStatus(idoc_ilqr_solve_box(stream, dt, static_cast<int>(problem), D.B, D.T, D.N, D.m, theta.untyped_data(), D.theta_dim,
                                        x0.untyped_data(), u_lo.untyped_data(), u_hi.untyped_data(), X->untyped_data(),
                                        U->untyped_data(), Nu->untyped_data(), static_cast<int>(maxiter), thres,
                                        use_line_search ? 1 : 0, ls_lower, ls_upper, ls_rho, static_cast<int>(ls_maxiter),
                                        static_cast<int32_t*>(iters->untyped_data()),
                                        static_cast<int32_t*>(ls_failed->untyped_data()), ws->untyped_data(), ws->size_bytes()),
                    "idoc_ilqr_solve_box");
}

ffi::Error IlqrVjp(cudaStream_t stream, ffi::AnyBuffer theta, ffi::AnyBuffer x0, ffi::AnyBuffer X, ffi::AnyBuffer U, ffi::AnyBuffer Nu,
                   ffi::AnyBuffer Xb, ffi::AnyBuffer Ub, ffi::AnyBuffer Nub, ffi::Result<ffi::AnyBuffer> gtheta,
                   ffi::Result<ffi::AnyBuffer> gx0, ffi::Result<ffi::AnyBuffer> ws, int64_t problem, int64_t systems, bool has_nub) {
  const int dt = DType(X);
  if (dt < 0) return ffi::Error::InvalidArgument("idoc_ilqr_vjp: dtype must be float32 or float64");
  IDOC_SAME_DTYPE(X, theta) IDOC_SAME_DTYPE(X, x0) IDOC_SAME_DTYPE(X, U) IDOC_SAME_DTYPE(X, Nu) IDOC_SAME_DTYPE(X, Xb) IDOC_SAME_DTYPE(X, Ub)
  auto dims = IlqrShapes(theta, X, U);
  if (dims.has_error()) return dims.error();
  const IlqrDims D = dims.value();
  if (D.N % systems != 0) return ffi::Error::InvalidArgument("idoc_ilqr_vjp: state width is not a multiple of `systems`");
  return IlqrStatus(idoc_ilqr_vjp(stream, dt, static_cast<int>(problem), D.B, D.T, static_cast<int>(D.N / systems), D.m,
                                  static_cast<int>(systems), theta.untyped_data(), D.theta_dim, x0.untyped_data(), X.untyped_data(),
                                  U.untyped_data(), Nu.untyped_data(), Xb.untyped_data(), Ub.untyped_data(),
                                  has_nub ? Nub.untyped_data() : nullptr, gtheta->untyped_data(), gx0->untyped_data(),
                                  ws->untyped_data(), ws->size_bytes()),
                    "idoc_ilqr_vjp");
}

ffi::Error IlqrVjpBox(cudaStream_t stream, ffi::AnyBuffer theta, ffi::AnyBuffer x0, ffi::AnyBuffer u_lo, ffi::AnyBuffer u_hi,
                      ffi::AnyBuffer X, ffi::AnyBuffer U, ffi::AnyBuffer Nu, ffi::AnyBuffer Xb, ffi::AnyBuffer Ub, ffi::AnyBuffer Nub,
                      ffi::Result<ffi::AnyBuffer> gtheta, ffi::Result<ffi::AnyBuffer> gx0, ffi::Result<ffi::AnyBuffer> ws,
                      int64_t problem, bool has_nub) {
  const int dt = DType(X);
  if (dt < 0) return ffi::Error::InvalidArgument("idoc_ilqr_vjp_box: dtype must be float32 or float64");
  auto dims = IlqrShapes(theta, X, U);
  if (dims.has_error()) return dims.error();
  const IlqrDims D = dims.value();
  return IlqrStatus(idoc_ilqr_vjp_box(stream, dt, static_cast<int>(problem), D.B, D.T, D.N, D.m, theta.untyped_data(), D.theta_dim,
                                      x0.untyped_data(), u_lo.untyped_data(), u_hi.untyped_data(), X.untyped_data(), U.untyped_data(),
                                      Nu.untyped_data(), Xb.untyped_data(), Ub.untyped_data(), has_nub ? Nub.untyped_data() : nullptr,
                                      gtheta->untyped_data(), gx0->untyped_data(), ws->untyped_data(), ws->size_bytes()),
                    "idoc_ilqr_vjp_box");
}

// make_lqr_approx (idoc/ilqr.py:127-158): mode 0 local, 1 global, 2 global + Lagrangian curvature (needs Nu)
ffi::Error IlqrLinearize(cudaStream_t stream, ffi::AnyBuffer theta, ffi::AnyBuffer x0, ffi::AnyBuffer X, ffi::AnyBuffer U,
                         ffi::AnyBuffer Nu, ffi::Result<ffi::AnyBuffer> Q, ffi::Result<ffi::AnyBuffer> q,
                         ffi::Result<ffi::AnyBuffer> Qf, ffi::Result<ffi::AnyBuffer> qf, ffi::Result<ffi::AnyBuffer> M,
                         ffi::Result<ffi::AnyBuffer> R, ffi::Result<ffi::AnyBuffer> r, ffi::Result<ffi::AnyBuffer> A,
                         ffi::Result<ffi::AnyBuffer> Bm, ffi::Result<ffi::AnyBuffer> d, int64_t problem, int64_t systems, int64_t mode) {
  const int dt = DType(X);
  if (dt < 0) return ffi::Error::InvalidArgument("idoc_ilqr_linearize: dtype must be float32 or float64");
  auto dims = IlqrShapes(theta, X, U);
  if (dims.has_error()) return dims.error();
  const IlqrDims D = dims.value();
  return IlqrStatus(idoc_ilqr_linearize(stream, dt, static_cast<int>(problem), D.B, D.T, static_cast<int>(D.N / systems), D.m,
                                        static_cast<int>(systems), theta.untyped_data(), D.theta_dim, x0.untyped_data(),
                                        X.untyped_data(), U.untyped_data(), mode == 2 ? Nu.untyped_data() : nullptr,
                                        static_cast<int>(mode), Q->untyped_data(), q->untyped_data(), Qf->untyped_data(),
                                        qf->untyped_data(), M->untyped_data(), R->untyped_data(), r->untyped_data(), A->untyped_data(),
                                        Bm->untyped_data(), d->untyped_data()),
                    "idoc_ilqr_linearize");
}

// ------------------------------------------------------------------------------------------------ blqr (shared controls)
// A[..., T, bs, n, n], Bm[..., T, bs, n, m]: P = product of the leading axes
struct BlqrDims {
  int64_t P;
  int T, bs, n, m;
};
ffi::ErrorOr<BlqrDims> BlqrShapes(const ffi::AnyBuffer& A, const ffi::AnyBuffer& Bm) {
  const auto da = A.dimensions(), db = Bm.dimensions();
  if (da.size() < 4 || db.size() != da.size()) return ffi::Unexpected(ffi::Error::InvalidArgument("blqr: A[..., T, bs, n, n], B[..., T, bs, n, m]"));
  BlqrDims d;
  d.P = 1;
  for (size_t i = 0; i + 4 < da.size(); i++) d.P *= da[i];
  d.T = static_cast<int>(da[da.size() - 4]);
  d.bs = static_cast<int>(da[da.size() - 3]);
  d.n = static_cast<int>(da[da.size() - 1]);
  d.m = static_cast<int>(db[db.size() - 1]);
  if (da[da.size() - 2] != d.n || db[db.size() - 2] != d.n || db[db.size() - 3] != d.bs)
    return ffi::Unexpected(ffi::Error::InvalidArgument("blqr: A must be bs x n x n, B bs x n x m"));
  return d;
}

ffi::Error BlqrFwd(cudaStream_t stream, ffi::AnyBuffer Q, ffi::AnyBuffer q, ffi::AnyBuffer Qf, ffi::AnyBuffer qf, ffi::AnyBuffer M,
                   ffi::AnyBuffer R, ffi::AnyBuffer r, ffi::AnyBuffer A, ffi::AnyBuffer Bm, ffi::AnyBuffer d, ffi::AnyBuffer x0,
                   ffi::Result<ffi::AnyBuffer> X, ffi::Result<ffi::AnyBuffer> U, ffi::Result<ffi::AnyBuffer> Nu,
                   ffi::Result<ffi::AnyBuffer> K, ffi::Result<ffi::AnyBuffer> k, ffi::Result<ffi::AnyBuffer> dCdc,
                   ffi::Result<ffi::AnyBuffer> status, ffi::Result<ffi::AnyBuffer> ws) {
  const int dt = DType(A);
  if (dt < 0) return ffi::Error::InvalidArgument("idoc_blqr_fwd: dtype must be float32 or float64");
  IDOC_SAME_DTYPE(A, Q) IDOC_SAME_DTYPE(A, q) IDOC_SAME_DTYPE(A, Qf) IDOC_SAME_DTYPE(A, qf) IDOC_SAME_DTYPE(A, M)
  IDOC_SAME_DTYPE(A, R) IDOC_SAME_DTYPE(A, r) IDOC_SAME_DTYPE(A, Bm) IDOC_SAME_DTYPE(A, d) IDOC_SAME_DTYPE(A, x0)
  auto dims = BlqrShapes(A, Bm);
  if (dims.has_error()) return dims.error();
  const BlqrDims D = dims.value();
  return Status(idoc_blqr_fwd(stream, dt, D.P, D.T, D.bs, D.n, D.m, Q.untyped_data(), q.untyped_data(), Qf.untyped_data(),
                              qf.untyped_data(), M.untyped_data(), R.untyped_data(), r.untyped_data(), A.untyped_data(),
                              Bm.untyped_data(), d.untyped_data(), x0.untyped_data(), X->untyped_data(), U->untyped_data(),
                              Nu->untyped_data(), K->untyped_data(), k->untyped_data(), dCdc->untyped_data(),
                              static_cast<int32_t*>(status->untyped_data()), ws->untyped_data(), ws->size_bytes()),
                "idoc_blqr_fwd");
}

ffi::Error BlqrVjp(cudaStream_t stream, ffi::AnyBuffer Q, ffi::AnyBuffer Qf, ffi::AnyBuffer M, ffi::AnyBuffer R, ffi::AnyBuffer A,
                   ffi::AnyBuffer Bm, ffi::AnyBuffer x0, ffi::AnyBuffer X, ffi::AnyBuffer U, ffi::AnyBuffer Nu, ffi::AnyBuffer Xb,
                   ffi::AnyBuffer Ub, ffi::AnyBuffer Nub, ffi::Result<ffi::AnyBuffer> gQ, ffi::Result<ffi::AnyBuffer> gq,
                   ffi::Result<ffi::AnyBuffer> gQf, ffi::Result<ffi::AnyBuffer> gqf, ffi::Result<ffi::AnyBuffer> gM,
                   ffi::Result<ffi::AnyBuffer> gR, ffi::Result<ffi::AnyBuffer> gr, ffi::Result<ffi::AnyBuffer> gA,
                   ffi::Result<ffi::AnyBuffer> gB, ffi::Result<ffi::AnyBuffer> gd, ffi::Result<ffi::AnyBuffer> gx0,
                   ffi::Result<ffi::AnyBuffer> ws, bool has_nub) {
  const int dt = DType(A);
  if (dt < 0) return ffi::Error::InvalidArgument("idoc_blqr_vjp: dtype must be float32 or float64");
  IDOC_SAME_DTYPE(A, Q) IDOC_SAME_DTYPE(A, Qf) IDOC_SAME_DTYPE(A, M) IDOC_SAME_DTYPE(A, R) IDOC_SAME_DTYPE(A, Bm) IDOC_SAME_DTYPE(A, x0)
  IDOC_SAME_DTYPE(A, X) IDOC_SAME_DTYPE(A, U) IDOC_SAME_DTYPE(A, Nu) IDOC_SAME_DTYPE(A, Xb) IDOC_SAME_DTYPE(A, Ub)
  auto dims = BlqrShapes(A, Bm);
  if (dims.has_error()) return dims.error();
  const BlqrDims D = dims.value();
  return Status(idoc_blqr_vjp(stream, dt, D.P, D.T, D.bs, D.n, D.m, Q.untyped_data(), Qf.untyped_data(), M.untyped_data(),
                              R.untyped_data(), A.untyped_data(), Bm.untyped_data(), x0.untyped_data(), X.untyped_data(),
                              U.untyped_data(), Nu.untyped_data(), Xb.untyped_data(), Ub.untyped_data(),
                              has_nub ? Nub.untyped_data() : nullptr, gQ->untyped_data(), gq->untyped_data(), gQf->untyped_data(),
                              gqf->untyped_data(), gM->untyped_data(), gR->untyped_data(), gr->untyped_data(), gA->untyped_data(),
                              gB->untyped_data(), gd->untyped_data(), gx0->untyped_data(), ws->untyped_data(), ws->size_bytes()),
                "idoc_blqr_vjp");
}

// ------------------------------------------------------------------------------------------------ bilqr, one kernel (+ limits)
// has_limits = 0: u_lo / u_hi are 1-element placeholders
ffi::Error BilqrSolveBox(cudaStream_t stream, ffi::AnyBuffer theta, ffi::AnyBuffer x0, ffi::AnyBuffer u_lo, ffi::AnyBuffer u_hi,
                         ffi::AnyBuffer X0, ffi::AnyBuffer U0, ffi::Result<ffi::AnyBuffer> X, ffi::Result<ffi::AnyBuffer> U,
                         ffi::Result<ffi::AnyBuffer> Nu, ffi::Result<ffi::AnyBuffer> iters, ffi::Result<ffi::AnyBuffer> ls_failed,
                         ffi::Result<ffi::AnyBuffer> ws, int64_t problem, int64_t systems, bool has_limits, int64_t maxiter,
                         double thres, bool use_line_search, double ls_lower, double ls_upper, double ls_rho, int64_t ls_maxiter) {
  const int dt = DType(X0);
  if (dt < 0) return ffi::Error::InvalidArgument("idoc_bilqr_solve_box: dtype must be float32 or float64");
  IDOC_SAME_DTYPE(X0, theta) IDOC_SAME_DTYPE(X0, x0) IDOC_SAME_DTYPE(X0, U0) IDOC_SAME_DTYPE(X0, u_lo) IDOC_SAME_DTYPE(X0, u_hi)
  auto dims = IlqrShapes(theta, X0, U0);
  if (dims.has_error()) return dims.error();
  const IlqrDims D = dims.value();
  if (systems <= 0 || D.N % systems != 0) return ffi::Error::InvalidArgument("idoc_bilqr_solve_box: state width is not a multiple of `systems`");
  if (cudaMemcpyAsync(X->untyped_data(), X0.untyped_data(), X0.size_bytes(), cudaMemcpyDeviceToDevice, stream) != cudaSuccess ||
      cudaMemcpyAsync(U->untyped_data(), U0.untyped_data(), U0.size_bytes(), cudaMemcpyDeviceToDevice, stream) != cudaSuccess)
    return ffi::Error::Internal("idoc_bilqr_solve_box: copying the initial trajectory failed");
  return IlqrStatus(idoc_bilqr_solve_box(stream, dt, static_cast<int>(problem), D.B, D.T, static_cast<int>(D.N / systems), D.m,
                                         static_cast<int>(systems), theta.untyped_data(), D.theta_dim, x0.untyped_data(),
                                         has_limits ? u_lo.untyped_data() : nullptr, has_limits ? u_hi.untyped_data() : nullptr,
                                         X->untyped_data(), U->untyped_data(), Nu->untyped_data(), static_cast<int>(maxiter), thres,
                                         use_line_search ? 1 : 0, ls_lower, ls_upper, ls_rho, static_cast<int>(ls_maxiter),
                                         static_cast<int32_t*>(iters->untyped_data()),
                                         static_cast<int32_t*>(ls_failed->untyped_data()), ws->untyped_data(), ws->size_bytes()),
                    "idoc_bilqr_solve_box");
}

ffi::Error BilqrVjpBox(cudaStream_t stream, ffi::AnyBuffer theta, ffi::AnyBuffer x0, ffi::AnyBuffer u_lo, ffi::AnyBuffer u_hi,
                       ffi::AnyBuffer X, ffi::AnyBuffer U, ffi::AnyBuffer Nu, ffi::AnyBuffer Xb, ffi::AnyBuffer Ub, ffi::AnyBuffer Nub,
                       ffi::Result<ffi::AnyBuffer> gtheta, ffi::Result<ffi::AnyBuffer> gx0, ffi::Result<ffi::AnyBuffer> ws,
                       int64_t problem, int64_t systems, bool has_limits, bool has_nub) {
  const int dt = DType(X);
  if (dt < 0) return ffi::Error::InvalidArgument("idoc_bilqr_vjp_box: dtype must be float32 or float64");
  auto dims = IlqrShapes(theta, X, U);
  if (dims.has_error()) return dims.error();
  const IlqrDims D = dims.value();
  if (systems <= 0 || D.N % systems != 0) return ffi::Error::InvalidArgument("idoc_bilqr_vjp_box: state width is not a multiple of `systems`");
  return IlqrStatus(idoc_bilqr_vjp_box(stream, dt, static_cast<int>(problem), D.B, D.T, static_cast<int>(D.N / systems), D.m,
                                       static_cast<int>(systems), theta.untyped_data(), D.theta_dim, x0.untyped_data(),
                                       has_limits ? u_lo.untyped_data() : nullptr, has_limits ? u_hi.untyped_data() : nullptr,
                                       X.untyped_data(), U.untyped_data(), Nu.untyped_data(), Xb.untyped_data(), Ub.untyped_data(),
                                       has_nub ? Nub.untyped_data() : nullptr, gtheta->untyped_data(), gx0->untyped_data(),
                                       ws->untyped_data(), ws->size_bytes()),
                    "idoc_bilqr_vjp_box");
}

}  // namespace

#define IDOC_ARGS_11 .Arg<ffi::AnyBuffer>().Arg<ffi::AnyBuffer>().Arg<ffi::AnyBuffer>().Arg<ffi::AnyBuffer>().Arg<ffi::AnyBuffer>() \
                     .Arg<ffi::AnyBuffer>().Arg<ffi::AnyBuffer>().Arg<ffi::AnyBuffer>().Arg<ffi::AnyBuffer>().Arg<ffi::AnyBuffer>() \
                     .Arg<ffi::AnyBuffer>()
#define IDOC_RET(N) IDOC_RET_##N
#define IDOC_RET_1 .Ret<ffi::AnyBuffer>()
#define IDOC_RET_2 IDOC_RET_1 IDOC_RET_1
#define IDOC_RET_3 IDOC_RET_2 IDOC_RET_1
#define IDOC_RET_5 IDOC_RET_3 IDOC_RET_2
#define IDOC_RET_6 IDOC_RET_3 IDOC_RET_3
#define IDOC_RET_7 IDOC_RET_5 IDOC_RET_2
#define IDOC_RET_8 IDOC_RET_5 IDOC_RET_3
#define IDOC_RET_10 IDOC_RET_5 IDOC_RET_5
#define IDOC_RET_12 IDOC_RET_10 IDOC_RET_2

XLA_FFI_DEFINE_HANDLER_SYMBOL(idoc_lqr_fwd_ffi, LqrFwd,
                              ffi::Ffi::Bind().Ctx<ffi::PlatformStream<cudaStream_t>>() IDOC_ARGS_11 IDOC_RET(8));
XLA_FFI_DEFINE_HANDLER_SYMBOL(idoc_lqr_backward_ffi, LqrBackward,
                              ffi::Ffi::Bind().Ctx<ffi::PlatformStream<cudaStream_t>>()
                                  .Arg<ffi::AnyBuffer>().Arg<ffi::AnyBuffer>().Arg<ffi::AnyBuffer>().Arg<ffi::AnyBuffer>().Arg<ffi::AnyBuffer>()
                                  .Arg<ffi::AnyBuffer>().Arg<ffi::AnyBuffer>().Arg<ffi::AnyBuffer>().Arg<ffi::AnyBuffer>().Arg<ffi::AnyBuffer>()
                                  IDOC_RET(5));
XLA_FFI_DEFINE_HANDLER_SYMBOL(idoc_lqr_vjp_ffi, LqrVjp,
                              ffi::Ffi::Bind().Ctx<ffi::PlatformStream<cudaStream_t>>() IDOC_ARGS_11 IDOC_RET(12)
                                  .Attr<bool>("has_nub").Attr<bool>("reduce_batch"));
XLA_FFI_DEFINE_HANDLER_SYMBOL(idoc_lqr_kkt_ffi, LqrKkt,
                              ffi::Ffi::Bind().Ctx<ffi::PlatformStream<cudaStream_t>>() IDOC_ARGS_11
                                  .Arg<ffi::AnyBuffer>().Arg<ffi::AnyBuffer>().Arg<ffi::AnyBuffer>() IDOC_RET(3));
XLA_FFI_DEFINE_HANDLER_SYMBOL(idoc_tilqr_fwd_ffi, TilqrFwd,
                              ffi::Ffi::Bind().Ctx<ffi::PlatformStream<cudaStream_t>>()
                                  .Arg<ffi::AnyBuffer>().Arg<ffi::AnyBuffer>().Arg<ffi::AnyBuffer>().Arg<ffi::AnyBuffer>().Arg<ffi::AnyBuffer>()
                                  .Arg<ffi::AnyBuffer>() IDOC_RET(5).Attr<int64_t>("horizon"));
XLA_FFI_DEFINE_HANDLER_SYMBOL(idoc_tilqr_vjp_ffi, TilqrVjp,
                              ffi::Ffi::Bind().Ctx<ffi::PlatformStream<cudaStream_t>>()
                                  .Arg<ffi::AnyBuffer>().Arg<ffi::AnyBuffer>().Arg<ffi::AnyBuffer>().Arg<ffi::AnyBuffer>().Arg<ffi::AnyBuffer>()
                                  .Arg<ffi::AnyBuffer>().Arg<ffi::AnyBuffer>().Arg<ffi::AnyBuffer>().Arg<ffi::AnyBuffer>().Arg<ffi::AnyBuffer>()
                                  IDOC_RET(7).Attr<int64_t>("horizon").Attr<bool>("has_nub"));
XLA_FFI_DEFINE_HANDLER_SYMBOL(idoc_ilqr_solve_ffi, IlqrSolve,
                              ffi::Ffi::Bind().Ctx<ffi::PlatformStream<cudaStream_t>>()
                                  .Arg<ffi::AnyBuffer>().Arg<ffi::AnyBuffer>().Arg<ffi::AnyBuffer>().Arg<ffi::AnyBuffer>() IDOC_RET(6)
                                  .Attr<int64_t>("problem").Attr<int64_t>("systems").Attr<int64_t>("maxiter").Attr<double>("thres")
                                  .Attr<bool>("use_line_search").Attr<double>("ls_lower").Attr<double>("ls_upper")
                                  .Attr<double>("ls_rho").Attr<int64_t>("ls_maxiter"));
XLA_FFI_DEFINE_HANDLER_SYMBOL(idoc_ilqr_solve_box_ffi, IlqrSolveBox,
                              ffi::Ffi::Bind().Ctx<ffi::PlatformStream<cudaStream_t>>()
                                  .Arg<ffi::AnyBuffer>().Arg<ffi::AnyBuffer>().Arg<ffi::AnyBuffer>().Arg<ffi::AnyBuffer>().Arg<ffi::AnyBuffer>()
                                  .Arg<ffi::AnyBuffer>() IDOC_RET(6)
                                  .Attr<int64_t>("problem").Attr<int64_t>("maxiter").Attr<double>("thres").Attr<bool>("use_line_search")
                                  .Attr<double>("ls_lower").Attr<double>("ls_upper").Attr<double>("ls_rho").Attr<int64_t>("ls_maxiter"));
XLA_FFI_DEFINE_HANDLER_SYMBOL(idoc_ilqr_vjp_ffi, IlqrVjp,
                              ffi::Ffi::Bind().Ctx<ffi::PlatformStream<cudaStream_t>>()
                                  .Arg<ffi::AnyBuffer>().Arg<ffi::AnyBuffer>().Arg<ffi::AnyBuffer>().Arg<ffi::AnyBuffer>().Arg<ffi::AnyBuffer>()
                                  .Arg<ffi::AnyBuffer>().Arg<ffi::AnyBuffer>().Arg<ffi::AnyBuffer>() IDOC_RET(3)
                                  .Attr<int64_t>("problem").Attr<int64_t>("systems").Attr<bool>("has_nub"));
XLA_FFI_DEFINE_HANDLER_SYMBOL(idoc_ilqr_vjp_box_ffi, IlqrVjpBox,
                              ffi::Ffi::Bind().Ctx<ffi::PlatformStream<cudaStream_t>>()
                                  .Arg<ffi::AnyBuffer>().Arg<ffi::AnyBuffer>().Arg<ffi::AnyBuffer>().Arg<ffi::AnyBuffer>().Arg<ffi::AnyBuffer>()
                                  .Arg<ffi::AnyBuffer>().Arg<ffi::AnyBuffer>().Arg<ffi::AnyBuffer>().Arg<ffi::AnyBuffer>().Arg<ffi::AnyBuffer>()
                                  IDOC_RET(3).Attr<int64_t>("problem").Attr<bool>("has_nub"));
XLA_FFI_DEFINE_HANDLER_SYMBOL(idoc_ilqr_linearize_ffi, IlqrLinearize,
                              ffi::Ffi::Bind().Ctx<ffi::PlatformStream<cudaStream_t>>()
                                  .Arg<ffi::AnyBuffer>().Arg<ffi::AnyBuffer>().Arg<ffi::AnyBuffer>().Arg<ffi::AnyBuffer>().Arg<ffi::AnyBuffer>()
                                  IDOC_RET(10).Attr<int64_t>("problem").Attr<int64_t>("systems").Attr<int64_t>("mode"));
XLA_FFI_DEFINE_HANDLER_SYMBOL(idoc_blqr_fwd_ffi, BlqrFwd,
                              ffi::Ffi::Bind().Ctx<ffi::PlatformStream<cudaStream_t>>() IDOC_ARGS_11 IDOC_RET(8));
XLA_FFI_DEFINE_HANDLER_SYMBOL(idoc_blqr_vjp_ffi, BlqrVjp,
                              ffi::Ffi::Bind().Ctx<ffi::PlatformStream<cudaStream_t>>() IDOC_ARGS_11
                                  .Arg<ffi::AnyBuffer>().Arg<ffi::AnyBuffer>() IDOC_RET(12).Attr<bool>("has_nub"));
XLA_FFI_DEFINE_HANDLER_SYMBOL(idoc_bilqr_solve_box_ffi, BilqrSolveBox,
                              ffi::Ffi::Bind().Ctx<ffi::PlatformStream<cudaStream_t>>()
                                  .Arg<ffi::AnyBuffer>().Arg<ffi::AnyBuffer>().Arg<ffi::AnyBuffer>().Arg<ffi::AnyBuffer>().Arg<ffi::AnyBuffer>()
                                  .Arg<ffi::AnyBuffer>() IDOC_RET(6)
                                  .Attr<int64_t>("problem").Attr<int64_t>("systems").Attr<bool>("has_limits").Attr<int64_t>("maxiter")
                                  .Attr<double>("thres").Attr<bool>("use_line_search").Attr<double>("ls_lower")
                                  .Attr<double>("ls_upper").Attr<double>("ls_rho").Attr<int64_t>("ls_maxiter"));
XLA_FFI_DEFINE_HANDLER_SYMBOL(idoc_bilqr_vjp_box_ffi, BilqrVjpBox,
                              ffi::Ffi::Bind().Ctx<ffi::PlatformStream<cudaStream_t>>()
                                  .Arg<ffi::AnyBuffer>().Arg<ffi::AnyBuffer>().Arg<ffi::AnyBuffer>().Arg<ffi::AnyBuffer>().Arg<ffi::AnyBuffer>()
                                  .Arg<ffi::AnyBuffer>().Arg<ffi::AnyBuffer>().Arg<ffi::AnyBuffer>().Arg<ffi::AnyBuffer>().Arg<ffi::AnyBuffer>()
                                  IDOC_RET(3).Attr<int64_t>("problem").Attr<int64_t>("systems").Attr<bool>("has_limits")
                                  .Attr<bool>("has_nub"));
