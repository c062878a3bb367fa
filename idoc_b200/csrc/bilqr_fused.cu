// Fused shared-control batch iLQR (idoc/bilqr.py:120-221) with optional control limits: idoc_bilqr_solve_box / idoc_bilqr_vjp_box.
//
// `systems` copies of one system (same theta, different x0) driven by ONE control sequence are a single system of state
// dimension systems * n whose dynamics / cost derivatives are block diagonal in the state (idoc/bilqr.py:93-117: cost and
// dynamics vmapped over the batch axis, costs summed).  SharedControl<P, NS, BS> presents that lift through the family
// interface of csrc/ilqr.cuh, so the one-kernel iLQR of csrc/ilqr_fused.cuh -- including its control-limited backward
// step (per-step box QP on the SHARED control, Tassa et al. 2014) and the active-set implicit VJP -- runs unchanged on
// it.  This is where BASELINE's north_star places the box step ("projected-Newton step of blqr/bilqr"); the reference
// has no control limits, so the limited path is checked by its own KKT / complementarity conditions, finite
// differences and a bounded quasi-Newton solve of the same problem (tests/test_gpu_bilqr.py); without limits it is
// cross-checked against the multi-kernel loop that the reference goldens pin.
#include "../../include/idoc_b200.h"

#include "ilqr.cuh"
#include "ilqr_fused.cuh"
#include "runtime.cuh"

namespace idoc {

int ilqr_fail(int code, const char* msg);  // ilqr_api.cu: sets idoc_ilqr_last_error()

template <typename P, int NS, int BS>
struct SharedControl {
  __host__ __device__ static int theta_dim(int, int m) { return P::theta_dim(NS, m); }
  template <typename T>
  __device__ static void dyn(int, int m, const T* th, const T* x, const T* u, T* f) {
#pragma unroll
    for (int s = 0; s < BS; s++) P::dyn(NS, m, th, x + s * NS, u, f + s * NS);
  }
  template <typename T>
  __device__ static T cost(int, int m, const T* th, const T* x, const T* u) {
    T c = T(0);
#pragma unroll
    for (int s = 0; s < BS; s++) c += P::cost(NS, m, th, x + s * NS, u);  // bilqr.py:139-143: summed over the systems
    return c;
  }
  template <typename T>
  __device__ static T costf(int, int m, const T* th, const T* x) {
    T c = T(0);
#pragma unroll
    for (int s = 0; s < BS; s++) c += P::costf(NS, m, th, x + s * NS);
    return c;
  }
  // Q', A' block diagonal; M', B' stacked; R', l_u summed (both accumulate in the family interface)
  template <typename T>
  __device__ static void derivs(int, int m, const T* th, const T* x, const T* u, const T* nu, T* lx, T* lu, T* Q, int ldq, T* R,
                                T* M, T* A, int lda, T* B, T* f) {
#pragma unroll
    for (int i = 0; i < NS * BS; i++)
#pragma unroll
      for (int j = 0; j < NS * BS; j++)
        if (i / NS != j / NS) { Q[i * ldq + j] = T(0); A[i * lda + j] = T(0); }
#pragma unroll
    for (int s = 0; s < BS; s++)
      P::derivs(NS, m, th, x + s * NS, u, nu != nullptr ? nu + s * NS : (const T*)nullptr, lx + s * NS, lu,
                Q + (s * NS) * ldq + s * NS, ldq, R, M + s * NS * m, A + (s * NS) * lda + s * NS, lda, B + s * NS * m, f + s * NS);
  }
  template <typename T>
  __device__ static void final_derivs(int, int m, const T* th, const T* x, T* lfx, T* Qf, int ld) {
#pragma unroll
    for (int i = 0; i < NS * BS; i++)
#pragma unroll
      for (int j = 0; j < NS * BS; j++)
        if (i / NS != j / NS) Qf[i * ld + j] = T(0);
#pragma unroll
    for (int s = 0; s < BS; s++) P::final_derivs(NS, m, th, x + s * NS, lfx + s * NS, Qf + (s * NS) * ld + s * NS, ld);
  }
  // theta is shared: its cotangent is the sum over the systems of the per-system pull-backs (csrc/ilqr.cuh, theta kernel)
  template <typename T>
  __device__ static T theta_vjp_step(int, int m, int e, const T* th, const T* x, const T* u, const T* nu, const LeafCot<T>& g) {
    T acc = T(0);
#pragma unroll
    for (int s = 0; s < BS; s++) {
      LeafCot<T> h = g;
      h.gQ = g.gQ + (s * NS) * g.ldq + s * NS; h.gq = g.gq + s * NS; h.gM = g.gM + s * NS * m;
      h.gA = g.gA + (s * NS) * g.lda + s * NS; h.gB = g.gB + s * NS * m; h.gd = g.gd + s * NS;
      acc += P::theta_vjp_step(NS, m, e, th, x + s * NS, u, nu + s * NS, h);
    }
    return acc;
  }
  template <typename T>
  __device__ static T theta_vjp_final(int, int m, int e, const T* th, const T* xT, const T* gQf, int ld, const T* gqf) {
    T acc = T(0);
#pragma unroll
    for (int s = 0; s < BS; s++) acc += P::theta_vjp_final(NS, m, e, th, xT + s * NS, gQf + (s * NS) * ld + s * NS, ld, gqf + s * NS);
    return acc;
  }
};

namespace {

size_t al256(size_t x) { return (x + 255) / 256 * 256; }

struct SolveCall {
  cudaStream_t s; int64_t B; int Th;
  const void *theta, *x0, *u_lo, *u_hi; void *X, *U, *Nu;
  int maxiter; double thres; int use_ls; double ls_lower, ls_upper, ls_rho; int ls_maxiter;
  int32_t *iters, *ls_failed; char* ws;
};
struct VjpCall {
  cudaStream_t s; int64_t B; int Th;
  const void *theta, *x0, *u_lo, *u_hi, *X, *U, *Nu, *Xb, *Ub, *Nub; void *gtheta, *gx0; char* ws;
};

template <typename P, typename T, int NS, int M, int BS, int TD>
int solve_t(const SolveCall& c) {
  using L = SharedControl<P, NS, BS>;
  constexpr int N = NS * BS;
  IlqrFusedArgs<T> f{};
  f.theta = (const T*)c.theta; f.x0 = (const T*)c.x0; f.X = (T*)c.X; f.U = (T*)c.U; f.Nu = (T*)c.Nu; f.scratch = (T*)c.ws;
  f.u_lo = (const T*)c.u_lo; f.u_hi = (const T*)c.u_hi;
  int* ints = (int*)(c.ws + al256(ilqr_fused_scratch_elems(c.Th, N, M) * (size_t)c.B * sizeof(T)));
  f.iters = c.iters ? c.iters : ints; f.ls_failed = c.ls_failed ? c.ls_failed : ints + c.B;
  f.nb = c.B; f.T_ = c.Th; f.theta_dim = P::theta_dim(NS, M); f.maxiter = c.maxiter; f.ls_maxiter = c.ls_maxiter; f.use_ls = c.use_ls;
  f.box_ls = tune_int("IDOC_BOX_LS", 0);
  f.thres = (T)c.thres; f.ls_lower = (T)c.ls_lower; f.ls_upper = (T)c.ls_upper; f.ls_rho = (T)c.ls_rho;
  { ProfScope ps(K_ILQR_FUSED, c.s); ilqr_fused_kernel<L, T, N, M, TD><<<(unsigned)((c.B + 31) / 32), 32, 0, c.s>>>(f); }
  IDOC_CK(cudaGetLastError());
  return IDOC_OK;
}

template <typename P, typename T, int NS, int M, int BS, int TD>
int vjp_t(const VjpCall& c) {
  using L = SharedControl<P, NS, BS>;
  constexpr int N = NS * BS;
  IlqrFusedVjpArgs<T> f{};
  f.theta = (const T*)c.theta; f.x0 = (const T*)c.x0; f.X = (const T*)c.X; f.U = (const T*)c.U; f.Nu = (const T*)c.Nu;
  f.Xb = (const T*)c.Xb; f.Ub = (const T*)c.Ub; f.Nub = (const T*)c.Nub; f.gtheta = (T*)c.gtheta; f.gx0 = (T*)c.gx0;
  f.scratch = (T*)c.ws; f.nb = c.B; f.T_ = c.Th; f.theta_dim = P::theta_dim(NS, M);
  f.u_lo = (const T*)c.u_lo; f.u_hi = (const T*)c.u_hi;
  { ProfScope ps(K_ILQR_FUSED, c.s); ilqr_fused_vjp_kernel<L, T, N, M, TD><<<(unsigned)((c.B + 31) / 32), 32, 0, c.s>>>(f); }
  IDOC_CK(cudaGetLastError());
  return IDOC_OK;
}

// The compiled lifts: (family, n, m, systems).  Thread-per-problem kernels: the lifted state systems * n stays <= 8.
#define IDOC_LIFTS(ROW)                    \
  ROW(IDOC_PROBLEM_PENDULUM, Pendulum, 2, 1, 2, 9) \
  ROW(IDOC_PROBLEM_PENDULUM, Pendulum, 2, 1, 4, 9) \
  ROW(IDOC_PROBLEM_CARTPOLE, Cartpole, 4, 1, 2, 14) \
  ROW(IDOC_PROBLEM_RELU_RNN, ReluRnn, 3, 1, 2, 0)  \
  ROW(IDOC_PROBLEM_RELU_AX, ReluAx, 3, 2, 2, 0)

bool lift_compiled(int problem, int n, int m, int systems) {
#define IDOC_ROW(ID, P, NS, MM, BS, TD) \
  if (problem == ID && n == NS && m == MM && systems == BS) return true;
  IDOC_LIFTS(IDOC_ROW)
#undef IDOC_ROW
  return false;
}

int check_call(int dtype, int problem, int64_t B, int T, int n, int m, int systems, int theta_dim) {
  if (dtype != IDOC_F32 && dtype != IDOC_F64) return ilqr_fail(IDOC_ERR_INVALID, "dtype must be IDOC_F32 or IDOC_F64");
  if (B <= 0 || T <= 0) return ilqr_fail(IDOC_ERR_INVALID, "need B, T > 0");
  if (!lift_compiled(problem, n, m, systems))
    return ilqr_fail(IDOC_ERR_UNSUPPORTED, "no fused shared-control kernel for this (family, n, m, systems): see idoc_bilqr_box_supported");
  if (theta_dim != idoc_ilqr_theta_dim(problem, n, m)) return ilqr_fail(IDOC_ERR_INVALID, "theta_dim does not match the problem");
  return IDOC_OK;
}

}  // namespace
}  // namespace idoc

using namespace idoc;

extern "C" {

int idoc_bilqr_box_supported(int problem, int n, int m, int systems) { return lift_compiled(problem, n, m, systems) ? 1 : 0; }

int idoc_bilqr_solve_box(void* stream, int dtype, int problem, int64_t B, int T, int n, int m, int systems, const void* theta,
                         int theta_dim, const void* x0, const void* u_lo, const void* u_hi, void* X, void* U, void* Nu,
                         int maxiter, double thres, int use_line_search, double ls_lower, double ls_upper, double ls_rho,
                         int ls_maxiter, int32_t* iters, int32_t* ls_failed, void* ws, size_t ws_bytes) {
  int rc = check_call(dtype, problem, B, T, n, m, systems, theta_dim);
  if (rc) return rc;
  if (!theta || !x0 || !X || !U || !Nu || !ws || (u_lo == nullptr) != (u_hi == nullptr))
    return ilqr_fail(IDOC_ERR_INVALID, "null pointer argument (u_lo and u_hi: both or neither)");
  if (ws_bytes < idoc_ilqr_ws_bytes(dtype, B, T, n, m, systems)) return ilqr_fail(IDOC_ERR_WORKSPACE, "ws too small");
  const SolveCall c{(cudaStream_t)stream, B, T, theta, x0, u_lo, u_hi, X, U, Nu, maxiter, thres, use_line_search, ls_lower,
                    ls_upper, ls_rho, ls_maxiter, iters, ls_failed, (char*)ws};
#define IDOC_ROW(ID, P, NS, MM, BS, TD)                                           \
  if (problem == ID && n == NS && m == MM && systems == BS)                \
    return dtype == IDOC_F32 ? solve_t<P, float, NS, MM, BS, TD>(c) : solve_t<P, double, NS, MM, BS, TD>(c);
  IDOC_LIFTS(IDOC_ROW)
#undef IDOC_ROW
  return ilqr_fail(IDOC_ERR_UNSUPPORTED, "unreachable");
}

int idoc_bilqr_vjp_box(void* stream, int dtype, int problem, int64_t B, int T, int n, int m, int systems, const void* theta,
                       int theta_dim, const void* x0, const void* u_lo, const void* u_hi, const void* X, const void* U,
                       const void* Nu, const void* Xb, const void* Ub, const void* Nub, void* gtheta, void* gx0, void* ws,
                       size_t ws_bytes) {
  int rc = check_call(dtype, problem, B, T, n, m, systems, theta_dim);
  if (rc) return rc;
  if (!theta || !x0 || !X || !U || !Nu || !Xb || !Ub || !gtheta || !gx0 || !ws || (u_lo == nullptr) != (u_hi == nullptr))
    return ilqr_fail(IDOC_ERR_INVALID, "null pointer argument (u_lo and u_hi: both or neither)");
  if (ws_bytes < idoc_ilqr_vjp_ws_bytes(dtype, B, T, n, m, systems)) return ilqr_fail(IDOC_ERR_WORKSPACE, "ws too small");
  const VjpCall c{(cudaStream_t)stream, B, T, theta, x0, u_lo, u_hi, X, U, Nu, Xb, Ub, Nub, gtheta, gx0, (char*)ws};
#define IDOC_ROW(ID, P, NS, MM, BS, TD)                                           \
  if (problem == ID && n == NS && m == MM && systems == BS)                \
    return dtype == IDOC_F32 ? vjp_t<P, float, NS, MM, BS, TD>(c) : vjp_t<P, double, NS, MM, BS, TD>(c);
  IDOC_LIFTS(IDOC_ROW)
#undef IDOC_ROW
  return ilqr_fail(IDOC_ERR_UNSUPPORTED, "unreachable");
}

}  // extern "C"
