// C ABI of the iLQR outer loop for compiled problems (include/idoc_b200.h: idoc_ilqr_*).
// Host side only enqueues kernels (plus, when allow_sync != 0, one 4-byte read-back per iteration to stop early);
// the LQR sub-problems go through idoc_lqr_backward / idoc_lqr_vjp of this same library.
#include "../../include/idoc_b200.h"

#include <cstdio>
#include <cstring>

#include "aux_kernels.cuh"
#include "ilqr.cuh"
#include "ilqr_fused.cuh"
#include <type_traits>
#include "runtime.cuh"

// User-defined families (csrc/families/*.cuh, each ending with IDOC_FAMILY(name, Struct)): __graft_entry__.build() lists
// them in gen/user_families.inc as includes (under IDOC_FAMILY_INCLUDES) and as IDOC_USER_FAMILY(id, "name", Struct)
// rows (ids from IDOC_PROBLEM_USER0 on, in file-name order).
// IDOC_FAMILIES_INC: the family table this translation unit is compiled with.  The product library uses the generated
// gen/user_families.inc; a RUN-TIME family plugin (idoc_b200.ilqr.define_family: a family given as source text at run time,
// compiled by nvcc into its own small shared library next to this one) compiles this same file with its one-row table and
// IDOC_PLUGIN_ONLY, which drops the built-in families from the dispatch.  The plugin exports the same idoc_ilqr_* entry
// points (loaded RTLD_LOCAL, linked -Bsymbolic) and calls idoc_lqr_backward / idoc_lqr_vjp of the product library.
#ifndef IDOC_FAMILIES_INC
#define IDOC_FAMILIES_INC "gen/user_families.inc"
#endif
#define IDOC_FAMILY(name, F)
namespace idoc {
#define IDOC_FAMILY_INCLUDES 1
#include IDOC_FAMILIES_INC
#undef IDOC_FAMILY_INCLUDES
}  // namespace idoc

using namespace idoc;

namespace {

template <typename P>
struct is_ad_family : std::false_type {};
template <typename F>
struct is_ad_family<AdFamily<F>> : std::true_type {};

thread_local char g_ierr[256] = "";
int ifail(int code, const char* msg) {
  snprintf(g_ierr, sizeof(g_ierr), "%s", msg);
  return code;
}

}  // namespace
namespace idoc {
int ilqr_fail(int code, const char* msg) { return ifail(code, msg); }  // for bilqr_fused.cu
}  // namespace idoc
namespace {

size_t al(size_t x) { return (x + 255) / 256 * 256; }

struct Layout {  // byte offsets inside the caller's workspace
  size_t Q, q, Qf, qf, M, R, r, A, B, d, lqr_ws, lqr_ws_bytes;
  size_t K, k, dCdc, Xn, Un, c, cn, alpha, ints;                       // solve
  size_t gQ, gq, gQf, gqf, gM, gR, gr, gA, gB, gd, ws2, ws2_bytes;     // vjp
  size_t total;
};

Layout make_layout(int dtype, int64_t B, int T, int n, int m, bool vjp) {  // n = lifted state dimension
  const size_t sz = dtype == IDOC_F32 ? 4 : 8, BT = (size_t)B * T;
  Layout L{};
  size_t o = 0;
  auto take = [&](size_t bytes) {
    size_t r = o;
    o += al(bytes);
    return r;
  };
  L.Q = take(BT * n * n * sz); L.q = take(BT * n * sz); L.Qf = take((size_t)B * n * n * sz); L.qf = take((size_t)B * n * sz);
  L.M = take(BT * n * m * sz); L.R = take(BT * m * m * sz); L.r = take(BT * m * sz); L.A = take(BT * n * n * sz);
  L.B = take(BT * n * m * sz); L.d = take(BT * n * sz);
  L.lqr_ws_bytes = idoc_lqr_ws_bytes(dtype, B, T, n, m);
  L.lqr_ws = take(L.lqr_ws_bytes);
  if (!vjp) {
    L.K = take(BT * m * n * sz); L.k = take(BT * m * sz); L.dCdc = take((size_t)B * 2 * sz);
    L.Xn = take(BT * n * sz); L.Un = take(BT * m * sz);
    L.c = take((size_t)B * sz); L.cn = take((size_t)B * sz); L.alpha = take((size_t)B * sz);
    L.ints = take(((size_t)B * 5 + 4) * sizeof(int));
  } else {
    L.gQ = take(BT * n * n * sz); L.gq = take(BT * n * sz); L.gQf = take((size_t)B * n * n * sz); L.gqf = take((size_t)B * n * sz);
    L.gM = take(BT * n * m * sz); L.gR = take(BT * m * m * sz); L.gr = take(BT * m * sz); L.gA = take(BT * n * n * sz);
    L.gB = take(BT * n * m * sz); L.gd = take(BT * n * sz);
    L.ws2_bytes = idoc_lqr_vjp_ws_bytes(dtype, B, T, n, m);
    L.ws2 = take(L.ws2_bytes);
  }
  L.total = o;
  return L;
}

__global__ void fill_int_kernel(int* p, long long n, int v) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) p[i] = v;
}

template <typename T>
void bind_leaves(IlqrArgs<T>& a, char* ws, const Layout& L) {
  a.Q = (T*)(ws + L.Q); a.q = (T*)(ws + L.q); a.Qf = (T*)(ws + L.Qf); a.qf = (T*)(ws + L.qf); a.Mx = (T*)(ws + L.M);
  a.R = (T*)(ws + L.R); a.r = (T*)(ws + L.r); a.A = (T*)(ws + L.A); a.B = (T*)(ws + L.B); a.d = (T*)(ws + L.d);
}

// One-kernel iLQR for the small single systems that have a fused instantiation (ilqr_fused.cuh); IDOC_ILQR_FUSED=0
// forces the multi-kernel loop below (the two paths are compared in tests/test_gpu_ilqr.py).
template <typename P, typename T, int N, int M, int TD>
int launch_fused(cudaStream_t s, int64_t B, int Th, const void* theta, const void* x0, void* X, void* U, void* Nu, int maxiter,
                 double thres, int use_ls, double ls_lower, double ls_upper, double ls_rho, int ls_maxiter, int32_t* iters_out,
                 int32_t* ls_failed_out, char* ws, const void* u_lo, const void* u_hi) {
  IlqrFusedArgs<T> f{};
  f.theta = (const T*)theta; f.x0 = (const T*)x0; f.X = (T*)X; f.U = (T*)U; f.Nu = (T*)Nu; f.scratch = (T*)ws;
  f.u_lo = (const T*)u_lo; f.u_hi = (const T*)u_hi;
  int* ints = (int*)(ws + al(ilqr_fused_scratch_elems(Th, N, M) * (size_t)B * sizeof(T)));
  f.iters = iters_out ? iters_out : ints; f.ls_failed = ls_failed_out ? ls_failed_out : ints + B;
  f.nb = B; f.T_ = Th; f.theta_dim = P::theta_dim(N, M); f.maxiter = maxiter; f.ls_maxiter = ls_maxiter; f.use_ls = use_ls;
  f.box_ls = tune_int("IDOC_BOX_LS", 0);  // measured (profiles/r2_box_line_search.jsonl): the monotone rules end in far worse optima on the cart-pole
  f.thres = (T)thres; f.ls_lower = (T)ls_lower; f.ls_upper = (T)ls_upper; f.ls_rho = (T)ls_rho;
  { ProfScope ps(K_ILQR_FUSED, s); ilqr_fused_kernel<P, T, N, M, TD><<<(unsigned)((B + 31) / 32), 32, 0, s>>>(f); }
  IDOC_CK(cudaGetLastError());
  return IDOC_OK;
}

template <typename P, typename T>
int solve_impl(cudaStream_t s, int dtype, int64_t B, int Th, int nsys, int m, int bs, const void* theta, const void* x0, void* X,
               void* U, void* Nu, int maxiter, double thres, int use_ls, double ls_lower, double ls_upper, double ls_rho,
               int ls_maxiter, int32_t* iters_out, int32_t* ls_failed_out, char* ws, int allow_sync, const void* u_lo = nullptr,
               const void* u_hi = nullptr) {
  const int n = nsys * bs;  // lifted
  if (bs == 1 && (u_lo != nullptr || tune_int("IDOC_ILQR_FUSED", 1))) {
#define IDOC_FUSED(NN, MM, TD)                                                                                              \
  if (nsys == NN && m == MM)                                                                                                \
    return launch_fused<P, T, NN, MM, TD>(s, B, Th, theta, x0, X, U, Nu, maxiter, thres, use_ls, ls_lower, ls_upper, ls_rho, \
                                          ls_maxiter, iters_out, ls_failed_out, ws, u_lo, u_hi);
    if constexpr (std::is_same<P, Pendulum>::value) { IDOC_FUSED(2, 1, 9) }
    if constexpr (is_ad_family<P>::value) { IDOC_FUSED(P::N, P::M, P::TD) }  // PendulumAd, Cartpole, user families
    if constexpr (std::is_same<P, ReluRnn>::value) { IDOC_FUSED(3, 1, 0) IDOC_FUSED(4, 2, 0) }
#undef IDOC_FUSED
  }
  if (u_lo != nullptr) return ifail(IDOC_ERR_UNSUPPORTED, "control limits: single systems with a fused kernel only");
  const Layout L = make_layout(dtype, B, Th, n, m, false);
  IlqrArgs<T> a{};
  a.theta = (const T*)theta; a.x0 = (const T*)x0; a.X = (T*)X; a.U = (T*)U;
  a.Xn = (T*)(ws + L.Xn); a.Un = (T*)(ws + L.Un); a.c = (T*)(ws + L.c); a.cn = (T*)(ws + L.cn); a.alpha = (T*)(ws + L.alpha);
  int* ints = (int*)(ws + L.ints);
  a.ls_it = ints; a.searching = ints + B; a.active = ints + 2 * B; a.iters = ints + 3 * B; a.ls_failed = ints + 4 * B;
  a.n_active = ints + 5 * B;
  bind_leaves(a, ws, L);
  a.K = (const T*)(ws + L.K); a.k = (const T*)(ws + L.k); a.dCdc = (const T*)(ws + L.dCdc);
  a.nb = B; a.T_ = Th; a.n = nsys; a.m = m; a.bs = bs; a.theta_dim = P::theta_dim(nsys, m);
  a.thres = (T)thres; a.ls_lower = (T)ls_lower; a.ls_upper = (T)ls_upper; a.ls_rho = (T)ls_rho;
  a.ls_maxiter = ls_maxiter; a.use_ls = use_ls;
  const unsigned gb = (unsigned)((B + 127) / 128), gbt = (unsigned)(((long long)B * Th + 127) / 128);

  IDOC_CK(cudaMemsetAsync(ints, 0, ((size_t)B * 5 + 4) * sizeof(int), s));
  { ProfScope ps(K_ILQR, s); fill_int_kernel<<<gb, 128, 0, s>>>(a.active, B, 1); }
  { ProfScope ps(K_ILQR, s); ilqr_simulate_kernel<P, T><<<gb, 128, 0, s>>>(a, 0); }  // c_old         ilqr.py:255-256
  IDOC_CK(cudaGetLastError());
  int n_active_host = 1;
  for (int it = 0; it < maxiter; it++) {
    { ProfScope ps(K_ILQR, s); ilqr_begin_iter_kernel<T><<<gb, 128, 0, s>>>(a); }
    { ProfScope ps(K_ILQR, s); ilqr_linearize_kernel<P, T><<<gbt, 128, 0, s>>>(a, 0); }   // local form   ilqr.py:230
    IDOC_CK(cudaGetLastError());
    int rc = idoc_lqr_backward(s, dtype, B, Th, n, m, a.Q, a.q, a.Qf, a.qf, a.Mx, a.R, a.r, a.A, a.B, a.d, (void*)a.K,
                               (void*)a.k, (void*)a.dCdc, nullptr, ws + L.lqr_ws, L.lqr_ws_bytes);  // ilqr.py:231-233
    if (rc) return ifail(rc, idoc_last_error());
    const int rounds = use_ls ? ls_maxiter : 1;
    for (int r = 0; r < rounds; r++) {
      { ProfScope ps(K_ILQR, s); ilqr_update_kernel<P, T><<<gb, 128, 0, s>>>(a); }         // f(alpha)     ilqr.py:235-236
      { ProfScope ps(K_ILQR, s); ilqr_ls_kernel<T><<<gb, 128, 0, s>>>(a); }                // line_search.py:12-27
    }
    { ProfScope ps(K_ILQR, s); ilqr_commit_kernel<T><<<(unsigned)B, 64, 0, s>>>(a); }      // ilqr.py:246-252
    IDOC_CK(cudaGetLastError());
    if (allow_sync) {
      IDOC_CK(cudaMemcpyAsync(&n_active_host, a.n_active, sizeof(int), cudaMemcpyDeviceToHost, s));
      IDOC_CK(cudaStreamSynchronize(s));
      if (n_active_host == 0) break;
    }
  }
  // global approximation at the solution and its co-states                                 ilqr.py:266-267
  { ProfScope ps(K_ILQR, s); ilqr_linearize_kernel<P, T><<<gbt, 128, 0, s>>>(a, 1); }
  FixupArgs<T> f{};
  f.Q = a.Q; f.q = a.q; f.Qf = a.Qf; f.qf = a.qf; f.Mx = a.Mx; f.A = a.A; f.X = a.X; f.U = a.U; f.Nu = (T*)Nu;
  f.status = nullptr; f.nb = B; f.T_ = Th; f.n = n; f.m = m;
  { ProfScope ps(K_FIXUP, s); lqr_costate_fixup_kernel<T><<<(unsigned)((B + 63) / 64), 64, 0, s>>>(f); }
  IDOC_CK(cudaGetLastError());
  if (iters_out) IDOC_CK(cudaMemcpyAsync(iters_out, a.iters, (size_t)B * sizeof(int), cudaMemcpyDeviceToDevice, s));
  if (ls_failed_out) IDOC_CK(cudaMemcpyAsync(ls_failed_out, a.ls_failed, (size_t)B * sizeof(int), cudaMemcpyDeviceToDevice, s));
  return IDOC_OK;
}

template <typename P, typename T>
int vjp_impl(cudaStream_t s, int dtype, int64_t B, int Th, int nsys, int m, int bs, const void* theta, const void* x0, const void* X,
             const void* U, const void* Nu, const void* Xb, const void* Ub, const void* Nub, void* gtheta, void* gx0,
             char* ws, const void* u_lo = nullptr, const void* u_hi = nullptr) {
  const int n = nsys * bs;  // lifted
  if (bs == 1 && (u_lo != nullptr || tune_int("IDOC_ILQR_FUSED", 1))) {
#define IDOC_FUSED(NN, MM, TD)                                                                                     \
  if (nsys == NN && m == MM) {                                                                                     \
    IlqrFusedVjpArgs<T> f{};                                                                                       \
    f.theta = (const T*)theta; f.x0 = (const T*)x0; f.X = (const T*)X; f.U = (const T*)U; f.Nu = (const T*)Nu;     \
    f.Xb = (const T*)Xb; f.Ub = (const T*)Ub; f.Nub = (const T*)Nub; f.gtheta = (T*)gtheta; f.gx0 = (T*)gx0;       \
    f.scratch = (T*)ws; f.nb = B; f.T_ = Th; f.theta_dim = P::theta_dim(NN, MM);                                   \
    f.u_lo = (const T*)u_lo; f.u_hi = (const T*)u_hi;                                                              \
    { ProfScope ps(K_ILQR_FUSED, s); ilqr_fused_vjp_kernel<P, T, NN, MM, TD><<<(unsigned)((B + 31) / 32), 32, 0, s>>>(f); } \
    IDOC_CK(cudaGetLastError());                                                                                   \
    return IDOC_OK;                                                                                                \
  }
    if constexpr (std::is_same<P, Pendulum>::value) { IDOC_FUSED(2, 1, 9) }
    if constexpr (is_ad_family<P>::value) { IDOC_FUSED(P::N, P::M, P::TD) }  // PendulumAd, Cartpole, user families
    if constexpr (std::is_same<P, ReluRnn>::value) { IDOC_FUSED(3, 1, 0) IDOC_FUSED(4, 2, 0) }
#undef IDOC_FUSED
  }
  if (u_lo != nullptr) return ifail(IDOC_ERR_UNSUPPORTED, "control limits: single systems with a fused kernel only");
  const Layout L = make_layout(dtype, B, Th, n, m, true);
  IlqrArgs<T> a{};
  a.theta = (const T*)theta; a.x0 = (const T*)x0; a.X = (T*)X; a.U = (T*)U; a.Nu = (const T*)Nu;
  bind_leaves(a, ws, L);
  a.nb = B; a.T_ = Th; a.n = nsys; a.m = m; a.bs = bs; a.theta_dim = P::theta_dim(nsys, m);
  const unsigned gbt = (unsigned)(((long long)B * Th + 127) / 128);
  // blocks of the adjoint system: Hessian of the Lagrangian at the solution
  { ProfScope ps(K_ILQR, s); ilqr_linearize_kernel<P, T><<<gbt, 128, 0, s>>>(a, 2); }
  IDOC_CK(cudaGetLastError());
  int rc = idoc_lqr_backward(s, dtype, B, Th, n, m, a.Q, a.q, a.Qf, a.qf, a.Mx, a.R, a.r, a.A, a.B, a.d, nullptr, nullptr,
                             nullptr, nullptr, ws + L.lqr_ws, L.lqr_ws_bytes);
  if (rc) return ifail(rc, idoc_last_error());
  rc = idoc_lqr_vjp(s, dtype, B, Th, n, m, a.Mx, a.A, a.B, x0, X, U, Nu, ws + L.lqr_ws, Xb, Ub, Nub, ws + L.gQ, ws + L.gq,
                    ws + L.gQf, ws + L.gqf, ws + L.gM, ws + L.gR, ws + L.gr, ws + L.gA, ws + L.gB, ws + L.gd, gx0,
                    ws + L.ws2, L.ws2_bytes, 0);
  if (rc) return ifail(rc, idoc_last_error());
  { ProfScope ps(K_ILQR, s);
    ilqr_theta_vjp_kernel<P, T><<<(unsigned)B, 128, 0, s>>>(a, (const T*)(ws + L.gQ), (const T*)(ws + L.gq),
                                                            (const T*)(ws + L.gQf), (const T*)(ws + L.gqf),
                                                            (const T*)(ws + L.gM), (const T*)(ws + L.gR),
                                                            (const T*)(ws + L.gr), (const T*)(ws + L.gA),
                                                            (const T*)(ws + L.gB), (const T*)(ws + L.gd), (T*)gtheta); }
  IDOC_CK(cudaGetLastError());
  return IDOC_OK;
}

int check(int dtype, int problem, int64_t B, int T, int n, int m, int bs, int theta_dim) {
  if (dtype != IDOC_F32 && dtype != IDOC_F64) return ifail(IDOC_ERR_INVALID, "dtype must be IDOC_F32 or IDOC_F64");
  if (B <= 0 || T <= 0 || n <= 0 || m <= 0 || bs <= 0 || n * bs > ILQR_MAXD || m > ILQR_MAXD)
    return ifail(IDOC_ERR_INVALID, "need B, T > 0, 0 < m <= 32 and 0 < n * systems <= 32");
#ifdef IDOC_PLUGIN_ONLY
  if (problem < IDOC_PROBLEM_USER0) return ifail(IDOC_ERR_UNSUPPORTED, "family plugin: user family ids only");
#endif
  if (problem == IDOC_PROBLEM_RELU_RNN || problem == IDOC_PROBLEM_RELU_AX) {
    if (theta_dim != ReluRnn::theta_dim(n, m)) return ifail(IDOC_ERR_INVALID, "theta_dim does not match the problem");
  } else if (problem == IDOC_PROBLEM_PENDULUM || problem == IDOC_PROBLEM_PENDULUM_AD) {
    if (n != 2 || m != 1 || theta_dim != 9) return ifail(IDOC_ERR_INVALID, "pendulum: n = 2, m = 1, theta_dim = 9");
  } else if (problem == IDOC_PROBLEM_CARTPOLE) {
    if (n != 4 || m != 1 || theta_dim != 14) return ifail(IDOC_ERR_INVALID, "cartpole: n = 4, m = 1, theta_dim = 14");
  } else {
#define IDOC_USER_FAMILY(ID, NAME, F)                                                                         \
  if (problem == ID) {                                                                                        \
    if (n != F::N || m != F::M || theta_dim != F::TD)                                                         \
      return ifail(IDOC_ERR_INVALID, NAME ": n, m, theta_dim must match the family's N, M, TD");              \
    return IDOC_OK;                                                                                           \
  }
#include IDOC_FAMILIES_INC
#undef IDOC_USER_FAMILY
    return ifail(IDOC_ERR_UNSUPPORTED, "unknown problem id");
  }
  return IDOC_OK;
}

}  // namespace

template <typename P, typename T>
static int linearize_impl(cudaStream_t s, int64_t B, int Th, int nsys, int m, int bs, const void* theta, const void* x0, const void* X,
                          const void* U, const void* Nu, int mode, void* Q, void* q, void* Qf, void* qf, void* M, void* R,
                          void* r, void* A, void* Bm, void* d) {
  IlqrArgs<T> a{};
  a.theta = (const T*)theta; a.x0 = (const T*)x0; a.X = (T*)X; a.U = (T*)U; a.Nu = (const T*)Nu;
  a.Q = (T*)Q; a.q = (T*)q; a.Qf = (T*)Qf; a.qf = (T*)qf; a.Mx = (T*)M; a.R = (T*)R; a.r = (T*)r; a.A = (T*)A; a.B = (T*)Bm;
  a.d = (T*)d;
  a.nb = B; a.T_ = Th; a.n = nsys; a.m = m; a.bs = bs; a.theta_dim = P::theta_dim(nsys, m);
  { ProfScope ps(K_ILQR, s); ilqr_linearize_kernel<P, T><<<(unsigned)(((long long)B * Th + 127) / 128), 128, 0, s>>>(a, mode); }
  IDOC_CK(cudaGetLastError());
  return IDOC_OK;
}


extern "C" {

const char* idoc_ilqr_last_error(void) { return g_ierr; }

int idoc_ilqr_theta_dim(int problem, int n, int m) {
  if (problem == IDOC_PROBLEM_RELU_RNN || problem == IDOC_PROBLEM_RELU_AX) return ReluRnn::theta_dim(n, m);
  if (problem == IDOC_PROBLEM_PENDULUM || problem == IDOC_PROBLEM_PENDULUM_AD) return 9;
  if (problem == IDOC_PROBLEM_CARTPOLE) return 14;
#define IDOC_USER_FAMILY(ID, NAME, F) \
  if (problem == ID) return F::TD;
#include IDOC_FAMILIES_INC
#undef IDOC_USER_FAMILY
  return -1;
}

// user-defined families compiled into this library (csrc/families/): count, and name / id / dimensions of the i-th
int idoc_ilqr_family_count(void) {
  int c = 0;
#define IDOC_USER_FAMILY(ID, NAME, F) c++;
#include IDOC_FAMILIES_INC
#undef IDOC_USER_FAMILY
  return c;
}
int idoc_ilqr_family_info(int i, const char** name, int* problem, int* n, int* m, int* theta_dim) {
  int c = 0;
#define IDOC_USER_FAMILY(ID, NAME, F)                                 \
  if (c++ == i) {                                                     \
    if (name) *name = NAME;                                           \
    if (problem) *problem = ID;                                       \
    if (n) *n = F::N;                                                 \
    if (m) *m = F::M;                                                 \
    if (theta_dim) *theta_dim = F::TD;                                \
    return IDOC_OK;                                                   \
  }
#include IDOC_FAMILIES_INC
#undef IDOC_USER_FAMILY
  return ifail(IDOC_ERR_INVALID, "family index out of range");
}

size_t idoc_ilqr_ws_bytes(int dtype, int64_t B, int T, int n, int m, int systems) {
  return make_layout(dtype, B, T, n * systems, m, false).total;
}
size_t idoc_ilqr_vjp_ws_bytes(int dtype, int64_t B, int T, int n, int m, int systems) {
  return make_layout(dtype, B, T, n * systems, m, true).total;
}

#define IDOC_FAMILY_DISPATCH_DEF 1  // defines IDOC_DISPATCH_USER(CALL): one IDOC_USER_FAMILY_CALL per user family
#include IDOC_FAMILIES_INC
#undef IDOC_FAMILY_DISPATCH_DEF
#define IDOC_USER_FAMILY_CALL(CALL, ID, F)                     \
  if (problem == ID) {                                         \
    if (dtype == IDOC_F32) return CALL(AdFamily<F>, float);    \
    return CALL(AdFamily<F>, double);                          \
  }
#ifdef IDOC_PLUGIN_ONLY
#define IDOC_DISPATCH(CALL) \
  IDOC_DISPATCH_USER(CALL)  \
  return ifail(IDOC_ERR_UNSUPPORTED, "family plugin: user family ids only");
#else
#define IDOC_DISPATCH(CALL)                                 \
  IDOC_DISPATCH_USER(CALL)                                  \
  if (problem == IDOC_PROBLEM_RELU_RNN) {                   \
    if (dtype == IDOC_F32) return CALL(ReluRnn, float);     \
    return CALL(ReluRnn, double);                           \
  }                                                         \
  if (problem == IDOC_PROBLEM_RELU_AX) {                    \
    if (dtype == IDOC_F32) return CALL(ReluAx, float);      \
    return CALL(ReluAx, double);                            \
  }                                                         \
  if (problem == IDOC_PROBLEM_CARTPOLE) {                   \
    if (dtype == IDOC_F32) return CALL(Cartpole, float);    \
    return CALL(Cartpole, double);                          \
  }                                                         \
  if (problem == IDOC_PROBLEM_PENDULUM_AD) {                \
    if (dtype == IDOC_F32) return CALL(PendulumAd, float);  \
    return CALL(PendulumAd, double);                        \
  }                                                         \
  if (dtype == IDOC_F32) return CALL(Pendulum, float);      \
  return CALL(Pendulum, double);
#endif

int idoc_ilqr_solve(void* stream, int dtype, int problem, int64_t B, int T, int n, int m, int systems, const void* theta,
                    int theta_dim, const void* x0, void* X, void* U, void* Nu, int maxiter, double thres,
                    int use_line_search, double ls_lower, double ls_upper, double ls_rho, int ls_maxiter, int32_t* iters,
                    int32_t* ls_failed, void* ws, size_t ws_bytes, int allow_sync) {
  int rc = check(dtype, problem, B, T, n, m, systems, theta_dim);
  if (rc) return rc;
  if (!theta || !x0 || !X || !U || !Nu || !ws) return ifail(IDOC_ERR_INVALID, "null pointer argument");
  if (ws_bytes < idoc_ilqr_ws_bytes(dtype, B, T, n, m, systems)) return ifail(IDOC_ERR_WORKSPACE, "ws too small");
  cudaStream_t s = (cudaStream_t)stream;
#define IDOC_CALL(P, TT)                                                                                                 \
  solve_impl<P, TT>(s, dtype, B, T, n, m, systems, theta, x0, X, U, Nu, maxiter, thres, use_line_search, ls_lower, ls_upper, \
                    ls_rho, ls_maxiter, iters, ls_failed, (char*)ws, allow_sync)
  IDOC_DISPATCH(IDOC_CALL)
#undef IDOC_CALL
}

int idoc_ilqr_linearize(void* stream, int dtype, int problem, int64_t B, int T, int n, int m, int systems, const void* theta,
                        int theta_dim, const void* x0, const void* X, const void* U, const void* Nu, int mode, void* Q,
                        void* q, void* Qf, void* qf, void* M, void* R, void* r, void* A, void* Bm, void* d) {
  int rc = check(dtype, problem, B, T, n, m, systems, theta_dim);
  if (rc) return rc;
  if (mode < 0 || mode > 2 || (mode == 2 && !Nu)) return ifail(IDOC_ERR_INVALID, "mode must be 0, 1 or 2 (2 needs Nu)");
  if (!theta || !x0 || !X || !U || !Q || !q || !Qf || !qf || !M || !R || !r || !A || !Bm || !d)
    return ifail(IDOC_ERR_INVALID, "null pointer argument");
  cudaStream_t s = (cudaStream_t)stream;
#define IDOC_CALL(P, TT) \
  linearize_impl<P, TT>(s, B, T, n, m, systems, theta, x0, X, U, Nu, mode, Q, q, Qf, qf, M, R, r, A, Bm, d)
  IDOC_DISPATCH(IDOC_CALL)
#undef IDOC_CALL
}

int idoc_ilqr_vjp(void* stream, int dtype, int problem, int64_t B, int T, int n, int m, int systems, const void* theta,
                  int theta_dim, const void* x0, const void* X, const void* U, const void* Nu, const void* Xb,
                  const void* Ub, const void* Nub, void* gtheta, void* gx0, void* ws, size_t ws_bytes) {
  int rc = check(dtype, problem, B, T, n, m, systems, theta_dim);
  if (rc) return rc;
  if (!theta || !x0 || !X || !U || !Nu || !Xb || !Ub || !gtheta || !gx0 || !ws) return ifail(IDOC_ERR_INVALID, "null pointer argument");
  if (ws_bytes < idoc_ilqr_vjp_ws_bytes(dtype, B, T, n, m, systems)) return ifail(IDOC_ERR_WORKSPACE, "ws too small");
  cudaStream_t s = (cudaStream_t)stream;
#define IDOC_CALL(P, TT) \
  vjp_impl<P, TT>(s, dtype, B, T, n, m, systems, theta, x0, X, U, Nu, Xb, Ub, Nub, gtheta, gx0, (char*)ws)
  IDOC_DISPATCH(IDOC_CALL)
#undef IDOC_CALL
}

int idoc_ilqr_solve_box(void* stream, int dtype, int problem, int64_t B, int T, int n, int m, const void* theta, int theta_dim,
                        const void* x0, const void* u_lo, const void* u_hi, void* X, void* U, void* Nu, int maxiter,
                        double thres, int use_line_search, double ls_lower, double ls_upper, double ls_rho, int ls_maxiter,
                        int32_t* iters, int32_t* ls_failed, void* ws, size_t ws_bytes) {
  int rc = check(dtype, problem, B, T, n, m, 1, theta_dim);
  if (rc) return rc;
  if (!theta || !x0 || !X || !U || !Nu || !ws || !u_lo || !u_hi) return ifail(IDOC_ERR_INVALID, "null pointer argument");
  if (ws_bytes < idoc_ilqr_ws_bytes(dtype, B, T, n, m, 1)) return ifail(IDOC_ERR_WORKSPACE, "ws too small");
  cudaStream_t s = (cudaStream_t)stream;
#define IDOC_CALL(P, TT)                                                                                              \
  solve_impl<P, TT>(s, dtype, B, T, n, m, 1, theta, x0, X, U, Nu, maxiter, thres, use_line_search, ls_lower, ls_upper, \
                    ls_rho, ls_maxiter, iters, ls_failed, (char*)ws, 0, u_lo, u_hi)
  IDOC_DISPATCH(IDOC_CALL)
#undef IDOC_CALL
}

int idoc_ilqr_vjp_box(void* stream, int dtype, int problem, int64_t B, int T, int n, int m, const void* theta, int theta_dim,
                      const void* x0, const void* u_lo, const void* u_hi, const void* X, const void* U, const void* Nu,
                      const void* Xb, const void* Ub, const void* Nub, void* gtheta, void* gx0, void* ws, size_t ws_bytes) {
  int rc = check(dtype, problem, B, T, n, m, 1, theta_dim);
  if (rc) return rc;
  if (!theta || !x0 || !X || !U || !Nu || !Xb || !Ub || !gtheta || !gx0 || !ws || !u_lo || !u_hi)
    return ifail(IDOC_ERR_INVALID, "null pointer argument");
  if (ws_bytes < idoc_ilqr_vjp_ws_bytes(dtype, B, T, n, m, 1)) return ifail(IDOC_ERR_WORKSPACE, "ws too small");
  cudaStream_t s = (cudaStream_t)stream;
#define IDOC_CALL(P, TT) \
  vjp_impl<P, TT>(s, dtype, B, T, n, m, 1, theta, x0, X, U, Nu, Xb, Ub, Nub, gtheta, gx0, (char*)ws, u_lo, u_hi)
  IDOC_DISPATCH(IDOC_CALL)
#undef IDOC_CALL
}

}  // extern "C"
