// iLQR outer loop for compiled problems (reference: idoc/ilqr.py:94-278, idoc/line_search.py:7-42).
//
// The reference's Problem holds Python callables differentiated by JAX; a kernel cannot call those, so a problem
// is a device functor (cost / dynamics and their derivatives, written once per problem family) selected by id.
// One iLQR iteration = linearise (local form, ilqr.py:143-146) -> Riccati sweep (the LQR kernels, lqr.py:62-107)
// -> backtracking line search over closed-loop rollouts (ilqr.py:196-219, line_search.py:9-40), all on the device,
// per-problem step sizes and convergence flags; the host only enqueues.
//
// Kernels here are shape-generic (runtime n, m <= 32): thread per (problem, timestep) for the linearisation,
// thread per problem for the sequential rollouts.
#pragma once
#include "common.cuh"

namespace idoc {

constexpr int ILQR_MAXD = 32;

// cotangents of the global LQR-approximation leaves of ONE system at one (problem, timestep), as the LQR VJP wrote them
// (gQ = sym(ux (x) x), gq = ux, gR = sym(uu (x) u), gr = uu, gM, gA = nu (x) ux + unu (x) x, gB = nu (x) uu + unu (x) u, gd = unu)
template <typename T>
struct LeafCot {
  const T *gQ, *gq, *gR, *gr, *gM, *gA, *gB, *gd;
  int ldq, lda;  // row strides of the (lifted) gQ and gA blocks
};

// msin / mcos: one name for plain scalars and for the AD types of ad.cuh (families are written once over the scalar type)
__device__ __forceinline__ double msin(double x) { return ::sin(x); }
__device__ __forceinline__ double mcos(double x) { return ::cos(x); }
__device__ __forceinline__ float msin(float x) { return ::sinf(x); }
__device__ __forceinline__ float mcos(float x) { return ::cosf(x); }
}  // namespace idoc
#include "ad.cuh"
namespace idoc {

// ---------------------------------------------------------------------------------------------------- problems
// A problem family is a struct of static device functions on ONE system (state n, control m):
//   dyn, cost, costf;
//   derivs(n, m, th, x, u, nu, lx, lu+=, Q (ld), R+=, M rows, A (ld), B rows, f): cost gradient / Hessians and
//       dynamics Jacobians written straight into the (strided) LQR leaves; if nu != nullptr the dynamics
//       curvature sum_i nu_i d2f_i is added to Q, R, M (Hessian of the Lagrangian, needed by the implicit VJP);
//   final_derivs; theta_vjp_step / theta_vjp_final: cotangent of theta element e given the cotangents of the
//       GLOBAL LQR-approximation leaves (the leaves as functions of theta at fixed (x, u)).
// With bs > 1 systems sharing one control sequence (bilqr, idoc/bilqr.py) the kernels call these per system and
// assemble the block-diagonal lift (A', Q' block diagonal; B', M' stacked; R', r' summed over systems).

// relu recurrent systems of the reference's tests.  PRE = false: f = A relu(x) + B u + 0.5, cost weights 1e-4
// (tests/test_ilqr.py:14-64);  PRE = true: f = relu(A x) + B u + 0.5, cost weights 1e-6 (tests/test_bilqr.py:39-83).
// l = 1/2 x'Qx + q'x + w u'Ru - w (1'u)(1'x) + w r'u,  lf = 1/2 x'Qf x;  theta = (Q, q, Qf, R, r, A, B) flattened.
template <bool PRE>
struct ReluSys {
  __host__ __device__ static int theta_dim(int n, int m) { return 3 * n * n + n + m * m + m + n * m; }
  struct Th {
    int oQ, oq, oQf, oR, or_, oA, oB;
    __host__ __device__ Th(int n, int m) {
      oQ = 0; oq = oQ + n * n; oQf = oq + n; oR = oQf + n * n; or_ = oR + m * m; oA = or_ + m; oB = oA + n * n;
    }
  };
  template <typename T>
  __device__ static T wgt() { return PRE ? T(1e-6) : T(1e-4); }
  // activation derivative pattern: g[j] multiplies column j (PRE = false) or row i (PRE = true) of A
  template <typename T>
  __device__ static T gate(int n, int m, const T* th, const T* x, int i) {
    if (!PRE) return x[i] > T(0) ? T(1) : T(0);
    const Th o(n, m);
    T z = T(0);
    for (int j = 0; j < n; j++) z += th[o.oA + i * n + j] * x[j];
    return z > T(0) ? T(1) : T(0);
  }
  template <typename T>
  __device__ static void dyn(int n, int m, const T* th, const T* x, const T* u, T* f) {
    const Th o(n, m);
    for (int i = 0; i < n; i++) {
      T s = T(0), bu = T(0);
      for (int j = 0; j < n; j++) s += th[o.oA + i * n + j] * (PRE ? x[j] : (x[j] > T(0) ? x[j] : T(0)));
      for (int j = 0; j < m; j++) bu += th[o.oB + i * m + j] * u[j];
      f[i] = (PRE ? (s > T(0) ? s : T(0)) : s) + bu + T(0.5);
    }
  }
  template <typename T>
  __device__ static T cost(int n, int m, const T* th, const T* x, const T* u) {
    const Th o(n, m);
    T lQ = T(0), lq = T(0), lR = T(0), lr = T(0), sx = T(0), su = T(0);
    for (int i = 0; i < n; i++) {
      T s = T(0);
      for (int j = 0; j < n; j++) s += th[o.oQ + i * n + j] * x[j];
      lQ += s * x[i];
      lq += th[o.oq + i] * x[i];
      sx += x[i];
    }
    for (int i = 0; i < m; i++) {
      T s = T(0);
      for (int j = 0; j < m; j++) s += th[o.oR + i * m + j] * u[j];
      lR += s * u[i];
      lr += th[o.or_ + i] * u[i];
      su += u[i];
    }
    return T(0.5) * lQ + lq + wgt<T>() * (lR - su * sx + lr);
  }
  template <typename T>
  __device__ static T costf(int n, int m, const T* th, const T* x) {
    const Th o(n, m);
    T l = T(0);
    for (int i = 0; i < n; i++) {
      T s = T(0);
      for (int j = 0; j < n; j++) s += th[o.oQf + i * n + j] * x[j];
      l += s * x[i];
    }
    return T(0.5) * l;
  }
  template <typename T>
  __device__ static void derivs(int n, int m, const T* th, const T* x, const T* u, const T* nu, T* lx, T* lu, T* Q, int ldq,
                                T* R, T* M, T* A, int lda, T* B, T* f) {
    (void)nu;  // piecewise-linear dynamics: no curvature
    const Th o(n, m);
    const T w = wgt<T>();
    T sx = T(0), su = T(0);
    for (int i = 0; i < n; i++) sx += x[i];
    for (int i = 0; i < m; i++) su += u[i];
    for (int i = 0; i < n; i++) {
      T s = th[o.oq + i] - w * su;
      const T gi = PRE ? gate(n, m, th, x, i) : T(0);
      for (int j = 0; j < n; j++) {
        const T h = T(0.5) * (th[o.oQ + i * n + j] + th[o.oQ + j * n + i]);
        Q[i * ldq + j] = h;
        s += h * x[j];
        A[i * lda + j] = (PRE ? gi : (x[j] > T(0) ? T(1) : T(0))) * th[o.oA + i * n + j];
      }
      lx[i] = s;
      for (int j = 0; j < m; j++) {
        M[i * m + j] = -w;
        B[i * m + j] = th[o.oB + i * m + j];
      }
    }
    for (int i = 0; i < m; i++) {
      T s = w * th[o.or_ + i] - w * sx;
      for (int j = 0; j < m; j++) {
        const T h = w * (th[o.oR + i * m + j] + th[o.oR + j * m + i]);
        R[i * m + j] += h;
        s += h * u[j];
      }
      lu[i] += s;
    }
    dyn(n, m, th, x, u, f);
  }
  template <typename T>
  __device__ static void final_derivs(int n, int m, const T* th, const T* x, T* lfx, T* Qf, int ld) {
    const Th o(n, m);
    for (int i = 0; i < n; i++) {
      T s = T(0);
      for (int j = 0; j < n; j++) {
        const T h = T(0.5) * (th[o.oQf + i * n + j] + th[o.oQf + j * n + i]);
        Qf[i * ld + j] = h;
        s += h * x[j];
      }
      lfx[i] = s;
    }
  }
  // leaves of one system as functions of theta at fixed (x,u): Q_t = sym(Q), q_t = q, R_t += w (R+R'), r_t += w r,
  // M_t const, A_t = A diag(g) (PRE = false) or diag(g) A (PRE = true), B_t = B, d_t = 0.5
  template <typename T>
  __device__ static T theta_vjp_step(int n, int m, int e, const T* th, const T* x, const T* /*u*/, const T* /*nu*/, const LeafCot<T>& g) {
    const T *gQ = g.gQ, *gq = g.gq, *gR = g.gR, *gr = g.gr, *gA = g.gA, *gB = g.gB;
    const int ldq = g.ldq, lda = g.lda;
    const Th o(n, m);
    const T w = wgt<T>();
    if (e < o.oq) { const int i = e / n, j = e % n; return T(0.5) * (gQ[i * ldq + j] + gQ[j * ldq + i]); }
    if (e < o.oQf) return gq[e - o.oq];
    if (e < o.oR) return T(0);
    if (e < o.or_) { const int k = e - o.oR, i = k / m, j = k % m; return w * (gR[i * m + j] + gR[j * m + i]); }
    if (e < o.oA) return w * gr[e - o.or_];
    if (e < o.oB) {
      const int k = e - o.oA, i = k / n, j = k % n;
      const T g = PRE ? gate(n, m, th, x, i) : (x[j] > T(0) ? T(1) : T(0));
      return g * gA[i * lda + j];
    }
    return gB[e - o.oB];
  }
  template <typename T>
  __device__ static T theta_vjp_final(int n, int m, int e, const T* /*th*/, const T* /*xT*/, const T* gQf, int ld, const T* /*gqf*/) {
    const Th o(n, m);
    if (e >= o.oQf && e < o.oR) { const int k = e - o.oQf, i = k / n, j = k % n; return T(0.5) * (gQf[i * ld + j] + gQf[j * ld + i]); }
    return T(0);
  }
};
using ReluRnn = ReluSys<false>;
using ReluAx = ReluSys<true>;

// damped pendulum swing-up (n = 2, m = 1), explicit Euler with step h:
//   th' = th + h w,   w' = w + h (-g_l sin(th) - b w + c u);   l = 1/2 (w1 (th-pi)^2 + w2 w^2 + wu u^2),  lf = 1/2 (f1 (th-pi)^2 + f2 w^2)
// theta = [h, g_l, b, c, w1, w2, wu, f1, f2]
struct Pendulum {
  __host__ __device__ static int theta_dim(int, int) { return 9; }
  template <typename T>
  __device__ static void dyn(int, int, const T* th, const T* x, const T* u, T* f) {
    f[0] = x[0] + th[0] * x[1];
    f[1] = x[1] + th[0] * (-th[1] * sin(x[0]) - th[2] * x[1] + th[3] * u[0]);
  }
  template <typename T>
  __device__ static T cost(int, int, const T* th, const T* x, const T* u) {
    const T e = x[0] - T(3.14159265358979323846);
    return T(0.5) * (th[4] * e * e + th[5] * x[1] * x[1] + th[6] * u[0] * u[0]);
  }
  template <typename T>
  __device__ static T costf(int, int, const T* th, const T* x) {
    const T e = x[0] - T(3.14159265358979323846);
    return T(0.5) * (th[7] * e * e + th[8] * x[1] * x[1]);
  }
  template <typename T>
  __device__ static void derivs(int, int, const T* th, const T* x, const T* u, const T* nu, T* lx, T* lu, T* Q, int ldq, T* R,
                                T* M, T* A, int lda, T* B, T* f) {
    const T e = x[0] - T(3.14159265358979323846);
    lx[0] = th[4] * e; lx[1] = th[5] * x[1]; lu[0] += th[6] * u[0];
    Q[0] = th[4]; Q[1] = T(0); Q[ldq] = T(0); Q[ldq + 1] = th[5];
    R[0] += th[6]; M[0] = T(0); M[1] = T(0);
    A[0] = T(1); A[1] = th[0]; A[lda] = -th[0] * th[1] * cos(x[0]); A[lda + 1] = T(1) - th[0] * th[2];
    B[0] = T(0); B[1] = th[0] * th[3];
    if (nu != nullptr) Q[0] += nu[1] * th[0] * th[1] * sin(x[0]);  // nu_1 * d2 f_1 / d th^2
    dyn(2, 1, th, x, u, f);
  }
  template <typename T>
  __device__ static void final_derivs(int, int, const T* th, const T* x, T* lfx, T* Qf, int ld) {
    const T e = x[0] - T(3.14159265358979323846);
    lfx[0] = th[7] * e; lfx[1] = th[8] * x[1];
    Qf[0] = th[7]; Qf[1] = T(0); Qf[ld] = T(0); Qf[ld + 1] = th[8];
  }
  // Leaves as functions of theta at fixed (x, u):  A = [[1, h], [-h g_l cos x1, 1 - h b]],  B = [0, h c],
  // d = f - A x - B u = [0, h g_l (x1 cos x1 - sin x1)];  the cost enters through l_x = [w1 (x1 - pi), w2 x2], l_u = wu u
  // (the Q x part of q = l_x - Q x - M u cancels against <gQ, dQ>, including the Lagrangian curvature put into Q11).
  template <typename T>
  __device__ static T theta_vjp_step(int, int, int e, const T* th, const T* x, const T* u, const T* /*nu*/, const LeafCot<T>& g) {
    const T s1 = sin(x[0]), c1 = cos(x[0]), lin = x[0] * c1 - s1;
    const T gA12 = g.gA[1], gA21 = g.gA[g.lda], gA22 = g.gA[g.lda + 1], gB2 = g.gB[1], gd2 = g.gd[1];
    switch (e) {
      case 0: return gA12 - th[1] * c1 * gA21 - th[2] * gA22 + th[3] * gB2 + th[1] * lin * gd2;  // h
      case 1: return -th[0] * c1 * gA21 + th[0] * lin * gd2;                                       // g/l
      case 2: return -th[0] * gA22;                                                                // b
      case 3: return th[0] * gB2;                                                                  // c
      case 4: return g.gq[0] * (x[0] - T(3.14159265358979323846));                                 // w1
      case 5: return g.gq[1] * x[1];                                                               // w2
      case 6: return g.gr[0] * u[0];                                                               // wu
      default: return T(0);
    }
  }
  template <typename T>
  __device__ static T theta_vjp_final(int, int, int e, const T*, const T* xT, const T*, int, const T* gqf) {
    if (e == 7) return gqf[0] * (xT[0] - T(3.14159265358979323846));  // f1
    if (e == 8) return gqf[1] * xT[1];                                 // f2
    return T(0);
  }
};

// ---------------------------------------------------------------------------------------------------- kernels
// ---- families written once over the scalar type and differentiated by csrc/ad.cuh (AdFamily<F>)
// the pendulum above, again: validates the AD path against the hand-written derivatives (tests/test_gpu_ilqr.py)
struct PendulumF {
  static constexpr int N = 2, M = 1, TD = 9;
  template <class S>
  __device__ static void dyn(const S* th, const S* x, const S* u, S* f) {
    f[0] = x[0] + th[0] * x[1];
    f[1] = x[1] + th[0] * (th[3] * u[0] - th[1] * msin(x[0]) - th[2] * x[1]);
  }
  template <class S>
  __device__ static S cost(const S* th, const S* x, const S* u) {
    const S e = x[0] - S(3.14159265358979323846);
    return S(0.5) * (th[4] * e * e + th[5] * x[1] * x[1] + th[6] * u[0] * u[0]);
  }
  template <class S>
  __device__ static S costf(const S* th, const S* x) {
    const S e = x[0] - S(3.14159265358979323846);
    return S(0.5) * (th[7] * e * e + th[8] * x[1] * x[1]);
  }
};
using PendulumAd = AdFamily<PendulumF>;

// cart-pole swing-up (BASELINE config 4): state (p, th, pdot, thdot), th measured from the hanging position (pi = upright),
// point mass m_p at distance l on a cart of mass m_c, force u on the cart, explicit Euler with step h:
//   pddot = (u + m_p sin th (l thdot^2 + g cos th)) / (m_c + m_p sin^2 th),   thddot = -(pddot cos th + g sin th) / l
//   l = 1/2 (w_p p^2 + w_th (th - pi)^2 + w_v pdot^2 + w_om thdot^2 + w_u u^2),  lf = 1/2 (f_p p^2 + f_th (th - pi)^2 + f_v pdot^2 + f_om thdot^2)
// theta = [h, m_c, m_p, l, g, w_p, w_th, w_v, w_om, w_u, f_p, f_th, f_v, f_om]
struct CartpoleF {
  static constexpr int N = 4, M = 1, TD = 14;
  template <class S>
  __device__ static void dyn(const S* th, const S* x, const S* u, S* f) {
    const S sn = msin(x[1]), cs = mcos(x[1]);
    const S pdd = (u[0] + th[2] * sn * (th[3] * x[3] * x[3] + th[4] * cs)) / (th[1] + th[2] * sn * sn);
    const S tdd = -(pdd * cs + th[4] * sn) / th[3];
    f[0] = x[0] + th[0] * x[2];
    f[1] = x[1] + th[0] * x[3];
    f[2] = x[2] + th[0] * pdd;
    f[3] = x[3] + th[0] * tdd;
  }
  template <class S>
  __device__ static S cost(const S* th, const S* x, const S* u) {
    const S e = x[1] - S(3.14159265358979323846);
    return S(0.5) * (th[5] * x[0] * x[0] + th[6] * e * e + th[7] * x[2] * x[2] + th[8] * x[3] * x[3] + th[9] * u[0] * u[0]);
  }
  template <class S>
  __device__ static S costf(const S* th, const S* x) {
    const S e = x[1] - S(3.14159265358979323846);
    return S(0.5) * (th[10] * x[0] * x[0] + th[11] * e * e + th[12] * x[2] * x[2] + th[13] * x[3] * x[3]);
  }
};
using Cartpole = AdFamily<CartpoleF>;

template <typename T>
struct IlqrArgs {
  const T* theta;  // [B, theta_dim]
  const T* x0;     // [B, N]           N = bs * n (lifted state)
  T *X, *U;        // current trajectory [B,T,N], [B,T,m]
  T *Xn, *Un;      // trial trajectory
  T *c, *cn;       // [B] cost of current / trial trajectory
  T* alpha;        // [B] line-search step
  int *ls_it, *searching, *active, *iters, *ls_failed, *n_active;
  T *Q, *q, *Qf, *qf, *Mx, *R, *r, *A, *B, *d;  // lifted LQR leaves (written by the linearisation)
  const T *K, *k, *dCdc;                          // gains and expected-change terms from the Riccati sweep
  const T* Nu;                                    // co-states (adjoint-block mode only)
  long long nb;
  int T_, n, m, bs, theta_dim;                    // n = state dimension of ONE system
  T thres, ls_lower, ls_upper, ls_rho;
  int ls_maxiter, use_ls;
};

// open-loop rollout + cost (ilqr.py:161-176; bilqr.py:93-117: cost summed over the systems)
template <typename P, typename T>
__global__ void ilqr_simulate_kernel(const IlqrArgs<T> a, int write_x) {
  const long long b = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= a.nb) return;
  const int n = a.n, m = a.m, bs = a.bs, N = n * bs, Th = a.T_;
  const T* th = a.theta + b * a.theta_dim;
  T x[ILQR_MAXD], xn[ILQR_MAXD], u[ILQR_MAXD];
  for (int i = 0; i < N; i++) x[i] = a.x0[b * N + i];
  T c = T(0);
  for (int t = 0; t < Th; t++) {
    for (int j = 0; j < m; j++) u[j] = a.U[(b * Th + t) * m + j];
    for (int s = 0; s < bs; s++) {
      c += P::cost(n, m, th, x + s * n, u);
      P::dyn(n, m, th, x + s * n, u, xn + s * n);
    }
    for (int i = 0; i < N; i++) {
      x[i] = xn[i];
      if (write_x) a.X[(b * Th + t) * N + i] = xn[i];
    }
  }
  for (int s = 0; s < bs; s++) c += P::costf(n, m, th, x + s * n);
  a.c[b] = c;
}

// LQR approximation around (X, U): mode 0 = local (ilqr.py:143-146), 1 = global (ilqr.py:139-142),
// 2 = blocks of the adjoint system (global + curvature of the dynamics weighted by the co-states)
template <typename P, typename T>
__global__ void ilqr_linearize_kernel(const IlqrArgs<T> a, int mode) {
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const int n = a.n, m = a.m, bs = a.bs, N = n * bs, Th = a.T_;
  if (idx >= a.nb * Th) return;
  const long long b = idx / Th;
  const int t = (int)(idx % Th);
  const T* th = a.theta + b * a.theta_dim;
  T x[ILQR_MAXD], u[ILQR_MAXD], lx[ILQR_MAXD], lu[ILQR_MAXD], f[ILQR_MAXD], nu[ILQR_MAXD];
  T* Q = a.Q + idx * N * N;
  T* A = a.A + idx * N * N;
  T* M = a.Mx + idx * N * m;
  T* Bm = a.B + idx * N * m;
  T* R = a.R + idx * m * m;
  for (int i = 0; i < N; i++) x[i] = t == 0 ? a.x0[b * N + i] : a.X[(idx - 1) * N + i];  // sX = [x0, X[:-1]]  ilqr.py:149
  for (int j = 0; j < m; j++) { u[j] = a.U[idx * m + j]; lu[j] = T(0); }
  if (mode == 2)
    for (int i = 0; i < N; i++) nu[i] = a.Nu[idx * N + i];
  if (bs > 1)
    for (int e = 0; e < N * N; e++) { Q[e] = T(0); A[e] = T(0); }
  for (int e = 0; e < m * m; e++) R[e] = T(0);
  for (int s = 0; s < bs; s++)
    P::derivs(n, m, th, x + s * n, u, mode == 2 ? nu + s * n : (const T*)nullptr, lx + s * n, lu, Q + (s * n) * N + s * n, N, R,
              M + s * n * m, A + (s * n) * N + s * n, N, Bm + s * n * m, f + s * n);
  for (int i = 0; i < N; i++) {
    T qv = lx[i], dv = T(0);
    if (mode != 0) {
      dv = f[i];
      for (int j = 0; j < N; j++) { qv -= Q[i * N + j] * x[j]; dv -= A[i * N + j] * x[j]; }
      for (int j = 0; j < m; j++) { qv -= M[i * m + j] * u[j]; dv -= Bm[i * m + j] * u[j]; }
    }
    a.q[idx * N + i] = qv;
    a.d[idx * N + i] = dv;
  }
  for (int i = 0; i < m; i++) {
    T rv = lu[i];
    if (mode != 0) {
      for (int j = 0; j < m; j++) rv -= R[i * m + j] * u[j];
      for (int j = 0; j < N; j++) rv -= M[j * m + i] * x[j];
    }
    a.r[idx * m + i] = rv;
  }
  if (t == Th - 1) {  // terminal cost at xf = X[-1]                                            ilqr.py:150-155
    T xf[ILQR_MAXD];
    T* Qf = a.Qf + b * N * N;
    for (int i = 0; i < N; i++) xf[i] = a.X[idx * N + i];
    if (bs > 1)
      for (int e = 0; e < N * N; e++) Qf[e] = T(0);
    for (int s = 0; s < bs; s++) P::final_derivs(n, m, th, xf + s * n, lx + s * n, Qf + (s * n) * N + s * n, N);
    for (int i = 0; i < N; i++) {
      T v = lx[i];
      if (mode != 0)
        for (int j = 0; j < N; j++) v -= Qf[i * N + j] * xf[j];
      a.qf[b * N + i] = v;
    }
  }
}

// closed-loop rollout around the previous trajectory with step size alpha (ilqr.py:196-219; bilqr.py:136-167)
template <typename P, typename T>
__global__ void ilqr_update_kernel(const IlqrArgs<T> a) {
  const long long b = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= a.nb) return;
  if (a.searching[b] != 1) return;
  const int n = a.n, m = a.m, bs = a.bs, N = n * bs, Th = a.T_;
  const T* th = a.theta + b * a.theta_dim;
  const T alpha = a.alpha[b];
  T xh[ILQR_MAXD], xn[ILQR_MAXD], uh[ILQR_MAXD];
  for (int i = 0; i < N; i++) xh[i] = a.x0[b * N + i];
  T l = T(0);
  for (int t = 0; t < Th; t++) {
    const long long idx = b * Th + t;
    const T* xo = t == 0 ? a.x0 + b * N : a.X + (idx - 1) * N;
    for (int j = 0; j < m; j++) {
      T du = alpha * a.k[idx * m + j];
      for (int i = 0; i < N; i++) du += a.K[(idx * m + j) * N + i] * (xh[i] - xo[i]);
      uh[j] = a.U[idx * m + j] + du;
      a.Un[idx * m + j] = uh[j];
    }
    for (int s = 0; s < bs; s++) {
      l += P::cost(n, m, th, xh + s * n, uh);
      P::dyn(n, m, th, xh + s * n, uh, xn + s * n);
    }
    for (int i = 0; i < N; i++) {
      xh[i] = xn[i];
      a.Xn[idx * N + i] = xn[i];
    }
  }
  for (int s = 0; s < bs; s++) l += P::costf(n, m, th, xh + s * n);
  a.cn[b] = l;
}

// one backtracking decision per problem (line_search.py:12-27) + iLQR convergence bookkeeping (ilqr.py:246-252)
template <typename T>
__global__ void ilqr_ls_kernel(const IlqrArgs<T> a) {
  const long long b = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= a.nb) return;
  if (a.searching[b] != 1) return;
  const T alpha = a.alpha[b];
  const T c0 = a.c[b], c1 = a.cn[b];
  bool carry_on = false;
  if (a.use_ls) {
    const T ec = alpha * alpha * a.dCdc[b * 2] + alpha * a.dCdc[b * 2 + 1];
    const T p = abs_((c0 - c1 + T(1e-10)) / (T(1e-10) + ec));
    carry_on = !(p > a.ls_lower && p < a.ls_upper);  // = (p <= lower) || (p >= upper), and a NaN ratio (overflowed trial) backtracks too
  }
  const int it = a.ls_it[b] + 1;
  if (carry_on && it < a.ls_maxiter) {
    a.alpha[b] = alpha * a.ls_rho;
    a.ls_it[b] = it;
    return;
  }
  // search over for this problem: the last tried trajectory is taken even if it failed (line_search.py:29-40)
  a.searching[b] = 2;  // 2 = commit pending
  if (carry_on) a.ls_failed[b] += 1;
}

// copy the accepted trial trajectory, update cost / convergence flags
template <typename T>
__global__ void ilqr_commit_kernel(const IlqrArgs<T> a) {
  const long long b = blockIdx.x;
  if (a.searching[b] != 2) return;
  const int n = a.n * a.bs, m = a.m, Th = a.T_;
  for (int e = threadIdx.x; e < Th * n; e += blockDim.x) a.X[b * Th * n + e] = a.Xn[b * Th * n + e];
  for (int e = threadIdx.x; e < Th * m; e += blockDim.x) a.U[b * Th * m + e] = a.Un[b * Th * m + e];
  __syncthreads();
  if (threadIdx.x == 0) {
    const T c0 = a.c[b], c1 = a.cn[b];
    const T pct = abs_((c0 - c1) / c0);  // ilqr.py:246
    a.c[b] = c1;
    a.searching[b] = 0;
    a.iters[b] += 1;
    if (!(pct > a.thres)) {
      a.active[b] = 0;
    } else {
      atomicAdd(a.n_active, 1);
    }
  }
}

// start of an iteration: every still-active problem searches from alpha = 1
template <typename T>
__global__ void ilqr_begin_iter_kernel(const IlqrArgs<T> a) {
  const long long b = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (b == 0) *a.n_active = 0;
  if (b >= a.nb) return;
  a.searching[b] = a.active[b] ? 1 : 0;
  a.alpha[b] = T(1);
  a.ls_it[b] = 0;
}

// theta cotangent: sum over t (and over the systems sharing theta) of the leaf cotangents pulled back through the
// global approximation
template <typename P, typename T>
__global__ void ilqr_theta_vjp_kernel(const IlqrArgs<T> a, const T* gQ, const T* gq, const T* gQf, const T* gqf, const T* gM,
                                      const T* gR, const T* gr, const T* gA, const T* gB, const T* gd, T* gtheta) {
  const long long b = blockIdx.x;
  const int n = a.n, m = a.m, bs = a.bs, N = n * bs, Th = a.T_;
  const T* th = a.theta + b * a.theta_dim;
  for (int e = threadIdx.x; e < a.theta_dim; e += blockDim.x) {
    T acc = T(0);
    for (int s = 0; s < bs; s++) {
      acc += P::theta_vjp_final(n, m, e, th, a.X + (b * Th + Th - 1) * N + s * n, gQf + b * N * N + (s * n) * N + s * n, N,
                                gqf + b * N + s * n);
      for (int t = 0; t < Th; t++) {
        const long long idx = b * Th + t;
        const T* x = (t == 0 ? a.x0 + b * N : a.X + (idx - 1) * N) + s * n;
        LeafCot<T> g;
        g.gQ = gQ + idx * N * N + (s * n) * N + s * n; g.ldq = N; g.gq = gq + idx * N + s * n;
        g.gR = gR + idx * m * m; g.gr = gr + idx * m; g.gM = gM + idx * N * m + s * n * m;
        g.gA = gA + idx * N * N + (s * n) * N + s * n; g.lda = N; g.gB = gB + idx * N * m + s * n * m; g.gd = gd + idx * N + s * n;
        acc += P::theta_vjp_step(n, m, e, th, x, a.U + idx * m, a.Nu + idx * N + s * n, g);
      }
    }
    gtheta[b * a.theta_dim + e] = acc;
  }
}

}  // namespace idoc
