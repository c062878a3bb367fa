// Forward-mode automatic differentiation on the device, for compiled iLQR problem families.
//
// The reference builds its LQR approximation with jax autodiff (idoc/ilqr.py:127-158: grad / jacfwd / jacobian of the
// user's cost and dynamics).  A CUDA kernel cannot trace Python, but a problem family written ONCE as templates over the
// scalar type (dyn, cost, costf) can be differentiated the same way: `Jet2<T, Z>` carries value, gradient and Hessian
// w.r.t. Z inputs (one evaluation gives l_x, l_u, l_xx, l_xu, l_uu, or f, f_x, f_u and the curvature of every f_i);
// `HDual<T>` is a hyper-dual number (value, d/de, d/dd, d2/de dd) used for the theta pull-back of the implicit VJP, where
// one mixed second derivative per theta element is needed.  `AdFamily<F>` adapts such a family to the functor
// interface of csrc/ilqr.cuh (derivs, final_derivs, theta_vjp_step, theta_vjp_final), so the hand-written derivative
// code of a family becomes optional; tests compare the two on the pendulum.
#pragma once
#include "common.cuh"

namespace idoc {

// ------------------------------------------------------------------------------------------------ second-order jet
template <typename T, int Z>
struct Jet2 {
  static constexpr int H = Z * (Z + 1) / 2;
  T v, g[Z], h[H];  // h packed upper: (i, j), i <= j at i*Z - i*(i-1)/2 + (j - i)
  __device__ Jet2() {}
  __device__ Jet2(T c) : v(c) {
#pragma unroll
    for (int i = 0; i < Z; i++) g[i] = T(0);
#pragma unroll
    for (int i = 0; i < H; i++) h[i] = T(0);
  }
  __device__ static Jet2 var(T c, int k) {
    Jet2 r(c);
#pragma unroll
    for (int i = 0; i < Z; i++)
      if (i == k) r.g[i] = T(1);
    return r;
  }
  __device__ static constexpr int hi(int i, int j) { return i * Z - i * (i - 1) / 2 + (j - i); }
  __device__ T hess(int i, int j) const { return i <= j ? h[hi(i, j)] : h[hi(j, i)]; }
};
// r = phi(a) with phi' = d1, phi'' = d2 at a.v
template <typename T, int Z>
__device__ Jet2<T, Z> jet_chain(const Jet2<T, Z>& a, T val, T d1, T d2) {
  Jet2<T, Z> r;
  r.v = val;
#pragma unroll
  for (int i = 0; i < Z; i++) r.g[i] = d1 * a.g[i];
#pragma unroll
  for (int i = 0; i < Z; i++)
#pragma unroll
    for (int j = i; j < Z; j++) r.h[Jet2<T, Z>::hi(i, j)] = d1 * a.h[Jet2<T, Z>::hi(i, j)] + d2 * a.g[i] * a.g[j];
  return r;
}
template <typename T, int Z>
__device__ Jet2<T, Z> operator+(const Jet2<T, Z>& a, const Jet2<T, Z>& b) {
  Jet2<T, Z> r;
  r.v = a.v + b.v;
#pragma unroll
  for (int i = 0; i < Z; i++) r.g[i] = a.g[i] + b.g[i];
#pragma unroll
  for (int i = 0; i < Jet2<T, Z>::H; i++) r.h[i] = a.h[i] + b.h[i];
  return r;
}
template <typename T, int Z>
__device__ Jet2<T, Z> operator-(const Jet2<T, Z>& a) {
  Jet2<T, Z> r;
  r.v = -a.v;
#pragma unroll
  for (int i = 0; i < Z; i++) r.g[i] = -a.g[i];
#pragma unroll
  for (int i = 0; i < Jet2<T, Z>::H; i++) r.h[i] = -a.h[i];
  return r;
}
template <typename T, int Z>
__device__ Jet2<T, Z> operator-(const Jet2<T, Z>& a, const Jet2<T, Z>& b) { return a + (-b); }
template <typename T, int Z>
__device__ Jet2<T, Z> operator*(const Jet2<T, Z>& a, const Jet2<T, Z>& b) {
  Jet2<T, Z> r;
  r.v = a.v * b.v;
#pragma unroll
  for (int i = 0; i < Z; i++) r.g[i] = a.g[i] * b.v + a.v * b.g[i];
#pragma unroll
  for (int i = 0; i < Z; i++)
#pragma unroll
    for (int j = i; j < Z; j++) {
      const int k = Jet2<T, Z>::hi(i, j);
      r.h[k] = a.h[k] * b.v + a.v * b.h[k] + a.g[i] * b.g[j] + a.g[j] * b.g[i];
    }
  return r;
}
template <typename T, int Z>
__device__ Jet2<T, Z> inv(const Jet2<T, Z>& a) {
  const T r = T(1) / a.v;
  return jet_chain(a, r, -r * r, T(2) * r * r * r);
}
template <typename T, int Z>
__device__ Jet2<T, Z> operator/(const Jet2<T, Z>& a, const Jet2<T, Z>& b) { return a * inv(b); }
template <typename T, int Z>
__device__ Jet2<T, Z> msin(const Jet2<T, Z>& a) { const T s = ::sin(a.v), c = ::cos(a.v); return jet_chain(a, s, c, -s); }
template <typename T, int Z>
__device__ Jet2<T, Z> mcos(const Jet2<T, Z>& a) { const T s = ::sin(a.v), c = ::cos(a.v); return jet_chain(a, c, -s, -c); }
// mixed scalar forms
template <typename T, int Z> __device__ Jet2<T, Z> operator+(const Jet2<T, Z>& a, T b) { Jet2<T, Z> r = a; r.v += b; return r; }
template <typename T, int Z> __device__ Jet2<T, Z> operator+(T b, const Jet2<T, Z>& a) { return a + b; }
template <typename T, int Z> __device__ Jet2<T, Z> operator-(const Jet2<T, Z>& a, T b) { return a + (-b); }
template <typename T, int Z> __device__ Jet2<T, Z> operator-(T b, const Jet2<T, Z>& a) { return (-a) + b; }
template <typename T, int Z> __device__ Jet2<T, Z> operator*(const Jet2<T, Z>& a, T b) {
  Jet2<T, Z> r;
  r.v = a.v * b;
#pragma unroll
  for (int i = 0; i < Z; i++) r.g[i] = a.g[i] * b;
#pragma unroll
  for (int i = 0; i < Jet2<T, Z>::H; i++) r.h[i] = a.h[i] * b;
  return r;
}
template <typename T, int Z> __device__ Jet2<T, Z> operator*(T b, const Jet2<T, Z>& a) { return a * b; }
template <typename T, int Z> __device__ Jet2<T, Z> operator/(const Jet2<T, Z>& a, T b) { return a * (T(1) / b); }
template <typename T, int Z> __device__ Jet2<T, Z> operator/(T b, const Jet2<T, Z>& a) { return inv(a) * b; }

// ------------------------------------------------------------------------------------------------ hyper-dual
template <typename T>
struct HDual {
  T v, e, d, ed;
  __device__ HDual() {}
  __device__ HDual(T c) : v(c), e(T(0)), d(T(0)), ed(T(0)) {}
  __device__ HDual(T c, T e_, T d_) : v(c), e(e_), d(d_), ed(T(0)) {}
};
template <typename T> __device__ HDual<T> hd_chain(const HDual<T>& a, T val, T d1, T d2) {
  HDual<T> r;
  r.v = val; r.e = d1 * a.e; r.d = d1 * a.d; r.ed = d1 * a.ed + d2 * a.e * a.d;
  return r;
}
template <typename T> __device__ HDual<T> operator+(const HDual<T>& a, const HDual<T>& b) {
  HDual<T> r; r.v = a.v + b.v; r.e = a.e + b.e; r.d = a.d + b.d; r.ed = a.ed + b.ed; return r;
}
template <typename T> __device__ HDual<T> operator-(const HDual<T>& a) { HDual<T> r; r.v = -a.v; r.e = -a.e; r.d = -a.d; r.ed = -a.ed; return r; }
template <typename T> __device__ HDual<T> operator-(const HDual<T>& a, const HDual<T>& b) { return a + (-b); }
template <typename T> __device__ HDual<T> operator*(const HDual<T>& a, const HDual<T>& b) {
  HDual<T> r;
  r.v = a.v * b.v; r.e = a.e * b.v + a.v * b.e; r.d = a.d * b.v + a.v * b.d;
  r.ed = a.ed * b.v + a.e * b.d + a.d * b.e + a.v * b.ed;
  return r;
}
template <typename T> __device__ HDual<T> inv(const HDual<T>& a) { const T r = T(1) / a.v; return hd_chain(a, r, -r * r, T(2) * r * r * r); }
template <typename T> __device__ HDual<T> operator/(const HDual<T>& a, const HDual<T>& b) { return a * inv(b); }
template <typename T> __device__ HDual<T> msin(const HDual<T>& a) { const T s = ::sin(a.v), c = ::cos(a.v); return hd_chain(a, s, c, -s); }
template <typename T> __device__ HDual<T> mcos(const HDual<T>& a) { const T s = ::sin(a.v), c = ::cos(a.v); return hd_chain(a, c, -s, -c); }
template <typename T> __device__ HDual<T> operator+(const HDual<T>& a, T b) { HDual<T> r = a; r.v += b; return r; }
template <typename T> __device__ HDual<T> operator+(T b, const HDual<T>& a) { return a + b; }
template <typename T> __device__ HDual<T> operator-(const HDual<T>& a, T b) { return a + (-b); }
template <typename T> __device__ HDual<T> operator-(T b, const HDual<T>& a) { return (-a) + b; }
template <typename T> __device__ HDual<T> operator*(const HDual<T>& a, T b) { HDual<T> r; r.v = a.v * b; r.e = a.e * b; r.d = a.d * b; r.ed = a.ed * b; return r; }
template <typename T> __device__ HDual<T> operator*(T b, const HDual<T>& a) { return a * b; }
template <typename T> __device__ HDual<T> operator/(const HDual<T>& a, T b) { return a * (T(1) / b); }
template <typename T> __device__ HDual<T> operator/(T b, const HDual<T>& a) { return inv(a) * b; }

// ------------------------------------------------------------------------------------------------ family adapter
// F provides, as templates over the scalar type S (T, Jet2<T, Z> or HDual<T>), with compile-time dimensions:
//   static constexpr int N, M, TD;
//   template <class S> static void dyn(const S* th, const S* x, const S* u, S* f);
//   template <class S> static S cost(const S* th, const S* x, const S* u);
//   template <class S> static S costf(const S* th, const S* x);
template <typename F>
struct AdFamily {
  static constexpr int N = F::N, M = F::M, TD = F::TD, Z = N + M;
  __host__ __device__ static int theta_dim(int, int) { return TD; }
  template <typename T>
  __device__ static void dyn(int, int, const T* th, const T* x, const T* u, T* f) { F::template dyn<T>(th, x, u, f); }
  template <typename T>
  __device__ static T cost(int, int, const T* th, const T* x, const T* u) { return F::template cost<T>(th, x, u); }
  template <typename T>
  __device__ static T costf(int, int, const T* th, const T* x) { return F::template costf<T>(th, x); }

  template <typename T>
  __device__ static void derivs(int, int, const T* th, const T* x, const T* u, const T* nu, T* lx, T* lu, T* Q, int ldq, T* R,
                                T* Mx, T* A, int lda, T* B, T* f) {
    using J = Jet2<T, Z>;
    J thj[TD], xj[N], uj[M], fj[N];
#pragma unroll
    for (int e = 0; e < TD; e++) thj[e] = J(th[e]);
#pragma unroll
    for (int i = 0; i < N; i++) xj[i] = J::var(x[i], i);
#pragma unroll
    for (int j = 0; j < M; j++) uj[j] = J::var(u[j], N + j);
    const J l = F::template cost<J>(thj, xj, uj);
    F::template dyn<J>(thj, xj, uj, fj);
#pragma unroll
    for (int i = 0; i < N; i++) {
      lx[i] = l.g[i];
      f[i] = fj[i].v;
#pragma unroll
      for (int j = 0; j < N; j++) A[i * lda + j] = fj[i].g[j];
#pragma unroll
      for (int j = 0; j < M; j++) B[i * M + j] = fj[i].g[N + j];
    }
#pragma unroll
    for (int j = 0; j < M; j++) lu[j] += l.g[N + j];
    // Hessian of the cost, plus the curvature of the dynamics weighted by the co-states (adjoint blocks only)
#pragma unroll
    for (int i = 0; i < Z; i++)
#pragma unroll
      for (int j = 0; j < Z; j++) {
        T hv = l.hess(i, j);
        if (nu != nullptr) {
#pragma unroll
          for (int k = 0; k < N; k++) hv += nu[k] * fj[k].hess(i, j);
        }
        if (i < N && j < N) Q[i * ldq + j] = hv;
        else if (i < N && j >= N) Mx[i * M + (j - N)] = hv;
        else if (i >= N && j >= N) R[(i - N) * M + (j - N)] += hv;
      }
  }
  template <typename T>
  __device__ static void final_derivs(int, int, const T* th, const T* x, T* lfx, T* Qf, int ld) {
    using J = Jet2<T, N>;
    J thj[TD], xj[N];
#pragma unroll
    for (int e = 0; e < TD; e++) thj[e] = J(th[e]);
#pragma unroll
    for (int i = 0; i < N; i++) xj[i] = J::var(x[i], i);
    const J l = F::template costf<J>(thj, xj);
#pragma unroll
    for (int i = 0; i < N; i++) {
      lfx[i] = l.g[i];
#pragma unroll
      for (int j = 0; j < N; j++) Qf[i * ld + j] = l.hess(i, j);
    }
  }
  // d/d theta_e of  S(theta) = dz . grad_z l + nu . (J_f dz) + unu . f   with dz = (ux, uu) = (gq, gr), unu = gd:
  // one hyper-dual evaluation (e along dz, d along theta_e) of cost and dynamics.
  template <typename T>
  __device__ static T theta_vjp_step(int, int, int e, const T* th, const T* x, const T* u, const T* nu, const LeafCot<T>& g) {
    using D = HDual<T>;
    D thd[TD], xd[N], ud[M], fd[N];
#pragma unroll
    for (int k = 0; k < TD; k++) thd[k] = D(th[k], T(0), k == e ? T(1) : T(0));
#pragma unroll
    for (int i = 0; i < N; i++) xd[i] = D(x[i], g.gq[i], T(0));
#pragma unroll
    for (int j = 0; j < M; j++) ud[j] = D(u[j], g.gr[j], T(0));
    const D l = F::template cost<D>(thd, xd, ud);
    F::template dyn<D>(thd, xd, ud, fd);
    T acc = l.ed;
#pragma unroll
    for (int i = 0; i < N; i++) acc += nu[i] * fd[i].ed + g.gd[i] * fd[i].d;
    return acc;
  }
  template <typename T>
  __device__ static T theta_vjp_final(int, int, int e, const T* th, const T* xT, const T*, int, const T* gqf) {
    using D = HDual<T>;
    D thd[TD], xd[N];
#pragma unroll
    for (int k = 0; k < TD; k++) thd[k] = D(th[k], T(0), k == e ? T(1) : T(0));
#pragma unroll
    for (int i = 0; i < N; i++) xd[i] = D(xT[i], gqf[i], T(0));
    return F::template costf<D>(thd, xd).ed;
  }
};

}  // namespace idoc
