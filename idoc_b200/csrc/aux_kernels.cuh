// Shape-generic auxiliary kernels: KKT residual (idoc/lqr.py:128-150) and the synthetic problem
// generator of SURVEY.md 8(d).  One thread per (problem, timestep); runtime n, m (n <= 32, m <= 32).
#pragma once
#include "lqr_geo.cuh"

namespace idoc {

template <typename T>
struct KktArgs {
  const T *Q, *q, *Qf, *qf, *Mx, *R, *r, *A, *B, *d, *x0, *X, *U, *Nu;
  T *rX, *rU, *rNu;
  long long nb;
  int T_, n, m;
};

template <typename T>
__global__ void lqr_kkt_kernel(const KktArgs<T> a) {
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const int Th = a.T_, n = a.n, m = a.m;
  if (idx >= a.nb * Th) return;
  const long long b = idx / Th;
  const int t = (int)(idx % Th);
  const T* X = a.X + (b * Th) * n;
  const T* U = a.U + (b * Th) * m;
  const T* Nu = a.Nu + (b * Th) * n;
  const T* sx = t == 0 ? a.x0 + b * n : X + (long long)(t - 1) * n;  // sX = [x0, X[:-1]]   lqr.py:147
  const T* At = a.A + idx * n * n;
  const T* Bt = a.B + idx * n * m;
  const T* Mt = a.Mx + idx * n * m;
  const T* Rt = a.R + idx * m * m;
  // dLdNu[t] = d + A sX + B U - X                                       lqr.py:149
  for (int i = 0; i < n; i++) {
    T s = a.d[idx * n + i] - X[(long long)t * n + i];
    for (int j = 0; j < n; j++) s += At[i * n + j] * sx[j];
    for (int j = 0; j < m; j++) s += Bt[i * m + j] * U[(long long)t * m + j];
    a.rNu[idx * n + i] = s;
  }
  // dLdU[t] = B^T Nu + sym(R) U + r + M^T sX                            lqr.py:148
  for (int c = 0; c < m; c++) {
    T s = a.r[idx * m + c];
    for (int i = 0; i < n; i++) s += Bt[i * m + c] * Nu[(long long)t * n + i] + Mt[i * m + c] * sx[i];
    for (int j = 0; j < m; j++) s += T(0.5) * (Rt[c * m + j] + Rt[j * m + c]) * U[(long long)t * m + j];
    a.rU[idx * m + c] = s;
  }
  // dLdX[t]                                                             lqr.py:140-146
  if (t < Th - 1) {
    const T* A1 = At + n * n;
    const T* M1 = Mt + n * m;
    const T* Q1 = a.Q + (idx + 1) * n * n;
    for (int i = 0; i < n; i++) {
      T s = a.q[(idx + 1) * n + i] - Nu[(long long)t * n + i];
      for (int j = 0; j < m; j++) s += M1[i * m + j] * U[(long long)(t + 1) * m + j];
      for (int j = 0; j < n; j++)
        s += T(0.5) * (Q1[i * n + j] + Q1[j * n + i]) * X[(long long)t * n + j] + A1[j * n + i] * Nu[(long long)(t + 1) * n + j];
      a.rX[idx * n + i] = s;
    }
  } else {
    const T* Qf = a.Qf + b * n * n;
    for (int i = 0; i < n; i++) {
      T s = a.qf[b * n + i] - Nu[(long long)t * n + i];
      for (int j = 0; j < n; j++) s += T(0.5) * (Qf[i * n + j] + Qf[j * n + i]) * X[(long long)t * n + j];
      a.rX[idx * n + i] = s;
    }
  }
}

// ------------------------------------------------------------------ co-state fix-up + non-finite report (rare paths)
// The rollout kernel forms Nu[t] = V_{t+1} x_{t+1} + v_{t+1}, which equals the reference's adjoint recursion
// (idoc/lqr.py:110-125) whenever no eigen-shift was taken.  For the (rare) problems whose status has ST_SHIFT
// set the trajectory is not an exact KKT point, so their co-states are recomputed with the reference's own
// recursion:  Nu[T-1] = Qf X[T-1] + qf;  Nu[t-1] = A[t]^T Nu[t] + Q[t] X[t-1] + q[t] + M[t] U[t].
//
// Non-finite solves.  The reference propagates NaN / Inf from its inputs (and from a failed factorisation) into X, U, Nu.
// The kernels' pivot clamps could instead turn such a problem into finite garbage, so: a NaN / Inf anywhere in the problem
// reaches the final state x_T through the recursions (v, V -> K, k -> u -> x), the Riccati sweeps flag NaN pivots and the
// tilqr factorisation failure themselves; every problem with either sign gets IDOC_STATUS_NONFINITE and NaN outputs.
template <typename T>
struct FixupArgs {
  const T *Q, *q, *Qf, *qf, *Mx, *A, *U;
  T *X, *Unf, *Nu;     // Unf: U again, writable (NaN fill)
  int* status;         // [B] workspace status word (read + updated); null = recompute the co-states of every problem
  int* ustatus;        // optional caller-visible copy
  long long nb;
  int T_, n, m;
};

template <typename T>
__global__ void lqr_costate_fixup_kernel(const FixupArgs<T> a) {
  const long long b = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= a.nb) return;
  int st = a.status != nullptr ? a.status[b] : ST_SHIFT;  // no status array (iLQR's final co-states): recompute every problem
  if (a.status != nullptr) {
    const T* xT = a.X + (b * a.T_ + a.T_ - 1) * a.n;
    bool bad = (st & ST_NONFINITE) != 0;
    for (int i = 0; i < a.n; i++) bad = bad || !(abs_(xT[i]) <= T(sizeof(T) == 8 ? 1.7e308 : 3.4e38));
    if (bad) {
      const T nan = (T)__int_as_float(0x7fc00000);
      for (long long e = 0; e < (long long)a.T_ * a.n; e++) { a.X[b * a.T_ * a.n + e] = nan; a.Nu[b * a.T_ * a.n + e] = nan; }
      for (long long e = 0; e < (long long)a.T_ * a.m; e++) a.Unf[b * a.T_ * a.m + e] = nan;
      st |= ST_NONFINITE;
      a.status[b] = st;
      if (a.ustatus != nullptr) a.ustatus[b] = st;
      return;
    }
  }
  if ((st & ST_SHIFT) == 0 || a.Q == nullptr) return;
  const int Th = a.T_, n = a.n, m = a.m;
  const T* X = a.X + b * Th * n;
  const T* U = a.U + b * Th * m;
  T* Nu = a.Nu + b * Th * n;
  const T* Qf = a.Qf + b * n * n;
  for (int i = 0; i < n; i++) {
    T s = a.qf[b * n + i];
    for (int j = 0; j < n; j++) s += T(0.5) * (Qf[i * n + j] + Qf[j * n + i]) * X[(long long)(Th - 1) * n + j];
    Nu[(long long)(Th - 1) * n + i] = s;
  }
  for (int t = Th - 1; t >= 1; t--) {
    const long long idx = b * Th + t;
    const T* At = a.A + idx * n * n;
    const T* Qt = a.Q + idx * n * n;
    const T* Mt = a.Mx + idx * n * m;
    for (int i = 0; i < n; i++) {
      T s = a.q[idx * n + i];
      for (int j = 0; j < n; j++)
        s += At[j * n + i] * Nu[(long long)t * n + j] + T(0.5) * (Qt[i * n + j] + Qt[j * n + i]) * X[(long long)(t - 1) * n + j];
      for (int j = 0; j < m; j++) s += Mt[i * m + j] * U[(long long)t * m + j];
      Nu[(long long)(t - 1) * n + i] = s;
    }
  }
}

// ------------------------------------------------------------------ generator
__device__ __forceinline__ uint64_t mix64(uint64_t z) {
  z += 0x9E3779B97F4A7C15ull;
  z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
  z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
  return z ^ (z >> 31);
}
struct Rng {
  uint64_t key, ctr;
  __device__ double uniform() {  // (0,1)
    const uint64_t r = mix64(key ^ mix64(ctr++));
    return ((double)(r >> 11) + 0.5) * (1.0 / 9007199254740992.0);
  }
  __device__ double normal() {
    const double u1 = uniform(), u2 = uniform();
    return sqrt(-2.0 * log(u1)) * cospi(2.0 * u2);
  }
};

template <typename T>
struct GenArgs {
  T *Q, *q, *Qf, *qf, *Mx, *R, *r, *A, *B, *d, *x0;
  long long nb;
  int T_, n, m;
  uint64_t seed;
};

constexpr int GEN_MAX = 32;

// W W^T / k + I  (symmetric positive definite), written row-major into dst
template <typename T>
__device__ void gen_spd(Rng& g, int k, T* dst, double* W) {
  for (int i = 0; i < k * k; i++) W[i] = g.normal();
  for (int i = 0; i < k; i++)
    for (int j = 0; j <= i; j++) {
      double s = 0.0;
      for (int l = 0; l < k; l++) s += W[i * k + l] * W[j * k + l];
      s = s / k + (i == j ? 1.0 : 0.0);
      dst[i * k + j] = (T)s;
      dst[j * k + i] = (T)s;
    }
}

template <typename T>
__global__ void lqr_generate_kernel(const GenArgs<T> a) {
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const int Th = a.T_, n = a.n, m = a.m;
  if (idx >= a.nb * (Th + 1)) return;
  double W[GEN_MAX * GEN_MAX];
  Rng g{mix64(a.seed), (uint64_t)idx << 20};
  if (idx >= a.nb * Th) {  // per-problem leaves
    const long long b = idx - a.nb * Th;
    gen_spd<T>(g, n, a.Qf + b * n * n, W);
    for (int i = 0; i < n; i++) a.qf[b * n + i] = (T)g.normal();
    for (int i = 0; i < n; i++) a.x0[b * n + i] = (T)g.normal();
    return;
  }
  // A = 0.5 * orthonormal factor of a Gaussian matrix (idoc/utils.py:9-13), by modified Gram-Schmidt
  for (int i = 0; i < n * n; i++) W[i] = g.normal();
  for (int c = 0; c < n; c++) {
    for (int p = 0; p < c; p++) {
      double dot = 0.0;
      for (int i = 0; i < n; i++) dot += W[i * n + p] * W[i * n + c];
      for (int i = 0; i < n; i++) W[i * n + c] -= dot * W[i * n + p];
    }
    double nrm = 0.0;
    for (int i = 0; i < n; i++) nrm += W[i * n + c] * W[i * n + c];
    nrm = 1.0 / sqrt(nrm);
    for (int i = 0; i < n; i++) W[i * n + c] *= nrm;
  }
  for (int i = 0; i < n * n; i++) a.A[idx * n * n + i] = (T)(0.5 * W[i]);
  for (int i = 0; i < n * m; i++) a.B[idx * n * m + i] = (T)g.normal();
  for (int i = 0; i < n * m; i++) a.Mx[idx * n * m + i] = (T)(0.1 * g.normal());
  for (int i = 0; i < n; i++) a.d[idx * n + i] = (T)g.normal();
  for (int i = 0; i < n; i++) a.q[idx * n + i] = (T)g.normal();
  for (int i = 0; i < m; i++) a.r[idx * m + i] = (T)g.normal();
  gen_spd<T>(g, n, a.Q + idx * n * n, W);
  gen_spd<T>(g, m, a.R + idx * m * m, W);
}

}  // namespace idoc
