// Register-resident small dense linear algebra used by the Riccati sweep:
// Cholesky + inverse factor of the m x m block Guu, and the (rare) eigen-shift of idoc/lqr.py:86-88.
#pragma once
#include "common.cuh"

namespace idoc {

__host__ __device__ constexpr int tril(int a, int b) { return a * (a + 1) / 2 + b; }  // b <= a

// Gs: symmetric M x M, packed upper row-major (triu).  On success Li = inverse of the lower Cholesky
// factor (packed lower row-major, tril): Gs^{-1} = Li^T Li.  Returns false on a non-positive pivot.
// If CLAMP, non-positive pivots are replaced by `tiny` (and *clamped is set) instead of failing.
template <typename T, int M, bool CLAMP>
IDOC_DEVICE bool chol_inv(const T* Gs, T* Li, T tiny, bool* clamped) {
  T Lf[tri(M)];
  T dinv[M];
  bool ok = true;
#pragma unroll
  for (int a = 0; a < M; a++) {
#pragma unroll
    for (int b = 0; b <= a; b++) {
      T s = Gs[triu(b, a, M)];
#pragma unroll
      for (int c = 0; c < b; c++) s -= Lf[tril(a, c)] * Lf[tril(b, c)];
      if (a == b) {
        if (!(s > T(0))) {
          if constexpr (CLAMP) {
            s = tiny;
            *clamped = true;
          } else {
            ok = false;
            s = T(1);
          }
        }
        dinv[a] = rsqrt_<T>(s);
        Lf[tril(a, a)] = s * dinv[a];
      } else {
        Lf[tril(a, b)] = s * dinv[b];
      }
    }
  }
#pragma unroll
  for (int b = 0; b < M; b++) {
    Li[tril(b, b)] = dinv[b];
#pragma unroll
    for (int a = b + 1; a < M; a++) {
      T s = T(0);
#pragma unroll
      for (int c = b; c < a; c++) s += Lf[tril(a, c)] * Li[tril(c, b)];
      Li[tril(a, b)] = -dinv[a] * s;
    }
  }
  return ok;
}

// Smallest eigenvalue of a symmetric M x M matrix (packed upper) by cyclic Jacobi sweeps.
// Rare path (only when Guu is within 1e-8 of singular or indefinite): kept out of line.
template <typename T, int M>
__device__ __noinline__ T min_eig_jacobi(const T* Gs) {
  T Amat[M][M];
  for (int i = 0; i < M; i++)
    for (int j = 0; j < M; j++) Amat[i][j] = Gs[triu(i, j, M)];
  if (M == 1) return Amat[0][0];
  for (int sweep = 0; sweep < 30; sweep++) {
    T off = T(0), diag = T(0);
    for (int i = 0; i < M; i++) {
      diag += Amat[i][i] * Amat[i][i];
      for (int j = i + 1; j < M; j++) off += Amat[i][j] * Amat[i][j];
    }
    const T eps = sizeof(T) == 8 ? T(1e-30) : T(1e-14);
    if (!(off > eps * (diag + off))) break;
    for (int p = 0; p < M - 1; p++)
      for (int q = p + 1; q < M; q++) {
        const T apq = Amat[p][q];
        if (apq == T(0)) continue;
        const T theta = (Amat[q][q] - Amat[p][p]) / (T(2) * apq);
        const T tt = (theta >= T(0) ? T(1) : T(-1)) / (abs_(theta) + sqrt(theta * theta + T(1)));
        const T c = T(1) / sqrt(tt * tt + T(1)), s = tt * c;
        for (int k = 0; k < M; k++) {  // A <- A J
          const T akp = Amat[k][p], akq = Amat[k][q];
          Amat[k][p] = c * akp - s * akq;
          Amat[k][q] = s * akp + c * akq;
        }
        for (int k = 0; k < M; k++) {  // A <- J^T A
          const T apk = Amat[p][k], aqk = Amat[q][k];
          Amat[p][k] = c * apk - s * aqk;
          Amat[q][k] = s * apk + c * aqk;
        }
      }
  }
  T mn = Amat[0][0];
  for (int i = 1; i < M; i++) mn = Amat[i][i] < mn ? Amat[i][i] : mn;
  return mn;
}

}  // namespace idoc
