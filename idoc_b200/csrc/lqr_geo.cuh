// Compile-time geometry of the time-varying LQR kernels: slab layouts in shared memory, the
// residual ("workspace") slot written by the Riccati sweep and consumed by the rollout and the VJP.
// Problem pytree / shapes follow the reference: idoc/lqr.py:23-35 (LQR), :55-59 (Params), with a leading
// batch axis:  Q[B,T,n,n] q[B,T,n] Qf[B,n,n] qf[B,n] M[B,T,n,m] R[B,T,m,m] r[B,T,m] A[B,T,n,n] B[B,T,n,m]
// d[B,T,n] x0[B,n];  X[B,T,n] (X[t] = x_{t+1}), U[B,T,m], Nu[B,T,n]  (idoc/typs.py:6-10).
#pragma once
#include "common.cuh"

namespace idoc {

template <typename T, int N, int M>
struct Geo {
  static constexpr int SZ = sizeof(T);
  static constexpr int E16 = 16 / SZ;
  static constexpr int NP = tri(N), MP = tri(M);
  static constexpr int p16(int w) { return rup(w, E16); }

  // per-(b,t) sizes of the problem arrays, elements
  static constexpr int wA = N * N, wB = N * M, wQ = N * N, wM = N * M, wR = M * M, wq = N, wr = M, wd = N;
  static constexpr int PER_STEP = wA + wB + wQ + wM + wR + wq + wr + wd;  // 2n^2+2nm+m^2+2n+m  (228 at n=8,m=4)
  static constexpr int ONCE = N * N + 2 * N;                              // Qf, qf, x0
  static constexpr int SOL_STEP = 2 * N + M;                              // X, Nu, U

  // copy granularity (bytes) that divides every per-step array size: 4, 8 or 16
  static constexpr int GRE = cgcd(cgcd(cgcd(N * N, N * M), cgcd(M * M, N)), cgcd(M, E16));
  static constexpr int GR = GRE * SZ;

  // ---- residual slot ws[b][t][WS]  (written by the Riccati sweep at step t)
  //   Linv : inverse Cholesky factor of Gtuu_t (packed lower, row-major)         MP
  //   s    : eigen-shift taken at step t (lqr.py:88), 0 in the regular case       1
  //   K,k  : gains of step t (lqr.py:89-90)                                       M*N, M
  //   v,Vp : value function ENTERING step t, i.e. (v_{t+1}, V_{t+1}) packed upper N, NP
  static constexpr int oLinv = 0;
  static constexpr int oS = MP;
  static constexpr int oK = p16(MP + 1);
  static constexpr int ok = oK + p16(M * N);
  static constexpr int ov = ok + p16(M);
  static constexpr int oVp = ov + p16(N);
  static constexpr int WS = oVp + p16(NP);

  // ---- VJP slot ws2[b][t][WS2]: k'_t and v'_{t+1} of the adjoint (vector-only) backward sweep
  static constexpr int o2k = 0;
  static constexpr int o2v = p16(M);
  static constexpr int WS2 = o2v + p16(N);
};

// per-problem status bits (mirrored as IDOC_STATUS_* in include/idoc_b200.h)
constexpr int ST_SHIFT = 1, ST_NONFINITE = 2, ST_CHOL_CLAMP = 4;

}  // namespace idoc
