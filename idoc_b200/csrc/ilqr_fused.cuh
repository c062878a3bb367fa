// Fused iLQR for small systems: the whole outer loop of idoc/ilqr.py:221-268 (linearise -> Riccati sweep -> backtracking
// line search -> update, until the relative cost change falls below thres) as ONE kernel with one thread per problem.
//
// The multi-kernel loop in ilqr_api.cu writes the LQR leaves of every iteration to HBM (linearise), reads them back in
// the Riccati sweep, and launches up to ls_maxiter rollouts + bookkeeping kernels per iteration.  For the small
// systems of BASELINE config 4 (pendulum: n = 2, m = 1) every block fits a thread's registers, so here the
// linearisation is produced inside the time-reversed sweep (no leaf ever touches memory), the
// gains K_t, k_t and the two trajectories (current / trial) live in a per-problem scratch laid out [t][i][B] (coalesced
// across the threads of a warp, L2-resident), and problems that converge early simply leave the loop.
//
// Same arithmetic as the unfused path:  backward step = idoc/lqr.py:77-95 (eigen-shift, unshifted-Guu value update,
// expected-change terms), rollout = idoc/ilqr.py:196-219, line search = idoc/line_search.py:9-40 (last trial accepted on
// failure), convergence = ilqr.py:246-252, co-states = lqr.py:110-125 on the global approximation (ilqr.py:266-267).
#pragma once
#include "ilqr.cuh"
#include "lqr_fwd.cuh"  // IDOC_EPS
#include "small_linalg.cuh"

namespace idoc {

template <typename T>
struct IlqrFusedArgs {
  const T *theta, *x0;
  T *X, *U, *Nu;      // in: initial trajectory (X, U); out: solution and co-states, [B,T,n] / [B,T,m]
  T* scratch;         // B * T * (2n + 2m + mn + m) scalars
  const T *u_lo, *u_hi;  // optional control limits [B, m]: control-limited iLQR (Tassa et al. 2014), no reference
                         // counterpart (BASELINE config 4 asks for box control limits; idoc itself has none: SURVEY D1)
  int *iters, *ls_failed;
  long long nb;
  int T_, theta_dim, maxiter, ls_maxiter, use_ls;
  int box_ls;  // line-search rule of bound-active iterations: 0 reference ratio window, 1 Armijo, 2 window + strict decrease
  T thres, ls_lower, ls_upper, ls_rho;
};

__host__ __device__ constexpr size_t ilqr_fused_scratch_elems(int T, int n, int m) {
  return (size_t)T * (2 * n + 2 * m + m * n + m);
}

// TD > 0: theta (TD scalars) is copied into the thread's registers; TD = 0: read through the global pointer
// x = Li^T Li b  (Li: packed lower inverse Cholesky factor, i.e. x = G^{-1} b)
template <typename T, int M>
IDOC_DEVICE void li_solve(const T* Li, const T* b, T* x) {
  T y[M];
#pragma unroll
  for (int i = 0; i < M; i++) {
    T s = T(0);
#pragma unroll
    for (int c = 0; c <= i; c++) s += Li[tril(i, c)] * b[c];
    y[i] = s;
  }
#pragma unroll
  for (int i = 0; i < M; i++) {
    T s = T(0);
#pragma unroll
    for (int c = i; c < M; c++) s += Li[tril(c, i)] * y[c];
    x[i] = s;
  }
}

// inverse Cholesky factor of G (packed upper) with the rows / columns of the components with st[i] != 0 replaced by the identity
template <typename T, int M>
IDOC_DEVICE void masked_chol_inv(const T* G, const int* st, T* Lim) {
  constexpr int MP = M * (M + 1) / 2;
  T Gm[MP];
#pragma unroll
  for (int i = 0; i < M; i++)
#pragma unroll
    for (int j = i; j < M; j++) Gm[triu(i, j, M)] = (st[i] != 0 || st[j] != 0) ? (i == j ? T(1) : T(0)) : G[triu(i, j, M)];
  bool cl = false;
  chol_inv<T, M, true>(Gm, Lim, T(IDOC_EPS) * T(1e-3), &cl);
}

// Box QP of one backward step of control-limited iLQR (Tassa, Mansard, Todorov 2014; no reference counterpart):
//     min_du  1/2 du^T G du + g^T du     s.t.  lo <= du <= hi          (G = Guu + shift I, positive definite, packed upper)
// by a primal active-set iteration on masked systems: clamp every free component that leaves the box, otherwise release
// the clamped component whose multiplier has the wrong sign most, re-solve; at most 3 M + 1 re-solves.  du enters as the
// unconstrained minimiser.  st[i] = -1 / 0 / +1: at the lower bound / free / at the upper bound; Lim: inverse Cholesky
// factor of the masked matrix (for the feedback gains of the free components).  Returns whether any component is clamped.
template <typename T, int M>
IDOC_DEVICE bool box_qp(const T* G, const T* g, const T* lo, const T* hi, T* du, int* st, T* Lim) {
  const T tol = sizeof(T) == 4 ? T(1e-5) : T(1e-11);
#pragma unroll
  for (int i = 0; i < M; i++) st[i] = 0;
  bool any = false;
  for (int it = 0; it < 3 * M + 1; it++) {
    bool changed = false;
#pragma unroll
    for (int i = 0; i < M; i++)
      if (st[i] == 0) {
        if (du[i] < lo[i]) { st[i] = -1; changed = true; }
        else if (du[i] > hi[i]) { st[i] = 1; changed = true; }
      }
    if (!changed) {
      int worst = -1;
      T wv = T(0);
#pragma unroll
      for (int i = 0; i < M; i++)
        if (st[i] != 0) {
          T lam = g[i];
#pragma unroll
          for (int j = 0; j < M; j++) lam += G[i <= j ? triu(i, j, M) : triu(j, i, M)] * du[j];
          const T viol = st[i] < 0 ? -lam : lam;  // lower bound: multiplier >= 0; upper bound: <= 0
          if (viol > tol * (T(1) + (g[i] < T(0) ? -g[i] : g[i])) && viol > wv) { wv = viol; worst = i; }
        }
      if (worst < 0) break;
#pragma unroll
      for (int i = 0; i < M; i++)
        if (i == worst) st[i] = 0;
    }
    any = false;
    T rhs[M];
#pragma unroll
    for (int i = 0; i < M; i++) {
      if (st[i] != 0) {
        rhs[i] = st[i] < 0 ? lo[i] : hi[i];
        any = true;
      } else {
        T r = -g[i];
#pragma unroll
        for (int j = 0; j < M; j++)
          if (st[j] != 0) r -= G[i <= j ? triu(i, j, M) : triu(j, i, M)] * (st[j] < 0 ? lo[j] : hi[j]);
        rhs[i] = r;
      }
    }
    masked_chol_inv<T, M>(G, st, Lim);
    li_solve<T, M>(Lim, rhs, du);
  }
#pragma unroll
  for (int i = 0; i < M; i++)
    if (st[i] != 0) du[i] = st[i] < 0 ? lo[i] : hi[i];
  return any;
}

template <typename P, typename T, int N, int M, int TD>
__global__ void __launch_bounds__(32) ilqr_fused_kernel(const IlqrFusedArgs<T> a) {
  const long long b = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= a.nb) return;
  const long long B = a.nb;
  const int Th = a.T_;
  constexpr int MP = tri(M);
  T thl[TD > 0 ? TD : 1];
  if (TD > 0) {
#pragma unroll
    for (int e = 0; e < TD; e++) thl[e] = a.theta[b * a.theta_dim + e];
  }
  const T* th = TD > 0 ? thl : a.theta + b * a.theta_dim;

  // scratch sections (each element strided by B): two trajectory buffers, gains
  T* S = a.scratch + b;
  const long long oX[2] = {0, (long long)Th * N};
  const long long oU[2] = {2LL * Th * N, 2LL * Th * N + (long long)Th * M};
  const long long oK = 2LL * Th * (N + M), ok = oK + (long long)Th * M * N;
  auto ld = [&](long long off) -> T { return S[off * B]; };
  auto st = [&](long long off, T v) { S[off * B] = v; };

  T x0[N];
#pragma unroll
  for (int i = 0; i < N; i++) x0[i] = a.x0[b * N + i];
  for (int t = 0; t < Th; t++) {
#pragma unroll
    for (int i = 0; i < N; i++) st(oX[0] + t * N + i, a.X[(b * Th + t) * N + i]);
#pragma unroll
    for (int j = 0; j < M; j++) st(oU[0] + t * M + j, a.U[(b * Th + t) * M + j]);
  }
  int cur = 0;

  // c_old: cost of the open-loop rollout of U from x0 (ilqr.py:255-256)
  T c;
  {
    T x[N], xn[N], u[M];
#pragma unroll
    for (int i = 0; i < N; i++) x[i] = x0[i];
    c = T(0);
    for (int t = 0; t < Th; t++) {
#pragma unroll
      for (int j = 0; j < M; j++) u[j] = ld(oU[0] + t * M + j);
      c += P::cost(N, M, th, x, u);
      P::dyn(N, M, th, x, u, xn);
#pragma unroll
      for (int i = 0; i < N; i++) x[i] = xn[i];
    }
    c += P::costf(N, M, th, x);
  }

  int iters = 0, ls_failed = 0;
  for (int it = 0; it < a.maxiter; it++) {
    // ---------------------------------------------------------------- backward sweep on the local approximation
    T V[N][N], v[N];
    T dC = T(0), dc = T(0);
    bool bound_active = false;  // control limits: some step of this backward sweep sits on a bound
    {
      T xf[N], Qf[N * N];
#pragma unroll
      for (int i = 0; i < N; i++) xf[i] = ld(oX[cur] + (Th - 1) * N + i);
      P::final_derivs(N, M, th, xf, v, Qf, N);  // local form: qf = lf_x
#pragma unroll
      for (int i = 0; i < N; i++)
#pragma unroll
        for (int j = 0; j < N; j++) V[i][j] = T(0.5) * (Qf[i * N + j] + Qf[j * N + i]);
    }
    // (x_t, u_t) of the NEXT step of the sweep are requested before the current step's arithmetic: the trajectory
    // loads do not depend on the recursion, so their L2 latency hides behind it
    T xp[N], up[M];
#pragma unroll
    for (int i = 0; i < N; i++) xp[i] = Th == 1 ? x0[i] : ld(oX[cur] + (Th - 2) * N + i);
#pragma unroll
    for (int j = 0; j < M; j++) up[j] = ld(oU[cur] + (Th - 1) * M + j);
    for (int t = Th - 1; t >= 0; t--) {
      T x[N], u[M], lx[N], lu[M], f[N], Q[N * N], R[M * M], Mx[N * M], A[N * N], Bm[N * M];
#pragma unroll
      for (int i = 0; i < N; i++) x[i] = xp[i];
#pragma unroll
      for (int j = 0; j < M; j++) {
        u[j] = up[j];
        lu[j] = T(0);
      }
      if (t > 0) {
#pragma unroll
        for (int i = 0; i < N; i++) xp[i] = t == 1 ? x0[i] : ld(oX[cur] + (t - 2) * N + i);
#pragma unroll
        for (int j = 0; j < M; j++) up[j] = ld(oU[cur] + (t - 1) * M + j);
      }
#pragma unroll
      for (int e = 0; e < M * M; e++) R[e] = T(0);
      P::derivs(N, M, th, x, u, (const T*)nullptr, lx, lu, Q, N, R, Mx, A, N, Bm, f);
      // W = V [A B]
      T WA[N][N], WB[N][M];
#pragma unroll
      for (int k = 0; k < N; k++) {
#pragma unroll
        for (int j = 0; j < N; j++) {
          T s = T(0);
#pragma unroll
          for (int l = 0; l < N; l++) s += V[k][l] * A[l * N + j];
          WA[k][j] = s;
        }
#pragma unroll
        for (int j = 0; j < M; j++) {
          T s = T(0);
#pragma unroll
          for (int l = 0; l < N; l++) s += V[k][l] * Bm[l * M + j];
          WB[k][j] = s;
        }
      }
      T Gxx[N][N], Gxu[N][M], Guu[M][M], gx[N], gu[M];
#pragma unroll
      for (int i = 0; i < N; i++) {
        T s = lx[i];
#pragma unroll
        for (int k = 0; k < N; k++) s += A[k * N + i] * v[k];
        gx[i] = s;
#pragma unroll
        for (int j = 0; j < N; j++) {
          T g = T(0.5) * (Q[i * N + j] + Q[j * N + i]);
#pragma unroll
          for (int k = 0; k < N; k++) g += A[k * N + i] * WA[k][j];
          Gxx[i][j] = g;
        }
#pragma unroll
        for (int j = 0; j < M; j++) {
          T g = Mx[i * M + j];
#pragma unroll
          for (int k = 0; k < N; k++) g += A[k * N + i] * WB[k][j];
          Gxu[i][j] = g;
        }
      }
#pragma unroll
      for (int i = 0; i < M; i++) {
        T s = lu[i];
#pragma unroll
        for (int k = 0; k < N; k++) s += Bm[k * M + i] * v[k];
        gu[i] = s;
#pragma unroll
        for (int j = 0; j < M; j++) {
          T g = T(0.5) * (R[i * M + j] + R[j * M + i]);
#pragma unroll
          for (int k = 0; k < N; k++) g += Bm[k * M + i] * WB[k][j];
          Guu[i][j] = g;
        }
      }
      T Gs[MP], Li[MP];
#pragma unroll
      for (int i = 0; i < M; i++)
#pragma unroll
        for (int j = i; j < M; j++) Gs[triu(i, j, M)] = T(0.5) * (Guu[i][j] + Guu[j][i]);
      T shift = T(0);
      {
        bool clamped = false;
        const bool ok = chol_inv<T, M, false>(Gs, Li, T(0), &clamped);
        T fro = T(0);
#pragma unroll
        for (int e = 0; e < MP; e++) fro += Li[e] * Li[e];
        if (!ok || !(fro < T(1.0 / IDOC_EPS))) {  // lqr.py:86-88
          T Gc[MP];
#pragma unroll
          for (int e = 0; e < MP; e++) Gc[e] = Gs[e];
          const T mn = min_eig_jacobi<T, M>(Gc);
          shift = T(IDOC_EPS) - mn;
          shift = shift > T(0) ? shift : T(0);
#pragma unroll
          for (int i = 0; i < M; i++) Gc[triu(i, i, M)] += shift;
          chol_inv<T, M, true>(Gc, Li, T(IDOC_EPS) * T(1e-3), &clamped);
        }
      }
      // k = -Gt^{-1} gu, K = -Gt^{-1} Gux   (Gt^{-1} = Li^T Li)
      T kk[M], K[M][N];
      {
        T y[M];
#pragma unroll
        for (int i = 0; i < M; i++) {
          T s = T(0);
#pragma unroll
          for (int c2 = 0; c2 <= i; c2++) s += Li[tril(i, c2)] * gu[c2];
          y[i] = s;
        }
#pragma unroll
        for (int i = 0; i < M; i++) {
          T s = T(0);
#pragma unroll
          for (int c2 = i; c2 < M; c2++) s += Li[tril(c2, i)] * y[c2];
          kk[i] = -s;
        }
      }
#pragma unroll
      for (int j = 0; j < N; j++) {
        T y[M];
#pragma unroll
        for (int i = 0; i < M; i++) {
          T s = T(0);
#pragma unroll
          for (int c2 = 0; c2 <= i; c2++) s += Li[tril(i, c2)] * Gxu[j][c2];
          y[i] = s;
        }
#pragma unroll
        for (int i = 0; i < M; i++) {
          T s = T(0);
#pragma unroll
          for (int c2 = i; c2 < M; c2++) s += Li[tril(c2, i)] * y[c2];
          K[i][j] = -s;
        }
      }
      // control limits: the step is the box QP  min 1/2 du^T Gt du + gu^T du, lo - u <= du <= hi - u; a clamped component has
      // no feedback (Tassa et al. 2014).  m = 1: the QP is a clamp; m > 1: active-set iteration (box_qp)
      bool clamped = false;  // every control clamped (m = 1), K = 0
      bool partial = false;  // m > 1: some component clamped
      if (M == 1 && a.u_lo != nullptr) {
        const T lo = a.u_lo[b * M] - u[0], hi = a.u_hi[b * M] - u[0];
        const T kc = kk[0] < lo ? lo : (kk[0] > hi ? hi : kk[0]);
        clamped = kc != kk[0];
        kk[0] = kc;
        if (clamped) {
#pragma unroll
          for (int j = 0; j < N; j++) K[0][j] = T(0);
        }
      }
      if (M > 1 && a.u_lo != nullptr) {
        T lo[M], hi[M], Gt[MP], Lim[MP];
        int stt[M];
#pragma unroll
        for (int i = 0; i < M; i++) { lo[i] = a.u_lo[b * M + i] - u[i]; hi[i] = a.u_hi[b * M + i] - u[i]; }
#pragma unroll
        for (int i = 0; i < M; i++)
#pragma unroll
          for (int j = i; j < M; j++) Gt[triu(i, j, M)] = Gs[triu(i, j, M)] + (i == j ? shift : T(0));
        partial = box_qp<T, M>(Gt, gu, lo, hi, kk, stt, Lim);
        if (partial) {  // gains of the free components: K_f = -Gt_ff^{-1} Gux_f, K_c = 0
#pragma unroll
          for (int j = 0; j < N; j++) {
            T rhs[M], col[M];
#pragma unroll
            for (int i = 0; i < M; i++) rhs[i] = stt[i] != 0 ? T(0) : -Gxu[j][i];
            li_solve<T, M>(Lim, rhs, col);
#pragma unroll
            for (int i = 0; i < M; i++) K[i][j] = stt[i] != 0 ? T(0) : col[i];
          }
        }
      }
      bound_active = bound_active || clamped || partial;
      {  // expected-change terms with the unshifted Guu (lqr.py:93-94)
        T quad = T(0), lin = T(0);
#pragma unroll
        for (int i = 0; i < M; i++) {
          T s = T(0);
#pragma unroll
          for (int j = 0; j < M; j++) s += Gs[triu(i, j, M)] * kk[j];
          quad += s * kk[i];
          lin += gu[i] * kk[i];
        }
        dC += T(0.5) * quad;
        dc += lin;
      }
      // v = gx + K^T (gu - s k),  V = Gxx + Gxu K - s K^T K  (upper triangle, mirrored); a clamped step (K = 0, k at the
      // bound) takes the general form of lqr.py:91-92 instead: v = gx + Gxu k, V = Gxx
      if (M > 1 && partial) {
        // partly clamped step: the general form of lqr.py:91-92 with the unshifted Guu,
        //   v = gx + Gxu k + K^T (gu + Guu k),   V = Gxx + Gxu K + K^T Gux + K^T Guu K
        T Gk[M], GK[M][N];
#pragma unroll
        for (int i = 0; i < M; i++) {
          T sk = gu[i];
#pragma unroll
          for (int l = 0; l < M; l++) sk += Gs[i <= l ? triu(i, l, M) : triu(l, i, M)] * kk[l];
          Gk[i] = sk;
#pragma unroll
          for (int j = 0; j < N; j++) {
            T sK = T(0);
#pragma unroll
            for (int l = 0; l < M; l++) sK += Gs[i <= l ? triu(i, l, M) : triu(l, i, M)] * K[l][j];
            GK[i][j] = sK;
          }
        }
#pragma unroll
        for (int j = 0; j < N; j++) {
          T sv = gx[j];
#pragma unroll
          for (int i = 0; i < M; i++) sv += Gxu[j][i] * kk[i] + K[i][j] * Gk[i];
          v[j] = sv;
        }
#pragma unroll
        for (int i = 0; i < N; i++)
#pragma unroll
          for (int j = i; j < N; j++) {
            T sV = Gxx[i][j];
#pragma unroll
            for (int c2 = 0; c2 < M; c2++) sV += Gxu[i][c2] * K[c2][j] + K[c2][i] * Gxu[j][c2] + K[c2][i] * GK[c2][j];
            V[i][j] = sV;
            V[j][i] = sV;
          }
      } else {
#pragma unroll
      for (int j = 0; j < N; j++) {
        T s = gx[j];
        if (clamped) {
#pragma unroll
          for (int i = 0; i < M; i++) s += Gxu[j][i] * kk[i];
        } else {
#pragma unroll
          for (int i = 0; i < M; i++) s += K[i][j] * (gu[i] - shift * kk[i]);
        }
        v[j] = s;
      }
#pragma unroll
      for (int i = 0; i < N; i++)
#pragma unroll
        for (int j = i; j < N; j++) {
          T s = Gxx[i][j];
#pragma unroll
          for (int c2 = 0; c2 < M; c2++) s += Gxu[i][c2] * K[c2][j] - shift * K[c2][i] * K[c2][j];
          V[i][j] = s;
          V[j][i] = s;
        }
      }
#pragma unroll
      for (int i = 0; i < M; i++) {
        st(ok + t * M + i, kk[i]);
#pragma unroll
        for (int j = 0; j < N; j++) st(oK + (t * M + i) * N + j, K[i][j]);
      }
    }
    // ---------------------------------------------------------------- backtracking line search over closed-loop rollouts
    const int nxt = cur ^ 1;
    T alpha = T(1), cn = T(0);
    int ls_it = 0;
    for (;;) {
      T xh[N], xn[N], uh[M];
#pragma unroll
      for (int i = 0; i < N; i++) xh[i] = x0[i];
      T l = T(0);
      // gains and reference trajectory of step t+1 are requested while step t is computed (they do not depend on xh)
      T kp[M], Kp[M][N], xop[N], uop[M];
      auto fetch = [&](int t) {
#pragma unroll
        for (int j = 0; j < M; j++) {
          kp[j] = ld(ok + t * M + j);
          uop[j] = ld(oU[cur] + t * M + j);
#pragma unroll
          for (int i = 0; i < N; i++) Kp[j][i] = ld(oK + (t * M + j) * N + i);
        }
#pragma unroll
        for (int i = 0; i < N; i++) xop[i] = t == 0 ? x0[i] : ld(oX[cur] + (t - 1) * N + i);
      };
      fetch(0);
      for (int t = 0; t < Th; t++) {
        T kc[M], Kc[M][N], xo[N], uo[M];
#pragma unroll
        for (int j = 0; j < M; j++) {
          kc[j] = kp[j];
          uo[j] = uop[j];
#pragma unroll
          for (int i = 0; i < N; i++) Kc[j][i] = Kp[j][i];
        }
#pragma unroll
        for (int i = 0; i < N; i++) xo[i] = xop[i];
        if (t + 1 < Th) fetch(t + 1);
#pragma unroll
        for (int j = 0; j < M; j++) {
          T du = alpha * kc[j];
#pragma unroll
          for (int i = 0; i < N; i++) du += Kc[j][i] * (xh[i] - xo[i]);
          uh[j] = uo[j] + du;
          if (a.u_lo != nullptr) uh[j] = uh[j] < a.u_lo[b * M + j] ? a.u_lo[b * M + j] : (uh[j] > a.u_hi[b * M + j] ? a.u_hi[b * M + j] : uh[j]);
          st(oU[nxt] + t * M + j, uh[j]);
        }
        l += P::cost(N, M, th, xh, uh);
        P::dyn(N, M, th, xh, uh, xn);
#pragma unroll
        for (int i = 0; i < N; i++) {
          xh[i] = xn[i];
          st(oX[nxt] + t * N + i, xn[i]);
        }
      }
      l += P::costf(N, M, th, xh);
      cn = l;
      bool carry_on = false;
      if (a.use_ls) {  // line_search.py:12-27
        const T ec = alpha * alpha * dC + alpha * dc;
        const T p = abs_((c - cn + T(1e-10)) / (T(1e-10) + ec));
        carry_on = !(p > a.ls_lower && p < a.ls_upper);  // NaN ratio (overflowed trial rollout) backtracks too
        // Control limits with an active bound (no reference counterpart): the rollout clamps u + alpha k + K dx again, so the
        // realised change falls short of the quadratic model, and the |actual / expected| window also admits cost INCREASES.
        // Two monotone alternatives are compiled in for iterations with an active bound (IDOC_BOX_LS): 1 = sufficient
        // decrease (Armijo, 1e-4 of the expected change), 2 = the window AND a strict decrease.  Measured at B = 16384
        // (profiles/r2_box_line_search.jsonl): they remove the iteration-cap stalls (cart-pole |u| <= 6: 1909 -> 0 / 2 problems
        // at the cap, pendulum |u| <= 8: 72 -> 0 / 5, 3.4x the solves/s) but on the cart-pole they END IN FAR WORSE OPTIMA
        // (median final cost 612 against 1.9; 2.3 - 2.8 k of 16 384 problems end upright against 14.1 k): the reference's
        // non-monotone rule is what lets the early, heavily clamped iterates escape the bang-bang basin.  Default: 0, the
        // reference's rule everywhere.
        if (a.u_lo != nullptr && bound_active) {
          if (a.box_ls == 1) carry_on = !(cn <= c + T(1e-4) * ec);
          else if (a.box_ls == 2) carry_on = carry_on || !(cn < c);
        }
      }
      ls_it++;
      if (a.use_ls && carry_on && ls_it < a.ls_maxiter) {
        alpha *= a.ls_rho;
        continue;
      }
      if (carry_on) ls_failed++;  // the last tried trajectory is taken even if it failed (line_search.py:29-40)
      break;
    }
    cur = nxt;
    const T pct = abs_((c - cn) / c);  // ilqr.py:246
    c = cn;
    iters++;
    if (!(pct > a.thres)) break;
  }
  // ---------------------------------------------------------------- solution + co-states (lqr.py:110-125 on the global approx)
  for (int t = 0; t < Th; t++) {
#pragma unroll
    for (int i = 0; i < N; i++) a.X[(b * Th + t) * N + i] = ld(oX[cur] + t * N + i);
#pragma unroll
    for (int j = 0; j < M; j++) a.U[(b * Th + t) * M + j] = ld(oU[cur] + t * M + j);
  }
  {
    T nu[N], xf[N], Qf[N * N];
#pragma unroll
    for (int i = 0; i < N; i++) xf[i] = ld(oX[cur] + (Th - 1) * N + i);
    P::final_derivs(N, M, th, xf, nu, Qf, N);  // Nu[T-1] = Qf x_T + qf = lf_x(x_T)
#pragma unroll
    for (int i = 0; i < N; i++) a.Nu[(b * Th + Th - 1) * N + i] = nu[i];
    for (int t = Th - 1; t >= 1; t--) {  // Nu[t-1] = A_t^T Nu[t] + Q_t x_t + M_t u_t + q_t = l_x(x_t, u_t) + A_t^T Nu[t]
      T x[N], u[M], lx[N], lu[M], f[N], Q[N * N], R[M * M], Mx[N * M], A[N * N], Bm[N * M], nn[N];
#pragma unroll
      for (int i = 0; i < N; i++) x[i] = ld(oX[cur] + (t - 1) * N + i);
#pragma unroll
      for (int j = 0; j < M; j++) {
        u[j] = ld(oU[cur] + t * M + j);
        lu[j] = T(0);
      }
#pragma unroll
      for (int e = 0; e < M * M; e++) R[e] = T(0);
      P::derivs(N, M, th, x, u, (const T*)nullptr, lx, lu, Q, N, R, Mx, A, N, Bm, f);
#pragma unroll
      for (int i = 0; i < N; i++) {
        T s = lx[i];
#pragma unroll
        for (int k = 0; k < N; k++) s += A[k * N + i] * nu[k];
        nn[i] = s;
      }
#pragma unroll
      for (int i = 0; i < N; i++) {
        nu[i] = nn[i];
        a.Nu[(b * Th + t - 1) * N + i] = nn[i];
      }
    }
  }
  a.iters[b] = iters;
  a.ls_failed[b] = ls_failed;
}

}  // namespace idoc

// ================================================================================================ fused implicit VJP
// The implicit-function VJP of the iLQR solution for the same small systems, one thread per problem, ONE kernel:
// the multi-kernel path (ilqr_api.cu: vjp_impl) linearises the Hessian of the Lagrangian into HBM leaves, runs the tuned
// Riccati sweep on them, the two adjoint sweeps, and a theta pull-back kernel over the 10 leaf cotangents.  Here the
// adjoint blocks are produced inside the sweeps (mode-2 derivatives: dynamics curvature weighted by Nu), the backward
// sweep carries the Riccati recursion AND the adjoint vector recursion (v', k') together, and the forward sweep
// accumulates the theta cotangent in registers from the per-step leaf cotangents, which never exist in memory.
// Same arithmetic as csrc/lqr_vjp.cuh and ilqr_theta_vjp_kernel.
namespace idoc {

template <typename T>
struct IlqrFusedVjpArgs {
  const T *theta, *x0, *X, *U, *Nu;  // solution
  const T *Xb, *Ub, *Nub;            // cotangents (Nub may be null)
  T *gtheta, *gx0;                   // [B, theta_dim], [B, n]
  T* scratch;                        // B * T * (mn + m + n*n + n) scalars
  const T *u_lo, *u_hi;              // optional control limits [B, m]: controls at a bound are an active set
  long long nb;
  int T_, theta_dim;
};

__host__ __device__ constexpr size_t ilqr_fused_vjp_scratch_elems(int T, int n, int m) {
  return (size_t)T * (m * n + m + n * n + n);
}

template <typename P, typename T, int N, int M, int TD>
__global__ void __launch_bounds__(32) ilqr_fused_vjp_kernel(const IlqrFusedVjpArgs<T> a) {
  const long long b = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= a.nb) return;
  const long long B = a.nb;
  const int Th = a.T_;
  constexpr int MP = tri(M), TDR = TD > 0 ? TD : 1;
  T thl[TDR];
  if (TD > 0) {
#pragma unroll
    for (int e = 0; e < TD; e++) thl[e] = a.theta[b * a.theta_dim + e];
  }
  const T* th = TD > 0 ? thl : a.theta + b * a.theta_dim;
  T* S = a.scratch + b;
  constexpr int WSTEP = M * N + M + N * N + N;  // per step: K | k' | V_{t+1} | v'_{t+1}
  auto ld = [&](long long off) -> T { return S[off * B]; };
  auto st = [&](long long off, T v) { S[off * B] = v; };
  const T* Xg = a.X + b * Th * N;
  const T* Ug = a.U + b * Th * M;
  const T* Ng = a.Nu + b * Th * N;
  T x0[N];
#pragma unroll
  for (int i = 0; i < N; i++) x0[i] = a.x0[b * N + i];

  // ---------------------------------------------------------------- backward: Riccati on the adjoint blocks + (v', k')
  T V[N][N], vp[N];
  {
    T xf[N], lfx[N], Qf[N * N];
#pragma unroll
    for (int i = 0; i < N; i++) xf[i] = Xg[(Th - 1) * N + i];
    P::final_derivs(N, M, th, xf, lfx, Qf, N);
#pragma unroll
    for (int i = 0; i < N; i++) {
      vp[i] = a.Xb[(b * Th + Th - 1) * N + i];  // v'_T = qf' = Xb[T-1]
#pragma unroll
      for (int j = 0; j < N; j++) V[i][j] = T(0.5) * (Qf[i * N + j] + Qf[j * N + i]);
    }
  }
  for (int t = Th - 1; t >= 0; t--) {
    T x[N], u[M], nu[N], lx[N], lu[M], f[N], Q[N * N], R[M * M], Mx[N * M], A[N * N], Bm[N * M];
#pragma unroll
    for (int i = 0; i < N; i++) {
      x[i] = t == 0 ? x0[i] : Xg[(t - 1) * N + i];
      nu[i] = Ng[t * N + i];
    }
#pragma unroll
    for (int j = 0; j < M; j++) { u[j] = Ug[t * M + j]; lu[j] = T(0); }
#pragma unroll
    for (int e = 0; e < M * M; e++) R[e] = T(0);
    P::derivs(N, M, th, x, u, nu, lx, lu, Q, N, R, Mx, A, N, Bm, f);
    const long long o = (long long)t * WSTEP;
#pragma unroll
    for (int i = 0; i < N; i++) {
      st(o + M * N + M + N * N + i, vp[i]);
#pragma unroll
      for (int j = 0; j < N; j++) st(o + M * N + M + i * N + j, V[i][j]);
    }
    T WA[N][N], WB[N][M];
#pragma unroll
    for (int k = 0; k < N; k++) {
#pragma unroll
      for (int j = 0; j < N; j++) {
        T s = T(0);
#pragma unroll
        for (int l = 0; l < N; l++) s += V[k][l] * A[l * N + j];
        WA[k][j] = s;
      }
#pragma unroll
      for (int j = 0; j < M; j++) {
        T s = T(0);
#pragma unroll
        for (int l = 0; l < N; l++) s += V[k][l] * Bm[l * M + j];
        WB[k][j] = s;
      }
    }
    // w = v' + V d', gx' = q' + A^T w, gu' = r' + B^T w  with q' = Xb[t-1] (0 at t = 0), r' = Ub[t], d' = Nub[t]
    T w[N], gx[N], gu[M];
#pragma unroll
    for (int i = 0; i < N; i++) {
      T s = vp[i];
      if (a.Nub != nullptr) {
#pragma unroll
        for (int l = 0; l < N; l++) s += V[i][l] * a.Nub[(b * Th + t) * N + l];
      }
      w[i] = s;
    }
    T Gxx[N][N], Gxu[N][M], Guu[M][M];
#pragma unroll
    for (int i = 0; i < N; i++) {
      T s = t > 0 ? a.Xb[(b * Th + t - 1) * N + i] : T(0);
#pragma unroll
      for (int k = 0; k < N; k++) s += A[k * N + i] * w[k];
      gx[i] = s;
#pragma unroll
      for (int j = 0; j < N; j++) {
        T g = T(0.5) * (Q[i * N + j] + Q[j * N + i]);
#pragma unroll
        for (int k = 0; k < N; k++) g += A[k * N + i] * WA[k][j];
        Gxx[i][j] = g;
      }
#pragma unroll
      for (int j = 0; j < M; j++) {
        T g = Mx[i * M + j];
#pragma unroll
        for (int k = 0; k < N; k++) g += A[k * N + i] * WB[k][j];
        Gxu[i][j] = g;
      }
    }
#pragma unroll
    for (int i = 0; i < M; i++) {
      T s = a.Ub[(b * Th + t) * M + i];
#pragma unroll
      for (int k = 0; k < N; k++) s += Bm[k * M + i] * w[k];
      gu[i] = s;
#pragma unroll
      for (int j = 0; j < M; j++) {
        T g = T(0.5) * (R[i * M + j] + R[j * M + i]);
#pragma unroll
        for (int k = 0; k < N; k++) g += Bm[k * M + i] * WB[k][j];
        Guu[i][j] = g;
      }
    }
    T Gs[MP], Li[MP];
#pragma unroll
    for (int i = 0; i < M; i++)
#pragma unroll
      for (int j = i; j < M; j++) Gs[triu(i, j, M)] = T(0.5) * (Guu[i][j] + Guu[j][i]);
    T shift = T(0);
    {
      bool clamped = false;
      const bool ok = chol_inv<T, M, false>(Gs, Li, T(0), &clamped);
      T fro = T(0);
#pragma unroll
      for (int e = 0; e < MP; e++) fro += Li[e] * Li[e];
      if (!ok || !(fro < T(1.0 / IDOC_EPS))) {
        T Gc[MP];
#pragma unroll
        for (int e = 0; e < MP; e++) Gc[e] = Gs[e];
        const T mn = min_eig_jacobi<T, M>(Gc);
        shift = T(IDOC_EPS) - mn;
        shift = shift > T(0) ? shift : T(0);
#pragma unroll
        for (int i = 0; i < M; i++) Gc[triu(i, i, M)] += shift;
        chol_inv<T, M, true>(Gc, Li, T(IDOC_EPS) * T(1e-3), &clamped);
      }
    }
    T kk[M], K[M][N];
    {
      T y[M];
#pragma unroll
      for (int i = 0; i < M; i++) {
        T s = T(0);
#pragma unroll
        for (int c2 = 0; c2 <= i; c2++) s += Li[tril(i, c2)] * gu[c2];
        y[i] = s;
      }
#pragma unroll
      for (int i = 0; i < M; i++) {
        T s = T(0);
#pragma unroll
        for (int c2 = i; c2 < M; c2++) s += Li[tril(c2, i)] * y[c2];
        kk[i] = -s;
      }
    }
#pragma unroll
    for (int j = 0; j < N; j++) {
      T y[M];
#pragma unroll
      for (int i = 0; i < M; i++) {
        T s = T(0);
#pragma unroll
        for (int c2 = 0; c2 <= i; c2++) s += Li[tril(i, c2)] * Gxu[j][c2];
        y[i] = s;
      }
#pragma unroll
      for (int i = 0; i < M; i++) {
        T s = T(0);
#pragma unroll
        for (int c2 = i; c2 < M; c2++) s += Li[tril(c2, i)] * y[c2];
        K[i][j] = -s;
      }
    }
    // a step whose control sits on a bound is an active constraint: u does not vary with the parameters, so the adjoint
    // system has no control freedom there (K = 0, k' = 0, v' = gx', V = Gxx)
    if (M == 1 && a.u_lo != nullptr && (u[0] <= a.u_lo[b * M] || u[0] >= a.u_hi[b * M])) {
      kk[0] = T(0);
#pragma unroll
      for (int j = 0; j < N; j++) K[0][j] = T(0);
    }
    if (M > 1 && a.u_lo != nullptr) {  // per component: the active ones drop out of the step's system (masked re-solve)
      int stt[M];
      bool any = false;
#pragma unroll
      for (int i = 0; i < M; i++) {
        stt[i] = (u[i] <= a.u_lo[b * M + i] || u[i] >= a.u_hi[b * M + i]) ? 1 : 0;
        any = any || stt[i] != 0;
      }
      if (any) {
        T Gt[MP], Lim[MP], rhs[M], col[M];
#pragma unroll
        for (int i = 0; i < M; i++)
#pragma unroll
          for (int j = i; j < M; j++) Gt[triu(i, j, M)] = Gs[triu(i, j, M)] + (i == j ? shift : T(0));
        masked_chol_inv<T, M>(Gt, stt, Lim);
#pragma unroll
        for (int i = 0; i < M; i++) rhs[i] = stt[i] != 0 ? T(0) : -gu[i];
        li_solve<T, M>(Lim, rhs, col);
#pragma unroll
        for (int i = 0; i < M; i++) kk[i] = stt[i] != 0 ? T(0) : col[i];
#pragma unroll
        for (int j = 0; j < N; j++) {
#pragma unroll
          for (int i = 0; i < M; i++) rhs[i] = stt[i] != 0 ? T(0) : -Gxu[j][i];
          li_solve<T, M>(Lim, rhs, col);
#pragma unroll
          for (int i = 0; i < M; i++) K[i][j] = stt[i] != 0 ? T(0) : col[i];
        }
      }
    }
#pragma unroll
    for (int j = 0; j < N; j++) {
      T s = gx[j];
#pragma unroll
      for (int i = 0; i < M; i++) s += K[i][j] * (gu[i] - shift * kk[i]);
      vp[j] = s;
    }
#pragma unroll
    for (int i = 0; i < N; i++)
#pragma unroll
      for (int j = i; j < N; j++) {
        T s = Gxx[i][j];
#pragma unroll
        for (int c2 = 0; c2 < M; c2++) s += Gxu[i][c2] * K[c2][j] - shift * K[c2][i] * K[c2][j];
        V[i][j] = s;
        V[j][i] = s;
      }
#pragma unroll
    for (int i = 0; i < M; i++) {
      st(o + M * N + i, kk[i]);
#pragma unroll
      for (int j = 0; j < N; j++) st(o + i * N + j, K[i][j]);
    }
  }
  // ---------------------------------------------------------------- forward: (ux, uU, uNu) and the theta / x0 cotangents
  T acc[TD > 0 ? TD : 1];
  T* gth = a.gtheta + b * a.theta_dim;
  if (TD > 0) {
#pragma unroll
    for (int e = 0; e < TD; e++) acc[e] = T(0);
  } else {
    for (int e = 0; e < a.theta_dim; e++) gth[e] = T(0);
  }
  T ux[N];
#pragma unroll
  for (int i = 0; i < N; i++) ux[i] = T(0);
  for (int t = 0; t < Th; t++) {
    T x[N], u[M], nu[N], lx[N], lu[M], f[N], Q[N * N], R[M * M], Mx[N * M], A[N * N], Bm[N * M];
#pragma unroll
    for (int i = 0; i < N; i++) {
      x[i] = t == 0 ? x0[i] : Xg[(t - 1) * N + i];
      nu[i] = Ng[t * N + i];
    }
#pragma unroll
    for (int j = 0; j < M; j++) { u[j] = Ug[t * M + j]; lu[j] = T(0); }
#pragma unroll
    for (int e = 0; e < M * M; e++) R[e] = T(0);
    P::derivs(N, M, th, x, u, nu, lx, lu, Q, N, R, Mx, A, N, Bm, f);
    const long long o = (long long)t * WSTEP;
    T uU[M], uxn[N], unu[N];
#pragma unroll
    for (int i = 0; i < M; i++) {
      T s = ld(o + M * N + i);
#pragma unroll
      for (int j = 0; j < N; j++) s += ld(o + i * N + j) * ux[j];
      uU[i] = s;
    }
#pragma unroll
    for (int i = 0; i < N; i++) {
      T s = a.Nub != nullptr ? a.Nub[(b * Th + t) * N + i] : T(0);
#pragma unroll
      for (int j = 0; j < N; j++) s += A[i * N + j] * ux[j];
#pragma unroll
      for (int j = 0; j < M; j++) s += Bm[i * M + j] * uU[j];
      uxn[i] = s;
    }
#pragma unroll
    for (int i = 0; i < N; i++) {
      T s = ld(o + M * N + M + N * N + i);
#pragma unroll
      for (int j = 0; j < N; j++) s += ld(o + M * N + M + i * N + j) * uxn[j];
      unu[i] = s;
    }
    // leaf cotangents of step t (csrc/lqr_vjp.cuh), local only
    T gA[N * N], gQ[N * N], gB[N * M], gMm[N * M], gR[M * M];
#pragma unroll
    for (int i = 0; i < N; i++) {
#pragma unroll
      for (int j = 0; j < N; j++) {
        gA[i * N + j] = nu[i] * ux[j] + unu[i] * x[j];
        gQ[i * N + j] = T(0.5) * (ux[i] * x[j] + x[i] * ux[j]);
      }
#pragma unroll
      for (int j = 0; j < M; j++) {
        gB[i * M + j] = nu[i] * uU[j] + unu[i] * u[j];
        gMm[i * M + j] = ux[i] * u[j] + x[i] * uU[j];
      }
    }
#pragma unroll
    for (int i = 0; i < M; i++)
#pragma unroll
      for (int j = 0; j < M; j++) gR[i * M + j] = T(0.5) * (uU[i] * u[j] + u[i] * uU[j]);
    LeafCot<T> g;
    g.gQ = gQ; g.ldq = N; g.gq = ux; g.gR = gR; g.gr = uU; g.gM = gMm; g.gA = gA; g.lda = N; g.gB = gB; g.gd = unu;
    if (TD > 0) {
#pragma unroll
      for (int e = 0; e < TD; e++) acc[e] += P::theta_vjp_step(N, M, e, th, x, u, nu, g);
    } else {
      for (int e = 0; e < a.theta_dim; e++) gth[e] += P::theta_vjp_step(N, M, e, th, x, u, nu, g);
    }
    if (t == 0) {  // gx0 = M[0] uU[0] + A[0]^T uNu[0]
#pragma unroll
      for (int i = 0; i < N; i++) {
        T s = T(0);
#pragma unroll
        for (int j = 0; j < M; j++) s += Mx[i * M + j] * uU[j];
#pragma unroll
        for (int k = 0; k < N; k++) s += A[k * N + i] * unu[k];
        a.gx0[b * N + i] = s;
      }
    }
    if (t == Th - 1) {  // terminal leaves: gQf = sym(uX[T-1] (x) X[T-1]), gqf = uX[T-1]
      T xT[N], gQf[N * N];
#pragma unroll
      for (int i = 0; i < N; i++) xT[i] = Xg[(Th - 1) * N + i];
#pragma unroll
      for (int i = 0; i < N; i++)
#pragma unroll
        for (int j = 0; j < N; j++) gQf[i * N + j] = T(0.5) * (uxn[i] * xT[j] + uxn[j] * xT[i]);
      if (TD > 0) {
#pragma unroll
        for (int e = 0; e < TD; e++) acc[e] += P::theta_vjp_final(N, M, e, th, xT, gQf, N, uxn);
      } else {
        for (int e = 0; e < a.theta_dim; e++) gth[e] += P::theta_vjp_final(N, M, e, th, xT, gQf, N, uxn);
      }
    }
#pragma unroll
    for (int i = 0; i < N; i++) ux[i] = uxn[i];
  }
  if (TD > 0) {
#pragma unroll
    for (int e = 0; e < TD; e++) gth[e] = acc[e];
  }
}

}  // namespace idoc
