// Shared-control batch LQR (reference: idoc/blqr.py — `batch_size` systems driven by ONE control sequence U[T, m]).
//
// The reference carries a dense [bs, n, bs, n] value tensor through einsums (idoc/blqr.py:63-125): O(T bs^2 n^3) work and
// O(bs^2 n^2) memory.  That recursion is the LQR recursion of the block-diagonal lift (n' = bs n, A' = blockdiag(A_b),
// Q' = blockdiag(Q_b), B' = vstack(B_b), M' = vstack(M_b), shared R, r), and the lifted value matrix has the structure
//
//        V'_t  =  blockdiag(D_b)  -  Z Z^T ,      Z in R^{n' x r},   r = m (T - 1 - t)
//
// because each backward step  V <- Gxx' - Gxu' S Gux'  (S = Gt^{-1} (Gt + s I) Gt^{-1}, the eigen-shifted solve of
// idoc/blqr.py:95-105 with the unshifted Guu in the update, :109-113) subtracts one rank-m term and the step's
// A'^T (.) A' keeps block-diagonal parts block-diagonal and maps Z to A'^T Z.  Carrying (D_b, Z) instead of V':
//
//   Z^T B' = sum_b Z_b^T B_b  (r x m),  Z^T d' = sum_b Z_b^T d_b            reductions over the systems
//   W_b = D_b B_b,  Gxu_b = M_b + A_b^T W_b - (A_b^T Z_b)(Z^T B'),  Guu = R + sum_b B_b^T W_b - (Z^T B')^T (Z^T B')
//   w_b = v_b + D_b d_b - Z_b (Z^T d'),  gx_b = q_b + A_b^T w_b,  gu = r + sum_b B_b^T w_b
//   k = -Gt^{-1} gu,  K_b = -Gt^{-1} Gxu_b^T,  v_b+ = gx_b + Gxu_b (2 I - Gt^{-1} Guu) k,
//   D_b+ = sym(Q_b + A_b^T D_b A_b),  Z+ = [A'^T Z | Gxu' C],  C C^T = S
//
// costs O(bs n^2 (n + m + r)) per step and O(bs n (n + r)) memory: linear in the number of systems (the dense lift is cubic;
// the round-1 host lift was capped at bs n <= 128).  Rollout: u_t = k_t + sum_b K_{t,b} x_b (one reduction per step); co-states
// by the reference's own recursion per system (idoc/blqr.py:160-182).  One CTA per blqr problem; threads own systems.
// The implicit VJP (custom_root(kkt, solve_cg), idoc/blqr.py:258) is the same solver run on the adjoint problem
// (q' = Xb shifted, qf' = Xb[T-1], r' = Ub, d' = Nub, x0' = 0: SURVEY 3.4) followed by per-(t, b) outer products.
//
// Arrays follow the reference's shapes with an optional leading problem axis P:
//   Q[P,T,bs,n,n] q[P,T,bs,n] Qf[P,bs,n,n] qf[P,bs,n] M[P,T,bs,n,m] R[P,T,m,m] r[P,T,m] A[P,T,bs,n,n] B[P,T,bs,n,m] d[P,T,bs,n]
//   x0[P,bs,n];  X[P,T,bs,n] U[P,T,m] Nu[P,T,bs,n];  K[P,T,bs,m,n] k[P,T,m].
#pragma once
#include "lqr_fwd.cuh"  // IDOC_EPS, status bits

namespace idoc {

constexpr int BL_NMAX = 16, BL_MMAX = 8, BL_THREADS = 256;

template <typename T>
struct BlqrArgs {
  const T *Q, *q, *Qf, *qf, *Mx, *R, *r, *A, *B, *d, *x0;
  T *X, *U, *Nu;   // outputs
  T *K, *k;        // gains [P,T,bs,m,n], [P,T,m] (always written: the rollout reads them)
  T* dCdc;         // optional [P,2]
  int* status;     // optional [P]
  T* ws;           // per problem: D[bs,n,n] v[bs,n] gx[bs,n] Gxu[bs,n,m] Z[rmax,bs,n] ZtB[rmax,m] Ztd[rmax]
  long long ws_stride;  // elements per problem
  int P, T_, bs, n, m;
  int q_shift;     // adjoint mode: q[t] is read from q[t-1] (zero at t = 0) — q' = Xb shifted by one step
  int zero_x0;     // adjoint mode: x0' = 0
};

__host__ __device__ inline long long blqr_ws_elems(int T_, int bs, int n, int m) {
  const long long rmax = (long long)m * T_;
  return (long long)bs * n * n + 2LL * bs * n + (long long)bs * n * m + rmax * bs * n + rmax * m + rmax + 64;
}

// deterministic block sum of `cnt` (<= 80) per-thread values: warp shuffles, then warp partials added in warp order
template <typename T, int CNT>
__device__ inline void block_sum_vec(T (&v)[CNT], int cnt, T* red /* [warps][CNT] */, T* out /* [CNT] */) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
  for (int e = 0; e < cnt; e++) {
    T x = v[e];
    for (int o = 16; o; o >>= 1) x += __shfl_xor_sync(0xffffffffu, x, o);
    if (lane == 0) red[warp * CNT + e] = x;
  }
  __syncthreads();
  if (threadIdx.x < cnt) {
    T s = T(0);
    for (int w = 0; w < nw; w++) s += red[w * CNT + threadIdx.x];
    out[threadIdx.x] = s;
  }
  __syncthreads();
}

template <typename T>
__global__ void __launch_bounds__(BL_THREADS) blqr_riccati_kernel(const BlqrArgs<T> a) {
  constexpr int NM = BL_NMAX, MM = BL_MMAX, RC = MM + MM * MM;
  const int p = blockIdx.x, tid = threadIdx.x, nth = blockDim.x;
  const int Th = a.T_, bs = a.bs, n = a.n, m = a.m;
  const long long rmax = (long long)m * Th;
  T* ws = a.ws + (long long)p * a.ws_stride;
  T* D = ws;                                   // [bs][n][n]
  T* v = D + (long long)bs * n * n;            // [bs][n]
  T* gxs = v + (long long)bs * n;              // [bs][n]
  T* Gxu = gxs + (long long)bs * n;            // [bs][n][m]
  T* Z = Gxu + (long long)bs * n * m;          // [rmax][bs][n]
  T* ZtB = Z + rmax * bs * n;                  // [rmax][m]
  T* Ztd = ZtB + rmax * m;                     // [rmax]
  __shared__ T red[(BL_THREADS / 32) * RC];
  __shared__ T sums[RC];                       // [gu (m) | sum_b B^T D B (m*m)]
  __shared__ T sGinv[MM * MM], sC[MM * MM], sk[MM], skv[MM], sZZ[MM * MM];
  __shared__ T sShift;
  __shared__ int sStat;
  const T* Qp = a.Q + (long long)p * Th * bs * n * n;
  const T* qp = a.q ? a.q + (long long)p * Th * bs * n : nullptr;
  const T* Mp = a.Mx ? a.Mx + (long long)p * Th * bs * n * m : nullptr;
  const T* Rp = a.R + (long long)p * Th * m * m;
  const T* rp = a.r ? a.r + (long long)p * Th * m : nullptr;
  const T* Ap = a.A + (long long)p * Th * bs * n * n;
  const T* Bp = a.B + (long long)p * Th * bs * n * m;
  const T* dp = a.d ? a.d + (long long)p * Th * bs * n : nullptr;
  T* Kp = a.K + (long long)p * Th * bs * m * n;
  T* kp = a.k + (long long)p * Th * m;
  if (tid == 0) sStat = 0;
  // V_T = sym(Qf), v_T = qf
  for (int b = tid; b < bs; b += nth) {
    const T* Qf = a.Qf + ((long long)p * bs + b) * n * n;
    for (int i = 0; i < n; i++) {
      for (int j = 0; j < n; j++) D[((long long)b * n + i) * n + j] = T(0.5) * (Qf[i * n + j] + Qf[j * n + i]);
      v[(long long)b * n + i] = a.qf ? a.qf[((long long)p * bs + b) * n + i] : T(0);
    }
  }
  T dC = T(0), dc = T(0);
  int r = 0;  // current rank of Z
  __syncthreads();
  for (int t = Th - 1; t >= 0; t--) {
    const T* At = Ap + (long long)t * bs * n * n;
    const T* Bt = Bp + (long long)t * bs * n * m;
    const T* Qt = Qp + (long long)t * bs * n * n;
    const T* Mt = Mp ? Mp + (long long)t * bs * n * m : nullptr;
    const T* dt = dp ? dp + (long long)t * bs * n : nullptr;
    const T* qt = nullptr;
    if (qp) qt = a.q_shift ? (t > 0 ? qp + (long long)(t - 1) * bs * n : nullptr) : qp + (long long)t * bs * n;
    // ---- P1: Z^T B' and Z^T d' (one warp per column j of Z; lanes stride over the (b, i) pairs)
    {
      const int lane = tid & 31, warp = tid >> 5, nw = nth >> 5;
      for (int j = warp; j < r; j += nw) {
        const T* Zj = Z + (long long)j * bs * n;
        T acc[MM + 1];
        for (int c = 0; c <= m; c++) acc[c] = T(0);
        for (int e = lane; e < bs * n; e += 32) {
          const T z = Zj[e];
          for (int c = 0; c < m; c++) acc[c] += z * Bt[(long long)e * m + c];
          if (dt) acc[m] += z * dt[e];
        }
        for (int c = 0; c <= m; c++)
          for (int o = 16; o; o >>= 1) acc[c] += __shfl_xor_sync(0xffffffffu, acc[c], o);
        if (lane == 0) {
          for (int c = 0; c < m; c++) ZtB[(long long)j * m + c] = acc[c];
          Ztd[j] = acc[m];
        }
      }
    }
    __syncthreads();
    // ---- P2: per system: w, gx, W = D B, A^T W, and the partial sums of gu and B^T D B
    T part[RC];
    for (int e = 0; e < RC; e++) part[e] = T(0);
    for (int b = tid; b < bs; b += nth) {
      const T* Ab = At + (long long)b * n * n;
      const T* Bb = Bt + (long long)b * n * m;
      const T* Db = D + (long long)b * n * n;
      T w[NM], W[NM][MM];
      for (int i = 0; i < n; i++) {
        T acc = v[(long long)b * n + i];
        if (dt)
          for (int l = 0; l < n; l++) acc += Db[i * n + l] * dt[(long long)b * n + l];
        w[i] = acc;
      }
      if (dt)
        for (int j = 0; j < r; j++) {
          const T zd = Ztd[j];
          const T* Zjb = Z + ((long long)j * bs + b) * n;
          for (int i = 0; i < n; i++) w[i] -= Zjb[i] * zd;
        }
      for (int i = 0; i < n; i++) {  // gx = q + A^T w
        T acc = qt ? qt[(long long)b * n + i] : T(0);
        for (int l = 0; l < n; l++) acc += Ab[l * n + i] * w[l];
        gxs[(long long)b * n + i] = acc;
      }
      for (int c = 0; c < m; c++) {  // gu partial = B^T w
        T acc = T(0);
        for (int l = 0; l < n; l++) acc += Bb[l * m + c] * w[l];
        part[c] += acc;
      }
      for (int i = 0; i < n; i++)
        for (int c = 0; c < m; c++) {
          T acc = T(0);
          for (int l = 0; l < n; l++) acc += Db[i * n + l] * Bb[l * m + c];
          W[i][c] = acc;
        }
      for (int i = 0; i < n; i++)
        for (int c = 0; c < m; c++) {  // Gxu (block-diagonal part) = M + A^T D B
          T acc = Mt ? Mt[((long long)b * n + i) * m + c] : T(0);
          for (int l = 0; l < n; l++) acc += Ab[l * n + i] * W[l][c];
          Gxu[((long long)b * n + i) * m + c] = acc;
        }
      for (int c1 = 0; c1 < m; c1++)
        for (int c2 = 0; c2 < m; c2++) {
          T acc = T(0);
          for (int l = 0; l < n; l++) acc += Bb[l * m + c1] * W[l][c2];
          part[m + c1 * m + c2] += acc;
        }
    }
    block_sum_vec<T, RC>(part, m + m * m, red, sums);
    // ---- P3: Z <- A'^T Z (thread per (j, b))
    for (long long e = tid; e < (long long)r * bs; e += nth) {
      const int b = (int)(e % bs);
      T* Zjb = Z + e * n;
      const T* Ab = At + (long long)b * n * n;
      T z[NM], az[NM];
      for (int i = 0; i < n; i++) z[i] = Zjb[i];
      for (int i = 0; i < n; i++) {
        T acc = T(0);
        for (int l = 0; l < n; l++) acc += Ab[l * n + i] * z[l];
        az[i] = acc;
      }
      for (int i = 0; i < n; i++) Zjb[i] = az[i];
    }
    // (Z^T B')^T (Z^T B'): m x m, warp 0 strides over j
    if (tid < 32) {
      T zz[MM * MM];
      for (int e = 0; e < m * m; e++) zz[e] = T(0);
      for (int j = tid; j < r; j += 32)
        for (int c1 = 0; c1 < m; c1++)
          for (int c2 = 0; c2 < m; c2++) zz[c1 * m + c2] += ZtB[(long long)j * m + c1] * ZtB[(long long)j * m + c2];
      for (int e = 0; e < m * m; e++) {
        for (int o = 16; o; o >>= 1) zz[e] += __shfl_xor_sync(0xffffffffu, zz[e], o);
        if (tid == 0) sZZ[e] = zz[e];
      }
    }
    __syncthreads();
    // ---- P4: Gxu_b -= (A^T Z)_b (Z^T B')
    for (int b = tid; b < bs; b += nth) {
      T g[NM][MM];
      for (int i = 0; i < n; i++)
        for (int c = 0; c < m; c++) g[i][c] = Gxu[((long long)b * n + i) * m + c];
      for (int j = 0; j < r; j++) {
        const T* Zjb = Z + ((long long)j * bs + b) * n;
        for (int i = 0; i < n; i++) {
          const T z = Zjb[i];
          for (int c = 0; c < m; c++) g[i][c] -= z * ZtB[(long long)j * m + c];
        }
      }
      for (int i = 0; i < n; i++)
        for (int c = 0; c < m; c++) Gxu[((long long)b * n + i) * m + c] = g[i][c];
    }
    // ---- P5: the m x m solve (thread 0): Guu, eigen-shift, Cholesky, k, and the factor C of S
    if (tid == 0) {
      T Guu[MM][MM], Gt[MM][MM], Lf[MM][MM], Li[MM][MM], gu[MM];
      const T* Rt = Rp + (long long)t * m * m;
      for (int c1 = 0; c1 < m; c1++) {
        gu[c1] = (rp ? rp[(long long)t * m + c1] : T(0)) + sums[c1];
        for (int c2 = 0; c2 < m; c2++)
          Guu[c1][c2] = T(0.5) * (Rt[c1 * m + c2] + Rt[c2 * m + c1]) + T(0.5) * (sums[m + c1 * m + c2] + sums[m + c2 * m + c1]) - sZZ[c1 * m + c2];
      }
      T shift = T(0);
      for (int attempt = 0; attempt < 2; attempt++) {
        bool ok = true;
        for (int i = 0; i < m; i++)
          for (int j = 0; j < m; j++) Gt[i][j] = Guu[i][j] + (i == j ? shift : T(0));
        for (int j = 0; j < m; j++) {  // Cholesky (lower)
          T s = Gt[j][j];
          for (int c = 0; c < j; c++) s -= Lf[j][c] * Lf[j][c];
          if (!(s > T(0))) {
            ok = false;
            if (attempt == 1) { s = T(IDOC_EPS) * T(1e-3); sStat |= ST_CHOL_CLAMP; } else s = T(1);
          }
          const T dj = sqrt(s);
          Lf[j][j] = dj;
          for (int i = j + 1; i < m; i++) {
            T x = Gt[i][j];
            for (int c = 0; c < j; c++) x -= Lf[i][c] * Lf[j][c];
            Lf[i][j] = x / dj;
          }
        }
        for (int c = 0; c < m; c++) {  // inverse factor, column c
          for (int i = 0; i < c; i++) Li[i][c] = T(0);
          Li[c][c] = T(1) / Lf[c][c];
          for (int i = c + 1; i < m; i++) {
            T acc = T(0);
            for (int l = c; l < i; l++) acc += Lf[i][l] * Li[l][c];
            Li[i][c] = -acc / Lf[i][i];
          }
        }
        T fro = T(0);
        for (int i = 0; i < m; i++)
          for (int j = 0; j <= i; j++) fro += Li[i][j] * Li[i][j];
        if (attempt == 1) break;
        if (ok && fro < T(1.0 / IDOC_EPS)) break;  // lambda_min >= 1 / |Linv|_F^2 > EPS: no shift (idoc/blqr.py:95-100)
        // lambda_min by Jacobi sweeps
        T Am[MM][MM];
        for (int i = 0; i < m; i++)
          for (int j = 0; j < m; j++) Am[i][j] = Guu[i][j];
        for (int sweep = 0; sweep < 30 && m > 1; sweep++) {
          T off = T(0), dg = T(0);
          for (int i = 0; i < m; i++) { dg += Am[i][i] * Am[i][i]; for (int j = i + 1; j < m; j++) off += Am[i][j] * Am[i][j]; }
          if (!(off > (sizeof(T) == 8 ? T(1e-30) : T(1e-14)) * (dg + off))) break;
          for (int pi = 0; pi < m - 1; pi++)
            for (int qi = pi + 1; qi < m; qi++) {
              const T apq = Am[pi][qi];
              if (apq == T(0)) continue;
              const T theta = (Am[qi][qi] - Am[pi][pi]) / (T(2) * apq);
              const T tt = (theta >= T(0) ? T(1) : T(-1)) / (abs_(theta) + sqrt(theta * theta + T(1)));
              const T cc = T(1) / sqrt(tt * tt + T(1)), sn = tt * cc;
              for (int l = 0; l < m; l++) { const T x1 = Am[l][pi], x2 = Am[l][qi]; Am[l][pi] = cc * x1 - sn * x2; Am[l][qi] = sn * x1 + cc * x2; }
              for (int l = 0; l < m; l++) { const T x1 = Am[pi][l], x2 = Am[qi][l]; Am[pi][l] = cc * x1 - sn * x2; Am[qi][l] = sn * x1 + cc * x2; }
            }
        }
        T mn = Am[0][0];
        for (int i = 1; i < m; i++) mn = Am[i][i] < mn ? Am[i][i] : mn;
        if (!(mn == mn)) sStat |= ST_NONFINITE;
        shift = T(IDOC_EPS) - mn;
        shift = shift > T(0) ? shift : T(0);
        if (shift > T(0)) sStat |= ST_SHIFT;
      }
      // Ginv = Li^T Li
      for (int i = 0; i < m; i++)
        for (int j = 0; j < m; j++) {
          T acc = T(0);
          for (int l = (i > j ? i : j); l < m; l++) acc += Li[l][i] * Li[l][j];
          sGinv[i * m + j] = acc;
        }
      T kk[MM], Gk[MM];
      for (int i = 0; i < m; i++) {
        T acc = T(0);
        for (int j = 0; j < m; j++) acc += sGinv[i * m + j] * gu[j];
        kk[i] = -acc;
        sk[i] = -acc;
      }
      T quad = T(0), lin = T(0);
      for (int i = 0; i < m; i++) {
        T acc = T(0);
        for (int j = 0; j < m; j++) acc += Guu[i][j] * kk[j];
        Gk[i] = acc;
        quad += acc * kk[i];
        lin += gu[i] * kk[i];
      }
      dC += T(0.5) * quad;  // idoc/blqr.py:116-117 (unshifted Guu)
      dc += lin;
      for (int i = 0; i < m; i++) {  // kv = (2 I - Ginv Guu) k
        T acc = T(0);
        for (int j = 0; j < m; j++) acc += sGinv[i * m + j] * Gk[j];
        skv[i] = T(2) * kk[i] - acc;
      }
      // S = Ginv (Gt + s I) Ginv = Ginv + s Ginv^2 = Li^T (I + s Li Li^T) Li;  C = Li^T chol(I + s Li Li^T)
      T Cm[MM][MM];
      if (shift > T(0)) {
        T Hm[MM][MM], Lh[MM][MM];
        for (int i = 0; i < m; i++)
          for (int j = 0; j < m; j++) {
            T acc = T(0);
            for (int l = 0; l <= (i < j ? i : j); l++) acc += Li[i][l] * Li[j][l];
            Hm[i][j] = (i == j ? T(1) : T(0)) + shift * acc;
          }
        for (int j = 0; j < m; j++) {
          T s = Hm[j][j];
          for (int c = 0; c < j; c++) s -= Lh[j][c] * Lh[j][c];
          const T dj = sqrt(s > T(0) ? s : T(1));
          Lh[j][j] = dj;
          for (int i = j + 1; i < m; i++) {
            T x = Hm[i][j];
            for (int c = 0; c < j; c++) x -= Lh[i][c] * Lh[j][c];
            Lh[i][j] = x / dj;
          }
          for (int i = 0; i < j; i++) Lh[i][j] = T(0);
        }
        for (int i = 0; i < m; i++)
          for (int j = 0; j < m; j++) {  // C = Li^T Lh
            T acc = T(0);
            for (int l = (i > j ? i : j); l < m; l++) acc += Li[l][i] * Lh[l][j];
            Cm[i][j] = acc;
          }
      } else {
        for (int i = 0; i < m; i++)
          for (int j = 0; j < m; j++) Cm[i][j] = Li[j][i];  // C = Li^T
      }
      for (int i = 0; i < m; i++)
        for (int j = 0; j < m; j++) sC[i * m + j] = Cm[i][j];
      sShift = shift;
      for (int c = 0; c < m; c++) kp[(long long)t * m + c] = kk[c];
    }
    __syncthreads();
    // ---- P6: per system: K_b, v_b+, D_b+, the new columns of Z
    for (int b = tid; b < bs; b += nth) {
      const T* Ab = At + (long long)b * n * n;
      const T* Qb = Qt + (long long)b * n * n;
      T* Db = D + (long long)b * n * n;
      T g[NM][MM];
      for (int i = 0; i < n; i++)
        for (int c = 0; c < m; c++) g[i][c] = Gxu[((long long)b * n + i) * m + c];
      T* Kb = Kp + ((long long)t * bs + b) * m * n;
      for (int c = 0; c < m; c++)
        for (int i = 0; i < n; i++) {  // K = -Ginv Gxu^T
          T acc = T(0);
          for (int l = 0; l < m; l++) acc += sGinv[c * m + l] * g[i][l];
          Kb[c * n + i] = -acc;
        }
      for (int i = 0; i < n; i++) {
        T acc = gxs[(long long)b * n + i];
        for (int c = 0; c < m; c++) acc += g[i][c] * skv[c];
        v[(long long)b * n + i] = acc;
      }
      T AD[NM][NM];  // A^T D
      for (int i = 0; i < n; i++)
        for (int j = 0; j < n; j++) {
          T acc = T(0);
          for (int l = 0; l < n; l++) acc += Ab[l * n + i] * Db[l * n + j];
          AD[i][j] = acc;
        }
      T Dn[NM][NM];
      for (int i = 0; i < n; i++)
        for (int j = 0; j < n; j++) {
          T acc = T(0.5) * (Qb[i * n + j] + Qb[j * n + i]);
          for (int l = 0; l < n; l++) acc += AD[i][l] * Ab[l * n + j];
          Dn[i][j] = acc;
        }
      for (int i = 0; i < n; i++)
        for (int j = 0; j < n; j++) Db[i * n + j] = T(0.5) * (Dn[i][j] + Dn[j][i]);
      for (int c = 0; c < m; c++) {  // Z[r + c][b][:] = Gxu_b C[:, c]
        T* Zn = Z + ((long long)(r + c) * bs + b) * n;
        for (int i = 0; i < n; i++) {
          T acc = T(0);
          for (int l = 0; l < m; l++) acc += g[i][l] * sC[l * m + c];
          Zn[i] = acc;
        }
      }
    }
    r += m;
    __syncthreads();
  }
  if (tid == 0) {
    if (a.dCdc) { a.dCdc[p * 2] = dC; a.dCdc[p * 2 + 1] = dc; }
    if (a.status) a.status[p] = sStat;
  }
}

// forward rollout (idoc/blqr.py:229-242) + co-states (idoc/blqr.py:160-182)
template <typename T>
__global__ void __launch_bounds__(BL_THREADS) blqr_rollout_kernel(const BlqrArgs<T> a) {
  constexpr int NM = BL_NMAX, MM = BL_MMAX;
  const int p = blockIdx.x, tid = threadIdx.x, nth = blockDim.x;
  const int Th = a.T_, bs = a.bs, n = a.n, m = a.m;
  __shared__ T red[(BL_THREADS / 32) * MM];
  __shared__ T su[MM];
  const T* Ap = a.A + (long long)p * Th * bs * n * n;
  const T* Bp = a.B + (long long)p * Th * bs * n * m;
  const T* dp = a.d ? a.d + (long long)p * Th * bs * n : nullptr;
  const T* Kp = a.K + (long long)p * Th * bs * m * n;
  const T* kp = a.k + (long long)p * Th * m;
  T* Xp = a.X + (long long)p * Th * bs * n;
  T* Up = a.U + (long long)p * Th * m;
  T* Nup = a.Nu + (long long)p * Th * bs * n;
  for (int t = 0; t < Th; t++) {
    T part[MM];
    for (int c = 0; c < m; c++) part[c] = T(0);
    for (int b = tid; b < bs; b += nth) {  // sum_b K_b x_b
      const T* xb = t > 0 ? Xp + ((long long)(t - 1) * bs + b) * n : (a.zero_x0 ? nullptr : a.x0 + ((long long)p * bs + b) * n);
      const T* Kb = Kp + ((long long)t * bs + b) * m * n;
      if (xb)
        for (int c = 0; c < m; c++) {
          T acc = T(0);
          for (int i = 0; i < n; i++) acc += Kb[c * n + i] * xb[i];
          part[c] += acc;
        }
    }
    block_sum_vec<T, MM>(part, m, red, su);
    if (tid < m) {
      su[tid] += kp[(long long)t * m + tid];
      Up[(long long)t * m + tid] = su[tid];
    }
    __syncthreads();
    for (int b = tid; b < bs; b += nth) {
      const T* xb = t > 0 ? Xp + ((long long)(t - 1) * bs + b) * n : (a.zero_x0 ? nullptr : a.x0 + ((long long)p * bs + b) * n);
      const T* Ab = Ap + ((long long)t * bs + b) * n * n;
      const T* Bb = Bp + ((long long)t * bs + b) * n * m;
      T x[NM];
      for (int i = 0; i < n; i++) x[i] = xb ? xb[i] : T(0);
      for (int i = 0; i < n; i++) {
        T acc = dp ? dp[((long long)t * bs + b) * n + i] : T(0);
        for (int l = 0; l < n; l++) acc += Ab[i * n + l] * x[l];
        for (int c = 0; c < m; c++) acc += Bb[i * m + c] * su[c];
        Xp[((long long)t * bs + b) * n + i] = acc;
      }
    }
    __syncthreads();
  }
  // co-states per system: Nu[T-1] = Qf X[T-1] + qf;  Nu[t-1] = A[t]^T Nu[t] + Q[t] X[t-1] + q[t] + M[t] U[t]
  const T* Qp = a.Q + (long long)p * Th * bs * n * n;
  const T* qp = a.q ? a.q + (long long)p * Th * bs * n : nullptr;
  const T* Mp = a.Mx ? a.Mx + (long long)p * Th * bs * n * m : nullptr;
  for (int b = tid; b < bs; b += nth) {
    const T* Qf = a.Qf + ((long long)p * bs + b) * n * n;
    T nu[NM], nn[NM];
    for (int i = 0; i < n; i++) {
      T acc = a.qf ? a.qf[((long long)p * bs + b) * n + i] : T(0);
      for (int j = 0; j < n; j++) acc += T(0.5) * (Qf[i * n + j] + Qf[j * n + i]) * Xp[((long long)(Th - 1) * bs + b) * n + j];
      nu[i] = acc;
      Nup[((long long)(Th - 1) * bs + b) * n + i] = acc;
    }
    for (int t = Th - 1; t >= 1; t--) {
      const T* Ab = Ap + ((long long)t * bs + b) * n * n;
      const T* Qb = Qp + ((long long)t * bs + b) * n * n;
      const T* xb = Xp + ((long long)(t - 1) * bs + b) * n;
      const T* qt = nullptr;
      if (qp) qt = a.q_shift ? qp + ((long long)(t - 1) * bs + b) * n : qp + ((long long)t * bs + b) * n;
      for (int i = 0; i < n; i++) {
        T acc = qt ? qt[i] : T(0);
        for (int j = 0; j < n; j++) acc += Ab[j * n + i] * nu[j] + T(0.5) * (Qb[i * n + j] + Qb[j * n + i]) * xb[j];
        if (Mp)
          for (int c = 0; c < m; c++) acc += Mp[(((long long)t * bs + b) * n + i) * m + c] * Up[(long long)t * m + c];
        nn[i] = acc;
      }
      for (int i = 0; i < n; i++) {
        nu[i] = nn[i];
        Nup[((long long)(t - 1) * bs + b) * n + i] = nn[i];
      }
    }
  }
}

// parameter cotangents from the forward solution (X, U, Nu) and the adjoint solution (uX, uU, uNu): SURVEY 3.4 per system
template <typename T>
struct BlqrGradArgs {
  const T *Mx, *A, *x0, *X, *U, *Nu, *uX, *uU, *uNu;
  T *gQ, *gq, *gQf, *gqf, *gM, *gR, *gr, *gA, *gB, *gd, *gx0;
  int P, T_, bs, n, m;
};
template <typename T>
__global__ void blqr_grad_kernel(const BlqrGradArgs<T> a) {
  const int Th = a.T_, bs = a.bs, n = a.n, m = a.m;
  const long long tot = (long long)a.P * Th * bs;
  const long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= tot) return;
  const int b = (int)(e % bs), t = (int)((e / bs) % Th);
  const long long p = e / ((long long)bs * Th);
  const long long tb = (p * Th + t) * bs + b;
  const T* nu = a.Nu + tb * n;
  const T* unu = a.uNu + tb * n;
  const T* sx = t > 0 ? a.X + (tb - bs) * n : a.x0 + (p * bs + b) * n;
  const T* sux = t > 0 ? a.uX + (tb - bs) * n : nullptr;
  const T* U = a.U + (p * Th + t) * m;
  const T* uU = a.uU + (p * Th + t) * m;
  for (int i = 0; i < n; i++) {
    const T suxi = sux ? sux[i] : T(0);
    for (int j = 0; j < n; j++) {
      const T suxj = sux ? sux[j] : T(0);
      a.gA[(tb * n + i) * n + j] = nu[i] * suxj + unu[i] * sx[j];
      a.gQ[(tb * n + i) * n + j] = T(0.5) * (suxi * sx[j] + sx[i] * suxj);
    }
    for (int c = 0; c < m; c++) {
      a.gB[(tb * n + i) * m + c] = nu[i] * uU[c] + unu[i] * U[c];
      a.gM[(tb * n + i) * m + c] = suxi * U[c] + sx[i] * uU[c];
    }
    a.gd[tb * n + i] = unu[i];
    a.gq[tb * n + i] = suxi;
  }
  if (b == 0) {
    for (int c1 = 0; c1 < m; c1++) {
      for (int c2 = 0; c2 < m; c2++) a.gR[((p * Th + t) * m + c1) * m + c2] = T(0.5) * (uU[c1] * U[c2] + U[c1] * uU[c2]);
      a.gr[(p * Th + t) * m + c1] = uU[c1];
    }
  }
  if (t == 0) {  // gx0 = M[0] uU[0] + A[0]^T uNu[0]
    const T* A0 = a.A + tb * n * n;
    for (int i = 0; i < n; i++) {
      T acc = T(0);
      if (a.Mx)
        for (int c = 0; c < m; c++) acc += a.Mx[(tb * n + i) * m + c] * uU[c];
      for (int l = 0; l < n; l++) acc += A0[l * n + i] * unu[l];
      a.gx0[(p * bs + b) * n + i] = acc;
    }
  }
  if (t == Th - 1) {  // gQf = sym(uX[T-1] (x) X[T-1]), gqf = uX[T-1]
    const T* xT = a.X + tb * n;
    const T* uxT = a.uX + tb * n;
    for (int i = 0; i < n; i++) {
      for (int j = 0; j < n; j++) a.gQf[((p * bs + b) * n + i) * n + j] = T(0.5) * (uxT[i] * xT[j] + xT[i] * uxT[j]);
      a.gqf[(p * bs + b) * n + i] = uxT[i];
    }
  }
}

}  // namespace idoc
