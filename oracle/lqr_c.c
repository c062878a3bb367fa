/* CPU ORACLE (test / baseline infrastructure, not product code) — plain C restatement of the reference's
 * time-varying LQR, /root/reference/idoc/lqr.py, one problem per OpenMP task.  Used only by bench.py's
 * cpu_baseline / `--impl reference` leg and by tests/ (checked there against oracle/lqr_np.py and the golden
 * vectors produced from the reference sources).  Compiled twice: REAL=double (symbols *_f64) and float (*_f32).
 *
 *   backward : lqr.py:62-107  (step body :77-95: eigh shift, general LU solve, unshifted-Guu value update)
 *   forward  : lqr.py:153-164
 *   adjoint  : lqr.py:110-125
 *   direct   : lqr.py:170-175
 *   vjp      : the exact implicit VJP that jaxopt.custom_root(kkt) (lqr.py:177) defines, computed as one
 *              adjoint LQR solve + outer products — same algorithm as oracle/lqr_np.py:vjp_exact.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#ifndef REAL
#define REAL double
#endif
#ifndef SUFFIX
#define SUFFIX f64
#endif
#define CAT_(a, b) a##_##b
#define CAT(a, b) CAT_(a, b)
#define FN(name) CAT(name, SUFFIX)

#define MAXD 32
#define EPS ((REAL)1e-8) /* lqr.py:73 */

typedef struct {
  const REAL *Q, *q, *Qf, *qf, *M, *R, *r, *A, *B, *d;
} lqr_t; /* one problem: Q[T,n,n] q[T,n] Qf[n,n] qf[n] M[T,n,m] R[T,m,m] r[T,m] A[T,n,n] B[T,n,m] d[T,n] */

/* smallest eigenvalue of a symmetric m x m matrix (cyclic Jacobi) — jnp.linalg.eigh(Guu)[0][0], lqr.py:86-87 */
static REAL min_eig(const REAL* G, int m) {
  REAL A[MAXD * MAXD];
  memcpy(A, G, sizeof(REAL) * m * m);
  for (int sweep = 0; sweep < 50; sweep++) {
    REAL off = 0, diag = 0;
    for (int i = 0; i < m; i++) {
      diag += A[i * m + i] * A[i * m + i];
      for (int j = i + 1; j < m; j++) off += A[i * m + j] * A[i * m + j];
    }
    if (!(off > (REAL)1e-32 * (diag + off))) break;
    for (int p = 0; p < m - 1; p++)
      for (int q = p + 1; q < m; q++) {
        REAL apq = A[p * m + q];
        if (apq == 0) continue;
        REAL theta = (A[q * m + q] - A[p * m + p]) / (2 * apq);
        REAL t = (theta >= 0 ? 1 : -1) / (fabs(theta) + sqrt(theta * theta + 1));
        REAL c = 1 / sqrt(t * t + 1), s = t * c;
        for (int k = 0; k < m; k++) {
          REAL akp = A[k * m + p], akq = A[k * m + q];
          A[k * m + p] = c * akp - s * akq;
          A[k * m + q] = s * akp + c * akq;
        }
        for (int k = 0; k < m; k++) {
          REAL apk = A[p * m + k], aqk = A[q * m + k];
          A[p * m + k] = c * apk - s * aqk;
          A[q * m + k] = s * apk + c * aqk;
        }
      }
  }
  REAL mn = A[0];
  for (int i = 1; i < m; i++)
    if (A[i * m + i] < mn) mn = A[i * m + i];
  return mn;
}

/* general solve  G X = Bm  (m x m, nrhs columns), LU with partial pivoting — jax.scipy.linalg.solve, lqr.py:89-90 */
static void lu_solve(const REAL* G, int m, REAL* Bm, int nrhs) {
  REAL A[MAXD * MAXD];
  memcpy(A, G, sizeof(REAL) * m * m);
  for (int c = 0; c < m; c++) {
    int piv = c;
    for (int i = c + 1; i < m; i++)
      if (fabs(A[i * m + c]) > fabs(A[piv * m + c])) piv = i;
    if (piv != c) {
      for (int j = 0; j < m; j++) { REAL t = A[c * m + j]; A[c * m + j] = A[piv * m + j]; A[piv * m + j] = t; }
      for (int j = 0; j < nrhs; j++) { REAL t = Bm[c * nrhs + j]; Bm[c * nrhs + j] = Bm[piv * nrhs + j]; Bm[piv * nrhs + j] = t; }
    }
    REAL inv = 1 / A[c * m + c];
    for (int i = c + 1; i < m; i++) {
      REAL f = A[i * m + c] * inv;
      if (f == 0) continue;
      for (int j = c + 1; j < m; j++) A[i * m + j] -= f * A[c * m + j];
      for (int j = 0; j < nrhs; j++) Bm[i * nrhs + j] -= f * Bm[c * nrhs + j];
    }
  }
  for (int c = m - 1; c >= 0; c--) {
    REAL inv = 1 / A[c * m + c];
    for (int j = 0; j < nrhs; j++) {
      REAL s = Bm[c * nrhs + j];
      for (int k = c + 1; k < m; k++) s -= A[c * m + k] * Bm[k * nrhs + j];
      Bm[c * nrhs + j] = s * inv;
    }
  }
}

/* lqr.py:62-107.  K[T,m,n], k[T,m]; dCdc[2] optional */
static void backward(const lqr_t* p, int T, int n, int m, REAL* K, REAL* k, REAL* dCdc) {
  REAL V[MAXD * MAXD], v[MAXD], VA[MAXD * MAXD], VB[MAXD * MAXD], Gxx[MAXD * MAXD], Guu[MAXD * MAXD], Gxu[MAXD * MAXD];
  REAL Gt[MAXD * MAXD], rhs[MAXD * (MAXD + 1)], Vd[MAXD], gx[MAXD], gu[MAXD], tmp[MAXD * MAXD], Gk[MAXD], w[MAXD];
  memcpy(V, p->Qf, sizeof(REAL) * n * n);
  memcpy(v, p->qf, sizeof(REAL) * n);
  REAL dC = 0, dc = 0;
  for (int t = T - 1; t >= 0; t--) {
    const REAL *A = p->A + (size_t)t * n * n, *B = p->B + (size_t)t * n * m, *Q = p->Q + (size_t)t * n * n;
    const REAL *M = p->M + (size_t)t * n * m, *R = p->R + (size_t)t * m * m;
    const REAL *q = p->q + (size_t)t * n, *r = p->r + (size_t)t * m, *d = p->d + (size_t)t * n;
    for (int i = 0; i < n; i++) {
      for (int j = 0; j < n; j++) { REAL s = 0; for (int l = 0; l < n; l++) s += V[i * n + l] * A[l * n + j]; VA[i * n + j] = s; }
      for (int j = 0; j < m; j++) { REAL s = 0; for (int l = 0; l < n; l++) s += V[i * n + l] * B[l * m + j]; VB[i * m + j] = s; }
      REAL s = 0; for (int l = 0; l < n; l++) s += V[i * n + l] * d[l]; Vd[i] = s;           /* lqr.py:83 */
    }
    for (int i = 0; i < n; i++)
      for (int j = 0; j < n; j++) { REAL s = Q[i * n + j]; for (int l = 0; l < n; l++) s += A[l * n + i] * VA[l * n + j]; tmp[i * n + j] = s; }
    for (int i = 0; i < n; i++) for (int j = 0; j < n; j++) Gxx[i * n + j] = (REAL)0.5 * (tmp[i * n + j] + tmp[j * n + i]); /* :80 */
    for (int i = 0; i < m; i++)
      for (int j = 0; j < m; j++) { REAL s = R[i * m + j]; for (int l = 0; l < n; l++) s += B[l * m + i] * VB[l * m + j]; tmp[i * m + j] = s; }
    for (int i = 0; i < m; i++) for (int j = 0; j < m; j++) Guu[i * m + j] = (REAL)0.5 * (tmp[i * m + j] + tmp[j * m + i]); /* :81 */
    for (int i = 0; i < n; i++)
      for (int j = 0; j < m; j++) { REAL s = M[i * m + j]; for (int l = 0; l < n; l++) s += A[l * n + i] * VB[l * m + j]; Gxu[i * m + j] = s; } /* :82 */
    for (int i = 0; i < n; i++) w[i] = v[i] + Vd[i];
    for (int i = 0; i < n; i++) { REAL s = q[i]; for (int l = 0; l < n; l++) s += A[l * n + i] * w[l]; gx[i] = s; }  /* :84 */
    for (int i = 0; i < m; i++) { REAL s = r[i]; for (int l = 0; l < n; l++) s += B[l * m + i] * w[l]; gu[i] = s; }  /* :85 */
    REAL mn = min_eig(Guu, m);                                                                 /* :86-87 */
    REAL shift = EPS - mn; if (shift < 0) shift = 0;
    memcpy(Gt, Guu, sizeof(REAL) * m * m);
    for (int i = 0; i < m; i++) Gt[i * m + i] += shift;                                        /* :88 */
    for (int i = 0; i < m; i++) { for (int j = 0; j < n; j++) rhs[i * (n + 1) + j] = Gxu[j * m + i]; rhs[i * (n + 1) + n] = gu[i]; }
    lu_solve(Gt, m, rhs, n + 1);
    REAL* Kt = K + (size_t)t * m * n; REAL* kt = k + (size_t)t * m;
    for (int i = 0; i < m; i++) { for (int j = 0; j < n; j++) Kt[i * n + j] = -rhs[i * (n + 1) + j]; kt[i] = -rhs[i * (n + 1) + n]; } /* :89-90 */
    /* V = sym(Gxx + Gxu K + K^T Gxu^T + K^T Guu K)   :91 */
    for (int i = 0; i < m; i++) for (int j = 0; j < n; j++) { REAL s = 0; for (int l = 0; l < m; l++) s += Guu[i * m + l] * Kt[l * n + j]; VB[i * n + j] = s; } /* Guu K */
    for (int i = 0; i < n; i++)
      for (int j = 0; j < n; j++) {
        REAL s = Gxx[i * n + j];
        for (int l = 0; l < m; l++) s += Gxu[i * m + l] * Kt[l * n + j] + Kt[l * n + i] * Gxu[j * m + l] + Kt[l * n + i] * VB[l * n + j];
        tmp[i * n + j] = s;
      }
    for (int i = 0; i < n; i++) for (int j = 0; j < n; j++) V[i * n + j] = (REAL)0.5 * (tmp[i * n + j] + tmp[j * n + i]);
    for (int i = 0; i < m; i++) { REAL s = 0; for (int l = 0; l < m; l++) s += Guu[i * m + l] * kt[l]; Gk[i] = s; }
    for (int i = 0; i < n; i++) {                                                               /* :92 */
      REAL s = gx[i];
      for (int l = 0; l < m; l++) s += Gxu[i * m + l] * kt[l] + Kt[l * n + i] * gu[l] + Kt[l * n + i] * Gk[l];
      v[i] = s;
    }
    for (int i = 0; i < m; i++) { dC += (REAL)0.5 * Gk[i] * kt[i]; dc += gu[i] * kt[i]; }       /* :93-94 */
  }
  if (dCdc) { dCdc[0] = dC; dCdc[1] = dc; }
}

/* lqr.py:153-164 */
static void forward(const lqr_t* p, const REAL* x0, const REAL* K, const REAL* k, int T, int n, int m, REAL* X, REAL* U) {
  REAL x[MAXD], xn[MAXD];
  memcpy(x, x0, sizeof(REAL) * n);
  for (int t = 0; t < T; t++) {
    const REAL *A = p->A + (size_t)t * n * n, *B = p->B + (size_t)t * n * m, *d = p->d + (size_t)t * n;
    REAL* u = U + (size_t)t * m;
    for (int i = 0; i < m; i++) { REAL s = k[(size_t)t * m + i]; for (int j = 0; j < n; j++) s += K[((size_t)t * m + i) * n + j] * x[j]; u[i] = s; }
    for (int i = 0; i < n; i++) {
      REAL s = d[i];
      for (int j = 0; j < n; j++) s += A[i * n + j] * x[j];
      for (int j = 0; j < m; j++) s += B[i * m + j] * u[j];
      xn[i] = s;
    }
    memcpy(x, xn, sizeof(REAL) * n);
    memcpy(X + (size_t)t * n, xn, sizeof(REAL) * n);
  }
}

/* lqr.py:110-125 */
static void adjoint(const lqr_t* p, const REAL* X, const REAL* U, int T, int n, int m, REAL* Nu) {
  REAL* nu = Nu + (size_t)(T - 1) * n;
  for (int i = 0; i < n; i++) { REAL s = p->qf[i]; for (int j = 0; j < n; j++) s += p->Qf[i * n + j] * X[(size_t)(T - 1) * n + j]; nu[i] = s; }
  for (int t = T - 1; t >= 1; t--) {
    const REAL *A = p->A + (size_t)t * n * n, *Q = p->Q + (size_t)t * n * n, *M = p->M + (size_t)t * n * m, *q = p->q + (size_t)t * n;
    const REAL* nut = Nu + (size_t)t * n;
    REAL* out = Nu + (size_t)(t - 1) * n;
    for (int i = 0; i < n; i++) {
      REAL s = q[i];
      for (int j = 0; j < n; j++) s += A[j * n + i] * nut[j] + Q[i * n + j] * X[(size_t)(t - 1) * n + j];
      for (int j = 0; j < m; j++) s += M[i * m + j] * U[(size_t)t * m + j];
      out[i] = s;
    }
  }
}

static lqr_t slice(int64_t b, int T, int n, int m, const REAL* Q, const REAL* q, const REAL* Qf, const REAL* qf, const REAL* M,
                   const REAL* R, const REAL* r, const REAL* A, const REAL* B, const REAL* d) {
  lqr_t p;
  p.Q = Q + (size_t)b * T * n * n; p.q = q + (size_t)b * T * n; p.Qf = Qf + (size_t)b * n * n; p.qf = qf + (size_t)b * n;
  p.M = M + (size_t)b * T * n * m; p.R = R + (size_t)b * T * m * m; p.r = r + (size_t)b * T * m;
  p.A = A + (size_t)b * T * n * n; p.B = B + (size_t)b * T * n * m; p.d = d + (size_t)b * T * n;
  return p;
}

int FN(oracle_threads)(void) {
#ifdef _OPENMP
  return omp_get_max_threads();
#else
  return 1;
#endif
}

/* thread count of the following calls (launchers such as torchrun export OMP_NUM_THREADS=1) */
void FN(oracle_set_threads)(int n) {
#ifdef _OPENMP
  if (n > 0) omp_set_num_threads(n);
#else
  (void)n;
#endif
}

/* direct (lqr.py:170-175) over a batch; K,k scratch [B,T,m,n],[B,T,m] are outputs too */
void FN(oracle_lqr_fwd)(int64_t Bn, int T, int n, int m, const REAL* Q, const REAL* q, const REAL* Qf, const REAL* qf,
                        const REAL* M, const REAL* R, const REAL* r, const REAL* A, const REAL* B, const REAL* d,
                        const REAL* x0, REAL* X, REAL* U, REAL* Nu, REAL* K, REAL* k) {
#pragma omp parallel for schedule(static)
  for (int64_t b = 0; b < Bn; b++) {
    lqr_t p = slice(b, T, n, m, Q, q, Qf, qf, M, R, r, A, B, d);
    REAL* Kb = K + (size_t)b * T * m * n; REAL* kb = k + (size_t)b * T * m;
    backward(&p, T, n, m, Kb, kb, 0);
    forward(&p, x0 + (size_t)b * n, Kb, kb, T, n, m, X + (size_t)b * T * n, U + (size_t)b * T * m);
    adjoint(&p, X + (size_t)b * T * n, U + (size_t)b * T * m, T, n, m, Nu + (size_t)b * T * n);
  }
}

/* exact implicit VJP, per-problem gradients (same algorithm as oracle/lqr_np.py:vjp_exact) */
void FN(oracle_lqr_vjp)(int64_t Bn, int T, int n, int m, const REAL* Q, const REAL* Qf, const REAL* M, const REAL* R,
                        const REAL* A, const REAL* B, const REAL* x0, const REAL* X, const REAL* U, const REAL* Nu,
                        const REAL* Xb, const REAL* Ub, const REAL* Nub,
                        REAL* gQ, REAL* gq, REAL* gQf, REAL* gqf, REAL* gM, REAL* gR, REAL* gr, REAL* gA, REAL* gB,
                        REAL* gd, REAL* gx0) {
#pragma omp parallel
  {
    REAL* qa = (REAL*)malloc(sizeof(REAL) * T * n);
    REAL* da = (REAL*)malloc(sizeof(REAL) * T * n);
    REAL* Ka = (REAL*)malloc(sizeof(REAL) * T * m * n);
    REAL* ka = (REAL*)malloc(sizeof(REAL) * T * m);
    REAL* uX = (REAL*)malloc(sizeof(REAL) * T * n);
    REAL* uU = (REAL*)malloc(sizeof(REAL) * T * m);
    REAL* uNu = (REAL*)malloc(sizeof(REAL) * T * n);
    REAL zero[MAXD];
    memset(zero, 0, sizeof(zero));
#pragma omp for schedule(static)
    for (int64_t b = 0; b < Bn; b++) {
      const REAL* xb = Xb + (size_t)b * T * n;
      memset(qa, 0, sizeof(REAL) * n);
      memcpy(qa + n, xb, sizeof(REAL) * (size_t)(T - 1) * n);              /* q'[t] = Xb[t-1] */
      if (Nub) memcpy(da, Nub + (size_t)b * T * n, sizeof(REAL) * T * n); else memset(da, 0, sizeof(REAL) * T * n);
      lqr_t p;
      p.Q = Q + (size_t)b * T * n * n; p.q = qa; p.Qf = Qf + (size_t)b * n * n; p.qf = xb + (size_t)(T - 1) * n;
      p.M = M + (size_t)b * T * n * m; p.R = R + (size_t)b * T * m * m; p.r = Ub + (size_t)b * T * m;
      p.A = A + (size_t)b * T * n * n; p.B = B + (size_t)b * T * n * m; p.d = da;
      backward(&p, T, n, m, Ka, ka, 0);
      forward(&p, zero, Ka, ka, T, n, m, uX, uU);
      adjoint(&p, uX, uU, T, n, m, uNu);
      const REAL *Xp = X + (size_t)b * T * n, *Up = U + (size_t)b * T * m, *Np = Nu + (size_t)b * T * n;
      for (int t = 0; t < T; t++) {
        const REAL* sx = t == 0 ? x0 + (size_t)b * n : Xp + (size_t)(t - 1) * n;
        const REAL* sux = t == 0 ? zero : uX + (size_t)(t - 1) * n;
        const REAL *u = Up + (size_t)t * m, *nu = Np + (size_t)t * n, *uu = uU + (size_t)t * m, *unu = uNu + (size_t)t * n;
        size_t o = (size_t)b * T + t;
        for (int i = 0; i < n; i++) {
          for (int j = 0; j < n; j++) {
            gA[(o * n + i) * n + j] = nu[i] * sux[j] + unu[i] * sx[j];
            gQ[(o * n + i) * n + j] = (REAL)0.5 * (sux[i] * sx[j] + sx[i] * sux[j]);
          }
          for (int j = 0; j < m; j++) {
            gB[(o * n + i) * m + j] = nu[i] * uu[j] + unu[i] * u[j];
            gM[(o * n + i) * m + j] = sux[i] * u[j] + sx[i] * uu[j];
          }
          gd[o * n + i] = unu[i];
          gq[o * n + i] = sux[i];
        }
        for (int i = 0; i < m; i++) {
          for (int j = 0; j < m; j++) gR[(o * m + i) * m + j] = (REAL)0.5 * (uu[i] * u[j] + u[i] * uu[j]);
          gr[o * m + i] = uu[i];
        }
      }
      const REAL *uxT = uX + (size_t)(T - 1) * n, *xT = Xp + (size_t)(T - 1) * n;
      for (int i = 0; i < n; i++) {
        for (int j = 0; j < n; j++) gQf[((size_t)b * n + i) * n + j] = (REAL)0.5 * (uxT[i] * xT[j] + xT[i] * uxT[j]);
        gqf[(size_t)b * n + i] = uxT[i];
        REAL s = 0;
        for (int j = 0; j < m; j++) s += p.M[i * m + j] * uU[j];
        for (int j = 0; j < n; j++) s += p.A[j * n + i] * uNu[j];
        gx0[(size_t)b * n + i] = s;
      }
    }
    free(qa); free(da); free(Ka); free(ka); free(uX); free(uU); free(uNu);
  }
}
