// Shape-generic LQR kernels: runtime (n, m), one CTA per problem, thread-strided loops.  Small problems
// (everything fits the CTA's shared memory) run with 32 threads per problem out of shared memory; large ones
// (n up to 128, e.g. the lifted shared-control batch LQR of idoc/blqr.py) run with 256 threads per problem out of a
// per-problem scratch region in global memory (L2-resident).  They are the fallback for every (n, m) that has no tuned
// instantiation in shapes.def (so that, like the reference, the entry points accept any problem size), and the
// engine behind the time-invariant solver (tilqr): with `ti` set, A, B, Q, R are indexed by problem only
// (idoc/tilqr.py:12-19), absent leaves (q, r, d, M = nullptr) are zero, and the VJP sums A/B/Q/R gradients over t.
//
// Same mathematics, residual-slot layout and status convention as the tuned kernels (lqr_fwd.cuh, lqr_vjp.cuh):
//   riccati : idoc/lqr.py:77-95      rollout : idoc/lqr.py:157-161 + Nu = V x + v      adjoint sweeps: SURVEY 3.4
#pragma once
#include "common.cuh"
#include "lqr_fwd.cuh"
#include "lqr_vjp.cuh"

namespace idoc {

constexpr int GEN_MAXD = 32;    // max control dimension m (per-thread column solves keep m values)
constexpr int GEN_MAXN = 128;   // max state dimension n

// CTA-wide sum (all threads must call; red = 32 scalars of shared memory)
template <typename T>
IDOC_DEVICE T block_sum(T v, T* red) {
  for (int o = 16; o; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  const int w = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
  __syncthreads();
  if ((threadIdx.x & 31) == 0) red[w] = v;
  __syncthreads();
  T tot = T(0);
  for (int i = 0; i < nw; i++) tot += red[i];
  __syncthreads();
  return tot;
}

struct GeoRT {
  int n, m, NP, MP, oLinv, oS, oK, ok, ov, oVp, WS, o2k, o2v, WS2;
  __host__ __device__ GeoRT(int n_, int m_, int sz) : n(n_), m(m_) {
    const int e16 = 16 / sz;
    auto p16 = [e16](int w) { return (w + e16 - 1) / e16 * e16; };
    NP = n * (n + 1) / 2;
    MP = m * (m + 1) / 2;
    oLinv = 0;
    oS = MP;
    oK = p16(MP + 1);
    ok = oK + p16(m * n);
    ov = ok + p16(m);
    oVp = ov + p16(n);
    WS = oVp + p16(NP);
    o2k = 0;
    o2v = p16(m);
    WS2 = o2v + p16(n);
  }
};

__host__ __device__ inline int tri_idx(int i, int j, int n) {  // packed upper, any order of (i, j)
  const int a = i < j ? i : j, b = i < j ? j : i;
  return a * n - a * (a - 1) / 2 + (b - a);
}
__host__ __device__ inline int tril_idx(int a, int b) { return a * (a + 1) / 2 + b; }  // b <= a

template <typename T>
struct GenFwdArgs {
  LqrFwdArgs<T> a;
  int n, m, ti;
};
template <typename T>
struct GenVjpArgs {
  LqrVjpArgs<T> a;
  int n, m, ti;
};

// coalesced copy of `cnt` elements global -> shared (or zero fill when src is null)
template <typename T>
IDOC_DEVICE void g2s(T* dst, const T* src, int cnt, int lane) {
  const int nt = blockDim.x;
  if (src != nullptr)
    for (int e = lane; e < cnt; e += nt) dst[e] = src[e];
  else
    for (int e = lane; e < cnt; e += nt) dst[e] = T(0);
}

// shared-memory footprint (elements) of one warp of the generic Riccati kernel
__host__ __device__ inline int gen_riccati_elems(int n, int m, int WS) {
  const int f = n + m;
  return 2 * n * n + 2 * n * m + m * m + 2 * n + m  // slab A B Q M R q r d
         + n * n + 2 * n                             // V, v, w
         + n * f                                     // W = V [A B]
         + n * n + n * m + 2 * m * m + n + m         // Gxx, Gxu, Guu, Gt(factor), gx, gu
         + m * m + m * n + 2 * m                     // Linv(full), K, k, yg
         + WS + 16;
}

template <typename T>
__global__ void gen_riccati_kernel(const GenFwdArgs<T> ga, T* scratch) {
  extern __shared__ __align__(16) char smem_raw[];
  const LqrFwdArgs<T>& a = ga.a;
  const int n = ga.n, m = ga.m, f = n + m, Th = a.T_;
  const GeoRT G(n, m, (int)sizeof(T));
  const int lane = threadIdx.x, NT = blockDim.x;
  const long long prob = blockIdx.x;
  __shared__ T red[32];
  T* s = scratch ? scratch + (size_t)prob * gen_riccati_elems(n, m, G.WS) : reinterpret_cast<T*>(smem_raw);
  T* sA = s;            T* sB = sA + n * n;  T* sQ = sB + n * m;  T* sM = sQ + n * n;  T* sR = sM + n * m;
  T* sq = sR + m * m;   T* sr = sq + n;      T* sd = sr + m;
  T* V = sd + n;        T* v = V + n * n;    T* w = v + n;
  T* W = w + n;         // n x f
  T* Gxx = W + n * f;   T* Gxu = Gxx + n * n;  T* Guu = Gxu + n * m;  T* Gt = Guu + m * m;
  T* gx = Gt + m * m;   T* gu = gx + n;
  T* Li = gu + m;       T* K = Li + m * m;   T* kk = K + m * n;  T* yg = kk + m;
  T* slot = yg + m;

  {  // V_T = sym(Qf), v_T = qf
    const T* Qf = a.Qf + prob * n * n;
    for (int e = lane; e < n * n; e += NT) {
      const int i = e / n, j = e % n;
      V[e] = T(0.5) * (Qf[i * n + j] + Qf[j * n + i]);
    }
    g2s(v, a.qf ? a.qf + prob * n : (const T*)nullptr, n, lane);
  }
  if (ga.ti) {  // time-invariant blocks: loaded once
    g2s(sA, a.A + prob * n * n, n * n, lane);
    g2s(sB, a.B + prob * n * m, n * m, lane);
    g2s(sQ, a.Q + prob * n * n, n * n, lane);
    g2s(sR, a.R + prob * m * m, m * m, lane);
    g2s(sM, (const T*)nullptr, n * m, lane);
    g2s(sq, (const T*)nullptr, n, lane);
    g2s(sr, (const T*)nullptr, m, lane);
    g2s(sd, (const T*)nullptr, n, lane);
  }
  T dC = T(0), dc = T(0);
  int stat = 0;
  __syncthreads();

  for (int t = Th - 1; t >= 0; t--) {
    const long long idx = prob * Th + t;
    if (!ga.ti) {
      g2s(sA, a.A + idx * n * n, n * n, lane);
      g2s(sB, a.B + idx * n * m, n * m, lane);
      g2s(sQ, a.Q + idx * n * n, n * n, lane);
      g2s(sM, a.Mx ? a.Mx + idx * n * m : (const T*)nullptr, n * m, lane);
      g2s(sR, a.R + idx * m * m, m * m, lane);
      g2s(sq, a.q ? a.q + idx * n : (const T*)nullptr, n, lane);
      g2s(sr, a.r ? a.r + idx * m : (const T*)nullptr, m, lane);
      g2s(sd, a.d ? a.d + idx * n : (const T*)nullptr, n, lane);
    }
    // slot: value function entering the step
    for (int e = lane; e < n * n; e += NT) {
      const int i = e / n, j = e % n;
      if (i <= j) slot[G.oVp + tri_idx(i, j, n)] = V[e];
    }
    for (int e = lane; e < n; e += NT) slot[G.ov + e] = v[e];
    __syncthreads();
    // w = v + V d
    for (int i = lane; i < n; i += NT) {
      T acc = v[i];
      for (int l = 0; l < n; l++) acc += V[i * n + l] * sd[l];
      w[i] = acc;
    }
    // W = V [A B]
    for (int e = lane; e < n * f; e += NT) {
      const int k = e / f, j = e % f;
      T acc = T(0);
      if (j < n) for (int l = 0; l < n; l++) acc += V[k * n + l] * sA[l * n + j];
      else for (int l = 0; l < n; l++) acc += V[k * n + l] * sB[l * m + (j - n)];
      W[e] = acc;
    }
    __syncthreads();
    // G = sym([Q M; M^T R]) + [A B]^T W, gx, gu
    for (int e = lane; e < n * n; e += NT) {
      const int i = e / n, j = e % n;
      T acc = T(0.5) * (sQ[i * n + j] + sQ[j * n + i]);
      for (int k = 0; k < n; k++) acc += sA[k * n + i] * W[k * f + j];
      Gxx[e] = acc;
    }
    for (int e = lane; e < n * m; e += NT) {
      const int i = e / m, j = e % m;
      T acc = sM[e];
      for (int k = 0; k < n; k++) acc += sA[k * n + i] * W[k * f + n + j];
      Gxu[e] = acc;
    }
    for (int e = lane; e < m * m; e += NT) {
      const int i = e / m, j = e % m;
      T acc = sR[e];
      for (int k = 0; k < n; k++) acc += sB[k * m + i] * W[k * f + n + j];
      Gt[e] = acc;  // unsymmetrised
    }
    for (int i = lane; i < n; i += NT) {
      T acc = sq[i];
      for (int k = 0; k < n; k++) acc += sA[k * n + i] * w[k];
      gx[i] = acc;
    }
    for (int i = lane; i < m; i += NT) {
      T acc = sr[i];
      for (int k = 0; k < n; k++) acc += sB[k * m + i] * w[k];
      gu[i] = acc;
    }
    __syncthreads();
    for (int e = lane; e < m * m; e += NT) {
      const int i = e / m, j = e % m;
      Guu[e] = T(0.5) * (Gt[i * m + j] + Gt[j * m + i]);
    }
    __syncthreads();
    // Cholesky of Guu (+ shift), right-looking, factor in Gt (lower)
    T shift = T(0);
    for (int attempt = 0; attempt < 2; attempt++) {
      for (int e = lane; e < m * m; e += NT) Gt[e] = Guu[e] + ((e / m == e % m) ? shift : T(0));
      __syncthreads();
      bool ok = true;
      for (int j = 0; j < m; j++) {
        T p = Gt[j * m + j];
        if (!(p > T(0))) {
          ok = false;
          if (attempt == 1) { p = T(IDOC_EPS) * T(1e-3); stat |= ST_CHOL_CLAMP; }
          else {
            p = T(1);
            // tilqr factors without a shift (sym_pos=True, idoc/tilqr.py:72): a failed pivot has no retry there; the
            // reference's Cholesky solve would return NaN, so the problem is flagged and its outputs set to NaN
            if (ga.ti) stat |= ST_CHOL_CLAMP | ST_NONFINITE;
          }
        }
        const T rs = rsqrt_<T>(p);
        __syncthreads();
        for (int i = j + lane; i < m; i += NT) Gt[i * m + j] = (i == j) ? p * rs : Gt[i * m + j] * rs;
        __syncthreads();
        for (int e = lane; e < (m - j - 1) * (m - j - 1); e += NT) {
          const int i = j + 1 + e / (m - j - 1), c = j + 1 + e % (m - j - 1);
          if (c <= i) Gt[i * m + c] -= Gt[i * m + j] * Gt[c * m + j];
        }
        __syncthreads();
      }
      // Linv (lower, full storage): column b by lane b
      for (int b = lane; b < m; b += NT) {
        for (int r2 = 0; r2 < b; r2++) Li[r2 * m + b] = T(0);
        Li[b * m + b] = T(1) / Gt[b * m + b];
        for (int i = b + 1; i < m; i++) {
          T acc = T(0);
          for (int c = b; c < i; c++) acc += Gt[i * m + c] * Li[c * m + b];
          Li[i * m + b] = -acc / Gt[i * m + i];
        }
      }
      __syncthreads();
      T fro = T(0);
      for (int e = lane; e < m * m; e += NT) fro += Li[e] * Li[e];
      fro = block_sum(fro, red);
      if (attempt == 1 || ga.ti) break;  // tilqr solves with sym_pos=True and no shift (idoc/tilqr.py:72)
      if (ok && fro < T(1.0 / IDOC_EPS)) break;
      // slow path: lambda_min by Jacobi on lane 0 (scratch: yg is too small -> use W, free at this point)
      T mn = T(0);
      if (lane == 0) {
        T* Am = W;
        for (int e = 0; e < m * m; e++) Am[e] = Guu[e];
        for (int sweep = 0; sweep < 30 && m > 1; sweep++) {
          T off = T(0), dg = T(0);
          for (int i = 0; i < m; i++) { dg += Am[i * m + i] * Am[i * m + i]; for (int j = i + 1; j < m; j++) off += Am[i * m + j] * Am[i * m + j]; }
          if (!(off > (sizeof(T) == 8 ? T(1e-30) : T(1e-14)) * (dg + off))) break;
          for (int p = 0; p < m - 1; p++)
            for (int q2 = p + 1; q2 < m; q2++) {
              const T apq = Am[p * m + q2];
              if (apq == T(0)) continue;
              const T theta = (Am[q2 * m + q2] - Am[p * m + p]) / (T(2) * apq);
              const T tt = (theta >= T(0) ? T(1) : T(-1)) / (abs_(theta) + sqrt(theta * theta + T(1)));
              const T c = T(1) / sqrt(tt * tt + T(1)), sn = tt * c;
              for (int k = 0; k < m; k++) { const T x1 = Am[k * m + p], x2 = Am[k * m + q2]; Am[k * m + p] = c * x1 - sn * x2; Am[k * m + q2] = sn * x1 + c * x2; }
              for (int k = 0; k < m; k++) { const T x1 = Am[p * m + k], x2 = Am[q2 * m + k]; Am[p * m + k] = c * x1 - sn * x2; Am[q2 * m + k] = sn * x1 + c * x2; }
            }
        }
        mn = Am[0];
        for (int i = 1; i < m; i++) mn = Am[i * m + i] < mn ? Am[i * m + i] : mn;
      }
      if (lane == 0) red[0] = mn;
      __syncthreads();
      mn = red[0];
      __syncthreads();
      shift = T(IDOC_EPS) - mn;
      if (!(mn == mn) || abs_(mn) > T(1e300)) stat |= ST_NONFINITE;  // NaN / Inf in Guu (the reference propagates NaN)
      shift = shift > T(0) ? shift : T(0);
      if (shift > T(0)) stat |= ST_SHIFT;
      __syncthreads();
    }
    // k = -Linv^T Linv gu ; K columns: lane j -> column j of K
    for (int b = lane; b < m; b += NT) {
      T acc = T(0);
      for (int c = 0; c <= b; c++) acc += Li[b * m + c] * gu[c];
      yg[b] = acc;
    }
    __syncthreads();
    for (int b = lane; b < m; b += NT) {
      T acc = T(0);
      for (int c = b; c < m; c++) acc += Li[c * m + b] * yg[c];
      kk[b] = -acc;
    }
    for (int j = lane; j < n; j += NT) {
      T y[GEN_MAXD];  // m <= GEN_MAXD
      for (int b = 0; b < m; b++) {
        T acc = T(0);
        for (int c = 0; c <= b; c++) acc += Li[b * m + c] * Gxu[j * m + c];
        y[b] = acc;
      }
      for (int b = 0; b < m; b++) {
        T acc = T(0);
        for (int c = b; c < m; c++) acc += Li[c * m + b] * y[c];
        K[b * n + j] = -acc;
      }
    }
    __syncthreads();
    {  // dC, dc with the unshifted Guu
      T quad = T(0), lin = T(0);
      for (int b = lane; b < m; b += NT) {
        T acc = T(0);
        for (int c = 0; c < m; c++) acc += Guu[b * m + c] * kk[c];
        quad += acc * kk[b];
        lin += gu[b] * kk[b];
      }
      quad = block_sum(quad, red);
      lin = block_sum(lin, red);
      dC += T(0.5) * quad;
      dc += lin;
    }
    // slot: Linv (packed lower), shift, K, k
    for (int e = lane; e < m * m; e += NT) {
      const int i = e / m, j = e % m;
      if (j <= i) slot[G.oLinv + tril_idx(i, j)] = Li[e];
    }
    if (lane == 0) slot[G.oS] = shift;
    for (int e = lane; e < m * n; e += NT) slot[G.oK + e] = K[e];
    for (int e = lane; e < m; e += NT) slot[G.ok + e] = kk[e];
    // v_new = gx + K^T (gu - shift k)
    for (int j = lane; j < n; j += NT) {
      T acc = gx[j];
      for (int b = 0; b < m; b++) acc += K[b * n + j] * (gu[b] - shift * kk[b]);
      w[j] = acc;  // staged; v is still needed? no: copy below after the sync
    }
    // V_new = Gxx + Gxu K - shift K^T K  (upper triangle mirrored)
    for (int e = lane; e < n * n; e += NT) {
      const int i = e / n, j = e % n;
      if (i <= j) {
        T acc = Gxx[i * n + j];
        for (int b = 0; b < m; b++) acc += Gxu[i * m + b] * K[b * n + j] - shift * K[b * n + i] * K[b * n + j];
        W[e] = acc;  // staged in W (free)
      }
    }
    __syncthreads();
    for (int e = lane; e < n * n; e += NT) {
      const int i = e / n, j = e % n;
      V[e] = i <= j ? W[i * n + j] : W[j * n + i];
    }
    for (int e = lane; e < n; e += NT) v[e] = w[e];
    // flush slot
    T* wsl = a.ws + idx * G.WS;
    for (int e = lane; e < G.WS; e += NT) wsl[e] = slot[e];
    if (a.K != nullptr) {
      for (int e = lane; e < m * n; e += NT) a.K[idx * m * n + e] = K[e];
      for (int e = lane; e < m; e += NT) a.k[idx * m + e] = kk[e];
    }
    __syncthreads();
  }
  if (lane == 0) {
    if (a.dCdc != nullptr) { a.dCdc[prob * 2] = dC; a.dCdc[prob * 2 + 1] = dc; }
    if (a.status != nullptr) a.status[prob] = stat;
    a.wstatus[prob] = stat;
  }
}

__host__ __device__ inline int gen_rollout_elems(int n, int m, int WS) { return n * n + n * m + 3 * n + m + WS + 16; }

template <typename T>
__global__ void gen_rollout_kernel(const GenFwdArgs<T> ga, T* scratch) {
  extern __shared__ __align__(16) char smem_raw[];
  const LqrFwdArgs<T>& a = ga.a;
  const int n = ga.n, m = ga.m, Th = a.T_;
  const GeoRT G(n, m, (int)sizeof(T));
  const int lane = threadIdx.x, NT = blockDim.x;
  const long long prob = blockIdx.x;
  __shared__ T red[32];
  T* s = scratch ? scratch + (size_t)prob * gen_rollout_elems(n, m, G.WS) : reinterpret_cast<T*>(smem_raw);
  T* sA = s; T* sB = sA + n * n; T* sd = sB + n * m; T* x = sd + n; T* xn = x + n; T* u = xn + n; T* slot = u + m;
  g2s(x, a.x0 + prob * n, n, lane);
  if (ga.ti) {
    g2s(sA, a.A + prob * n * n, n * n, lane);
    g2s(sB, a.B + prob * n * m, n * m, lane);
    g2s(sd, (const T*)nullptr, n, lane);
  }
  for (int t = 0; t < Th; t++) {
    const long long idx = prob * Th + t;
    if (!ga.ti) {
      g2s(sA, a.A + idx * n * n, n * n, lane);
      g2s(sB, a.B + idx * n * m, n * m, lane);
      g2s(sd, a.d ? a.d + idx * n : (const T*)nullptr, n, lane);
    }
    g2s(slot, a.ws + idx * G.WS, G.WS, lane);
    __syncthreads();
    for (int b = lane; b < m; b += NT) {
      T acc = slot[G.ok + b];
      for (int j = 0; j < n; j++) acc += slot[G.oK + b * n + j] * x[j];
      u[b] = acc;
      a.U[idx * m + b] = acc;
    }
    __syncthreads();
    for (int i = lane; i < n; i += NT) {
      T acc = sd[i];
      for (int j = 0; j < n; j++) acc += sA[i * n + j] * x[j];
      for (int b = 0; b < m; b++) acc += sB[i * m + b] * u[b];
      xn[i] = acc;
      a.X[idx * n + i] = acc;
    }
    __syncthreads();
    for (int i = lane; i < n; i += NT) {
      T acc = slot[G.ov + i];
      for (int j = 0; j < n; j++) acc += slot[G.oVp + tri_idx(i, j, n)] * xn[j];
      a.Nu[idx * n + i] = acc;
      x[i] = xn[i];
    }
    __syncthreads();
  }
}

__host__ __device__ inline int gen_adj_elems(int n, int m, int WS, int WS2) {
  return n * n + n * m + WS + WS2 + 6 * n + 3 * m + 2 * (2 * n * n + n * m + m * m) + 32;
}

// adjoint backward sweep: v', k'   (ws2 slot t = [k'_t | v'_{t+1}])
template <typename T>
__global__ void gen_adj_bwd_kernel(const GenVjpArgs<T> ga, T* scratch) {
  extern __shared__ __align__(16) char smem_raw[];
  const LqrVjpArgs<T>& a = ga.a;
  const int n = ga.n, m = ga.m, Th = a.T_;
  const GeoRT G(n, m, (int)sizeof(T));
  const int lane = threadIdx.x, NT = blockDim.x;
  const long long prob = blockIdx.x;
  __shared__ T red[32];
  T* s = scratch ? scratch + (size_t)prob * gen_adj_elems(n, m, G.WS, G.WS2) : reinterpret_cast<T*>(smem_raw);
  T* sA = s; T* sB = sA + n * n; T* slot = sB + n * m; T* vp = slot + G.WS; T* w = vp + n; T* gx = w + n; T* gu = gx + n;
  T* yg = gu + m; T* kk = yg + m;
  g2s(vp, a.Xb + (prob * Th + Th - 1) * n, n, lane);  // v'_T = qf' = Xb[T-1]
  if (ga.ti) {
    g2s(sA, a.A + prob * n * n, n * n, lane);
    g2s(sB, a.B + prob * n * m, n * m, lane);
  }
  __syncthreads();
  for (int t = Th - 1; t >= 0; t--) {
    const long long idx = prob * Th + t;
    if (!ga.ti) {
      g2s(sA, a.A + idx * n * n, n * n, lane);
      g2s(sB, a.B + idx * n * m, n * m, lane);
    }
    g2s(slot, a.ws + idx * G.WS, G.WS, lane);
    T* out = a.ws2 + idx * G.WS2;
    for (int e = lane; e < n; e += NT) out[G.o2v + e] = vp[e];
    __syncthreads();
    for (int i = lane; i < n; i += NT) {
      T acc = vp[i];
      if (a.Nub != nullptr)
        for (int l = 0; l < n; l++) acc += slot[G.oVp + tri_idx(i, l, n)] * a.Nub[idx * n + l];
      w[i] = acc;
    }
    __syncthreads();
    for (int j = lane; j < n; j += NT) {
      T acc = t > 0 ? a.Xb[(idx - 1) * n + j] : T(0);
      for (int k = 0; k < n; k++) acc += sA[k * n + j] * w[k];
      gx[j] = acc;
    }
    for (int b = lane; b < m; b += NT) {
      T acc = a.Ub[idx * m + b];
      for (int k = 0; k < n; k++) acc += sB[k * m + b] * w[k];
      gu[b] = acc;
    }
    __syncthreads();
    for (int b = lane; b < m; b += NT) {
      T acc = T(0);
      for (int c = 0; c <= b; c++) acc += slot[G.oLinv + tril_idx(b, c)] * gu[c];
      yg[b] = acc;
    }
    __syncthreads();
    for (int b = lane; b < m; b += NT) {
      T acc = T(0);
      for (int c = b; c < m; c++) acc += slot[G.oLinv + tril_idx(c, b)] * yg[c];
      kk[b] = -acc;
      out[G.o2k + b] = -acc;
    }
    __syncthreads();
    const T shift = slot[G.oS];
    for (int j = lane; j < n; j += NT) {
      T acc = gx[j];
      for (int b = 0; b < m; b++) acc += slot[G.oK + b * n + j] * (gu[b] - shift * kk[b]);
      w[j] = acc;
    }
    __syncthreads();
    for (int e = lane; e < n; e += NT) vp[e] = w[e];
    __syncthreads();
  }
}

// adjoint forward sweep + gradients
template <typename T>
__global__ void gen_adj_fwd_kernel(const GenVjpArgs<T> ga, T* scratch) {
  extern __shared__ __align__(16) char smem_raw[];
  const LqrVjpArgs<T>& a = ga.a;
  const int n = ga.n, m = ga.m, Th = a.T_;
  const GeoRT G(n, m, (int)sizeof(T));
  const int lane = threadIdx.x, NT = blockDim.x;
  const long long prob = blockIdx.x;
  __shared__ T red[32];
  const bool reduce = a.reduce_batch != 0;
  T* s = scratch ? scratch + (size_t)prob * gen_adj_elems(n, m, G.WS, G.WS2) : reinterpret_cast<T*>(smem_raw);
  T* sA = s; T* sB = sA + n * n; T* slot = sB + n * m; T* s2 = slot + G.WS;
  T* ux = s2 + G.WS2; T* uxn = ux + n; T* sX = uxn + n; T* nu = sX + n; T* unu = nu + n; T* dp = unu + n;
  T* uU = dp + n; T* Uv = uU + m; T* tmpm = Uv + m;
  T* accA = tmpm + m; T* accB = accA + n * n; T* accQ = accB + n * m; T* accR = accQ + n * n;  // time sums (ti)
  for (int e = lane; e < n; e += NT) { ux[e] = T(0); sX[e] = a.x0[prob * n + e]; }
  if (ga.ti) {
    g2s(sA, a.A + prob * n * n, n * n, lane);
    g2s(sB, a.B + prob * n * m, n * m, lane);
    for (int e = lane; e < 2 * n * n + n * m + m * m; e += NT) accA[e] = T(0);
  }
  __syncthreads();
  for (int t = 0; t < Th; t++) {
    const long long idx = prob * Th + t;
    if (!ga.ti) {
      g2s(sA, a.A + idx * n * n, n * n, lane);
      g2s(sB, a.B + idx * n * m, n * m, lane);
    }
    g2s(slot, a.ws + idx * G.WS, G.WS, lane);
    g2s(s2, a.ws2 + idx * G.WS2, G.WS2, lane);
    g2s(Uv, a.U + idx * m, m, lane);
    g2s(nu, a.Nu + idx * n, n, lane);
    g2s(dp, a.Nub ? a.Nub + idx * n : (const T*)nullptr, n, lane);
    __syncthreads();
    for (int b = lane; b < m; b += NT) {
      T acc = s2[G.o2k + b];
      for (int j = 0; j < n; j++) acc += slot[G.oK + b * n + j] * ux[j];
      uU[b] = acc;
    }
    __syncthreads();
    for (int i = lane; i < n; i += NT) {
      T acc = dp[i];
      for (int j = 0; j < n; j++) acc += sA[i * n + j] * ux[j];
      for (int b = 0; b < m; b++) acc += sB[i * m + b] * uU[b];
      uxn[i] = acc;
    }
    __syncthreads();
    for (int i = lane; i < n; i += NT) {
      T acc = s2[G.o2v + i];
      for (int j = 0; j < n; j++) acc += slot[G.oVp + tri_idx(i, j, n)] * uxn[j];
      unu[i] = acc;
    }
    __syncthreads();
    // gradients of step t
    if (ga.ti) {
      for (int e = lane; e < n * n; e += NT) {
        const int i = e / n, j = e % n;
        accA[e] += nu[i] * ux[j] + unu[i] * sX[j];
        accQ[e] += T(0.5) * (ux[i] * sX[j] + sX[i] * ux[j]);
      }
      for (int e = lane; e < n * m; e += NT) { const int i = e / m, j = e % m; accB[e] += nu[i] * uU[j] + unu[i] * Uv[j]; }
      for (int e = lane; e < m * m; e += NT) { const int i = e / m, j = e % m; accR[e] += T(0.5) * (uU[i] * Uv[j] + Uv[i] * uU[j]); }
    } else {
      // vectors-only mode (batch-summed gradients, grad_reduce.cuh): gq = sux, gr = uU, gd = uNu per problem, no outer product
      if (!reduce) {
        for (int e = lane; e < n * n; e += NT) {
          const int i = e / n, j = e % n;
          a.gA[idx * n * n + e] = nu[i] * ux[j] + unu[i] * sX[j];
          a.gQ[idx * n * n + e] = T(0.5) * (ux[i] * sX[j] + sX[i] * ux[j]);
        }
        for (int e = lane; e < n * m; e += NT) {
          const int i = e / m, j = e % m;
          a.gB[idx * n * m + e] = nu[i] * uU[j] + unu[i] * Uv[j];
          a.gM[idx * n * m + e] = ux[i] * Uv[j] + sX[i] * uU[j];
        }
        for (int e = lane; e < m * m; e += NT) {
          const int i = e / m, j = e % m;
          a.gR[idx * m * m + e] = T(0.5) * (uU[i] * Uv[j] + Uv[i] * uU[j]);
        }
      }
      for (int e = lane; e < n; e += NT) { a.gd[idx * n + e] = unu[e]; a.gq[idx * n + e] = ux[e]; }
      for (int e = lane; e < m; e += NT) a.gr[idx * m + e] = uU[e];
    }
    if (t == 0) {  // gx0 = M[0] uU[0] + A[0]^T uNu[0]
      for (int i = lane; i < n; i += NT) {
        T acc = T(0);
        if (a.Mx != nullptr && !ga.ti) for (int b = 0; b < m; b++) acc += a.Mx[(prob * Th) * n * m + i * m + b] * uU[b];
        for (int k = 0; k < n; k++) acc += sA[k * n + i] * unu[k];
        a.gx0[prob * n + i] = acc;
      }
    }
    if (t == Th - 1) {
      const T* XT = a.X + idx * n;
      if (!reduce || ga.ti) {
        T* gQf = a.gQf + prob * n * n;
        for (int e = lane; e < n * n; e += NT) {
          const int i = e / n, j = e % n;
          gQf[e] = T(0.5) * (uxn[i] * XT[j] + uxn[j] * XT[i]);
        }
      }
      if (a.gqf != nullptr)
        for (int e = lane; e < n; e += NT) a.gqf[prob * n + e] = uxn[e];
    }
    __syncthreads();
    for (int e = lane; e < n; e += NT) { ux[e] = uxn[e]; sX[e] = a.X[idx * n + e]; }
    __syncthreads();
  }
  if (ga.ti) {  // time-summed gradients of the time-invariant leaves (idoc/tilqr.py:38-54 in adjoint form)
    for (int e = lane; e < n * n; e += NT) { a.gA[prob * n * n + e] = accA[e]; a.gQ[prob * n * n + e] = accQ[e]; }
    for (int e = lane; e < n * m; e += NT) a.gB[prob * n * m + e] = accB[e];
    for (int e = lane; e < m * m; e += NT) a.gR[prob * m * m + e] = accR[e];
  }
}

}  // namespace idoc
