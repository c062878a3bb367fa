// Forward solve of the batched time-varying LQR (reference: idoc/lqr.py:62-107 backward,
// :153-164 forward, :110-125 adjoint, composed at :170-175), as two sm_100a kernels:
//
//   lqr_riccati_kernel : time-reversed Riccati sweep, t = T-1..0.  Streams A,B,Q,M,R,q,r,d of step t
//                        (PER_STEP scalars) once, writes the residual slot ws[b][t] (gains K_t,k_t, inverse
//                        Cholesky factor of Gtuu_t, eigen-shift, and the value function (V_{t+1},v_{t+1})
//                        entering the step).
//   lqr_rollout_kernel : forward sweep, t = 0..T-1.  Streams A,B,d and the residual slot, writes X,U and
//                        the co-states as Nu[t] = V_{t+1} x_{t+1} + v_{t+1}  (equal to the reference's
//                        adjoint recursion lqr.py:110-125 for an LQR solution; it removes the third sweep).
//
// One group of L lanes per problem, G = 32/L problems per warp, warps fully independent (see common.cuh).
#pragma once
#include "loader.cuh"
#include "lqr_geo.cuh"
#include "small_linalg.cuh"

namespace idoc {

template <typename T>
struct LqrFwdArgs {
  const T *Q, *q, *Qf, *qf, *Mx, *R, *r, *A, *B, *d, *x0;
  T *X, *U, *Nu;  // outputs
  T *K, *k;       // optional user-visible gains [B,T,m,n], [B,T,m] (may be null)
  T* ws;          // residual slots [B,T,WS]
  T* dCdc;        // optional [B,2]: expected-change terms of lqr.py:93-94,104-105
  int* status;    // optional [B]
  int* wstatus;   // [B] status words kept at the tail of ws (always written; drives the co-state fix-up)
  long long nb;   // batch size
  int T_;         // horizon
};

constexpr double IDOC_EPS = 1e-8;  // lqr.py:73

// ======================================================================================= Riccati
template <typename T, int N, int M, int L, int NBUF>
struct RiccatiCfg {
  using Ge = Geo<T, N, M>;
  static constexpr int SZ = Ge::SZ;
  static constexpr int G = 32 / L;
  static constexpr int CX = cdiv(N, L), CU = cdiv(M, L);
  static constexpr int PQ = cdiv(Ge::NP, L);
  // in-slab (elements)
  static constexpr int iA = 0;
  static constexpr int iB = iA + Ge::p16(Ge::wA);
  static constexpr int iQ = iB + Ge::p16(Ge::wB);
  static constexpr int iM = iQ + Ge::p16(Ge::wQ);
  static constexpr int iR = iM + Ge::p16(Ge::wM);
  static constexpr int iq = iR + Ge::p16(Ge::wR);
  static constexpr int ir = iq + Ge::p16(Ge::wq);
  static constexpr int id = ir + Ge::p16(Ge::wr);
  static constexpr int IN = id + Ge::p16(Ge::wd);
  // exchange area (elements).  Vf (full V_new, written in phase 7) aliases Guu|Gxu|gu, which are dead after
  // phase 6; Y is read in phase 7 and stays separate.
  static constexpr int eGuu = 0;
  static constexpr int eGxu = eGuu + Ge::p16(M * M);
  static constexpr int egu = eGxu + Ge::p16(N * M);
  static constexpr int eVf = 0;
  static constexpr int eY = cmax(egu + Ge::p16(M), Ge::p16(N * N));
  static constexpr int EX = eY + Ge::p16(M * N);
  // group slab
  static constexpr int oIn = 0;
  static constexpr int oOut = NBUF * IN;
  static constexpr int oEx = oOut + 2 * Ge::WS;
  static constexpr int GROUP_BYTES = odd16((oEx + EX) * SZ);
  static constexpr int WARP_BYTES = WARP_HDR_BYTES + G * GROUP_BYTES;
  static constexpr bool TMA = Ge::GR == 16;
  static constexpr int LOAD_CHUNKS = Ge::PER_STEP * SZ / Ge::GR;
  static constexpr int RL = cdiv(LOAD_CHUNKS, 32);
  static constexpr int RS = cdiv(Ge::WS * SZ / 16, 32);
};

template <typename T, int N, int M, int L, int NBUF, int WARPS>
__global__ void __launch_bounds__(WARPS * 32) lqr_riccati_kernel(const LqrFwdArgs<T> a) {
  using C = RiccatiCfg<T, N, M, L, NBUF>;
  using Ge = Geo<T, N, M>;
  constexpr int SZ = C::SZ, G = C::G, CX = C::CX, CU = C::CU, NP = Ge::NP, MP = Ge::MP, WS = Ge::WS;
  constexpr int RA_N = row_align(N, SZ), RA_M = row_align(M, SZ);
  extern __shared__ __align__(16) char smem_raw[];

  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int grp = lane / L, sub = lane % L;
  const long long prob0 = ((long long)blockIdx.x * WARPS + warp) * G;
  if (prob0 >= a.nb) return;  // warp-uniform; no CTA-wide barriers anywhere below
  char* wregion = smem_raw + warp * C::WARP_BYTES;
  char* wbase = wregion + WARP_HDR_BYTES;
  T* gs = reinterpret_cast<T*>(wbase + grp * C::GROUP_BYTES);
  const long long last = a.nb - 1;
  const long long prob = prob0 + grp;
  const bool active = prob <= last;
  const long long probc = active ? prob : last;  // idle groups shadow the last problem, stores masked
  const int Th = a.T_;

  SlabLoader<C::TMA, G, Ge::GR, 8, C::RL, 0> ld;
  ld.init(wregion, prob0, last, Th, C::GROUP_BYTES, lane);
  ld.add(a.A, Ge::wA * SZ, 0, Ge::wA * SZ, C::iA * SZ);
  ld.add(a.B, Ge::wB * SZ, 0, Ge::wB * SZ, C::iB * SZ);
  ld.add(a.Q, Ge::wQ * SZ, 0, Ge::wQ * SZ, C::iQ * SZ);
  ld.add(a.Mx, Ge::wM * SZ, 0, Ge::wM * SZ, C::iM * SZ);
  ld.add(a.R, Ge::wR * SZ, 0, Ge::wR * SZ, C::iR * SZ);
  ld.add(a.q, Ge::wq * SZ, 0, Ge::wq * SZ, C::iq * SZ);
  ld.add(a.r, Ge::wr * SZ, 0, Ge::wr * SZ, C::ir * SZ);
  ld.add(a.d, Ge::wd * SZ, 0, Ge::wd * SZ, C::id * SZ);
  SlabStorer<C::TMA, G, 16, 1, C::RS> st;
  st.init(prob0, last, Th, C::GROUP_BYTES, lane);
  st.add(a.ws, WS * SZ, 0, WS * SZ, 0);

  auto issue_loads = [&](int t, int buf) { ld.issue(wbase, (C::oIn + buf * C::IN) * SZ, t, buf); };

  // own columns
  int jx[CX], ju[CU];
  bool vx[CX], vu[CU];
#pragma unroll
  for (int c = 0; c < CX; c++) {
    const int j = sub * CX + c;
    vx[c] = j < N;
    jx[c] = vx[c] ? j : N - 1;
  }
#pragma unroll
  for (int c = 0; c < CU; c++) {
    const int j = sub * CU + c;
    vu[c] = j < M;
    ju[c] = vu[c] ? j : M - 1;
  }
  // own entries of the packed symmetric V: offsets of (i,j) and (j,i) in the full exchange matrix
  int pkA[C::PQ], pkB[C::PQ];
  {
    int e = 0;
#pragma unroll
    for (int i = 0; i < N; i++)
#pragma unroll
      for (int j = i; j < N; j++) {
#pragma unroll
        for (int qq = 0; qq < C::PQ; qq++)
          if (e == sub + L * qq) {
            pkA[qq] = i * N + j;
            pkB[qq] = j * N + i;
          }
        e++;
      }
  }

  issue_loads(Th - 1, (Th - 1) % NBUF);

  // V_T = sym(Qf), v_T = qf into the slab that step T-1 reads
  {
    T* cur = gs + C::oOut + ((Th - 1) & 1) * WS;
    const T* Qf = a.Qf + probc * (N * N);
    const T* qf = a.qf + probc * N;
#pragma unroll
    for (int qq = 0; qq < C::PQ; qq++) {
      const int e = sub + L * qq;
      if (e < NP) cur[Ge::oVp + e] = T(0.5) * (Qf[pkA[qq]] + Qf[pkB[qq]]);
    }
    for (int i = sub; i < N; i += L) cur[Ge::ov + i] = qf[i];
  }
  T dC = T(0), dc = T(0);
  int stat = 0;

  for (int t = Th - 1; t >= 0; t--) {
    const T* in = gs + C::oIn + (t % NBUF) * C::IN;
    T* cur = gs + C::oOut + (t & 1) * WS;
    T* nxt = gs + C::oOut + ((t & 1) ^ 1) * WS;
    T* ex = gs + C::oEx;

    st.drain();
    ld.wait(t % NBUF);
    __syncwarp();
    if (NBUF > 1 && t > 0) issue_loads(t - 1, (t - 1) % NBUF);

    // ---- 1. w = v + V d                                          (lqr.py:83-85: Vd, A^T v + A^T Vd)
    T Vr[NP], w[N];
    lds<T, NP, 16>(cur + Ge::oVp, Vr);
    {
      T vv[N], dd[N];
      lds<T, N, 16>(cur + Ge::ov, vv);
      lds<T, N, 16>(in + C::id, dd);
#pragma unroll
      for (int k = 0; k < N; k++) {
        T acc = vv[k];
#pragma unroll
        for (int l = 0; l < N; l++) acc += Vr[triu(k, l, N)] * dd[l];
        w[k] = acc;
      }
    }
    // ---- 2. own columns f of [A B], and W = V f
    T Wx[CX][N], Wu[CU][N], gx[CX], gu_[CU];
#pragma unroll
    for (int c = 0; c < CX; c++) {
      T f[N];
#pragma unroll
      for (int k = 0; k < N; k++) f[k] = in[C::iA + k * N + jx[c]];
      T acc = in[C::iq + jx[c]];
#pragma unroll
      for (int k = 0; k < N; k++) acc += f[k] * w[k];
      gx[c] = acc;  // gx = q + A^T (v + V d)                        lqr.py:84
#pragma unroll
      for (int k = 0; k < N; k++) {
        T s = T(0);
#pragma unroll
        for (int l = 0; l < N; l++) s += Vr[triu(k, l, N)] * f[l];
        Wx[c][k] = s;
      }
    }
#pragma unroll
    for (int c = 0; c < CU; c++) {
      T f[N];
#pragma unroll
      for (int k = 0; k < N; k++) f[k] = in[C::iB + k * M + ju[c]];
      T acc = in[C::ir + ju[c]];
#pragma unroll
      for (int k = 0; k < N; k++) acc += f[k] * w[k];
      gu_[c] = acc;  // gu = r + B^T (v + V d)                       lqr.py:85
#pragma unroll
      for (int k = 0; k < N; k++) {
        T s = T(0);
#pragma unroll
        for (int l = 0; l < N; l++) s += Vr[triu(k, l, N)] * f[l];
        Wu[c][k] = s;
      }
    }
    // ---- 3. own columns of G = [Q M; M^T R] + [A B]^T V [A B]      lqr.py:80-82
    T Gx[CX][N], Gxu[CU][N], Guu[CU][M];
#pragma unroll
    for (int c = 0; c < CX; c++)
#pragma unroll
      for (int i = 0; i < N; i++) Gx[c][i] = in[C::iQ + i * N + jx[c]];
#pragma unroll
    for (int c = 0; c < CU; c++) {
#pragma unroll
      for (int i = 0; i < N; i++) Gxu[c][i] = in[C::iM + i * M + ju[c]];
#pragma unroll
      for (int b = 0; b < M; b++) Guu[c][b] = in[C::iR + b * M + ju[c]];
    }
#pragma unroll
    for (int k = 0; k < N; k++) {
      T Ar[N], Br[M];
      lds<T, N, RA_N>(in + C::iA + k * N, Ar);
      lds<T, M, RA_M>(in + C::iB + k * M, Br);
#pragma unroll
      for (int c = 0; c < CX; c++)
#pragma unroll
        for (int i = 0; i < N; i++) Gx[c][i] += Ar[i] * Wx[c][k];
#pragma unroll
      for (int c = 0; c < CU; c++) {
#pragma unroll
        for (int i = 0; i < N; i++) Gxu[c][i] += Ar[i] * Wu[c][k];
#pragma unroll
        for (int b = 0; b < M; b++) Guu[c][b] += Br[b] * Wu[c][k];
      }
    }
    // ---- 4. publish the control-block columns
#pragma unroll
    for (int c = 0; c < CU; c++)
      if (vu[c]) {
#pragma unroll
        for (int i = 0; i < N; i++) ex[C::eGxu + i * M + ju[c]] = Gxu[c][i];
#pragma unroll
        for (int b = 0; b < M; b++) ex[C::eGuu + b * M + ju[c]] = Guu[c][b];
        ex[C::egu + ju[c]] = gu_[c];
      }
    __syncwarp();
    if (NBUF == 1 && t > 0) issue_loads(t - 1, 0);  // all reads of the in-slab are done

    // ---- 5. every lane: Guu -> (shift) -> inverse Cholesky factor; k   lqr.py:81,86-90
    T Gs[MP], Li[MP], guv[M], kk[M];
    T shift = T(0);
    {
      T Gm[M * M];
      lds<T, M * M, 16>(ex + C::eGuu, Gm);
      lds<T, M, 16>(ex + C::egu, guv);
#pragma unroll
      for (int b = 0; b < M; b++)
#pragma unroll
        for (int c2 = b; c2 < M; c2++) Gs[triu(b, c2, M)] = T(0.5) * (Gm[b * M + c2] + Gm[c2 * M + b]);
      bool clamped = false;
      bool ok = chol_inv<T, M, false>(Gs, Li, T(0), &clamped);
      T fro = T(0);
#pragma unroll
      for (int e = 0; e < MP; e++) fro += Li[e] * Li[e];
      // lambda_min(Guu) >= 1/||Linv||_F^2: if that exceeds EPS no shift is needed (the common case)
      if (!ok || !(fro < T(1.0 / IDOC_EPS))) {
        const T mn = min_eig_jacobi<T, M>(Gs);
        shift = T(IDOC_EPS) - mn;
        shift = shift > T(0) ? shift : T(0);  // lqr.py:88
        T Gt[MP];
#pragma unroll
        for (int e = 0; e < MP; e++) Gt[e] = Gs[e];
#pragma unroll
        for (int b = 0; b < M; b++) Gt[triu(b, b, M)] += shift;
        chol_inv<T, M, true>(Gt, Li, T(IDOC_EPS) * T(1e-3), &clamped);
        if (shift > T(0)) stat |= ST_SHIFT;
        if (clamped) stat |= ST_CHOL_CLAMP;
      }
      T yg[M];
#pragma unroll
      for (int b = 0; b < M; b++) {
        T s = T(0);
#pragma unroll
        for (int c2 = 0; c2 <= b; c2++) s += Li[tril(b, c2)] * guv[c2];
        yg[b] = s;
      }
#pragma unroll
      for (int b = 0; b < M; b++) {
        T s = T(0);
#pragma unroll
        for (int c2 = b; c2 < M; c2++) s += Li[tril(c2, b)] * yg[c2];
        kk[b] = -s;  // k = -Gtuu^{-1} gu                              lqr.py:90
      }
      // expected-change terms with the UNSHIFTED Guu                  lqr.py:93-94
      T quad = T(0), lin = T(0);
#pragma unroll
      for (int b = 0; b < M; b++) {
        T s = T(0);
#pragma unroll
        for (int c2 = 0; c2 < M; c2++) s += Gs[triu(b, c2, M)] * kk[c2];
        quad += s * kk[b];
        lin += guv[b] * kk[b];
      }
      dC += T(0.5) * quad;
      dc += lin;
    }
    // ---- 6. own state columns: Y = Linv Gux, K = -Linv^T Y, v_new     lqr.py:89,92
    T Yc[CX][M], Kc[CX][M];
#pragma unroll
    for (int c = 0; c < CX; c++) {
      T grow[M];
      lds<T, M, RA_M>(ex + C::eGxu + jx[c] * M, grow);
#pragma unroll
      for (int b = 0; b < M; b++) {
        T s = T(0);
#pragma unroll
        for (int c2 = 0; c2 <= b; c2++) s += Li[tril(b, c2)] * grow[c2];
        Yc[c][b] = s;
      }
#pragma unroll
      for (int b = 0; b < M; b++) {
        T s = T(0);
#pragma unroll
        for (int c2 = b; c2 < M; c2++) s += Li[tril(c2, b)] * Yc[c][c2];
        Kc[c][b] = -s;
      }
      // v = gx + Gxu k + K^T gu + K^T Guu k  ==  gx + K^T (gu - shift k)   (Guu = Gtuu - shift I)
      T vn = gx[c];
#pragma unroll
      for (int b = 0; b < M; b++) vn += Kc[c][b] * (guv[b] - shift * kk[b]);
      if (vx[c]) {
#pragma unroll
        for (int b = 0; b < M; b++) {
          cur[Ge::oK + b * N + jx[c]] = Kc[c][b];
          ex[C::eY + b * N + jx[c]] = Yc[c][b];
        }
        nxt[Ge::ov + jx[c]] = vn;
      }
    }
    if (sub == 0) {
      sts<T, MP, 16>(cur + Ge::oLinv, Li);
      cur[Ge::oS] = shift;
      sts<T, M, 16>(cur + Ge::ok, kk);
    }
    __syncwarp();
    // ---- 7. V_new columns: Gxx + Gxu K + K^T Gux + K^T Guu K == Gxx - Y^T Y - shift K^T K   lqr.py:91
    {
      T Yf[M * N];
      lds<T, M * N, 16>(ex + C::eY, Yf);
#pragma unroll
      for (int c = 0; c < CX; c++)
#pragma unroll
        for (int i = 0; i < N; i++) {
          T s = Gx[c][i];
#pragma unroll
          for (int b = 0; b < M; b++) s -= Yf[b * N + i] * Yc[c][b];
          Gx[c][i] = s;
        }
      if (shift > T(0)) {
        T Kf[M * N];
        lds<T, M * N, 16>(cur + Ge::oK, Kf);
#pragma unroll
        for (int c = 0; c < CX; c++)
#pragma unroll
          for (int i = 0; i < N; i++) {
            T s = T(0);
#pragma unroll
            for (int b = 0; b < M; b++) s += Kf[b * N + i] * Kc[c][b];
            Gx[c][i] -= shift * s;
          }
      }
#pragma unroll
      for (int c = 0; c < CX; c++)
        if (vx[c]) {
#pragma unroll
          for (int i = 0; i < N; i++) ex[C::eVf + i * N + jx[c]] = Gx[c][i];
        }
    }
    // flush the completed slot t to HBM (bulk async stores, or coalesced 16-byte stores on the cp.async path)
    st.store(wbase, (C::oOut + (t & 1) * WS) * SZ, t);
    if (a.K != nullptr && active) {
      for (int e = sub; e < M * N; e += L) a.K[(prob * Th + t) * (M * N) + e] = cur[Ge::oK + e];
      for (int e = sub; e < M; e += L) a.k[(prob * Th + t) * M + e] = cur[Ge::ok + e];
    }
    __syncwarp();
    // ---- 8. symmetrise + pack V_new into the next slab                  lqr.py:91 (symmetrize)
#pragma unroll
    for (int qq = 0; qq < C::PQ; qq++) {
      const int e = sub + L * qq;
      if (e < NP) nxt[Ge::oVp + e] = T(0.5) * (ex[C::eVf + pkA[qq]] + ex[C::eVf + pkB[qq]]);
    }
    // (the __syncwarp at the top of the next iteration orders these writes before the next reads)
  }
  st.drain();
  if (active && sub == 0) {
    if (a.dCdc != nullptr) {
      a.dCdc[prob * 2 + 0] = dC;
      a.dCdc[prob * 2 + 1] = dc;
    }
    if (a.status != nullptr) a.status[prob] = stat;
    a.wstatus[prob] = stat;
  }
}

// ======================================================================================= rollout
template <typename T, int N, int M, int L, int NBUF, int OC>
struct RolloutCfg {
  using Ge = Geo<T, N, M>;
  static constexpr int SZ = Ge::SZ;
  static constexpr int G = 32 / L;
  static constexpr int CX = cdiv(N, L), CU = cdiv(M, L);
  // in-slab: A | B | d | ws[oK .. WS)  (K, k, v, Vp)
  static constexpr int iA = 0;
  static constexpr int iB = iA + Ge::p16(Ge::wA);
  static constexpr int id = iB + Ge::p16(Ge::wB);
  static constexpr int iW = id + Ge::p16(Ge::wd);  // ws sub-range starts here; field f at iW + (Ge::of - Ge::oK)
  static constexpr int WSUB = Ge::WS - Ge::oK;
  static constexpr int IN = iW + WSUB;
  // state + out staging: x | u | out chunk of OC steps: X (OC*N) | U (OC*M) | Nu (OC*N), each 16B padded
  static constexpr int sx = 0;
  static constexpr int su = sx + Ge::p16(N);
  static constexpr int soX = su + Ge::p16(M);
  static constexpr int soU = soX + Ge::p16(OC * N);
  static constexpr int soN = soU + Ge::p16(OC * M);
  static constexpr int ST = soN + Ge::p16(OC * N);
  static constexpr int oIn = 0;
  static constexpr int oSt = NBUF * IN;
  static constexpr int GROUP_BYTES = odd16((oSt + ST) * SZ);
  static constexpr int WARP_BYTES = WARP_HDR_BYTES + G * GROUP_BYTES;
  static constexpr bool TMA = Ge::GR == 16;
  static constexpr int LOAD_CHUNKS = (Ge::wA + Ge::wB + Ge::wd) * SZ / Ge::GR + WSUB * SZ / Ge::GR;
  static constexpr int RL = cdiv(LOAD_CHUNKS, 32);
};

template <typename T, int N, int M, int L, int NBUF, int OC, int WARPS>
__global__ void __launch_bounds__(WARPS * 32) lqr_rollout_kernel(const LqrFwdArgs<T> a) {
  using C = RolloutCfg<T, N, M, L, NBUF, OC>;
  using Ge = Geo<T, N, M>;
  constexpr int SZ = C::SZ, G = C::G, CX = C::CX, CU = C::CU, NP = Ge::NP, WS = Ge::WS;
  constexpr int RA_N = row_align(N, SZ), RA_M = row_align(M, SZ);
  extern __shared__ __align__(16) char smem_raw[];

  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int grp = lane / L, sub = lane % L;
  const long long prob0 = ((long long)blockIdx.x * WARPS + warp) * G;
  if (prob0 >= a.nb) return;
  char* wregion = smem_raw + warp * C::WARP_BYTES;
  char* wbase = wregion + WARP_HDR_BYTES;
  T* gs = reinterpret_cast<T*>(wbase + grp * C::GROUP_BYTES);
  const long long last = a.nb - 1;
  const long long prob = prob0 + grp;
  const bool active = prob <= last;
  const long long probc = active ? prob : last;
  const int Th = a.T_;

  SlabLoader<C::TMA, G, Ge::GR, 4, C::RL, 0> ld;
  ld.init(wregion, prob0, last, Th, C::GROUP_BYTES, lane);
  ld.add(a.A, Ge::wA * SZ, 0, Ge::wA * SZ, C::iA * SZ);
  ld.add(a.B, Ge::wB * SZ, 0, Ge::wB * SZ, C::iB * SZ);
  ld.add(a.d, Ge::wd * SZ, 0, Ge::wd * SZ, C::id * SZ);
  ld.add(a.ws, WS * SZ, Ge::oK * SZ, C::WSUB * SZ, C::iW * SZ);

  auto issue_loads = [&](int t, int buf) { ld.issue(wbase, (C::oIn + buf * C::IN) * SZ, t, buf); };
  issue_loads(0, 0);

  T* st = gs + C::oSt;
  for (int i = sub; i < N; i += L) st[C::sx + i] = a.x0[probc * N + i];

  for (int t = 0; t < Th; t++) {
    const T* in = gs + C::oIn + (t % NBUF) * C::IN;
    const T* wsl = in + C::iW - Ge::oK;  // so that wsl[Ge::oX] addresses field X of the slot
    const int oc = t % OC;
    ld.wait(t % NBUF);
    __syncwarp();
    if (NBUF > 1 && t + 1 < Th) issue_loads(t + 1, (t + 1) % NBUF);

    // u = K x + k   (own rows)                                         lqr.py:159
    T x[N];
    lds<T, N, 16>(st + C::sx, x);
#pragma unroll
    for (int c = 0; c < CU; c++) {
      const int b = sub * CU + c;
      if (b < M) {
        T Kr[N];
        lds<T, N, RA_N>(wsl + Ge::oK + b * N, Kr);
        T s = wsl[Ge::ok + b];
#pragma unroll
        for (int j = 0; j < N; j++) s += Kr[j] * x[j];
        st[C::su + b] = s;
        st[C::soU + oc * M + b] = s;
      }
    }
    __syncwarp();
    // x' = A x + B u + d   (own rows)                                  lqr.py:160
    T u[M];
    lds<T, M, 16>(st + C::su, u);
    T xn[CX];
#pragma unroll
    for (int c = 0; c < CX; c++) {
      const int i = sub * CX + c;
      const int ic = i < N ? i : N - 1;
      T Ar[N], Br[M];
      lds<T, N, RA_N>(in + C::iA + ic * N, Ar);
      lds<T, M, RA_M>(in + C::iB + ic * M, Br);
      T s = in[C::id + ic];
#pragma unroll
      for (int j = 0; j < N; j++) s += Ar[j] * x[j];
#pragma unroll
      for (int b = 0; b < M; b++) s += Br[b] * u[b];
      xn[c] = s;
    }
    __syncwarp();  // everyone has read x
#pragma unroll
    for (int c = 0; c < CX; c++) {
      const int i = sub * CX + c;
      if (i < N) {
        st[C::sx + i] = xn[c];
        st[C::soX + oc * N + i] = xn[c];
      }
    }
    __syncwarp();
    // Nu[t] = V_{t+1} x_{t+1} + v_{t+1}   (own rows; V packed symmetric)
    {
      T xf[N], Vr[NP];
      lds<T, N, 16>(st + C::sx, xf);
      lds<T, NP, 16>(wsl + Ge::oVp, Vr);
#pragma unroll
      for (int i = 0; i < N; i++) {
        if (i / CX == sub) {  // row i is mine
          T s = wsl[Ge::ov + i];
#pragma unroll
          for (int j = 0; j < N; j++) s += Vr[triu(i, j, N)] * xf[j];
          st[C::soN + oc * N + i] = s;
        }
      }
    }
    __syncwarp();
    if (NBUF == 1 && t + 1 < Th) issue_loads(t + 1, 0);
    // flush OC buffered steps of X,U,Nu
    if (oc == OC - 1 || t == Th - 1) {
      const int nst = oc + 1, t0 = t - oc;
      if (active) {
        T* Xg = a.X + (prob * Th + t0) * N;
        T* Ug = a.U + (prob * Th + t0) * M;
        T* Ng = a.Nu + (prob * Th + t0) * N;
        for (int e = sub; e < nst * N; e += L) {
          Xg[e] = st[C::soX + e];
          Ng[e] = st[C::soN + e];
        }
        for (int e = sub; e < nst * M; e += L) Ug[e] = st[C::soU + e];
      }
      __syncwarp();
    }
  }
}

}  // namespace idoc
