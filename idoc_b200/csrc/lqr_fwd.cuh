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
// BM (bulk mode): low 3 bits = how many of the large input segments (A, Q, B, M in that order) are moved by
// bulk async copies instead of cp.async; bit 3 = the residual slot is written back by a bulk copy; bit 4 (with bit 3) =
// ONE slot buffer per problem instead of two: the buffer holds [Linv s K k] of step t next to [v V] of step t-1
// (the value function leaving step t), and the end of step t stores the first range to slot t and the second to
// slot t-1 with two bulk copies (slot T-1's [v V] = (qf, sym Qf) is stored before the sweep).
template <typename T, int N, int M, int L, int NBUF, int BM = 0>
struct RiccatiCfg {
  using Ge = Geo<T, N, M>;
  static constexpr int LB = BM & 7, SB = (BM >> 3) & 1, SS = (BM >> 4) & 1;
  // bit 5: the exchange area aliases the in-slab range of Q (dead after the start of phase 3); Q is then bulk-loaded
  // at the END of the step on its own mbarrier and awaited just before phase 3 (saves the exchange area's bytes)
  static constexpr int EXA = (BM >> 5) & 1;
  // bit 6: R, q, r do not pass through shared memory: every lane fetches the entries of its own columns (q: CX, r: CU,
  // R: CU x M scalars) straight from global memory into registers one step ahead (frees their slab range: at n=8, m=4
  // in fp64 that is what lets a 12th warp fit an SM)
  static constexpr int REGQ = (BM >> 6) & 1;
  static_assert(!SS || SB, "the single-buffered slot is written back by bulk copies");
  static_assert(!EXA || (LB >= 2 && NBUF == 1 && SB), "the aliased exchange area needs Q on the bulk path and a single in-slab");
  static constexpr int SZ = Ge::SZ;
  static constexpr int G = 32 / L;
  static constexpr int CX = cdiv(N, L), CU = cdiv(M, L);
  // own columns can be moved as one vector when they never straddle a row end
  static constexpr bool VECX = CX > 1 && N % CX == 0;
  static constexpr bool VECU = CU > 1 && M % CU == 0;
  // in-slab (elements)
  static constexpr int iA = 0;
  static constexpr int iB = iA + Ge::p16(Ge::wA);
  static constexpr int iQ = iB + Ge::p16(Ge::wB);
  static constexpr int iM = iQ + Ge::p16(Ge::wQ);
  static constexpr int iR = iM + Ge::p16(Ge::wM);
  static constexpr int iq = iR + (REGQ ? 0 : Ge::p16(Ge::wR));
  static constexpr int ir = iq + (REGQ ? 0 : Ge::p16(Ge::wq));
  static constexpr int id = ir + (REGQ ? 0 : Ge::p16(Ge::wr));
  static constexpr int IN = id + Ge::p16(Ge::wd);
  // exchange area (elements): the control-block columns every lane of the group needs
  static constexpr int eGuu = 0;
  static constexpr int eGxu = eGuu + Ge::p16(M * M);
  static constexpr int egu = eGxu + Ge::p16(N * M);
  static constexpr int EX = egu + Ge::p16(M);
  // group slab
  static constexpr int oIn = 0;
  static constexpr int oOut = NBUF * IN;
  static constexpr int oEx = EXA ? oIn + iQ : oOut + (SS ? 1 : 2) * Ge::WS;
  static_assert(!EXA || EX <= Ge::p16(Ge::wQ), "exchange area larger than Q");
  // bits 8-10: extra 16-byte units between the slabs of consecutive groups (which banks the G groups of a warp start on)
  static constexpr int PADG = (BM >> 8) & 7;
  static constexpr int GROUP_BYTES = odd16((EXA ? oOut + (SS ? 1 : 2) * Ge::WS : oEx + EX) * SZ) + 16 * PADG;
  static constexpr int BAR_OFF = G * GROUP_BYTES;  // two mbarriers per warp (bulk mode only)
  static constexpr int WARP_BYTES = G * GROUP_BYTES + (BM ? 16 : 0);
  static constexpr int BULK_ELEMS = (LB > 0 ? Ge::wA : 0) + (LB > 1 ? Ge::wQ : 0) + (LB > 2 ? Ge::wB : 0) + (LB > 3 ? Ge::wM : 0);
  static constexpr int LOAD_CHUNKS = (Ge::PER_STEP - BULK_ELEMS - (REGQ ? Ge::wR + Ge::wq + Ge::wr : 0)) * SZ / Ge::GR;
  static constexpr int RL = cdiv(LOAD_CHUNKS, 32);
  static constexpr int RS = cdiv(Ge::WS * SZ / 16, 32);
  static_assert(LB == 0 || (Ge::wA * SZ) % 16 == 0, "bulk copies move multiples of 16 bytes");
  static_assert(LB < 3 || (Ge::wB * SZ) % 16 == 0, "bulk copies move multiples of 16 bytes");
};

// CNT own columns (first one = j0) of row `row` of a row-major matrix with LD columns
template <typename T, int CNT, int LD, bool VEC>
IDOC_DEVICE void lds_cols(const T* mat, int row, int j0, const int* jcl, T* dst) {
  if constexpr (VEC) {
    lds<T, CNT, align_of(cgcd(CNT, LD), (int)sizeof(T))>(mat + row * LD + j0, dst);
  } else {
#pragma unroll
    for (int c = 0; c < CNT; c++) dst[c] = mat[row * LD + jcl[c]];
  }
}
template <typename T, int CNT, int LD, bool VEC>
IDOC_DEVICE void sts_cols(T* mat, int row, int j0, const int* jcl, const bool* valid, const T* src) {
  if constexpr (VEC) {
    if (valid[0]) sts<T, CNT, align_of(cgcd(CNT, LD), (int)sizeof(T))>(mat + row * LD + j0, src);
  } else {
#pragma unroll
    for (int c = 0; c < CNT; c++)
      if (valid[c]) mat[row * LD + jcl[c]] = src[c];
  }
}

// resident CTAs per SM that shared memory allows (1 KB reserved per CTA): passed to __launch_bounds__ so that ptxas keeps
// the register count low enough for the same number of CTAs
__host__ __device__ constexpr int smem_ctas(int warp_bytes, int warps) {
  int c = (228 * 1024) / (warps * warp_bytes + 1024);
  return c > 32 ? 32 : (c < 1 ? 1 : c);
}

template <typename T, int N, int M, int L, int NBUF, int WARPS, int BM = 0>
__global__ void __launch_bounds__(WARPS * 32, smem_ctas(RiccatiCfg<T, N, M, L, NBUF, BM>::WARP_BYTES, WARPS))
    lqr_riccati_kernel(const LqrFwdArgs<T> a) {
  using C = RiccatiCfg<T, N, M, L, NBUF, BM>;
  using Ge = Geo<T, N, M>;
  constexpr int SZ = C::SZ, G = C::G, CX = C::CX, CU = C::CU, NP = Ge::NP, MP = Ge::MP, WS = Ge::WS;
  constexpr int RA_N = row_align(N, SZ), RA_M = row_align(M, SZ);
  extern __shared__ __align__(16) char smem_raw[];

  const int lane = threadIdx.x & 31;
  const int warp = WARPS == 1 ? 0 : __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0);  // warp-uniform for the compiler
  const int grp = lane / L, sub = lane % L;
  const long long prob0 = ((long long)blockIdx.x * WARPS + warp) * G;
  if (prob0 >= a.nb) return;  // warp-uniform; no CTA-wide barriers anywhere below
  char* wbase = smem_raw + warp * C::WARP_BYTES;
  const uint32_t wbase32 = smem_u32(wbase);
  T* gs = reinterpret_cast<T*>(wbase + grp * C::GROUP_BYTES);
  const long long last = a.nb - 1;
  const long long prob = prob0 + grp;
  const bool active = prob <= last;
  const long long probc = active ? prob : last;  // idle groups shadow the last problem, stores masked
  const int Th = a.T_;

  SlabLoader<G, Ge::GR, C::RL, 0, C::LB, C::EXA> ld;
  ld.init(prob0, last, Th, C::GROUP_BYTES, lane, reinterpret_cast<uint64_t*>(wbase + C::BAR_OFF));
  if (C::LB > 0) ld.add_bulk(a.A, Ge::wA * SZ, 0, Ge::wA * SZ, C::iA * SZ); else ld.add(a.A, Ge::wA * SZ, 0, Ge::wA * SZ, C::iA * SZ);
  if (C::LB > 1 && !C::EXA) ld.add_bulk(a.Q, Ge::wQ * SZ, 0, Ge::wQ * SZ, C::iQ * SZ);
  if (C::LB <= 1) ld.add(a.Q, Ge::wQ * SZ, 0, Ge::wQ * SZ, C::iQ * SZ);
  if (C::LB > 2) ld.add_bulk(a.B, Ge::wB * SZ, 0, Ge::wB * SZ, C::iB * SZ); else ld.add(a.B, Ge::wB * SZ, 0, Ge::wB * SZ, C::iB * SZ);
  if (C::LB > 3) ld.add_bulk(a.Mx, Ge::wM * SZ, 0, Ge::wM * SZ, C::iM * SZ); else ld.add(a.Mx, Ge::wM * SZ, 0, Ge::wM * SZ, C::iM * SZ);
  if (C::LB > 1 && C::EXA) ld.add_bulk(a.Q, Ge::wQ * SZ, 0, Ge::wQ * SZ, C::iQ * SZ);  // the late segment comes last
  if (!C::REGQ) {
    ld.add(a.R, Ge::wR * SZ, 0, Ge::wR * SZ, C::iR * SZ);
    ld.add(a.q, Ge::wq * SZ, 0, Ge::wq * SZ, C::iq * SZ);
    ld.add(a.r, Ge::wr * SZ, 0, Ge::wr * SZ, C::ir * SZ);
  }
  ld.add(a.d, Ge::wd * SZ, 0, Ge::wd * SZ, C::id * SZ);
  SlabStorer<G, 16, C::RS> st;
  BulkStorer<G, C::SB ? 1 : 0> bst;
  if (C::SB) {
    bst.init(prob0, last, Th, C::GROUP_BYTES, lane);
    bst.add(a.ws, WS * SZ, 0, WS * SZ, 0);
  } else {
    st.init(prob0, last, Th, C::GROUP_BYTES, lane);
    st.add(a.ws, WS * SZ, 0, WS * SZ, 0);
  }

  // own columns of the state block (jx) and of the control block (ju); out-of-range ones are clamped + masked
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
  const int jx0 = C::VECX ? (vx[0] ? sub * CX : N - CX) : jx[0];
  const int ju0 = C::VECU ? (vu[0] ? sub * CU : M - CU) : ju[0];
  if (C::VECX) {
#pragma unroll
    for (int c = 0; c < CX; c++) jx[c] = jx0 + c;
  }
  if (C::VECU) {
#pragma unroll
    for (int c = 0; c < CU; c++) ju[c] = ju0 + c;
  }

  ld.issue(wbase32, C::oIn * SZ + ((Th - 1) % NBUF) * C::IN * SZ, Th - 1, (Th - 1) % NBUF);
  ld.issue_late(wbase32, C::oIn * SZ, Th - 1);
  // own-column entries of q, r, R of the NEXT step to be processed (REGQ mode), requested one step ahead
  T qreg[CX], rreg[CU], Rreg[CU][M];
  auto fetch_qrR = [&](int t) {
    const long long bt = probc * Th + t;
#pragma unroll
    for (int c = 0; c < CX; c++) qreg[c] = __ldg(a.q + bt * N + jx[c]);
#pragma unroll
    for (int c = 0; c < CU; c++) {
      rreg[c] = __ldg(a.r + bt * M + ju[c]);
#pragma unroll
      for (int b = 0; b < M; b++) Rreg[c][b] = __ldg(a.R + bt * (M * M) + b * M + ju[c]);
    }
  };
  if (C::REGQ) fetch_qrR(Th - 1);

  // V_T = sym(Qf) (packed upper), v_T = qf, into the slab that step T-1 reads
  {
    T* curV = gs + C::oOut + (C::SS ? -Ge::ov : ((Th - 1) & 1) * WS);  // single-slot mode: [v V] first (see below)
    const T* Qf = a.Qf + probc * (N * N);
    const T* qf = a.qf + probc * N;
#pragma unroll
    for (int c = 0; c < CX; c++)
#pragma unroll
      for (int i = 0; i < N; i++)
        if (vx[c] && i <= jx[c])
          curV[Ge::oVp + (i * N - i * (i - 1) / 2 - i) + jx[c]] = T(0.5) * (Qf[i * N + jx[c]] + Qf[jx[c] * N + i]);
    for (int i = sub; i < N; i += L) curV[Ge::ov + i] = qf[i];
  }
  if (C::SS) {  // slot T-1 receives (v_T, V_T)
    fence_proxy_async();
    __syncwarp();
    bst.store_range(wbase32, C::oOut * SZ, Th - 1, Th, Ge::ov * SZ, 0, (WS - Ge::ov) * SZ);
  }
  T dC = T(0), dc = T(0);
  int stat = 0;

  for (int t = Th - 1; t >= 0; t--) {
    const T* in = gs + C::oIn + (t % NBUF) * C::IN;
    // Single-slot mode keeps ONE buffer laid out [v V | Linv s K k]: the first range is the value function leaving
    // step t (slot t-1), the second belongs to slot t, and the two are adjacent in global memory
    // (ws[t-1][ov..WS) is followed by ws[t][0..ov)), so one bulk copy per step stores both.
    T* cur = gs + C::oOut + (C::SS ? WS - Ge::ov : (t & 1) * WS);            // fields Linv, s, K, k of step t
    T* curV = gs + C::oOut + (C::SS ? -Ge::ov : (t & 1) * WS);              // fields v, V entering step t
    T* nxt = gs + C::oOut + (C::SS ? -Ge::ov : ((t & 1) ^ 1) * WS);         // fields v, V leaving step t
    T* ex = gs + C::oEx;

    ld.wait(t % NBUF);
    __syncwarp();
    if (NBUF > 1 && t > 0) ld.issue(wbase32, C::oIn * SZ + ((t - 1) % NBUF) * C::IN * SZ, t - 1, (t - 1) % NBUF);

    // ---- 1. w = v + V d                                          (lqr.py:83-85: Vd, A^T v + A^T Vd)
    T Vr[NP], w[N];
    lds<T, NP, 16>(curV + Ge::oVp, Vr);
    {
      T vv[N], dd[N];
      lds<T, N, 16>(curV + Ge::ov, vv);
      lds<T, N, 16>(in + C::id, dd);
#pragma unroll
      for (int k = 0; k < N; k++) {
        T acc = vv[k];
#pragma unroll
        for (int l = 0; l < N; l++) acc += Vr[triu(k, l, N)] * dd[l];
        w[k] = acc;
      }
    }
    // ---- 2. own columns f of [A B], W = V f, and the linear terms
    T Wx[CX][N], Wu[CU][N], gx[CX], gu_[CU];
    {
      T f[N][CX];
#pragma unroll
      for (int k = 0; k < N; k++) lds_cols<T, CX, N, C::VECX>(in + C::iA, k, jx0, jx, f[k]);
#pragma unroll
      for (int c = 0; c < CX; c++) {
        T acc = C::REGQ ? qreg[c] : in[C::iq + jx[c]];
#pragma unroll
        for (int k = 0; k < N; k++) acc += f[k][c] * w[k];
        gx[c] = acc;  // gx = q + A^T (v + V d)                        lqr.py:84
#pragma unroll
        for (int k = 0; k < N; k++) {
          T s = T(0);
#pragma unroll
          for (int l = 0; l < N; l++) s += Vr[triu(k, l, N)] * f[l][c];
          Wx[c][k] = s;
        }
      }
    }
    {
      T f[N][CU];
#pragma unroll
      for (int k = 0; k < N; k++) lds_cols<T, CU, M, C::VECU>(in + C::iB, k, ju0, ju, f[k]);
#pragma unroll
      for (int c = 0; c < CU; c++) {
        T acc = C::REGQ ? rreg[c] : in[C::ir + ju[c]];
#pragma unroll
        for (int k = 0; k < N; k++) acc += f[k][c] * w[k];
        gu_[c] = acc;  // gu = r + B^T (v + V d)                       lqr.py:85
#pragma unroll
        for (int k = 0; k < N; k++) {
          T s = T(0);
#pragma unroll
          for (int l = 0; l < N; l++) s += Vr[triu(k, l, N)] * f[l][c];
          Wu[c][k] = s;
        }
      }
    }
    // ---- 3. own columns of G = sym([Q M; M^T R]) + [A B]^T V [A B]  lqr.py:80-82
    ld.wait_late();  // Q of this step (aliased-exchange mode)
    T Gx[CX][N], Gxu[CU][N], Guu[CU][M];
#pragma unroll
    for (int i = 0; i < N; i++) {
      T qc[CX];
      lds_cols<T, CX, N, C::VECX>(in + C::iQ, i, jx0, jx, qc);
#pragma unroll
      for (int c = 0; c < CX; c++) Gx[c][i] = qc[c];
      T mc[CU];
      lds_cols<T, CU, M, C::VECU>(in + C::iM, i, ju0, ju, mc);
#pragma unroll
      for (int c = 0; c < CU; c++) Gxu[c][i] = mc[c];
    }
#pragma unroll
    for (int c = 0; c < CX; c++) {  // symmetrise Q: 0.5 (column j + row j)
      T qr[N];
      lds<T, N, RA_N>(in + C::iQ + jx[c] * N, qr);
#pragma unroll
      for (int i = 0; i < N; i++) Gx[c][i] = T(0.5) * (Gx[c][i] + qr[i]);
    }
#pragma unroll
    for (int b = 0; b < M; b++) {
      T rc[CU];
      if (!C::REGQ) lds_cols<T, CU, M, C::VECU>(in + C::iR, b, ju0, ju, rc);
#pragma unroll
      for (int c = 0; c < CU; c++) Guu[c][b] = C::REGQ ? Rreg[c][b] : rc[c];
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
    if (C::EXA) __syncwarp();  // every lane has read Q: its range becomes the exchange area
#pragma unroll
    for (int i = 0; i < N; i++) {
      T tmp[CU];
#pragma unroll
      for (int c = 0; c < CU; c++) tmp[c] = Gxu[c][i];
      sts_cols<T, CU, M, C::VECU>(ex + C::eGxu, i, ju0, ju, vu, tmp);
    }
#pragma unroll
    for (int b = 0; b < M; b++) {
      T tmp[CU];
#pragma unroll
      for (int c = 0; c < CU; c++) tmp[c] = Guu[c][b];
      sts_cols<T, CU, M, C::VECU>(ex + C::eGuu, b, ju0, ju, vu, tmp);
    }
    sts_cols<T, CU, M, C::VECU>(ex + C::egu, 0, ju0, ju, vu, gu_);
    bst.wait_read();  // the bulk store of slot t+1 has left the buffer that phase 6 overwrites
    __syncwarp();
    if (NBUF == 1 && t > 0) ld.issue(wbase32, C::oIn * SZ, t - 1, 0);  // all reads of the in-slab are done
    if (C::REGQ && t > 0) fetch_qrR(t - 1);                             // this step's q, r, R have been consumed

    // ---- 5. every lane: Guu -> (shift) -> inverse Cholesky factor; k   lqr.py:81,86-90
    T Li[MP], guv[M], kk[M];
    T shift = T(0);
    {
      T Gs[MP];
      {
        T Gm[M * M];
        lds<T, M * M, 16>(ex + C::eGuu, Gm);
#pragma unroll
        for (int b = 0; b < M; b++)
#pragma unroll
          for (int c2 = b; c2 < M; c2++) Gs[triu(b, c2, M)] = T(0.5) * (Gm[b * M + c2] + Gm[c2 * M + b]);
      }
      lds<T, M, 16>(ex + C::egu, guv);
      bool clamped = false;
      bool ok = chol_inv<T, M, false>(Gs, Li, T(0), &clamped);
      T fro = T(0);
#pragma unroll
      for (int e = 0; e < MP; e++) fro += Li[e] * Li[e];
      // lambda_min(Guu) >= 1/||Linv||_F^2: if that exceeds EPS no shift is needed (the common case)
      if (!ok || !(fro < T(1.0 / IDOC_EPS))) {
        T Gc[MP];  // private copy: the out-of-line eigen-solver takes an address
#pragma unroll
        for (int e = 0; e < MP; e++) Gc[e] = Gs[e];
        const T mn = min_eig_jacobi<T, M>(Gc);
        if (!(mn == mn) || abs_(mn) > T(1e300)) stat |= ST_NONFINITE;  // NaN / Inf in Guu (the reference propagates NaN)
        shift = T(IDOC_EPS) - mn;
        shift = shift > T(0) ? shift : T(0);  // lqr.py:88
#pragma unroll
        for (int b = 0; b < M; b++) Gc[triu(b, b, M)] += shift;
        chol_inv<T, M, true>(Gc, Li, T(IDOC_EPS) * T(1e-3), &clamped);
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
    // ---- 6. own state columns: K = -Gtuu^{-1} Gux, v_new, V_new      lqr.py:89,91,92
    T Kc[CX][M];
#pragma unroll
    for (int c = 0; c < CX; c++) {
      T grow[M], y[M];
      lds<T, M, RA_M>(ex + C::eGxu + jx[c] * M, grow);
#pragma unroll
      for (int b = 0; b < M; b++) {
        T s = T(0);
#pragma unroll
        for (int c2 = 0; c2 <= b; c2++) s += Li[tril(b, c2)] * grow[c2];
        y[b] = s;
      }
#pragma unroll
      for (int b = 0; b < M; b++) {
        T s = T(0);
#pragma unroll
        for (int c2 = b; c2 < M; c2++) s += Li[tril(c2, b)] * y[c2];
        Kc[c][b] = -s;
      }
    }
    {
      // v = gx + Gxu k + K^T gu + K^T Guu k  ==  gx + K^T (gu - shift k)   (Guu = Gtuu - shift I)
      T vn[CX];
#pragma unroll
      for (int c = 0; c < CX; c++) {
        T s = gx[c];
#pragma unroll
        for (int b = 0; b < M; b++) s += Kc[c][b] * (guv[b] - shift * kk[b]);
        vn[c] = s;
      }
      sts_cols<T, CX, N, C::VECX>(nxt + Ge::ov, 0, jx0, jx, vx, vn);
#pragma unroll
      for (int b = 0; b < M; b++) {
        T tmp[CX];
#pragma unroll
        for (int c = 0; c < CX; c++) tmp[c] = Kc[c][b];
        sts_cols<T, CX, N, C::VECX>(cur + Ge::oK, b, jx0, jx, vx, tmp);
      }
    }
    // V = sym(Gxx + Gxu K + K^T Gux + K^T Guu K) == Gxx + Gxu K - shift K^T K.  Only the upper triangle is formed
    // (entry (i,j), i <= j, by the owner of column j) and written straight into the packed slot of the next step;
    // this equals the reference's symmetrised update up to rounding.
    {
      T Gf[N * M];  // all of Gxu in one batch of loads (row-by-row loads left the FMAs waiting on shared memory)
      lds<T, N * M, 16>(ex + C::eGxu, Gf);
#pragma unroll
      for (int i = 0; i < N; i++) {
#pragma unroll
        for (int c = 0; c < CX; c++) {
          T s = Gx[c][i];
#pragma unroll
          for (int b = 0; b < M; b++) s += Gf[i * M + b] * Kc[c][b];
          if (vx[c] && i <= jx[c]) nxt[Ge::oVp + (i * N - i * (i - 1) / 2 - i) + jx[c]] = s;
        }
      }
    }
    if (sub == 0) {
      sts<T, MP, 16>(cur + Ge::oLinv, Li);
      cur[Ge::oS] = shift;
      sts<T, M, 16>(cur + Ge::ok, kk);
    }
    if (C::SB) fence_proxy_async();
    __syncwarp();
    if (__any_sync(0xffffffffu, shift > T(0))) {  // rare: the -shift K^T K term needs the other lanes' K columns
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
            if (vx[c] && i <= jx[c]) nxt[Ge::oVp + (i * N - i * (i - 1) / 2 - i) + jx[c]] -= shift * s;
          }
      }
      if (C::SS) {  // V_t is stored at the end of this step: publish the correction to the async proxy too
        fence_proxy_async();
        __syncwarp();
      }
    }
    // flush the completed slot t to HBM (coalesced 16-byte streaming stores by the whole warp)
    if (C::EXA && t > 0) ld.issue_late(wbase32, C::oIn * SZ, t - 1);  // the exchange area is dead: refill Q
    if (C::SS) {
      // bytes [0, WS) of the buffer go to ws[t-1][ov..] ++ ws[t][..ov); at t = 0 only the second range exists
      if (t > 0) bst.store_range(wbase32, C::oOut * SZ, t - 1, Th, Ge::ov * SZ, 0, WS * SZ);
      else bst.store_range(wbase32, C::oOut * SZ, 0, Th, 0, (WS - Ge::ov) * SZ, Ge::ov * SZ);
    } else if (C::SB) {
      bst.store(wbase32, (C::oOut + (t & 1) * WS) * SZ, t, Th);
    } else {
      st.store(wbase, (C::oOut + (t & 1) * WS) * SZ, t);
    }
    if (a.K != nullptr && active) {
      for (int e = sub; e < M * N; e += L) a.K[(prob * Th + t) * (M * N) + e] = cur[Ge::oK + e];
      for (int e = sub; e < M; e += L) a.k[(prob * Th + t) * M + e] = cur[Ge::ok + e];
    }
    // (the __syncwarp at the top of the next iteration orders this step's writes before the next reads)
  }
  bst.drain();
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
// BM (bulk mode): low 3 bits = how many of the large input segments (A, ws[K..V], B in that order) are loaded by bulk
// copies; bit 6 = X, U, Nu of a step are stored straight from the owning lanes' registers (no OC-step staging in
// shared memory: the L lanes of a group cover one or two whole 32-byte sectors per store instruction).
// TB = 2 (even horizons): the loads of two consecutive timesteps travel as one request per array (twice the segment size);
// field f of step `par` of the request sits at TB * iF + par * width(f) of the in-slab.
template <typename T, int N, int M, int L, int NBUF, int OC_, int BM = 0, int TB = 1, int PADG = 0>
struct RolloutCfg {
  using Ge = Geo<T, N, M>;
  static constexpr int LB = BM & 7;
  static constexpr bool DS = (BM & 64) != 0;
  static constexpr int OC = DS ? 0 : OC_;
  static constexpr int SZ = Ge::SZ;
  static constexpr int G = 32 / L;
  static constexpr int CX = cdiv(N, L), CU = cdiv(M, L);
  // in-slab: A | B | d | ws[oK .. WS)  (K, k, v, Vp)
  static constexpr int iA = 0;
  static constexpr int iB = iA + Ge::p16(Ge::wA);
  static constexpr int id = iB + Ge::p16(Ge::wB);
  static constexpr int iW = id + Ge::p16(Ge::wd);  // ws sub-range starts here; field f at iW + (Ge::of - Ge::oK)
  static constexpr int WSUB = Ge::WS - Ge::oK;
  static constexpr int IN = TB * (iW + WSUB);
  static_assert(TB == 1 || ((Ge::wA * SZ) % 16 == 0 && (Ge::wB * SZ) % 16 == 0 && (Ge::wd * SZ) % 16 == 0 && (WSUB * SZ) % 16 == 0 &&
                            (Ge::WS * SZ) % 16 == 0),
                "paired timesteps need every per-step segment to be a multiple of 16 bytes");
  // state + out staging: x | u | out chunk of OC steps: X (OC*N) | U (OC*M) | Nu (OC*N), each 16B padded
  static constexpr int sx = 0;
  static constexpr int su = sx + Ge::p16(N);
  static constexpr int soX = su + Ge::p16(M);
  static constexpr int soU = soX + Ge::p16(OC * N);
  static constexpr int soN = soU + Ge::p16(OC * M);
  static constexpr int ST = soN + Ge::p16(OC * N);
  static constexpr int oIn = 0;
  static constexpr int oSt = NBUF * IN;
  // PADG: extra 16-byte units between the slabs of consecutive groups (which banks the G groups of a warp start on)
  static constexpr int GROUP_BYTES = odd16((oSt + ST) * SZ) + 16 * PADG;
  static constexpr int BAR_OFF = G * GROUP_BYTES;
  static constexpr int WARP_BYTES = G * GROUP_BYTES + (LB ? 16 : 0);
  static constexpr int NBULK = (LB > 0 ? 1 : 0) + (LB > 1 ? TB : 0) + (LB > 2 ? 1 : 0);
  static constexpr int BULK_LOAD = (LB > 0 ? Ge::wA : 0) + (LB > 1 ? WSUB : 0) + (LB > 2 ? Ge::wB : 0);
  static constexpr int LOAD_CHUNKS = TB * (Ge::wA + Ge::wB + Ge::wd + WSUB - BULK_LOAD) * SZ / Ge::GR;
  static constexpr int RL = cdiv(LOAD_CHUNKS, 32);
  static_assert(LB == 0 || ((Ge::wA * SZ) % 16 == 0 && (Ge::wB * SZ) % 16 == 0), "bulk copies move multiples of 16 bytes");
};

// CNT consecutive elements owned by this lane, stored with one vector store when they form an aligned 4/8/16-byte unit
template <typename T, int CNT, int LD>
IDOC_DEVICE void stg_own(T* g, const T* v, int i0) {
  constexpr int BYTES = CNT * (int)sizeof(T);
  if constexpr ((BYTES == 16 || BYTES == 8 || BYTES == 4) && LD % CNT == 0) {
    stcs_regs<T, CNT>(g + i0, v);
  } else {
#pragma unroll
    for (int c = 0; c < CNT; c++)
      if (i0 + c < LD) g[i0 + c] = v[c];
  }
}

template <typename T, int N, int M, int L, int NBUF, int OC_, int WARPS, int BM = 0, int TB = 1, int PADG = 0>
__global__ void __launch_bounds__(WARPS * 32) lqr_rollout_kernel(const LqrFwdArgs<T> a) {
  using C = RolloutCfg<T, N, M, L, NBUF, OC_, BM, TB, PADG>;
  constexpr bool DS = C::DS;
  constexpr int OC = DS ? 1 : OC_;
  using Ge = Geo<T, N, M>;
  constexpr int SZ = C::SZ, G = C::G, CX = C::CX, CU = C::CU, NP = Ge::NP, WS = Ge::WS;
  constexpr int RA_N = row_align(N, SZ), RA_M = row_align(M, SZ);
  extern __shared__ __align__(16) char smem_raw[];

  const int lane = threadIdx.x & 31;
  const int warp = WARPS == 1 ? 0 : __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0);  // warp-uniform for the compiler
  const int grp = lane / L, sub = lane % L;
  const long long prob0 = ((long long)blockIdx.x * WARPS + warp) * G;
  if (prob0 >= a.nb) return;
  char* wbase = smem_raw + warp * C::WARP_BYTES;
  const uint32_t wbase32 = smem_u32(wbase);
  T* gs = reinterpret_cast<T*>(wbase + grp * C::GROUP_BYTES);
  const long long last = a.nb - 1;
  const long long prob = prob0 + grp;
  const bool active = prob <= last;
  const long long probc = active ? prob : last;
  const int Th = a.T_;

  const int Tp = Th / TB;  // requests: one per TB timesteps (the launcher guarantees TB | Th)
  SlabLoader<G, Ge::GR, C::RL, 0, C::NBULK> ld;
  ld.init(prob0, last, Tp, C::GROUP_BYTES, lane, reinterpret_cast<uint64_t*>(wbase + C::BAR_OFF));
  auto add_arr = [&](const T* base, int w, int off, bool bulk) {
    if (bulk) ld.add_bulk(base, TB * w * SZ, 0, TB * w * SZ, TB * off * SZ); else ld.add(base, TB * w * SZ, 0, TB * w * SZ, TB * off * SZ);
  };
  add_arr(a.A, Ge::wA, C::iA, C::LB > 0);
#pragma unroll
  for (int par = 0; par < TB; par++) {  // ws[K..V] of each step of the request
    if (C::LB > 1) ld.add_bulk(a.ws, TB * WS * SZ, (par * WS + Ge::oK) * SZ, C::WSUB * SZ, (TB * C::iW + par * C::WSUB) * SZ);
    else ld.add(a.ws, TB * WS * SZ, (par * WS + Ge::oK) * SZ, C::WSUB * SZ, (TB * C::iW + par * C::WSUB) * SZ);
  }
  add_arr(a.B, Ge::wB, C::iB, C::LB > 2);
  add_arr(a.d, Ge::wd, C::id, false);

  auto issue_loads = [&](int tp, int buf) { ld.issue(wbase32, (C::oIn + buf * C::IN) * SZ, tp, buf); };
  issue_loads(0, 0);

  T* st = gs + C::oSt;
  for (int i = sub; i < N; i += L) st[C::sx + i] = a.x0[probc * N + i];

  for (int tp = 0; tp < Tp; tp++) {
    const T* inb = gs + C::oIn + (tp % NBUF) * C::IN;
    ld.wait(tp % NBUF);
    __syncwarp();
    if (NBUF > 1 && tp + 1 < Tp) issue_loads(tp + 1, (tp + 1) % NBUF);
#pragma unroll
   for (int par = 0; par < TB; par++) {
    const int t = tp * TB + par;
    if (par > 0) __syncwarp();
    const T* inA = inb + TB * C::iA + par * Ge::wA;
    const T* inB = inb + TB * C::iB + par * Ge::wB;
    const T* ind = inb + TB * C::id + par * Ge::wd;
    const T* wsl = inb + TB * C::iW + par * C::WSUB - Ge::oK;  // so that wsl[Ge::oX] addresses field X of the slot
    const int oc = t % OC;

    // u = K x + k   (own rows)                                         lqr.py:159
    T x[N], uo[CU];
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
        if constexpr (!DS) st[C::soU + oc * M + b] = s;
        uo[c] = s;
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
      lds<T, N, RA_N>(inA + ic * N, Ar);
      lds<T, M, RA_M>(inB + ic * M, Br);
      T s = ind[ic];
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
        if constexpr (!DS) st[C::soX + oc * N + i] = xn[c];
      }
    }
    __syncwarp();
    // Nu[t] = V_{t+1} x_{t+1} + v_{t+1}   (own rows; V packed symmetric)
    T nuo[CX];
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
          if constexpr (!DS) st[C::soN + oc * N + i] = s;
          nuo[i % CX] = s;
        }
      }
    }
    if constexpr (DS) {
      if (active) {
        const long long bt = prob * Th + t;
        stg_own<T, CX, N>(a.X + bt * N, xn, sub * CX);
        stg_own<T, CX, N>(a.Nu + bt * N, nuo, sub * CX);
        if (sub * CU < M) stg_own<T, CU, M>(a.U + bt * M, uo, sub * CU);
      }
    }
    if (NBUF == 1 || !DS) __syncwarp();
    // flush OC buffered steps of X,U,Nu
    if (!DS && (oc == OC - 1 || t == Th - 1)) {
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
    if (NBUF == 1 && tp + 1 < Tp) issue_loads(tp + 1, 0);
  }
}

}  // namespace idoc
