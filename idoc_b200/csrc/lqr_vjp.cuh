// Implicit-function VJP of the batched LQR solve (reference: idoc/lqr.py:177,
// jaxopt.implicit_diff.custom_root(kkt, solve=solve_cg) around `direct`).
//
// The linear system jaxopt solves by CG,  (d kkt / d s)^T u = -w,  is the KKT system of ANOTHER LQR with the
// same quadratic blocks and linear terms taken from the cotangent w = (Xb, Ub, Nub):
//     x0' = 0,  q'[0] = 0,  q'[t] = Xb[t-1],  qf' = Xb[T-1],  r' = Ub,  d' = Nub          
// so its gains, Cholesky factors and value matrices are those of the forward solve (saved in ws), and only
// the O(n^2)-per-step *vector* recursion has to be redone:
//
//   lqr_adj_bwd_kernel : t = T-1..0   v'_t = gx' + K^T (gu' - s k'),  k'_t = -Gtuu^{-1} gu'
//                        gx' = q' + A^T (v'_{t+1} + V_{t+1} d'),  gu' = r' + B^T (v'_{t+1} + V_{t+1} d')
//   lqr_adj_fwd_kernel : t = 0..T-1   uU = K ux + k',  ux' = A ux + B uU + d',  uNu = V_{t+1} ux' + v'_{t+1}
//                        and the parameter cotangents as outer products, written once, coalesced:
//       gA = Nu (x) sux + uNu (x) sX     gB = Nu (x) uU + uNu (x) U      gd = uNu
//       gQ = sym(sux (x) sX)             gq = sux        gR = sym(uU (x) U)      gr = uU
//       gM = sux (x) U + sX (x) uU       gQf = sym(uX[T-1] (x) X[T-1])   gqf = uX[T-1]
//       gx0 = M[0] uU[0] + A[0]^T uNu[0]
//   (sX[t] = x_t = [x0, X[:-1]],  sux[t] = [0, uX[:-1]];  sym because kkt applies .symm(), lqr.py:130.)
#pragma once
#include "loader.cuh"
#include "lqr_geo.cuh"
#include "small_linalg.cuh"

namespace idoc {

template <typename T>
struct LqrVjpArgs {
  const T *Mx, *A, *B, *x0;         // problem leaves the VJP needs
  const T *X, *U, *Nu;              // forward solution
  const T* ws;                      // forward residual slots
  const T *Xb, *Ub, *Nub;           // cotangents (Nub may be null)
  T *gQ, *gq, *gQf, *gqf, *gM, *gR, *gr, *gA, *gB, *gd, *gx0;
  T* ws2;                           // [B,T,WS2]
  long long nb;
  int T_;
  int reduce_batch;  // != 0: "vectors only" mode for the batch-summed gradients (grad_reduce.cuh): the sweep stores
                     // sux -> gq, uU -> gr, uNu -> gd (per problem, [B,T,.]) and uX[T-1] -> gqf ([B,n]); no outer products
};

// ======================================================================================= adjoint backward
// BM (bulk mode): how many of the large input segments (A, ws[Linv..K], B in that order) are loaded by bulk copies.
// TB = 2 (even horizons): the loads of two consecutive timesteps travel as one request per array (twice the segment size);
// field f of step `par` of the request sits at TB * iF + par * width(f) of the in-slab.
template <typename T, int N, int M, int L, int NBUF, bool HAS_NUB, int BM = 0, int TB = 1, int PADG = 0>
struct AdjBwdCfg {
  using Ge = Geo<T, N, M>;
  static constexpr int LB = BM & 7;
  static constexpr int SZ = Ge::SZ;
  static constexpr int G = 32 / L;
  static constexpr int CX = cdiv(N, L), CU = cdiv(M, L);
  static constexpr int WPRE = Ge::ok;  // ws prefix [0, ok): Linv, s, K
  static constexpr int iA = 0;
  static constexpr int iB = iA + Ge::p16(Ge::wA);
  static constexpr int iW = iB + Ge::p16(Ge::wB);           // ws prefix
  static constexpr int iV = iW + WPRE;                      // Vp (only if HAS_NUB)
  static constexpr int iUb = iV + (HAS_NUB ? Ge::p16(Ge::NP) : 0);
  static constexpr int iNb = iUb + Ge::p16(M);              // Nub[t] (only if HAS_NUB)
  static constexpr int iXb = iNb + (HAS_NUB ? Ge::p16(N) : 0);  // Xb[t-1]
  static constexpr int IN = TB * (iXb + Ge::p16(N));
  static constexpr int WV = Ge::p16(Ge::NP);
  static_assert(TB == 1 || ((Ge::wA * SZ) % 16 == 0 && (Ge::wB * SZ) % 16 == 0 && (WPRE * SZ) % 16 == 0 && (M * SZ) % 16 == 0 &&
                            (N * SZ) % 16 == 0 && (Ge::WS * SZ) % 16 == 0),
                "paired timesteps need every per-step segment to be a multiple of 16 bytes");
  static constexpr int oIn = 0;
  static constexpr int oOut = NBUF * IN;         // 2 x WS2 slots
  static constexpr int oGu = oOut + 2 * Ge::WS2;  // exchange: gu' (M)
  static constexpr int GROUP_BYTES = odd16((oGu + Ge::p16(M)) * SZ) + 16 * PADG;  // PADG: see RolloutCfg
  static constexpr int BAR_OFF = G * GROUP_BYTES;
  static constexpr int WARP_BYTES = G * GROUP_BYTES + (LB ? 16 : 0);
  static constexpr int NBULK = (LB > 0 ? 1 : 0) + (LB > 1 ? TB : 0) + (LB > 2 ? 1 : 0);
  static constexpr int BULK_LOAD = (LB > 0 ? Ge::wA : 0) + (LB > 1 ? WPRE : 0) + (LB > 2 ? Ge::wB : 0);
  static_assert(LB == 0 || ((Ge::wA * SZ) % 16 == 0 && (Ge::wB * SZ) % 16 == 0), "bulk copies move multiples of 16 bytes");
  static constexpr int LOAD_CHUNKS =
      TB * ((Ge::wA + Ge::wB + M + (HAS_NUB ? N : 0)) * SZ + (WPRE + (HAS_NUB ? Ge::p16(Ge::NP) : 0)) * SZ - BULK_LOAD * SZ) / Ge::GR;
  static constexpr int RL = cdiv(LOAD_CHUNKS, 32);
  static constexpr int RQ = cdiv(TB * N * SZ / Ge::GR, 32);
  static constexpr int RS = cdiv(Ge::WS2 * SZ / 16, 32);
};

template <typename T, int N, int M, int L, int NBUF, bool HAS_NUB, int WARPS, int BM = 0, int TB = 1, int PADG = 0>
__global__ void __launch_bounds__(WARPS * 32) lqr_adj_bwd_kernel(const LqrVjpArgs<T> a) {
  using C = AdjBwdCfg<T, N, M, L, NBUF, HAS_NUB, BM, TB, PADG>;
  using Ge = Geo<T, N, M>;
  constexpr int SZ = C::SZ, G = C::G, CX = C::CX, CU = C::CU, NP = Ge::NP, MP = Ge::MP, WS = Ge::WS, WS2 = Ge::WS2;
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
  SlabLoader<G, Ge::GR, C::RL, C::RQ, C::NBULK> ld;
  ld.init(prob0, last, Tp, C::GROUP_BYTES, lane, reinterpret_cast<uint64_t*>(wbase + C::BAR_OFF));
  auto add_arr = [&](const T* base, int w, int off, bool bulk) {
    if (bulk) ld.add_bulk(base, TB * w * SZ, 0, TB * w * SZ, TB * off * SZ); else ld.add(base, TB * w * SZ, 0, TB * w * SZ, TB * off * SZ);
  };
  add_arr(a.A, Ge::wA, C::iA, C::LB > 0);
#pragma unroll
  for (int par = 0; par < TB; par++) {  // ws[Linv..K] of each step of the request
    if (C::LB > 1) ld.add_bulk(a.ws, TB * WS * SZ, par * WS * SZ, C::WPRE * SZ, (TB * C::iW + par * C::WPRE) * SZ);
    else ld.add(a.ws, TB * WS * SZ, par * WS * SZ, C::WPRE * SZ, (TB * C::iW + par * C::WPRE) * SZ);
  }
  add_arr(a.B, Ge::wB, C::iB, C::LB > 2);
  if constexpr (HAS_NUB) {
#pragma unroll
    for (int par = 0; par < TB; par++) ld.add(a.ws, TB * WS * SZ, (par * WS + Ge::oVp) * SZ, C::WV * SZ, (TB * C::iV + par * C::WV) * SZ);
  }
  add_arr(a.Ub, M, C::iUb, false);
  if constexpr (HAS_NUB) add_arr(a.Nub, N, C::iNb, false);
  // q'[t] = Xb[t-1].  TB = 1: the segment of step t-1, not loaded at t = 0.  TB = 2: request tp needs Xb[2tp-1], Xb[2tp]: the
  // array seen from Xb[1] on, request tp-1 (not loaded at tp = 0, whose second step takes Xb[0] from registers)
  ld.add(a.Xb + (TB - 1) * N, TB * N * SZ, 0, TB * N * SZ, TB * C::iXb * SZ, -1);
  SlabStorer<G, 16, C::RS> st;
  st.init(prob0, last, Th, C::GROUP_BYTES, lane);
  st.add(a.ws2, WS2 * SZ, 0, WS2 * SZ, 0);

  auto issue_loads = [&](int t, int buf) { ld.issue(wbase32, (C::oIn + buf * C::IN) * SZ, t, buf); };
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

  issue_loads(Tp - 1, (Tp - 1) % NBUF);
  {  // v'_T = qf' = Xb[T-1]
    T* cur = gs + C::oOut + ((Th - 1) & 1) * WS2;
    for (int i = sub; i < N; i += L) cur[Ge::o2v + i] = a.Xb[(probc * Th + Th - 1) * N + i];
  }
  T xb0[CX];  // Xb[0] (TB = 2 only: q' of step 1)
  if constexpr (TB > 1) {
#pragma unroll
    for (int c = 0; c < CX; c++) xb0[c] = a.Xb[(probc * Th) * N + jx[c]];
  }

  for (int tp = Tp - 1; tp >= 0; tp--) {
    const T* inb = gs + C::oIn + (tp % NBUF) * C::IN;
    ld.wait(tp % NBUF);
    __syncwarp();
    if (NBUF > 1 && tp > 0) issue_loads(tp - 1, (tp - 1) % NBUF);
#pragma unroll
   for (int par = TB - 1; par >= 0; par--) {
    const int t = tp * TB + par;
    if (par < TB - 1) __syncwarp();  // the previous step's store has read the slot this step writes
    const T* inA = inb + TB * C::iA + par * Ge::wA;
    const T* inB = inb + TB * C::iB + par * Ge::wB;
    const T* wsl = inb + TB * C::iW + par * C::WPRE;  // prefix: field offsets as in Geo
    const T* inUb = inb + TB * C::iUb + par * M;
    const T* inXb = inb + TB * C::iXb + par * N;
    T* cur = gs + C::oOut + (t & 1) * WS2;
    T* nxt = gs + C::oOut + ((t & 1) ^ 1) * WS2;
    T* exg = gs + C::oGu;

    T w[N];
    lds<T, N, 16>(cur + Ge::o2v, w);
    if constexpr (HAS_NUB) {
      T Vr[NP], dd[N];
      lds<T, NP, 16>(inb + TB * C::iV + par * C::WV, Vr);
      lds<T, N, 16>(inb + TB * C::iNb + par * N, dd);
#pragma unroll
      for (int k = 0; k < N; k++) {
        T acc = w[k];
#pragma unroll
        for (int l = 0; l < N; l++) acc += Vr[triu(k, l, N)] * dd[l];
        w[k] = acc;
      }
    }
    T gx[CX];
#pragma unroll
    for (int c = 0; c < CX; c++) {
      T acc = t > 0 ? ((TB > 1 && t == 1) ? xb0[c] : inXb[jx[c]]) : T(0);  // q'[t] = Xb[t-1], q'[0] = 0
#pragma unroll
      for (int k = 0; k < N; k++) acc += inA[k * N + jx[c]] * w[k];
      gx[c] = acc;
    }
#pragma unroll
    for (int c = 0; c < CU; c++) {
      T acc = inUb[ju[c]];
#pragma unroll
      for (int k = 0; k < N; k++) acc += inB[k * M + ju[c]] * w[k];
      if (vu[c]) exg[ju[c]] = acc;
    }
    __syncwarp();
    T guv[M], kk[M], Li[MP];
    lds<T, M, 16>(exg, guv);
    lds<T, MP, 16>(wsl + Ge::oLinv, Li);
    const T shift = wsl[Ge::oS];
    {
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
        kk[b] = -s;
      }
    }
#pragma unroll
    for (int c = 0; c < CX; c++) {
      T vn = gx[c];
#pragma unroll
      for (int b = 0; b < M; b++) vn += wsl[Ge::oK + b * N + jx[c]] * (guv[b] - shift * kk[b]);
      if (vx[c]) nxt[Ge::o2v + jx[c]] = vn;
    }
    if (sub == 0) sts<T, M, 16>(cur + Ge::o2k, kk);
    __syncwarp();
    if (NBUF == 1 && par == 0 && tp > 0) issue_loads(tp - 1, 0);
    st.store(wbase, (C::oOut + (t & 1) * WS2) * SZ, t);
   }
  }
}

// ======================================================================================= adjoint forward + grads
// BM (bulk mode): low 3 bits = how many of the large input segments (A, V, K, B in that order) are loaded by bulk
// async copies; bits 3-5 = how many of the large gradient segments (gA, gQ, gB, gM) are stored by bulk copies.
template <typename T, int N, int M, int L, int NBUF, bool HAS_NUB, int BM = 0>
struct AdjFwdCfg {
  using Ge = Geo<T, N, M>;
  static constexpr int LB = BM & 7, SB = (BM >> 3) & 7;
  static constexpr int SZ = Ge::SZ;
  static constexpr int G = 32 / L;
  static constexpr int CX = cdiv(N, L), CU = cdiv(M, L);
  static constexpr int WK = Ge::ok - Ge::oK;  // K range
  static constexpr int iA = 0;
  static constexpr int iB = iA + Ge::p16(Ge::wA);
  static constexpr int iK = iB + Ge::p16(Ge::wB);
  static constexpr int iV = iK + WK;
  static constexpr int iW2 = iV + Ge::p16(Ge::NP);
  static constexpr int iU = iW2 + Ge::WS2;
  static constexpr int iNu = iU + Ge::p16(M);
  static constexpr int iNb = iNu + Ge::p16(N);
  static constexpr int iSX = iNb + (HAS_NUB ? Ge::p16(N) : 0);
  static constexpr int IN = iSX + Ge::p16(N);
  // out-slab: gradients of step t, same order as the copy plan
  static constexpr int gA = 0;
  static constexpr int gB = gA + Ge::p16(Ge::wA);
  static constexpr int gQ = gB + Ge::p16(Ge::wB);
  static constexpr int gM = gQ + Ge::p16(Ge::wQ);
  static constexpr int gR = gM + Ge::p16(Ge::wM);
  static constexpr int gq = gR + Ge::p16(Ge::wR);
  static constexpr int gr = gq + Ge::p16(Ge::wq);
  static constexpr int gd = gr + Ge::p16(Ge::wr);
  static constexpr int OUT = gd + Ge::p16(Ge::wd);
  // state: ux | uU | uxn
  static constexpr int sux = 0;
  static constexpr int suU = sux + Ge::p16(N);
  static constexpr int sxn = suU + Ge::p16(M);
  static constexpr int ST = sxn + Ge::p16(N);
  static constexpr int oIn = 0;
  static constexpr int oOut = NBUF * IN;
  static constexpr int oSt = oOut + OUT;
  static constexpr int GROUP_BYTES = odd16((oSt + ST) * SZ);
  static constexpr int BAR_OFF = G * GROUP_BYTES;
  static constexpr int WARP_BYTES = G * GROUP_BYTES + (LB ? 16 : 0);
  static constexpr int BULK_LOAD = (LB > 0 ? Ge::wA : 0) + (LB > 1 ? Ge::p16(Ge::NP) : 0) + (LB > 2 ? WK : 0) + (LB > 3 ? Ge::wB : 0);
  static constexpr int BULK_STORE = (SB > 0 ? Ge::wA : 0) + (SB > 1 ? Ge::wQ : 0) + (SB > 2 ? Ge::wB : 0) + (SB > 3 ? Ge::wM : 0);
  static constexpr int LOAD_CHUNKS =
      ((Ge::wA + Ge::wB + M + N + (HAS_NUB ? N : 0)) * SZ + (WK + Ge::p16(Ge::NP) + Ge::WS2) * SZ - BULK_LOAD * SZ) / Ge::GR;
  static constexpr int RL = cdiv(LOAD_CHUNKS, 32);
  static constexpr int RQ = cdiv(N * SZ / Ge::GR, 32);
  static constexpr int RS = cdiv((Ge::PER_STEP - BULK_STORE) * SZ / Ge::GR, 32);
  static_assert(BM == 0 || ((Ge::wA * SZ) % 16 == 0 && (Ge::wB * SZ) % 16 == 0 && (WK * SZ) % 16 == 0), "bulk copies move multiples of 16 bytes");
};

template <typename T, int N, int M, int L, int NBUF, bool HAS_NUB, int WARPS, int BM = 0>
__global__ void __launch_bounds__(WARPS * 32) lqr_adj_fwd_kernel(const LqrVjpArgs<T> a) {
  using C = AdjFwdCfg<T, N, M, L, NBUF, HAS_NUB, BM>;
  using Ge = Geo<T, N, M>;
  constexpr int SZ = C::SZ, G = C::G, CX = C::CX, CU = C::CU, NP = Ge::NP, WS = Ge::WS, WS2 = Ge::WS2;
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
  const bool vo = a.reduce_batch != 0;  // vectors only (see LqrVjpArgs)

  SlabLoader<G, Ge::GR, C::RL, C::RQ, C::LB> ld;
  ld.init(prob0, last, Th, C::GROUP_BYTES, lane, reinterpret_cast<uint64_t*>(wbase + C::BAR_OFF));
  if (C::LB > 0) ld.add_bulk(a.A, Ge::wA * SZ, 0, Ge::wA * SZ, C::iA * SZ); else ld.add(a.A, Ge::wA * SZ, 0, Ge::wA * SZ, C::iA * SZ);
  if (C::LB > 1) ld.add_bulk(a.ws, WS * SZ, Ge::oVp * SZ, Ge::p16(NP) * SZ, C::iV * SZ); else ld.add(a.ws, WS * SZ, Ge::oVp * SZ, Ge::p16(NP) * SZ, C::iV * SZ);
  if (C::LB > 2) ld.add_bulk(a.ws, WS * SZ, Ge::oK * SZ, C::WK * SZ, C::iK * SZ); else ld.add(a.ws, WS * SZ, Ge::oK * SZ, C::WK * SZ, C::iK * SZ);
  if (C::LB > 3) ld.add_bulk(a.B, Ge::wB * SZ, 0, Ge::wB * SZ, C::iB * SZ); else ld.add(a.B, Ge::wB * SZ, 0, Ge::wB * SZ, C::iB * SZ);
  ld.add(a.ws2, WS2 * SZ, 0, WS2 * SZ, C::iW2 * SZ);
  ld.add(a.U, M * SZ, 0, M * SZ, C::iU * SZ);
  ld.add(a.Nu, N * SZ, 0, N * SZ, C::iNu * SZ);
  if constexpr (HAS_NUB) ld.add(a.Nub, N * SZ, 0, N * SZ, C::iNb * SZ);
  ld.add(a.X, N * SZ, 0, N * SZ, C::iSX * SZ, -1);  // sX[t] = X[t-1]; x0 at t = 0
  SlabStorer<G, Ge::GR, C::RS> st;
  st.init(prob0, last, Th, C::GROUP_BYTES, lane);
  BulkStorer<G, C::SB> bst;
  bst.init(prob0, last, Th, C::GROUP_BYTES, lane);
  if (!vo) {
    if (C::SB > 0) bst.add(a.gA, Ge::wA * SZ, 0, Ge::wA * SZ, C::gA * SZ); else st.add(a.gA, Ge::wA * SZ, 0, Ge::wA * SZ, C::gA * SZ);
    if (C::SB > 1) bst.add(a.gQ, Ge::wQ * SZ, 0, Ge::wQ * SZ, C::gQ * SZ); else st.add(a.gQ, Ge::wQ * SZ, 0, Ge::wQ * SZ, C::gQ * SZ);
    if (C::SB > 2) bst.add(a.gB, Ge::wB * SZ, 0, Ge::wB * SZ, C::gB * SZ); else st.add(a.gB, Ge::wB * SZ, 0, Ge::wB * SZ, C::gB * SZ);
    if (C::SB > 3) bst.add(a.gM, Ge::wM * SZ, 0, Ge::wM * SZ, C::gM * SZ); else st.add(a.gM, Ge::wM * SZ, 0, Ge::wM * SZ, C::gM * SZ);
    st.add(a.gR, Ge::wR * SZ, 0, Ge::wR * SZ, C::gR * SZ);
  }
  st.add(a.gq, Ge::wq * SZ, 0, Ge::wq * SZ, C::gq * SZ);
  st.add(a.gr, Ge::wr * SZ, 0, Ge::wr * SZ, C::gr * SZ);
  st.add(a.gd, Ge::wd * SZ, 0, Ge::wd * SZ, C::gd * SZ);

  auto issue_loads = [&](int t, int buf) { ld.issue(wbase32, (C::oIn + buf * C::IN) * SZ, t, buf); };
  issue_loads(0, 0);
  T* stt = gs + C::oSt;
  T* out = gs + C::oOut;
  for (int i = sub; i < N; i += L) {
    stt[C::sux + i] = T(0);                                 // ux_0 = x0' = 0
    gs[C::oIn + C::iSX + i] = a.x0[probc * N + i];          // sX[0] = x0 (buffer 0 is read at t = 0)
  }

  for (int t = 0; t < Th; t++) {
    const T* in = gs + C::oIn + (t % NBUF) * C::IN;
    ld.wait(t % NBUF);
    __syncwarp();
    if (NBUF > 1 && t + 1 < Th) issue_loads(t + 1, (t + 1) % NBUF);

    T ux[N], sX[N], Uv[M];
    lds<T, N, 16>(stt + C::sux, ux);
    lds<T, N, 16>(in + C::iSX, sX);
    lds<T, M, 16>(in + C::iU, Uv);
    // uU = K ux + k'   (own rows)
#pragma unroll
    for (int c = 0; c < CU; c++) {
      const int b = sub * CU + c;
      if (b < M) {
        T Kr[N];
        lds<T, N, RA_N>(in + C::iK + b * N, Kr);
        T s = in[C::iW2 + Ge::o2k + b];
#pragma unroll
        for (int j = 0; j < N; j++) s += Kr[j] * ux[j];
        stt[C::suU + b] = s;
      }
    }
    __syncwarp();
    T uU[M];
    lds<T, M, 16>(stt + C::suU, uU);
    // ux' = A ux + B uU + d'   (own rows)
#pragma unroll
    for (int c = 0; c < CX; c++) {
      const int i = sub * CX + c;
      if (i < N) {
        T Ar[N], Br[M];
        lds<T, N, RA_N>(in + C::iA + i * N, Ar);
        lds<T, M, RA_M>(in + C::iB + i * M, Br);
        T s = HAS_NUB ? in[C::iNb + i] : T(0);
#pragma unroll
        for (int j = 0; j < N; j++) s += Ar[j] * ux[j];
#pragma unroll
        for (int b = 0; b < M; b++) s += Br[b] * uU[b];
        stt[C::sxn + i] = s;
      }
    }
    bst.wait_read();  // the bulk stores of step t-1 have left the out-slab
    __syncwarp();
    // uNu = V_{t+1} ux' + v'_{t+1}  (own rows), and the gradient rows this lane owns
    {
      T xn[N], Vr[NP];
      lds<T, N, 16>(stt + C::sxn, xn);
      lds<T, NP, 16>(in + C::iV, Vr);
#pragma unroll
      for (int i = 0; i < N; i++) {
        if (i / CX == sub) {
          T unu = in[C::iW2 + Ge::o2v + i];
#pragma unroll
          for (int j = 0; j < N; j++) unu += Vr[triu(i, j, N)] * xn[j];
          const T nu = in[C::iNu + i];
          if (!vo) {
            T rowA[N], rowQ[N], rowB[M], rowM[M];
#pragma unroll
            for (int j = 0; j < N; j++) {
              rowA[j] = nu * ux[j] + unu * sX[j];
              rowQ[j] = T(0.5) * (ux[i] * sX[j] + sX[i] * ux[j]);
            }
#pragma unroll
            for (int b = 0; b < M; b++) {
              rowB[b] = nu * uU[b] + unu * Uv[b];
              rowM[b] = ux[i] * Uv[b] + sX[i] * uU[b];
            }
            sts<T, N, RA_N>(out + C::gA + i * N, rowA);
            sts<T, N, RA_N>(out + C::gQ + i * N, rowQ);
            sts<T, M, RA_M>(out + C::gB + i * M, rowB);
            sts<T, M, RA_M>(out + C::gM + i * M, rowM);
          }
          out[C::gd + i] = unu;
          out[C::gq + i] = ux[i];
          stt[C::sux + i] = xn[i];  // becomes sux of the next step (all lanes have read the old one)
        }
      }
#pragma unroll
      for (int b = 0; b < M; b++) {
        if (b / CU == sub) {
          if (!vo) {
            T rowR[M];
#pragma unroll
            for (int c2 = 0; c2 < M; c2++) rowR[c2] = T(0.5) * (uU[b] * Uv[c2] + Uv[b] * uU[c2]);
            sts<T, M, RA_M>(out + C::gR + b * M, rowR);
          }
          out[C::gr + b] = uU[b];
        }
      }
      if (C::SB) fence_proxy_async();
      __syncwarp();
      if (NBUF == 1 && t + 1 < Th) issue_loads(t + 1, 0);
      if (t == 0 && active) {
        // gx0 = M[0] uU[0] + A[0]^T uNu[0]          (uNu[0] = the gd row just published)
        const T* M0 = a.Mx + (prob * Th) * (N * M);
        const T* A0 = a.A + (prob * Th) * (N * N);
        for (int i = sub; i < N; i += L) {
          T s = T(0);
#pragma unroll
          for (int b = 0; b < M; b++) s += M0[i * M + b] * uU[b];
#pragma unroll
          for (int k = 0; k < N; k++) s += A0[k * N + i] * out[C::gd + k];
          a.gx0[prob * N + i] = s;
        }
      }
      if (t == Th - 1 && active) {
        // gQf = sym(uX[T-1] (x) X[T-1]),  gqf = uX[T-1]
        const T* XT = a.X + (prob * Th + Th - 1) * N;
        T* gQf = a.gQf + prob * (N * N);
        T* gqf = a.gqf + prob * N;
        if (!vo) {
          for (int e = sub; e < N * N; e += L) {
            const int i = e / N, j = e % N;
            gQf[e] = T(0.5) * (stt[C::sxn + i] * XT[j] + stt[C::sxn + j] * XT[i]);
          }
        }
        for (int e = sub; e < N; e += L) gqf[e] = stt[C::sxn + e];
      }
    }
    // write the gradient slab of step t (vectors-only mode: gq, gr, gd alone)
    if (!vo) bst.store(wbase32, C::oOut * SZ, t, Th);
    st.store(wbase, C::oOut * SZ, t);
    __syncwarp();
  }
  bst.drain();
}

// ======================================================================================= adjoint forward, direct stores
// Same sweep as lqr_adj_fwd_kernel, but the gradient slab is never staged in shared memory: every gradient leaf of a
// step is an outer-product form  s * (a1 (x) b1 + a2 (x) b2)  of vectors that all lanes of the group can read
// (Nu, uNu, sux, sX, uU, U), so lane `sub` computes the 16-byte chunks  sub, sub + L, sub + 2L, ...  of each leaf
// straight into registers and stores them with streaming 16-byte stores (the L lanes of a group cover 16 L contiguous
// bytes per instruction: whole 32-byte sectors).  That removes the 228-scalar out-slab (1.8 KB of 5.1 KB per problem
// in fp64: 5 -> 8 resident warps per SM), its STS + LDS traffic and one pass through the LSU data pipe.
// Requires n and m to be multiples of the 16-byte chunk (2 doubles / 4 floats).  In vectors-only mode (batch-summed
// gradients, grad_reduce.cuh) the same sweep stores gq, gr, gd, gqf and gx0 and skips every outer product.
template <typename T, int N, int M>
constexpr bool adj_fwd_direct_ok() {
  return N % (16 / (int)sizeof(T)) == 0 && M % (16 / (int)sizeof(T)) == 0;
}

// TB = 2: the loads of TWO consecutive timesteps travel as one request per array (every [B][T][w] array is contiguous in
// t, so the pair is one segment of twice the size: in fp32 the 256-byte A becomes 512 bytes, the 16-byte U 32 bytes).
// The in-slab holds field f of the pair at TB * iF + par * width(f); the horizon must be even (odd horizons take TB = 1);
// X[t-1] of the pair's first step is carried over from the previous pair in the state area.
template <typename T, int N, int M, int L, int NBUF, bool HAS_NUB, int BM = 0, int TB = 1, int PADG = 0>
struct AdjFwdDCfg {
  using Ge = Geo<T, N, M>;
  using Base = AdjFwdCfg<T, N, M, L, NBUF, HAS_NUB, 0>;
  static constexpr int LB = BM & 7;
  static constexpr int OCG = (BM >> 3) & 7;  // the small leaves (gR, gq, gr, gd) are staged for OCG steps and flushed as
                                             // OCG-step contiguous blocks (0: stored every step like the large ones)
  static constexpr int SZ = Ge::SZ;
  static constexpr int G = 32 / L;
  static constexpr int CX = cdiv(N, L), CU = cdiv(M, L);
  static constexpr int WK = Base::WK;
  static constexpr int WV = Ge::p16(Ge::NP);
  static constexpr int iA = Base::iA, iB = Base::iB, iK = Base::iK, iV = Base::iV, iW2 = Base::iW2, iU = Base::iU, iNu = Base::iNu,
                       iNb = Base::iNb, iSX = Base::iSX, IN = TB * Base::IN;
  static_assert(TB == 1 || ((Ge::wA * SZ) % 16 == 0 && (Ge::wB * SZ) % 16 == 0 && (WK * SZ) % 16 == 0 && (Ge::WS2 * SZ) % 16 == 0 &&
                            (M * SZ) % 16 == 0 && (N * SZ) % 16 == 0 && (Ge::WS * SZ) % 16 == 0),
                "paired timesteps need every per-step segment to be a multiple of 16 bytes");
  // state: ux (two copies, ping-pong over t) | uU | uNu | X carried over from the previous pair (TB = 2)
  static constexpr int sux = 0;
  static constexpr int suU = sux + 2 * Ge::p16(N);
  static constexpr int sun = suU + Ge::p16(M);
  static constexpr int sXp = sun + Ge::p16(N);
  static constexpr int sgR = sXp + (TB > 1 ? Ge::p16(N) : 0);     // staging: OCG steps of gR | gq | gr | gd
  static constexpr int sgq = sgR + OCG * M * M;
  static constexpr int sgr = sgq + OCG * N;
  static constexpr int sgd = sgr + OCG * M;
  static constexpr int ST = sgd + OCG * N;
  static constexpr int oIn = 0;
  static constexpr int oSt = NBUF * IN;
  static constexpr int GROUP_BYTES = odd16((oSt + ST) * SZ) + 16 * PADG;  // PADG: see RolloutCfg
  static constexpr int BAR_OFF = G * GROUP_BYTES;
  static constexpr int WARP_BYTES = G * GROUP_BYTES + (LB ? 16 : 0);
  static constexpr int NBULK = (LB > 0 ? 1 : 0) + (LB > 1 ? TB : 0) + (LB > 2 ? TB : 0) + (LB > 3 ? 1 : 0);
  static constexpr int BULK_LOAD = (LB > 0 ? Ge::wA : 0) + (LB > 1 ? WV : 0) + (LB > 2 ? WK : 0) + (LB > 3 ? Ge::wB : 0);
  static constexpr int MAIN_LOAD = Ge::wA + Ge::wB + WK + WV + Ge::WS2 + M + N + (HAS_NUB ? N : 0) + (TB > 1 ? N : 0);
  static constexpr int LOAD_CHUNKS = TB * (MAIN_LOAD - BULK_LOAD) * SZ / Ge::GR;
  static constexpr int RL = cdiv(LOAD_CHUNKS, 32);
  static constexpr int RQ = TB > 1 ? 0 : cdiv(N * SZ / Ge::GR, 32);
};

// g[row][col] = s * (a1[row] b1[col] + a2[row] b2[col]) for the 16-byte chunks sub, sub+L, ... of a row-major NR x NC leaf
// (a*, b* point into shared memory; NC is a multiple of the chunk)
template <typename T, int E, bool STREAM>
IDOC_DEVICE void put_chunk(T* dst, const T* v) {
  if constexpr (STREAM) stcs_regs<T, E>(dst, v); else sts<T, E, 16>(dst, v);
}
template <typename T, int NR, int NC, int L, bool HALF, bool STREAM = true>
IDOC_DEVICE void store_outer2(T* g, int sub, const T* a1, const T* b1, const T* a2, const T* b2) {
  constexpr int E = 16 / (int)sizeof(T), RC = NC / E, NCH = NR * RC, KMAX = cdiv(NCH, L);
  static_assert(NC % E == 0, "row length must be a multiple of the 16-byte chunk");
  if constexpr (L % RC == 0) {
    // the lane's column chunk is the same for every k; its rows are k * (L / RC) + sub / RC
    const int cp = sub % RC, r0 = sub / RC;
    T c1[E], c2[E];
    lds<T, E, 16>(b1 + cp * E, c1);
    lds<T, E, 16>(b2 + cp * E, c2);
#pragma unroll
    for (int k = 0; k < KMAX; k++) {
      const int row = r0 + k * (L / RC);
      if (NCH % L == 0 || row < NR) {
        const T x1 = a1[row], x2 = a2[row];
        T v[E];
#pragma unroll
        for (int e = 0; e < E; e++) v[e] = HALF ? T(0.5) * (x1 * c1[e] + x2 * c2[e]) : x1 * c1[e] + x2 * c2[e];
        put_chunk<T, E, STREAM>(g + (row * RC + cp) * E, v);
      }
    }
  } else {
#pragma unroll
    for (int k = 0; k < KMAX; k++) {
      const int c = sub + L * k;
      if (NCH % L == 0 || c < NCH) {
        const int row = c / RC, cp = c % RC;
        T c1[E], c2[E];
        lds<T, E, 16>(b1 + cp * E, c1);
        lds<T, E, 16>(b2 + cp * E, c2);
        const T x1 = a1[row], x2 = a2[row];
        T v[E];
#pragma unroll
        for (int e = 0; e < E; e++) v[e] = HALF ? T(0.5) * (x1 * c1[e] + x2 * c2[e]) : x1 * c1[e] + x2 * c2[e];
        put_chunk<T, E, STREAM>(g + c * E, v);
      }
    }
  }
}
// g[0..NC) = a[0..NC): the 16-byte chunks sub, sub+L, ... of a vector leaf
template <typename T, int NC, int L, bool STREAM = true>
IDOC_DEVICE void store_vec(T* g, int sub, const T* a) {
  constexpr int E = 16 / (int)sizeof(T), NCH = NC / E;
#pragma unroll
  for (int k = 0; k < cdiv(NCH, L); k++) {
    const int c = sub + L * k;
    if (c < NCH) {
      if constexpr (STREAM) st_global_cs<16>(g + c * E, a + c * E);
      else *reinterpret_cast<uint4*>(g + c * E) = *reinterpret_cast<const uint4*>(a + c * E);
    }
  }
}
// flush `cnt` elements (a multiple of the 16-byte chunk) staged in shared memory to global, chunks sub, sub+L, ...
template <typename T, int MAXCNT, int L>
IDOC_DEVICE void flush_block(T* g, int sub, const T* stage, int cnt) {
  constexpr int E = 16 / (int)sizeof(T);
#pragma unroll
  for (int k = 0; k < cdiv(MAXCNT / E, L); k++) {
    const int c = sub + L * k;
    if (c * E < cnt) st_global_cs<16>(g + c * E, stage + c * E);
  }
}

template <typename T, int N, int M, int L, int NBUF, bool HAS_NUB, int WARPS, int BM = 0, int TB = 1, int PADG = 0>
__global__ void __launch_bounds__(WARPS * 32) lqr_adj_fwd_direct_kernel(const LqrVjpArgs<T> a) {
  using C = AdjFwdDCfg<T, N, M, L, NBUF, HAS_NUB, BM, TB, PADG>;
  using Ge = Geo<T, N, M>;
  constexpr int SZ = C::SZ, G = C::G, CX = C::CX, CU = C::CU, NP = Ge::NP, WS = Ge::WS, WS2 = Ge::WS2;
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
  const bool vo = a.reduce_batch != 0;  // vectors only (see LqrVjpArgs): gq, gr, gd, gqf, gx0 and no outer product

  SlabLoader<G, Ge::GR, C::RL, C::RQ, C::NBULK> ld;
  ld.init(prob0, last, Tp, C::GROUP_BYTES, lane, reinterpret_cast<uint64_t*>(wbase + C::BAR_OFF));
  // a [B][T][w] array: the TB steps of a request are contiguous; a sub-range of the residual slot: one segment per step
  auto add_arr = [&](const T* base, int w, int off, bool bulk) {
    if (bulk) ld.add_bulk(base, TB * w * SZ, 0, TB * w * SZ, TB * off * SZ); else ld.add(base, TB * w * SZ, 0, TB * w * SZ, TB * off * SZ);
  };
  auto add_slot = [&](int field, int w, int off, bool bulk) {
#pragma unroll
    for (int par = 0; par < TB; par++) {
      if (bulk) ld.add_bulk(a.ws, TB * WS * SZ, (par * WS + field) * SZ, w * SZ, (TB * off + par * w) * SZ);
      else ld.add(a.ws, TB * WS * SZ, (par * WS + field) * SZ, w * SZ, (TB * off + par * w) * SZ);
    }
  };
  add_arr(a.A, Ge::wA, C::iA, C::LB > 0);
  add_slot(Ge::oVp, C::WV, C::iV, C::LB > 1);
  add_slot(Ge::oK, C::WK, C::iK, C::LB > 2);
  add_arr(a.B, Ge::wB, C::iB, C::LB > 3);
  add_arr(a.ws2, WS2, C::iW2, false);
  add_arr(a.U, M, C::iU, false);
  add_arr(a.Nu, N, C::iNu, false);
  if constexpr (HAS_NUB) add_arr(a.Nub, N, C::iNb, false);
  if constexpr (TB == 1) ld.add(a.X, N * SZ, 0, N * SZ, C::iSX * SZ, -1);  // sX[t] = X[t-1]; x0 at t = 0
  else add_arr(a.X, N, C::iSX, false);                                     // X[t], X[t+1]: sX of step t comes from the carry

  auto issue_loads = [&](int tp, int buf) { ld.issue(wbase32, (C::oIn + buf * C::IN) * SZ, tp, buf); };
  issue_loads(0, 0);
  T* stt = gs + C::oSt;
  for (int i = sub; i < N; i += L) {
    stt[C::sux + i] = T(0);  // ux_0 = x0' = 0
    if constexpr (TB == 1) gs[C::oIn + C::iSX + i] = a.x0[probc * N + i];  // sX[0] = x0 (buffer 0 is read at t = 0)
    else stt[C::sXp + i] = a.x0[probc * N + i];
  }

  for (int tp = 0; tp < Tp; tp++) {
    const T* inb = gs + C::oIn + (tp % NBUF) * C::IN;
    ld.wait(tp % NBUF);
    __syncwarp();
    if (NBUF > 1 && tp + 1 < Tp) issue_loads(tp + 1, (tp + 1) % NBUF);
#pragma unroll
    for (int par = 0; par < TB; par++) {
      const int t = tp * TB + par;
      if (par > 0) __syncwarp();  // the previous step's stores have read the state area
      // field of width w at in-slab offset off, step `par` of the request
      auto F = [&](int off, int w) -> const T* { return inb + TB * off + par * w; };
      const T* inA = F(C::iA, Ge::wA);
      const T* inB = F(C::iB, Ge::wB);
      const T* inK = F(C::iK, C::WK);
      const T* inV = F(C::iV, C::WV);
      const T* inW2 = F(C::iW2, WS2);
      const T* Uv = F(C::iU, M);
      const T* nu = F(C::iNu, N);
      const T* sX = TB == 1 ? inb + C::iSX : (par == 0 ? stt + C::sXp : inb + TB * C::iSX + (par - 1) * N);
      T* uxs = stt + C::sux + (t & 1) * Ge::p16(N);         // sux[t]
      T* uxn = stt + C::sux + ((t & 1) ^ 1) * Ge::p16(N);   // ux' = sux[t+1]

      T ux[N];
      lds<T, N, 16>(uxs, ux);
      // uU = K ux + k'   (own rows)
#pragma unroll
      for (int c = 0; c < CU; c++) {
        const int b = sub * CU + c;
        if (b < M) {
          T Kr[N];
          lds<T, N, RA_N>(inK + b * N, Kr);
          T s = inW2[Ge::o2k + b];
#pragma unroll
          for (int j = 0; j < N; j++) s += Kr[j] * ux[j];
          stt[C::suU + b] = s;
        }
      }
      __syncwarp();
      T uU[M];
      lds<T, M, 16>(stt + C::suU, uU);
      // ux' = A ux + B uU + d'   (own rows)
#pragma unroll
      for (int c = 0; c < CX; c++) {
        const int i = sub * CX + c;
        if (i < N) {
          T Ar[N], Br[M];
          lds<T, N, RA_N>(inA + i * N, Ar);
          lds<T, M, RA_M>(inB + i * M, Br);
          T s = T(0);
          if constexpr (HAS_NUB) s = F(C::iNb, N)[i];
#pragma unroll
          for (int j = 0; j < N; j++) s += Ar[j] * ux[j];
#pragma unroll
          for (int b = 0; b < M; b++) s += Br[b] * uU[b];
          uxn[i] = s;
        }
      }
      __syncwarp();
      // uNu = V_{t+1} ux' + v'_{t+1}  (own rows)
      {
        T xn[N], Vr[NP];
        lds<T, N, 16>(uxn, xn);
        lds<T, NP, 16>(inV, Vr);
#pragma unroll
        for (int i = 0; i < N; i++) {
          if (i / CX == sub) {
            T unu = inW2[Ge::o2v + i];
#pragma unroll
            for (int j = 0; j < N; j++) unu += Vr[triu(i, j, N)] * xn[j];
            stt[C::sun + i] = unu;
          }
        }
      }
      __syncwarp();
      // the gradient leaves of step t, straight from registers
      if (active) {
        const long long bt = prob * Th + t;
        const T* unu = stt + C::sun;
        const T* uUs = stt + C::suU;
        if (!vo) {
          store_outer2<T, N, N, L, false>(a.gA + bt * Ge::wA, sub, nu, uxs, unu, sX);
          store_outer2<T, N, N, L, true>(a.gQ + bt * Ge::wQ, sub, uxs, sX, sX, uxs);
          store_outer2<T, N, M, L, false>(a.gB + bt * Ge::wB, sub, nu, uUs, unu, Uv);
          store_outer2<T, N, M, L, false>(a.gM + bt * Ge::wM, sub, uxs, Uv, sX, uUs);
        }
        if constexpr (C::OCG == 0) {
          if (!vo) store_outer2<T, M, M, L, true>(a.gR + bt * Ge::wR, sub, uUs, Uv, Uv, uUs);
          store_vec<T, N, L>(a.gq + bt * Ge::wq, sub, uxs);
          store_vec<T, M, L>(a.gr + bt * Ge::wr, sub, uUs);
          store_vec<T, N, L>(a.gd + bt * Ge::wd, sub, unu);
        }
        if (t == 0) {
          // gx0 = M[0] uU[0] + A[0]^T uNu[0]
          const T* M0 = a.Mx + (prob * Th) * (N * M);
          for (int i = sub; i < N; i += L) {
            T s = T(0);
#pragma unroll
            for (int b = 0; b < M; b++) s += M0[i * M + b] * uU[b];
#pragma unroll
            for (int k = 0; k < N; k++) s += inA[k * N + i] * unu[k];
            a.gx0[prob * N + i] = s;
          }
        }
        if (t == Th - 1) {
          // gQf = sym(uX[T-1] (x) X[T-1]),  gqf = uX[T-1]
          const T* XT = a.X + (prob * Th + Th - 1) * N;
          T* gQf = a.gQf + prob * (N * N);
          T* gqf = a.gqf + prob * N;
          if (!vo) {
            for (int e = sub; e < N * N; e += L) {
              const int i = e / N, j = e % N;
              gQf[e] = T(0.5) * (uxn[i] * XT[j] + uxn[j] * XT[i]);
            }
          }
          for (int e = sub; e < N; e += L) gqf[e] = uxn[e];
        }
      }
      if constexpr (C::OCG > 0) {
        const int oc = t % C::OCG;
        if (!vo) store_outer2<T, M, M, L, true, false>(stt + C::sgR + oc * (M * M), sub, stt + C::suU, Uv, Uv, stt + C::suU);
        store_vec<T, N, L, false>(stt + C::sgq + oc * N, sub, uxs);
        store_vec<T, M, L, false>(stt + C::sgr + oc * M, sub, stt + C::suU);
        store_vec<T, N, L, false>(stt + C::sgd + oc * N, sub, stt + C::sun);
        if (oc == C::OCG - 1 || t == Th - 1) {
          __syncwarp();
          if (active) {
            const long long bt0 = prob * Th + (t - oc);
            if (!vo) flush_block<T, C::OCG * M * M, L>(a.gR + bt0 * Ge::wR, sub, stt + C::sgR, (oc + 1) * M * M);
            flush_block<T, C::OCG * N, L>(a.gq + bt0 * Ge::wq, sub, stt + C::sgq, (oc + 1) * N);
            flush_block<T, C::OCG * M, L>(a.gr + bt0 * Ge::wr, sub, stt + C::sgr, (oc + 1) * M);
            flush_block<T, C::OCG * N, L>(a.gd + bt0 * Ge::wd, sub, stt + C::sgd, (oc + 1) * N);
          }
        }
      }
    }
    if constexpr (TB > 1) {  // carry X of the request's last step: sX of the next request's first step
      __syncwarp();          // (the first step's stores have read the previous carry)
      for (int i = sub; i < N; i += L) stt[C::sXp + i] = inb[TB * C::iSX + (TB - 1) * N + i];
    }
    if (NBUF == 1) {
      __syncwarp();
      if (tp + 1 < Tp) issue_loads(tp + 1, 0);
    }
  }
}


}  // namespace idoc
