// Mid-size Riccati sweep on the tensor cores (fp32 problems, n a multiple of 16, m a multiple of 8; one warp per problem at
// n = 32, two problems per warp at n = 16).
//
// Same mathematics, arguments, residual-slot layout and status bits as mid_riccati_kernel (csrc/mid.cuh; reference:
// idoc/lqr.py:77-95, idoc/tilqr.py:56-78): only the dense products of a step change pipe —
//     W   = V [A B]                                         idoc/lqr.py:80-85 (the V A, V B factors)
//     Gxx = sym(Q) + A^T W_A,  Gux = M^T + B^T W_A,  Guu = sym(R) + B^T W_B           idoc/lqr.py:80-82
//     V+  = Gxx + Gux^T K - s K^T K                                                   idoc/lqr.py:91
// run as warp-level mma.sync.m16n8k8 TF32 products with FP32 accumulators.  TF32 keeps 10 mantissa bits, far from
// north_star's 1e-4 bar over a 200-step recursion, so every operand is split into  x = hi + lo  (hi = tf32(x),
// lo = tf32(x - hi)) and each product is three MMAs  lo*hi + hi*lo + hi*hi  ("3xTF32"): the dropped lo*lo term is 2^-22
// relative, i.e. FP32-level accuracy (parity is asserted at 1e-4 against the fp64 oracle like every other fp32 path).
//
// Why: the FFMA version of this sweep is bound by the SHARED-MEMORY pipe, not by the math pipe (profiles/r1e_*: LSU
// wavefronts 67 - 88 %, FMA pipe 17 %): a 4x8 register tile needs 12 operand scalars from shared memory per 32 FMAs and every
// LDS.128 costs four wavefronts however many lanes share its address.  An MMA fragment feeds 1024 multiply-adds from 6
// scalars per lane, and fragments are reused across the tiles of a product (A across the n-tiles, B across the m-tiles).
// Shared-memory matrices get row strides chosen so that fragment loads are conflict-free:
//     A-operand read as A[i][k] from a row-major matrix: stride = 4 (mod 32) (bank = 4 g + tig);
//     operands read "transposed" (element (i, k) at [k][i]: A^T from F, B operands from F / W / K): stride = 8 or 24 (mod 32)
//     (bank = 8 tig + g).
#pragma once
#include "mid.cuh"

namespace idoc {

__host__ __device__ constexpr int ld_mod(int w, int r8, int modulo) {  // smallest ld >= w with ld % modulo == r8 (or modulo - r8)
  int ld = w;
  while (!((ld % modulo) == r8 || (ld % modulo) == modulo - r8)) ld++;
  return ld;
}

// Fragment geometry per scalar type.  float: mma.sync.m16n8k8 TF32 (3xTF32 split), 16 x 8 accumulator tiles, 4 values per lane;
// double: mma.sync.m8n8k4 FP64 (DMMA, exact), 8 x 8 tiles, 2 values per lane.  In both, accumulator value e of lane
// (g = lane / 4, tig = lane % 4) is element (g + 8 (e / 2), 2 tig + (e & 1)) of its tile.
// Row strides that make fragment loads conflict-free (banks are 4 bytes wide; an 8-byte access takes half a warp per wavefront):
//   float : A[i][k] reads want stride = 4 (mod 8) (bank 4 g + tig); transposed / B-operand reads = 8 or 24 (mod 32) (8 tig + g)
//   double: both patterns want stride = 4 or 12 (mod 16) doubles (g, tig in 0..3 within a half warp)
template <typename T>
struct MmaGeo;
template <>
struct MmaGeo<float> {
  static constexpr int TM = 16, TK = 8, CE = 4;
  static constexpr int ld_a(int w) { return ld_mod(w, 4, 8); }
  static constexpr int ld_t(int w) { return ld_mod(w, 8, 32); }
  static constexpr int ld_q(int w) { return ld_mod(w, 8, 32); }
};
template <>
struct MmaGeo<double> {
  static constexpr int TM = 8, TK = 4, CE = 2;
  static constexpr int ld_a(int w) { return ld_mod(w, 4, 8); }
  static constexpr int ld_t(int w) { return ld_mod(w, 4, 16); }
  static constexpr int ld_q(int w) { return ld_mod(w, 8, 16); }  // 16-byte accumulator-image loads: rows 64 bytes apart
};

template <typename T, int N, int M, bool TI>
struct MmaCfgT {
  using MG = MmaGeo<T>;
  static_assert(N % 16 == 0 && M % 8 == 0 && N <= 32 && M <= 16, "mma sweep: n in {16, 32}, m in {8, 16}");
  using Ge = Geo<T, N, M>;
  static constexpr int F = N + M;
  static constexpr int LDV = MG::ld_a(N + 1);       // V read as A[i][k] = V[i][k]
  static constexpr int LDF = MG::ld_t(F);           // F = [A | B]: element (i, k) at [k][i] (A^T, B^T) and B operand [k][n]
  static constexpr int LDW = LDF;                   // W: B operand [k][n]
  static constexpr int LDQ = MG::ld_q(N);           // sym(Q): accumulator-fragment initial values (row pairs of 2 columns)
  static constexpr int LDG = MG::ld_t(N);           // Gux [M][N]: element (i, k) of Gux^T at [k][i]
  static constexpr int LDK = MG::ld_t(N);           // K [M][N]: B operand [k][n]
  static constexpr int MTN = N / MG::TM, MTM = (M + MG::TM - 1) / MG::TM, NTN = N / 8, NTM = M / 8, NTF = F / 8;
  // RF (time-invariant, n = 16): the operands of a step that never change stay in REGISTERS for the whole sweep (float: already
  // split into their TF32 halves) -- the B fragments of F = [A | B] for W = V F and the accumulator image of sym(Q) for Gxx --
  // so the step neither loads nor splits them again and sym(Q) has no shared-memory copy (more resident warps)
  static constexpr bool RF = TI && N == 16;
  static constexpr int oF = 0;
  static constexpr int oQ = oF + N * LDF;
  static constexpr int oMx = oQ + (RF ? 0 : N * LDQ);  // [N][M]
  static constexpr int oR = oMx + (TI ? 0 : N * M);   // [M][M]
  static constexpr int oq = oR + M * M;
  static constexpr int or_ = oq + (TI ? 0 : N);
  static constexpr int od = or_ + (TI ? 0 : M);
  static constexpr int oV = od + (TI ? 0 : N);        // [N][LDV]
  static constexpr int ov = oV + N * LDV;
  static constexpr int ow = ov + N;
  static constexpr int oW = ow + N;                   // [N][LDW]
  static constexpr int oGux = oW + N * LDW;           // [M][LDG]
  // W is dead once the G blocks are formed: Guu, its working copy, the inverse factor, K and the Jacobi scratch live in its
  // range when they fit (n = 32), after Gux otherwise
  static constexpr bool ALIAS = 4 * M * M + M * LDK <= N * LDW;
  static constexpr int oGuu = ALIAS ? oW : oGux + M * LDG;
  static constexpr int oGt = oGuu + M * M;
  static constexpr int oLi = oGt + M * M;
  static constexpr int oK = oLi + M * M;              // [M][LDK]
  static constexpr int oJac = oK + M * LDK;
  static constexpr int ogx = ALIAS ? oGux + M * LDG : oJac + M * M;
  static constexpr int ogu = ogx + N;
  static constexpr int okk = ogu + M;
  static constexpr int ELEMS = (okk + M + 8 + 31) / 32 * 32;
  // two warps per problem: CTAs per SM that shared memory allows (at most 8: 128 registers per thread), the register cap of that build
  static constexpr int SMEM_CTAS = (227 * 1024) / (ELEMS * (int)sizeof(T) + 1024);
  static constexpr int NW_MINB = SMEM_CTAS > 8 ? 8 : (SMEM_CTAS < 1 ? 1 : SMEM_CTAS);
};
template <int N, int M, bool TI>
using MmaCfg = MmaCfgT<float, N, M, TI>;

// x = hi + lo with hi on the TF32 grid (10 mantissa bits).  sm_100 has no single-instruction cvt.rna.tf32 (ptxas expands it to a
// compare, a predicated integer add and, for the residual, a mask: 6 instructions per operand word for hi and lo together), so
// the split is done on the bit pattern: hi = round-half-up to 13 dropped bits (integer add + mask; Inf / NaN keep their class),
// lo = x - hi exactly (|lo| <= 2^-11 |x|), then the same half-ulp add WITHOUT the mask: the tensor core reads only the upper 19
// bits of a TF32 operand, so the add makes its truncation a rounding (error <= 2^-22 |x|, the size of the lo*lo term 3xTF32
// drops anyway).  4 instructions per operand word.
IDOC_DEVICE void split_tf32(float x, uint32_t& hi, uint32_t& lo) {
  hi = (__float_as_uint(x) + 0x1000u) & 0xffffe000u;
  lo = __float_as_uint(x - __uint_as_float(hi)) + 0x1000u;
}
IDOC_DEVICE void mma_tf32(float (&c)[4], const uint32_t (&a)[4], const uint32_t (&b)[2]) {
  asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};\n"
               : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
               : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
}
// c += a b with a = ah + al, b = bh + bl (small terms first)
IDOC_DEVICE void mma3(float (&c)[4], const uint32_t (&ah)[4], const uint32_t (&al)[4], const uint32_t (&bh)[2], const uint32_t (&bl)[2]) {
  mma_tf32(c, al, bh);
  mma_tf32(c, ah, bl);
  mma_tf32(c, ah, bh);
}
// A fragment (16 x 8, rows i0.., k0..): element (i, k) at p[i * si + k * sk]; lane (g = lane / 4, tig = lane % 4) holds
// (g, tig), (g + 8, tig), (g, tig + 4), (g + 8, tig + 4).  HALF: rows >= 8 are outside the matrix (m = 8): zero.
template <bool HALF = false>
IDOC_DEVICE void load_a(const float* p, int si, int sk, int g, int tig, uint32_t (&ah)[4], uint32_t (&al)[4]) {
  split_tf32(p[g * si + tig * sk], ah[0], al[0]);
  split_tf32(p[g * si + (tig + 4) * sk], ah[2], al[2]);
  if (HALF) {
    ah[1] = al[1] = ah[3] = al[3] = 0u;
  } else {
    split_tf32(p[(g + 8) * si + tig * sk], ah[1], al[1]);
    split_tf32(p[(g + 8) * si + (tig + 4) * sk], ah[3], al[3]);
  }
}
// B fragment (8 x 8, k0.., n0..): element (k, n) at p[k * sk + n * sn]; lane holds (tig, g), (tig + 4, g)
IDOC_DEVICE void load_b(const float* p, int sk, int sn, int g, int tig, uint32_t (&bh)[2], uint32_t (&bl)[2]) {
  split_tf32(p[tig * sk + g * sn], bh[0], bl[0]);
  split_tf32(p[(tig + 4) * sk + g * sn], bh[1], bl[1]);
}

IDOC_DEVICE void dmma884(double (&c)[2], double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n" : "+d"(c[0]), "+d"(c[1]) : "d"(a), "d"(b));
}

// Typed fragment operations used by the sweep below
template <typename T>
struct MmaOps;
template <>
struct MmaOps<float> {
  struct FA { uint32_t h[4], l[4]; };
  struct FB { uint32_t h[2], l[2]; };
  template <bool HALF>
  static IDOC_DEVICE void lda(const float* p, int si, int sk, int g, int tig, FA& f) { load_a<HALF>(p, si, sk, g, tig, f.h, f.l); }
  static IDOC_DEVICE void ldb(const float* p, int sk, int sn, int g, int tig, FB& f) { load_b(p, sk, sn, g, tig, f.h, f.l); }
  static IDOC_DEVICE void mma(float (&c)[4], const FA& a, const FB& b) { mma3(c, a.h, a.l, b.h, b.l); }
  static IDOC_DEVICE void st2(float* p, float x, float y) { *reinterpret_cast<float2*>(p) = make_float2(x, y); }
  static IDOC_DEVICE void ld2(const float* p, float& x, float& y) { const float2 v = *reinterpret_cast<const float2*>(p); x = v.x; y = v.y; }
};
template <>
struct MmaOps<double> {
  struct FA { double v; };
  struct FB { double v; };
  template <bool HALF>
  static IDOC_DEVICE void lda(const double* p, int si, int sk, int g, int tig, FA& f) { f.v = p[g * si + tig * sk]; }
  static IDOC_DEVICE void ldb(const double* p, int sk, int sn, int g, int tig, FB& f) { f.v = p[tig * sk + g * sn]; }
  static IDOC_DEVICE void mma(double (&c)[2], const FA& a, const FB& b) { dmma884(c, a.v, b.v); }
  static IDOC_DEVICE void st2(double* p, double x, double y) { *reinterpret_cast<double2*>(p) = make_double2(x, y); }
  static IDOC_DEVICE void ld2(const double* p, double& x, double& y) { const double2 v = *reinterpret_cast<const double2*>(p); x = v.x; y = v.y; }
};

// G problems per warp (G = 2 at n = 16): the vector phases (slot stores, matrix-vector products, the m x m factorisation, the
// gain solves) run with 32 / G lanes per problem exactly as in mid_riccati_kernel, so no lane idles there; the MMA phases are
// warp-wide and visit the G problems one after the other (operands from that problem's shared-memory block, accumulators of
// all G problems in registers).  Full-warp barriers separate the two kinds of phase.
#ifndef IDOC_MMA_RF_MINB
#define IDOC_MMA_RF_MINB 14  // 14 resident warps (the shared-memory limit): 128 registers with ~60 B of spills measured 12 % faster than 167 registers and 12 warps
#endif
// NW warps per problem (NW = 2 at n = 32, where one problem per warp leaves an SM with 7 resident warps): the accumulator
// tiles of every product are split by n-tile between the warps (each warp loads all A fragments and its own B fragments), the
// element-wise loops of the vector phases run over all 32 NW lanes, the m x m factorisation stays on warp 0, and the barriers
// between phases are CTA barriers.
template <typename T, int N, int M, bool TI, int G = 1, int NW = 1>
__global__ void __launch_bounds__(32 * NW, NW > 1 ? MmaCfgT<T, N, M, TI>::NW_MINB : ((sizeof(T) == 4 && MmaCfgT<T, N, M, TI>::RF && G == 2) ? IDOC_MMA_RF_MINB : 1))
mid_riccati_mma_kernel(const GenFwdArgs<T> ga) {
  using C = MmaCfgT<T, N, M, TI>;
  using Ge = Geo<T, N, M>;
  using MG = MmaGeo<T>;
  using Op = MmaOps<T>;
  using FA = typename Op::FA;
  using FB = typename Op::FB;
  constexpr int TM = MG::TM, TK = MG::TK, CE = MG::CE;
  constexpr int WS = Ge::WS, LDV = C::LDV, LDF = C::LDF, LDW = C::LDW, LDQ = C::LDQ, LDG = C::LDG, LDK = C::LDK;
  constexpr int MTN = C::MTN, MTM = C::MTM, NTN = C::NTN, NTM = C::NTM, NTF = C::NTF;
  constexpr int LP = NW > 1 ? 32 * NW : 32 / G;  // lanes of a problem in the vector phases
  constexpr int SW = NW > 1 ? 32 : LP;           // width of the shuffles / reductions of the factorisation (warp 0 when NW > 1)
  constexpr int NTFW = (NTF + NW - 1) / NW, NTNW = (NTN + NW - 1) / NW, NTMW = (NTM + NW - 1) / NW;  // n-tiles per warp
  static_assert(NW == 1 || (NW == 2 && G == 1 && !MmaCfgT<T, N, M, TI>::RF && NTN % NW == 0), "two warps per problem: one problem per CTA");
  constexpr int GSTRIDE = C::ELEMS + (G > 1 ? 16 : 0);  // 16 mod 32 words: the problems of a warp sit on different bank halves
  constexpr bool HALF_M = M < TM;  // float, m = 8: the B^T tile has 8 valid rows of 16
  static_assert(G == 1 || G == 2, "problems per warp");
  static_assert(LP >= N && LP >= M, "one lane per row in the vector phases");
  extern __shared__ __align__(16) char smem_raw[];
  const int tid = threadIdx.x & 31, wid = threadIdx.x >> 5, g = tid >> 2, tig = tid & 3;  // MMA fragment coordinates
  const int lane = threadIdx.x % LP, grp = threadIdx.x / LP;  // vector phases: lane within the problem's group
  const int ntf0 = wid * NTFW, ntn0 = wid * NTNW, ntm0 = wid * NTMW;  // first n-tile of this warp
  auto psync = [&]() { if constexpr (NW > 1) __syncthreads(); else __syncwarp(); };  // every lane that works on the warp's / CTA's problems
  T* s0 = reinterpret_cast<T*>(smem_raw);
  T* s = s0 + grp * GSTRIDE;
  const LqrFwdArgs<T>& a = ga.a;
  constexpr bool ti = TI;
  const int Th = a.T_;
  const long long praw = (long long)blockIdx.x * G + grp;
  const long long prob = praw < a.nb ? praw : a.nb - 1;  // an idle group of the last warp shadows the last problem
  const unsigned gmask = G == 1 ? 0xffffffffu : (0xffffu << (16 * grp));
  auto gsync = [&]() { if constexpr (NW > 1) __syncthreads(); else __syncwarp(gmask); };  // the lanes of ONE problem
  T* sshift = s + C::okk + M;  // NW > 1: the shift of the step, from warp 0 to the others
  T* sF = s + C::oF; T* sQ = s + C::oQ; T* sMx = s + C::oMx; T* sR = s + C::oR; T* sq = s + C::oq; T* sr = s + C::or_;
  T* sd = s + C::od; T* sV = s + C::oV; T* sv = s + C::ov; T* sw = s + C::ow;  T* sGux = s + C::oGux;
  T* sGuu = s + C::oGuu; T* sGt = s + C::oGt; T* sLi = s + C::oLi; T* sK = s + C::oK; T* sgx = s + C::ogx; T* sgu = s + C::ogu;
  T* skk = s + C::okk;

  {  // V_T = sym(Qf), v_T = qf
    const T* Qf = a.Qf + prob * N * N;
    for (int e = lane; e < N * N; e += LP) {
      const int i = e / N, j = e % N;
      sV[i * LDV + j] = T(0.5) * (Qf[i * N + j] + Qf[j * N + i]);
    }
    mid_g2s<T, LP>(sv, a.qf ? a.qf + prob * N : (const T*)nullptr, N, lane);
  }
  if (ti) {
    const T* A = a.A + prob * N * N;
    const T* B = a.B + prob * N * M;
    const T* Q = a.Q + prob * N * N;
    for (int e = lane; e < N * N; e += LP) sF[(e / N) * LDF + (e % N)] = A[e];
    for (int e = lane; e < N * M; e += LP) sF[(e / M) * LDF + N + (e % M)] = B[e];
    if constexpr (!C::RF) {
      for (int e = lane; e < N * N; e += LP) sQ[(e / N) * LDQ + (e % N)] = T(0.5) * (Q[e] + Q[(e % N) * N + e / N]);
    }
    mid_g2s<T, LP>(sR, a.R + prob * M * M, M * M, lane);
  }
  auto issue_tv = [&](int t) {  // 16-byte cp.async copies, rows into the padded strides
    constexpr int E = 16 / (int)sizeof(T);
    const long long idx = prob * Th + t;
    for (int c = lane; c < N * N / E; c += LP) cp_async<16>(smem_u32(sF + (c / (N / E)) * LDF + (c % (N / E)) * E), a.A + idx * N * N + c * E);
    for (int c = lane; c < N * M / E; c += LP) cp_async<16>(smem_u32(sF + (c / (M / E)) * LDF + N + (c % (M / E)) * E), a.B + idx * N * M + c * E);
    for (int c = lane; c < N * N / E; c += LP) cp_async<16>(smem_u32(sQ + (c / (N / E)) * LDQ + (c % (N / E)) * E), a.Q + idx * N * N + c * E);
    if (a.Mx != nullptr) mid_cp_range16<T, LP>(sMx, a.Mx + idx * N * M, N * M, lane);
    mid_cp_range16<T, LP>(sR, a.R + idx * M * M, M * M, lane);
    if (a.q != nullptr) mid_cp_range16<T, LP>(sq, a.q + idx * N, N, lane);
    if (a.r != nullptr) mid_cp_range16<T, LP>(sr, a.r + idx * M, M, lane);
    if (a.d != nullptr) mid_cp_range16<T, LP>(sd, a.d + idx * N, N, lane);
    cp_async_commit();
  };
  if (!ti) {
    mid_g2s<T, LP>(sMx, (const T*)nullptr, N * M, lane);  // absent leaves are zero
    mid_g2s<T, LP>(sq, (const T*)nullptr, N, lane);
    mid_g2s<T, LP>(sr, (const T*)nullptr, M, lane);
    mid_g2s<T, LP>(sd, (const T*)nullptr, N, lane);
    issue_tv(Th - 1);
  }
  T dC = T(0), dc = T(0);
  int stat = 0;
  psync();
  // register-resident fragments (RF): built once from the shared-memory F of each problem of the warp and from Q in global memory
  constexpr int RG = C::RF ? G : 1, RK = C::RF ? N / TK : 1, RNF = C::RF ? NTF : 1, RMT = C::RF ? MTN : 1, RNT = C::RF ? NTN : 1;
  FB fb[RG][RK][RNF];
  T qacc[RG][RMT][RNT][CE];
  if constexpr (C::RF) {
#pragma unroll
    for (int pg = 0; pg < G; pg++) {
      const T* pF = s0 + pg * GSTRIDE + C::oF;
#pragma unroll
      for (int ks = 0; ks < N / TK; ks++)
#pragma unroll
        for (int nt = 0; nt < NTF; nt++) Op::ldb(pF + (ks * TK) * LDF + nt * 8, LDF, 1, g, tig, fb[pg][ks][nt]);
      const long long pr = (long long)blockIdx.x * G + pg;
      const T* Q = a.Q + (pr < a.nb ? pr : a.nb - 1) * N * N;
#pragma unroll
      for (int mt = 0; mt < MTN; mt++)
#pragma unroll
        for (int nt = 0; nt < NTN; nt++)
#pragma unroll
          for (int e = 0; e < CE; e++) {
            const int r = mt * TM + g + (e >= 2 ? 8 : 0), c = nt * 8 + 2 * tig + (e & 1);
            qacc[pg][mt][nt][e] = T(0.5) * (Q[r * N + c] + Q[c * N + r]);
          }
    }
  }

  for (int t = Th - 1; t >= 0; t--) {
    const long long idx = prob * Th + t;
    T* wsl = a.ws + idx * WS;
    if (!ti) {
      cp_async_wait<0>();
      gsync();
      // symmetrise Q in place along wrapped diagonals (lane l averages (l, l+d) with (l+d, l))
      if (lane < N) {
#pragma unroll 4
        for (int d = 1; d <= N / 2; d++) {
          int c = lane + d;
          c = c >= N ? c - N : c;
          if (d < N / 2 || lane < N / 2) {
            const T mm = T(0.5) * (sQ[lane * LDQ + c] + sQ[c * LDQ + lane]);
            sQ[lane * LDQ + c] = mm;
            sQ[c * LDQ + lane] = mm;
          }
        }
      }
    }
    // slot: the value function entering the step (packed upper), straight to global
    if (lane < N) {
#pragma unroll
      for (int i = 0; i < N; i++)
        if (i <= lane) wsl[Ge::oVp + (i * N - i * (i - 1) / 2 - i) + lane] = sV[i * LDV + lane];
      wsl[Ge::ov + lane] = TI ? T(0) : sv[lane];
    }
    // ---- w = v + V d.  tilqr (TI) has no affine terms at all (idoc/tilqr.py:12-19: Q, Qf, R, A, B only): v, w, gx, gu, k and
    // the expected-change terms are identically zero and their arithmetic is compiled out
    if constexpr (!TI) {
      if (lane < N) {
        sw[lane] = col_dot4<T, N>(sV + lane, LDV, sd, sv[lane]);
      }
    }
    psync();  // sym(Q) of every problem of the warp is in place: warp-wide products from here
    // ---- W = V [A B]   (M = N rows, N = F columns, K = N)
#pragma unroll
    for (int pg = 0; pg < G; pg++) {
      const T* pV = s0 + pg * GSTRIDE + C::oV;
      const T* pF = s0 + pg * GSTRIDE + C::oF;
      T* pW = s0 + pg * GSTRIDE + C::oW;
      T c[MTN][NTFW][CE];
#pragma unroll
      for (int mt = 0; mt < MTN; mt++)
#pragma unroll
        for (int nt = 0; nt < NTFW; nt++)
#pragma unroll
          for (int e = 0; e < CE; e++) c[mt][nt][e] = T(0);
#pragma unroll
      for (int ks = 0; ks < N / TK; ks++) {
        FA fa[MTN];
#pragma unroll
        for (int mt = 0; mt < MTN; mt++) Op::template lda<false>(pV + (mt * TM) * LDV + ks * TK, LDV, 1, g, tig, fa[mt]);
#pragma unroll
        for (int i = 0; i < NTFW; i++) {
          const int nt = ntf0 + i;
          if (NTF % NW != 0 && nt >= NTF) continue;  // warp-uniform
          if constexpr (C::RF) {
#pragma unroll
            for (int mt = 0; mt < MTN; mt++) Op::mma(c[mt][i], fa[mt], fb[pg][ks][i]);
          } else {
            FB b;
            Op::ldb(pF + (ks * TK) * LDF + nt * 8, LDF, 1, g, tig, b);
#pragma unroll
            for (int mt = 0; mt < MTN; mt++) Op::mma(c[mt][i], fa[mt], b);
          }
        }
      }
      // accumulator fragment: (g, 2 tig), (g, 2 tig + 1) [, (g + 8, 2 tig), (g + 8, 2 tig + 1)]
#pragma unroll
      for (int mt = 0; mt < MTN; mt++)
#pragma unroll
        for (int i = 0; i < NTFW; i++) {
          const int nt = ntf0 + i;
          if (NTF % NW != 0 && nt >= NTF) continue;
          Op::st2(pW + (mt * TM + g) * LDW + nt * 8 + 2 * tig, c[mt][i][0], c[mt][i][1]);
          if constexpr (CE == 4) Op::st2(pW + (mt * TM + g + 8) * LDW + nt * 8 + 2 * tig, c[mt][i][2], c[mt][i][3]);
        }
    }
    psync();
    // ---- G blocks: Gxx stays in accumulator fragments until V+ is formed
    T gxx[G][MTN][NTNW][CE];  // this warp's n-tiles: columns 8 (ntn0 + i) ..
#pragma unroll
    for (int pg = 0; pg < G; pg++) {
      const T* pb = s0 + pg * GSTRIDE;
      const T *pF = pb + C::oF, *pQ = pb + C::oQ, *pMx = pb + C::oMx, *pR = pb + C::oR, *pW = pb + C::oW;
      T *pGux = s0 + pg * GSTRIDE + C::oGux, *pGt = s0 + pg * GSTRIDE + C::oGt;
      T gux[MTM][NTNW][CE], guu[MTM][NTMW][CE];
#pragma unroll
      for (int mt = 0; mt < MTN; mt++)
#pragma unroll
        for (int i = 0; i < NTNW; i++) {
          const int nt = ntn0 + i;
          if constexpr (C::RF) {
#pragma unroll
            for (int e = 0; e < CE; e++) gxx[pg][mt][i][e] = qacc[pg][mt][i][e];
          } else {
            Op::ld2(pQ + (mt * TM + g) * LDQ + nt * 8 + 2 * tig, gxx[pg][mt][i][0], gxx[pg][mt][i][1]);
            if constexpr (CE == 4) Op::ld2(pQ + (mt * TM + g + 8) * LDQ + nt * 8 + 2 * tig, gxx[pg][mt][i][2], gxx[pg][mt][i][3]);
          }
        }
#pragma unroll
      for (int mu = 0; mu < MTM; mu++) {
#pragma unroll
        for (int i = 0; i < NTNW; i++) {  // Gux[c][j] = M[j][c] + ...: rows c = mu TM + g (+ 8), columns j = nt*8 + 2 tig (+1)
#pragma unroll
          for (int e = 0; e < CE; e++) {
            const int row = mu * TM + g + (e >= 2 ? 8 : 0), col = (ntn0 + i) * 8 + 2 * tig + (e & 1);
            gux[mu][i][e] = (TI || row >= M) ? T(0) : pMx[col * M + row];
          }
        }
#pragma unroll
        for (int i = 0; i < NTMW; i++) {
#pragma unroll
          for (int e = 0; e < CE; e++) {
            const int row = mu * TM + g + (e >= 2 ? 8 : 0), col = (ntm0 + i) * 8 + 2 * tig + (e & 1);
            guu[mu][i][e] = (row >= M || (NTM % NW != 0 && ntm0 + i >= NTM)) ? T(0) : pR[row * M + col];
          }
        }
      }
#pragma unroll
      for (int ks = 0; ks < N / TK; ks++) {
        FA fa[MTN], fu[MTM];
#pragma unroll
        for (int mt = 0; mt < MTN; mt++) Op::template lda<false>(pF + (ks * TK) * LDF + mt * TM, 1, LDF, g, tig, fa[mt]);  // A^T[i][k] = F[k][i]
#pragma unroll
        for (int mu = 0; mu < MTM; mu++) Op::template lda<HALF_M>(pF + (ks * TK) * LDF + N + mu * TM, 1, LDF, g, tig, fu[mu]);  // B^T[i][k] = F[k][N + i]
#pragma unroll
        for (int i = 0; i < NTNW; i++) {
          FB b;
          Op::ldb(pW + (ks * TK) * LDW + (ntn0 + i) * 8, LDW, 1, g, tig, b);
#pragma unroll
          for (int mt = 0; mt < MTN; mt++) Op::mma(gxx[pg][mt][i], fa[mt], b);
#pragma unroll
          for (int mu = 0; mu < MTM; mu++) Op::mma(gux[mu][i], fu[mu], b);
        }
#pragma unroll
        for (int i = 0; i < NTMW; i++) {
          if (NTM % NW != 0 && ntm0 + i >= NTM) continue;  // warp-uniform
          FB b;
          Op::ldb(pW + (ks * TK) * LDW + N + (ntm0 + i) * 8, LDW, 1, g, tig, b);
#pragma unroll
          for (int mu = 0; mu < MTM; mu++) Op::mma(guu[mu][i], fu[mu], b);
        }
      }
#pragma unroll
      for (int mu = 0; mu < MTM; mu++)
#pragma unroll
        for (int i = 0; i < NTNW; i++) {
          const int nt = ntn0 + i;
          Op::st2(pGux + (mu * TM + g) * LDG + nt * 8 + 2 * tig, gux[mu][i][0], gux[mu][i][1]);
          if constexpr (CE == 4 && !HALF_M) Op::st2(pGux + (mu * TM + g + 8) * LDG + nt * 8 + 2 * tig, gux[mu][i][2], gux[mu][i][3]);
        }
      if constexpr (C::ALIAS) psync();  // every lane has read W: its range becomes Guu / factor / K
#pragma unroll
      for (int mu = 0; mu < MTM; mu++)
#pragma unroll
        for (int i = 0; i < NTMW; i++) {
          const int nt = ntm0 + i;
          if (NTM % NW != 0 && nt >= NTM) continue;
          Op::st2(pGt + (mu * TM + g) * M + nt * 8 + 2 * tig, guu[mu][i][0], guu[mu][i][1]);  // unsymmetrised
          if constexpr (CE == 4 && !HALF_M) Op::st2(pGt + (mu * TM + g + 8) * M + nt * 8 + 2 * tig, guu[mu][i][2], guu[mu][i][3]);
        }
    }
    // gx = q + A^T w, gu = r + B^T w
    if constexpr (!TI) {
      if (lane < N) {
        sgx[lane] = col_dot4<T, N>(sF + lane, LDF, sw, sq[lane]);
      }
      if (lane < M) {
        sgu[lane] = col_dot4<T, N>(sF + N + lane, LDF, sw, sr[lane]);
      }
    }
    psync();
    if (!ti && t > 0) issue_tv(t - 1);  // F, Q, M, R, q, r, d of this step are dead from here on
    if constexpr (!TI) {
      for (int e = lane; e < M * M; e += LP) {
        const int i = e / M, j = e % M;
        sGuu[e] = T(0.5) * (sGt[i * M + j] + sGt[j * M + i]);
      }
      gsync();
    }
    // ---- Cholesky of Guu (+ shift): identical to mid_riccati_kernel (lanes own rows, shuffles inside the factorisation);
    // the problems of a warp may take different paths through the retry: barriers and shuffles are scoped to the group
    T shift = T(0);
    if (NW == 1 || wid == 0)  // the factorisation is a single-warp chain (shuffles)
    for (int attempt = 0; attempt < 2; attempt++) {
      T gr[M], dinv[M];
      {
        const int row = lane < M ? lane : 0;
        if constexpr (TI) {  // sym(Guu) row straight into registers (no shift, no Jacobi path: nothing else reads it)
          T rw[M];
          lds<T, M, align_of(M, (int)sizeof(T))>(sGt + row * M, rw);
#pragma unroll
          for (int c = 0; c < M; c++) gr[c] = T(0.5) * (rw[c] + sGt[c * M + row]);
          __syncwarp(gmask);  // the factor below is written over sGt
        } else {
          lds<T, M, align_of(M, (int)sizeof(T))>(sGuu + row * M, gr);
#pragma unroll
          for (int c = 0; c < M; c++) gr[c] += (c == lane) ? shift : T(0);  // select, not a divergent branch per column
        }
      }
      bool ok = true;
#pragma unroll
      for (int j = 0; j < M; j++) {
        T p = __shfl_sync(gmask, gr[j], j, SW);
        if (!(p > T(0))) {
          ok = false;
          if (attempt == 1) { p = T(IDOC_EPS) * T(1e-3); stat |= ST_CHOL_CLAMP; }
          else {
            p = T(1);
            if (ti) stat |= ST_CHOL_CLAMP | ST_NONFINITE;  // tilqr: no shift, no retry (idoc/tilqr.py:72)
          }
        }
        const T rs = rsqrt_<T>(p);
        dinv[j] = rs;
        const T lij = (lane == j) ? p * rs : gr[j] * rs;
        gr[j] = lij;
#pragma unroll
        for (int c = j + 1; c < M; c++) gr[c] -= lij * __shfl_sync(gmask, lij, c, SW);
      }
      if (lane < M) sts<T, M, align_of(M, (int)sizeof(T))>(sGt + lane * M, gr);
      __syncwarp(gmask);
      T li[M];
#pragma unroll
      for (int i = 0; i < M; i++) li[i] = (i == lane) ? dinv[i] : T(0);
#pragma unroll
      for (int i = 1; i < M; i++) {
        T Lrow[M];
        lds<T, M, align_of(M, (int)sizeof(T))>(sGt + i * M, Lrow);
        T acc = T(0);
#pragma unroll
        for (int c = 0; c < i; c++) acc += Lrow[c] * li[c];
        if (i > lane) li[i] = -acc * dinv[i];
      }
      if (lane < M) {
#pragma unroll
        for (int i = 0; i < M; i++) sLi[i * M + lane] = li[i];
      }
      __syncwarp(gmask);
      T fro = T(0);
      if (lane < M) {
#pragma unroll
        for (int i = 0; i < M; i++) fro += li[i] * li[i];
      }
      fro = group_sum<T, SW>(fro, gmask);
      if (attempt == 1 || ti) break;
      if (ok && fro < T(1.0 / IDOC_EPS)) break;
      T mn = T(0);
      if (lane == 0) {  // slow path: lambda_min by Jacobi sweeps
        T* Am = s + C::oJac;
        for (int e = 0; e < M * M; e++) Am[e] = sGuu[e];
        for (int sweep = 0; sweep < 30; sweep++) {
          T off = T(0), dg = T(0);
          for (int i = 0; i < M; i++) { dg += Am[i * M + i] * Am[i * M + i]; for (int j = i + 1; j < M; j++) off += Am[i * M + j] * Am[i * M + j]; }
          if (!(off > (sizeof(T) == 8 ? T(1e-30) : T(1e-14)) * (dg + off))) break;
          for (int p = 0; p < M - 1; p++)
            for (int q2 = p + 1; q2 < M; q2++) {
              const T apq = Am[p * M + q2];
              if (apq == T(0)) continue;
              const T theta = (Am[q2 * M + q2] - Am[p * M + p]) / (T(2) * apq);
              const T tt = (theta >= T(0) ? T(1) : T(-1)) / (abs_(theta) + sqrt(theta * theta + T(1)));
              const T cc = T(1) / sqrt(tt * tt + T(1)), sn = tt * cc;
              for (int k = 0; k < M; k++) { const T x1 = Am[k * M + p], x2 = Am[k * M + q2]; Am[k * M + p] = cc * x1 - sn * x2; Am[k * M + q2] = sn * x1 + cc * x2; }
              for (int k = 0; k < M; k++) { const T x1 = Am[p * M + k], x2 = Am[q2 * M + k]; Am[p * M + k] = cc * x1 - sn * x2; Am[q2 * M + k] = sn * x1 + cc * x2; }
            }
        }
        mn = Am[0];
        for (int i = 1; i < M; i++) mn = Am[i * M + i] < mn ? Am[i * M + i] : mn;
      }
      mn = __shfl_sync(gmask, mn, 0, SW);
      shift = T(IDOC_EPS) - mn;
      if (!(mn == mn) || abs_(mn) > (sizeof(T) == 8 ? T(1e300) : T(1e30))) stat |= ST_NONFINITE;
      shift = shift > T(0) ? shift : T(0);
      if (shift > T(0)) stat |= ST_SHIFT;
      __syncwarp(gmask);
    }
    if constexpr (NW > 1) {  // the factor, its inverse and the shift, from warp 0 to the other warp(s)
      if (threadIdx.x == 0) *sshift = shift;
      __syncthreads();
      shift = *sshift;
    }
    // ---- k = -Gt^{-1} gu, K = -Gt^{-1} Gux
    if constexpr (!TI) {
      if (lane < M) {
        T a0 = T(0), a1 = T(0);  // y = Linv gu: row `lane`, entries c <= lane (unrolled, predicated: no data-dependent trip count)
#pragma unroll
        for (int c = 0; c < M; c += 2) {
          a0 += c <= lane ? sLi[lane * M + c] * sgu[c] : T(0);
          a1 += c + 1 <= lane ? sLi[lane * M + c + 1] * sgu[c + 1] : T(0);
        }
        skk[lane] = a0 + a1;
      }
      gsync();
      T kmine = T(0);
      if (lane < M) {
        T a0 = T(0), a1 = T(0);  // k = -Linv^T y: column `lane`, entries c >= lane
#pragma unroll
        for (int c = 0; c < M; c += 2) {
          a0 += c >= lane ? sLi[c * M + lane] * skk[c] : T(0);
          a1 += c + 1 >= lane ? sLi[(c + 1) * M + lane] * skk[c + 1] : T(0);
        }
        kmine = -(a0 + a1);
      }
      gsync();
      if (lane < M) skk[lane] = kmine;
    }
    if (lane < N) {
      const int j = lane;
      T gc[M], kc[M];
#pragma unroll
      for (int c = 0; c < M; c++) {
        gc[c] = sGux[c * LDG + j];
        kc[c] = T(0);
      }
#pragma unroll
      for (int b = 0; b < M; b++) {
        T Lr[M];
        lds<T, M, align_of(M, (int)sizeof(T))>(sLi + b * M, Lr);
        T acc = T(0);
#pragma unroll
        for (int c = 0; c <= b; c++) acc += Lr[c] * gc[c];
#pragma unroll
        for (int c = 0; c <= b; c++) kc[c] -= Lr[c] * acc;
      }
#pragma unroll
      for (int c = 0; c < M; c++) sK[c * LDK + j] = kc[c];
    }
    gsync();
    if constexpr (!TI) {  // dC, dc with the unshifted Guu (lqr.py:93-94)
      T quad = T(0), lin = T(0);
      if (lane < M) {
        T acc = T(0);
        for (int c = 0; c < M; c++) acc += sGuu[lane * M + c] * skk[c];
        quad = acc * skk[lane];
        lin = sgu[lane] * skk[lane];
      }
      if (NW == 1 || wid == 0) {
        dC += T(0.5) * group_sum<T, SW>(quad, gmask);
        dc += group_sum<T, SW>(lin, gmask);
      }
    }
    // ---- slot: Linv (packed lower), shift, K, k ; user-visible gains
    for (int e = lane; e < M * M; e += LP) {
      const int i = e / M, j = e % M;
      if (j <= i) wsl[Ge::oLinv + tril_idx(i, j)] = sLi[e];
    }
    if (lane == 0) wsl[Ge::oS] = shift;
    if constexpr (LP == N) {  // element lane + N b is (b, lane)
#pragma unroll
      for (int b = 0; b < M; b++) wsl[Ge::oK + b * N + lane] = sK[b * LDK + lane];
    } else {
      for (int e = lane; e < M * N; e += LP) wsl[Ge::oK + e] = sK[(e / N) * LDK + (e % N)];
    }
    if (lane < M) wsl[Ge::ok + lane] = TI ? T(0) : skk[lane];
    if (a.K != nullptr) {
      for (int e = lane; e < M * N; e += LP) a.K[idx * M * N + e] = sK[(e / N) * LDK + (e % N)];
      if (lane < M) a.k[idx * M + lane] = TI ? T(0) : skk[lane];
    }
    // ---- v+ = gx + K^T (gu - s k),  V+ = Gxx + Gux^T K - s K^T K
    T vnew = T(0);
    if constexpr (!TI) {
      if (lane < N) {
        T acc = sgx[lane];
        for (int b = 0; b < M; b++) acc += sK[b * LDK + lane] * (sgu[b] - shift * skk[b]);
        vnew = acc;
      }
    }
    psync();  // K of every problem of the warp is written; every read of sV / sv of this step is done
#pragma unroll
    for (int pg = 0; pg < G; pg++) {
      const T *pGux = s0 + pg * GSTRIDE + C::oGux, *pK = s0 + pg * GSTRIDE + C::oK;
      T* pV = s0 + pg * GSTRIDE + C::oV;
#pragma unroll
      for (int ks = 0; ks < M / TK; ks++) {
        FA fa[MTN];
#pragma unroll
        for (int mt = 0; mt < MTN; mt++) Op::template lda<false>(pGux + (ks * TK) * LDG + mt * TM, 1, LDG, g, tig, fa[mt]);  // Gux^T[i][k] = Gux[k][i]
#pragma unroll
        for (int i = 0; i < NTNW; i++) {
          FB b;
          Op::ldb(pK + (ks * TK) * LDK + (ntn0 + i) * 8, LDK, 1, g, tig, b);
#pragma unroll
          for (int mt = 0; mt < MTN; mt++) Op::mma(gxx[pg][mt][i], fa[mt], b);
        }
      }
      const T sh = TI ? T(0) : (NW > 1 ? shift : __shfl_sync(0xffffffffu, shift, pg * LP));
      if (!TI && sh > T(0)) {  // rare: - s K^T K
        T kk2[MTN][NTNW][CE];
#pragma unroll
        for (int mt = 0; mt < MTN; mt++)
#pragma unroll
          for (int nt = 0; nt < NTNW; nt++)
#pragma unroll
            for (int e = 0; e < CE; e++) kk2[mt][nt][e] = T(0);
#pragma unroll
        for (int ks = 0; ks < M / TK; ks++) {
          FA fa[MTN];
#pragma unroll
          for (int mt = 0; mt < MTN; mt++) Op::template lda<false>(pK + (ks * TK) * LDK + mt * TM, 1, LDK, g, tig, fa[mt]);
#pragma unroll
          for (int i = 0; i < NTNW; i++) {
            FB b;
            Op::ldb(pK + (ks * TK) * LDK + (ntn0 + i) * 8, LDK, 1, g, tig, b);
#pragma unroll
            for (int mt = 0; mt < MTN; mt++) Op::mma(kk2[mt][i], fa[mt], b);
          }
        }
#pragma unroll
        for (int mt = 0; mt < MTN; mt++)
#pragma unroll
          for (int nt = 0; nt < NTNW; nt++)
#pragma unroll
            for (int e = 0; e < CE; e++) gxx[pg][mt][nt][e] -= sh * kk2[mt][nt][e];
      }
      // The owner of an upper-triangle entry writes it AND its mirror image, so V stays exactly symmetric (the reference
      // symmetrises V; the slot keeps the upper triangle).  Row-wise store: bank 20 g + 2 tig (2-way); mirrored store:
      // bank 8 tig + g (+ LDV e): conflict-free, which is what replaces the separate mirror pass of the FFMA sweep.
#pragma unroll
      for (int mt = 0; mt < MTN; mt++)
#pragma unroll
        for (int i = 0; i < NTNW; i++)
#pragma unroll
          for (int e = 0; e < CE; e++) {
            const int r = mt * TM + g + (e >= 2 ? 8 : 0), c = (ntn0 + i) * 8 + 2 * tig + (e & 1);
            if (r <= c) {
              pV[r * LDV + c] = gxx[pg][mt][i][e];
              if (r < c) pV[c * LDV + r] = gxx[pg][mt][i][e];
            }
          }
    }
    if constexpr (!TI) {
      if (lane < N) sv[lane] = vnew;
    }
    psync();
  }
  if (lane == 0) {
    if (a.dCdc != nullptr) { a.dCdc[prob * 2] = dC; a.dCdc[prob * 2 + 1] = dc; }
    if (a.status != nullptr) a.status[prob] = stat;
    a.wstatus[prob] = stat;
  }
}

}  // namespace idoc
