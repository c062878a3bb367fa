// Mid-size Riccati sweep: one warp per problem, compile-time (N, M) with N, M multiples of 8 (the BASELINE shapes
// n=16,m=8 (tilqr, config 3) and n=32,m=16 (config 5)), register-tiled matrix products over shared-memory operands.
// Drop-in for gen_riccati_kernel (csrc/generic.cuh): same arguments, same residual-slot layout, same status bits, so
// the rollout / adjoint kernels and every host path are unchanged.  Reference: idoc/lqr.py:77-95 (time-varying) and
// idoc/tilqr.py:56-78 (`ti`: A, B, Q, R indexed by problem only, q = r = d = M = 0, no eigen-shift).
//
// The 32 lanes form a 4 x 8 grid; for an R x C output, lane (lr, lc) owns the (R/4) x (C/8) tile at rows lr*R/4, columns
// lc*C/8, and every product is an outer-product accumulation over k whose two operand slices are CONTIGUOUS pieces of
// row k of a row-major shared-memory matrix (vector loads, broadcast across the lanes that share lr or lc, no bank
// conflicts):
//     W   = V [A B]                       rows of V (symmetric: row k = column k)  x  rows of F = [A | B]
//     Gxx = sym(Q) + A^T W_A,  Gux = M^T + B^T W_A,  Guu = sym(R) + B^T W_B          rows of F  x  rows of W
//     V+  = Gxx + Gux^T K - s K^T K       rows of Gux (stored transposed, M x N)  x  rows of K
// Gxx never leaves the registers of the lane that owns its tile.  The m x m Cholesky runs with lanes owning rows, the
// triangular solves with lanes owning right-hand-side columns.
#pragma once
#include "generic.cuh"

namespace idoc {

// TI: time-invariant problems (tilqr) have no M, q, r, d: their buffers and the arithmetic on them are compiled out (at
// n = 16, m = 8 that is what lets a 14th CTA fit an SM)
template <typename T, int N, int M, int LP = 32, bool TI = false>
struct MidCfg {
  static_assert(N % 8 == 0 && M % 8 == 0 && M <= 32 && N <= 32, "mid kernels: n, m multiples of 8, at most 32");
  using Ge = Geo<T, N, M>;
  static constexpr int F = N + M;
  static_assert(LP == 32 || LP == 16, "lanes per problem");
  static constexpr int G = 32 / LP, LC = LP / 4;  // the LP lanes of a problem form a 4 x LC grid
  static constexpr int TRN = N / 4, TCN = N / LC, TCF = F / LC, TRM = M / 4, TCM = M / LC;
  // shared-memory layout (elements)
  static constexpr int oF = 0;                 // [N][F]  A | B
  static constexpr int oQ = oF + N * F;        // [N][N]
  static constexpr int oMx = oQ + N * N;       // [N][M]
  static constexpr int oR = oMx + (TI ? 0 : N * M);  // [M][M]
  static constexpr int oq = oR + M * M;               // q[N] r[M] d[N]
  static constexpr int or_ = oq + (TI ? 0 : N);
  static constexpr int od = or_ + (TI ? 0 : M);
  static constexpr int oV = od + (TI ? 0 : N);        // [N][N]
  static constexpr int ov = oV + N * N;        // v[N] w[N]
  static constexpr int ow = ov + N;
  static constexpr int oW = ow + N;            // [N][F]
  static constexpr int oGux = oW + N * F;      // [M][N]
  // W is dead once the G blocks are formed: Guu, its working copy, the inverse factor, K (and the M x M scratch of the
  // eigen-shift path) live in its range when they fit (one barrier before the first of them is written)
  static constexpr bool ALIAS = N * F >= 4 * M * M + M * N;
  static constexpr int oGuu = ALIAS ? oW : oGux + M * N;  // [M][M] (symmetrised, unshifted)
  static constexpr int oGt = oGuu + M * M;     // [M][M] working copy -> Cholesky factor
  static constexpr int oLi = oGt + M * M;      // [M][M] inverse factor (lower)
  static constexpr int oK = oLi + M * M;       // [M][N]
  static constexpr int oJac = ALIAS ? oK + M * N : oW;    // [M][M] Jacobi scratch of the eigen-shift path
  static constexpr int ogx = ALIAS ? oGux + M * N : oK + M * N;  // gx[N] gu[M] k[M]
  static constexpr int ogu = ogx + N;
  static constexpr int okk = ogu + M;
  static constexpr int ELEMS = okk + M + 8;
  // per-problem stride: 16 mod 32 words, so that the two problems of a warp (LP = 16) sit on different bank halves
  static constexpr int GSTRIDE = (ELEMS + 31) / 32 * 32 + (G > 1 ? 16 : 0);
};

// acc[TR][TC] += sum_k a_k[rows] (x) b_k[cols]; A-operand row k at pa + k*lda (TR contiguous), B-operand at pb + k*ldb
template <typename T, int TR, int TC, int KD>
IDOC_DEVICE void tile_mac(T (&acc)[TR][TC], const T* pa, int lda, const T* pb, int ldb) {
#pragma unroll 8
  for (int k = 0; k < KD; k++) {
    T av[TR], bv[TC];
    lds<T, TR, align_of(TR, (int)sizeof(T))>(pa + k * lda, av);
    lds<T, TC, align_of(TC, (int)sizeof(T))>(pb + k * ldb, bv);
#pragma unroll
    for (int i = 0; i < TR; i++) {
      if constexpr (TC % 2 == 0) {  // fp32: one packed FMA (FFMA2) per two columns, common.cuh
#pragma unroll
        for (int j = 0; j < TC; j += 2) fma_pair<T>(acc[i][j], acc[i][j + 1], bv[j], bv[j + 1], av[i]);
      } else {
#pragma unroll
        for (int j = 0; j < TC; j++) acc[i][j] += av[i] * bv[j];
      }
    }
  }
}

// 16-byte cp.async of `cnt` contiguous elements (a multiple of the 16-byte chunk)
template <typename T, int LP>
IDOC_DEVICE void mid_cp_range16(T* dst, const T* src, int cnt, int lane) {
  constexpr int E = 16 / (int)sizeof(T);
  for (int c = lane; c < cnt / E; c += LP) cp_async<16>(smem_u32(dst + c * E), src + c * E);
}
// plain copy (or zero fill when src is null) by the LP lanes of a problem
template <typename T, int LP>
IDOC_DEVICE void mid_g2s(T* dst, const T* src, int cnt, int lane) {
  if (src != nullptr)
    for (int e = lane; e < cnt; e += LP) dst[e] = src[e];
  else
    for (int e = lane; e < cnt; e += LP) dst[e] = T(0);
}
template <typename T, int LP>
IDOC_DEVICE T group_sum(T v, unsigned mask) {
#pragma unroll
  for (int o = LP / 2; o; o >>= 1) v += __shfl_xor_sync(mask, v, o);
  return v;
}

template <typename T>
IDOC_DEVICE T warp_sum(T v) {
  for (int o = 16; o; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// sum_k col[k * ld] * x[k] on four accumulators (a quarter of the dependent-FMA chain of the plain loop); x: shared-memory vector
template <typename T, int K>
IDOC_DEVICE T col_dot4(const T* col, int ld, const T* x, T init) {
  static_assert(K % 4 == 0, "four accumulators");
  T a0 = init, a1 = T(0), a2 = T(0), a3 = T(0);
#pragma unroll
  for (int k = 0; k < K; k += 4) {
    a0 += col[k * ld] * x[k];
    a1 += col[(k + 1) * ld] * x[k + 1];
    a2 += col[(k + 2) * ld] * x[k + 2];
    a3 += col[(k + 3) * ld] * x[k + 3];
  }
  return (a0 + a1) + (a2 + a3);
}

template <typename T, int N, int M, int LP, bool TI>
__global__ void __launch_bounds__(32) mid_riccati_kernel(const GenFwdArgs<T> ga) {
  using C = MidCfg<T, N, M, LP, TI>;
  using Ge = Geo<T, N, M>;
  constexpr int F = C::F, TRN = C::TRN, TCN = C::TCN, TCF = C::TCF, TRM = C::TRM, TCM = C::TCM, WS = Ge::WS;
  extern __shared__ __align__(16) char smem_raw[];
  // `lane` is the lane within the problem's group of LP lanes; a warp owns 32 / LP problems (idle groups of the last warp
  // shadow the last problem: their plain stores rewrite identical values, their atomics are masked)
  const int lane = threadIdx.x % LP, grp = threadIdx.x / LP, lr = lane / C::LC, lc = lane % C::LC;
  T* s = reinterpret_cast<T*>(smem_raw) + grp * C::GSTRIDE;
  const LqrFwdArgs<T>& a = ga.a;
  constexpr bool ti = TI;  // the launcher picks the instantiation from ga.ti
  const int Th = a.T_;
  const long long praw = (long long)blockIdx.x * C::G + grp;
  const long long prob = praw < a.nb ? praw : a.nb - 1;
  // the two problems of a warp (LP = 16) may take different paths through the eigen-shift retry: every barrier and
  // shuffle below is scoped to the problem's own lanes
  const unsigned gmask = LP == 32 ? 0xffffffffu : (0xffffu << (16 * grp));
  T* sF = s + C::oF; T* sQ = s + C::oQ; T* sMx = s + C::oMx; T* sR = s + C::oR; T* sq = s + C::oq; T* sr = s + C::or_;
  T* sd = s + C::od; T* sV = s + C::oV; T* sv = s + C::ov; T* sw = s + C::ow; T* sW = s + C::oW; T* sGux = s + C::oGux;
  T* sGuu = s + C::oGuu; T* sGt = s + C::oGt; T* sLi = s + C::oLi; T* sK = s + C::oK; T* sgx = s + C::ogx; T* sgu = s + C::ogu;
  T* skk = s + C::okk;

  auto load_F = [&](const T* A, const T* B) {
    for (int e = lane; e < N * N; e += LP) sF[(e / N) * F + (e % N)] = A[e];
    for (int e = lane; e < N * M; e += LP) sF[(e / M) * F + N + (e % M)] = B[e];
  };
  auto load_symQ = [&](const T* Q) {  // sQ = 1/2 (Q + Q^T): the transposed operand comes from global (L1/L2), not shared
    for (int e = lane; e < N * N; e += LP) sQ[e] = T(0.5) * (Q[e] + Q[(e % N) * N + e / N]);
  };
  {  // V_T = sym(Qf), v_T = qf
    const T* Qf = a.Qf + prob * N * N;
    for (int e = lane; e < N * N; e += LP) {
      const int i = e / N, j = e % N;
      sV[e] = T(0.5) * (Qf[i * N + j] + Qf[j * N + i]);
    }
    mid_g2s<T, LP>(sv, a.qf ? a.qf + prob * N : (const T*)nullptr, N, lane);
  }
  if (ti) {
    load_F(a.A + prob * N * N, a.B + prob * N * M);
    load_symQ(a.Q + prob * N * N);
    mid_g2s<T, LP>(sR, a.R + prob * M * M, M * M, lane);
  }
  // time-varying leaves: 16-byte cp.async copies (rows of A and B interleaved into F = [A | B]; Q lands unsymmetrised),
  // issued as soon as the previous step has consumed its inputs
  auto issue_tv = [&](int t) {
    constexpr int E = 16 / (int)sizeof(T);
    const long long idx = prob * Th + t;
    for (int c = lane; c < N * N / E; c += LP) cp_async<16>(smem_u32(sF + (c / (N / E)) * F + (c % (N / E)) * E), a.A + idx * N * N + c * E);
    for (int c = lane; c < N * M / E; c += LP) cp_async<16>(smem_u32(sF + (c / (M / E)) * F + N + (c % (M / E)) * E), a.B + idx * N * M + c * E);
    mid_cp_range16<T, LP>(sQ, a.Q + idx * N * N, N * N, lane);
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
  __syncwarp(gmask);

  for (int t = Th - 1; t >= 0; t--) {
    const long long idx = prob * Th + t;
    T* wsl = a.ws + idx * WS;
    if (!ti) {  // the time-varying leaves of this step were requested during the previous one
      cp_async_wait<0>();
      __syncwarp(gmask);
      // symmetrise Q in place along wrapped diagonals: lane l averages (l, l+d) with (l+d, l); both the row-wise and the
      // column-wise access of a diagonal fall on distinct banks (the transposed read in the Gxx tiles was an 8-way
      // conflict).  Every unordered pair belongs to exactly one (lane, d); the barrier after W = V [A B] publishes it.
      static_assert(LP >= N, "one lane per row of Q");
      if (lane < N) {
#pragma unroll 4
        for (int d = 1; d <= N / 2; d++) {
          int c = lane + d;
          c = c >= N ? c - N : c;
          if (d < N / 2 || lane < N / 2) {
            const T m = T(0.5) * (sQ[lane * N + c] + sQ[c * N + lane]);
            sQ[lane * N + c] = m;
            sQ[c * N + lane] = m;
          }
        }
      }
    }
    // slot: the value function entering the step (packed upper), straight to global
    if constexpr (LP == N) {  // lane = column: row i of the packed triangle at a compile-time offset
#pragma unroll
      for (int i = 0; i < N; i++)
        if (i <= lane) wsl[Ge::oVp + (i * N - i * (i - 1) / 2 - i) + lane] = sV[i * N + lane];
    } else {
#pragma unroll 4
      for (int e = lane; e < N * N; e += LP) {
        const int i = e / N, j = e % N;
        if (i <= j) wsl[Ge::oVp + (i * N - i * (i - 1) / 2 - i) + j] = sV[e];
      }
    }
    if (lane < N) wsl[Ge::ov + lane] = TI ? T(0) : sv[lane];
    __syncwarp(gmask);

    // ---- w = v + V d (lanes own entries; column access of the symmetric V is conflict-free).  tilqr (TI) has no affine terms
    // at all (idoc/tilqr.py:12-19: Q, Qf, R, A, B only): v, w, gx, gu, k and the expected-change terms are identically zero
    // and their arithmetic is compiled out
    if constexpr (!TI) {
      if (lane < N) {
        sw[lane] = col_dot4<T, N>(sV + lane, N, sd, sv[lane]);
      }
    }
    // ---- W = V [A B]
    {
      T acc[TRN][TCF];
#pragma unroll
      for (int i = 0; i < TRN; i++)
#pragma unroll
        for (int j = 0; j < TCF; j++) acc[i][j] = T(0);
      tile_mac<T, TRN, TCF, N>(acc, sV + lr * TRN, N, sF + lc * TCF, F);
#pragma unroll
      for (int i = 0; i < TRN; i++)  // one vector store per tile row
        sts<T, TCF, align_of(TCF, (int)sizeof(T))>(sW + (lr * TRN + i) * F + lc * TCF, acc[i]);
    }
    __syncwarp(gmask);
    // ---- G blocks
    T gxx[TRN][TCN];  // stays in registers until V+ is formed
#pragma unroll
    for (int i = 0; i < TRN; i++)
#pragma unroll
      for (int j = 0; j < TCN; j++) {
        const int r = lr * TRN + i, c = lc * TCN + j;  // sQ holds sym(Q)
        gxx[i][j] = sQ[r * N + c];
      }
    tile_mac<T, TRN, TCN, N>(gxx, sF + lr * TRN, F, sW + lc * TCN, F);
    {
      T gux[TRM][TCN];  // Gux[c][i] = M[i][c] + sum_k B[k][c] W_A[k][i]
#pragma unroll
      for (int i = 0; i < TRM; i++)
#pragma unroll
        for (int j = 0; j < TCN; j++) gux[i][j] = TI ? T(0) : sMx[(lc * TCN + j) * M + lr * TRM + i];
      tile_mac<T, TRM, TCN, N>(gux, sF + N + lr * TRM, F, sW + lc * TCN, F);
#pragma unroll
      for (int i = 0; i < TRM; i++) sts<T, TCN, align_of(TCN, (int)sizeof(T))>(sGux + (lr * TRM + i) * N + lc * TCN, gux[i]);
      T guu[TRM][TCM];
#pragma unroll
      for (int i = 0; i < TRM; i++)
#pragma unroll
        for (int j = 0; j < TCM; j++) guu[i][j] = sR[(lr * TRM + i) * M + lc * TCM + j];
      tile_mac<T, TRM, TCM, N>(guu, sF + N + lr * TRM, F, sW + N + lc * TCM, F);
      if constexpr (C::ALIAS) __syncwarp(gmask);  // every lane has read W: its range becomes Guu / factor / K
#pragma unroll
      for (int i = 0; i < TRM; i++) sts<T, TCM, align_of(TCM, (int)sizeof(T))>(sGt + (lr * TRM + i) * M + lc * TCM, guu[i]);  // unsymmetrised
    }
    // gx = q + A^T w, gu = r + B^T w
    if constexpr (!TI) {
      if (lane < N) {
        sgx[lane] = col_dot4<T, N>(sF + lane, F, sw, sq[lane]);
      }
      if (lane < M) {
        sgu[lane] = col_dot4<T, N>(sF + N + lane, F, sw, sr[lane]);
      }
    }
    __syncwarp(gmask);
    if (!ti && t > 0) issue_tv(t - 1);  // F, Q, M, R, q, r, d of this step are dead from here on
    for (int e = lane; e < M * M; e += LP) {
      const int i = e / M, j = e % M;
      sGuu[e] = T(0.5) * (sGt[i * M + j] + sGt[j * M + i]);
    }
    __syncwarp(gmask);
    // ---- Cholesky of Guu (+ shift): lanes own rows; factor in sGt (lower), inverse factor in sLi
    T shift = T(0);
    for (int attempt = 0; attempt < 2; attempt++) {
      // Lane i holds row i of Guu in registers; the right-looking factorisation exchanges the pivot and the column
      // entries with shuffles (no shared memory, no barrier inside the loop).  Entries above the diagonal are garbage.
      T g[M], dinv[M];
      {
        const int row = lane < M ? lane : 0;
        lds<T, M, align_of(M, (int)sizeof(T))>(sGuu + row * M, g);
#pragma unroll
        for (int c = 0; c < M; c++) g[c] += (c == lane) ? shift : T(0);  // select, not a divergent branch per column
      }
      bool ok = true;
#pragma unroll
      for (int j = 0; j < M; j++) {
        T p = __shfl_sync(gmask, g[j], j, LP);
        if (!(p > T(0))) {
          ok = false;
          if (attempt == 1) { p = T(IDOC_EPS) * T(1e-3); stat |= ST_CHOL_CLAMP; }
          else {
            p = T(1);
            // tilqr factors without a shift (sym_pos=True, idoc/tilqr.py:72): a failed pivot has no retry there; the
            // reference's Cholesky solve would return NaN, so the problem is flagged and its outputs set to NaN
            if (ti) stat |= ST_CHOL_CLAMP | ST_NONFINITE;
          }
        }
        const T rs = rsqrt_<T>(p);
        dinv[j] = rs;
        const T lij = (lane == j) ? p * rs : g[j] * rs;  // L[lane][j] for lane >= j
        g[j] = lij;
#pragma unroll
        for (int c = j + 1; c < M; c++) g[c] -= lij * __shfl_sync(gmask, lij, c, LP);  // used for c <= lane only
      }
      if (lane < M) sts<T, M, align_of(M, (int)sizeof(T))>(sGt + lane * M, g);  // rows of L (lower part valid)
      __syncwarp(gmask);
      // column `lane` of the inverse factor: li[i] = -(sum_{c<i} L[i][c] li[c]) / L[i][i], li[c] = 0 for c < lane.
      // Every lane reads the same L entries (broadcast), so the loop is uniform and divergence-free.
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
      fro = group_sum<T, LP>(fro, gmask);
      if (attempt == 1 || ti) break;  // tilqr solves with sym_pos=True and no shift (idoc/tilqr.py:72)
      if (ok && fro < T(1.0 / IDOC_EPS)) break;
      // slow path: lambda_min by Jacobi on lane 0 (scratch: inside the dead W range)
      T mn = T(0);
      if (lane == 0) {
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
      mn = __shfl_sync(gmask, mn, 0, LP);
      shift = T(IDOC_EPS) - mn;
      if (!(mn == mn) || abs_(mn) > T(1e300)) stat |= ST_NONFINITE;  // NaN / Inf in Guu (the reference propagates NaN)
      shift = shift > T(0) ? shift : T(0);
      if (shift > T(0)) stat |= ST_SHIFT;
      __syncwarp(gmask);
    }
    // ---- k = -Gt^{-1} gu (lane 0..M-1 cooperate through shared memory), K = -Gt^{-1} Gux (lanes own columns)
    if constexpr (!TI) {
      if (lane < M) {
        T a0 = T(0), a1 = T(0);  // y = Linv gu: row `lane`, entries c <= lane (unrolled, predicated: no data-dependent trip count)
#pragma unroll
        for (int c = 0; c < M; c += 2) {
          a0 += c <= lane ? sLi[lane * M + c] * sgu[c] : T(0);
          a1 += c + 1 <= lane ? sLi[lane * M + c + 1] * sgu[c + 1] : T(0);
        }
        skk[lane] = a0 + a1;  // y
      }
      __syncwarp(gmask);
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
      __syncwarp(gmask);
      if (lane < M) skk[lane] = kmine;
    }
    if (lane < N) {  // K[:, lane] = -Linv^T (Linv Gux[:, lane]); rows of Linv are broadcast vector loads
      const int j = lane;
      T gc[M], y[M], kc[M];
#pragma unroll
      for (int c = 0; c < M; c++) {
        gc[c] = sGux[c * N + j];
        kc[c] = T(0);
      }
#pragma unroll
      for (int b = 0; b < M; b++) {
        T Lr[M];
        lds<T, M, align_of(M, (int)sizeof(T))>(sLi + b * M, Lr);
        T acc = T(0);
#pragma unroll
        for (int c = 0; c <= b; c++) acc += Lr[c] * gc[c];
        y[b] = acc;
#pragma unroll
        for (int c = 0; c <= b; c++) kc[c] -= Lr[c] * acc;  // K[c][j] = -sum_{b >= c} Linv[b][c] y[b]
      }
#pragma unroll
      for (int c = 0; c < M; c++) sK[c * N + j] = kc[c];
    }
    __syncwarp(gmask);
    if constexpr (!TI) {  // dC, dc with the unshifted Guu (lqr.py:93-94)
      T quad = T(0), lin = T(0);
      if (lane < M) {
        T acc = T(0);
        for (int c = 0; c < M; c++) acc += sGuu[lane * M + c] * skk[c];
        quad = acc * skk[lane];
        lin = sgu[lane] * skk[lane];
      }
      dC += T(0.5) * group_sum<T, LP>(quad, gmask);
      dc += group_sum<T, LP>(lin, gmask);
    }
    // ---- slot: Linv (packed lower), shift, K, k ; user-visible gains
    for (int e = lane; e < M * M; e += LP) {
      const int i = e / M, j = e % M;
      if (j <= i) wsl[Ge::oLinv + tril_idx(i, j)] = sLi[e];
    }
    if (lane == 0) wsl[Ge::oS] = shift;
    for (int e = lane; e < M * N; e += LP) wsl[Ge::oK + e] = sK[e];
    if (lane < M) wsl[Ge::ok + lane] = TI ? T(0) : skk[lane];
    if (a.K != nullptr) {
      for (int e = lane; e < M * N; e += LP) a.K[idx * M * N + e] = sK[e];
      if (lane < M) a.k[idx * M + lane] = TI ? T(0) : skk[lane];
    }
    // ---- v+ = gx + K^T (gu - s k),  V+ = Gxx + Gux^T K - s K^T K
    T vnew = T(0);
    if constexpr (!TI) {
      if (lane < N) {
        T acc = sgx[lane];
        for (int b = 0; b < M; b++) acc += sK[b * N + lane] * (sgu[b] - shift * skk[b]);
        vnew = acc;
      }
    }
    tile_mac<T, TRN, TCN, M>(gxx, sGux + lr * TRN, N, sK + lc * TCN, N);
    if (shift > T(0)) {
      T kk2[TRN][TCN];
#pragma unroll
      for (int i = 0; i < TRN; i++)
#pragma unroll
        for (int j = 0; j < TCN; j++) kk2[i][j] = T(0);
      tile_mac<T, TRN, TCN, M>(kk2, sK + lr * TRN, N, sK + lc * TCN, N);
#pragma unroll
      for (int i = 0; i < TRN; i++)
#pragma unroll
        for (int j = 0; j < TCN; j++) gxx[i][j] -= shift * kk2[i][j];
    }
    __syncwarp(gmask);  // every read of sV / sv of this step is done
    // the owner of an upper-triangle entry writes it and its mirror image (the reference symmetrises V; the slot keeps
    // the upper triangle), so V stays exactly symmetric without a second pass
#pragma unroll
    for (int i = 0; i < TRN; i++)
#pragma unroll
      for (int j = 0; j < TCN; j++) {
        const int r = lr * TRN + i, c = lc * TCN + j;
        if (r <= c) sV[r * N + c] = gxx[i][j];
      }
    if constexpr (!TI) {
      if (lane < N) sv[lane] = vnew;
    }
    __syncwarp(gmask);
    // mirror image along diagonals (lane l copies (l, l+d) to (l+d, l): distinct banks on both sides; the column-wise
    // store straight from the tiles was an 8-way conflict)
    if (lane < N) {
#pragma unroll 4
      for (int d = 1; d < N; d++) {
        const int c = lane + d;
        if (c < N) sV[c * N + lane] = sV[lane * N + c];
      }
    }
    __syncwarp(gmask);
  }
  if (lane == 0) {
    if (a.dCdc != nullptr) { a.dCdc[prob * 2] = dC; a.dCdc[prob * 2 + 1] = dc; }
    if (a.status != nullptr) a.status[prob] = stat;
    a.wstatus[prob] = stat;
  }
}

}  // namespace idoc

// ================================================================================================ vector sweeps
// Rollout and adjoint sweeps for the same shapes: one warp per problem, compile-time sizes, the same arithmetic and
// slot / argument conventions as gen_rollout_kernel / gen_adj_bwd_kernel / gen_adj_fwd_kernel (csrc/generic.cuh).
// A and B sit in shared memory with ODD row strides (n+1, m+1), so that both "lane owns a row" (A x) and "lane owns a
// column" (A^T w) products are bank-conflict free; K x (K row-major in the slot) is split over 32/m lanes per row and
// reduced with shuffles; the time-invariant gradient sums (tilqr) stay in registers across the whole sweep.
namespace idoc {

// TI: time-invariant problems (tilqr) keep the rows / columns of A and B that a lane multiplies with in REGISTERS for the whole
// sweep: no A | B block in shared memory and no operand loads for it.  NST prefetch stages: step t computes out of one while
// the cp.async copies of steps t+1 .. t+NST-1 are in flight.  Measured (profiles/r2_mid_nst_ab.jsonl): more than two stages
// is SLOWER at every shape (the sweeps are bound by shared-memory wavefronts and resident warps, not by bytes in flight).
template <typename T, int N, int M, int LP = 32, bool TI = false, int NST = 2>
struct MidVec {
  using Ge = Geo<T, N, M>;
  static_assert(NST >= 2 && NST <= 4, "prefetch stages");
  // row strides of the shared-memory A and B: one 16-byte chunk of padding, so that rows start 16-byte aligned (16-byte cp.async
  // copies, vector loads of a lane's own row) AND the eight lanes of a quarter warp reading their rows 16 bytes at a time fall
  // on eight different bank groups (stride = 4 mod 8 words); column reads (lane = column) are stride-1 and conflict-free anyway
  static constexpr int E16 = 16 / (int)sizeof(T);
  static constexpr int LDA = N + E16, LDB = M + E16;
  static constexpr int G = 32 / LP;
  static constexpr int LPR = LP / M;   // lanes per row of K
  static constexpr int KE = N / LPR;   // contiguous K elements per lane
  static_assert(LP % M == 0 && N % LPR == 0, "K x split");
  static constexpr int VL = N > M ? N : M;
  // RAB: A, B rows / columns in registers (time-invariant problems, where the register budget allows: every fp32 shape, the
  // smaller fp64 ones); otherwise time-invariant problems hold ONE A | B block in shared memory, in front of the stages
  static constexpr bool RAB = TI && (sizeof(T) == 4 || N * M <= 192);
  static constexpr int AB = (N * LDA + N * LDB + 3) / 4 * 4;  // A | B
  static constexpr int oB = N * LDA;                           // B inside the A | B block
  // one prefetch stage: everything a step reads from global memory
  static constexpr int oSlot = TI ? 0 : AB;               // time-varying: A | B is part of every stage
  static constexpr int oS2 = oSlot + Ge::WS;
  static constexpr int oPv = oS2 + Ge::WS2;               // 5 prefetched vectors
  static constexpr int STAGE = (oPv + 5 * VL + 3) / 4 * 4;
  static constexpr int oSt = (TI && !RAB) ? AB : 0;       // first stage
  static constexpr int oVec = oSt + NST * STAGE;          // working vectors
  static constexpr int ELEMS = oVec + 10 * VL + 8;
  static constexpr int GSTRIDE = (ELEMS + 31) / 32 * 32 + (G > 1 ? 16 : 0);
};

// element-wise cp.async of a row-major R x C matrix into a shared-memory matrix with row stride LD
template <typename T, int R, int Cn, int LD, int LP>
IDOC_DEVICE void mid_cp_matrix(T* dst, const T* src, int lane) {
#pragma unroll 4
  for (int e = lane; e < R * Cn; e += LP) cp_async<(int)sizeof(T)>(smem_u32(dst + (e / Cn) * LD + (e % Cn)), src + e);
}
// 16-byte cp.async of a row-major R x C matrix (C a multiple of the 16-byte chunk) into rows of stride LD
template <typename T, int R, int Cn, int LD, int LP>
IDOC_DEVICE void mid_cp_matrix16(T* dst, const T* src, int lane) {
  constexpr int E = 16 / (int)sizeof(T), CPR = Cn / E;
  static_assert(Cn % E == 0 && LD % E == 0, "16-byte rows");
#pragma unroll 4
  for (int c = lane; c < R * CPR; c += LP) cp_async<16>(smem_u32(dst + (c / CPR) * LD + (c % CPR) * E), src + c * E);
}
template <typename T, int N, int M, int LP>
IDOC_DEVICE void mid_cp_AB(T* ab, const T* A, const T* B, int lane) {
  constexpr int E = 16 / (int)sizeof(T);
  mid_cp_matrix16<T, N, N, N + E, LP>(ab, A, lane);
  mid_cp_matrix16<T, N, M, M + E, LP>(ab + N * (N + E), B, lane);
}
// 16-byte cp.async of elements [lo, hi) (multiples of the 16-byte chunk) of a contiguous global record
template <typename T, int LP>
IDOC_DEVICE void mid_cp_range(T* dst, const T* src, int lo, int hi, int lane) {
  constexpr int E = 16 / (int)sizeof(T);
  for (int c = lo / E + lane; c < hi / E; c += LP) cp_async<16>(smem_u32(dst + c * E), src + c * E);
}
template <typename T>
IDOC_DEVICE void mid_cp_elem(T* dst, const T* src) {
  cp_async<(int)sizeof(T)>(smem_u32(dst), src);
}
// 16-byte cp.async of CNT contiguous elements, CNT / E <= LP chunks: one chunk per lane (the per-step vectors -- U, Nu, X, the
// cotangents -- are 16-byte aligned rows of 16-byte aligned arrays: n, m multiples of 8, base pointers checked at the C ABI)
template <typename T, int CNT, int LP>
IDOC_DEVICE void mid_cp_vec(T* dst, const T* src, int lane) {
  constexpr int E = 16 / (int)sizeof(T);
  static_assert(CNT % E == 0 && CNT / E <= LP, "one 16-byte chunk per lane");
  if (lane < CNT / E) cp_async<16>(smem_u32(dst + lane * E), src + lane * E);
}
// mid_cp_range with compile-time bounds: fully unrolled, no loop-carried pointer arithmetic
template <typename T, int LO, int HI, int LP>
IDOC_DEVICE void mid_cp_range_c(T* dst, const T* src, int lane) {
  constexpr int E = 16 / (int)sizeof(T), C0 = LO / E, C1 = HI / E, NCH = C1 - C0;
  static_assert(LO % E == 0 && HI % E == 0, "16-byte bounds");
#pragma unroll
  for (int i = 0; i < (NCH + LP - 1) / LP; i++) {
    const int c = C0 + i * LP + lane;
    if ((i + 1) * LP <= NCH || c < C1) cp_async<16>(smem_u32(dst + c * E), src + c * E);
  }
}
// (K x)[b] for b = lane / LPR, valid in every lane of the row's group
template <typename T, int N, int M, int LP>
IDOC_DEVICE T mid_kx(const T* K, const T* x, int lane) {
  using V = MidVec<T, N, M, LP>;
  const int b = lane / V::LPR, c0 = (lane % V::LPR) * V::KE;
  T kv[V::KE], xv[V::KE];
  lds<T, V::KE, align_of(V::KE, (int)sizeof(T))>(K + b * N + c0, kv);
  lds<T, V::KE, align_of(V::KE, (int)sizeof(T))>(x + c0, xv);
  T acc = T(0);
#pragma unroll
  for (int c = 0; c < V::KE; c++) acc += kv[c] * xv[c];
#pragma unroll
  for (int o = V::LPR / 2; o; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
  return acc;
}
// sum_j V(i, j) x[j] with V symmetric, packed upper row-major
template <typename T, int N>
IDOC_DEVICE T mid_symdot(const T* Vp, const T* x, int i) {
  T acc = T(0);
  for (int j = 0; j < i; j++) acc += Vp[j * N - j * (j - 1) / 2 + (i - j)] * x[j];
  const int ro = i * N - i * (i - 1) / 2 - i;
  for (int j = i; j < N; j++) acc += Vp[ro + j] * x[j];
  return acc;
}

// The same product with x in registers, fully unrolled and branch-free: for r <= i the entry V(r, i) sits at off(r) + i - r
// (consecutive lanes read consecutive words of row r), for r > i the entry V(i, r) at off(i) + r - i (lane-dependent row; at
// n = 16 the 16 row offsets fall on 16 different banks).  The loop-carried index arithmetic and the data-dependent trip counts
// of mid_symdot made it 58 % of the rollout's instructions (profiles/r2_mid16_rollout_lines.txt).
template <typename T, int N>
IDOC_DEVICE T mid_symdot_r(const T* Vp, const T (&x)[N], int i) {
  const T* pa = Vp + i;
  const T* pb = Vp + (i * N - i * (i - 1) / 2 - i);
  T acc0 = T(0), acc1 = T(0);
#pragma unroll
  for (int r = 0; r < N; r++) {
    const T v = (i >= r) ? pa[r * N - r * (r - 1) / 2 - r] : pb[r];
    if (r & 1) acc1 += v * x[r];
    else acc0 += v * x[r];
  }
  return acc0 + acc1;
}
// sum_j a[j] x[j] with two accumulators (halves the dependent chain)
template <typename T, int K>
IDOC_DEVICE T mid_dot_r(const T (&a)[K], const T (&x)[K], T init) {
  T acc0 = init, acc1 = T(0);
#pragma unroll
  for (int j = 0; j < K; j++) {
    if (j & 1) acc1 += a[j] * x[j];
    else acc0 += a[j] * x[j];
  }
  return acc0 + acc1;
}
template <typename T, int K>
IDOC_DEVICE void mid_vload(const T* p, T (&r)[K]) {  // p: 16-byte aligned shared-memory vector
  lds<T, K, 16>(p, r);
}

template <typename T, int N, int M, int LP, bool TI, int NST>
__global__ void __launch_bounds__(32) mid_rollout_kernel(const GenFwdArgs<T> ga) {
  using V = MidVec<T, N, M, LP, TI, NST>;
  using Ge = Geo<T, N, M>;
  extern __shared__ __align__(16) char smem_raw[];
  const int lane = threadIdx.x % LP, grp = threadIdx.x / LP;
  T* s = reinterpret_cast<T*>(smem_raw) + grp * V::GSTRIDE;
  const LqrFwdArgs<T>& a = ga.a;
  const int Th = a.T_;
  const long long praw = (long long)blockIdx.x * V::G + grp;
  const long long prob = praw < a.nb ? praw : a.nb - 1;
  T* x = s + V::oVec; T* xn = x + V::VL; T* u = xn + V::VL;
  auto issue = [&](int t, int st) {  // everything step t reads: A, B, d (time-varying) and slot[K..V]
    if (t < Th) {
      T* g = s + V::oSt + st * V::STAGE;
      const long long idx = prob * Th + t;
      if constexpr (!TI) {
        mid_cp_AB<T, N, M, LP>(g, a.A + idx * N * N, a.B + idx * N * M, lane);
        if (a.d != nullptr) mid_cp_vec<T, N, LP>(g + V::oPv, a.d + idx * N, lane);
      }
      mid_cp_range_c<T, Ge::oK, Ge::WS, LP>(g + V::oSlot, a.ws + idx * Ge::WS, lane);
    }
    cp_async_commit();  // one group per step, empty past the horizon: the wait below counts groups
  };
  if (lane < N) {
    x[lane] = a.x0[prob * N + lane];
#pragma unroll
    for (int st = 0; st < NST; st++) s[V::oSt + st * V::STAGE + V::oPv + lane] = T(0);  // d = 0 unless loaded
  }
  T Ar[V::RAB ? N : 1], Br[V::RAB ? M : 1];  // time-invariant: row `lane` of A and of B
  if constexpr (V::RAB) {
    const int row = lane < N ? lane : 0;
#pragma unroll
    for (int j = 0; j < N; j++) Ar[j] = a.A[prob * N * N + row * N + j];
#pragma unroll
    for (int b = 0; b < M; b++) Br[b] = a.B[prob * N * M + row * M + b];
  } else if constexpr (TI) {
    mid_cp_AB<T, N, M, LP>(s, a.A + prob * N * N, a.B + prob * N * M, lane);  // joins the first group
  }
#pragma unroll
  for (int st = 0; st < NST - 1; st++) issue(st, st);
  int cur = 0;
  for (int t = 0; t < Th; t++) {
    const long long idx = prob * Th + t;
    const T* g = s + V::oSt + cur * V::STAGE;
    const T *sA = TI ? s : g, *sB = sA + V::oB, *slot = g + V::oSlot, *sd = g + V::oPv;
    cp_async_wait<NST - 2>();
    __syncwarp();
    {
      const int nx = cur == 0 ? NST - 1 : cur - 1;  // the stage step t - 1 computed out of
      issue(t + NST - 1, nx);
    }
    {  // u = K x + k                                                          lqr.py:159
      const T ku = mid_kx<T, N, M, LP>(slot + Ge::oK, x, lane);
      if (lane % V::LPR == 0) {
        const int b = lane / V::LPR;
        const T uv = ku + slot[Ge::ok + b];
        u[b] = uv;
        a.U[idx * M + b] = uv;
      }
    }
    __syncwarp();
    if (lane < N) {  // x' = A x + B u + d                                      lqr.py:160
      T xr[N], ur[M];
      mid_vload<T, N>(x, xr);
      mid_vload<T, M>(u, ur);
      T acc;
      if constexpr (V::RAB) {
        acc = mid_dot_r<T, N>(Ar, xr, sd[lane]) + mid_dot_r<T, M>(Br, ur, T(0));
      } else {
        T ar[N], br[M];
        mid_vload<T, N>(sA + lane * V::LDA, ar);  // the lane's own rows: vector loads
        mid_vload<T, M>(sB + lane * V::LDB, br);
        acc = mid_dot_r<T, N>(ar, xr, sd[lane]) + mid_dot_r<T, M>(br, ur, T(0));
      }
      xn[lane] = acc;
      a.X[idx * N + lane] = acc;
    }
    __syncwarp();
    if (lane < N) {  // Nu[t] = V_{t+1} x_{t+1} + v_{t+1}
      T xr[N];
      mid_vload<T, N>(xn, xr);
      a.Nu[idx * N + lane] = slot[Ge::ov + lane] + mid_symdot_r<T, N>(slot + Ge::oVp, xr, lane);
      x[lane] = xn[lane];
    }
    cur = cur + 1 == NST ? 0 : cur + 1;
  }
}

// adjoint backward sweep: v', k'   (ws2 slot t = [k'_t | v'_{t+1}])
template <typename T, int N, int M, int LP, bool TI, int NST>
__global__ void __launch_bounds__(32) mid_adj_bwd_kernel(const GenVjpArgs<T> ga) {
  using V = MidVec<T, N, M, LP, TI, NST>;
  using Ge = Geo<T, N, M>;
  extern __shared__ __align__(16) char smem_raw[];
  const int lane = threadIdx.x % LP, grp = threadIdx.x / LP;
  T* s = reinterpret_cast<T*>(smem_raw) + grp * V::GSTRIDE;
  const LqrVjpArgs<T>& a = ga.a;
  const bool has_nub = a.Nub != nullptr;
  const int Th = a.T_;
  const long long praw = (long long)blockIdx.x * V::G + grp;
  const long long prob = praw < a.nb ? praw : a.nb - 1;
  T* vp = s + V::oVec; T* w = vp + V::VL; T* gx = w + V::VL; T* gu = gx + V::VL; T* yg = gu + V::VL; T* kk = yg + V::VL;
  auto issue = [&](int t, int st) {  // A, B (time-varying), slot[Linv s K] (+ V with a Nu cotangent), Ub[t], Xb[t-1], Nub[t]
    if (t >= 0) {
      T* g = s + V::oSt + st * V::STAGE;
      const long long idx = prob * Th + t;
      if constexpr (!TI) mid_cp_AB<T, N, M, LP>(g, a.A + idx * N * N, a.B + idx * N * M, lane);
      if (has_nub) mid_cp_range_c<T, 0, Ge::WS, LP>(g + V::oSlot, a.ws + idx * Ge::WS, lane);
      else mid_cp_range_c<T, 0, Ge::ok, LP>(g + V::oSlot, a.ws + idx * Ge::WS, lane);
      mid_cp_vec<T, M, LP>(g + V::oPv, a.Ub + idx * M, lane);
      if (t > 0) mid_cp_vec<T, N, LP>(g + V::oPv + V::VL, a.Xb + (idx - 1) * N, lane);
      if (has_nub) mid_cp_vec<T, N, LP>(g + V::oPv + 2 * V::VL, a.Nub + idx * N, lane);
    }
    cp_async_commit();
  };
  if (lane < N) vp[lane] = a.Xb[(prob * Th + Th - 1) * N + lane];  // v'_T = qf' = Xb[T-1]
  T Ac[V::RAB ? N : 1], Bc[V::RAB ? N : 1];  // time-invariant: column `lane` of A and of B
  if constexpr (V::RAB) {
    const int ca = lane < N ? lane : 0, cb = lane < M ? lane : 0;
#pragma unroll
    for (int k = 0; k < N; k++) { Ac[k] = a.A[prob * N * N + k * N + ca]; Bc[k] = a.B[prob * N * M + k * M + cb]; }
  } else if constexpr (TI) {
    mid_cp_AB<T, N, M, LP>(s, a.A + prob * N * N, a.B + prob * N * M, lane);
  }
#pragma unroll
  for (int st = 0; st < NST - 1; st++) issue(Th - 1 - st, st);
  int cur = 0;
  for (int t = Th - 1; t >= 0; t--) {
    const long long idx = prob * Th + t;
    const T* g = s + V::oSt + cur * V::STAGE;
    const T *sA = TI ? s : g, *sB = sA + V::oB, *slot = g + V::oSlot, *ub = g + V::oPv, *xb = ub + V::VL, *nb = xb + V::VL;
    cp_async_wait<NST - 2>();
    __syncwarp();
    {
      const int nx = cur == 0 ? NST - 1 : cur - 1;
      issue(t - (NST - 1), nx);
    }
    T* out = a.ws2 + idx * Ge::WS2;
    if (lane < N) {
      out[Ge::o2v + lane] = vp[lane];
      T wv = vp[lane];
      if (has_nub) {
        T nr[N];
        mid_vload<T, N>(nb, nr);
        wv += mid_symdot_r<T, N>(slot + Ge::oVp, nr, lane);
      }
      w[lane] = wv;
    }
    __syncwarp();
    {
      T wr[N];
      mid_vload<T, N>(w, wr);
      if (lane < N) {  // gx' = q' + A^T w
        const T init = t > 0 ? xb[lane] : T(0);
        if constexpr (V::RAB) {
          gx[lane] = mid_dot_r<T, N>(Ac, wr, init);
        } else {
          T ac[N];
#pragma unroll
          for (int k = 0; k < N; k++) ac[k] = sA[k * V::LDA + lane];
          gx[lane] = mid_dot_r<T, N>(ac, wr, init);
        }
      }
      if (lane < M) {  // gu' = r' + B^T w
        if constexpr (V::RAB) {
          gu[lane] = mid_dot_r<T, N>(Bc, wr, ub[lane]);
        } else {
          T bc[N];
#pragma unroll
          for (int k = 0; k < N; k++) bc[k] = sB[k * V::LDB + lane];
          gu[lane] = mid_dot_r<T, N>(bc, wr, ub[lane]);
        }
      }
    }
    __syncwarp();
    if (lane < M) {  // y = Linv gu' (row `lane` of the packed lower factor; unrolled, predicated)
      const T* row = slot + Ge::oLinv + tril_idx(lane, 0);
      T a0 = T(0), a1 = T(0);
#pragma unroll
      for (int c = 0; c < M; c += 2) {
        a0 += c <= lane ? row[c] * gu[c] : T(0);
        a1 += c + 1 <= lane ? row[c + 1] * gu[c + 1] : T(0);
      }
      yg[lane] = a0 + a1;
    }
    __syncwarp();
    if (lane < M) {
      T a0 = T(0), a1 = T(0);  // k' = -Linv^T y (column `lane`)
#pragma unroll
      for (int c = 0; c < M; c += 2) {
        a0 += c >= lane ? slot[Ge::oLinv + tril_idx(c, 0) + lane] * yg[c] : T(0);
        a1 += c + 1 >= lane ? slot[Ge::oLinv + tril_idx(c + 1, 0) + lane] * yg[c + 1] : T(0);
      }
      const T acc = a0 + a1;
      kk[lane] = -acc;
      out[Ge::o2k + lane] = -acc;
    }
    __syncwarp();
    const T shift = slot[Ge::oS];
    T vnew = T(0);
    if (lane < N) {
      T acc = gx[lane];
#pragma unroll 8
      for (int b = 0; b < M; b++) acc += slot[Ge::oK + b * N + lane] * (gu[b] - shift * kk[b]);
      vnew = acc;
    }
    __syncwarp();
    if (lane < N) vp[lane] = vnew;
    cur = cur + 1 == NST ? 0 : cur + 1;
  }
}

// adjoint forward sweep + gradients
template <typename T, int N, int M, int LP, bool TI, int NST>
__global__ void __launch_bounds__(32) mid_adj_fwd_kernel(const GenVjpArgs<T> ga) {
  using V = MidVec<T, N, M, LP, TI, NST>;
  using Ge = Geo<T, N, M>;
  extern __shared__ __align__(16) char smem_raw[];
  const int lane = threadIdx.x % LP, grp = threadIdx.x / LP;
  T* s = reinterpret_cast<T*>(smem_raw) + grp * V::GSTRIDE;
  const LqrVjpArgs<T>& a = ga.a;
  constexpr bool ti = TI;
  const bool reduce = a.reduce_batch != 0, has_nub = a.Nub != nullptr;
  const int Th = a.T_;
  const long long praw = (long long)blockIdx.x * V::G + grp;
  const bool active = praw < a.nb;
  const long long prob = active ? praw : a.nb - 1;
  T* ux = s + V::oVec; T* uxn = ux + V::VL; T* unu = uxn + V::VL; T* uU = unu + V::VL;
  constexpr int EA = TI ? N * N / LP : 1, EB = TI ? N * M / LP : 1, ER = TI ? (M * M + LP - 1) / LP : 1;
  T accA[EA], accQ[EA], accB[EB], accR[ER];  // time sums of the time-invariant leaves: element e = lane + 32 k
#pragma unroll
  for (int k = 0; k < EA; k++) { accA[k] = T(0); accQ[k] = T(0); }
#pragma unroll
  for (int k = 0; k < EB; k++) accB[k] = T(0);
#pragma unroll
  for (int k = 0; k < ER; k++) accR[k] = T(0);
  auto issue = [&](int t, int st) {  // A, B (time-varying), slot[K..V], ws2, U[t], Nu[t], Nub[t], sX[t] = X[t-1] (x0 at 0)
    if (t < Th) {
      T* g = s + V::oSt + st * V::STAGE;
      const long long idx = prob * Th + t;
      if constexpr (!TI) mid_cp_AB<T, N, M, LP>(g, a.A + idx * N * N, a.B + idx * N * M, lane);
      mid_cp_range_c<T, Ge::oK, Ge::WS, LP>(g + V::oSlot, a.ws + idx * Ge::WS, lane);
      mid_cp_range_c<T, 0, Ge::WS2, LP>(g + V::oS2, a.ws2 + idx * Ge::WS2, lane);
      mid_cp_vec<T, M, LP>(g + V::oPv, a.U + idx * M, lane);
      mid_cp_vec<T, N, LP>(g + V::oPv + V::VL, a.Nu + idx * N, lane);
      if (has_nub) mid_cp_vec<T, N, LP>(g + V::oPv + 2 * V::VL, a.Nub + idx * N, lane);
      mid_cp_vec<T, N, LP>(g + V::oPv + 3 * V::VL, t > 0 ? a.X + (idx - 1) * N : a.x0 + prob * N, lane);
    }
    cp_async_commit();
  };
  if (lane < N) {
    ux[lane] = T(0);
#pragma unroll
    for (int st = 0; st < NST; st++) s[V::oSt + st * V::STAGE + V::oPv + 2 * V::VL + lane] = T(0);  // d' = 0 unless a Nu cotangent is given
  }
  T Ar[V::RAB ? N : 1], Br[V::RAB ? M : 1];  // time-invariant: row `lane` of A and of B
  if constexpr (V::RAB) {
    const int row = lane < N ? lane : 0;
#pragma unroll
    for (int j = 0; j < N; j++) Ar[j] = a.A[prob * N * N + row * N + j];
#pragma unroll
    for (int b = 0; b < M; b++) Br[b] = a.B[prob * N * M + row * M + b];
  } else if constexpr (TI) {
    mid_cp_AB<T, N, M, LP>(s, a.A + prob * N * N, a.B + prob * N * M, lane);  // joins the first group
  }
#pragma unroll
  for (int st = 0; st < NST - 1; st++) issue(st, st);
  int cur = 0;
  for (int t = 0; t < Th; t++) {
    const long long idx = prob * Th + t;
    const T* g = s + V::oSt + cur * V::STAGE;
    const T *sA = TI ? s : g, *sB = sA + V::oB, *slot = g + V::oSlot, *s2 = g + V::oS2;
    const T *Uv = g + V::oPv, *nu = Uv + V::VL, *dp = nu + V::VL, *sX = dp + V::VL;
    cp_async_wait<NST - 2>();
    __syncwarp();
    {
      const int nx = cur == 0 ? NST - 1 : cur - 1;
      issue(t + NST - 1, nx);
    }
    {  // uU = K ux + k'
      const T ku = mid_kx<T, N, M, LP>(slot + Ge::oK, ux, lane);
      if (lane % V::LPR == 0) uU[lane / V::LPR] = ku + s2[Ge::o2k + lane / V::LPR];
    }
    __syncwarp();
    T uxr[N], uUr[M];  // ux, uU in registers: operands of the products and of the gradient outer products below
    mid_vload<T, N>(ux, uxr);
    mid_vload<T, M>(uU, uUr);
    if (lane < N) {  // ux' = A ux + B uU + d'
      T acc;
      if constexpr (V::RAB) {
        acc = mid_dot_r<T, N>(Ar, uxr, dp[lane]) + mid_dot_r<T, M>(Br, uUr, T(0));
      } else {
        T ar[N], br[M];
        mid_vload<T, N>(sA + lane * V::LDA, ar);  // the lane's own rows: vector loads
        mid_vload<T, M>(sB + lane * V::LDB, br);
        acc = mid_dot_r<T, N>(ar, uxr, dp[lane]) + mid_dot_r<T, M>(br, uUr, T(0));
      }
      uxn[lane] = acc;
    }
    __syncwarp();
    if (lane < N) {
      T xr[N];
      mid_vload<T, N>(uxn, xr);
      unu[lane] = s2[Ge::o2v + lane] + mid_symdot_r<T, N>(slot + Ge::oVp, xr, lane);
    }
    __syncwarp();
    // gradients of step t: element e = lane + LP k of each leaf
    if constexpr (V::RAB && LP == N && LP % M == 0) {
      // e = lane + N k  ->  (i, j) = (k, lane) for the N x N leaves; the N x M and M x M leaves have RPK = LP / M rows per k:
      // (i, j) = (RPK k + lane / M, lane % M).  Row operands come from registers (vector loads), column operands are one scalar
      // load each: ~24 shared-memory loads per step instead of ~180 dynamically indexed ones.
      constexpr int RPK = LP / M;
      T nur[N], unur[N], sXr[N], Uvr[M];
      mid_vload<T, N>(nu, nur);
      mid_vload<T, N>(unu, unur);
      mid_vload<T, N>(sX, sXr);
      mid_vload<T, M>(Uv, Uvr);
      const int jm = lane % M, hi = lane / M;
      const T uxl = ux[lane], sXl = sX[lane], uUl = uU[jm], Uvl = Uv[jm];
#pragma unroll
      for (int k = 0; k < EA; k++) {
        accA[k] += nur[k] * uxl + unur[k] * sXl;
        accQ[k] += T(0.5) * (uxr[k] * sXl + sXr[k] * uxl);
      }
#pragma unroll
      for (int k = 0; k < EB; k++) {
        T nui = nur[k * RPK], unui = unur[k * RPK];
#pragma unroll
        for (int q = 1; q < RPK; q++) { nui = hi == q ? nur[k * RPK + q] : nui; unui = hi == q ? unur[k * RPK + q] : unui; }
        accB[k] += nui * uUl + unui * Uvl;
      }
#pragma unroll
      for (int k = 0; k < ER; k++) {
        if (k * RPK < M) {
          T uUi = uUr[k * RPK < M ? k * RPK : 0], Uvi = Uvr[k * RPK < M ? k * RPK : 0];
#pragma unroll
          for (int q = 1; q < RPK; q++) {
            const int r = k * RPK + q < M ? k * RPK + q : 0;
            uUi = hi == q ? uUr[r] : uUi;
            Uvi = hi == q ? Uvr[r] : Uvi;
          }
          if (lane + LP * k < M * M) accR[k] += T(0.5) * (uUi * Uvl + Uvi * uUl);
        }
      }
    } else if constexpr (TI) {
#pragma unroll
      for (int k = 0; k < EA; k++) {
        const int e = lane + LP * k, i = e / N, j = e % N;
        accA[k] += nu[i] * ux[j] + unu[i] * sX[j];
        accQ[k] += T(0.5) * (ux[i] * sX[j] + sX[i] * ux[j]);
      }
#pragma unroll
      for (int k = 0; k < EB; k++) {
        const int e = lane + LP * k, i = e / M, j = e % M;
        accB[k] += nu[i] * uU[j] + unu[i] * Uv[j];
      }
#pragma unroll
      for (int k = 0; k < ER; k++) {
        const int e = lane + LP * k, i = e / M, j = e % M;
        if (e < M * M) accR[k] += T(0.5) * (uU[i] * Uv[j] + Uv[i] * uU[j]);
      }
    } else {
      // vectors-only mode (batch-summed gradients, grad_reduce.cuh): gq = sux, gr = uU, gd = uNu per problem, no outer product
      if (!reduce) {
#pragma unroll 4
        for (int e = lane; e < N * N; e += LP) {
          const int i = e / N, j = e % N;
          a.gA[idx * N * N + e] = nu[i] * ux[j] + unu[i] * sX[j];
          a.gQ[idx * N * N + e] = T(0.5) * (ux[i] * sX[j] + sX[i] * ux[j]);
        }
#pragma unroll 4
        for (int e = lane; e < N * M; e += LP) {
          const int i = e / M, j = e % M;
          a.gB[idx * N * M + e] = nu[i] * uU[j] + unu[i] * Uv[j];
          a.gM[idx * N * M + e] = ux[i] * Uv[j] + sX[i] * uU[j];
        }
        for (int e = lane; e < M * M; e += LP) {
          const int i = e / M, j = e % M;
          a.gR[idx * M * M + e] = T(0.5) * (uU[i] * Uv[j] + Uv[i] * uU[j]);
        }
      }
      if (lane < N) { a.gd[idx * N + lane] = unu[lane]; a.gq[idx * N + lane] = ux[lane]; }
      if (lane < M) a.gr[idx * M + lane] = uU[lane];
    }
    if (t == 0 && lane < N) {  // gx0 = M[0] uU[0] + A[0]^T uNu[0]
      T acc = T(0);
      if (a.Mx != nullptr && !ti)
        for (int b = 0; b < M; b++) acc += a.Mx[(prob * Th) * N * M + lane * M + b] * uU[b];
      if constexpr (V::RAB) {
        for (int k = 0; k < N; k++) acc += a.A[prob * N * N + k * N + lane] * unu[k];
      } else {
#pragma unroll 8
        for (int k = 0; k < N; k++) acc += sA[k * V::LDA + lane] * unu[k];
      }
      a.gx0[prob * N + lane] = acc;
    }
    if (t == Th - 1) {
      const T* XT = a.X + idx * N;
      if (!reduce || ti) {
        T* gQf = a.gQf + prob * N * N;
        for (int e = lane; e < N * N; e += LP) {
          const int i = e / N, j = e % N;
          gQf[e] = T(0.5) * (uxn[i] * XT[j] + uxn[j] * XT[i]);
        }
      }
      if (a.gqf != nullptr && lane < N) a.gqf[prob * N + lane] = uxn[lane];
    }
    __syncwarp();
    if (lane < N) ux[lane] = uxn[lane];
    cur = cur + 1 == NST ? 0 : cur + 1;
  }
  if constexpr (TI) {  // time-summed gradients of the time-invariant leaves (idoc/tilqr.py:38-54 in adjoint form)
#pragma unroll
    for (int k = 0; k < EA; k++) { a.gA[prob * N * N + lane + LP * k] = accA[k]; a.gQ[prob * N * N + lane + LP * k] = accQ[k]; }
#pragma unroll
    for (int k = 0; k < EB; k++) a.gB[prob * N * M + lane + LP * k] = accB[k];
#pragma unroll
    for (int k = 0; k < ER; k++)
      if (lane + LP * k < M * M) a.gR[prob * M * M + lane + LP * k] = accR[k];
  }
}

}  // namespace idoc
