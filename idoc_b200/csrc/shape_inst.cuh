// Instantiation of the LQR kernels for one (n, m) shape: included by the generated per-shape translation
// units (shape_<n>_<m>.cu), which define IDOC_N, IDOC_M, IDOC_L32, IDOC_L64 and optionally IDOC_TUNABLE.
#pragma once
#include "lqr_fwd.cuh"
#include "lqr_vjp.cuh"
#include "runtime.cuh"

namespace idoc {

template <typename T, int N, int M, int L, int NBUF_R, int NBUF>
static int run_fwd_v(cudaStream_t s, const LqrFwdArgs<T>& a, bool do_rollout) {
  {
    using C = RiccatiCfg<T, N, M, L, NBUF_R>;
    static_assert(C::WARP_BYTES <= MAX_SMEM, "riccati slab does not fit shared memory");
    constexpr int W = pick_warps(C::WARP_BYTES);
    auto kern = lqr_riccati_kernel<T, N, M, L, NBUF_R, W>;
    const size_t smem = (size_t)W * C::WARP_BYTES;
    IDOC_CK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const long long per_block = (long long)W * C::G;
    const unsigned grid = (unsigned)((a.nb + per_block - 1) / per_block);
    {
      ProfScope ps(K_RICCATI, s);
      kern<<<grid, W * 32, smem, s>>>(a);
    }
    IDOC_CK(cudaGetLastError());
  }
  if (do_rollout) {
    constexpr int OC = 4;
    using C = RolloutCfg<T, N, M, L, NBUF, OC>;
    constexpr int W = pick_warps(C::WARP_BYTES);
    auto kern = lqr_rollout_kernel<T, N, M, L, NBUF, OC, W>;
    const size_t smem = (size_t)W * C::WARP_BYTES;
    IDOC_CK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const long long per_block = (long long)W * C::G;
    const unsigned grid = (unsigned)((a.nb + per_block - 1) / per_block);
    {
      ProfScope ps(K_ROLLOUT, s);
      kern<<<grid, W * 32, smem, s>>>(a);
    }
    IDOC_CK(cudaGetLastError());
  }
  return 0;
}

template <typename T, int N, int M, int L, int NBUF, bool HAS_NUB>
static int run_vjp_v(cudaStream_t s, const LqrVjpArgs<T>& a) {
  {
    using C = AdjBwdCfg<T, N, M, L, NBUF, HAS_NUB>;
    constexpr int W = pick_warps(C::WARP_BYTES);
    auto kern = lqr_adj_bwd_kernel<T, N, M, L, NBUF, HAS_NUB, W>;
    const size_t smem = (size_t)W * C::WARP_BYTES;
    IDOC_CK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const long long per_block = (long long)W * C::G;
    const unsigned grid = (unsigned)((a.nb + per_block - 1) / per_block);
    {
      ProfScope ps(K_ADJ_BWD, s);
      kern<<<grid, W * 32, smem, s>>>(a);
    }
    IDOC_CK(cudaGetLastError());
  }
  {
    using C = AdjFwdCfg<T, N, M, L, NBUF, HAS_NUB>;
    constexpr int W = pick_warps(C::WARP_BYTES);
    auto kern = lqr_adj_fwd_kernel<T, N, M, L, NBUF, HAS_NUB, W>;
    const size_t smem = (size_t)W * C::WARP_BYTES;
    IDOC_CK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const long long per_block = (long long)W * C::G;
    const unsigned grid = (unsigned)((a.nb + per_block - 1) / per_block);
    {
      ProfScope ps(K_ADJ_FWD, s);
      kern<<<grid, W * 32, smem, s>>>(a);
    }
    IDOC_CK(cudaGetLastError());
  }
  return 0;
}

template <typename T, int N, int M, int L, int NBUF>
static int run_vjp_n(cudaStream_t s, const LqrVjpArgs<T>& a) {
  if (a.Nub != nullptr) return run_vjp_v<T, N, M, L, NBUF, true>(s, a);
  return run_vjp_v<T, N, M, L, NBUF, false>(s, a);
}

}  // namespace idoc

#define IDOC_CAT2(a, b) a##b
#define IDOC_CAT(a, b) IDOC_CAT2(a, b)
#define IDOC_SYM(base) IDOC_CAT(IDOC_CAT(IDOC_CAT(IDOC_CAT(base, _), IDOC_N), _), IDOC_M)

namespace idoc {

// Buffering defaults (measured, profiles/r1_notes.md): the Riccati sweep re-issues its loads as soon as the
// in-slab has been consumed (phase 4), so one buffer suffices and leaves room for more resident warps; the three
// streaming kernels read their slab until the end of the step and are double-buffered.
template <typename T>
static int fwd_entry(cudaStream_t s, const LqrFwdArgs<T>& a, bool rollout) {
  constexpr int LD = sizeof(T) == 4 ? IDOC_L32 : IDOC_L64;
#ifdef IDOC_TUNABLE
  const int L = tune_int("IDOC_L", LD), NBR = tune_int("IDOC_NBUF_R", 1), NB = tune_int("IDOC_NBUF", 2);
#define IDOC_V(LL, RR, SS) \
  if (L == LL && NBR == RR && NB == SS) return run_fwd_v<T, IDOC_N, IDOC_M, LL, RR, SS>(s, a, rollout);
  IDOC_V(4, 1, 1) IDOC_V(4, 1, 2) IDOC_V(4, 2, 1) IDOC_V(4, 2, 2) IDOC_V(8, 1, 1) IDOC_V(8, 1, 2) IDOC_V(8, 2, 1) IDOC_V(8, 2, 2)
#undef IDOC_V
#endif
  return run_fwd_v<T, IDOC_N, IDOC_M, LD, 1, 2>(s, a, rollout);
}
template <typename T>
static int vjp_entry(cudaStream_t s, const LqrVjpArgs<T>& a) {
  constexpr int LD = sizeof(T) == 4 ? IDOC_L32 : IDOC_L64;
#ifdef IDOC_TUNABLE
  const int L = tune_int("IDOC_L", LD), NB = tune_int("IDOC_NBUF", 2);
  if (L == 4 && NB == 1) return run_vjp_n<T, IDOC_N, IDOC_M, 4, 1>(s, a);
  if (L == 4 && NB == 2) return run_vjp_n<T, IDOC_N, IDOC_M, 4, 2>(s, a);
  if (L == 8 && NB == 1) return run_vjp_n<T, IDOC_N, IDOC_M, 8, 1>(s, a);
  if (L == 8 && NB == 2) return run_vjp_n<T, IDOC_N, IDOC_M, 8, 2>(s, a);
#endif
  return run_vjp_n<T, IDOC_N, IDOC_M, LD, 2>(s, a);
}

int IDOC_SYM(lqr_fwd_f32)(cudaStream_t s, const LqrFwdArgs<float>& a, bool rollout) { return fwd_entry<float>(s, a, rollout); }
int IDOC_SYM(lqr_fwd_f64)(cudaStream_t s, const LqrFwdArgs<double>& a, bool rollout) { return fwd_entry<double>(s, a, rollout); }
int IDOC_SYM(lqr_vjp_f32)(cudaStream_t s, const LqrVjpArgs<float>& a) { return vjp_entry<float>(s, a); }
int IDOC_SYM(lqr_vjp_f64)(cudaStream_t s, const LqrVjpArgs<double>& a) { return vjp_entry<double>(s, a); }

}  // namespace idoc
