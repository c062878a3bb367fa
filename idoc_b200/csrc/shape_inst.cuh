// Instantiation of the LQR kernels for one (n, m) shape: included by the generated per-shape translation
// units (shape_<n>_<m>_{fwd,vjp}.cu), which define IDOC_N, IDOC_M, IDOC_L32, IDOC_L64, IDOC_PART_FWD or
// IDOC_PART_VJP, and optionally IDOC_TUNABLE (compiles the variant table that the IDOC_* environment
// overrides select from; used by the tuning sweeps on the GPU box).
#pragma once
#include "lqr_fwd.cuh"
#include "lqr_vjp.cuh"
#include "runtime.cuh"

namespace idoc {

template <typename Kern, typename Args>
static int launch_sweep(Kern kern, int id, int warps, int G, size_t smem, cudaStream_t s, const Args& a) {
#ifdef IDOC_TUNABLE
  if (tune_int("IDOC_PAD_KERNEL", -1) == id) smem += (size_t)tune_int("IDOC_SMEM_PAD", 0);  // occupancy sensitivity probe
#endif
  IDOC_CK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  const long long per_block = (long long)warps * G;
  const unsigned grid = (unsigned)((a.nb + per_block - 1) / per_block);
  {
    ProfScope ps(id, s);
    kern<<<grid, warps * 32, smem, s>>>(a);
  }
  IDOC_CK(cudaGetLastError());
  return 0;
}

template <typename T, int N, int M, int L, int NBUF, int BM>
static int launch_riccati(cudaStream_t s, const LqrFwdArgs<T>& a) {
  using C = RiccatiCfg<T, N, M, L, NBUF, BM>;
  static_assert(C::WARP_BYTES <= MAX_SMEM, "riccati slab does not fit shared memory");
  constexpr int W = pick_warps(C::WARP_BYTES);
  return launch_sweep(lqr_riccati_kernel<T, N, M, L, NBUF, W, BM>, K_RICCATI, W, C::G, (size_t)W * C::WARP_BYTES, s, a);
}
// BM of the three streaming kernels: bit 7 = the loads of TWO consecutive timesteps travel as one request per array (even
// horizons), bit 8 = FOUR (horizons that are multiples of 4; falls back to two, then one); bits 9-11 = PADG, extra
// 16-byte units between the slabs of consecutive groups.
template <typename T, int N, int M, int L, int NBUF, int BM, int TB>
static int launch_rollout_tb(cudaStream_t s, const LqrFwdArgs<T>& a) {
  constexpr int OC = 4, B7 = BM & 127, PADG = (BM >> 9) & 7;
  using C = RolloutCfg<T, N, M, L, NBUF, OC, B7, TB, PADG>;
  constexpr int W = pick_warps(C::WARP_BYTES);
  return launch_sweep(lqr_rollout_kernel<T, N, M, L, NBUF, OC, W, B7, TB, PADG>, K_ROLLOUT, W, C::G, (size_t)W * C::WARP_BYTES, s, a);
}
template <typename T, int N, int M, int L, int NBUF, int BM>
static int launch_rollout(cudaStream_t s, const LqrFwdArgs<T>& a) {
  if constexpr ((BM & 256) != 0) {
    if (a.T_ % 4 == 0) return launch_rollout_tb<T, N, M, L, NBUF, BM, 4>(s, a);
  }
  if constexpr ((BM & (128 | 256)) != 0) {
    if (a.T_ % 2 == 0) return launch_rollout_tb<T, N, M, L, NBUF, BM, 2>(s, a);
  }
  return launch_rollout_tb<T, N, M, L, NBUF, BM, 1>(s, a);
}
template <typename T, int N, int M, int L, int NBUF, bool HAS_NUB, int BM, int TB>
static int launch_adj_bwd_tb(cudaStream_t s, const LqrVjpArgs<T>& a) {
  constexpr int B7 = BM & 127, PADG = (BM >> 9) & 7;
  using C = AdjBwdCfg<T, N, M, L, NBUF, HAS_NUB, B7, TB, PADG>;
  constexpr int W = pick_warps(C::WARP_BYTES);
  return launch_sweep(lqr_adj_bwd_kernel<T, N, M, L, NBUF, HAS_NUB, W, B7, TB, PADG>, K_ADJ_BWD, W, C::G, (size_t)W * C::WARP_BYTES, s, a);
}
template <typename T, int N, int M, int L, int NBUF, bool HAS_NUB, int BM>
static int launch_adj_bwd(cudaStream_t s, const LqrVjpArgs<T>& a) {
  if constexpr ((BM & 128) != 0) {
    if (a.T_ % 2 == 0) return launch_adj_bwd_tb<T, N, M, L, NBUF, HAS_NUB, BM, 2>(s, a);
  }
  return launch_adj_bwd_tb<T, N, M, L, NBUF, HAS_NUB, BM, 1>(s, a);
}
template <typename T, int N, int M, int L, int NBUF, bool HAS_NUB, int BM, int TB>
static int launch_adj_fwd_tb(cudaStream_t s, const LqrVjpArgs<T>& a) {
  constexpr int B6 = BM & 63, PADG = (BM >> 9) & 7;
  using C = AdjFwdDCfg<T, N, M, L, NBUF, HAS_NUB, B6, TB, PADG>;
  constexpr int W = pick_warps(C::WARP_BYTES);
  return launch_sweep(lqr_adj_fwd_direct_kernel<T, N, M, L, NBUF, HAS_NUB, W, B6, TB, PADG>, K_ADJ_FWD, W, C::G,
                      (size_t)W * C::WARP_BYTES, s, a);
}
template <typename T, int N, int M, int L, int NBUF, bool HAS_NUB, int BM>
static int launch_adj_fwd(cudaStream_t s, const LqrVjpArgs<T>& a) {
  // BM bit 6: direct (unstaged) gradient stores, per-problem gradients on chunk-aligned shapes only
  if constexpr ((BM & 64) != 0 && adj_fwd_direct_ok<T, N, M>()) {
    if constexpr ((BM & 256) != 0) {
      if (a.T_ % 4 == 0) return launch_adj_fwd_tb<T, N, M, L, NBUF, HAS_NUB, BM, 4>(s, a);
    }
    if constexpr ((BM & (128 | 256)) != 0) {
      if (a.T_ % 2 == 0) return launch_adj_fwd_tb<T, N, M, L, NBUF, HAS_NUB, BM, 2>(s, a);
    }
    return launch_adj_fwd_tb<T, N, M, L, NBUF, HAS_NUB, BM, 1>(s, a);
  }
  // staged kernel (shapes whose rows are not chunk-aligned): bits 3-5 mean bulk STORES there
  constexpr int BMS = (BM & 64) ? (BM & 7) : (BM & 63);
  using C = AdjFwdCfg<T, N, M, L, NBUF, HAS_NUB, BMS>;
  constexpr int W = pick_warps(C::WARP_BYTES);
  return launch_sweep(lqr_adj_fwd_kernel<T, N, M, L, NBUF, HAS_NUB, W, BMS>, K_ADJ_FWD, W, C::G, (size_t)W * C::WARP_BYTES, s, a);
}

}  // namespace idoc

#define IDOC_CAT2(a, b) a##b
#define IDOC_CAT(a, b) IDOC_CAT2(a, b)
#define IDOC_SYM(base) IDOC_CAT(IDOC_CAT(IDOC_CAT(IDOC_CAT(base, _), IDOC_N), _), IDOC_M)

// Defaults (measured on B200, profiles/r1_notes.md): buffering and bulk mode per kernel and dtype.  The Riccati sweep
// re-issues its loads as soon as the in-slab has been consumed (phase 4), so one buffer suffices and leaves room for more
// resident warps; the three streaming kernels read their slab until the end of the step and are double-buffered.
// Bulk (TMA) copies are used for the >= 256-byte segments only, i.e. at n = 8 in these shapes.
namespace idoc {
template <typename T, int N, int M>
struct Defaults {
  static constexpr bool BIG = N * N * (int)sizeof(T) >= 256 && (N * N * (int)sizeof(T)) % 16 == 0 && (N * M * (int)sizeof(T)) % 16 == 0;
  // f64: A,B,M bulk early + Q bulk late under the aliased exchange area + single-buffered slot + q, r, R fetched
  // straight into registers (12 resident warps); f32: A,Q bulk + bulk slot store
  static constexpr int RIC_NB = 1, RIC_BM = BIG ? (sizeof(T) == 8 ? 124 : 10 + 2 * 256) : 0;  // f32: + two 16-byte units between group slabs
  // BM bit 7 (the three streaming kernels): the loads of two consecutive timesteps travel as one request per array (even
  // horizons; odd ones take the unpaired kernel) -- twice the bytes per request, half the requests
  static constexpr int ROLL_NB = 2, ROLL_BM = BIG ? (sizeof(T) == 8 ? 2 : 131) : 0;   // A, ws[K..V] (f32: and B) bulk; f32 paired
  static constexpr int ABWD_NB = 2, ABWD_BM = BIG ? 130 : 0;                          // A and ws[Linv..K] bulk, paired
  // direct (unstaged) stores of the large gradient leaves, small leaves flushed as 4-step blocks, paired loads;
  // f64: A, V, K, B bulk; f32: A bulk and FOUR timesteps per request when 4 divides the horizon (bit 8)
  static constexpr int AFWD_NB = 2, AFWD_BM = BIG ? (sizeof(T) == 8 ? 228 : 353) : 0;
  // vectors-only mode (batch-summed gradients): the sweep writes 20 instead of 228 scalars per problem-step, so it behaves
  // like the rollout (a read stream): smaller per-warp footprints win (tools/tune_vo.py, profiles/r2_tune_vo_*.jsonl:
  // f64 2.39 -> 2.04 ms with A alone on the bulk path and single-step requests, f32 1.28 -> 1.00 ms with two steps per request)
  static constexpr int AFWD_VO_BM = BIG ? (sizeof(T) == 8 ? 97 : 228) : 0;
};
}  // namespace idoc

namespace idoc {

#ifdef IDOC_PART_FWD
template <typename T>
static int fwd_entry(cudaStream_t s, const LqrFwdArgs<T>& a, bool rollout) {
  constexpr int L = sizeof(T) == 4 ? IDOC_L32 : IDOC_L64;
  using D = Defaults<T, IDOC_N, IDOC_M>;
  int rc = -1;
#ifdef IDOC_TUNABLE
  {
    const int nb = tune_int("IDOC_NBUF_R", -1), bm = tune_int("IDOC_BM_RIC", -1);
#define IDOC_V(NB, BMV) \
  if (rc < 0 && nb == NB && bm == BMV) rc = launch_riccati<T, IDOC_N, IDOC_M, L, NB, BMV>(s, a);
    IDOC_V(1, 0) IDOC_V(1, 10) IDOC_V(1, 26) IDOC_V(1, 60) IDOC_V(1, 124) IDOC_V(1, 122) IDOC_V(1, 90)
    IDOC_V(1, 380) IDOC_V(1, 636) IDOC_V(1, 892) IDOC_V(1, 1148) IDOC_V(1, 1404) IDOC_V(1, 1660) IDOC_V(1, 1916)
    IDOC_V(1, 266) IDOC_V(1, 522) IDOC_V(1, 778) IDOC_V(1, 1034) IDOC_V(1, 1290) IDOC_V(1, 1546) IDOC_V(1, 1802)
#undef IDOC_V
  }
#endif
  if (rc < 0) rc = launch_riccati<T, IDOC_N, IDOC_M, L, D::RIC_NB, D::RIC_BM>(s, a);
  if (rc != 0 || !rollout) return rc;
  rc = -1;
#ifdef IDOC_TUNABLE
  {
    const int nb = tune_int("IDOC_NBUF_O", -1), bm = tune_int("IDOC_BM_ROLL", -1);
#define IDOC_V(NB, BMV) \
  if (rc < 0 && nb == NB && bm == BMV) rc = launch_rollout<T, IDOC_N, IDOC_M, L, NB, BMV>(s, a);
    IDOC_V(1, 0) IDOC_V(2, 0) IDOC_V(2, 1) IDOC_V(2, 2) IDOC_V(2, 64) IDOC_V(2, 65) IDOC_V(2, 66) IDOC_V(1, 64) IDOC_V(1, 65)
    IDOC_V(2, 128) IDOC_V(2, 129) IDOC_V(2, 130) IDOC_V(2, 131) IDOC_V(2, 192) IDOC_V(2, 194) IDOC_V(2, 195)
    IDOC_V(2, 258) IDOC_V(2, 259)
    IDOC_V(2, 514) IDOC_V(2, 1026) IDOC_V(2, 1538) IDOC_V(2, 2050) IDOC_V(2, 2562) IDOC_V(2, 3074) IDOC_V(2, 3586)
    IDOC_V(2, 643) IDOC_V(2, 1155) IDOC_V(2, 1667) IDOC_V(2, 2179) IDOC_V(2, 2691) IDOC_V(2, 3203) IDOC_V(2, 3715)
#undef IDOC_V
  }
#endif
  if (rc < 0) rc = launch_rollout<T, IDOC_N, IDOC_M, L, D::ROLL_NB, D::ROLL_BM>(s, a);
  return rc;
}
int IDOC_SYM(lqr_fwd_f32)(cudaStream_t s, const LqrFwdArgs<float>& a, bool rollout) { return fwd_entry<float>(s, a, rollout); }
int IDOC_SYM(lqr_fwd_f64)(cudaStream_t s, const LqrFwdArgs<double>& a, bool rollout) { return fwd_entry<double>(s, a, rollout); }
#endif  // IDOC_PART_FWD

#ifdef IDOC_PART_VJP
template <typename T, bool HAS_NUB>
static int vjp_entry_n(cudaStream_t s, const LqrVjpArgs<T>& a) {
  constexpr int L = sizeof(T) == 4 ? IDOC_L32 : IDOC_L64;
  using D = Defaults<T, IDOC_N, IDOC_M>;
  int rc = -1;
#ifdef IDOC_TUNABLE
  if (!HAS_NUB) {
    const int nb = tune_int("IDOC_NBUF_B", -1), bm = tune_int("IDOC_BM_ABWD", -1);
#define IDOC_V(NB, BMV) \
  if (rc < 0 && nb == NB && bm == BMV) rc = launch_adj_bwd<T, IDOC_N, IDOC_M, L, NB, false, BMV>(s, a);
    IDOC_V(1, 0) IDOC_V(1, 2) IDOC_V(1, 3) IDOC_V(2, 0) IDOC_V(2, 1) IDOC_V(2, 2) IDOC_V(2, 3)
    IDOC_V(2, 128) IDOC_V(2, 129) IDOC_V(2, 130) IDOC_V(2, 131)
    IDOC_V(2, 642) IDOC_V(2, 1154) IDOC_V(2, 1666) IDOC_V(2, 2178) IDOC_V(2, 2690) IDOC_V(2, 3202) IDOC_V(2, 3714)
#undef IDOC_V
  }
#endif
  if (rc < 0) rc = launch_adj_bwd<T, IDOC_N, IDOC_M, L, D::ABWD_NB, HAS_NUB, D::ABWD_BM>(s, a);
  if (rc != 0) return rc;
  rc = -1;
#ifdef IDOC_TUNABLE
  if (!HAS_NUB) {
    const int nb = tune_int("IDOC_NBUF_F", -1), bm = tune_int("IDOC_BM_AFWD", -1);
#define IDOC_V(NB, BMV) \
  if (rc < 0 && nb == NB && bm == BMV) rc = launch_adj_fwd<T, IDOC_N, IDOC_M, L, NB, false, BMV>(s, a);
    IDOC_V(2, 0) IDOC_V(2, 2) IDOC_V(2, 64) IDOC_V(2, 66) IDOC_V(2, 96) IDOC_V(2, 98) IDOC_V(2, 82) IDOC_V(1, 98)
    IDOC_V(2, 224) IDOC_V(2, 225) IDOC_V(2, 226) IDOC_V(2, 228) IDOC_V(1, 224) IDOC_V(1, 225)
    IDOC_V(2, 100) IDOC_V(2, 99) IDOC_V(2, 97) IDOC_V(1, 100) IDOC_V(1, 228)
    IDOC_V(2, 352) IDOC_V(2, 353) IDOC_V(2, 356)
    IDOC_V(2, 740) IDOC_V(2, 1252) IDOC_V(2, 1764) IDOC_V(2, 2276) IDOC_V(2, 2788) IDOC_V(2, 3300) IDOC_V(2, 3812)
    IDOC_V(2, 865) IDOC_V(2, 1377) IDOC_V(2, 1889) IDOC_V(2, 2401) IDOC_V(2, 2913) IDOC_V(2, 3425) IDOC_V(2, 3937)
#undef IDOC_V
  }
#endif
  if (rc < 0 && a.reduce_batch && D::AFWD_VO_BM != D::AFWD_BM) rc = launch_adj_fwd<T, IDOC_N, IDOC_M, L, D::AFWD_NB, HAS_NUB, D::AFWD_VO_BM>(s, a);
  if (rc < 0) rc = launch_adj_fwd<T, IDOC_N, IDOC_M, L, D::AFWD_NB, HAS_NUB, D::AFWD_BM>(s, a);
  return rc;
}
template <typename T>
static int vjp_entry(cudaStream_t s, const LqrVjpArgs<T>& a) {
  if (a.Nub != nullptr) return vjp_entry_n<T, true>(s, a);
  return vjp_entry_n<T, false>(s, a);
}
int IDOC_SYM(lqr_vjp_f32)(cudaStream_t s, const LqrVjpArgs<float>& a) { return vjp_entry<float>(s, a); }
int IDOC_SYM(lqr_vjp_f64)(cudaStream_t s, const LqrVjpArgs<double>& a) { return vjp_entry<double>(s, a); }
#endif  // IDOC_PART_VJP

}  // namespace idoc
