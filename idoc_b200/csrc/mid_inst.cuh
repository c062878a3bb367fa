// Instantiation of the mid-size kernels for one (n, m): included by the generated translation units mid_<n>_<m>.cu
// (IDOC_N, IDOC_M, IDOC_LP defined by the build), which export mid_sweep_<n>_<m>_{f32,f64}(which, stream, args).
#pragma once
#include "mid.cuh"
#include "mid_mma.cuh"
#include "runtime.cuh"

namespace idoc {

#ifndef IDOC_MID_MMA_NW_DEFAULT
#define IDOC_MID_MMA_NW_DEFAULT 1
#endif
#ifndef IDOC_MID_DMMA_DEFAULT
#define IDOC_MID_DMMA_DEFAULT(N, M) (((N) == 16 && (M) == 16) ? 0 : 1)
#endif

// vector sweeps (which = 1 rollout, 2 adjoint backward, 3 adjoint forward + gradients): instantiation by (time-invariant,
// prefetch stages)
template <typename T, int N, int M, int LP, bool TI, int NST>
static int mid_launch_vec_t(int which, cudaStream_t s, const void* args) {
  using V = MidVec<T, N, M, LP, TI, NST>;
  const size_t smem = (size_t)V::G * V::GSTRIDE * sizeof(T);
  if (smem > (size_t)MAX_SMEM) return -2;  // IDOC_ERR_UNSUPPORTED (cannot happen for the compiled shapes)
  if (which == 1) {
    const GenFwdArgs<T>& g = *static_cast<const GenFwdArgs<T>*>(args);
    auto kern = mid_rollout_kernel<T, N, M, LP, TI, NST>;
    IDOC_CK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    ProfScope ps(K_ROLLOUT, s);
    kern<<<(unsigned)((g.a.nb + V::G - 1) / V::G), 32, smem, s>>>(g);
  } else if (which == 2) {
    const GenVjpArgs<T>& g = *static_cast<const GenVjpArgs<T>*>(args);
    auto kern = mid_adj_bwd_kernel<T, N, M, LP, TI, NST>;
    IDOC_CK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    ProfScope ps(K_ADJ_BWD, s);
    kern<<<(unsigned)((g.a.nb + V::G - 1) / V::G), 32, smem, s>>>(g);
  } else {
    const GenVjpArgs<T>& g = *static_cast<const GenVjpArgs<T>*>(args);
    auto kern = mid_adj_fwd_kernel<T, N, M, LP, TI, NST>;
    IDOC_CK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    ProfScope ps(K_ADJ_FWD, s);
    kern<<<(unsigned)((g.a.nb + V::G - 1) / V::G), 32, smem, s>>>(g);
  }
  IDOC_CK(cudaGetLastError());
  return 0;
}

template <typename T, int N, int M, int LP>
static int mid_launch_vec(int which, cudaStream_t s, const void* args) {
  const int ti = which == 1 ? static_cast<const GenFwdArgs<T>*>(args)->ti : static_cast<const GenVjpArgs<T>*>(args)->ti;
  // two prefetch stages: three and four measured slower at every shape (profiles/r2_mid_nst_ab.jsonl)
  if (ti) return mid_launch_vec_t<T, N, M, LP, true, 2>(which, s, args);
  return mid_launch_vec_t<T, N, M, LP, false, 2>(which, s, args);
}

// which: 0 = Riccati sweep, 1 = rollout, 2 = adjoint backward, 3 = adjoint forward + gradients
template <typename T, int N, int M, int LP>
static int mid_launch(int which, cudaStream_t s, const void* args) {
  if (which == 0) {
    const GenFwdArgs<T>& g = *static_cast<const GenFwdArgs<T>*>(args);
    auto go = [&](auto kern, size_t smem, int G) -> int {
      IDOC_CK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
      {
        ProfScope ps(K_RICCATI, s);
        kern<<<(unsigned)((g.a.nb + G - 1) / G), 32, smem, s>>>(g);
      }
      IDOC_CK(cudaGetLastError());
      return 0;
    };
    // n a multiple of 16: the tensor-core sweep (csrc/mid_mma.cuh: 3xTF32 mma.sync products in fp32, DMMA m8n8k4 in fp64).
    // IDOC_MID_MMA = 0 / 1 forces the register-tile / tensor-core sweep (the tests compare the two); default: the tensor-core
    // sweep where it measured faster (profiles/r2_mma_ab_f32.jsonl, r2_mma_ab_f64.jsonl).
    if constexpr (N % 16 == 0 && M % 8 == 0 && N <= 32 && M <= 16) {
      const int dflt = sizeof(T) == 4 ? ((N == 16 && M == 16) ? 0 : 1) : IDOC_MID_DMMA_DEFAULT(N, M);
      const int want = tune_int("IDOC_MID_MMA", dflt);
      if (want) {
        // problems per warp: 2 at n = 16 (the vector phases keep every lane busy), 1 at n = 32; IDOC_MID_MMA_G overrides
        constexpr int GD = N == 16 ? 2 : 1;
        const int G = N == 16 ? tune_int("IDOC_MID_MMA_G", GD) : 1;
        auto gom = [&](auto kern, size_t smem, int Gk, int nw = 1) -> int {
          IDOC_CK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
          {
            ProfScope ps(K_RICCATI, s);
            kern<<<(unsigned)((g.a.nb + Gk - 1) / Gk), 32 * nw, smem, s>>>(g);
          }
          IDOC_CK(cudaGetLastError());
          return 0;
        };
        if constexpr (N == 16) {
          if (G == 2) {
            if (g.ti) return gom(mid_riccati_mma_kernel<T, N, M, true, 2>, (size_t)(2 * MmaCfgT<T, N, M, true>::ELEMS + 16) * sizeof(T), 2);
            return gom(mid_riccati_mma_kernel<T, N, M, false, 2>, (size_t)(2 * MmaCfgT<T, N, M, false>::ELEMS + 16) * sizeof(T), 2);
          }
        }
#ifdef IDOC_MMA_TWO_WARPS
        // Two warps per problem at n = 32 (mid_riccati_mma_kernel<..., NW = 2>: n-tiles split between the warps, CTA barriers):
        // correct (the tests ran it), measured 17 - 23 % SLOWER in fp32 and within +-5 % in fp64 (profiles/r2_mma_two_warps_ab.jsonl:
        // the factorisation chain stays on one warp while the other waits at the barriers), so it is not compiled by default.
        if constexpr (N == 32) {
          if (tune_int("IDOC_MID_MMA_NW", IDOC_MID_MMA_NW_DEFAULT) == 2) {
            if (g.ti) return gom(mid_riccati_mma_kernel<T, N, M, true, 1, 2>, (size_t)MmaCfgT<T, N, M, true>::ELEMS * sizeof(T), 1, 2);
            return gom(mid_riccati_mma_kernel<T, N, M, false, 1, 2>, (size_t)MmaCfgT<T, N, M, false>::ELEMS * sizeof(T), 1, 2);
          }
        }
#endif
        if (g.ti) return gom(mid_riccati_mma_kernel<T, N, M, true, 1>, (size_t)MmaCfgT<T, N, M, true>::ELEMS * sizeof(T), 1);
        return gom(mid_riccati_mma_kernel<T, N, M, false, 1>, (size_t)MmaCfgT<T, N, M, false>::ELEMS * sizeof(T), 1);
      }
    }
    if (g.ti) {  // time-invariant (tilqr): the instantiation without M, q, r, d
      using C = MidCfg<T, N, M, LP, true>;
      return go(mid_riccati_kernel<T, N, M, LP, true>, (size_t)C::G * C::GSTRIDE * sizeof(T), C::G);
    }
    using C = MidCfg<T, N, M, LP, false>;
    return go(mid_riccati_kernel<T, N, M, LP, false>, (size_t)C::G * C::GSTRIDE * sizeof(T), C::G);
  }
  return mid_launch_vec<T, N, M, LP>(which, s, args);
}

#define IDOC_CAT2(a, b) a##b
#define IDOC_CAT(a, b) IDOC_CAT2(a, b)
#define IDOC_MIDSYM(base) IDOC_CAT(IDOC_CAT(IDOC_CAT(IDOC_CAT(base, _), IDOC_N), _), IDOC_M)

int IDOC_MIDSYM(mid_sweep_f32)(int which, cudaStream_t s, const void* args) { return mid_launch<float, IDOC_N, IDOC_M, IDOC_LP>(which, s, args); }
int IDOC_MIDSYM(mid_sweep_f64)(int which, cudaStream_t s, const void* args) { return mid_launch<double, IDOC_N, IDOC_M, IDOC_LP>(which, s, args); }

}  // namespace idoc
