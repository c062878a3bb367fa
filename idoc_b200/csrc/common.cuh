// Common device helpers for the idoc_b200 kernels (sm_100a).
//
// Execution model shared by every time-sweep kernel in this library:
//   * a *group* of L lanes (L in {1,2,4,8,16,32}) owns one LQR problem; a warp owns G = 32/L problems;
//   * every warp is self-contained: it streams the per-timestep slabs of its own G problems from HBM
//     into its private shared-memory region with 16-byte cp.async (coalesced: consecutive lanes copy
//     consecutive 16-byte chunks of one problem's slab), waits, computes, and writes results back
//     through shared memory with 16-byte coalesced stores.  No CTA-wide barriers on the hot path.
//   * lanes of a group exchange data only through shared memory + __syncwarp().
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace idoc {

#define IDOC_DEVICE __device__ __forceinline__
#ifndef IDOC_CPASYNC_CA
#define IDOC_CPASYNC_CA 0
#endif

__host__ __device__ constexpr int cdiv(int a, int b) { return (a + b - 1) / b; }
__host__ __device__ constexpr int rup(int a, int b) { return cdiv(a, b) * b; }
__host__ __device__ constexpr int tri(int n) { return n * (n + 1) / 2; }
__host__ __device__ constexpr int cgcd(int a, int b) { return b == 0 ? a : cgcd(b, a % b); }
__host__ __device__ constexpr int cmin(int a, int b) { return a < b ? a : b; }
__host__ __device__ constexpr int cmax(int a, int b) { return a > b ? a : b; }

// index into row-major packed upper triangle of a symmetric n x n matrix
__host__ __device__ constexpr int triu(int i, int j, int n) {
  return i <= j ? (i * n - i * (i - 1) / 2 + (j - i)) : (j * n - j * (j - 1) / 2 + (i - j));
}

// round a byte count up to an odd multiple of 16 B: group slabs laid out at this stride never
// collide on shared-memory banks when different groups read the same slab offset with LDS.128.
__host__ __device__ constexpr int odd16(int bytes) {
  int q = cdiv(bytes, 16);
  return ((q & 1) ? q : q + 1) * 16;
}

// ------------------------------------------------------------------ async copies
IDOC_DEVICE uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

template <int BYTES>
IDOC_DEVICE void cp_async(uint32_t smem_dst, const void* gsrc) {
  static_assert(BYTES == 4 || BYTES == 8 || BYTES == 16, "cp.async size");
  if constexpr (BYTES == 16 && !IDOC_CPASYNC_CA) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(smem_dst), "l"(gsrc) : "memory");
  } else {
    asm volatile("cp.async.ca.shared.global [%0], [%1], %2;\n" ::"r"(smem_dst), "l"(gsrc), "n"(BYTES) : "memory");
  }
}
IDOC_DEVICE void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::: "memory"); }
template <int N_>
IDOC_DEVICE void cp_async_wait() {
  asm volatile("cp.async.wait_group %0;\n" ::"n"(N_) : "memory");
}

template <int BYTES>
struct RawVec;
template <>
struct RawVec<16> {
  using type = uint4;
};
template <>
struct RawVec<8> {
  using type = uint2;
};
template <>
struct RawVec<4> {
  using type = uint32_t;
};

// streaming (evict-first) global store of BYTES
template <int BYTES>
IDOC_DEVICE void st_global_cs(void* g, const void* s) {
  using V = typename RawVec<BYTES>::type;
  __stcs(reinterpret_cast<V*>(g), *reinterpret_cast<const V*>(s));
}

// streaming store of CNT register values (CNT * sizeof(T) in {4, 8, 16}; g aligned to that size)
template <typename T, int CNT>
IDOC_DEVICE void stcs_regs(T* g, const T* v) {
  constexpr int BYTES = CNT * (int)sizeof(T);
  static_assert(BYTES == 4 || BYTES == 8 || BYTES == 16, "vector store size");
  if constexpr (sizeof(T) == 8) {
    if constexpr (CNT == 2) __stcs(reinterpret_cast<double2*>(g), make_double2((double)v[0], (double)v[1]));
    else __stcs(reinterpret_cast<double*>(g), (double)v[0]);
  } else {
    if constexpr (CNT == 4) __stcs(reinterpret_cast<float4*>(g), make_float4((float)v[0], (float)v[1], (float)v[2], (float)v[3]));
    else if constexpr (CNT == 2) __stcs(reinterpret_cast<float2*>(g), make_float2((float)v[0], (float)v[1]));
    else __stcs(reinterpret_cast<float*>(g), (float)v[0]);
  }
}

// ------------------------------------------------------------------ shared-memory vector access
// Load CNT consecutive elements of T from shared memory; ALIGN = known byte alignment of p.
template <typename T, int CNT, int ALIGN>
IDOC_DEVICE void lds(const T* p, T* dst) {
  constexpr int E16 = 16 / (int)sizeof(T), E8 = 8 / (int)sizeof(T);
  constexpr int n16 = (ALIGN >= 16) ? CNT / E16 : 0;
  constexpr int o8 = n16 * E16;
  constexpr int n8 = (ALIGN >= 8 && E8 >= 1) ? (CNT - o8) / E8 : 0;
  constexpr int o1 = o8 + n8 * E8;
#pragma unroll
  for (int i = 0; i < n16; i++) {
    const uint4 v = *reinterpret_cast<const uint4*>(p + i * E16);
    T tmp[E16];
    *reinterpret_cast<uint4*>(tmp) = v;
#pragma unroll
    for (int e = 0; e < E16; e++) dst[i * E16 + e] = tmp[e];
  }
#pragma unroll
  for (int i = 0; i < n8; i++) {
    const uint2 v = *reinterpret_cast<const uint2*>(p + o8 + i * E8);
    T tmp[E8 > 0 ? E8 : 1];
    *reinterpret_cast<uint2*>(tmp) = v;
#pragma unroll
    for (int e = 0; e < E8; e++) dst[o8 + i * E8 + e] = tmp[e];
  }
#pragma unroll
  for (int i = o1; i < CNT; i++) dst[i] = p[i];
}

template <typename T, int CNT, int ALIGN>
IDOC_DEVICE void sts(T* p, const T* src) {
  constexpr int E16 = 16 / (int)sizeof(T), E8 = 8 / (int)sizeof(T);
  constexpr int n16 = (ALIGN >= 16) ? CNT / E16 : 0;
  constexpr int o8 = n16 * E16;
  constexpr int n8 = (ALIGN >= 8 && E8 >= 1) ? (CNT - o8) / E8 : 0;
  constexpr int o1 = o8 + n8 * E8;
#pragma unroll
  for (int i = 0; i < n16; i++) {
    T tmp[E16];
#pragma unroll
    for (int e = 0; e < E16; e++) tmp[e] = src[i * E16 + e];
    *reinterpret_cast<uint4*>(p + i * E16) = *reinterpret_cast<const uint4*>(tmp);
  }
#pragma unroll
  for (int i = 0; i < n8; i++) {
    T tmp[E8 > 0 ? E8 : 1];
#pragma unroll
    for (int e = 0; e < E8; e++) tmp[e] = src[o8 + i * E8 + e];
    *reinterpret_cast<uint2*>(p + o8 + i * E8) = *reinterpret_cast<const uint2*>(tmp);
  }
#pragma unroll
  for (int i = o1; i < CNT; i++) p[i] = src[i];
}

// byte alignment of element offset `off` (elements of size sz) from a 16B-aligned base
__host__ __device__ constexpr int align_of(int off, int sz) {
  int b = off * sz;
  return (b % 16 == 0) ? 16 : ((b % 8 == 0) ? 8 : ((b % 4 == 0) ? 4 : sz));
}
// alignment guaranteed for every row of a row-major matrix with `cols` columns
__host__ __device__ constexpr int row_align(int cols, int sz) { return align_of(cols, sz); }

// ------------------------------------------------------------------ math
template <typename T>
IDOC_DEVICE T rsqrt_(T x);
template <>
IDOC_DEVICE float rsqrt_<float>(float x) {
  // the bare MUFU.RSQ (<= 2 ulp for normal arguments; rsqrtf() wraps it in a denormal-rescaling sequence that triples the
  // instruction count) and one Newton step (~1 ulp).  Pivots are checked > 0 by the callers; a denormal pivot (flushed to
  // zero: +inf, then NaN after the Newton step) is a failed factorisation either way and is reported through the NONFINITE status.
  float y;
  asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y * (1.5f - 0.5f * x * y * y);
}
template <>
IDOC_DEVICE double rsqrt_<double>(double x) {
  return rsqrt(x);  // <= 1 ulp (CUDA math library), no division / sqrt sequence
}
template <typename T>
IDOC_DEVICE T abs_(T x) {
  return x < T(0) ? -x : x;
}

// ------------------------------------------------------------------ packed FP32 multiply-adds
// sm_100 issues two FP32 FMAs on a register pair as ONE instruction (FFMA2, `__ffma2_rn`): measured 109 multiply-adds
// per clock per SM at 1.7 warp-instructions per clock against 124 at 3.9 for FFMA (tools/mma_rate.cu), i.e. the same
// arithmetic for 44 % of the issue slots.  Same rounding as fmaf.  Used by the register-tiled products of the mid-size
// kernels (mid.cuh: -4 % / -8 % on the fp32 Riccati sweep at n = 16 / 32); in the n = 8 Riccati sweep the register-pair
// moves it needs cancel the saving (1.44 vs 1.42 ms), so that kernel keeps scalar FMAs.
#ifndef IDOC_FFMA2
#define IDOC_FFMA2 1
#endif
// acc[i] += x[i] * s  for i < CNT (register arrays)
template <typename T, int CNT>
IDOC_DEVICE void axpy_regs(T* acc, const T* x, T s) {
  if constexpr (IDOC_FFMA2 && sizeof(T) == 4 && CNT % 2 == 0) {
    const float2 s2 = make_float2(s, s);
#pragma unroll
    for (int i = 0; i < CNT; i += 2) {
      const float2 r = __ffma2_rn(make_float2(x[i], x[i + 1]), s2, make_float2(acc[i], acc[i + 1]));
      acc[i] = r.x;
      acc[i + 1] = r.y;
    }
  } else {
#pragma unroll
    for (int i = 0; i < CNT; i++) acc[i] += x[i] * s;
  }
}
// (acc0, acc1) += (x0, x1) * s
template <typename T>
IDOC_DEVICE void fma_pair(T& acc0, T& acc1, T x0, T x1, T s) {
  if constexpr (IDOC_FFMA2 && sizeof(T) == 4) {
    const float2 r = __ffma2_rn(make_float2(x0, x1), make_float2(s, s), make_float2(acc0, acc1));
    acc0 = r.x;
    acc1 = r.y;
  } else {
    acc0 += x0 * s;
    acc1 += x1 * s;
  }
}

// ------------------------------------------------------------------ per-lane copy plans
// A plan describes, for one lane, which 'GR'-byte chunks of a problem's slab it moves each step.
// Segments are appended in slab order; all lanes call add() with the same arguments.
template <int R, int GR>
struct CopyPlan {
  const char* gp[R];  // global address of this lane's chunk for (problem 0, t 0) of the segment
  int stride[R];      // bytes between consecutive (problem,t) slabs of that segment (0 = lane idle)
  int soff[R];        // byte offset of the chunk inside the group's shared-memory slab
  int nchunks;

  IDOC_DEVICE void reset() {
#pragma unroll
    for (int r = 0; r < R; r++) {
      gp[r] = nullptr;
      stride[r] = 0;
      soff[r] = 0;
    }
    nchunks = 0;
  }
  // bytes: size of one (problem,t) slab of this segment in global memory (also the stride);
  // sub_off/sub_bytes: the sub-range of it that is copied; smem_off: where it lands in the slab.
  IDOC_DEVICE void add(const void* base, int bytes, int sub_off, int sub_bytes, int smem_off, int lane) {
    const int n = sub_bytes / GR;
#pragma unroll
    for (int r = 0; r < R; r++) {
      const int c = lane + 32 * r - nchunks;
      if (c >= 0 && c < n) {
        gp[r] = static_cast<const char*>(base) + sub_off + c * GR;
        stride[r] = bytes;
        soff[r] = smem_off + c * GR;
      }
    }
    nchunks += n;
  }
  // global -> shared for one group (slab index idx = problem*T + t < 2^31; smem_slab = 32-bit shared address)
  IDOC_DEVICE void load(uint32_t smem_slab, unsigned idx) const {
#pragma unroll
    for (int r = 0; r < R; r++)
      if (stride[r]) cp_async<GR>(smem_slab + soff[r], gp[r] + (unsigned long long)idx * (unsigned)stride[r]);
  }
  // shared -> global for one group
  IDOC_DEVICE void store(const char* smem_slab, unsigned idx) const {
#pragma unroll
    for (int r = 0; r < R; r++)
      if (stride[r])
        st_global_cs<GR>(const_cast<char*>(gp[r]) + (unsigned long long)idx * (unsigned)stride[r], smem_slab + soff[r]);
  }
};

// ------------------------------------------------------------------ bulk async copies (TMA engine) + mbarriers
// Used for the large (>= 256 B) per-step segments only: the engine is request-rate bound on small ones
// (loader.cuh).  A bulk copy writes shared memory through the async proxy at full width instead of the
// 32-byte-sector LDGSTS path, which relieves the LSU data pipe.
IDOC_DEVICE void mbar_init(uint64_t* bar, int count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;\n" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
IDOC_DEVICE void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;\n" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
IDOC_DEVICE void mbar_wait(uint64_t* bar, uint32_t parity) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "WAIT_LOOP:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@p bra DONE;\n"
      "bra WAIT_LOOP;\n"
      "DONE:\n"
      "}\n" ::"r"(smem_u32(bar)),
      "r"(parity)
      : "memory");
}
// one deterministic leader lane of a fully converged warp (all 32 lanes must execute this)
IDOC_DEVICE bool elect_one() {
  uint32_t pred;
  asm volatile(
      "{\n"
      ".reg .pred P;\n"
      "elect.sync _|P, 0xffffffff;\n"
      "selp.u32 %0, 1, 0, P;\n"
      "}\n"
      : "=r"(pred));
  return pred != 0;
}
IDOC_DEVICE void fence_mbar_init() { asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory"); }
IDOC_DEVICE void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory"); }
IDOC_DEVICE void bulk_g2s(uint32_t smem_dst, const void* gsrc, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];\n" ::"r"(smem_dst),
               "l"(gsrc), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}
IDOC_DEVICE void bulk_s2g(void* gdst, uint32_t smem_src, uint32_t bytes) {
  asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;\n" ::"l"(gdst), "r"(smem_src), "r"(bytes)
               : "memory");
}
IDOC_DEVICE void bulk_commit() { asm volatile("cp.async.bulk.commit_group;\n" ::: "memory"); }
IDOC_DEVICE void bulk_wait_read0() { asm volatile("cp.async.bulk.wait_group.read 0;\n" ::: "memory"); }
IDOC_DEVICE void bulk_wait_all0() { asm volatile("cp.async.bulk.wait_group 0;\n" ::: "memory"); }

}  // namespace idoc
