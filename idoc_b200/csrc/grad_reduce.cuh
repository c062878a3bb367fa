// Batch-summed parameter gradients of the LQR implicit VJP (idoc_lqr_vjp with reduce_batch = 1): the path's one
// reduction over problems, and the buffer the multi-GPU all-reduce combines.
//
// Every gradient leaf of a timestep is a sum of outer products of six per-problem vectors (lqr_vjp.cuh header):
//     gA = Nu (x) sux + uNu (x) sX     gB = Nu (x) uU + uNu (x) U      gd = uNu
//     gQ = sym(sux (x) sX)             gq = sux        gR = sym(uU (x) U)      gr = uU
//     gM = sux (x) U + sX (x) uU       gQf = sym(uX[T-1] (x) X[T-1])   gqf = uX[T-1]
// so summed over the batch each leaf is a GEMM with the batch as its contraction axis:  [gA | gB](t) = Lt^T R with
// rows (Nu_b, uNu_b) on the left and ((sux_b | uU_b), (sX_b | U_b)) on the right, K = 2B.  The adjoint forward sweep
// therefore runs in "vectors only" mode (it stores sux, uU, uNu: 20 of the 228 scalars per problem-step at n=8, m=4,
// and no outer product), and the two kernels here form the sums:
//
//   grad_reduce_partial_kernel : grid (T+1, NC chunks of the batch, tile groups).  The CTA stages the operand vectors of 48
//                                problems at a time into shared memory (16-byte cp.async, double-buffered); a lane owns up
//                                to MAXT 4x4 register tiles of the step's outputs and accumulates them over its warp's share
//                                of the staged problems (4 x 16-byte shared loads per 32 multiply-adds).  The four warps of
//                                a CTA are summed in a fixed order through shared memory and the CTA writes one partial
//                                [t][PER_STEP] row of chunk c.
//   grad_reduce_final_kernel   : sums the NC partials of every output element in chunk order, symmetrises gQ, gR, gQf
//                                (kkt applies .symm(), idoc/lqr.py:130) and scatters into the ten leaf buffers.
//
// No atomics anywhere: the result is bit-reproducible run to run (round 1 summed with floating-point atomicAdd: one per
// element per warp-step, 187 M of them at config 2, slower than writing every per-problem gradient).
#pragma once
#include "common.cuh"

namespace idoc {

template <typename T>
struct GradReduceArgs {
  const T *X, *U, *Nu, *x0;     // forward solution [B,T,n], [B,T,m], [B,T,n]; x0 [B,n]
  const T *uq, *ur, *ud, *uqf;  // adjoint solution: sux [B,T,n], uU [B,T,m], uNu [B,T,n], uX[T-1] [B,n]
  T* part;                      // [NC][T * PER_STEP + n*n + n]
  T *gQ, *gq, *gQf, *gqf, *gM, *gR, *gr, *gA, *gB, *gd;  // outputs, shapes without the batch axis
  long long nb;
  int T_, n, m, NC;
  int stage;  // problems staged in shared memory at a time (gr_stage_problems)
};

constexpr int GR_TILE = 4, GR_MAXT = 2, GR_WARPS = 4;

__host__ __device__ inline int gr_per_step(int n, int m) { return 2 * n * (n + m) + m * m + 2 * n + m; }
__host__ __device__ inline long long gr_part_elems(int T_, int n, int m) { return (long long)T_ * gr_per_step(n, m) + n * n + n; }
// 4x4 tile tasks of one timestep: [gA|gB] and [gQraw|gM] (n x (n+m) each, column tiles never straddle the n|m boundary),
// gRraw (m x m), and the vector sums gd, gq (n) and gr (m) as one-column tiles
__host__ __device__ inline int gr_tasks_step(int n, int m) {
  const int rn = cdiv(n, GR_TILE), rm = cdiv(m, GR_TILE);
  return 2 * rn * (rn + rm) + rm * rm + 2 * rn + rm;
}
__host__ __device__ inline int gr_tasks_final(int n) {
  const int rn = cdiv(n, GR_TILE);
  return rn * rn + rn;
}

// ---- staging.  The operand vectors of one (problem, timestep) are six short segments of six different arrays
// (Nu, uNu, sux, sX: n scalars; uU, U: m scalars), each 32 - 128 bytes, a problem stride apart from its neighbour's.
// A CTA stages them for GR_STAGE problems at a time into shared memory with 16-byte cp.async copies, double-buffered, so that
// dozens of KB per SM are in flight while the tiles of the previous stage are accumulated (the first version read the
// operands straight from global memory inside the FMA loop: latency-bound at 16 resident warps, 5 % of the HBM peak).
// Row layout of a staged problem (elements; every segment padded to a multiple of 4):
//     [ Nu | uNu | sux | sX | uU | U ]        final pseudo-step:  [ uX[T-1] | X[T-1] ]
constexpr int GR_STAGE_MAX = 48;  // problems per stage (fewer when the staged rows of a large shape would not fit)

__host__ __device__ inline int gr_pad4(int w) { return (w + 3) / 4 * 4; }
// row stride of a staged problem: the six segments + one 4-element pad when that makes the stride an ODD number of 16-byte
// (fp32) / 32-byte (fp64) units, so that the problems a warp walks side by side start on different banks
__host__ __device__ inline int gr_row_elems(int n, int m) {
  const int w = 4 * gr_pad4(n) + 2 * gr_pad4(m);
  return ((w / 4) & 1) ? w : w + 4;
}

template <typename T>
struct GrSeg {
  const T* p;        // segment of problem 0 of the batch at this timestep
  long long stride;  // elements between consecutive problems
  int len, off;      // valid elements, offset inside the staged row
};

// the six (two at the final pseudo-step) segments of timestep t
template <typename T>
__device__ inline int gr_segments(const GradReduceArgs<T>& a, int t, GrSeg<T>* sg) {
  const int n = a.n, m = a.m, Th = a.T_, n4 = gr_pad4(n), m4 = gr_pad4(m);
  const long long sn = (long long)Th * n, sm = (long long)Th * m;
  if (t == Th) {
    sg[0] = GrSeg<T>{a.uqf, (long long)n, n, 0};
    sg[1] = GrSeg<T>{a.X + (long long)(Th - 1) * n, sn, n, n4};
    return 2;
  }
  sg[0] = GrSeg<T>{a.Nu + (long long)t * n, sn, n, 0};
  sg[1] = GrSeg<T>{a.ud + (long long)t * n, sn, n, n4};
  sg[2] = GrSeg<T>{a.uq + (long long)t * n, sn, n, 2 * n4};
  sg[3] = t > 0 ? GrSeg<T>{a.X + (long long)(t - 1) * n, sn, n, 3 * n4} : GrSeg<T>{a.x0, (long long)n, n, 3 * n4};  // sX[t] = x_t
  sg[4] = GrSeg<T>{a.ur + (long long)t * m, sm, m, 4 * n4};
  sg[5] = GrSeg<T>{a.U + (long long)t * m, sm, m, 4 * n4 + m4};
  return 6;
}

// Staging copies without per-chunk index arithmetic: the 16-byte chunks of ONE problem's segments are dealt to the 32 lanes of
// a warp once (chunk c of the concatenated segments -> lane c % 32, at most GR_PLAN chunks per lane), the warps of the CTA take
// the problems of a stage round-robin; per stage a lane only forms  base + problem * stride  per owned chunk.
constexpr int GR_PLAN = 3;  // <= 96 chunks per problem (n = 32, m = 16 in fp64: 80)
template <typename T>
struct GrCopyPlan {
  const T* base[GR_PLAN];      // chunk of problem 0 of the batch (nullptr: lane idle for this entry)
  long long stride[GR_PLAN];   // elements between consecutive problems of that segment
  int dst[GR_PLAN];            // element offset inside the staged row
  __device__ inline void init(const GrSeg<T>* sg, int nseg, int lane) {
    constexpr int E = 16 / (int)sizeof(T);
#pragma unroll
    for (int k = 0; k < GR_PLAN; k++) {
      base[k] = nullptr;
      stride[k] = 0;
      dst[k] = 0;
      int c = lane + 32 * k;  // chunk index inside the problem
      for (int sgi = 0; sgi < nseg; sgi++) {
        const int cps = sg[sgi].len / E;
        if (c >= 0 && c < cps) {
          base[k] = sg[sgi].p + c * E;
          stride[k] = sg[sgi].stride;
          dst[k] = sg[sgi].off + c * E;
          c = -1;
        } else if (c >= 0) {
          c -= cps;
        }
      }
    }
  }
  // problems bs .. bs + cnt - 1 into rows 0 .. cnt - 1 of `dstbuf` (row stride RW); warp `warp` of `nw` takes rows warp, warp + nw, ..
  __device__ inline void issue(T* dstbuf, int RW, long long bs, int cnt, int warp, int nw) const {
    for (int pl = warp; pl < cnt; pl += nw) {
#pragma unroll
      for (int k = 0; k < GR_PLAN; k++)
        if (base[k] != nullptr) cp_async<16>(smem_u32(dstbuf + pl * RW + dst[k]), base[k] + (bs + pl) * stride[k]);
    }
    cp_async_commit();
  }
};

// A tile task in terms of staged-row offsets: acc[i][j] += row[L[s] + i] * row[R[s] + j], s = 0, 1 (offset < 0: term absent;
// R < 0 with `ones`: the vector-sum tasks multiply by (1, 0, 0, 0)).  Edge tiles read the zero padding of their segment.
struct GrTask {
  int L[2], R[2];
  int out, ld, nr, nc;  // partial-row offset of element (0,0), row stride, valid rows / columns
  bool ones;
};

__host__ __device__ inline GrTask gr_decode(int n, int m, int Th, int t, int task) {
  const int rn = cdiv(n, GR_TILE), rm = cdiv(m, GR_TILE), ct = rn + rm;
  const int n4 = gr_pad4(n), m4 = gr_pad4(m);
  const int oNu = 0, oUNu = n4, oSux = 2 * n4, oSX = 3 * n4, oUU = 4 * n4, oU = 4 * n4 + m4;
  GrTask k;
  k.L[0] = k.L[1] = k.R[0] = k.R[1] = -1;
  k.ones = false;
  k.nr = k.nc = 0;
  k.out = 0;
  k.ld = 1;
  if (t == Th) {  // final pseudo-step: Pf = uX[T-1] (x) X[T-1], sum of uX[T-1]   (row = [uX[T-1] | X[T-1]])
    if (task < rn * rn) {
      const int i0 = (task / rn) * GR_TILE, j0 = (task % rn) * GR_TILE;
      k.L[0] = i0;
      k.R[0] = n4 + j0;
      k.out = i0 * n + j0; k.ld = n; k.nr = min(GR_TILE, n - i0); k.nc = min(GR_TILE, n - j0);
    } else {
      const int i0 = (task - rn * rn) * GR_TILE;
      k.L[0] = i0;
      k.ones = true;
      k.out = n * n + i0; k.ld = 1; k.nr = min(GR_TILE, n - i0); k.nc = 1;
    }
    return k;
  }
  const int nP = rn * ct;
  if (task < 2 * nP) {
    const bool p2 = task >= nP;
    const int tt = p2 ? task - nP : task;
    const int i0 = (tt / ct) * GR_TILE, cti = tt % ct;
    const bool mcol = cti >= rn;
    const int j0 = (mcol ? cti - rn : cti) * GR_TILE;
    if (!p2) {  // [gA | gB] = Nu (x) [sux | uU] + uNu (x) [sX | U]
      k.L[0] = oNu + i0;
      k.L[1] = oUNu + i0;
      k.R[0] = mcol ? oUU + j0 : oSux + j0;
      k.R[1] = mcol ? oU + j0 : oSX + j0;
    } else {    // [gQraw | gM] = sux (x) [sX | U] + sX (x) [0 | uU]
      k.L[0] = oSux + i0;
      k.R[0] = mcol ? oU + j0 : oSX + j0;
      if (mcol) {
        k.L[1] = oSX + i0;
        k.R[1] = oUU + j0;
      }
    }
    k.out = (p2 ? n * (n + m) : 0) + i0 * (n + m) + (mcol ? n : 0) + j0;
    k.ld = n + m; k.nr = min(GR_TILE, n - i0); k.nc = min(GR_TILE, (mcol ? m : n) - j0);
    return k;
  }
  int tt = task - 2 * nP;
  const int o3 = 2 * n * (n + m);
  if (tt < rm * rm) {  // gRraw = uU (x) U
    const int i0 = (tt / rm) * GR_TILE, j0 = (tt % rm) * GR_TILE;
    k.L[0] = oUU + i0;
    k.R[0] = oU + j0;
    k.out = o3 + i0 * m + j0; k.ld = m; k.nr = min(GR_TILE, m - i0); k.nc = min(GR_TILE, m - j0);
    return k;
  }
  tt -= rm * rm;
  k.ones = true;
  k.ld = 1; k.nc = 1;
  if (tt < rn) {            // gd = sum uNu
    k.L[0] = oUNu + tt * GR_TILE;
    k.out = o3 + m * m + tt * GR_TILE; k.nr = min(GR_TILE, n - tt * GR_TILE);
  } else if (tt < 2 * rn) {  // gq = sum sux
    tt -= rn;
    k.L[0] = oSux + tt * GR_TILE;
    k.out = o3 + m * m + n + tt * GR_TILE; k.nr = min(GR_TILE, n - tt * GR_TILE);
  } else {                   // gr = sum uU
    tt -= 2 * rn;
    k.L[0] = oUU + tt * GR_TILE;
    k.out = o3 + m * m + 2 * n + tt * GR_TILE; k.nr = min(GR_TILE, m - tt * GR_TILE);
  }
  return k;
}

template <typename T>
__device__ __forceinline__ void gr_lds4(const T* p, T* v) {
  if constexpr (sizeof(T) == 4) {
    const float4 x = *reinterpret_cast<const float4*>(p);
    v[0] = x.x; v[1] = x.y; v[2] = x.z; v[3] = x.w;
  } else {
    const double2 x = *reinterpret_cast<const double2*>(p), y = *(reinterpret_cast<const double2*>(p) + 1);
    v[0] = x.x; v[1] = x.y; v[2] = y.x; v[3] = y.y;
  }
}

// VEC: n and m are multiples of 4, every segment moves as 16-byte cp.async chunks; otherwise element-wise copies and the
// padding of every staged row is zeroed once (edge tiles multiply it)
template <typename T, bool VEC>
__global__ void __launch_bounds__(GR_WARPS * 32) grad_reduce_partial_kernel(const GradReduceArgs<T> a) {
  constexpr int SLOTS = 32 * GR_MAXT, EL = GR_TILE * GR_TILE, NTH = GR_WARPS * 32;
  extern __shared__ __align__(16) char gr_smem[];
  const int t = blockIdx.x, chunk = blockIdx.y, group = blockIdx.z;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int n = a.n, m = a.m, Th = a.T_;
  const int ntask = t == Th ? gr_tasks_final(n) : gr_tasks_step(n, m);
  // a warp's SLOTS accumulator slots (MAXT per lane) hold `ntg` tasks for each of `psplit` problems walked side by side
  const int ntg = min(SLOTS, ntask - group * SLOTS);  // tasks of this tile group
  if (ntg <= 0) return;  // the final pseudo-step has fewer tasks than a regular step: its upper tile groups are empty
  const int psplit = SLOTS / ntg;
  const long long per = (a.nb + a.NC - 1) / a.NC;
  const long long b0 = (long long)chunk * per, b1 = b0 + per < a.nb ? b0 + per : a.nb;
  const int RW = gr_row_elems(n, m), GR_STAGE = a.stage;
  T* stage = reinterpret_cast<T*>(gr_smem);                       // [2][GR_STAGE][RW]
  T* red = stage + 2 * GR_STAGE * RW;                             // [GR_WARPS][EL][SLOTS] (aliases nothing: own range)

  GrSeg<T> sg[6];
  const int nseg = gr_segments<T>(a, t, sg);
  if (!VEC) {  // zero every staged row once: the copies below only write the valid elements
    for (int e = tid; e < 2 * GR_STAGE * RW; e += NTH) stage[e] = T(0);
    __syncthreads();
  }
  GrCopyPlan<T> plan;
  bool use_plan = false;
  if (VEC) {
    int cpp = 0;
    for (int sgi = 0; sgi < nseg; sgi++) cpp += sg[sgi].len / (16 / (int)sizeof(T));
    use_plan = cpp <= 32 * GR_PLAN;
    if (use_plan) plan.init(sg, nseg, lane);
  }
  auto issue = [&](long long bs, int buf) {  // stage problems bs .. bs + GR_STAGE - 1 (clamped to the chunk)
    T* dst = stage + buf * GR_STAGE * RW;
    const int cnt = (int)(b1 - bs < GR_STAGE ? b1 - bs : GR_STAGE);
    if constexpr (VEC) {
      if (use_plan) {
        plan.issue(dst, RW, bs, cnt, warp, GR_WARPS);
        return;
      }
      constexpr int E = 16 / (int)sizeof(T);
      for (int sgi = 0; sgi < nseg; sgi++) {  // large shapes (more chunks per problem than the plan holds)
        const int cps = sg[sgi].len / E;
        for (int c = tid; c < cnt * cps; c += NTH) {
          const int pl = c / cps, cc = c % cps;
          cp_async<16>(smem_u32(dst + pl * RW + sg[sgi].off + cc * E), sg[sgi].p + (bs + pl) * sg[sgi].stride + cc * E);
        }
      }
    } else {
      for (int sgi = 0; sgi < nseg; sgi++) {
        const int len = sg[sgi].len;
        for (int c = tid; c < cnt * len; c += NTH) {
          const int pl = c / len, cc = c % len;
          cp_async<(int)sizeof(T)>(smem_u32(dst + pl * RW + sg[sgi].off + cc), sg[sgi].p + (bs + pl) * sg[sgi].stride + cc);
        }
      }
    }
    cp_async_commit();
  };

  GrTask task[GR_MAXT];
  int sub[GR_MAXT];
  T acc[GR_MAXT][GR_TILE][GR_TILE];
#pragma unroll
  for (int k = 0; k < GR_MAXT; k++) {
    const int slot = k * 32 + lane;
    sub[k] = slot / ntg;
    task[k] = gr_decode(n, m, Th, t, group * SLOTS + slot % ntg);
    if (sub[k] >= psplit) task[k].L[0] = task[k].L[1] = -1;  // idle slot
#pragma unroll
    for (int i = 0; i < GR_TILE; i++)
#pragma unroll
      for (int j = 0; j < GR_TILE; j++) acc[k][i][j] = T(0);
  }
  const int stride = GR_WARPS * psplit;
  const int nstage = (int)((b1 - b0 + GR_STAGE - 1) / GR_STAGE);
  if (nstage > 0) issue(b0, 0);
  for (int st = 0; st < nstage; st++) {
    const long long bs = b0 + (long long)st * GR_STAGE;
    if (st + 1 < nstage) {
      issue(bs + GR_STAGE, (st + 1) & 1);
      cp_async_wait<1>();
    } else {
      cp_async_wait<0>();
    }
    __syncthreads();
    const T* buf = stage + (st & 1) * GR_STAGE * RW;
    const int cnt = (int)(b1 - bs < GR_STAGE ? b1 - bs : GR_STAGE);
    for (int pb = warp * psplit; pb < cnt; pb += stride) {
#pragma unroll
      for (int k = 0; k < GR_MAXT; k++) {
        const int pl = pb + sub[k];
        if (pl >= cnt) continue;
        const T* row = buf + pl * RW;
#pragma unroll
        for (int s = 0; s < 2; s++) {
          if (task[k].L[s] < 0) continue;
          T l[GR_TILE], r[GR_TILE];
          gr_lds4<T>(row + task[k].L[s], l);
          if (task[k].ones) {
            r[0] = T(1); r[1] = r[2] = r[3] = T(0);
          } else {
            gr_lds4<T>(row + task[k].R[s], r);
          }
#pragma unroll
          for (int i = 0; i < GR_TILE; i++)
#pragma unroll
            for (int j = 0; j < GR_TILE; j++) acc[k][i][j] += l[i] * r[j];
        }
      }
    }
    __syncthreads();  // the buffer is refilled two stages later: everyone is done reading it
  }
  // every slot's sums go through shared memory and are added in (warp, sub-batch) order: deterministic
#pragma unroll
  for (int k = 0; k < GR_MAXT; k++)
#pragma unroll
    for (int i = 0; i < GR_TILE; i++)
#pragma unroll
      for (int j = 0; j < GR_TILE; j++) red[(warp * EL + i * GR_TILE + j) * SLOTS + k * 32 + lane] = acc[k][i][j];
  __syncthreads();
  T* prow = a.part + (long long)chunk * gr_part_elems(Th, n, m) + (long long)t * gr_per_step(n, m);
  for (int x = tid; x < ntg * EL; x += NTH) {
    const int tl = x / EL, el = x % EL, i = el / GR_TILE, j = el % GR_TILE;
    const GrTask tk = gr_decode(n, m, Th, t, group * SLOTS + tl);
    if (i >= tk.nr || j >= tk.nc) continue;
    T v = T(0);
    for (int w = 0; w < GR_WARPS; w++)
      for (int sb = 0; sb < psplit; sb++) v += red[(w * EL + el) * SLOTS + sb * ntg + tl];
    prow[tk.out + i * tk.ld + j] = v;
  }
}
// problems per stage and dynamic shared memory of the partial kernel (two staged buffers + the cross-warp reduction area)
inline int gr_stage_problems(int n, int m, size_t elem) {
  const size_t red = (size_t)GR_WARPS * GR_TILE * GR_TILE * 32 * GR_MAXT * elem, budget = 150 * 1024;
  const size_t fit = (budget - red) / ((size_t)2 * gr_row_elems(n, m) * elem);
  return (int)(fit > (size_t)GR_STAGE_MAX ? GR_STAGE_MAX : (fit < 1 ? 1 : fit));
}
inline size_t gr_partial_smem(int n, int m, size_t elem) {
  return ((size_t)2 * gr_stage_problems(n, m, elem) * gr_row_elems(n, m) + (size_t)GR_WARPS * GR_TILE * GR_TILE * 32 * GR_MAXT) * elem;
}

// one thread per output element: ordered sum over the NC chunk partials, symmetrisation, scatter into the leaves
template <typename T>
__global__ void grad_reduce_final_kernel(const GradReduceArgs<T> a) {
  const int n = a.n, m = a.m, Th = a.T_;
  const int PS = gr_per_step(n, m);
  const long long PE = gr_part_elems(Th, n, m);
  const long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= PE) return;
  auto sum = [&](long long off) {
    T s = T(0);
    for (int c = 0; c < a.NC; c++) s += a.part[(long long)c * PE + off];
    return s;
  };
  const long long t = e / PS;
  const int w = (int)(e - t * PS);
  const long long base = t * PS;
  if (t == Th) {  // [Pf (n x n) | sum uX[T-1]]
    if (w < n * n) {
      const int i = w / n, j = w % n;
      a.gQf[w] = T(0.5) * (sum(base + w) + sum(base + j * n + i));
    } else {
      a.gqf[w - n * n] = sum(base + w);
    }
    return;
  }
  const int nP = n * (n + m);
  if (w < nP) {
    const int i = w / (n + m), c = w % (n + m);
    if (c < n) a.gA[t * n * n + i * n + c] = sum(e); else a.gB[t * n * m + i * m + (c - n)] = sum(e);
  } else if (w < 2 * nP) {
    const int ww = w - nP, i = ww / (n + m), c = ww % (n + m);
    if (c < n) a.gQ[t * n * n + i * n + c] = T(0.5) * (sum(e) + sum(base + nP + c * (n + m) + i));
    else a.gM[t * n * m + i * m + (c - n)] = sum(e);
  } else if (w < 2 * nP + m * m) {
    const int ww = w - 2 * nP, i = ww / m, j = ww % m;
    a.gR[t * m * m + ww] = T(0.5) * (sum(e) + sum(base + 2 * nP + j * m + i));
  } else {
    const int ww = w - 2 * nP - m * m;
    if (ww < n) a.gd[t * n + ww] = sum(e);
    else if (ww < 2 * n) a.gq[t * n + (ww - n)] = sum(e);
    else a.gr[t * m + (ww - 2 * n)] = sum(e);
  }
}

}  // namespace idoc
