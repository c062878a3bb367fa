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
//   grad_reduce_partial_kernel : grid (T+1, NC chunks of the batch, tile groups).  A lane owns up to MAXT 4x4 register
//                                tiles of the step's outputs and walks its warp's share of the chunk's problems, reading
//                                the operand vectors straight from global memory (read-only path): 4 x 16-byte loads per
//                                32 multiply-adds.  The four warps of a CTA are summed in a fixed order through shared
//                                memory and the CTA writes one partial [t][PER_STEP] row of chunk c.
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

// one operand of a rank-1 term: element i of problem b is p[b * stride + min(i, len - 1)]
template <typename T>
struct GrOperand {
  const T* p;
  long long stride;
  int len;  // valid elements from p on (edge tiles clamp their loads)
};
template <typename T>
struct GrTask {
  GrOperand<T> L[2], R[2];  // acc[i][j] += L[s][i] * R[s][j], s = 0, 1 (p == nullptr: term absent)
  int out, ld, nr, nc;      // partial-row offset of element (0,0), row stride, valid rows / columns
  bool ones;                // vector-sum task: R = (1, 0, 0, 0)
};

template <typename T>
__device__ inline GrTask<T> gr_decode(const GradReduceArgs<T>& a, int t, int task) {
  const int n = a.n, m = a.m, Th = a.T_;
  const int rn = cdiv(n, GR_TILE), rm = cdiv(m, GR_TILE), ct = rn + rm;
  const long long sn = (long long)Th * n, sm = (long long)Th * m;
  GrTask<T> k;
  k.L[0].p = k.L[1].p = k.R[0].p = k.R[1].p = nullptr;
  k.ones = false;
  k.nr = k.nc = 0;
  auto vec = [&](const T* base, long long stride, int off, int len) { return GrOperand<T>{base + off, stride, len - off}; };
  if (t == Th) {  // final pseudo-step: Pf = uX[T-1] (x) X[T-1], sum of uX[T-1]
    const T* XT = a.X + (long long)(Th - 1) * n;
    if (task < rn * rn) {
      const int i0 = (task / rn) * GR_TILE, j0 = (task % rn) * GR_TILE;
      k.L[0] = vec(a.uqf, n, i0, n);
      k.R[0] = vec(XT, sn, j0, n);
      k.out = i0 * n + j0; k.ld = n; k.nr = min(GR_TILE, n - i0); k.nc = min(GR_TILE, n - j0);
    } else {
      const int i0 = (task - rn * rn) * GR_TILE;
      k.L[0] = vec(a.uqf, n, i0, n);
      k.ones = true;
      k.out = n * n + i0; k.ld = 1; k.nr = min(GR_TILE, n - i0); k.nc = 1;
    }
    return k;
  }
  const T* Nu = a.Nu + (long long)t * n;
  const T* uNu = a.ud + (long long)t * n;
  const T* sux = a.uq + (long long)t * n;
  const T* U = a.U + (long long)t * m;
  const T* uU = a.ur + (long long)t * m;
  const T* sX = t > 0 ? a.X + (long long)(t - 1) * n : a.x0;  // sX[t] = x_t = X[t-1], x0 at t = 0
  const long long ssX = t > 0 ? sn : (long long)n;
  const int nP = rn * ct;
  if (task < 2 * nP) {
    const bool p2 = task >= nP;
    const int tt = p2 ? task - nP : task;
    const int i0 = (tt / ct) * GR_TILE, cti = tt % ct;
    const bool mcol = cti >= rn;
    const int j0 = (mcol ? cti - rn : cti) * GR_TILE;
    if (!p2) {  // [gA | gB] = Nu (x) [sux | uU] + uNu (x) [sX | U]
      k.L[0] = vec(Nu, sn, i0, n);
      k.L[1] = vec(uNu, sn, i0, n);
      k.R[0] = mcol ? vec(uU, sm, j0, m) : vec(sux, sn, j0, n);
      k.R[1] = mcol ? vec(U, sm, j0, m) : vec(sX, ssX, j0, n);
    } else {    // [gQraw | gM] = sux (x) [sX | U] + sX (x) [0 | uU]
      k.L[0] = vec(sux, sn, i0, n);
      k.R[0] = mcol ? vec(U, sm, j0, m) : vec(sX, ssX, j0, n);
      if (mcol) {
        k.L[1] = vec(sX, ssX, i0, n);
        k.R[1] = vec(uU, sm, j0, m);
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
    k.L[0] = vec(uU, sm, i0, m);
    k.R[0] = vec(U, sm, j0, m);
    k.out = o3 + i0 * m + j0; k.ld = m; k.nr = min(GR_TILE, m - i0); k.nc = min(GR_TILE, m - j0);
    return k;
  }
  tt -= rm * rm;
  k.ones = true;
  k.ld = 1; k.nc = 1;
  if (tt < rn) {            // gd = sum uNu
    k.L[0] = vec(uNu, sn, tt * GR_TILE, n);
    k.out = o3 + m * m + tt * GR_TILE; k.nr = min(GR_TILE, n - tt * GR_TILE);
  } else if (tt < 2 * rn) {  // gq = sum sux
    tt -= rn;
    k.L[0] = vec(sux, sn, tt * GR_TILE, n);
    k.out = o3 + m * m + n + tt * GR_TILE; k.nr = min(GR_TILE, n - tt * GR_TILE);
  } else {                   // gr = sum uU
    tt -= 2 * rn;
    k.L[0] = vec(uU, sm, tt * GR_TILE, m);
    k.out = o3 + m * m + 2 * n + tt * GR_TILE; k.nr = min(GR_TILE, m - tt * GR_TILE);
  }
  return k;
}

template <typename T, bool VEC>
__device__ __forceinline__ void gr_load4(const GrOperand<T>& o, long long b, T* v) {
  const T* p = o.p + b * o.stride;
  if constexpr (VEC) {
    if constexpr (sizeof(T) == 4) {
      const float4 x = __ldg(reinterpret_cast<const float4*>(p));
      v[0] = x.x; v[1] = x.y; v[2] = x.z; v[3] = x.w;
    } else {
      const double2 x = __ldg(reinterpret_cast<const double2*>(p)), y = __ldg(reinterpret_cast<const double2*>(p) + 1);
      v[0] = x.x; v[1] = x.y; v[2] = y.x; v[3] = y.y;
    }
  } else {
#pragma unroll
    for (int i = 0; i < GR_TILE; i++) v[i] = __ldg(p + (i < o.len ? i : o.len - 1));
  }
}

template <typename T, bool VEC>
__global__ void __launch_bounds__(GR_WARPS * 32) grad_reduce_partial_kernel(const GradReduceArgs<T> a) {
  constexpr int SLOTS = 32 * GR_MAXT, EL = GR_TILE * GR_TILE;
  __shared__ T red[GR_WARPS][EL][SLOTS];
  const int t = blockIdx.x, chunk = blockIdx.y, group = blockIdx.z;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int ntask = t == a.T_ ? gr_tasks_final(a.n) : gr_tasks_step(a.n, a.m);
  // a warp's SLOTS accumulator slots (MAXT per lane) hold `ntg` tasks for each of `psplit` problems walked side by side
  const int ntg = min(SLOTS, ntask - group * SLOTS);  // tasks of this tile group
  if (ntg <= 0) return;  // the final pseudo-step has fewer tasks than a regular step: its upper tile groups are empty
  const int psplit = SLOTS / ntg;
  const long long per = (a.nb + a.NC - 1) / a.NC;
  const long long b0 = (long long)chunk * per, b1 = b0 + per < a.nb ? b0 + per : a.nb;

  GrTask<T> task[GR_MAXT];
  int sub[GR_MAXT];
  T acc[GR_MAXT][GR_TILE][GR_TILE];
#pragma unroll
  for (int k = 0; k < GR_MAXT; k++) {
    const int slot = k * 32 + lane;
    sub[k] = slot / ntg;
    task[k] = gr_decode<T>(a, t, group * SLOTS + slot % ntg);
    if (sub[k] >= psplit) task[k].L[0].p = task[k].L[1].p = nullptr;  // idle slot
#pragma unroll
    for (int i = 0; i < GR_TILE; i++)
#pragma unroll
      for (int j = 0; j < GR_TILE; j++) acc[k][i][j] = T(0);
  }
  const int stride = GR_WARPS * psplit;
#pragma unroll 2
  for (long long bb = b0 + warp * psplit; bb < b1; bb += stride) {
#pragma unroll
    for (int k = 0; k < GR_MAXT; k++) {
      const long long b = bb + sub[k];
      if (b >= b1) continue;
#pragma unroll
      for (int s = 0; s < 2; s++) {
        if (task[k].L[s].p == nullptr) continue;
        T l[GR_TILE], r[GR_TILE];
        gr_load4<T, VEC>(task[k].L[s], b, l);
        if (task[k].ones) {
          r[0] = T(1); r[1] = r[2] = r[3] = T(0);
        } else {
          gr_load4<T, VEC>(task[k].R[s], b, r);
        }
#pragma unroll
        for (int i = 0; i < GR_TILE; i++)
#pragma unroll
          for (int j = 0; j < GR_TILE; j++) acc[k][i][j] += l[i] * r[j];
      }
    }
  }
  // every slot's sums go through shared memory and are added in (warp, sub-batch) order: deterministic
#pragma unroll
  for (int k = 0; k < GR_MAXT; k++)
#pragma unroll
    for (int i = 0; i < GR_TILE; i++)
#pragma unroll
      for (int j = 0; j < GR_TILE; j++) red[warp][i * GR_TILE + j][k * 32 + lane] = acc[k][i][j];
  __syncthreads();
  T* row = a.part + (long long)chunk * gr_part_elems(a.T_, a.n, a.m) + (long long)t * gr_per_step(a.n, a.m);
  for (int x = threadIdx.x; x < ntg * EL; x += GR_WARPS * 32) {
    const int tl = x / EL, el = x % EL, i = el / GR_TILE, j = el % GR_TILE;
    const GrTask<T> tk = gr_decode<T>(a, t, group * SLOTS + tl);
    if (i >= tk.nr || j >= tk.nc) continue;
    T v = T(0);
    for (int w = 0; w < GR_WARPS; w++)
      for (int sb = 0; sb < psplit; sb++) v += red[w][el][sb * ntg + tl];
    row[tk.out + i * tk.ld + j] = v;
  }
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
