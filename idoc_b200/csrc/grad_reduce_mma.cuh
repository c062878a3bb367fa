// Batch-summed gradients on the tensor cores: the batch contraction of grad_reduce.cuh as warp-level MMAs.
//
// Summed over the batch, every gradient block of a timestep is a GEMM whose contraction axis is the batch (grad_reduce.cuh):
//     P1 = [gA | gB]      = sum_b  Nu_b (x) [sux_b | uU_b]  +  uNu_b (x) [sX_b | U_b]          (n x (n + m), K = 2 B)
//     P2 = [gQraw | gM]   = sum_b  sux_b (x) [sX_b | U_b]   +  sX_b (x) [0 | uU_b]
//     P3 = gRraw          = sum_b  uU_b (x) U_b                                                  (m x m,       K = B)
// i.e. the dense contraction north_star reserves the tensor cores for.  One MMA consumes the staged vectors of FOUR problems
// (K = 2 terms x 2 problems per DMMA m8n8k4 in fp64; K = 2 terms x 4 problems per m16n8k8 TF32 MMA in fp32):
//   fp64: mma.sync.aligned.m8n8k4.row.col.f64 (SASS DMMA) — exact FP64 multiply-adds, 2 operand scalars per lane per 256 FMAs
//         against 8 per 16 FMAs with 4x4 register tiles (the FFMA kernel is bound by its shared-memory loads);
//   fp32: mma.sync.aligned.m16n8k8 TF32 with the 3xTF32 split of mid_mma.cuh (x = hi + lo, lo*hi + hi*lo + hi*hi): FP32-level
//         accuracy, 6 operand scalars per lane per 1024 FMAs.
// Staging (16-byte cp.async, double-buffered), chunking over the batch, the partial-row layout and the ordered final reduction
// are those of grad_reduce.cuh (same `part` buffer, same grad_reduce_final_kernel): only the accumulation changes pipe.  The four
// warps of a CTA take the staged problems in interleaved groups of four and are added in warp order: bit-reproducible.
// Shapes: n a multiple of 8 (<= 32), m in {4, 8, 16}; everything else keeps the FFMA kernel.
#pragma once
#include "grad_reduce.cuh"
#include "mid_mma.cuh"  // TF32 helpers

namespace idoc {


__host__ __device__ inline int gr_pad8(int w) { return (w + 7) / 8 * 8; }
// staged row of the MMA kernels: the six segments padded to multiples of 8 (+ one 8-element pad for an odd stride in 32-byte
// units); final pseudo-step: [uX[T-1] | X[T-1]]
__host__ __device__ inline int grm_row_elems(int n, int m) {
  const int w = 4 * gr_pad8(n) + 2 * gr_pad8(m);
  return ((w / 8) & 1) ? w : w + 8;
}
constexpr int GRM_STAGE = 32;  // problems per stage (a multiple of 4 x GR_WARPS / 2)
inline size_t grm_smem_bytes(int n, int m, size_t elem, int acc_per_lane) {
  // two staged buffers, re-used after the last stage as the cross-warp reduction area [GR_WARPS][acc_per_lane][32], + the
  // vector-sum area [GR_WARPS][row]
  const size_t stage = (size_t)2 * GRM_STAGE * grm_row_elems(n, m), red = (size_t)GR_WARPS * acc_per_lane * 32;
  return ((stage > red ? stage : red) + (size_t)GR_WARPS * grm_row_elems(n, m)) * elem;
}

// One tile block of a step: rows r0 .. r0 + 8 RT, column tiles of 8, operand offsets inside the staged row
struct GrmBlock {
  int L[2], R[2];   // staged-row offsets of the two terms' operands (offset < 0: term absent)
  int out, ld;      // partial-row offset of element (0, 0) of the block, row stride
  int rows, cols;   // valid rows / columns (stores are masked)
};

// the blocks of timestep t (t == T: the final pseudo-step); returns their number (<= 5)
__device__ inline int grm_blocks(int n, int m, int Th, int t, GrmBlock* b) {
  const int n8 = gr_pad8(n), m8 = gr_pad8(m);
  const int oNu = 0, oUNu = n8, oSux = 2 * n8, oSX = 3 * n8, oUU = 4 * n8, oU = 4 * n8 + m8;
  if (t == Th) {  // Pf = uX[T-1] (x) X[T-1]
    b[0] = GrmBlock{{0, -1}, {n8, -1}, 0, n, n, n};
    return 1;
  }
  const int nP = n * (n + m);
  b[0] = GrmBlock{{oNu, oUNu}, {oSux, oSX}, 0, n + m, n, n};                    // gA
  b[1] = GrmBlock{{oNu, oUNu}, {oUU, oU}, n, n + m, n, m};                      // gB
  b[2] = GrmBlock{{oSux, -1}, {oSX, -1}, nP, n + m, n, n};                      // gQraw
  b[3] = GrmBlock{{oSux, oSX}, {oU, oUU}, nP + n, n + m, n, m};                 // gM
  b[4] = GrmBlock{{oUU, -1}, {oU, -1}, 2 * nP, m, m, m};                        // gRraw
  return 5;
}

// ------------------------------------------------------------------------------------------------ fp64: DMMA m8n8k4
// N8 = ceil(n / 8), M8 = ceil(m / 8); ZS: the row tiles of the blocks are dealt round-robin to ZS CTAs (blockIdx.z) so that
// the accumulators of one warp fit the register file (n = 32: 52 tiles of 2 doubles per lane -> 2 x 28)
template <int N8, int M8, int ZS>
__global__ void __launch_bounds__(GR_WARPS * 32) grad_reduce_dmma_kernel(const GradReduceArgs<double> a) {
  using T = double;
  constexpr int NTH = GR_WARPS * 32;
  constexpr int RT = (N8 + ZS - 1) / ZS;                                     // row tiles of an n-row block per CTA
  constexpr int NACC = 2 * (RT * N8 + RT * M8 + RT * N8 + RT * M8 + M8 * M8);  // doubles per lane
  extern __shared__ __align__(16) char gr_smem[];
  const int t = blockIdx.x, chunk = blockIdx.y, zs = blockIdx.z;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, g = lane >> 2, kk = lane & 3;
  const int n = a.n, m = a.m, Th = a.T_;
  const long long per = (a.nb + a.NC - 1) / a.NC;
  const long long b0 = (long long)chunk * per, b1 = b0 + per < a.nb ? b0 + per : a.nb;
  const int RW = grm_row_elems(n, m);
  T* stage = reinterpret_cast<T*>(gr_smem);           // [2][GRM_STAGE][RW]
  T* red = stage;                                     // [GR_WARPS][NACC][32]: the staged buffers are dead by then
  T* vsum = stage + (2 * GRM_STAGE * RW > GR_WARPS * NACC * 32 ? 2 * GRM_STAGE * RW : GR_WARPS * NACC * 32);  // [GR_WARPS][RW]
  GrmBlock blk[5];
  const int nblk = grm_blocks(n, m, Th, t, blk);
  GrSeg<T> sg[6];
  const int nseg = gr_segments<T>(a, t, sg);
  {  // staged-row offsets of the MMA layout (segments padded to 8)
    const int n8 = gr_pad8(n), m8 = gr_pad8(m);
    if (t == Th) { sg[0].off = 0; sg[1].off = n8; }
    else { sg[0].off = 0; sg[1].off = n8; sg[2].off = 2 * n8; sg[3].off = 3 * n8; sg[4].off = 4 * n8; sg[5].off = 4 * n8 + m8; }
  }
  for (int e = tid; e < 2 * GRM_STAGE * RW; e += NTH) stage[e] = T(0);  // pads (and unused rows) must read as zero
  __syncthreads();
  GrCopyPlan<T> plan;
  plan.init(sg, nseg, lane);
  auto issue = [&](long long bs, int buf) {
    const int cnt = (int)(b1 - bs < GRM_STAGE ? b1 - bs : GRM_STAGE);
    plan.issue(stage + buf * GRM_STAGE * RW, RW, bs, cnt, warp, GR_WARPS);
  };
  T acc[NACC / 2][2];
#pragma unroll
  for (int i = 0; i < NACC / 2; i++) acc[i][0] = acc[i][1] = T(0);
  T vs[8];  // vector sums: elements lane, lane + 32, ... of the staged row (RW <= 200 at n = 32, m = 16)
#pragma unroll
  for (int i = 0; i < 8; i++) vs[i] = T(0);
  const int nstage = (int)((b1 - b0 + GRM_STAGE - 1) / GRM_STAGE);
  if (nstage > 0) issue(b0, 0);
  for (int st = 0; st < nstage; st++) {
    const long long bs = b0 + (long long)st * GRM_STAGE;
    if (st + 1 < nstage) {
      issue(bs + GRM_STAGE, (st + 1) & 1);
      cp_async_wait<1>();
    } else {
      cp_async_wait<0>();
    }
    __syncthreads();
    T* buf = stage + (st & 1) * GRM_STAGE * RW;
    const int cnt = (int)(b1 - bs < GRM_STAGE ? b1 - bs : GRM_STAGE);
    if (cnt < GRM_STAGE) {  // last stage of the chunk: rows beyond the batch hold an earlier stage's problems
      for (int e = tid; e < (GRM_STAGE - cnt) * RW; e += NTH) buf[cnt * RW + e] = T(0);
      __syncthreads();
    }
    for (int q = warp; q * 4 < cnt; q += GR_WARPS) {  // this warp's groups of four problems
      const T* rows = buf + (q * 4) * RW;
      int ai = 0;
#pragma unroll
      for (int bi = 0; bi < 5; bi++) {
        if (bi >= nblk) break;
        const bool two = blk[bi].L[1] >= 0;
        const bool mrows = bi == 4;                       // gRraw: m rows
        const bool mcols = bi == 1 || bi == 3 || bi == 4;  // m columns
        const int RTB = mrows ? M8 : RT, CTB = mcols ? M8 : N8;
#pragma unroll
        for (int rt = 0; rt < (N8 > M8 ? RT : (RT > M8 ? RT : M8)); rt++) {
          if (rt >= RTB) break;
          const int rtile = mrows ? rt : rt * ZS + zs;  // global row tile (m-row blocks live in CTA 0 only)
          const bool live = mrows ? (zs == 0) : (rtile < N8);
          // A operands of the two DMMAs of this group: DMMA j covers problems 2 j, 2 j + 1 (two-term) or all four (one-term, j = 0)
          T a0 = T(0), a1 = T(0);
          if (live) {
            if (two) {
              a0 = rows[(kk >> 1) * RW + blk[bi].L[kk & 1] + rtile * 8 + g];
              a1 = rows[(2 + (kk >> 1)) * RW + blk[bi].L[kk & 1] + rtile * 8 + g];
            } else {
              a0 = rows[kk * RW + blk[bi].L[0] + rtile * 8 + g];
            }
          }
#pragma unroll
          for (int ct = 0; ct < (N8 > M8 ? N8 : M8); ct++) {
            if (ct >= CTB) break;
            if (live) {
              if (two) {
                const T b0v = rows[(kk >> 1) * RW + blk[bi].R[kk & 1] + ct * 8 + g];
                const T b1v = rows[(2 + (kk >> 1)) * RW + blk[bi].R[kk & 1] + ct * 8 + g];
                dmma884(acc[ai], a0, b0v);
                dmma884(acc[ai], a1, b1v);
              } else {
                const T b0v = rows[kk * RW + blk[bi].R[0] + ct * 8 + g];
                dmma884(acc[ai], a0, b0v);
              }
            }
            ai++;
          }
        }
      }
      // vector sums over the same four problems
#pragma unroll
      for (int i = 0; i < 8; i++) {
        const int e = lane + 32 * i;
        if (e < RW) vs[i] += (rows[e] + rows[RW + e]) + (rows[2 * RW + e] + rows[3 * RW + e]);
      }
    }
    __syncthreads();
  }
  // cross-warp sums in warp order, then the partial row
#pragma unroll
  for (int i = 0; i < NACC / 2; i++) {
    red[(warp * NACC + 2 * i) * 32 + lane] = acc[i][0];
    red[(warp * NACC + 2 * i + 1) * 32 + lane] = acc[i][1];
  }
#pragma unroll
  for (int i = 0; i < 8; i++)
    if (lane + 32 * i < RW) vsum[warp * RW + lane + 32 * i] = vs[i];
  __syncthreads();
  T* prow = a.part + (long long)chunk * gr_part_elems(Th, n, m) + (long long)t * gr_per_step(n, m);
  if (warp == 0) {
    int ai = 0;
    for (int bi = 0; bi < nblk; bi++) {
      const bool mrows = bi == 4, mcols = bi == 1 || bi == 3 || bi == 4;
      const int RTB = mrows ? M8 : RT, CTB = mcols ? M8 : N8;
      for (int rt = 0; rt < RTB; rt++) {
        const int rtile = mrows ? rt : rt * ZS + zs;
        const bool live = mrows ? (zs == 0) : (rtile < N8);
        for (int ct = 0; ct < CTB; ct++) {
          if (live) {
            const int row = rtile * 8 + g, col = ct * 8 + 2 * kk;
            for (int e = 0; e < 2; e++) {
              T v = T(0);
              for (int w = 0; w < GR_WARPS; w++) v += red[(w * NACC + 2 * ai + e) * 32 + lane];
              if (row < blk[bi].rows && col + e < blk[bi].cols) prow[blk[bi].out + row * blk[bi].ld + col + e] = v;
            }
          }
          ai++;
        }
      }
    }
  }
  if (zs == 0) {  // vector sums: gd = sum uNu, gq = sum sux, gr = sum uU (final pseudo-step: gqf = sum uX[T-1])
    const int n8 = gr_pad8(n);
    for (int e = tid; e < RW; e += NTH) {
      T v = T(0);
      for (int w = 0; w < GR_WARPS; w++) v += vsum[w * RW + e];
      if (t == Th) {
        if (e < n) prow[n * n + e] = v;
      } else {
        const int o3 = 2 * n * (n + m) + m * m;
        if (e >= n8 && e < n8 + n) prow[o3 + (e - n8)] = v;                   // uNu -> gd
        else if (e >= 2 * n8 && e < 2 * n8 + n) prow[o3 + n + (e - 2 * n8)] = v;  // sux -> gq
        else if (e >= 4 * n8 && e < 4 * n8 + m) prow[o3 + 2 * n + (e - 4 * n8)] = v;  // uU -> gr
      }
    }
  }
}

// ------------------------------------------------------------------------------------------------ fp32: 3xTF32 m16n8k8
// Row tiles of 16 (MT = ceil(n / 16)), column tiles of 8; one MMA consumes four problems x two terms (a one-term block feeds
// zeros as its second term).
template <int N8, int M8>
__global__ void __launch_bounds__(GR_WARPS * 32) grad_reduce_tf32_kernel(const GradReduceArgs<float> a) {
  using T = float;
  constexpr int NTH = GR_WARPS * 32;
  constexpr int MTN = (N8 + 1) / 2, MTM = (M8 + 1) / 2;
  constexpr int NTILE = MTN * N8 + MTN * M8 + MTN * N8 + MTN * M8 + MTM * M8;
  constexpr int NACC = 4 * NTILE;
  extern __shared__ __align__(16) char gr_smem[];
  const int t = blockIdx.x, chunk = blockIdx.y;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, g = lane >> 2, tig = lane & 3;
  const int n = a.n, m = a.m, Th = a.T_;
  const long long per = (a.nb + a.NC - 1) / a.NC;
  const long long b0 = (long long)chunk * per, b1 = b0 + per < a.nb ? b0 + per : a.nb;
  const int RW = grm_row_elems(n, m);
  T* stage = reinterpret_cast<T*>(gr_smem);
  T* red = stage;  // the staged buffers are dead by then
  T* vsum = stage + (2 * GRM_STAGE * RW > GR_WARPS * NACC * 32 ? 2 * GRM_STAGE * RW : GR_WARPS * NACC * 32);
  GrmBlock blk[5];
  const int nblk = grm_blocks(n, m, Th, t, blk);
  GrSeg<T> sg[6];
  const int nseg = gr_segments<T>(a, t, sg);
  {
    const int n8 = gr_pad8(n), m8 = gr_pad8(m);
    if (t == Th) { sg[0].off = 0; sg[1].off = n8; }
    else { sg[0].off = 0; sg[1].off = n8; sg[2].off = 2 * n8; sg[3].off = 3 * n8; sg[4].off = 4 * n8; sg[5].off = 4 * n8 + m8; }
  }
  for (int e = tid; e < 2 * GRM_STAGE * RW; e += NTH) stage[e] = T(0);
  __syncthreads();
  GrCopyPlan<T> plan;
  plan.init(sg, nseg, lane);
  auto issue = [&](long long bs, int buf) {
    const int cnt = (int)(b1 - bs < GRM_STAGE ? b1 - bs : GRM_STAGE);
    plan.issue(stage + buf * GRM_STAGE * RW, RW, bs, cnt, warp, GR_WARPS);
  };
  float acc[NTILE][4];
#pragma unroll
  for (int i = 0; i < NTILE; i++) acc[i][0] = acc[i][1] = acc[i][2] = acc[i][3] = 0.f;
  float vs[8];
#pragma unroll
  for (int i = 0; i < 8; i++) vs[i] = 0.f;
  const int nstage = (int)((b1 - b0 + GRM_STAGE - 1) / GRM_STAGE);
  if (nstage > 0) issue(b0, 0);
  for (int st = 0; st < nstage; st++) {
    const long long bs = b0 + (long long)st * GRM_STAGE;
    if (st + 1 < nstage) {
      issue(bs + GRM_STAGE, (st + 1) & 1);
      cp_async_wait<1>();
    } else {
      cp_async_wait<0>();
    }
    __syncthreads();
    T* buf = stage + (st & 1) * GRM_STAGE * RW;
    const int cnt = (int)(b1 - bs < GRM_STAGE ? b1 - bs : GRM_STAGE);
    if (cnt < GRM_STAGE) {
      for (int e = tid; e < (GRM_STAGE - cnt) * RW; e += NTH) buf[cnt * RW + e] = T(0);
      __syncthreads();
    }
    for (int q = warp; q * 4 < cnt; q += GR_WARPS) {
      const T* rows = buf + (q * 4) * RW;
      // k index of the fragment -> (term, problem): k = tig: term tig & 1, problem tig >> 1; k = tig + 4: problem 2 + (tig >> 1)
      const int s = tig & 1, p0 = tig >> 1, p1 = 2 + (tig >> 1);
      int ai = 0;
#pragma unroll
      for (int bi = 0; bi < 5; bi++) {
        if (bi >= nblk) break;
        const bool mrows = bi == 4, mcols = bi == 1 || bi == 3 || bi == 4;
        const int MTB = mrows ? MTM : MTN, CTB = mcols ? M8 : N8;
        const int lo = blk[bi].L[s], ro = blk[bi].R[s];  // < 0: this lane's term is absent (zero operand)
        const int nrows = blk[bi].rows;
#pragma unroll
        for (int mt = 0; mt < (MTN > MTM ? MTN : MTM); mt++) {
          if (mt >= MTB) break;
          uint32_t ah[4], al[4];
          {
            const int r0 = mt * 16 + g, r1 = r0 + 8;
            const float x0 = (lo >= 0 && r0 < nrows) ? rows[p0 * RW + lo + r0] : 0.f;
            const float x1 = (lo >= 0 && r1 < nrows) ? rows[p0 * RW + lo + r1] : 0.f;
            const float x2 = (lo >= 0 && r0 < nrows) ? rows[p1 * RW + lo + r0] : 0.f;
            const float x3 = (lo >= 0 && r1 < nrows) ? rows[p1 * RW + lo + r1] : 0.f;
            split_tf32(x0, ah[0], al[0]);
            split_tf32(x1, ah[1], al[1]);
            split_tf32(x2, ah[2], al[2]);
            split_tf32(x3, ah[3], al[3]);
          }
#pragma unroll
          for (int ct = 0; ct < (N8 > M8 ? N8 : M8); ct++) {
            if (ct >= CTB) break;
            uint32_t bh[2], bl[2];
            const float y0 = ro >= 0 ? rows[p0 * RW + ro + ct * 8 + g] : 0.f;
            const float y1 = ro >= 0 ? rows[p1 * RW + ro + ct * 8 + g] : 0.f;
            split_tf32(y0, bh[0], bl[0]);
            split_tf32(y1, bh[1], bl[1]);
            mma3(acc[ai], ah, al, bh, bl);
            ai++;
          }
        }
      }
#pragma unroll
      for (int i = 0; i < 8; i++) {
        const int e = lane + 32 * i;
        if (e < RW) vs[i] += (rows[e] + rows[RW + e]) + (rows[2 * RW + e] + rows[3 * RW + e]);
      }
    }
    __syncthreads();
  }
#pragma unroll
  for (int i = 0; i < NTILE; i++)
#pragma unroll
    for (int e = 0; e < 4; e++) red[(warp * NACC + 4 * i + e) * 32 + lane] = acc[i][e];
#pragma unroll
  for (int i = 0; i < 8; i++)
    if (lane + 32 * i < RW) vsum[warp * RW + lane + 32 * i] = vs[i];
  __syncthreads();
  T* prow = a.part + (long long)chunk * gr_part_elems(Th, n, m) + (long long)t * gr_per_step(n, m);
  if (warp == 0) {
    int ai = 0;
    for (int bi = 0; bi < nblk; bi++) {
      const bool mrows = bi == 4, mcols = bi == 1 || bi == 3 || bi == 4;
      const int MTB = mrows ? MTM : MTN, CTB = mcols ? M8 : N8;
      for (int mt = 0; mt < MTB; mt++)
        for (int ct = 0; ct < CTB; ct++) {
          for (int e = 0; e < 4; e++) {
            const int row = mt * 16 + g + (e >= 2 ? 8 : 0), col = ct * 8 + 2 * tig + (e & 1);
            T v = 0.f;
            for (int w = 0; w < GR_WARPS; w++) v += red[(w * NACC + 4 * ai + e) * 32 + lane];
            if (row < blk[bi].rows && col < blk[bi].cols) prow[blk[bi].out + row * blk[bi].ld + col] = v;
          }
          ai++;
        }
    }
  }
  {
    const int n8 = gr_pad8(n);
    for (int e = tid; e < RW; e += NTH) {
      T v = 0.f;
      for (int w = 0; w < GR_WARPS; w++) v += vsum[w * RW + e];
      if (t == Th) {
        if (e < n) prow[n * n + e] = v;
      } else {
        const int o3 = 2 * n * (n + m) + m * m;
        if (e >= n8 && e < n8 + n) prow[o3 + (e - n8)] = v;
        else if (e >= 2 * n8 && e < 2 * n8 + n) prow[o3 + n + (e - 2 * n8)] = v;
        else if (e >= 4 * n8 && e < 4 * n8 + m) prow[o3 + 2 * n + (e - 4 * n8)] = v;
      }
    }
  }
}

}  // namespace idoc
