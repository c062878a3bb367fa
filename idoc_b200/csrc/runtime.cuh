// Glue shared by the per-shape translation units and the C-ABI file: launch accounting / optional
// per-kernel event timing, error reporting, launch-geometry selection, tuning overrides.
#pragma once
#include <cuda_runtime.h>
#include <cstdlib>

#include "common.cuh"

namespace idoc {

enum { K_RICCATI = 0, K_ROLLOUT = 1, K_FIXUP = 2, K_ADJ_BWD = 3, K_ADJ_FWD = 4, K_KKT = 5, K_GEN = 6,
       K_TILQR_FWD = 7, K_TILQR_VJP = 8, K_ILQR = 9, K_MISC = 10, K_ILQR_FUSED = 11, K_GRAD_REDUCE = 12 };

// implemented in idoc_b200.cu
void* prof_launch_begin(int id, cudaStream_t s);
void prof_launch_end(void* tok, cudaStream_t s);
int report_cuda_error(cudaError_t e, const char* where);

struct ProfScope {
  cudaStream_t s;
  void* tok;
  ProfScope(int id, cudaStream_t s_) : s(s_), tok(prof_launch_begin(id, s_)) {}
  ~ProfScope() {
    if (tok) prof_launch_end(tok, s);
  }
};

#define IDOC_CK(call)                                             \
  do {                                                            \
    cudaError_t e_ = (call);                                      \
    if (e_ != cudaSuccess) return ::idoc::report_cuda_error(e_, #call); \
  } while (0)

constexpr int MAX_SMEM = 227 * 1024;

// warps per CTA that maximise resident warps per SM for a kernel needing `warp_bytes` of shared memory per
// warp (1 KB of shared memory is reserved per resident CTA; at most 32 CTAs and 64 warps per SM)
constexpr int resident_warps(int warp_bytes, int w) {
  int ctas = (228 * 1024) / (w * warp_bytes + 1024);
  ctas = ctas > 32 ? 32 : ctas;
  int warps = ctas * w;
  return warps > 64 ? 64 : warps;
}
constexpr int pick_warps(int warp_bytes) {
  int best = 1;
  for (int w = 1; w <= 4; w++)
    if (w * warp_bytes <= MAX_SMEM && resident_warps(warp_bytes, w) >= resident_warps(warp_bytes, best)) best = w;
  return best;
}

inline int tune_int(const char* name, int dflt) {
  const char* v = std::getenv(name);
  return v ? std::atoi(v) : dflt;
}

}  // namespace idoc
