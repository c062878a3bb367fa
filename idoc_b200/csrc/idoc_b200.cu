// C ABI of idoc_b200 (see include/idoc_b200.h): argument validation, (dtype, n, m) dispatch to the
// compiled kernel instantiations, launch configuration, and the small CUDA-runtime helper surface.
#include "../../include/idoc_b200.h"

#include <atomic>
#include <cstdio>
#include <cstring>

#include <mutex>
#include <vector>

#include "aux_kernels.cuh"
#include "blqr.cuh"
#include "generic.cuh"
#include "grad_reduce.cuh"
#include "grad_reduce_mma.cuh"
#include "lqr_fwd.cuh"
#include "lqr_vjp.cuh"
#include "runtime.cuh"

using namespace idoc;

static thread_local char g_err[512] = "";
static std::atomic<long long> g_launches{0};

static int fail(int code, const char* fmt, const char* detail = "") {
  snprintf(g_err, sizeof(g_err), fmt, detail);
  return code;
}
static int cuda_fail(cudaError_t e, const char* where) {
  snprintf(g_err, sizeof(g_err), "%s: %s", where, cudaGetErrorString(e));
  return IDOC_ERR_CUDA;
}
#define CK(call)                                  \
  do {                                            \
    cudaError_t e_ = (call);                      \
    if (e_ != cudaSuccess) return cuda_fail(e_, #call); \
  } while (0)

// ------------------------------------------------------------------------------------------ per-kernel timing
// Optional CUDA-event bracket around every kernel this library launches (bench.py's roofline legs).
struct ProfRec {
  int id;
  cudaEvent_t a, b;
};
static std::mutex g_prof_mu;
static std::vector<ProfRec> g_prof;        // records of the current profiling window
static std::vector<cudaEvent_t> g_evpool;  // recycled events
static std::atomic<bool> g_prof_on{false};

namespace idoc {
void* prof_launch_begin(int id, cudaStream_t s) {
  g_launches++;
  if (!g_prof_on.load(std::memory_order_relaxed)) return nullptr;
  std::lock_guard<std::mutex> lk(g_prof_mu);
  cudaEvent_t ev[2];
  for (auto& e : ev) {
    if (!g_evpool.empty()) {
      e = g_evpool.back();
      g_evpool.pop_back();
    } else {
      cudaEventCreate(&e);
    }
  }
  cudaEventRecord(ev[0], s);
  g_prof.push_back({id, ev[0], ev[1]});
  return (void*)ev[1];
}
void prof_launch_end(void* tok, cudaStream_t s) { cudaEventRecord((cudaEvent_t)tok, s); }
int report_cuda_error(cudaError_t e, const char* where) { return cuda_fail(e, where); }
}  // namespace idoc

static bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }

// ------------------------------------------------------------------------------------------ shapes
// One translation unit per compiled (n, m) (shape_<n>_<m>.cu, generated from shapes.def by the build).
namespace idoc {
#define IDOC_SHAPE(N_, M_, L32, L64, NB32, NB64, TUN)                                       \
  int lqr_fwd_f32_##N_##_##M_(cudaStream_t, const LqrFwdArgs<float>&, bool);               \
  int lqr_fwd_f64_##N_##_##M_(cudaStream_t, const LqrFwdArgs<double>&, bool);              \
  int lqr_vjp_f32_##N_##_##M_(cudaStream_t, const LqrVjpArgs<float>&);                     \
  int lqr_vjp_f64_##N_##_##M_(cudaStream_t, const LqrVjpArgs<double>&);
#include "shapes.def"
#undef IDOC_SHAPE
}  // namespace idoc

static bool tuned_shape(int n, int m) {
  if (tune_int("IDOC_FORCE_GENERIC", 0)) return false;
#define IDOC_SHAPE(N_, M_, L32, L64, NB32, NB64, TUN) \
  if (n == N_ && m == M_) return true;
#include "shapes.def"
#undef IDOC_SHAPE
  return false;
}
static bool generic_shape(int n, int m) { return n >= 1 && m >= 1 && n <= GEN_MAXN && m <= GEN_MAXD; }

// generic (runtime-shape) launches ------------------------------------------------------------------------
// A problem whose working set fits a CTA's shared memory runs there with one warp; larger ones run with 256 threads
// out of a per-problem scratch region placed after the residual slots in the caller's workspace.
constexpr size_t GEN_SMEM_LIMIT = 160 * 1024;
template <typename T>
static size_t gen_fwd_scratch_elems(int n, int m) {
  const GeoRT G(n, m, (int)sizeof(T));
  const size_t e = (size_t)cmax(gen_riccati_elems(n, m, G.WS), gen_rollout_elems(n, m, G.WS));
  return e * sizeof(T) > GEN_SMEM_LIMIT ? e : 0;
}
template <typename T>
static size_t gen_vjp_scratch_elems(int n, int m) {
  const GeoRT G(n, m, (int)sizeof(T));
  const size_t e = (size_t)gen_adj_elems(n, m, G.WS, G.WS2);
  return e * sizeof(T) > GEN_SMEM_LIMIT ? e : 0;
}
static size_t status_bytes(int64_t B) { return (((size_t)B * 4 + 15) / 16) * 16; }

// ---- batch-summed gradients (reduce_batch = 1): scratch appended to the VJP workspace -------------------------------
// [uq B*T*n | ur B*T*m | ud B*T*n | uqf B*n | partials NC * (T*PER_STEP + n*n + n)], every region 16-byte aligned.
// NC batch chunks: enough CTAs to fill the GPU ((T+1) * NC of them), at least 256 problems per chunk.
static int reduce_chunks(int64_t B) {
  const int64_t c = B / 256;
  return (int)(c < 1 ? 1 : (c > 64 ? 64 : c));
}
struct ReduceLayout {
  size_t uq, ur, ud, uqf, part, total;  // element offsets from the start of the scratch, total elements
  int NC;
};
static ReduceLayout reduce_layout(int64_t B, int T, int n, int m) {
  auto up = [](size_t e) { return (e + 3) / 4 * 4; };  // 16 bytes in fp32, 32 in fp64
  ReduceLayout r;
  r.NC = reduce_chunks(B);
  r.uq = 0;
  r.ur = r.uq + up((size_t)B * T * n);
  r.ud = r.ur + up((size_t)B * T * m);
  r.uqf = r.ud + up((size_t)B * T * n);
  r.part = r.uqf + up((size_t)B * n);
  r.total = r.part + up((size_t)r.NC * (size_t)gr_part_elems(T, n, m));
  return r;
}
// threads per problem of the generic kernels: the matrix loops have O(n (n + m)) independent outputs per step
static int gen_threads(int n, bool big) {
  const int t = tune_int("IDOC_GEN_THREADS", 0);
  if (t > 0) return t;
  (void)n;  // measured (tools/bench_configs.py): with the GPU full, one warp per problem is the fastest for n <= 32
  return big ? 256 : 32;
}

// mid-size warp-per-problem kernels (csrc/mid.cuh): one translation unit per (n, m) of mid_shapes.def
namespace idoc {
#define IDOC_MID_SHAPE(N_, M_, LP_)                                             \
  int mid_sweep_f32_##N_##_##M_(int which, cudaStream_t s, const void* args);   \
  int mid_sweep_f64_##N_##_##M_(int which, cudaStream_t s, const void* args);
#include "mid_shapes.def"
#undef IDOC_MID_SHAPE
}  // namespace idoc
typedef int (*MidSweep)(int, cudaStream_t, const void*);
template <typename T>
static MidSweep mid_sweep(int n, int m) {
  if (tune_int("IDOC_NO_MID", 0)) return nullptr;
#define IDOC_MID_SHAPE(N_, M_, LP_) \
  if (n == N_ && m == M_) return sizeof(T) == 4 ? mid_sweep_f32_##N_##_##M_ : mid_sweep_f64_##N_##_##M_;
#include "mid_shapes.def"
#undef IDOC_MID_SHAPE
  return nullptr;
}

template <typename T>
static int gen_fwd(cudaStream_t s, int n, int m, const LqrFwdArgs<T>& a, bool rollout, int ti) {
  GenFwdArgs<T> g{a, n, m, ti};
  const GeoRT G(n, m, (int)sizeof(T));
  const size_t big = gen_fwd_scratch_elems<T>(n, m);
  T* scratch = big ? reinterpret_cast<T*>(reinterpret_cast<char*>(a.wstatus) + status_bytes(a.nb)) : nullptr;
  const int threads = gen_threads(n, big != 0);
  const MidSweep mid = mid_sweep<T>(n, m);
  if (mid) {
    const int rc = mid(0, s, &g);
    if (rc) return rc;
  } else {
    const size_t smem = big ? 0 : (size_t)gen_riccati_elems(n, m, G.WS) * sizeof(T);
    auto kern = gen_riccati_kernel<T>;
    CK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)GEN_SMEM_LIMIT));
    ProfScope ps(K_RICCATI, s);
    kern<<<(unsigned)a.nb, threads, smem, s>>>(g, scratch);
  }
  CK(cudaGetLastError());
  if (rollout && mid) return mid(1, s, &g);
  if (rollout) {
    const size_t smem = big ? 0 : (size_t)gen_rollout_elems(n, m, G.WS) * sizeof(T);
    auto kern = gen_rollout_kernel<T>;
    CK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)GEN_SMEM_LIMIT));
    ProfScope ps(K_ROLLOUT, s);
    kern<<<(unsigned)a.nb, threads, smem, s>>>(g, scratch);
  }
  CK(cudaGetLastError());
  return IDOC_OK;
}
template <typename T>
static int gen_vjp(cudaStream_t s, int n, int m, const LqrVjpArgs<T>& a, int ti) {
  GenVjpArgs<T> g{a, n, m, ti};
  const GeoRT G(n, m, (int)sizeof(T));
  const size_t big = gen_vjp_scratch_elems<T>(n, m);
  T* scratch = big ? a.ws2 + (size_t)a.nb * a.T_ * G.WS2 : nullptr;
  const int threads = gen_threads(n, big != 0);
  if (const MidSweep mid = mid_sweep<T>(n, m)) {
    const int rc = mid(2, s, &g);
    if (rc) return rc;
    return mid(3, s, &g);
  }
  const size_t smem = big ? 0 : (size_t)gen_adj_elems(n, m, G.WS, G.WS2) * sizeof(T);
  {
    auto kern = gen_adj_bwd_kernel<T>;
    CK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)GEN_SMEM_LIMIT));
    ProfScope ps(K_ADJ_BWD, s);
    kern<<<(unsigned)a.nb, threads, smem, s>>>(g, scratch);
  }
  CK(cudaGetLastError());
  {
    auto kern = gen_adj_fwd_kernel<T>;
    CK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)GEN_SMEM_LIMIT));
    ProfScope ps(K_ADJ_FWD, s);
    kern<<<(unsigned)a.nb, threads, smem, s>>>(g, scratch);
  }
  CK(cudaGetLastError());
  return IDOC_OK;
}

template <typename T>
static size_t ws_elems(int n, int m, bool vjp) {
  if (!generic_shape(n, m)) return 0;
  const GeoRT G(n, m, (int)sizeof(T));
  return vjp ? G.WS2 : G.WS;
}
// bytes of the regular VJP workspace: ws2 slots + per-problem scratch of the large generic shapes (16-byte multiple)
template <typename T>
static size_t vjp_base_bytes(int64_t B, int T_, int n, int m) {
  const size_t e = ws_elems<T>(n, m, true);
  const size_t sc = tuned_shape(n, m) ? 0 : gen_vjp_scratch_elems<T>(n, m);
  const size_t b = (e * (size_t)B * (size_t)T_ + sc * (size_t)B) * sizeof(T);
  return (b + 15) / 16 * 16;
}
static int check_common(int dtype, int64_t B, int T, int n, int m) {
  if (dtype != IDOC_F32 && dtype != IDOC_F64) return fail(IDOC_ERR_INVALID, "dtype must be IDOC_F32 or IDOC_F64");
  if (B <= 0 || T <= 0 || n <= 0 || m <= 0) return fail(IDOC_ERR_INVALID, "B, T, n, m must be positive");
  if ((double)B * (double)T >= 2147483647.0) return fail(IDOC_ERR_INVALID, "B*T must be < 2^31");
  if (!idoc_lqr_supported(dtype, n, m)) return fail(IDOC_ERR_UNSUPPORTED, "unsupported shape: need 1 <= n <= 128, 1 <= m <= 32");
  return IDOC_OK;
}

static int fwd_dispatch(cudaStream_t s, int n, int m, const LqrFwdArgs<float>& a, bool rollout) {
#define IDOC_SHAPE(N_, M_, L32, L64, NB32, NB64, TUN) \
  if (n == N_ && m == M_) return lqr_fwd_f32_##N_##_##M_(s, a, rollout);
#include "shapes.def"
#undef IDOC_SHAPE
  return fail(IDOC_ERR_UNSUPPORTED, "no kernel compiled for this (n, m)");
}
static int fwd_dispatch(cudaStream_t s, int n, int m, const LqrFwdArgs<double>& a, bool rollout) {
#define IDOC_SHAPE(N_, M_, L32, L64, NB32, NB64, TUN) \
  if (n == N_ && m == M_) return lqr_fwd_f64_##N_##_##M_(s, a, rollout);
#include "shapes.def"
#undef IDOC_SHAPE
  return fail(IDOC_ERR_UNSUPPORTED, "no kernel compiled for this (n, m)");
}
static int vjp_dispatch(cudaStream_t s, int n, int m, const LqrVjpArgs<float>& a) {
#define IDOC_SHAPE(N_, M_, L32, L64, NB32, NB64, TUN) \
  if (n == N_ && m == M_) return lqr_vjp_f32_##N_##_##M_(s, a);
#include "shapes.def"
#undef IDOC_SHAPE
  return fail(IDOC_ERR_UNSUPPORTED, "no kernel compiled for this (n, m)");
}
static int vjp_dispatch(cudaStream_t s, int n, int m, const LqrVjpArgs<double>& a) {
#define IDOC_SHAPE(N_, M_, L32, L64, NB32, NB64, TUN) \
  if (n == N_ && m == M_) return lqr_vjp_f64_##N_##_##M_(s, a);
#include "shapes.def"
#undef IDOC_SHAPE
  return fail(IDOC_ERR_UNSUPPORTED, "no kernel compiled for this (n, m)");
}

template <typename T>
static int fwd_typed(void* stream, int64_t B, int T_, int n, int m, const void* Q, const void* q, const void* Qf,
                     const void* qf, const void* M, const void* R, const void* r, const void* A, const void* Bm,
                     const void* d, const void* x0, void* X, void* U, void* Nu, void* K, void* k, void* dCdc,
                     int32_t* status, void* ws, bool rollout) {
  LqrFwdArgs<T> a;
  a.Q = (const T*)Q; a.q = (const T*)q; a.Qf = (const T*)Qf; a.qf = (const T*)qf; a.Mx = (const T*)M;
  a.R = (const T*)R; a.r = (const T*)r; a.A = (const T*)A; a.B = (const T*)Bm; a.d = (const T*)d;
  a.x0 = (const T*)x0; a.X = (T*)X; a.U = (T*)U; a.Nu = (T*)Nu; a.K = (T*)K; a.k = (T*)k;
  a.ws = (T*)ws; a.dCdc = (T*)dCdc; a.status = status; a.nb = B; a.T_ = T_;
  a.wstatus = reinterpret_cast<int*>(reinterpret_cast<char*>(ws) + ws_elems<T>(n, m, false) * (size_t)B * (size_t)T_ * sizeof(T));
  {
    int rc = tuned_shape(n, m) ? fwd_dispatch((cudaStream_t)stream, n, m, a, rollout)
                               : gen_fwd<T>((cudaStream_t)stream, n, m, a, rollout, 0);
    if (rc) return rc;
  }
  if (rollout) {
    FixupArgs<T> f;
    f.Q = a.Q; f.q = a.q; f.Qf = a.Qf; f.qf = a.qf; f.Mx = a.Mx; f.A = a.A; f.X = a.X; f.U = a.U; f.Unf = a.U; f.Nu = a.Nu;
    f.status = a.wstatus; f.ustatus = a.status; f.nb = a.nb; f.T_ = a.T_; f.n = n; f.m = m;
    cudaStream_t s = (cudaStream_t)stream;
    {
      ProfScope ps(K_FIXUP, s);
      lqr_costate_fixup_kernel<T><<<(unsigned)((a.nb + 255) / 256), 256, 0, s>>>(f);
    }
    CK(cudaGetLastError());
  }
  return IDOC_OK;
}

// returns 0 = launched, 1 = no instantiation for this shape (caller falls back to the FFMA kernel), < 0 = error
template <int N8, int M8, int ZS>
static int launch_dmma(cudaStream_t s, const GradReduceArgs<double>& g, unsigned NC) {
  constexpr int RT = (N8 + ZS - 1) / ZS, NACC = 2 * (2 * RT * N8 + 2 * RT * M8 + M8 * M8);
  const size_t smem = grm_smem_bytes(g.n, g.m, sizeof(double), NACC);
  auto kern = grad_reduce_dmma_kernel<N8, M8, ZS>;
  CK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  ProfScope ps(K_GRAD_REDUCE, s);
  kern<<<dim3((unsigned)(g.T_ + 1), NC, ZS), GR_WARPS * 32, smem, s>>>(g);
  CK(cudaGetLastError());
  return 0;
}
template <int N8, int M8>
static int launch_tf32(cudaStream_t s, const GradReduceArgs<float>& g, unsigned NC) {
  constexpr int MTN = (N8 + 1) / 2, MTM = (M8 + 1) / 2, NACC = 4 * (2 * MTN * N8 + 2 * MTN * M8 + MTM * M8);
  const size_t smem = grm_smem_bytes(g.n, g.m, sizeof(float), NACC);
  auto kern = grad_reduce_tf32_kernel<N8, M8>;
  CK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  ProfScope ps(K_GRAD_REDUCE, s);
  kern<<<dim3((unsigned)(g.T_ + 1), NC, 1), GR_WARPS * 32, smem, s>>>(g);
  CK(cudaGetLastError());
  return 0;
}
static int launch_grad_reduce_mma(cudaStream_t s, const GradReduceArgs<double>& g, unsigned NC) {
  const int n8 = (g.n + 7) / 8, m8 = (g.m + 7) / 8;
  if (n8 == 1 && m8 == 1) return launch_dmma<1, 1, 1>(s, g, NC);
  if (n8 == 2 && m8 == 1) return launch_dmma<2, 1, 1>(s, g, NC);
  if (n8 == 2 && m8 == 2) return launch_dmma<2, 2, 1>(s, g, NC);
  if (n8 == 3 && m8 == 1) return launch_dmma<3, 1, 2>(s, g, NC);
  if (n8 == 4 && m8 == 1) return launch_dmma<4, 1, 2>(s, g, NC);
  if (n8 == 4 && m8 == 2) return launch_dmma<4, 2, 2>(s, g, NC);
  return 1;
}
static int launch_grad_reduce_mma(cudaStream_t s, const GradReduceArgs<float>& g, unsigned NC) {
  const int n8 = (g.n + 7) / 8, m8 = (g.m + 7) / 8;
  if (n8 == 1 && m8 == 1) return launch_tf32<1, 1>(s, g, NC);
  if (n8 == 2 && m8 == 1) return launch_tf32<2, 1>(s, g, NC);
  if (n8 == 2 && m8 == 2) return launch_tf32<2, 2>(s, g, NC);
  if (n8 == 3 && m8 == 1) return launch_tf32<3, 1>(s, g, NC);
  if (n8 == 4 && m8 == 1) return launch_tf32<4, 1>(s, g, NC);
  if (n8 == 4 && m8 == 2) return launch_tf32<4, 2>(s, g, NC);
  return 1;
}

template <typename T>
static int vjp_typed(void* stream, int64_t B, int T_, int n, int m, const void* M, const void* A, const void* Bm,
                     const void* x0, const void* X, const void* U, const void* Nu, const void* ws, const void* Xb,
                     const void* Ub, const void* Nub, void* gQ, void* gq, void* gQf, void* gqf, void* gM, void* gR,
                     void* gr, void* gA, void* gB, void* gd, void* gx0, void* ws2, int reduce_batch) {
  LqrVjpArgs<T> a;
  a.Mx = (const T*)M; a.A = (const T*)A; a.B = (const T*)Bm; a.x0 = (const T*)x0;
  a.X = (const T*)X; a.U = (const T*)U; a.Nu = (const T*)Nu; a.ws = (const T*)ws;
  a.Xb = (const T*)Xb; a.Ub = (const T*)Ub; a.Nub = (const T*)Nub;
  a.gQ = (T*)gQ; a.gq = (T*)gq; a.gQf = (T*)gQf; a.gqf = (T*)gqf; a.gM = (T*)gM; a.gR = (T*)gR; a.gr = (T*)gr;
  a.gA = (T*)gA; a.gB = (T*)gB; a.gd = (T*)gd; a.gx0 = (T*)gx0; a.ws2 = (T*)ws2;
  a.nb = B; a.T_ = T_; a.reduce_batch = reduce_batch;
  cudaStream_t s = (cudaStream_t)stream;
  if (!reduce_batch) {
    if (!tuned_shape(n, m)) return gen_vjp<T>(s, n, m, a, 0);
    return vjp_dispatch(s, n, m, a);
  }
  // Batch-summed gradients: the sweeps run in vectors-only mode (they store sux, uU, uNu, uX[T-1] per problem into the
  // scratch behind the regular VJP workspace), then the two kernels of grad_reduce.cuh form the sums over the batch.
  const ReduceLayout rl = reduce_layout(B, T_, n, m);
  T* scratch = reinterpret_cast<T*>(reinterpret_cast<char*>(ws2) + vjp_base_bytes<T>(B, T_, n, m));
  LqrVjpArgs<T> v = a;
  v.gq = scratch + rl.uq; v.gr = scratch + rl.ur; v.gd = scratch + rl.ud; v.gqf = scratch + rl.uqf;
  {
    const int rc = tuned_shape(n, m) ? vjp_dispatch(s, n, m, v) : gen_vjp<T>(s, n, m, v, 0);
    if (rc) return rc;
  }
  GradReduceArgs<T> g;
  g.X = a.X; g.U = a.U; g.Nu = a.Nu; g.x0 = a.x0;
  g.uq = v.gq; g.ur = v.gr; g.ud = v.gd; g.uqf = v.gqf; g.part = scratch + rl.part;
  g.gQ = a.gQ; g.gq = a.gq; g.gQf = a.gQf; g.gqf = a.gqf; g.gM = a.gM; g.gR = a.gR; g.gr = a.gr; g.gA = a.gA; g.gB = a.gB; g.gd = a.gd;
  g.nb = B; g.T_ = T_; g.n = n; g.m = m; g.NC = rl.NC; g.stage = gr_stage_problems(n, m, sizeof(T));
  const int slots = 32 * GR_MAXT;
  const int groups = (cmax(gr_tasks_step(n, m), gr_tasks_final(n)) + slots - 1) / slots;
  const dim3 grid((unsigned)(T_ + 1), (unsigned)rl.NC, (unsigned)groups);
  // tensor-core contraction (grad_reduce_mma.cuh) for n in {8, 16, 24, 32}, m in {4, 8, 16}; IDOC_GR_MMA=0 keeps the FFMA kernel
  const bool mma_shape = n % 8 == 0 && n <= 32 && (m == 4 || m == 8 || m == 16) && (n + 7) / 8 >= (m + 7) / 8;
  // measured (profiles/r2_grad_reduce_mma_ab.jsonl): fp64 DMMA wins at every shape, 3xTF32 from n = 16 on (at n = 8 half of
  // every 16-row MMA tile is padding)
  if (mma_shape && tune_int("IDOC_GR_MMA", (sizeof(T) == 8 || n >= 16) ? 1 : 0)) {
    const int rc = launch_grad_reduce_mma(s, g, (unsigned)rl.NC);
    if (rc != 1) {
      if (rc) return rc;
      ProfScope ps(K_GRAD_REDUCE, s);
      const long long pe = gr_part_elems(T_, n, m);
      grad_reduce_final_kernel<T><<<(unsigned)((pe + 255) / 256), 256, 0, s>>>(g);
      CK(cudaGetLastError());
      return IDOC_OK;
    }
  }
  {
    const size_t smem = gr_partial_smem(n, m, sizeof(T));
    const bool vec = n % 4 == 0 && m % 4 == 0;
    auto kern = vec ? grad_reduce_partial_kernel<T, true> : grad_reduce_partial_kernel<T, false>;
    CK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    ProfScope ps(K_GRAD_REDUCE, s);
    kern<<<grid, GR_WARPS * 32, smem, s>>>(g);
  }
  CK(cudaGetLastError());
  {
    ProfScope ps(K_GRAD_REDUCE, s);
    const long long pe = gr_part_elems(T_, n, m);
    grad_reduce_final_kernel<T><<<(unsigned)((pe + 255) / 256), 256, 0, s>>>(g);
  }
  CK(cudaGetLastError());
  return IDOC_OK;
}

template <typename T>
static int kkt_typed(cudaStream_t s, int64_t B, int T_, int n, int m, const void* Q, const void* q, const void* Qf,
                     const void* qf, const void* M, const void* R, const void* r, const void* A, const void* Bm,
                     const void* d, const void* x0, const void* X, const void* U, const void* Nu, void* rX, void* rU,
                     void* rNu) {
  KktArgs<T> a;
  a.Q = (const T*)Q; a.q = (const T*)q; a.Qf = (const T*)Qf; a.qf = (const T*)qf; a.Mx = (const T*)M;
  a.R = (const T*)R; a.r = (const T*)r; a.A = (const T*)A; a.B = (const T*)Bm; a.d = (const T*)d; a.x0 = (const T*)x0;
  a.X = (const T*)X; a.U = (const T*)U; a.Nu = (const T*)Nu; a.rX = (T*)rX; a.rU = (T*)rU; a.rNu = (T*)rNu;
  a.nb = B; a.T_ = T_; a.n = n; a.m = m;
  const long long tot = (long long)B * T_;
  {
    ProfScope ps(K_KKT, s);
    lqr_kkt_kernel<T><<<(unsigned)((tot + 127) / 128), 128, 0, s>>>(a);
  }
  CK(cudaGetLastError());
  return IDOC_OK;
}

template <typename T>
static int gen_typed(cudaStream_t s, int64_t B, int T_, int n, int m, uint64_t seed, void* Q, void* q, void* Qf,
                     void* qf, void* M, void* R, void* r, void* A, void* Bm, void* d, void* x0) {
  GenArgs<T> a;
  a.Q = (T*)Q; a.q = (T*)q; a.Qf = (T*)Qf; a.qf = (T*)qf; a.Mx = (T*)M; a.R = (T*)R; a.r = (T*)r;
  a.A = (T*)A; a.B = (T*)Bm; a.d = (T*)d; a.x0 = (T*)x0; a.nb = B; a.T_ = T_; a.n = n; a.m = m; a.seed = seed;
  const long long tot = (long long)B * (T_ + 1);
  {
    ProfScope ps(K_GEN, s);
    lqr_generate_kernel<T><<<(unsigned)((tot + 63) / 64), 64, 0, s>>>(a);
  }
  CK(cudaGetLastError());
  return IDOC_OK;
}

extern "C" {


const char* idoc_version(void) { return "idoc_b200 0.1 (sm_100a)"; }
const char* idoc_last_error(void) { return g_err; }
int64_t idoc_launch_count(void) { return (int64_t)g_launches.load(); }

int idoc_profile_begin(void) {
  std::lock_guard<std::mutex> lk(g_prof_mu);
  for (auto& r : g_prof) {
    g_evpool.push_back(r.a);
    g_evpool.push_back(r.b);
  }
  g_prof.clear();
  g_prof_on = true;
  return IDOC_OK;
}
int idoc_profile_end(int max_records, int* ids, float* ms, int* count) {
  g_prof_on = false;
  std::lock_guard<std::mutex> lk(g_prof_mu);
  int n = 0;
  for (auto& r : g_prof) {
    if (n >= max_records) break;
    CK(cudaEventSynchronize(r.b));
    float t = 0.f;
    CK(cudaEventElapsedTime(&t, r.a, r.b));
    ids[n] = r.id;
    ms[n] = t;
    n++;
  }
  *count = n;
  for (auto& r : g_prof) {
    g_evpool.push_back(r.a);
    g_evpool.push_back(r.b);
  }
  g_prof.clear();
  return IDOC_OK;
}

int idoc_lqr_supported(int dtype, int n, int m) {
  if (dtype != IDOC_F32 && dtype != IDOC_F64) return 0;
  if (tuned_shape(n, m)) return 1;
  return generic_shape(n, m) ? 2 : 0;
}

size_t idoc_lqr_ws_bytes(int dtype, int64_t B, int T, int n, int m) {
  const size_t e = dtype == IDOC_F32 ? ws_elems<float>(n, m, false) : ws_elems<double>(n, m, false);
  if (e == 0) return 0;
  // residual slots + one status word per problem (16-byte aligned) + per-problem scratch of the large generic shapes
  const size_t sc = tuned_shape(n, m) ? 0 : (dtype == IDOC_F32 ? gen_fwd_scratch_elems<float>(n, m) * 4 : gen_fwd_scratch_elems<double>(n, m) * 8);
  return e * (size_t)B * (size_t)T * (dtype == IDOC_F32 ? 4 : 8) + status_bytes(B) + sc * (size_t)B;
}
size_t idoc_lqr_vjp_ws_bytes(int dtype, int64_t B, int T, int n, int m) {
  if (!generic_shape(n, m)) return 0;
  return dtype == IDOC_F32 ? vjp_base_bytes<float>(B, T, n, m) : vjp_base_bytes<double>(B, T, n, m);
}
size_t idoc_lqr_vjp_reduce_ws_bytes(int dtype, int64_t B, int T, int n, int m) {
  if (!generic_shape(n, m)) return 0;
  return idoc_lqr_vjp_ws_bytes(dtype, B, T, n, m) + reduce_layout(B, T, n, m).total * (dtype == IDOC_F32 ? 4 : 8);
}

int idoc_lqr_fwd(void* stream, int dtype, int64_t B, int T, int n, int m, const void* Q, const void* q,
                 const void* Qf, const void* qf, const void* M, const void* R, const void* r, const void* A,
                 const void* Bm, const void* d, const void* x0, void* X, void* U, void* Nu, void* K, void* k,
                 void* dCdc, int32_t* status, void* ws, size_t ws_bytes) {
  int rc = check_common(dtype, B, T, n, m);
  if (rc) return rc;
  const void* req[] = {Q, q, Qf, qf, M, R, r, A, Bm, d, x0, X, U, Nu, ws};
  for (const void* p : req) {
    if (p == nullptr) return fail(IDOC_ERR_INVALID, "null pointer argument");
    if (!aligned16(p)) return fail(IDOC_ERR_INVALID, "base pointers must be 16-byte aligned");
  }
  if ((K == nullptr) != (k == nullptr)) return fail(IDOC_ERR_INVALID, "K and k must both be given or both be NULL");
  if (ws_bytes < idoc_lqr_ws_bytes(dtype, B, T, n, m)) return fail(IDOC_ERR_WORKSPACE, "ws too small");
  if (dtype == IDOC_F32)
    return fwd_typed<float>(stream, B, T, n, m, Q, q, Qf, qf, M, R, r, A, Bm, d, x0, X, U, Nu, K, k, dCdc, status, ws, true);
  return fwd_typed<double>(stream, B, T, n, m, Q, q, Qf, qf, M, R, r, A, Bm, d, x0, X, U, Nu, K, k, dCdc, status, ws, true);
}

int idoc_lqr_backward(void* stream, int dtype, int64_t B, int T, int n, int m, const void* Q, const void* q,
                      const void* Qf, const void* qf, const void* M, const void* R, const void* r, const void* A,
                      const void* Bm, const void* d, void* K, void* k, void* dCdc, int32_t* status, void* ws,
                      size_t ws_bytes) {
  int rc = check_common(dtype, B, T, n, m);
  if (rc) return rc;
  const void* req[] = {Q, q, Qf, qf, M, R, r, A, Bm, d, ws};
  for (const void* p : req) {
    if (p == nullptr) return fail(IDOC_ERR_INVALID, "null pointer argument");
    if (!aligned16(p)) return fail(IDOC_ERR_INVALID, "base pointers must be 16-byte aligned");
  }
  if ((K == nullptr) != (k == nullptr)) return fail(IDOC_ERR_INVALID, "K and k must both be given or both be NULL");
  if (ws_bytes < idoc_lqr_ws_bytes(dtype, B, T, n, m)) return fail(IDOC_ERR_WORKSPACE, "ws too small");
  if (dtype == IDOC_F32)
    return fwd_typed<float>(stream, B, T, n, m, Q, q, Qf, qf, M, R, r, A, Bm, d, nullptr, nullptr, nullptr, nullptr, K, k, dCdc, status, ws, false);
  return fwd_typed<double>(stream, B, T, n, m, Q, q, Qf, qf, M, R, r, A, Bm, d, nullptr, nullptr, nullptr, nullptr, K, k, dCdc, status, ws, false);
}

int idoc_lqr_vjp(void* stream, int dtype, int64_t B, int T, int n, int m, const void* M, const void* A,
                 const void* Bm, const void* x0, const void* X, const void* U, const void* Nu, const void* ws,
                 const void* Xb, const void* Ub, const void* Nub, void* gQ, void* gq, void* gQf, void* gqf, void* gM,
                 void* gR, void* gr, void* gA, void* gB, void* gd, void* gx0, void* ws2, size_t ws2_bytes,
                 int reduce_batch) {
  int rc = check_common(dtype, B, T, n, m);
  if (rc) return rc;
  const void* req[] = {M, A, Bm, x0, X, U, Nu, ws, Xb, Ub, gQ, gq, gQf, gqf, gM, gR, gr, gA, gB, gd, gx0, ws2};
  for (const void* p : req) {
    if (p == nullptr) return fail(IDOC_ERR_INVALID, "null pointer argument");
    if (!aligned16(p)) return fail(IDOC_ERR_INVALID, "base pointers must be 16-byte aligned");
  }
  if (Nub != nullptr && !aligned16(Nub)) return fail(IDOC_ERR_INVALID, "base pointers must be 16-byte aligned");
  if (ws2_bytes < (reduce_batch ? idoc_lqr_vjp_reduce_ws_bytes(dtype, B, T, n, m) : idoc_lqr_vjp_ws_bytes(dtype, B, T, n, m)))
    return fail(IDOC_ERR_WORKSPACE, reduce_batch ? "ws2 too small (reduce_batch = 1 needs idoc_lqr_vjp_reduce_ws_bytes)" : "ws2 too small");
  if (dtype == IDOC_F32)
    return vjp_typed<float>(stream, B, T, n, m, M, A, Bm, x0, X, U, Nu, ws, Xb, Ub, Nub, gQ, gq, gQf, gqf, gM, gR, gr, gA, gB, gd, gx0, ws2, reduce_batch);
  return vjp_typed<double>(stream, B, T, n, m, M, A, Bm, x0, X, U, Nu, ws, Xb, Ub, Nub, gQ, gq, gQf, gqf, gM, gR, gr, gA, gB, gd, gx0, ws2, reduce_batch);
}

int idoc_lqr_kkt(void* stream, int dtype, int64_t B, int T, int n, int m, const void* Q, const void* q,
                 const void* Qf, const void* qf, const void* M, const void* R, const void* r, const void* A,
                 const void* Bm, const void* d, const void* x0, const void* X, const void* U, const void* Nu,
                 void* rX, void* rU, void* rNu) {
  if (dtype != IDOC_F32 && dtype != IDOC_F64) return fail(IDOC_ERR_INVALID, "dtype must be IDOC_F32 or IDOC_F64");
  if (B <= 0 || T <= 0 || n <= 0 || m <= 0) return fail(IDOC_ERR_INVALID, "B, T, n, m must be positive");
  const void* req[] = {Q, q, Qf, qf, M, R, r, A, Bm, d, x0, X, U, Nu, rX, rU, rNu};
  for (const void* p : req)
    if (p == nullptr) return fail(IDOC_ERR_INVALID, "null pointer argument");
  if (dtype == IDOC_F32)
    return kkt_typed<float>((cudaStream_t)stream, B, T, n, m, Q, q, Qf, qf, M, R, r, A, Bm, d, x0, X, U, Nu, rX, rU, rNu);
  return kkt_typed<double>((cudaStream_t)stream, B, T, n, m, Q, q, Qf, qf, M, R, r, A, Bm, d, x0, X, U, Nu, rX, rU, rNu);
}

int idoc_lqr_generate(void* stream, int dtype, int64_t B, int T, int n, int m, uint64_t seed, void* Q, void* q,
                      void* Qf, void* qf, void* M, void* R, void* r, void* A, void* Bm, void* d, void* x0) {
  if (dtype != IDOC_F32 && dtype != IDOC_F64) return fail(IDOC_ERR_INVALID, "dtype must be IDOC_F32 or IDOC_F64");
  if (B <= 0 || T <= 0 || n <= 0 || m <= 0 || n > GEN_MAX || m > GEN_MAX)
    return fail(IDOC_ERR_INVALID, "generator needs 0 < n, m <= 32");
  if (dtype == IDOC_F32) return gen_typed<float>((cudaStream_t)stream, B, T, n, m, seed, Q, q, Qf, qf, M, R, r, A, Bm, d, x0);
  return gen_typed<double>((cudaStream_t)stream, B, T, n, m, seed, Q, q, Qf, qf, M, R, r, A, Bm, d, x0);
}

// ------------------------------------------------------------------------------------------ tilqr
}  // extern "C"
template <typename T>
static int tilqr_fwd_typed(void* stream, int64_t B, int T_, int n, int m, const void* Q, const void* Qf, const void* R,
                           const void* A, const void* Bm, const void* x0, void* X, void* U, void* Nu, void* K,
                           int32_t* status, void* ws) {
  LqrFwdArgs<T> a{};
  a.Q = (const T*)Q; a.Qf = (const T*)Qf; a.R = (const T*)R; a.A = (const T*)A; a.B = (const T*)Bm; a.x0 = (const T*)x0;
  a.X = (T*)X; a.U = (T*)U; a.Nu = (T*)Nu; a.K = (T*)K; a.k = nullptr; a.ws = (T*)ws; a.status = status;
  a.nb = B; a.T_ = T_;
  a.wstatus = reinterpret_cast<int*>(reinterpret_cast<char*>(ws) + ws_elems<T>(n, m, false) * (size_t)B * (size_t)T_ * sizeof(T));
  if (K != nullptr) return fail(IDOC_ERR_INVALID, "tilqr: pass K = NULL (gains live in ws)");
  cudaStream_t s = (cudaStream_t)stream;
  const int rc = gen_fwd<T>(s, n, m, a, true, 1);
  if (rc) return rc;
  FixupArgs<T> f{};  // no eigen-shift in tilqr: only the non-finite report (failed factorisation, NaN inputs)
  f.X = a.X; f.U = a.U; f.Unf = a.U; f.Nu = a.Nu; f.status = a.wstatus; f.ustatus = a.status;
  f.nb = a.nb; f.T_ = a.T_; f.n = n; f.m = m;
  {
    ProfScope ps(K_FIXUP, s);
    lqr_costate_fixup_kernel<T><<<(unsigned)((a.nb + 255) / 256), 256, 0, s>>>(f);
  }
  CK(cudaGetLastError());
  return IDOC_OK;
}
template <typename T>
static int tilqr_vjp_typed(void* stream, int64_t B, int T_, int n, int m, const void* A, const void* Bm, const void* x0,
                           const void* X, const void* U, const void* Nu, const void* ws, const void* Xb, const void* Ub,
                           const void* Nub, void* gQ, void* gQf, void* gR, void* gA, void* gB, void* gx0, void* ws2) {
  LqrVjpArgs<T> a{};
  a.A = (const T*)A; a.B = (const T*)Bm; a.x0 = (const T*)x0; a.X = (const T*)X; a.U = (const T*)U; a.Nu = (const T*)Nu;
  a.ws = (const T*)ws; a.Xb = (const T*)Xb; a.Ub = (const T*)Ub; a.Nub = (const T*)Nub;
  a.gQ = (T*)gQ; a.gQf = (T*)gQf; a.gR = (T*)gR; a.gA = (T*)gA; a.gB = (T*)gB; a.gx0 = (T*)gx0; a.ws2 = (T*)ws2;
  a.nb = B; a.T_ = T_; a.reduce_batch = 0;
  return gen_vjp<T>((cudaStream_t)stream, n, m, a, 1);
}
extern "C" {
int idoc_tilqr_fwd(void* stream, int dtype, int64_t B, int T, int n, int m, const void* Q, const void* Qf, const void* R,
                   const void* A, const void* Bm, const void* x0, void* X, void* U, void* Nu, int32_t* status, void* ws,
                   size_t ws_bytes) {
  int rc = check_common(dtype, B, T, n, m);
  if (rc) return rc;
  const void* req[] = {Q, Qf, R, A, Bm, x0, X, U, Nu, ws};
  for (const void* p : req) {
    if (p == nullptr) return fail(IDOC_ERR_INVALID, "null pointer argument");
    if (!aligned16(p)) return fail(IDOC_ERR_INVALID, "base pointers must be 16-byte aligned");
  }
  if (ws_bytes < idoc_lqr_ws_bytes(dtype, B, T, n, m)) return fail(IDOC_ERR_WORKSPACE, "ws too small");
  if (dtype == IDOC_F32) return tilqr_fwd_typed<float>(stream, B, T, n, m, Q, Qf, R, A, Bm, x0, X, U, Nu, nullptr, status, ws);
  return tilqr_fwd_typed<double>(stream, B, T, n, m, Q, Qf, R, A, Bm, x0, X, U, Nu, nullptr, status, ws);
}
int idoc_tilqr_vjp(void* stream, int dtype, int64_t B, int T, int n, int m, const void* A, const void* Bm, const void* x0,
                   const void* X, const void* U, const void* Nu, const void* ws, const void* Xb, const void* Ub,
                   const void* Nub, void* gQ, void* gQf, void* gR, void* gA, void* gB, void* gx0, void* ws2,
                   size_t ws2_bytes) {
  int rc = check_common(dtype, B, T, n, m);
  if (rc) return rc;
  const void* req[] = {A, Bm, x0, X, U, Nu, ws, Xb, Ub, gQ, gQf, gR, gA, gB, gx0, ws2};
  for (const void* p : req) {
    if (p == nullptr) return fail(IDOC_ERR_INVALID, "null pointer argument");
    if (!aligned16(p)) return fail(IDOC_ERR_INVALID, "base pointers must be 16-byte aligned");
  }
  if (Nub != nullptr && !aligned16(Nub)) return fail(IDOC_ERR_INVALID, "base pointers must be 16-byte aligned");
  if (ws2_bytes < idoc_lqr_vjp_ws_bytes(dtype, B, T, n, m)) return fail(IDOC_ERR_WORKSPACE, "ws2 too small");
  if (dtype == IDOC_F32)
    return tilqr_vjp_typed<float>(stream, B, T, n, m, A, Bm, x0, X, U, Nu, ws, Xb, Ub, Nub, gQ, gQf, gR, gA, gB, gx0, ws2);
  return tilqr_vjp_typed<double>(stream, B, T, n, m, A, Bm, x0, X, U, Nu, ws, Xb, Ub, Nub, gQ, gQf, gR, gA, gB, gx0, ws2);
}

// ------------------------------------------------------------------------------------------ runtime helpers
int idoc_device_count(void) {
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess) return 0;
  return n;
}
int idoc_set_device(int dev) { CK(cudaSetDevice(dev)); return IDOC_OK; }
int idoc_malloc(void** dptr, size_t bytes) { CK(cudaMalloc(dptr, bytes)); return IDOC_OK; }
int idoc_free(void* dptr) { CK(cudaFree(dptr)); return IDOC_OK; }
int idoc_host_alloc(void** hptr, size_t bytes) { CK(cudaHostAlloc(hptr, bytes, cudaHostAllocDefault)); return IDOC_OK; }
int idoc_host_free(void* hptr) { CK(cudaFreeHost(hptr)); return IDOC_OK; }
int idoc_memset(void* dptr, int value, size_t bytes, void* stream) {
  CK(cudaMemsetAsync(dptr, value, bytes, (cudaStream_t)stream));
  return IDOC_OK;
}
int idoc_memcpy_h2d(void* dst, const void* src, size_t bytes, void* stream) {
  if (stream) CK(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyHostToDevice, (cudaStream_t)stream));
  else CK(cudaMemcpy(dst, src, bytes, cudaMemcpyHostToDevice));
  return IDOC_OK;
}
int idoc_memcpy_d2h(void* dst, const void* src, size_t bytes, void* stream) {
  if (stream) CK(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToHost, (cudaStream_t)stream));
  else CK(cudaMemcpy(dst, src, bytes, cudaMemcpyDeviceToHost));
  return IDOC_OK;
}
int idoc_memcpy_d2d(void* dst, const void* src, size_t bytes, void* stream) {
  CK(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToDevice, (cudaStream_t)stream));
  return IDOC_OK;
}
int idoc_stream_create(void** stream) {
  cudaStream_t s;
  CK(cudaStreamCreateWithFlags(&s, cudaStreamNonBlocking));
  *stream = (void*)s;
  return IDOC_OK;
}
int idoc_stream_destroy(void* stream) { CK(cudaStreamDestroy((cudaStream_t)stream)); return IDOC_OK; }
int idoc_stream_sync(void* stream) {
  if (stream) CK(cudaStreamSynchronize((cudaStream_t)stream));
  else CK(cudaDeviceSynchronize());
  return IDOC_OK;
}
int idoc_event_create(void** ev) {
  cudaEvent_t e;
  CK(cudaEventCreate(&e));
  *ev = (void*)e;
  return IDOC_OK;
}
int idoc_event_destroy(void* ev) { CK(cudaEventDestroy((cudaEvent_t)ev)); return IDOC_OK; }
int idoc_event_record(void* ev, void* stream) { CK(cudaEventRecord((cudaEvent_t)ev, (cudaStream_t)stream)); return IDOC_OK; }
int idoc_stream_wait_event(void* stream, void* ev) {
  CK(cudaStreamWaitEvent((cudaStream_t)stream, (cudaEvent_t)ev, 0));
  return IDOC_OK;
}
int idoc_event_sync(void* ev) { CK(cudaEventSynchronize((cudaEvent_t)ev)); return IDOC_OK; }
int idoc_event_elapsed_ms(void* a, void* b, float* ms) {
  CK(cudaEventSynchronize((cudaEvent_t)b));
  CK(cudaEventElapsedTime(ms, (cudaEvent_t)a, (cudaEvent_t)b));
  return IDOC_OK;
}
int idoc_mem_info(size_t* free_bytes, size_t* total_bytes) { CK(cudaMemGetInfo(free_bytes, total_bytes)); return IDOC_OK; }
int idoc_device_props(int* sm_count, int* cc_major, int* cc_minor, char* name, int name_len) {
  int dev = 0;
  CK(cudaGetDevice(&dev));
  cudaDeviceProp p;
  CK(cudaGetDeviceProperties(&p, dev));
  if (sm_count) *sm_count = p.multiProcessorCount;
  if (cc_major) *cc_major = p.major;
  if (cc_minor) *cc_minor = p.minor;
  if (name && name_len > 0) {
    strncpy(name, p.name, (size_t)name_len - 1);
    name[name_len - 1] = 0;
  }
  return IDOC_OK;
}


}  // extern "C"

// ------------------------------------------------------------------------------------------ blqr (shared controls)
template <typename T>
static int blqr_solve(cudaStream_t s, int64_t P, int T_, int bs, int n, int m, const void* Q, const void* q, const void* Qf,
                      const void* qf, const void* M, const void* R, const void* r, const void* A, const void* Bm, const void* d,
                      const void* x0, void* X, void* U, void* Nu, void* K, void* k, void* dCdc, int32_t* status, T* ws,
                      bool rollout, int q_shift, int zero_x0) {
  BlqrArgs<T> a;
  a.Q = (const T*)Q; a.q = (const T*)q; a.Qf = (const T*)Qf; a.qf = (const T*)qf; a.Mx = (const T*)M; a.R = (const T*)R;
  a.r = (const T*)r; a.A = (const T*)A; a.B = (const T*)Bm; a.d = (const T*)d; a.x0 = (const T*)x0;
  a.X = (T*)X; a.U = (T*)U; a.Nu = (T*)Nu; a.dCdc = (T*)dCdc; a.status = status;
  a.P = (int)P; a.T_ = T_; a.bs = bs; a.n = n; a.m = m; a.q_shift = q_shift; a.zero_x0 = zero_x0;
  a.ws_stride = blqr_ws_elems(T_, bs, n, m);
  a.ws = ws;
  T* gains = ws + (size_t)P * a.ws_stride;  // K[P,T,bs,m,n] k[P,T,m] when the caller does not ask for them
  a.K = K ? (T*)K : gains;
  a.k = k ? (T*)k : gains + (size_t)P * T_ * bs * m * n;
  {
    ProfScope ps(K_RICCATI, s);
    blqr_riccati_kernel<T><<<(unsigned)P, BL_THREADS, 0, s>>>(a);
  }
  CK(cudaGetLastError());
  if (rollout) {
    ProfScope ps(K_ROLLOUT, s);
    blqr_rollout_kernel<T><<<(unsigned)P, BL_THREADS, 0, s>>>(a);
  }
  CK(cudaGetLastError());
  return IDOC_OK;
}
static size_t blqr_fwd_elems(int64_t P, int T_, int bs, int n, int m) {
  return (size_t)P * ((size_t)blqr_ws_elems(T_, bs, n, m) + (size_t)T_ * bs * m * n + (size_t)T_ * m);
}
static int blqr_check(int dtype, int64_t P, int T_, int bs, int n, int m) {
  if (dtype != IDOC_F32 && dtype != IDOC_F64) return fail(IDOC_ERR_INVALID, "dtype must be IDOC_F32 or IDOC_F64");
  if (P <= 0 || T_ <= 0 || bs <= 0 || n <= 0 || m <= 0) return fail(IDOC_ERR_INVALID, "P, T, batch_size, n, m must be positive");
  if (n > BL_NMAX || m > BL_MMAX) return fail(IDOC_ERR_UNSUPPORTED, "blqr: per-system n <= 16, m <= 8 (any number of systems)");
  return IDOC_OK;
}
extern "C" {
size_t idoc_blqr_ws_bytes(int dtype, int64_t P, int T, int batch_size, int n, int m) {
  return blqr_fwd_elems(P, T, batch_size, n, m) * (dtype == IDOC_F32 ? 4 : 8);
}
size_t idoc_blqr_vjp_ws_bytes(int dtype, int64_t P, int T, int batch_size, int n, int m) {
  // adjoint solve workspace + its gains + qf' [P,bs,n] + adjoint solution uX, uNu [P,T,bs,n], uU [P,T,m]
  const size_t e = blqr_fwd_elems(P, T, batch_size, n, m) + (size_t)P * batch_size * n + 2 * (size_t)P * T * batch_size * n + (size_t)P * T * m;
  return e * (dtype == IDOC_F32 ? 4 : 8);
}
int idoc_blqr_fwd(void* stream, int dtype, int64_t P, int T, int batch_size, int n, int m, const void* Q, const void* q,
                  const void* Qf, const void* qf, const void* M, const void* R, const void* r, const void* A, const void* Bm,
                  const void* d, const void* x0, void* X, void* U, void* Nu, void* K, void* k, void* dCdc, int32_t* status,
                  void* ws, size_t ws_bytes) {
  int rc = blqr_check(dtype, P, T, batch_size, n, m);
  if (rc) return rc;
  const bool rollout = X != nullptr;
  const void* req[] = {Q, q, Qf, qf, M, R, r, A, Bm, d, ws};
  for (const void* p : req)
    if (p == nullptr) return fail(IDOC_ERR_INVALID, "null pointer argument");
  if (rollout && (x0 == nullptr || U == nullptr || Nu == nullptr)) return fail(IDOC_ERR_INVALID, "null pointer argument");
  if (ws_bytes < idoc_blqr_ws_bytes(dtype, P, T, batch_size, n, m)) return fail(IDOC_ERR_WORKSPACE, "ws too small");
  if (dtype == IDOC_F32)
    return blqr_solve<float>((cudaStream_t)stream, P, T, batch_size, n, m, Q, q, Qf, qf, M, R, r, A, Bm, d, x0, X, U, Nu, K, k, dCdc,
                             status, (float*)ws, rollout, 0, 0);
  return blqr_solve<double>((cudaStream_t)stream, P, T, batch_size, n, m, Q, q, Qf, qf, M, R, r, A, Bm, d, x0, X, U, Nu, K, k, dCdc,
                            status, (double*)ws, rollout, 0, 0);
}
}  // extern "C"
template <typename T>
static int blqr_vjp_typed(cudaStream_t s, int64_t P, int T_, int bs, int n, int m, const void* Q, const void* Qf, const void* M,
                          const void* R, const void* A, const void* Bm, const void* x0, const void* X, const void* U,
                          const void* Nu, const void* Xb, const void* Ub, const void* Nub, void* gQ, void* gq, void* gQf,
                          void* gqf, void* gM, void* gR, void* gr, void* gA, void* gB, void* gd, void* gx0, T* ws) {
  const size_t fe = blqr_fwd_elems(P, T_, bs, n, m);
  T* qfp = ws + fe;                                   // qf' = Xb[T-1]        [P, bs, n]
  T* uX = qfp + (size_t)P * bs * n;                   // adjoint solution
  T* uNu = uX + (size_t)P * T_ * bs * n;
  T* uU = uNu + (size_t)P * T_ * bs * n;
  // qf'[p] = Xb[p, T-1]: one strided copy
  CK(cudaMemcpy2DAsync(qfp, (size_t)bs * n * sizeof(T), (const T*)Xb + (size_t)(T_ - 1) * bs * n, (size_t)T_ * bs * n * sizeof(T),
                       (size_t)bs * n * sizeof(T), (size_t)P, cudaMemcpyDeviceToDevice, s));
  // adjoint problem: same quadratic blocks, q'[t] = Xb[t-1] (q_shift), r' = Ub, d' = Nub, x0' = 0
  int rc = blqr_solve<T>(s, P, T_, bs, n, m, Q, Xb, Qf, qfp, M, R, Ub, A, Bm, Nub, nullptr, uX, uU, uNu, nullptr, nullptr, nullptr,
                         nullptr, ws, true, 1, 1);
  if (rc) return rc;
  BlqrGradArgs<T> g;
  g.Mx = (const T*)M; g.A = (const T*)A; g.x0 = (const T*)x0; g.X = (const T*)X; g.U = (const T*)U; g.Nu = (const T*)Nu;
  g.uX = uX; g.uU = uU; g.uNu = uNu;
  g.gQ = (T*)gQ; g.gq = (T*)gq; g.gQf = (T*)gQf; g.gqf = (T*)gqf; g.gM = (T*)gM; g.gR = (T*)gR; g.gr = (T*)gr; g.gA = (T*)gA;
  g.gB = (T*)gB; g.gd = (T*)gd; g.gx0 = (T*)gx0;
  g.P = (int)P; g.T_ = T_; g.bs = bs; g.n = n; g.m = m;
  const long long tot = (long long)P * T_ * bs;
  {
    ProfScope ps(K_ADJ_FWD, s);
    blqr_grad_kernel<T><<<(unsigned)((tot + 127) / 128), 128, 0, s>>>(g);
  }
  CK(cudaGetLastError());
  return IDOC_OK;
}
extern "C" int idoc_blqr_vjp(void* stream, int dtype, int64_t P, int T, int batch_size, int n, int m, const void* Q, const void* Qf,
                             const void* M, const void* R, const void* A, const void* Bm, const void* x0, const void* X,
                             const void* U, const void* Nu, const void* Xb, const void* Ub, const void* Nub, void* gQ, void* gq,
                             void* gQf, void* gqf, void* gM, void* gR, void* gr, void* gA, void* gB, void* gd, void* gx0, void* ws,
                             size_t ws_bytes) {
  int rc = blqr_check(dtype, P, T, batch_size, n, m);
  if (rc) return rc;
  const void* req[] = {Q, Qf, M, R, A, Bm, x0, X, U, Nu, Xb, Ub, gQ, gq, gQf, gqf, gM, gR, gr, gA, gB, gd, gx0, ws};
  for (const void* p : req)
    if (p == nullptr) return fail(IDOC_ERR_INVALID, "null pointer argument");
  if (ws_bytes < idoc_blqr_vjp_ws_bytes(dtype, P, T, batch_size, n, m)) return fail(IDOC_ERR_WORKSPACE, "ws too small");
  if (dtype == IDOC_F32)
    return blqr_vjp_typed<float>((cudaStream_t)stream, P, T, batch_size, n, m, Q, Qf, M, R, A, Bm, x0, X, U, Nu, Xb, Ub, Nub, gQ, gq,
                                 gQf, gqf, gM, gR, gr, gA, gB, gd, gx0, (float*)ws);
  return blqr_vjp_typed<double>((cudaStream_t)stream, P, T, batch_size, n, m, Q, Qf, M, R, A, Bm, x0, X, U, Nu, Xb, Ub, Nub, gQ, gq,
                                gQf, gqf, gM, gR, gr, gA, gB, gd, gx0, (double*)ws);
}

// ------------------------------------------------------------------------------------------ FMA peak probe
// Measured FP32 / FP64 FMA throughput of this device (the roofline denominator of the FMA-bound tilqr configuration,
// "the builder must measure them with an FMA micro-benchmark in the same run"): every thread runs 16 independent
// multiply-add chains out of registers, 8 CTAs of 256 threads per SM, timed with CUDA events.
template <typename T>
__global__ void __launch_bounds__(256) fma_peak_kernel(T* out, int iters) {
  T c[16];
#pragma unroll
  for (int j = 0; j < 16; j++) c[j] = T(threadIdx.x + j);
  const T a = T(1.0001), b = T(0.5);
  for (int i = 0; i < iters; i++)
#pragma unroll
    for (int j = 0; j < 16; j++) c[j] = c[j] * a + b;
  T acc = T(0);
#pragma unroll
  for (int j = 0; j < 16; j++) acc += c[j];
  out[(size_t)blockIdx.x * blockDim.x + threadIdx.x] = acc;
}
template <typename T>
static int fma_peak_typed(cudaStream_t s, double* tflops) {
  int dev = 0, sms = 0;
  CK(cudaGetDevice(&dev));
  CK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
  const int grid = sms * 8, iters = sizeof(T) == 4 ? 16384 : 8192;
  T* out = nullptr;
  CK(cudaMalloc(&out, (size_t)grid * 256 * sizeof(T)));
  cudaEvent_t e0, e1;
  CK(cudaEventCreate(&e0));
  CK(cudaEventCreate(&e1));
  float best = 1e30f;
  for (int rep = 0; rep < 4; rep++) {
    CK(cudaEventRecord(e0, s));
    fma_peak_kernel<T><<<grid, 256, 0, s>>>(out, iters);
    CK(cudaEventRecord(e1, s));
    CK(cudaEventSynchronize(e1));
    float ms = 0.f;
    CK(cudaEventElapsedTime(&ms, e0, e1));
    if (rep > 0 && ms < best) best = ms;
  }
  cudaEventDestroy(e0);
  cudaEventDestroy(e1);
  cudaFree(out);
  g_launches += 4;
  *tflops = 2.0 * 16.0 * iters * (double)grid * 256.0 / (best * 1e-3) / 1e12;
  return IDOC_OK;
}
extern "C" int idoc_fma_peak(void* stream, int dtype, double* tflops) {
  if (tflops == nullptr) return fail(IDOC_ERR_INVALID, "null pointer argument");
  if (dtype == IDOC_F32) return fma_peak_typed<float>((cudaStream_t)stream, tflops);
  if (dtype == IDOC_F64) return fma_peak_typed<double>((cudaStream_t)stream, tflops);
  return fail(IDOC_ERR_INVALID, "dtype must be IDOC_F32 or IDOC_F64");
}

// ------------------------------------------------------------------------------------------ NCCL (lazy, optional)
// The only exchange step of the path: all-reduce of batch-reduced parameter gradients (idoc_lqr_vjp with
// reduce_batch = 1) across the GPUs of a box.  libnccl is dlopen'ed on first use so that the library loads
// (and every single-GPU entry point works) on machines without NCCL.
#include <dlfcn.h>
struct Id128 {  // ncclUniqueId is passed by value: 128 bytes
  char b[128];
};
struct NcclApi {
  void* h = nullptr;
  int (*GetUniqueId)(void*) = nullptr;
  int (*CommInitRank)(void**, int, Id128, int) = nullptr;
  int (*AllReduce)(const void*, void*, size_t, int, int, void*, cudaStream_t) = nullptr;
  int (*CommDestroy)(void*) = nullptr;
  const char* (*GetErrorString)(int) = nullptr;
};
static NcclApi g_nccl;
static std::mutex g_nccl_mu;
static int nccl_load() {
  std::lock_guard<std::mutex> lk(g_nccl_mu);
  if (g_nccl.h) return IDOC_OK;
  void* h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
  if (!h) h = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
  if (!h) return fail(IDOC_ERR_NCCL, "cannot dlopen libnccl.so.2: %s", dlerror());
  g_nccl.GetUniqueId = (int (*)(void*))dlsym(h, "ncclGetUniqueId");
  g_nccl.CommInitRank = (int (*)(void**, int, Id128, int))dlsym(h, "ncclCommInitRank");
  g_nccl.AllReduce = (int (*)(const void*, void*, size_t, int, int, void*, cudaStream_t))dlsym(h, "ncclAllReduce");
  g_nccl.CommDestroy = (int (*)(void*))dlsym(h, "ncclCommDestroy");
  g_nccl.GetErrorString = (const char* (*)(int))dlsym(h, "ncclGetErrorString");
  if (!g_nccl.GetUniqueId || !g_nccl.CommInitRank || !g_nccl.AllReduce || !g_nccl.CommDestroy)
    return fail(IDOC_ERR_NCCL, "libnccl is missing a required symbol");
  g_nccl.h = h;
  return IDOC_OK;
}
static int nccl_fail(int rc, const char* where) {
  snprintf(g_err, sizeof(g_err), "%s: %s", where, g_nccl.GetErrorString ? g_nccl.GetErrorString(rc) : "nccl error");
  return IDOC_ERR_NCCL;
}

extern "C" {
int idoc_nccl_unique_id(void* id128) {
  int rc = nccl_load();
  if (rc) return rc;
  rc = g_nccl.GetUniqueId(id128);
  return rc ? nccl_fail(rc, "ncclGetUniqueId") : IDOC_OK;
}
int idoc_nccl_init(const void* id128, int rank, int nranks, void** comm) {
  int rc = nccl_load();
  if (rc) return rc;
  Id128 id;
  memcpy(&id, id128, sizeof(id));
  rc = g_nccl.CommInitRank(comm, nranks, id, rank);
  return rc ? nccl_fail(rc, "ncclCommInitRank") : IDOC_OK;
}
int idoc_nccl_allreduce_sum(void* comm, void* stream, int dtype, void* buf, size_t count) {
  if (!g_nccl.h) return fail(IDOC_ERR_NCCL, "NCCL not initialised");
  if (dtype != IDOC_F32 && dtype != IDOC_F64) return fail(IDOC_ERR_INVALID, "dtype must be IDOC_F32 or IDOC_F64");
  const int nccl_dtype = dtype == IDOC_F32 ? 7 : 8;  // ncclFloat32 = 7, ncclFloat64 = 8
  int rc = g_nccl.AllReduce(buf, buf, count, nccl_dtype, /*ncclSum*/ 0, comm, (cudaStream_t)stream);
  return rc ? nccl_fail(rc, "ncclAllReduce") : IDOC_OK;
}
int idoc_nccl_destroy(void* comm) {
  if (!g_nccl.h) return IDOC_OK;
  int rc = g_nccl.CommDestroy(comm);
  return rc ? nccl_fail(rc, "ncclCommDestroy") : IDOC_OK;
}
}  // extern "C"
