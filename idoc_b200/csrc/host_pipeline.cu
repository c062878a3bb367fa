// Host-buffer entry points of the C ABI (include/idoc_b200.h, "host pipeline"): the whole hot path — fwd solve + implicit VJP —
// for a batch of LQR problems that lives in (pinned) HOST memory, results back in host memory.  This is the call a host
// program without device arrays makes (the reference's users pass NumPy / jax arrays to Solver.implicit and jax.grad,
// idoc/lqr.py:167-180, tests/test_lqr.py:65-90), and what bench.py's `e2e` number times.
//
// The batch is streamed through the device in chunks on three CUDA streams: the H2D copy of chunk c+1, the kernels of chunk c
// and the D2H copy of chunk c-1 overlap (PCIe is full duplex), each chunk in one of `nslots` device slots that are recycled in
// order.  Cotangent of the VJP: (X, U, 0), i.e. the loss 1/2 |X|^2 + 1/2 |U|^2 of the reference's tests (tests/test_lqr.py:52-58).
// reduce_batch = 1: gradients summed over the batch — every chunk reduces on the device (grad_reduce.cuh) and sends ONE
// contiguous buffer of the ten leaves; idoc_lqr_host_finish adds the per-chunk sums on the host.
#include "../../include/idoc_b200.h"

#include <cuda_runtime.h>

#include <cstdint>
#include <cstdio>
#include <cstring>
#include <vector>

namespace {

thread_local char h_err[256] = "";

struct Slot {
  void* prob[11];   // Q q Qf qf M R r A B d x0
  void *X, *U, *Nu, *ws, *ws2;
  int32_t* status;
  void* grad[11];   // per-problem mode: 11 leaves; batch-summed mode: grad[0..9] are views of gbuf, grad[10] = gx0
  void* gbuf;
  cudaEvent_t e_in, e_comp, e_out;
  bool used;
};

struct Pipe {
  int dtype, T, n, m, nslots, reduce;
  int64_t chunk;
  size_t sz, ws_bytes, ws2_bytes, gbuf_elems;
  size_t row[11];      // elements per problem of every problem leaf (x0 last)
  size_t goff[10];     // element offsets of the ten leaves inside the batch-summed buffer
  size_t gcnt[10];     // elements of each batch-summed leaf
  std::vector<Slot> slots;
  cudaStream_t s_in, s_comp, s_out;
  cudaEvent_t ev0, ev1;
  // batch-summed mode: pinned staging [2][max_chunks][gbuf_elems], two batches in flight (ping-pong), so that back-to-back
  // enqueued batches (wait = 0) never synchronise on the batch right before them
  void* stage;
  int64_t stage_chunks;
  struct Pending {
    bool active;
    int64_t chunks;
    void* g[10];
    cudaEvent_t done;
  } pend[2];
  int parity;
  uint64_t h2d, d2h;
};

int flush_pending(Pipe* p, int par);

int fail(int code, const char* msg) {
  snprintf(h_err, sizeof(h_err), "%s", msg);
  return code;
}
#define HCK(call)                                                     \
  do {                                                                \
    cudaError_t e_ = (call);                                          \
    if (e_ != cudaSuccess) {                                          \
      snprintf(h_err, sizeof(h_err), "%s: %s", #call, cudaGetErrorString(e_)); \
      return IDOC_ERR_CUDA;                                           \
    }                                                                 \
  } while (0)

void leaf_rows(int T, int n, int m, size_t* row) {
  const size_t t = (size_t)T;
  row[0] = t * n * n; row[1] = t * n; row[2] = (size_t)n * n; row[3] = n; row[4] = t * n * m; row[5] = t * m * m; row[6] = t * m;
  row[7] = t * n * n; row[8] = t * n * m; row[9] = t * n; row[10] = n;
}

// wait for the batch staged in half `par` and add its per-chunk sums into the host gradient leaves
int flush_pending(Pipe* p, int par) {
  Pipe::Pending& pd = p->pend[par];
  if (!pd.active) return IDOC_OK;
  HCK(cudaEventSynchronize(pd.done));
  const char* base = (const char*)p->stage + (size_t)par * p->stage_chunks * p->gbuf_elems * p->sz;
  for (int k = 0; k < 10; k++) {
    for (size_t e = 0; e < p->gcnt[k]; e++) {
      if (p->dtype == IDOC_F64) {
        double acc = 0.0;
        for (int64_t c = 0; c < pd.chunks; c++) acc += ((const double*)base)[(size_t)c * p->gbuf_elems + p->goff[k] + e];
        ((double*)pd.g[k])[e] = acc;
      } else {
        float acc = 0.f;
        for (int64_t c = 0; c < pd.chunks; c++) acc += ((const float*)base)[(size_t)c * p->gbuf_elems + p->goff[k] + e];
        ((float*)pd.g[k])[e] = acc;
      }
    }
  }
  pd.active = false;
  return IDOC_OK;
}

}  // namespace

extern "C" {

const char* idoc_lqr_host_last_error(void) { return h_err; }

int idoc_lqr_host_create(void** handle, int dtype, int64_t chunk, int T, int n, int m, int nslots, int reduce_batch) {
  if (handle == nullptr) return fail(IDOC_ERR_INVALID, "null handle");
  if (dtype != IDOC_F32 && dtype != IDOC_F64) return fail(IDOC_ERR_INVALID, "dtype must be IDOC_F32 or IDOC_F64");
  if (chunk <= 0 || T <= 0 || n <= 0 || m <= 0 || nslots < 2) return fail(IDOC_ERR_INVALID, "chunk, T, n, m positive; nslots >= 2");
  if (!idoc_lqr_supported(dtype, n, m)) return fail(IDOC_ERR_UNSUPPORTED, "unsupported (n, m)");
  Pipe* p = new Pipe();
  p->dtype = dtype; p->T = T; p->n = n; p->m = m; p->nslots = nslots; p->reduce = reduce_batch ? 1 : 0; p->chunk = chunk;
  p->sz = dtype == IDOC_F32 ? 4 : 8;
  leaf_rows(T, n, m, p->row);
  p->ws_bytes = idoc_lqr_ws_bytes(dtype, chunk, T, n, m);
  p->ws2_bytes = p->reduce ? idoc_lqr_vjp_reduce_ws_bytes(dtype, chunk, T, n, m) : idoc_lqr_vjp_ws_bytes(dtype, chunk, T, n, m);
  size_t o = 0;
  const size_t e16 = 16 / p->sz;
  for (int k = 0; k < 10; k++) {
    p->goff[k] = o;
    p->gcnt[k] = p->row[k];
    o += (p->row[k] + e16 - 1) / e16 * e16;
  }
  p->gbuf_elems = o;
  p->stage = nullptr;
  p->stage_chunks = 0;
  p->parity = 0;
  for (auto& pd : p->pend) {
    pd.active = false;
    pd.chunks = 0;
    pd.done = nullptr;
  }
  p->h2d = p->d2h = 0;
  HCK(cudaStreamCreateWithFlags(&p->s_in, cudaStreamNonBlocking));
  HCK(cudaStreamCreateWithFlags(&p->s_comp, cudaStreamNonBlocking));
  HCK(cudaStreamCreateWithFlags(&p->s_out, cudaStreamNonBlocking));
  HCK(cudaEventCreate(&p->ev0));
  HCK(cudaEventCreate(&p->ev1));
  for (auto& pd : p->pend) HCK(cudaEventCreateWithFlags(&pd.done, cudaEventDisableTiming));
  p->slots.resize(nslots);
  for (auto& s : p->slots) {
    memset(&s, 0, sizeof(s));
    for (int k = 0; k < 11; k++) HCK(cudaMalloc(&s.prob[k], p->row[k] * chunk * p->sz));
    HCK(cudaMalloc(&s.X, (size_t)chunk * T * n * p->sz));
    HCK(cudaMalloc(&s.U, (size_t)chunk * T * m * p->sz));
    HCK(cudaMalloc(&s.Nu, (size_t)chunk * T * n * p->sz));
    HCK(cudaMalloc(&s.status, (size_t)chunk * 4));
    HCK(cudaMalloc(&s.ws, p->ws_bytes));
    HCK(cudaMalloc(&s.ws2, p->ws2_bytes));
    if (p->reduce) {
      HCK(cudaMalloc(&s.gbuf, p->gbuf_elems * p->sz));
      for (int k = 0; k < 10; k++) s.grad[k] = (char*)s.gbuf + p->goff[k] * p->sz;
      HCK(cudaMalloc(&s.grad[10], p->row[10] * chunk * p->sz));
    } else {
      for (int k = 0; k < 11; k++) HCK(cudaMalloc(&s.grad[k], p->row[k] * chunk * p->sz));
    }
    HCK(cudaEventCreateWithFlags(&s.e_in, cudaEventDisableTiming));
    HCK(cudaEventCreateWithFlags(&s.e_comp, cudaEventDisableTiming));
    HCK(cudaEventCreateWithFlags(&s.e_out, cudaEventDisableTiming));
  }
  *handle = p;
  return IDOC_OK;
}

int idoc_lqr_host_finish(void* handle) {
  Pipe* p = static_cast<Pipe*>(handle);
  if (p == nullptr) return fail(IDOC_ERR_INVALID, "null handle");
  HCK(cudaStreamSynchronize(p->s_out));
  for (int k = 0; k < 2; k++) {  // older batch first
    const int rc = flush_pending(p, (p->parity + k) & 1);
    if (rc) return rc;
  }
  return IDOC_OK;
}

int idoc_lqr_host_solve_vjp(void* handle, int64_t B, const void* Q, const void* q, const void* Qf, const void* qf, const void* M,
                            const void* R, const void* r, const void* A, const void* Bm, const void* d, const void* x0, void* X,
                            void* U, void* Nu, void* gQ, void* gq, void* gQf, void* gqf, void* gM, void* gR, void* gr, void* gA,
                            void* gB, void* gd, void* gx0, int wait, float* ms) {
  Pipe* p = static_cast<Pipe*>(handle);
  if (p == nullptr || B <= 0) return fail(IDOC_ERR_INVALID, "null handle or empty batch");
  const void* hp[11] = {Q, q, Qf, qf, M, R, r, A, Bm, d, x0};
  void* hg[11] = {gQ, gq, gQf, gqf, gM, gR, gr, gA, gB, gd, gx0};
  for (int k = 0; k < 11; k++)
    if (hp[k] == nullptr || hg[k] == nullptr) return fail(IDOC_ERR_INVALID, "null host pointer");
  if (X == nullptr || U == nullptr || Nu == nullptr) return fail(IDOC_ERR_INVALID, "null host pointer");
  const int64_t nchunks = (B + p->chunk - 1) / p->chunk;
  const int par = p->parity;
  if (p->reduce) {
    if (nchunks > p->stage_chunks) {  // (re)allocate the staging area: flush everything that still points into the old one
      const int rc = idoc_lqr_host_finish(p);
      if (rc) return rc;
      if (p->stage) HCK(cudaFreeHost(p->stage));
      HCK(cudaHostAlloc(&p->stage, (size_t)2 * nchunks * p->gbuf_elems * p->sz, cudaHostAllocDefault));
      p->stage_chunks = nchunks;
    }
    const int rc = flush_pending(p, par);  // the batch before the previous one used this half
    if (rc) return rc;
  }
  const size_t sz = p->sz;
  const int T = p->T, n = p->n, m = p->m;
  p->h2d = p->d2h = 0;
  HCK(cudaEventRecord(p->ev0, p->s_in));
  int64_t b0 = 0;
  for (int64_t c = 0; c < nchunks; c++) {
    const int64_t nb = B - b0 < p->chunk ? B - b0 : p->chunk;
    Slot& s = p->slots[(size_t)(c % p->nslots)];
    if (s.used) HCK(cudaStreamWaitEvent(p->s_in, s.e_out, 0));  // the slot's previous results have left the device
    s.used = true;
    for (int k = 0; k < 11; k++) {
      const size_t bytes = p->row[k] * (size_t)nb * sz;
      HCK(cudaMemcpyAsync(s.prob[k], (const char*)hp[k] + p->row[k] * (size_t)b0 * sz, bytes, cudaMemcpyHostToDevice, p->s_in));
      p->h2d += bytes;
    }
    HCK(cudaEventRecord(s.e_in, p->s_in));
    HCK(cudaStreamWaitEvent(p->s_comp, s.e_in, 0));
    int rc = idoc_lqr_fwd(p->s_comp, p->dtype, nb, T, n, m, s.prob[0], s.prob[1], s.prob[2], s.prob[3], s.prob[4], s.prob[5], s.prob[6],
                          s.prob[7], s.prob[8], s.prob[9], s.prob[10], s.X, s.U, s.Nu, nullptr, nullptr, nullptr, s.status, s.ws,
                          p->ws_bytes);
    if (rc) return fail(rc, idoc_last_error());
    rc = idoc_lqr_vjp(p->s_comp, p->dtype, nb, T, n, m, s.prob[4], s.prob[7], s.prob[8], s.prob[10], s.X, s.U, s.Nu, s.ws, s.X, s.U,
                      nullptr, s.grad[0], s.grad[1], s.grad[2], s.grad[3], s.grad[4], s.grad[5], s.grad[6], s.grad[7], s.grad[8],
                      s.grad[9], s.grad[10], s.ws2, p->ws2_bytes, p->reduce);
    if (rc) return fail(rc, idoc_last_error());
    HCK(cudaEventRecord(s.e_comp, p->s_comp));
    HCK(cudaStreamWaitEvent(p->s_out, s.e_comp, 0));
    const size_t sol[3] = {(size_t)T * n, (size_t)T * m, (size_t)T * n};
    void* hs[3] = {X, U, Nu};
    void* ds[3] = {s.X, s.U, s.Nu};
    for (int k = 0; k < 3; k++) {
      const size_t bytes = sol[k] * (size_t)nb * sz;
      HCK(cudaMemcpyAsync((char*)hs[k] + sol[k] * (size_t)b0 * sz, ds[k], bytes, cudaMemcpyDeviceToHost, p->s_out));
      p->d2h += bytes;
    }
    if (p->reduce) {
      const size_t bytes = p->gbuf_elems * sz;
      HCK(cudaMemcpyAsync((char*)p->stage + ((size_t)par * p->stage_chunks + (size_t)c) * bytes, s.gbuf, bytes, cudaMemcpyDeviceToHost, p->s_out));
      const size_t bx = p->row[10] * (size_t)nb * sz;
      HCK(cudaMemcpyAsync((char*)gx0 + p->row[10] * (size_t)b0 * sz, s.grad[10], bx, cudaMemcpyDeviceToHost, p->s_out));
      p->d2h += bytes + bx;
    } else {
      for (int k = 0; k < 11; k++) {
        const size_t bytes = p->row[k] * (size_t)nb * sz;
        HCK(cudaMemcpyAsync((char*)hg[k] + p->row[k] * (size_t)b0 * sz, s.grad[k], bytes, cudaMemcpyDeviceToHost, p->s_out));
        p->d2h += bytes;
      }
    }
    HCK(cudaEventRecord(s.e_out, p->s_out));
    b0 += nb;
  }
  HCK(cudaEventRecord(p->ev1, p->s_out));
  if (p->reduce) {
    Pipe::Pending& pd = p->pend[par];
    pd.active = true;
    pd.chunks = nchunks;
    for (int k = 0; k < 10; k++) pd.g[k] = hg[k];
    HCK(cudaEventRecord(pd.done, p->s_out));
    p->parity ^= 1;
  }
  if (wait) {
    HCK(cudaEventSynchronize(p->ev1));
    if (ms) HCK(cudaEventElapsedTime(ms, p->ev0, p->ev1));
    return idoc_lqr_host_finish(p);
  }
  return IDOC_OK;
}

int idoc_lqr_host_elapsed_ms(void* handle, float* ms) {
  Pipe* p = static_cast<Pipe*>(handle);
  if (p == nullptr || ms == nullptr) return fail(IDOC_ERR_INVALID, "null argument");
  HCK(cudaEventSynchronize(p->ev1));
  HCK(cudaEventElapsedTime(ms, p->ev0, p->ev1));
  return IDOC_OK;
}

int idoc_lqr_host_streams(void* handle, void** s_in, void** s_out) {
  Pipe* p = static_cast<Pipe*>(handle);
  if (p == nullptr) return fail(IDOC_ERR_INVALID, "null handle");
  if (s_in) *s_in = p->s_in;
  if (s_out) *s_out = p->s_out;
  return IDOC_OK;
}

int idoc_lqr_host_bytes(void* handle, uint64_t* h2d, uint64_t* d2h) {
  Pipe* p = static_cast<Pipe*>(handle);
  if (p == nullptr) return fail(IDOC_ERR_INVALID, "null handle");
  if (h2d) *h2d = p->h2d;
  if (d2h) *d2h = p->d2h;
  return IDOC_OK;
}

int idoc_lqr_host_destroy(void* handle) {
  Pipe* p = static_cast<Pipe*>(handle);
  if (p == nullptr) return IDOC_OK;
  cudaStreamSynchronize(p->s_in);
  cudaStreamSynchronize(p->s_comp);
  cudaStreamSynchronize(p->s_out);
  for (auto& s : p->slots) {
    for (int k = 0; k < 11; k++) cudaFree(s.prob[k]);
    cudaFree(s.X); cudaFree(s.U); cudaFree(s.Nu); cudaFree(s.status); cudaFree(s.ws); cudaFree(s.ws2);
    if (p->reduce) {
      cudaFree(s.gbuf);
      cudaFree(s.grad[10]);
    } else {
      for (int k = 0; k < 11; k++) cudaFree(s.grad[k]);
    }
    cudaEventDestroy(s.e_in); cudaEventDestroy(s.e_comp); cudaEventDestroy(s.e_out);
  }
  if (p->stage) cudaFreeHost(p->stage);
  for (auto& pd : p->pend) cudaEventDestroy(pd.done);
  cudaStreamDestroy(p->s_in); cudaStreamDestroy(p->s_comp); cudaStreamDestroy(p->s_out);
  cudaEventDestroy(p->ev0); cudaEventDestroy(p->ev1);
  delete p;
  return IDOC_OK;
}

}  // extern "C"
