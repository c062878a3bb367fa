/* idoc_b200 — C ABI of the B200-native batched LQR hot path.
 *
 * Drop-in boundary for tachukao/idoc's solver entry points.  The reference has no FFI of its own
 * (pure Python/JAX); its boundary is the Python call  Solver.direct(params) / .implicit(params) /
 * .kkt(state, params)  on the pytrees of idoc/lqr.py:16-59 and idoc/typs.py:6-16.  Each entry point
 * below cites the reference function it replaces.  An XLA typed-FFI handler (or any other host
 * binding: ctypes, cgo, JNI ...) only unpacks its buffers and forwards to these functions.
 *
 * Conventions
 *   - every pointer is a DEVICE pointer owned by the caller unless the name ends in _host;
 *   - arrays are dense row-major with a leading batch axis (what jax.vmap of the reference sees):
 *       Q[B,T,n,n] q[B,T,n] Qf[B,n,n] qf[B,n] M[B,T,n,m] R[B,T,m,m] r[B,T,m] A[B,T,n,n] Bm[B,T,n,m]
 *       d[B,T,n] x0[B,n];   X[B,T,n] (X[t] = x_{t+1}, x0 excluded), U[B,T,m], Nu[B,T,n];
 *     base pointers must be 16-byte aligned;
 *   - Q, R, Qf are used through their symmetric parts (the reference's kkt applies .symm(),
 *     idoc/lqr.py:130; its tests symmetrise before calling direct, tests/test_lqr.py:61,66);
 *   - calls only enqueue work on `stream` (no allocation, no synchronisation): safe under CUDA graphs
 *     and XLA; scratch comes from caller-provided workspaces sized by the *_ws_bytes functions;
 *   - return value: 0 = OK, <0 = error (IDOC_ERR_*).  Numerical events (eigen-shift taken, Cholesky
 *     pivot clamped) are not errors: they are reported per problem in `status` (IDOC_STATUS_* bits).
 */
#ifndef IDOC_B200_H
#define IDOC_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define IDOC_F32 0
#define IDOC_F64 1

#define IDOC_OK 0
#define IDOC_ERR_INVALID (-1)     /* null pointer, non-positive size, misaligned base pointer */
#define IDOC_ERR_UNSUPPORTED (-2) /* (n, m, dtype) has no compiled kernel                       */
#define IDOC_ERR_CUDA (-3)        /* launch / runtime failure; see idoc_last_error()            */
#define IDOC_ERR_WORKSPACE (-4)   /* workspace too small                                        */
#define IDOC_ERR_NCCL (-5)

#define IDOC_STATUS_SHIFT 1      /* eigen-shift of idoc/lqr.py:86-88 was taken at some step */
#define IDOC_STATUS_NONFINITE 2
#define IDOC_STATUS_CHOL_CLAMP 4 /* a Cholesky pivot of the shifted Guu had to be clamped   */

const char* idoc_version(void);
const char* idoc_last_error(void);
/* 1: tuned kernels compiled for (n, m) (csrc/shapes.def); 2: served by the shape-generic kernels
 * (any 1 <= n <= 128, 1 <= m <= 32); 0: unsupported */
int idoc_lqr_supported(int dtype, int n, int m);

/* ---- time-varying LQR ---------------------------------------------------------------------------
 * idoc_lqr_fwd replaces  idoc.lqr.build(T).direct(params)   (idoc/lqr.py:170-175), i.e.
 *   backward (lqr.py:62-107) -> forward (lqr.py:153-164) -> adjoint (lqr.py:110-125).
 * Outputs X,U,Nu; optional gains K[B,T,m,n], k[B,T,m] (lqr.Gains, lqr.py:16-20; NULL to skip),
 * optional dCdc[B,2] (the expected_change terms of lqr.py:93-94,104-105), optional status[B].
 * `ws` receives the residuals the VJP needs (idoc_lqr_ws_bytes bytes). */
size_t idoc_lqr_ws_bytes(int dtype, int64_t B, int T, int n, int m);
int idoc_lqr_fwd(void* stream, int dtype, int64_t B, int T, int n, int m,
                 const void* Q, const void* q, const void* Qf, const void* qf, const void* M,
                 const void* R, const void* r, const void* A, const void* Bm, const void* d, const void* x0,
                 void* X, void* U, void* Nu, void* K, void* k, void* dCdc, int32_t* status,
                 void* ws, size_t ws_bytes);

/* idoc_lqr_backward replaces idoc.lqr.backward(lqr, T, return_expected_change=True) (lqr.py:62-107)
 * alone (used by the iLQR outer loop, idoc/ilqr.py:231-233): gains K,k + dCdc, residuals in ws. */
int idoc_lqr_backward(void* stream, int dtype, int64_t B, int T, int n, int m,
                      const void* Q, const void* q, const void* Qf, const void* qf, const void* M,
                      const void* R, const void* r, const void* A, const void* Bm, const void* d,
                      void* K, void* k, void* dCdc, int32_t* status, void* ws, size_t ws_bytes);

/* idoc_lqr_vjp replaces the backward pass of  idoc.lqr.build(T).implicit  (lqr.py:177:
 * jaxopt.implicit_diff.custom_root(kkt, solve=solve_cg)): given the cotangent (Xb,Ub,Nub) of the
 * solution it returns the cotangent of every Params leaf, computed exactly as one adjoint LQR
 * solve that reuses the forward factorisation in `ws` (instead of the reference's CG iteration).
 * Nub may be NULL (loss independent of Nu, as in every reference test).
 * reduce_batch = 0: g* have the shapes of their primals (per-problem gradients);
 * reduce_batch = 1: g* are summed over the batch axis (shapes without B; overwritten, no zeroing needed); gx0 stays
 *                   per-problem.  The sums are formed without atomics (the sweeps store the adjoint solution, a second
 *                   pass contracts over the batch tile by tile and adds chunk partials in a fixed order), so the result
 *                   is bit-reproducible; ws2 must then hold idoc_lqr_vjp_reduce_ws_bytes bytes.  When the ten leaves
 *                   are carved from ONE buffer (any order, 16-byte aligned), that buffer is what a single
 *                   idoc_nccl_allreduce_sum call combines across GPUs. */
size_t idoc_lqr_vjp_ws_bytes(int dtype, int64_t B, int T, int n, int m);
size_t idoc_lqr_vjp_reduce_ws_bytes(int dtype, int64_t B, int T, int n, int m);
int idoc_lqr_vjp(void* stream, int dtype, int64_t B, int T, int n, int m,
                 const void* M, const void* A, const void* Bm, const void* x0,
                 const void* X, const void* U, const void* Nu, const void* ws,
                 const void* Xb, const void* Ub, const void* Nub,
                 void* gQ, void* gq, void* gQf, void* gqf, void* gM, void* gR, void* gr,
                 void* gA, void* gB, void* gd, void* gx0,
                 void* ws2, size_t ws2_bytes, int reduce_batch);

/* idoc_lqr_kkt replaces idoc.lqr.kkt(state, params) (lqr.py:128-150): residuals rX,rU,rNu. */
int idoc_lqr_kkt(void* stream, int dtype, int64_t B, int T, int n, int m,
                 const void* Q, const void* q, const void* Qf, const void* qf, const void* M,
                 const void* R, const void* r, const void* A, const void* Bm, const void* d, const void* x0,
                 const void* X, const void* U, const void* Nu, void* rX, void* rU, void* rNu);

/* ---- time-invariant LQR -------------------------------------------------------------------------
 * idoc_tilqr_fwd replaces idoc.tilqr.build(T).direct (idoc/tilqr.py:56-99): Q[B,n,n] Qf[B,n,n] R[B,m,m]
 * A[B,n,n] Bm[B,n,m] x0[B,n] -> X[B,T,n] U[B,T,m] Nu[B,T,n].  ws as for idoc_lqr_fwd.
 * idoc_tilqr_vjp replaces the backward pass of custom_root(kkt)(direct) (idoc/tilqr.py:101): cotangents of
 * Q, Qf, R, A, B (summed over time, per problem) and x0. */
int idoc_tilqr_fwd(void* stream, int dtype, int64_t B, int T, int n, int m,
                   const void* Q, const void* Qf, const void* R, const void* A, const void* Bm, const void* x0,
                   void* X, void* U, void* Nu, int32_t* status, void* ws, size_t ws_bytes);
int idoc_tilqr_vjp(void* stream, int dtype, int64_t B, int T, int n, int m,
                   const void* A, const void* Bm, const void* x0, const void* X, const void* U, const void* Nu,
                   const void* ws, const void* Xb, const void* Ub, const void* Nub,
                   void* gQ, void* gQf, void* gR, void* gA, void* gB, void* gx0, void* ws2, size_t ws2_bytes);

/* ---- batch LQR with SHARED controls ---------------------------------------------------------------
 * idoc_blqr_fwd replaces idoc.blqr.build(T).direct (idoc/blqr.py:245-261: backward :128-157 -> forward :229-242 -> adjoint
 * :160-182): `batch_size` systems driven by ONE control sequence.  Shapes as in the reference with an optional leading
 * problem axis P:  Q[P,T,bs,n,n] q[P,T,bs,n] Qf[P,bs,n,n] qf[P,bs,n] M[P,T,bs,n,m] R[P,T,m,m] r[P,T,m] A[P,T,bs,n,n]
 * Bm[P,T,bs,n,m] d[P,T,bs,n] x0[P,bs,n] -> X[P,T,bs,n] U[P,T,m] Nu[P,T,bs,n]; optional gains K[P,T,bs,m,n], k[P,T,m]
 * (blqr.Gains, idoc/blqr.py:17-21), dCdc[P,2], status[P].  X = U = Nu = NULL: backward pass only (idoc/blqr.py:128-157).
 * The reference's dense [bs,n,bs,n] value tensor is carried as bs diagonal n x n blocks minus a low-rank term
 * (csrc/blqr.cuh): work and memory are linear in batch_size; per-system n <= 16, m <= 8.
 * idoc_blqr_vjp replaces the backward pass of custom_root(kkt, solve_cg)(direct) (idoc/blqr.py:258): the cotangent of every
 * Params leaf (shapes of the primals) from the cotangent (Xb, Ub, Nub) of the solution; Nub may be NULL. */
size_t idoc_blqr_ws_bytes(int dtype, int64_t P, int T, int batch_size, int n, int m);
size_t idoc_blqr_vjp_ws_bytes(int dtype, int64_t P, int T, int batch_size, int n, int m);
int idoc_blqr_fwd(void* stream, int dtype, int64_t P, int T, int batch_size, int n, int m,
                  const void* Q, const void* q, const void* Qf, const void* qf, const void* M, const void* R, const void* r,
                  const void* A, const void* Bm, const void* d, const void* x0, void* X, void* U, void* Nu, void* K, void* k,
                  void* dCdc, int32_t* status, void* ws, size_t ws_bytes);
int idoc_blqr_vjp(void* stream, int dtype, int64_t P, int T, int batch_size, int n, int m,
                  const void* Q, const void* Qf, const void* M, const void* R, const void* A, const void* Bm, const void* x0,
                  const void* X, const void* U, const void* Nu, const void* Xb, const void* Ub, const void* Nub,
                  void* gQ, void* gq, void* gQf, void* gqf, void* gM, void* gR, void* gr, void* gA, void* gB, void* gd,
                  void* gx0, void* ws, size_t ws_bytes);

/* ---- iLQR for compiled problems ----------------------------------------------------------------
 * The reference's ilqr.Problem holds Python callables (idoc/ilqr.py:94-117); a kernel cannot call those, so a
 * problem family is a device functor selected by id, parameterised by theta[B, theta_dim]:
 *   IDOC_PROBLEM_RELU_RNN  the reference's own test system (tests/test_ilqr.py:14-64), any n, m <= 32,
 *                          theta = (Q, q, Qf, R, r, A, B) flattened in that order;
 *   IDOC_PROBLEM_PENDULUM  damped pendulum swing-up, n = 2, m = 1, theta = [h, g/l, b, c, w1, w2, wu, f1, f2].
 * idoc_ilqr_solve replaces ilqr.build(problem, maxiter, thres, line_search).direct(init, params)
 * (idoc/ilqr.py:221-268 with line_search.py:7-42): X, U hold the initial trajectory on entry and the solution on
 * exit, Nu receives the co-states (ilqr.py:266-267); iters / ls_failed are optional per-problem counters.
 * allow_sync != 0 lets the call read one int per iteration back to stop as soon as every problem has converged.
 * idoc_ilqr_vjp replaces the backward pass of custom_root(kkt, solve_cg(tol=1e-8))(ilqr) (idoc/ilqr.py:270-272):
 * cotangents of theta and x0 from the cotangent of the solution, by one adjoint LQR solve on the Hessian of the
 * Lagrangian at the solution.
 *
 * `systems` >= 1 is the batch size of the reference's bilqr (idoc/bilqr.py:19-221): that many copies of the system
 * (same theta, different initial states) driven by ONE control sequence, cost summed over the copies.  States are
 * then laid out lifted: x0[B, systems*n], X/Nu[B, T, systems*n], U[B, T, m]; systems * n <= 32.  systems = 1 is
 * plain iLQR (idoc/ilqr.py). */
#define IDOC_PROBLEM_RELU_RNN 0
#define IDOC_PROBLEM_PENDULUM 1
#define IDOC_PROBLEM_RELU_AX 2 /* f = relu(A x) + B u + 0.5, cost weights 1e-6: tests/test_bilqr.py:39-83 */
/* Families written once over the scalar type and differentiated on the device by forward-mode AD (csrc/ad.cuh), the
 * counterpart of the reference's jax.grad / jacfwd in make_lqr_approx (idoc/ilqr.py:127-158):
 *   IDOC_PROBLEM_CARTPOLE     cart-pole swing-up, n = 4, m = 1,
 *                             theta = [h, m_c, m_p, l, g, w_p, w_th, w_v, w_om, w_u, f_p, f_th, f_v, f_om];
 *   IDOC_PROBLEM_PENDULUM_AD  the pendulum again through the AD path (validates it against the hand-written one). */
#define IDOC_PROBLEM_CARTPOLE 3
#define IDOC_PROBLEM_PENDULUM_AD 4
/* User-defined families: csrc/families/<file>.cuh defines dynamics, cost and final cost once over a generic scalar
 * (reference: the Python callables of ilqr.Problem, idoc/ilqr.py:94-124, differentiated there by JAX and here by the device
 * forward-mode AD of csrc/ad.cuh) and registers itself with IDOC_FAMILY(name, Struct); the build assigns problem ids from
 * IDOC_PROBLEM_USER0 on, in file-name order.  Look families up by name: */
#define IDOC_PROBLEM_USER0 16
int idoc_ilqr_family_count(void);
/* name / problem id / n / m / theta_dim of the i-th user family (any output pointer may be NULL) */
int idoc_ilqr_family_info(int i, const char** name, int* problem, int* n, int* m, int* theta_dim);
const char* idoc_ilqr_last_error(void);
int idoc_ilqr_theta_dim(int problem, int n, int m);
size_t idoc_ilqr_ws_bytes(int dtype, int64_t B, int T, int n, int m, int systems);
size_t idoc_ilqr_vjp_ws_bytes(int dtype, int64_t B, int T, int n, int m, int systems);
int idoc_ilqr_solve(void* stream, int dtype, int problem, int64_t B, int T, int n, int m, int systems,
                    const void* theta, int theta_dim, const void* x0, void* X, void* U, void* Nu,
                    int maxiter, double thres, int use_line_search, double ls_lower, double ls_upper, double ls_rho,
                    int ls_maxiter, int32_t* iters, int32_t* ls_failed, void* ws, size_t ws_bytes, int allow_sync);
/* idoc_ilqr_linearize replaces make_lqr_approx(problem, params, local)(X, U) (idoc/ilqr.py:127-158): the LQR
 * leaves around a trajectory; mode 0 = local (deviation) form, 1 = global form (what ilqr's kkt feeds to lqr.kkt,
 * ilqr.py:192-194), 2 = global form with the dynamics curvature weighted by Nu added to Q, R, M. */
int idoc_ilqr_linearize(void* stream, int dtype, int problem, int64_t B, int T, int n, int m, int systems,
                        const void* theta, int theta_dim, const void* x0, const void* X, const void* U, const void* Nu,
                        int mode, void* Q, void* q, void* Qf, void* qf, void* M, void* R, void* r, void* A, void* Bm,
                        void* d);
int idoc_ilqr_vjp(void* stream, int dtype, int problem, int64_t B, int T, int n, int m, int systems,
                  const void* theta, int theta_dim, const void* x0, const void* X, const void* U, const void* Nu,
                  const void* Xb, const void* Ub, const void* Nub, void* gtheta, void* gx0, void* ws, size_t ws_bytes);

/* Control-limited iLQR (u_lo <= u_t <= u_hi, limits per problem [B, m]; single systems with a fused kernel):
 * every backward step solves its box QP (a clamp when m = 1, an active-set iteration otherwise) and a clamped control
 * carries no feedback (Tassa, Mansard, Todorov 2014); the implicit VJP treats the controls that sit on a bound as an
 * active set.  No reference counterpart: idoc has
 * no control limits (its "blqr" is shared-control BATCH LQR); BASELINE config 4 asks for them.  Parity unpinned. */
int idoc_ilqr_solve_box(void* stream, int dtype, int problem, int64_t B, int T, int n, int m, const void* theta,
                        int theta_dim, const void* x0, const void* u_lo, const void* u_hi, void* X, void* U, void* Nu,
                        int maxiter, double thres, int use_line_search, double ls_lower, double ls_upper, double ls_rho,
                        int ls_maxiter, int32_t* iters, int32_t* ls_failed, void* ws, size_t ws_bytes);
int idoc_ilqr_vjp_box(void* stream, int dtype, int problem, int64_t B, int T, int n, int m, const void* theta,
                      int theta_dim, const void* x0, const void* u_lo, const void* u_hi, const void* X, const void* U,
                      const void* Nu, const void* Xb, const void* Ub, const void* Nub, void* gtheta, void* gx0, void* ws,
                      size_t ws_bytes);

/* Shared-control batch iLQR (idoc/bilqr.py:120-221) as ONE kernel, with optional control limits on the shared control
 * (u_lo, u_hi [B, m], both null = no limits).  `systems` copies of one system (same theta, x0 [B, systems * n]) driven by one
 * control sequence: the kernels of idoc_ilqr_solve_box / idoc_ilqr_vjp_box on the block-diagonal lift (csrc/bilqr_fused.cu).
 * Compiled for the (family, n, m, systems) combinations idoc_bilqr_box_supported() answers 1 for; everything else returns
 * IDOC_ERR_UNSUPPORTED (unlimited problems then go through idoc_ilqr_solve with systems > 1).  Workspace sizes:
 * idoc_ilqr_ws_bytes / idoc_ilqr_vjp_ws_bytes with the same systems.  Limits: no reference counterpart, parity unpinned. */
int idoc_bilqr_box_supported(int problem, int n, int m, int systems);
int idoc_bilqr_solve_box(void* stream, int dtype, int problem, int64_t B, int T, int n, int m, int systems, const void* theta,
                         int theta_dim, const void* x0, const void* u_lo, const void* u_hi, void* X, void* U, void* Nu,
                         int maxiter, double thres, int use_line_search, double ls_lower, double ls_upper, double ls_rho,
                         int ls_maxiter, int32_t* iters, int32_t* ls_failed, void* ws, size_t ws_bytes);
int idoc_bilqr_vjp_box(void* stream, int dtype, int problem, int64_t B, int T, int n, int m, int systems, const void* theta,
                       int theta_dim, const void* x0, const void* u_lo, const void* u_hi, const void* X, const void* U,
                       const void* Nu, const void* Xb, const void* Ub, const void* Nub, void* gtheta, void* gx0, void* ws,
                       size_t ws_bytes);

/* ---- host pipeline: the hot path for a batch that lives in HOST memory ----------------------------
 * What a host program without device arrays calls (the reference's users hand NumPy / jax arrays to Solver.implicit and
 * jax.grad, idoc/lqr.py:167-180, tests/test_lqr.py:65-90).  Every pointer of idoc_lqr_host_solve_vjp is a HOST pointer
 * (page-locked memory, e.g. idoc_host_alloc, for the copies to overlap).  The batch is streamed through the device in chunks of
 * `chunk` problems over three streams (H2D of chunk c+1, kernels of chunk c, D2H of chunk c-1) in `nslots` >= 2 device slots:
 * fwd solve (X, U, Nu out) + implicit VJP with cotangent (X, U, 0), the loss 1/2|X|^2 + 1/2|U|^2 of tests/test_lqr.py:52-58.
 * reduce_batch = 0: g* have the shapes of their primals.  reduce_batch = 1: the ten LQR leaves are summed over the batch (shapes
 * without B; each chunk reduces on the device, the per-chunk sums are added on the host), gx0 stays per problem.
 * wait = 0 only enqueues (consecutive batches overlap); idoc_lqr_host_finish waits for everything enqueued and, in
 * batch-summed mode, completes the host-side sums (the g* of a wait = 0 call are valid after it).  Not thread-safe per handle. */
const char* idoc_lqr_host_last_error(void);
int idoc_lqr_host_create(void** handle, int dtype, int64_t chunk, int T, int n, int m, int nslots, int reduce_batch);
int idoc_lqr_host_solve_vjp(void* handle, int64_t B,
                            const void* Q, const void* q, const void* Qf, const void* qf, const void* M, const void* R,
                            const void* r, const void* A, const void* Bm, const void* d, const void* x0,
                            void* X, void* U, void* Nu,
                            void* gQ, void* gq, void* gQf, void* gqf, void* gM, void* gR, void* gr, void* gA, void* gB, void* gd,
                            void* gx0, int wait, float* ms);
int idoc_lqr_host_finish(void* handle);
int idoc_lqr_host_elapsed_ms(void* handle, float* ms);                 /* device time of the last enqueued batch (waits for it) */
int idoc_lqr_host_streams(void* handle, void** s_in, void** s_out);    /* for event timing across several batches */
int idoc_lqr_host_bytes(void* handle, uint64_t* h2d, uint64_t* d2h);   /* bytes copied by the last batch */
int idoc_lqr_host_destroy(void* handle);

/* ---- runtime helpers (so host bindings need nothing but this library) ---------------------------- */
int idoc_device_count(void);
int idoc_set_device(int dev);
int idoc_malloc(void** dptr, size_t bytes);
int idoc_free(void* dptr);
int idoc_host_alloc(void** hptr, size_t bytes); /* pinned */
int idoc_host_free(void* hptr);
int idoc_memset(void* dptr, int value, size_t bytes, void* stream);
int idoc_memcpy_h2d(void* dst, const void* src_host, size_t bytes, void* stream); /* async if stream != NULL */
int idoc_memcpy_d2h(void* dst_host, const void* src, size_t bytes, void* stream);
int idoc_memcpy_d2d(void* dst, const void* src, size_t bytes, void* stream);
int idoc_stream_create(void** stream);
int idoc_stream_destroy(void* stream);
int idoc_stream_sync(void* stream); /* NULL = device synchronize */
int idoc_event_create(void** ev);
int idoc_event_destroy(void* ev);
int idoc_event_record(void* ev, void* stream);
int idoc_stream_wait_event(void* stream, void* ev);
int idoc_event_sync(void* ev);
int idoc_event_elapsed_ms(void* ev_start, void* ev_stop, float* ms); /* synchronises on ev_stop */
int idoc_mem_info(size_t* free_bytes, size_t* total_bytes);
int idoc_device_props(int* sm_count, int* cc_major, int* cc_minor, char* name, int name_len);
/* measured FMA throughput of the current device in TFLOP/s (register-resident multiply-add chains, CUDA events; synchronises):
 * the roofline denominator of the FMA-bound time-invariant configuration */
int idoc_fma_peak(void* stream, int dtype, double* tflops);
/* number of kernel launches this library has issued in this process (bench evidence) */
int64_t idoc_launch_count(void);

/* per-kernel CUDA-event timing of the launches made between begin and end (ids: 0 riccati, 1 rollout,
 * 2 costate-fixup, 3 adjoint-backward, 4 adjoint-forward+grads, 5 kkt, 6 generator) */
int idoc_profile_begin(void);
int idoc_profile_end(int max_records, int* ids, float* ms, int* count);

/* ---- multi-GPU: the one exchange step of the path.  Gradients summed over the batch
 * (idoc_lqr_vjp with reduce_batch = 1) are all-reduced across the GPUs of a box with NCCL over NVLink.
 * libnccl.so.2 is dlopen'ed on first use.  id128 is an ncclUniqueId (128 bytes) created on one rank and
 * distributed by the host program (any channel: MPI, torch.distributed store, a file). */
int idoc_nccl_unique_id(void* id128);
int idoc_nccl_init(const void* id128, int rank, int nranks, void** comm);
int idoc_nccl_allreduce_sum(void* comm, void* stream, int dtype, void* buf, size_t count);
int idoc_nccl_destroy(void* comm);

/* synthetic problem generator on the device (SURVEY 8(d) distribution; counter-based RNG) */
int idoc_lqr_generate(void* stream, int dtype, int64_t B, int T, int n, int m, uint64_t seed,
                      void* Q, void* q, void* Qf, void* qf, void* M, void* R, void* r, void* A, void* Bm,
                      void* d, void* x0);

#ifdef __cplusplus
}
#endif
#endif /* IDOC_B200_H */
