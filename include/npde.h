/*
 * npde.h -- C ABI of libnpde_b200.so: the learned-iterator hot path of
 * ermongroup/Neural-PDE-Solver as hand-written sm_100a CUDA kernels.
 *
 * The reference has no FFI: its boundary for this path is the Python factory
 * models/get_iterator.py:3-48 plus the nn.Module contract
 * models/iterators/iterator.py:6-44 and the free functions of
 * utils/heat_utils.py.  Every entry point below replaces one of those Python
 * call sites (cited per function, paths relative to the reference root); the
 * Python mirror in neural-pde-solver_b200/ binds them with ctypes, and
 * INTEGRATION.md shows the stub a reference maintainer would add.
 *
 * Conventions
 *  - all tensors fp32, C-contiguous, caller-owned DEVICE memory:
 *      fields x, f, gt, y : (B, N, N)
 *      bc, NPDE_BC_SQUARE : (B, 4) = [top, bottom, left, right]  (heat_utils.py:29-42)
 *      bc, NPDE_BC_MASK   : (B, 2, N, N) = [values, mask], mask in {0,1}
 *      H inputs/outputs   : (B, N-2, N-2)
 *      conv weights       : (L, 3, 3), cross-correlation, no bias (conv_iterator.py:12-15)
 *  - enqueue-only on `stream` (a cudaStream_t passed as void*); never
 *    synchronises, never allocates device memory: scratch comes from the
 *    caller (npde_workspace_bytes), must be 256-byte aligned.  (npde_iterate may
 *    create and destroy a private capture stream and an executable CUDA graph
 *    per call on launch-bound sizes; no state survives the call.)
 *  - returns NPDE_OK (0), a negative NPDE_ERR_* argument error, or a positive
 *    cudaError_t.  No global mutable state; thread-safe per stream.
 *  - arithmetic order is fixed (explicit fmaf chains) so results
 *    are bit-identical run to run and to oracle/npde_oracle.c.
 */
#ifndef NPDE_H_
#define NPDE_H_
#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

#define NPDE_OK 0
#define NPDE_ERR_NULL (-1)      /* required pointer is NULL */
#define NPDE_ERR_SIZE (-2)      /* B < 1, N < 3, or a multigrid level of even size */
#define NPDE_ERR_BC (-3)        /* bc_kind not one of NPDE_BC_*  (heat_utils.py:40-42 raises) */
#define NPDE_ERR_ACT (-4)       /* unknown activation (iterator.py:25-26 raises NotImplementedError) */
#define NPDE_ERR_LAYERS (-5)    /* n_layers / smoothing counts out of range */
#define NPDE_ERR_WORKSPACE (-6) /* workspace missing, misaligned or too small */
#define NPDE_ERR_ARG (-7)       /* other invalid argument */

#define NPDE_BC_SQUARE 0
#define NPDE_BC_MASK 1

#define NPDE_ACT_NONE 0    /* iterator.py:23-24 */
#define NPDE_ACT_CLAMP 1   /* iterator.py:21-22  x.clamp(0,1) -- the reference default (base_args.py:47) */
#define NPDE_ACT_SIGMOID 2 /* iterator.py:19-20 */

#define NPDE_AGG_MAX 0  /* heat_utils.py:126-127 */
#define NPDE_AGG_MEAN 1 /* heat_utils.py:128-129 */

#define NPDE_INIT_ZERO 0   /* heat_utils.py:80-82,86-87 */
#define NPDE_INIT_RANDOM 1 /* heat_utils.py:77-79,88-89: the uniform numbers come from the caller's generator */
#define NPDE_INIT_AVG 2    /* heat_utils.py:90-91 (square domain; ignored on a mask geometry like the reference) */

#define NPDE_IT_JACOBI 0    /* models/iterators/jacobi_iterator.py:7-20  */
#define NPDE_IT_CONV 1      /* models/iterators/conv_iterator.py:8-54    */
#define NPDE_IT_MULTIGRID 2 /* models/iterators/jacobi_iterator.py:23-78 */
#define NPDE_IT_UNET 3      /* models/iterators/unet_iterator.py:8-108   */

#define NPDE_MAX_CONV_LAYERS 8
#define NPDE_MAX_LEVELS 12
#define NPDE_SMALL_MAX_N 65 /* npde_iterate: up to this grid size the whole Jacobi / Conv k-step loop is one kernel */
#define NPDE_MAX_UNET_SMOOTH 8 /* npde_unet_step_bwd: pre / post smoothing convs per level (the reference uses <= 4) */

/* operations for npde_workspace_bytes */
#define NPDE_WS_REDUCE 0     /* npde_fd_error, npde_l2_error */
#define NPDE_WS_CONV_FWD 1   /* npde_conv_H, npde_conv_step_fwd */
#define NPDE_WS_CONV_BWD 2   /* npde_conv_step_bwd */
#define NPDE_WS_MG 3         /* npde_mg_step */
#define NPDE_WS_UNET_FWD 4   /* npde_unet_H, npde_unet_step_fwd */
#define NPDE_WS_UNET_BWD 5   /* npde_unet_step_bwd */
#define NPDE_WS_CG 6         /* npde_cg_step */
#define NPDE_WS_CONV_BPTT 7  /* npde_conv_iterate_bwd (npde_conv_iterate_tape uses NPDE_WS_CONV_FWD) */
#define NPDE_WS_ITERATE 16   /* npde_iterate: pass NPDE_WS_ITERATE + NPDE_IT_* of the iterator */

int npde_version(void);

/* Statistics: kernels enqueued by this library since it was loaded (bench.py's gpu_launches). */
unsigned long long npde_debug_launch_count(void);
const char *npde_strerror(int code);

/* Scratch bytes needed by operation `op` (NPDE_WS_*).  bc_kind, n_layers, pre,
 * post, k as in the call that will use it (pass 0 where not applicable). */
size_t npde_workspace_bytes(int op, int B, int N, int bc_kind, int n_layers, int pre, int post, int k);

/* G: utils/heat_utils.py:44-59 set_boundary (in place on x). */
int npde_set_boundary(float *x, const float *bc, int bc_kind, int B, int N, void *stream);

/* utils/heat_utils.py:70-94 initialize, in place on x (B,N,N).  Square domain: interior <- 0 | rnd | mean(bc), ring
 * untouched; mask geometry: x <- G(0 | rnd).  rnd: uniform numbers from the caller's generator, (B,N-2,N-2) on the
 * square domain and (B,N,N) on a mask geometry; NULL unless mode == NPDE_INIT_RANDOM. */
int npde_initialize(float *x, const float *bc, int bc_kind, int mode, const float *rnd, int B, int N, void *stream);

/* utils/heat_utils.py:61-68 pad_boundary: xin (B,N-2,N-2) -> y (B,N,N) = G(zero_pad(xin)). */
int npde_pad_boundary(const float *xin, const float *bc, int bc_kind, float *y, int B, int N, void *stream);

/* Psi then G: utils/heat_utils.py:96-106 fd_step == JacobiIterator.forward
 * (jacobi_iterator.py:12-17).  f may be NULL.  y must not alias x. */
int npde_fd_step(const float *x, const float *bc, int bc_kind, const float *f, float *y, int B, int N,
                 void *stream);

/* utils/heat_utils.py:108-132 fd_error -> err (B).  aggregate = NPDE_AGG_*. */
int npde_fd_error(const float *x, const float *bc, int bc_kind, const float *f, int aggregate, float *err,
                  int B, int N, void *ws, size_t ws_bytes, void *stream);

/* utils/heat_utils.py:134-144 l2_error -> err (B). */
int npde_l2_error(const float *x, const float *gt, float *err, int B, int N, void *ws, size_t ws_bytes,
                  void *stream);

/* models/iterators/conjugate_gradient.py:15-47 ConjugateGradient.forward: n_iters conjugate-gradient iterations
 * on the 5-point operator, boundary ring of x kept; x (B,N,N) -> y (B,N,N), y != x.  The reference takes no
 * boundary description here (bc is unused, :15) and no source (f must be None: `f + r` does not broadcast, :24-25).
 * Per-problem alpha / beta stay on the device (deterministic two-level fp32 sums). */
int npde_cg_step(const float *x, int n_iters, float *y, int B, int N, void *ws, size_t ws_bytes, void *stream);

/* models/iterators/conv_iterator.py:19-31 ConvIterator.H on r (B,N-2,N-2).
 * bc_kind == NPDE_BC_MASK <=> iterator.is_bc_mask (bc may be NULL otherwise). */
int npde_conv_H(const float *r, const float *bc, int bc_kind, const float *w, int n_layers, float *out, int B,
                int N, void *ws, size_t ws_bytes, void *stream);

/* models/iterators/conv_iterator.py:33-51 ConvIterator.forward:
 * y = G(pad(act(Psi(x) - f + H(Psi(x) - f - x)))).
 * w is the DEVICE copy of the weights (always required); w_host is an optional
 * HOST copy of the same (n_layers,3,3) values.  With w_host, the square domain
 * and 1 <= n_layers <= 4 the whole step is ONE fused kernel whose weights
 * travel as kernel parameters (constant bank); otherwise it runs layer by layer
 * from w.  The library never reads device memory on the host, hence the two
 * pointers. */
int npde_conv_step_fwd(const float *x, const float *bc, int bc_kind, const float *f, const float *w,
                       const float *w_host, int n_layers, int act, float *y, int B, int N, void *ws,
                       size_t ws_bytes, void *stream);

/* Backward of one ConvIterator.forward (autograd in the reference:
 * models/heat_model.py:52-53,71-72).  gout = dL/dy (B,N,N);
 * gw (n_layers,3,3) is OVERWRITTEN with the batch-summed weight gradient;
 * gx (B,N,N) may be NULL (the reference detaches its inputs, heat_model.py:48-50). */
int npde_conv_step_bwd(const float *x, const float *bc, int bc_kind, const float *f, const float *w,
                       int n_layers, int act, const float *gout, float *gw, float *gx, int B, int N, void *ws,
                       size_t ws_bytes, void *stream);

/* Backprop through k unrolled ConvIterator steps (BASELINE config 4; the reference gets it from autograd over a
 * Python loop of forward calls, models/heat_model.py:48-53 without the detach).
 * npde_conv_iterate_tape: tape (k,B,N,N) receives x_0 = x, x_1, ..., x_{k-1} and y = x_k  (k >= 1).
 * npde_conv_iterate_bwd:  gout = dL/dx_k; gw (n_layers,3,3) is OVERWRITTEN with the weight gradient summed over the
 *   k steps in fp32, last step first (the order autograd accumulates in); gx = dL/dx_0 may be NULL. */
int npde_conv_iterate_tape(const float *x, const float *bc, int bc_kind, const float *f, const float *w,
                           const float *w_host, int n_layers, int act, int k, float *tape, float *y, int B, int N,
                           void *ws, size_t ws_bytes, void *stream);
int npde_conv_iterate_bwd(const float *tape, const float *bc, int bc_kind, const float *f, const float *w,
                          int n_layers, int act, int k, const float *gout, float *gw, float *gx, int B, int N,
                          void *ws, size_t ws_bytes, void *stream);

/* utils/heat_utils.py:352-360 subsample: x (B,N,N) -> y (B,M,M), M=(N-1)/2+1 (exact injection). */
int npde_subsample(const float *x, float *y, int B, int N, void *stream);

/* utils/heat_utils.py:329-338 restriction: x (B,N,N) -> y (B,M,M), ring <- G(bc_sub) on the coarse grid. */
int npde_restrict(const float *x, const float *bc_sub, int bc_kind, float *y, int B, int N, void *stream);

/* utils/heat_utils.py:340-350 interpolation: x (B,M,M) -> y (B,2M-1,2M-1), then G with the fine bc. */
int npde_prolong(const float *x, const float *bc, int bc_kind, float *y, int B, int M, void *stream);

/* models/iterators/jacobi_iterator.py:34-75 MultigridIterator.forward (one V-cycle). */
int npde_mg_step(const float *x, const float *bc, int bc_kind, const float *f, int n_layers, int pre, int post,
                 float *y, int B, int N, void *ws, size_t ws_bytes, void *stream);

/* models/iterators/unet_iterator.py:36-85 UNetIterator.H on r (B,N-2,N-2).
 * w_first (n_layers*pre,3,3), w_pool (n_layers-1,3,3), w_second (n_layers*post,3,3). */
int npde_unet_H(const float *r, const float *bc, int bc_kind, const float *w_first, const float *w_pool,
                const float *w_second, int n_layers, int pre, int post, float *out, int B, int N, void *ws,
                size_t ws_bytes, void *stream);

/* models/iterators/unet_iterator.py:87-105 UNetIterator.forward. */
int npde_unet_step_fwd(const float *x, const float *bc, int bc_kind, const float *f, const float *w_first,
                       const float *w_pool, const float *w_second, int n_layers, int pre, int post, int act,
                       float *y, int B, int N, void *ws, size_t ws_bytes, void *stream);

/* Same with an optional HOST copy of the weights, w_host = first | pool | second concatenated as in
 * npde_iterate_desc.w.  With it (and 1 <= pre, post <= 4) the finest level runs as two fused streaming kernels
 * (Psi + residual + pre-smoothing convs; post-smoothing convs + skip + y + act + G); without it this is
 * npde_unet_step_fwd. */
int npde_unet_step_fwd_hw(const float *x, const float *bc, int bc_kind, const float *f, const float *w_first,
                          const float *w_pool, const float *w_second, const float *w_host, int n_layers, int pre,
                          int post, int act, float *y, int B, int N, void *ws, size_t ws_bytes, void *stream);

/* Backward of one UNetIterator.forward (autograd in the reference, models/heat_model.py:52-53,71-72 with
 * --iterator unet): recomputes the forward with a tape of every layer input, then walks it back.
 * gout = dL/dy (B,N,N); gradients OVERWRITE gw_first/gw_pool/gw_second; gx (B,N,N) may be NULL.
 * pre, post <= NPDE_MAX_UNET_SMOOTH. */
int npde_unet_step_bwd(const float *x, const float *bc, int bc_kind, const float *f, const float *w_first,
                       const float *w_pool, const float *w_second, int n_layers, int pre, int post, int act,
                       const float *gout, float *gw_first, float *gw_pool, float *gw_second, float *gx, int B,
                       int N, void *ws, size_t ws_bytes, void *stream);

/* k applications of one iterator without returning to the host: replaces the
 * Python loops utils/heat_utils.py:170-178 (calculate_errors),
 * models/heat_model.py:49-50,61-62 (train), runtime.py:78-79, generation.py:86-93.
 * Jacobi and Conv run as temporally blocked register-streaming kernels. */
typedef struct npde_iterate_desc {
  int iterator;        /* NPDE_IT_* */
  int B, N;
  int bc_kind;
  int n_layers;        /* conv layers, or multigrid / unet depth */
  int pre, post;       /* multigrid / unet smoothing counts */
  int act;             /* NPDE_ACT_* (conv, unet) */
  int k;               /* number of applications, >= 0 */
  int temporal_depth;  /* 0 = library default: grids up to NPDE_SMALL_MAX_N run the whole Jacobi / Conv loop in one
                        * on-chip kernel, larger ones stream with 4 Jacobi steps per HBM round trip; 1,2,4,8 = streaming
                        * kernels with that many Jacobi steps per round trip (Conv: one step per launch, first-generation
                        * strip kernel); -1 = Conv on the square domain: always the wide kernel on the batch-interleaved
                        * layout (npde_pack), which the default picks by itself for large enough problems */
  const float *x_in;   /* (B,N,N) */
  float *x_out;        /* (B,N,N); may alias x_in */
  const float *bc;
  const float *f;      /* NULL: Laplace */
  const float *w;      /* DEVICE weights. conv: (n_layers,3,3); unet: first | pool | second concatenated */
  const float *w_host; /* optional HOST copy of w (enables the fused conv kernel, see npde_conv_step_fwd) */
  const float *gt;     /* NULL, or ground truth for err_l2 */
  float *err_l2;       /* NULL, or (k,B): row t = l2_error(x_{t+1}, gt)  (heat_utils.py:172, unnormalised) */
  float *err_fd;       /* NULL, or (k+1,B): row t = fd_error(x_t, 'max'), t = 0..k (generation.py:88) */
  void *workspace;
  size_t workspace_bytes;
} npde_iterate_desc;

int npde_iterate(const npde_iterate_desc *d, void *stream);

/* generation.py:72-99 get_solution as ONE device-side loop: apply the iterator of `d` (d->k, err_l2, err_fd and
 * temporal_depth are ignored) until max_b fd_error(x_b, 'max') < tol, looking at the residual every `check_every`
 * applications like the reference's `(i + 1) % 100 == 0`, at most max_iters applications (rounded up to a multiple of
 * check_every).  The loop is a CUDA graph with a WHILE node whose condition the residual kernel writes on the device:
 * the call only enqueues, the host never waits.  result: 16 bytes of DEVICE memory, npde_solve_result.
 * Workspace: npde_workspace_bytes(NPDE_WS_ITERATE + iterator, ..., k = check_every). */
typedef struct npde_solve_result {
  int iterations;  /* applications performed (0 if x_in already met the tolerance) */
  int converged;   /* 1 iff the last residual checked is < tol */
  float residual;  /* the last residual checked: max_b fd_error */
  float residual0; /* max_b fd_error of x_in */
} npde_solve_result;
int npde_solve(const npde_iterate_desc *d, float tol, int check_every, int max_iters, void *result, void *stream);

/* ---- persistent device layout for drivers that keep fields resident between calls ----------------------
 * The loops of generation.py:86-93 / runtime.py:78-79 / heat_model.py:49-50 call the iterator on the SAME
 * tensors over and over.  A driver may convert them once into the library's batch-interleaved "W8" layout
 * and step there, so that no call pays a layout pass:
 *   element (b,i,c) of a (B,N,N) field  ->  i*pitch + (b*G8 + c/8)*8 + p8(c%8),
 *   G8 = ceil(N/8), pitch = 8*B*G8 floats, p8 = position of column c0..c7 in (c0,c4,c1,c5,c2,c6,c3,c7);
 *   padding columns hold zeros.  Buffers: npde_packed_floats(B,N) floats, 32-byte aligned. */
size_t npde_packed_floats(int B, int N);
int npde_pack(const float *x, float *packed, int B, int N, void *stream);    /* (B,N,N) -> W8 */
int npde_unpack(const float *packed, float *x, int B, int N, void *stream);  /* W8 -> (B,N,N) */
/* ConvIterator.forward (models/iterators/conv_iterator.py:33-51) on packed fields, square domain (bc (B,4)),
 * 1 <= n_layers <= 4, w_host = HOST weights (n_layers,3,3); f_packed: packed source or NULL; in != out. */
int npde_conv_step_packed(const float *in_packed, float *out_packed, const float *bc, const float *f_packed,
                          const float *w_host, int n_layers, int act, int B, int N, void *stream);
/* k applications of Phi_H ping-ponging between two W8 fields: application t reads (t even ? a : b) and writes the
 * other one, so the k-th iterate is in (k even ? a : b).  The k-step loop of heat_utils.py:163-181 / generation.py:
 * 86-93 for a driver that keeps its iterate in the packed layout between residual checks (no layout pass per call). */
int npde_conv_iterate_packed(float *a_packed, float *b_packed, const float *bc, const float *f_packed,
                             const float *w_host, int n_layers, int act, int B, int N, int k, void *stream);
/* fd_error(x, bc, f, aggregate='max') (utils/heat_utils.py:108-121, square domain) of a W8 field -> err[B];
 * bit-identical to npde_fd_error on the unpacked field. */
int npde_fd_error_packed(const float *packed, const float *f_packed, float *err, int B, int N, void *stream);

#ifdef __cplusplus
}
#endif
#endif /* NPDE_H_ */
