// npde_grid_ops.cu -- the grid operators of utils/heat_utils.py as one fused
// kernel each: G (set_boundary / pad_boundary), Psi+G (fd_step, single step on
// the reference layout), the error norms (fd_error, l2_error) and the multigrid
// transfer operators (subsample, restriction+G, interpolation+G).
//
// These are HBM-bound elementwise / 5-point kernels: one thread per output
// cell, x fastest so every warp reads and writes whole 128-byte lines; the
// neighbour reuse is served by L1.  The k-step hot loop does not use them --
// see npde_stream.cu for the temporally blocked register-streaming kernels.
#include "npde_common.cuh"
#include "npde_internal.h"

namespace {

constexpr int TX = 32, TY = 8;

inline dim3 grid2d(int W, int H, int B) { return dim3((W + TX - 1) / TX, (H + TY - 1) / TY, B); }

// ---------------------------------------------------------------- G -------
__global__ void k_set_boundary_mask(float *x, BcView bc, int N) {
  int j = blockIdx.x * TX + threadIdx.x, i = blockIdx.y * TY + threadIdx.y, b = blockIdx.z;
  if (i >= N || j >= N) return;
  size_t o = ((size_t)b * N + i) * N + j;
  x[o] = npde_apply_G(x[o], bc, b, i, j, N);
}

// square: one thread per ring cell; rows skip the corner columns (columns win).
__global__ void k_set_boundary_square(float *x, BcView bc, int N) {
  int t = blockIdx.x * blockDim.x + threadIdx.x, b = blockIdx.y;
  if (t >= 4 * N) return;
  int side = t / N, p = t - side * N;
  const float *c = bc.bc + (size_t)b * 4;
  float *xb = x + (size_t)b * N * N;
  if (side == 0) { if (p > 0 && p < N - 1) xb[p] = __ldg(c + 0); }
  else if (side == 1) { if (p > 0 && p < N - 1) xb[(size_t)(N - 1) * N + p] = __ldg(c + 1); }
  else if (side == 2) xb[(size_t)p * N] = __ldg(c + 2);
  else xb[(size_t)p * N + N - 1] = __ldg(c + 3);
}

__global__ void k_pad_boundary(const float *__restrict__ xin, BcView bc, float *__restrict__ y, int N) {
  int j = blockIdx.x * TX + threadIdx.x, i = blockIdx.y * TY + threadIdx.y, b = blockIdx.z;
  if (i >= N || j >= N) return;
  int S = N - 2;
  bool interior = i >= 1 && i <= N - 2 && j >= 1 && j <= N - 2;
  float v = interior ? __ldg(xin + ((size_t)b * S + (i - 1)) * S + (j - 1)) : 0.0f;
  y[((size_t)b * N + i) * N + j] = npde_apply_G(v, bc, b, i, j, N);
}

// utils/heat_utils.py:70-94 initialize.  Square domain: the interior takes 0 / rnd / mean(bc) and the ring is left
// alone (in place, :85-90); mean = (((top + bottom) + left) + right) * 0.25, ATen's order for a row of four.
// Mask geometry: x <- G(0 | rnd) (:78-83).  rnd: (B,N-2,N-2) on the square domain, (B,N,N) on a mask geometry.
__global__ void k_initialize(float *x, BcView bc, int mode, const float *__restrict__ rnd, int N) {
  int j = blockIdx.x * TX + threadIdx.x, i = blockIdx.y * TY + threadIdx.y, b = blockIdx.z;
  if (i >= N || j >= N) return;
  const size_t o = ((size_t)b * N + i) * N + j;
  if (bc.kind == NPDE_BC_MASK) {
    const float v = mode == NPDE_INIT_RANDOM ? __ldg(rnd + o) : 0.0f;
    x[o] = npde_apply_G(v, bc, b, i, j, N);
    return;
  }
  if (i < 1 || i > N - 2 || j < 1 || j > N - 2) return;
  float v = 0.0f;
  if (mode == NPDE_INIT_RANDOM) {
    v = __ldg(rnd + ((size_t)b * (N - 2) + (i - 1)) * (N - 2) + (j - 1));
  } else if (mode == NPDE_INIT_AVG) {
    const float *c = bc.bc + (size_t)b * 4;
    v = __fmul_rn(__fadd_rn(__fadd_rn(__fadd_rn(__ldg(c), __ldg(c + 1)), __ldg(c + 2)), __ldg(c + 3)), 0.25f);
  }
  x[o] = v;
}

// ------------------------------------------------------------ Psi + G -----
__global__ void k_fd_step(const float *__restrict__ x, BcView bc, const float *__restrict__ f,
                          float *__restrict__ y, int N) {
  int j = blockIdx.x * TX + threadIdx.x, i = blockIdx.y * TY + threadIdx.y, b = blockIdx.z;
  if (i >= N || j >= N) return;
  const float *xb = x + (size_t)b * N * N;
  size_t o = (size_t)i * N + j;
  float u = i > 0 ? __ldg(xb + o - N) : 0.0f;
  float l = j > 0 ? __ldg(xb + o - 1) : 0.0f;
  float r = j < N - 1 ? __ldg(xb + o + 1) : 0.0f;
  float d = i < N - 1 ? __ldg(xb + o + N) : 0.0f;
  float v = npde_psi(u, l, r, d);
  if (f) v = __fsub_rn(v, __ldg(f + (size_t)b * N * N + o));
  y[(size_t)b * N * N + o] = npde_apply_G(v, bc, b, i, j, N);
}

__device__ __forceinline__ int view_col(int j, int perm) { return npde_view_col(j, perm); }

// -------------------------------------------------------- reductions ------
// grid (chunks, B); every CTA reduces a band of rows, writes one partial, and
// the last CTA of each problem folds the partials in index order, so the
// result is deterministic (no floating-point atomics).
struct ReduceWs {
  double *part;       // (B, chunks)
  unsigned *counter;  // (B), zeroed by the host wrapper
};

constexpr int RED_THREADS = 256;

template <bool IS_MAX>
__device__ __forceinline__ void block_reduce_finish(double v, ReduceWs ws, int chunks, float *err, int b,
                                                    double denom, bool take_sqrt) {
  __shared__ double sh[RED_THREADS / 32];
  __shared__ bool last;
  int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  if (IS_MAX) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmax(v, __shfl_xor_sync(NPDE_FULL_MASK, v, o));
  } else {
    v = npde_warp_sum(v);
  }
  if (lane == 0) sh[warp] = v;
  __syncthreads();
  if (threadIdx.x == 0) {
    double t = sh[0];
    for (int w = 1; w < RED_THREADS / 32; ++w) t = IS_MAX ? fmax(t, sh[w]) : t + sh[w];
    ws.part[(size_t)b * chunks + blockIdx.x] = t;
    __threadfence();
    unsigned done = atomicAdd(ws.counter + b, 1u);
    last = (done == (unsigned)chunks - 1);
  }
  __syncthreads();
  if (last && threadIdx.x == 0) {
    __threadfence();
    volatile double *p = ws.part + (size_t)b * chunks;
    double t = p[0];
    for (int c = 1; c < chunks; ++c) t = IS_MAX ? fmax(t, p[c]) : t + p[c];
    if (IS_MAX) err[b] = (float)t;
    else err[b] = take_sqrt ? sqrtf((float)(t / denom)) : (float)(t / denom);
    ws.counter[b] = 0u;  // leave the counters clean: back-to-back reductions on one workspace need no memset
  }
}

template <bool IS_MAX>
__global__ void __launch_bounds__(RED_THREADS)
k_fd_error(const float *__restrict__ x, int xp, long xbs, int xperm, BcView bc, const float *__restrict__ f,
           float *err, ReduceWs ws, int N, int rows_per_chunk) {
  int b = blockIdx.y, chunks = gridDim.x;
  const float *xb = x + (size_t)b * xbs;
  const float *fb = f ? f + (size_t)b * N * N : nullptr;
  const bool mask = bc.kind == NPDE_BC_MASK;
  const float *m = mask ? bc.bc + ((size_t)b * 2 + 1) * N * N : nullptr;
  int lo = mask ? 0 : 1, hi = mask ? N : N - 1;  // heat_utils.py:113-121
  int r0 = lo + blockIdx.x * rows_per_chunk, r1 = min(r0 + rows_per_chunk, hi);
  int W = hi - lo;
  // a warp walks along a row (coalesced, neighbours served by L1), the 8 warps of the CTA take
  // rows r0+w, r0+w+8, ...; max is accumulated in fp32 (exact, order independent), sums in fp64
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  float vmax = 0.0f;
  double vsum = 0.0;
  for (int i = r0 + ty; i < r1; i += RED_THREADS / 32) {
    const float *row = xb + (size_t)i * xp;
    for (int j = lo + tx; j < hi; j += 32) {
      size_t o = (size_t)i * N + j;
      const int jc = view_col(j, xperm);
      float u = i > 0 ? __ldg(row - xp + jc) : 0.0f;
      float l = j > 0 ? __ldg(row + view_col(j - 1, xperm)) : 0.0f;
      float c = __ldg(row + jc);
      float r = j < N - 1 ? __ldg(row + view_col(j + 1, xperm)) : 0.0f;
      float d = i < N - 1 ? __ldg(row + xp + jc) : 0.0f;
      float v = npde_residual(u, l, c, r, d);
      if (fb) v = __fsub_rn(v, __ldg(fb + o));
      if (m) v = __fmul_rn(v, __fsub_rn(1.0f, __ldg(m + o)));
      if (IS_MAX) vmax = fmaxf(vmax, fabsf(v));
      else vsum += (double)fabsf(v);
    }
  }
  double acc = IS_MAX ? (double)vmax : vsum;
  block_reduce_finish<IS_MAX>(acc, ws, chunks, err, b, (double)W * (double)W, false);
}

__global__ void __launch_bounds__(RED_THREADS)
k_l2_error(const float *__restrict__ x, int xp, long xbs, int xperm, const float *__restrict__ gt, float *err,
           ReduceWs ws, int N, int rows_per_chunk) {
  int b = blockIdx.y, chunks = gridDim.x;
  size_t n = (size_t)N * N;
  const int r0 = blockIdx.x * rows_per_chunk, r1 = min(r0 + rows_per_chunk, N);
  const float *xb = x + (size_t)b * xbs, *gb = gt + (size_t)b * n;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  double acc = 0.0;
  // a warp per row, lanes along it: the same cell -> thread mapping (and hence summation order) for every layout
  for (int i = r0 + warp; i < r1; i += RED_THREADS / 32) {
    const float *xr = xb + (size_t)i * xp, *gr = gb + (size_t)i * N;
    for (int j = lane; j < N; j += 32) {
      float d = __fsub_rn(__ldg(xr + view_col(j, xperm)), __ldg(gr + j));
      acc += (double)__fmul_rn(d, d);
    }
  }
  block_reduce_finish<false>(acc, ws, chunks, err, b, (double)n, true);
}

__global__ void k_copy_field(const float *__restrict__ src, int sp, long sb, int sperm, float *__restrict__ dst,
                             int dp, long db, int dperm, int N) {
  int j = blockIdx.x * TX + threadIdx.x, i = blockIdx.y * TY + threadIdx.y, b = blockIdx.z;
  if (i >= N || j >= N) return;
  dst[(size_t)b * db + (size_t)i * dp + view_col(j, dperm)] =
      __ldg(src + (size_t)b * sb + (size_t)i * sp + view_col(j, sperm));
}

// ------------------------------------------------------ transfer ops ------
// y[I,J] = scale * x[2I,2J]   (subsample; scale 4 gives f_sub, jacobi_iterator.py:47)
__global__ void k_subsample(const float *__restrict__ x, size_t x_bstride, float *__restrict__ y, int N, int M,
                            float scale) {
  int J = blockIdx.x * TX + threadIdx.x, I = blockIdx.y * TY + threadIdx.y, b = blockIdx.z;
  if (I >= M || J >= M) return;
  float v = __ldg(x + (size_t)b * x_bstride + (size_t)2 * I * N + 2 * J);
  y[((size_t)b * M + I) * M + J] = scale == 1.0f ? v : __fmul_rn(scale, v);
}

__global__ void k_restrict(const float *__restrict__ x, int xp, long xbs, int xperm, BcView bc_sub,
                           float *__restrict__ y, int yp, long ybs, int yperm, int N, int M) {
  int J = blockIdx.x * TX + threadIdx.x, I = blockIdx.y * TY + threadIdx.y, b = blockIdx.z;
  if (I >= M || J >= M) return;
  float v = 0.0f;
  if (I >= 1 && I <= M - 2 && J >= 1 && J <= M - 2) {
    const float *row = x + (size_t)b * xbs + (size_t)(2 * I) * xp;
    const int jc = view_col(2 * J, xperm);
    v = __fmul_rn(0.125f, __ldg(row - xp + jc));
    v = __fmaf_rn(0.125f, __ldg(row + view_col(2 * J - 1, xperm)), v);
    v = __fmaf_rn(0.5f, __ldg(row + jc), v);
    v = __fmaf_rn(0.125f, __ldg(row + view_col(2 * J + 1, xperm)), v);
    v = __fmaf_rn(0.125f, __ldg(row + xp + jc), v);
  }
  y[(size_t)b * ybs + (size_t)I * yp + view_col(J, yperm)] = npde_apply_G(v, bc_sub, b, I, J, M);
}

// utils/heat_utils.py:340-350 interpolation: bilinear x2, align_corners=True (copy / 2-point / 4-point
// means; see DESIGN.md for the two centre orders), then G with the fine bc.
// One thread per COARSE cell (I,J): it loads its 2x2 coarse neighbourhood once and writes the
// four fine cells (2I,2J), (2I,2J+1), (2I+1,2J), (2I+1,2J+1): no parity divergence, a quarter of the threads.
// Any field view on both sides (the V-cycle keeps its levels in the pair-permuted layout).
__global__ void k_prolong4(const float *__restrict__ x, int xp, long xbs, int xperm, BcView bc,
                           float *__restrict__ y, int yp, long ybs, int yperm, int M, int N, int nested) {
  int J = blockIdx.x * TX + threadIdx.x, I = blockIdx.y * TY + threadIdx.y, b = blockIdx.z;
  if (I >= M || J >= M) return;
  const float *xb = x + (size_t)b * xbs;
  const bool hasJ = J + 1 < M, hasI = I + 1 < M;
  const int c0 = view_col(J, xperm), c1 = hasJ ? view_col(J + 1, xperm) : c0;
  const float a = __ldg(xb + (size_t)I * xp + c0);
  const float bq = hasJ ? __ldg(xb + (size_t)I * xp + c1) : 0.0f;
  const float c = hasI ? __ldg(xb + (size_t)(I + 1) * xp + c0) : 0.0f;
  const float d = (hasI && hasJ) ? __ldg(xb + (size_t)(I + 1) * xp + c1) : 0.0f;
  float *yb = y + (size_t)b * ybs;
  const int i = 2 * I, j = 2 * J;
  yb[(size_t)i * yp + view_col(j, yperm)] = npde_apply_G(a, bc, b, i, j, N);
  if (hasJ)
    yb[(size_t)i * yp + view_col(j + 1, yperm)] = npde_apply_G(__fmul_rn(0.5f, __fadd_rn(a, bq)), bc, b, i, j + 1, N);
  if (hasI) {
    yb[(size_t)(i + 1) * yp + view_col(j, yperm)] = npde_apply_G(__fmul_rn(0.5f, __fadd_rn(a, c)), bc, b, i + 1, j, N);
    if (hasJ) {
      float sum = nested ? __fadd_rn(__fadd_rn(a, bq), __fadd_rn(c, d)) : __fadd_rn(__fadd_rn(__fadd_rn(a, bq), c), d);
      yb[(size_t)(i + 1) * yp + view_col(j + 1, yperm)] = npde_apply_G(__fmul_rn(0.25f, sum), bc, b, i + 1, j + 1, N);
    }
  }
}

}  // namespace

// ------------------------------------------------------------ host side ----
namespace npde {

int check_common(const void *a, const void *b, int B, int N) {
  if (!a || !b) return NPDE_ERR_NULL;
  if (B < 1 || N < 3) return NPDE_ERR_SIZE;
  return NPDE_OK;
}

int check_bc(const float *bc, int bc_kind) {
  if (bc_kind != NPDE_BC_SQUARE && bc_kind != NPDE_BC_MASK) return NPDE_ERR_BC;
  if (!bc) return NPDE_ERR_NULL;
  return NPDE_OK;
}

int reduce_chunks(int N) {
  long cells = (long)N * N;
  int c = (int)((cells + 32767) / 32768);
  return c < 1 ? 1 : (c > 1024 ? 1024 : c);
}

size_t reduce_ws_bytes(int B, int N) {
  return npde_align_up(sizeof(double) * (size_t)B * reduce_chunks(N), 256) +
         npde_align_up(sizeof(unsigned) * (size_t)B, 256);
}

static int reduce_ws_carve(void *ws, size_t ws_bytes, int B, int N, ReduceWs *out, cudaStream_t st,
                           bool counters_clean = false) {
  if (!ws || ((uintptr_t)ws & 255)) return NPDE_ERR_WORKSPACE;
  if (ws_bytes < reduce_ws_bytes(B, N)) return NPDE_ERR_WORKSPACE;
  size_t part_bytes = npde_align_up(sizeof(double) * (size_t)B * reduce_chunks(N), 256);
  out->part = reinterpret_cast<double *>(ws);
  out->counter = reinterpret_cast<unsigned *>(reinterpret_cast<char *>(ws) + part_bytes);
  if (counters_clean) return NPDE_OK;
  cudaError_t e = cudaMemsetAsync(out->counter, 0, sizeof(unsigned) * (size_t)B, st);
  return e == cudaSuccess ? NPDE_OK : (int)e;
}

int reduce_ws_clear(void *ws, size_t ws_bytes, int B, int N, cudaStream_t st) {
  ReduceWs r;
  return reduce_ws_carve(ws, ws_bytes, B, N, &r, st, false);
}

int set_boundary(float *x, const float *bc, int bc_kind, int B, int N, cudaStream_t st) {
  BcView v{bc, bc_kind};
  if (bc_kind == NPDE_BC_MASK) {
    NPDE_LAUNCH(k_set_boundary_mask, grid2d(N, N, B), dim3(TX, TY), 0, st, x, v, N);
  } else {
    NPDE_LAUNCH(k_set_boundary_square, dim3((4 * N + 127) / 128, B), dim3(128), 0, st, x, v, N);
  }
  NPDE_CHECK_LAUNCH();
  return NPDE_OK;
}

int initialize(float *x, const float *bc, int bc_kind, int mode, const float *rnd, int B, int N, cudaStream_t st) {
  BcView v{bc, bc_kind};
  NPDE_LAUNCH(k_initialize, grid2d(N, N, B), dim3(TX, TY), 0, st, x, v, mode, rnd, N);
  NPDE_CHECK_LAUNCH();
  return NPDE_OK;
}

int pad_boundary(const float *xin, const float *bc, int bc_kind, float *y, int B, int N, cudaStream_t st) {
  BcView v{bc, bc_kind};
  NPDE_LAUNCH(k_pad_boundary, grid2d(N, N, B), dim3(TX, TY), 0, st, xin, v, y, N);
  NPDE_CHECK_LAUNCH();
  return NPDE_OK;
}

int fd_step_simple(const float *x, const float *bc, int bc_kind, const float *f, float *y, int B, int N,
                   cudaStream_t st) {
  BcView v{bc, bc_kind};
  NPDE_LAUNCH(k_fd_step, grid2d(N, N, B), dim3(TX, TY), 0, st, x, v, f, y, N);
  NPDE_CHECK_LAUNCH();
  return NPDE_OK;
}

int fd_error(const float *x, const float *bc, int bc_kind, const float *f, int aggregate, float *err, int B,
             int N, void *ws, size_t ws_bytes, cudaStream_t st) {
  return fd_error_view(ref_view(x, N), bc, bc_kind, f, aggregate, err, B, N, ws, ws_bytes, st);
}

int fd_error_view(const FieldView &xv, const float *bc, int bc_kind, const float *f, int aggregate, float *err,
                  int B, int N, void *ws, size_t ws_bytes, cudaStream_t st, bool counters_clean) {
  if (bc_kind == NPDE_BC_SQUARE && aggregate == NPDE_AGG_MAX && N >= 3)  // the streaming kernel (npde_stream.cu)
    return fd_error_max_stream(xv, ref_view(f, N), err, B, N, st);
  ReduceWs r;
  int rc = reduce_ws_carve(ws, ws_bytes, B, N, &r, st, counters_clean);
  if (rc) return rc;
  BcView v{bc, bc_kind};
  int chunks = reduce_chunks(N);
  int rows = (bc_kind == NPDE_BC_MASK) ? N : N - 2;
  int rpc = (rows + chunks - 1) / chunks;
  chunks = (rows + rpc - 1) / rpc;  // no empty trailing chunks
  if (aggregate == NPDE_AGG_MAX) {
    NPDE_LAUNCH(k_fd_error<true>, dim3(chunks, B), dim3(RED_THREADS), 0, st, xv.p, xv.pitch, xv.bstride, xv.perm, v,
                f, err, r, N, rpc);
  } else {
    NPDE_LAUNCH(k_fd_error<false>, dim3(chunks, B), dim3(RED_THREADS), 0, st, xv.p, xv.pitch, xv.bstride, xv.perm, v,
                f, err, r, N, rpc);
  }
  NPDE_CHECK_LAUNCH();
  return NPDE_OK;
}

int l2_error(const float *x, const float *gt, float *err, int B, int N, void *ws, size_t ws_bytes,
             cudaStream_t st) {
  return l2_error_view(ref_view(x, N), gt, err, B, N, ws, ws_bytes, st);
}

int l2_error_view(const FieldView &xv, const float *gt, float *err, int B, int N, void *ws, size_t ws_bytes,
                  cudaStream_t st, bool counters_clean) {
  ReduceWs r;
  int rc = reduce_ws_carve(ws, ws_bytes, B, N, &r, st, counters_clean);
  if (rc) return rc;
  int chunks = reduce_chunks(N);
  int rpc = (N + chunks - 1) / chunks;
  chunks = (N + rpc - 1) / rpc;
  NPDE_LAUNCH(k_l2_error, dim3(chunks, B), dim3(RED_THREADS), 0, st, xv.p, xv.pitch, xv.bstride, xv.perm, gt, err, r,
              N, rpc);
  NPDE_CHECK_LAUNCH();
  return NPDE_OK;
}

int subsample_strided(const float *x, size_t x_bstride, float *y, int B, int N, float scale, cudaStream_t st) {
  int M = (N - 1) / 2 + 1;
  NPDE_LAUNCH(k_subsample, grid2d(M, M, B), dim3(TX, TY), 0, st, x, x_bstride, y, N, M, scale);
  NPDE_CHECK_LAUNCH();
  return NPDE_OK;
}

int subsample(const float *x, float *y, int B, int N, float scale, cudaStream_t st) {
  return subsample_strided(x, (size_t)N * N, y, B, N, scale, st);
}

int copy_field(const FieldView &src, const FieldViewMut &dst, int B, int N, cudaStream_t st) {
  NPDE_LAUNCH(k_copy_field, grid2d(N, N, B), dim3(TX, TY), 0, st, src.p, src.pitch, src.bstride, src.perm, dst.p,
              dst.pitch, dst.bstride, dst.perm, N);
  NPDE_CHECK_LAUNCH();
  return NPDE_OK;
}

int restrict_(const float *x, const float *bc_sub, int bc_kind, float *y, int B, int N, cudaStream_t st) {
  return restrict_view(ref_view(x, N), bc_sub, bc_kind, ref_view_mut(y, (N - 1) / 2 + 1), B, N, st);
}

int restrict_view(const FieldView &x, const float *bc_sub, int bc_kind, const FieldViewMut &y, int B, int N,
                  cudaStream_t st) {
  int M = (N - 1) / 2 + 1;
  BcView v{bc_sub, bc_kind};
  NPDE_LAUNCH(k_restrict, grid2d(M, M, B), dim3(TX, TY), 0, st, x.p, x.pitch, x.bstride, x.perm, v, y.p, y.pitch,
              y.bstride, y.perm, N, M);
  NPDE_CHECK_LAUNCH();
  return NPDE_OK;
}

int prolong(const float *x, const float *bc, int bc_kind, float *y, int B, int M, cudaStream_t st) {
  return prolong_view(ref_view(x, M), bc, bc_kind, ref_view_mut(y, 2 * M - 1), B, M, st);
}

int prolong_view(const FieldView &x, const float *bc, int bc_kind, const FieldViewMut &y, int B, int M,
                 cudaStream_t st) {
  int N = 2 * M - 1;
  BcView v{bc, bc_kind};
  int nested = ((long)M * M > 1024) ? 1 : 0;  // ATen CPU bilinear order switch, see DESIGN.md
  NPDE_LAUNCH(k_prolong4, grid2d(M, M, B), dim3(TX, TY), 0, st, x.p, x.pitch, x.bstride, x.perm, v, y.p, y.pitch,
              y.bstride, y.perm, M, N, nested);
  NPDE_CHECK_LAUNCH();
  return NPDE_OK;
}

}  // namespace npde
