// npde_cg.cu -- ConjugateGradient.forward (models/iterators/conjugate_gradient.py:15-47) as CUDA kernels.
//
// The reference runs n_iters CG iterations on L x = 0 (L = loss_kernel *, the 5-point operator of the hot path), with
// per-problem scalars alpha = r.r / p.Ap and beta = r'.r' / r.r obtained from torch.sum reductions (one host-visible
// tensor each).  Here one CG iteration is TWO kernels and the scalars never leave the device:
//
//   k_cg_init      r = -(L x) on the interior, y = x, per-block partial sums of r.r
//   k_cg_direction p' = r + beta p (beta from the partials; p' = r on the first iteration), Ap' = L p' with a
//                  zero ring (pad_boundary(p, 0)), partial sums of p'.Ap'
//   k_cg_update    alpha from the partials; y += alpha p' on ALL cells (the ring adds alpha*0, which is what
//                  propagates a NaN alpha exactly like the reference), r -= alpha Ap', partial sums of r.r
//
// Every reduction is deterministic: a block owns CG_ROWS rows of one problem, each warp sums its row lane-strided
// and by a shuffle tree, warps are added in index order, and the consumers add the per-block partials of their
// problem in a fixed order as well.  Element-wise arithmetic rounds like the reference's separate tensor ops
// (alpha*p then +; no contraction).  The sums themselves are fp32 like torch's but in a different association,
// so results agree with the reference to rounding (tests: <= 1e-5 relative), not bit for bit.
#include "npde_internal.h"

namespace npde {
namespace {

constexpr int CG_ROWS = 8;  // rows (= warps) per block

struct CgFields {
  float *y;          // (B,N,N) iterate
  float *r, *ap;     // (B,N,N), interior used
  float *p[2];       // (B,N,N) ping-pong search directions, ring = 0
  float *rr[2];      // (B,parts) partial sums of r.r, ping-pong over iterations
  float *pap;        // (B,parts) partial sums of p.Ap
};

__device__ __forceinline__ float warp_sum_f(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = __fadd_rn(v, __shfl_xor_sync(NPDE_FULL_MASK, v, o));
  return v;
}

// sum of this block's CG_ROWS row sums, in warp order; valid in thread (0,0)
__device__ __forceinline__ float block_sum(float row_sum, float *smem) {
  if (threadIdx.x == 0) smem[threadIdx.y] = row_sum;
  __syncthreads();
  float s = 0.0f;
  if (threadIdx.x == 0 && threadIdx.y == 0)
    for (int w = 0; w < CG_ROWS; ++w) s = __fadd_rn(s, smem[w]);
  __syncthreads();
  return s;
}

// every thread of the block gets the sum of `parts` partials of its problem (fixed order)
__device__ __forceinline__ float total_of(const float *partials, int parts, float *smem) {
  if (threadIdx.y == 0) {
    float s = 0.0f;
    for (int k = threadIdx.x; k < parts; k += 32) s = __fadd_rn(s, partials[k]);
    s = warp_sum_f(s);
    if (threadIdx.x == 0) smem[CG_ROWS] = s;
  }
  __syncthreads();
  float t = smem[CG_ROWS];
  __syncthreads();
  return t;
}

__device__ __forceinline__ float at(const float *f, int N, int i, int j) { return f[(size_t)i * N + j]; }

__global__ void __launch_bounds__(32 * CG_ROWS) k_cg_init(const float *__restrict__ x, CgFields c, int N, int parts) {
  __shared__ float smem[CG_ROWS + 1];
  const int b = blockIdx.y, i = blockIdx.x * CG_ROWS + threadIdx.y;
  const size_t base = (size_t)b * N * N;
  const float *xb = x + base;
  float acc = 0.0f;
  if (i < N) {
    const bool inner_row = i > 0 && i < N - 1;
    for (int j = threadIdx.x; j < N; j += 32) {
      size_t o = base + (size_t)i * N + j;
      c.y[o] = xb[(size_t)i * N + j];
      float r = 0.0f;
      if (inner_row && j > 0 && j < N - 1) {
        r = -npde_residual(at(xb, N, i - 1, j), at(xb, N, i, j - 1), at(xb, N, i, j), at(xb, N, i, j + 1),
                           at(xb, N, i + 1, j));
        acc = __fadd_rn(acc, __fmul_rn(r, r));
      }
      c.r[o] = r;
    }
  }
  float s = block_sum(warp_sum_f(acc), smem);
  if (threadIdx.x == 0 && threadIdx.y == 0) c.rr[0][(size_t)b * parts + blockIdx.x] = s;
}

// p' at (i,j): r + beta*p on the interior, 0 on the ring
__device__ __forceinline__ float dir_at(const float *r, const float *p, float beta, bool first, int N, int i, int j) {
  if (i <= 0 || j <= 0 || i >= N - 1 || j >= N - 1) return 0.0f;
  size_t o = (size_t)i * N + j;
  return first ? r[o] : __fadd_rn(r[o], __fmul_rn(beta, p[o]));
}

__global__ void __launch_bounds__(32 * CG_ROWS) k_cg_direction(CgFields c, int N, int parts, int it) {
  __shared__ float smem[CG_ROWS + 1];
  const int b = blockIdx.y, i = blockIdx.x * CG_ROWS + threadIdx.y;
  const bool first = it == 0;
  float beta = 0.0f;
  if (!first) {
    float rr_new = total_of(c.rr[it & 1] + (size_t)b * parts, parts, smem);
    float rr_old = total_of(c.rr[(it - 1) & 1] + (size_t)b * parts, parts, smem);
    beta = __fdiv_rn(rr_new, rr_old);
  }
  const size_t base = (size_t)b * N * N;
  const float *r = c.r + base, *p_old = c.p[(it + 1) & 1] + base;
  float *p_new = c.p[it & 1] + base, *ap = c.ap + base;
  float acc = 0.0f;
  if (i < N) {
    for (int j = threadIdx.x; j < N; j += 32) {
      float pc = dir_at(r, p_old, beta, first, N, i, j);
      p_new[(size_t)i * N + j] = pc;
      if (i > 0 && i < N - 1 && j > 0 && j < N - 1) {
        float a = npde_residual(dir_at(r, p_old, beta, first, N, i - 1, j), dir_at(r, p_old, beta, first, N, i, j - 1),
                                pc, dir_at(r, p_old, beta, first, N, i, j + 1),
                                dir_at(r, p_old, beta, first, N, i + 1, j));
        ap[(size_t)i * N + j] = a;
        acc = __fadd_rn(acc, __fmul_rn(pc, a));
      }
    }
  }
  float s = block_sum(warp_sum_f(acc), smem);
  if (threadIdx.x == 0 && threadIdx.y == 0) c.pap[(size_t)b * parts + blockIdx.x] = s;
}

__global__ void __launch_bounds__(32 * CG_ROWS) k_cg_update(CgFields c, int N, int parts, int it) {
  __shared__ float smem[CG_ROWS + 1];
  const int b = blockIdx.y, i = blockIdx.x * CG_ROWS + threadIdx.y;
  float rr = total_of(c.rr[it & 1] + (size_t)b * parts, parts, smem);
  float pap = total_of(c.pap + (size_t)b * parts, parts, smem);
  const float alpha = __fdiv_rn(rr, pap);
  const size_t base = (size_t)b * N * N;
  const float *p = c.p[it & 1] + base, *ap = c.ap + base;
  float *y = c.y + base, *r = c.r + base;
  float acc = 0.0f;
  if (i < N) {
    for (int j = threadIdx.x; j < N; j += 32) {
      size_t o = (size_t)i * N + j;
      y[o] = __fadd_rn(y[o], __fmul_rn(alpha, p[o]));
      if (i > 0 && i < N - 1 && j > 0 && j < N - 1) {
        float rn = __fsub_rn(r[o], __fmul_rn(alpha, ap[o]));
        r[o] = rn;
        acc = __fadd_rn(acc, __fmul_rn(rn, rn));
      }
    }
  }
  float s = block_sum(warp_sum_f(acc), smem);
  if (threadIdx.x == 0 && threadIdx.y == 0) c.rr[(it + 1) & 1][(size_t)b * parts + blockIdx.x] = s;
}

}  // namespace

static inline int cg_parts(int N) { return (N + CG_ROWS - 1) / CG_ROWS; }

size_t cg_ws_bytes(int B, int N) {
  size_t field = npde_align_up(sizeof(float) * (size_t)B * N * N, 256);
  size_t part = npde_align_up(sizeof(float) * (size_t)B * cg_parts(N), 256);
  return 4 * field + 3 * part;
}

int cg_step(const float *x, int n_iters, float *y, int B, int N, void *ws, cudaStream_t st) {
  char *base = reinterpret_cast<char *>(ws);
  size_t field = npde_align_up(sizeof(float) * (size_t)B * N * N, 256);
  size_t part = npde_align_up(sizeof(float) * (size_t)B * cg_parts(N), 256);
  CgFields c;
  c.y = y;
  c.r = reinterpret_cast<float *>(base);
  c.ap = reinterpret_cast<float *>(base + field);
  c.p[0] = reinterpret_cast<float *>(base + 2 * field);
  c.p[1] = reinterpret_cast<float *>(base + 3 * field);
  c.rr[0] = reinterpret_cast<float *>(base + 4 * field);
  c.rr[1] = reinterpret_cast<float *>(base + 4 * field + part);
  c.pap = reinterpret_cast<float *>(base + 4 * field + 2 * part);
  const int parts = cg_parts(N);
  dim3 grid(parts, B), block(32, CG_ROWS);
  NPDE_LAUNCH(k_cg_init, grid, block, 0, st, x, c, N, parts);
  NPDE_CHECK_LAUNCH();
  for (int it = 0; it < n_iters; ++it) {
    NPDE_LAUNCH(k_cg_direction, grid, block, 0, st, c, N, parts, it);
    NPDE_CHECK_LAUNCH();
    NPDE_LAUNCH(k_cg_update, grid, block, 0, st, c, N, parts, it);
    NPDE_CHECK_LAUNCH();
  }
  return NPDE_OK;
}

}  // namespace npde
