// npde_common.cuh -- shared device helpers for libnpde_b200 (sm_100a).
//
// Arithmetic policy: every operation whose rounding matters is an explicit
// intrinsic (__fmaf_rn / __fadd_rn / __fsub_rn / __fmul_rn), which nvcc never
// contracts or reassociates.  The orders follow what the reference's ATen CPU
// path produces (row-major fmaf chains; see DESIGN.md "Arithmetic contract").
#pragma once
#ifdef NPDE_EMU
#include "cuda_emu.h"  // tests/emu: host-side SIMT emulator, never part of the product build
#define NPDE_LAUNCH(kernel, grid, block, smem, stream, ...) \
  (emu::trace(#kernel), emu::launch((grid), (block), (smem), [=]() { kernel(__VA_ARGS__); }))
#define NPDE_DYN_SMEM(type, name) type *name = reinterpret_cast<type *>(emu::S().dyn_smem)
#else
#include <cuda_runtime.h>
#define NPDE_LAUNCH(kernel, grid, block, smem, stream, ...) \
  (npde_note_launch(), kernel<<<(grid), (block), (smem), (stream)>>>(__VA_ARGS__))
#define NPDE_DYN_SMEM(type, name) \
  extern __shared__ __align__(16) unsigned char name##_raw[]; \
  type *name = reinterpret_cast<type *>(name##_raw)
#endif
#include <stddef.h>
#include <stdint.h>

#include "../../include/npde.h"

// statistics only (npde_debug_launch_count): number of kernels this library has enqueued
void npde_note_launch();

#define NPDE_FULL_MASK 0xffffffffu

// Boundary description handed to every kernel by value (lives in the constant bank).
struct BcView {
  const float *bc;  // square: (B,4); mask: (B,2,N,N)
  int kind;         // NPDE_BC_*
};

// Up to NPDE_MAX_CONV_LAYERS 3x3 kernels by value: FFMA reads them straight from c[0x0][..].
struct ConvW {
  float w[NPDE_MAX_CONV_LAYERS][9];
};

__device__ __forceinline__ float npde_act(float z, int act) {
  if (act == NPDE_ACT_CLAMP) return fminf(fmaxf(z, 0.0f), 1.0f);
  if (act == NPDE_ACT_SIGMOID) return 1.0f / (1.0f + expf(-z));
  return z;
}

__device__ __forceinline__ float npde_act_grad(float z, int act) {
  if (act == NPDE_ACT_CLAMP) return (z >= 0.0f && z <= 1.0f) ? 1.0f : 0.0f;
  if (act == NPDE_ACT_SIGMOID) {
    float s = 1.0f / (1.0f + expf(-z));
    return s * (1.0f - s);
  }
  return 1.0f;
}

// column j inside a row of a field view: the pair-permuted layout stores every aligned group of four
// columns (c0,c1,c2,c3) as (c0,c2,c1,c3) (npde_internal.h FieldView)
__device__ __forceinline__ int npde_view_col(int j, int perm) {
  return perm ? ((j & ~3) | ((j & 1) << 1) | ((j >> 1) & 1)) : j;
}

// An H x H array that may live inside a frame: element (i,j) at p[b*bstride + (i+off)*pitch + col(j+off)].
// A plain (B,H,H) array is {p, H, H*H, 0, 0}; the interior of a (B,N,N) frame is {p, pitch, bstride, perm, 1}.
struct ArrView {
  const float *p;
  int pitch;
  long bstride;
  int perm, off;
};
struct ArrViewMut {
  float *p;
  int pitch;
  long bstride;
  int perm, off;
};

// Jacobi 5-point average, taps in the oracle's order: up, left, right, down.
__device__ __forceinline__ float npde_psi(float u, float l, float r, float d) {
  float acc = __fmul_rn(0.25f, u);
  acc = __fmaf_rn(0.25f, l, acc);
  acc = __fmaf_rn(0.25f, r, acc);
  acc = __fmaf_rn(0.25f, d, acc);
  return acc;
}

// loss-kernel chain: up, left, centre(-1), right, down  (heat_utils.py:13-15).
__device__ __forceinline__ float npde_residual(float u, float l, float c, float r, float d) {
  float acc = __fmul_rn(0.25f, u);
  acc = __fmaf_rn(0.25f, l, acc);
  acc = __fmaf_rn(-1.0f, c, acc);
  acc = __fmaf_rn(0.25f, r, acc);
  acc = __fmaf_rn(0.25f, d, acc);
  return acc;
}

// G at one cell.  Square: ring cells take the bc scalars, columns win the
// corners (heat_utils.py:55-58).  Mask: x*(1-m)+v with two roundings (:53).
__device__ __forceinline__ float npde_apply_G(float v, const BcView &bc, int b, int i, int j, int N) {
  if (bc.kind == NPDE_BC_MASK) {
    const float *val = bc.bc + (size_t)b * 2 * N * N;
    const float *msk = val + (size_t)N * N;
    size_t o = (size_t)i * N + j;
    return __fadd_rn(__fmul_rn(v, __fsub_rn(1.0f, __ldg(msk + o))), __ldg(val + o));
  }
  const float *c = bc.bc + (size_t)b * 4;
  if (j == 0) return __ldg(c + 2);
  if (j == N - 1) return __ldg(c + 3);
  if (i == 0) return __ldg(c + 0);
  if (i == N - 1) return __ldg(c + 1);
  return v;
}

__device__ __forceinline__ float npde_warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(NPDE_FULL_MASK, v, o));
  return v;
}
__device__ __forceinline__ double npde_warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(NPDE_FULL_MASK, v, o);
  return v;
}

static inline size_t npde_align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

#define NPDE_CHECK_LAUNCH()                      \
  do {                                           \
    cudaError_t e__ = cudaGetLastError();        \
    if (e__ != cudaSuccess) return (int)e__;     \
  } while (0)
