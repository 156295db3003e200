// npde_unet_coarse.cu -- the coarse levels of the U-Net shaped H (models/iterators/unet_iterator.py:52-85) in ONE
// kernel, one CTA per problem, every level tensor in shared memory.
//
// From a level whose interior is at most 127^2 downwards a problem's level tensors fit on chip: the kernel pools the
// skip tensor of the level above (stride-2 conv, read from global memory), runs the pre-smoothing convs, pooling convs,
// post-smoothing convs (+ skip) and pad -> bilinear x2 -> crop upsamplings of all deeper levels between shared-memory
// planes, and writes the upsampled correction of the level above back to global memory -- what unet_levels() does
// with 4 + 6 (L - l0) - 2 thread-per-cell launches and HBM round trips (12 launches, 149 + 127 us at 256 x 257^2 for
// U-Net(3,2,2); the finest level runs in the head / tail stream kernels either side).
//
// Arithmetic: per output the row-major fmaf chain of k_conv3x3 (first tap a plain product), then the level's
// foreground multiply, then the skip add; upsampling as k_pad_up_crop (incl. ATen's nested / flat order switch) -- so
// results are bit-identical to the layered kernels and the oracle.
//
// Planes are zero-bordered ((s+2)^2, the conv's zero padding and the upsampling's pad-by-one for free).  Level l keeps
// three planes A_l, B_l (ping-pong) and K_l (skip); the planes of level l+1 live inside A_l, which is dead while the
// deeper levels run (3 (s_{l+1}+2)^2 <= (s_l+2)^2 for s_l >= 5): 3 x 129^2 floats = 200 KB at 257^2.
#include "npde_common.cuh"
#include "npde_internal.h"

namespace {

constexpr int CT = 1024;  // threads per CTA: 32 x 32 cells per sweep

struct CoarseParams {
  ArrView src;       // skip tensor of level l0-1 (interior s_src x s_src), pooled into level l0
  int s_src;
  ArrViewMut up;     // upsampled correction of level l0-1 ((2 s_l0 + 1)^2 = s_src^2)
  const float *wf, *wp, *wsnd;  // device weights, indexed as in unet_levels(): wf[l*pre+j], wp[l], wsnd[(L-1-l)*post+j]
  const float *mask[NPDE_MAX_LEVELS];  // level masks ((s_l+2)^2 planes) or NULL
  size_t mbs[NPDE_MAX_LEVELS];
  int s[NPDE_MAX_LEVELS];
  int L, pre, post, l0;
};

// weights of one conv: `w` points into the shared-memory copy made at kernel start (a global load per conv phase
// would put a DRAM / L2 latency in front of each of the ~14 phases of a problem)
__device__ __forceinline__ void load_w(const float *w, float (&k)[9]) {
#pragma unroll
  for (int t = 0; t < 9; ++t) k[t] = w[t];
}
constexpr int CW_MAX = NPDE_MAX_LEVELS * (2 * NPDE_MAX_UNET_SMOOTH + 1) * 9;

// stride-1 conv (pad 1) between bordered planes of the same pitch P = s + 2: a thread owns one column of a chunk of
// rows and slides a 3x3 register window down it -- 3 shared-memory loads per output instead of 9, the row-major fmaf
// chain of k_conv3x3.  32 warps = 4 column groups x 8 row chunks.  Offsets are 32-bit element indices (the planes are
// shared memory: no 64-bit pointer arithmetic in the loop).
template <bool MASK, bool ADD>
__device__ __forceinline__ void conv_same(const float *__restrict__ in, float *__restrict__ out, int s, const float (&k)[9],
                                          const float *__restrict__ mask_plane, const float *__restrict__ add) {
  const int P = s + 2;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  const int j = tx + 32 * (ty & 3);
  const int rows = (s + 7) >> 3, i0 = (ty >> 2) * rows, i1 = min(i0 + rows, s);
  if (j >= s || i0 >= i1) return;
  int q = i0 * P + j;             // bordered (i0, j) = interior (i0 - 1, j - 1): top-left tap of output row i0
  int o = (i0 + 1) * P + (j + 1);  // output cell (and its mask / skip cell)
  float u0 = in[q], u1 = in[q + 1], u2 = in[q + 2];
  float m0 = in[q + P], m1 = in[q + P + 1], m2 = in[q + P + 2];
  q += 2 * P;
#pragma unroll 4
  for (int i = i0; i < i1; ++i) {
    const float d0 = in[q], d1 = in[q + 1], d2 = in[q + 2];
    float fg = 1.0f;
    if (MASK) fg = __fsub_rn(1.0f, __ldg(mask_plane + o));
    float acc = __fmul_rn(k[0], u0);
    acc = __fmaf_rn(k[1], u1, acc);
    acc = __fmaf_rn(k[2], u2, acc);
    acc = __fmaf_rn(k[3], m0, acc);
    acc = __fmaf_rn(k[4], m1, acc);
    acc = __fmaf_rn(k[5], m2, acc);
    acc = __fmaf_rn(k[6], d0, acc);
    acc = __fmaf_rn(k[7], d1, acc);
    acc = __fmaf_rn(k[8], d2, acc);
    if (MASK) acc = __fmul_rn(acc, fg);
    if (ADD) acc = __fadd_rn(acc, add[o]);
    out[o] = acc;
    q += P;
    o += P;
    u0 = m0; u1 = m1; u2 = m2;
    m0 = d0; m1 = d1; m2 = d2;
  }
}

// out (bordered, interior s_out^2) = conv3x3(in) [* fg] [+ add]; in: bordered plane of pitch s_in + 2
// stride 1: pad 1, s_in == s_out; stride 2: pad 0, s_out = (s_in - 3) / 2 + 1
__device__ void conv_plane(const float *in, int s_in, float *out, int s_out, const float *w, int stride,
                           const float *mask_plane, const float *add) {
  float k[9];
  load_w(w, k);
  if (stride == 1) {
    if (mask_plane) {
      if (add) conv_same<true, true>(in, out, s_out, k, mask_plane, add);
      else conv_same<true, false>(in, out, s_out, k, mask_plane, add);
    } else {
      if (add) conv_same<false, true>(in, out, s_out, k, mask_plane, add);
      else conv_same<false, false>(in, out, s_out, k, mask_plane, add);
    }
    return;
  }
  const int Pi = s_in + 2, Po = s_out + 2;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  for (int i = ty; i < s_out; i += 32)
    for (int j = tx; j < s_out; j += 32) {
      const float *q = in + (2 * i + 1) * Pi + (2 * j + 1);  // stride 2, pad 0: interior (2i, 2j)
      float acc = __fmul_rn(k[0], q[0]);
      acc = __fmaf_rn(k[1], q[1], acc);
      acc = __fmaf_rn(k[2], q[2], acc);
      acc = __fmaf_rn(k[3], q[Pi], acc);
      acc = __fmaf_rn(k[4], q[Pi + 1], acc);
      acc = __fmaf_rn(k[5], q[Pi + 2], acc);
      acc = __fmaf_rn(k[6], q[2 * Pi], acc);
      acc = __fmaf_rn(k[7], q[2 * Pi + 1], acc);
      acc = __fmaf_rn(k[8], q[2 * Pi + 2], acc);
      if (mask_plane) acc = __fmul_rn(acc, __fsub_rn(1.0f, __ldg(mask_plane + (i + 1) * Po + (j + 1))));
      if (add) acc = __fadd_rn(acc, add[(i + 1) * Po + (j + 1)]);
      out[(i + 1) * Po + (j + 1)] = acc;
    }
}

// the pooling conv of level l0-1 -> l0, input read from global memory through an array view.  `stage` (the B and K
// planes, unused at this point) gives every warp room for the three source rows of the output row it is working on:
// rows are fetched with coalesced loads (27 in flight per lane at 257^2) and the nine taps of an output come from
// shared memory, instead of nine dependent-latency gathers per output.
__device__ void pool_from_global(const float *pb, const ArrView &v, int s_in, float *out, int s_out, const float *w,
                                 const float *mask_plane, float *stage, int stage_floats) {
  float k[9];
  load_w(w, k);
  const int Po = s_out + 2;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  const int width = s_in + 2 * v.off + 3;  // floats of a source row that can be addressed (permuted groups of four)
  const int rowlen = min(v.pitch, (width + 3) & ~3);
  if (3 * rowlen * 32 <= stage_floats) {
    float *mine = stage + (size_t)ty * 3 * rowlen;
    for (int i = ty; i < s_out; i += 32) {
      __syncwarp();
#pragma unroll
      for (int a = 0; a < 3; ++a) {
        const float *src = pb + (size_t)(2 * i + a + v.off) * v.pitch;
        // (the pair permutation of a frame row is undone here, once per element: it is its own inverse)
        for (int c = tx; c < rowlen; c += 32) mine[a * rowlen + npde_view_col(c, v.perm)] = __ldg(src + c);
      }
      __syncwarp();
      for (int j = tx; j < s_out; j += 32) {
        const float *t0 = mine + 2 * j + v.off;
        float acc = 0.0f;
#pragma unroll
        for (int a = 0; a < 3; ++a)
#pragma unroll
          for (int c = 0; c < 3; ++c) {
            const float t = t0[a * rowlen + c];
            acc = (a == 0 && c == 0) ? __fmul_rn(k[0], t) : __fmaf_rn(k[a * 3 + c], t, acc);
          }
        if (mask_plane) acc = __fmul_rn(acc, __fsub_rn(1.0f, __ldg(mask_plane + (size_t)(i + 1) * Po + (j + 1))));
        out[(size_t)(i + 1) * Po + (j + 1)] = acc;
      }
    }
    return;
  }
  for (int i = ty; i < s_out; i += 32)
    for (int j = tx; j < s_out; j += 32) {
      const int i0 = 2 * i, j0 = 2 * j;  // stride 2, pad 0: always inside the s_in x s_in interior
      float t[9];
#pragma unroll
      for (int a = 0; a < 3; ++a)
#pragma unroll
        for (int c = 0; c < 3; ++c)
          t[a * 3 + c] = __ldg(pb + (size_t)(i0 + a + v.off) * v.pitch + npde_view_col(j0 + c + v.off, v.perm));
      float acc = __fmul_rn(k[0], t[0]);
#pragma unroll
      for (int q = 1; q < 9; ++q) acc = __fmaf_rn(k[q], t[q], acc);
      if (mask_plane) acc = __fmul_rn(acc, __fsub_rn(1.0f, __ldg(mask_plane + (size_t)(i + 1) * Po + (j + 1))));
      out[(size_t)(i + 1) * Po + (j + 1)] = acc;
    }
}

// crop_1(bilinear_x2(r with its zero border)) -- k_pad_up_crop.  r: bordered plane of interior s; output interior
// (2s+1)^2.  A thread takes one cell (I, J) of the bordered plane, 0 <= I, J <= s, and produces the up to four outputs
// whose upsampled coordinates are (2I + di, 2J + dj): one set of four loads, no divergent parity cases.
// store(i, j, v) receives the cropped coordinates.
template <class Store>
__device__ __forceinline__ void up_plane(const float *r, int s, int nested, Store store) {
  const int P = s + 2, so = 2 * s + 1;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  for (int I = ty; I <= s; I += 32)
    for (int J = tx; J <= s; J += 32) {
      const float *q = r + I * P + J;
      const float a = q[0], b = q[1], c = q[P], d = q[P + 1];
      const int i = 2 * I - 1, j = 2 * J - 1;  // cropped coordinates of the (even, even) output
      const bool i_ok = I >= 1, j_ok = J >= 1, i2_ok = i + 1 < so, j2_ok = j + 1 < so;
      if (i_ok && j_ok) store(i, j, a);
      if (i_ok && j2_ok) store(i, j + 1, __fmul_rn(0.5f, __fadd_rn(a, b)));
      if (i2_ok && j_ok) store(i + 1, j, __fmul_rn(0.5f, __fadd_rn(a, c)));
      if (i2_ok && j2_ok) {
        const float sum = nested ? __fadd_rn(__fadd_rn(a, b), __fadd_rn(c, d)) : __fadd_rn(__fadd_rn(__fadd_rn(a, b), c), d);
        store(i + 1, j + 1, __fmul_rn(0.25f, sum));
      }
    }
}

__device__ void clear(float *p, int n) {
  for (int t = threadIdx.x; t < n; t += CT) p[t] = 0.0f;
}
__device__ void clear_border(float *p, int s) {
  const int P = s + 2;
  for (int t = threadIdx.x; t < P; t += CT) {
    p[t] = 0.0f;
    p[(size_t)(P - 1) * P + t] = 0.0f;
    p[(size_t)t * P] = 0.0f;
    p[(size_t)t * P + P - 1] = 0.0f;
  }
}

__global__ void __launch_bounds__(CT, 1)
k_unet_coarse(const CoarseParams p) {
  NPDE_DYN_SMEM(float, sm);
  __shared__ float wsm[CW_MAX];
  const int b = blockIdx.x, L = p.L, pre = p.pre, post = p.post, l0 = p.l0;
  // all weights of the step: first | pool | second
  const int nf = 9 * L * pre, np = 9 * (L - 1), ns = 9 * L * post;
  for (int t = threadIdx.x; t < nf + np + ns; t += CT)
    wsm[t] = t < nf ? __ldg(p.wf + t) : (t < nf + np ? __ldg(p.wp + (t - nf)) : __ldg(p.wsnd + (t - nf - np)));
  const float *wf = wsm, *wp = wsm + nf, *wsnd = wsm + nf + np;
  // planes of level l: A = sm, B = sm + n_l, K = sm + 2 n_l (levels nest inside A)
  auto n_of = [&](int l) { return (p.s[l] + 2) * (p.s[l] + 2); };
  auto mk = [&](int l) -> const float * { return p.mask[l] ? p.mask[l] + (size_t)b * p.mbs[l] : nullptr; };

  // ---- down ----
  clear(sm, 3 * n_of(l0));
  __syncthreads();
  pool_from_global(p.src.p + (size_t)b * p.src.bstride, p.src, p.s_src, sm, p.s[l0], wp + 9 * (l0 - 1), mk(l0),
                   sm + n_of(l0), 2 * n_of(l0));
  __syncthreads();
  clear(sm + n_of(l0), 2 * n_of(l0));  // the staging rows lived in B and K
  __syncthreads();
  for (int l = l0; l < L; ++l) {
    const int n = n_of(l), s = p.s[l];
    float *A = sm, *B = sm + n, *K = sm + 2 * n;
    float *cur = A;
    for (int j = 0; j < pre; ++j) {
      float *dst = (j == pre - 1) ? K : (cur == A ? B : A);
      conv_plane(cur, s, dst, s, wf + 9 * (l * pre + j), 1, mk(l), nullptr);
      __syncthreads();
      cur = dst;
    }
    if (l < L - 1) {
      // the next level's planes live inside A (dead now): clear them, then pool K into the next A
      clear(sm, 3 * n_of(l + 1));
      __syncthreads();
      conv_plane(K, s, sm, p.s[l + 1], wp + 9 * l, 2, mk(l + 1), nullptr);
      __syncthreads();
    }
  }
  // ---- up ----
  const float *res = nullptr;  // result plane of the level below
  for (int l = L - 1; l >= l0; --l) {
    const int n = n_of(l), s = p.s[l];
    float *A = sm, *B = sm + n, *K = sm + 2 * n;
    float *cur;
    if (l == L - 1) {
      cur = K;
    } else {
      // upsample the result of level l+1 (inside A) into B, then A's border, overwritten by the deeper planes, is
      // zeroed again before A serves as a conv input
      const int nested = ((long)(p.s[l + 1] + 2) * (p.s[l + 1] + 2) > 1024) ? 1 : 0;
      const float *m = mk(l);
      const int P = s + 2;
      up_plane(res, p.s[l + 1], nested, [&](int i, int j, float v) {
        if (m) v = __fmul_rn(v, __fsub_rn(1.0f, __ldg(m + (i + 1) * P + (j + 1))));
        B[(i + 1) * P + (j + 1)] = v;
      });
      __syncthreads();
      clear_border(A, s);
      __syncthreads();
      cur = B;
    }
    const int iup = L - 1 - l;
    for (int j = 0; j < post; ++j) {
      float *dst = (cur == A) ? B : A;  // never K: the skip tensor is added by the last conv
      conv_plane(cur, s, dst, s, wsnd + 9 * (iup * post + j), 1, mk(l), j == post - 1 ? K : nullptr);
      __syncthreads();
      cur = dst;
    }
    res = cur;
  }
  // ---- the upsampled correction of level l0-1, to global memory ----
  {
    const int s = p.s[l0], so = 2 * s + 1;
    const int nested = ((long)(s + 2) * (s + 2) > 1024) ? 1 : 0;
    const float *m = p.mask[l0 - 1] ? p.mask[l0 - 1] + (size_t)b * p.mbs[l0 - 1] : nullptr;
    float *ob = p.up.p + (size_t)b * p.up.bstride;
    const ArrViewMut up = p.up;
    const bool frame = up.perm == 1 && up.off == 1 && (up.pitch & 3) == 0 && (up.bstride & 3) == 0 &&
                       ((((uintptr_t)up.p) & 15) == 0) && up.pitch >= ((so + 2 + 3) & ~3);
    if (frame) {
      // by pairs of coarse cells (I, 2m), (I, 2m+1): their 2 x 4 outputs are one aligned 16-byte group of each of
      // two frame rows (k_pad_up_crop_quad<true>) -- one 128-bit store per row instead of four 4-byte ones
      const int P = s + 2, tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
      for (int I = ty; I <= s; I += 32)
        for (int mq = tx; 2 * mq <= s; mq += 32) {
          const int J = 2 * mq;
          const float *q = res + I * P + J;
          float a[2][3];
#pragma unroll
          for (int di = 0; di < 2; ++di)
#pragma unroll
            for (int dj = 0; dj < 3; ++dj) a[di][dj] = (J + dj <= s + 1) ? q[di * P + dj] : 0.0f;
#pragma unroll
          for (int di = 0; di < 2; ++di) {
            const int i = 2 * I + di - 1;
            if (i < 0 || i >= so) continue;
            float v[4];
#pragma unroll
            for (int dj = 0; dj < 4; ++dj) {
              const int c = dj >> 1;
              const float p00 = a[0][c], p01 = a[0][c + 1], p10 = a[1][c], p11 = a[1][c + 1];
              if (di == 0 && !(dj & 1)) v[dj] = p00;
              else if (di == 0) v[dj] = __fmul_rn(0.5f, __fadd_rn(p00, p01));
              else if (!(dj & 1)) v[dj] = __fmul_rn(0.5f, __fadd_rn(p00, p10));
              else {
                const float sum = nested ? __fadd_rn(__fadd_rn(p00, p01), __fadd_rn(p10, p11))
                                         : __fadd_rn(__fadd_rn(__fadd_rn(p00, p01), p10), p11);
                v[dj] = __fmul_rn(0.25f, sum);
              }
              const int j = 2 * J + dj - 1;
              if (m && j >= 0 && j < so) v[dj] = __fmul_rn(v[dj], __fsub_rn(1.0f, __ldg(m + (i + 1) * (so + 2) + (j + 1))));
            }
            const int j0 = 2 * J - 1;
            float *g = ob + (size_t)(i + 1) * up.pitch + 4 * mq;
            if (j0 >= 0 && j0 + 3 < so) {
              *reinterpret_cast<float4 *>(g) = make_float4(v[0], v[2], v[1], v[3]);
            } else {
#pragma unroll
              for (int dj = 0; dj < 4; ++dj)
                if (j0 + dj >= 0 && j0 + dj < so) g[npde_view_col(dj, 1)] = v[dj];
            }
          }
        }
    } else {
      up_plane(res, s, nested, [&](int i, int j, float v) {
        if (m) v = __fmul_rn(v, __fsub_rn(1.0f, __ldg(m + (i + 1) * (so + 2) + (j + 1))));
        ob[(i + up.off) * up.pitch + npde_view_col(j + up.off, up.perm)] = v;
      });
    }
  }
}

size_t coarse_smem(const int *s, int l0) { return sizeof(float) * 3 * (size_t)(s[l0] + 2) * (s[l0] + 2); }

}  // namespace

namespace npde {

// levels l0 .. L-1 fit on chip, nest, and have the smoothing convs the kernel assumes
bool unet_coarse_supported(const int *s, int L, int l0, int pre, int post, int B) {
  if (l0 < 1 || l0 >= L || pre < 1 || post < 1 || pre > NPDE_MAX_UNET_SMOOTH || post > NPDE_MAX_UNET_SMOOTH ||
      L > NPDE_MAX_LEVELS)
    return false;
#ifdef NPDE_EMU
  if (s[l0] > 127) return false;
  (void)B;
#else
  // One CTA per problem, ~46 us whatever the batch: below ~16 problems the layered kernels, which spread a level over
  // all SMs and replay from a CUDA graph, are faster (B = 1, 257^2: 42 vs 58 us per U-Net step).
  if (B < 16) return false;
  if (coarse_smem(s, l0) + sizeof(float) * CW_MAX > 227 * 1024) return false;
#endif
  for (int l = l0; l + 1 < L; ++l)
    if (3 * (s[l + 1] + 2) * (s[l + 1] + 2) > (s[l] + 2) * (s[l] + 2)) return false;
  for (int l = l0; l < L; ++l)
    if (s[l] < 1) return false;
  return true;
}

int unet_coarse(const ArrView &src, int s_src, const ArrViewMut &up, const float *wf, const float *wp,
                const float *wsnd, float *const *mask, const size_t *mbs, const int *s, int L, int pre, int post, int l0,
                int B, cudaStream_t st) {
  CoarseParams p;
  p.src = src; p.s_src = s_src; p.up = up;
  p.wf = wf; p.wp = wp; p.wsnd = wsnd;
  for (int l = 0; l < NPDE_MAX_LEVELS; ++l) {
    p.mask[l] = l < L ? mask[l] : nullptr;
    p.mbs[l] = l < L ? mbs[l] : 0;
    p.s[l] = l < L ? s[l] : 0;
  }
  p.L = L; p.pre = pre; p.post = post; p.l0 = l0;
  const size_t smem = coarse_smem(s, l0);
#ifndef NPDE_EMU
  if (smem > 48 * 1024) {
    const cudaError_t e = cudaFuncSetAttribute(k_unet_coarse, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return (int)e;
  }
#endif
  NPDE_LAUNCH(k_unet_coarse, dim3((unsigned)B), dim3(CT), smem, st, p);
  NPDE_CHECK_LAUNCH();
  return NPDE_OK;
}

}  // namespace npde
