// npde_bwd_stream.cu -- backward of one Phi_H step (ConvIterator.forward, models/iterators/conv_iterator.py:33-51,
// differentiated by autograd in models/heat_model.py:52-53,71-72) on the square domain as ONE register-streaming pass.
//
// Given the step's input x and the gradient g' of its output, the kernel recomputes the forward intermediates row by
// row (y = Psi(x) - f, r_0 = y - x, r_k = conv(r_{k-1}, w_k), z = y + r_L; the oracle's rounding order, so act'(z)
// sees the forward's z bit for bit) and runs the adjoint right behind them in the same stream of rows:
//     g_L = g' * act'(z)                               (interior cells; G makes the ring independent of x)
//     dW_l += sum_cells g_l (x) r_{l-1},   g_{l-1} = conv^T(g_l, w_l)            l = L .. 1
//     g_y = g_L + g_0,   g_x = Psi^T(g_y) - g_0
// A warp owns a strip of 32 x 4 columns of the batch's VIRTUAL row (every image padded to a multiple of four columns,
// images side by side, so strips span images: no per-image remainder strips at 257^2) and a band of rows; two lanes
// on either side are halo (forward radius L+1 plus adjoint radius L+1 <= 8 columns), the band is entered 2L+2 rows
// early.  Rows that a later stage needs again (r_{l-1} for dW_l, y for z, g_L for g_y) wait in LANE-PRIVATE shared
// memory delay lines (no synchronisation); neighbours' columns come from warp shuffles.  The 9L weight-gradient sums
// are accumulated per lane in fp32 over the few hundred cells a lane owns, then in fp64 across lanes and CTAs in a
// fixed order (k_bwd_finish): deterministic, and within 1e-6 of autograd's own (unspecified-order) fp32 sums.
//
// Replaces, per step, the layered chain k_psi_residual + L x k_conv3x3 + k_bwd_gz + L x k_bwd_layer + k_bwd_gx
// (npde_conv_ops.cu; still used for mask geometries and more than three layers).
#include <stdlib.h>

#include "npde_common.cuh"
#include "npde_internal.h"

namespace {

constexpr int BH = 2;            // halo lanes per side
constexpr int BV = 32 - 2 * BH;  // lanes whose columns the strip owns
constexpr int BROW = 24;         // bytes of one delayed r row of a lane: 6 floats (own 4 + both neighbours); 64-bit accesses
                                 // at this stride are bank-conflict free
constexpr int BROW4 = 16;        // bytes of one delayed y / g_L row of a lane

struct BwdParams {
  const float *x, *f, *gout;
  float *gx;
  double *part;  // [ctas][9 * L] weight-gradient partials
  int B, N, Gp, groups;  // Gp = ceil(N / 4) groups of four columns per image, groups = B * Gp
  int strips, bands, band_rows;
  const float *w;  // (L, 9) weights in device memory
};

// Programmatic dependent launch (as in npde_wide.cu): a grid may become resident while its predecessor in the stream
// drains; it touches global memory only after griddepcontrol.wait.  A backward pass is a chain
// k_bwd_stream -> k_bwd_finish -> k_bwd_stream -> ...: every link waits for its predecessor, which itself completed
// only after ITS predecessor, so step t+1 sees both the gradient field and the partial sums of step t.
#ifdef NPDE_EMU
__device__ __forceinline__ void bw_pdl_sync() {}
#else
__device__ __forceinline__ void bw_pdl_sync() {
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
  asm volatile("griddepcontrol.wait;" ::: "memory");
}
#endif

struct Row6 {
  float v[6];  // v[0] = column -1 (left lane's last), v[1..4] = own columns, v[5] = column 4 (right lane's first)
};

__device__ __forceinline__ void bw_halo(Row6 &r, int lane_l, int lane_r) {
  r.v[0] = __shfl_sync(NPDE_FULL_MASK, r.v[4], lane_l);
  r.v[5] = __shfl_sync(NPDE_FULL_MASK, r.v[1], lane_r);
}
__device__ __forceinline__ void bw_zero(Row6 &r) {
#pragma unroll
  for (int c = 0; c < 6; ++c) r.v[c] = 0.0f;
}
__device__ __forceinline__ void bw_store(char *slot, const Row6 &r) {
  float2 *p = reinterpret_cast<float2 *>(slot);
#pragma unroll
  for (int c = 0; c < 3; ++c) p[c] = make_float2(r.v[2 * c], r.v[2 * c + 1]);
}
__device__ __forceinline__ void bw_load(const char *slot, Row6 &r) {
  const float2 *p = reinterpret_cast<const float2 *>(slot);
#pragma unroll
  for (int c = 0; c < 3; ++c) {
    const float2 v = p[c];
    r.v[2 * c] = v.x;
    r.v[2 * c + 1] = v.y;
  }
}
__device__ __forceinline__ void bw_store4(char *slot, const float (&v)[4]) {
  *reinterpret_cast<float4 *>(slot) = make_float4(v[0], v[1], v[2], v[3]);
}
__device__ __forceinline__ void bw_load4(const char *slot, float (&v)[4]) {
  const float4 q = *reinterpret_cast<const float4 *>(slot);
  v[0] = q.x; v[1] = q.y; v[2] = q.z; v[3] = q.w;
}

template <int ACT>
__device__ __forceinline__ float bw_act_grad(float z) {
  if (ACT == NPDE_ACT_CLAMP) return (z >= 0.0f && z <= 1.0f) ? 1.0f : 0.0f;
  if (ACT == NPDE_ACT_SIGMOID) {
    const float s = 1.0f / (1.0f + expf(-z));
    return s * (1.0f - s);
  }
  return 1.0f;
}

// Slots of the lane-private delay lines (rows).  With x row `it` arriving in a step:
//   y row it-1 is written, y row it-1-L read                         -> L + 1 slots
//   g_L row it-1-L written, g_L row it-1-2L read                     -> L + 1 slots
//   r_{l-1} row it-l written (l = 1..L), row it-2L-2+l read          -> 2(L-l) + 3 slots
template <int L>
struct BwdLayout {
  static constexpr int y_slots = L + 1, g_slots = L + 1;
  static constexpr int r_total = (L - 1) * (L + 3) + 3;  // sum over l of r_slots(l)
  static constexpr int rows = r_total + y_slots + g_slots;
  static constexpr size_t r_bytes = (size_t)r_total * BROW * 32;
  static constexpr size_t bytes = r_bytes + (size_t)(y_slots + g_slots) * BROW4 * 32;
};
__host__ __device__ __forceinline__ int bw_r_slots(int L, int l) { return 2 * (L - l) + 3; }
__host__ __device__ __forceinline__ int bw_r_base(int L, int l) { return (l - 1) * (2 * L + 3 - l); }

template <int L, int ACT, bool HASF>
__global__ void __launch_bounds__(32, 12)
k_bwd_stream(const BwdParams p) {
  NPDE_DYN_SMEM(char, smem);
  using LY = BwdLayout<L>;
  bw_pdl_sync();
  const int lane = threadIdx.x & 31, N = p.N;
  const int lane_l = (lane + 31) & 31, lane_r = (lane + 1) & 31;
  const int strip = (int)blockIdx.x % p.strips, band = (int)blockIdx.x / p.strips;
  const int i0 = band * p.band_rows, i1 = min(i0 + p.band_rows, N);
  float acc[L][9], wr[L][9];
#pragma unroll
  for (int l = 0; l < L; ++l)
#pragma unroll
    for (int t = 0; t < 9; ++t) {
      acc[l][t] = 0.0f;
      wr[l][t] = __ldg(p.w + l * 9 + t);  // the C ABI of the backward passes the weights in device memory only
    }

  if (i0 < i1) {
    // this lane's group of four columns in the virtual row
    const int G = strip * BV - BH + lane;
    const bool in_range = G >= 0 && G < p.groups;
    const int b = in_range ? G / p.Gp : 0;
    const int col0 = in_range ? (G - b * p.Gp) * 4 : 0;
    const bool owner = in_range && lane >= BH && lane < BH + BV;
    bool exists[4], interior[4];
#pragma unroll
    for (int c = 0; c < 4; ++c) {
      exists[c] = in_range && col0 + c < N;
      interior[c] = in_range && col0 + c >= 1 && col0 + c <= N - 2;
    }
    const size_t img = (size_t)b * N * N + col0;
    const float *px = p.x + img, *pg = p.gout + img, *pf = HASF ? p.f + img : nullptr;
    float *po = p.gx ? p.gx + img : nullptr;
    char *mine = smem + (size_t)lane * BROW;  // r row slot s of this lane: mine + s * SLOT
    char *mine4 = smem + LY::r_bytes + (size_t)lane * BROW4;  // y slots, then g_L slots: mine4 + s * SLOT4
    constexpr size_t SLOT = 32 * BROW, SLOT4 = 32 * BROW4;
    for (int s = 0; s < LY::r_total; ++s) {
      Row6 z;
      bw_zero(z);
      bw_store(mine + s * SLOT, z);
    }
    for (int s = 0; s < LY::y_slots + LY::g_slots; ++s) {
      const float z4[4] = {0.0f, 0.0f, 0.0f, 0.0f};
      bw_store4(mine4 + s * SLOT4, z4);
    }
    // `rowp` points at this lane's first column of `row` (it is only dereferenced when the row exists)
    auto load4 = [&](const float *rowp, int row, float (&v)[4]) {
      const bool ok = row >= 0 && row < N;
#pragma unroll
      for (int c = 0; c < 4; ++c) v[c] = (ok && exists[c]) ? __ldg(rowp + c) : 0.0f;
    };

    // register state
    Row6 xw[3];          // x rows it-2, it-1, it
    float S[L][2][4];    // forward conv stage k: partial sums of the two unfinished output rows
    Row6 gw[L][3];       // gw[l-1]: rows n-2, n-1, n of g_l (n = the newest row, it-2L-1+l)
    Row6 gyw[3];         // g_y rows n0-2, n0-1, n0 (n0 = it-2L-1)
    float g0prev[4];     // g_0 row n0-1
#pragma unroll
    for (int k = 0; k < 3; ++k) {
      bw_zero(xw[k]);
      bw_zero(gyw[k]);
    }
#pragma unroll
    for (int l = 0; l < L; ++l) {
#pragma unroll
      for (int k = 0; k < 3; ++k) bw_zero(gw[l][k]);
#pragma unroll
      for (int c = 0; c < 4; ++c) S[l][0][c] = S[l][1][c] = 0.0f;
    }
#pragma unroll
    for (int c = 0; c < 4; ++c) g0prev[c] = 0.0f;

    const int it_first = max(i0 - 2 * L - 2, 0), it_last = i1 + 2 * L + 1;
    float xnext[4], fnext[4], gnext[4];  // x row it, f row it-1, g' row it-1-L: loaded one step ahead
    // running row pointers of the three input streams and of the output (pointer arithmetic only: no dereference
    // outside [0, N))
    const float *rx = px + (ptrdiff_t)it_first * N;
    const float *rf = HASF ? pf + (ptrdiff_t)(it_first - 1) * N : nullptr;
    const float *rg = pg + (ptrdiff_t)(it_first - 1 - L) * N;
    ptrdiff_t ro = (ptrdiff_t)(it_first - 2 * L - 2) * N;  // offset of the output row of this step from po
    load4(rx, it_first, xnext);
    if (HASF) load4(rf, it_first - 1, fnext);
    load4(rg, it_first - 1 - L, gnext);

    for (int it = it_first; it <= it_last; ++it) {
      // ---- inputs of this step; start the loads of the next one ----
#pragma unroll
      for (int k = 0; k < 2; ++k) xw[k] = xw[k + 1];
#pragma unroll
      for (int c = 0; c < 4; ++c) xw[2].v[c + 1] = xnext[c];
      float fcur[4], gcur[4];
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        fcur[c] = HASF ? fnext[c] : 0.0f;
        gcur[c] = gnext[c];
      }
      rx += N;
      rg += N;
      load4(rx, it + 1, xnext);
      if (HASF) {
        rf += N;
        load4(rf, it, fnext);
      }
      load4(rg, it - L, gnext);
      bw_halo(xw[2], lane_l, lane_r);

      // ---- y = Psi(x) - f and r_0 = y - x for row a0 = it-1 (oracle psi_residual) ----
      const int a0 = it - 1;
      const bool a0_int = a0 >= 1 && a0 <= N - 2;
      Row6 rin;  // the row a conv stage consumes
      float y[4];
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        float v = __fmul_rn(0.25f, xw[0].v[c + 1]);
        v = __fmaf_rn(0.25f, xw[1].v[c], v);
        v = __fmaf_rn(0.25f, xw[1].v[c + 2], v);
        v = __fmaf_rn(0.25f, xw[2].v[c + 1], v);
        if (HASF) v = __fsub_rn(v, fcur[c]);
        const bool ok = a0_int && interior[c];
        y[c] = ok ? v : 0.0f;
        rin.v[c + 1] = ok ? __fsub_rn(v, xw[1].v[c + 1]) : 0.0f;
      }
      bw_store4(mine4 + (size_t)((it + 4 * LY::y_slots) % LY::y_slots) * SLOT4, y);
      bw_halo(rin, lane_l, lane_r);

      // ---- forward conv stages k = 1..L: consume r_{k-1} row it-k, finish r_k row it-k-1 (oracle conv3x3_img order:
      //      w0..w8 row-major, first tap a plain product; two partial-sum rows per stage) ----
      float rL[4];
#pragma unroll
      for (int k = 1; k <= L; ++k) {
        // r_{k-1} row it-k waits for the weight gradient of layer k
        bw_store(mine + (size_t)(bw_r_base(L, k) + (it + 8 * bw_r_slots(L, k)) % bw_r_slots(L, k)) * SLOT, rin);
        const float *w = wr[k - 1];
        float fin[4];
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          float m = S[k - 1][0][c];  // taps w0..w5 done
          m = __fmaf_rn(w[6], rin.v[c], m);
          m = __fmaf_rn(w[7], rin.v[c + 1], m);
          m = __fmaf_rn(w[8], rin.v[c + 2], m);
          fin[c] = m;
          float t = S[k - 1][1][c];  // taps w0..w2 done
          t = __fmaf_rn(w[3], rin.v[c], t);
          t = __fmaf_rn(w[4], rin.v[c + 1], t);
          t = __fmaf_rn(w[5], rin.v[c + 2], t);
          S[k - 1][0][c] = t;
          float n = __fmul_rn(w[0], rin.v[c]);
          n = __fmaf_rn(w[1], rin.v[c + 1], n);
          n = __fmaf_rn(w[2], rin.v[c + 2], n);
          S[k - 1][1][c] = n;
        }
        const int ak = it - k - 1;
        const bool ak_int = ak >= 1 && ak <= N - 2;
        if (k < L) {
#pragma unroll
          for (int c = 0; c < 4; ++c) rin.v[c + 1] = (ak_int && interior[c]) ? fin[c] : 0.0f;
          bw_halo(rin, lane_l, lane_r);
        } else {
#pragma unroll
          for (int c = 0; c < 4; ++c) rL[c] = fin[c];
        }
      }

      // ---- g_L = g' * act'(y + r_L) on row aL = it-1-L ----
      const int aL = it - 1 - L;
      const bool aL_int = aL >= 1 && aL <= N - 2;
      Row6 gnew;
      {
        float ys[4], gl[4];
        bw_load4(mine4 + (size_t)((it - L + 4 * LY::y_slots) % LY::y_slots) * SLOT4, ys);
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          const float z = __fadd_rn(ys[c], rL[c]);
          gl[c] = (aL_int && interior[c]) ? __fmul_rn(gcur[c], bw_act_grad<ACT>(z)) : 0.0f;
          gnew.v[c + 1] = gl[c];
        }
        bw_store4(mine4 + (size_t)(LY::y_slots + (it + 4 * LY::g_slots) % LY::g_slots) * SLOT4, gl);
      }

      // ---- adjoint layers l = L..1: g_l row n = it-2L-1+l enters its window; centre row ac = n-1.  The first row of
      //      g_L that an owned output depends on is i0-1-L = the one of step i0: the adjoint (60 % of a step's work)
      //      sleeps through the first half of the band's warm-up ----
      if (it >= i0) {
#pragma unroll
      for (int l = L; l >= 1; --l) {
        bw_halo(gnew, lane_l, lane_r);
        Row6 (&W)[3] = gw[l - 1];
        W[0] = W[1];
        W[1] = W[2];
        W[2] = gnew;
        const int ac = it - 2 * L - 2 + l;
        // weight gradient: dW_l[p][q] += g_l[ac-p+1][c] * r_{l-1}[ac][c+q-1]; a row of g_l counts in the band that owns it
        Row6 r;
        bw_load(mine + (size_t)(bw_r_base(L, l) + (ac + l + 8 * bw_r_slots(L, l)) % bw_r_slots(L, l)) * SLOT, r);
#pragma unroll
        for (int pp = 0; pp < 3; ++pp) {
          const int a = ac - pp + 1;
          if (a >= i0 && a < i1) {
#pragma unroll
            for (int q = 0; q < 3; ++q)
#pragma unroll
              for (int c = 0; c < 4; ++c)
                acc[l - 1][pp * 3 + q] = __fmaf_rn(W[2 - pp].v[c + 1], r.v[c + q], acc[l - 1][pp * 3 + q]);
          }
        }
        // g_{l-1}[ac][c] = sum_{p,q} w_l[p][q] * g_l[ac-p+1][c-q+1]
        const float *w = wr[l - 1];
        const bool ac_int = ac >= 1 && ac <= N - 2;
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          float o[3];  // one short chain per window row (the adjoint's summation order is free)
#pragma unroll
          for (int pp = 0; pp < 3; ++pp) {
            o[pp] = __fmul_rn(w[pp * 3], W[2 - pp].v[c + 2]);
            o[pp] = __fmaf_rn(w[pp * 3 + 1], W[2 - pp].v[c + 1], o[pp]);
            o[pp] = __fmaf_rn(w[pp * 3 + 2], W[2 - pp].v[c], o[pp]);
          }
          gnew.v[c + 1] = (ac_int && interior[c]) ? __fadd_rn(__fadd_rn(o[0], o[1]), o[2]) : 0.0f;
        }
      }

      // ---- g_y = g_L + g_0 on row n0 = it-2L-1; g_x = Psi^T(g_y) - g_0 on row n0-1 ----
      const int n0 = it - 2 * L - 1;
      {
        float gs[4];
        bw_load4(mine4 + (size_t)(LY::y_slots + (it - L + 4 * LY::g_slots) % LY::g_slots) * SLOT4, gs);
        gyw[0] = gyw[1];
        gyw[1] = gyw[2];
#pragma unroll
        for (int c = 0; c < 4; ++c) gyw[2].v[c + 1] = __fadd_rn(gs[c], gnew.v[c + 1]);
        bw_halo(gyw[2], lane_l, lane_r);
      }
      const int ao = n0 - 1;
      if (ao >= i0 && ao < i1 && owner && p.gx != nullptr) {  // gx == NULL: only the weight gradient is wanted
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          float v = __fmul_rn(0.25f, gyw[2].v[c + 1]);
          v = __fmaf_rn(0.25f, gyw[1].v[c + 2], v);
          v = __fmaf_rn(0.25f, gyw[1].v[c], v);
          v = __fmaf_rn(0.25f, gyw[0].v[c + 1], v);
          if (ao >= 1 && ao <= N - 2 && interior[c]) v = __fsub_rn(v, g0prev[c]);
          if (exists[c]) po[ro + c] = v;
        }
      }
#pragma unroll
      for (int c = 0; c < 4; ++c) g0prev[c] = gnew.v[c + 1];
      }  // adjoint
      ro += N;
    }
    if (!owner) {
#pragma unroll
      for (int l = 0; l < L; ++l)
#pragma unroll
        for (int t = 0; t < 9; ++t) acc[l][t] = 0.0f;
    }
  }
  // ---- per-CTA partials, fp64, fixed order ----
#pragma unroll
  for (int l = 0; l < L; ++l)
#pragma unroll
    for (int t = 0; t < 9; ++t) {
      const double v = npde_warp_sum((double)acc[l][t]);
      if (lane == 0) p.part[(size_t)blockIdx.x * (9 * L) + l * 9 + t] = v;
    }
}

// gw[t] (= or +=) sum over CTAs of part[cta][t], one block per tap: per-thread sums over a fixed stride, then a fixed tree
constexpr int FIN_THREADS = 128;
__global__ void __launch_bounds__(FIN_THREADS)
k_bwd_finish(const double *__restrict__ part, int ctas, int taps, float *gw, int accumulate) {
  __shared__ double sh[FIN_THREADS];
  bw_pdl_sync();
  const int t = blockIdx.x;
  double v = 0.0;
  for (int k = threadIdx.x; k < ctas; k += FIN_THREADS) v += part[(size_t)k * taps + t];
  sh[threadIdx.x] = v;
  __syncthreads();
  for (int o = FIN_THREADS / 2; o > 0; o >>= 1) {
    if ((int)threadIdx.x < o) sh[threadIdx.x] += sh[threadIdx.x + o];
    __syncthreads();
  }
  if (threadIdx.x == 0) gw[t] = accumulate ? __fadd_rn(gw[t], (float)sh[0]) : (float)sh[0];
}

// launch with programmatic stream serialization (plain launch in the emulator and inside a stream capture)
template <class K, class... A>
int bw_launch(K kern, dim3 grid, dim3 block, size_t smem, cudaStream_t st, A... args) {
#ifndef NPDE_EMU
  cudaStreamCaptureStatus cap = cudaStreamCaptureStatusNone;
  if (cudaStreamIsCapturing(st, &cap) == cudaSuccess && cap == cudaStreamCaptureStatusNone && !getenv("NPDE_BWD_NO_PDL")) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid;
    cfg.blockDim = block;
    cfg.dynamicSmemBytes = smem;
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    npde_note_launch();
    const cudaError_t e = cudaLaunchKernelEx(&cfg, kern, args...);
    return e == cudaSuccess ? NPDE_OK : (int)e;
  }
#endif
  NPDE_LAUNCH(kern, grid, block, smem, st, args...);
  NPDE_CHECK_LAUNCH();
  return NPDE_OK;
}

struct BwdGeom {
  int Gp, groups, strips;
};
BwdGeom bwd_geom(int B, int N) {
  BwdGeom g;
  g.Gp = (N + 3) / 4;
  g.groups = B * g.Gp;
  g.strips = (g.groups + BV - 1) / BV;
  return g;
}

int bwd_slots() {
#ifdef NPDE_EMU
  return 5;
#else
  int dev = 0, sms = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) return 148 * 8;
  if (cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || sms < 1) sms = 148;
  return sms * 12;
#endif
}

int bwd_bands(int strips, int N, int L) {
#ifdef NPDE_EMU
  const int minrows = 5;
#else
  const int minrows = 4 * L + 4;  // a band pays 4L + 3 extra row steps
#endif
  int bands = bwd_slots() / (strips > 0 ? strips : 1);
  if (bands > N / minrows) bands = N / minrows;
  return bands < 1 ? 1 : bands;
}

template <int L, int ACT, bool HASF>
int launch_bwd(BwdParams &p, cudaStream_t st) {
  const size_t smem = BwdLayout<L>::bytes;
#ifndef NPDE_EMU
  if (smem > 48 * 1024) cudaFuncSetAttribute(k_bwd_stream<L, ACT, HASF>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
#endif
  dim3 grid((unsigned)((long)p.strips * p.bands)), block(32);
  return bw_launch(k_bwd_stream<L, ACT, HASF>, grid, block, smem, st, p);
}

template <int L>
int launch_bwd_l(BwdParams &p, int act, cudaStream_t st) {
  switch (act) {
    case NPDE_ACT_NONE: return p.f ? launch_bwd<L, NPDE_ACT_NONE, true>(p, st) : launch_bwd<L, NPDE_ACT_NONE, false>(p, st);
    case NPDE_ACT_CLAMP: return p.f ? launch_bwd<L, NPDE_ACT_CLAMP, true>(p, st) : launch_bwd<L, NPDE_ACT_CLAMP, false>(p, st);
    case NPDE_ACT_SIGMOID:
      return p.f ? launch_bwd<L, NPDE_ACT_SIGMOID, true>(p, st) : launch_bwd<L, NPDE_ACT_SIGMOID, false>(p, st);
    default: return NPDE_ERR_ACT;
  }
}

}  // namespace

namespace npde {

bool conv_bwd_stream_supported(int bc_kind, int n_layers) {
  return bc_kind == NPDE_BC_SQUARE && n_layers >= 1 && n_layers <= 3;
}

size_t conv_bwd_stream_ws_bytes(int B, int N, int n_layers) {
  const BwdGeom g = bwd_geom(B, N);
  const long ctas = (long)g.strips * bwd_bands(g.strips, N, n_layers);
  return npde_align_up(sizeof(double) * (size_t)ctas * 9 * (size_t)n_layers, 256);
}

int conv_bwd_stream(const float *x, const float *f, const float *w_dev, int n_layers, int act, const float *gout,
                    float *gw, bool accumulate, float *gx, int B, int N, void *ws, size_t ws_bytes, cudaStream_t st) {
  if (!x || !gout || !gw || !w_dev || !ws) return NPDE_ERR_NULL;
  if (ws_bytes < conv_bwd_stream_ws_bytes(B, N, n_layers)) return NPDE_ERR_WORKSPACE;
  const BwdGeom g = bwd_geom(B, N);
  BwdParams p;
  p.x = x; p.f = f; p.gout = gout; p.gx = gx;
  p.part = reinterpret_cast<double *>(ws);
  p.B = B; p.N = N; p.Gp = g.Gp; p.groups = g.groups; p.strips = g.strips;
  p.bands = bwd_bands(g.strips, N, n_layers);
  p.band_rows = (N + p.bands - 1) / p.bands;
  p.bands = (N + p.band_rows - 1) / p.band_rows;
  p.w = w_dev;
  int rc;
  switch (n_layers) {
    case 1: rc = launch_bwd_l<1>(p, act, st); break;
    case 2: rc = launch_bwd_l<2>(p, act, st); break;
    case 3: rc = launch_bwd_l<3>(p, act, st); break;
    default: return NPDE_ERR_LAYERS;
  }
  if (rc != NPDE_OK) return rc;
  const int ctas = p.strips * p.bands, taps = 9 * n_layers;
  return bw_launch(k_bwd_finish, dim3((unsigned)taps), dim3(FIN_THREADS), 0, st, (const double *)p.part, ctas, taps, gw,
                   accumulate ? 1 : 0);
}

}  // namespace npde
