// npde_stream.cu -- the hot loop: fused, register-streaming kernels for
//   * Phi_H = ConvIterator.forward (models/iterators/conv_iterator.py:33-51):
//       Psi -> residual -> L masked 3x3 convs -> +y -> act -> pad -> G   in ONE pass, and
//   * T temporally blocked Jacobi steps (utils/heat_utils.py:96-106 applied T times)
//     per HBM round trip.
//
// Design ("warp strips"):
//   - A warp owns a strip of 128 grid columns and marches down the rows of a
//     band.  Lane l holds columns {2l, 2l+1} of the left half AND {64+2l, 65+2l}
//     of the right half, packed as float2 {left, right}: every arithmetic
//     instruction is a packed FFMA2/FADD2/FMUL2 (fma.rn.f32x2, new on sm_100),
//     which does two IEEE-rn FMAs per issue slot with naturally aligned operands.
//   - Each pipeline stage keeps a 3-row sliding window in registers (rows rotate
//     by a 3x unrolled loop, no register moves).  Horizontal neighbours come from
//     warp shuffles; nothing is exchanged between warps, so there is no
//     __syncthreads anywhere: a strip recomputes a halo of (L+1) columns instead.
//   - Conv stages are software-pipelined two rows apart so that all stages of a
//     step are independent instruction streams (ILP instead of occupancy).
//   - The Jacobi value y of a row is needed 2L steps later; it waits in a
//     per-lane shared-memory delay line (private slots, no synchronisation).
//   - Input rows are prefetched one step ahead into registers and several rows
//     ahead into L2 (prefetch.global.L2).
//   - Arithmetic order is the oracle's: row-major fmaf chain, first tap a plain
//     product (oracle/npde_oracle.c conv3x3_img / orc_fd_step), so iterates are
//     bit-identical to the CPU restatement.
#include "npde_common.cuh"
#include "npde_internal.h"

namespace {

typedef float2 f2;

#ifdef NPDE_EMU
#define NPDE_GRID_CONSTANT
__device__ __forceinline__ f2 fma2(f2 a, f2 b, f2 c) { return make_float2(fmaf(a.x, b.x, c.x), fmaf(a.y, b.y, c.y)); }
__device__ __forceinline__ f2 mul2(f2 a, f2 b) { return make_float2(__fmul_rn(a.x, b.x), __fmul_rn(a.y, b.y)); }
__device__ __forceinline__ f2 add2(f2 a, f2 b) { return make_float2(__fadd_rn(a.x, b.x), __fadd_rn(a.y, b.y)); }
__device__ __forceinline__ f2 sub2(f2 a, f2 b) { return make_float2(__fsub_rn(a.x, b.x), __fsub_rn(a.y, b.y)); }
__device__ __forceinline__ void prefetch_l2(const void *) {}
#else
#define NPDE_GRID_CONSTANT __grid_constant__
typedef unsigned long long u64;
__device__ __forceinline__ u64 pk(f2 a) { return *reinterpret_cast<u64 *>(&a); }
__device__ __forceinline__ f2 upk(u64 a) { return *reinterpret_cast<f2 *>(&a); }
// Packed fp32: two independent round-to-nearest operations per instruction.
// The only products that are later added (first conv tap, 0.25*u) feed an fma's
// addend, which ptxas cannot contract, so the rounding sequence is exactly the
// scalar fmaf chain's.
__device__ __forceinline__ f2 fma2(f2 a, f2 b, f2 c) {
  u64 d;
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(pk(a)), "l"(pk(b)), "l"(pk(c)));
  return upk(d);
}
__device__ __forceinline__ f2 mul2(f2 a, f2 b) {
  u64 d;
  asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(pk(a)), "l"(pk(b)));
  return upk(d);
}
__device__ __forceinline__ f2 add2(f2 a, f2 b) {
  u64 d;
  asm("add.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(pk(a)), "l"(pk(b)));
  return upk(d);
}
__device__ __forceinline__ f2 sub2(f2 a, f2 b) {
  u64 d;
  asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(pk(a)), "l"(pk(b)));
  return upk(d);
}
__device__ __forceinline__ void prefetch_l2(const void *p) { asm volatile("prefetch.global.L2 [%0];" ::"l"(p)); }
#endif

__device__ __forceinline__ f2 bc2(float s) { return make_float2(s, s); }
__device__ __forceinline__ f2 zero2() { return make_float2(0.0f, 0.0f); }
__device__ __forceinline__ f2 shfl2(f2 v, int src) {
  return make_float2(__shfl_sync(NPDE_FULL_MASK, v.x, src), __shfl_sync(NPDE_FULL_MASK, v.y, src));
}

constexpr int STRIP = 128;        // columns per warp strip
constexpr int HALF = STRIP / 2;   // columns per half (the two lanes of a packed value)
constexpr int L2_AHEAD = 6;       // rows of L2 prefetch distance

// Column bookkeeping of one lane: bit0 = left-half col 0, bit1 = left-half col 1,
// bit2 = right-half col 0, bit3 = right-half col 1.
struct LaneCols {
  int colA, colB;
  unsigned in_img;    // column exists (0 <= c < N)
  unsigned interior;  // 1 <= c <= N-2
  unsigned store;     // column belongs to this strip's valid output range
};

__device__ __forceinline__ LaneCols make_cols(int cs, int lane, int N, int vs, int ve) {
  LaneCols c;
  c.colA = cs + 2 * lane;
  c.colB = c.colA + HALF;
  c.in_img = c.interior = c.store = 0u;
#pragma unroll
  for (int q = 0; q < 4; ++q) {
    int col = ((q & 2) ? c.colB : c.colA) + (q & 1);
    if (col >= 0 && col < N) c.in_img |= 1u << q;
    if (col >= 1 && col <= N - 2) c.interior |= 1u << q;
    if (col >= vs && col < ve) c.store |= 1u << q;
  }
  return c;
}

// One grid row (this lane's four columns) -> o0 = {A col0, B col0}, o1 = {A col1, B col1}.
// VEC: rows are 8-byte aligned and the strip starts on an even column, so each
// half is one 64-bit load.
template <bool VEC>
__device__ __forceinline__ void load_row(const float *__restrict__ row, const LaneCols &c, f2 &o0, f2 &o1) {
  float a0 = 0.0f, a1 = 0.0f, b0 = 0.0f, b1 = 0.0f;
  if (VEC && (c.in_img & 3u) == 3u) {
    float2 t = __ldg(reinterpret_cast<const float2 *>(row + c.colA));
    a0 = t.x;
    a1 = t.y;
  } else {
    if (c.in_img & 1u) a0 = __ldg(row + c.colA);
    if (c.in_img & 2u) a1 = __ldg(row + c.colA + 1);
  }
  if (VEC && (c.in_img & 12u) == 12u) {
    float2 t = __ldg(reinterpret_cast<const float2 *>(row + c.colB));
    b0 = t.x;
    b1 = t.y;
  } else {
    if (c.in_img & 4u) b0 = __ldg(row + c.colB);
    if (c.in_img & 8u) b1 = __ldg(row + c.colB + 1);
  }
  o0 = make_float2(a0, b0);
  o1 = make_float2(a1, b1);
}

template <bool VEC>
__device__ __forceinline__ void store_row(float *__restrict__ row, const LaneCols &c, f2 o0, f2 o1) {
  if (VEC && (c.store & 3u) == 3u) {
    *reinterpret_cast<float2 *>(row + c.colA) = make_float2(o0.x, o1.x);
  } else {
    if (c.store & 1u) row[c.colA] = o0.x;
    if (c.store & 2u) row[c.colA + 1] = o1.x;
  }
  if (VEC && (c.store & 12u) == 12u) {
    *reinterpret_cast<float2 *>(row + c.colB) = make_float2(o0.y, o1.y);
  } else {
    if (c.store & 4u) row[c.colB] = o0.y;
    if (c.store & 8u) row[c.colB + 1] = o1.y;
  }
}

__device__ __forceinline__ void prefetch_row(const float *row, const LaneCols &c) {
  if (c.in_img & 1u) prefetch_l2(row + c.colA);
  if (c.in_img & 4u) prefetch_l2(row + c.colB);
}

// Horizontal halos of a freshly produced row.  w[1], w[2] are this lane's own
// columns; w[0] / w[3] become the columns to their left / right.  The seam
// between the halves: lane 31's right neighbour (left half) is lane 0's
// right-half col 0, lane 0's left neighbour (right half) is lane 31's left-half
// col 1.  The strip's outer edges receive don't-care values (they only reach
// cells outside the strip's valid range).
__device__ __forceinline__ void exchange_halo(f2 (&w)[4], int lane) {
  f2 l = shfl2(w[2], (lane + 31) & 31);
  f2 r = shfl2(w[1], (lane + 1) & 31);
  if (lane == 0) l.y = l.x;
  if (lane == 31) r.x = r.y;
  w[0] = l;
  w[3] = r;
}

// Zero the components whose column is not an interior column (bits as LaneCols).
__device__ __forceinline__ void keep_interior(f2 &o0, f2 &o1, unsigned interior) {
  if (!(interior & 1u)) o0.x = 0.0f;
  if (!(interior & 2u)) o1.x = 0.0f;
  if (!(interior & 4u)) o0.y = 0.0f;
  if (!(interior & 8u)) o1.y = 0.0f;
}

// G on the square domain for one output row: ring rows take top/bottom, then the
// ring columns take left/right (columns win the corners, heat_utils.py:55-58).
__device__ __forceinline__ void apply_G_square(f2 &o0, f2 &o1, int row, int N, const LaneCols &c, float top,
                                               float bottom, float left, float right) {
  if (row == 0) {
    o0 = bc2(top);
    o1 = o0;
  } else if (row == N - 1) {
    o0 = bc2(bottom);
    o1 = o0;
  }
  if (c.interior != 15u) {
    if (c.colA == 0) o0.x = left;
    if (c.colA == N - 1) o0.x = right;
    if (c.colA + 1 == N - 1) o1.x = right;  // colA is even or <0, so colA+1 is never 0
    if (c.colA + 1 == 0) o1.x = left;
    if (c.colB == 0) o0.y = left;
    if (c.colB == N - 1) o0.y = right;
    if (c.colB + 1 == N - 1) o1.y = right;
    if (c.colB + 1 == 0) o1.y = left;
  }
}

// 3x3 cross-correlation of a 3-row window for this lane's column c (0 or 1):
// row-major fmaf chain, first tap a plain product (oracle conv3x3_img).
__device__ __forceinline__ f2 conv9(const f2 (&t)[4], const f2 (&m)[4], const f2 (&b)[4], const float *w, int c) {
  f2 acc = mul2(bc2(w[0]), t[c]);
  acc = fma2(bc2(w[1]), t[c + 1], acc);
  acc = fma2(bc2(w[2]), t[c + 2], acc);
  acc = fma2(bc2(w[3]), m[c], acc);
  acc = fma2(bc2(w[4]), m[c + 1], acc);
  acc = fma2(bc2(w[5]), m[c + 2], acc);
  acc = fma2(bc2(w[6]), b[c], acc);
  acc = fma2(bc2(w[7]), b[c + 1], acc);
  acc = fma2(bc2(w[8]), b[c + 2], acc);
  return acc;
}

// Jacobi average, taps in the oracle's order up, left, right, down (orc_fd_step).
__device__ __forceinline__ f2 psi2(f2 u, f2 l, f2 r, f2 d) {
  const f2 q = bc2(0.25f);
  f2 acc = mul2(q, u);
  acc = fma2(q, l, acc);
  acc = fma2(q, r, acc);
  acc = fma2(q, d, acc);
  return acc;
}

__device__ __forceinline__ f2 act2(f2 z, int act) {
  return make_float2(npde_act(z.x, act), npde_act(z.y, act));
}

struct StripItem {
  int b, strip, band;
};
__device__ __forceinline__ bool decode_item(int B, int strips, int bands, StripItem &it) {
  long item = (long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (item >= (long)B * strips * bands) return false;
  it.strip = (int)(item % strips);
  it.band = (int)((item / strips) % bands);
  it.b = (int)(item / ((long)strips * bands));
  return true;
}

// ------------------------------------------------------------ Phi_H (Conv) ---
struct ConvStreamParams {
  const float *in;
  float *out;
  const float *f;    // reference layout (pitch N) or NULL
  const float *bc4;  // (B,4)
  long in_bstride, out_bstride;
  int in_pitch, out_pitch;
  int B, N, act;
  int strips, bands, band_rows, wv;
  float w[4][9];
};

template <int L, bool VIN, bool VOUT>
__global__ void __launch_bounds__(npde::STREAM_THREADS)
k_conv_stream(const NPDE_GRID_CONSTANT ConvStreamParams p) {
  constexpr int R = L + 1;           // dependency radius of Phi_H
  constexpr int RC = (R + 1) & ~1;   // column halo, even: pair loads stay aligned
  constexpr int D = 2 * L + 1;       // delay-line depth: y(row) is consumed 2L steps later
  NPDE_DYN_SMEM(float4, ydl);        // [warp][D][32]: per-lane private slots

  StripItem item;
  if (!decode_item(p.B, p.strips, p.bands, item)) return;
  const int lane = threadIdx.x & 31;
  float4 *delay = ydl + (size_t)(threadIdx.x >> 5) * D * 32 + lane;

  const int N = p.N;
  const int vs = item.strip * p.wv, ve = min(vs + p.wv, N);
  const LaneCols c = make_cols(vs - RC, lane, N, vs, ve);
  const int i0 = item.band * p.band_rows, i1 = min(i0 + p.band_rows, N);
  const float *xin = p.in + (size_t)item.b * p.in_bstride;
  float *xout = p.out + (size_t)item.b * p.out_bstride;
  const float *fin = p.f ? p.f + (size_t)item.b * N * N : nullptr;
  const float *bcb = p.bc4 + (size_t)item.b * 4;
  const float top = __ldg(bcb + 0), bottom = __ldg(bcb + 1), left = __ldg(bcb + 2), right = __ldg(bcb + 3);

  f2 xw[3][4];     // x rows it-2, it-1, it (ring over the step phase)
  f2 rw[L][3][4];  // r_0 .. r_{L-1}: the three newest rows of each stage
#pragma unroll
  for (int s = 0; s < 3; ++s)
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      xw[s][q] = zero2();
#pragma unroll
      for (int k = 0; k < L; ++k) rw[k][s][q] = zero2();
    }
#pragma unroll
  for (int s = 0; s < D; ++s) delay[s * 32] = make_float4(0.f, 0.f, 0.f, 0.f);

  // input rows it = it_first .. it_last; the output row of a step is it - 1 - 2L
  const int it_first = max(i0 - R, 0);  // rows above the grid are zeros == the initial state
  const int it_last = i1 + 2 * L;
  const int x_last = min(i1 - 1 + R, N - 1);  // last input row that exists and matters
  f2 pf[2], pff[2];                            // x row / f row prefetched for the next step
  pf[0] = pf[1] = pff[0] = pff[1] = zero2();
  load_row<VIN>(xin + (size_t)it_first * p.in_pitch, c, pf[0], pf[1]);
  int wslot = 0;  // delay-line write slot of this step (step index mod D)

  for (int it_base = it_first; it_base <= it_last; it_base += 3) {
#pragma unroll
    for (int ph = 0; ph < 3; ++ph) {
      const int it = it_base + ph;
      const int s_new = ph, s_m1 = (ph + 2) % 3, s_m2 = (ph + 1) % 3;  // rows of step t, t-1, t-2

      // ---- stage L: r_L row ro, out = G(act(y + r_L)) ------------------------------
      const int ro = it - 1 - 2 * L;
      const int rslot = (wslot + 1 == D) ? 0 : wslot + 1;  // written 2L steps ago
      if (ro >= i0 && ro < i1) {
        f2 o0, o1;
        if (ro >= 1 && ro <= N - 2) {
          f2 h0 = conv9(rw[L - 1][s_new], rw[L - 1][s_m2], rw[L - 1][s_m1], p.w[L - 1], 0);
          f2 h1 = conv9(rw[L - 1][s_new], rw[L - 1][s_m2], rw[L - 1][s_m1], p.w[L - 1], 1);
          float4 yv = delay[rslot * 32];
          o0 = act2(add2(make_float2(yv.x, yv.y), h0), p.act);
          o1 = act2(add2(make_float2(yv.z, yv.w), h1), p.act);
        } else {
          o0 = o1 = zero2();
        }
        apply_G_square(o0, o1, ro, N, c, top, bottom, left, right);
        store_row<VOUT>(xout + (size_t)ro * p.out_pitch, c, o0, o1);
      }

      // ---- stages L-1 .. 1: r_k row (it - 1 - 2k) from r_{k-1} ------------------------
#pragma unroll
      for (int k = L - 1; k >= 1; --k) {
        const int rk = it - 1 - 2 * k;
        f2 o0, o1;
        if (rk >= 1 && rk <= N - 2) {
          o0 = conv9(rw[k - 1][s_new], rw[k - 1][s_m2], rw[k - 1][s_m1], p.w[k - 1], 0);
          o1 = conv9(rw[k - 1][s_new], rw[k - 1][s_m2], rw[k - 1][s_m1], p.w[k - 1], 1);
          if (c.interior != 15u) keep_interior(o0, o1, c.interior);
        } else {
          o0 = o1 = zero2();
        }
        rw[k][s_new][1] = o0;
        rw[k][s_new][2] = o1;
        exchange_halo(rw[k][s_new], lane);
      }

      // ---- stage 0: x row it arrives; y and r_0 for row it-1 ---------------------------
      xw[s_new][1] = pf[0];
      xw[s_new][2] = pf[1];
      {
        const int r0 = it - 1;
        f2 y0 = psi2(xw[s_m2][1], xw[s_m1][0], xw[s_m1][2], xw[s_new][1]);
        f2 y1 = psi2(xw[s_m2][2], xw[s_m1][1], xw[s_m1][3], xw[s_new][2]);
        if (fin) {
          y0 = sub2(y0, pff[0]);
          y1 = sub2(y1, pff[1]);
        }
        f2 q0 = sub2(y0, xw[s_m1][1]);
        f2 q1 = sub2(y1, xw[s_m1][2]);
        if (r0 >= 1 && r0 <= N - 2) {
          if (c.interior != 15u) keep_interior(q0, q1, c.interior);
        } else {
          q0 = q1 = zero2();
        }
        delay[wslot * 32] = make_float4(y0.x, y0.y, y1.x, y1.y);
        rw[0][s_new][1] = q0;
        rw[0][s_new][2] = q1;
        exchange_halo(rw[0][s_new], lane);
      }
      exchange_halo(xw[s_new], lane);

      // ---- fetch the next rows -----------------------------------------------------------
      {
        const int nx = it + 1;
        if (nx <= x_last) {
          load_row<VIN>(xin + (size_t)nx * p.in_pitch, c, pf[0], pf[1]);
        } else {
          pf[0] = pf[1] = zero2();
        }
        if (fin) {  // f row for the next step's stage 0 is row `it`
          if (it >= 1 && it <= N - 2) load_row<false>(fin + (size_t)it * N, c, pff[0], pff[1]);
          else pff[0] = pff[1] = zero2();
          if (it + L2_AHEAD <= x_last) prefetch_row(fin + (size_t)(it + L2_AHEAD) * N, c);
        }
        if (nx + L2_AHEAD <= x_last) prefetch_row(xin + (size_t)(nx + L2_AHEAD) * p.in_pitch, c);
      }
      wslot = (wslot + 1 == D) ? 0 : wslot + 1;
    }
  }
}

// --------------------------------------------------------- T Jacobi steps ---
struct JacobiStreamParams {
  const float *in;
  float *out;
  const float *f;    // reference layout (pitch N) or NULL; only with T == 1
  const float *bc4;  // (B,4)
  long in_bstride, out_bstride;
  int in_pitch, out_pitch;
  int B, N;
  int strips, bands, band_rows, wv;
};

template <int T, bool VIN, bool VOUT>
__global__ void __launch_bounds__(npde::STREAM_THREADS)
k_jacobi_stream(const NPDE_GRID_CONSTANT JacobiStreamParams p) {
  constexpr int RC = (T + 1) & ~1;
  StripItem item;
  if (!decode_item(p.B, p.strips, p.bands, item)) return;
  const int lane = threadIdx.x & 31;
  const int N = p.N;
  const int vs = item.strip * p.wv, ve = min(vs + p.wv, N);
  const LaneCols c = make_cols(vs - RC, lane, N, vs, ve);
  const int i0 = item.band * p.band_rows, i1 = min(i0 + p.band_rows, N);
  const float *xin = p.in + (size_t)item.b * p.in_bstride;
  float *xout = p.out + (size_t)item.b * p.out_bstride;
  const float *fin = (T == 1 && p.f) ? p.f + (size_t)item.b * N * N : nullptr;
  const float *bcb = p.bc4 + (size_t)item.b * 4;
  const float top = __ldg(bcb + 0), bottom = __ldg(bcb + 1), left = __ldg(bcb + 2), right = __ldg(bcb + 3);

  // win[0] = x, win[s] = iterate after s steps (s < T); rows of step t, t-1, t-2
  f2 win[T][3][4];
#pragma unroll
  for (int s = 0; s < T; ++s)
#pragma unroll
    for (int r = 0; r < 3; ++r)
#pragma unroll
      for (int q = 0; q < 4; ++q) win[s][r][q] = zero2();

  // input rows it; stage s (1..T) produces row it - s in the same step (forward
  // order: the 5-point stencil needs no halo on its newest row)
  const int it_first = max(i0 - T, 0);
  const int it_last = i1 - 1 + T;
  const int x_last = min(it_last, N - 1);
  f2 pf[2], pff[2];
  pf[0] = pf[1] = pff[0] = pff[1] = zero2();
  load_row<VIN>(xin + (size_t)it_first * p.in_pitch, c, pf[0], pf[1]);
  if (fin && it_first >= 1) load_row<false>(fin + (size_t)(it_first - 1) * N, c, pff[0], pff[1]);

  for (int it_base = it_first; it_base <= it_last; it_base += 3) {
#pragma unroll
    for (int ph = 0; ph < 3; ++ph) {
      const int it = it_base + ph;
      const int s_new = ph, s_m1 = (ph + 2) % 3, s_m2 = (ph + 1) % 3;
      win[0][s_new][1] = pf[0];
      win[0][s_new][2] = pf[1];
#pragma unroll
      for (int s = 1; s <= T; ++s) {
        const int row = it - s;
        f2 o0 = psi2(win[s - 1][s_m2][1], win[s - 1][s_m1][0], win[s - 1][s_m1][2], win[s - 1][s_new][1]);
        f2 o1 = psi2(win[s - 1][s_m2][2], win[s - 1][s_m1][1], win[s - 1][s_m1][3], win[s - 1][s_new][2]);
        if (T == 1 && fin) {
          o0 = sub2(o0, pff[0]);
          o1 = sub2(o1, pff[1]);
        }
        apply_G_square(o0, o1, row, N, c, top, bottom, left, right);
        if (s < T) {
          // rows above/below the grid must read as zero padding for the next stage's ring
          // cells only, and those are overwritten by G: no masking needed.
          win[s < T ? s : 0][s_new][1] = o0;
          win[s < T ? s : 0][s_new][2] = o1;
        } else if (row >= i0 && row < i1) {
          store_row<VOUT>(xout + (size_t)row * p.out_pitch, c, o0, o1);
        }
      }
#pragma unroll
      for (int s = 0; s < T; ++s) exchange_halo(win[s][s_new], lane);

      const int nx = it + 1;
      if (nx <= x_last) {
        load_row<VIN>(xin + (size_t)nx * p.in_pitch, c, pf[0], pf[1]);
      } else {
        pf[0] = pf[1] = zero2();
      }
      if (T == 1 && fin) {  // next step's output row is `it`
        if (it <= N - 1) load_row<false>(fin + (size_t)it * N, c, pff[0], pff[1]);
        if (it + L2_AHEAD <= x_last) prefetch_row(fin + (size_t)(it + L2_AHEAD) * N, c);
      }
      if (nx + L2_AHEAD <= x_last) prefetch_row(xin + (size_t)(nx + L2_AHEAD) * p.in_pitch, c);
    }
  }
}

// ------------------------------------------------------------- host side ---
// Work split: B * strips * bands warp items.  Bands exist only to fill the
// machine when B * strips is small; every band re-streams 2R halo rows.
struct Split {
  int strips, bands, band_rows, wv;
};

Split make_split(int B, int N, int halo_cols, int halo_rows, int resident_warps) {
  Split s;
  s.wv = STRIP - 2 * halo_cols;
  s.strips = (N + s.wv - 1) / s.wv;
  long per_band = (long)B * s.strips;
  int bands = (int)(resident_warps / (per_band > 0 ? per_band : 1));
  int max_bands = N / (8 * (halo_rows > 0 ? halo_rows : 1));  // keep the halo re-streaming <= 25%
  if (bands > max_bands) bands = max_bands;
  if (bands < 1) bands = 1;
  s.band_rows = (N + bands - 1) / bands;
  s.bands = (N + s.band_rows - 1) / s.band_rows;
  return s;
}

int resident_warps_cached(const void *kernel, size_t smem, int *cache) {
#ifdef NPDE_EMU
  (void)kernel; (void)smem; (void)cache;
  return 8;  // small on purpose: exercises several bands per image in the emulator
#else
  if (*cache > 0) return *cache;
  int dev = 0, sms = 0, blocks = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) return 148 * 16;
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&blocks, kernel, npde::STREAM_THREADS, smem) != cudaSuccess ||
      blocks < 1)
    blocks = 1;
  *cache = sms * blocks * (npde::STREAM_THREADS / 32);
  return *cache;
#endif
}

inline bool vec_ok(const void *base, long bstride, int pitch) {
  return ((uintptr_t)base % 8 == 0) && (bstride % 2 == 0) && (pitch % 2 == 0);
}

template <int L>
int launch_conv(ConvStreamParams &p, cudaStream_t st) {
  constexpr int R = L + 1, RC = (R + 1) & ~1, D = 2 * L + 1;
  const size_t smem = sizeof(float4) * D * 32 * (npde::STREAM_THREADS / 32);
  const bool vin = vec_ok(p.in, p.in_bstride, p.in_pitch), vout = vec_ok(p.out, p.out_bstride, p.out_pitch);
  static int cache = 0;
  int resident = resident_warps_cached((const void *)k_conv_stream<L, true, true>, smem, &cache);
  Split s = make_split(p.B, p.N, RC, R, resident);
  p.strips = s.strips; p.bands = s.bands; p.band_rows = s.band_rows; p.wv = s.wv;
  long items = (long)p.B * s.strips * s.bands;
  const int wpb = npde::STREAM_THREADS / 32;
  dim3 grid((unsigned)((items + wpb - 1) / wpb)), block(npde::STREAM_THREADS);
  if (vin && vout) NPDE_LAUNCH((k_conv_stream<L, true, true>), grid, block, smem, st, p);
  else if (vin) NPDE_LAUNCH((k_conv_stream<L, true, false>), grid, block, smem, st, p);
  else if (vout) NPDE_LAUNCH((k_conv_stream<L, false, true>), grid, block, smem, st, p);
  else NPDE_LAUNCH((k_conv_stream<L, false, false>), grid, block, smem, st, p);
  NPDE_CHECK_LAUNCH();
  return NPDE_OK;
}

template <int T>
int launch_jacobi(JacobiStreamParams &p, cudaStream_t st) {
  constexpr int RC = (T + 1) & ~1;
  const bool vin = vec_ok(p.in, p.in_bstride, p.in_pitch), vout = vec_ok(p.out, p.out_bstride, p.out_pitch);
  static int cache = 0;
  int resident = resident_warps_cached((const void *)k_jacobi_stream<T, true, true>, 0, &cache);
  Split s = make_split(p.B, p.N, RC, T, resident);
  p.strips = s.strips; p.bands = s.bands; p.band_rows = s.band_rows; p.wv = s.wv;
  long items = (long)p.B * s.strips * s.bands;
  const int wpb = npde::STREAM_THREADS / 32;
  dim3 grid((unsigned)((items + wpb - 1) / wpb)), block(npde::STREAM_THREADS);
  if (vin && vout) NPDE_LAUNCH((k_jacobi_stream<T, true, true>), grid, block, 0, st, p);
  else if (vin) NPDE_LAUNCH((k_jacobi_stream<T, true, false>), grid, block, 0, st, p);
  else if (vout) NPDE_LAUNCH((k_jacobi_stream<T, false, true>), grid, block, 0, st, p);
  else NPDE_LAUNCH((k_jacobi_stream<T, false, false>), grid, block, 0, st, p);
  NPDE_CHECK_LAUNCH();
  return NPDE_OK;
}

}  // namespace

namespace npde {

bool conv_stream_supported(int bc_kind, int n_layers, const float *w_host) {
  return bc_kind == NPDE_BC_SQUARE && n_layers >= 1 && n_layers <= 4 && w_host != nullptr;
}

int conv_stream(const FieldView &in, const FieldViewMut &out, const float *bc4, const float *f, const float *w_host,
                int n_layers, int act, int B, int N, cudaStream_t st) {
  ConvStreamParams p;
  p.in = in.p; p.in_pitch = in.pitch; p.in_bstride = in.bstride;
  p.out = out.p; p.out_pitch = out.pitch; p.out_bstride = out.bstride;
  p.f = f; p.bc4 = bc4; p.B = B; p.N = N; p.act = act;
  for (int l = 0; l < 4; ++l)
    for (int t = 0; t < 9; ++t) p.w[l][t] = (l < n_layers) ? w_host[l * 9 + t] : 0.0f;
  switch (n_layers) {
    case 1: return launch_conv<1>(p, st);
    case 2: return launch_conv<2>(p, st);
    case 3: return launch_conv<3>(p, st);
    case 4: return launch_conv<4>(p, st);
    default: return NPDE_ERR_LAYERS;
  }
}

int jacobi_stream_max_depth(const float *f) { return f ? 1 : 8; }

int jacobi_stream(const FieldView &in, const FieldViewMut &out, const float *bc4, const float *f, int depth, int B,
                  int N, cudaStream_t st) {
  JacobiStreamParams p;
  p.in = in.p; p.in_pitch = in.pitch; p.in_bstride = in.bstride;
  p.out = out.p; p.out_pitch = out.pitch; p.out_bstride = out.bstride;
  p.f = f; p.bc4 = bc4; p.B = B; p.N = N;
  if (f && depth != 1) return NPDE_ERR_ARG;
  switch (depth) {
    case 1: return launch_jacobi<1>(p, st);
    case 2: return launch_jacobi<2>(p, st);
    case 4: return launch_jacobi<4>(p, st);
    case 8: return launch_jacobi<8>(p, st);
    default: return NPDE_ERR_ARG;
  }
}

}  // namespace npde
