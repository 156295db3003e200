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
#ifndef NPDE_L2_AHEAD
#define NPDE_L2_AHEAD 6
#endif
constexpr int L2_AHEAD = NPDE_L2_AHEAD;  // rows of L2 prefetch distance (0: no prefetch instructions)

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
// Four 32-bit loads straight into the halves of the packed registers: a 64-bit load per half
// would need four register moves to re-pair {A,B}, and rows of a 2^n+1 grid are not 8-byte
// aligned anyway.  The two loads of a half hit the same 128-byte lines (L1).
__device__ __forceinline__ void load_row(const float *__restrict__ row, const LaneCols &c, f2 &o0, f2 &o1) {
  o0.x = (c.in_img & 1u) ? __ldg(row + c.colA) : 0.0f;
  o1.x = (c.in_img & 2u) ? __ldg(row + c.colA + 1) : 0.0f;
  o0.y = (c.in_img & 4u) ? __ldg(row + c.colB) : 0.0f;
  o1.y = (c.in_img & 8u) ? __ldg(row + c.colB + 1) : 0.0f;
}
// same for a strip whose 128 columns all exist; p points at this lane's first column
__device__ __forceinline__ void load_row_full(const float *__restrict__ p, f2 &o0, f2 &o1) {
  o0.x = __ldg(p);
  o1.x = __ldg(p + 1);
  o0.y = __ldg(p + HALF);
  o1.y = __ldg(p + HALF + 1);
}

// mask bits as LaneCols: which of the four columns this lane writes
__device__ __forceinline__ void store_row(float *__restrict__ p, unsigned mask, f2 o0, f2 o1) {
  if (mask & 1u) p[0] = o0.x;
  if (mask & 2u) p[1] = o1.x;
  if (mask & 4u) p[HALF] = o0.y;
  if (mask & 8u) p[HALF + 1] = o1.y;
}

__device__ __forceinline__ void prefetch_row(const float *p, unsigned mask) {
  if (mask & 1u) prefetch_l2(p);
  if (mask & 4u) prefetch_l2(p + HALF);
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
// One warp per CTA: the work item depends on blockIdx only, so ptxas can prove that the band /
// strip bookkeeping and the loop-versioning branches are warp-uniform and keep the conv weights
// in uniform registers (FFMA2 R, R, UR.F32, R) instead of re-loading them from the constant bank.
__device__ __forceinline__ bool decode_item(int B, int strips, int bands, StripItem &it) {
  long item = (long)blockIdx.x;
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

__device__ __forceinline__ float and_bits(float v, unsigned m) { return __uint_as_float(__float_as_uint(v) & m); }
__device__ __forceinline__ float sel_bits(float v, unsigned keep, unsigned val) {
  return __uint_as_float((__float_as_uint(v) & keep) | val);  // one LOP3
}

// iterator.py:17-27.  Sigmoid is kept out of line: it is rare and its expf/division
// expansion would otherwise sit in the middle of the streaming loop.
__device__ __noinline__ float sigmoid_slow(float z) { return 1.0f / (1.0f + expf(-z)); }
__device__ __forceinline__ f2 act_pair(f2 z, int act) {
  if (act == NPDE_ACT_CLAMP) return make_float2(fminf(fmaxf(z.x, 0.0f), 1.0f), fminf(fmaxf(z.y, 0.0f), 1.0f));
  if (act == NPDE_ACT_SIGMOID) return make_float2(sigmoid_slow(z.x), sigmoid_slow(z.y));
  return z;
}

// Streaming form of the row-major 3x3 chain.  A row of the stage input (with its
// halos) arrives once and serves three output rows: it FINISHES the chain of the
// row above it (taps w6..w8), CONTINUES the chain of its own row (w3..w5) and
// STARTS the chain of the row below (w0..w2, first tap a plain product).  Per
// output the operation order is exactly w0..w8, so results equal the windowed
// form bit for bit, but only two partial sums per column stay live instead of a
// 3-row window.
struct ConvAcc {
  f2 mid[2];  // after w0..w5: finished by the next row
  f2 top[2];  // after w0..w2
};
__device__ __forceinline__ void conv_feed(const f2 (&row)[4], const float *w, ConvAcc &a, f2 &out0, f2 &out1) {
#pragma unroll
  for (int c = 0; c < 2; ++c) {
    f2 fin = fma2(bc2(w[6]), row[c], a.mid[c]);
    fin = fma2(bc2(w[7]), row[c + 1], fin);
    fin = fma2(bc2(w[8]), row[c + 2], fin);
    f2 mid = fma2(bc2(w[3]), row[c], a.top[c]);
    mid = fma2(bc2(w[4]), row[c + 1], mid);
    mid = fma2(bc2(w[5]), row[c + 2], mid);
    f2 top = mul2(bc2(w[0]), row[c]);
    top = fma2(bc2(w[1]), row[c + 1], top);
    top = fma2(bc2(w[2]), row[c + 2], top);
    a.mid[c] = mid;
    a.top[c] = top;
    if (c == 0) out0 = fin; else out1 = fin;
  }
}

template <int L>
struct ConvState {
  f2 xb[4];      // x row it-1 with halos
  f2 yu[2];      // 0.25 * x[it-2]: first tap of Psi for output row it-1
  f2 rb[L][4];   // rb[k]: the row of r_k produced in the previous step, with halos
  ConvAcc acc[L];
  f2 pf[2][2], pff[2][2];  // [parity of it]: x row `it` / f row it-1, fetched two steps earlier
};

struct ConvLaneCtx {
  unsigned keep[4];  // per column (A0, A1, B0, B1): all ones if interior, else 0
  unsigned gval[4];  // Dirichlet value bits of a ring column, else 0
  unsigned in_img, store;  // LaneCols bit masks, pinned in registers
  const float *pin;  // this lane's first column in the NEXT input row to fetch
  const float *pfin; // same for f (row `it`)
  float *pout;       // this lane's first column in the output row of the current step
  char *delay;       // this lane's slot 0 of the y delay line (shared memory)
  float top, bottom;
  int N, i0, i1, x_last, in_pitch, out_pitch, lane_l, lane_r;
  bool first, last;  // lane 0 / lane 31
};

#ifndef NPDE_EMU
// Keeps a loop-invariant per-lane constant in a register: without this ptxas
// rematerialises the column masks from the column index in every step (2-3
// instructions each) to save registers.
__device__ __forceinline__ void pin_reg(unsigned &v) { asm volatile("" : "+r"(v)); }
#else
__device__ __forceinline__ void pin_reg(unsigned &) {}
#endif

__device__ __forceinline__ void exchange_halo_ctx(f2 (&w)[4], const ConvLaneCtx &x) {
  f2 l = shfl2(w[2], x.lane_l);
  f2 r = shfl2(w[1], x.lane_r);
  if (x.first) l.y = l.x;
  if (x.last) r.x = r.y;
  w[0] = l;
  w[3] = r;
}

template <int ACT>
__device__ __forceinline__ f2 act_t(f2 z) {
  if (ACT == NPDE_ACT_CLAMP) return make_float2(fminf(fmaxf(z.x, 0.0f), 1.0f), fminf(fmaxf(z.y, 0.0f), 1.0f));
  if (ACT == NPDE_ACT_SIGMOID) return make_float2(sigmoid_slow(z.x), sigmoid_slow(z.y));
  return z;
}

// One step: input row `it` has arrived in s.pf.  Stage k (0..L) works on row it-1-2k.
// STEADY: every stage row (it-1-2k, k=0..L) is an interior row of the grid: no ring-row handling.
//         Whether the output row belongs to this band and whether the rows to fetch exist are
//         uniform predicates evaluated in every step (a few instructions), so the warm-up and
//         drain steps of a band in the middle of the grid run the same fast code.
// EDGE:   the strip contains ring or out-of-grid columns: loads are predicated per column,
//         stage outputs are masked to the interior columns and G fixes the ring columns.
// DS = delay-line slots (power of two >= 2L+1), 512 bytes per slot and warp.
// PH = parity slot of the row buffers used by this step (the loop alternates 0,1).
template <int L, int ACT, bool HASF, bool STEADY, bool EDGE, int PH>
__device__ __forceinline__ void conv_step(ConvState<L> &s, ConvLaneCtx &x, const ConvStreamParams &p, int it,
                                          unsigned &woff) {
  f2 (&pf)[2] = s.pf[PH];
  f2 (&pff)[2] = s.pff[PH];
  constexpr int DS = (2 * L + 1 <= 8) ? 8 : 16;
  constexpr unsigned RING = DS * 512u - 1u;
  const int N = x.N;
  const unsigned roff = (woff + (DS - 2 * L) * 512u) & RING;  // y written 2L steps ago

  // ---- stages L .. 1 (reverse order: each reads the row its producer made LAST step) ----
#pragma unroll
  for (int k = L; k >= 1; --k) {
    const int rk = it - 1 - 2 * k;
    f2 h0, h1;
    conv_feed(s.rb[k - 1], p.w[k - 1], s.acc[k - 1], h0, h1);
    if (k == L) {
      if (rk >= x.i0 && rk < x.i1) {
        f2 o0, o1;
        if (STEADY || (rk >= 1 && rk <= N - 2)) {
          float4 yv = *reinterpret_cast<const float4 *>(x.delay + roff);
          o0 = act_t<ACT>(add2(make_float2(yv.x, yv.y), h0));
          o1 = act_t<ACT>(add2(make_float2(yv.z, yv.w), h1));
        } else {
          o0 = o1 = bc2(rk == 0 ? x.top : x.bottom);  // ring row: pad then G
        }
        if (EDGE) {  // pad + G on the columns: non-interior columns take the Dirichlet value
          o0.x = sel_bits(o0.x, x.keep[0], x.gval[0]);
          o1.x = sel_bits(o1.x, x.keep[1], x.gval[1]);
          o0.y = sel_bits(o0.y, x.keep[2], x.gval[2]);
          o1.y = sel_bits(o1.y, x.keep[3], x.gval[3]);
        }
        store_row(x.pout, x.store, o0, o1);
      }
      if (STEADY || rk >= 0) x.pout += x.out_pitch;  // invariant: pout points at row max(rk, 0)
    } else {
      if (!STEADY && !(rk >= 1 && rk <= N - 2)) h0 = h1 = zero2();
      if (EDGE) {
        h0.x = and_bits(h0.x, x.keep[0]);
        h1.x = and_bits(h1.x, x.keep[1]);
        h0.y = and_bits(h0.y, x.keep[2]);
        h1.y = and_bits(h1.y, x.keep[3]);
      }
      s.rb[k][1] = h0;
      s.rb[k][2] = h1;
      exchange_halo_ctx(s.rb[k], x);
    }
  }

  // ---- stage 0: y = Psi(x) - f and r_0 = y - x for row it-1 ------------------------------
  {
    const int r0 = it - 1;
    const f2 q = bc2(0.25f);
    f2 y0 = fma2(q, s.xb[0], s.yu[0]);  // taps: up (in yu), left, right, down
    y0 = fma2(q, s.xb[2], y0);
    y0 = fma2(q, pf[0], y0);
    f2 y1 = fma2(q, s.xb[1], s.yu[1]);
    y1 = fma2(q, s.xb[3], y1);
    y1 = fma2(q, pf[1], y1);
    if (HASF) {
      y0 = sub2(y0, pff[0]);
      y1 = sub2(y1, pff[1]);
    }
    f2 q0 = sub2(y0, s.xb[1]);
    f2 q1 = sub2(y1, s.xb[2]);
    if (!STEADY && !(r0 >= 1 && r0 <= N - 2)) q0 = q1 = zero2();
    if (EDGE) {
      q0.x = and_bits(q0.x, x.keep[0]);
      q1.x = and_bits(q1.x, x.keep[1]);
      q0.y = and_bits(q0.y, x.keep[2]);
      q1.y = and_bits(q1.y, x.keep[3]);
    }
    *reinterpret_cast<float4 *>(x.delay + woff) = make_float4(y0.x, y0.y, y1.x, y1.y);
    s.yu[0] = mul2(q, s.xb[1]);  // first tap of the next output row (row it): up = x[it-1]
    s.yu[1] = mul2(q, s.xb[2]);
    s.rb[0][1] = q0;
    s.rb[0][2] = q1;
    exchange_halo_ctx(s.rb[0], x);
    s.xb[1] = pf[0];
    s.xb[2] = pf[1];
    exchange_halo_ctx(s.xb, x);
  }

  // ---- refill this parity's buffers: x row it+2 and f row it+1 (consumed two steps from now) ----
  {
    const int nx = it + 2;
    if (nx <= x.x_last) {
      if (EDGE) {
        pf[0].x = (x.in_img & 1u) ? __ldg(x.pin) : 0.0f;
        pf[1].x = (x.in_img & 2u) ? __ldg(x.pin + 1) : 0.0f;
        pf[0].y = (x.in_img & 4u) ? __ldg(x.pin + HALF) : 0.0f;
        pf[1].y = (x.in_img & 8u) ? __ldg(x.pin + HALF + 1) : 0.0f;
      } else {
        load_row_full(x.pin, pf[0], pf[1]);
      }
      if (L2_AHEAD > 0 && nx + L2_AHEAD <= x.x_last)  // never prefetch past the rows this band may read
        prefetch_row(x.pin + (size_t)L2_AHEAD * x.in_pitch, EDGE ? x.in_img : 15u);
      x.pin += x.in_pitch;
    } else {
      pf[0] = pf[1] = zero2();
    }
    if (HASF) {
      const int nf = it + 1;  // f row used by the stage 0 of step it+2
      if (nf <= x.x_last && (STEADY || (nf >= 1 && nf <= N - 2))) {
        if (EDGE) {
          pff[0].x = (x.in_img & 1u) ? __ldg(x.pfin) : 0.0f;
          pff[1].x = (x.in_img & 2u) ? __ldg(x.pfin + 1) : 0.0f;
          pff[0].y = (x.in_img & 4u) ? __ldg(x.pfin + HALF) : 0.0f;
          pff[1].y = (x.in_img & 8u) ? __ldg(x.pfin + HALF + 1) : 0.0f;
        } else {
          load_row_full(x.pfin, pff[0], pff[1]);
        }
      } else {
        pff[0] = pff[1] = zero2();
      }
      if (L2_AHEAD > 0 && nf + L2_AHEAD <= x.x_last) prefetch_row(x.pfin + (size_t)L2_AHEAD * N, EDGE ? x.in_img : 15u);
      x.pfin += N;
    }
  }
  woff = (woff + 512u) & RING;
}

template <int L, int ACT, bool HASF>
#ifndef NPDE_CONV_MIN_CTAS
#define NPDE_CONV_MIN_CTAS 16
#endif
__global__ void __launch_bounds__(npde::STREAM_THREADS, NPDE_CONV_MIN_CTAS)
k_conv_stream(const NPDE_GRID_CONSTANT ConvStreamParams p) {
  constexpr int R = L + 1;           // dependency radius of Phi_H
  constexpr int RC = (R + 1) & ~1;   // column halo (even)
  constexpr int DS = (2 * L + 1 <= 8) ? 8 : 16;
  NPDE_DYN_SMEM(float4, ydl);        // [warp][DS][32]: per-lane private slots

  StripItem item;
  if (!decode_item(p.B, p.strips, p.bands, item)) return;
  ConvLaneCtx x;
  const int lane = threadIdx.x & 31;
  x.lane_l = (lane + 31) & 31;
  x.lane_r = (lane + 1) & 31;
  x.first = lane == 0;
  x.last = lane == 31;
  x.delay = reinterpret_cast<char *>(ydl + lane);
  x.N = p.N;
  const int N = p.N;
  const int vs = item.strip * p.wv, ve = min(vs + p.wv, N);
  const int cs = vs - RC;
  const LaneCols c = make_cols(cs, lane, N, vs, ve);
  x.in_img = c.in_img;
  x.store = c.store;
  pin_reg(x.in_img);
  pin_reg(x.store);
  x.i0 = item.band * p.band_rows;
  x.i1 = min(x.i0 + p.band_rows, N);
  x.in_pitch = p.in_pitch;
  x.out_pitch = p.out_pitch;
  const float *bcb = p.bc4 + (size_t)item.b * 4;
  x.top = __ldg(bcb + 0);
  x.bottom = __ldg(bcb + 1);
  const float left = __ldg(bcb + 2), right = __ldg(bcb + 3);
#pragma unroll
  for (int q = 0; q < 4; ++q) {
    const int col = ((q & 2) ? c.colB : c.colA) + (q & 1);
    x.keep[q] = (c.interior >> q & 1u) ? 0xffffffffu : 0u;
    x.gval[q] = col == 0 ? __float_as_uint(left) : (col == N - 1 ? __float_as_uint(right) : 0u);
    pin_reg(x.keep[q]);
    pin_reg(x.gval[q]);
  }
  // the strip needs the EDGE code iff one of its 128 columns is not an interior column
  const bool edge_strip = cs < 1 || cs + STRIP - 1 > N - 2;

  ConvState<L> s;
#pragma unroll
  for (int q = 0; q < 4; ++q) {
    s.xb[q] = zero2();
#pragma unroll
    for (int k = 0; k < L; ++k) s.rb[k][q] = zero2();
  }
#pragma unroll
  for (int k = 0; k < L; ++k) s.acc[k].mid[0] = s.acc[k].mid[1] = s.acc[k].top[0] = s.acc[k].top[1] = zero2();
  s.yu[0] = s.yu[1] = zero2();
#pragma unroll
  for (int h = 0; h < 2; ++h) s.pf[h][0] = s.pf[h][1] = s.pff[h][0] = s.pff[h][1] = zero2();
#pragma unroll
  for (int d = 0; d < DS; ++d) *reinterpret_cast<float4 *>(x.delay + d * 512) = make_float4(0.f, 0.f, 0.f, 0.f);

  // input rows it = it_first .. it_last; the output row of a step is it - 1 - 2L
  const int it_first = max(x.i0 - R, 0);  // rows above the grid are zeros == the initial state
  const int it_last = x.i1 + 2 * L;
  x.x_last = min(x.i1 - 1 + R, N - 1);    // last input row that exists and matters
  const float *xin = p.in + (size_t)item.b * p.in_bstride;
  const float *fimg = HASF ? p.f + (size_t)item.b * N * N : nullptr;
  // buffers of parity 0 / 1 hold rows it_first / it_first+1 (f: one row behind)
  load_row(xin + (size_t)it_first * p.in_pitch, c, s.pf[0][0], s.pf[0][1]);
  if (it_first + 1 <= x.x_last) load_row(xin + (size_t)(it_first + 1) * p.in_pitch, c, s.pf[1][0], s.pf[1][1]);
  if (HASF) {
    if (it_first - 1 >= 1) load_row(fimg + (size_t)(it_first - 1) * N, c, s.pff[0][0], s.pff[0][1]);
    if (it_first >= 1 && it_first <= N - 2) load_row(fimg + (size_t)it_first * N, c, s.pff[1][0], s.pff[1][1]);
  }
  // running pointers: pin -> row it+2 of x, pfin -> row it+1 of f, pout -> output row it-1-2L (clamped at 0)
  x.pin = xin + (size_t)(it_first + 2) * p.in_pitch + c.colA;
  x.pfin = HASF ? fimg + (size_t)(it_first + 1) * N + c.colA : nullptr;
  x.pout = p.out + (size_t)item.b * p.out_bstride + (size_t)max(it_first - 1 - 2 * L, 0) * p.out_pitch + c.colA;
  unsigned woff = 0;

  // steady range of `it`: all stage rows it-1-2L .. it-1 are interior rows (and so is f row it+1)
  const int steady_lo = 2 + 2 * L;
  const int steady_hi = min(N - 3, it_last);
  int it = it_first;
  while (it <= it_last) {
    // parity bookkeeping: steps alternate buffer 0,1 starting with 0 at it_first
    if (((it - it_first) & 1) == 0 && it >= steady_lo && it + 1 <= steady_hi) {
      if (edge_strip) {
        do {
          conv_step<L, ACT, HASF, true, true, 0>(s, x, p, it, woff);
          conv_step<L, ACT, HASF, true, true, 1>(s, x, p, it + 1, woff);
          it += 2;
        } while (it + 1 <= steady_hi);
      } else {
        do {
          conv_step<L, ACT, HASF, true, false, 0>(s, x, p, it, woff);
          conv_step<L, ACT, HASF, true, false, 1>(s, x, p, it + 1, woff);
          it += 2;
        } while (it + 1 <= steady_hi);
      }
    } else if (((it - it_first) & 1) == 0) {
      conv_step<L, ACT, HASF, false, true, 0>(s, x, p, it, woff);
      ++it;
    } else {
      conv_step<L, ACT, HASF, false, true, 1>(s, x, p, it, woff);
      ++it;
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

template <int T>
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
  load_row(xin + (size_t)it_first * p.in_pitch, c, pf[0], pf[1]);
  if (fin && it_first >= 1) load_row(fin + (size_t)(it_first - 1) * N, c, pff[0], pff[1]);

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
          store_row(xout + (size_t)row * p.out_pitch + c.colA, c.store, o0, o1);
        }
      }
#pragma unroll
      for (int s = 0; s < T; ++s) exchange_halo(win[s][s_new], lane);

      const int nx = it + 1;
      if (nx <= x_last) {
        load_row(xin + (size_t)nx * p.in_pitch, c, pf[0], pf[1]);
      } else {
        pf[0] = pf[1] = zero2();
      }
      if (T == 1 && fin) {  // next step's output row is `it`
        if (it <= N - 1) load_row(fin + (size_t)it * N, c, pff[0], pff[1]);
        if (it + L2_AHEAD <= x_last) prefetch_row(fin + (size_t)(it + L2_AHEAD) * N + c.colA, c.in_img);
      }
      if (nx + L2_AHEAD <= x_last) prefetch_row(xin + (size_t)(nx + L2_AHEAD) * p.in_pitch + c.colA, c.in_img);
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

template <int L, int ACT>
int launch_conv_f(ConvStreamParams &p, dim3 grid, dim3 block, size_t smem, cudaStream_t st) {
  if (p.f) NPDE_LAUNCH((k_conv_stream<L, ACT, true>), grid, block, smem, st, p);
  else NPDE_LAUNCH((k_conv_stream<L, ACT, false>), grid, block, smem, st, p);
  NPDE_CHECK_LAUNCH();
  return NPDE_OK;
}

template <int L>
int launch_conv(ConvStreamParams &p, cudaStream_t st) {
  constexpr int R = L + 1, RC = (R + 1) & ~1, DS = (2 * L + 1 <= 8) ? 8 : 16;
  const size_t smem = sizeof(float4) * DS * 32 * (npde::STREAM_THREADS / 32);
  static int cache = 0;
  int resident = resident_warps_cached((const void *)k_conv_stream<L, NPDE_ACT_CLAMP, false>, smem, &cache);
  Split s = make_split(p.B, p.N, RC, R, resident);
  p.strips = s.strips; p.bands = s.bands; p.band_rows = s.band_rows; p.wv = s.wv;
  long items = (long)p.B * s.strips * s.bands;
  const int wpb = npde::STREAM_THREADS / 32;
  dim3 grid((unsigned)((items + wpb - 1) / wpb)), block(npde::STREAM_THREADS);
  switch (p.act) {
    case NPDE_ACT_NONE: return launch_conv_f<L, NPDE_ACT_NONE>(p, grid, block, smem, st);
    case NPDE_ACT_CLAMP: return launch_conv_f<L, NPDE_ACT_CLAMP>(p, grid, block, smem, st);
    case NPDE_ACT_SIGMOID: return launch_conv_f<L, NPDE_ACT_SIGMOID>(p, grid, block, smem, st);
    default: return NPDE_ERR_ACT;
  }
}

template <int T>
int launch_jacobi(JacobiStreamParams &p, cudaStream_t st) {
  constexpr int RC = (T + 1) & ~1;
  static int cache = 0;
  int resident = resident_warps_cached((const void *)k_jacobi_stream<T>, 0, &cache);
  Split s = make_split(p.B, p.N, RC, T, resident);
  p.strips = s.strips; p.bands = s.bands; p.band_rows = s.band_rows; p.wv = s.wv;
  long items = (long)p.B * s.strips * s.bands;
  const int wpb = npde::STREAM_THREADS / 32;
  dim3 grid((unsigned)((items + wpb - 1) / wpb)), block(npde::STREAM_THREADS);
  NPDE_LAUNCH((k_jacobi_stream<T>), grid, block, 0, st, p);
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
