// npde_stream.cu -- the hot loop: fused, register-streaming kernels for
//   * Phi_H = ConvIterator.forward (models/iterators/conv_iterator.py:33-51):
//       Psi -> residual -> L masked 3x3 convs -> +y -> act -> pad -> G   in ONE pass, and
//   * T temporally blocked Jacobi steps (utils/heat_utils.py:96-106 applied T times)
//     per HBM round trip.
//
// Design ("warp strips"):
//   - A warp owns a strip of 128 grid columns and marches down the rows of a band.  Lane l
//     holds the four adjacent columns c0..c3 = cs+4l .. cs+4l+3 as two packed pairs
//     Q0 = {c0, c2}, Q1 = {c1, c3}.  With this pairing the centre and the inner horizontal
//     taps of a 3-wide stencil are naturally aligned pairs (out Q0: centre Q0, right Q1;
//     out Q1: left Q0, centre Q1) and run as packed FFMA2 (fma.rn.f32x2, new on sm_100: two
//     IEEE-rn FMAs per issue slot); only the two outer taps (left of Q0 = {c-1, c1}, right of
//     Q1 = {c2, c4}) mix registers of different pairs and run as two scalar FFMAs on the
//     halves of the accumulator pair.  c-1 and c4 come from the neighbour lanes by ONE warp
//     shuffle each per row; nothing is exchanged between warps, so there is no
//     __syncthreads anywhere: a strip recomputes a halo of whole lanes instead.
//   - The internal ping-pong fields of the k-step loop use a padded pitch and store every
//     aligned group of four columns as (c0, c2, c1, c3) ("pair-permuted layout"): one 128-bit
//     load delivers Q0 and Q1 as register pairs, one 128-bit store writes a lane's outputs.
//     Reference-layout tensors (pitch 2^n+1, rows only 4-byte aligned) use 32-bit accesses.
//   - Each conv stage keeps three partial-sum rows that are updated in place and swap roles
//     every step, and the three live x rows rotate the same way; the step index modulo 3 is a template
//     parameters of a 3x unrolled loop, so the rotation costs no register moves.
//   - Conv stages are software-pipelined two rows apart so that all stages of a step are
//     independent instruction streams (ILP instead of occupancy).
//   - The Jacobi value y of a row is needed 2L steps later; it waits in a per-lane
//     shared-memory delay line (private slots, no synchronisation).
//   - Input rows are loaded one step ahead straight into their register slot and several
//     rows ahead into L2 (prefetch.global.L2).
//   - Arithmetic order is the oracle's: row-major fmaf chain, first tap a plain product
//     (oracle/npde_oracle.c conv3x3_img / orc_fd_step), so iterates are bit-identical to the
//     CPU restatement.
#include <atomic>

#include "npde_common.cuh"
#include "npde_internal.h"

#include "npde_f2.cuh"

namespace {

constexpr int STRIP = 128;        // columns per warp strip: lane l owns the 4 adjacent columns cs+4l .. cs+4l+3
constexpr int LANE_COLS = 4;
#ifndef NPDE_L2_AHEAD
#define NPDE_L2_AHEAD 4
#endif
constexpr int L2_AHEAD = NPDE_L2_AHEAD;  // rows of L2 prefetch distance (0: no prefetch instructions)
#ifndef NPDE_RING
#define NPDE_RING 8
#endif
// Rows of x kept in flight per warp by cp.async in the fast loops (a ring of 512-byte slots in
// shared memory; RING_D - 1 rows are under way while one is consumed).  0: plain loads + L2 prefetch.
constexpr int RING_D = NPDE_RING;
static_assert((RING_D & (RING_D - 1)) == 0, "the ring size must be a power of two (slot offsets wrap with a mask)");
#ifndef NPDE_RING_PF
#define NPDE_RING_PF 0
#endif
constexpr int RING_PF = NPDE_RING_PF;  // additional L2 prefetch distance (rows beyond the ring), 0 = none

// Column bookkeeping of one lane: bit q <-> column c0 + q
// (bit0 = lo(Q0), bit1 = lo(Q1), bit2 = hi(Q0), bit3 = hi(Q1)).
struct LaneCols {
  int c0;
  unsigned in_img;    // column exists (0 <= c < N)
  unsigned interior;  // 1 <= c <= N-2
  unsigned store;     // column belongs to this strip's valid output range
};

__device__ __forceinline__ LaneCols make_cols(int cs, int lane, int N, int vs, int ve) {
  LaneCols c;
  c.c0 = cs + LANE_COLS * lane;
  c.in_img = c.interior = c.store = 0u;
#pragma unroll
  for (int q = 0; q < 4; ++q) {
    int col = c.c0 + q;
    if (col >= 0 && col < N) c.in_img |= 1u << q;
    if (col >= 1 && col <= N - 2) c.interior |= 1u << q;
    if (col >= vs && col < ve) c.store |= 1u << q;
  }
  return c;
}

// One grid row, this lane's four columns -> q0 = Q0 = {c0,c2}, q1 = Q1 = {c1,c3}.  p points at
// the lane's group of four.
// perm (pair-permuted padded layout, memory order c0,c2,c1,c3): one 128-bit load, enabled per lane.
// natural layout: four 32-bit loads with per-column predicates (any alignment).
__device__ __forceinline__ void load_row(const float *__restrict__ p, bool perm, bool lane_ok, unsigned in_img,
                                         f2 &q0, f2 &q1) {
  if (perm) {
    if (lane_ok) ld_pairs(p, q0, q1);
    else q0 = q1 = zero2();
  } else {
    const float v0 = (in_img & 1u) ? __ldg(p) : 0.0f;
    const float v1 = (in_img & 2u) ? __ldg(p + 1) : 0.0f;
    const float v2 = (in_img & 4u) ? __ldg(p + 2) : 0.0f;
    const float v3 = (in_img & 8u) ? __ldg(p + 3) : 0.0f;
    q0 = mk2(v0, v2);
    q1 = mk2(v1, v3);
  }
}

// o0 = {c0,c2}, o1 = {c1,c3}.  perm: one 128-bit store per enabled lane (a lane stores all of
// its columns or none: strip ranges are multiples of 4); natural: per-column predicates.
__device__ __forceinline__ void store_row(float *__restrict__ p, bool perm, bool lane_ok, unsigned mask, f2 o0,
                                          f2 o1) {
  if (perm) {
    if (lane_ok) st_pairs(p, o0, o1);
  } else {
    if (mask & 1u) p[0] = lo(o0);
    if (mask & 2u) p[1] = lo(o1);
    if (mask & 4u) p[2] = hi(o0);
    if (mask & 8u) p[3] = hi(o1);
  }
}

// same from the four column values c0..c3 (fresh scalars: ptxas allocates them as one quad)
__device__ __forceinline__ void store_cols(float *__restrict__ p, bool perm, bool lane_ok, unsigned mask, float v0,
                                           float v1, float v2, float v3) {
  if (perm) {
    if (lane_ok) *reinterpret_cast<float4 *>(p) = make_float4(v0, v2, v1, v3);
  } else {
    if (mask & 1u) p[0] = v0;
    if (mask & 2u) p[1] = v1;
    if (mask & 4u) p[2] = v2;
    if (mask & 8u) p[3] = v3;
  }
}

// L2 prefetch of the 512-byte row segment of a strip
#ifndef NPDE_PF_LANES
#define NPDE_PF_LANES 1  // every lane: the segment is not line-aligned (a sparser prefetch misses its tail line)
#endif
__device__ __forceinline__ void prefetch_row(const float *p, int lane, unsigned in_img) {
  if ((lane & (NPDE_PF_LANES - 1)) == 0 && (in_img & 1u)) prefetch_l2(p);
}

// The two outer horizontal neighbours of a row: c-1 = c3 of the left lane, c4 = c0 of the right
// lane.  The strip's outer edges (lane 0 left, lane 31 right) receive wrapped-around
// don't-care values; they only reach cells outside the strip's valid range.
__device__ __forceinline__ void halo_scalars(f2 q0, f2 q1, int lane_l, int lane_r, float &hl, float &hr) {
  hl = __shfl_sync(NPDE_FULL_MASK, hi(q1), lane_l);
  hr = __shfl_sync(NPDE_FULL_MASK, lo(q0), lane_r);
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
    float v0 = lo(o0), v1 = lo(o1), v2 = hi(o0), v3 = hi(o1);
    if (c.c0 == 0) v0 = left;
    if (c.c0 == N - 1) v0 = right;
    if (c.c0 + 1 == 0) v1 = left;
    if (c.c0 + 1 == N - 1) v1 = right;
    if (c.c0 + 2 == 0) v2 = left;
    if (c.c0 + 2 == N - 1) v2 = right;
    if (c.c0 + 3 == 0) v3 = left;
    if (c.c0 + 3 == N - 1) v3 = right;
    o0 = mk2(v0, v2);
    o1 = mk2(v1, v3);
  }
}

// Jacobi average of one row (centre pairs m0, m1 with outer neighbours hl, hr; rows above u*
// and below d*), taps in the oracle's order up, left, right, down (orc_fd_step).
__device__ __forceinline__ void psi_row(f2 u0, f2 u1, f2 m0, f2 m1, float hl, float hr, f2 d0, f2 d1, f2 &o0,
                                        f2 &o1) {
  const f2 q = bc2(0.25f);
  o0 = mul2(q, u0);
  o0 = fma_lohi(0.25f, hl, lo(m1), o0);  // left of {c0,c2} = {c-1, c1}
  o0 = fma2(q, m1, o0);                  // right of {c0,c2} = {c1, c3}
  o0 = fma2(q, d0, o0);
  o1 = mul2(q, u1);
  o1 = fma2(q, m0, o1);                  // left of {c1,c3} = {c0, c2}
  o1 = fma_lohi(0.25f, hi(m0), hr, o1);  // right of {c1,c3} = {c2, c4}
  o1 = fma2(q, d1, o1);
}

// Work partition: B * strips * bands warp items, the strip index fastest.  The warps that share a
// band of an image walk down the same rows at the same time, so their 512-byte row segments hit
// the same open DRAM pages (a partition into equal-cost pieces at unrelated rows measured 7% slower).
// One warp per CTA: the work item depends on blockIdx only, so ptxas can prove that the band /
// strip bookkeeping and the loop-versioning branches are warp-uniform and keeps the conv weights
// in uniform registers (FFMA2 R, R, UR.F32, R) instead of re-loading them from the constant bank.
struct StripItem {
  int b, strip, i0, i1;  // image, strip, rows [i0, i1)
};
struct Split {
  int strips, bands, band_rows, wv;
};
__host__ __device__ __forceinline__ bool strip_is_edge(int strip, int wv, int rc, int N) {
  const int cs = strip * wv - rc;
  return cs < 1 || cs + STRIP - 1 > N - 2;
}
__device__ __forceinline__ bool decode_item(int B, int N, const Split &sp, StripItem &it) {
  long item = (long)blockIdx.x;
  if (item >= (long)B * sp.strips * sp.bands) return false;
  it.strip = (int)(item % sp.strips);
  const int band = (int)((item / sp.strips) % sp.bands);
  it.b = (int)(item / ((long)sp.strips * sp.bands));
  it.i0 = band * sp.band_rows;
  it.i1 = min(it.i0 + sp.band_rows, N);
  return true;
}

// ------------------------------------------------------------ Phi_H (Conv) ---
constexpr int ACT_HEAD = 3;  // internal "activation": the kernel is the head of a U-Net step (see conv_step)

struct ConvStreamParams {
  const float *in;
  float *out;
  float *yout;       // ACT_HEAD only: second output (y = Psi(x) - f), same layout as out
  const float *f;    // NULL: Laplace
  const float *bc4;  // (B,4)
  long in_bstride, out_bstride, f_bstride;
  int in_pitch, out_pitch, f_pitch;
  int in_perm, out_perm, f_perm;  // field is in the pair-permuted padded layout (128-bit accesses)
  // mask geometry (bc = [values, mask], heat_utils.py:29-42): the two planes as field views, else NULL
  const float *bv, *bm;
  long bv_bstride, bm_bstride;
  int bv_pitch, bm_pitch, bv_perm, bm_perm;
  int B, N, act;
  Split split;
  float w[4][9];
};

__device__ __forceinline__ float sel_bits(float v, unsigned keep, unsigned val) {
  return __uint_as_float((__float_as_uint(v) & keep) | val);  // one LOP3
}

// iterator.py:17-27 applied to z = y + h.  Sigmoid is kept out of line: it is rare and its
// expf/division expansion would otherwise sit in the middle of the streaming loop.
__device__ __noinline__ float sigmoid_slow(float z) { return 1.0f / (1.0f + expf(-z)); }
template <int ACT>
__device__ __forceinline__ float add_act(float y, float h) {
  if (ACT == NPDE_ACT_CLAMP) {
#ifdef NPDE_EMU
    return fminf(fmaxf(__fadd_rn(y, h), 0.0f), 1.0f);
#else
    float z;  // one FADD.SAT: round(y + h) clamped to [0, 1]
    asm("add.rn.sat.f32 %0, %1, %2;" : "=f"(z) : "f"(y), "f"(h));
    return z;
#endif
  }
  const float z = __fadd_rn(y, h);
  if (ACT == NPDE_ACT_SIGMOID) return sigmoid_slow(z);
  return z;
}

// Streaming form of the row-major 3x3 chain.  A row of the stage input (centre pairs c0p, c1p
// and its outer neighbours hl, hr) arrives once and serves three output rows: it FINISHES the
// chain of the row above it (taps w6..w8), CONTINUES the chain of its own row (w3..w5) and
// STARTS the chain of the row below (w0..w2, first tap a plain product).  Per output the
// operation order is exactly w0..w8 (oracle conv3x3_img), so results equal the windowed form
// bit for bit, but only partial sums stay live instead of a 3-row window.  The three
// partial-sum rows of a stage are updated IN PLACE and swap roles every step
// (fin -> top -> mid -> fin).
__device__ __forceinline__ void conv_taps(f2 c0p, f2 c1p, float hl, float hr, float wl, float wc, float wr,
                                          f2 &a0, f2 &a1) {
  a0 = fma_lohi(wl, hl, lo(c1p), a0);  // out {c0,c2}: left {c-1,c1}, centre {c0,c2}, right {c1,c3}
  a0 = fma2(bc2(wc), c0p, a0);
  a0 = fma2(bc2(wr), c1p, a0);
  a1 = fma2(bc2(wl), c0p, a1);         // out {c1,c3}: left {c0,c2}, centre {c1,c3}, right {c2,c4}
  a1 = fma2(bc2(wc), c1p, a1);
  a1 = fma_lohi(wr, hi(c0p), hr, a1);
}
template <int L>
struct ConvState {
  f2 X[2][2];      // X[P]: x row `it` (loaded during the previous step), X[P^1]: row it-1 (centre row of Psi)
  float xhl, xhr;  // outer neighbours of x row it-1
  f2 yu[2];        // 0.25 * x[it-2]: first tap of Psi for output row it-1
  f2 S[L][2][2];   // S[k-1]: the two live partial-sum rows of conv stage k.  S[k-1][P] has seen taps
                   // w0..w5 of its output row (this step finishes it), S[k-1][P^1] taps w0..w2.
  f2 pff[2];       // f row it-1, fetched one step earlier
  f2 pm[2];        // mask geometry: mask row it-1, fetched one step earlier
  f2 pv[2][2];     // mask geometry: pv[P] = value row it-1-L (this step's output row), pv[P^1] is being fetched
};

struct ConvLaneCtx {
  unsigned keep[4];  // per column c0..c3: all ones if interior, else 0
  unsigned gval[4];  // Dirichlet value bits of a ring column, else 0
  unsigned in_img, store;  // LaneCols bit masks
  bool in_perm, out_perm, f_perm;  // layouts (uniform)
  bool ld_ok, st_ok; // permuted layout: this lane's 128-bit load / store is enabled
  const float *pin;  // this lane's group of four in the NEXT input row to fetch
  const float *pfin; // same for f
  const float *pmin, *pvin;  // mask geometry: next mask row / value row to fetch
  float *pout;       // this lane's group of four in the output row of the current step
  float *pyout;      // ACT_HEAD: same for the y output
  char *delay;       // this lane's slot 0 of the y delay line (shared memory)
  char *ring;        // this lane's 16 bytes in slot 0 of the x row ring (shared memory)
  unsigned rslot;    // ring slot (byte offset) that holds x row it+1
  unsigned nat_off;  // reference-layout ring fill: byte offset of column cs + lane inside a slot
  unsigned nat_img;  // ... and bit k = column cs + lane + 32k exists
  float top, bottom;
  int N, i0, i1, x_last, f_last, in_pitch, out_pitch, f_pitch, bm_pitch, bv_pitch, lane, lane_l, lane_r;
  bool bm_perm, bv_perm;
};

#ifndef NPDE_EMU
// Keeps a loop-invariant per-lane constant in a register: without this ptxas
// rematerialises the column masks from the column index in every step (2-3
// instructions each) to save registers.
__device__ __forceinline__ void pin_reg(unsigned &v) { asm volatile("" : "+r"(v)); }
#else
__device__ __forceinline__ void pin_reg(unsigned &) {}
#endif

// Start the asynchronous copy of one row segment (this strip's 128 columns) into a ring slot, which always
// holds the pair-permuted order.  p = this lane's group of four in that row.
//   permuted field:  one 16-byte copy per lane;
//   reference layout (rows only 4-byte aligned): four coalesced 4-byte copies per lane -- lane l moves columns
//   cs + l + 32k, k = 0..3, each straight to its permuted position; out-of-image columns are zero-filled.
__device__ __forceinline__ void ring_fill(char *slot_lane0 /* slot base + 16 * lane */, const float *p, bool perm,
                                          bool ld_ok, const ConvLaneCtx &x) {
  if (perm) {
    if (ld_ok) cp_async16(slot_lane0, p);
  } else {
    // lanes write each other's bytes here: every lane must be done reading the slot's previous row
    // (converged warps are; the barrier makes it explicit)
    __syncwarp();
    char *slot = slot_lane0 - 16 * x.lane;
    const float *row = p - 3 * x.lane;  // column cs + lane
#pragma unroll
    for (int k = 0; k < 4; ++k)
      cp_async4_zfill(slot + x.nat_off + 128 * k, row + 32 * k, (x.nat_img >> k) & 1u);
  }
}

// One step: input row `it` has arrived in s.X[P].  Stage 0 (Psi, residual) works on row it-1 and
// conv stage k (1..L) on row it-1-k, each consuming the row the previous stage produced in THIS
// step, so no finished rows stay live across steps; latency between the stages is covered by
// the other warps of the scheduler.
// MAIN:   every stage row (it-1-k, k=0..L) is an interior row of the grid, the output row
//         belongs to this band and the rows to fetch exist: no row tests at all.
// !MAIN:  generic step (band warm-up / drain, ring rows): uniform row tests everywhere.
// EDGE:   the strip contains ring or out-of-grid columns: loads are predicated per lane or
//         column, stage outputs are masked to the interior columns and G fixes the ring columns.
// PERM:   all fields are in the pair-permuted padded layout (128-bit accesses); otherwise the
//         layout of each field is a run-time flag.
// P:      step parity.  The two x row slots and the two partial-sum rows of every stage swap
//         roles each step, so a 2x unrolled loop needs no register moves and its body
//         (~260 instructions) stays inside the 6 KB L0 instruction cache.
// DS = delay-line slots (power of two >= L+1), 512 bytes per slot and warp.
// RINGED: (fast loops on permuted fields only) x rows arrive through the cp.async ring.
// MASK:   mask geometry (L-shape, cylinders, ...): bc = [values v, mask m], m in {0,1}.  G is
//         x*(1-m)+v at every cell and every stage output is multiplied by the foreground 1-m
//         (conv_iterator.py:23-30).  Per row the kernel builds the AND-masks "interior cell and
//         m == 0" once (stage 0) and keeps them L steps in a second shared-memory delay line; they
//         replace the column masks of EDGE (MASK implies EDGE) and the ring-row tests.
// ACT == ACT_HEAD: head of a U-Net step (unet_iterator.py:87-97 + the first level's pre-smoothing convs):
//         instead of act(y + h) and G the epilogue stores r_L (masked like every stage output) to `out`
//         and y to `yout`, both as frames whose ring is never read.
template <int L, int ACT, bool HASF, bool MASK, bool MAIN, bool EDGE, bool PERM, int P, bool RINGED = false>
__device__ __forceinline__ void conv_step(ConvState<L> &s, ConvLaneCtx &x, const ConvStreamParams &p, int it,
                                          unsigned &woff) {
  constexpr int DS = (L + 1 <= 4) ? 4 : 8;
  constexpr unsigned RING = DS * 512u - 1u;
  constexpr unsigned KEEP_OFF = DS * 512u;  // the keep-mask delay line sits behind the y delay line (MASK only)
  const int N = x.N;
  const unsigned roff = (woff + (DS - L) * 512u) & RING;  // y written L steps ago
  const bool in_perm = PERM || x.in_perm, out_perm = PERM || x.out_perm, f_perm = PERM || x.f_perm;
  const bool ld_ok = !EDGE || x.ld_ok;
  f2 rc0, rc1;    // the row travelling down the stages: centre pairs ...
  float rl, rr;   // ... and outer neighbours

  // ---- stage 0: y = Psi(x) - f and r_0 = y - x for row it-1 ------------------------------
  {
    const f2 (&D)[2] = s.X[P];      // row it: the "down" tap
    f2 (&C)[2] = s.X[P ^ 1];        // centre row it-1
    const int r0 = it - 1;
    const f2 q = bc2(0.25f);
    f2 y0 = fma_lohi(0.25f, s.xhl, lo(C[1]), s.yu[0]);  // taps: up (in yu), left, right, down
    y0 = fma2(q, C[1], y0);
    y0 = fma2(q, D[0], y0);
    f2 y1 = fma2(q, C[0], s.yu[1]);
    y1 = fma_lohi(0.25f, hi(C[0]), s.xhr, y1);
    y1 = fma2(q, D[1], y1);
    if (HASF) {
      y0 = sub2(y0, s.pff[0]);
      y1 = sub2(y1, s.pff[1]);
      const int nf = it;  // f row used by the stage 0 of the next step
      if (RINGED) {
        // x.pfin points at f row it + RING_D - 1: it joins the copy group of x row it + RING_D below
        if (it + RING_D - 1 <= x.f_last) {
          constexpr unsigned RMASK_F = (RING_D > 0 ? RING_D : 1) * 512u - 1u;
          ring_fill(x.ring + RING_D * 512 + ((x.rslot + (RING_D - 1) * 512u) & RMASK_F), x.pfin, f_perm, ld_ok, x);
          x.pfin += x.f_pitch;
        }
      } else {
        if (MAIN || (nf <= x.x_last && nf >= 1 && nf <= N - 2)) {
          load_row(x.pfin, f_perm, ld_ok, EDGE ? x.in_img : 15u, s.pff[0], s.pff[1]);
        } else {
          s.pff[0] = s.pff[1] = zero2();
        }
        if (L2_AHEAD > 0 && nf + L2_AHEAD <= x.x_last)
          prefetch_row(x.pfin + (size_t)L2_AHEAD * x.f_pitch, x.lane, EDGE ? x.in_img : 15u);
        x.pfin += x.f_pitch;
      }
    }
    rc0 = sub2(y0, C[0]);
    rc1 = sub2(y1, C[1]);
    if (MASK) {
      // keep masks of row it-1: interior cell (row and column) and not a Dirichlet cell
      const bool row_in = MAIN || (r0 >= 1 && r0 <= N - 2);
      uint4 kq;
      kq.x = (row_in && lo(s.pm[0]) == 0.0f) ? x.keep[0] : 0u;
      kq.y = (row_in && lo(s.pm[1]) == 0.0f) ? x.keep[1] : 0u;
      kq.z = (row_in && hi(s.pm[0]) == 0.0f) ? x.keep[2] : 0u;
      kq.w = (row_in && hi(s.pm[1]) == 0.0f) ? x.keep[3] : 0u;
      *reinterpret_cast<uint4 *>(x.delay + KEEP_OFF + woff + 8 * x.lane) = kq;  // 16 bytes per lane and slot
      rc0 = and2(rc0, kq.x, kq.z);
      rc1 = and2(rc1, kq.y, kq.w);
      // fetch mask row `it` (next step) and value row it-L (epilogue row of the next step)
      const int nm = it, nv = it - L;
      if (MAIN || (nm >= 1 && nm <= min(x.x_last, N - 2))) {
        load_row(x.pmin, PERM || x.bm_perm, ld_ok, x.in_img, s.pm[0], s.pm[1]);
        if (L2_AHEAD > 0 && nm + L2_AHEAD <= x.x_last)
          prefetch_row(x.pmin + (size_t)L2_AHEAD * x.bm_pitch, x.lane, x.in_img);
      } else {
        s.pm[0] = s.pm[1] = bc2(1.0f);
      }
      x.pmin += x.bm_pitch;
      if (MAIN || (nv >= 0 && nv <= N - 1 && nv < x.i1)) {
        load_row(x.pvin, PERM || x.bv_perm, ld_ok, x.in_img, s.pv[P ^ 1][0], s.pv[P ^ 1][1]);
        if (L2_AHEAD > 0 && nv + L2_AHEAD < x.i1) prefetch_row(x.pvin + (size_t)L2_AHEAD * x.bv_pitch, x.lane, x.in_img);
      } else {
        s.pv[P ^ 1][0] = s.pv[P ^ 1][1] = zero2();
      }
      if (MAIN || nv >= 0) x.pvin += x.bv_pitch;  // invariant: pvin points at row max(nv + 1, 0)
    } else {
      if (!MAIN && !(r0 >= 1 && r0 <= N - 2)) rc0 = rc1 = zero2();
      if (EDGE) {
        rc0 = and2(rc0, x.keep[0], x.keep[2]);
        rc1 = and2(rc1, x.keep[1], x.keep[3]);
      }
    }
    halo_scalars(rc0, rc1, x.lane_l, x.lane_r, rl, rr);
    // y waits L steps in the delay line; the two pairs go to separate 256-byte halves of the slot so
    // that they stay independent 64-bit stores (one 128-bit store would need them in one aligned quad)
    *reinterpret_cast<f2 *>(x.delay + woff) = y0;
    *reinterpret_cast<f2 *>(x.delay + woff + 256) = y1;
    s.yu[0] = mul2(q, C[0]);  // first tap of the next output row (row it): up = x[it-1]
    s.yu[1] = mul2(q, C[1]);
    halo_scalars(D[0], D[1], x.lane_l, x.lane_r, s.xhl, s.xhr);  // row `it` is the next centre row
    // ---- row it-1 is dead: fetch x row it+1 into its slot (consumed by the next step) ----
    const int nx = it + 1;
    if (RINGED) {
      // x.pin points at row it + RING_D: start its copy into the slot that was read one step ago
      constexpr unsigned RMASK = (RING_D > 0 ? RING_D : 1) * 512u - 1u;
      if (it + RING_D <= x.x_last) {
        ring_fill(x.ring + ((x.rslot + (RING_D - 1) * 512u) & RMASK), x.pin, in_perm, ld_ok, x);
        if (RING_PF > 0 && it + RING_D + RING_PF <= x.x_last)
          prefetch_row(x.pin + (size_t)RING_PF * x.in_pitch, x.lane, EDGE ? x.in_img : 15u);
        x.pin += x.in_pitch;
      }
      cp_async_commit();
    } else if (MAIN || nx <= x.x_last) {
      load_row(x.pin, in_perm, ld_ok, EDGE ? x.in_img : 15u, C[0], C[1]);
      if (L2_AHEAD > 0 && nx + L2_AHEAD <= x.x_last)  // never prefetch past the rows this band may read
        prefetch_row(x.pin + (size_t)L2_AHEAD * x.in_pitch, x.lane, EDGE ? x.in_img : 15u);
      x.pin += x.in_pitch;
    } else {
      C[0] = C[1] = zero2();
    }
  }

  // ---- conv stages 1 .. L on rows it-2 .. it-1-L ------------------------------------------
#pragma unroll
  for (int k = 1; k <= L; ++k) {
    const int rk = it - 1 - k;
    const float *w = p.w[k - 1];
    f2 (&mid)[2] = s.S[k - 1][P];      // taps w0..w5 done: finished by this row
    f2 (&top)[2] = s.S[k - 1][P ^ 1];  // taps w0..w2 done
    f2 fin0 = mid[0], fin1 = mid[1];
    conv_taps(rc0, rc1, rl, rr, w[6], w[7], w[8], fin0, fin1);
    conv_taps(rc0, rc1, rl, rr, w[3], w[4], w[5], top[0], top[1]);  // becomes the "mid" row of the next step
    mid[0] = mul_lohi(w[0], rl, lo(rc1));                            // fresh "top" row of the next step
    mid[0] = fma2(bc2(w[1]), rc0, mid[0]);
    mid[0] = fma2(bc2(w[2]), rc1, mid[0]);
    mid[1] = mul2(bc2(w[0]), rc0);
    mid[1] = fma2(bc2(w[1]), rc1, mid[1]);
    mid[1] = fma_lohi(w[2], hi(rc0), rr, mid[1]);
    if (k < L) {
      if (MASK) {  // keep masks of row it-1-k, built k steps ago
        const uint4 kq = *reinterpret_cast<const uint4 *>(x.delay + KEEP_OFF + 8 * x.lane +
                                                          ((woff + (DS - k) * 512u) & RING));
        fin0 = and2(fin0, kq.x, kq.z);
        fin1 = and2(fin1, kq.y, kq.w);
      } else {
        if (!MAIN && !(rk >= 1 && rk <= N - 2)) fin0 = fin1 = zero2();
        if (EDGE) {
          fin0 = and2(fin0, x.keep[0], x.keep[2]);
          fin1 = and2(fin1, x.keep[1], x.keep[3]);
        }
      }
      rc0 = fin0;
      rc1 = fin1;
      halo_scalars(rc0, rc1, x.lane_l, x.lane_r, rl, rr);
    } else if (ACT == ACT_HEAD) {
      if (MAIN || (rk >= x.i0 && rk < x.i1)) {
        if (MASK) {
          const uint4 kq = *reinterpret_cast<const uint4 *>(x.delay + KEEP_OFF + 8 * x.lane + roff);
          fin0 = and2(fin0, kq.x, kq.z);
          fin1 = and2(fin1, kq.y, kq.w);
        } else {
          if (!MAIN && !(rk >= 1 && rk <= N - 2)) fin0 = fin1 = zero2();
          if (EDGE) {
            fin0 = and2(fin0, x.keep[0], x.keep[2]);
            fin1 = and2(fin1, x.keep[1], x.keep[3]);
          }
        }
        const f2 ya = *reinterpret_cast<const f2 *>(x.delay + roff);
        const f2 yb = *reinterpret_cast<const f2 *>(x.delay + roff + 256);
        store_row(x.pout, out_perm, x.st_ok, EDGE ? x.store : 15u, fin0, fin1);
        store_row(x.pyout, out_perm, x.st_ok, EDGE ? x.store : 15u, ya, yb);
      }
      if (MAIN || rk >= 0) {
        x.pout += x.out_pitch;
        x.pyout += x.out_pitch;
      }
    } else {
      if (MAIN || (rk >= x.i0 && rk < x.i1)) {
        float o0, o1, o2, o3;  // columns c0..c3
        if (MASK) {
          // G on the mask geometry: pad(act(y + h)) * (1 - m) + v at every cell (heat_utils.py:51-53);
          // m is 0 or 1, so the product is a select: keep == 0 on ring cells and Dirichlet cells
          const f2 ya = *reinterpret_cast<const f2 *>(x.delay + roff);
          const f2 yb = *reinterpret_cast<const f2 *>(x.delay + roff + 256);
          const uint4 kq = *reinterpret_cast<const uint4 *>(x.delay + KEEP_OFF + 8 * x.lane + roff);
          o0 = __fadd_rn(sel_bits(add_act<ACT>(lo(ya), lo(fin0)), kq.x, 0u), lo(s.pv[P][0]));
          o2 = __fadd_rn(sel_bits(add_act<ACT>(hi(ya), hi(fin0)), kq.z, 0u), hi(s.pv[P][0]));
          o1 = __fadd_rn(sel_bits(add_act<ACT>(lo(yb), lo(fin1)), kq.y, 0u), lo(s.pv[P][1]));
          o3 = __fadd_rn(sel_bits(add_act<ACT>(hi(yb), hi(fin1)), kq.w, 0u), hi(s.pv[P][1]));
        } else if (MAIN || (rk >= 1 && rk <= N - 2)) {
          const f2 ya = *reinterpret_cast<const f2 *>(x.delay + roff);
          const f2 yb = *reinterpret_cast<const f2 *>(x.delay + roff + 256);
          o0 = add_act<ACT>(lo(ya), lo(fin0));
          o2 = add_act<ACT>(hi(ya), hi(fin0));
          o1 = add_act<ACT>(lo(yb), lo(fin1));
          o3 = add_act<ACT>(hi(yb), hi(fin1));
        } else {
          o0 = o1 = o2 = o3 = (rk == 0 ? x.top : x.bottom);  // ring row: pad then G
        }
        if (EDGE && !MASK) {  // pad + G on the columns: non-interior columns take the Dirichlet value
          o0 = sel_bits(o0, x.keep[0], x.gval[0]);
          o1 = sel_bits(o1, x.keep[1], x.gval[1]);
          o2 = sel_bits(o2, x.keep[2], x.gval[2]);
          o3 = sel_bits(o3, x.keep[3], x.gval[3]);
        }
        store_cols(x.pout, out_perm, x.st_ok, EDGE ? x.store : 15u, o0, o1, o2, o3);
      }
      if (MAIN || rk >= 0) x.pout += x.out_pitch;  // invariant: pout points at row max(rk, 0)
    }
  }
  if (RINGED) {  // row it+1 has landed (at most RING_D - 1 younger copies pending): move it to registers
    constexpr unsigned RMASK = (RING_D > 0 ? RING_D : 1) * 512u - 1u;
    cp_async_wait<(RING_D > 0 ? RING_D - 1 : 0)>();
    if (!PERM) __syncwarp();  // reference-layout fills: a lane's bytes were copied by four other lanes
    // two 64-bit reads: the pairs stay independent registers (a 128-bit read lands in a quad and is
    // then copied into the two row slots)
    s.X[P ^ 1][0] = *reinterpret_cast<const f2 *>(x.ring + x.rslot);
    s.X[P ^ 1][1] = *reinterpret_cast<const f2 *>(x.ring + x.rslot + 8);
    if (HASF) {  // f row `it` travelled in the same copy group
      s.pff[0] = *reinterpret_cast<const f2 *>(x.ring + RING_D * 512 + x.rslot);
      s.pff[1] = *reinterpret_cast<const f2 *>(x.ring + RING_D * 512 + x.rslot + 8);
    }
    x.rslot = (x.rslot + 512u) & RMASK;
  }
  woff = (woff + 512u) & RING;
}

#ifndef NPDE_CONV_MIN_CTAS
#define NPDE_CONV_MIN_CTAS 16
#endif
template <int L, int ACT, bool HASF, bool MASK>
__global__ void __launch_bounds__(npde::STREAM_THREADS, NPDE_CONV_MIN_CTAS)
k_conv_stream(const NPDE_GRID_CONSTANT ConvStreamParams p) {
  constexpr int R = L + 1;           // dependency radius of Phi_H
  constexpr int RC = (R + 3) & ~3;   // column halo: whole lanes
  constexpr int DS = (L + 1 <= 4) ? 4 : 8;
  NPDE_DYN_SMEM(float4, ydl);        // [DS] slots of 512 bytes: two 256-byte halves, 8 bytes per lane each

  StripItem item;
  if (!decode_item(p.B, p.N, p.split, item)) return;
  ConvLaneCtx x;
  const int lane = threadIdx.x & 31;
  x.lane = lane;
  x.lane_l = (lane + 31) & 31;
  x.lane_r = (lane + 1) & 31;
  x.delay = reinterpret_cast<char *>(ydl) + 8 * lane;
  x.ring = reinterpret_cast<char *>(ydl) + (MASK ? 2 : 1) * DS * 512 + 16 * lane;  // behind the delay line(s)
  x.rslot = 0;
  x.N = p.N;
  const int N = p.N;
  const int vs = item.strip * p.split.wv, ve = min(vs + p.split.wv, N);
  const int cs = vs - RC;
  const LaneCols c = make_cols(cs, lane, N, vs, ve);
  x.in_img = c.in_img;
  x.store = c.store;
  pin_reg(x.in_img);
  pin_reg(x.store);
  x.in_perm = p.in_perm != 0;
  x.out_perm = p.out_perm != 0;
  x.f_perm = p.f_perm != 0;
  x.ld_ok = c.c0 >= 0 && c.c0 < N;  // c0 is a multiple of 4 and a permuted pitch is >= round_up(N, 4)
  {
    // column cs + lane (+ 32k) lives in lane (lane / 4 + 8k)'s group, at the permuted position of lane % 4
    const int q = lane & 3;
    x.nat_off = (unsigned)((lane >> 2) * 16 + (((q & 1) << 1) | (q >> 1)) * 4);
    x.nat_img = 0u;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const int col = cs + lane + 32 * k;
      if (col >= 0 && col < N) x.nat_img |= 1u << k;
    }
  }
  x.st_ok = c.store != 0u;          // strip ranges are multiples of 4: a lane stores all of its columns or none
  x.i0 = item.i0;
  x.i1 = item.i1;
  x.in_pitch = p.in_pitch;
  x.out_pitch = p.out_pitch;
  x.f_pitch = p.f_pitch;
  float left = 0.0f, right = 0.0f;
  x.top = x.bottom = 0.0f;
  if (!MASK) {
    const float *bcb = p.bc4 + (size_t)item.b * 4;
    x.top = __ldg(bcb + 0);
    x.bottom = __ldg(bcb + 1);
    left = __ldg(bcb + 2);
    right = __ldg(bcb + 3);
  }
  x.bm_pitch = p.bm_pitch;
  x.bv_pitch = p.bv_pitch;
  x.bm_perm = p.bm_perm != 0;
  x.bv_perm = p.bv_perm != 0;
#pragma unroll
  for (int q = 0; q < 4; ++q) {
    const int col = c.c0 + q;
    x.keep[q] = (c.interior >> q & 1u) ? 0xffffffffu : 0u;
    x.gval[q] = col == 0 ? __float_as_uint(left) : (col == N - 1 ? __float_as_uint(right) : 0u);
    pin_reg(x.keep[q]);
    pin_reg(x.gval[q]);
  }
  // the strip needs the EDGE code iff one of its 128 columns is not an interior column
  const bool edge_strip = MASK || strip_is_edge(item.strip, p.split.wv, RC, N);
  const bool all_perm = x.in_perm && x.out_perm && (!HASF || x.f_perm) && (!MASK || (x.bm_perm && x.bv_perm));

  ConvState<L> s;
#pragma unroll
  for (int h = 0; h < 2; ++h) {
    s.X[h][0] = s.X[h][1] = zero2();
    s.yu[h] = s.pff[h] = zero2();
    s.pm[h] = bc2(1.0f);
    s.pv[h][0] = s.pv[h][1] = zero2();
#pragma unroll
    for (int k = 0; k < L; ++k) s.S[k][h][0] = s.S[k][h][1] = zero2();
  }
  s.xhl = s.xhr = 0.0f;
#pragma unroll
  for (int d = 0; d < (MASK ? 2 : 1) * DS * 2; ++d) *reinterpret_cast<f2 *>(x.delay + d * 256) = zero2();
#pragma unroll
  for (int d = 0; d < (HASF ? 2 : 1) * RING_D; ++d)
    *reinterpret_cast<float4 *>(x.ring + d * 512) = make_float4(0.f, 0.f, 0.f, 0.f);

  // input rows it = it_first .. it_last; the output row of a step is it - 1 - L
  const int it_first = max(x.i0 - R, 0);  // rows above the grid are zeros == the initial state
  const int it_last = x.i1 + L;
  x.x_last = min(x.i1 - 1 + R, N - 1);    // last input row that exists and matters
  x.f_last = min(x.x_last, N - 2);        // last row of f that is ever read
  const float *xin = p.in + (size_t)item.b * p.in_bstride + c.c0;
  const float *fimg = HASF ? p.f + (size_t)item.b * p.f_bstride + c.c0 : nullptr;
  // slot 0 receives row it_first (the first step runs with P = 0); f: one row behind
  load_row(xin + (size_t)it_first * p.in_pitch, x.in_perm, x.ld_ok, c.in_img, s.X[0][0], s.X[0][1]);
  if (HASF && it_first - 1 >= 1)
    load_row(fimg + (size_t)(it_first - 1) * p.f_pitch, x.f_perm, x.ld_ok, c.in_img, s.pff[0], s.pff[1]);
  // running pointers: pin -> row it+1 of x, pfin -> row it of f, pout -> output row it-1-L (clamped at 0)
  x.pin = xin + (size_t)(it_first + 1) * p.in_pitch;
  x.pfin = HASF ? fimg + (size_t)it_first * p.f_pitch : nullptr;
  x.pmin = x.pvin = nullptr;
  if (MASK) {
    const float *mimg = p.bm + (size_t)item.b * p.bm_bstride + c.c0;
    const float *vimg = p.bv + (size_t)item.b * p.bv_bstride + c.c0;
    if (it_first - 1 >= 1)
      load_row(mimg + (size_t)(it_first - 1) * p.bm_pitch, x.bm_perm, x.ld_ok, c.in_img, s.pm[0], s.pm[1]);
    x.pmin = mimg + (size_t)it_first * p.bm_pitch;               // mask row `it`
    x.pvin = vimg + (size_t)max(it_first - L, 0) * p.bv_pitch;   // value row max(it - L, 0)
  }
  x.pout = p.out + (size_t)item.b * p.out_bstride + (size_t)max(it_first - 1 - L, 0) * p.out_pitch + c.c0;
  x.pyout = ACT == ACT_HEAD ? p.yout + (size_t)item.b * p.out_bstride +
                                  (size_t)max(it_first - 1 - L, 0) * p.out_pitch + c.c0
                            : nullptr;
  unsigned woff = 0;

  // MAIN range of `it`: stage rows it-1-L .. it-1 interior, output row it-1-L inside the band,
  // x row it+1 exists and f row `it` is an interior row
  const int main_lo = max(L + 2, x.i0 + L + 1);
  const int main_hi = min(N - 2, x.x_last - 1);
  int it = it_first;
  while (it <= it_last) {
    if (it >= main_lo && it + 1 <= main_hi) {
      // the fast loops iterate in place: the register assignment of a variant is set up once per band
#define NPDE_MAIN_LOOP(EDGE_, PERM_, RINGED_)                                            \
  do {                                                                                   \
    conv_step<L, ACT, HASF, MASK, true, EDGE_, PERM_, 0, RINGED_>(s, x, p, it, woff);    \
    conv_step<L, ACT, HASF, MASK, true, EDGE_, PERM_, 1, RINGED_>(s, x, p, it + 1, woff);\
    it += 2;                                                                             \
  } while (it + 1 <= main_hi)
      if (RING_D > 0) {
        // fill the ring: x row `it` is in registers, rows it+1 .. it+RING_D-1 start now
        x.rslot = 0;
#pragma unroll
        for (int d = 0; d < RING_D - 1; ++d) {
          if (it + 1 + d <= x.x_last) ring_fill(x.ring + d * 512, x.pin + (size_t)d * x.in_pitch, x.in_perm, x.ld_ok, x);
          if (HASF && it + d <= x.f_last)  // f runs one row behind x
            ring_fill(x.ring + (RING_D + d) * 512, x.pfin + (size_t)d * x.f_pitch, x.f_perm, x.ld_ok, x);
          cp_async_commit();
        }
        x.pin += (size_t)min(RING_D - 1, max(x.x_last - it, 0)) * x.in_pitch;
        if (HASF) x.pfin += (size_t)min(RING_D - 1, max(x.f_last - it + 1, 0)) * x.f_pitch;
        if (!all_perm) NPDE_MAIN_LOOP(true, false, true);  // any mix of layouts: run-time flags per field
        else if (edge_strip) NPDE_MAIN_LOOP(true, true, true);
        else if (!MASK) NPDE_MAIN_LOOP(false, true, true);
        cp_async_wait<0>();
        // back to plain loads: x.pin must point at row it+1 again; row `it` is in registers (slot of parity 0)
        x.pin = xin + (size_t)(it + 1) * p.in_pitch;
        if (HASF) x.pfin = fimg + (size_t)it * p.f_pitch;
      } else {
        if (!all_perm) NPDE_MAIN_LOOP(true, false, false);
        else if (edge_strip) NPDE_MAIN_LOOP(true, true, false);
        else if (!MASK) NPDE_MAIN_LOOP(false, true, false);
      }
#undef NPDE_MAIN_LOOP
    } else {
      conv_step<L, ACT, HASF, MASK, false, true, false, 0>(s, x, p, it, woff);
      if (it + 1 <= it_last) conv_step<L, ACT, HASF, MASK, false, true, false, 1>(s, x, p, it + 1, woff);
      it += 2;
    }
  }
}

// ------------------------------------------------- tail of a U-Net step ---
// unet_iterator.py:70-76 at the finest level + :99-105: the L post-smoothing convs applied to r (already
// multiplied by the foreground mask), + skip, then out = G(pad(act(y + h))).  Same strip / band streaming and
// conv chain as k_conv_stream, without Psi: row `it` of r arrives, stage k (1..L) finishes row it-k, the
// epilogue handles row it-L.  rin, skip and y are frames with one common layout (their ring is never used:
// every row is ANDed with the keep masks "interior cell [and m == 0]" on arrival).  One code path with
// uniform row tests (no MAIN / EDGE versions): this kernel replaces four launches of the layered path and is
// not the headline.
struct TailStreamParams {
  const float *rin, *skip, *y;
  float *out;
  const float *bc4;
  const float *bv, *bm;
  long in_bstride, out_bstride, bv_bstride, bm_bstride;
  int in_pitch, out_pitch, bv_pitch, bm_pitch;
  int in_perm, out_perm, bv_perm, bm_perm;
  int B, N;
  Split split;
  float w[4][9];
};

template <int L>
struct TailState {
  f2 S[L][2][2];  // partial-sum rows of the conv stages, as in ConvState
  f2 pr[2];       // r row `it` (fetched one step earlier)
  f2 ps[2], py[2], pv[2], pm[2];  // skip / y / value rows of the next output row, mask row of the next input row
};

struct TailCtx {
  unsigned keep[4], gval[4], in_img, store;
  bool in_perm, out_perm, bv_perm, bm_perm, ld_ok, st_ok;
  const float *prin, *pskip, *py, *pvin, *pmin;
  float *pout;
  char *kdelay;  // this lane's 16 bytes in slot 0 of the keep-mask delay line (MASK)
  float top, bottom;
  int N, i0, i1, in_pitch, out_pitch, bv_pitch, bm_pitch, lane, lane_l, lane_r;
};

template <int L, int ACT, bool MASK, int P>
__device__ __forceinline__ void tail_step(TailState<L> &s, TailCtx &x, const TailStreamParams &p, int it,
                                          unsigned &woff) {
  constexpr int DS = (L + 1 <= 4) ? 4 : 8;
  constexpr unsigned RING = DS * 512u - 1u;
  const int N = x.N;
  // ---- arrival of r row `it`: keep masks of that row, mask, halo ----
  const bool row_in = it >= 1 && it <= N - 2;
  uint4 kq;
  kq.x = (row_in && (!MASK || lo(s.pm[0]) == 0.0f)) ? x.keep[0] : 0u;
  kq.y = (row_in && (!MASK || lo(s.pm[1]) == 0.0f)) ? x.keep[1] : 0u;
  kq.z = (row_in && (!MASK || hi(s.pm[0]) == 0.0f)) ? x.keep[2] : 0u;
  kq.w = (row_in && (!MASK || hi(s.pm[1]) == 0.0f)) ? x.keep[3] : 0u;
  *reinterpret_cast<uint4 *>(x.kdelay + woff) = kq;
  f2 rc0 = and2(s.pr[0], kq.x, kq.z), rc1 = and2(s.pr[1], kq.y, kq.w);
  float rl, rr;
  halo_scalars(rc0, rc1, x.lane_l, x.lane_r, rl, rr);
  // fetch r row it+1 and (MASK) mask row it+1
  {
    const int nx = it + 1;
    if (nx <= N - 2 && nx <= x.i1 - 1 + L) {
      load_row(x.prin, x.in_perm, x.ld_ok, x.in_img, s.pr[0], s.pr[1]);
      if (MASK) load_row(x.pmin, x.bm_perm, x.ld_ok, x.in_img, s.pm[0], s.pm[1]);
      if (L2_AHEAD > 0 && nx + L2_AHEAD <= N - 2) {
        prefetch_row(x.prin + (size_t)L2_AHEAD * x.in_pitch, x.lane, x.in_img);
        if (MASK) prefetch_row(x.pmin + (size_t)L2_AHEAD * x.bm_pitch, x.lane, x.in_img);
      }
    } else {
      s.pr[0] = s.pr[1] = zero2();
      if (MASK) s.pm[0] = s.pm[1] = bc2(1.0f);
    }
    x.prin += x.in_pitch;
    if (MASK) x.pmin += x.bm_pitch;
  }
#pragma unroll
  for (int k = 1; k <= L; ++k) {
    const int rk = it - k;
    const float *w = p.w[k - 1];
    f2 (&mid)[2] = s.S[k - 1][P];
    f2 (&top)[2] = s.S[k - 1][P ^ 1];
    f2 fin0 = mid[0], fin1 = mid[1];
    conv_taps(rc0, rc1, rl, rr, w[6], w[7], w[8], fin0, fin1);
    conv_taps(rc0, rc1, rl, rr, w[3], w[4], w[5], top[0], top[1]);
    mid[0] = mul_lohi(w[0], rl, lo(rc1));
    mid[0] = fma2(bc2(w[1]), rc0, mid[0]);
    mid[0] = fma2(bc2(w[2]), rc1, mid[0]);
    mid[1] = mul2(bc2(w[0]), rc0);
    mid[1] = fma2(bc2(w[1]), rc1, mid[1]);
    mid[1] = fma_lohi(w[2], hi(rc0), rr, mid[1]);
    const uint4 kk = *reinterpret_cast<const uint4 *>(x.kdelay + ((woff + (DS - k) * 512u) & RING));  // row it-k
    fin0 = and2(fin0, kk.x, kk.z);
    fin1 = and2(fin1, kk.y, kk.w);
    if (k < L) {
      rc0 = fin0;
      rc1 = fin1;
      halo_scalars(rc0, rc1, x.lane_l, x.lane_r, rl, rr);
    } else {
      if (rk >= x.i0 && rk < x.i1) {
        // h = r + skip (:76), z = y + h, act, pad, G
        const float h0 = __fadd_rn(lo(fin0), lo(s.ps[0])), h2 = __fadd_rn(hi(fin0), hi(s.ps[0]));
        const float h1 = __fadd_rn(lo(fin1), lo(s.ps[1])), h3 = __fadd_rn(hi(fin1), hi(s.ps[1]));
        float o0 = add_act<ACT>(lo(s.py[0]), h0), o2 = add_act<ACT>(hi(s.py[0]), h2);
        float o1 = add_act<ACT>(lo(s.py[1]), h1), o3 = add_act<ACT>(hi(s.py[1]), h3);
        if (MASK) {
          o0 = __fadd_rn(sel_bits(o0, kk.x, 0u), lo(s.pv[0]));
          o2 = __fadd_rn(sel_bits(o2, kk.z, 0u), hi(s.pv[0]));
          o1 = __fadd_rn(sel_bits(o1, kk.y, 0u), lo(s.pv[1]));
          o3 = __fadd_rn(sel_bits(o3, kk.w, 0u), hi(s.pv[1]));
        } else {
          if (rk == 0 || rk == N - 1) o0 = o1 = o2 = o3 = (rk == 0 ? x.top : x.bottom);  // ring row: pad then G
          o0 = sel_bits(o0, x.keep[0], x.gval[0]);
          o1 = sel_bits(o1, x.keep[1], x.gval[1]);
          o2 = sel_bits(o2, x.keep[2], x.gval[2]);
          o3 = sel_bits(o3, x.keep[3], x.gval[3]);
        }
        store_cols(x.pout, x.out_perm, x.st_ok, x.store, o0, o1, o2, o3);
      }
      if (rk >= 0) x.pout += x.out_pitch;
      // skip / y / value rows of the next output row rk + 1
      const int nr = rk + 1;
      if (nr >= 0) {
        if (nr >= 1 && nr <= N - 2 && nr < x.i1) {
          load_row(x.pskip, x.in_perm, x.ld_ok, x.in_img, s.ps[0], s.ps[1]);
          load_row(x.py, x.in_perm, x.ld_ok, x.in_img, s.py[0], s.py[1]);
        } else {
          s.ps[0] = s.ps[1] = s.py[0] = s.py[1] = zero2();
        }
        x.pskip += x.in_pitch;
        x.py += x.in_pitch;
        if (MASK) {
          if (nr <= N - 1 && nr < x.i1) load_row(x.pvin, x.bv_perm, x.ld_ok, x.in_img, s.pv[0], s.pv[1]);
          else s.pv[0] = s.pv[1] = zero2();
          x.pvin += x.bv_pitch;
        }
      }
    }
  }
  woff = (woff + 512u) & RING;
}

template <int L, int ACT, bool MASK>
__global__ void __launch_bounds__(npde::STREAM_THREADS)
k_tail_stream(const NPDE_GRID_CONSTANT TailStreamParams p) {
  constexpr int RC = (L + 3) & ~3;
  constexpr int DS = (L + 1 <= 4) ? 4 : 8;
  NPDE_DYN_SMEM(float4, kdl);  // [DS][32] keep masks, 16 bytes per lane and slot
  StripItem item;
  if (!decode_item(p.B, p.N, p.split, item)) return;
  TailCtx x;
  const int lane = threadIdx.x & 31, N = p.N;
  x.lane = lane;
  x.lane_l = (lane + 31) & 31;
  x.lane_r = (lane + 1) & 31;
  x.N = N;
  x.kdelay = reinterpret_cast<char *>(kdl) + 16 * lane;
  const int vs = item.strip * p.split.wv, ve = min(vs + p.split.wv, N);
  const LaneCols c = make_cols(vs - RC, lane, N, vs, ve);
  x.in_img = c.in_img;
  x.store = c.store;
  x.in_perm = p.in_perm != 0; x.out_perm = p.out_perm != 0; x.bv_perm = p.bv_perm != 0; x.bm_perm = p.bm_perm != 0;
  x.ld_ok = c.c0 >= 0 && c.c0 < N;
  x.st_ok = c.store != 0u;
  x.i0 = item.i0; x.i1 = item.i1;
  x.in_pitch = p.in_pitch; x.out_pitch = p.out_pitch; x.bv_pitch = p.bv_pitch; x.bm_pitch = p.bm_pitch;
  float left = 0.f, right = 0.f;
  x.top = x.bottom = 0.f;
  if (!MASK) {
    const float *bcb = p.bc4 + (size_t)item.b * 4;
    x.top = __ldg(bcb + 0); x.bottom = __ldg(bcb + 1); left = __ldg(bcb + 2); right = __ldg(bcb + 3);
  }
#pragma unroll
  for (int q = 0; q < 4; ++q) {
    const int col = c.c0 + q;
    x.keep[q] = (c.interior >> q & 1u) ? 0xffffffffu : 0u;
    x.gval[q] = col == 0 ? __float_as_uint(left) : (col == N - 1 ? __float_as_uint(right) : 0u);
  }
  TailState<L> s;
#pragma unroll
  for (int h = 0; h < 2; ++h) {
    s.pr[h] = s.ps[h] = s.py[h] = s.pv[h] = zero2();
    s.pm[h] = bc2(1.0f);
#pragma unroll
    for (int k = 0; k < L; ++k) s.S[k][h][0] = s.S[k][h][1] = zero2();
  }
#pragma unroll
  for (int d = 0; d < DS; ++d) *reinterpret_cast<float4 *>(x.kdelay + d * 512) = make_float4(0.f, 0.f, 0.f, 0.f);

  const int it_first = max(x.i0 - L, 0), it_last = x.i1 - 1 + L;
  const size_t lane_in = (size_t)item.b * p.in_bstride + c.c0;
  // r row it_first (+ its mask row) is fetched here; every step fetches the next one
  if (it_first >= 1 && it_first <= N - 2) {
    load_row(p.rin + lane_in + (size_t)it_first * p.in_pitch, x.in_perm, x.ld_ok, c.in_img, s.pr[0], s.pr[1]);
    if (MASK)
      load_row(p.bm + (size_t)item.b * p.bm_bstride + c.c0 + (size_t)it_first * p.bm_pitch, x.bm_perm, x.ld_ok, c.in_img,
               s.pm[0], s.pm[1]);
  }
  x.prin = p.rin + lane_in + (size_t)(it_first + 1) * p.in_pitch;
  x.pmin = MASK ? p.bm + (size_t)item.b * p.bm_bstride + c.c0 + (size_t)(it_first + 1) * p.bm_pitch : nullptr;
  // skip / y / value rows are fetched one step ahead of their output row; the first fetch is for row nr0
  const int nr0 = max(it_first - L + 1, 0);
  x.pskip = p.skip + lane_in + (size_t)nr0 * p.in_pitch;
  x.py = p.y + lane_in + (size_t)nr0 * p.in_pitch;
  x.pvin = MASK ? p.bv + (size_t)item.b * p.bv_bstride + c.c0 + (size_t)nr0 * p.bv_pitch : nullptr;
  x.pout = p.out + (size_t)item.b * p.out_bstride + (size_t)max(it_first - L, 0) * p.out_pitch + c.c0;
  unsigned woff = 0;
  for (int it = it_first; it <= it_last; it += 2) {
    tail_step<L, ACT, MASK, 0>(s, x, p, it, woff);
    if (it + 1 <= it_last) tail_step<L, ACT, MASK, 1>(s, x, p, it + 1, woff);
  }
}

// --------------------------------------------------------- T Jacobi steps ---
struct JacobiStreamParams {
  const float *in;
  float *out;
  const float *f;    // NULL, or the source (only with T == 1)
  const float *bc4;  // (B,4), or NULL on a mask geometry
  long in_bstride, out_bstride, f_bstride;
  int in_pitch, out_pitch, f_pitch;
  int in_perm, out_perm, f_perm;  // field is in the pair-permuted padded layout (128-bit accesses)
  // mask geometry (only with T == 1): the two planes of bc = [values, mask] as field views
  const float *bv, *bm;
  long bv_bstride, bm_bstride;
  int bv_pitch, bm_pitch, bv_perm, bm_perm;
  int B, N;
  Split split;
};

// MASK: G = y*(1-m)+v at every cell, Psi zero-padded on the whole grid (heat_utils.py:51-53,96-106).
// With T > 1 stage s needs the bc rows of grid row it - s: every arriving mask / value row is turned into
// "keep bits" (m == 0 and the cell exists) plus values and parked in a lane-private shared-memory ring of
// JAC_RING rows (32 bytes per lane and row), from which the T stages read their row; rows and columns outside
// the grid come out as exact zeros, which is the zero padding the next stage expects.
constexpr int JAC_RING = 8;  // rows, power of two, > largest T on a mask geometry
constexpr size_t jacobi_mask_smem(int T) { return T > 1 ? (size_t)JAC_RING * 2 * 32 * 16 : 0; }
// HASF (T > 1 only; the T == 1 kernels test p.f at run time): a source term is present, its rows wait in a ring too.
template <int T, bool MASK = false, bool HASF = false>
__global__ void __launch_bounds__(npde::STREAM_THREADS)
k_jacobi_stream(const NPDE_GRID_CONSTANT JacobiStreamParams p) {
  static_assert(!HASF || T > 1, "T == 1 handles the source at run time");
  static_assert(!MASK || T < JAC_RING, "mask ring too short");
  constexpr bool MRING = MASK && T > 1;
  constexpr int RC = (T + 3) & ~3;  // column halo: whole lanes
  StripItem item;
  if (!decode_item(p.B, p.N, p.split, item)) return;
  const int lane = threadIdx.x & 31;
  const int lane_l = (lane + 31) & 31, lane_r = (lane + 1) & 31;
  const int N = p.N;
  const int vs = item.strip * p.split.wv, ve = min(vs + p.split.wv, N);
  const LaneCols c = make_cols(vs - RC, lane, N, vs, ve);
  const int i0 = item.i0, i1 = item.i1;
  const float *xin = p.in + (size_t)item.b * p.in_bstride + c.c0;
  float *xout = p.out + (size_t)item.b * p.out_bstride + c.c0;
  constexpr bool FRING = HASF && T > 1;  // source rows wait in a shared-memory ring like the bc rows
  const float *fin = ((T == 1 || HASF) && p.f) ? p.f + (size_t)item.b * p.f_bstride + c.c0 : nullptr;
  float top = 0.f, bottom = 0.f, left = 0.f, right = 0.f;
  if (!MASK) {
    const float *bcb = p.bc4 + (size_t)item.b * 4;
    top = __ldg(bcb + 0), bottom = __ldg(bcb + 1), left = __ldg(bcb + 2), right = __ldg(bcb + 3);
  }
  const float *vin = MASK ? p.bv + (size_t)item.b * p.bv_bstride + c.c0 : nullptr;
  const float *min_ = MASK ? p.bm + (size_t)item.b * p.bm_bstride + c.c0 : nullptr;
  // permuted layout: a lane whose first column exists moves its whole group of four; the pad
  // columns it then touches only feed ring / out-of-grid cells and are rewritten by G or ignored
  const bool in_perm = p.in_perm != 0, out_perm = p.out_perm != 0, f_perm = p.f_perm != 0;
  const bool ld_ok = c.c0 >= 0 && c.c0 < N;
  const bool st_ok = c.store != 0u;

  // win[0] = x, win[s] = iterate after s steps (s < T); rows of step t, t-1, t-2 (centre pairs
  // and outer neighbours)
  f2 win[T][3][2];
  float wl[T][3], wr[T][3];
#pragma unroll
  for (int s = 0; s < T; ++s)
#pragma unroll
    for (int r = 0; r < 3; ++r) {
      win[s][r][0] = win[s][r][1] = zero2();
      wl[s][r] = wr[s][r] = 0.0f;
    }

  // input rows it; stage s (1..T) produces row it - s in the same step (forward
  // order: the 5-point stencil needs no halo on its newest row)
  const int it_first = max(i0 - T, 0);
  const int it_last = i1 - 1 + T;
  const int x_last = min(it_last, N - 1);
  f2 pf[2], pff[2], pm[2], pv[2];
  pf[0] = pf[1] = pff[0] = pff[1] = pv[0] = pv[1] = zero2();
  pm[0] = pm[1] = bc2(1.0f);
  NPDE_DYN_SMEM(uint4, mring);  // MRING: slot (row & 7): [keep bits of 32 lanes][values of 32 lanes]
  if (MRING) {
#pragma unroll
    for (int r = 0; r < JAC_RING; ++r) {
      mring[r * 64 + lane] = make_uint4(0u, 0u, 0u, 0u);
      mring[r * 64 + 32 + lane] = make_uint4(0u, 0u, 0u, 0u);
    }
  }
  uint4 *fring = mring + (MRING ? JAC_RING * 64 : 0);  // FRING: slot (row & 7): the source row, 16 bytes per lane
  if (FRING) {
#pragma unroll
    for (int r = 0; r < JAC_RING; ++r) fring[r * 32 + lane] = make_uint4(0u, 0u, 0u, 0u);
  }
  // mask geometries: the ring need not be Dirichlet, so columns beyond the grid (pad columns of the permuted
  // layout, whatever they hold) must read as the zero padding of the stencil
  const unsigned img0 = (c.in_img & 1u) ? 0xffffffffu : 0u, img1 = (c.in_img & 2u) ? 0xffffffffu : 0u;
  const unsigned img2 = (c.in_img & 4u) ? 0xffffffffu : 0u, img3 = (c.in_img & 8u) ? 0xffffffffu : 0u;
  load_row(xin + (size_t)it_first * p.in_pitch, in_perm, ld_ok, c.in_img, pf[0], pf[1]);
  if (MASK) {
    pf[0] = and2(pf[0], img0, img2);
    pf[1] = and2(pf[1], img1, img3);
  }
  if (T == 1 && fin && it_first >= 1)
    load_row(fin + (size_t)(it_first - 1) * p.f_pitch, f_perm, ld_ok, c.in_img, pff[0], pff[1]);
  if (MASK && !MRING && it_first >= 1) {  // mask / value rows of the first output row it_first - 1
    load_row(min_ + (size_t)(it_first - 1) * p.bm_pitch, p.bm_perm != 0, ld_ok, c.in_img, pm[0], pm[1]);
    load_row(vin + (size_t)(it_first - 1) * p.bv_pitch, p.bv_perm != 0, ld_ok, c.in_img, pv[0], pv[1]);
  }

  for (int it_base = it_first; it_base <= it_last; it_base += 3) {
#pragma unroll
    for (int ph = 0; ph < 3; ++ph) {
      const int it = it_base + ph;
      const int s_new = ph, s_m1 = (ph + 2) % 3, s_m2 = (ph + 1) % 3;
      win[0][s_new][0] = pf[0];
      win[0][s_new][1] = pf[1];
#pragma unroll
      for (int s = 1; s <= T; ++s) {
        const int row = it - s;
        f2 o0, o1;
        psi_row(win[s - 1][s_m2][0], win[s - 1][s_m2][1], win[s - 1][s_m1][0], win[s - 1][s_m1][1],
                wl[s - 1][s_m1], wr[s - 1][s_m1], win[s - 1][s_new][0], win[s - 1][s_new][1], o0, o1);
        if (T == 1 && fin) {
          o0 = sub2(o0, pff[0]);
          o1 = sub2(o1, pff[1]);
        } else if (FRING && row >= 0 && row <= N - 1) {  // source row of grid row `row`, parked s steps ago
          const uint4 fr = fring[(row & (JAC_RING - 1)) * 32 + lane];
          o0 = sub2(o0, mk2(__uint_as_float(fr.x), __uint_as_float(fr.z)));
          o1 = sub2(o1, mk2(__uint_as_float(fr.y), __uint_as_float(fr.w)));
        }
        if (MRING) {  // keep bits / values of grid row `row`, parked s steps ago
          if (row >= 0 && row <= N - 1) {
            const uint4 k = mring[(row & (JAC_RING - 1)) * 64 + lane];
            const uint4 v = mring[(row & (JAC_RING - 1)) * 64 + 32 + lane];
            o0 = and2(o0, k.x, k.z);
            o1 = and2(o1, k.y, k.w);
            o0 = mk2(__fadd_rn(lo(o0), __uint_as_float(v.x)), __fadd_rn(hi(o0), __uint_as_float(v.z)));
            o1 = mk2(__fadd_rn(lo(o1), __uint_as_float(v.y)), __fadd_rn(hi(o1), __uint_as_float(v.w)));
          } else {
            o0 = o1 = zero2();
          }
        } else if (MASK) {  // m is 0 or 1: x*(1-m) is a select, then + v; columns beyond the grid stay exact zeros
          o0 = and2(o0, ((c.in_img & 1u) && lo(pm[0]) == 0.0f) ? 0xffffffffu : 0u,
                    ((c.in_img & 4u) && hi(pm[0]) == 0.0f) ? 0xffffffffu : 0u);
          o1 = and2(o1, ((c.in_img & 2u) && lo(pm[1]) == 0.0f) ? 0xffffffffu : 0u,
                    ((c.in_img & 8u) && hi(pm[1]) == 0.0f) ? 0xffffffffu : 0u);
          o0 = mk2(__fadd_rn(lo(o0), (c.in_img & 1u) ? lo(pv[0]) : 0.0f), __fadd_rn(hi(o0), (c.in_img & 4u) ? hi(pv[0]) : 0.0f));
          o1 = mk2(__fadd_rn(lo(o1), (c.in_img & 2u) ? lo(pv[1]) : 0.0f), __fadd_rn(hi(o1), (c.in_img & 8u) ? hi(pv[1]) : 0.0f));
        } else {
          apply_G_square(o0, o1, row, N, c, top, bottom, left, right);
        }
        if (s < T) {
          // rows above/below the grid must read as zero padding for the next stage's ring
          // cells only, and those are overwritten by G: no masking needed.
          win[s < T ? s : 0][s_new][0] = o0;
          win[s < T ? s : 0][s_new][1] = o1;
        } else if (row >= i0 && row < i1) {
          store_row(xout + (size_t)row * p.out_pitch, out_perm, st_ok, c.store, o0, o1);
        }
      }
#pragma unroll
      for (int s = 0; s < T; ++s)
        halo_scalars(win[s][s_new][0], win[s][s_new][1], lane_l, lane_r, wl[s][s_new], wr[s][s_new]);

      const int nx = it + 1;
      if (nx <= x_last) {
        load_row(xin + (size_t)nx * p.in_pitch, in_perm, ld_ok, c.in_img, pf[0], pf[1]);
        if (MASK) {
          pf[0] = and2(pf[0], img0, img2);
          pf[1] = and2(pf[1], img1, img3);
        }
      } else {
        pf[0] = pf[1] = zero2();
      }
      if ((T == 1 || HASF) && fin) {  // next step's stage-1 output row is `it`
        if (it <= N - 1) load_row(fin + (size_t)it * p.f_pitch, f_perm, ld_ok, c.in_img, pff[0], pff[1]);
        if (FRING && it <= N - 1)
          fring[(it & (JAC_RING - 1)) * 32 + lane] =
              make_uint4(__float_as_uint(lo(pff[0])), __float_as_uint(lo(pff[1])), __float_as_uint(hi(pff[0])),
                         __float_as_uint(hi(pff[1])));
        if (L2_AHEAD > 0 && it + L2_AHEAD <= x_last)
          prefetch_row(fin + (size_t)(it + L2_AHEAD) * p.f_pitch, lane, c.in_img);
      }
      if (MASK && it <= N - 1 && it < (MRING ? it_last : i1)) {
        load_row(min_ + (size_t)it * p.bm_pitch, p.bm_perm != 0, ld_ok, c.in_img, pm[0], pm[1]);
        load_row(vin + (size_t)it * p.bv_pitch, p.bv_perm != 0, ld_ok, c.in_img, pv[0], pv[1]);
        if (MRING) {  // row `it` is first used by stage 1 of the next step
          uint4 k, v;
          k.x = ((c.in_img & 1u) && lo(pm[0]) == 0.0f) ? 0xffffffffu : 0u;
          k.y = ((c.in_img & 2u) && lo(pm[1]) == 0.0f) ? 0xffffffffu : 0u;
          k.z = ((c.in_img & 4u) && hi(pm[0]) == 0.0f) ? 0xffffffffu : 0u;
          k.w = ((c.in_img & 8u) && hi(pm[1]) == 0.0f) ? 0xffffffffu : 0u;
          v.x = (c.in_img & 1u) ? __float_as_uint(lo(pv[0])) : 0u;
          v.y = (c.in_img & 2u) ? __float_as_uint(lo(pv[1])) : 0u;
          v.z = (c.in_img & 4u) ? __float_as_uint(hi(pv[0])) : 0u;
          v.w = (c.in_img & 8u) ? __float_as_uint(hi(pv[1])) : 0u;
          mring[(it & (JAC_RING - 1)) * 64 + lane] = k;
          mring[(it & (JAC_RING - 1)) * 64 + 32 + lane] = v;
        }
        if (L2_AHEAD > 0 && it + L2_AHEAD < i1) {
          prefetch_row(min_ + (size_t)(it + L2_AHEAD) * p.bm_pitch, lane, c.in_img);
          prefetch_row(vin + (size_t)(it + L2_AHEAD) * p.bv_pitch, lane, c.in_img);
        }
      }
      if (L2_AHEAD > 0 && nx + L2_AHEAD <= x_last)
        prefetch_row(xin + (size_t)(nx + L2_AHEAD) * p.in_pitch, lane, c.in_img);
    }
  }
}

// ------------------------------------------------- residual (fd_error) ---
// utils/heat_utils.py:108-132 on the square domain with aggregate 'max': per problem
// max_{interior} |1/4 u + 1/4 l - c + 1/4 r + 1/4 d - f| (the loss-kernel fmaf chain, oracle orc_fd_error).
// Same strip / band streaming as the Jacobi kernel, nothing is stored: every warp folds its maximum
// into err[b] with an integer atomicMax on the float bits (non-negative floats order like integers,
// a maximum is order independent, so the result is deterministic).  err must be zeroed beforehand.
struct ResidStreamParams {
  const float *in;
  const float *f;  // NULL or the source
  float *err;      // (B)
  long in_bstride, f_bstride;
  int in_pitch, f_pitch;
  int in_perm, f_perm;
  int B, N;
  Split split;
};

__global__ void __launch_bounds__(npde::STREAM_THREADS)
k_resid_stream(const NPDE_GRID_CONSTANT ResidStreamParams p) {
  constexpr int RC = 4;
  StripItem item;
  if (!decode_item(p.B, p.N, p.split, item)) return;
  const int lane = threadIdx.x & 31;
  const int lane_l = (lane + 31) & 31, lane_r = (lane + 1) & 31;
  const int N = p.N;
  const int vs = item.strip * p.split.wv, ve = min(vs + p.split.wv, N);
  const LaneCols c = make_cols(vs - RC, lane, N, vs, ve);
  const int o0 = max(item.i0, 1), o1 = min(item.i1, N - 1);  // rows whose residual this band owns
  if (o0 >= o1) return;
  const float *xin = p.in + (size_t)item.b * p.in_bstride + c.c0;
  const float *fin = p.f ? p.f + (size_t)item.b * p.f_bstride + c.c0 : nullptr;
  const bool in_perm = p.in_perm != 0, f_perm = p.f_perm != 0;
  const bool ld_ok = c.c0 >= 0 && c.c0 < N;
  const unsigned own = c.store & c.interior;  // columns this lane contributes
  const unsigned k0 = (own & 1u) ? 0xffffffffu : 0u, k1 = (own & 2u) ? 0xffffffffu : 0u,
                 k2 = (own & 4u) ? 0xffffffffu : 0u, k3 = (own & 8u) ? 0xffffffffu : 0u;

  f2 win[3][2];
  float wl[3], wr[3];
#pragma unroll
  for (int r = 0; r < 3; ++r) {
    win[r][0] = win[r][1] = zero2();
    wl[r] = wr[r] = 0.0f;
  }
  const int it_first = o0 - 1, it_last = o1;  // input rows o0-1 .. o1 (all exist: 0 <= o0-1, o1 <= N-1)
  f2 pf[2], pff[2];
  pff[0] = pff[1] = zero2();
  load_row(xin + (size_t)it_first * p.in_pitch, in_perm, ld_ok, c.in_img, pf[0], pf[1]);
  float vmax = 0.0f;
  const f2 q = bc2(0.25f), neg1 = bc2(-1.0f);
  for (int it_base = it_first; it_base <= it_last; it_base += 3) {
#pragma unroll
    for (int ph = 0; ph < 3; ++ph) {
      const int it = it_base + ph;
      if (it > it_last) break;
      const int s_new = ph, s_m1 = (ph + 2) % 3, s_m2 = (ph + 1) % 3;
      win[s_new][0] = pf[0];
      win[s_new][1] = pf[1];
      const int row = it - 1;  // centre row of this step
      if (row >= o0) {
        const f2 m0 = win[s_m1][0], m1 = win[s_m1][1];
        f2 a0 = mul2(q, win[s_m2][0]);
        a0 = fma_lohi(0.25f, wl[s_m1], lo(m1), a0);
        a0 = fma2(neg1, m0, a0);
        a0 = fma2(q, m1, a0);
        a0 = fma2(q, win[s_new][0], a0);
        f2 a1 = mul2(q, win[s_m2][1]);
        a1 = fma2(q, m0, a1);
        a1 = fma2(neg1, m1, a1);
        a1 = fma_lohi(0.25f, hi(m0), wr[s_m1], a1);
        a1 = fma2(q, win[s_new][1], a1);
        if (fin) {
          a0 = sub2(a0, pff[0]);
          a1 = sub2(a1, pff[1]);
        }
        a0 = and2(a0, k0 & 0x7fffffffu, k2 & 0x7fffffffu);  // |.| and column ownership in one mask
        a1 = and2(a1, k1 & 0x7fffffffu, k3 & 0x7fffffffu);
        vmax = fmaxf(vmax, fmaxf(fmaxf(lo(a0), hi(a0)), fmaxf(lo(a1), hi(a1))));
      }
      halo_scalars(win[s_new][0], win[s_new][1], lane_l, lane_r, wl[s_new], wr[s_new]);
      const int nx = it + 1;
      if (nx <= it_last) {
        load_row(xin + (size_t)nx * p.in_pitch, in_perm, ld_ok, c.in_img, pf[0], pf[1]);
        if (L2_AHEAD > 0 && nx + L2_AHEAD <= it_last)
          prefetch_row(xin + (size_t)(nx + L2_AHEAD) * p.in_pitch, lane, c.in_img);
      }
      if (fin && it >= o0 && it < o1) {  // f row of the next step's centre row
        load_row(fin + (size_t)it * p.f_pitch, f_perm, ld_ok, c.in_img, pff[0], pff[1]);
        if (L2_AHEAD > 0 && it + L2_AHEAD < o1) prefetch_row(fin + (size_t)(it + L2_AHEAD) * p.f_pitch, lane, c.in_img);
      }
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) vmax = fmaxf(vmax, __shfl_xor_sync(NPDE_FULL_MASK, vmax, o));
  if (lane == 0) atomicMax(reinterpret_cast<int *>(p.err) + item.b, __float_as_int(vmax));
}

// ------------------------------------------------------------- host side ---
// Bands exist only to fill the machine when B * strips is small; every band re-streams
// 2 * halo_rows rows and pays the pipeline fill, so bands are kept >= 8 * halo_rows rows long.
Split make_split(int B, int N, int halo_cols, int halo_rows, int resident_warps) {
  Split s;
  s.wv = STRIP - 2 * halo_cols;
  s.strips = (N + s.wv - 1) / s.wv;
  long per_band = (long)B * s.strips;
  int bands = (int)(resident_warps / (per_band > 0 ? per_band : 1));
#ifndef NPDE_BAND_FACTOR
#define NPDE_BAND_FACTOR 2  // bands of >= 2 * halo rows: small problems are latency bound, more warps win (12.3 -> 8.2 us per launch at 32 x 65^2)
#endif
  int max_bands = N / (NPDE_BAND_FACTOR * (halo_rows > 0 ? halo_rows : 1));
  if (bands > max_bands) bands = max_bands;
  if (bands < 1) bands = 1;
  s.band_rows = (N + bands - 1) / bands;
  s.bands = (N + s.band_rows - 1) / s.band_rows;
  return s;
}

// Resident warps of a kernel on the CURRENT device.  The occupancy answer is a property of (kernel, device): the
// cache is one relaxed atomic per device (a racing first call on two threads computes the same value twice).
constexpr int MAX_DEVICES = 64;
struct OccCache {
  std::atomic<int> v[MAX_DEVICES];
  OccCache() { for (auto &a : v) a.store(0, std::memory_order_relaxed); }
};
int resident_warps_cached(const void *kernel, size_t smem, OccCache *cache) {
#ifdef NPDE_EMU
  (void)kernel; (void)smem; (void)cache;
  return 8;  // small on purpose: exercises several bands per image in the emulator
#else
  int dev = 0, sms = 0, blocks = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= MAX_DEVICES) return 148 * 16;
  const int hit = cache->v[dev].load(std::memory_order_relaxed);
  if (hit > 0) return hit;
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&blocks, kernel, npde::STREAM_THREADS, smem) != cudaSuccess ||
      blocks < 1)
    blocks = 1;
  const int val = sms * blocks * (npde::STREAM_THREADS / 32);
  cache->v[dev].store(val, std::memory_order_relaxed);
  return val;
#endif
}

template <int L, int ACT>
int launch_conv_f(ConvStreamParams &p, dim3 grid, dim3 block, size_t smem, cudaStream_t st) {
  if (p.bm) {
    if (p.f) NPDE_LAUNCH((k_conv_stream<L, ACT, true, true>), grid, block, smem, st, p);
    else NPDE_LAUNCH((k_conv_stream<L, ACT, false, true>), grid, block, smem, st, p);
  } else {
    if (p.f) NPDE_LAUNCH((k_conv_stream<L, ACT, true, false>), grid, block, smem, st, p);
    else NPDE_LAUNCH((k_conv_stream<L, ACT, false, false>), grid, block, smem, st, p);
  }
  NPDE_CHECK_LAUNCH();
  return NPDE_OK;
}

template <int L>
int launch_conv(ConvStreamParams &p, cudaStream_t st) {
  constexpr int R = L + 1, RC = (R + 3) & ~3, DS = (L + 1 <= 4) ? 4 : 8;
  // y delay line (+ keep-mask delay line) + x row ring (+ f row ring)
  const size_t smem = (size_t)512 * ((p.bm ? 2 : 1) * DS + (p.f ? 2 : 1) * RING_D) * (npde::STREAM_THREADS / 32);
  static OccCache cache[4];
#ifdef NPDE_DEV_ONLY_CONV3
  int resident = resident_warps_cached((const void *)k_conv_stream<L, NPDE_ACT_CLAMP, false, false>, smem, &cache[0]);
#else
  int resident =
      p.bm ? (p.f ? resident_warps_cached((const void *)k_conv_stream<L, NPDE_ACT_CLAMP, true, true>, smem, &cache[3])
                  : resident_warps_cached((const void *)k_conv_stream<L, NPDE_ACT_CLAMP, false, true>, smem, &cache[2]))
           : (p.f ? resident_warps_cached((const void *)k_conv_stream<L, NPDE_ACT_CLAMP, true, false>, smem, &cache[1])
                  : resident_warps_cached((const void *)k_conv_stream<L, NPDE_ACT_CLAMP, false, false>, smem, &cache[0]));
#endif
  p.split = make_split(p.B, p.N, RC, R, resident);
  dim3 grid((unsigned)((long)p.B * p.split.strips * p.split.bands)), block(npde::STREAM_THREADS);
#ifdef NPDE_DEV_ONLY_CONV3  // tools/ab_bench.py: quick-to-compile tuning builds, clamp / no source only
  if (p.act != NPDE_ACT_CLAMP || p.f || p.bm) return NPDE_ERR_ARG;
  NPDE_LAUNCH((k_conv_stream<L, NPDE_ACT_CLAMP, false, false>), grid, block, smem, st, p);
  NPDE_CHECK_LAUNCH();
  return NPDE_OK;
#else
  switch (p.act) {
    case ACT_HEAD: return launch_conv_f<L, ACT_HEAD>(p, grid, block, smem, st);
    case NPDE_ACT_NONE: return launch_conv_f<L, NPDE_ACT_NONE>(p, grid, block, smem, st);
    case NPDE_ACT_CLAMP: return launch_conv_f<L, NPDE_ACT_CLAMP>(p, grid, block, smem, st);
    case NPDE_ACT_SIGMOID: return launch_conv_f<L, NPDE_ACT_SIGMOID>(p, grid, block, smem, st);
    default: return NPDE_ERR_ACT;
  }
#endif
}

template <int T, bool MASK, bool HASF>
int launch_jacobi_f(JacobiStreamParams &p, cudaStream_t st) {
  constexpr int RC = (T + 3) & ~3;
  static OccCache cache;
  // mask ring (keep bits + values) and source ring
  const size_t smem = (MASK ? jacobi_mask_smem(T) : 0) + (HASF ? (size_t)JAC_RING * 32 * 16 : 0);
  int resident = resident_warps_cached((const void *)k_jacobi_stream<T, MASK, HASF>, smem, &cache);
  p.split = make_split(p.B, p.N, RC, T, resident);
  dim3 grid((unsigned)((long)p.B * p.split.strips * p.split.bands)), block(npde::STREAM_THREADS);
  NPDE_LAUNCH((k_jacobi_stream<T, MASK, HASF>), grid, block, smem, st, p);
  NPDE_CHECK_LAUNCH();
  return NPDE_OK;
}
template <int T, bool MASK = false>
int launch_jacobi(JacobiStreamParams &p, cudaStream_t st) {
  if constexpr (T > 1 && T <= 4) {
    if (p.f) return launch_jacobi_f<T, MASK, true>(p, st);
  }
  return launch_jacobi_f<T, MASK, false>(p, st);
}

}  // namespace

namespace npde {

bool conv_stream_supported(int bc_kind, int n_layers, const float *w_host) {
  return (bc_kind == NPDE_BC_SQUARE || bc_kind == NPDE_BC_MASK) && n_layers >= 1 && n_layers <= 4 &&
         w_host != nullptr;
}

// the pair-permuted layout needs 16-byte aligned rows and room for a whole lane at the right edge
static bool perm_ok(const void *p, int pitch, long bstride, int perm, int N) {
  if (!p || !perm) return true;
  return ((uintptr_t)p & 15) == 0 && (pitch & 3) == 0 && (bstride & 3) == 0 && pitch >= ((N + 3) & ~3);
}

static int conv_stream_impl(const FieldView &in, const FieldViewMut &out, float *yout, const BcFields &bc,
                            const FieldView &f, const float *w_host, int n_layers, int act, int B, int N,
                            cudaStream_t st) {
  ConvStreamParams p;
  p.yout = yout;
  p.bc4 = bc.bc4;
  p.bv = bc.values.p; p.bv_pitch = bc.values.pitch; p.bv_bstride = bc.values.bstride; p.bv_perm = bc.values.perm;
  p.bm = bc.mask.p; p.bm_pitch = bc.mask.pitch; p.bm_bstride = bc.mask.bstride; p.bm_perm = bc.mask.perm;
  if ((bc.bc4 == nullptr) == (bc.mask.p == nullptr) || (bc.mask.p == nullptr) != (bc.values.p == nullptr))
    return NPDE_ERR_BC;
  if (!perm_ok(p.bv, p.bv_pitch, p.bv_bstride, p.bv_perm, N) || !perm_ok(p.bm, p.bm_pitch, p.bm_bstride, p.bm_perm, N))
    return NPDE_ERR_ARG;
  p.in = in.p; p.in_pitch = in.pitch; p.in_bstride = in.bstride;
  p.out = out.p; p.out_pitch = out.pitch; p.out_bstride = out.bstride;
  p.f = f.p; p.f_pitch = f.pitch; p.f_bstride = f.bstride;
  p.B = B; p.N = N; p.act = act;
  p.in_perm = in.perm; p.out_perm = out.perm; p.f_perm = f.p ? f.perm : 0;
  if (!perm_ok(in.p, in.pitch, in.bstride, in.perm, N) || !perm_ok(out.p, out.pitch, out.bstride, out.perm, N) ||
      !perm_ok(f.p, f.pitch, f.bstride, f.perm, N))
    return NPDE_ERR_ARG;
  for (int l = 0; l < 4; ++l)
    for (int t = 0; t < 9; ++t) p.w[l][t] = (l < n_layers) ? w_host[l * 9 + t] : 0.0f;
  switch (n_layers) {
#ifndef NPDE_DEV_ONLY_CONV3
    case 1: return launch_conv<1>(p, st);
    case 2: return launch_conv<2>(p, st);
    case 4: return launch_conv<4>(p, st);
#endif
    case 3: return launch_conv<3>(p, st);
    default: return NPDE_ERR_LAYERS;
  }
}

int conv_stream(const FieldView &in, const FieldViewMut &out, const BcFields &bc, const FieldView &f,
                const float *w_host, int n_layers, int act, int B, int N, cudaStream_t st) {
  if (act < NPDE_ACT_NONE || act > NPDE_ACT_SIGMOID) return NPDE_ERR_ACT;
  return conv_stream_impl(in, out, nullptr, bc, f, w_host, n_layers, act, B, N, st);
}

// Head of a U-Net step: y = Psi(x) - f and r = the n_layers pre-smoothing convs of the finest level applied
// to the masked residual, both written as frames (rout / yout share pitch, batch stride and layout).
int conv_head_stream(const FieldView &in, const BcFields &bc, const FieldView &f, const float *w_host, int n_layers,
                     const FieldViewMut &rout, float *yout, int B, int N, cudaStream_t st) {
  return conv_stream_impl(in, rout, yout, bc, f, w_host, n_layers, ACT_HEAD, B, N, st);
}

// Tail of a U-Net step: out = G(pad(act(y + (convs(rin) + skip)))); rin, skip, y share one frame layout.
template <int L, int ACT>
static int launch_tail(TailStreamParams &p, cudaStream_t st) {
  constexpr int RC = (L + 3) & ~3, DS = (L + 1 <= 4) ? 4 : 8;
  const size_t smem = (size_t)512 * DS;
  static OccCache cache[2];
  int resident = p.bm ? resident_warps_cached((const void *)k_tail_stream<L, ACT, true>, smem, &cache[1])
                      : resident_warps_cached((const void *)k_tail_stream<L, ACT, false>, smem, &cache[0]);
  p.split = make_split(p.B, p.N, RC, L, resident);
  dim3 grid((unsigned)((long)p.B * p.split.strips * p.split.bands)), block(npde::STREAM_THREADS);
  if (p.bm) NPDE_LAUNCH((k_tail_stream<L, ACT, true>), grid, block, smem, st, p);
  else NPDE_LAUNCH((k_tail_stream<L, ACT, false>), grid, block, smem, st, p);
  NPDE_CHECK_LAUNCH();
  return NPDE_OK;
}
template <int L>
static int launch_tail_act(TailStreamParams &p, int act, cudaStream_t st) {
  switch (act) {
    case NPDE_ACT_NONE: return launch_tail<L, NPDE_ACT_NONE>(p, st);
    case NPDE_ACT_CLAMP: return launch_tail<L, NPDE_ACT_CLAMP>(p, st);
    case NPDE_ACT_SIGMOID: return launch_tail<L, NPDE_ACT_SIGMOID>(p, st);
    default: return NPDE_ERR_ACT;
  }
}

int conv_tail_stream(const FieldView &rin, const float *skip, const float *y, const BcFields &bc, const float *w_host,
                     int n_layers, int act, const FieldViewMut &out, int B, int N, cudaStream_t st) {
#ifdef NPDE_DEV_ONLY_CONV3
  (void)rin; (void)skip; (void)y; (void)bc; (void)w_host; (void)n_layers; (void)act; (void)out; (void)B; (void)N; (void)st;
  return NPDE_ERR_ARG;
#else
  TailStreamParams p;
  p.rin = rin.p; p.skip = skip; p.y = y;
  p.in_pitch = rin.pitch; p.in_bstride = rin.bstride; p.in_perm = rin.perm;
  p.out = out.p; p.out_pitch = out.pitch; p.out_bstride = out.bstride; p.out_perm = out.perm;
  p.bc4 = bc.bc4;
  p.bv = bc.values.p; p.bv_pitch = bc.values.pitch; p.bv_bstride = bc.values.bstride; p.bv_perm = bc.values.perm;
  p.bm = bc.mask.p; p.bm_pitch = bc.mask.pitch; p.bm_bstride = bc.mask.bstride; p.bm_perm = bc.mask.perm;
  if ((bc.bc4 == nullptr) == (bc.mask.p == nullptr)) return NPDE_ERR_BC;
  if (!perm_ok(rin.p, rin.pitch, rin.bstride, rin.perm, N) || !perm_ok(out.p, out.pitch, out.bstride, out.perm, N) ||
      !perm_ok(p.bv, p.bv_pitch, p.bv_bstride, p.bv_perm, N) || !perm_ok(p.bm, p.bm_pitch, p.bm_bstride, p.bm_perm, N))
    return NPDE_ERR_ARG;
  p.B = B; p.N = N;
  for (int l = 0; l < 4; ++l)
    for (int t = 0; t < 9; ++t) p.w[l][t] = (l < n_layers) ? w_host[l * 9 + t] : 0.0f;
  switch (n_layers) {
    case 1: return launch_tail_act<1>(p, act, st);
    case 2: return launch_tail_act<2>(p, act, st);
    case 3: return launch_tail_act<3>(p, act, st);
    case 4: return launch_tail_act<4>(p, act, st);
    default: return NPDE_ERR_LAYERS;
  }
#endif
}

int jacobi_stream_max_depth(const float *f) { return f ? 4 : 8; }

int fd_error_max_stream(const FieldView &x, const FieldView &f, float *err, int B, int N, cudaStream_t st) {
  ResidStreamParams p;
  p.in = x.p; p.in_pitch = x.pitch; p.in_bstride = x.bstride; p.in_perm = x.perm;
  p.f = f.p; p.f_pitch = f.pitch; p.f_bstride = f.bstride; p.f_perm = f.p ? f.perm : 0;
  p.err = err; p.B = B; p.N = N;
  if (!perm_ok(x.p, x.pitch, x.bstride, x.perm, N) || !perm_ok(f.p, f.pitch, f.bstride, f.perm, N)) return NPDE_ERR_ARG;
  static OccCache cache;
  int resident = resident_warps_cached((const void *)k_resid_stream, 0, &cache);
  p.split = make_split(B, N, 4, 1, resident);
  cudaError_t e = cudaMemsetAsync(err, 0, sizeof(float) * (size_t)B, st);
  if (e != cudaSuccess) return (int)e;
  dim3 grid((unsigned)((long)B * p.split.strips * p.split.bands)), block(npde::STREAM_THREADS);
  NPDE_LAUNCH(k_resid_stream, grid, block, 0, st, p);
  NPDE_CHECK_LAUNCH();
  return NPDE_OK;
}

int jacobi_stream(const FieldView &in, const FieldViewMut &out, const BcFields &bc, const FieldView &f, int depth,
                  int B, int N, cudaStream_t st) {
  JacobiStreamParams p;
  p.bc4 = bc.bc4;
  p.bv = bc.values.p; p.bv_pitch = bc.values.pitch; p.bv_bstride = bc.values.bstride; p.bv_perm = bc.values.perm;
  p.bm = bc.mask.p; p.bm_pitch = bc.mask.pitch; p.bm_bstride = bc.mask.bstride; p.bm_perm = bc.mask.perm;
  if ((bc.bc4 == nullptr) == (bc.mask.p == nullptr) || (bc.mask.p == nullptr) != (bc.values.p == nullptr))
    return NPDE_ERR_BC;
  if (!perm_ok(p.bv, p.bv_pitch, p.bv_bstride, p.bv_perm, N) || !perm_ok(p.bm, p.bm_pitch, p.bm_bstride, p.bm_perm, N))
    return NPDE_ERR_ARG;
  if (p.bm && depth > 4) return NPDE_ERR_ARG;
  p.in = in.p; p.in_pitch = in.pitch; p.in_bstride = in.bstride;
  p.out = out.p; p.out_pitch = out.pitch; p.out_bstride = out.bstride;
  p.f = f.p; p.f_pitch = f.pitch; p.f_bstride = f.bstride;
  p.B = B; p.N = N;
  p.in_perm = in.perm; p.out_perm = out.perm; p.f_perm = f.p ? f.perm : 0;
  if (!perm_ok(in.p, in.pitch, in.bstride, in.perm, N) || !perm_ok(out.p, out.pitch, out.bstride, out.perm, N) ||
      !perm_ok(f.p, f.pitch, f.bstride, f.perm, N))
    return NPDE_ERR_ARG;
  if (f.p && depth > 4) return NPDE_ERR_ARG;  // the source ring holds 8 rows
  switch (depth) {
    case 1: return p.bm ? launch_jacobi<1, true>(p, st) : launch_jacobi<1>(p, st);
#ifndef NPDE_DEV_ONLY_CONV3
    case 2: return p.bm ? launch_jacobi<2, true>(p, st) : launch_jacobi<2>(p, st);
    case 4: return p.bm ? launch_jacobi<4, true>(p, st) : launch_jacobi<4>(p, st);
    case 8: return launch_jacobi<8>(p, st);
#endif
    default: return NPDE_ERR_ARG;
  }
}

}  // namespace npde
