// npde_wide.cu -- the k-step loop of Phi_H = ConvIterator.forward (models/iterators/conv_iterator.py:33-51)
// on the square domain, second generation of the fused register-streaming kernel ("wide strips").
//
// What changed against k_conv_stream (npde_stream.cu), and why:
//   * Batch-interleaved field layout "W8".  Row i of ALL images is one contiguous virtual row
//         [ image 0 row i | image 1 row i | ... | image B-1 row i ],
//     every image padded to G8 = ceil(N/8) groups of eight columns, and inside a group the columns are
//     stored as (c0,c4,c1,c5,c2,c6,c3,c7).  A warp strip is 32 consecutive groups of the VIRTUAL row, so
//     strips tile B*G8*8 columns instead of each image's 2^n+1 separately: at 64 x 1025^2 the computed
//     columns drop from 1152 to 1100 per image row (halo waste 12.4% -> 7.3%), and at 257^2 from 384 to 284.
//     Ring columns and the padding between two images are ordinary "not interior" columns (the EDGE masks).
//   * Eight adjacent columns per lane as four packed pairs Qk = {c_k, c_k+4}.  With this pairing the left
//     and right operands of a 3-wide stencil are whole register pairs for every output pair except the two
//     at the lane's ends, and those two are built with one register move each ({c-1,c3}, {c4,c8}), so EVERY
//     tap of every stage is one FFMA2; one SHFL pair serves eight columns instead of four; a lane moves its
//     row with one 256-bit access (LDG/STG.E.ENL2.256, new on sm_100).  About 170 instructions per
//     8-column row step instead of 2 x 129, for the same 264 FMA-pipe cycles.
//     (At L = 3 the two straddling taps run as scalar FFMA on the halves instead: NPDE_W_OUTER_SCALAR.)
//   * Register pairs are carried as two 32-bit values (NPDE_F2_HALVES, npde_f2.cuh) and packed only inside the
//     packed instruction: ptxas then keeps the loop-carried partial sums in place (no ~80 MOVs per two row steps).
//   * Work partition: bands x strips one-warp CTAs in one wave, the strips of a band in lock step; band counts of
//     interior and edge strips from a cost model, occupancy enforced through the shared-memory request, launches
//     chained with programmatic dependent launch (launch_wide_f); the band warm-up runs store-free.
// The arithmetic is unchanged: per output the oracle's order (oracle/npde_oracle.c conv3x3_img, orc_fd_step),
// explicit fma.rn / mul.rn / sub.rn only, so iterates stay bit-identical to k_conv_stream and the CPU oracle.
#include <stdlib.h>

#include <algorithm>
#include <atomic>

#include "npde_common.cuh"
#include "npde_internal.h"

// Register pairs are carried as two 32-bit values in this translation unit (npde_f2.cuh): with 64-bit values ptxas
// copies every loop-carried partial-sum row home at the end of the 2x unrolled loop (~80 MOVs per two row steps).
#if !defined(NPDE_F2_PACKED64) && !defined(NPDE_SCALAR_F2) && !defined(NPDE_F2_HALVES)
#define NPDE_F2_HALVES 1
#endif
#include "npde_f2.cuh"

namespace {

constexpr int WCOLS = 8;      // columns per lane
constexpr int WVALID = 30;    // lanes 1..30 of a strip store; lanes 0 and 31 are the halo (radius <= 8 columns)
constexpr int WSLOT = 1024;   // bytes of one row segment of a strip (32 lanes x 8 columns x 4 bytes)
constexpr int WMAX_EDGE = 1024;  // edge strips that can get their own band count (else: uniform bands)
#ifndef NPDE_W_TMA
#define NPDE_W_TMA 0  // 1: row segments arrive by TMA bulk copies (one elected lane, mbarrier completion), see wring_fill
#endif
// Row segments in flight per warp (cp.async ring in shared memory), by layer count.  L >= 3 (fp32-pipe bound, three
// warps per sub-partition): 8 -- 102.0 vs 106.3 us per launch at 64 x 1025^2, 0.758 vs 0.737 of the roofline sustained
// on the same box; 16 halves the occupancy (141 us).  L <= 2 (memory bound): 4 -- a deeper ring scatters the DRAM
// accesses in flight over more rows (L = 1: 111.3 vs 113.6 us, L = 2: 113.1 vs 115.9 us).  NPDE_WRING=n (tuning
// builds) forces one depth for every L.
template <int L>
struct WRingOf {
#ifdef NPDE_WRING
  static constexpr int value = NPDE_WRING;
#else
  static constexpr int value = (L >= 3 && !NPDE_W_TMA) ? 8 : 4;
#endif
  static_assert((value & (value - 1)) == 0 && value >= 2, "ring size: power of two");
};
constexpr int WRING_TMA = 4;  // (NPDE_W_TMA builds: wring_fill is not templated on L)
#ifndef NPDE_WIDE_MIN_CTAS
#define NPDE_WIDE_MIN_CTAS 8
#endif
#ifndef NPDE_W_SKEW
#define NPDE_W_SKEW 0
#endif
// NPDE_W_SKEW = 1: the stages of a step are SKEWED by one step -- conv stage k consumes the row that stage k-1
// finished in the PREVIOUS step (kept in registers with its two halo values) instead of the row it finishes in the
// same step.  Nothing computed in a step depends on anything else computed in that step, so the serial chain
// Psi -> SHFL -> stage 1 -> SHFL -> ... -> stage L -> store (four shuffle latencies and four FFMA2 chains per row
// step, with two resident warps per scheduler to hide them) disappears; the price is L more warm-up steps per band
// and a y delay line of 2L instead of L rows.  Same arithmetic per cell, so results stay bit-identical.
#if NPDE_W_SKEW && NPDE_W_TMA
#error "NPDE_W_SKEW and NPDE_W_TMA are separate experiments"
#endif

struct WideParams {
  const float *in;
  float *out;
  const float *f;    // NULL: Laplace; else the source in the W8 layout
  const float *bc4;  // (B,4)
  int B, N, G8, VG;  // G8 groups of 8 columns per image, VG = B*G8 groups per virtual row
  int pitch;         // floats per virtual row = 8*VG
  int strips;        // ceil(VG / WVALID)
  int bands, band_rows;  // a CTA (one warp) takes rows [band*band_rows, (band+1)*band_rows) of one strip
  // Edge strips (ring / padding columns inside: the masked loop costs ~1.24x the interior loop per row) are cut into
  // bands_edge >= bands shorter bands so that every warp of the single wave finishes at the same time.  CTAs
  // [0, strips*bands) are (strip, band < bands); the following n_edge*(bands_edge - bands) CTAs are the extra bands
  // of the edge strips, listed in edge_strip[].
  int bands_edge, band_rows_edge, n_edge;
  int act;
  float w[4][9];
  unsigned short edge_strip[WMAX_EDGE];
};

// does the strip need the EDGE code?  (not if its 32 groups are interior groups of ONE image)
__host__ __device__ __forceinline__ bool wide_strip_is_edge(int strip, int N, int G8, int VG) {
  const int v0 = strip * WVALID - 1;
  if (v0 < 0 || v0 + 31 >= VG) return true;
  const int b0 = v0 / G8, g0 = v0 - b0 * G8;
  const int g_hi = (N - 9) >= 0 ? (N - 9) / 8 : -1;  // last group whose eight columns are all interior
  return !(g0 >= 1 && g0 + 31 <= g_hi);
}

// ---- 256-bit global and 128-bit shared accesses of a lane's row: pairs q[k] = {c_k, c_k+4} ----------------
#ifdef NPDE_EMU
__device__ __forceinline__ void ld256(const float *p, f2 (&q)[4]) {
#pragma unroll
  for (int k = 0; k < 4; ++k) q[k] = mk2(p[2 * k], p[2 * k + 1]);
}
__device__ __forceinline__ void st256_cols(float *p, const float (&o)[8]) {
#pragma unroll
  for (int k = 0; k < 4; ++k) { p[2 * k] = o[k]; p[2 * k + 1] = o[k + 4]; }
}
__device__ __forceinline__ void lds128(const char *s, f2 &a, f2 &b) {
  const float *p = reinterpret_cast<const float *>(s);
  a = mk2(p[0], p[1]);
  b = mk2(p[2], p[3]);
}
__device__ __forceinline__ void sts128(char *s, f2 a, f2 b) {
  float *p = reinterpret_cast<float *>(s);
  p[0] = a.x; p[1] = a.y; p[2] = b.x; p[3] = b.y;
}
__device__ __forceinline__ void cp_async16_zfill(void *smem, const float *g, bool valid) {
  if (valid) memcpy(smem, g, 16);
  else memset(smem, 0, 16);
}
#else
#if defined(NPDE_SCALAR_F2) || defined(NPDE_F2_HALVES)
__device__ __forceinline__ void ld256(const float *p, f2 (&q)[4]) {
  float v[8];
  asm volatile("ld.global.nc.v8.f32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
               : "=f"(v[0]), "=f"(v[1]), "=f"(v[2]), "=f"(v[3]), "=f"(v[4]), "=f"(v[5]), "=f"(v[6]), "=f"(v[7]) : "l"(p));
#pragma unroll
  for (int k = 0; k < 4; ++k) q[k] = mk2(v[2 * k], v[2 * k + 1]);
}
#else
__device__ __forceinline__ void ld256(const float *p, f2 (&q)[4]) {
  asm volatile("ld.global.nc.v4.b64 {%0,%1,%2,%3}, [%4];" : "=l"(q[0]), "=l"(q[1]), "=l"(q[2]), "=l"(q[3]) : "l"(p));
}
#endif
// columns c0..c7 -> memory order (c0,c4,c1,c5,c2,c6,c3,c7), one STG.E.ENL2.256
__device__ __forceinline__ void st256_cols(float *p, const float (&o)[8]) {
  asm volatile("st.global.v8.f32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"l"(p), "f"(o[0]), "f"(o[4]), "f"(o[1]),
               "f"(o[5]), "f"(o[2]), "f"(o[6]), "f"(o[3]), "f"(o[7])
               : "memory");
}
#if defined(NPDE_SCALAR_F2) || defined(NPDE_F2_HALVES)
__device__ __forceinline__ void lds128(const char *s, f2 &a, f2 &b) {
  const float4 v = *reinterpret_cast<const float4 *>(s);
  a = mk2(v.x, v.y);
  b = mk2(v.z, v.w);
}
__device__ __forceinline__ void sts128(char *s, f2 a, f2 b) {
  *reinterpret_cast<float4 *>(s) = make_float4(a.x, a.y, b.x, b.y);
}
#else
__device__ __forceinline__ void lds128(const char *s, f2 &a, f2 &b) {
  const ulonglong2 v = *reinterpret_cast<const ulonglong2 *>(s);
  a = v.x;
  b = v.y;
}
__device__ __forceinline__ void sts128(char *s, f2 a, f2 b) {
  *reinterpret_cast<ulonglong2 *>(s) = make_ulonglong2(a, b);
}
#endif
__device__ __forceinline__ void cp_async16_zfill(void *smem, const float *g, bool valid) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"((unsigned)__cvta_generic_to_shared(smem)), "l"(g),
               "r"(valid ? 16 : 0)
               : "memory");
}
__device__ __forceinline__ void pin_reg(unsigned &v) { asm volatile("" : "+r"(v)); }
#endif
#ifdef NPDE_EMU
__device__ __forceinline__ void pin_reg(unsigned &) {}
#endif

// ---- TMA (bulk asynchronous copy engine) primitives: mbarrier + cp.async.bulk global -> shared -------------------
#ifdef NPDE_EMU
__device__ __forceinline__ void mbar_init(void *, unsigned) {}
__device__ __forceinline__ void mbar_fence_init() {}
__device__ __forceinline__ void mbar_wait(void *, unsigned) {}
__device__ __forceinline__ void tma_load_1d(void *dst, const void *src, unsigned bytes, void *) { memcpy(dst, src, bytes); }
#else
__device__ __forceinline__ void mbar_init(void *bar, unsigned count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"((unsigned)__cvta_generic_to_shared(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_fence_init() {
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // generic-proxy zero fill of the slots before the async writes
}
__device__ __forceinline__ void mbar_wait(void *bar, unsigned parity) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "WAIT_%=:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@!p bra WAIT_%=;\n"
      "}" ::"r"((unsigned)__cvta_generic_to_shared(bar)), "r"(parity) : "memory");
}
// one thread: announce `bytes` on the barrier and start the bulk copy that will complete them
__device__ __forceinline__ void tma_load_1d(void *dst, const void *src, unsigned bytes, void *bar) {
  const unsigned b = (unsigned)__cvta_generic_to_shared(bar);
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(b), "r"(bytes) : "memory");
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                   (unsigned)__cvta_generic_to_shared(dst)),
               "l"(src), "r"(bytes), "r"(b)
               : "memory");
}
#endif

#ifdef NPDE_EMU
__device__ __forceinline__ void pdl_wait_prior_grids() {}
__device__ __forceinline__ void pdl_launch_dependents() {}
#else
__device__ __forceinline__ void pdl_wait_prior_grids() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
#endif

__device__ __forceinline__ float wsel_bits(float v, unsigned keep, unsigned val) {
  return __uint_as_float((__float_as_uint(v) & keep) | val);
}
__device__ __noinline__ float wsigmoid_slow(float z) { return 1.0f / (1.0f + expf(-z)); }
template <int ACT>
__device__ __forceinline__ float wadd_act(float y, float h) {
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
  if (ACT == NPDE_ACT_SIGMOID) return wsigmoid_slow(z);
  return z;
}

template <int L>
struct WState {
  f2 X[2][4];      // X[P]: x row `it` (arrived during the previous step), X[P^1]: row it-1 (centre row of Psi)
  float xhl, xhr;  // outer neighbours of x row it-1: column c-1 (left lane's c7) and c8 (right lane's c0)
  f2 yu[4];        // 0.25 * x[it-2]: first tap of Psi for output row it-1
  f2 S[L][2][4];   // S[k-1][P]: partial sums of stage k that have seen taps w0..w5 (this step finishes them),
                   // S[k-1][P^1]: taps w0..w2
  f2 pff[4];       // f row it-1, fetched one step earlier
#if NPDE_W_SKEW
  f2 R[L][4];          // R[k]: the row stage k finished in the previous step (k = 0: the residual r_0), masked
  float Rl[L], Rr[L];  // its outer neighbours (columns c-1 and c8)
#endif
};

// rows of the y delay line (power of two): y waits L steps, 2L with skewed stages (read before write when equal)
template <int L>
struct WDelay {
  static constexpr int slots = NPDE_W_SKEW ? ((2 * L <= 4) ? 4 : 8) : ((L + 1 <= 4) ? 4 : 8);
};

struct WCtx {
  unsigned keep[8];  // per column c0..c7: all ones if interior column, else 0 (EDGE strips)
  unsigned gval[8];  // Dirichlet value bits of a ring column, else 0
  bool ld_ok, st_ok; // the lane's group exists / belongs to the strip's valid range
  bool fill_ok[2];   // the two 16-byte chunks this lane copies into a ring slot exist
  const float *pin, *pfin;    // this lane's group in the next x / f row fetched with a plain load (generic steps)
  const float *fill, *ffill;  // chunk `lane` of the next row segment to start into the ring (main loop)
  const float *safe; // any valid, 16-byte aligned address of the input field
  float *pout;       // this lane's group in the output row of the current step
  char *delay;       // slot 0 of the y delay line (shared memory)
  char *ring;        // slot 0 of the x row ring (the f ring follows it)
  unsigned a_off, b_off;      // this lane's two 16-byte chunks inside a slot ({Q0,Q1} and {Q2,Q3}), swizzled
  unsigned fill_off;          // where chunk `lane` of a segment goes inside a slot (chunk lane+32: + 512)
  unsigned rslot;    // ring slot (byte offset) that holds x row it+1
  // NPDE_W_TMA: the part of this strip's 1 KB segment that lies inside the virtual row (edge strips: the rest of a
  // slot stays zero), the number of rows started / consumed so far (slot = n % WRING, barrier phase = n / WRING)
  int tma_src, tma_dst, tma_bytes;
  float wreg[4][9];  // NPDE_W_TMA: the weights in ordinary registers (the uniform registers serve the TMA operands;
                     // with the weights there too ptxas re-loads them from the constant bank 19 times per two steps)
  int pending;       // rows started and not yet consumed
  unsigned phase;    // barrier phase of the slot `rslot` (flips when rslot wraps)
  char *bars;        // WRING mbarriers behind the rings
  float top, bottom;
  int N, i0, i1, x_last, f_last, pitch, lane_l, lane_r;
};

// Taps of a 3-wide stencil row on the four output pairs a[0..3] of a lane.  Operand row r[0..3] = pairs
// {c_k, c_k+4}; hl / hr = the outer neighbours c-1 / c8.  Every tap whose left / right operand is a whole pair of
// r is one packed FFMA2.  The two that are not -- the left tap of a[0] = {c0,c4} needs {c-1,c3}, the right tap of
// a[3] = {c3,c7} needs {c4,c8} -- either build that pair (NPDE_W_OUTER_SCALAR = 0: register moves) or run as two
// scalar FFMA on the halves (= 1; the halves representation of f2 makes that free of any packing).
#ifndef NPDE_W_OUTER_SCALAR
#define NPDE_W_OUTER_SCALAR 3  // the layer count L that runs the two outer taps as scalar FFMA (measured: L = 3
                               // 116.3 vs 117.4 us sustained; L = 4 is 1% slower with it); 0: never
#endif
template <bool OS>
__device__ __forceinline__ f2 wtap_l(float w, float hl, const f2 (&r)[4], f2 acc) {
  if (OS) return fma_lohi(w, hl, lo(r[3]), acc);
  return fma2(bc2(w), mk2(hl, lo(r[3])), acc);
}
template <bool OS>
__device__ __forceinline__ f2 wtap_r(float w, float hr, const f2 (&r)[4], f2 acc) {
  if (OS) return fma_lohi(w, hi(r[0]), hr, acc);
  return fma2(bc2(w), mk2(hi(r[0]), hr), acc);
}
template <bool OS>
__device__ __forceinline__ f2 wtap_l_first(float w, float hl, const f2 (&r)[4]) {
  if (OS) return mul_lohi(w, hl, lo(r[3]));
  return mul2(bc2(w), mk2(hl, lo(r[3])));
}
// three taps (wl, wc, wr) of one stencil row applied to the four output pairs a[0..3]
template <bool OS>
__device__ __forceinline__ void wtaps(const f2 (&r)[4], float hl, float hr, float wl, float wc, float wr, f2 (&a)[4]) {
  const f2 l2 = bc2(wl), c2 = bc2(wc), r2 = bc2(wr);
  a[0] = wtap_l<OS>(wl, hl, r, a[0]);  a[0] = fma2(c2, r[0], a[0]); a[0] = fma2(r2, r[1], a[0]);
  a[1] = fma2(l2, r[0], a[1]); a[1] = fma2(c2, r[1], a[1]); a[1] = fma2(r2, r[2], a[1]);
  a[2] = fma2(l2, r[1], a[2]); a[2] = fma2(c2, r[2], a[2]); a[2] = fma2(r2, r[3], a[2]);
  a[3] = fma2(l2, r[2], a[3]); a[3] = fma2(c2, r[3], a[3]); a[3] = wtap_r<OS>(wr, hr, r, a[3]);
}
// same for the first stencil row of an output: the first tap is a plain product (oracle conv3x3_img)
template <bool OS>
__device__ __forceinline__ void wtaps_first(const f2 (&r)[4], float hl, float hr, float wl, float wc, float wr,
                                            f2 (&a)[4]) {
  const f2 l2 = bc2(wl), c2 = bc2(wc), r2 = bc2(wr);
  a[0] = wtap_l_first<OS>(wl, hl, r);  a[0] = fma2(c2, r[0], a[0]); a[0] = fma2(r2, r[1], a[0]);
  a[1] = mul2(l2, r[0]); a[1] = fma2(c2, r[1], a[1]); a[1] = fma2(r2, r[2], a[1]);
  a[2] = mul2(l2, r[1]); a[2] = fma2(c2, r[2], a[2]); a[2] = fma2(r2, r[3], a[2]);
  a[3] = mul2(l2, r[2]); a[3] = fma2(c2, r[3], a[3]); a[3] = wtap_r<OS>(wr, hr, r, a[3]);
}
__device__ __forceinline__ void whalo(const f2 (&r)[4], int lane_l, int lane_r, float &hl, float &hr) {
  hl = __shfl_sync(NPDE_FULL_MASK, hi(r[3]), lane_l);  // c7 of the left lane
  hr = __shfl_sync(NPDE_FULL_MASK, lo(r[0]), lane_r);  // c0 of the right lane
}

// start the copy of one row segment (this strip's 32 groups = 1 KB) into a ring slot: lane l moves chunks l and l+32
template <bool EDGE>
__device__ __forceinline__ void wring_fill(char *slot, const float *chunk, const WCtx &x) {
#if NPDE_W_TMA
  // `chunk` is the (warp-uniform) start of the segment.  One bulk copy of its in-row part by an elected lane;
  // completion is signalled on the slot's mbarrier (x and f rows share it).
  if ((threadIdx.x & 31) == 0 && x.tma_bytes > 0)
    tma_load_1d(slot + x.tma_dst, chunk + x.tma_src, (unsigned)x.tma_bytes,
                x.bars + 8 * ((((unsigned)(slot - x.ring)) >> 10) & (unsigned)(WRING_TMA - 1)));
  return;
#endif
  __syncwarp();  // lanes write each other's bytes: everybody is done reading the slot's previous row
  if (EDGE) {
    // a chunk outside the virtual row is zero-filled (source size 0; the address passed is a valid one anyway)
    cp_async16_zfill(slot + x.fill_off, x.fill_ok[0] ? chunk : x.safe, x.fill_ok[0]);
    cp_async16_zfill(slot + x.fill_off + 512, x.fill_ok[1] ? chunk + 4 * 32 : x.safe, x.fill_ok[1]);
  } else {
    cp_async16(slot + x.fill_off, chunk);
    cp_async16(slot + x.fill_off + 512, chunk + 4 * 32);
  }
}

// One step: input row `it` is in s.X[P].  Stage 0 (Psi, residual) works on row it-1; conv stage k (1..L) finishes row
// it-1-k from the row stage k-1 produced in this step (NPDE_W_SKEW = 0), or row it-1-2k from the row stage k-1
// produced in the previous step (NPDE_W_SKEW = 1); the epilogue stores the row stage L finished.
// MAIN: every stage row is an interior row, the output row belongs to this piece and the rows to fetch exist
//       (no row tests); x (and f) rows arrive through the cp.async ring.
// EDGE: the strip contains ring / padding / out-of-range columns: stage outputs are masked to the interior
//       columns and G inserts the Dirichlet values.
// P:    step parity (the two x slots and the two partial-sum rows of a stage swap roles every step; a variant
//       that rotates three slots per stage to avoid consume-then-overwrite register moves needs a 3x unrolled
//       loop of 10 KB and measured 4% slower: instruction fetch).

// ---- stage 0: y = Psi(x) - f and r_0 = y - x for row it-1 -> (rc, rl, rr); y into the delay line; fetch x row it+1
template <int L, int ACT, bool HASF, bool MAIN, bool EDGE, int P>
__device__ __forceinline__ void wide_stage0(WState<L> &s, WCtx &x, int it, unsigned woff, f2 (&rc)[4], float &rl,
                                            float &rr) {
  constexpr int WRING = WRingOf<L>::value;
  constexpr unsigned RMASK = WRING * (unsigned)WSLOT - 1u;
  const int N = x.N;
  constexpr bool OS = (L == NPDE_W_OUTER_SCALAR);
  const f2 (&D)[4] = s.X[P];
  f2 (&C)[4] = s.X[P ^ 1];
  const int r0 = it - 1;
  const f2 q = bc2(0.25f);
  f2 y[4];  // taps in the oracle's order: up (in yu), left, right, down
  y[0] = wtap_l<OS>(0.25f, s.xhl, C, s.yu[0]); y[0] = fma2(q, C[1], y[0]); y[0] = fma2(q, D[0], y[0]);
  y[1] = fma2(q, C[0], s.yu[1]); y[1] = fma2(q, C[2], y[1]); y[1] = fma2(q, D[1], y[1]);
  y[2] = fma2(q, C[1], s.yu[2]); y[2] = fma2(q, C[3], y[2]); y[2] = fma2(q, D[2], y[2]);
  y[3] = fma2(q, C[2], s.yu[3]); y[3] = wtap_r<OS>(0.25f, s.xhr, C, y[3]); y[3] = fma2(q, D[3], y[3]);
  if (HASF) {
#pragma unroll
    for (int k = 0; k < 4; ++k) y[k] = sub2(y[k], s.pff[k]);
    if (MAIN) {
      // x.ffill points at f row it + WRING - 1: it joins the copy group of x row it + WRING below
      if (NPDE_W_TMA ? (it + WRING <= x.x_last) : (it + WRING - 1 <= x.f_last)) {  // TMA: x and f rows share a barrier
        wring_fill<EDGE>(x.ring + WRING * WSLOT + ((x.rslot + (WRING - 1) * (unsigned)WSLOT) & RMASK), x.ffill, x);
        x.ffill += x.pitch;
      }
    } else {
      const int nf = it;  // f row used by stage 0 of the next step
      if (nf <= x.x_last && nf >= 1 && nf <= N - 2 && x.ld_ok) ld256(x.pfin, s.pff);
      else s.pff[0] = s.pff[1] = s.pff[2] = s.pff[3] = zero2();
      x.pfin += x.pitch;
    }
  }
#pragma unroll
  for (int k = 0; k < 4; ++k) rc[k] = sub2(y[k], C[k]);
  if (!MAIN && !(r0 >= 1 && r0 <= N - 2)) rc[0] = rc[1] = rc[2] = rc[3] = zero2();
  if (EDGE) {
#pragma unroll
    for (int k = 0; k < 4; ++k) rc[k] = and2(rc[k], x.keep[k], x.keep[k + 4]);
  }
  whalo(rc, x.lane_l, x.lane_r, rl, rr);
  sts128(x.delay + woff + x.a_off, y[0], y[1]);  // y waits in the delay line
  sts128(x.delay + woff + x.b_off, y[2], y[3]);
#pragma unroll
  for (int k = 0; k < 4; ++k) s.yu[k] = mul2(q, C[k]);  // first tap of the next output row: up = x[it-1]
  whalo(D, x.lane_l, x.lane_r, s.xhl, s.xhr);           // row `it` is the next centre row
  // ---- row it-1 is dead: fetch x row it+1 ----
  if (MAIN) {
    if (it + WRING <= x.x_last) {  // x.fill points at row it + WRING: into the slot that was read one step ago
      wring_fill<EDGE>(x.ring + ((x.rslot + (WRING - 1) * (unsigned)WSLOT) & RMASK), x.fill, x);
      x.fill += x.pitch;
      if (NPDE_W_TMA) ++x.pending;
    }
    if (!NPDE_W_TMA) cp_async_commit();
  } else {
    // warm-up / drain / ring-row steps: a plain load (running these steps from the cp.async ring as well
    // measured 4% slower overall)
    if (it + 1 <= x.x_last && x.ld_ok) ld256(x.pin, C);
    else C[0] = C[1] = C[2] = C[3] = zero2();
    x.pin += x.pitch;
  }
}

// ---- conv stage k (1..L): consumes the row (in, il, ir) of stage k-1, finishes row rk.  k < L: the finished row
// goes to (out, ol, orr) (may alias the input); k == L: epilogue with the y row at `roff` of the delay line.
template <int L, int ACT, bool MAIN, bool EDGE, int P, bool STORE = true>
__device__ __forceinline__ void wide_conv_stage(WState<L> &s, WCtx &x, const WideParams &p, int k, int rk, unsigned roff,
                                                const f2 (&in)[4], float il, float ir, f2 (&out)[4], float &ol,
                                                float &orr) {
  const int N = x.N;
#if NPDE_W_TMA
  const float *w = x.wreg[k - 1];
#else
  const float *w = p.w[k - 1];
#endif
  f2 (&mid)[4] = s.S[k - 1][P];      // taps w0..w5 done: finished by this row
  f2 (&top)[4] = s.S[k - 1][P ^ 1];  // taps w0..w2 done
  f2 fin[4] = {mid[0], mid[1], mid[2], mid[3]};
  constexpr bool OS = (L == NPDE_W_OUTER_SCALAR);
  wtaps<OS>(in, il, ir, w[6], w[7], w[8], fin);
  wtaps<OS>(in, il, ir, w[3], w[4], w[5], top);        // becomes the "mid" row of the next step
  wtaps_first<OS>(in, il, ir, w[0], w[1], w[2], mid);  // fresh "top" row of the next step
  if (k < L) {
    if (!MAIN && !(rk >= 1 && rk <= N - 2)) fin[0] = fin[1] = fin[2] = fin[3] = zero2();
#pragma unroll
    for (int j = 0; j < 4; ++j) out[j] = EDGE ? and2(fin[j], x.keep[j], x.keep[j + 4]) : fin[j];
    whalo(out, x.lane_l, x.lane_r, ol, orr);
  } else {
    // MAIN steps also run the band's warm-up (pipeline filling from the zero state, rows above the band) as
    // STORE = false: those output rows are not this piece's, the whole epilogue is skipped
    if (MAIN ? (STORE && (!NPDE_W_SKEW || rk >= x.i0)) : (rk >= x.i0 && rk < x.i1)) {
      float o[8];
      if (MAIN || (rk >= 1 && rk <= N - 2)) {
        f2 ya, yb, yc, yd;
        lds128(x.delay + roff + x.a_off, ya, yb);
        lds128(x.delay + roff + x.b_off, yc, yd);
        o[0] = wadd_act<ACT>(lo(ya), lo(fin[0])); o[4] = wadd_act<ACT>(hi(ya), hi(fin[0]));
        o[1] = wadd_act<ACT>(lo(yb), lo(fin[1])); o[5] = wadd_act<ACT>(hi(yb), hi(fin[1]));
        o[2] = wadd_act<ACT>(lo(yc), lo(fin[2])); o[6] = wadd_act<ACT>(hi(yc), hi(fin[2]));
        o[3] = wadd_act<ACT>(lo(yd), lo(fin[3])); o[7] = wadd_act<ACT>(hi(yd), hi(fin[3]));
      } else {
#pragma unroll
        for (int c = 0; c < 8; ++c) o[c] = (rk == 0 ? x.top : x.bottom);  // ring row: pad then G
      }
      if (EDGE) {  // pad + G on the columns: non-interior columns take the Dirichlet value (or 0: padding)
#pragma unroll
        for (int c = 0; c < 8; ++c) o[c] = wsel_bits(o[c], x.keep[c], x.gval[c]);
      }
      if (x.st_ok) st256_cols(x.pout, o);
    }
    if (MAIN || rk >= 0) x.pout += x.pitch;  // invariant: pout points at row max(rk + 1, 0) afterwards
  }
}

template <int L, int ACT, bool HASF, bool MAIN, bool EDGE, int P, bool STORE = true>
__device__ __forceinline__ void wide_step(WState<L> &s, WCtx &x, const WideParams &p, int it, unsigned &woff) {
  constexpr int WRING = WRingOf<L>::value;
  constexpr int DS = WDelay<L>::slots;
  constexpr unsigned DMASK = DS * (unsigned)WSLOT - 1u;
  constexpr unsigned RMASK = WRING * (unsigned)WSLOT - 1u;
#if NPDE_W_SKEW
  // y of row it-1-2L was written 2L steps ago (DS == 2L: the slot this step's stage 0 overwrites afterwards)
  const unsigned roff = (woff + (DS - 2 * L) * (unsigned)WSLOT) & DMASK;
  // last stage first: stage k reads what stage k-1 left in the previous step, then stage k-1 replaces it
#pragma unroll
  for (int k = L; k >= 1; --k)
    wide_conv_stage<L, ACT, MAIN, EDGE, P, STORE>(s, x, p, k, it - 1 - 2 * k, roff, s.R[k - 1], s.Rl[k - 1], s.Rr[k - 1],
                                           s.R[k < L ? k : 0], s.Rl[k < L ? k : 0], s.Rr[k < L ? k : 0]);
  wide_stage0<L, ACT, HASF, MAIN, EDGE, P>(s, x, it, woff, s.R[0], s.Rl[0], s.Rr[0]);
#else
  const unsigned roff = (woff + (DS - L) * (unsigned)WSLOT) & DMASK;  // y written L steps ago
  f2 rc[4];
  float rl, rr;
  wide_stage0<L, ACT, HASF, MAIN, EDGE, P>(s, x, it, woff, rc, rl, rr);
#pragma unroll
  for (int k = 1; k <= L; ++k)
    wide_conv_stage<L, ACT, MAIN, EDGE, P, STORE>(s, x, p, k, it - 1 - k, roff, rc, rl, rr, rc, rl, rr);
#endif
  if (MAIN) {  // row it+1 has landed (at most WRING - 1 younger groups pending): move it to registers
#if NPDE_W_TMA
    if (x.tma_bytes > 0) mbar_wait(x.bars + (x.rslot >> 7), x.phase);  // 8 bytes per barrier, 1 KB per slot
    --x.pending;
    if (x.rslot == (unsigned)(WRING - 1) * WSLOT) x.phase ^= 1u;  // the ring wraps after this slot
#else
    cp_async_wait<WRING - 1>();
    __syncwarp();  // a lane's bytes were copied by other lanes
#endif
    lds128(x.ring + x.rslot + x.a_off, s.X[P ^ 1][0], s.X[P ^ 1][1]);
    lds128(x.ring + x.rslot + x.b_off, s.X[P ^ 1][2], s.X[P ^ 1][3]);
    if (HASF) {  // f row `it` travelled in the same copy group
      lds128(x.ring + WRING * WSLOT + x.rslot + x.a_off, s.pff[0], s.pff[1]);
      lds128(x.ring + WRING * WSLOT + x.rslot + x.b_off, s.pff[2], s.pff[3]);
    }
    x.rslot = (x.rslot + (unsigned)WSLOT) & RMASK;
  }
  woff = (woff + (unsigned)WSLOT) & DMASK;
}

// rows [i0, i1) of one strip
template <int L, int ACT, bool HASF>
__device__ __forceinline__ void wide_piece(const WideParams &p, int strip, int i0, int i1, char *smem) {
  constexpr int WRING = WRingOf<L>::value;
  constexpr int R = L + 1;
  constexpr int DS = WDelay<L>::slots;
  constexpr int LAT = NPDE_W_SKEW ? 2 * L : L;  // steps between the arrival of x row i+1 and the store of output row i
  WCtx x;
  const int lane = threadIdx.x & 31, N = p.N;
  x.lane_l = (lane + 31) & 31;
  x.lane_r = (lane + 1) & 31;
  x.N = N;
  x.i0 = i0;
  x.i1 = i1;
#ifdef NPDE_W_NOMEM
  x.pitch = 0;  // DEVELOPMENT: every row step touches the same row (compute-only timing; results are garbage)
#else
  x.pitch = p.pitch;
#endif
  x.safe = p.in;
  x.delay = smem;
  x.ring = smem + DS * WSLOT;
  x.rslot = 0;
  const int sw = NPDE_W_TMA ? 0 : (lane >> 2) & 1;  // a bulk copy lands linearly: no chunk swizzle
  x.a_off = 32u * lane + 16u * sw;
  x.b_off = 32u * lane + 16u * (1 - sw);
  x.fill_off = 16u * (lane ^ ((lane >> 3) & 1));
  // virtual group of this lane, its image and first column
  const int v0 = strip * WVALID - 1;
  const int vg = v0 + lane;
  const bool in_range = vg >= 0 && vg < p.VG;
  const int b = in_range ? vg / p.G8 : 0;
  const int col0 = in_range ? (vg - b * p.G8) * WCOLS : 0;
  x.ld_ok = in_range;
  x.st_ok = in_range && lane >= 1 && lane <= WVALID;
  {
    const int va = v0 + (lane >> 1), vb = v0 + 16 + (lane >> 1);  // groups of chunks `lane` and `lane + 32`
    x.fill_ok[0] = va >= 0 && va < p.VG;
    x.fill_ok[1] = vb >= 0 && vb < p.VG;
  }
  const float *bcb = p.bc4 + (size_t)b * 4;
  x.top = __ldg(bcb + 0);
  x.bottom = __ldg(bcb + 1);
  const float left = __ldg(bcb + 2), right = __ldg(bcb + 3);
#pragma unroll
  for (int c = 0; c < 8; ++c) {
    const int col = col0 + c;
    const bool interior = in_range && col >= 1 && col <= N - 2;
    x.keep[c] = interior ? 0xffffffffu : 0u;
    x.gval[c] = !in_range ? 0u : (col == 0 ? __float_as_uint(left) : (col == N - 1 ? __float_as_uint(right) : 0u));
    pin_reg(x.keep[c]);
    pin_reg(x.gval[c]);
  }
  {
    // in-row part of the strip's segment [v0, v0 + 32) groups of the virtual row [0, VG)
    const int g_lo = max(v0, 0), g_hi = min(v0 + 32, p.VG);
    x.tma_src = (g_lo - v0) * WCOLS;                 // floats from the segment start
    x.tma_dst = (g_lo - v0) * WCOLS * 4;             // bytes from the slot start
    x.tma_bytes = max(g_hi - g_lo, 0) * WCOLS * 4;
#pragma unroll
    for (int l = 0; l < 4; ++l)
#pragma unroll
      for (int t = 0; t < 9; ++t) {
        float v = l < L ? p.w[l][t] : 0.0f;
#if NPDE_W_TMA && !defined(NPDE_EMU)
        asm volatile("" : "+f"(v));  // opaque: a per-thread copy, not a uniform value
#endif
        x.wreg[l][t] = v;
      }
    x.pending = 0;
    x.phase = 0;
    x.bars = smem + (DS + (HASF ? 2 : 1) * WRING) * WSLOT;
  }
  // uniform: the strip needs the EDGE code unless its 32 groups are interior groups of ONE image
  const bool edge_strip = wide_strip_is_edge(strip, N, p.G8, p.VG);

  WState<L> s;
#pragma unroll
  for (int h = 0; h < 2; ++h)
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      s.X[h][k] = zero2();
#pragma unroll
      for (int l = 0; l < L; ++l) s.S[l][h][k] = zero2();
    }
#pragma unroll
  for (int k = 0; k < 4; ++k) s.yu[k] = s.pff[k] = zero2();
  s.xhl = s.xhr = 0.0f;
#if NPDE_W_SKEW
#pragma unroll
  for (int l = 0; l < L; ++l) {
    s.Rl[l] = s.Rr[l] = 0.0f;
#pragma unroll
    for (int k = 0; k < 4; ++k) s.R[l][k] = zero2();
  }
#endif
  __syncwarp();  // a previous piece of this warp is done with the shared memory
#pragma unroll
  for (int d = 0; d < DS; ++d) {
    sts128(x.delay + d * WSLOT + x.a_off, zero2(), zero2());
    sts128(x.delay + d * WSLOT + x.b_off, zero2(), zero2());
  }
#if NPDE_W_TMA
#pragma unroll
  for (int d = 0; d < (HASF ? 2 : 1) * WRING; ++d) {  // edge strips copy only the in-row part of a segment
    sts128(x.ring + d * WSLOT + x.a_off, zero2(), zero2());
    sts128(x.ring + d * WSLOT + x.b_off, zero2(), zero2());
  }
#endif
  __syncwarp();

  // input rows it = it_first .. it_last; the output row of a step is it - 1 - LAT
  const int it_first = max(i0 - R, 0);  // rows above the grid are zeros == the initial state
  const int it_last = i1 + LAT;
  // last input row that exists and matters, plus one (fetched, never used: it lets the row-test-free loop run the
  // piece's last step as well)
  x.x_last = min(i1 - 1 + R + 1, N - 1);
  x.f_last = min(x.x_last, N - 2);      // last row of f that is ever read
  const size_t lane_off = (size_t)max(vg, 0) * WCOLS;
  const float *xin = p.in + lane_off;
  const float *fin_ = HASF ? p.f + lane_off : nullptr;
  // chunk `lane` of this strip's row segment; strip 0 starts one group left of the row (never dereferenced there)
  const long seg_off = (long)v0 * WCOLS + (NPDE_W_TMA ? 0 : 4 * lane);  // TMA: the (warp-uniform) segment start
  // the first step runs with parity 0: slot 0 receives row it_first
  if (x.ld_ok) ld256(xin + (size_t)it_first * p.pitch, s.X[0]);
  if (HASF && it_first - 1 >= 1 && x.ld_ok) ld256(fin_ + (size_t)(it_first - 1) * p.pitch, s.pff);
  // running pointers: pin -> row it+1 of x, pfin -> row it of f, pout -> output row it-1-L (clamped at 0)
  x.pin = xin + (size_t)(it_first + 1) * p.pitch;
  x.pfin = HASF ? fin_ + (size_t)it_first * p.pitch : nullptr;
  x.fill = x.ffill = nullptr;
  x.pout = p.out + lane_off + (size_t)max(it_first - 1 - LAT, 0) * p.pitch;
  unsigned woff = 0;

  // MAIN range of `it`: stage rows it-1-LAT .. it-1 interior, output row it-1-LAT inside the piece, x row it+1 exists
  // (the warm-up steps of a band that does not start at the top of the grid qualify too: their stores are suppressed)
  const int main_lo = LAT + 2;
  const int main_hi = min(N - 2, x.x_last - 1);
  // Steps up to i0 + LAT produce rows above the band: the store-free loop runs whole pairs of them.  From it_first
  // (= i0 - R unless the band starts at the top, where nothing is suppressed) a band has LAT + R + 1 = 2L + 2 such
  // steps, an even number, and `it` moves in pairs from it_first, so the first storing pair starts exactly at row i0.
  // (With skewed stages the count is 3L + 2: the storing loop keeps a row test there, see wide_conv_stage.)
  const int warm_hi = min(i0 + LAT, main_hi);
  int it = it_first;  // (it - it_first) is even at the top of this loop: parities 0, 1 per trip
  while (it <= it_last) {
    if (it >= main_lo && it + 1 <= main_hi) {
      // fill the ring: x row `it` is in registers, rows it+1 .. it+WRING-1 start now (f: one row behind)
      x.rslot = 0;
#if NPDE_W_TMA
      // the slot numbering restarts at 0 with every main loop: bring the barrier phases along by re-initialising
      // them (nothing is in flight here: the previous loop drained its copies)
      if (lane == 0)
        for (int d = 0; d < WRING; ++d) mbar_init(x.bars + 8 * d, HASF ? 2u : 1u);
      mbar_fence_init();
      __syncwarp();
      x.pending = 0;
      x.phase = 0;
#endif
      x.fill = p.in + (size_t)(it + 1) * p.pitch + seg_off;
      if (HASF) x.ffill = p.f + (size_t)it * p.pitch + seg_off;
#define NPDE_W_MAIN_LOOP(EDGE_)                                                     \
  do {                                                                              \
    _Pragma("unroll") for (int d = 0; d < WRING - 1; ++d) {                         \
      if (it + 1 + d <= x.x_last) {                                                 \
        wring_fill<EDGE_>(x.ring + d * WSLOT, x.fill, x);                           \
        x.fill += x.pitch;                                                          \
        if (NPDE_W_TMA) ++x.pending;                                                \
      }                                                                             \
      if (HASF && (NPDE_W_TMA ? (it + 1 + d <= x.x_last) : (it + d <= x.f_last))) {  \
        wring_fill<EDGE_>(x.ring + (WRING + d) * WSLOT, x.ffill, x);                \
        x.ffill += x.pitch;                                                         \
      }                                                                             \
      cp_async_commit();                                                            \
    }                                                                               \
    while (it + 1 <= warm_hi) { /* the band's warm-up: an even number of steps (2 LAT + 2), nothing stored */ \
      wide_step<L, ACT, HASF, true, EDGE_, 0, false>(s, x, p, it, woff);            \
      wide_step<L, ACT, HASF, true, EDGE_, 1, false>(s, x, p, it + 1, woff);        \
      it += 2;                                                                      \
    }                                                                               \
    while (it + 1 <= main_hi) {                                                     \
      wide_step<L, ACT, HASF, true, EDGE_, 0>(s, x, p, it, woff);                   \
      wide_step<L, ACT, HASF, true, EDGE_, 1>(s, x, p, it + 1, woff);               \
      it += 2;                                                                      \
    }                                                                               \
  } while (0)
      if (edge_strip) NPDE_W_MAIN_LOOP(true);
      else NPDE_W_MAIN_LOOP(false);
#undef NPDE_W_MAIN_LOOP
#if NPDE_W_TMA
      // rows started but not consumed (the band ended first): their copies must land before the slots are reused
      // or the CTA exits
      while (x.pending > 0) {
        if (x.tma_bytes > 0) mbar_wait(x.bars + (x.rslot >> 7), x.phase);
        if (x.rslot == (unsigned)(WRING - 1) * WSLOT) x.phase ^= 1u;
        x.rslot = (x.rslot + (unsigned)WSLOT) & (unsigned)(WRING * WSLOT - 1);
        --x.pending;
      }
#else
      cp_async_wait<0>();
#endif
      __syncwarp();
      // back to plain loads: row `it` is in registers (slot of parity 0)
      x.pin = xin + (size_t)(it + 1) * p.pitch;
      if (HASF) x.pfin = fin_ + (size_t)it * p.pitch;
    } else {
      wide_step<L, ACT, HASF, false, true, 0>(s, x, p, it, woff);
      if (it + 1 <= it_last) wide_step<L, ACT, HASF, false, true, 1>(s, x, p, it + 1, woff);
      it += 2;
    }
  }
}

template <int L, int ACT, bool HASF>
__global__ void __launch_bounds__(npde::STREAM_THREADS, NPDE_WIDE_MIN_CTAS)
k_conv_wide(const NPDE_GRID_CONSTANT WideParams p) {
  NPDE_DYN_SMEM(float4, wsm);
  // strip index fastest: the warps of a band walk down the same rows at the same time, so their 1 KB row segments
  // are contiguous in memory (one long DRAM burst per row of the band).  The work item depends on blockIdx only:
  // ptxas can prove that the strip bookkeeping and the loop versioning are warp-uniform and keeps the weights in
  // uniform registers.
  // Programmatic dependent launch: this grid may become resident while the previous Phi_H launch of the same stream
  // is still draining (its launch latency and CTA ramp hide behind that tail); nothing of global memory is touched
  // before the previous grid has completed and flushed, and the next launch may be scheduled from now on.
  pdl_launch_dependents();
  pdl_wait_prior_grids();
  int strip, band;
  const int uniform_ctas = p.strips * p.bands;
  if ((int)blockIdx.x < uniform_ctas) {
    strip = (int)blockIdx.x % p.strips;
    band = (int)blockIdx.x / p.strips;
  } else {  // the extra bands of the edge strips
    const int e = (int)blockIdx.x - uniform_ctas;
    strip = p.edge_strip[e % p.n_edge];
    band = p.bands + e / p.n_edge;
  }
  const int rows = (p.bands_edge > p.bands && wide_strip_is_edge(strip, p.N, p.G8, p.VG)) ? p.band_rows_edge : p.band_rows;
  const int i0 = band * rows, i1 = min(i0 + rows, p.N);
  if (i0 < i1) wide_piece<L, ACT, HASF>(p, strip, i0, i1, reinterpret_cast<char *>(wsm));
}

// ---- layout conversion: reference (B,N,N) <-> W8 ----------------------------------------------------------
// One warp per (row, 32 groups): coalesced 4-byte accesses on the reference side (rows of a 2^n+1 grid are only
// 4-byte aligned), a 1 KB staging row in shared memory, one 256-bit access per lane on the W8 side.
// Padding columns of the W8 field are written as zeros.
constexpr int PACK_WARPS = 4;
// One warp per (row, image, 32 groups of that image): the image index is warp-uniform, so the element loop has no
// integer division (a warp that straddled two images needed one per element: 185 -> ~130 us per 64 x 1025^2 field).
template <bool TO_WIDE>
__global__ void __launch_bounds__(32 * PACK_WARPS)
k_wide_convert(const float *__restrict__ src, float *__restrict__ dst, int B, int N, int G8, int VG, int segs_per_img) {
  __shared__ __align__(16) float stage[PACK_WARPS][256];
  const int lane = threadIdx.x & 31, wrp = threadIdx.x >> 5;
  const int seg = blockIdx.x * PACK_WARPS + wrp;  // 32 groups of one image
  const int row = blockIdx.y;
  const int b = seg / segs_per_img;
  if (b >= B) return;
  const int g0 = (seg - b * segs_per_img) * 32;   // first group of the segment inside the image
  const int ng = min(32, G8 - g0);                // groups that exist
  const size_t wrow = (size_t)row * VG * 8 + ((size_t)b * G8 + g0) * 8;  // W8 side: start of the segment
  const float *rrow = src + ((size_t)b * N + row) * N;                   // reference side: start of the image row
  float *rrow_out = dst + ((size_t)b * N + row) * N;
  float *st = stage[wrp];
  if (TO_WIDE) {
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      const int e = 32 * k + lane;  // element of the segment in column order
      const int col = g0 * 8 + e, c = e & 7;
      const float v = col < N ? __ldg(rrow + col) : 0.0f;
      st[(e & ~7) | ((c & 3) << 1) | (c >> 2)] = v;  // (c0,c4,c1,c5,c2,c6,c3,c7)
    }
    __syncwarp();
    if (lane < ng) {
      const float4 a = *reinterpret_cast<const float4 *>(st + 8 * lane), bq = *reinterpret_cast<const float4 *>(st + 8 * lane + 4);
      float4 *d = reinterpret_cast<float4 *>(dst + wrow + (size_t)lane * 8);
      d[0] = a;
      d[1] = bq;
    }
  } else {
    if (lane < ng) {
      const float4 *sp = reinterpret_cast<const float4 *>(src + wrow + (size_t)lane * 8);
      *reinterpret_cast<float4 *>(st + 8 * lane) = __ldg(sp);
      *reinterpret_cast<float4 *>(st + 8 * lane + 4) = __ldg(sp + 1);
    }
    __syncwarp();
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      const int e = 32 * k + lane;
      const int col = g0 * 8 + e, c = e & 7;
      if (col < N && (e >> 3) < ng) rrow_out[col] = st[(e & ~7) | ((c & 3) << 1) | (c >> 2)];
    }
  }
}

// ---- fd_error('max') on a W8 field (utils/heat_utils.py:108-121, square domain): a solver loop that keeps its
// iterate in the packed layout checks the residual without unpacking.  One thread per group of eight columns walks
// down RES_ROWS rows with a three-row register window; the outer neighbours c-1 / c8 are the last / first element
// of the adjacent groups of the same image (L1 hits: the neighbour threads load them as part of their own group).
// Per cell the loss-kernel chain npde_residual (bit-identical to k_fd_error); maximum per image with an integer
// atomicMax on the float bits (non-negative floats order like integers, a maximum is order independent).
constexpr int RES_ROWS = 32, RES_THREADS = 128;
__device__ __forceinline__ void res_load(const float *row, int g, int G8, float (&c)[10]) {
  // c[1..8] = columns 0..7 of group g, c[0] = column -1, c[9] = column 8 (0 outside the image)
  const float4 a = __ldg(reinterpret_cast<const float4 *>(row + 8 * g)), b = __ldg(reinterpret_cast<const float4 *>(row + 8 * g + 4));
  c[1] = a.x; c[5] = a.y; c[2] = a.z; c[6] = a.w; c[3] = b.x; c[7] = b.y; c[4] = b.z; c[8] = b.w;
  c[0] = g > 0 ? __ldg(row + 8 * (g - 1) + 7) : 0.0f;
  c[9] = g + 1 < G8 ? __ldg(row + 8 * (g + 1)) : 0.0f;
}
__global__ void __launch_bounds__(RES_THREADS)
k_wide_resid(const float *__restrict__ x, const float *__restrict__ f, float *err, int B, int N, int G8, int VG) {
  const int vg = blockIdx.x * RES_THREADS + threadIdx.x;
  const int i0 = 1 + blockIdx.y * RES_ROWS, i1 = min(i0 + RES_ROWS, N - 1);
  float vmax = 0.0f;
  int b = min(vg, VG - 1) / G8;
  if (vg < VG && i0 < i1) {
    const int g = vg - b * G8;
    const size_t pitch = (size_t)VG * 8;
    const float *img = x + (size_t)b * G8 * 8;           // this image's first group in a virtual row
    const float *fim = f ? f + (size_t)b * G8 * 8 : nullptr;
    float up[10], ce[10], dn[10];
    res_load(img + (size_t)(i0 - 1) * pitch, g, G8, up);
    res_load(img + (size_t)i0 * pitch, g, G8, ce);
    for (int i = i0; i < i1; ++i) {
      res_load(img + (size_t)(i + 1) * pitch, g, G8, dn);
      float fv[8];
      if (fim) {
        const float *fr = fim + (size_t)i * pitch + 8 * g;
        const float4 a = __ldg(reinterpret_cast<const float4 *>(fr)), q = __ldg(reinterpret_cast<const float4 *>(fr + 4));
        fv[0] = a.x; fv[4] = a.y; fv[1] = a.z; fv[5] = a.w; fv[2] = q.x; fv[6] = q.y; fv[3] = q.z; fv[7] = q.w;
      }
#pragma unroll
      for (int c = 0; c < 8; ++c) {
        const int col = 8 * g + c;
        if (col >= 1 && col <= N - 2) {
          float v = npde_residual(up[c + 1], ce[c], ce[c + 1], ce[c + 2], dn[c + 1]);
          if (fim) v = __fsub_rn(v, fv[c]);
          vmax = fmaxf(vmax, fabsf(v));
        }
      }
#pragma unroll
      for (int c = 0; c < 10; ++c) { up[c] = ce[c]; ce[c] = dn[c]; }
    }
  }
  // lanes of a warp may belong to two images (an image boundary inside the warp; b grows with the lane)
  const int b_first = __shfl_sync(NPDE_FULL_MASK, b, 0), b_last = __shfl_sync(NPDE_FULL_MASK, b, 31);
  if (b_first == b_last) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) vmax = fmaxf(vmax, __shfl_xor_sync(NPDE_FULL_MASK, vmax, o));
    if ((threadIdx.x & 31) == 0 && vmax > 0.0f) atomicMax(reinterpret_cast<int *>(err) + b, __float_as_int(vmax));
  } else if (vmax > 0.0f) {
    atomicMax(reinterpret_cast<int *>(err) + b, __float_as_int(vmax));
  }
}

struct WideGeom {
  int G8, VG, pitch, strips;
};
WideGeom wide_geom(int B, int N) {
  WideGeom g;
  g.G8 = (N + 7) / 8;
  g.VG = B * g.G8;
  g.pitch = g.VG * 8;
  g.strips = (g.VG + WVALID - 1) / WVALID;
  return g;
}

template <int L>
constexpr size_t wide_smem(bool hasf) {
  constexpr int WRING = WRingOf<L>::value;
  return (size_t)WSLOT * (WDelay<L>::slots + (hasf ? 2 : 1) * WRING) + (NPDE_W_TMA ? 8 * WRING : 0);
}

// resident warps of the kernel on the current device (cached per device)
int wide_slots(const void *kernel, size_t smem, std::atomic<int> *cache /* one entry per device */) {
#ifdef NPDE_EMU
  (void)kernel; (void)smem; (void)cache;
  return 8;  // small on purpose: several bands per strip in the emulator
#else
  int dev = 0, sms = 0, blocks = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return 148 * 8;
  const int hit = cache[dev].load(std::memory_order_relaxed);
  if (hit > 0) return hit;
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&blocks, kernel, npde::STREAM_THREADS, smem) != cudaSuccess || blocks < 1)
    blocks = 1;
  // one-warp CTAs are dealt to the four SM sub-partitions in turn: keep their count per SM a multiple of four
  if (blocks >= 4) blocks &= ~3;
  cache[dev].store(sms * blocks, std::memory_order_relaxed);
  return sms * blocks;
#endif
}

// 4 x SMs of the current device
int wide_layer() {
#ifdef NPDE_EMU
  return 4;
#else
  static std::atomic<int> cache[64];
  int dev = 0, sms = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return 4 * 148;
  const int hit = cache[dev].load(std::memory_order_relaxed);
  if (hit > 0) return hit;
  if (cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || sms < 1) sms = 148;
  cache[dev].store(4 * sms, std::memory_order_relaxed);
  return 4 * sms;
#endif
}

int wide_env_int(const char *name, int dflt) {
#ifdef NPDE_EMU
  (void)name;
  return dflt;
#else
  const char *v = getenv(name);
  return v ? atoi(v) : dflt;
#endif
}

template <int L, int ACT, bool HASF>
int launch_wide_f(WideParams &p, cudaStream_t st) {
  static std::atomic<int> cache[64];  // zero-initialised (static storage)
  const size_t smem = wide_smem<L>(HASF);
  auto kern = k_conv_wide<L, ACT, HASF>;
#ifndef NPDE_EMU
  if (smem > 48 * 1024) cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
#endif
  int slots = wide_slots((const void *)kern, smem, cache);
  // Work partition: bands x strips one-warp CTAs in ONE wave, the strips of a band in lock step.  How many bands?
  // A piece of h rows costs h + warm-up row steps (2L + 3: halo rows and pipeline fill),
  // and a row step of w = ceil(CTAs / (4 x SMs)) resident warps per SM sub-partition costs each of them 1.3 / 2.0 /
  // 2.76 units for w = 1 / 2 / 3 (measured: a lone warp is latency bound, the third warp still adds throughput):
  // the launch takes c(w) x (h + warm-up).  Minimising that over the band count picks 6 bands on three warps for
  // 64 x 1025^2 (108 us; 4 bands on two warps: 115 us) and 33 bands of 32 rows on two warps for 8 x 1025^2
  // (17.9 us; 41 bands of 25 rows on three warps: 21.8 us) -- profiles/r2_wide_ab.txt.
  // Everything tried to also use the warp slots that whole bands leave idle lost more on the memory side than it
  // won (flattened (strip,row) runs, diagonal lock step, a flattened tail), except one: more, shorter bands for the
  // slower edge strips (below).  NPDE_WIDE_BANDS=b (tuning) forces b bands.
#ifdef NPDE_EMU
  const int minrows = 6;
  const int layer = 4;
#else
  const int minrows = wide_env_int("NPDE_WIDE_MINROWS", 16);
  const int layer = wide_layer();  // one-warp CTAs that put one warp on every SM sub-partition
#endif
  const int nstrips = p.strips > 0 ? p.strips : 1;
  const int force_bands = wide_env_int("NPDE_WIDE_BANDS", 0);
  // edge strips (a ring / padding column inside): the masked loop costs EDGE_COST x the interior loop per row step
  const double EDGE_COST = 1.18;
  int n_edge = 0;
  if (force_bands <= 0 && p.strips <= 65535 && wide_env_int("NPDE_WIDE_EDGE_BANDS", 1) != 0) {
    for (int s = 0; s < p.strips && n_edge <= WMAX_EDGE; ++s)
      if (wide_strip_is_edge(s, p.N, p.G8, p.VG)) {
        if (n_edge < WMAX_EDGE) p.edge_strip[n_edge] = (unsigned short)s;
        ++n_edge;
      }
    if (n_edge >= p.strips || n_edge > WMAX_EDGE) n_edge = 0;  // all strips alike (or too many to list): uniform bands
  }
  const int n_int = nstrips - n_edge;
  auto real_bands = [&](int b) {  // the bands that exist when N rows are cut into pieces of ceil(N / b) rows
    const int h = (p.N + b - 1) / b;
    return (p.N + h - 1) / h;
  };
  int bmax = p.N / minrows;
  if (bmax < 1) bmax = 1;
  int bands = 1, bands_e = 1;
  {
    // interior strips get b bands, edge strips as many (>= b) as still fit into the same warps-per-sub-partition layer:
    // minimise c(w) x max(h_i + warm, EDGE_COST (h_e + warm)) over b
    const double warm = 2 * L + 3;
    const int wmax_dev = slots / layer > 0 ? slots / layer : 1;
    const int wcap = std::min(wide_env_int("NPDE_WIDE_W", 16), wmax_dev);  // tuning: at most this many warps per sub-partition
    double best = 0.0;
    for (int b = 1; b <= bmax; ++b) {
      if (real_bands(b) != b) continue;
      const long base = (long)nstrips * b;
      int w = (int)((base + layer - 1) / layer);
      if (w > wcap) {
        if (b > 1) continue;
        w = wcap;  // a single band that does not fit into one wave: nothing to choose
      }
      const long cap = std::min<long>(slots, (long)w * layer);
      int be = b;
      if (n_edge > 0) {
        be = b + (int)((cap - base) / n_edge);
        if (be > bmax) be = bmax;
        while (be > b && real_bands(be) != be) --be;
        // more edge bands than balance needs only add warm-up work
        while (be > b && EDGE_COST * ((p.N + be - 2) / (be - 1) + warm) <= (p.N + b - 1) / b + warm &&
               real_bands(be - 1) == be - 1)
          --be;
      }
      const double hi = (p.N + b - 1) / b + warm, he = EDGE_COST * ((p.N + be - 1) / be + warm);
      const double cost = (w < 2 ? 1.3 : 2.0 + 0.76 * (w - 2)) * (n_edge > 0 && he > hi ? he : hi);
      if (best == 0.0 || cost < best) {
        best = cost;
        bands = b;
        bands_e = be;
      }
    }
  }
  if (force_bands > 0) bands = bands_e = force_bands;
  // The block scheduler fills an SM to its occupancy limit before it moves on: a grid meant to put w warps on every
  // sub-partition must not be ALLOWED more than 4w CTAs per SM, or it leaves whole SMs idle.  The limit is set
  // through the shared-memory request (the unused part is never touched).
  size_t smem_launch = smem;
#ifndef NPDE_EMU
  {
    int w = (int)(((long)n_int * bands + (long)n_edge * bands_e + layer - 1) / layer);
    const int wmax = slots / layer;
    if (w < 1) w = 1;
    if (w < wmax && force_bands <= 0) {
      const size_t per_cta = (size_t)(228 * 1024) / (size_t)(4 * w) - 1024 - 256;  // 1 KB per CTA is the system's
      if (per_cta > smem_launch) smem_launch = per_cta;
    }
    if (smem_launch > 48 * 1024) {
      static std::atomic<int> optin[64];
      int dev = 0;
      cudaGetDevice(&dev);
      if (dev < 0 || dev >= 64 || optin[dev].load(std::memory_order_relaxed) < (int)smem_launch) {
        cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_launch);
        if (dev >= 0 && dev < 64) optin[dev].store((int)smem_launch, std::memory_order_relaxed);
      }
    }
  }
#endif
  p.band_rows = (p.N + bands - 1) / bands;
  p.bands = (p.N + p.band_rows - 1) / p.band_rows;
  p.bands_edge = p.bands;
  p.band_rows_edge = p.band_rows;
  p.n_edge = 0;
  long ctas = (long)p.bands * p.strips;
  if (n_edge > 0 && bands_e > p.bands) {
    const int rows = (p.N + bands_e - 1) / bands_e;
    const int nb = (p.N + rows - 1) / rows;
    if (nb > p.bands) {
      p.bands_edge = nb;
      p.band_rows_edge = rows;
      p.n_edge = n_edge;
      ctas += (long)n_edge * (nb - p.bands);
    }
  }
  dim3 grid((unsigned)ctas), block(npde::STREAM_THREADS);
#ifndef NPDE_EMU
  // Programmatic stream serialization (see k_conv_wide); not inside a stream capture, where the plain edge is kept.
  cudaStreamCaptureStatus cap = cudaStreamCaptureStatusNone;
  if (wide_env_int("NPDE_WIDE_PDL", 1) != 0 && cudaStreamIsCapturing(st, &cap) == cudaSuccess &&
      cap == cudaStreamCaptureStatusNone) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid;
    cfg.blockDim = block;
    cfg.dynamicSmemBytes = smem_launch;
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    npde_note_launch();
    const cudaError_t e = cudaLaunchKernelEx(&cfg, kern, p);
    if (e != cudaSuccess) return (int)e;
    return NPDE_OK;
  }
#endif
  NPDE_LAUNCH((k_conv_wide<L, ACT, HASF>), grid, block, smem_launch, st, p);
  NPDE_CHECK_LAUNCH();
  return NPDE_OK;
}

template <int L>
int launch_wide(WideParams &p, cudaStream_t st) {
  switch (p.act) {
    case NPDE_ACT_NONE:
      return p.f ? launch_wide_f<L, NPDE_ACT_NONE, true>(p, st) : launch_wide_f<L, NPDE_ACT_NONE, false>(p, st);
    case NPDE_ACT_CLAMP:
      return p.f ? launch_wide_f<L, NPDE_ACT_CLAMP, true>(p, st) : launch_wide_f<L, NPDE_ACT_CLAMP, false>(p, st);
    case NPDE_ACT_SIGMOID:
      return p.f ? launch_wide_f<L, NPDE_ACT_SIGMOID, true>(p, st) : launch_wide_f<L, NPDE_ACT_SIGMOID, false>(p, st);
    default: return NPDE_ERR_ACT;
  }
}

}  // namespace

namespace npde {

size_t wide_field_floats(int B, int N) {
  const WideGeom g = wide_geom(B, N);
  return (size_t)g.pitch * N;
}

bool conv_wide_supported(int bc_kind, int n_layers, const float *w_host, int B, int N) {
  return bc_kind == NPDE_BC_SQUARE && n_layers >= 1 && n_layers <= 4 && w_host != nullptr && B >= 1 && N >= 3;
}

int wide_pack(const float *x_ref, float *wide, int B, int N, cudaStream_t st) {
  const WideGeom g = wide_geom(B, N);
  const int spi = (g.G8 + 31) / 32;  // segments of 32 groups per image
  dim3 grid((unsigned)((B * spi + PACK_WARPS - 1) / PACK_WARPS), (unsigned)N), block(32 * PACK_WARPS);
  NPDE_LAUNCH(k_wide_convert<true>, grid, block, 0, st, x_ref, wide, B, N, g.G8, g.VG, spi);
  NPDE_CHECK_LAUNCH();
  return NPDE_OK;
}

int wide_unpack(const float *wide, float *x_ref, int B, int N, cudaStream_t st) {
  const WideGeom g = wide_geom(B, N);
  const int spi = (g.G8 + 31) / 32;
  dim3 grid((unsigned)((B * spi + PACK_WARPS - 1) / PACK_WARPS), (unsigned)N), block(32 * PACK_WARPS);
  NPDE_LAUNCH(k_wide_convert<false>, grid, block, 0, st, wide, x_ref, B, N, g.G8, g.VG, spi);
  NPDE_CHECK_LAUNCH();
  return NPDE_OK;
}

int wide_fd_error_max(const float *wide, const float *f_wide, float *err, int B, int N, cudaStream_t st) {
  if (!wide || !err) return NPDE_ERR_NULL;
  if (((uintptr_t)wide | (uintptr_t)f_wide) & 31) return NPDE_ERR_ARG;
  const WideGeom g = wide_geom(B, N);
#ifdef NPDE_EMU
  memset(err, 0, sizeof(float) * (size_t)B);
#else
  const cudaError_t e = cudaMemsetAsync(err, 0, sizeof(float) * (size_t)B, st);
  if (e != cudaSuccess) return (int)e;
#endif
  if (N < 3) return NPDE_OK;
  dim3 grid((unsigned)((g.VG + RES_THREADS - 1) / RES_THREADS), (unsigned)((N - 2 + RES_ROWS - 1) / RES_ROWS)), block(RES_THREADS);
  NPDE_LAUNCH(k_wide_resid, grid, block, 0, st, wide, f_wide, err, B, N, g.G8, g.VG);
  NPDE_CHECK_LAUNCH();
  return NPDE_OK;
}

int conv_wide(const float *in, float *out, const float *bc4, const float *f_wide, const float *w_host, int n_layers,
              int act, int B, int N, cudaStream_t st) {
  if (!in || !out || !bc4 || !w_host) return NPDE_ERR_NULL;
  if (in == out) return NPDE_ERR_ARG;
  if (((uintptr_t)in | (uintptr_t)out | (uintptr_t)f_wide) & 31) return NPDE_ERR_ARG;
  const WideGeom g = wide_geom(B, N);
  WideParams p;
  p.in = in; p.out = out; p.f = f_wide; p.bc4 = bc4;
  p.B = B; p.N = N; p.G8 = g.G8; p.VG = g.VG; p.pitch = g.pitch; p.strips = g.strips;
  p.bands = 1; p.band_rows = N; p.act = act;
  for (int l = 0; l < 4; ++l)
    for (int t = 0; t < 9; ++t) p.w[l][t] = (l < n_layers) ? w_host[l * 9 + t] : 0.0f;
  switch (n_layers) {
    case 1: return launch_wide<1>(p, st);
    case 2: return launch_wide<2>(p, st);
    case 3: return launch_wide<3>(p, st);
    case 4: return launch_wide<4>(p, st);
    default: return NPDE_ERR_LAYERS;
  }
}

}  // namespace npde
