// npde_api.cu -- the extern "C" surface of libnpde_b200 (include/npde.h):
// argument validation, workspace carving and the host-side schedules that
// string kernels together (V-cycle, UNet, backward, the k-step loop).  Nothing
// here synchronises or allocates.
#include <stdlib.h>

#include <string.h>

#include <atomic>
#include <mutex>
#include <new>
#include <vector>

#include "npde_common.cuh"
#include "npde_internal.h"

static std::atomic<unsigned long long> g_launches{0};
void npde_note_launch() { g_launches.fetch_add(1, std::memory_order_relaxed); }

namespace {

// Bump allocator over the caller's workspace.  With base == nullptr it only
// measures, so npde_workspace_bytes and the real carve share one code path.
struct Bump {
  char *base;
  size_t off = 0;
  explicit Bump(void *b) : base(reinterpret_cast<char *>(b)) {}
  template <class T> T *take(size_t count) {
    size_t bytes = npde_align_up(sizeof(T) * count, 256);
    T *p = base ? reinterpret_cast<T *>(base + off) : nullptr;
    off += bytes;
    return p;
  }
  void *take_bytes(size_t bytes) { return take<char>(bytes); }
};

inline int ws_ok(const void *ws, size_t have, size_t need) {
  if (need == 0) return NPDE_OK;
  if (!ws || ((uintptr_t)ws & 255) || have < need) return NPDE_ERR_WORKSPACE;
  return NPDE_OK;
}

inline cudaStream_t S(void *stream) { return reinterpret_cast<cudaStream_t>(stream); }

inline int d2d(void *dst, const void *src, size_t bytes, cudaStream_t st) {
  if (dst == src || bytes == 0) return NPDE_OK;
  cudaError_t e = cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToDevice, st);
  return e == cudaSuccess ? NPDE_OK : (int)e;
}

#define TRY(expr)            \
  do {                       \
    int rc__ = (expr);       \
    if (rc__) return rc__;   \
  } while (0)

inline const float *mask0(const float *bc, int bc_kind, int N) {
  return (bc && bc_kind == NPDE_BC_MASK) ? bc + (size_t)N * N : nullptr;
}

// ----------------------------------------------------------------- Conv ---
struct ConvFwdWs {
  float *y_int, *t0, *t1;
  void carve(Bump &b, int B, int N) {
    size_t n = (size_t)B * (N - 2) * (N - 2);
    y_int = b.take<float>(n);
    t0 = b.take<float>(n);
    t1 = b.take<float>(n);
  }
};

// H on a tensor that is ALREADY multiplied by fg (or on the square path).
// src may be const input; result lands in `out`.
int conv_layers(const float *src, const float *mask, size_t mbs, const float *w, int L, float *t0, float *t1,
                float *out, int B, int S, cudaStream_t st) {
  const float *cur = src;
  for (int l = 0; l < L; ++l) {
    float *dst = (l == L - 1) ? out : (cur == t0 ? t1 : t0);
    TRY(npde::conv3x3(cur, S, w + 9 * l, 1, 1, mask, mbs, dst, S, B, st));
    cur = dst;
  }
  if (L == 0) TRY(d2d(out, src, sizeof(float) * (size_t)B * S * S, st));
  return NPDE_OK;
}

int conv_H_impl(const float *r, const float *bc, int bc_kind, const float *w, int L, float *out, int B, int N,
                void *ws, cudaStream_t st) {
  Bump b(ws);
  ConvFwdWs c;
  c.carve(b, B, N);
  int S = N - 2;
  const float *mask = mask0(bc, bc_kind, N);
  size_t mbs = 2 * (size_t)N * N;
  const float *src = r;
  if (mask) {  // conv_iterator.py:23-25
    TRY(d2d(c.y_int, r, sizeof(float) * (size_t)B * S * S, st));
    TRY(npde::mul_fg(c.y_int, mask, mbs, S, B, st));
    src = c.y_int;
  }
  return conv_layers(src, mask, mbs, w, L, c.t0, c.t1, out, B, S, st);
}

int conv_step_layered(const float *x, const float *bc, int bc_kind, const float *f, const float *w, int L,
                      int act, float *y, int B, int N, void *ws, cudaStream_t st) {
  Bump b(ws);
  ConvFwdWs c;
  c.carve(b, B, N);
  int S = N - 2;
  const float *mask = mask0(bc, bc_kind, N);
  size_t mbs = 2 * (size_t)N * N;
  TRY(npde::psi_residual(x, f, mask, mbs, c.y_int, c.t0, B, N, st));
  float *h = (L == 0) ? c.t0 : ((L & 1) ? c.t1 : c.t0);
  // ping-pong t0 -> t1 -> t0 ...; the last layer writes into `h`
  const float *cur = c.t0;
  for (int l = 0; l < L; ++l) {
    float *dst = (cur == c.t0) ? c.t1 : c.t0;
    TRY(npde::conv3x3(cur, S, w + 9 * l, 1, 1, mask, mbs, dst, S, B, st));
    cur = dst;
  }
  (void)h;
  return npde::combine(c.y_int, cur, bc, bc_kind, act, y, B, N, st);
}

struct ConvBwdWs {
  float *y_int, *gz, *g, *t;
  float *rs[NPDE_MAX_CONV_LAYERS + 1];
  void *wg;
  size_t wg_bytes;
  unsigned *counters;
  void carve(Bump &b, int B, int N, int L) {
    size_t n = (size_t)B * (N - 2) * (N - 2);
    y_int = b.take<float>(n);
    gz = b.take<float>(n);
    g = b.take<float>(n);
    t = b.take<float>(n);
    for (int l = 0; l <= L; ++l) rs[l] = b.take<float>(n);
    size_t a = npde::weight_grad_ws_bytes(B, N - 2), c = npde::bwd_layer_ws_bytes(B, N - 2);
    size_t d = npde::conv_bwd_stream_ws_bytes(B, N, 3);  // the fused streaming backward (npde_bwd_stream.cu)
    wg_bytes = a > c ? a : c;
    if (d > wg_bytes) wg_bytes = d;
    wg = b.take_bytes(wg_bytes);
    counters = b.take<unsigned>(NPDE_MAX_CONV_LAYERS);
  }
};

// ------------------------------------------------------------- V-cycle ---
// Level fields live in the pair-permuted padded layout (the Jacobi stream kernel then moves 128 bits per
// lane; restriction / prolongation read and write the same views).  On mask geometries every level also
// keeps permuted copies of its two bc planes for the smoother (made once per V-cycle and level).
struct MgLevel {
  int N;
  npde::FieldViewMut p[2];  // ping-pong fields
  float *f;                 // 4^l * injected source (level >= 1), or NULL
  float *bc;                // injected [values, mask] (mask path, level >= 1)
  float *vperm, *mperm;     // mask path: the level's bc planes in the permuted layout (smoother's 128-bit loads)
};
struct MgWs {
  MgLevel lv[NPDE_MAX_LEVELS];
  void carve(Bump &b, int B, int N, int bc_kind, int L, bool has_f) {
    int n = N;
    const bool mask = bc_kind == NPDE_BC_MASK;
    for (int l = 0; l < L; ++l) {
      size_t cells = (size_t)B * n * n;
      int pitch = (n + 3) & ~3;
      long bstride = (long)npde_align_up((size_t)pitch * n, 64);
      lv[l].N = n;
      for (int h = 0; h < 2; ++h) lv[l].p[h] = npde::FieldViewMut{b.take<float>((size_t)B * bstride), pitch, bstride, 1};
      lv[l].f = (l > 0 && has_f) ? b.take<float>(cells) : nullptr;
      lv[l].bc = (l > 0 && mask) ? b.take<float>(2 * cells) : nullptr;
      lv[l].vperm = mask ? b.take<float>((size_t)B * bstride) : nullptr;
      lv[l].mperm = mask ? b.take<float>((size_t)B * bstride) : nullptr;
      n = (n - 1) / 2 + 1;
    }
  }
};

// n Jacobi steps starting from `cur` on the level's ping-pong fields (never writing into an external source);
// `last`, if given, receives the final iterate directly.  cur is updated to the view that holds the result.
int smooth(npde::FieldView &cur, MgLevel &lv, const float *bc, int bc_kind, const float *f, int n, int B,
           const npde::FieldViewMut *last, cudaStream_t st);

int vcycle(MgWs &m, int l, int L, const npde::FieldView &src, const float *bc, const float *f, int bc_kind, int pre,
           int post, int B, const npde::FieldViewMut *out, npde::FieldView *result, cudaStream_t st) {
  MgLevel &lv = m.lv[l];
  npde::FieldView cur = src;
  const bool coarsest = l == L - 1;
  if (bc_kind == NPDE_BC_MASK && pre + post > 0) {
    npde::BcFields nat = npde::bc_fields(bc, bc_kind, lv.N);
    TRY(npde::copy_field(nat.values, npde::FieldViewMut{lv.vperm, lv.p[0].pitch, lv.p[0].bstride, 1}, B, lv.N, st));
    TRY(npde::copy_field(nat.mask, npde::FieldViewMut{lv.mperm, lv.p[0].pitch, lv.p[0].bstride, 1}, B, lv.N, st));
  }
  // the very last operation of the level may write the caller's tensor directly
  TRY(smooth(cur, lv, bc, bc_kind, f, pre, B, (coarsest && post == 0) ? out : nullptr, st));  // jacobi_iterator.py:41-42
  if (!coarsest) {
    MgLevel &nx = m.lv[l + 1];
    const float *f_sub = nullptr, *bc_sub = bc;
    if (f) {  // :46-47
      TRY(npde::subsample(f, nx.f, B, lv.N, 4.0f, st));
      f_sub = nx.f;
    }
    if (bc_kind == NPDE_BC_MASK) {  // :51-54 values and mask are both injected
      TRY(npde::subsample(bc, nx.bc, 2 * B, lv.N, 1.0f, st));
      bc_sub = nx.bc;
    }
    TRY(npde::restrict_view(cur, bc_sub, bc_kind, nx.p[0], B, lv.N, st));  // :58
    npde::FieldView sub;
    TRY(vcycle(m, l + 1, L, npde::as_const(nx.p[0]), bc_sub, f_sub, bc_kind, pre, post, B, nullptr, &sub, st));  // :60
    const npde::FieldViewMut &dst = (post == 0 && out) ? *out : ((cur.p == lv.p[0].p) ? lv.p[1] : lv.p[0]);
    TRY(npde::prolong_view(sub, bc, bc_kind, dst, B, nx.N, st));  // :62 replaces x
    cur = npde::as_const(dst);
  }
  TRY(smooth(cur, lv, bc, bc_kind, f, post, B, out, st));  // :65-66
  *result = cur;
  return NPDE_OK;
}

}  // namespace

namespace npde {
// n Jacobi steps src -> buf0 / buf1 (ping-pong views); *result is the view that holds the last iterate; the
// final launch writes `last` if given.  Square domain without a source runs temporally blocked launches of
// depth 4/2/1; otherwise one launch per step.
int jacobi_steps_view(const FieldView &src, const FieldViewMut &buf0, const FieldViewMut &buf1, const BcFields &bcf,
                      const float *f, int n, int B, int N, const FieldViewMut *last, FieldView *result,
                      cudaStream_t st) {
  FieldView cur = src;
  int flip = 0, left = n;
  const bool square = bcf.bc4 != nullptr;
  while (left > 0) {
    // temporal blocking everywhere: square domain and mask geometries, with and without a source term
    (void)square;
    int depth = left >= 4 ? 4 : (left >= 2 ? 2 : 1);
    const bool final_launch = left - depth == 0;
    const FieldViewMut &dst = (final_launch && last) ? *last : (flip ? buf1 : buf0);
    if (!dst.p || dst.p == cur.p) return NPDE_ERR_ARG;
    TRY(jacobi_stream(cur, dst, bcf, ref_view(f, N), depth, B, N, st));
    left -= depth;
    cur = as_const(dst);
    flip ^= 1;
  }
  *result = cur;
  return NPDE_OK;
}

int jacobi_steps(const float *src, float *buf0, float *buf1, const float *bc, int bc_kind, const float *f, int n,
                 int B, int N, const float **result, cudaStream_t st) {
  if (bc_kind == NPDE_BC_MASK) {  // single steps on the reference layout: the coalesced thread-per-cell kernel
    const float *cur = src;
    float *dst = buf0;
    for (int t = 0; t < n; ++t) {
      if (!dst) return NPDE_ERR_ARG;
      TRY(fd_step_simple(cur, bc, bc_kind, f, dst, B, N, st));
      cur = dst;
      dst = (dst == buf0) ? buf1 : buf0;
    }
    *result = cur;
    return NPDE_OK;
  }
  FieldView res;
  TRY(jacobi_steps_view(ref_view(src, N), ref_view_mut(buf0, N), ref_view_mut(buf1, N), bc_fields(bc, bc_kind, N), f,
                        n, B, N, nullptr, &res, st));
  *result = res.p;
  return NPDE_OK;
}
}  // namespace npde

namespace {

int smooth(npde::FieldView &cur, MgLevel &lv, const float *bc, int bc_kind, const float *f, int n, int B,
           const npde::FieldViewMut *last, cudaStream_t st) {
  if (n <= 0) return NPDE_OK;
  // first destination must not be `cur`
  const npde::FieldViewMut &b0 = (cur.p == lv.p[0].p) ? lv.p[1] : lv.p[0];
  const npde::FieldViewMut &b1 = (b0.p == lv.p[0].p) ? lv.p[1] : lv.p[0];
  npde::BcFields bcf = npde::bc_fields(bc, bc_kind, lv.N);
  if (bc_kind == NPDE_BC_MASK) {  // the permuted copies made by vcycle() for this level
    bcf.values = npde::FieldView{lv.vperm, lv.p[0].pitch, lv.p[0].bstride, 1};
    bcf.mask = npde::FieldView{lv.mperm, lv.p[0].pitch, lv.p[0].bstride, 1};
  }
  npde::FieldView res;
  TRY(npde::jacobi_steps_view(cur, b0, b1, bcf, f, n, B, lv.N, last, &res, st));
  cur = res;
  return NPDE_OK;
}

// ---------------------------------------------------------------- UNet ---
struct UnetWs {
  int s[NPDE_MAX_LEVELS];        // interior size per level
  float *mask[NPDE_MAX_LEVELS];  // level masks (B, s+2, s+2); level 0 aliases bc
  size_t mbs[NPDE_MAX_LEVELS];
  float *skip[NPDE_MAX_LEVELS];
  float *t0, *t1, *y_int;
  void carve(Bump &b, int B, int N, int L, bool is_mask) {
    int n = N;
    for (int l = 0; l < L; ++l) {
      s[l] = n - 2;
      skip[l] = b.take<float>((size_t)B * s[l] * s[l]);
      mask[l] = (is_mask && l > 0) ? b.take<float>((size_t)B * n * n) : nullptr;
      mbs[l] = (size_t)n * n;
      n = (n - 1) / 2 + 1;
    }
    size_t n0 = (size_t)B * (N - 2) * (N - 2);
    t0 = b.take<float>(n0);
    t1 = b.take<float>(n0);
    y_int = b.take<float>(n0);
  }
};

// Tape + gradient buffers of npde_unet_step_bwd.
struct UnetBwdWs {
  UnetWs base;                                        // level sizes, masks, y_int (skip / t0 / t1 unused)
  float *F[NPDE_MAX_LEVELS][NPDE_MAX_UNET_SMOOTH + 1];   // F[l][j]: input of pre-smoothing conv j; F[l][pre] = skip_l
  float *Sx[NPDE_MAX_LEVELS][NPDE_MAX_UNET_SMOOTH + 1];  // Sx[l][j]: input of post-smoothing conv j
  float *A[NPDE_MAX_LEVELS];                          // Sx[l][post] + skip_l
  float *gskip[NPDE_MAX_LEVELS];
  float *gz, *g[2];
  void *wg;
  size_t wg_bytes;
  unsigned *counter;  // block counter of the fused backward-layer kernel
  void carve(Bump &b, int B, int N, int L, int pre, int post, bool is_mask) {
    base.carve(b, B, N, L, is_mask);
    for (int l = 0; l < L; ++l) {
      size_t n = (size_t)B * base.s[l] * base.s[l];
      for (int j = 0; j <= pre; ++j) F[l][j] = b.take<float>(n);
      for (int j = 0; j <= post; ++j) Sx[l][j] = (j == 0 && l == L - 1) ? F[l][pre] : b.take<float>(n);
      A[l] = b.take<float>(n);
      gskip[l] = b.take<float>(n);
    }
    size_t n0 = (size_t)B * (N - 2) * (N - 2);
    gz = b.take<float>(n0);
    g[0] = b.take<float>(n0);
    g[1] = b.take<float>(n0);
    size_t wa = npde::weight_grad_ws_bytes(B, N - 2), wb = npde::bwd_layer_ws_bytes(B, N - 2);
    wg_bytes = wa > wb ? wa : wb;
    wg = b.take_bytes(wg_bytes);
    counter = b.take<unsigned>(1);
  }
};

int unet_build_masks(UnetWs &u, const float *bc, int bc_kind, int B, int N, int L, cudaStream_t st) {
  if (bc_kind != NPDE_BC_MASK || !bc) {
    for (int l = 0; l < L; ++l) u.mask[l] = nullptr;
    return NPDE_OK;
  }
  // level 0 mask is the second plane of bc (batch stride 2*N*N)
  u.mask[0] = const_cast<float *>(bc) + (size_t)N * N;
  u.mbs[0] = 2 * (size_t)N * N;
  int n = N;
  for (int l = 1; l < L; ++l) {  // unet_iterator.py:44-48
    TRY(npde::subsample_strided(u.mask[l - 1], u.mbs[l - 1], u.mask[l], B, n, 1.0f, st));
    n = (n - 1) / 2 + 1;
  }
  return NPDE_OK;
}

// unet_iterator.py:52-85 on a tensor already multiplied by masks[0]; src (B,s0,s0) -> out
// unet_iterator.py:52-85 from level `l0` down and back up to level l0, on a tensor that is already multiplied
// by masks[l0]; src (B,s_l0,s_l0).  *result = the level-l0 tensor after its post-smoothing convs and skip add
// (what the reference upsamples next, or h when l0 == 0); it lives in u.t0 / u.t1.
// skip0: where the level-l0 skip tensor is kept (u.skip[l0] unless the caller already has it).
int unet_levels(UnetWs &u, int l0, const float *src, const float *wf, const float *wp, const float *wsnd, int L,
                int pre, int post, const float **result, int B, cudaStream_t st) {
  const float *cur = src;
  auto other = [&](const float *c) -> float * { return (c == u.t0) ? u.t1 : u.t0; };
  int up_from = L - 1;  // the level the way up starts at
  for (int l = l0; l < L; ++l) {
    int s = u.s[l];
    for (int j = 0; j < pre; ++j) {
      // the last pre-smoothing conv writes the skip tensor in place (:61), no copy
      float *dst = (j == pre - 1) ? u.skip[l] : other(cur);
      TRY(npde::conv3x3(cur, s, wf + 9 * (l * pre + j), 1, 1, u.mask[l], u.mbs[l], dst, s, B, st));
      cur = dst;
    }
    if (pre == 0) TRY(d2d(u.skip[l], cur, sizeof(float) * (size_t)B * s * s, st));
    if (l < L - 1 && npde::unet_coarse_supported(u.s, L, l + 1, pre, post, B) && !getenv("NPDE_UNET_LAYERED")) {
      // every deeper level fits in shared memory: pool, levels l+1 .. L-1 and the upsampling back to this level in
      // ONE kernel (npde_unet_coarse.cu); cur = the upsampled correction of level l, as after the loop below
      float *dst = other(cur);
      TRY(npde::unet_coarse(ArrView{u.skip[l], s, (long)s * s, 0, 0}, s, ArrViewMut{dst, s, (long)s * s, 0, 0}, wf, wp,
                            wsnd, u.mask, u.mbs, u.s, L, pre, post, l + 1, B, st));
      cur = dst;
      up_from = l;
      break;
    }
    if (l < L - 1) {
      float *dst = other(cur);
      TRY(npde::conv3x3(cur, s, wp + 9 * l, 2, 0, u.mask[l + 1], u.mbs[l + 1], dst, u.s[l + 1], B, st));  // :63-66
      cur = dst;
    }
  }
  for (int i = L - 1 - up_from; i < L - l0; ++i) {
    int l = L - 1 - i, s = u.s[l];
    for (int j = 0; j < post; ++j) {
      // the last post-smoothing conv adds the skip connection in its epilogue (:76)
      float *dst = other(cur);
      TRY(npde::conv3x3(cur, s, wsnd + 9 * (i * post + j), 1, 1, u.mask[l], u.mbs[l], dst, s, B, st,
                        j == post - 1 ? u.skip[l] : nullptr));
      cur = dst;
    }
    if (post == 0) {
      float *acc = other(cur);  // never the caller's input, never the skip tensor
      TRY(d2d(acc, cur, sizeof(float) * (size_t)B * s * s, st));
      TRY(npde::add_inplace(acc, u.skip[l], (size_t)B * s * s, st));  // :76
      cur = acc;
    }
    if (l > l0) {
      float *dst = other(cur);
      TRY(npde::pad_up_crop(cur, s, u.mask[l - 1], u.mbs[l - 1], dst, B, st));  // :78-83
      cur = dst;
    }
  }
  *result = cur;
  return NPDE_OK;
}

int unet_core(UnetWs &u, const float *src, const float *wf, const float *wp, const float *wsnd, int L, int pre,
              int post, float *out, int B, cudaStream_t st) {
  const float *res = nullptr;
  TRY(unet_levels(u, 0, src, wf, wp, wsnd, L, pre, post, &res, B, st));
  return d2d(out, res, sizeof(float) * (size_t)B * u.s[0] * u.s[0], st);
}

// Frames of the fused finest level (npde_unet_step_fwd with host weights): skip_0, y and the upsampled
// coarse correction, all in the pair-permuted padded layout.
struct UnetFusedWs {
  float *skip0, *y0, *up0;
  int pitch;
  long bstride;
  void carve(Bump &b, int B, int N) {
    pitch = (N + 3) & ~3;
    bstride = (long)npde_align_up((size_t)pitch * N, 64);
    skip0 = b.take<float>((size_t)B * bstride);
    y0 = b.take<float>((size_t)B * bstride);
    up0 = b.take<float>((size_t)B * bstride);
  }
};

int level_sizes_ok(int N, int L) {
  int n = N;
  for (int l = 0; l < L - 1; ++l) {
    if (!(n & 1) || n < 5) return 0;
    n = (n - 1) / 2 + 1;
  }
  return n >= 3;
}

// npde_iterate: two ping-pong fields (padded pitch: 16-byte aligned rows), the
// reduction scratch and the largest per-step workspace any iterator may need.
constexpr int ITER_SEG = 32;  // fused launches per captured segment of the Jacobi / Conv loop (even)
struct IterWs {
  float *buf[2];
  float *stage_l2, *stage_fd;  // (ITER_SEG, B) error rows of one replayed segment
  float *bc_copy;              // stable copy of the caller's bc (graph replay), or nullptr
  size_t bc_floats;
  float *fpad;  // the source in the padded layout (Jacobi / Conv fast path)
  float *vpad, *mpad;  // mask geometry: the two bc planes in the padded layout (Conv fast path)
  int pitch;
  long bstride;
  void *red;
  size_t red_bytes;
  void *step;
  size_t step_bytes;
  void carve(Bump &b, int iterator, int B, int N, int bc_kind, int L, int pre, int post) {
    pitch = (N + 3) & ~3;
    bstride = (long)npde_align_up((size_t)pitch * N, 64);
    size_t cells = (size_t)B * bstride;  // >= B*N*N, so the generic loop can use them unpadded
    if (iterator == NPDE_IT_CONV && bc_kind == NPDE_BC_SQUARE) {  // the wide kernel's batch-interleaved layout
      const size_t wide = npde::wide_field_floats(B, N);
      if (wide > cells) cells = wide;
    }
    buf[0] = b.take<float>(cells);
    buf[1] = b.take<float>(cells);
    fpad = (iterator == NPDE_IT_JACOBI || iterator == NPDE_IT_CONV) ? b.take<float>(cells) : nullptr;
    const bool mask_fast = (iterator == NPDE_IT_CONV || iterator == NPDE_IT_JACOBI) && bc_kind == NPDE_BC_MASK;
    vpad = mask_fast ? b.take<float>(cells) : nullptr;
    mpad = mask_fast ? b.take<float>(cells) : nullptr;
    red_bytes = npde::reduce_ws_bytes(B, N);
    red = b.take_bytes(red_bytes);
    stage_l2 = b.take<float>((size_t)ITER_SEG * B);
    stage_fd = b.take<float>((size_t)ITER_SEG * B);
    // graph-eligible sizes: the caller's bc is copied here so that cached segments never refer to caller memory
    bc_floats = ((size_t)B * N * N <= ((size_t)8 << 20)) ? (bc_kind == NPDE_BC_MASK ? 2 * (size_t)B * N * N : 4 * (size_t)B) : 0;
    bc_copy = bc_floats ? b.take<float>(bc_floats) : nullptr;
    step_bytes = 0;
    if (iterator == NPDE_IT_CONV) step_bytes = npde_workspace_bytes(NPDE_WS_CONV_FWD, B, N, bc_kind, L, 0, 0, 0);
    if (iterator == NPDE_IT_MULTIGRID) step_bytes = npde_workspace_bytes(NPDE_WS_MG, B, N, bc_kind, L, pre, post, 0);
    if (iterator == NPDE_IT_UNET) step_bytes = npde_workspace_bytes(NPDE_WS_UNET_FWD, B, N, bc_kind, L, pre, post, 0);
    step = b.take_bytes(step_bytes);
  }
};

// ------------------------------------------------------------ CUDA graphs ---
// Small problems are launch bound (a 32 x 65^2 Conv3 step is one 3 us kernel; a U-Net step at 1 x 1025^2 is ~25
// kernels of a few us each), so the k-step loop replays a short captured segment instead of issuing every launch
// from the host: the segment (an even number of steps, ping-pong buffers back where they started) is captured
// once on a private stream, instantiated, and launched into the caller's stream as often as it fits.  The
// caller's stream may be the legacy default stream, which cannot capture -- hence the private stream.
// Instantiated segments are cached per (shape, iterator, every pointer and every by-value weight the captured
// launches contain): the drivers call npde_iterate over and over with the same workspace (calculate_errors in chunks,
// solve_to_tolerance per check interval, runtime.py once per problem), and capture + instantiation cost as much as a
// few dozen small launches.  The caller's bc tensor is copied into the workspace first (see npde_iterate) so that
// the segment only refers to memory that stays the same from call to call.
struct GraphKey {
  int dev, iterator, B, N, bc_kind, n_layers, pre, post, act, adv, want_l2, want_fd, n_whost;
  const void *bc, *f, *w, *gt, *ws;
  unsigned long long whost_hash;
};
#ifndef NPDE_EMU
struct GraphCache {
  struct Entry {
    GraphKey key;
    cudaGraph_t graph;
    cudaGraphExec_t exec;
    unsigned long long stamp;
  };
  std::mutex mu;
  std::vector<Entry> entries;
  unsigned long long clock = 0;
  static constexpr size_t kMax = 16;
  cudaGraphExec_t find(const GraphKey &k) {
    std::lock_guard<std::mutex> lock(mu);
    for (auto &e : entries)
      if (memcmp(&e.key, &k, sizeof(GraphKey)) == 0) {
        e.stamp = ++clock;
        return e.exec;
      }
    return nullptr;
  }
  void put(const GraphKey &k, cudaGraph_t g, cudaGraphExec_t ex) {
    std::lock_guard<std::mutex> lock(mu);
    if (entries.size() >= kMax) {  // evict the least recently used (an in-flight launch completes first)
      size_t lru = 0;
      for (size_t i = 1; i < entries.size(); ++i)
        if (entries[i].stamp < entries[lru].stamp) lru = i;
      cudaGraphExecDestroy(entries[lru].exec);
      cudaGraphDestroy(entries[lru].graph);
      entries.erase(entries.begin() + lru);
    }
    entries.push_back(Entry{k, g, ex, ++clock});
  }
};
GraphCache g_graphs;
#endif

inline GraphKey make_graph_key(const npde_iterate_desc *d, int adv, const void *ws, size_t n_whost) {
  GraphKey k;
  memset(&k, 0, sizeof(k));  // padding bytes take part in the comparison
  int dev = 0;
#ifndef NPDE_EMU
  cudaGetDevice(&dev);
#endif
  k.dev = dev; k.iterator = d->iterator; k.B = d->B; k.N = d->N; k.bc_kind = d->bc_kind; k.n_layers = d->n_layers;
  k.pre = d->pre; k.post = d->post; k.act = d->act; k.adv = adv;
  k.want_l2 = d->err_l2 != nullptr; k.want_fd = d->err_fd != nullptr;
  k.bc = d->bc; k.f = d->f; k.w = d->w; k.gt = d->err_l2 ? d->gt : nullptr; k.ws = ws;
  k.n_whost = d->w_host ? (int)n_whost : -1;
  unsigned long long h = 1469598103934665603ull;  // FNV-1a over the by-value weights
  if (d->w_host) {
    const unsigned char *p = reinterpret_cast<const unsigned char *>(d->w_host);
    for (size_t i = 0; i < n_whost * sizeof(float); ++i) h = (h ^ p[i]) * 1099511628211ull;
  }
  k.whost_hash = h;
  return k;
}

struct StepGraph {
#ifndef NPDE_EMU
  cudaStream_t cs = nullptr;
  cudaGraph_t graph = nullptr;
  cudaGraphExec_t exec = nullptr;
  bool capturing = false, cached = false, owned = true;
  GraphKey key;
  // true: a cached instantiation of this very segment exists -- nothing to capture
  bool lookup(const GraphKey &k) {
    key = k;
    exec = g_graphs.find(k);
    cached = exec != nullptr;
    if (cached) owned = false;
    return cached;
  }
  // false: graphs are not usable here (caller is capturing itself, or stream creation failed) -> plain launches
  bool begin(cudaStream_t caller) {
    cudaStreamCaptureStatus status = cudaStreamCaptureStatusNone;
    if (cudaStreamIsCapturing(caller, &status) != cudaSuccess || status != cudaStreamCaptureStatusNone) {
      cudaGetLastError();
      return false;
    }
    if (cudaStreamCreateWithFlags(&cs, cudaStreamNonBlocking) != cudaSuccess) { cudaGetLastError(); return false; }
    if (cudaStreamBeginCapture(cs, cudaStreamCaptureModeRelaxed) != cudaSuccess) { cudaGetLastError(); return false; }
    capturing = true;
    return true;
  }
  int abort_capture() {
    if (capturing) { cudaStreamEndCapture(cs, &graph); capturing = false; cudaGetLastError(); }
    return NPDE_OK;
  }
  // ends the capture, instantiates and hands the result to the cache; on failure the caller falls back to plain launches
  int finish() {
    if (cached) return NPDE_OK;
    cudaError_t e = cudaStreamEndCapture(cs, &graph);
    capturing = false;
    if (e == cudaSuccess) e = cudaGraphInstantiate(&exec, graph, 0);
    if (e != cudaSuccess) { cudaGetLastError(); return (int)e; }
    g_graphs.put(key, graph, exec);
    owned = false;
    return NPDE_OK;
  }
  int launch_once(cudaStream_t caller) {
    cudaError_t e = cudaGraphLaunch(exec, caller);
    return e == cudaSuccess ? NPDE_OK : (int)e;
  }
  ~StepGraph() {
    abort_capture();
    if (owned) {
      if (exec) cudaGraphExecDestroy(exec);
      if (graph) cudaGraphDestroy(graph);
    }
    if (cs) cudaStreamDestroy(cs);
  }
#endif
};

// replaying pays once the host cannot keep the device busy: few cells per launch and enough steps to amortise
// the capture + instantiation
inline bool graph_worthwhile(size_t cells, int k) {
#ifdef NPDE_EMU
  (void)cells; (void)k;
  return false;
#else
  return cells <= ((size_t)8 << 20) && k >= 16;
#endif
}

// the wide kernel wants enough strip-rows for every resident warp to get a piece of useful length
inline bool wide_worthwhile(int B, int N) {
#ifdef NPDE_EMU
  (void)B;
  return N >= 9;
#else
  const char *v = getenv("NPDE_WIDE");
  if (v) return atoi(v) != 0;
  const long groups = (long)B * ((N + 7) / 8);
  return (groups + 29) / 30 * (long)N >= 4096;
#endif
}

// Backward of one Conv step given its input x (recomputes the layer tape), SURVEY.md Appendix B.
// gw is overwritten, or accumulated into in fp32 when `accumulate`.
int conv_step_bwd_impl(ConvBwdWs &c, const float *x, const float *bc, int bc_kind, const float *f, const float *w,
                       int L, int act, const float *gout, float *gw, bool accumulate, float *gx, int B, int N,
                       cudaStream_t st) {
  // square domain, up to three layers: recompute + adjoint + weight gradient in ONE streaming pass
  if (npde::conv_bwd_stream_supported(bc_kind, L) && !getenv("NPDE_BWD_LAYERED"))
    return npde::conv_bwd_stream(x, f, w, L, act, gout, gw, accumulate, gx, B, N, c.wg, c.wg_bytes, st);
  int Sz = N - 2;
  size_t bytes = sizeof(float) * (size_t)B * Sz * Sz;
  const float *mask = mask0(bc, bc_kind, N);
  size_t mbs = 2 * (size_t)N * N;
  // forward with tape
  TRY(npde::psi_residual(x, f, mask, mbs, c.y_int, c.rs[0], B, N, st));
  for (int l = 0; l < L; ++l) TRY(npde::conv3x3(c.rs[l], Sz, w + 9 * l, 1, 1, mask, mbs, c.rs[l + 1], Sz, B, st));
  // backward
  TRY(npde::bwd_gz(gout, c.y_int, c.rs[L], mask, mbs, act, c.gz, B, N, st));
  {
    cudaError_t e = cudaMemsetAsync(c.counters, 0, sizeof(unsigned) * NPDE_MAX_CONV_LAYERS, st);
    if (e != cudaSuccess) return (int)e;
  }
  // layer by layer: one fused pass per layer gives both the weight gradient and the gradient of the layer input.
  // gz itself must survive for the last kernel, so on mask geometries (in-place fg multiply) work on a copy.
  const float *g = c.gz;
  float *outs[2] = {c.g, c.t};
  int o = 0;
  if (mask) {
    TRY(d2d(c.g, c.gz, bytes, st));
    g = c.g;
    o = 1;
  }
  for (int l = L - 1; l >= 0; --l) {
    if (mask) TRY(npde::mul_fg(const_cast<float *>(g), mask, mbs, Sz, B, st));
    TRY(npde::bwd_layer(g, c.rs[l], w + 9 * l, outs[o], Sz, gw + 9 * l, B, c.wg, c.wg_bytes, c.counters + l, accumulate,
                        st));
    g = outs[o];
    o ^= 1;
  }
  if (gx) {
    if (mask) TRY(npde::mul_fg(const_cast<float *>(g), mask, mbs, Sz, B, st));
    TRY(npde::bwd_gx(c.gz, g, gx, B, N, st));
  }
  return NPDE_OK;
}


// ------------------------------------------------------------------ npde_solve ---
// One warp: largest residual of the batch, bookkeeping, and (on the device) the WHILE condition of the loop.
struct SolveResult {
  int iterations, converged;
  float residual, residual0;
};
#ifdef NPDE_EMU
__global__ void k_solve_check(const float *err, int B, float tol, int add, int max_iters, SolveResult *r, int first) {
#else
__global__ void k_solve_check(const float *err, int B, float tol, int add, int max_iters, SolveResult *r, int first,
                              cudaGraphConditionalHandle handle) {
#endif
  float m = 0.0f;
  bool nan = false;
  for (int b = threadIdx.x; b < B; b += 32) {
    const float e = err[b];
    nan = nan || (e != e);
    m = fmaxf(m, e);
  }
  m = npde_warp_max(m);
  if (npde_warp_max(nan ? 1.0f : 0.0f) > 0.0f) m = __int_as_float(0x7fc00000);  // a NaN residual never converges
  if (threadIdx.x == 0) {
    if (first) {
      r->iterations = 0;
      r->residual0 = m;
    } else {
      r->iterations += add;
    }
    r->residual = m;
    r->converged = (m < tol) ? 1 : 0;
#ifndef NPDE_EMU
    cudaGraphSetConditional(handle, (!(m < tol) && r->iterations < max_iters) ? 1u : 0u);
#endif
  }
}

}  // namespace

// ======================================================================= C ABI
extern "C" {

int npde_version(void) { return 200; }  // major * 100 + minor: 2.0 = round 2 (packed solver loop, npde_solve, streaming backward)

unsigned long long npde_debug_launch_count(void) { return g_launches.load(std::memory_order_relaxed); }

const char *npde_strerror(int code) {
  switch (code) {
    case NPDE_OK: return "ok";
    case NPDE_ERR_NULL: return "required pointer is NULL";
    case NPDE_ERR_SIZE: return "bad batch / grid size";
    case NPDE_ERR_BC: return "bad boundary-condition encoding";
    case NPDE_ERR_ACT: return "unknown activation";
    case NPDE_ERR_LAYERS: return "layer / smoothing count out of range";
    case NPDE_ERR_WORKSPACE: return "workspace missing, misaligned or too small";
    case NPDE_ERR_ARG: return "invalid argument";
    default: return code > 0 ? cudaGetErrorString((cudaError_t)code) : "unknown error";
  }
}

int npde_set_boundary(float *x, const float *bc, int bc_kind, int B, int N, void *stream) {
  TRY(npde::check_common(x, x, B, N));
  TRY(npde::check_bc(bc, bc_kind));
  return npde::set_boundary(x, bc, bc_kind, B, N, S(stream));
}

int npde_initialize(float *x, const float *bc, int bc_kind, int mode, const float *rnd, int B, int N, void *stream) {
  TRY(npde::check_common(x, x, B, N));
  TRY(npde::check_bc(bc, bc_kind));
  if (mode < NPDE_INIT_ZERO || mode > NPDE_INIT_AVG) return NPDE_ERR_ARG;
  if (mode == NPDE_INIT_RANDOM && !rnd) return NPDE_ERR_NULL;
  if (bc_kind == NPDE_BC_MASK && mode == NPDE_INIT_AVG) return NPDE_OK;  // heat_utils.py:77-83: no branch for 'avg'
  return npde::initialize(x, bc, bc_kind, mode, rnd, B, N, S(stream));
}

int npde_pad_boundary(const float *xin, const float *bc, int bc_kind, float *y, int B, int N, void *stream) {
  TRY(npde::check_common(xin, y, B, N));
  TRY(npde::check_bc(bc, bc_kind));
  return npde::pad_boundary(xin, bc, bc_kind, y, B, N, S(stream));
}

int npde_fd_step(const float *x, const float *bc, int bc_kind, const float *f, float *y, int B, int N,
                 void *stream) {
  TRY(npde::check_common(x, y, B, N));
  TRY(npde::check_bc(bc, bc_kind));
  if (x == y) return NPDE_ERR_ARG;
  const float *res = nullptr;
  return npde::jacobi_steps(x, y, nullptr, bc, bc_kind, f, 1, B, N, &res, S(stream));
}

int npde_fd_error(const float *x, const float *bc, int bc_kind, const float *f, int aggregate, float *err,
                  int B, int N, void *ws, size_t ws_bytes, void *stream) {
  TRY(npde::check_common(x, err, B, N));
  TRY(npde::check_bc(bc, bc_kind));
  if (aggregate != NPDE_AGG_MAX && aggregate != NPDE_AGG_MEAN) return NPDE_ERR_ARG;
  return npde::fd_error(x, bc, bc_kind, f, aggregate, err, B, N, ws, ws_bytes, S(stream));
}

int npde_cg_step(const float *x, int n_iters, float *y, int B, int N, void *ws, size_t ws_bytes, void *stream) {
  TRY(npde::check_common(x, y, B, N));
  if (n_iters < 0) return NPDE_ERR_ARG;
  if (x == y) return NPDE_ERR_ARG;  // the first kernel reads neighbours of x while writing y
  TRY(ws_ok(ws, ws_bytes, npde::cg_ws_bytes(B, N)));
  return npde::cg_step(x, n_iters, y, B, N, ws, S(stream));
}

int npde_l2_error(const float *x, const float *gt, float *err, int B, int N, void *ws, size_t ws_bytes,
                  void *stream) {
  TRY(npde::check_common(x, gt, B, N));
  if (!err) return NPDE_ERR_NULL;
  return npde::l2_error(x, gt, err, B, N, ws, ws_bytes, S(stream));
}

int npde_conv_H(const float *r, const float *bc, int bc_kind, const float *w, int n_layers, float *out, int B,
                int N, void *ws, size_t ws_bytes, void *stream) {
  TRY(npde::check_common(r, out, B, N));
  if (bc_kind != NPDE_BC_SQUARE && bc_kind != NPDE_BC_MASK) return NPDE_ERR_BC;
  if (bc_kind == NPDE_BC_MASK && !bc) return NPDE_ERR_NULL;
  if (n_layers < 0 || n_layers > NPDE_MAX_CONV_LAYERS) return NPDE_ERR_LAYERS;
  if (n_layers > 0 && !w) return NPDE_ERR_NULL;
  TRY(ws_ok(ws, ws_bytes, npde_workspace_bytes(NPDE_WS_CONV_FWD, B, N, bc_kind, n_layers, 0, 0, 0)));
  return conv_H_impl(r, bc, bc_kind, w, n_layers, out, B, N, ws, S(stream));
}

int npde_conv_step_fwd(const float *x, const float *bc, int bc_kind, const float *f, const float *w,
                       const float *w_host, int n_layers, int act, float *y, int B, int N, void *ws,
                       size_t ws_bytes, void *stream) {
  TRY(npde::check_common(x, y, B, N));
  TRY(npde::check_bc(bc, bc_kind));
  if (act < NPDE_ACT_NONE || act > NPDE_ACT_SIGMOID) return NPDE_ERR_ACT;
  if (n_layers < 0 || n_layers > NPDE_MAX_CONV_LAYERS) return NPDE_ERR_LAYERS;
  if (n_layers > 0 && !w) return NPDE_ERR_NULL;
  if (x == y) return NPDE_ERR_ARG;
  if (npde::conv_stream_supported(bc_kind, n_layers, w_host))
    return npde::conv_stream(npde::ref_view(x, N), npde::ref_view_mut(y, N), npde::bc_fields(bc, bc_kind, N),
                             npde::ref_view(f, N), w_host, n_layers, act, B, N, S(stream));
  TRY(ws_ok(ws, ws_bytes, npde_workspace_bytes(NPDE_WS_CONV_FWD, B, N, bc_kind, n_layers, 0, 0, 0)));
  return conv_step_layered(x, bc, bc_kind, f, w, n_layers, act, y, B, N, ws, S(stream));
}

int npde_conv_step_bwd(const float *x, const float *bc, int bc_kind, const float *f, const float *w,
                       int n_layers, int act, const float *gout, float *gw, float *gx, int B, int N, void *ws,
                       size_t ws_bytes, void *stream) {
  TRY(npde::check_common(x, gout, B, N));
  TRY(npde::check_bc(bc, bc_kind));
  if (act < NPDE_ACT_NONE || act > NPDE_ACT_SIGMOID) return NPDE_ERR_ACT;
  if (n_layers < 1 || n_layers > NPDE_MAX_CONV_LAYERS) return NPDE_ERR_LAYERS;
  if (!w || !gw) return NPDE_ERR_NULL;
  TRY(ws_ok(ws, ws_bytes, npde_workspace_bytes(NPDE_WS_CONV_BWD, B, N, bc_kind, n_layers, 0, 0, 0)));
  Bump b(ws);
  ConvBwdWs c;
  c.carve(b, B, N, n_layers);
  return conv_step_bwd_impl(c, x, bc, bc_kind, f, w, n_layers, act, gout, gw, false, gx, B, N, S(stream));
}

int npde_conv_iterate_tape(const float *x, const float *bc, int bc_kind, const float *f, const float *w,
                           const float *w_host, int n_layers, int act, int k, float *tape, float *y, int B, int N,
                           void *ws, size_t ws_bytes, void *stream) {
  TRY(npde::check_common(x, y, B, N));
  TRY(npde::check_bc(bc, bc_kind));
  if (!tape) return NPDE_ERR_NULL;
  if (k < 1) return NPDE_ERR_ARG;
  const size_t n = (size_t)B * N * N;
  TRY(d2d(tape, x, sizeof(float) * n, S(stream)));
  for (int t = 0; t < k; ++t) {
    float *dst = (t + 1 < k) ? tape + (size_t)(t + 1) * n : y;
    TRY(npde_conv_step_fwd(tape + (size_t)t * n, bc, bc_kind, f, w, w_host, n_layers, act, dst, B, N, ws, ws_bytes,
                           stream));
  }
  return NPDE_OK;
}

int npde_conv_iterate_bwd(const float *tape, const float *bc, int bc_kind, const float *f, const float *w,
                          int n_layers, int act, int k, const float *gout, float *gw, float *gx, int B, int N,
                          void *ws, size_t ws_bytes, void *stream) {
  TRY(npde::check_common(tape, gout, B, N));
  TRY(npde::check_bc(bc, bc_kind));
  if (act < NPDE_ACT_NONE || act > NPDE_ACT_SIGMOID) return NPDE_ERR_ACT;
  if (n_layers < 1 || n_layers > NPDE_MAX_CONV_LAYERS) return NPDE_ERR_LAYERS;
  if (!w || !gw) return NPDE_ERR_NULL;
  if (k < 1) return NPDE_ERR_ARG;
  TRY(ws_ok(ws, ws_bytes, npde_workspace_bytes(NPDE_WS_CONV_BPTT, B, N, bc_kind, n_layers, 0, 0, 0)));
  Bump b(ws);
  ConvBwdWs c;
  c.carve(b, B, N, n_layers);
  const size_t n = (size_t)B * N * N;
  float *g[2] = {b.take<float>(n), b.take<float>(n)};
  const float *gin = gout;
  // last step first: the order in which autograd adds the per-step weight gradients (fp32)
  for (int t = k - 1; t >= 0; --t) {
    float *gnext = (t > 0) ? g[t & 1] : gx;
    TRY(conv_step_bwd_impl(c, tape + (size_t)t * n, bc, bc_kind, f, w, n_layers, act, gin, gw, t != k - 1, gnext, B, N,
                           S(stream)));
    gin = gnext;
  }
  return NPDE_OK;
}

int npde_subsample(const float *x, float *y, int B, int N, void *stream) {
  TRY(npde::check_common(x, y, B, N));
  if (!(N & 1)) return NPDE_ERR_SIZE;
  return npde::subsample(x, y, B, N, 1.0f, S(stream));
}

int npde_restrict(const float *x, const float *bc_sub, int bc_kind, float *y, int B, int N, void *stream) {
  TRY(npde::check_common(x, y, B, N));
  TRY(npde::check_bc(bc_sub, bc_kind));
  if (!(N & 1) || N < 5) return NPDE_ERR_SIZE;
  return npde::restrict_(x, bc_sub, bc_kind, y, B, N, S(stream));
}

int npde_prolong(const float *x, const float *bc, int bc_kind, float *y, int B, int M, void *stream) {
  TRY(npde::check_common(x, y, B, M));
  TRY(npde::check_bc(bc, bc_kind));
  return npde::prolong(x, bc, bc_kind, y, B, M, S(stream));
}

int npde_mg_step(const float *x, const float *bc, int bc_kind, const float *f, int n_layers, int pre, int post,
                 float *y, int B, int N, void *ws, size_t ws_bytes, void *stream) {
  TRY(npde::check_common(x, y, B, N));
  TRY(npde::check_bc(bc, bc_kind));
  if (n_layers < 1 || n_layers > NPDE_MAX_LEVELS || pre < 0 || post < 0) return NPDE_ERR_LAYERS;
  if (!level_sizes_ok(N, n_layers)) return NPDE_ERR_SIZE;
  if (x == y) return NPDE_ERR_ARG;
  TRY(ws_ok(ws, ws_bytes, npde_workspace_bytes(NPDE_WS_MG, B, N, bc_kind, n_layers, pre, post, 0)));
  cudaStream_t st = S(stream);
  Bump b(ws);
  MgWs m;
  m.carve(b, B, N, bc_kind, n_layers, f != nullptr);
  npde::FieldViewMut out = npde::ref_view_mut(y, N);
  npde::FieldView res;
  TRY(vcycle(m, 0, n_layers, npde::ref_view(x, N), bc, f, bc_kind, pre, post, B, &out, &res, st));
  if (res.p != y) TRY(npde::copy_field(res, out, B, N, st));  // pre == post == 0 on a single level
  return NPDE_OK;
}

int npde_unet_H(const float *r, const float *bc, int bc_kind, const float *w_first, const float *w_pool,
                const float *w_second, int n_layers, int pre, int post, float *out, int B, int N, void *ws,
                size_t ws_bytes, void *stream) {
  TRY(npde::check_common(r, out, B, N));
  if (bc_kind != NPDE_BC_SQUARE && bc_kind != NPDE_BC_MASK) return NPDE_ERR_BC;
  if (bc_kind == NPDE_BC_MASK && !bc) return NPDE_ERR_NULL;
  if (n_layers < 1 || n_layers > NPDE_MAX_LEVELS || pre < 0 || post < 0) return NPDE_ERR_LAYERS;
  if (!level_sizes_ok(N, n_layers)) return NPDE_ERR_SIZE;
  if ((pre > 0 && !w_first) || (post > 0 && !w_second) || (n_layers > 1 && !w_pool)) return NPDE_ERR_NULL;
  TRY(ws_ok(ws, ws_bytes, npde_workspace_bytes(NPDE_WS_UNET_FWD, B, N, bc_kind, n_layers, pre, post, 0)));
  cudaStream_t st = S(stream);
  Bump b(ws);
  UnetWs u;
  u.carve(b, B, N, n_layers, bc_kind == NPDE_BC_MASK);
  TRY(unet_build_masks(u, bc, bc_kind, B, N, n_layers, st));
  const float *src = r;
  if (u.mask[0]) {  // unet_iterator.py:50
    TRY(d2d(u.y_int, r, sizeof(float) * (size_t)B * u.s[0] * u.s[0], st));
    TRY(npde::mul_fg(u.y_int, u.mask[0], u.mbs[0], u.s[0], B, st));
    src = u.y_int;
  }
  return unet_core(u, src, w_first, w_pool, w_second, n_layers, pre, post, out, B, st);
}

int npde_unet_step_fwd(const float *x, const float *bc, int bc_kind, const float *f, const float *w_first,
                       const float *w_pool, const float *w_second, int n_layers, int pre, int post, int act,
                       float *y, int B, int N, void *ws, size_t ws_bytes, void *stream) {
  TRY(npde::check_common(x, y, B, N));
  TRY(npde::check_bc(bc, bc_kind));
  if (act < NPDE_ACT_NONE || act > NPDE_ACT_SIGMOID) return NPDE_ERR_ACT;
  if (n_layers < 1 || n_layers > NPDE_MAX_LEVELS || pre < 0 || post < 0) return NPDE_ERR_LAYERS;
  if (!level_sizes_ok(N, n_layers)) return NPDE_ERR_SIZE;
  if ((pre > 0 && !w_first) || (post > 0 && !w_second) || (n_layers > 1 && !w_pool)) return NPDE_ERR_NULL;
  if (x == y) return NPDE_ERR_ARG;
  TRY(ws_ok(ws, ws_bytes, npde_workspace_bytes(NPDE_WS_UNET_FWD, B, N, bc_kind, n_layers, pre, post, 0)));
  cudaStream_t st = S(stream);
  Bump b(ws);
  UnetWs u;
  u.carve(b, B, N, n_layers, bc_kind == NPDE_BC_MASK);
  TRY(unet_build_masks(u, bc, bc_kind, B, N, n_layers, st));
  // r0 lands in a dedicated buffer so unet_core can ping-pong t0/t1 freely
  float *r0 = u.skip[0];  // skip[0] is overwritten later, after r0 has been consumed? no: keep separate
  (void)r0;
  float *h = b.take<float>((size_t)B * u.s[0] * u.s[0]);
  float *r0buf = b.take<float>((size_t)B * u.s[0] * u.s[0]);
  TRY(npde::psi_residual(x, f, u.mask[0], u.mbs[0], u.y_int, r0buf, B, N, st));
  TRY(unet_core(u, r0buf, w_first, w_pool, w_second, n_layers, pre, post, h, B, st));
  return npde::combine(u.y_int, h, bc, bc_kind, act, y, B, N, st);
}

// The finest level of a U-Net step fused into two streaming kernels (head: Psi, residual and the
// pre-smoothing convs; tail: post-smoothing convs, + skip, + y, act, G); the coarser levels run layer by layer
// in between.  Needs host copies of the weights (they travel as kernel parameters).
static bool unet_fused_ok(int n_layers, int pre, int post, const float *w_host) {
  return w_host != nullptr && pre >= 1 && pre <= 4 && post >= 1 && post <= 4 && n_layers >= 1;
}

int npde_unet_step_fwd_hw(const float *x, const float *bc, int bc_kind, const float *f, const float *w_first,
                          const float *w_pool, const float *w_second, const float *w_host, int n_layers, int pre,
                          int post, int act, float *y, int B, int N, void *ws, size_t ws_bytes, void *stream) {
  if (!unet_fused_ok(n_layers, pre, post, w_host))
    return npde_unet_step_fwd(x, bc, bc_kind, f, w_first, w_pool, w_second, n_layers, pre, post, act, y, B, N, ws,
                              ws_bytes, stream);
  TRY(npde::check_common(x, y, B, N));
  TRY(npde::check_bc(bc, bc_kind));
  if (act < NPDE_ACT_NONE || act > NPDE_ACT_SIGMOID) return NPDE_ERR_ACT;
  if (n_layers > NPDE_MAX_LEVELS) return NPDE_ERR_LAYERS;
  if (!level_sizes_ok(N, n_layers)) return NPDE_ERR_SIZE;
  if (!w_first || !w_second || (n_layers > 1 && !w_pool)) return NPDE_ERR_NULL;
  if (x == y) return NPDE_ERR_ARG;
  TRY(ws_ok(ws, ws_bytes, npde_workspace_bytes(NPDE_WS_UNET_FWD, B, N, bc_kind, n_layers, pre, post, 0)));
  cudaStream_t st = S(stream);
  const int L = n_layers;
  Bump b(ws);
  UnetWs u;
  u.carve(b, B, N, L, bc_kind == NPDE_BC_MASK);
  b.take<float>((size_t)B * u.s[0] * u.s[0]);  // (layout of the layered path: h, r0)
  b.take<float>((size_t)B * u.s[0] * u.s[0]);
  UnetFusedWs fw;
  fw.carve(b, B, N);
  TRY(unet_build_masks(u, bc, bc_kind, B, N, L, st));
  const float *wh_first = w_host, *wh_second = w_host + 9 * ((size_t)L * pre + (L - 1));
  npde::BcFields bcf = npde::bc_fields(bc, bc_kind, N);
  npde::FieldViewMut skipv = {fw.skip0, fw.pitch, fw.bstride, 1};
  // head: x -> y (frame), skip_0 = pre-smoothed masked residual (frame)
  TRY(npde::conv_head_stream(npde::ref_view(x, N), bcf, npde::ref_view(f, N), wh_first, pre, skipv, fw.y0, B, N, st));
  npde::FieldView rin = npde::as_const(skipv);  // single level: the post convs start from skip_0 itself
  if (L > 1 && npde::unet_coarse_supported(u.s, L, 1, pre, post, B) && !getenv("NPDE_UNET_LAYERED")) {
    // levels 1 .. L-1 fit on chip (N <= 257): pooling conv of level 0, all coarser levels and the upsampling into the
    // level-0 frame in one kernel, one CTA per problem
    TRY(npde::unet_coarse(ArrView{fw.skip0, fw.pitch, fw.bstride, 1, 1}, u.s[0],
                          ArrViewMut{fw.up0, fw.pitch, fw.bstride, 1, 1}, w_first, w_pool, w_second, u.mask, u.mbs, u.s,
                          L, pre, post, 1, B, st));
    rin = npde::FieldView{fw.up0, fw.pitch, fw.bstride, 1};
  } else if (L > 1) {
    // pooling conv of level 0 reads the skip frame, then levels 1 .. L-1 layer by layer
    float *lvl1 = u.t0;
    TRY(npde::conv3x3_view(ArrView{fw.skip0, fw.pitch, fw.bstride, 1, 1}, u.s[0], w_pool, 2, 0, u.mask[1], u.mbs[1], lvl1,
                           u.s[1], B, st));
    const float *res = nullptr;
    TRY(unet_levels(u, 1, lvl1, w_first, w_pool, w_second, L, pre, post, &res, B, st));
    TRY(npde::pad_up_crop_view(res, u.s[1], u.mask[0], u.mbs[0], ArrViewMut{fw.up0, fw.pitch, fw.bstride, 1, 1}, B, st));
    rin = npde::FieldView{fw.up0, fw.pitch, fw.bstride, 1};
  }
  // tail: post-smoothing convs of level 0, + skip_0, + y, act, pad, G
  return npde::conv_tail_stream(rin, fw.skip0, fw.y0, bcf, wh_second + 9 * (size_t)(L - 1) * post, post, act,
                                npde::ref_view_mut(y, N), B, N, st);
}

int npde_unet_step_bwd(const float *x, const float *bc, int bc_kind, const float *f, const float *w_first,
                       const float *w_pool, const float *w_second, int n_layers, int pre, int post, int act,
                       const float *gout, float *gw_first, float *gw_pool, float *gw_second, float *gx, int B,
                       int N, void *ws, size_t ws_bytes, void *stream) {
  TRY(npde::check_common(x, gout, B, N));
  TRY(npde::check_bc(bc, bc_kind));
  if (act < NPDE_ACT_NONE || act > NPDE_ACT_SIGMOID) return NPDE_ERR_ACT;
  if (n_layers < 1 || n_layers > NPDE_MAX_LEVELS || pre < 0 || post < 0 || pre > NPDE_MAX_UNET_SMOOTH ||
      post > NPDE_MAX_UNET_SMOOTH)
    return NPDE_ERR_LAYERS;
  if (!level_sizes_ok(N, n_layers)) return NPDE_ERR_SIZE;
  if ((pre > 0 && (!w_first || !gw_first)) || (post > 0 && (!w_second || !gw_second)) ||
      (n_layers > 1 && (!w_pool || !gw_pool)))
    return NPDE_ERR_NULL;
  TRY(ws_ok(ws, ws_bytes, npde_workspace_bytes(NPDE_WS_UNET_BWD, B, N, bc_kind, n_layers, pre, post, 0)));
  cudaStream_t st = S(stream);
  const int L = n_layers;
  Bump b(ws);
  UnetBwdWs u;
  u.carve(b, B, N, L, pre, post, bc_kind == NPDE_BC_MASK);
  TRY(unet_build_masks(u.base, bc, bc_kind, B, N, L, st));
  float *const *mask = u.base.mask;
  const size_t *mbs = u.base.mbs;
  const int *sz = u.base.s;
  auto cells = [&](int l) { return (size_t)B * sz[l] * sz[l]; };

  // ---- forward with tape (unet_iterator.py:36-85): every layer input is kept -------------------
  TRY(npde::psi_residual(x, f, mask[0], mbs[0], u.base.y_int, u.F[0][0], B, N, st));
  for (int l = 0; l < L; ++l) {
    for (int j = 0; j < pre; ++j)
      TRY(npde::conv3x3(u.F[l][j], sz[l], w_first + 9 * (l * pre + j), 1, 1, mask[l], mbs[l], u.F[l][j + 1], sz[l], B,
                        st));
    if (l < L - 1)
      TRY(npde::conv3x3(u.F[l][pre], sz[l], w_pool + 9 * l, 2, 0, mask[l + 1], mbs[l + 1], u.F[l + 1][0], sz[l + 1], B,
                        st));
  }
  for (int i = 0; i < L; ++i) {
    const int l = L - 1 - i;
    for (int j = 0; j < post; ++j)
      TRY(npde::conv3x3(u.Sx[l][j], sz[l], w_second + 9 * (i * post + j), 1, 1, mask[l], mbs[l], u.Sx[l][j + 1], sz[l],
                        B, st));
    TRY(d2d(u.A[l], u.Sx[l][post], sizeof(float) * cells(l), st));
    TRY(npde::add_inplace(u.A[l], u.F[l][pre], cells(l), st));  // :76 skip connection
    if (i < L - 1) TRY(npde::pad_up_crop(u.A[l], sz[l], mask[l - 1], mbs[l - 1], u.Sx[l - 1][0], B, st));
  }
  const float *h = u.A[0];

  // ---- backward ---------------------------------------------------------------------------------
  TRY(npde::bwd_gz(gout, u.base.y_int, h, mask[0], mbs[0], act, u.gz, B, N, st));
  float *g = u.g[0], *other = u.g[1];
  auto swap = [&]() { float *t = g; g = other; other = t; };
  // backward of one same-size conv layer: weight gradient and input gradient in one fused pass over g
  auto layer_bwd = [&](const float *gin, const float *r, const float *w9, float *gout_, int S, float *gw9) -> int {
    cudaError_t e = cudaMemsetAsync(u.counter, 0, sizeof(unsigned), st);
    if (e != cudaSuccess) return (int)e;
    return npde::bwd_layer(gin, r, w9, gout_, S, gw9, B, u.wg, u.wg_bytes, u.counter, false, st);
  };
  TRY(d2d(g, u.gz, sizeof(float) * cells(0), st));
  for (int i = L - 1; i >= 0; --i) {
    const int l = L - 1 - i;
    if (i < L - 1) {  // undo "* fg, crop, upsample, pad" that followed this level (:78-83)
      if (mask[l - 1]) TRY(npde::mul_fg(g, mask[l - 1], mbs[l - 1], sz[l - 1], B, st));
      TRY(npde::pad_up_crop_transpose(g, sz[l], other, B, st));
      swap();
    }
    TRY(d2d(u.gskip[l], g, sizeof(float) * cells(l), st));
    for (int j = post - 1; j >= 0; --j) {
      if (mask[l]) TRY(npde::mul_fg(g, mask[l], mbs[l], sz[l], B, st));
      TRY(layer_bwd(g, u.Sx[l][j], w_second + 9 * (i * post + j), other, sz[l], gw_second + 9 * (i * post + j)));
      swap();
    }
  }
  // g = dL/dT_{L-1} through the post-smoothing path of the deepest level; add the skip path
  TRY(npde::add_inplace(g, u.gskip[L - 1], cells(L - 1), st));
  for (int l = L - 1; l >= 0; --l) {
    for (int j = pre - 1; j >= 0; --j) {
      if (mask[l]) TRY(npde::mul_fg(g, mask[l], mbs[l], sz[l], B, st));
      TRY(layer_bwd(g, u.F[l][j], w_first + 9 * (l * pre + j), other, sz[l], gw_first + 9 * (l * pre + j)));
      swap();
    }
    if (l > 0) {  // the stride-2 pooling conv that produced this level's input (:63-66)
      if (mask[l]) TRY(npde::mul_fg(g, mask[l], mbs[l], sz[l], B, st));
      TRY(npde::weight_grad(g, sz[l], u.F[l - 1][pre], sz[l - 1], 2, 0, gw_pool + 9 * (l - 1), B, u.wg, u.wg_bytes,
                            st));
      TRY(npde::conv3x3_transpose(g, sz[l], w_pool + 9 * (l - 1), 2, 0, other, sz[l - 1], B, st));
      swap();
      TRY(npde::add_inplace(g, u.gskip[l - 1], cells(l - 1), st));
    }
  }
  if (gx) {
    if (mask[0]) TRY(npde::mul_fg(g, mask[0], mbs[0], sz[0], B, st));
    TRY(npde::bwd_gx(u.gz, g, gx, B, N, st));
  }
  return NPDE_OK;
}

size_t npde_workspace_bytes(int op, int B, int N, int bc_kind, int n_layers, int pre, int post, int k) {
  (void)k;
  if (B < 1 || N < 3) return 0;
  Bump b(nullptr);
  switch (op) {
    case NPDE_WS_REDUCE:
      return npde::reduce_ws_bytes(B, N);
    case NPDE_WS_CONV_FWD: {
      ConvFwdWs c;
      c.carve(b, B, N);
      return b.off;
    }
    case NPDE_WS_CONV_BWD: {
      ConvBwdWs c;
      int L = n_layers < 0 ? 0 : (n_layers > NPDE_MAX_CONV_LAYERS ? NPDE_MAX_CONV_LAYERS : n_layers);
      c.carve(b, B, N, L);
      return b.off;
    }
    case NPDE_WS_MG: {
      if (n_layers < 1 || n_layers > NPDE_MAX_LEVELS) return 0;
      MgWs m;
      m.carve(b, B, N, bc_kind, n_layers, true);
      return b.off;
    }
    case NPDE_WS_UNET_FWD: {
      if (n_layers < 1 || n_layers > NPDE_MAX_LEVELS) return 0;
      UnetWs u;
      u.carve(b, B, N, n_layers, bc_kind == NPDE_BC_MASK);
      b.take<float>((size_t)B * (N - 2) * (N - 2));
      b.take<float>((size_t)B * (N - 2) * (N - 2));
      UnetFusedWs fw;  // frames of the fused finest level (npde_unet_step_fwd_hw)
      fw.carve(b, B, N);
      return b.off;
    }
    case NPDE_WS_UNET_BWD: {
      if (n_layers < 1 || n_layers > NPDE_MAX_LEVELS || pre < 0 || post < 0 || pre > NPDE_MAX_UNET_SMOOTH ||
          post > NPDE_MAX_UNET_SMOOTH)
        return 0;
      UnetBwdWs u;
      u.carve(b, B, N, n_layers, pre, post, bc_kind == NPDE_BC_MASK);
      return b.off;
    }
    case NPDE_WS_CG:
      return npde::cg_ws_bytes(B, N);
    case NPDE_WS_CONV_BPTT: {
      ConvBwdWs c;
      int L = n_layers < 0 ? 0 : (n_layers > NPDE_MAX_CONV_LAYERS ? NPDE_MAX_CONV_LAYERS : n_layers);
      c.carve(b, B, N, L);
      b.take<float>((size_t)B * N * N);
      b.take<float>((size_t)B * N * N);
      return b.off;
    }
    case NPDE_WS_ITERATE + NPDE_IT_JACOBI:
    case NPDE_WS_ITERATE + NPDE_IT_CONV:
    case NPDE_WS_ITERATE + NPDE_IT_MULTIGRID:
    case NPDE_WS_ITERATE + NPDE_IT_UNET: {
      IterWs w;
      w.carve(b, op - NPDE_WS_ITERATE, B, N, bc_kind, n_layers, pre, post);
      return b.off;
    }
    default:
      return 0;
  }
}

int npde_iterate(const npde_iterate_desc *d, void *stream) {
  if (!d) return NPDE_ERR_NULL;
  TRY(npde::check_common(d->x_in, d->x_out, d->B, d->N));
  TRY(npde::check_bc(d->bc, d->bc_kind));
  if (d->k < 0) return NPDE_ERR_ARG;
  if (d->err_l2 && !d->gt) return NPDE_ERR_NULL;
  const int B = d->B, N = d->N, it = d->iterator;
  if (it < NPDE_IT_JACOBI || it > NPDE_IT_UNET) return NPDE_ERR_ARG;
  if (it == NPDE_IT_CONV || it == NPDE_IT_UNET) {
    if (d->act < NPDE_ACT_NONE || d->act > NPDE_ACT_SIGMOID) return NPDE_ERR_ACT;
  }
  if (it == NPDE_IT_CONV && (d->n_layers < 0 || d->n_layers > NPDE_MAX_CONV_LAYERS)) return NPDE_ERR_LAYERS;
  if (it == NPDE_IT_CONV && d->n_layers > 0 && !d->w) return NPDE_ERR_NULL;
  if (it == NPDE_IT_MULTIGRID || it == NPDE_IT_UNET) {
    if (d->n_layers < 1 || d->n_layers > NPDE_MAX_LEVELS || d->pre < 0 || d->post < 0) return NPDE_ERR_LAYERS;
    if (!level_sizes_ok(N, d->n_layers)) return NPDE_ERR_SIZE;
    if (it == NPDE_IT_UNET && !d->w) return NPDE_ERR_NULL;
  }
  int depth = d->temporal_depth <= 0 ? 4 : d->temporal_depth;  // 4 measured fastest at 1025^2 (halo 4 = one lane)
  if (d->temporal_depth < -1 || (depth != 1 && depth != 2 && depth != 4 && depth != 8)) return NPDE_ERR_ARG;
  // grids that fit in shared memory: all k steps (and their error rows) in one kernel, x read and written once.
  // An explicit temporal_depth selects the streaming kernels below instead.
  if (d->temporal_depth == 0 && d->k >= 1 && npde::small_iterate_supported(it, N, d->n_layers, d->w_host))
    return npde::small_iterate(it, d->x_in, d->x_out, d->bc, d->bc_kind, d->f, d->w_host, d->n_layers, d->act, d->k,
                               d->gt, d->err_l2, d->err_fd, B, N, S(stream));
  size_t need = npde_workspace_bytes(NPDE_WS_ITERATE + it, B, N, d->bc_kind, d->n_layers, d->pre, d->post, d->k);
  TRY(ws_ok(d->workspace, d->workspace_bytes, need));
  cudaStream_t st = S(stream);
  Bump b(d->workspace);
  IterWs w;
  w.carve(b, it, B, N, d->bc_kind, d->n_layers, d->pre, d->post);
  const size_t cells = (size_t)B * N * N;
  const bool want_err = d->err_l2 || d->err_fd;
  npde_iterate_desc dd = *d;
#ifndef NPDE_EMU
  if (w.bc_copy && graph_worthwhile(cells, d->k)) {
    TRY(d2d(w.bc_copy, d->bc, sizeof(float) * w.bc_floats, st));
    dd.bc = w.bc_copy;
  }
#endif
  d = &dd;
  const size_t n_whost = it == NPDE_IT_CONV ? 9 * (size_t)d->n_layers
                        : it == NPDE_IT_UNET ? 9 * ((size_t)d->n_layers * (d->pre + d->post) + d->n_layers - 1) : 0;

  // ---- fast path: fused stream kernels on pair-permuted ping-pong buffers.  Error rows, if asked for,
  // come from the reduction kernels reading the same buffers (one launch per row, no host round trip);
  // they force one iterator application per launch.
  const bool conv_fast = it == NPDE_IT_CONV && npde::conv_stream_supported(d->bc_kind, d->n_layers, d->w_host);
  const bool jac_fast = it == NPDE_IT_JACOBI;  // mask geometries: one step per launch
  if (conv_fast || jac_fast) {
    npde::FieldView cur = npde::ref_view(d->x_in, N);
    if (d->err_fd)
      TRY(npde::fd_error_view(cur, d->bc, d->bc_kind, d->f, NPDE_AGG_MAX, d->err_fd, B, N, w.red, w.red_bytes, st));
    if (d->k == 0) return d2d(d->x_out, d->x_in, sizeof(float) * cells, st);
    // ---- wide path (npde_wide.cu): pack once, k launches on the batch-interleaved W8 layout, unpack once.
    // Taken when there are enough strip-rows to fill the machine and no per-step error rows are wanted; an
    // explicit temporal_depth selects the first-generation stream kernel below.
    if (conv_fast && !want_err && d->k >= 3 && (d->temporal_depth == -1 || (d->temporal_depth == 0 && wide_worthwhile(B, N))) &&
        npde::conv_wide_supported(d->bc_kind, d->n_layers, d->w_host, B, N)) {
      const float *fw = nullptr;
      if (d->f) {
        TRY(npde::wide_pack(d->f, w.fpad, B, N, st));
        fw = w.fpad;
      }
      TRY(npde::wide_pack(d->x_in, w.buf[0], B, N, st));
      for (int t = 0; t < d->k; ++t)
        TRY(npde::conv_wide(w.buf[t & 1], w.buf[(t + 1) & 1], d->bc, fw, d->w_host, d->n_layers, d->act, B, N, st));
      return npde::wide_unpack(w.buf[d->k & 1], d->x_out, B, N, st);
    }
    npde::FieldViewMut pad[2] = {{w.buf[0], w.pitch, w.bstride, 1}, {w.buf[1], w.pitch, w.bstride, 1}};
    npde::FieldViewMut last = npde::ref_view_mut(d->x_out, N);
    // the source is read in every step: give it 16-byte aligned rows once (128-bit loads)
    npde::FieldView fv = npde::ref_view(d->f, N);
    if (d->f && d->k >= 3) {
      npde::FieldViewMut fp = {w.fpad, w.pitch, w.bstride, 1};
      TRY(npde::copy_field(fv, fp, B, N, st));
      fv = npde::as_const(fp);
    }
    // same for the two planes of a mask-geometry bc tensor
    npde::BcFields bcf = npde::bc_fields(d->bc, d->bc_kind, N);
    if (d->bc_kind == NPDE_BC_MASK && d->k >= 3) {
      npde::FieldViewMut vp = {w.vpad, w.pitch, w.bstride, 1}, mp = {w.mpad, w.pitch, w.bstride, 1};
      TRY(npde::copy_field(bcf.values, vp, B, N, st));
      TRY(npde::copy_field(bcf.mask, mp, B, N, st));
      bcf.values = npde::as_const(vp);
      bcf.mask = npde::as_const(mp);
    }
    int left = d->k, flip = 0, done = 0;
    const int full_adv =
        (jac_fast && !want_err) ? (((d->bc_kind == NPDE_BC_MASK || d->f) && depth > 4) ? 4 : depth) : 1;
    auto launch = [&](const npde::FieldView &src, const npde::FieldViewMut &dst, int adv, cudaStream_t s) -> int {
      if (conv_fast) return npde::conv_stream(src, dst, bcf, fv, d->w_host, d->n_layers, d->act, B, N, s);
      return npde::jacobi_stream(src, dst, bcf, fv, adv, B, N, s);
    };
#ifndef NPDE_EMU
    // launch-bound sizes: x_in -> pad[0] once, then replay a captured segment of SEG launches pad[0] -> pad[1] -> ...
    // -> pad[0]; the plain loop below finishes the remainder and writes the caller's tensor
    constexpr int SEG = ITER_SEG;
    if (graph_worthwhile(cells, d->k) && left > full_adv * (SEG + 1)) {
      StepGraph sg;
      const bool hit = sg.lookup(make_graph_key(d, full_adv, d->workspace, n_whost));
      if (hit || sg.begin(st)) {
        // error rows of a replayed segment land in staging rows (fixed addresses) and are copied to their place
        // in the caller's (k, B) / (k+1, B) arrays after every replay
        auto rows = [&](const npde::FieldView &v, float *l2_row, float *fd_row, cudaStream_t s) -> int {
          if (d->err_l2) TRY(npde::l2_error_view(v, d->gt, l2_row, B, N, w.red, w.red_bytes, s, true));
          if (d->err_fd)
            TRY(npde::fd_error_view(v, d->bc, d->bc_kind, d->f, NPDE_AGG_MAX, fd_row, B, N, w.red, w.red_bytes, s, true));
          return NPDE_OK;
        };
        if (want_err) TRY(npde::reduce_ws_clear(w.red, w.red_bytes, B, N, st));  // once; reductions keep them clean
        int rc = NPDE_OK;
        for (int i = 0; i < SEG && rc == NPDE_OK && !hit; ++i) {
          rc = launch(npde::as_const(pad[i & 1]), pad[(i + 1) & 1], full_adv, sg.cs);
          if (rc == NPDE_OK && want_err)
            rc = rows(npde::as_const(pad[(i + 1) & 1]), w.stage_l2 + (size_t)i * B, w.stage_fd + (size_t)i * B, sg.cs);
        }
        if (rc != NPDE_OK) { sg.abort_capture(); return rc; }
        TRY(launch(cur, pad[0], full_adv, st));
        left -= full_adv;
        done += full_adv;
        if (want_err)
          TRY(rows(npde::as_const(pad[0]), d->err_l2 ? d->err_l2 + (size_t)(done - 1) * B : nullptr,
                   d->err_fd ? d->err_fd + (size_t)done * B : nullptr, st));
        int times = (left - 1) / (full_adv * SEG);
        if (sg.finish() != NPDE_OK) times = 0;  // no graph after all: the plain loop below does the rest
        for (int r = 0; r < times; ++r) {
          TRY(sg.launch_once(st));
          if (d->err_l2)
            TRY(d2d(d->err_l2 + (size_t)done * B, w.stage_l2, sizeof(float) * (size_t)SEG * B, st));
          if (d->err_fd)
            TRY(d2d(d->err_fd + (size_t)(done + 1) * B, w.stage_fd, sizeof(float) * (size_t)SEG * B, st));
          left -= full_adv * SEG;
          done += full_adv * SEG;
        }
        cur = npde::as_const(pad[0]);
        flip = 1;
      }
    }
#endif
    while (left > 0) {
      int adv = full_adv;
      while (adv > left) adv >>= 1;
      // the final launch writes the caller's tensor directly (reference layout)
      bool final_launch = (left - adv == 0);
      bool alias = final_launch && (const void *)d->x_out == (const void *)cur.p;  // k == 1 in place
      npde::FieldViewMut dst = (final_launch && !alias) ? last : pad[flip];
      TRY(launch(cur, dst, adv, st));
      left -= adv;
      done += adv;
      if (final_launch && alias) {
        TRY(npde::copy_field(npde::as_const(dst), last, B, N, st));
        dst = last;
      }
      cur = npde::as_const(dst);
      flip ^= 1;
      if (d->err_l2)
        TRY(npde::l2_error_view(cur, d->gt, d->err_l2 + (size_t)(done - 1) * B, B, N, w.red, w.red_bytes, st));
      if (d->err_fd)
        TRY(npde::fd_error_view(cur, d->bc, d->bc_kind, d->f, NPDE_AGG_MAX, d->err_fd + (size_t)done * B, B, N, w.red,
                                w.red_bytes, st));
    }
    return NPDE_OK;
  }

  // ---- generic loop: one step function per application, norms by the reduction kernels ----
  float *bufA = w.buf[0], *bufB = w.buf[1];
  const float *cur = d->x_in;
  const float *wf = d->w, *wp = nullptr, *wsnd = nullptr;
  if (it == NPDE_IT_UNET) {
    wp = wf + 9 * (size_t)d->n_layers * d->pre;
    wsnd = wp + 9 * (size_t)(d->n_layers - 1);
  }
  auto step = [&](const float *src, float *dst, void *s) -> int {
    switch (it) {
      case NPDE_IT_JACOBI: {
        const float *res = nullptr;
        return npde::jacobi_steps(src, dst, nullptr, d->bc, d->bc_kind, d->f, 1, B, N, &res, S(s));
      }
      case NPDE_IT_CONV:
        return npde_conv_step_fwd(src, d->bc, d->bc_kind, d->f, d->w, d->w_host, d->n_layers, d->act, dst, B, N,
                                  w.step, w.step_bytes, s);
      case NPDE_IT_MULTIGRID:
        return npde_mg_step(src, d->bc, d->bc_kind, d->f, d->n_layers, d->pre, d->post, dst, B, N, w.step,
                            w.step_bytes, s);
      default:
        return npde_unet_step_fwd_hw(src, d->bc, d->bc_kind, d->f, wf, wp, wsnd, d->w_host, d->n_layers, d->pre,
                                     d->post, d->act, dst, B, N, w.step, w.step_bytes, s);
    }
  };
  int t = 0;
#ifndef NPDE_EMU
  // launch-bound sizes: x_in -> A once, then replay the captured pair of steps A -> B -> A
  if (!want_err && graph_worthwhile(cells, d->k)) {
    StepGraph sg;
    const bool hit = sg.lookup(make_graph_key(d, 2, d->workspace, n_whost));
    if (hit || sg.begin(st)) {
      int rc = hit ? NPDE_OK : step(bufA, bufB, sg.cs);
      if (rc == NPDE_OK && !hit) rc = step(bufB, bufA, sg.cs);
      if (rc != NPDE_OK) { sg.abort_capture(); return rc; }
      TRY(step(cur, bufA, stream));
      int times = (d->k - 1) / 2;
      if (sg.finish() != NPDE_OK) times = 0;  // no graph after all: the plain loop below does the rest
      for (int r = 0; r < times; ++r) TRY(sg.launch_once(st));
      cur = bufA;
      t = 1 + 2 * times;
    }
  }
#endif
  for (; t < d->k; ++t) {
    if (d->err_fd)
      TRY(npde::fd_error(cur, d->bc, d->bc_kind, d->f, NPDE_AGG_MAX, d->err_fd + (size_t)t * B, B, N, w.red,
                         w.red_bytes, st));
    float *nxt = (cur == bufA) ? bufB : bufA;
    TRY(step(cur, nxt, stream));
    cur = nxt;
    if (d->err_l2) TRY(npde::l2_error(cur, d->gt, d->err_l2 + (size_t)t * B, B, N, w.red, w.red_bytes, st));
  }
  if (d->err_fd)
    TRY(npde::fd_error(cur, d->bc, d->bc_kind, d->f, NPDE_AGG_MAX, d->err_fd + (size_t)d->k * B, B, N, w.red,
                       w.red_bytes, st));
  return d2d(d->x_out, cur, sizeof(float) * cells, st);
}

int npde_solve(const npde_iterate_desc *d, float tol, int check_every, int max_iters, void *result, void *stream) {
  if (!d || !result) return NPDE_ERR_NULL;
  TRY(npde::check_common(d->x_in, d->x_out, d->B, d->N));
  TRY(npde::check_bc(d->bc, d->bc_kind));
  if (check_every < 1 || max_iters < 1 || !(tol > 0.0f)) return NPDE_ERR_ARG;
  const int B = d->B, N = d->N, it = d->iterator;
  if (it < NPDE_IT_JACOBI || it > NPDE_IT_UNET) return NPDE_ERR_ARG;
  if ((it == NPDE_IT_CONV || it == NPDE_IT_UNET) && (d->act < NPDE_ACT_NONE || d->act > NPDE_ACT_SIGMOID))
    return NPDE_ERR_ACT;
  if (it == NPDE_IT_CONV && (d->n_layers < 0 || d->n_layers > NPDE_MAX_CONV_LAYERS)) return NPDE_ERR_LAYERS;
  if (it == NPDE_IT_CONV && d->n_layers > 0 && !d->w) return NPDE_ERR_NULL;
  if (it == NPDE_IT_MULTIGRID || it == NPDE_IT_UNET) {
    if (d->n_layers < 1 || d->n_layers > NPDE_MAX_LEVELS || d->pre < 0 || d->post < 0) return NPDE_ERR_LAYERS;
    if (!level_sizes_ok(N, d->n_layers)) return NPDE_ERR_SIZE;
    if (it == NPDE_IT_UNET && !d->w) return NPDE_ERR_NULL;
  }
  size_t need = npde_workspace_bytes(NPDE_WS_ITERATE + it, B, N, d->bc_kind, d->n_layers, d->pre, d->post, check_every);
  TRY(ws_ok(d->workspace, d->workspace_bytes, need));
  cudaStream_t st = S(stream);
  Bump b(d->workspace);
  IterWs w;
  w.carve(b, it, B, N, d->bc_kind, d->n_layers, d->pre, d->post);
  SolveResult *res = reinterpret_cast<SolveResult *>(result);
  float *err = w.stage_fd;  // (B) residual row of the current check
  const bool conv_fast = it == NPDE_IT_CONV && npde::conv_stream_supported(d->bc_kind, d->n_layers, d->w_host);
  const bool fast = conv_fast || it == NPDE_IT_JACOBI;  // stream kernels on pair-permuted fields
  npde::FieldViewMut pad[2] = {{w.buf[0], w.pitch, w.bstride, 1}, {w.buf[1], w.pitch, w.bstride, 1}};
  if (!fast) pad[0] = npde::ref_view_mut(w.buf[0], N), pad[1] = npde::ref_view_mut(w.buf[1], N);
  npde::FieldView fv = npde::ref_view(d->f, N);
  npde::BcFields bcf = npde::bc_fields(d->bc, d->bc_kind, N);
  const float *wf = d->w, *wp = nullptr, *wsnd = nullptr;
  if (it == NPDE_IT_UNET) {
    wp = wf + 9 * (size_t)d->n_layers * d->pre;
    wsnd = wp + 9 * (size_t)(d->n_layers - 1);
  }
  // prologue: x_in -> pad[0] (the source and the bc planes get their 128-bit layout once), residual of x_0
  auto prologue = [&](cudaStream_t s) -> int {
    if (fast && d->f) {
      npde::FieldViewMut fp = {w.fpad, w.pitch, w.bstride, 1};
      TRY(npde::copy_field(fv, fp, B, N, s));
      fv = npde::as_const(fp);
    }
    if (fast && d->bc_kind == NPDE_BC_MASK) {
      npde::FieldViewMut vp = {w.vpad, w.pitch, w.bstride, 1}, mp = {w.mpad, w.pitch, w.bstride, 1};
      TRY(npde::copy_field(bcf.values, vp, B, N, s));
      TRY(npde::copy_field(bcf.mask, mp, B, N, s));
      bcf.values = npde::as_const(vp);
      bcf.mask = npde::as_const(mp);
    }
    TRY(npde::copy_field(npde::ref_view(d->x_in, N), pad[0], B, N, s));
    TRY(npde::reduce_ws_clear(w.red, w.red_bytes, B, N, s));
    return npde::fd_error_view(npde::as_const(pad[0]), d->bc, d->bc_kind, d->f, NPDE_AGG_MAX, err, B, N, w.red,
                               w.red_bytes, s, true);
  };
  // one check interval: check_every applications pad[0] -> ... -> pad[0], then the residual row
  auto body = [&](cudaStream_t s) -> int {
    int cur = 0, left = check_every;
    while (left > 0) {
      int adv = 1;
      if (it == NPDE_IT_JACOBI) adv = left >= 4 ? 4 : (left >= 2 ? 2 : 1);
      const npde::FieldView src = npde::as_const(pad[cur]);
      const npde::FieldViewMut &dst = pad[cur ^ 1];
      if (it == NPDE_IT_JACOBI) TRY(npde::jacobi_stream(src, dst, bcf, fv, adv, B, N, s));
      else if (conv_fast) TRY(npde::conv_stream(src, dst, bcf, fv, d->w_host, d->n_layers, d->act, B, N, s));
      else if (it == NPDE_IT_CONV)
        TRY(npde_conv_step_fwd(src.p, d->bc, d->bc_kind, d->f, d->w, d->w_host, d->n_layers, d->act, dst.p, B, N, w.step,
                               w.step_bytes, s));
      else if (it == NPDE_IT_MULTIGRID)
        TRY(npde_mg_step(src.p, d->bc, d->bc_kind, d->f, d->n_layers, d->pre, d->post, dst.p, B, N, w.step, w.step_bytes, s));
      else
        TRY(npde_unet_step_fwd_hw(src.p, d->bc, d->bc_kind, d->f, wf, wp, wsnd, d->w_host, d->n_layers, d->pre, d->post,
                                  d->act, dst.p, B, N, w.step, w.step_bytes, s));
      left -= adv;
      cur ^= 1;
    }
    if (cur != 0) TRY(npde::copy_field(npde::as_const(pad[1]), pad[0], B, N, s));
    return npde::fd_error_view(npde::as_const(pad[0]), d->bc, d->bc_kind, d->f, NPDE_AGG_MAX, err, B, N, w.red,
                               w.red_bytes, s, true);
  };
  auto epilogue = [&](cudaStream_t s) -> int {
    return npde::copy_field(npde::as_const(pad[0]), npde::ref_view_mut(d->x_out, N), B, N, s);
  };
#ifdef NPDE_EMU
  // host emulator: "device" memory is host memory, so the loop condition is simply read back
  TRY(prologue(st));
  NPDE_LAUNCH(k_solve_check, dim3(1), dim3(32), 0, st, err, B, tol, 0, max_iters, res, 1);
  while (!res->converged && res->iterations < max_iters) {
    TRY(body(st));
    NPDE_LAUNCH(k_solve_check, dim3(1), dim3(32), 0, st, err, B, tol, check_every, max_iters, res, 0);
  }
  return epilogue(st);
#else
  // One graph: prologue -> WHILE(residual >= tol and iterations < max_iters) { body } -> epilogue.  The condition is
  // written by k_solve_check on the device, so the host neither waits nor reads anything until the caller does.
  cudaStreamCaptureStatus status = cudaStreamCaptureStatusNone;
  if (cudaStreamIsCapturing(st, &status) != cudaSuccess || status != cudaStreamCaptureStatusNone) {
    cudaGetLastError();
    return NPDE_ERR_ARG;  // cannot build a graph inside the caller's capture
  }
  struct Guard {
    cudaStream_t cs = nullptr, cs2 = nullptr;
    cudaGraph_t g = nullptr;
    cudaGraphExec_t ex = nullptr;
    ~Guard() {
      if (ex) cudaGraphExecDestroy(ex);  // an in-flight launch completes; resources are released afterwards
      if (g) cudaGraphDestroy(g);
      if (cs) cudaStreamDestroy(cs);
      if (cs2) cudaStreamDestroy(cs2);
    }
  } G;
#define CU(expr) do { cudaError_t e__ = (expr); if (e__ != cudaSuccess) { cudaGetLastError(); return (int)e__; } } while (0)
  CU(cudaStreamCreateWithFlags(&G.cs, cudaStreamNonBlocking));
  CU(cudaStreamCreateWithFlags(&G.cs2, cudaStreamNonBlocking));
  CU(cudaGraphCreate(&G.g, 0));
  cudaGraphConditionalHandle handle;
  CU(cudaGraphConditionalHandleCreate(&handle, G.g, 1, cudaGraphCondAssignDefault));
  CU(cudaStreamBeginCaptureToGraph(G.cs, G.g, nullptr, nullptr, 0, cudaStreamCaptureModeRelaxed));
  int rc = prologue(G.cs);
  if (rc == NPDE_OK) {
    NPDE_LAUNCH(k_solve_check, dim3(1), dim3(32), 0, G.cs, err, B, tol, 0, max_iters, res, 1, handle);
    if (cudaGetLastError() != cudaSuccess) rc = NPDE_ERR_ARG;
  }
  cudaGraphNode_t cond = nullptr;
  cudaGraph_t body_graph = nullptr;
  if (rc == NPDE_OK) {
    const cudaGraphNode_t *deps = nullptr;
    size_t n_deps = 0;
    cudaGraph_t gq = nullptr;
    unsigned long long id = 0;
    cudaError_t e = cudaStreamGetCaptureInfo(G.cs, &status, &id, &gq, &deps, &n_deps);
    cudaGraphNodeParams np = {};
    np.type = cudaGraphNodeTypeConditional;
    np.conditional.handle = handle;
    np.conditional.type = cudaGraphCondTypeWhile;
    np.conditional.size = 1;
    if (e == cudaSuccess) e = cudaGraphAddNode(&cond, G.g, deps, n_deps, &np);
    if (e == cudaSuccess) {
      body_graph = np.conditional.phGraph_out[0];
      e = cudaStreamUpdateCaptureDependencies(G.cs, &cond, 1, cudaStreamSetCaptureDependencies);
    }
    if (e != cudaSuccess) rc = (int)e;
  }
  if (rc == NPDE_OK) rc = epilogue(G.cs);
  {
    cudaGraph_t done = nullptr;
    cudaError_t e = cudaStreamEndCapture(G.cs, &done);
    if (rc == NPDE_OK && e != cudaSuccess) rc = (int)e;
  }
  if (rc != NPDE_OK) { cudaGetLastError(); return rc; }
  CU(cudaStreamBeginCaptureToGraph(G.cs2, body_graph, nullptr, nullptr, 0, cudaStreamCaptureModeRelaxed));
  rc = body(G.cs2);
  if (rc == NPDE_OK) {
    NPDE_LAUNCH(k_solve_check, dim3(1), dim3(32), 0, G.cs2, err, B, tol, check_every, max_iters, res, 0, handle);
    if (cudaGetLastError() != cudaSuccess) rc = NPDE_ERR_ARG;
  }
  {
    cudaError_t e = cudaStreamEndCapture(G.cs2, nullptr);
    if (rc == NPDE_OK && e != cudaSuccess) rc = (int)e;
  }
  if (rc != NPDE_OK) { cudaGetLastError(); return rc; }
  CU(cudaGraphInstantiate(&G.ex, G.g, 0));
  CU(cudaGraphLaunch(G.ex, st));
#undef CU
  return NPDE_OK;
#endif
}

size_t npde_packed_floats(int B, int N) { return (B > 0 && N > 0) ? npde::wide_field_floats(B, N) : 0; }

int npde_pack(const float *x, float *packed, int B, int N, void *stream) {
  TRY(npde::check_common(x, packed, B, N));
  if ((uintptr_t)packed & 31) return NPDE_ERR_ARG;
  return npde::wide_pack(x, packed, B, N, S(stream));
}

int npde_unpack(const float *packed, float *x, int B, int N, void *stream) {
  TRY(npde::check_common(packed, x, B, N));
  if ((uintptr_t)packed & 31) return NPDE_ERR_ARG;
  return npde::wide_unpack(packed, x, B, N, S(stream));
}

int npde_conv_step_packed(const float *in_packed, float *out_packed, const float *bc, const float *f_packed,
                          const float *w_host, int n_layers, int act, int B, int N, void *stream) {
  TRY(npde::check_common(in_packed, out_packed, B, N));
  if (!bc || !w_host) return NPDE_ERR_NULL;
  if (act < NPDE_ACT_NONE || act > NPDE_ACT_SIGMOID) return NPDE_ERR_ACT;
  if (n_layers < 1 || n_layers > 4) return NPDE_ERR_LAYERS;
  return npde::conv_wide(in_packed, out_packed, bc, f_packed, w_host, n_layers, act, B, N, S(stream));
}

int npde_conv_iterate_packed(float *a_packed, float *b_packed, const float *bc, const float *f_packed,
                             const float *w_host, int n_layers, int act, int B, int N, int k, void *stream) {
  TRY(npde::check_common(a_packed, b_packed, B, N));
  if (!bc || !w_host) return NPDE_ERR_NULL;
  if (act < NPDE_ACT_NONE || act > NPDE_ACT_SIGMOID) return NPDE_ERR_ACT;
  if (n_layers < 1 || n_layers > 4) return NPDE_ERR_LAYERS;
  if (k < 0) return NPDE_ERR_ARG;
  float *buf[2] = {a_packed, b_packed};
  for (int t = 0; t < k; ++t)
    TRY(npde::conv_wide(buf[t & 1], buf[(t + 1) & 1], bc, f_packed, w_host, n_layers, act, B, N, S(stream)));
  return NPDE_OK;
}

int npde_fd_error_packed(const float *packed, const float *f_packed, float *err, int B, int N, void *stream) {
  if (B < 1 || N < 1) return NPDE_ERR_SIZE;
  return npde::wide_fd_error_max(packed, f_packed, err, B, N, S(stream));
}

}  // extern "C"
