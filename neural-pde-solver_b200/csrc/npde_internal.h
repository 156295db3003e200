// npde_internal.h -- host-side entry points shared between the translation units
// of libnpde_b200.  All functions enqueue on `st` and return NPDE_OK / error.
#pragma once
#include "npde_common.cuh"

namespace npde {

int check_common(const void *a, const void *b, int B, int N);
int check_bc(const float *bc, int bc_kind);

// A (B,N,N) field with an arbitrary row pitch / batch stride (floats).  The reference layout
// is {p, N, N*N, perm = 0}.  The internal ping-pong buffers of the k-step loop use the
// "pair-permuted padded layout" (perm = 1): 16-byte aligned rows of pitch >= round_up(N, 4)
// in which every aligned group of four columns (c0,c1,c2,c3) is stored as (c0,c2,c1,c3), the
// register-pair order of the stream kernels (npde_stream.cu), so a lane moves its four
// columns with one 128-bit access.
struct FieldView {
  const float *p;
  int pitch;
  long bstride;
  int perm;
};
struct FieldViewMut {
  float *p;
  int pitch;
  long bstride;
  int perm;
};
inline FieldView ref_view(const float *p, int N) { return FieldView{p, N, (long)N * N, 0}; }  // p may be nullptr
inline FieldViewMut ref_view_mut(float *p, int N) { return FieldViewMut{p, N, (long)N * N, 0}; }
// layout conversion / copy between any two views
int copy_field(const FieldView &src, const FieldViewMut &dst, int B, int N, cudaStream_t st);  // npde_grid_ops.cu
inline FieldView as_const(const FieldViewMut &v) { return FieldView{v.p, v.pitch, v.bstride, v.perm}; }

// square domain, 1 <= n_layers <= 4 and a HOST copy of the weights (they travel as kernel parameters)
bool conv_stream_supported(int bc_kind, int n_layers, const float *w_host);
// npde_small.cu: the whole k-step loop of a grid that fits in shared memory (N <= NPDE_SMALL_MAX_N) in one kernel
bool small_iterate_supported(int iterator, int N, int n_layers, const float *w_host);
int small_iterate(int iterator, const float *x_in, float *x_out, const float *bc, int bc_kind, const float *f,
                  const float *w_host, int n_layers, int act, int k, const float *gt, float *err_l2, float *err_fd,
                  int B, int N, cudaStream_t st);
// npde_cg.cu
size_t cg_ws_bytes(int B, int N);
int cg_step(const float *x, int n_iters, float *y, int B, int N, void *ws, cudaStream_t st);
// npde_grid_ops.cu
int reduce_chunks(int N);
size_t reduce_ws_bytes(int B, int N);
int set_boundary(float *x, const float *bc, int bc_kind, int B, int N, cudaStream_t st);
int initialize(float *x, const float *bc, int bc_kind, int mode, const float *rnd, int B, int N, cudaStream_t st);
int pad_boundary(const float *xin, const float *bc, int bc_kind, float *y, int B, int N, cudaStream_t st);
int fd_step_simple(const float *x, const float *bc, int bc_kind, const float *f, float *y, int B, int N,
                   cudaStream_t st);
int fd_error(const float *x, const float *bc, int bc_kind, const float *f, int aggregate, float *err, int B,
             int N, void *ws, size_t ws_bytes, cudaStream_t st);
int l2_error(const float *x, const float *gt, float *err, int B, int N, void *ws, size_t ws_bytes,
             cudaStream_t st);
// same on any field view (the internal buffers of the k-step loop); bc, f and gt stay in the reference layout
int fd_error_view(const FieldView &x, const float *bc, int bc_kind, const float *f, int aggregate, float *err,
                  int B, int N, void *ws, size_t ws_bytes, cudaStream_t st, bool counters_clean = false);
// counters_clean: the reduction counters of `ws` are known to be zero (every reduction leaves them so)
int reduce_ws_clear(void *ws, size_t ws_bytes, int B, int N, cudaStream_t st);
int l2_error_view(const FieldView &x, const float *gt, float *err, int B, int N, void *ws, size_t ws_bytes,
                  cudaStream_t st, bool counters_clean = false);
int subsample(const float *x, float *y, int B, int N, float scale, cudaStream_t st);
// same, input images `x_bstride` floats apart (the mask plane of a (B,2,N,N) bc tensor)
int subsample_strided(const float *x, size_t x_bstride, float *y, int B, int N, float scale, cudaStream_t st);
int restrict_(const float *x, const float *bc_sub, int bc_kind, float *y, int B, int N, cudaStream_t st);
int prolong(const float *x, const float *bc, int bc_kind, float *y, int B, int M, cudaStream_t st);
// same on field views (the V-cycle keeps its levels in the pair-permuted layout); bc in the reference layout
int restrict_view(const FieldView &x, const float *bc_sub, int bc_kind, const FieldViewMut &y, int B, int N,
                  cudaStream_t st);
int prolong_view(const FieldView &x, const float *bc, int bc_kind, const FieldViewMut &y, int B, int M,
                 cudaStream_t st);

// npde_conv_ops.cu -- layer-by-layer building blocks (UNet, deep conv stacks, backward)
// mask: level mask (B, S+2, S+2) with batch stride mask_bstride floats, or NULL.
// add: optional (B,Ho,Ho) tensor added to the (masked) result -- the U-Net's skip connection fused into the conv
int conv3x3(const float *in, int Hi, const float *w9, int stride, int pad, const float *mask,
            size_t mask_bstride, float *out, int Ho, int B, cudaStream_t st, const float *add = nullptr);
int conv3x3_transpose(const float *g, int Hg, const float *w9, int stride, int pad, float *out, int Ho, int B,
                      cudaStream_t st);
// same with the input read through an array view / the output written through one (frames of the fused U-Net path)
int conv3x3_view(const ArrView &in, int Hi, const float *w9, int stride, int pad, const float *mask,
                 size_t mask_bstride, float *out, int Ho, int B, cudaStream_t st, const float *add = nullptr);
int pad_up_crop_view(const float *r, int s, const float *mask_out, size_t mask_bstride, const ArrViewMut &out, int B,
                     cudaStream_t st);
int mul_fg(float *r, const float *mask, size_t mask_bstride, int S, int B, cudaStream_t st);
int add_inplace(float *a, const float *b, size_t n, cudaStream_t st);
int psi_residual(const float *x, const float *f, const float *mask, size_t mask_bstride, float *y_int,
                 float *r0, int B, int N, cudaStream_t st);
int combine(const float *y_int, const float *h, const float *bc, int bc_kind, int act, float *out, int B, int N,
            cudaStream_t st);
int pad_up_crop(const float *r, int s, const float *mask_out, size_t mask_bstride, float *out, int B,
                cudaStream_t st);
int pad_up_crop_transpose(const float *g, int s, float *out, int B, cudaStream_t st);
int bwd_gz(const float *gout, const float *y_int, const float *rL, const float *mask, size_t mask_bstride,
           int act, float *gz, int B, int N, cudaStream_t st);
// gw9 (9 floats) <- sum_{b,i,j} g[b,i,j] * r[b, i*stride + a - pad, j*stride + c - pad]
// accumulate: gw9 += (fp32) instead of gw9 = (backprop through several steps sums in fp32, like autograd)
int weight_grad(const float *g, int Hg, const float *r, int Hr, int stride, int pad, float *gw9, int B,
                void *ws, size_t ws_bytes, cudaStream_t st, bool accumulate = false);
size_t weight_grad_ws_bytes(int B, int Hg);
// fused backward layer of a same-size stack: t = conv^T(g; w9), gw9 (+)= g (x) r; *counter must be 0
size_t bwd_layer_ws_bytes(int B, int S);
int bwd_layer(const float *g, const float *r, const float *w9, float *t, int S, float *gw9, int B, void *ws,
              size_t ws_bytes, unsigned *counter, bool accumulate, cudaStream_t st);
int bwd_gx(const float *gz, const float *g0, float *gx, int B, int N, cudaStream_t st);

// npde_stream.cu -- fused register-streaming kernels (Conv L-layer Phi_H, Jacobi T-step)
constexpr int STREAM_THREADS = 32;  // one warp strip per CTA (see decode_item in npde_stream.cu)

// Boundary data of the stream kernels: the (B,4) scalars of the square domain, or the two planes of a
// (B,2,N,N) mask-geometry bc tensor as field views (reference layout: pitch N, batch stride 2*N*N).
struct BcFields {
  const float *bc4;
  FieldView values, mask;
};
inline BcFields bc_fields(const float *bc, int bc_kind, int N) {
  if (bc_kind == NPDE_BC_MASK)
    return BcFields{nullptr, FieldView{bc, N, 2L * N * N, 0}, FieldView{bc + (size_t)N * N, N, 2L * N * N, 0}};
  return BcFields{bc, FieldView{nullptr, 0, 0, 0}, FieldView{nullptr, 0, 0, 0}};
}
// f.p == nullptr: Laplace.  perm views are accessed with 128-bit loads / stores, natural ones with 32-bit.
int conv_stream(const FieldView &in, const FieldViewMut &out, const BcFields &bc, const FieldView &f,
                const float *w_host, int n_layers, int act, int B, int N, cudaStream_t st);
// head of a U-Net step (npde_stream.cu): r = pre-smoothing convs of the masked residual, y = Psi(x) - f; rout and
// yout are frames with the same pitch / batch stride / layout
int conv_head_stream(const FieldView &in, const BcFields &bc, const FieldView &f, const float *w_host, int n_layers,
                     const FieldViewMut &rout, float *yout, int B, int N, cudaStream_t st);
// tail of a U-Net step: out = G(pad(act(y + (post-smoothing convs(rin) + skip)))); rin, skip and y are frames with
// rin's pitch / batch stride / layout
int conv_tail_stream(const FieldView &rin, const float *skip, const float *y, const BcFields &bc, const float *w_host,
                     int n_layers, int act, const FieldViewMut &out, int B, int N, cudaStream_t st);
// depth in {1,2,4,8} Jacobi steps per launch (1 when f != NULL); square domain only
int jacobi_stream_max_depth(const float *f);
// fd_error(x, 'max') on the square domain, streaming (any view of x and f); zeroes and fills err (B)
int fd_error_max_stream(const FieldView &x, const FieldView &f, float *err, int B, int N, cudaStream_t st);
// mask geometries (bc.mask.p != nullptr): depth 1 only
int jacobi_stream(const FieldView &in, const FieldViewMut &out, const BcFields &bc, const FieldView &f, int depth,
                  int B, int N, cudaStream_t st);


// npde_unet_coarse.cu -- U-Net levels l0 .. L-1 (interiors <= 127^2) in one kernel, one CTA per problem, on chip:
// pools `src` (the skip tensor of level l0-1), runs the deeper levels, writes the upsampled correction of level l0-1
bool unet_coarse_supported(const int *s, int L, int l0, int pre, int post, int B);
int unet_coarse(const ArrView &src, int s_src, const ArrViewMut &up, const float *wf, const float *wp,
                const float *wsnd, float *const *mask, const size_t *mbs, const int *s, int L, int pre, int post, int l0,
                int B, cudaStream_t st);
// npde_bwd_stream.cu -- backward of one Conv step as one register-streaming pass (square domain, 1..3 layers)
bool conv_bwd_stream_supported(int bc_kind, int n_layers);
size_t conv_bwd_stream_ws_bytes(int B, int N, int n_layers);
int conv_bwd_stream(const float *x, const float *f, const float *w_dev, int n_layers, int act, const float *gout,
                    float *gw, bool accumulate, float *gx, int B, int N, void *ws, size_t ws_bytes, cudaStream_t st);
// npde_wide.cu -- second-generation fused Phi_H kernel on the batch-interleaved "W8" layout (square domain):
// element (b,i,c) of a (B,N,N) field lives at  i*pitch + (b*G8 + c/8)*8 + perm8(c%8),  G8 = ceil(N/8),
// pitch = 8*B*G8, perm8 = (c0,c4,c1,c5,c2,c6,c3,c7); padding columns hold zeros.
size_t wide_field_floats(int B, int N);
bool conv_wide_supported(int bc_kind, int n_layers, const float *w_host, int B, int N);
int wide_pack(const float *x_ref, float *wide, int B, int N, cudaStream_t st);
int wide_unpack(const float *wide, float *x_ref, int B, int N, cudaStream_t st);
// fd_error('max') of a W8 field (f_wide: the source in the W8 layout or nullptr) -> err[B]
int wide_fd_error_max(const float *wide, const float *f_wide, float *err, int B, int N, cudaStream_t st);
// one Phi_H application W8 -> W8 (in != out, 32-byte aligned); f_wide: the source in the W8 layout or nullptr
int conv_wide(const float *in, float *out, const float *bc4, const float *f_wide, const float *w_host, int n_layers,
              int act, int B, int N, cudaStream_t st);

}  // namespace npde
