// ubench_stream.cu -- HBM bandwidth of the access patterns the stream kernels can use: every warp
// (one per CTA, 16 per SM) copies 512-byte row segments top to bottom, one row in flight in
// registers and AHEAD rows ahead in L2 (prefetch.global.L2), like k_jacobi_stream<1> without math.
//   pattern 0: row-major padded images (pitch 1028 floats), warp = (image, band, strip): the
//              segments of a warp are 4112 bytes apart, the 9 strips of a row are adjacent.
//   pattern 1: strip-major tiles: the segments of a warp are contiguous (480 or 512 bytes per row).
// Design input for DESIGN.md (memory layout of the internal ping-pong fields); not part of the library.
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>

__device__ __forceinline__ void prefetch_l2(const void *p) { asm volatile("prefetch.global.L2 [%0];" ::"l"(p)); }

template <int AHEAD>
__global__ void __launch_bounds__(32, 16)
k_copy(const float4 *__restrict__ in, float4 *__restrict__ out, long warp_stride, long row_stride, int rows,
       int strips, int bands, long band_stride, long img_stride, int lanes_used, int mode) {
  // strides in float4 units
  const int lane = threadIdx.x;
  long base;
  if (mode == 0) {
    const int strip = blockIdx.x % strips, band = (blockIdx.x / strips) % bands, b = blockIdx.x / (strips * bands);
    base = (long)b * img_stride + (long)band * band_stride + (long)strip * warp_stride;
  } else {
    base = (long)blockIdx.x * warp_stride;
  }
  const float4 *pi = in + base + lane;
  float4 *po = out + base + lane;
  const bool on = lane < lanes_used;
  float4 v = on ? __ldg(pi) : make_float4(0, 0, 0, 0);
  for (int r = 0; r < rows; ++r) {
    float4 nv = v;
    if (r + 1 < rows && on) nv = __ldg(pi + row_stride);
    if (AHEAD > 0 && r + 1 + AHEAD < rows && on) prefetch_l2(pi + (1 + AHEAD) * row_stride);
    v.x += 1.0f;
    if (on) *po = v;
    v = nv;
    pi += row_stride;
    po += row_stride;
  }
}

int main() {
  const int B = 64, N = 1025, strips = 9, bands = 4, band_rows = 257;
  const long pitch4 = 1028 / 4;
  size_t bytes = (size_t)B * 1028 * 1028 * 4 + (1 << 20);
  float4 *in, *out;
  cudaMalloc(&in, bytes * 2);
  cudaMalloc(&out, bytes * 2);
  cudaMemset(in, 0, bytes * 2);
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  const int warps = B * strips * bands;
  auto run = [&](const char *name, auto kern, long wstride, long rstride, int rows, long bstride, long istride,
                 int lanes, int mode, double useful_bytes) {
    for (int i = 0; i < 3; ++i) kern<<<warps, 32>>>(in, out, wstride, rstride, rows, strips, bands, bstride, istride, lanes, mode);
    cudaEventRecord(e0);
    const int reps = 20;
    for (int i = 0; i < reps; ++i) {
      kern<<<warps, 32>>>(in, out, wstride, rstride, rows, strips, bands, bstride, istride, lanes, mode);
      float4 *t = (float4 *)in; in = out; out = t;  // ping-pong like the solver
    }
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms;
    cudaEventElapsedTime(&ms, e0, e1);
    printf("%-58s %7.1f us/launch  %6.0f GB/s  (%s)\n", name, ms * 1e3 / reps, useful_bytes / (ms / reps * 1e-3) / 1e9,
           cudaGetErrorString(cudaGetLastError()));
  };
  // bytes moved: read + write of what the lanes touch
  const double rows_total = (double)warps * band_rows;
  // pattern 0: row-major, pitch 1028 floats; a warp's segment = 512 B at column 120*strip (480 B apart)
  run("row-major pitch 1028, 512B segments, ahead 4", k_copy<4>, 30, pitch4, band_rows, band_rows * pitch4, 1028 * pitch4, 32, 0, rows_total * 512 * 2);
  run("row-major pitch 1028, 480B segments, ahead 4", k_copy<4>, 30, pitch4, band_rows, band_rows * pitch4, 1028 * pitch4, 30, 0, rows_total * 480 * 2);
  run("row-major pitch 1056 (128B rows), 480B segs, ahead 4", k_copy<4>, 30, 1056 / 4, band_rows, band_rows * (1056 / 4), 1028 * (1056 / 4), 30, 0, rows_total * 480 * 2);
  run("row-major pitch 1028, 480B segments, ahead 8", k_copy<8>, 30, pitch4, band_rows, band_rows * pitch4, 1028 * pitch4, 30, 0, rows_total * 480 * 2);
  run("row-major pitch 1028, 480B segments, ahead 0", k_copy<0>, 30, pitch4, band_rows, band_rows * pitch4, 1028 * pitch4, 30, 0, rows_total * 480 * 2);
  // pattern 1: strip-major: warp-contiguous streams
  run("strip-major 480B rows contiguous, ahead 4", k_copy<4>, (long)band_rows * 30, 30, band_rows, 0, 0, 30, 1, rows_total * 480 * 2);
  run("strip-major 512B rows contiguous, ahead 4", k_copy<4>, (long)band_rows * 32, 32, band_rows, 0, 0, 32, 1, rows_total * 512 * 2);
  run("strip-major 480B rows contiguous, ahead 8", k_copy<8>, (long)band_rows * 30, 30, band_rows, 0, 0, 30, 1, rows_total * 480 * 2);
  run("strip-major 480B rows contiguous, ahead 16", k_copy<16>, (long)band_rows * 30, 30, band_rows, 0, 0, 30, 1, rows_total * 480 * 2);
  run("strip-major 480B rows contiguous, ahead 0", k_copy<0>, (long)band_rows * 30, 30, band_rows, 0, 0, 30, 1, rows_total * 480 * 2);
  return 0;
}
