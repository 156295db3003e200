// ubench_stream3.cu -- DRAM page locality of the L2 prefetch: same copy as ubench_stream2.cu mode 0
// (row-major padded images, 9 strips x 4 bands x 64 images = 2304 one-warp CTAs, LDG one row
// ahead), but the L2 prefetch is issued by ONE warp per (image, band) for the whole image row
// block: cp.async.bulk.prefetch.L2 of C consecutive rows (C * 4112 bytes, contiguous) every C rows,
// A rows ahead of consumption, instead of 480 bytes per warp and row.
// Design input for DESIGN.md; not part of the library.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ void prefetch_l2(const void *p) { asm volatile("prefetch.global.L2 [%0];" ::"l"(p)); }
__device__ __forceinline__ void bulk_prefetch_l2(const void *p, uint32_t bytes) {
  asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(p), "r"(bytes) : "memory");
}
struct Geo {
  long warp_stride, row_stride, band_stride, img_stride;  // float4 units
  int rows, strips, bands, lanes;
};

// MODE 0: per-warp per-row prefetch (AHEAD rows); MODE 1: strip-0 warp prefetches C rows x full row every C rows
template <int MODE, int AHEAD, int C>
__global__ void __launch_bounds__(32, 16) k_copy(const float4 *__restrict__ in, float4 *__restrict__ out, Geo g) {
  const int lane = threadIdx.x;
  const int strip = blockIdx.x % g.strips, band = (blockIdx.x / g.strips) % g.bands, b = blockIdx.x / (g.strips * g.bands);
  const long band_base = (long)b * g.img_stride + (long)band * g.band_stride;
  const float4 *pi = in + band_base + (long)strip * g.warp_stride + lane;
  float4 *po = out + band_base + (long)strip * g.warp_stride + lane;
  const float4 *prow = in + band_base;  // start of the band's first row
  const bool on = lane < g.lanes;
  const uint32_t chunk_bytes = (uint32_t)(C * g.row_stride * 16);
  if (MODE == 1 && strip == 0 && lane == 0) {
    for (int r = 0; r < AHEAD && r < g.rows; r += C) bulk_prefetch_l2(prow + (long)r * g.row_stride, chunk_bytes);
  }
  float4 v = on ? __ldg(pi) : make_float4(0, 0, 0, 0);
  for (int r = 0; r < g.rows; ++r) {
    float4 nv = v;
    if (r + 1 < g.rows && on) nv = __ldg(pi + g.row_stride);
    if (MODE == 0) {
      if (AHEAD > 0 && r + 1 + AHEAD < g.rows && on) prefetch_l2(pi + (1 + AHEAD) * g.row_stride);
    } else {
      if (strip == 0 && lane == 0 && (r % C) == 0 && r + AHEAD + C <= g.rows)
        bulk_prefetch_l2(prow + (long)(r + AHEAD) * g.row_stride, chunk_bytes);
    }
    v.x += 1.0f;
    if (on) *po = v;
    v = nv;
    pi += g.row_stride;
    po += g.row_stride;
  }
}

int main() {
  const int B = 64, strips = 9, bands = 4, band_rows = 257;
  const long pitch4 = 1028 / 4;
  size_t bytes = (size_t)B * 1028 * 1028 * 4 + (1 << 20);
  float4 *in, *out;
  cudaMalloc(&in, bytes);
  cudaMalloc(&out, bytes);
  cudaMemset(in, 0, bytes);
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  const int warps = B * strips * bands;
  Geo g{30, pitch4, band_rows * pitch4, 1028 * pitch4, band_rows, strips, bands, 30};
  const double useful = (double)warps * band_rows * 480 * 2;
  auto run = [&](const char *name, auto kern) {
    for (int i = 0; i < 3; ++i) kern<<<warps, 32>>>(in, out, g);
    cudaEventRecord(e0);
    const int reps = 20;
    for (int i = 0; i < reps; ++i) {
      kern<<<warps, 32>>>(in, out, g);
      float4 *t = in; in = out; out = t;
    }
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms;
    cudaEventElapsedTime(&ms, e0, e1);
    printf("%-52s %7.1f us/launch  %6.0f GB/s  (%s)\n", name, ms * 1e3 / reps, useful / (ms / reps * 1e-3) / 1e9,
           cudaGetErrorString(cudaGetLastError()));
  };
  run("per-warp row prefetch, 4 ahead", k_copy<0, 4, 1>);
  run("group prefetch C=1 row,  A=4", k_copy<1, 4, 1>);
  run("group prefetch C=2 rows, A=4", k_copy<1, 4, 2>);
  run("group prefetch C=4 rows, A=4", k_copy<1, 4, 4>);
  run("group prefetch C=4 rows, A=8", k_copy<1, 8, 4>);
  run("group prefetch C=8 rows, A=8", k_copy<1, 8, 8>);
  run("group prefetch C=8 rows, A=16", k_copy<1, 16, 8>);
  run("group prefetch C=16 rows, A=16", k_copy<1, 16, 16>);
  run("group prefetch C=16 rows, A=32", k_copy<1, 32, 16>);
  run("group prefetch C=32 rows, A=32", k_copy<1, 32, 32>);
  return 0;
}
