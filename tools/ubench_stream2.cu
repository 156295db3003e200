// ubench_stream2.cu -- how to keep enough bytes in flight with 16 one-warp CTAs per SM (the
// register budget of the Conv3 stream kernel): copy of 480/512-byte row segments, row-major
// padded images (pitch 1028 floats) as in ubench_stream.cu, with
//   mode 0: LDG one row ahead + prefetch.global.L2 AHEAD rows ahead        (what npde_stream.cu did)
//   mode 1: cp.async (LDGSTS, 16 B per lane) ring of D rows in shared memory, no L2 prefetch
//   mode 2: cp.async.bulk (TMA 1-D bulk copy) + mbarrier ring of D rows, issued by one lane
// Design input for DESIGN.md; not part of the library.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ void prefetch_l2(const void *p) { asm volatile("prefetch.global.L2 [%0];" ::"l"(p)); }
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

struct Geo {
  long warp_stride, row_stride, band_stride, img_stride;  // float4 units
  int rows, strips, bands, lanes;
};
__device__ __forceinline__ long warp_base(const Geo &g) {
  const int strip = blockIdx.x % g.strips, band = (blockIdx.x / g.strips) % g.bands, b = blockIdx.x / (g.strips * g.bands);
  return (long)b * g.img_stride + (long)band * g.band_stride + (long)strip * g.warp_stride;
}

template <int AHEAD>
__global__ void __launch_bounds__(32, 16) k_ldg(const float4 *__restrict__ in, float4 *__restrict__ out, Geo g) {
  const int lane = threadIdx.x;
  const float4 *pi = in + warp_base(g) + lane;
  float4 *po = out + warp_base(g) + lane;
  const bool on = lane < g.lanes;
  float4 v = on ? __ldg(pi) : make_float4(0, 0, 0, 0);
  for (int r = 0; r < g.rows; ++r) {
    float4 nv = v;
    if (r + 1 < g.rows && on) nv = __ldg(pi + g.row_stride);
    if (AHEAD > 0 && r + 1 + AHEAD < g.rows && on) prefetch_l2(pi + (1 + AHEAD) * g.row_stride);
    v.x += 1.0f;
    if (on) *po = v;
    v = nv;
    pi += g.row_stride;
    po += g.row_stride;
  }
}

template <int D>
__global__ void __launch_bounds__(32, 16) k_cpasync(const float4 *__restrict__ in, float4 *__restrict__ out, Geo g) {
  __shared__ float4 ring[D][32];
  const int lane = threadIdx.x;
  const float4 *pi = in + warp_base(g) + lane;
  float4 *po = out + warp_base(g) + lane;
  const bool on = lane < g.lanes;
  const uint32_t s0 = smem_u32(&ring[0][lane]);
#pragma unroll
  for (int d = 0; d < D - 1; ++d) {
    if (d < g.rows && on) asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(s0 + d * 512), "l"(pi + (long)d * g.row_stride));
    asm volatile("cp.async.commit_group;");
  }
  int slot = 0, fill = D - 1;
  for (int r = 0; r < g.rows; ++r) {
    if (r + D - 1 < g.rows && on)
      asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(s0 + fill * 512), "l"(pi + (long)(D - 1) * g.row_stride));
    asm volatile("cp.async.commit_group;");
    asm volatile("cp.async.wait_group %0;" ::"n"(D - 1));
    float4 v = ring[slot][lane];
    v.x += 1.0f;
    if (on) *po = v;
    pi += g.row_stride;
    po += g.row_stride;
    slot = (slot + 1 == D) ? 0 : slot + 1;
    fill = (fill + 1 == D) ? 0 : fill + 1;
  }
}

template <int D>
__global__ void __launch_bounds__(32, 16) k_bulk(const float4 *__restrict__ in, float4 *__restrict__ out, Geo g) {
  __shared__ __align__(128) float4 ring[D][32];
  __shared__ __align__(8) unsigned long long bar[D];
  const int lane = threadIdx.x;
  const float4 *pi = in + warp_base(g);  // row start of the segment (lane 0)
  float4 *po = out + warp_base(g) + lane;
  const bool on = lane < g.lanes;
  const uint32_t bytes = g.lanes * 16;
  if (lane == 0) {
#pragma unroll
    for (int d = 0; d < D; ++d) asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar[d])));
    asm volatile("fence.mbarrier_init.release.cluster;");
  }
  __syncwarp();
  if (lane == 0) {
#pragma unroll
    for (int d = 0; d < D - 1; ++d)
      if (d < g.rows) {
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(&bar[d])), "r"(bytes));
        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                         smem_u32(&ring[d][0])),
                     "l"(pi + (long)d * g.row_stride), "r"(bytes), "r"(smem_u32(&bar[d]))
                     : "memory");
      }
  }
  int slot = 0, fill = D - 1;
  uint32_t phase = 0;
  for (int r = 0; r < g.rows; ++r) {
    __syncwarp();  // all lanes have read the slot that is refilled now
    if (lane == 0 && r + D - 1 < g.rows) {
      asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(&bar[fill])), "r"(bytes));
      asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                       smem_u32(&ring[fill][0])),
                   "l"(pi + (long)(D - 1) * g.row_stride), "r"(bytes), "r"(smem_u32(&bar[fill]))
                   : "memory");
    }
    uint32_t done = 0;
    while (!done)
      asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0, 1, 0, p; }"
                   : "=r"(done)
                   : "r"(smem_u32(&bar[slot])), "r"(phase)
                   : "memory");
    float4 v = ring[slot][lane];
    v.x += 1.0f;
    if (on) *po = v;
    pi += g.row_stride;
    po += g.row_stride;
    if (++slot == D) { slot = 0; phase ^= 1; }
    if (++fill == D) fill = 0;
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
    printf("%-44s %7.1f us/launch  %6.0f GB/s  (%s)\n", name, ms * 1e3 / reps, useful / (ms / reps * 1e-3) / 1e9,
           cudaGetErrorString(cudaGetLastError()));
  };
  run("LDG +1 row, L2 prefetch 4 rows", k_ldg<4>);
  run("LDG +1 row, L2 prefetch 2 rows", k_ldg<2>);
  run("LDG +1 row, L2 prefetch 6 rows", k_ldg<6>);
  run("cp.async ring D=4", k_cpasync<4>);
  run("cp.async ring D=6", k_cpasync<6>);
  run("cp.async ring D=8", k_cpasync<8>);
  run("cp.async ring D=12", k_cpasync<12>);
  run("cp.async ring D=16", k_cpasync<16>);
  run("bulk+mbarrier ring D=4", k_bulk<4>);
  run("bulk+mbarrier ring D=8", k_bulk<8>);
  run("bulk+mbarrier ring D=12", k_bulk<12>);
  run("bulk+mbarrier ring D=16", k_bulk<16>);
  return 0;
}
