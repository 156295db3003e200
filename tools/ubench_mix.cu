// ubench_mix.cu -- issue-slot / FMA-pipe model of sm_100a for the mix npde_stream.cu executes:
// packed FFMA2, scalar FFMA, integer ALU and SHFL, alone and interleaved.  Design input for
// DESIGN.md ("what bounds the Conv3 kernel"); not part of the library.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/_build/ubench_mix tools/ubench_mix.cu
#include <cstdio>
#include <cuda_runtime.h>
typedef unsigned long long u64;
__device__ __forceinline__ u64 ffma2(u64 a, u64 b, u64 c) {
  u64 d;
  asm volatile("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c));
  return d;
}
__device__ __forceinline__ float ffma(float a, float b, float c) {
  float d;
  asm volatile("fma.rn.f32 %0, %1, %2, %3;" : "=f"(d) : "f"(a), "f"(b), "f"(c));
  return d;
}
__device__ __forceinline__ unsigned lop(unsigned a, unsigned b) {
  unsigned d;
  asm volatile("xor.b32 %0, %1, %2;" : "=r"(d) : "r"(a), "r"(b));
  return d;
}
constexpr int ITERS = 2048;
// MODE: n2 packed, n1 scalar, na alu, ns shfl per unrolled group of 8 accumulator updates
template <int N2, int N1, int NA, int NS>
__global__ void k_mix(float *out, float w, long long *cyc) {
  u64 a2[8];
  float a1[8];
  unsigned ai[8];
  float sh[4];
  const u64 ww = ((u64)__float_as_uint(w) << 32) | __float_as_uint(w);
  const u64 x2 = ((u64)__float_as_uint(threadIdx.x * 1e-3f) << 32) | __float_as_uint(threadIdx.x * 2e-3f);
  const float x1 = threadIdx.x * 1e-3f;
#pragma unroll
  for (int i = 0; i < 8; ++i) { a2[i] = i; a1[i] = i; ai[i] = i + threadIdx.x; }
#pragma unroll
  for (int i = 0; i < 4; ++i) sh[i] = threadIdx.x + i;
  long long t0 = clock64();
  for (int it = 0; it < ITERS; ++it) {
#pragma unroll
    for (int r = 0; r < 4; ++r) {
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        if (i < N2) a2[i] = ffma2(ww, x2, a2[i]);
        if (i < N1) a1[i] = ffma(w, x1, a1[i]);
        if (i < NA) ai[i] = lop(ai[i], 0x5a5a5a5au + r);
        if (i < NS && r == 0) sh[i & 3] = __shfl_sync(0xffffffffu, sh[i & 3], (threadIdx.x + 1) & 31);
      }
    }
  }
  long long t1 = clock64();
  float s = 0;
#pragma unroll
  for (int i = 0; i < 8; ++i) s += __uint_as_float((unsigned)a2[i]) + __uint_as_float((unsigned)(a2[i] >> 32)) + a1[i] + ai[i];
#pragma unroll
  for (int i = 0; i < 4; ++i) s += sh[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
}
template <int N2, int N1, int NA, int NS>
void run(float *out, long long *cyc, int warps_per_smsp) {
  int threads = 128 * warps_per_smsp;  // warps are distributed round-robin over the 4 SMSPs
  for (int i = 0; i < 2; ++i) k_mix<N2, N1, NA, NS><<<148, threads>>>(out, 1.0001f, cyc);
  cudaDeviceSynchronize();
  long long c = 0;
  cudaMemcpy(&c, cyc, sizeof(c), cudaMemcpyDeviceToHost);
  double groups = (double)ITERS * 4;  // per warp
  double per_group = (double)c / groups / warps_per_smsp;  // SMSP cycles per group of one warp
  printf("FFMA2 x%d + FFMA x%d + LOP x%d + SHFL x%.2f  warps/SMSP=%d : %.2f cyc per group  (pipe model 2*N2+N1 = %d, issue = %.2f)  err=%s\n",
         N2, N1, NA, NS / 4.0, warps_per_smsp, per_group, 2 * N2 + N1, N2 + N1 + NA + NS / 4.0, cudaGetErrorString(cudaGetLastError()));
}
int main() {
  float *out; long long *cyc;
  cudaMalloc(&out, 148 * 1024 * sizeof(float));
  cudaMalloc(&cyc, 8);
  for (int w = 1; w <= 4; w *= 2) {
    run<8, 0, 0, 0>(out, cyc, w);
    run<0, 8, 0, 0>(out, cyc, w);
    run<4, 4, 0, 0>(out, cyc, w);
    run<8, 8, 0, 0>(out, cyc, w);
    run<8, 0, 8, 0>(out, cyc, w);
    run<8, 0, 4, 0>(out, cyc, w);
    run<4, 4, 4, 0>(out, cyc, w);
    run<4, 4, 4, 4>(out, cyc, w);
    run<4, 4, 2, 8>(out, cyc, w);
    run<0, 8, 8, 0>(out, cyc, w);
    printf("\n");
  }
  return 0;
}
