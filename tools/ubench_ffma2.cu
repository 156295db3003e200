// ubench_ffma2.cu -- issue cost of packed fp32 FMAs on sm_100a by operand kind: register pair, uniform register
// broadcast (UR.F32, what ptxas emits for a kernel-parameter weight), immediate.  Design input for
// npde_wide.cu; not part of the library.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/_build/ubench_ffma2 tools/ubench_ffma2.cu
#include <cstdio>
#include <cuda_runtime.h>
typedef unsigned long long u64;
__device__ __forceinline__ u64 ffma2(u64 a, u64 b, u64 c) {
  u64 d;
  asm volatile("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c));
  return d;
}
__device__ __forceinline__ u64 pack(float lo, float hi) {
  u64 d;
  asm("mov.b64 %0, {%1, %2};" : "=l"(d) : "f"(lo), "f"(hi));
  return d;
}
struct P { float w[8]; };
constexpr int ITERS = 4096;
// MODE 0: b operand = per-thread register pairs; 1: uniform (kernel parameter) broadcast; 2: immediate;
// 3: uniform values copied into per-thread register pairs first
template <int MODE, int NACC>
__global__ void k(float *out, const __grid_constant__ P p, long long *cyc) {
  u64 acc[NACC], x[4];
#pragma unroll
  for (int i = 0; i < NACC; ++i) acc[i] = pack(i, i + 1);
#pragma unroll
  for (int i = 0; i < 4; ++i) x[i] = pack(threadIdx.x * 1e-3f + i, threadIdx.x * 2e-3f);
  u64 wr[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    float v = p.w[i];
    if (MODE == 0) v += threadIdx.x * 1e-9f;
    if (MODE == 3) asm volatile("" : "+f"(v));  // opaque: a per-thread register copy of a uniform value
    wr[i] = pack(v, v);
  }
  long long t0 = clock64();
  for (int it = 0; it < ITERS; ++it) {
#pragma unroll
    for (int r = 0; r < 8; ++r) {
#pragma unroll
      for (int i = 0; i < NACC; ++i) {
        u64 b;
        if (MODE == 1) b = pack(p.w[r], p.w[r]);
        else if (MODE == 2) b = pack(0.25f, 0.25f);
        else b = wr[r];
        acc[i] = ffma2(x[(i + r) & 3], b, acc[i]);
      }
    }
  }
  long long t1 = clock64();
  float s = 0;
#pragma unroll
  for (int i = 0; i < NACC; ++i) s += __uint_as_float((unsigned)acc[i]) + __uint_as_float((unsigned)(acc[i] >> 32));
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
}
template <int MODE, int NACC>
void run(float *out, long long *cyc, int warps_per_smsp, const char *what) {
  P p;
  for (int i = 0; i < 8; ++i) p.w[i] = 1.0f + 1e-4f * i;
  for (int i = 0; i < 2; ++i) k<MODE, NACC><<<148, 128 * warps_per_smsp>>>(out, p, cyc);
  cudaDeviceSynchronize();
  long long c = 0;
  cudaMemcpy(&c, cyc, sizeof(c), cudaMemcpyDeviceToHost);
  double n = (double)ITERS * 8 * NACC * warps_per_smsp;  // FFMA2 per scheduler
  printf("%-34s acc=%2d warps/SMSP=%d : %.2f cycles per FFMA2 (scheduler)  %s\n", what, NACC, warps_per_smsp, (double)c / n,
         cudaGetErrorString(cudaGetLastError()));
}
int main() {
  float *out; long long *cyc;
  cudaMalloc(&out, 148 * 1024 * sizeof(float));
  cudaMalloc(&cyc, 8);
  for (int w = 1; w <= 2; ++w) {
    run<0, 8>(out, cyc, w, "b = register pair");
    run<1, 8>(out, cyc, w, "b = uniform broadcast (UR.F32)");
    run<2, 8>(out, cyc, w, "b = immediate");
    run<3, 8>(out, cyc, w, "b = register copy of uniform");
    run<0, 4>(out, cyc, w, "b = register pair");
    run<1, 4>(out, cyc, w, "b = uniform broadcast (UR.F32)");
    run<1, 2>(out, cyc, w, "b = uniform broadcast (UR.F32)");
    run<0, 2>(out, cyc, w, "b = register pair");
    run<1, 1>(out, cyc, w, "b = uniform broadcast (UR.F32)");
    run<0, 1>(out, cyc, w, "b = register pair");
  }
  return 0;
}
