// ubench_fma.cu -- measures fp32 FMA issue rates on sm_100a: scalar FFMA (register and
// constant-bank weight operand) vs packed FFMA2 (fma.rn.f32x2), alone and mixed with LDS.64.
// Design input for npde_stream.cu (DESIGN.md "Why packed pairs").  Not part of the library.
#include <cstdio>
#include <cuda_runtime.h>

__device__ __forceinline__ float2 ffma2(float2 a, float2 b, float2 c) {
  unsigned long long ra = *reinterpret_cast<unsigned long long *>(&a), rb = *reinterpret_cast<unsigned long long *>(&b),
                     rc = *reinterpret_cast<unsigned long long *>(&c), rd;
  asm volatile("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(rd) : "l"(ra), "l"(rb), "l"(rc));
  return *reinterpret_cast<float2 *>(&rd);
}

constexpr int ITERS = 4096, NACC = 8;

__global__ void k_ffma_reg(float *out, float w) {
  float a[NACC];
  float x = threadIdx.x * 1e-3f, ww = w + threadIdx.x * 1e-9f;
#pragma unroll
  for (int i = 0; i < NACC; ++i) a[i] = i;
  for (int it = 0; it < ITERS; ++it) {
#pragma unroll
    for (int i = 0; i < NACC; ++i) a[i] = __fmaf_rn(ww, x, a[i]);
#pragma unroll
    for (int i = 0; i < NACC; ++i) a[i] = __fmaf_rn(ww, a[(i + 1) % NACC], a[i]);
  }
  float s = 0;
#pragma unroll
  for (int i = 0; i < NACC; ++i) s += a[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
__global__ void k_ffma_const(float *out, float w) {  // w stays in the constant bank
  float a[NACC];
  float x = threadIdx.x * 1e-3f;
#pragma unroll
  for (int i = 0; i < NACC; ++i) a[i] = i;
  for (int it = 0; it < ITERS; ++it) {
#pragma unroll
    for (int i = 0; i < NACC; ++i) a[i] = __fmaf_rn(w, x, a[i]);
#pragma unroll
    for (int i = 0; i < NACC; ++i) a[i] = __fmaf_rn(w, a[(i + 1) % NACC], a[i]);
  }
  float s = 0;
#pragma unroll
  for (int i = 0; i < NACC; ++i) s += a[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
__global__ void k_ffma2(float *out, float w) {
  float2 a[NACC];
  float2 x = make_float2(threadIdx.x * 1e-3f, threadIdx.x * 2e-3f), ww = make_float2(w, w);
#pragma unroll
  for (int i = 0; i < NACC; ++i) a[i] = make_float2(i, -i);
  for (int it = 0; it < ITERS; ++it) {
#pragma unroll
    for (int i = 0; i < NACC; ++i) a[i] = ffma2(ww, x, a[i]);
#pragma unroll
    for (int i = 0; i < NACC; ++i) a[i] = ffma2(ww, a[(i + 1) % NACC], a[i]);
  }
  float s = 0;
#pragma unroll
  for (int i = 0; i < NACC; ++i) s += a[i].x + a[i].y;
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
// FFMA2 + one LDS.64 per 4 FFMA2 (the stencil's expected mix)
__global__ void k_ffma2_lds(float *out, float w) {
  __shared__ float2 sm[1024 + 64];
  for (int i = threadIdx.x; i < 1024 + 64; i += blockDim.x) sm[i] = make_float2(i * 1e-4f, i * 2e-4f);
  __syncthreads();
  float2 a[NACC];
  float2 ww = make_float2(w, w);
#pragma unroll
  for (int i = 0; i < NACC; ++i) a[i] = make_float2(i, -i);
  int p = threadIdx.x;
  for (int it = 0; it < ITERS; ++it) {
    float2 x0 = sm[p], x1 = sm[p + 32];
    float2 x2 = sm[p + 1], x3 = sm[p + 33];
    p = (p + 2) & 1023;
#pragma unroll
    for (int i = 0; i < NACC; ++i) a[i] = ffma2(ww, (i & 1) ? x1 : x0, a[i]);
#pragma unroll
    for (int i = 0; i < NACC; ++i) a[i] = ffma2(ww, (i & 1) ? x3 : x2, a[i]);
  }
  float s = 0;
#pragma unroll
  for (int i = 0; i < NACC; ++i) s += a[i].x + a[i].y;
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
__global__ void k_ffma_lds(float *out, float w) {
  __shared__ float sm[1024 + 64];
  for (int i = threadIdx.x; i < 1024 + 64; i += blockDim.x) sm[i] = i * 1e-4f;
  __syncthreads();
  float a[NACC];
#pragma unroll
  for (int i = 0; i < NACC; ++i) a[i] = i;
  int p = threadIdx.x;
  for (int it = 0; it < ITERS; ++it) {
    float x0 = sm[p], x1 = sm[p + 32], x2 = sm[p + 1], x3 = sm[p + 33];
    p = (p + 2) & 1023;
#pragma unroll
    for (int i = 0; i < NACC; ++i) a[i] = __fmaf_rn(w, (i & 1) ? x1 : x0, a[i]);
#pragma unroll
    for (int i = 0; i < NACC; ++i) a[i] = __fmaf_rn(w, (i & 1) ? x3 : x2, a[i]);
  }
  float s = 0;
#pragma unroll
  for (int i = 0; i < NACC; ++i) s += a[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <class K> void run(const char *name, K kern, int fma_per_instr, float *out, int threads) {
  int dev = 0, sms = 0, khz = 0;
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, dev);
  int blocks = sms * (2048 / threads);
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  for (int i = 0; i < 3; ++i) kern<<<blocks, threads>>>(out, 1.0001f);
  cudaEventRecord(e0);
  for (int i = 0; i < 5; ++i) kern<<<blocks, threads>>>(out, 1.0001f);
  cudaEventRecord(e1);
  cudaEventSynchronize(e1);
  float ms = 0;
  cudaEventElapsedTime(&ms, e0, e1);
  ms /= 5;
  double fmas = (double)blocks * threads * ITERS * NACC * 2 * fma_per_instr;
  double per_clk_sm = fmas / (ms * 1e-3) / sms / (khz * 1e3);
  printf("%-14s threads=%4d  %.3f ms  %.2f Tfma/s  %.1f fma/clk/SM (at max clock %d MHz)  err=%s\n", name, threads, ms,
         fmas / ms * 1e-9, per_clk_sm, khz / 1000, cudaGetErrorString(cudaGetLastError()));
}

int main() {
  float *out;
  cudaMalloc(&out, sizeof(float) * 148 * 2048 * 4);
  for (int threads : {256, 1024}) {
    run("ffma_reg", k_ffma_reg, 1, out, threads);
    run("ffma_const", k_ffma_const, 1, out, threads);
    run("ffma2", k_ffma2, 2, out, threads);
    run("ffma_lds", k_ffma_lds, 1, out, threads);
    run("ffma2_lds", k_ffma2_lds, 2, out, threads);
  }
  return 0;
}
