// cuda_emu.h -- a tiny single-threaded SIMT emulator (TEST INFRASTRUCTURE ONLY).
//
// The build container has nvcc but no GPU.  To debug index logic before
// spending GPU-box minutes, the product's .cu sources can be compiled with g++
// against this header (-DNPDE_EMU) into tests/emu/_build/libnpde_emu.so, which
// exports the same C ABI (include/npde.h) but takes HOST pointers.  Every CUDA
// thread of a block is a ucontext fiber; __syncthreads / __syncwarp / shuffles
// are cooperative barriers; blocks run one after another.  It is slow and only
// used by `tests/test_emu_*.py` on tiny grids.  The product package never loads
// this library: neural-pde-solver_b200/_lib.py only ever opens libnpde_b200.so.
#pragma once
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <ucontext.h>

#include <functional>
#include <vector>

#define __global__
#define __device__
#define __host__
#define __forceinline__ inline __attribute__((always_inline))
#define __launch_bounds__(...)
#define __noinline__ __attribute__((noinline))
#define __shared__ static
#define __constant__ static
#define __align__(n) __attribute__((aligned(n)))

struct uint3 { unsigned x, y, z; };
struct dim3 {
  unsigned x, y, z;
  dim3(unsigned x_ = 1, unsigned y_ = 1, unsigned z_ = 1) : x(x_), y(y_), z(z_) {}
};
struct __attribute__((aligned(16))) float4 { float x, y, z, w; };
struct __attribute__((aligned(8))) float2 { float x, y; };
struct __attribute__((aligned(16))) int4 { int x, y, z, w; };
struct __attribute__((aligned(16))) uint4 { unsigned x, y, z, w; };
static inline float4 make_float4(float x, float y, float z, float w) { return float4{x, y, z, w}; }
static inline float2 make_float2(float x, float y) { return float2{x, y}; }
static inline uint4 make_uint4(unsigned x, unsigned y, unsigned z, unsigned w) { return uint4{x, y, z, w}; }

typedef int cudaError_t;
typedef void *cudaStream_t;
enum { cudaSuccess = 0, cudaErrorInvalidValue = 1 };
enum cudaMemcpyKind { cudaMemcpyHostToDevice = 1, cudaMemcpyDeviceToHost = 2, cudaMemcpyDeviceToDevice = 3, cudaMemcpyDefault = 4 };
enum cudaFuncAttribute { cudaFuncAttributeMaxDynamicSharedMemorySize = 8 };
static inline cudaError_t cudaGetLastError() { return cudaSuccess; }
static inline cudaError_t cudaPeekAtLastError() { return cudaSuccess; }
static inline const char *cudaGetErrorString(cudaError_t) { return "emu"; }
static inline cudaError_t cudaMemsetAsync(void *p, int v, size_t n, cudaStream_t) { memset(p, v, n); return cudaSuccess; }
static inline cudaError_t cudaMemcpyAsync(void *d, const void *s, size_t n, cudaMemcpyKind, cudaStream_t) { memmove(d, s, n); return cudaSuccess; }
template <class F> static inline cudaError_t cudaFuncSetAttribute(F, cudaFuncAttribute, int) { return cudaSuccess; }

namespace emu {
struct Barrier { int count = 0; unsigned gen = 0; };
struct Fiber {
  ucontext_t ctx;
  char *stack = nullptr;
  uint3 tid;
  int lane = 0, warp = 0;
  bool done = false;
};
struct Warp {
  Barrier bar;
  int live = 0;
  uint64_t slot[32];
};
struct State {
  ucontext_t sched;
  std::vector<Fiber> fibers;
  std::vector<Warp> warps;
  Barrier block_bar;
  int live = 0;
  Fiber *cur = nullptr;
  uint3 blk;
  dim3 bdim, gdim;
  char *dyn_smem = nullptr;
  std::function<void()> *body = nullptr;
};
inline State &S() { static State s; return s; }
static const size_t kStack = 256 * 1024;

inline void yield() { State &s = S(); swapcontext(&s.cur->ctx, &s.sched); }
inline void release(Barrier &b) { b.count = 0; b.gen++; }
inline void barrier_wait(Barrier &b, int &live) {
  unsigned g = b.gen;
  if (++b.count >= live) release(b);
  else while (b.gen == g) yield();
}
inline void trampoline() {
  State &s = S();
  (*s.body)();
  Fiber *f = s.cur;
  f->done = true;
  // an exited thread counts as arrived at every later barrier
  s.live--;
  if (s.block_bar.count > 0 && s.block_bar.count >= s.live) release(s.block_bar);
  Warp &w = s.warps[f->warp];
  w.live--;
  if (w.bar.count > 0 && w.bar.count >= w.live) release(w.bar);
  swapcontext(&f->ctx, &s.sched);
}

// NPDE_EMU_TRACE=<file>: append the name of every launched kernel (tests assert on it)
inline void trace(const char *kernel) {
  const char *path = getenv("NPDE_EMU_TRACE");
  if (!path) return;
  if (FILE *f = fopen(path, "a")) { fprintf(f, "%s\n", kernel); fclose(f); }
}

template <class F> inline void launch(dim3 grid, dim3 block, size_t smem, F &&fn) {
  State &s = S();
  int nthreads = block.x * block.y * block.z;
  int nwarps = (nthreads + 31) / 32;
  std::function<void()> body = fn;
  s.body = &body;
  s.bdim = block;
  s.gdim = grid;
  s.fibers.assign(nthreads, Fiber());
  for (auto &f : s.fibers) f.stack = (char *)malloc(kStack);
  std::vector<char> smem_buf(smem + 16);
  for (unsigned bz = 0; bz < grid.z; ++bz)
    for (unsigned by = 0; by < grid.y; ++by)
      for (unsigned bx = 0; bx < grid.x; ++bx) {
        s.blk = uint3{bx, by, bz};
        memset(smem_buf.data(), 0xFF, smem_buf.size());  // NaN-fill: catch uninitialised reads
        s.dyn_smem = (char *)(((uintptr_t)smem_buf.data() + 15) & ~(uintptr_t)15);
        s.warps.assign(nwarps, Warp());
        s.block_bar = Barrier();
        s.live = nthreads;
        for (int t = 0; t < nthreads; ++t) {
          Fiber &f = s.fibers[t];
          f.done = false;
          f.tid = uint3{t % block.x, (t / block.x) % block.y, t / (block.x * block.y)};
          f.lane = t % 32;
          f.warp = t / 32;
          s.warps[f.warp].live++;
          getcontext(&f.ctx);
          f.ctx.uc_stack.ss_sp = f.stack;
          f.ctx.uc_stack.ss_size = kStack;
          f.ctx.uc_link = &s.sched;
          makecontext(&f.ctx, (void (*)())trampoline, 0);
        }
        int remaining = nthreads;
        long spins = 0;
        while (remaining > 0) {
          remaining = 0;
          for (int t = 0; t < nthreads; ++t) {
            Fiber &f = s.fibers[t];
            if (f.done) continue;
            s.cur = &f;
            swapcontext(&s.sched, &f.ctx);
            if (!f.done) remaining++;
          }
          if (++spins > 50000000L) { fprintf(stderr, "emu: deadlock suspected\n"); abort(); }
        }
      }
  for (auto &f : s.fibers) free(f.stack);
  s.fibers.clear();
  s.body = nullptr;
}
}  // namespace emu

#define threadIdx (emu::S().cur->tid)
#define blockIdx (emu::S().blk)
#define blockDim (emu::S().bdim)
#define gridDim (emu::S().gdim)

static inline void __syncthreads() { emu::State &s = emu::S(); emu::barrier_wait(s.block_bar, s.live); }
static inline void __syncwarp(unsigned = 0xffffffffu) {
  emu::State &s = emu::S();
  emu::Warp &w = s.warps[s.cur->warp];
  emu::barrier_wait(w.bar, w.live);
}
static inline void __threadfence() {}
static inline void __threadfence_block() {}

namespace emu {
template <class T> inline T shfl_from(T v, int src, bool valid) {
  State &s = S();
  Warp &w = s.warps[s.cur->warp];
  uint64_t bits = 0;
  memcpy(&bits, &v, sizeof(T));
  w.slot[s.cur->lane] = bits;
  barrier_wait(w.bar, w.live);
  T r = v;
  if (valid) memcpy(&r, &w.slot[src], sizeof(T));
  barrier_wait(w.bar, w.live);
  return r;
}
inline int lane() { return S().cur->lane; }
}  // namespace emu
template <class T> static inline T __shfl_sync(unsigned, T v, int src, int width = 32) {
  int base = emu::lane() / width * width;
  return emu::shfl_from(v, base + (src % width), true);
}
template <class T> static inline T __shfl_up_sync(unsigned, T v, unsigned d, int width = 32) {
  int l = emu::lane();
  return emu::shfl_from(v, l - (int)d, (l % width) >= (int)d);
}
template <class T> static inline T __shfl_down_sync(unsigned, T v, unsigned d, int width = 32) {
  int l = emu::lane();
  return emu::shfl_from(v, l + (int)d, (l % width) + (int)d < width);
}
template <class T> static inline T __shfl_xor_sync(unsigned, T v, int m, int width = 32) {
  int l = emu::lane();
  (void)width;
  return emu::shfl_from(v, l ^ m, true);
}

// atomics (single OS thread: plain read-modify-write)
template <class T> static inline T atomicAdd(T *p, T v) { T o = *p; *p = o + v; return o; }
template <class T> static inline T atomicMax(T *p, T v) { T o = *p; if (v > o) *p = v; return o; }
template <class T> static inline T atomicMin(T *p, T v) { T o = *p; if (v < o) *p = v; return o; }
static inline unsigned atomicInc(unsigned *p, unsigned lim) { unsigned o = *p; *p = (o >= lim) ? 0 : o + 1; return o; }

// intrinsics (compile with -ffp-contract=off so nothing is fused behind our back)
static inline float __fmaf_rn(float a, float b, float c) { return fmaf(a, b, c); }
static inline float __fadd_rn(float a, float b) { volatile float r = a + b; return r; }
static inline float __fsub_rn(float a, float b) { volatile float r = a - b; return r; }
static inline float __fdiv_rn(float a, float b) { volatile float r = a / b; return r; }
static inline float __fmul_rn(float a, float b) { volatile float r = a * b; return r; }
static inline float __fdividef(float a, float b) { return a / b; }
static inline unsigned __float_as_uint(float f) { unsigned u; memcpy(&u, &f, 4); return u; }
static inline float __uint_as_float(unsigned u) { float f; memcpy(&f, &u, 4); return f; }
static inline int __float_as_int(float f) { int u; memcpy(&u, &f, 4); return u; }
static inline float __int_as_float(int u) { float f; memcpy(&f, &u, 4); return f; }
template <class T> static inline T __ldg(const T *p) { return *p; }
static inline int min(int a, int b) { return a < b ? a : b; }
static inline int max(int a, int b) { return a > b ? a : b; }
static inline double __longlong_as_double(long long v) { double d; memcpy(&d, &v, 8); return d; }
