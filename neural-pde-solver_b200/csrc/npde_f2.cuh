// npde_f2.cuh -- packed-fp32 register pairs and the asynchronous-copy primitives shared by the
// register-streaming kernels (npde_stream.cu, npde_wide.cu).  Everything lives in an anonymous
// namespace: each translation unit gets its own inlined copies.
#pragma once
#include "npde_common.cuh"

namespace {

// f2 = two fp32 values in one 64-bit register pair.  On the device it is an opaque 64-bit
// integer that is packed / unpacked explicitly (mov.b64), so ptxas sees exactly one register
// pair per value and every packed instruction takes it as is; in the emulator it is a struct.
// NPDE_SCALAR_F2 (tuning builds): the struct representation on the device too, i.e. scalar FFMA / FMUL only.
#ifdef NPDE_EMU
#define NPDE_GRID_CONSTANT
#else
#define NPDE_GRID_CONSTANT __grid_constant__
#endif
#if !defined(NPDE_EMU) && defined(NPDE_F2_HALVES)
// Tuning build: a pair is carried as two 32-bit values and packed only inside the packed instruction, so that every
// loop-carried value is a 32-bit register for ptxas (which coalesces those across the back edge without copies).
struct __align__(8) f2 { float x, y; };
__device__ __forceinline__ f2 mk2(float lo, float hi) { f2 r; r.x = lo; r.y = hi; return r; }
__device__ __forceinline__ float lo(f2 a) { return a.x; }
__device__ __forceinline__ float hi(f2 a) { return a.y; }
__device__ __forceinline__ f2 fma2(f2 a, f2 b, f2 c) {
  f2 d;
  asm("{\n.reg .b64 ra, rb, rc, rd;\nmov.b64 ra, {%2, %3};\nmov.b64 rb, {%4, %5};\nmov.b64 rc, {%6, %7};\n"
      "fma.rn.f32x2 rd, ra, rb, rc;\nmov.b64 {%0, %1}, rd;\n}"
      : "=f"(d.x), "=f"(d.y) : "f"(a.x), "f"(a.y), "f"(b.x), "f"(b.y), "f"(c.x), "f"(c.y));
  return d;
}
__device__ __forceinline__ f2 mul2(f2 a, f2 b) {
  f2 d;
  asm("{\n.reg .b64 ra, rb, rd;\nmov.b64 ra, {%2, %3};\nmov.b64 rb, {%4, %5};\n"
      "mul.rn.f32x2 rd, ra, rb;\nmov.b64 {%0, %1}, rd;\n}"
      : "=f"(d.x), "=f"(d.y) : "f"(a.x), "f"(a.y), "f"(b.x), "f"(b.y));
  return d;
}
__device__ __forceinline__ f2 sub2(f2 a, f2 b) {
  f2 d;
  asm("{\n.reg .b64 ra, rb, rd;\nmov.b64 ra, {%2, %3};\nmov.b64 rb, {%4, %5};\n"
      "sub.rn.f32x2 rd, ra, rb;\nmov.b64 {%0, %1}, rd;\n}"
      : "=f"(d.x), "=f"(d.y) : "f"(a.x), "f"(a.y), "f"(b.x), "f"(b.y));
  return d;
}
__device__ __forceinline__ f2 and2(f2 a, unsigned mlo, unsigned mhi) {
  return mk2(__uint_as_float(__float_as_uint(a.x) & mlo), __uint_as_float(__float_as_uint(a.y) & mhi));
}
#elif defined(NPDE_EMU) || defined(NPDE_SCALAR_F2)
struct __align__(8) f2 { float x, y; };
__device__ __forceinline__ f2 mk2(float lo, float hi) { f2 r; r.x = lo; r.y = hi; return r; }
__device__ __forceinline__ float lo(f2 a) { return a.x; }
__device__ __forceinline__ float hi(f2 a) { return a.y; }
__device__ __forceinline__ f2 fma2(f2 a, f2 b, f2 c) { return mk2(__fmaf_rn(a.x, b.x, c.x), __fmaf_rn(a.y, b.y, c.y)); }
__device__ __forceinline__ f2 mul2(f2 a, f2 b) { return mk2(__fmul_rn(a.x, b.x), __fmul_rn(a.y, b.y)); }
__device__ __forceinline__ f2 sub2(f2 a, f2 b) { return mk2(__fsub_rn(a.x, b.x), __fsub_rn(a.y, b.y)); }
// bitwise AND with a pair of 32-bit masks
__device__ __forceinline__ f2 and2(f2 a, unsigned mlo, unsigned mhi) {
  return mk2(__uint_as_float(__float_as_uint(a.x) & mlo), __uint_as_float(__float_as_uint(a.y) & mhi));
}
#else
typedef unsigned long long f2;
__device__ __forceinline__ f2 mk2(float lo, float hi) {
  f2 d;
  asm("mov.b64 %0, {%1, %2};" : "=l"(d) : "f"(lo), "f"(hi));
  return d;
}
__device__ __forceinline__ float lo(f2 a) {
  float l, h;
  asm("mov.b64 {%0, %1}, %2;" : "=f"(l), "=f"(h) : "l"(a));
  return l;
}
__device__ __forceinline__ float hi(f2 a) {
  float l, h;
  asm("mov.b64 {%0, %1}, %2;" : "=f"(l), "=f"(h) : "l"(a));
  return h;
}
// Packed fp32 (new on sm_100): two independent round-to-nearest operations per instruction.
// The only products that are later added (first conv tap, 0.25*u) feed an fma's addend, which
// ptxas cannot contract, so the rounding sequence is exactly the scalar fmaf chain's.
__device__ __forceinline__ f2 fma2(f2 a, f2 b, f2 c) {
  f2 d;
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c));
  return d;
}
__device__ __forceinline__ f2 mul2(f2 a, f2 b) {
  f2 d;
  asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
  return d;
}
__device__ __forceinline__ f2 sub2(f2 a, f2 b) {
  f2 d;
  asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
  return d;
}
__device__ __forceinline__ f2 and2(f2 a, unsigned mlo, unsigned mhi) {
  return a & (((f2)mhi << 32) | (f2)mlo);
}
#endif
#ifdef NPDE_EMU
__device__ __forceinline__ void prefetch_l2(const void *) {}
// 16 bytes at p -> two pairs {p[0], p[1]}, {p[2], p[3]}
__device__ __forceinline__ void ld_pairs(const float *p, f2 &a, f2 &b) {
  a = mk2(p[0], p[1]);
  b = mk2(p[2], p[3]);
}
__device__ __forceinline__ void st_pairs(float *p, f2 a, f2 b) {
  p[0] = a.x; p[1] = a.y; p[2] = b.x; p[3] = b.y;
}
// asynchronous 16-byte global -> shared copy and its group bookkeeping: synchronous in the emulator
__device__ __forceinline__ void cp_async16(void *smem, const float *g) { memcpy(smem, g, 16); }
__device__ __forceinline__ void cp_async4_zfill(void *smem, const float *g, bool valid) {
  float v = valid ? *g : 0.0f;
  memcpy(smem, &v, 4);
}
__device__ __forceinline__ void cp_async_commit() {}
template <int PENDING> __device__ __forceinline__ void cp_async_wait() {}
#else
__device__ __forceinline__ void prefetch_l2(const void *p) { asm volatile("prefetch.global.L2 [%0];" ::"l"(p)); }
#if defined(NPDE_SCALAR_F2) || defined(NPDE_F2_HALVES)
__device__ __forceinline__ void ld_pairs(const float *p, f2 &a, f2 &b) {
  const float4 v = __ldg(reinterpret_cast<const float4 *>(p));
  a = mk2(v.x, v.y);
  b = mk2(v.z, v.w);
}
__device__ __forceinline__ void st_pairs(float *p, f2 a, f2 b) {
  *reinterpret_cast<float4 *>(p) = make_float4(a.x, a.y, b.x, b.y);
}
#else
__device__ __forceinline__ void ld_pairs(const float *p, f2 &a, f2 &b) {
  const ulonglong2 v = __ldg(reinterpret_cast<const ulonglong2 *>(p));
  a = v.x;
  b = v.y;
}
__device__ __forceinline__ void st_pairs(float *p, f2 a, f2 b) {
  *reinterpret_cast<ulonglong2 *>(p) = make_ulonglong2(a, b);
}
#endif
// cp.async (LDGSTS): 16 bytes per lane straight from L2 into shared memory, no register staging
__device__ __forceinline__ void cp_async16(void *smem, const float *g) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"((unsigned)__cvta_generic_to_shared(smem)), "l"(g)
               : "memory");
}
// 4-byte variant with zero fill (src-size 0 writes zeros): coalesced element-wise copies of a row that is
// only 4-byte aligned (reference layout)
__device__ __forceinline__ void cp_async4_zfill(void *smem, const float *g, bool valid) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 4, %2;" ::"r"((unsigned)__cvta_generic_to_shared(smem)), "l"(g),
               "r"(valid ? 4 : 0)
               : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int PENDING> __device__ __forceinline__ void cp_async_wait() {
  asm volatile("cp.async.wait_group %0;" ::"n"(PENDING) : "memory");
}
#endif

__device__ __forceinline__ f2 bc2(float s) { return mk2(s, s); }
__device__ __forceinline__ f2 zero2() { return mk2(0.0f, 0.0f); }
// acc + w * {a_lo, a_hi} on the two halves separately (two scalar FFMAs: the operand pair is
// not a register pair)
__device__ __forceinline__ f2 fma_lohi(float w, float a_lo, float a_hi, f2 acc) {
  return mk2(__fmaf_rn(w, a_lo, lo(acc)), __fmaf_rn(w, a_hi, hi(acc)));
}
__device__ __forceinline__ f2 mul_lohi(float w, float a_lo, float a_hi) {
  return mk2(__fmul_rn(w, a_lo), __fmul_rn(w, a_hi));
}

}  // namespace
