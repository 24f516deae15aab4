// f32x2.cuh -- packed float32 x 2 arithmetic (sm_100: fma.rn.f32x2 / mul.rn.f32x2 / add.rn.f32x2, SASS FFMA2 / FMUL2 / FADD2):
// one instruction per PAIR of values, each half rounded exactly like the scalar instruction.  Used by the FX kernels
// (lo = left channel, hi = right channel) and by the TCN epilogue (two adjacent output channels).
#pragma once

namespace mst {
namespace f2 {

typedef unsigned long long u64;

__device__ __forceinline__ u64 pk(float lo, float hi) {
  u64 r;
  asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
  return r;
}
__device__ __forceinline__ float lo_of(u64 v) {
  float a, b;
  asm("mov.b64 {%0, %1}, %2;" : "=f"(a), "=f"(b) : "l"(v));
  (void)b;
  return a;
}
__device__ __forceinline__ float hi_of(u64 v) {
  float a, b;
  asm("mov.b64 {%0, %1}, %2;" : "=f"(a), "=f"(b) : "l"(v));
  (void)a;
  return b;
}
__device__ __forceinline__ u64 dup(float v) { return pk(v, v); }
__device__ __forceinline__ u64 fma2(u64 a, u64 b, u64 c) {
  u64 d;
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c));
  return d;
}
__device__ __forceinline__ u64 mul2(u64 a, u64 b) {
  u64 d;
  asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
  return d;
}
__device__ __forceinline__ u64 add2(u64 a, u64 b) {
  u64 d;
  asm("add.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
  return d;
}

}  // namespace f2
}  // namespace mst
