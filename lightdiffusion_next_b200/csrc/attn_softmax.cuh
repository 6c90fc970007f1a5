// Softmax-side helpers shared by the tensor-memory attention kernels (attention5.cu: d = 40, attention6.cu: d = 80).
#pragma once
#include "common.h"
#include "ptx.cuh"

namespace ldn {
namespace asm_sm {

__device__ __forceinline__ float ex2m(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ float ex2p(float x) {
  x = fmaxf(x, -125.0f);
  const float t = x + 12582912.0f;
  const float n = t - 12582912.0f;
  const float f = x - n;
  float p = fmaf(f, 0.0555041086f, 0.2402265070f);
  p = fmaf(p, f, 0.6931471806f);
  p = fmaf(p, f, 1.0f);
  return __int_as_float(__float_as_int(p) + (__float_as_int(t) << 23));
}
// exp2(s * sc + neg_m) for one 8-column chunk of a score row, packed to 4 bf16x2 words.
// mode 0: MUFU ex2; mode 1: degree-3 polynomial on the FMA pipe (same arithmetic as ex2p, two lanes per instruction);
// mode 2: experiment only (no exponential). `mode` is a compile-time constant after unrolling.
__device__ __forceinline__ void exp8_pack(const uint32_t* s, float sc, float neg_m, int mode, uint32_t* w) {
  const float2 sc2 = make_float2(sc, sc), nm2 = make_float2(neg_m, neg_m);
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    float2 e = ffma2(make_float2(__uint_as_float(s[2 * k]), __uint_as_float(s[2 * k + 1])), sc2, nm2);
    if (mode == 1) {
      e.x = fmaxf(e.x, -125.0f);
      e.y = fmaxf(e.y, -125.0f);
      const float2 t = fadd2(e, make_float2(12582912.0f, 12582912.0f));
      const float2 n = fadd2(t, make_float2(-12582912.0f, -12582912.0f));
      const float2 f = ffma2(n, make_float2(-1.0f, -1.0f), e);
      float2 q = ffma2(f, make_float2(0.0555041086f, 0.0555041086f), make_float2(0.2402265070f, 0.2402265070f));
      q = ffma2(q, f, make_float2(0.6931471806f, 0.6931471806f));
      q = ffma2(q, f, make_float2(1.0f, 1.0f));
      e.x = __int_as_float(__float_as_int(q.x) + (__float_as_int(t.x) << 23));
      e.y = __int_as_float(__float_as_int(q.y) + (__float_as_int(t.y) << 23));
    } else if (mode == 0) {
      e.x = ex2m(e.x);
      e.y = ex2m(e.y);
    } else {
      e.x *= 0.001f;
      e.y *= 0.001f;
    }
    w[k] = pack_bf16x2(e.x, e.y);
  }
}

// exp2 of eight scores that already carry scale and offset (attention9.cu, folded variant: the tensor core produced
// s * scale * log2(e) - m through an extra operand column), packed to 4 bf16x2 words.  Same modes as exp8_pack.
__device__ __forceinline__ void exp8_pack_raw(const uint32_t* s, int mode, uint32_t* w) {
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    float2 e = make_float2(__uint_as_float(s[2 * k]), __uint_as_float(s[2 * k + 1]));
    if (mode == 1) {
      e.x = fmaxf(e.x, -125.0f);
      e.y = fmaxf(e.y, -125.0f);
      const float2 t = fadd2(e, make_float2(12582912.0f, 12582912.0f));
      const float2 n = fadd2(t, make_float2(-12582912.0f, -12582912.0f));
      const float2 f = ffma2(n, make_float2(-1.0f, -1.0f), e);
      float2 q = ffma2(f, make_float2(0.0555041086f, 0.0555041086f), make_float2(0.2402265070f, 0.2402265070f));
      q = ffma2(q, f, make_float2(0.6931471806f, 0.6931471806f));
      q = ffma2(q, f, make_float2(1.0f, 1.0f));
      e.x = __int_as_float(__float_as_int(q.x) + (__float_as_int(t.x) << 23));
      e.y = __int_as_float(__float_as_int(q.y) + (__float_as_int(t.y) << 23));
    } else if (mode == 0) {
      e.x = ex2m(e.x);
      e.y = ex2m(e.y);
    } else {
      e.x *= 0.001f;
      e.y *= 0.001f;
    }
    w[k] = pack_bf16x2(e.x, e.y);
  }
}
__device__ __forceinline__ uint32_t pin3(uint32_t v) {
  asm volatile("mov.u32 %0, %0;" : "+r"(v));
  return v;
}
__device__ __forceinline__ void tmem_st16(uint32_t taddr, const uint32_t* v) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16};" ::"r"(taddr),
      "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]), "r"(v[8]), "r"(v[9]),
      "r"(v[10]), "r"(v[11]), "r"(v[12]), "r"(v[13]), "r"(v[14]), "r"(v[15])
      : "memory");
}
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }


// exp8_pack that also accumulates the eight fp32 exponentials into a packed running sum (softmax denominator kept in
// registers when V^T carries no ones row).
__device__ __forceinline__ void exp8_pack_sum(const uint32_t* s, float sc, float neg_m, int mode, uint32_t* w, float2& acc) {
  const float2 sc2 = make_float2(sc, sc), nm2 = make_float2(neg_m, neg_m);
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    float2 e = ffma2(make_float2(__uint_as_float(s[2 * k]), __uint_as_float(s[2 * k + 1])), sc2, nm2);
    if (mode == 1) {
      e.x = fmaxf(e.x, -125.0f);
      e.y = fmaxf(e.y, -125.0f);
      const float2 t = fadd2(e, make_float2(12582912.0f, 12582912.0f));
      const float2 n = fadd2(t, make_float2(-12582912.0f, -12582912.0f));
      const float2 f = ffma2(n, make_float2(-1.0f, -1.0f), e);
      float2 q = ffma2(f, make_float2(0.0555041086f, 0.0555041086f), make_float2(0.2402265070f, 0.2402265070f));
      q = ffma2(q, f, make_float2(0.6931471806f, 0.6931471806f));
      q = ffma2(q, f, make_float2(1.0f, 1.0f));
      e.x = __int_as_float(__float_as_int(q.x) + (__float_as_int(t.x) << 23));
      e.y = __int_as_float(__float_as_int(q.y) + (__float_as_int(t.y) << 23));
    } else {
      e.x = ex2m(e.x);
      e.y = ex2m(e.y);
    }
    acc = fadd2(acc, e);
    w[k] = pack_bf16x2(e.x, e.y);
  }
}

}  // namespace asm_sm
}  // namespace ldn
