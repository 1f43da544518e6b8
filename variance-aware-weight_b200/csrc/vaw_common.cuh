// vaw_common.cuh — shared device/host helpers for the sm_100a kernels.
// Everything here is internal to the shared library; the public surface is include/vaw_b200.h.
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <cuda.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>

// ----------------------------------------------------------------------------------------------
// error plumbing (C ABI: every entry point returns 0 or a negative code; text via vaw_last_error)
// ----------------------------------------------------------------------------------------------
#define VAW_OK 0
#define VAW_ERR_INVALID (-1)
#define VAW_ERR_CUDA (-2)
#define VAW_ERR_UNSUPPORTED (-3)

void vaw_set_error(const char* fmt, ...);

#define VAW_CHECK_ARG(cond, ...)                 \
  do {                                           \
    if (!(cond)) {                               \
      vaw_set_error(__VA_ARGS__);                \
      return VAW_ERR_INVALID;                    \
    }                                            \
  } while (0)

#define VAW_CUDA_TRY(expr)                                                              \
  do {                                                                                  \
    cudaError_t _e = (expr);                                                            \
    if (_e != cudaSuccess) {                                                            \
      vaw_set_error("%s:%d CUDA error %s: %s", __FILE__, __LINE__, cudaGetErrorName(_e), \
                    cudaGetErrorString(_e));                                            \
      return VAW_ERR_CUDA;                                                              \
    }                                                                                   \
  } while (0)

void vaw_note_launch();  // bumps the library-wide kernel launch counter (vaw_launch_count)
#define VAW_LAUNCH_CHECK()            \
  do {                                \
    vaw_note_launch();                \
    VAW_CUDA_TRY(cudaGetLastError()); \
  } while (0)

int vaw_num_sms();  // cached SM count of the current device

// ----------------------------------------------------------------------------------------------
// small device helpers
// ----------------------------------------------------------------------------------------------
#ifdef __CUDACC__

typedef __nv_bfloat16 bf16;
typedef __nv_bfloat162 bf162;

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

__device__ __forceinline__ uint32_t pack_bf16(float lo, float hi) {
  bf162 v = __floats2bfloat162_rn(lo, hi);
  return *reinterpret_cast<uint32_t*>(&v);
}
__device__ __forceinline__ float2 unpack_bf16(uint32_t u) {
  bf162 v = *reinterpret_cast<bf162*>(&u);
  return __bfloat1622float2(v);
}

// streaming 128-bit global accesses (inputs read once / outputs written once)
__device__ __forceinline__ void prefetch_l2(const void* p) { asm volatile("prefetch.global.L2 [%0];" ::"l"(p)); }
__device__ __forceinline__ float4 ldg_stream_f4(const float4* p) {
  float4 r;
  asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];"
               : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w)
               : "l"(p));
  return r;
}
__device__ __forceinline__ uint4 ldg_stream_u4(const uint4* p) {
  uint4 r;
  asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];"
               : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w)
               : "l"(p));
  return r;
}
__device__ __forceinline__ uint2 ldg_stream_u2(const uint2* p) {
  uint2 r;
  asm volatile("ld.global.nc.L1::no_allocate.v2.u32 {%0,%1}, [%2];" : "=r"(r.x), "=r"(r.y) : "l"(p));
  return r;
}
__device__ __forceinline__ void stg_stream_f4(float4* p, float4 v) {
  asm volatile("st.global.L1::no_allocate.v4.f32 [%0], {%1,%2,%3,%4};" ::"l"(p), "f"(v.x), "f"(v.y),
               "f"(v.z), "f"(v.w)
               : "memory");
}
__device__ __forceinline__ void stg_stream_u4(uint4* p, uint4 v) {
  asm volatile("st.global.L1::no_allocate.v4.u32 [%0], {%1,%2,%3,%4};" ::"l"(p), "r"(v.x), "r"(v.y),
               "r"(v.z), "r"(v.w)
               : "memory");
}
__device__ __forceinline__ void stg_stream_u2(uint2* p, uint2 v) {
  asm volatile("st.global.L1::no_allocate.v2.u32 [%0], {%1,%2};" ::"l"(p), "r"(v.x), "r"(v.y) : "memory");
}

// activation functions (forward value and derivative), fp32 with hardware-approximate transcendentals
// (tanh.approx / ex2.approx / rcp.approx: relative error ~2^-11, below the bf16 resolution of every consumer).
__device__ __forceinline__ float tanh_fast(float x) {
  float y;
  asm("tanh.approx.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ float sigmoid_fast(float x) { return __fdividef(1.f, 1.f + __expf(-x)); }
__device__ __forceinline__ float gelu_tanh_f(float x) {
  const float k0 = 0.7978845608028654f, k1 = 0.044715f;
  const float u = k0 * x * fmaf(k1, x * x, 1.f);
  return 0.5f * x * (1.f + tanh_fast(u));
}
__device__ __forceinline__ float gelu_tanh_grad_f(float x) {
  const float k0 = 0.7978845608028654f, k1 = 0.044715f;
  const float x2 = x * x;
  const float th = tanh_fast(k0 * x * fmaf(k1, x2, 1.f));
  const float du = k0 * fmaf(3.f * k1, x2, 1.f);
  return 0.5f * (1.f + th) + 0.5f * x * (1.f - th * th) * du;
}
// Two elements per instruction (Blackwell FFMA2 / FMUL2): same formulas, half the issue slots.  Used by the GEMM
// epilogues, where the per-element math - not the main loop - bounded the short-K GEMMs.
__device__ __forceinline__ float2 gelu_tanh_f2(float2 x) {
  const float k0 = 0.7978845608028654f, k1 = 0.044715f;
  const float2 x2 = __fmul2_rn(x, x);
  const float2 u = __fmul2_rn(x, __ffma2_rn(x2, make_float2(k0 * k1, k0 * k1), make_float2(k0, k0)));
  const float2 t = make_float2(tanh_fast(u.x), tanh_fast(u.y));
  const float2 hx = __fmul2_rn(x, make_float2(0.5f, 0.5f));
  return __ffma2_rn(hx, t, hx);
}
__device__ __forceinline__ float2 gelu_tanh_grad_f2(float2 x) {
  const float k0 = 0.7978845608028654f, k1 = 0.044715f;
  const float2 x2 = __fmul2_rn(x, x);
  const float2 u = __fmul2_rn(x, __ffma2_rn(x2, make_float2(k0 * k1, k0 * k1), make_float2(k0, k0)));
  const float2 t = make_float2(tanh_fast(u.x), tanh_fast(u.y));
  const float2 omt2 = __ffma2_rn(make_float2(-t.x, -t.y), t, make_float2(1.f, 1.f));
  const float2 du = __ffma2_rn(x2, make_float2(3.f * k0 * k1, 3.f * k0 * k1), make_float2(k0, k0));
  const float2 a = __fmul2_rn(__fmul2_rn(x, make_float2(0.5f, 0.5f)), omt2);
  const float2 b = __ffma2_rn(t, make_float2(0.5f, 0.5f), make_float2(0.5f, 0.5f));
  return __ffma2_rn(a, du, b);
}
// exact-GELU via erf(z) ~ 1 - (a1 t + ... + a5 t^5) exp(-z^2), t = 1/(1 + p|z|)  (Abramowitz-Stegun 7.1.26, |err| < 1.5e-7)
__device__ __forceinline__ float erf_fast(float z) {
  const float az = fabsf(z);
  const float t = __fdividef(1.f, fmaf(0.3275911f, az, 1.f));
  const float poly = t * fmaf(t, fmaf(t, fmaf(t, fmaf(t, 1.061405429f, -1.453152027f), 1.421413741f), -0.284496736f), 0.254829592f);
  const float r = 1.f - poly * __expf(-az * az);
  return copysignf(r, z);
}
__device__ __forceinline__ float gelu_erf_f(float x) { return 0.5f * x * (1.f + erf_fast(x * 0.7071067811865476f)); }
__device__ __forceinline__ float gelu_erf_grad_f(float x) {
  const float cdf = 0.5f * (1.f + erf_fast(x * 0.7071067811865476f));
  const float pdf = 0.3989422804014327f * __expf(-0.5f * x * x);
  return cdf + x * pdf;
}
__device__ __forceinline__ float silu_f(float x) { return x * sigmoid_fast(x); }
__device__ __forceinline__ float silu_grad_f(float x) {
  const float s = sigmoid_fast(x);
  return s * (1.f + x * (1.f - s));
}

#endif  // __CUDACC__
