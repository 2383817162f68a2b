// Shared device/host helpers for the tcvom_b200 kernels (sm_100a).
#pragma once
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include <atomic>

#include "../../include/tcvom_b200.h"

namespace tcv {

extern thread_local char g_err[512];
extern std::atomic<long long> g_launches;

int fail(int code, const char* fmt, ...);
// Checks the launch, counts it, returns TCV_OK or sets the error string.
int launched(const char* what);

#define TCV_REQUIRE(cond, ...)                           \
  do {                                                   \
    if (!(cond)) return tcv::fail(TCV_ERR_INVALID, __VA_ARGS__); \
  } while (0)

#define TCV_CUDA(call)                                                              \
  do {                                                                              \
    cudaError_t e_ = (call);                                                        \
    if (e_ != cudaSuccess)                                                          \
      return tcv::fail(TCV_ERR_CUDA, "%s: %s", #call, cudaGetErrorString(e_));      \
  } while (0)

static inline cudaStream_t S(tcv_stream_t s) { return reinterpret_cast<cudaStream_t>(s); }

// ---- split-bf16 helpers -------------------------------------------------------------
__device__ __forceinline__ void split_bf16(float x, __nv_bfloat16& hi, __nv_bfloat16& lo) {
  hi = __float2bfloat16_rn(x);
  lo = __float2bfloat16_rn(x - __bfloat162float(hi));
}

__device__ __forceinline__ float bf16_bits_to_float(uint16_t b) { return __uint_as_float(((uint32_t)b) << 16); }

// 8 consecutive channels: 16 B from the hi plane + 16 B from the lo plane -> 8 floats
__device__ __forceinline__ void load8(const __nv_bfloat16* hi, long long plane, float* f) {
  uint4 a = *reinterpret_cast<const uint4*>(hi);
  uint4 b = *reinterpret_cast<const uint4*>(hi + plane);
  const uint32_t aw[4] = {a.x, a.y, a.z, a.w};
  const uint32_t bw[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    f[2 * i] = __uint_as_float(aw[i] << 16) + __uint_as_float(bw[i] << 16);
    f[2 * i + 1] = __uint_as_float(aw[i] & 0xffff0000u) + __uint_as_float(bw[i] & 0xffff0000u);
  }
}

__device__ __forceinline__ void load4(const __nv_bfloat16* hi, long long plane, float* f) {
  uint2 a = *reinterpret_cast<const uint2*>(hi);
  uint2 b = *reinterpret_cast<const uint2*>(hi + plane);
  f[0] = __uint_as_float(a.x << 16) + __uint_as_float(b.x << 16);
  f[1] = __uint_as_float(a.x & 0xffff0000u) + __uint_as_float(b.x & 0xffff0000u);
  f[2] = __uint_as_float(a.y << 16) + __uint_as_float(b.y << 16);
  f[3] = __uint_as_float(a.y & 0xffff0000u) + __uint_as_float(b.y & 0xffff0000u);
}

// eight consecutive fp32 per-channel parameters (32-byte aligned) as two 16-byte loads instead of eight scalar ones
__device__ __forceinline__ void ldf8(const float* p, float* f) {
  const float4 a = __ldg(reinterpret_cast<const float4*>(p));
  const float4 b = __ldg(reinterpret_cast<const float4*>(p) + 1);
  f[0] = a.x; f[1] = a.y; f[2] = a.z; f[3] = a.w; f[4] = b.x; f[5] = b.y; f[6] = b.z; f[7] = b.w;
}

__device__ __forceinline__ float load1(const __nv_bfloat16* hi, long long plane) {
  return __bfloat162float(hi[0]) + __bfloat162float(hi[plane]);
}

// two floats -> packed bf16x2 hi word and lo word (one cvt.rn.bf16x2.f32 per plane)
__device__ __forceinline__ void split2_bf16(float x0, float x1, uint32_t& hi, uint32_t& lo) {
  const __nv_bfloat162 h = __floats2bfloat162_rn(x0, x1);
  hi = *reinterpret_cast<const uint32_t*>(&h);
  const float r0 = x0 - __uint_as_float(hi << 16), r1 = x1 - __uint_as_float(hi & 0xffff0000u);
  const __nv_bfloat162 l = __floats2bfloat162_rn(r0, r1);
  lo = *reinterpret_cast<const uint32_t*>(&l);
}

// activation applied to a register array with the (uniform) selector hoisted out of the element loop
template <int N>
__device__ __forceinline__ void apply_act_n(float* f, int act) {
  if (act == TCV_ACT_RELU) {
#pragma unroll
    for (int j = 0; j < N; ++j) f[j] = fmaxf(f[j], 0.f);
  } else if (act == TCV_ACT_LEAKY02) {
#pragma unroll
    for (int j = 0; j < N; ++j) f[j] = f[j] > 0.f ? f[j] : 0.2f * f[j];
  } else if (act == TCV_ACT_TANH01) {
#pragma unroll
    for (int j = 0; j < N; ++j) f[j] = (tanhf(f[j]) + 1.0f) * 0.5f;
  } else if (act == TCV_ACT_LEAKY001) {
#pragma unroll
    for (int j = 0; j < N; ++j) f[j] = f[j] > 0.f ? f[j] : 0.01f * f[j];
  } else if (act == TCV_ACT_CLAMP01) {
#pragma unroll
    for (int j = 0; j < N; ++j) f[j] = fminf(fmaxf(f[j], 0.f), 1.f);
  } else if (act == TCV_ACT_RELU6) {
#pragma unroll
    for (int j = 0; j < N; ++j) f[j] = fminf(fmaxf(f[j], 0.f), 6.f);
  }
}

__device__ __forceinline__ uint32_t pack2(__nv_bfloat16 a, __nv_bfloat16 b) {
  return (uint32_t)__bfloat16_as_ushort(a) | ((uint32_t)__bfloat16_as_ushort(b) << 16);
}

__device__ __forceinline__ void store8(__nv_bfloat16* hi, long long plane, const float* f) {
  uint32_t h[4], l[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) split2_bf16(f[2 * i], f[2 * i + 1], h[i], l[i]);
  *reinterpret_cast<uint4*>(hi) = make_uint4(h[0], h[1], h[2], h[3]);
  *reinterpret_cast<uint4*>(hi + plane) = make_uint4(l[0], l[1], l[2], l[3]);
}

__device__ __forceinline__ void store4(__nv_bfloat16* hi, long long plane, const float* f) {
  uint32_t h[2], l[2];
#pragma unroll
  for (int i = 0; i < 2; ++i) split2_bf16(f[2 * i], f[2 * i + 1], h[i], l[i]);
  *reinterpret_cast<uint2*>(hi) = make_uint2(h[0], h[1]);
  *reinterpret_cast<uint2*>(hi + plane) = make_uint2(l[0], l[1]);
}

__device__ __forceinline__ void store1(__nv_bfloat16* hi, long long plane, float f) {
  __nv_bfloat16 h, l;
  split_bf16(f, h, l);
  hi[0] = h;
  hi[plane] = l;
}

__device__ __forceinline__ int reflect(int i, int n) {
  // torch 'reflect' padding (no edge repeat); valid for -n < i < 2n-1
  if (i < 0) i = -i;
  if (i >= n) i = 2 * n - 2 - i;
  return i;
}

__device__ __forceinline__ float apply_act(float t, int act) {
  switch (act) {
    case TCV_ACT_RELU: return fmaxf(t, 0.f);
    case TCV_ACT_LEAKY02: return t > 0.f ? t : 0.2f * t;
    case TCV_ACT_TANH01: return (tanhf(t) + 1.0f) * 0.5f;
    case TCV_ACT_LEAKY001: return t > 0.f ? t : 0.01f * t;
    case TCV_ACT_CLAMP01: return fminf(fmaxf(t, 0.f), 1.f);
    case TCV_ACT_RELU6: return fminf(fmaxf(t, 0.f), 6.f);
    default: return t;
  }
}

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

}  // namespace tcv
