// common.cuh -- shared helpers for libaide_b200 (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <string>

#include "../../include/aide_b200.h"

namespace aide {

// A column sum out[j] = sum_r src[r * ld + j] that rides along with the split-K reduction of a weight gradient (the conv
// bias gradient of the same unit: one launch less per unit).  aide_conv3x3_wgrad_ex posts it for the calling thread,
// the next launch_wgrad_reduce() on that thread takes it.
struct ColSumJob {
  const float* src;
  int rows, ld, cols;
  float* dst;
};
void wgrad_reduce_post_job(const ColSumJob& job);
bool wgrad_reduce_take_job(ColSumJob* job);


// ------------------------------------------------------------------ error plumbing
void set_error(const char* fmt, ...);
int cuda_fail(cudaError_t e, const char* what, const char* file, int line);

#define AIDE_CUDA(expr)                                                          \
  do {                                                                           \
    cudaError_t _e = (expr);                                                     \
    if (_e != cudaSuccess) return ::aide::cuda_fail(_e, #expr, __FILE__, __LINE__); \
  } while (0)
// every kernel launch of the library is followed by exactly one AIDE_CHECK_LAUNCH(): it also feeds the
// launch counter exported as aide_launch_count() (bench.py's "gpu_launches")
void count_launch();
#define AIDE_CHECK_LAUNCH()          \
  do {                               \
    ::aide::count_launch();          \
    AIDE_CUDA(cudaGetLastError());   \
  } while (0)
#define AIDE_REQUIRE(cond, ...)                 \
  do {                                          \
    if (!(cond)) {                              \
      ::aide::set_error(__VA_ARGS__);           \
      return 1;                                 \
    }                                           \
  } while (0)

constexpr int kNumSMs = 148;  // B200

inline int ceil_div(long long a, long long b) { return (int)((a + b - 1) / b); }

// ------------------------------------------------------------------ operand-format views
struct View {
  void* p0;
  void* p1;
  int ctot;
  int coff;
};
struct CView {
  const void* p0;
  const void* p1;
  int ctot;
  int coff;
};

__device__ __forceinline__ float tf32_rn(float x) {
  uint32_t r;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x));
  return __uint_as_float(r);
}

// ---- AIDE_FMT_F16X2: x * 2^s = hi + lo with hi = fp16(x * 2^s), lo = fp16(x * 2^s - hi) ------------------------
// Two fp16 planes carry 22 significant bits; three kind::f16 MMAs (hi*hi + hi*lo + lo*hi) reproduce the fp32
// product at TWICE the tcgen05 rate of kind::tf32 and with half the operand bytes of TF32X2.  fp16 has a narrow
// exponent range, so every tensor is stored pre-scaled by a power of two (exact): activations by 2^8 (full 22-bit
// precision for |x| in [5e-4, 255], absolute error <= 1e-10 below), weights by 2^12, gradients by a per-tensor
// power of two chosen on the device.  Conversions saturate at +-65504 instead of producing inf.
constexpr float kF16ActScale = 256.f;
constexpr float kF16WScale = 4096.f;
// Sticky saturation flag (one per translation unit, read through aide_f16_saturated()): set when a pre-scaled value
// leaves the fp16 range, i.e. |activation| >= 255.87, |weight| >= 16 or a gradient beyond its dynamic scale's head-room.
// The result is then clipped, which silently breaks the fp32-parity claim -- tests and users can check the flag.
static __device__ unsigned int g_f16_sat_tu = 0;
void register_f16_sat_reader(unsigned (*fn)(bool));
namespace {
struct F16SatRegistrar {
  F16SatRegistrar() {
    register_f16_sat_reader([](bool reset) -> unsigned {
      unsigned v = 0;
      if (cudaMemcpyFromSymbol(&v, g_f16_sat_tu, sizeof(v)) != cudaSuccess) {
        cudaGetLastError();
        return 0u;
      }
      if (reset && v) {
        const unsigned z = 0;
        cudaMemcpyToSymbol(g_f16_sat_tu, &z, sizeof(z));
      }
      return v;
    });
  }
};
static F16SatRegistrar g_f16_sat_registrar;
}  // namespace
__device__ __forceinline__ void f16_split(float xs, __half& hi, __half& lo) {
  if (fabsf(xs) > 65504.f) g_f16_sat_tu = 1u;          // sticky (benign race: every writer stores 1)
  const float c = fminf(fmaxf(xs, -65504.f), 65504.f);
  hi = __float2half_rn(c);
  lo = __float2half_rn(c - __half2float(hi));
}

// Load/store 4 consecutive channels (c % 4 == 0, view channel offsets are multiples of 4).
template <int FMT>
__device__ __forceinline__ float4 ld4(const void* p0, const void* p1, size_t e) {
  if constexpr (FMT == AIDE_FMT_F32) {
    return *reinterpret_cast<const float4*>(reinterpret_cast<const float*>(p0) + e);
  } else if constexpr (FMT == AIDE_FMT_TF32X2) {
    float4 a = *reinterpret_cast<const float4*>(reinterpret_cast<const float*>(p0) + e);
    float4 b = *reinterpret_cast<const float4*>(reinterpret_cast<const float*>(p1) + e);
    return make_float4(a.x + b.x, a.y + b.y, a.z + b.z, a.w + b.w);
  } else if constexpr (FMT == AIDE_FMT_F16X2) {
    const uint2 rh = *reinterpret_cast<const uint2*>(reinterpret_cast<const __half*>(p0) + e);
    const uint2 rl = *reinterpret_cast<const uint2*>(reinterpret_cast<const __half*>(p1) + e);
    const float2 h0 = __half22float2(*reinterpret_cast<const __half2*>(&rh.x));
    const float2 h1 = __half22float2(*reinterpret_cast<const __half2*>(&rh.y));
    const float2 l0 = __half22float2(*reinterpret_cast<const __half2*>(&rl.x));
    const float2 l1 = __half22float2(*reinterpret_cast<const __half2*>(&rl.y));
    constexpr float inv = 1.0f / kF16ActScale;
    return make_float4((h0.x + l0.x) * inv, (h0.y + l0.y) * inv, (h1.x + l1.x) * inv, (h1.y + l1.y) * inv);
  } else {
    uint2 r = *reinterpret_cast<const uint2*>(reinterpret_cast<const __nv_bfloat16*>(p0) + e);
    __nv_bfloat162 lo = *reinterpret_cast<__nv_bfloat162*>(&r.x);
    __nv_bfloat162 hi = *reinterpret_cast<__nv_bfloat162*>(&r.y);
    float2 a = __bfloat1622float2(lo), b = __bfloat1622float2(hi);
    return make_float4(a.x, a.y, b.x, b.y);
  }
}

template <int FMT>
__device__ __forceinline__ void st4(void* p0, void* p1, size_t e, float4 v) {
  if constexpr (FMT == AIDE_FMT_F32) {
    *reinterpret_cast<float4*>(reinterpret_cast<float*>(p0) + e) = v;
  } else if constexpr (FMT == AIDE_FMT_TF32X2) {
    float4 h = make_float4(tf32_rn(v.x), tf32_rn(v.y), tf32_rn(v.z), tf32_rn(v.w));
    float4 l = make_float4(v.x - h.x, v.y - h.y, v.z - h.z, v.w - h.w);
    *reinterpret_cast<float4*>(reinterpret_cast<float*>(p0) + e) = h;
    *reinterpret_cast<float4*>(reinterpret_cast<float*>(p1) + e) = l;
  } else if constexpr (FMT == AIDE_FMT_F16X2) {
    __half h[4], l[4];
    f16_split(v.x * kF16ActScale, h[0], l[0]);
    f16_split(v.y * kF16ActScale, h[1], l[1]);
    f16_split(v.z * kF16ActScale, h[2], l[2]);
    f16_split(v.w * kF16ActScale, h[3], l[3]);
    *reinterpret_cast<uint2*>(reinterpret_cast<__half*>(p0) + e) = *reinterpret_cast<const uint2*>(h);
    *reinterpret_cast<uint2*>(reinterpret_cast<__half*>(p1) + e) = *reinterpret_cast<const uint2*>(l);
  } else {
    __nv_bfloat162 lo = __floats2bfloat162_rn(v.x, v.y);
    __nv_bfloat162 hi = __floats2bfloat162_rn(v.z, v.w);
    uint2 r;
    r.x = *reinterpret_cast<uint32_t*>(&lo);
    r.y = *reinterpret_cast<uint32_t*>(&hi);
    *reinterpret_cast<uint2*>(reinterpret_cast<__nv_bfloat16*>(p0) + e) = r;
  }
}

template <int FMT>
__device__ __forceinline__ float ld1(const void* p0, const void* p1, size_t e) {
  if constexpr (FMT == AIDE_FMT_F32) {
    return reinterpret_cast<const float*>(p0)[e];
  } else if constexpr (FMT == AIDE_FMT_TF32X2) {
    return reinterpret_cast<const float*>(p0)[e] + reinterpret_cast<const float*>(p1)[e];
  } else if constexpr (FMT == AIDE_FMT_F16X2) {
    return (__half2float(reinterpret_cast<const __half*>(p0)[e]) + __half2float(reinterpret_cast<const __half*>(p1)[e])) *
           (1.0f / kF16ActScale);
  } else {
    return __bfloat162float(reinterpret_cast<const __nv_bfloat16*>(p0)[e]);
  }
}
template <int FMT>
__device__ __forceinline__ void st1(void* p0, void* p1, size_t e, float v) {
  if constexpr (FMT == AIDE_FMT_F32) {
    reinterpret_cast<float*>(p0)[e] = v;
  } else if constexpr (FMT == AIDE_FMT_TF32X2) {
    float h = tf32_rn(v);
    reinterpret_cast<float*>(p0)[e] = h;
    reinterpret_cast<float*>(p1)[e] = v - h;
  } else if constexpr (FMT == AIDE_FMT_F16X2) {
    __half h, l;
    f16_split(v * kF16ActScale, h, l);
    reinterpret_cast<__half*>(p0)[e] = h;
    reinterpret_cast<__half*>(p1)[e] = l;
  } else {
    reinterpret_cast<__nv_bfloat16*>(p0)[e] = __float2bfloat16_rn(v);
  }
}

inline int fmt_elem_bytes(int fmt) { return (fmt == AIDE_FMT_BF16 || fmt == AIDE_FMT_F16X2) ? 2 : 4; }
inline int fmt_planes(int fmt) { return (fmt == AIDE_FMT_TF32X2 || fmt == AIDE_FMT_F16X2) ? 2 : 1; }
inline bool fmt_valid(int fmt) { return fmt >= 0 && fmt <= 3; }

// dispatch a templated launcher on the runtime format
#define AIDE_DISPATCH_FMT(fmt, ...)                                   \
  switch (fmt) {                                                      \
    case AIDE_FMT_F32: { constexpr int FMT = AIDE_FMT_F32; __VA_ARGS__; break; }       \
    case AIDE_FMT_TF32X2: { constexpr int FMT = AIDE_FMT_TF32X2; __VA_ARGS__; break; } \
    case AIDE_FMT_BF16: { constexpr int FMT = AIDE_FMT_BF16; __VA_ARGS__; break; }     \
    case AIDE_FMT_F16X2: { constexpr int FMT = AIDE_FMT_F16X2; __VA_ARGS__; break; }   \
    default: ::aide::set_error("bad operand format %d", fmt); return 1;                 \
  }

inline cudaStream_t as_stream(void* s) { return reinterpret_cast<cudaStream_t>(s); }

// out[j] = sum_r partial[r*ld + j] (j < cols), fp64 accumulation in fixed order (defined in bn.cu)
int launch_reduce_rows(const float* partial, int rows, int ld, int cols, float* out, cudaStream_t st);

// warp / block reductions ------------------------------------------------------------------
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

}  // namespace aide
