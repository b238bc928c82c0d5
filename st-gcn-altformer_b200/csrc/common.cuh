// Shared device/host helpers for the altformer_b200 kernels (sm_100a only).
#pragma once
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/altformer_b200.h"

namespace afb {

using bf16 = __nv_bfloat16;

// ---- host-side error plumbing ------------------------------------------------------------
void set_error(const char* fmt, ...);
int check_launch(const char* what);  // returns 0 or a cudaError code, records message
int next_stream_dir();                // 0 = ascending, 1 = descending row order for this launch (api.cu)

#define AFB_REQUIRE(cond, ...)          \
  do {                                  \
    if (!(cond)) {                      \
      ::afb::set_error(__VA_ARGS__);    \
      return AFB_ERR_INVALID;           \
    }                                   \
  } while (0)

static inline cudaStream_t as_stream(void* s) { return reinterpret_cast<cudaStream_t>(s); }
static inline int ceil_div(int64_t a, int64_t b) { return (int)((a + b - 1) / b); }

// ---- programmatic dependent launch (PDL) ------------------------------------------------------------------------
// In the small-batch regimes (batch 32 inference, one 256-sequence batch split over 8 GPUs) the step is ~260 launches of
// 5-15 us each, most of them one-CTA-per-SM kernels with a 2-4 us prologue (barrier init, TMEM allocation, descriptor
// prefetch).  Launched with the programmatic-stream-serialization attribute, a kernel's CTAs become resident as soon as the
// previous kernel's CTAs leave their SMs, run that prologue, and block in pdl_wait() until the previous grid has completed
// and flushed; EVERY thread calls pdl_wait() before its first access to global memory.  pdl_launch_dependents() (first
// instruction of a kernel) lets the next launch do the same to us.  Both are no-ops in a normally launched grid.
// AFB_PDL=0 launches everything the ordinary way; `small` = the launch is short enough for its prologue to matter.
__device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
bool pdl_enabled();
bool pdl_force_all();
template <typename... KArgs, typename... Args>
inline cudaError_t launch_pdl(bool small, void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, Args&&... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = (small && pdl_enabled()) ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
}

// ---- typed global load/store (fp32 math everywhere) ---------------------------------------
template <typename T> __device__ __forceinline__ float ldf(const T* p);
template <> __device__ __forceinline__ float ldf<float>(const float* p) { return *p; }
template <> __device__ __forceinline__ float ldf<bf16>(const bf16* p) { return __bfloat162float(*p); }
template <typename T> __device__ __forceinline__ void stf(T* p, float v);
template <> __device__ __forceinline__ void stf<float>(float* p, float v) { *p = v; }
template <> __device__ __forceinline__ void stf<bf16>(bf16* p, float v) { *p = __float2bfloat16_rn(v); }

// two consecutive elements (address must be aligned to 2 elements)
template <typename T> __device__ __forceinline__ float2 ld2(const T* p);
template <> __device__ __forceinline__ float2 ld2<float>(const float* p) { return *reinterpret_cast<const float2*>(p); }
template <> __device__ __forceinline__ float2 ld2<bf16>(const bf16* p) {
  return __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(p));
}
template <typename T> __device__ __forceinline__ void st2(T* p, float a, float b);
template <> __device__ __forceinline__ void st2<float>(float* p, float a, float b) {
  *reinterpret_cast<float2*>(p) = make_float2(a, b);
}
template <> __device__ __forceinline__ void st2<bf16>(bf16* p, float a, float b) {
  *reinterpret_cast<__nv_bfloat162*>(p) = __floats2bfloat162_rn(a, b);
}

// four consecutive elements (address aligned to 4 elements)
template <typename T> __device__ __forceinline__ float4 ld4(const T* p);
template <> __device__ __forceinline__ float4 ld4<float>(const float* p) { return *reinterpret_cast<const float4*>(p); }
template <> __device__ __forceinline__ float4 ld4<bf16>(const bf16* p) {
  uint2 raw = *reinterpret_cast<const uint2*>(p);
  float2 a = __bfloat1622float2(*reinterpret_cast<__nv_bfloat162*>(&raw.x));
  float2 b = __bfloat1622float2(*reinterpret_cast<__nv_bfloat162*>(&raw.y));
  return make_float4(a.x, a.y, b.x, b.y);
}
template <typename T> __device__ __forceinline__ void st4(T* p, float4 v);
template <> __device__ __forceinline__ void st4<float>(float* p, float4 v) { *reinterpret_cast<float4*>(p) = v; }
template <> __device__ __forceinline__ void st4<bf16>(bf16* p, float4 v) {
  __nv_bfloat162 a = __floats2bfloat162_rn(v.x, v.y), b = __floats2bfloat162_rn(v.z, v.w);
  uint2 raw;
  raw.x = *reinterpret_cast<uint32_t*>(&a);
  raw.y = *reinterpret_cast<uint32_t*>(&b);
  *reinterpret_cast<uint2*>(p) = raw;
}

// runtime-dtype scalar access for epilogues (dt: AFB_F32 / AFB_BF16)
__device__ __forceinline__ float ld_dyn(const void* base, int64_t idx, int dt) {
  return dt == AFB_BF16 ? __bfloat162float(reinterpret_cast<const bf16*>(base)[idx])
                        : reinterpret_cast<const float*>(base)[idx];
}
__device__ __forceinline__ float2 ld2_dyn(const void* base, int64_t idx, int dt) {
  return dt == AFB_BF16 ? ld2<bf16>(reinterpret_cast<const bf16*>(base) + idx)
                        : ld2<float>(reinterpret_cast<const float*>(base) + idx);
}
__device__ __forceinline__ void st2_dyn(void* base, int64_t idx, int dt, float a, float b) {
  if (dt == AFB_BF16) st2<bf16>(reinterpret_cast<bf16*>(base) + idx, a, b);
  else st2<float>(reinterpret_cast<float*>(base) + idx, a, b);
}

// ---- warp / block reductions ----------------------------------------------------------------
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
__device__ __forceinline__ double warp_sum_d(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// exact (erf) GELU and its derivative -- nn.GELU default (model_ST.py:17)
// Phi(x) = 0.5 (1 + erf(x / sqrt 2)) via Abramowitz-Stegun 7.1.26 (|error| <= 1.5e-7 on erf): one exp, one
// reciprocal and five FMAs -- the same exp(-x^2/2) also yields the normal pdf needed by the derivative.
__device__ __forceinline__ void gelu_parts(float x, float& cdf, float& u) {
  const float z = fabsf(x) * 0.70710678118654752f;
  const float t = __fdividef(1.0f, fmaf(0.3275911f, z, 1.0f));
  u = __expf(-z * z);
  float poly = fmaf(1.061405429f, t, -1.453152027f);
  poly = fmaf(poly, t, 1.421413741f);
  poly = fmaf(poly, t, -0.284496736f);
  poly = fmaf(poly, t, 0.254829592f);
  const float h = 0.5f * poly * t * u;  // 0.5 (1 - erf(z))
  cdf = x >= 0.f ? 1.0f - h : h;
}
__device__ __forceinline__ float gelu_f(float x) {
  float cdf, u;
  gelu_parts(x, cdf, u);
  return x * cdf;
}
__device__ __forceinline__ float gelu_grad_f(float x) {
  float cdf, u;
  gelu_parts(x, cdf, u);
  return fmaf(x * 0.3989422804014327f, u, cdf);
}

// bf16-mode GELU: tanh form on the MUFU.TANH unit (1 special-function op instead of 2).  |gelu_tanh - gelu_erf|
// <= 4.7e-4, below the bf16 resolution of the stored value; the fp32 parity mode keeps the erf form above.
__device__ __forceinline__ float tanh_approx(float x) {
  float y;
  asm("tanh.approx.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ float gelu_fast_f(float x) {
  const float u = x * fmaf(0.0356774081f, x * x, 0.7978845608f);   // sqrt(2/pi) (x + 0.044715 x^3)
  const float hx = 0.5f * x;
  return fmaf(hx, tanh_approx(u), hx);
}
__device__ __forceinline__ float gelu_fast_grad_f(float x) {
  const float x2 = x * x;
  const float th = tanh_approx(x * fmaf(0.0356774081f, x2, 0.7978845608f));
  const float du = fmaf(0.1070322243f, x2, 0.7978845608f);           // d/dx of the tanh argument
  return fmaf(0.5f * x * du, fmaf(-th, th, 1.0f), fmaf(0.5f, th, 0.5f));
}

}  // namespace afb
