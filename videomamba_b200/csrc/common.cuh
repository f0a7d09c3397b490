// Shared helpers for libvmb200 kernels (sm_100a only).
#pragma once
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/vmb200.h"

namespace vmb {

// ---- error plumbing (thread-local message, never abort) ---------------------------------
void set_error(const char* fmt, ...);
int cuda_fail(cudaError_t e, const char* what);

#define VMB_CHECK_ARG(cond, ...)      \
  do {                                \
    if (!(cond)) {                    \
      ::vmb::set_error(__VA_ARGS__);  \
      return VMB_ERR_INVALID;         \
    }                                 \
  } while (0)

#define VMB_UNSUPPORTED(...)          \
  do {                                \
    ::vmb::set_error(__VA_ARGS__);    \
    return VMB_ERR_UNSUPPORTED;       \
  } while (0)

#define VMB_CUDA(call)                                        \
  do {                                                        \
    cudaError_t e__ = (call);                                 \
    if (e__ != cudaSuccess) return ::vmb::cuda_fail(e__, #call); \
  } while (0)

extern unsigned long long g_launches;  // kernels enqueued by this library (vmb_launch_count)

#define VMB_LAUNCH_CHECK(name)                                   \
  do {                                                           \
    ++::vmb::g_launches;                                         \
    cudaError_t e__ = cudaGetLastError();                        \
    if (e__ != cudaSuccess) return ::vmb::cuda_fail(e__, name);  \
  } while (0)

inline cudaStream_t as_stream(vmb_stream_t s) { return reinterpret_cast<cudaStream_t>(s); }
inline int dtype_size(int dt) { return dt == VMB_BF16 ? 2 : 4; }
inline bool dtype_ok(int dt) { return dt == VMB_F32 || dt == VMB_BF16; }
int sm_count();

// ---- per-stage timing (vmb_prof_*): a scope brackets the launches of one stage with events --
extern bool g_prof_on;
void prof_begin(int kind, cudaStream_t st, int* slot);
void prof_end(int slot, cudaStream_t st);
struct ProfScope {
  int slot = -1;
  cudaStream_t st;
  ProfScope(int kind, cudaStream_t s) : st(s) { if (g_prof_on) prof_begin(kind, s, &slot); }
  ~ProfScope() { if (slot >= 0) prof_end(slot, st); }
};

// ---- element conversion -------------------------------------------------------------------
template <typename T> __device__ __forceinline__ float to_f32(T v);
template <> __device__ __forceinline__ float to_f32<float>(float v) { return v; }
template <> __device__ __forceinline__ float to_f32<__nv_bfloat16>(__nv_bfloat16 v) {
  return __bfloat162float(v);
}
template <typename T> __device__ __forceinline__ T from_f32(float v);
template <> __device__ __forceinline__ float from_f32<float>(float v) { return v; }
template <> __device__ __forceinline__ __nv_bfloat16 from_f32<__nv_bfloat16>(float v) {
  return __float2bfloat16_rn(v);
}

// Load element i of an array whose element type is only known at run time (small side inputs).
__device__ __forceinline__ float load_as_f32(const void* p, int64_t i, int dtype) {
  return dtype == VMB_BF16 ? __bfloat162float(reinterpret_cast<const __nv_bfloat16*>(p)[i])
                           : reinterpret_cast<const float*>(p)[i];
}
__device__ __forceinline__ void store_from_f32(void* p, int64_t i, int dtype, float v) {
  if (dtype == VMB_BF16) reinterpret_cast<__nv_bfloat16*>(p)[i] = __float2bfloat16_rn(v);
  else reinterpret_cast<float*>(p)[i] = v;
}

// ---- math -----------------------------------------------------------------------------------
constexpr float kLog2e = 1.4426950408889634f;
constexpr float kLn2 = 0.6931471805599453f;

// Accurate variants (fp32 parity path, 1e-5 budget end to end).
__device__ __forceinline__ float softplus_accurate(float x) {
  // torch.nn.functional.softplus: identity above threshold 20, log1p(exp(x)) below.
  return x > 20.f ? x : log1pf(expf(x));
}
__device__ __forceinline__ float silu_accurate(float x) { return x / (1.f + expf(-x)); }

// Fast variants (bf16 path, 2e-2 budget): MUFU ex2/lg2/rcp.
__device__ __forceinline__ float ex2_approx(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ float lg2_approx(float x) {
  float y;
  asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ float rcp_approx(float x) {
  float y;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ float softplus_fast(float x) {
  if (x > 20.f) return x;
  const float e = ex2_approx(x * kLog2e);
  // log1p(e): for tiny e the 1+e rounding would dominate, use the series instead
  return e < 0.00390625f ? e * (1.f - 0.5f * e + 0.33333334f * e * e)
                         : lg2_approx(1.f + e) * kLn2;
}
__device__ __forceinline__ float tanh_approx(float x) {
  float y;
  asm("tanh.approx.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
// x * sigmoid(x) with sigmoid(x) = 0.5 * tanh(x / 2) + 0.5: one MUFU instead of ex2 + rcp
__device__ __forceinline__ float silu_fast(float x) {
  return x * fmaf(tanh_approx(0.5f * x), 0.5f, 0.5f);
}

template <bool kAccurate> __device__ __forceinline__ float softplus_f(float x) {
  if constexpr (kAccurate) return softplus_accurate(x);
  else return softplus_fast(x);
}
template <bool kAccurate> __device__ __forceinline__ float silu_f(float x) {
  if constexpr (kAccurate) return silu_accurate(x);
  else return silu_fast(x);
}
template <bool kAccurate> __device__ __forceinline__ float exp2_f(float x) {
  if constexpr (kAccurate) return exp2f(x);
  else return ex2_approx(x);
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

}  // namespace vmb
