// Fused residual-add + RMSNorm / LayerNorm (HBM-bound, one warp per row).
//
// Stands in for rms_norm_fn / layer_norm_fn as called by the reference at
// models/videomamba/videomamba.py:151-166 and :902-918.  Per row: read x (and the fp32 residual
// stream), write the fp32 sum back (prenorm) and the normalised row in x's dtype.  The row is
// held in registers between the reduction and the scale, so every byte crosses HBM once.
// Algorithmic bytes per row (bf16 x, fp32 residual): 2D + 4D read, 4D + 2D written.
#include "common.cuh"

namespace vmb {
namespace {

constexpr int kRowsPerCta = 8;  // 8 warps, one row each

template <typename T> struct Vec4;  // 4 consecutive elements
template <> struct Vec4<float> {
  using type = float4;
  static __device__ __forceinline__ void unpack(const float4& v, float* f) {
    f[0] = v.x; f[1] = v.y; f[2] = v.z; f[3] = v.w;
  }
  static __device__ __forceinline__ float4 pack(const float* f) {
    return make_float4(f[0], f[1], f[2], f[3]);
  }
};
template <> struct Vec4<__nv_bfloat16> {
  using type = uint2;
  static __device__ __forceinline__ void unpack(const uint2& v, float* f) {
    const __nv_bfloat162 a = *reinterpret_cast<const __nv_bfloat162*>(&v.x);
    const __nv_bfloat162 b = *reinterpret_cast<const __nv_bfloat162*>(&v.y);
    f[0] = __low2float(a); f[1] = __high2float(a); f[2] = __low2float(b); f[3] = __high2float(b);
  }
  static __device__ __forceinline__ uint2 pack(const float* f) {
    __nv_bfloat162 a = __floats2bfloat162_rn(f[0], f[1]);
    __nv_bfloat162 b = __floats2bfloat162_rn(f[2], f[3]);
    uint2 r;
    r.x = *reinterpret_cast<uint32_t*>(&a);
    r.y = *reinterpret_cast<uint32_t*>(&b);
    return r;
  }
};

// kIters * 128 >= dim; dim % 4 == 0.  TX: x / y type, TR: residual-in type, TO: residual-out type,
// TW: weight type.
template <typename TX, typename TR, typename TO, typename TW, int kIters, bool kRms>
__global__ void __launch_bounds__(kRowsPerCta * 32)
add_norm_kernel(const TX* __restrict__ x, int64_t ldx, const TR* __restrict__ residual,
                const TW* __restrict__ weight, const TW* __restrict__ bias, TX* __restrict__ y,
                TO* __restrict__ residual_out, int64_t rows, int dim, float eps) {
  const int lane = threadIdx.x & 31;
  // CTAs walk the rows back to front: the producer (out_proj) wrote the last rows most recently, so
  // they are the ones still in L2; the consumer (in_proj) starts at row 0, which this kernel writes last
  const int64_t row = (int64_t)(gridDim.x - 1 - blockIdx.x) * kRowsPerCta + (threadIdx.x >> 5);
  if (row >= rows) return;
  const int nvec = dim >> 2;
  float v[kIters][4];
  float sum = 0.f, sumsq = 0.f;
  const TX* xr = x + row * ldx;
#pragma unroll
  for (int it = 0; it < kIters; ++it) {
    const int c = it * 32 + lane;
    if (c < nvec) {
      typename Vec4<TX>::type raw = reinterpret_cast<const typename Vec4<TX>::type*>(xr)[c];
      Vec4<TX>::unpack(raw, v[it]);
      if (residual != nullptr) {
        float r[4];
        typename Vec4<TR>::type rr =
            reinterpret_cast<const typename Vec4<TR>::type*>(residual + row * (int64_t)dim)[c];
        Vec4<TR>::unpack(rr, r);
#pragma unroll
        for (int j = 0; j < 4; ++j) v[it][j] += r[j];
      }
      if (residual_out != nullptr) {
        reinterpret_cast<typename Vec4<TO>::type*>(residual_out + row * (int64_t)dim)[c] =
            Vec4<TO>::pack(v[it]);
      }
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        sum += v[it][j];
        sumsq += v[it][j] * v[it][j];
      }
    } else {
#pragma unroll
      for (int j = 0; j < 4; ++j) v[it][j] = 0.f;
    }
  }
  const float inv_dim = 1.f / (float)dim;
  float mean = 0.f, rstd;
  if constexpr (kRms) {
    sumsq = warp_sum(sumsq);
    rstd = rsqrtf(sumsq * inv_dim + eps);
  } else {
    mean = warp_sum(sum) * inv_dim;
    float var = 0.f;
#pragma unroll
    for (int it = 0; it < kIters; ++it) {
      const int c = it * 32 + lane;
      if (c < nvec) {
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const float d = v[it][j] - mean;
          var += d * d;
        }
      }
    }
    var = warp_sum(var) * inv_dim;
    rstd = rsqrtf(var + eps);
  }
#pragma unroll
  for (int it = 0; it < kIters; ++it) {
    const int c = it * 32 + lane;
    if (c < nvec) {
      float w[4], b[4] = {0.f, 0.f, 0.f, 0.f}, o[4];
      Vec4<TW>::unpack(reinterpret_cast<const typename Vec4<TW>::type*>(weight)[c], w);
      if (bias != nullptr)
        Vec4<TW>::unpack(reinterpret_cast<const typename Vec4<TW>::type*>(bias)[c], b);
#pragma unroll
      for (int j = 0; j < 4; ++j) o[j] = (v[it][j] - mean) * rstd * w[j] + b[j];
      reinterpret_cast<typename Vec4<TX>::type*>(y + row * (int64_t)dim)[c] = Vec4<TX>::pack(o);
    }
  }
}

template <typename TX, typename TR, typename TO, typename TW, bool kRms>
int launch_iters(const void* x, int64_t ldx, const void* residual, const void* weight,
                 const void* bias, void* y, void* residual_out, int64_t rows, int dim, float eps,
                 cudaStream_t st) {
  const unsigned grid = (unsigned)((rows + kRowsPerCta - 1) / kRowsPerCta);
  const int nvec = dim / 4;
#define VMB_AN_LAUNCH(IT)                                                                      \
  add_norm_kernel<TX, TR, TO, TW, IT, kRms><<<grid, kRowsPerCta * 32, 0, st>>>(                \
      (const TX*)x, ldx, (const TR*)residual, (const TW*)weight, (const TW*)bias, (TX*)y,      \
      (TO*)residual_out, rows, dim, eps)
  if (nvec <= 32 * 1) VMB_AN_LAUNCH(1);
  else if (nvec <= 32 * 2) VMB_AN_LAUNCH(2);
  else if (nvec <= 32 * 3) VMB_AN_LAUNCH(3);
  else if (nvec <= 32 * 5) VMB_AN_LAUNCH(5);
  else if (nvec <= 32 * 9) VMB_AN_LAUNCH(9);
  else if (nvec <= 32 * 16) VMB_AN_LAUNCH(16);
  else VMB_UNSUPPORTED("add_norm: dim %d > 2048 not supported", dim);
#undef VMB_AN_LAUNCH
  VMB_LAUNCH_CHECK("add_norm_kernel");
  return VMB_OK;
}

template <typename TX, typename TR, typename TO, typename TW>
int launch_rms(bool rms, const void* x, int64_t ldx, const void* residual, const void* weight,
               const void* bias, void* y, void* residual_out, int64_t rows, int dim, float eps,
               cudaStream_t st) {
  return rms ? launch_iters<TX, TR, TO, TW, true>(x, ldx, residual, weight, bias, y, residual_out,
                                                  rows, dim, eps, st)
             : launch_iters<TX, TR, TO, TW, false>(x, ldx, residual, weight, bias, y,
                                                   residual_out, rows, dim, eps, st);
}

}  // namespace
}  // namespace vmb

extern "C" int vmb_add_norm_fwd(const void* x, int x_dtype, int64_t ldx, const void* residual,
                                int residual_dtype, const void* weight, const void* bias,
                                int w_dtype, void* y, void* residual_out, int residual_out_dtype,
                                int64_t rows, int dim, float eps, int is_rms,
                                vmb_stream_t stream) {
  using namespace vmb;
  using bf16 = __nv_bfloat16;
  VMB_CHECK_ARG(dtype_ok(x_dtype) && dtype_ok(w_dtype), "add_norm: bad dtype");
  VMB_CHECK_ARG(rows >= 0 && dim > 0, "add_norm: bad sizes rows=%lld dim=%d", (long long)rows, dim);
  if (rows == 0) return VMB_OK;  // empty batch: pointers may legitimately be null
  VMB_CHECK_ARG(x && weight && y, "add_norm: null x / weight / y");
  if (dim % 4 != 0 || ldx % 4 != 0) VMB_UNSUPPORTED("add_norm: dim and ldx must be multiples of 4");
  if (!residual) residual_dtype = VMB_F32;
  if (!residual_out) residual_out_dtype = VMB_F32;
  VMB_CHECK_ARG(dtype_ok(residual_dtype) && dtype_ok(residual_out_dtype), "add_norm: bad dtype");
  {
    // rows are read and written as 4-element vectors: 16 B for fp32, 8 B for bf16
    auto aligned = [](const void* p, int dt) {
      return p == nullptr || reinterpret_cast<uintptr_t>(p) % (dt == VMB_F32 ? 16 : 8) == 0;
    };
    if (!aligned(x, x_dtype) || !aligned(y, x_dtype) || !aligned(residual, residual_dtype) ||
        !aligned(residual_out, residual_out_dtype) || !aligned(weight, w_dtype) || !aligned(bias, w_dtype))
      VMB_UNSUPPORTED("add_norm: base pointers must be aligned to 4 elements (16 B fp32 / 8 B bf16)");
  }
  cudaStream_t st = as_stream(stream);
  ProfScope ps(VMB_PROF_ADD_NORM, st);
  const bool rms = is_rms != 0;
  const int key = (x_dtype << 3) | (residual_dtype << 2) | (residual_out_dtype << 1) | w_dtype;
#define VMB_AN_CASE(K, TX, TR, TO, TW)                                                          \
  case K:                                                                                       \
    return launch_rms<TX, TR, TO, TW>(rms, x, ldx, residual, weight, bias, y, residual_out,    \
                                      rows, dim, eps, st)
  switch (key) {
    VMB_AN_CASE(0b0000, float, float, float, float);
    VMB_AN_CASE(0b1001, bf16, float, float, bf16);   // bf16 model, fp32 residual stream
    VMB_AN_CASE(0b1111, bf16, bf16, bf16, bf16);     // bf16 model, residual_in_fp32=False
    VMB_AN_CASE(0b1101, bf16, bf16, float, bf16);
    VMB_AN_CASE(0b1011, bf16, float, bf16, bf16);
    VMB_AN_CASE(0b1000, bf16, float, float, float);  // bf16 activations, fp32 norm weights
    VMB_AN_CASE(0b0001, float, float, float, bf16);
    default:
      VMB_UNSUPPORTED("add_norm: dtype combination x=%d res=%d res_out=%d w=%d not built",
                      x_dtype, residual_dtype, residual_out_dtype, w_dtype);
  }
#undef VMB_AN_CASE
}

// ---- refiner fusion gate --------------------------------------------------------------------
// out = s * fwd + (1 - s) * bwd,  s = sigmoid(g1 (+ g2)):  the sigmoid and the blend of
// BiMambaRefinerBlock (models/refiner_backbone.py:129-134) in one pass; g1 / g2 are the two halves of
// Linear(cat[fwd, bwd]) computed as two projections, so the concatenated tensor never exists.
namespace vmb {
namespace {
template <typename T, bool kAccurate>
__global__ void __launch_bounds__(256)
gate_blend_kernel(const T* __restrict__ g1, const T* __restrict__ g2, const T* __restrict__ fwd,
                  const T* __restrict__ bwd, T* __restrict__ out, int64_t nvec) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= nvec) return;
  using V = typename Vec4<T>::type;
  float a[4], b[4], f[4], r[4], o[4];
  Vec4<T>::unpack(reinterpret_cast<const V*>(g1)[i], a);
  if (g2 != nullptr) {
    Vec4<T>::unpack(reinterpret_cast<const V*>(g2)[i], b);
#pragma unroll
    for (int j = 0; j < 4; ++j) a[j] += b[j];
  }
  Vec4<T>::unpack(reinterpret_cast<const V*>(fwd)[i], f);
  Vec4<T>::unpack(reinterpret_cast<const V*>(bwd)[i], r);
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const float s = kAccurate ? 1.f / (1.f + expf(-a[j])) : fmaf(tanh_approx(0.5f * a[j]), 0.5f, 0.5f);
    o[j] = s * f[j] + (1.f - s) * r[j];
  }
  reinterpret_cast<V*>(out)[i] = Vec4<T>::pack(o);
}
}  // namespace
}  // namespace vmb

extern "C" int vmb_gate_blend_fwd(const void* g1, const void* g2, const void* fwd, const void* bwd,
                                  void* out, int64_t n, int dtype, vmb_stream_t stream) {
  using namespace vmb;
  VMB_CHECK_ARG(dtype_ok(dtype), "gate_blend: bad dtype");
  VMB_CHECK_ARG(n >= 0, "gate_blend: bad size %lld", (long long)n);
  if (n == 0) return VMB_OK;
  VMB_CHECK_ARG(g1 && fwd && bwd && out, "gate_blend: null g1 / fwd / bwd / out");
  if (n % 4 != 0) VMB_UNSUPPORTED("gate_blend: element count must be a multiple of 4");
  {
    const uintptr_t al = dtype == VMB_F32 ? 16 : 8;
    const void* ptrs[5] = {g1, g2, fwd, bwd, out};
    for (const void* p : ptrs)
      if (p != nullptr && reinterpret_cast<uintptr_t>(p) % al != 0)
        VMB_UNSUPPORTED("gate_blend: pointers must be aligned to 4 elements (16 B fp32 / 8 B bf16)");
  }
  const int64_t nvec = n / 4;
  const unsigned grid = (unsigned)((nvec + 255) / 256);
  cudaStream_t st = as_stream(stream);
  if (dtype == VMB_BF16)
    gate_blend_kernel<__nv_bfloat16, false><<<grid, 256, 0, st>>>(
        (const __nv_bfloat16*)g1, (const __nv_bfloat16*)g2, (const __nv_bfloat16*)fwd,
        (const __nv_bfloat16*)bwd, (__nv_bfloat16*)out, nvec);
  else
    gate_blend_kernel<float, true><<<grid, 256, 0, st>>>((const float*)g1, (const float*)g2,
                                                          (const float*)fwd, (const float*)bwd,
                                                          (float*)out, nvec);
  VMB_LAUNCH_CHECK("gate_blend_kernel");
  return VMB_OK;
}
