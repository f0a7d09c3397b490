// Fused single-token decode step of the mixer: everything between in_proj and out_proj of the
// reference's Mamba.step (models/videomamba/mamba_simple.py:466-494) in ONE kernel --
//     causal_conv1d_update (:468-474)  ->  x_proj (:476)  ->  split, dt_proj without bias (:477-479)
//     ->  selective_state_update with dt_bias / softplus / D skip / SiLU(z) gate (:483-494)
// -- with both states updated in place.  The reference issues four operator calls here (two of them
// cuBLAS GEMMs with M = batch); at M = 1..32 each is a few microseconds of launch for nanoseconds of
// work.  One CTA per batch row; the row's conv output, x_dbl and delta live in shared memory.
// Rounding points follow the reference: conv output, x_dbl and delta_raw are rounded to the model
// dtype between the stages; the state update runs in fp32.
#include "internal.h"

namespace vmb {
namespace {

constexpr int kStepThreads = 256;

template <typename T, bool kAccurate>
__global__ void __launch_bounds__(kStepThreads)
mixer_step_kernel(const vmb_step_args a) {
  extern __shared__ float smem[];
  const int Di = a.Di, N = a.N, R = a.R, W = a.W, X = a.R + 2 * a.N;
  float* s_xc = smem;              // [Di]  conv output (rounded to T)
  float* s_xdb = smem + Di;        // [X]   x_dbl row (rounded to T)
  const int b = blockIdx.x;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const T* xz = reinterpret_cast<const T*>(a.xz) + (int64_t)b * a.xz_bstride;
  const T* w_conv = reinterpret_cast<const T*>(a.w_conv);
  const T* b_conv = reinterpret_cast<const T*>(a.b_conv);

  // ---- conv step: roll the window, store it in the state's dtype, convolve the STORED values ----
  for (int c = tid; c < Di; c += kStepThreads) {
    const int64_t base = ((int64_t)b * Di + c) * W;
    float acc = b_conv ? to_f32<T>(b_conv[c]) : 0.f;
    for (int k = 0; k < W; ++k) {
      const float v = k + 1 < W ? load_as_f32(a.conv_state, base + k + 1, a.cs_dtype) : to_f32<T>(xz[c]);
      store_from_f32(a.conv_state, base + k, a.cs_dtype, v);
      const float r = a.cs_dtype == VMB_BF16 ? __bfloat162float(__float2bfloat16_rn(v)) : v;
      acc = fmaf(to_f32<T>(w_conv[(int64_t)c * W + k]), r, acc);
    }
    s_xc[c] = to_f32<T>(from_f32<T>(silu_f<kAccurate>(acc)));
  }
  __syncthreads();

  // ---- x_proj: X dot products of length Di, one warp per output row (coalesced weight rows) ----
  const T* w_x = reinterpret_cast<const T*>(a.w_x);
  for (int j = warp; j < X; j += kStepThreads / 32) {
    const T* wr = w_x + (int64_t)j * Di;
    float acc = 0.f;
    for (int c = lane; c < Di; c += 32) acc = fmaf(s_xc[c], to_f32<T>(wr[c]), acc);
    acc = warp_sum(acc);
    if (lane == 0) s_xdb[j] = to_f32<T>(from_f32<T>(acc));
  }
  __syncthreads();

  // ---- dt_proj (no bias) + state update + D skip + gate, one channel per thread --------------------
  const T* w_dt = reinterpret_cast<const T*>(a.w_dt);
  T* y = reinterpret_cast<T*>(a.y) + (int64_t)b * a.y_bstride;
  const float* Bm = s_xdb + R;
  const float* Cm = s_xdb + R + N;
  for (int c = tid; c < Di; c += kStepThreads) {
    float draw = 0.f;
    for (int r = 0; r < R; ++r) draw = fmaf(s_xdb[r], to_f32<T>(w_dt[(int64_t)c * R + r]), draw);
    draw = to_f32<T>(from_f32<T>(draw));                       // F.linear output in the model dtype
    const float dv = softplus_f<kAccurate>(draw + (a.dt_bias ? a.dt_bias[c] : 0.f));
    const float xv = s_xc[c];
    const float du = dv * xv;
    const int64_t base = ((int64_t)b * Di + c) * N;
    float acc = 0.f;
    for (int n = 0; n < N; ++n) {
      const float e = exp2_f<kAccurate>(dv * a.A2[(int64_t)c * N + n]);
      const float hn = fmaf(e, load_as_f32(a.ssm_state, base + n, a.ss_dtype), du * Bm[n]);
      store_from_f32(a.ssm_state, base + n, a.ss_dtype, hn);
      acc = fmaf(hn, Cm[n], acc);
    }
    if (a.Dskip) acc = fmaf(a.Dskip[c], xv, acc);
    acc *= silu_f<kAccurate>(to_f32<T>(xz[Di + c]));
    y[c] = from_f32<T>(acc);
  }
}

}  // namespace
}  // namespace vmb

extern "C" int vmb_mixer_step_fwd(const vmb_step_args* a, vmb_stream_t stream) {
  using namespace vmb;
  VMB_CHECK_ARG(a != nullptr, "mixer_step: null args");
  VMB_CHECK_ARG(dtype_ok(a->dtype) && dtype_ok(a->cs_dtype) && dtype_ok(a->ss_dtype), "mixer_step: bad dtype");
  VMB_CHECK_ARG(a->B >= 0 && a->Di > 0 && a->N > 0 && a->R > 0 && a->W > 0, "mixer_step: bad sizes");
  if (a->B == 0) return VMB_OK;
  VMB_CHECK_ARG(a->xz && a->conv_state && a->ssm_state && a->w_conv && a->w_x && a->w_dt && a->A2 && a->y,
                "mixer_step: null tensor");
  const size_t smem = (size_t)(a->Di + a->R + 2 * a->N) * sizeof(float);
  if (smem > 200 * 1024) VMB_UNSUPPORTED("mixer_step: d_inner %d too large for the shared-memory row", a->Di);
  cudaStream_t st = as_stream(stream);
  ProfScope ps(VMB_PROF_OTHER, st);
  if (a->dtype == VMB_F32) {
    if (smem > 48 * 1024)
      VMB_CUDA(cudaFuncSetAttribute(mixer_step_kernel<float, true>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                    (int)smem));
    mixer_step_kernel<float, true><<<a->B, kStepThreads, smem, st>>>(*a);
  } else {
    if (smem > 48 * 1024)
      VMB_CUDA(cudaFuncSetAttribute(mixer_step_kernel<__nv_bfloat16, false>,
                                    cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    mixer_step_kernel<__nv_bfloat16, false><<<a->B, kStepThreads, smem, st>>>(*a);
  }
  VMB_LAUNCH_CHECK("mixer_step_kernel");
  return VMB_OK;
}
