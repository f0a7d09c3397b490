// Selective scan, any-shape / true-fp32 kernel, plus the single-step state update.
//
// Semantics = the reference's _selective_scan_ref (models/videomamba/mamba_simple.py:30-106):
//   delta = softplus(delta_raw + bias); h = exp(delta*A)*h + delta*B_t*u_t; y = <C_t,h> + D*u;
//   y *= silu(z); h_last = h (fp32).
// One thread owns one (batch, channel) chain with all N states in registers and walks the
// sequence in order, so the work per (token, channel, state) is the minimum: one exp2, one mul
// for the drive, two FMAs.  B_t / C_t rows of a token chunk are staged once per CTA in shared
// memory (fp32) and read back as warp broadcasts.  This kernel takes delta_raw as a tensor (the
// projection ran as a separate GEMM, rounding delta_raw to the model dtype exactly where the
// reference does, mamba_simple.py:413-414); the production bf16 shapes use scan_fast.cu.
#include "common.cuh"

namespace vmb {
namespace {

constexpr int kThreads = 128;
constexpr int kTok = 32;  // tokens staged per step

template <typename T, int NMAX, bool kAccurate>
__global__ void __launch_bounds__(kThreads)
scan_generic_kernel(const vmb_scan_args a) {
  __shared__ float sB[kTok][NMAX];
  __shared__ float sC[kTok][NMAX];
  const int tid = threadIdx.x;
  const int b = blockIdx.y;
  const int d = blockIdx.x * kThreads + tid;
  const bool valid = d < a.Di;
  const int N = a.N, L = a.L;

  float A2[NMAX], h[NMAX];
#pragma unroll
  for (int n = 0; n < NMAX; ++n) {
    A2[n] = (valid && n < N) ? a.A2[(int64_t)d * N + n] : 0.f;
    h[n] = 0.f;
    if (valid && n < N && a.h0 != nullptr)
      h[n] = load_as_f32(a.h0, ((int64_t)b * a.Di + d) * N + n, a.h0_dtype);
  }
  const float Dv = (valid && a.D) ? a.D[d] : 0.f;
  const float bias = (valid && a.dt_bias) ? a.dt_bias[d] : 0.f;
  const T* u = reinterpret_cast<const T*>(a.u) + (int64_t)b * a.u_bstride + d;
  const T* dl = reinterpret_cast<const T*>(a.delta) + (int64_t)b * a.d_bstride + d;
  const T* z = a.z ? reinterpret_cast<const T*>(a.z) + (int64_t)b * a.z_bstride + d : nullptr;
  T* y = reinterpret_cast<T*>(a.y) + (int64_t)b * a.y_bstride + d;
  const T* bc = reinterpret_cast<const T*>(a.bc) + (int64_t)b * a.bc_bstride;
  // physical row of logical token i (forward / whole-sequence reversal / frame-axis reversal)
  const int F = a.frame_len;
  auto row = [&](int i) -> int64_t {
    if (!a.reverse) return (int64_t)i;
    if (F <= 0) return (int64_t)(L - 1 - i);
    const int f = i / F;
    return (int64_t)(L - (f + 1) * F + (i - f * F));
  };

  for (int c0 = 0; c0 < L; c0 += kTok) {
    const int nt = min(kTok, L - c0);
    for (int e = tid; e < nt * N; e += kThreads) {
      const int t = e / N, n = e % N;
      const int64_t r = row(c0 + t);
      sB[t][n] = to_f32<T>(bc[r * a.bc_tstride + a.b_off + n]);
      sC[t][n] = to_f32<T>(bc[r * a.bc_tstride + a.c_off + n]);
    }
    __syncthreads();
    if (valid) {
#pragma unroll 2
      for (int t = 0; t < nt; ++t) {
        const int64_t r = row(c0 + t);
        const float uv = to_f32<T>(u[r * a.u_tstride]);
        float dv = to_f32<T>(dl[r * a.d_tstride]) + bias;
        if (a.softplus) dv = softplus_f<kAccurate>(dv);
        const float du = dv * uv;
        float acc = 0.f;
#pragma unroll
        for (int n = 0; n < NMAX; ++n) {
          if (n < N) {
            const float e = exp2_f<kAccurate>(dv * A2[n]);
            h[n] = fmaf(e, h[n], du * sB[t][n]);
            acc = fmaf(h[n], sC[t][n], acc);
          }
        }
        acc = fmaf(Dv, uv, acc);
        if (z != nullptr) acc *= silu_f<kAccurate>(to_f32<T>(z[r * a.z_tstride]));
        y[r * a.y_tstride] = from_f32<T>(acc);
      }
    }
    __syncthreads();
  }
  if (valid && a.h_last != nullptr) {
#pragma unroll
    for (int n = 0; n < NMAX; ++n)
      if (n < N) a.h_last[((int64_t)b * a.Di + d) * N + n] = h[n];
  }
}

template <typename T, bool kAccurate>
__global__ void state_update_kernel(void* __restrict__ state, int state_dtype,
                                    const T* __restrict__ x, int64_t x_bs,
                                    const T* __restrict__ dt, int64_t dt_bs,
                                    const float* __restrict__ A2, const T* __restrict__ Bm,
                                    int64_t b_bs, const T* __restrict__ Cm, int64_t c_bs,
                                    const float* __restrict__ D, const T* __restrict__ z,
                                    int64_t z_bs, const float* __restrict__ dt_bias, int softplus,
                                    T* __restrict__ y, int64_t y_bs, int Di, int N) {
  const int d = blockIdx.x * blockDim.x + threadIdx.x;
  const int b = blockIdx.y;
  if (d >= Di) return;
  const float xv = to_f32<T>(x[(int64_t)b * x_bs + d]);
  float dv = to_f32<T>(dt[(int64_t)b * dt_bs + d]) + (dt_bias ? dt_bias[d] : 0.f);
  if (softplus) dv = softplus_f<kAccurate>(dv);
  const float du = dv * xv;
  float acc = 0.f;
  const int64_t base = ((int64_t)b * Di + d) * N;
  for (int n = 0; n < N; ++n) {
    const float e = exp2_f<kAccurate>(dv * A2[(int64_t)d * N + n]);
    const float hn = fmaf(e, load_as_f32(state, base + n, state_dtype),
                          du * to_f32<T>(Bm[(int64_t)b * b_bs + n]));
    store_from_f32(state, base + n, state_dtype, hn);
    acc = fmaf(hn, to_f32<T>(Cm[(int64_t)b * c_bs + n]), acc);
  }
  if (D) acc = fmaf(D[d], xv, acc);
  if (z) acc *= silu_f<kAccurate>(to_f32<T>(z[(int64_t)b * z_bs + d]));
  y[(int64_t)b * y_bs + d] = from_f32<T>(acc);
}

}  // namespace

int scan_generic(const vmb_scan_args& a, cudaStream_t st) {
  dim3 grid((a.Di + kThreads - 1) / kThreads, a.B);
#define VMB_SG(T, NMAX, ACC) scan_generic_kernel<T, NMAX, ACC><<<grid, kThreads, 0, st>>>(a)
  if (a.dtype == VMB_F32) {
    if (a.N <= 16) VMB_SG(float, 16, true);
    else if (a.N <= 64) VMB_SG(float, 64, true);
    else VMB_UNSUPPORTED("selective_scan: d_state=%d > 64 not supported", a.N);
  } else {
    if (a.N <= 16) VMB_SG(__nv_bfloat16, 16, false);
    else if (a.N <= 64) VMB_SG(__nv_bfloat16, 64, false);
    else VMB_UNSUPPORTED("selective_scan: d_state=%d > 64 not supported", a.N);
  }
#undef VMB_SG
  VMB_LAUNCH_CHECK("scan_generic_kernel");
  return VMB_OK;
}

}  // namespace vmb

extern "C" int vmb_selective_state_update(void* state, int state_dtype, const void* x, int64_t x_bs,
                                          const void* dt, int64_t dt_bs, const float* A2,
                                          const void* Bm, int64_t b_bs, const void* Cm,
                                          int64_t c_bs, const float* D, const void* z, int64_t z_bs,
                                          const float* dt_bias, int softplus, void* y, int64_t y_bs,
                                          int B, int Di, int N, int dtype, vmb_stream_t stream) {
  using namespace vmb;
  VMB_CHECK_ARG(state && x && dt && A2 && Bm && Cm && y, "state_update: null pointer");
  VMB_CHECK_ARG(dtype_ok(dtype) && dtype_ok(state_dtype), "state_update: bad dtype");
  VMB_CHECK_ARG(B <= 65535, "state_update: batch %d > 65535", B);
  if (B <= 0) return VMB_OK;
  dim3 grid((Di + 127) / 128, B);
  cudaStream_t st = as_stream(stream);
  if (dtype == VMB_F32)
    state_update_kernel<float, true><<<grid, 128, 0, st>>>(
        state, state_dtype, (const float*)x, x_bs, (const float*)dt, dt_bs, A2, (const float*)Bm,
        b_bs, (const float*)Cm, c_bs, D, (const float*)z, z_bs, dt_bias, softplus, (float*)y, y_bs,
        Di, N);
  else
    state_update_kernel<__nv_bfloat16, false><<<grid, 128, 0, st>>>(
        state, state_dtype, (const __nv_bfloat16*)x, x_bs, (const __nv_bfloat16*)dt, dt_bs, A2,
        (const __nv_bfloat16*)Bm, b_bs, (const __nv_bfloat16*)Cm, c_bs, D,
        (const __nv_bfloat16*)z, z_bs, dt_bias, softplus, (__nv_bfloat16*)y, y_bs, Di, N);
  VMB_LAUNCH_CHECK("state_update_kernel");
  return VMB_OK;
}
