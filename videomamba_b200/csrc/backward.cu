// Backward kernels of the block's non-scan operators (the reference
// advertises a differentiable state carry, scripts/check_streaming_state.py:57-60, and wraps the
// mixer in activation checkpointing, models/videomamba/videomamba.py:168-206).
//
//   * add_norm_bwd:  backward of the fused residual add + RMSNorm / LayerNorm
//                    (forward: addnorm.cu; reference call sites videomamba.py:151-166, :902-918)
//   * conv1d_bwd:    backward of the depthwise causal conv + SiLU with streaming history
//                    (forward: conv1d.cu; reference mamba_simple.py:381-404), including the gradient
//                    that arrives through the returned conv state and the one that leaves through
//                    the incoming conv state
//   * transpose / column-sum helpers: the projection backward is two calls of the forward GEMM
//     (dX = dY * W, dW = dY^T * X) on transposed operands plus a column sum for the bias.
//
// Every reduction over rows (dw, db) is two-stage: per-CTA partials in a caller-provided workspace,
// then one reduce launch -- no atomics, results are run-to-run deterministic.  All math is fp32;
// element types are resolved at run time (these kernels are not on the forward hot path).
#include <algorithm>
#include <type_traits>

#include "internal.h"

namespace vmb {
namespace {

// ---- transpose: in (rows, cols) with row stride ld -> out (cols, rows) with row stride ldo ----------
template <typename T>
__global__ void __launch_bounds__(256)
transpose_kernel(const T* __restrict__ in, int64_t ld, T* __restrict__ out, int64_t ldo, int64_t rows, int cols) {
  __shared__ T tile[32][33];
  const int64_t r0 = (int64_t)blockIdx.x * 32;            // row tiles on grid.x (rows can be millions)
  const int c0 = blockIdx.y * 32;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;   // 32 x 8
#pragma unroll
  for (int i = 0; i < 32; i += 8) {
    const int64_t r = r0 + ty + i;
    const int c = c0 + tx;
    if (r < rows && c < cols) tile[ty + i][tx] = in[r * ld + c];
  }
  __syncthreads();
#pragma unroll
  for (int i = 0; i < 32; i += 8) {
    const int c = c0 + ty + i;
    const int64_t r = r0 + tx;
    if (r < rows && c < cols) out[(int64_t)c * ldo + r] = tile[tx][ty + i];
  }
}

// ---- partial[P][n] (fp32) -> out[n] -------------------------------------------------------------
__global__ void reduce_partials_kernel(const float* __restrict__ partial, int P, int64_t n,
                                       void* __restrict__ out, int out_dtype) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  float s = 0.f;
  for (int p = 0; p < P; ++p) s += partial[(int64_t)p * n + i];
  store_from_f32(out, i, out_dtype, s);
}

// Many partial rows, few columns (per-CTA partials of the conv / add-norm backward: P in the hundreds, n a few
// thousand): CTA = 32 columns x 8 row lanes, fixed summation order (deterministic).  In-place use (out == row 0
// of partial) is safe: a column is read and written by one CTA only, the write follows the barrier.
__global__ void __launch_bounds__(256)
reduce_partials_tall_kernel(const float* __restrict__ partial, int P, int64_t n, void* __restrict__ out,
                            int out_dtype) {
  __shared__ float red[8][33];
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  const int64_t c = (int64_t)blockIdx.x * 32 + tx;
  float s = 0.f;
  if (c < n)
    for (int p = ty; p < P; p += 8) s += partial[(int64_t)p * n + c];
  red[ty][tx] = s;
  __syncthreads();
  if (ty == 0 && c < n) {
    float t = 0.f;
#pragma unroll
    for (int j = 0; j < 8; ++j) t += red[j][tx];
    store_from_f32(out, c, out_dtype, t);
  }
}

// ---- column sums of x (M, N) with row stride ld: CTA = 32 columns x 8 row lanes over a row chunk --
__global__ void __launch_bounds__(256)
colsum_partial_kernel(const void* __restrict__ x, int dtype, int64_t ld, int64_t M, int N,
                      int64_t rows_per_chunk, float* __restrict__ partial) {
  __shared__ float red[8][33];
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  const int c = blockIdx.x * 32 + tx;
  const int64_t r0 = (int64_t)blockIdx.y * rows_per_chunk;
  const int64_t r1 = min(M, r0 + rows_per_chunk);
  float s = 0.f;
  if (c < N)
    for (int64_t r = r0 + ty; r < r1; r += 8) s += load_as_f32(x, r * ld + c, dtype);
  red[ty][tx] = s;
  __syncthreads();
  if (ty == 0 && c < N) {
#pragma unroll
    for (int j = 1; j < 8; ++j) s += red[j][tx];
    partial[(int64_t)blockIdx.y * N + c] = s;
  }
}

// ---- add + norm backward -------------------------------------------------------------------------
// One warp per row, warps stride over the rows; a lane owns columns lane, lane + 32, ...
// dacc = rstd * (w*dy - xhat * mean(w*dy*xhat) [- mean(w*dy) for LayerNorm]) + d(residual_out).
constexpr int kBwdWarps = 8;

template <int kIters, bool kRms>
__global__ void __launch_bounds__(kBwdWarps * 32)
add_norm_bwd_kernel(const void* __restrict__ x, int x_dtype, int64_t ldx, const void* __restrict__ residual,
                    int res_dtype, const void* __restrict__ weight, int w_dtype,
                    const void* __restrict__ dy, const void* __restrict__ dres_out, int dres_out_dtype,
                    void* __restrict__ dx, void* __restrict__ dres, float* __restrict__ partial_w,
                    float* __restrict__ partial_b, int64_t rows, int dim, float eps) {
  __shared__ float red[kBwdWarps][kIters * 32];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  float w[kIters], dw[kIters], db[kIters];
#pragma unroll
  for (int it = 0; it < kIters; ++it) {
    const int c = it * 32 + lane;
    w[it] = c < dim ? load_as_f32(weight, c, w_dtype) : 0.f;
    dw[it] = db[it] = 0.f;
  }
  const float inv_dim = 1.f / (float)dim;
  for (int64_t row = (int64_t)blockIdx.x * kBwdWarps + warp; row < rows; row += (int64_t)gridDim.x * kBwdWarps) {
    float v[kIters], g[kIters];
    float sum = 0.f, sumsq = 0.f;
#pragma unroll
    for (int it = 0; it < kIters; ++it) {
      const int c = it * 32 + lane;
      v[it] = g[it] = 0.f;
      if (c < dim) {
        v[it] = load_as_f32(x, row * ldx + c, x_dtype);
        if (residual != nullptr) v[it] += load_as_f32(residual, row * (int64_t)dim + c, res_dtype);
        g[it] = load_as_f32(dy, row * (int64_t)dim + c, x_dtype);
        sum += v[it];
        sumsq += v[it] * v[it];
      }
    }
    float mean = 0.f, rstd;
    if constexpr (kRms) {
      rstd = rsqrtf(warp_sum(sumsq) * inv_dim + eps);
    } else {
      mean = warp_sum(sum) * inv_dim;
      float var = 0.f;
#pragma unroll
      for (int it = 0; it < kIters; ++it)
        if (it * 32 + lane < dim) var += (v[it] - mean) * (v[it] - mean);
      rstd = rsqrtf(warp_sum(var) * inv_dim + eps);
    }
    float c1 = 0.f, c2 = 0.f;                       // mean(w*dy*xhat), mean(w*dy)
#pragma unroll
    for (int it = 0; it < kIters; ++it) {
      const float xhat = (v[it] - mean) * rstd;     // 0 beyond dim (v = mean = 0 there only for RMS; masked below)
      if (it * 32 + lane < dim) {
        dw[it] += g[it] * xhat;
        db[it] += g[it];
        const float wg = w[it] * g[it];
        c1 += wg * xhat;
        c2 += wg;
        v[it] = xhat;
        g[it] = wg;
      }
    }
    c1 = warp_sum(c1) * inv_dim;
    c2 = kRms ? 0.f : warp_sum(c2) * inv_dim;
#pragma unroll
    for (int it = 0; it < kIters; ++it) {
      const int c = it * 32 + lane;
      if (c < dim) {
        float d = rstd * (g[it] - v[it] * c1 - c2);
        if (dres_out != nullptr) d += load_as_f32(dres_out, row * (int64_t)dim + c, dres_out_dtype);
        store_from_f32(dx, row * (int64_t)dim + c, x_dtype, d);
        if (dres != nullptr) store_from_f32(dres, row * (int64_t)dim + c, res_dtype, d);
      }
    }
  }
  // CTA partials of dw / db
#pragma unroll
  for (int pass = 0; pass < 2; ++pass) {
    __syncthreads();
#pragma unroll
    for (int it = 0; it < kIters; ++it) red[warp][it * 32 + lane] = pass ? db[it] : dw[it];
    __syncthreads();
    float* dst = pass ? partial_b : partial_w;
    if (dst != nullptr)
      for (int c = threadIdx.x; c < dim; c += kBwdWarps * 32) {
        float s = 0.f;
#pragma unroll
        for (int j = 0; j < kBwdWarps; ++j) s += red[j][c];
        dst[(int64_t)blockIdx.x * dim + c] = s;
      }
  }
}

// Vectorised variant for the production type combinations (x / dy / dx / weight of type T, residual stream
// and its gradients fp32): a lane owns 4 consecutive columns per step, 8-byte (bf16) / 16-byte (fp32) accesses.
template <typename T> struct V4;
template <> struct V4<float> {
  static __device__ __forceinline__ void ld(const float* p, float (&f)[4]) {
    const float4 v = *reinterpret_cast<const float4*>(p);
    f[0] = v.x; f[1] = v.y; f[2] = v.z; f[3] = v.w;
  }
  static __device__ __forceinline__ void st(float* p, const float (&f)[4]) {
    *reinterpret_cast<float4*>(p) = make_float4(f[0], f[1], f[2], f[3]);
  }
};
template <> struct V4<__nv_bfloat16> {
  static __device__ __forceinline__ void ld(const __nv_bfloat16* p, float (&f)[4]) {
    const uint2 v = *reinterpret_cast<const uint2*>(p);
    f[0] = __uint_as_float(v.x << 16); f[1] = __uint_as_float(v.x & 0xffff0000u);
    f[2] = __uint_as_float(v.y << 16); f[3] = __uint_as_float(v.y & 0xffff0000u);
  }
  static __device__ __forceinline__ void st(__nv_bfloat16* p, const float (&f)[4]) {
    __nv_bfloat162 a = __floats2bfloat162_rn(f[0], f[1]), b = __floats2bfloat162_rn(f[2], f[3]);
    uint2 r;
    r.x = *reinterpret_cast<uint32_t*>(&a);
    r.y = *reinterpret_cast<uint32_t*>(&b);
    *reinterpret_cast<uint2*>(p) = r;
  }
};

template <typename T, int kIters, bool kRms>        // kIters * 128 >= dim, dim % 4 == 0
__global__ void __launch_bounds__(kBwdWarps * 32)
add_norm_bwd_vec_kernel(const T* __restrict__ x, int64_t ldx, const float* __restrict__ residual,
                        const T* __restrict__ weight, const T* __restrict__ dy,
                        const float* __restrict__ dres_out, T* __restrict__ dx, float* __restrict__ dres,
                        float* __restrict__ partial_w, float* __restrict__ partial_b, int64_t rows, int dim,
                        float eps) {
  __shared__ float red[kBwdWarps][kIters * 128];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int nvec = dim >> 2;
  float w[kIters][4], dw[kIters][4], db[kIters][4];
#pragma unroll
  for (int it = 0; it < kIters; ++it) {
    const int c = it * 32 + lane;
#pragma unroll
    for (int j = 0; j < 4; ++j) w[it][j] = dw[it][j] = db[it][j] = 0.f;
    if (c < nvec) V4<T>::ld(weight + 4 * c, w[it]);
  }
  const float inv_dim = 1.f / (float)dim;
  for (int64_t row = (int64_t)blockIdx.x * kBwdWarps + warp; row < rows; row += (int64_t)gridDim.x * kBwdWarps) {
    float v[kIters][4], g[kIters][4], ro[kIters][4];
    float sum = 0.f, sumsq = 0.f;
    // every load of the row is issued up front (the incoming residual gradient too: loading it after the two
    // reductions would give the row a second, dependent trip to memory)
#pragma unroll
    for (int it = 0; it < kIters; ++it) {
      const int c = it * 32 + lane;
#pragma unroll
      for (int j = 0; j < 4; ++j) v[it][j] = g[it][j] = ro[it][j] = 0.f;
      if (c < nvec) {
        V4<T>::ld(x + row * ldx + 4 * c, v[it]);
        V4<T>::ld(dy + row * (int64_t)dim + 4 * c, g[it]);
        if (dres_out != nullptr) V4<float>::ld(dres_out + row * (int64_t)dim + 4 * c, ro[it]);
      }
    }
#pragma unroll
    for (int it = 0; it < kIters; ++it) {
      const int c = it * 32 + lane;
      if (c < nvec) {
        if (residual != nullptr) {
          float r[4];
          V4<float>::ld(residual + row * (int64_t)dim + 4 * c, r);
#pragma unroll
          for (int j = 0; j < 4; ++j) v[it][j] += r[j];
        }
#pragma unroll
        for (int j = 0; j < 4; ++j) { sum += v[it][j]; sumsq += v[it][j] * v[it][j]; }
      }
    }
    float mean = 0.f, rstd;
    if constexpr (kRms) {
      rstd = rsqrtf(warp_sum(sumsq) * inv_dim + eps);
    } else {
      mean = warp_sum(sum) * inv_dim;
      float var = 0.f;
#pragma unroll
      for (int it = 0; it < kIters; ++it)
        if (it * 32 + lane < nvec)
#pragma unroll
          for (int j = 0; j < 4; ++j) var += (v[it][j] - mean) * (v[it][j] - mean);
      rstd = rsqrtf(warp_sum(var) * inv_dim + eps);
    }
    float c1 = 0.f, c2 = 0.f;
#pragma unroll
    for (int it = 0; it < kIters; ++it)
      if (it * 32 + lane < nvec)
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const float xhat = (v[it][j] - mean) * rstd;
          dw[it][j] += g[it][j] * xhat;
          db[it][j] += g[it][j];
          const float wg = w[it][j] * g[it][j];
          c1 += wg * xhat;
          c2 += wg;
          v[it][j] = xhat;
          g[it][j] = wg;
        }
    c1 = warp_sum(c1) * inv_dim;
    c2 = kRms ? 0.f : warp_sum(c2) * inv_dim;
#pragma unroll
    for (int it = 0; it < kIters; ++it) {
      const int c = it * 32 + lane;
      if (c < nvec) {
        float d[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) d[j] = rstd * (g[it][j] - v[it][j] * c1 - c2) + ro[it][j];
        V4<T>::st(dx + row * (int64_t)dim + 4 * c, d);
        if (dres != nullptr) V4<float>::st(dres + row * (int64_t)dim + 4 * c, d);
      }
    }
  }
#pragma unroll
  for (int pass = 0; pass < 2; ++pass) {
    __syncthreads();
#pragma unroll
    for (int it = 0; it < kIters; ++it)
#pragma unroll
      for (int j = 0; j < 4; ++j) red[warp][(it * 32 + lane) * 4 + j] = pass ? db[it][j] : dw[it][j];
    __syncthreads();
    float* dst = pass ? partial_b : partial_w;
    if (dst != nullptr)
      for (int c = threadIdx.x; c < dim; c += kBwdWarps * 32) {
        float s = 0.f;
#pragma unroll
        for (int j = 0; j < kBwdWarps; ++j) s += red[j][c];
        dst[(int64_t)blockIdx.x * dim + c] = s;
      }
  }
}

template <typename T>
bool launch_add_norm_bwd_vec(const void* x, int64_t ldx, const void* residual, const void* weight, const void* dy,
                             const void* dres_out, void* dx, void* dres, float* pw, float* pb, int64_t rows,
                             int dim, float eps, bool rms, int ctas, cudaStream_t st) {
#define VMB_ANV(IT)                                                                                         \
  do {                                                                                                      \
    if (rms)                                                                                                \
      add_norm_bwd_vec_kernel<T, IT, true><<<ctas, kBwdWarps * 32, 0, st>>>(                                \
          (const T*)x, ldx, (const float*)residual, (const T*)weight, (const T*)dy, (const float*)dres_out, \
          (T*)dx, (float*)dres, pw, pb, rows, dim, eps);                                                    \
    else                                                                                                    \
      add_norm_bwd_vec_kernel<T, IT, false><<<ctas, kBwdWarps * 32, 0, st>>>(                               \
          (const T*)x, ldx, (const float*)residual, (const T*)weight, (const T*)dy, (const float*)dres_out, \
          (T*)dx, (float*)dres, pw, pb, rows, dim, eps);                                                    \
  } while (0)
  if (dim <= 128 * 2) VMB_ANV(2);
  else if (dim <= 128 * 3) VMB_ANV(3);
  else if (dim <= 128 * 5) VMB_ANV(5);
  else if (dim <= 128 * 9) VMB_ANV(9);
  else return false;
#undef VMB_ANV
  return true;
}

// ---- causal conv + SiLU backward -------------------------------------------------------------------
// Thread = channel; a CTA walks one chunk of kConvChunk tokens of one sequence in order, carrying the last
// W inputs and the last W pre-activation gradients in registers.  hist[j], j in [-(W-1), L): x[j] for
// j >= 0, conv_state_in[W + j] for j < 0 (zeros without a state).
constexpr int kConvChunk = 32;
constexpr int kConvWMax = 8;

__global__ void __launch_bounds__(128)
conv1d_bwd_kernel(const void* __restrict__ x, int64_t x_bs, int64_t x_ts, const void* __restrict__ weight,
                  const void* __restrict__ bias, const void* __restrict__ cs_in, int cs_in_dtype,
                  const void* __restrict__ dy, const void* __restrict__ dcs_out, int dcs_out_dtype,
                  void* __restrict__ dx, int64_t dx_bs, int64_t dx_ts, void* __restrict__ dcs_in,
                  float* __restrict__ partial, int B, int L, int Di, int W, int silu, int dtype) {
  const int d = blockIdx.x * 128 + threadIdx.x;
  const int b = blockIdx.y;
  const int chunk = blockIdx.z;
  if (d >= Di) return;
  const int c0 = chunk * kConvChunk;
  const int c1 = min(L, c0 + kConvChunk);           // this CTA owns dx[c0, c1) (and the history for chunk 0)
  float w[kConvWMax], hist[kConvWMax], dpre[kConvWMax], dwa[kConvWMax];
#pragma unroll
  for (int k = 0; k < kConvWMax; ++k) {
    w[k] = k < W ? load_as_f32(weight, (int64_t)d * W + k, dtype) : 0.f;
    hist[k] = dpre[k] = dwa[k] = 0.f;
  }
  const float bv = bias ? load_as_f32(bias, d, dtype) : 0.f;
  float dba = 0.f;
  auto hist_at = [&](int j) -> float {               // j in [-(W-1), L)
    if (j >= 0) return load_as_f32(x, (int64_t)b * x_bs + (int64_t)j * x_ts + d, dtype);
    return cs_in ? load_as_f32(cs_in, ((int64_t)b * Di + d) * W + (W + j), cs_in_dtype) : 0.f;
  };
  auto dcs_out_at = [&](int i) -> float {            // gradient of conv_state_out[b, d, i]
    return dcs_out ? load_as_f32(dcs_out, ((int64_t)b * Di + d) * W + i, dcs_out_dtype) : 0.f;
  };
  // hist[k] holds hist_at(l - (W-1) + k) for the current l; prime with the W-1 values before c0
  for (int k = 1; k < W; ++k) hist[k] = hist_at(c0 - W + k);
  // dpre window: dpre[k] = dpre(l - (W-1) + k).  Tokens [c0, c1 + W - 1) are visited: the last W - 1 only
  // complete the gradients of this chunk's inputs (their own dw / db belong to the next chunk).
  const int lend = min(L, c1 + W - 1);
  for (int l = c0; l < c1 + W - 1; ++l) {
    // shift the windows
    for (int k = 0; k + 1 < W; ++k) { hist[k] = hist[k + 1]; dpre[k] = dpre[k + 1]; }
    float g = 0.f;
    if (l < lend) {
      hist[W - 1] = hist_at(l);
      float pre = bv;
      for (int k = 0; k < W; ++k) pre = fmaf(w[k], hist[k], pre);
      g = load_as_f32(dy, ((int64_t)b * L + l) * Di + d, dtype);
      if (silu) {
        const float s = 1.f / (1.f + expf(-pre));
        g *= s * (1.f + pre * (1.f - s));
      }
      if (l < c1) {
        for (int k = 0; k < W; ++k) dwa[k] = fmaf(g, hist[k], dwa[k]);
        dba += g;
      }
    }
    dpre[W - 1] = g;
    // input j = l - (W-1) now has all its consumers: d hist[j] = sum_i w[W-1-i] * dpre(j + i)
    const int j = l - (W - 1);
    if (j >= c0 || (chunk == 0 && j >= -(W - 1))) {
      float s = 0.f;
      for (int i = 0; i < W; ++i) s = fmaf(w[W - 1 - i], dpre[i], s);
      // dpre entries for tokens < 0 are zero by construction (the window starts empty)
      if (j >= 0) {
        if (j < c1) {
          if (j >= L - W) s += dcs_out_at(j - (L - W));
          store_from_f32(dx, (int64_t)b * dx_bs + (int64_t)j * dx_ts + d, dtype, s);
        }
      } else if (dcs_in != nullptr) {
        const int si = W + j;                         // history slot
        if (si >= L) s += dcs_out_at(si - L);         // L < W: old history is still part of the new state
        store_from_f32(dcs_in, ((int64_t)b * Di + d) * W + si, cs_in_dtype, s);
      }
    }
  }
  if (chunk == 0 && dcs_in != nullptr) {              // slot 0 never reaches the conv; it may reach the new state
    float s = 0.f;
    if (0 >= L) s += dcs_out_at(0 - L);
    store_from_f32(dcs_in, ((int64_t)b * Di + d) * W, cs_in_dtype, s);
  }
  // partial[(b * nchunks + chunk)][d][W + 1]
  float* p = partial + (((int64_t)b * gridDim.z + chunk) * Di + d) * (W + 1);
  for (int k = 0; k < W; ++k) p[k] = dwa[k];
  p[W] = dba;
}

// Fast variant: compile-time W (windows stay in registers), a thread owns TWO adjacent channels (4-byte bf16 /
// 8-byte fp32 accesses), fast exp.  Same walk and the same partial layout as conv1d_bwd_kernel.
template <typename T> struct Pair;
template <> struct Pair<float> {
  static __device__ __forceinline__ float2 ld(const float* p) { return *reinterpret_cast<const float2*>(p); }
  static __device__ __forceinline__ void st(float* p, float2 v) { *reinterpret_cast<float2*>(p) = v; }
};
template <> struct Pair<__nv_bfloat16> {
  static __device__ __forceinline__ float2 ld(const __nv_bfloat16* p) {
    const uint32_t v = *reinterpret_cast<const uint32_t*>(p);
    return make_float2(__uint_as_float(v << 16), __uint_as_float(v & 0xffff0000u));
  }
  static __device__ __forceinline__ void st(__nv_bfloat16* p, float2 v) {
    __nv_bfloat162 r = __floats2bfloat162_rn(v.x, v.y);
    *reinterpret_cast<uint32_t*>(p) = *reinterpret_cast<uint32_t*>(&r);
  }
};

// kStaged (bf16, 16-byte aligned rows): the x and dy rows of the chunk are brought into shared memory with
// 16-byte cp.async first (38 + 35 rows x 512 B: ~37 KB per CTA, six CTAs per SM), the walk then reads
// shared memory.  Without it every token of the dependent walk exposes a global-load latency and 28 resident
// warps x two 128-byte requests keep only ~7 KB in flight per SM (1.5 TB/s).
template <typename T, int W, bool kStaged>
__global__ void __launch_bounds__(128)
conv1d_bwd_pair_kernel(const T* __restrict__ x, int64_t x_bs, int64_t x_ts, const T* __restrict__ weight,
                       const T* __restrict__ bias, const void* __restrict__ cs_in, int cs_in_dtype,
                       const T* __restrict__ dy, const void* __restrict__ dcs_out, int dcs_out_dtype,
                       T* __restrict__ dx, int64_t dx_bs, int64_t dx_ts, void* __restrict__ dcs_in,
                       float* __restrict__ partial, int L, int Di, int silu) {
  const int d = (blockIdx.x * 128 + threadIdx.x) * 2;
  const int b = blockIdx.y;
  const int chunk = blockIdx.z;
  if (!kStaged && d >= Di) return;
  const int dw_ = min(d, Di - 2);                     // (staged: every thread helps with the copies first)
  const int c0 = chunk * kConvChunk;
  const int c1 = min(L, c0 + kConvChunk);
  float2 w[W], hist[W], dpre[W], dwa[W];
#pragma unroll
  for (int k = 0; k < W; ++k) {
    w[k] = make_float2(to_f32<T>(weight[(int64_t)dw_ * W + k]), to_f32<T>(weight[(int64_t)(dw_ + 1) * W + k]));
    hist[k] = dpre[k] = dwa[k] = make_float2(0.f, 0.f);
  }
  const float2 bv = bias ? Pair<T>::ld(bias + dw_) : make_float2(0.f, 0.f);
  float2 dba = make_float2(0.f, 0.f);
  const T* xp = x + (int64_t)b * x_bs + d;
  const T* gp = dy + (int64_t)b * L * Di + d;
  T* op = dx + (int64_t)b * dx_bs + d;
  // staged rows: sx[r] = x row max(c0 - (W-1), 0) + r, sg[r] = dy row c0 + r, 256 channels (512 B) per row
  extern __shared__ __align__(16) uint8_t conv_smem[];
  constexpr int kRowsMax = kConvChunk + 2 * (W - 1);  // x rows [c0 - (W-1), c1 + W - 1); dy rows [c0, c1 + W - 1)
  const int xr0 = max(c0 - (W - 1), 0);
  const T* sx = reinterpret_cast<const T*>(conv_smem) + 2 * threadIdx.x;
  const T* sg = sx + kRowsMax * 256;
  if constexpr (kStaged) {
    const int d0 = blockIdx.x * 256;
    const int xrows = min(L, c1 + W - 1) - xr0, grows = min(L, c1 + W - 1) - c0;
    const uint32_t sb = static_cast<uint32_t>(__cvta_generic_to_shared(conv_smem));
    for (int i = threadIdx.x; i < xrows * 32; i += 128) {
      const int r = i >> 5, ch = (i & 31) * 8;
      if (d0 + ch < Di) {
        const T* src = x + (int64_t)b * x_bs + (int64_t)(xr0 + r) * x_ts + d0 + ch;
        asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(sb + (r * 256 + ch) * 2), "l"(src) : "memory");
      }
    }
    for (int i = threadIdx.x; i < grows * 32; i += 128) {
      const int r = i >> 5, ch = (i & 31) * 8;
      if (d0 + ch < Di) {
        const T* src = dy + ((int64_t)b * L + c0 + r) * Di + d0 + ch;
        asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(sb + ((kRowsMax + r) * 256 + ch) * 2), "l"(src) : "memory");
      }
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
    asm volatile("cp.async.wait_group 0;" ::: "memory");
    __syncthreads();
    if (d >= Di) return;
  }
  auto hist_at = [&](int j) -> float2 {
    if (kStaged && j >= 0) return Pair<T>::ld(sx + (j - xr0) * 256);
    if (j >= 0) return Pair<T>::ld(xp + (int64_t)j * x_ts);
    if (!cs_in) return make_float2(0.f, 0.f);
    const int64_t o = ((int64_t)b * Di + d) * W + (W + j);
    return make_float2(load_as_f32(cs_in, o, cs_in_dtype), load_as_f32(cs_in, o + W, cs_in_dtype));
  };
  auto dcs_out_at = [&](int i) -> float2 {
    if (!dcs_out) return make_float2(0.f, 0.f);
    const int64_t o = ((int64_t)b * Di + d) * W + i;
    return make_float2(load_as_f32(dcs_out, o, dcs_out_dtype), load_as_f32(dcs_out, o + W, dcs_out_dtype));
  };
#pragma unroll
  for (int k = 1; k < W; ++k) hist[k] = hist_at(c0 - W + k);
  const int lend = min(L, c1 + W - 1);
#pragma unroll 4
  for (int l = c0; l < c1 + W - 1; ++l) {
#pragma unroll
    for (int k = 0; k + 1 < W; ++k) { hist[k] = hist[k + 1]; dpre[k] = dpre[k + 1]; }
    float2 g = make_float2(0.f, 0.f);
    if (l < lend) {
      hist[W - 1] = hist_at(l);
      float2 pre = bv;
#pragma unroll
      for (int k = 0; k < W; ++k) { pre.x = fmaf(w[k].x, hist[k].x, pre.x); pre.y = fmaf(w[k].y, hist[k].y, pre.y); }
      g = kStaged ? Pair<T>::ld(sg + (l - c0) * 256) : Pair<T>::ld(gp + (int64_t)l * Di);
      if (silu) {
        // bf16: sigmoid from one MUFU.EX2 + one MUFU.RCP (the IEEE division and __expf's range handling were a
        // quarter of the loop's instructions); fp32 keeps the accurate form
        float sgx, sgy;
        if constexpr (std::is_same<T, float>::value) {
          sgx = 1.f / (1.f + __expf(-pre.x));
          sgy = 1.f / (1.f + __expf(-pre.y));
        } else {
          sgx = rcp_approx(1.f + ex2_approx(-kLog2e * pre.x));
          sgy = rcp_approx(1.f + ex2_approx(-kLog2e * pre.y));
        }
        g.x *= sgx * (1.f + pre.x * (1.f - sgx));
        g.y *= sgy * (1.f + pre.y * (1.f - sgy));
      }
      if (l < c1) {
#pragma unroll
        for (int k = 0; k < W; ++k) { dwa[k].x = fmaf(g.x, hist[k].x, dwa[k].x); dwa[k].y = fmaf(g.y, hist[k].y, dwa[k].y); }
        dba.x += g.x; dba.y += g.y;
      }
    }
    dpre[W - 1] = g;
    const int j = l - (W - 1);
    if (j >= c0 || (chunk == 0 && j >= -(W - 1))) {
      float2 s = make_float2(0.f, 0.f);
#pragma unroll
      for (int i = 0; i < W; ++i) { s.x = fmaf(w[W - 1 - i].x, dpre[i].x, s.x); s.y = fmaf(w[W - 1 - i].y, dpre[i].y, s.y); }
      if (j >= 0) {
        if (j < c1) {
          if (j >= L - W) { const float2 e = dcs_out_at(j - (L - W)); s.x += e.x; s.y += e.y; }
          Pair<T>::st(op + (int64_t)j * dx_ts, s);
        }
      } else if (dcs_in != nullptr) {
        const int si = W + j;
        if (si >= L) { const float2 e = dcs_out_at(si - L); s.x += e.x; s.y += e.y; }
        const int64_t o = ((int64_t)b * Di + d) * W + si;
        store_from_f32(dcs_in, o, cs_in_dtype, s.x);
        store_from_f32(dcs_in, o + W, cs_in_dtype, s.y);
      }
    }
  }
  if (chunk == 0 && dcs_in != nullptr) {              // slot 0 never reaches the conv (L >= 1 here)
    const int64_t o = ((int64_t)b * Di + d) * W;
    store_from_f32(dcs_in, o, cs_in_dtype, 0.f);
    store_from_f32(dcs_in, o + W, cs_in_dtype, 0.f);
  }
  float* p = partial + (((int64_t)b * gridDim.z + chunk) * Di + d) * (W + 1);
#pragma unroll
  for (int k = 0; k < W; ++k) { p[k] = dwa[k].x; p[W + 1 + k] = dwa[k].y; }
  p[W] = dba.x;
  p[2 * W + 1] = dba.y;
}

template <typename T, int W>
void launch_conv_bwd_pair(const void* x, int64_t x_bs, int64_t x_ts, const void* weight, const void* bias,
                          const void* cs_in, int cs_in_dtype, const void* dy, const void* dcs_out, int dcs_out_dtype,
                          void* dx, int64_t dx_bs, int64_t dx_ts, void* dcs_in, float* partial, int B, int L, int Di,
                          int silu, int chunks, cudaStream_t st) {
  auto al16 = [](const void* p) { return reinterpret_cast<uintptr_t>(p) % 16 == 0; };
  if (std::is_same<T, __nv_bfloat16>::value && Di % 8 == 0 && x_bs % 8 == 0 && x_ts % 8 == 0 && al16(x) && al16(dy)) {
    constexpr int smem = 2 * (kConvChunk + 2 * (W - 1)) * 256 * 2;
    static bool attr_set = false;
    if (!attr_set) {
      cudaFuncSetAttribute(conv1d_bwd_pair_kernel<T, W, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
      attr_set = true;
    }
    conv1d_bwd_pair_kernel<T, W, true><<<dim3((Di / 2 + 127) / 128, B, chunks), 128, smem, st>>>(
        (const T*)x, x_bs, x_ts, (const T*)weight, (const T*)bias, cs_in, cs_in_dtype, (const T*)dy, dcs_out,
        dcs_out_dtype, (T*)dx, dx_bs, dx_ts, dcs_in, partial, L, Di, silu);
    return;
  }
  conv1d_bwd_pair_kernel<T, W, false><<<dim3((Di / 2 + 127) / 128, B, chunks), 128, 0, st>>>(
      (const T*)x, x_bs, x_ts, (const T*)weight, (const T*)bias, cs_in, cs_in_dtype, (const T*)dy, dcs_out,
      dcs_out_dtype, (T*)dx, dx_bs, dx_ts, dcs_in, partial, L, Di, silu);
}

}  // namespace

int reduce_partials(const float* partial, int P, int64_t n, void* out, int out_dtype, cudaStream_t st) {
  if (n <= 0) return VMB_OK;
  if (P >= 32 && n <= (int64_t)P * 4096) {        // tall and narrow: one thread per column would walk P rows alone
    reduce_partials_tall_kernel<<<(unsigned)((n + 31) / 32), 256, 0, st>>>(partial, P, n, out, out_dtype);
    VMB_LAUNCH_CHECK("reduce_partials_tall_kernel");
    return VMB_OK;
  }
  reduce_partials_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(partial, P, n, out, out_dtype);
  VMB_LAUNCH_CHECK("reduce_partials_kernel");
  return VMB_OK;
}

int64_t colsum_chunks(int64_t M) { return std::max<int64_t>(1, std::min<int64_t>(512, (M + 511) / 512)); }

}  // namespace vmb

using namespace vmb;

extern "C" int vmb_transpose_2d(const void* in, int64_t ld, void* out, int64_t ldo, int64_t rows, int cols,
                                int dtype, vmb_stream_t stream) {
  VMB_CHECK_ARG(dtype_ok(dtype), "transpose: bad dtype %d", dtype);
  VMB_CHECK_ARG(rows >= 0 && cols >= 0 && ld >= cols && ldo >= rows, "transpose: bad sizes");
  if (rows == 0 || cols == 0) return VMB_OK;
  VMB_CHECK_ARG(in && out, "transpose: null pointer");
  VMB_CHECK_ARG((cols + 31) / 32 <= 65535 && (rows + 31) / 32 <= 2147483647ll, "transpose: matrix too large");
  cudaStream_t st = as_stream(stream);
  dim3 grid((unsigned)((rows + 31) / 32), (unsigned)((cols + 31) / 32));
  if (dtype == VMB_F32)
    transpose_kernel<float><<<grid, 256, 0, st>>>(reinterpret_cast<const float*>(in), ld,
                                                  reinterpret_cast<float*>(out), ldo, rows, cols);
  else
    transpose_kernel<__nv_bfloat16><<<grid, 256, 0, st>>>(reinterpret_cast<const __nv_bfloat16*>(in), ld,
                                                          reinterpret_cast<__nv_bfloat16*>(out), ldo, rows, cols);
  VMB_LAUNCH_CHECK("transpose_kernel");
  return VMB_OK;
}

extern "C" int64_t vmb_colsum_workspace_bytes(int64_t M, int N) {
  if (M <= 0 || N <= 0) return 0;
  return colsum_chunks(M) * N * (int64_t)sizeof(float);
}

extern "C" int vmb_colsum(const void* x, int64_t ld, int64_t M, int N, int dtype, void* out, int out_dtype,
                          void* workspace, int64_t workspace_bytes, vmb_stream_t stream) {
  VMB_CHECK_ARG(dtype_ok(dtype) && dtype_ok(out_dtype), "colsum: bad dtype");
  VMB_CHECK_ARG(M >= 0 && N > 0 && ld >= N, "colsum: bad sizes");
  VMB_CHECK_ARG(out != nullptr, "colsum: null output");
  cudaStream_t st = as_stream(stream);
  if (M == 0) {
    VMB_CUDA(cudaMemsetAsync(out, 0, (size_t)N * dtype_size(out_dtype), st));
    return VMB_OK;
  }
  VMB_CHECK_ARG(x != nullptr, "colsum: null input");
  const int64_t P = colsum_chunks(M);
  VMB_CHECK_ARG(workspace && workspace_bytes >= P * N * (int64_t)sizeof(float), "colsum: workspace too small");
  const int64_t per = (M + P - 1) / P;
  float* partial = reinterpret_cast<float*>(workspace);
  colsum_partial_kernel<<<dim3((N + 31) / 32, (unsigned)P), 256, 0, st>>>(x, dtype, ld, M, N, per, partial);
  VMB_LAUNCH_CHECK("colsum_partial_kernel");
  { const int rc = reduce_partials(partial, (int)P, N, out, out_dtype, st); if (rc != VMB_OK) return rc; }
  return VMB_OK;
}

namespace {
int add_norm_bwd_ctas(int64_t rows) {
  return (int)std::max<int64_t>(1, std::min<int64_t>((rows + kBwdWarps - 1) / kBwdWarps, 4ll * sm_count()));
}
}  // namespace

extern "C" int64_t vmb_add_norm_bwd_workspace_bytes(int64_t rows, int dim) {
  if (rows <= 0 || dim <= 0) return 0;
  return 2ll * add_norm_bwd_ctas(rows) * dim * (int64_t)sizeof(float);
}

extern "C" int vmb_add_norm_bwd(const void* x, int x_dtype, int64_t ldx, const void* residual,
                                int residual_dtype, const void* weight, int w_dtype, const void* dy,
                                const void* dresidual_out, int dresidual_out_dtype, void* dx,
                                void* dresidual, float* dweight, float* dbias, int64_t rows, int dim,
                                float eps, int is_rms, void* workspace, int64_t workspace_bytes,
                                vmb_stream_t stream) {
  VMB_CHECK_ARG(dtype_ok(x_dtype) && dtype_ok(w_dtype), "add_norm_bwd: bad dtype");
  VMB_CHECK_ARG(rows >= 0 && dim > 0 && ldx >= dim, "add_norm_bwd: bad sizes");
  if (!residual) residual_dtype = VMB_F32;
  if (!dresidual_out) dresidual_out_dtype = VMB_F32;
  VMB_CHECK_ARG(dtype_ok(residual_dtype) && dtype_ok(dresidual_out_dtype), "add_norm_bwd: bad dtype");
  VMB_CHECK_ARG(dresidual == nullptr || residual != nullptr, "add_norm_bwd: dresidual without a residual");
  cudaStream_t st = as_stream(stream);
  if (rows == 0) {
    if (dweight) VMB_CUDA(cudaMemsetAsync(dweight, 0, (size_t)dim * sizeof(float), st));
    if (dbias) VMB_CUDA(cudaMemsetAsync(dbias, 0, (size_t)dim * sizeof(float), st));
    return VMB_OK;
  }
  VMB_CHECK_ARG(x && weight && dy && dx, "add_norm_bwd: null x / weight / dy / dx");
  const int ctas = add_norm_bwd_ctas(rows);
  VMB_CHECK_ARG(workspace && workspace_bytes >= 2ll * ctas * dim * (int64_t)sizeof(float),
                "add_norm_bwd: workspace too small");
  float* pw = reinterpret_cast<float*>(workspace);
  float* pb = pw + (int64_t)ctas * dim;
  {
    auto al = [](const void* p, int bytes) { return p == nullptr || reinterpret_cast<uintptr_t>(p) % bytes == 0; };
    const int vb = x_dtype == VMB_F32 ? 16 : 8;
    const bool vec = w_dtype == x_dtype && residual_dtype == VMB_F32 && dresidual_out_dtype == VMB_F32 &&
                     dim % 4 == 0 && ldx % 4 == 0 && al(x, vb) && al(dy, vb) && al(dx, vb) && al(weight, vb) &&
                     al(residual, 16) && al(dresidual_out, 16) && al(dresidual, 16);
    if (vec) {
      const bool ok = x_dtype == VMB_F32
          ? launch_add_norm_bwd_vec<float>(x, ldx, residual, weight, dy, dresidual_out, dx, dresidual, pw, pb,
                                           rows, dim, eps, is_rms != 0, ctas, st)
          : launch_add_norm_bwd_vec<__nv_bfloat16>(x, ldx, residual, weight, dy, dresidual_out, dx, dresidual,
                                                   pw, pb, rows, dim, eps, is_rms != 0, ctas, st);
      if (ok) {
        VMB_LAUNCH_CHECK("add_norm_bwd_vec_kernel");
        if (dweight) { int rc = reduce_partials(pw, ctas, dim, dweight, VMB_F32, st); if (rc) return rc; }
        if (dbias) { int rc = reduce_partials(pb, ctas, dim, dbias, VMB_F32, st); if (rc) return rc; }
        return VMB_OK;
      }
    }
  }
#define VMB_ANB(IT)                                                                                      \
  do {                                                                                                   \
    if (is_rms)                                                                                          \
      add_norm_bwd_kernel<IT, true><<<ctas, kBwdWarps * 32, 0, st>>>(                                    \
          x, x_dtype, ldx, residual, residual_dtype, weight, w_dtype, dy, dresidual_out,                 \
          dresidual_out_dtype, dx, dresidual, pw, pb, rows, dim, eps);                                   \
    else                                                                                                 \
      add_norm_bwd_kernel<IT, false><<<ctas, kBwdWarps * 32, 0, st>>>(                                   \
          x, x_dtype, ldx, residual, residual_dtype, weight, w_dtype, dy, dresidual_out,                 \
          dresidual_out_dtype, dx, dresidual, pw, pb, rows, dim, eps);                                   \
  } while (0)
  if (dim <= 32 * 8) VMB_ANB(8);
  else if (dim <= 32 * 12) VMB_ANB(12);
  else if (dim <= 32 * 18) VMB_ANB(18);
  else if (dim <= 32 * 36) VMB_ANB(36);
  else VMB_UNSUPPORTED("add_norm_bwd: dim %d > 1152 not supported", dim);
#undef VMB_ANB
  VMB_LAUNCH_CHECK("add_norm_bwd_kernel");
  if (dweight) { const int rc = reduce_partials(pw, ctas, dim, dweight, VMB_F32, st); if (rc) return rc; }
  if (dbias) { const int rc = reduce_partials(pb, ctas, dim, dbias, VMB_F32, st); if (rc) return rc; }
  return VMB_OK;
}

extern "C" int64_t vmb_causal_conv1d_bwd_workspace_bytes(int B, int L, int Di, int W) {
  if (B <= 0 || L <= 0 || Di <= 0 || W <= 0) return 0;
  const int64_t chunks = (L + kConvChunk - 1) / kConvChunk;
  return (int64_t)B * chunks * Di * (W + 1) * (int64_t)sizeof(float);
}

extern "C" int vmb_causal_conv1d_bwd(const void* x, int64_t x_bstride, int64_t x_tstride, const void* weight,
                                     const void* bias, const void* conv_state_in, int cs_in_dtype,
                                     const void* dy, const void* dconv_state_out, int dcs_out_dtype,
                                     void* dx, int64_t dx_bstride, int64_t dx_tstride, void* dconv_state_in,
                                     float* dweight, float* dbias, int B, int L, int Di, int W, int silu,
                                     int dtype, void* workspace, int64_t workspace_bytes, vmb_stream_t stream) {
  if (dx_bstride == 0 && dx_tstride == 0) { dx_tstride = Di; dx_bstride = (int64_t)L * Di; }
  VMB_CHECK_ARG(dtype_ok(dtype), "conv1d_bwd: bad dtype %d", dtype);
  VMB_CHECK_ARG(B >= 0 && L >= 0 && Di > 0 && W > 0, "conv1d_bwd: bad sizes");
  if (W > kConvWMax) VMB_UNSUPPORTED("conv1d_bwd: d_conv %d > %d not supported", W, kConvWMax);
  if (!conv_state_in) cs_in_dtype = VMB_F32;
  if (!dconv_state_out) dcs_out_dtype = VMB_F32;
  VMB_CHECK_ARG(dtype_ok(cs_in_dtype) && dtype_ok(dcs_out_dtype), "conv1d_bwd: bad state dtype");
  VMB_CHECK_ARG(dconv_state_in == nullptr || conv_state_in != nullptr, "conv1d_bwd: dconv_state_in without a state");
  VMB_CHECK_ARG(B <= 65535, "conv1d_bwd: batch %d > 65535", B);
  cudaStream_t st = as_stream(stream);
  if (B == 0 || L == 0) {
    if (dweight) VMB_CUDA(cudaMemsetAsync(dweight, 0, (size_t)Di * W * sizeof(float), st));
    if (dbias) VMB_CUDA(cudaMemsetAsync(dbias, 0, (size_t)Di * sizeof(float), st));
    if (L == 0 && B > 0 && dconv_state_in != nullptr) {
      // empty chunk: the new state IS the old one
      if (dconv_state_out && dcs_out_dtype == cs_in_dtype)
        VMB_CUDA(cudaMemcpyAsync(dconv_state_in, dconv_state_out, (size_t)B * Di * W * dtype_size(cs_in_dtype),
                                 cudaMemcpyDeviceToDevice, st));
      else if (!dconv_state_out)
        VMB_CUDA(cudaMemsetAsync(dconv_state_in, 0, (size_t)B * Di * W * dtype_size(cs_in_dtype), st));
      else VMB_UNSUPPORTED("conv1d_bwd: empty sequence with mixed state dtypes");
    }
    return VMB_OK;
  }
  VMB_CHECK_ARG(x && weight && dy && dx, "conv1d_bwd: null x / weight / dy / dx");
  const int chunks = (L + kConvChunk - 1) / kConvChunk;
  VMB_CHECK_ARG(chunks <= 65535, "conv1d_bwd: sequence too long");
  const int64_t need = (int64_t)B * chunks * Di * (W + 1) * (int64_t)sizeof(float);
  VMB_CHECK_ARG(workspace && workspace_bytes >= need, "conv1d_bwd: workspace too small");
  float* partial = reinterpret_cast<float*>(workspace);
  // fast path: even channel count, pair-aligned rows, d_conv 2..4
  const int esz = dtype_size(dtype);
  auto al = [&](const void* p) { return reinterpret_cast<uintptr_t>(p) % (2 * esz) == 0; };
  VMB_CHECK_ARG(dx_tstride >= Di, "conv1d_bwd: dx token stride %lld < Di", (long long)dx_tstride);
  const bool pair_ok = Di % 2 == 0 && x_bstride % 2 == 0 && x_tstride % 2 == 0 && dx_bstride % 2 == 0 &&
                       dx_tstride % 2 == 0 && al(x) && al(dy) && al(dx) &&
                       (bias == nullptr || al(bias)) && W >= 2 && W <= 4;
  if (pair_ok) {
#define VMB_CBP(T, WW)                                                                                          \
  launch_conv_bwd_pair<T, WW>(x, x_bstride, x_tstride, weight, bias, conv_state_in, cs_in_dtype, dy,            \
                              dconv_state_out, dcs_out_dtype, dx, dx_bstride, dx_tstride, dconv_state_in, partial, B, L, \
                              Di, silu, chunks, st)
    if (dtype == VMB_F32) { if (W == 4) VMB_CBP(float, 4); else if (W == 3) VMB_CBP(float, 3); else VMB_CBP(float, 2); }
    else { if (W == 4) VMB_CBP(__nv_bfloat16, 4); else if (W == 3) VMB_CBP(__nv_bfloat16, 3); else VMB_CBP(__nv_bfloat16, 2); }
#undef VMB_CBP
    VMB_LAUNCH_CHECK("conv1d_bwd_pair_kernel");
  } else {
    conv1d_bwd_kernel<<<dim3((Di + 127) / 128, B, chunks), 128, 0, st>>>(
        x, x_bstride, x_tstride, weight, bias, conv_state_in, cs_in_dtype, dy, dconv_state_out, dcs_out_dtype,
        dx, dx_bstride, dx_tstride, dconv_state_in, partial, B, L, Di, W, silu, dtype);
    VMB_LAUNCH_CHECK("conv1d_bwd_kernel");
  }
  // partial rows are [dw_0 .. dw_{W-1}, db] per channel: reduce into one (Di, W + 1) fp32 table, then split
  const int64_t n = (int64_t)Di * (W + 1);
  float* table = partial;   // in place: row 0 of the partials becomes the sum
  if (B * chunks > 1) {
    const int rc = reduce_partials(partial, B * chunks, n, table, VMB_F32, st);
    if (rc != VMB_OK) return rc;
  }
  if (dweight) VMB_CUDA(cudaMemcpy2DAsync(dweight, (size_t)W * sizeof(float), table, (size_t)(W + 1) * sizeof(float),
                                          (size_t)W * sizeof(float), Di, cudaMemcpyDeviceToDevice, st));
  if (dbias) VMB_CUDA(cudaMemcpy2DAsync(dbias, sizeof(float), table + W, (size_t)(W + 1) * sizeof(float),
                                        sizeof(float), Di, cudaMemcpyDeviceToDevice, st));
  return VMB_OK;
}
