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

// ---- causal conv + SiLU backward -------------------------------------------------------------------
// Thread = channel; a CTA walks one chunk of kConvChunk tokens of one sequence in order, carrying the last
// W inputs and the last W pre-activation gradients in registers.  hist[j], j in [-(W-1), L): x[j] for
// j >= 0, conv_state_in[W + j] for j < 0 (zeros without a state).
constexpr int kConvChunk = 64;
constexpr int kConvWMax = 8;

__global__ void __launch_bounds__(128)
conv1d_bwd_kernel(const void* __restrict__ x, int64_t x_bs, int64_t x_ts, const void* __restrict__ weight,
                  const void* __restrict__ bias, const void* __restrict__ cs_in, int cs_in_dtype,
                  const void* __restrict__ dy, const void* __restrict__ dcs_out, int dcs_out_dtype,
                  void* __restrict__ dx, void* __restrict__ dcs_in, float* __restrict__ partial,
                  int B, int L, int Di, int W, int silu, int dtype) {
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
          store_from_f32(dx, ((int64_t)b * L + j) * Di + d, dtype, s);
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

}  // namespace

int reduce_partials(const float* partial, int P, int64_t n, void* out, int out_dtype, cudaStream_t st) {
  if (n <= 0) return VMB_OK;
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
  reduce_partials_kernel<<<(N + 255) / 256, 256, 0, st>>>(partial, (int)P, N, out, out_dtype);
  VMB_LAUNCH_CHECK("reduce_partials_kernel");
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
  if (dweight) {
    reduce_partials_kernel<<<(dim + 255) / 256, 256, 0, st>>>(pw, ctas, dim, dweight, VMB_F32);
    VMB_LAUNCH_CHECK("reduce_partials_kernel");
  }
  if (dbias) {
    reduce_partials_kernel<<<(dim + 255) / 256, 256, 0, st>>>(pb, ctas, dim, dbias, VMB_F32);
    VMB_LAUNCH_CHECK("reduce_partials_kernel");
  }
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
                                     void* dx, void* dconv_state_in, float* dweight, float* dbias, int B,
                                     int L, int Di, int W, int silu, int dtype, void* workspace,
                                     int64_t workspace_bytes, vmb_stream_t stream) {
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
  conv1d_bwd_kernel<<<dim3((Di + 127) / 128, B, chunks), 128, 0, st>>>(
      x, x_bstride, x_tstride, weight, bias, conv_state_in, cs_in_dtype, dy, dconv_state_out, dcs_out_dtype,
      dx, dconv_state_in, partial, B, L, Di, W, silu, dtype);
  VMB_LAUNCH_CHECK("conv1d_bwd_kernel");
  // partial rows are [dw_0 .. dw_{W-1}, db] per channel: reduce into one (Di, W + 1) fp32 table, then split
  const int64_t n = (int64_t)Di * (W + 1);
  float* table = partial;   // in place: row 0 of the partials becomes the sum
  if (B * chunks > 1) {
    reduce_partials_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(partial, B * chunks, n, table, VMB_F32);
    VMB_LAUNCH_CHECK("reduce_partials_kernel");
  }
  if (dweight) VMB_CUDA(cudaMemcpy2DAsync(dweight, (size_t)W * sizeof(float), table, (size_t)(W + 1) * sizeof(float),
                                          (size_t)W * sizeof(float), Di, cudaMemcpyDeviceToDevice, st));
  if (dbias) VMB_CUDA(cudaMemcpy2DAsync(dbias, sizeof(float), table + W, (size_t)(W + 1) * sizeof(float),
                                        sizeof(float), Di, cudaMemcpyDeviceToDevice, st));
  return VMB_OK;
}
