// Back glue of the token path: pooling over the visible patch tokens + pool_norm, and the row gather
// of the masked (visible-token) path.
//
// vmb_pool_norm_fwd stands in for the pooling block of PretrainVideoMamba.forward
// (models/videomamba/videomamba.py:983-1063): mean over the patch tokens (all of them, or per frame
// with keep_temporal), combined with the CLS row per pool_type, then nn.LayerNorm (pool_norm).
// Rounding points follow the reference's torch ops in the model dtype: the mean is accumulated in
// fp32 and rounded, `cls + avg` is rounded, LayerNorm runs in fp32 and rounds once.
// Two launches: partial sums over row chunks (every CTA streams a contiguous slab, HBM-bound: the
// token tensor is read once), then one CTA per output row finishes mean / combine / LayerNorm.
// No atomics: results are run-to-run deterministic.
//
// vmb_gather_rows stands in for `tokens.gather(1, visible...)` (videomamba.py:826-836).
#include "internal.h"

namespace vmb {
namespace {

constexpr int kPoolRows = 64;      // rows per partial-sum CTA
constexpr int kPoolThreads = 128;

// ws[(b * G + g) * chunks + chunk][C] = sum of rows [chunk * kPoolRows, ...) of group g.
// Thread (cx, ry) sums rows ry, ry + kPoolRy, ... of VEC consecutive channels (one 16-byte load per row);
// the kPoolRy row slices are added through shared memory in a fixed order (deterministic).
constexpr int kPoolRy = 4;
template <typename T, int VEC>
__global__ void __launch_bounds__(kPoolThreads)
pool_partial_kernel(const T* __restrict__ x, int64_t x_bs, int64_t x_ts, int first_row, int per, int C,
                    int chunks, float* __restrict__ ws) {
  __shared__ float part[kPoolRy][kPoolThreads / kPoolRy][VEC];
  const int chunk = blockIdx.x, g = blockIdx.y, b = blockIdx.z;
  const int r0 = chunk * kPoolRows, r1 = min(per, r0 + kPoolRows);
  const T* base = x + (int64_t)b * x_bs + (int64_t)(first_row + g * per) * x_ts;
  float* out = ws + ((int64_t)(b * gridDim.y + g) * chunks + chunk) * C;
  constexpr int kCx = kPoolThreads / kPoolRy;            // channel groups per pass
  const int cx = threadIdx.x % kCx, ry = threadIdx.x / kCx;
  for (int c0 = 0; c0 < C; c0 += kCx * VEC) {
    const int c = c0 + cx * VEC;
    float acc[VEC];
#pragma unroll
    for (int v = 0; v < VEC; ++v) acc[v] = 0.f;
    if (c < C) {
      for (int r = r0 + ry; r < r1; r += kPoolRy) {
        const T* p = base + (int64_t)r * x_ts + c;
        if constexpr (VEC == 8) {                          // bf16: one 16-byte load
          const uint4 q = *reinterpret_cast<const uint4*>(p);
          const uint32_t w[4] = {q.x, q.y, q.z, q.w};
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            acc[2 * k] += __uint_as_float(w[k] << 16);
            acc[2 * k + 1] += __uint_as_float(w[k] & 0xffff0000u);
          }
        } else if constexpr (VEC == 4) {                   // fp32: one 16-byte load
          const float4 q = *reinterpret_cast<const float4*>(p);
          acc[0] += q.x; acc[1] += q.y; acc[2] += q.z; acc[3] += q.w;
        } else {
          acc[0] += to_f32<T>(*p);
        }
      }
    }
#pragma unroll
    for (int v = 0; v < VEC; ++v) part[ry][cx][v] = acc[v];
    __syncthreads();
    if (ry == 0 && c < C) {
#pragma unroll
      for (int v = 0; v < VEC; ++v) {
        float t = part[0][cx][v];
#pragma unroll
        for (int k = 1; k < kPoolRy; ++k) t += part[k][cx][v];
        out[c + v] = t;
      }
    }
    __syncthreads();
  }
}

// One CTA per output row (b, o).  mode: 0 cls, 1 cls + avg, 2 cat[cls, avg], 3 avg.
template <typename T>
__global__ void __launch_bounds__(kPoolThreads)
pool_finish_kernel(const T* __restrict__ x, int64_t x_bs, const float* __restrict__ ws, int G, int per,
                   int C, int chunks, int mode, const T* __restrict__ ln_w, const T* __restrict__ ln_b,
                   float eps, T* __restrict__ out, int out_rows) {
  extern __shared__ float row[];                 // [C] the row LayerNorm sees (already rounded to T)
  __shared__ float red[2][kPoolThreads / 32];
  const int o = blockIdx.x, b = blockIdx.y;
  const bool is_cls_row = mode == 0 || (mode == 2 && o == 0);
  const int g = mode == 2 ? o - 1 : o;           // group whose mean this row uses
  const T* cls = x + (int64_t)b * x_bs;          // token 0 (only read when the mode has a CLS row)
  float s1 = 0.f;
  for (int c = threadIdx.x; c < C; c += kPoolThreads) {
    float v;
    if (is_cls_row) {
      v = to_f32<T>(cls[c]);
    } else {
      const float* p = ws + (int64_t)(b * G + g) * chunks * C + c;
      float acc = 0.f;
      for (int k = 0; k < chunks; ++k) acc += p[(int64_t)k * C];
      v = to_f32<T>(from_f32<T>(acc / (float)per));                      // patches.mean(...) in T
      if (mode == 1) v = to_f32<T>(from_f32<T>(to_f32<T>(cls[c]) + v));  // cls_token + avg in T
    }
    row[c] = v;
    s1 += v;
  }
  // LayerNorm over C in fp32 (two passes over the shared row: mean, then variance)
  s1 = warp_sum(s1);
  if ((threadIdx.x & 31) == 0) red[0][threadIdx.x >> 5] = s1;
  __syncthreads();
  float mean = 0.f;
  for (int i = 0; i < kPoolThreads / 32; ++i) mean += red[0][i];
  mean /= (float)C;
  float s2 = 0.f;
  for (int c = threadIdx.x; c < C; c += kPoolThreads) {
    const float d = row[c] - mean;
    s2 += d * d;
  }
  s2 = warp_sum(s2);
  if ((threadIdx.x & 31) == 0) red[1][threadIdx.x >> 5] = s2;
  __syncthreads();
  float var = 0.f;
  for (int i = 0; i < kPoolThreads / 32; ++i) var += red[1][i];
  const float rstd = rsqrtf(var / (float)C + eps);
  T* dst = out + ((int64_t)b * out_rows + o) * C;
  for (int c = threadIdx.x; c < C; c += kPoolThreads) {
    float v = (row[c] - mean) * rstd;
    if (ln_w) v *= to_f32<T>(ln_w[c]);
    if (ln_b) v += to_f32<T>(ln_b[c]);
    dst[c] = from_f32<T>(v);
  }
}

template <typename V>
__global__ void gather_rows_kernel(const V* __restrict__ src, int64_t src_bs_vecs, int64_t src_ts_vecs,
                                   const int64_t* __restrict__ index, int n_per_batch, int row_vecs,
                                   V* __restrict__ dst) {
  const int64_t r = blockIdx.x;                  // output row (b, i)
  const int b = (int)(r / n_per_batch);
  const V* s = src + (int64_t)b * src_bs_vecs + index[r] * src_ts_vecs;
  V* d = dst + r * row_vecs;
  for (int i = threadIdx.x; i < row_vecs; i += blockDim.x) d[i] = s[i];
}

template <typename T>
int pool_launch(const void* x, int64_t x_bs, int64_t x_ts, int B, int G, int per, int C, int has_cls, int mode,
                const void* ln_w, const void* ln_b, float eps, void* out, float* ws, cudaStream_t st) {
  const int chunks = (per + kPoolRows - 1) / kPoolRows;
  const int out_rows = mode == 0 ? 1 : (mode == 2 ? G + 1 : G);
  if (mode != 0) {
    dim3 grid(chunks, G, B);
    constexpr int kVec = 16 / (int)sizeof(T);
    const bool vec = C % kVec == 0 && x_bs % kVec == 0 && x_ts % kVec == 0 &&
                     reinterpret_cast<uintptr_t>(x) % 16 == 0;
    if (vec)
      pool_partial_kernel<T, kVec><<<grid, kPoolThreads, 0, st>>>((const T*)x, x_bs, x_ts, has_cls ? 1 : 0, per,
                                                                  C, chunks, ws);
    else
      pool_partial_kernel<T, 1><<<grid, kPoolThreads, 0, st>>>((const T*)x, x_bs, x_ts, has_cls ? 1 : 0, per, C,
                                                               chunks, ws);
    VMB_LAUNCH_CHECK("pool_partial_kernel");
  }
  dim3 grid(out_rows, B);
  pool_finish_kernel<T><<<grid, kPoolThreads, (size_t)C * sizeof(float), st>>>(
      (const T*)x, x_bs, ws, G, per, C, chunks, mode, (const T*)ln_w, (const T*)ln_b, eps, (T*)out, out_rows);
  VMB_LAUNCH_CHECK("pool_finish_kernel");
  return VMB_OK;
}

}  // namespace
}  // namespace vmb

extern "C" int64_t vmb_pool_norm_workspace_bytes(int B, int G, int per, int C) {
  if (B <= 0 || G <= 0 || per <= 0 || C <= 0) return 0;
  const int64_t chunks = (per + vmb::kPoolRows - 1) / vmb::kPoolRows;
  return (int64_t)B * G * chunks * C * 4;
}

extern "C" int vmb_pool_norm_fwd(const void* x, int64_t x_bstride, int64_t x_tstride, int B, int G, int per,
                                 int C, int has_cls, int mode, const void* ln_weight, const void* ln_bias,
                                 float eps, void* out, void* workspace, int64_t workspace_bytes, int dtype,
                                 vmb_stream_t stream) {
  using namespace vmb;
  VMB_CHECK_ARG(dtype_ok(dtype), "pool_norm: bad dtype %d", dtype);
  VMB_CHECK_ARG(B >= 0 && G >= 1 && per >= 0 && C > 0, "pool_norm: bad sizes");
  VMB_CHECK_ARG(mode >= 0 && mode <= 3, "pool_norm: mode %d (0 cls, 1 cls+avg, 2 cls_cat_avg, 3 avg)", mode);
  VMB_CHECK_ARG(mode == 3 || has_cls, "pool_norm: this mode needs the CLS row (token 0)");
  VMB_CHECK_ARG(mode == 0 || per >= 1, "pool_norm: no patch tokens to average");
  VMB_CHECK_ARG(B <= 65535 && G <= 65535, "pool_norm: batch / groups > 65535");
  if (B == 0) return VMB_OK;
  VMB_CHECK_ARG(x && out, "pool_norm: null x / out");
  if (C * sizeof(float) > 48 * 1024) VMB_UNSUPPORTED("pool_norm: C=%d too wide", C);
  const int64_t need = mode == 0 ? 0 : vmb_pool_norm_workspace_bytes(B, G, per, C);
  VMB_CHECK_ARG(need == 0 || (workspace && workspace_bytes >= need), "pool_norm: workspace too small");
  cudaStream_t st = as_stream(stream);
  ProfScope ps(VMB_PROF_OTHER, st);
  if (dtype == VMB_F32)
    return pool_launch<float>(x, x_bstride, x_tstride, B, G, per, C, has_cls, mode, ln_weight, ln_bias, eps, out,
                              (float*)workspace, st);
  return pool_launch<__nv_bfloat16>(x, x_bstride, x_tstride, B, G, per, C, has_cls, mode, ln_weight, ln_bias,
                                    eps, out, (float*)workspace, st);
}

extern "C" int vmb_gather_rows(const void* src, int64_t src_bstride, int64_t src_tstride, const int64_t* index,
                               int B, int n_per_batch, int C, void* dst, int dtype, vmb_stream_t stream) {
  using namespace vmb;
  VMB_CHECK_ARG(dtype_ok(dtype), "gather_rows: bad dtype %d", dtype);
  VMB_CHECK_ARG(B >= 0 && n_per_batch >= 0 && C > 0, "gather_rows: bad sizes");
  const int64_t rows = (int64_t)B * n_per_batch;
  if (rows == 0) return VMB_OK;
  VMB_CHECK_ARG(src && index && dst, "gather_rows: null pointer");
  VMB_CHECK_ARG(rows < (1ll << 31), "gather_rows: too many rows");
  const int64_t es = dtype_size(dtype);
  cudaStream_t st = as_stream(stream);
  const bool v16 = (C * es) % 16 == 0 && (src_bstride * es) % 16 == 0 && (src_tstride * es) % 16 == 0 &&
                   reinterpret_cast<uintptr_t>(src) % 16 == 0 && reinterpret_cast<uintptr_t>(dst) % 16 == 0;
  if (v16) {
    const int rv = (int)(C * es / 16);
    gather_rows_kernel<uint4><<<(unsigned)rows, rv >= 128 ? 128 : 64, 0, st>>>(
        (const uint4*)src, src_bstride * es / 16, src_tstride * es / 16, index, n_per_batch, rv, (uint4*)dst);
  } else if (dtype == VMB_BF16) {
    gather_rows_kernel<uint16_t><<<(unsigned)rows, 128, 0, st>>>((const uint16_t*)src, src_bstride, src_tstride,
                                                                  index, n_per_batch, C, (uint16_t*)dst);
  } else {
    gather_rows_kernel<uint32_t><<<(unsigned)rows, 128, 0, st>>>((const uint32_t*)src, src_bstride, src_tstride,
                                                                  index, n_per_batch, C, (uint32_t*)dst);
  }
  VMB_LAUNCH_CHECK("gather_rows_kernel");
  return VMB_OK;
}
