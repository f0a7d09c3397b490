// Weight gradient of a projection: dW (N, K) = dY^T X, dY (M, N), X (M, K), bf16, summed over the M
// tokens (backward of the nn.Linear calls of the mixer, reference mamba_simple.py:333-339, :409,
// :413-414, :445-446, which the reference differentiates through torch autograd).
//
// The production shapes (N or K >= 128, M >= 256) run on tcgen05 (wgrad_tc.cu, dispatched by the entry point
// below); this mma.sync kernel keeps the rest.
// The contraction runs over the TOKEN axis, which is the slow axis of both operands, and the output is
// small (at most 1536 x 384): the forward GEMM (gemm_tc.cu) would need transposed copies of two
// activation tensors and would run on a handful of CTAs.  Here both operands are read in place,
// token-major rows, and the token axis is split over the grid:
//   * CTA = one 128 (n) x 128 (k) output tile x one slice of the tokens; 8 warps, each 64 x 32 of the tile
//   * 32-token steps: the dY and X rows of the step are brought in with 16-byte cp.async into a
//     double-buffered shared-memory stage (rows padded to 272 bytes: conflict-free ldmatrix)
//   * both mma operands come out of [token][feature] tiles through ldmatrix.trans (the token axis is the
//     k index of mma.sync m16n8k16 for A = dY^T and the row index of B = X)
//   * fp32 partial tiles per token slice, summed by reduce_partials (two-stage, deterministic).
#include <algorithm>

#include "internal.h"

namespace vmb {
namespace {

constexpr int kTile = 128;                 // output tile edge (n and k)
constexpr int kStep = 32;                  // tokens per pipeline step
constexpr int kPitch = 136;                // bf16 elements per shared-memory row (272 bytes)
constexpr int kThreads = 256;

__device__ __forceinline__ void cp_async16(uint32_t dst, const void* src, bool ok) {
  const int n = ok ? 16 : 0;               // src-size 0: zero fill
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(src), "r"(n) : "memory");
}
__device__ __forceinline__ void ldmatrix_x4_trans(uint32_t addr, uint32_t (&r)[4]) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0, %1, %2, %3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3])
               : "r"(addr));
}
__device__ __forceinline__ void mma16816(float (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile(
      "mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0, %1, %2, %3}, {%4, %5, %6, %7}, {%8, %9}, "
      "{%0, %1, %2, %3};"
      : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
      : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}

__global__ void __launch_bounds__(kThreads)
wgrad_kernel(const __nv_bfloat16* __restrict__ dy, int64_t ldy, const __nv_bfloat16* __restrict__ x,
             int64_t ldx, float* __restrict__ partial, int64_t M, int N, int K, int64_t rows_per_split) {
  extern __shared__ __align__(16) uint8_t smem_raw[];
  // stage s: dY tile [kStep][kPitch], X tile [kStep][kPitch]
  constexpr int kTileBytes = kStep * kPitch * 2;
  const uint32_t sbase = static_cast<uint32_t>(__cvta_generic_to_shared(smem_raw));
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int wn = warp >> 2, wk = warp & 3;           // warp tile: rows [64 wn, +64) of n, cols [32 wk, +32) of k
  const int n0 = blockIdx.x * kTile, k0 = blockIdx.y * kTile;
  const int64_t m_begin = (int64_t)blockIdx.z * rows_per_split;
  const int64_t m_end = min(M, m_begin + rows_per_split);
  const int nsteps = m_end > m_begin ? (int)((m_end - m_begin + kStep - 1) / kStep) : 0;

  // loader: 32 rows x 16 chunks of 16 bytes per operand = 512 chunks each, 2 per thread per operand
  auto load_stage = [&](int step, int stg) {
    const int64_t m0 = m_begin + (int64_t)step * kStep;
#pragma unroll
    for (int i = 0; i < 2; ++i) {
      const int c = tid + i * kThreads;              // 0..511
      const int r = c >> 4, ch = c & 15;
      const int64_t m = m0 + r;
      const bool row_ok = m < m_end;
      const int ncol = n0 + ch * 8, kcol = k0 + ch * 8;
      const uint32_t dst = sbase + stg * 2 * kTileBytes + (r * kPitch + ch * 8) * 2;
      cp_async16(dst, dy + (row_ok ? m : 0) * ldy + (ncol < N ? ncol : 0), row_ok && ncol < N);
      cp_async16(dst + kTileBytes, x + (row_ok ? m : 0) * ldx + (kcol < K ? kcol : 0), row_ok && kcol < K);
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
  };

  float acc[4][4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j)
#pragma unroll
      for (int e = 0; e < 4; ++e) acc[i][j][e] = 0.f;

  if (nsteps > 0) load_stage(0, 0);
  for (int step = 0; step < nsteps; ++step) {
    const int stg = step & 1;
    if (step + 1 < nsteps) {
      load_stage(step + 1, stg ^ 1);
      asm volatile("cp.async.wait_group 1;" ::: "memory");
    } else {
      asm volatile("cp.async.wait_group 0;" ::: "memory");
    }
    __syncthreads();
    const uint32_t sA = sbase + stg * 2 * kTileBytes;              // dY tile [token][n]
    const uint32_t sB = sA + kTileBytes;                           // X tile  [token][k]
#pragma unroll
    for (int ks = 0; ks < kStep / 16; ++ks) {
      // A = dY^T: fragment (16 n x 16 tokens) from the [token][n] tile, transposed on load.
      // ldmatrix.trans x4: matrices {tok 0-7, n 0-7}, {tok 0-7, n 8-15}, {tok 8-15, n 0-7}, {tok 8-15, n 8-15}
      // -> a0 (n 0-7, tok 0-7), a1 (n 8-15, tok 0-7), a2 (n 0-7, tok 8-15), a3 (n 8-15, tok 8-15)
      uint32_t af[4][4];
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const int tok = ks * 16 + (lane & 7) + 8 * (lane >> 4);
        const int col = wn * 64 + i * 16 + 8 * ((lane >> 3) & 1);
        ldmatrix_x4_trans(sA + (tok * kPitch + col) * 2, af[i]);
      }
      // B = X: fragment (16 tokens x 8 k) pairs: one x4.trans covers k 0-15 of the warp tile:
      // matrices {tok 0-7, k 0-7}, {tok 8-15, k 0-7}, {tok 0-7, k 8-15}, {tok 8-15, k 8-15}
      uint32_t bf[2][4];
#pragma unroll
      for (int j = 0; j < 2; ++j) {
        const int tok = ks * 16 + (lane & 7) + 8 * ((lane >> 3) & 1);
        const int col = wk * 32 + j * 16 + 8 * (lane >> 4);
        ldmatrix_x4_trans(sB + (tok * kPitch + col) * 2, bf[j]);
      }
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j)
          mma16816(acc[i][j], af[i], bf[j >> 1][2 * (j & 1)], bf[j >> 1][2 * (j & 1) + 1]);
    }
    __syncthreads();                                               // stage free for the load after next
  }
  // partial tile: partial[split][n][k]
  float* out = partial + (int64_t)blockIdx.z * N * K;
  const int g = lane >> 2, tig = lane & 3;
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int n = n0 + wn * 64 + i * 16 + g;
      const int k = k0 + wk * 32 + j * 8 + 2 * tig;
      if (k < K) {
        // K is a multiple of 8 (checked by the host), so k + 1 < K as well
        if (n < N) *reinterpret_cast<float2*>(out + (int64_t)n * K + k) = make_float2(acc[i][j][0], acc[i][j][1]);
        if (n + 8 < N) *reinterpret_cast<float2*>(out + (int64_t)(n + 8) * K + k) = make_float2(acc[i][j][2], acc[i][j][3]);
      }
    }
}

int wgrad_splits(int64_t M, int N, int K) {
  const int tiles = ((N + kTile - 1) / kTile) * ((K + kTile - 1) / kTile);
  const int64_t max_by_rows = std::max<int64_t>(1, M / (4 * kStep));      // at least 128 tokens per slice
  const int64_t want = std::max<int64_t>(1, (3ll * sm_count() + tiles - 1) / tiles);
  return (int)std::min<int64_t>(std::min<int64_t>(want, max_by_rows), 512);
}

}  // namespace
}  // namespace vmb

using namespace vmb;

extern "C" int64_t vmb_linear_wgrad_workspace_bytes(int64_t M, int N, int K) {
  if (M <= 0 || N <= 0 || K <= 0) return 0;
  const int splits = std::max(wgrad_splits(M, N, K), (N >= 128 || K >= 128) ? wgrad_tc_splits(M, N, K) : 1);
  return (int64_t)splits * N * K * (int64_t)sizeof(float);
}

extern "C" int vmb_linear_wgrad(const void* dy, int64_t ldy, const void* x, int64_t ldx, void* dw, int dw_dtype,
                                int64_t M, int N, int K, void* workspace, int64_t workspace_bytes,
                                vmb_stream_t stream) {
  VMB_CHECK_ARG(dtype_ok(dw_dtype), "linear_wgrad: bad output dtype");
  VMB_CHECK_ARG(M >= 0 && N > 0 && K > 0 && ldy >= N && ldx >= K, "linear_wgrad: bad sizes");
  VMB_CHECK_ARG(dw != nullptr, "linear_wgrad: null output");
  cudaStream_t st = as_stream(stream);
  if (M == 0) {
    VMB_CUDA(cudaMemsetAsync(dw, 0, (size_t)N * K * dtype_size(dw_dtype), st));
    return VMB_OK;
  }
  VMB_CHECK_ARG(dy && x, "linear_wgrad: null operand");
  auto al16 = [](const void* p) { return reinterpret_cast<uintptr_t>(p) % 16 == 0; };
  if (!al16(dy) || !al16(x) || ldy % 8 != 0 || ldx % 8 != 0 || N % 8 != 0 || K % 8 != 0)
    VMB_UNSUPPORTED("linear_wgrad: operands must be bf16 with 16-byte aligned rows and N, K multiples of 8");
  if (wgrad_tc_supported(dy, ldy, x, ldx, M, N, K)) {   // tcgen05: both operands MN-major, read in place
    const int sp = wgrad_tc_splits(M, N, K);
    VMB_CHECK_ARG(workspace && workspace_bytes >= (int64_t)sp * N * K * (int64_t)sizeof(float),
                  "linear_wgrad: workspace too small");
    const int rc = wgrad_tc(dy, ldy, x, ldx, reinterpret_cast<float*>(workspace), M, N, K, sp, st);
    if (rc != VMB_OK) return rc;
    return reduce_partials(reinterpret_cast<const float*>(workspace), sp, (int64_t)N * K, dw, dw_dtype, st);
  }
  const int splits = wgrad_splits(M, N, K);
  VMB_CHECK_ARG(workspace && workspace_bytes >= (int64_t)splits * N * K * (int64_t)sizeof(float),
                "linear_wgrad: workspace too small");
  const int64_t per = ((M + splits - 1) / splits + kStep - 1) / kStep * kStep;
  constexpr int smem = 2 * 2 * kStep * kPitch * 2;
  dim3 grid((N + kTile - 1) / kTile, (K + kTile - 1) / kTile, splits);
  wgrad_kernel<<<grid, kThreads, smem, st>>>(reinterpret_cast<const __nv_bfloat16*>(dy), ldy,
                                             reinterpret_cast<const __nv_bfloat16*>(x), ldx,
                                             reinterpret_cast<float*>(workspace), M, N, K, per);
  VMB_LAUNCH_CHECK("wgrad_kernel");
  return reduce_partials(reinterpret_cast<const float*>(workspace), splits, (int64_t)N * K, dw, dw_dtype, st);
}
