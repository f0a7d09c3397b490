// CUDA-core dense projection  C[M,N] = A[M,K] * W[N,K]^T (+ bias), fp32 accumulate.
//
// This is the TRUE-fp32 projection path (the 1e-5 parity mode cannot use tf32/bf16 tensor
// cores) and the any-shape path for the odd geometries the reference's tests use
// (d_model 8/16, dt_rank 1, d_state 4/8).  The bf16 production shapes go to gemm_tc.cu.
// Both operands are K-contiguous ("TN"), exactly how nn.Linear stores its weight
// (reference: models/videomamba/mamba_simple.py:218-220, :235-240, :279-281).
#include "common.cuh"

namespace vmb {
namespace {

constexpr int kBK = 16;

template <typename T, int BM, int BN, int TM, int TN, bool kVec>
__global__ void __launch_bounds__((BM / TM) * (BN / TN))
linear_simt_kernel(const T* __restrict__ A, int64_t lda, const T* __restrict__ W, int64_t ldw,
                   const T* __restrict__ bias, T* __restrict__ C, int64_t ldc, int64_t M, int N,
                   int K) {
  constexpr int kThreads = (BM / TM) * (BN / TN);
  __shared__ float As[kBK][BM + 4];
  __shared__ float Ws[kBK][BN + 4];
  const int tid = threadIdx.x;
  const int tx = tid % (BN / TN);
  const int ty = tid / (BN / TN);
  const int64_t m0 = (int64_t)blockIdx.y * BM;
  const int n0 = blockIdx.x * BN;

  float acc[TM][TN];
#pragma unroll
  for (int i = 0; i < TM; ++i)
#pragma unroll
    for (int j = 0; j < TN; ++j) acc[i][j] = 0.f;

  for (int k0 = 0; k0 < K; k0 += kBK) {
    if constexpr (kVec) {
      // 4 consecutive k per load; rows beyond M/N are clamped to zero.
      constexpr int kQuads = kBK / 4;
      for (int e = tid; e < BM * kQuads; e += kThreads) {
        const int r = e / kQuads, q = e % kQuads;
        float f[4] = {0.f, 0.f, 0.f, 0.f};
        if (m0 + r < M) {
          const T* p = A + (m0 + r) * lda + k0 + q * 4;
          if constexpr (sizeof(T) == 4) {
            const float4 v = *reinterpret_cast<const float4*>(p);
            f[0] = v.x; f[1] = v.y; f[2] = v.z; f[3] = v.w;
          } else {
            const uint2 v = *reinterpret_cast<const uint2*>(p);
            const __nv_bfloat162 a = *reinterpret_cast<const __nv_bfloat162*>(&v.x);
            const __nv_bfloat162 b = *reinterpret_cast<const __nv_bfloat162*>(&v.y);
            f[0] = __low2float(a); f[1] = __high2float(a);
            f[2] = __low2float(b); f[3] = __high2float(b);
          }
        }
#pragma unroll
        for (int j = 0; j < 4; ++j) As[q * 4 + j][r] = f[j];
      }
      for (int e = tid; e < BN * kQuads; e += kThreads) {
        const int r = e / kQuads, q = e % kQuads;
        float f[4] = {0.f, 0.f, 0.f, 0.f};
        if (n0 + r < N) {
          const T* p = W + (int64_t)(n0 + r) * ldw + k0 + q * 4;
          if constexpr (sizeof(T) == 4) {
            const float4 v = *reinterpret_cast<const float4*>(p);
            f[0] = v.x; f[1] = v.y; f[2] = v.z; f[3] = v.w;
          } else {
            const uint2 v = *reinterpret_cast<const uint2*>(p);
            const __nv_bfloat162 a = *reinterpret_cast<const __nv_bfloat162*>(&v.x);
            const __nv_bfloat162 b = *reinterpret_cast<const __nv_bfloat162*>(&v.y);
            f[0] = __low2float(a); f[1] = __high2float(a);
            f[2] = __low2float(b); f[3] = __high2float(b);
          }
        }
#pragma unroll
        for (int j = 0; j < 4; ++j) Ws[q * 4 + j][r] = f[j];
      }
    } else {
      for (int e = tid; e < BM * kBK; e += kThreads) {
        const int r = e / kBK, k = e % kBK;
        float f = 0.f;
        if (m0 + r < M && k0 + k < K) f = to_f32<T>(A[(m0 + r) * lda + k0 + k]);
        As[k][r] = f;
      }
      for (int e = tid; e < BN * kBK; e += kThreads) {
        const int r = e / kBK, k = e % kBK;
        float f = 0.f;
        if (n0 + r < N && k0 + k < K) f = to_f32<T>(W[(int64_t)(n0 + r) * ldw + k0 + k]);
        Ws[k][r] = f;
      }
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < kBK; ++k) {
      float a[TM], b[TN];
#pragma unroll
      for (int i = 0; i < TM; ++i) a[i] = As[k][ty * TM + i];
#pragma unroll
      for (int j = 0; j < TN; ++j) b[j] = Ws[k][tx * TN + j];
#pragma unroll
      for (int i = 0; i < TM; ++i)
#pragma unroll
        for (int j = 0; j < TN; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
    }
    __syncthreads();
  }

#pragma unroll
  for (int i = 0; i < TM; ++i) {
    const int64_t m = m0 + ty * TM + i;
    if (m >= M) continue;
#pragma unroll
    for (int j = 0; j < TN; ++j) {
      const int n = n0 + tx * TN + j;
      if (n >= N) continue;
      float v = acc[i][j];
      if (bias != nullptr) v += to_f32<T>(bias[n]);
      C[m * ldc + n] = from_f32<T>(v);
    }
  }
}

template <typename T>
int launch(const void* A, int64_t lda, const void* W, int64_t ldw, const void* bias, void* C,
           int64_t ldc, int64_t M, int N, int K, cudaStream_t st) {
  constexpr int kAlign = 4;  // elements per vector load
  const bool vec = (K % kBK == 0) && (lda % kAlign == 0) && (ldw % kAlign == 0) &&
                   (reinterpret_cast<uintptr_t>(A) % (4 * sizeof(T)) == 0) &&
                   (reinterpret_cast<uintptr_t>(W) % (4 * sizeof(T)) == 0);
  const bool big = (N >= 96 && M >= 96);
#define VMB_LS(BM, BN, TM, TN, VEC)                                                            \
  do {                                                                                         \
    dim3 grid((unsigned)((N + BN - 1) / BN), (unsigned)((M + BM - 1) / BM));                   \
    if (grid.y > 65535u) VMB_UNSUPPORTED("linear_simt: M too large for this tile");            \
    linear_simt_kernel<T, BM, BN, TM, TN, VEC><<<grid, (BM / TM) * (BN / TN), 0, st>>>(        \
        (const T*)A, lda, (const T*)W, ldw, (const T*)bias, (T*)C, ldc, M, N, K);              \
  } while (0)
  if (big) {
    if (vec) VMB_LS(128, 128, 8, 8, true); else VMB_LS(128, 128, 8, 8, false);
  } else if (M >= 4096) {
    // tall-skinny (x_proj / dt_proj): keep CTAs tall so the grid stays within limits
    if (vec) VMB_LS(128, 32, 8, 2, true); else VMB_LS(128, 32, 8, 2, false);
  } else {
    if (vec) VMB_LS(32, 32, 2, 2, true); else VMB_LS(32, 32, 2, 2, false);
  }
#undef VMB_LS
  VMB_LAUNCH_CHECK("linear_simt_kernel");
  return VMB_OK;
}

}  // namespace

int linear_simt(const void* A, int64_t lda, const void* W, int64_t ldw, const void* bias, void* C,
                int64_t ldc, int64_t M, int N, int K, int dtype, cudaStream_t st) {
  if (dtype == VMB_F32) return launch<float>(A, lda, W, ldw, bias, C, ldc, M, N, K, st);
  return launch<__nv_bfloat16>(A, lda, W, ldw, bias, C, ldc, M, N, K, st);
}

}  // namespace vmb
