// Weight gradient of a projection on the 5th-generation tensor cores:  dW (N, K) = dY^T X, dY (M, N) and
// X (M, K) bf16 token-major, summed over the M tokens (backward of the nn.Linear calls of the mixer, reference
// mamba_simple.py:333-339, :445-446, which the reference differentiates through torch autograd).
//
// The contraction runs over the TOKEN axis -- the slow axis of both operands -- so for tcgen05.mma both
// operands are MN-major: a TMA box of 64 tokens x 64 features with the 128-byte swizzle IS the canonical
// MN-major SWIZZLE_128B atom (rows = k index, 128 bytes = 64 elements of the M / N index); the 64-feature
// atoms of a tile sit one box (8 KB) apart (leading byte offset), the 8-token groups 1 KB apart (stride byte
// offset), and a K = 16 step advances the descriptor by 2 KB.  No transposed copies, no ldmatrix.
//   * CTA = one 128 (n) x BNo (k) tile of dW x one slice of the tokens (grid z); 192 threads:
//     warp 0 TMA producer (2 + BNo / 64 boxes per 64-token step into a 4-6 deep ring), warp 1 MMA issuer
//     (M 128 x N BNo x K 16, fp32 accumulator in TMEM) and TMEM owner, warps 2-5 epilogue (tcgen05.ld,
//     fp32 partial tile to global memory);
//   * the per-slice partial tiles are summed by reduce_partials (two-stage, deterministic).
// A projection with fewer than 128 outputs (x_proj: N = 64) runs with the operands swapped -- the tile rows are
// the K features, its columns the N outputs -- and the epilogue writes the transpose.
// The mma.sync kernel of wgrad.cu keeps the shapes this one does not take (N and K both below 128, tiny M).
#include <algorithm>

#include "internal.h"

namespace vmb {
namespace {

constexpr int kTok = 64;                  // tokens per pipeline step (one box)
constexpr int kBoxBytes = kTok * 128;     // 64 tokens x 64 bf16
constexpr int kThreads = 192;
constexpr int BMo = 128;                  // rows of dW per tile (UMMA M)

__host__ __device__ constexpr int stages_for(int bno) { return bno >= 256 ? 4 : (bno >= 192 ? 5 : 6); }
__host__ __device__ constexpr int tmem_cols_for(int bno) { return bno <= 64 ? 64 : (bno <= 128 ? 128 : 256); }
__host__ __device__ constexpr size_t smem_bytes_for(int bno) {
  return 1024 + (size_t)stages_for(bno) * (2 + bno / 64) * kBoxBytes + 256;
}

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return static_cast<uint32_t>(__cvta_generic_to_shared(p)); }
__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(pred));
  return pred != 0;
}
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  uint32_t done;
  do {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(done)
        : "r"(bar), "r"(parity)
        : "memory");
  } while (!done);
}
__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      ::"r"(dst), "l"(map), "r"(bar), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void tc_mma_bf16(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc,
                                            uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void tc_ld_32x32(uint32_t taddr, uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]),
        "=r"(r[7]), "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]),
        "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]),
        "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]),
        "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr));
}
__device__ __forceinline__ void tc_wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// UMMA shared-memory descriptor of an MN-major, 128B-swizzled operand: 64-element (128-byte) rows of the
// M / N index, one row per k (token); 8-row groups 1 KB apart (stride byte offset), 64-element atoms one box
// apart (leading byte offset).  Bits: [0,14) start >> 4, [16,30) LBO >> 4, [32,46) SBO >> 4, [46,48) version 1,
// [61,64) layout 2 (SWIZZLE_128B).
__device__ __forceinline__ uint64_t umma_desc_mn_sw128(uint32_t smem_addr) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr >> 4) & 0x3FFF);
  d |= (uint64_t)(kBoxBytes >> 4) << 16;
  d |= (uint64_t)(1024 >> 4) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)2 << 61;
  return d;
}
// kind::f16 instruction descriptor: fp32 accumulate, A and B bf16, BOTH MN-major (bits 15 and 16 set)
__host__ __device__ constexpr uint32_t umma_idesc_bf16_mn(int m, int n) {
  return (1u << 4) | (1u << 7) | (1u << 10) | (1u << 15) | (1u << 16) | ((uint32_t)(n >> 3) << 17) |
         ((uint32_t)(m >> 4) << 24);
}

// Tile rows = features of operand A (Fa of them), columns = features of operand B (Fb); element (r, c) of the
// slice's partial goes to partial[slice * Fa * Fb + r * ld_row + c * ld_col]: (A, B) = (dY, X) with ld_row = K,
// ld_col = 1, or swapped (A, B) = (X, dY) with ld_row = 1, ld_col = K.
template <int BNo>
__global__ void __launch_bounds__(kThreads, 1)
wgrad_tc_kernel(const __grid_constant__ CUtensorMap map_dy, const __grid_constant__ CUtensorMap map_x,
                float* __restrict__ partial, int64_t M, int N, int K, int64_t rows_per_split, int64_t ld_row,
                int64_t ld_col) {
  constexpr int kStages = stages_for(BNo);
  constexpr int kXBoxes = BNo / 64;
  constexpr uint32_t kStageBytes = (2 + kXBoxes) * kBoxBytes;
  constexpr uint32_t kIdesc = umma_idesc_bf16_mn(BMo, BNo);
  constexpr int kTmemCols = tmem_cols_for(BNo);

  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t bars = smem_base + kStages * kStageBytes;
  auto full_bar = [&](int s) { return bars + 8u * s; };
  auto empty_bar = [&](int s) { return bars + 8u * (kStages + s); };
  const uint32_t tfull_bar = bars + 8u * (2 * kStages);
  const uint32_t tmem_slot = bars + 8u * (2 * kStages + 1);
  volatile uint32_t* tmem_slot_ptr =
      reinterpret_cast<volatile uint32_t*>(smem_raw + (tmem_slot - smem_u32(smem_raw)));

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int n0 = blockIdx.x * BMo, k0 = blockIdx.y * BNo;
  const int64_t m_begin = (int64_t)blockIdx.z * rows_per_split;
  const int64_t m_end = min(M, m_begin + rows_per_split);
  const int nsteps = m_end > m_begin ? (int)((m_end - m_begin + kTok - 1) / kTok) : 0;

  if (warp == 1 && elect_one()) {
    for (int s = 0; s < kStages; ++s) { mbar_init(full_bar(s), 1); mbar_init(empty_bar(s), 1); }
    mbar_init(tfull_bar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_slot),
                 "r"((uint32_t)kTmemCols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot_ptr;

  if (warp == 0) {
    // ===== TMA producer =====
    if (elect_one()) {
      int stage = 0;
      uint32_t phase = 0;
      for (int step = 0; step < nsteps; ++step) {
        const int m = (int)(m_begin + (int64_t)step * kTok);
        mbar_wait(empty_bar(stage), phase ^ 1);
        mbar_expect_tx(full_bar(stage), kStageBytes);
        const uint32_t dst = smem_base + stage * kStageBytes;
        tma_load_2d(dst, &map_dy, full_bar(stage), n0, m);
        tma_load_2d(dst + kBoxBytes, &map_dy, full_bar(stage), n0 + 64, m);
#pragma unroll
        for (int j = 0; j < kXBoxes; ++j)
          tma_load_2d(dst + (2 + j) * kBoxBytes, &map_x, full_bar(stage), k0 + 64 * j, m);
        if (++stage == kStages) { stage = 0; phase ^= 1; }
      }
    }
  } else if (warp == 1) {
    // ===== MMA issuer =====
    if (elect_one()) {
      int stage = 0;
      uint32_t phase = 0;
      for (int step = 0; step < nsteps; ++step) {
        mbar_wait(full_bar(stage), phase);
        tc_fence_after();
        const uint32_t base = smem_base + stage * kStageBytes;
        const uint64_t a_desc = umma_desc_mn_sw128(base);
        const uint64_t b_desc = umma_desc_mn_sw128(base + 2 * kBoxBytes);
#pragma unroll
        for (int ks = 0; ks < kTok / 16; ++ks)       // 16 tokens = two 8-row groups = 2 KB: +128 in (addr >> 4)
          tc_mma_bf16(tmem_base, a_desc + 128u * ks, b_desc + 128u * ks, kIdesc, (step | ks) != 0);
        tc_commit(empty_bar(stage));
        if (++stage == kStages) { stage = 0; phase ^= 1; }
      }
      tc_commit(tfull_bar);
    }
  } else {
    // ===== epilogue: accumulator row = n feature, column = k feature =====
    const int ew = warp & 3;
    const int n = n0 + ew * 32 + lane;
    float* out = partial + (int64_t)blockIdx.z * N * K + (int64_t)n * ld_row + (int64_t)k0 * ld_col;
    if (nsteps > 0) {
      mbar_wait(tfull_bar, 0);
      tc_fence_after();
    }
#pragma unroll 1
    for (int c = 0; c < BNo; c += 32) {
      if (k0 + c >= K) break;
      uint32_t v[32];
      if (nsteps > 0) {
        tc_ld_32x32(tmem_base + ((uint32_t)(ew * 32) << 16) + c, v);
        tc_wait_ld();
      } else {
#pragma unroll
        for (int j = 0; j < 32; ++j) v[j] = 0u;
      }
      if (n < N) {
        if (ld_col == 1) {
#pragma unroll
          for (int j = 0; j < 32; j += 4)
            if (k0 + c + j < K)                      // K % 8 == 0: whole float4s
              *reinterpret_cast<float4*>(out + c + j) =
                  make_float4(__uint_as_float(v[j]), __uint_as_float(v[j + 1]), __uint_as_float(v[j + 2]),
                              __uint_as_float(v[j + 3]));
        } else {                                     // transposed: a warp's 32 rows are contiguous per column
#pragma unroll
          for (int j = 0; j < 32; ++j)
            if (k0 + c + j < K) out[(int64_t)(c + j) * ld_col] = __uint_as_float(v[j]);
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"((uint32_t)kTmemCols)
                 : "memory");
  }
}

int bno_for(int K) { return K > 128 ? (K % 256 == 0 && K % 192 != 0 ? 256 : 192) : (K > 64 ? 128 : 64); }

template <int BNo>
int launch(const CUtensorMap& mdy, const CUtensorMap& mx, float* partial, int64_t M, int N, int K, int splits,
           int64_t ld_row, int64_t ld_col, cudaStream_t st) {
  constexpr size_t smem = smem_bytes_for(BNo);
  static bool attr_set[64] = {false};
  int dev = 0;
  VMB_CUDA(cudaGetDevice(&dev));
  if (dev < 0 || dev >= 64 || !attr_set[dev]) {
    VMB_CUDA(cudaFuncSetAttribute(wgrad_tc_kernel<BNo>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    if (dev >= 0 && dev < 64) attr_set[dev] = true;
  }
  const int64_t per = ((M + splits - 1) / splits + kTok - 1) / kTok * kTok;
  dim3 grid((N + BMo - 1) / BMo, (K + BNo - 1) / BNo, splits);
  wgrad_tc_kernel<BNo><<<grid, kThreads, smem, st>>>(mdy, mx, partial, M, N, K, per, ld_row, ld_col);
  VMB_LAUNCH_CHECK("wgrad_tc_kernel");
  return VMB_OK;
}

}  // namespace

bool wgrad_tc_supported(const void* dy, int64_t ldy, const void* x, int64_t ldx, int64_t M, int N, int K) {
  auto al16 = [](const void* p) { return reinterpret_cast<uintptr_t>(p) % 16 == 0; };
  return al16(dy) && al16(x) && ldy % 8 == 0 && ldx % 8 == 0 && N % 8 == 0 && K % 8 == 0 && N >= 8 && K >= 8 &&
         (N >= BMo || K >= BMo) && M >= 4 * kTok && M < (1ll << 31);
}

int wgrad_tc_splits(int64_t M, int N, int K) {
  if (N < BMo) std::swap(N, K);                   // swapped roles (see wgrad_tc)
  const int bno = bno_for(K);
  const int tiles = ((N + BMo - 1) / BMo) * ((K + bno - 1) / bno);
  const int64_t max_by_rows = std::max<int64_t>(1, M / (4 * kTok));     // at least 256 tokens per slice
  const int64_t want = std::max<int64_t>(1, sm_count() / tiles);        // one CTA per SM (shared memory)
  return (int)std::min<int64_t>(std::min<int64_t>(want, max_by_rows), 512);
}

int wgrad_tc(const void* dy, int64_t ldy, const void* x, int64_t ldx, float* partial, int64_t M, int N, int K,
             int splits, cudaStream_t st) {
  CUtensorMap ma, mb;
  int rc;
  int64_t ld_row = K, ld_col = 1;
  if (N < BMo) {                                  // fewer than 128 outputs: tile rows = the K features, write dW transposed
    std::swap(dy, x); std::swap(ldy, ldx);
    ld_row = 1; ld_col = K;
    std::swap(N, K);
  }
  if ((rc = make_tensor_map_2d_bf16_sw128(&ma, dy, M, N, ldy, kTok))) return rc;
  if ((rc = make_tensor_map_2d_bf16_sw128(&mb, x, M, K, ldx, kTok))) return rc;
  switch (bno_for(K)) {
    case 64: return launch<64>(ma, mb, partial, M, N, K, splits, ld_row, ld_col, st);
    case 128: return launch<128>(ma, mb, partial, M, N, K, splits, ld_row, ld_col, st);
    case 192: return launch<192>(ma, mb, partial, M, N, K, splits, ld_row, ld_col, st);
    default: return launch<256>(ma, mb, partial, M, N, K, splits, ld_row, ld_col, st);
  }
}

}  // namespace vmb
