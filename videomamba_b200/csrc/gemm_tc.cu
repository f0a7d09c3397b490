// bf16 dense projection on the 5th-generation tensor cores:  C[M,N] = A[M,K] * W[N,K]^T (+ bias).
//
// Stands in for the cuBLAS calls behind nn.Linear / F.linear on the reference's mixer path
// (models/videomamba/mamba_simple.py:333-339 in_proj, :409 x_proj, :445-446 out_proj) and behind
// the Conv3d patch embedding (models/videomamba/videomamba.py:359-368, kernel == stride).
//
// Two kernels share the design below: gemm_tc_kernel (one CTA per tile of 128 x BN) and
// gemm_tc_pair_kernel (cta_group::2: the two CTAs of a cluster share a 256 x BN tile, each holding
// half of the W tile -- the default for N % 256 == 0 or N % 192 == 0, see its header further down).
//
// Design (B200 / sm_100a, one persistent CTA per SM, 192 threads, warp-specialised):
//   warp 0   TMA producer: A tile (128 x 64) and W tile (BN x 64) per k-block into a kStages-deep
//            ring of 128B-swizzled shared-memory buffers (cp.async.bulk.tensor + mbarrier tx).
//   warp 1   MMA issuer: one elected lane issues tcgen05.mma (M=128, N=BN, K=16, bf16 x bf16 ->
//            fp32) with both operands read from shared memory through UMMA descriptors;
//            accumulators live in TMEM (2 x BN columns: double buffered across tiles).
//            It also owns the TMEM allocation.
//   warps 2-5 epilogue: tcgen05.ld the accumulator (thread = output row), add bias, round once to
//            bf16, stage 128 x 64 sub-tiles in swizzled shared memory and write them with TMA
//            stores (coalesced, rows beyond M / columns beyond N are clipped by the hardware).
// Tiles are walked n-fastest so the CTAs running at the same time share the A rows in L2.
// Both operands are K-contiguous, exactly how nn.Linear stores its weight, so no transposes.
// Activation epilogue (act_from, vmb_linear_fwd_act): SiLU on the output columns [act_from, N) before the
// rounding -- in_proj stores x | SiLU(z), the gate of mamba_simple.py:423-435, and the fused scan only
// multiplies.  With act_from == N / 2 the CTA-pair kernel interleaves the halves: a 256-column tile is 128 x
// columns (leader's W rows) + the 128 gate columns of the same channels (peer's W rows), so every tile has the
// same epilogue work (with whole x tiles and z tiles the pairs holding more z tiles set the time).
#include <cuda.h>

#include <climits>
#include <cstdlib>
#include <type_traits>
#include <mutex>
#include <unordered_map>

#include "internal.h"

namespace vmb {
namespace {

constexpr int BM = 128;       // rows of C per tile (UMMA M)
constexpr int BK = 64;        // k-block: 64 bf16 = one 128-byte swizzle row
constexpr int UMMA_K = 16;
constexpr int kThreads = 192;
constexpr int kEpiWarp0 = 2;  // first epilogue warp (a warp reads the TMEM lane quadrant warp % 4)
constexpr int kSubN = 64;     // columns per epilogue sub-tile (128 bytes of bf16)

__host__ __device__ constexpr int stages_for(int bn) { return bn <= 128 ? 6 : 4; }
__host__ __device__ constexpr int tmem_cols_for(int bn) { return bn <= 64 ? 128 : (bn <= 128 ? 256 : 512); }
__host__ __device__ constexpr size_t smem_bytes_for(int bn, int stages) {
  return 1024 /* alignment slack */ + (size_t)stages * (BM * BK * 2 + bn * BK * 2) +
         2 * (BM * kSubN * 2) + 256 /* barriers + tmem pointer */;
}

// ---- PTX wrappers ----------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}
__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile(
      "{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(pred));
  return pred != 0;
}
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes)
               : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
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
__device__ __forceinline__ void fence_barrier_init() {
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void fence_proxy_async_smem() {
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap* map, uint32_t bar,
                                            int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      ::"r"(dst), "l"(map), "r"(bar), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ void tma_store_2d(const CUtensorMap* map, uint32_t src, int c0, int c1) {
  asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];"
               ::"l"(map), "r"(src), "r"(c0), "r"(c1)
               : "memory");
}
__device__ __forceinline__ void tma_store_commit() {
  asm volatile("cp.async.bulk.commit_group;" ::: "memory");
}
template <int kPending> __device__ __forceinline__ void tma_store_wait_read() {
  asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(kPending) : "memory");
}
__device__ __forceinline__ void tma_store_wait_all() {
  asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
}
__device__ __forceinline__ void prefetch_tmap(const CUtensorMap* map) {
  asm volatile("prefetch.tensormap [%0];" ::"l"(map) : "memory");
}
__device__ __forceinline__ void tc_fence_before() {
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
}
__device__ __forceinline__ void tc_fence_after() {
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
}
__device__ __forceinline__ void tc_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar)
               : "memory");
}
__device__ __forceinline__ void tc_mma_bf16(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc,
                                            uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// 32 lanes x 32 consecutive fp32 columns: thread l of the warp receives row (lane base + l).
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
__device__ __forceinline__ void tc_wait_ld() {
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void epi_bar_sync() {  // named barrier 1: the 128 epilogue threads
  asm volatile("bar.sync 1, 128;" ::: "memory");
}
// SiLU on a thread's 64 accumulator values before the rounding to bf16 (in_proj's z half: the scan then
// multiplies by the gate as it is).  x * (0.5 * tanh(x / 2) + 0.5): silu_fast's arithmetic, packed.
__device__ __forceinline__ void epi_silu(uint32_t (&v0)[32], uint32_t (&v1)[32]) {
  const float2 half2 = make_float2(0.5f, 0.5f);
#pragma unroll
  for (int j = 0; j < 32; j += 2) {
    float2 a = make_float2(__uint_as_float(v0[j]), __uint_as_float(v0[j + 1]));
    float2 b = make_float2(__uint_as_float(v1[j]), __uint_as_float(v1[j + 1]));
    const float2 ha = __fmul2_rn(a, half2), hb = __fmul2_rn(b, half2);
    const float2 ta = make_float2(tanh_approx(ha.x), tanh_approx(ha.y));
    const float2 tb = make_float2(tanh_approx(hb.x), tanh_approx(hb.y));
    a = __fmul2_rn(a, __ffma2_rn(ta, half2, half2));
    b = __fmul2_rn(b, __ffma2_rn(tb, half2, half2));
    v0[j] = __float_as_uint(a.x); v0[j + 1] = __float_as_uint(a.y);
    v1[j] = __float_as_uint(b.x); v1[j + 1] = __float_as_uint(b.y);
  }
}

// UMMA shared-memory descriptor of a K-major, 128B-swizzled operand tile whose rows are 128 bytes
// (64 bf16) and whose 8-row groups are 1024 bytes apart (dense):
//   bits [0,14) start address >> 4, [16,30) leading byte offset >> 4 (unused for swizzled K-major:
//   1), [32,46) stride byte offset >> 4 (= 1024 >> 4), [46,48) version = 1, [61,64) layout = 2.
__device__ __forceinline__ uint64_t umma_desc_sw128(uint32_t smem_addr) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr >> 4) & 0x3FFF);
  d |= (uint64_t)1 << 16;
  d |= (uint64_t)(1024 >> 4) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)2 << 61;
  return d;
}
// Instruction descriptor, kind::f16: fp32 accumulate (bits 4-5 = 1), A and B bf16 (bits 7-9 and
// 10-12 = 1), both K-major (bits 15, 16 = 0), N >> 3 at bits 17-22, M >> 4 at bits 24-28.
__host__ __device__ constexpr uint32_t umma_idesc_bf16(int m, int n) {
  return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(m >> 4) << 24);
}

// kStages == stages_for(BN): the stand-alone configuration (fills the SM's shared memory).
// kStages == 2: the small-footprint configuration (< 100 KB of shared memory, <= 104 registers) that
// can be co-resident with the scan CTAs of another step running on a second stream.
template <int BN, int kStages>
__global__ void __launch_bounds__(kThreads, kStages == 2 ? 3 : 1)
gemm_tc_kernel(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_w,
               const __grid_constant__ CUtensorMap map_c, const __nv_bfloat16* __restrict__ bias,
               int M, int N, int K, int act_from) {
  constexpr int kTmemCols = tmem_cols_for(BN);
  constexpr uint32_t kABytes = BM * BK * 2;
  constexpr uint32_t kWBytes = BN * BK * 2;
  constexpr uint32_t kSubBytes = BM * kSubN * 2;
  constexpr uint32_t kIdesc = umma_idesc_bf16(BM, BN);

  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t smem_a = smem_base;
  const uint32_t smem_w = smem_a + kStages * kABytes;
  const uint32_t smem_c = smem_w + kStages * kWBytes;
  const uint32_t bars = smem_c + 2 * kSubBytes;
  // barrier layout (8 bytes each): full[kStages] | empty[kStages] | tmem_full[2] | tmem_empty[2]
  auto full_bar = [&](int s) { return bars + 8u * s; };
  auto empty_bar = [&](int s) { return bars + 8u * (kStages + s); };
  auto tfull_bar = [&](int s) { return bars + 8u * (2 * kStages + s); };
  auto tempty_bar = [&](int s) { return bars + 8u * (2 * kStages + 2 + s); };
  const uint32_t tmem_slot = bars + 8u * (2 * kStages + 4);
  uint8_t* smem_gen = smem_raw + (smem_base - smem_u32(smem_raw));   // generic view of smem_base
  volatile uint32_t* tmem_slot_ptr =
      reinterpret_cast<volatile uint32_t*>(smem_gen + (tmem_slot - smem_base));

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int m_tiles = (M + BM - 1) / BM;
  const int n_tiles = (N + BN - 1) / BN;
  const int k_blocks = (K + BK - 1) / BK;
  const int num_tiles = m_tiles * n_tiles;

  if (warp == 0 && elect_one()) {
    prefetch_tmap(&map_a);
    prefetch_tmap(&map_w);
    prefetch_tmap(&map_c);
  }
  if (warp == 1 && elect_one()) {
    for (int s = 0; s < kStages; ++s) {
      mbar_init(full_bar(s), 1);
      mbar_init(empty_bar(s), 1);
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(tfull_bar(s), 1);
      mbar_init(tempty_bar(s), 128);
    }
    fence_barrier_init();
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_slot),
                 "r"((uint32_t)kTmemCols)
                 : "memory");
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
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
        const int m0 = (tile / n_tiles) * BM;
        const int n0 = (tile % n_tiles) * BN;
        for (int kb = 0; kb < k_blocks; ++kb) {
          mbar_wait(empty_bar(stage), phase ^ 1);
          mbar_expect_tx(full_bar(stage), kABytes + kWBytes);
          tma_load_2d(smem_a + stage * kABytes, &map_a, full_bar(stage), kb * BK, m0);
          tma_load_2d(smem_w + stage * kWBytes, &map_w, full_bar(stage), kb * BK, n0);
          if (++stage == kStages) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // ===== MMA issuer =====
    if (elect_one()) {
      int stage = 0;
      uint32_t phase = 0;
      int local = 0;
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++local) {
        const int acc = local & 1;
        const uint32_t acc_phase = (local >> 1) & 1;
        mbar_wait(tempty_bar(acc), acc_phase ^ 1);   // epilogue drained this accumulator
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + acc * BN;
        for (int kb = 0; kb < k_blocks; ++kb) {
          mbar_wait(full_bar(stage), phase);
          tc_fence_after();
          const uint64_t a_desc = umma_desc_sw128(smem_a + stage * kABytes);
          const uint64_t b_desc = umma_desc_sw128(smem_w + stage * kWBytes);
#pragma unroll
          for (int k = 0; k < BK / UMMA_K; ++k) {
            // advance 16 elements = 32 bytes along K inside the swizzle row: +2 in (addr >> 4)
            tc_mma_bf16(d_tmem, a_desc + 2u * k, b_desc + 2u * k, kIdesc, (kb | k) != 0);
          }
          tc_commit(empty_bar(stage));               // smem slot free once these MMAs retire
          if (kb == k_blocks - 1) tc_commit(tfull_bar(acc));
          if (++stage == kStages) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp >= kEpiWarp0) {
    // ===== epilogue =====
    const int ew = warp & 3;                         // TMEM lanes [32 ew, 32 ew + 32)
    const int row = ew * 32 + lane;                  // row of the tile this thread owns
    const bool issuer = threadIdx.x == kEpiWarp0 * 32;
    int local = 0;
    int buf = 0;
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++local) {
      const int acc = local & 1;
      const uint32_t acc_phase = (local >> 1) & 1;
      const int m0 = (tile / n_tiles) * BM;
      const int n0 = (tile % n_tiles) * BN;
      mbar_wait(tfull_bar(acc), acc_phase);
      tc_fence_after();
      const uint32_t t_row = tmem_base + acc * BN + ((uint32_t)(ew * 32) << 16);
#pragma unroll 1
      for (int sub = 0; sub < BN / kSubN; ++sub) {
        if (n0 + sub * kSubN >= N) break;            // whole sub-tile beyond N (uniform)
        uint32_t v0[32], v1[32];
        tc_ld_32x32(t_row + sub * kSubN, v0);
        tc_ld_32x32(t_row + sub * kSubN + 32, v1);
        tc_wait_ld();
        if (sub == BN / kSubN - 1 || n0 + (sub + 1) * kSubN >= N) {
          // accumulator fully read: hand it back to the MMA warp
          tc_fence_before();
          mbar_arrive(tempty_bar(acc));
        }
        if (bias != nullptr) {
          const int nb = n0 + sub * kSubN;
#pragma unroll
          for (int j = 0; j < 32; ++j) {
            const float b0 = nb + j < N ? __bfloat162float(bias[nb + j]) : 0.f;
            const float b1 = nb + 32 + j < N ? __bfloat162float(bias[nb + 32 + j]) : 0.f;
            v0[j] = __float_as_uint(__uint_as_float(v0[j]) + b0);
            v1[j] = __float_as_uint(__uint_as_float(v1[j]) + b1);
          }
        }
        if (n0 + sub * kSubN >= act_from) epi_silu(v0, v1);   // gate columns (uniform per sub-tile)
        // the store that last read this staging buffer (two sub-tiles ago) must be done reading
        if (issuer) tma_store_wait_read<1>();
        epi_bar_sync();
        uint8_t* dst = smem_gen + (smem_c - smem_base) + buf * kSubBytes + row * 128;
#pragma unroll
        for (int c = 0; c < 8; ++c) {                // 8 chunks of 16 bytes = 8 bf16 each
          uint32_t p[4];
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            const int j = c * 8 + q * 2;
            const float lo = __uint_as_float(j < 32 ? v0[j] : v1[j - 32]);
            const float hi = __uint_as_float(j + 1 < 32 ? v0[j + 1] : v1[j + 1 - 32]);
            __nv_bfloat162 h = __floats2bfloat162_rn(lo, hi);
            p[q] = *reinterpret_cast<uint32_t*>(&h);
          }
          // 128B swizzle: 16-byte chunk index XOR (row mod 8)
          *reinterpret_cast<uint4*>(dst + ((c ^ (row & 7)) << 4)) = make_uint4(p[0], p[1], p[2], p[3]);
        }
        fence_proxy_async_smem();
        epi_bar_sync();
        if (issuer) {
          tma_store_2d(&map_c, smem_c + buf * kSubBytes, n0 + sub * kSubN, m0);
          tma_store_commit();
        }
        buf ^= 1;
      }
    }
    if (issuer) tma_store_wait_all();
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base),
                 "r"((uint32_t)kTmemCols)
                 : "memory");
  }
}

// ---- CTA-pair variant (cta_group::2) ---------------------------------------------------------
// Two CTAs of a cluster (the two SMs of a TPC) work on one 256 x BN tile: CTA r owns rows
// [128 r, 128 r + 128) of the tile (its A rows in its shared memory, its accumulator rows in its
// TMEM) and HALF of the W tile (rows [r BN/2, (r+1) BN/2)); the tensor cores of both SMs read both
// halves.  Per 128 x BN of output an SM therefore ingests 128 A rows + BN/2 W rows instead of
// 128 + BN: the 1-CTA kernel is bound by the ~12.4 TB/s the L2 can deliver (in_proj: 288 KB per
// 128 x 256 tile -> 109 us; out_proj: 491 KB per 128 x 192 tile -> 62 us, both exactly what is
// measured), and the pair needs 1.5x / 1.43x less.
//   * both CTAs run a TMA producer; every load signals the LEADER's (rank 0) full barrier
//     (cp.async.bulk.tensor .cta_group::2), the leader expects the bytes of both CTAs;
//   * only the leader issues tcgen05.mma.cta_group::2 (M = 256); its commits are multicast to the
//     empty / accumulator-full barriers of both CTAs;
//   * the epilogue warps of both CTAs hand the accumulator back on the leader's barrier
//     (remote mbarrier.arrive for rank 1);  TMEM is allocated / freed by warp 1 of both CTAs.
__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ uint32_t mapa_cluster(uint32_t smem_addr, uint32_t rank) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(smem_addr), "r"(rank));
  return r;
}
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t cluster_addr) {
  // default .release.cta semantics: the hand-over orders TMEM reads (tcgen05 fences), not generic memory;
  // a cluster-scope release here costs a memory barrier per thread per tile (measured: +20 % kernel time)
  asm volatile("mbarrier.arrive.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ void tma_load_2d_pair(uint32_t dst, const CUtensorMap* map, uint32_t leader_bar,
                                                 int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      ::"r"(dst), "l"(map), "r"(leader_bar), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ void tc_commit_pair(uint32_t bar) {   // arrives on `bar` in both CTAs
  asm volatile(
      "tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
      ::"r"(bar), "h"((uint16_t)3)
      : "memory");
}
__device__ __forceinline__ void tc_mma_bf16_pair(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc,
                                                 uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
      : "memory");
}

__host__ __device__ constexpr size_t smem_bytes_pair(int bn, int stages) {
  return 1024 + (size_t)stages * (BM * BK * 2 + (bn / 2) * BK * 2) + 2 * (BM * kSubN * 2) + 256;
}
__host__ __device__ constexpr int stages_pair(int) { return 6; }   // 6 x (16 KB A + <= 16 KB W half)

template <int BN, int kStages>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(kThreads, kStages <= 3 ? 3 : 1)
gemm_tc_pair_kernel(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_w,
                    const __grid_constant__ CUtensorMap map_c, const __nv_bfloat16* __restrict__ bias,
                    int M, int N, int K, int act_from, int interleave) {
  constexpr int kTmemCols = tmem_cols_for(BN);
  constexpr uint32_t kABytes = BM * BK * 2;
  constexpr uint32_t kWBytes = (BN / 2) * BK * 2;          // this CTA's half of the W tile
  constexpr uint32_t kSubBytes = BM * kSubN * 2;
  constexpr uint32_t kIdesc = umma_idesc_bf16(2 * BM, BN);  // M = 256 across the pair

  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t smem_a = smem_base;
  const uint32_t smem_w = smem_a + kStages * kABytes;
  const uint32_t smem_c = smem_w + kStages * kWBytes;
  const uint32_t bars = smem_c + 2 * kSubBytes;
  auto full_bar = [&](int s) { return bars + 8u * s; };
  auto empty_bar = [&](int s) { return bars + 8u * (kStages + s); };
  auto tfull_bar = [&](int s) { return bars + 8u * (2 * kStages + s); };
  auto tempty_bar = [&](int s) { return bars + 8u * (2 * kStages + 2 + s); };
  const uint32_t tmem_slot = bars + 8u * (2 * kStages + 4);
  uint8_t* smem_gen = smem_raw + (smem_base - smem_u32(smem_raw));
  volatile uint32_t* tmem_slot_ptr =
      reinterpret_cast<volatile uint32_t*>(smem_gen + (tmem_slot - smem_base));

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const bool leader = rank == 0;
  const int pair = blockIdx.x >> 1;
  const int num_pairs = gridDim.x >> 1;
  const int m_tiles = (M + 2 * BM - 1) / (2 * BM);
  const int n_tiles = (N + BN - 1) / BN;
  const int k_blocks = (K + BK - 1) / BK;
  const int num_tiles = m_tiles * n_tiles;

  if (warp == 0 && elect_one()) {
    prefetch_tmap(&map_a);
    prefetch_tmap(&map_w);
    prefetch_tmap(&map_c);
  }
  if (warp == 1 && elect_one()) {
    for (int s = 0; s < kStages; ++s) {
      mbar_init(full_bar(s), 1);       // leader: its own arrive.expect_tx; bytes come from both CTAs
      mbar_init(empty_bar(s), 1);      // one multicast commit per phase
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(tfull_bar(s), 1);
      mbar_init(tempty_bar(s), 8);     // leader: the 4 epilogue warps of both CTAs
    }
    fence_barrier_init();
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_slot),
                 "r"((uint32_t)kTmemCols)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  cluster_sync_all();                  // barriers of both CTAs initialised before any remote signal
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot_ptr;

  if (warp == 0) {
    // ===== TMA producer (both CTAs) =====
    if (elect_one()) {
      int stage = 0;
      uint32_t phase = 0;
      for (int tile = pair; tile < num_tiles; tile += num_pairs) {
        const int m0 = (tile / n_tiles) * (2 * BM) + (int)rank * BM;
        // W rows of this CTA's half of the tile; interleaved (in_proj with the gate): the leader holds 128 x
        // columns, the peer the 128 z columns of the same channels, so every tile carries the same epilogue work
        const int n0 = interleave ? (tile % n_tiles) * (BN / 2) + (int)rank * act_from
                                  : (tile % n_tiles) * BN + (int)rank * (BN / 2);
        for (int kb = 0; kb < k_blocks; ++kb) {
          mbar_wait(empty_bar(stage), phase ^ 1);
          if (leader) mbar_expect_tx(full_bar(stage), 2 * (kABytes + kWBytes));
          const uint32_t lbar = mapa_cluster(full_bar(stage), 0);
          tma_load_2d_pair(smem_a + stage * kABytes, &map_a, lbar, kb * BK, m0);
          tma_load_2d_pair(smem_w + stage * kWBytes, &map_w, lbar, kb * BK, n0);
          if (++stage == kStages) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // ===== MMA issuer (leader CTA only) =====
    if (leader && elect_one()) {
      int stage = 0;
      uint32_t phase = 0;
      int local = 0;
      for (int tile = pair; tile < num_tiles; tile += num_pairs, ++local) {
        const int acc = local & 1;
        const uint32_t acc_phase = (local >> 1) & 1;
        mbar_wait(tempty_bar(acc), acc_phase ^ 1);   // both epilogues drained this accumulator
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + acc * BN;
        for (int kb = 0; kb < k_blocks; ++kb) {
          mbar_wait(full_bar(stage), phase);
          tc_fence_after();
          const uint64_t a_desc = umma_desc_sw128(smem_a + stage * kABytes);
          const uint64_t b_desc = umma_desc_sw128(smem_w + stage * kWBytes);
#pragma unroll
          for (int k = 0; k < BK / UMMA_K; ++k)
            tc_mma_bf16_pair(d_tmem, a_desc + 2u * k, b_desc + 2u * k, kIdesc, (kb | k) != 0);
          tc_commit_pair(empty_bar(stage));          // frees the slot in both CTAs
          if (kb == k_blocks - 1) tc_commit_pair(tfull_bar(acc));
          if (++stage == kStages) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp >= kEpiWarp0) {
    // ===== epilogue (both CTAs, own 128 rows) =====
    const int ew = warp & 3;
    const int row = ew * 32 + lane;
    const bool issuer = threadIdx.x == kEpiWarp0 * 32;
    int local = 0;
    int buf = 0;
    for (int tile = pair; tile < num_tiles; tile += num_pairs, ++local) {
      const int acc = local & 1;
      const uint32_t acc_phase = (local >> 1) & 1;
      const int m0 = (tile / n_tiles) * (2 * BM) + (int)rank * BM;
      const int n0 = (tile % n_tiles) * BN;
      mbar_wait(tfull_bar(acc), acc_phase);
      tc_fence_after();
      const uint32_t t_row = tmem_base + acc * BN + ((uint32_t)(ew * 32) << 16);
#pragma unroll 1
      for (int sub = 0; sub < BN / kSubN; ++sub) {
        // first output column of the sub-tile (interleaved: accumulator columns [0, BN/2) are x, the rest z)
        const int nb = interleave ? (n0 >> 1) + (sub & 1) * kSubN + (sub >= BN / kSubN / 2 ? act_from : 0)
                                  : n0 + sub * kSubN;
        const bool beyond = nb >= N;                 // whole sub-tile beyond N (uniform)
        uint32_t v0[32], v1[32];
        if (!beyond) {
          tc_ld_32x32(t_row + sub * kSubN, v0);
          tc_ld_32x32(t_row + sub * kSubN + 32, v1);
          tc_wait_ld();
        }
        if (sub == BN / kSubN - 1) {                 // accumulator fully read: one arrival per warp
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive_cluster(mapa_cluster(tempty_bar(acc), 0));
        }
        if (beyond) continue;
        if (bias != nullptr) {
#pragma unroll
          for (int j = 0; j < 32; ++j) {
            const float b0 = nb + j < N ? __bfloat162float(bias[nb + j]) : 0.f;
            const float b1 = nb + 32 + j < N ? __bfloat162float(bias[nb + 32 + j]) : 0.f;
            v0[j] = __float_as_uint(__uint_as_float(v0[j]) + b0);
            v1[j] = __float_as_uint(__uint_as_float(v1[j]) + b1);
          }
        }
        if (nb >= act_from) epi_silu(v0, v1);        // gate columns (uniform per sub-tile)
        if (issuer) tma_store_wait_read<1>();
        epi_bar_sync();
        uint8_t* dst = smem_gen + (smem_c - smem_base) + buf * kSubBytes + row * 128;
#pragma unroll
        for (int c = 0; c < 8; ++c) {
          uint32_t p[4];
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            const int j = c * 8 + q * 2;
            const float lo = __uint_as_float(j < 32 ? v0[j] : v1[j - 32]);
            const float hi = __uint_as_float(j + 1 < 32 ? v0[j + 1] : v1[j + 1 - 32]);
            __nv_bfloat162 h = __floats2bfloat162_rn(lo, hi);
            p[q] = *reinterpret_cast<uint32_t*>(&h);
          }
          *reinterpret_cast<uint4*>(dst + ((c ^ (row & 7)) << 4)) = make_uint4(p[0], p[1], p[2], p[3]);
        }
        fence_proxy_async_smem();
        epi_bar_sync();
        if (issuer) {
          tma_store_2d(&map_c, smem_c + buf * kSubBytes, nb, m0);
          tma_store_commit();
        }
        buf ^= 1;
      }
    }
    if (issuer) tma_store_wait_all();
  }

  __syncwarp();                        // the elected lanes rejoin their warps before the aligned barrier
  tc_fence_before();
  cluster_sync_all();                  // the leader's MMAs read the peer's shared memory and TMEM
  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base),
                 "r"((uint32_t)kTmemCols)
                 : "memory");
  }
}

// ---- causal conv1d + SiLU fused into the x_proj projection ------------------------------------
// x_dbl[M, BN] = SiLU(conv(x))[M, K] * W[BN, K]^T, and xc = SiLU(conv(x)) is written on the way (the
// scan reads it as u).  Stands in for the causal_conv1d_fn -> rearrange -> x_proj sequence of
// models/videomamba/mamba_simple.py:381-409 for the stateless forward walk (no conv_state in / out,
// no reversal: those keep the separate kernels).  The conv output tile never returns from HBM:
//   warp 0     TMA producer: per k-block (64 channels) the RAW x rows [m0 - 3, m0 + 128) x 64
//              (unswizzled box; rows before the tensor start are zero-filled) and the W tile.
//   warps 6-21 conv (two groups of 8 warps, alternate k-blocks): thread = 4 channels x 8 rows; slides
//              the 4-tap window over its 11 raw rows
//              (same FFMA2 order as conv1d_ring_kernel: bit-identical xc), writes the bf16 results
//              to the k-block's A stage in the UMMA 128-byte-swizzle layout AND to xc in HBM, then
//              fence.proxy.async + mbarrier arrive.  Rows where a new sequence starts (row % L < 3)
//              drop the taps that would reach into the previous clip.
//   warp 1     MMA issuer (tcgen05.mma M 128 x N BN x K 16 from the A stage and the W tile).
//   warps 2-5  epilogue: TMEM -> bf16 -> staging -> TMA store of the x_dbl tile.
// HBM traffic per token: x read once, xc written once, x_dbl written once (the separate kernels
// re-read xc: + Di * 2 bytes).
#ifndef VMB_CX_GROUPS
#define VMB_CX_GROUPS 2
#endif
constexpr int kCxGroups = VMB_CX_GROUPS;                                 // conv warp groups (8 warps each)
constexpr int kCxThreads = 192 + 256 * kCxGroups;   // TMA, MMA, 4 epilogue warps, kCxGroups x 8 conv warps
constexpr int kCxRawRows = BM + 3;
constexpr uint32_t kCxRawBytes = kCxRawRows * BK * 2;                       // 16 768
constexpr uint32_t kCxRawStage = (kCxRawBytes + 1023u) / 1024u * 1024u;     // 17 408
constexpr int kCxAStages = 2;            // conv-output (A operand) ring: held only from conv to MMA
__host__ __device__ constexpr size_t smem_bytes_cx(int bn, int stages, int K) {
  return 1024 + (size_t)stages * (kCxRawStage + bn * BK * 2) + (size_t)kCxAStages * (BM * BK * 2) +
         BM * kSubN * 2 + 512 + (size_t)K * 10;       // conv taps (K x 4) and bias (K), bf16
}

template <int BN, int kStages, bool kSilu>
__global__ void __launch_bounds__(kCxThreads, 1)
conv_xproj_kernel(const __grid_constant__ CUtensorMap map_x, const __grid_constant__ CUtensorMap map_w,
                  const __grid_constant__ CUtensorMap map_c, const __nv_bfloat16* __restrict__ cw,
                  const __nv_bfloat16* __restrict__ cb, __nv_bfloat16* __restrict__ xc, int64_t xc_ld,
                  int M, int K, int L) {
  static_assert(BN <= kSubN, "conv_xproj: one epilogue sub-tile");
  constexpr int kTmemCols = tmem_cols_for(BN);
  constexpr uint32_t kABytes = BM * BK * 2;
  constexpr uint32_t kWBytes = BN * BK * 2;
  constexpr uint32_t kIdesc = umma_idesc_bf16(BM, BN);

  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t smem_x = smem_base;                                   // raw x stages
  const uint32_t smem_w = smem_x + kStages * kCxRawStage;
  const uint32_t smem_a = smem_w + kStages * kWBytes;
  const uint32_t smem_c = smem_a + kCxAStages * kABytes;
  const uint32_t bars = smem_c + BM * kSubN * 2;
  auto raw_full = [&](int s) { return bars + 8u * s; };
  auto raw_empty = [&](int s) { return bars + 8u * (kStages + s); };
  auto a_full = [&](int s) { return bars + 8u * (2 * kStages + s); };
  auto a_empty = [&](int s) { return bars + 8u * (2 * kStages + kCxAStages + s); };
  auto tfull_bar = [&](int s) { return bars + 8u * (2 * kStages + 2 * kCxAStages + s); };
  auto tempty_bar = [&](int s) { return bars + 8u * (2 * kStages + 2 * kCxAStages + 2 + s); };
  const uint32_t tmem_slot = bars + 8u * (2 * kStages + 2 * kCxAStages + 4);
  const uint32_t smem_cw = bars + 512;                                 // conv taps, then bias
  uint8_t* smem_gen = smem_raw + (smem_base - smem_u32(smem_raw));
  volatile uint32_t* tmem_slot_ptr =
      reinterpret_cast<volatile uint32_t*>(smem_gen + (tmem_slot - smem_base));

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int m_tiles = (M + BM - 1) / BM;
  const int k_blocks = K / BK;

  if (warp == 0 && elect_one()) {
    prefetch_tmap(&map_x);
    prefetch_tmap(&map_w);
    prefetch_tmap(&map_c);
  }
  if (warp == 1 && elect_one()) {
    for (int s = 0; s < kStages; ++s) {
      mbar_init(raw_full(s), 1);
      mbar_init(raw_empty(s), 256 + 1);     // the conv threads (raw rows read) + the MMA commit (W read)
    }
    for (int s = 0; s < kCxAStages; ++s) {
      mbar_init(a_full(s), 256);
      mbar_init(a_empty(s), 1);
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(tfull_bar(s), 1);
      mbar_init(tempty_bar(s), 128);
    }
    fence_barrier_init();
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_slot),
                 "r"((uint32_t)kTmemCols)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  {   // conv taps (K x 4) and bias (K) into shared memory, once per CTA
    uint4* dst = reinterpret_cast<uint4*>(smem_gen + (smem_cw - smem_base));
    const int nw = K / 2;                            // 16-byte chunks of taps (4 taps x 2 channels each)
    for (int q = threadIdx.x; q < nw; q += blockDim.x) dst[q] = __ldg(reinterpret_cast<const uint4*>(cw) + q);
    const int nb = K / 8;
    for (int q = threadIdx.x; q < nb; q += blockDim.x)
      dst[nw + q] = cb != nullptr ? __ldg(reinterpret_cast<const uint4*>(cb) + q) : make_uint4(0u, 0u, 0u, 0u);
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot_ptr;

  // tiles are walked back to front: in_proj wrote the last rows of x most recently
  auto tile_m0 = [&](int i) { return (m_tiles - 1 - i) * BM; };

  if (warp == 0) {
    // ===== TMA producer =====
    if (elect_one()) {
      int stage = 0;
      uint32_t phase = 0;
      for (int i = blockIdx.x; i < m_tiles; i += gridDim.x) {
        const int m0 = tile_m0(i);
        for (int kb = 0; kb < k_blocks; ++kb) {
          mbar_wait(raw_empty(stage), phase ^ 1);
          mbar_expect_tx(raw_full(stage), kCxRawBytes + kWBytes);
          tma_load_2d(smem_x + stage * kCxRawStage, &map_x, raw_full(stage), kb * BK, m0 - 3);
          tma_load_2d(smem_w + stage * kWBytes, &map_w, raw_full(stage), kb * BK, 0);
          if (++stage == kStages) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // ===== MMA issuer =====
    if (elect_one()) {
      int n = 0;                                     // running k-block counter
      int local = 0;
      for (int i = blockIdx.x; i < m_tiles; i += gridDim.x, ++local) {
        const int acc = local & 1;
        const uint32_t acc_phase = (local >> 1) & 1;
        mbar_wait(tempty_bar(acc), acc_phase ^ 1);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + acc * BN;
        for (int kb = 0; kb < k_blocks; ++kb, ++n) {
          const int rs = n % kStages, as = n % kCxAStages;
          mbar_wait(raw_full(rs), (uint32_t)(n / kStages) & 1u);        // W tile landed
          mbar_wait(a_full(as), (uint32_t)(n / kCxAStages) & 1u);       // conv output of this k-block staged
          tc_fence_after();
          const uint64_t a_desc = umma_desc_sw128(smem_a + as * kABytes);
          const uint64_t b_desc = umma_desc_sw128(smem_w + rs * kWBytes);
#pragma unroll
          for (int k = 0; k < BK / UMMA_K; ++k)
            tc_mma_bf16(d_tmem, a_desc + 2u * k, b_desc + 2u * k, kIdesc, (kb | k) != 0);
          tc_commit(a_empty(as));
          tc_commit(raw_empty(rs));
          if (kb == k_blocks - 1) tc_commit(tfull_bar(acc));
        }
      }
    }
  } else if (warp < 6) {
    // ===== epilogue: one 128 x BN tile of x_dbl =====
    const int ew = warp & 3;
    const int row = ew * 32 + lane;
    const bool issuer = threadIdx.x == kEpiWarp0 * 32;
    int local = 0;
    for (int i = blockIdx.x; i < m_tiles; i += gridDim.x, ++local) {
      const int acc = local & 1;
      const uint32_t acc_phase = (local >> 1) & 1;
      const int m0 = tile_m0(i);
      mbar_wait(tfull_bar(acc), acc_phase);
      tc_fence_after();
      const uint32_t t_row = tmem_base + acc * BN + ((uint32_t)(ew * 32) << 16);
      uint32_t v0[32], v1[32];
      tc_ld_32x32(t_row, v0);
      if (BN > 32) tc_ld_32x32(t_row + 32, v1);
      tc_wait_ld();
      tc_fence_before();
      mbar_arrive(tempty_bar(acc));
      if (issuer) tma_store_wait_read<0>();          // the previous tile's store has read the staging buffer
      epi_bar_sync();
      uint8_t* dst = smem_gen + (smem_c - smem_base) + row * 128;
#pragma unroll
      for (int c = 0; c < BN / 8; ++c) {
        uint32_t p[4];
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          const int j = c * 8 + q * 2;
          const float lo = __uint_as_float(j < 32 ? v0[j] : v1[j - 32]);
          const float hi = __uint_as_float(j + 1 < 32 ? v0[j + 1] : v1[j + 1 - 32]);
          __nv_bfloat162 h = __floats2bfloat162_rn(lo, hi);
          p[q] = *reinterpret_cast<uint32_t*>(&h);
        }
        *reinterpret_cast<uint4*>(dst + ((c ^ (row & 7)) << 4)) = make_uint4(p[0], p[1], p[2], p[3]);
      }
      fence_proxy_async_smem();
      epi_bar_sync();
      if (issuer) {
        tma_store_2d(&map_c, smem_c, 0, m0);
        tma_store_commit();
      }
    }
    if (issuer) tma_store_wait_all();
  } else {
    // ===== conv warps: kCxGroups groups of 256 threads take the k-blocks round robin.  A conv warp issues
    // about one instruction per 6 cycles (dependent FFMA2 / MUFU chains), so throughput scales with the
    // number of resident conv warps: thread = 4 channels (quad q4 of the k-block) x rows [8 rg, 8 rg + 8),
    // 16 conv warps per SM at <= 88 registers =====
    const int ct = threadIdx.x - 6 * 32;
    const int grp = ct >> 8;
    const int q4 = ct & 15, rg = (ct & 255) >> 4;
    const uint4* cws = reinterpret_cast<const uint4*>(smem_gen + (smem_cw - smem_base));
    int n = 0;                                       // running k-block counter (all roles agree on it)
    for (int i = blockIdx.x; i < m_tiles; i += gridDim.x) {
      const int m0 = tile_m0(i);
      const int mrow = m0 + rg * 8;                  // first output row of this thread
      const int t_first = mrow % L;                  // its position in its sequence
      for (int kb = 0; kb < k_blocks; ++kb, ++n) {
        if (n % kCxGroups != grp) continue;
        const int stage = n % kStages;                          // raw ring
        const uint32_t phase = (uint32_t)(n / kStages) & 1u;
        const int as = n % kCxAStages;                          // A ring
        const uint32_t aphase = (uint32_t)(n / kCxAStages) & 1u;
        const int c0 = kb * BK + q4 * 4;             // first of this thread's 4 channels
        float2 w2[2][4], b2[2];
        {
          const uint4 wv[2] = {cws[c0 / 2], cws[c0 / 2 + 1]};
#pragma unroll
          for (int p = 0; p < 2; ++p) {              // 8 bf16: channel 2p taps 0..3, channel 2p+1 taps 0..3
            const uint32_t q[4] = {wv[p].x, wv[p].y, wv[p].z, wv[p].w};
            const float e0[4] = {__uint_as_float(q[0] << 16), __uint_as_float(q[0] & 0xffff0000u),
                                 __uint_as_float(q[1] << 16), __uint_as_float(q[1] & 0xffff0000u)};
            const float e1[4] = {__uint_as_float(q[2] << 16), __uint_as_float(q[2] & 0xffff0000u),
                                 __uint_as_float(q[3] << 16), __uint_as_float(q[3] & 0xffff0000u)};
#pragma unroll
            for (int k = 0; k < 4; ++k) w2[p][k] = make_float2(e0[k], e1[k]);
          }
          const uint2 bv = reinterpret_cast<const uint2*>(cws + K / 2)[c0 / 4];
          b2[0] = make_float2(__uint_as_float(bv.x << 16), __uint_as_float(bv.x & 0xffff0000u));
          b2[1] = make_float2(__uint_as_float(bv.y << 16), __uint_as_float(bv.y & 0xffff0000u));
        }
        mbar_wait(raw_full(stage), phase);
        mbar_wait(a_empty(as), aphase ^ 1);
        const uint8_t* rawp = smem_gen + (smem_x - smem_base) + stage * kCxRawStage + q4 * 8;
        uint8_t* ap = smem_gen + (smem_a - smem_base) + as * kABytes;
        auto unpack = [](const uint2& v, float2 (&f)[2]) {
          f[0] = make_float2(__uint_as_float(v.x << 16), __uint_as_float(v.x & 0xffff0000u));
          f[1] = make_float2(__uint_as_float(v.y << 16), __uint_as_float(v.y & 0xffff0000u));
        };
        // raw row j of the stage is tensor row m0 - 3 + j: this thread's window starts at 8 rg.
        // kEdge: a sequence starts inside (or just before) these rows -- taps that would reach into the
        // previous clip are dropped; the common case has no such row and no checks.
        auto run_rows = [&](auto edge_tag) {
          constexpr bool kEdge = decltype(edge_tag)::value;
          float2 win[3][2];
#pragma unroll
          for (int k = 0; k < 3; ++k) {
            unpack(*reinterpret_cast<const uint2*>(rawp + (rg * 8 + k) * 128), win[k]);
            if (kEdge && t_first - 3 + k < 0) {
              win[k][0] = make_float2(0.f, 0.f);
              win[k][1] = make_float2(0.f, 0.f);
            }
          }
          int t = t_first;
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            float2 xin[2];
            unpack(*reinterpret_cast<const uint2*>(rawp + (rg * 8 + 3 + j) * 128), xin);
            if (kEdge && t == L) {                   // a new sequence starts at this row
              t = 0;
#pragma unroll
              for (int k = 0; k < 3; ++k) {
                win[k][0] = make_float2(0.f, 0.f);
                win[k][1] = make_float2(0.f, 0.f);
              }
            }
            uint32_t ov[2];
#pragma unroll
            for (int q = 0; q < 2; ++q) {
              float2 acc = __ffma2_rn(w2[q][0], win[0][q], b2[q]);
              acc = __ffma2_rn(w2[q][1], win[1][q], acc);
              acc = __ffma2_rn(w2[q][2], win[2][q], acc);
              acc = __ffma2_rn(w2[q][3], xin[q], acc);
              if (kSilu) {   // x * (0.5 * tanh(x / 2) + 0.5), packed: same per-lane arithmetic as silu_fast
                const float2 hx = __fmul2_rn(acc, make_float2(0.5f, 0.5f));
                const float2 th = make_float2(tanh_approx(hx.x), tanh_approx(hx.y));
                acc = __fmul2_rn(acc, __ffma2_rn(th, make_float2(0.5f, 0.5f), make_float2(0.5f, 0.5f)));
              }
              const __nv_bfloat162 h = __floats2bfloat162_rn(acc.x, acc.y);
              ov[q] = *reinterpret_cast<const uint32_t*>(&h);
              win[0][q] = win[1][q]; win[1][q] = win[2][q]; win[2][q] = xin[q];
            }
            const int r = rg * 8 + j;                // row of the tile
            const uint2 out = make_uint2(ov[0], ov[1]);
            // 128B swizzle of the A stage: 16-byte chunk (q4 >> 1) XOR (row mod 8), this quad's half of it
            *reinterpret_cast<uint2*>(ap + r * 128 + ((((q4 >> 1) ^ (r & 7)) << 4) | ((q4 & 1) << 3))) = out;
            if (!kEdge || m0 + r < M) *reinterpret_cast<uint2*>(xc + (int64_t)(m0 + r) * xc_ld + c0) = out;
            if (kEdge) ++t;
          }
        };
        if (t_first < 3 || t_first + 8 > L || m0 + BM > M) run_rows(std::true_type{});
        else run_rows(std::false_type{});
        fence_proxy_async_smem();                    // generic-proxy writes of the A stage -> tensor core reads
        mbar_arrive(a_full(as));
        mbar_arrive(raw_empty(stage));
      }
    }
  }

  __syncwarp();
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base),
                 "r"((uint32_t)kTmemCols)
                 : "memory");
  }
}

// ---- host side: tensor maps ----------------------------------------------------------------
using EncodeFn = CUresult (*)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                              const cuuint64_t*, const cuuint32_t*, const cuuint32_t*,
                              CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion,
                              CUtensorMapFloatOOBfill);

EncodeFn encode_fn() {
  static EncodeFn fn = [] {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess ||
        q != cudaDriverEntryPointSuccess)
      p = nullptr;
    return reinterpret_cast<EncodeFn>(p);
  }();
  return fn;
}

struct MapKey {
  const void* ptr; int64_t rows, cols, ld; int box_rows;
  bool operator==(const MapKey& o) const {
    return ptr == o.ptr && rows == o.rows && cols == o.cols && ld == o.ld && box_rows == o.box_rows;
  }
};
struct MapKeyHash {
  size_t operator()(const MapKey& k) const {
    size_t h = std::hash<const void*>()(k.ptr);
    auto mix = [&](int64_t v) { h ^= std::hash<int64_t>()(v) + 0x9e3779b97f4a7c15ull + (h << 6) + (h >> 2); };
    mix(k.rows); mix(k.cols); mix(k.ld); mix(k.box_rows);
    return h;
  }
};

// 2-D bf16 row-major tensor (rows x cols, row stride ld elements), box = box_rows x 64, 128B swizzle.
int make_map(CUtensorMap* out, const void* ptr, int64_t rows, int64_t cols, int64_t ld, int box_rows) {
  static std::mutex mu;
  static std::unordered_map<MapKey, CUtensorMap, MapKeyHash> cache;
  const MapKey key{ptr, rows, cols, ld, box_rows};
  {
    std::lock_guard<std::mutex> lk(mu);
    auto it = cache.find(key);
    if (it != cache.end()) { *out = it->second; return VMB_OK; }
  }
  EncodeFn fn = encode_fn();
  if (!fn) { set_error("gemm_tc: cuTensorMapEncodeTiled entry point not available"); return VMB_ERR_CUDA; }
  const cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
  const cuuint64_t strides[1] = {(cuuint64_t)ld * 2};
  const cuuint32_t box[2] = {(cuuint32_t)BK, (cuuint32_t)box_rows};
  const cuuint32_t estr[2] = {1, 1};
  const CUresult r = fn(out, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(ptr), dims, strides,
                        box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                        CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) { set_error("gemm_tc: cuTensorMapEncodeTiled failed (%d)", (int)r); return VMB_ERR_CUDA; }
  std::lock_guard<std::mutex> lk(mu);
  if (cache.size() > 4096) cache.clear();
  cache.emplace(key, *out);
  return VMB_OK;
}

}  // namespace

// 2-D bf16 row-major map with 64-column (128-byte, 128B-swizzled) boxes of box_rows rows (wgrad_tc.cu).
int make_tensor_map_2d_bf16_sw128(CUtensorMap* out, const void* ptr, int64_t rows, int64_t cols, int64_t ld,
                                  int box_rows) {
  return make_map(out, ptr, rows, cols, ld, box_rows);
}

// 3-D bf16 tensor map (innermost dimension contiguous), box = box0 x box1 x 1, used by the scan's
// TMA staging (scan_fast.cu).  Cached like the 2-D maps.
int make_tensor_map_3d_bf16(CUtensorMap* out, const void* ptr, uint64_t d0, uint64_t d1, uint64_t d2,
                            uint64_t stride1_bytes, uint64_t stride2_bytes, uint32_t box0,
                            uint32_t box1, bool swizzle128) {
  struct Key {
    const void* ptr; uint64_t d0, d1, d2, s1, s2; uint32_t b0, b1; bool sw;
    bool operator==(const Key& o) const {
      return ptr == o.ptr && d0 == o.d0 && d1 == o.d1 && d2 == o.d2 && s1 == o.s1 && s2 == o.s2 &&
             b0 == o.b0 && b1 == o.b1 && sw == o.sw;
    }
  };
  struct KeyHash {
    size_t operator()(const Key& k) const {
      size_t h = std::hash<const void*>()(k.ptr);
      auto mix = [&](uint64_t v) { h ^= std::hash<uint64_t>()(v) + 0x9e3779b97f4a7c15ull + (h << 6) + (h >> 2); };
      mix(k.d0); mix(k.d1); mix(k.d2); mix(k.s1); mix(k.s2); mix(k.b0); mix(k.b1); mix(k.sw);
      return h;
    }
  };
  static std::mutex mu;
  static std::unordered_map<Key, CUtensorMap, KeyHash> cache;
  const Key key{ptr, d0, d1, d2, stride1_bytes, stride2_bytes, box0, box1, swizzle128};
  {
    std::lock_guard<std::mutex> lk(mu);
    auto it = cache.find(key);
    if (it != cache.end()) { *out = it->second; return VMB_OK; }
  }
  EncodeFn fn = encode_fn();
  if (!fn) { set_error("tensor map: cuTensorMapEncodeTiled entry point not available"); return VMB_ERR_CUDA; }
  const cuuint64_t dims[3] = {d0, d1, d2};
  const cuuint64_t strides[2] = {stride1_bytes, stride2_bytes};
  const cuuint32_t box[3] = {box0, box1, 1};
  const cuuint32_t estr[3] = {1, 1, 1};
  const CUresult r = fn(out, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, const_cast<void*>(ptr), dims, strides, box,
                        estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                        swizzle128 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_NONE,
                        CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) { set_error("tensor map: cuTensorMapEncodeTiled (3-D) failed (%d)", (int)r); return VMB_ERR_CUDA; }
  std::lock_guard<std::mutex> lk(mu);
  if (cache.size() > 4096) cache.clear();
  cache.emplace(key, *out);
  return VMB_OK;
}

namespace {

template <int BN, int kStages>
int launch(const void* A, int64_t lda, const void* W, int64_t ldw, const void* bias, void* C,
           int64_t ldc, int64_t M, int N, int K, int act_from, cudaStream_t st) {
  CUtensorMap ma, mw, mc;
  int rc;
  if ((rc = make_map(&ma, A, M, K, lda, BM))) return rc;
  if ((rc = make_map(&mw, W, N, K, ldw, BN))) return rc;
  if ((rc = make_map(&mc, C, M, N, ldc, BM))) return rc;
  static bool attr_set[64] = {false};
  constexpr size_t smem = smem_bytes_for(BN, kStages);
  int dev = 0;
  VMB_CUDA(cudaGetDevice(&dev));
  if (dev < 0 || dev >= 64 || !attr_set[dev]) {
    VMB_CUDA(cudaFuncSetAttribute(gemm_tc_kernel<BN, kStages>,
                                  cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    if (dev >= 0 && dev < 64) attr_set[dev] = true;
  }
  const int64_t tiles = ((M + BM - 1) / BM) * ((N + BN - 1) / BN);
  const int grid = (int)std::min<int64_t>(tiles, sm_count());
  gemm_tc_kernel<BN, kStages><<<grid, kThreads, smem, st>>>(ma, mw, mc, (const __nv_bfloat16*)bias,
                                                            (int)M, N, K, act_from);
  VMB_LAUNCH_CHECK("gemm_tc_kernel");
  return VMB_OK;
}

template <int BN, int kStages>
int launch_pair(const void* A, int64_t lda, const void* W, int64_t ldw, const void* bias, void* C,
                int64_t ldc, int64_t M, int N, int K, int act_from, cudaStream_t st) {
  CUtensorMap ma, mw, mc;
  int rc;
  if ((rc = make_map(&ma, A, M, K, lda, BM))) return rc;
  if ((rc = make_map(&mw, W, N, K, ldw, BN / 2))) return rc;
  if ((rc = make_map(&mc, C, M, N, ldc, BM))) return rc;
  static bool attr_set[64] = {false};
  constexpr size_t smem = smem_bytes_pair(BN, kStages);
  static_assert(smem <= 227 * 1024, "pair kernel: shared memory budget");
  int dev = 0;
  VMB_CUDA(cudaGetDevice(&dev));
  if (dev < 0 || dev >= 64 || !attr_set[dev]) {
    VMB_CUDA(cudaFuncSetAttribute(gemm_tc_pair_kernel<BN, kStages>,
                                  cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    if (dev >= 0 && dev < 64) attr_set[dev] = true;
  }
  const int64_t tiles = ((M + 2 * BM - 1) / (2 * BM)) * ((N + BN - 1) / BN);
  const int pairs = (int)std::min<int64_t>(tiles, sm_count() / 2);
  // x | gate projections (act_from == N / 2, whole 128-column halves): interleaved tiles
  const int interleave = BN == 256 && N % 2 == 0 && act_from == N / 2 && act_from % (BN / 2) == 0;
  gemm_tc_pair_kernel<BN, kStages><<<2 * pairs, kThreads, smem, st>>>(ma, mw, mc, (const __nv_bfloat16*)bias,
                                                                      (int)M, N, K, act_from, interleave);
  VMB_LAUNCH_CHECK("gemm_tc_pair_kernel");
  return VMB_OK;
}

// 2-D bf16 row-major tensor, box = box_rows x 64 columns, NO swizzle (raw rows for the conv warps).
int make_map_plain(CUtensorMap* out, const void* ptr, int64_t rows, int64_t cols, int64_t ld, int box_rows) {
  EncodeFn fn = encode_fn();
  if (!fn) { set_error("gemm_tc: cuTensorMapEncodeTiled entry point not available"); return VMB_ERR_CUDA; }
  const cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
  const cuuint64_t strides[1] = {(cuuint64_t)ld * 2};
  const cuuint32_t box[2] = {(cuuint32_t)BK, (cuuint32_t)box_rows};
  const cuuint32_t estr[2] = {1, 1};
  const CUresult r = fn(out, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(ptr), dims, strides,
                        box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
                        CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) { set_error("gemm_tc: cuTensorMapEncodeTiled (plain) failed (%d)", (int)r); return VMB_ERR_CUDA; }
  return VMB_OK;
}

template <int BN, int kStages>
int launch_cx(const void* x, int64_t x_ld, const void* cw, const void* cb, const void* W, int64_t ldw,
              void* xc, int64_t xc_ld, void* C, int64_t ldc, int64_t M, int N, int K, int L, int silu,
              cudaStream_t st) {
  CUtensorMap mx, mw, mc;
  int rc;
  if ((rc = make_map_plain(&mx, x, M, K, x_ld, kCxRawRows))) return rc;
  if ((rc = make_map(&mw, W, N, K, ldw, BN))) return rc;
  if ((rc = make_map(&mc, C, M, N, ldc, BM))) return rc;
  const size_t smem = smem_bytes_cx(BN, kStages, K);
  if (smem > 227 * 1024) VMB_UNSUPPORTED("conv_xproj: K too large for the shared-memory budget");
  static bool attr_set[64][2] = {};
  int dev = 0;
  VMB_CUDA(cudaGetDevice(&dev));
  const int si = silu ? 1 : 0;
  if (dev < 0 || dev >= 64 || !attr_set[dev][si]) {
    if (silu)
      VMB_CUDA(cudaFuncSetAttribute(conv_xproj_kernel<BN, kStages, true>,
                                    cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    else
      VMB_CUDA(cudaFuncSetAttribute(conv_xproj_kernel<BN, kStages, false>,
                                    cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    if (dev >= 0 && dev < 64) attr_set[dev][si] = true;
  }
  const int64_t tiles = (M + BM - 1) / BM;
  const int grid = (int)std::min<int64_t>(tiles, sm_count());
  if (silu)
    conv_xproj_kernel<BN, kStages, true><<<grid, kCxThreads, smem, st>>>(
        mx, mw, mc, (const __nv_bfloat16*)cw, (const __nv_bfloat16*)cb, (__nv_bfloat16*)xc, xc_ld, (int)M, K, L);
  else
    conv_xproj_kernel<BN, kStages, false><<<grid, kCxThreads, smem, st>>>(
        mx, mw, mc, (const __nv_bfloat16*)cw, (const __nv_bfloat16*)cb, (__nv_bfloat16*)xc, xc_ld, (int)M, K, L);
  VMB_LAUNCH_CHECK("conv_xproj_kernel");
  return VMB_OK;
}

}  // namespace

bool gemm_tc_supported(const void* A, int64_t lda, const void* W, int64_t ldw, const void* C,
                       int64_t ldc, int64_t M, int N, int K) {
  auto al16 = [](const void* p) { return reinterpret_cast<uintptr_t>(p) % 16 == 0; };
  // TMA: 16-byte aligned bases and row pitches; keep tiny problems on the CUDA-core kernel
  return al16(A) && al16(W) && al16(C) && lda % 8 == 0 && ldw % 8 == 0 && ldc % 8 == 0 &&
         K % 8 == 0 && N % 8 == 0 && K >= 16 && N >= 16 && M >= 1 && M < (1ll << 31) - BM &&
         M * (int64_t)N >= 64 * 64;
}

int gemm_tc(const void* A, int64_t lda, const void* W, int64_t ldw, const void* bias, void* C,
            int64_t ldc, int64_t M, int N, int K, cudaStream_t st, int act_from) {
  // act_from: SiLU on the output columns [act_from, N) before the rounding; whole 64-column sub-tiles only
  if (act_from < 0 || act_from >= N) act_from = INT_MAX;
  else if (act_from % kSubN != 0) return VMB_ERR_UNSUPPORTED;
  if (M >= 4 * BM) {   // CTA pairs: 256-row tiles, half a W tile per SM
    if (N % 256 == 0) return launch_pair<256, stages_pair(256)>(A, lda, W, ldw, bias, C, ldc, M, N, K, act_from, st);
    if (N % 192 == 0) return launch_pair<192, stages_pair(192)>(A, lda, W, ldw, bias, C, ldc, M, N, K, act_from, st);
  }
  if (N <= 64) return launch<64, stages_for(64)>(A, lda, W, ldw, bias, C, ldc, M, N, K, act_from, st);
  if (N <= 128) return launch<128, stages_for(128)>(A, lda, W, ldw, bias, C, ldc, M, N, K, act_from, st);
  if (N % 256 == 0) return launch<256, stages_for(256)>(A, lda, W, ldw, bias, C, ldc, M, N, K, act_from, st);
  if (N % 192 == 0 || N < 256)
    return launch<192, stages_for(192)>(A, lda, W, ldw, bias, C, ldc, M, N, K, act_from, st);
  return launch<256, stages_for(256)>(A, lda, W, ldw, bias, C, ldc, M, N, K, act_from, st);
}

bool conv_xproj_supported(const void* x, int64_t x_ld, const void* cw, const void* cb, const void* W,
                          int64_t ldw, const void* xc, int64_t xc_ld, const void* C, int64_t ldc,
                          int64_t M, int N, int K, int L) {
  auto al16 = [](const void* p) { return reinterpret_cast<uintptr_t>(p) % 16 == 0; };
  return N == 64 && K % BK == 0 && K >= BK && smem_bytes_cx(64, 6, K) <= 227 * 1024 && M >= 4 * BM && M < (1ll << 31) - 2 * BM && L >= 4 &&
         al16(x) && al16(cw) && (cb == nullptr || al16(cb)) && al16(W) && al16(xc) && al16(C) &&
         x_ld % 8 == 0 && ldw % 8 == 0 && xc_ld % 8 == 0 && ldc % 8 == 0;
}

int conv_xproj_tc(const void* x, int64_t x_ld, const void* cw, const void* cb, const void* W,
                  int64_t ldw, void* xc, int64_t xc_ld, void* C, int64_t ldc, int64_t M, int N, int K,
                  int L, int silu, cudaStream_t st) {
  return launch_cx<64, 6>(x, x_ld, cw, cb, W, ldw, xc, xc_ld, C, ldc, M, N, K, L, silu, st);
}

}  // namespace vmb
