// Fused selective scan for the bf16 production shapes (d_state = 16):
//     dt_proj  ->  + dt_bias  ->  softplus  ->  S6 recurrence  ->  + D*u  ->  * SiLU(z)
// in ONE kernel, so delta (B, L, Di) never exists in HBM.  Per token the kernel reads u (the conv
// output), z and the x_dbl row [dt_low | B | C] and writes y: (3*Di + Xp) * 2 bytes.
//
// Stands in for the dt_proj GEMM + selective_scan_fn pair of the reference
// (models/videomamba/mamba_simple.py:413-414 and :423-435 through _selective_scan_with_state
// :109-172; semantics of _selective_scan_ref :30-106).  Difference in rounding points: the
// reference rounds delta_raw to bf16 between the two ops, here it stays fp32 (closer to exact).
//
// The kernel is bound by instruction issue and the XU (MUFU) pipe, not by HBM: 16 exponentials per
// (token, channel) in fp32 state (profiles/).  Layout:
//   * unit = (batch b, 16 channels); the sequence is walked in tiles of 16 tokens.  One lane issues
//     three cp.async.bulk.tensor box copies per tile (u 16 x 16, z 16 x 16, x_dbl Xp x 16) that
//     complete on an mbarrier; rows beyond the sequence are zero-filled by the TMA unit; the
//     reversed direction loads the box in memory order and the consumers mirror the row index
//     (compile-time kRev, so every shared-memory offset stays an immediate).  128-byte x_dbl rows
//     (dt_rank 24) use the 128-byte swizzle and are read through the same XOR.
//   * phase A (per tile): delta_raw[16 tokens x 16 channels] = dt_low[16 x R] * w_dt^T on the tensor
//     pipe (mma.sync m16n8k16, A fragments by ldmatrix straight from the staged bf16 rows, w_dt
//     fragments in registers); softplus and delta*u on the accumulator fragments -> shared memory
//     entries {delta_a, delta_b, du_a, du_b} for the channel pair (a, b) = (g, g + 8).
//   * phase B (per token): a lane owns the 2 channels x 4 states of an mma A fragment whose rows are
//     the unit's 16 channels and whose k index is the 16 states (channels g, g+8; states
//     2*tig + {0,1}, 2*tig + 8 + {0,1}).  The pair (a, b) of one state shares a packed register:
//     decay and update run on fma.rn.f32x2, the states are rounded once to bf16 pairs -- exactly the
//     fragment registers -- and ONE HMMA multiplies them with C_t (replicated over the 8 columns of
//     the B operand), so every lane of a channel receives <C_t, h_t>: no shuffle.
//   * the decay factors exp2(delta * A2[n]) come from one of three evaluators (template kExp):
//       general, MUFU:  one MUFU.EX2 per factor (8 per lane-token) -- shipped for general A;
//       general, split: kExp of the lane's four state pairs use a Cody-Waite range reduction +
//                       degree-3 polynomial on the FMA pipe (packed), the rest MUFU.  Measured (B200,
//                       profiles/r02_scan_evaluators.jsonl): one pair is a wash, two or more lose 8-40 %
//                       (10 issue slots per pair for the 2 MUFU it replaces) -- measurement builds only
//                       (-DVMB_SCAN_LAB);
//       geometric A:    A[d][n] = (n+1) * A[d][0] (exact S4D-real structure, reference
//                       mamba_simple.py:265-272): two MUFU per channel (r = e^(delta*A0) and
//                       r^(2*tig+1)), the other factors by packed multiplies: -13 %.  Selected by the
//                       caller (weights are checked when they are loaded), never guessed in the kernel.
//   * two kernels share that code: scan1w (one warp per unit: tile staging, phase A, recurrence and
//     finalisation in one instruction stream; small footprint, two launches can share an SM) and
//     scan2w (helper warp: TMA, gathers, phase A, finalisation; consumer warp: recurrence only --
//     faster when the batch alone leaves the schedulers short of warps).
//   * small batches split the sequence into segments: a first pass computes every segment's end
//     state from a zero start (recurrence only) and sum(delta); a tiny kernel chains the carries;
//     the second pass runs all segments concurrently from their true initial states.  Exact.
//   * z_gate (template kGate): the staged z tile already holds SiLU(z) -- the in_proj epilogue applied it
//     (gemm_tc.cu, vmb_linear_fwd_act) -- and the finalisation is one multiply: -0.5 MUFU and -1.1 other
//     instructions per warp-token (392 -> 376 us per launch shared, profiles/r02_gate_ab.jsonl).  The
//     training forward (kCkpt) and the first pass of the split never take it.
//   * reverse = 1 walks the sequence back to front (the flipped branch of BiMambaRefinerBlock,
//     models/refiner_backbone.py:61-68, :92-135, without flip copies); with frame_len > 0 only the
//     frame axis is reversed (4-D input): every frame is cut into its own tiles (tile_geom).
#include <algorithm>

#include "internal.h"

namespace vmb {
namespace {

constexpr int kCh = 16;                 // channels per unit (8 channel pairs x 4 state quads)
constexpr int kTT = 16;                 // tokens per tile
constexpr int kRowBytes = 48;           // y tile rows: 32 B of channels + 16 B pad
constexpr int kN = 16;
constexpr int kExpGeo = 8;              // kExp value of the geometric-A evaluator
#ifndef VMB_SCAN_POLY_PAIRS
#define VMB_SCAN_POLY_PAIRS 0           // general-A default: state pairs per lane on the FMA pipe
#endif
constexpr int kExpDefault = VMB_SCAN_POLY_PAIRS;

// x_dbl row pitch the host packs for dt_rank R (ops.xdbl_pitch): a compile-time constant here, so
// every shared-memory offset of the kernels is an immediate
__host__ __device__ constexpr int xp_of(int R) { return (R + 2 * kN + 15) / 16 * 16; }

// softplus(x) = max(x, 0) + log1p(exp(-|x|)).  log1p through lg2(1 + e): its absolute error is
// <= 1 ulp of 1.0 (6e-8), which is what matters for a step size that enters exp(delta*A) and
// delta*u linearly; identity above 20 (the reference's threshold) falls out in fp32.
__device__ __forceinline__ float softplus_mufu(float x) {
  const float e = ex2_approx(-fabsf(x) * kLog2e);
  return fmaf(lg2_approx(1.f + e), kLn2, fmaxf(x, 0.f));
}
__device__ __forceinline__ void ldmatrix_x4(uint32_t addr, uint32_t (&r)[4]) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0, %1, %2, %3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3])
               : "r"(addr));
}
__device__ __forceinline__ void mma_bf16_16816(float (&d)[4], const uint32_t (&a)[4], uint32_t b0,
                                               uint32_t b1) {
  asm volatile(
      "mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0, %1, %2, %3}, {%4, %5, %6, %7}, "
      "{%8, %9}, {%0, %1, %2, %3};"
      : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
      : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
__device__ __forceinline__ float bf16lo(uint32_t v) { return __uint_as_float(v << 16); }
__device__ __forceinline__ float bf16hi(uint32_t v) { return __uint_as_float(v & 0xffff0000u); }
__device__ __forceinline__ uint32_t pack_bf16x2(float lo, float hi) {
  uint32_t r;
  asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi), "f"(lo));
  return r;
}

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ bool mbar_try(uint32_t bar, uint32_t parity) {
  uint32_t done;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(done)
      : "r"(bar), "r"(parity)
      : "memory");
  return done != 0;
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  while (!mbar_try(bar, parity)) {}
}
// wait with back-off: a helper warp that is ahead must not spin on the issue slots the consumer needs
__device__ __forceinline__ void mbar_wait_sleep(uint32_t bar, uint32_t parity) {
  while (!mbar_try(bar, parity)) __nanosleep(256);
}
__device__ __forceinline__ void tma_load_3d(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0,
                                            int c1, int c2) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
      ::"r"(dst), "l"(map), "r"(bar), "r"(c0), "r"(c1), "r"(c2)
      : "memory");
}

// ---- decay factors --------------------------------------------------------------------------------
__device__ __forceinline__ float2 ex2_pair(float2 x) { return make_float2(ex2_approx(x.x), ex2_approx(x.y)); }

// 2^x for x <= 0 on the FMA / ALU pipes, two values per instruction: Cody-Waite split x = n + f with
// n = round(x) (magic-number add), 2^f on [-0.5, 0.5] by the degree-3 minimax polynomial constrained to
// p(0) = 1 (relative error 1.0e-4; the decay RATE log2(p) is within 4.5e-4 relative for every f, so
// slowly decaying states keep their time constant), exponent inserted by an integer add.  The clamp
// keeps n inside the exponent field (results below 2^-125 are indistinguishable from 0 here).
__device__ __forceinline__ float2 exp2_poly_pair(float2 x) {
  x.x = fmaxf(x.x, -125.f);
  x.y = fmaxf(x.y, -125.f);
  const float2 magic = make_float2(12582912.f, 12582912.f);          // 1.5 * 2^23
  const float2 t = __fadd2_rn(x, magic);                             // low mantissa bits = round(x)
  const float2 nf = __fadd2_rn(t, make_float2(-12582912.f, -12582912.f));
  const float2 f = __fadd2_rn(x, make_float2(-nf.x, -nf.y));
  float2 p = make_float2(0.05500893294811249f, 0.05500893294811249f);
  p = __ffma2_rn(p, f, make_float2(0.2422109693288803f, 0.2422109693288803f));
  p = __ffma2_rn(p, f, make_float2(0.6932829022407532f, 0.6932829022407532f));
  p = __ffma2_rn(p, f, make_float2(1.f, 1.f));
  return make_float2(__int_as_float(__float_as_int(p.x) + (__float_as_int(t.x) << 23)),
                     __int_as_float(__float_as_int(p.y) + (__float_as_int(t.y) << 23)));
}

// The state a lane carries: its 2 channels (a = g, b = g + 8 of the unit) x 4 states, packed as
// (a, b) pairs per state; A holds the matching A2 = A * log2(e) entries.
template <int kExp, bool kStateOnly>
struct Lane {
  float2 A[4];      // general: A2 of states {2tig, 2tig+1, 2tig+8, 2tig+9}; geometric: [0] = state 2tig, [1] = state 0
  float2 h[4];
  float sum_a, sum_b;

  // seg > 0 (second pass of the sequence split): the state at the start of segment `seg` is chained here from the
  // records of the first pass, h <- exp2(A2 * S[s]) * h + H[s] for s < seg (H[s] = end state of segment s from a zero
  // start, S[s] = its sum of delta): a few dozen instructions per preceding segment instead of a kernel between
  // the passes.
  __device__ __forceinline__ void init(const FastScanArgs& a, int ch_a, int tig, int64_t hoff_a, int seg, int b) {
    const float* pa = a.A2 + (int64_t)ch_a * kN;
    const float* pb = pa + 8 * kN;
    if constexpr (kExp == kExpGeo) {
      A[0] = make_float2(pa[2 * tig], pb[2 * tig]);
      A[1] = make_float2(pa[0], pb[0]);
      A[2] = A[3] = make_float2(0.f, 0.f);
    } else {
      const int n[4] = {2 * tig, 2 * tig + 1, 2 * tig + 8, 2 * tig + 9};
#pragma unroll
      for (int i = 0; i < 4; ++i) A[i] = make_float2(pa[n[i]], pb[n[i]]);
    }
    const int64_t hoff_b = hoff_a + 8 * kN;
    auto ld = [&](int64_t off) -> float {
      if (kStateOnly) return 0.f;
      return a.h0 ? load_as_f32(a.h0, off, a.h0_dtype) : 0.f;
    };
    h[0] = make_float2(ld(hoff_a), ld(hoff_b));
    h[1] = make_float2(ld(hoff_a + 1), ld(hoff_b + 1));
    h[2] = make_float2(ld(hoff_a + 8), ld(hoff_b + 8));
    h[3] = make_float2(ld(hoff_a + 9), ld(hoff_b + 9));
    sum_a = sum_b = 0.f;
    if constexpr (!kStateOnly) {
      if (seg > 0) {
        const int n[4] = {2 * tig, 2 * tig + 1, 2 * tig + 8, 2 * tig + 9};
        float2 An[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) An[i] = make_float2(pa[n[i]], pb[n[i]]);
        const int64_t seg_stride = (int64_t)a.B * a.Di * kN;
        const float* H = a.seg_ws;
        const float* S = a.seg_ws + (int64_t)a.nseg * seg_stride + (int64_t)b * a.Di + ch_a;
        const int64_t s_stride = (int64_t)a.B * a.Di;
        const int off[4] = {0, 1, 8, 9};
#pragma unroll 4
        for (int sg = 0; sg < seg; ++sg) {
          const float sa = S[sg * s_stride], sb = S[sg * s_stride + 8];
          const float* Hs = H + sg * seg_stride;
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            h[i].x = fmaf(exp2f(An[i].x * sa), h[i].x, Hs[hoff_a + off[i]]);
            h[i].y = fmaf(exp2f(An[i].y * sb), h[i].y, Hs[hoff_b + off[i]]);
          }
        }
      }
    }
  }

  __device__ __forceinline__ void store(float* dst, int64_t hoff_a) const {
    const int64_t hoff_b = hoff_a + 8 * kN;
    *reinterpret_cast<float2*>(dst + hoff_a) = make_float2(h[0].x, h[1].x);
    *reinterpret_cast<float2*>(dst + hoff_a + 8) = make_float2(h[2].x, h[3].x);
    *reinterpret_cast<float2*>(dst + hoff_b) = make_float2(h[0].y, h[1].y);
    *reinterpret_cast<float2*>(dst + hoff_b + 8) = make_float2(h[2].y, h[3].y);
  }

  // record of the state for the backward (scan_bwd_fast.cu): 256 floats per (unit, 4-token group), lane l
  // owns floats [4l, 4l + 4) and [128 + 4l, 128 + 4l + 4)
  __device__ __forceinline__ void checkpoint(float4* rec) const {
    rec[0] = make_float4(h[0].x, h[0].y, h[1].x, h[1].y);
    rec[32] = make_float4(h[2].x, h[2].y, h[3].x, h[3].y);
  }

  // dd = {delta_a, delta_b, du_a, du_b}; Bv = B_t of the lane's four states; c0 / c1 = the lane's
  // bf16 pairs of C_t (mma B fragment).  d[0] / d[2] return <C_t, h_t> of channels a / b.
  __device__ __forceinline__ void token(const float4 dd, const float4 Bv, uint32_t c0, uint32_t c1,
                                        float (&d)[4]) {
    const float2 dab = make_float2(dd.x, dd.y), du = make_float2(dd.z, dd.w);
    float2 e[4];
    if constexpr (kExp == kExpGeo) {
      const float2 e1 = ex2_pair(__fmul2_rn(dab, A[0]));       // r^(2tig+1)
      const float2 r = ex2_pair(__fmul2_rn(dab, A[1]));        // r = 2^(delta * A2[0])
      const float2 r2 = __fmul2_rn(r, r), r4 = __fmul2_rn(r2, r2), r8 = __fmul2_rn(r4, r4);
      e[0] = e1;
      e[1] = __fmul2_rn(e1, r);
      e[2] = __fmul2_rn(e1, r8);
      e[3] = __fmul2_rn(e[1], r8);
    } else {
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const float2 x = __fmul2_rn(dab, A[i]);
        // the polynomial takes the high-index (usually fastest decaying) states first
        if (i >= 4 - kExp) e[i] = exp2_poly_pair(x);
        else e[i] = ex2_pair(x);
      }
    }
    h[0] = __ffma2_rn(e[0], h[0], __fmul2_rn(du, make_float2(Bv.x, Bv.x)));
    h[1] = __ffma2_rn(e[1], h[1], __fmul2_rn(du, make_float2(Bv.y, Bv.y)));
    h[2] = __ffma2_rn(e[2], h[2], __fmul2_rn(du, make_float2(Bv.z, Bv.z)));
    h[3] = __ffma2_rn(e[3], h[3], __fmul2_rn(du, make_float2(Bv.w, Bv.w)));
    if constexpr (kStateOnly) {
      sum_a += dd.x;
      sum_b += dd.y;
    } else {
      const uint32_t af[4] = {pack_bf16x2(h[0].x, h[1].x), pack_bf16x2(h[0].y, h[1].y),
                              pack_bf16x2(h[2].x, h[3].x), pack_bf16x2(h[2].y, h[3].y)};
      d[0] = d[1] = d[2] = d[3] = 0.f;
      mma_bf16_16816(d, af, c0, c1);
    }
  }
};

// Tile -> (first memory row of its 16-row box, how many of its logical rows belong to the walk).
// Forward: rows [16 tile, +16).  kRev: the box is loaded in memory order and read mirrored.  Frame-axis
// reversal (frame_len > 0, forward code path): frames of frame_len tokens back to front, every frame
// cut into its own tiles so a tile never straddles two frames (the rows a box reads beyond the end of
// its frame are treated like the rows beyond the end of a sequence).
template <bool kRev>
__device__ __forceinline__ void tile_geom(int tile, int L_total, int L_end, int frame_len, int tpf,
                                          int& row0, int& valid) {
  if (!kRev && frame_len > 0) {
    const int f = tile / tpf, j = tile - f * tpf;
    row0 = L_total - (f + 1) * frame_len + j * kTT;
    valid = min(kTT, frame_len - j * kTT);
  } else {
    const int t0 = tile * kTT;
    row0 = kRev ? L_total - kTT - t0 : t0;                    // may be < 0: the TMA unit zero-fills
    valid = min(kTT, L_end - t0);
  }
}

// phase A of one tile for one warp: dt projection on the tensor pipe, softplus, delta * u -> sdd.
// xa_addr: ldmatrix addresses (per k-step) of the x_dbl tile; su: u tile (dense 32-byte rows).
template <int KST, bool kRev>
__device__ __forceinline__ void phase_a(const uint32_t (&xa_addr)[KST], const uint32_t (&bfrag)[2][KST][2],
                                        const float (&bias)[2][2], const uint8_t* su, uint8_t* sdd,
                                        int g, int tig, int valid) {
  float acc[2][4];
#pragma unroll
  for (int n = 0; n < 2; ++n)
#pragma unroll
    for (int i = 0; i < 4; ++i) acc[n][i] = 0.f;
#pragma unroll
  for (int ks = 0; ks < KST; ++ks) {
    uint32_t af[4];
    ldmatrix_x4(xa_addr[ks], af);
    mma_bf16_16816(acc[0], af, bfrag[0][ks][0], bfrag[0][ks][1]);
    mma_bf16_16816(acc[1], af, bfrag[1][ks][0], bfrag[1][ks][1]);
  }
#pragma unroll
  for (int half = 0; half < 2; ++half) {
    const int tl = g + 8 * half;                              // token row within the tile
    const int sr = kRev ? kTT - 1 - tl : tl;
    const bool pad = tl >= valid;                             // rows beyond the walk: delta = 0 keeps the state
    const uint32_t ua = *reinterpret_cast<const uint32_t*>(su + sr * 32 + (2 * tig) * 2);
    const uint32_t ub = *reinterpret_cast<const uint32_t*>(su + sr * 32 + (8 + 2 * tig) * 2);
#pragma unroll
    for (int i = 0; i < 2; ++i) {
      float da = softplus_mufu(acc[0][2 * half + i] + bias[0][i]);
      float db = softplus_mufu(acc[1][2 * half + i] + bias[1][i]);
      if (pad) { da = 0.f; db = 0.f; }
      const float uav = i ? bf16hi(ua) : bf16lo(ua), ubv = i ? bf16hi(ub) : bf16lo(ub);
      // entry of channel pair (2tig + i, 2tig + i + 8) of token tl; the XOR keeps the 128-bit reads of
      // the recurrence (8 pairs x 16 B per token) and these writes conflict free
      *reinterpret_cast<float4*>(sdd + tl * 128 + (((2 * tig + i) ^ (tl & 1)) << 4)) =
          make_float4(da, db, da * uav, db * ubv);
    }
  }
}

template <int R>
__device__ __forceinline__ void load_wdt(const FastScanArgs& a, int cw, int g, int tig,
                                         uint32_t (&bfrag)[2][(R + 15) / 16][2], float (&bias)[2][2]) {
  constexpr int KST = (R + 15) / 16;
  const __nv_bfloat16* wd = reinterpret_cast<const __nv_bfloat16*>(a.w_dt_pad);
#pragma unroll
  for (int n = 0; n < 2; ++n) {
    const __nv_bfloat16* wr = wd + (int64_t)(cw + 8 * n + g) * a.Rp;
#pragma unroll
    for (int ks = 0; ks < KST; ++ks) {
      bfrag[n][ks][0] = *reinterpret_cast<const uint32_t*>(wr + 16 * ks + 2 * tig);
      bfrag[n][ks][1] = *reinterpret_cast<const uint32_t*>(wr + 16 * ks + 8 + 2 * tig);
    }
    bias[n][0] = a.dt_bias ? a.dt_bias[cw + 8 * n + 2 * tig] : 0.f;
    bias[n][1] = a.dt_bias ? a.dt_bias[cw + 8 * n + 2 * tig + 1] : 0.f;
  }
}

// =================================================================================================
// scan1w: one warp per unit.
// =================================================================================================
namespace one_warp {

struct Plan {
  int x0, u0, z0, b, c, dd, y, bar, total, xb;
  __host__ __device__ constexpr int u(int s) const { return u0 + s * (kTT * 32); }
  __host__ __device__ constexpr int z(int s) const { return z0 + s * (kTT * 32); }
};
__host__ __device__ constexpr Plan plan(int Xp) {
  Plan p{};
  int off = 0;
  p.xb = Xp * 2;                                  // dense x_dbl rows
  p.x0 = off; off += kTT * p.xb;                  // 1024-byte aligned (swizzle atom) -- base of the carve-up
  p.u0 = off; off += 2 * kTT * 32;
  p.z0 = off; off += 2 * kTT * 32;
  p.b = off; off += kTT * kN * 4;                 // [token][tig][B(2tig), B(2tig+1), B(2tig+8), B(2tig+9)] fp32
  p.c = off; off += kTT * kN * 2;                 // [token][C0..15] bf16, as in the row
  p.dd = off; off += kTT * 8 * 16;                // [token][pair ^ (token & 1)]{delta_a, delta_b, du_a, du_b}
  p.y = off; off += kTT * kRowBytes;
  p.bar = off; off += 16;                         // two mbarriers (one per stage)
  p.total = off + 1024;                           // + alignment slack
  return p;
}

template <int R, bool kStateOnly, bool kRev, int kExp, bool kCkpt = false, bool kGate = false>
__global__ void __launch_bounds__(32, 18)
scan1w_kernel(const FastScanArgs a, const __grid_constant__ CUtensorMap map_u,
              const __grid_constant__ CUtensorMap map_z, const __grid_constant__ CUtensorMap map_x) {
  extern __shared__ uint8_t smem_raw[];
  constexpr int KST = (R + 15) / 16;
  constexpr Plan sp = plan(xp_of(R));
  constexpr int XB = sp.xb;
  constexpr bool kSwz = XB == 128;                // 128-byte rows: TMA 128B swizzle
  const uint32_t sbase = (static_cast<uint32_t>(__cvta_generic_to_shared(smem_raw)) + 1023u) & ~1023u;
  uint8_t* const smem = smem_raw + (sbase - static_cast<uint32_t>(__cvta_generic_to_shared(smem_raw)));
  using bf16 = __nv_bfloat16;

  const int lane = threadIdx.x;
  const int g = lane >> 2, tig = lane & 3;
  const int cw = blockIdx.x * kCh;
  const int b = blockIdx.y;
  const int seg = blockIdx.z;
  const int tbeg = seg * a.seg_len;
  const int L = min(a.L, tbeg + a.seg_len);
  const int64_t seg_stride = (int64_t)a.B * a.Di * kN;
  float* const wsH = a.seg_ws;
  float* const wsS = a.seg_ws + (int64_t)a.nseg * seg_stride;

  Lane<kExp, kStateOnly> st;
  const int64_t hoff_a = ((int64_t)b * a.Di + cw + g) * kN + 2 * tig;
  st.init(a, cw + g, tig, hoff_a, seg, b);
  const float Da = a.D ? a.D[cw + g] : 0.f, Db = a.D ? a.D[cw + g + 8] : 0.f;
  uint32_t bfrag[2][KST][2];
  float bias[2][2];
  load_wdt<R>(a, cw, g, tig, bfrag, bias);

  bf16* yg = reinterpret_cast<bf16*>(a.y) + (int64_t)b * a.y_bs + cw;
  const int y_ts = (int)a.y_ts;
  // training forward: the state before every 4-token group, for the backward pass
  float4* const ckrec = kCkpt ? reinterpret_cast<float4*>(a.ckpt) +
                                    ((int64_t)(b * gridDim.x + blockIdx.x) * ((a.L + 3) / 4)) * 64 + lane
                              : nullptr;

  // shared-memory row of logical tile row r: the reversed direction holds the box in memory order
  auto srow = [](int r) { return kRev ? kTT - 1 - r : r; };
  // byte address (relative to the tile) of byte offset o in logical row r of the x_dbl tile
  auto xoff = [&](int r, int o) {
    const int sr = srow(r);
    return kSwz ? sr * XB + ((((o >> 4) ^ (sr & 7)) << 4) | (o & 15)) : sr * XB + o;
  };

  const uint32_t bar0 = sbase + sp.bar;
  constexpr uint32_t kTileBytes = kTT * 32 * (kStateOnly ? 1 : 2) + kTT * XB;
  if (lane == 0) {
    mbar_init(bar0, 1);
    mbar_init(bar0 + 8, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncwarp();
  const int F = kRev ? 0 : a.frame_len;
  const int tpf = F > 0 ? (F + kTT - 1) / kTT : 0;
  auto issue = [&](int tile, int stg) {            // lane 0 only
    int row0, valid;
    tile_geom<kRev>(tile, a.L, L, F, tpf, row0, valid);
    const uint32_t bar = bar0 + 8 * stg;
    mbar_expect_tx(bar, kTileBytes);
    tma_load_3d(sbase + sp.x0, &map_x, bar, 0, row0, b);
    tma_load_3d(sbase + sp.u(stg), &map_u, bar, cw, row0, b);
    if (!kStateOnly) tma_load_3d(sbase + sp.z(stg), &map_z, bar, cw, row0, b);
  };

  const int tile_lo = tbeg / kTT;
  const int ntiles = F > 0 ? (a.L / F) * tpf : (L + kTT - 1) / kTT;
  if (lane == 0) issue(tile_lo, 0);
  bf16* const sy = reinterpret_cast<bf16*>(smem + sp.y);

  // loop-invariant per-lane offsets
  uint32_t xa_addr[KST];                            // ldmatrix rows of phase A
  {
    const int row = (lane & 7) + 8 * ((lane >> 3) & 1);
#pragma unroll
    for (int ks = 0; ks < KST; ++ks) xa_addr[ks] = sbase + sp.x0 + xoff(row, 32 * ks + 16 * (lane >> 4));
  }
  // B / C gathers: element e = lane + 32 i -> row (lane >> 3) + 4 i, pair p = lane & 7
  const int gp = lane & 7, gr = lane >> 3;
  // rows r and r + 8 share their swizzle phase, so two offsets (rows gr, gr + 4) cover all four
  int xb_off[2], xc_off[2];
#pragma unroll
  for (int i = 0; i < 2; ++i) {
    xb_off[i] = xoff(gr + 4 * i, (R + 2 * gp) * 2);
    xc_off[i] = xoff(gr + 4 * i, (R + kN + 2 * gp) * 2);
  }
  constexpr int kRow8 = kRev ? -8 * XB : 8 * XB;    // logical row + 8 in shared-memory bytes
  const int bslot = (gp & 3) * 4 + (gp >> 2) * 2;

  for (int tile = tile_lo; tile < ntiles; ++tile) {
    const int it = tile - tile_lo;
    const int stg = it & 1;
    int row0, valid;
    tile_geom<kRev>(tile, a.L, L, F, tpf, row0, valid);
    mbar_wait(bar0 + 8 * stg, (it >> 1) & 1);
    __syncwarp();                                  // tile landed; last tile's smem readers are done
    const uint8_t* su = smem + sp.u(stg);
    const uint8_t* sz = smem + sp.z(stg);
    const uint8_t* sx = smem + sp.x0;

    // ---- B_t of the tile to fp32, in fragment order (a lane's 4 states contiguous) --------------
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const uint32_t v = *reinterpret_cast<const uint32_t*>(sx + xb_off[i & 1] + (i >> 1) * kRow8);
      *reinterpret_cast<float2*>(smem + sp.b + (gr + 4 * i) * (kN * 4) + bslot * 4) =
          make_float2(bf16lo(v), bf16hi(v));
    }
    if constexpr (!kStateOnly) {
#pragma unroll
      for (int i = 0; i < 4; ++i)
        *reinterpret_cast<uint32_t*>(smem + sp.c + (gr + 4 * i) * (kN * 2) + gp * 4) =
            *reinterpret_cast<const uint32_t*>(sx + xc_off[i & 1] + (i >> 1) * kRow8);
    }
    phase_a<KST, kRev>(xa_addr, bfrag, bias, su, smem + sp.dd, g, tig, valid);
    __syncwarp();                                  // B / C / dd tiles visible; raw x_dbl rows no longer needed
    if (tile + 1 < ntiles && lane == 0) {          // prefetch the next tile behind the recurrence
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // our reads before the TMA's writes
      issue(tile + 1, stg ^ 1);
    }

    // ---- phase B: the recurrence; <C, h> by one HMMA per token -------------------------------------
    const uint8_t* sdd0 = smem + sp.dd + (g << 4);
    const uint8_t* sdd1 = smem + sp.dd + ((g ^ 1) << 4);
    const uint8_t* sb = smem + sp.b + tig * 16;
    const uint8_t* sc = smem + sp.c + tig * 4;
#pragma unroll
    for (int tg = 0; tg < kTT; tg += 4) {
      float ya = 0.f, yb = 0.f;
      if constexpr (kCkpt) {
        if (row0 + tg < a.L) st.checkpoint(ckrec + (int64_t)((row0 + tg) >> 2) * 64);
      }
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const int t = tg + i;
        const float4 dd = *reinterpret_cast<const float4*>(((t & 1) ? sdd1 : sdd0) + t * 128);
        const float4 Bv = *reinterpret_cast<const float4*>(sb + t * (kN * 4));
        uint32_t c0 = 0, c1 = 0;
        if constexpr (!kStateOnly) {
          c0 = *reinterpret_cast<const uint32_t*>(sc + t * (kN * 2));
          c1 = *reinterpret_cast<const uint32_t*>(sc + t * (kN * 2) + 16);
        }
        float d[4];
        st.token(dd, Bv, c0, c1, d);
        if constexpr (!kStateOnly) {
          if (tig == i) { ya = d[0]; yb = d[2]; }
        }
      }
      if constexpr (!kStateOnly) {
        // lane tig finalises token tg + tig of its two channels; its shared-memory row is
        // tg + tig (forward) or 15 - tg - tig (reversed)
        const int tf = tg + tig;
        const int so = (kRev ? (kTT - 1 - tg) * 32 : tg * 32) + (kRev ? -tig * 32 : tig * 32) + g * 2;
        const float ua = __bfloat162float(*reinterpret_cast<const bf16*>(su + so));
        const float ub = __bfloat162float(*reinterpret_cast<const bf16*>(su + so + 16));
        const float za = __bfloat162float(*reinterpret_cast<const bf16*>(sz + so));
        const float zb = __bfloat162float(*reinterpret_cast<const bf16*>(sz + so + 16));
        // z_gate: the in_proj epilogue already applied SiLU (the staged tile holds the gate itself)
        const float ga = kGate ? za : silu_fast(za);
        const float gb = kGate ? zb : silu_fast(zb);
        sy[tf * (kRowBytes / 2) + g] = __float2bfloat16_rn(fmaf(Da, ua, ya) * ga);
        sy[tf * (kRowBytes / 2) + g + 8] = __float2bfloat16_rn(fmaf(Db, ub, yb) * gb);
      }
    }
    if constexpr (!kStateOnly) {
      __syncwarp();
      const int row = lane >> 1, ch = lane & 1;
      if (row < valid)
        *reinterpret_cast<uint4*>(yg + ((row0 + (kRev ? kTT - 1 - row : row)) * y_ts + ch * 8)) =
            *reinterpret_cast<const uint4*>(smem + sp.y + row * kRowBytes + ch * 16);
    }
  }

  if constexpr (kStateOnly) {
    st.store(wsH + seg * seg_stride, hoff_a);
    if (tig == 0) {
      float* ss = wsS + ((int64_t)seg * a.B + b) * a.Di + cw + g;
      ss[0] = st.sum_a;
      ss[8] = st.sum_b;
    }
    return;
  }
  if (seg != a.nseg - 1) return;
  if (a.h_last != nullptr) st.store(a.h_last, hoff_a);
}

}  // namespace one_warp

// =================================================================================================
// scan2w: TWO warps per unit.  In scan1w one warp runs, per 16-token tile, the B / C gathers, phase A,
// the recurrence and the finalisation one after the other; the recurrence is latency bound (dependent
// MUFU / FFMA2 chains) and everything else sits in the same in-order instruction stream.  Here a
// HELPER warp does everything that does not depend on the state -- TMA issue, gathers, phase A for
// tile i, then + D*u, * SiLU(z) and the store of tile i - 1 -- while the CONSUMER warp only runs the
// recurrence and the <C, h> contraction, handing the raw sums back through shared memory.  dd / B /
// C / y_raw tiles are double buffered, u / z triple buffered (read by the helper one tile later), two
// mbarrier pairs carry the hand-over.
// =================================================================================================
namespace two_warp {

struct Plan {
  int x0, u0, z0, b, c, dd, yr, bar, total, xb;
  __host__ __device__ constexpr int x(int s) const { return x0 + s * ((kTT * xb + 1023) / 1024 * 1024); }
  __host__ __device__ constexpr int u(int s) const { return u0 + s * (kTT * 32); }
  __host__ __device__ constexpr int z(int s) const { return z0 + s * (kTT * 32); }
  __host__ __device__ constexpr int B(int s) const { return b + s * (kTT * kN * 4); }
  __host__ __device__ constexpr int C(int s) const { return c + s * (kTT * kN * 2); }
  __host__ __device__ constexpr int D(int s) const { return dd + s * (kTT * 8 * 16); }
  __host__ __device__ constexpr int Y(int s) const { return yr + s * (kTT * kCh * 4); }
};
__host__ __device__ constexpr Plan plan(int Xp) {
  Plan p{};
  int off = 0;
  p.xb = Xp * 2;
  const int xt = (kTT * p.xb + 1023) / 1024 * 1024;   // keep both x tiles on swizzle-atom boundaries
  p.x0 = off; off += 2 * xt;
  p.u0 = off; off += 3 * kTT * 32;
  p.z0 = off; off += 3 * kTT * 32;
  p.b = off; off += 2 * kTT * kN * 4;
  p.c = off; off += 2 * kTT * kN * 2;
  p.dd = off; off += 2 * kTT * 8 * 16;
  p.yr = off; off += 2 * kTT * kCh * 4;               // raw <C, h> sums, fp32 [token][channel]
  p.bar = off; off += 64;                             // tma[3], full[2], done[2], warp slot ids
  p.total = off + 1024;
  return p;
}

template <int R, bool kStateOnly, bool kRev, int kExp, bool kCkpt = false, bool kGate = false>
__global__ void __launch_bounds__(64, 11)
scan2w_kernel(const FastScanArgs a, const __grid_constant__ CUtensorMap map_u,
              const __grid_constant__ CUtensorMap map_z, const __grid_constant__ CUtensorMap map_x) {
  extern __shared__ uint8_t smem_raw[];
  constexpr int KST = (R + 15) / 16;
  constexpr int XB = xp_of(R) * 2;
  constexpr Plan sp = plan(xp_of(R));
  constexpr bool kSwz = XB == 128;
  const uint32_t sbase = (static_cast<uint32_t>(__cvta_generic_to_shared(smem_raw)) + 1023u) & ~1023u;
  uint8_t* const smem = smem_raw + (sbase - static_cast<uint32_t>(__cvta_generic_to_shared(smem_raw)));
  using bf16 = __nv_bfloat16;

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int g = lane >> 2, tig = lane & 3;
  const int cw = blockIdx.x * kCh;
  const int b = blockIdx.y;
  const int seg = blockIdx.z;                      // sequence split: segment [tbeg, L) of the sequence
  const int tbeg = seg * a.seg_len;
  const int L = min(a.L, tbeg + a.seg_len);
  const int F = kRev ? 0 : a.frame_len;
  const int tpf = F > 0 ? (F + kTT - 1) / kTT : 0;
  const int tile_lo = tbeg / kTT;
  const int ntiles = F > 0 ? (a.L / F) * tpf : (L + kTT - 1) / kTT;   // global tile indices [tile_lo, ntiles)
  const int64_t seg_stride = (int64_t)a.B * a.Di * kN;
  float* const wsH = a.seg_ws;
  float* const wsS = a.seg_ws + (int64_t)a.nseg * seg_stride;

  const uint32_t bar0 = sbase + sp.bar;
  auto tma_bar = [&](int s) { return bar0 + 8u * s; };          // tile % 3
  auto full_bar = [&](int s) { return bar0 + 24u + 8u * s; };   // helper -> consumer, tile & 1
  auto done_bar = [&](int s) { return bar0 + 40u + 8u * s; };   // consumer -> helper, tile & 1
  if (threadIdx.x == 0) {
    for (int s = 0; s < 3; ++s) mbar_init(tma_bar(s), 1);
    for (int s = 0; s < 2; ++s) { mbar_init(full_bar(s), 1); mbar_init(done_bar(s), 1); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  // Roles by hardware warp slot: a CTA's two warps sit on neighbouring schedulers (slot % 4); if warp 1
  // were always the consumer, every MUFU-heavy warp of the SM would share two of the four schedulers.
  // Both warps derive the roles from the same two slot numbers, so they always agree.
  volatile uint32_t* slot_ids = reinterpret_cast<volatile uint32_t*>(smem + sp.bar + 56);
  {
    uint32_t wid;
    asm volatile("mov.u32 %0, %%warpid;" : "=r"(wid));
    if (lane == 0) slot_ids[warp] = wid;
  }
  __syncthreads();
  bool is_helper;
  {
    const uint32_t w0 = slot_ids[0], w1 = slot_ids[1];
    const uint32_t lo = min(w0, w1), hi = max(w0, w1);
    const uint32_t consumer = ((lo >> 2) & 1u) ? hi : lo;
    is_helper = slot_ids[warp] != consumer;        // compares the STORED numbers: the two warps always differ
  }

  auto srow = [](int r) { return kRev ? kTT - 1 - r : r; };
  auto xoff = [&](int r, int o) {
    const int sr = srow(r);
    return kSwz ? sr * XB + ((((o >> 4) ^ (sr & 7)) << 4) | (o & 15)) : sr * XB + o;
  };
  const int nit = ntiles - tile_lo;

  if (is_helper) {
    // ================================ helper ================================
    uint32_t bfrag[2][KST][2];
    float bias[2][2];
    load_wdt<R>(a, cw, g, tig, bfrag, bias);
    // finalisation: lane = (token row lane >> 1, channel half lane & 1): 8 channels of one token
    const int frow = lane >> 1, fch = (lane & 1) * 8;
    float Dv[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) Dv[j] = a.D ? a.D[cw + fch + j] : 0.f;
    bf16* yg = reinterpret_cast<bf16*>(a.y) + (int64_t)b * a.y_bs + cw + fch;
    const int y_ts = (int)a.y_ts;

    // `it` = tile - tile_lo indexes the buffers and barrier phases; TMA coordinates use the global tile
    constexpr uint32_t kTileBytes = kTT * 32 * (kStateOnly ? 1 : 2) + kTT * XB;
    auto issue = [&](int it) {                      // lane 0 only
      int row0, valid;
      tile_geom<kRev>(tile_lo + it, a.L, L, F, tpf, row0, valid);
      const uint32_t bar = tma_bar(it % 3);
      mbar_expect_tx(bar, kTileBytes);
      tma_load_3d(sbase + sp.x(it & 1), &map_x, bar, 0, row0, b);
      tma_load_3d(sbase + sp.u(it % 3), &map_u, bar, cw, row0, b);
      if (!kStateOnly) tma_load_3d(sbase + sp.z(it % 3), &map_z, bar, cw, row0, b);
    };
    if (lane == 0) {
      issue(0);
      if (nit > 1) issue(1);
    }
    const int gp = lane & 7, gr = lane >> 3;
    const int bslot = (gp & 3) * 4 + (gp >> 2) * 2;
    constexpr int kRow8 = kRev ? -8 * XB : 8 * XB;

    auto finalize = [&](int it) {                   // + D*u, * SiLU(z), store: the tile's raw sums are in Y(it & 1)
      const int pb = it & 1;
      mbar_wait_sleep(done_bar(pb), (uint32_t)(it >> 1) & 1u);
      if (kStateOnly) return;                       // first pass of the split: only the buffer hand-back matters
      int row0, valid;
      tile_geom<kRev>(tile_lo + it, a.L, L, F, tpf, row0, valid);
      const uint8_t* su = smem + sp.u(it % 3) + srow(frow) * 32 + fch * 2;
      const uint8_t* sz = smem + sp.z(it % 3) + srow(frow) * 32 + fch * 2;
      const float* yr = reinterpret_cast<const float*>(smem + sp.Y(pb)) + frow * kCh + fch;
      const uint4 uv = *reinterpret_cast<const uint4*>(su);
      const uint4 zv = *reinterpret_cast<const uint4*>(sz);
      const float4 y0 = *reinterpret_cast<const float4*>(yr), y1 = *reinterpret_cast<const float4*>(yr + 4);
      const uint32_t uw[4] = {uv.x, uv.y, uv.z, uv.w}, zw[4] = {zv.x, zv.y, zv.z, zv.w};
      const float yv[8] = {y0.x, y0.y, y0.z, y0.w, y1.x, y1.y, y1.z, y1.w};
      uint32_t o[4];
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const float zl = bf16lo(zw[q]), zh = bf16hi(zw[q]);
        const float lo = fmaf(Dv[2 * q], bf16lo(uw[q]), yv[2 * q]) * (kGate ? zl : silu_fast(zl));
        const float hi = fmaf(Dv[2 * q + 1], bf16hi(uw[q]), yv[2 * q + 1]) * (kGate ? zh : silu_fast(zh));
        o[q] = pack_bf16x2(lo, hi);
      }
      if (frow < valid)
        *reinterpret_cast<uint4*>(yg + (int64_t)(row0 + (kRev ? kTT - 1 - frow : frow)) * y_ts) =
            make_uint4(o[0], o[1], o[2], o[3]);
    };

    for (int it = 0; it < nit; ++it) {
      const int pb = it & 1;
      int row0, valid;
      tile_geom<kRev>(tile_lo + it, a.L, L, F, tpf, row0, valid);
      mbar_wait(tma_bar(it % 3), (uint32_t)(it / 3) & 1u);
      // buffers pb were last read by the consumer for tile it - 2, which finalize(it - 2) has waited for
      const uint8_t* sx = smem + sp.x(pb);
      const uint8_t* su = smem + sp.u(it % 3);
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const uint32_t v = *reinterpret_cast<const uint32_t*>(sx + xoff(gr + 4 * (i & 1), (R + 2 * gp) * 2) + (i >> 1) * kRow8);
        *reinterpret_cast<float2*>(smem + sp.B(pb) + (gr + 4 * i) * (kN * 4) + bslot * 4) =
            make_float2(bf16lo(v), bf16hi(v));
        if (!kStateOnly)
          *reinterpret_cast<uint32_t*>(smem + sp.C(pb) + (gr + 4 * i) * (kN * 2) + gp * 4) =
              *reinterpret_cast<const uint32_t*>(sx + xoff(gr + 4 * (i & 1), (R + kN + 2 * gp) * 2) + (i >> 1) * kRow8);
      }
      {
        uint32_t xa_addr[KST];
        const int row = (lane & 7) + 8 * ((lane >> 3) & 1);
#pragma unroll
        for (int ks = 0; ks < KST; ++ks) xa_addr[ks] = sbase + sp.x(pb) + xoff(row, 32 * ks + 16 * (lane >> 4));
        phase_a<KST, kRev>(xa_addr, bfrag, bias, su, smem + sp.D(pb), g, tig, valid);
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(full_bar(pb));    // tile staged for the consumer
      if (it > 0) finalize(it - 1);
      __syncwarp();
      // x(pb) has been read by this tile's gathers, u / z slot (it + 2) % 3 by finalize(it - 1) just now:
      // refill them for tile it + 2 (one whole iteration ahead of their use)
      if (lane == 0 && it + 2 < nit) {
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        issue(it + 2);
      }
    }
    finalize(nit - 1);
  } else {
    // ================================ consumer ================================
    Lane<kExp, kStateOnly> st;
    const int64_t hoff_a = ((int64_t)b * a.Di + cw + g) * kN + 2 * tig;
    st.init(a, cw + g, tig, hoff_a, seg, b);
    float4* const ckrec = kCkpt ? reinterpret_cast<float4*>(a.ckpt) +
                                      ((int64_t)(b * gridDim.x + blockIdx.x) * ((a.L + 3) / 4)) * 64 + lane
                                : nullptr;
    for (int it = 0; it < nit; ++it) {
      const int pb = it & 1;
      const int tok0 = (tile_lo + it) * kTT;           // kCkpt: forward walk without frames, tile = 16 consecutive tokens
      mbar_wait(full_bar(pb), (uint32_t)(it >> 1) & 1u);
      const uint8_t* sdd0 = smem + sp.D(pb) + (g << 4);
      const uint8_t* sdd1 = smem + sp.D(pb) + ((g ^ 1) << 4);
      const uint8_t* sb = smem + sp.B(pb) + tig * 16;
      const uint8_t* sc = smem + sp.C(pb) + tig * 4;
      float* yr = reinterpret_cast<float*>(smem + sp.Y(pb));
#pragma unroll
      for (int tg = 0; tg < kTT; tg += 4) {
        float ya = 0.f, yb = 0.f;
        if constexpr (kCkpt) {
          if (tok0 + tg < a.L) st.checkpoint(ckrec + (int64_t)((tok0 + tg) >> 2) * 64);
        }
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const int t = tg + i;
          const float4 dd = *reinterpret_cast<const float4*>(((t & 1) ? sdd1 : sdd0) + t * 128);
          const float4 Bv = *reinterpret_cast<const float4*>(sb + t * (kN * 4));
          uint32_t c0 = 0, c1 = 0;
          if constexpr (!kStateOnly) {
            c0 = *reinterpret_cast<const uint32_t*>(sc + t * (kN * 2));
            c1 = *reinterpret_cast<const uint32_t*>(sc + t * (kN * 2) + 16);
          }
          float d[4];
          st.token(dd, Bv, c0, c1, d);
          if constexpr (!kStateOnly) {
            if (tig == i) { ya = d[0]; yb = d[2]; }
          }
        }
        if constexpr (!kStateOnly) {
          yr[(tg + tig) * kCh + g] = ya;             // lane tig hands over token tg + tig of its two channels
          yr[(tg + tig) * kCh + g + 8] = yb;
        }
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(done_bar(pb));      // raw sums written, dd / B / C tiles pb free
    }
    if constexpr (kStateOnly) {
      st.store(wsH + seg * seg_stride, hoff_a);
      if (tig == 0) {
        float* ss = wsS + ((int64_t)seg * a.B + b) * a.Di + cw + g;
        ss[0] = st.sum_a;
        ss[8] = st.sum_b;
      }
    } else if (seg == a.nseg - 1 && a.h_last != nullptr) {
      st.store(a.h_last, hoff_a);
    }
  }
}

}  // namespace two_warp

// Segments only when the batch alone cannot fill the GPU (one warp per 16 channels of a sequence).
constexpr int kMinSegFloor = 32;      // smallest segment (sizes the workspace)
// min_seg = 0: automatic.  A lone warp needs ~230 clk per token, three per scheduler ~120 each; the split costs one
// more dependent launch (the state pass; the carries are chained by the second pass itself) and ~1.5x the work.
// Measured as 24 dependent launches in a CUDA graph (profiles/r02_scan_split_latency.jsonl, us per launch, B = 1:
// L 196 unsplit 17.5 / 32-token segments 14.6, L 392 28.3 / 16.7 (128-token segments 24.1), L 784 49.9 / 21.3;
// B = 4: L 196 17.8 / 18.6, L 392 28.6 / 24.5, 96-token segments 26.3): 32 tokens up to one unit per SM, 96 above.
void plan_segments(const FastScanArgs& a, int min_seg, int* nseg, int* seg_len) {
  const int64_t warps = (int64_t)a.B * (a.Di / kCh);
  const int64_t want = 12ll * sm_count();       // 3 warps per scheduler
  if (min_seg <= 0) min_seg = warps <= sm_count() ? 32 : 96;
  *nseg = 1;
  *seg_len = (a.L + kTT - 1) / kTT * kTT;
  if (warps >= want || a.L < 3 * min_seg) return;
  int n = (int)std::min<int64_t>((want + warps - 1) / warps, a.L / min_seg);
  if (n < 3) return;                           // two segments cost 1.26x the work for less than that in occupancy
  const int len = ((a.L + n - 1) / n + kTT - 1) / kTT * kTT;
  *nseg = (a.L + len - 1) / len;
  *seg_len = len;
}
int64_t seg_ws_bytes_for(int nseg, int B, int Di) {    // H (nseg,B,Di,N) | S (nseg,B,Di), fp32
  return ((int64_t)nseg * B * Di * kN + (int64_t)nseg * B * Di) * 4;
}

int tensor_maps(const FastScanArgs& a, CUtensorMap* mu, CUtensorMap* mz, CUtensorMap* mx) {
  const uint64_t L = (uint64_t)a.L, B = (uint64_t)a.B;
  auto bs = [&](int64_t v, int64_t ts) { return (uint64_t)((a.B > 1 ? v : ts * a.L) * 2); };
  int rc;
  if ((rc = make_tensor_map_3d_bf16(mu, a.u, (uint64_t)a.Di, L, B, (uint64_t)a.u_ts * 2, bs(a.u_bs, a.u_ts),
                                    kCh, kTT, false)))
    return rc;
  if ((rc = make_tensor_map_3d_bf16(mz, a.z, (uint64_t)a.Di, L, B, (uint64_t)a.z_ts * 2, bs(a.z_bs, a.z_ts),
                                    kCh, kTT, false)))
    return rc;
  return make_tensor_map_3d_bf16(mx, a.xdbl, (uint64_t)a.Xp, L, B, (uint64_t)a.x_ts * 2, bs(a.x_bs, a.x_ts),
                                 (uint32_t)a.Xp, kTT, a.Xp * 2 == 128);
}

template <int R, int kExp, bool kTwoWarp>
int launch(const FastScanArgs& a0, cudaStream_t st, bool split) {
  FastScanArgs a = a0;
  if (!a.reverse) a.frame_len = 0;
  // measurement aid (tune / 100: 1 .. 4 = segments of at least 32 / 64 / 96 / 128 tokens instead of the automatic choice)
  const int ms_sel = a.tune / 100;
  plan_segments(a, ms_sel >= 1 && ms_sel <= 4 ? 32 * ms_sel : 0, &a.nseg, &a.seg_len);
  if (!split || a.frame_len > 0 || (a.nseg > 1 && (a.seg_ws == nullptr ||
                                a.seg_ws_bytes < seg_ws_bytes_for(a.nseg, a.B, a.Di)))) {
    a.nseg = 1;                                // no workspace given: run unsplit (still correct)
    a.seg_len = (a.L + kTT - 1) / kTT * kTT;
  }
  CUtensorMap mu, mz, mx;
  if (int rc = tensor_maps(a, &mu, &mz, &mx)) return rc;
  constexpr int smem = kTwoWarp ? two_warp::plan(xp_of(R)).total : one_warp::plan(xp_of(R)).total;
  constexpr int threads = kTwoWarp ? 64 : 32;
  auto run = [&](dim3 grid, auto state_only) {
    constexpr bool kSO = decltype(state_only)::value;
    const bool rev = a.reverse && a.frame_len == 0;
    if constexpr (kTwoWarp) {
      if (!kSO && a.z_gate) {                       // z holds the gate itself (inference mixer; never with ckpt)
        if (rev) two_warp::scan2w_kernel<R, false, true, kExp, false, true><<<grid, threads, smem, st>>>(a, mu, mz, mx);
        else two_warp::scan2w_kernel<R, false, false, kExp, false, true><<<grid, threads, smem, st>>>(a, mu, mz, mx);
      } else if (rev) two_warp::scan2w_kernel<R, kSO, true, kExp><<<grid, threads, smem, st>>>(a, mu, mz, mx);
      else if (!kSO && a.ckpt) two_warp::scan2w_kernel<R, false, false, kExp, true><<<grid, threads, smem, st>>>(a, mu, mz, mx);
      else two_warp::scan2w_kernel<R, kSO, false, kExp><<<grid, threads, smem, st>>>(a, mu, mz, mx);
    } else {
      if (!kSO && a.z_gate) {
        if (rev) one_warp::scan1w_kernel<R, false, true, kExp, false, true><<<grid, threads, smem, st>>>(a, mu, mz, mx);
        else one_warp::scan1w_kernel<R, false, false, kExp, false, true><<<grid, threads, smem, st>>>(a, mu, mz, mx);
      } else if (rev) one_warp::scan1w_kernel<R, kSO, true, kExp><<<grid, threads, smem, st>>>(a, mu, mz, mx);
      else if (!kSO && a.ckpt) one_warp::scan1w_kernel<R, false, false, kExp, true><<<grid, threads, smem, st>>>(a, mu, mz, mx);
      else one_warp::scan1w_kernel<R, kSO, false, kExp><<<grid, threads, smem, st>>>(a, mu, mz, mx);
    }
  };
  if (a.nseg > 1) {
    run(dim3(a.Di / kCh, a.B, a.nseg - 1), std::true_type{});
    VMB_LAUNCH_CHECK("scan kernel <state pass>");
  }
  run(dim3(a.Di / kCh, a.B, a.nseg), std::false_type{});
  VMB_LAUNCH_CHECK("scan kernel");
  return VMB_OK;
}

template <int R, int kExp>
int launch_layout(const FastScanArgs& a, cudaStream_t st, bool two_warp, bool split) {
  return two_warp ? launch<R, kExp, true>(a, st, split) : launch<R, kExp, false>(a, st, split);
}

template <int R>
int launch_exp(const FastScanArgs& a, cudaStream_t st, int exp_sel, bool two_warp, bool split) {
  switch (exp_sel) {
    case kExpGeo: return launch_layout<R, kExpGeo>(a, st, two_warp, split);
    case kExpDefault: return launch_layout<R, kExpDefault>(a, st, two_warp, split);
    default: break;
  }
#ifdef VMB_SCAN_LAB
  // measurement builds only: every general-A evaluator for the Small width (tools/scan_sweep.py)
  if constexpr (R == 24) {
    switch (exp_sel) {
      case 0: return launch_layout<R, 0>(a, st, two_warp, split);
      case 1: return launch_layout<R, 1>(a, st, two_warp, split);
      case 2: return launch_layout<R, 2>(a, st, two_warp, split);
      case 3: return launch_layout<R, 3>(a, st, two_warp, split);
      case 4: return launch_layout<R, 4>(a, st, two_warp, split);
      default: break;
    }
  }
#endif
  VMB_UNSUPPORTED("scan_fast: evaluator %d is not part of this build", exp_sel);
}

}  // namespace

bool scan_fast_supported(const FastScanArgs& a) {
  auto al16 = [](const void* p) { return reinterpret_cast<uintptr_t>(p) % 16 == 0; };
  const bool r_ok = a.R == 12 || a.R == 24 || a.R == 36;
  const int kst16 = (a.R + 15) / 16 * 16;      // dt projection reads x_dbl / w_dt columns [0, kst16)
  return a.N == kN && r_ok && a.Di % kCh == 0 && a.Xp == xp_of(a.R) && a.Rp >= kst16 && a.Rp % 2 == 0 &&
         a.B >= 1 && a.B <= 65535 && a.L >= 1 &&
         a.w_dt_pad != nullptr && reinterpret_cast<uintptr_t>(a.w_dt_pad) % 4 == 0 &&
         al16(a.u) && al16(a.z) && al16(a.xdbl) && al16(a.y) && al16(a.A2) &&
         (a.h_last == nullptr || al16(a.h_last)) &&
         a.u_bs % 8 == 0 && a.u_ts % 8 == 0 && a.z_bs % 8 == 0 && a.z_ts % 8 == 0 &&
         a.x_bs % 8 == 0 && a.x_ts % 8 == 0 && a.y_bs % 8 == 0 && a.y_ts % 8 == 0;
}

int64_t scan_fast_workspace_bytes(int B, int L, int Di, int N) {
  FastScanArgs a;
  a.B = B; a.L = L; a.Di = Di; a.N = N;
  int nseg, seg_len;
  plan_segments(a, kMinSegFloor, &nseg, &seg_len);      // upper bound over every minimum the tune field can select
  if (nseg <= 1) return 0;
  return seg_ws_bytes_for(nseg, B, Di);
}

int scan_fast(const FastScanArgs& a, cudaStream_t st) {
  // Layout (tune / 10: 0 automatic, 1 one warp per unit, 2 two warps per unit, 3 two warps + sequence
  // split).  Automatic: between 2 and 9.5 units per SM the two-warp kernel wins without splitting the
  // sequence (B200, L = 3137 / 6273: -12 % at 7.8 units per SM, -26 % at 2.6, -30 % at 3.9 for the
  // Middle width); fewer units need the sequence split, more fill the schedulers anyway and the
  // one-warp kernel's smaller footprint lets two launches share the SMs (profiles/).
  const int layout = (a.tune / 10) % 10, exp_tune = a.tune % 10;
  const int64_t units = (int64_t)a.B * (a.Di / kCh);
  const bool two_warp = layout == 2 || layout == 3 || (layout == 0 && 2 * units < 19ll * sm_count());
  const bool split = layout == 0 ? (!two_warp || units < 2ll * sm_count()) : layout != 2;
  // Evaluator (tune % 10: 0 automatic, 1..5 = general A with 0..4 polynomial pairs, 9 = geometric).
  int exp_sel = a.a_geometric ? kExpGeo : kExpDefault;
  if (exp_tune == 9) exp_sel = kExpGeo;
  else if (exp_tune >= 1 && exp_tune <= 5) exp_sel = exp_tune - 1;
  switch (a.R) {
    case 12: return launch_exp<12>(a, st, exp_sel, two_warp, split);
    case 24: return launch_exp<24>(a, st, exp_sel, two_warp, split);
    case 36: return launch_exp<36>(a, st, exp_sel, two_warp, split);
    default: VMB_UNSUPPORTED("scan_fast: dt_rank %d not built", a.R);
  }
}

}  // namespace vmb
