// Backward of the selective scan for the bf16 production shapes (d_state = 16, Di % 16 == 0): the
// lane layout of the fused forward (scan_fast.cu) applied to the reverse recurrence.
//
// Math (reference _selective_scan_ref, models/videomamba/mamba_simple.py:30-106; derivation in the
// header of scan_bwd.cu): walking t = L-1 .. 0 with g = dLoss/dh_t,
//   g += dy_t C_t;  dC_t[n] = sum_d dy h_t;  dB_t[n] = sum_d g delta u;  s1[d] = sum_n g B_t[n]
//   du = delta s1 + dy D;  ddelta = sum_n g a h_{t-1} A[n] + u s1;  dA[n] += g a h_{t-1} delta;  g *= a
//
// Layout.  unit = (batch, 16 channels) = ONE WARP; a lane owns 2 channels (a = g, b = g + 8) x 4 states
// (2 tig + {0, 1, 8, 9}) -- the registers of an mma A fragment whose rows are the unit's channels and
// whose k index is the state; the channel pair (a, b) of one state shares a packed fp32 register, so
// both recurrences run on mul / fma.rn.f32x2.  EVERY reduction is an HMMA (m16n8k16, bf16 in, fp32 out):
//   over the states:   <C_t, h_t> (ypre, for dz), <B_t, g> (s1), sum_n g a h_{t-1} A2[n] (against ones)
//   over the channels: dC_t = H_t^T dy_t and dB_t = G_t^T (delta u)_t -- the fragment is transposed in
//                      registers (4 movmatrix) and multiplied with the per-channel vector.
// The vector operand of token i of a 4-token sub-chunk is masked to columns {2i, 2i+1} of the B operand
// and the four tokens accumulate into one accumulator: afterwards lane tig holds the sums of token
// t0 + tig of its two channels (its two states for dB / dC) -- no shuffle, no select -- and finalises
// that token.  The mask is an ADDRESS: the operand rows [B | C | dy | delta*u] (bf16, permuted so a lane's
// two words are one 64-bit load) are adjacent in shared memory with a zero region behind them, and one
// per-lane displacement (0 for the lanes of columns {2i, 2i+1}, a constant for the others) turns the address
// of any of them into an address of zeros, 16 banks away from the real row.
//
// h_{t-1} in reverse order: state records, 1 KB per (unit, 4 tokens) = the lane registers of the state before
// every 4-token sub-chunk.  The fused training forward writes them (scan_fast.cu, kCkpt instantiations;
// vmb_scan_bwd_args.fwd_ckpt); without them pass 1 here (scan_ckpt_fast_kernel) walks the sequence forward
// first.  Pass 2 streams the records through a ring of three 1 KB cp.async.bulk copies two sub-chunks ahead,
// recomputes the 4 states of the sub-chunk (a_t and a_t h_{t-1} stay in registers; the h_t-only products --
// <C_t, h_t>, dC_t -- run there, beside the exponentials) and runs the reverse recurrence on them.
// Tiles of 16 tokens arrive by TMA (six 16 x 16 boxes on an mbarrier, double buffered; pass 1 uses 16-byte
// cp.async); per-channel scalars (softplus and its derivative from one exponential, the gate terms) are computed
// once per tile (phase A: lane (g, tig) prepares tokens tig + 4j of channels g, g + 8); the stage's consumed
// u / delta / z arrays then stage du / ddelta_raw / dz, written back by TMA stores that clip at L.
// dB / dC go to one slab row per (unit, token) (summed over the units by scan_bwd.cu), dA / dD / d(dt_bias)
// to per-batch partials.  No atomics: deterministic.  Measurements: profiles/r02_scan_bwd_fast_ncu.txt.
#include <algorithm>

#include "internal.h"

namespace vmb {
namespace {

constexpr int kTT = 16;                   // tokens per tile
constexpr int kN = 16;
using bf16 = __nv_bfloat16;

// shared memory of a warp (bytes).  Pass 2: two TMA stages of six raw arrays (16 rows x 32 B each); once
// phase A has consumed a stage its u / delta / z arrays become the staging rows of du / ddelta / dz (TMA
// stores).  Phase A also writes the operand rows the HMMAs read, permuted so a lane's two words are one
// 64-bit load: [t][tig]{(2tig, 2tig+1), (2tig+8, 2tig+9)} for B_t, C_t (states) and dy, delta*u (channels).
// [Bp | Cp | Yh | Uh] are adjacent and a zero region follows: ONE per-lane displacement turns the address
// of any of the four into an address of zeros (operand masking), 16 banks away from the real rows.
constexpr int kRaw = kTT * 32;
constexpr int oU = 0, oDl = kRaw, oZ = 2 * kRaw, oGo = 3 * kRaw, oB = 4 * kRaw, oC = 5 * kRaw;
constexpr int kStage = 6 * kRaw;
constexpr int oBp = 2 * kStage, oCp = oBp + kRaw, oYh = oCp + kRaw, oUh = oYh + kRaw;
constexpr int oZero = oUh + kRaw;
constexpr int kZeroBytes = 4 * kRaw + 128;
constexpr int kZDisp = oZero + 64 - oBp;  // real row address -> zeros
constexpr int oDD = oZero + kZeroBytes;   // [t][pair ^ ((t & 3) << 1)] {delta_a, delta_b, du_a, du_b}
constexpr int oDY = oDD + kTT * 128;      // same indexing: {dy_a, dy_b, delta_a, delta_b}
constexpr int oFin = oDY + kTT * 128;     // [j][lane] bf16 pairs {sig_a, sig_b | dzf_a, dzf_b | u_a, u_b | -} (private to the lane)
constexpr int oRing = oFin + 4 * 32 * 16; // three 1 KB checkpoint records in flight (bulk copies)
constexpr int oBar = oRing + 3 * 1024;    // mbarriers: two stages, three ring slots
constexpr int kSmemBwd = oBar + 64;
// B / C columns that do not start on a 16-byte boundary (x_dbl behind a dt_rank of 12 or 36: TMA boxes must):
// ONE box of 40 columns from the aligned column below B_t covers [B | C]; it lands in two extra 1280-byte slots
// (80-byte rows) behind the plan above and the repack step reads it at the 4-byte granular offset
constexpr int kWideCols = 40;
constexpr int kWideBytes = kTT * kWideCols * 2;
constexpr int oWide = oBar + 128;
static_assert(oDD % 128 == 0 && oRing % 128 == 0 && oZero % 128 == 0, "smem plan");
static_assert(11 * (kSmemBwd + 1024) <= 227 * 1024, "11 one-warp CTAs per SM");
constexpr int kStage1 = 3 * kRaw;         // pass 1 stages u, delta_raw, B only
constexpr int kSmemCkpt = 2 * kStage1 + kTT * 128;

__device__ __forceinline__ float bf16lo(uint32_t v) { return __uint_as_float(v << 16); }
__device__ __forceinline__ float bf16hi(uint32_t v) { return __uint_as_float(v & 0xffff0000u); }
__device__ __forceinline__ uint32_t pack_bf16x2(float lo, float hi) {
  uint32_t r;
  asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi), "f"(lo));
  return r;
}
__device__ __forceinline__ float ldbf(const uint8_t* p) { return __bfloat162float(*reinterpret_cast<const bf16*>(p)); }
__device__ __forceinline__ void mma_acc(float (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile(
      "mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0, %1, %2, %3}, {%4, %5, %6, %7}, "
      "{%8, %9}, {%0, %1, %2, %3};"
      : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
      : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
__device__ __forceinline__ uint32_t movm(uint32_t v) {
  uint32_t r;
  asm volatile("movmatrix.sync.aligned.m8n8.trans.b16 %0, %1;" : "=r"(r) : "r"(v));
  return r;
}
// A fragment of X (16 x 16) -> A fragment of X^T
__device__ __forceinline__ void transpose_frag(const uint32_t (&a)[4], uint32_t (&t)[4]) {
  t[0] = movm(a[0]);
  t[1] = movm(a[2]);
  t[2] = movm(a[1]);
  t[3] = movm(a[3]);
}
__device__ __forceinline__ void cp_async16(uint32_t dst, const void* src, int bytes) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(src), "r"(bytes) : "memory");
}
__device__ __forceinline__ void cp_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N> __device__ __forceinline__ void cp_wait() {
  asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory");
}
__device__ __forceinline__ float2 ex2_pair(float2 x) { return make_float2(ex2_approx(x.x), ex2_approx(x.y)); }
__device__ __forceinline__ float2 bc2(float v) { return make_float2(v, v); }

// softplus and its derivative (sigmoid) from one exponential; identity / 1 above the reference's threshold
// falls out in fp32
__device__ __forceinline__ void softplus_sig(float x, bool sp, float& d, float& s) {
  if (!sp) { d = x; s = 1.f; return; }
  const float e = ex2_approx(-fabsf(x) * kLog2e);
  const float r = rcp_approx(1.f + e);
  d = fmaf(lg2_approx(1.f + e), kLn2, fmaxf(x, 0.f));
  s = x >= 0.f ? r : e * r;
}

// 16-byte cp.async of one staged array: lane -> (row = lane >> 1, half = lane & 1); rows beyond L are zero-filled
__device__ __forceinline__ void stage_rows(uint32_t dst, const void* base, int64_t bs, int64_t ts, int b, int col,
                                           int t0, int L, int lane) {
  const int row = lane >> 1, half = lane & 1;
  const int t = t0 + row;
  const bf16* src = reinterpret_cast<const bf16*>(base) + (int64_t)b * bs + (int64_t)min(t, L - 1) * ts + col + half * 8;
  cp_async16(dst + row * 32 + half * 16, src, t < L ? 16 : 0);
}

// ---- mbarrier / TMA (async proxy) ----------------------------------------------------------------------
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
__device__ __forceinline__ void tma_load_3d(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1, int c2) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
      ::"r"(dst), "l"(map), "r"(bar), "r"(c0), "r"(c1), "r"(c2)
      : "memory");
}
__device__ __forceinline__ void tma_store_3d(const CUtensorMap* map, uint32_t src, int c0, int c1, int c2) {
  asm volatile("cp.async.bulk.tensor.3d.global.shared::cta.bulk_group [%0, {%2, %3, %4}], [%1];"
               ::"l"(map), "r"(src), "r"(c0), "r"(c1), "r"(c2)
               : "memory");
}
__device__ __forceinline__ void bulk_load(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               ::"r"(dst), "l"(src), "r"(bytes), "r"(bar)
               : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_read_all() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

// the same for a column offset that is only 4-byte aligned (B_t rows of x_dbl behind a dt_rank of 12 or 36):
// 16 rows x 8 words, four 4-byte copies per lane
__device__ __forceinline__ void stage_rows_w(uint32_t dst, const void* base, int64_t bs, int64_t ts, int b, int col,
                                             int t0, int L, int lane) {
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const int w = lane + 32 * k, row = w >> 3, cw = w & 7;
    const int t = t0 + row;
    const bf16* src = reinterpret_cast<const bf16*>(base) + (int64_t)b * bs + (int64_t)min(t, L - 1) * ts + col + cw * 2;
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4, %2;" ::"r"(dst + row * 32 + cw * 4), "l"(src), "r"(t < L ? 4 : 0)
                 : "memory");
  }
}

// the lane's state registers <-> the checkpoint record of a (unit, sub-chunk): 256 floats, lane l owns
// floats [4l, 4l+4) and [128 + 4l, 128 + 4l + 4)
struct Rec { float4 lo, hi; };
__device__ __forceinline__ void rec_to(const Rec& r, float2 (&h)[4]) {
  h[0] = make_float2(r.lo.x, r.lo.y); h[1] = make_float2(r.lo.z, r.lo.w);
  h[2] = make_float2(r.hi.x, r.hi.y); h[3] = make_float2(r.hi.z, r.hi.w);
}

// Sequence split of the reverse walk (small batches: B * Di / 16 units do not fill the GPU).  g enters the
// recurrence linearly: over a segment, g_out = P * g_in + G with P = prod a_t = exp2(A2 * sum delta) and G the
// result for g_in = 0.  scan_bwd_carry_kernel computes (G, sum delta) of every segment but the first from a
// zero start (recurrence only), scan_bwd_chain_kernel chains g_in from the last segment down, and pass 2 then
// runs all segments concurrently from their true g_in (the forward states come from the records anyway).
struct SegPlan {
  int nseg, seg_tiles;      // segments of seg_tiles 16-token tiles (the last one may be shorter)
  float* G;                 // [seg][b * nunits + unit][256] fp32, record layout (lane l: floats 4l.., 128 + 4l..)
  float* S;                 // [seg][b * nunits + unit][16]  sum of delta per channel of the unit
  float* gin;               // [seg][b * nunits + unit][256] g entering segment seg from the future side
};

// ---------------------------------------------------------------------------------------------------
// pass 1: forward walk, state before every 4-token sub-chunk -> ckpt[((b * nunits + unit) * nck + k) * 256]
// ---------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(32, 16)
scan_ckpt_fast_kernel(const vmb_scan_bwd_args a, float* __restrict__ ckpt, int nck) {
  __shared__ __align__(128) uint8_t smem[kSmemCkpt];
  const uint32_t sbase = static_cast<uint32_t>(__cvta_generic_to_shared(smem));
  const int lane = threadIdx.x, g = lane >> 2, tig = lane & 3;
  const int unit = blockIdx.x, cw = unit * 16, b = blockIdx.y;
  const int L = a.L;
  const bool sp = a.softplus != 0;
  const int ns[4] = {2 * tig, 2 * tig + 1, 2 * tig + 8, 2 * tig + 9};
  float2 A2[4], h[4];
  {
    const float* pa = a.A2 + (int64_t)(cw + g) * kN;
    const float* pb = pa + 8 * kN;
    const int64_t ha = ((int64_t)b * a.Di + cw + g) * kN, hb = ha + 8 * kN;
#pragma unroll
    for (int s = 0; s < 4; ++s) {
      A2[s] = make_float2(pa[ns[s]], pb[ns[s]]);
      h[s] = a.h0 ? make_float2(load_as_f32(a.h0, ha + ns[s], a.h0_dtype), load_as_f32(a.h0, hb + ns[s], a.h0_dtype))
                  : make_float2(0.f, 0.f);
    }
  }
  const float bias_a = a.dt_bias ? a.dt_bias[cw + g] : 0.f, bias_b = a.dt_bias ? a.dt_bias[cw + g + 8] : 0.f;
  float4* pck = reinterpret_cast<float4*>(ckpt + ((int64_t)(b * gridDim.x + unit) * nck) * 256) + lane;
  const int ntiles = (L + kTT - 1) / kTT;
  auto stage = [&](int tile, int stg) {
    const uint32_t d = sbase + stg * kStage1;
    stage_rows(d, a.u, a.u_bstride, a.u_tstride, b, cw, tile * kTT, L, lane);
    stage_rows(d + kRaw, a.delta, a.d_bstride, a.d_tstride, b, cw, tile * kTT, L, lane);
    if (a.b_off & 7) stage_rows_w(d + 2 * kRaw, a.bc, a.bc_bstride, a.bc_tstride, b, a.b_off, tile * kTT, L, lane);
    else stage_rows(d + 2 * kRaw, a.bc, a.bc_bstride, a.bc_tstride, b, a.b_off, tile * kTT, L, lane);
    cp_commit();
  };
  stage(0, 0);
  uint8_t* const sdd = smem + 2 * kStage1;
  int posoff[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) posoff[i] = (g ^ (i << 1)) << 4;
  const int pos_own = (g ^ (tig << 1)) << 4;        // entry of the tokens this lane prepares (t & 3 == tig)
  for (int tile = 0; tile < ntiles; ++tile) {
    const int stg = tile & 1;
    if (tile + 1 < ntiles) { stage(tile + 1, stg ^ 1); cp_wait<1>(); } else { cp_wait<0>(); }
    __syncwarp();
    const uint8_t* raw = smem + stg * kStage1;
    const int nvalid = min(kTT, L - tile * kTT);
    // phase A: lane prepares tokens tig + 4j of channels g, g + 8
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int t = tig + 4 * j;
      const uint8_t* r = raw + t * 32 + g * 2;
      const float ua = ldbf(r), ub = ldbf(r + 16);
      float da, db, sa, sb;
      softplus_sig(ldbf(r + kRaw) + bias_a, sp, da, sa);
      softplus_sig(ldbf(r + kRaw + 16) + bias_b, sp, db, sb);
      if (t >= nvalid) { da = 0.f; db = 0.f; }
      *reinterpret_cast<float4*>(sdd + t * 128 + pos_own) = make_float4(da, db, da * ua, db * ub);
    }
    __syncwarp();
    const uint8_t* sB = raw + 2 * kRaw + 4 * tig;
    const uint8_t* sD = sdd;
    const int nsub = (nvalid + 3) >> 2;
#pragma unroll 1
    for (int c = 0; c < nsub; ++c, sB += 4 * 32, sD += 4 * 128) {
      pck[0] = make_float4(h[0].x, h[0].y, h[1].x, h[1].y);
      pck[32] = make_float4(h[2].x, h[2].y, h[3].x, h[3].y);
      pck += 64;
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const float4 dd = *reinterpret_cast<const float4*>(sD + i * 128 + posoff[i]);
        const uint32_t br0 = *reinterpret_cast<const uint32_t*>(sB + i * 32);
        const uint32_t br1 = *reinterpret_cast<const uint32_t*>(sB + i * 32 + 16);
        const float Bv[4] = {bf16lo(br0), bf16hi(br0), bf16lo(br1), bf16hi(br1)};
        const float2 dab = make_float2(dd.x, dd.y), du = make_float2(dd.z, dd.w);
#pragma unroll
        for (int s = 0; s < 4; ++s)
          h[s] = __ffma2_rn(ex2_pair(__fmul2_rn(dab, A2[s])), h[s], __fmul2_rn(du, bc2(Bv[s])));
      }
    }
    __syncwarp();                                  // sdd / the stage are rewritten by the next iteration
  }
}

// ---------------------------------------------------------------------------------------------------
// sequence split, pass A: g over a segment from a zero start, and the segment's sum of delta
// ---------------------------------------------------------------------------------------------------
constexpr int kStageC = 4 * kRaw;         // delta_raw, z, dout, C
constexpr int kSmemCarry = 2 * kStageC + kTT * 128;

__global__ void __launch_bounds__(32, 16)
scan_bwd_carry_kernel(const vmb_scan_bwd_args a, const SegPlan sp_) {
  __shared__ __align__(128) uint8_t smem[kSmemCarry];
  const uint32_t sbase = static_cast<uint32_t>(__cvta_generic_to_shared(smem));
  const int lane = threadIdx.x, g = lane >> 2, tig = lane & 3;
  const int unit = blockIdx.x, cw = unit * 16, b = blockIdx.y, seg = blockIdx.z + 1;
  const int L = a.L;
  const bool sp = a.softplus != 0, has_z = a.z != nullptr;
  const int ns[4] = {2 * tig, 2 * tig + 1, 2 * tig + 8, 2 * tig + 9};
  float2 A2[4], gg[4];
  {
    const float* pa = a.A2 + (int64_t)(cw + g) * kN;
    const float* pb = pa + 8 * kN;
#pragma unroll
    for (int s = 0; s < 4; ++s) {
      A2[s] = make_float2(pa[ns[s]], pb[ns[s]]);
      gg[s] = make_float2(0.f, 0.f);
    }
  }
  const float bias_a = a.dt_bias ? a.dt_bias[cw + g] : 0.f, bias_b = a.dt_bias ? a.dt_bias[cw + g + 8] : 0.f;
  float sum_a = 0.f, sum_b = 0.f;
  const int ntiles = (L + kTT - 1) / kTT;
  const int tile_lo = seg * sp_.seg_tiles, tile_hi = min(ntiles, tile_lo + sp_.seg_tiles);
  auto stage = [&](int tile, int stg) {
    const uint32_t d = sbase + stg * kStageC;
    stage_rows(d, a.delta, a.d_bstride, a.d_tstride, b, cw, tile * kTT, L, lane);
    if (has_z) stage_rows(d + kRaw, a.z, a.z_bstride, a.z_tstride, b, cw, tile * kTT, L, lane);
    stage_rows(d + 2 * kRaw, a.dout, a.dout_bstride, a.dout_tstride, b, cw, tile * kTT, L, lane);
    if (a.c_off & 7) stage_rows_w(d + 3 * kRaw, a.bc, a.bc_bstride, a.bc_tstride, b, a.c_off, tile * kTT, L, lane);
    else stage_rows(d + 3 * kRaw, a.bc, a.bc_bstride, a.bc_tstride, b, a.c_off, tile * kTT, L, lane);
    cp_commit();
  };
  if (tile_hi > tile_lo) stage(tile_hi - 1, (tile_hi - 1) & 1);
  uint8_t* const sdy = smem + 2 * kStageC;
  int posoff[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) posoff[i] = (g ^ (i << 1)) << 4;
  const int pos_own = (g ^ (tig << 1)) << 4;
  for (int tile = tile_hi - 1; tile >= tile_lo; --tile) {
    const int stg = tile & 1;
    if (tile > tile_lo) { stage(tile - 1, stg ^ 1); cp_wait<1>(); } else { cp_wait<0>(); }
    __syncwarp();
    const uint8_t* raw = smem + stg * kStageC;
    const int nvalid = min(kTT, L - tile * kTT);
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int t = tig + 4 * j;
      const uint8_t* r = raw + t * 32 + g * 2;
      float da, db, sa, sb;
      softplus_sig(ldbf(r) + bias_a, sp, da, sa);
      softplus_sig(ldbf(r + 16) + bias_b, sp, db, sb);
      if (t >= nvalid) { da = 0.f; db = 0.f; }
      sum_a += da;
      sum_b += db;
      float dya = ldbf(r + 2 * kRaw), dyb = ldbf(r + 2 * kRaw + 16);          // zero beyond the sequence
      if (has_z) {
        const float za = ldbf(r + kRaw), zb = ldbf(r + kRaw + 16);
        dya *= za * fmaf(tanh_approx(0.5f * za), 0.5f, 0.5f);
        dyb *= zb * fmaf(tanh_approx(0.5f * zb), 0.5f, 0.5f);
      }
      *reinterpret_cast<float4*>(sdy + t * 128 + pos_own) = make_float4(dya, dyb, da, db);
    }
    __syncwarp();
    const uint8_t* sC = raw + 3 * kRaw + 4 * tig;
#pragma unroll
    for (int t = kTT - 1; t >= 0; --t) {
      const float4 yd = *reinterpret_cast<const float4*>(sdy + t * 128 + posoff[t & 3]);
      const uint32_t cr0 = *reinterpret_cast<const uint32_t*>(sC + t * 32);
      const uint32_t cr1 = *reinterpret_cast<const uint32_t*>(sC + t * 32 + 16);
      const float Cv[4] = {bf16lo(cr0), bf16hi(cr0), bf16lo(cr1), bf16hi(cr1)};
      const float2 dy = make_float2(yd.x, yd.y), dlt = make_float2(yd.z, yd.w);
#pragma unroll
      for (int s = 0; s < 4; ++s)
        gg[s] = __fmul2_rn(__ffma2_rn(dy, bc2(Cv[s]), gg[s]), ex2_pair(__fmul2_rn(dlt, A2[s])));
    }
    __syncwarp();
  }
  const int64_t slot = (int64_t)seg * gridDim.y * gridDim.x + (int64_t)b * gridDim.x + unit;
  float4* rec = reinterpret_cast<float4*>(sp_.G + slot * 256) + lane;
  rec[0] = make_float4(gg[0].x, gg[0].y, gg[1].x, gg[1].y);
  rec[32] = make_float4(gg[2].x, gg[2].y, gg[3].x, gg[3].y);
  // sum of delta: the four tig lanes of a channel pair hold partial sums over their tokens
  sum_a += __shfl_xor_sync(0xffffffffu, sum_a, 1);  sum_b += __shfl_xor_sync(0xffffffffu, sum_b, 1);
  sum_a += __shfl_xor_sync(0xffffffffu, sum_a, 2);  sum_b += __shfl_xor_sync(0xffffffffu, sum_b, 2);
  if (tig == 0) {
    sp_.S[slot * 16 + g] = sum_a;
    sp_.S[slot * 16 + g + 8] = sum_b;
  }
}

// g_in of the last segment = dh_last (or 0); g_in(j - 1) = exp2(A2 * S_j) * g_in(j) + G_j.  One thread per
// record element: e -> lane (e & 127) >> 2, register q = (e & 3) + 4 (e >> 7): state pair q >> 1, channel q & 1.
__global__ void scan_bwd_chain_kernel(const vmb_scan_bwd_args a, const SegPlan sp_, int nunits) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int64_t per_seg = (int64_t)a.B * nunits * 256;
  if (i >= per_seg) return;
  const int e = (int)(i & 255);
  const int64_t bu = i >> 8;                        // b * nunits + unit
  const int unit = (int)(bu % nunits), b = (int)(bu / nunits);
  const int lane = (e & 127) >> 2, q = (e & 3) + 4 * (e >> 7);
  const int g = lane >> 2, tig = lane & 3, s = q >> 1;
  const int ch = unit * 16 + g + 8 * (q & 1);
  const int n = 2 * tig + (s & 1) + 8 * (s >> 1);
  const float A2 = a.A2[(int64_t)ch * kN + n];
  float gv = a.dh_last ? a.dh_last[((int64_t)b * a.Di + ch) * kN + n] : 0.f;
  for (int j = sp_.nseg - 1; j >= 1; --j) {
    sp_.gin[(int64_t)j * per_seg + i] = gv;
    gv = fmaf(exp2f(A2 * sp_.S[((int64_t)j * a.B * nunits + bu) * 16 + g + 8 * (q & 1)]), gv,
              sp_.G[(int64_t)j * per_seg + i]);
  }
  sp_.gin[i] = gv;
}

// ---------------------------------------------------------------------------------------------------
// pass 2: reverse walk (of one segment: blockIdx.z)
// ---------------------------------------------------------------------------------------------------
struct BwdMaps { CUtensorMap u, dl, z, go, bc, du, dd, dz; };   // bc: 16-column boxes, or the 40-column box of kWide

// 168 registers: the register file is split per scheduler (16 K each), so 3 one-warp CTAs per scheduler = 12
// per SM need <= 170 (1536 units at batch 32 are 10.4 per SM); 184 would leave 2 per scheduler
template <bool kWide>
__global__ void __launch_bounds__(32, 12)
scan_bwd_fast_kernel(const vmb_scan_bwd_args a, const __grid_constant__ BwdMaps maps, const float* __restrict__ ckpt,
                     int nck, float* __restrict__ bc_slabs, float* __restrict__ pA, float* __restrict__ pD,
                     float* __restrict__ pBias, const SegPlan sp_) {
  __shared__ __align__(1024) uint8_t smem[kWide ? oWide + 2 * kWideBytes : kSmemBwd];
  const uint32_t sbase = static_cast<uint32_t>(__cvta_generic_to_shared(smem));
  const int lane = threadIdx.x, g = lane >> 2, tig = lane & 3;
  const int unit = blockIdx.x, cw = unit * 16, b = blockIdx.y;
  const int L = a.L, Di = a.Di;
  const bool sp = a.softplus != 0, has_z = a.z != nullptr;
  const int ns[4] = {2 * tig, 2 * tig + 1, 2 * tig + 8, 2 * tig + 9};
  const int64_t ha = ((int64_t)b * Di + cw + g) * kN, hb = ha + 8 * kN;
  const int seg = blockIdx.z;
  float2 A2[4], gg[4], dA[4];
  {
    const float* pa = a.A2 + (int64_t)(cw + g) * kN;
    const float* pb = pa + 8 * kN;
#pragma unroll
    for (int s = 0; s < 4; ++s) {
      A2[s] = make_float2(pa[ns[s]], pb[ns[s]]);
      gg[s] = a.dh_last ? make_float2(a.dh_last[ha + ns[s]], a.dh_last[hb + ns[s]]) : make_float2(0.f, 0.f);
      dA[s] = make_float2(0.f, 0.f);
    }
    if (seg < sp_.nseg - 1) {                       // g entering this segment from the later ones
      const float4* r = reinterpret_cast<const float4*>(
                            sp_.gin + (((int64_t)seg * gridDim.y + b) * gridDim.x + unit) * 256) + lane;
      rec_to(Rec{r[0], r[32]}, gg);
    }
  }
  const float bias_a = a.dt_bias ? a.dt_bias[cw + g] : 0.f, bias_b = a.dt_bias ? a.dt_bias[cw + g + 8] : 0.f;
  const float Da = a.D ? a.D[cw + g] : 0.f, Db = a.D ? a.D[cw + g + 8] : 0.f;
  float dD_a = 0.f, dD_b = 0.f, dBias_a = 0.f, dBias_b = 0.f;
  // operand masks: the vector of token i of a sub-chunk lives in columns {2i, 2i+1} of the B operand
  uint32_t onesm[4];                                // the all-ones vector (bf16 pairs), masked the same way
  int mdisp[4];                                     // displacement of this lane's operand reads for token i
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    onesm[i] = (g >> 1) == i ? 0x3f803f80u : 0u;
    mdisp[i] = (g >> 1) == i ? 0 : kZDisp;
  }
  int posoff[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) posoff[i] = (g ^ (i << 1)) << 4;
  const int pos_own = (g ^ (tig << 1)) << 4;        // entry of the tokens this lane prepares / finalises (t & 3 == tig)

  const int ntiles = (L + kTT - 1) / kTT;
  const int tile_lo = seg * sp_.seg_tiles, tile_hi = min(ntiles, tile_lo + sp_.seg_tiles);
  const int k_lo = 4 * tile_lo, k_hi = min(nck, 4 * tile_hi);   // sub-chunks (records) [k_lo, k_hi) of the segment
  const uint32_t bar0 = sbase + oBar;               // +8 s: stage s; +16 + 8 r: ring slot r
  if (lane == 0) {
#pragma unroll
    for (int i = 0; i < 5; ++i) mbar_init(bar0 + 8 * i, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  // zero region (never written again)
  for (int i = lane; i < kZeroBytes / 16; i += 32) reinterpret_cast<uint4*>(smem + oZero)[i] = make_uint4(0u, 0u, 0u, 0u);
  __syncwarp();
  const uint32_t kTileBytes = (has_z ? 4 : 3) * kRaw + (kWide ? kWideBytes : 2 * kRaw);
  const int bc_col0 = a.b_off & ~7;                 // kWide: first column of the 40-column box
  auto issue_tile = [&](int tile, int stg) {        // lane 0
    const uint32_t d = sbase + stg * kStage, bar = bar0 + 8 * stg;
    const int t0 = tile * kTT;
    mbar_expect_tx(bar, kTileBytes);
    tma_load_3d(d + oU, &maps.u, bar, cw, t0, b);
    tma_load_3d(d + oDl, &maps.dl, bar, cw, t0, b);
    if (has_z) tma_load_3d(d + oZ, &maps.z, bar, cw, t0, b);
    tma_load_3d(d + oGo, &maps.go, bar, cw, t0, b);
    if constexpr (kWide) {
      tma_load_3d(sbase + oWide + stg * kWideBytes, &maps.bc, bar, bc_col0, t0, b);
    } else {
      tma_load_3d(d + oB, &maps.bc, bar, a.b_off, t0, b);
      tma_load_3d(d + oC, &maps.bc, bar, a.c_off, t0, b);
    }
  };
  // checkpoint records: 1 KB bulk copies into a ring of three, two sub-chunks ahead
  const float* rec_g = ckpt + ((int64_t)(b * gridDim.x + unit) * nck) * 256;
  auto issue_rec = [&](int k, int slot) {           // lane 0; record k -> ring slot
    mbar_expect_tx(bar0 + 16 + 8 * slot, 1024);
    bulk_load(sbase + oRing + slot * 1024, rec_g + (int64_t)k * 256, 1024, bar0 + 16 + 8 * slot);
  };
  if (lane == 0) {
    issue_tile(tile_hi - 1, (tile_hi - 1) & 1);
    issue_rec(k_hi - 1, 0);
    if (k_hi - 2 >= k_lo) issue_rec(k_hi - 2, 1);
  }
  int krec = k_hi - 3;                              // next record to fetch
  int slot_cur = 0, slot_nxt = 2;                   // slot of this sub-chunk's record / of the record fetched now
  uint32_t ring_par = 0;                            // bit r: parity to wait for on slot r

  // output rows of the token this lane finalises in the current sub-chunk (4 k + tig), stepped back 4 tokens
  // per sub-chunk: dB / dC slab row (states g, g + 8)
  int tok = 4 * (k_hi - 1) + tig;
  float* slab_p = bc_slabs + (((int64_t)unit * a.B + b) * L + tok) * 32 + g;

  uint8_t* const sdd = smem + oDD;
  uint8_t* const sdy = smem + oDY;
  uint4* const sfin = reinterpret_cast<uint4*>(smem + oFin) + lane;
  uint32_t par_stage = 0;                           // bit s: parity to wait for on stage s

  for (int tile = tile_hi - 1; tile >= tile_lo; --tile) {
    const int stg = tile & 1;
    mbar_wait(bar0 + 8 * stg, (par_stage >> stg) & 1u);
    par_stage ^= 1u << stg;
    uint8_t* const raw = smem + stg * kStage;
    const int nvalid = min(kTT, L - tile * kTT);

    // ---- phase A: per-(token, channel) scalars of the tile ------------------------------------------
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int t = tig + 4 * j;
      const uint8_t* r = raw + t * 32 + g * 2;
      const float ua = ldbf(r + oU), ub = ldbf(r + oU + 16);
      float da, db, sa, sb;
      softplus_sig(ldbf(r + oDl) + bias_a, sp, da, sa);
      softplus_sig(ldbf(r + oDl + 16) + bias_b, sp, db, sb);
      if (t >= nvalid) { da = 0.f; db = 0.f; }
      const float ga = ldbf(r + oGo), gb = ldbf(r + oGo + 16);   // zero beyond the sequence
      float dya = ga, dyb = gb, dzfa = 0.f, dzfb = 0.f;
      if (has_z) {
        const float za = ldbf(r + oZ), zb = ldbf(r + oZ + 16);
        const float s_a = fmaf(tanh_approx(0.5f * za), 0.5f, 0.5f), s_b = fmaf(tanh_approx(0.5f * zb), 0.5f, 0.5f);
        dya = ga * za * s_a;
        dyb = gb * zb * s_b;
        dzfa = ga * s_a * fmaf(za, 1.f - s_a, 1.f);
        dzfb = gb * s_b * fmaf(zb, 1.f - s_b, 1.f);
      }
      dD_a = fmaf(dya, ua, dD_a);
      dD_b = fmaf(dyb, ub, dD_b);
      const float dua = da * ua, dub = db * ub;
      *reinterpret_cast<float4*>(sdd + t * 128 + pos_own) = make_float4(da, db, dua, dub);
      *reinterpret_cast<float4*>(sdy + t * 128 + pos_own) = make_float4(dya, dyb, da, db);
      sfin[j * 32] = make_uint4(pack_bf16x2(sa, sb), pack_bf16x2(dzfa, dzfb), pack_bf16x2(ua, ub), 0u);
      // channel c of the unit -> word (c & 7) >> 1 of half c >> 3, element c & 1: permuted position
      // [t][tig' = (c & 7) >> 1]{half 0 word, half 1 word}
      uint8_t* w = smem + t * 32 + (g >> 1) * 8 + (g & 1) * 2;
      *reinterpret_cast<bf16*>(w + oYh) = __float2bfloat16_rn(dya);
      *reinterpret_cast<bf16*>(w + oYh + 4) = __float2bfloat16_rn(dyb);
      *reinterpret_cast<bf16*>(w + oUh) = __float2bfloat16_rn(dua);
      *reinterpret_cast<bf16*>(w + oUh + 4) = __float2bfloat16_rn(dub);
    }
    {                                               // B_t / C_t rows -> permuted rows
      const int row = lane >> 1, half = lane & 1;
      uint4 vb, vc;
      if constexpr (kWide) {
        const uint32_t* w = reinterpret_cast<const uint32_t*>(smem + oWide + stg * kWideBytes + row * (kWideCols * 2) +
                                                              (a.b_off - bc_col0) * 2 + half * 16);
        vb = make_uint4(w[0], w[1], w[2], w[3]);
        vc = make_uint4(w[8], w[9], w[10], w[11]);    // C_t = B_t + 16 columns
      } else {
        vb = *reinterpret_cast<const uint4*>(raw + oB + row * 32 + half * 16);
        vc = *reinterpret_cast<const uint4*>(raw + oC + row * 32 + half * 16);
      }
      uint32_t* pb = reinterpret_cast<uint32_t*>(smem + oBp + row * 32 + half * 4);
      uint32_t* pc = reinterpret_cast<uint32_t*>(smem + oCp + row * 32 + half * 4);
      pb[0] = vb.x; pb[2] = vb.y; pb[4] = vb.z; pb[6] = vb.w;
      pc[0] = vc.x; pc[2] = vc.y; pc[4] = vc.z; pc[6] = vc.w;
    }
    __syncwarp();
    // the stage's raw rows are consumed: its u / delta / z arrays now stage du / ddelta / dz.  The other
    // stage is free for the next tile once the stores of the previous tile have read their rows.
    if (lane == 0 && tile > tile_lo) {
      bulk_wait_read_all();
      fence_async_smem();
      issue_tile(tile - 1, stg ^ 1);
    }

    // ---- phase B: sub-chunks back to front -------------------------------------------------------------
    const int nsub = (nvalid + 3) >> 2;
    const uint8_t* sP = smem + oBp + 8 * tig + (nsub - 1) * (4 * 32);     // permuted operand rows of the sub-chunk
    const uint8_t* sD = sdd + (nsub - 1) * (4 * 128);
    const uint8_t* sY = sdy + (nsub - 1) * (4 * 128);
    bf16* so = reinterpret_cast<bf16*>(raw + (4 * (nsub - 1) + tig) * 32) + g;    // output staging row of the lane's token
#pragma unroll 1
    for (int c = nsub - 1; c >= 0; --c, sP -= 4 * 32, sD -= 4 * 128, sY -= 4 * 128, so -= 4 * 16) {
      float2 h[4], q[4][4], an[4][4];
      if (lane == 0 && krec >= k_lo) {              // the record two sub-chunks before this one (its slot was read last sub-chunk)
        fence_async_smem();
        issue_rec(krec, slot_nxt);
      }
      --krec;
      mbar_wait(bar0 + 16 + 8 * slot_cur, (ring_par >> slot_cur) & 1u);
      ring_par ^= 1u << slot_cur;
      {
        const uint8_t* rp = smem + oRing + slot_cur * 1024 + lane * 16;
        const Rec r{*reinterpret_cast<const float4*>(rp), *reinterpret_cast<const float4*>(rp + 512)};
        rec_to(r, h);
        slot_nxt = slot_cur;
        slot_cur = slot_cur == 2 ? 0 : slot_cur + 1;
      }
      float acc_y[4] = {0.f, 0.f, 0.f, 0.f}, acc_s[4] = {0.f, 0.f, 0.f, 0.f}, acc_w[4] = {0.f, 0.f, 0.f, 0.f};
      float acc_c[4] = {0.f, 0.f, 0.f, 0.f}, acc_b[4] = {0.f, 0.f, 0.f, 0.f};
      // forward recompute of the sub-chunk (keeps a_t and a_t h_{t-1}); everything that needs h_t only --
      // <C_t, h_t> and dC_t -- runs here, beside the exponentials
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const float4 dd = *reinterpret_cast<const float4*>(sD + i * 128 + posoff[i]);
        const uint2 br = *reinterpret_cast<const uint2*>(sP + i * 32);
        const float Bv[4] = {bf16lo(br.x), bf16hi(br.x), bf16lo(br.y), bf16hi(br.y)};
        const float2 dab = make_float2(dd.x, dd.y), du = make_float2(dd.z, dd.w);
#pragma unroll
        for (int s = 0; s < 4; ++s) {
          an[i][s] = ex2_pair(__fmul2_rn(dab, A2[s]));
          q[i][s] = __fmul2_rn(an[i][s], h[s]);
          h[s] = __ffma2_rn(du, bc2(Bv[s]), q[i][s]);
        }
        const uint32_t hf[4] = {pack_bf16x2(h[0].x, h[1].x), pack_bf16x2(h[0].y, h[1].y),
                                pack_bf16x2(h[2].x, h[3].x), pack_bf16x2(h[2].y, h[3].y)};
        const uint2 cm = *reinterpret_cast<const uint2*>(sP + mdisp[i] + i * 32 + (oCp - oBp));
        mma_acc(acc_y, hf, cm.x, cm.y);
        uint32_t tf[4];
        transpose_frag(hf, tf);
        const uint2 ym = *reinterpret_cast<const uint2*>(sP + mdisp[i] + i * 32 + (oYh - oBp));
        mma_acc(acc_c, tf, ym.x, ym.y);
      }
      // reverse recurrence
#pragma unroll
      for (int i = 3; i >= 0; --i) {
        const float4 yd = *reinterpret_cast<const float4*>(sY + i * 128 + posoff[i]);     // dy_a, dy_b, delta_a, delta_b
        const float2 dy = make_float2(yd.x, yd.y), dlt = make_float2(yd.z, yd.w);
        const uint2 cr = *reinterpret_cast<const uint2*>(sP + i * 32 + (oCp - oBp));
        const float Cv[4] = {bf16lo(cr.x), bf16hi(cr.x), bf16lo(cr.y), bf16hi(cr.y)};
#pragma unroll
        for (int s = 0; s < 4; ++s) gg[s] = __ffma2_rn(dy, bc2(Cv[s]), gg[s]);       // dLoss / dh_t
        // g a h_{t-1}: dA and the state part of ddelta
        {
          float2 w[4];
#pragma unroll
          for (int s = 0; s < 4; ++s) {
            const float2 gq = __fmul2_rn(gg[s], q[i][s]);
            dA[s] = __ffma2_rn(gq, dlt, dA[s]);
            w[s] = __fmul2_rn(gq, A2[s]);
          }
          const uint32_t wf[4] = {pack_bf16x2(w[0].x, w[1].x), pack_bf16x2(w[0].y, w[1].y),
                                  pack_bf16x2(w[2].x, w[3].x), pack_bf16x2(w[2].y, w[3].y)};
          mma_acc(acc_w, wf, onesm[i], onesm[i]);
        }
        // g: <B_t, g> and dB_t
        {
          const uint32_t gf[4] = {pack_bf16x2(gg[0].x, gg[1].x), pack_bf16x2(gg[0].y, gg[1].y),
                                  pack_bf16x2(gg[2].x, gg[3].x), pack_bf16x2(gg[2].y, gg[3].y)};
          const uint2 bm = *reinterpret_cast<const uint2*>(sP + mdisp[i] + i * 32);
          mma_acc(acc_s, gf, bm.x, bm.y);
          uint32_t tf[4];
          transpose_frag(gf, tf);
          const uint2 um = *reinterpret_cast<const uint2*>(sP + mdisp[i] + i * 32 + (oUh - oBp));
          mma_acc(acc_b, tf, um.x, um.y);
        }
#pragma unroll
        for (int s = 0; s < 4; ++s) gg[s] = __fmul2_rn(gg[s], an[i][s]);              // -> dLoss / dh_{t-1} (partial)
      }
      // lane tig finalises token 4c + tig of channels g, g + 8 (states g, g + 8 for dB / dC); rows beyond
      // the sequence are clipped by the TMA stores
      {
        const float4 yd = *reinterpret_cast<const float4*>(sY + tig * 128 + pos_own);
        const uint4 finp = sfin[c * 32];
        const float sg_a = bf16lo(finp.x), sg_b = bf16hi(finp.x), dzf_a = bf16lo(finp.y), dzf_b = bf16hi(finp.y);
        const float ua = bf16lo(finp.z), ub = bf16hi(finp.z);
        const float dr_a = fmaf(kLn2, acc_w[0], ua * acc_s[0]) * sg_a;
        const float dr_b = fmaf(kLn2, acc_w[2], ub * acc_s[2]) * sg_b;
        so[oU / 2] = __float2bfloat16_rn(fmaf(yd.z, acc_s[0], yd.x * Da));            // du
        so[oU / 2 + 8] = __float2bfloat16_rn(fmaf(yd.w, acc_s[2], yd.y * Db));
        so[oDl / 2] = __float2bfloat16_rn(dr_a);                                      // ddelta_raw
        so[oDl / 2 + 8] = __float2bfloat16_rn(dr_b);
        so[oZ / 2] = __float2bfloat16_rn(dzf_a * fmaf(Da, ua, acc_y[0]));             // dz
        so[oZ / 2 + 8] = __float2bfloat16_rn(dzf_b * fmaf(Db, ub, acc_y[2]));
        if (tok < L) {
          dBias_a += dr_a;
          dBias_b += dr_b;
          slab_p[0] = acc_b[0];
          slab_p[8] = acc_b[2];
          slab_p[16] = acc_c[0];
          slab_p[24] = acc_c[2];
        }
        tok -= 4;
        slab_p -= 4 * 32;
      }
    }
    __syncwarp();                                   // staging rows complete; the tile's shared memory is rewritten next
    if (lane == 0) {
      fence_async_smem();
      const uint32_t d = sbase + stg * kStage;
      tma_store_3d(&maps.du, d + oU, cw, tile * kTT, b);
      tma_store_3d(&maps.dd, d + oDl, cw, tile * kTT, b);
      if (a.dz) tma_store_3d(&maps.dz, d + oZ, cw, tile * kTT, b);
      bulk_commit();
    }
  }
  if (lane == 0) bulk_wait_read_all();

  // ---- per-unit results ----------------------------------------------------------------------------------
  const int64_t part = (int64_t)seg * a.B * Di;     // partial rows: (segment, batch)
#pragma unroll
  for (int s = 0; s < 4; ++s) {
    if (a.dh0 && seg == 0) {
      a.dh0[ha + ns[s]] = gg[s].x;
      a.dh0[hb + ns[s]] = gg[s].y;
    }
    pA[part * kN + ha + ns[s]] = dA[s].x;
    pA[part * kN + hb + ns[s]] = dA[s].y;
  }
  // dD / d(dt_bias): the four tig lanes of a channel pair hold partial sums over their tokens
  dD_a += __shfl_xor_sync(0xffffffffu, dD_a, 1);  dD_b += __shfl_xor_sync(0xffffffffu, dD_b, 1);
  dD_a += __shfl_xor_sync(0xffffffffu, dD_a, 2);  dD_b += __shfl_xor_sync(0xffffffffu, dD_b, 2);
  dBias_a += __shfl_xor_sync(0xffffffffu, dBias_a, 1);  dBias_b += __shfl_xor_sync(0xffffffffu, dBias_b, 1);
  dBias_a += __shfl_xor_sync(0xffffffffu, dBias_a, 2);  dBias_b += __shfl_xor_sync(0xffffffffu, dBias_b, 2);
  if (tig == 0) {
    const int64_t o = part + (int64_t)b * Di + cw + g;
    pD[o] = dD_a;  pD[o + 8] = dD_b;
    pBias[o] = dBias_a;  pBias[o + 8] = dBias_b;
  }
}

}  // namespace

bool scan_bwd_fast_supported(const vmb_scan_bwd_args& a) {
  auto al16 = [](const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; };
  auto s8 = [](int64_t v) { return (v & 7) == 0; };
  if (a.dtype != VMB_BF16 || a.N != kN || a.Di % 16 != 0 || a.L < 1 || a.B < 1) return false;
  if (!al16(a.u) || !al16(a.delta) || !al16(a.dout) || !al16(a.bc) || (a.z && !al16(a.z))) return false;
  if (!s8(a.u_bstride) || !s8(a.u_tstride) || !s8(a.d_bstride) || !s8(a.d_tstride) || !s8(a.dout_bstride) ||
      !s8(a.dout_tstride) || !s8(a.bc_bstride) || !s8(a.bc_tstride) || (a.b_off & 1) || (a.c_off & 1) || a.b_off < 0 ||
      a.c_off < 0)
    return false;
  // B / C columns: 16-byte aligned starts, or [B | C] adjacent behind any even offset (one 40-column TMA box;
  // 4-byte copies in pass 1)
  if (((a.b_off | a.c_off) & 7) && a.c_off != a.b_off + kN) return false;
  if (a.z && (!s8(a.z_bstride) || !s8(a.z_tstride))) return false;
  if (!al16(a.du) || !al16(a.ddelta) || (a.dz && (!al16(a.dz) || !s8(a.dz_bstride) || !s8(a.dz_tstride)))) return false;
  return true;
}

int64_t scan_bwd_fast_ckpt_bytes(int B, int L, int Di) {
  return (int64_t)B * ((Di + 15) / 16) * ((L + 3) / 4) * 256 * 4;
}

// Segments only when the batch alone leaves the GPU short of warps (one warp per 16 channels of a sequence);
// at least 8 tiles (128 tokens) each: the carry pass costs about a third of the walk it parallelises.
int scan_bwd_fast_segments(int B, int L, int Di, int* seg_tiles) {
  const int ntiles = (L + kTT - 1) / kTT;
  const int64_t units = (int64_t)B * ((Di + 15) / 16);
  int nseg = 1;
  if (units < 6ll * sm_count() && ntiles >= 24) {
    nseg = (int)std::min<int64_t>((10ll * sm_count() + units - 1) / units, ntiles / 8);
    if (nseg < 3) nseg = 1;
  }
  const int per = (ntiles + nseg - 1) / nseg;
  if (seg_tiles) *seg_tiles = per;
  return (ntiles + per - 1) / per;
}

// bytes behind the caller's slabs / partials that the split needs: G, S, gin
int64_t scan_bwd_fast_seg_bytes(int B, int L, int Di) {
  const int nseg = scan_bwd_fast_segments(B, L, Di, nullptr);
  if (nseg <= 1) return 0;
  const int64_t slots = (int64_t)nseg * B * ((Di + 15) / 16);
  return slots * (256 + 16 + 256) * 4;
}

int scan_bwd_fast(const vmb_scan_bwd_args& a, float* ckpt, float* slabs, float* pA, float* pD, float* pBias,
                  float* seg_ws, int* nparts, cudaStream_t st) {
  const int nck = (a.L + 3) / 4;
  SegPlan sp{};
  sp.nseg = seg_ws ? scan_bwd_fast_segments(a.B, a.L, a.Di, &sp.seg_tiles) : 1;
  if (sp.nseg <= 1) { sp.nseg = 1; sp.seg_tiles = (a.L + kTT - 1) / kTT; }
  const int nunits = a.Di / 16;
  {
    const int64_t slots = (int64_t)sp.nseg * a.B * nunits;
    sp.G = seg_ws;
    sp.S = seg_ws ? seg_ws + slots * 256 : nullptr;
    sp.gin = seg_ws ? seg_ws + slots * (256 + 16) : nullptr;
  }
  *nparts = sp.nseg * a.B;                          // rows of the dA / dD / d(dt_bias) partials
  dim3 grid(nunits, a.B, sp.nseg);
  {                                                 // 11 CTAs x 20 KB per SM need the largest shared-memory split
    static bool carved[64] = {false};
    int dev = 0;
    VMB_CUDA(cudaGetDevice(&dev));
    if (dev < 0 || dev >= 64 || !carved[dev]) {
      cudaFuncSetAttribute(scan_bwd_fast_kernel<false>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
      cudaFuncSetAttribute(scan_bwd_fast_kernel<true>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
      cudaFuncSetAttribute(scan_ckpt_fast_kernel, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
      if (dev >= 0 && dev < 64) carved[dev] = true;
    }
  }
  // 3-D maps (channel x token x batch), 16 x 16 boxes; a batch of one still needs a valid outer stride
  BwdMaps m;
  const uint64_t L = (uint64_t)a.L, B = (uint64_t)a.B, Di = (uint64_t)a.Di;
  const bool wide = ((a.b_off | a.c_off) & 7) != 0;
  auto map_of = [&](CUtensorMap* out, const void* p, uint64_t d0, int64_t ts, int64_t bs, uint32_t box0 = 16) {
    return make_tensor_map_3d_bf16(out, p, d0, L, B, (uint64_t)ts * 2, (uint64_t)(a.B > 1 ? bs : ts * a.L) * 2, box0, kTT,
                                   false);
  };
  int rc;
  if ((rc = map_of(&m.u, a.u, Di, a.u_tstride, a.u_bstride))) return rc;
  if ((rc = map_of(&m.dl, a.delta, Di, a.d_tstride, a.d_bstride))) return rc;
  if (a.z) { if ((rc = map_of(&m.z, a.z, Di, a.z_tstride, a.z_bstride))) return rc; } else m.z = m.u;
  if ((rc = map_of(&m.go, a.dout, Di, a.dout_tstride, a.dout_bstride))) return rc;
  if ((rc = map_of(&m.bc, a.bc, (uint64_t)std::max(a.b_off, a.c_off) + 16, a.bc_tstride, a.bc_bstride,
                   wide ? kWideCols : 16)))
    return rc;
  if ((rc = map_of(&m.du, a.du, Di, a.Di, (int64_t)a.L * a.Di))) return rc;
  if ((rc = map_of(&m.dd, a.ddelta, Di, a.Di, (int64_t)a.L * a.Di))) return rc;
  if (a.dz) {
    const bool dense = a.dz_tstride == 0;
    if ((rc = map_of(&m.dz, a.dz, Di, dense ? a.Di : a.dz_tstride, dense ? (int64_t)a.L * a.Di : a.dz_bstride))) return rc;
  } else {
    m.dz = m.du;
  }
  if (a.fwd_ckpt) {                                 // the fused forward already wrote the records
    VMB_CHECK_ARG(reinterpret_cast<uintptr_t>(a.fwd_ckpt) % 16 == 0, "selective_scan_bwd: fwd_ckpt not 16-byte aligned");
    ckpt = const_cast<float*>(a.fwd_ckpt);
  } else {
    scan_ckpt_fast_kernel<<<dim3(nunits, a.B), 32, 0, st>>>(a, ckpt, nck);
    VMB_LAUNCH_CHECK("scan_ckpt_fast_kernel");
  }
  if (sp.nseg > 1) {
    scan_bwd_carry_kernel<<<dim3(nunits, a.B, sp.nseg - 1), 32, 0, st>>>(a, sp);
    VMB_LAUNCH_CHECK("scan_bwd_carry_kernel");
    const int64_t n = (int64_t)a.B * nunits * 256;
    scan_bwd_chain_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(a, sp, nunits);
    VMB_LAUNCH_CHECK("scan_bwd_chain_kernel");
  }
  if (wide) scan_bwd_fast_kernel<true><<<grid, 32, 0, st>>>(a, m, ckpt, nck, slabs, pA, pD, pBias, sp);
  else scan_bwd_fast_kernel<false><<<grid, 32, 0, st>>>(a, m, ckpt, nck, slabs, pA, pD, pBias, sp);
  VMB_LAUNCH_CHECK("scan_bwd_fast_kernel");
  return VMB_OK;
}

}  // namespace vmb
