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
// that token (du, ddelta_raw, dz stores; dB / dC into its slab row).
//
// h_{t-1} in reverse order: pass 1 (ckpt kernel) walks forward and stores the state before every
// 4-token sub-chunk (1 KB per warp and sub-chunk, coalesced 128-bit stores); pass 2 reloads it one
// sub-chunk ahead, recomputes the 4 states (keeping h_{t-1} and a_t in registers) and runs the reverse
// recurrence on them.  Tiles of 16 tokens are staged by 16-byte cp.async, double buffered.
// Per-channel scalars (softplus, its derivative, the gate terms) are computed once per tile (phase A:
// lane (g, tig) prepares tokens tig + 4j of channels g, g + 8).  No atomics: deterministic.
#include "internal.h"

namespace vmb {
namespace {

constexpr int kTT = 16;                   // tokens per tile
constexpr int kN = 16;
using bf16 = __nv_bfloat16;

// shared memory of a warp (bytes).  Pass 2 stage: the six staged arrays, then the two bf16 vectors phase A
// derives; [B | C | Yh | Uh] are adjacent and the zero region sits behind both stages, so ONE per-lane
// displacement turns the address of any of the four into an address of zeros (operand masking).
constexpr int kRaw = kTT * 32;            // one staged array: 16 rows x 16 channels / states bf16
constexpr int oU = 0, oDl = kRaw, oZ = 2 * kRaw, oGo = 3 * kRaw, oB = 4 * kRaw, oC = 5 * kRaw;
constexpr int oYh = 6 * kRaw;             // [t][16 channels] bf16 dy        (written by phase A)
constexpr int oUh = 7 * kRaw;             // [t][16 channels] bf16 delta * u (written by phase A)
constexpr int kStage = 8 * kRaw;
constexpr int oZero = 2 * kStage;         // 4 * kRaw (+ 64: masked lanes read 16 banks away from the real rows) bytes of zeros
constexpr int kZeroBytes = 4 * kRaw + 64;
constexpr int oDD = oZero + kZeroBytes + 64;    // [t][pair ^ ((t & 3) << 1)] {delta_a, delta_b, du_a, du_b}
constexpr int oDY = oDD + kTT * 128;      // same indexing: {dy_a, dy_b, u_a, u_b}
constexpr int oFin = oDY + kTT * 128;     // [j][lane] bf16 {sig_a, sig_b, dzf_a, dzf_b} (private to the lane)
constexpr int oRing = oFin + 4 * 32 * 8;  // three 1 KB checkpoint records in flight (cp.async ring)
constexpr int kSmemBwd = oRing + 3 * 1024;
static_assert(kSmemBwd == 8192 + 2112 + 64 + 4096 + 1024 + 3072 && oDD % 128 == 0 && oRing % 128 == 0, "smem plan");
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

// the lane's state registers <-> the checkpoint record of a (unit, sub-chunk): 256 floats, lane l owns
// floats [4l, 4l+4) and [128 + 4l, 128 + 4l + 4)
struct Rec { float4 lo, hi; };
__device__ __forceinline__ void rec_to(const Rec& r, float2 (&h)[4]) {
  h[0] = make_float2(r.lo.x, r.lo.y); h[1] = make_float2(r.lo.z, r.lo.w);
  h[2] = make_float2(r.hi.x, r.hi.y); h[3] = make_float2(r.hi.z, r.hi.w);
}

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
    stage_rows(d + 2 * kRaw, a.bc, a.bc_bstride, a.bc_tstride, b, a.b_off, tile * kTT, L, lane);
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
// pass 2: reverse walk
// ---------------------------------------------------------------------------------------------------
// 168 registers: the register file is split per scheduler (16 K each), so 3 one-warp CTAs per scheduler = 12
// per SM need <= 170 (1536 units at batch 32 are 10.4 per SM); 184 would leave 2 per scheduler
__global__ void __launch_bounds__(32, 12)
scan_bwd_fast_kernel(const vmb_scan_bwd_args a, const float* __restrict__ ckpt, int nck,
                     float* __restrict__ bc_slabs, float* __restrict__ pA, float* __restrict__ pD,
                     float* __restrict__ pBias) {
  __shared__ __align__(128) uint8_t smem[kSmemBwd];
  const uint32_t sbase = static_cast<uint32_t>(__cvta_generic_to_shared(smem));
  const int lane = threadIdx.x, g = lane >> 2, tig = lane & 3;
  const int unit = blockIdx.x, cw = unit * 16, b = blockIdx.y;
  const int L = a.L, Di = a.Di;
  const bool sp = a.softplus != 0, has_z = a.z != nullptr;
  const int ns[4] = {2 * tig, 2 * tig + 1, 2 * tig + 8, 2 * tig + 9};
  const int64_t ha = ((int64_t)b * Di + cw + g) * kN, hb = ha + 8 * kN;
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
  }
  const float bias_a = a.dt_bias ? a.dt_bias[cw + g] : 0.f, bias_b = a.dt_bias ? a.dt_bias[cw + g + 8] : 0.f;
  const float Da = a.D ? a.D[cw + g] : 0.f, Db = a.D ? a.D[cw + g + 8] : 0.f;
  float dD_a = 0.f, dD_b = 0.f, dBias_a = 0.f, dBias_b = 0.f;
  // operand masks: the vector of token i of a sub-chunk lives in columns {2i, 2i+1} of the B operand
  uint32_t onesm[4];                                // the all-ones vector (bf16 pairs), masked the same way
#pragma unroll
  for (int i = 0; i < 4; ++i) onesm[i] = (g >> 1) == i ? 0x3f803f80u : 0u;
  int posoff[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) posoff[i] = (g ^ (i << 1)) << 4;
  const int pos_own = (g ^ (tig << 1)) << 4;        // entry of the tokens this lane prepares / finalises (t & 3 == tig)

  const int ntiles = (L + kTT - 1) / kTT;
  auto stage = [&](int tile, int stg) {
    const uint32_t d = sbase + stg * kStage;
    const int t0 = tile * kTT;
    stage_rows(d + oU, a.u, a.u_bstride, a.u_tstride, b, cw, t0, L, lane);
    stage_rows(d + oDl, a.delta, a.d_bstride, a.d_tstride, b, cw, t0, L, lane);
    if (has_z) stage_rows(d + oZ, a.z, a.z_bstride, a.z_tstride, b, cw, t0, L, lane);
    stage_rows(d + oGo, a.dout, a.dout_bstride, a.dout_tstride, b, cw, t0, L, lane);
    stage_rows(d + oB, a.bc, a.bc_bstride, a.bc_tstride, b, a.b_off, t0, L, lane);
    stage_rows(d + oC, a.bc, a.bc_bstride, a.bc_tstride, b, a.c_off, t0, L, lane);
    cp_commit();
  };
  stage(ntiles - 1, (ntiles - 1) & 1);

  // zero region (never written again)
  for (int i = lane; i < kZeroBytes / 16; i += 32) reinterpret_cast<uint4*>(smem + oZero)[i] = make_uint4(0u, 0u, 0u, 0u);

  // checkpoint records: 16-byte cp.async into a ring of three, two sub-chunks ahead.  cp.async groups in
  // commit order: every tile top commits one group (the next tile's rows), every sub-chunk top one (a
  // record; possibly empty), so "all but the newest N" is exact bookkeeping (see the waits).
  const float* rec_g = ckpt + ((int64_t)(b * gridDim.x + unit) * nck) * 256 + lane * 4;
  auto fetch_rec = [&](int k, uint32_t slot_off) {  // record k -> ring slot (empty group when k < 0)
    if (k >= 0) {
      const float* src = rec_g + (int64_t)k * 256;
      cp_async16(sbase + slot_off + lane * 16, src, 16);
      cp_async16(sbase + slot_off + 512 + lane * 16, src + 128, 16);
    }
    cp_commit();
  };
  uint32_t ring_cur = oRing, ring_n1 = oRing + 1024, ring_n2 = oRing + 2048;
  fetch_rec(nck - 1, ring_cur);
  fetch_rec(nck - 2, ring_n1);
  int krec = nck - 3;                               // next record to fetch

  // output rows of the token this lane finalises in the current sub-chunk (4 k + tig), stepped back 4 tokens
  // per sub-chunk; channels g / g + 8 of the unit
  int tok = 4 * (nck - 1) + tig;
  const int64_t row_last = (int64_t)b * L + tok;
  bf16* du_p = reinterpret_cast<bf16*>(a.du) + row_last * Di + cw + g;
  bf16* dd_p = reinterpret_cast<bf16*>(a.ddelta) + row_last * Di + cw + g;
  bf16* dz_p = a.dz ? reinterpret_cast<bf16*>(a.dz) + row_last * Di + cw + g : nullptr;
  float* slab_p = bc_slabs + (((int64_t)unit * a.B + b) * L + tok) * 32 + g;
  const int64_t step_o = 4 * (int64_t)Di;

  uint8_t* const sdd = smem + oDD;
  uint8_t* const sdy = smem + oDY;
  uint2* const sfin = reinterpret_cast<uint2*>(smem + oFin) + lane;

  for (int tile = ntiles - 1; tile >= 0; --tile) {
    const int stg = tile & 1;
    if (tile > 0) stage(tile - 1, stg ^ 1); else cp_commit();
    cp_wait<3>();                                   // newer: that group and at most two records
    __syncwarp();
    const uint8_t* raw = smem + stg * kStage;
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
      *reinterpret_cast<float4*>(sdy + t * 128 + pos_own) = make_float4(dya, dyb, ua, ub);
      sfin[j * 32] = make_uint2(pack_bf16x2(sa, sb), pack_bf16x2(dzfa, dzfb));
      uint8_t* w = smem + stg * kStage + t * 32 + g * 2;
      *reinterpret_cast<bf16*>(w + oYh) = __float2bfloat16_rn(dya);
      *reinterpret_cast<bf16*>(w + oYh + 16) = __float2bfloat16_rn(dyb);
      *reinterpret_cast<bf16*>(w + oUh) = __float2bfloat16_rn(dua);
      *reinterpret_cast<bf16*>(w + oUh + 16) = __float2bfloat16_rn(dub);
    }
    __syncwarp();

    // ---- phase B: sub-chunks back to front -------------------------------------------------------------
    const int nsub = (nvalid + 3) >> 2;
    const uint8_t* sB = raw + oB + 4 * tig + (nsub - 1) * (4 * 32);
    const uint8_t* sC = raw + oC + 4 * tig + (nsub - 1) * (4 * 32);
    const uint8_t* sD = sdd + (nsub - 1) * (4 * 128);
    const uint8_t* sY = sdy + (nsub - 1) * (4 * 128);
    // token i of a sub-chunk: lanes of columns {2i, 2i+1} read the real [B | C | Yh | Uh] rows, the others zeros
    const int zdisp = oZero + 64 - (stg * kStage + oB);
    const uint8_t* sM[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) sM[i] = sB + ((g >> 1) == i ? 0 : zdisp);
#pragma unroll 1
    for (int c = nsub - 1; c >= 0; --c, sB -= 4 * 32, sC -= 4 * 32, sD -= 4 * 128, sY -= 4 * 128) {
      float2 h[4], q[4][4], an[4][4];
      fetch_rec(krec--, ring_n2);                   // the record two sub-chunks before this one
      cp_wait<2>();                                 // this sub-chunk's record has landed
      {
        const Rec r{*reinterpret_cast<const float4*>(smem + ring_cur + lane * 16),
                    *reinterpret_cast<const float4*>(smem + ring_cur + 512 + lane * 16)};
        rec_to(r, h);
        const uint32_t t = ring_cur; ring_cur = ring_n1; ring_n1 = ring_n2; ring_n2 = t;
      }
      float acc_y[4] = {0.f, 0.f, 0.f, 0.f}, acc_s[4] = {0.f, 0.f, 0.f, 0.f}, acc_w[4] = {0.f, 0.f, 0.f, 0.f};
      float acc_c[4] = {0.f, 0.f, 0.f, 0.f}, acc_b[4] = {0.f, 0.f, 0.f, 0.f};
      // forward recompute of the sub-chunk (keeps a_t and a_t h_{t-1}); everything that needs h_t only --
      // <C_t, h_t> and dC_t -- runs here, beside the exponentials
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const float4 dd = *reinterpret_cast<const float4*>(sD + i * 128 + posoff[i]);
        const uint32_t br0 = *reinterpret_cast<const uint32_t*>(sB + i * 32);
        const uint32_t br1 = *reinterpret_cast<const uint32_t*>(sB + i * 32 + 16);
        const float Bv[4] = {bf16lo(br0), bf16hi(br0), bf16lo(br1), bf16hi(br1)};
        const float2 dab = make_float2(dd.x, dd.y), du = make_float2(dd.z, dd.w);
#pragma unroll
        for (int s = 0; s < 4; ++s) {
          an[i][s] = ex2_pair(__fmul2_rn(dab, A2[s]));
          q[i][s] = __fmul2_rn(an[i][s], h[s]);
          h[s] = __ffma2_rn(du, bc2(Bv[s]), q[i][s]);
        }
        const uint32_t hf[4] = {pack_bf16x2(h[0].x, h[1].x), pack_bf16x2(h[0].y, h[1].y),
                                pack_bf16x2(h[2].x, h[3].x), pack_bf16x2(h[2].y, h[3].y)};
        const uint32_t c0 = *reinterpret_cast<const uint32_t*>(sM[i] + i * 32 + (oC - oB));
        const uint32_t c1 = *reinterpret_cast<const uint32_t*>(sM[i] + i * 32 + (oC - oB) + 16);
        mma_acc(acc_y, hf, c0, c1);
        uint32_t tf[4];
        transpose_frag(hf, tf);
        const uint32_t v0 = *reinterpret_cast<const uint32_t*>(sM[i] + i * 32 + (oYh - oB));
        const uint32_t v1 = *reinterpret_cast<const uint32_t*>(sM[i] + i * 32 + (oYh - oB) + 16);
        mma_acc(acc_c, tf, v0, v1);
      }
      // reverse recurrence
#pragma unroll
      for (int i = 3; i >= 0; --i) {
        const float4 yu = *reinterpret_cast<const float4*>(sY + i * 128 + posoff[i]);
        const float2 dlt = *reinterpret_cast<const float2*>(sD + i * 128 + posoff[i]);
        const float2 dy = make_float2(yu.x, yu.y);
        const uint32_t cr0 = *reinterpret_cast<const uint32_t*>(sC + i * 32);
        const uint32_t cr1 = *reinterpret_cast<const uint32_t*>(sC + i * 32 + 16);
        const float Cv[4] = {bf16lo(cr0), bf16hi(cr0), bf16lo(cr1), bf16hi(cr1)};
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
          const uint32_t br0 = *reinterpret_cast<const uint32_t*>(sM[i] + i * 32);
          const uint32_t br1 = *reinterpret_cast<const uint32_t*>(sM[i] + i * 32 + 16);
          mma_acc(acc_s, gf, br0, br1);
          uint32_t tf[4];
          transpose_frag(gf, tf);
          const uint32_t v0 = *reinterpret_cast<const uint32_t*>(sM[i] + i * 32 + (oUh - oB));
          const uint32_t v1 = *reinterpret_cast<const uint32_t*>(sM[i] + i * 32 + (oUh - oB) + 16);
          mma_acc(acc_b, tf, v0, v1);
        }
#pragma unroll
        for (int s = 0; s < 4; ++s) gg[s] = __fmul2_rn(gg[s], an[i][s]);              // -> dLoss / dh_{t-1} (partial)
      }
      // lane tig finalises token 4c + tig of channels g, g + 8 (states g, g + 8 for dB / dC)
      {
        const float4 dd = *reinterpret_cast<const float4*>(sD + tig * 128 + pos_own);
        const float4 yu = *reinterpret_cast<const float4*>(sY + tig * 128 + pos_own);
        const uint2 finp = sfin[c * 32];
        const float4 fin = make_float4(bf16lo(finp.x), bf16hi(finp.x), bf16lo(finp.y), bf16hi(finp.y));
        if (tok < L) {
          const float du_a = fmaf(dd.x, acc_s[0], yu.x * Da), du_b = fmaf(dd.y, acc_s[2], yu.y * Db);
          const float dr_a = fmaf(kLn2, acc_w[0], yu.z * acc_s[0]) * fin.x;
          const float dr_b = fmaf(kLn2, acc_w[2], yu.w * acc_s[2]) * fin.y;
          du_p[0] = __float2bfloat16_rn(du_a);
          du_p[8] = __float2bfloat16_rn(du_b);
          dd_p[0] = __float2bfloat16_rn(dr_a);
          dd_p[8] = __float2bfloat16_rn(dr_b);
          dBias_a += dr_a;
          dBias_b += dr_b;
          if (dz_p) {
            dz_p[0] = __float2bfloat16_rn(fin.z * fmaf(Da, yu.z, acc_y[0]));
            dz_p[8] = __float2bfloat16_rn(fin.w * fmaf(Db, yu.w, acc_y[2]));
          }
          slab_p[0] = acc_b[0];
          slab_p[8] = acc_b[2];
          slab_p[16] = acc_c[0];
          slab_p[24] = acc_c[2];
        }
        tok -= 4;
        du_p -= step_o;
        dd_p -= step_o;
        if (dz_p) dz_p -= step_o;
        slab_p -= 4 * 32;
      }
#pragma unroll
      for (int i = 0; i < 4; ++i) sM[i] -= 4 * 32;
    }
    __syncwarp();                                  // the tile's shared memory is rewritten by the next iteration
  }

  // ---- per-unit results ----------------------------------------------------------------------------------
#pragma unroll
  for (int s = 0; s < 4; ++s) {
    if (a.dh0) {
      a.dh0[ha + ns[s]] = gg[s].x;
      a.dh0[hb + ns[s]] = gg[s].y;
    }
    pA[ha + ns[s]] = dA[s].x;
    pA[hb + ns[s]] = dA[s].y;
  }
  // dD / d(dt_bias): the four tig lanes of a channel pair hold partial sums over their tokens
  dD_a += __shfl_xor_sync(0xffffffffu, dD_a, 1);  dD_b += __shfl_xor_sync(0xffffffffu, dD_b, 1);
  dD_a += __shfl_xor_sync(0xffffffffu, dD_a, 2);  dD_b += __shfl_xor_sync(0xffffffffu, dD_b, 2);
  dBias_a += __shfl_xor_sync(0xffffffffu, dBias_a, 1);  dBias_b += __shfl_xor_sync(0xffffffffu, dBias_b, 1);
  dBias_a += __shfl_xor_sync(0xffffffffu, dBias_a, 2);  dBias_b += __shfl_xor_sync(0xffffffffu, dBias_b, 2);
  if (tig == 0) {
    const int64_t o = (int64_t)b * Di + cw + g;
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
      !s8(a.dout_tstride) || !s8(a.bc_bstride) || !s8(a.bc_tstride) || !s8(a.b_off) || !s8(a.c_off))
    return false;
  if (a.z && (!s8(a.z_bstride) || !s8(a.z_tstride))) return false;
  return true;
}

int64_t scan_bwd_fast_ckpt_bytes(int B, int L, int Di) {
  return (int64_t)B * ((Di + 15) / 16) * ((L + 3) / 4) * 256 * 4;
}

int scan_bwd_fast(const vmb_scan_bwd_args& a, float* ckpt, float* slabs, float* pA, float* pD, float* pBias,
                  cudaStream_t st) {
  const int nck = (a.L + 3) / 4;
  dim3 grid(a.Di / 16, a.B);
  static const bool carve = [] {                    // 11 CTAs x 18.1 KB per SM need the largest shared-memory split
    cudaFuncSetAttribute(scan_bwd_fast_kernel, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
    cudaFuncSetAttribute(scan_ckpt_fast_kernel, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
    return true;
  }();
  (void)carve;
  scan_ckpt_fast_kernel<<<grid, 32, 0, st>>>(a, ckpt, nck);
  VMB_LAUNCH_CHECK("scan_ckpt_fast_kernel");
  scan_bwd_fast_kernel<<<grid, 32, 0, st>>>(a, ckpt, nck, slabs, pA, pD, pBias);
  VMB_LAUNCH_CHECK("scan_bwd_fast_kernel");
  return VMB_OK;
}

}  // namespace vmb
