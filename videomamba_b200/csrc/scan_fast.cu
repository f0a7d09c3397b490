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
// Work decomposition (the kernel is bound by MUFU ex2 and issue slots, not by HBM):
//   * CTA = (batch b, 32 channels), 128 threads; walks the sequence in tiles of 32 tokens.
//   * a channel's 16 states are split over 4 adjacent lanes (4 states each): 4x the threads of a
//     thread-per-channel scan, no redundant exponentials;
//   * per group of 4 tokens each of the 4 lanes "owns" one token: it does the token's dt_proj dot
//     product (weights in registers), softplus, SiLU(z) and D*u once, and the group exchanges
//     delta / delta*u with width-4 shuffles; the per-token partial outputs of the 4 lanes are
//     combined with a 3-shuffle transpose-reduce that leaves each owner with its token's y.
//   * tiles of u, z and x_dbl are staged in shared memory with 16-byte cp.async copies, double
//     buffered; x_dbl is expanded to fp32 once per tile so B_t / C_t / dt_low are read back as
//     conflict-free 128-bit broadcasts; y leaves through shared memory as 16-byte stores.
//   * reverse = 1 walks the sequence back to front (tile rows are gathered in logical order), which
//     is what the flipped branch of BiMambaRefinerBlock needs without any torch.flip copy.
#include <cstdlib>

#include "internal.h"

namespace vmb {
namespace {

constexpr int kCh = 32;                 // channels per CTA
constexpr int kThreads = 128;           // 32 channels x 4 lanes
constexpr int kTT = 32;                 // tokens per tile
constexpr int kUZRowBytes = 80;         // 64 B of channels + 16 B pad (bank spread for 4-row reads)
constexpr int kN = 16;
constexpr int kYRowBytes = 80;          // 64 B of channels + 16 B pad

__host__ __device__ constexpr int dt_stride(int R) {  // floats per sDT row: odd number of 16-B groups
  return ((R / 4) % 2 == 1) ? R : R + 4;
}

struct Smem {   // byte offsets; u / z / x are double buffered: stage s lives at base + s * stride
  int u0, z0, x0, xstride, bc, dt, y, total;
  __host__ __device__ int u(int s) const { return u0 + s * (kTT * kUZRowBytes); }
  __host__ __device__ int z(int s) const { return z0 + s * (kTT * kUZRowBytes); }
  __host__ __device__ int x(int s) const { return x0 + s * xstride; }
};
__host__ __device__ inline Smem smem_plan(int R, int Xp) {
  Smem s;
  int off = 0;
  s.u0 = off; off += 2 * kTT * kUZRowBytes;
  s.z0 = off; off += 2 * kTT * kUZRowBytes;
  s.xstride = kTT * Xp * 2;
  s.x0 = off; off += 2 * s.xstride;
  s.bc = off; off += kTT * 2 * kN * 4;
  s.dt = off; off += kTT * dt_stride(R) * 4;
  s.y = off; off += kTT * kYRowBytes;
  s.total = off;
  return s;
}

__device__ __forceinline__ void cp_async16(uint32_t dst, const void* src, bool valid) {
  const int sz = valid ? 16 : 0;   // src-size 0: the 16 destination bytes are zero-filled
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(src), "r"(sz) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int kPending> __device__ __forceinline__ void cp_async_wait() {
  asm volatile("cp.async.wait_group %0;" ::"n"(kPending) : "memory");
}
__device__ __forceinline__ float tanh_approx(float x) {
  float y;
  asm("tanh.approx.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
// softplus with the reference's semantics (identity above 20 falls out of the formula in fp32)
__device__ __forceinline__ float softplus_mufu(float x) {
  const float e = ex2_approx(-fabsf(x) * kLog2e);                       // exp(-|x|) in (0, 1]
  const float big = lg2_approx(1.f + e) * kLn2;                         // log1p(e), e not tiny
  const float small = e * fmaf(e, fmaf(e, 0.33333334f, -0.5f), 1.f);    // e - e^2/2 + e^3/3
  return fmaxf(x, 0.f) + (e < 0.015625f ? small : big);
}
// z * sigmoid(z) with sigmoid(z) = 0.5 * tanh(z / 2) + 0.5 (one MUFU)
__device__ __forceinline__ float silu_tanh(float z) {
  return z * fmaf(tanh_approx(0.5f * z), 0.5f, 0.5f);
}

template <int R, bool kPacked>
__global__ void __launch_bounds__(kThreads, 6)
scan_fast_kernel(const FastScanArgs a) {
  extern __shared__ __align__(16) uint8_t smem[];
  constexpr int X = R + 2 * kN;
  constexpr int kDts = dt_stride(R);
  const Smem sp = smem_plan(R, a.Xp);
  const uint32_t sbase = static_cast<uint32_t>(__cvta_generic_to_shared(smem));

  const int tid = threadIdx.x;
  const int lane = tid & 31;
  const int j = tid & 3;                       // state quad [4j, 4j+4) and token slot in a group
  const int cl = tid >> 2;                     // channel within the CTA
  const int c0 = blockIdx.x * kCh;
  const int c = c0 + cl;
  const int b = blockIdx.y;
  const int L = a.L;
  using bf16 = __nv_bfloat16;

  // ---- per-thread constants ---------------------------------------------------------------
  float A2[4], h[4], w[R];
  {
    const float4 av = *reinterpret_cast<const float4*>(a.A2 + (int64_t)c * kN + 4 * j);
    A2[0] = av.x; A2[1] = av.y; A2[2] = av.z; A2[3] = av.w;
    const int64_t hoff = ((int64_t)b * a.Di + c) * kN + 4 * j;
    if (a.h0 != nullptr) {
#pragma unroll
      for (int n = 0; n < 4; ++n) h[n] = load_as_f32(a.h0, hoff + n, a.h0_dtype);
    } else {
#pragma unroll
      for (int n = 0; n < 4; ++n) h[n] = 0.f;
    }
    const bf16* wr = reinterpret_cast<const bf16*>(a.w_dt_pad) + (int64_t)c * a.Rp;
#pragma unroll
    for (int r = 0; r < R; ++r) w[r] = __bfloat162float(wr[r]);
  }
  const float Dv = a.D ? a.D[c] : 0.f;
  const float bias = a.dt_bias ? a.dt_bias[c] : 0.f;

  const bf16* ug = reinterpret_cast<const bf16*>(a.u) + (int64_t)b * a.u_bs + c0;
  const bf16* zg = reinterpret_cast<const bf16*>(a.z) + (int64_t)b * a.z_bs + c0;
  const bf16* xg = reinterpret_cast<const bf16*>(a.xdbl) + (int64_t)b * a.x_bs;
  bf16* yg = reinterpret_cast<bf16*>(a.y) + (int64_t)b * a.y_bs + c0;
  const int xchunks = a.Xp / 8;                // 16-byte chunks per x_dbl row
  auto phys = [&](int t) -> int64_t { return a.reverse ? (int64_t)(L - 1 - t) : (int64_t)t; };

  auto issue_tile = [&](int tile, int st) {
    const int t0 = tile * kTT;
    {
      const int row = tid >> 2, ch = tid & 3;  // 32 rows x 4 chunks of 8 channels
      const int t = t0 + row;
      const bool ok = t < L;
      const int64_t pr = ok ? phys(t) : 0;
      cp_async16(sbase + sp.u(st) + row * kUZRowBytes + ch * 16, ug + pr * a.u_ts + ch * 8, ok);
      cp_async16(sbase + sp.z(st) + row * kUZRowBytes + ch * 16, zg + pr * a.z_ts + ch * 8, ok);
    }
    for (int e = tid; e < kTT * xchunks; e += kThreads) {
      const int row = e / xchunks, ch = e - row * xchunks;
      const int t = t0 + row;
      const bool ok = t < L;
      const int64_t pr = ok ? phys(t) : 0;
      cp_async16(sbase + sp.x(st) + (row * xchunks + ch) * 16, xg + pr * a.x_ts + ch * 8, ok);
    }
  };

  const int ntiles = (L + kTT - 1) / kTT;
  issue_tile(0, 0);
  cp_async_commit();

  for (int tile = 0; tile < ntiles; ++tile) {
    const int st = tile & 1;
    const int t0 = tile * kTT;
    if (tile + 1 < ntiles) issue_tile(tile + 1, st ^ 1);
    cp_async_commit();
    cp_async_wait<1>();
    __syncthreads();

    // ---- expand this tile's x_dbl rows to fp32: dt_low -> sDT, [B | C] -> sBC ----------------
    {
      const uint32_t* xr = reinterpret_cast<const uint32_t*>(smem + sp.x(st));
      float* sdt = reinterpret_cast<float*>(smem + sp.dt);
      float* sbc = reinterpret_cast<float*>(smem + sp.bc);
      const int xw = a.Xp / 2;                 // 32-bit words per row
      for (int e = tid; e < kTT * (X / 2); e += kThreads) {
        const int row = e / (X / 2), p = e - row * (X / 2);
        const uint32_t v = xr[row * xw + p];
        const float lo = __uint_as_float(v << 16), hi = __uint_as_float(v & 0xffff0000u);
        float* dst = (2 * p < R) ? sdt + row * kDts + 2 * p : sbc + row * (2 * kN) + (2 * p - R);
        *reinterpret_cast<float2*>(dst) = make_float2(lo, hi);
      }
    }
    __syncthreads();

    // ---- the scan over the tile, 4 tokens per step -----------------------------------------------
    {
      const bf16* su = reinterpret_cast<const bf16*>(smem + sp.u(st));
      const bf16* sz = reinterpret_cast<const bf16*>(smem + sp.z(st));
      const float* sdt = reinterpret_cast<const float*>(smem + sp.dt);
      const float* sbc = reinterpret_cast<const float*>(smem + sp.bc);
      bf16* sy = reinterpret_cast<bf16*>(smem + sp.y);
      const int src0 = lane & ~3;
#pragma unroll 2
      for (int g = 0; g < kTT / 4; ++g) {
        const int row = 4 * g + j;
        // owner part: this lane's token of the group
        const float uval = __bfloat162float(su[row * (kUZRowBytes / 2) + cl]);
        const float zval = __bfloat162float(sz[row * (kUZRowBytes / 2) + cl]);
        float acc0 = bias, acc1 = 0.f;
        const float4* dtr = reinterpret_cast<const float4*>(sdt + row * kDts);
#pragma unroll
        for (int r4 = 0; r4 < R / 4; ++r4) {
          const float4 v = dtr[r4];
          acc0 = fmaf(v.x, w[4 * r4 + 0], acc0);
          acc1 = fmaf(v.y, w[4 * r4 + 1], acc1);
          acc0 = fmaf(v.z, w[4 * r4 + 2], acc0);
          acc1 = fmaf(v.w, w[4 * r4 + 3], acc1);
        }
        float delta = softplus_mufu(acc0 + acc1);
        if (t0 + row >= L) delta = 0.f;        // padding token: decay 1, drive 0 -> state untouched
        const float du = delta * uval;
        const float gate = silu_tanh(zval);
        const float skip = Dv * uval;

        float p[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          const float dk = __shfl_sync(0xffffffffu, delta, src0 | k);
          const float duk = __shfl_sync(0xffffffffu, du, src0 | k);
          const float4 Bv = *reinterpret_cast<const float4*>(sbc + (4 * g + k) * (2 * kN) + 4 * j);
          const float4 Cv = *reinterpret_cast<const float4*>(sbc + (4 * g + k) * (2 * kN) + kN + 4 * j);
          if constexpr (kPacked) {
            const float2 d2 = make_float2(dk, dk), du2 = make_float2(duk, duk);
            const float2 x01 = __fmul2_rn(d2, make_float2(A2[0], A2[1]));
            const float2 x23 = __fmul2_rn(d2, make_float2(A2[2], A2[3]));
            const float2 e01 = make_float2(ex2_approx(x01.x), ex2_approx(x01.y));
            const float2 e23 = make_float2(ex2_approx(x23.x), ex2_approx(x23.y));
            const float2 b01 = __fmul2_rn(du2, make_float2(Bv.x, Bv.y));
            const float2 b23 = __fmul2_rn(du2, make_float2(Bv.z, Bv.w));
            const float2 h01 = __ffma2_rn(e01, make_float2(h[0], h[1]), b01);
            const float2 h23 = __ffma2_rn(e23, make_float2(h[2], h[3]), b23);
            h[0] = h01.x; h[1] = h01.y; h[2] = h23.x; h[3] = h23.y;
            float2 q = __fmul2_rn(h01, make_float2(Cv.x, Cv.y));
            q = __ffma2_rn(h23, make_float2(Cv.z, Cv.w), q);
            p[k] = q.x + q.y;
          } else {
            const float e0 = ex2_approx(dk * A2[0]), e1 = ex2_approx(dk * A2[1]);
            const float e2 = ex2_approx(dk * A2[2]), e3 = ex2_approx(dk * A2[3]);
            h[0] = fmaf(e0, h[0], duk * Bv.x);
            h[1] = fmaf(e1, h[1], duk * Bv.y);
            h[2] = fmaf(e2, h[2], duk * Bv.z);
            h[3] = fmaf(e3, h[3], duk * Bv.w);
            p[k] = fmaf(h[0], Cv.x, h[1] * Cv.y) + fmaf(h[2], Cv.z, h[3] * Cv.w);
          }
        }
        // transpose-reduce: lane j ends up with sum over the 4 lanes of p[j]
        const bool odd = j & 1, hi2 = j & 2;
        const float keep0 = odd ? p[1] : p[0], send0 = odd ? p[0] : p[1];
        const float keep1 = odd ? p[3] : p[2], send1 = odd ? p[2] : p[3];
        const float q0 = keep0 + __shfl_xor_sync(0xffffffffu, send0, 1);
        const float q1 = keep1 + __shfl_xor_sync(0xffffffffu, send1, 1);
        const float keep = hi2 ? q1 : q0, send = hi2 ? q0 : q1;
        const float ysum = keep + __shfl_xor_sync(0xffffffffu, send, 2);
        sy[row * (kYRowBytes / 2) + cl] = __float2bfloat16_rn((ysum + skip) * gate);
      }
    }
    __syncthreads();

    // ---- y tile out: 32 rows x 64 bytes as 16-byte stores ------------------------------------------
    {
      const int row = tid >> 2, ch = tid & 3;
      const int t = t0 + row;
      if (t < L) {
        const uint4 v = *reinterpret_cast<const uint4*>(smem + sp.y + row * kYRowBytes + ch * 16);
        *reinterpret_cast<uint4*>(yg + phys(t) * a.y_ts + ch * 8) = v;
      }
    }
    // the next iteration's first __syncthreads orders these reads of sY before its next writes
  }

  if (a.h_last != nullptr) {
    *reinterpret_cast<float4*>(a.h_last + ((int64_t)b * a.Di + c) * kN + 4 * j) =
        make_float4(h[0], h[1], h[2], h[3]);
  }
}

int variant() {
  static int v = [] {
    const char* e = std::getenv("VMB_SCAN_VARIANT");
    return e ? std::atoi(e) : 0;
  }();
  return v;
}

template <int R>
int launch(const FastScanArgs& a, cudaStream_t st) {
  const Smem sp = smem_plan(R, a.Xp);
  dim3 grid(a.Di / kCh, a.B);
  const bool packed = (variant() & 1) == 0;
  auto kern = packed ? scan_fast_kernel<R, true> : scan_fast_kernel<R, false>;
  if (sp.total > 48 * 1024)
    VMB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, sp.total));
  kern<<<grid, kThreads, sp.total, st>>>(a);
  VMB_LAUNCH_CHECK("scan_fast_kernel");
  return VMB_OK;
}

}  // namespace

bool scan_fast_supported(const FastScanArgs& a) {
  auto al16 = [](const void* p) { return reinterpret_cast<uintptr_t>(p) % 16 == 0; };
  const bool r_ok = a.R == 12 || a.R == 24 || a.R == 36;
  return a.N == kN && r_ok && a.Di % kCh == 0 && a.Xp % 8 == 0 && a.Xp >= a.R + 2 * kN &&
         a.Rp >= a.R && a.B >= 1 && a.B <= 65535 && a.L >= 1 && a.w_dt_pad != nullptr &&
         al16(a.u) && al16(a.z) && al16(a.xdbl) && al16(a.y) && al16(a.A2) &&
         (a.h_last == nullptr || al16(a.h_last)) &&
         a.u_bs % 8 == 0 && a.u_ts % 8 == 0 && a.z_bs % 8 == 0 && a.z_ts % 8 == 0 &&
         a.x_bs % 8 == 0 && a.x_ts % 8 == 0 && a.y_bs % 8 == 0 && a.y_ts % 8 == 0 &&
         (variant() & 2) == 0;
}

int scan_fast(const FastScanArgs& a, cudaStream_t st) {
  switch (a.R) {
    case 12: return launch<12>(a, st);
    case 24: return launch<24>(a, st);
    case 36: return launch<36>(a, st);
    default: VMB_UNSUPPORTED("scan_fast: dt_rank %d not built", a.R);
  }
}

}  // namespace vmb
