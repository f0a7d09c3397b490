// Depthwise causal conv1d (+SiLU), token-major, with streaming history (HBM-bound).
//
// Stands in for causal_conv1d_fn and the torch.cat / F.pad state handling around it
// (reference models/videomamba/mamba_simple.py:381-404) and for causal_conv1d_update (:468-474).
// A thread owns VEC consecutive channels (one 128-bit load per token) and slides a W-tap
// window down a chunk of tokens, so x is read once (+ W-1 halo rows per chunk) and y written once.
// Algorithmic bytes per token: 2 * Di * sizeof(T).
// Two kernels: conv1d_fwd_kernel (any dtype / width, all loads of a 16-token chunk up front) and
// conv1d_ring_kernel (bf16, d_conv 4, the production shape: rows staged by cp.async into a private
// ring of shared-memory slots, 0.83 of the HBM peak; CTAs walk batch and chunks back to front so the
// kernel starts on the rows in_proj wrote last -- DESIGN.md 3.3, 3.4).

#include "common.cuh"

namespace vmb {
namespace {

constexpr int kChunk = 16;  // tokens per thread along the sequence

template <typename T, int VEC> struct VecIO;
template <> struct VecIO<float, 4> {
  using raw = float4;
  static __device__ __forceinline__ raw ldg(const float* p) { return *reinterpret_cast<const float4*>(p); }
  static __device__ __forceinline__ raw zero() { return make_float4(0.f, 0.f, 0.f, 0.f); }
  static __device__ __forceinline__ void unpack(const raw& v, float* f) {
    f[0] = v.x; f[1] = v.y; f[2] = v.z; f[3] = v.w;
  }
  static __device__ __forceinline__ void store(float* p, const float* f) {
    *reinterpret_cast<float4*>(p) = make_float4(f[0], f[1], f[2], f[3]);
  }
};
template <> struct VecIO<__nv_bfloat16, 8> {
  using raw = uint4;
  static __device__ __forceinline__ raw ldg(const __nv_bfloat16* p) { return *reinterpret_cast<const uint4*>(p); }
  static __device__ __forceinline__ raw zero() { return make_uint4(0u, 0u, 0u, 0u); }
  static __device__ __forceinline__ void unpack(const raw& v, float* f) {
    const uint32_t w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      f[2 * i] = __uint_as_float(w[i] << 16);
      f[2 * i + 1] = __uint_as_float(w[i] & 0xffff0000u);
    }
  }
  static __device__ __forceinline__ void store(__nv_bfloat16* p, const float* f) {
    uint32_t w[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      __nv_bfloat162 h = __floats2bfloat162_rn(f[2 * i], f[2 * i + 1]);
      w[i] = *reinterpret_cast<uint32_t*>(&h);
    }
    *reinterpret_cast<uint4*>(p) = make_uint4(w[0], w[1], w[2], w[3]);
  }
};
template <typename T> struct VecIO<T, 1> {
  using raw = T;
  static __device__ __forceinline__ raw ldg(const T* p) { return *p; }
  static __device__ __forceinline__ raw zero() { return from_f32<T>(0.f); }
  static __device__ __forceinline__ void unpack(const raw& v, float* f) { f[0] = to_f32<T>(v); }
  static __device__ __forceinline__ void store(T* p, const float* f) { *p = from_f32<T>(f[0]); }
};

// A thread owns VEC consecutive channels and kChunk consecutive tokens.  All kChunk + W - 1 input
// rows of the chunk are requested before any arithmetic (independent 128-bit loads in flight),
// then the W-tap window slides down the rows held in registers.
template <typename T, int VEC, int W, bool kAccurate>
__global__ void __launch_bounds__(128)
conv1d_fwd_kernel(const T* __restrict__ x, int64_t x_bs, int64_t x_ts, const T* __restrict__ weight,
                  const T* __restrict__ bias, const void* __restrict__ cs_in, int cs_in_dtype,
                  T* __restrict__ y, int64_t y_bs, int64_t y_ts, void* __restrict__ cs_out,
                  int cs_out_dtype, int L, int Di, int silu, int reverse, int frame_len) {
  using IO = VecIO<T, VEC>;
  const int c0 = (blockIdx.x * blockDim.x + threadIdx.x) * VEC;
  if (c0 >= Di) return;
  const int b = blockIdx.z;
  const int t0 = blockIdx.y * kChunk;
  const T* xb = x + (int64_t)b * x_bs + c0;
  T* yb = y + (int64_t)b * y_bs + c0;
  // physical row of logical token i: forward, whole-sequence reversal, or frame-axis reversal (frames of
  // frame_len tokens back to front, tokens inside a frame front to back -- the 4-D flip of
  // BiMambaRefinerBlock, models/refiner_backbone.py:61-68)
  auto row = [&](int i) -> int64_t {
    if (!reverse) return (int64_t)i;
    if (frame_len <= 0) return (int64_t)(L - 1 - i);
    const int f = i / frame_len;
    return (int64_t)(L - (f + 1) * frame_len + (i - f * frame_len));
  };

  // rows t0 - (W-1) .. t0 + kChunk - 1 of the logical sequence (negative: carried state / zeros)
  typename IO::raw raw[kChunk + W - 1];
#pragma unroll
  for (int i = 0; i < kChunk + W - 1; ++i) {
    const int t = t0 + i - (W - 1);
    raw[i] = (t >= 0 && t < L) ? IO::ldg(xb + row(t) * x_ts) : IO::zero();
  }
  float w[VEC][W], bv[VEC];
#pragma unroll
  for (int v = 0; v < VEC; ++v) {
#pragma unroll
    for (int k = 0; k < W; ++k) w[v][k] = to_f32<T>(weight[(int64_t)(c0 + v) * W + k]);
    bv[v] = bias ? to_f32<T>(bias[c0 + v]) : 0.f;
  }

  float win[W][VEC];  // win[k] = hist[t + k - (W-1)]
#pragma unroll
  for (int k = 0; k < W - 1; ++k) {
    const int t = t0 + k - (W - 1);
    if (t < 0 && cs_in != nullptr) {
#pragma unroll
      for (int v = 0; v < VEC; ++v)
        win[k + 1][v] = load_as_f32(cs_in, ((int64_t)b * Di + c0 + v) * W + (W + t), cs_in_dtype);
    } else {
      IO::unpack(raw[k], win[k + 1]);
    }
  }
#pragma unroll
  for (int i = 0; i < kChunk; ++i) {
    const int t = t0 + i;
    if (t >= L) break;
#pragma unroll
    for (int k = 0; k < W - 1; ++k)
#pragma unroll
      for (int v = 0; v < VEC; ++v) win[k][v] = win[k + 1][v];
    IO::unpack(raw[i + W - 1], win[W - 1]);
    float o[VEC];
#pragma unroll
    for (int v = 0; v < VEC; ++v) {
      float acc = bv[v];
#pragma unroll
      for (int k = 0; k < W; ++k) acc = fmaf(w[v][k], win[k][v], acc);
      o[v] = silu ? silu_f<kAccurate>(acc) : acc;
    }
    IO::store(yb + row(t) * y_ts, o);
  }

  // The CTA that owns the final chunk also emits the next conv state: hist[L-W .. L-1].
  if (cs_out != nullptr && t0 < L && t0 + kChunk >= L) {
#pragma unroll
    for (int k = 0; k < W; ++k) {
      const int t = L - W + k;
      float f[VEC];
      if (t >= 0) {
        IO::unpack(IO::ldg(xb + row(t) * x_ts), f);
      } else if (cs_in != nullptr) {
#pragma unroll
        for (int v = 0; v < VEC; ++v)
          f[v] = load_as_f32(cs_in, ((int64_t)b * Di + c0 + v) * W + (W + t), cs_in_dtype);
      } else {
#pragma unroll
        for (int v = 0; v < VEC; ++v) f[v] = 0.f;
      }
#pragma unroll
      for (int v = 0; v < VEC; ++v)
        store_from_f32(cs_out, ((int64_t)b * Di + c0 + v) * W + k, cs_out_dtype, f[v]);
    }
  }
}

// Production shape (bf16, d_conv 4, 16-byte aligned rows): a thread owns 8 channels and STREAMS down
// `tok` tokens; channel pairs are processed with packed FFMA2.
constexpr int kSub = 8;     // shortest sequence the streaming kernel takes is 2 * kSub tokens

__device__ __forceinline__ void unpack_pairs(const uint4& v, float2 (&f)[4]) {
  const uint32_t w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
  for (int i = 0; i < 4; ++i)
    f[i] = make_float2(__uint_as_float(w[i] << 16), __uint_as_float(w[i] & 0xffff0000u));
}

// Ring staging: the rows a thread is going to consume are brought
// in by 16-byte cp.async into a private ring of shared-memory slots ([slot][thread], conflict free),
// kRing rows deep, so the bytes in flight per SM no longer depend on registers: 5 CTAs x 96 threads
// x 8-12 rows x 16 B = 60-90 KB outstanding per SM at the shipped depth of 12, what the HBM
// latency-bandwidth product needs.  A thread only ever reads slots it filled itself, so cp.async.wait_group is the only
// synchronisation.  Groups of kGrp rows are committed together; group k + kRing / kGrp is issued
// into the slots of group k right after group k has been consumed.
constexpr int kGrp = 4;

__device__ __forceinline__ void cp_async16_zfill(uint32_t dst, const void* src, bool valid) {
  const int sz = valid ? 16 : 0;
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(src), "r"(sz) : "memory");
}
template <int kPending> __device__ __forceinline__ void cp_async_wait_group() {
  asm volatile("cp.async.wait_group %0;" ::"n"(kPending) : "memory");
}

template <bool kSilu, int kRing>
__global__ void __launch_bounds__(96, 5)
conv1d_ring_kernel(const __nv_bfloat16* __restrict__ x, int64_t x_bs, int64_t x_ts,
                   const __nv_bfloat16* __restrict__ weight, const __nv_bfloat16* __restrict__ bias,
                   const void* __restrict__ cs_in, int cs_in_dtype, __nv_bfloat16* __restrict__ y,
                   int64_t y_bs, int64_t y_ts, void* __restrict__ cs_out, int cs_out_dtype, int L,
                   int Di, int reverse, int tok) {
  constexpr int W = 4;
  constexpr int kAhead = kRing / kGrp;              // groups in flight
  extern __shared__ __align__(16) uint8_t ring_smem[];
  const int c0 = (blockIdx.x * blockDim.x + threadIdx.x) * 8;
  if (c0 >= Di) return;
  // CTAs are dispatched x-, then y-, then z-fastest: walk batch and chunks back to front, so the kernel
  // starts on the rows in_proj wrote last (still in L2) and x_proj, which starts at row 0, finds
  // this kernel's most recent output there
  const int b = gridDim.z - 1 - blockIdx.z;
  const int t0 = (gridDim.y - 1 - blockIdx.y) * tok;
  const int tend = min(L, t0 + tok);                // tokens [t0, tend)
  const __nv_bfloat16* xb = x + (int64_t)b * x_bs + c0;
  __nv_bfloat16* yb = y + (int64_t)b * y_bs + c0;
  const int dir = reverse ? -1 : 1;
  const int r0 = reverse ? L - 1 : 0;               // logical token t lives at row r0 + dir * t
  const int xs = (int)x_ts, ys = (int)y_ts;
  const uint32_t slot0 = static_cast<uint32_t>(__cvta_generic_to_shared(ring_smem)) + threadIdx.x * 16;
  const uint32_t slot_pitch = blockDim.x * 16;
  const uint8_t* const my = ring_smem + threadIdx.x * 16;
  auto issue_group = [&](int grp, int pos) {        // rows t0 + grp * kGrp + i -> slots pos * kGrp + i
#pragma unroll
    for (int i = 0; i < kGrp; ++i) {
      const int t = t0 + grp * kGrp + i;
      const bool ok = t < tend;
      cp_async16_zfill(slot0 + (pos * kGrp + i) * slot_pitch, xb + (int64_t)(r0 + dir * (ok ? t : t0)) * xs, ok);
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
  };
  const int ngroups = (tend - t0 + kGrp - 1) / kGrp;
#pragma unroll
  for (int k = 0; k < kAhead; ++k) issue_group(k, k);   // groups beyond the chunk are zero fills

  auto ldrow = [&](int t) -> uint4 {
    return (t >= 0 && t < L) ? __ldg(reinterpret_cast<const uint4*>(xb + (int64_t)(r0 + dir * t) * xs))
                             : make_uint4(0u, 0u, 0u, 0u);
  };
  uint4 hist[W - 1];
#pragma unroll
  for (int k = 0; k < W - 1; ++k) hist[k] = ldrow(t0 - (W - 1) + k);

  float2 w2[4][W], b2[4];
  {
    const uint4 wv[4] = {__ldg(reinterpret_cast<const uint4*>(weight + (int64_t)c0 * W)),
                         __ldg(reinterpret_cast<const uint4*>(weight + (int64_t)c0 * W + 8)),
                         __ldg(reinterpret_cast<const uint4*>(weight + (int64_t)c0 * W + 16)),
                         __ldg(reinterpret_cast<const uint4*>(weight + (int64_t)c0 * W + 24))};
#pragma unroll
    for (int p = 0; p < 4; ++p) {                   // 8 bf16: channel 2p taps 0..3, channel 2p+1 taps 0..3
      const uint32_t q[4] = {wv[p].x, wv[p].y, wv[p].z, wv[p].w};
      const float e0[4] = {__uint_as_float(q[0] << 16), __uint_as_float(q[0] & 0xffff0000u),
                           __uint_as_float(q[1] << 16), __uint_as_float(q[1] & 0xffff0000u)};
      const float e1[4] = {__uint_as_float(q[2] << 16), __uint_as_float(q[2] & 0xffff0000u),
                           __uint_as_float(q[3] << 16), __uint_as_float(q[3] & 0xffff0000u)};
#pragma unroll
      for (int k = 0; k < W; ++k) w2[p][k] = make_float2(e0[k], e1[k]);
    }
    if (bias) {
      float2 t[4];
      unpack_pairs(__ldg(reinterpret_cast<const uint4*>(bias + c0)), t);
#pragma unroll
      for (int p = 0; p < 4; ++p) b2[p] = t[p];
    } else {
#pragma unroll
      for (int p = 0; p < 4; ++p) b2[p] = make_float2(0.f, 0.f);
    }
  }

  float2 win[W - 1][4];                             // rows t-3, t-2, t-1
#pragma unroll
  for (int k = 0; k < W - 1; ++k) {
    const int t = t0 - (W - 1) + k;
    if (t < 0 && cs_in != nullptr) {
#pragma unroll
      for (int p = 0; p < 4; ++p)
        win[k][p] = make_float2(
            load_as_f32(cs_in, ((int64_t)b * Di + c0 + 2 * p) * W + (W + t), cs_in_dtype),
            load_as_f32(cs_in, ((int64_t)b * Di + c0 + 2 * p + 1) * W + (W + t), cs_in_dtype));
    } else {
      unpack_pairs(hist[k], win[k]);
    }
  }

  int pos = 0;
#pragma unroll 1
  for (int grp = 0; grp < ngroups; ++grp) {
    cp_async_wait_group<kAhead - 1>();              // group grp has landed (own writes: no barrier needed)
    uint4 cur[kGrp];
#pragma unroll
    for (int i = 0; i < kGrp; ++i)
      cur[i] = *reinterpret_cast<const uint4*>(my + (size_t)(pos * kGrp + i) * slot_pitch);
    issue_group(grp + kAhead, pos);                 // refill the slots just read (program order: after the loads)
    pos = pos + 1 == kAhead ? 0 : pos + 1;
    const int ts = t0 + grp * kGrp;
#pragma unroll
    for (int i = 0; i < kGrp; ++i) {
      float2 xin[4];
      unpack_pairs(cur[i], xin);
      uint32_t o[4];
#pragma unroll
      for (int p = 0; p < 4; ++p) {
        float2 acc = __ffma2_rn(w2[p][0], win[0][p], b2[p]);
        acc = __ffma2_rn(w2[p][1], win[1][p], acc);
        acc = __ffma2_rn(w2[p][2], win[2][p], acc);
        acc = __ffma2_rn(w2[p][3], xin[p], acc);
        if (kSilu) { acc.x = silu_fast(acc.x); acc.y = silu_fast(acc.y); }
        const __nv_bfloat162 h = __floats2bfloat162_rn(acc.x, acc.y);
        o[p] = *reinterpret_cast<const uint32_t*>(&h);
        win[0][p] = win[1][p]; win[1][p] = win[2][p]; win[2][p] = xin[p];
      }
      if (ts + i < tend)
        *reinterpret_cast<uint4*>(yb + (int64_t)(r0 + dir * (ts + i)) * ys) = make_uint4(o[0], o[1], o[2], o[3]);
    }
  }
  cp_async_wait_group<0>();                         // nothing of ours may still be landing when the CTA retires

  // The thread that owns the final chunk also emits the next conv state: hist[L-W .. L-1] (pre-conv x).
  if (cs_out != nullptr && t0 < L && t0 + tok >= L) {
#pragma unroll
    for (int k = 0; k < W; ++k) {
      const int t = L - W + k;
      float2 f[4];
      if (t >= 0) {
        unpack_pairs(ldrow(t), f);
      } else if (cs_in != nullptr) {
#pragma unroll
        for (int p = 0; p < 4; ++p)
          f[p] = make_float2(
              load_as_f32(cs_in, ((int64_t)b * Di + c0 + 2 * p) * W + (W + t), cs_in_dtype),
              load_as_f32(cs_in, ((int64_t)b * Di + c0 + 2 * p + 1) * W + (W + t), cs_in_dtype));
      } else {
#pragma unroll
        for (int p = 0; p < 4; ++p) f[p] = make_float2(0.f, 0.f);
      }
#pragma unroll
      for (int p = 0; p < 4; ++p) {
        store_from_f32(cs_out, ((int64_t)b * Di + c0 + 2 * p) * W + k, cs_out_dtype, f[p].x);
        store_from_f32(cs_out, ((int64_t)b * Di + c0 + 2 * p + 1) * W + k, cs_out_dtype, f[p].y);
      }
    }
  }
}

template <typename T, int W>
__global__ void conv1d_update_kernel(const T* __restrict__ x, int64_t x_bs, void* __restrict__ cs,
                                     int cs_dtype, const T* __restrict__ weight,
                                     const T* __restrict__ bias, T* __restrict__ y, int64_t y_bs,
                                     int B, int Di, int silu) {
  const int d = blockIdx.x * blockDim.x + threadIdx.x;
  const int b = blockIdx.y;
  if (d >= Di) return;
  const int64_t base = ((int64_t)b * Di + d) * W;
  float win[W];
#pragma unroll
  for (int k = 0; k < W - 1; ++k) win[k] = load_as_f32(cs, base + k + 1, cs_dtype);
  win[W - 1] = to_f32<T>(x[(int64_t)b * x_bs + d]);
  // the state is stored in its own dtype first; the conv then reads the ROUNDED values, as
  // the reference does (it convolves conv_state after writing x into it)
  float acc = bias ? to_f32<T>(bias[d]) : 0.f;
#pragma unroll
  for (int k = 0; k < W; ++k) {
    store_from_f32(cs, base + k, cs_dtype, win[k]);
    const float r = load_as_f32(cs, base + k, cs_dtype);
    acc = fmaf(to_f32<T>(weight[(int64_t)d * W + k]), r, acc);
  }
  y[(int64_t)b * y_bs + d] = from_f32<T>(silu ? silu_accurate(acc) : acc);
}

template <typename T, int VEC, int W>
int launch_fwd(const void* x, int64_t x_bs, int64_t x_ts, const void* weight, const void* bias,
               const void* cs_in, int cs_in_dtype, void* y, int64_t y_bs, int64_t y_ts,
               void* cs_out, int cs_out_dtype, int B, int L, int Di, int silu, int reverse,
               int frame_len, cudaStream_t st) {
  const int cthreads = (Di + VEC - 1) / VEC;
  if constexpr (sizeof(T) == 2 && VEC == 8 && W == 4) {
    if (!(reverse && frame_len > 0) &&          // the frame-axis walk runs on the chunk kernel below
        reinterpret_cast<uintptr_t>(weight) % 16 == 0 &&
        (bias == nullptr || reinterpret_cast<uintptr_t>(bias) % 16 == 0) && L >= 2 * kSub) {
      const int blk = cthreads >= 96 ? 96 : ((cthreads + 31) / 32) * 32;
      const int gx = (cthreads + blk - 1) / blk;
      constexpr int kRing = 12;   // 18 KB per CTA; 12 / 16 / 24 rows measured: 55.9 / 56.8 / 57.0 us alone, the
                                  // smaller footprint shares SMs with another step's scan more easily
      const int tokr = 40;   // tokens per thread; measured 24 .. 128 at batch 32: 56.3 us at 40, 58.4 at 64, 64.8 at 128
      const size_t smem = (size_t)kRing * blk * 16;
      dim3 g(gx, (L + tokr - 1) / tokr, B);
      if (silu)
        conv1d_ring_kernel<true, kRing><<<g, blk, smem, st>>>((const T*)x, x_bs, x_ts, (const T*)weight,
                                                              (const T*)bias, cs_in, cs_in_dtype, (T*)y, y_bs,
                                                              y_ts, cs_out, cs_out_dtype, L, Di, reverse, tokr);
      else
        conv1d_ring_kernel<false, kRing><<<g, blk, smem, st>>>((const T*)x, x_bs, x_ts, (const T*)weight,
                                                               (const T*)bias, cs_in, cs_in_dtype, (T*)y, y_bs,
                                                               y_ts, cs_out, cs_out_dtype, L, Di, reverse, tokr);
      VMB_LAUNCH_CHECK("conv1d_ring_kernel");
      return VMB_OK;
    }
  }
  const int block = cthreads >= 128 ? 128 : ((cthreads + 31) / 32) * 32;
  dim3 grid((cthreads + block - 1) / block, (L + kChunk - 1) / kChunk, B);
  constexpr bool kAccurate = sizeof(T) == 4;
  conv1d_fwd_kernel<T, VEC, W, kAccurate><<<grid, block, 0, st>>>(
      (const T*)x, x_bs, x_ts, (const T*)weight, (const T*)bias, cs_in, cs_in_dtype, (T*)y, y_bs,
      y_ts, cs_out, cs_out_dtype, L, Di, silu, reverse, frame_len);
  VMB_LAUNCH_CHECK("conv1d_fwd_kernel");
  return VMB_OK;
}

template <typename T, int VEC>
int dispatch_w(int W, const void* x, int64_t x_bs, int64_t x_ts, const void* weight,
               const void* bias, const void* cs_in, int cs_in_dtype, void* y, int64_t y_bs,
               int64_t y_ts, void* cs_out, int cs_out_dtype, int B, int L, int Di, int silu,
               int reverse, int frame_len, cudaStream_t st) {
#define VMB_CW(WW)                                                                             \
  case WW:                                                                                     \
    return launch_fwd<T, VEC, WW>(x, x_bs, x_ts, weight, bias, cs_in, cs_in_dtype, y, y_bs,    \
                                  y_ts, cs_out, cs_out_dtype, B, L, Di, silu, reverse, frame_len, st)
  switch (W) {
    VMB_CW(1); VMB_CW(2); VMB_CW(3); VMB_CW(4);
    default: VMB_UNSUPPORTED("causal_conv1d: d_conv=%d not supported (1..4)", W);
  }
#undef VMB_CW
}

}  // namespace
}  // namespace vmb

extern "C" int vmb_causal_conv1d_fwd(const void* x, int64_t x_bs, int64_t x_ts, const void* weight,
                                     const void* bias, const void* cs_in, int cs_in_dtype, void* y,
                                     int64_t y_bs, int64_t y_ts, void* cs_out, int cs_out_dtype,
                                     int B, int L, int Di, int W, int silu, int reverse, int frame_len,
                                     int dtype,
                                     vmb_stream_t stream) {
  using namespace vmb;
  VMB_CHECK_ARG(x && weight && y, "causal_conv1d: null x / weight / y");
  VMB_CHECK_ARG(dtype_ok(dtype), "causal_conv1d: bad dtype %d", dtype);
  VMB_CHECK_ARG(B >= 0 && L >= 0 && Di > 0, "causal_conv1d: bad sizes");
  if (!cs_in) cs_in_dtype = VMB_F32;
  if (!cs_out) cs_out_dtype = VMB_F32;
  VMB_CHECK_ARG(dtype_ok(cs_in_dtype) && dtype_ok(cs_out_dtype), "causal_conv1d: bad state dtype");
  VMB_CHECK_ARG(B <= 65535, "causal_conv1d: batch %d > 65535", B);
  VMB_CHECK_ARG(frame_len >= 0 && (frame_len == 0 || L % frame_len == 0),
                "causal_conv1d: L=%d is not a whole number of frames of %d tokens", L, frame_len);
  if (B == 0) return VMB_OK;
  if (L == 0) {
    // nothing to convolve; the state just carries over (or stays zero)
    if (cs_out) {
      const size_t n = (size_t)B * Di * W;
      if (cs_in && cs_in_dtype == cs_out_dtype)
        VMB_CUDA(cudaMemcpyAsync(cs_out, cs_in, n * dtype_size(cs_out_dtype),
                                 cudaMemcpyDeviceToDevice, as_stream(stream)));
      else if (!cs_in)
        VMB_CUDA(cudaMemsetAsync(cs_out, 0, n * dtype_size(cs_out_dtype), as_stream(stream)));
      else
        VMB_UNSUPPORTED("causal_conv1d: L=0 with a state dtype change");
    }
    return VMB_OK;
  }
  cudaStream_t st = as_stream(stream);
  const int vec = dtype == VMB_BF16 ? 8 : 4;
  const bool aligned = Di % vec == 0 && x_bs % vec == 0 && x_ts % vec == 0 && y_bs % vec == 0 &&
                       y_ts % vec == 0 && reinterpret_cast<uintptr_t>(x) % 16 == 0 &&
                       reinterpret_cast<uintptr_t>(y) % 16 == 0;
  if (dtype == VMB_BF16) {
    if (aligned)
      return dispatch_w<__nv_bfloat16, 8>(W, x, x_bs, x_ts, weight, bias, cs_in, cs_in_dtype, y,
                                          y_bs, y_ts, cs_out, cs_out_dtype, B, L, Di, silu,
                                          reverse, frame_len, st);
    return dispatch_w<__nv_bfloat16, 1>(W, x, x_bs, x_ts, weight, bias, cs_in, cs_in_dtype, y, y_bs,
                                        y_ts, cs_out, cs_out_dtype, B, L, Di, silu, reverse, frame_len, st);
  }
  if (aligned)
    return dispatch_w<float, 4>(W, x, x_bs, x_ts, weight, bias, cs_in, cs_in_dtype, y, y_bs, y_ts,
                                cs_out, cs_out_dtype, B, L, Di, silu, reverse, frame_len, st);
  return dispatch_w<float, 1>(W, x, x_bs, x_ts, weight, bias, cs_in, cs_in_dtype, y, y_bs, y_ts,
                              cs_out, cs_out_dtype, B, L, Di, silu, reverse, frame_len, st);
}

extern "C" int vmb_causal_conv1d_update(const void* x, int64_t x_bs, void* conv_state, int cs_dtype,
                                        const void* weight, const void* bias, void* y,
                                        int64_t y_bs, int B, int Di, int W, int silu, int dtype,
                                        vmb_stream_t stream) {
  using namespace vmb;
  VMB_CHECK_ARG(x && conv_state && weight && y, "conv1d_update: null pointer");
  VMB_CHECK_ARG(dtype_ok(dtype) && dtype_ok(cs_dtype), "conv1d_update: bad dtype");
  VMB_CHECK_ARG(B <= 65535, "conv1d_update: batch %d > 65535", B);
  if (B <= 0) return VMB_OK;
  cudaStream_t st = as_stream(stream);
  dim3 grid((Di + 127) / 128, B);
#define VMB_CU(T, WW)                                                                          \
  conv1d_update_kernel<T, WW><<<grid, 128, 0, st>>>((const T*)x, x_bs, conv_state, cs_dtype,   \
                                                    (const T*)weight, (const T*)bias, (T*)y,   \
                                                    y_bs, B, Di, silu)
  if (dtype == VMB_BF16) {
    switch (W) {
      case 1: VMB_CU(__nv_bfloat16, 1); break;
      case 2: VMB_CU(__nv_bfloat16, 2); break;
      case 3: VMB_CU(__nv_bfloat16, 3); break;
      case 4: VMB_CU(__nv_bfloat16, 4); break;
      default: VMB_UNSUPPORTED("conv1d_update: d_conv=%d not supported", W);
    }
  } else {
    switch (W) {
      case 1: VMB_CU(float, 1); break;
      case 2: VMB_CU(float, 2); break;
      case 3: VMB_CU(float, 3); break;
      case 4: VMB_CU(float, 4); break;
      default: VMB_UNSUPPORTED("conv1d_update: d_conv=%d not supported", W);
    }
  }
#undef VMB_CU
  VMB_LAUNCH_CHECK("conv1d_update_kernel");
  return VMB_OK;
}
