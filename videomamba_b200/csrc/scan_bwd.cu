// Backward of the selective scan: the C entry point, the final reductions, and
// the any-shape kernels (true fp32, d_state <= 16, any Di / strides).  The bf16 production shapes run the
// kernels of scan_bwd_fast.cu (dispatched in run() below); both write the same slab / partial layouts.
//
// Forward (reference _selective_scan_ref, models/videomamba/mamba_simple.py:30-106):
//   delta = softplus(draw + bias);  a_t[n] = exp(delta_t A[n]);  h_t = a_t h_{t-1} + delta_t u_t B_t
//   ypre_t = <C_t, h_t> + D u_t;   out_t = ypre_t * silu(z_t);   h_last = h_L
// Backward, given dout and d(h_last), walking t = L-1 .. 0 with g = dLoss/dh_t:
//   dy = dout silu(z);  dz = dout ypre silu'(z);  g += dy C_t;  dC_t = sum_d dy h_t
//   dB_t = sum_d g delta u;  du = dy D + sum_n g delta B_t;  ddelta = sum_n g (a A h_{t-1} + u B_t)
//   dA[n] += g a h_{t-1} delta;  g *= a  (-> t-1);   ddraw = ddelta sigmoid(draw + bias);  dh0 = g after t = 0
//
// h_{t-1} is needed in reverse order.  Pass 1 (scan_ckpt_kernel) walks the sequence forward and stores
// the state at the start of every 8-token chunk ([batch][chunk][channel][state]: a warp's two channels are
// one 128-byte line); pass 2 (scan_bwd_kernel) walks the chunks back to
// front, recomputes the 8 states of a chunk from its checkpoint and runs the reverse recurrence on them.
//
// Thread layout: ONE LANE PER (channel, state) -- a half-warp is a channel, a warp two channels, a CTA
// of 8 warps 16 adjacent channels of one sequence -- so a batch of 8 already gives 3072 warps and the
// per-token dependent chain is one exp2 + one FMA.  Per-channel scalars of a chunk (delta, its
// derivative, u; the gate terms of dout / z) are computed once by the lanes of the half-warp (lane n < 8:
// token n; lane n >= 8: token n - 8) and handed round by shuffles.  Sums over the 16 states (ypre, du,
// ddelta) are transposing butterflies over the chunk's 8 tokens (8 / 15 shuffles per chunk instead of
// 4 per token and value).  dB_t / dC_t need a sum over channels: one shuffle adds the two channels of
// a warp, the 8 warps of a CTA meet in shared memory once per chunk, every CTA writes one slab row
// {dB_t[0..15], dC_t[0..15]} per token and a last kernel sums the Di / 16 slabs.  dA, dD and d(dt_bias)
// are summed over the batch the same way.  No atomics: results are deterministic.
#include <algorithm>

#include "internal.h"

namespace vmb {
namespace {

constexpr int kT = 8;          // tokens per chunk
constexpr int kWarps = 8;      // warps per CTA
constexpr int kChan = 2 * kWarps;   // channels per CTA
constexpr int kNMax = 16;

// sum over the 16 lanes of a half-warp of 8 per-lane values: afterwards every lane holds the total of
// value (lane >> 1) & 7  (8 shuffles)
__device__ __forceinline__ float halfwarp_sum8(float (&v)[8], int lane) {
#pragma unroll
  for (int s = 8, w = 4; w >= 1; s >>= 1, w >>= 1) {
    const bool up = (lane & s) != 0;
#pragma unroll
    for (int k = 0; k < w; ++k) {
      const float keep = up ? v[k + w] : v[k];
      const float send = up ? v[k] : v[k + w];
      v[k] = keep + __shfl_xor_sync(0xffffffffu, send, s);
    }
  }
  return v[0] + __shfl_xor_sync(0xffffffffu, v[0], 1);
}
// sum over the 16 lanes of a half-warp of 16 per-lane values: lane n ends with the total of value n & 15
__device__ __forceinline__ float halfwarp_sum16(float (&v)[16], int lane) {
#pragma unroll
  for (int s = 8; s >= 1; s >>= 1) {
    const bool up = (lane & s) != 0;
#pragma unroll
    for (int k = 0; k < s; ++k) {
      const float keep = up ? v[k + s] : v[k];
      const float send = up ? v[k] : v[k + s];
      v[k] = keep + __shfl_xor_sync(0xffffffffu, send, s);
    }
  }
  return v[0];
}

template <typename T, bool kAccurate>
__global__ void __launch_bounds__(kWarps * 32)
scan_ckpt_kernel(const vmb_scan_bwd_args a, float* __restrict__ ckpt, int nchunks) {
  __shared__ float sB[2][kT][kNMax];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int n = lane & 15, half = lane & 16;       // half = first lane of this channel's half-warp
  const int b = blockIdx.y;
  const int d = (blockIdx.x * kWarps + warp) * 2 + (lane >> 4);
  const int N = a.N, L = a.L;
  const bool valid = d < a.Di && n < N;
  const int dc = min(d, a.Di - 1);                  // clamped channel for loads (results of invalid lanes are unused)
  const float A2 = valid ? a.A2[(int64_t)d * N + n] : 0.f;
  float h = (valid && a.h0) ? load_as_f32(a.h0, ((int64_t)b * a.Di + d) * N + n, a.h0_dtype) : 0.f;
  const float bias = a.dt_bias ? a.dt_bias[dc] : 0.f;
  // per-lane walking pointers (advanced by one chunk per iteration: no 64-bit index arithmetic in the loop)
  const bool prep = n < kT;                         // lane n < 8 prepares token n of its channel
  const T* pu = reinterpret_cast<const T*>(a.u) + (int64_t)b * a.u_bstride + dc + (int64_t)n * a.u_tstride;
  const T* pd = reinterpret_cast<const T*>(a.delta) + (int64_t)b * a.d_bstride + dc + (int64_t)n * a.d_tstride;
  const int64_t su = (int64_t)kT * a.u_tstride, sd = (int64_t)kT * a.d_tstride;
  const bool tile_lane = tid < kT * kNMax && (tid & 15) < N;
  const T* pb = reinterpret_cast<const T*>(a.bc) + (int64_t)b * a.bc_bstride + (int64_t)(tid >> 4) * a.bc_tstride +
                a.b_off + (tid & 15);
  const int64_t sbc = (int64_t)kT * a.bc_tstride;
  float* pc = ckpt + (((int64_t)b * nchunks) * a.Di + d) * N + n;
  const int64_t sc = (int64_t)a.Di * N;
  struct Pre { float dv, uv, tb; };
  auto prefetch = [&](int nt) {                     // the next chunk's loads, one chunk ahead (see scan_bwd_kernel)
    Pre q{0.f, 0.f, 0.f};
    if (prep && n < nt) {
      q.uv = to_f32<T>(*pu);
      q.dv = to_f32<T>(*pd);
    }
    if (tile_lane && (tid >> 4) < nt) q.tb = to_f32<T>(*pb);
    pu += su; pd += sd; pb += sbc;
    return q;
  };
  Pre nxt = prefetch(min(kT, L));
  for (int c = 0; c < nchunks; ++c) {
    const int nt = min(kT, L - c * kT);
    const Pre cur = nxt;
    float (*tile)[kNMax] = sB[c & 1];               // double buffered: one barrier per chunk
    if (tid < kT * kNMax) tile[tid >> 4][tid & 15] = cur.tb;
    if (c + 1 < nchunks) nxt = prefetch(min(kT, L - (c + 1) * kT));
    float dv = 0.f;
    const float uv = cur.uv;
    if (prep && n < nt) {
      dv = cur.dv + bias;
      if (a.softplus) dv = softplus_f<kAccurate>(dv);
    }
    __syncthreads();
    if (valid) *pc = h;
    pc += sc;
#pragma unroll
    for (int t = 0; t < kT; ++t) {
      const float dt = __shfl_sync(0xffffffffu, dv, half + t);
      const float ut = __shfl_sync(0xffffffffu, uv, half + t);
      h = fmaf(exp2_f<kAccurate>(dt * A2), h, dt * ut * tile[t][n]);
    }
  }
}

template <typename T, bool kAccurate>
__global__ void __launch_bounds__(kWarps * 32)
scan_bwd_kernel(const vmb_scan_bwd_args a, const float* __restrict__ ckpt, int nchunks,
                float* __restrict__ bc_slabs, float* __restrict__ pA, float* __restrict__ pD,
                float* __restrict__ pBias) {
  __shared__ float sB[kT][kNMax];
  __shared__ float sC[kT][kNMax];
  __shared__ float wbuf[kWarps][kT][32];           // per warp: {dB_t[n], dC_t[n]} summed over its two channels
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int n = lane & 15, half = lane & 16;
  const int b = blockIdx.y;
  const int d = (blockIdx.x * kWarps + warp) * 2 + (lane >> 4);
  const int N = a.N, L = a.L;
  const int64_t Di = a.Di;
  const bool chan_ok = d < a.Di;
  const bool valid = chan_ok && n < N;
  const int dc = min(d, a.Di - 1);
  const bool lo = n < kT;                           // lanes 0..7 of a channel: delta / u of token n
  const int tk = n & 7;                             // token this lane prepares and finalises

  const float A2 = valid ? a.A2[(int64_t)d * N + n] : 0.f;
  const float An = A2 * kLn2;
  float g = (valid && a.dh_last) ? a.dh_last[((int64_t)b * Di + d) * N + n] : 0.f;
  float dA = 0.f, dD = 0.f, dBias = 0.f;
  const float Dv = a.D ? a.D[dc] : 0.f;
  const float bias = a.dt_bias ? a.dt_bias[dc] : 0.f;
  // Per-lane walking pointers, stepped back one chunk per iteration (no 64-bit index arithmetic in the loop).
  // Lane roles within a channel: lanes 0..7 read delta / u and write du of token n; lanes 8..15 read
  // dout / z and write ddelta / dz of token n - 8.
  const int64_t last0 = (int64_t)(nchunks - 1) * kT;                 // first token of the last chunk
  const T* pa;  const T* pb2;  int64_t sa, sb2;                     // the lane's two scalar streams
  if (lo) {
    pa = reinterpret_cast<const T*>(a.delta) + (int64_t)b * a.d_bstride + dc + (last0 + tk) * a.d_tstride;
    pb2 = reinterpret_cast<const T*>(a.u) + (int64_t)b * a.u_bstride + dc + (last0 + tk) * a.u_tstride;
    sa = (int64_t)kT * a.d_tstride;  sb2 = (int64_t)kT * a.u_tstride;
  } else {
    pa = reinterpret_cast<const T*>(a.dout) + (int64_t)b * a.dout_bstride + dc + (last0 + tk) * a.dout_tstride;
    pb2 = a.z ? reinterpret_cast<const T*>(a.z) + (int64_t)b * a.z_bstride + dc + (last0 + tk) * a.z_tstride : nullptr;
    sa = (int64_t)kT * a.dout_tstride;  sb2 = (int64_t)kT * a.z_tstride;
  }
  const bool has_z = a.z != nullptr;
  const bool load_b = lo || has_z;
  const float* pck = ckpt + (((int64_t)b * nchunks + (nchunks - 1)) * Di + d) * N + n;
  const int64_t sck = Di * N;
  const bool tile_lane = tid < kT * kNMax && (tid & 15) < N;
  const T* ptb = reinterpret_cast<const T*>(a.bc) + (int64_t)b * a.bc_bstride + (last0 + (tid >> 4)) * a.bc_tstride +
                 (tid & 15);
  const int64_t stb = (int64_t)kT * a.bc_tstride;
  const int b_off = a.b_off, c_off = a.c_off;
  // outputs: lanes 0..7 -> du, lanes 8..15 -> ddelta (and dz)
  T* po = (lo ? reinterpret_cast<T*>(a.du) : reinterpret_cast<T*>(a.ddelta)) + ((int64_t)b * L + last0 + tk) * Di + dc;
  const int64_t dz_ts = a.dz_tstride ? a.dz_tstride : Di, dz_bs = a.dz_tstride ? a.dz_bstride : (int64_t)L * Di;
  T* pz = a.dz ? reinterpret_cast<T*>(a.dz) + (int64_t)b * dz_bs + (last0 + tk) * dz_ts + dc : nullptr;
  const int64_t so = (int64_t)kT * Di, sz = (int64_t)kT * dz_ts;
  // slab of this CTA: [(slab * B + b) * L + t][32] = {dB_t[0..15], dC_t[0..15]} summed over its 16 channels
  float* pslab = bc_slabs + (((int64_t)blockIdx.x * a.B + b) * L + last0 + (tid >> 5)) * 32 + (tid & 31);

  // Global loads of a chunk (two per-channel scalars per lane, the checkpointed state, the B / C tile
  // elements) are issued one chunk AHEAD into registers: with 8 tokens of work per chunk and CTA-wide
  // barriers, a load issued where it is needed would expose its full latency every chunk.
  struct Pre { float r0, r1, hck, tb, tc; };
  auto prefetch = [&](int nt) {
    Pre q{0.f, 0.f, 0.f, 0.f, 0.f};
    if (tk < nt && chan_ok) {
      q.r0 = to_f32<T>(*pa);
      if (load_b) q.r1 = to_f32<T>(*pb2);
    }
    if (valid) q.hck = *pck;
    if (tile_lane && (tid >> 4) < nt) {
      q.tb = to_f32<T>(ptb[b_off]);
      q.tc = to_f32<T>(ptb[c_off]);
    }
    pa -= sa; pb2 -= sb2; pck -= sck; ptb -= stb;
    return q;
  };
  Pre nxt = prefetch(min(kT, L - (int)last0));
  for (int c = nchunks - 1; c >= 0; --c) {
    const int nt = min(kT, L - c * kT);
    const Pre cur = nxt;
    if (tid < kT * kNMax) {
      sB[tid >> 4][tid & 15] = cur.tb;
      sC[tid >> 4][tid & 15] = cur.tc;
    }
    if (c > 0) nxt = prefetch(kT);
    // ---- per-channel scalars of the chunk: lanes 0..7 {delta, softplus', u}, lanes 8..15 {dy, dz factor, -}
    float p0 = 0.f, p1 = 0.f, p2 = 0.f;
    if (tk < nt && chan_ok) {                        // lanes of a channel beyond Di contribute zeros everywhere
      if (lo) {
        const float raw = cur.r0 + bias;
        p0 = raw;
        p1 = 1.f;
        if (a.softplus) {
          p0 = softplus_f<kAccurate>(raw);
          // d softplus = sigmoid; torch's softplus is the identity above its threshold (20)
          p1 = raw > 20.f ? 1.f : 1.f / (1.f + (kAccurate ? expf(-raw) : ex2_approx(-raw * kLog2e)));
        }
        p2 = cur.r1;
      } else {
        const float gout = cur.r0;
        p0 = gout;
        if (has_z) {
          const float zv = cur.r1;
          const float s = 1.f / (1.f + (kAccurate ? expf(-zv) : ex2_approx(-zv * kLog2e)));
          p0 = gout * zv * s;                        // dy
          p1 = gout * s * (1.f + zv * (1.f - s));    // dz = p1 * ypre
        }
      }
    }
    __syncthreads();                                 // tiles staged (and last chunk's wbuf consumed)
    // ---- recompute the chunk forward from its checkpoint -------------------------------------------
    float h = cur.hck;
    float hp[kT], an[kT], yc[kT];
#pragma unroll
    for (int t = 0; t < kT; ++t) {
      const float dt = __shfl_sync(0xffffffffu, p0, half + t);
      const float ut = __shfl_sync(0xffffffffu, p2, half + t);
      hp[t] = h;
      an[t] = exp2_f<kAccurate>(dt * A2);
      h = fmaf(an[t], h, dt * ut * sB[t][n]);
      yc[t] = h * sC[t][n];
    }
    // ypre of token (lane >> 1) & 7 of this channel (without the D skip), then to the lane that owns dz
    const float ysum = halfwarp_sum8(yc, lane);
    const float ypre_tk = __shfl_sync(0xffffffffu, ysum, half + 2 * tk);
    // ---- reverse recurrence over the chunk -----------------------------------------------------------
    float red[16];                                   // [0..7] du contributions, [8..15] ddelta contributions
    float hn = h;                                    // h_t of the token being processed
#pragma unroll
    for (int t = kT - 1; t >= 0; --t) {
      const float dt = __shfl_sync(0xffffffffu, p0, half + t);
      const float ut = __shfl_sync(0xffffffffu, p2, half + t);
      const float dy = __shfl_sync(0xffffffffu, p0, half + 8 + t);
      const float bt = sB[t][n], ct = sC[t][n];
      g = fmaf(dy, ct, g);                           // dLoss / dh_t
      const float vC = dy * hn;                      // -> dC_t[n]
      const float vB = g * dt * ut;                  // -> dB_t[n]
      const float ahp = an[t] * hp[t];
      red[t] = g * dt * bt;
      red[8 + t] = g * fmaf(ahp, An, ut * bt);
      dA = fmaf(g * ahp, dt, dA);
      g *= an[t];
      hn = hp[t];
      // the two channels of the warp: lanes 0..15 collect dB, lanes 16..31 collect dC
      const float mine = half ? vC : vB, other = half ? vB : vC;
      wbuf[warp][t][lane] = mine + __shfl_xor_sync(0xffffffffu, other, 16);
    }
    // lane n < 8: du of token n; lane n >= 8: ddelta of token n - 8
    const float tot = halfwarp_sum16(red, lane);
    const float dy_tk = __shfl_sync(0xffffffffu, p0, half + 8 + tk);     // for the lanes 0..7
    const float sg_tk = __shfl_sync(0xffffffffu, p1, half + tk);         // for the lanes 8..15
    const float u_tk = __shfl_sync(0xffffffffu, p2, half + tk);
    if (chan_ok && tk < nt) {
      if (lo) {
        *po = from_f32<T>(fmaf(dy_tk, Dv, tot));
      } else {
        const float draw = tot * sg_tk;
        *po = from_f32<T>(draw);
        dBias += draw;
        dD = fmaf(p0, u_tk, dD);
        if (pz) *pz = from_f32<T>(p1 * fmaf(Dv, u_tk, ypre_tk));
      }
    }
    po -= so;
    if (pz) pz -= sz;
    __syncthreads();                                 // wbuf complete
    {
      const int t = tid >> 5, k = tid & 31;          // 8 tokens x 32 values
      if (t < nt) {
        float s = 0.f;
#pragma unroll
        for (int w = 0; w < kWarps; ++w) s += wbuf[w][t][k];
        *pslab = s;
      }
      pslab -= kT * 32;
    }
    // the next iteration's first __syncthreads orders these wbuf reads before its writes
  }
  if (valid) {
    if (a.dh0) a.dh0[((int64_t)b * Di + d) * N + n] = g;
    pA[((int64_t)b * Di + d) * N + n] = dA;
  }
  // dD / d(dt_bias): lanes 8..15 of a channel hold per-token partial sums
  dD += __shfl_xor_sync(0xffffffffu, dD, 1);  dBias += __shfl_xor_sync(0xffffffffu, dBias, 1);
  dD += __shfl_xor_sync(0xffffffffu, dD, 2);  dBias += __shfl_xor_sync(0xffffffffu, dBias, 2);
  dD += __shfl_xor_sync(0xffffffffu, dD, 4);  dBias += __shfl_xor_sync(0xffffffffu, dBias, 4);
  if (chan_ok && n == 8) {
    pD[(int64_t)b * Di + d] = dD;
    pBias[(int64_t)b * Di + d] = dBias;
  }
}

// dbc[b][t][b_off + n] = sum over slabs of dB, dbc[b][t][c_off + n] likewise for dC
template <typename T>
__global__ void scan_bwd_bc_kernel(const float* __restrict__ slabs, int nslabs, int64_t rows /* B*L */, int N,
                                   T* __restrict__ dbc, int64_t dbc_ts, int b_off, int c_off) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;   // (row, 32)
  if (i >= rows * 32) return;
  const int64_t row = i >> 5;
  const int k = (int)(i & 31);
  float s = 0.f;
  for (int p = 0; p < nslabs; ++p) s += slabs[((int64_t)p * rows + row) * 32 + k];
  const int n = k & 15;
  if (n < N) dbc[row * dbc_ts + (k < 16 ? b_off : c_off) + n] = from_f32<T>(s);
}

struct Plan {
  int nchunks, nslabs;
  int64_t ckpt, slabs, pA, pD, pBias, seg, total;
};
Plan plan(int B, int L, int Di, int N) {
  Plan p{};
  p.nchunks = (L + kT - 1) / kT;
  p.nslabs = (Di + kChan - 1) / kChan;
  auto up = [](int64_t v) { return (v + 255) / 256 * 256; };
  int64_t off = 0;
  // the larger of the two checkpoint layouts (every 8 tokens here, every 4 in scan_bwd_fast.cu)
  p.ckpt = off; off += up(std::max((int64_t)B * p.nchunks * N * Di * 4, scan_bwd_fast_ckpt_bytes(B, L, Di)));
  p.slabs = off; off += up((int64_t)p.nslabs * B * L * 32 * 4);
  const int64_t parts = (int64_t)B * (N == 16 ? scan_bwd_fast_segments(B, L, Di, nullptr) : 1);   // (segment, batch) rows
  p.pA = off; off += up(parts * Di * N * 4);
  p.pD = off; off += up(parts * Di * 4);
  p.pBias = off; off += up(parts * Di * 4);
  p.seg = off; off += up(N == 16 ? scan_bwd_fast_seg_bytes(B, L, Di) : 0);
  p.total = off;
  return p;
}

template <typename T, bool kAccurate>
int run(const vmb_scan_bwd_args& a, cudaStream_t st) {
  const Plan p = plan(a.B, a.L, a.Di, a.N);
  char* base = reinterpret_cast<char*>((reinterpret_cast<uintptr_t>(a.workspace) + 255) / 256 * 256);
  VMB_CHECK_ARG(a.workspace && a.workspace_bytes >= p.total + (base - (char*)a.workspace),
                "selective_scan_bwd: workspace too small");
  float* ckpt = reinterpret_cast<float*>(base + p.ckpt);
  float* slabs = reinterpret_cast<float*>(base + p.slabs);
  float* pA = reinterpret_cast<float*>(base + p.pA);
  float* pD = reinterpret_cast<float*>(base + p.pD);
  float* pBias = reinterpret_cast<float*>(base + p.pBias);
  int nparts = a.B;
  if (scan_bwd_fast_supported(a)) {
    float* seg_ws = scan_bwd_fast_seg_bytes(a.B, a.L, a.Di) > 0 ? reinterpret_cast<float*>(base + p.seg) : nullptr;
    const int rc = scan_bwd_fast(a, ckpt, slabs, pA, pD, pBias, seg_ws, &nparts, st);
    if (rc != VMB_OK) return rc;
  } else {
    dim3 grid((a.Di + kChan - 1) / kChan, a.B);
    scan_ckpt_kernel<T, kAccurate><<<grid, kWarps * 32, 0, st>>>(a, ckpt, p.nchunks);
    VMB_LAUNCH_CHECK("scan_ckpt_kernel");
    scan_bwd_kernel<T, kAccurate><<<grid, kWarps * 32, 0, st>>>(a, ckpt, p.nchunks, slabs, pA, pD, pBias);
    VMB_LAUNCH_CHECK("scan_bwd_kernel");
  }
  const int64_t rows = (int64_t)a.B * a.L;
  scan_bwd_bc_kernel<T><<<(unsigned)((rows * 32 + 255) / 256), 256, 0, st>>>(
      slabs, p.nslabs, rows, a.N, reinterpret_cast<T*>(a.dbc), a.dbc_tstride, a.b_off, a.c_off);
  VMB_LAUNCH_CHECK("scan_bwd_bc_kernel");
  int rc;
  if (a.dA && (rc = reduce_partials(pA, nparts, (int64_t)a.Di * a.N, a.dA, VMB_F32, st))) return rc;
  if (a.dD && (rc = reduce_partials(pD, nparts, a.Di, a.dD, VMB_F32, st))) return rc;
  if (a.ddt_bias && (rc = reduce_partials(pBias, nparts, a.Di, a.ddt_bias, VMB_F32, st))) return rc;
  return VMB_OK;
}

}  // namespace
}  // namespace vmb

using namespace vmb;

extern "C" int64_t vmb_scan_bwd_ckpt_bytes(int B, int L, int Di) {
  if (B <= 0 || L <= 0 || Di <= 0) return 0;
  return scan_bwd_fast_ckpt_bytes(B, L, Di);
}

extern "C" int64_t vmb_selective_scan_bwd_workspace_bytes(int B, int L, int Di, int N) {
  if (B <= 0 || L <= 0 || Di <= 0 || N <= 0) return 0;
  return plan(B, L, Di, N).total + 256;
}

extern "C" int vmb_selective_scan_bwd(const vmb_scan_bwd_args* a, vmb_stream_t stream) {
  VMB_CHECK_ARG(a != nullptr, "selective_scan_bwd: null args");
  VMB_CHECK_ARG(dtype_ok(a->dtype), "selective_scan_bwd: bad dtype %d", a->dtype);
  VMB_CHECK_ARG(a->B >= 0 && a->L >= 0 && a->Di > 0 && a->N > 0, "selective_scan_bwd: bad sizes");
  VMB_CHECK_ARG(a->B <= 65535, "selective_scan_bwd: batch %d > 65535", a->B);
  if (a->N > kNMax) VMB_UNSUPPORTED("selective_scan_bwd: d_state %d > %d not supported", a->N, kNMax);
  VMB_CHECK_ARG(!a->h0 || dtype_ok(a->h0_dtype), "selective_scan_bwd: bad h0 dtype");
  cudaStream_t st = as_stream(stream);
  if (a->B == 0 || a->L == 0) {
    if (a->dA) VMB_CUDA(cudaMemsetAsync(a->dA, 0, (size_t)a->Di * a->N * 4, st));
    if (a->dD) VMB_CUDA(cudaMemsetAsync(a->dD, 0, (size_t)a->Di * 4, st));
    if (a->ddt_bias) VMB_CUDA(cudaMemsetAsync(a->ddt_bias, 0, (size_t)a->Di * 4, st));
    if (a->B > 0 && a->dh0) {                       // empty sequence: h_last = h0
      if (a->dh_last) VMB_CUDA(cudaMemcpyAsync(a->dh0, a->dh_last, (size_t)a->B * a->Di * a->N * 4, cudaMemcpyDeviceToDevice, st));
      else VMB_CUDA(cudaMemsetAsync(a->dh0, 0, (size_t)a->B * a->Di * a->N * 4, st));
    }
    return VMB_OK;
  }
  VMB_CHECK_ARG(a->u && a->delta && a->bc && a->A2 && a->dout && a->du && a->ddelta && a->dbc,
                "selective_scan_bwd: null tensor");
  VMB_CHECK_ARG(a->z == nullptr || a->dz != nullptr, "selective_scan_bwd: z without dz");
  VMB_CHECK_ARG(a->dz_tstride == 0 || a->dz_tstride >= a->Di, "selective_scan_bwd: dz token stride < Di");
  if (a->dtype == VMB_F32) return run<float, true>(*a, st);
  return run<__nv_bfloat16, false>(*a, st);
}
