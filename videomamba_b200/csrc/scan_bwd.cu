// Backward of the selective scan.
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
// the state at the start of every 8-token chunk; pass 2 (scan_bwd_kernel) walks the chunks back to
// front, recomputes the 8 states of a chunk from its checkpoint into shared memory and runs the
// reverse recurrence on them.  One thread owns one (batch, channel) chain with all N <= 16 states in
// registers, so the reductions over n are in-thread; dB_t / dC_t need a sum over channels: a warp
// reduces its 32 channels with a transposing butterfly (31 shuffles for the 32 values of a token) and
// writes one 128-byte row per token into a per-warp slab, a last kernel sums the Di / 32 slabs.  dA, dD
// and d(dt_bias) are summed over the batch the same way.  No atomics: results are deterministic.
#include <algorithm>

#include "internal.h"

namespace vmb {
namespace {

constexpr int kT = 8;          // tokens per chunk
constexpr int kThr = 128;      // channels per CTA
constexpr int kNMax = 16;

template <typename T, bool kAccurate>
__global__ void __launch_bounds__(kThr)
scan_ckpt_kernel(const vmb_scan_bwd_args a, float* __restrict__ ckpt, int nchunks) {
  __shared__ float sB[kT][kNMax];
  const int tid = threadIdx.x;
  const int b = blockIdx.y;
  const int d = blockIdx.x * kThr + tid;
  const bool valid = d < a.Di;
  const int N = a.N, L = a.L;
  float A2[kNMax], h[kNMax];
#pragma unroll
  for (int n = 0; n < kNMax; ++n) {
    A2[n] = (valid && n < N) ? a.A2[(int64_t)d * N + n] : 0.f;
    h[n] = (valid && n < N && a.h0) ? load_as_f32(a.h0, ((int64_t)b * a.Di + d) * N + n, a.h0_dtype) : 0.f;
  }
  const float bias = (valid && a.dt_bias) ? a.dt_bias[d] : 0.f;
  const T* u = reinterpret_cast<const T*>(a.u) + (int64_t)b * a.u_bstride + d;
  const T* dl = reinterpret_cast<const T*>(a.delta) + (int64_t)b * a.d_bstride + d;
  const T* bc = reinterpret_cast<const T*>(a.bc) + (int64_t)b * a.bc_bstride;
  for (int c = 0; c < nchunks; ++c) {
    const int t0 = c * kT, nt = min(kT, L - t0);
    for (int e = tid; e < nt * N; e += kThr) sB[e / N][e % N] = to_f32<T>(bc[(int64_t)(t0 + e / N) * a.bc_tstride + a.b_off + e % N]);
    __syncthreads();
    if (valid) {
#pragma unroll
      for (int n = 0; n < kNMax; ++n)
        if (n < N) ckpt[(((int64_t)b * nchunks + c) * N + n) * a.Di + d] = h[n];
      for (int t = 0; t < nt; ++t) {
        const float uv = to_f32<T>(u[(int64_t)(t0 + t) * a.u_tstride]);
        float dv = to_f32<T>(dl[(int64_t)(t0 + t) * a.d_tstride]) + bias;
        if (a.softplus) dv = softplus_f<kAccurate>(dv);
        const float du = dv * uv;
#pragma unroll
        for (int n = 0; n < kNMax; ++n)
          if (n < N) h[n] = fmaf(exp2_f<kAccurate>(dv * A2[n]), h[n], du * sB[t][n]);
      }
    }
    __syncthreads();
  }
}

template <typename T, bool kAccurate>
__global__ void __launch_bounds__(kThr)
scan_bwd_kernel(const vmb_scan_bwd_args a, const float* __restrict__ ckpt, int nchunks,
                float* __restrict__ bc_slabs, float* __restrict__ pA, float* __restrict__ pD,
                float* __restrict__ pBias) {
  extern __shared__ float hist[];                 // [kT][kNMax][kThr]: h_{t-1} of the chunk's tokens
  __shared__ float sB[kT][kNMax];
  __shared__ float sC[kT][kNMax];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int b = blockIdx.y;
  const int d = blockIdx.x * kThr + tid;
  const bool valid = d < a.Di;
  const int N = a.N, L = a.L;
  const int64_t Di = a.Di;

  float A2[kNMax], g[kNMax], dA[kNMax];
#pragma unroll
  for (int n = 0; n < kNMax; ++n) {
    A2[n] = (valid && n < N) ? a.A2[(int64_t)d * N + n] : 0.f;
    g[n] = (valid && n < N && a.dh_last) ? a.dh_last[((int64_t)b * Di + d) * N + n] : 0.f;
    dA[n] = 0.f;
  }
  const float Dv = (valid && a.D) ? a.D[d] : 0.f;
  const float bias = (valid && a.dt_bias) ? a.dt_bias[d] : 0.f;
  float dD = 0.f, dBias = 0.f;
  const T* u = reinterpret_cast<const T*>(a.u) + (int64_t)b * a.u_bstride + d;
  const T* dl = reinterpret_cast<const T*>(a.delta) + (int64_t)b * a.d_bstride + d;
  const T* z = a.z ? reinterpret_cast<const T*>(a.z) + (int64_t)b * a.z_bstride + d : nullptr;
  const T* go = reinterpret_cast<const T*>(a.dout) + (int64_t)b * a.dout_bstride + d;
  const T* bc = reinterpret_cast<const T*>(a.bc) + (int64_t)b * a.bc_bstride;
  T* du_out = reinterpret_cast<T*>(a.du) + (int64_t)b * L * Di + d;
  T* dd_out = reinterpret_cast<T*>(a.ddelta) + (int64_t)b * L * Di + d;
  T* dz_out = a.dz ? reinterpret_cast<T*>(a.dz) + (int64_t)b * L * Di + d : nullptr;
  // slab of this warp: [(slab * B + b) * L + t][32] = {dB_t[0..15], dC_t[0..15]} summed over its 32 channels
  float* slab = bc_slabs + (((int64_t)(blockIdx.x * (kThr / 32) + warp) * a.B + b) * L) * 32;

  for (int c = nchunks - 1; c >= 0; --c) {
    const int t0 = c * kT, nt = min(kT, L - t0);
    for (int e = tid; e < nt * N; e += kThr) {
      const int t = e / N, n = e % N;
      sB[t][n] = to_f32<T>(bc[(int64_t)(t0 + t) * a.bc_tstride + a.b_off + n]);
      sC[t][n] = to_f32<T>(bc[(int64_t)(t0 + t) * a.bc_tstride + a.c_off + n]);
    }
    __syncthreads();
    float dlt[kT], sg[kT], ypre[kT];
    // ---- recompute the chunk forward from its checkpoint -------------------------------------------
    {
      float h[kNMax];
#pragma unroll
      for (int n = 0; n < kNMax; ++n)
        h[n] = (valid && n < N) ? ckpt[(((int64_t)b * nchunks + c) * N + n) * Di + d] : 0.f;
#pragma unroll
      for (int t = 0; t < kT; ++t) {
        dlt[t] = sg[t] = ypre[t] = 0.f;
        if (t < nt && valid) {
          const float uv = to_f32<T>(u[(int64_t)(t0 + t) * a.u_tstride]);
          const float raw = to_f32<T>(dl[(int64_t)(t0 + t) * a.d_tstride]) + bias;
          float dv = raw;
          sg[t] = 1.f;
          if (a.softplus) {
            dv = softplus_f<kAccurate>(raw);
            // d softplus = sigmoid; torch's softplus is the identity above its threshold (20)
            sg[t] = raw > 20.f ? 1.f : 1.f / (1.f + (kAccurate ? expf(-raw) : ex2_approx(-raw * kLog2e)));
          }
          dlt[t] = dv;
          const float duv = dv * uv;
          float acc = 0.f;
#pragma unroll
          for (int n = 0; n < kNMax; ++n) {
            if (n < N) {
              hist[(t * kNMax + n) * kThr + tid] = h[n];
              h[n] = fmaf(exp2_f<kAccurate>(dv * A2[n]), h[n], duv * sB[t][n]);
              acc = fmaf(h[n], sC[t][n], acc);
            }
          }
          ypre[t] = fmaf(Dv, uv, acc);
        }
      }
    }
    // ---- reverse recurrence over the chunk -----------------------------------------------------------
#pragma unroll
    for (int t = kT - 1; t >= 0; --t) {
      if (t < nt) {                                 // uniform over the CTA
        float v[32];                                // {dB contributions, dC contributions} of this thread
#pragma unroll
        for (int i = 0; i < 32; ++i) v[i] = 0.f;
        if (valid) {
          const int64_t row = t0 + t;
          const float uv = to_f32<T>(u[row * a.u_tstride]);
          const float gout = to_f32<T>(go[row * a.dout_tstride]);
          float dy = gout;
          if (z != nullptr) {
            const float zv = to_f32<T>(z[row * a.z_tstride]);
            const float s = 1.f / (1.f + (kAccurate ? expf(-zv) : ex2_approx(-zv * kLog2e)));
            dy = gout * zv * s;
            if (dz_out) dz_out[row * Di] = from_f32<T>(gout * ypre[t] * s * (1.f + zv * (1.f - s)));
          }
          const float dv = dlt[t];
          const float duv = dv * uv;
          float dut = dy * Dv, ddel = 0.f;
          dD = fmaf(dy, uv, dD);
#pragma unroll
          for (int n = 0; n < kNMax; ++n) {
            if (n < N) {
              const float hp = hist[(t * kNMax + n) * kThr + tid];
              const float an = exp2_f<kAccurate>(dv * A2[n]);
              const float bt = sB[t][n], ct = sC[t][n];
              const float hn = fmaf(an, hp, duv * bt);
              g[n] = fmaf(dy, ct, g[n]);
              v[16 + n] = dy * hn;
              v[n] = g[n] * duv;
              dut = fmaf(g[n] * dv, bt, dut);
              const float ahp = an * hp;
              ddel = fmaf(g[n], fmaf(ahp, A2[n] * kLn2, uv * bt), ddel);
              dA[n] = fmaf(g[n] * ahp, dv, dA[n]);
              g[n] *= an;
            }
          }
          const float draw = ddel * sg[t];
          dBias += draw;
          du_out[row * Di] = from_f32<T>(dut);
          dd_out[row * Di] = from_f32<T>(draw);
        }
        // transposing butterfly: afterwards lane i holds the warp's sum of v[i]
#pragma unroll
        for (int s = 16; s >= 1; s >>= 1) {
          const bool up = (lane & s) != 0;
#pragma unroll
          for (int k = 0; k < s; ++k) {
            const float keep = up ? v[k + s] : v[k];
            const float send = up ? v[k] : v[k + s];
            v[k] = keep + __shfl_xor_sync(0xffffffffu, send, s);
          }
        }
        slab[(int64_t)(t0 + t) * 32 + lane] = v[0];
      }
    }
    __syncthreads();
  }
  if (valid) {
#pragma unroll
    for (int n = 0; n < kNMax; ++n) {
      if (n < N) {
        if (a.dh0) a.dh0[((int64_t)b * Di + d) * N + n] = g[n];
        pA[((int64_t)b * Di + d) * N + n] = dA[n];
      }
    }
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
  int64_t ckpt, slabs, pA, pD, pBias, total;
};
Plan plan(int B, int L, int Di, int N) {
  Plan p{};
  p.nchunks = (L + kT - 1) / kT;
  p.nslabs = (Di + kThr - 1) / kThr * (kThr / 32);
  auto up = [](int64_t v) { return (v + 255) / 256 * 256; };
  int64_t off = 0;
  p.ckpt = off; off += up((int64_t)B * p.nchunks * N * Di * 4);
  p.slabs = off; off += up((int64_t)p.nslabs * B * L * 32 * 4);
  p.pA = off; off += up((int64_t)B * Di * N * 4);
  p.pD = off; off += up((int64_t)B * Di * 4);
  p.pBias = off; off += up((int64_t)B * Di * 4);
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
  dim3 grid((a.Di + kThr - 1) / kThr, a.B);
  scan_ckpt_kernel<T, kAccurate><<<grid, kThr, 0, st>>>(a, ckpt, p.nchunks);
  VMB_LAUNCH_CHECK("scan_ckpt_kernel");
  constexpr int smem = kT * kNMax * kThr * 4;
  static bool attr_set = false;
  if (!attr_set) {
    VMB_CUDA(cudaFuncSetAttribute(scan_bwd_kernel<T, kAccurate>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    attr_set = true;
  }
  scan_bwd_kernel<T, kAccurate><<<grid, kThr, smem, st>>>(a, ckpt, p.nchunks, slabs, pA, pD, pBias);
  VMB_LAUNCH_CHECK("scan_bwd_kernel");
  const int64_t rows = (int64_t)a.B * a.L;
  scan_bwd_bc_kernel<T><<<(unsigned)((rows * 32 + 255) / 256), 256, 0, st>>>(
      slabs, p.nslabs, rows, a.N, reinterpret_cast<T*>(a.dbc), a.dbc_tstride, a.b_off, a.c_off);
  VMB_LAUNCH_CHECK("scan_bwd_bc_kernel");
  int rc;
  if (a.dA && (rc = reduce_partials(pA, a.B, (int64_t)a.Di * a.N, a.dA, VMB_F32, st))) return rc;
  if (a.dD && (rc = reduce_partials(pD, a.B, a.Di, a.dD, VMB_F32, st))) return rc;
  if (a.ddt_bias && (rc = reduce_partials(pBias, a.B, a.Di, a.ddt_bias, VMB_F32, st))) return rc;
  return VMB_OK;
}

}  // namespace
}  // namespace vmb

using namespace vmb;

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
  if (a->dtype == VMB_F32) return run<float, true>(*a, st);
  return run<__nv_bfloat16, false>(*a, st);
}
