// Front glue of the token path (HBM-bound copies), as two kernels:
//   vmb_patchify      clip (B, C, T, H, W) -> patch rows (B*t*h*w, C*k*ph*pw): the im2col of the
//                     reference's Conv3d with kernel == stride (models/videomamba/videomamba.py:359-368,
//                     PatchEmbed), which turns the patch embedding into one dense projection.
//   vmb_embed_tokens  patch tokens (B, t, hw, D) + spatial pos (hw, D) + temporal pos (t, D), with the
//                     optional CLS row (cls_token + pos_embed[0]) written at position 0
//                     (videomamba.py:806-823): one read and one write of the token tensor instead
//                     of two adds and a torch.cat.
#include <algorithm>

#include "common.cuh"

namespace vmb {
namespace {

// One warp builds ONE patch row (C*k*ph*pw contiguous elements): lanes walk its VEC-element vectors, so
// the stores are fully coalesced runs and every output line is written once, completely, by one warp
// (the first kernel walked image rows: coalesced reads, but every 32-byte piece of a patch row arrived
// from a different warp at a different time -- 120 us for 308 MB).  The reads are pw-element segments
// (32 bytes at patch 16 / bf16) whose neighbours along x belong to the next patch, i.e. the next warp.
template <typename V, int VEC>
__global__ void __launch_bounds__(256)
patchify_kernel(const V* __restrict__ x, V* __restrict__ cols, int C, int T, int H, int W, int k,
                int ph, int pw, int t, int h, int w, int64_t rows) {
  const int64_t r = blockIdx.x * (int64_t)(blockDim.x >> 5) + (threadIdx.x >> 5);   // patch row (b, tp, py, px)
  if (r >= rows) return;
  int64_t q = r;
  const int px = (int)(q % w); q /= w;
  const int py = (int)(q % h); q /= h;
  const int tp = (int)(q % t);
  const int64_t b = q / t;
  const int vpw = pw / VEC;                        // vectors per pw-element segment
  const int nseg = C * k * ph;                     // segments (c, dt, dy) of the patch row
  const int nvec = nseg * vpw;
  const int64_t dst = r * (int64_t)nvec;
  for (int i = threadIdx.x & 31; i < nvec; i += 32) {
    const int sgm = i / vpw, dx = (i % vpw) * VEC;
    const int dy = sgm % ph;
    const int dt = (sgm / ph) % k;
    const int c = sgm / (ph * k);
    const int64_t src = ((((b * C + c) * T + (tp * k + dt)) * H + (py * ph + dy)) * (int64_t)W + px * pw + dx);
    cols[dst + i] = x[src / VEC];
  }
}

// tokens out[b][cls + (tp * hw + s)][:] = patches[b][tp][s][:] + spatial[s][:] + temporal[tp][:];
// out[b][0][:] = cls_row[:] when has_cls.  One thread per 8 (bf16) / 4 (fp32) channels.
template <typename T, int VEC>
__global__ void __launch_bounds__(256)
embed_tokens_kernel(const T* __restrict__ patches, const T* __restrict__ spatial,
                    const T* __restrict__ temporal, const T* __restrict__ cls_row, T* __restrict__ out,
                    int64_t B, int t, int hw, int D, int has_cls) {
  const int dv = D / VEC;
  const int64_t L = (int64_t)has_cls + (int64_t)t * hw;
  const int64_t total = B * L * dv;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total;
       i += (int64_t)gridDim.x * blockDim.x) {
    const int d = (int)(i % dv) * VEC;
    const int64_t tokrow = i / dv;
    const int64_t b = tokrow / L;
    const int64_t l = tokrow % L;
    // whole-vector loads / stores (VEC elements = 16 bytes on the aligned path, 1 element otherwise)
    struct alignas(sizeof(T) * VEC) Vec { T e[VEC]; };
    Vec o;
    if (has_cls && l == 0) {
      o = *reinterpret_cast<const Vec*>(cls_row + d);
    } else {
      const int64_t p = l - has_cls;
      const int tp = (int)(p / hw), s = (int)(p % hw);
      const Vec a = *reinterpret_cast<const Vec*>(patches + ((b * t + tp) * hw + s) * D + d);
      const Vec sp = *reinterpret_cast<const Vec*>(spatial + (int64_t)s * D + d);
      const Vec te = *reinterpret_cast<const Vec*>(temporal + (int64_t)tp * D + d);
#pragma unroll
      for (int e = 0; e < VEC; ++e) {
        // two roundings, as the reference's two separate adds in the model dtype
        const float first = to_f32<T>(from_f32<T>(to_f32<T>(a.e[e]) + to_f32<T>(sp.e[e])));
        o.e[e] = from_f32<T>(first + to_f32<T>(te.e[e]));
      }
    }
    *reinterpret_cast<Vec*>(out + tokrow * D + d) = o;
  }
}

}  // namespace
}  // namespace vmb

using namespace vmb;

extern "C" int vmb_patchify(const void* x, void* cols, int64_t B, int C, int T, int H, int W, int k,
                            int ph, int pw, int dtype, vmb_stream_t stream) {
  VMB_CHECK_ARG(dtype_ok(dtype), "patchify: bad dtype %d", dtype);
  VMB_CHECK_ARG(B >= 0 && C > 0 && T > 0 && H > 0 && W > 0 && k > 0 && ph > 0 && pw > 0,
                "patchify: bad sizes");
  const int t = T / k, h = H / ph, w = W / pw;
  if (B == 0 || t == 0 || h == 0 || w == 0) return VMB_OK;
  VMB_CHECK_ARG(x && cols, "patchify: null pointer");
  cudaStream_t st = as_stream(stream);
  const int es = dtype_size(dtype);
  auto grid_for = [](int64_t rows) { return (unsigned)((rows + 7) / 8); };
  // widest vector that divides the patch width and keeps every source / destination run aligned
  int vec_bytes = 16;
  while (vec_bytes > es && ((pw * es) % vec_bytes != 0 || (W * es) % vec_bytes != 0 ||
                            reinterpret_cast<uintptr_t>(x) % vec_bytes != 0 ||
                            reinterpret_cast<uintptr_t>(cols) % vec_bytes != 0))
    vec_bytes /= 2;
  const int64_t nv = B * (int64_t)t * h * w;                  // patch rows, one warp each
  // the element type only matters through its size: move raw words
  if (vec_bytes == 16) {
    if (es == 2) patchify_kernel<uint4, 8><<<grid_for(nv), 256, 0, st>>>((const uint4*)x, (uint4*)cols, C, T, H, W, k, ph, pw, t, h, w, nv);
    else patchify_kernel<uint4, 4><<<grid_for(nv), 256, 0, st>>>((const uint4*)x, (uint4*)cols, C, T, H, W, k, ph, pw, t, h, w, nv);
  } else if (vec_bytes == 8) {
    if (es == 2) patchify_kernel<uint2, 4><<<grid_for(nv), 256, 0, st>>>((const uint2*)x, (uint2*)cols, C, T, H, W, k, ph, pw, t, h, w, nv);
    else patchify_kernel<uint2, 2><<<grid_for(nv), 256, 0, st>>>((const uint2*)x, (uint2*)cols, C, T, H, W, k, ph, pw, t, h, w, nv);
  } else if (vec_bytes == 4) {
    if (es == 2) patchify_kernel<uint32_t, 2><<<grid_for(nv), 256, 0, st>>>((const uint32_t*)x, (uint32_t*)cols, C, T, H, W, k, ph, pw, t, h, w, nv);
    else patchify_kernel<uint32_t, 1><<<grid_for(nv), 256, 0, st>>>((const uint32_t*)x, (uint32_t*)cols, C, T, H, W, k, ph, pw, t, h, w, nv);
  } else {
    patchify_kernel<uint16_t, 1><<<grid_for(nv), 256, 0, st>>>((const uint16_t*)x, (uint16_t*)cols, C, T, H, W, k, ph, pw, t, h, w, nv);
  }
  VMB_LAUNCH_CHECK("patchify_kernel");
  return VMB_OK;
}

extern "C" int vmb_embed_tokens(const void* patches, const void* spatial, const void* temporal,
                                const void* cls_row, void* out, int64_t B, int t, int hw, int D,
                                int dtype, vmb_stream_t stream) {
  VMB_CHECK_ARG(dtype_ok(dtype), "embed_tokens: bad dtype %d", dtype);
  VMB_CHECK_ARG(B >= 0 && t >= 0 && hw >= 0 && D > 0, "embed_tokens: bad sizes");
  const int has_cls = cls_row != nullptr;
  const int64_t L = has_cls + (int64_t)t * hw;
  if (B == 0 || L == 0) return VMB_OK;
  VMB_CHECK_ARG(out && (t * (int64_t)hw == 0 || (patches && spatial && temporal)), "embed_tokens: null pointer");
  cudaStream_t st = as_stream(stream);
  auto al16 = [](const void* p) { return reinterpret_cast<uintptr_t>(p) % 16 == 0; };
  const int vec = dtype == VMB_BF16 ? 8 : 4;
  const bool v = D % vec == 0 && al16(patches) && al16(spatial) && al16(temporal) && al16(out) &&
                 (!has_cls || al16(cls_row));
  const int64_t n = B * L * (v ? D / vec : D);
  const unsigned grid = (unsigned)std::min<int64_t>((n + 255) / 256, 148 * 16);
  if (dtype == VMB_BF16) {
    using T = __nv_bfloat16;
    if (v) embed_tokens_kernel<T, 8><<<grid, 256, 0, st>>>((const T*)patches, (const T*)spatial, (const T*)temporal, (const T*)cls_row, (T*)out, B, t, hw, D, has_cls);
    else embed_tokens_kernel<T, 1><<<grid, 256, 0, st>>>((const T*)patches, (const T*)spatial, (const T*)temporal, (const T*)cls_row, (T*)out, B, t, hw, D, has_cls);
  } else {
    if (v) embed_tokens_kernel<float, 4><<<grid, 256, 0, st>>>((const float*)patches, (const float*)spatial, (const float*)temporal, (const float*)cls_row, (float*)out, B, t, hw, D, has_cls);
    else embed_tokens_kernel<float, 1><<<grid, 256, 0, st>>>((const float*)patches, (const float*)spatial, (const float*)temporal, (const float*)cls_row, (float*)out, B, t, hw, D, has_cls);
  }
  VMB_LAUNCH_CHECK("embed_tokens_kernel");
  return VMB_OK;
}
