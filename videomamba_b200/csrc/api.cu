// C ABI of libvmb200 (see include/vmb200.h): argument checking, kernel selection, and the
// whole-mixer orchestration that stands where the body of the reference's Mamba.forward
// (models/videomamba/mamba_simple.py:332-446) issues its operator calls.
#include <cstdarg>
#include <cstdio>
#include <mutex>
#include <vector>

#include "internal.h"

namespace vmb {

static thread_local char g_err[512] = "";

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

int cuda_fail(cudaError_t e, const char* what) {
  set_error("CUDA error in %s: %s (%s)", what, cudaGetErrorName(e), cudaGetErrorString(e));
  return VMB_ERR_CUDA;
}

int sm_count() {
  static int cached[64] = {0};
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return 148;
  if (cached[dev] == 0) {
    int n = 0;
    if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0)
      n = 148;
    cached[dev] = n;
  }
  return cached[dev];
}

// ---- per-stage timing ------------------------------------------------------------------------
bool g_prof_on = false;
unsigned long long g_launches = 0;
namespace {
struct ProfRec { int kind; cudaEvent_t a, b; bool closed; };
std::mutex g_prof_mu;
std::vector<ProfRec> g_prof_log;
std::vector<cudaEvent_t> g_prof_free;
cudaEvent_t prof_event() {
  if (!g_prof_free.empty()) { cudaEvent_t e = g_prof_free.back(); g_prof_free.pop_back(); return e; }
  cudaEvent_t e = nullptr;
  cudaEventCreate(&e);
  return e;
}
}  // namespace
void prof_begin(int kind, cudaStream_t st, int* slot) {
  std::lock_guard<std::mutex> lk(g_prof_mu);
  ProfRec r{kind, prof_event(), prof_event(), false};
  if (!r.a || !r.b || cudaEventRecord(r.a, st) != cudaSuccess) { (void)cudaGetLastError(); return; }
  g_prof_log.push_back(r);
  *slot = (int)g_prof_log.size() - 1;
}
void prof_end(int slot, cudaStream_t st) {
  std::lock_guard<std::mutex> lk(g_prof_mu);
  if (slot < 0 || slot >= (int)g_prof_log.size()) return;
  if (cudaEventRecord(g_prof_log[slot].b, st) == cudaSuccess) g_prof_log[slot].closed = true;
  else (void)cudaGetLastError();
}

namespace {

inline int64_t align_up(int64_t v, int64_t a) { return (v + a - 1) / a * a; }
inline int xdbl_pitch(int R, int N) { return (int)align_up(R + 2 * N, 16); }

struct MixerWorkspace {
  int64_t xz, xc, xdbl, delta, y, seg, seg_bytes, total;  // byte offsets
};

MixerWorkspace plan_workspace(int B, int L, int Di, int N, int R, int dtype) {
  const int64_t es = dtype_size(dtype);
  const int64_t M = (int64_t)B * L;
  MixerWorkspace w;
  int64_t off = 0;
  w.xz = off;    off = align_up(off + M * 2 * Di * es, 256);
  w.xc = off;    off = align_up(off + M * Di * es, 256);
  w.xdbl = off;  off = align_up(off + M * xdbl_pitch(R, N) * es, 256);
  w.delta = off; off = align_up(off + M * Di * es, 256);
  w.y = off;     off = align_up(off + M * Di * es, 256);
  w.seg_bytes = dtype == VMB_BF16 ? scan_fast_workspace_bytes(B, L, Di, N) : 0;
  w.seg = off;   off = align_up(off + w.seg_bytes, 256);
  w.total = off;
  return w;
}

template <typename V>
__global__ void state_rows_kernel(const V* __restrict__ src, V* __restrict__ dst,
                                  const int32_t* __restrict__ index, int64_t row_vecs,
                                  bool gather) {
  const int r = blockIdx.y;
  const int64_t p = index[r];
  const V* s = gather ? src + p * row_vecs : src + (int64_t)r * row_vecs;
  V* d = gather ? dst + (int64_t)r * row_vecs : dst + p * row_vecs;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < row_vecs;
       i += (int64_t)gridDim.x * blockDim.x)
    d[i] = s[i];
}

int state_rows(const void* src, void* dst, const int32_t* index, int n_rows, int64_t row_elems,
               int dtype, bool gather, cudaStream_t st) {
  VMB_CHECK_ARG(dtype_ok(dtype), "state_gather/scatter: bad dtype");
  VMB_CHECK_ARG(n_rows >= 0 && n_rows <= 65535 && row_elems >= 0, "state_gather/scatter: bad sizes");
  if (n_rows == 0 || row_elems == 0) return VMB_OK;
  VMB_CHECK_ARG(src && dst && index, "state_gather/scatter: null pointer");
  const int64_t bytes = row_elems * dtype_size(dtype);
  const bool v16 = bytes % 16 == 0 && reinterpret_cast<uintptr_t>(src) % 16 == 0 &&
                   reinterpret_cast<uintptr_t>(dst) % 16 == 0;
  if (v16) {
    const int64_t nv = bytes / 16;
    dim3 grid((unsigned)std::min<int64_t>((nv + 255) / 256, 64), n_rows);
    state_rows_kernel<uint4><<<grid, 256, 0, st>>>((const uint4*)src, (uint4*)dst, index, nv,
                                                   gather);
  } else if (dtype == VMB_BF16) {
    dim3 grid((unsigned)std::min<int64_t>((row_elems + 255) / 256, 64), n_rows);
    state_rows_kernel<uint16_t><<<grid, 256, 0, st>>>((const uint16_t*)src, (uint16_t*)dst, index,
                                                      row_elems, gather);
  } else {
    dim3 grid((unsigned)std::min<int64_t>((row_elems + 255) / 256, 64), n_rows);
    state_rows_kernel<uint32_t><<<grid, 256, 0, st>>>((const uint32_t*)src, (uint32_t*)dst, index,
                                                      row_elems, gather);
  }
  VMB_LAUNCH_CHECK("state_rows_kernel");
  return VMB_OK;
}

}  // namespace
}  // namespace vmb

using namespace vmb;

extern "C" int vmb_abi_version(void) { return VMB_ABI_VERSION; }
extern "C" const char* vmb_last_error(void) { return g_err; }

extern "C" int vmb_device_info(int* sms, int* cc_major, int* cc_minor) {
  int dev = 0;
  VMB_CUDA(cudaGetDevice(&dev));
  int n = 0, ma = 0, mi = 0;
  VMB_CUDA(cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev));
  VMB_CUDA(cudaDeviceGetAttribute(&ma, cudaDevAttrComputeCapabilityMajor, dev));
  VMB_CUDA(cudaDeviceGetAttribute(&mi, cudaDevAttrComputeCapabilityMinor, dev));
  if (sms) *sms = n;
  if (cc_major) *cc_major = ma;
  if (cc_minor) *cc_minor = mi;
  return VMB_OK;
}

extern "C" int vmb_linear_fwd(const void* A, int64_t lda, const void* W, int64_t ldw,
                              const void* bias, void* C, int64_t ldc, int64_t M, int N, int K,
                              int dtype, vmb_stream_t stream) {
  VMB_CHECK_ARG(dtype_ok(dtype), "linear: bad dtype %d", dtype);
  VMB_CHECK_ARG(M >= 0 && N > 0 && K > 0, "linear: bad sizes M=%lld N=%d K=%d", (long long)M, N, K);
  VMB_CHECK_ARG(M == 0 || (A && W && C), "linear: null A / W / C");
  VMB_CHECK_ARG(lda >= K && ldw >= K && ldc >= N, "linear: row stride smaller than row");
  if (M == 0) return VMB_OK;
  cudaStream_t st = as_stream(stream);
  if (dtype == VMB_BF16 && gemm_tc_supported(A, lda, W, ldw, C, ldc, M, N, K))
    return gemm_tc(A, lda, W, ldw, bias, C, ldc, M, N, K, st);
  return linear_simt(A, lda, W, ldw, bias, C, ldc, M, N, K, dtype, st);
}

extern "C" int vmb_linear_fwd_act(const void* A, int64_t lda, const void* W, int64_t ldw,
                                  const void* bias, void* C, int64_t ldc, int64_t M, int N, int K,
                                  int dtype, int act_from, vmb_stream_t stream) {
  VMB_CHECK_ARG(dtype_ok(dtype), "linear_act: bad dtype %d", dtype);
  VMB_CHECK_ARG(M >= 0 && N > 0 && K > 0, "linear_act: bad sizes M=%lld N=%d K=%d", (long long)M, N, K);
  VMB_CHECK_ARG(M == 0 || (A && W && C), "linear_act: null A / W / C");
  VMB_CHECK_ARG(lda >= K && ldw >= K && ldc >= N, "linear_act: row stride smaller than row");
  VMB_CHECK_ARG(act_from >= 0 && act_from <= N, "linear_act: act_from=%d outside [0, %d]", act_from, N);
  if (act_from == N) return vmb_linear_fwd(A, lda, W, ldw, bias, C, ldc, M, N, K, dtype, stream);
  if (dtype != VMB_BF16 || act_from % 64 != 0 || !gemm_tc_supported(A, lda, W, ldw, C, ldc, M, N, K))
    VMB_UNSUPPORTED("linear_act: the activation epilogue exists on the tensor-core kernel only (bf16, TMA-aligned "
                    "operands, act_from a multiple of 64)");
  if (M == 0) return VMB_OK;
  return gemm_tc(A, lda, W, ldw, bias, C, ldc, M, N, K, as_stream(stream), act_from);
}

extern "C" int vmb_selective_scan_fwd(const vmb_scan_args* a, vmb_stream_t stream) {
  VMB_CHECK_ARG(a != nullptr, "selective_scan: null args");
  VMB_CHECK_ARG(a->u && a->delta && a->bc && a->A2 && a->y, "selective_scan: null tensor");
  VMB_CHECK_ARG(dtype_ok(a->dtype), "selective_scan: bad dtype %d", a->dtype);
  VMB_CHECK_ARG(a->B >= 0 && a->L >= 0 && a->Di > 0 && a->N > 0, "selective_scan: bad sizes");
  VMB_CHECK_ARG(a->B <= 65535, "selective_scan: batch %d > 65535", a->B);
  VMB_CHECK_ARG(!a->h0 || dtype_ok(a->h0_dtype), "selective_scan: bad h0 dtype");
  VMB_CHECK_ARG(a->frame_len >= 0 && (a->frame_len == 0 || a->L % a->frame_len == 0),
                "selective_scan: L=%d is not a whole number of frames of %d tokens", a->L, a->frame_len);
  if (a->B == 0) return VMB_OK;
  return scan_generic(*a, as_stream(stream));
}

extern "C" int64_t vmb_fused_scan_workspace_bytes(int B, int L, int Di, int N) {
  if (B <= 0 || L <= 0 || Di <= 0 || N <= 0) return 0;
  return scan_fast_workspace_bytes(B, L, Di, N);
}

extern "C" int vmb_selective_scan_fused_fwd(const vmb_fused_scan_args* a, vmb_stream_t stream) {
  VMB_CHECK_ARG(a != nullptr, "fused_scan: null args");
  VMB_CHECK_ARG(a->B >= 0 && a->L >= 0 && a->Di > 0 && a->N > 0 && a->R > 0, "fused_scan: bad sizes");
  if (a->B == 0 || a->L == 0) {
    if (a->L == 0 && a->B > 0 && a->h_last) VMB_UNSUPPORTED("fused_scan: empty sequence with state output");
    return VMB_OK;
  }
  VMB_CHECK_ARG(a->u && a->z && a->xdbl && a->w_dt && a->A2 && a->y, "fused_scan: null tensor");
  VMB_CHECK_ARG(!a->h0 || dtype_ok(a->h0_dtype), "fused_scan: bad h0 dtype");
  FastScanArgs f;
  f.u = a->u; f.u_bs = a->u_bstride; f.u_ts = a->u_tstride;
  f.z = a->z; f.z_bs = a->z_bstride; f.z_ts = a->z_tstride;
  f.xdbl = a->xdbl; f.x_bs = a->x_bstride; f.x_ts = a->x_tstride;
  f.w_dt_pad = a->w_dt; f.A2 = a->A2; f.D = a->D; f.dt_bias = a->dt_bias;
  f.h0 = a->h0; f.h0_dtype = a->h0_dtype;
  f.y = a->y; f.y_bs = a->y_bstride; f.y_ts = a->y_tstride; f.h_last = a->h_last;
  f.B = a->B; f.L = a->L; f.Di = a->Di; f.N = a->N; f.R = a->R; f.Rp = a->Rp; f.Xp = a->Xp;
  f.reverse = a->reverse; f.a_geometric = a->a_geometric; f.tune = a->tune;
  f.frame_len = a->reverse ? a->frame_len : 0;
  VMB_CHECK_ARG(f.frame_len >= 0 && (f.frame_len == 0 || a->L % f.frame_len == 0),
                "fused_scan: L=%d is not a whole number of frames of %d tokens", a->L, f.frame_len);
  VMB_CHECK_ARG(a->bwd_ckpt == nullptr || (!a->reverse && reinterpret_cast<uintptr_t>(a->bwd_ckpt) % 16 == 0),
                "fused_scan: bwd_ckpt needs the forward walk and 16-byte alignment");
  f.ckpt = a->bwd_ckpt;
  f.z_gate = a->z_gate != 0;
  VMB_CHECK_ARG(!(f.z_gate && f.ckpt), "fused_scan: the training forward (bwd_ckpt) needs the raw z, not the gate");
  f.seg_ws = reinterpret_cast<float*>(a->workspace);
  f.seg_ws_bytes = a->workspace ? a->workspace_bytes : 0;
  VMB_CHECK_ARG(reinterpret_cast<uintptr_t>(a->workspace) % 16 == 0, "fused_scan: workspace not 16-byte aligned");
  if (!scan_fast_supported(f)) VMB_UNSUPPORTED("fused_scan: shape / alignment not covered by the fused kernel");
  ProfScope ps(VMB_PROF_SCAN, as_stream(stream));
  return scan_fast(f, as_stream(stream));
}

extern "C" int64_t vmb_mixer_workspace_bytes(int B, int L, int D, int Di, int N, int R,
                                             int dtype) {
  (void)D;
  if (B < 0 || L < 0 || Di <= 0 || N <= 0 || R <= 0 || !dtype_ok(dtype)) return -1;
  return plan_workspace(B, L, Di, N, R, dtype).total + 256;
}

extern "C" int vmb_mixer_fwd(const vmb_mixer_args* p, vmb_stream_t stream) {
  VMB_CHECK_ARG(p != nullptr, "mixer: null args");
  VMB_CHECK_ARG(p->hidden && p->out && p->w_in && p->w_conv && p->w_x && p->w_dt && p->w_out &&
                    p->A2 && p->Dskip && p->dt_bias,
                "mixer: null tensor");
  VMB_CHECK_ARG(dtype_ok(p->dtype), "mixer: bad dtype %d", p->dtype);
  VMB_CHECK_ARG(p->B >= 0 && p->L >= 0 && p->D > 0 && p->Di > 0 && p->N > 0 && p->R > 0 &&
                    p->W > 0,
                "mixer: bad sizes");
  if (p->B == 0 || p->L == 0) {
    if (p->L == 0 && p->B > 0 && (p->conv_state_out || p->ssm_state_out))
      VMB_UNSUPPORTED("mixer: empty sequence with state output");
    return VMB_OK;
  }
  const int B = p->B, L = p->L, D = p->D, Di = p->Di, N = p->N, R = p->R;
  const int X = R + 2 * N, Xw = xdbl_pitch(R, N);
  const int64_t M = (int64_t)B * L;
  const MixerWorkspace ws = plan_workspace(B, L, Di, N, R, p->dtype);
  VMB_CHECK_ARG(p->workspace != nullptr, "mixer: null workspace");
  char* base = reinterpret_cast<char*>(align_up(reinterpret_cast<int64_t>(p->workspace), 256));
  VMB_CHECK_ARG(p->workspace_bytes >= ws.total + (base - (char*)p->workspace),
                "mixer: workspace too small (%lld < %lld)", (long long)p->workspace_bytes,
                (long long)ws.total);
  VMB_CHECK_ARG(p->h_bstride == (int64_t)L * p->h_tstride && p->o_bstride == (int64_t)L * p->o_tstride,
                "mixer: hidden / out must be uniformly strided over (batch, token)");
  void* xz = base + ws.xz;
  void* xc = base + ws.xc;
  void* xdbl = base + ws.xdbl;
  void* delta = base + ws.delta;
  void* y = base + ws.y;
  const int64_t es = dtype_size(p->dtype);
  cudaStream_t st = as_stream(stream);
  int rc;

  FastScanArgs f;
  f.u = xc; f.u_bs = (int64_t)L * Di; f.u_ts = Di;
  f.z = (char*)xz + (int64_t)Di * es; f.z_bs = (int64_t)L * 2 * Di; f.z_ts = 2 * Di;
  f.xdbl = xdbl; f.x_bs = (int64_t)L * Xw; f.x_ts = Xw;
  f.w_dt_pad = p->w_dt_pad; f.A2 = p->A2; f.D = p->Dskip; f.dt_bias = p->dt_bias;
  f.h0 = p->ssm_state_in; f.h0_dtype = p->ss_in_dtype;
  f.y = y; f.y_bs = (int64_t)L * Di; f.y_ts = Di; f.h_last = p->ssm_state_out;
  f.B = B; f.L = L; f.Di = Di; f.N = N; f.R = R; f.Rp = p->Rp; f.Xp = p->Xp; f.reverse = p->reverse;
  f.a_geometric = p->a_geometric; f.tune = p->scan_tune;
  f.frame_len = p->reverse ? p->frame_len : 0;
  VMB_CHECK_ARG(f.frame_len >= 0 && (f.frame_len == 0 || L % f.frame_len == 0),
                "mixer: L=%d is not a whole number of frames of %d tokens", L, f.frame_len);
  f.seg_ws = ws.seg_bytes ? reinterpret_cast<float*>(base + ws.seg) : nullptr;
  f.seg_ws_bytes = ws.seg_bytes;
  const bool fast_ok = p->dtype == VMB_BF16 && p->w_x_pad && p->w_dt_pad && p->Xp == Xw &&
                       scan_fast_supported(f);
  if (p->path == 2 && !fast_ok) VMB_UNSUPPORTED("mixer: fast path requested but not available");
  const bool fast = fast_ok && p->path != 1;

  // in_proj (mamba_simple.py:333-339): xz (M, 2Di), x = [:, :Di], z = [:, Di:].  On request (gate_in_proj) and
  // on the fast path the epilogue of the tensor-core projection applies SiLU to the z half before its one
  // rounding to bf16, and the scan multiplies by the stored gate as it is (mamba_simple.py:423-435 applies
  // SiLU to the rounded z inside the scan: same number of roundings, at the other side of the activation).
  const bool gate = p->gate_in_proj != 0 && fast && Di % 64 == 0 &&
                    gemm_tc_supported(p->hidden, p->h_tstride, p->w_in, D, xz, 2 * Di, M, 2 * Di, D);
  {
    ProfScope ps(VMB_PROF_IN_PROJ, st);
    rc = gate ? vmb_linear_fwd_act(p->hidden, p->h_tstride, p->w_in, D, p->b_in, xz, 2 * Di, M, 2 * Di, D,
                                   p->dtype, Di, stream)
              : vmb_linear_fwd(p->hidden, p->h_tstride, p->w_in, D, p->b_in, xz, 2 * Di, M, 2 * Di, D,
                               p->dtype, stream);
  }
  if (rc) return rc;
  f.z_gate = gate ? 1 : 0;

  // Stateless forward walk on the fast path, on request (fuse_conv_xproj): the conv runs inside the x_proj
  // projection (its output tile goes from the conv warps to the tensor cores through shared memory and to
  // HBM once).  Alone it saves 14 us per layer, but it is a persistent kernel that fills the SM's shared
  // memory, so with several forwards in flight it cannot share SMs with another step's scan the way the
  // small conv CTAs do (bench.py: serial 20.6 -> 20.2 ms, 3 in flight 17.6 -> 18.5 ms): the caller decides.
  const bool fuse_conv_xproj = p->fuse_conv_xproj != 0;
  const bool fused_conv = fuse_conv_xproj && fast && p->W == 4 && !p->reverse && p->conv_state_in == nullptr &&
                          p->conv_state_out == nullptr &&
                          conv_xproj_supported(xz, 2 * Di, p->w_conv, p->b_conv, p->w_x_pad, Di, xc, Di,
                                               xdbl, Xw, M, Xw, Di, L);
  if (fused_conv) {
    ProfScope ps(VMB_PROF_CONV, st);
    rc = conv_xproj_tc(xz, 2 * Di, p->w_conv, p->b_conv, p->w_x_pad, Di, xc, Di, xdbl, Xw, M, Xw, Di, L,
                       1, st);
    if (rc) return rc;
  } else {
    // causal conv + SiLU with optional history (mamba_simple.py:381-404)
    ProfScope ps(VMB_PROF_CONV, st);
    rc = vmb_causal_conv1d_fwd(xz, (int64_t)L * 2 * Di, 2 * Di, p->w_conv, p->b_conv,
                               p->conv_state_in, p->cs_in_dtype, xc, (int64_t)L * Di, Di,
                               p->conv_state_out, p->cs_out_dtype, B, L, Di, p->W, 1, p->reverse,
                               f.frame_len, p->dtype, stream);
    if (rc) return rc;
  }

  if (fast) {
    // x_proj with zero-padded weight rows: x_dbl (M, Xp) = [dt_low | B | C | 0]
    if (!fused_conv) {
      ProfScope ps(VMB_PROF_X_PROJ, st);
      rc = vmb_linear_fwd(xc, Di, p->w_x_pad, Di, nullptr, xdbl, Xw, M, Xw, Di, p->dtype, stream);
      if (rc) return rc;
    }
    {
      ProfScope ps(VMB_PROF_SCAN, st);
      rc = scan_fast(f, st);
    }
    if (rc) return rc;
  } else {
    // x_proj (mamba_simple.py:409) and dt_proj (:413-414): delta_raw rounded to the model dtype
    {
      ProfScope ps(VMB_PROF_X_PROJ, st);
      rc = vmb_linear_fwd(xc, Di, p->w_x, Di, nullptr, xdbl, Xw, M, X, Di, p->dtype, stream);
    }
    if (rc) return rc;
    {
      ProfScope ps(VMB_PROF_DT_PROJ, st);
      rc = vmb_linear_fwd(xdbl, Xw, p->w_dt, R, nullptr, delta, Di, M, Di, R, p->dtype, stream);
    }
    if (rc) return rc;
    vmb_scan_args s;
    s.u = xc; s.u_bstride = (int64_t)L * Di; s.u_tstride = Di;
    s.delta = delta; s.d_bstride = (int64_t)L * Di; s.d_tstride = Di;
    s.z = f.z; s.z_bstride = f.z_bs; s.z_tstride = f.z_ts;
    s.bc = xdbl; s.bc_bstride = (int64_t)L * Xw; s.bc_tstride = Xw; s.b_off = R; s.c_off = R + N;
    s.A2 = p->A2; s.D = p->Dskip; s.dt_bias = p->dt_bias;
    s.h0 = p->ssm_state_in; s.h0_dtype = p->ss_in_dtype;
    s.y = y; s.y_bstride = (int64_t)L * Di; s.y_tstride = Di; s.h_last = p->ssm_state_out;
    s.B = B; s.L = L; s.Di = Di; s.N = N; s.dtype = p->dtype; s.softplus = 1;
    s.reverse = p->reverse; s.frame_len = f.frame_len;
    {
      ProfScope ps(VMB_PROF_SCAN, st);
      rc = vmb_selective_scan_fwd(&s, stream);
    }
    if (rc) return rc;
  }
  ProfScope ps_out(VMB_PROF_OUT_PROJ, st);
  // out_proj (mamba_simple.py:445-446)
  return vmb_linear_fwd(y, Di, p->w_out, Di, p->b_out, p->out, p->o_tstride, M, D, Di, p->dtype,
                        stream);
}

extern "C" int vmb_conv_xproj_fwd(const void* x, int64_t x_ld, const void* w_conv, const void* b_conv,
                                  const void* w_x, int64_t w_x_ld, void* xc, int64_t xc_ld, void* x_dbl,
                                  int64_t x_dbl_ld, int64_t M, int N, int Di, int L, vmb_stream_t stream) {
  using namespace vmb;
  VMB_CHECK_ARG(M >= 0 && N > 0 && Di > 0 && L > 0, "conv_xproj: bad sizes");
  if (M == 0) return VMB_OK;
  VMB_CHECK_ARG(x && w_conv && w_x && xc && x_dbl, "conv_xproj: null pointer");
  VMB_CHECK_ARG(M % L == 0, "conv_xproj: M must be a whole number of sequences of L tokens");
  if (!conv_xproj_supported(x, x_ld, w_conv, b_conv, w_x, w_x_ld, xc, xc_ld, x_dbl, x_dbl_ld, M, N, Di, L))
    VMB_UNSUPPORTED("conv_xproj: shape / alignment outside the fused kernel (N == 64, Di %% 64 == 0, "
                    "M >= 512, 16-byte aligned rows)");
  cudaStream_t st = as_stream(stream);
  ProfScope ps(VMB_PROF_CONV, st);
  return conv_xproj_tc(x, x_ld, w_conv, b_conv, w_x, w_x_ld, xc, xc_ld, x_dbl, x_dbl_ld, M, N, Di, L, 1, st);
}

extern "C" int vmb_state_gather(const void* pool, const int32_t* index, void* batch, int n_rows,
                                int64_t row_elems, int dtype, vmb_stream_t stream) {
  return state_rows(pool, batch, index, n_rows, row_elems, dtype, true, as_stream(stream));
}

extern "C" int vmb_state_scatter(void* pool, const int32_t* index, const void* batch, int n_rows,
                                 int64_t row_elems, int dtype, vmb_stream_t stream) {
  return state_rows(batch, pool, index, n_rows, row_elems, dtype, false, as_stream(stream));
}

extern "C" int vmb_prof_enable(int on) {
  g_prof_on = on != 0;
  return VMB_OK;
}

extern "C" int vmb_prof_read(double* ms_sum, int64_t* launches, int reset) {
  VMB_CHECK_ARG(ms_sum && launches, "prof_read: null output");
  std::lock_guard<std::mutex> lk(g_prof_mu);
  for (const ProfRec& r : g_prof_log) {
    if (!r.closed || r.kind < 0 || r.kind >= VMB_PROF_KINDS) continue;
    VMB_CUDA(cudaEventSynchronize(r.b));
    float ms = 0.f;
    VMB_CUDA(cudaEventElapsedTime(&ms, r.a, r.b));
    ms_sum[r.kind] += ms;
    launches[r.kind] += 1;
  }
  if (reset) {
    for (const ProfRec& r : g_prof_log) { g_prof_free.push_back(r.a); g_prof_free.push_back(r.b); }
    g_prof_log.clear();
  }
  return VMB_OK;
}

extern "C" int64_t vmb_launch_count(void) { return (int64_t)g_launches; }
