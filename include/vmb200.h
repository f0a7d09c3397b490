/*
 * vmb200.h -- C ABI of libvmb200.so: the B200 (sm_100a) implementation of the VideoMamba
 * per-block Mamba-mixer hot path.
 *
 * The reference (tannerhoalst/VideoMamba) is pure Python and has no FFI of its own; the seam
 * this library replaces is the set of third-party operator calls the reference makes
 * (SURVEY.md section 8b).  Each entry point below names the reference call site it stands in
 * for (paths relative to the reference root).  Signatures use plain pointers and sizes only:
 * every pointer is a DEVICE pointer unless stated otherwise, every call only ENQUEUES work on
 * `stream` (a cudaStream_t / CUstream passed as void*), never synchronises, never allocates
 * and never frees.  Return value: 0 on success, negative on error (VMB_ERR_*); the message of
 * the last error on the calling thread is available from vmb_last_error().
 *
 * Layout convention: activations are TOKEN-MAJOR, i.e. (batch, token, channel) with channel
 * stride 1 and explicit batch / token strides counted in ELEMENTS.  (The reference works
 * channel-major (B, D, L) because the upstream kernels do; a (B, L, D) buffer viewed as
 * (B, D, L) is the same memory.)  State tensors keep the reference layouts:
 * conv_state (B, d_inner, d_conv), ssm_state (B, d_inner, d_state), both contiguous.
 */
#ifndef VMB200_H_
#define VMB200_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define VMB_ABI_VERSION 5

/* element types */
#define VMB_F32 0
#define VMB_BF16 1

/* status codes */
#define VMB_OK 0
#define VMB_ERR_INVALID (-1)     /* bad argument (null pointer, bad size, bad dtype) */
#define VMB_ERR_UNSUPPORTED (-2) /* shape / option outside what the kernels cover */
#define VMB_ERR_CUDA (-3)        /* a CUDA call failed; see vmb_last_error() */

typedef void* vmb_stream_t;

#if defined(VMB_BUILDING) && defined(__GNUC__)
#define VMB_API __attribute__((visibility("default")))
#else
#define VMB_API
#endif

VMB_API int vmb_abi_version(void);
VMB_API const char* vmb_last_error(void);
/* Host-side query: SM count and compute capability of the current device. */
VMB_API int vmb_device_info(int* sm_count, int* cc_major, int* cc_minor);

/* ------------------------------------------------------------------------------------------
 * Fused residual-add + RMSNorm / LayerNorm.
 * Replaces mamba_ssm.ops.triton.layer_norm.{rms_norm_fn, layer_norm_fn} as called at
 * models/videomamba/videomamba.py:151-166 (prenorm=True) and :902-918 (prenorm=False).
 *   acc = float(x) (+ float(residual));  residual_out = acc (optional);
 *   y   = x_dtype( norm(acc) * weight (+ bias) ),  norm = RMS (is_rms) or LayerNorm.
 * x: (rows, dim) with row stride ldx; residual / y / residual_out: (rows, dim) contiguous.
 * ---------------------------------------------------------------------------------------- */
VMB_API int vmb_add_norm_fwd(const void* x, int x_dtype, int64_t ldx,
                     const void* residual, int residual_dtype,     /* nullable */
                     const void* weight, const void* bias, int w_dtype, /* bias nullable */
                     void* y,                                      /* x_dtype */
                     void* residual_out, int residual_out_dtype,   /* nullable */
                     int64_t rows, int dim, float eps, int is_rms, vmb_stream_t stream);

/* ------------------------------------------------------------------------------------------
 * Refiner fusion gate: out = s * fwd + (1 - s) * bwd with s = sigmoid(g1 (+ g2)), element-wise over
 * n contiguous elements (n % 4 == 0).  Replaces the nn.Sigmoid of `fusion_gate` and the blend
 * `gate * out_fwd + (1.0 - gate) * out_bwd` of BiMambaRefinerBlock.forward
 * (models/refiner_backbone.py:40-43, :129-134).  g1 / g2 are the two halves of
 * Linear(cat[out_fwd, out_bwd]) (g2 nullable when the caller projected the concatenation itself).
 * fp32 math, one rounding to `dtype`.
 * ---------------------------------------------------------------------------------------- */
VMB_API int vmb_gate_blend_fwd(const void* g1, const void* g2 /* nullable */, const void* fwd,
                       const void* bwd, void* out, int64_t n, int dtype, vmb_stream_t stream);

/* ------------------------------------------------------------------------------------------
 * Dense projection  C[M,N] = A[M,K] * W[N,K]^T (+ bias[N]),  fp32 accumulate.
 * Replaces the cuBLAS calls behind in_proj / x_proj / dt_proj / out_proj:
 * models/videomamba/mamba_simple.py:333-339, :409, :413, :445-446 (and :464, :476-479, :496).
 * dtype VMB_BF16 with TMA-compatible shapes runs on tcgen05 tensor cores (TMEM accumulators);
 * VMB_F32 and odd shapes run a CUDA-core kernel in true fp32.
 * lda / ldw / ldc are row strides in elements.
 * ---------------------------------------------------------------------------------------- */
VMB_API int vmb_linear_fwd(const void* A, int64_t lda, const void* W, int64_t ldw,
                   const void* bias,                              /* nullable, same dtype */
                   void* C, int64_t ldc, int64_t M, int N, int K, int dtype,
                   vmb_stream_t stream);
/* The same projection with SiLU applied to the output columns [act_from, N) in the epilogue (on the fp32
 * accumulator, before the one rounding to bf16).  in_proj with act_from = d_inner stores x | SiLU(z): the gate
 * of mamba_simple.py:423-435 computed where the accumulator already sits in registers, so the fused scan
 * (vmb_fused_scan_args.z_gate) only multiplies.  Tensor-core kernel only: VMB_BF16, TMA-compatible operands,
 * act_from a multiple of 64 (or == N: no activation); VMB_ERR_UNSUPPORTED otherwise. */
VMB_API int vmb_linear_fwd_act(const void* A, int64_t lda, const void* W, int64_t ldw,
                   const void* bias,                              /* nullable, same dtype */
                   void* C, int64_t ldc, int64_t M, int N, int K, int dtype, int act_from,
                   vmb_stream_t stream);

/* ------------------------------------------------------------------------------------------
 * Causal conv1d (d_conv 4) + SiLU fused into the x_proj projection (bf16, stateless forward walk):
 *   xc[m, :]    = SiLU(bias + sum_k w[:, k] * x[m + k - 3, :])   with x[row before its sequence] = 0,
 *   x_dbl[m, :] = xc[m, :] * w_x^T                                (fp32 accumulate, one rounding).
 * Replaces causal_conv1d_fn + rearrange + x_proj at models/videomamba/mamba_simple.py:381-409 when no
 * conv_state is carried: the conv output goes to the tensor cores through shared memory and to HBM
 * once.  Rows are (batch, token) flattened: M = B * L rows of Di channels with row pitch x_ld; L tokens
 * per sequence.  w_conv (Di, 4), b_conv (Di) nullable, w_x (N, Di) with N == 64 (x_proj weight zero
 * padded to the x_dbl pitch).  Results are bit-identical to vmb_causal_conv1d_fwd followed by
 * vmb_linear_fwd.  Returns VMB_ERR_UNSUPPORTED for shapes outside that (callers then use the pair).
 * ---------------------------------------------------------------------------------------- */
VMB_API int vmb_conv_xproj_fwd(const void* x, int64_t x_ld, const void* w_conv, const void* b_conv,
                       const void* w_x, int64_t w_x_ld, void* xc, int64_t xc_ld, void* x_dbl,
                       int64_t x_dbl_ld, int64_t M, int N, int Di, int L, vmb_stream_t stream);

/* ------------------------------------------------------------------------------------------
 * Depthwise causal conv1d (+ optional SiLU), token-major, with streaming history.
 * Replaces causal_conv1d.causal_conv1d_fn at models/videomamba/mamba_simple.py:383-399 together
 * with the torch.cat / F.pad state handling around it (:381-404):
 *   hist = [conv_state_in (B,Di,W) or zeros ; x];  y[l] = act(bias + sum_k w[k]*hist[W+l+k-(W-1)])
 *   conv_state_out = last W columns of hist (pre-conv inputs).
 * reverse != 0: the logical sequence is the physical one read back to front (token i of the
 * logical sequence is row L-1-i); history and state columns are in logical order.
 * reverse != 0 with frame_len > 0 (L a multiple of it): FRAME-AXIS reversal -- frames of frame_len
 * tokens are walked back to front, the tokens inside a frame front to back (logical token i is row
 * L - (i / frame_len + 1) * frame_len + i % frame_len): the 4-D flip of BiMambaRefinerBlock
 * (models/refiner_backbone.py:61-68) without the two gather copies.  frame_len = 0 otherwise.
 * ---------------------------------------------------------------------------------------- */
VMB_API int vmb_causal_conv1d_fwd(const void* x, int64_t x_bstride, int64_t x_tstride,
                          const void* weight /* (Di, W) */, const void* bias /* nullable */,
                          const void* conv_state_in, int cs_in_dtype,  /* nullable */
                          void* y, int64_t y_bstride, int64_t y_tstride,
                          void* conv_state_out, int cs_out_dtype,      /* nullable */
                          int B, int L, int Di, int W, int silu, int reverse, int frame_len, int dtype,
                          vmb_stream_t stream);

/* Single-token conv step: rolls conv_state (B,Di,W) in place, returns act(conv).
 * Replaces causal_conv1d.causal_conv1d_update at models/videomamba/mamba_simple.py:468-474. */
VMB_API int vmb_causal_conv1d_update(const void* x /* (B,Di) */, int64_t x_bstride,
                             void* conv_state, int cs_dtype,
                             const void* weight, const void* bias,
                             void* y /* (B,Di) */, int64_t y_bstride,
                             int B, int Di, int W, int silu, int dtype, vmb_stream_t stream);

/* ------------------------------------------------------------------------------------------
 * Selective scan (S6 recurrence), token-major.
 * Replaces mamba_ssm.ops.selective_scan_interface.selective_scan_fn as reached through
 * _selective_scan_with_state (models/videomamba/mamba_simple.py:109-172, call :423-435) with the
 * semantics of _selective_scan_ref (:30-106):
 *   delta = softplus(delta_raw + dt_bias);  h = exp(delta*A)*h + delta*B_t*u_t;
 *   y_t = <C_t,h> + D*u_t;  y_t *= silu(z_t);  h_last = h after the final token (fp32).
 * u, delta_raw, z, y: (B, L, Di) views.  B_t / C_t are read from one token-major buffer `bc`
 * of row stride bc_tstride: B at columns [b_off, b_off+N), C at [c_off, c_off+N).
 * A2 = A * log2(e) with A = -exp(A_log) (fp32, (Di,N)).  reverse != 0 walks tokens L-1..0.
 * ---------------------------------------------------------------------------------------- */
typedef struct vmb_scan_args {
  const void* u;      int64_t u_bstride, u_tstride;
  const void* delta;  int64_t d_bstride, d_tstride;
  const void* z;      int64_t z_bstride, z_tstride;      /* nullable */
  const void* bc;     int64_t bc_bstride, bc_tstride;    int32_t b_off, c_off;
  const float* A2;        /* (Di, N) fp32, A*log2(e) */
  const float* D;         /* (Di) fp32, nullable */
  const float* dt_bias;   /* (Di) fp32, nullable */
  const void* h0;     int32_t h0_dtype;                  /* (B,Di,N), nullable */
  void* y;            int64_t y_bstride, y_tstride;
  float* h_last;                                         /* (B,Di,N) fp32, nullable */
  int32_t B, L, Di, N;
  int32_t dtype;          /* element type of u / delta / z / bc / y */
  int32_t softplus;       /* apply softplus to delta_raw + dt_bias */
  int32_t reverse;
  int32_t frame_len;      /* with reverse != 0: frame-axis reversal (see vmb_causal_conv1d_fwd); else 0 */
} vmb_scan_args;
VMB_API int vmb_selective_scan_fwd(const vmb_scan_args* args, vmb_stream_t stream);

/* ------------------------------------------------------------------------------------------
 * Fused dt_proj + selective scan (bf16, d_state = 16): the production form of the two reference
 * steps models/videomamba/mamba_simple.py:413-414 (dt_proj) and :423-435 (selective_scan_fn),
 * so that delta (B, L, Di) is never materialised:
 *   delta = softplus(w_dt[d, :R] . xdbl[t, :R] + dt_bias[d]);  then the recurrence above with
 *   B_t = xdbl[t, R:R+N], C_t = xdbl[t, R+N:R+2N].
 * u, z, y: (B, L, Di) token-major views; xdbl: (B, L, Xp) rows [dt_low | B | C | pad];
 * w_dt: (Di, Rp) bf16, Rp >= R rounded up to a multiple of 16, columns >= R zero (the projection
 * runs in 16-wide k-steps on the tensor pipe).  Returns VMB_ERR_UNSUPPORTED for shapes the
 * fused kernel does not cover (the caller then uses vmb_linear_fwd + vmb_selective_scan_fwd).
 * ---------------------------------------------------------------------------------------- */
typedef struct vmb_fused_scan_args {
  const void* u;      int64_t u_bstride, u_tstride;
  const void* z;      int64_t z_bstride, z_tstride;
  const void* xdbl;   int64_t x_bstride, x_tstride;
  const void* w_dt;       /* (Di, Rp) bf16 */
  const float* A2;        /* (Di, N) fp32, A*log2(e) */
  const float* D;         /* (Di) fp32, nullable */
  const float* dt_bias;   /* (Di) fp32, nullable */
  const void* h0;     int32_t h0_dtype;                  /* (B,Di,N), nullable */
  void* y;            int64_t y_bstride, y_tstride;
  float* h_last;                                         /* (B,Di,N) fp32, nullable */
  int32_t B, L, Di, N, R, Rp, Xp;
  int32_t reverse;
  /* optional scratch of vmb_fused_scan_workspace_bytes(...) bytes (device, 16-byte aligned): lets
   * small batches split the sequence into concurrently processed segments (exact two-pass carry);
   * without it the call still works, one warp per 16 channels of a sequence. */
  void* workspace;    int64_t workspace_bytes;
  /* a_geometric != 0: the caller guarantees A2[d][n] == (n+1) * A2[d][0] for every channel (the exact
   * S4D-real structure of the reference's initialisation, models/videomamba/mamba_simple.py:265-272;
   * check it when the weights are loaded): the decay factors of a channel then come from two
   * exponentials and packed multiplies instead of sixteen exponentials.  0 = general A. */
  int32_t a_geometric;
  /* measurement aid, 0 = automatic: 100 * minimum segment of the sequence split (1 .. 4 = 32 .. 128 tokens)
   * + 10 * layout (1 one warp per unit, 2 two warps, 3 two warps + sequence
   * split) + evaluator (1 = MUFU only, 2..5 = 1..4 of a lane's four state pairs on the FMA-pipe
   * polynomial, 9 = geometric).  Evaluators other than the built default and 9 exist only in
   * measurement builds (-DVMB_SCAN_LAB) and return VMB_ERR_UNSUPPORTED otherwise. */
  int32_t tune;
  int32_t frame_len;      /* with reverse != 0: frame-axis reversal (see vmb_causal_conv1d_fwd); else 0 */
  /* training forward (reverse == 0 only), nullable: vmb_scan_bwd_ckpt_bytes(B, L, Di) bytes (device, 16-byte
   * aligned) that receive the state before every 4-token group -- handed to vmb_selective_scan_bwd as
   * vmb_scan_bwd_args.fwd_ckpt, the backward then skips its own forward pass over the sequence. */
  float* bwd_ckpt;
  /* z_gate != 0: `z` already holds SiLU(z) (vmb_linear_fwd_act applied it in the in_proj epilogue): the
   * scan multiplies by it as it is.  0 = `z` is the raw in_proj output (mamba_simple.py:369). */
  int32_t z_gate;
} vmb_fused_scan_args;
VMB_API int64_t vmb_fused_scan_workspace_bytes(int B, int L, int Di, int N);
VMB_API int vmb_selective_scan_fused_fwd(const vmb_fused_scan_args* args, vmb_stream_t stream);

/* Single recurrent step, state (B,Di,N) updated in place.  Replaces
 * mamba_ssm.ops.triton.selective_state_update at models/videomamba/mamba_simple.py:483-494
 * (and the per-token fallback loop :158-171).  x, dt, z, y: (B,Di); Bm, Cm: (B,N). */
VMB_API int vmb_selective_state_update(void* state, int state_dtype,
                               const void* x, int64_t x_bstride,
                               const void* dt, int64_t dt_bstride,
                               const float* A2, const void* Bm, int64_t b_bstride,
                               const void* Cm, int64_t c_bstride,
                               const float* D, const void* z, int64_t z_bstride,
                               const float* dt_bias, int softplus,
                               void* y, int64_t y_bstride,
                               int B, int Di, int N, int dtype, vmb_stream_t stream);

/* ------------------------------------------------------------------------------------------
 * Fused single-token decode step: everything between in_proj and out_proj of Mamba.step
 * (models/videomamba/mamba_simple.py:466-494) in one kernel -- causal_conv1d_update (:468-474),
 * x_proj (:476), split + dt_proj without bias (:477-479), selective_state_update with dt_bias,
 * softplus, D skip and the SiLU(z) gate (:483-494).  conv_state (B, Di, W) and ssm_state (B, Di, N)
 * are updated IN PLACE, as the reference's step does.  xz (B, 2Di) is the in_proj output
 * (x = [:, :Di], z = [:, Di:]); y (B, Di) is the gated output that goes into out_proj.  The conv
 * output, x_dbl and delta_raw are rounded to `dtype` between the stages like the reference's
 * separate calls; the state update runs in fp32.  Any Di / N / R / W; one CTA per batch row.
 * ---------------------------------------------------------------------------------------- */
typedef struct vmb_step_args {
  const void* xz;      int64_t xz_bstride;               /* (B, 2Di) */
  void* conv_state;    int32_t cs_dtype;                 /* (B, Di, W) contiguous, in place */
  void* ssm_state;     int32_t ss_dtype;                 /* (B, Di, N) contiguous, in place */
  const void* w_conv;  /* (Di, W) */
  const void* b_conv;  /* (Di) nullable */
  const void* w_x;     /* (R+2N, Di) contiguous */
  const void* w_dt;    /* (Di, R) contiguous */
  const float* A2;     /* (Di, N) fp32, A*log2(e) */
  const float* Dskip;  /* (Di) fp32, nullable */
  const float* dt_bias;/* (Di) fp32, nullable */
  void* y;             int64_t y_bstride;                /* (B, Di) */
  int32_t B, Di, N, R, W;
  int32_t dtype;       /* element type of xz / weights / y */
} vmb_step_args;
VMB_API int vmb_mixer_step_fwd(const vmb_step_args* args, vmb_stream_t stream);

/* ------------------------------------------------------------------------------------------
 * Whole mixer block: in_proj -> conv(+SiLU) -> x_proj -> dt_proj -> scan(+gate) -> out_proj.
 * Replaces the body of Mamba.forward, models/videomamba/mamba_simple.py:332-446, for the
 * stateless, (conv_state, ssm_state)-streaming and ssm-only cases, and stands where
 * mamba_ssm's mamba_inner_fn (:352-366) stands on the reference's fast path.
 * All weights are in `dtype`; A2 / D / dt_bias are fp32 (prepared once per weight load).
 * `workspace` must hold vmb_mixer_workspace_bytes(...) bytes (device memory, 256-B aligned).
 * ---------------------------------------------------------------------------------------- */
typedef struct vmb_mixer_args {
  const void* hidden;  int64_t h_bstride, h_tstride;     /* (B, L, D) */
  void* out;           int64_t o_bstride, o_tstride;     /* (B, L, D) */
  const void* w_in;    /* (2Di, D)  */
  const void* b_in;    /* (2Di) nullable */
  const void* w_conv;  /* (Di, W)   */
  const void* b_conv;  /* (Di) nullable */
  const void* w_x;     /* (R+2N, Di) */
  const void* w_dt;    /* (Di, R)   */
  const void* w_out;   /* (D, Di)   */
  const void* b_out;   /* (D) nullable */
  const float* A2;     /* (Di, N) fp32 */
  const float* Dskip;  /* (Di) fp32 */
  const float* dt_bias;/* (Di) fp32 */
  /* fast-path operands prepared at weight-load time (nullable => generic path):
   * w_x_pad (Xp, Di) zero-padded rows, w_dt_pad (Di, Rp) zero-padded columns */
  const void* w_x_pad; const void* w_dt_pad; int32_t Xp, Rp;
  const void* conv_state_in;  int32_t cs_in_dtype;       /* nullable */
  void* conv_state_out;       int32_t cs_out_dtype;      /* nullable */
  const void* ssm_state_in;   int32_t ss_in_dtype;       /* nullable */
  float* ssm_state_out;                                  /* nullable, fp32 */
  void* workspace;     int64_t workspace_bytes;
  int32_t B, L, D, Di, N, R, W;
  int32_t dtype;
  int32_t reverse;     /* walk tokens L-1..0 (conv and scan) */
  int32_t path;        /* 0 = auto, 1 = force generic kernels, 2 = force fast kernels */
  int32_t a_geometric; /* as in vmb_fused_scan_args */
  int32_t scan_tune;   /* as vmb_fused_scan_args.tune */
  int32_t frame_len;   /* with reverse != 0: frame-axis reversal (see vmb_causal_conv1d_fwd); else 0 */
  int32_t fuse_conv_xproj; /* != 0: stateless forward walks run conv + x_proj as ONE kernel (vmb_conv_xproj_fwd):
                            * less HBM traffic and faster for a single forward in flight, slower when several
                            * forwards share the GPU (it fills the SMs' shared memory); bit-identical results */
  int32_t gate_in_proj;    /* != 0: on the fast path in_proj stores SiLU(z) (vmb_linear_fwd_act) and the scan multiplies
                            * by it as it is: one activation less in the scan's instruction stream; the gate is
                            * rounded to bf16 after the activation instead of before it (same bf16 tolerance,
                            * not bit-identical to gate_in_proj == 0) */
} vmb_mixer_args;
VMB_API int64_t vmb_mixer_workspace_bytes(int B, int L, int D, int Di, int N, int R, int dtype);
VMB_API int vmb_mixer_fwd(const vmb_mixer_args* args, vmb_stream_t stream);

/* ------------------------------------------------------------------------------------------
 * Streaming state I/O: move per-stream state rows between a resident pool and a batch.
 * Serves the (conv_state, ssm_state) carry of the streaming contract
 * (models/videomamba/streaming.py:77-92, models/videomamba/videomamba.py:526-549): row i of
 * `batch` <-> row index[i] of `pool`; rows are `row_elems` elements of `dtype`.
 * ---------------------------------------------------------------------------------------- */
VMB_API int vmb_state_gather(const void* pool, const int32_t* index, void* batch,
                     int n_rows, int64_t row_elems, int dtype, vmb_stream_t stream);
VMB_API int vmb_state_scatter(void* pool, const int32_t* index, const void* batch,
                      int n_rows, int64_t row_elems, int dtype, vmb_stream_t stream);

/* ------------------------------------------------------------------------------------------
 * Front glue of the token path.
 * vmb_patchify: clip x (B, C, T, H, W), contiguous -> patch rows cols (B*t*h*w, C*k*ph*pw) with
 * t = T/k, h = H/ph, w = W/pw (trailing remainders are cropped), row order (b, t, y, x), column
 * order (c, dt, dy, dx).  It is the im2col of the reference's Conv3d whose kernel equals its stride
 * (PatchEmbed, models/videomamba/videomamba.py:359-368); the projection itself is vmb_linear_fwd.
 * vmb_embed_tokens: out (B, has_cls + t*hw, D) = patches (B, t, hw, D) + spatial (hw, D) +
 * temporal (t, D), each add rounded to `dtype` like the reference's two adds, and row 0 =
 * cls_row (D) when cls_row is not NULL (models/videomamba/videomamba.py:806-823: position
 * embeddings, CLS placement; continuation chunks pass NULL).
 * ---------------------------------------------------------------------------------------- */
VMB_API int vmb_patchify(const void* x, void* cols, int64_t B, int C, int T, int H, int W, int k,
                 int ph, int pw, int dtype, vmb_stream_t stream);
VMB_API int vmb_embed_tokens(const void* patches, const void* spatial, const void* temporal,
                     const void* cls_row,                          /* nullable */
                     void* out, int64_t B, int t, int hw, int D, int dtype, vmb_stream_t stream);

/* ------------------------------------------------------------------------------------------
 * Back glue of the token path.
 * vmb_pool_norm_fwd: the pooling block of PretrainVideoMamba.forward
 * (models/videomamba/videomamba.py:983-1063).  x (B, has_cls + G*per, C) are the tokens after the final
 * norm (token 0 = CLS when has_cls); group g = patch tokens [g*per, (g+1)*per) (G = 1: all patch tokens;
 * G = temporal tokens, per = tokens per frame: keep_temporal).  mode 0 `cls`: LN(cls) -> out (B, 1, C);
 * 1 `cls+avg`: LN(cls + mean_g) -> (B, G, C); 2 `cls_cat_avg`: LN(cat[cls, mean_g]) -> (B, 1 + G, C);
 * 3 `avg`: LN(mean_g) -> (B, G, C).  LN = nn.LayerNorm(C) (pool_norm: weight / bias nullable, eps).  The mean
 * and the sum with CLS are rounded to `dtype` like the reference's torch ops.  workspace: device scratch of
 * vmb_pool_norm_workspace_bytes(B, G, per, C) bytes (unused in mode 0).
 * vmb_gather_rows: dst (B, n, C) = src[b, index[b, i], :] -- the visible-token gather of the masked path
 * (videomamba.py:826-836); index (B, n) int64 on the device.
 * ---------------------------------------------------------------------------------------- */
VMB_API int64_t vmb_pool_norm_workspace_bytes(int B, int G, int per, int C);
VMB_API int vmb_pool_norm_fwd(const void* x, int64_t x_bstride, int64_t x_tstride, int B, int G, int per, int C,
                      int has_cls, int mode, const void* ln_weight, const void* ln_bias, float eps,
                      void* out, void* workspace, int64_t workspace_bytes, int dtype, vmb_stream_t stream);
VMB_API int vmb_gather_rows(const void* src, int64_t src_bstride, int64_t src_tstride, const int64_t* index,
                    int B, int n_per_batch, int C, void* dst, int dtype, vmb_stream_t stream);

/* ------------------------------------------------------------------------------------------
 * Backward passes (training).  The reference gets its gradients from
 * autograd through the third-party operators: the streaming-training check differentiates through
 * the carried (conv_state, ssm_state) (scripts/check_streaming_state.py:47-60) and Block.forward wraps
 * the mixer in activation checkpointing (models/videomamba/videomamba.py:168-206).  These entry points
 * are the backward halves of the forward operators above, for the FORWARD token order (the reversed
 * walk of BiMambaRefinerBlock is differentiated through flip copies).  fp32 math; every reduction over
 * rows / batch is two-stage through the caller's workspace (no atomics: deterministic).  Parameter
 * gradients (dweight, dbias, dA, dD, ddt_bias) are written, not accumulated, as fp32.
 *
 * vmb_add_norm_bwd: backward of vmb_add_norm_fwd.  x / residual / weight are the forward inputs, dy the
 *   gradient of y (x's dtype), dresidual_out the gradient of residual_out (nullable).  dx (rows, dim,
 *   x's dtype); dresidual (residual's dtype, nullable) receives the same values; dweight / dbias (dim)
 *   fp32, nullable.
 * vmb_causal_conv1d_bwd: backward of vmb_causal_conv1d_fwd (reverse = 0).  dy (B, L, Di) contiguous;
 *   dconv_state_out = gradient of the returned state (nullable); dx a (B, L, Di) view with element strides
 *   dx_bstride / dx_tstride (channel stride 1; 0, 0 = contiguous), e.g. the x half of a d(xz) buffer;
 *   dconv_state_in (conv_state_in's dtype, nullable); dweight (Di, W), dbias (Di) fp32, nullable.
 * vmb_selective_scan_bwd: backward of vmb_selective_scan_fwd (reverse = 0, d_state <= 16).
 *   dout (B, L, Di) view; dh_last (B, Di, N) fp32 nullable.  du / ddelta / dz (B, L, Di) contiguous in
 *   `dtype` (ddelta is the gradient of delta_raw, i.e. through the softplus); dbc has bc's layout: the
 *   kernel writes columns [b_off, b_off+N) and [c_off, c_off+N) of every row and leaves the rest alone;
 *   dA is the gradient with respect to A (= A2 / log2(e)), (Di, N) fp32; dD, ddt_bias (Di) fp32; dh0
 *   (B, Di, N) fp32; all four nullable.
 * vmb_transpose_2d / vmb_colsum: out (cols, rows; row stride ldo) = in^T; out (N) = column sums of x (M, N).  The
 *   projection backward is two vmb_linear_fwd calls on transposed operands (dX = dY W, dW = dY^T X)
 *   plus a column sum for the bias.
 * ---------------------------------------------------------------------------------------- */
VMB_API int64_t vmb_add_norm_bwd_workspace_bytes(int64_t rows, int dim);
VMB_API int vmb_add_norm_bwd(const void* x, int x_dtype, int64_t ldx,
                     const void* residual, int residual_dtype,       /* nullable */
                     const void* weight, int w_dtype,
                     const void* dy,                                 /* x_dtype */
                     const void* dresidual_out, int dresidual_out_dtype, /* nullable */
                     void* dx, void* dresidual /* nullable */,
                     float* dweight, float* dbias,                   /* nullable */
                     int64_t rows, int dim, float eps, int is_rms,
                     void* workspace, int64_t workspace_bytes, vmb_stream_t stream);
VMB_API int64_t vmb_causal_conv1d_bwd_workspace_bytes(int B, int L, int Di, int W);
VMB_API int vmb_causal_conv1d_bwd(const void* x, int64_t x_bstride, int64_t x_tstride,
                          const void* weight, const void* bias /* nullable */,
                          const void* conv_state_in, int cs_in_dtype,     /* nullable */
                          const void* dy,
                          const void* dconv_state_out, int dcs_out_dtype, /* nullable */
                          void* dx, int64_t dx_bstride, int64_t dx_tstride, /* 0, 0 = contiguous (B, L, Di) */
                          void* dconv_state_in /* nullable */,
                          float* dweight, float* dbias,                   /* nullable */
                          int B, int L, int Di, int W, int silu, int dtype,
                          void* workspace, int64_t workspace_bytes, vmb_stream_t stream);
typedef struct vmb_scan_bwd_args {
  const void* u;      int64_t u_bstride, u_tstride;
  const void* delta;  int64_t d_bstride, d_tstride;
  const void* z;      int64_t z_bstride, z_tstride;      /* nullable */
  const void* bc;     int64_t bc_bstride, bc_tstride;    int32_t b_off, c_off;
  const float* A2;        /* (Di, N) fp32, A*log2(e) */
  const float* D;         /* (Di) fp32, nullable */
  const float* dt_bias;   /* (Di) fp32, nullable */
  const void* h0;     int32_t h0_dtype;                  /* (B,Di,N), nullable */
  const void* dout;   int64_t dout_bstride, dout_tstride;
  const float* dh_last;                                  /* (B,Di,N) fp32, nullable */
  void* du;  void* ddelta;  void* dz;                    /* (B,L,Di) contiguous; dz nullable iff z is */
  int64_t dz_bstride, dz_tstride;                        /* dz only: element strides of a (B,L,Di) VIEW (e.g. the z half
                                                            of a d(xz) buffer); 0, 0 = contiguous */
  void* dbc;          int64_t dbc_tstride;               /* rows (B*L), same columns as bc */
  float* dA;  float* dD;  float* ddt_bias;  float* dh0;  /* nullable */
  void* workspace;    int64_t workspace_bytes;
  int32_t B, L, Di, N;
  int32_t dtype;
  int32_t softplus;
  /* nullable: the records vmb_selective_scan_fused_fwd wrote through bwd_ckpt for the SAME u / delta / B / h0
   * (bf16, d_state 16, Di % 16 == 0 only; ignored by the generic kernels, which always recompute) */
  const float* fwd_ckpt;
} vmb_scan_bwd_args;
VMB_API int64_t vmb_selective_scan_bwd_workspace_bytes(int B, int L, int Di, int N);
/* bytes of the state records of one (B, L, Di) scan with d_state 16: 1 KB per (batch, 16 channels, 4 tokens) */
VMB_API int64_t vmb_scan_bwd_ckpt_bytes(int B, int L, int Di);
VMB_API int vmb_selective_scan_bwd(const vmb_scan_bwd_args* args, vmb_stream_t stream);
/* Weight gradient of a projection without transposed copies: dw (N, K) = dy^T x, dy (M, N) row stride
 * ldy, x (M, K) row stride ldx, both bf16 with 16-byte aligned rows (N, K, ldy, ldx multiples of 8); dw in
 * dw_dtype.  Tensor-core kernel with the token axis split over the grid (csrc/wgrad.cu).  Returns
 * VMB_ERR_UNSUPPORTED for other layouts (callers then use vmb_transpose_2d + vmb_linear_fwd). */
VMB_API int64_t vmb_linear_wgrad_workspace_bytes(int64_t M, int N, int K);
VMB_API int vmb_linear_wgrad(const void* dy, int64_t ldy, const void* x, int64_t ldx, void* dw, int dw_dtype,
                     int64_t M, int N, int K, void* workspace, int64_t workspace_bytes,
                     vmb_stream_t stream);
VMB_API int vmb_transpose_2d(const void* in, int64_t ld, void* out, int64_t ldo, int64_t rows, int cols,
                     int dtype, vmb_stream_t stream);
VMB_API int64_t vmb_colsum_workspace_bytes(int64_t M, int N);
VMB_API int vmb_colsum(const void* x, int64_t ld, int64_t M, int N, int dtype, void* out, int out_dtype,
               void* workspace, int64_t workspace_bytes, vmb_stream_t stream);

/* ------------------------------------------------------------------------------------------
 * Per-stage device timing (measurement aid, off by default; nothing like it exists in the
 * reference).  While enabled, every entry point above brackets the kernels it enqueues with
 * cudaEvents on `stream`, tagged with the stage they belong to.  vmb_prof_read synchronises
 * the recorded events, adds their elapsed milliseconds and launch counts per stage into the
 * caller's arrays (VMB_PROF_KINDS entries each, HOST pointers) and optionally clears the log.
 * Must be disabled while a stream is being captured into a CUDA graph.
 * ---------------------------------------------------------------------------------------- */
#define VMB_PROF_ADD_NORM 0
#define VMB_PROF_IN_PROJ 1
#define VMB_PROF_CONV 2
#define VMB_PROF_X_PROJ 3
#define VMB_PROF_DT_PROJ 4
#define VMB_PROF_SCAN 5
#define VMB_PROF_OUT_PROJ 6
#define VMB_PROF_OTHER 7
#define VMB_PROF_KINDS 8
VMB_API int vmb_prof_enable(int on);
VMB_API int vmb_prof_read(double* ms_sum, int64_t* launches, int reset);
/* Number of kernels this library has enqueued so far in this process (all threads). */
VMB_API int64_t vmb_launch_count(void);

#ifdef __cplusplus
}
#endif
#endif /* VMB200_H_ */
