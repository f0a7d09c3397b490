// Internal (non-ABI) entry points shared between the translation units of libvmb200.
#pragma once
#include <cuda.h>

#include "common.cuh"

namespace vmb {

// linear_simt.cu -- CUDA-core projection (true fp32 / any shape)
int linear_simt(const void* A, int64_t lda, const void* W, int64_t ldw, const void* bias, void* C,
                int64_t ldc, int64_t M, int N, int K, int dtype, cudaStream_t st);

// gemm_tc.cu -- tcgen05 / TMEM / TMA bf16 projection.  Returns VMB_ERR_UNSUPPORTED (without
// setting an error the caller must report) when the shape does not fit the tensor-core kernel.
bool gemm_tc_supported(const void* A, int64_t lda, const void* W, int64_t ldw, const void* C,
                       int64_t ldc, int64_t M, int N, int K);
// act_from in [0, N), a multiple of 64: SiLU on the output columns [act_from, N) in the epilogue (fp32, before the
// one rounding to bf16); anything else = no activation.
int gemm_tc(const void* A, int64_t lda, const void* W, int64_t ldw, const void* bias, void* C,
            int64_t ldc, int64_t M, int N, int K, cudaStream_t st, int act_from = -1);

// gemm_tc.cu -- cached 3-D bf16 tensor map (dims d0 (contiguous) x d1 x d2, byte strides of d1 / d2,
// box0 x box1 x 1 boxes, optional 128-byte swizzle) for TMA staging outside the GEMM.
int make_tensor_map_3d_bf16(CUtensorMap* out, const void* ptr, uint64_t d0, uint64_t d1, uint64_t d2,
                            uint64_t stride1_bytes, uint64_t stride2_bytes, uint32_t box0,
                            uint32_t box1, bool swizzle128);

int make_tensor_map_2d_bf16_sw128(CUtensorMap* out, const void* ptr, int64_t rows, int64_t cols, int64_t ld,
                                  int box_rows);

// wgrad_tc.cu -- weight gradient dW (N, K) = dY^T X on tcgen05 (both operands MN-major: the token axis is the
// contraction).  partial: [splits][N][K] fp32 (summed by the caller); supported() decides, splits() sizes.
bool wgrad_tc_supported(const void* dy, int64_t ldy, const void* x, int64_t ldx, int64_t M, int N, int K);
int wgrad_tc_splits(int64_t M, int N, int K);
int wgrad_tc(const void* dy, int64_t ldy, const void* x, int64_t ldx, float* partial, int64_t M, int N, int K,
             int splits, cudaStream_t st);

// gemm_tc.cu -- causal conv (d_conv 4, bf16) + SiLU fused into the x_proj projection: x (M rows of K
// channels, row pitch x_ld; rows are (batch, token) flattened, L tokens per sequence), conv taps cw
// (K, 4) / bias cb (K), W (N = 64, K) -> xc (M, K) and C = xc * W^T (M, N).  Stateless forward walk only.
bool conv_xproj_supported(const void* x, int64_t x_ld, const void* cw, const void* cb, const void* W,
                          int64_t ldw, const void* xc, int64_t xc_ld, const void* C, int64_t ldc,
                          int64_t M, int N, int K, int L);
int conv_xproj_tc(const void* x, int64_t x_ld, const void* cw, const void* cb, const void* W,
                  int64_t ldw, void* xc, int64_t xc_ld, void* C, int64_t ldc, int64_t M, int N, int K,
                  int L, int silu, cudaStream_t st);

// backward.cu -- out[n] = sum over P rows of partial[P][n] (fp32 partials)
int reduce_partials(const float* partial, int P, int64_t n, void* out, int out_dtype, cudaStream_t st);

// scan_bwd_fast.cu -- selective-scan backward for bf16, d_state 16, Di % 16 == 0, 16-byte aligned rows:
// the checkpoint pass + the reverse pass; slabs / pA / pD / pBias laid out as in scan_bwd.cu, which
// runs the final reductions.
bool scan_bwd_fast_supported(const vmb_scan_bwd_args& a);
int64_t scan_bwd_fast_ckpt_bytes(int B, int L, int Di);
int scan_bwd_fast_segments(int B, int L, int Di, int* seg_tiles);   // sequence split for small batches (1 = none)
int64_t scan_bwd_fast_seg_bytes(int B, int L, int Di);              // scratch of the split (seg_ws)
// *nparts = rows of the pA / pD / pBias partials written (segments x batch)
int scan_bwd_fast(const vmb_scan_bwd_args& a, float* ckpt, float* slabs, float* pA, float* pD, float* pBias,
                  float* seg_ws, int* nparts, cudaStream_t st);

// scan_generic.cu
int scan_generic(const vmb_scan_args& a, cudaStream_t st);

// scan_fast.cu -- production bf16 scan with the dt projection fused in.
struct FastScanArgs {
  const void* u;  int64_t u_bs, u_ts;       // conv output (B, L, Di)
  const void* z;  int64_t z_bs, z_ts;       // gate (B, L, Di)
  const void* xdbl; int64_t x_bs, x_ts;     // (B, L, Xp): [0,R) dt_low | [R,R+N) B | [R+N,R+2N) C
  const void* w_dt_pad;                     // (Di, Rp) bf16, zero padded
  const float* A2; const float* D; const float* dt_bias;
  const void* h0; int h0_dtype;
  void* y; int64_t y_bs, y_ts;
  float* h_last;
  int B, L, Di, N, R, Rp, Xp;
  int reverse;
  int frame_len = 0;     // with reverse: frame-axis reversal, frames of frame_len tokens (0 = whole-sequence reversal)
  int a_geometric = 0;   // caller's promise: A2[d][n] == (n+1) * A2[d][0] (checked when the weights are loaded)
  int z_gate = 0;        // z already holds SiLU(z) (applied by the in_proj epilogue): the scan only multiplies
  int tune = 0;          // measurement aid: 10 * layout + evaluator, 0 = automatic (see scan_fast())
  float* ckpt = nullptr; // forward walk only: records of the state before every 4-token group (scan_bwd_fast.cu layout)
  // sequence split for small batches (filled by scan_fast itself): nseg segments of seg_len tokens,
  // carried through seg_ws = [H (nseg,B,Di,N) | S (nseg,B,Di)] fp32 (pass 2 chains them itself)
  float* seg_ws = nullptr;
  int64_t seg_ws_bytes = 0;
  int nseg = 1, seg_len = 0;
};
int64_t scan_fast_workspace_bytes(int B, int L, int Di, int N);
bool scan_fast_supported(const FastScanArgs& a);
int scan_fast(const FastScanArgs& a, cudaStream_t st);

}  // namespace vmb
