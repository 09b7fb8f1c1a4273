// Persistent warp-specialised bf16 GEMM on tcgen05 + TMA:   C[M,N] = epilogue(A[M,K] * W[N,K]^T)
// Interface only; the kernel lives in gemm_tc.cu.
// A and W are K-major bf16 (row pitch in elements, multiple of 8), fp32 accumulation in TMEM (two accumulator
// stages so the epilogue of tile i overlaps the MMAs of tile i+1).  The epilogue fuses bias, GELU/ReLU, an fp32
// residual read at the destination row, a window-unpartition / pixel-shuffle row remap and an fp32 or bf16 output.
#pragma once
#include <cuda_bf16.h>
#include <cuda_runtime.h>

namespace cvb {

enum { GEMM_ACT_NONE = 0, GEMM_ACT_GELU = 1, GEMM_ACT_RELU = 2 };
enum { GEMM_MAP_IDENTITY = 0, GEMM_MAP_UNWINDOW = 1, GEMM_MAP_SHUFFLE2 = 2, GEMM_MAP_POOL2 = 3, GEMM_MAP_QPOOL = 4 };

struct GemmEpilogue {
  const float* bias = nullptr;       // [N] (for SHUFFLE2: [N/4], indexed by output channel)
  int act = GEMM_ACT_NONE;
  int res_before_act = 0;            // 0: act(acc+bias)+res   1: act(acc+bias+res)
  const float* res = nullptr;        // fp32 residual, read at the destination row/column
  long long ld_res = 0;
  long long res_row_mod = 0;         // > 0: residual row = dest row % res_row_mod (per-image broadcast tables)
  float* out_f32 = nullptr;
  long long ld_f32 = 0;
  __nv_bfloat16* out_bf16 = nullptr;
  long long ld_bf16 = 0;
  int map_mode = GEMM_MAP_IDENTITY;
  // UNWINDOW: source row = ((b*nwy + wy)*nwx + wx)*ws*ws + ty*ws + tx  ->  dest row (b*H + wy*ws+ty)*W + wx*ws+tx,
  //           rows that fall in the window padding (y >= H or x >= W) are dropped.
  // POOL2:    window-major source rows (ws = 4 or 8, no window padding); the output row (b, y/2, x/2) on the H/2 x W/2 grid
  //           receives the MAXIMUM over the 2 x 2 source rows (Hiera's Q-pool shortcut: maxpool2x2(proj(norm(x)))),
  //           fp32 output, no residual — the four rows of a group sit in one warp's TMEM lanes, so the pooling is two
  //           shuffles per element in the epilogue and only a quarter of the rows is ever written.
  // QPOOL:    16-bit identity output (TMA stores) for the columns >= pool_cols; the columns < pool_cols (the q part of a
  //           Q-pooled block's qkv) are 2 x 2 max-pooled over window-major rows (ws = 4 or 8) and written to `pool_out`
  //           in pooled window-major order (row = win*(ws/2)^2 + (ty/2)*(ws/2) + tx/2); their full-resolution copy is
  //           never written (nothing reads it).
  // SHUFFLE2: source row (b,y,x) on an H x W grid, column n = (dy*2+dx)*Cout + co -> dest row
  //           (b*2H + 2y+dy)*2W + 2x+dx, dest column co   (ConvTranspose2d kernel 2 stride 2)
  int ws = 0, nwx = 0, nwy = 0, H = 0, W = 0, cout = 0;
  int pool_cols = 0;                 // QPOOL
  __nv_bfloat16* pool_out = nullptr;
  long long ld_pool = 0;
  int fp16 = 0;                      // operands (A, W) and the 16-bit output are IEEE half instead of bf16
  int gelu_h2 = 0;                   // (set by the launcher) fp16 output + tanh-form GELU: evaluate the GELU in half2 (act.cuh)
};

struct GemmProblem {
  int M, N, K;
  int n_tiles_m, n_tiles_n;
  int rotate;  // epilogue: the column-chunk group a warp takes rotates from tile to tile (balances BN / 32 chunks over EW / 4 groups)
  int res_stage;  // epilogue: fp32 residual loaded lanes-along-columns and passed through the staging buffer
  int res_l2pf;   // epilogue: in-place residual (res == out) boxes of a warp's next tile bulk-prefetched into L2
  unsigned long long* trace;  // debug timeline of CTA 0 (cv_gemm_set_trace): slot = event * 256 + local tile index, else nullptr
};

// Host launcher.  Returns CV_OK or a CV_ERR_* status (message via cv_last_error()).
int gemm_tc_launch(const __nv_bfloat16* A, long long lda, const __nv_bfloat16* W, long long ldw, int M, int N, int K,
                   const GemmEpilogue& epi, int num_sms, cudaStream_t st);

}  // namespace cvb
