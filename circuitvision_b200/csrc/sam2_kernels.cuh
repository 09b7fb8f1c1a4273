// Bandwidth-bound / small CUDA-core kernels of the SAM 2.1 path (everything that is not a large GEMM or the Hiera
// attention): preprocessing + im2col, LayerNorm with fused window partition, Q / shortcut max-pooling, the
// token-side decoder arithmetic (fp32), LayerNorm2d+GELU, mask product + stability counters, and the fused
// x4-upsample + multi-kernel refinement + threshold tail.  Declarations only; definitions in sam2_kernels.cu.
#pragma once
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>

namespace cvb {

// ---- a1 + patch-embed operand: u8 HWC image (already 1024x1024) -> normalised 7x7/stride-4 patches as a
// two-term bf16 split [B*256*256, 2*PE_K] (hi | lo), K = 147 padded to PE_K.
constexpr int PE_K = 152;
constexpr int PE_K8 = 168;  // uint8 path: 7 kernel rows x (21 taps + 3 zeros), see k_im2col_u8raw
int launch_im2col_u8(const uint8_t* img, int B, int S, const float* mean, const float* inv_std, int swap_rb,
                     __nv_bfloat16* out, cudaStream_t st);
// raw-pixel variant: out [B*256*256, PE_K8] = bf16(pixel value), normalisation folded into "pe.w8" / "pos8"
// Every `fp16` argument below selects IEEE half (saturating conversion) instead of bf16 for the 16-bit operand
// buffers; the `__nv_bfloat16*` types are then just opaque 16-bit storage.
int launch_im2col_u8raw(const uint8_t* img, int B, int S, int swap_rb, int fp16, __nv_bfloat16* out, cudaStream_t st);
int launch_im2col_f32(const float* img_chw, int B, int S, int fp16, __nv_bfloat16* out, cudaStream_t st);

// ---- a1 for crops that are not 1024x1024: ToTensor (/255) -> bilinear antialias resize to SxS (the arithmetic of
// F.interpolate(mode="bilinear", antialias=True, align_corners=False): width pass then height pass, triangle filter
// whose support grows with the down-scale factor) -> Normalize.  tmp: float [H, S, 3].  out: float CHW [3,S,S].
int launch_preprocess_aa(const uint8_t* img_hwc, int H, int W, int S, const float* mean, const float* inv_std, int swap_rb,
                         float* tmp, float* out_chw, cudaStream_t st);

// batched page form: geom = B records {int64 byte offset of the page, int page width, x0, y0, x1, y1, pad} (32 bytes each, device);
// image b = crop window of page b; tmp: float [B, max_hc, S, 3]; out: float [B, 3, S, S]
int launch_preprocess_pages(const uint8_t* pages, const void* geom, int B, int max_hc, int S, const float* mean,
                            const float* inv_std, int swap_rb, float* tmp, float* out, cudaStream_t st);

// ---- LayerNorm over the channel dim with optional window partition (zero rows for window padding).
// ws == 0: identity row order.  gamma == nullptr: plain fp32 -> bf16 cast.  out_f32 optional (normalised, fp32).
int launch_ln_rows(const float* X, long long n_src_rows, int C, const float* gamma, const float* beta, float eps,
                   int B, int H, int W, int ws, int fp16, __nv_bfloat16* out_bf16, float* out_f32, cudaStream_t st, __nv_bfloat16* raw16 = nullptr);

// ---- Q pooling: Qp[(b,win,py,px), c] = max_{2x2} QKV[(b,win,2py+dy,2px+dx), c], c < Cq (bf16, window-major)
int launch_pool_q(const __nv_bfloat16* qkv, long long ld, int n_windows, int ws, int Cq, int fp16, __nv_bfloat16* qp,
                  cudaStream_t st);
// ---- shortcut pooling: S fp32 window-major [B*nwy*nwx*ws*ws, C] -> R fp32 grid order [B*(H/2)*(W/2), C]
int launch_pool_shortcut(const float* S, int B, int H, int W, int ws, int C, float* R, cudaStream_t st);
// ---- Y[(b,y,x), :] += L[(b,y/2,x/2), :]   (FPN top-down nearest x2)
int launch_add_nearest2(float* Y, const float* L, int B, int H, int W, int C, cudaStream_t st);

// ---- decoder, token side (fp32)
// C[R,N] = act(A[R,K] * W[N,K]^T + bias) (+ res);  act: 0 none 1 gelu 2 relu 3 sigmoid
int launch_tok_linear(const float* A, long long lda, const float* W, const float* bias, int R, int N, int K, int act,
                      const float* res, long long ld_res, float* C, long long ldc, cudaStream_t st);
// Y = LN(X + add) * g + b  (add may be null; add_rows_mod > 0: add row = r % add_rows_mod), C = 256
int launch_tok_add_ln(const float* X, const float* add, int add_rows_mod, const float* g, const float* b, float eps,
                      int R, int C, float* Y, cudaStream_t st);
// Y = X + P[r % mod]
int launch_tok_add_bcast(const float* X, const float* P, int mod, int R, int C, float* Y, cudaStream_t st);
// self attention of T tokens per image, heads x d (T <= 64), q/k/v/out [B*T, heads*d]
int launch_tok_self_attn(const float* q, const float* k, const float* v, int B, int T, int heads, int d, float* out,
                         cudaStream_t st);
// tokens -> image: q [T, heads*d] (q_img_stride 0 = same for every image) against K/V [B*Nk, ld] fp32
int launch_attn_t2i(const float* q, long long q_img_stride, const float* K, const float* V, long long ld_kv, int B,
                    int T, int Nk, int heads, int d, float* out, float* scratch, cudaStream_t st);
size_t attn_t2i_scratch_floats(int B, int T, int heads, int d);
// image -> tokens: Q [B*Nq, ld_q] fp32, K/V [B*T, heads*d] fp32, out bf16 [B*Nq, heads*d]
int launch_attn_i2t(const float* Q, long long ld_q, const float* K, const float* V, int B, int Nq, int T, int heads,
                    int d, int fp16, __nv_bfloat16* out, cudaStream_t st);

// ---- upscaling: LayerNorm2d (channels-last rows of C=64) + GELU -> bf16
int launch_ln2d_gelu(const float* X, long long rows, int C, const float* g, const float* b, float eps, int fp16,
                     __nv_bfloat16* out, cudaStream_t st);
// ---- masks[b,k,p] = sum_c hyper[b,k,c] * U[(b,p),c]  (k<4, c<32) ; counts[b] = {#(m0 > delta), #(m0 > -delta)}
int launch_mask_product(const float* U, const float* hyper, int B, int P, float delta, float* masks,
                        unsigned int* counts, cudaStream_t st);
// ---- dynamic multimask via stability: sel[b] in 0..3, low_res[b] = masks[b, sel], iou_out[b] = iou[b, sel]
int launch_select_mask(const float* masks, const float* iou, const unsigned int* counts, int B, int P, float thresh,
                       float* low_res, float* iou_out, int* sel, cudaStream_t st);

// ---- tail: bilinear x4 (align_corners=False) of low_res 256^2 -> 1024^2, 4-branch conv refinement (k=3,5,7,11,
// 4 channels each, exact GELU, 1x1 combine), optional fp32 high-res logits, optional threshold (> 0) -> u8 {0,255}
// with per-image extents (min x, min y, max x, max y) of the foreground.
struct RefineWeights {
  const float* w[4];   // [4, k, k]
  const float* b[4];   // [4]
  const float* cw;     // [16]
  float cb;
  int use_refine;
  const float* comp = nullptr;  // [16 phases][25 taps][16 channels] composite low-res stencils ("ref.comp"), optional
};
// src_full = 1: `low_res` is already a [B,1,1024,1024] map (no upsample; MultiKernelRefinement.forward alone)
int launch_tail(const float* low_res, int src_full, int B, const RefineWeights& rw, float* high_res, uint8_t* mask_u8, int* extents,
                cudaStream_t st);
// bilinear resize (align_corners=False) of fp32 [B,1,1024,1024] logits to (H,W) + threshold -> u8 + extents
int launch_resize_threshold(const float* high_res, int B, int S, int H, int W, float* out_logits, uint8_t* mask_u8,
                            int* extents, cudaStream_t st);

// ---- debug tap: counter[0] += number of 16-bit entries equal to +-65504 (a saturated cvt.rn.satfinite.f16 result)
int launch_count_sat16(const uint16_t* p, long long n, unsigned int* counter, cudaStream_t st);

}  // namespace cvb
