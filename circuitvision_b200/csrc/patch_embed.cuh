// Fused patch embedding of the Hiera trunk (reference: sam2 PatchEmbed, Conv2d(3, E, 7, stride 4, padding 3), reached from
// src/sam2_infer.py:226 `image_encoder(x)`): raw uint8 pixels -> conv as a tcgen05 GEMM -> + folded bias / positional embedding
// -> fp32 residual stream X0, and optionally LayerNorm(norm1 of block 0) -> 16-bit operand in 8 x 8 window-major row order.
#pragma once
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>

namespace cvb {

struct PatchEmbedArgs {
  const uint8_t* img;          // [B, S, S, 3] uint8, HWC
  int B, S;                    // S % 512 == 0 (128-token tiles along x)
  int swap_rb;                 // BGR input
  int fp16;                    // operand format of W and of the normalised output (1 = fp16, 0 = bf16)
  const __nv_bfloat16* W;      // [E, 168] 16-bit, K order (ky, kx * 3 + c) padded 21 -> 24 per kernel row ("pe.w8")
  int E;                       // 96 / 112 / 144 (E % 16 == 0, E <= 160)
  const float* pos;            // [G * G, E] fp32: positional embedding + bias - mean term ("pos8"), G = S / 4
  float* X0;                   // [B * G * G, E] fp32 out
  // optional: norm1 of block 0 on the fly.  A16 [B * G * G, E] 16-bit, rows in 8 x 8 window-major order
  const float* gamma;
  const float* beta;
  float eps;
  __nv_bfloat16* A16;          // nullptr = no LayerNorm output
  unsigned int* sat_counter;   // fp16 saturation counter (debug), may be nullptr
};

bool patch_embed_supported(int E, int S);
int patch_embed_launch(const PatchEmbedArgs& a, int num_sms, cudaStream_t st);

}  // namespace cvb
