// Fused Hiera MLP half-block on tcgen05 + TMA (interface; kernel in mlp_fused.cu):
//
//     X  <-  X + fc2( GELU( fc1( LayerNorm(X) ) ) )                 (sam2 MultiScaleBlock: x + mlp(norm2(x)))
//
// for one 128-row tile at a time, with the LayerNorm'd 16-bit operand, the 4C-wide hidden activation and the fc2
// accumulator never leaving the SM: HBM sees one fp32 read and one fp32 write of the residual stream (plus the 16-bit
// weights, streamed from L2), instead of LN write/read + hidden write/read + residual read/write of the unfused chain.
#pragma once
#include <cuda_bf16.h>
#include <cuda_runtime.h>

namespace cvb {

struct MlpFusedArgs {
  float* X = nullptr;                  // fp32 residual stream [M, C], updated in place
  int M = 0, C = 0;
  const float* gamma = nullptr;        // norm2 weight / bias [C]
  const float* beta = nullptr;
  float eps = 1e-6f;
  const __nv_bfloat16* W1 = nullptr;   // fc1 weight [4C, C], 16-bit, K-major
  const float* b1 = nullptr;           // [4C]
  const __nv_bfloat16* W2 = nullptr;   // fc2 weight [C, 4C], 16-bit, K-major
  const float* b2 = nullptr;           // [C]
  int fp16 = 0;                        // operand format: 0 bf16, 1 IEEE half
  unsigned int* sat_counter = nullptr; // debug: counts hidden activations that saturated in the fp16 conversion
  unsigned long long* trace = nullptr; // debug: CTA 0 appends (event id << 48 | %globaltimer ns) records, trace[0] = count
};

// true when a fused instantiation exists for this width (the caller falls back to LN + fc1 + fc2 GEMMs otherwise)
bool mlp_fused_supported(int C);
int mlp_fused_launch(const MlpFusedArgs& a, int num_sms, cudaStream_t st);

}  // namespace cvb
