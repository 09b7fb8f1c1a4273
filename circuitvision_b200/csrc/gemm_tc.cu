// Host launcher + C-ABI test entry of the tcgen05 GEMM (gemm_tc.cuh).
#include "gemm_tc.cuh"

#include "common.cuh"

namespace cvb {

template <int BN>
static int launch_bn(const __nv_bfloat16* A, long long lda, const __nv_bfloat16* W, long long ldw, int M, int N, int K,
                     const GemmEpilogue& epi, int num_sms, cudaStream_t st) {
  static bool attr_set = false;
  constexpr int smem = gemm_smem_bytes<BN>();
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(k_gemm_tc<BN>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    if (e != cudaSuccess) return cvb_fail_cuda(e, "cudaFuncSetAttribute(k_gemm_tc)");
    attr_set = true;
  }
  CUtensorMap ta, tw;
  if (!tc_host::make_tmap_bf16(&ta, A, (uint64_t)M, (uint64_t)K, (uint64_t)lda, GEMM_BM) ||
      !tc_host::make_tmap_bf16(&tw, W, (uint64_t)N, (uint64_t)K, (uint64_t)ldw, BN))
    return cvb_fail(CV_ERR_CUDA, "cuTensorMapEncodeTiled failed (GEMM operands)");
  GemmProblem p;
  p.M = M; p.N = N; p.K = K;
  p.n_tiles_m = (M + GEMM_BM - 1) / GEMM_BM;
  p.n_tiles_n = (N + BN - 1) / BN;
  long long tiles = (long long)p.n_tiles_m * p.n_tiles_n;
  int grid = (int)(tiles < num_sms ? tiles : num_sms);
  cvb_next_work(2.0 * (double)M * (double)N * (double)K);
  if (cvb_profile_on()) {
    char nm[96];
    snprintf(nm, sizeof(nm), "gemm M%d N%d K%d bn%d%s%s%s", M, N, K, BN, epi.out_bf16 ? " ->bf16" : "", epi.out_f32 ? " ->f32" : "",
             epi.res ? " +res" : "");
    cvb_next_name(nm);
  }
  CVB_LAUNCH((k_gemm_tc<BN>), dim3(grid), dim3(GEMM_THREADS), smem, st, ta, tw, p, epi);
  return CV_OK;
}

int gemm_tc_launch(const __nv_bfloat16* A, long long lda, const __nv_bfloat16* W, long long ldw, int M, int N, int K,
                   const GemmEpilogue& epi, int num_sms, cudaStream_t st) {
  if (M <= 0 || N <= 0 || K <= 0) return cvb_fail(CV_ERR_INVALID, "gemm: non-positive size");
  if ((N % 32) || (K % 8) || (lda % 8) || (ldw % 8)) return cvb_fail(CV_ERR_INVALID, "gemm: N%32, K%8, lda%8, ldw%8 must be 0");
  if (((uintptr_t)A | (uintptr_t)W) & 15) return cvb_fail(CV_ERR_INVALID, "gemm: operands must be 16-byte aligned");
  if (epi.map_mode == GEMM_MAP_SHUFFLE2 && (epi.cout % 32)) return cvb_fail(CV_ERR_INVALID, "gemm: shuffle needs cout%32==0");
  if (N % 192 == 0) return launch_bn<192>(A, lda, W, ldw, M, N, K, epi, num_sms, st);
  if (N % 128 == 0) return launch_bn<128>(A, lda, W, ldw, M, N, K, epi, num_sms, st);
  if (N % 96 == 0) return launch_bn<96>(A, lda, W, ldw, M, N, K, epi, num_sms, st);
  if (N % 64 == 0) return launch_bn<64>(A, lda, W, ldw, M, N, K, epi, num_sms, st);
  return launch_bn<32>(A, lda, W, ldw, M, N, K, epi, num_sms, st);
}

int device_sm_count() {
  static int n = 0;
  if (!n) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
    if (n <= 0) n = 148;
  }
  return n;
}

}  // namespace cvb

using namespace cvb;

extern "C" int cv_gemm_bf16(const void* A, long long lda, const void* W, long long ldw, int M, int N, int K,
                            const float* bias, int act, const float* residual, long long ld_res, float* out_f32,
                            long long ld_f32, void* out_bf16, long long ld_bf16, void* stream) {
  cvb_reset_launches();
  if (!A || !W || (!out_f32 && !out_bf16)) return cvb_fail(CV_ERR_INVALID, "cv_gemm_bf16: null pointer");
  GemmEpilogue e;
  e.bias = bias;
  e.act = act;
  e.res = residual;
  e.ld_res = ld_res;
  e.out_f32 = out_f32;
  e.ld_f32 = ld_f32;
  e.out_bf16 = (__nv_bfloat16*)out_bf16;
  e.ld_bf16 = ld_bf16;
  return gemm_tc_launch((const __nv_bfloat16*)A, lda, (const __nv_bfloat16*)W, ldw, M, N, K, e, device_sm_count(),
                        (cudaStream_t)stream);
}
