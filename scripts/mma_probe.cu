// tcgen05.mma issue/throughput probe (gpurun only): how many SM cycles does one UTCHMMA of shape 128 x N x 16 cost when
// the issuing thread keeps the tensor pipe's queue full?  Variants: N, SS vs TS (A from TMEM), one dependent accumulator
// chain vs two alternating accumulators, cta_group 1 vs 2.  Operands are zeros (only the timing matters).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I circuitvision_b200/csrc -o /tmp/mma_probe scripts/mma_probe.cu -lcuda
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "tc05.cuh"

constexpr int ITERS = 4096;  // MMAs per CTA

template <int N, int TS, int ALT, int CG>
__global__ void __launch_bounds__(128, 1) k_probe(unsigned long long* cycles) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  uint8_t* sA = smem;               // 128 x 64 16-bit, K-major swizzled (16 KB)
  uint8_t* sB = smem + 16384;       // up to 256 x 64 (32 KB)
  uint64_t* bar = (uint64_t*)(smem + 16384 + 32768);
  uint32_t* slot = (uint32_t*)(bar + 2);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int i = threadIdx.x; i < (16384 + 32768) / 16; i += 128) ((uint4*)smem)[i] = make_uint4(0, 0, 0, 0);
  if (threadIdx.x == 0) {
    tc::mbar_init(&bar[0], 1);
    tc::fence_barrier_init();
  }
  tc::fence_proxy_async_smem();
  if (warp == 0) {
    if (CG == 2) tc::tmem_alloc2<512>(slot);
    else tc::tmem_alloc<512>(slot);
  }
  tc::tc_fence_before();
  __syncthreads();
  if (CG == 2) tc::cluster_sync();
  tc::tc_fence_after();
  const uint32_t tmem = *slot;
  const int rank = CG == 2 ? (int)tc::cluster_ctarank() : 0;
  long long t0 = 0, t1 = 0;
  if (warp == 1 && rank == 0) {
    const uint32_t idesc = tc::idesc_bf16(128 * CG, N, false, false, true);
    const uint64_t da = tc::desc_kmajor(tc::smem_u32(sA)), db = tc::desc_kmajor(tc::smem_u32(sB));
    t0 = clock64();
    for (int it = 0; it < ITERS / 4; it++) {
      if (tc::elect_one()) {
#pragma unroll
        for (int k = 0; k < 4; k++) {
          const uint32_t d = tmem + ((ALT && (k & 1)) ? 256 : 0);
          if (TS) tc::mma_f16_ts(d, tmem + 288 + k * 8, db + 2 * k, idesc, 1u);
          else if (CG == 2) tc::mma_f16_ss2(d, da + 2 * k, db + 2 * k, idesc, 1u);
          else tc::mma_f16_ss(d, da + 2 * k, db + 2 * k, idesc, 1u);
        }
      }
      __syncwarp();
    }
    if (tc::elect_one()) {
      if (CG == 2) tc::mma_commit2(&bar[0]);
      else tc::mma_commit(&bar[0]);
    }
    __syncwarp();
    tc::mbar_wait(&bar[0], 0);
    t1 = clock64();
    if (lane == 0) cycles[blockIdx.x] = (unsigned long long)(t1 - t0);
  }
  tc::tc_fence_before();
  __syncthreads();
  if (CG == 2) tc::cluster_sync();
  if (warp == 0) {
    tc::tc_fence_after();
    if (CG == 2) tc::tmem_dealloc2<512>(tmem);
    else tc::tmem_dealloc<512>(tmem);
  }
}

template <int N, int TS, int ALT, int CG>
static void run(const char* name) {
  unsigned long long* d;
  cudaMalloc(&d, 148 * 8);
  cudaMemset(d, 0, 148 * 8);
  auto kern = k_probe<N, TS, ALT, CG>;
  const int smem = 1024 + 16384 + 32768 + 256;
  cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  for (int rep = 0; rep < 2; rep++) {
    cudaEventRecord(e0);
    if (CG == 2) {
      cudaLaunchConfig_t cfg = {};
      cfg.gridDim = dim3(148);
      cfg.blockDim = dim3(128);
      cfg.dynamicSmemBytes = smem;
      cudaLaunchAttribute attr[1];
      attr[0].id = cudaLaunchAttributeClusterDimension;
      attr[0].val.clusterDim.x = 2;
      attr[0].val.clusterDim.y = 1;
      attr[0].val.clusterDim.z = 1;
      cfg.attrs = attr;
      cfg.numAttrs = 1;
      cudaLaunchKernelEx(&cfg, kern, d);
    } else {
      kern<<<148, 128, smem>>>(d);
    }
    cudaEventRecord(e1);
    cudaError_t err = cudaDeviceSynchronize();
    if (err != cudaSuccess) {
      printf("%-34s FAILED: %s\n", name, cudaGetErrorString(err));
      return;
    }
  }
  float ms = 0;
  cudaEventElapsedTime(&ms, e0, e1);
  std::vector<unsigned long long> h(148);
  cudaMemcpy(h.data(), d, 148 * 8, cudaMemcpyDeviceToHost);
  unsigned long long mx = 0;
  for (auto v : h) mx = v > mx ? v : mx;
  const double cyc = (double)mx / ITERS;
  const double nominal = (double)N * CG / 2.0 / CG;  // 128 x N x 16 MACs at 4096 MAC/clk/SM = N / 2 cycles (per SM)
  const double flops = 2.0 * 128 * CG * N * 16 * ITERS * (148 / CG);
  printf("%-34s %7.1f cycles/MMA (nominal %5.1f)  kernel %.3f ms -> %7.1f TFLOP/s\n", name, cyc, nominal, ms, flops / ms / 1e9);
  cudaFree(d);
}

int main() {
  run<64, 0, 0, 1>("SS N=64  one chain");
  run<96, 0, 0, 1>("SS N=96  one chain");
  run<128, 0, 0, 1>("SS N=128 one chain");
  run<192, 0, 0, 1>("SS N=192 one chain");
  run<256, 0, 0, 1>("SS N=256 one chain");
  run<64, 0, 1, 1>("SS N=64  two accumulators");
  run<128, 0, 1, 1>("SS N=128 two accumulators");
  run<256, 0, 1, 1>("SS N=256 two accumulators");
  run<64, 1, 0, 1>("TS N=64  one chain");
  run<96, 1, 0, 1>("TS N=96  one chain");
  run<128, 1, 0, 1>("TS N=128 one chain");
  run<192, 1, 0, 1>("TS N=192 one chain");
  run<128, 0, 0, 2>("SS N=128 cta_group::2 (M=256)");
  run<256, 0, 0, 2>("SS N=256 cta_group::2 (M=256)");
  return 0;
}
