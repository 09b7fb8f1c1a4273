// tcgen05.mma issue/throughput probe (gpurun only): how many SM cycles does one UTCHMMA of shape 128 x N x 16 cost when
// the issuing thread keeps the tensor pipe's queue full?  Variants: N, SS vs TS (A from TMEM), one dependent accumulator
// chain vs two alternating accumulators, cta_group 1 vs 2.  Operands are zeros (only the timing matters).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I circuitvision_b200/csrc -o /tmp/mma_probe scripts/mma_probe.cu -lcuda
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "tc05.cuh"

constexpr int ITERS = 65536;  // MMAs per CTA

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

// SS N=128 MMAs with a tcgen05.commit after every GROUP MMAs (to a barrier nobody waits on): does commit stall the issuer?
// Also reports the cycles the issuing thread itself needed to get through the loop (issue time), not only the completion time.
template <int GROUP>
__global__ void __launch_bounds__(128, 1) k_probe_commit(unsigned long long* cycles, unsigned long long* issue_cycles) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  uint8_t* sA = smem;
  uint8_t* sB = smem + 16384;
  uint64_t* bar = (uint64_t*)(smem + 16384 + 32768);
  uint32_t* slot = (uint32_t*)(bar + 2);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int i = threadIdx.x; i < (16384 + 32768) / 16; i += 128) ((uint4*)smem)[i] = make_uint4(0, 0, 0, 0);
  if (threadIdx.x == 0) {
    tc::mbar_init(&bar[0], 1);
    tc::mbar_init(&bar[1], 1);
    tc::fence_barrier_init();
  }
  tc::fence_proxy_async_smem();
  if (warp == 0) tc::tmem_alloc<512>(slot);
  tc::tc_fence_before();
  __syncthreads();
  tc::tc_fence_after();
  const uint32_t tmem = *slot;
  if (warp == 1) {
    const uint32_t idesc = tc::idesc_bf16(128, 128, false, false, true);
    const uint64_t da = tc::desc_kmajor(tc::smem_u32(sA)), db = tc::desc_kmajor(tc::smem_u32(sB));
    const long long t0 = clock64();
    for (int it = 0; it < ITERS / GROUP; it++) {
      if (tc::elect_one()) {
#pragma unroll
        for (int k = 0; k < GROUP; k++) tc::mma_f16_ss(tmem, da + 2 * (k & 3), db + 2 * (k & 3), idesc, 1u);
        tc::mma_commit(&bar[1]);
      }
      __syncwarp();
    }
    const long long ti = clock64();
    if (tc::elect_one()) tc::mma_commit(&bar[0]);
    __syncwarp();
    tc::mbar_wait(&bar[0], 0);
    const long long t1 = clock64();
    if (lane == 0) {
      cycles[blockIdx.x] = (unsigned long long)(t1 - t0);
      issue_cycles[blockIdx.x] = (unsigned long long)(ti - t0);
    }
  }
  tc::tc_fence_before();
  __syncthreads();
  if (warp == 0) {
    tc::tc_fence_after();
    tc::tmem_dealloc<512>(tmem);
  }
}

template <int GROUP>
static void run_commit(const char* name) {
  unsigned long long* d;
  cudaMalloc(&d, 2 * 148 * 8);
  auto kern = k_probe_commit<GROUP>;
  const int smem = 1024 + 16384 + 32768 + 256;
  cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  for (int rep = 0; rep < 2; rep++) {
    kern<<<148, 128, smem>>>(d, d + 148);
    cudaError_t err = cudaDeviceSynchronize();
    if (err != cudaSuccess) {
      printf("%-34s FAILED: %s\n", name, cudaGetErrorString(err));
      return;
    }
  }
  std::vector<unsigned long long> h(296);
  cudaMemcpy(h.data(), d, 296 * 8, cudaMemcpyDeviceToHost);
  printf("%-34s %7.1f cycles/MMA to completion, %7.1f cycles/MMA of issue time (SM 0)\n", name, (double)h[0] / ITERS, (double)h[148] / ITERS);
  cudaFree(d);
}

// attention-like mix: per group 6 SS MMAs 128x128x16 (K-major B) into S + 8 TS MMAs 128x96x16 (A from TMEM, MN-major B) into O
template <int MNB, int RND = 0>
__global__ void __launch_bounds__(128, 1) k_probe_attn(unsigned long long* cycles) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  uint8_t* sQ = smem;               // 32 KB
  uint8_t* sK = smem + 32768;       // 32 KB
  uint8_t* sV = smem + 65536;       // 32 KB
  uint64_t* bar = (uint64_t*)(smem + 98304);
  uint32_t* slot = (uint32_t*)(bar + 2);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int i = threadIdx.x; i < 98304 / 16; i += 128) {
    uint32_t h = (uint32_t)i * 2654435761u + blockIdx.x * 40503u;
    // random fp16 values in (-2, 2): sign | exponent 12..15 | random mantissa
    auto rnd = [&]() { h = h * 1664525u + 1013904223u; return ((h >> 8) & 0x83FFu) | (((h >> 20) & 3u) + 12u) << 10; };
    uint4 v = make_uint4(0, 0, 0, 0);
    if (RND) v = make_uint4(rnd() | rnd() << 16, rnd() | rnd() << 16, rnd() | rnd() << 16, rnd() | rnd() << 16);
    ((uint4*)smem)[i] = v;
  }
  if (threadIdx.x == 0) {
    tc::mbar_init(&bar[0], 1);
    tc::fence_barrier_init();
  }
  tc::fence_proxy_async_smem();
  if (warp == 0) tc::tmem_alloc<512>(slot);
  tc::tc_fence_before();
  __syncthreads();
  tc::tc_fence_after();
  const uint32_t tmem = *slot;
  if (RND) {  // random P (A operand of the TS MMAs) in the S columns
    uint32_t w[32];
    uint32_t h = threadIdx.x * 747796405u + 12345u;
    for (int c = 0; c < 8; c++) {
#pragma unroll
      for (int i = 0; i < 32; i++) { h = h * 1664525u + 1013904223u; w[i] = (h & 0x03FF03FFu) | 0x38003800u; }
      tc::tmem_st_32x32(tmem + ((uint32_t)(warp * 32) << 16) + c * 32, w);
    }
    tc::tmem_st_wait();
    tc::tc_fence_before();
    __syncthreads();
    tc::tc_fence_after();
  }
  if (warp == 1) {
    const uint32_t idesc_qk = tc::idesc_bf16(128, 128, false, false, true);
    const uint32_t idesc_pv = tc::idesc_bf16(128, 96, false, MNB != 0, true);
    const uint64_t dq = tc::desc_kmajor(tc::smem_u32(sQ)), dk = tc::desc_kmajor(tc::smem_u32(sK));
    const uint64_t dv = MNB ? tc::smem_desc_sw128(tc::smem_u32(sV), 128 * 128, 1024) : tc::desc_kmajor(tc::smem_u32(sV));
    const long long t0 = clock64();
    for (int it = 0; it < ITERS / 16; it++) {
      const int g = it & 1;
      if (tc::elect_one()) {
#pragma unroll
        for (int k = 0; k < 6; k++) {
          const uint32_t off = ((k >> 2) * (128 * 128) + (k & 3) * 32) >> 4;
          tc::mma_f16_ss(tmem + g * 128, dq + off, dk + off, idesc_qk, k > 0 ? 1u : 0u);
        }
#pragma unroll
        for (int k = 0; k < 8; k++)
          tc::mma_f16_ts(tmem + 256 + g * 96, tmem + (g ^ 1) * 128 + k * 8, MNB ? dv + (k * 2048 >> 4) : dv + 2 * (k & 3), idesc_pv, 1u);
      }
      __syncwarp();
    }
    if (tc::elect_one()) tc::mma_commit(&bar[0]);
    __syncwarp();
    tc::mbar_wait(&bar[0], 0);
    const long long t1 = clock64();
    if (lane == 0) cycles[blockIdx.x] = (unsigned long long)(t1 - t0);
  }
  tc::tc_fence_before();
  __syncthreads();
  if (warp == 0) {
    tc::tc_fence_after();
    tc::tmem_dealloc<512>(tmem);
  }
}

// the same mix while NW other warps stream tcgen05.ld (+ optional st) over other TMEM columns, as softmax / epilogue warps do
// WITH_ST == 2: the other warps run an arithmetic loop (FFMA + MUFU.EX2 + conversions, no TMEM access) instead
template <int NW, int WITH_ST>
__global__ void __launch_bounds__(128 + 32 * NW, 1) k_probe_attn_ld(unsigned long long* cycles, unsigned int* sink) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  uint8_t* sQ = smem;
  uint8_t* sK = smem + 32768;
  uint8_t* sV = smem + 65536;
  uint64_t* bar = (uint64_t*)(smem + 98304);
  uint32_t* slot = (uint32_t*)(bar + 2);
  volatile int* stop = (volatile int*)(bar + 4);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int i = threadIdx.x; i < 98304 / 16; i += blockDim.x) ((uint4*)smem)[i] = make_uint4(0, 0, 0, 0);
  if (threadIdx.x == 0) {
    tc::mbar_init(&bar[0], 1);
    tc::fence_barrier_init();
    *stop = 0;
  }
  tc::fence_proxy_async_smem();
  if (warp == 0) tc::tmem_alloc<512>(slot);
  tc::tc_fence_before();
  __syncthreads();
  tc::tc_fence_after();
  const uint32_t tmem = *slot;
  if (warp == 1) {
    const uint32_t idesc_qk = tc::idesc_bf16(128, 128, false, false, true);
    const uint32_t idesc_pv = tc::idesc_bf16(128, 96, false, true, true);
    const uint64_t dq = tc::desc_kmajor(tc::smem_u32(sQ)), dk = tc::desc_kmajor(tc::smem_u32(sK));
    const uint64_t dv = tc::smem_desc_sw128(tc::smem_u32(sV), 128 * 128, 1024);
    const long long t0 = clock64();
    for (int it = 0; it < ITERS / 16; it++) {
      const int g = it & 1;
      if (tc::elect_one()) {
#pragma unroll
        for (int k = 0; k < 6; k++) {
          const uint32_t off = ((k >> 2) * (128 * 128) + (k & 3) * 32) >> 4;
          tc::mma_f16_ss(tmem + g * 128, dq + off, dk + off, idesc_qk, k > 0 ? 1u : 0u);
        }
#pragma unroll
        for (int k = 0; k < 8; k++)
          tc::mma_f16_ts(tmem + 256 + g * 96, tmem + (g ^ 1) * 128 + k * 8, dv + (k * 2048 >> 4), idesc_pv, 1u);
      }
      __syncwarp();
    }
    if (tc::elect_one()) tc::mma_commit(&bar[0]);
    __syncwarp();
    tc::mbar_wait(&bar[0], 0);
    const long long t1 = clock64();
    if (lane == 0) {
      cycles[blockIdx.x] = (unsigned long long)(t1 - t0);
      *stop = 1;
    }
  } else if (warp >= 4) {
    const uint32_t base = tmem + ((uint32_t)((warp & 3) * 32) << 16) + 448;  // columns 448..511: nobody else uses them
    unsigned int acc = 0;
    unsigned long long n = 0;
    if (WITH_ST == 2) {
      float x[16];
#pragma unroll
      for (int i = 0; i < 16; i++) x[i] = (float)(lane + i) * 0.01f;
      while (!*stop) {
#pragma unroll
        for (int r = 0; r < 8; r++) {
#pragma unroll
          for (int i = 0; i < 16; i++) {
            float y;
            asm volatile("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(fmaf(x[i], 0.999f, -0.5f)));
            x[i] = y * 0.5f + x[(i + 1) & 15] * 0.25f;
          }
        }
        n++;
      }
#pragma unroll
      for (int i = 0; i < 16; i++) acc += __float_as_uint(x[i]);
    }
    while (!*stop) {
      uint32_t v[32];
      tc::tmem_ld_32x32(base + ((n & 1) ? 32 : 0), v);
      tc::tmem_ld_wait();
#pragma unroll
      for (int i = 0; i < 32; i++) acc += v[i];
      if (WITH_ST) {
        uint32_t w[16];
#pragma unroll
        for (int i = 0; i < 16; i++) w[i] = acc + i;
        tc::tmem_st_32x16(base + ((n & 1) ? 32 : 0), w);
        tc::tmem_st_wait();
      }
      n++;
    }
    if (lane == 0) {
      sink[blockIdx.x * 32 + warp] = acc;
      if (warp == 4) cycles[148 + blockIdx.x] = n;
    }
  }
  tc::tc_fence_before();
  __syncthreads();
  if (warp == 0) {
    tc::tc_fence_after();
    tc::tmem_dealloc<512>(tmem);
  }
}

template <int NW, int WITH_ST>
static void run_attn_ld(const char* name) {
  unsigned long long* d;
  unsigned int* sink;
  cudaMalloc(&d, 2 * 148 * 8);
  cudaMalloc(&sink, 148 * 32 * 4);
  auto kern = k_probe_attn_ld<NW, WITH_ST>;
  const int smem = 1024 + 98304 + 256;
  cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  for (int rep = 0; rep < 2; rep++) {
    kern<<<148, 128 + 32 * NW, smem>>>(d, sink);
    cudaError_t err = cudaDeviceSynchronize();
    if (err != cudaSuccess) {
      printf("%-34s FAILED: %s\n", name, cudaGetErrorString(err));
      return;
    }
  }
  std::vector<unsigned long long> h(296);
  cudaMemcpy(h.data(), d, 296 * 8, cudaMemcpyDeviceToHost);
  unsigned long long mx = 0;
  for (int i = 0; i < 148; i++) mx = h[i] > mx ? h[i] : mx;
  printf("%-34s %7.1f cycles per group (nominal 768); one streaming warp did %.1f ld rounds per group\n", name, (double)mx / (ITERS / 16),
         (double)h[148] / (ITERS / 16));
  cudaFree(d);
  cudaFree(sink);
}

template <int MNB, int RND = 0>
static void run_attn(const char* name) {
  unsigned long long* d;
  cudaMalloc(&d, 148 * 8);
  auto kern = k_probe_attn<MNB, RND>;
  const int smem = 1024 + 98304 + 256;
  cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  for (int rep = 0; rep < 6; rep++) {
    if (rep == 1) cudaEventRecord(e0);
    kern<<<148, 128, smem>>>(d);
  }
  cudaEventRecord(e1);
  cudaError_t err = cudaDeviceSynchronize();
  if (err != cudaSuccess) {
    printf("%-34s FAILED: %s\n", name, cudaGetErrorString(err));
    return;
  }
  float ms = 0;
  cudaEventElapsedTime(&ms, e0, e1);
  std::vector<unsigned long long> h(148);
  cudaMemcpy(h.data(), d, 148 * 8, cudaMemcpyDeviceToHost);
  unsigned long long mx = 0;
  for (auto v : h) mx = v > mx ? v : mx;
  printf("%-34s %7.1f cycles per (6 QK + 8 PV) group (nominal 768); %.3f ms per launch -> %.2f GHz effective\n", name,
         (double)mx / (ITERS / 16), ms / 5, (double)mx / (ms / 5 * 1e6));
  cudaFree(d);
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
  run_commit<4>("SS N=128, commit every 4 MMAs");
  run_commit<16>("SS N=128, commit every 16 MMAs");
  run_commit<64>("SS N=128, commit every 64 MMAs");
  run_attn_ld<4, 0>("attention mix + 4 warps tcgen05.ld");
  run_attn_ld<8, 0>("attention mix + 8 warps tcgen05.ld");
  run_attn_ld<16, 0>("attention mix + 16 warps tcgen05.ld");
  run_attn_ld<8, 2>("attention mix + 8 warps FMA/MUFU");
  run_attn_ld<16, 2>("attention mix + 16 warps FMA/MUFU");
  run_attn_ld<8, 1>("attention mix + 8 warps ld+st");
  run_attn_ld<16, 1>("attention mix + 16 warps ld+st");
  run_attn<0>("attention mix, V K-major");
  run_attn<1>("attention mix, V MN-major");
  run_attn<1, 1>("attention mix, random operands");
  return 0;
}
