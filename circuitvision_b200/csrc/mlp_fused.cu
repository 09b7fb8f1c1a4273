// Fused LayerNorm -> fc1 -> GELU -> fc2 -> +residual for one Hiera MLP half-block (interface: mlp_fused.cuh).
//
// Persistent CTA of 16 warps per SM; a work item is a 128-row tile of the fp32 residual stream X [M, C]:
//
//   warps 12-15  LayerNorm producers: read the fp32 rows, normalise, write the 16-bit A operand [128, C] into shared memory
//                in the K-major 128-byte-swizzled layout tcgen05 reads (double-buffered across tiles when it fits)
//   warp 0       TMA producer: streams W1 / W2 boxes ([rows, 64] 16-bit, 128B swizzle) from L2 through a ring
//   warp 1       MMA issuer (one lane).  The hidden dimension 4C is cut into chunks of HC columns; per chunk g
//                    fc1(g):  S[g&1]  = A * W1[chunk g]^T            (SS MMA, fp32 accumulators in TMEM, 128 x HC)
//                    fc2(g):  Y      += H[g&1] * W2[:, chunk g]^T     (TS MMA: A operand = H read from TMEM)
//                issued as fc1(g), fc2(g-1), fc1(g+1), ... across tile boundaries, so the tensor pipe works on the next chunk
//                while the epilogue warps turn S into H
//   warps 4-11   epilogue: per chunk  tcgen05.ld S -> +b1 -> GELU -> 16-bit -> tcgen05.st H over the head of the same TMEM
//                columns (a warp only overwrites columns it has already loaded); per tile  tcgen05.ld Y -> +b2 -> + fp32
//                residual -> swizzled staging -> TMA store of the fp32 rows (in place)
//
// TMEM: Y at column 0 (two buffers for C = 96), S/H double buffer above it; 512 columns allocated.
// The hidden activation and the normalised operand never reach HBM; HBM sees X once in, once out.
#include "mlp_fused.cuh"

#include <stdlib.h>

#include "act.cuh"
#include "common.cuh"
#include "tc05.cuh"

namespace cvb {

constexpr int MLP_SMEM_MAX = 232448;
constexpr int MLP_THREADS = 512;

template <int C_>
struct MlpCfg {
  static constexpr int C = C_;
  static constexpr int HC = (C == 96 || C == 192) ? 128 : 64;      // hidden columns per chunk (= N of fc1, K of fc2)
  static constexpr int NY = (C == 96) ? 2 : 1;                      // Y accumulator buffers
  static constexpr int NCH = 4 * C / HC;                            // chunks per tile
  static constexpr int NKB1 = (C + 63) / 64;                        // 64-wide K-blocks of fc1
  static constexpr int KS1_LAST = (C - (NKB1 - 1) * 64) / 16;       // 16-wide k-steps in the last K-block
  static constexpr int NSPLIT = C > 256 ? 2 : 1;                    // fc2 N = C is issued as NSPLIT MMAs of N2 columns
  static constexpr int N2 = C / NSPLIT;
  static constexpr int NKB2 = HC / 64;                              // K-blocks of fc2 per chunk
  static constexpr int UNIT_ROWS = HC > N2 ? HC : N2;               // rows of the largest weight box
  static constexpr int STAGE_BYTES = UNIT_ROWS * 128;
  static constexpr int A_BYTES = NKB1 * 128 * 128;
  static constexpr int OUT_BYTES = 8 * 4096;                        // one 32x32 fp32 staging box per epilogue warp
  static constexpr int FIXED = 1024 + 512 + OUT_BYTES;
  static constexpr int NA = ((MLP_SMEM_MAX - FIXED - 2 * A_BYTES) / STAGE_BYTES >= 4) ? 2 : 1;
  static constexpr int NSTAGES_RAW = (MLP_SMEM_MAX - FIXED - NA * A_BYTES) / STAGE_BYTES;
  static constexpr int NSTAGES = NSTAGES_RAW > 8 ? 8 : NSTAGES_RAW;
  static constexpr int SMEM = FIXED + NA * A_BYTES + NSTAGES * STAGE_BYTES;
  static constexpr int SBASE = ((NY * C + 63) / 64) * 64;           // first TMEM column of the S/H double buffer
  // LayerNorm producers: LN_L lanes per row, LN_V float4 per lane (C = 4 * LN_L * LN_V)
  static constexpr int LN_L = (C == 96 || C == 224 || C == 288) ? 8 : (C == 192) ? 16 : (C == 384) ? 32 : 4;
  static constexpr int LN_V = C / (4 * LN_L);
  static_assert(4 * C % HC == 0, "chunking");
  static_assert(C % 16 == 0 && N2 % 16 == 0 && N2 <= 256, "MMA shapes");
  static_assert(SBASE + 2 * HC <= 512, "TMEM budget");
  static_assert(NSTAGES >= 3, "weight ring");
  static_assert(4 * LN_L * LN_V == C, "LayerNorm lane layout");
  static_assert(STAGE_BYTES % 1024 == 0 && A_BYTES % 1024 == 0, "swizzle atoms");
};

// With a single A buffer the first fc1 of a tile has to wait for the LayerNorm producers, which in turn wait for the last
// fc1 of the previous tile: the pending fc2 is issued first so that the tensor pipe is not idle behind that wait.
template <class K>
__device__ __forceinline__ bool fc2_first(int j) { return K::NA == 1 && j == 0; }

template <int C_>
__global__ void __launch_bounds__(MLP_THREADS, 1)
k_mlp_fused(const __grid_constant__ CUtensorMap tmap_w1, const __grid_constant__ CUtensorMap tmap_w2,
            const __grid_constant__ CUtensorMap tmap_out, MlpFusedArgs p) {
  using K = MlpCfg<C_>;
  constexpr int C = K::C, HC = K::HC;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  uint8_t* sA = smem;                                  // [NA][NKB1][128 rows][128 B]
  uint8_t* sW = sA + K::NA * K::A_BYTES;               // [NSTAGES][STAGE_BYTES]
  uint8_t* sOut = sW + K::NSTAGES * K::STAGE_BYTES;    // [8 warps][4096]
  uint64_t* bars = (uint64_t*)(sOut + K::OUT_BYTES);
  uint64_t* w_full = bars;                    // [NSTAGES]
  uint64_t* w_empty = w_full + K::NSTAGES;    // [NSTAGES]
  uint64_t* a_full = w_empty + K::NSTAGES;    // [2]
  uint64_t* a_empty = a_full + 2;             // [2]
  uint64_t* s_full = a_empty + 2;             // [2]
  uint64_t* h_full = s_full + 2;              // [2]
  uint64_t* y_full = h_full + 2;              // [2]
  uint64_t* y_empty = y_full + 2;             // [2]
  uint32_t* tmem_slot = (uint32_t*)(y_empty + 2);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int n_tiles = (p.M + 127) / 128;
  const int my_tiles = (n_tiles - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;  // tiles blockIdx.x, += gridDim.x
  const int n_chunks = my_tiles * K::NCH;

  if (warp == 0 && lane == 0) {
    tc::prefetch_tmap(&tmap_w1);
    tc::prefetch_tmap(&tmap_w2);
    tc::prefetch_tmap(&tmap_out);
  }
  if (warp == 1 && lane == 0) {
    for (int i = 0; i < K::NSTAGES; i++) {
      tc::mbar_init(&w_full[i], 1);
      tc::mbar_init(&w_empty[i], 1);
    }
    for (int i = 0; i < 2; i++) {
      tc::mbar_init(&a_full[i], 4);
      tc::mbar_init(&a_empty[i], 1);
      tc::mbar_init(&s_full[i], 1);
      tc::mbar_init(&h_full[i], 8);
      tc::mbar_init(&y_full[i], 1);
      tc::mbar_init(&y_empty[i], 8);
    }
    tc::fence_barrier_init();
  }
  if (warp == 2) tc::tmem_alloc<512>(tmem_slot);
  tc::tc_fence_before();
  __syncthreads();
  tc::tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ===================== TMA producer: weight boxes in the order the MMA warp consumes them
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      auto load_w1 = [&](int j) {
        for (int kb = 0; kb < K::NKB1; kb++) {
          tc::mbar_wait(&w_empty[stage], phase ^ 1);
          tc::mbar_arrive_expect_tx(&w_full[stage], HC * 128);
          tc::tma_load_2d(sW + stage * K::STAGE_BYTES, &tmap_w1, &w_full[stage], kb * 64, j * HC);
          if (++stage == K::NSTAGES) { stage = 0; phase ^= 1; }
        }
      };
      auto load_w2 = [&](int j) {
        for (int kb = 0; kb < K::NKB2; kb++)
          for (int sp = 0; sp < K::NSPLIT; sp++) {
            tc::mbar_wait(&w_empty[stage], phase ^ 1);
            tc::mbar_arrive_expect_tx(&w_full[stage], K::N2 * 128);
            tc::tma_load_2d(sW + stage * K::STAGE_BYTES, &tmap_w2, &w_full[stage], j * HC + kb * 64, sp * K::N2);
            if (++stage == K::NSTAGES) { stage = 0; phase ^= 1; }
          }
      };
      for (int g = 0; g <= n_chunks; g++) {
        const int j = g % K::NCH, jp = (g + K::NCH - 1) % K::NCH;
        const bool swap = fc2_first<K>(j);
        if (swap && g >= 1) load_w2(jp);
        if (g < n_chunks) load_w1(j);
        if (!swap && g >= 1) load_w2(jp);
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer
    if (lane == 0) {
      const uint32_t idesc1 = tc::idesc_bf16(128, HC, false, false, p.fp16 != 0);
      const uint32_t idesc2 = tc::idesc_bf16(128, K::N2, false, false, p.fp16 != 0);
      int stage = 0;
      uint32_t phase = 0;
      auto fc1 = [&](int g) {
        const int it = g / K::NCH, j = g - it * K::NCH;
        const int ab = it % K::NA;
        if (j == 0) {
          tc::mbar_wait(&a_full[ab], (uint32_t)(it / K::NA) & 1u);
          tc::tc_fence_after();
        }
        const uint32_t d = tmem_base + K::SBASE + (g & 1) * HC;
        const uint32_t a0 = tc::smem_u32(sA + ab * K::A_BYTES);
        for (int kb = 0; kb < K::NKB1; kb++) {
          tc::mbar_wait(&w_full[stage], phase);
          tc::tc_fence_after();
          const uint32_t b0 = tc::smem_u32(sW + stage * K::STAGE_BYTES);
          const int ks = kb == K::NKB1 - 1 ? K::KS1_LAST : 4;
          for (int k = 0; k < ks; k++)
            tc::mma_f16_ss(d, tc::desc_kmajor(a0 + kb * 16384 + k * 32), tc::desc_kmajor(b0 + k * 32), idesc1,
                           (kb > 0 || k > 0) ? 1u : 0u);
          tc::mma_commit(&w_empty[stage]);
          if (++stage == K::NSTAGES) { stage = 0; phase ^= 1; }
        }
        tc::mma_commit(&s_full[g & 1]);
        if (j == K::NCH - 1) tc::mma_commit(&a_empty[ab]);
      };
      auto fc2 = [&](int g) {
        const int it = g / K::NCH, j = g - it * K::NCH;
        const int yb = it % K::NY;
        if (j == 0) {
          tc::mbar_wait(&y_empty[yb], ((uint32_t)(it / K::NY) & 1u) ^ 1u);
          tc::tc_fence_after();
        }
        tc::mbar_wait(&h_full[g & 1], (uint32_t)(g >> 1) & 1u);
        tc::tc_fence_after();
        const uint32_t hbase = tmem_base + K::SBASE + (g & 1) * HC;
        for (int kb = 0; kb < K::NKB2; kb++)
          for (int sp = 0; sp < K::NSPLIT; sp++) {
            tc::mbar_wait(&w_full[stage], phase);
            tc::tc_fence_after();
            const uint32_t b0 = tc::smem_u32(sW + stage * K::STAGE_BYTES);
            const uint32_t d = tmem_base + yb * C + sp * K::N2;
            for (int k = 0; k < 4; k++) {
              // k-step kk covers hidden elements [16 kk, 16 kk + 16) of the chunk; the epilogue warp pair (column halves)
              // left its 16-bit H at the head of its own half of the S columns: 2 elements per 32-bit column
              const int e = (kb * 4 + k) * 16;
              const int half = e / (HC / 2);
              const uint32_t a = hbase + half * (HC / 2) + (e - half * (HC / 2)) / 2;
              tc::mma_f16_ts(d, a, tc::desc_kmajor(b0 + k * 32), idesc2, (j > 0 || kb > 0 || k > 0) ? 1u : 0u);
            }
            tc::mma_commit(&w_empty[stage]);
            if (++stage == K::NSTAGES) { stage = 0; phase ^= 1; }
          }
        if (j == K::NCH - 1) tc::mma_commit(&y_full[yb]);
      };
      for (int g = 0; g <= n_chunks; g++) {
        const bool swap = fc2_first<K>(g % K::NCH);
        if (swap && g >= 1) fc2(g - 1);
        if (g < n_chunks) fc1(g);
        if (!swap && g >= 1) fc2(g - 1);
      }
    }
  } else if (warp >= 12) {
    // ===================== LayerNorm producers: fp32 rows -> normalised 16-bit A operand (swizzled K-major)
    constexpr int L = K::LN_L, V = K::LN_V, RPW = 32 / L, RPP = 4 * RPW, PASSES = 128 / RPP, U = V <= 3 ? 4 : 1;
    static_assert(PASSES % U == 0, "pass unrolling");
    const int pw = warp - 12, sub = lane % L, rsub = lane / L;
    for (int it = 0; it < my_tiles; it++) {
      const int tile = (int)blockIdx.x + it * (int)gridDim.x;
      const int ab = it % K::NA;
      tc::mbar_wait(&a_empty[ab], ((uint32_t)(it / K::NA) & 1u) ^ 1u);
      uint8_t* A = sA + ab * K::A_BYTES;
      for (int p0 = 0; p0 < PASSES; p0 += U) {
        float4 v[U][V];
        int rows[U];
#pragma unroll
        for (int u = 0; u < U; u++) {
          const int r = (p0 + u) * RPP + pw * RPW + rsub;
          rows[u] = r;
          const long long gr = (long long)tile * 128 + r;
          if (gr < p.M) {
            const float4* xr = (const float4*)(p.X + gr * C);
#pragma unroll
            for (int k = 0; k < V; k++) v[u][k] = xr[sub + k * L];
          } else {
#pragma unroll
            for (int k = 0; k < V; k++) v[u][k] = make_float4(0.f, 0.f, 0.f, 0.f);
          }
        }
#pragma unroll
        for (int u = 0; u < U; u++) {
          float s = 0.f;
#pragma unroll
          for (int k = 0; k < V; k++) s += v[u][k].x + v[u][k].y + v[u][k].z + v[u][k].w;
#pragma unroll
          for (int o = L / 2; o; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
          const float mean = s * (1.0f / C);
          float q = 0.f;
#pragma unroll
          for (int k = 0; k < V; k++) {
            const float a = v[u][k].x - mean, b = v[u][k].y - mean, c = v[u][k].z - mean, d = v[u][k].w - mean;
            q += a * a + b * b + c * c + d * d;
          }
#pragma unroll
          for (int o = L / 2; o; o >>= 1) q += __shfl_xor_sync(0xffffffffu, q, o);
          const float rstd = rsqrtf(q * (1.0f / C) + p.eps);
          const int r = rows[u];
#pragma unroll
          for (int k = 0; k < V; k++) {
            const int i4 = sub + k * L;  // float4 index within the row; columns 4 i4 .. 4 i4 + 3
            const float4 g = __ldg((const float4*)p.gamma + i4), bb = __ldg((const float4*)p.beta + i4);
            const float o0 = (v[u][k].x - mean) * rstd * g.x + bb.x, o1 = (v[u][k].y - mean) * rstd * g.y + bb.y;
            const float o2 = (v[u][k].z - mean) * rstd * g.z + bb.z, o3 = (v[u][k].w - mean) * rstd * g.w + bb.w;
            const int col = i4 * 4, kb = col >> 6, ch = (col & 63) >> 3, hf = (col & 7) >> 2;
            *(uint2*)(A + kb * 16384 + r * 128 + (((ch ^ (r & 7)) << 4) | (hf << 3))) =
                make_uint2(tc::pack16(p.fp16, o0, o1), tc::pack16(p.fp16, o2, o3));
          }
        }
      }
      tc::fence_proxy_async_smem();  // generic-proxy writes -> visible to the tensor core's async-proxy reads
      __syncwarp();
      if (lane == 0) tc::mbar_arrive(&a_full[ab]);
    }
  } else if (warp >= 4) {
    // ===================== epilogue warps: quadrant = warp % 4 (TMEM lanes), column half = (warp - 4) / 4
    const int quad = warp & 3, half = (warp - 4) >> 2;
    const uint32_t lane_addr = tmem_base + ((uint32_t)(quad * 32) << 16);
    uint8_t* buf = sOut + (warp - 4) * 4096;
    unsigned int nsat = 0;
    for (int it = 0; it < my_tiles; it++) {
      const int tile = (int)blockIdx.x + it * (int)gridDim.x;
      for (int j = 0; j < K::NCH; j++) {
        const int g = it * K::NCH + j;
        tc::mbar_wait(&s_full[g & 1], (uint32_t)(g >> 1) & 1u);
        tc::tc_fence_after();
        const uint32_t sbase = lane_addr + K::SBASE + (g & 1) * HC + half * (HC / 2);
#pragma unroll
        for (int i = 0; i < HC / 64; i++) {
          uint32_t v[32];
          tc::tmem_ld_32x32(sbase + 32 * i, v);
          const float4* bp = (const float4*)(p.b1 + j * HC + half * (HC / 2) + 32 * i);
          float4 bv[8];
#pragma unroll
          for (int q = 0; q < 8; q++) bv[q] = __ldg(bp + q);
          tc::tmem_ld_wait();
          float f[32];
#pragma unroll
          for (int q = 0; q < 8; q++) {
            up2(add2(pk2(__uint_as_float(v[4 * q + 0]), __uint_as_float(v[4 * q + 1])), pk2(bv[q].x, bv[q].y)), f[4 * q + 0], f[4 * q + 1]);
            up2(add2(pk2(__uint_as_float(v[4 * q + 2]), __uint_as_float(v[4 * q + 3])), pk2(bv[q].z, bv[q].w)), f[4 * q + 2], f[4 * q + 3]);
          }
#pragma unroll
          for (int q = 0; q < 32; q += 2) gelu_tanh2(f[q], f[q + 1]);
          uint32_t pk[16];
          if (p.fp16) {
#pragma unroll
            for (int q = 0; q < 16; q++) pk[q] = tc::pack16(1, f[2 * q], f[2 * q + 1]);
          } else {
#pragma unroll
            for (int q = 0; q < 16; q++) pk[q] = tc::pack16(0, f[2 * q], f[2 * q + 1]);
          }
          if (p.sat_counter && p.fp16) {
#pragma unroll
            for (int q = 0; q < 16; q++) nsat += ((pk[q] & 0x7FFFu) == 0x7BFFu) + (((pk[q] >> 16) & 0x7FFFu) == 0x7BFFu);
          }
          tc::tmem_st_32x16(sbase + 16 * i, pk);  // H over columns this warp has already consumed
        }
        tc::tmem_st_wait();
        tc::tc_fence_before();
        __syncwarp();
        if (lane == 0) tc::mbar_arrive(&h_full[g & 1]);
      }
      // ---- tile epilogue: Y + b2 + residual -> fp32 rows, in place
      const int yb = it % K::NY;
      tc::mbar_wait(&y_full[yb], (uint32_t)(it / K::NY) & 1u);
      tc::tc_fence_after();
      const long long row0 = (long long)tile * 128 + quad * 32, myrow = row0 + lane;
      for (int c = half; c < (C + 31) / 32; c += 2) {
        const int col0 = c * 32, ncols = C - col0;
        float4 rv[8], bv[8];
        if (myrow < p.M) {
          const float4* rp = (const float4*)(p.X + myrow * C + col0);
#pragma unroll
          for (int q = 0; q < 8; q++) rv[q] = 4 * q < ncols ? rp[q] : make_float4(0.f, 0.f, 0.f, 0.f);
        } else {
#pragma unroll
          for (int q = 0; q < 8; q++) rv[q] = make_float4(0.f, 0.f, 0.f, 0.f);
        }
#pragma unroll
        for (int q = 0; q < 8; q++) bv[q] = 4 * q < ncols ? __ldg((const float4*)(p.b2 + col0) + q) : make_float4(0.f, 0.f, 0.f, 0.f);
        uint32_t v[32];
        tc::tmem_ld_32x32(lane_addr + yb * C + col0, v);
        if (lane == 0) tc::tma_store_wait_read<0>();  // the previous store of this warp has drained the staging box
        __syncwarp();
        tc::tmem_ld_wait();
#pragma unroll
        for (int q = 0; q < 8; q++) {
          float4 o;
          o.x = __uint_as_float(v[4 * q + 0]) + bv[q].x + rv[q].x;
          o.y = __uint_as_float(v[4 * q + 1]) + bv[q].y + rv[q].y;
          o.z = __uint_as_float(v[4 * q + 2]) + bv[q].z + rv[q].z;
          o.w = __uint_as_float(v[4 * q + 3]) + bv[q].w + rv[q].w;
          *(float4*)(buf + lane * 128 + ((q ^ (lane & 7)) << 4)) = o;
        }
        tc::fence_proxy_async_smem();
        __syncwarp();
        if (lane == 0 && row0 < p.M) {
          tc::tma_store_2d(&tmap_out, buf, col0, (int)row0);
          tc::tma_store_commit();
        }
      }
      tc::tc_fence_before();
      __syncwarp();
      if (lane == 0) tc::mbar_arrive(&y_empty[yb]);
    }
    if (lane == 0) tc::tma_store_wait<0>();
    if (p.sat_counter) {
      nsat = __reduce_add_sync(0xffffffffu, nsat);
      if (lane == 0 && nsat) atomicAdd(p.sat_counter, nsat);
    }
  }
  tc::tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    tc::tc_fence_after();
    tc::tmem_dealloc<512>(tmem_base);
  }
}

template <int C_>
static int launch_mlp(const MlpFusedArgs& a, int num_sms, cudaStream_t st) {
  using K = MlpCfg<C_>;
  static std::atomic<unsigned long long> attr_set{0};
  auto kern = k_mlp_fused<C_>;
  if (cvb_once_per_device(attr_set)) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, K::SMEM);
    if (e != cudaSuccess) return cvb_fail_cuda(e, "cudaFuncSetAttribute(k_mlp_fused)");
  }
  CUtensorMap t1, t2, to;
  if (!tc_host::make_tmap_bf16(&t1, a.W1, (uint64_t)4 * C_, (uint64_t)C_, (uint64_t)C_, K::HC) ||
      !tc_host::make_tmap_bf16(&t2, a.W2, (uint64_t)C_, (uint64_t)4 * C_, (uint64_t)4 * C_, K::N2) ||
      !tc_host::make_tmap_2d(&to, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, a.X, (uint64_t)a.M, (uint64_t)C_, (uint64_t)C_ * 4, 32, 32,
                             CU_TENSOR_MAP_SWIZZLE_128B))
    return cvb_fail(CV_ERR_CUDA, "cuTensorMapEncodeTiled failed (fused MLP)");
  const int n_tiles = (a.M + 127) / 128;
  const int grid = n_tiles < num_sms ? n_tiles : num_sms;
  cvb_next_work(16.0 * (double)a.M * (double)C_ * (double)C_);  // two GEMMs of 2 * M * C * 4C flop
  if (cvb_profile_on()) {
    char nm[96];
    snprintf(nm, sizeof(nm), "mlp M%d C%d hc%d fused LN+fc1+GELU+fc2+res", a.M, C_, K::HC);
    cvb_next_name(nm);
  }
  CVB_LAUNCH(kern, dim3(grid), dim3(MLP_THREADS), K::SMEM, st, t1, t2, to, a);
  return CV_OK;
}

bool mlp_fused_supported(int C) {
  static const bool on = getenv("CVB_MLP_FUSED") ? atoi(getenv("CVB_MLP_FUSED")) != 0 : true;
  return on && (C == 96 || C == 112 || C == 144 || C == 192 || C == 224 || C == 288);
}

int mlp_fused_launch(const MlpFusedArgs& a, int num_sms, cudaStream_t st) {
  if (!a.X || !a.gamma || !a.beta || !a.W1 || !a.b1 || !a.W2 || !a.b2 || a.M <= 0)
    return cvb_fail(CV_ERR_INVALID, "mlp_fused: null argument");
  if ((((uintptr_t)a.X | (uintptr_t)a.W1 | (uintptr_t)a.W2 | (uintptr_t)a.b1 | (uintptr_t)a.b2 | (uintptr_t)a.gamma |
        (uintptr_t)a.beta) & 15) != 0)
    return cvb_fail(CV_ERR_INVALID, "mlp_fused: pointers must be 16-byte aligned");
  switch (a.C) {
    case 96: return launch_mlp<96>(a, num_sms, st);
    case 112: return launch_mlp<112>(a, num_sms, st);
    case 144: return launch_mlp<144>(a, num_sms, st);
    case 192: return launch_mlp<192>(a, num_sms, st);
    case 224: return launch_mlp<224>(a, num_sms, st);
    case 288: return launch_mlp<288>(a, num_sms, st);
  }
  return cvb_fail(CV_ERR_INVALID, "mlp_fused: unsupported width");
}

}  // namespace cvb

using namespace cvb;

// Test entry (tests/test_mlp_fused_gpu.py): the fused half-block on caller tensors.
extern "C" int cv_mlp_fused(float* X, int M, int C, const float* gamma, const float* beta, float eps, const void* W1,
                            const float* b1, const void* W2, const float* b2, int operand_fp16, void* stream) {
  cvb_reset_launches();
  if (!mlp_fused_supported(C)) return cvb_fail(CV_ERR_INVALID, "cv_mlp_fused: width without a fused instantiation");
  MlpFusedArgs a;
  a.X = X; a.M = M; a.C = C; a.gamma = gamma; a.beta = beta; a.eps = eps;
  a.W1 = (const __nv_bfloat16*)W1; a.b1 = b1; a.W2 = (const __nv_bfloat16*)W2; a.b2 = b2; a.fp16 = operand_fp16;
  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  return mlp_fused_launch(a, sms, (cudaStream_t)stream);
}
