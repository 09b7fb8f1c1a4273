// Fused LayerNorm -> fc1 -> GELU -> fc2 -> +residual for one Hiera MLP half-block (interface: mlp_fused.cuh).
//
// Persistent CTA of 28 warps per SM; a work item is a 128-row tile of the fp32 residual stream X [M, C]:
//
//   warp 3       TMA producer of the fp32 rows: slices of the tile through a small shared-memory ring (deep prefetch)
//   warps 24-27  LayerNorm producers: read the fp32 rows from that ring, normalise, write the 16-bit A operand [128, C] into
//                shared memory in the K-major 128-byte-swizzled layout tcgen05 reads (double-buffered across tiles when it fits)
//   warp 0       TMA producer: streams W1 / W2 boxes ([rows, 64] 16-bit, 128B swizzle) from L2 through a ring
//   warp 1       MMA issuer (one lane).  The hidden dimension 4C is cut into chunks of HC columns; per chunk g
//                    fc1(g):  S[g&1]  = A * W1[chunk g]^T            (SS MMA, fp32 accumulators in TMEM, 128 x HC)
//                    fc2(g):  Y      += H[g&1] * W2[:, chunk g]^T     (TS MMA: A operand = H read from TMEM)
//                issued as fc1(g), fc2(g-1), fc1(g+1), ... across tile boundaries, so the tensor pipe works on the next chunk
//                while the epilogue warps turn S into H
//   warps 4-19   chunk epilogue (four warps per TMEM lane quadrant, each owning a quarter of the chunk's columns): tcgen05.ld S -> +b1
//                -> GELU -> 16-bit -> tcgen05.st H over the head of the same TMEM columns (a warp only overwrites columns it has
//                already loaded)
//   warps 20-23  store warps, per tile: tcgen05.ld Y -> +b2 -> + fp32 residual -> swizzled staging -> TMA store (in place)
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

template <int C_>
struct MlpCfg {
  static constexpr int C = C_;
  // warp roles: 0 weight TMA, 1 MMA, 2 TMEM alloc, 3 X TMA, then EW chunk-epilogue warps (S -> GELU -> H), SW store warps
  // (Y + b2 + residual -> X, off the chunk pipeline's critical path) and 4 LayerNorm producers.  With two Y buffers the
  // store phase overlaps a whole tile, so one store warp per quadrant is enough and the epilogue gets four; with one Y buffer
  // the store phase gates the next tile's first fc2 and gets two per quadrant.
  static constexpr int EW = (C <= 192) ? 16 : 8;
  static constexpr int SW = (C <= 192) ? 4 : 8;
  static constexpr int THREADS = (4 + EW + SW + 4) * 32;
  static constexpr int HC = (C == 96) ? 128 : 64;                   // hidden columns per chunk (= N of fc1, K of fc2)
  // Y accumulator buffers: two whenever 2 C + 2 HC columns fit the 512 of TMEM — with one, the first fc2 of a tile waits for the
  // store warps to drain the previous tile (6 us at C = 192, hidden behind only two chunks of fc1)
  static constexpr int NY = (C <= 192) ? 2 : 1;
  static constexpr int NCH = 4 * C / HC;                            // chunks per tile
  static constexpr int NKB1 = (C + 63) / 64;                        // 64-wide K-blocks of fc1
  static constexpr int KS1_LAST = (C - (NKB1 - 1) * 64) / 16;       // 16-wide k-steps in the last K-block
  static constexpr int NSPLIT = C >= 192 ? 2 : 1;                   // fc2 N = C is issued as NSPLIT MMAs of N2 columns (boxes <= 18 KB)
  static constexpr int N2 = C / NSPLIT;
  static constexpr int NKB2 = HC / 64;                              // K-blocks of fc2 per chunk
  static constexpr int UNIT_ROWS = HC > N2 ? HC : N2;               // rows of the largest weight box
  static constexpr int STAGE_BYTES = UNIT_ROWS * 128;
  static constexpr int A_BYTES = NKB1 * 128 * 128;
  static constexpr int OUT_BYTES = SW * 2048;                   // one 32 x 16 fp32 staging box per store warp
  static constexpr int BIAS_BYTES = ((5 * C * 4 + 1023) / 1024) * 1024;  // b1 [4C] | b2 [C] copied to shared memory once
  static constexpr int PC = HC / (EW / 4);                      // hidden columns per epilogue warp and chunk
  // LayerNorm producers: LN_L lanes per row, LN_V float4 per lane (C = 4 * LN_L * LN_V)
  static constexpr int LN_L = (C == 96 || C == 224 || C == 288) ? 8 : (C == 192) ? 16 : (C == 384) ? 32 : 4;
  static constexpr int LN_V = C / (4 * LN_L);
  // fp32 X ring: the rows reach the LayerNorm producers through TMA (slices of XR rows), so that tens of KB are in flight
  // per SM without holding them in registers (with plain loads the four producer warps had 12 KB in flight and their
  // global-load latency, not the tensor pipe or the epilogue, set the pace: ncu, profiles/r2_mlp_fused_notes.md)
  static constexpr int RPP = 4 * (32 / LN_L);                       // rows the four producer warps cover per pass
  static constexpr int XR = (2 * RPP * C * 4 <= 16384 && 2 * RPP <= 128) ? 2 * RPP : RPP;  // rows per slice
  static constexpr int NXB = C > 256 ? 2 : 1;                       // column boxes per slice (TMA box dims are <= 256)
  static constexpr int XBC = C / NXB;
  static constexpr int X_BYTES = XR * C * 4;
  static constexpr int XSLOTS = 2;
  static constexpr int FIXED = 1024 + 512 + OUT_BYTES + BIAS_BYTES + XSLOTS * X_BYTES;
  // A operand buffers: two (LayerNorm of tile t+1 under the MMAs of tile t) when four weight stages still fit beside them
  static constexpr int NA = ((MLP_SMEM_MAX - FIXED - 2 * A_BYTES) / STAGE_BYTES >= 4) ? 2 : 1;
  static constexpr int NSTAGES_RAW = (MLP_SMEM_MAX - FIXED - NA * A_BYTES) / STAGE_BYTES;
  static constexpr int NSTAGES = NSTAGES_RAW > 8 ? 8 : NSTAGES_RAW;
  static constexpr int SMEM = FIXED + NA * A_BYTES + NSTAGES * STAGE_BYTES;
  static constexpr int SBASE = ((NY * C + 63) / 64) * 64;           // first TMEM column of the S/H double buffer
  static_assert(4 * C % HC == 0, "chunking");
  static_assert(128 % XR == 0 && X_BYTES % 1024 == 0, "X ring slices");
  static_assert(C % 16 == 0 && N2 % 16 == 0 && N2 <= 256, "MMA shapes");
  static_assert(SBASE + 2 * HC <= 512, "TMEM budget");
  static_assert(NSTAGES >= 3, "weight ring");
  static_assert(4 * LN_L * LN_V == C, "LayerNorm lane layout");
  static_assert(STAGE_BYTES % 1024 == 0 && A_BYTES % 1024 == 0, "swizzle atoms");
};

// With a single A buffer the first fc1 of a tile has to wait for the LayerNorm producers, which in turn wait for the last
// fc1 of the previous tile: the pending fc2 is issued first so that the tensor pipe is not idle behind that wait.
// Debug timeline (scripts/mlp_trace.py): one record per pipeline event of CTA 0.
__device__ __forceinline__ void mlp_trace(const MlpFusedArgs& p, unsigned ev, unsigned idx) {
  if (p.trace && blockIdx.x == 0) {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    // fixed slot per (event, index), plain store: an atomic per record costs the recording warp ~0.5 us and distorts the timeline
    if (ev < 12 && idx < 320) p.trace[1 + ev * 320 + idx] = ((unsigned long long)(ev * 4096u + idx) << 44) | (t & 0xFFFFFFFFFFFull);
  }
}

template <class K>
__device__ __forceinline__ bool fc2_first(int j) { return K::NA == 1 && j == 0; }

// ILV: issue the MMAs of a k-step alternately on two independent accumulator halves (fc1: two N = HC/2 MMAs; fc2 with
// NSPLIT = 2: the two column halves of Y) instead of running one dependent accumulation chain after the other.
template <int C_, int ILV>
__global__ void __launch_bounds__(MlpCfg<C_>::THREADS, 1)
k_mlp_fused(const __grid_constant__ CUtensorMap tmap_w1, const __grid_constant__ CUtensorMap tmap_w2,
            const __grid_constant__ CUtensorMap tmap_out, const __grid_constant__ CUtensorMap tmap_x, MlpFusedArgs p) {
  using K = MlpCfg<C_>;
  constexpr int C = K::C, HC = K::HC;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  uint8_t* sA = smem;                                  // [NA][NKB1][128 rows][128 B]
  uint8_t* sW = sA + K::NA * K::A_BYTES;               // [NSTAGES][STAGE_BYTES]
  uint8_t* sOut = sW + K::NSTAGES * K::STAGE_BYTES;    // [16 warps][2048]
  uint8_t* sX = sOut + K::OUT_BYTES;                   // [XSLOTS][NXB][XR rows][XBC] fp32
  float* sB1 = (float*)(sX + K::XSLOTS * K::X_BYTES);  // [4C]
  float* sB2 = sB1 + 4 * C;                            // [C]
  uint64_t* bars = (uint64_t*)((uint8_t*)sB1 + K::BIAS_BYTES);
  uint64_t* w_full = bars;                    // [NSTAGES]
  uint64_t* w_empty = w_full + K::NSTAGES;    // [NSTAGES]
  uint64_t* a_full = w_empty + K::NSTAGES;    // [2]
  uint64_t* a_empty = a_full + 2;             // [2]
  uint64_t* s_full = a_empty + 2;             // [2]
  uint64_t* h_full = s_full + 2;              // [2]
  uint64_t* y_full = h_full + 2;              // [2]
  uint64_t* y_empty = y_full + 2;             // [2]
  uint64_t* x_full = y_empty + 2;             // [XSLOTS]
  uint64_t* x_empty = x_full + 4;             // [XSLOTS]
  uint32_t* tmem_slot = (uint32_t*)(x_empty + 4);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int n_tiles = (p.M + 127) / 128;
  const int my_tiles = (n_tiles - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;  // tiles blockIdx.x, += gridDim.x
  const int n_chunks = my_tiles * K::NCH;

  if (warp == 0 && lane == 0) {
    tc::prefetch_tmap(&tmap_w1);
    tc::prefetch_tmap(&tmap_w2);
    tc::prefetch_tmap(&tmap_out);
    tc::prefetch_tmap(&tmap_x);
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
      tc::mbar_init(&h_full[i], K::EW);
      tc::mbar_init(&y_full[i], 1);
      tc::mbar_init(&y_empty[i], K::SW);
    }
    for (int i = 0; i < K::XSLOTS; i++) {
      tc::mbar_init(&x_full[i], 1);
      tc::mbar_init(&x_empty[i], 4);
    }
    tc::fence_barrier_init();
  }
  if (warp == 2) tc::tmem_alloc<512>(tmem_slot);
  for (int i = threadIdx.x; i < 5 * C; i += K::THREADS) sB1[i] = i < 4 * C ? p.b1[i] : p.b2[i - 4 * C];
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
  } else if (warp == 3) {
    // ===================== TMA producer of the fp32 rows (its own thread: a full weight ring must not hold it back)
    if (lane == 0) {
      constexpr int SL = 128 / K::XR;
      for (int q = 0; q < my_tiles * SL; q++) {
        const int it = q / SL, sl = q - it * SL, slot = q % K::XSLOTS;
        const int row = ((int)blockIdx.x + it * (int)gridDim.x) * 128 + sl * K::XR;
        tc::mbar_wait(&x_empty[slot], ((uint32_t)(q / K::XSLOTS) & 1u) ^ 1u);
        tc::mbar_arrive_expect_tx(&x_full[slot], K::X_BYTES);
        for (int xb = 0; xb < K::NXB; xb++)
          tc::tma_load_2d(sX + slot * K::X_BYTES + xb * (K::XR * K::XBC * 4), &tmap_x, &x_full[slot], xb * K::XBC, row);
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer.  The whole warp runs the control flow (converged: descriptors and barrier addresses
    // live in uniform registers) and one elected lane issues; inside an `if (lane == 0)` region ptxas wrapped every
    // tcgen05.mma in an ELECT / R2UR.BROADCAST / BRA.U.ANY sequence — 16 instructions and ~150 cycles per MMA, more than the
    // 48-64 cycles the N = 96 / 128 MMAs of this kernel take on the tensor pipe (timeline in profiles/r2_mlp_fused_notes.md).
    {
      const uint32_t idesc1 = tc::idesc_bf16(128, ILV ? HC / 2 : HC, false, false, p.fp16 != 0);
      const uint32_t idesc2 = tc::idesc_bf16(128, K::N2, false, false, p.fp16 != 0);
      int stage = 0;
      uint32_t phase = 0;
      auto fc1 = [&](int g) {
        const int it = g / K::NCH, j = g - it * K::NCH;
        const int ab = it % K::NA;
        if (j == 0) {
          tc::mbar_wait(&a_full[ab], (uint32_t)(it / K::NA) & 1u);
          tc::tc_fence_after();
          if (lane == 0) mlp_trace(p, 1, g);  // MMA: A operand of the tile is there
        }
        const uint32_t d = tmem_base + K::SBASE + (g & 1) * HC;
        const uint32_t a0 = tc::smem_u32(sA + ab * K::A_BYTES);
        for (int kb = 0; kb < K::NKB1; kb++) {
          tc::mbar_wait(&w_full[stage], phase);
          tc::tc_fence_after();
          const uint64_t da = tc::desc_kmajor(a0 + kb * 16384), db = tc::desc_kmajor(tc::smem_u32(sW + stage * K::STAGE_BYTES));
          const int ks = kb == K::NKB1 - 1 ? K::KS1_LAST : 4;
          if (tc::elect_one()) {
#pragma unroll
            for (int k = 0; k < 4; k++)
              if (k < ks) {
                tc::mma_f16_ss(d, da + 2 * k, db + 2 * k, idesc1, (kb > 0 || k > 0) ? 1u : 0u);  // +32 B per k-step
                if (ILV)  // second half of the chunk's columns: B rows HC/2.. (HC/2 * 128 B further), D columns + HC/2
                  tc::mma_f16_ss(d + HC / 2, da + 2 * k, db + (HC / 2 * 128 >> 4) + 2 * k, idesc1, (kb > 0 || k > 0) ? 1u : 0u);
              }
            tc::mma_commit(&w_empty[stage]);
          }
          __syncwarp();
          if (++stage == K::NSTAGES) { stage = 0; phase ^= 1; }
        }
        if (tc::elect_one()) {
          tc::mma_commit(&s_full[g & 1]);
          if (j == K::NCH - 1) tc::mma_commit(&a_empty[ab]);
        }
        __syncwarp();
        if (lane == 0) mlp_trace(p, 2, g);  // MMA: fc1(g) issued
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
        if (lane == 0) mlp_trace(p, 3, g);  // MMA: H(g) is there
        const uint32_t hbase = tmem_base + K::SBASE + (g & 1) * HC;
        if (ILV && K::NSPLIT == 2) {
          for (int kb = 0; kb < K::NKB2; kb++) {
            const int st0 = stage, st1 = (stage + 1 == K::NSTAGES) ? 0 : stage + 1;
            const uint32_t ph1 = (stage + 1 == K::NSTAGES) ? phase ^ 1 : phase;
            tc::mbar_wait(&w_full[st0], phase);
            tc::mbar_wait(&w_full[st1], ph1);
            tc::tc_fence_after();
            const uint64_t db0 = tc::desc_kmajor(tc::smem_u32(sW + st0 * K::STAGE_BYTES));
            const uint64_t db1 = tc::desc_kmajor(tc::smem_u32(sW + st1 * K::STAGE_BYTES));
            const uint32_t d = tmem_base + yb * C;
            if (tc::elect_one()) {
#pragma unroll
              for (int k = 0; k < 4; k++) {
                const int e = (kb * 4 + k) * 16;
                const int part = e / K::PC;
                const uint32_t a = hbase + part * K::PC + (e - part * K::PC) / 2;
                const uint32_t acc = (j > 0 || kb > 0 || k > 0) ? 1u : 0u;
                tc::mma_f16_ts(d, a, db0 + 2 * k, idesc2, acc);
                tc::mma_f16_ts(d + K::N2, a, db1 + 2 * k, idesc2, acc);
              }
              tc::mma_commit(&w_empty[st0]);
              tc::mma_commit(&w_empty[st1]);
            }
            __syncwarp();
            for (int z = 0; z < 2; z++)
              if (++stage == K::NSTAGES) { stage = 0; phase ^= 1; }
          }
        } else
        for (int kb = 0; kb < K::NKB2; kb++)
          for (int sp = 0; sp < K::NSPLIT; sp++) {
            tc::mbar_wait(&w_full[stage], phase);
            tc::tc_fence_after();
            const uint64_t db = tc::desc_kmajor(tc::smem_u32(sW + stage * K::STAGE_BYTES));
            const uint32_t d = tmem_base + yb * C + sp * K::N2;
            if (tc::elect_one()) {
#pragma unroll
              for (int k = 0; k < 4; k++) {
                // k-step kk covers hidden elements [16 kk, 16 kk + 16) of the chunk; each of the four epilogue warps of a
                // quadrant left its 16-bit H at the head of its own quarter of the S columns: 2 elements per 32-bit column
                const int e = (kb * 4 + k) * 16;
                const int part = e / K::PC;
                const uint32_t a = hbase + part * K::PC + (e - part * K::PC) / 2;
                tc::mma_f16_ts(d, a, db + 2 * k, idesc2, (j > 0 || kb > 0 || k > 0) ? 1u : 0u);
              }
              tc::mma_commit(&w_empty[stage]);
            }
            __syncwarp();
            if (++stage == K::NSTAGES) { stage = 0; phase ^= 1; }
          }
        if (j == K::NCH - 1) {
          if (tc::elect_one()) tc::mma_commit(&y_full[yb]);
          __syncwarp();
        }
        if (lane == 0) mlp_trace(p, 4, g);  // MMA: fc2(g) issued
      };
      for (int g = 0; g <= n_chunks; g++) {
        const bool swap = fc2_first<K>(g % K::NCH);
        if (swap && g >= 1) fc2(g - 1);
        if (g < n_chunks) fc1(g);
        if (!swap && g >= 1) fc2(g - 1);
      }
    }
  } else if (warp >= 4 + K::EW + K::SW) {
    // ===================== LayerNorm producers: fp32 rows -> normalised 16-bit A operand (swizzled K-major)
    constexpr int L = K::LN_L, V = K::LN_V, RPW = 32 / L, RPP = 4 * RPW, SL = 128 / K::XR, PPS = K::XR / RPP;
    const int pw = warp - 4 - K::EW - K::SW, sub = lane % L, rsub = lane / L;
    constexpr bool HOIST = V <= 3;  // this lane's columns are the same in every pass: keep their gamma / beta in registers
    float4 gam[HOIST ? V : 1], bet[HOIST ? V : 1];
    if (HOIST) {
#pragma unroll
      for (int k = 0; k < V; k++) {
        gam[k] = __ldg((const float4*)p.gamma + sub + k * L);
        bet[k] = __ldg((const float4*)p.beta + sub + k * L);
      }
    }
    for (int it = 0; it < my_tiles; it++) {
      const int ab = it % K::NA;
      uint8_t* A = sA + ab * K::A_BYTES;
      for (int sl = 0; sl < SL; sl++) {
        const int q = it * SL + sl, slot = q % K::XSLOTS;
        tc::mbar_wait(&x_full[slot], (uint32_t)(q / K::XSLOTS) & 1u);
        if (sl == 0) {
          if (pw == 0 && lane == 0) mlp_trace(p, 9, it);  // LN: first slice of the tile is there
          tc::mbar_wait(&a_empty[ab], ((uint32_t)(it / K::NA) & 1u) ^ 1u);
          if (pw == 0 && lane == 0) mlp_trace(p, 10, it);  // LN: A buffer is free
        }
        const uint8_t* xs = sX + slot * K::X_BYTES;
#pragma unroll
        for (int ps = 0; ps < PPS; ps++) {
          const int rl = ps * RPP + pw * RPW + rsub;  // row within the slice
          const int r = sl * K::XR + rl;              // row within the tile (rows beyond M arrive as zeros: TMA fill)
          float4 v[V];
#pragma unroll
          for (int k = 0; k < V; k++) {
            const int col = (sub + k * L) * 4, xb = col / K::XBC;
            v[k] = *(const float4*)(xs + xb * (K::XR * K::XBC * 4) + rl * (K::XBC * 4) + (col - xb * K::XBC) * 4);
          }
          float s = 0.f;
#pragma unroll
          for (int k = 0; k < V; k++) s += v[k].x + v[k].y + v[k].z + v[k].w;
#pragma unroll
          for (int o = L / 2; o; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
          const float mean = s * (1.0f / C);
          float qq = 0.f;
#pragma unroll
          for (int k = 0; k < V; k++) {
            const float a = v[k].x - mean, b = v[k].y - mean, c = v[k].z - mean, d = v[k].w - mean;
            qq += a * a + b * b + c * c + d * d;
          }
#pragma unroll
          for (int o = L / 2; o; o >>= 1) qq += __shfl_xor_sync(0xffffffffu, qq, o);
          const float rstd = rsqrtf(qq * (1.0f / C) + p.eps);
#pragma unroll
          for (int k = 0; k < V; k++) {
            const int i4 = sub + k * L;  // float4 index within the row; columns 4 i4 .. 4 i4 + 3
            const float4 g = HOIST ? gam[HOIST ? k : 0] : __ldg((const float4*)p.gamma + i4);
            const float4 bb = HOIST ? bet[HOIST ? k : 0] : __ldg((const float4*)p.beta + i4);
            const float o0 = (v[k].x - mean) * rstd * g.x + bb.x, o1 = (v[k].y - mean) * rstd * g.y + bb.y;
            const float o2 = (v[k].z - mean) * rstd * g.z + bb.z, o3 = (v[k].w - mean) * rstd * g.w + bb.w;
            const int col = i4 * 4, kb = col >> 6, ch = (col & 63) >> 3, hf = (col & 7) >> 2;
            *(uint2*)(A + kb * 16384 + r * 128 + (((ch ^ (r & 7)) << 4) | (hf << 3))) =
                make_uint2(tc::pack16(p.fp16, o0, o1), tc::pack16(p.fp16, o2, o3));
          }
        }
        __syncwarp();
        if (lane == 0) tc::mbar_arrive(&x_empty[slot]);  // the slice has been read: the TMA thread may refill the slot
      }
      tc::fence_proxy_async_smem();  // generic-proxy writes -> visible to the tensor core's async-proxy reads
      __syncwarp();
      if (lane == 0) tc::mbar_arrive(&a_full[ab]);
      if (pw == 0 && lane == 0) mlp_trace(p, 11, it);  // LN: A operand written
    }
  } else if (warp >= 4 + K::EW) {
    // ===================== store warps: quadrant = warp % 4 (TMEM lanes), unit parity = (warp - 12) / 4.
    // Y + b2 + residual -> fp32 rows, in place, in units of 16 columns.  They only ever wait for y_full, so the chunk
    // epilogue warps run straight on into the next tile (with the same warps doing both, the 4-5 us of this epilogue
    // stalled the chunk pipeline at every tile boundary: timeline in profiles/r2_mlp_fused_notes.md).
    const int quad = warp & 3, uh = (warp - 4 - K::EW) >> 2;
    constexpr int USTEP = K::SW / 4;  // store warps per quadrant
    const uint32_t lane_addr = tmem_base + ((uint32_t)(quad * 32) << 16);
    uint8_t* buf = sOut + (warp - 4 - K::EW) * 2048;
    for (int it = 0; it < my_tiles; it++) {
      const int tile = (int)blockIdx.x + it * (int)gridDim.x;
      const int yb = it % K::NY;
      const long long row0 = (long long)tile * 128 + quad * 32, myrow = row0 + lane;
      const bool live = myrow < p.M;
      float4 rv[4];
      auto load_res = [&](int u) {  // the residual of the first unit is requested before the wait on the accumulator
        const float4* rp = (const float4*)(p.X + myrow * C + u * 16);
#pragma unroll
        for (int q = 0; q < 4; q++) rv[q] = live ? rp[q] : make_float4(0.f, 0.f, 0.f, 0.f);
      };
      if (uh < C / 16) load_res(uh);
      tc::mbar_wait(&y_full[yb], (uint32_t)(it / K::NY) & 1u);
      tc::tc_fence_after();
      if (warp == 4 + K::EW && lane == 0) mlp_trace(p, 7, it);  // store: Y of the tile is there
      for (int u = uh; u < C / 16; u += USTEP) {
        const int col0 = u * 16;
        uint32_t v[16];
        tc::tmem_ld_32x16(lane_addr + yb * C + col0, v);
        if (lane == 0) tc::tma_store_wait_read<0>();  // the previous store of this warp has drained the staging box
        __syncwarp();
        tc::tmem_ld_wait();
        const float4* bp = (const float4*)(sB2 + col0);
#pragma unroll
        for (int q = 0; q < 4; q++) {
          const float4 bv = bp[q];
          float4 o;
          o.x = __uint_as_float(v[4 * q + 0]) + bv.x + rv[q].x;
          o.y = __uint_as_float(v[4 * q + 1]) + bv.y + rv[q].y;
          o.z = __uint_as_float(v[4 * q + 2]) + bv.z + rv[q].z;
          o.w = __uint_as_float(v[4 * q + 3]) + bv.w + rv[q].w;
          // 32 x 16 fp32 box, 64-byte rows, 64B swizzle: 16-byte chunk q of row r lives at chunk q ^ ((r >> 1) & 3)
          *(float4*)(buf + lane * 64 + ((q ^ ((lane >> 1) & 3)) << 4)) = o;
        }
        if (u + USTEP < C / 16) load_res(u + USTEP);
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
      if (warp == 4 + K::EW && lane == 0) mlp_trace(p, 8, it);  // store: tile stored
    }
    if (lane == 0) tc::tma_store_wait<0>();
  } else if (warp >= 4) {
    // ===================== chunk-epilogue warps: quadrant = warp % 4 (TMEM lanes), column half = (warp - 4) / 4
    constexpr int PC = K::PC;
    const int quad = warp & 3, part = (warp - 4) >> 2;
    const uint32_t lane_addr = tmem_base + ((uint32_t)(quad * 32) << 16);
    unsigned int nsat = 0;
    for (int g = 0; g < n_chunks; g++) {
      const int j = g % K::NCH;
      tc::mbar_wait(&s_full[g & 1], (uint32_t)(g >> 1) & 1u);
      tc::tc_fence_after();
      if (warp == 4 && lane == 0) mlp_trace(p, 5, g);  // epilogue: S(g) is there
      const uint32_t sbase = lane_addr + K::SBASE + (g & 1) * HC + part * PC;
      constexpr int SUBW = PC >= 32 ? 32 : 16;  // columns per TMEM load
#pragma unroll
      for (int i = 0; i < PC / SUBW; i++) {
        uint32_t v[SUBW];
        if (SUBW == 32) tc::tmem_ld_32x32(sbase + 32 * i, *(uint32_t(*)[32])v);
        else tc::tmem_ld_32x16(sbase + 16 * i, *(uint32_t(*)[16])v);
        const float4* bp = (const float4*)(sB1 + j * HC + part * PC + SUBW * i);  // same address for every lane: broadcast reads
        tc::tmem_ld_wait();
        uint32_t pk[SUBW / 2];
#pragma unroll
        for (int q = 0; q < SUBW / 4; q++) {
          const float4 bv = bp[q];
          float f0, f1, f2, f3;
          up2(add2(pk2(__uint_as_float(v[4 * q + 0]), __uint_as_float(v[4 * q + 1])), pk2(bv.x, bv.y)), f0, f1);
          up2(add2(pk2(__uint_as_float(v[4 * q + 2]), __uint_as_float(v[4 * q + 3])), pk2(bv.z, bv.w)), f2, f3);
          gelu_tanh2(f0, f1);
          gelu_tanh2(f2, f3);
          if (p.fp16) { pk[2 * q] = tc::pack16(1, f0, f1); pk[2 * q + 1] = tc::pack16(1, f2, f3); }
          else { pk[2 * q] = tc::pack16(0, f0, f1); pk[2 * q + 1] = tc::pack16(0, f2, f3); }
        }
        if (p.sat_counter && p.fp16) {
#pragma unroll
          for (int q = 0; q < SUBW / 2; q++) nsat += ((pk[q] & 0x7FFFu) == 0x7BFFu) + (((pk[q] >> 16) & 0x7FFFu) == 0x7BFFu);
        }
        // H over columns this warp has already consumed
        if (SUBW == 32) tc::tmem_st_32x16(sbase + 16 * i, *(uint32_t(*)[16])pk);
        else tc::tmem_st_32x8(sbase + 8 * i, *(uint32_t(*)[8])pk);
      }
      tc::tmem_st_wait();
      tc::tc_fence_before();
      __syncwarp();
      if (lane == 0) tc::mbar_arrive(&h_full[g & 1]);
      if (warp == 4 && lane == 0) mlp_trace(p, 6, g);  // epilogue: H(g) written
    }
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

template <int C_, int ILV>
static int launch_mlp_v(const MlpFusedArgs& a, int num_sms, cudaStream_t st) {
  using K = MlpCfg<C_>;
  static std::atomic<unsigned long long> attr_set{0};
  auto kern = k_mlp_fused<C_, ILV>;
  if (cvb_once_per_device(attr_set)) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, K::SMEM);
    if (e != cudaSuccess) return cvb_fail_cuda(e, "cudaFuncSetAttribute(k_mlp_fused)");
  }
  CUtensorMap t1, t2, to;
  if (!tc_host::make_tmap_bf16(&t1, a.W1, (uint64_t)4 * C_, (uint64_t)C_, (uint64_t)C_, K::HC) ||
      !tc_host::make_tmap_bf16(&t2, a.W2, (uint64_t)C_, (uint64_t)4 * C_, (uint64_t)4 * C_, K::N2) ||
      !tc_host::make_tmap_2d(&to, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, a.X, (uint64_t)a.M, (uint64_t)C_, (uint64_t)C_ * 4, 16, 32,
                             CU_TENSOR_MAP_SWIZZLE_64B))
    return cvb_fail(CV_ERR_CUDA, "cuTensorMapEncodeTiled failed (fused MLP)");
  CUtensorMap tx;
  if (!tc_host::make_tmap_2d(&tx, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, a.X, (uint64_t)a.M, (uint64_t)C_, (uint64_t)C_ * 4, K::XBC, K::XR,
                             CU_TENSOR_MAP_SWIZZLE_NONE))
    return cvb_fail(CV_ERR_CUDA, "cuTensorMapEncodeTiled failed (fused MLP, X rows)");
  const int n_tiles = (a.M + 127) / 128;
  const int grid = n_tiles < num_sms ? n_tiles : num_sms;
  cvb_next_work(16.0 * (double)a.M * (double)C_ * (double)C_);  // two GEMMs of 2 * M * C * 4C flop
  if (cvb_profile_on()) {
    char nm[96];
    snprintf(nm, sizeof(nm), "mlp M%d C%d hc%d fused LN+fc1+GELU+fc2+res", a.M, C_, K::HC);
    cvb_next_name(nm);
  }
  CVB_LAUNCH(kern, dim3(grid), dim3(K::THREADS), K::SMEM, st, t1, t2, to, tx, a);
  return CV_OK;
}

template <int C_>
static int launch_mlp(const MlpFusedArgs& a, int num_sms, cudaStream_t st) {
  static const int ilv = getenv("CVB_MLP_ILV") ? atoi(getenv("CVB_MLP_ILV")) : 0;
  return ilv ? launch_mlp_v<C_, 1>(a, num_sms, st) : launch_mlp_v<C_, 0>(a, num_sms, st);
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
static unsigned long long* g_mlp_trace = nullptr;
// Debug: device buffer of >= 4001 uint64 (zeroed) that the next cv_mlp_fused calls fill with CTA 0's pipeline timeline.
extern "C" int cv_mlp_fused_set_trace(void* device_buffer) {
  g_mlp_trace = (unsigned long long*)device_buffer;
  return CV_OK;
}

extern "C" int cv_mlp_fused(float* X, int M, int C, const float* gamma, const float* beta, float eps, const void* W1,
                            const float* b1, const void* W2, const float* b2, int operand_fp16, void* stream) {
  cvb_reset_launches();
  if (!mlp_fused_supported(C)) return cvb_fail(CV_ERR_INVALID, "cv_mlp_fused: width without a fused instantiation");
  MlpFusedArgs a;
  a.X = X; a.M = M; a.C = C; a.gamma = gamma; a.beta = beta; a.eps = eps;
  a.W1 = (const __nv_bfloat16*)W1; a.b1 = b1; a.W2 = (const __nv_bfloat16*)W2; a.b2 = b2; a.fp16 = operand_fp16;
  a.trace = g_mlp_trace;
  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  return mlp_fused_launch(a, sms, (cudaStream_t)stream);
}
