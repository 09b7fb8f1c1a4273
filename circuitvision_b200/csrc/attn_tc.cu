// Block-diagonal ("windowed") flash attention on tcgen05 + TMA, head_dim 96, bf16 operands, fp32 softmax.
// One kernel serves every attention of the Hiera trunk: tokens are stored window-major, a query
// row of window w attends exactly the keys of window w, with Wq query tokens and Wkv key tokens per window
// (Wq == Wkv for ordinary blocks, Wq = Wkv/4 for Q-pooled blocks, Wq = Wkv = tokens-per-image for global blocks).
// Zero-pad tokens inside a window are ordinary keys (the reference does not mask them).
//
// CTA = one 128-row query tile of one head.  Warp roles: warp0 TMA producer, warp1 MMA issuer (one thread),
// warp2 TMEM allocator, warps 4-7 softmax (thread = query row).  Per 128-key tile:
//   S = Q K^T  (tcgen05.mma 128x128x96, S in TMEM)  ->  softmax warps read S, write P (bf16) into shared memory
//   in the K-major 128B-swizzled layout  ->  O_tile = P V (tcgen05.mma 128x96x128, V consumed MN-major straight
//   from its row-major tile)  ->  softmax warps fold O_tile into fp32 registers with the online-softmax rescale.
#include <stdlib.h>

#include "common.cuh"
#include "tc05.cuh"
#include "act.cuh"

namespace cvb {

constexpr int ATT_BM = 128;
constexpr int ATT_BN = 128;
// head dim D = 64 or 96 (Hiera's 56 and 72 are zero-padded to these by the weight folding); a 128-row operand tile is
// ceil(D / 64) swizzle atoms of 128 rows x 128 bytes
template <int D>
constexpr int att_tile_bytes() { return ((D + 63) / 64) * ATT_BM * 128; }

struct AttnParams {
  int Mq, Mkv, Wq, Wkv, heads;
  int qcol0, kcol0, vcol0;  // column of head 0 inside the Q / K / V tensor maps
  float scale_log2;         // softmax scale * log2(e)
  int fp16;                 // q/k/v/P/out are IEEE half instead of bf16
  int skip_chunks;          // windowed kernel: skip score chunks no row of the warp can see
  int q_rows;               // windowed kernel: query rows per work item (<= 128), a whole number of windows or an even
                            // split of one window, so that an item never straddles windows it does not need
  __nv_bfloat16* out;       // [Mq, heads*96]
  long long ld_out;
  unsigned long long* trace;  // debug timeline of CTA (0, 0) of the global kernel (cv_attn_set_trace), else nullptr
};

// Debug timeline (scripts/attn_trace.py): one record (event, index, %globaltimer) per pipeline event of CTA (0, 0).
// Records go to a shared-memory array (a global atomic per event costs the recording warp ~0.5 us and distorts the timeline);
// slot = ev * 64 + idx (idx < 64), dumped by attn_trace_dump at the end of the kernel.
constexpr int ATT_TRACE_SLOTS = 8 * 64;
__device__ __forceinline__ void attn_trace(const AttnParams& p, unsigned long long* tbuf, unsigned ev, unsigned idx) {
  if (p.trace && blockIdx.x == 0 && blockIdx.y == 0 && idx < 64) {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    tbuf[ev * 64 + idx] = ((unsigned long long)(ev * 4096u + idx) << 44) | (t & 0xFFFFFFFFFFFull);
  }
}
__device__ __forceinline__ void attn_trace_dump(const AttnParams& p, const unsigned long long* tbuf) {
  if (p.trace && blockIdx.x == 0 && blockIdx.y == 0) {
    for (int i = threadIdx.x; i < ATT_TRACE_SLOTS; i += blockDim.x) p.trace[1 + i] = tbuf[i];
    if (threadIdx.x == 0) p.trace[0] = ATT_TRACE_SLOTS;
  }
}

__device__ __forceinline__ float ex2(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

// ------------------------------------------------------------------------------------------------ global attention
// Wq == Wkv (one window = all tokens of an image), Wkv % 128 == 0, Wq % 256 == 0: the three global blocks of stage 3
// (4096 x 4096 per head) — 85 of the trunk's 86 attention GFLOP per image.
// CTA = TWO 128-row query tiles of one head sharing every K/V tile; 12 warps:
//   warp0 TMA producer, warp1 MMA issuer, warp2 TMEM allocator, warps 4-7 softmax of tile 0, warps 8-11 softmax of tile 1.
// TMEM: S0 | S1 (128 fp32 columns each) and O0 | O1 (96 columns each).  Per key tile and query tile:
//   S = Q K^T (SS MMA)  ->  softmax warps read S ONCE, write P = 2^(s*scale - m_ref) as packed bf16 over the first 64
//   columns of S (tcgen05.st)  ->  O += P V with P as the TMEM A operand (TS MMA, V consumed MN-major from smem).
// The tensor pipe ping-pongs between the two query tiles, so softmax of one overlaps the MMAs of the other.
// m_ref is a lazily updated reference maximum: P uses the reference carried in from the previous tiles and O / l are
// rescaled (tcgen05.ld -> mul -> tcgen05.st) only when a tile's maximum exceeds it by more than 2^8 — rare after the
// first tile, and exact either way because the same reference scales numerator and denominator.
constexpr int AG_THREADS = 384;
template <int D>
constexpr int ag_smem() { return 1024 + att_tile_bytes<D>() * (2 /*Q0,Q1*/ + 2 /*K*/ + 2 /*V*/) + 256; }

// 2^x for a PAIR of scores on the FMA pipe (Cody-Waite split + degree-3 minimax polynomial, relative error 7.7e-5 — a sixth of
// an fp16 ulp of P): x = n + f with n = round(x) taken from the low mantissa bits of x + 1.5 * 2^23, 2^f on [-0.5, 0.5] by
// Horner, 2^n by an integer add into the exponent field.  The softmax warps of the global kernel are bound by the XU pipe
// (one MUFU.EX2 per score: 56 % busy in ncu while the tensor pipe sat at 40 %), so a third of the scores go through here.
__device__ __forceinline__ void ex2_poly2(float x0, float x1, float& e0, float& e1) {
  x0 = fmaxf(x0, -125.f);
  x1 = fmaxf(x1, -125.f);
  const uint64_t x = pk2(x0, x1);
  const uint64_t xf = add2(x, pk2(12582912.f, 12582912.f));
  const uint64_t r = add2(xf, pk2(-12582912.f, -12582912.f));
  const uint64_t f = fma2(r, pk2(-1.f, -1.f), x);
  uint64_t q = fma2(f, pk2(0.05508868f, 0.05508868f), pk2(0.24260405f, 0.24260405f));
  q = fma2(q, f, pk2(0.69327624f, 0.69327624f));
  q = fma2(q, f, pk2(0.99992894f, 0.99992894f));
  float q0, q1, n0, n1;
  up2(q, q0, q1);
  up2(xf, n0, n1);
  e0 = __int_as_float(__float_as_int(q0) + (__float_as_int(n0) << 23));
  e1 = __int_as_float(__float_as_int(q1) + (__float_as_int(n1) << 23));
}

template <int D, bool EXP_FMA>
__global__ void __launch_bounds__(AG_THREADS, 1)
k_attn_global(const __grid_constant__ CUtensorMap tq, const __grid_constant__ CUtensorMap tk,
              const __grid_constant__ CUtensorMap tv, AttnParams p) {
  constexpr int ATT_D = D, ATT_TILE_BYTES = att_tile_bytes<D>();
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  uint8_t* sQ = smem;                      // [2]
  uint8_t* sK = sQ + 2 * ATT_TILE_BYTES;   // [2]
  uint8_t* sV = sK + 2 * ATT_TILE_BYTES;   // [2]
  uint64_t* bars = (uint64_t*)(sV + 2 * ATT_TILE_BYTES);
  uint64_t* q_full = bars;         // 1
  uint64_t* k_full = bars + 1;     // [2]
  uint64_t* k_empty = bars + 3;    // [2]
  uint64_t* v_full = bars + 5;     // [2]
  uint64_t* v_empty = bars + 7;    // [2]
  uint64_t* s_full = bars + 9;     // [2] per query tile
  uint64_t* p_full = bars + 11;    // [2]
  uint64_t* o_full = bars + 13;    // [2]
  uint32_t* tmem_slot = (uint32_t*)(bars + 15);

  const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0), lane = threadIdx.x & 31;  // shuffle: warp-uniform for the compiler
  const int q0 = blockIdx.x * 2 * ATT_BM;
  const int head = blockIdx.y;
  const int kv_start = (q0 / p.Wq) * p.Wkv;
  const int n_tiles = p.Wkv / ATT_BN;

  if (warp == 0 && lane == 0) {
    tc::prefetch_tmap(&tq);
    tc::prefetch_tmap(&tk);
    tc::prefetch_tmap(&tv);
  }
  if (warp == 1 && lane == 0) {
    tc::mbar_init(q_full, 1);
    for (int i = 0; i < 2; i++) {
      tc::mbar_init(&k_full[i], 1);
      tc::mbar_init(&k_empty[i], 1);
      tc::mbar_init(&v_full[i], 1);
      tc::mbar_init(&v_empty[i], 1);
      tc::mbar_init(&s_full[i], 1);
      tc::mbar_init(&p_full[i], 4);  // one arrival per softmax warp of the tile
      tc::mbar_init(&o_full[i], 1);
    }
    tc::fence_barrier_init();
  }
  if (warp == 2) tc::tmem_alloc<512>(tmem_slot);
  tc::tc_fence_before();
  __syncthreads();
  tc::tc_fence_after();
  const uint32_t tmem_base = __shfl_sync(0xffffffffu, *tmem_slot, 0);

  if (warp == 0) {
    if (lane == 0) {
      tc::mbar_arrive_expect_tx(q_full, 2 * ATT_TILE_BYTES);
      for (int g = 0; g < 2; g++) {
        tc::tma_load_2d(sQ + g * ATT_TILE_BYTES, &tq, q_full, p.qcol0 + head * ATT_D, q0 + g * ATT_BM);
        if (D > 64) tc::tma_load_2d(sQ + g * ATT_TILE_BYTES + ATT_BM * 128, &tq, q_full, p.qcol0 + head * ATT_D + 64, q0 + g * ATT_BM);
      }
      for (int j = 0; j < n_tiles; j++) {
        const int s = j & 1;
        const uint32_t par = ((j >> 1) & 1) ^ 1;
        const int row = kv_start + j * ATT_BN;
        tc::mbar_wait(&k_empty[s], par);
        tc::mbar_arrive_expect_tx(&k_full[s], ATT_TILE_BYTES);
        uint8_t* k = sK + s * ATT_TILE_BYTES;
        tc::tma_load_2d(k, &tk, &k_full[s], p.kcol0 + head * ATT_D, row);
        if (D > 64) tc::tma_load_2d(k + ATT_BN * 128, &tk, &k_full[s], p.kcol0 + head * ATT_D + 64, row);
        tc::mbar_wait(&v_empty[s], par);
        tc::mbar_arrive_expect_tx(&v_full[s], ATT_TILE_BYTES);
        uint8_t* v = sV + s * ATT_TILE_BYTES;
        tc::tma_load_2d(v, &tv, &v_full[s], p.vcol0 + head * ATT_D, row);
        if (D > 64) tc::tma_load_2d(v + ATT_BN * 128, &tv, &v_full[s], p.vcol0 + head * ATT_D + 64, row);
      }
    }
  } else if (warp == 1) {
    // MMA issuer: the whole warp runs the control flow (converged), one elected lane issues — inside an `if (lane == 0)`
    // region every tcgen05.mma was wrapped in an ELECT / R2UR.BROADCAST / BRA.U.ANY loop (~25 dependent instructions)
    {
      const uint32_t idesc_qk = tc::idesc_bf16(ATT_BM, ATT_BN, false, false, p.fp16 != 0);
      const uint32_t idesc_pv = tc::idesc_bf16(ATT_BM, ATT_D, false, true, p.fp16 != 0);  // A = P (TMEM, K-major), B = V MN-major
      auto issue_qk = [&](int g, int j, uint64_t* also) {
        const uint64_t dq = tc::desc_kmajor(tc::smem_u32(sQ + g * ATT_TILE_BYTES));
        const uint64_t dk = tc::desc_kmajor(tc::smem_u32(sK + (j & 1) * ATT_TILE_BYTES));
        if (tc::elect_one()) {
#pragma unroll
          for (int k = 0; k < ATT_D / 16; k++) {
            const uint32_t off = ((k >> 2) * (ATT_BM * 128) + (k & 3) * 32) >> 4;  // descriptor address units of 16 bytes
            tc::mma_f16_ss(tmem_base + g * 128, dq + off, dk + off, idesc_qk, k > 0 ? 1u : 0u);
          }
          tc::mma_commit(&s_full[g]);
          if (also) tc::mma_commit(also);
        }
        __syncwarp();
      };
      auto issue_pv = [&](int g, int j, uint64_t* also) {
        const uint64_t dv = tc::smem_desc_sw128(tc::smem_u32(sV + (j & 1) * ATT_TILE_BYTES), ATT_BN * 128, 1024);
        if (tc::elect_one()) {
#pragma unroll
          for (int k = 0; k < ATT_BN / 16; k++)
            // A: 16 keys = 8 packed columns of the P region; B: 16 key rows (2048 B) per k-step of the MN-major V tile
            tc::mma_f16_ts(tmem_base + 256 + g * ATT_D, tmem_base + g * 128 + k * 8, dv + (k * 2048 >> 4), idesc_pv,
                           (j > 0 || k > 0) ? 1u : 0u);
          tc::mma_commit(&o_full[g]);
          if (also) tc::mma_commit(also);
        }
        __syncwarp();
      };
      tc::mbar_wait(q_full, 0);
      tc::mbar_wait(&k_full[0], 0);
      tc::tc_fence_after();
      issue_qk(0, 0, nullptr);
      issue_qk(1, 0, &k_empty[0]);
      for (int j = 0; j < n_tiles; j++) {
        const bool more = j + 1 < n_tiles;
        tc::mbar_wait(&v_full[j & 1], (j >> 1) & 1);
        tc::mbar_wait(&p_full[0], j & 1);
        tc::tc_fence_after();
        issue_pv(0, j, nullptr);
        if (more) {
          tc::mbar_wait(&k_full[(j + 1) & 1], ((j + 1) >> 1) & 1);
          tc::tc_fence_after();
          issue_qk(0, j + 1, nullptr);
        }
        tc::mbar_wait(&p_full[1], j & 1);
        tc::tc_fence_after();
        issue_pv(1, j, &v_empty[j & 1]);
        if (more) issue_qk(1, j + 1, &k_empty[(j + 1) & 1]);
      }
    }
  } else if (warp >= 4) {
    const int g = (warp - 4) >> 2;  // query tile of this warp
    const int quad = warp & 3;
    const int r = quad * 32 + lane;
    const int grow = q0 + g * ATT_BM + r;
    const uint32_t lane_sel = (uint32_t)(quad * 32) << 16;
    const uint32_t tS = tmem_base + lane_sel + g * 128;
    const uint32_t tO = tmem_base + lane_sel + 256 + g * ATT_D;
    float m_ref = 0.f, l = 0.f;
    for (int j = 0; j < n_tiles; j++) {
      tc::mbar_wait(&s_full[g], j & 1);
      tc::tc_fence_after();
      if (j == 0) {
        float mx = -INFINITY;
#pragma unroll
        for (int c = 0; c < ATT_BN / 32; c++) {
          uint32_t v[32];
          tc::tmem_ld_32x32(tS + c * 32, v);
          tc::tmem_ld_wait();
#pragma unroll
          for (int i = 0; i < 32; i++) mx = fmaxf(mx, __uint_as_float(v[i]));
        }
        m_ref = mx * p.scale_log2;
      }
      // per PAIR of scores: one FMNMX3 (running max), one packed FFMA2 (scale and shift), two ex2, one packed FADD2
      // (row sum) and one pack — the softmax warps are issue-bound, so every instruction per element counts
      float tmax = -INFINITY;
      uint64_t rs2 = pk2(0.f, 0.f);
      const uint64_t sc2 = pk2(p.scale_log2, p.scale_log2), mr2 = pk2(-m_ref, -m_ref);
      // the TMEM load of chunk c + 1 is in flight while chunk c is exponentiated
      uint32_t v[2][32];
      tc::tmem_ld_32x32(tS, v[0]);
#pragma unroll
      for (int c = 0; c < ATT_BN / 32; c++) {
        tc::tmem_ld_wait();
        if (c + 1 < ATT_BN / 32) tc::tmem_ld_32x32(tS + (c + 1) * 32, v[(c + 1) & 1]);
        uint32_t pk[16];
        float pe[32];
#pragma unroll
        for (int i = 0; i < 32; i += 2) {
          const float s0 = __uint_as_float(v[c & 1][i]), s1 = __uint_as_float(v[c & 1][i + 1]);
          asm("max.f32 %0, %0, %1, %2;" : "+f"(tmax) : "f"(s0), "f"(s1));
          float x0, x1;
          up2(fma2(pk2(s0, s1), sc2, mr2), x0, x1);
          if (EXP_FMA && ((i >> 1) % 3) == 2) {  // every third pair: exponentials on the FMA pipe
            ex2_poly2(x0, x1, pe[i], pe[i + 1]);
          } else {
            pe[i] = ex2(x0);
            pe[i + 1] = ex2(x1);
          }
          const uint64_t pp = pk2(pe[i], pe[i + 1]);
          asm("add.rn.f32x2 %0, %0, %1;" : "+l"(rs2) : "l"(pp));
        }
        // one uniform branch per chunk on the operand format instead of a predicated pair of conversions per score pair
        if (p.fp16) {
#pragma unroll
          for (int i = 0; i < 16; i++) pk[i] = tc::pack16(1, pe[2 * i], pe[2 * i + 1]);
        } else {
#pragma unroll
          for (int i = 0; i < 16; i++) pk[i] = tc::pack16(0, pe[2 * i], pe[2 * i + 1]);
        }
        tc::tmem_st_32x16(tS + c * 16, pk);  // P over the already-consumed head of S
      }
      float rs_lo, rs_hi;
      up2(rs2, rs_lo, rs_hi);
      const float rowsum = rs_lo + rs_hi;
      tc::tmem_st_wait();
      tc::tc_fence_before();
      __syncwarp();
      if (lane == 0) tc::mbar_arrive(&p_full[g]);
      l += rowsum;
      const float tm = tmax * p.scale_log2;
      const bool need = tm > m_ref + 8.0f;
      if (__any_sync(0xffffffffu, need)) {
        // rescale O and l to the new reference once P V of this tile has been accumulated
        tc::mbar_wait(&o_full[g], j & 1);
        tc::tc_fence_after();
        const float alpha = need ? ex2(m_ref - tm) : 1.0f;
#pragma unroll
        for (int c = 0; c < ATT_D / 32; c++) {
          uint32_t v[32];
          tc::tmem_ld_32x32(tO + c * 32, v);
          tc::tmem_ld_wait();
#pragma unroll
          for (int i = 0; i < 32; i++) v[i] = __float_as_uint(__uint_as_float(v[i]) * alpha);
          tc::tmem_st_32x32(tO + c * 32, v);
        }
        tc::tmem_st_wait();
        tc::tc_fence_before();
        l *= alpha;
        if (need) m_ref = tm;
      }
    }
    tc::mbar_wait(&o_full[g], (n_tiles - 1) & 1);
    tc::tc_fence_after();
    if (grow < p.Mq) {
      const float inv = 1.f / l;
      __nv_bfloat16* o = p.out + (long long)grow * p.ld_out + head * ATT_D;
#pragma unroll
      for (int c = 0; c < ATT_D / 32; c++) {
        uint32_t v[32];
        tc::tmem_ld_32x32(tO + c * 32, v);
        tc::tmem_ld_wait();
#pragma unroll
        for (int i = 0; i < 32; i += 8) {
          uint32_t w[4];
#pragma unroll
          for (int k = 0; k < 4; k++) {
            w[k] = tc::pack16(p.fp16, __uint_as_float(v[i + 2 * k]) * inv, __uint_as_float(v[i + 2 * k + 1]) * inv);
          }
          *(uint4*)(o + c * 32 + i) = make_uint4(w[0], w[1], w[2], w[3]);
        }
      }
    } else {
      // keep the warp-collective TMEM loads aligned for rows past Mq
#pragma unroll
      for (int c = 0; c < ATT_D / 32; c++) {
        uint32_t v[32];
        tc::tmem_ld_32x32(tO + c * 32, v);
        tc::tmem_ld_wait();
      }
    }
  }
  tc::tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    tc::tc_fence_after();
    tc::tmem_dealloc<512>(tmem_base);
  }
}

// Two softmax threads per query row (k_attn_global2): ncu of k_attn_global showed its softmax warps neither XU- nor issue-bound
// (two warps per scheduler at IPC 0.33: every warp is one long dependent stream of TMEM load -> 32 exponentials -> TMEM store) and
// waiting for S 36 % of the time because softmax -> P V -> Q K^T of ONE query tile is a serial chain (P overwrites S).  Here a
// row's 128 scores are split between two warps of the same TMEM lane quadrant (columns 0-63 / 64-127): twice the warps per
// scheduler hide the latencies, and the chain's softmax link is half as long.  The two threads of a row agree on the reference
// maximum through a small shared-memory exchange and a 64-thread named barrier per key tile; the P of the upper half lives in
// the upper half's own S columns (64..95), so no thread overwrites scores its partner has not read yet.
constexpr int AG2_THREADS = 640;
__device__ __forceinline__ void pair_sync(int id) { asm volatile("bar.sync %0, 64;" ::"r"(id) : "memory"); }

template <int D, bool EXP_FMA>
__global__ void __launch_bounds__(AG2_THREADS, 1)
k_attn_global2(const __grid_constant__ CUtensorMap tq, const __grid_constant__ CUtensorMap tk,
               const __grid_constant__ CUtensorMap tv, AttnParams p) {
  constexpr int ATT_D = D, ATT_TILE_BYTES = att_tile_bytes<D>(), DH = D / 2;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  uint8_t* sQ = smem;                      // [2]
  uint8_t* sK = sQ + 2 * ATT_TILE_BYTES;   // [2]
  uint8_t* sV = sK + 2 * ATT_TILE_BYTES;   // [2]
  float* xch = (float*)(sV + 2 * ATT_TILE_BYTES);  // [3 slots][2 tiles][2 halves][128 rows]: tile maxima (two slots) and row sums
  unsigned long long* tbuf = (unsigned long long*)(xch + 3 * 2 * 2 * 128);  // debug timeline (ATT_TRACE_SLOTS records)
  uint64_t* bars = (uint64_t*)(tbuf + ATT_TRACE_SLOTS);
  uint64_t* q_full = bars;         // 1
  uint64_t* k_full = bars + 1;     // [2]
  uint64_t* k_empty = bars + 3;    // [2]
  uint64_t* v_full = bars + 5;     // [2]
  uint64_t* v_empty = bars + 7;    // [2]
  uint64_t* s_full = bars + 9;     // [2] per query tile
  uint64_t* p_full = bars + 11;    // [2]
  uint64_t* o_full = bars + 13;    // [2]
  uint32_t* tmem_slot = (uint32_t*)(bars + 15);

  // warp index through a shuffle: tells the compiler it is warp-uniform, so the role branches are uniform branches and the MMA
  // warp's descriptor arithmetic can live in uniform registers (CUTLASS's canonical_warp_idx_sync trick)
  const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0), lane = threadIdx.x & 31;
  const int q0 = blockIdx.x * 2 * ATT_BM;
  const int head = blockIdx.y;
  const int kv_start = (q0 / p.Wq) * p.Wkv;
  const int n_tiles = p.Wkv / ATT_BN;

  if (warp == 0 && lane == 0) {
    tc::prefetch_tmap(&tq);
    tc::prefetch_tmap(&tk);
    tc::prefetch_tmap(&tv);
  }
  if (warp == 1 && lane == 0) {
    tc::mbar_init(q_full, 1);
    for (int i = 0; i < 2; i++) {
      tc::mbar_init(&k_full[i], 1);
      tc::mbar_init(&k_empty[i], 1);
      tc::mbar_init(&v_full[i], 1);
      tc::mbar_init(&v_empty[i], 1);
      tc::mbar_init(&s_full[i], 1);
      tc::mbar_init(&p_full[i], 8);  // one arrival per softmax warp of the tile
      tc::mbar_init(&o_full[i], 1);
    }
    tc::fence_barrier_init();
  }
  if (warp == 2) tc::tmem_alloc<512>(tmem_slot);
  if (p.trace)
    for (int i = threadIdx.x; i < ATT_TRACE_SLOTS; i += AG2_THREADS) tbuf[i] = 0ull;
  tc::tc_fence_before();
  __syncthreads();
  tc::tc_fence_after();
  const uint32_t tmem_base = __shfl_sync(0xffffffffu, *tmem_slot, 0);

  if (warp == 0) {
    if (lane == 0) {
      tc::mbar_arrive_expect_tx(q_full, 2 * ATT_TILE_BYTES);
      for (int g = 0; g < 2; g++) {
        tc::tma_load_2d(sQ + g * ATT_TILE_BYTES, &tq, q_full, p.qcol0 + head * ATT_D, q0 + g * ATT_BM);
        if (D > 64) tc::tma_load_2d(sQ + g * ATT_TILE_BYTES + ATT_BM * 128, &tq, q_full, p.qcol0 + head * ATT_D + 64, q0 + g * ATT_BM);
      }
      for (int j = 0; j < n_tiles; j++) {
        const int s = j & 1;
        const uint32_t par = ((j >> 1) & 1) ^ 1;
        const int row = kv_start + j * ATT_BN;
        tc::mbar_wait(&k_empty[s], par);
        tc::mbar_arrive_expect_tx(&k_full[s], ATT_TILE_BYTES);
        uint8_t* k = sK + s * ATT_TILE_BYTES;
        tc::tma_load_2d(k, &tk, &k_full[s], p.kcol0 + head * ATT_D, row);
        if (D > 64) tc::tma_load_2d(k + ATT_BN * 128, &tk, &k_full[s], p.kcol0 + head * ATT_D + 64, row);
        tc::mbar_wait(&v_empty[s], par);
        tc::mbar_arrive_expect_tx(&v_full[s], ATT_TILE_BYTES);
        uint8_t* v = sV + s * ATT_TILE_BYTES;
        tc::tma_load_2d(v, &tv, &v_full[s], p.vcol0 + head * ATT_D, row);
        if (D > 64) tc::tma_load_2d(v + ATT_BN * 128, &tv, &v_full[s], p.vcol0 + head * ATT_D + 64, row);
      }
    }
  } else if (warp == 1) {
    const uint32_t idesc_qk = tc::idesc_bf16(ATT_BM, ATT_BN, false, false, p.fp16 != 0);
    const uint32_t idesc_pv = tc::idesc_bf16(ATT_BM, ATT_D, false, true, p.fp16 != 0);  // A = P (TMEM, K-major), B = V MN-major
    auto issue_qk = [&](int g, int j, uint64_t* also) {
      const uint64_t dq = tc::desc_kmajor(tc::smem_u32(sQ + g * ATT_TILE_BYTES));
      const uint64_t dk = tc::desc_kmajor(tc::smem_u32(sK + (j & 1) * ATT_TILE_BYTES));
      if (tc::elect_one()) {
#pragma unroll
        for (int k = 0; k < ATT_D / 16; k++) {
          const uint32_t off = ((k >> 2) * (ATT_BM * 128) + (k & 3) * 32) >> 4;  // descriptor address units of 16 bytes
          tc::mma_f16_ss(tmem_base + g * 128, dq + off, dk + off, idesc_qk, k > 0 ? 1u : 0u);
        }
        tc::mma_commit(&s_full[g]);
        if (also) tc::mma_commit(also);
      }
      __syncwarp();
    };
    auto issue_pv = [&](int g, int j, uint64_t* also) {
      const uint64_t dv = tc::smem_desc_sw128(tc::smem_u32(sV + (j & 1) * ATT_TILE_BYTES), ATT_BN * 128, 1024);
      if (tc::elect_one()) {
#pragma unroll
        for (int k = 0; k < ATT_BN / 16; k++)
          // A: 16 keys = 8 packed columns; keys 0-63 sit at columns 0-31 of the tile's S region, keys 64-127 at columns 64-95
          tc::mma_f16_ts(tmem_base + 256 + g * ATT_D, tmem_base + g * 128 + (k >> 2) * 64 + (k & 3) * 8, dv + (k * 2048 >> 4), idesc_pv,
                         (j > 0 || k > 0) ? 1u : 0u);
        tc::mma_commit(&o_full[g]);
        if (also) tc::mma_commit(also);
      }
      __syncwarp();
    };
    tc::mbar_wait(q_full, 0);
    tc::mbar_wait(&k_full[0], 0);
    tc::tc_fence_after();
    issue_qk(0, 0, nullptr);
    issue_qk(1, 0, &k_empty[0]);
    for (int j = 0; j < n_tiles; j++) {
      const bool more = j + 1 < n_tiles;
      tc::mbar_wait(&v_full[j & 1], (j >> 1) & 1);
      tc::mbar_wait(&p_full[0], j & 1);
      tc::tc_fence_after();
      if (lane == 0) attn_trace(p, tbuf, 1, j * 2);
      issue_pv(0, j, nullptr);
      if (lane == 0) attn_trace(p, tbuf, 6, j * 2);
      if (more) {
        tc::mbar_wait(&k_full[(j + 1) & 1], ((j + 1) >> 1) & 1);
        tc::tc_fence_after();
        if (lane == 0) attn_trace(p, tbuf, 7, j * 2);
        issue_qk(0, j + 1, nullptr);
      }
      if (lane == 0) attn_trace(p, tbuf, 2, j * 2);
      tc::mbar_wait(&p_full[1], j & 1);
      tc::tc_fence_after();
      if (lane == 0) attn_trace(p, tbuf, 1, j * 2 + 1);
      issue_pv(1, j, &v_empty[j & 1]);
      if (lane == 0) attn_trace(p, tbuf, 6, j * 2 + 1);
      if (more) issue_qk(1, j + 1, &k_empty[(j + 1) & 1]);
      if (lane == 0) attn_trace(p, tbuf, 2, j * 2 + 1);
    }
  } else if (warp >= 4) {
    const int g = (warp - 4) >> 3;         // query tile of this warp
    const int ch = ((warp - 4) >> 2) & 1;  // column half of the score tile
    const int quad = warp & 3;
    const int r = quad * 32 + lane;
    const int grow = q0 + g * ATT_BM + r;
    const int pair_id = 1 + g * 4 + quad;  // named barrier of the two warps that share these 32 rows
    const uint32_t lane_sel = (uint32_t)(quad * 32) << 16;
    const uint32_t tS = tmem_base + lane_sel + g * 128 + ch * 64;   // my 64 score columns; my P goes over their first 32
    const uint32_t tO = tmem_base + lane_sel + 256 + g * ATT_D + ch * DH;
    float* xme = xch + (g * 2 + ch) * 128 + r;        // + slot * 512
    float* xpt = xch + (g * 2 + (ch ^ 1)) * 128 + r;
    float m_ref = 0.f, l = 0.f;
    for (int j = 0; j < n_tiles; j++) {
      tc::mbar_wait(&s_full[g], j & 1);
      tc::tc_fence_after();
      if (ch == 0 && quad == 0 && lane == 0) attn_trace(p, tbuf, 4, j * 2 + g);
      if (j == 0) {
        float mx = -INFINITY;
#pragma unroll
        for (int c = 0; c < 2; c++) {
          uint32_t v[32];
          tc::tmem_ld_32x32(tS + c * 32, v);
          tc::tmem_ld_wait();
#pragma unroll
          for (int i = 0; i < 32; i++) mx = fmaxf(mx, __uint_as_float(v[i]));
        }
        xme[2 * 512] = mx;
        pair_sync(pair_id);
        mx = fmaxf(mx, xpt[2 * 512]);
        m_ref = mx * p.scale_log2;
      }
      float tmax = -INFINITY;
      uint64_t rs2 = pk2(0.f, 0.f);
      const uint64_t sc2 = pk2(p.scale_log2, p.scale_log2), mr2 = pk2(-m_ref, -m_ref);
      uint32_t v[2][32];
      tc::tmem_ld_32x32(tS, v[0]);
#pragma unroll
      for (int c = 0; c < 2; c++) {
        tc::tmem_ld_wait();
        if (c + 1 < 2) tc::tmem_ld_32x32(tS + (c + 1) * 32, v[(c + 1) & 1]);
        uint32_t pk[16];
        float pe[32];
#pragma unroll
        for (int i = 0; i < 32; i += 2) {
          const float s0 = __uint_as_float(v[c & 1][i]), s1 = __uint_as_float(v[c & 1][i + 1]);
          asm("max.f32 %0, %0, %1, %2;" : "+f"(tmax) : "f"(s0), "f"(s1));
          float x0, x1;
          up2(fma2(pk2(s0, s1), sc2, mr2), x0, x1);
          if (EXP_FMA && ((i >> 1) % 3) == 2) {  // every third pair: exponentials on the FMA pipe
            ex2_poly2(x0, x1, pe[i], pe[i + 1]);
          } else {
            pe[i] = ex2(x0);
            pe[i + 1] = ex2(x1);
          }
          const uint64_t pp = pk2(pe[i], pe[i + 1]);
          asm("add.rn.f32x2 %0, %0, %1;" : "+l"(rs2) : "l"(pp));
        }
        if (p.fp16) {
#pragma unroll
          for (int i = 0; i < 16; i++) pk[i] = tc::pack16(1, pe[2 * i], pe[2 * i + 1]);
        } else {
#pragma unroll
          for (int i = 0; i < 16; i++) pk[i] = tc::pack16(0, pe[2 * i], pe[2 * i + 1]);
        }
        tc::tmem_st_32x16(tS + c * 16, pk);  // P over the already-consumed head of my own score columns
      }
      float rs_lo, rs_hi;
      up2(rs2, rs_lo, rs_hi);
      const float rowsum = rs_lo + rs_hi;
      tc::tmem_st_wait();
      tc::tc_fence_before();
      __syncwarp();
      if (lane == 0) tc::mbar_arrive(&p_full[g]);
      if (ch == 0 && quad == 0 && lane == 0) attn_trace(p, tbuf, 5, j * 2 + g);
      l += rowsum;
      // both threads of the row must take the same rescale decision: exchange the half-row maxima (slot = tile parity)
      xme[(j & 1) * 512] = tmax;
      pair_sync(pair_id);
      tmax = fmaxf(tmax, xpt[(j & 1) * 512]);
      const float tm = tmax * p.scale_log2;
      const bool need = tm > m_ref + 8.0f;
      if (__any_sync(0xffffffffu, need)) {
        // rescale my half of O (and my partial l) to the new reference once P V of this tile has been accumulated
        tc::mbar_wait(&o_full[g], j & 1);
        tc::tc_fence_after();
        const float alpha = need ? ex2(m_ref - tm) : 1.0f;
#pragma unroll
        for (int c = 0; c < DH / 16; c++) {
          uint32_t w[16];
          tc::tmem_ld_32x16(tO + c * 16, w);
          tc::tmem_ld_wait();
#pragma unroll
          for (int i = 0; i < 16; i++) w[i] = __float_as_uint(__uint_as_float(w[i]) * alpha);
          tc::tmem_st_32x16(tO + c * 16, w);
        }
        tc::tmem_st_wait();
        tc::tc_fence_before();
        l *= alpha;
        if (need) m_ref = tm;
      }
    }
    tc::mbar_wait(&o_full[g], (n_tiles - 1) & 1);
    tc::tc_fence_after();
    xme[2 * 512] = l;  // (slot 2 was last used before tile 0's first barrier)
    pair_sync(pair_id);
    l += xpt[2 * 512];
    const float inv = 1.f / l;
    __nv_bfloat16* o = p.out + (long long)grow * p.ld_out + head * ATT_D + ch * DH;
#pragma unroll
    for (int c = 0; c < DH / 16; c++) {
      uint32_t w[16];
      tc::tmem_ld_32x16(tO + c * 16, w);
      tc::tmem_ld_wait();
      if (grow < p.Mq) {
        uint32_t q[8];
#pragma unroll
        for (int k = 0; k < 8; k++) q[k] = tc::pack16(p.fp16, __uint_as_float(w[2 * k]) * inv, __uint_as_float(w[2 * k + 1]) * inv);
        *(uint4*)(o + c * 16) = make_uint4(q[0], q[1], q[2], q[3]);
        *(uint4*)(o + c * 16 + 8) = make_uint4(q[4], q[5], q[6], q[7]);
      }
    }
  }
  tc::tc_fence_before();
  __syncthreads();
  attn_trace_dump(p, tbuf);
  if (warp == 2) {
    tc::tc_fence_after();
    tc::tmem_dealloc<512>(tmem_base);
  }
}

// ------------------------------------------------------------------------------------------------ windowed attention
// Persistent kernel for every block-diagonal case that is not global (Hiera's 8x8 / 4x4 / 14x14 / 7x7 windows and their
// Q-pooled variants).  These are HBM-bound (a few hundred keys per query), so the design goal is streaming efficiency:
//   * work item = (128-row query tile, head); a CTA walks pairs of items, one item per "slot" (slot = S/O TMEM region +
//     Q buffer + 4 softmax warps), so the softmax of one item overlaps the loads and MMAs of the other;
//   * K/V tiles of both slots flow through ONE 4-entry TMA ring in exactly the order the MMA warp consumes them;
//   * P stays in TMEM (TS MMA), O accumulates in TMEM; per-row window masking; the reference keeps zero-pad tokens as
//     ordinary keys, so padding is NOT masked — only keys of other windows are.
// The running maximum is fixed at the first tile that has visible keys and O / l are rescaled only when a later tile
// exceeds it by more than 2^8 (same exactness argument as the global kernel).
constexpr int AW_THREADS = 384;
// K/V ring depth: the ring feeds both slots in MMA consumption order and a V tile stays in it until its P is ready, so
// a shallow ring serialises the TMA latency behind the softmax; 5 tiles of 32 KB (head dim 96) / 8 of 16 KB (64) fit.
template <int D>
constexpr int aw_ring() { return D > 64 ? 5 : 8; }
static_assert(1024 + 32768 * (2 + 5) + 256 <= 232448, "windowed attention shared memory");
template <int D>
constexpr int aw_smem() { return 1024 + att_tile_bytes<D>() * (2 + aw_ring<D>()) + 256; }

struct WinItem {
  int q0, kv_lo, n_kt, head, valid;
};
__device__ __forceinline__ WinItem win_item(const AttnParams& p, int n_items, int it) {
  WinItem w;
  w.valid = it < n_items;
  int qt = it / p.heads;
  w.head = it - qt * p.heads;
  w.q0 = qt * p.q_rows;
  int r_last = min(w.q0 + p.q_rows - 1, p.Mq - 1);
  w.kv_lo = (w.q0 / p.Wq) * p.Wkv;
  int kv_hi = (r_last / p.Wq + 1) * p.Wkv;
  w.n_kt = w.valid ? (kv_hi - w.kv_lo + ATT_BN - 1) / ATT_BN : 0;
  return w;
}

// UNI: every warp's rows lie in one window (8 x 8 windows with 128-row items, 14 x 14 with window-aligned 98-row items), so whole
// 32-column chunks are visible to all of them and skip the per-element visibility tests (they were half of the softmax
// instructions); the other geometries keep the masked code only (a kernel carrying both paths ran them 10-15 % slower).
template <int D, bool UNI>
__global__ void __launch_bounds__(AW_THREADS, 1)
k_attn_win(const __grid_constant__ CUtensorMap tq, const __grid_constant__ CUtensorMap tk,
           const __grid_constant__ CUtensorMap tv, AttnParams p, int n_items) {
  constexpr int ATT_D = D, ATT_TILE_BYTES = att_tile_bytes<D>(), AW_RING = aw_ring<D>();
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  uint8_t* sQ = smem;                       // [2]
  uint8_t* sR = sQ + 2 * ATT_TILE_BYTES;    // K/V ring [AW_RING]
  uint64_t* bars = (uint64_t*)(sR + AW_RING * ATT_TILE_BYTES);
  uint64_t* q_full = bars;                  // [2]
  uint64_t* q_empty = bars + 2;             // [2]
  uint64_t* r_full = bars + 4;              // [AW_RING]
  uint64_t* r_empty = bars + 4 + AW_RING;   // [AW_RING]
  uint64_t* s_full = bars + 4 + 2 * AW_RING;  // [2]
  uint64_t* p_full = s_full + 2;            // [2]
  uint64_t* o_full = s_full + 4;            // [2]
  uint64_t* o_free = s_full + 6;            // [2]
  uint32_t* tmem_slot = (uint32_t*)(s_full + 8);

  const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0), lane = threadIdx.x & 31;  // shuffle: warp-uniform for the compiler
  const int n_pairs = (n_items + 1) / 2;

  if (warp == 0 && lane == 0) {
    tc::prefetch_tmap(&tq);
    tc::prefetch_tmap(&tk);
    tc::prefetch_tmap(&tv);
  }
  if (warp == 1 && lane == 0) {
    for (int i = 0; i < 2; i++) {
      tc::mbar_init(&q_full[i], 1);
      tc::mbar_init(&q_empty[i], 1);
      tc::mbar_init(&s_full[i], 1);
      tc::mbar_init(&p_full[i], 4);
      tc::mbar_init(&o_full[i], 1);
      tc::mbar_init(&o_free[i], 4);
    }
    for (int i = 0; i < AW_RING; i++) {
      tc::mbar_init(&r_full[i], 1);
      tc::mbar_init(&r_empty[i], 1);
    }
    tc::fence_barrier_init();
  }
  if (warp == 2) tc::tmem_alloc<512>(tmem_slot);
  tc::tc_fence_before();
  __syncthreads();
  tc::tc_fence_after();
  const uint32_t tmem_base = __shfl_sync(0xffffffffu, *tmem_slot, 0);

  if (warp == 0) {
    // ===================== TMA producer: Q per slot, K/V through the ring in MMA consumption order
    if (lane == 0) {
      uint32_t ring = 0, ring_uses = 0;
      uint32_t q_uses[2] = {0, 0};
      auto load_kv = [&](const CUtensorMap* tm, int col, int row) {
        tc::mbar_wait(&r_empty[ring], ((ring_uses / AW_RING) & 1) ^ 1);
        tc::mbar_arrive_expect_tx(&r_full[ring], ATT_TILE_BYTES);
        uint8_t* dst = sR + ring * ATT_TILE_BYTES;
        tc::tma_load_2d(dst, tm, &r_full[ring], col, row);
        if (D > 64) tc::tma_load_2d(dst + ATT_BN * 128, tm, &r_full[ring], col + 64, row);
        ring = (ring + 1) % AW_RING;
        ring_uses++;
      };
      for (int pp = blockIdx.x; pp < n_pairs; pp += gridDim.x) {
        WinItem it[2] = {win_item(p, n_items, 2 * pp), win_item(p, n_items, 2 * pp + 1)};
        for (int g = 0; g < 2; g++)
          if (it[g].valid) {
            tc::mbar_wait(&q_empty[g], (q_uses[g] & 1) ^ 1);
            tc::mbar_arrive_expect_tx(&q_full[g], ATT_TILE_BYTES);
            uint8_t* dst = sQ + g * ATT_TILE_BYTES;
            tc::tma_load_2d(dst, &tq, &q_full[g], p.qcol0 + it[g].head * ATT_D, it[g].q0);
            if (D > 64) tc::tma_load_2d(dst + ATT_BM * 128, &tq, &q_full[g], p.qcol0 + it[g].head * ATT_D + 64, it[g].q0);
            q_uses[g]++;
          }
        for (int g = 0; g < 2; g++)
          if (it[g].n_kt > 0) load_kv(&tk, p.kcol0 + it[g].head * ATT_D, it[g].kv_lo);
        const int nmax = max(it[0].n_kt, it[1].n_kt);
        for (int j = 0; j < nmax; j++)
          for (int g = 0; g < 2; g++)
            if (j < it[g].n_kt) {
              load_kv(&tv, p.vcol0 + it[g].head * ATT_D, it[g].kv_lo + j * ATT_BN);
              if (j + 1 < it[g].n_kt) load_kv(&tk, p.kcol0 + it[g].head * ATT_D, it[g].kv_lo + (j + 1) * ATT_BN);
            }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer (converged warp, one elected lane issues: see k_attn_global)
    {
      const uint32_t idesc_qk = tc::idesc_bf16(ATT_BM, ATT_BN, false, false, p.fp16 != 0);
      const uint32_t idesc_pv = tc::idesc_bf16(ATT_BM, ATT_D, false, true, p.fp16 != 0);
      uint32_t ring = 0, ring_uses = 0;
      uint32_t q_uses[2] = {0, 0}, p_uses[2] = {0, 0}, item_uses[2] = {0, 0};
      auto ring_wait = [&]() -> uint32_t {
        tc::mbar_wait(&r_full[ring], (ring_uses / AW_RING) & 1);
        tc::tc_fence_after();
        return tc::smem_u32(sR + ring * ATT_TILE_BYTES);
      };
      auto ring_advance = [&]() {
        ring = (ring + 1) % AW_RING;
        ring_uses++;
      };
      auto issue_qk = [&](int g, bool last) {
        const uint64_t dk = tc::desc_kmajor(ring_wait());
        const uint64_t dq = tc::desc_kmajor(tc::smem_u32(sQ + g * ATT_TILE_BYTES));
        if (tc::elect_one()) {
#pragma unroll
          for (int k = 0; k < ATT_D / 16; k++) {
            const uint32_t off = ((k >> 2) * (ATT_BM * 128) + (k & 3) * 32) >> 4;
            tc::mma_f16_ss(tmem_base + g * 128, dq + off, dk + off, idesc_qk, k > 0 ? 1u : 0u);
          }
          tc::mma_commit(&s_full[g]);
          tc::mma_commit(&r_empty[ring]);
          if (last) tc::mma_commit(&q_empty[g]);
        }
        __syncwarp();
        ring_advance();
      };
      for (int pp = blockIdx.x; pp < n_pairs; pp += gridDim.x) {
        WinItem it[2] = {win_item(p, n_items, 2 * pp), win_item(p, n_items, 2 * pp + 1)};
        for (int g = 0; g < 2; g++)
          if (it[g].n_kt > 0) {
            tc::mbar_wait(&q_full[g], q_uses[g] & 1);
            q_uses[g]++;
            issue_qk(g, it[g].n_kt == 1);
          }
        const int nmax = max(it[0].n_kt, it[1].n_kt);
        for (int j = 0; j < nmax; j++)
          for (int g = 0; g < 2; g++)
            if (j < it[g].n_kt) {
              tc::mbar_wait(&p_full[g], p_uses[g] & 1);
              p_uses[g]++;
              if (j == 0) {
                // the previous item of this slot has been read out of O
                tc::mbar_wait(&o_free[g], (item_uses[g] & 1) ^ 1);
                item_uses[g]++;
              }
              tc::tc_fence_after();
              const uint64_t dv = tc::smem_desc_sw128(ring_wait(), ATT_BN * 128, 1024);
              if (tc::elect_one()) {
#pragma unroll
                for (int k = 0; k < ATT_BN / 16; k++)
                  tc::mma_f16_ts(tmem_base + 256 + g * ATT_D, tmem_base + g * 128 + k * 8, dv + (k * 2048 >> 4), idesc_pv,
                                 (j > 0 || k > 0) ? 1u : 0u);
                tc::mma_commit(&o_full[g]);
                tc::mma_commit(&r_empty[ring]);
              }
              __syncwarp();
              ring_advance();
              if (j + 1 < it[g].n_kt) issue_qk(g, j + 2 == it[g].n_kt);
            }
      }
    }
  } else if (warp >= 4) {
    // ===================== softmax + output: warps 4-7 slot 0, warps 8-11 slot 1
    const int g = (warp - 4) >> 2;
    const int quad = warp & 3;
    const int r = quad * 32 + lane;
    const uint32_t lane_sel = (uint32_t)(quad * 32) << 16;
    const uint32_t tS = tmem_base + lane_sel + g * 128;
    const uint32_t tO = tmem_base + lane_sel + 256 + g * ATT_D;
    uint32_t s_uses = 0, pv_done = 0;  // cumulative tiles of this slot
    for (int pp = blockIdx.x; pp < n_pairs; pp += gridDim.x) {
      const WinItem it = win_item(p, n_items, 2 * pp + g);
      if (it.n_kt == 0) continue;
      const int grow = it.q0 + r;
      const bool row_ok = grow < p.Mq && r < p.q_rows;  // the TMA box always brings 128 rows; the tail belongs to the next item
      const int wq = (row_ok ? grow : it.q0) / p.Wq;
      const int vis0 = wq * p.Wkv - it.kv_lo;  // first visible key relative to the item's first key tile
      float m_ref = -INFINITY, l = 0.f;
      for (int j = 0; j < it.n_kt; j++) {
        const int c_lo = max(vis0 - j * ATT_BN, 0), c_hi = row_ok ? min(vis0 + p.Wkv - j * ATT_BN, ATT_BN) : 0;
        // 32-column chunks of this key tile that at least one row of the WARP can see: with small windows most of the
        // 128 x 128 score tile is other windows' keys (a warp of the Q-pooled 16/64 case sees one key tile in four, a
        // warp of the 8 x 8 case two chunks in four), and those chunks need neither the maximum nor the exponentials —
        // only zeros in P.
        const bool sees = c_hi > c_lo;
        const int cb = p.skip_chunks ? (__reduce_min_sync(0xffffffffu, sees ? c_lo : ATT_BN) >> 5) : 0;
        const int ce = p.skip_chunks ? ((__reduce_max_sync(0xffffffffu, sees ? c_hi : 0) + 31) >> 5) : ATT_BN / 32;
        // columns every row of the warp sees: chunks inside [u_lo, u_hi) need no per-element visibility test (all chunks of the
        // 8 x 8 and 14 x 14 window cases except the last partial one; the masks were half of the softmax instructions there)
        // (rows past the item's last row are never stored: whatever they compute is "don't care", they do not restrict the range)
        const int u_lo = UNI ? __reduce_max_sync(0xffffffffu, row_ok ? (sees ? c_lo : ATT_BN) : 0) : ATT_BN;
        const int u_hi = UNI ? __reduce_min_sync(0xffffffffu, row_ok ? (sees ? c_hi : 0) : ATT_BN) : 0;
        tc::mbar_wait(&s_full[g], s_uses & 1);
        s_uses++;
        tc::tc_fence_after();
        float tmx = -INFINITY;
#pragma unroll
        for (int c = 0; c < ATT_BN / 32; c++) {
          if (c < cb || c >= ce) continue;
          uint32_t v[32];
          tc::tmem_ld_32x32(tS + c * 32, v);
          tc::tmem_ld_wait();
          if (UNI && c * 32 >= u_lo && c * 32 + 32 <= u_hi) {
            float m0 = __uint_as_float(v[0]), m1 = __uint_as_float(v[1]), m2 = __uint_as_float(v[2]), m3 = __uint_as_float(v[3]);
#pragma unroll
            for (int i = 4; i < 32; i += 4) {
              m0 = fmaxf(m0, __uint_as_float(v[i]));
              m1 = fmaxf(m1, __uint_as_float(v[i + 1]));
              m2 = fmaxf(m2, __uint_as_float(v[i + 2]));
              m3 = fmaxf(m3, __uint_as_float(v[i + 3]));
            }
            tmx = fmaxf(tmx, fmaxf(fmaxf(m0, m1), fmaxf(m2, m3)));
          } else {
#pragma unroll
            for (int i = 0; i < 32; i++) {
              int col = c * 32 + i;
              if (col >= c_lo && col < c_hi) tmx = fmaxf(tmx, __uint_as_float(v[i]));
            }
          }
        }
        const float tm = tmx * p.scale_log2;  // -inf when the row sees nothing in this tile
        const bool first = (m_ref == -INFINITY) && (tmx != -INFINITY);
        const bool need = (m_ref != -INFINITY) && (tm > m_ref + 8.0f);
        if (first) m_ref = tm;
        if (__any_sync(0xffffffffu, need)) {
          // O holds the tiles accumulated so far: wait for the last P V of this item, then rescale
          tc::mbar_wait(&o_full[g], (pv_done - 1) & 1);
          tc::tc_fence_after();
          const float alpha = need ? ex2(m_ref - tm) : 1.0f;
#pragma unroll
          for (int c = 0; c < ATT_D / 32; c++) {
            uint32_t v[32];
            tc::tmem_ld_32x32(tO + c * 32, v);
            tc::tmem_ld_wait();
#pragma unroll
            for (int i = 0; i < 32; i++) v[i] = __float_as_uint(__uint_as_float(v[i]) * alpha);
            tc::tmem_st_32x32(tO + c * 32, v);
          }
          tc::tmem_st_wait();
          l *= alpha;
          if (need) m_ref = tm;
        }
        float rowsum = 0.f;
        const float mr = (m_ref == -INFINITY) ? 0.f : m_ref;
        // P overwrites S in place (16-bit, half the columns): every visible chunk is read before any chunk is written
        uint32_t pk[ATT_BN / 32][16];
#pragma unroll
        for (int c = 0; c < ATT_BN / 32; c++) {
          if (c < cb || c >= ce) {
#pragma unroll
            for (int i = 0; i < 16; i++) pk[c][i] = 0u;
            continue;
          }
          uint32_t v[32];
          tc::tmem_ld_32x32(tS + c * 32, v);
          tc::tmem_ld_wait();
          if (UNI && c * 32 >= u_lo && c * 32 + 32 <= u_hi) {
            float rs0 = 0.f, rs1 = 0.f;
#pragma unroll
            for (int i = 0; i < 32; i += 2) {
              const float p0 = ex2(fmaf(__uint_as_float(v[i]), p.scale_log2, -mr));
              const float p1 = ex2(fmaf(__uint_as_float(v[i + 1]), p.scale_log2, -mr));
              rs0 += p0;
              rs1 += p1;
              pk[c][i >> 1] = tc::pack16(p.fp16, p0, p1);
            }
            rowsum += rs0 + rs1;
          } else {
#pragma unroll
            for (int i = 0; i < 32; i += 2) {
              const int col = c * 32 + i;
              float p0 = ex2(fmaf(__uint_as_float(v[i]), p.scale_log2, -mr));
              float p1 = ex2(fmaf(__uint_as_float(v[i + 1]), p.scale_log2, -mr));
              if (col < c_lo || col >= c_hi) p0 = 0.f;
              if (col + 1 < c_lo || col + 1 >= c_hi) p1 = 0.f;
              rowsum += p0 + p1;
              pk[c][i >> 1] = tc::pack16(p.fp16, p0, p1);
            }
          }
        }
#pragma unroll
        for (int c = 0; c < ATT_BN / 32; c++) tc::tmem_st_32x16(tS + c * 16, pk[c]);
        tc::tmem_st_wait();
        tc::tc_fence_before();
        __syncwarp();
        if (lane == 0) tc::mbar_arrive(&p_full[g]);
        l += rowsum;
        pv_done++;
      }
      tc::mbar_wait(&o_full[g], (pv_done - 1) & 1);
      tc::tc_fence_after();
      const float inv = l > 0.f ? 1.f / l : 0.f;
      __nv_bfloat16* o = p.out + (long long)grow * p.ld_out + it.head * ATT_D;
#pragma unroll
      for (int c = 0; c < ATT_D / 32; c++) {
        uint32_t v[32];
        tc::tmem_ld_32x32(tO + c * 32, v);
        tc::tmem_ld_wait();
        if (row_ok) {
#pragma unroll
          for (int i = 0; i < 32; i += 8) {
            uint32_t w[4];
#pragma unroll
            for (int k = 0; k < 4; k++)
              w[k] = tc::pack16(p.fp16, __uint_as_float(v[i + 2 * k]) * inv, __uint_as_float(v[i + 2 * k + 1]) * inv);
            *(uint4*)(o + c * 32 + i) = make_uint4(w[0], w[1], w[2], w[3]);
          }
        }
      }
      tc::tc_fence_before();
      __syncwarp();
      if (lane == 0) tc::mbar_arrive(&o_free[g]);
    }
  }
  tc::tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    tc::tc_fence_after();
    tc::tmem_dealloc<512>(tmem_base);
  }
}

int device_sm_count();
static unsigned long long* g_attn_trace = nullptr;

template <int D>
static int attn_launch_d(const CUtensorMap& tq, const CUtensorMap& tk, const CUtensorMap& tv, const AttnParams& p, bool glob,
                         cudaStream_t st) {
  if (glob) {
    static std::atomic<unsigned long long> attr_set_g{0};
    if (cvb_once_per_device(attr_set_g)) {
      cudaError_t e = cudaFuncSetAttribute(k_attn_global<D, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, ag_smem<D>());
      if (e == cudaSuccess) e = cudaFuncSetAttribute(k_attn_global<D, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, ag_smem<D>());
      if (e != cudaSuccess) return cvb_fail_cuda(e, "cudaFuncSetAttribute(k_attn_global)");
    }
    // CVB_ATTN_EXPFMA=0: every exponential on the XU pipe (A/B switch)
    static const bool exp_fma = getenv("CVB_ATTN_EXPFMA") ? atoi(getenv("CVB_ATTN_EXPFMA")) != 0 : true;
    // two softmax threads per row (k_attn_global2): same speed within run-to-run noise (5.47-5.97 ms against 5.47-5.81 ms per
    // step over repeated same-box runs), so the one-thread-per-row kernel stays the default; scripts/attn_trace.py uses this one
    static const bool two = getenv("CVB_ATTN_G2") ? atoi(getenv("CVB_ATTN_G2")) != 0 : false;
    if (two) {
      constexpr int smem2 = ag_smem<D>() + 3 * 2 * 2 * 128 * 4 + ATT_TRACE_SLOTS * 8;
      static std::atomic<unsigned long long> attr_set_2{0};
      if (cvb_once_per_device(attr_set_2)) {
        cudaError_t e = cudaFuncSetAttribute(k_attn_global2<D, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem2);
        if (e == cudaSuccess) e = cudaFuncSetAttribute(k_attn_global2<D, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem2);
        if (e != cudaSuccess) return cvb_fail_cuda(e, "cudaFuncSetAttribute(k_attn_global2)");
      }
      if (exp_fma) {
        CVB_LAUNCH((k_attn_global2<D, true>), dim3(p.Mq / (2 * ATT_BM), p.heads), dim3(AG2_THREADS), smem2, st, tq, tk, tv, p);
      } else {
        CVB_LAUNCH((k_attn_global2<D, false>), dim3(p.Mq / (2 * ATT_BM), p.heads), dim3(AG2_THREADS), smem2, st, tq, tk, tv, p);
      }
      return CV_OK;
    }
    if (exp_fma) {
      CVB_LAUNCH((k_attn_global<D, true>), dim3(p.Mq / (2 * ATT_BM), p.heads), dim3(AG_THREADS), ag_smem<D>(), st, tq, tk, tv, p);
    } else {
      CVB_LAUNCH((k_attn_global<D, false>), dim3(p.Mq / (2 * ATT_BM), p.heads), dim3(AG_THREADS), ag_smem<D>(), st, tq, tk, tv, p);
    }
    return CV_OK;
  }
  static std::atomic<unsigned long long> attr_set_w{0};
  if (cvb_once_per_device(attr_set_w)) {
    cudaError_t e = cudaFuncSetAttribute(k_attn_win<D, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, aw_smem<D>());
    if (e == cudaSuccess) e = cudaFuncSetAttribute(k_attn_win<D, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, aw_smem<D>());
    if (e != cudaSuccess) return cvb_fail_cuda(e, "cudaFuncSetAttribute(k_attn_win)");
  }
  const int n_items = ((p.Mq + p.q_rows - 1) / p.q_rows) * p.heads;
  const int n_pairs = (n_items + 1) / 2;
  const int grid = n_pairs < device_sm_count() ? n_pairs : device_sm_count();
  // rows 32 w .. 32 w + 31 of an item in one window: windows of a multiple of 32 rows with 32-aligned items, or items that do
  // not cross a window at all
  static const int uni_on = getenv("CVB_ATTN_UNI") ? atoi(getenv("CVB_ATTN_UNI")) : 1;
  const bool uni = uni_on && ((p.Wq % 32 == 0 && p.q_rows % 32 == 0) || (p.Wq % p.q_rows == 0));
  if (uni) {
    CVB_LAUNCH((k_attn_win<D, true>), dim3(grid), dim3(AW_THREADS), aw_smem<D>(), st, tq, tk, tv, p, n_items);
  } else {
    CVB_LAUNCH((k_attn_win<D, false>), dim3(grid), dim3(AW_THREADS), aw_smem<D>(), st, tq, tk, tv, p, n_items);
  }
  return CV_OK;
}

// q/k/v: 16-bit matrices [Mq|Mkv, ld*] whose columns [col0 + h*D, col0 + (h+1)*D) hold head h; D = 64 or 96 is the
// (zero-padded) head dim of the buffers, `scale` the softmax scale of the real head dim.
int attn_tc_launch(const __nv_bfloat16* q, long long ldq, int qcols, int qcol0, const __nv_bfloat16* k, long long ldk,
                   int kcols, int kcol0, const __nv_bfloat16* v, long long ldv, int vcols, int vcol0, int Mq, int Mkv,
                   int Wq, int Wkv, int heads, int D, float scale, __nv_bfloat16* out, long long ld_out, int fp16,
                   cudaStream_t st) {
  if (D != 64 && D != 96) return cvb_fail(CV_ERR_INVALID, "attention: padded head dim must be 64 or 96");
  if (Mq <= 0 || Mkv <= 0 || Wq <= 0 || Wkv <= 0 || heads <= 0) return cvb_fail(CV_ERR_INVALID, "attention: bad sizes");
  if ((long long)(Mq / Wq) * Wkv > Mkv || (Mq % Wq)) return cvb_fail(CV_ERR_INVALID, "attention: window counts of Q and K/V differ");
  if ((ldq % 8) || (ldk % 8) || (ldv % 8) || (ld_out % 8) || (qcol0 % 8) || (kcol0 % 8) || (vcol0 % 8))
    return cvb_fail(CV_ERR_INVALID, "attention: pitches / column offsets must be multiples of 8 elements");
  CUtensorMap tq, tk, tv;
  if (!tc_host::make_tmap_bf16(&tq, q, (uint64_t)Mq, (uint64_t)qcols, (uint64_t)ldq, ATT_BM) ||
      !tc_host::make_tmap_bf16(&tk, k, (uint64_t)Mkv, (uint64_t)kcols, (uint64_t)ldk, ATT_BN) ||
      !tc_host::make_tmap_bf16(&tv, v, (uint64_t)Mkv, (uint64_t)vcols, (uint64_t)ldv, ATT_BN))
    return cvb_fail(CV_ERR_CUDA, "cuTensorMapEncodeTiled failed (attention operands)");
  AttnParams p;
  p.Mq = Mq; p.Mkv = Mkv; p.Wq = Wq; p.Wkv = Wkv; p.heads = heads;
  p.qcol0 = qcol0; p.kcol0 = kcol0; p.vcol0 = vcol0;
  p.scale_log2 = scale * 1.4426950408889634f;
  p.fp16 = fp16;
  {
    // same-box A/B (tiny, B = 64): 8x8 windows 1.06 -> 0.72 ms, 4x4 windows 0.52 -> 0.39 ms, every other case equal or
    // slightly better; CVB_ATTN_SKIP = 0 turns it off
    static const int force = getenv("CVB_ATTN_SKIP") ? atoi(getenv("CVB_ATTN_SKIP")) : 1;
    p.skip_chunks = force != 0;
    // window-aligned items: 14 x 14 windows (196 rows) -> two items of 98 rows with exactly the window's 196 keys (two key
    // tiles instead of three to four); 7 x 7 windows (49 rows) -> 98 rows = two whole windows; windows that divide 128
    // keep 128-row items.  CVB_ATTN_ALIGN = 0 restores 128-row items everywhere.
    static const int align = getenv("CVB_ATTN_ALIGN") ? atoi(getenv("CVB_ATTN_ALIGN")) : 1;
    p.q_rows = ATT_BM;
    if (align) {
      if (Wq > ATT_BM) {
        const int parts = (Wq + ATT_BM - 1) / ATT_BM;
        if (Wq % parts == 0) p.q_rows = Wq / parts;
      } else if (ATT_BM % Wq) {
        p.q_rows = (ATT_BM / Wq) * Wq;
      }
    }
  }
  p.out = out; p.ld_out = ld_out;
  p.trace = g_attn_trace;
  const bool glob = Wq == Wkv && (Wkv % ATT_BN) == 0 && (Wq % (2 * ATT_BM)) == 0 && (Mq % (2 * ATT_BM)) == 0;
  // algorithmic flops: every query row against the keys of its own window, QK^T and PV (padded head dim as executed)
  cvb_next_work(4.0 * (double)Mq * (double)Wkv * D * heads);
  if (cvb_profile_on()) {
    char nm[96];
    snprintf(nm, sizeof(nm), "%s Mq%d Wq%d Wkv%d h%d d%d", glob ? "attn_global" : "attn", Mq, Wq, Wkv, heads, D);
    cvb_next_name(nm);
  }
  return D == 64 ? attn_launch_d<64>(tq, tk, tv, p, glob, st) : attn_launch_d<96>(tq, tk, tv, p, glob, st);
}

}  // namespace cvb

using namespace cvb;

extern "C" int cv_attn_set_trace(void* device_buffer) {
  g_attn_trace = (unsigned long long*)device_buffer;
  return CV_OK;
}

extern "C" int cv_attention_bf16(const void* qkv_q, long long ldq, int qcols, int qcol0, const void* qkv_k,
                                 long long ldk, int kcols, int kcol0, const void* qkv_v, long long ldv, int vcols,
                                 int vcol0, int Mq, int Mkv, int Wq, int Wkv, int heads, int head_dim, float scale,
                                 void* out, long long ld_out, void* stream) {
  cvb_reset_launches();
  if (!qkv_q || !qkv_k || !qkv_v || !out) return cvb_fail(CV_ERR_INVALID, "cv_attention_bf16: null pointer");
  if (head_dim != 64 && head_dim != 96) return cvb_fail(CV_ERR_INVALID, "cv_attention_bf16: head_dim must be 64 or 96 (pad 56 / 72 with zeros)");
  return attn_tc_launch((const __nv_bfloat16*)qkv_q, ldq, qcols, qcol0, (const __nv_bfloat16*)qkv_k, ldk, kcols, kcol0,
                        (const __nv_bfloat16*)qkv_v, ldv, vcols, vcol0, Mq, Mkv, Wq, Wkv, heads, head_dim, scale,
                        (__nv_bfloat16*)out, ld_out, 0, (cudaStream_t)stream);
}
