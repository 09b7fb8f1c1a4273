// Persistent warp-specialised bf16 GEMM on tcgen05 + TMA (interface: gemm_tc.cuh) + its C-ABI test entry.
//
// CTA = 384 threads: warp0 TMA producer (one lane), warp1 MMA issuer (one lane), warp2 TMEM allocator, warps 4..11
// epilogue.  Mainloop: 128 x BN x 64 k-blocks through a 4-stage smem ring (128B-swizzled K-major tiles), accumulators
// in two TMEM stages so the epilogue of tile i overlaps the MMAs of tile i+1.
//
// Epilogue, per warp and 32-column chunk (thread = accumulator row = TMEM lane):
//   tcgen05.ld 32 columns -> +bias -> activation -> (+ fp32 residual) -> swizzled per-warp smem buffer
//   * identity row map: one TMA store (cp.async.bulk.tensor) of the 32x32 box per chunk, double-buffered;
//   * window-unpartition / pixel-shuffle maps: the buffer is read back row-segment-wise and written with
//     coalesced 128-byte (fp32) / 64-byte (bf16) stores at the remapped rows.
// The epilogue configuration is a template parameter pack so that the hot instantiations carry no per-element
// branches (an earlier all-runtime version was instruction-cache bound: 23 % no_inst stalls in ncu); a value of -1
// keeps that flag a runtime value for the general C-ABI entry point.
#include "gemm_tc.cuh"

#include <stdlib.h>

#include "common.cuh"
#include "tc05.cuh"
#include "act.cuh"

namespace cvb {

constexpr int GEMM_BM = 128;
constexpr int GEMM_BK = 64;
constexpr int GEMM_SMEM_MAX = 232448;

// CG = CTAs cooperating on one accumulator tile (cta_group): 1 = 128 x BN per CTA; 2 = a CTA pair computes 256 x BN,
// each CTA holding its own 128 rows of A and HALF of the B tile, so the operand bytes per CTA and k-block drop from
// 16 + BN/8 KB to 16 + BN/16 KB — more k-blocks in flight for the same shared memory and less L2->SM traffic per flop,
// which is what bounds the K >= 384 shapes (ncu: the MMA thread spins on the full barriers, not on the epilogue).
//
// EW = epilogue warps (8 or 16).  Every configuration runs 8: 16 warps (640 threads; the 16-bit residual-free epilogues
// fit 92 registers) were measured on the small-K shapes and changed nothing (profiles/r1_gemm_epilogue_experiments.md) —
// those epilogues are bound by instruction + MUFU count per element, not by latency hiding.
// EB = bytes of one 32 x 32 staging box (fp32 4096, 16-bit 2048).
template <int BN, int CG>
constexpr int gemm_stage_bytes() { return GEMM_BM * 128 + (BN / CG) * 128; }
template <int BN, int CG, int EW>
constexpr int gemm_epi_bufs() { return (CG == 2 && BN == 256 && EW == 8) ? 1 : 2; }
template <int BN, int CG, int EW, int EB>
constexpr int gemm_stages() {
  int s = (GEMM_SMEM_MAX - 1024 - 256 - EW * gemm_epi_bufs<BN, CG, EW>() * EB) / gemm_stage_bytes<BN, CG>();
  return s > 8 ? 8 : (CG == 1 && s > 4 ? 4 : s);
}
// A-stationary CTA pairs (AST): the A rows of an M tile pair stay in shared memory (AST 64-wide k-blocks, K <= 448) while the
// pair walks every N tile of that row block; only the weights stream through the ring.  The K = 384 / 448 shapes are bound by
// the L2 slice throughput (gemm_trace: the MMA loop waits for operands at 38-42 B/clk/SM = 6300 B/clk chip-wide); this halves
// the L2 bytes per flop (A once per row block instead of once per tile).
// (the template value AST is the number of resident k-blocks: 6 for K <= 384, 7 for K <= 448)
constexpr int AST_MAX = 7;
template <int BN, int EW, int EB, int NKB>
constexpr int gemm_ast_stages() {
  int s = (GEMM_SMEM_MAX - 1024 - 512 - EW * 2 * EB - NKB * GEMM_BM * 128) / ((BN / 2) * 128);
  return s > 8 ? 8 : s;
}
template <int BN, int EW, int EB, int NKB>
constexpr int gemm_ast_smem_bytes() {
  return 1024 + NKB * GEMM_BM * 128 + gemm_ast_stages<BN, EW, EB, NKB>() * (BN / 2) * 128 + EW * 2 * EB + 512;
}

template <int BN, int CG, int EW, int EB>
constexpr int gemm_smem_bytes() {
  return 1024 /*align slack*/ + gemm_stages<BN, CG, EW, EB>() * gemm_stage_bytes<BN, CG>() + EW * gemm_epi_bufs<BN, CG, EW>() * EB + 256;
}

// Row indices fit 32 bits (M is an int and a destination row never exceeds the source row count): the remap runs on
// 32-bit unsigned divisions — the 64-bit ones cost ~1300 instructions per thread and tile on the windowed proj GEMMs.
__device__ __forceinline__ long long gemm_dest_row(const GemmEpilogue& e, int map, long long r64) {
  if (map == GEMM_MAP_UNWINDOW) {
    const unsigned r = (unsigned)r64;
    const unsigned ws = (unsigned)e.ws, w2 = ws * ws;
    const unsigned win = r / w2, t = r - win * w2;
    const unsigned per_img = (unsigned)(e.nwx * e.nwy);
    const unsigned b = win / per_img, wi = win - b * per_img;
    const unsigned wy = wi / (unsigned)e.nwx, wx = wi - wy * (unsigned)e.nwx;
    const unsigned ty = t / ws, tx = t - ty * ws;
    const unsigned y = wy * ws + ty, x = wx * ws + tx;
    if (y >= (unsigned)e.H || x >= (unsigned)e.W) return -1;
    return (long long)((b * (unsigned)e.H + y) * (unsigned)e.W + x);
  }
  if (map == GEMM_MAP_QPOOL) {
    const unsigned r = (unsigned)r64;
    const unsigned ws = (unsigned)e.ws, w2 = ws * ws;
    const unsigned win = r / w2, t = r - win * w2;
    const unsigned ty = t / ws, tx = t - ty * ws;
    if ((ty | tx) & 1u) return -1;
    return (long long)(win * (w2 >> 2) + (ty >> 1) * (ws >> 1) + (tx >> 1));
  }
  if (map == GEMM_MAP_POOL2) {
    // anchor rows (even ty, even tx) own the pooled output row; every other row returns -1
    const unsigned r = (unsigned)r64;
    const unsigned ws = (unsigned)e.ws, w2 = ws * ws;
    const unsigned win = r / w2, t = r - win * w2;
    const unsigned per_img = (unsigned)(e.nwx * e.nwy);
    const unsigned b = win / per_img, wi = win - b * per_img;
    const unsigned wy = wi / (unsigned)e.nwx, wx = wi - wy * (unsigned)e.nwx;
    const unsigned ty = t / ws, tx = t - ty * ws;
    if ((ty | tx) & 1u) return -1;
    const unsigned y = wy * ws + ty, x = wx * ws + tx;
    if (y >= (unsigned)e.H || x >= (unsigned)e.W) return -1;
    return (long long)((b * (unsigned)(e.H >> 1) + (y >> 1)) * (unsigned)(e.W >> 1) + (x >> 1));
  }
  return r64;
}

template <int V>
__device__ __forceinline__ int pick(int runtime) { return V < 0 ? runtime : V; }

// Debug timeline (scripts/gemm_trace.py): plain global stores into fixed slots, no atomics (an atomic costs the recording
// warp ~0.5 us and distorts what it measures).
__device__ __forceinline__ void gemm_trace(const GemmProblem& p, unsigned ev, int idx) {
  if (p.trace && blockIdx.x == 0 && idx < 256) {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    p.trace[1 + ev * 256 + idx] = ((unsigned long long)(ev * 4096u + (unsigned)idx) << 44) | (t & 0xFFFFFFFFFFFull);
  }
}

__device__ __forceinline__ float apply_act(int act, float x) {
  if (act == GEMM_ACT_GELU) return gelu_fast(x);
  if (act == GEMM_ACT_RELU) return fmaxf(x, 0.f);
  return x;
}
constexpr int GEMM_ACT_GELU_TANH = 4;  // internal: tanh-form GELU for 16-bit outputs, one MUFU per element (act.cuh)
constexpr int GEMM_ACT_GELU_SIG = 3;  // internal: sigmoid-form GELU for 16-bit outputs (act.cuh)
__device__ __forceinline__ void apply_act2(int act, float& a, float& b) {
  if (act == GEMM_ACT_GELU_TANH) {
    gelu_tanh2(a, b);
  } else if (act == GEMM_ACT_GELU_SIG) {
    gelu_sig2(a, b);
  } else if (act == GEMM_ACT_GELU) {
    gelu_fast2(a, b);
  } else if (act == GEMM_ACT_RELU) {
    a = fmaxf(a, 0.f);
    b = fmaxf(b, 0.f);
  }
}

// ACT / RES / OUT / MAP / RBA: compile-time epilogue configuration, -1 = read from GemmEpilogue at run time.
// OUT: 0 = fp32, 1 = bf16, 2 = both (runtime-only).
template <int BN, int ACT, int RES, int OUT, int MAP, int RBA, int CG, int EW, int AST = 0>
__global__ void __launch_bounds__(128 + 32 * EW, 1)
k_gemm_tc(const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_w,
          const __grid_constant__ CUtensorMap tmap_out, GemmProblem p, GemmEpilogue e) {
  static_assert(BN % 32 == 0 && BN >= 32 && BN <= 256, "BN");
  static_assert(CG == 1 || CG == 2, "CG");
  static_assert(EW == 8 || EW == 16, "EW");
  static_assert(AST == 0 || CG == 2, "A-stationary tiles are a CTA-pair configuration");
  constexpr int GEMM_EPI_BUF = (OUT == 1 && (MAP == GEMM_MAP_IDENTITY || MAP == GEMM_MAP_QPOOL)) ? 2048 : 4096;  // 16-bit TMA box or fp32 staging
  constexpr int GEMM_STAGES = AST ? gemm_ast_stages<BN, EW, GEMM_EPI_BUF, AST ? AST : 1>() : gemm_stages<BN, CG, EW, GEMM_EPI_BUF>();
  constexpr int EPI_BUFS = AST ? 2 : gemm_epi_bufs<BN, CG, EW>();
  static_assert(GEMM_STAGES >= 2, "smem ring");
  constexpr int A_BYTES = GEMM_BM * 128, B_BYTES = (BN / CG) * 128;
  constexpr int A_SLOTS = AST ? AST : GEMM_STAGES;  // A: resident k-blocks of the row block, or a ring like B
  constexpr int TMEM_COLS = (2 * BN <= 64) ? 64 : (2 * BN <= 128) ? 128 : (2 * BN <= 256) ? 256 : 512;
  constexpr bool TMA_OUT = (MAP == GEMM_MAP_IDENTITY && (OUT == 0 || OUT == 1)) || (MAP == GEMM_MAP_QPOOL && OUT == 1);
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  uint8_t* sA = smem;
  uint8_t* sB = smem + A_SLOTS * A_BYTES;
  uint8_t* epi = sB + GEMM_STAGES * B_BYTES;  // [EW warps][EPI_BUFS][GEMM_EPI_BUF], 1024-byte aligned
  uint64_t* bars = (uint64_t*)(epi + EW * EPI_BUFS * GEMM_EPI_BUF);
  uint64_t* full = bars;                     // [STAGES]
  uint64_t* empty = bars + GEMM_STAGES;      // [STAGES]
  uint64_t* tfull = bars + 2 * GEMM_STAGES;  // [2]
  uint64_t* tempty = tfull + 2;              // [2]
  uint64_t* a_full = tempty + 2;             // [AST_MAX]  (AST only)
  uint64_t* a_empty = a_full + AST_MAX;      // [AST_MAX]
  uint32_t* tmem_slot = (uint32_t*)(a_empty + AST_MAX);

  const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0), lane = threadIdx.x & 31;  // shuffle: warp-uniform for the compiler
  // CG == 2: a tile is 256 rows; CTA `rank` of the pair owns rows [tile_m*256 + rank*128, +128)
  const int rank = CG == 2 ? (int)tc::cluster_ctarank() : 0;
  // AST: the loop index runs over (row block of this pair, N tile) with the N tile fastest; otherwise tiles are dealt round-robin
  const int n_mt = (p.n_tiles_m + CG - 1) / CG;
  const int my_mt = AST ? (n_mt - (int)(blockIdx.x / CG) + (int)(gridDim.x / CG) - 1) / (int)(gridDim.x / CG) : 0;
  const int n_tiles = AST ? my_mt * p.n_tiles_n : n_mt * p.n_tiles_n;
  const int tile0 = AST ? 0 : blockIdx.x / CG, tile_stride = AST ? 1 : gridDim.x / CG;
  // tile index -> (row block, N tile)
  auto tile_mn = [&](int t, int& tm, int& nb) {
    if (AST) {
      const int i = t / p.n_tiles_n;
      nb = t - i * p.n_tiles_n;
      tm = (int)(blockIdx.x / CG) + i * (int)(gridDim.x / CG);
    } else {
      tm = t / p.n_tiles_n;
      nb = t - tm * p.n_tiles_n;
    }
  };
  const int n_kb = (p.K + GEMM_BK - 1) / GEMM_BK;

  if (warp == 0 && lane == 0) {
    tc::prefetch_tmap(&tmap_a);
    tc::prefetch_tmap(&tmap_w);
    if (TMA_OUT) tc::prefetch_tmap(&tmap_out);
  }
  if (warp == 1 && lane == 0) {
    for (int i = 0; i < GEMM_STAGES; i++) {
      tc::mbar_init(&full[i], 1);
      tc::mbar_init(&empty[i], 1);
    }
    for (int i = 0; i < 2; i++) {
      tc::mbar_init(&tfull[i], 1);
      tc::mbar_init(&tempty[i], EW * CG);  // the leader collects the epilogue warps of both CTAs
    }
    if (AST)
      for (int i = 0; i < AST_MAX; i++) {
        tc::mbar_init(&a_full[i], 1);
        tc::mbar_init(&a_empty[i], 1);
      }
    tc::fence_barrier_init();
  }
  if (warp == 2) {
    if (CG == 2) tc::tmem_alloc2<TMEM_COLS>(tmem_slot);
    else tc::tmem_alloc<TMEM_COLS>(tmem_slot);
  }
  tc::tc_fence_before();
  __syncthreads();
  if (CG == 2) tc::cluster_sync();  // the peer's barriers are initialised before anything arrives on them
  tc::tc_fence_after();
  const uint32_t tmem_base = __shfl_sync(0xffffffffu, *tmem_slot, 0);

  if (warp == 0) {
    // ===================== TMA producer
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      int lt = 0;
      for (int t = tile0; t < n_tiles; t += tile_stride, lt++) {
        int tm, nb;
        tile_mn(t, tm, nb);
        const int mb = tm * CG + rank;
        for (int kb = 0; kb < n_kb; kb++) {
          if (AST) {
            if (nb == 0) {
              // k-block kb of the next row block: its slot is free once the LAST N tile of the previous row block has consumed it
              // (released k-block by k-block, so these loads run under that tile's remaining MMAs)
              const uint32_t rb = (uint32_t)(t / p.n_tiles_n);
              tc::mbar_wait(&a_empty[kb], (rb & 1) ^ 1);
              if (rank == 0) tc::mbar_arrive_expect_tx(&a_full[kb], 2 * A_BYTES);
              tc::tma_load_2d_2sm(sA + kb * A_BYTES, &tmap_a, &a_full[kb], kb * GEMM_BK, mb * GEMM_BM);
            }
            tc::mbar_wait(&empty[stage], phase ^ 1);
            if (kb == 0) gemm_trace(p, 5, lt);
            if (rank == 0) tc::mbar_arrive_expect_tx(&full[stage], 2 * B_BYTES);
            tc::tma_load_2d_2sm(sB + stage * B_BYTES, &tmap_w, &full[stage], kb * GEMM_BK, nb * BN + rank * (BN / 2));
            if (++stage == GEMM_STAGES) { stage = 0; phase ^= 1; }
            continue;
          }
          tc::mbar_wait(&empty[stage], phase ^ 1);
          if (kb == 0) gemm_trace(p, 5, lt);
          if (CG == 2) {
            // both CTAs load their half; the bytes of both are counted on the leader's full barrier
            if (rank == 0) tc::mbar_arrive_expect_tx(&full[stage], 2 * (A_BYTES + B_BYTES));
            tc::tma_load_2d_2sm(sA + stage * A_BYTES, &tmap_a, &full[stage], kb * GEMM_BK, mb * GEMM_BM);
            tc::tma_load_2d_2sm(sB + stage * B_BYTES, &tmap_w, &full[stage], kb * GEMM_BK, nb * BN + rank * (BN / 2));
          } else {
            tc::mbar_arrive_expect_tx(&full[stage], A_BYTES + B_BYTES);
            tc::tma_load_2d(sA + stage * A_BYTES, &tmap_a, &full[stage], kb * GEMM_BK, mb * GEMM_BM);
            tc::tma_load_2d(sB + stage * B_BYTES, &tmap_w, &full[stage], kb * GEMM_BK, nb * BN);
          }
          if (++stage == GEMM_STAGES) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer.  The whole warp runs the control flow (converged, so descriptors and barrier addresses
    // stay in uniform registers) and one elected lane issues.  Inside an `if (lane == 0)` region ptxas wrapped every
    // tcgen05.mma in an ELECT / R2UR.BROADCAST / BRA.U.ANY loop: 28 dependent instructions per MMA, as long as the 128
    // tensor-pipe cycles of a 256-wide MMA itself — the K >= 384 shapes sat at 0.55-0.60 of the sustained peak behind it.
    if (rank == 0) {
      const uint32_t idesc = tc::idesc_bf16(GEMM_BM * CG, BN, false, false, e.fp16 != 0);
      int stage = 0;
      uint32_t phase = 0;
      int acc = 0;
      uint32_t acc_phase = 0;
      int lt = 0;
      for (int t = tile0; t < n_tiles; t += tile_stride, lt++) {
        tc::mbar_wait(&tempty[acc], acc_phase ^ 1);
        tc::tc_fence_after();
        if (lane == 0) gemm_trace(p, 0, lt);
        const uint32_t d_tmem = tmem_base + acc * BN;
        const int rb = AST ? t / p.n_tiles_n : 0, nb_ast = AST ? t - rb * p.n_tiles_n : 0;
        for (int kb = 0; kb < n_kb; kb++) {
          if (AST && nb_ast == 0) tc::mbar_wait(&a_full[kb], rb & 1);
          tc::mbar_wait(&full[stage], phase);
          tc::tc_fence_after();
          if (kb == 0 && lane == 0) gemm_trace(p, 1, lt);
          const uint64_t da = tc::desc_kmajor(tc::smem_u32(sA + (AST ? kb : stage) * A_BYTES));
          const uint64_t db = tc::desc_kmajor(tc::smem_u32(sB + stage * B_BYTES));
          const int ksteps = min(GEMM_BK / 16, (p.K - kb * GEMM_BK + 15) / 16);
          if (tc::elect_one()) {
#pragma unroll
            for (int k = 0; k < GEMM_BK / 16; k++)
              if (k < ksteps) {  // + 32 bytes per k-step: 2 units of the descriptor's 16-byte address field
                if (CG == 2) tc::mma_f16_ss2(d_tmem, da + 2 * k, db + 2 * k, idesc, (kb > 0 || k > 0) ? 1u : 0u);
                else tc::mma_f16_ss(d_tmem, da + 2 * k, db + 2 * k, idesc, (kb > 0 || k > 0) ? 1u : 0u);
              }
            if (CG == 2) tc::mma_commit2(&empty[stage]);
            else tc::mma_commit(&empty[stage]);
            if (AST && nb_ast == p.n_tiles_n - 1) tc::mma_commit2(&a_empty[kb]);  // last N tile of the row block: A slot kb is free
          }
          __syncwarp();
          if (++stage == GEMM_STAGES) { stage = 0; phase ^= 1; }
        }
        if (tc::elect_one()) {
          if (CG == 2) tc::mma_commit2(&tfull[acc]);
          else tc::mma_commit(&tfull[acc]);
        }
        __syncwarp();
        if (lane == 0) gemm_trace(p, 2, lt);
        if (++acc == 2) { acc = 0; acc_phase ^= 1; }
      }
    }
  } else if (warp >= 4) {
    // ===================== epilogue: EW warps; warp handles TMEM lanes 32*(warp%4).. and column chunks c == ew/4 (mod EW/4)
    const int act = pick<ACT>(e.act);
    const bool has_res = pick<RES>(e.res != nullptr) != 0;
    const int map = pick<MAP>(e.map_mode);
    const bool rba = pick<RBA>(e.res_before_act) != 0;
    const bool out_f32 = (OUT < 0) ? (e.out_f32 != nullptr) : (OUT == 0 || OUT == 2);
    const bool out_b16 = (OUT < 0) ? (e.out_bf16 != nullptr) : (OUT == 1 || OUT == 2);
    const int ew = warp - 4;
    const int quad = warp & 3;  // TMEM lane quadrant this warp may access (= warp id mod 4)
    const int cgroup = ew >> 2;
    uint8_t* bufs = epi + ew * EPI_BUFS * GEMM_EPI_BUF;
    int acc = 0;
    uint32_t acc_phase = 0;
    uint32_t nbuf = 0;
    const int sub_row = lane >> 3, sub_c = lane & 7;
    int cg_rot = cgroup;
    int lt = -1;
    for (int t = tile0; t < n_tiles; t += tile_stride) {
      lt++;
      int tm, nb;
      tile_mn(t, tm, nb);
      const int mb = tm * CG + rank;
      const long long row0 = (long long)mb * GEMM_BM + quad * 32;
      const long long myrow = row0 + lane;
      // remapped destination rows: every lane maps its own row once, the row-segment owners fetch it by shuffle
      // (a destination row index never exceeds the source row count, so it fits an int)
      int dest[8];
      if (MAP == GEMM_MAP_QPOOL) dest[0] = (myrow < p.M) ? (int)gemm_dest_row(e, map, myrow) : -1;
      if (!TMA_OUT) {
        int dmine = (myrow < p.M) ? (int)gemm_dest_row(e, map, myrow) : -1;
        if (map == GEMM_MAP_SHUFFLE2 && dmine >= 0) {
          // (b, y, x) on the H x W grid -> row of pixel (2y, 2x) on the 2H x 2W grid; the chunk's (dy, dx) is added below
          const unsigned hw = (unsigned)(e.H * e.W);
          const unsigned img = (unsigned)dmine / hw, rem = (unsigned)dmine - img * hw;
          const unsigned y = rem / (unsigned)e.W, xx = rem - y * (unsigned)e.W;
          dmine = (int)((img * 2u * (unsigned)e.H + 2u * y) * 2u * (unsigned)e.W + 2u * xx);
        }
        if (MAP == GEMM_MAP_POOL2) {
          dest[0] = dmine;  // the anchor lane writes its pooled row itself
        } else {
#pragma unroll
          for (int it = 0; it < 8; it++) dest[it] = __shfl_sync(0xffffffffu, dmine, it * 4 + sub_row);
        }
      }
      tc::mbar_wait(&tfull[acc], acc_phase);
      tc::tc_fence_after();
      if (ew == 0 && lane == 0) gemm_trace(p, 3, lt);
#pragma unroll 1
      // BN / 32 chunks over EW / 4 warp groups: with 3 chunks and 2 groups one group would take two chunks of EVERY tile;
      // rotating the start group evens that out over consecutive tiles (the accumulators are double-buffered, so a group
      // that finishes early moves on to the next tile)
      const int c_first = cg_rot;
      if ((BN / 32) % (EW / 4) != 0 && p.rotate) cg_rot = (cg_rot + (EW / 4) - ((BN / 32) % (EW / 4))) % (EW / 4);
      // In-place residual stream (res == out): ask for the residual boxes this warp will add in its NEXT tile to be brought
      // into L2 now — one bulk-prefetch instruction per 32 x 32 box through the output's tensor map, no registers, no
      // shared memory.  The row-per-thread residual loads of the K <= 384 shapes then wait for an L2 hit instead of HBM
      // (ncu: a third of the epilogue warps' stall samples sat on the first use of those loads).  Every residual element
      // is read by exactly one tile, so nothing is prefetched twice.
      if (TMA_OUT && OUT == 0 && RES != 0 && p.res_l2pf && lane == 0) {
        const int t2 = t + tile_stride;
        if (t2 < n_tiles) {
          int tm2, nb2;
          tile_mn(t2, tm2, nb2);
          const long long r2 = (long long)(tm2 * CG + rank) * GEMM_BM + quad * 32;
          if (r2 < p.M)
            for (int c2 = cg_rot; c2 < BN / 32; c2 += EW / 4) {
              const int col2 = nb2 * BN + c2 * 32;
              if (col2 < p.N) tc::tma_prefetch_2d(&tmap_out, col2, (int)r2);
            }
        }
      }
      // window un-partition epilogues (rows scattered by the map, plain stores): every lane prefetches the 128-byte
      // pieces of its own destination row of the next tile
      if (!TMA_OUT && MAP != GEMM_MAP_POOL2 && p.res_l2pf && map == GEMM_MAP_UNWINDOW && has_res) {
        const int t2 = t + tile_stride;
        if (t2 < n_tiles) {
          int tm2, nb2;
          tile_mn(t2, tm2, nb2);
          const long long m2 = (long long)(tm2 * CG + rank) * GEMM_BM + quad * 32 + lane;
          const long long d2 = m2 < p.M ? gemm_dest_row(e, map, m2) : -1;
          if (d2 >= 0) {
            const float* rp2 = e.res + d2 * e.ld_res + nb2 * BN;
            for (int c2 = cg_rot; c2 < BN / 32; c2 += EW / 4)
              if (nb2 * BN + c2 * 32 < p.N) asm volatile("prefetch.global.L2 [%0];" ::"l"(rp2 + c2 * 32));
          }
        }
      }
      for (int c = c_first; c < BN / 32; c += EW / 4) {
        const int col0 = nb * BN + c * 32;
        const int ncols = p.N - col0;  // valid columns of this chunk: >= 32, 16 (N % 32 == 16) or <= 0
        uint8_t* buf = bufs + (nbuf % EPI_BUFS) * GEMM_EPI_BUF;
        // The staging buffers alternate with the TMA-store groups: wait_group.read<EPI_BUFS - 1> below frees the buffer used
        // EPI_BUFS uses ago only if every use in between committed exactly one group.  The pooled q chunks of a Q-pooled qkv GEMM
        // write their result with ordinary stores (no group), so they must not advance the alternation: with nbuf++ here, a
        // k / v chunk could overwrite a box its predecessor's TMA store was still reading (seen as run-to-run differences of
        // one attention window, scripts/sam2_determinism_probe.py).
        if (!(MAP == GEMM_MAP_QPOOL && col0 < e.pool_cols)) nbuf++;
        // residual for the TMA path: this thread's own row, 32 consecutive floats (issued before the TMEM wait)
        float4 rv[8];
        // residual through the staging buffer (fp32 boxes): loaded with lanes ALONG the columns (8 lanes per 128-byte row piece,
        // 4 rows per instruction = 4 L1 wavefronts instead of the 32 of a row-per-thread load), handed to the row-owning thread
        // through the buffer that then carries the result
        const bool res_staged = TMA_OUT && OUT == 0 && RES == 1 && p.res_stage;
        if (res_staged) {
          const int prow = lane >> 3, pcol = lane & 7;
#pragma unroll
          for (int i = 0; i < 8; i++) {
            const long long r = row0 + prow + 4 * i;
            const long long rr = e.res_row_mod > 0 ? r % e.res_row_mod : r;
            rv[i] = (r < p.M && pcol * 4 < ncols) ? *(const float4*)(e.res + rr * e.ld_res + col0 + pcol * 4) : make_float4(0.f, 0.f, 0.f, 0.f);
          }
        } else if (TMA_OUT && has_res) {
          if (myrow < p.M) {
            long long rr = e.res_row_mod > 0 ? myrow % e.res_row_mod : myrow;
            const float4* rp = (const float4*)(e.res + rr * e.ld_res + col0);
#pragma unroll
            for (int j = 0; j < 8; j++) rv[j] = 4 * j < ncols ? rp[j] : make_float4(0.f, 0.f, 0.f, 0.f);
          } else {
#pragma unroll
            for (int j = 0; j < 8; j++) rv[j] = make_float4(0.f, 0.f, 0.f, 0.f);
          }
        }
        uint32_t v[32];
        tc::tmem_ld_32x32(tmem_base + ((uint32_t)(quad * 32) << 16) + acc * BN + c * 32, v);
        // SHUFFLE2: this 32-column chunk lies inside one (dy,dx) group; bias/out column = co
        int q = 0, ocol0 = col0;
        if (map == GEMM_MAP_SHUFFLE2) { q = col0 / e.cout; ocol0 = col0 - q * e.cout; }
        // remapped path: final destination rows and their residual segments, all loads issued up front (the output may
        // alias the residual, so the compiler cannot hoist these loads above the stores of an earlier iteration itself)
        long long dfin[8];
        float4 rres[8];
        if (!TMA_OUT && MAP != GEMM_MAP_POOL2) {
#pragma unroll
          for (int it = 0; it < 8; it++) {
            long long d = dest[it];
            if (d >= 0 && map == GEMM_MAP_SHUFFLE2) d += (long long)(q >> 1) * 2 * e.W + (q & 1);
            dfin[it] = d;
            rres[it] = make_float4(0.f, 0.f, 0.f, 0.f);
            if (has_res && d >= 0 && sub_c * 4 < ncols) {
              long long rr = e.res_row_mod > 0 ? d % e.res_row_mod : d;
              rres[it] = *(const float4*)(e.res + rr * e.ld_res + ocol0 + sub_c * 4);
            }
          }
        }
        float4 bv[8];
        if (e.bias) {
#pragma unroll
          for (int j = 0; j < 8; j++) bv[j] = 4 * j < ncols ? __ldg((const float4*)(e.bias + ocol0) + j) : make_float4(0.f, 0.f, 0.f, 0.f);
        } else {
#pragma unroll
          for (int j = 0; j < 8; j++) bv[j] = make_float4(0.f, 0.f, 0.f, 0.f);
        }
        if (TMA_OUT) {
          if (lane == 0) tc::tma_store_wait_read<EPI_BUFS - 1>();  // the store that last read `buf` has drained it
          __syncwarp();
        }
        if (res_staged) {
          const int prow = lane >> 3, pcol = lane & 7;
#pragma unroll
          for (int i = 0; i < 8; i++) {
            const int rr = prow + 4 * i;
            *(float4*)(buf + rr * 128 + ((pcol ^ (rr & 7)) << 4)) = rv[i];
          }
          __syncwarp();
#pragma unroll
          for (int j = 0; j < 8; j++) rv[j] = *(const float4*)(buf + lane * 128 + ((j ^ (lane & 7)) << 4));
        }
        tc::tmem_ld_wait();
        float f[32];
#pragma unroll
        for (int j = 0; j < 8; j++) {
          // packed fp32x2 adds: the accumulator registers of tcgen05.ld and the bias float4 are natural pairs
          up2(add2(pk2(__uint_as_float(v[4 * j + 0]), __uint_as_float(v[4 * j + 1])), pk2(bv[j].x, bv[j].y)), f[4 * j + 0],
              f[4 * j + 1]);
          up2(add2(pk2(__uint_as_float(v[4 * j + 2]), __uint_as_float(v[4 * j + 3])), pk2(bv[j].z, bv[j].w)), f[4 * j + 2],
              f[4 * j + 3]);
        }
        if (TMA_OUT) {
          if (has_res && rba) {
#pragma unroll
            for (int j = 0; j < 8; j++) { f[4 * j] += rv[j].x; f[4 * j + 1] += rv[j].y; f[4 * j + 2] += rv[j].z; f[4 * j + 3] += rv[j].w; }
          }
          // fp16 output + tanh-form GELU: the activation runs on the packed halves below (one MUFU per two elements)
          const bool act_h2 = ACT == GEMM_ACT_GELU_TANH && OUT == 1 && e.gelu_h2;
          if (!act_h2) {
#pragma unroll
            for (int i = 0; i < 32; i += 2) apply_act2(act, f[i], f[i + 1]);
          }
          if (has_res && !rba) {
#pragma unroll
            for (int j = 0; j < 8; j++) { f[4 * j] += rv[j].x; f[4 * j + 1] += rv[j].y; f[4 * j + 2] += rv[j].z; f[4 * j + 3] += rv[j].w; }
          }
          if (MAP == GEMM_MAP_QPOOL && col0 < e.pool_cols) {
            // q columns of a Q-pooled block: the chunk is staged as 16-bit rows (rounding is monotonic, so the maximum of the
            // rounded values is the rounded maximum), then lane (j, cg) = (lane / 4, lane % 4) reduces pooled token j of
            // this 32-row block — rows (a, a + 1, a + ws, a + ws + 1) of its anchor a — over 8 columns and stores 16 bytes:
            // ~25 instructions per thread instead of the 64 shuffles + 64 maxima of a register-resident 2 x 2 maximum
            const int wsl = e.ws;
#pragma unroll
            for (int j = 0; j < 4; j++) {
              uint32_t w[4];
#pragma unroll
              for (int k = 0; k < 4; k++) w[k] = tc::pack16(e.fp16, f[8 * j + 2 * k], f[8 * j + 2 * k + 1]);
              *(uint4*)(buf + lane * 64 + ((j ^ ((lane >> 1) & 3)) << 4)) = make_uint4(w[0], w[1], w[2], w[3]);
            }
            __syncwarp();
            {
              const int pj = lane >> 2, cgq = lane & 3, half = wsl >> 1;
              const int a = (pj / half) * 2 * wsl + (pj % half) * 2;
              const int dd = __shfl_sync(0xffffffffu, dest[0], a);
              uint4 m = *(const uint4*)(buf + a * 64 + ((cgq ^ ((a >> 1) & 3)) << 4));
#pragma unroll
              for (int s2 = 1; s2 < 4; s2++) {
                const int r = a + (s2 & 1) + (s2 >> 1) * wsl;
                const uint4 o = *(const uint4*)(buf + r * 64 + ((cgq ^ ((r >> 1) & 3)) << 4));
                m.x = tc::max16x2(e.fp16, m.x, o.x); m.y = tc::max16x2(e.fp16, m.y, o.y);
                m.z = tc::max16x2(e.fp16, m.z, o.z); m.w = tc::max16x2(e.fp16, m.w, o.w);
              }
              if (dd >= 0) *(uint4*)(e.pool_out + (long long)dd * e.ld_pool + col0 + cgq * 8) = m;
            }
            __syncwarp();
            continue;
          }
          if (OUT == 0) {
            // fp32 box: 128-byte rows, 128B swizzle (16-byte chunk j of row r lives at chunk j ^ (r & 7))
#pragma unroll
            for (int j = 0; j < 8; j++)
              *(float4*)(buf + lane * 128 + ((j ^ (lane & 7)) << 4)) = make_float4(f[4 * j], f[4 * j + 1], f[4 * j + 2], f[4 * j + 3]);
          } else {
            // bf16 box: 64-byte rows, 64B swizzle (chunk j of row r lives at chunk j ^ ((r >> 1) & 3))
#pragma unroll
            for (int j = 0; j < 4; j++) {
              uint32_t w[4];
              // one uniform branch per 8 values instead of a predicated pair of conversions per value pair
              if (act_h2) {
#pragma unroll
                for (int k = 0; k < 4; k++) w[k] = gelu_tanh_h2(tc::pack16(1, f[8 * j + 2 * k], f[8 * j + 2 * k + 1]));
              } else if (e.fp16) {
#pragma unroll
                for (int k = 0; k < 4; k++) w[k] = tc::pack16(1, f[8 * j + 2 * k], f[8 * j + 2 * k + 1]);
              } else {
#pragma unroll
                for (int k = 0; k < 4; k++) w[k] = tc::pack16(0, f[8 * j + 2 * k], f[8 * j + 2 * k + 1]);
              }
              *(uint4*)(buf + lane * 64 + ((j ^ ((lane >> 1) & 3)) << 4)) = make_uint4(w[0], w[1], w[2], w[3]);
            }
          }
          tc::fence_proxy_async_smem();
          __syncwarp();
          if (lane == 0 && col0 < p.N && row0 < p.M) {
            tc::tma_store_2d(&tmap_out, buf, col0, (int)row0);
            tc::tma_store_commit();
          }
        } else if (MAP == GEMM_MAP_POOL2) {
          // 2 x 2 maximum over rows (a, a + 1, a + ws, a + ws + 1) of each anchor row a: the chunk goes through the swizzled fp32
          // staging buffer and lane (j, cg) = (lane / 4, lane % 4) reduces pooled token j over 8 columns (two float4 per
          // row) and stores 32 contiguous bytes; 4 lanes cover a 128-byte line of the pooled row
          const int wsl = e.ws;
#pragma unroll
          for (int j = 0; j < 8; j++)
            *(float4*)(buf + lane * 128 + ((j ^ (lane & 7)) << 4)) = make_float4(f[4 * j], f[4 * j + 1], f[4 * j + 2], f[4 * j + 3]);
          __syncwarp();
          {
            const int pj = lane >> 2, cgq = lane & 3, half = wsl >> 1;
            const int a = (pj / half) * 2 * wsl + (pj % half) * 2;
            const int dd = __shfl_sync(0xffffffffu, dest[0], a);
            float4 m0 = *(const float4*)(buf + a * 128 + (((2 * cgq) ^ (a & 7)) << 4));
            float4 m1 = *(const float4*)(buf + a * 128 + (((2 * cgq + 1) ^ (a & 7)) << 4));
#pragma unroll
            for (int s2 = 1; s2 < 4; s2++) {
              const int r = a + (s2 & 1) + (s2 >> 1) * wsl;
              const float4 o0 = *(const float4*)(buf + r * 128 + (((2 * cgq) ^ (r & 7)) << 4));
              const float4 o1 = *(const float4*)(buf + r * 128 + (((2 * cgq + 1) ^ (r & 7)) << 4));
              m0.x = fmaxf(m0.x, o0.x); m0.y = fmaxf(m0.y, o0.y); m0.z = fmaxf(m0.z, o0.z); m0.w = fmaxf(m0.w, o0.w);
              m1.x = fmaxf(m1.x, o1.x); m1.y = fmaxf(m1.y, o1.y); m1.z = fmaxf(m1.z, o1.z); m1.w = fmaxf(m1.w, o1.w);
            }
            if (dd >= 0 && col0 + cgq * 8 < p.N) {
              float4* o = (float4*)(e.out_f32 + (long long)dd * e.ld_f32 + col0 + cgq * 8);
              o[0] = m0;
              if (col0 + cgq * 8 + 4 < p.N) o[1] = m1;
            }
          }
          __syncwarp();
        } else {
          if (!rba) {
#pragma unroll
            for (int i = 0; i < 32; i += 2) apply_act2(act, f[i], f[i + 1]);
          }
#pragma unroll
          for (int j = 0; j < 8; j++)
            *(float4*)(buf + lane * 128 + ((j ^ (lane & 7)) << 4)) = make_float4(f[4 * j], f[4 * j + 1], f[4 * j + 2], f[4 * j + 3]);
          __syncwarp();
          if (col0 < p.N) {
#pragma unroll
            for (int it = 0; it < 8; it++) {
              const long long d = dfin[it];
              if (d < 0 || sub_c * 4 >= ncols) continue;
              const int r = it * 4 + sub_row;
              float4 x = *(const float4*)(buf + r * 128 + ((sub_c ^ (r & 7)) << 4));
              const int oc = ocol0 + sub_c * 4;
              if (has_res) { x.x += rres[it].x; x.y += rres[it].y; x.z += rres[it].z; x.w += rres[it].w; }
              if (rba) { apply_act2(act, x.x, x.y); apply_act2(act, x.z, x.w); }
              if (out_f32) *(float4*)(e.out_f32 + d * e.ld_f32 + oc) = x;
              if (out_b16) {
                uint2 pk;
                pk.x = tc::pack16(e.fp16, x.x, x.y);
                pk.y = tc::pack16(e.fp16, x.z, x.w);
                *(uint2*)(e.out_bf16 + d * e.ld_bf16 + oc) = pk;
              }
            }
          }
          __syncwarp();
        }
      }
      // all TMEM reads of this accumulator stage are complete (tmem_ld_wait above): hand it back to the MMA warp
      tc::tc_fence_before();
      __syncwarp();
      if (lane == 0) {
        if (CG == 2) tc::mbar_arrive_leader_relaxed(&tempty[acc]);  // (tmem_ld_wait + tcgen05 fence above: the reads are complete)
        else tc::mbar_arrive(&tempty[acc]);
        if (ew == 0) gemm_trace(p, 4, lt);
      }
      if (++acc == 2) { acc = 0; acc_phase ^= 1; }
    }
    if (TMA_OUT && lane == 0) tc::tma_store_wait<0>();  // global writes complete before the CTA retires
  }
  tc::tc_fence_before();
  __syncthreads();
  if (CG == 2) tc::cluster_sync();  // no CTA of the pair retires while the other may still signal it
  if (warp == 2) {
    tc::tc_fence_after();
    if (CG == 2) tc::tmem_dealloc2<TMEM_COLS>(tmem_base);
    else tc::tmem_dealloc<TMEM_COLS>(tmem_base);
  }
}

static unsigned long long* g_gemm_trace = nullptr;

template <int BN, int ACT, int RES, int OUT, int MAP, int RBA, int CG = 1, int EW = 8, int AST = 0>
static int launch_cfg(const __nv_bfloat16* A, long long lda, const __nv_bfloat16* W, long long ldw, int M, int N, int K,
                      const GemmEpilogue& epi, int num_sms, cudaStream_t st) {
  static std::atomic<unsigned long long> attr_set{0};
  constexpr int GEMM_THREADS = 128 + 32 * EW;
  constexpr int EB = (OUT == 1 && (MAP == GEMM_MAP_IDENTITY || MAP == GEMM_MAP_QPOOL)) ? 2048 : 4096;
  constexpr int smem = AST ? gemm_ast_smem_bytes<BN, EW, EB, AST ? AST : 1>() : gemm_smem_bytes<BN, CG, EW, EB>();
  static_assert(smem <= GEMM_SMEM_MAX, "shared memory budget");
  static_assert(!AST || gemm_ast_stages<BN, EW, EB, AST ? AST : 1>() >= 2, "A-stationary ring");
  if (AST && K > AST * GEMM_BK) return cvb_fail(CV_ERR_INVALID, "gemm: row block does not fit the resident A slots");
  auto kern = k_gemm_tc<BN, ACT, RES, OUT, MAP, RBA, CG, EW, AST>;
  if (cvb_once_per_device(attr_set)) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    if (e != cudaSuccess) return cvb_fail_cuda(e, "cudaFuncSetAttribute(k_gemm_tc)");
  }
  CUtensorMap ta, tw, to;
  if (!tc_host::make_tmap_bf16(&ta, A, (uint64_t)M, (uint64_t)K, (uint64_t)lda, GEMM_BM) ||
      !tc_host::make_tmap_bf16(&tw, W, (uint64_t)N, (uint64_t)K, (uint64_t)ldw, BN / CG))
    return cvb_fail(CV_ERR_CUDA, "cuTensorMapEncodeTiled failed (GEMM operands)");
  constexpr bool TMA_OUT = (MAP == GEMM_MAP_IDENTITY && (OUT == 0 || OUT == 1)) || (MAP == GEMM_MAP_QPOOL && OUT == 1);
  if (TMA_OUT) {
    bool ok = OUT == 0 ? tc_host::make_tmap_2d(&to, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, epi.out_f32, (uint64_t)M, (uint64_t)N,
                                               (uint64_t)epi.ld_f32 * 4, 32, 32, CU_TENSOR_MAP_SWIZZLE_128B)
                       : tc_host::make_tmap_2d(&to, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, epi.out_bf16, (uint64_t)M, (uint64_t)N,
                                               (uint64_t)epi.ld_bf16 * 2, 32, 32, CU_TENSOR_MAP_SWIZZLE_64B);
    if (!ok) return cvb_fail(CV_ERR_CUDA, "cuTensorMapEncodeTiled failed (GEMM output)");
  } else {
    to = ta;
  }
  GemmProblem p;
  p.M = M; p.N = N; p.K = K;
  p.n_tiles_m = (M + GEMM_BM - 1) / GEMM_BM;
  p.n_tiles_n = (N + BN - 1) / BN;
  static const int rotate_on = getenv("CVB_GEMM_ROT") ? atoi(getenv("CVB_GEMM_ROT")) : 1;
  p.rotate = rotate_on;
  p.trace = g_gemm_trace;
  static const int res_stage_on = getenv("CVB_GEMM_RES_STAGE") ? atoi(getenv("CVB_GEMM_RES_STAGE")) : 1;
  // same-box A/B (10 reps): N384 K1536 843 -> 982 TFLOP/s, N768 K3072 1178 -> 1292, N192 K192 / N384 K384 +9 %; the 96-wide
  // shapes, already at 5.5 TB/s, lose 3 % to the extra shared-memory round trip and keep the direct loads
  p.res_stage = res_stage_on && N >= 128;
  static const int res_l2pf_on = getenv("CVB_GEMM_RES_L2PF") ? atoi(getenv("CVB_GEMM_RES_L2PF")) : 1;
  // same-box A/B (scripts/gpu_ab.sh CVB_GEMM_RES_L2PF, two alternating runs): N192 K192 +res 0.911 -> 0.820 ms (2 launches), N384 K384
  // +res 0.884 -> 0.772 (4), its windowed form 0.762 -> 0.719 (3), N96 K96 +res 0.912 -> 0.867; the long-K shapes, whose MMA loop is
  // bound by operand delivery, LOSE to the extra requests (N384 K1536 2.40 -> 2.47 ms over 7 launches, N768 K3072 +5 %): K <= 512 only
  p.res_l2pf = res_l2pf_on && K <= 512 && epi.res != nullptr && epi.res_row_mod == 0 &&
               (TMA_OUT ? (OUT == 0 && epi.res == epi.out_f32 && epi.ld_res == epi.ld_f32) : true);
  long long tiles = (long long)((p.n_tiles_m + CG - 1) / CG) * (AST ? 1 : p.n_tiles_n);  // AST: work unit = a row block
  const int max_groups = num_sms / CG;
  int grid = (int)(tiles < max_groups ? tiles : max_groups) * CG;
  cvb_next_work(2.0 * (double)M * (double)N * (double)K);
  if (cvb_profile_on()) {
    char nm[96];
    snprintf(nm, sizeof(nm), "gemm M%d N%d K%d bn%d%s%s%s%s%s%s", M, N, K, BN, epi.out_bf16 ? " ->bf16" : "", epi.out_f32 ? " ->f32" : "",
             epi.res ? " +res" : "", ACT < 0 ? " generic" : "", CG == 2 ? (AST ? " 2cta-ast" : " 2cta") : "", EW == 16 ? " ew16" : "");
    cvb_next_name(nm);
  }
  if (CG == 2) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(grid);
    cfg.blockDim = dim3(GEMM_THREADS);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = 2;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    bool prof = cvb_profile_on();
    if (prof) cvb_profile_begin("k_gemm_tc", st, cvb_take_work());
    cudaError_t le = cudaLaunchKernelEx(&cfg, kern, ta, tw, to, p, epi);
    if (prof) cvb_profile_end(st);
    if (le != cudaSuccess) return cvb_fail_cuda(le, "launch k_gemm_tc (CTA pair)");
    cvb_count_launch();
    return CV_OK;
  }
  CVB_LAUNCH(kern, dim3(grid), dim3(GEMM_THREADS), smem, st, ta, tw, to, p, epi);
  return CV_OK;
}

template <int BN>
static int launch_bn(const __nv_bfloat16* A, long long lda, const __nv_bfloat16* W, long long ldw, int M, int N, int K,
                     const GemmEpilogue& e, int num_sms, cudaStream_t st) {
#define CVB_GEMM_ARGS A, lda, W, ldw, M, N, K, e, num_sms, st
  const bool res = e.res != nullptr, f32 = e.out_f32 != nullptr, b16 = e.out_bf16 != nullptr;
  // TMA-store outputs need a 16-byte aligned base and pitch; everything else takes the coalesced-store path
  const bool tma_ok = f32 ? (((uintptr_t)e.out_f32 & 15) == 0 && (e.ld_f32 % 4) == 0)
                          : (((uintptr_t)e.out_bf16 & 15) == 0 && (e.ld_bf16 % 8) == 0);
  if (e.map_mode == GEMM_MAP_IDENTITY && tma_ok && !(f32 && b16)) {
    // CTA pairs for the K >= 256 shapes that are bound by operand delivery, not by HBM
    static const bool pairs_on = getenv("CVB_GEMM_PAIRS") ? atoi(getenv("CVB_GEMM_PAIRS")) != 0 : true;
    static const int pair_min_bn = getenv("CVB_PAIR_MINBN") ? atoi(getenv("CVB_PAIR_MINBN")) : 256;
    // GELU form of the 16-bit-output epilogues (act.cuh): 2 = tanh form (default), 1 = sigmoid form, 0 = A&S erf
    static const int gelu_form = getenv("CVB_GELU_FORM") ? atoi(getenv("CVB_GELU_FORM")) : 2;
    const bool gelu_sig = gelu_form == 1, gelu_tanh = gelu_form == 2;
    // 192- / 224-wide tiles: only the long-K residual GEMMs (fc2 of stages 3-4) gain from the pair (their operand stream is what
    // binds them: +7 % at N384 K1536, +10 % at N448 K1792); the K = 384 qkv shapes lose 7 % to the pair's extra synchronisation
    static const int pair_fc2_on = getenv("CVB_PAIR_FC2") ? atoi(getenv("CVB_PAIR_FC2")) : 1;
    static const int pair_qkv_on = getenv("CVB_PAIR_QKV") ? atoi(getenv("CVB_PAIR_QKV")) : 1;
    const bool pair_fc2 = pair_fc2_on && BN >= 192 && f32 && res && K >= 1024;
    // ... and, once the accumulator hand-back of a pair no longer paid a cluster-scope release fence, the 16-bit K = 384 / 448 qkv
    // shapes as well (N1152 K384: 920 -> 1080 TFLOP/s)
    const bool pair_qkv = pair_qkv_on && BN >= 192 && b16 && !res;
    if (pairs_on && BN >= 128 && (BN >= pair_min_bn || pair_fc2 || pair_qkv) && K >= 256 && M >= 1024) {
      // A-stationary row blocks for the L2-bound K <= 448 shapes with several N tiles (16-bit outputs: qkv, fc1)
      static const int ast_on = getenv("CVB_GEMM_AST") ? atoi(getenv("CVB_GEMM_AST")) : 1;
      if constexpr (BN >= 192) {
        if (ast_on && K <= AST_MAX * GEMM_BK && (N + BN - 1) / BN >= 3 && b16 && !res) {
          if (K <= 6 * GEMM_BK) {
            if (e.act == GEMM_ACT_NONE) return launch_cfg<BN, 0, 0, 1, 0, 0, 2, 8, 6>(CVB_GEMM_ARGS);
          } else {
            if (e.act == GEMM_ACT_NONE) return launch_cfg<BN, 0, 0, 1, 0, 0, 2, 8, 7>(CVB_GEMM_ARGS);
          }
          // (the GELU shapes with their 16 epilogue warps keep the ring: 64 KB of staging leave the weight ring 3 stages beside a
          //  resident A, and N1792 K448 fell from 1100 to 990 TFLOP/s)
        }
      }
      if (b16 && !res && e.act == GEMM_ACT_NONE) return launch_cfg<BN, 0, 0, 1, 0, 0, 2>(CVB_GEMM_ARGS);
      // 16 epilogue warps for the GELU epilogue of the CTA-pair shapes (fc1 of stages 3-4, epilogue-bound): 844 -> 880 TFLOP/s on
      // M262144 N1536 K384 once the MMA issue no longer co-limited (CVB_GELU_EW16=0 restores 8)
      static const int gelu_ew16 = getenv("CVB_GELU_EW16") ? atoi(getenv("CVB_GELU_EW16")) : 1;
      if (b16 && !res && e.act == GEMM_ACT_GELU && gelu_tanh && gelu_ew16) return launch_cfg<BN, 4, 0, 1, 0, 0, 2, 16>(CVB_GEMM_ARGS);
      if (b16 && !res && e.act == GEMM_ACT_GELU && gelu_tanh) return launch_cfg<BN, 4, 0, 1, 0, 0, 2>(CVB_GEMM_ARGS);
      if (b16 && !res && e.act == GEMM_ACT_GELU && gelu_sig) return launch_cfg<BN, 3, 0, 1, 0, 0, 2>(CVB_GEMM_ARGS);
      if (b16 && !res && e.act == GEMM_ACT_GELU) return launch_cfg<BN, 1, 0, 1, 0, 0, 2>(CVB_GEMM_ARGS);
      if (f32 && res && e.act == GEMM_ACT_NONE) return launch_cfg<BN, 0, 1, 0, 0, 0, 2>(CVB_GEMM_ARGS);
      if (f32 && !res && e.act == GEMM_ACT_NONE) return launch_cfg<BN, 0, 0, 0, 0, 0, 2>(CVB_GEMM_ARGS);
    }
    if (b16 && !res && e.act == GEMM_ACT_NONE) return launch_cfg<BN, 0, 0, 1, 0, 0>(CVB_GEMM_ARGS);  // qkv
    if (b16 && !res && e.act == GEMM_ACT_GELU && gelu_tanh) return launch_cfg<BN, 4, 0, 1, 0, 0>(CVB_GEMM_ARGS);
    if (b16 && !res && e.act == GEMM_ACT_GELU && gelu_sig) return launch_cfg<BN, 3, 0, 1, 0, 0>(CVB_GEMM_ARGS);
    if (b16 && !res && e.act == GEMM_ACT_GELU) return launch_cfg<BN, 1, 0, 1, 0, 0>(CVB_GEMM_ARGS);  // mlp fc1
    if (f32 && res && e.act == GEMM_ACT_NONE) return launch_cfg<BN, 0, 1, 0, 0, 0>(CVB_GEMM_ARGS);   // fc2, global proj, tables
    if (f32 && !res && e.act == GEMM_ACT_NONE) return launch_cfg<BN, 0, 0, 0, 0, 0>(CVB_GEMM_ARGS);  // shortcut, neck
  }
  if (e.map_mode == GEMM_MAP_UNWINDOW && f32 && !b16 && res && e.act == GEMM_ACT_NONE)
    return launch_cfg<BN, 0, 1, 0, 1, 0>(CVB_GEMM_ARGS);                                             // windowed proj
  if (e.map_mode == GEMM_MAP_QPOOL) {
    const bool tma16 = b16 && !f32 && ((uintptr_t)e.out_bf16 & 15) == 0 && (e.ld_bf16 % 8) == 0;
    if (!tma16 || res || e.act != GEMM_ACT_NONE || (e.ws != 4 && e.ws != 8) || !e.pool_out || (e.pool_cols % 32) ||
        ((uintptr_t)e.pool_out & 15) || (e.ld_pool % 8))
      return cvb_fail(CV_ERR_INVALID, "gemm: QPOOL needs a 16-bit TMA output, no residual / activation, windows of 4 or 8");
    return launch_cfg<BN, 0, 0, 1, 4, 0>(CVB_GEMM_ARGS);                                             // qkv of a Q-pooled block
  }
  if (e.map_mode == GEMM_MAP_POOL2) {
    if (!f32 || b16 || res || e.act != GEMM_ACT_NONE || (e.ws != 4 && e.ws != 8) || (e.H % e.ws) || (e.W % e.ws))
      return cvb_fail(CV_ERR_INVALID, "gemm: POOL2 needs fp32 output, no residual / activation, windows of 4 or 8 without padding");
    return launch_cfg<BN, 0, 0, 0, 3, 0>(CVB_GEMM_ARGS);                                             // Q-pool shortcut
  }
  if (e.map_mode == GEMM_MAP_SHUFFLE2 && f32 && !b16 && res) {
    if (e.act == GEMM_ACT_NONE) return launch_cfg<BN, 0, 1, 0, 2, 0>(CVB_GEMM_ARGS);                 // upscale 1
    if (e.act == GEMM_ACT_GELU && e.res_before_act) return launch_cfg<BN, 1, 1, 0, 2, 1>(CVB_GEMM_ARGS);  // upscale 2
  }
  return launch_cfg<BN, -1, -1, -1, -1, -1>(CVB_GEMM_ARGS);                                          // anything else
#undef CVB_GEMM_ARGS
}

int gemm_tc_launch(const __nv_bfloat16* A, long long lda, const __nv_bfloat16* W, long long ldw, int M, int N, int K,
                   const GemmEpilogue& epi_in, int num_sms, cudaStream_t st) {
  // CVB_GELU_H2=0: keep the fp32x2 GELU for fp16 outputs as well (A/B switch)
  static const int gelu_h2_on = getenv("CVB_GELU_H2") ? atoi(getenv("CVB_GELU_H2")) : 1;
  GemmEpilogue epi = epi_in;
  epi.gelu_h2 = gelu_h2_on && epi.fp16 && epi.act == GEMM_ACT_GELU && epi.out_bf16 && !epi.out_f32 && !epi.res;
  if (M <= 0 || N <= 0 || K <= 0) return cvb_fail(CV_ERR_INVALID, "gemm: non-positive size");
  if ((N % 16) || (K % 8) || (lda % 8) || (ldw % 8)) return cvb_fail(CV_ERR_INVALID, "gemm: N%16, K%8, lda%8, ldw%8 must be 0");
  if (((uintptr_t)A | (uintptr_t)W) & 15) return cvb_fail(CV_ERR_INVALID, "gemm: operands must be 16-byte aligned");
  if (!epi.out_f32 && !epi.out_bf16) return cvb_fail(CV_ERR_INVALID, "gemm: no output");
  if (epi.map_mode == GEMM_MAP_SHUFFLE2 && (epi.cout % 32)) return cvb_fail(CV_ERR_INVALID, "gemm: shuffle needs cout%32==0");
  if (epi.bias && ((uintptr_t)epi.bias & 15)) return cvb_fail(CV_ERR_INVALID, "gemm: bias must be 16-byte aligned");
  if (epi.res && (((uintptr_t)epi.res & 15) || (epi.ld_res % 4))) return cvb_fail(CV_ERR_INVALID, "gemm: residual alignment");
  if (N % 32) {
    // N = 112 / 144 (Hiera base+ / large stage-1 width): one n-tile wider than N, the last chunk is half valid
    if (epi.map_mode == GEMM_MAP_SHUFFLE2) return cvb_fail(CV_ERR_INVALID, "gemm: shuffle needs N%32==0");
    if (N <= 128) return launch_bn<128>(A, lda, W, ldw, M, N, K, epi, num_sms, st);
    if (N <= 192) return launch_bn<192>(A, lda, W, ldw, M, N, K, epi, num_sms, st);
    if (N <= 256) return launch_bn<256>(A, lda, W, ldw, M, N, K, epi, num_sms, st);
    return cvb_fail(CV_ERR_INVALID, "gemm: N % 32 == 16 is supported up to N = 256");
  }
  if (N % 256 == 0 && K >= 256 && M >= 1024) return launch_bn<256>(A, lda, W, ldw, M, N, K, epi, num_sms, st);
  if (N % 192 == 0) return launch_bn<192>(A, lda, W, ldw, M, N, K, epi, num_sms, st);
  // Hiera base+ widths (112 << s): 224 / 448 / 896 columns are whole 224-wide tiles (they used to fall to 64- or even
  // 32-wide tiles: 26 % of the tensor peak on the base+ fc2 GEMM)
  if (N % 224 == 0) return launch_bn<224>(A, lda, W, ldw, M, N, K, epi, num_sms, st);
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

extern "C" int cv_gemm_ex(const void* A, long long lda, const void* W, long long ldw, int M, int N, int K,
                          const cv_gemm_epilogue* x, void* stream) {
  cvb_reset_launches();
  if (!A || !W || !x || (!x->out_f32 && !x->out_16)) return cvb_fail(CV_ERR_INVALID, "cv_gemm_ex: null pointer");
  GemmEpilogue e;
  e.bias = x->bias;
  e.act = x->act;
  e.res_before_act = x->res_before_act;
  e.res = x->residual;
  e.ld_res = x->ld_res;
  e.res_row_mod = x->res_row_mod;
  e.out_f32 = x->out_f32;
  e.ld_f32 = x->ld_f32;
  e.out_bf16 = (__nv_bfloat16*)x->out_16;
  e.ld_bf16 = x->ld_16;
  e.map_mode = x->map_mode;
  e.ws = x->ws; e.nwx = x->nwx; e.nwy = x->nwy; e.H = x->H; e.W = x->W; e.cout = x->cout;
  e.pool_cols = x->pool_cols;
  e.pool_out = (__nv_bfloat16*)x->pool_out;
  e.ld_pool = x->ld_pool;
  e.fp16 = x->operand_fp16;
  return gemm_tc_launch((const __nv_bfloat16*)A, lda, (const __nv_bfloat16*)W, ldw, M, N, K, e, device_sm_count(),
                        (cudaStream_t)stream);
}

extern "C" int cv_gemm_set_trace(void* device_buffer) {
  g_gemm_trace = (unsigned long long*)device_buffer;
  return CV_OK;
}
