// Persistent warp-specialised bf16 GEMM on tcgen05 + TMA:   C[M,N] = epilogue(A[M,K] * W[N,K]^T)
// A and W are K-major bf16 (row pitch in elements, multiple of 8), fp32 accumulation in TMEM (two accumulator
// stages so the epilogue of tile i overlaps the MMAs of tile i+1).  The epilogue fuses bias, GELU/ReLU, an fp32
// residual read at the destination row, a window-unpartition / pixel-shuffle row remap and fp32 and/or bf16
// outputs, staged through per-warp shared memory so global stores are 128-byte row segments.
#pragma once
#include "tc05.cuh"

namespace cvb {

enum { GEMM_ACT_NONE = 0, GEMM_ACT_GELU = 1, GEMM_ACT_RELU = 2 };
enum { GEMM_MAP_IDENTITY = 0, GEMM_MAP_UNWINDOW = 1, GEMM_MAP_SHUFFLE2 = 2 };

struct GemmEpilogue {
  const float* bias = nullptr;       // [N] (for SHUFFLE2: [N/4], indexed by output channel)
  int act = GEMM_ACT_NONE;
  int res_before_act = 0;            // 0: act(acc+bias)+res   1: act(acc+bias+res)
  const float* res = nullptr;        // fp32 residual, read at the destination row/column
  long long ld_res = 0;
  long long res_row_mod = 0;         // > 0: residual row = dest row % res_row_mod (per-image broadcast tables)
  float* out_f32 = nullptr;
  long long ld_f32 = 0;
  __nv_bfloat16* out_bf16 = nullptr;
  long long ld_bf16 = 0;
  int map_mode = GEMM_MAP_IDENTITY;
  // UNWINDOW: source row = ((b*nwy + wy)*nwx + wx)*ws*ws + ty*ws + tx  ->  dest row (b*H + wy*ws+ty)*W + wx*ws+tx,
  //           rows that fall in the window padding (y >= H or x >= W) are dropped.
  // SHUFFLE2: source row (b,y,x) on an H x W grid, column n = (dy*2+dx)*Cout + co -> dest row
  //           (b*2H + 2y+dy)*2W + 2x+dx, dest column co   (ConvTranspose2d kernel 2 stride 2)
  int ws = 0, nwx = 0, nwy = 0, H = 0, W = 0, cout = 0;
};

struct GemmProblem {
  int M, N, K;
  int n_tiles_m, n_tiles_n;
};

constexpr int GEMM_BM = 128;
constexpr int GEMM_BK = 64;
constexpr int GEMM_STAGES = 4;
constexpr int GEMM_THREADS = 384;  // warp0 TMA, warp1 MMA, warp2 TMEM alloc, warp3 idle, warps 4..11 epilogue
constexpr int GEMM_STAGE_LD = 36;  // staging row pitch in floats (32 + 4: conflict-free float4 rows)

template <int BN>
constexpr int gemm_smem_bytes() {
  return 1024 /*align slack*/ + GEMM_STAGES * (GEMM_BM * 128 + BN * 128) + 8 * 32 * GEMM_STAGE_LD * 4 + 256;
}

__device__ __forceinline__ float gelu_erf(float x) { return 0.5f * x * (1.0f + erff(x * 0.70710678118654752440f)); }

__device__ __forceinline__ long long gemm_dest_row(const GemmEpilogue& e, long long r) {
  if (e.map_mode == GEMM_MAP_UNWINDOW) {
    int w2 = e.ws * e.ws;
    long long win = r / w2;
    int t = (int)(r - win * w2);
    int per_img = e.nwx * e.nwy;
    long long b = win / per_img;
    int wi = (int)(win - b * per_img);
    int wy = wi / e.nwx, wx = wi - wy * e.nwx;
    int ty = t / e.ws, tx = t - ty * e.ws;
    int y = wy * e.ws + ty, x = wx * e.ws + tx;
    if (y >= e.H || x >= e.W) return -1;
    return (b * e.H + y) * e.W + x;
  }
  return r;
}

template <int BN>
__global__ void __launch_bounds__(GEMM_THREADS, 1)
k_gemm_tc(const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_w, GemmProblem p,
          GemmEpilogue e) {
  static_assert(BN % 32 == 0 && BN >= 32 && BN <= 256, "BN");
  constexpr int A_BYTES = GEMM_BM * 128, B_BYTES = BN * 128;
  constexpr int TMEM_COLS = (2 * BN <= 64) ? 64 : (2 * BN <= 128) ? 128 : (2 * BN <= 256) ? 256 : 512;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  uint8_t* sA = smem;
  uint8_t* sB = smem + GEMM_STAGES * A_BYTES;
  float* staging = (float*)(smem + GEMM_STAGES * (A_BYTES + B_BYTES));
  uint64_t* bars = (uint64_t*)(staging + 8 * 32 * GEMM_STAGE_LD);
  uint64_t* full = bars;                     // [STAGES]
  uint64_t* empty = bars + GEMM_STAGES;      // [STAGES]
  uint64_t* tfull = bars + 2 * GEMM_STAGES;  // [2]
  uint64_t* tempty = tfull + 2;              // [2]
  uint32_t* tmem_slot = (uint32_t*)(tempty + 2);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int n_tiles = p.n_tiles_m * p.n_tiles_n;
  const int n_kb = (p.K + GEMM_BK - 1) / GEMM_BK;

  if (warp == 0 && lane == 0) {
    tc::prefetch_tmap(&tmap_a);
    tc::prefetch_tmap(&tmap_w);
  }
  if (warp == 1 && lane == 0) {
    for (int i = 0; i < GEMM_STAGES; i++) {
      tc::mbar_init(&full[i], 1);
      tc::mbar_init(&empty[i], 1);
    }
    for (int i = 0; i < 2; i++) {
      tc::mbar_init(&tfull[i], 1);
      tc::mbar_init(&tempty[i], 8);
    }
    tc::fence_barrier_init();
  }
  if (warp == 2) tc::tmem_alloc<TMEM_COLS>(tmem_slot);
  tc::tc_fence_before();
  __syncthreads();
  tc::tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ===================== TMA producer
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int t = blockIdx.x; t < n_tiles; t += gridDim.x) {
        int mb = t / p.n_tiles_n, nb = t - mb * p.n_tiles_n;
        for (int kb = 0; kb < n_kb; kb++) {
          tc::mbar_wait(&empty[stage], phase ^ 1);
          tc::mbar_arrive_expect_tx(&full[stage], A_BYTES + B_BYTES);
          tc::tma_load_2d(sA + stage * A_BYTES, &tmap_a, &full[stage], kb * GEMM_BK, mb * GEMM_BM);
          tc::tma_load_2d(sB + stage * B_BYTES, &tmap_w, &full[stage], kb * GEMM_BK, nb * BN);
          if (++stage == GEMM_STAGES) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer (one thread)
    if (lane == 0) {
      constexpr uint32_t idesc = tc::idesc_bf16(GEMM_BM, BN, false, false);
      int stage = 0;
      uint32_t phase = 0;
      int acc = 0;
      uint32_t acc_phase = 0;
      for (int t = blockIdx.x; t < n_tiles; t += gridDim.x) {
        tc::mbar_wait(&tempty[acc], acc_phase ^ 1);
        tc::tc_fence_after();
        uint32_t d_tmem = tmem_base + acc * BN;
        for (int kb = 0; kb < n_kb; kb++) {
          tc::mbar_wait(&full[stage], phase);
          tc::tc_fence_after();
          uint32_t a0 = tc::smem_u32(sA + stage * A_BYTES), b0 = tc::smem_u32(sB + stage * B_BYTES);
          int ksteps = min(GEMM_BK / 16, (p.K - kb * GEMM_BK + 15) / 16);
          for (int k = 0; k < ksteps; k++)
            tc::mma_f16_ss(d_tmem, tc::desc_kmajor(a0 + k * 32), tc::desc_kmajor(b0 + k * 32), idesc,
                           (kb > 0 || k > 0) ? 1u : 0u);
          tc::mma_commit(&empty[stage]);
          if (++stage == GEMM_STAGES) { stage = 0; phase ^= 1; }
        }
        tc::mma_commit(&tfull[acc]);
        if (++acc == 2) { acc = 0; acc_phase ^= 1; }
      }
    }
  } else if (warp >= 4) {
    // ===================== epilogue: 8 warps; warp e handles TMEM lanes 32*(e%4).. and column chunks c with c%2 == e/4
    const int ew = warp - 4;
    const int quad = warp & 3;  // TMEM lane quadrant this warp may access (= warp id mod 4)
    const int cgroup = ew >> 2;
    float* stg = staging + ew * 32 * GEMM_STAGE_LD;
    int acc = 0;
    uint32_t acc_phase = 0;
    const int sub_row = lane >> 3, sub_col = (lane & 7) * 4;
    for (int t = blockIdx.x; t < n_tiles; t += gridDim.x) {
      int mb = t / p.n_tiles_n, nb = t - mb * p.n_tiles_n;
      long long row0 = (long long)mb * GEMM_BM + quad * 32;
      // destination rows of the 8 rows this thread copies out
      long long dest[8];
#pragma unroll
      for (int it = 0; it < 8; it++) {
        long long r = row0 + it * 4 + sub_row;
        dest[it] = (r < p.M) ? gemm_dest_row(e, r) : -1;
      }
      tc::mbar_wait(&tfull[acc], acc_phase);
      tc::tc_fence_after();
      for (int c = cgroup; c < BN / 32; c += 2) {
        uint32_t v[32];
        tc::tmem_ld_32x32(tmem_base + ((uint32_t)(quad * 32) << 16) + acc * BN + c * 32, v);
        tc::tmem_ld_wait();
        int col0 = nb * BN + c * 32;
        if (col0 < p.N) {
          // SHUFFLE2: this 32-column chunk lies inside one (dy,dx) group; bias/out column = co
          int q = 0, ocol0 = col0;
          if (e.map_mode == GEMM_MAP_SHUFFLE2) { q = col0 / e.cout; ocol0 = col0 - q * e.cout; }
          float f[32];
#pragma unroll
          for (int i = 0; i < 32; i++) {
            float x = __uint_as_float(v[i]);
            if (e.bias) x += __ldg(e.bias + ocol0 + i);
            if (!e.res_before_act) {
              if (e.act == GEMM_ACT_GELU) x = gelu_erf(x);
              else if (e.act == GEMM_ACT_RELU) x = fmaxf(x, 0.f);
            }
            f[i] = x;
          }
#pragma unroll
          for (int i = 0; i < 32; i += 4)
            *(float4*)(stg + lane * GEMM_STAGE_LD + i) = make_float4(f[i], f[i + 1], f[i + 2], f[i + 3]);
          __syncwarp();
#pragma unroll
          for (int it = 0; it < 8; it++) {
            long long d = dest[it];
            if (d < 0) continue;
            float4 x = *(const float4*)(stg + (it * 4 + sub_row) * GEMM_STAGE_LD + sub_col);
            if (e.map_mode == GEMM_MAP_SHUFFLE2) {
              long long img = d / ((long long)e.H * e.W);
              int rem = (int)(d - img * e.H * e.W);
              int y = rem / e.W, xx = rem - y * e.W;
              d = (img * 2 * e.H + 2 * y + (q >> 1)) * 2 * e.W + 2 * xx + (q & 1);
            }
            int oc = ocol0 + sub_col;
            if (e.res) {
              long long rr = e.res_row_mod > 0 ? d % e.res_row_mod : d;
              float4 r = *(const float4*)(e.res + rr * e.ld_res + oc);
              x.x += r.x; x.y += r.y; x.z += r.z; x.w += r.w;
            }
            if (e.res_before_act) {
              if (e.act == GEMM_ACT_GELU) { x.x = gelu_erf(x.x); x.y = gelu_erf(x.y); x.z = gelu_erf(x.z); x.w = gelu_erf(x.w); }
              else if (e.act == GEMM_ACT_RELU) { x.x = fmaxf(x.x, 0.f); x.y = fmaxf(x.y, 0.f); x.z = fmaxf(x.z, 0.f); x.w = fmaxf(x.w, 0.f); }
            }
            if (e.out_f32) *(float4*)(e.out_f32 + d * e.ld_f32 + oc) = x;
            if (e.out_bf16) {
              __nv_bfloat162 lo = __floats2bfloat162_rn(x.x, x.y), hi = __floats2bfloat162_rn(x.z, x.w);
              uint2 pk;
              pk.x = *(uint32_t*)&lo;
              pk.y = *(uint32_t*)&hi;
              *(uint2*)(e.out_bf16 + d * e.ld_bf16 + oc) = pk;
            }
          }
          __syncwarp();
        }
      }
      tc::tc_fence_before();
      __syncwarp();
      if (lane == 0) tc::mbar_arrive(&tempty[acc]);
      if (++acc == 2) { acc = 0; acc_phase ^= 1; }
    }
  }
  tc::tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    tc::tc_fence_after();
    tc::tmem_dealloc<TMEM_COLS>(tmem_base);
  }
}

// Host launcher.  Returns a cudaError_t-compatible int (0 = ok); -1 = tensor-map failure.
int gemm_tc_launch(const __nv_bfloat16* A, long long lda, const __nv_bfloat16* W, long long ldw, int M, int N, int K,
                   const GemmEpilogue& epi, int num_sms, cudaStream_t st);

}  // namespace cvb
