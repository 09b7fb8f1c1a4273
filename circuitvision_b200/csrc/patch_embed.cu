// Fused patch embedding (interface: patch_embed.cuh).
//
// The 7 x 7 / stride 4 convolution of the Hiera trunk is a GEMM [tokens, 168] x [168, E] whose left operand is a gather of raw
// pixels.  Round 1 materialised that operand (k_im2col_u8raw: 22 MB per image written, read again by k_gemm_tc) and then ran
// k_ln_rows over the result for block 0's norm1.  Here one persistent kernel does all three: per 128-token tile
//
//   warps 4-7    operand producers: thread = token; for each of the 7 kernel rows 24 raw bytes -> 24 16-bit values (exact) -> three
//                16-byte chunks of the K-major 128B-swizzled A tile in shared memory (double-buffered)
//   warp 0       loads the weights [E, 168] once (three TMA boxes, resident for the whole kernel)
//   warp 1       MMA issuer: 11 tcgen05.mma (128 x E x 16) per tile into one of two TMEM accumulators
//   warps 8-15   epilogue, two groups of four (group g owns accumulator g, i.e. every second tile; thread = token row):
//                x = acc + pos -> swizzled staging -> TMA store to X0; with LayerNorm: x is also written back to TMEM, the row
//                statistics are taken in two more sweeps (mean, then squared deviations, as k_ln_rows does) and the normalised
//                16-bit row goes out in 8 x 8 window-major order (64-byte pieces, 8 consecutive tokens = 8 consecutive rows)
//
// HBM sees 3 MB of pixels in and 24 (+12) MB out per image instead of 3 + 22 + 22 + 24 (+ 24 + 12).
#include "patch_embed.cuh"

#include <stdlib.h>

#include "common.cuh"
#include "tc05.cuh"

namespace cvb {

constexpr int PEK = 168;             // 7 kernel rows x 24 (21 taps + 3 zeros)
constexpr int PE_THREADS = 16 * 32;  // 4 control + 4 producer + 8 epilogue warps
constexpr int PE_A_BYTES = 3 * 128 * 128;  // three 64-wide K blocks of a 128-row tile
constexpr int PE_STAGE_BYTES = 4096;       // one 32 x 32 fp32 staging box

static inline int pe_smem_bytes(int E) { return 1024 + 3 * E * 128 + 2 * PE_A_BYTES + 8 * 2 * PE_STAGE_BYTES + 256; }

template <bool SWAP>
__global__ void __launch_bounds__(PE_THREADS, 1)
k_patch_embed(const __grid_constant__ CUtensorMap tmap_w, const __grid_constant__ CUtensorMap tmap_x, PatchEmbedArgs a, int n_tiles) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  const int E = a.E;
  uint8_t* sW = smem;                       // [3][E rows][128 B]
  uint8_t* sA = sW + 3 * E * 128;           // [2][3][128 rows][128 B]   (E * 128 is a multiple of 1024 for E % 8 == 0)
  uint8_t* sS = sA + 2 * PE_A_BYTES;        // [8 warps][2][4096]
  uint64_t* bars = (uint64_t*)(sS + 8 * 2 * PE_STAGE_BYTES);
  uint64_t* w_full = bars;                  // [1]
  uint64_t* a_full = bars + 1;              // [2]
  uint64_t* a_empty = bars + 3;             // [2]
  uint64_t* t_full = bars + 5;              // [2]
  uint64_t* t_empty = bars + 7;             // [2]
  uint32_t* tmem_slot = (uint32_t*)(bars + 9);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int G = a.S >> 2;                   // tokens per image row
  const int tiles_per_row = G >> 7;
  const int tiles_per_img = G * tiles_per_row;

  if (warp == 0 && lane == 0) {
    tc::prefetch_tmap(&tmap_w);
    tc::prefetch_tmap(&tmap_x);
  }
  if (warp == 1 && lane == 0) {
    tc::mbar_init(w_full, 1);
    for (int i = 0; i < 2; i++) {
      tc::mbar_init(&a_full[i], 4);
      tc::mbar_init(&a_empty[i], 1);
      tc::mbar_init(&t_full[i], 1);
      tc::mbar_init(&t_empty[i], 4);
    }
    tc::fence_barrier_init();
  }
  if (warp == 2) tc::tmem_alloc<512>(tmem_slot);
  // K columns 168..191 of both A buffers are never produced: zero them once (the last k-step reads 160..175)
  for (int i = threadIdx.x; i < 2 * 128; i += PE_THREADS) {
    const int buf = i >> 7, r = i & 127;
    uint8_t* base = sA + buf * PE_A_BYTES + 2 * 16384 + r * 128;
#pragma unroll
    for (int c = 5; c < 8; c++) *(uint4*)(base + ((c ^ (r & 7)) << 4)) = make_uint4(0u, 0u, 0u, 0u);
  }
  tc::fence_proxy_async_smem();
  tc::tc_fence_before();
  __syncthreads();
  tc::tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    if (lane == 0) {
      tc::mbar_arrive_expect_tx(w_full, 3 * E * 128);
      for (int kb = 0; kb < 3; kb++) tc::tma_load_2d(sW + kb * E * 128, &tmap_w, w_full, kb * 64, 0);
    }
  } else if (warp == 1) {
    // ===================== MMA issuer (converged warp, one elected lane issues)
    const uint32_t idesc = tc::idesc_bf16(128, E, false, false, a.fp16 != 0);
    tc::mbar_wait(w_full, 0);
    int it = 0;
    for (int t = blockIdx.x; t < n_tiles; t += gridDim.x, it++) {
      const int s = it & 1;
      const uint32_t ph = (it >> 1) & 1;
      tc::mbar_wait(&t_empty[s], ph ^ 1);
      tc::mbar_wait(&a_full[s], ph);
      tc::tc_fence_after();
      const uint32_t d_tmem = tmem_base + s * E;
      const uint32_t a0 = tc::smem_u32(sA + s * PE_A_BYTES), w0 = tc::smem_u32(sW);
      if (tc::elect_one()) {
#pragma unroll
        for (int kb = 0; kb < 3; kb++) {
          const uint64_t da = tc::desc_kmajor(a0 + kb * 16384), dw = tc::desc_kmajor(w0 + kb * E * 128);
#pragma unroll
          for (int k = 0; k < 4; k++)
            if (kb < 2 || k < 3) tc::mma_f16_ss(d_tmem, da + 2 * k, dw + 2 * k, idesc, (kb | k) ? 1u : 0u);
        }
        tc::mma_commit(&a_empty[s]);
        tc::mma_commit(&t_full[s]);
      }
      __syncwarp();
    }
  } else if (warp >= 4 && warp < 8) {
    // ===================== operand producers: thread = token of the tile
    const int r = threadIdx.x - 128;
    int it = 0;
    for (int t = blockIdx.x; t < n_tiles; t += gridDim.x, it++) {
      const int s = it & 1;
      const uint32_t ph = (it >> 1) & 1;
      const int b = t / tiles_per_img, rem = t - b * tiles_per_img;
      const int y = rem / tiles_per_row, x = (rem - y * tiles_per_row) * 128 + r;
      // 7 kernel rows x 6 words: bytes 12 x - 12 .. 12 x + 11 of image row 4 y - 3 + ky (pixels 4 x - 3 .. 4 x + 3 are bytes
      // 12 x - 9 .. 12 x + 11); only token 0 has words left of the row, rows outside the image are zero padding
      uint32_t w[7][6];
      const int w0 = 3 * x - 3;
#pragma unroll
      for (int ky = 0; ky < 7; ky++) {
        const int iy = y * 4 - 3 + ky;
        const bool row_ok = iy >= 0 && iy < a.S;
        const uint32_t* row = (const uint32_t*)(a.img + ((size_t)b * a.S + (row_ok ? iy : 0)) * a.S * 3);
#pragma unroll
        for (int i = 0; i < 6; i++) w[ky][i] = (row_ok && w0 + i >= 0) ? __ldg(row + w0 + i) : 0u;
      }
      tc::mbar_wait(&a_empty[s], ph ^ 1);
      uint8_t* arow = sA + s * PE_A_BYTES + r * 128;
#pragma unroll
      for (int ky = 0; ky < 7; ky++) {
        // byte j of the 21-byte segment -> fp32 by the 2^23 trick (PRMT builds 0x4B0000bb), exact for 0..255
        auto val = [&](int j) -> float {
          if (j >= 21) return 0.f;
          if (SWAP) j = 3 * (j / 3) + 2 - (j % 3);  // BGR input: channel c of pixel kx lives at byte 2 - c
          const int byte = j + 3;
          return __uint_as_float(__byte_perm(w[ky][byte >> 2], 0x4B000000u, 0x7540 + (byte & 3))) - 8388608.0f;
        };
        uint32_t o[12];
#pragma unroll
        for (int i = 0; i < 12; i++) o[i] = tc::pack16(a.fp16, val(2 * i), val(2 * i + 1));
#pragma unroll
        for (int i = 0; i < 3; i++) {
          const int q = ky * 3 + i;  // 16-byte chunk of the 336-byte operand row
          *(uint4*)(arow + (q >> 3) * 16384 + (((q & 7) ^ (r & 7)) << 4)) = make_uint4(o[4 * i], o[4 * i + 1], o[4 * i + 2], o[4 * i + 3]);
        }
      }
      tc::fence_proxy_async_smem();
      __syncwarp();
      if (lane == 0) tc::mbar_arrive(&a_full[s]);
    }
  } else if (warp >= 8) {
    // ===================== epilogue: group g = accumulator g = tiles of local parity g; thread = token row
    const int g = (warp - 8) >> 2, quad = warp & 3;
    const int r = quad * 32 + lane;
    uint8_t* stage = sS + (warp - 8) * 2 * PE_STAGE_BYTES;
    const uint32_t tacc = tmem_base + ((uint32_t)(quad * 32) << 16) + g * E;
    const int nch = (E + 31) >> 5;
    const bool ln = a.A16 != nullptr;
    const float inv_e = 1.0f / (float)E;
    uint32_t nbuf = 0;
    int it = 0;
    for (int t = blockIdx.x; t < n_tiles; t += gridDim.x, it++) {
      if ((it & 1) != g) continue;
      const uint32_t ph = (it >> 1) & 1;
      const int b = t / tiles_per_img, rem = t - b * tiles_per_img;
      const int y = rem / tiles_per_row, x0 = (rem - y * tiles_per_row) * 128;
      const float* prow = a.pos + (size_t)(rem * 128 + r) * E;   // row of the positional table
      tc::mbar_wait(&t_full[g], ph);
      tc::tc_fence_after();
      float sum = 0.f;
#pragma unroll 1
      for (int c = 0; c < nch; c++) {
        const int ncols = E - c * 32;  // >= 32 or 16
        float4 rv[8];
#pragma unroll
        for (int j = 0; j < 8; j++) rv[j] = 4 * j < ncols ? __ldg((const float4*)(prow + c * 32) + j) : make_float4(0.f, 0.f, 0.f, 0.f);
        uint32_t v[32];
        tc::tmem_ld_32x32(tacc + c * 32, v);
        uint8_t* buf = stage + (nbuf & 1) * PE_STAGE_BYTES;
        nbuf++;
        if (lane == 0) tc::tma_store_wait_read<1>();
        __syncwarp();
        tc::tmem_ld_wait();
#pragma unroll
        for (int j = 0; j < 8; j++) {
          float4 o;
          o.x = __uint_as_float(v[4 * j + 0]) + rv[j].x;
          o.y = __uint_as_float(v[4 * j + 1]) + rv[j].y;
          o.z = __uint_as_float(v[4 * j + 2]) + rv[j].z;
          o.w = __uint_as_float(v[4 * j + 3]) + rv[j].w;
          if (4 * j < ncols) sum += (o.x + o.y) + (o.z + o.w);
          v[4 * j + 0] = __float_as_uint(o.x); v[4 * j + 1] = __float_as_uint(o.y);
          v[4 * j + 2] = __float_as_uint(o.z); v[4 * j + 3] = __float_as_uint(o.w);
          *(float4*)(buf + lane * 128 + ((j ^ (lane & 7)) << 4)) = o;
        }
        if (ln) {
          if (ncols >= 32) tc::tmem_st_32x32(tacc + c * 32, v);
          else tc::tmem_st_32x16(tacc + c * 32, *(const uint32_t(*)[16]) & v[0]);
        }
        tc::fence_proxy_async_smem();
        __syncwarp();
        if (lane == 0) {
          tc::tma_store_2d(&tmap_x, buf, c * 32, (int)((long long)t * 128 + quad * 32));
          tc::tma_store_commit();
        }
      }
      if (ln) {
        tc::tmem_st_wait();
        const float mean = sum * inv_e;
        float q = 0.f;
#pragma unroll 1
        for (int c = 0; c < nch; c++) {
          const int ncols = E - c * 32;
          uint32_t v[32];
          tc::tmem_ld_32x32(tacc + c * 32, v);
          tc::tmem_ld_wait();
#pragma unroll
          for (int i = 0; i < 32; i++) {
            const float d = __uint_as_float(v[i]) - mean;
            if (i < ncols) q = fmaf(d, d, q);
          }
        }
        const float rstd = rsqrtf(q * inv_e + a.eps);
        // destination rows: token (y, x0 + r) of image b in 8 x 8 window-major order
        const int nw = G >> 3;
        const int xx = x0 + r;
        const long long drow = (((long long)b * nw + (y >> 3)) * nw + (xx >> 3)) * 64 + (y & 7) * 8 + (xx & 7);
        if (lane == 0) tc::tma_store_wait_read<0>();  // the staging buffers are reused as the 16-bit staging
        __syncwarp();
#pragma unroll 1
        for (int c = 0; c < nch; c++) {
          const int ncols = E - c * 32;
          uint32_t v[32];
          tc::tmem_ld_32x32(tacc + c * 32, v);
          tc::tmem_ld_wait();
          // 32 rows x 64 bytes, pitch 80 bytes (conflict-free 16-byte accesses both ways)
          uint8_t* sb = stage + (c & 1) * PE_STAGE_BYTES;
#pragma unroll
          for (int j = 0; j < 4; j++) {
            uint32_t w[4];
#pragma unroll
            for (int k = 0; k < 2; k++) {
              const bool okc = 4 * (2 * j + k) < ncols;
              const float4 gg = okc ? __ldg((const float4*)(a.gamma + c * 32) + 2 * j + k) : make_float4(0.f, 0.f, 0.f, 0.f);
              const float4 bb = okc ? __ldg((const float4*)(a.beta + c * 32) + 2 * j + k) : make_float4(0.f, 0.f, 0.f, 0.f);
              const float o0 = (__uint_as_float(v[8 * j + 4 * k + 0]) - mean) * rstd * gg.x + bb.x;
              const float o1 = (__uint_as_float(v[8 * j + 4 * k + 1]) - mean) * rstd * gg.y + bb.y;
              const float o2 = (__uint_as_float(v[8 * j + 4 * k + 2]) - mean) * rstd * gg.z + bb.z;
              const float o3 = (__uint_as_float(v[8 * j + 4 * k + 3]) - mean) * rstd * gg.w + bb.w;
              w[2 * k] = tc::pack16(a.fp16, o0, o1);
              w[2 * k + 1] = tc::pack16(a.fp16, o2, o3);
            }
            *(uint4*)(sb + lane * 80 + j * 16) = make_uint4(w[0], w[1], w[2], w[3]);
          }
          __syncwarp();
          // copy out: lane -> (row lane / 4 + 8 i, 16-byte piece lane % 4): 4 lanes write 64 contiguous bytes of a row
#pragma unroll
          for (int i = 0; i < 4; i++) {
            const int rr = (lane >> 2) + 8 * i, pc = lane & 3;
            const long long d = __shfl_sync(0xffffffffu, drow, rr);
            if (pc * 8 < ncols)
              *(uint4*)(a.A16 + d * E + c * 32 + pc * 8) = *(const uint4*)(sb + rr * 80 + pc * 16);
          }
          __syncwarp();
        }
      }
      tc::tc_fence_before();
      __syncwarp();
      if (lane == 0) tc::mbar_arrive(&t_empty[g]);
    }
    if (lane == 0) tc::tma_store_wait<0>();
  }
  tc::tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    tc::tc_fence_after();
    tc::tmem_dealloc<512>(tmem_base);
  }
}

bool patch_embed_supported(int E, int S) {
  static const int on = getenv("CVB_PATCH_FUSED") ? atoi(getenv("CVB_PATCH_FUSED")) : 1;
  return on && (E % 16) == 0 && E >= 32 && E <= 160 && S > 0 && (S % 512) == 0;
}

int patch_embed_launch(const PatchEmbedArgs& a, int num_sms, cudaStream_t st) {
  if (!patch_embed_supported(a.E, a.S)) return cvb_fail(CV_ERR_INVALID, "patch_embed: unsupported width / image size");
  if (((uintptr_t)a.img & 3) || ((uintptr_t)a.W & 15) || ((uintptr_t)a.pos & 15) || ((uintptr_t)a.X0 & 15))
    return cvb_fail(CV_ERR_INVALID, "patch_embed: misaligned pointer");
  if (a.A16 && (!a.gamma || !a.beta || ((uintptr_t)a.A16 & 15) || ((uintptr_t)a.gamma & 15) || ((uintptr_t)a.beta & 15)))
    return cvb_fail(CV_ERR_INVALID, "patch_embed: LayerNorm output needs aligned gamma / beta / A16");
  const int G = a.S / 4;
  const long long rows = (long long)a.B * G * G;
  const int n_tiles = (int)(rows / 128);
  const int smem = pe_smem_bytes(a.E);
  if (smem > 232448) return cvb_fail(CV_ERR_INVALID, "patch_embed: shared memory budget");
  static std::atomic<unsigned long long> attr_set{0};
  if (cvb_once_per_device(attr_set)) {
    cudaError_t e = cudaFuncSetAttribute(k_patch_embed<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 232448);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(k_patch_embed<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 232448);
    if (e != cudaSuccess) return cvb_fail_cuda(e, "cudaFuncSetAttribute(k_patch_embed)");
  }
  CUtensorMap tw, tx;
  if (!tc_host::make_tmap_bf16(&tw, a.W, (uint64_t)a.E, (uint64_t)PEK, (uint64_t)PEK, (uint32_t)a.E) ||
      !tc_host::make_tmap_2d(&tx, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, a.X0, (uint64_t)rows, (uint64_t)a.E, (uint64_t)a.E * 4, 32, 32,
                             CU_TENSOR_MAP_SWIZZLE_128B))
    return cvb_fail(CV_ERR_CUDA, "cuTensorMapEncodeTiled failed (patch embed)");
  cvb_next_work(2.0 * (double)rows * a.E * 147);
  if (cvb_profile_on()) {
    char nm[96];
    snprintf(nm, sizeof(nm), "patch_embed M%lld E%d%s", rows, a.E, a.A16 ? " +LN1" : "");
    cvb_next_name(nm);
  }
  const int grid = n_tiles < num_sms ? n_tiles : num_sms;
  if (a.swap_rb) {
    CVB_LAUNCH((k_patch_embed<true>), dim3(grid), dim3(PE_THREADS), smem, st, tw, tx, a, n_tiles);
  } else {
    CVB_LAUNCH((k_patch_embed<false>), dim3(grid), dim3(PE_THREADS), smem, st, tw, tx, a, n_tiles);
  }
  return CV_OK;
}

}  // namespace cvb
