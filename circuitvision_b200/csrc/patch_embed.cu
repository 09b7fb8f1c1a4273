// Fused patch embedding (interface: patch_embed.cuh).
//
// The 7 x 7 / stride 4 convolution of the Hiera trunk is a GEMM [tokens, 168] x [168, E] whose left operand is a gather of raw
// pixels.  Round 1 materialised that operand (k_im2col_u8raw: 22 MB per image written, read again by k_gemm_tc) and then ran
// k_ln_rows over the result for block 0's norm1.  Here one persistent kernel does all three: per 128-token tile
//
//   warp 3       raw pixel rows: seven 1.5 KB bulk copies (cp.async.bulk) per tile into a two-slot ring, two tiles ahead
//   warps 4-7    operand producers: thread = token; for each of the 7 kernel rows 24 raw bytes -> 24 16-bit values (exact) -> three
//                16-byte chunks of the K-major 128B-swizzled A tile in shared memory (double-buffered)
//   warp 0       loads the weights [E, 168] once (three TMA boxes, resident for the whole kernel)
//   warp 1       MMA issuer: 11 tcgen05.mma (128 x E x 16) per tile into one of FOUR TMEM accumulators
//   warps 8-15   epilogue, two groups of four (group g takes every second tile; thread = token row): x = acc + pos -> swizzled
//                staging -> TMA store to X0.  With LayerNorm: x is also written back to TMEM, the row statistics are combined from
//                per-chunk (sum, centred sum of squares) pairs (Chan's formula: as robust as the two-sweep form of k_ln_rows, one
//                sweep less) and a second sweep normalises: 16-bit rows out in 8 x 8 window-major order (8 consecutive tokens =
//                8 consecutive rows: four 8-row TMA stores per 32-column chunk)
//
// HBM sees 3 MB of pixels in and 24 (+12) MB out per image instead of 3 + 22 + 22 + 24 (+ 24 + 12).
// Measured steps (tiny, 64 crops, B200): im2col 0.48 + GEMM 1.05 + k_ln_rows 0.46 ms -> 1.25 ms (first fused version, row-per-thread
// global loads: the LSU wavefront count, 32 per instruction, was the limiter) -> 0.99 ms (pixels by bulk copy, positional rows
// loaded lanes-along-columns, 16-bit rows by TMA) -> see profiles/README.md for the current figure.
#include "patch_embed.cuh"

#include <stdlib.h>

#include "common.cuh"
#include "tc05.cuh"

namespace cvb {

constexpr int PEK = 168;             // 7 kernel rows x 24 (21 taps + 3 zeros)
constexpr int PE_THREADS = 16 * 32;  // 4 control + 4 producer + 8 epilogue warps
constexpr int PE_A_BYTES = 3 * 128 * 128;  // three 64-wide K blocks of a 128-row tile
constexpr int PE_STAGE_BYTES = 4096;       // one 32 x 32 fp32 staging box
constexpr int PE_RAW_PITCH = 1552;         // bytes of one raw pixel row segment: 16 (alignment slack) + 128 tokens x 12
constexpr int PE_RAW_SLOT = 7 * PE_RAW_PITCH;  // 7 kernel rows; one extra segment of zeros after both slots (rows outside the image)
constexpr int PE_NACC = 4;                 // TMEM accumulators (4 x E <= 512 columns)

static inline int pe_smem_bytes(int E) {
  return 1024 + 3 * E * 128 + 2 * PE_A_BYTES + 8 * 2 * PE_STAGE_BYTES + 2 * PE_RAW_SLOT + PE_RAW_PITCH + 512;
}

__device__ __forceinline__ void sts128(uint32_t addr, float4 v) {
  asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}
__device__ __forceinline__ void sts128u(uint32_t addr, uint4 v) {
  asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}
__device__ __forceinline__ float4 lds128(uint32_t addr) {
  float4 v;
  asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(addr) : "memory");
  return v;
}
__device__ __forceinline__ uint32_t lds32(uint32_t addr) {
  uint32_t v;
  asm volatile("ld.shared.b32 %0, [%1];" : "=r"(v) : "r"(addr) : "memory");
  return v;
}

template <int E, bool SWAP, bool FP16>
__global__ void __launch_bounds__(PE_THREADS, 1)
k_patch_embed(const __grid_constant__ CUtensorMap tmap_w, const __grid_constant__ CUtensorMap tmap_x,
              const __grid_constant__ CUtensorMap tmap_a16, PatchEmbedArgs a, int n_tiles) {
  static_assert(E % 16 == 0 && PE_NACC * E <= 512, "E");
  constexpr int NCH = (E + 31) / 32;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  uint8_t* sW = smem;                       // [3][E rows][128 B]
  uint8_t* sA = sW + 3 * E * 128;           // [2][3][128 rows][128 B]   (E * 128 is a multiple of 1024 for E % 8 == 0)
  uint8_t* sS = sA + 2 * PE_A_BYTES;        // [8 warps][2][4096]
  uint8_t* sR = sS + 8 * 2 * PE_STAGE_BYTES;  // [2] raw pixel rows of a tile (7 row segments + a zero segment)
  uint8_t* sZ = sR + 2 * PE_RAW_SLOT;         // zero segment
  uint64_t* bars = (uint64_t*)(sZ + PE_RAW_PITCH);
  uint64_t* w_full = bars;                  // [1]
  uint64_t* a_full = bars + 1;              // [2]
  uint64_t* a_empty = bars + 3;             // [2]
  uint64_t* r_full = bars + 5;              // [2]
  uint64_t* r_empty = bars + 7;             // [2]
  uint64_t* t_full = bars + 9;              // [PE_NACC]
  uint64_t* t_empty = bars + 9 + PE_NACC;   // [PE_NACC]
  uint32_t* tmem_slot = (uint32_t*)(bars + 9 + 2 * PE_NACC);

  const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0), lane = threadIdx.x & 31;  // shuffle: warp-uniform for the compiler
  const int G = a.S >> 2;                   // tokens per image row
  const int tiles_per_row = G >> 7;
  const int tiles_per_img = G * tiles_per_row;

  if (warp == 0 && lane == 0) {
    tc::prefetch_tmap(&tmap_w);
    tc::prefetch_tmap(&tmap_x);
    if (a.A16) tc::prefetch_tmap(&tmap_a16);
  }
  if (warp == 1 && lane == 0) {
    tc::mbar_init(w_full, 1);
    for (int i = 0; i < 2; i++) {
      tc::mbar_init(&a_full[i], 4);
      tc::mbar_init(&a_empty[i], 1);
      tc::mbar_init(&r_full[i], 1);
      tc::mbar_init(&r_empty[i], 4);
    }
    for (int i = 0; i < PE_NACC; i++) {
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
  for (int i = threadIdx.x; i < PE_RAW_PITCH / 16; i += PE_THREADS) *(uint4*)(sZ + i * 16) = make_uint4(0u, 0u, 0u, 0u);
  tc::fence_proxy_async_smem();
  tc::tc_fence_before();
  __syncthreads();
  tc::tc_fence_after();
  const uint32_t tmem_base = __shfl_sync(0xffffffffu, *tmem_slot, 0);

  if (warp == 0) {
    if (lane == 0) {
      tc::mbar_arrive_expect_tx(w_full, 3 * E * 128);
      for (int kb = 0; kb < 3; kb++) tc::tma_load_2d(sW + kb * E * 128, &tmap_w, w_full, kb * 64, 0);
    }
  } else if (warp == 1) {
    // ===================== MMA issuer (converged warp, one elected lane issues)
    const uint32_t idesc = tc::idesc_bf16(128, E, false, false, FP16);
    tc::mbar_wait(w_full, 0);
    int it = 0;
    for (int t = blockIdx.x; t < n_tiles; t += gridDim.x, it++) {
      const int s = it & 1, acc = it & (PE_NACC - 1);
      tc::mbar_wait(&t_empty[acc], ((it / PE_NACC) & 1) ^ 1);
      tc::mbar_wait(&a_full[s], (it >> 1) & 1);
      tc::tc_fence_after();
      const uint32_t d_tmem = tmem_base + acc * E;
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
        tc::mma_commit(&t_full[acc]);
      }
      __syncwarp();
    }
  } else if (warp == 3) {
    // ===================== raw pixel rows: 7 bulk copies per tile (bytes 12 x0 - 16 .. 12 x0 + 1536 of image rows 4 y - 3 ..
    // 4 y + 3; the first tile of a row starts at byte 0 and lands 16 bytes into the segment), two tiles ahead of the producers
    if (lane == 0) {
      int it = 0;
      for (int t = blockIdx.x; t < n_tiles; t += gridDim.x, it++) {
        const int s = it & 1;
        const uint32_t ph = (it >> 1) & 1;
        const int b = t / tiles_per_img, rem = t - b * tiles_per_img;
        const int y = rem / tiles_per_row, xt = rem - y * tiles_per_row;
        tc::mbar_wait(&r_empty[s], ph ^ 1);
        const int lo = max(0, 3 - 4 * y), hi = min(7, a.S + 3 - 4 * y);  // kernel rows inside the image
        const uint32_t bytes = xt ? PE_RAW_PITCH : PE_RAW_PITCH - 16;
        if (!xt) {  // the 16 slack bytes in front of a segment are "left of the image" here: zero (an earlier tile's copy filled them)
          for (int ky = lo; ky < hi; ky++) *(uint4*)(sR + s * PE_RAW_SLOT + ky * PE_RAW_PITCH) = make_uint4(0u, 0u, 0u, 0u);
        }
        tc::mbar_arrive_expect_tx(&r_full[s], bytes * (uint32_t)(hi - lo));
        for (int ky = lo; ky < hi; ky++) {
          const uint8_t* src = a.img + ((size_t)b * a.S + (y * 4 - 3 + ky)) * a.S * 3 + (xt ? xt * 1536 - 16 : 0);
          tc::bulk_load_1d(sR + s * PE_RAW_SLOT + ky * PE_RAW_PITCH + (xt ? 0 : 16), src, bytes, &r_full[s]);
        }
      }
    }
  } else if (warp >= 4 && warp < 8) {
    // ===================== operand producers: thread = token of the tile
    const int r = threadIdx.x - 128;
    int it = 0;
    for (int t = blockIdx.x; t < n_tiles; t += gridDim.x, it++) {
      const int s = it & 1;
      const uint32_t ph = (it >> 1) & 1;
      const int rem = t % tiles_per_img;
      const int y = rem / tiles_per_row;
      // 7 kernel rows x 6 words: bytes 12 x - 12 .. 12 x + 11 of image row 4 y - 3 + ky (pixels 4 x - 3 .. 4 x + 3 are bytes
      // 12 x - 9 .. 12 x + 11) = words 3 r + 1 .. 3 r + 6 of the row's segment (stride 3 words: conflict-free); kernel rows
      // outside the image read the zero segment, the bytes left of the image are the zeroed slack of the segment
      uint32_t w[7][6];
      tc::mbar_wait(&r_full[s], ph);
      const uint32_t raw = tc::smem_u32(sR + s * PE_RAW_SLOT) + (3 * r + 1) * 4, zseg = tc::smem_u32(sZ) + (3 * r + 1) * 4;
#pragma unroll
      for (int ky = 0; ky < 7; ky++) {
        const int iy = y * 4 - 3 + ky;
        const uint32_t seg = (iy >= 0 && iy < a.S) ? raw + ky * PE_RAW_PITCH : zseg;
#pragma unroll
        for (int i = 0; i < 6; i++) w[ky][i] = lds32(seg + i * 4);
      }
      // (the raw slot is NOT released here: ld.shared followed by an mbarrier arrive does not order the loads' data return
      //  before the next bulk copy into the slot — the copy engine overwrote bytes some lanes had not received yet, seen as
      //  run-to-run differences of a few tokens per ~10 images; the slot is released below, after every word has been consumed)
      tc::mbar_wait(&a_empty[s], ph ^ 1);
      const uint32_t arow = tc::smem_u32(sA + s * PE_A_BYTES) + r * 128;
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
        for (int i = 0; i < 12; i++) o[i] = tc::pack16(FP16 ? 1 : 0, val(2 * i), val(2 * i + 1));
#pragma unroll
        for (int i = 0; i < 3; i++) {
          const int q = ky * 3 + i;  // 16-byte chunk of the 336-byte operand row
          sts128u(arow + (q >> 3) * 16384 + (((q & 7) ^ (r & 7)) << 4), make_uint4(o[4 * i], o[4 * i + 1], o[4 * i + 2], o[4 * i + 3]));
        }
      }
      tc::fence_proxy_async_smem();
      __syncwarp();
      if (lane == 0) {
        tc::mbar_arrive(&a_full[s]);
        tc::mbar_arrive(&r_empty[s]);  // every loaded word has been consumed by the conversions above
      }
    }
  } else if (warp >= 8) {
    // ===================== epilogue: group g takes the tiles of local parity g (accumulators g and g + 2); thread = token row
    const int ew = warp - 8, g = ew >> 2, quad = warp & 3;
    uint8_t* stage_p = sS + ew * 2 * PE_STAGE_BYTES;
    const uint32_t stage = tc::smem_u32(stage_p);
    const bool ln = a.A16 != nullptr;
    constexpr float inv_e = 1.0f / (float)E;
    uint32_t nbuf = 0;  // staging boxes used so far by this warp (buffer = nbuf & 1)
    int it = 0;
    // Positional-embedding rows: loaded with lanes ALONG the columns (8 lanes per 128-byte row piece, 4 rows per instruction: 4
    // L1 wavefronts instead of the 32 of a row-per-thread load), one chunk ahead in registers, and passed to the row-owning
    // thread through the staging buffer that then carries the chunk's result
    const int prow_l = lane >> 3, pcol_l = lane & 7;
    for (int t = blockIdx.x; t < n_tiles; t += gridDim.x, it++) {
      if ((it & 1) != g) continue;
      const int acc = it & (PE_NACC - 1);
      const uint32_t tacc = tmem_base + ((uint32_t)(quad * 32) << 16) + acc * E;
      const int b = t / tiles_per_img, rem = t - b * tiles_per_img;
      const int y = rem / tiles_per_row, x0 = (rem - y * tiles_per_row) * 128;
      const float* pbase = a.pos + (size_t)(rem * 128 + quad * 32 + prow_l) * E + pcol_l * 4;
      float4 pr[8];
#pragma unroll
      for (int i = 0; i < 8; i++)
        pr[i] = (pcol_l * 4 < E) ? __ldg((const float4*)(pbase + (size_t)(4 * i) * E)) : make_float4(0.f, 0.f, 0.f, 0.f);
      tc::mbar_wait(&t_full[acc], (it / PE_NACC) & 1);
      tc::tc_fence_after();
      float csum[NCH], cm2[NCH];
#pragma unroll
      for (int c = 0; c < NCH; c++) {
        const int ncols = (E - c * 32) >= 32 ? 32 : 16;
        uint32_t v[32];
        tc::tmem_ld_32x32(tacc + c * 32, v);
        const int bsel = nbuf & 1;
        const uint32_t buf = stage + bsel * PE_STAGE_BYTES;
        nbuf++;
        if (lane == 0) tc::tma_store_wait_read<1>();
        __syncwarp();
#pragma unroll
        for (int i = 0; i < 8; i++) {
          const int rr = prow_l + 4 * i;
          sts128(buf + rr * 128 + ((pcol_l ^ (rr & 7)) << 4), pr[i]);
        }
        __syncwarp();
        tc::tmem_ld_wait();
        float s1 = 0.f;
#pragma unroll
        for (int j = 0; j < 8; j++) {
          const uint32_t sp = buf + lane * 128 + ((j ^ (lane & 7)) << 4);
          float4 o = lds128(sp);
          o.x += __uint_as_float(v[4 * j + 0]);
          o.y += __uint_as_float(v[4 * j + 1]);
          o.z += __uint_as_float(v[4 * j + 2]);
          o.w += __uint_as_float(v[4 * j + 3]);
          if (4 * j < ncols) s1 += (o.x + o.y) + (o.z + o.w);
          v[4 * j + 0] = __float_as_uint(o.x); v[4 * j + 1] = __float_as_uint(o.y);
          v[4 * j + 2] = __float_as_uint(o.z); v[4 * j + 3] = __float_as_uint(o.w);
          sts128(sp, o);
        }
        tc::fence_proxy_async_smem();
        __syncwarp();
        if (lane == 0) {
          tc::tma_store_2d(&tmap_x, stage_p + bsel * PE_STAGE_BYTES, c * 32, (int)((long long)t * 128 + quad * 32));
          tc::tma_store_commit();
        }
        // the next chunk's positional rows are requested only now: the proxy fence above is a MEMBAR that waits for every
        // outstanding load of the thread, so loads issued before it would be waited for there instead of overlapping
        if (c + 1 < NCH) {
#pragma unroll
          for (int i = 0; i < 8; i++)
            pr[i] = ((c + 1) * 32 + pcol_l * 4 < E) ? __ldg((const float4*)(pbase + (size_t)(4 * i) * E + (c + 1) * 32))
                                                    : make_float4(0.f, 0.f, 0.f, 0.f);
        }
        if (ln) {
          if (ncols >= 32) tc::tmem_st_32x32(tacc + c * 32, v);
          else tc::tmem_st_32x16(tacc + c * 32, *(const uint32_t(*)[16]) & v[0]);
          // centred sum of squares of this chunk (values still in registers)
          const float mc = s1 * (1.0f / (float)ncols);
          float m2 = 0.f;
#pragma unroll
          for (int i = 0; i < 32; i++)
            if (i < ncols) {
              const float d = __uint_as_float(v[i]) - mc;
              m2 = fmaf(d, d, m2);
            }
          csum[c] = s1;
          cm2[c] = m2;
        }
      }
      if (ln) {
        float sum = 0.f;
#pragma unroll
        for (int c = 0; c < NCH; c++) sum += csum[c];
        const float mean = sum * inv_e;
        float q = 0.f;
#pragma unroll
        for (int c = 0; c < NCH; c++) {
          const float nc = (E - c * 32) >= 32 ? 32.f : 16.f;
          const float d = csum[c] * (1.0f / nc) - mean;
          q += cm2[c] + nc * d * d;
        }
        const float rstd = rsqrtf(q * inv_e + a.eps);
        const float nmr = -mean * rstd;
        // destination rows: tokens (y, x0 + quad * 32 + 8 k .. + 7) of image b are 8 consecutive rows in 8 x 8 window-major order
        const int nw = G >> 3;
        const long long drow0 = (((long long)b * nw + (y >> 3)) * nw + ((x0 + quad * 32) >> 3)) * 64 + (y & 7) * 8;
        tc::tmem_st_wait();
#pragma unroll
        for (int c = 0; c < NCH; c++) {
          const int ncols = (E - c * 32) >= 32 ? 32 : 16;
          uint32_t v[32];
          tc::tmem_ld_32x32(tacc + c * 32, v);
          // 16-bit box: 64-byte rows, 64B swizzle (chunk j of row r lives at chunk j ^ ((r >> 1) & 3)), four 8-row TMA stores
          const int bsel = nbuf & 1;
          const uint32_t sb = stage + bsel * PE_STAGE_BYTES;
          nbuf++;
          if (lane == 0) tc::tma_store_wait_read<1>();
          __syncwarp();
          tc::tmem_ld_wait();
#pragma unroll
          for (int j = 0; j < 4; j++) {
            uint32_t w[4];
#pragma unroll
            for (int k = 0; k < 2; k++) {
              const bool okc = 4 * (2 * j + k) < ncols;
              const float4 gg = okc ? __ldg((const float4*)(a.gamma + c * 32) + 2 * j + k) : make_float4(0.f, 0.f, 0.f, 0.f);
              const float4 bb = okc ? __ldg((const float4*)(a.beta + c * 32) + 2 * j + k) : make_float4(0.f, 0.f, 0.f, 0.f);
              // (x - mean) * rstd * gamma + beta  as  fma(fma(x, rstd, -mean * rstd), gamma, beta)
              const float o0 = fmaf(fmaf(__uint_as_float(v[8 * j + 4 * k + 0]), rstd, nmr), gg.x, bb.x);
              const float o1 = fmaf(fmaf(__uint_as_float(v[8 * j + 4 * k + 1]), rstd, nmr), gg.y, bb.y);
              const float o2 = fmaf(fmaf(__uint_as_float(v[8 * j + 4 * k + 2]), rstd, nmr), gg.z, bb.z);
              const float o3 = fmaf(fmaf(__uint_as_float(v[8 * j + 4 * k + 3]), rstd, nmr), gg.w, bb.w);
              w[2 * k] = tc::pack16(FP16 ? 1 : 0, o0, o1);
              w[2 * k + 1] = tc::pack16(FP16 ? 1 : 0, o2, o3);
            }
            sts128u(sb + lane * 64 + ((j ^ ((lane >> 1) & 3)) << 4), make_uint4(w[0], w[1], w[2], w[3]));
          }
          tc::fence_proxy_async_smem();
          __syncwarp();
          if (lane == 0) {
#pragma unroll
            for (int k = 0; k < 4; k++)
              tc::tma_store_2d(&tmap_a16, stage_p + bsel * PE_STAGE_BYTES + k * 512, c * 32, (int)(drow0 + k * 64));
            tc::tma_store_commit();
          }
        }
      }
      tc::tc_fence_before();
      __syncwarp();
      if (lane == 0) tc::mbar_arrive(&t_empty[acc]);
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
  return on && (E == 96 || E == 112) && S > 0 && (S % 512) == 0 && pe_smem_bytes(E) <= 232448;
}

template <int E, bool SWAP, bool FP16>
static int pe_launch_t(const CUtensorMap& tw, const CUtensorMap& tx, const CUtensorMap& ta, const PatchEmbedArgs& a, int n_tiles,
                       int grid, cudaStream_t st) {
  static std::atomic<unsigned long long> attr_set{0};
  auto kern = k_patch_embed<E, SWAP, FP16>;
  if (cvb_once_per_device(attr_set)) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, pe_smem_bytes(E));
    if (e != cudaSuccess) return cvb_fail_cuda(e, "cudaFuncSetAttribute(k_patch_embed)");
  }
  CVB_LAUNCH(kern, dim3(grid), dim3(PE_THREADS), pe_smem_bytes(E), st, tw, tx, ta, a, n_tiles);
  return CV_OK;
}

template <int E>
static int pe_launch_e(const CUtensorMap& tw, const CUtensorMap& tx, const CUtensorMap& ta, const PatchEmbedArgs& a, int n_tiles,
                       int grid, cudaStream_t st) {
  if (a.swap_rb)
    return a.fp16 ? pe_launch_t<E, true, true>(tw, tx, ta, a, n_tiles, grid, st) : pe_launch_t<E, true, false>(tw, tx, ta, a, n_tiles, grid, st);
  return a.fp16 ? pe_launch_t<E, false, true>(tw, tx, ta, a, n_tiles, grid, st) : pe_launch_t<E, false, false>(tw, tx, ta, a, n_tiles, grid, st);
}

int patch_embed_launch(const PatchEmbedArgs& a, int num_sms, cudaStream_t st) {
  if (!patch_embed_supported(a.E, a.S)) return cvb_fail(CV_ERR_INVALID, "patch_embed: unsupported width / image size");
  if (((uintptr_t)a.img & 15) || ((uintptr_t)a.W & 15) || ((uintptr_t)a.pos & 15) || ((uintptr_t)a.X0 & 15))
    return cvb_fail(CV_ERR_INVALID, "patch_embed: misaligned pointer");
  if (a.A16 && (!a.gamma || !a.beta || ((uintptr_t)a.A16 & 15) || ((uintptr_t)a.gamma & 15) || ((uintptr_t)a.beta & 15)))
    return cvb_fail(CV_ERR_INVALID, "patch_embed: LayerNorm output needs aligned gamma / beta / A16");
  const int G = a.S / 4;
  const long long rows = (long long)a.B * G * G;
  const int n_tiles = (int)(rows / 128);
  CUtensorMap tw, tx, ta;
  if (!tc_host::make_tmap_bf16(&tw, a.W, (uint64_t)a.E, (uint64_t)PEK, (uint64_t)PEK, (uint32_t)a.E) ||
      !tc_host::make_tmap_2d(&tx, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, a.X0, (uint64_t)rows, (uint64_t)a.E, (uint64_t)a.E * 4, 32, 32,
                             CU_TENSOR_MAP_SWIZZLE_128B))
    return cvb_fail(CV_ERR_CUDA, "cuTensorMapEncodeTiled failed (patch embed)");
  ta = tx;
  if (a.A16 && !tc_host::make_tmap_2d(&ta, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, a.A16, (uint64_t)rows, (uint64_t)a.E, (uint64_t)a.E * 2, 32,
                                      8, CU_TENSOR_MAP_SWIZZLE_64B))
    return cvb_fail(CV_ERR_CUDA, "cuTensorMapEncodeTiled failed (patch embed, norm1 output)");
  cvb_next_work(2.0 * (double)rows * a.E * 147);
  if (cvb_profile_on()) {
    char nm[96];
    snprintf(nm, sizeof(nm), "patch_embed M%lld E%d%s", rows, a.E, a.A16 ? " +LN1" : "");
    cvb_next_name(nm);
  }
  const int grid = n_tiles < num_sms ? n_tiles : num_sms;
  if (a.E == 96) return pe_launch_e<96>(tw, tx, ta, a, n_tiles, grid, st);
  return pe_launch_e<112>(tw, tx, ta, a, n_tiles, grid, st);
}

}  // namespace cvb
