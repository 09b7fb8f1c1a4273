// Definitions of the CUDA-core kernels declared in sam2_kernels.cuh.  Reference behaviour: SURVEY.md §B.2/B.3
// (sam2 package modules as called from /root/reference/src/sam2_infer.py:220-275) and sam2_infer.py:29-189.
#include "sam2_kernels.cuh"

#include <stdlib.h>

#include <math.h>

#include "common.cuh"
#include "tc05.cuh"
#include "act.cuh"

namespace cvb {

__device__ __forceinline__ float gelu_exact(float x) { return gelu_fast(x); }
__device__ __forceinline__ float warp_sum(float v) {
  for (int o = 16; o; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float warp_max(float v) {
  for (int o = 16; o; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

// ------------------------------------------------------------------------------------------------ im2col
// One thread per (token, tap-row): writes the 7 taps x 3 channels of kernel row ky.  K index = c*49 + ky*7 + kx
// (nn.Conv2d weight [Cout, 3, 7, 7] flattened), hi part at [0,PE_K), lo part at [PE_K, 2*PE_K).
__device__ __forceinline__ __nv_bfloat16 to16(int fp16, float v) {
  uint32_t w = tc::pack16(fp16, v, 0.f);
  unsigned short lo = (unsigned short)(w & 0xFFFFu);
  return *(__nv_bfloat16*)&lo;
}
__device__ __forceinline__ float from16(int fp16, __nv_bfloat16 h) {
  return fp16 ? __half2float(*(__half*)&h) : __bfloat162float(h);
}

template <bool U8>
__global__ void __launch_bounds__(256) k_im2col(const void* __restrict__ img_, int B, int S, float m0, float m1, float m2,
                                                float s0, float s1, float s2, int swap_rb, int fp16,
                                                __nv_bfloat16* __restrict__ out) {
  const int G = S / 4;
  long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  long long total = (long long)B * G * G * 8;  // 7 tap rows + 1 padding writer
  if (idx >= total) return;
  int ky = (int)(idx & 7);
  long long tok = idx >> 3;
  int x = (int)(tok % G);
  int y = (int)((tok / G) % G);
  int b = (int)(tok / ((long long)G * G));
  __nv_bfloat16* o = out + tok * (2 * PE_K);
  if (ky == 7) {  // zero the K padding 147..151 of both halves
    for (int k = 147; k < PE_K; k++) { o[k] = to16(fp16, 0.f); o[PE_K + k] = to16(fp16, 0.f); }
    return;
  }
  int iy = y * 4 - 3 + ky;
  const float mean[3] = {m0, m1, m2}, istd[3] = {s0, s1, s2};
  for (int c = 0; c < 3; c++) {
    for (int kx = 0; kx < 7; kx++) {
      int ix = x * 4 - 3 + kx;
      float v = 0.f;
      if (iy >= 0 && iy < S && ix >= 0 && ix < S) {
        if (U8) {
          const uint8_t* im = (const uint8_t*)img_;
          int cc = swap_rb ? 2 - c : c;
          // ToTensor (/255) then Normalize: (p/255 - mean) / std   (sam2_infer.py:41-46)
          v = ((float)im[(((long long)b * S + iy) * S + ix) * 3 + cc] / 255.0f - mean[c]) * istd[c];
        } else {
          const float* im = (const float*)img_;
          v = im[(((long long)b * 3 + c) * S + iy) * S + ix];
        }
      }
      __nv_bfloat16 hi = to16(fp16, v);
      __nv_bfloat16 lo = to16(fp16, v - from16(fp16, hi));
      int k = c * 49 + ky * 7 + kx;
      o[k] = hi;
      o[PE_K + k] = lo;
    }
  }
}

int launch_im2col_u8(const uint8_t* img, int B, int S, const float* mean, const float* inv_std, int swap_rb,
                     __nv_bfloat16* out, cudaStream_t st) {
  long long total = (long long)B * (S / 4) * (S / 4) * 8;
  cvb_next_work((double)B * S * S * 3 + (double)B * (S / 4) * (S / 4) * 2 * PE_K * 2);
  CVB_LAUNCH((k_im2col<true>), dim3((unsigned)((total + 255) / 256)), dim3(256), 0, st, img, B, S, mean[0], mean[1],
             mean[2], inv_std[0], inv_std[1], inv_std[2], swap_rb, 0, out);
  return CV_OK;
}

// Raw-pixel patch operand: out[tok, k] = fp16/bf16(pixel value) (0..255 is exact in both), zero outside the image and in
// the K padding.  ToTensor's 1/255 and Normalize's mean/std are folded into the GEMM weights and the per-token additive
// table at load time (sam2_weights.fold_state_dict: "pe.w8", "pos8").
// K order of the u8 path: k = ky*24 + kx*3 + c (21 taps of kernel row ky + 3 zeros): the 21 taps are 21 CONTIGUOUS bytes
// of the HWC image row 4y-3+ky starting at byte 12x-9, which is byte 3 of an aligned 32-bit word for every token, and the
// 24 converted values are three aligned 16-byte stores.  Thread = (token, ky); consecutive threads write consecutive
// 48-byte pieces of the operand (row pitch 336 B), so the 22 MB/image write stream is fully coalesced.
// (The previous version — 152-wide rows built element by element through shared memory — was issue-bound at 90 % issue
// utilisation and 1.1 TB/s: ~30 instructions per element for the index arithmetic; this one needs ~3.)
template <bool SWAP, bool FP16>
__global__ void __launch_bounds__(224) k_im2col_u8raw(const uint8_t* __restrict__ img, int S, __nv_bfloat16* __restrict__ out) {
  const int G = S / 4;
  const int t = blockIdx.x * 224 + threadIdx.x;  // (token in row, ky), ky fastest
  const int x = t / 7, ky = t - x * 7;
  const int y = blockIdx.y, b = blockIdx.z;
  if (x >= G) return;
  const int iy = y * 4 - 3 + ky;
  uint32_t w[6];
#pragma unroll
  for (int i = 0; i < 6; i++) w[i] = 0u;
  if (iy >= 0 && iy < S) {
    const uint32_t* row = (const uint32_t*)(img + ((size_t)b * S + iy) * S * 3);
    const int w0 = 3 * x - 3;  // word that holds byte 12x - 9 in its top byte
#pragma unroll
    for (int i = 0; i < 6; i++)
      if (w0 + i >= 0) w[i] = __ldg(row + w0 + i);  // only token 0 has words left of the row (pixels -3..-1)
  }
  // byte j of the 21-byte segment -> fp32 by the 2^23 trick (PRMT builds 0x4B0000bb), exact for 0..255
  auto val = [&](int j) -> float {
    if (j >= 21) return 0.f;
    if (SWAP) j = 3 * (j / 3) + 2 - (j % 3);  // BGR input: channel c of pixel kx lives at byte 2 - c
    const int byte = j + 3;                    // position inside the 24 loaded bytes
    return __uint_as_float(__byte_perm(w[byte >> 2], 0x4B000000u, 0x7540 + (byte & 3))) - 8388608.0f;
  };
  uint32_t o[12];
#pragma unroll
  for (int i = 0; i < 12; i++) o[i] = tc::pack16(FP16 ? 1 : 0, val(2 * i), val(2 * i + 1));
  uint4* dst = (uint4*)(out + (((size_t)b * G + y) * G + x) * PE_K8 + ky * 24);
  dst[0] = make_uint4(o[0], o[1], o[2], o[3]);
  dst[1] = make_uint4(o[4], o[5], o[6], o[7]);
  dst[2] = make_uint4(o[8], o[9], o[10], o[11]);
}

int launch_im2col_u8raw(const uint8_t* img, int B, int S, int swap_rb, int fp16, __nv_bfloat16* out, cudaStream_t st) {
  if (S % 4 || ((uintptr_t)img & 3)) return cvb_fail(CV_ERR_INVALID, "im2col: image side must be a multiple of 4 and 4-byte aligned");
  const int G = S / 4;
  cvb_next_work((double)B * S * S * 3 + (double)B * G * G * PE_K8 * 2);
  dim3 grid((G * 7 + 223) / 224, G, B), blk(224);
  if (swap_rb) {
    if (fp16) { CVB_LAUNCH((k_im2col_u8raw<true, true>), grid, blk, 0, st, img, S, out); }
    else { CVB_LAUNCH((k_im2col_u8raw<true, false>), grid, blk, 0, st, img, S, out); }
  } else {
    if (fp16) { CVB_LAUNCH((k_im2col_u8raw<false, true>), grid, blk, 0, st, img, S, out); }
    else { CVB_LAUNCH((k_im2col_u8raw<false, false>), grid, blk, 0, st, img, S, out); }
  }
  return CV_OK;
}

int launch_im2col_f32(const float* img, int B, int S, int fp16, __nv_bfloat16* out, cudaStream_t st) {
  long long total = (long long)B * (S / 4) * (S / 4) * 8;
  cvb_next_work((double)B * S * S * 12 + (double)B * (S / 4) * (S / 4) * 2 * PE_K * 2);
  CVB_LAUNCH((k_im2col<false>), dim3((unsigned)((total + 255) / 256)), dim3(256), 0, st, img, B, S, 0.f, 0.f, 0.f, 1.f,
             1.f, 1.f, 0, fp16, out);
  return CV_OK;
}

// ------------------------------------------------------------------------------------------------ antialias resize
struct AATap { int lo, n; float center, invscale; };
__device__ __forceinline__ AATap aa_tap(int i, float scale, int in_size) {
  // upsample_bilinear2d_aa: support = max(scale, 1), taps [center - support + 0.5, center + support + 0.5)
  float support = scale >= 1.f ? scale : 1.f;
  float center = scale * (i + 0.5f);
  AATap t;
  t.invscale = scale >= 1.f ? 1.f / scale : 1.f;
  t.lo = max((int)(center - support + 0.5f), 0);
  t.n = min((int)(center + support + 0.5f), in_size) - t.lo;
  t.center = center;
  return t;
}
__device__ __forceinline__ float aa_w(const AATap& t, int j) {
  float x = fabsf((j + t.lo - t.center + 0.5f) * t.invscale);
  return x < 1.f ? 1.f - x : 0.f;
}

__global__ void __launch_bounds__(256) k_aa_width(const uint8_t* __restrict__ img, int H, int W, int S, float scale, int swap_rb,
                                                  float* __restrict__ tmp) {
  int x = blockIdx.x * 64 + (threadIdx.x & 63), y = blockIdx.y * 4 + (threadIdx.x >> 6);
  if (x >= S || y >= H) return;
  AATap t = aa_tap(x, scale, W);
  float tot = 0.f, a0 = 0.f, a1 = 0.f, a2 = 0.f;
  for (int j = 0; j < t.n; j++) tot += aa_w(t, j);
  for (int j = 0; j < t.n; j++) {
    float w = aa_w(t, j) / tot;
    const uint8_t* p = img + ((long long)y * W + t.lo + j) * 3;
    a0 += w * ((float)p[swap_rb ? 2 : 0] / 255.0f);
    a1 += w * ((float)p[1] / 255.0f);
    a2 += w * ((float)p[swap_rb ? 0 : 2] / 255.0f);
  }
  float* o = tmp + ((long long)y * S + x) * 3;
  o[0] = a0; o[1] = a1; o[2] = a2;
}

__global__ void __launch_bounds__(256) k_aa_height(const float* __restrict__ tmp, int H, int S, float scale, float m0, float m1,
                                                   float m2, float s0, float s1, float s2, float* __restrict__ out) {
  int x = blockIdx.x * 64 + (threadIdx.x & 63), y = blockIdx.y * 4 + (threadIdx.x >> 6);
  if (x >= S || y >= S) return;
  AATap t = aa_tap(y, scale, H);
  float tot = 0.f, a0 = 0.f, a1 = 0.f, a2 = 0.f;
  for (int j = 0; j < t.n; j++) tot += aa_w(t, j);
  for (int j = 0; j < t.n; j++) {
    float w = aa_w(t, j) / tot;
    const float* p = tmp + ((long long)(t.lo + j) * S + x) * 3;
    a0 += w * p[0]; a1 += w * p[1]; a2 += w * p[2];
  }
  long long pl = (long long)S * S, o = (long long)y * S + x;
  out[o] = (a0 - m0) * s0;
  out[pl + o] = (a1 - m1) * s1;
  out[2 * pl + o] = (a2 - m2) * s2;
}

int launch_preprocess_aa(const uint8_t* img, int H, int W, int S, const float* mean, const float* inv_std, int swap_rb,
                         float* tmp, float* out, cudaStream_t st) {
  float sx = (float)W / (float)S, sy = (float)H / (float)S;
  CVB_LAUNCH(k_aa_width, dim3((S + 63) / 64, (H + 3) / 4), dim3(256), 0, st, img, H, W, S, sx, swap_rb, tmp);
  CVB_LAUNCH(k_aa_height, dim3((S + 63) / 64, (S + 3) / 4), dim3(256), 0, st, tmp, H, S, sy, mean[0], mean[1], mean[2],
             inv_std[0], inv_std[1], inv_std[2], out);
  return CV_OK;
}

// Batched form for whole pages (analysis_pipeline.py:177-208): image b is the crop window
// [x0, x1) x [y0, y1) of page b (uint8 HWC at pages + off[b], page width pw[b]) — the array the reference obtains by slicing
// (circuit_analyzer.py:1246) and hands to SAM2Transforms.  Same arithmetic as the single-image kernels above, one grid
// z-slice per image; tmp is [B][max_hc][S][3] floats.
struct PageGeom { long long off; int pw, x0, y0, x1, y1, pad; };

__global__ void __launch_bounds__(256) k_aa_width_pages(const uint8_t* __restrict__ pages, const PageGeom* __restrict__ geom, int S,
                                                        int swap_rb, int max_hc, float* __restrict__ tmp) {
  const PageGeom g = geom[blockIdx.z];
  const int Hc = g.y1 - g.y0, Wc = g.x1 - g.x0;
  int x = blockIdx.x * 64 + (threadIdx.x & 63), y = blockIdx.y * 4 + (threadIdx.x >> 6);
  if (x >= S || y >= Hc) return;
  const float scale = (float)Wc / (float)S;
  AATap t = aa_tap(x, scale, Wc);
  float tot = 0.f, a0 = 0.f, a1 = 0.f, a2 = 0.f;
  for (int j = 0; j < t.n; j++) tot += aa_w(t, j);
  const uint8_t* rowp = pages + g.off + ((long long)(g.y0 + y) * g.pw + g.x0) * 3;
  for (int j = 0; j < t.n; j++) {
    float w = aa_w(t, j) / tot;
    const uint8_t* p = rowp + (long long)(t.lo + j) * 3;
    a0 += w * ((float)p[swap_rb ? 2 : 0] / 255.0f);
    a1 += w * ((float)p[1] / 255.0f);
    a2 += w * ((float)p[swap_rb ? 0 : 2] / 255.0f);
  }
  float* o = tmp + (((long long)blockIdx.z * max_hc + y) * S + x) * 3;
  o[0] = a0; o[1] = a1; o[2] = a2;
}

__global__ void __launch_bounds__(256) k_aa_height_pages(const float* __restrict__ tmp, const PageGeom* __restrict__ geom, int S,
                                                         int max_hc, float m0, float m1, float m2, float s0, float s1, float s2,
                                                         float* __restrict__ out) {
  const PageGeom g = geom[blockIdx.z];
  const int Hc = g.y1 - g.y0;
  int x = blockIdx.x * 64 + (threadIdx.x & 63), y = blockIdx.y * 4 + (threadIdx.x >> 6);
  if (x >= S || y >= S) return;
  const float scale = (float)Hc / (float)S;
  AATap t = aa_tap(y, scale, Hc);
  float tot = 0.f, a0 = 0.f, a1 = 0.f, a2 = 0.f;
  for (int j = 0; j < t.n; j++) tot += aa_w(t, j);
  const float* base = tmp + (long long)blockIdx.z * max_hc * S * 3;
  for (int j = 0; j < t.n; j++) {
    float w = aa_w(t, j) / tot;
    const float* p = base + ((long long)(t.lo + j) * S + x) * 3;
    a0 += w * p[0]; a1 += w * p[1]; a2 += w * p[2];
  }
  long long pl = (long long)S * S, o = (long long)y * S + x;
  float* ob = out + (long long)blockIdx.z * 3 * pl;
  ob[o] = (a0 - m0) * s0;
  ob[pl + o] = (a1 - m1) * s1;
  ob[2 * pl + o] = (a2 - m2) * s2;
}

int launch_preprocess_pages(const uint8_t* pages, const void* geom, int B, int max_hc, int S, const float* mean,
                            const float* inv_std, int swap_rb, float* tmp, float* out, cudaStream_t st) {
  CVB_LAUNCH(k_aa_width_pages, dim3((S + 63) / 64, (max_hc + 3) / 4, B), dim3(256), 0, st, pages, (const PageGeom*)geom, S, swap_rb,
             max_hc, tmp);
  CVB_LAUNCH(k_aa_height_pages, dim3((S + 63) / 64, (S + 3) / 4, B), dim3(256), 0, st, tmp, (const PageGeom*)geom, S, max_hc, mean[0],
             mean[1], mean[2], inv_std[0], inv_std[1], inv_std[2], out);
  return CV_OK;
}

// ------------------------------------------------------------------------------------------------ LayerNorm rows
// One warp per destination row.  C % 4 == 0, C <= 1536.
__global__ void __launch_bounds__(256) k_ln_rows(const float* __restrict__ X, int C, const float* __restrict__ gamma,
                                                 const float* __restrict__ beta, float eps, int H, int W, int ws,
                                                 int nwx, int nwy, long long n_dst, int fp16, __nv_bfloat16* __restrict__ ob,
                                                 float* __restrict__ of) {
  long long r = (long long)blockIdx.x * 8 + (threadIdx.x >> 5);
  if (r >= n_dst) return;
  int lane = threadIdx.x & 31;
  long long src = r;
  if (ws > 0) {
    int w2 = ws * ws;
    long long win = r / w2;
    int t = (int)(r - win * w2);
    int per = nwx * nwy;
    long long b = win / per;
    int wi = (int)(win - b * per);
    int wy = wi / nwx, wx = wi - wy * nwx;
    int ty = t / ws, tx = t - ty * ws;
    int y = wy * ws + ty, x = wx * ws + tx;
    src = (y < H && x < W) ? (b * H + y) * W + x : -1;
  }
  const int nv = C >> 2;  // float4 per row
  if (src < 0) {
    for (int i = lane; i < nv; i += 32) {
      if (ob) *(uint2*)(ob + r * C + i * 4) = make_uint2(0u, 0u);
      if (of) *(float4*)(of + r * C + i * 4) = make_float4(0.f, 0.f, 0.f, 0.f);
    }
    return;
  }
  float4 v[12];
  const float4* xr = (const float4*)(X + src * C);
  float s = 0.f;
#pragma unroll
  for (int k = 0; k < 12; k++) {
    int i = lane + k * 32;
    if (i < nv) {
      v[k] = xr[i];
      s += v[k].x + v[k].y + v[k].z + v[k].w;
    }
  }
  float mean = 0.f, rstd = 1.f;
  if (gamma) {
    mean = warp_sum(s) / C;
    float q = 0.f;
#pragma unroll
    for (int k = 0; k < 12; k++) {
      int i = lane + k * 32;
      if (i < nv) {
        float a = v[k].x - mean, b = v[k].y - mean, c = v[k].z - mean, d = v[k].w - mean;
        q += a * a + b * b + c * c + d * d;
      }
    }
    rstd = rsqrtf(warp_sum(q) / C + eps);
  }
#pragma unroll
  for (int k = 0; k < 12; k++) {
    int i = lane + k * 32;
    if (i < nv) {
      float4 o = v[k];
      if (gamma) {
        float4 g = ((const float4*)gamma)[i], bb = ((const float4*)beta)[i];
        o.x = (o.x - mean) * rstd * g.x + bb.x;
        o.y = (o.y - mean) * rstd * g.y + bb.y;
        o.z = (o.z - mean) * rstd * g.z + bb.z;
        o.w = (o.w - mean) * rstd * g.w + bb.w;
      }
      if (of) *(float4*)(of + r * C + i * 4) = o;
      if (ob) {
        *(uint2*)(ob + r * C + i * 4) = make_uint2(tc::pack16(fp16, o.x, o.y), tc::pack16(fp16, o.z, o.w));
      }
    }
  }
}

// L lanes per row (32/L rows per warp in flight), V float4 per lane: C = 4*L*V.  Narrow rows (C = 96: 384 bytes) keep
// four rows per warp in flight instead of one, which is what a latency-bound streaming kernel needs.
template <int L, int V>
__global__ void __launch_bounds__(256) k_ln_rows_t(const float* __restrict__ X, const float* __restrict__ gamma,
                                                   const float* __restrict__ beta, float eps, int H, int W, int ws, int nwx,
                                                   int nwy, long long n_dst, int fp16, __nv_bfloat16* __restrict__ ob,
                                                   float* __restrict__ of, __nv_bfloat16* __restrict__ raw16) {
  constexpr int C = 4 * L * V, RPW = 32 / L;
  const int lane = threadIdx.x & 31, sub = lane % L;
  const long long r = ((long long)blockIdx.x * 8 + (threadIdx.x >> 5)) * RPW + lane / L;
  const bool live = r < n_dst;
  long long src = live ? r : -1;
  if (live && ws > 0) {
    int w2 = ws * ws;
    long long win = r / w2;
    int t = (int)(r - win * w2);
    int per = nwx * nwy;
    long long b = win / per;
    int wi = (int)(win - b * per);
    int wy = wi / nwx, wx = wi - wy * nwx;
    int ty = t / ws, tx = t - ty * ws;
    int y = wy * ws + ty, x = wx * ws + tx;
    src = (y < H && x < W) ? (b * H + y) * W + x : -1;
  }
  float4 v[V];
  float s = 0.f;
  if (src >= 0) {
    const float4* xr = (const float4*)(X + src * C);
#pragma unroll
    for (int k = 0; k < V; k++) {
      v[k] = xr[sub + k * L];
      s += v[k].x + v[k].y + v[k].z + v[k].w;
    }
  } else {
#pragma unroll
    for (int k = 0; k < V; k++) v[k] = make_float4(0.f, 0.f, 0.f, 0.f);
  }
  float mean = 0.f, rstd = 1.f;
  if (gamma) {
#pragma unroll
    for (int o = L / 2; o; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    mean = s * (1.0f / C);
    float q = 0.f;
#pragma unroll
    for (int k = 0; k < V; k++) {
      float a = v[k].x - mean, b = v[k].y - mean, c = v[k].z - mean, d = v[k].w - mean;
      q += a * a + b * b + c * c + d * d;
    }
#pragma unroll
    for (int o = L / 2; o; o >>= 1) q += __shfl_xor_sync(0xffffffffu, q, o);
    rstd = rsqrtf(q * (1.0f / C) + eps);
  }
  if (!live) return;
#pragma unroll
  for (int k = 0; k < V; k++) {
    const int i = sub + k * L;
    float4 o = v[k];
    // 16-bit copy of the un-normalised source row, in source order (the neck's lateral conv reads the stage output as a GEMM
    // operand: one read of the fp32 stream serves both this copy and norm1 of the next stage's first block)
    if (raw16 && src >= 0) *(uint2*)(raw16 + src * C + i * 4) = make_uint2(tc::pack16(fp16, o.x, o.y), tc::pack16(fp16, o.z, o.w));
    if (gamma && src >= 0) {
      float4 g = __ldg((const float4*)gamma + i), bb = __ldg((const float4*)beta + i);
      o.x = (o.x - mean) * rstd * g.x + bb.x;
      o.y = (o.y - mean) * rstd * g.y + bb.y;
      o.z = (o.z - mean) * rstd * g.z + bb.z;
      o.w = (o.w - mean) * rstd * g.w + bb.w;
    }
    if (of) *(float4*)(of + r * C + i * 4) = o;
    if (ob) {
      *(uint2*)(ob + r * C + i * 4) = make_uint2(tc::pack16(fp16, o.x, o.y), tc::pack16(fp16, o.z, o.w));
    }
  }
}

template <int L, int V>
static int launch_ln_t(const float* X, const float* gamma, const float* beta, float eps, int H, int W, int ws, int nwx, int nwy,
                       long long n_dst, int fp16, __nv_bfloat16* ob, float* of, __nv_bfloat16* raw16, cudaStream_t st) {
  const long long rows_per_block = 8 * (32 / L);
  CVB_LAUNCH((k_ln_rows_t<L, V>), dim3((unsigned)((n_dst + rows_per_block - 1) / rows_per_block)), dim3(256), 0, st, X, gamma,
             beta, eps, H, W, ws, nwx, nwy, n_dst, fp16, ob, of, raw16);
  return CV_OK;
}

int launch_ln_rows(const float* X, long long n_src_rows, int C, const float* gamma, const float* beta, float eps, int B,
                   int H, int W, int ws, int fp16, __nv_bfloat16* out_bf16, float* out_f32, cudaStream_t st, __nv_bfloat16* raw16) {
  if ((C & 3) || C > 1536) return cvb_fail(CV_ERR_INVALID, "ln_rows: C must be a multiple of 4 and <= 1536");
  int nwx = 0, nwy = 0;
  long long n_dst = n_src_rows;
  if (ws > 0) {
    nwx = (W + ws - 1) / ws;
    nwy = (H + ws - 1) / ws;
    n_dst = (long long)B * nwx * nwy * ws * ws;
  }
  cvb_next_work((double)n_src_rows * C * 4 + (double)n_dst * C * (out_bf16 ? 2 : 0) + (double)n_dst * C * (out_f32 ? 4 : 0) +
                (double)n_src_rows * C * (raw16 ? 2 : 0));
  if (cvb_profile_on()) {
    char nm[96];
    snprintf(nm, sizeof(nm), "ln_rows R%lld C%d ws%d%s%s", n_src_rows, C, ws, gamma ? "" : " cast", raw16 ? " +copy" : "");
    cvb_next_name(nm);
  }
#define CVB_LN_CASE(CC, LL, VV) \
  if (C == CC) return launch_ln_t<LL, VV>(X, gamma, beta, eps, H, W, ws, nwx, nwy, n_dst, fp16, out_bf16, out_f32, raw16, st);
  CVB_LN_CASE(96, 8, 3) CVB_LN_CASE(192, 16, 3) CVB_LN_CASE(384, 32, 3) CVB_LN_CASE(768, 32, 6) CVB_LN_CASE(256, 16, 4)
  CVB_LN_CASE(64, 8, 2) CVB_LN_CASE(112, 4, 7) CVB_LN_CASE(224, 8, 7) CVB_LN_CASE(448, 16, 7) CVB_LN_CASE(896, 32, 7)
  CVB_LN_CASE(144, 4, 9) CVB_LN_CASE(288, 8, 9) CVB_LN_CASE(576, 16, 9) CVB_LN_CASE(1152, 32, 9)
#undef CVB_LN_CASE
  if (raw16) return cvb_fail(CV_ERR_INVALID, "ln_rows: the 16-bit source copy needs one of the templated widths");
  CVB_LAUNCH(k_ln_rows, dim3((unsigned)((n_dst + 7) / 8)), dim3(256), 0, st, X, C, gamma, beta, eps, H, W, ws, nwx, nwy,
             n_dst, fp16, out_bf16, out_f32);
  return CV_OK;
}

// ------------------------------------------------------------------------------------------------ pooling
__global__ void __launch_bounds__(256) k_pool_q(const __nv_bfloat16* __restrict__ qkv, long long ld, int ws, int Cq, int fp16,
                                                long long n_dst, __nv_bfloat16* __restrict__ qp) {
  const int chunks = Cq >> 3;
  long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= n_dst * chunks) return;
  long long r = idx / chunks;
  int c = (int)(idx - r * chunks) * 8;
  int hw = ws >> 1, h2 = hw * hw;
  long long win = r / h2;
  int t = (int)(r - win * h2);
  int py = t / hw, px = t - py * hw;
  long long base = win * ws * ws;
  __nv_bfloat162 m[4];
  bool first = true;
  for (int dy = 0; dy < 2; dy++)
    for (int dx = 0; dx < 2; dx++) {
      long long sr = base + (2 * py + dy) * ws + 2 * px + dx;
      uint4 u = *(const uint4*)(qkv + sr * ld + c);
      const __nv_bfloat162* p = (const __nv_bfloat162*)&u;
      for (int k = 0; k < 4; k++) {
        if (first) m[k] = p[k];
        else if (fp16) { __half2 r = __hmax2(*(const __half2*)&m[k], *(const __half2*)&p[k]); m[k] = *(__nv_bfloat162*)&r; }
        else m[k] = __hmax2(m[k], p[k]);
      }
      first = false;
    }
  *(uint4*)(qp + r * Cq + c) = *(uint4*)m;
}

int launch_pool_q(const __nv_bfloat16* qkv, long long ld, int n_windows, int ws, int Cq, int fp16, __nv_bfloat16* qp,
                  cudaStream_t st) {
  if ((ws & 1) || (Cq & 7)) return cvb_fail(CV_ERR_INVALID, "pool_q: odd window or Cq%8 != 0");
  long long n_dst = (long long)n_windows * (ws / 2) * (ws / 2);
  long long total = n_dst * (Cq / 8);
  cvb_next_work((double)n_dst * Cq * 2 * 5);
  CVB_LAUNCH(k_pool_q, dim3((unsigned)((total + 255) / 256)), dim3(256), 0, st, qkv, ld, ws, Cq, fp16, n_dst, qp);
  return CV_OK;
}

__global__ void __launch_bounds__(256) k_pool_shortcut(const float* __restrict__ S, int H, int W, int ws, int nwx, int nwy,
                                                       int C, long long n_dst, float* __restrict__ R) {
  const int chunks = C >> 2;
  long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= n_dst * chunks) return;
  long long r = idx / chunks;
  int c = (int)(idx - r * chunks) * 4;
  int H2 = H >> 1, W2 = W >> 1;
  int x2 = (int)(r % W2);
  int y2 = (int)((r / W2) % H2);
  long long b = r / ((long long)W2 * H2);
  float4 m = make_float4(-INFINITY, -INFINITY, -INFINITY, -INFINITY);
  for (int dy = 0; dy < 2; dy++)
    for (int dx = 0; dx < 2; dx++) {
      int y = 2 * y2 + dy, x = 2 * x2 + dx;
      int wy = y / ws, ty = y - wy * ws, wx = x / ws, tx = x - wx * ws;
      long long sr = ((b * nwy + wy) * nwx + wx) * (long long)(ws * ws) + ty * ws + tx;
      float4 v = *(const float4*)(S + sr * C + c);
      m.x = fmaxf(m.x, v.x); m.y = fmaxf(m.y, v.y); m.z = fmaxf(m.z, v.z); m.w = fmaxf(m.w, v.w);
    }
  *(float4*)(R + r * C + c) = m;
}

int launch_pool_shortcut(const float* S, int B, int H, int W, int ws, int C, float* R, cudaStream_t st) {
  int nwx = (W + ws - 1) / ws, nwy = (H + ws - 1) / ws;
  long long n_dst = (long long)B * (H / 2) * (W / 2);
  long long total = n_dst * (C / 4);
  cvb_next_work((double)n_dst * C * 4 * 5);
  CVB_LAUNCH(k_pool_shortcut, dim3((unsigned)((total + 255) / 256)), dim3(256), 0, st, S, H, W, ws, nwx, nwy, C, n_dst, R);
  return CV_OK;
}

__global__ void __launch_bounds__(256) k_add_nearest2(float* __restrict__ Y, const float* __restrict__ L, int H, int W,
                                                      int C, long long total) {
  long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= total) return;
  const int chunks = C >> 2;
  long long r = idx / chunks;
  int c = (int)(idx - r * chunks) * 4;
  int x = (int)(r % W), y = (int)((r / W) % H);
  long long b = r / ((long long)W * H);
  long long lr = (b * (H / 2) + y / 2) * (W / 2) + x / 2;
  float4 a = *(float4*)(Y + r * C + c);
  float4 l = *(const float4*)(L + lr * C + c);
  a.x += l.x; a.y += l.y; a.z += l.z; a.w += l.w;
  *(float4*)(Y + r * C + c) = a;
}

int launch_add_nearest2(float* Y, const float* L, int B, int H, int W, int C, cudaStream_t st) {
  long long total = (long long)B * H * W * (C / 4);
  cvb_next_work((double)B * H * W * C * 4 * 2.25);
  CVB_LAUNCH(k_add_nearest2, dim3((unsigned)((total + 255) / 256)), dim3(256), 0, st, Y, L, H, W, C, total);
  return CV_OK;
}

// ------------------------------------------------------------------------------------------------ token-side linear (fp32)
// C[R,N] = act(A[R,K] * W[N,K]^T + bias) (+ res).  64 x 64 output tile per CTA, 16 x 16 threads, 4 x 4 micro-tile, K
// tiles of 32 double-buffered with cp.async: the K loop of these small GEMMs (R = 38 tokens per image) is
// latency-bound, so the next tile's global loads are in flight while the current one is multiplied.
constexpr int TL_BK = 32, TL_LD = 64 + 4;
__device__ __forceinline__ void cp_async4(float* smem_dst, const float* gsrc, bool ok) {
  unsigned d = (unsigned)__cvta_generic_to_shared(smem_dst);
  int sz = ok ? 4 : 0;  // src-size 0 -> zero fill
  asm volatile("cp.async.ca.shared.global [%0], [%1], 4, %2;" ::"r"(d), "l"(gsrc), "r"(sz) : "memory");
}
__global__ void __launch_bounds__(256) k_tok_linear(const float* __restrict__ A, long long lda, const float* __restrict__ Wt,
                                                    const float* __restrict__ bias, int R, int N, int K, int act,
                                                    const float* __restrict__ res, long long ld_res, float* __restrict__ Cm,
                                                    long long ldc) {
  __shared__ __align__(16) float sA[2][TL_BK][TL_LD];
  __shared__ __align__(16) float sW[2][TL_BK][TL_LD];
  const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
  const int r0 = blockIdx.y * 64, n0 = blockIdx.x * 64;
  const int n_kt = (K + TL_BK - 1) / TL_BK;
  auto stage = [&](int kt, int buf) {
    const int k0 = kt * TL_BK;
    // 64 rows x 32 k per operand; consecutive threads take consecutive k of one row (coalesced 128-byte rows)
#pragma unroll
    for (int i = 0; i < 8; i++) {
      const int e = threadIdx.x + i * 256;
      const int rr = e >> 5, kk = e & 31;
      const int gk = k0 + kk;
      const int gr = r0 + rr, gn = n0 + rr;
      const bool oka = gr < R && gk < K, okw = gn < N && gk < K;
      cp_async4(&sA[buf][kk][rr], A + (oka ? (long long)gr * lda + gk : 0), oka);
      cp_async4(&sW[buf][kk][rr], Wt + (okw ? (long long)gn * K + gk : 0), okw);
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
  };
  float acc[4][4] = {};
  stage(0, 0);
  for (int kt = 0; kt < n_kt; kt++) {
    const int buf = kt & 1;
    if (kt + 1 < n_kt) {
      stage(kt + 1, buf ^ 1);
      asm volatile("cp.async.wait_group 1;" ::: "memory");
    } else {
      asm volatile("cp.async.wait_group 0;" ::: "memory");
    }
    __syncthreads();
#pragma unroll
    for (int kk = 0; kk < TL_BK; kk++) {
      const float4 a4 = *(const float4*)&sA[buf][kk][ty * 4];
      const float4 w4 = *(const float4*)&sW[buf][kk][tx * 4];
      const float a[4] = {a4.x, a4.y, a4.z, a4.w}, w[4] = {w4.x, w4.y, w4.z, w4.w};
#pragma unroll
      for (int i = 0; i < 4; i++)
#pragma unroll
        for (int j = 0; j < 4; j++) acc[i][j] = fmaf(a[i], w[j], acc[i][j]);
    }
    __syncthreads();
  }
  for (int i = 0; i < 4; i++) {
    int gr = r0 + ty * 4 + i;
    if (gr >= R) continue;
    for (int j = 0; j < 4; j++) {
      int gn = n0 + tx * 4 + j;
      if (gn >= N) continue;
      float x = acc[i][j] + (bias ? bias[gn] : 0.f);
      if (act == 1) x = gelu_exact(x);
      else if (act == 2) x = fmaxf(x, 0.f);
      else if (act == 3) x = 1.f / (1.f + expf(-x));
      if (res) x += res[(long long)gr * ld_res + gn];
      Cm[(long long)gr * ldc + gn] = x;
    }
  }
}

int launch_tok_linear(const float* A, long long lda, const float* W, const float* bias, int R, int N, int K, int act,
                      const float* res, long long ld_res, float* C, long long ldc, cudaStream_t st) {
  cvb_next_work(2.0 * R * (double)N * K);
  if (cvb_profile_on()) {
    char nm[64];
    snprintf(nm, sizeof(nm), "tok_linear R%d N%d K%d act%d", R, N, K, act);
    cvb_next_name(nm);
  }
  CVB_LAUNCH(k_tok_linear, dim3((N + 63) / 64, (R + 63) / 64), dim3(256), 0, st, A, lda, W, bias, R, N, K, act, res, ld_res,
             C, ldc);
  return CV_OK;
}

// Y = LN(X + add) * g + b ; one warp per row, C % 32 == 0, C <= 512
__global__ void __launch_bounds__(256) k_tok_add_ln(const float* __restrict__ X, const float* __restrict__ add, int mod,
                                                    const float* __restrict__ g, const float* __restrict__ b, float eps,
                                                    int R, int C, float* __restrict__ Y) {
  int r = blockIdx.x * 8 + (threadIdx.x >> 5);
  if (r >= R) return;
  int lane = threadIdx.x & 31;
  float v[16];
  int n = C >> 5;
  float s = 0.f;
  for (int k = 0; k < n; k++) {
    int c = lane + k * 32;
    float x = X[(long long)r * C + c];
    if (add) x += add[(long long)(mod > 0 ? r % mod : r) * C + c];
    v[k] = x;
    s += x;
  }
  float mean = warp_sum(s) / C, q = 0.f;
  for (int k = 0; k < n; k++) { float d = v[k] - mean; q += d * d; }
  float rstd = rsqrtf(warp_sum(q) / C + eps);
  for (int k = 0; k < n; k++) {
    int c = lane + k * 32;
    Y[(long long)r * C + c] = (v[k] - mean) * rstd * g[c] + b[c];
  }
}

int launch_tok_add_ln(const float* X, const float* add, int add_rows_mod, const float* g, const float* b, float eps, int R,
                      int C, float* Y, cudaStream_t st) {
  if ((C & 31) || C > 512) return cvb_fail(CV_ERR_INVALID, "tok_add_ln: C%32 / C<=512");
  CVB_LAUNCH(k_tok_add_ln, dim3((R + 7) / 8), dim3(256), 0, st, X, add, add_rows_mod, g, b, eps, R, C, Y);
  return CV_OK;
}

__global__ void k_tok_add_bcast(const float* __restrict__ X, const float* __restrict__ P, int mod, long long total, int C,
                                float* __restrict__ Y) {
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total) return;
  long long r = i / C;
  int c = (int)(i - r * C);
  Y[i] = X[i] + P[(r % mod) * C + c];
}

int launch_tok_add_bcast(const float* X, const float* P, int mod, int R, int C, float* Y, cudaStream_t st) {
  long long total = (long long)R * C;
  CVB_LAUNCH(k_tok_add_bcast, dim3((unsigned)((total + 255) / 256)), dim3(256), 0, st, X, P, mod, total, C, Y);
  return CV_OK;
}

// self attention among T (<= 64) tokens: one CTA per (image, head), 64 threads, thread = query
__global__ void __launch_bounds__(64) k_tok_self_attn(const float* __restrict__ q, const float* __restrict__ k,
                                                      const float* __restrict__ v, int T, int heads, int d,
                                                      float* __restrict__ out) {
  __shared__ float sk[64][33], sv[64][33];
  int b = blockIdx.x, h = blockIdx.y, t = threadIdx.x;
  int C = heads * d;
  for (int i = threadIdx.x; i < T * d; i += 64) {
    int tt = i / d, dd = i - tt * d;
    sk[tt][dd] = k[((long long)b * T + tt) * C + h * d + dd];
    sv[tt][dd] = v[((long long)b * T + tt) * C + h * d + dd];
  }
  __syncthreads();
  if (t >= T) return;
  float qv[32];
  for (int i = 0; i < d; i++) qv[i] = q[((long long)b * T + t) * C + h * d + i];
  float sc[64];
  float mx = -INFINITY, scale = rsqrtf((float)d);
  for (int j = 0; j < T; j++) {
    float s = 0.f;
    for (int i = 0; i < d; i++) s = fmaf(qv[i], sk[j][i], s);
    s *= scale;
    sc[j] = s;
    mx = fmaxf(mx, s);
  }
  float l = 0.f;
  for (int j = 0; j < T; j++) { sc[j] = expf(sc[j] - mx); l += sc[j]; }
  float inv = 1.f / l;
  for (int i = 0; i < d; i++) {
    float o = 0.f;
    for (int j = 0; j < T; j++) o = fmaf(sc[j], sv[j][i], o);
    out[((long long)b * T + t) * C + h * d + i] = o * inv;
  }
}

int launch_tok_self_attn(const float* q, const float* k, const float* v, int B, int T, int heads, int d, float* out,
                         cudaStream_t st) {
  if (T > 64 || d > 32) return cvb_fail(CV_ERR_INVALID, "tok_self_attn: T<=64, d<=32");
  CVB_LAUNCH(k_tok_self_attn, dim3(B, heads), dim3(64), 0, st, q, k, v, T, heads, d, out);
  return CV_OK;
}

// tokens -> image attention, d == 16.  CTA = (key split, head, image): stages its 128 keys/values of one head in shared
// memory; thread = one decoder token (query) running an online softmax over the split's keys.  Every lane reads the
// same K/V address (broadcast LDS.128, no bank conflicts) and keeps its (max, sum, O[16]) in registers, so there are no
// cross-lane reductions at all; the partials per (query, split) are merged by k_attn_t2i_combine.
constexpr int T2I_SPLIT_KEYS = 128;  // 16 KB of shared memory per CTA -> 14 CTAs per SM
__global__ void __launch_bounds__(64) k_attn_t2i_part(const float* __restrict__ q, long long q_img_stride,
                                                      const float* __restrict__ K, const float* __restrict__ V,
                                                      long long ld_kv, int T, int Nk, int heads, float* __restrict__ part) {
  __shared__ __align__(16) float sk[T2I_SPLIT_KEYS * 16];
  __shared__ __align__(16) float sv[T2I_SPLIT_KEYS * 16];
  const int split = blockIdx.x, h = blockIdx.y, b = blockIdx.z;
  const int nsplit = gridDim.x;
  const int k0 = split * T2I_SPLIT_KEYS;
  const int nk = min(T2I_SPLIT_KEYS, Nk - k0);
  for (int i = threadIdx.x; i < nk * 4; i += 64) {
    int j = i >> 2, c = (i & 3) * 4;
    long long row = (long long)b * Nk + k0 + j;
    *(float4*)(sk + j * 16 + c) = *(const float4*)(K + row * ld_kv + h * 16 + c);
    *(float4*)(sv + j * 16 + c) = *(const float4*)(V + row * ld_kv + h * 16 + c);
  }
  __syncthreads();
  const int t = threadIdx.x;
  if (t >= T) return;
  const int C = heads * 16;
  const float* qp = q + (long long)b * q_img_stride + (long long)t * C + h * 16;
  // packed fp32x2 arithmetic along the head dimension: q, k, v and the output accumulator are 8 pairs each, so a key
  // costs 8 + 8 FFMA2 instead of 16 + 16 FFMA (the loop is issue-bound: one thread per query token)
  uint64_t qv[8];
#pragma unroll
  for (int i = 0; i < 8; i++) qv[i] = pk2(__ldg(qp + 2 * i) * 0.36067376f, __ldg(qp + 2 * i + 1) * 0.36067376f);  // 1/sqrt(16) * log2(e)
  float m = -INFINITY, l = 0.f;
  uint64_t o[8];
#pragma unroll
  for (int i = 0; i < 8; i++) o[i] = 0ull;
  for (int j0 = 0; j0 < nk; j0 += 8) {
    // 8 keys per step: one rescale per step instead of per key
    float sc[8], mx = m;
#pragma unroll
    for (int u = 0; u < 8; u++) {
      float sdot = -INFINITY;
      if (j0 + u < nk) {
        const ulonglong2* kr = (const ulonglong2*)(sk + (j0 + u) * 16);
        uint64_t acc = 0ull;
#pragma unroll
        for (int c = 0; c < 4; c++) {
          const ulonglong2 kk = kr[c];
          acc = fma2(qv[2 * c], kk.x, acc);
          acc = fma2(qv[2 * c + 1], kk.y, acc);
        }
        float a0, a1;
        up2(acc, a0, a1);
        sdot = a0 + a1;
      }
      sc[u] = sdot;
      mx = fmaxf(mx, sdot);
    }
    float alpha;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(alpha) : "f"(m - mx));  // m = -inf -> 0
    l *= alpha;
    const uint64_t al2 = pk2(alpha, alpha);
#pragma unroll
    for (int i = 0; i < 8; i++) o[i] = mul2(o[i], al2);
    m = mx;
#pragma unroll
    for (int u = 0; u < 8; u++) {
      float pe;
      asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(pe) : "f"(sc[u] - mx));  // masked keys: ex2(-inf) = 0
      l += pe;
      if (j0 + u < nk) {
        const ulonglong2* vr = (const ulonglong2*)(sv + (j0 + u) * 16);
        const uint64_t pe2 = pk2(pe, pe);
#pragma unroll
        for (int c = 0; c < 4; c++) {
          const ulonglong2 vv = vr[c];
          o[2 * c] = fma2(pe2, vv.x, o[2 * c]);
          o[2 * c + 1] = fma2(pe2, vv.y, o[2 * c + 1]);
        }
      }
    }
  }
  // partial in natural-log convention of the combine kernel: stored max is in log2 units -> convert
  float* pp = part + ((((long long)b * heads + h) * T + t) * nsplit + split) * 18;
  pp[0] = m * 0.69314718056f;
  pp[1] = l;
#pragma unroll
  for (int i = 0; i < 8; i++) up2(o[i], pp[2 + 2 * i], pp[3 + 2 * i]);
}

__global__ void k_attn_t2i_combine(const float* __restrict__ part, int T, int heads, int nsplit, long long total,
                                   float* __restrict__ out) {
  long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;  // (b, h, t, i)
  if (idx >= total) return;
  int i = (int)(idx & 15);
  long long bht = idx >> 4;
  int t = (int)(bht % T);
  int h = (int)((bht / T) % heads);
  long long b = bht / ((long long)T * heads);
  const float* p = part + bht * nsplit * 18;
  float mx = -INFINITY;
  for (int s = 0; s < nsplit; s++) mx = fmaxf(mx, p[s * 18]);
  float l = 0.f, o = 0.f;
  for (int s = 0; s < nsplit; s++) {
    float w = expf(p[s * 18] - mx);
    l += w * p[s * 18 + 1];
    o += w * p[s * 18 + 2 + i];
  }
  out[(b * T + t) * (heads * 16) + h * 16 + i] = o / l;
}

size_t attn_t2i_scratch_floats(int B, int T, int heads, int d) {
  (void)d;
  return (size_t)B * heads * T * 32 * 18 + 64;
}

int launch_attn_t2i(const float* q, long long q_img_stride, const float* K, const float* V, long long ld_kv, int B, int T,
                    int Nk, int heads, int d, float* out, float* scratch, cudaStream_t st) {
  if (d != 16) return cvb_fail(CV_ERR_INVALID, "attn_t2i: head dim must be 16");
  int nsplit = (Nk + T2I_SPLIT_KEYS - 1) / T2I_SPLIT_KEYS;
  if (nsplit > 32) return cvb_fail(CV_ERR_INVALID, "attn_t2i: at most 4096 keys");
  cvb_next_work(4.0 * B * (double)T * Nk * heads * d);
  if (T > 64) return cvb_fail(CV_ERR_INVALID, "attn_t2i: at most 64 query tokens");
  CVB_LAUNCH(k_attn_t2i_part, dim3(nsplit, heads, B), dim3(64), 0, st, q, q_img_stride, K, V, ld_kv, T, Nk, heads, scratch);
  long long total = (long long)B * heads * T * 16;
  CVB_LAUNCH(k_attn_t2i_combine, dim3((unsigned)((total + 255) / 256)), dim3(256), 0, st, scratch, T, heads, nsplit, total,
             out);
  return CV_OK;
}

// image -> tokens attention, d == 16: thread = one image token (query row), looping over the heads; the image's T
// decoder tokens (K/V, fp32) sit in shared memory and every lane of a warp reads the same K/V address (broadcast, no
// bank conflicts); T is a template parameter so the score array lives in registers.
template <int T>
__global__ void __launch_bounds__(256) k_attn_i2t(const float* __restrict__ Q, long long ld_q, const float* __restrict__ K,
                                                  const float* __restrict__ V, int Nq, int heads, int fp16,
                                                  __nv_bfloat16* __restrict__ out) {
  extern __shared__ float sm[];
  const int C = heads * 16;
  float* sk = sm;
  float* sv = sm + T * C;
  const int b = blockIdx.y;
  for (int i = threadIdx.x; i < T * C / 4; i += 256) {
    ((float4*)sk)[i] = ((const float4*)(K + (long long)b * T * C))[i];
    ((float4*)sv)[i] = ((const float4*)(V + (long long)b * T * C))[i];
  }
  __syncthreads();
  const int tq = blockIdx.x * 256 + threadIdx.x;
  if (tq >= Nq) return;
  const long long row = (long long)b * Nq + tq;
  for (int h = 0; h < heads; h++) {
    float qv[16];
    const float4* qp = (const float4*)(Q + row * ld_q + h * 16);
#pragma unroll
    for (int c = 0; c < 4; c++) {
      float4 x = qp[c];
      // 1/sqrt(16) and log2(e) folded into q so the softmax is a bare ex2
      qv[c * 4] = x.x * 0.36067376f; qv[c * 4 + 1] = x.y * 0.36067376f;
      qv[c * 4 + 2] = x.z * 0.36067376f; qv[c * 4 + 3] = x.w * 0.36067376f;
    }
    float sc[T];
    float mx = -INFINITY;
#pragma unroll
    for (int j = 0; j < T; j++) {
      const float4* kr = (const float4*)(sk + j * C + h * 16);
      float s = 0.f;
#pragma unroll
      for (int c = 0; c < 4; c++) {
        float4 kk = kr[c];
        s = fmaf(qv[c * 4], kk.x, s); s = fmaf(qv[c * 4 + 1], kk.y, s);
        s = fmaf(qv[c * 4 + 2], kk.z, s); s = fmaf(qv[c * 4 + 3], kk.w, s);
      }
      sc[j] = s;
      mx = fmaxf(mx, s);
    }
    float l = 0.f, o[16];
#pragma unroll
    for (int i = 0; i < 16; i++) o[i] = 0.f;
#pragma unroll
    for (int j = 0; j < T; j++) {
      float pe;
      asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(pe) : "f"(sc[j] - mx));
      l += pe;
      const float4* vr = (const float4*)(sv + j * C + h * 16);
#pragma unroll
      for (int c = 0; c < 4; c++) {
        float4 vv = vr[c];
        o[c * 4] = fmaf(pe, vv.x, o[c * 4]); o[c * 4 + 1] = fmaf(pe, vv.y, o[c * 4 + 1]);
        o[c * 4 + 2] = fmaf(pe, vv.z, o[c * 4 + 2]); o[c * 4 + 3] = fmaf(pe, vv.w, o[c * 4 + 3]);
      }
    }
    const float inv = 1.f / l;
    uint32_t w[8];
#pragma unroll
    for (int i = 0; i < 8; i++) w[i] = tc::pack16(fp16, o[2 * i] * inv, o[2 * i + 1] * inv);
    uint4* dst = (uint4*)(out + row * C + h * 16);
    dst[0] = make_uint4(w[0], w[1], w[2], w[3]);
    dst[1] = make_uint4(w[4], w[5], w[6], w[7]);
  }
}

int launch_attn_i2t(const float* Q, long long ld_q, const float* K, const float* V, int B, int Nq, int T, int heads, int d,
                    int fp16, __nv_bfloat16* out, cudaStream_t st) {
  if (d != 16 || T != 38 || heads > 16) return cvb_fail(CV_ERR_INVALID, "attn_i2t: built for d == 16, T == 38 decoder tokens");
  size_t smem = (size_t)2 * T * heads * 16 * sizeof(float);
  cvb_next_work(4.0 * B * (double)T * Nq * heads * d);
  CVB_LAUNCH((k_attn_i2t<38>), dim3((Nq + 255) / 256, B), dim3(256), smem, st, Q, ld_q, K, V, Nq, heads, fp16, out);
  return CV_OK;
}

// ------------------------------------------------------------------------------------------------ LayerNorm2d + GELU (C == 64)
__global__ void __launch_bounds__(256) k_ln2d_gelu(const float* __restrict__ X, long long rows, const float* __restrict__ g,
                                                   const float* __restrict__ b, float eps, int fp16,
                                                   __nv_bfloat16* __restrict__ out) {
  // 16 lanes per row (float4 each), 2 rows per warp
  long long r = ((long long)blockIdx.x * 256 + threadIdx.x) >> 4;
  int l = threadIdx.x & 15;
  bool ok = r < rows;
  float4 v = ok ? *(const float4*)(X + r * 64 + l * 4) : make_float4(0.f, 0.f, 0.f, 0.f);
  float s = v.x + v.y + v.z + v.w;
  for (int o = 8; o; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  float mean = s * (1.f / 64.f);
  float a0 = v.x - mean, a1 = v.y - mean, a2 = v.z - mean, a3 = v.w - mean;
  float q = a0 * a0 + a1 * a1 + a2 * a2 + a3 * a3;
  for (int o = 8; o; o >>= 1) q += __shfl_xor_sync(0xffffffffu, q, o);
  // LayerNorm2d (sam2): (x - u) / sqrt(var + eps) * w + b
  float rstd = 1.f / sqrtf(q * (1.f / 64.f) + eps);
  if (!ok) return;
  float4 gg = ((const float4*)g)[l], bb = ((const float4*)b)[l];
  float y0 = gelu_exact(a0 * rstd * gg.x + bb.x), y1 = gelu_exact(a1 * rstd * gg.y + bb.y);
  float y2 = gelu_exact(a2 * rstd * gg.z + bb.z), y3 = gelu_exact(a3 * rstd * gg.w + bb.w);
  *(uint2*)(out + r * 64 + l * 4) = make_uint2(tc::pack16(fp16, y0, y1), tc::pack16(fp16, y2, y3));
}

__global__ void __launch_bounds__(256) k_count_sat16(const uint4* __restrict__ p, long long n16, const uint16_t* __restrict__ tail,
                                                     int n_tail, unsigned int* __restrict__ counter) {
  unsigned int c = 0;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n16; i += (long long)gridDim.x * blockDim.x) {
    const uint4 v = p[i];
    const uint32_t w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
    for (int j = 0; j < 4; j++) c += ((w[j] & 0x7FFFu) == 0x7BFFu) + (((w[j] >> 16) & 0x7FFFu) == 0x7BFFu);
  }
  if (blockIdx.x == 0 && (int)threadIdx.x < n_tail) c += (tail[threadIdx.x] & 0x7FFFu) == 0x7BFFu;
  c = __reduce_add_sync(0xffffffffu, c);
  if ((threadIdx.x & 31) == 0 && c) atomicAdd(counter, c);
}

int launch_count_sat16(const uint16_t* p, long long n, unsigned int* counter, cudaStream_t st) {
  if (((uintptr_t)p & 15) != 0) return cvb_fail(CV_ERR_INVALID, "count_sat16: buffer must be 16-byte aligned");
  const long long n16 = n / 8;
  const int grid = (int)std::min<long long>((n16 + 255) / 256 + 1, 148 * 8);
  CVB_LAUNCH(k_count_sat16, dim3(grid), dim3(256), 0, st, (const uint4*)p, n16, p + n16 * 8, (int)(n - n16 * 8), counter);
  return CV_OK;
}

int launch_ln2d_gelu(const float* X, long long rows, int C, const float* g, const float* b, float eps, int fp16,
                     __nv_bfloat16* out, cudaStream_t st) {
  if (C != 64) return cvb_fail(CV_ERR_INVALID, "ln2d_gelu: C must be 64");
  cvb_next_work((double)rows * 64 * 6);
  CVB_LAUNCH(k_ln2d_gelu, dim3((unsigned)((rows * 16 + 255) / 256)), dim3(256), 0, st, X, rows, g, b, eps, fp16, out);
  return CV_OK;
}

// ------------------------------------------------------------------------------------------------ mask product
__global__ void __launch_bounds__(256) k_mask_product(const float* __restrict__ U, const float* __restrict__ hyper, int P,
                                                      float delta, float* __restrict__ masks, unsigned int* __restrict__ counts) {
  __shared__ float sh[4 * 32];
  int b = blockIdx.y;
  if (threadIdx.x < 128) sh[threadIdx.x] = hyper[(long long)b * 128 + threadIdx.x];
  __syncthreads();
  int p = blockIdx.x * 256 + threadIdx.x;
  unsigned ci = 0, cu = 0;
  if (p < P) {
    const float4* u = (const float4*)(U + ((long long)b * P + p) * 32);
    float m[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
    for (int c = 0; c < 8; c++) {
      float4 x = u[c];
#pragma unroll
      for (int k = 0; k < 4; k++) {
        const float* hk = sh + k * 32 + c * 4;
        m[k] = fmaf(hk[0], x.x, m[k]); m[k] = fmaf(hk[1], x.y, m[k]);
        m[k] = fmaf(hk[2], x.z, m[k]); m[k] = fmaf(hk[3], x.w, m[k]);
      }
    }
#pragma unroll
    for (int k = 0; k < 4; k++) masks[((long long)b * 4 + k) * P + p] = m[k];
    ci = m[0] > delta;
    cu = m[0] > -delta;
  }
  unsigned bi = __ballot_sync(0xffffffffu, ci), bu = __ballot_sync(0xffffffffu, cu);
  if ((threadIdx.x & 31) == 0) {
    if (bi) atomicAdd(counts + 2 * b, __popc(bi));
    if (bu) atomicAdd(counts + 2 * b + 1, __popc(bu));
  }
}

int launch_mask_product(const float* U, const float* hyper, int B, int P, float delta, float* masks, unsigned int* counts,
                        cudaStream_t st) {
  CVB_CHECK(cudaMemsetAsync(counts, 0, (size_t)B * 2 * sizeof(unsigned int), st));
  cvb_next_work((double)B * P * (32 + 4) * 4);
  CVB_LAUNCH(k_mask_product, dim3((P + 255) / 256, B), dim3(256), 0, st, U, hyper, P, delta, masks, counts);
  return CV_OK;
}

__global__ void __launch_bounds__(256) k_select_mask(const float* __restrict__ masks, const float* __restrict__ iou,
                                                     const unsigned int* __restrict__ counts, int P, float thresh,
                                                     float* __restrict__ low, float* __restrict__ iou_out, int* __restrict__ sel) {
  int b = blockIdx.y;
  // stability = area_i / area_u (1 when area_u == 0); stable -> token 0, else best-IoU of tokens 1..3 (first max)
  float ai = (float)counts[2 * b], au = (float)counts[2 * b + 1];
  float stab = au > 0.f ? ai / au : 1.f;
  int s = 0;
  if (!(stab >= thresh)) {
    const float* io = iou + b * 4;
    s = 1;
    if (io[2] > io[s]) s = 2;
    if (io[3] > io[s]) s = 3;
  }
  int p = blockIdx.x * 256 + threadIdx.x;
  if (p < P) low[(long long)b * P + p] = masks[((long long)b * 4 + s) * P + p];
  if (p == 0) { iou_out[b] = iou[b * 4 + s]; sel[b] = s; }
}

int launch_select_mask(const float* masks, const float* iou, const unsigned int* counts, int B, int P, float thresh,
                       float* low_res, float* iou_out, int* sel, cudaStream_t st) {
  CVB_LAUNCH(k_select_mask, dim3((P + 255) / 256, B), dim3(256), 0, st, masks, iou, counts, P, thresh, low_res, iou_out, sel);
  return CV_OK;
}

// ------------------------------------------------------------------------------------------------ tail
// CTA = 32 x 32 output pixels.  up[42][44]: bilinearly upsampled logits of the tile + halo 5 (zero outside the
// 1024^2 image: Conv2d padding='same' pads the upsampled map with zeros).  Thread = 4 consecutive pixels of a row.
constexpr int TL = 32, THALO = 5, TUP = TL + 2 * THALO;
// weights tap-major with the 4 channels of a tap contiguous, so one broadcast LDS.128 feeds 16 FMAs (4 channels x 4
// pixels); with scalar weight loads the kernel was LSU-bound (one shared-memory wavefront per 4 FMAs).
// A packed-fp32x2 (FFMA2) version of this stencil — channel pairs as accumulators, tile stored duplicated — issues
// half the instructions and runs at the same speed (3.72 vs 3.58 ms for 64 images; ncu: FMA pipe 60 % busy at 4 warps
// per sub-partition, FFMA2 occupies the pipe for two cycles), so the scalar form stays.
struct __align__(16) TailConst {
  float4 w3[9], w5[25], w7[49], w11[121];
  float b[16], cw[16], cb;
};

__global__ void __launch_bounds__(256) k_tail(const float* __restrict__ low, int src_full, RefineWeights rw, float* __restrict__ high,
                                              uint8_t* __restrict__ mask, int* __restrict__ extents, int border_only) {
  __shared__ __align__(16) float up[TUP][TUP + 2];
  __shared__ TailConst tc_;
  // the interior tiles belong to k_tail_phase
  if (border_only && blockIdx.x > 0 && blockIdx.x < gridDim.x - 1 && blockIdx.y > 0 && blockIdx.y < gridDim.y - 1) return;
  const int S = 1024, LS = 256;
  const int b = blockIdx.z;
  const int X0 = blockIdx.x * TL, Y0 = blockIdx.y * TL;
  const float* lo = low + (long long)b * (src_full ? S * S : LS * LS);
  if (rw.use_refine) {
    // global layout [channel][k*k] -> shared [tap][channel]
    for (int i = threadIdx.x; i < 4 * 9; i += 256) ((float*)tc_.w3)[(i % 9) * 4 + i / 9] = rw.w[0][i];
    for (int i = threadIdx.x; i < 4 * 25; i += 256) ((float*)tc_.w5)[(i % 25) * 4 + i / 25] = rw.w[1][i];
    for (int i = threadIdx.x; i < 4 * 49; i += 256) ((float*)tc_.w7)[(i % 49) * 4 + i / 49] = rw.w[2][i];
    for (int i = threadIdx.x; i < 4 * 121; i += 256) ((float*)tc_.w11)[(i % 121) * 4 + i / 121] = rw.w[3][i];
    if (threadIdx.x < 16) {
      tc_.b[threadIdx.x] = rw.b[threadIdx.x >> 2][threadIdx.x & 3];
      tc_.cw[threadIdx.x] = rw.cw[threadIdx.x];
    }
    if (threadIdx.x == 0) tc_.cb = rw.cb;
  }
  for (int i = threadIdx.x; i < TUP * TUP; i += 256) {
    int uy = i / TUP, ux = i - uy * TUP;
    int Y = Y0 - THALO + uy, X = X0 - THALO + ux;
    float v = 0.f;
    if (Y >= 0 && Y < S && X >= 0 && X < S && src_full) {
      v = lo[Y * S + X];
    } else if (Y >= 0 && Y < S && X >= 0 && X < S) {
      // F.interpolate(bilinear, align_corners=False): src = max(0, (dst + 0.5) * 0.25 - 0.5)
      float sy = fmaxf(0.f, (Y + 0.5f) * 0.25f - 0.5f), sx = fmaxf(0.f, (X + 0.5f) * 0.25f - 0.5f);
      int y0 = (int)sy, x0 = (int)sx;
      int y1 = y0 + (y0 < LS - 1 ? 1 : 0), x1 = x0 + (x0 < LS - 1 ? 1 : 0);
      float ly = sy - y0, lx = sx - x0;
      float hy = 1.f - ly, hx = 1.f - lx;
      v = hy * (hx * lo[y0 * LS + x0] + lx * lo[y0 * LS + x1]) + ly * (hx * lo[y1 * LS + x0] + lx * lo[y1 * LS + x1]);
    }
    up[uy][ux] = v;
  }
  __syncthreads();
  const int ty = threadIdx.x >> 3, tx = (threadIdx.x & 7) * 4;  // 32 rows x 8 groups of 4 pixels
  float res[4];
  if (rw.use_refine) {
    float acc[4][4][4];  // [branch][channel][pixel]
#pragma unroll
    for (int a = 0; a < 4; a++)
#pragma unroll
      for (int c = 0; c < 4; c++)
#pragma unroll
        for (int p = 0; p < 4; p++) acc[a][c][p] = 0.f;
#pragma unroll 1
    for (int dy = -5; dy <= 5; dy++) {
      float row[16];
      {
        // tx is a multiple of 4 and the row pitch is 44 floats: four aligned 16-byte loads cover row[0..15] (14 used)
        const float4* rp = (const float4*)&up[ty + THALO + dy][tx];
#pragma unroll
        for (int i = 0; i < 4; i++) {
          float4 v = rp[i];
          row[4 * i] = v.x; row[4 * i + 1] = v.y; row[4 * i + 2] = v.z; row[4 * i + 3] = v.w;
        }
      }
#define CVB_TAP(ACC, WV, OFF)                                   \
  {                                                             \
    const float4 wv = (WV);                                     \
    _Pragma("unroll") for (int p = 0; p < 4; p++) {             \
      ACC[0][p] = fmaf(wv.x, row[(OFF) + p], ACC[0][p]);        \
      ACC[1][p] = fmaf(wv.y, row[(OFF) + p], ACC[1][p]);        \
      ACC[2][p] = fmaf(wv.z, row[(OFF) + p], ACC[2][p]);        \
      ACC[3][p] = fmaf(wv.w, row[(OFF) + p], ACC[3][p]);        \
    }                                                           \
  }
      // branch 3: k = 11
#pragma unroll
      for (int dx = 0; dx < 11; dx++) CVB_TAP(acc[3], tc_.w11[(dy + 5) * 11 + dx], dx)
      if (dy >= -3 && dy <= 3) {
#pragma unroll
        for (int dx = 0; dx < 7; dx++) CVB_TAP(acc[2], tc_.w7[(dy + 3) * 7 + dx], dx + 2)
      }
      if (dy >= -2 && dy <= 2) {
#pragma unroll
        for (int dx = 0; dx < 5; dx++) CVB_TAP(acc[1], tc_.w5[(dy + 2) * 5 + dx], dx + 3)
      }
      if (dy >= -1 && dy <= 1) {
#pragma unroll
        for (int dx = 0; dx < 3; dx++) CVB_TAP(acc[0], tc_.w3[(dy + 1) * 3 + dx], dx + 4)
      }
#undef CVB_TAP
    }
#pragma unroll
    for (int p = 0; p < 4; p++) {
      float o = tc_.cb;
#pragma unroll
      for (int a = 0; a < 4; a++)
#pragma unroll
        for (int c = 0; c < 4; c += 2) {
          float g0 = acc[a][c][p] + tc_.b[a * 4 + c], g1 = acc[a][c + 1][p] + tc_.b[a * 4 + c + 1];
          gelu_fast2(g0, g1);
          o = fmaf(tc_.cw[a * 4 + c], g0, o);
          o = fmaf(tc_.cw[a * 4 + c + 1], g1, o);
        }
      res[p] = o;
    }
  } else {
#pragma unroll
    for (int p = 0; p < 4; p++) res[p] = up[ty + THALO][tx + THALO + p];
  }
  const int Y = Y0 + ty, X = X0 + tx;
  long long o = ((long long)b * S + Y) * S + X;
  if (high) *(float4*)(high + o) = make_float4(res[0], res[1], res[2], res[3]);
  if (mask) {
    uint32_t m = 0;
    int xmin = INT_MAX, xmax = -1;
#pragma unroll
    for (int p = 0; p < 4; p++)
      if (res[p] > 0.0f) {
        m |= 0xFFu << (8 * p);
        xmin = min(xmin, X + p);
        xmax = max(xmax, X + p);
      }
    *(uint32_t*)(mask + o) = m;
    if (extents) {
      int ymin = xmax >= 0 ? Y : INT_MAX, ymax = xmax >= 0 ? Y : -1;
      for (int k = 16; k; k >>= 1) {
        xmin = min(xmin, __shfl_xor_sync(0xffffffffu, xmin, k));
        xmax = max(xmax, __shfl_xor_sync(0xffffffffu, xmax, k));
        ymin = min(ymin, __shfl_xor_sync(0xffffffffu, ymin, k));
        ymax = max(ymax, __shfl_xor_sync(0xffffffffu, ymax, k));
      }
      if ((threadIdx.x & 31) == 0 && xmax >= 0) {
        atomicMin(extents + 4 * b, xmin);
        atomicMin(extents + 4 * b + 1, ymin);
        atomicMax(extents + 4 * b + 2, xmax);
        atomicMax(extents + 4 * b + 3, ymax);
      }
    }
  }
}

// Interior of the fused tail on the LOW-RES grid (sam2_weights.refine_phase_tables): away from the image border the
// composition  conv_k(upsample_x4(low))  is, for each of the 16 output phases (Y % 4, X % 4), one position-independent
// 5 x 5 stencil on the 256^2 logits — 25 taps x 16 channels = 400 FMA per output pixel instead of 816.
// CTA = 64 x 32 output pixels = 16 x 8 low-res cells; warp w evaluates phases 2w and 2w+1 (so the composite-weight
// address is warp-uniform: one broadcast LDS.128 feeds 16 FMAs), lane = cell row x group of 4 cells in a row.
// The 32 border tiles of each side keep the general kernel above (k_tail, border_only).
constexpr int TP_W = 64, TP_H = 32, TP_CX = TP_W / 4, TP_CY = TP_H / 4, TP_LD = TP_W + 1;
__global__ void __launch_bounds__(256, 2) k_tail_phase(const float* __restrict__ low, const float* __restrict__ comp,
                                                       RefineWeights rw, float* __restrict__ high, uint8_t* __restrict__ mask,
                                                       int* __restrict__ extents) {
  __shared__ __align__(16) float s_comp[16 * 25 * 16];
  __shared__ float s_low[TP_CY + 4][TP_CX + 4 + 1];
  __shared__ float s_out[TP_H][TP_LD];
  __shared__ float s_b[16], s_cw[16];
  const int S = 1024, LS = 256;
  const int b = blockIdx.z;
  const int X0 = 32 + blockIdx.x * TP_W, Y0 = 32 + blockIdx.y * TP_H;
  const int cx0 = X0 >> 2, cy0 = Y0 >> 2;
  const float* lo = low + (size_t)b * LS * LS;
  for (int i = threadIdx.x; i < 16 * 25 * 16 / 4; i += 256) ((float4*)s_comp)[i] = __ldg((const float4*)comp + i);
  for (int i = threadIdx.x; i < (TP_CY + 4) * (TP_CX + 4); i += 256) {
    const int r = i / (TP_CX + 4), c = i - r * (TP_CX + 4);
    s_low[r][c] = lo[(cy0 - 2 + r) * LS + cx0 - 2 + c];
  }
  if (threadIdx.x < 16) {
    s_b[threadIdx.x] = rw.b[threadIdx.x >> 2][threadIdx.x & 3];
    s_cw[threadIdx.x] = rw.cw[threadIdx.x];
  }
  __syncthreads();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int cy = lane >> 2, cg = (lane & 3) * 4;
  float win[5][8];  // low-res rows cy-2 .. cy+2, columns cg-2 .. cg+5 of the tile
#pragma unroll
  for (int tu = 0; tu < 5; tu++)
#pragma unroll
    for (int c = 0; c < 8; c++) win[tu][c] = s_low[cy + tu][cg + c];
#pragma unroll 1
  for (int ph = 0; ph < 2; ph++) {
    const int p = warp * 2 + ph;
    const int py = p >> 2, px = p & 3;
    const float4* T = (const float4*)s_comp + p * 25 * 4;
    float acc[16][4];
#pragma unroll
    for (int c = 0; c < 16; c++)
#pragma unroll
      for (int j = 0; j < 4; j++) acc[c][j] = 0.f;
#pragma unroll
    for (int tu = 0; tu < 5; tu++)
#pragma unroll
      for (int tv = 0; tv < 5; tv++) {
#pragma unroll
        for (int q = 0; q < 4; q++) {
          const float4 w = T[(tu * 5 + tv) * 4 + q];
#pragma unroll
          for (int j = 0; j < 4; j++) {
            const float v = win[tu][j + tv];
            acc[4 * q + 0][j] = fmaf(w.x, v, acc[4 * q + 0][j]);
            acc[4 * q + 1][j] = fmaf(w.y, v, acc[4 * q + 1][j]);
            acc[4 * q + 2][j] = fmaf(w.z, v, acc[4 * q + 2][j]);
            acc[4 * q + 3][j] = fmaf(w.w, v, acc[4 * q + 3][j]);
          }
        }
      }
#pragma unroll
    for (int j = 0; j < 4; j++) {
      float o = rw.cb;
#pragma unroll
      for (int c = 0; c < 16; c += 2) {
        float g0 = acc[c][j] + s_b[c], g1 = acc[c + 1][j] + s_b[c + 1];
        gelu_fast2(g0, g1);
        o = fmaf(s_cw[c], g0, o);
        o = fmaf(s_cw[c + 1], g1, o);
      }
      s_out[cy * 4 + py][(cg + j) * 4 + px] = o;
    }
  }
  __syncthreads();
  // write-out: thread = 8 consecutive pixels of one tile row
  const int ty = threadIdx.x >> 3, tx = (threadIdx.x & 7) * 8;
  const int Y = Y0 + ty, X = X0 + tx;
  const size_t o = ((size_t)b * S + Y) * S + X;
  float r[8];
#pragma unroll
  for (int k = 0; k < 8; k++) r[k] = s_out[ty][tx + k];
  if (high) {
    *(float4*)(high + o) = make_float4(r[0], r[1], r[2], r[3]);
    *(float4*)(high + o + 4) = make_float4(r[4], r[5], r[6], r[7]);
  }
  if (mask) {
    uint32_t m0 = 0, m1 = 0;
    int xmin = INT_MAX, xmax = -1;
#pragma unroll
    for (int k = 0; k < 8; k++)
      if (r[k] > 0.0f) {
        if (k < 4) m0 |= 0xFFu << (8 * k);
        else m1 |= 0xFFu << (8 * (k - 4));
        xmin = min(xmin, X + k);
        xmax = max(xmax, X + k);
      }
    *(uint2*)(mask + o) = make_uint2(m0, m1);
    if (extents) {
      int ymin = xmax >= 0 ? Y : INT_MAX, ymax = xmax >= 0 ? Y : -1;
      for (int k = 16; k; k >>= 1) {
        xmin = min(xmin, __shfl_xor_sync(0xffffffffu, xmin, k));
        xmax = max(xmax, __shfl_xor_sync(0xffffffffu, xmax, k));
        ymin = min(ymin, __shfl_xor_sync(0xffffffffu, ymin, k));
        ymax = max(ymax, __shfl_xor_sync(0xffffffffu, ymax, k));
      }
      if ((threadIdx.x & 31) == 0 && xmax >= 0) {
        atomicMin(extents + 4 * b, xmin);
        atomicMin(extents + 4 * b + 1, ymin);
        atomicMax(extents + 4 * b + 2, xmax);
        atomicMax(extents + 4 * b + 3, ymax);
      }
    }
  }
}

__global__ void k_init_extents(int* e, int B) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < B) { e[4 * i] = INT_MAX; e[4 * i + 1] = INT_MAX; e[4 * i + 2] = -1; e[4 * i + 3] = -1; }
}

int launch_tail(const float* low_res, int src_full, int B, const RefineWeights& rw, float* high_res, uint8_t* mask_u8,
                int* extents, cudaStream_t st) {
  if (extents) CVB_LAUNCH(k_init_extents, dim3((B + 127) / 128), dim3(128), 0, st, extents, B);
  cvb_next_work((double)B * (256.0 * 256 * 4 + 1024.0 * 1024 * ((high_res ? 4 : 0) + (mask_u8 ? 1 : 0))));
  // low-res source + refinement: interior on the phase tables, the one-tile border frame on the general stencil
  static const bool phase_on = getenv("CVB_TAIL_PHASE") ? atoi(getenv("CVB_TAIL_PHASE")) != 0 : true;
  const bool phase = phase_on && !src_full && rw.use_refine && rw.comp != nullptr;
  CVB_LAUNCH(k_tail, dim3(1024 / TL, 1024 / TL, B), dim3(256), 0, st, low_res, src_full, rw, high_res, mask_u8, extents,
             phase ? 1 : 0);
  if (phase)
    CVB_LAUNCH(k_tail_phase, dim3((1024 - 64) / TP_W, (1024 - 64) / TP_H, B), dim3(256), 0, st, low_res, rw.comp, rw, high_res,
               mask_u8, extents);
  return CV_OK;
}

// bilinear (align_corners=False) resize of the 1024^2 logits to the caller's (H, W) + threshold (sam2_infer.py:127,
// circuit_analyzer.py:356)
__global__ void __launch_bounds__(256) k_resize_threshold(const float* __restrict__ hi, int S, int H, int W, float sy_, float sx_,
                                                          float* __restrict__ out, uint8_t* __restrict__ mask,
                                                          int* __restrict__ extents) {
  int b = blockIdx.z;
  int x = blockIdx.x * 32 + (threadIdx.x & 31), y = blockIdx.y * 8 + (threadIdx.x >> 5);
  bool fg = false;
  if (x < W && y < H) {
    float sy = fmaxf(0.f, (y + 0.5f) * sy_ - 0.5f), sx = fmaxf(0.f, (x + 0.5f) * sx_ - 0.5f);
    int y0 = min((int)sy, S - 1), x0 = min((int)sx, S - 1);
    int y1 = y0 + (y0 < S - 1 ? 1 : 0), x1 = x0 + (x0 < S - 1 ? 1 : 0);
    float ly = sy - y0, lx = sx - x0, hy = 1.f - ly, hx = 1.f - lx;
    const float* p = hi + (long long)b * S * S;
    float v = hy * (hx * p[y0 * S + x0] + lx * p[y0 * S + x1]) + ly * (hx * p[y1 * S + x0] + lx * p[y1 * S + x1]);
    long long o = ((long long)b * H + y) * W + x;
    if (out) out[o] = v;
    fg = v > 0.0f;
    if (mask) mask[o] = fg ? 255 : 0;
  }
  if (extents) {
    int xmin = fg ? x : INT_MAX, xmax = fg ? x : -1, ymin = fg ? y : INT_MAX, ymax = fg ? y : -1;
    for (int k = 16; k; k >>= 1) {
      xmin = min(xmin, __shfl_xor_sync(0xffffffffu, xmin, k));
      xmax = max(xmax, __shfl_xor_sync(0xffffffffu, xmax, k));
      ymin = min(ymin, __shfl_xor_sync(0xffffffffu, ymin, k));
      ymax = max(ymax, __shfl_xor_sync(0xffffffffu, ymax, k));
    }
    if ((threadIdx.x & 31) == 0 && xmax >= 0) {
      atomicMin(extents + 4 * b, xmin);
      atomicMin(extents + 4 * b + 1, ymin);
      atomicMax(extents + 4 * b + 2, xmax);
      atomicMax(extents + 4 * b + 3, ymax);
    }
  }
}

int launch_resize_threshold(const float* high_res, int B, int S, int H, int W, float* out_logits, uint8_t* mask_u8,
                            int* extents, cudaStream_t st) {
  if (extents) CVB_LAUNCH(k_init_extents, dim3((B + 127) / 128), dim3(128), 0, st, extents, B);
  // area_pixel_compute_scale (align_corners=False, no scale_factor): scale = in / out
  float sy = (float)S / (float)H, sx = (float)S / (float)W;
  CVB_LAUNCH(k_resize_threshold, dim3((W + 31) / 32, (H + 7) / 8, B), dim3(256), 0, st, high_res, S, H, W, sy, sx,
             out_logits, mask_u8, extents);
  return CV_OK;
}

}  // namespace cvb
