// Node / connection analysis on device and the native-resolution CCL kernel.
// Reference behaviour: /root/reference/src/circuit_analyzer.py:1286-1605 and helpers (see include/cv_b200.h).
// Everything is batched over images (blockIdx.y / blockIdx.z = image) and HBM-bound integer work:
// coalesced 16-byte accesses where rows are contiguous, shared-memory halo tiles for the stencils,
// warp ballots for run detection, union-find with atomicMin for labelling.
#include <cuda_runtime.h>
#include <stdint.h>
#include <limits.h>

#include "../../include/cv_b200.h"
#include "common.cuh"
#include "node_prims.cuh"

namespace cvb {

// ------------------------------------------------------------------------------------------------
// a12: emptied = mask.copy(); emptied[box] = 0
// ------------------------------------------------------------------------------------------------
__global__ void k_copy16(const uint4* __restrict__ src, uint4* __restrict__ dst, size_t n16,
                         const uint8_t* __restrict__ src8, uint8_t* __restrict__ dst8, size_t n) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  size_t stride = (size_t)gridDim.x * blockDim.x;
  for (size_t k = i; k < n16; k += stride) dst[k] = __ldg(src + k);
  for (size_t k = n16 * 16 + i; k < n; k += stride) dst8[k] = src8[k];
}

// numpy_slices: the terminal-reclassification caller (:2245-2249) writes mask[max(0,ymin):min(H,ymax), ...] = 0 without
// the emptiness guard of :1341, so a negative upper bound wraps like a Python slice (end = H + ymax).
__global__ void k_zero_boxes(uint8_t* __restrict__ emptied, int H, int W, const cv_box* __restrict__ boxes,
                             const int32_t* __restrict__ box_offsets, int numpy_slices) {
  int b = blockIdx.y;
  int i = box_offsets[b] + blockIdx.x;
  if (i >= box_offsets[b + 1]) return;
  cv_box bx = boxes[i];
  if (!(bx.flags & CV_BOX_ZERO_IN_MASK)) return;
  int y0 = max(0, bx.ymin), y1 = min(H, bx.ymax);
  int x0 = max(0, bx.xmin), x1 = min(W, bx.xmax);
  if (numpy_slices) {
    if (y1 < 0) y1 = max(0, H + y1);
    if (x1 < 0) x1 = max(0, W + x1);
  }
  if (y0 >= y1 || x0 >= x1) return;
  uint8_t* img = emptied + (size_t)b * H * W;
  int bw = x1 - x0;
  int total = bw * (y1 - y0);
  for (int t = threadIdx.x; t < total; t += blockDim.x) {
    int yy = t / bw, xx = t - yy * bw;
    img[(size_t)(y0 + yy) * W + x0 + xx] = 0;
  }
}

// ------------------------------------------------------------------------------------------------
// a13: cv2.resize(emptied, (w', 600)) INTER_LINEAR, bit-exact fixed point
// ------------------------------------------------------------------------------------------------
__global__ void k_resize(const uint8_t* __restrict__ src, int H, int W, uint8_t* __restrict__ dst, int h, int w) {
  int b = blockIdx.z;
  int dx = blockIdx.x * blockDim.x + threadIdx.x;
  int dy = blockIdx.y * blockDim.y + threadIdx.y;
  if (dx >= w || dy >= h) return;
  const uint8_t* s = src + (size_t)b * H * W;
  ResizeTap tx = resize_tap(dx, w, W, true);
  ResizeTap ty = resize_tap(dy, h, H, false);
  const uint8_t* r0 = s + (size_t)ty.i0 * W;
  const uint8_t* r1 = s + (size_t)ty.i1 * W;
  int a = resize_hpass(r0[tx.i0], r0[tx.i1], tx);
  int c = resize_hpass(r1[tx.i0], r1[tx.i1], tx);
  dst[(size_t)b * h * w + (size_t)dy * w + dx] = resize_vpass(a, c, ty);
}

// ------------------------------------------------------------------------------------------------
// a14: GaussianBlur 5x5 (8.8 fixed point, reflect101) -> 5x5 max -> 5x5 min, one shared-memory tile pass.
// Also accumulates the image sum that get_contours' mean>127 test needs (a15).
// ------------------------------------------------------------------------------------------------
#define ENH_T 32
__global__ void __launch_bounds__(256) k_enhance(const uint8_t* __restrict__ src, uint8_t* __restrict__ dst, int h,
                                                 int w, unsigned long long* __restrict__ sums) {
  __shared__ uint8_t s_src[ENH_T + 12][ENH_T + 12];
  __shared__ uint16_t s_h[ENH_T + 12][ENH_T + 8];
  __shared__ uint8_t s_bl[ENH_T + 8][ENH_T + 8];
  __shared__ uint8_t s_di[ENH_T + 4][ENH_T + 4];
  __shared__ uint8_t s_t1[ENH_T + 8][ENH_T + 4];  // row maxima
  __shared__ uint8_t s_t2[ENH_T + 4][ENH_T];      // row minima
  __shared__ unsigned int s_sum;
  int b = blockIdx.z;
  const uint8_t* img = src + (size_t)b * h * w;
  int tx0 = blockIdx.x * ENH_T, ty0 = blockIdx.y * ENH_T;
  int tid = threadIdx.x;
  if (tid == 0) s_sum = 0;
  for (int t = tid; t < (ENH_T + 12) * (ENH_T + 12); t += 256) {
    int ly = t / (ENH_T + 12), lx = t - ly * (ENH_T + 12);
    int gy = reflect101(ty0 + ly - 6, h), gx = reflect101(tx0 + lx - 6, w);
    s_src[ly][lx] = img[(size_t)gy * w + gx];
  }
  __syncthreads();
  // The tile was loaded through reflect101, so s_src[r][c] already holds img[reflect(gy)][reflect(gx)] for
  // the halo coordinates; the 5-tap windows below therefore see BORDER_REFLECT_101 without further index math.
  for (int t = tid; t < (ENH_T + 12) * (ENH_T + 8); t += 256) {
    int ly = t / (ENH_T + 8), lx = t - ly * (ENH_T + 8);
    s_h[ly][lx] = (uint16_t)gauss5_h(s_src[ly][lx], s_src[ly][lx + 1], s_src[ly][lx + 2], s_src[ly][lx + 3],
                                     s_src[ly][lx + 4]);
  }
  __syncthreads();
  for (int t = tid; t < (ENH_T + 8) * (ENH_T + 8); t += 256) {
    int ly = t / (ENH_T + 8), lx = t - ly * (ENH_T + 8);
    s_bl[ly][lx] = gauss5_v(s_h[ly][lx], s_h[ly + 1][lx], s_h[ly + 2][lx], s_h[ly + 3][lx], s_h[ly + 4][lx]);
  }
  __syncthreads();
  // Morphology sees in-image pixels only: positions outside the image become 0 for the dilation (a maximum over
  // non-negative values ignores them) and 255 for the erosion.  Both 5x5 windows are separable: 5 + 5 taps instead of 25.
  for (int t = tid; t < (ENH_T + 8) * (ENH_T + 8); t += 256) {
    int ly = t / (ENH_T + 8), lx = t - ly * (ENH_T + 8);
    int gy = ty0 + ly - 4, gx = tx0 + lx - 4;
    if (gy < 0 || gy >= h || gx < 0 || gx >= w) s_bl[ly][lx] = 0;
  }
  __syncthreads();
  // dilate (two 3x3 iterations == 5x5 max), rows then columns
  for (int t = tid; t < (ENH_T + 8) * (ENH_T + 4); t += 256) {
    int ly = t / (ENH_T + 4), lx = t - ly * (ENH_T + 4);
    const uint8_t* r = &s_bl[ly][lx];
    s_t1[ly][lx] = (uint8_t)max(max(max((int)r[0], (int)r[1]), max((int)r[2], (int)r[3])), (int)r[4]);
  }
  __syncthreads();
  for (int t = tid; t < (ENH_T + 4) * (ENH_T + 4); t += 256) {
    int ly = t / (ENH_T + 4), lx = t - ly * (ENH_T + 4);
    int gy = ty0 + ly - 2, gx = tx0 + lx - 2;
    int m = max(max(max((int)s_t1[ly][lx], (int)s_t1[ly + 1][lx]), max((int)s_t1[ly + 2][lx], (int)s_t1[ly + 3][lx])),
                (int)s_t1[ly + 4][lx]);
    s_di[ly][lx] = (gy < 0 || gy >= h || gx < 0 || gx >= w) ? (uint8_t)255 : (uint8_t)m;
  }
  __syncthreads();
  // erode (5x5 min), rows then columns
  for (int t = tid; t < (ENH_T + 4) * ENH_T; t += 256) {
    int ly = t / ENH_T, lx = t - ly * ENH_T;
    const uint8_t* r = &s_di[ly][lx];
    s_t2[ly][lx] = (uint8_t)min(min(min((int)r[0], (int)r[1]), min((int)r[2], (int)r[3])), (int)r[4]);
  }
  __syncthreads();
  unsigned int local = 0;
  for (int t = tid; t < ENH_T * ENH_T; t += 256) {
    int ly = t / ENH_T, lx = t - ly * ENH_T;
    int gy = ty0 + ly, gx = tx0 + lx;
    if (gy < h && gx < w) {
      int m = min(min(min((int)s_t2[ly][lx], (int)s_t2[ly + 1][lx]), min((int)s_t2[ly + 2][lx], (int)s_t2[ly + 3][lx])),
                  (int)s_t2[ly + 4][lx]);
      dst[(size_t)b * h * w + (size_t)gy * w + gx] = (uint8_t)m;
      local += m;
    }
  }
  for (int o = 16; o; o >>= 1) local += __shfl_xor_sync(0xffffffffu, local, o);
  if ((tid & 31) == 0) atomicAdd(&s_sum, local);
  __syncthreads();
  if (tid == 0) atomicAdd(&sums[b], (unsigned long long)s_sum);
}

// a15 prelude: mean>127 inversion, 255->1 mutation of the returned array, binary image for contouring
// Also emits the binary image as BITS in flat pixel order (bit i of the image = pixel i, 32 per word): the border
// tracers keep that copy in shared memory (600 x 600 pixels = 45 KB).
__global__ void k_binarize(const uint8_t* __restrict__ enh_raw, uint8_t* __restrict__ enhanced_out,
                           uint8_t* __restrict__ bin, int n_per_image, const unsigned long long* __restrict__ sums,
                           cv_image_result* __restrict__ results, uint32_t* __restrict__ bits, int words_per_image) {
  int b = blockIdx.y;
  bool inv = sums[b] > 127ull * (unsigned long long)n_per_image;  // cv2.mean(img)[0] > 127
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i == 0) results[b].inverted = inv ? 1 : 0;
  bool f = false;
  if (i < n_per_image) {
    size_t p = (size_t)b * n_per_image + i;
    uint8_t v = enh_raw[p];
    if (enhanced_out) enhanced_out[p] = inv ? v : (v == 255 ? (uint8_t)1 : v);
    f = inv ? (v != 255) : (v != 0);
    bin[p] = (uint8_t)f;
  }
  const unsigned m = __ballot_sync(0xffffffffu, f);
  if (bits && (threadIdx.x & 31) == 0 && (i >> 5) < words_per_image) bits[(size_t)b * words_per_image + (i >> 5)] = m;
}

// ------------------------------------------------------------------------------------------------
// union-find connected-component labelling.  Parent of x is L[x] - OFF (OFF = 1 lets 0 mean background
// in the caller-visible label image; OFF = 0 is used internally where background is labelled as well).
// ------------------------------------------------------------------------------------------------
template <int OFF>
__device__ __forceinline__ int uf_find(const int* L, int x) {
  int y = __ldcg(L + x) - OFF;
  while (y != x) {
    x = y;
    y = __ldcg(L + x) - OFF;
  }
  return x;
}

// One warp = 32 consecutive pixels of one row.  Initial label = start of the pixel's horizontal run inside
// the warp segment (ballot + clz), so that only run boundaries need union operations afterwards.
template <int OFF, bool WITH_BG>
__global__ void __launch_bounds__(256) k_ccl_init(const uint8_t* __restrict__ bin, int* __restrict__ L, int h, int w) {
  int b = blockIdx.z;
  int x = blockIdx.x * 32 + threadIdx.x;
  int y = blockIdx.y * blockDim.y + threadIdx.y;
  if (y >= h) return;
  bool in = x < w;
  size_t base = (size_t)b * h * w;
  int p = y * w + x;
  bool f = in && bin[base + p] != 0;
  unsigned fgbits = __ballot_sync(0xffffffffu, f);
  unsigned bgbits = __ballot_sync(0xffffffffu, in && !f);
  if (!in) return;
  unsigned mine = f ? fgbits : bgbits;
  unsigned below = ~mine & ((1u << threadIdx.x) - 1u);
  int start = below ? (32 - __clz(below)) : 0;
  int lbl = p - (int)threadIdx.x + start;
  if (f || WITH_BG) L[base + p] = lbl + OFF;
  else L[base + p] = 0;  // only reachable with OFF == 1: background = 0
}

// Word-parallel merge / flatten on the flat bit plane (bit i = pixel i; k_binarize): one thread owns the 32 pixels
// (y, 32k .. 32k+31) as a register word, so a row pair is compared with a handful of bit operations and every CONTACT
// between a run of this row and a run of the row above is one union — instead of five byte loads and a chain of
// predicates per pixel.  k_ccl_init labelled every pixel with the start of its run inside the same 32-pixel segment,
// so a run's first pixel is its representative.  Foreground merges 8-connected, background 4-connected.
__device__ __forceinline__ uint32_t row_word(const uint32_t* __restrict__ B, int w, int y, int k, uint32_t* valid) {
  const int nv = min(32, w - 32 * k);
  const uint32_t vm = nv >= 32 ? 0xFFFFFFFFu : ((1u << nv) - 1u);
  const uint32_t f = (uint32_t)y * (uint32_t)w + 32u * (uint32_t)k;
  const uint32_t lo = __ldg(B + (f >> 5)), hi = __ldg(B + (f >> 5) + 1);  // the plane is padded by >= 1 word
  *valid = vm;
  return __funnelshift_r(lo, hi, f & 31) & vm;
}
__device__ __forceinline__ bool flat_bit(const uint32_t* __restrict__ B, uint32_t i) { return (__ldg(B + (i >> 5)) >> (i & 31)) & 1u; }
__device__ __forceinline__ int w_run_len(uint32_t word, int s) {  // consecutive set bits from bit s (set)
  const uint32_t inv = ~(word >> s);
  const int l = __ffs(inv) - 1;
  return (l < 0 || l > 32 - s) ? 32 - s : l;
}
__device__ __forceinline__ int w_run_start(uint32_t word, int x) {  // first bit of the run that contains set bit x
  const uint32_t below = ~word & ((x ? (1u << x) : 1u) - 1u);
  return below ? 32 - __clz(below) : 0;
}
__device__ __forceinline__ uint32_t w_run_mask(int s, int len) { return (len >= 32 ? 0xFFFFFFFFu : ((1u << len) - 1u)) << s; }

// find with path halving (a visited node is re-pointed at its grandparent with atomicMin: parents only ever become
// smaller ancestors, so concurrent walkers stay correct).  Without it a tall structure — a vertical wire, the paper
// background — becomes a chain as long as the image is high and every later find walks all of it.
__device__ __forceinline__ int uf_find_halve(int* L, int x) {
  while (true) {
    const int y = __ldcg(L + x);
    if (y == x) return x;
    const int z = __ldcg(L + y);
    if (z == y) return y;
    atomicMin(L + x, z);
    x = z;
  }
}
__device__ __forceinline__ void uf_union_halve(int* L, int a, int b) {
  bool done;
  do {
    a = uf_find_halve(L, a);
    b = uf_find_halve(L, b);
    if (a < b) {
      const int old = atomicMin(L + b, a);
      done = (old == b);
      b = old;
    } else if (b < a) {
      const int old = atomicMin(L + a, b);
      done = (old == a);
      a = old;
    } else {
      done = true;
    }
  } while (!done);
}

template <int CONN>
__device__ __forceinline__ void merge_class(int* L, uint32_t cur, uint32_t up, int pc, int pu) {
  // every run of `cur` against the runs of `up` it touches (CONN 8: diagonals inside the word included)
  while (cur) {
    const int st = __ffs(cur) - 1;
    const int len = w_run_len(cur, st);
    const uint32_t rm = w_run_mask(st, len);
    cur &= ~rm;
    uint32_t aw = rm;
    if (CONN == 8) aw |= (rm << 1) | (rm >> 1);
    uint32_t cand = aw & up;
    while (cand) {
      const int us = __ffs(cand) - 1;
      const int ust = w_run_start(up, us);
      cand &= ~w_run_mask(ust, w_run_len(up, ust));
      uf_union_halve(L, pc + st, pu + ust);
    }
  }
}

__global__ void __launch_bounds__(256) k_ccl_merge_words(const uint32_t* __restrict__ bits_all, int words_per_image,
                                                         int* __restrict__ Lall, int h, int w) {
  const int b = blockIdx.y;
  const int nwr = (w + 31) >> 5;
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= h * nwr) return;
  const int y = t / nwr, k = t - y * nwr;
  const uint32_t* B = bits_all + (size_t)b * words_per_image;
  int* L = Lall + (size_t)b * h * w;
  uint32_t vm;
  const uint32_t fg = row_word(B, w, y, k, &vm);
  const uint32_t bg = ~fg & vm;
  const int pc = y * w + 32 * k;
  // runs that continue across the 32-pixel segment boundary (k_ccl_init labels per segment)
  if (k > 0 && flat_bit(B, (uint32_t)pc - 1u) == (bool)(fg & 1u)) uf_union_halve(L, pc, pc - 1);
  if (y == 0) return;
  uint32_t vmu;
  const uint32_t upf = row_word(B, w, y - 1, k, &vmu);
  const uint32_t upb = ~upf & vm;
  const int pu = pc - w;
  merge_class<8>(L, fg, upf, pc, pu);
  merge_class<4>(L, bg, upb, pc, pu);
  // foreground diagonals across the segment boundary
  if (k > 0 && (fg & 1u) && flat_bit(B, (uint32_t)pu - 1u)) uf_union_halve(L, pc, pu - 1);
  if (32 * k + 32 < w && (fg >> 31) && flat_bit(B, (uint32_t)pu + 32u)) uf_union_halve(L, pc + 31, pu + 32);
}

__global__ void __launch_bounds__(256) k_ccl_flatten_words(const uint32_t* __restrict__ bits_all, int words_per_image,
                                                           int* __restrict__ Lall, int h, int w) {
  const int b = blockIdx.y;
  const int nwr = (w + 31) >> 5;
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= h * nwr) return;
  const int y = t / nwr, k = t - y * nwr;
  const uint32_t* B = bits_all + (size_t)b * words_per_image;
  int* L = Lall + (size_t)b * h * w;
  uint32_t vm;
  const uint32_t fg = row_word(B, w, y, k, &vm);
  const int pc = y * w + 32 * k;
#pragma unroll
  for (int cls = 0; cls < 2; cls++) {
    uint32_t cur = cls ? (~fg & vm) : fg;
    while (cur) {
      const int st = __ffs(cur) - 1;
      const int len = w_run_len(cur, st);
      cur &= ~w_run_mask(st, len);
      const int first = pc + st;
      const int r = uf_find<0>(L, first);
      // roots keep pointing at themselves, so concurrent finds stay valid; the run's pixels all named `first`
      if (r != first)
        for (int i = 0; i < len; i++) L[first + i] = r;
    }
  }
}

// background regions that touch the image frame (4-connected to cv2's implicit zero border)
__global__ void k_frame_flags(const uint8_t* __restrict__ bin, const int* __restrict__ Lall, uint8_t* __restrict__ frame,
                              int h, int w) {
  int b = blockIdx.y;
  int t = blockIdx.x * blockDim.x + threadIdx.x;
  int per = 2 * w + 2 * h;
  if (t >= per) return;
  int x, y;
  if (t < w) { x = t; y = 0; }
  else if (t < 2 * w) { x = t - w; y = h - 1; }
  else if (t < 2 * w + h) { x = 0; y = t - 2 * w; }
  else { x = w - 1; y = t - 2 * w - h; }
  size_t base = (size_t)b * h * w;
  int p = y * w + x;
  if (bin[base + p] == 0) frame[base + Lall[base + p]] = 1;
}

// External components (cv2 RETR_EXTERNAL) listed in cv2's order = descending raster index of the first pixel.
// One CTA per image.
#define CVB_MAX_ROWS 12000  // row counters live in dynamic shared memory (4 B per image row, below the 48 KB default)
__global__ void __launch_bounds__(1024) k_list_external(const uint8_t* __restrict__ bin, const int* __restrict__ Lall,
                                                        const uint8_t* __restrict__ frame, int h, int w,
                                                        int* __restrict__ cand, int max_external,
                                                        cv_image_result* __restrict__ results) {
  extern __shared__ int s_cnt[];  // [h]
  __shared__ int s_total;
  int b = blockIdx.x;
  size_t base = (size_t)b * h * w;
  const uint8_t* im = bin + base;
  const int* L = Lall + base;
  const uint8_t* fr = frame + base;
  int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  auto is_ext = [&](int x, int y) -> bool {
    if (x >= w) return false;
    int p = y * w + x;
    if (im[p] == 0 || L[p] != p) return false;
    return x == 0 || fr[L[p - 1]] != 0;
  };
  for (int y = warp; y < h; y += 32) {
    int cnt = 0;
    for (int x0 = 0; x0 < w; x0 += 32) cnt += __popc(__ballot_sync(0xffffffffu, is_ext(x0 + lane, y)));
    if (lane == 0) s_cnt[y] = cnt;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    int run = 0;
    for (int y = h - 1; y >= 0; y--) {
      int c = s_cnt[y];
      s_cnt[y] = run;  // number of externals in rows below
      run += c;
    }
    s_total = run;
  }
  __syncthreads();
  int nchunks = (w + 31) / 32;
  for (int y = warp; y < h; y += 32) {
    int run = s_cnt[y];
    for (int c = nchunks - 1; c >= 0; c--) {
      bool e = is_ext(c * 32 + lane, y);
      unsigned bits = __ballot_sync(0xffffffffu, e);
      if (e) {
        unsigned right = (lane == 31) ? 0u : (bits >> (lane + 1));
        int idx = run + __popc(right);
        if (idx < max_external) cand[(size_t)b * max_external + idx] = y * w + c * 32 + lane;
      }
      run += __popc(bits);
    }
  }
  if (threadIdx.x == 0) {
    int tot = s_total;
    if (tot > max_external) {
      results[b].status |= CV_STATUS_EXTERNAL_OVERFLOW;
      tot = max_external;
    }
    results[b].n_external = tot;
  }
}

struct CandStat {
  int32_t nverts, xmin, ymin, xmax, ymax, pad;
  long long a00, a01;
};

// a15: follow every external border once without storing it: vertex count, extents, polygon sums
__global__ void k_trace_stats(const uint8_t* __restrict__ bin, int h, int w, const int* __restrict__ cand,
                              int max_external, const cv_image_result* __restrict__ results,
                              CandStat* __restrict__ stats) {
  int b = blockIdx.y;
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= results[b].n_external) return;
  int p = cand[(size_t)b * max_external + i];
  ContourStats st = trace_outer_simple<uint8_t>(bin + (size_t)b * h * w, w, h, p % w, p / w, nullptr, 0);
  CandStat cs;
  cs.nverts = st.nverts;
  cs.xmin = st.xmin; cs.ymin = st.ymin; cs.xmax = st.xmax; cs.ymax = st.ymax; cs.pad = -1;
  cs.a00 = st.a00; cs.a01 = st.a01;
  stats[(size_t)b * max_external + i] = cs;
}

// Shared-memory variants of the two border walks: the walk is a chain of dependent neighbour probes (latency-bound,
// one thread per component), so each CTA first copies the image's BIT plane into shared memory and probes there
// (~25 cycles instead of an L2 round trip per probe).  TRACE_CTAS CTAs per image share its candidates.
constexpr int TRACE_CTAS = 8, TRACE_THREADS = 128;
constexpr int TRACE_PAD_WORDS = 4;  // one zero word in front of the plane (so pixel -1 is addressable) + 16-byte alignment

// The shared-memory plane: word 0..3 are zero, pixel i of the image is bit (i + 128) of the array, one extra zero word at
// the end (the funnel shift below reads the word after the one that holds pixel x+1).
__device__ __forceinline__ void load_bits_to_smem(uint32_t* sm, const uint32_t* __restrict__ g, int words) {
  if (threadIdx.x < TRACE_PAD_WORDS) sm[threadIdx.x] = 0u;
  if (threadIdx.x < 4) sm[TRACE_PAD_WORDS + words + threadIdx.x] = 0u;
  for (int i = threadIdx.x; i < words / 4; i += blockDim.x) ((uint4*)(sm + TRACE_PAD_WORDS))[i] = __ldg((const uint4*)g + i);
  __syncthreads();
}

// 8-neighbour occupancy of pixel (x, y) in OpenCV chain-code order (bit 0 = E, 1 = NE, 2 = N, 3 = NW, 4 = W, 5 = SW, 6 = S,
// 7 = SE; y grows downwards), pixels outside the image read as background.  Three independent row reads instead of up to
// eight dependent probes: the border walk is one long dependency chain, so the probes' latency is its speed.
__device__ __forceinline__ uint32_t neighbours8(const uint32_t* sm, int w, int h, int x, int y) {
  auto row3 = [&](int yy) -> uint32_t {  // bits (x-1, x, x+1) of row yy
    if ((unsigned)yy >= (unsigned)h) return 0u;
    const uint32_t i = (uint32_t)yy * (uint32_t)w + (uint32_t)x + (TRACE_PAD_WORDS * 32 - 1);
    const uint32_t lo = sm[i >> 5], hi = sm[(i >> 5) + 1];
    return __funnelshift_r(lo, hi, i & 31) & 7u;
  };
  uint32_t up = row3(y - 1), mid = row3(y), dn = row3(y + 1);
  const uint32_t keep = (x > 0 ? 1u : 0u) | 2u | (x + 1 < w ? 4u : 0u);  // flat bit order: row ends touch the next row
  up &= keep; mid &= keep; dn &= keep;
  return ((mid >> 2) & 1u) | (((up >> 2) & 1u) << 1) | (((up >> 1) & 1u) << 2) | ((up & 1u) << 3) | ((mid & 1u) << 4) |
         ((dn & 1u) << 5) | (((dn >> 1) & 1u) << 6) | (((dn >> 2) & 1u) << 7);
}
__device__ __forceinline__ int chain_dx(int s) { return (int)((0x901Au >> (2 * s)) & 3u) - 1; }  // 1 1 0 -1 -1 -1 0 1
__device__ __forceinline__ int chain_dy(int s) { return (int)((0xA901u >> (2 * s)) & 3u) - 1; }  // 0 -1 -1 -1 0 1 1 1

// A long border (the wire network of a schematic: tens of thousands of steps) is ONE dependent chain, so the first walk
// (statistics: vertex count, extents, polygon sums — needed by the area filter) cannot be split.  It drops a CHECKPOINT
// every TRACE_CK steps, though: the walker's full state.  The second walk, which only has to write the vertices of the
// contours that passed the filter, then starts one thread per checkpoint and each covers TRACE_CK steps — microseconds
// instead of another full-length chain.
constexpr int TRACE_CK = 256;
struct __align__(16) TraceCk {
  int32_t cand;     // candidate index within the image
  int32_t x, y;     // current pixel
  int32_t s;        // search direction state (-1: isolated pixel, the contour is its single vertex)
  int32_t prev_s;
  int32_t nverts;   // vertices emitted before this checkpoint
  int32_t pad0, pad1;
};

// trace_outer_fg (node_prims.cuh) on the shared-memory bit plane: same visiting order, same vertex selection, same sums.
// ck / n_ck / ck_cap: checkpoint pool of the image (nullptr: none are recorded); overflow sets *n_ck beyond ck_cap.
__device__ ContourStats trace_outer_bits(const uint32_t* sm, int w, int h, int x0, int y0, int32_t* out, int cap,
                                         TraceCk* ck, int* n_ck, int ck_cap, int cand) {
  ContourStats st;
  st.nverts = 0;
  st.xmin = st.xmax = x0;
  st.ymin = st.ymax = y0;
  st.a00 = st.a01 = st.a10 = 0;
  long long first_x = 0, first_y = 0, prev_x = 0, prev_y = 0;
  auto emit = [&](int x, int y) {
    if (out && st.nverts < cap) { out[2 * st.nverts] = x; out[2 * st.nverts + 1] = y; }
    if (st.nverts == 0) { first_x = x; first_y = y; }
    else {
      long long dxy = prev_x * (long long)y - (long long)x * prev_y;
      st.a00 += dxy;
      st.a01 += dxy * (prev_y + y);
      st.a10 += dxy * (prev_x + x);
    }
    prev_x = x; prev_y = y;
    st.xmin = min(st.xmin, x); st.xmax = max(st.xmax, x);
    st.ymin = min(st.ymin, y); st.ymax = max(st.ymax, y);
    st.nverts++;
  };
  auto checkpoint = [&](int x, int y, int s, int prev_s) {
    if (!ck) return;
    const int i = atomicAdd(n_ck, 1);
    if (i < ck_cap) {
      TraceCk c;
      c.cand = cand; c.x = x; c.y = y; c.s = s; c.prev_s = prev_s; c.nverts = st.nverts; c.pad0 = c.pad1 = 0;
      ck[i] = c;
    }
  };
  // first neighbour clockwise, starting just after West: directions 3, 2, 1, 0, 7, 6, 5, 4
  const uint32_t n0 = neighbours8(sm, w, h, x0, y0);
  if (n0 == 0u) {
    checkpoint(x0, y0, -1, 0);
    emit(x0, y0);  // isolated pixel
  } else {
    // rotate so that direction 3 becomes the top bit of a byte and scan downwards
    const uint32_t r0 = ((n0 | (n0 << 8)) >> 4) & 0xFFu;  // bit j = direction (j + 4) & 7
    int s = (31 - __clz(r0) + 4) & 7;                      // highest j first: j = 7 is direction 3
    const int x1 = x0 + chain_dx(s), y1 = y0 + chain_dy(s);
    int x3 = x0, y3 = y0;
    int prev_s = s ^ 4;
    for (int step = 0;; step++) {
      if ((step & (TRACE_CK - 1)) == 0) checkpoint(x3, y3, s, prev_s);
      // counter-clockwise search for the next border pixel, starting after the back-pointer
      const uint32_t nb = neighbours8(sm, w, h, x3, y3);
      const uint32_t rot = ((nb | (nb << 8)) >> (s + 1)) & 0xFFu;  // bit j = direction (s + 1 + j) & 7
      const int sn = rot ? ((s + 1 + (__ffs(rot) - 1)) & 7) : (s & 7);
      const int x4 = x3 + chain_dx(sn), y4 = y3 + chain_dy(sn);
      if (sn != prev_s) emit(x3, y3);
      prev_s = sn;
      if (x4 == x0 && y4 == y0 && x3 == x1 && y3 == y1) break;
      x3 = x4; y3 = y4;
      s = (sn + 4) & 7;
    }
  }
  if (st.nverts > 0) {
    long long dxy = prev_x * first_y - first_x * prev_y;
    st.a00 += dxy;
    st.a01 += dxy * (prev_y + first_y);
    st.a10 += dxy * (prev_x + first_x);
  }
  return st;
}

// second walk, one checkpoint: up to TRACE_CK steps from the recorded state, vertices written at their final positions
__device__ void trace_segment_bits(const uint32_t* sm, int w, int h, int x0, int y0, const TraceCk& c, int32_t* out, int cap) {
  if (c.s < 0) {  // isolated pixel
    if (cap > 0) { out[0] = x0; out[1] = y0; }
    return;
  }
  // the walk ends when it re-enters the start pixel's first step: (x1, y1) is the first neighbour found from (x0, y0)
  const uint32_t n0 = neighbours8(sm, w, h, x0, y0);
  const uint32_t r0 = ((n0 | (n0 << 8)) >> 4) & 0xFFu;
  const int s0 = (31 - __clz(r0) + 4) & 7;
  const int x1 = x0 + chain_dx(s0), y1 = y0 + chain_dy(s0);
  int x3 = c.x, y3 = c.y, s = c.s, prev_s = c.prev_s, nv = c.nverts;
  for (int step = 0; step < TRACE_CK; step++) {
    const uint32_t nb = neighbours8(sm, w, h, x3, y3);
    const uint32_t rot = ((nb | (nb << 8)) >> (s + 1)) & 0xFFu;
    const int sn = rot ? ((s + 1 + (__ffs(rot) - 1)) & 7) : (s & 7);
    const int x4 = x3 + chain_dx(sn), y4 = y3 + chain_dy(sn);
    if (sn != prev_s) {
      if (nv < cap) { out[2 * nv] = x3; out[2 * nv + 1] = y3; }
      nv++;
    }
    prev_s = sn;
    if (x4 == x0 && y4 == y0 && x3 == x1 && y3 == y1) break;
    x3 = x4; y3 = y4;
    s = (sn + 4) & 7;
  }
}

__global__ void __launch_bounds__(TRACE_THREADS) k_trace_stats_bits(const uint32_t* __restrict__ bits, int words_per_image,
                                                                    int h, int w, const int* __restrict__ cand,
                                                                    int max_external,
                                                                    const cv_image_result* __restrict__ results,
                                                                    CandStat* __restrict__ stats, TraceCk* __restrict__ ck,
                                                                    int* __restrict__ n_ck, int ck_cap) {
  extern __shared__ __align__(16) uint32_t s_bits[];
  const int b = blockIdx.y;
  const int n = results[b].n_external;
  if ((int)(blockIdx.x * TRACE_THREADS) >= n) return;  // whole CTA: nothing to walk
  load_bits_to_smem(s_bits, bits + (size_t)b * words_per_image, words_per_image);
  for (int i = blockIdx.x * TRACE_THREADS + threadIdx.x; i < n; i += TRACE_CTAS * TRACE_THREADS) {
    const int p = cand[(size_t)b * max_external + i];
    const ContourStats st = trace_outer_bits(s_bits, w, h, p % w, p / w, nullptr, 0, ck ? ck + (size_t)b * ck_cap : nullptr,
                                             n_ck + b, ck_cap, i);
    CandStat cs;
    cs.nverts = st.nverts;
    cs.xmin = st.xmin; cs.ymin = st.ymin; cs.xmax = st.xmax; cs.ymax = st.ymax; cs.pad = -1;
    cs.a00 = st.a00; cs.a01 = st.a01;
    stats[(size_t)b * max_external + i] = cs;
  }
}

// second walk from the checkpoints (stats[i].pad = id of the kept contour, -1 when the candidate was filtered out).
// Images whose checkpoint pool overflowed are left to the sequential kernel below (`overflowed` selects which).
__global__ void __launch_bounds__(TRACE_THREADS) k_trace_points_ck(const uint32_t* __restrict__ bits, int words_per_image,
                                                                   int h, int w, const int* __restrict__ cand,
                                                                   int max_external, const CandStat* __restrict__ stats,
                                                                   const cv_contour* __restrict__ contours, int max_contours,
                                                                   int max_points, const TraceCk* __restrict__ ck,
                                                                   const int* __restrict__ n_ck, int ck_cap,
                                                                   int32_t* __restrict__ points) {
  extern __shared__ __align__(16) uint32_t s_bits[];
  const int b = blockIdx.y;
  const int n = n_ck[b];
  if (n > ck_cap || (int)(blockIdx.x * TRACE_THREADS) >= n) return;
  load_bits_to_smem(s_bits, bits + (size_t)b * words_per_image, words_per_image);
  for (int i = blockIdx.x * TRACE_THREADS + threadIdx.x; i < n; i += gridDim.x * TRACE_THREADS) {
    const TraceCk c = ck[(size_t)b * ck_cap + i];
    const int id = stats[(size_t)b * max_external + c.cand].pad;
    if (id < 0) continue;
    const cv_contour& ct = contours[(size_t)b * max_contours + id];
    if (ct.offset + ct.nverts > max_points) continue;
    const int p = cand[(size_t)b * max_external + c.cand];
    trace_segment_bits(s_bits, w, h, p % w, p / w, c, points + ((size_t)b * max_points + ct.offset) * 2, ct.nverts);
  }
}

__global__ void __launch_bounds__(TRACE_THREADS) k_trace_points_bits(const uint32_t* __restrict__ bits, int words_per_image,
                                                                     int h, int w, const cv_contour* __restrict__ contours,
                                                                     int max_contours, int max_points,
                                                                     const cv_image_result* __restrict__ results,
                                                                     int32_t* __restrict__ points, const int* __restrict__ n_ck,
                                                                     int ck_cap) {
  extern __shared__ __align__(16) uint32_t s_bits[];
  const int b = blockIdx.y;
  if (n_ck && n_ck[b] <= ck_cap) return;  // this image was written from its checkpoints
  const int n = results[b].n_contours;
  if ((int)(blockIdx.x * TRACE_THREADS) >= n) return;
  load_bits_to_smem(s_bits, bits + (size_t)b * words_per_image, words_per_image);
  for (int i = blockIdx.x * TRACE_THREADS + threadIdx.x; i < n; i += TRACE_CTAS * TRACE_THREADS) {
    const cv_contour& ct = contours[(size_t)b * max_contours + i];
    if (ct.offset + ct.nverts > max_points) continue;
    trace_outer_bits(s_bits, w, h, ct.start_x, ct.start_y, points + ((size_t)b * max_points + ct.offset) * 2, ct.nverts,
                     nullptr, nullptr, 0, 0);
  }
}

// area filter (contourArea / (h*w) > 0.0004) + ids + point-pool offsets; one CTA per image
__global__ void __launch_bounds__(1024) k_filter_contours(const int* __restrict__ cand, CandStat* __restrict__ stats,
                                                          int max_external, int h, int w, double area_thr,
                                                          cv_contour* __restrict__ contours, int max_contours,
                                                          int max_points, cv_image_result* __restrict__ results) {
  __shared__ int s_wk[32], s_wp[32];
  __shared__ int s_run_k, s_run_p;
  int b = blockIdx.x;
  int nE = results[b].n_external;
  int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  if (threadIdx.x == 0) { s_run_k = 0; s_run_p = 0; }
  __syncthreads();
  int status = 0;
  for (int c0 = 0; c0 < nE; c0 += 1024) {
    int i = c0 + threadIdx.x;
    CandStat cs;
    int keep = 0, np = 0;
    if (i < nE) {
      cs = stats[(size_t)b * max_external + i];
      keep = area_passes(cs.a00, h, w, area_thr) ? 1 : 0;
      np = keep ? cs.nverts : 0;
    }
    // block exclusive scan of (keep, np)
    int k = keep, q = np;
    for (int o = 1; o < 32; o <<= 1) {
      int tk = __shfl_up_sync(0xffffffffu, k, o), tq = __shfl_up_sync(0xffffffffu, q, o);
      if (lane >= o) { k += tk; q += tq; }
    }
    if (lane == 31) { s_wk[warp] = k; s_wp[warp] = q; }
    __syncthreads();
    if (warp == 0) {
      int a = s_wk[lane], c = s_wp[lane];
      for (int o = 1; o < 32; o <<= 1) {
        int ta = __shfl_up_sync(0xffffffffu, a, o), tc = __shfl_up_sync(0xffffffffu, c, o);
        if (lane >= o) { a += ta; c += tc; }
      }
      s_wk[lane] = a; s_wp[lane] = c;
    }
    __syncthreads();
    int base_k = s_run_k + (warp ? s_wk[warp - 1] : 0);
    int base_p = s_run_p + (warp ? s_wp[warp - 1] : 0);
    int id = base_k + k - keep;
    int off = base_p + q - np;
    if (keep) {
      if (id < max_contours && off + np <= max_points) {
        int p = cand[(size_t)b * max_external + i];
        cv_contour ct;
        ct.start_x = p % w; ct.start_y = p / w;
        ct.offset = off; ct.nverts = cs.nverts;
        ct.xmin = cs.xmin; ct.ymin = cs.ymin; ct.xmax = cs.xmax; ct.ymax = cs.ymax;
        ct.a00 = cs.a00; ct.a01 = cs.a01;
        ct.new_id = -1; ct.ncomp = 0; ct.has_source = 0;
        int cy;
        ct.centroid_y = centroid_y(cs.a00, cs.a01, &cy) ? cy : INT_MIN;
        contours[(size_t)b * max_contours + id] = ct;
        stats[(size_t)b * max_external + i].pad = id;  // read by the checkpointed second walk
      } else {
        status |= (id >= max_contours) ? CV_STATUS_CONTOUR_OVERFLOW : CV_STATUS_POINT_OVERFLOW;
      }
    }
    __syncthreads();
    if (threadIdx.x == 1023) { s_run_k = base_k + k; s_run_p = base_p + q; }
    __syncthreads();
  }
  if (status) atomicOr(&results[b].status, status);
  __syncthreads();
  if (threadIdx.x == 0) {
    int nk = s_run_k, npnt = s_run_p;
    // on overflow the stored prefix is still self-consistent: ids/offsets are monotone
    if (nk > max_contours) nk = max_contours;
    results[b].n_contours = nk;
    results[b].n_points = npnt > max_points ? max_points : npnt;
  }
}

// a15: second walk of the kept borders, now writing the SIMPLE vertex lists
__global__ void k_trace_points(const uint8_t* __restrict__ bin, int h, int w, const cv_contour* __restrict__ contours,
                               int max_contours, int max_points, const cv_image_result* __restrict__ results,
                               int32_t* __restrict__ points) {
  int b = blockIdx.y;
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= results[b].n_contours) return;
  const cv_contour& ct = contours[(size_t)b * max_contours + i];
  if (ct.offset + ct.nverts > max_points) return;
  trace_outer_simple<uint8_t>(bin + (size_t)b * h * w, w, h, ct.start_x, ct.start_y,
                              points + ((size_t)b * max_points + ct.offset) * 2, ct.nverts);
}

// ------------------------------------------------------------------------------------------------
// a16: box <-> contour contact.  One CTA per image; each warp takes one box of the current chunk of 32
// boxes and walks the contours in id order (AABB pre-test, then a ballot scan of the vertex list for the
// first vertex that is "near").  Hits are staged in shared memory and flushed in box-major order so the
// pair list has the reference's append order.
// ------------------------------------------------------------------------------------------------
#define CVB_HITS_PER_BOX 64
// contours of one image in id order against one box (a whole warp): AABB pre-test, then a ballot scan of the vertex list for
// the first "near" vertex.  emit(hit index, pair) is called by every lane for each hit; returns the number of hits.
template <class Emit>
__device__ __forceinline__ int contact_scan_box(const cv_box& bx, int bi, const cv_contour* __restrict__ cts, int nK,
                                                const int32_t* __restrict__ pts, int lane, Emit emit) {
  int nh = 0;
  for (int k = 0; k < nK; k++) {
    const cv_contour& ct = cts[k];
    int cx = ct.xmin, cy = ct.ymin, cxm = ct.xmax + 1, cym = ct.ymax + 1;  // rect x+w, y+h
    if (bx.rxmax < cx || bx.rxmin > cxm || bx.rymax < cy || bx.rymin > cym) continue;
    int n = ct.nverts;
    const int32_t* cp = pts + (size_t)ct.offset * 2;
    int hit = -1, hx = 0, hy = 0;
    for (int v0 = 0; v0 < n && hit < 0; v0 += 32) {
      int v = v0 + lane;
      bool near = false;
      int px = 0, py = 0;
      if (v < n) {
        px = cp[2 * v]; py = cp[2 * v + 1];
        near = point_near_box(px, py, bx.rxmin, bx.rymin, bx.rxmax, bx.rymax, bx.thresh);
      }
      unsigned m = __ballot_sync(0xffffffffu, near);
      if (m) {
        int src = __ffs(m) - 1;
        hit = v0 + src;
        hx = __shfl_sync(0xffffffffu, px, src);
        hy = __shfl_sync(0xffffffffu, py, src);
      }
    }
    if (hit >= 0) {
      cv_pair pr; pr.contour = k; pr.box = bi; pr.px = hx; pr.py = hy;
      emit(nh, pr);
      nh++;
    }
  }
  return nh;
}

__global__ void __launch_bounds__(1024) k_contact(const cv_box* __restrict__ boxes, const int32_t* __restrict__ box_offsets,
                                                  const cv_contour* __restrict__ contours, int max_contours,
                                                  const int32_t* __restrict__ points, int max_points,
                                                  cv_pair* __restrict__ pairs, int max_pairs,
                                                  cv_image_result* __restrict__ results) {
  __shared__ cv_pair s_hits[32][CVB_HITS_PER_BOX];
  __shared__ int s_nh[32];
  __shared__ int s_base[33];
  __shared__ int s_written;
  int b = blockIdx.x;
  int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  int bo = box_offsets[b], nb = box_offsets[b + 1] - bo;
  int nK = results[b].n_contours;
  const cv_contour* cts = contours + (size_t)b * max_contours;
  const int32_t* pts = points + (size_t)b * max_points * 2;
  cv_pair* out = pairs + (size_t)b * max_pairs;
  if (threadIdx.x == 0) s_written = 0;
  __syncthreads();
  int status = 0;
  for (int c0 = 0; c0 < nb; c0 += 32) {
    int bi = c0 + warp;
    int nh = 0;
    cv_box bx;
    bool scan = false;
    if (bi < nb) {
      bx = boxes[bo + bi];
      scan = (bx.flags & CV_BOX_IS_COMPONENT) != 0;
    }
    if (scan)
      nh = contact_scan_box(bx, bi, cts, nK, pts, lane, [&](int idx, const cv_pair& pr) {
        if (idx < CVB_HITS_PER_BOX && lane == 0) s_hits[warp][idx] = pr;
      });
    if (lane == 0) s_nh[warp] = nh;
    __syncthreads();
    if (threadIdx.x == 0) {
      int run = s_written;
      for (int i = 0; i < 32; i++) { s_base[i] = run; run += s_nh[i]; }
      s_base[32] = run;
    }
    __syncthreads();
    const int base = s_base[warp];
    for (int j = lane; j < min(nh, CVB_HITS_PER_BOX); j += 32) {
      int dst = base + j;
      if (dst < max_pairs) out[dst] = s_hits[warp][j];
      else status |= CV_STATUS_PAIR_OVERFLOW;
    }
    if (nh > CVB_HITS_PER_BOX) {
      // a box that touches more contours than the staging rows hold (noisy masks): the same walk once more, hits beyond the
      // staged ones go straight to their slots
      contact_scan_box(bx, bi, cts, nK, pts, lane, [&](int idx, const cv_pair& pr) {
        if (idx >= CVB_HITS_PER_BOX) {
          if (base + idx < max_pairs) { if (lane == 0) out[base + idx] = pr; }
          else status |= CV_STATUS_PAIR_OVERFLOW;
        }
      });
    }
    __syncthreads();
    if (threadIdx.x == 0) s_written = s_base[32];
    __syncthreads();
  }
  if (status) atomicOr(&results[b].status, status);
  __syncthreads();
  if (threadIdx.x == 0) results[b].n_pairs = min(s_written, max_pairs);
}

// ------------------------------------------------------------------------------------------------
// a16 tail + a17: uid de-duplication, ground selection, renumbering (circuit_analyzer.py:1418-1582).
// Tiny and sequential: one thread per image.  The pair list is compacted in place (dropped duplicates).
// ------------------------------------------------------------------------------------------------
__global__ void k_assemble(const cv_box* __restrict__ boxes, const int32_t* __restrict__ box_offsets,
                           cv_contour* __restrict__ contours, int max_contours, cv_pair* __restrict__ pairs,
                           int max_pairs, cv_image_result* __restrict__ results, int B) {
  int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= B) return;
  cv_contour* cts = contours + (size_t)b * max_contours;
  cv_pair* pr = pairs + (size_t)b * max_pairs;
  const cv_box* bx = boxes + box_offsets[b];
  int nK = results[b].n_contours, nP = results[b].n_pairs;
  // de-duplicate (node, persistent_uid): keep the first occurrence in append order
  int out = 0;
  for (int i = 0; i < nP; i++) {
    cv_pair p = pr[i];
    int g = bx[p.box].uid_group;
    bool dup = false;
    if (g != p.box) {  // only boxes that share a uid with an earlier box can be duplicates
      for (int j = 0; j < out && !dup; j++) dup = (pr[j].contour == p.contour) && (bx[pr[j].box].uid_group == g);
    }
    if (!dup) {
      pr[out++] = p;
      cts[p.contour].ncomp++;
      if (bx[p.box].flags & CV_BOX_IS_SOURCE) cts[p.contour].has_source = 1;
    }
  }
  results[b].n_pairs = out;
  int n_valid = 0, max_conn = 0;
  for (int k = 0; k < nK; k++)
    if (cts[k].ncomp > 0) { n_valid++; max_conn = max(max_conn, cts[k].ncomp); }
  results[b].ground = -1;
  results[b].n_nodes = 0;
  if (n_valid == 0) return;
  // ground: source-connected valid node with the largest centroid_y (stable => lowest id on ties) :1475-1497
  int ground = -1;
  long long best = LLONG_MIN;
  bool any_src = false;
  for (int k = 0; k < nK; k++)
    if (cts[k].ncomp > 0 && cts[k].has_source) {
      long long cy = cts[k].centroid_y == INT_MIN ? (LLONG_MIN + 1) : (long long)cts[k].centroid_y;
      if (!any_src || cy > best) { best = cy; ground = k; }
      any_src = true;
    }
  if (!any_src) {  // :1499-1524 fallback among the nodes with the most components
    int n_max = 0, first_max = -1;
    for (int k = 0; k < nK; k++)
      if (cts[k].ncomp == max_conn) { if (first_max < 0) first_max = k; n_max++; }
    if (n_max > 1) {
      bool any = false;
      for (int k = 0; k < nK; k++)
        if (cts[k].ncomp == max_conn) {
          long long cy = cts[k].centroid_y == INT_MIN ? (LLONG_MIN + 1) : (long long)cts[k].centroid_y;
          if (!any || cy > best) { best = cy; ground = k; }
          any = true;
        }
    } else {
      ground = first_max;
    }
  }
  results[b].ground = ground;
  cts[ground].new_id = 0;
  int nxt = 1;
  for (int k = 0; k < nK; k++) {  // :1556-1568
    if (k == ground || cts[k].ncomp == 0) continue;
    if (cts[k].ncomp >= 2 || (nxt == 1 && n_valid == 2)) cts[k].new_id = nxt++;
  }
  results[b].n_nodes = nxt;
}

__global__ void k_init_results(cv_image_result* results, unsigned long long* sums, int* n_ck, int B) {
  int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= B) return;
  n_ck[b] = 0;
  cv_image_result r;
  r.n_external = r.n_contours = r.n_points = r.n_pairs = r.n_nodes = 0;
  r.ground = -1; r.inverted = 0; r.status = 0;
  results[b] = r;
  sums[b] = 0;
}


// ------------------------------------------------------------------------------------------------
// Terminal reclassification (circuit_analyzer.py:2217-2310): page -> grey -> adaptive threshold ->
// box masking -> external contours at NATIVE resolution -> per-terminal count of contours with a "near" vertex.
// ------------------------------------------------------------------------------------------------
// grey value of the RGB page with the reference's channel swap; thread = 4 pixels (12 B in, 4 B out)
__global__ void __launch_bounds__(256) k_gray_page(const uint8_t* __restrict__ rgb, uint8_t* __restrict__ gray, size_t n_px,
                                                   int aligned) {
  size_t q = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  size_t p0 = q * 4;
  if (p0 >= n_px) return;
  if (aligned && p0 + 4 <= n_px) {
    const uint32_t* src = (const uint32_t*)(rgb + p0 * 3);
    uint32_t a = __ldg(src), b = __ldg(src + 1), c = __ldg(src + 2);
    // bytes: a = R0 G0 B0 R1 | b = G1 B1 R2 G2 | c = B2 R3 G3 B3
    int g0 = gray_of_rgb_page(a & 255, (a >> 8) & 255, (a >> 16) & 255);
    int g1 = gray_of_rgb_page(a >> 24, b & 255, (b >> 8) & 255);
    int g2 = gray_of_rgb_page((b >> 16) & 255, b >> 24, c & 255);
    int g3 = gray_of_rgb_page((c >> 8) & 255, (c >> 16) & 255, c >> 24);
    *(uint32_t*)(gray + p0) = (uint32_t)g0 | ((uint32_t)g1 << 8) | ((uint32_t)g2 << 16) | ((uint32_t)g3 << 24);
  } else {
    for (size_t p = p0; p < p0 + 4 && p < n_px; p++)
      gray[p] = (uint8_t)gray_of_rgb_page(rgb[3 * p], rgb[3 * p + 1], rgb[3 * p + 2]);
  }
}

// cv2.adaptiveThreshold(MEAN_C, BINARY_INV, 31, 21): 64 x 32 output tile, replicate-border halo of 15 in shared memory,
// separable sliding-window sums (rows: thread = (row, 16-column segment); columns: thread = (column, 8-row segment)).
#define ADT_W 64
#define ADT_H 32
__global__ void __launch_bounds__(256) k_adaptive31(const uint8_t* __restrict__ gray, uint8_t* __restrict__ out, int H, int W) {
  __shared__ uint8_t s_in[ADT_H + 30][ADT_W + 32];   // 94 used columns, padded row pitch
  __shared__ uint16_t s_row[ADT_H + 30][ADT_W + 2];  // horizontal 31-sums (<= 7905)
  const int b = blockIdx.z, x0 = blockIdx.x * ADT_W, y0 = blockIdx.y * ADT_H;
  const uint8_t* im = gray + (size_t)b * H * W;
  for (int t = threadIdx.x; t < (ADT_H + 30) * (ADT_W + 30); t += 256) {
    int ly = t / (ADT_W + 30), lx = t - ly * (ADT_W + 30);
    int gy = min(max(y0 + ly - 15, 0), H - 1), gx = min(max(x0 + lx - 15, 0), W - 1);
    s_in[ly][lx] = im[(size_t)gy * W + gx];
  }
  __syncthreads();
  for (int t = threadIdx.x; t < (ADT_H + 30) * 4; t += 256) {
    int ly = t >> 2, seg = (t & 3) * 16;
    int acc = 0;
#pragma unroll
    for (int k = 0; k < 31; k++) acc += s_in[ly][seg + k];
    s_row[ly][seg] = (uint16_t)acc;
#pragma unroll
    for (int j = 1; j < 16; j++) {
      acc += (int)s_in[ly][seg + j + 30] - (int)s_in[ly][seg + j - 1];
      s_row[ly][seg + j] = (uint16_t)acc;
    }
  }
  __syncthreads();
  {
    int lx = threadIdx.x & 63, seg = (threadIdx.x >> 6) * 8;
    int gx = x0 + lx;
    int acc = 0;
#pragma unroll
    for (int k = 0; k < 31; k++) acc += s_row[seg + k][lx];
#pragma unroll
    for (int j = 0; j < 8; j++) {
      if (j) acc += (int)s_row[seg + j + 30][lx] - (int)s_row[seg + j - 1][lx];
      int gy = y0 + seg + j;
      if (gx < W && gy < H) out[(size_t)b * H * W + (size_t)gy * W + gx] = adaptive_inv_31_21(s_in[seg + j + 15][lx + 15], acc);
    }
  }
}

__global__ void __launch_bounds__(256) k_image_sum(const uint8_t* __restrict__ img, int n_per_image,
                                                   unsigned long long* __restrict__ sums) {
  int b = blockIdx.y;
  const uint8_t* im = img + (size_t)b * n_per_image;
  unsigned int local = 0;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n_per_image; i += gridDim.x * blockDim.x) local += im[i];
  for (int o = 16; o; o >>= 1) local += __shfl_xor_sync(0xffffffffu, local, o);
  if ((threadIdx.x & 31) == 0 && local) atomicAdd(&sums[b], (unsigned long long)local);
}

// counts[box] = number of contours (id order irrelevant) with at least one vertex "near" the box in the reference's
// sense (:811-846, threshold 10, NO bounding-rectangle pre-test here — :2279-2285), -1 for boxes that are not terminals.
// One CTA per image, one warp per box; lanes scan a contour's vertices 32 at a time.
__global__ void __launch_bounds__(1024) k_terminal_counts(const cv_box* __restrict__ boxes, const int32_t* __restrict__ box_offsets,
                                                          const cv_contour* __restrict__ contours, int max_contours,
                                                          const int32_t* __restrict__ points, int max_points,
                                                          const cv_image_result* __restrict__ results, int thresh,
                                                          int32_t* __restrict__ counts) {
  int b = blockIdx.x;
  int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  int bo = box_offsets[b], nb = box_offsets[b + 1] - bo;
  int nK = results[b].n_contours;
  const cv_contour* cts = contours + (size_t)b * max_contours;
  const int32_t* pts = points + (size_t)b * max_points * 2;
  for (int bi = warp; bi < nb; bi += 32) {
    cv_box bx = boxes[bo + bi];
    int n = -1;
    if (bx.flags & CV_BOX_IS_TERMINAL) {
      n = 0;
      for (int k = 0; k < nK; k++) {
        const cv_contour& ct = cts[k];
        const int32_t* cp = pts + (size_t)ct.offset * 2;
        bool hit = false;
        for (int v0 = 0; v0 < ct.nverts && !hit; v0 += 32) {
          int v = v0 + lane;
          bool near = v < ct.nverts && point_near_box(cp[2 * v], cp[2 * v + 1], bx.xmin, bx.ymin, bx.xmax, bx.ymax, thresh);
          hit = __ballot_sync(0xffffffffu, near) != 0u;
        }
        n += hit ? 1 : 0;
      }
    }
    if (lane == 0) counts[bo + bi] = n;
  }
}

}  // namespace cvb

using namespace cvb;

static inline size_t align256(size_t x) { return (x + 255) & ~(size_t)255; }

static cv_nodes_caps resolve_caps(const cv_nodes_caps* caps) {
  cv_nodes_caps c;
  c.max_external = 32768; c.max_contours = 2560; c.max_points = 262144; c.max_pairs = 8192;
  if (caps) {
    if (caps->max_external > 0) c.max_external = caps->max_external;
    if (caps->max_contours > 0) c.max_contours = caps->max_contours;
    if (caps->max_points > 0) c.max_points = caps->max_points;
    if (caps->max_pairs > 0) c.max_pairs = caps->max_pairs;
  }
  return c;
}

extern "C" int cv_nodes_resized_width(int H, int W) {
  if (H <= 0 || W <= 0) return 0;
  double aspect = (double)W / (double)H;   // circuit_analyzer.py:800
  return (int)(600.0 * aspect);            // :803 int(new_height * aspect_ratio)
}

struct NodesWs {
  uint8_t* enh_raw; uint8_t* bin; uint8_t* frame; int* labels; int* cand; CandStat* stats; unsigned long long* sums;
  uint32_t* bits; int words;  // bit plane of `bin`, `words` 32-bit words per image (multiple of 4)
  TraceCk* ck; int* n_ck; int ck_cap;  // checkpoints of the first border walk
  size_t total;
};

static inline int bit_words(size_t n_pixels) { return (int)((((n_pixels + 31) / 32) + 3) & ~(size_t)3); }
// one checkpoint per candidate + one per TRACE_CK border steps (a border has at most ~4 steps per pixel; in practice a
// few per cent of that) — an image that needs more falls back to the sequential second walk
static inline int trace_ck_cap(int max_external, size_t n_pixels) { return max_external + (int)(n_pixels / 64) + 64; }
constexpr size_t TRACE_SMEM_MAX = 200 * 1024;  // bit planes up to 1.6 Mpixel stay in shared memory

// the two border walks, from shared memory when the bit plane fits
static int launch_traces_stats(const uint32_t* bits, int words, const uint8_t* bin, int h, int w, const int* cand,
                               const cv_nodes_caps& c, int B, const cv_image_result* results, CandStat* stats, TraceCk* ck,
                               int* n_ck, int ck_cap, cudaStream_t st) {
  const size_t smem = ((size_t)words + TRACE_PAD_WORDS + 4) * 4;
  if (smem <= TRACE_SMEM_MAX) {
    static std::atomic<unsigned long long> attr{0};
    if (cvb_once_per_device(attr))
      CVB_CHECK(cudaFuncSetAttribute(k_trace_stats_bits, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)TRACE_SMEM_MAX));
    CVB_LAUNCH(k_trace_stats_bits, dim3(TRACE_CTAS, B), dim3(TRACE_THREADS), smem, st, bits, words, h, w, cand,
               c.max_external, results, stats, ck, n_ck, ck_cap);
  } else {
    CVB_LAUNCH(k_trace_stats, dim3((c.max_external + 63) / 64, B), dim3(64), 0, st, bin, h, w, cand, c.max_external, results,
               stats);
  }
  return CV_OK;
}
static int launch_traces_points(const uint32_t* bits, int words, const uint8_t* bin, int h, int w, const int* cand,
                                const CandStat* stats, const cv_contour* contours, const cv_nodes_caps& c, int B,
                                const cv_image_result* results, int32_t* points, const TraceCk* ck, const int* n_ck,
                                int ck_cap, cudaStream_t st) {
  const size_t smem = ((size_t)words + TRACE_PAD_WORDS + 4) * 4;
  if (smem <= TRACE_SMEM_MAX) {
    static std::atomic<unsigned long long> attr{0};
    if (cvb_once_per_device(attr)) {
      CVB_CHECK(cudaFuncSetAttribute(k_trace_points_bits, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)TRACE_SMEM_MAX));
      CVB_CHECK(cudaFuncSetAttribute(k_trace_points_ck, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)TRACE_SMEM_MAX));
    }
    // every checkpoint is an independent walk of at most TRACE_CK steps ...
    CVB_LAUNCH(k_trace_points_ck, dim3(2 * TRACE_CTAS, B), dim3(TRACE_THREADS), smem, st, bits, words, h, w, cand,
               c.max_external, stats, contours, c.max_contours, c.max_points, ck, n_ck, ck_cap, points);
    // ... and the images whose checkpoint pool overflowed (if any) repeat the full-length walk
    CVB_LAUNCH(k_trace_points_bits, dim3(TRACE_CTAS, B), dim3(TRACE_THREADS), smem, st, bits, words, h, w, contours,
               c.max_contours, c.max_points, results, points, n_ck, ck_cap);
  } else {
    CVB_LAUNCH(k_trace_points, dim3((c.max_contours + 63) / 64, B), dim3(64), 0, st, bin, h, w, contours, c.max_contours,
               c.max_points, results, points);
  }
  return CV_OK;
}

static NodesWs carve_nodes_ws(void* base, int B, int h, int w, const cv_nodes_caps& c) {
  NodesWs ws;
  size_t n = (size_t)B * h * w, off = 0;
  char* p = (char*)base;
  ws.enh_raw = (uint8_t*)(p + off); off += align256(n);
  ws.bin = (uint8_t*)(p + off); off += align256(n);
  ws.frame = (uint8_t*)(p + off); off += align256(n);
  ws.labels = (int*)(p + off); off += align256(n * 4);
  ws.cand = (int*)(p + off); off += align256((size_t)B * c.max_external * 4);
  ws.stats = (CandStat*)(p + off); off += align256((size_t)B * c.max_external * sizeof(CandStat));
  ws.sums = (unsigned long long*)(p + off); off += align256((size_t)B * 8);
  ws.words = bit_words((size_t)h * w);
  ws.bits = (uint32_t*)(p + off); off += align256((size_t)B * ws.words * 4);
  ws.ck_cap = trace_ck_cap(c.max_external, (size_t)h * w);
  ws.ck = (TraceCk*)(p + off); off += align256((size_t)B * ws.ck_cap * sizeof(TraceCk));
  ws.n_ck = (int*)(p + off); off += align256((size_t)B * 4);
  ws.total = off;
  return ws;
}

extern "C" size_t cv_nodes_workspace_bytes(int B, int H, int W, const cv_nodes_caps* caps) {
  int w = cv_nodes_resized_width(H, W);
  if (B <= 0 || w <= 0) return 0;
  return carve_nodes_ws(nullptr, B, 600, w, resolve_caps(caps)).total;
}

extern "C" int cv_nodes_analyze(const uint8_t* masks, int B, int H, int W, const cv_box* boxes,
                                const int32_t* box_offsets, int max_boxes_per_image, uint8_t* emptied,
                                uint8_t* resized, uint8_t* enhanced, cv_contour* contours, int32_t* points,
                                cv_pair* pairs, cv_image_result* results, const cv_nodes_caps* caps, void* workspace,
                                size_t workspace_bytes, void* stream_) {
  cvb_reset_launches();
  if (!masks || !emptied || !resized || !enhanced || !contours || !points || !pairs || !results || !box_offsets ||
      !workspace || B <= 0 || H <= 0 || W <= 0)
    return cvb_fail(CV_ERR_INVALID, "cv_nodes_analyze: null pointer or non-positive size");
  const int h = 600;
  const int w = cv_nodes_resized_width(H, W);
  if (w <= 0) return cvb_fail(CV_ERR_INVALID, "cv_nodes_analyze: resized width is 0 (extreme aspect ratio)");
  if (h > CVB_MAX_ROWS) return cvb_fail(CV_ERR_INVALID, "cv_nodes_analyze: internal row limit");
  if ((long long)h * w >= (1ll << 30)) return cvb_fail(CV_ERR_INVALID, "cv_nodes_analyze: resized image too large");
  cv_nodes_caps c = resolve_caps(caps);
  NodesWs ws = carve_nodes_ws(workspace, B, h, w, c);
  if (workspace_bytes < ws.total) return cvb_fail(CV_ERR_INVALID, "cv_nodes_analyze: workspace too small");
  cudaStream_t st = (cudaStream_t)stream_;
  const size_t n_full = (size_t)B * H * W;
  const int n_small = h * w;

  CVB_LAUNCH(k_init_results, dim3((B + 127) / 128), dim3(128), 0, st, results, ws.sums, ws.n_ck, B);
  {  // a12
    bool al = (((uintptr_t)masks | (uintptr_t)emptied) & 15) == 0;
    size_t n16 = al ? n_full / 16 : 0;
    size_t want = (n16 + 255) / 256;
    int grid = (int)(want < 1 ? 1 : (want > 148 * 16 ? 148 * 16 : want));
    cvb_next_work(2.0 * (double)n_full);  // read mask + write emptied
    CVB_LAUNCH(k_copy16, dim3(grid), dim3(256), 0, st, (const uint4*)masks, (uint4*)emptied, n16, masks, emptied, n_full);
    if (max_boxes_per_image > 0 && boxes)
      CVB_LAUNCH(k_zero_boxes, dim3(max_boxes_per_image, B), dim3(256), 0, st, emptied, H, W, boxes, box_offsets, 0);
  }
  // a13: each destination row reads two source rows (whole 32-byte sectors when down-scaling < 32x) + writes itself
  cvb_next_work((double)B * h * (2.0 * W + w));
  CVB_LAUNCH(k_resize, dim3((w + 31) / 32, (h + 7) / 8, B), dim3(32, 8), 0, st, emptied, H, W, resized, h, w);
  // a14 + a15 prelude
  cvb_next_work(2.0 * (double)B * n_small);
  CVB_LAUNCH(k_enhance, dim3((w + ENH_T - 1) / ENH_T, (h + ENH_T - 1) / ENH_T, B), dim3(256), 0, st, resized, ws.enh_raw,
             h, w, ws.sums);
  CVB_LAUNCH(k_binarize, dim3((ws.words * 32 + 255) / 256, B), dim3(256), 0, st, ws.enh_raw, enhanced, ws.bin, n_small,
             ws.sums, results, ws.bits, ws.words);
  // a15: labelling (8-connected foreground, 4-connected background) -> external components -> borders
  CVB_CHECK(cudaMemsetAsync(ws.frame, 0, (size_t)B * n_small, st));
  dim3 cg((w + 31) / 32, (h + 7) / 8, B), cb(32, 8);
  cvb_next_work(5.0 * (double)B * n_small);  // 1 B/px mask read + 4 B/px label write
  CVB_LAUNCH((k_ccl_init<0, true>), cg, cb, 0, st, ws.bin, ws.labels, h, w);
  {
    const int n_words = h * ((w + 31) / 32);
    cvb_next_work(0.25 * (double)B * n_small);  // two rows of bits per word
    CVB_LAUNCH(k_ccl_merge_words, dim3((n_words + 255) / 256, B), dim3(256), 0, st, ws.bits, ws.words, ws.labels, h, w);
    cvb_next_work(8.0 * (double)B * n_small);  // label read + write
    CVB_LAUNCH(k_ccl_flatten_words, dim3((n_words + 255) / 256, B), dim3(256), 0, st, ws.bits, ws.words, ws.labels, h, w);
  }
  CVB_LAUNCH(k_frame_flags, dim3((2 * w + 2 * h + 255) / 256, B), dim3(256), 0, st, ws.bin, ws.labels, ws.frame, h, w);
  CVB_LAUNCH(k_list_external, dim3(B), dim3(1024), (size_t)h * sizeof(int), st, ws.bin, ws.labels, ws.frame, h, w, ws.cand,
             c.max_external, results);
  if (int rc = launch_traces_stats(ws.bits, ws.words, ws.bin, h, w, ws.cand, c, B, results, ws.stats, ws.ck, ws.n_ck, ws.ck_cap, st)) return rc;
  CVB_LAUNCH(k_filter_contours, dim3(B), dim3(1024), 0, st, ws.cand, ws.stats, c.max_external, h, w, 0.0004, contours,
             c.max_contours, c.max_points, results);
  if (int rc = launch_traces_points(ws.bits, ws.words, ws.bin, h, w, ws.cand, ws.stats, contours, c, B, results, points, ws.ck, ws.n_ck, ws.ck_cap, st)) return rc;
  // a16, a17
  if (boxes) {
    CVB_LAUNCH(k_contact, dim3(B), dim3(1024), 0, st, boxes, box_offsets, contours, c.max_contours, points, c.max_points,
               pairs, c.max_pairs, results);
    CVB_LAUNCH(k_assemble, dim3((B + 31) / 32), dim3(32), 0, st, boxes, box_offsets, contours, c.max_contours, pairs,
               c.max_pairs, results, B);
  }
  return CV_OK;
}

// ------------------------------------------------------------------ result compaction before the device->host copy
// The result tables are fixed-capacity (max_contours x 64 B + max_pairs x 16 B + max_points x 8 B per image, mostly empty).
// cv_nodes_pack gathers the used prefixes of all B images into one contiguous blob so that only the bytes that carry
// information cross PCIe:   header[b] = {byte offset of image b's section, n_contours, n_pairs, n_points} (int64 x 4),
// header[B].offset = total bytes;  section = contours[nK] | pairs[nP] | points[nPts] (x, y int32 pairs), 16-byte aligned.
__global__ void __launch_bounds__(1024) k_pack_offsets(const cv_image_result* __restrict__ results, int B,
                                                       long long* __restrict__ header) {
  __shared__ long long part[1024];
  // B <= a few thousand: one block, serial chunks of 1024 with a running base
  long long base = 0;
  for (int b0 = 0; b0 < B; b0 += 1024) {
    const int b = b0 + threadIdx.x;
    long long sz = 0, nK = 0, nP = 0, nPt = 0;
    if (b < B) {
      nK = results[b].n_contours; nP = results[b].n_pairs; nPt = results[b].n_points;
      sz = (nK * (long long)sizeof(cv_contour) + nP * (long long)sizeof(cv_pair) + nPt * 8 + 15) & ~15ll;
    }
    part[threadIdx.x] = sz;
    __syncthreads();
    for (int d = 1; d < 1024; d <<= 1) {
      long long v = threadIdx.x >= d ? part[threadIdx.x - d] : 0;
      __syncthreads();
      part[threadIdx.x] += v;
      __syncthreads();
    }
    if (b < B) {
      header[4 * b + 0] = base + part[threadIdx.x] - sz;
      header[4 * b + 1] = nK; header[4 * b + 2] = nP; header[4 * b + 3] = nPt;
    }
    const long long tot = part[1023];
    __syncthreads();
    base += tot;
  }
  if (threadIdx.x == 0) { header[4 * B] = base; header[4 * B + 1] = header[4 * B + 2] = header[4 * B + 3] = 0; }
}

__global__ void __launch_bounds__(256) k_pack_copy(const cv_contour* __restrict__ contours, int max_contours,
                                                   const int32_t* __restrict__ points, int max_points,
                                                   const cv_pair* __restrict__ pairs, int max_pairs,
                                                   const long long* __restrict__ header, uint8_t* __restrict__ blob,
                                                   long long blob_cap) {
  const int b = blockIdx.y;
  const long long off = header[4 * b], nK = header[4 * b + 1], nP = header[4 * b + 2], nPt = header[4 * b + 3];
  const long long w0 = nK * (long long)(sizeof(cv_contour) / 8), w1 = w0 + nP * (long long)(sizeof(cv_pair) / 8), w2 = w1 + nPt;
  if (off + w2 * 8 > blob_cap) return;  // caller checks header[B] against the capacity
  const unsigned long long* sK = (const unsigned long long*)(contours + (size_t)b * max_contours);
  const unsigned long long* sP = (const unsigned long long*)(pairs + (size_t)b * max_pairs);
  const unsigned long long* sT = (const unsigned long long*)(points + (size_t)b * max_points * 2);
  unsigned long long* dst = (unsigned long long*)(blob + off);
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < w2; i += (long long)gridDim.x * blockDim.x)
    dst[i] = i < w0 ? sK[i] : (i < w1 ? sP[i - w0] : sT[i - w1]);
}

extern "C" int cv_nodes_pack(const cv_contour* contours, const int32_t* points, const cv_pair* pairs,
                             const cv_image_result* results, int B, const cv_nodes_caps* caps, long long* header,
                             uint8_t* blob, long long blob_bytes, void* stream) {
  cvb_reset_launches();
  if (!contours || !points || !pairs || !results || !header || !blob || B <= 0 || blob_bytes <= 0)
    return cvb_fail(CV_ERR_INVALID, "cv_nodes_pack: bad argument");
  cv_nodes_caps c = resolve_caps(caps);
  cudaStream_t st = (cudaStream_t)stream;
  CVB_LAUNCH(k_pack_offsets, dim3(1), dim3(1024), 0, st, results, B, header);
  cvb_next_work(0.0);
  CVB_LAUNCH(k_pack_copy, dim3(16, B), dim3(256), 0, st, contours, c.max_contours, points, c.max_points, pairs, c.max_pairs,
             header, blob, (long long)blob_bytes);
  return CV_OK;
}

// ------------------------------------------------------------------ terminal reclassification entry point
struct TermWs {
  uint8_t* gray; uint8_t* bin; uint8_t* frame; int* labels; int* cand; CandStat* stats; unsigned long long* sums;
  uint32_t* bits; int words;
  TraceCk* ck; int* n_ck; int ck_cap;
  size_t total;
};

static TermWs carve_term_ws(void* base, int B, int H, int W, const cv_nodes_caps& c) {
  TermWs ws;
  size_t n = (size_t)B * H * W, off = 0;
  char* p = (char*)base;
  ws.gray = (uint8_t*)(p + off); off += align256(n);
  ws.bin = (uint8_t*)(p + off); off += align256(n);
  ws.frame = (uint8_t*)(p + off); off += align256(n);
  ws.labels = (int*)(p + off); off += align256(n * 4);
  ws.cand = (int*)(p + off); off += align256((size_t)B * c.max_external * 4);
  ws.stats = (CandStat*)(p + off); off += align256((size_t)B * c.max_external * sizeof(CandStat));
  ws.sums = (unsigned long long*)(p + off); off += align256((size_t)B * 8);
  ws.words = bit_words((size_t)H * W);
  ws.bits = (uint32_t*)(p + off); off += align256((size_t)B * ws.words * 4);
  ws.ck_cap = trace_ck_cap(c.max_external, (size_t)H * W);
  ws.ck = (TraceCk*)(p + off); off += align256((size_t)B * ws.ck_cap * sizeof(TraceCk));
  ws.n_ck = (int*)(p + off); off += align256((size_t)B * 4);
  ws.total = off;
  return ws;
}

extern "C" size_t cv_terminals_workspace_bytes(int B, int H, int W, const cv_nodes_caps* caps) {
  if (B <= 0 || H <= 0 || W <= 0) return 0;
  return carve_term_ws(nullptr, B, H, W, resolve_caps(caps)).total;
}

extern "C" int cv_terminals_analyze(const uint8_t* pages_rgb, int B, int H, int W, const cv_box* boxes,
                                    const int32_t* box_offsets, int max_boxes_per_image, int n_boxes_total,
                                    uint8_t* wire_mask, int32_t* box_counts, cv_contour* contours, int32_t* points,
                                    cv_image_result* results, const cv_nodes_caps* caps, void* workspace,
                                    size_t workspace_bytes, void* stream_) {
  cvb_reset_launches();
  if (!pages_rgb || !wire_mask || !contours || !points || !results || !box_offsets || !workspace || B <= 0 || H <= 0 || W <= 0)
    return cvb_fail(CV_ERR_INVALID, "cv_terminals_analyze: null pointer or non-positive size");
  if (n_boxes_total > 0 && (!boxes || !box_counts)) return cvb_fail(CV_ERR_INVALID, "cv_terminals_analyze: boxes without a count array");
  if (H > CVB_MAX_ROWS) return cvb_fail(CV_ERR_INVALID, "cv_terminals_analyze: more than 12000 image rows");
  if ((long long)H * W >= (1ll << 30)) return cvb_fail(CV_ERR_INVALID, "cv_terminals_analyze: image too large");
  cv_nodes_caps c = resolve_caps(caps);
  TermWs ws = carve_term_ws(workspace, B, H, W, c);
  if (workspace_bytes < ws.total) return cvb_fail(CV_ERR_INVALID, "cv_terminals_analyze: workspace too small");
  cudaStream_t st = (cudaStream_t)stream_;
  const int n_img = H * W;
  const size_t n_px = (size_t)B * n_img;

  CVB_LAUNCH(k_init_results, dim3((B + 127) / 128), dim3(128), 0, st, results, ws.sums, ws.n_ck, B);
  // segment_circuit (:313-319 via :2231-2234)
  cvb_next_work(4.0 * (double)n_px);
  CVB_LAUNCH(k_gray_page, dim3((unsigned)((n_px / 4 + 256) / 256)), dim3(256), 0, st, pages_rgb, ws.gray, n_px,
             (int)(((uintptr_t)pages_rgb & 3) == 0 && ((uintptr_t)ws.gray & 3) == 0));
  cvb_next_work(2.0 * (double)n_px);
  CVB_LAUNCH(k_adaptive31, dim3((W + ADT_W - 1) / ADT_W, (H + ADT_H - 1) / ADT_H, B), dim3(256), 0, st, ws.gray, wire_mask, H, W);
  // :2238-2249 box masking with NumPy slice semantics
  if (max_boxes_per_image > 0 && boxes)
    CVB_LAUNCH(k_zero_boxes, dim3(max_boxes_per_image, B), dim3(256), 0, st, wire_mask, H, W, boxes, box_offsets, 1);
  // get_contours(prelim_wire_mask, 0.0001) at native resolution (:2252, :388-412)
  cvb_next_work((double)n_px);
  CVB_LAUNCH(k_image_sum, dim3(min((n_img + 255) / 256, 148 * 8), B), dim3(256), 0, st, wire_mask, n_img, ws.sums);
  CVB_LAUNCH(k_binarize, dim3((ws.words * 32 + 255) / 256, B), dim3(256), 0, st, wire_mask, (uint8_t*)nullptr, ws.bin, n_img,
             ws.sums, results, ws.bits, ws.words);
  CVB_CHECK(cudaMemsetAsync(ws.frame, 0, n_px, st));
  dim3 cg((W + 31) / 32, (H + 7) / 8, B), cb(32, 8);
  cvb_next_work(5.0 * (double)n_px);
  CVB_LAUNCH((k_ccl_init<0, true>), cg, cb, 0, st, ws.bin, ws.labels, H, W);
  {
    const int n_words = H * ((W + 31) / 32);
    cvb_next_work(0.25 * (double)n_px);
    CVB_LAUNCH(k_ccl_merge_words, dim3((n_words + 255) / 256, B), dim3(256), 0, st, ws.bits, ws.words, ws.labels, H, W);
    cvb_next_work(8.0 * (double)n_px);
    CVB_LAUNCH(k_ccl_flatten_words, dim3((n_words + 255) / 256, B), dim3(256), 0, st, ws.bits, ws.words, ws.labels, H, W);
  }
  CVB_LAUNCH(k_frame_flags, dim3((2 * W + 2 * H + 255) / 256, B), dim3(256), 0, st, ws.bin, ws.labels, ws.frame, H, W);
  CVB_LAUNCH(k_list_external, dim3(B), dim3(1024), (size_t)H * sizeof(int), st, ws.bin, ws.labels, ws.frame, H, W, ws.cand,
             c.max_external, results);
  if (int rc = launch_traces_stats(ws.bits, ws.words, ws.bin, H, W, ws.cand, c, B, results, ws.stats, ws.ck, ws.n_ck, ws.ck_cap, st)) return rc;
  CVB_LAUNCH(k_filter_contours, dim3(B), dim3(1024), 0, st, ws.cand, ws.stats, c.max_external, H, W, 0.0001, contours,
             c.max_contours, c.max_points, results);
  if (int rc = launch_traces_points(ws.bits, ws.words, ws.bin, H, W, ws.cand, ws.stats, contours, c, B, results, points, ws.ck, ws.n_ck, ws.ck_cap, st)) return rc;
  // :2270-2287 contacts of the terminals, threshold 10
  if (n_boxes_total > 0)
    CVB_LAUNCH(k_terminal_counts, dim3(B), dim3(1024), 0, st, boxes, box_offsets, contours, c.max_contours, points,
               c.max_points, results, 10, box_counts);
  return CV_OK;
}

// native-resolution CCL (cv_ccl_label, BASELINE cfg 4) lives in ccl_tiles.cu
