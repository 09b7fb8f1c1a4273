// Native-resolution connected-component labelling (cv_ccl_label, BASELINE cfg 4) — tiled, bit-parallel union-find.
//
// Not a reference code path: the oracle is cv2.connectedComponents up to renaming; labels are
// canonical: labels[p] = 1 + min linear index of p's component, 0 for background.
//
// Traffic plan (the path is HBM-bound; algorithmic bytes = 1 B/px mask read + 4 B/px label write), three passes:
//   pass A  k_ccl_scan       : one WARP per tile of 1024 x 64 pixels, lane = one 32-pixel word column, rows top to bottom.
//                              Reads the mask once (1 B/px), writes the bit plane (1 bit/px), first[word] = label of the word's
//                              first sub-run (4 B per 32 px, coalesced), and parent entries IN the label image only for roots
//                              and for the second and later sub-runs of a word.  No block barriers, no shared atomics: labels
//                              flow down the rows through two shared-memory rows of slots per warp, ballot / shuffle resolve
//                              the runs that span several words, and only real merges (two different labels meeting) touch the
//                              global union-find.
//   pass B  k_ccl_seams      : unions across the tile seams (rows y = 64 k: 1.6 % of the image; columns x = 1024 k), one launch.
//   pass C  k_ccl_write      : the label image is written exactly once (4 B/px): root by pointer chasing from first[word],
//                              16-byte stores.  It also counts the roots.
// Total ~ 5.3 B/px; 0.60 of the HBM copy peak at 256 images per call, 0.45 at 16 (tails of the two big passes).  The algorithm is modelled lane by lane in oracle/ccl_scan_model.py (held to cv2 on the CPU by
// tests/test_ccl_model_cpu.py); round 1's block-level shared-memory union-find + fix-up pass (6+ B/px, 72 % issue-bound,
// 35 % of its stalls at __syncthreads) reached 0.32 of the HBM copy peak.
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdlib.h>

#include "../../include/cv_b200.h"
#include "common.cuh"

namespace cvb {


// ---- global union-find on the label image (parent of x = L[x] - 1; 0 = background)
__device__ __forceinline__ int gfind(const int* L, int x) {
  int y = __ldcg(L + x) - 1;
  while (y != x) {
    x = y;
    y = __ldcg(L + x) - 1;
  }
  return x;
}
// find with path halving: a visited node is re-pointed at its grandparent (atomicMin: parents only become smaller
// ancestors, concurrent walkers stay correct), so the seam unions do not build tile-to-tile chains
__device__ __forceinline__ int gfind_halve(int* L, int x) {
  while (true) {
    const int y = __ldcg(L + x) - 1;
    if (y == x) return x;
    const int z = __ldcg(L + y) - 1;
    if (z == y) return y;
    atomicMin(L + x, z + 1);
    x = z;
  }
}
__device__ __forceinline__ void gunion(int* L, int a, int b) {
  // first hop without a write: a and b may be ordinary pixels, and pass C relies on ordinary pixels keeping the
  // tile-local root pass A gave them — only root entries may be re-pointed
  a = __ldcg(L + a) - 1;
  b = __ldcg(L + b) - 1;
  bool done;
  do {
    a = gfind_halve(L, a);
    b = gfind_halve(L, b);
    if (a < b) {
      int old = atomicMin(L + b, a + 1) - 1;
      done = (old == b);
      b = old;
    } else if (b < a) {
      int old = atomicMin(L + a, b + 1) - 1;
      done = (old == a);
      a = old;
    } else {
      done = true;
    }
  } while (!done);
}

// union of two ELEMENTS (run-start pixels with a parent entry)
__device__ __forceinline__ void gunion_roots(int* L, int a, int b) { gunion(L, a, b); }
// union of two LABELS (elements that carry a parent entry: roots, former roots), no first hop
__device__ __forceinline__ void gunion_labels(int* L, int a, int b) {
  bool done;
  do {
    a = gfind_halve(L, a);
    b = gfind_halve(L, b);
    if (a < b) {
      int old = atomicMin(L + b, a + 1) - 1;
      done = (old == b);
      b = old;
    } else if (b < a) {
      int old = atomicMin(L + a, b + 1) - 1;
      done = (old == a);
      a = old;
    } else {
      done = true;
    }
  } while (!done);
}

// start of the sub-run of `word` that contains set bit x
__device__ __forceinline__ int run_start(uint32_t word, int x) {
  uint32_t below = ~word & ((x ? (1u << x) : 1u) - 1u);
  return below ? 32 - __clz(below) : 0;
}
// number of consecutive set bits of `word` starting at bit s (bit s is set)
__device__ __forceinline__ int run_len(uint32_t word, int s) {
  uint32_t inv = ~(word >> s);
  int l = __ffs(inv) - 1;  // inv == 0 -> -1
  return (l < 0 || l > 32 - s) ? 32 - s : l;
}
__device__ __forceinline__ uint32_t run_mask(int s, int len) {
  return (len >= 32 ? 0xFFFFFFFFu : ((1u << len) - 1u)) << s;
}

// 32 mask bytes of row y starting at x -> bit i set iff pixel x+i is non-zero (out-of-image pixels are 0)
__device__ __forceinline__ uint32_t load_word(const uint8_t* __restrict__ im, int H, int W, int y, int x, bool vec_ok) {
  if (y >= H || x >= W) return 0u;
  const uint8_t* p = im + (size_t)y * W + x;
  uint32_t w = 0;
  if (vec_ok && x + 32 <= W) {
    const uint4 a = __ldg((const uint4*)p), b = __ldg((const uint4*)p + 1);
    const uint32_t v[8] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w};
#pragma unroll
    for (int k = 0; k < 8; k++) {
      // MSB of each byte := (byte != 0); the multiply gathers the four MSBs (bits 7,15,23,31) into bits 28..31
      uint32_t m = (((v[k] & 0x7F7F7F7Fu) + 0x7F7F7F7Fu) | v[k]) & 0x80808080u;
      w |= ((m * 0x00204081u) >> 28) << (4 * k);
    }
  } else {
    const int n = min(32, W - x);
    for (int i = 0; i < n; i++) w |= (uint32_t)(p[i] != 0) << i;
  }
  return w;
}

// The same in two steps, so that the 32 raw bytes of a row can wait in registers while the row before it is processed (the
// conversion is the first use of the loaded data: done right after the load, it exposes the whole memory latency every row).
// `fast` lanes (whole word inside the image, 16-byte aligned rows) hold the raw bytes; the others hold the finished word in a.x.
struct RawWord {
  uint4 a, b;
};
__device__ __forceinline__ RawWord load_raw(const uint8_t* __restrict__ im, int H, int W, int y, int x, bool fast) {
  RawWord r;
  if (fast) {
    const uint4* p = (const uint4*)(im + (size_t)y * W + x);
    r.a = __ldg(p);
    r.b = __ldg(p + 1);
  } else {
    r.a = make_uint4(load_word(im, H, W, y, x, false), 0u, 0u, 0u);
    r.b = make_uint4(0u, 0u, 0u, 0u);
  }
  return r;
}
__device__ __forceinline__ uint32_t raw_to_word(const RawWord& r, bool fast) {
  if (!fast) return r.a.x;
  const uint32_t v[8] = {r.a.x, r.a.y, r.a.z, r.a.w, r.b.x, r.b.y, r.b.z, r.b.w};
  uint32_t w = 0;
#pragma unroll
  for (int k = 0; k < 8; k++) {
    uint32_t m = (((v[k] & 0x7F7F7F7Fu) + 0x7F7F7F7Fu) | v[k]) & 0x80808080u;
    w |= ((m * 0x00204081u) >> 28) << (4 * k);
  }
  return w;
}

// ================================================================================================
// pass A: warp-sequential scan.  One warp per tile of 1024 x CS_TH pixels, lane = one 32-pixel word column, rows top to
// bottom.  Model: oracle/ccl_scan_model.py (pass_a_tile) — same names, same order.
//
// Element of the union-find = a maximal horizontal run inside the tile row, identified by the linear index of its first
// pixel; its parent entry lives in the label image AT that pixel (L[start] = parent + 1), nowhere else.  A run takes the
// label (ancestor index) of a run it touches in the row above, so vertical structures never build chains; a run that
// touches nothing becomes a root; touching two different labels is a (rare) union with global atomics.
// lab[parity][lane][start bit >> 1] = label of the word-local sub-run starting at that bit (two runs cannot start at
// adjacent bits), for the previous and the current row.
// ================================================================================================
constexpr int CS_TH = 64;        // rows per tile (32: twice the seam work, 0.582 -> 0.597 of the HBM peak at 128 images; 128: the
                                 // scan's tail grows, 0.594, and 16 images no longer fill the GPU)
constexpr int CS_WARPS = 8;      // tiles (stacked vertically) per CTA (4 / 6 / 8 measured the same at 128 images)
constexpr int CS_LSTRIDE = 16;   // 16 slots per lane; the slot index is XOR-swizzled with the lane so that the 32 lanes hit 32 banks
                                 // (4 KB per warp: 7 CTAs = 56 warps per SM, the whole 16 x 4096^2 problem in one wave)
constexpr int CS_INF = 0x7FFFFFFF;

__device__ __forceinline__ int cs_slot(int lane, int slot) { return lane * CS_LSTRIDE + (slot ^ ((lane >> 1) & 15)); }

__device__ __forceinline__ void cs_combine(int* L, int& cd, int t) {
  if (cd == CS_INF) {
    cd = t;
  } else if (t != cd) {
    gunion_roots(L, cd, t);
    cd = min(cd, t);
  }
}

template <int CONN>
__global__ void __launch_bounds__(CS_WARPS * 32) k_ccl_scan(const uint8_t* __restrict__ masks, int* __restrict__ labels,
                                                           uint32_t* __restrict__ bits_all, int* __restrict__ first_all, int H,
                                                           int W, int wpr, int vec_ok) {
  __shared__ int lab_s[CS_WARPS][2][32 * CS_LSTRIDE];
  const int warp = threadIdx.x >> 5, c = threadIdx.x & 31;
  const int b = blockIdx.z, x0 = blockIdx.x * 1024, y0 = (blockIdx.y * CS_WARPS + warp) * CS_TH;
  if (y0 >= H) return;
  const uint8_t* im = masks + (size_t)b * H * W;
  int* L = labels + (size_t)b * H * W;
  uint32_t* bits = bits_all + (size_t)b * H * wpr;
  int* first = first_all + (size_t)b * H * wpr;
  const int wc = blockIdx.x * 32 + c;
  const bool valid = wc < wpr;
  int(*lab)[32 * CS_LSTRIDE] = lab_s[warp];
  uint32_t up = 0, upl = 0, upr = 0;
  const int rows = min(CS_TH, H - y0);
  // the rows of the tile are walked one after the other (each depends on the labels of the one above): row r + 1 waits
  // converted, row r + 2 as raw bytes whose load was issued a whole row step before its first use.  (An up-front
  // prefetch.global.L2 of the whole tile helped the 16-image batch of round 1 and costs 1-2 % at 128+ images: removed.)
  const int xw = x0 + c * 32;
  const bool fast = valid && vec_ok && xw + 32 <= W;
  uint32_t w_next = valid ? raw_to_word(load_raw(im, H, W, y0, xw, fast), fast) : 0u;
  RawWord raw2;
  raw2.a = raw2.b = make_uint4(0u, 0u, 0u, 0u);
  if (valid && rows > 1) raw2 = load_raw(im, H, W, y0 + 1, xw, fast);
  for (int r = 0; r < rows; r++) {
    const int y = y0 + r, par = r & 1;
    const uint32_t w = w_next;
    w_next = (valid && r + 1 < rows) ? raw_to_word(raw2, fast) : 0u;     // loaded one row step ago
    if (valid && r + 2 < rows) raw2 = load_raw(im, H, W, y + 2, xw, fast);  // first use one row step from now
    if (valid) bits[(size_t)y * wpr + wc] = w;
    uint32_t wl = __shfl_up_sync(0xffffffffu, w, 1), wr = __shfl_down_sync(0xffffffffu, w, 1);
    if (c == 0) wl = 0u;
    if (c == 31) wr = 0u;
    const bool cin = (w & 1u) && (wl >> 31), cout = (w >> 31) && (wr & 1u);
    const bool full = w == 0xFFFFFFFFu, brk = !(full && cin);
    const uint32_t Bm = __ballot_sync(0xffffffffu, brk);
    const int origin = cin ? 31 - __clz(Bm & ((1u << c) - 1u)) : c;  // lane 0 never has cin, so the mask is non-empty
    const int rowbase = y * W + x0 + c * 32;
    int* lp = lab[par ^ 1];  // previous row
    int* lc = lab[par];
    int ch = CS_INF, ct = CS_INF, tail_st = 0;
    const int first_st = __ffs(w) - 1;  // start bit of the word's first sub-run: its label goes to first[], not to the label image
    for (uint32_t s = w & ~(w << 1); s; s &= s - 1) {
      const int st = __ffs(s) - 1;
      const int len = run_len(w, st);
      const uint32_t rm = run_mask(st, len);
      int cd = CS_INF;
      if (r > 0) {
        uint32_t aw = rm;
        if (CONN == 8) aw |= (rm << 1) | (rm >> 1);
        uint32_t ov = aw & up;
        while (ov) {
          const int u = __ffs(ov) - 1;
          const int us = run_start(up, u);
          ov &= ~run_mask(us, run_len(up, us));
          cs_combine(L, cd, lp[cs_slot(c, us >> 1)]);
        }
        if (CONN == 8) {
          if ((rm & 1u) && (upl >> 31)) cs_combine(L, cd, lp[cs_slot(c - 1, run_start(upl, 31) >> 1)]);
          if ((rm >> 31) && (upr & 1u)) cs_combine(L, cd, lp[cs_slot(c + 1, 0)]);
        }
      }
      const bool is_head = st == 0 && cin, is_tail = (rm >> 31) && cout;
      if (is_head) {
        ch = cd;  // joins the run that started in lane `origin`
      } else if (is_tail) {
        ct = cd;
        tail_st = st;
      } else {
        const int m = cd == CS_INF ? rowbase + st : cd;  // nothing touched: a root, named by its first pixel
        // parent entries in the label image: roots, and sub-runs after the first of their word (the first one is found through
        // first[]): a 4-byte store per run dirtied a 32-byte sector of the label image per run (read-modify-write in DRAM)
        if (cd == CS_INF || st != first_st) L[rowbase + st] = m + 1;
        lc[cs_slot(c, st >> 1)] = m;
      }
    }
    // runs that span words: minimum over the portions, unions between portions that disagree.  Most rows have none (only
    // horizontal structures cross word boundaries): one ballot skips the whole step.
    if (__ballot_sync(0xffffffffu, cin) != 0u) {
      // segmented backward min-scan over the followers of each origin: the groups are contiguous lane ranges, so five shuffle
      // steps do it (__match_any + __reduce_min on sub-masks compiled to a loop over ~30 singleton groups per row)
      const int key = cin ? origin : 32 + c;
      int gmin = cin ? ch : CS_INF;
#pragma unroll
      for (int d = 1; d < 32; d <<= 1) {
        const int v2 = __shfl_down_sync(0xffffffffu, gmin, d), k2 = __shfl_down_sync(0xffffffffu, key, d);
        if (c + d < 32 && k2 == key) gmin = min(gmin, v2);
      }
      const int from_right = __shfl_down_sync(0xffffffffu, gmin, 1);  // first follower: minimum over the whole group
      const bool tail_org = cout && brk;  // my tail run starts in my word and continues to the right
      int my_m = 0;
      const int my_start = rowbase + tail_st;
      if (tail_org) {
        const int tot = min(ct, from_right);
        my_m = tot == CS_INF ? my_start : tot;
        if (tot == CS_INF || tail_st != first_st) L[my_start] = my_m + 1;  // before the union: a new root needs its entry
        if (ct != CS_INF && ct != my_m) gunion_roots(L, ct, my_m);
        lc[cs_slot(c, tail_st >> 1)] = my_m;
      }
      const int mo = __shfl_sync(0xffffffffu, my_m, origin);
      if (cin) {
        if (ch != CS_INF && ch != mo) gunion_roots(L, ch, mo);
        lc[cs_slot(c, 0)] = mo;
      }
    }
    // first[word] = label of the word's first sub-run (an element of the same component as every pixel of that sub-run): pass C
    // starts its root chase there — a coalesced 4-byte read per word instead of a 32-byte sector of the label image per word
    if (valid && w) first[(size_t)y * wpr + wc] = lc[cs_slot(c, (__ffs(w) - 1) >> 1)];
    __syncwarp();  // label slots and parent entries of this row are visible to every lane before the next row reads them
    up = w; upl = wl; upr = wr;
  }
}

// label (an element of the same component that carries a parent entry) of the sub-run of word (y, wc) that contains set bit
// `bit`: first[] for the first sub-run of a word (whether it starts there or entered from the left), one hop through the
// sub-run's own parent entry otherwise
__device__ __forceinline__ int cs_label_of(const int* L, const uint32_t* __restrict__ bits, const int* __restrict__ first, int W, int wpr,
                                           int y, int wc, int bit) {
  const uint32_t wv = bits[(size_t)y * wpr + wc];
  const int st = run_start(wv, bit);
  if (st == __ffs(wv) - 1) return first[(size_t)y * wpr + wc];
  return __ldcg(L + (size_t)y * W + wc * 32 + st) - 1;
}

// pass B, horizontal seams: rows y = k CS_TH against the row above, one 32-pixel word per thread
template <int CONN>
__device__ __forceinline__ void ccl_seam_row_word(int* L, const uint32_t* __restrict__ bits, const int* __restrict__ first, int H, int W,
                                                  int wpr, int y, int wc) {
  if (y >= H || wc >= wpr) return;
  const uint32_t wv = bits[(size_t)y * wpr + wc];
  if (!wv) return;
  const uint32_t uv = bits[(size_t)(y - 1) * wpr + wc];
  bool ul = false, ur = false;
  if (CONN == 8) {
    ul = wc > 0 && (bits[(size_t)(y - 1) * wpr + wc - 1] >> 31);
    ur = wc + 1 < wpr && (bits[(size_t)(y - 1) * wpr + wc + 1] & 1u);
  }
  if (!uv && !ul && !ur) return;
  for (uint32_t s = wv & ~(wv << 1); s; s &= s - 1) {
    const int st = __ffs(s) - 1;
    const int len = run_len(wv, st);
    const uint32_t rm = run_mask(st, len);
    uint32_t aw = rm;
    if (CONN == 8) aw |= (rm << 1) | (rm >> 1);
    uint32_t ov = aw & uv;
    const bool dl = CONN == 8 && (rm & 1u) && ul, dr = CONN == 8 && (rm >> 31) && ur;
    if (!ov && !dl && !dr) continue;
    const int me = cs_label_of(L, bits, first, W, wpr, y, wc, st);
    while (ov) {
      const int u = __ffs(ov) - 1;
      const int us = run_start(uv, u);
      ov &= ~run_mask(us, run_len(uv, us));
      gunion_labels(L, me, cs_label_of(L, bits, first, W, wpr, y - 1, wc, us));
    }
    if (dl) gunion_labels(L, me, cs_label_of(L, bits, first, W, wpr, y - 1, wc - 1, 31));
    if (dr) gunion_labels(L, me, cs_label_of(L, bits, first, W, wpr, y - 1, wc + 1, 0));
  }
}

// pass B, vertical seams: columns x = k 1024 against column x - 1, one row per thread
template <int CONN>
__device__ __forceinline__ void ccl_seam_col_pixel(int* L, const uint32_t* __restrict__ bits, const int* __restrict__ first, int H, int W,
                                                   int wpr, long long t) {
  const int n_seams = (W - 1) / 1024;
  if (t >= (long long)n_seams * H) return;
  const int wc = (int)(t / H + 1) * 32, y = (int)(t % H);
  const bool cur = bits[(size_t)y * wpr + wc] & 1u, left = bits[(size_t)y * wpr + wc - 1] >> 31;
  if (cur && left) {
    gunion_labels(L, cs_label_of(L, bits, first, W, wpr, y, wc, 0), cs_label_of(L, bits, first, W, wpr, y, wc - 1, 31));
  } else if (CONN == 8 && cur) {
    const int me = cs_label_of(L, bits, first, W, wpr, y, wc, 0);
    if (y > 0 && (bits[(size_t)(y - 1) * wpr + wc - 1] >> 31)) gunion_labels(L, me, cs_label_of(L, bits, first, W, wpr, y - 1, wc - 1, 31));
    if (y + 1 < H && (bits[(size_t)(y + 1) * wpr + wc - 1] >> 31)) gunion_labels(L, me, cs_label_of(L, bits, first, W, wpr, y + 1, wc - 1, 31));
  } else if (CONN == 8 && left) {
    const int me = cs_label_of(L, bits, first, W, wpr, y, wc - 1, 31);
    if (y > 0 && (bits[(size_t)(y - 1) * wpr + wc] & 1u)) gunion_labels(L, me, cs_label_of(L, bits, first, W, wpr, y - 1, wc, 0));
    if (y + 1 < H && (bits[(size_t)(y + 1) * wpr + wc] & 1u)) gunion_labels(L, me, cs_label_of(L, bits, first, W, wpr, y + 1, wc, 0));
  }
}

// pass B in one launch: blocks [0, row_blocks) take the horizontal seams (block = 128 words of one seam row), the blocks after
// them the vertical seams (128 seam pixels each) — the two sets of unions are independent, so they run side by side
template <int CONN>
__global__ void __launch_bounds__(128) k_ccl_seams(int* __restrict__ labels, const uint32_t* __restrict__ bits_all,
                                                    const int* __restrict__ first_all, int H, int W, int wpr, int row_blocks_x,
                                                    int row_blocks) {
  const int b = blockIdx.z;
  int* L = labels + (size_t)b * H * W;
  const uint32_t* bits = bits_all + (size_t)b * H * wpr;
  const int* first = first_all + (size_t)b * H * wpr;
  const int bx = blockIdx.x;
  if (bx < row_blocks) {
    const int sy = bx / row_blocks_x, sx = bx - sy * row_blocks_x;
    ccl_seam_row_word<CONN>(L, bits, first, H, W, wpr, (sy + 1) * CS_TH, sx * 128 + (int)threadIdx.x);
  } else {
    ccl_seam_col_pixel<CONN>(L, bits, first, H, W, wpr, (long long)(bx - row_blocks) * 128 + threadIdx.x);
  }
}

// pass B', flatten: a vertical structure that crosses k tile seams leaves a chain root_k -> root_k-1 -> ... -> root_0 (union by
// minimum index links every tile's root under the one above it), and pass C would walk it from every pixel.  All those
// roots are runs of the FIRST row of a tile, so re-pointing the run starts of the seam rows at their final root (3 % of the
// rows, path-halving finds) bounds every later chase by a few hops.
__global__ void __launch_bounds__(128) k_ccl_flatten_rows(int* __restrict__ labels, const uint32_t* __restrict__ bits_all,
                                                          int* __restrict__ first_all, int H, int W, int wpr) {
  const int b = blockIdx.z, y = (blockIdx.y + 1) * CS_TH, wc = blockIdx.x * 128 + threadIdx.x;
  if (y >= H || wc >= wpr) return;
  int* L = labels + (size_t)b * H * W;
  const uint32_t* bits = bits_all + (size_t)b * H * wpr;
  int* first = first_all + (size_t)b * H * wpr;
  const uint32_t wv = bits[(size_t)y * wpr + wc];
  if (!wv) return;
  {  // first sub-run of the word: its label lives in first[]
    const int lab = first[(size_t)y * wpr + wc];
    const int r = gfind_halve(L, lab);
    if (r != lab) first[(size_t)y * wpr + wc] = r;
  }
  uint32_t starts = wv & ~(wv << 1);
  starts &= starts - 1;  // further sub-runs: their own parent entries
  for (; starts; starts &= starts - 1) {
    const int s = y * W + wc * 32 + __ffs(starts) - 1;
    const int r = gfind_halve(L, s);
    if (r != s) atomicMin(L + s, r + 1);
  }
}

// pass C: the label image is written ONCE.  One warp per row segment of 1024 pixels (= one tile row), lane = one 32-pixel word:
// the lane resolves the label of its word's first sub-run (first[word] from pass A, root by pointer chasing), then the warp writes the 4 KB of labels cooperatively — per 16-byte store a lane fetches the word
// and the label of the word it is writing by shuffle, so a store instruction covers 512 contiguous bytes.  Words that hold
// more than one run (noise, text: ~1 % of the words) are written by their own lane afterwards.  (A thread-per-4-pixels
// version needed 1.1 instructions per pixel and was bound by them: 0.70 ms for 16 x 4096^2; this one needs ~0.15.)
// The parent entries sit at run-start pixels and this pass overwrites them with root + 1 — still a valid parent pointer for
// any thread that chases through them concurrently.
constexpr int CW_WARPS = 8;  // warps per CTA
constexpr int CW_R = 4;      // rows per warp: their pointer chases are independent and run side by side

// rare path of pass C: this lane's word holds several runs — resolve each and write the lane's own 32 pixels
__device__ __noinline__ int ccl_write_multi(int* L, int* dst, uint32_t word, int base, int x, int W, int first_lab, int vec_ok) {
  int n_roots = 0;
  bool first = true;
  for (uint32_t s = word & ~(word << 1); s; s &= s - 1) {
    const int st = __ffs(s) - 1;
    int l2 = first_lab;
    if (!first) {  // a further run always starts inside the word: it has its own parent entry
      const int root = gfind(L, base + st);
      l2 = root + 1;
      n_roots += (root == base + st);
    }
    first = false;
    const int len = run_len(word, st);
    for (int k = st; k < st + len; k++)
      if (x + k < W) dst[k] = l2;
  }
  // zeros between the runs
  for (uint32_t z = ~word; z; z &= z - 1) {
    const int k = __ffs(z) - 1;
    if (x + k < W) dst[k] = 0;
  }
  (void)vec_ok;
  return n_roots;
}

__global__ void __launch_bounds__(CW_WARPS * 32, 5) k_ccl_write(int* __restrict__ labels, const uint32_t* __restrict__ bits_all,
                                                           const int* __restrict__ first_all, int H, int W, int wpr, int vec_ok,
                                                           int* __restrict__ ncomp, int* __restrict__ partial) {
  const int b = blockIdx.z, y0 = (blockIdx.y * CW_WARPS + (threadIdx.x >> 5)) * CW_R, c = threadIdx.x & 31;
  if (y0 >= H) return;
  int* L = labels + (size_t)b * H * W;
  const uint32_t* bits = bits_all + (size_t)b * H * wpr;
  const int* first = first_all + (size_t)b * H * wpr;
  const int wc = blockIdx.x * 32 + c;
  uint32_t word[CW_R];
  int cur[CW_R], par[CW_R], start[CW_R];
  bool cont[CW_R];
#pragma unroll
  for (int i = 0; i < CW_R; i++) word[i] = (wc < wpr && y0 + i < H) ? bits[(size_t)(y0 + i) * wpr + wc] : 0u;
#pragma unroll
  for (int i = 0; i < CW_R; i++) {
    uint32_t wl = __shfl_up_sync(0xffffffffu, word[i], 1);
    if (c == 0) wl = 0u;
    cont[i] = (word[i] & 1u) && (wl >> 31);  // the first run of my word entered from the left
    start[i] = (y0 + i) * W + wc * 32 + __ffs(word[i]) - 1;  // the run's own element, when it starts in this word (!cont)
    cur[i] = word[i] ? __ldcs(first + (size_t)(y0 + i) * wpr + wc) : -1;  // pass A's label of my first sub-run
  }
#pragma unroll
  for (int i = 0; i < CW_R; i++) par[i] = cur[i] >= 0 ? __ldcg(L + cur[i]) - 1 : -1;
  for (bool moving = true; moving;) {  // all chains advance together
    moving = false;
#pragma unroll
    for (int i = 0; i < CW_R; i++)
      if (par[i] != cur[i]) {
        cur[i] = par[i];
        par[i] = __ldcg(L + cur[i]) - 1;
        moving = true;
      }
  }
  int n_roots = 0;
  const int sub = c >> 3, sh = (c & 7) * 4;
#pragma unroll
  for (int i = 0; i < CW_R; i++) {
    const int y = y0 + i;
    if (y >= H) break;
    const int lab = cur[i] + 1;  // 0 for an empty word
    if (word[i] && !cont[i]) n_roots += (cur[i] == start[i]);
    const uint32_t starts = word[i] & ~(word[i] << 1);
    const bool multi = (starts & (starts - 1)) != 0u;
    const uint32_t mm = __ballot_sync(0xffffffffu, multi);
    int* row = L + (size_t)y * W + blockIdx.x * 1024;
#pragma unroll
    for (int q = 0; q < 8; q++) {
      const int src = 4 * q + sub;
      const uint32_t wsrc = __shfl_sync(0xffffffffu, word[i], src);
      const int lsrc = __shfl_sync(0xffffffffu, lab, src);
      const uint32_t nib = (wsrc >> sh) & 0xFu;
      const int x = blockIdx.x * 1024 + src * 32 + sh;
      if (((mm >> src) & 1u) || x >= W) continue;
      const int4 o = make_int4((nib & 1u) ? lsrc : 0, (nib & 2u) ? lsrc : 0, (nib & 4u) ? lsrc : 0, (nib & 8u) ? lsrc : 0);
      int* dst = row + src * 32 + sh;
      if (vec_ok && x + 4 <= W) {
        *(int4*)dst = o;
      } else {
        const int v[4] = {o.x, o.y, o.z, o.w};
        for (int k = 0; k < 4; k++)
          if (x + k < W) dst[k] = v[k];
      }
    }
    if (multi) n_roots += ccl_write_multi(L, row + c * 32, word[i], y * W + wc * 32, wc * 32, W, lab, vec_ok);
  }
  if (ncomp) {
    for (int o = 16; o; o >>= 1) n_roots += __shfl_xor_sync(0xffffffffu, n_roots, o);
    if (c == 0 && n_roots) {
      if (partial) atomicAdd(partial + ((size_t)b * 32 + ((blockIdx.x + blockIdx.y) & 31)) * 32, n_roots);
      else atomicAdd(ncomp + b, n_roots);
    }
  }
}

__global__ void k_ccl_count_finish(const int* __restrict__ partial, int* __restrict__ ncomp, int B) {
  int b = blockIdx.x, l = threadIdx.x;
  int v = partial[((size_t)b * 32 + l) * 32];
  for (int o = 16; o; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  if (l == 0 && b < B) ncomp[b] = v;
}

}  // namespace cvb

using namespace cvb;

static size_t ccl_partial_bytes(int B) { return (size_t)(B > 0 ? B : 1) * 32 * 32 * sizeof(int); }
static size_t ccl_plane_bytes(int B, int H, int W) { return (((size_t)B * H * ((W + 31) / 32) * 4) + 255) & ~(size_t)255; }

template <int CONN>
static int ccl_run(const uint8_t* masks, int B, int H, int W, int32_t* labels, int32_t* n_components, int* partial,
                   uint32_t* bits, int* first, cudaStream_t st) {
  const int vec_ok = (W % 16 == 0) && (((uintptr_t)masks & 15) == 0) && (((uintptr_t)labels & 15) == 0);
  const int wpr = (W + 31) / 32;
  const double px = (double)B * H * W;
  cvb_next_work(1.0 * px);
  static std::atomic<unsigned long long> carveout_set{0};
  if (cvb_once_per_device(carveout_set))  // 7 CTAs x 32 KB of label slots per SM: ask for the large shared-memory split
    CVB_CHECK(cudaFuncSetAttribute(k_ccl_scan<CONN>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
  // (a specialisation without the bounds tests and the byte-wise load path compiled to 40 registers = 6 CTAs per SM and was no
  //  faster at 128 images and slower at 16, where 1024 CTAs then fill 1.15 waves)
  CVB_LAUNCH((k_ccl_scan<CONN>), dim3((W + 1023) / 1024, (H + CS_TH * CS_WARPS - 1) / (CS_TH * CS_WARPS), B), dim3(CS_WARPS * 32), 0,
             st, masks, labels, bits, first, H, W, wpr, vec_ok);
  const int seams_h = (H - 1) / CS_TH;
  const long long vpx = (long long)((W - 1) / 1024) * H;
  const int row_blocks_x = (wpr + 127) / 128, row_blocks = row_blocks_x * seams_h;
  const long long col_blocks = (vpx + 127) / 128;
  if (row_blocks + col_blocks > 0)
    CVB_LAUNCH((k_ccl_seams<CONN>), dim3((unsigned)(row_blocks + col_blocks), 1, B), dim3(128), 0, st, labels, bits, first, H, W, wpr,
               row_blocks_x, row_blocks);
  if (seams_h > 0)
    CVB_LAUNCH(k_ccl_flatten_rows, dim3((wpr + 127) / 128, seams_h, B), dim3(128), 0, st, labels, bits, first, H, W, wpr);
  cvb_next_work(4.0 * px);
  CVB_LAUNCH(k_ccl_write, dim3((W + 1023) / 1024, (H + CW_WARPS * CW_R - 1) / (CW_WARPS * CW_R), B), dim3(CW_WARPS * 32), 0, st, labels, bits, first, H, W, wpr, vec_ok, n_components,
             partial);
  if (n_components && partial) CVB_LAUNCH(k_ccl_count_finish, dim3(B), dim3(32), 0, st, partial, n_components, B);
  return CV_OK;
}

extern "C" size_t cv_ccl_workspace_bytes(int B, int H, int W) {
  // spread component counters | bit plane (1 bit / pixel) | first[] (one int per 32-pixel word)
  return ccl_partial_bytes(B) + 2 * ccl_plane_bytes(B, H, W);
}

extern "C" int cv_ccl_label(const uint8_t* masks, int B, int H, int W, int connectivity, int32_t* labels,
                            int32_t* n_components, void* workspace, size_t workspace_bytes, void* stream_) {
  cvb_reset_launches();
  if (!masks || !labels || B <= 0 || H <= 0 || W <= 0)
    return cvb_fail(CV_ERR_INVALID, "cv_ccl_label: null pointer or non-positive size");
  if (connectivity != 4 && connectivity != 8) return cvb_fail(CV_ERR_INVALID, "cv_ccl_label: connectivity must be 4 or 8");
  if ((long long)H * W >= (1ll << 31) - 2) return cvb_fail(CV_ERR_INVALID, "cv_ccl_label: image too large for int32 labels");
  if (B > 65535 || H > 65535) return cvb_fail(CV_ERR_INVALID, "cv_ccl_label: batch or height too large");
  if (!workspace || workspace_bytes < cv_ccl_workspace_bytes(B, H, W))
    return cvb_fail(CV_ERR_INVALID, "cv_ccl_label: workspace of cv_ccl_workspace_bytes(B, H, W) bytes required");
  cudaStream_t st = (cudaStream_t)stream_;
  int* partial = (int*)workspace;
  uint32_t* bits = (uint32_t*)((uint8_t*)workspace + ccl_partial_bytes(B));
  int* first = (int*)((uint8_t*)bits + ccl_plane_bytes(B, H, W));
  if (n_components) {
    CVB_CHECK(cudaMemsetAsync(partial, 0, ccl_partial_bytes(B), st));
    CVB_CHECK(cudaMemsetAsync(n_components, 0, (size_t)B * 4, st));
  }
  return connectivity == 8 ? ccl_run<8>(masks, B, H, W, labels, n_components, partial, bits, first, st)
                           : ccl_run<4>(masks, B, H, W, labels, n_components, partial, bits, first, st);
}
