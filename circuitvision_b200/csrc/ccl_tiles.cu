// Native-resolution connected-component labelling (cv_ccl_label, BASELINE cfg 4) — tiled, bit-parallel union-find.
//
// Not a reference code path: the oracle is cv2.connectedComponents up to renaming; labels are
// canonical: labels[p] = 1 + min linear index of p's component, 0 for background.
//
// Traffic plan (the kernel is HBM-bound; algorithmic bytes = 1 B/px mask read + 4 B/px label write):
//   pass A  k_ccl_tile_label : each CTA loads a 256 x 32 tile of the mask as BITS (one 32-pixel word per thread), labels it
//                              entirely in shared memory (elements = maximal horizontal runs; unions between adjacent
//                              rows by bit overlap; atomicMin union-find on a 16 KB parent array, one slot per pixel pair) and writes the label
//                              image once, fully coalesced: 1 + 4 B/px.  It also emits a per-word "dirty" bitmap
//                              (word holds a run of a component that touches the tile border) and counts the
//                              components that stay inside their tile.
//   pass B  k_ccl_seams_h/_v : only the tile seams (3.5 % of the image; horizontal seams one 32-pixel word per thread, one
//                              union per run contact) merge components across tiles with the
//                              global atomicMin union-find on the label image (roots point at roots; the finds halve the
//                              paths they walk, which replaced a separate compression pass).
//   pass C  k_ccl_tile_fixup : one thread per 32-pixel word re-reads the mask bits (1 B/px, no full label read), takes
//                              the tile-local root named at each sub-run's first pixel, looks up its global root and
//                              rewrites only the runs whose component changed (sparse row segments); counts roots.
// Total ≈ 6 B/px + seams instead of the 16+ B/px of a per-pixel init / merge / flatten pipeline.
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdlib.h>

#include "../../include/cv_b200.h"
#include "common.cuh"

namespace cvb {

constexpr int CT_W = 256, CT_H = 32, CT_WORDS = CT_W / 32, CT_THREADS = CT_H * CT_WORDS;  // 256 x 32 tile, 256 threads

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

// ---- shared-memory union-find on sub-run start positions (tile-local pixel index)
// Slot of element x = x >> 1: two horizontally adjacent pixels are never both the first pixel of a (sub-)run, so the
// parent array needs one slot per pixel PAIR (8 KB per tile instead of 16 KB: twice the resident tiles per SM).
__device__ __forceinline__ int sfind(const volatile int* P, int x) {
  int y = P[x >> 1];
  while (y != x) {
    x = y;
    y = P[x >> 1];
  }
  return x;
}
__device__ __forceinline__ void sunion(int* P, int a, int b) {
  bool done;
  do {
    a = sfind(P, a);
    b = sfind(P, b);
    if (a < b) {
      int old = atomicMin(P + (b >> 1), a);
      done = (old == b);
      b = old;
    } else if (b < a) {
      int old = atomicMin(P + (a >> 1), b);
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

// Element of the union-find = a maximal horizontal run of the tile row (it may span several 32-pixel words), identified
// by the tile-local index of its first pixel.  fid[r][c] = element of the run that ENTERS word c of row r from the left
// (valid when the word's bit 0 is set and the previous word's bit 31 is set).
__device__ __forceinline__ bool carries_in(const uint32_t (*bits)[CT_WORDS], int r, int c) {
  return c > 0 && (bits[r][c] & 1u) && (bits[r][c - 1] >> 31);
}
__device__ __forceinline__ int elem_of(const uint32_t (*bits)[CT_WORDS], const int (*fid)[CT_WORDS], int r, int c, int st) {
  return (st == 0 && carries_in(bits, r, c)) ? fid[r][c] : r * CT_W + c * 32 + st;
}

// Thread <-> word mapping: lane = row, warp = word column (r = t % 32, c = t / 32).  A vertical wire then keeps whole
// warps busy instead of one lane in every warp, which matters because the union-find code is divergent.
//
// Builds the tile's union-find in shared memory.  On return (after the trailing __syncthreads) the forest is complete:
// sfind(P, e) gives the ROOT (tile-local pixel index of the component's first pixel in raster order) of run element e.
template <int CONN>
__device__ __forceinline__ uint32_t tile_label(const uint8_t* __restrict__ im, int H, int W, int x0, int y0, bool vec_ok,
                                               uint32_t (*bits)[CT_WORDS], int (*fid)[CT_WORDS], int* P) {
  const int r = threadIdx.x % CT_H, c = threadIdx.x / CT_H;
  const uint32_t w = load_word(im, H, W, y0 + r, x0 + c * 32, vec_ok);
  const int base = r * CT_W + c * 32;
  bits[r][c] = w;
  __syncthreads();
  // the run that enters my word from the left starts in the nearest word to the left that is not completely set
  bool cin = false;
  int first = base;
  if (c > 0 && (w & 1u) && (bits[r][c - 1] >> 31)) {
    cin = true;
    for (int k = c - 1; k >= 0; k--) {
      const uint32_t lw = bits[r][k];
      const int ls = run_start(lw, 31);
      first = r * CT_W + k * 32 + ls;
      if (ls > 0 || k == 0 || !(bits[r][k - 1] >> 31)) break;
    }
  }
  fid[r][c] = first;
  uint32_t starts = w & ~(w << 1);
  if (cin) starts &= ~1u;  // a continued run is not an element of its own
  for (uint32_t s = starts; s; s &= s - 1) {
    int i = __ffs(s) - 1;
    P[(base + i) >> 1] = base + i;
  }
  __syncthreads();
  if (w && r > 0) {
    const uint32_t up = bits[r - 1][c];
    const uint32_t upl = (c > 0) ? bits[r - 1][c - 1] : 0u, upr = (c + 1 < CT_WORDS) ? bits[r - 1][c + 1] : 0u;
    const bool up_cin = (c > 0) && (up & 1u) && (upl >> 31);
    uint32_t cur = w;
    while (cur) {
      const int s = __ffs(cur) - 1;
      const int len = run_len(cur, s);
      const uint32_t rm = run_mask(s, len);
      cur &= ~rm;
      const int me = (s == 0 && cin) ? first : base + s;
      uint32_t aw = rm;
      if (CONN == 8) aw |= (rm << 1) | (rm >> 1);
      uint32_t cand = aw & up;
      while (cand) {
        const int us = __ffs(cand) - 1;
        const int st = run_start(up, us);
        cand &= ~run_mask(st, run_len(up, st));
        // both runs continue from the previous word and already touch there: that thread made the union
        if (s == 0 && cin && st == 0 && up_cin) continue;
        sunion(P, me, elem_of(bits, fid, r - 1, c, st));
      }
      if (CONN == 8) {
        if (s == 0 && !cin && (upl >> 31)) sunion(P, me, elem_of(bits, fid, r - 1, c - 1, run_start(upl, 31)));
        if (s + len == 32 && !(up >> 31) && (upr & 1u)) sunion(P, me, elem_of(bits, fid, r - 1, c + 1, 0));
      }
    }
  }
  __syncthreads();
  return w;
}

// dirty[(b*n_tiles + tile)*4 + warp] bit l: the word owned by thread 32*warp + l holds a sub-run of a component that
// touches the tile border — only such components can be merged by the seam pass, so only those words are revisited by
// pass C.  Components that stay inside their tile are final after pass A and are counted here.
template <int CONN>
__global__ void __launch_bounds__(CT_THREADS, 8) k_ccl_tile_label(const uint8_t* __restrict__ masks, int* __restrict__ labels,
                                                                int H, int W, int vec_ok, uint32_t* __restrict__ dirty,
                                                                int* __restrict__ ncomp, int* __restrict__ partial) {
  __shared__ uint32_t bits[CT_H][CT_WORDS];
  __shared__ int fid[CT_H][CT_WORDS];
  __shared__ int P[CT_H * CT_W / 2];
  __shared__ uint32_t touch[CT_H * CT_W / 32];  // bit per tile-local pixel: root of a border-touching component
  __shared__ int closed_roots;
  const int b = blockIdx.z, x0 = blockIdx.x * CT_W, y0 = blockIdx.y * CT_H;
  const uint8_t* im = masks + (size_t)b * H * W;
  int* L = labels + (size_t)b * H * W;
  if (threadIdx.x < CT_H * CT_W / 32) touch[threadIdx.x] = 0u;
  if (threadIdx.x == 0) closed_roots = 0;
  const uint32_t w = tile_label<CONN>(im, H, W, x0, y0, vec_ok != 0, bits, fid, P);
  int* G = P;  // after the second sync below the slots hold the GLOBAL label of the sub-run starting there
  const int r = threadIdx.x % CT_H, c = threadIdx.x / CT_H;
  const int base = r * CT_W + c * 32;
  const bool cin = carries_in(bits, r, c);
  const uint32_t sub = w & ~(w << 1);  // every sub-run of my word, continued or not
  // root and global label of every sub-run; roots of components that touch the tile border are marked.  The root is
  // parked in the slot of base + i: for a run element that is path compression, for a continued sub-run the slot is unused.
  for (uint32_t s = sub; s; s &= s - 1) {
    const int i = __ffs(s) - 1;
    const int root = sfind(P, (i == 0 && cin) ? fid[r][c] : base + i);
    const bool edge = r == 0 || r == CT_H - 1 || (c == 0 && i == 0) || (c == CT_WORDS - 1 && i + run_len(w, i) == 32);
    if (edge) atomicOr(&touch[root >> 5], 1u << (root & 31));
    if (root != base + i) P[(base + i) >> 1] = root;
  }
  __syncthreads();
  bool my_dirty = false;
  int n_closed = 0;
  for (uint32_t s = sub; s; s &= s - 1) {
    const int i = __ffs(s) - 1;
    const int root = P[(base + i) >> 1];
    const bool t = (touch[root >> 5] >> (root & 31)) & 1u;
    my_dirty |= t;
    n_closed += (!t && root == base + i);  // first pixel of a component that cannot change any more
    G[(base + i) >> 1] = (y0 + root / CT_W) * W + x0 + (root % CT_W) + 1;  // nobody else reads my slots any more
  }
  if (my_dirty) n_closed = 0;  // pass C revisits this word and counts every root in it
  const uint32_t dmask = __ballot_sync(0xffffffffu, my_dirty);
  if (dirty && (threadIdx.x & 31) == 0)
    dirty[(((size_t)b * gridDim.y + blockIdx.y) * gridDim.x + blockIdx.x) * (CT_THREADS / 32) + (threadIdx.x >> 5)] = dmask;
  if (ncomp) {
    for (int o = 16; o; o >>= 1) n_closed += __shfl_xor_sync(0xffffffffu, n_closed, o);
    if ((threadIdx.x & 31) == 0 && n_closed) atomicAdd(&closed_roots, n_closed);
  }
  __syncthreads();
  if (ncomp && dirty && threadIdx.x == 0 && closed_roots) {
    if (partial) atomicAdd(partial + ((size_t)b * 32 + ((blockIdx.x + blockIdx.y) & 31)) * 32, closed_roots);
    else atomicAdd(ncomp + b, closed_roots);
  }
  // coalesced label write: each warp takes rows warp, warp + n_warps, ...; per 128-pixel segment of the row lane l owns
  // pixels 4l..4l+3
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int sh = (lane & 7) * 4;
  for (int rr = warp; rr < CT_H; rr += CT_THREADS / 32) {
    const int y = y0 + rr;
    if (y >= H) break;
#pragma unroll
    for (int seg = 0; seg < CT_W / 128; seg++) {
      const int wc = seg * 4 + (lane >> 3);
      const uint32_t word = bits[rr][wc];
      const uint32_t nib = (word >> sh) & 0xFu;
      int out[4] = {0, 0, 0, 0};
      if (nib) {
        const int gb = rr * CT_W + wc * 32;
        if (nib == 0xFu) {
          out[0] = out[1] = out[2] = out[3] = G[(gb + run_start(word, sh)) >> 1];
        } else {
#pragma unroll
          for (int k = 0; k < 4; k++)
            if ((nib >> k) & 1u) out[k] = G[(gb + run_start(word, sh + k)) >> 1];
        }
      }
      const int x = x0 + seg * 128 + lane * 4;
      int* dst = L + (size_t)y * W + x;
      if (vec_ok && x + 4 <= W) {
        *(int4*)dst = make_int4(out[0], out[1], out[2], out[3]);
      } else {
        for (int k = 0; k < 4; k++)
          if (x + k < W) dst[k] = out[k];
      }
    }
  }
}

// Seams.  Horizontal seams (rows y = k*CT_H, 3.1 % of the image) are handled one 32-pixel WORD per thread: the seam row
// and the row above it are read as bits, and every (run below, run above) contact is one union between run
// representatives — all pixels of a run carry the same tile-local root after pass A, so any pixel of it starts the
// same find.  Vertical seams (columns x = k*CT_W, 0.8 %) stay one pixel per thread.
template <int CONN>
__global__ void __launch_bounds__(128) k_ccl_seams_h(const uint8_t* __restrict__ masks, int* __restrict__ labels, int H, int W,
                                                      int vec_ok) {
  const int b = blockIdx.z;
  const uint8_t* im = masks + (size_t)b * H * W;
  int* L = labels + (size_t)b * H * W;
  const int y = (blockIdx.y + 1) * CT_H;
  const int x0 = (blockIdx.x * 128 + threadIdx.x) * 32;
  if (y >= H || x0 >= W) return;
  const uint32_t cur = load_word(im, H, W, y, x0, vec_ok != 0);
  if (!cur) return;
  const uint32_t up = load_word(im, H, W, y - 1, x0, vec_ok != 0);
  const int rowc = y * W + x0, rowu = (y - 1) * W + x0;
  bool ul = false, ur = false;
  if (CONN == 8) {
    ul = x0 > 0 && im[rowu - 1] != 0;
    ur = x0 + 32 < W && im[rowu + 32] != 0;
  }
  if (!up && !ul && !ur) return;
  uint32_t rest = cur;
  while (rest) {
    const int st = __ffs(rest) - 1;
    const int len = run_len(rest, st);
    const uint32_t rm = run_mask(st, len);
    rest &= ~rm;
    uint32_t aw = rm;
    if (CONN == 8) aw |= (rm << 1) | (rm >> 1);
    uint32_t cand = aw & up;
    while (cand) {
      const int us = __ffs(cand) - 1;
      const int ust = run_start(up, us);
      cand &= ~run_mask(ust, run_len(up, ust));
      gunion(L, rowc + st, rowu + ust);
    }
    if (CONN == 8) {
      if (st == 0 && ul) gunion(L, rowc, rowu - 1);
      if (st + len == 32 && ur) gunion(L, rowc + 31, rowu + 32);
    }
  }
}

template <int CONN>
__global__ void __launch_bounds__(256) k_ccl_seams_v(const uint8_t* __restrict__ masks, int* __restrict__ labels, int H, int W) {
  const int b = blockIdx.z;
  const uint8_t* im = masks + (size_t)b * H * W;
  int* L = labels + (size_t)b * H * W;
  const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const int n_seams = (W - 1) / CT_W;
  if (t >= (long long)n_seams * H) return;
  const int x = (int)(t / H + 1) * CT_W, y = (int)(t % H);
  const int p = y * W + x;
  if (!im[p]) return;
  if (im[p - 1]) {
    gunion(L, p, p - 1);
  } else if (CONN == 8) {
    if (y > 0 && im[p - W - 1]) gunion(L, p, p - W - 1);
    if (y + 1 < H && im[p + W - 1]) gunion(L, p, p + W - 1);
  }
}

// pass C: same thread <-> word mapping as pass A.  Only words flagged dirty are revisited.  The label pass A wrote at
// the LAST pixel of a sub-run names the run's tile-local root (only root pixels — always the first pixel of a run — are
// modified by the seam unions); if that root was merged into another component, the whole run is rewritten.
// With dirty == nullptr every word is visited (and every root is counted here).
__global__ void __launch_bounds__(CT_THREADS) k_ccl_tile_fixup(const uint8_t* __restrict__ masks, int* __restrict__ labels, int H,
                                                                int W, int vec_ok, const uint32_t* __restrict__ dirty,
                                                                int* __restrict__ ncomp, int* __restrict__ partial) {
  const int b = blockIdx.z, x0 = blockIdx.x * CT_W, y0 = blockIdx.y * CT_H;
  const int lane = threadIdx.x & 31;
  uint32_t dmask = 0xffffffffu;
  if (dirty)
    dmask = __ldg(dirty + (((size_t)b * gridDim.y + blockIdx.y) * gridDim.x + blockIdx.x) * (CT_THREADS / 32) + (threadIdx.x >> 5));
  if (dmask == 0u) return;  // whole warp: nothing in its 8 rows can have changed
  const uint8_t* im = masks + (size_t)b * H * W;
  int* L = labels + (size_t)b * H * W;
  const int r = threadIdx.x % CT_H, c = threadIdx.x / CT_H;
  const int y = y0 + r, x = x0 + c * 32;
  int n_roots = 0;
  if ((dmask >> lane) & 1u) {
    const uint32_t w = load_word(im, H, W, y, x, vec_ok != 0);
    for (uint32_t s = w & ~(w << 1); s; s &= s - 1) {
      const int i = __ffs(s) - 1;
      const int len = run_len(w, i);
      const int px = y * W + x + i;
      const int ref = __ldcg(L + px + len - 1);  // tile-local root + 1 (a run's last pixel is never a root unless len == 1)
      const int f = gfind(L, ref - 1);
      if (ref != f + 1) {
        // rewrite the run: 16-byte stores over its aligned middle part (a 16-pixel wire crossing is 4 stores, not 16)
        const int v = f + 1;
        int k = 0;
        if (vec_ok) {
          for (; k < len && ((px + k) & 3); k++) L[px + k] = v;
          for (; k + 4 <= len; k += 4) *(int4*)(L + px + k) = make_int4(v, v, v, v);
        }
        for (; k < len; k++) L[px + k] = v;
      }
      // this run starts at the first pixel of its component (pass A counted the roots of the words it left clean)
      n_roots += (f == px);
    }
  }
  if (ncomp) {
    for (int o = 16; o; o >>= 1) n_roots += __shfl_xor_sync(0xffffffffu, n_roots, o);
    if (lane == 0 && n_roots) {
      if (partial) atomicAdd(partial + ((size_t)b * 32 + ((blockIdx.x + blockIdx.y + 7) & 31)) * 32, n_roots);
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

template <int CONN>
static int ccl_run(const uint8_t* masks, int B, int H, int W, int32_t* labels, int32_t* n_components, int* partial,
                   uint32_t* dirty, cudaStream_t st) {
  const int vec_ok = (W % 16 == 0) && (((uintptr_t)masks & 15) == 0) && (((uintptr_t)labels & 15) == 0);
  dim3 tg((W + CT_W - 1) / CT_W, (H + CT_H - 1) / CT_H, B);
  const double px = (double)B * H * W;
  cvb_next_work(5.0 * px);
  static std::atomic<unsigned long long> carveout_set{0};
  if (cvb_once_per_device(carveout_set)) {
    // 8 resident tiles x ~19 KB: ask for the large shared-memory split (the default heuristic leaves room for fewer)
    CVB_CHECK(cudaFuncSetAttribute(k_ccl_tile_label<CONN>, cudaFuncAttributePreferredSharedMemoryCarveout,
                                   cudaSharedmemCarveoutMaxShared));
  }
  CVB_LAUNCH((k_ccl_tile_label<CONN>), tg, dim3(CT_THREADS), 0, st, masks, labels, H, W, vec_ok, dirty, n_components, partial);
  const int seams_h = (H - 1) / CT_H, seams_v = (W - 1) / CT_W;
  const int words = (W + 31) / 32;
  const long long vpx = (long long)seams_v * H;
  if (seams_h > 0)
    CVB_LAUNCH((k_ccl_seams_h<CONN>), dim3((words + 127) / 128, seams_h, B), dim3(128), 0, st, masks, labels, H, W, vec_ok);
  if (vpx > 0)
    CVB_LAUNCH((k_ccl_seams_v<CONN>), dim3((unsigned)((vpx + 255) / 256), 1, B), dim3(256), 0, st, masks, labels, H, W);
  CVB_LAUNCH(k_ccl_tile_fixup, tg, dim3(CT_THREADS), 0, st, masks, labels, H, W, vec_ok, dirty, n_components, partial);
  if (n_components && partial) CVB_LAUNCH(k_ccl_count_finish, dim3(B), dim3(32), 0, st, partial, n_components, B);
  return CV_OK;
}

static size_t ccl_dirty_bytes(int B, int H, int W) {
  return (size_t)B * ((W + CT_W - 1) / CT_W) * ((H + CT_H - 1) / CT_H) * (CT_THREADS / 32) * sizeof(uint32_t);
}
static size_t ccl_partial_bytes(int B) { return (size_t)(B > 0 ? B : 1) * 32 * 32 * sizeof(int); }

extern "C" size_t cv_ccl_workspace_bytes(int B, int H, int W) {
  // labels are resolved in place in the caller's label image; the workspace holds the spread component counters and
  // the per-tile dirty bitmap
  return ccl_partial_bytes(B) + ccl_dirty_bytes(B, H, W);
}

extern "C" int cv_ccl_label(const uint8_t* masks, int B, int H, int W, int connectivity, int32_t* labels,
                            int32_t* n_components, void* workspace, size_t workspace_bytes, void* stream_) {
  cvb_reset_launches();
  if (!masks || !labels || B <= 0 || H <= 0 || W <= 0)
    return cvb_fail(CV_ERR_INVALID, "cv_ccl_label: null pointer or non-positive size");
  if (connectivity != 4 && connectivity != 8) return cvb_fail(CV_ERR_INVALID, "cv_ccl_label: connectivity must be 4 or 8");
  if ((long long)H * W >= (1ll << 31) - 2) return cvb_fail(CV_ERR_INVALID, "cv_ccl_label: image too large for int32 labels");
  if (B > 65535) return cvb_fail(CV_ERR_INVALID, "cv_ccl_label: batch too large");
  cudaStream_t st = (cudaStream_t)stream_;
  int* partial = nullptr;
  uint32_t* dirty = nullptr;
  if (workspace && workspace_bytes >= cv_ccl_workspace_bytes(B, H, W)) {
    partial = (int*)workspace;
    dirty = (uint32_t*)((uint8_t*)workspace + ccl_partial_bytes(B));
    if (n_components) CVB_CHECK(cudaMemsetAsync(partial, 0, ccl_partial_bytes(B), st));
  }
  if (n_components) CVB_CHECK(cudaMemsetAsync(n_components, 0, (size_t)B * 4, st));
  return connectivity == 8 ? ccl_run<8>(masks, B, H, W, labels, n_components, partial, dirty, st)
                           : ccl_run<4>(masks, B, H, W, labels, n_components, partial, dirty, st);
}
