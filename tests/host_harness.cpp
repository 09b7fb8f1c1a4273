// CPU unit-test harness for the __host__ __device__ arithmetic in
// circuitvision_b200/csrc/node_prims.cuh.  Compiled with g++ by tests/conftest.py
// (no CUDA needed) and compared against cv2 in tests/test_node_prims_cpu.py.
// It exists so the fixed-point / contour-following logic the CUDA kernels execute can be
// verified in the GPU-less build container; it is never loaded by the product.
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <vector>
#include <algorithm>
#include "../circuitvision_b200/csrc/node_prims.cuh"

using namespace cvb;

extern "C" {

void hh_resize(const uint8_t* src, int H, int W, uint8_t* dst, int h, int w) {
  std::vector<int> row0(w), row1(w);
  for (int dy = 0; dy < h; dy++) {
    ResizeTap ty = resize_tap(dy, h, H, false);
    for (int dx = 0; dx < w; dx++) {
      ResizeTap tx = resize_tap(dx, w, W, true);
      int r0 = resize_hpass(src[(size_t)ty.i0 * W + tx.i0], src[(size_t)ty.i0 * W + tx.i1], tx);
      int r1 = resize_hpass(src[(size_t)ty.i1 * W + tx.i0], src[(size_t)ty.i1 * W + tx.i1], tx);
      dst[(size_t)dy * w + dx] = resize_vpass(r0, r1, ty);
    }
  }
}

void hh_enhance(const uint8_t* src, int h, int w, uint8_t* dst) {
  std::vector<uint32_t> hp((size_t)h * w);
  std::vector<uint8_t> bl((size_t)h * w), di((size_t)h * w);
  for (int y = 0; y < h; y++)
    for (int x = 0; x < w; x++) {
      const uint8_t* r = src + (size_t)y * w;
      hp[(size_t)y * w + x] = gauss5_h(r[reflect101(x - 2, w)], r[reflect101(x - 1, w)], r[x],
                                       r[reflect101(x + 1, w)], r[reflect101(x + 2, w)]);
    }
  for (int y = 0; y < h; y++)
    for (int x = 0; x < w; x++)
      bl[(size_t)y * w + x] = gauss5_v(hp[(size_t)reflect101(y - 2, h) * w + x], hp[(size_t)reflect101(y - 1, h) * w + x],
                                       hp[(size_t)y * w + x], hp[(size_t)reflect101(y + 1, h) * w + x],
                                       hp[(size_t)reflect101(y + 2, h) * w + x]);
  for (int y = 0; y < h; y++)
    for (int x = 0; x < w; x++) {
      int m = 0;
      for (int dy = -2; dy <= 2; dy++)
        for (int dx = -2; dx <= 2; dx++) {
          int yy = y + dy, xx = x + dx;
          if (yy < 0 || xx < 0 || yy >= h || xx >= w) continue;
          m = std::max(m, (int)bl[(size_t)yy * w + xx]);
        }
      di[(size_t)y * w + x] = (uint8_t)m;
    }
  for (int y = 0; y < h; y++)
    for (int x = 0; x < w; x++) {
      int m = 255;
      for (int dy = -2; dy <= 2; dy++)
        for (int dx = -2; dx <= 2; dx++) {
          int yy = y + dy, xx = x + dx;
          if (yy < 0 || xx < 0 || yy >= h || xx >= w) continue;
          m = std::min(m, (int)di[(size_t)yy * w + xx]);
        }
      dst[(size_t)y * w + x] = (uint8_t)m;
    }
}

// External contours in cv2 order.  pts: (x,y) pairs, offsets[n+1], stats per contour:
// [a00, a01, xmin, ymin, xmax, ymax].  Returns number of contours (or -1 on overflow).
int hh_external_contours(const uint8_t* img, int h, int w, int32_t* pts, int pts_cap, int32_t* offsets,
                         long long* stats, int max_contours) {
  // foreground 8-conn labels, background 4-conn labels with frame flag (simple BFS flood fill)
  std::vector<int> lab((size_t)h * w, -1);
  std::vector<char> bg_frame;  // per bg label
  std::vector<int> stack;
  int nlab = 0;
  std::vector<int> fg_first;  // raster-first pixel per fg label (label discovered in raster order)
  std::vector<char> is_fg_label;
  for (int p = 0; p < h * w; p++) {
    if (lab[p] >= 0) continue;
    bool f = img[p] != 0;
    int id = nlab++;
    is_fg_label.push_back(f);
    fg_first.push_back(p);
    bg_frame.push_back(0);
    stack.clear();
    stack.push_back(p);
    lab[p] = id;
    while (!stack.empty()) {
      int q = stack.back();
      stack.pop_back();
      int qx = q % w, qy = q / w;
      if (!f && (qx == 0 || qy == 0 || qx == w - 1 || qy == h - 1)) bg_frame[id] = 1;
      for (int dy = -1; dy <= 1; dy++)
        for (int dx = -1; dx <= 1; dx++) {
          if (!dx && !dy) continue;
          if (!f && dx && dy) continue;  // background is 4-connected
          int xx = qx + dx, yy = qy + dy;
          if (xx < 0 || yy < 0 || xx >= w || yy >= h) continue;
          int r = yy * w + xx;
          if (lab[r] >= 0 || ((img[r] != 0) != f)) continue;
          lab[r] = id;
          stack.push_back(r);
        }
    }
  }
  int n = 0, used = 0;
  offsets[0] = 0;
  for (int id = nlab - 1; id >= 0; id--) {  // descending raster index of the first pixel
    if (!is_fg_label[id]) continue;
    int p = fg_first[id];
    int x = p % w, y = p / w;
    bool external = (x == 0) || bg_frame[lab[p - 1]];
    if (!external) continue;
    if (n >= max_contours) return -1;
    ContourStats st = trace_outer_simple<uint8_t>(img, w, h, x, y, pts + 2 * used, pts_cap - used);
    if (used + st.nverts > pts_cap) return -1;
    used += st.nverts;
    offsets[n + 1] = used;
    stats[6 * n + 0] = st.a00;
    stats[6 * n + 1] = st.a01;
    stats[6 * n + 2] = st.xmin;
    stats[6 * n + 3] = st.ymin;
    stats[6 * n + 4] = st.xmax;
    stats[6 * n + 5] = st.ymax;
    n++;
  }
  return n;
}

int hh_area_passes(long long a00, int h, int w, double thr) { return area_passes(a00, h, w, thr) ? 1 : 0; }

int hh_centroid_y(long long a00, long long a01, int* cy) { return centroid_y(a00, a01, cy) ? 1 : 0; }

int hh_point_near_box(int px, int py, int xmin, int ymin, int xmax, int ymax, int t) {
  return point_near_box(px, py, xmin, ymin, xmax, ymax, t) ? 1 : 0;
}

// segment_circuit on the RGB page: grey (with the reference's RGB/BGR swap) + adaptive threshold 31 / 21
void hh_segment_circuit(const uint8_t* rgb, int H, int W, uint8_t* out) {
  std::vector<int> g((size_t)H * W);
  for (size_t i = 0; i < (size_t)H * W; i++) g[i] = gray_of_rgb_page(rgb[3 * i], rgb[3 * i + 1], rgb[3 * i + 2]);
  auto cl = [](int v, int n) { return v < 0 ? 0 : (v >= n ? n - 1 : v); };
  for (int y = 0; y < H; y++)
    for (int x = 0; x < W; x++) {
      int s = 0;
      for (int dy = -15; dy <= 15; dy++)
        for (int dx = -15; dx <= 15; dx++) s += g[(size_t)cl(y + dy, H) * W + cl(x + dx, W)];
      out[(size_t)y * W + x] = adaptive_inv_31_21(g[(size_t)y * W + x], s);
    }
}

int hh_gray(int c0, int c1, int c2) { return gray_rgb2gray(c0, c1, c2); }

}  // extern "C"
