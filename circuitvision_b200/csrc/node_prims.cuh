// Per-pixel / per-component arithmetic of the node-analysis path, written once as
// __host__ __device__ inline functions so the CUDA kernels (node_kernels.cu) and the CPU
// unit-test harness (tests/host_harness.cpp, g++) execute the same lines.  This is product
// code, not the oracle: the oracle is cv2 itself (oracle/node_oracle.py).
//
// Behavioural spec (all of it is OpenCV 4.13 behaviour the reference relies on):
//   resize  : cv2.resize(INTER_LINEAR, u8)  <- /root/reference/src/circuit_analyzer.py:806
//   blur    : cv2.GaussianBlur((5,5),1) u8  <- circuit_analyzer.py:304
//   morph   : dilate/erode 3x3 x2           <- circuit_analyzer.py:308,311
//   contour : findContours(EXTERNAL,SIMPLE) <- circuit_analyzer.py:404 ; contourArea :410 ;
//             boundingRect :412 ; moments :1481
#pragma once
#include <stdint.h>
#include <math.h>

#if defined(__CUDACC__)
#define CV_HD __host__ __device__ __forceinline__
#else
#define CV_HD inline
#endif

namespace cvb {

// ---------------------------------------------------------------- resize
struct ResizeTap {
  int i0, i1;  // clamped source indices
  int c0, c1;  // 11-bit fixed-point weights (sum 2048)
};

CV_HD int cv_round_half_even(float v) {
#if defined(__CUDA_ARCH__)
  return __float2int_rn(v);
#else
  return (int)lrintf(v);  // default FE_TONEAREST == round-half-even
#endif
}

// Coefficients of destination index d when resizing an axis of n_src samples to n_dst samples.
// x axis clamps the fraction at both ends, y axis only clamps the row indices (OpenCV quirk).
CV_HD ResizeTap resize_tap(int d, int n_dst, int n_src, bool is_x) {
  double inv_scale = (double)n_dst / (double)n_src;
  double scale = 1.0 / inv_scale;
  float f = (float)((d + 0.5) * scale - 0.5);
  int s = (int)floorf(f);
  f -= (float)s;
  if (is_x) {
    if (s < 0) { s = 0; f = 0.f; }
    if (s >= n_src - 1) { s = n_src - 1; f = 0.f; }
  }
  ResizeTap t;
  t.c0 = cv_round_half_even((1.f - f) * 2048.f);
  t.c1 = cv_round_half_even(f * 2048.f);
  int a = s, b = s + 1;
  t.i0 = a < 0 ? 0 : (a > n_src - 1 ? n_src - 1 : a);
  t.i1 = b < 0 ? 0 : (b > n_src - 1 ? n_src - 1 : b);
  return t;
}

CV_HD int resize_hpass(int p0, int p1, const ResizeTap& tx) { return p0 * tx.c0 + p1 * tx.c1; }

CV_HD uint8_t resize_vpass(int r0, int r1, const ResizeTap& ty) {
  int v = (((ty.c0 * (r0 >> 4)) >> 16) + ((ty.c1 * (r1 >> 4)) >> 16) + 2) >> 2;
  return (uint8_t)v;  // weights sum to 2048 => always within 0..255
}

// ---------------------------------------------------------------- blur / morphology
CV_HD int reflect101(int i, int n) {
  if (n == 1) return 0;
  while (i < 0 || i >= n) {
    if (i < 0) i = -i;
    else i = 2 * n - 2 - i;
  }
  return i;
}

// 8.8 fixed-point Gaussian taps for ksize 5, sigma 1
#define CVB_G0 14
#define CVB_G1 62
#define CVB_G2 104

CV_HD uint32_t gauss5_h(int a, int b, int c, int d, int e) {
  return (uint32_t)(CVB_G0 * a + CVB_G1 * b + CVB_G2 * c + CVB_G1 * d + CVB_G0 * e);  // <= 65280
}
CV_HD uint8_t gauss5_v(uint32_t a, uint32_t b, uint32_t c, uint32_t d, uint32_t e) {
  uint32_t v = CVB_G0 * a + CVB_G1 * b + CVB_G2 * c + CVB_G1 * d + CVB_G0 * e;
  return (uint8_t)((v + 32768u) >> 16);
}

// ---------------------------------------------------------------- contour following
// 8-direction chain codes, OpenCV order: 0=E 1=NE 2=N 3=NW 4=W 5=SW 6=S 7=SE (y grows downwards)
CV_HD int code_dx(int s) { return (s == 0 || s == 1 || s == 7) ? 1 : ((s >= 3 && s <= 5) ? -1 : 0); }
CV_HD int code_dy(int s) { return (s >= 1 && s <= 3) ? -1 : ((s >= 5 && s <= 7) ? 1 : 0); }

struct ContourStats {
  int32_t nverts;
  int32_t xmin, ymin, xmax, ymax;
  long long a00;  // sum(x_{i-1}*y_i - x_i*y_{i-1})       (2 * signed area)
  long long a01;  // sum(dxy * (y_{i-1} + y_i))           (6 * m01 before sign)
  long long a10;  // sum(dxy * (x_{i-1} + x_i))
};

// Follows the outer border of the 8-connected component whose raster-first pixel is (x0,y0),
// reproducing OpenCV's border follower + CHAIN_APPROX_SIMPLE vertex selection.  `img` is any
// image where foreground = non-zero; pixels outside [0,w)x[0,h) read as background.
// If `out` is non-null the vertices are written as (x,y) int32 pairs (capacity `cap` vertices).
// Returns stats; nverts counts every vertex even beyond cap.
// `fg(x, y)` is the foreground predicate (false outside the image); the byte-image wrapper follows below.
template <typename Fg>
CV_HD ContourStats trace_outer_fg(Fg fg, int x0, int y0, int32_t* out, int cap) {
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
    if (x < st.xmin) st.xmin = x;
    if (x > st.xmax) st.xmax = x;
    if (y < st.ymin) st.ymin = y;
    if (y > st.ymax) st.ymax = y;
    st.nverts++;
  };
  // clockwise search for the first neighbour, starting just after West
  int s = 4;
  const int s_stop = 4;
  int x1 = 0, y1 = 0;
  bool found = false;
  do {
    s = (s - 1) & 7;
    x1 = x0 + code_dx(s);
    y1 = y0 + code_dy(s);
    if (fg(x1, y1)) { found = true; break; }
  } while (s != s_stop);
  if (!found) {
    emit(x0, y0);  // isolated pixel
  } else {
    int x3 = x0, y3 = y0;  // current pixel
    int prev_s = s ^ 4;
    for (;;) {
      // counter-clockwise search for the next border pixel, starting after the back-pointer
      int x4 = 0, y4 = 0;
      int k = s;
      for (;;) {
        k++;
        x4 = x3 + code_dx(k & 7);
        y4 = y3 + code_dy(k & 7);
        if (fg(x4, y4)) break;
        if (k >= s + 8) break;  // cannot happen for a component with >= 2 pixels
      }
      s = k & 7;
      if (s != prev_s) emit(x3, y3);
      prev_s = s;
      bool done = (x4 == x0 && y4 == y0 && x3 == x1 && y3 == y1);
      if (done) break;
      x3 = x4; y3 = y4;
      s = (s + 4) & 7;
    }
  }
  // close the polygon (term i = 0 uses the last vertex as predecessor)
  if (st.nverts > 0) {
    long long dxy = prev_x * first_y - first_x * prev_y;
    st.a00 += dxy;
    st.a01 += dxy * (prev_y + first_y);
    st.a10 += dxy * (prev_x + first_x);
  }
  return st;
}

template <typename Pix>
CV_HD ContourStats trace_outer_simple(const Pix* img, int w, int h, int x0, int y0, int32_t* out, int cap) {
  return trace_outer_fg(
      [=](int x, int y) -> bool { return x >= 0 && y >= 0 && x < w && y < h && img[(size_t)y * w + x] != 0; }, x0, y0, out,
      cap);
}

// cv2.contourArea(c) / (h*w) > thr  evaluated exactly as the reference does in Python doubles
CV_HD bool area_passes(long long a00, int h, int w, double thr) {
  double area = fabs((double)a00 * 0.5);
  double normalizer = (double)((long long)h * (long long)w);
  return area / normalizer > thr;
}

// int(M['m01'] / M['m00']) of cv2.moments(contour); returns false when m00 == 0
CV_HD bool centroid_y(long long a00, long long a01, int* cy) {
  if (!(fabs((double)a00) > 1.1920928955078125e-07)) return false;  // FLT_EPSILON gate in contourMoments
  double db1_2 = a00 > 0 ? 0.5 : -0.5;
  double db1_6 = a00 > 0 ? 0.16666666666666666666666666666667 : -0.16666666666666666666666666666667;
  double m00 = (double)a00 * db1_2;
  double m01 = (double)a01 * db1_6;
  if (m00 == 0) return false;
  *cy = (int)(m01 / m00);
  return true;
}

// reference is_point_near_bbox (circuit_analyzer.py:811-846): inside (inclusive) OR within t of any
// of the four infinite edge lines.
// ---------------------------------------------------------------- segment_circuit (circuit_analyzer.py:313-319)
// cv2.cvtColor(COLOR_RGB2GRAY) on u8: 15-bit fixed point, first channel weighted 0.299 (OpenCV 4.13; exhaustively
// checked over all 2^24 triples in tests/test_node_prims_cpu.py).
CV_HD int gray_rgb2gray(int c0, int c1, int c2) { return (c0 * 9798 + c1 * 19235 + c2 * 3735 + 16384) >> 15; }
// The reference feeds RGB2GRAY a BGR copy of the RGB page (:2231 + :316): in terms of the RGB input the red channel
// gets the blue coefficient and vice versa.
CV_HD int gray_of_rgb_page(int r, int g, int b) { return gray_rgb2gray(b, g, r); }
// cv2.adaptiveThreshold(ADAPTIVE_THRESH_MEAN_C, THRESH_BINARY_INV, 31, 21): mean = boxFilter 31x31 (BORDER_REPLICATE)
// rounded to nearest (no ties: 961 is odd), foreground where src - mean <= -21.
CV_HD int box_mean_31(int window_sum) { return (2 * window_sum + 961) / 1922; }
CV_HD uint8_t adaptive_inv_31_21(int src, int window_sum) { return (src - box_mean_31(window_sum) <= -21) ? 255 : 0; }

CV_HD bool point_near_box(int px, int py, int xmin, int ymin, int xmax, int ymax, int t) {
  if (xmin <= px && px <= xmax && ymin <= py && py <= ymax) return true;
  int dl = px - xmin; dl = dl < 0 ? -dl : dl;
  int dr = px - xmax; dr = dr < 0 ? -dr : dr;
  int dt = py - ymin; dt = dt < 0 ? -dt : dt;
  int db = py - ymax; db = db < 0 ? -db : db;
  return dl <= t || dr <= t || dt <= t || db <= t;
}

}  // namespace cvb
