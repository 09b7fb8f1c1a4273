"""Host execution of the __host__ __device__ arithmetic the CUDA node kernels run
(circuitvision_b200/csrc/node_prims.cuh) against cv2 — resize, blur+morphology, external contours."""
import ctypes

import cv2
import numpy as np

from circuitvision_b200 import synth

u8p = ctypes.POINTER(ctypes.c_uint8)
i32p = ctypes.POINTER(ctypes.c_int32)
i64p = ctypes.POINTER(ctypes.c_longlong)


def _contours(L, img):
    h, w = img.shape
    cap = 4 * h * w
    pts = np.empty((cap, 2), np.int32)
    offs = np.empty(h * w + 2, np.int32)
    stats = np.empty((h * w + 1, 6), np.int64)
    n = L.hh_external_contours(img.ctypes.data_as(u8p), h, w, pts.ctypes.data_as(i32p), cap,
                               offs.ctypes.data_as(i32p), stats.ctypes.data_as(i64p), h * w + 1)
    assert n >= 0
    return [pts[offs[i]:offs[i + 1]].copy() for i in range(n)], stats[:n].copy()


def test_resize_bit_exact(host_harness):
    rng = np.random.default_rng(0)
    for (H, W) in [(1024, 1024), (493, 712), (720, 1280), (1200, 1200), (300, 400), (601, 777), (37, 53), (2048, 2048)]:
        src = rng.integers(0, 256, (H, W), dtype=np.uint8)
        nw = int(600 * (W / H))
        dst = np.empty((600, nw), np.uint8)
        host_harness.hh_resize(src.ctypes.data_as(u8p), H, W, dst.ctypes.data_as(u8p), 600, nw)
        assert np.array_equal(dst, cv2.resize(src, (nw, 600))), (H, W)


def test_enhance_bit_exact(host_harness):
    rng = np.random.default_rng(1)
    k = np.ones((3, 3), np.uint8)
    for (h, w) in [(600, 600), (600, 866), (37, 53), (5, 7)]:
        src = rng.integers(0, 256, (h, w), dtype=np.uint8)
        dst = np.empty_like(src)
        host_harness.hh_enhance(src.ctypes.data_as(u8p), h, w, dst.ctypes.data_as(u8p))
        ref = cv2.erode(cv2.dilate(cv2.GaussianBlur(src, (5, 5), 1), k, iterations=2), k, iterations=2)
        assert np.array_equal(dst, ref), (h, w)


def _check(L, img):
    img = np.ascontiguousarray(img)
    ref, _ = cv2.findContours(img.copy(), cv2.RETR_EXTERNAL, cv2.CHAIN_APPROX_SIMPLE)
    got, stats = _contours(L, img)
    assert len(ref) == len(got)
    for r, g, s in zip(ref, got, stats):
        assert np.array_equal(r.reshape(-1, 2), g)
        assert abs(int(s[0])) * 0.5 == cv2.contourArea(r)
        x, y, ww, hh = cv2.boundingRect(r)
        assert (x, y, x + ww - 1, y + hh - 1) == tuple(int(v) for v in s[2:6])
        M = cv2.moments(r)
        cy = ctypes.c_int(0)
        f = L.hh_centroid_y(ctypes.c_longlong(int(s[0])), ctypes.c_longlong(int(s[1])), ctypes.byref(cy))
        if M["m00"] != 0:
            assert f and cy.value == int(M["m01"] / M["m00"])
        else:
            assert not f


def test_external_contours_match_cv2(host_harness):
    rng = np.random.default_rng(2)
    for i in range(40):
        h = int(rng.integers(1, 70))
        w = int(rng.integers(1, 70))
        p = float(rng.uniform(0.1, 0.9))
        _check(host_harness, ((rng.random((h, w)) < p) * 255).astype(np.uint8))
    for i in range(8):
        _check(host_harness, synth.random_blob_mask(i, 160, 220, p=0.5, smooth=2))
    a = np.zeros((30, 30), np.uint8)
    a[5:25, 5:25] = 255
    a[8:22, 8:22] = 0
    a[12:18, 12:18] = 255
    a[14:16, 14:16] = 0
    _check(host_harness, a)  # nested rings: only the outer ring is external
    _check(host_harness, np.zeros((9, 9), np.uint8))
    _check(host_harness, np.full((9, 9), 7, np.uint8))
    m, _, _ = synth.make_schematic(0, 1024)
    k = np.ones((3, 3), np.uint8)
    e = cv2.erode(cv2.dilate(cv2.GaussianBlur(cv2.resize(m, (600, 600)), (5, 5), 1), k, iterations=2), k, iterations=2)
    _check(host_harness, e)


def test_point_near_box_matches_reference_rule(host_harness):
    from oracle.node_oracle import is_point_near_bbox
    rng = np.random.default_rng(3)
    for _ in range(2000):
        px, py = (int(v) for v in rng.integers(-5, 60, 2))
        x0, y0 = (int(v) for v in rng.integers(0, 40, 2))
        x1, y1 = x0 + int(rng.integers(0, 20)), y0 + int(rng.integers(0, 20))
        t = int(rng.choice([6, 8, 20]))
        b = {"xmin": x0, "ymin": y0, "xmax": x1, "ymax": y1}
        assert bool(host_harness.hh_point_near_box(px, py, x0, y0, x1, y1, t)) == is_point_near_bbox(px, py, b, t)


def test_segment_circuit_bit_exact(host_harness):
    """grey conversion (incl. the reference's RGB/BGR swap, circuit_analyzer.py:2231+316) and the 31/21 adaptive
    threshold, executed from node_prims.cuh on the host, against cv2."""
    rng = np.random.default_rng(3)
    for (h, w) in [(97, 131), (31, 31), (12, 40), (200, 64)]:
        rgb = rng.integers(0, 256, (h, w, 3), dtype=np.uint8)
        if h == 200:
            rgb = cv2.GaussianBlur(rgb, (0, 0), 3)
        out = np.empty((h, w), np.uint8)
        host_harness.hh_segment_circuit(rgb.ctypes.data_as(u8p), h, w, out.ctypes.data_as(u8p))
        bgr = cv2.cvtColor(rgb, cv2.COLOR_RGB2BGR)
        grey = cv2.cvtColor(bgr, cv2.COLOR_RGB2GRAY)
        ref = cv2.adaptiveThreshold(grey, 255, cv2.ADAPTIVE_THRESH_MEAN_C, cv2.THRESH_BINARY_INV, 31, 21)
        assert np.array_equal(out, ref), (h, w)
    # grey formula on a lattice of triples + random triples
    tri = rng.integers(0, 256, (4096, 3), dtype=np.uint8)
    ref = cv2.cvtColor(tri.reshape(64, 64, 3), cv2.COLOR_RGB2GRAY).reshape(-1)
    got = np.array([host_harness.hh_gray(int(a), int(b), int(c)) for a, b, c in tri], np.uint8)
    assert np.array_equal(got, ref)
