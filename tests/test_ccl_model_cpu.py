"""The lane-by-lane model of the warp-sequential CCL (oracle/ccl_scan_model.py, the algorithm csrc/ccl_tiles.cu implements)
against cv2.connectedComponents with canonical labels (1 + min linear index of the component)."""
import numpy as np
import pytest

from oracle import ccl_scan_model, node_oracle


def _cases():
    rng = np.random.default_rng(3)
    out = []
    for (h, w, p) in [(40, 1100, 0.5), (70, 2100, 0.35), (33, 64, 0.6), (65, 1024, 0.45), (20, 3000, 0.62)]:
        out.append((rng.random((h, w)) < p).astype(np.uint8) * 255)
    m = np.zeros((96, 2048), np.uint8)  # long runs through several words and across the 1024-pixel tile seam, U shapes
    m[5, 10:2040] = 1
    m[5:90, 10] = 1
    m[5:90, 2039] = 1
    m[89, 10:700] = 1
    m[40, 1000:1060] = 1
    m[41:60, 1023] = 1
    m[41:60, 1024] = 1
    m[70:75, 960:1100] = 1
    out.append(m)
    m2 = np.ones((64, 1088), np.uint8)  # everything set: full words everywhere
    m2[31:33, 500] = 0
    out.append(m2)
    d = np.zeros((70, 1060), np.uint8)  # diagonals only (8-connected staircase across word, row-tile and column-tile seams)
    for i in range(70):
        d[i, 990 + i] = 1
    d[10, 5], d[11, 4], d[12, 3] = 1, 1, 1
    out.append(d)
    return out


@pytest.mark.parametrize("conn", [8, 4])
def test_model_matches_cv2(conn):
    for k, m in enumerate(_cases()):
        want = node_oracle.ccl_labels_min_index(m, conn)
        got = ccl_scan_model.label(m, conn)
        assert np.array_equal(got, want), (k, conn, int((got != want).sum()))
