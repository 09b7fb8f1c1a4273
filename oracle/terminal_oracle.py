"""TEST INFRASTRUCTURE ONLY — CPU restatement of the reference's preliminary terminal reclassification.

Follows `/root/reference/src/circuit_analyzer.py`:
  reclassify_terminals_based_on_connectivity :2217-2310
  segment_circuit                            :313-319   (cvtColor RGB2GRAY on the BGR copy + adaptiveThreshold 31 / 21)
  get_contours(area_threshold=0.0001)        :388-412
  is_point_near_bbox(..., 10)                :811-846
called from `/root/reference/src/analysis_pipeline.py:124-127` on the full, uncropped RGB image with the NMS'd YOLO
boxes (integer pixel coordinates, :276-287).

Parity pinning: the reference has no tests for this path; the restatement is pinned by executing the unmodified
reference method (oracle/ref_loader.py + oracle/gen_golden.py --terminals) on seeded inputs; the fixtures live in
tests/golden/terminal_golden.npz and tests/test_oracle_golden.py checks this file against them.

Only tests/, __graft_entry__.smoke() and bench.py's CPU-baseline legs may import this module.
"""
from __future__ import annotations

import cv2
import numpy as np

from .node_oracle import PRESERVE_IN_MASK, get_contours, is_point_near_bbox

RECLASS_THRESHOLD_PX = 10   # :2276
RECLASS_AREA_THRESHOLD = 0.0001  # :2252


def segment_circuit_from_rgb(image_rgb: np.ndarray) -> np.ndarray:
    """:2234 + :313-319.  The reference converts RGB->BGR and then applies COLOR_RGB2GRAY to that BGR array, i.e. the
    grey value weights the ORIGINAL red channel with the blue coefficient and vice versa."""
    bgr = cv2.cvtColor(image_rgb, cv2.COLOR_RGB2BGR)
    grey = cv2.cvtColor(bgr.copy(), cv2.COLOR_RGB2GRAY)
    return cv2.adaptiveThreshold(grey, 255, cv2.ADAPTIVE_THRESH_MEAN_C, cv2.THRESH_BINARY_INV, 31, 21)


def prelim_wire_mask(image_rgb: np.ndarray, boxes) -> np.ndarray:
    """:2238-2249 — adaptive-threshold mask with every box that is not preserved zeroed (slice semantics of NumPy:
    negative `ymax`/`xmax` after min() would wrap, the reference clamps only the lower bound with max(0, .))."""
    m = segment_circuit_from_rgb(image_rgb).copy()
    for b in boxes:
        if b.get("class") not in PRESERVE_IN_MASK:
            ymin, ymax = int(b["ymin"]), int(b["ymax"])
            xmin, xmax = int(b["xmin"]), int(b["xmax"])
            m[max(0, ymin):min(m.shape[0], ymax), max(0, xmin):min(m.shape[1], xmax)] = 0
    return m


def terminal_contact_counts(image_rgb: np.ndarray, boxes):
    """Returns (counts, mask, contours): counts[i] = number of distinct prelim contours with a vertex 'near' box i
    (only evaluated for class == 'terminal', -1 otherwise)."""
    mask = prelim_wire_mask(image_rgb, boxes)
    contours = get_contours(mask.copy(), RECLASS_AREA_THRESHOLD)
    counts = []
    for b in boxes:
        if b.get("class") != "terminal":
            counts.append(-1)
            continue
        n = 0
        for c in contours:
            for p in c["contour"]:
                if is_point_near_bbox(int(p[0][0]), int(p[0][1]), b, RECLASS_THRESHOLD_PX):
                    n += 1
                    break
        counts.append(n)
    return counts, mask, contours


def reclassify_terminals(image_rgb: np.ndarray, boxes, class_names=None):
    """:2217-2310 — modifies `boxes` in place exactly like the reference and returns the per-box contact counts."""
    counts, _, _ = terminal_contact_counts(image_rgb, boxes)
    vdc_id = None
    if class_names:
        for num_id, name in class_names.items():
            if name == "voltage.dc":
                vdc_id = num_id
                break
    for b, n in zip(boxes, counts):
        if n >= 2:
            b["original_yolo_class_if_reclassified"] = b["class"]
            b["class"] = "voltage.dc"
            if vdc_id is not None:
                b["_yolo_class_id_temp"] = vdc_id
            b["was_reclassified_from_terminal"] = True
    return counts
