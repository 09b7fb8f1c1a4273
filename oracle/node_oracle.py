"""TEST INFRASTRUCTURE ONLY — CPU restatement of the reference's node/connection analysis.

Follows `/root/reference/src/circuit_analyzer.py`:
  get_node_connections      :1286-1605
  resize_image_keep_aspect  :787-809      resize_bboxes :461-477
  enhance_lines             :289-311      get_contours  :388-412 (visualisation 414-458 omitted)
  is_point_near_bbox        :811-846
The pixel arithmetic lives in OpenCV (`cv2` 4.13.0 in this image, the reference's own dependency,
requirements.txt), which this restatement calls exactly where the reference calls it.

Parity pinning: the reference ships no tests or golden vectors for this path, so the
restatement is pinned by *executing the unmodified reference function* (oracle/ref_loader.py) in the build
container on the seeded inputs of oracle/gen_golden.py; the outputs are committed under tests/golden/ and
tests/test_oracle_golden.py checks this file against them (bit-exact nodes / contours / images / netlist).

Only tests/, __graft_entry__.smoke() and bench.py's CPU-baseline legs may import this module.
"""
from __future__ import annotations

from copy import deepcopy

import cv2
import numpy as np

NON_COMPONENTS = frozenset(["text", "junction", "crossover", "vss", "explanatory", "circuit"])  # :51
SOURCE_COMPONENTS = frozenset(
    ["voltage.ac", "voltage.dc", "voltage.dependent", "current.dc", "current.dependent"])  # :52
PRESERVE_IN_MASK = ("crossover", "junction", "circuit", "vss")  # :1326
THRESH_8 = ("diode", "diode.light_emitting", "diode.zener", "transistor.bjt", "transistor.fet")  # :1411


def empty_boxes(mask: np.ndarray, boxes) -> np.ndarray:
    """:1308-1345 — copy the mask and zero every box whose class is not preserved."""
    emptied = mask.copy()
    H, W = emptied.shape[:2]
    for b in boxes:
        if b["class"] not in PRESERVE_IN_MASK:
            ymin, ymax = max(0, int(b["ymin"])), min(H, int(b["ymax"]))
            xmin, xmax = max(0, int(b["xmin"])), min(W, int(b["xmax"]))
            if ymin < ymax and xmin < xmax:
                emptied[ymin:ymax, xmin:xmax] = 0
    return emptied


def resize_bboxes(boxes, width_scale, height_scale):
    """:461-477"""
    out = []
    for b in boxes:
        r = b.copy()
        r["xmin"] = int(b["xmin"] * width_scale)
        r["ymin"] = int(b["ymin"] * height_scale)
        r["xmax"] = int(b["xmax"] * width_scale)
        r["ymax"] = int(b["ymax"] * height_scale)
        out.append(r)
    return out


def resize_keep_aspect(image, boxes, new_height=600):
    """:787-809"""
    h, w = image.shape[:2]
    new_w = int(new_height * (w / h))
    resized = cv2.resize(image, (new_w, new_height))
    return resized, resize_bboxes(boxes, new_w / w, new_height / h)


def enhance_lines(img):
    """:289-311"""
    blurred = cv2.GaussianBlur(img, (5, 5), 1)
    k = np.ones((3, 3), np.uint8)
    return cv2.erode(cv2.dilate(blurred, k, iterations=2), k, iterations=2)


def get_contours(img, area_threshold=0.0004):
    """:388-412 — NB mutates `img` (255→1) unless the inversion branch made a fresh array."""
    if cv2.mean(img)[0] > 127:
        img = 255 - img
    img[img == 255] = 1
    contours, _ = cv2.findContours(img, cv2.RETR_EXTERNAL, cv2.CHAIN_APPROX_SIMPLE)
    normalizer = img.shape[0] * img.shape[1]
    contours = [c for c in contours if cv2.contourArea(c) / normalizer > area_threshold]
    return [{"id": i, "contour": c, "area": cv2.contourArea(c) / normalizer, "rectangle": cv2.boundingRect(c)}
            for i, c in enumerate(contours)]


def is_point_near_bbox(px, py, b, t):
    """:811-846 — distance to the four infinite edge LINES (reference quirk)."""
    if b["xmin"] <= px <= b["xmax"] and b["ymin"] <= py <= b["ymax"]:
        return True
    return (abs(px - b["xmin"]) <= t or abs(px - b["xmax"]) <= t or
            abs(py - b["ymin"]) <= t or abs(py - b["ymax"]) <= t)


def pixel_threshold(cls: str) -> int:
    """:1404-1415"""
    if cls in SOURCE_COMPONENTS:
        return 20
    if cls in THRESH_8:
        return 8
    return 6


def _centroid_y(contour):
    M = cv2.moments(contour)
    if M["m00"] != 0:
        return int(M["m01"] / M["m00"])
    return -float("inf")


def get_node_connections(mask, boxes):
    """:1286-1605 minus visualisations.  Returns (new_nodes_list, emptied, enhanced, resized, conn_points)."""
    emptied = empty_boxes(mask, boxes)
    resized, rboxes = resize_keep_aspect(emptied, boxes)
    enhanced = enhance_lines(resized)
    contours = get_contours(enhanced)
    nodes = {c["id"]: {"id": c["id"], "components": [], "contour": c["contour"]} for c in contours}
    conn_points = []
    for i, rb in enumerate(rboxes):  # :1380
        if rb["class"] in NON_COMPONENTS:
            continue
        t = pixel_threshold(rb["class"])
        for c in contours:
            cx, cy, cw, ch = c["rectangle"]
            if rb["xmax"] < cx or rb["xmin"] > cx + cw or rb["ymax"] < cy or rb["ymin"] > cy + ch:
                continue
            for p in c["contour"]:
                px, py = int(p[0][0]), int(p[0][1])
                if is_point_near_bbox(px, py, rb, t):
                    tgt = deepcopy(rboxes[i])
                    ref = tgt.get("persistent_uid")
                    if ref is None:
                        ref = (tgt["class"], tgt["xmin"], tgt["ymin"], tgt["xmax"], tgt["ymax"])
                    present = False
                    for e in nodes[c["id"]]["components"]:
                        er = e.get("persistent_uid")
                        if er is None:
                            er = (e["class"], e["xmin"], e["ymin"], e["xmax"], e["ymax"])
                        if er == ref:
                            present = True
                            break
                    if not present:
                        nodes[c["id"]]["components"].append(tgt)
                        conn_points.append((px, py))
                    break
    valid = {k: v for k, v in nodes.items() if v["components"]}  # :1451
    if not valid:
        return [], emptied, enhanced, resized, conn_points
    max_conn = max(len(v["components"]) for v in valid.values())
    with_max = [k for k, v in valid.items() if len(v["components"]) == max_conn]
    by_id = {c["id"]: c for c in contours}
    ground = None
    cands = []
    for k, v in valid.items():  # :1475
        if any(comp["class"] in SOURCE_COMPONENTS for comp in v["components"]):
            cands.append({"id": k, "centroid_y": _centroid_y(by_id[k]["contour"])})
    if cands:
        cands.sort(key=lambda x: x["centroid_y"], reverse=True)
        ground = cands[0]["id"]
    else:  # :1499-1545
        if with_max:
            if len(with_max) > 1:
                det = [{"id": k, "centroid_y": _centroid_y(by_id[k]["contour"])} for k in with_max]
                det.sort(key=lambda x: x["centroid_y"], reverse=True)
                ground = det[0]["id"]
            else:
                ground = with_max[0]
        if ground is None:
            ground = list(valid.keys())[0]
    out = [{"id": 0, "components": valid[ground]["components"], "contour": valid[ground]["contour"]}]
    nxt = 1
    for k in sorted(x for x in valid if x != ground):  # :1558
        v = valid[k]
        if len(v["components"]) >= 2 or (len(out) == 1 and len(valid) == 2 and len(v["components"]) > 0):
            out.append({"id": nxt, "components": v["components"], "contour": v["contour"]})
            nxt += 1
    return out, emptied, enhanced, resized, conn_points


def node_signature(nodes):
    """Bit-exact comparable form: [(id, [uids...], contour int32 array)]."""
    return [(int(n["id"]), [c.get("persistent_uid") for c in n["components"]],
             np.asarray(n["contour"], np.int32).reshape(-1, 2)) for n in nodes]


def ccl_labels_min_index(mask: np.ndarray, connectivity: int = 8) -> np.ndarray:
    """Oracle for the native-resolution CCL kernel (BASELINE cfg 4; not a reference code path):
    cv2.connectedComponents relabelled canonically — label = 1 + min linear index of the component, 0 = bg."""
    n, lab = cv2.connectedComponents((mask != 0).astype(np.uint8), connectivity=connectivity, ltype=cv2.CV_32S)
    H, W = mask.shape
    idx = np.arange(H * W, dtype=np.int64).reshape(H, W)
    mins = np.full(n, H * W, np.int64)
    np.minimum.at(mins, lab.ravel(), idx.ravel())
    out = (mins[lab] + 1).astype(np.int32)
    out[lab == 0] = 0
    return out
