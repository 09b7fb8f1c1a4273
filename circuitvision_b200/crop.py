"""YOLO-cluster crop of the page (host side).

Behavioural mirror of `/root/reference/src/circuit_analyzer.py`:
  crop_image_and_adjust_bboxes          :937-1284   (called at analysis_pipeline.py:177 with padding=80)
  _are_bboxes_proximal_for_clustering   :892-928
  _component_has_nearby_text            :930-935
It decides WHICH window of the page goes through SAM 2.1 and shifts the boxes into it: box geometry on a few dozen
rectangles, microseconds of host work.  The pixel side of that step — cropping, the antialiased resize to 1024² and the
normalisation of the uint8 upload — already runs on the device inside `segment_with_sam2` (`cv_sam2_preprocess`); the
returned crop is a zero-copy NumPy view exactly like the reference's slice.  Parity: tests/test_crop_cpu.py compares
window, adjusted boxes and the whole `crop_debug_info` dict with the unmodified reference on randomized pages (build
container) and with committed fixtures (tests/golden/crop_golden.json).
"""
from __future__ import annotations

import math
from copy import deepcopy

NON_COMPONENTS = frozenset(["text", "junction", "crossover", "vss", "explanatory", "circuit"])  # :51
NOT_CLUSTERED = frozenset(["text", "explanatory", "circuit", "vss", "crossover"])                # :995
TEXT_INCLUSION_PADDING = 20      # :1193
TEXT_REACH = 150                 # :1199 expanded_check_padding
MAX_BASIS_FRACTION = 0.90        # :1176


def _rect(b):
    if isinstance(b, dict):
        return b["xmin"], b["ymin"], b["xmax"], b["ymax"]
    return b


def boxes_proximal(a, b, threshold=50) -> bool:
    """:892-928 — overlapping, or both axis gaps within `threshold`."""
    ax0, ay0, ax1, ay1 = _rect(a)
    bx0, by0, bx1, by1 = _rect(b)
    if not (ax1 < bx0 or ax0 > bx1 or ay1 < by0 or ay0 > by1):
        return True
    gap_x = bx0 - ax1 if ax1 < bx0 else (ax0 - bx1 if ax0 > bx1 else 0)
    gap_y = by0 - ay1 if ay1 < by0 else (ay0 - by1 if ay0 > by1 else 0)
    return gap_x <= threshold and gap_y <= threshold


def has_nearby_text(component, texts, threshold=30) -> bool:
    """:930-935"""
    return any(boxes_proximal(component, t, threshold) for t in texts)


def _mean_diagonal(boxes) -> float:
    w = sum(b["xmax"] - b["xmin"] for b in boxes) / len(boxes)
    h = sum(b["ymax"] - b["ymin"] for b in boxes) / len(boxes)
    return math.sqrt(w ** 2 + h ** 2)


def _clusters(elements, threshold):
    """Connected groups under `boxes_proximal`, members in the reference's depth-first visiting order (:1024-1041: the
    neighbour pushed LAST is visited first)."""
    n = len(elements)
    near = [[] for _ in range(n)]
    for i in range(n):
        for j in range(i + 1, n):
            if boxes_proximal(elements[i], elements[j], threshold):
                near[i].append(j)
                near[j].append(i)
    seen = [False] * n
    groups = []
    for root in range(n):
        if seen[root]:
            continue
        members, stack = [], [root]
        while stack:
            u = stack.pop()
            if seen[u]:
                continue
            seen[u] = True
            members.append(elements[u])
            stack.extend(v for v in near[u] if not seen[v])
        groups.append(members)
    return groups


def crop_image_and_adjust_bboxes(image_to_crop, all_yolo_bboxes_input, padding=20, non_components=NON_COMPONENTS):
    """:937-1284.  Returns (cropped image view | the input image, adjusted bbox list, crop_debug_info)."""
    page_h, page_w = image_to_crop.shape[:2]
    boxes = all_yolo_bboxes_input
    info = {
        "crop_applied": False, "reason_for_no_crop": None, "original_image_dims": (page_w, page_h),
        "num_total_yolo_bboxes": len(boxes), "num_component_type_bboxes": 0, "num_text_type_bboxes": 0,
        "clustering_proximity_threshold": None, "num_clusters_found": None, "main_cluster_info": None,
        "crop_decision_source": "unknown", "crop_basis_bbox_before_padding": None, "padding_value": padding,
        "window_after_main_padding": None, "text_bboxes_that_expanded_crop": [], "final_crop_window_abs": None,
        "cropped_image_dims": (page_w, page_h),
    }

    def unchanged(reason):
        info["reason_for_no_crop"] = reason
        return image_to_crop, [deepcopy(b) for b in boxes], info

    texts = [b for b in boxes if b.get("class") == "text"]
    info["num_component_type_bboxes"] = sum(1 for b in boxes if b.get("class") not in non_components)
    info["num_text_type_bboxes"] = len(texts)
    elements = [b for b in boxes if b.get("class") not in NOT_CLUSTERED]
    if not elements:
        info["crop_decision_source"] = "no_crop_due_to_no_clustering_elements"
        return unchanged("no_elements_for_clustering")

    # :1002-1020 — the clustering radius follows the mean component diagonal
    sized = [e for e in elements if e.get("class") != "junction"]
    if sized:
        diag = _mean_diagonal(sized)
        radius = max(int(diag * 2.0), 30)
    else:
        diag = _mean_diagonal(elements)
        radius = max(int(diag * 2.5), 20)
    info["clustering_proximity_threshold"] = radius
    groups = _clusters(elements, radius)
    info["num_clusters_found"] = len(groups)

    # :1056-1150 — clusters ranked by (components with a text label nearby, size); without any labelled component in
    # the best one the largest cluster wins
    text_radius = max(int((diag if diag > 0 else 30) * 0.75), 25)
    ranked = []
    for gid, members in enumerate(groups):
        parts = [b for b in members if b.get("class") != "junction"]
        labelled = sum(1 for b in parts if has_nearby_text(b, texts, text_radius))
        ranked.append({"bboxes": members, "score": (labelled, len(members)), "id": gid, "text_assoc_count": labelled,
                       "total_elements_in_cluster": len(members), "actual_components_in_cluster": len(parts)})
    ranked.sort(key=lambda c: c["score"], reverse=True)
    best = ranked[0]
    if best["text_assoc_count"] == 0 and best["actual_components_in_cluster"] > 0:
        chosen = max(groups, key=len)
        info["crop_decision_source"] = "main_cluster_fallback_no_text_assoc_in_best_with_components"
        rec = next((c for c in ranked if c["bboxes"] == chosen), None) or best
    else:
        chosen, rec = best["bboxes"], best
        info["crop_decision_source"] = "main_yolo_cluster_scored_by_text_assoc"
    info["main_cluster_info"] = {"num_elements": len(chosen), "text_assoc_count": rec["text_assoc_count"],
                                 "score": rec["score"], "id": rec["id"],
                                 "example_uid": chosen[0].get("persistent_uid")}
    basis = (min(b["xmin"] for b in chosen), min(b["ymin"] for b in chosen),
             max(b["xmax"] for b in chosen), max(b["ymax"] for b in chosen))
    info["crop_basis_bbox_before_padding"] = basis

    # :1170-1179 — a basis that already covers the page is not worth a crop
    bx0, by0, bx1, by1 = basis
    page_area = float(page_h * page_w)
    if page_area > 0 and (float(max(0, bx1 - bx0)) * float(max(0, by1 - by0))) / page_area > MAX_BASIS_FRACTION:
        return unchanged("crop_basis_bbox_too_large")

    x0, y0 = float(max(0, bx0 - padding)), float(max(0, by0 - padding))
    x1, y1 = float(min(page_w, bx1 + padding)), float(min(page_h, by1 + padding))
    info["window_after_main_padding"] = (int(round(x0)), int(round(y0)), int(round(x1)), int(round(y1)))
    # :1195-1223 — text boxes within reach of the window pull it outwards (in list order: the window grows as it goes)
    for t in texts:
        tx0, ty0, tx1, ty1 = float(t["xmin"]), float(t["ymin"]), float(t["xmax"]), float(t["ymax"])
        if tx1 < x0 - TEXT_REACH or tx0 > x1 + TEXT_REACH or ty1 < y0 - TEXT_REACH or ty0 > y1 + TEXT_REACH:
            continue
        grown = (min(x0, max(0, tx0 - TEXT_INCLUSION_PADDING)), min(y0, max(0, ty0 - TEXT_INCLUSION_PADDING)),
                 max(x1, min(page_w, tx1 + TEXT_INCLUSION_PADDING)), max(y1, min(page_h, ty1 + TEXT_INCLUSION_PADDING)))
        if grown != (x0, y0, x1, y1):
            info["text_bboxes_that_expanded_crop"].append({
                "uid": t.get("persistent_uid"), "class": t.get("class"),
                "coords_original": (t["xmin"], t["ymin"], t["xmax"], t["ymax"]),
                "coords_text_box_abs": (tx0, ty0, tx1, ty1)})
        x0, y0, x1, y1 = grown

    wx0, wy0 = max(0, int(round(x0))), max(0, int(round(y0)))
    wx1, wy1 = min(page_w, int(round(x1))), min(page_h, int(round(y1)))
    info["final_crop_window_abs"] = (wx0, wy0, wx1, wy1)
    if wx0 >= wx1 or wy0 >= wy1:
        return unchanged("invalid_region_after_expansion")
    cropped = image_to_crop[wy0:wy1, wx0:wx1]
    new_h, new_w = cropped.shape[:2]
    info["cropped_image_dims"] = (new_w, new_h)
    info["crop_applied"] = True

    # :1259-1281 — every box moves into the window, is clipped to it and dropped when nothing is left
    moved = []
    for b in boxes:
        m = deepcopy(b)
        m["xmin"], m["ymin"] = max(0, b["xmin"] - wx0), max(0, b["ymin"] - wy0)
        m["xmax"], m["ymax"] = min(new_w, b["xmax"] - wx0), min(new_h, b["ymax"] - wy0)
        if m["xmax"] > m["xmin"] and m["ymax"] > m["ymin"]:
            moved.append(m)
    return cropped, moved, info
