"""TEST INFRASTRUCTURE ONLY.  Generates tests/golden/node_golden.npz by running the UNMODIFIED reference
`CircuitAnalyzer.get_node_connections` + `generate_netlist_from_nodes` (via oracle/ref_loader.py) on the
deterministic cases of `golden_cases()`.  Runs only in the build container (needs /root/reference).

    python -m oracle.gen_golden
"""
from __future__ import annotations

import hashlib
import json
import os
import sys

import cv2
import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from circuitvision_b200 import synth  # noqa: E402

GOLDEN = os.path.join(ROOT, "tests", "golden", "node_golden.npz")
GOLDEN_TERMINALS = os.path.join(ROOT, "tests", "golden", "terminal_golden.npz")
CLASS_NAMES = {4: "terminal", 7: "voltage.dc", 10: "resistor"}  # subset of the reference's classes.json ids


def _box(cls, x0, y0, x1, y1):
    return {"class": cls, "confidence": 0.5, "xmin": x0, "ymin": y0, "xmax": x1, "ymax": y1,
            "persistent_uid": f"{cls}_{round(x0)}_{round(y0)}_{round(x1)}_{round(y1)}"}


def golden_cases():
    """name -> (mask u8 [H,W], boxes).  Fully deterministic; no file reads (must regenerate on the GPU box)."""
    cases = {}
    for s in range(6):
        m, b, _ = synth.make_schematic(s, 1024)
        cases[f"schem1024_s{s}"] = (m, b)
    m, b, _ = synth.make_schematic(11, 2048, grid=12, dense=True)
    cases["dense2048_s11"] = (m, b)
    m, b, _ = synth.make_schematic(12, 4096)
    cases["dense4096_s12"] = (m, b)
    m, b, _ = synth.make_schematic(13, 4096, grid=24, dense=False)
    cases["grid24_4096_s13"] = (m, b)
    # non-square crops (aspect != 1 exercises new_w = int(600*W/H))
    m, b, _ = synth.make_schematic(3, 1024)
    cases["crop_720x1000"] = (np.ascontiguousarray(m[100:820, 10:1010]),
                              [dict(x, xmin=x["xmin"] - 10, xmax=x["xmax"] - 10, ymin=x["ymin"] - 100,
                                    ymax=x["ymax"] - 100) for x in b])
    cases["crop_1000x640"] = (np.ascontiguousarray(m[12:1012, 200:840]),
                              [dict(x, xmin=x["xmin"] - 200.5, xmax=x["xmax"] - 200.25, ymin=x["ymin"] - 12.75,
                                    ymax=x["ymax"] - 12.0) for x in b])  # float coords + boxes off the crop
    # edge cases
    cases["empty_mask"] = (np.zeros((512, 512), np.uint8), [_box("resistor", 10, 10, 60, 40)])
    m, b, _ = synth.make_schematic(4, 1024)
    cases["no_boxes"] = (m, [])
    cases["only_noncomponents"] = (m, [x for x in b if x["class"] in ("text", "junction")])
    inv = 255 - m  # mostly white => get_contours inversion branch (circuit_analyzer.py:398)
    cases["inverted_white"] = (inv, [x for x in b if x["class"] == "junction"])
    # noisy hand-drawn-like mask: blobs + nested rings + spurs, sources only => ground fallbacks
    rb = synth.random_blob_mask(5, 700, 900, p=0.52, smooth=3)
    rb[100:300, 100:400] = 255
    rb[130:270, 130:370] = 0
    rb[170:230, 200:300] = 255
    cases["blobs_700x900"] = (rb, [_box("resistor", 90, 180, 140, 220), _box("capacitor.unpolarized", 380, 150, 420, 200),
                                   _box("diode", 600, 300, 660, 340), _box("transistor.bjt", 300, 500, 380, 560),
                                   _box("resistor", 90, 180, 140, 220),  # duplicate uid
                                   _box("terminal", 500, 100, 520, 120), _box("gnd", 700, 600, 740, 640)])
    m, b, _ = synth.make_schematic(5, 1024)
    cases["two_valid_nodes"] = (m, b[:2])
    cases["single_component"] = (m, b[:1])
    # values other than {0,255} in the mask (any non-zero is foreground downstream)
    m2 = (m // 255) * 37
    cases["mask_value_37"] = (m2.astype(np.uint8), b)
    # the two real schematics the reference ships (hand-drawn photo: noisy mask, many vertices; printed bridge), mask =
    # segment_circuit(page), hand-written boxes (oracle/real_cases.py; pages travel as tests/golden/real_pages.npz)
    from oracle import real_cases
    for name, (rgb, mask, boxes) in real_cases.real_cases().items():
        cases[f"real_{name}"] = (mask, boxes)
    return cases


def terminal_cases():
    """name -> (rgb u8 [H,W,3], boxes with 'terminal' entries).  Deterministic (seeded); no file reads."""
    cases = {}

    def paper(seed, size, noise, grid=None):
        m, b, rgb = synth.make_schematic(seed, size, grid=grid, dense=False, render_rgb=True)
        rng = np.random.default_rng(7000 + seed)
        img = rgb.astype(np.float32)
        if noise:
            yy, xx = np.mgrid[0:size, 0:size].astype(np.float32)
            shade = 30.0 * np.sin(xx / size * 3.1) * np.cos(yy / size * 2.3)  # uneven illumination
            img = img * 0.85 + shade[..., None] + rng.normal(0.0, noise, img.shape).astype(np.float32)
        return m, b, np.clip(img, 0, 255).astype(np.uint8), rng

    def add_terminals(m, b, rng, n, size):
        ys, xs = np.nonzero(m)
        out = list(b)
        for k in range(n):
            j = int(rng.integers(len(ys)))
            cx, cy = int(xs[j]), int(ys[j])
            w, h = int(rng.integers(8, 40)), int(rng.integers(8, 40))
            if k % 5 == 4:  # off-wire terminal
                cx, cy = int(rng.integers(size)), int(rng.integers(size))
            out.append(_box("terminal", cx - w // 2, cy - h // 2, cx + w - w // 2, cy + h - h // 2))
        return out

    for seed, size, noise in ((0, 1024, 0.0), (1, 1024, 6.0), (2, 1024, 12.0), (3, 768, 6.0)):
        m, b, rgb, rng = paper(seed, size, noise)
        cases[f"paper{size}_s{seed}_n{int(noise)}"] = (rgb, add_terminals(m, b, rng, 10, size))
    m, b, rgb, rng = paper(4, 2048, 5.0, grid=10)
    cases["paper2048_s4"] = (rgb, add_terminals(m, b, rng, 24, 2048))
    # non-square crop, terminals clipped by the image border, duplicate terminal, boxes partly outside
    m, b, rgb, rng = paper(5, 1024, 4.0)
    bb = add_terminals(m, b, rng, 8, 1024)
    crop = np.ascontiguousarray(rgb[40:800, 100:1000])
    bb = [dict(x, xmin=x["xmin"] - 100, xmax=x["xmax"] - 100, ymin=x["ymin"] - 40, ymax=x["ymax"] - 40) for x in bb]
    bb.append(_box("terminal", -12, 300, 14, 330))
    bb.append(_box("terminal", 880, 740, 930, 790))
    bb.append(dict(bb[-1]))
    cases["crop_760x900"] = (crop, bb)
    # no terminals at all / only terminals / blank page / inverted page (dark paper: mean of the mask > 127)
    m, b, rgb, rng = paper(6, 1024, 3.0)
    cases["no_terminals"] = (rgb, b)
    cases["only_terminals"] = (rgb, [x for x in add_terminals(m, [], rng, 12, 1024)])
    cases["blank_page"] = (np.full((512, 640, 3), 250, np.uint8), [_box("terminal", 100, 100, 140, 130)])
    chk = np.indices((600, 600)).sum(0) % 2 * 255  # checkerboard: the adaptive threshold marks half the pixels
    cases["checkerboard"] = (np.repeat(chk[..., None], 3, 2).astype(np.uint8), [_box("terminal", 200, 200, 260, 240),
                                                                               _box("resistor", 300, 300, 420, 340)])
    # channel-dependent colours: exercises the RGB/BGR swap of segment_circuit
    col = np.zeros((512, 512, 3), np.uint8)
    col[..., 0] = 200
    col[..., 2] = 40
    col[100:110, 50:450] = (10, 10, 250)
    col[300:310, 50:450] = (250, 10, 10)
    col[100:310, 240:250] = (10, 250, 10)
    cases["colour_channels"] = (col, [_box("terminal", 230, 190, 262, 222), _box("terminal", 40, 90, 70, 120)])
    from oracle import real_cases
    for name, (rgb, mask, boxes) in real_cases.real_cases().items():
        cases[f"real_{name}"] = (rgb, boxes)
    return cases


def main_terminals():
    from oracle import ref_loader
    store, meta = {}, {}
    for name, (rgb, boxes) in terminal_cases().items():
        out, n_contours, counts = ref_loader.reference_reclassify(rgb, boxes, CLASS_NAMES)
        A = ref_loader.load_reference_analyzer()
        mask = A.segment_circuit(cv2.cvtColor(rgb, cv2.COLOR_RGB2BGR))
        meta[name] = {"shape": list(rgb.shape), "rgb_sha256": _sha(rgb), "mask_sha256": _sha(mask),
                      "n_contours": n_contours, "terminal_counts": counts,
                      "classes": [b["class"] for b in out],
                      "reclassified": [bool(b.get("was_reclassified_from_terminal", False)) for b in out],
                      "yolo_ids": [b.get("_yolo_class_id_temp") for b in out],
                      "orig": [b.get("original_yolo_class_if_reclassified") for b in out]}
        print(name, rgb.shape, len(boxes), "boxes ->", sum(meta[name]["reclassified"]), "reclassified of",
              sum(b["class"] == "terminal" for b in boxes), "terminals")
    store["meta_json"] = np.frombuffer(json.dumps(meta, sort_keys=True).encode(), np.uint8)
    np.savez_compressed(GOLDEN_TERMINALS, **store)
    print("wrote", GOLDEN_TERMINALS, os.path.getsize(GOLDEN_TERMINALS), "bytes; cv2", cv2.__version__)


GOLDEN_CROPS = os.path.join(ROOT, "tests", "golden", "crop_golden.json")


def _jsonable(x):
    return json.loads(json.dumps(x, default=lambda o: o.item() if hasattr(o, "item") else str(o)))


def main_crops():
    """tests/golden/crop_golden.json: the unmodified reference crop_image_and_adjust_bboxes (circuit_analyzer.py:937) on
    the seeded pages of oracle/crop_cases.py, padding 80 (analysis_pipeline.py:181) and the default 20."""
    import contextlib
    import copy
    import io
    from oracle import ref_loader
    from oracle.crop_cases import random_page
    A = ref_loader.load_reference_analyzer()
    out = {}
    for seed in range(84):
        img, boxes = random_page(seed)
        for pad in (80, 20):
            with contextlib.redirect_stdout(io.StringIO()):
                rimg, rb, rinfo = A.crop_image_and_adjust_bboxes(img, copy.deepcopy(boxes), padding=pad)
            out[f"{seed}/{pad}"] = {"shape": list(rimg.shape), "sha": _sha(rimg), "boxes": _jsonable(rb), "info": _jsonable(rinfo)}
    json.dump(out, open(GOLDEN_CROPS, "w"), sort_keys=True)
    n = sum(1 for v in out.values() if v["info"]["crop_applied"])
    print("wrote", GOLDEN_CROPS, os.path.getsize(GOLDEN_CROPS), "bytes;", n, "of", len(out), "calls cropped")


def _sha(a: np.ndarray) -> str:
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def main():
    from oracle import ref_loader
    store = {}
    meta = {}
    for name, (mask, boxes) in golden_cases().items():
        nodes, emptied, enhanced, cimg, fviz, cpts, text = ref_loader.reference_node_analysis(mask, boxes)
        # the connection points themselves (cyan discs of the third drawing): recovered from the oracle restatement,
        # which the test suite pins to these very fixtures
        from oracle import node_oracle
        conn_pts = node_oracle.get_node_connections(mask, boxes)[4]
        meta[name] = {
            "shape": list(mask.shape),
            "n_boxes": len(boxes),
            "nodes": [{"id": int(n["id"]), "uids": [c["persistent_uid"] for c in n["components"]],
                       "comp_xyxy": [[c["xmin"], c["ymin"], c["xmax"], c["ymax"]] for c in n["components"]]}
                      for n in nodes],
            "netlist": text,
            "emptied_sha256": _sha(emptied),
            "enhanced_sha256": _sha(enhanced),
            "enhanced_shape": list(enhanced.shape),
            "mask_sha256": _sha(mask),
            # the three debug drawings the reference returns (:414-458, :1585-1603), byte for byte
            "contour_viz_sha256": _sha(cimg), "final_viz_sha256": _sha(fviz), "points_viz_sha256": _sha(cpts),
            "connection_points": [[int(p[0]), int(p[1])] for p in conn_pts],
        }
        for n in nodes:
            store[f"{name}/contour{int(n['id'])}"] = np.asarray(n["contour"], np.int32).reshape(-1, 2)
        print(name, mask.shape, len(boxes), "boxes ->", len(nodes), "nodes")
    store["meta_json"] = np.frombuffer(json.dumps(meta, sort_keys=True).encode(), np.uint8)
    os.makedirs(os.path.dirname(GOLDEN), exist_ok=True)
    np.savez_compressed(GOLDEN, **store)
    print("wrote", GOLDEN, os.path.getsize(GOLDEN), "bytes; cv2", cv2.__version__)


if __name__ == "__main__":
    if "--crops" in sys.argv:
        main_crops()
    elif "--terminals" in sys.argv:
        main_terminals()
    else:
        main()
