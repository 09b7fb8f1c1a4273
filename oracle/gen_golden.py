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
    return cases


def _sha(a: np.ndarray) -> str:
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def main():
    from oracle import ref_loader
    store = {}
    meta = {}
    for name, (mask, boxes) in golden_cases().items():
        nodes, emptied, enhanced, cimg, fviz, cpts, text = ref_loader.reference_node_analysis(mask, boxes)
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
        }
        for n in nodes:
            store[f"{name}/contour{int(n['id'])}"] = np.asarray(n["contour"], np.int32).reshape(-1, 2)
        print(name, mask.shape, len(boxes), "boxes ->", len(nodes), "nodes")
    store["meta_json"] = np.frombuffer(json.dumps(meta, sort_keys=True).encode(), np.uint8)
    os.makedirs(os.path.dirname(GOLDEN), exist_ok=True)
    np.savez_compressed(GOLDEN, **store)
    print("wrote", GOLDEN, os.path.getsize(GOLDEN), "bytes; cv2", cv2.__version__)


if __name__ == "__main__":
    main()
