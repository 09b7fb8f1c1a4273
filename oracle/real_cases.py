"""TEST INFRASTRUCTURE ONLY.  The two real schematics the reference ships (`/root/reference/static/images/circuits_1.jpg`,
a 720x1280 hand-drawn photo, and `Unbalanced_Wheatstone_bridge.png`, 493x712) as parity inputs: the only non-synthetic
images available offline.  YOLO is absent, so the component boxes below are hand-written from the pictures (class names
of the reference's classes.json).  Mask source for the node analysis = the reference's `segment_circuit`
(circuit_analyzer.py:313-319), as in its terminal-reclassification step.

`make_real_pages()` (build container only) decodes the two files with cv2 exactly as the app does (cv2.imread ->
COLOR_BGR2RGB, analysis_pipeline.py:20) and stores the RGB arrays losslessly (PNG bytes) in tests/golden/real_pages.npz, so
the cases regenerate on the GPU box without /root/reference."""
from __future__ import annotations

import os

import cv2
import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PAGES = os.path.join(ROOT, "tests", "golden", "real_pages.npz")
SOURCES = {"circuits_1": "circuits_1.jpg", "wheatstone": "Unbalanced_Wheatstone_bridge.png"}


def _b(cls, x0, y0, x1, y1):
    return {"class": cls, "confidence": 0.9, "xmin": x0, "ymin": y0, "xmax": x1, "ymax": y1,
            "persistent_uid": f"{cls}_{x0}_{y0}_{x1}_{y1}"}


BOXES = {
    # hand-drawn loop: 5 V source, 3 ohm and 10 ohm resistors, 2 A current source, four text labels
    "circuits_1": [
        _b("voltage.dc", 208, 348, 352, 484), _b("resistor", 432, 196, 532, 252), _b("resistor", 676, 318, 728, 432),
        _b("current.dc", 864, 332, 998, 432), _b("text", 436, 112, 584, 192), _b("text", 98, 380, 206, 474),
        _b("text", 562, 362, 682, 428), _b("text", 1002, 346, 1112, 408),
        _b("terminal", 684, 214, 716, 246), _b("terminal", 262, 596, 296, 630), _b("terminal", 40, 40, 70, 70),
    ],
    # printed Wheatstone bridge: battery, R1..R5, four junction dots, labels
    "wheatstone": [
        _b("voltage.dc", 84, 204, 152, 302), _b("resistor", 384, 134, 446, 196), _b("resistor", 544, 134, 606, 196),
        _b("resistor", 452, 228, 528, 264), _b("resistor", 384, 304, 446, 366), _b("resistor", 544, 304, 606, 366),
        _b("junction", 480, 68, 506, 94), _b("junction", 320, 234, 346, 260), _b("junction", 642, 234, 668, 260),
        _b("junction", 480, 404, 506, 430), _b("text", 312, 92, 398, 166), _b("text", 596, 92, 668, 158),
        _b("text", 446, 192, 534, 228), _b("text", 322, 322, 412, 392), _b("text", 572, 322, 662, 392),
        _b("text", 10, 236, 78, 268), _b("terminal", 104, 22, 132, 50), _b("terminal", 484, 452, 504, 482),
    ],
}


def make_real_pages(ref_root="/root/reference"):
    out = {}
    for name, fn in SOURCES.items():
        bgr = cv2.imread(os.path.join(ref_root, "static", "images", fn))
        assert bgr is not None, fn
        rgb = cv2.cvtColor(bgr, cv2.COLOR_BGR2RGB)
        ok, png = cv2.imencode(".png", rgb)  # lossless container for the decoded pixels (channel order kept as stored)
        assert ok
        out[name] = png
    np.savez(PAGES, **out)
    return PAGES


def load_pages():
    """name -> (H,W,3) uint8 RGB page, identical on every box."""
    z = np.load(PAGES)
    return {k: cv2.imdecode(z[k], cv2.IMREAD_UNCHANGED) for k in SOURCES}


def real_cases():
    """name -> (rgb page, wire mask = segment_circuit(page), boxes)."""
    from .terminal_oracle import segment_circuit_from_rgb
    out = {}
    for name, rgb in load_pages().items():
        out[name] = (rgb, segment_circuit_from_rgb(rgb), [dict(b) for b in BOXES[name]])
    return out
