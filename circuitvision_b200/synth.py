"""Deterministic synthetic "Manhattan schematic" generator.

Produces, for a seed (= image index): a clean wire mask (uint8 {0,255}), the list of
component bounding boxes in the dict format of the reference's YOLO wrapper
(`/root/reference/src/circuit_analyzer.py:267-287`, persistent_uid format of :285),
and an RGB rendering (white paper, black wires, glyph-filled component boxes) used as
the SAM 2.1 input crop.  Pure NumPy; used by tests, the benchmark and the oracle so
that both sides of every parity comparison see identical bytes.
"""
from __future__ import annotations

import numpy as np

COMPONENT_CLASSES = (
    "resistor",
    "capacitor.unpolarized",
    "inductor",
    "diode",
    "voltage.dc",
    "current.dc",
)


def _uid(cls: str, xmin: int, ymin: int, xmax: int, ymax: int) -> str:
    return f"{cls}_{round(xmin)}_{round(ymin)}_{round(xmax)}_{round(ymax)}"


def _box(cls: str, xmin: int, ymin: int, xmax: int, ymax: int, conf: float = 0.9) -> dict:
    return {
        "class": cls,
        "confidence": conf,
        "xmin": int(xmin),
        "ymin": int(ymin),
        "xmax": int(xmax),
        "ymax": int(ymax),
        "persistent_uid": _uid(cls, xmin, ymin, xmax, ymax),
    }


def make_schematic(seed: int, size: int = 1024, grid: int | None = None, dense: bool | None = None,
                   render_rgb: bool = False):
    """Return (mask[H,W] u8, boxes list[dict], rgb[H,W,3] u8 | None).

    grid: lattice of junction points (6 @1024², 24 for the "dense" 4096² variant).
    dense: adds 0.2 % salt noise (the cfg-4 variant).
    """
    S = int(size)
    if grid is None:
        grid = 6 if S <= 2048 else 24
    if dense is None:
        dense = S > 2048
    G = int(grid)
    rng = np.random.default_rng(seed)
    mask = np.zeros((S, S), np.uint8)
    th = max(3, S // 256)
    margin = S // (G + 1)
    step = (S - 2 * margin) // (G - 1)
    pts = [[(margin + i * step, margin + j * step) for i in range(G)] for j in range(G)]
    boxes: list[dict] = []
    h2 = th // 2

    def hline(y, x0, x1):
        mask[max(0, y - h2): y - h2 + th, x0: x1 + 1] = 255

    def vline(x, y0, y1):
        mask[y0: y1 + 1, max(0, x - h2): x - h2 + th] = 255

    long_side = max(8, step // 5)
    short_side = max(6, step // 8)
    for j in range(G):
        for i in range(G):
            x, y = pts[j][i]
            # horizontal edge to the right
            if i + 1 < G and rng.random() < 0.8:
                x1 = pts[j][i + 1][0]
                hline(y, x, x1)
                if rng.random() < 0.6:
                    cx = (x + x1) // 2
                    cls = COMPONENT_CLASSES[int(rng.integers(len(COMPONENT_CLASSES)))]
                    boxes.append(_box(cls, cx - long_side // 2, y - short_side // 2,
                                      cx + long_side // 2, y + short_side // 2))
            # vertical edge downward
            if j + 1 < G and rng.random() < 0.8:
                y1 = pts[j + 1][i][1]
                vline(x, y, y1)
                if rng.random() < 0.6:
                    cy = (y + y1) // 2
                    cls = COMPONENT_CLASSES[int(rng.integers(len(COMPONENT_CLASSES)))]
                    boxes.append(_box(cls, x - short_side // 2, cy - long_side // 2,
                                      x + short_side // 2, cy + long_side // 2))
    # text boxes off the wires (cell centres)
    for _ in range(G):
        i = int(rng.integers(G - 1))
        j = int(rng.integers(G - 1))
        cx = pts[j][i][0] + step // 2
        cy = pts[j][i][1] + step // 2
        w = max(6, step // 6)
        boxes.append(_box("text", cx - w, cy - w // 2, cx + w, cy + w // 2, 0.8))
        mask[cy - w // 4: cy + w // 4, cx - w // 2: cx + w // 2] = 255  # glyph pixels the box masks away
    # junction boxes on lattice points
    for _ in range(max(1, G // 3)):
        i = int(rng.integers(G))
        j = int(rng.integers(G))
        x, y = pts[j][i]
        r = max(4, th * 2)
        boxes.append(_box("junction", x - r, y - r, x + r, y + r, 0.7))
    if dense:
        n = int(0.002 * S * S)
        ys = rng.integers(0, S, n)
        xs = rng.integers(0, S, n)
        mask[ys, xs] = 255
    rgb = None
    if render_rgb:
        rgb = np.full((S, S, 3), 255, np.uint8)
        rgb[mask > 0] = 0
        for k, b in enumerate(boxes):
            if b["class"] in ("text", "junction"):
                continue
            x0, y0, x1, y1 = b["xmin"], b["ymin"], b["xmax"], b["ymax"]
            shade = 40 + 30 * (COMPONENT_CLASSES.index(b["class"]))
            rgb[y0:y1, x0:x1] = (shade, 255 - shade, (shade * 3) % 256)
            rgb[y0:y1:3, x0:x1] = 0  # glyph stripes
    return mask, boxes, rgb


def make_batch(seeds, size: int = 1024, **kw):
    """Stack masks for a list of seeds; boxes stay a list of lists."""
    masks, boxes = [], []
    for s in seeds:
        m, b, _ = make_schematic(int(s), size, **kw)
        masks.append(m)
        boxes.append(b)
    return np.stack(masks), boxes


def random_blob_mask(seed: int, h: int, w: int, p: float = 0.5, smooth: int = 0) -> np.ndarray:
    """Unstructured random mask for edge-case parity tests (nested rings, spurs, single pixels)."""
    rng = np.random.default_rng(seed)
    m = (rng.random((h, w)) < p).astype(np.uint8) * 255
    if smooth:
        k = smooth
        acc = np.zeros((h, w), np.float32)
        pad = np.pad(m.astype(np.float32), k, mode="edge")
        for dy in range(2 * k + 1):
            for dx in range(2 * k + 1):
                acc += pad[dy:dy + h, dx:dx + w]
        m = (acc / ((2 * k + 1) ** 2) > 127).astype(np.uint8) * 255
    return m
