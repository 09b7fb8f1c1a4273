import numpy as np
CLASSES = ["resistor", "capacitor.unpolarized", "inductor", "diode", "voltage.dc", "current.dc", "gnd", "terminal",
           "junction", "junction", "text", "text", "text", "crossover", "vss", "explanatory", "circuit", "transistor.bjt"]


def random_page(seed):
    """Deterministic page + YOLO-style boxes: a few clusters of components with labels, stray text, junction-only pages,
    empty pages, pages covered by one cluster."""
    rng = np.random.default_rng(seed)
    H, W = int(rng.integers(200, 1400)), int(rng.integers(200, 1600))
    mode = seed % 7
    boxes = []

    def add(cls, cx, cy, w, h):
        x0, y0 = int(cx - w // 2), int(cy - h // 2)
        boxes.append({"class": cls, "confidence": 0.5, "xmin": x0, "ymin": y0, "xmax": x0 + int(w), "ymax": y0 + int(h),
                      "persistent_uid": f"{cls}_{x0}_{y0}_{x0 + int(w)}_{y0 + int(h)}"})

    n_clusters = int(rng.integers(1, 4)) if mode != 3 else 0
    for _ in range(n_clusters):
        ccx, ccy = rng.integers(0, W), rng.integers(0, H)
        spread = int(rng.integers(40, 400))
        for _ in range(int(rng.integers(1, 12))):
            cls = CLASSES[int(rng.integers(len(CLASSES)))]
            if mode == 4:
                cls = "junction"
            add(cls, ccx + rng.integers(-spread, spread + 1), ccy + rng.integers(-spread, spread + 1),
                rng.integers(6, 120), rng.integers(6, 120))
    for _ in range(int(rng.integers(0, 6))):
        add("text", rng.integers(-50, W + 50), rng.integers(-50, H + 50), rng.integers(10, 200), rng.integers(8, 60))
    if mode == 5:  # one box that covers the page
        add("resistor", W // 2, H // 2, W - 4, H - 4)
    if mode == 6 and boxes:  # float coordinates (boxes before round())
        for b in boxes:
            b["xmin"] += 0.25
            b["ymax"] -= 0.5
    img = np.zeros((H, W, 3), np.uint8)
    img[::7, ::5] = (seed % 251)
    return img, boxes
