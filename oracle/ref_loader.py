"""TEST INFRASTRUCTURE ONLY.  Loads the *unmodified* reference `src/circuit_analyzer.py`
from /root/reference with its absent third-party imports mock-stubbed.

Only works in the build container (the GPU box has no /root/reference); it is used by
`oracle/gen_golden.py` to produce the committed fixtures under tests/golden/ and by
CPU tests that are skipped when the reference tree is missing.
"""
from __future__ import annotations

import contextlib
import importlib
import io
import os
import sys
from unittest.mock import MagicMock

REF_ROOT = "/root/reference"
_STUBS = [
    "ultralytics", "matplotlib", "matplotlib.pyplot", "sam2", "sam2.build_sam",
    "sam2.sam2_image_predictor", "sam2.modeling", "sam2.modeling.sam2_base", "sam2.utils",
    "sam2.utils.misc", "peft", "google", "google.genai", "google.genai.types", "streamlit",
    "PySpice", "PySpice.Spice", "PySpice.Spice.Netlist", "PySpice.Unit", "openai", "dotenv",
    "groq",
]
_analyzer = None


def available() -> bool:
    return os.path.isfile(os.path.join(REF_ROOT, "src", "circuit_analyzer.py"))


def load_reference_analyzer():
    """Return a `CircuitAnalyzer(use_sam2=False, debug=True)` built from the reference's own code."""
    global _analyzer
    if _analyzer is not None:
        return _analyzer
    if not available():
        raise RuntimeError("reference tree not present (only available in the build container)")
    for name in _STUBS:
        try:
            importlib.import_module(name)
        except Exception:
            sys.modules[name] = MagicMock()
    if REF_ROOT not in sys.path:
        sys.path.insert(0, REF_ROOT)
    os.environ.pop("GEMINI_API_KEY", None)
    cwd = os.getcwd()
    os.chdir(REF_ROOT)  # classes.json is read cwd-relative (utils.py:102)
    try:
        with contextlib.redirect_stdout(io.StringIO()):
            ca = importlib.import_module("src.circuit_analyzer")
            _analyzer = ca.CircuitAnalyzer(use_sam2=False, debug=True)
            _analyzer.show_image = lambda *a, **k: None
    finally:
        os.chdir(cwd)
    return _analyzer


def reference_node_analysis(mask, boxes):
    """Run the reference get_node_connections + netlist consumer; returns the parity artifacts."""
    import copy
    A = load_reference_analyzer()
    with contextlib.redirect_stdout(io.StringIO()):
        nodes, emptied, enhanced, cimg, fviz, cpts = A.get_node_connections(
            None, None if mask is None else mask.copy(), copy.deepcopy(boxes))
        try:
            text = "\n".join(A.stringify_line(l) for l in A.generate_netlist_from_nodes(copy.deepcopy(nodes)))
        except Exception as e:  # consumer failure is itself a parity artifact
            text = f"<netlist error: {type(e).__name__}>"
    return nodes, emptied, enhanced, cimg, fviz, cpts, text


def reference_reclassify(image_rgb, boxes, class_names=None):
    """Run the reference reclassify_terminals_based_on_connectivity (circuit_analyzer.py:2217) in place on a deep copy
    of `boxes`; returns the modified list.  `class_names` stands in for `self.yolo.model.names` (ultralytics is absent)."""
    import copy
    A = load_reference_analyzer()
    out = copy.deepcopy(boxes)
    A.yolo = MagicMock()
    A.yolo.model.names = dict(class_names or {})
    log = io.StringIO()
    with contextlib.redirect_stdout(log):
        A.reclassify_terminals_based_on_connectivity(image_rgb.copy(), out)
    # the reference prints (debug=True) the contour count and, per terminal in list order, its distinct-contour count
    import re
    text = log.getvalue()
    found = re.search(r"Prelim Reclass: Found (\d+) contours", text)
    counts = [int(x) for x in re.findall(r"connected to (\d+) distinct contours", text)]
    return out, (int(found.group(1)) if found else None), counts
