"""GPU parity of the preliminary terminal reclassification (cv_terminals_analyze through the Python drop-in) against
the fixtures the unmodified reference produced (tests/golden/terminal_golden.npz, reclassify_terminals_based_on_connectivity
circuit_analyzer.py:2217) and against the CPU oracle (oracle/terminal_oracle.py).  Bit-exact: adaptive-threshold wire mask,
contour vertex arrays, per-terminal distinct-contour counts, final classes and bookkeeping keys."""
import copy

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def analyzer():
    from circuitvision_b200.circuit_analyzer import CircuitAnalyzer
    from oracle.gen_golden import CLASS_NAMES
    return CircuitAnalyzer(use_sam2=False, debug=True, device=0, class_names=CLASS_NAMES)


def test_golden_fixtures_from_unmodified_reference(analyzer, terminal_golden, terminal_cases):
    for name, (rgb, boxes) in terminal_cases.items():
        g = terminal_golden[name]
        out = copy.deepcopy(boxes)
        assert analyzer.reclassify_terminals_based_on_connectivity(rgb, out) is None
        assert [b["class"] for b in out] == g["classes"], name
        assert [bool(b.get("was_reclassified_from_terminal", False)) for b in out] == g["reclassified"], name
        assert [b.get("_yolo_class_id_temp") for b in out] == g["yolo_ids"], name
        assert [b.get("original_yolo_class_if_reclassified") for b in out] == g["orig"], name
        r = analyzer._ta().analyze(rgb[None], [boxes])
        assert r.n_contours(0) == g["n_contours"], name
        if g["n_contours"]:
            assert [int(c) for c in r.counts(0) if c >= 0] == g["terminal_counts"], name


def test_oracle_parity_masks_contours_counts(analyzer, terminal_cases):
    from oracle import terminal_oracle
    for name, (rgb, boxes) in terminal_cases.items():
        counts, mask, contours = terminal_oracle.terminal_contact_counts(rgb, boxes)
        r = analyzer._ta().analyze(rgb[None], [boxes])
        assert np.array_equal(r.wire_mask[0].cpu().numpy(), mask), name
        got = r.page_contours(0)
        assert len(got) == len(contours), name
        for a, c in zip(got, contours):
            assert np.array_equal(a, c["contour"]), name
        assert [int(c) for c in r.counts(0)] == counts, name


def test_batch_and_numpy_slice_semantics(analyzer):
    """Two pages in one call; a box whose upper bounds are negative wraps like the reference's NumPy slice (:2248)."""
    from circuitvision_b200 import synth
    from oracle import terminal_oracle
    pages, lists = [], []
    for seed in (31, 32):
        m, b, rgb = synth.make_schematic(seed, 1024, render_rgb=True)
        ys, xs = np.nonzero(m)
        extra = [{"class": "terminal", "xmin": int(xs[k]) - 9, "ymin": int(ys[k]) - 7, "xmax": int(xs[k]) + 11,
                  "ymax": int(ys[k]) + 8, "persistent_uid": f"t{k}"} for k in (5, len(xs) // 2, len(xs) - 7)]
        extra.append({"class": "text", "xmin": 300, "ymin": -40, "xmax": 700, "ymax": -30, "persistent_uid": "wrap_y"})
        extra.append({"class": "resistor", "xmin": -50, "ymin": 100, "xmax": -20, "ymax": 400, "persistent_uid": "wrap_x"})
        pages.append(rgb)
        lists.append(b + extra)
    want = [copy.deepcopy(l) for l in lists]
    ref_counts = [terminal_oracle.reclassify_terminals(p, l, analyzer.class_names) for p, l in zip(pages, want)]
    got = [copy.deepcopy(l) for l in lists]
    r = analyzer.reclassify_terminals_batch(np.stack(pages), got)
    for b in range(2):
        assert [int(c) for c in r.counts(b)] == ref_counts[b]
        assert got[b] == want[b]
        assert np.array_equal(r.wire_mask[b].cpu().numpy(), terminal_oracle.prelim_wire_mask(pages[b], lists[b]))


def test_tall_page_and_capacity_growth():
    """3000-row page (row counters beyond the 600-row node path) with noisy paper: thousands of external candidates."""
    from circuitvision_b200 import terminals
    from oracle import terminal_oracle
    rng = np.random.default_rng(9)
    page = np.full((3000, 700, 3), 235, np.uint8)
    page[::97, :] = 20
    page[:, 350:354] = 20
    page[rng.random((3000, 700)) < 0.01] = 0
    boxes = [{"class": "terminal", "xmin": 340, "ymin": 1000, "xmax": 364, "ymax": 1030, "persistent_uid": "t"},
             {"class": "resistor", "xmin": 330, "ymin": 2000, "xmax": 374, "ymax": 2100, "persistent_uid": "r"}]
    ta = terminals.TerminalAnalyzer(0, caps=dict(max_external=256, max_contours=8, max_points=512))
    r = ta.analyze(page[None], [boxes])
    counts, mask, contours = terminal_oracle.terminal_contact_counts(page, boxes)
    assert np.array_equal(r.wire_mask[0].cpu().numpy(), mask)
    assert r.n_contours(0) == len(contours)
    assert [int(c) for c in r.counts(0)] == counts
    for a, c in zip(r.page_contours(0), contours):
        assert np.array_equal(a, c["contour"])


def test_rejects_bad_input(analyzer):
    from circuitvision_b200._lib import CvError
    with pytest.raises(CvError):
        analyzer.reclassify_terminals_based_on_connectivity(np.zeros((10, 10), np.uint8), [])
    with pytest.raises(CvError):
        analyzer.reclassify_terminals_based_on_connectivity(
            np.zeros((64, 64, 3), np.uint8), [{"class": "terminal", "xmin": 1.5, "ymin": 2, "xmax": 9, "ymax": 9}])
