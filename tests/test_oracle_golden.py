"""The CPU oracle (oracle/node_oracle.py) against the fixtures produced by the unmodified reference."""
import hashlib

import numpy as np
import pytest

from oracle import node_oracle, ref_loader


def _sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def test_oracle_matches_reference_golden(golden, golden_cases):
    z, meta = golden
    assert set(meta) == set(golden_cases)
    for name, (mask, boxes) in golden_cases.items():
        g = meta[name]
        assert _sha(mask) == g["mask_sha256"], f"{name}: generator drifted"
        nodes, emptied, enhanced, resized, pts = node_oracle.get_node_connections(mask, boxes)
        assert _sha(emptied) == g["emptied_sha256"], name
        assert _sha(enhanced) == g["enhanced_sha256"], name
        assert list(enhanced.shape) == g["enhanced_shape"], name
        assert len(nodes) == len(g["nodes"]), name
        for n, gn in zip(nodes, g["nodes"]):
            assert int(n["id"]) == gn["id"]
            assert [c["persistent_uid"] for c in n["components"]] == gn["uids"], name
            assert [[c["xmin"], c["ymin"], c["xmax"], c["ymax"]] for c in n["components"]] == gn["comp_xyxy"], name
            assert np.array_equal(np.asarray(n["contour"]).reshape(-1, 2), z[f"{name}/contour{gn['id']}"]), name


@pytest.mark.skipif(not ref_loader.available(), reason="reference tree only exists in the build container")
def test_oracle_matches_live_reference_extra_seeds():
    from circuitvision_b200 import synth
    for seed in (21, 22):
        mask, boxes, _ = synth.make_schematic(seed, 1024)
        rn, remp, renh, *_rest, text = ref_loader.reference_node_analysis(mask, boxes)
        nodes, emptied, enhanced, _, _ = node_oracle.get_node_connections(mask, boxes)
        assert np.array_equal(remp, emptied) and np.array_equal(renh, enhanced)
        a, b = node_oracle.node_signature(rn), node_oracle.node_signature(nodes)
        assert len(a) == len(b)
        for (i1, u1, c1), (i2, u2, c2) in zip(a, b):
            assert i1 == i2 and u1 == u2 and np.array_equal(c1, c2)


def test_ccl_oracle_canonical_labels():
    m = np.zeros((6, 8), np.uint8)
    m[0, 0] = 255
    m[1, 1] = 255          # 8-connected to (0,0)
    m[4, 5:8] = 9
    lab = node_oracle.ccl_labels_min_index(m, 8)
    assert lab[0, 0] == 1 and lab[1, 1] == 1
    assert (lab[4, 5:8] == 4 * 8 + 5 + 1).all()
    assert lab[2, 2] == 0
    lab4 = node_oracle.ccl_labels_min_index(m, 4)
    assert lab4[1, 1] == 1 * 8 + 1 + 1


def test_terminal_oracle_matches_reference_golden(terminal_golden, terminal_cases):
    """oracle/terminal_oracle.py against the fixtures of the unmodified reference reclassify_terminals_based_on_connectivity
    (circuit_analyzer.py:2217): adaptive-threshold mask, contour count, per-terminal distinct-contour counts, final
    classes and the bookkeeping keys."""
    import copy
    import cv2
    from oracle import terminal_oracle
    from oracle.gen_golden import CLASS_NAMES
    assert set(terminal_golden) == set(terminal_cases)
    for name, (rgb, boxes) in terminal_cases.items():
        g = terminal_golden[name]
        assert _sha(rgb) == g["rgb_sha256"], f"{name}: generator drifted"
        assert _sha(terminal_oracle.segment_circuit_from_rgb(rgb)) == g["mask_sha256"], name
        counts, mask, contours = terminal_oracle.terminal_contact_counts(rgb, boxes)
        assert len(contours) == g["n_contours"], name
        if g["n_contours"]:  # the reference's debug build returns before the per-terminal loop when nothing was found
            assert [c for c in counts if c >= 0] == g["terminal_counts"], name
        out = copy.deepcopy(boxes)
        terminal_oracle.reclassify_terminals(rgb, out, CLASS_NAMES)
        assert [b["class"] for b in out] == g["classes"], name
        assert [bool(b.get("was_reclassified_from_terminal", False)) for b in out] == g["reclassified"], name
        assert [b.get("_yolo_class_id_temp") for b in out] == g["yolo_ids"], name
        assert [b.get("original_yolo_class_if_reclassified") for b in out] == g["orig"], name
