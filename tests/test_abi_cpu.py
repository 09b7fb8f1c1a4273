"""The C-ABI library loads without a GPU and exports every symbol include/cv_b200.h declares."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
    txt = open(os.path.join(ROOT, "include", "cv_b200.h")).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    return sorted(set(re.findall(r"\b(cv_[a-z0-9_]+)\s*\(", txt)))


@pytest.fixture(scope="module")
def lib():
    from circuitvision_b200 import build, _lib
    build.build()
    return _lib.load()


def test_header_symbols_exported(lib):
    syms = _declared_symbols()
    assert "cv_nodes_analyze" in syms and "cv_ccl_label" in syms
    for s in syms:
        assert hasattr(lib, s), f"{s} declared in include/cv_b200.h but not exported"


def test_host_only_entry_points(lib):
    assert lib.cv_version().decode().startswith("circuitvision_b200")
    assert lib.cv_nodes_resized_width(1024, 1024) == 600
    assert lib.cv_nodes_resized_width(493, 712) == int(600 * (712 / 493))
    assert lib.cv_nodes_resized_width(720, 1280) == int(600 * (1280 / 720))
    assert lib.cv_nodes_workspace_bytes(2, 1024, 1024, None) > 2 * 600 * 600 * 7


def test_invalid_arguments_fail_loudly(lib):
    rc = lib.cv_nodes_analyze(None, 0, 0, 0, None, None, 0, None, None, None, None, None, None, None, None, None, 0, None)
    assert rc != 0 and b"cv_nodes_analyze" in lib.cv_last_error()
    rc = lib.cv_ccl_label(None, 1, 4, 4, 8, None, None, None, 0, None)
    assert rc != 0


def test_product_path_refuses_to_run_without_gpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from circuitvision_b200 import CvError
    from circuitvision_b200.nodes import NodeAnalyzer
    with pytest.raises(CvError):
        NodeAnalyzer(0)


def test_box_packing_matches_reference_rules():
    from circuitvision_b200.nodes import pack_boxes
    boxes = [{"class": "resistor", "xmin": 10.7, "ymin": -3.2, "xmax": 50.9, "ymax": 20.0, "persistent_uid": "a"},
             {"class": "junction", "xmin": 1, "ymin": 2, "xmax": 3, "ymax": 4, "persistent_uid": "b"},
             {"class": "voltage.dc", "xmin": 1, "ymin": 2, "xmax": 3, "ymax": 4, "persistent_uid": "a"},
             {"class": "diode", "xmin": 5, "ymin": 6, "xmax": 7, "ymax": 8}]
    rec, offs, rb, mx = pack_boxes([boxes, []], 720, 1000)
    assert list(offs) == [0, 4, 4] and mx == 4
    nw = int(600 * (1000 / 720))
    assert rec["xmin"][0] == 10 and rec["ymin"][0] == -3          # int() truncates toward zero
    assert rec["rxmin"][0] == int(10.7 * (nw / 1000)) and rec["rymin"][0] == int(-3.2 * (600 / 720))
    assert list(rec["flags"]) == [3, 0, 7, 3]
    assert list(rec["thresh"]) == [6, 6, 20, 8]
    assert list(rec["uid_group"]) == [0, 1, 0, 3]
    assert rb[0][0]["xmin"] == rec["rxmin"][0]
