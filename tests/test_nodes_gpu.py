"""GPU parity of the node / connection analysis (through the C ABI) against the CPU oracle and against the
fixtures the unmodified reference produced (tests/golden/node_golden.npz).  Bit-exact: emptied mask, enhanced
image, node ids, component uid lists, contour vertex arrays."""
import hashlib

import numpy as np
import pytest

from circuitvision_b200 import synth

pytestmark = pytest.mark.gpu


def _sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


@pytest.fixture(scope="module")
def analyzer():
    from circuitvision_b200.circuit_analyzer import CircuitAnalyzer
    return CircuitAnalyzer(use_sam2=False, debug=True, device=0)


def _assert_nodes_equal(got, ref, name=""):
    from oracle.node_oracle import node_signature
    a, b = node_signature(got), node_signature(ref)
    assert len(a) == len(b), f"{name}: {len(a)} nodes vs {len(b)}"
    for (i1, u1, c1), (i2, u2, c2) in zip(a, b):
        assert i1 == i2, name
        assert u1 == u2, f"{name}: node {i1} components {u1} vs {u2}"
        assert np.array_equal(c1, c2), f"{name}: node {i1} contour differs"


def test_golden_fixtures_from_unmodified_reference(analyzer, golden, golden_cases):
    z, meta = golden
    for name, (mask, boxes) in golden_cases.items():
        g = meta[name]
        nodes, emptied, enhanced, cviz, fviz, pviz = analyzer.get_node_connections(None, mask, boxes)
        assert _sha(emptied) == g["emptied_sha256"], name
        assert list(enhanced.shape) == g["enhanced_shape"], name
        assert _sha(enhanced) == g["enhanced_sha256"], name
        assert len(nodes) == len(g["nodes"]), name
        for n, gn in zip(nodes, g["nodes"]):
            assert n["id"] == gn["id"], name
            assert [c["persistent_uid"] for c in n["components"]] == gn["uids"], name
            assert [[c["xmin"], c["ymin"], c["xmax"], c["ymax"]] for c in n["components"]] == gn["comp_xyxy"], name
            assert np.array_equal(n["contour"].reshape(-1, 2), z[f"{name}/contour{gn['id']}"]), name
            assert n["contour"].dtype == np.int32 and n["contour"].shape[1:] == (1, 2)
        assert cviz.shape == enhanced.shape + (3,) and fviz.shape == cviz.shape and pviz.shape == cviz.shape
        # the three debug drawings (circuit_analyzer.py:414-458, :1585-1603): same cv2 calls on device-produced contours /
        # connection points / resized image => byte-identical to what the unmodified reference returned
        assert _sha(cviz) == g["contour_viz_sha256"], name
        assert _sha(fviz) == g["final_viz_sha256"], name
        assert _sha(pviz) == g["points_viz_sha256"], name
        # netlist connectivity: the drop-in's own generate_netlist_from_nodes / stringify_line on the device-produced
        # node table against the text the unmodified reference printed for this case
        text = "\n".join(analyzer.stringify_line(l) for l in analyzer.generate_netlist_from_nodes(nodes))
        assert text == g["netlist"], name


def test_oracle_parity_fresh_seeds(analyzer):
    from oracle import node_oracle
    for seed in range(100, 112):
        mask, boxes, _ = synth.make_schematic(seed, 1024)
        ref_nodes, ref_emp, ref_enh, ref_res, ref_pts = node_oracle.get_node_connections(mask, boxes)
        nodes, emptied, enhanced, *_ = analyzer.get_node_connections(None, mask, boxes)
        assert np.array_equal(emptied, ref_emp) and np.array_equal(enhanced, ref_enh), seed
        _assert_nodes_equal(nodes, ref_nodes, f"seed {seed}")


def test_oracle_parity_random_blobs_and_shapes(analyzer):
    """Hand-drawn-like masks: nested rings, spurs, single pixels, non-square crops, float boxes."""
    from oracle import node_oracle
    rng = np.random.default_rng(7)
    classes = ["resistor", "voltage.dc", "diode", "text", "junction", "terminal", "gnd", "transistor.bjt"]
    for i, (h, w) in enumerate([(493, 712), (720, 1280), (600, 600), (333, 1000), (1500, 700), (64, 64)]):
        mask = synth.random_blob_mask(50 + i, h, w, p=0.5, smooth=2 + (i % 3))
        boxes = []
        for k in range(14):
            x0, y0 = float(rng.uniform(-10, w - 20)), float(rng.uniform(-10, h - 20))
            bw, bh = float(rng.uniform(5, w / 4)), float(rng.uniform(5, h / 4))
            cls = classes[int(rng.integers(len(classes)))]
            boxes.append({"class": cls, "xmin": x0, "ymin": y0, "xmax": x0 + bw, "ymax": y0 + bh,
                          "persistent_uid": f"{cls}_{k % 11}"})  # some duplicate uids
        ref_nodes, ref_emp, ref_enh, _, _ = node_oracle.get_node_connections(mask, boxes)
        nodes, emptied, enhanced, *_ = analyzer.get_node_connections(None, mask, boxes)
        assert np.array_equal(emptied, ref_emp), (h, w)
        assert np.array_equal(enhanced, ref_enh), (h, w)
        _assert_nodes_equal(nodes, ref_nodes, f"blob {h}x{w}")


def test_batch_equals_stack_of_singles(analyzer):
    from oracle import node_oracle
    seeds = list(range(200, 216))
    masks, boxes = synth.make_batch(seeds, 1024)
    r = analyzer.get_node_connections_batch(masks, boxes)
    assert r.launches > 0
    emp = r.emptied.cpu().numpy()
    enh = r.enhanced.cpu().numpy()
    for b, s in enumerate(seeds):
        ref_nodes, ref_emp, ref_enh, _, ref_pts = node_oracle.get_node_connections(masks[b], boxes[b])
        assert np.array_equal(emp[b], ref_emp) and np.array_equal(enh[b], ref_enh), s
        _assert_nodes_equal(r.nodes(b), ref_nodes, f"batch seed {s}")
        assert r.connection_points(b) == [tuple(p) for p in ref_pts]


def test_mask_none_and_input_not_mutated(analyzer):
    ctx = np.zeros((50, 70, 3), np.uint8)
    out = analyzer.get_node_connections(ctx, None, [])
    assert out[0] == [] and all(o.shape == (50, 70, 3) for o in out[1:])
    mask, boxes, _ = synth.make_schematic(3, 1024)
    m0 = mask.copy()
    b0 = [dict(b) for b in boxes]
    analyzer.get_node_connections(None, mask, boxes)
    assert np.array_equal(mask, m0) and boxes == b0


def test_full_size_4096_properties(analyzer):
    """BASELINE cfg-4 size: properties that do not need the oracle — idempotence of box masking, emptied is a
    subset of the mask, every node contour lies on foreground of the enhanced image, ids are 0..n-1."""
    import torch
    masks, boxes = synth.make_batch([900, 901], 4096)
    r = analyzer.get_node_connections_batch(masks, boxes)
    emp = r.emptied.cpu().numpy()
    assert ((emp != 0) <= (masks != 0)).all()
    r2 = analyzer.get_node_connections_batch(emp.copy(), boxes)
    assert torch.equal(r2.emptied.cpu(), torch.from_numpy(emp))
    enh = r.enhanced.cpu().numpy()
    for b in range(2):
        nodes = r.nodes(b)
        assert [n["id"] for n in nodes] == list(range(len(nodes)))
        for n in nodes:
            pts = n["contour"].reshape(-1, 2)
            assert (enh[b][pts[:, 1], pts[:, 0]] != 0).all()
    # oracle on one of them (cv2 finishes a 4096² image in tens of ms)
    from oracle import node_oracle
    ref_nodes, ref_emp, ref_enh, _, _ = node_oracle.get_node_connections(masks[0], boxes[0])
    assert np.array_equal(emp[0], ref_emp) and np.array_equal(enh[0], ref_enh)
    _assert_nodes_equal(r.nodes(0), ref_nodes, "4096")


@pytest.mark.parametrize("conn", [4, 8])
def test_native_ccl_matches_cv2(conn):
    import torch
    from circuitvision_b200.nodes import ccl_label
    from oracle.node_oracle import ccl_labels_min_index
    imgs = [synth.random_blob_mask(5, 257, 300, p=0.5, smooth=0), synth.random_blob_mask(6, 257, 300, p=0.55, smooth=2)]
    d = torch.from_numpy(np.stack(imgs)).cuda()
    lab, cnt = ccl_label(d, conn)
    lab = lab.cpu().numpy()
    for i, m in enumerate(imgs):
        ref = ccl_labels_min_index(m, conn)
        assert np.array_equal(lab[i], ref)
        assert int(cnt[i]) == len(np.unique(ref)) - 1
    m, _, _ = synth.make_schematic(77, 4096)
    d = torch.from_numpy(m[None]).cuda()
    lab, cnt = ccl_label(d, conn)
    assert np.array_equal(lab[0].cpu().numpy(), ccl_labels_min_index(m, conn))
    # random textures over many tiles: percolating blobs, thin diagonal structures, runs that cross every seam
    for seed, (hh, ww), p, smooth in [(11, (513, 1030), 0.45, 0), (12, (513, 1030), 0.55, 1), (13, (700, 900), 0.62, 0),
                                      (14, (1024, 768), 0.5, 2), (15, (333, 2049), 0.58, 0), (16, (2048, 2048), 0.593, 0)]:
        m = synth.random_blob_mask(seed, hh, ww, p=p, smooth=smooth)
        lab, cnt = ccl_label(torch.from_numpy(m[None]).cuda(), conn)
        ref = ccl_labels_min_index(m, conn)
        assert np.array_equal(lab[0].cpu().numpy(), ref), (seed, conn)
        assert int(cnt[0]) == len(np.unique(ref)) - 1, (seed, conn)
    # edge cases: all background / all foreground / single row
    for arr in (np.zeros((1, 9, 33), np.uint8), np.full((1, 9, 33), 255, np.uint8), np.full((1, 1, 70), 3, np.uint8)):
        lab, cnt = ccl_label(torch.from_numpy(arr).cuda(), conn)
        assert np.array_equal(lab[0].cpu().numpy(), ccl_labels_min_index(arr[0], conn))


def test_shared_analyzer_from_several_threads(analyzer):
    """The app shares ONE analyzer across its session threads (app.py:134; SURVEY §8(b) threading): concurrent calls must
    serialise internally and every caller must get its own, correct result."""
    import threading
    from oracle import node_oracle
    cases = [synth.make_schematic(300 + i, 1024)[:2] for i in range(6)]
    want = [node_oracle.get_node_connections(m, b) for m, b in cases]
    got, errors = [None] * len(cases), []

    def work(i):
        try:
            for _ in range(3):
                got[i] = analyzer.get_node_connections(None, cases[i][0], cases[i][1])
        except Exception as e:  # noqa: BLE001
            errors.append(e)

    threads = [threading.Thread(target=work, args=(i,)) for i in range(len(cases))]
    for t in threads:
        t.start()
    for t in threads:
        t.join()
    assert not errors, errors
    for i, (nodes, emptied, enhanced, *_rest) in enumerate(got):
        rn, remp, renh, _, _ = want[i]
        assert np.array_equal(emptied, remp) and np.array_equal(enhanced, renh), i
        _assert_nodes_equal(nodes, rn, f"thread case {i}")


def _ring_of_blobs_case():
    """One component box with 80+ separate wire blobs just outside its perimeter (a noisy hand-drawn page does this): every
    blob is a kept contour in contact with the same box — more than the 64 contacts k_contact stages per box."""
    h, w = 600, 1000
    mask = np.zeros((h, w), np.uint8)
    x0, y0, x1, y1 = 100, 100, 900, 500
    for x in range(x0, x1 - 19, 30):
        mask[y0 - 20:y0 + 4, x:x + 20] = 255  # straddles the edge: the part inside the box is emptied, the rest touches it
        mask[y1 - 4:y1 + 20, x:x + 20] = 255
    for y in range(y0, y1 - 19, 30):
        mask[y:y + 20, x0 - 20:x0 + 4] = 255
        mask[y:y + 20, x1 - 4:x1 + 20] = 255
    boxes = [{"class": "resistor", "xmin": x0, "ymin": y0, "xmax": x1, "ymax": y1, "persistent_uid": "resistor_0"},
             # a second detection of the same rectangle: nodes need two components to be kept (circuit_analyzer.py:1558)
             {"class": "capacitor.unpolarized", "xmin": x0, "ymin": y0, "xmax": x1, "ymax": y1, "persistent_uid": "cap_0"}]
    return mask, boxes


def test_box_touching_more_than_64_contours(analyzer):
    from oracle import node_oracle
    mask, boxes = _ring_of_blobs_case()
    ref_nodes, ref_emp, ref_enh, _, ref_pts = node_oracle.get_node_connections(mask, boxes)
    assert sum(any(c["persistent_uid"] == "resistor_0" for c in n["components"]) for n in ref_nodes) > 64
    nodes, emptied, enhanced, *_ = analyzer.get_node_connections(None, mask, boxes)
    assert np.array_equal(emptied, ref_emp) and np.array_equal(enhanced, ref_enh)
    _assert_nodes_equal(nodes, ref_nodes, "ring of blobs")
    r = analyzer.get_node_connections_batch(np.stack([mask, mask]), [boxes, boxes])
    for b in range(2):
        _assert_nodes_equal(r.nodes(b), ref_nodes, f"ring of blobs, batch item {b}")
        assert r.connection_points(b) == [tuple(p) for p in ref_pts]
