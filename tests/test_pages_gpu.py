"""the batched page path (PagePipeline: raw uint8 pages up, device crop + antialiased resize + normalise,
SAM 2.1 on the batch, logits back at crop resolution, node analysis per crop) against the per-page drop-in sequence of
analysis_pipeline.py:177-246 (crop_image_and_adjust_bboxes -> segment_with_sam2 -> get_node_connections) and against the
oracle's SAM2Transforms arithmetic on the very crop the reference would slice."""
import copy

import numpy as np
import pytest
import torch

from circuitvision_b200 import synth

pytestmark = pytest.mark.gpu


def _iou(a, b):
    inter, union = int((a & b).sum()), int((a | b).sum())
    return inter / union if union else 1.0


def _page(seed, hw, at):
    m, boxes, rgb = synth.make_schematic(seed, 1024, render_rgb=True)
    page = np.full(hw + (3,), 255, np.uint8)
    y0, x0 = at
    h, w = min(1024, hw[0] - y0), min(1024, hw[1] - x0)
    page[y0:y0 + h, x0:x0 + w] = rgb[:h, :w]
    moved = [dict(b, xmin=b["xmin"] + x0, xmax=b["xmax"] + x0, ymin=b["ymin"] + y0, ymax=b["ymax"] + y0) for b in boxes
             if b["xmax"] < w and b["ymax"] < h]
    return page, moved


def test_page_pipeline_matches_per_page_drop_in_calls():
    from oracle import node_oracle, sam2_oracle
    from circuitvision_b200 import sam2_infer
    from circuitvision_b200.circuit_analyzer import CircuitAnalyzer
    from circuitvision_b200.pipeline import PagePipeline
    ref = sam2_oracle.build_oracle("tiny", seed=0)
    model = sam2_infer.get_modified_sam2("tiny", None, device="cuda:0", use_refinement_layer=True)
    model.load_state_dict(ref.state_dict())
    model.set_max_batch(3)
    A = CircuitAnalyzer(sam2_model=model, sam2_transforms=sam2_infer.SAM2Transforms(1024, 0.0), debug=True, device=0)
    pages, lists = zip(*[_page(61, (1400, 1700), (150, 300)), _page(62, (1100, 1300), (40, 60)), _page(63, (2000, 1500), (700, 200))])
    pipe = PagePipeline(model, padding=80)
    res = pipe.run(list(pages), [copy.deepcopy(l) for l in lists])
    assert pipe.h2d_bytes <= sum(p.size for p in pages) + 3 * 32  # raw uint8 pages only
    for b, (page, boxes) in enumerate(zip(pages, lists)):
        crop, moved, info = A.crop_image_and_adjust_bboxes(page, copy.deepcopy(boxes), padding=80)
        assert info["crop_applied"] and res[b]["crop_info"]["final_crop_window_abs"] == info["final_crop_window_abs"]
        assert res[b]["boxes"] == moved
        # SAM2Transforms on the crop (the :343 swap included), fp32 round-off
        want_x = sam2_oracle.preprocess_rgb(np.ascontiguousarray(crop[:, :, ::-1]))
        assert (pipe.last_input[b].cpu() - want_x).abs().max().item() <= 2e-5
        mask1, _, ext1 = A.segment_with_sam2(np.ascontiguousarray(crop))
        assert res[b]["mask"].shape == crop.shape[:2]
        assert _iou(res[b]["mask"] > 0, mask1 > 0) >= 0.998  # batched == single up to attention-tile rounding
        # node analysis of the pipeline's own mask: bit-exact against the CPU oracle on that mask
        rn, remp, renh, _, _ = node_oracle.get_node_connections(res[b]["mask"], moved)
        assert np.array_equal(res[b]["emptied"], remp) and np.array_equal(res[b]["enhanced"], renh)
        a, c = node_oracle.node_signature(res[b]["nodes"]), node_oracle.node_signature(rn)
        assert len(a) == len(c) and all(x[0] == y[0] and x[1] == y[1] and np.array_equal(x[2], y[2]) for x, y in zip(a, c))
