"""TEST INFRASTRUCTURE ONLY.  Generates tests/golden/sam2_golden.npz by EXECUTING the reference's own SAM 2.1 code
(oracle/ref_sam2.py: unmodified `SAM2Transforms`, `MultiKernelRefinement`, `SAM2ImageWrapper.forward` from
/root/reference/src/sam2_infer.py and `CircuitAnalyzer.segment_with_sam2` from src/circuit_analyzer.py:321-386) on the
deterministic cases of `sam2_cases()`; the third-party `sam2` package underneath is the restatement of
oracle/sam2_oracle.py (the only part that cannot be executed offline).  Build container only.

    python -m oracle.gen_sam2_golden

Inputs regenerate anywhere from seeds (synth.make_schematic; weights = sam2_oracle.build_oracle(variant, seed) under
torch's CPU generator), so the fixture stores outputs only:
  low      [256,256] f32   low_res_masks of the wrapper (:252-260)
  iou      f32             iou_predictions
  high_sub [128,128] f32   high_res_masks[::8, ::8] after x4 bilinear + refinement (:263-272)
  x_sub    [3,64,64] f32   SAM2Transforms output [::16, ::16] (what the wrapper was fed), x_sum its float64 sum
  mask     packed bits     (final logits > 0) at the ORIGINAL image size (:354-356)
  bbox     4 ints          extent box (:364-370), -1s when None
"""
from __future__ import annotations

import json
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from circuitvision_b200 import synth  # noqa: E402

GOLDEN = os.path.join(ROOT, "tests", "golden", "sam2_golden.npz")


def case_image(seed: int, hw):
    """The (H,W,3) uint8 array handed to segment_with_sam2 (the pipeline passes RGB into the 'bgr' argument, §D.3)."""
    _, _, rgb = synth.make_schematic(seed, 1024, render_rgb=True)
    if tuple(hw) != (1024, 1024):
        pad_h, pad_w = max(0, hw[0] - 1024), max(0, hw[1] - 1024)
        rgb = np.pad(rgb, ((0, pad_h), (0, pad_w), (0, 0)), constant_values=255)[:hw[0], :hw[1]]
    return np.ascontiguousarray(rgb)


def sam2_cases():
    """name -> (variant, weight seed, image seed, (H, W)).  base_plus s5..s8 are the four probes of the batch-256 case."""
    c = {}
    for s in (5, 6, 7):
        c[f"tiny_s{s}"] = ("tiny", 0, s, (1024, 1024))
    c["tiny_s21_720x1280"] = ("tiny", 0, 21, (720, 1280))
    c["tiny_s22_1300x900"] = ("tiny", 0, 22, (1300, 900))
    c["small_s5"] = ("small", 0, 5, (1024, 1024))
    for s in (5, 6, 7, 8):
        c[f"base_plus_s{s}"] = ("base_plus", 0, s, (1024, 1024))
    c["base_plus_s9_900x1100"] = ("base_plus", 0, 9, (900, 1100))
    c["large_s5"] = ("large", 0, 5, (1024, 1024))
    c["large_s6_720x1280"] = ("large", 0, 6, (720, 1280))
    return c


def tail_cases():
    """Stand-alone inputs for the two reference classes that need no sam2 at all."""
    g = torch.Generator().manual_seed(3)
    x = torch.randn(2, 1, 1024, 1024, generator=g)
    return x


def main():
    from oracle import ref_sam2
    assert ref_sam2.available(), "needs /root/reference"
    out, meta = {}, {}
    wrappers = {}
    for name, (variant, wseed, iseed, hw) in sam2_cases().items():
        t0 = time.time()
        if (variant, wseed) not in wrappers:
            wrappers.clear()  # one large model in memory at a time
            wrappers[(variant, wseed)] = ref_sam2.build_reference_wrapper(variant, wseed)[0]
        w = wrappers[(variant, wseed)]
        img = case_image(iseed, hw)
        fed = {}
        orig = w.forward

        def tap(images, *a, _orig=orig, **k):
            fed["x"] = images.detach().clone()
            return _orig(images, *a, **k)

        w.forward = tap
        try:
            mask, colored, bbox, taps = ref_sam2.reference_segment(w, img)
        finally:
            w.forward = orig
        assert mask is not None and mask.shape == tuple(hw)
        assert np.array_equal(colored[:, :, 1], mask) and not colored[:, :, 0].any() and not colored[:, :, 2].any()
        out[name + "/low"] = taps["low"][0, 0].numpy().astype(np.float32)
        out[name + "/iou"] = taps["iou"].numpy().astype(np.float32).reshape(1)
        out[name + "/high_sub"] = taps["high"][0, 0, ::8, ::8].numpy().astype(np.float32)
        out[name + "/x_sub"] = fed["x"][0, :, ::16, ::16].numpy().astype(np.float32)
        out[name + "/mask"] = np.packbits(mask > 0)
        out[name + "/bbox"] = np.array(bbox if bbox is not None else (-1, -1, -1, -1), np.int32)
        meta[name] = {"variant": variant, "weight_seed": wseed, "image_seed": iseed, "hw": list(hw),
                      "x_sum": float(fed["x"].double().sum()), "fg": float((mask > 0).mean()),
                      "low_std": float(taps["low"].std()), "high_std": float(taps["high"].std())}
        print(f"{name}: fg {meta[name]['fg']:.3f} bbox {bbox} iou {float(taps['iou']):.4f} ({time.time() - t0:.1f} s)", flush=True)
    # MultiKernelRefinement / postprocess_masks of the reference on seeded logits (weights: tiny seed-0 refinement head)
    w = ref_sam2.build_reference_wrapper("tiny", 0)[0]
    x = tail_cases()
    with torch.no_grad():
        r = w.refinement_layer(x)
    out["refine/out_sub"] = r[:, 0, ::8, ::8].numpy().astype(np.float32)
    out["refine/out_border"] = torch.cat([r[:, 0, :12].flatten(1), r[:, 0, -12:].flatten(1), r[:, 0, :, :12].flatten(1),
                                          r[:, 0, :, -12:].flatten(1)], 1).numpy().astype(np.float32)
    tr = ref_sam2.reference_transforms()
    for hw in [(493, 712), (1500, 1100)]:
        p = tr.postprocess_masks(x, hw)
        out[f"post/{hw[0]}x{hw[1]}_sub"] = p[:, 0, ::4, ::4].numpy().astype(np.float32)
    out["meta_json"] = np.frombuffer(json.dumps(meta).encode(), dtype=np.uint8)
    np.savez_compressed(GOLDEN, **out)
    print("wrote", GOLDEN, os.path.getsize(GOLDEN) >> 10, "KiB")


if __name__ == "__main__":
    main()
