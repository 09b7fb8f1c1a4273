"""TEST INFRASTRUCTURE ONLY.  Runs the reference's OWN SAM 2.1 wrapper code, unmodified, in the build container:

    /root/reference/src/sam2_infer.py        SAM2Transforms :29-128, MultiKernelRefinement :130-189,
                                              SAM2ImageWrapper.__init__/forward :191-275
    /root/reference/src/circuit_analyzer.py  CircuitAnalyzer.segment_with_sam2 :321-386

The only thing that is NOT the reference's code is the third-party `sam2` package it calls into (un-vendored,
un-pinned git dependency, requirements.txt:12 — not installable offline).  `ShimSAM2Base` stands in for the object
`build_sam2()` returns: it exposes exactly the attributes the wrapper touches (sam2_infer.py:226-265) —
`image_encoder(images) -> {"backbone_fpn": [...]}`, `sam_mask_decoder(.conv_s0, .conv_s1, __call__ with upstream's
keyword arguments)`, `sam_prompt_encoder.get_dense_pe()`, `_prepare_backbone_features(out)`, `image_size` — over the
restated modules of oracle/sam2_oracle.py, with the same parameter tree, so `oracle.state_dict()` loads 1:1 into the
reference wrapper.  `peft` is not needed: LoRA enters through merged weights (random init has lora_B = 0).

Works only where /root/reference exists; used by oracle/gen_sam2_golden.py (fixtures under tests/golden/) and by CPU
tests that skip when the tree is missing.
"""
from __future__ import annotations

import contextlib
import importlib
import io
import sys
from unittest.mock import MagicMock

import numpy as np
import torch

from . import ref_loader, sam2_oracle

_mod = None


def available() -> bool:
    return ref_loader.available()


def reference_module():
    """The reference's src/sam2_infer.py imported unmodified (sam2 / peft mocked: it only imports names from them at
    module level and uses them inside get_modified_sam2, which the fixtures do not call)."""
    global _mod
    if _mod is not None:
        return _mod
    if not available():
        raise RuntimeError("reference tree not present (only available in the build container)")
    for name in ("sam2", "sam2.build_sam", "sam2.sam2_image_predictor", "sam2.modeling", "sam2.modeling.sam2_base",
                 "sam2.utils", "sam2.utils.misc", "peft"):
        try:
            importlib.import_module(name)
        except Exception:
            sys.modules[name] = MagicMock()
    if ref_loader.REF_ROOT not in sys.path:
        sys.path.insert(0, ref_loader.REF_ROOT)
    with contextlib.redirect_stdout(io.StringIO()):
        _mod = importlib.import_module("src.sam2_infer")
    return _mod


class _ShimImageEncoder(sam2_oracle.ImageEncoder):
    """upstream ImageEncoder.forward returns a dict; the wrapper reads and rewrites out["backbone_fpn"] (:226-232)."""

    def forward(self, x):
        fpn = list(super().forward(x))
        return {"vision_features": fpn[-1], "vision_pos_enc": [None] * len(fpn), "backbone_fpn": fpn}


class _ShimMaskDecoder(sam2_oracle.MaskDecoder):
    """upstream MaskDecoder.forward keyword interface (sam2_infer.py:252-260) -> 4 outputs."""

    def forward(self, image_embeddings, image_pe, sparse_prompt_embeddings, dense_prompt_embeddings, multimask_output,
                repeat_image, high_res_features=None):
        assert multimask_output is False and repeat_image is True and high_res_features is not None
        masks, iou, aux = super().forward(image_embeddings, image_pe, sparse_prompt_embeddings, dense_prompt_embeddings,
                                          high_res_features)
        return masks, iou, None, aux["obj"]


class ShimSAM2Base(torch.nn.Module):
    image_size = sam2_oracle.IMAGE_SIZE

    def __init__(self, variant):
        super().__init__()
        self.image_encoder = _ShimImageEncoder(variant)
        self.sam_prompt_encoder = sam2_oracle.PromptEncoder()
        self.sam_mask_decoder = _ShimMaskDecoder()

    def _prepare_backbone_features(self, backbone_out):
        """upstream SAM2Base._prepare_backbone_features: the last 3 FPN levels flattened to (HW, B, C)."""
        maps = backbone_out["backbone_fpn"][-3:]
        pos = backbone_out["vision_pos_enc"][-3:]
        feat_sizes = [(x.shape[-2], x.shape[-1]) for x in maps]
        vision_feats = [x.flatten(2).permute(2, 0, 1) for x in maps]
        return backbone_out, vision_feats, pos, feat_sizes


def build_reference_wrapper(variant="tiny", seed=0):
    """The reference's SAM2ImageWrapper (its own class, its own forward) around the shim, carrying exactly the weights of
    `sam2_oracle.build_oracle(variant, seed)` (strict load: the two parameter trees have identical names)."""
    m = reference_module()
    oracle = sam2_oracle.build_oracle(variant, seed=seed)
    torch.manual_seed(12345)  # the wrapper's own randn parameters are overwritten by the strict load below
    w = m.SAM2ImageWrapper(ShimSAM2Base(variant), embedding_r=4, use_refinement=True, refinement_kernel_sizes=[3, 5, 7, 11])
    res = w.load_state_dict(oracle.state_dict(), strict=True)
    assert not res.missing_keys and not res.unexpected_keys
    return w.eval(), oracle


def reference_transforms():
    """SAM2Transforms exactly as circuit_analyzer.py:245-250 constructs it."""
    return reference_module().SAM2Transforms(resolution=1024, mask_threshold=0, max_hole_area=0, max_sprinkle_area=0)


def reference_segment(wrapper, image_np_bgr):
    """The reference's CircuitAnalyzer.segment_with_sam2 (:321-386), unmodified, on the CPU.
    -> (mask u8 {0,255}, coloured image, extent box) + the tensors the wrapper produced on the way."""
    A = ref_loader.load_reference_analyzer()
    A.use_sam2 = True
    A.sam2_model = wrapper
    A.sam2_transforms = reference_transforms()
    A.sam2_device = torch.device("cpu")
    taps = {}
    orig_forward = wrapper.forward

    def tapped(*a, **k):
        out = orig_forward(*a, **k)
        taps["high"], taps["low"], taps["iou"] = (t.detach().clone() for t in out)
        return out

    wrapper.forward = tapped
    try:
        with contextlib.redirect_stdout(io.StringIO()):
            mask, colored, bbox = A.segment_with_sam2(np.ascontiguousarray(image_np_bgr))
    finally:
        wrapper.forward = orig_forward
        A.use_sam2 = False
        A.sam2_model = None
    return mask, colored, bbox, taps
