"""SAM 2.1 oracle pinning (CPU).

1. tests/golden/sam2_golden.npz was produced by EXECUTING the reference's own code (oracle/ref_sam2.py: the unmodified
   SAM2Transforms / MultiKernelRefinement / SAM2ImageWrapper.forward of /root/reference/src/sam2_infer.py and
   CircuitAnalyzer.segment_with_sam2 of src/circuit_analyzer.py:321-386) — only the third-party `sam2` package under it is a
   restatement.  The travelling restatement (oracle/sam2_oracle.py) must reproduce those fixtures on any box.
2. In the build container the reference classes are executed live against the restatement.
3. The restated third-party modules are cross-checked against the independent `transformers` implementation for base+ and
   large too (SURVEY.md §C.2 config recipe; tiny is covered in test_sam2_oracle_cpu.py)."""
import json
import os

import numpy as np
import pytest
import torch

from oracle import gen_sam2_golden, ref_sam2, sam2_oracle
from test_sam2_oracle_cpu import _to_hf_key

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def sam2_golden():
    z = np.load(os.path.join(ROOT, "tests", "golden", "sam2_golden.npz"))
    return z, json.loads(bytes(z["meta_json"]).decode())


def _iou(a, b):
    inter, union = (a & b).sum(), (a | b).sum()
    return inter / union if union else 1.0


def test_fixture_is_complete(sam2_golden):
    z, meta = sam2_golden
    cases = gen_sam2_golden.sam2_cases()
    assert set(meta) == set(cases)
    for name, (variant, wseed, iseed, hw) in cases.items():
        m = meta[name]
        assert (m["variant"], m["weight_seed"], m["image_seed"], tuple(m["hw"])) == (variant, wseed, iseed, tuple(hw))
        assert z[name + "/low"].shape == (256, 256) and z[name + "/high_sub"].shape == (128, 128)
        assert z[name + "/mask"].size == (hw[0] * hw[1] + 7) // 8
    assert {"tiny", "small", "base_plus", "large"} == {m["variant"] for m in meta.values()}
    assert any(tuple(m["hw"]) != (1024, 1024) for m in meta.values())


@pytest.mark.parametrize("name", ["tiny_s7", "tiny_s21_720x1280"])
def test_restatement_reproduces_reference_fixture(sam2_golden, name):
    """The travelling oracle == what the reference's own wrapper + segment_with_sam2 produced (fp32 round-off across hosts)."""
    z, meta = sam2_golden
    m = meta[name]
    oracle = sam2_oracle.build_oracle(m["variant"], seed=m["weight_seed"])
    img = gen_sam2_golden.case_image(m["image_seed"], m["hw"])
    x = sam2_oracle.preprocess_rgb(img[:, :, ::-1])[None]  # circuit_analyzer.py:343 swap
    assert np.allclose(x[0, :, ::16, ::16].numpy(), z[name + "/x_sub"], atol=1e-5)
    assert abs(float(x.double().sum()) - m["x_sum"]) <= 1e-3 * max(1.0, abs(m["x_sum"]))
    with torch.no_grad():
        high, low, iou = oracle(x)
    std = m["low_std"]
    assert np.abs(low[0, 0].numpy() - z[name + "/low"]).max() <= 2e-3 * std
    assert abs(float(iou) - float(z[name + "/iou"][0])) <= 1e-5
    assert np.abs(high[0, 0, ::8, ::8].numpy() - z[name + "/high_sub"]).max() <= 2e-3 * m["high_std"]
    mask, _, bbox = sam2_oracle.segment(oracle, img)
    want = np.unpackbits(z[name + "/mask"])[:mask.size].reshape(mask.shape).astype(bool)
    assert _iou(mask > 0, want) >= 0.9995
    assert tuple(int(v) for v in z[name + "/bbox"]) == tuple(bbox)


needs_ref = pytest.mark.skipif(not ref_sam2.available(), reason="reference tree only exists in the build container")


@needs_ref
def test_reference_wrapper_executes_and_equals_restatement():
    """SAM2ImageWrapper.forward + segment_with_sam2 of the reference, executed, vs the restatement: identical arithmetic."""
    w, oracle = ref_sam2.build_reference_wrapper("tiny", 0)
    img = gen_sam2_golden.case_image(21, (720, 1280))
    mask, colored, bbox, taps = ref_sam2.reference_segment(w, img)
    m2, _, b2 = sam2_oracle.segment(oracle, img)
    assert np.array_equal(mask, m2) and tuple(bbox) == tuple(b2)
    x = sam2_oracle.preprocess_rgb(img[:, :, ::-1])[None]
    with torch.no_grad():
        h, l, i = oracle(x)
    assert torch.allclose(l, taps["low"], atol=1e-5) and torch.allclose(h, taps["high"], atol=1e-5)
    assert torch.allclose(i, taps["iou"], atol=1e-6)


@needs_ref
def test_reference_tail_classes_equal_restatement():
    """MultiKernelRefinement (:130-189), SAM2Transforms.__call__ / forward_batch / postprocess_masks (:29-128) executed."""
    from PIL import Image
    w, oracle = ref_sam2.build_reference_wrapper("tiny", 0)
    x = gen_sam2_golden.tail_cases()[:1]
    with torch.no_grad():
        assert torch.allclose(w.refinement_layer(x), oracle.refinement_layer(x), atol=1e-6)
    tr = ref_sam2.reference_transforms()
    rng = np.random.default_rng(0)
    for hw in [(1024, 1024), (493, 712), (1500, 1100)]:
        img = rng.integers(0, 256, hw + (3,), dtype=np.uint8)
        assert torch.allclose(tr(Image.fromarray(img)), sam2_oracle.preprocess_rgb(img), atol=1e-6)
        assert torch.allclose(tr.postprocess_masks(x, hw), sam2_oracle.postprocess_masks(x, hw), atol=1e-6)
    b = tr.forward_batch([Image.fromarray(rng.integers(0, 256, (300, 400, 3), dtype=np.uint8)) for _ in range(2)])
    assert b.shape == (2, 3, 1024, 1024)


def _hf_config(variant):
    from transformers import Sam2Config
    v = sam2_oracle.VARIANTS[variant]
    e = v["embed"]
    return Sam2Config(vision_config={
        "backbone_config": {"hidden_size": e, "num_attention_heads": v["heads"], "blocks_per_stage": list(v["stages"]),
                            "embed_dim_per_stage": [e, 2 * e, 4 * e, 8 * e],
                            "num_attention_heads_per_stage": [v["heads"] * 2 ** s for s in range(4)],
                            "global_attention_blocks": list(v["global_blocks"]),
                            "window_size_per_stage": list(v["window_spec"]),
                            "window_positional_embedding_background_size": list(v["bkg"])},
        "backbone_channel_list": [8 * e, 4 * e, 2 * e, e]})


@pytest.mark.parametrize("variant", ["base_plus", "large"])
def test_restated_sam2_package_matches_transformers(variant):
    """The part that cannot be executed (facebookresearch/sam2) vs `transformers`' independent implementation, base+ / large."""
    from transformers import Sam2Model
    from circuitvision_b200 import synth
    oracle = sam2_oracle.build_oracle(variant, seed=0)
    hf = Sam2Model(_hf_config(variant)).eval()
    sd = {}
    for k, v in oracle.state_dict().items():
        hk = _to_hf_key(k)
        if hk is not None:
            sd[hk] = v
    missing, unexpected = hf.load_state_dict(sd, strict=False)
    assert not unexpected, unexpected
    needed = [m for m in missing if m.startswith(("vision_encoder.backbone", "vision_encoder.neck.convs", "mask_decoder",
                                                  "shared_image_embedding"))]
    assert not needed, needed
    _, _, rgb = synth.make_schematic(5, 1024, render_rgb=True)
    x = sam2_oracle.preprocess_rgb(rgb)[None]
    with torch.no_grad():
        high, low, iou, aux = oracle(x, return_aux=True)
        fpn = hf.vision_encoder(pixel_values=x).fpn_hidden_states
        for a, b in zip(aux["fpn"], fpn):
            assert torch.allclose(a, b, atol=5e-4, rtol=1e-4), (a - b).abs().max()
        s0, s1 = hf.mask_decoder.conv_s0(fpn[0]), hf.mask_decoder.conv_s1(fpn[1])
        dense = (oracle.dense_embedding1 @ oracle.dense_embedding2).view(1, 256, 64, 64)
        masks, hiou, _, _ = hf.mask_decoder(
            image_embeddings=fpn[2], image_positional_embeddings=hf.get_image_wide_positional_embeddings(),
            sparse_prompt_embeddings=oracle.sparse_embedding[:, None], dense_prompt_embeddings=dense,
            multimask_output=False, high_resolution_features=[s0, s1])
    ref_low = masks[:, 0]
    assert (low - ref_low).abs().max().item() < 5e-4 * max(1.0, ref_low.abs().max().item())
    assert torch.allclose(iou, hiou[:, 0], atol=1e-5)
