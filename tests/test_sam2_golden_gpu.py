"""GPU parity of the SAM 2.1 path against fixtures produced by EXECUTING the reference's own code
(tests/golden/sam2_golden.npz, oracle/gen_sam2_golden.py: the unmodified SAM2Transforms / MultiKernelRefinement /
SAM2ImageWrapper.forward of src/sam2_infer.py and CircuitAnalyzer.segment_with_sam2 of src/circuit_analyzer.py:321-386 over
a shim of the third-party sam2 package).  No CPU forward runs here: weights regenerate from seeds, inputs from the
synthetic generator, and the CUDA path (through the drop-in classes -> C ABI) is compared with the stored outputs.

Tolerances (16-bit tensor-core operands, fp32 accumulation / residual stream / softmax / LayerNorm; written per test):
  thresholded mask IoU vs the reference >= 0.99 (BASELINE.json north_star), fp16 default held to >= 0.997 on tiny;
  low-res logits: max |err| <= 2 % and mean |err| <= 0.4 % of the logit standard deviation; predicted IoU |err| <= 1e-3;
  SAM2Transforms output: fp32 round-off (2e-5)."""
import json
import os

import numpy as np
import pytest
import torch

from oracle import gen_sam2_golden, sam2_oracle

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def sam2_golden():
    z = np.load(os.path.join(ROOT, "tests", "golden", "sam2_golden.npz"))
    return z, json.loads(bytes(z["meta_json"]).decode())


def _iou(a, b):
    inter, union = int((a & b).sum()), int((a | b).sum())
    return inter / union if union else 1.0


def _model(variant, seed=0, dtype=torch.float16):
    from circuitvision_b200 import sam2_infer
    ref = sam2_oracle.build_oracle(variant, seed=seed)  # weights only; its forward is never called in this file
    model = sam2_infer.get_modified_sam2(variant, None, device="cuda:0", use_refinement_layer=True)
    res = model.load_state_dict(ref.state_dict())
    assert not res.missing_keys and not res.unexpected_keys
    model.set_operand_dtype(dtype)
    model.engine().set_debug(True)  # count saturated fp16 conversions (tc::pack16 clips silently otherwise)
    return model


def _check_case(model, z, meta, name, iou_gate, low_gate=(0.02, 0.004)):
    from circuitvision_b200 import sam2_infer
    from circuitvision_b200.circuit_analyzer import CircuitAnalyzer
    m = meta[name]
    img = gen_sam2_golden.case_image(m["image_seed"], m["hw"])
    tr = sam2_infer.SAM2Transforms(1024, 0.0)
    A = CircuitAnalyzer(sam2_model=model, sam2_transforms=tr, debug=True, device=0)
    # --- the drop-in call (circuit_analyzer.py:321-386)
    mask, colored, bbox = A.segment_with_sam2(img)
    want = np.unpackbits(z[name + "/mask"])[:m["hw"][0] * m["hw"][1]].reshape(m["hw"]).astype(bool)
    assert mask.shape == tuple(m["hw"]) and mask.dtype == np.uint8 and set(np.unique(mask)) <= {0, 255}
    iou_fg = _iou(mask > 0, want)
    assert iou_fg >= iou_gate, (name, iou_fg)
    wb = tuple(int(v) for v in z[name + "/bbox"])
    assert bbox == (wb if wb[0] >= 0 else None), (name, bbox, wb)
    assert np.array_equal(colored[:, :, 1], mask) and not colored[:, :, 0].any() and not colored[:, :, 2].any()
    # --- wrapper level (sam2_infer.py:49-51, :220-275): transforms output, low-res logits, IoU head, refined high-res logits
    x = tr(np.ascontiguousarray(img[:, :, ::-1]))  # the :343 channel swap
    assert np.abs(x[:, ::16, ::16].cpu().numpy() - z[name + "/x_sub"]).max() <= 2e-5
    assert abs(float(x.double().sum()) - m["x_sum"]) <= 1e-4 * max(1.0, abs(m["x_sum"])) + 5.0
    model.set_max_batch(1)
    high, low, iou = model(x[None])
    torch.cuda.synchronize()
    d = np.abs(low[0, 0].cpu().numpy() - z[name + "/low"])
    assert d.max() <= low_gate[0] * m["low_std"] and d.mean() <= low_gate[1] * m["low_std"], (name, d.max() / m["low_std"], d.mean() / m["low_std"])
    assert abs(float(iou) - float(z[name + "/iou"][0])) <= 1e-3
    hs = high[0, 0, ::8, ::8].cpu().numpy()
    ref_hs = z[name + "/high_sub"]
    assert np.abs(hs - ref_hs).max() <= 0.04 * m["high_std"], (name, np.abs(hs - ref_hs).max() / m["high_std"])
    # a gate that cannot be vacuous when the mask is almost all foreground / background: threshold at the reference median
    med = float(np.median(ref_hs))
    bal = _iou(hs > med, ref_hs > med)
    assert bal >= 0.97, (name, bal)
    return iou_fg, d.max() / m["low_std"], bal


@pytest.mark.parametrize("variant", ["tiny", "small", "base_plus", "large"])
def test_cuda_path_matches_reference_executed_fixtures(sam2_golden, variant):
    z, meta = sam2_golden
    model = _model(variant)
    names = [n for n, m in meta.items() if m["variant"] == variant]
    assert names
    for n in names:
        _check_case(model, z, meta, n, iou_gate=0.997 if variant == "tiny" else 0.99,
                    low_gate=(0.015, 0.003) if variant == "tiny" else (0.02, 0.004))
    # no fp16 conversion of the last forward saturated (tc::pack16 uses cvt.rn.satfinite, which would clip silently)
    assert model.engine().saturation_count() == 0
    del model
    torch.cuda.empty_cache()


def test_bf16_operand_mode_against_reference_fixtures(sam2_golden):
    """north_star's nominal operand format (bf16, 8-bit significand) against the reference-executed fixtures.  With
    random-init weights bf16 operand rounding alone costs more than the 0.99 gate allows on some images (measured on the
    B200: IoU 0.987 on tiny_s5; CPU emulation of the roundings, scripts/error_budget.py: 0.993 +- 0.004) — DESIGN.md §2
    states the shortfall; the gate here is the floor bf16 is held to, the default fp16 operand format meets >= 0.997."""
    z, meta = sam2_golden
    model = _model("tiny", dtype=torch.bfloat16)
    for n in ("tiny_s5", "tiny_s6", "tiny_s7", "tiny_s21_720x1280"):
        _check_case(model, z, meta, n, iou_gate=0.98, low_gate=(0.08, 0.015))
    del model
    torch.cuda.empty_cache()


def test_baseplus_batch256_chunk64_matches_fixtures(sam2_golden):
    """BASELINE configs[2] as bench.py runs it: 256 base+ crops in engine passes of 64.  Four probes sit at the chunk
    boundaries (image 0, 63, 64, 255); their masks and low-res logits must match the per-image reference fixtures."""
    z, meta = sam2_golden
    model = _model("base_plus")
    model.set_max_batch(64)
    probes = {0: "base_plus_s5", 63: "base_plus_s6", 64: "base_plus_s7", 255: "base_plus_s8"}
    filler = [gen_sam2_golden.case_image(100 + i, (1024, 1024)) for i in range(6)]
    batch = np.stack([gen_sam2_golden.case_image(meta[probes[i]]["image_seed"], (1024, 1024)) if i in probes else filler[i % 6]
                      for i in range(256)])
    d = torch.from_numpy(batch).cuda()
    eng = model.engine()
    r = eng.forward(d, 0, True, want_high=False, want_low=True, want_mask=True)  # uint8 crops, :343 swap on the device
    torch.cuda.synchronize()
    assert eng.launches > 4 * 100
    for i, name in probes.items():
        m = meta[name]
        want = np.unpackbits(z[name + "/mask"])[:1024 * 1024].reshape(1024, 1024).astype(bool)
        assert _iou(r["mask"][i].cpu().numpy() > 0, want) >= 0.99, name
        dl = np.abs(r["low"][i, 0].cpu().numpy() - z[name + "/low"])
        assert dl.max() <= 0.03 * m["low_std"] and dl.mean() <= 0.004 * m["low_std"], (name, dl.max() / m["low_std"])
    assert eng.saturation_count() == 0
    del model, eng, d
    torch.cuda.empty_cache()


def test_reference_tail_fixtures(sam2_golden):
    """MultiKernelRefinement (:130-189) and postprocess_masks (:88-128) executed by the reference on seeded logits."""
    from circuitvision_b200 import sam2_infer
    z, _ = sam2_golden
    model = _model("tiny")
    x = gen_sam2_golden.tail_cases().cuda()
    got = model.refinement_layer(x)
    scale = max(1.0, float(np.abs(z["refine/out_sub"]).max()))
    assert np.abs(got[:, 0, ::8, ::8].cpu().numpy() - z["refine/out_sub"]).max() <= 1e-5 * scale
    border = torch.cat([got[:, 0, :12].flatten(1), got[:, 0, -12:].flatten(1), got[:, 0, :, :12].flatten(1),
                        got[:, 0, :, -12:].flatten(1)], 1).cpu().numpy()
    assert np.abs(border - z["refine/out_border"]).max() <= 1e-5 * scale  # 'same' zero padding at the image frame
    tr = sam2_infer.SAM2Transforms(1024, 0.0)
    for hw in [(493, 712), (1500, 1100)]:
        p = tr.postprocess_masks(x, hw)
        assert np.abs(p[:, 0, ::4, ::4].cpu().numpy() - z[f"post/{hw[0]}x{hw[1]}_sub"]).max() <= 1e-5


def test_peft_checkpoint_of_large_with_reference_lora_targets():
    """a PEFT-keyed state dict of the LARGE variant with LoRA on every module the reference targets
    (circuit_analyzer.py:156-199: decoder attention / MLPs, iou head, conv_s0/s1, neck.convs.{2,3}, trunk blocks 44 and 47)
    goes through get_modified_sam2 + load_state_dict (circuit_analyzer.py:203-233); the CUDA forward must match the oracle
    RUNNING the explicit A/B factors (oracle/lora.py), and must differ from the un-adapted model."""
    from circuitvision_b200 import sam2_infer, synth
    from oracle import lora
    ref = sam2_oracle.build_oracle("large", seed=0)
    base_sd = {k: v.clone() for k, v in ref.state_dict().items()}
    applied, peft_sd = lora.apply_lora(ref, r=4, alpha=16.0, seed=7, b_std=0.05)
    assert set(applied) == set(lora.REFERENCE_TARGETS)  # all 36 targets exist in large
    model = sam2_infer.get_modified_sam2("large", None, device="cuda:0", use_peft=True, lora_rank=4, lora_alpha=16,
                                         lora_dropout=0.3, lora_target_modules=lora.REFERENCE_TARGETS, use_wrapper=True,
                                         trainable_embedding_r=4, use_refinement_layer=True, refinement_kernels=[3, 5, 7, 11],
                                         kernel_channels=2)
    res = model.load_state_dict({"state_dict": peft_sd}["state_dict"])
    assert not res.missing_keys and not res.unexpected_keys
    assert sorted(model.last_load_report["merged"]) == sorted("sam2_model." + t for t in applied)
    x = sam2_oracle.preprocess_rgb(synth.make_schematic(5, 1024, render_rgb=True)[2])[None]
    with torch.no_grad():
        rh, rl, ri = ref(x)  # explicit-factor LoRA forward, fp32 CPU
    model.set_max_batch(1)
    high, low, iou = model(x.cuda())
    torch.cuda.synchronize()
    std = rl.std().item()
    d = (low.cpu() - rl).abs()
    assert d.max().item() <= 0.02 * std and d.mean().item() <= 0.004 * std, (d.max().item() / std, d.mean().item() / std)
    assert (iou.cpu() - ri).abs().max().item() <= 1e-3
    assert _iou((high.cpu() > 0).numpy(), (rh > 0).numpy()) >= 0.99
    # the adapters matter: the same model with the base weights only is measurably different
    model.load_state_dict(base_sd)
    _, low0, _ = model(x.cuda())
    assert (low0.cpu() - rl).abs().mean().item() > 5 * d.mean().item()
    del model
    torch.cuda.empty_cache()
