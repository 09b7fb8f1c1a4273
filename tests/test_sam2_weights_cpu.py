"""Host logic of the SAM 2.1 path that needs no GPU: parameter-name parity with the oracle (and therefore with the
upstream module tree), load-time folding checked against the oracle's own sub-modules, checkpoint ingestion
(PEFT prefix + LoRA merge), and the loud failure of the compute entry points without a CUDA device."""
import os

import pytest
import torch
import torch.nn.functional as F

from circuitvision_b200 import sam2_infer, sam2_weights
from oracle import sam2_oracle


@pytest.fixture(scope="module")
def oracle():
    return sam2_oracle.build_oracle("tiny", seed=0)


@pytest.fixture(scope="module")
def folded(oracle):
    m = sam2_infer.get_modified_sam2("tiny", None, device="cpu", use_refinement_layer=True)
    res = m.load_state_dict(oracle.state_dict())
    assert not res.missing_keys and not res.unexpected_keys
    return sam2_weights.fold_state_dict(m.state_dict(), m.sam2_model.variant, True)


@pytest.mark.parametrize("variant", ["tiny", "small", "base_plus", "large"])
def test_parameter_names_and_shapes_match_oracle(variant):
    ours = sam2_weights.SAM2Params(sam2_weights.VARIANTS[variant]).state_dict()
    theirs = sam2_oracle.SAM2Base(variant).state_dict()
    assert set(ours) == set(theirs)
    for k in ours:
        assert ours[k].shape == theirs[k].shape, k
    assert sam2_weights.block_plan(sam2_weights.VARIANTS[variant]) == sam2_oracle.block_plan(variant)


def test_variant_from_reference_yaml():
    p = "/root/reference/models/configs/sam2.1_hiera_l.yaml"
    if not os.path.exists(p):
        pytest.skip("reference tree not mounted")
    assert sam2_weights.variant_from_yaml(p) == sam2_weights.VARIANTS["large"]


def test_folded_constants_match_oracle_modules(oracle, folded):
    dec = oracle.sam2_model.sam_mask_decoder
    tr = dec.transformer
    with torch.no_grad():
        tokens = torch.cat([dec.obj_score_token.weight, dec.iou_token.weight, dec.mask_tokens.weight,
                            oracle.sparse_embedding[0]], 0)
        assert torch.allclose(folded["tok0"], tokens)
        l0 = tr.layers[0]
        q1 = l0.norm1(l0.self_attn(tokens[None], tokens[None], tokens[None]))[0]
        assert torch.allclose(folded["l0.q1"], q1, atol=1e-5)
        qc = l0.cross_attn_token_to_image.q_proj(q1 + tokens)
        assert torch.allclose(folded["l0.t2i.qc"], qc, atol=1e-5)
        pe = oracle.sam2_model.sam_prompt_encoder.get_dense_pe().flatten(2)[0].t()  # [4096, 256]
        kk = l0.cross_attn_token_to_image.k_proj
        assert torch.allclose(folded["l0.t2i.kpe"][:, :128], pe @ kk.weight.t(), atol=1e-4)
        assert not folded["l0.t2i.kpe"][:, 128:].any()
        dense = (oracle.dense_embedding1 @ oracle.dense_embedding2)[0].t()
        assert torch.allclose(folded["dense"], dense, atol=1e-5)
        pos = oracle.sam2_model.image_encoder.trunk.pos((256, 256))[0].reshape(65536, 96)
        assert torch.allclose(folded["pos"], pos, atol=1e-6)
        # neck lateral conv composed with conv_s0 (1x1 o 1x1)
        x = torch.randn(1, 96, 8, 8)
        want = dec.conv_s0(oracle.sam2_model.image_encoder.neck.convs[3].conv(x))
        got = F.conv2d(x, folded["s0.w"].float()[:, :, None, None], folded["s0.b"])
        assert torch.allclose(got, want, atol=2e-2)  # bf16 weights
        # ConvTranspose2d(k2,s2) as a GEMM + pixel shuffle
        x = torch.randn(1, 256, 4, 4)
        want = dec.output_upscaling[0](x)
        y = x[0].flatten(1).t() @ folded["up1.w"].float().t()  # [16, 4*64]
        y = y.view(4, 4, 2, 2, 64).permute(4, 0, 2, 1, 3).reshape(64, 8, 8) + folded["up1.b"][:, None, None]
        assert torch.allclose(y, want[0], atol=3e-2)
    w = folded["pe.w"].float()
    assert w.shape == (96, 2 * sam2_weights.PE_K) and torch.equal(w[:, :152], w[:, 152:]) and not w[:, 147:152].any()
    assert all(t.is_contiguous() and t.dtype in (torch.float32, torch.bfloat16) for t in folded.values())


def test_checkpoint_ingestion_merges_lora(oracle):
    base = oracle.state_dict()
    target = "sam2_model.sam_mask_decoder.transformer.layers.0.self_attn.k_proj"
    conv_t = "sam2_model.sam_mask_decoder.conv_s0"
    g = torch.Generator().manual_seed(1)
    peft = {}
    for k, v in base.items():
        if k.startswith("sam2_model."):
            k2 = "sam2_model.base_model.model." + k[len("sam2_model."):]
        else:
            k2 = k
        for t in (target, conv_t):
            tp = "sam2_model.base_model.model." + t[len("sam2_model."):]
            if k2 == tp + ".weight":
                k2 = tp + ".base_layer.weight"
            elif k2 == tp + ".bias":
                k2 = tp + ".base_layer.bias"
        peft[k2] = v
    A1, B1 = torch.randn(4, 256, generator=g), torch.randn(256, 4, generator=g)
    A2, B2 = torch.randn(4, 256, 1, 1, generator=g), torch.randn(32, 4, 1, 1, generator=g)
    pre = "sam2_model.base_model.model."
    peft[pre + target[11:] + ".lora_A.default.weight"], peft[pre + target[11:] + ".lora_B.default.weight"] = A1, B1
    peft[pre + conv_t[11:] + ".lora_A.default.weight"], peft[pre + conv_t[11:] + ".lora_B.default.weight"] = A2, B2
    peft["sam2_model.base_model.model.memory_attention.layers.0.linear1.weight"] = torch.zeros(3, 3)  # not on the path
    m = sam2_infer.get_modified_sam2("tiny", None, device="cpu", use_refinement_layer=True, lora_rank=4, lora_alpha=16)
    res = m.load_state_dict({"state_dict": peft})
    assert not res.missing_keys and not res.unexpected_keys
    assert sorted(m.last_load_report["merged"]) == sorted([target, conv_t])
    assert m.last_load_report["ignored"] == ["sam2_model.memory_attention.layers.0.linear1.weight"]
    sd = m.state_dict()
    assert torch.allclose(sd[target + ".weight"], base[target + ".weight"] + 4.0 * (B1 @ A1), atol=1e-5)
    assert torch.allclose(sd[conv_t + ".weight"], base[conv_t + ".weight"] + 4.0 * (B2.flatten(1) @ A2.flatten(1))[:, :, None, None],
                          atol=1e-5)
    untouched = "sam2_model.image_encoder.trunk.blocks.3.attn.qkv.weight"
    assert torch.equal(sd[untouched], base[untouched])


def test_api_surface_matches_reference_signatures():
    import inspect
    sig = inspect.signature(sam2_infer.get_modified_sam2)
    assert list(sig.parameters)[:14] == ["model_cfg_path", "checkpoint_path", "device", "use_high_res_features", "use_peft",
                                         "lora_rank", "lora_alpha", "lora_dropout", "lora_target_modules", "use_wrapper",
                                         "trainable_embedding_r", "use_refinement_layer", "refinement_kernels",
                                         "kernel_channels"]
    assert sig.parameters["lora_rank"].default == 12 and sig.parameters["refinement_kernels"].default == [3, 5, 7, 11]
    f = inspect.signature(sam2_infer.SAM2ImageWrapper.forward)
    assert list(f.parameters) == ["self", "images", "points", "point_labels", "masks_prompt", "multimask_output"]
    w = inspect.signature(sam2_infer.SAM2ImageWrapper.__init__)
    assert w.parameters["refinement_kernel_sizes"].default == [3, 5, 7, 9, 11] and w.parameters["embedding_r"].default == 4
    t = sam2_infer.SAM2Transforms(resolution=1024, mask_threshold=0.0, max_hole_area=0.0, max_sprinkle_area=0.0)
    c = t.transform_coords(torch.tensor([[[100.0, 50.0]]]), normalize=True, orig_hw=(500, 1000))
    assert torch.allclose(c, torch.tensor([[[102.4, 102.4]]]))
    assert t.transform_boxes(torch.tensor([[0.0, 0.0, 1.0, 1.0]])).shape == (1, 2, 2)
    m = sam2_infer.get_modified_sam2("tiny", None, device="cpu", use_refinement_layer=True)
    assert m.sam2_model.image_size == 1024  # circuit_analyzer.py:237-238


def test_no_cpu_fallback():
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from circuitvision_b200 import CvError
    m = sam2_infer.get_modified_sam2("tiny", None, device="cpu", use_refinement_layer=True)
    with pytest.raises(CvError):
        m(torch.zeros(1, 3, 1024, 1024))
    with pytest.raises(CvError):
        sam2_infer.SAM2Transforms(1024, 0.0)(torch.zeros(4, 4, 3, dtype=torch.uint8).numpy())
    with pytest.raises(CvError):
        m.refinement_layer(torch.zeros(1, 1, 1024, 1024))
