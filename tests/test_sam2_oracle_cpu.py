"""The fp32 SAM 2.1 restatement (oracle/sam2_oracle.py) against the independent implementation of the same
architecture in `transformers` (modeling_sam2.py) with identical weights — SURVEY.md §C.2.  fp32 round-off only."""
import numpy as np
import pytest
import torch

from oracle import sam2_oracle


def _to_hf_key(k: str):
    """oracle (upstream-style) parameter name -> transformers Sam2Model parameter name."""
    if not k.startswith("sam2_model."):
        return None
    k = k[len("sam2_model."):]
    if k == "sam_prompt_encoder.pe_layer.positional_encoding_gaussian_matrix":
        return "shared_image_embedding.positional_embedding"
    if k.startswith("image_encoder.trunk."):
        k = "vision_encoder.backbone." + k[len("image_encoder.trunk."):]
        k = k.replace("patch_embed.proj.", "patch_embed.projection.")
        k = k.replace(".norm1.", ".layer_norm1.").replace(".norm2.", ".layer_norm2.")
        k = k.replace(".mlp.layers.0.", ".mlp.proj_in.").replace(".mlp.layers.1.", ".mlp.proj_out.")
        return k
    if k.startswith("image_encoder.neck.convs."):
        return "vision_encoder.neck.convs." + k[len("image_encoder.neck.convs."):].replace(".conv.", ".")
    if k.startswith("sam_mask_decoder."):
        k = "mask_decoder." + k[len("sam_mask_decoder."):]
        k = k.replace(".out_proj.", ".o_proj.")
        for i in range(1, 5):
            k = k.replace(f".norm{i}.", f".layer_norm{i}.")
        k = k.replace("transformer.norm_final_attn.", "transformer.layer_norm_final_attn.")
        k = k.replace("output_upscaling.0.", "upscale_conv1.").replace("output_upscaling.1.", "upscale_layer_norm.")
        k = k.replace("output_upscaling.3.", "upscale_conv2.")
        if ".mlp.layers." in k:  # two-layer transformer MLP
            k = k.replace(".mlp.layers.0.", ".mlp.proj_in.").replace(".mlp.layers.1.", ".mlp.proj_out.")
        elif ".layers." in k and ("hypernetworks" in k or "prediction_head" in k or "obj_score_head" in k):
            k = k.replace(".layers.0.", ".proj_in.").replace(".layers.1.", ".layers.0.").replace(".layers.2.", ".proj_out.")
        return k
    return None


@pytest.fixture(scope="module")
def pair():
    from transformers import Sam2Config, Sam2Model
    torch.manual_seed(0)
    oracle = sam2_oracle.build_oracle("tiny", seed=0)
    hf = Sam2Model(Sam2Config()).eval()
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
    return oracle, hf


def test_block_plan_tiny():
    plan = sam2_oracle.block_plan("tiny")
    assert len(plan) == 12
    assert plan[0] == (96, 96, 1, 8, False) and plan[1] == (96, 192, 2, 8, True) and plan[3] == (192, 384, 4, 4, True)
    assert [p[3] for p in plan[4:10]] == [14, 0, 14, 0, 14, 0] and plan[10] == (384, 768, 8, 14, True)
    assert plan[11] == (768, 768, 8, 7, False)
    assert len(sam2_oracle.block_plan("base_plus")) == 24 and len(sam2_oracle.block_plan("large")) == 48


def test_oracle_matches_transformers_sam2(pair):
    from circuitvision_b200 import synth
    oracle, hf = pair
    _, _, rgb = synth.make_schematic(5, 1024, render_rgb=True)
    x = sam2_oracle.preprocess_rgb(rgb)[None]
    with torch.no_grad():
        high, low, iou, aux = oracle(x, return_aux=True)
        fpn = hf.vision_encoder(pixel_values=x).fpn_hidden_states
        for a, b in zip(aux["fpn"], fpn):
            assert torch.allclose(a, b, atol=2e-4, rtol=1e-4), (a - b).abs().max()
        s0 = hf.mask_decoder.conv_s0(fpn[0])
        s1 = hf.mask_decoder.conv_s1(fpn[1])
        dense = (oracle.dense_embedding1 @ oracle.dense_embedding2).view(1, 256, 64, 64)
        masks, hiou, _, _ = hf.mask_decoder(
            image_embeddings=fpn[2], image_positional_embeddings=hf.get_image_wide_positional_embeddings(),
            sparse_prompt_embeddings=oracle.sparse_embedding[:, None], dense_prompt_embeddings=dense,
            multimask_output=False, high_resolution_features=[s0, s1])
    ref_low = masks[:, 0]
    scale = ref_low.abs().max().item()
    assert (low - ref_low).abs().max().item() < 2e-4 * max(1.0, scale)
    assert torch.allclose(iou, hiou[:, 0], atol=1e-5)
    assert high.shape == (1, 1, 1024, 1024) and low.shape == (1, 1, 256, 256) and iou.shape == (1, 1)
    # logits are not degenerate with this init (SURVEY §7 hard part 4)
    assert low.std().item() > 1e-3


def test_batched_equals_stack_of_singles():
    from circuitvision_b200 import synth
    oracle = sam2_oracle.build_oracle("tiny", seed=1)
    xs = torch.stack([sam2_oracle.preprocess_rgb(synth.make_schematic(s, 1024, render_rgb=True)[2]) for s in (1, 2)])
    with torch.no_grad():
        hb, lb, ib = oracle(xs)
        for i in range(2):
            h1, l1, i1 = oracle(xs[i:i + 1])
            assert torch.allclose(lb[i], l1[0], atol=1e-4) and torch.allclose(ib[i], i1[0], atol=1e-5)


def test_preprocess_matches_torchvision():
    from torchvision.transforms import Normalize, Resize, ToTensor
    from PIL import Image
    rng = np.random.default_rng(0)
    for hw in [(1024, 1024), (493, 712), (1500, 1100)]:
        img = rng.integers(0, 256, hw + (3,), dtype=np.uint8)
        ref = Normalize(sam2_oracle.MEAN, sam2_oracle.STD)(Resize((1024, 1024))(ToTensor()(Image.fromarray(img))))
        got = sam2_oracle.preprocess_rgb(img)
        assert torch.allclose(ref, got, atol=1e-5), hw
