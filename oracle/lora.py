"""TEST INFRASTRUCTURE ONLY.  LoRA as `peft` runs it at inference (un-merged): y = base(x) + (alpha / r) * B(A(x)) on
every module the reference targets (`/root/reference/src/circuit_analyzer.py:156-199`, r = 4, alpha = 16 at :210-211),
applied to the fp32 oracle, plus the PEFT-style checkpoint of that model (key layout of `/root/reference/src/
sam2_infer.py:396` `sam2_model.base_model.model.` + `<module>.base_layer.{weight,bias}` /
`<module>.lora_A.default.weight` / `<module>.lora_B.default.weight`), i.e. what `circuit_analyzer.py:227-233` feeds to
`load_state_dict`.  `peft` itself is not installable offline; its Linear / Conv2d LoRA forward is the formula above
(dropout is the identity in eval mode)."""
from __future__ import annotations

import torch
import torch.nn as nn

# circuit_analyzer.py:156-199 (base_parts + added_parts), verbatim module paths of the upstream sam2 tree
BASE_PARTS = [f"sam_mask_decoder.transformer.layers.{l}.{a}.{p}" for a in ("self_attn", "cross_attn_token_to_image")
              for l in (0, 1) for p in ("k_proj", "q_proj", "v_proj", "out_proj")] + \
             [f"sam_mask_decoder.transformer.layers.{l}.mlp.layers.{j}" for l in (0, 1) for j in (0, 1)]
ADDED_PARTS = [
    "sam_mask_decoder.iou_prediction_head.layers.2", "sam_mask_decoder.conv_s0", "sam_mask_decoder.conv_s1",
    "image_encoder.neck.convs.2.conv", "image_encoder.neck.convs.3.conv",
    "image_encoder.trunk.blocks.44.attn.qkv", "image_encoder.trunk.blocks.44.mlp.layers.0", "image_encoder.trunk.blocks.44.proj",
    "image_encoder.trunk.blocks.47.attn.qkv", "image_encoder.trunk.blocks.47.mlp.layers.0",
] + [f"sam_mask_decoder.transformer.layers.{l}.cross_attn_image_to_token.{p}" for l in (0, 1) for p in ("q_proj", "k_proj", "v_proj")]
REFERENCE_TARGETS = BASE_PARTS + ADDED_PARTS


class LoraLinear(nn.Module):
    def __init__(self, base: nn.Linear, r: int, alpha: float, g: torch.Generator, b_std: float):
        super().__init__()
        self.base_layer = base
        self.lora_A = nn.Linear(base.in_features, r, bias=False)
        self.lora_B = nn.Linear(r, base.out_features, bias=False)
        self.scaling = alpha / r
        with torch.no_grad():
            self.lora_A.weight.copy_(torch.randn(self.lora_A.weight.shape, generator=g) / base.in_features ** 0.5)
            self.lora_B.weight.copy_(torch.randn(self.lora_B.weight.shape, generator=g) * b_std)

    def forward(self, x):
        return self.base_layer(x) + self.scaling * self.lora_B(self.lora_A(x))


class LoraConv1x1(nn.Module):
    def __init__(self, base: nn.Conv2d, r: int, alpha: float, g: torch.Generator, b_std: float):
        super().__init__()
        assert base.kernel_size == (1, 1)
        self.base_layer = base
        self.lora_A = nn.Conv2d(base.in_channels, r, 1, bias=False)
        self.lora_B = nn.Conv2d(r, base.out_channels, 1, bias=False)
        self.scaling = alpha / r
        with torch.no_grad():
            self.lora_A.weight.copy_(torch.randn(self.lora_A.weight.shape, generator=g) / base.in_channels ** 0.5)
            self.lora_B.weight.copy_(torch.randn(self.lora_B.weight.shape, generator=g) * b_std)

    def forward(self, x):
        return self.base_layer(x) + self.scaling * self.lora_B(self.lora_A(x))


def _resolve(root: nn.Module, path: str):
    cur = root
    parts = path.split(".")
    for p in parts[:-1]:
        cur = cur[int(p)] if p.isdigit() else getattr(cur, p)
    return cur, parts[-1]


def apply_lora(oracle, targets=None, r: int = 4, alpha: float = 16.0, seed: int = 7, b_std: float = 0.02):
    """Wraps the target modules of `oracle.sam2_model` IN PLACE with explicit-factor LoRA modules (non-zero B so the update
    matters) and returns (applied target paths, PEFT-style state dict of the whole wrapper).  Targets that do not exist in
    this variant (trunk blocks 44 / 47 exist only in `large`) are skipped, as `peft` would simply not match them."""
    g = torch.Generator().manual_seed(seed)
    applied = []
    for t in (REFERENCE_TARGETS if targets is None else targets):
        try:
            parent, leaf = _resolve(oracle.sam2_model, t)
            mod = parent[int(leaf)] if leaf.isdigit() else getattr(parent, leaf)
        except (AttributeError, IndexError):
            continue
        if isinstance(mod, nn.Linear):
            new = LoraLinear(mod, r, alpha, g, b_std)
        elif isinstance(mod, nn.Conv2d):
            new = LoraConv1x1(mod, r, alpha, g, b_std)
        else:
            raise TypeError(f"{t}: {type(mod).__name__} is not a LoRA target type")
        if leaf.isdigit():
            parent[int(leaf)] = new
        else:
            setattr(parent, leaf, new)
        applied.append(t)
    sd = {}
    for k, v in oracle.state_dict().items():
        if k.startswith("sam2_model."):
            k = "sam2_model.base_model.model." + k[len("sam2_model."):]
            k = k.replace(".lora_A.weight", ".lora_A.default.weight").replace(".lora_B.weight", ".lora_B.default.weight")
        sd[k] = v.detach().clone()
    return applied, sd
