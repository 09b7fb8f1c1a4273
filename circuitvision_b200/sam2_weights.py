"""Parameter container and load-time weight folding for the SAM 2.1 image path.

The reference builds its model with the third-party `sam2` package (`/root/reference/src/sam2_infer.py:333`
`build_sam2`) and then reads three sub-modules of it (`image_encoder`, `sam_prompt_encoder`, `sam_mask_decoder`,
sam2_infer.py:226-260).  This module provides

* `SAM2Params` — an `nn.Module` tree that owns exactly those sub-modules' PARAMETERS under the upstream
  names (`image_encoder.trunk.blocks.N.attn.qkv.weight`, ...), so `state_dict()` / `load_state_dict()` are
  interchangeable with the reference's checkpoints for the image path.  It has no forward(): all arithmetic runs in
  libcv_b200.so.
* `fold_state_dict` — turns a wrapper state dict into the named device tensors the C engine consumes
  (include/cv_b200.h `cv_sam2_set_tensor`): bf16 GEMM operands, fp32 everything else, with every
  input-independent piece of the dataflow evaluated once here (positional embeddings, dense prompt, dense PE and its
  projections, layer-0 token self-attention, neck∘conv_s0/conv_s1 products, ConvTranspose weights re-laid-out as GEMM
  operands).  This is weight preparation, not a compute fallback: it never sees an image.
* `normalize_state_dict` — checkpoint ingestion: strips the PEFT prefix
  (`sam2_model.base_model.model.`, sam2_infer.py:396), merges LoRA pairs `W += (alpha/r)·B·A`
  (circuit_analyzer.py:203-223: r=4, alpha=16) and drops upstream keys outside the image path.
"""
from __future__ import annotations

import math

import numpy as np
import torch
import torch.nn as nn
import torch.nn.functional as F

VARIANTS = {
    # embed, heads, stages, global blocks, window spec, pos-embed background   (SURVEY §B.1;
    # large = /root/reference/models/configs/sam2.1_hiera_l.yaml:11-16,26)
    "tiny": dict(embed=96, heads=1, stages=(1, 2, 7, 2), global_blocks=(5, 7, 9), window_spec=(8, 4, 14, 7), bkg=(7, 7)),
    "small": dict(embed=96, heads=1, stages=(1, 2, 11, 2), global_blocks=(7, 10, 13), window_spec=(8, 4, 14, 7), bkg=(7, 7)),
    "base_plus": dict(embed=112, heads=2, stages=(2, 3, 16, 3), global_blocks=(12, 16, 20), window_spec=(8, 4, 14, 7),
                      bkg=(14, 14)),
    "large": dict(embed=144, heads=2, stages=(2, 6, 36, 4), global_blocks=(23, 33, 43), window_spec=(8, 4, 16, 8), bkg=(7, 7)),
}
MEAN = (0.485, 0.456, 0.406)  # sam2_infer.py:41-42
STD = (0.229, 0.224, 0.225)
PE_K8 = 168  # uint8 patch operand: 7 kernel rows x (21 taps + 3 zeros); must equal csrc/sam2_kernels.cuh PE_K8
PE_K = 152  # patch-embed K (3*7*7 = 147) padded to a multiple of 8; must equal csrc/sam2_kernels.cuh PE_K
REFINE_KERNELS = (3, 5, 7, 11)  # the kernel sizes the fused tail kernel is built for (circuit_analyzer.py:218)


def variant_from_yaml(path: str) -> dict:
    """Trunk hyper-parameters from an upstream hydra yaml (e.g. models/configs/sam2.1_hiera_l.yaml:6-28) without
    hydra: only `model.image_encoder.trunk.*` is read; absent keys take the upstream Hiera defaults."""
    import yaml
    with open(path) as f:
        y = yaml.safe_load(f)
    t = y["model"]["image_encoder"]["trunk"]
    return dict(embed=int(t.get("embed_dim", 96)), heads=int(t.get("num_heads", 1)),
                stages=tuple(t.get("stages", (2, 3, 16, 3))), global_blocks=tuple(t.get("global_att_blocks", (12, 16, 20))),
                window_spec=tuple(t.get("window_spec", (8, 4, 14, 7))),
                bkg=tuple(t.get("window_pos_embed_bkg_spatial_size", (14, 14))))


def padded_head_dim(v: dict) -> int:
    """Head dim of the attention kernels' buffers: the real head dim zero-padded to 64 or 96 (csrc/attn_tc.cu)."""
    hd = v["embed"] // v["heads"]
    if hd > 96 or hd % 8:
        raise ValueError(f"head dim {hd} is not supported (multiple of 8, <= 96)")
    return 64 if hd <= 64 else 96


def block_plan(v: dict):
    """[(dim_in, dim_out, heads, window, q_pool)] per trunk block: the first block of stage s>0 widens
    the channels, pools Q and still uses the previous stage's window."""
    plan, total = [], 0
    for s, nb in enumerate(v["stages"]):
        for b in range(nb):
            first = s > 0 and b == 0
            dim_out = v["embed"] * 2 ** s
            dim_in = v["embed"] * 2 ** (s - 1) if first else dim_out
            ws = v["window_spec"][s - 1] if first else v["window_spec"][s]
            if total in v["global_blocks"]:
                ws = 0
            plan.append((dim_in, dim_out, v["heads"] * 2 ** s, ws, first))
            total += 1
    return plan


# ------------------------------------------------------------------------------------------ parameter tree
def _put(root: nn.Module, path: str, mod: nn.Module):
    parts = path.split(".")
    cur = root
    for p in parts[:-1]:
        if p not in cur._modules:
            cur.add_module(p, nn.Module())
        cur = cur._modules[p]
    cur.add_module(parts[-1], mod)


class _Buffer(nn.Module):
    def __init__(self, name, t):
        super().__init__()
        self.register_buffer(name, t)


class SAM2Params(nn.Module):
    """Parameters of the three sub-modules the wrapper touches, under the upstream names.  Leaf modules are stock
    torch layers used purely as initialised parameter holders (PyTorch-default `reset_parameters()`, the
    'upstream-style' random init of SURVEY §7)."""

    image_size = 1024  # read by the reference at circuit_analyzer.py:237-238

    def __init__(self, variant: dict):
        super().__init__()
        self.variant = dict(variant)
        self.use_high_res_features_in_sam = True
        E = variant["embed"]
        lin, ln = nn.Linear, nn.LayerNorm
        P = "image_encoder.trunk."
        _put(self, P + "patch_embed.proj", nn.Conv2d(3, E, 7, 4, 3))
        trunk = self._modules["image_encoder"]._modules["trunk"]
        trunk.pos_embed = nn.Parameter(torch.zeros(1, E, *variant["bkg"]))
        ws0 = variant["window_spec"][0]
        trunk.pos_embed_window = nn.Parameter(torch.zeros(1, E, ws0, ws0))
        nn.init.trunc_normal_(trunk.pos_embed, std=0.02)
        nn.init.trunc_normal_(trunk.pos_embed_window, std=0.02)
        for i, (ci, co, _h, _ws, _pool) in enumerate(block_plan(variant)):
            b = f"{P}blocks.{i}."
            _put(self, b + "norm1", ln(ci, eps=1e-6))
            _put(self, b + "attn.qkv", lin(ci, 3 * co))
            _put(self, b + "attn.proj", lin(co, co))
            _put(self, b + "norm2", ln(co, eps=1e-6))
            _put(self, b + "mlp.layers.0", lin(co, 4 * co))
            _put(self, b + "mlp.layers.1", lin(4 * co, co))
            if ci != co:
                _put(self, b + "proj", lin(ci, co))
        for j, c in enumerate([8 * E, 4 * E, 2 * E, E]):
            _put(self, f"image_encoder.neck.convs.{j}.conv", nn.Conv2d(c, 256, 1))
        _put(self, "sam_prompt_encoder.pe_layer", _Buffer("positional_encoding_gaussian_matrix", torch.randn(2, 128)))
        D = "sam_mask_decoder."

        def attn(path, inner):
            for n, (a, b_) in dict(q_proj=(256, inner), k_proj=(256, inner), v_proj=(256, inner), out_proj=(inner, 256)).items():
                _put(self, f"{path}.{n}", lin(a, b_))

        for l in range(2):
            t = f"{D}transformer.layers.{l}."
            attn(t + "self_attn", 256)
            _put(self, t + "norm1", ln(256))
            attn(t + "cross_attn_token_to_image", 128)
            _put(self, t + "norm2", ln(256))
            _put(self, t + "mlp.layers.0", lin(256, 2048))
            _put(self, t + "mlp.layers.1", lin(2048, 256))
            _put(self, t + "norm3", ln(256))
            _put(self, t + "norm4", ln(256))
            attn(t + "cross_attn_image_to_token", 128)
        attn(D + "transformer.final_attn_token_to_image", 128)
        _put(self, D + "transformer.norm_final_attn", ln(256))
        _put(self, D + "iou_token", nn.Embedding(1, 256))
        _put(self, D + "mask_tokens", nn.Embedding(4, 256))
        _put(self, D + "obj_score_token", nn.Embedding(1, 256))
        _put(self, D + "output_upscaling.0", nn.ConvTranspose2d(256, 64, 2, 2))
        _put(self, D + "output_upscaling.1", ln(64, eps=1e-6))  # LayerNorm2d: weight/bias [64]
        _put(self, D + "output_upscaling.3", nn.ConvTranspose2d(64, 32, 2, 2))
        _put(self, D + "conv_s0", nn.Conv2d(256, 32, 1))
        _put(self, D + "conv_s1", nn.Conv2d(256, 64, 1))
        for k in range(4):
            for j, (a, b_) in enumerate([(256, 256), (256, 256), (256, 32)]):
                _put(self, f"{D}output_hypernetworks_mlps.{k}.layers.{j}", lin(a, b_))
        for j, (a, b_) in enumerate([(256, 256), (256, 256), (256, 4)]):
            _put(self, f"{D}iou_prediction_head.layers.{j}", lin(a, b_))
        for j, (a, b_) in enumerate([(256, 256), (256, 256), (256, 1)]):
            _put(self, f"{D}pred_obj_score_head.layers.{j}", lin(a, b_))

    def forward(self, *a, **k):
        raise RuntimeError("SAM2Params holds parameters only; the forward pass runs in libcv_b200.so via SAM2ImageWrapper")


# ------------------------------------------------------------------------------------------ checkpoint ingestion
PEFT_PREFIX = "sam2_model.base_model.model."  # sam2_infer.py:396


def normalize_state_dict(sd: dict, lora_alpha: float = 16.0, lora_rank=None, wanted=None):
    """Reference fine-tuned checkpoint (PEFT-wrapped names, LoRA factor pairs) -> plain wrapper state dict.

    `…X.base_layer.weight` + `…X.lora_A.default.weight` [r,in] + `…X.lora_B.default.weight` [out,r]
        ->  `…X.weight` = base + (alpha / r) · B · A            (Linear; 1x1 Conv2d factors are [r,in,1,1]/[out,r,1,1])
    Returns (state_dict, report) where report lists merged modules and ignored keys.  `wanted`: optional set of
    keys to keep (everything else, e.g. the video-memory modules of upstream checkpoints, is dropped)."""
    if "state_dict" in sd and isinstance(sd["state_dict"], dict):  # circuit_analyzer.py:229-231
        sd = sd["state_dict"]
    elif "model" in sd and isinstance(sd["model"], dict):  # upstream sam2.1_hiera_*.pt
        sd = {"sam2_model." + k: v for k, v in sd["model"].items()}
    flat = {}
    for k, v in sd.items():
        if k.startswith(PEFT_PREFIX):
            k = "sam2_model." + k[len(PEFT_PREFIX):]
        flat[k] = v
    out, merged, lora = {}, [], {}
    for k, v in flat.items():
        if ".lora_A." in k or ".lora_B." in k:
            mod, rest = k.split(".lora_")
            lora.setdefault(mod, {})[rest[0]] = v
        elif ".lora_dropout" in k or ".lora_embedding" in k:
            continue
        else:
            out[k.replace(".base_layer.", ".")] = v
    for mod, ab in lora.items():
        if "A" not in ab or "B" not in ab or mod + ".weight" not in out:
            raise KeyError(f"incomplete LoRA triple for '{mod}'")
        A, B = ab["A"].double(), ab["B"].double()
        r = A.shape[0] if lora_rank is None else lora_rank
        W = out[mod + ".weight"]
        delta = (B.flatten(1) @ A.flatten(1)) * (lora_alpha / r)
        out[mod + ".weight"] = (W.double() + delta.reshape(W.shape)).to(W.dtype)
        merged.append(mod)
    ignored = []
    if wanted is not None:
        ignored = sorted(k for k in out if k not in wanted)
        out = {k: v for k, v in out.items() if k in wanted}
    return out, dict(merged=sorted(merged), ignored=ignored)


# ------------------------------------------------------------------------------------------ folding
_OPERAND_DTYPE = torch.bfloat16  # set per fold_state_dict call


def _bf16(t: torch.Tensor) -> torch.Tensor:
    """fp32/fp64 -> the 16-bit tensor-core operand format of this fold (bf16, or IEEE half saturated to +-65504)."""
    t = t.float().contiguous()
    if _OPERAND_DTYPE == torch.float16:
        t = t.clamp(-65504.0, 65504.0)
    return t.to(_OPERAND_DTYPE)


def _f32(t: torch.Tensor) -> torch.Tensor:
    return t.float().contiguous()


def dense_pe(gauss: torch.Tensor, size: int = 64) -> torch.Tensor:
    """sam_prompt_encoder.get_dense_pe() (sam2_infer.py:254) as a token-major [size², 256] table."""
    grid = torch.ones((size, size), dtype=torch.float32)
    y = (grid.cumsum(0) - 0.5) / size
    x = (grid.cumsum(1) - 0.5) / size
    c = 2 * torch.stack([x, y], -1) - 1
    c = 2 * np.pi * (c @ gauss.float())
    return torch.cat([torch.sin(c), torch.cos(c)], -1).reshape(size * size, 256)


def _attn_tokens(q, k, v, heads):
    T = q.shape[0]
    sp = lambda t: t.reshape(t.shape[0], heads, -1).transpose(0, 1)
    qh, kh, vh = sp(q), sp(k), sp(v)
    a = torch.softmax(qh @ kh.transpose(-1, -2) / math.sqrt(qh.shape[-1]), -1)
    return (a @ vh).transpose(0, 1).reshape(T, -1)


def upsample4_taps(rel: int):
    """Bilinear x4 (align_corners=False) away from the borders: output row 4*Y0 + rel reads low-res rows Y0 + i0 and
    Y0 + i0 + 1 with weights (1 - f, f); src = (4 Y0 + rel + 0.5) / 4 - 0.5 = Y0 + (rel - 1.5) / 4."""
    t = (rel - 1.5) / 4.0
    i0 = int(np.floor(t))
    return i0, t - i0


def refine_phase_tables(branch_weights) -> torch.Tensor:
    """Composite weights of  conv_k( upsample_x4(low) )  for the interior of the image, where the composition is a
    position-independent 5 x 5 stencil on the LOW-RES grid for each of the 16 output phases (py, px) = (Y % 4, X % 4):

        Z[a,c][4 Y0 + py][4 X0 + px] = sum_{tu,tv} T[py,px][tu,tv][a*4+c] * low[Y0 - 2 + tu][X0 - 2 + tv]

    (every tap dy in [-5,5] of the widest 11 x 11 branch reads low-res rows Y0-2 .. Y0+2; phases 0 and 3 leave one
    border row of the window at zero).  204 taps per channel group collapse to 25: csrc/sam2_kernels.cu k_tail_phase.
    Evaluated in float64.  branch_weights: four [4,k,k] tensors (k = 3,5,7,11).
    Returns [16 phases, 25 taps, 16 channels] float32."""
    T = np.zeros((4, 4, 5, 5, 16), np.float64)  # [py][px][tu][tv][ch]
    for a, w in enumerate(branch_weights):
        w = np.asarray(w.detach().cpu().double().numpy()) if hasattr(w, "detach") else np.asarray(w, np.float64)
        k = w.shape[-1]
        r = (k - 1) // 2
        B = np.zeros((4, k, 5))  # [phase][tap d + r][window position]
        for p in range(4):
            for d in range(-r, r + 1):
                i0, f = upsample4_taps(p + d)
                for ii, wt in ((i0, 1.0 - f), (i0 + 1, f)):
                    if wt != 0.0:
                        assert 0 <= ii + 2 < 5, (p, d, ii)
                        B[p, d + r, ii + 2] += wt
        for py in range(4):
            for px in range(4):
                for c in range(4):
                    T[py, px, :, :, a * 4 + c] = B[py].T @ w[c] @ B[px]
    return torch.from_numpy(T.reshape(16, 25, 16)).float()


def fold_state_dict(sd: dict, variant: dict, use_refinement: bool, operand_dtype=torch.bfloat16) -> dict:
    """Wrapper state dict (keys `sam2_model.…`, `dense_embedding1/2`, `sparse_embedding`, `refinement_layer.…`)
    -> {engine tensor name: contiguous CPU tensor (float32, or `operand_dtype` for tensor-core operands)}.
    Load-time only."""
    global _OPERAND_DTYPE
    if operand_dtype not in (torch.bfloat16, torch.float16):
        raise ValueError("operand_dtype must be torch.bfloat16 or torch.float16")
    _OPERAND_DTYPE = operand_dtype
    g = lambda k: sd["sam2_model." + k].detach().cpu().double()
    E = variant["embed"]
    out = {}
    # ---- patch embed: K index = c*49 + ky*7 + kx, padded to PE_K, duplicated for the (hi | lo) pixel split
    w = g("image_encoder.trunk.patch_embed.proj.weight").reshape(E, 147)
    w = F.pad(w, (0, PE_K - 147))
    out["pe.w"] = _bf16(torch.cat([w, w], 1))
    out["pe.b"] = _f32(g("image_encoder.trunk.patch_embed.proj.bias"))
    # ---- positional embedding: bicubic(background -> 256x256) + tiled window embedding
    pe = sd["sam2_model.image_encoder.trunk.pos_embed"].detach().cpu().float()
    pw = sd["sam2_model.image_encoder.trunk.pos_embed_window"].detach().cpu().float()
    pos = F.interpolate(pe, size=(256, 256), mode="bicubic")
    pos = pos + pw.tile([x // y for x, y in zip(pos.shape, pw.shape)])
    out["pos"] = _f32(pos[0].permute(1, 2, 0).reshape(65536, E))
    # uint8 path: A = raw pixel values (exact in bf16).  conv(Normalize(p/255)) = sum_inb (w / (255 std_c)) p
    #   - sum_inb w mean_c / std_c + b   ->  weights "pe.w8", and one additive per-token table "pos8" that also absorbs
    #   the conv bias, the positional embedding and the border dependence of the mean term (zero padding of the
    #   NORMALISED image: out-of-bounds taps contribute nothing).
    w4 = g("image_encoder.trunk.patch_embed.proj.weight")  # [E,3,7,7]
    std = torch.tensor(STD, dtype=torch.float64)[None, :, None, None]
    w8 = _bf16(F.pad((w4 / (255.0 * std)).reshape(E, 147), (0, PE_K - 147)))  # K order (c, ky, kx) for the mean term
    # the uint8 operand is written with K order (ky, kx, c), each kernel row padded from 21 to 24 entries: a kernel row's
    # 21 taps are contiguous bytes of the HWC image and land as three aligned 16-byte stores (k_im2col_u8raw)
    out["pe.w8"] = F.pad(w8[:, :147].reshape(E, 3, 7, 7).permute(0, 2, 3, 1).reshape(E, 7, 21), (0, 3)).reshape(E, PE_K8).contiguous()
    # the mean term uses the SAME bf16-rounded weights, so the sum equals round(w') . (p - 255 mean): the rounding error
    # stays relative to the normalised pixel value instead of to the raw 0..255 value
    mean_img = (255.0 * torch.tensor(MEAN, dtype=torch.float64))[None, :, None, None].expand(1, 3, 1024, 1024)
    mterm = F.conv2d(mean_img, w8.double()[:, :147].reshape(E, 3, 7, 7), None, stride=4, padding=3)[0]  # [E,256,256]
    pos8 = pos[0].double() + g("image_encoder.trunk.patch_embed.proj.bias")[:, None, None] - mterm
    out["pos8"] = _f32(pos8.permute(1, 2, 0).reshape(65536, E))
    # ---- trunk blocks
    hd = E // variant["heads"]          # real head dim (96 / 56 / 72)
    D = padded_head_dim(variant)        # head dim of the attention buffers (64 or 96): extra columns are exact zeros
    for i, (ci, co, nh, _ws, _pool) in enumerate(block_plan(variant)):
        b, o = f"image_encoder.trunk.blocks.{i}.", f"b{i}."
        out[o + "n1.g"], out[o + "n1.b"] = _f32(g(b + "norm1.weight")), _f32(g(b + "norm1.bias"))
        wq, bq, wp = g(b + "attn.qkv.weight"), g(b + "attn.qkv.bias"), g(b + "attn.proj.weight")
        if D != hd:
            # rows (part, head, hd) -> (part, head, D) with zero rows / bias; proj columns (head, hd) -> (head, D)
            wq = F.pad(wq.reshape(3, nh, hd, ci), (0, 0, 0, D - hd)).reshape(3 * nh * D, ci)
            bq = F.pad(bq.reshape(3, nh, hd), (0, D - hd)).reshape(3 * nh * D)
            wp = F.pad(wp.reshape(co, nh, hd), (0, D - hd)).reshape(co, nh * D)
        out[o + "qkv.w"], out[o + "qkv.b"] = _bf16(wq), _f32(bq)
        out[o + "proj.w"], out[o + "proj.b"] = _bf16(wp), _f32(g(b + "attn.proj.bias"))
        out[o + "n2.g"], out[o + "n2.b"] = _f32(g(b + "norm2.weight")), _f32(g(b + "norm2.bias"))
        out[o + "fc1.w"], out[o + "fc1.b"] = _bf16(g(b + "mlp.layers.0.weight")), _f32(g(b + "mlp.layers.0.bias"))
        out[o + "fc2.w"], out[o + "fc2.b"] = _bf16(g(b + "mlp.layers.1.weight")), _f32(g(b + "mlp.layers.1.bias"))
        if ci != co:
            out[o + "sc.w"], out[o + "sc.b"] = _bf16(g(b + "proj.weight")), _f32(g(b + "proj.bias"))
    # ---- neck (convs[0] = deepest level) and conv_s0/conv_s1 composed with their lateral 1x1 convs
    nw = [g(f"image_encoder.neck.convs.{j}.conv.weight").flatten(1) for j in range(4)]
    nb = [g(f"image_encoder.neck.convs.{j}.conv.bias") for j in range(4)]
    out["neck3.w"], out["neck3.b"] = _bf16(nw[0]), _f32(nb[0])
    out["neck2.w"], out["neck2.b"] = _bf16(nw[1]), _f32(nb[1])
    D = "sam_mask_decoder."
    ws1, bs1 = g(D + "conv_s1.weight").flatten(1), g(D + "conv_s1.bias")
    ws0, bs0 = g(D + "conv_s0.weight").flatten(1), g(D + "conv_s0.bias")
    out["s1.w"], out["s1.b"] = _bf16(ws1 @ nw[2]), _f32(ws1 @ nb[2] + bs1)
    out["s0.w"], out["s0.b"] = _bf16(ws0 @ nw[3]), _f32(ws0 @ nb[3] + bs0)
    # ---- learned prompts (sam2_infer.py:207-209,250) and the dense positional encoding
    de1, de2 = sd["dense_embedding1"].detach().cpu().double(), sd["dense_embedding2"].detach().cpu().double()
    out["dense"] = _f32((de1 @ de2)[0].t())  # [4096, 256]
    tokens = torch.cat([g(D + "obj_score_token.weight"), g(D + "iou_token.weight"), g(D + "mask_tokens.weight"),
                        sd["sparse_embedding"].detach().cpu().double()[0]], 0)  # [38, 256]
    out["tok0"] = _f32(tokens)
    kpe = dense_pe(sd["sam2_model.sam_prompt_encoder.pe_layer.positional_encoding_gaussian_matrix"].detach().cpu()).double()

    def lin(x, name):
        return x @ g(name + ".weight").t() + g(name + ".bias")

    def t2i(prefix, o):  # tokens -> image cross attention: K|V from one GEMM on the image tokens
        wk, wv = g(prefix + ".k_proj.weight"), g(prefix + ".v_proj.weight")
        out[o + ".kv.w"] = _bf16(torch.cat([wk, wv], 0))
        out[o + ".kv.b"] = _f32(torch.cat([g(prefix + ".k_proj.bias"), g(prefix + ".v_proj.bias")]))
        out[o + ".kpe"] = _f32(torch.cat([kpe @ wk.t(), torch.zeros(4096, wv.shape[0], dtype=torch.float64)], 1))
        out[o + ".o.w"], out[o + ".o.b"] = _f32(g(prefix + ".out_proj.weight")), _f32(g(prefix + ".out_proj.bias"))

    for l in range(2):
        t, o = f"{D}transformer.layers.{l}.", f"l{l}"
        if l == 0:
            # skip_first_layer_pe: queries = norm1(self_attn(tokens, tokens, tokens)) — input independent
            sa = t + "self_attn"
            a = _attn_tokens(lin(tokens, sa + ".q_proj"), lin(tokens, sa + ".k_proj"), lin(tokens, sa + ".v_proj"), 8)
            q1 = F.layer_norm(lin(a, sa + ".out_proj"), (256,), g(t + "norm1.weight"), g(t + "norm1.bias"), 1e-5)
            out["l0.q1"] = _f32(q1)
            out["l0.t2i.qc"] = _f32(lin(q1 + tokens, t + "cross_attn_token_to_image.q_proj"))
        else:
            for n, s in dict(q="q_proj", k="k_proj", v="v_proj", o="out_proj").items():
                out[f"{o}.sa.{n}.w"], out[f"{o}.sa.{n}.b"] = _f32(g(f"{t}self_attn.{s}.weight")), _f32(g(f"{t}self_attn.{s}.bias"))
            out[o + ".n1.g"], out[o + ".n1.b"] = _f32(g(t + "norm1.weight")), _f32(g(t + "norm1.bias"))
            out[o + ".t2i.q.w"] = _f32(g(t + "cross_attn_token_to_image.q_proj.weight"))
            out[o + ".t2i.q.b"] = _f32(g(t + "cross_attn_token_to_image.q_proj.bias"))
        t2i(t + "cross_attn_token_to_image", o + ".t2i")
        for n in (2, 3, 4):
            out[f"{o}.n{n}.g"], out[f"{o}.n{n}.b"] = _f32(g(f"{t}norm{n}.weight")), _f32(g(f"{t}norm{n}.bias"))
        out[o + ".mlp1.w"], out[o + ".mlp1.b"] = _f32(g(t + "mlp.layers.0.weight")), _f32(g(t + "mlp.layers.0.bias"))
        out[o + ".mlp2.w"], out[o + ".mlp2.b"] = _f32(g(t + "mlp.layers.1.weight")), _f32(g(t + "mlp.layers.1.bias"))
        c = t + "cross_attn_image_to_token"
        wq = g(c + ".q_proj.weight")
        out[o + ".i2t.q.w"], out[o + ".i2t.q.b"] = _bf16(wq), _f32(g(c + ".q_proj.bias"))
        out[o + ".i2t.qpe"] = _f32(kpe @ wq.t())
        out[o + ".i2t.k.w"], out[o + ".i2t.k.b"] = _f32(g(c + ".k_proj.weight")), _f32(g(c + ".k_proj.bias"))
        out[o + ".i2t.v.w"], out[o + ".i2t.v.b"] = _f32(g(c + ".v_proj.weight")), _f32(g(c + ".v_proj.bias"))
        out[o + ".i2t.o.w"], out[o + ".i2t.o.b"] = _bf16(g(c + ".out_proj.weight")), _f32(g(c + ".out_proj.bias"))
    f = D + "transformer.final_attn_token_to_image"
    out["fin.q.w"], out["fin.q.b"] = _f32(g(f + ".q_proj.weight")), _f32(g(f + ".q_proj.bias"))
    t2i(f, "fin")
    out["fin.n.g"], out["fin.n.b"] = _f32(g(D + "transformer.norm_final_attn.weight")), _f32(g(D + "transformer.norm_final_attn.bias"))
    for j in range(3):
        out[f"iou.{j}.w"] = _f32(g(f"{D}iou_prediction_head.layers.{j}.weight"))
        out[f"iou.{j}.b"] = _f32(g(f"{D}iou_prediction_head.layers.{j}.bias"))
        for k in range(4):
            out[f"hyp{k}.{j}.w"] = _f32(g(f"{D}output_hypernetworks_mlps.{k}.layers.{j}.weight"))
            out[f"hyp{k}.{j}.b"] = _f32(g(f"{D}output_hypernetworks_mlps.{k}.layers.{j}.bias"))
    # ---- upscaling: ConvTranspose2d(k2,s2) weight [Cin, Cout, dy, dx] -> GEMM operand [(dy*2+dx)*Cout + co, Cin]
    u1, u2 = g(D + "output_upscaling.0.weight"), g(D + "output_upscaling.3.weight")
    out["up1.w"], out["up1.b"] = _bf16(u1.permute(2, 3, 1, 0).reshape(4 * 64, 256)), _f32(g(D + "output_upscaling.0.bias"))
    out["up2.w"], out["up2.b"] = _bf16(u2.permute(2, 3, 1, 0).reshape(4 * 32, 64)), _f32(g(D + "output_upscaling.3.bias"))
    out["upln.g"], out["upln.b"] = _f32(g(D + "output_upscaling.1.weight")), _f32(g(D + "output_upscaling.1.bias"))
    # ---- refinement head (sam2_infer.py:130-189)
    if use_refinement:
        for j, k in enumerate(REFINE_KERNELS):
            wj = sd[f"refinement_layer.conv_branches.{j}.weight"].detach().cpu()
            if tuple(wj.shape) != (4, 1, k, k):
                raise ValueError(f"refinement branch {j}: weight {tuple(wj.shape)}; the fused tail kernel is built for "
                                 f"kernel sizes {REFINE_KERNELS} with 4 channels per branch")
            out[f"ref{j}.w"] = _f32(wj.reshape(4, k, k))
            out[f"ref{j}.b"] = _f32(sd[f"refinement_layer.conv_branches.{j}.bias"].detach().cpu())
        out["ref.comp"] = _f32(refine_phase_tables([out[f"ref{j}.w"] for j in range(4)]))
        out["refc.w"] = _f32(sd["refinement_layer.combiner_conv.weight"].detach().cpu().reshape(16))
        out["refc.b"] = _f32(sd["refinement_layer.combiner_conv.bias"].detach().cpu().reshape(1))
    return out
