"""TEST INFRASTRUCTURE ONLY — fp32 CPU restatement of the reference's SAM 2.1 path.

Follows `/root/reference/src/sam2_infer.py`:
    SAM2Transforms            :29-128   (ToTensor -> Resize((R,R)) -> Normalize ; postprocess_masks = bilinear resize)
    MultiKernelRefinement     :130-189
    SAM2ImageWrapper.forward  :191-275  (encoder -> conv_s0/conv_s1 -> mask decoder with learned prompts ->
                                         bilinear x4 -> refinement)
and, for the arithmetic that lives in the un-vendored, un-pinned third-party `sam2` package
(`requirements.txt:12`: git+https://github.com/facebookresearch/sam2.git, no tag), the published SAM 2.1
architecture: Hiera trunk + FPN neck (hyper-parameters from `/root/reference/models/configs/sam2.1_hiera_l.yaml:6-28`
for large; SURVEY.md §B.1 for tiny/small/base+), prompt-encoder dense PE, two-way-transformer mask decoder with
`dynamic_multimask_via_stability`.  Parameter names mirror the upstream module tree
(`sam2_model.image_encoder.trunk.blocks.N.attn.qkv.weight`, ...) so that a reference checkpoint maps 1:1.

PARITY UNPINNED by reference tests: the reference ships no test, golden vector or fixture for this path and `sam2`
cannot be installed here (no network).  The restatement is cross-checked in tests/test_sam2_oracle_cpu.py against the
independent implementation of the same architecture in the image's `transformers` (modeling_sam2.py) with
identical weights — evidence, not ground truth.

Only tests/, __graft_entry__.smoke() and bench.py's CPU-baseline legs may import this module.
"""
from __future__ import annotations

import math

import numpy as np
import torch
import torch.nn as nn
import torch.nn.functional as F

VARIANTS = {
    # embed, heads, stages, global blocks, window spec, pos-embed background
    "tiny": dict(embed=96, heads=1, stages=(1, 2, 7, 2), global_blocks=(5, 7, 9), window_spec=(8, 4, 14, 7), bkg=(7, 7)),
    "small": dict(embed=96, heads=1, stages=(1, 2, 11, 2), global_blocks=(7, 10, 13), window_spec=(8, 4, 14, 7), bkg=(7, 7)),
    "base_plus": dict(embed=112, heads=2, stages=(2, 3, 16, 3), global_blocks=(12, 16, 20), window_spec=(8, 4, 14, 7), bkg=(14, 14)),
    "large": dict(embed=144, heads=2, stages=(2, 6, 36, 4), global_blocks=(23, 33, 43), window_spec=(8, 4, 16, 8), bkg=(7, 7)),
}
IMAGE_SIZE = 1024
MEAN = (0.485, 0.456, 0.406)
STD = (0.229, 0.224, 0.225)


def block_plan(variant: str):
    """[(dim_in, dim_out, heads, window, q_pool)] for every trunk block."""
    v = VARIANTS[variant]
    plan, total = [], 0
    for s, nb in enumerate(v["stages"]):
        for b in range(nb):
            first = s > 0 and b == 0
            dim_out = v["embed"] * 2 ** s
            dim_in = v["embed"] * 2 ** (s - 1) if first else dim_out
            ws = v["window_spec"][s - 1] if first else v["window_spec"][s]
            if total in v["global_blocks"]:
                ws = 0
            plan.append((dim_in, dim_out, v["heads"] * 2 ** s, ws, first and s <= 3))
            total += 1
    return plan


# ------------------------------------------------------------------------------------------ modules
class MLP(nn.Module):
    def __init__(self, d_in, d_hidden, d_out, n_layers, act, sigmoid=False):
        super().__init__()
        dims = [d_in] + [d_hidden] * (n_layers - 1) + [d_out]
        self.layers = nn.ModuleList(nn.Linear(a, b) for a, b in zip(dims[:-1], dims[1:]))
        self.act, self.sigmoid = act, sigmoid

    def forward(self, x):
        for i, l in enumerate(self.layers):
            x = l(x)
            if i + 1 < len(self.layers):
                x = self.act(x)
        return torch.sigmoid(x) if self.sigmoid else x


class MultiScaleAttention(nn.Module):
    def __init__(self, dim, dim_out, heads, q_pool):
        super().__init__()
        self.heads, self.q_pool = heads, q_pool
        self.qkv = nn.Linear(dim, dim_out * 3)
        self.proj = nn.Linear(dim_out, dim_out)

    def forward(self, x):  # x: (Bw, H, W, C)
        B, H, W, _ = x.shape
        qkv = self.qkv(x).reshape(B, H * W, 3, self.heads, -1)
        q, k, v = torch.unbind(qkv, 2)
        if self.q_pool:
            q = q.reshape(B, H, W, -1).permute(0, 3, 1, 2)
            q = F.max_pool2d(q, 2, 2).permute(0, 2, 3, 1)
            H, W = q.shape[1:3]
            q = q.reshape(B, H * W, self.heads, -1)
        q, k, v = (t.transpose(1, 2) for t in (q, k, v))
        d = q.shape[-1]
        a = torch.softmax((q @ k.transpose(-1, -2)) * d ** -0.5, dim=-1)
        x = (a @ v).transpose(1, 2).reshape(B, H, W, -1)
        return self.proj(x)


def window_partition(x, ws):
    B, H, W, C = x.shape
    ph, pw = (ws - H % ws) % ws, (ws - W % ws) % ws
    x = F.pad(x, (0, 0, 0, pw, 0, ph))
    Hp, Wp = H + ph, W + pw
    x = x.view(B, Hp // ws, ws, Wp // ws, ws, C).permute(0, 1, 3, 2, 4, 5).reshape(-1, ws, ws, C)
    return x, (Hp, Wp)


def window_unpartition(w, ws, pad_hw, hw):
    Hp, Wp = pad_hw
    H, W = hw
    B = w.shape[0] // (Hp * Wp // ws // ws)
    x = w.view(B, Hp // ws, Wp // ws, ws, ws, -1).permute(0, 1, 3, 2, 4, 5).reshape(B, Hp, Wp, -1)
    return x[:, :H, :W, :]


class MultiScaleBlock(nn.Module):
    def __init__(self, dim, dim_out, heads, window, q_pool):
        super().__init__()
        self.dim, self.dim_out, self.window, self.q_pool = dim, dim_out, window, q_pool
        self.norm1 = nn.LayerNorm(dim, eps=1e-6)
        self.attn = MultiScaleAttention(dim, dim_out, heads, q_pool)
        self.norm2 = nn.LayerNorm(dim_out, eps=1e-6)
        self.mlp = MLP(dim_out, dim_out * 4, dim_out, 2, nn.GELU())
        if dim != dim_out:
            self.proj = nn.Linear(dim, dim_out)

    def forward(self, x):  # (B, H, W, C)
        shortcut = x
        x = self.norm1(x)
        if self.dim != self.dim_out:
            s = self.proj(x).permute(0, 3, 1, 2)
            shortcut = F.max_pool2d(s, 2, 2).permute(0, 2, 3, 1)
        ws = self.window
        if ws > 0:
            H, W = x.shape[1:3]
            x, pad_hw = window_partition(x, ws)  # zero padding AFTER norm1: pad tokens get qkv = bias, unmasked
        x = self.attn(x)
        if self.q_pool:
            ws = self.window // 2
            H, W = shortcut.shape[1:3]
            pad_hw = (H + (ws - H % ws) % ws, W + (ws - W % ws) % ws)
        if self.window > 0:
            x = window_unpartition(x, ws, pad_hw, (H, W))
        x = shortcut + x
        return x + self.mlp(self.norm2(x))


class PatchEmbed(nn.Module):
    def __init__(self, dim):
        super().__init__()
        self.proj = nn.Conv2d(3, dim, kernel_size=7, stride=4, padding=3)

    def forward(self, x):
        return self.proj(x).permute(0, 2, 3, 1)


class Hiera(nn.Module):
    def __init__(self, variant):
        super().__init__()
        v = VARIANTS[variant]
        self.patch_embed = PatchEmbed(v["embed"])
        self.pos_embed = nn.Parameter(torch.zeros(1, v["embed"], *v["bkg"]))
        self.pos_embed_window = nn.Parameter(torch.zeros(1, v["embed"], v["window_spec"][0], v["window_spec"][0]))
        self.blocks = nn.ModuleList(MultiScaleBlock(*b) for b in block_plan(variant))
        self.stage_ends = list(np.cumsum(v["stages"]) - 1)

    def pos(self, hw):
        pe = F.interpolate(self.pos_embed, size=hw, mode="bicubic")
        pe = pe + self.pos_embed_window.tile([x // y for x, y in zip(pe.shape, self.pos_embed_window.shape)])
        return pe.permute(0, 2, 3, 1)

    def forward(self, x):
        x = self.patch_embed(x)
        x = x + self.pos(x.shape[1:3])
        outs = []
        for i, blk in enumerate(self.blocks):
            x = blk(x)
            if i in self.stage_ends:
                outs.append(x.permute(0, 3, 1, 2))
        return outs


class NeckConv(nn.Module):
    def __init__(self, c_in, d):
        super().__init__()
        self.conv = nn.Conv2d(c_in, d, 1)


class FpnNeck(nn.Module):
    """convs[0] takes the deepest (widest) level; top-down nearest x2 add only on level 2 (fpn_top_down_levels [2,3])."""

    def __init__(self, channel_list, d=256):
        super().__init__()
        self.convs = nn.ModuleList(NeckConv(c, d) for c in channel_list)

    def forward(self, xs):
        n = len(self.convs) - 1
        out = [None] * len(self.convs)
        prev = None
        for i in range(n, -1, -1):
            lat = self.convs[n - i].conv(xs[i])
            if i in (2, 3) and prev is not None:
                prev = lat + F.interpolate(prev, scale_factor=2.0, mode="nearest")
            else:
                prev = lat
            out[i] = prev
        return out[:-1]  # scalp = 1: drop the 32x32 level


class ImageEncoder(nn.Module):
    def __init__(self, variant):
        super().__init__()
        e = VARIANTS[variant]["embed"]
        self.trunk = Hiera(variant)
        self.neck = FpnNeck([e * 8, e * 4, e * 2, e])

    def forward(self, x):
        return self.neck(self.trunk(x))


class Attention(nn.Module):
    def __init__(self, dim, heads, downsample=1):
        super().__init__()
        self.heads = heads
        inner = dim // downsample
        self.q_proj, self.k_proj, self.v_proj = nn.Linear(dim, inner), nn.Linear(dim, inner), nn.Linear(dim, inner)
        self.out_proj = nn.Linear(inner, dim)

    def forward(self, q, k, v):
        B = max(q.shape[0], k.shape[0])
        sp = lambda t: t.expand(B, -1, -1).reshape(B, t.shape[1], self.heads, -1).transpose(1, 2)
        q, k, v = sp(self.q_proj(q)), sp(self.k_proj(k)), sp(self.v_proj(v))
        a = torch.softmax((q @ k.transpose(-1, -2)) / math.sqrt(q.shape[-1]), dim=-1)
        o = (a @ v).transpose(1, 2).reshape(B, -1, q.shape[-1] * self.heads)
        return self.out_proj(o)


class TwoWayBlock(nn.Module):
    def __init__(self, dim, heads, mlp_dim, skip_first_layer_pe):
        super().__init__()
        self.self_attn = Attention(dim, heads)
        self.norm1 = nn.LayerNorm(dim)
        self.cross_attn_token_to_image = Attention(dim, heads, 2)
        self.norm2 = nn.LayerNorm(dim)
        self.mlp = MLP(dim, mlp_dim, dim, 2, nn.ReLU())
        self.norm3 = nn.LayerNorm(dim)
        self.norm4 = nn.LayerNorm(dim)
        self.cross_attn_image_to_token = Attention(dim, heads, 2)
        self.skip_first_layer_pe = skip_first_layer_pe

    def forward(self, queries, keys, query_pe, key_pe):
        if self.skip_first_layer_pe:
            queries = self.self_attn(queries, queries, queries)
        else:
            q = queries + query_pe
            queries = queries + self.self_attn(q, q, queries)
        queries = self.norm1(queries)
        q, k = queries + query_pe, keys + key_pe
        queries = self.norm2(queries + self.cross_attn_token_to_image(q, k, keys))
        queries = self.norm3(queries + self.mlp(queries))
        q, k = queries + query_pe, keys + key_pe
        keys = self.norm4(keys + self.cross_attn_image_to_token(k, q, queries))
        return queries, keys


class TwoWayTransformer(nn.Module):
    def __init__(self, dim=256, heads=8, mlp_dim=2048, depth=2):
        super().__init__()
        self.layers = nn.ModuleList(TwoWayBlock(dim, heads, mlp_dim, i == 0) for i in range(depth))
        self.final_attn_token_to_image = Attention(dim, heads, 2)
        self.norm_final_attn = nn.LayerNorm(dim)

    def forward(self, image_embedding, image_pe, point_embedding):
        B, C, H, W = image_embedding.shape
        keys = image_embedding.flatten(2).permute(0, 2, 1)
        key_pe = image_pe.flatten(2).permute(0, 2, 1)
        queries = point_embedding
        for layer in self.layers:
            queries, keys = layer(queries, keys, point_embedding, key_pe)
        q, k = queries + point_embedding, keys + key_pe
        queries = self.norm_final_attn(queries + self.final_attn_token_to_image(q, k, keys))
        return queries, keys


class LayerNorm2d(nn.Module):
    def __init__(self, c, eps=1e-6):
        super().__init__()
        self.weight, self.bias, self.eps = nn.Parameter(torch.ones(c)), nn.Parameter(torch.zeros(c)), eps

    def forward(self, x):
        u = x.mean(1, keepdim=True)
        s = (x - u).pow(2).mean(1, keepdim=True)
        x = (x - u) / torch.sqrt(s + self.eps)
        return self.weight[:, None, None] * x + self.bias[:, None, None]


class MaskDecoder(nn.Module):
    def __init__(self, dim=256):
        super().__init__()
        self.transformer = TwoWayTransformer(dim)
        self.iou_token = nn.Embedding(1, dim)
        self.mask_tokens = nn.Embedding(4, dim)
        self.obj_score_token = nn.Embedding(1, dim)
        self.output_upscaling = nn.Sequential(
            nn.ConvTranspose2d(dim, dim // 4, 2, 2), LayerNorm2d(dim // 4), nn.GELU(),
            nn.ConvTranspose2d(dim // 4, dim // 8, 2, 2), nn.GELU())
        self.conv_s0 = nn.Conv2d(dim, dim // 8, 1)
        self.conv_s1 = nn.Conv2d(dim, dim // 4, 1)
        self.output_hypernetworks_mlps = nn.ModuleList(MLP(dim, dim, dim // 8, 3, nn.ReLU()) for _ in range(4))
        self.iou_prediction_head = MLP(dim, 256, 4, 3, nn.ReLU(), sigmoid=True)
        self.pred_obj_score_head = MLP(dim, dim, 1, 3, nn.ReLU())

    def forward(self, image_embeddings, image_pe, sparse, dense, high_res_features):
        """multimask_output=False, repeat_image=True, eval mode (dynamic multimask via stability).  Batched ==
        stack of independent B=1 calls (the reference only ever runs B=1; SURVEY §7 hard part 5)."""
        B = image_embeddings.shape[0]
        tokens = torch.cat([self.obj_score_token.weight, self.iou_token.weight, self.mask_tokens.weight], 0)
        tokens = torch.cat([tokens[None].expand(sparse.shape[0], -1, -1), sparse], 1).expand(B, -1, -1)
        src = image_embeddings + dense
        hs, src = self.transformer(src, image_pe.expand(B, -1, -1, -1), tokens)
        iou_token_out, mask_tokens_out = hs[:, 1], hs[:, 2:6]
        src = src.transpose(1, 2).reshape(B, -1, *image_embeddings.shape[2:])
        dc1, ln1, act1, dc2, act2 = self.output_upscaling
        feat_s0, feat_s1 = high_res_features
        up = act1(ln1(dc1(src) + feat_s1))
        up = act2(dc2(up) + feat_s0)
        hyper = torch.stack([self.output_hypernetworks_mlps[i](mask_tokens_out[:, i]) for i in range(4)], 1)
        b, c, h, w = up.shape
        masks = (hyper @ up.view(b, c, h * w)).view(b, 4, h, w)
        iou = self.iou_prediction_head(iou_token_out)
        obj = self.pred_obj_score_head(hs[:, 0])
        # dynamic_multimask_via_stability (delta 0.05, thresh 0.98)
        flat = masks[:, 0].flatten(1)
        area_i = (flat > 0.05).sum(-1).float()
        area_u = (flat > -0.05).sum(-1).float()
        stability = torch.where(area_u > 0, area_i / area_u, torch.ones_like(area_u))
        best = torch.argmax(iou[:, 1:], dim=-1)
        ar = torch.arange(B)
        stable = stability >= 0.98
        out_masks = torch.where(stable[:, None, None], masks[:, 0], masks[ar, best + 1])[:, None]
        out_iou = torch.where(stable, iou[:, 0], iou[ar, best + 1])[:, None]
        aux = dict(all_masks=masks, all_iou=iou, stability=stability, best=best, stable=stable, obj=obj)
        return out_masks, out_iou, aux


class PositionEmbeddingRandom(nn.Module):
    def __init__(self, num_pos_feats=128, scale=1.0):
        super().__init__()
        self.register_buffer("positional_encoding_gaussian_matrix", scale * torch.randn((2, num_pos_feats)))

    def forward(self, size=(64, 64)):
        h, w = size
        grid = torch.ones((h, w), dtype=torch.float32)
        y = (grid.cumsum(0) - 0.5) / h
        x = (grid.cumsum(1) - 0.5) / w
        c = 2 * torch.stack([x, y], -1) - 1
        c = 2 * np.pi * (c @ self.positional_encoding_gaussian_matrix)
        return torch.cat([torch.sin(c), torch.cos(c)], -1).permute(2, 0, 1)[None]


class PromptEncoder(nn.Module):
    def __init__(self):
        super().__init__()
        self.pe_layer = PositionEmbeddingRandom(128)

    def get_dense_pe(self):
        return self.pe_layer((64, 64))


class SAM2Base(nn.Module):
    image_size = IMAGE_SIZE

    def __init__(self, variant):
        super().__init__()
        self.image_encoder = ImageEncoder(variant)
        self.sam_prompt_encoder = PromptEncoder()
        self.sam_mask_decoder = MaskDecoder()


class MultiKernelRefinement(nn.Module):
    """sam2_infer.py:130-189 (intermediate channels are hard-coded to 4 by the wrapper, :215)."""

    def __init__(self, kernel_sizes=(3, 5, 7, 11), ch=4):
        super().__init__()
        self.conv_branches = nn.ModuleList(nn.Conv2d(1, ch, k, padding="same") for k in kernel_sizes)
        self.activation = nn.GELU()
        self.combiner_conv = nn.Conv2d(len(kernel_sizes) * ch, 1, 1)

    def forward(self, x):
        return self.combiner_conv(torch.cat([self.activation(b(x)) for b in self.conv_branches], 1))


class SAM2ImageWrapperOracle(nn.Module):
    """sam2_infer.py:191-275."""

    def __init__(self, variant="tiny", embedding_r=4, use_refinement=True, refinement_kernel_sizes=(3, 5, 7, 11)):
        super().__init__()
        self.variant = variant
        self.sam2_model = SAM2Base(variant)
        self.dense_embedding1 = nn.Parameter(torch.randn(1, 256, embedding_r))
        self.dense_embedding2 = nn.Parameter(torch.randn(1, embedding_r, 64 * 64))
        self.sparse_embedding = nn.Parameter(torch.randn(1, 32, 256))
        self.refinement_layer = MultiKernelRefinement(refinement_kernel_sizes) if use_refinement else None

    def forward(self, images, return_aux=False):
        enc = self.sam2_model.image_encoder
        trunk_out = enc.trunk(images)
        fpn = enc.neck(trunk_out)
        dec = self.sam2_model.sam_mask_decoder
        s0, s1 = dec.conv_s0(fpn[0]), dec.conv_s1(fpn[1])
        dense = (self.dense_embedding1 @ self.dense_embedding2).view(1, 256, 64, 64)
        low, iou, aux = dec(fpn[2], self.sam2_model.sam_prompt_encoder.get_dense_pe(), self.sparse_embedding, dense, [s0, s1])
        high = F.interpolate(low, size=(IMAGE_SIZE, IMAGE_SIZE), mode="bilinear", align_corners=False)
        if self.refinement_layer is not None:
            high = self.refinement_layer(high)
        if return_aux:
            aux.update(fpn=fpn, s0=s0, s1=s1, trunk=trunk_out)
            return high, low, iou, aux
        return high, low, iou


# ------------------------------------------------------------------------------------------ init / helpers
REFINE_SEED = 1102


def build_oracle(variant="tiny", seed=0, use_refinement=True, refine_seed=REFINE_SEED):
    """Deterministic random init, 'upstream-style': PyTorch-default reset_parameters() on every Linear / Conv /
    ConvTranspose / LayerNorm / Embedding, trunc-normal(0.02) positional embeddings, randn wrapper prompts
    (SURVEY §7 hard part 4 — NOT the HF std-0.02 init, which yields degenerate logits).

    The refinement head is re-initialised (same PyTorch-default recipe) under its own seed: a default-init
    MultiKernelRefinement adds a random offset several times larger than the spread of its output, so for most
    seeds the thresholded mask is all-0 or all-1 and an IoU gate would be vacuous.  refine_seed=1102 was picked by
    scanning 300 seeds for a foreground fraction near 30 % on schematic inputs; both sides of every comparison
    use the same weights, so the choice only makes the test non-trivial."""
    torch.manual_seed(seed)
    m = SAM2ImageWrapperOracle(variant, use_refinement=use_refinement)
    t = m.sam2_model.image_encoder.trunk
    nn.init.trunc_normal_(t.pos_embed, std=0.02)
    nn.init.trunc_normal_(t.pos_embed_window, std=0.02)
    if m.refinement_layer is not None:
        torch.manual_seed(refine_seed)
        for mod in m.refinement_layer.modules():
            if isinstance(mod, nn.Conv2d):
                mod.reset_parameters()
    return m.eval()


def preprocess_rgb(rgb_u8: np.ndarray) -> torch.Tensor:
    """SAM2Transforms.__call__ (:49-51) for an (H,W,3) uint8 array: ToTensor, Resize((1024,1024)) (bilinear,
    antialias; identity at 1024²), Normalize."""
    x = torch.from_numpy(np.ascontiguousarray(rgb_u8)).permute(2, 0, 1).float() / 255.0
    if x.shape[1:] != (IMAGE_SIZE, IMAGE_SIZE):
        x = F.interpolate(x[None], size=(IMAGE_SIZE, IMAGE_SIZE), mode="bilinear", align_corners=False, antialias=True)[0]
    mean = torch.tensor(MEAN)[:, None, None]
    std = torch.tensor(STD)[:, None, None]
    return (x - mean) / std


def postprocess_masks(masks: torch.Tensor, orig_hw) -> torch.Tensor:
    """SAM2Transforms.postprocess_masks (:88-128) with hole / sprinkle areas 0 (circuit_analyzer.py:245-250)."""
    return F.interpolate(masks.float(), orig_hw, mode="bilinear", align_corners=False)


def segment(model, image_np_bgr: np.ndarray):
    """circuit_analyzer.py:343-370: channel swap, transform, forward, postprocess, threshold, extent bbox."""
    rgb = image_np_bgr[:, :, ::-1]
    x = preprocess_rgb(rgb)[None]
    with torch.no_grad():
        high, low, iou = model(x)
    logits = postprocess_masks(high, rgb.shape[:2])[0, 0]
    mask = (logits > 0.0).numpy().astype(np.uint8) * 255
    ys, xs = np.nonzero(mask)
    bbox = (int(xs.min()), int(ys.min()), int(xs.max()) + 1, int(ys.max()) + 1) if len(xs) else None
    return mask, logits.numpy(), bbox
