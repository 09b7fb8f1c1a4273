"""CPU error-budget study: which bf16 roundings of the SAM 2.1 path cost mask IoU (oracle with selective rounding)."""
import sys, os, copy
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, torch.nn as nn
from oracle import sam2_oracle
from circuitvision_b200 import synth

torch.set_num_threads(8)
ref = sam2_oracle.build_oracle("tiny", seed=0)
seeds = [21, 5]
xs = torch.stack([sam2_oracle.preprocess_rgb(synth.make_schematic(s, 1024, render_rgb=True)[2]) for s in seeds])
with torch.no_grad():
    rh, rl, ri = ref(xs)

def bf(t): return t.to(torch.bfloat16).float()
def fp16(t, s=1.0): return (t * s).to(torch.float16).float() / s

def iou(a, b):
    a, b = a > 0, b > 0
    return [((a[i] & b[i]).sum().item() / max((a[i] | b[i]).sum().item(), 1)) for i in range(a.shape[0])]

def run(name, mod_fn):
    m = copy.deepcopy(ref)
    hooks = mod_fn(m) or []
    with torch.no_grad():
        h, l, i = m(xs)
    d = (l - rl).abs()
    print(f"{name:50s} IoU {['%.4f' % v for v in iou(h, rh)]}  low max|err|/std {d.max().item()/rl.std().item():.4f} mean {d.mean().item()/rl.std().item():.5f}")
    for hk in hooks: hk.remove()

def pe_bf16(m):
    w = m.sam2_model.image_encoder.trunk.patch_embed.proj.weight
    w.data = bf(w.data)
def pe_fp16(m):
    w = m.sam2_model.image_encoder.trunk.patch_embed.proj.weight
    w.data = fp16(w.data, 4.0)
def trunk_w(m):
    for mod in m.sam2_model.image_encoder.trunk.blocks.modules():
        if isinstance(mod, nn.Linear): mod.weight.data = bf(mod.weight.data)
def trunk_act(m):
    hs = []
    for mod in m.sam2_model.image_encoder.trunk.blocks.modules():
        if isinstance(mod, nn.Linear):
            hs.append(mod.register_forward_pre_hook(lambda mod, inp: (bf(inp[0]),)))
    return hs
def trunk_qkv_out(m):
    hs = []
    for blk in m.sam2_model.image_encoder.trunk.blocks:
        hs.append(blk.attn.qkv.register_forward_hook(lambda mod, inp, out: bf(out)))
    return hs
def neck_dec_w(m):
    for mod in list(m.sam2_model.image_encoder.neck.modules()) + list(m.sam2_model.sam_mask_decoder.modules()):
        if isinstance(mod, (nn.Conv2d, nn.ConvTranspose2d)): mod.weight.data = bf(mod.weight.data)
def all_of(*fns):
    def f(m):
        hs = []
        for fn in fns: hs += fn(m) or []
        return hs
    return f
def f16(t): return t.to(torch.float16).float()
def trunk_w16(m):
    for mod in m.sam2_model.image_encoder.trunk.blocks.modules():
        if isinstance(mod, nn.Linear): mod.weight.data = f16(mod.weight.data)
def trunk_act16(m):
    hs = []
    for mod in m.sam2_model.image_encoder.trunk.blocks.modules():
        if isinstance(mod, nn.Linear):
            hs.append(mod.register_forward_pre_hook(lambda mod, inp: (f16(inp[0]),)))
    for blk in m.sam2_model.image_encoder.trunk.blocks:
        hs.append(blk.attn.qkv.register_forward_hook(lambda mod, inp, out: f16(out)))
    return hs
def neck_dec_act(m):
    hs = []
    for mod in list(m.sam2_model.image_encoder.neck.modules()) + list(m.sam2_model.sam_mask_decoder.modules()):
        if isinstance(mod, (nn.Conv2d, nn.ConvTranspose2d)):
            hs.append(mod.register_forward_pre_hook(lambda mod, inp: (bf(inp[0]),)))
    return hs
def dec_tok16(m):
    """token side of the decoder (two-way transformer linears, hyper / iou MLPs) with fp16 weights and inputs"""
    hs = []
    for mod in m.sam2_model.sam_mask_decoder.modules():
        if isinstance(mod, nn.Linear):
            mod.weight.data = f16(mod.weight.data)
            hs.append(mod.register_forward_pre_hook(lambda mod, inp: (f16(inp[0]),)))
    return hs
def dec_tok_bf(m):
    hs = []
    for mod in m.sam2_model.sam_mask_decoder.modules():
        if isinstance(mod, nn.Linear):
            mod.weight.data = bf(mod.weight.data)
            hs.append(mod.register_forward_pre_hook(lambda mod, inp: (bf(inp[0]),)))
    return hs
def _dec_sel(pred, cast):
    def f(m):
        hs = []
        for name, mod in m.sam2_model.sam_mask_decoder.named_modules():
            if isinstance(mod, nn.Linear) and pred(name):
                mod.weight.data = cast(mod.weight.data)
                hs.append(mod.register_forward_pre_hook(lambda mod, inp: (cast(inp[0]),)))
        return hs
    return f
_img_side = lambda n: any(k in n for k in ("token_to_image.k_proj", "token_to_image.v_proj", "image_to_token.q_proj", "image_to_token.out_proj"))
which = sys.argv[1:] or ["pe_bf16", "pe_fp16", "trunk_w", "trunk_act", "trunk_qkv_out", "neck_dec_w", "all"]
table = dict(pe_bf16=pe_bf16, pe_fp16=pe_fp16, trunk_w=trunk_w, trunk_act=trunk_act, trunk_qkv_out=trunk_qkv_out, neck_dec_w=neck_dec_w,
             trunk_bf16=all_of(trunk_w, trunk_act, trunk_qkv_out), trunk_fp16=all_of(trunk_w16, trunk_act16), neck_dec_act=neck_dec_act,
             trunk_bf16_dec_act=all_of(trunk_w, trunk_act, trunk_qkv_out, neck_dec_act),
             dec_img16=_dec_sel(_img_side, f16), dec_img_bf=_dec_sel(_img_side, bf),
             dec_hyper16=_dec_sel(lambda n: "hypernetworks" in n or "iou_prediction" in n, f16),
             dec_tokside16=_dec_sel(lambda n: "transformer" in n and not _img_side(n), f16),
             dec_tok16=dec_tok16, dec_tok_bf=dec_tok_bf,
             all=all_of(pe_bf16, trunk_w, trunk_act, trunk_qkv_out, neck_dec_w), all_fp16pe=all_of(pe_fp16, trunk_w, trunk_act, trunk_qkv_out, neck_dec_w))
for k in which: run(k, table[k])
