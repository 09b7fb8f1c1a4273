"""GPU diagnostic: per-tap error of the CUDA SAM 2.1 path against the fp32 oracle (not a test; run with gpurun)."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from oracle import sam2_oracle
from circuitvision_b200 import sam2_infer, synth

variant = sys.argv[1] if len(sys.argv) > 1 else "tiny"
B = int(sys.argv[2]) if len(sys.argv) > 2 else 2
ref = sam2_oracle.build_oracle(variant, seed=0)
model = sam2_infer.get_modified_sam2(variant, None, device="cuda:0", use_refinement_layer=True)
print(model.load_state_dict(ref.state_dict()))
xs = torch.stack([sam2_oracle.preprocess_rgb(synth.make_schematic(5 + i, 1024, render_rgb=True)[2]) for i in range(B)])
t = time.time()
with torch.no_grad():
    rh, rl, ri, aux = ref(xs, return_aux=True)
print("oracle s", time.time() - t)
model.set_max_batch(B)
if len(sys.argv) > 3:
    model.set_operand_dtype(torch.bfloat16 if sys.argv[3] == "bf16" else torch.float16)
high, low, iou = model(xs.cuda())
torch.cuda.synchronize()
eng = model.engine()
E = model.sam2_model.variant["embed"]


def rep(name, got, want):
    got, want = got.float().cpu(), want.float()
    d = (got - want).abs()
    print(f"{name:10s} max|err| {d.max().item():.3e}  rel-to-std {d.max().item() / (want.std().item() + 1e-12):.3e}  "
          f"mean|err|/std {d.mean().item() / (want.std().item() + 1e-12):.3e}  std {want.std().item():.3e}  nan {torch.isnan(got).any().item()}")


for s in range(4):
    hw = 256 >> s
    C = E << s
    got = eng.read_buffer(f"X{s}", (B, hw, hw, C))
    rep(f"trunk{s}", got, aux["trunk"][s].permute(0, 2, 3, 1))
rep("s0", eng.read_buffer("s0", (B, 256, 256, 32)), aux["s0"].permute(0, 2, 3, 1))
rep("s1", eng.read_buffer("s1", (B, 128, 128, 64)), aux["s1"].permute(0, 2, 3, 1))
rep("masks", eng.read_buffer("masks", (B, 4, 256, 256)), aux["all_masks"])
rep("iou4", eng.read_buffer("iou4", (B, 4)), aux["all_iou"])
print("sel", eng.read_buffer("sel", (B,), torch.int32).cpu().tolist(), "oracle stable", aux["stable"].tolist(), "best", (aux["best"] + 1).tolist())
rep("low", low, rl)
rep("iou", iou, ri)
rep("high", high, rh)
a, b = high.cpu() > 0, rh > 0
for i in range(B):
    inter, union = (a[i] & b[i]).sum().item(), (a[i] | b[i]).sum().item()
    print(f"image {i}: fg {b[i].float().mean().item():.3f}  IoU {inter / max(union, 1):.5f}")
print("launches", eng.launches)
# timing
for bb in (B,):
    x = xs.cuda()
    for _ in range(2):
        model(x)
    torch.cuda.synchronize()
    t = time.time()
    for _ in range(5):
        model(x)
    torch.cuda.synchronize()
    print(f"B={bb}: {(time.time() - t) / 5 * 1e3:.2f} ms per forward")
