"""Per-kernel CUDA-event table of one SAM 2.1 forward (cv_profile_*), plus whole-forward timing.  gpurun only."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from circuitvision_b200 import sam2_infer, _lib

variant = sys.argv[1] if len(sys.argv) > 1 else "tiny"
B = int(sys.argv[2]) if len(sys.argv) > 2 else 16
chunk = int(sys.argv[3]) if len(sys.argv) > 3 else B
torch.manual_seed(0)
model = sam2_infer.get_modified_sam2(variant, None, device="cuda:0", use_refinement_layer=True)
model.set_max_batch(chunk)
x = torch.randint(0, 256, (B, 1024, 1024, 3), dtype=torch.uint8, device="cuda")
eng = model.engine()
for _ in range(2):
    eng.forward(x, 0, True, want_high=False, want_low=False, want_mask=True)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
N = 5
for _ in range(N):
    eng.forward(x, 0, True, want_high=False, want_low=False, want_mask=True)
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / N
print(f"{variant} B={B} chunk={chunk}: {ms:.2f} ms per forward = {B / ms * 1e3:.1f} crops/s; launches {eng.launches}")
lib = _lib.load()
lib.cv_profile_reset()
lib.cv_profile_enable(1)
eng.forward(x, 0, True, want_high=False, want_low=False, want_mask=True)
torch.cuda.synchronize()
lib.cv_profile_enable(0)
tab = sorted(_lib.profile_table(), key=lambda r: -r["ms"])
tot = sum(r["ms"] for r in tab)
print(f"sum of kernel times {tot:.2f} ms")
for r in tab:
    rate = r["work"] / (r["ms"] * 1e-3) if r["ms"] > 0 else 0
    print(f"{r['name'][:38]:38s} n={r['launches']:4d} {r['ms']:8.3f} ms {100 * r['ms'] / tot:5.1f}%  work/s {rate / 1e12:8.3f} T")
