"""Determinism of the fused patch embedding alone (CVB_SAM2_STOP_AFTER_PE=1): X0 after several runs must be bit-identical."""
import os, sys
os.environ["CVB_SAM2_STOP_AFTER_PE"] = "1"
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
variant, n, reps = sys.argv[1], int(sys.argv[2]), int(sys.argv[3])
import numpy as np, torch
from circuitvision_b200 import sam2_infer
from oracle import gen_sam2_golden
E = {"tiny": 96, "base_plus": 112}[variant]
batch = np.stack([gen_sam2_golden.case_image(100 + i % 6, (1024, 1024)) for i in range(n)])
d = torch.from_numpy(batch).cuda()
m = sam2_infer.build_random_init(variant, device=torch.device("cuda:0"), seed=0, max_batch=n)
eng = m.engine()
snaps = []
for _ in range(reps):
    try:
        eng.forward(d, 0, True, want_high=False, want_low=True, want_mask=False)
    except Exception as e:
        print("forward:", e)
    torch.cuda.synchronize()
    snaps.append(eng.read_buffer("X0", (n, 65536, E)).clone())
    torch.cuda.synchronize()
for k in range(1, reps):
    neq = (snaps[0] != snaps[k])
    rows = neq.any(dim=2)
    print(f"run 0 vs {k}: {int(rows.sum())} token rows differ in {int(rows.any(dim=1).sum())} images")
    idx = rows.nonzero()[:12].tolist()
    for (i, r) in idx:
        c = neq[i, r].nonzero().flatten().tolist()
        print(f"   image {i} token (y {r // 256}, x {r % 256}): {len(c)} columns {c[:6]}.., run0 {snaps[0][i, r, c[0]].item():.5f} vs {snaps[k][i, r, c[0]].item():.5f}")
