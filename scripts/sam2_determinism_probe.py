"""Run-to-run determinism of the SAM 2.1 engine: the same batch several times through one engine; every output and every stage
buffer must be bit-identical (no kernel here uses floating-point atomics).  gpurun only.
usage: sam2_determinism_probe.py <variant> <n_images> [reps]   (A/B switches through the CVB_* environment)"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
variant, n = sys.argv[1], int(sys.argv[2])
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 3
import numpy as np, torch
from circuitvision_b200 import sam2_infer
from oracle import gen_sam2_golden

E = {"tiny": 96, "small": 96, "base_plus": 112, "large": 144}[variant]
batch = np.stack([gen_sam2_golden.case_image(100 + i % 6, (1024, 1024)) for i in range(n)])
d = torch.from_numpy(batch).cuda()
m = sam2_infer.build_random_init(variant, device=torch.device("cuda:0"), seed=0, max_batch=n)
eng = m.engine()
bufs = [("X0", (n, 65536 * E)), ("X1", (n, 16384 * 2 * E)), ("X2", (n, 4096 * 4 * E)), ("X3", (n, 1024 * 8 * E)),
        ("s0", (n, 65536 * 32)), ("s1", (n, 16384 * 64)), ("keys32", (n, 4096 * 256))]
runs = []
for _ in range(reps):
    r = eng.forward(d, 0, True, want_high=False, want_low=True, want_mask=False)
    torch.cuda.synchronize()
    snap = {"low": r["low"].float().reshape(n, -1).clone()}
    for name, shape in bufs:
        snap[name] = eng.read_buffer(name, shape).clone()
    torch.cuda.synchronize()
    runs.append(snap)
sw = {k: v for k, v in os.environ.items() if k.startswith("CVB_")}
print(f"{variant} n={n} {sw}")
for k in range(1, reps):
    line = []
    for name in ["X0", "X1", "X2", "X3", "s0", "s1", "keys32", "low"]:
        a, b = runs[0][name], runs[k][name]
        ne = (a != b).sum(dim=1)
        imgs = [(i, int(ne[i])) for i in range(n) if ne[i] > 0]
        line.append(f"{name}: {len(imgs)} imgs {imgs[:4]}")
    print(f"  run 0 vs {k}: " + " | ".join(line))
# geometry of the X0 differences of the first affected image (token rows -> (y, x) on the 256 x 256 grid)
a, b = runs[0]["X0"], runs[1]["X0"]
ne = (a != b).sum(dim=1)
for i in range(n):
    if ne[i] > 0:
        rows = ((a[i].view(65536, E) != b[i].view(65536, E)).sum(dim=1) > 0).nonzero().flatten()
        ys, xs = (rows // 256).tolist(), (rows % 256).tolist()
        cols = ((a[i].view(65536, E) != b[i].view(65536, E)).sum(dim=0) > 0).sum().item()
        err = (a[i] - b[i]).abs().max().item()
        print(f"  image {i}: {len(rows)} token rows differ, y {min(ys)}..{max(ys)}, x {min(xs)}..{max(xs)}, {cols} of {E} columns, max |diff| {err:.4f}")
        break
