"""Run-to-run determinism of the integer paths: node analysis (1024^2 and dense 4096^2), terminal reclassification and
cv_ccl_label, each several times on the same inputs; every output must be identical.  gpurun only."""
import os, sys, copy
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from circuitvision_b200 import synth, nodes, _lib
from circuitvision_b200.circuit_analyzer import CircuitAnalyzer
from oracle.node_oracle import node_signature

reps = int(sys.argv[1]) if len(sys.argv) > 1 else 4
A = CircuitAnalyzer(use_sam2=False, debug=True, device=0)

def sig(r, B):
    out = []
    for b in range(B):
        out.append(([(i, u, c.tobytes()) for i, u, c in node_signature(r.nodes(b))], r.connection_points(b)))
    return out

for S, B, seeds in ((1024, 64, range(300, 364)), (4096, 16, range(900, 916))):
    masks, boxes = synth.make_batch(list(seeds), S)
    base = None
    bad = 0
    for k in range(reps):
        r = A.get_node_connections_batch(masks, boxes)
        cur = (sig(r, B), r.emptied.cpu().numpy().tobytes(), r.enhanced.cpu().numpy().tobytes())
        if base is None:
            base = cur
        elif cur != base:
            bad += 1
    print(f"node analysis {S}^2 x {B}: {bad} of {reps - 1} repeats differ")

# terminals
pages, blists = [], []
for s in range(8):
    m, bx, rgb = synth.make_schematic(500 + s, 1024, render_rgb=True)
    ys, xs = np.nonzero(m)
    bx = bx + [{"class": "terminal", "xmin": int(xs[k]) - 10, "ymin": int(ys[k]) - 8, "xmax": int(xs[k]) + 10, "ymax": int(ys[k]) + 8,
                "persistent_uid": f"t{k}"} for k in (5, len(xs) // 3, len(xs) // 2)]
    pages.append(rgb); blists.append(bx)
base, bad = None, 0
for k in range(reps):
    bl = copy.deepcopy(blists)
    A.reclassify_terminals_batch(np.stack(pages), bl)
    cur = [[b["class"] for b in l] for l in bl]
    if base is None: base = cur
    elif cur != base: bad += 1
print(f"terminal reclassification 8 pages: {bad} of {reps - 1} repeats differ")

# native CCL
masks, _ = synth.make_batch(list(range(900, 916)), 4096)
d = torch.from_numpy(masks).cuda()
base, bad = None, 0
for k in range(reps):
    lab, cnt = nodes.ccl_label(d, 8, True)
    torch.cuda.synchronize()
    cur = (lab.clone(), cnt.clone())
    if base is None: base = cur
    elif not (torch.equal(cur[0], base[0]) and torch.equal(cur[1], base[1])): bad += 1
print(f"cv_ccl_label 16 x 4096^2: {bad} of {reps - 1} repeats differ")
