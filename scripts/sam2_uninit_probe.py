"""Read-before-write probe of the SAM 2.1 engine's work buffers: the same batch through two engines whose buffers were filled
with different bytes at allocation (CVB_SAM2_FILL, optionally one buffer only) must give identical outputs.  gpurun only.
usage: sam2_uninit_probe.py <variant> <n_images> <chunk> <fill byte> [buffer name]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
variant, n, chunk, fill = sys.argv[1], int(sys.argv[2]), int(sys.argv[3]), sys.argv[4]
only = sys.argv[5] if len(sys.argv) > 5 else None
import numpy as np, torch
from circuitvision_b200 import sam2_infer
from oracle import gen_sam2_golden

batch = np.stack([gen_sam2_golden.case_image(100 + i % 6, (1024, 1024)) for i in range(n)])
d = torch.from_numpy(batch).cuda()

def run(fill_byte):
    if fill_byte is None:
        os.environ.pop("CVB_SAM2_FILL", None)
    else:
        os.environ["CVB_SAM2_FILL"] = str(fill_byte)
        if only:
            os.environ["CVB_SAM2_FILL_ONLY"] = only
    m = sam2_infer.build_random_init(variant, device=torch.device("cuda:0"), seed=0, max_batch=chunk)
    r = m.engine().forward(d, 0, True, want_high=False, want_low=True, want_mask=False)
    torch.cuda.synchronize()
    low = r["low"].float().cpu().clone()
    del m, r
    torch.cuda.empty_cache()
    return low

a = run(0)
b = run(int(fill))
diff = (a - b).abs().amax(dim=(1, 2, 3)) / a.std()
nan = torch.isnan(b).flatten(1).any(1)
bad = [(i, float(diff[i])) for i in range(n) if not (diff[i] == 0) or nan[i]]
print(f"{variant} n={n} chunk={chunk} fill={fill} only={only}: {len(bad)} images differ", bad[:12])
