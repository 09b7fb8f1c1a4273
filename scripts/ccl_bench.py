"""Native-resolution CCL (cv_ccl_label, BASELINE cfg 4): achieved GB/s at 5 B/px against the HBM copy peak.  gpurun only."""
import sys, os, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from circuitvision_b200 import synth, nodes, _lib

B = int(sys.argv[1]) if len(sys.argv) > 1 else 16
S = 4096
lib = _lib.load()
base = torch.from_numpy(np.stack([synth.make_schematic(900 + i, S)[0] for i in range(8)])).cuda()
pool = [base[(torch.arange(B, device="cuda") + p) % 8].contiguous() for p in range(2)]  # 2 x B x 16 MiB
labels = torch.empty((B, S, S), dtype=torch.int32, device="cuda")
counts = torch.empty((B,), dtype=torch.int32, device="cuda")
st = torch.cuda.current_stream().cuda_stream
ws_bytes = lib.cv_ccl_workspace_bytes(B, S, S)
ws = torch.empty(ws_bytes, dtype=torch.uint8, device="cuda")
def run(m):
    _lib.check(lib.cv_ccl_label(m.data_ptr(), B, S, S, 8, labels.data_ptr(), counts.data_ptr(), ws.data_ptr(), ws_bytes, st), "cv_ccl_label")
for p in pool:
    run(p)
torch.cuda.synchronize()
lib.cv_profile_reset(); lib.cv_profile_enable(1)
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
N = 6
e0.record()
for i in range(N):
    run(pool[i % 2])
e1.record(); torch.cuda.synchronize()
lib.cv_profile_enable(0)
ms = e0.elapsed_time(e1) / N
bytes_alg = B * S * S * 5.0
peak = json.load(open(os.path.join(os.path.dirname(__file__), "..", "MEASURED_PEAKS.json")))["hbm_gbs"] if os.path.exists(os.path.join(os.path.dirname(__file__), "..", "MEASURED_PEAKS.json")) else 6650.0
print(f"cv_ccl_label B={B} {S}^2 conn 8: {ms:.3f} ms per batch = {B / ms * 1e3:.1f} images/s; algorithmic 5 B/px -> {bytes_alg / ms / 1e6:.1f} GB/s = {bytes_alg / ms / 1e6 / peak:.3f} of HBM copy peak {peak}")
print("components per image:", counts[:4].tolist())
for r in sorted(_lib.profile_table(), key=lambda r: -r["ms"]):
    print(f"  {r['name'][:40]:40s} n={r['launches']:3d} {r['ms'] / N:8.3f} ms/batch")
