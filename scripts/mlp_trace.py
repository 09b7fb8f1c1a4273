"""Pipeline timeline of CTA 0 of the fused MLP kernel (cv_mlp_fused_set_trace).  gpurun only.  usage: mlp_trace.py C M"""
import ctypes as C, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from circuitvision_b200 import _lib
Cw = int(sys.argv[1]) if len(sys.argv) > 1 else 192
M = int(sys.argv[2]) if len(sys.argv) > 2 else 1048576
lib = _lib.load()
lib.cv_mlp_fused_set_trace.argtypes = [C.c_void_p]
g = torch.Generator(device="cuda").manual_seed(0)
X = torch.randn(M, Cw, device="cuda", generator=g)
gamma, beta = torch.ones(Cw, device="cuda"), torch.zeros(Cw, device="cuda")
w1 = (torch.randn(4 * Cw, Cw, device="cuda", generator=g) / Cw ** 0.5).half()
w2 = (torch.randn(Cw, 4 * Cw, device="cuda", generator=g) / (4 * Cw) ** 0.5).half()
b1, b2 = torch.zeros(4 * Cw, device="cuda"), torch.zeros(Cw, device="cuda")
st = torch.cuda.current_stream().cuda_stream
def run():
    _lib.check(lib.cv_mlp_fused(X.data_ptr(), M, Cw, gamma.data_ptr(), beta.data_ptr(), C.c_float(1e-6), w1.data_ptr(), b1.data_ptr(),
                                w2.data_ptr(), b2.data_ptr(), 1, st), "cv_mlp_fused")
run(); run(); torch.cuda.synchronize()
buf = torch.zeros(4001, dtype=torch.int64, device="cuda")  # 1 + 12 events x 320 slots
lib.cv_mlp_fused_set_trace(buf.data_ptr())
run(); torch.cuda.synchronize()
lib.cv_mlp_fused_set_trace(None)
h = buf.cpu().numpy().astype("uint64")
recs = sorted(((int(v) & 0xFFFFFFFFFFF, int(v) >> 44) for v in h[1:] if v))
t0 = recs[0][0]
names = {1: "MMA a_full", 2: "MMA fc1 issued", 3: "MMA h_full", 4: "MMA fc2 issued", 5: "EPI s_full", 6: "EPI H done", 7: "EPI y_full",
         8: "EPI tile stored", 9: "LN x ready", 10: "LN a_empty", 11: "LN A written"}
import os
lo = int(os.environ.get('SKIP', '0'))
for t, code in recs[lo:lo + int(os.environ.get('NREC', '260'))]:
    print(f"{(t - t0) / 1000:9.2f} us  {names.get(code // 4096, code // 4096):16s} {code % 4096}")
