"""Stand-alone timing of the fused MLP half-block (cv_mlp_fused) on encoder shapes; ncu target.  gpurun only.
usage: mlp_bench.py [C] [M] [reps]"""
import ctypes as C, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from circuitvision_b200 import _lib

Cw = int(sys.argv[1]) if len(sys.argv) > 1 else 96
M = int(sys.argv[2]) if len(sys.argv) > 2 else 4194304
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 5
lib = _lib.load()
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
run(); torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(reps):
    run()
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / reps
fl = 16.0 * M * Cw * Cw
print(f"mlp_fused C={Cw} M={M}: {ms:.3f} ms = {fl / ms / 1e9:.1f} TFLOP/s, {M * Cw * 8 / ms / 1e6:.0f} GB/s of X traffic, {ms * 1e3 / ((M + 127) // 128 / 148):.2f} us per tile per SM")
