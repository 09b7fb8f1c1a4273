"""Pipeline timeline of CTA 0 of k_gemm_tc (cv_gemm_set_trace).  gpurun only.  usage: gemm_trace.py M N K act res b16 [first] [count]"""
import ctypes as C, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from circuitvision_b200 import _lib
M, N, K, act, res, b16 = (int(x) for x in sys.argv[1:7])
first = int(sys.argv[7]) if len(sys.argv) > 7 else 4
count = int(sys.argv[8]) if len(sys.argv) > 8 else 6
lib = _lib.load()
lib.cv_gemm_set_trace.argtypes = [C.c_void_p]
A = torch.randn(M, K, device="cuda").half()
W = (torch.randn(N, K, device="cuda") * K ** -0.5).half()
bias = torch.randn(N, device="cuda")
X = torch.randn(M, N, device="cuda") if res else None
o16 = torch.empty(M, N, device="cuda", dtype=torch.float16) if b16 else None
o32 = None if b16 else (X if res else torch.empty(M, N, device="cuda"))
st = torch.cuda.current_stream().cuda_stream
e = _lib.cv_gemm_epilogue()
e.bias = bias.data_ptr(); e.act = act; e.operand_fp16 = 1
if res: e.residual = X.data_ptr(); e.ld_res = N
if b16: e.out_16 = o16.data_ptr(); e.ld_16 = N
else: e.out_f32 = o32.data_ptr(); e.ld_f32 = N
def run():
    _lib.check(lib.cv_gemm_ex(A.data_ptr(), K, W.data_ptr(), K, M, N, K, C.byref(e), st), "cv_gemm_ex")
run(); run(); torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record(); run(); e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1)
print(f"M{M} N{N} K{K} act{act} res{res} b16{b16}: {ms * 1e3:.1f} us = {2.0 * M * N * K / ms / 1e9:.1f} TFLOP/s")
buf = torch.zeros(1 + 8 * 256, dtype=torch.int64, device="cuda")
lib.cv_gemm_set_trace(buf.data_ptr())
run(); torch.cuda.synchronize()
lib.cv_gemm_set_trace(None)
h = buf.cpu().numpy().astype("uint64")
recs = sorted(((int(v) & 0xFFFFFFFFFFF, int(v) >> 44) for v in h[1:] if v))
names = {0: "MMA acc free", 1: "MMA first k-block there", 2: "MMA tile issued", 3: "EPI acc ready", 4: "EPI tile done", 5: "TMA tile start"}
t0 = recs[0][0]
for t, code in recs:
    ev, idx = code // 4096, code % 4096
    if first <= idx < first + count:
        print(f"{(t - t0) / 1000:9.2f} us  {names.get(ev, ev):24s} tile {idx}")
