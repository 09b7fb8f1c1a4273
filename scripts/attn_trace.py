"""Pipeline timeline of CTA (0, 0) of the global attention kernel (cv_attn_set_trace).  gpurun only.  usage: attn_trace.py [images]"""
import ctypes as C, os, sys
os.environ.setdefault("CVB_ATTN_G2", "1")  # the instrumented kernel
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from circuitvision_b200 import _lib
B = int(sys.argv[1]) if len(sys.argv) > 1 else 16
lib = _lib.load()
lib.cv_attn_set_trace.argtypes = [C.c_void_p]
H, D, T = 4, 96, 4096
M = B * T
g = torch.Generator(device="cuda").manual_seed(0)
qkv = (torch.randn(M, 3 * H * D, device="cuda", generator=g) * 0.5).to(torch.bfloat16)
out = torch.empty(M, H * D, device="cuda", dtype=torch.bfloat16)
st = torch.cuda.current_stream().cuda_stream
def run():
    rc = lib.cv_attention_bf16(qkv.data_ptr(), 3 * H * D, 3 * H * D, 0, qkv.data_ptr(), 3 * H * D, 3 * H * D, H * D, qkv.data_ptr(),
                               3 * H * D, 3 * H * D, 2 * H * D, M, M, T, T, H, D, C.c_float(D ** -0.5), out.data_ptr(), H * D, st)
    _lib.check(rc, "cv_attention_bf16")
run(); run(); torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record(); run(); e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1)
print(f"global attention B={B}: {ms:.3f} ms = {4.0 * M * T * D * H / ms / 1e9:.1f} TFLOP/s")
buf = torch.zeros(4001, dtype=torch.int64, device="cuda")
lib.cv_attn_set_trace(buf.data_ptr())
run(); torch.cuda.synchronize()
lib.cv_attn_set_trace(None)
h = buf.cpu().numpy().astype("uint64")
n = int(min(h[0], 4000))
recs = sorted(((int(v) & 0xFFFFFFFFFFF, int(v) >> 44) for v in h[1:1 + n] if v))
t0 = recs[0][0]
names = {1: "MMA P ready", 2: "MMA QK issued", 4: "SMX S ready", 5: "SMX P written", 6: "MMA PV issued", 7: "MMA K tile there"}
lo = int(os.environ.get("SKIP", "0"))
for t, code in recs[lo:lo + int(os.environ.get("NREC", "120"))]:
    idx = code % 4096
    print(f"{(t - t0) / 1000:9.2f} us  {names.get(code // 4096, code // 4096):18s} tile {idx >> 1:3d} q{idx & 1}")
