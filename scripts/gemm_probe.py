"""Runs a list of GEMM shapes once each (after one warm launch) — the target of `ncu --set full` captures.  gpurun only.
usage: gemm_probe.py M,N,K,act,res,b16 [...]"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from circuitvision_b200 import _lib

lib = _lib.load()
st = torch.cuda.current_stream().cuda_stream
reps = int(os.environ.get("REPS", "3"))
for spec in sys.argv[1:]:
    M, N, K, act, res, b16 = (int(x) for x in spec.split(","))
    A = torch.randn(M, K, device="cuda").to(torch.bfloat16)
    W = torch.randn(N, K, device="cuda").to(torch.bfloat16)
    bias = torch.randn(N, device="cuda")
    R = torch.randn(M, N, device="cuda") if res else None
    o32 = None if b16 else (R if res == 2 else torch.empty(M, N, device="cuda"))  # res == 2: in-place residual stream
    o16 = torch.empty(M, N, device="cuda", dtype=torch.bfloat16) if b16 else None
    def run():
        rc = lib.cv_gemm_bf16(A.data_ptr(), K, W.data_ptr(), K, M, N, K, bias.data_ptr(), act, R.data_ptr() if res else None, N,
                              o32.data_ptr() if o32 is not None else None, N, o16.data_ptr() if b16 else None, N, st)
        _lib.check(rc, "gemm")
    run(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        run()
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    byts = M * K * 2 + N * K * 2 + M * N * (2 if b16 else 4) + (M * N * 4 if res else 0)
    print(f"M{M} N{N} K{K} act{act} res{res} {'b16' if b16 else 'f32'}: {ms*1e3:9.1f} us  {2*M*N*K/ms/1e9:8.1f} TFLOP/s  {byts/ms/1e6:8.1f} GB/s", flush=True)
