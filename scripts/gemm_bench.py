"""Times the tcgen05 GEMM (C ABI) on the encoder's shapes.  gpurun only.  usage: gemm_bench.py [reps]"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from circuitvision_b200 import _lib

lib = _lib.load()
reps = int(sys.argv[1]) if len(sys.argv) > 1 else 5
B = 16
shapes = [  # M, N, K, act, res, bf16_out
    (B * 65536, 576, 96, 0, False, True),
    (B * 65536, 96, 96, 0, True, False),
    (B * 4096, 1536, 384, 1, False, True),
    (B * 4096, 384, 1536, 0, True, False),
    (B * 4096, 1152, 384, 0, False, True),
    (B * 1024, 3072, 768, 1, False, True),
    (8192, 8192, 8192, 0, False, True),
]
st = torch.cuda.current_stream().cuda_stream
for (M, N, K, act, res, b16) in shapes:
    A = torch.randn(M, K, device="cuda").to(torch.bfloat16)
    W = torch.randn(N, K, device="cuda").to(torch.bfloat16)
    bias = torch.randn(N, device="cuda")
    R = torch.randn(M, N, device="cuda") if res else None
    o32 = None if b16 else torch.empty(M, N, device="cuda")
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
    print(f"M{M} N{N} K{K} act{act} res{int(res)} {'bf16' if b16 else 'f32'}: {ms*1e3:9.1f} us  {2*M*N*K/ms/1e9:8.1f} TFLOP/s  {byts/ms/1e6:8.1f} GB/s")
