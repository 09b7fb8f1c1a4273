"""tcgen05 GEMM (through the C ABI) against a plain PyTorch fp32 reference of the same op on the same
bf16-rounded operands.  Tolerance: fp32 accumulation order only -> 2e-3 relative to the output scale
(bf16 outputs: + one bf16 rounding, 2^-8 relative)."""
import pytest
import torch

pytestmark = pytest.mark.gpu


def _run(M, N, K, bias=True, act=0, res=False, bf16_out=False, lda_pad=0):
    from circuitvision_b200 import _lib
    lib = _lib.load()
    g = torch.Generator(device="cpu").manual_seed(M * 7 + N * 3 + K)
    A = torch.randn(M, K + lda_pad, generator=g).cuda().to(torch.bfloat16)
    W = (torch.randn(N, K, generator=g) / K ** 0.5).cuda().to(torch.bfloat16)
    b = torch.randn(N, generator=g).cuda() if bias else None
    R = torch.randn(M, N, generator=g).cuda() if res else None
    out32 = torch.full((M, N), float("nan"), device="cuda")
    out16 = torch.empty((M, N), device="cuda", dtype=torch.bfloat16) if bf16_out else None
    rc = lib.cv_gemm_bf16(A.data_ptr(), K + lda_pad, W.data_ptr(), K, M, N, K, b.data_ptr() if bias else None, act,
                          R.data_ptr() if res else None, N, out32.data_ptr(), N,
                          out16.data_ptr() if bf16_out else None, N, torch.cuda.current_stream().cuda_stream)
    _lib.check(rc, "cv_gemm_bf16")
    torch.cuda.synchronize()
    ref = A[:, :K].float() @ W.float().t()
    if bias:
        ref = ref + b
    if act == 1:
        ref = torch.nn.functional.gelu(ref)
    elif act == 2:
        ref = torch.relu(ref)
    if res:
        ref = ref + R
    scale = ref.abs().max().item() + 1e-6
    err = (out32 - ref).abs().max().item()
    assert err <= 2e-3 * scale, f"M{M} N{N} K{K}: max err {err} vs scale {scale}"
    if bf16_out:
        err16 = (out16.float() - ref).abs().max().item()
        assert err16 <= (2e-3 + 2 ** -8) * scale


@pytest.mark.parametrize("M,N,K", [(128, 32, 64), (256, 96, 96), (1000, 288, 96), (4900, 1152, 384), (333, 64, 72),
                                   (128, 192, 768), (4096, 256, 3072), (38, 2048, 256), (65536, 96, 96)])
def test_gemm_shapes(M, N, K):
    _run(M, N, K)


def test_gemm_epilogues():
    _run(512, 384, 192, bias=True, act=1, res=True, bf16_out=True)
    _run(640, 128, 256, bias=False, act=2, res=False, bf16_out=True)
    _run(300, 96, 96, bias=True, act=0, res=True, lda_pad=8)


def test_gemm_rejects_bad_shapes():
    from circuitvision_b200 import _lib
    lib = _lib.load()
    x = torch.zeros(64, 64, device="cuda", dtype=torch.bfloat16)
    o = torch.zeros(64, 64, device="cuda")
    assert lib.cv_gemm_bf16(x.data_ptr(), 64, x.data_ptr(), 64, 64, 33, 64, None, 0, None, 0, o.data_ptr(), 64, None, 0, None) != 0
