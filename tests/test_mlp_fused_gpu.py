"""k_mlp_fused (LayerNorm -> fc1 -> GELU -> fc2 -> +residual in one tcgen05 kernel, csrc/mlp_fused.cu) through the C ABI
(cv_mlp_fused) against plain PyTorch fp32 of the same half-block, for every width with a fused instantiation
(Hiera stage 1-2 widths of tiny/small 96/192, base+ 112/224, large 144/288) and both 16-bit operand formats.

Tolerance: the kernel rounds the normalised operand and the hidden activation to 16 bits (as the unfused GEMM chain does);
against fp32 with the same 16-bit-rounded weights the update (out - x) is held to max |err| <= 1.5 % (fp16) / 6 % (bf16) of
its standard deviation and mean |err| <= 0.2 % / 1 %."""
import ctypes as C

import pytest
import torch

pytestmark = pytest.mark.gpu


def _run(M, Cw, dtype, seed=0):
    from circuitvision_b200 import _lib
    lib = _lib.load()
    g = torch.Generator(device="cpu").manual_seed(seed)
    x = torch.randn(M, Cw, generator=g) * 1.5 + 0.3
    gamma, beta = 1.0 + 0.2 * torch.randn(Cw, generator=g), 0.1 * torch.randn(Cw, generator=g)
    w1 = (torch.randn(4 * Cw, Cw, generator=g) / Cw ** 0.5).to(dtype)
    b1 = 0.1 * torch.randn(4 * Cw, generator=g)
    w2 = (torch.randn(Cw, 4 * Cw, generator=g) / (4 * Cw) ** 0.5).to(dtype)
    b2 = 0.1 * torch.randn(Cw, generator=g)
    dev = "cuda:0"
    X = x.to(dev).contiguous()
    t = [v.to(dev).contiguous() for v in (gamma, beta, w1, b1, w2, b2)]
    rc = lib.cv_mlp_fused(X.data_ptr(), M, Cw, t[0].data_ptr(), t[1].data_ptr(), C.c_float(1e-6), t[2].data_ptr(), t[3].data_ptr(),
                          t[4].data_ptr(), t[5].data_ptr(), int(dtype == torch.float16), torch.cuda.current_stream().cuda_stream)
    _lib.check(rc, "cv_mlp_fused")
    torch.cuda.synchronize()
    xd = x.to(dev)
    ln = torch.nn.functional.layer_norm(xd, (Cw,), t[0], t[1], 1e-6)
    h = torch.nn.functional.gelu(ln @ t[2].float().t() + t[3])
    upd = h @ t[4].float().t() + t[5]
    return X - xd, upd


@pytest.mark.parametrize("Cw", [96, 112, 144, 192, 224, 288])
def test_fused_mlp_matches_fp32(Cw):
    M = 128 * 450 + 37  # more tiles than 2 x 148 CTAs can hold at once, ragged last tile
    for dtype, gmax, gmean in ((torch.float16, 0.015, 0.002), (torch.bfloat16, 0.06, 0.01)):
        got, want = _run(M, Cw, dtype)
        std = want.std().item()
        d = (got - want).abs()
        assert d.max().item() <= gmax * std and d.mean().item() <= gmean * std, (Cw, dtype, d.max().item() / std, d.mean().item() / std)


def test_small_and_single_tile():
    for M in (1, 127, 128, 129, 4096):
        got, want = _run(M, 96, torch.float16, seed=M)
        std = want.std().item() if M > 1 else want.abs().mean().item()
        assert (got - want).abs().max().item() <= 0.02 * std, M
