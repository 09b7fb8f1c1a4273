"""The GEMM instantiations the SAM 2.1 engine actually launches (cv_gemm_ex = the engine's launcher): CTA pairs, 16 epilogue
warps, tanh-form GELU, fp16 operands, and the four row maps (window un-partition, ConvTranspose pixel shuffle, 2x2 max-pool of the
Q-pool shortcut, Q-pooled qkv) — each against a plain PyTorch fp32 reference of the same op on the same rounded operands
(reference modules: sam2 hieradet.window_unpartition / do_pool / MultiScaleAttention q-pool, mask_decoder ConvTranspose2d;
reached from src/sam2_infer.py:226-232).  Tolerances: fp32 outputs 2e-3 of the output scale (accumulation order), 16-bit outputs one
more rounding (2^-8 bf16, 2^-10 fp16); the GELU epilogue of 16-bit outputs is the tanh-form minimax (act.cuh), |error| <= 3e-4 |x|."""
import ctypes as C

import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu


def _lib():
    from circuitvision_b200 import _lib as L
    return L, L.load()


def _ops(M, N, K, fp16, seed, wscale=None):
    g = torch.Generator(device="cpu").manual_seed(seed)
    dt = torch.float16 if fp16 else torch.bfloat16
    A = torch.randn(M, K, generator=g).cuda().to(dt)
    W = (torch.randn(N, K, generator=g) * (wscale if wscale else K ** -0.5)).cuda().to(dt)
    b = torch.randn(N, generator=g).cuda()
    return A, W, b, dt


def _call(A, Wt, M, N, K, /, **kw):
    L, lib = _lib()
    e = L.cv_gemm_epilogue()
    for k, v in kw.items():
        setattr(e, k, v.data_ptr() if isinstance(v, torch.Tensor) else v)
    rc = lib.cv_gemm_ex(A.data_ptr(), K, Wt.data_ptr(), K, M, N, K, C.byref(e), torch.cuda.current_stream().cuda_stream)
    L.check(rc, "cv_gemm_ex")
    torch.cuda.synchronize()


def _close(out, ref, tol):
    scale = ref.abs().max().item() + 1e-6
    err = (out.float() - ref).abs().max().item()
    assert err <= tol * scale, f"max err {err} vs scale {scale} (tol {tol})"


@pytest.mark.parametrize("fp16", [0, 1])
@pytest.mark.parametrize("M,N,K", [(4096, 1536, 384), (2048, 3072, 768), (1024 + 128, 768, 256)])
def test_pair_gelu_16bit(M, N, K, fp16):
    """fc1 of stages 3-4: BN = 256 CTA pairs, 16 epilogue warps, tanh-form GELU, 16-bit output (odd tile counts included)."""
    A, W, b, dt = _ops(M, N, K, fp16, 11)
    out = torch.empty(M, N, device="cuda", dtype=dt)
    _call(A, W, M, N, K, bias=b, act=1, out_16=out, ld_16=N, operand_fp16=fp16)
    ref = F.gelu(A.float() @ W.float().t() + b)
    _close(out, ref, 2e-3 + (2 ** -10 if fp16 else 2 ** -8) + 3e-4)


@pytest.mark.parametrize("M,N,K", [(4096, 384, 1536), (3072, 448, 1792), (2048, 768, 3072)])
def test_pair_residual_f32(M, N, K):
    """fc2 of stages 3-4: long-K residual GEMMs (CTA pairs at BN = 192 / 224 / 256), in place on the residual stream."""
    A, W, b, dt = _ops(M, N, K, 1, 12)
    X = torch.randn(M, N, device="cuda")
    ref = A.float() @ W.float().t() + b + X
    _call(A, W, M, N, K, bias=b, residual=X, ld_res=N, out_f32=X, ld_f32=N, operand_fp16=1)
    _close(X, ref, 2e-3)


@pytest.mark.parametrize("ws,H,Wd,Cn", [(8, 64, 64, 96), (4, 32, 32, 192), (14, 64, 64, 384), (7, 32, 32, 768)])
def test_unwindow_residual(ws, H, Wd, Cn):
    """attention proj + window_unpartition + residual: source rows window-major (with Hiera's zero padding for 14 / 7), destination
    rows image-major, padded rows dropped."""
    Bn = 2
    nwx, nwy = -(-Wd // ws), -(-H // ws)
    M, K = Bn * nwx * nwy * ws * ws, Cn
    A, W, b, dt = _ops(M, Cn, K, 1, 13)
    X = torch.randn(Bn * H * Wd, Cn, device="cuda")
    y = (A.float() @ W.float().t() + b).view(Bn, nwy, nwx, ws, ws, Cn).permute(0, 1, 3, 2, 4, 5).reshape(Bn, nwy * ws, nwx * ws, Cn)
    ref = X + y[:, :H, :Wd].reshape(-1, Cn)
    _call(A, W, M, Cn, K, bias=b, residual=X, ld_res=Cn, out_f32=X, ld_f32=Cn, map_mode=1, ws=ws, nwx=nwx, nwy=nwy, H=H, W=Wd,
          operand_fp16=1)
    _close(X, ref, 2e-3)


@pytest.mark.parametrize("ws,H,Cin,Cout", [(8, 64, 96, 192), (4, 32, 192, 384)])
def test_pool2_shortcut(ws, H, Cin, Cout):
    """Q-pool shortcut do_pool(proj(x)): 2x2 max over window-major rows -> image-major rows of the pooled grid."""
    Bn, Wd = 2, H
    nw = H // ws
    M = Bn * H * Wd
    A, W, b, dt = _ops(M, Cout, Cin, 1, 14)
    out = torch.full((Bn * (H // 2) * (Wd // 2), Cout), float("nan"), device="cuda")
    _call(A, W, M, Cout, Cin, bias=b, out_f32=out, ld_f32=Cout, map_mode=3, ws=ws, nwx=nw, nwy=nw, H=H, W=Wd, operand_fp16=1)
    y = (A.float() @ W.float().t() + b).view(Bn, nw, nw, ws, ws, Cout).permute(0, 1, 3, 2, 4, 5).reshape(Bn, H, Wd, Cout)
    ref = F.max_pool2d(y.permute(0, 3, 1, 2), 2, 2).permute(0, 2, 3, 1).reshape(-1, Cout)
    _close(out, ref, 2e-3)


@pytest.mark.parametrize("fp16", [0, 1])
@pytest.mark.parametrize("ws,H,Cin,Cp", [(8, 64, 96, 192), (4, 32, 192, 384)])
def test_qpool_qkv(ws, H, Cin, Cp, fp16):
    """qkv of a Q-pooled block: k | v columns as 16-bit rows, the q columns 2x2 max-pooled into pooled window-major rows."""
    Bn, Wd = 2, H
    nw = H // ws
    M, N = Bn * H * Wd, 3 * Cp
    A, W, b, dt = _ops(M, N, Cin, fp16, 15)
    qkv = torch.zeros(M, N, device="cuda", dtype=dt)
    qp = torch.full((M // 4, Cp), float("nan"), device="cuda", dtype=dt)
    _call(A, W, M, N, Cin, bias=b, out_16=qkv, ld_16=N, map_mode=4, ws=ws, pool_cols=Cp, pool_out=qp, ld_pool=Cp, operand_fp16=fp16)
    y = A.float() @ W.float().t() + b
    tol = 2e-3 + (2 ** -10 if fp16 else 2 ** -8)
    _close(qkv[:, Cp:], y[:, Cp:], tol)
    q = y[:, :Cp].view(-1, ws, ws, Cp)  # [windows, ty, tx, C] (rows are window-major)
    ref = F.max_pool2d(q.permute(0, 3, 1, 2), 2, 2).permute(0, 2, 3, 1).reshape(-1, Cp)
    _close(qp, ref, tol)


@pytest.mark.parametrize("act,rba", [(0, 0), (1, 1)])
def test_shuffle2_upscale(act, rba):
    """mask-decoder upscaling: ConvTranspose2d(k = 2, s = 2) as a GEMM whose columns are (dy, dx, co), plus the high-resolution
    feature added before (upscale 2: GELU(dc2(x) + feat_s0)) or after the activation."""
    Bn, H, Wd, Cin, Co = 2, 32, 32, 256, 64
    M, N = Bn * H * Wd, 4 * Co
    A, W, _, dt = _ops(M, N, Cin, 1, 16)
    g = torch.Generator(device="cpu").manual_seed(17)
    b = torch.randn(Co, generator=g).cuda()
    R = torch.randn(Bn * 2 * H * 2 * Wd, Co, generator=g).cuda()
    out = torch.full_like(R, float("nan"))
    _call(A, W, M, N, Cin, bias=b, act=act, res_before_act=rba, residual=R, ld_res=Co, out_f32=out, ld_f32=Co, map_mode=2, H=H, W=Wd,
          cout=Co, operand_fp16=1)
    y = (A.float() @ W.float().t()).view(Bn, H, Wd, 2, 2, Co).permute(0, 1, 3, 2, 4, 5).reshape(Bn * 2 * H * 2 * Wd, Co) + b
    ref = F.gelu(y + R) if (act and rba) else (F.gelu(y) + R if act else y + R)
    _close(out, ref, 2e-3)


def test_residual_table_rows():
    """positional-embedding style residual: one table shared by every image (res_row_mod)."""
    Bn, T, Cn, K = 3, 1024, 96, 168
    A, W, b, dt = _ops(Bn * T, Cn, K, 1, 18)
    tab = torch.randn(T, Cn, device="cuda")
    out = torch.empty(Bn * T, Cn, device="cuda")
    _call(A, W, Bn * T, Cn, K, residual=tab, ld_res=Cn, res_row_mod=T, out_f32=out, ld_f32=Cn, operand_fp16=1)
    ref = A.float() @ W.float().t() + tab.repeat(Bn, 1)
    _close(out, ref, 2e-3)
