"""Block-diagonal flash attention (tcgen05) against a plain PyTorch fp32 reference on the same bf16 operands.
Tolerance: P is rounded to bf16 before P·V (2^-9 relative per term) and the output is bf16 -> 1.5e-2 absolute on
outputs of unit scale."""
import pytest
import torch

pytestmark = pytest.mark.gpu
D = 96


def _ref(q, k, v, Wq, Wkv, heads, scale):
    Mq, Mkv = q.shape[0], k.shape[0]
    nw = Mq // Wq
    qf = q.float().view(nw, Wq, heads, D).permute(0, 2, 1, 3)
    kf = k.float()[: nw * Wkv].view(nw, Wkv, heads, D).permute(0, 2, 1, 3)
    vf = v.float()[: nw * Wkv].view(nw, Wkv, heads, D).permute(0, 2, 1, 3)
    a = torch.softmax(qf @ kf.transpose(-1, -2) * scale, dim=-1)
    return (a @ vf).permute(0, 2, 1, 3).reshape(Mq, heads * D)


def _run(nwin, Wq, Wkv, heads, pooled, seed=0, amp=1.0):
    from circuitvision_b200 import _lib
    lib = _lib.load()
    g = torch.Generator().manual_seed(seed)
    C = heads * D
    Mkv, Mq = nwin * Wkv, nwin * Wq
    qkv = (torch.randn(Mkv, 3 * C, generator=g) * amp).cuda().to(torch.bfloat16)
    if pooled:
        q = (torch.randn(Mq, C, generator=g) * amp).cuda().to(torch.bfloat16)
        qa = (q, C, C, 0)
    else:
        q = qkv[:, :C]
        qa = (qkv, 3 * C, 3 * C, 0)
    out = torch.full((Mq, C), float("nan"), device="cuda", dtype=torch.bfloat16)
    scale = D ** -0.5
    rc = lib.cv_attention_bf16(qa[0].data_ptr(), qa[1], qa[2], qa[3], qkv.data_ptr(), 3 * C, 3 * C, C,
                               qkv.data_ptr(), 3 * C, 3 * C, 2 * C, Mq, Mkv, Wq, Wkv, heads, D, scale,
                               out.data_ptr(), C, torch.cuda.current_stream().cuda_stream)
    _lib.check(rc, "cv_attention_bf16")
    torch.cuda.synchronize()
    ref = _ref(q, qkv[:, C:2 * C], qkv[:, 2 * C:], Wq, Wkv, heads, scale)
    err = (out.float() - ref).abs().max().item()
    assert torch.isfinite(out.float()).all()
    assert err < 1.5e-2 * max(1.0, ref.abs().max().item()), f"nwin{nwin} Wq{Wq} Wkv{Wkv} h{heads}: err {err}"


@pytest.mark.parametrize("nwin,Wq,Wkv,heads,pooled", [
    (2, 128, 128, 1, False),     # exactly one tile per window
    (8, 64, 64, 1, False),       # stage-1 windows (8x8)
    (64, 16, 16, 2, False),      # stage-2 windows (4x4)
    (8, 16, 64, 2, True),        # Q-pooled 8x8 -> 4x4
    (5, 196, 196, 4, False),     # stage-3 windows (14x14), ragged against the 128 tiles
    (5, 49, 196, 8, True),       # Q-pooled 14x14 -> 7x7
    (7, 49, 49, 8, False),       # stage-4 windows (7x7)
    (2, 4096, 4096, 4, False),   # global attention over 64x64 tokens, 2 images
    (3, 300, 300, 1, False),     # odd sizes: last q tile partly out of range
])
def test_attention_matches_fp32(nwin, Wq, Wkv, heads, pooled):
    _run(nwin, Wq, Wkv, heads, pooled)


def test_attention_large_logits():
    _run(4, 196, 196, 2, False, seed=3, amp=4.0)


def test_global_attention_lazy_rescale():
    """Two-query-tile global kernel: large logits make tile maxima jump by far more than the 2^8 rescale threshold."""
    _run(1, 1024, 1024, 2, False, seed=5, amp=4.0)
    _run(2, 512, 512, 1, False, seed=6, amp=3.0)
    _run(1, 256, 256, 1, False, seed=7, amp=0.5)
