"""LayerNorm, fused attention (RoPE prologue) and the preprocess kernels against torch / the oracle."""
import math

import numpy as np
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu

from cbas_b200 import _lib  # noqa: E402
from cbas_b200.encoder import aa_bilinear_taps, rope_tables  # noqa: E402
from oracle import encoder as oenc  # noqa: E402
from tests.gpu_util import attention, attention_tc, layernorm, rel_err, stream  # noqa: E402


@pytest.mark.parametrize("D", [384, 768, 1024])
def test_layernorm(D):
    x = torch.randn(1003, D, device="cuda") * 3 + 0.5
    g = torch.randn(D, device="cuda")
    b = torch.randn(D, device="cuda")
    out = layernorm(x, g, b).float()
    want = F.layer_norm(x, (D,), g, b, 1e-5)
    assert rel_err(out, want) < 5e-3  # bf16 output


def _rope_ref(q, k, cos, sin):
    # modeling_dinov3_vit.py:203-207,238-268 restated on [B,H,T,64] tensors; cos/sin [Np,32] -> tile(2)
    cos, sin = torch.cat([cos, cos], -1), torch.cat([sin, sin], -1)
    P = q.shape[-2] - cos.shape[-2]

    def rot(x):
        x1, x2 = x[..., :32], x[..., 32:]
        return torch.cat((-x2, x1), dim=-1)

    def ap(x):
        xp = x[..., P:, :]
        return torch.cat([x[..., :P, :], xp * cos + rot(xp) * sin], dim=-2)
    return ap(q), ap(k)


@pytest.mark.parametrize("side,heads,frames", [(224, 12, 3), (256, 6, 2), (64, 12, 5), (32, 16, 4)])
def test_attention_with_rope(side, heads, frames):
    n = side // 16
    T, P, D = n * n + 5, 5, heads * 64
    cos, sin = rope_tables(n, n)
    cos, sin = cos.cuda(), sin.cuda()
    qkv = (torch.randn(frames * T, 3 * D, device="cuda") * 1.5).to(torch.bfloat16)
    out = attention(qkv, cos, sin, frames, T, P, heads).float()
    x = qkv.float().view(frames, T, 3, heads, 64).permute(2, 0, 3, 1, 4)  # [3,B,H,T,64]
    q, k = _rope_ref(x[0], x[1], cos, sin)
    want = F.scaled_dot_product_attention(q, k, x[2], scale=0.125).permute(0, 2, 1, 3).reshape(frames * T, D)
    e = rel_err(out, want)
    assert e < 1.5e-2, f"attention rel err {e}"


def _qkv_with_f16_v(frames, T, D, heads, prefix=0, rope=False):
    """QKV buffer as the encoder's GEMM lays it out for the tcgen05 kernels: v IEEE f16 bits; q, k bf16, or f16 as well
    where cbas_b200_attention_tc_qk_f16 says so (frames of at most 256 tokens).  Returns the buffer and the exact
    values it encodes as [3, B, H, T, 64] fp32."""
    raw = torch.randn(frames * T, 3 * D, device="cuda") * 1.5
    qkv = raw.to(torch.bfloat16)
    f16_from = 0 if _lib.lib().cbas_b200_attention_tc_qk_f16(T) else 2 * D
    x16 = raw[:, f16_from:].to(torch.float16)
    qkv.view(torch.int16)[:, f16_from:] = x16.view(torch.int16)
    vals = torch.cat([qkv[:, :f16_from].float(), x16.float()], dim=1)
    return qkv, vals.view(frames, T, 3, heads, 64).permute(2, 0, 3, 1, 4)


@pytest.mark.parametrize("side,heads,frames", [(224, 12, 3), (224, 12, 40), (128, 6, 5), (160, 16, 7), (224, 12, 1)])
def test_attention_tcgen05(side, heads, frames):
    """tcgen05 kernel (S and PV on the 5th-gen tensor cores, P in TMEM) vs torch SDPA; no RoPE inside."""
    n = side // 16
    T, D = n * n + 5, heads * 64
    qkv, x = _qkv_with_f16_v(frames, T, D, heads)
    out = attention_tc(qkv, frames, T, heads).float()
    want = F.scaled_dot_product_attention(x[0], x[1], x[2], scale=0.125).permute(0, 2, 1, 3).reshape(frames * T, D)
    e = rel_err(out, want)
    assert e < 1.5e-2, f"tcgen05 attention rel err {e}"


@pytest.mark.parametrize("side,heads,frames", [(224, 12, 3), (224, 6, 37), (128, 12, 5), (176, 16, 2), (208, 12, 9)])
def test_attention_tcgen05_rope_prologue(side, heads, frames):
    n = side // 16
    T, P, D = n * n + 5, 5, heads * 64
    cos, sin = rope_tables(n, n)
    cos, sin = cos.cuda(), sin.cuda()
    qkv, x = _qkv_with_f16_v(frames, T, D, heads, P, True)
    out = attention_tc(qkv, frames, T, heads, cos, sin, P).float()
    q, k = _rope_ref(x[0], x[1], cos, sin)
    want = F.scaled_dot_product_attention(q, k, x[2], scale=0.125).permute(0, 2, 1, 3).reshape(frames * T, D)
    e = rel_err(out, want)
    assert e < 1.5e-2, f"tcgen05 attention + RoPE prologue rel err {e}"
    qkv_bf = x.permute(1, 3, 0, 2, 4).reshape(frames * T, 3 * D).to(torch.bfloat16)
    legacy = attention(qkv_bf.contiguous(), cos, sin, frames, T, P, heads).float()  # the mma.sync kernel, bf16 V
    assert rel_err(out, legacy) < 1.5e-2


@pytest.mark.parametrize("T,heads,frames,rope_side", [(261, 12, 3, 16), (261, 12, 37, 16), (329, 12, 4, 0), (329, 6, 11, 18),
                                                     (272, 16, 2, 0), (384, 12, 2, 0), (257, 12, 3, 0)])
def test_attention_tcgen05_key_split(T, heads, frames, rope_side):
    """257..384 tokens per frame (256-px frames; DINOv2-with-registers): the key-split kernel merges two partial
    softmaxes per query tile; with and without the RoPE prologue, vs torch SDPA and vs the mma.sync kernel."""
    P, D = 5, heads * 64
    qkv, x = _qkv_with_f16_v(frames, T, D, heads, P, bool(rope_side))
    if rope_side:
        assert rope_side * rope_side + P == T
        cos, sin = rope_tables(rope_side, rope_side)
        cos, sin = cos.cuda(), sin.cuda()
        out = attention_tc(qkv, frames, T, heads, cos, sin, P).float()
        q, k = _rope_ref(x[0], x[1], cos, sin)
    else:
        out = attention_tc(qkv, frames, T, heads).float()
        q, k = x[0], x[1]
    want = F.scaled_dot_product_attention(q, k, x[2], scale=0.125).permute(0, 2, 1, 3).reshape(frames * T, D)
    e = rel_err(out, want)
    assert e < 1.5e-2, f"key-split tcgen05 attention rel err {e}"
    assert torch.isfinite(out).all()


@pytest.mark.parametrize("impl", [0, 1], ids=["auto_tcgen05", "mma_sync"])
def test_attention_every_token_count_residue(impl):
    """Token counts sweeping every residue mod 16 and both sides of every kernel boundary (one tile / two tiles /
    key-split / mma.sync-only), without RoPE and with a synthetic table: masks, partial tiles, clipped stores."""
    try:
        heads, frames, P = 6, 3, 5
        D = heads * 64
        worst = 0.0
        for T in [6, 17, 31, 64, 100, 127, 128, 129, 143, 160, 177, 191, 206, 222, 239, 255, 256, 257, 270, 288, 303,
                  319, 336, 350, 367, 384, 385, 430, 512, 592]:
            for with_rope in (False, True):
                qkv, x = _qkv_with_f16_v(frames, T, D, heads, P, with_rope)
                if with_rope:
                    ang = torch.rand(T - P, 32, device="cuda") * 6.28
                    cos, sin = torch.cos(ang).contiguous(), torch.sin(ang).contiguous()
                    q, k = _rope_ref(x[0], x[1], cos, sin)
                else:
                    cos = sin = None
                    q, k = x[0], x[1]
                if impl == 1 or not _lib.lib().cbas_b200_attention_tc_supported(T, P, 1 if with_rope else 0):
                    buf = x.permute(1, 3, 0, 2, 4).reshape(frames * T, 3 * D).to(torch.bfloat16).contiguous()
                    xb = buf.float().view(frames, T, 3, heads, 64).permute(2, 0, 3, 1, 4)
                    q, k = _rope_ref(xb[0], xb[1], cos, sin) if with_rope else (xb[0], xb[1])
                    v_ref = xb[2]
                    out = attention(buf, cos, sin, frames, T, P, heads).float()
                else:
                    v_ref = x[2]
                    out = attention_tc(qkv, frames, T, heads, cos, sin, P).float()
                want = F.scaled_dot_product_attention(q, k, v_ref, scale=0.125).permute(0, 2, 1, 3).reshape(frames * T, D)
                e = rel_err(out, want)
                worst = max(worst, e)
                assert torch.isfinite(out).all() and e < 1.5e-2, f"T={T} rope={with_rope}: rel err {e}"
        print(f"[parity] attention sweep impl {impl}: worst rel err {worst:.3e}")
    finally:
        pass


def test_preprocess_strided_and_unaligned_frames():
    """Frames that are views into a larger buffer: odd base offset, padded rows, padded frames (the C ABI takes byte
    strides; the vectorised paths must fall back where 16-byte alignment does not hold)."""
    H, W, n, S = 96, 128, 3, 64
    frames = oenc.synthetic_frames(n, H, W, seed=4, structured=False)
    rs, fs, off = W * 3 + 7, (W * 3 + 7) * H + 29, 3
    buf = torch.zeros(off + n * fs + 64, dtype=torch.uint8)
    for f in range(n):
        for y in range(H):
            o = off + f * fs + y * rs
            buf[o:o + W * 3] = torch.from_numpy(frames[f, y].reshape(-1))
    buf = buf.cuda()
    base = buf.data_ptr() + off
    # green
    A = torch.empty(n * (H // 16) * (W // 16), 256, device="cuda", dtype=torch.bfloat16)
    _lib.check(_lib.lib().cbas_b200_preprocess_green(base, A.data_ptr(), n, H, W, fs, rs, stream()), "green")
    g = torch.from_numpy(frames[..., 1].astype(np.float32))
    want = g.view(n, H // 16, 16, W // 16, 16).permute(0, 1, 3, 2, 4).reshape(-1, 256)
    assert torch.equal(A.float().cpu(), want)
    # resize + normalise: same result as from a contiguous copy
    ymin, wy = aa_bilinear_taps(H, S)
    xmin, wx = aa_bilinear_taps(W, S)
    t = lambda a, dt: torch.from_numpy(a).to("cuda", dt).contiguous()
    ymin_d, wy_d, xmin_d, wx_d = t(ymin, torch.int32), t(wy, torch.float32), t(xmin, torch.int32), t(wx, torch.float32)
    ns = S // 16
    outs = []
    fc = torch.from_numpy(frames).cuda()
    for ptr, fstride, rstride in ((base, fs, rs), (fc.data_ptr(), H * W * 3, W * 3)):
        B = torch.empty(n * ns * ns, 768, device="cuda", dtype=torch.bfloat16)
        _lib.check(_lib.lib().cbas_b200_preprocess_resize(
            ptr, B.data_ptr(), n, H, W, fstride, rstride, S, ymin_d.data_ptr(), wy_d.data_ptr(), wy.shape[1],
            xmin_d.data_ptr(), wx_d.data_ptr(), wx.shape[1], stream()), "resize")
        outs.append(B.float().cpu())
    assert float((outs[0] - outs[1]).abs().max()) <= 2.0 ** -7  # per-pixel kernel vs tiled kernel: <= 1 bf16 ulp


def test_rope_tables_match_hf_module():
    from transformers import DINOv3ViTConfig
    from transformers.models.dinov3_vit.modeling_dinov3_vit import DINOv3ViTRopePositionEmbedding
    rope = DINOv3ViTRopePositionEmbedding(DINOv3ViTConfig(hidden_size=768, num_attention_heads=12)).eval()
    cos, sin = rope(torch.zeros(1, 3, 224, 224))
    c, s = rope_tables(14, 14)
    np.testing.assert_allclose(cos[:, :32].numpy(), c.numpy(), atol=1e-6)
    np.testing.assert_allclose(cos[:, 32:].numpy(), c.numpy(), atol=1e-6)
    np.testing.assert_allclose(sin[:, :32].numpy(), s.numpy(), atol=1e-6)


def test_preprocess_green_exact():
    frames = oenc.synthetic_frames(3, 64, 96, seed=1, structured=False)
    f = torch.from_numpy(frames).cuda()
    A = torch.empty(3 * 4 * 6, 256, device="cuda", dtype=torch.bfloat16)
    _lib.check(_lib.lib().cbas_b200_preprocess_green(f.data_ptr(), A.data_ptr(), 3, 64, 96, 64 * 96 * 3, 96 * 3,
                                                     stream()), "green")
    g = torch.from_numpy(frames[..., 1].astype(np.float32))  # [3,64,96]
    want = g.view(3, 4, 16, 6, 16).permute(0, 1, 3, 2, 4).reshape(3 * 24, 256)
    assert torch.equal(A.float().cpu(), want)  # bytes are exact in bf16


@pytest.mark.parametrize("tiled", [2, 1, 0], ids=["column_per_thread", "tiled", "per_pixel"])
@pytest.mark.parametrize("H,W,S", [(256, 256, 224), (96, 128, 64), (224, 224, 224), (48, 48, 64), (480, 640, 224)])
def test_preprocess_resize_matches_oracle(H, W, S, tiled):
    _lib.check(_lib.lib().cbas_b200_debug_resize_tiled(tiled), "knob")
    frames = oenc.synthetic_frames(2, H, W, seed=2)
    want = oenc.preprocess_processor(frames, S)  # [2,3,S,S]
    ns = S // 16
    want = want.view(2, 3, ns, 16, ns, 16).permute(0, 2, 4, 1, 3, 5).reshape(2 * ns * ns, 768)
    ymin, wy = aa_bilinear_taps(H, S)
    xmin, wx = aa_bilinear_taps(W, S)
    t = lambda a, dt: torch.from_numpy(a).to("cuda", dt).contiguous()
    ymin_d, wy_d, xmin_d, wx_d = t(ymin, torch.int32), t(wy, torch.float32), t(xmin, torch.int32), t(wx, torch.float32)
    f = torch.from_numpy(frames).cuda()
    A = torch.empty(2 * ns * ns, 768, device="cuda", dtype=torch.bfloat16)
    _lib.check(_lib.lib().cbas_b200_preprocess_resize(
        f.data_ptr(), A.data_ptr(), 2, H, W, H * W * 3, W * 3, S, ymin_d.data_ptr(), wy_d.data_ptr(), wy.shape[1],
        xmin_d.data_ptr(), wx_d.data_ptr(), wx.shape[1], stream()), "resize")
    if tiled == 2:  # the production kernel does the same arithmetic in the same order as the general tiled one
        B = torch.empty_like(A)
        _lib.lib().cbas_b200_debug_resize_tiled(1)
        _lib.check(_lib.lib().cbas_b200_preprocess_resize(
            f.data_ptr(), B.data_ptr(), 2, H, W, H * W * 3, W * 3, S, ymin_d.data_ptr(), wy_d.data_ptr(), wy.shape[1],
            xmin_d.data_ptr(), wx_d.data_ptr(), wx.shape[1], stream()), "resize")
        assert torch.equal(A, B)
    _lib.lib().cbas_b200_debug_resize_tiled(2)
    got = A.float().cpu()
    # <= 1 bf16 ulp: |x| <= 2.7 -> ulp 2^-7 at most
    err = (got - want).abs()
    assert float(err.max()) <= 2.0 ** -7 + 1e-6, float(err.max())
    assert float((err / want.abs().clamp_min(0.25)).max()) < 2.0 ** -7
