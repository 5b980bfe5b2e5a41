"""Helpers for the -m gpu tests: call the C ABI (include/cbas_b200.h) on torch CUDA tensors."""
import numpy as np
import torch

from cbas_b200 import _lib


def stream():
    return torch.cuda.current_stream().cuda_stream


def gemm(a_bf16, w_bf16, bias=None, epi=0, out=None):
    """out = epilogue(a @ w.T + bias); epi 0 bf16, 1 gelu bf16, 2 f32 += , 4 f32, 5 gelu f32."""
    M, K = a_bf16.shape
    N = w_bf16.shape[0]
    if out is None:
        out = torch.empty(M, N, device="cuda", dtype=torch.bfloat16 if epi in (0, 1) else torch.float32)
    rc = _lib.lib().cbas_b200_gemm_bf16(a_bf16.data_ptr(), w_bf16.data_ptr(),
                                        bias.data_ptr() if bias is not None else None, out.data_ptr(), M, N, K, epi,
                                        stream())
    _lib.check(rc, "gemm")
    return out


def layernorm(x_f32, g, b, eps=1e-5):
    rows, D = x_f32.shape
    out = torch.empty(rows, D, device="cuda", dtype=torch.bfloat16)
    _lib.check(_lib.lib().cbas_b200_layernorm(x_f32.data_ptr(), g.data_ptr(), b.data_ptr(), out.data_ptr(), rows, D,
                                              eps, stream()), "layernorm")
    return out


def attention(qkv_bf16, cos, sin, frames, T, prefix, heads):
    D = heads * 64
    out = torch.empty(frames * T, D, device="cuda", dtype=torch.bfloat16)
    _lib.check(_lib.lib().cbas_b200_attention(qkv_bf16.data_ptr(), out.data_ptr(),
                                              cos.data_ptr() if cos is not None else None,
                                              sin.data_ptr() if sin is not None else None,
                                              frames, T, prefix, heads, stream()), "attention")
    return out


def rel_err(got, want):
    got, want = got.double(), want.double()
    return float((got - want).abs().max() / want.abs().max().clamp_min(1e-30))


def cosine_rows(a, b):
    a, b = a.double(), b.double()
    return (a * b).sum(-1) / (a.norm(dim=-1) * b.norm(dim=-1)).clamp_min(1e-30)


def attention_tc(qkv_bf16, frames, T, heads, cos=None, sin=None, prefix=0):
    D = heads * 64
    out = torch.zeros(frames * T, D, device="cuda", dtype=torch.bfloat16)
    _lib.check(_lib.lib().cbas_b200_attention_tc(qkv_bf16.data_ptr(), out.data_ptr(),
                                                 cos.data_ptr() if cos is not None else None,
                                                 sin.data_ptr() if sin is not None else None, frames, T, prefix, heads,
                                                 stream()), "attention_tc")
    return out


# ---- fused LayerNorm (csrc/gemm_tcgen05.cuh): kernel-level entry points
LN_STAT_SLOTS, LN_STAT_FLOATS = 16, 36  # csrc/gemm_tcgen05.cuh: sums[16], squares[16], shift, padding


def ln_stats_init(h_f32):
    """-> (hb bf16 [rows, D], stats f32 [rows, 36]) of h (h itself is left as it is)."""
    rows, D = h_f32.shape
    hb = torch.empty(rows, D, device="cuda", dtype=torch.bfloat16)
    stats = torch.empty(rows, LN_STAT_FLOATS, device="cuda", dtype=torch.float32)
    _lib.check(_lib.lib().cbas_b200_ln_stats_init(h_f32.data_ptr(), hb.data_ptr(), stats.data_ptr(), rows, D, stream()),
               "ln_stats_init")
    return hb, stats


def gemm_resid_ln(a_bf16, w_bf16, bias, h_f32, stats_in):
    """h += a @ w.T + bias in place -> (hb bf16, stats_out)."""
    M, K = a_bf16.shape
    N = w_bf16.shape[0]
    hb = torch.empty(M, N, device="cuda", dtype=torch.bfloat16)
    stats = torch.zeros(M, LN_STAT_FLOATS, device="cuda", dtype=torch.float32)
    _lib.check(_lib.lib().cbas_b200_gemm_resid_ln(
        a_bf16.data_ptr(), w_bf16.data_ptr(), bias.data_ptr() if bias is not None else None, h_f32.data_ptr(),
        hb.data_ptr(), stats_in.data_ptr(), stats.data_ptr(), M, N, K, stream()), "gemm_resid_ln")
    return hb, stats


def gemm_ln_a(hb, stats, w_folded_bf16, c1, c2, epi=0, eps=1e-5):
    """bf16 [M, N] = epi(LN(h) W^T + b) from the shifted copy and the row statistics."""
    M, K = hb.shape
    N = w_folded_bf16.shape[0]
    out = torch.empty(M, N, device="cuda", dtype=torch.bfloat16)
    _lib.check(_lib.lib().cbas_b200_gemm_ln_a(hb.data_ptr(), stats.data_ptr(), w_folded_bf16.data_ptr(), c1.data_ptr(),
                                              c2.data_ptr(), out.data_ptr(), M, N, K, epi, eps, stream()), "gemm_ln_a")
    return out
