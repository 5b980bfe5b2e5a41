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
