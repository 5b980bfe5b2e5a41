"""A few stand-alone launches of the tcgen05 attention at the bench geometry (for ncu / timing).
usage: attn_one.py [variant] [T] [frames]"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from cbas_b200 import _lib
from cbas_b200.encoder import rope_tables
from tests.gpu_util import attention_tc
variant = int(sys.argv[1]) if len(sys.argv) > 1 else 0
T = int(sys.argv[2]) if len(sys.argv) > 2 else 201
frames = int(sys.argv[3]) if len(sys.argv) > 3 else 512
heads = 12
side = {201: 14, 261: 16}.get(T)
cos, sin = rope_tables(side, side); cos, sin = cos.cuda(), sin.cuda()
qkv = torch.randn(frames * T, 3 * heads * 64, device="cuda").to(torch.float16).view(torch.bfloat16)  # f16 bits (q4 kernel operands)
for _ in range(3): attention_tc(qkv, frames, T, heads, cos, sin, 5)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(10): attention_tc(qkv, frames, T, heads, cos, sin, 5)
e1.record(); torch.cuda.synchronize()
us = e0.elapsed_time(e1) * 100
print(f"variant {variant} T {T} frames {frames}: {us:.1f} us per launch, {4 * T * T * 64 * frames * heads / us / 1e6:.0f} TFLOP/s")
