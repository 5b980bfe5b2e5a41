"""Stage timeline of the tcgen05 attention kernel (CTA 0, first items): clock64 stamps -> per-stage cycles."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from cbas_b200 import _lib
from cbas_b200.encoder import rope_tables
from tests.gpu_util import attention_tc
frames, heads, T = 512, 12, 201
cos, sin = rope_tables(14, 14); cos, sin = cos.cuda(), sin.cuda()
qkv = (torch.randn(frames * T, 3 * heads * 64, device="cuda")).to(torch.bfloat16)
for _ in range(3): attention_tc(qkv, frames, T, heads, cos, sin, 5)
tr = torch.zeros(64 * 16, dtype=torch.int64, device="cuda")
_lib.lib().cbas_b200_debug_attention_trace(tr.data_ptr())
attention_tc(qkv, frames, T, heads, cos, sin, 5); torch.cuda.synchronize()
_lib.lib().cbas_b200_debug_attention_trace(None)
t = tr.cpu().numpy().reshape(64, 16)
names = ["mma:loop_top", "mma:qk_ready", "mma:S0_issued", "mma:S1_issued", "mma:PV0_issued", "mma:PV1_issued",
         "sm:s_full", "sm:pass1_done", "sm:max_xchg", "sm:p_full_arrived", "sm:rotate_done", "sm:o_full", "sm:o_empty_arrived"]
base = t[4, 0]
for it in range(4, 12):
    print("item", it, " ".join(f"{n.split(':')[1]}={t[it, i] - t[it, 0]}" for i, n in enumerate(names)), " | item period", t[it + 1, 0] - t[it, 0])
