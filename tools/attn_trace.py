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
S = 24
names = {0: "mma0:top", 1: "mma0:S", 2: "mma0:PV", 3: "mma1:top", 4: "mma1:S", 5: "mma1:PV",
         6: "sm0:s_full", 7: "sm0:max", 8: "sm0:p_full", 9: "sm0:o_full", 10: "sm0:done",
         11: "sm1:s_full", 12: "sm1:max", 13: "sm1:p_full", 14: "sm1:o_full", 15: "sm1:done",
         16: "rot:start", 17: "rot:done", 18: "sm0:o_loaded", 19: "sm0:staged", 20: "sm0:bar", 21: "mma0:S_go",
         22: "mma0:PV_go"}
for mode in [0]:
    tr = torch.zeros(64 * S, dtype=torch.int64, device="cuda")
    _lib.lib().cbas_b200_debug_attention_trace(tr.data_ptr())
    attention_tc(qkv, frames, T, heads, cos, sin, 5); torch.cuda.synchronize()
    _lib.lib().cbas_b200_debug_attention_trace(None)
    t = tr.cpu().numpy().reshape(64, S)
    for it in range(4, 10):
        print("item", it, " ".join(f"{n}={t[it, i] - t[it, 0]}" for i, n in names.items()), " | item period", t[it + 1, 0] - t[it, 0])
    print("  mean S issue", np.mean(t[4:40, 1] - t[4:40, 21]), "mean PV issue", np.mean(t[4:40, 2] - t[4:40, 22]),
          "period", np.mean(t[5:41, 0] - t[4:40, 0]))
