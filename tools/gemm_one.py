"""Run a handful of launches of one GEMM configuration (ncu target).  usage: gemm_one.py N K cg epi [epi ...]"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from cbas_b200 import _lib
from tests.gpu_util import gemm
N, K, cg = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])
epis = [int(e) for e in sys.argv[4:]]
M = 102912
a = torch.randn(M, K, device="cuda").to(torch.bfloat16)
w = (torch.randn(N, K, device="cuda") * 0.05).to(torch.bfloat16)
b = torch.randn(N, device="cuda")
_lib.lib().cbas_b200_debug_gemm_cta_group(cg)
for epi in epis:
    out = torch.zeros(M, N, device="cuda", dtype=torch.bfloat16 if epi in (0, 1) else torch.float32)
    for _ in range(3):
        gemm(a, w, b, epi=epi, out=out)
torch.cuda.synchronize()
print("ok")
