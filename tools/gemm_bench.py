"""Time the tcgen05 GEMM (kernel-level C ABI) against cuBLAS on the encoder's shapes.  Not part of bench.py."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from cbas_b200 import _lib
from tests.gpu_util import gemm

def timeit(fn, iters=20, warm=5):
    for _ in range(warm): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters

M = 102912
shapes = [("qkv", 2304, 768), ("up", 3072, 768), ("down", 768, 3072), ("proj", 768, 768)]
for name, N, K in shapes:
    a = torch.randn(M, K, device="cuda").to(torch.bfloat16)
    w = (torch.randn(N, K, device="cuda") * 0.05).to(torch.bfloat16)
    b = torch.randn(N, device="cuda")
    flops = 2.0 * M * N * K
    out16 = torch.empty(M, N, device="cuda", dtype=torch.bfloat16)
    out32 = torch.zeros(M, N, device="cuda", dtype=torch.float32)
    res = {}
    res["cublas"] = timeit(lambda: torch.matmul(a, w.T, out=out16))
    for cg in (1, 2):
        _lib.lib().cbas_b200_debug_gemm_cta_group(cg)
        for epi, o in ((0, out16), (1, out16), (2, out32), (4, out32)):
            res[f"cg{cg}_epi{epi}"] = timeit(lambda: gemm(a, w, b, epi=epi, out=o))
    _lib.lib().cbas_b200_debug_gemm_cta_group(0)
    print(name, {k: f"{v*1e3:.0f}us {flops/v/1e9:.0f}TF" for k, v in res.items()}, flush=True)
