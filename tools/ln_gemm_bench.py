"""Isolated timing of the encoder's GEMMs with and without the fused-LayerNorm epilogues (kernel-level C ABI).
usage: ln_gemm_bench.py [iters] [only ...]   only: names from {qkv, up, proj, down}; prints one JSON line per case.
Works with an older library too (CBAS_B200_LIB=...): cases whose entry point is missing are skipped."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from cbas_b200 import _lib
from tests import gpu_util as G

iters = int(sys.argv[1]) if len(sys.argv) > 1 else 30
only = set(sys.argv[2:])
M, D, I = 102912, 768, 3072
has_ln = hasattr(_lib.lib(), "cbas_b200_gemm_ln_a")


def timeit(fn, iters=iters, warm=4):
    for _ in range(warm): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters * 1000.0  # us


def report(name, us, flops):
    print(json.dumps({"case": name, "us": round(us, 1), "tflops": round(flops / us / 1e6, 0)}), flush=True)


g = torch.Generator(device="cuda").manual_seed(0)
h = torch.randn(M, D, device="cuda", generator=g)
if has_ln:
    hb, st = G.ln_stats_init(h)
xn = torch.randn(M, D, device="cuda", generator=g).to(torch.bfloat16)
for name, N in (("qkv", 3 * D), ("up", I)):
    if only and name not in only: continue
    w = (torch.randn(N, D, device="cuda", generator=g) * 0.03).to(torch.bfloat16)
    b = torch.randn(N, device="cuda", generator=g)
    c1 = w.float().sum(1).contiguous()
    out = torch.empty(M, N, device="cuda", dtype=torch.bfloat16)
    epi = 1 if name == "up" else 0
    fl = 2.0 * M * N * D
    report(f"{name} plain epi{epi}", timeit(lambda: G.gemm(xn, w, b, epi=epi, out=out)), fl)
    if has_ln:
        lib = _lib.lib()
        def f():
            _lib.check(lib.cbas_b200_gemm_ln_a(hb.data_ptr(), st.data_ptr(), w.data_ptr(), c1.data_ptr(), b.data_ptr(),
                                              out.data_ptr(), M, N, D, epi, 1e-5, G.stream()), "ln_a")
        report(f"{name} LN-consumer epi{epi}", timeit(f), fl)
    del w, out
for name, K in (("proj", D), ("down", I)):
    if only and name not in only: continue
    a = torch.randn(M, K, device="cuda", generator=g).to(torch.bfloat16)
    w = (torch.randn(D, K, device="cuda", generator=g) * 0.01).to(torch.bfloat16)
    b = torch.randn(D, device="cuda", generator=g) * 0.01
    fl = 2.0 * M * D * K
    hh = h.clone()
    report(f"{name} plain resid (TMA reduce)", timeit(lambda: G.gemm(a, w, b, epi=2, out=hh)), fl)
    if has_ln:
        lib = _lib.lib()
        hb2 = torch.empty(M, D, device="cuda", dtype=torch.bfloat16)
        st2 = torch.zeros(M, G.LN_STAT_FLOATS, device="cuda")
        def f():
            _lib.check(lib.cbas_b200_gemm_resid_ln(a.data_ptr(), w.data_ptr(), b.data_ptr(), hh.data_ptr(), hb2.data_ptr(),
                                                  st.data_ptr(), st2.data_ptr(), M, D, K, G.stream()), "resid_ln")
        report(f"{name} LN-producer", timeit(f), fl)
    del a, w
