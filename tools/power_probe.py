"""Which kernels pull the GPU into the power cap?  Runs each kernel family alone for ~1.5 s and samples NVML."""
import os, sys, time, threading
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch, pynvml
from cbas_b200 import _lib
from cbas_b200.encoder import rope_tables
from tests.gpu_util import gemm, attention_tc, layernorm
pynvml.nvmlInit(); h = pynvml.nvmlDeviceGetHandleByIndex(0)
print("power limit W:", pynvml.nvmlDeviceGetEnforcedPowerLimit(h) / 1000)
M = 102912
a768 = torch.randn(M, 768, device="cuda").to(torch.bfloat16)
a3072 = torch.randn(M, 3072, device="cuda").to(torch.bfloat16)
w_up = (torch.randn(3072, 768, device="cuda") * 0.05).to(torch.bfloat16)
w_dn = (torch.randn(768, 3072, device="cuda") * 0.05).to(torch.bfloat16)
w_qkv = (torch.randn(2304, 768, device="cuda") * 0.05).to(torch.bfloat16)
b3072, b768, b2304 = torch.zeros(3072, device="cuda"), torch.zeros(768, device="cuda"), torch.zeros(2304, device="cuda")
out_up = torch.empty(M, 3072, device="cuda", dtype=torch.bfloat16)
out_qkv = torch.empty(M, 2304, device="cuda", dtype=torch.bfloat16)
hres = torch.randn(M, 768, device="cuda")
qkv = torch.randn(M, 2304, device="cuda").to(torch.bfloat16)
cos, sin = rope_tables(14, 14); cos, sin = cos.cuda(), sin.cuda()
xn = torch.empty(M, 768, device="cuda", dtype=torch.bfloat16)
g = torch.ones(768, device="cuda"); bt = torch.zeros(768, device="cuda")
zeros_a = torch.zeros_like(a768)
w_up_t = w_up.t().contiguous()
sq = torch.randn(8192, 8192, device="cuda").to(torch.bfloat16)
cases = {
    "cuBLAS up shape": (lambda: torch.matmul(a768, w_up.t(), out=out_up), 2 * M * 768 * 3072),
    "cuBLAS down shape": (lambda: torch.matmul(a3072, w_dn.t(), out=xn), 2 * M * 768 * 3072),
    "cuBLAS 8192^3": (lambda: torch.matmul(sq, sq), 2 * 8192 ** 3),
    "up GEMM, plain bf16 out": (lambda: gemm(a768, w_up, b3072, epi=0, out=out_up), 2 * M * 768 * 3072),
    "up GEMM (GELU)": (lambda: gemm(a768, w_up, b3072, epi=1, out=out_up), 2 * M * 768 * 3072),
    "up GEMM, zero A": (lambda: gemm(zeros_a, w_up, b3072, epi=1, out=out_up), 2 * M * 768 * 3072),
    "QKV GEMM": (lambda: gemm(a768, w_qkv, b2304, epi=0, out=out_qkv), 2 * M * 768 * 2304),
    "down GEMM (+resid)": (lambda: gemm(a3072, w_dn, b768, epi=2, out=hres), 2 * M * 768 * 3072),
    "attention": (lambda: attention_tc(qkv, 512, 201, 12, cos, sin, 5), 4 * 201 * 201 * 64 * 12 * 512),
    "layernorm": (lambda: layernorm(hres, g, bt), 0),
}
for name, (fn, flops) in cases.items():
    for _ in range(3): fn()
    torch.cuda.synchronize()
    samples = []
    stop = threading.Event()
    def samp():
        while not stop.is_set():
            samples.append((pynvml.nvmlDeviceGetPowerUsage(h) / 1000, pynvml.nvmlDeviceGetClockInfo(h, pynvml.NVML_CLOCK_SM)))
            time.sleep(0.02)
    t = threading.Thread(target=samp); t.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    n = 0; e0.record(); t0 = time.time()
    while time.time() - t0 < 1.5:
        for _ in range(20): fn()
        n += 20
        torch.cuda.synchronize()
    e1.record(); torch.cuda.synchronize(); stop.set(); t.join()
    ms = e0.elapsed_time(e1) / n
    s = np.array(samples[len(samples) // 3:])
    print(f"{name:22s} {ms * 1000:8.1f} us  {flops / ms / 1e9:7.0f} TFLOP/s  power {s[:, 0].mean():6.0f} W  SM clock {np.median(s[:, 1]):5.0f} MHz")
