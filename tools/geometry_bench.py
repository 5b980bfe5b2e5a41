"""Device-timed frames/s of encode_u8 for other geometries / encoder families than bench.py's headline config.
usage: geometry_bench.py [model side preprocess]...   e.g.  synthetic:vitb16 256 reference  synthetic:dinov2reg-b14 256 reference"""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from cbas_b200 import _lib
from cbas_b200.encoder import DinoEncoder

TAGS = ["preprocess", "patch_gemm", "layernorm", "qkv_gemm", "attention", "proj_gemm", "up_gemm", "down_gemm", "final_ln"]
args = sys.argv[1:] or ["synthetic:vitb16", "256", "reference", "synthetic:dinov2reg-b14", "256", "reference"]
for i in range(0, len(args), 3):
    model, side, pre = args[i], int(args[i + 1]), args[i + 2]
    src = side if pre == "reference" else 256
    enc = DinoEncoder(model, "cuda", preprocess=pre, image_size=side, max_frames=512)
    frames = [torch.randint(0, 256, (512, src, src, 3), dtype=torch.uint8, device="cuda") for _ in range(2)]
    for _ in range(3):
        enc.encode_u8(frames[0])
    torch.cuda.synchronize()
    _lib.profile_enable(True)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    steps = 10
    e0.record()
    for s in range(steps):
        enc.encode_u8(frames[s & 1])
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / steps
    prof = _lib.profile_read()
    _lib.profile_enable(False)
    nat = next(iter(enc._native.values()))
    print(json.dumps({"model": model, "side": side, "preprocess": pre, "tokens": nat.tokens, "frames_per_s": 512 / ms * 1e3,
                      "ms_per_512": ms, "kernels_ms": {k: round(v[0] / steps, 3) for k, v in prof.items() if k in TAGS and v[1]}}))
    del enc, frames
    torch.cuda.empty_cache()
