"""A/B of the attention implementations inside the full encoder step, one process, alternating.
usage: ab_attention.py [impl ...]   (2 = tcgen05, 1 = mma.sync, 0 = auto)"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from cbas_b200 import _lib
from cbas_b200.encoder import DinoEncoder
impls = [int(a) for a in sys.argv[1:]] or [2, 1]
enc = DinoEncoder("synthetic:vitb16", "cuda", preprocess="processor", image_size=224, max_frames=512)
frames = [torch.randint(0, 256, (512, 256, 256, 3), dtype=torch.uint8, device="cuda") for _ in range(2)]
outs = {}
for impl in impls:
    enc.set_option(_lib.OPT_ATTENTION_IMPL, impl)
    for _ in range(3): outs[impl] = enc.encode_u8(frames[0]).clone()
torch.cuda.synchronize()
for rep in range(3):
    for impl in impls:
        enc.set_option(_lib.OPT_ATTENTION_IMPL, impl)
        _lib.profile_enable(True)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        steps = 20
        e0.record()
        for s in range(steps): enc.encode_u8(frames[s & 1])
        e1.record(); torch.cuda.synchronize()
        prof = _lib.profile_read(); _lib.profile_enable(False)
        print(f"impl {impl}: {e0.elapsed_time(e1) / steps:.3f} ms/step  attention {prof['attention'][0] / steps:.3f} ms")
a, b = outs[impls[0]], outs[impls[-1]]
print("max |diff| between first and last impl, relative:", float((a - b).abs().max() / a.abs().max()))
