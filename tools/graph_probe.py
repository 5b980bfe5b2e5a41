"""Would replaying the encoder step as a CUDA graph shorten it?  (inter-kernel launch gaps: 178 launches per step)"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from cbas_b200.encoder import DinoEncoder
enc = DinoEncoder("synthetic:vitb16", "cuda", preprocess="processor", image_size=224, max_frames=512)
frames = torch.randint(0, 256, (512, 256, 256, 3), dtype=torch.uint8, device="cuda")
out = torch.empty(512, 768, device="cuda")
for _ in range(5): enc.encode_u8(frames, out=out)
torch.cuda.synchronize()
ref = out.clone()
g = torch.cuda.CUDAGraph()
s = torch.cuda.Stream()
with torch.cuda.stream(s):
    enc.encode_u8(frames, out=out)
    torch.cuda.synchronize()
    with torch.cuda.graph(g, stream=s):
        enc.encode_u8(frames, out=out)
torch.cuda.synchronize()
def t(fn, steps=30):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / steps
for rep in range(3):
    a = t(lambda: enc.encode_u8(frames, out=out)); b = t(g.replay)
    print("eager %.3f ms/step   graph %.3f ms/step" % (a, b))
out.zero_(); g.replay(); torch.cuda.synchronize()
print("graph output equals eager:", torch.equal(out, ref))
