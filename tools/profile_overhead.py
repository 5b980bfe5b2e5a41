"""Does the per-kernel CUDA-event profiling change the step time?  (bench.py times with it enabled.)"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from cbas_b200 import _lib
from cbas_b200.encoder import DinoEncoder
enc = DinoEncoder("synthetic:vitb16", "cuda", preprocess="processor", image_size=224, max_frames=512)
frames = [torch.randint(0, 256, (512, 256, 256, 3), dtype=torch.uint8, device="cuda") for _ in range(2)]
for _ in range(5): enc.encode_u8(frames[0])
torch.cuda.synchronize()
def run(steps, prof):
    _lib.profile_enable(prof)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for s in range(steps): enc.encode_u8(frames[s & 1])
    e1.record(); torch.cuda.synchronize()
    _lib.profile_enable(False)
    return e0.elapsed_time(e1) / steps
for rep in range(3):
    print("profiling on : %.3f ms/step" % run(20, True), " off: %.3f ms/step" % run(20, False))
