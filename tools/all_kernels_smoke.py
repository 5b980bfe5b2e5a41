"""Small pass over every kernel family (meant for compute-sanitizer; the tool is closed on this GPU pool, so it
serves as a quick all-kernels smoke run): encoder (both preprocess modes, tcgen05 + key-split +
mma.sync attention, pruned last block), DINOv2-with-registers, head (default and 128/2-layer), actogram."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from cbas_b200 import _lib
from cbas_b200.encoder import DinoEncoder
from cbas_b200.classifier_head import ClassifierLSTMDeltas, actogram_bins
from oracle import head as ohead

torch.manual_seed(0)
for model, side, pre, n in (("synthetic:vits16", 224, "reference", 3), ("synthetic:vits16", 256, "reference", 2),
                            ("synthetic:vits16", 224, "processor", 2), ("synthetic:dinov2reg-s14", 112, "reference", 2),
                            ("synthetic:vits16", 384, "reference", 1)):
    enc = DinoEncoder(model, "cuda", preprocess=pre, image_size=side if pre == "processor" else 224, max_frames=4)
    src = side if pre == "reference" else 256
    frames = torch.randint(0, 256, (n, src, src, 3), dtype=torch.uint8, device="cuda")
    out = enc.encode_u8(frames)
    assert torch.isfinite(out).all()
    print(model, side, pre, tuple(out.shape))
    del enc
for hs, layers, acc in ((64, 1, True), (128, 2, False)):
    sd = ohead.make_head_state(768, 9, 128, hs, seed=1, lstm_layers=layers, use_acceleration=acc)
    m = ClassifierLSTMDeltas(768, 9, lstm_hidden_size=hs, lstm_layers=layers, use_acceleration=acc)
    m.load_state_dict(sd)
    m = m.to("cuda").eval()
    emb = torch.randn(200, 768, device="cuda").half()
    probs = m.infer_embeddings(emb)
    assert torch.isfinite(probs).all()
    bins = actogram_bins(probs, 0, 0.1, 64)
    print("head", hs, layers, acc, tuple(probs.shape), bins.tolist())
torch.cuda.synchronize()
print("ok")
