"""Golden vectors for the non-default head configurations, from the REAL reference module.  TEST INFRASTRUCTURE.

Imports /root/reference/backend/classifier_head.py unmodified (it needs only torch) and records
ClassifierLSTMDeltas.forward for the hyper-parameters the reference's sweep uses (sweep_runner.py:106-108:
lstm_hidden_size 128, lstm_layers 2) and for use_acceleration=False.  Weights come from
oracle.head.make_head_state(seed), so the fixture stores only outputs.

    python oracle/gen_golden_head_variants.py        ->  tests/golden/head_variants.npz
"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, "/root/reference/backend")
import classifier_head  # noqa: E402  (the reference)
from oracle import head as ohead  # noqa: E402

VARIANTS = {  # name -> (lstm_hidden_size, lstm_layers, use_acceleration)
    "h128_l2": (128, 2, True),
    "h128_l1": (128, 1, True),
    "h64_l2_noacc": (64, 2, False),
    "h64_l1_noacc": (64, 1, False),
}
X_SEED, STATE_SCALE = 21, 2.0


def main():
    torch.manual_seed(0)
    x = torch.from_numpy(np.random.default_rng(X_SEED).standard_normal((12, 31, 768)).astype(np.float16)).float()
    out = {"x_seed": X_SEED, "state_scale": STATE_SCALE}
    for i, (name, (hs, layers, acc)) in enumerate(VARIANTS.items()):
        sd = ohead.make_head_state(768, 9, 128, hs, seed=30 + i, scale=STATE_SCALE, lstm_layers=layers,
                                   use_acceleration=acc)
        m = classifier_head.ClassifierLSTMDeltas(768, 9, seq_len=31, use_acceleration=acc, lstm_hidden_size=hs,
                                                 lstm_layers=layers).eval()
        m.load_state_dict(sd, strict=True)
        with torch.no_grad():
            logits, rawm = m(x)
        out[name + ":logits"], out[name + ":rawm"] = logits.numpy(), rawm.numpy()
        out[name + ":cfg"] = np.array([hs, layers, int(acc), 30 + i])
        print(name, logits.shape, rawm.shape)
    np.savez_compressed(os.path.join(ROOT, "tests", "golden", "head_variants.npz"), **out)


if __name__ == "__main__":
    main()
