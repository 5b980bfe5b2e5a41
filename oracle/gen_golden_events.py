"""Golden vectors for event extraction, from the REAL reference methods.  TEST INFRASTRUCTURE.

Calls Dataset.predictions_to_instances and Dataset.predictions_to_instances_with_confidence of
/root/reference/backend/cbas.py (imported unmodified behind the same I/O stand-ins as oracle/gen_golden.py) on a
synthetic probability CSV and stores the instance lists.

    python -m oracle.gen_golden_events     ->   tests/golden/events.npz
"""
import json
import os
import tempfile
import types

import numpy as np

from oracle import gen_golden as gg


def make_probs(seed=41, n=1500, C=5):
    """Piecewise-constant dominant behaviour with noisy confidence: long runs, flicker, sub-threshold gaps."""
    rng = np.random.default_rng(seed)
    lab = np.repeat(rng.integers(0, C, size=n // 25 + 1), rng.integers(1, 60, size=n // 25 + 1))[:n]
    while len(lab) < n:
        lab = np.concatenate([lab, lab])[:n]
    conf = np.clip(0.75 + 0.25 * np.sin(np.arange(n) / 9.0) + 0.1 * rng.standard_normal(n), 0.21, 0.99)
    p = np.full((n, C), 0.0)
    for i in range(n):
        rest = rng.dirichlet(np.ones(C - 1)) * (1 - conf[i])
        p[i, np.arange(C) != lab[i]] = rest
        p[i, lab[i]] = conf[i]
    return p.astype(np.float32)


def main():
    gg._install_stubs()
    import cbas  # the reference
    import gui_state
    behaviors = ["eating", "drinking", "rearing", "grooming", "background"]
    p = make_probs()
    out = {"seed": 41, "n": len(p), "behaviors": np.array(behaviors)}
    with tempfile.TemporaryDirectory() as td:
        csv = os.path.join(td, "cam1_00001_m1_outputs.csv")
        cbas.pd.DataFrame(p, columns=behaviors).to_csv(csv, index=False)
        fake_self = types.SimpleNamespace(config={"behaviors": behaviors})
        gui_state.proj = types.SimpleNamespace(path=td)
        for thr in (0.7, 0.5, 0.95):
            inst = cbas.Dataset.predictions_to_instances(fake_self, csv, "m1", threshold=thr)
            for d in inst:
                d["video"] = os.path.basename(d["video"])
            out[f"inst_thr{thr}"] = json.dumps(inst)
        for win in (1, 5, 8):
            inst, _ = cbas.Dataset.predictions_to_instances_with_confidence(fake_self, csv, "m1", smoothing_window=win)
            out[f"conf_win{win}"] = json.dumps(inst)
    np.savez_compressed(os.path.join(gg.OUT, "events.npz"), **out)
    print({k: (len(json.loads(str(v))) if k.startswith(("inst", "conf")) else None) for k, v in out.items()})


if __name__ == "__main__":
    main()
