"""File-to-file throughput vs in-process decode threads: synthetic mp4 -> cbas_b200.encode_file -> _cls.h5.
usage: e2e_file_threads.py [frames] [threads ...]; prints one JSON line per setting and mode."""
import json, os, sys, tempfile, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from tools.e2e_file_bench import make_clip


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 9000
    settings = [int(a) for a in sys.argv[2:]] or [1, 2, 3, 4, 6]
    with tempfile.TemporaryDirectory() as td:
        clip = os.path.join(td, "clip.mp4")
        make_clip(clip, n)
        import torch
        from cbas_b200 import cbas, store
        from cbas_b200.encoder import DinoEncoder
        cbas.DECODE_WORKERS = 0
        for mode, side in (("processor", 224), ("reference", 256)):
            enc = DinoEncoder("synthetic:vitb16", "cuda", preprocess=mode, image_size=side)
            cbas.DECODE_THREADS = 1
            cbas.encode_file(enc, clip)  # warm-up
            ref = None
            for t in settings:
                cbas.DECODE_THREADS = t
                torch.cuda.synchronize(); t0 = time.time()
                out = cbas.encode_file(enc, clip)
                torch.cuda.synchronize(); dt = time.time() - t0
                with store.EmbeddingReader(out) as r:
                    emb = r.read(0, r.shape[0])
                same = True if ref is None else bool(np.array_equal(ref, emb))
                ref = emb if ref is None else ref
                print(json.dumps({"stage": f"mp4 -> _cls.h5, ViT-B/16 {mode} {side}px", "decode_threads": t, "frames": n,
                                  "fps": round(n / dt), "seconds": round(dt, 3), "bit_identical_to_first": same,
                                  "cores": os.cpu_count()}), flush=True)
                os.remove(out)


if __name__ == "__main__":
    main()
