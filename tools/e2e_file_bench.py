"""File-to-file throughput: synthetic mp4 -> cbas_b200.encode_file -> _cls.h5, with the reference's in-thread decode
(DECODE_WORKERS=0) and with the parallel decoder.  usage: e2e_file_bench.py [frames] [workers ...]
Prints one JSON line per setting.  (Host decode is outside bench.py's device-timed metric)"""
import json, os, sys, tempfile, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np


def make_clip(path, n, side=256):
    import cv2
    vw = cv2.VideoWriter(path, cv2.VideoWriter_fourcc(*"mp4v"), 30.0, (side, side))
    assert vw.isOpened(), "no mp4 encoder in this OpenCV build"
    rng = np.random.default_rng(0)
    yy, xx = np.mgrid[0:side, 0:side].astype(np.float32)
    for i in range(n):
        cx, cy = (0.2 + 0.6 * ((i * 0.013) % 1.0)) * side, (0.3 + 0.4 * ((i * 0.021) % 1.0)) * side
        blob = 160.0 * np.exp(-((xx - cx) ** 2 + (yy - cy) ** 2) / (2.0 * (0.1 * side) ** 2))
        f = np.clip(60 + 0.4 * xx[..., None] + blob[..., None] * np.array([0.6, 0.8, 1.0]) + rng.normal(0, 6, (side, side, 3)), 0, 255)
        vw.write(f.astype(np.uint8))
    vw.release()


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 9000
    settings = [int(a) for a in sys.argv[2:]] or [0, 4, 8, 16]
    decode_only = os.environ.get("DECODE_ONLY") == "1"
    with tempfile.TemporaryDirectory() as td:
        clip = os.path.join(td, "clip.mp4")
        t0 = time.time(); make_clip(clip, n); print(f"# wrote {n} frames in {time.time() - t0:.1f}s", file=sys.stderr)
        if decode_only:
            from cbas_b200.decode import ParallelVideoReader
            for w in settings:
                if w == 0: continue
                t0 = time.time(); r = ParallelVideoReader(clip, workers=w); k = 0
                for i in range(0, len(r), 512): k += len(r.get_batch(range(i, min(i + 512, len(r)))))
                dt = time.time() - t0; r.close()
                print(json.dumps({"stage": "decode_only", "workers": w, "frames": k, "fps": k / dt, "cores": os.cpu_count()}))
            return
        import torch
        from cbas_b200 import cbas
        from cbas_b200.encoder import DinoEncoder
        enc = DinoEncoder("synthetic:vitb16", "cuda", preprocess="processor", image_size=224)
        cbas.DECODE_WORKERS = 0
        cbas.encode_file(enc, clip)  # warm-up: library load, workspace, first launches
        for w in settings:
            cbas.DECODE_WORKERS = w
            torch.cuda.synchronize(); t0 = time.time()
            out = cbas.encode_file(enc, clip)
            torch.cuda.synchronize(); dt = time.time() - t0
            print(json.dumps({"stage": "mp4 -> _cls.h5 (decode + H2D + ViT-B/16 224px + D2H + store)", "decode_workers": w,
                              "frames": n, "fps": n / dt, "seconds": dt, "cores": os.cpu_count()}))
            os.remove(out)


if __name__ == "__main__":
    main()
