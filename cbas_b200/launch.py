"""Multi-GPU backlog launcher: one process per GPU, videos sharded across ranks, no collective on the data path.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P \\
        -m cbas_b200.launch --videos 'recordings/**/*.mp4' --encoder synthetic:vitb16 [--model-dir models/M]

Every rank derives the same assignment (`parallel.partition_videos`: longest first by file size, each video to the
least-loaded rank - SURVEY.md 8e), runs `encode_file` on its videos (skipping those whose `_cls.h5` is already
stamped for this encoder, the reference's resume rule, startup_page.py:92-124) and, when a model bundle is given,
`infer_file` on the results (skipping existing CSVs like label_train_page.py:1875-1877).  With `--actogram BEHAVIOUR`
the per-video bin vectors are summed over all ranks (`parallel.allreduce_bins`: the path's only collective, a few kB)
and rank 0 prints the group actogram.  Works single-process (no torchrun) as well.
"""
from __future__ import annotations

import argparse
import glob
import json
import os
import time
import types

import numpy as np
import torch
import torch.distributed as dist


def main(argv=None) -> int:
    ap = argparse.ArgumentParser(description=__doc__, formatter_class=argparse.RawDescriptionHelpFormatter)
    ap.add_argument("--videos", required=True, help="glob of video files (.mp4 / .npy)")
    ap.add_argument("--encoder", default="synthetic:vitb16", help="model identifier for DinoEncoder")
    ap.add_argument("--preprocess", default="reference", choices=["reference", "processor"])
    ap.add_argument("--model-dir", default=None, help="model bundle (model.pth + model_meta.json) for infer_file")
    ap.add_argument("--actogram", default=None, help="behaviour to bin and reduce across ranks")
    ap.add_argument("--framerate", type=float, default=10.0)
    ap.add_argument("--bin-minutes", type=int, default=30)
    ap.add_argument("--threshold", type=float, default=0.5)
    ap.add_argument("--backend", default=None, help="torch.distributed backend (default: nccl, gloo if ranks share a GPU)")
    args = ap.parse_args(argv)

    from . import bundle, cbas, gui_state, parallel, store
    from .classifier_head import actogram_bins
    from .encoder import DinoEncoder

    rank, world = int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    n_dev = torch.cuda.device_count()
    if n_dev == 0:
        raise SystemExit("cbas_b200.launch needs CUDA devices (there is no CPU fallback)")
    dev = torch.device("cuda", local % n_dev)
    torch.cuda.set_device(dev)
    if world > 1:
        backend = args.backend or ("nccl" if world <= n_dev else "gloo")
        dist.init_process_group(backend, **({"device_id": dev} if backend == "nccl" else {}))

    paths = sorted(glob.glob(args.videos, recursive=True))
    paths = [p for p in paths if not p.endswith("_cls.h5")]
    costs = [float(os.path.getsize(p)) for p in paths]
    mine = parallel.partition_videos(paths, costs, world)[rank]

    gui_state.proj = types.SimpleNamespace(encoder_model_identifier=args.encoder, path=os.getcwd())
    enc = DinoEncoder(args.encoder, dev, preprocess=args.preprocess)
    head = meta = None
    if args.model_dir:
        head, meta = bundle.load_model_bundle(args.model_dir, args.encoder, device=dev, in_features=enc.hidden_size)
    name = os.path.basename(os.path.normpath(args.model_dir)) if args.model_dir else None

    t0 = time.perf_counter()
    frames = encoded = classified = 0
    local_bins = {}
    for p in mine:
        h5 = os.path.splitext(p)[0] + "_cls.h5"
        fresh = False
        if os.path.exists(h5):
            with store.EmbeddingReader(h5) as r:
                fresh = r.attrs.get("encoder_model_identifier") == args.encoder
        if not fresh:
            h5 = cbas.encode_file(enc, p)
            encoded += 1
        if h5 is None:
            continue
        with store.EmbeddingReader(h5) as r:
            frames += r.shape[0]
        if head is not None:
            hp = meta["hyperparameters"]
            csv = h5.replace("_cls.h5", f"_{name}_outputs.csv")
            if not os.path.exists(csv):
                csv = cbas.infer_file(h5, head, name, hp["behaviors"], hp["seq_len"], device=dev,
                                      temperature=float(meta.get("calibration", {}).get("temperature", 1.0)))
                classified += 1
            if args.actogram and csv:
                import pandas as pd
                probs = torch.from_numpy(pd.read_csv(csv)[hp["behaviors"]].to_numpy(dtype=np.float32)).to(dev)
                b = hp["behaviors"].index(args.actogram)
                local_bins[p] = actogram_bins(probs, b, args.threshold, int(args.bin_minutes * args.framerate * 60)).cpu()
    torch.cuda.synchronize(dev)
    dt = time.perf_counter() - t0

    stats = torch.tensor([frames, encoded, classified, dt], dtype=torch.float64)
    group_bins = None
    if world > 1:
        gathered = [torch.zeros_like(stats) for _ in range(world)]
        dist.all_gather(gathered, stats.to(dev) if dist.get_backend() == "nccl" else stats)
        per_rank = [g.cpu().tolist() for g in gathered]
    else:
        per_rank = [stats.tolist()]
    if args.actogram and head is not None:
        # every rank needs every video's bin count to lay the vectors out identically
        n_bins = {}
        if world > 1:
            dist.barrier()  # every rank's `_cls.h5` files are published
        for p in paths:
            h5 = os.path.splitext(p)[0] + "_cls.h5"
            if os.path.exists(h5):
                with store.EmbeddingReader(h5) as r:
                    n_bins[p] = -(-r.shape[0] // int(args.bin_minutes * args.framerate * 60))
        summed = parallel.allreduce_bins({k: v for k, v in local_bins.items() if k in n_bins}, n_bins)
        length = max(n_bins.values()) if n_bins else 0
        group_bins = [int(sum(int(v[i]) for v in summed.values() if i < len(v))) for i in range(length)]
    if rank == 0:
        wall = max(r[3] for r in per_rank)
        total = sum(r[0] for r in per_rank)
        print(json.dumps({"videos": len(paths), "world_size": world, "frames": int(total), "seconds": wall,
                          "frames_per_s": total / wall if wall > 0 else None,
                          "per_rank": [{"frames": int(r[0]), "encoded": int(r[1]), "classified": int(r[2]), "seconds": r[3]}
                                       for r in per_rank],
                          "actogram": {"behaviour": args.actogram, "bins": group_bins} if group_bins is not None else None}))
    if world > 1:
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    raise SystemExit(main())
