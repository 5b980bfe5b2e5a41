"""Multi-GPU sharding of the hot path: independent videos, no exchange step.

One process per GPU (torchrun / torch.distributed); every rank computes the same deterministic assignment of
videos to ranks, encodes and classifies its own shard with its own weight replica, and writes its own per-video
artefacts.  Nothing crosses GPUs on the data path.  The only cross-rank value the workload ever needs is the
per-camera actogram (BASELINE config 5): bin-count vectors, a few kB, summed with one all-reduce (NCCL on GPUs,
gloo in the CPU tests).
"""
from __future__ import annotations

from typing import Dict, List, Optional, Sequence

import torch
import torch.distributed as dist


def partition_videos(paths: Sequence[str], costs: Optional[Sequence[float]], world_size: int) -> List[List[str]]:
    """Longest-processing-time-first assignment: videos sorted by cost (frame count / file size; path breaks
    ties so every rank derives the same answer), each given to the currently least-loaded rank."""
    if world_size < 1:
        raise ValueError("world_size must be >= 1")
    costs = list(costs) if costs is not None else [1.0] * len(paths)
    if len(costs) != len(paths):
        raise ValueError("one cost per path")
    order = sorted(range(len(paths)), key=lambda i: (-float(costs[i]), paths[i]))
    load = [0.0] * world_size
    shards: List[List[str]] = [[] for _ in range(world_size)]
    for i in order:
        r = min(range(world_size), key=lambda k: (load[k], k))
        shards[r].append(paths[i])
        load[r] += float(costs[i])
    return shards


def split_frame_range(n_frames: int, world_size: int, halo: int = 0) -> List[range]:
    """One long video on several GPUs: contiguous spans; `halo` extends each span on both sides (the head
    needs +-seq_len//2 embeddings of context, cbas.py:503-504) clipped to the video."""
    base, extra = divmod(n_frames, world_size)
    out, start = [], 0
    for r in range(world_size):
        stop = start + base + (1 if r < extra else 0)
        out.append(range(max(0, start - halo), min(n_frames, stop + halo)))
        start = stop
    return out


def allreduce_bins(local: Dict[str, torch.Tensor], n_bins: Dict[str, int]) -> Dict[str, torch.Tensor]:
    """Sum per-camera actogram bin vectors over all ranks.  `local` holds this rank's partial counts for the
    cameras it processed (missing cameras count as zero); `n_bins` gives every camera's bin count on all ranks."""
    if not dist.is_available() or not dist.is_initialized():
        return {k: local.get(k, torch.zeros(n, dtype=torch.int64)).to(torch.int64) for k, n in n_bins.items()}
    dev = torch.device("cuda", torch.cuda.current_device()) if dist.get_backend() == "nccl" else torch.device("cpu")
    names = sorted(n_bins)
    flat = torch.zeros(sum(n_bins[k] for k in names), dtype=torch.int64, device=dev)
    off = 0
    for k in names:
        if k in local:
            flat[off:off + n_bins[k]] = local[k].to(dev, torch.int64)
        off += n_bins[k]
    dist.all_reduce(flat, op=dist.ReduceOp.SUM)
    out, off = {}, 0
    for k in names:
        out[k] = flat[off:off + n_bins[k]].cpu()
        off += n_bins[k]
    return out
