"""Streamed chunk pipeline: host frames -> H2D -> encode -> D2H, double-buffered on three CUDA streams.

The reference's encode loop (cbas.py:423-440) is strictly serial per 512-frame chunk: CPU decode, a pageable
synchronous H2D of fp32 frames, the forward pass, a synchronous D2H, an HDF5 write.  Here the uint8 frames of
chunk i+1 cross PCIe (pinned staging, copy stream) while chunk i is in the ViT (compute stream) and the
embeddings of chunk i-1 travel back (second copy stream); the host thread only blocks when it needs a slot
whose previous occupant has not finished.
"""
from __future__ import annotations

from typing import Callable, Iterable, Optional, Union

import numpy as np
import torch

from .encoder import DinoEncoder

Chunk = Union[np.ndarray, torch.Tensor]


class StreamedEncoder:
    """Reusable staging for one encoder and one frame geometry."""

    def __init__(self, encoder: DinoEncoder, frame_hw, chunk_size: int = 512, depth: int = 2):
        self.enc = encoder
        self.chunk = int(chunk_size)
        self.depth = int(depth)
        H, W = frame_hw
        dev = encoder.device
        D = encoder.hidden_size
        self.dev_in = [torch.empty(self.chunk, H, W, 3, dtype=torch.uint8, device=dev) for _ in range(depth)]
        self.dev_out = [torch.empty(self.chunk, D, dtype=torch.float32, device=dev) for _ in range(depth)]
        self.host_in = [torch.empty(self.chunk, H, W, 3, dtype=torch.uint8).pin_memory() for _ in range(depth)]
        self.host_out = [torch.empty(self.chunk, D, dtype=torch.float32).pin_memory() for _ in range(depth)]
        self.s_in = torch.cuda.Stream(device=dev)
        self.s_out = torch.cuda.Stream(device=dev)
        self.ev_in = [torch.cuda.Event() for _ in range(depth)]
        self.ev_comp = [torch.cuda.Event() for _ in range(depth)]
        self.ev_out = [torch.cuda.Event() for _ in range(depth)]
        self.ev_free = [torch.cuda.Event() for _ in range(depth)]  # dev_in[slot] consumed by the ViT
        self.h2d_bytes = 0
        self.d2h_bytes = 0

    def run(self, chunks: Iterable[Chunk], sink: Callable[[np.ndarray], None]) -> int:
        """Encode every chunk ([n<=chunk,H,W,3] uint8, numpy or (pinned) CPU tensor) in order; `sink` receives the
        float32 [n,D] embeddings of each chunk, in order, as a view that is only valid during the call.
        Returns the number of frames encoded."""
        compute = torch.cuda.current_stream(self.enc.device)
        pending = []  # (slot, n)
        total = 0

        def drain_one():
            slot, n = pending.pop(0)
            self.ev_out[slot].synchronize()
            sink(self.host_out[slot][:n].numpy())

        for i, ch in enumerate(chunks):
            slot = i % self.depth
            if len(pending) == self.depth:
                drain_one()  # frees `slot` (its D2H is done, so its compute and H2D are too)
            n = int(ch.shape[0])
            if n == 0:
                continue
            if n > self.chunk:
                raise ValueError("chunk larger than the pipeline's chunk_size")
            if isinstance(ch, np.ndarray):
                src = torch.from_numpy(np.ascontiguousarray(ch))
            else:
                src = ch.contiguous()
            if not src.is_pinned():
                self.host_in[slot][:n].copy_(src)  # pageable -> pinned staging (host memcpy)
                src = self.host_in[slot][:n]
            with torch.cuda.stream(self.s_in):
                self.dev_in[slot][:n].copy_(src, non_blocking=True)
                self.ev_in[slot].record(self.s_in)
            self.h2d_bytes += src.numel()
            compute.wait_event(self.ev_in[slot])
            self.enc.encode_u8(self.dev_in[slot][:n], out=self.dev_out[slot][:n])
            self.ev_comp[slot].record(compute)
            with torch.cuda.stream(self.s_out):
                self.s_out.wait_event(self.ev_comp[slot])
                self.host_out[slot][:n].copy_(self.dev_out[slot][:n], non_blocking=True)
                self.ev_out[slot].record(self.s_out)
            self.d2h_bytes += n * self.enc.hidden_size * 4
            pending.append((slot, n))
            total += n
        while pending:
            drain_one()
        return total
