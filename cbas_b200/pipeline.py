"""Streamed chunk pipeline: host frames -> H2D -> encode -> D2H, double-buffered on three CUDA streams.

The reference's encode loop (cbas.py:423-440) is strictly serial per 512-frame chunk: CPU decode, a pageable
synchronous H2D of fp32 frames, the forward pass, a synchronous D2H, an HDF5 write.  Here the uint8 frames of
chunk i+1 cross PCIe (pinned staging, copy stream) while chunk i is in the ViT (compute stream) and the
embeddings of chunk i-1 travel back (second copy stream); the host thread only blocks when it needs a slot
whose previous occupant has not finished.

Two ways in:
  run(chunks, sink)          chunks are arrays the caller already holds (pageable numpy, pinned tensors, views of a
                             cudaHostRegister-ed decode ring); pageable ones are staged through pinned memory.
  run_reader(reader, ...)    the pipeline asks `reader.read_into(start, stop, out)` to decode / copy each chunk
                             STRAIGHT INTO the pinned staging slot (no intermediate array, no second memcpy), on a
                             helper thread one chunk ahead of the GPU submission.
"""
from __future__ import annotations

import threading
from typing import Callable, Iterable, Optional, Union

import numpy as np
import torch

from .encoder import DinoEncoder

Chunk = Union[np.ndarray, torch.Tensor]


def _is_device_readable(t: torch.Tensor) -> bool:
    """Host memory the copy engine can read asynchronously: torch-pinned or cudaHostRegister-ed."""
    try:
        return t.is_pinned()
    except RuntimeError:
        return False


class StreamedEncoder:
    """Reusable staging for one encoder and one frame geometry."""

    def __init__(self, encoder: DinoEncoder, frame_hw, chunk_size: int = 512, depth: int = 2, planes: bool = False):
        """planes: chunks are [n,H,W] green planes instead of [n,H,W,3] RGB frames ('reference' preprocessing keeps
        only the green channel, cbas.py:431: a third of the bytes to decode-copy and to move across PCIe)."""
        self.enc = encoder
        self.chunk = int(chunk_size)
        self.depth = int(depth)
        H, W = frame_hw
        self.frame_hw = (int(H), int(W))
        self.planes = bool(planes)
        shape = (self.chunk, H, W) if planes else (self.chunk, H, W, 3)
        dev = encoder.device
        D = encoder.hidden_size
        self.dev_in = [torch.empty(shape, dtype=torch.uint8, device=dev) for _ in range(depth)]
        self.dev_out = [torch.empty(self.chunk, D, dtype=torch.float32, device=dev) for _ in range(depth)]
        self.host_in = [torch.empty(shape, dtype=torch.uint8).pin_memory() for _ in range(depth)]
        self.host_out = [torch.empty(self.chunk, D, dtype=torch.float32).pin_memory() for _ in range(depth)]
        self.s_in = torch.cuda.Stream(device=dev)
        self.s_out = torch.cuda.Stream(device=dev)
        self.ev_in = [torch.cuda.Event() for _ in range(depth)]
        self.ev_comp = [torch.cuda.Event() for _ in range(depth)]
        # the host thread WAITS on these (once per chunk): blocking events put it to sleep instead of spinning on a core -
        # with one process per GPU and two or three host threads each, eight ranks otherwise fight over a 16-core host
        self.ev_out = [torch.cuda.Event(blocking=True) for _ in range(depth)]
        self.h2d_bytes = 0
        self.d2h_bytes = 0

    # ------------------------------------------------------------------------------------------ internals
    def _quiesce(self) -> None:
        """After a failure: nothing of this pipeline may still be in flight when the staging buffers are reused."""
        try:
            self.s_in.synchronize()
            torch.cuda.current_stream(self.enc.device).synchronize()
            self.s_out.synchronize()
        except Exception:
            pass

    def _submit(self, slot: int, n: int, src: torch.Tensor, compute) -> None:
        """H2D of `src` (device-readable host memory) -> encode -> D2H, all asynchronous."""
        with torch.cuda.stream(self.s_in):
            self.dev_in[slot][:n].copy_(src, non_blocking=True)
            self.ev_in[slot].record(self.s_in)
        self.h2d_bytes += src.numel()
        compute.wait_event(self.ev_in[slot])
        if self.planes:
            self.enc.encode_u8_plane(self.dev_in[slot][:n], out=self.dev_out[slot][:n])
        else:
            self.enc.encode_u8(self.dev_in[slot][:n], out=self.dev_out[slot][:n])
        self.ev_comp[slot].record(compute)
        with torch.cuda.stream(self.s_out):
            self.s_out.wait_event(self.ev_comp[slot])
            self.host_out[slot][:n].copy_(self.dev_out[slot][:n], non_blocking=True)
            self.ev_out[slot].record(self.s_out)
        self.d2h_bytes += n * self.enc.hidden_size * 4

    # ------------------------------------------------------------------------------------------ public
    def run(self, chunks: Iterable[Chunk], sink: Callable[[np.ndarray], None]) -> int:
        """Encode every chunk ([n<=chunk,H,W,3] uint8 - [n,H,W] with planes=True -, numpy or CPU tensor) in order; `sink` receives the float32
        [n,D] embeddings of each chunk, in order, as a view that is only valid during the call.
        Returns the number of frames encoded."""
        compute = torch.cuda.current_stream(self.enc.device)
        pending = []  # (slot, n), oldest first
        total = 0
        used = 0      # non-empty chunks so far: slots rotate over THOSE, so a slot is never reused while pending

        def drain_one():
            slot, n = pending.pop(0)
            self.ev_out[slot].synchronize()
            sink(self.host_out[slot][:n].numpy())

        try:
            for ch in chunks:
                n = int(ch.shape[0])
                if n == 0:
                    continue
                if n > self.chunk:
                    raise ValueError("chunk larger than the pipeline's chunk_size")
                slot = used % self.depth
                used += 1
                if len(pending) == self.depth:
                    drain_one()  # the oldest pending chunk holds exactly this slot
                if isinstance(ch, np.ndarray):
                    src = torch.from_numpy(np.ascontiguousarray(ch))
                else:
                    src = ch.contiguous()
                if not _is_device_readable(src):
                    self.host_in[slot][:n].copy_(src)  # pageable -> pinned staging (host memcpy)
                    src = self.host_in[slot][:n]
                self._submit(slot, n, src, compute)
                pending.append((slot, n))
                total += n
            while pending:
                drain_one()
        except BaseException:
            self._quiesce()
            raise
        return total

    def run_reader(self, reader, video_len: int, sink: Callable[[np.ndarray], None],
                   progress_callback: Optional[Callable[[float], None]] = None) -> int:
        """Encode frames [0, video_len) of `reader` in chunks.  `reader.read_into(start, stop, out)` fills the uint8
        array `out` ([stop-start,H,W,3], a view of pinned memory) and runs on a helper thread one chunk ahead, so
        decode / page-cache copies overlap the GPU; `progress_callback(percent)` is called from the CALLING thread
        once per chunk, after the chunk has been read (cbas.py:427-429)."""
        compute = torch.cuda.current_stream(self.enc.device)
        starts = list(range(0, video_len, self.chunk))
        # Several decoders when the reader can provide them (`reader.clone()` / `reader.parallel_readers`): reader
        # thread j owns chunks j, j+R, j+2R, ... and its own capture, so whole chunks decode side by side (OpenCV
        # releases the GIL inside read / cvtColor) - the in-process counterpart of decode.ParallelVideoReader without
        # the start-up cost of worker processes.  Staging slots: one per pipeline stage plus one per reader, rounded up
        # to a multiple of R so that the chunks sharing a slot belong to the same thread (their order is then fixed).
        R = max(1, min(int(getattr(reader, "parallel_readers", 1)), len(starts)))
        readers = [reader]
        try:
            for _ in range(R - 1):
                readers.append(reader.clone())
        except Exception:
            for r in readers[1:]:
                r.close()
            readers, R = [reader], 1
        n_host = -(-(self.depth + R) // R) * R
        while len(self.host_in) < n_host:
            self.host_in.append(torch.empty_like(self.host_in[0]).pin_memory())
        ready = [threading.Event() for _ in starts]
        free = [threading.Event() for _ in range(n_host)]
        for f in free:
            f.set()
        errors = []
        stop = threading.Event()

        def read_loop(j):
            try:
                for k in range(j, len(starts), R):
                    s = starts[k]
                    hs = k % n_host
                    while not free[hs].wait(0.05):
                        if stop.is_set():
                            return
                    free[hs].clear()
                    e = min(s + self.chunk, video_len)
                    readers[j].read_into(s, e, self.host_in[hs][:e - s].numpy())
                    ready[k].set()
            except BaseException as exc:  # re-raised in the calling thread
                errors.append(exc)
                for r in ready:
                    r.set()

        threads = [threading.Thread(target=read_loop, args=(j,), name=f"cbas-b200-reader-{j}", daemon=True) for j in range(R)]
        for th in threads:
            th.start()
        pending = []  # (slot, n, host_slot)
        total = 0

        def drain_one():
            slot, n, hs = pending.pop(0)
            self.ev_out[slot].synchronize()  # D2H done => this chunk's H2D is done: its staging slot is free again
            free[hs].set()
            sink(self.host_out[slot][:n].numpy())

        try:
            for k, s in enumerate(starts):
                n = min(s + self.chunk, video_len) - s
                ready[k].wait()
                if errors:
                    raise errors[0]
                if progress_callback:
                    progress_callback((s + n) / video_len * 100)
                slot = k % self.depth
                if len(pending) == self.depth:
                    drain_one()
                hs = k % n_host
                self._submit(slot, n, self.host_in[hs][:n], compute)
                pending.append((slot, n, hs))
                total += n
            while pending:
                drain_one()
        except BaseException:
            stop.set()
            self._quiesce()
            raise
        finally:
            stop.set()
            for th in threads:
                th.join(timeout=10)
            for r in readers[1:]:
                r.close()
        return total
