"""Parallel CPU video decode feeding the encoder.

The reference decodes every 512-frame chunk with `decord.VideoReader(path, ctx=cpu(0)).get_batch(...)` in the
encoding thread (cbas.py:402,425): one core, ~2 k frames/s for a 256x256 clip, more than ten times slower than one
B200 encodes.  `ParallelVideoReader` keeps that two-call contract (`len(reader)`, `reader.get_batch(range(i, j))`
for consecutive chunks) but decodes chunks ahead of the caller in a pool of worker PROCESSES (OpenCV's FFmpeg
backend, each worker with its own capture, seeking to its chunk), into a ring of shared-memory chunk buffers.  The
array `get_batch` returns is a view of a ring slot and stays valid until two further chunks have been requested -
long enough for the streamed encoder, which has at most two chunks in flight.  The ring is pinned with
cudaHostRegister, so the encoder's copy engine reads the slots directly: no staging memcpy in the encoding thread.

Decode is host work and is excluded from the device-timed benchmark; this module exists so that the
end-to-end path - file in, `_cls.h5` out - is not throttled to one core.
"""
from __future__ import annotations

import multiprocessing as mp
import os
from multiprocessing import shared_memory
from typing import Dict, Optional

import numpy as np

from . import _decode_worker


def default_workers() -> int:
    """Decode processes for one GPU's pipeline: the host's cores, minus two for the Python threads that feed the GPU."""
    return max(1, (os.cpu_count() or 2) - 2)


class ParallelVideoReader:
    """green_only: the ring holds [n,H,W] green planes instead of [n,H,W,3] RGB frames (REFERENCE preprocessing keeps
    nothing else, cbas.py:431).  register_cuda: pin the ring with cudaHostRegister so the copy engine reads the slots
    directly (no staging memcpy in the encoding thread); silently skipped when CUDA is not initialised / available."""

    HOLD = 2  # slots that stay valid behind the newest one: the streamed pipeline has at most two chunks in flight

    def __init__(self, path: str, workers: Optional[int] = None, chunk: int = 512, depth: Optional[int] = None,
                 green_only: bool = False, register_cuda: bool = True):
        import cv2
        from .cbas import _exact_frame_count
        if not os.path.exists(path):
            raise FileNotFoundError(path)
        cap = cv2.VideoCapture(path)
        if not cap.isOpened():
            raise RuntimeError(f"could not open video '{path}'")
        self._len = _exact_frame_count(cap, path)
        self.h, self.w = int(cap.get(cv2.CAP_PROP_FRAME_HEIGHT)), int(cap.get(cv2.CAP_PROP_FRAME_WIDTH))
        cap.release()
        self.path, self.chunk, self.green_only = path, int(chunk), bool(green_only)
        self.frame_shape = (self.h, self.w) if green_only else (self.h, self.w, 3)
        self.n_chunks = (self._len + self.chunk - 1) // self.chunk
        self.workers = max(1, min(int(workers or default_workers()), max(1, self.n_chunks)))
        # slots: one per worker in flight + the newest handed to the caller + HOLD older ones still being copied
        self.depth = int(depth or self.workers + 1 + self.HOLD)
        if self.depth < 2 + self.HOLD:
            raise ValueError(f"depth must be at least {2 + self.HOLD}")
        self._registered = False
        self._procs, self._shm = [], None
        self._next_submit = 0            # next chunk index to hand to a worker
        self._next_yield = 0             # next chunk index the caller will ask for
        self._done: Dict[int, tuple] = {}  # chunk index -> (slot, n)
        self._free = list(range(self.depth))
        self._held = []                  # slots lent to the caller, oldest first
        if self._len == 0:
            return
        ctx = mp.get_context("spawn")    # never fork a process that may hold a CUDA context
        nbytes = self.depth * self.chunk * int(np.prod(self.frame_shape))
        self._shm = shared_memory.SharedMemory(create=True, size=nbytes)
        self._ring = np.ndarray((self.depth, self.chunk) + self.frame_shape, dtype=np.uint8, buffer=self._shm.buf)
        if register_cuda:
            self._register(nbytes)
        self._tasks, self._results = ctx.Queue(), ctx.Queue()
        for _ in range(self.workers):
            p = ctx.Process(target=_decode_worker.run, daemon=True,
                            args=(path, self._shm.name, self.depth, self.chunk, self.h, self.w, self.green_only,
                                  self._tasks, self._results))
            p.start()
            self._procs.append(p)
        self._pump()

    def __len__(self) -> int:
        return self._len

    def _register(self, nbytes: int) -> None:
        """cudaHostRegister the ring: torch then sees its views as pinned and copies from them asynchronously."""
        try:
            import torch
            if not torch.cuda.is_available():
                return
            self._ring_ptr = self._ring.ctypes.data
            rc = torch.cuda.cudart().cudaHostRegister(self._ring_ptr, nbytes, 0)
            self._registered = int(rc) == 0
        except Exception:
            self._registered = False

    def _pump(self) -> None:
        """Hand free ring slots to the workers, in chunk order."""
        while self._free and self._next_submit < self.n_chunks:
            k = self._next_submit
            self._tasks.put((k, k * self.chunk, min((k + 1) * self.chunk, self._len), self._free.pop()))
            self._next_submit += 1

    def get_batch(self, indices) -> np.ndarray:
        """Frames `indices` (a consecutive range that is the next chunk) as uint8 RGB [n,H,W,3]."""
        idx = range(indices.start, indices.stop) if isinstance(indices, range) else list(indices)
        if len(idx) == 0:
            return np.zeros((0,) + self.frame_shape, np.uint8)
        k = idx[0] // self.chunk
        if idx[0] != k * self.chunk or idx[-1] != min((k + 1) * self.chunk, self._len) - 1 or k != self._next_yield:
            raise ValueError("ParallelVideoReader serves consecutive chunk-aligned ranges in order "
                             f"(expected chunk {self._next_yield}, got frames {idx[0]}..{idx[-1]})")
        # the caller (and any asynchronous copy it started) is done with everything but the last HOLD chunks
        while len(self._held) > self.HOLD:
            self._free.append(self._held.pop(0))
        self._pump()
        while k not in self._done:
            try:
                index, slot, n, err = self._results.get(timeout=2.0)
            except Exception:  # queue.Empty: make sure somebody is still decoding
                dead = [p.exitcode for p in self._procs if not p.is_alive()]
                if dead:
                    self.close()
                    raise RuntimeError(f"a decode worker of '{self.path}' exited unexpectedly (exit codes {dead})")
                continue
            if err is not None:
                self.close()
                raise RuntimeError(f"decode worker failed on chunk {index} of '{self.path}':\n{err}")
            self._done[index] = (slot, n)
        slot, n = self._done.pop(k)
        self._held.append(slot)
        self._next_yield += 1
        return self._ring[slot, :n]

    def close(self) -> None:
        if self._procs:
            for _ in self._procs:
                self._tasks.put(None)
            for p in self._procs:
                p.join(timeout=5)
                if p.is_alive():
                    p.terminate()
            self._procs = []
        if self._shm is not None:
            if self._registered:
                try:
                    import torch
                    torch.cuda.synchronize()  # no copy engine may still be reading the ring
                    torch.cuda.cudart().cudaHostUnregister(self._ring_ptr)
                except Exception:
                    pass
                self._registered = False
            self._ring = None
            try:
                self._shm.close()
                self._shm.unlink()
            except (FileNotFoundError, BufferError):
                pass
            self._shm = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
