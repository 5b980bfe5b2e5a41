"""Worker process of cbas_b200.decode.ParallelVideoReader.  Imports only numpy and OpenCV (never torch / CUDA), so a
spawned worker starts in a fraction of a second and cannot touch the GPU context of its parent."""
from __future__ import annotations

import traceback
from multiprocessing import shared_memory

import numpy as np


def run(path: str, shm_name: str, slots: int, chunk: int, h: int, w: int, green_only: bool, tasks, results) -> None:
    import cv2
    shm = shared_memory.SharedMemory(name=shm_name)
    try:
        shape = (slots, chunk, h, w) if green_only else (slots, chunk, h, w, 3)
        ring = np.ndarray(shape, dtype=np.uint8, buffer=shm.buf)
        cap = cv2.VideoCapture(path)
        pos = 0
        while True:
            task = tasks.get()
            if task is None:
                break
            index, start, end, slot = task
            try:
                if not cap.isOpened():
                    raise RuntimeError(f"could not open video '{path}'")
                if start != pos:
                    # the FFmpeg backend seeks to the preceding keyframe and decodes forward to `start`
                    cap.set(cv2.CAP_PROP_POS_FRAMES, start)
                for i in range(start, end):
                    ok, bgr = cap.read()
                    if not ok:
                        raise RuntimeError(f"decode failed at frame {i} of '{path}'")
                    if green_only:
                        # green is channel 1 in BGR and in RGB alike: no colour conversion, a third of the bytes
                        np.copyto(ring[slot, i - start], bgr[:, :, 1])
                    else:
                        cv2.cvtColor(bgr, cv2.COLOR_BGR2RGB, dst=ring[slot, i - start])
                pos = end
                results.put((index, slot, end - start, None))
            except Exception:  # reported to the parent, which raises it in the caller's thread
                pos = -1
                results.put((index, slot, 0, traceback.format_exc()))
        cap.release()
    finally:
        shm.close()
