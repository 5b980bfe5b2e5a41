"""Host-side mirror of the hot-path entry points of `backend/cbas.py`:

    encode_file(encoder, path, progress_callback=None) -> str | None        cbas.py:399-456
    infer_file(file_path, model, dataset_name, behaviors, seq_len, device=None, temperature=1.0) -> str | None
                                                                             cbas.py:458-572
    Actogram(...).binned_activity                                            cbas.py:958-1007 (numeric part)

Same signatures, file naming, return values and error behaviour as the reference, so `workthreads.EncodeThread`
and `ClassificationThread` can call them unchanged; all GPU work goes through libcbas_b200.so.
"""
from __future__ import annotations

import os
import re
import traceback
from typing import Callable, Iterator, List, Optional, Tuple

import numpy as np
import torch

from . import gui_state, store
from .classifier_head import ClassifierLSTMDeltas, actogram_bins
from .encoder import DinoEncoder

CHUNK_SIZE = 512  # cbas.py:48
# worker processes that decode chunks ahead of encode_file (cbas_b200/decode.py); 0 = decode in the encoding process
# like the reference (on a helper thread, one chunk ahead), -1 = one worker per host core minus two.
# Set from the environment or by assigning cbas.DECODE_WORKERS.
DECODE_WORKERS = int(os.environ.get("CBAS_B200_DECODE_WORKERS", "0"))
# In-process decode threads per video (each with its own OpenCV capture, whole 512-frame chunks side by side):
# 0 = two when the host has at least eight cores (each capture already runs FFmpeg's own frame threads on every core:
# measured on the 16-core B200 host, mp4 -> _cls.h5 goes 13.2k -> 17.4k frames/s with two and falls off beyond three,
# profiles/e2e_file_threads_r02.jsonl); 1 = the reference's single decoder.
DECODE_THREADS = int(os.environ.get("CBAS_B200_DECODE_THREADS", "0"))


# ----------------------------------------------------------------------------------------------- video decode
def _exact_frame_count(cap, path: str) -> int:
    """Number of frames OpenCV can actually decode.  CAP_PROP_FRAME_COUNT is a container estimate that may exceed it
    (the reference uses decord's exact len(), cbas.py:402-403); when the estimate's last frame cannot be read, count by
    grabbing (no pixel decode) so that encode_file never asks for a frame that does not exist."""
    import cv2
    est = max(0, int(cap.get(cv2.CAP_PROP_FRAME_COUNT)))
    if est == 0:
        return 0
    cap.set(cv2.CAP_PROP_POS_FRAMES, est - 1)
    ok = cap.grab()
    if ok and not cap.grab():
        cap.set(cv2.CAP_PROP_POS_FRAMES, 0)
        return est
    # estimate is off (too long, or too short): count what is there
    cap.release()
    cap.open(path)
    n = 0
    while cap.grab():
        n += 1
    cap.release()
    cap.open(path)
    if n != est:
        print(f"Warning: {os.path.basename(path)}: container reports {est} frames, {n} can be decoded; using {n}.")
    return n


class VideoReader:
    """CPU video decode behind the two calls encode_file makes on decord.VideoReader (cbas.py:402,425):
    len(reader) and reader.get_batch(indices) -> uint8 RGB [n,H,W,3].  Uses decord when it is installed (the
    reference's decoder), OpenCV's FFmpeg backend otherwise; `.npy` files ([N,H,W,3] uint8) are read directly
    (synthetic clips, tests).  Decode/IO errors propagate to the caller like the reference's."""

    def __init__(self, path: str):
        self.path = path
        self._cap = None
        self._arr = None
        self._decord = None
        self._pos = 0
        self._fd = None
        if path.lower().endswith(".npy"):
            self._arr = np.load(path, mmap_mode="r")
            if self._arr.ndim != 4 or self._arr.shape[-1] != 3 or self._arr.dtype != np.uint8:
                raise ValueError(f"{path}: expected a uint8 [N,H,W,3] array")
            self._len = int(self._arr.shape[0])
            self.frame_hw = (int(self._arr.shape[1]), int(self._arr.shape[2]))
            # whole frames are read with pread() straight into the caller's (pinned) buffer: the kernel copies out of
            # the page cache without the per-page faults a fresh memory map costs (one per 4 KB, every call)
            if self._arr.flags["C_CONTIGUOUS"] and hasattr(os, "preadv"):
                self._fd = os.open(path, os.O_RDONLY)
                self._data_offset = int(self._arr.offset)
            return
        try:
            import decord  # type: ignore
            self._decord = decord.VideoReader(path, ctx=decord.cpu(0))
            self._len = len(self._decord)
            self.frame_hw = tuple(int(v) for v in self._decord[0].shape[:2]) if self._len else None
            return
        except ImportError:
            pass
        import cv2
        if not os.path.exists(path):
            raise FileNotFoundError(path)
        cap = cv2.VideoCapture(path)
        if not cap.isOpened():
            raise RuntimeError(f"could not open video '{path}'")
        self._cap = cap
        self._len = _exact_frame_count(cap, path)
        self.frame_hw = (int(cap.get(cv2.CAP_PROP_FRAME_HEIGHT)), int(cap.get(cv2.CAP_PROP_FRAME_WIDTH)))

    def __len__(self) -> int:
        return self._len

    @property
    def parallel_readers(self) -> int:
        """How many independent decoders of this file are worth running side by side (pipeline.run_reader gives each
        its own thread and whole chunks): OpenCV captures decode concurrently and seek frame-exactly on the
        constant-frame-rate files CBAS records; `.npy` clips already copy with a thread pool and decord readers are
        kept single (the reference's one-reader behaviour)."""
        if self._cap is None:
            return 1
        return DECODE_THREADS if DECODE_THREADS > 0 else (2 if (os.cpu_count() or 1) >= 8 else 1)

    def clone(self) -> "VideoReader":
        """A second capture of the same file with its own decode position (frame count and geometry are inherited,
        not re-derived)."""
        if self._cap is None:
            raise RuntimeError("only OpenCV-backed readers clone")
        import cv2
        cap = cv2.VideoCapture(self.path)
        if not cap.isOpened():
            raise RuntimeError(f"could not open video '{self.path}'")
        other = VideoReader.__new__(VideoReader)
        other.path, other._cap, other._arr, other._decord, other._pos, other._fd = self.path, cap, None, None, 0, None
        other._len, other.frame_hw = self._len, self.frame_hw
        return other

    def get_batch(self, indices) -> np.ndarray:
        idx = list(indices)
        if self._arr is not None:
            return np.ascontiguousarray(self._arr[idx[0]:idx[-1] + 1]) if idx else np.zeros((0,) + self._arr.shape[1:], np.uint8)
        if self._decord is not None:
            return self._decord.get_batch(idx).asnumpy()
        import cv2
        frames = []
        for i in idx:
            if i != self._pos:
                self._cap.set(cv2.CAP_PROP_POS_FRAMES, i)
                self._pos = i
            ok, bgr = self._cap.read()
            if not ok:
                raise RuntimeError(f"decode failed at frame {i} of '{self.path}'")
            self._pos += 1
            frames.append(cv2.cvtColor(bgr, cv2.COLOR_BGR2RGB))
        return np.stack(frames) if frames else np.zeros((0, 0, 0, 3), np.uint8)

    def read_into(self, start: int, stop: int, out: np.ndarray) -> None:
        """Frames [start, stop) straight into `out` (uint8 [stop-start,H,W,3], typically pinned staging memory of the
        streamed pipeline): one copy from the page cache / decoder, none through an intermediate array."""
        n = stop - start
        if n <= 0:
            return
        green = out.ndim == 3  # [n,H,W]: the caller wants the green plane only
        if self._arr is not None:
            if not green and self._fd is not None and out.flags["C_CONTIGUOUS"]:
                frame_bytes = int(np.prod(self._arr.shape[1:]))
                dst = memoryview(out).cast("B")
                parts = min(_COPY_THREADS, n)
                cuts = [n * i // parts for i in range(parts + 1)]

                def pread_part(ab):
                    lo, hi = ab[0] * frame_bytes, ab[1] * frame_bytes
                    off = self._data_offset + start * frame_bytes
                    while lo < hi:  # pread may return short counts
                        got = os.preadv(self._fd, [dst[lo:hi]], off + lo)
                        if got <= 0:
                            raise IOError(f"short read from '{self.path}'")
                        lo += got

                if parts <= 1:
                    pread_part((0, n))
                else:
                    list(_copy_pool().map(pread_part, zip(cuts[:-1], cuts[1:])))
                return
            src = self._arr[start:stop, :, :, 1] if green else self._arr[start:stop]
            parts = min(_COPY_THREADS, n)
            if parts <= 1:
                np.copyto(out, src)
                return
            # numpy releases the GIL for plain copies: a few threads reach memory bandwidth, one does not
            cuts = [n * i // parts for i in range(parts + 1)]
            list(_copy_pool().map(lambda ab: np.copyto(out[ab[0]:ab[1]], src[ab[0]:ab[1]]), zip(cuts[:-1], cuts[1:])))
            return
        if self._decord is not None:
            frames = self._decord.get_batch(list(range(start, stop))).asnumpy()
            np.copyto(out, frames[:, :, :, 1] if green else frames)
            return
        import cv2
        for k, i in enumerate(range(start, stop)):
            if i != self._pos:
                self._cap.set(cv2.CAP_PROP_POS_FRAMES, i)
                self._pos = i
            ok, bgr = self._cap.read()
            if not ok:
                raise RuntimeError(f"decode failed at frame {i} of '{self.path}'")
            self._pos += 1
            if green:
                np.copyto(out[k], bgr[:, :, 1])  # channel 1 is green in BGR and RGB alike
            else:
                cv2.cvtColor(bgr, cv2.COLOR_BGR2RGB, dst=out[k])

    def close(self):
        if self._cap is not None:
            self._cap.release()
        if getattr(self, "_fd", None) is not None:
            os.close(self._fd)
            self._fd = None


class _SpanReader:
    """Frames [start, stop) of another reader behind the same three calls (len, get_batch, read_into): what one rank
    sees of a long video that several GPUs encode together (parallel.split_frame_range)."""

    def __init__(self, reader, start: int, stop: int):
        n = len(reader)
        if not 0 <= start <= stop <= n:
            raise ValueError(f"frame range [{start}, {stop}) outside the video's {n} frames")
        self._r, self._start, self._stop = reader, int(start), int(stop)
        self.frame_hw = getattr(reader, "frame_hw", None)
        if hasattr(reader, "read_into"):
            self.read_into = lambda s, e, out: reader.read_into(self._start + s, self._start + e, out)
        if hasattr(reader, "clone"):
            self.parallel_readers = getattr(reader, "parallel_readers", 1)
            self.clone = lambda: _SpanReader(reader.clone(), self._start, self._stop)

    def __len__(self) -> int:
        return self._stop - self._start

    def get_batch(self, indices):
        return self._r.get_batch([self._start + int(i) for i in indices])

    def close(self):
        self._r.close()


_COPY_THREADS = max(1, min(4, (os.cpu_count() or 1) // 2))
_pool = None


def _copy_pool():
    global _pool
    if _pool is None:
        from concurrent.futures import ThreadPoolExecutor
        _pool = ThreadPoolExecutor(max_workers=_COPY_THREADS, thread_name_prefix="cbas-b200-copy")
    return _pool


# ----------------------------------------------------------------------------------------------- encode
def _make_pipeline(encoder, frame_hw: Tuple[int, int], planes: bool = False):
    """Streaming H2D -> ViT -> D2H pipeline for one encoder/geometry (cached on the encoder object)."""
    from .pipeline import StreamedEncoder
    cache = encoder.__dict__.setdefault("_pipelines", {})
    key = (tuple(frame_hw), CHUNK_SIZE, bool(planes))
    pipe = cache.get(key)
    if pipe is None:
        pipe = StreamedEncoder(encoder, frame_hw, CHUNK_SIZE, depth=2, planes=planes)
        cache[key] = pipe
    return pipe


def encode_file(encoder, path: str, progress_callback: Optional[Callable[[float], None]] = None, *,
                frame_range: Optional[Tuple[int, int]] = None, out_path: Optional[str] = None) -> Optional[str]:
    """Encode a video into `<video>_cls.h5` (cbas.py:399-456).

    Returns the output path, or None for a video with no frames; decode, I/O and compute errors are raised after
    the `.tmp` file has been removed.  The finished file appears atomically (os.replace).  `progress_callback`
    receives the percentage after each 512-frame chunk has been decoded, from the calling thread.

    Beyond the reference's signature (keyword-only, for one long video on several GPUs - launch.py --split-video):
    `frame_range=(start, stop)` encodes only those frames, `out_path` names the file they go to."""
    if not isinstance(encoder, DinoEncoder):
        raise TypeError("cbas_b200.encode_file needs a cbas_b200.DinoEncoder (there is no PyTorch fallback path)")
    # decoder errors propagate, as in the reference (cbas.py:400-402)
    # 'reference' preprocessing keeps only the green channel (cbas.py:431): decode-copy and ship that plane alone
    green_only = encoder.preprocess == "reference"
    if DECODE_WORKERS != 0 and not path.lower().endswith(".npy"):
        from .decode import ParallelVideoReader, default_workers
        reader = ParallelVideoReader(path, workers=DECODE_WORKERS if DECODE_WORKERS > 0 else default_workers(),
                                     chunk=CHUNK_SIZE, green_only=green_only)
    else:
        reader = VideoReader(path)
    if frame_range is not None:
        try:
            reader = _SpanReader(reader, frame_range[0], frame_range[1])
        except Exception:
            reader.close()
            raise
    video_len = len(reader)
    if video_len == 0:
        print(f"Warning: Video {path} contains no frames. Skipping.")
        reader.close()
        return None

    out_file_path = out_path or (os.path.splitext(path)[0] + "_cls.h5")
    tmp_file_path = out_file_path + ".tmp"
    writer = None
    try:
        attrs = {}
        if gui_state.proj:
            attrs["encoder_model_identifier"] = gui_state.proj.encoder_model_identifier
            attrs["schema_version"] = store.SCHEMA_VERSION
        writer = store.EmbeddingWriter(tmp_file_path, encoder.hidden_size, attrs)

        def chunks() -> Iterator[np.ndarray]:
            for i in range(0, video_len, CHUNK_SIZE):
                end_index = min(i + CHUNK_SIZE, video_len)
                frames_np = reader.get_batch(range(i, end_index))
                if progress_callback:
                    progress_callback((end_index / video_len) * 100)
                yield frames_np

        def sink(emb: np.ndarray) -> None:
            writer.append(emb)  # float32 -> float16 cast on store, like dset[-n:] = embeddings_out (cbas.py:438)
            writer.flush()

        if hasattr(reader, "read_into") and getattr(reader, "frame_hw", None):
            # in-thread readers: decode / copy each chunk straight into the pipeline's pinned staging, one chunk ahead
            pipe = _make_pipeline(encoder, tuple(reader.frame_hw), planes=green_only)
            with torch.no_grad():
                pipe.run_reader(reader, video_len, sink, progress_callback)
        else:
            # the parallel decoder hands out views of its (cudaHostRegister-ed) shared-memory ring
            it = chunks()
            first = next(it)
            pipe = _make_pipeline(encoder, tuple(first.shape[1:3]), planes=first.ndim == 3)

            def all_chunks():
                yield first
                yield from it

            with torch.no_grad():
                pipe.run(all_chunks(), sink)
        writer.close()
        writer = None
        os.replace(tmp_file_path, out_file_path)
        print(f"Successfully encoded {os.path.basename(path)} to {os.path.basename(out_file_path)}")
        return out_file_path
    except Exception as e:
        print(f"ERROR during encoding for {path}: {e}")
        if writer is not None:
            writer.abort()
        if os.path.exists(tmp_file_path):
            try:
                os.remove(tmp_file_path)
            except OSError:
                pass
        raise e
    finally:
        reader.close()


# ----------------------------------------------------------------------------------------------- infer
# frames classified per read of the `_cls.h5` file (the reference streams 20 000 at a time to bound host memory,
# cbas.py:482; a B200 wants longer launches: 262 144 ViT-B rows are 0.4 GB of f16 on the host and on the device)
INFERENCE_CHUNK_SIZE = 262144


def _as_native_head(model, device: torch.device) -> ClassifierLSTMDeltas:
    """Accept our head, or the reference's `classifier_head.ClassifierLSTMDeltas` instance (same state_dict)."""
    if isinstance(model, ClassifierLSTMDeltas):
        p = next(model.parameters())
        return model if p.device == device else model.to(device)
    if type(model).__name__.startswith("ClassifierLSTMDeltas"):
        head = ClassifierLSTMDeltas(
            in_features=model.in_features, out_features=model.out_features, seq_len=model.seq_len,
            bottleneck_dim=model.cls_ln.normalized_shape[0], use_acceleration=model.use_acceleration,
            ema_alpha=model.ema_alpha, center_window_size=model.sw,
            lstm_hidden_size=model.lstm.hidden_size, lstm_layers=model.lstm.num_layers)
        head.load_state_dict(model.state_dict(), strict=True)
        return head.to(device)
    raise TypeError(f"unsupported head architecture '{type(model).__name__}' (v3 inference needs ClassifierLSTMDeltas)")


def infer_file(file_path: str, model, dataset_name: str, behaviors: List[str], seq_len: int, device=None,
               temperature=1.0) -> Optional[str]:
    """Run the head over one `_cls.h5` and write `<video>_<dataset_name>_outputs.csv` (cbas.py:458-572).

    One probability row per frame (window centred on the frame, replicate padding at the ends of the video,
    softmax(logits / max(1e-3, temperature))), columns = behaviors, no index.  Like the reference, any failure is
    printed with its traceback and reported as None rather than raised."""
    output_file = file_path.replace("_cls.h5", f"_{dataset_name}_outputs.csv")
    if device is None:
        device = torch.device("cuda" if torch.cuda.is_available() else "cpu")
    device = torch.device(device)
    if device.type == "cuda" and device.index is None:
        device = torch.device("cuda", torch.cuda.current_device())
    try:
        import pandas as pd
        if device.type != "cuda":
            raise RuntimeError("cbas_b200.infer_file runs on CUDA only (no CPU fallback)")
        head = _as_native_head(model, device)
        if seq_len != head.seq_len:
            raise ValueError(f"seq_len {seq_len} does not match the model's window ({head.seq_len})")
        if len(behaviors) != head.out_features:
            raise ValueError("behaviors does not match the model's output width")
        half = seq_len // 2
        with store.EmbeddingReader(file_path) as f:
            total_frames = f.shape[0]
            if total_frames == 0:
                print(f"Warning: HDF5 file {file_path} is empty.")
                return None
            # Bounded memory like the reference (cbas.py:482,497-508): INFERENCE_CHUNK_SIZE target frames at a time,
            # read with +-seq_len//2 frames of context; only the probabilities (a few bytes per frame) accumulate.
            # infer_embeddings replicate-pads the ends of what it is given, which is the video's own padding for the
            # first and last chunk and falls on discarded context rows everywhere else.
            parts = []
            for start_idx in range(0, total_frames, INFERENCE_CHUNK_SIZE):
                end_idx = min(start_idx + INFERENCE_CHUNK_SIZE, total_frames)
                read_start, read_end = max(0, start_idx - half), min(total_frames, end_idx + half)
                emb = np.ascontiguousarray(f.read(read_start, read_end))
                if emb.dtype != np.float16:
                    emb = emb.astype(np.float16)
                emb_dev = torch.from_numpy(emb).to(device, non_blocking=False)
                with torch.no_grad():
                    p_dev = head.infer_embeddings(emb_dev, temperature=float(temperature))
                parts.append(p_dev[start_idx - read_start:end_idx - read_start].cpu().numpy())
                del emb_dev, p_dev
            probs = np.concatenate(parts) if len(parts) > 1 else parts[0]
        if len(probs) != total_frames:
            print(f"Warning: Prediction count ({len(probs)}) != Frame count ({total_frames}).")
        pd.DataFrame(probs, columns=behaviors).to_csv(output_file, index=False)
        return output_file
    except Exception as e:
        print(f"Error during buffered inference on {file_path}: {e}")
        traceback.print_exc()
        return None


# ----------------------------------------------------------------------------------------------- training
def train_lstm_model(*args, **kwargs):
    """Head training behind the reference's name and signature (cbas.py:1274-1422); see cbas_b200.training."""
    from .training import train_lstm_model as _train
    return _train(*args, **kwargs)


# ----------------------------------------------------------------------------------------------- actogram
class Actogram:
    """Numeric part of cbas.Actogram (cbas.py:958-1007): `binned_activity`, `binsize_frames`.  Same constructor
    arguments; the PNG rendering (`blob`) is GUI work and stays None.  Events and bin sums are computed by the
    CUDA kernel when a GPU is present."""

    def __init__(self, behavior: str, framerate: float, start: float, binsize_minutes: int, threshold: float,
                 lightcycle: str, plot_acrophase: bool = False, base_color: str = None, directory: str = None,
                 model: str = None, preloaded_df=None):
        self.behavior = behavior
        self.framerate, self.start_hour_on_plot = float(framerate), float(start)
        self.threshold, self.bin_size_minutes = float(threshold), int(binsize_minutes)
        self.plot_acrophase = plot_acrophase
        self.lightcycle_str = {"LL": "1" * 24, "DD": "0" * 24}.get(lightcycle, "1" * 12 + "0" * 12)
        self.blob = None
        self.binned_activity = []
        if self.framerate <= 0 or self.bin_size_minutes <= 0:
            return
        self.binsize_frames = int(self.bin_size_minutes * self.framerate * 60)
        if self.binsize_frames <= 0:
            return
        import pandas as pd
        tables = []
        if preloaded_df is not None:
            if self.behavior in preloaded_df.columns:
                tables.append(preloaded_df)
        elif directory and model:
            csvs = [os.path.join(directory, f) for f in os.listdir(directory) if f.endswith(f"_{model}_outputs.csv")]
            if not csvs:
                return
            try:
                csvs.sort(key=lambda p: int(re.search(r"_(\d+)_" + model, os.path.basename(p)).group(1)))
            except (AttributeError, ValueError):
                csvs.sort()
            for p in csvs:
                df = pd.read_csv(p)
                if df.empty or self.behavior not in df.columns:
                    continue
                tables.append(df)
        else:
            return
        if not tables:
            return
        # every table contributes its own frames; the other-behaviour maximum is taken inside each table's columns
        events_parts = []
        for df in tables:
            cols = list(df.columns)
            probs = np.ascontiguousarray(df.to_numpy(dtype=np.float64))  # compared in float64, like the reference
            events_parts.append((probs, cols.index(self.behavior)))
        self.binned_activity = [float(v) for v in _bin_tables(events_parts, self.threshold, self.binsize_frames)]


def _bin_tables(parts, threshold: float, bin_frames: int) -> np.ndarray:
    """Concatenate the per-file event streams and sum them in bins of bin_frames (last partial bin kept)."""
    if len({(p.shape[1], b) for p, b in parts}) == 1 and torch.cuda.is_available():
        probs = np.concatenate([p for p, _ in parts]) if len(parts) > 1 else parts[0][0]
        return actogram_bins(torch.from_numpy(probs).cuda(), parts[0][1], threshold, bin_frames).cpu().numpy()
    if not torch.cuda.is_available():
        raise RuntimeError("cbas_b200.Actogram needs a CUDA device (no CPU fallback)")
    # files with different column sets: per-file events on the GPU (bin size 1), then the bin sums
    ev = [actogram_bins(torch.from_numpy(p).cuda(), b, threshold, 1) for p, b in parts]
    ev = torch.cat(ev).to(torch.float32)
    n = ev.numel()
    nb = (n + bin_frames - 1) // bin_frames
    pad = torch.zeros(nb * bin_frames - n, device=ev.device)
    return torch.cat([ev, pad]).view(nb, bin_frames).sum(1).to(torch.int64).cpu().numpy()
