"""Work scheduling for the hot path: mirrors of `EncodeThread` (workthreads.py:267-362) and
`ClassificationThread` (workthreads.py:365-519) without the GUI transport.

Each thread pops file paths from the shared lists in `gui_state` under their locks, runs
`cbas.encode_file` / `cbas.infer_file` inside its own CUDA stream, logs and skips files that fail, and the encode
thread hands finished `_cls.h5` files to the classifier when a live-inference model is set
(workthreads.py:323-328).  eel progress callbacks are replaced by optional Python callables.

Multi-GPU: start one EncodeThread/ClassificationThread pair per device - the queues are shared,
so videos are distributed dynamically and no collective is involved - or run one process per GPU (bench.py,
cbas_b200.launch).  In the one-process form each device needs its own encoder: put them in
`gui_state.dino_encoders` keyed by device string; every native handle remembers its device and makes it current for
the duration of a call, so the threads need no device bookkeeping of their own.
"""
from __future__ import annotations

import os
import threading
import time
import traceback
from datetime import datetime
from typing import Callable, Optional

import torch

from . import bundle, cbas, gui_state

print_lock = threading.Lock()


def log_message(message: str, level: str = "INFO") -> None:
    """Timestamped, lock-protected print (workthreads.py:74-96; the GUI log queue is out of scope)."""
    with print_lock:
        print(f"[{datetime.now().strftime('%H:%M:%S')}] [{level}] {message}", flush=True)


def fit_temperature(model, val_loader, device):
    """Temperature calibration of a trained head (workthreads.py:103-137); see cbas_b200.training."""
    from .training import fit_temperature as _fit
    return _fit(model, val_loader, device)


class EncodeThread(threading.Thread):
    def __init__(self, device_str: str = "cuda", progress: Optional[Callable[[dict], None]] = None,
                 poll_seconds: float = 0.2):
        super().__init__(daemon=True)
        self.device = torch.device(device_str)
        self.cuda_stream = torch.cuda.Stream(device=self.device) if self.device.type == "cuda" else None
        self.progress = progress
        self.poll = poll_seconds
        self.total_initial_tasks = 0
        self.tasks_processed_in_batch = 0
        self._stop_flag = threading.Event()

    def stop(self):
        self._stop_flag.set()

    def _encoder(self):
        """This thread's encoder: the one registered for its device, else the single shared one."""
        return gui_state.dino_encoders.get(str(self.device)) or gui_state.dino_encoder

    def run(self):
        while not self._stop_flag.is_set():
            if self._encoder() is None:
                time.sleep(self.poll)
                continue
            with gui_state.encode_lock:
                file_to_encode = gui_state.encode_tasks.pop(0) if gui_state.encode_tasks else None
            if not file_to_encode:
                self.tasks_processed_in_batch = 0
                time.sleep(self.poll)
                continue
            name = os.path.basename(file_to_encode)
            last = [-10]

            def progress_updater(percent, _name=name):
                if self.progress:
                    self.progress({"overall_processed": self.tasks_processed_in_batch,
                                   "current_percent": percent, "current_file": _name})
                if int(percent) // 10 > last[0] // 10:
                    last[0] = int(percent)
                    log_message(f"Encoding '{_name}': {int(percent)}%", "INFO")

            try:
                if self.cuda_stream is not None:
                    with torch.cuda.stream(self.cuda_stream):
                        out_file = cbas.encode_file(self._encoder(), file_to_encode, progress_updater)
                else:
                    out_file = cbas.encode_file(self._encoder(), file_to_encode, progress_updater)
                if out_file:
                    log_message(f"Finished encoding: {name}", "INFO")
                    if gui_state.live_inference_model_name:
                        with gui_state.classify_lock:
                            gui_state.classify_tasks.append(out_file)
                else:
                    raise RuntimeError("Encoder returned None (likely video read error or empty video).")
            except Exception as e:
                log_message(f"Failed to encode '{name}': {e}", "ERROR")
            finally:
                self.tasks_processed_in_batch += 1


class ClassificationThread(threading.Thread):
    def __init__(self, device_str: str = "cuda", model_dirs: Optional[dict] = None, poll_seconds: float = 0.2,
                 on_new_data: Optional[Callable[[str], None]] = None):
        super().__init__(daemon=True)
        self.device = torch.device(device_str)
        self.cuda_stream = torch.cuda.Stream(device=self.device) if self.device.type == "cuda" else None
        self.model_dirs = model_dirs or {}  # model name -> bundle directory (gui_state.proj.models in the reference)
        self.poll = poll_seconds
        self.on_new_data = on_new_data
        self._stop_flag = threading.Event()

    def stop(self):
        self._stop_flag.set()

    def _load_model(self, model_name):
        log_message(f"Loading model bundle '{model_name}'...", "INFO")
        try:
            model_dir = self.model_dirs.get(model_name)
            if model_dir is None and gui_state.proj is not None and hasattr(gui_state.proj, "models"):
                model_obj = gui_state.proj.models.get(model_name)
                model_dir = getattr(model_obj, "path", None)
            if not model_dir:
                raise ValueError(f"Model '{model_name}' not found.")
            project_encoder = getattr(gui_state.proj, "encoder_model_identifier", None) if gui_state.proj else None
            in_features = getattr(gui_state.dino_encoder, "hidden_size", 768) if gui_state.dino_encoder else 768
            model, meta = bundle.load_model_bundle(model_dir, project_encoder, self.device, in_features)
            gui_state.live_inference_model_object = model
            log_message(f"Model '{model_name}' loaded successfully.", "INFO")
            return model, meta
        except Exception as e:
            log_message(f"Error loading model bundle '{model_name}': {e}", "ERROR")
            traceback.print_exc()
            gui_state.live_inference_model_object = None
            return None, None

    def run(self):
        last_model_name = None
        torch_model, meta = None, None
        while not self._stop_flag.is_set():
            model_name = gui_state.live_inference_model_name
            if not model_name:
                last_model_name = None
                time.sleep(self.poll)
                continue
            if model_name != last_model_name:
                torch_model, meta = self._load_model(model_name)
                last_model_name = model_name
                if torch_model is None:
                    gui_state.live_inference_model_name = None
                    continue
            with gui_state.classify_lock:
                file_to_classify = gui_state.classify_tasks.pop(0) if gui_state.classify_tasks else None
            if not file_to_classify:
                time.sleep(self.poll)
                continue
            try:
                hp = meta["hyperparameters"]
                args = dict(file_path=file_to_classify, model=torch_model, dataset_name=model_name,
                            behaviors=hp["behaviors"], seq_len=hp.get("seq_len", 31), device=self.device,
                            temperature=float(meta.get("calibration", {}).get("temperature", 1.0)))
                if self.cuda_stream is not None:
                    with torch.cuda.stream(self.cuda_stream):
                        out = cbas.infer_file(**args)
                else:
                    out = cbas.infer_file(**args)
                if out:
                    log_message(f"Finished classifying: {os.path.basename(file_to_classify)}", "INFO")
                    if self.on_new_data:
                        self.on_new_data(out)
            except Exception as e:
                log_message(f"Failed to classify '{os.path.basename(file_to_classify)}': {e}", "ERROR")


def start_threads(device_str: str = "cuda", model_dirs: Optional[dict] = None):
    """Start one encode and one classification worker on `device_str` (workthreads.py:1245-1273)."""
    enc, cls = EncodeThread(device_str), ClassificationThread(device_str, model_dirs)
    enc.start()
    cls.start()
    return enc, cls
