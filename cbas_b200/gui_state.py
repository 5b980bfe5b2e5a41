"""Shared state the hot path reads, mirroring the subset of backend/gui_state.py:31-104 that
encode_file / the worker threads touch: the loaded project (for the H5 attribute stamp, cbas.py:414-416),
the encoder object, and the two work queues with their locks (gui_state.py:76-84)."""
from __future__ import annotations

import threading
from typing import Any, List, Optional

proj: Optional[Any] = None                 # object with .encoder_model_identifier (cbas.Project in the reference)
dino_encoder: Optional[Any] = None         # cbas_b200.encoder.DinoEncoder once a project is loaded
# one encoder per device for the thread-pair-per-GPU deployment ("cuda:1" -> DinoEncoder on cuda:1); a worker thread
# takes the entry of its own device and falls back to `dino_encoder` (the reference has a single encoder object)
dino_encoders: dict = {}

encode_tasks: List[str] = []
encode_lock = threading.Lock()
classify_tasks: List[str] = []
classify_lock = threading.Lock()
live_inference_model_name: Optional[str] = None
live_inference_model_object: Optional[Any] = None

HEADLESS_MODE = True
