"""Self-describing model bundle: `model.pth` + `config.yaml` + `model_meta.json`.

Mirrors the save side of `workthreads.py:856-886` and the load side of `workthreads.py:372-451`
(ClassificationThread._load_model): metadata-driven architecture selection, legacy fallback when
model_meta.json is absent, encoder-identifier check, inference of `lstm_hidden_size` / `lstm_layers` from the
weight shapes when the metadata lacks them, `load_state_dict(strict=False)`.
"""
from __future__ import annotations

import json
import os
from typing import Any, Dict, List, Optional, Tuple

import torch
import yaml

from .classifier_head import ClassifierLSTMDeltas

BUNDLE_SCHEMA = "1.0"


class EncoderMismatch(RuntimeError):
    pass


def save_model_bundle(model_dir: str, model: ClassifierLSTMDeltas, name: str, behaviors: List[str], seq_len: int,
                      encoder_model_identifier: str, temperature: float = 1.0, cbas_commit_hash: str = "unknown",
                      training_run_info: Optional[Dict[str, Any]] = None) -> None:
    """Write the three bundle files exactly as the reference's trainer does (workthreads.py:856-886)."""
    os.makedirs(model_dir, exist_ok=True)
    torch.save({k: v.detach().cpu() for k, v in model.state_dict().items()}, os.path.join(model_dir, "model.pth"))
    config = {"name": name, "behaviors": list(behaviors), "seq_len": int(seq_len), "architecture": type(model).__name__}
    with open(os.path.join(model_dir, "config.yaml"), "w") as f:
        yaml.dump(config, f, allow_unicode=True)
    meta = {
        "model_bundle_schema": BUNDLE_SCHEMA,
        "cbas_commit_hash": cbas_commit_hash,
        "encoder_model_identifier": encoder_model_identifier,
        "head_architecture_version": type(model).__name__,
        "hyperparameters": {
            "behaviors": list(behaviors),
            "seq_len": int(seq_len),
            "use_acceleration": bool(getattr(model, "use_acceleration", True)),
            "lstm_hidden_size": int(getattr(model.lstm, "hidden_size", 64)),
            "lstm_layers": int(getattr(model.lstm, "num_layers", 1)),
        },
        "training_run_info": dict(training_run_info or {}),
        "calibration": {"temperature": float(temperature)},
    }
    with open(os.path.join(model_dir, "model_meta.json"), "w") as f:
        json.dump(meta, f, indent=4)


def load_model_bundle(model_dir: str, project_encoder_identifier: Optional[str] = None, device="cuda",
                      in_features: int = 768) -> Tuple[ClassifierLSTMDeltas, Dict[str, Any]]:
    """Load a bundle and return (head in eval mode on `device`, metadata) - workthreads.py:372-451.

    Raises EncoderMismatch when the bundle was trained on another encoder than the project's (the reference
    logs the error and refuses the model), and NotImplementedError for the legacy (pre-delta) head, which v3's
    infer_file cannot drive either."""
    cfg_path = os.path.join(model_dir, "config.yaml")
    config = {}
    if os.path.exists(cfg_path):
        with open(cfg_path) as f:
            config = yaml.safe_load(f) or {}
    meta_path = os.path.join(model_dir, "model_meta.json")
    if not os.path.exists(meta_path):
        meta = {"head_architecture_version": "ClassifierLegacyLSTM", "hyperparameters": dict(config),
                "encoder_model_identifier": project_encoder_identifier}
    else:
        with open(meta_path) as f:
            meta = json.load(f)
    model_encoder = meta.get("encoder_model_identifier")
    if model_encoder and project_encoder_identifier and model_encoder != project_encoder_identifier:
        raise EncoderMismatch(f"Encoder mismatch! Project is for '{project_encoder_identifier}', but model was trained "
                              f"with '{model_encoder}'. Please re-encode videos.")
    arch = meta.get("head_architecture_version", "ClassifierLegacyLSTM")
    hp = meta.get("hyperparameters", {}) or {}
    if "behaviors" not in hp:
        hp["behaviors"] = config.get("behaviors", [])
    if "seq_len" not in hp:
        hp["seq_len"] = config.get("seq_len", 31)
    meta["hyperparameters"] = hp
    weights_path = os.path.join(model_dir, "model.pth")
    try:
        weights = torch.load(weights_path, map_location="cpu", weights_only=True)
    except TypeError:
        weights = torch.load(weights_path, map_location="cpu")
    if not arch.startswith("ClassifierLSTMDeltas"):
        raise NotImplementedError(f"head architecture '{arch}' (legacy) is not supported by the B200 inference path")
    if "lstm_hidden_size" not in hp:
        ref = weights.get("attention_head.weight", weights.get("lin2.weight"))
        inferred = ref.shape[1] // 2 if ref is not None else 0
        hp["lstm_hidden_size"] = inferred if inferred else 64
    if "lstm_layers" not in hp:
        keys = [int(k.split("weight_ih_l")[1].split("_")[0]) for k in weights if "lstm.weight_ih_l" in k]
        hp["lstm_layers"] = max(keys) + 1 if keys else 1
    model = ClassifierLSTMDeltas(in_features=in_features, out_features=len(hp["behaviors"]), seq_len=hp["seq_len"],
                                 lstm_hidden_size=hp["lstm_hidden_size"], lstm_layers=hp["lstm_layers"])
    model.load_state_dict(weights, strict=False)
    model.to(device).eval()
    return model, meta
