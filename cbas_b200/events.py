"""Event extraction from per-frame probabilities: the step right after the head.

Mirrors `Dataset.predictions_to_instances` (backend/cbas.py:903-929) and
`Dataset.predictions_to_instances_with_confidence` (backend/cbas.py:931-956): same arguments, same dictionaries,
same edge cases - but the per-row pandas loop is replaced by run-length encoding on arrays, so a 1 M-frame file
takes milliseconds instead of minutes.  Both accept either the CSV `infer_file` wrote or the probability array
itself (e.g. straight from `ClassifierLSTMDeltas.infer_embeddings`, still on the GPU), so events and actogram
bins can be produced without a CSV round trip.
"""
from __future__ import annotations

import os
from typing import List, Optional, Sequence, Tuple, Union

import numpy as np

ArrayOrPath = Union[str, np.ndarray, "torch.Tensor"]  # noqa: F821


def _probs(source: ArrayOrPath, behaviors: Sequence[str]):
    """-> (float64 [N, len(behaviors)], DataFrame or None).  CSV columns are selected by behaviour name like the
    reference (`df[behaviors]`); arrays are taken to be in `behaviors` order."""
    if isinstance(source, str):
        import pandas as pd
        df = pd.read_csv(source)
        if not behaviors or any(b not in df.columns for b in behaviors):
            return None, df
        return df[list(behaviors)].to_numpy(dtype=np.float64), df
    if hasattr(source, "detach"):
        source = source.detach().float().cpu().numpy()
    a = np.asarray(source, dtype=np.float64)
    if a.ndim != 2 or a.shape[1] != len(behaviors):
        raise ValueError(f"expected probabilities of shape [N, {len(behaviors)}], got {a.shape}")
    return a, None


def _labels(p: np.ndarray) -> Tuple[np.ndarray, np.ndarray]:
    """pandas `idxmax(axis=1)` / `max(axis=1)`: first maximum wins, NaN entries are skipped."""
    if p.shape[0] == 0:
        return np.zeros(0, np.int64), np.zeros(0, np.float64)
    q = np.where(np.isnan(p), -np.inf, p)
    lab = q.argmax(axis=1)
    return lab, p[np.arange(len(p)), lab]


def predictions_to_instances(source: ArrayOrPath, model_name: str, behaviors: Sequence[str], threshold: float = 0.7,
                             video: Optional[str] = None) -> List[dict]:
    """cbas.py:903-929.  An event is a maximal run of frames whose top probability is >= threshold and whose top
    behaviour does not change; `{"video", "start", "label", "end"}` with inclusive frame indices."""
    try:
        p, _ = _probs(source, behaviors)
    except FileNotFoundError:
        return []
    if p is None or not len(behaviors):
        return []
    if video is None:
        video = source.replace(f"_{model_name}_outputs.csv", ".mp4") if isinstance(source, str) else ""
    lab, mx = _labels(p)
    n = len(lab)
    if n == 0:
        return []
    above = mx >= threshold
    # a new event starts at every above-threshold frame that follows a below-threshold frame or a label change
    prev_above = np.concatenate([[False], above[:-1]])
    prev_lab = np.concatenate([[-1], lab[:-1]])
    starts = np.flatnonzero(above & (~prev_above | (prev_lab != lab)))
    # it ends at the last frame before the next start or the next below-threshold frame
    nxt_break = np.concatenate([~above[1:] | (lab[1:] != lab[:-1]), [True]])
    ends = np.flatnonzero(above & nxt_break)
    return [{"video": video, "start": int(s), "label": behaviors[int(lab[s])], "end": int(e)}
            for s, e in zip(starts, ends)]


def predictions_to_instances_with_confidence(source: ArrayOrPath, model_name: str, behaviors: Sequence[str],
                                             threshold: float = 0.5, smoothing_window: int = 1,
                                             project_path: Optional[str] = None, video: Optional[str] = None):
    """cbas.py:931-956.  Every frame belongs to a block (no thresholding; `threshold` is accepted and unused, as in
    the reference): blocks are maximal runs of the (optionally median-filtered) top behaviour, confidence = mean top
    probability over the block.  Returns (instances, DataFrame or None) like the reference."""
    try:
        p, df = _probs(source, behaviors)
    except FileNotFoundError:
        return [], None
    if p is None or not len(behaviors):
        return [], df
    lab, mx = _labels(p)
    n = len(lab)
    if smoothing_window > 1:
        if smoothing_window % 2 == 0:
            smoothing_window += 1
        from scipy.signal import medfilt
        grp = medfilt(lab, kernel_size=smoothing_window).astype(np.int64)  # zero-padded median, like the reference
    else:
        grp = lab
    if video is None:
        video = source.replace(f"_{model_name}_outputs.csv", ".mp4") if isinstance(source, str) else ""
    if project_path is not None and video:
        video = os.path.relpath(video, start=project_path).replace("\\", "/")
    if df is not None:  # the reference annotates the frame it returns
        df["predicted_label"] = [behaviors[i] for i in lab]
        df["max_prob"] = mx
        df["label_for_grouping"] = [behaviors[i] for i in grp]
    if n == 0:
        return [], df
    starts = np.flatnonzero(np.concatenate([[True], grp[1:] != grp[:-1]]))
    ends = np.concatenate([starts[1:] - 1, [n - 1]])
    csum = np.concatenate([[0.0], np.cumsum(mx)])
    out = []
    for s, e in zip(starts, ends):
        # mean over the block exactly as pandas does it (sum / count), not from the running sum, to stay within
        # an ulp of the reference for long files
        conf = float(mx[s:e + 1].mean()) if e - s < 4096 else float((csum[e + 1] - csum[s]) / (e - s + 1))
        out.append({"video": video, "start": int(s), "end": int(e), "label": behaviors[int(grp[s])],
                    "confidence": conf})
    return out, df
