"""CPU restatement of the numeric part of Actogram.__init__ (cbas.py:969-999).  TEST INFRASTRUCTURE.

Per frame:  event = (p_b * [max_{b' != b} p_b' < p_b]) >= threshold
Bins:       bins[k] = sum(events[k*bs : (k+1)*bs]),  bs = int(bin_minutes * framerate * 60), last partial bin kept.
"""
from __future__ import annotations

import numpy as np


def binsize_frames(bin_minutes: int, framerate: float) -> int:
    return int(int(bin_minutes) * float(framerate) * 60)


def actogram_bins(probs: np.ndarray, behavior: int, threshold: float, bin_frames: int) -> np.ndarray:
    """probs [N,C] float -> int64 counts [ceil(N/bin_frames)].  Comparison arithmetic follows pandas/numpy in
    the reference: strict `<` against the max of the other columns, product with the 0/1 mask, `>=` threshold.
    With a single behaviour column the reference's max over zero columns is NaN and `NaN < p` is False."""
    probs = np.asarray(probs)
    n, c = probs.shape
    if bin_frames <= 0 or n == 0:
        return np.zeros(0, np.int64)
    p = probs[:, behavior]
    if c > 1:
        # DataFrame.max(axis=1) skips NaN cells (skipna=True); a row of nothing but NaN stays NaN and `NaN < p` is False
        others = np.fmax.reduce(np.delete(probs, behavior, axis=1), axis=1)
        is_max = others < p
    else:
        is_max = np.zeros(n, bool)
    events = (p * is_max >= threshold).astype(np.float64)
    nb = (n + bin_frames - 1) // bin_frames
    return np.array([events[k * bin_frames:(k + 1) * bin_frames].sum() for k in range(nb)]).astype(np.int64)
