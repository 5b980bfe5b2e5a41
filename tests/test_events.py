"""Event extraction (cbas_b200/events.py) against instance lists produced by the reference's own
Dataset.predictions_to_instances[_with_confidence] (oracle/gen_golden_events.py -> tests/golden/events.npz)."""
import json
import os

import numpy as np
import pandas as pd
import pytest

from cbas_b200 import events
from oracle.gen_golden_events import make_probs


@pytest.fixture(scope="module")
def gold(golden_dir):
    return np.load(os.path.join(golden_dir, "events.npz"))


@pytest.fixture(scope="module")
def csv_path(tmp_path_factory, gold):
    p = make_probs(int(gold["seed"]), int(gold["n"]), len(gold["behaviors"]))
    path = str(tmp_path_factory.mktemp("ev") / "cam1_00001_m1_outputs.csv")
    pd.DataFrame(p, columns=[str(b) for b in gold["behaviors"]]).to_csv(path, index=False)
    return path


@pytest.mark.parametrize("thr", [0.7, 0.5, 0.95])
def test_instances_match_reference(gold, csv_path, thr):
    behaviors = [str(b) for b in gold["behaviors"]]
    want = json.loads(str(gold[f"inst_thr{thr}"]))
    got = events.predictions_to_instances(csv_path, "m1", behaviors, threshold=thr)
    for d in got:
        d["video"] = os.path.basename(d["video"])
    assert got == want
    # same result from the array, without the CSV round trip
    p = pd.read_csv(csv_path)[behaviors].to_numpy()
    got2 = events.predictions_to_instances(p, "m1", behaviors, threshold=thr, video="cam1_00001.mp4")
    assert got2 == want


@pytest.mark.parametrize("win", [1, 5, 8])
def test_instances_with_confidence_match_reference(gold, csv_path, win):
    behaviors = [str(b) for b in gold["behaviors"]]
    want = json.loads(str(gold[f"conf_win{win}"]))
    got, df = events.predictions_to_instances_with_confidence(csv_path, "m1", behaviors, smoothing_window=win,
                                                              project_path=os.path.dirname(csv_path))
    assert df is not None and "label_for_grouping" in df.columns
    assert len(got) == len(want)
    for g, w in zip(got, want):
        assert (g["video"], g["start"], g["end"], g["label"]) == (w["video"], w["start"], w["end"], w["label"])
        assert abs(g["confidence"] - w["confidence"]) <= 1e-12


def test_edge_cases(tmp_path):
    behaviors = ["a", "b"]
    assert events.predictions_to_instances(str(tmp_path / "missing.csv"), "m", behaviors) == []
    assert events.predictions_to_instances_with_confidence(str(tmp_path / "missing.csv"), "m", behaviors) == ([], None)
    bad = str(tmp_path / "x_m_outputs.csv")
    pd.DataFrame({"a": [0.9], "c": [0.1]}).to_csv(bad, index=False)
    assert events.predictions_to_instances(bad, "m", behaviors) == []          # a behaviour column is missing
    assert events.predictions_to_instances(np.zeros((0, 2)), "m", behaviors) == []
    one = events.predictions_to_instances(np.array([[0.9, 0.1]]), "m", behaviors, threshold=0.7, video="v.mp4")
    assert one == [{"video": "v.mp4", "start": 0, "label": "a", "end": 0}]
    # ties go to the first column (pandas idxmax), an event that runs to the last frame is closed there
    p = np.array([[0.5, 0.5], [0.8, 0.2], [0.2, 0.8], [0.2, 0.8]])
    assert events.predictions_to_instances(p, "m", behaviors, threshold=0.5, video="v") == [
        {"video": "v", "start": 0, "label": "a", "end": 1}, {"video": "v", "start": 2, "label": "b", "end": 3}]
