"""The C ABI on its own: a plain-C host program (examples/c_abi_head_demo.c, built here with gcc against
include/cbas_b200.h) drives libcbas_b200.so without Python or torch in the process; its output must equal what the
Python mirror gets from the same library and agree with the CPU oracle."""
import os
import shutil
import subprocess

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

from cbas_b200.classifier_head import ClassifierLSTMDeltas, actogram_bins  # noqa: E402
from oracle import head as ohead  # noqa: E402

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.parametrize("hs,layers,acc", [(64, 1, True), (128, 2, False)])
def test_plain_c_host_matches_python_mirror(tmp_path, hs, layers, acc):
    gcc = shutil.which("gcc")
    if gcc is None:
        pytest.skip("no C compiler on this box")
    cuda_lib = next((p for p in ("/usr/local/cuda/lib64", "/usr/local/cuda/targets/x86_64-linux/lib")
                     if os.path.exists(os.path.join(p, "libcudart.so"))), None)
    if cuda_lib is None:
        pytest.skip("libcudart.so not found")
    exe = str(tmp_path / "c_abi_head_demo")
    libdir = os.path.join(ROOT, "cbas_b200")
    subprocess.run([gcc, "-O2", "-Wall", "-I", os.path.join(ROOT, "include"), os.path.join(ROOT, "examples", "c_abi_head_demo.c"),
                    "-o", exe, "-L", libdir, "-lcbas_b200", "-L", cuda_lib, "-lcudart",
                    f"-Wl,-rpath,{libdir}", f"-Wl,-rpath,{cuda_lib}"], check=True)
    F, C, T, n, temp = 384, 7, 31, 700, 1.3
    sd = ohead.make_head_state(F, C, 128, hs, seed=77, scale=2.0, lstm_layers=layers, use_acceleration=acc)
    d = tmp_path / "data"
    d.mkdir()
    for k, v in sd.items():
        if k not in ("gate", "attention_temp"):
            v.numpy().astype(np.float32).tofile(str(d / f"{k}.f32"))
    np.array([float(sd["gate"]), float(sd["attention_temp"])], np.float32).tofile(str(d / "scalars.f32"))
    emb = (np.random.default_rng(5).standard_normal((n, F)) * 1.2).astype(np.float16)
    emb.tofile(str(d / "emb.f16"))
    (d / "cfg.txt").write_text(f"{F} {C} {T} {hs} {layers} {int(acc)} {n} {temp}\n")
    r = subprocess.run([exe, str(d)], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stderr + r.stdout
    assert r.stdout.startswith("ok:")
    probs = np.fromfile(str(d / "probs.f32"), np.float32).reshape(n, C)
    bins = np.fromfile(str(d / "bins.i32"), np.int32)

    head = ClassifierLSTMDeltas(F, C, seq_len=T, lstm_hidden_size=hs, lstm_layers=layers, use_acceleration=acc)
    head.load_state_dict(sd)
    head = head.to("cuda").eval()
    want = head.infer_embeddings(torch.from_numpy(emb).cuda(), temperature=temp)
    assert np.array_equal(probs, want.cpu().numpy())  # same library, same kernels: bitwise
    assert np.array_equal(bins, actogram_bins(want, 0, 0.1, 50).cpu().numpy())
    oracle = ohead.infer_windows(emb, sd, seq_len=T, temperature=temp)
    assert np.abs(probs - oracle).max() <= 1e-3 and (probs.argmax(1) == oracle.argmax(1)).mean() >= 0.999
