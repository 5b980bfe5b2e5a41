"""Golden vector for CBAS's default encoder family (DINOv2-with-registers) from the REAL reference encode_file.
TEST INFRASTRUCTURE.

Same stand-ins as oracle/gen_golden.py (decord / h5py / matplotlib only); the reference's own DinoEncoder loads a
random-init Dinov2WithRegistersModel saved with save_pretrained and encode_file runs unmodified (cbas.py:399-456).

    python -m oracle.gen_golden_dinov2     ->   tests/golden/encode_file_dinov2reg.npz
"""
import os
import tempfile
import types
from unittest import mock

import numpy as np

from oracle import gen_golden as gg


def main():
    gg._install_stubs()
    import cbas  # the reference
    import gui_state
    from oracle import encoder as oenc
    frames = oenc.synthetic_frames(5, 128, 128, seed=17)
    model = oenc.build_hf_dinov2_model("dinov2reg-b14", seed=4, init_scale=3.0, num_hidden_layers=6)  # the reference hard-codes D = 768 (cbas.py:677)
    with tempfile.TemporaryDirectory() as td:
        model.save_pretrained(td)
        enc = cbas.DinoEncoder(td, device="cpu")
    gui_state.proj = types.SimpleNamespace(encoder_model_identifier="synthetic:dinov2reg-b14")
    gg._VIDEOS["/mem/clip2.mp4"] = frames
    with mock.patch.object(cbas.os, "replace", gg._os_replace_h5), \
            mock.patch.object(cbas.os.path, "exists", lambda p: False):
        out_path = cbas.encode_file(enc, "/mem/clip2.mp4", None)
    cls = gg._H5[out_path]["datasets"]["cls"]
    np.savez_compressed(os.path.join(gg.OUT, "encode_file_dinov2reg.npz"), cls=cls, frames_seed=17, model_seed=4,
                        init_scale=3.0, layers=6, side=128)
    print(out_path, cls.shape, cls.dtype)


if __name__ == "__main__":
    main()
