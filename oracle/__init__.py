"""CPU oracle for the cbas_b200 hot path.  TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import this
package, and only as the checker or the timed CPU baseline.  Nothing under cbas_b200/ imports it; the product
path fails loudly if libcbas_b200.so is missing.

Parity pinning: the reference ships no tests, golden vectors or fixtures, so this oracle is
pinned against OUTPUTS OF THE REFERENCE ITSELF, produced in the authoring container by the committed scripts and
stored under tests/golden/ (the GPU box has no /root/reference; the fixtures travel instead):
  * oracle/gen_golden.py imports /root/reference/backend/cbas.py and classifier_head.py UNMODIFIED - only the
    absent I/O modules (decord, h5py, matplotlib) are replaced by in-memory stand-ins - and records
      encode_file            -> encode_file_vitb.npz   (DINOv3 ViT-B/16, the `_cls.h5` content, layout, attributes)
      ClassifierLSTMDeltas   -> head_tiny.npz, head_default.npz (forward outputs, smoothed / delta streams)
      infer_file             -> infer_file.npz         (the CSV probabilities of the real window loop)
      Actogram               -> actogram.npz           (bin counts)
  * oracle/gen_golden_head_variants.py -> head_variants.npz (lstm_hidden_size 128, two LSTM layers, no acceleration)
  * oracle/gen_golden_events.py        -> events.npz        (Dataset.predictions_to_instances[_with_confidence])
  * oracle/gen_golden_dinov2.py        -> encode_file_dinov2reg.npz (encode_file with a DINOv2-with-registers model)
tests/test_oracle.py checks every restatement in this package against those fixtures (and live against the reference
module when /root/reference exists).  The ViT arithmetic itself lives in third-party `transformers`
(requirements.txt:26 `transformers>=4.53.3`; 5.5.0 installed here and on the GPU box): oracle/encoder.py calls that very
implementation (DINOv3ViTModel / Dinov2WithRegistersModel, random-init from a seed - the gated hub weights are not
available offline) around a restatement of the reference preprocessing (cbas.py:431,672-677).
"""
