"""CPU oracle for the cbas_b200 hot path.  TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import this
package, and only as the checker or the timed CPU baseline.  Nothing under cbas_b200/ imports it; the product
path fails loudly if libcbas_b200.so is missing.

Parity pinning: the reference ships no tests, golden vectors or fixtures, so this oracle is
pinned against outputs of the reference itself:
  * head:    oracle/gen_golden.py imports /root/reference/backend/classifier_head.py unchanged and stores its
             outputs under tests/golden/ (head_tiny.npz, head_default.npz); tests/test_oracle.py checks this
             restatement against them (and live against the reference module when /root/reference exists).
  * encoder: the ViT arithmetic lives in third-party `transformers` (requirements.txt:26 `transformers>=4.53.3`;
             5.5.0 installed here and on the GPU box).  oracle/encoder.py calls that very implementation
             (DINOv3ViTModel, random-init from a seed) around a restatement of the reference preprocessing
             (cbas.py:431,672-677); tests/golden/encoder_vits.npz pins its output for a seeded model.
  * windows / actogram: restated from cbas.py:497-551 and cbas.py:969-999; cbas.py itself cannot be imported
             (decord, h5py, matplotlib absent), so these two are pinned by closed-form cases in
             tests/test_oracle.py only -> "parity unpinned" against live reference outputs for those rows.
"""
