"""cbas_b200 - B200-native (sm_100a) implementation of CBAS's streamed video encoder and behaviour-inference
path, behind the reference's own Python surface (backend/cbas.py, backend/classifier_head.py,
backend/workthreads.py).  All GPU work happens in libcbas_b200.so (hand-written CUDA, C ABI in
include/cbas_b200.h); there is no CPU fallback."""

__version__ = "0.1.0"
