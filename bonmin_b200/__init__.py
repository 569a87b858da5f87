"""bonmin_b200 — B200-native batched NLP-relaxation engine behind Bonmin's plugin interfaces.

Python here is plumbing (ctypes over the C ABI in include/bnlp.h, instance generators, host-side
mirrors used by tests and bench.py).  All arithmetic of the hot path runs in libbnlp.so (CUDA, sm_100a)."""
