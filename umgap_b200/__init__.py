"""umgap_b200: B200-native implementation of UMGAP's per-read classification hot path.

The product is ``umgap_b200/lib/libumgap_gpu.so`` (hand-written CUDA for sm_100a behind the C ABI of
``include/umgap_gpu.h``) and the ``umgap`` CLI built on it; :mod:`umgap_b200.capi` is the ctypes
mirror used by the tests and the benchmark.  There is no CPU fallback.
"""
__version__ = "0.1.0"
