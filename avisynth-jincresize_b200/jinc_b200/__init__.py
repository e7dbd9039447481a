"""Python-side helpers of the B200 EWA-Jinc resampler: ctypes bindings for the C ABI
(include/jinc_b200.h) and the mini-host driver used by tests and bench.py.
The product itself is native: csrc/ (CUDA + C ABI) and plugin/ (AviSynth+ C plugin)."""
