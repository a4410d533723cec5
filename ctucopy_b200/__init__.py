"""ctucopy_b200 -- B200-native hot path of CtuCopy behind a C ABI.

The product is `libctucopy_b200.so` (hand-written sm_100a kernels + `extern "C"` entry
points declared in include/ctucopy_b200.h) and the `ctucopy_b200` command-line host.
This package is the thin ctypes binding used by the tests, bench.py and Python callers.
There is no CPU implementation here: importing works anywhere, creating a Handle needs
a CUDA device.
"""
from .api import (  # noqa: F401
    Config,
    CtuError,
    Handle,
    Plan,
    Result,
    design_filter_bank,
    extract,
    extract_features,
    extract_g711,
    g711_table,
    lib,
    lib_path,
    parse_config,
    vad_debug_files,
)

__all__ = ["Config", "CtuError", "Handle", "Plan", "Result", "design_filter_bank", "extract", "extract_features", "extract_g711", "g711_table", "lib", "lib_path", "parse_config", "vad_debug_files"]
