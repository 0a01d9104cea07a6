"""ppcseq_b200 -- B200-native hot path of stemangiola/ppcseq.

The product is `libppcseq_b200.so` (hand-written sm_100a CUDA behind a C ABI, include/ppcseq_b200.h).
This package is the thin host-side mirror of the reference's interface for that path; it binds the
library with ctypes and FAILS LOUDLY when the library is missing -- there is no CPU fallback.
"""
from ._lib import lib, library_path, PpcseqError  # noqa: F401
from .model import NBModel, layout  # noqa: F401
from .fit import Fit  # noqa: F401

__all__ = ["lib", "library_path", "PpcseqError", "NBModel", "layout", "Fit"]
