"""fastoverlap_b200 -- B200-native (CUDA sm_100a, FP64) implementation of FASTOVERLAP's
overlap-maximisation hot path behind the reference's Python API.

Public names mirror reference fastoverlap/__init__.py:2-17.  The compute back end is
libfastoverlap_b200.so (include/fastoverlap_b200.h) loaded with ctypes; there is no CPU
fallback -- constructing a Context without the library or without a GPU raises
FastOverlapError.
"""
from ._lib import Context, FastOverlapError, default_context, load_library, library_path
from .periodic import PeriodicAlign
from .soft import SOFT
from .spherical import SphericalAlign, SphericalHarmonicAlign
from .wrappers import (SphericalAlignFortran, SphericalHarmonicAlignFortran,
                               PeriodicAlignFortran)

__all__ = ["SphericalAlign", "SphericalHarmonicAlign", "PeriodicAlign", "SOFT",
           "SphericalAlignFortran", "SphericalHarmonicAlignFortran", "PeriodicAlignFortran", "Context",
           "FastOverlapError", "default_context"]
