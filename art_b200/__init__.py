"""art_b200 -- B200-native raw-development hot path (drop-in for artpixls/ART's rtengine path).

The product is `libart_hotpath.so` (hand-written sm_100a CUDA behind the C-ABI in
include/art_hotpath.h).  This package is the host-side mirror of the reference's
C++ surface for that path (rtengine/rawimagesource.h, rtengine/improcfun.h),
bound to the library with ctypes.  There is no CPU fallback: importing works
anywhere (so the ABI can be inspected), computing without a GPU raises.
"""
from .api import (HotPath, HotPathError, lib_path, load_library, ABI_SYMBOLS,  # noqa: F401
                  BAYER_AMAZE, BAYER_RCD, XTRANS_3PASS, XTRANS_1PASS)
from .rawimagesource import RawImageSource  # noqa: F401
from .improcfun import ImProcFunctions  # noqa: F401
from .batchqueue import BatchQueue  # noqa: F401
from . import synth  # noqa: F401

__all__ = ["HotPath", "HotPathError", "RawImageSource", "BatchQueue", "lib_path", "load_library",
           "ABI_SYMBOLS", "BAYER_AMAZE", "BAYER_RCD", "XTRANS_3PASS", "XTRANS_1PASS", "synth"]
