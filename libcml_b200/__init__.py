"""libcml_b200 -- B200-native photometric bundle adjustment (the DSO sliding-window hot path of libCML).

The product is libcmlba.so (hand-written sm_100a CUDA behind the C ABI of include/cmlba.h).  This package
holds its sources (csrc/), a ctypes binding that mirrors CML::Optimization::DSOBundleAdjustment's public
methods (binding.py), the window-snapshot format (cmlw.py) and the synthetic-window generator (synth.py).
There is no CPU fallback: without the compiled extension / a CUDA device every call raises.
"""
from . import cmlw, synth  # noqa: F401
from .binding import DSOBundleAdjustment, CmlbaError, load_library, lib_path  # noqa: F401
from .tracker import DSOTracker  # noqa: F401
from .tracer import DSOTracer  # noqa: F401
from .imgprep import CaptureImageGenerator  # noqa: F401
from .selector import PixelSelector  # noqa: F401
from .fast import FAST  # noqa: F401
