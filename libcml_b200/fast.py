"""ctypes mirror of CML::Features::FAST over the C ABI of include/cmlfast.h (SURVEY.md 8f NEXT #4: first unit of the ORB extractor).

    corners, scores = FAST(max_w, max_h).compute(image_u8, threshold)     -> features/corner/FAST.cpp:3-13
No CPU fallback.
"""
import ctypes as C

import numpy as np

from .binding import CmlbaError, load_library

FAST_SYMBOLS = ["cmlfast_create", "cmlfast_destroy", "cmlfast_last_error", "cmlfast_compute"]
_bound = False


def _bind(lib):
    global _bound
    if _bound:
        return lib
    vp = C.c_void_p
    lib.cmlfast_create.argtypes = [C.c_int, C.c_int, C.c_int, C.POINTER(vp)]
    lib.cmlfast_destroy.argtypes = [vp]
    lib.cmlfast_last_error.restype = C.c_char_p
    lib.cmlfast_last_error.argtypes = [vp]
    lib.cmlfast_compute.argtypes = [vp, C.POINTER(C.c_uint8), C.c_int, C.c_int, C.c_int, C.c_int, C.POINTER(C.c_int32), C.POINTER(C.c_int32), C.POINTER(C.c_int32),
                                    C.POINTER(C.c_float)]
    _bound = True
    return lib


class FAST:
    def __init__(self, max_width, max_height, device=0):
        self.lib = _bind(load_library())
        self.h = C.c_void_p()
        rc = self.lib.cmlfast_create(device, int(max_width), int(max_height), C.byref(self.h))
        if rc != 0:
            raise CmlbaError(rc, self.lib.cmlfast_last_error(None).decode())
        self.last_gpu_ms = 0.0

    def close(self):
        if getattr(self, "h", None):
            self.lib.cmlfast_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def compute(self, image, threshold, capacity=None):
        """image [h][w] uint8.  Returns (corners [n][2] int32 in raster order, scores [n] int32) = the Corner list with responses of the reference."""
        img = np.ascontiguousarray(image, dtype=np.uint8)
        h, w = img.shape
        cap = int(capacity if capacity is not None else (w * h) // 4)
        xy = np.empty((max(cap, 1), 2), np.int32); sc = np.empty(max(cap, 1), np.int32)
        n = C.c_int32(); ms = C.c_float()
        rc = self.lib.cmlfast_compute(self.h, img.ctypes.data_as(C.POINTER(C.c_uint8)), w, h, int(threshold), cap, xy.ctypes.data_as(C.POINTER(C.c_int32)),
                                      sc.ctypes.data_as(C.POINTER(C.c_int32)), C.byref(n), C.byref(ms))
        if rc != 0:
            raise CmlbaError(rc, self.lib.cmlfast_last_error(self.h).decode())
        self.last_gpu_ms = ms.value
        self.last_count = n.value
        k = min(n.value, cap)
        return xy[:k].copy(), sc[:k].copy()
