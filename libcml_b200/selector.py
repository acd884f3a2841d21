"""ctypes mirror of CML::Features::PixelSelector over the C ABI of include/cmlsel.h (SURVEY.md 8f NEXT #4, PixelSelector part).

    sel = PixelSelector(w, h); corners, types = sel.compute(capture, density)      -> PixelSelector.cpp:367-384
`capture` is a device-resident CaptureImage (libcml_b200.imgprep).  No CPU fallback.
"""
import ctypes as C

import numpy as np

from .binding import CmlbaError, load_library

SEL_SYMBOLS = ["cmlsel_create", "cmlsel_destroy", "cmlsel_last_error", "cmlsel_set_potential", "cmlsel_get_potential", "cmlsel_compute", "cmlsel_read"]
_bound = False


def _bind(lib):
    global _bound
    if _bound:
        return lib
    vp, fp = C.c_void_p, C.POINTER(C.c_float)
    lib.cmlsel_create.argtypes = [C.c_int, C.c_int, C.c_int, C.POINTER(vp)]
    lib.cmlsel_destroy.argtypes = [vp]
    lib.cmlsel_last_error.restype = C.c_char_p
    lib.cmlsel_last_error.argtypes = [vp]
    lib.cmlsel_set_potential.argtypes = [vp, C.c_int]
    lib.cmlsel_get_potential.argtypes = [vp]
    lib.cmlsel_compute.argtypes = [vp, C.POINTER(vp), C.c_float, C.c_int, C.c_float, C.c_int, fp, fp, C.POINTER(C.c_int32), fp]
    lib.cmlsel_read.restype = C.c_int64
    lib.cmlsel_read.argtypes = [vp, C.c_char_p, vp, C.c_int64]
    _bound = True
    return lib


class PixelSelector:
    def __init__(self, width, height, device=0):
        self.lib = _bind(load_library())
        self.width, self.height = int(width), int(height)
        self.h = C.c_void_p()
        rc = self.lib.cmlsel_create(device, self.width, self.height, C.byref(self.h))
        if rc != 0:
            raise CmlbaError(rc, self.lib.cmlsel_last_error(None).decode())
        self.last_gpu_ms = 0.0

    def close(self):
        if getattr(self, "h", None):
            self.lib.cmlsel_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _ck(self, rc):
        if rc != 0:
            raise CmlbaError(rc, self.lib.cmlsel_last_error(self.h).decode())

    def setPotential(self, v):
        self._ck(self.lib.cmlsel_set_potential(self.h, int(v)))

    @property
    def currentPotential(self):
        return int(self.lib.cmlsel_get_potential(self.h))

    def compute(self, capture, density, recursionsLeft=1, thFactor=1.0, capacity=None):
        """Returns (corners [n][2] float32, types [n] float32) in the reference's emission order."""
        if capture.gen.sizes[0] != (self.width, self.height):
            raise ValueError("capture size differs from the selector's")
        lv = (C.c_void_p * 3)(*[capture.devicePtr(f"texel{l}") for l in range(3)])
        cap = int(capacity or (self.width * self.height) // 4)
        xy = np.empty((cap, 2), np.float32); ty = np.empty(cap, np.float32)
        n = C.c_int32(); ms = C.c_float()
        self._ck(self.lib.cmlsel_compute(self.h, lv, float(density), int(recursionsLeft), float(thFactor), cap, xy.ctypes.data_as(C.POINTER(C.c_float)),
                                         ty.ctypes.data_as(C.POINTER(C.c_float)), C.byref(n), C.byref(ms)))
        self.last_gpu_ms = ms.value
        k = min(n.value, cap)
        return xy[:k].copy(), ty[:k].copy()

    def read(self, name, shape):
        out = np.empty(shape, np.float32)
        n = self.lib.cmlsel_read(self.h, name.encode(), out.ctypes.data_as(C.c_void_p), out.nbytes)
        if n < 0:
            raise CmlbaError(int(n), self.lib.cmlsel_last_error(self.h).decode())
        return out
