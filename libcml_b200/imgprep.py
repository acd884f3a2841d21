"""ctypes mirror of the per-frame part of CML::CaptureImageGenerator over the C ABI of include/cmlimg.h (SURVEY.md 8f NEXT #3).

    gen = CaptureImageGenerator(in_w, in_h, out_w, out_h); gen.setLut(lut); gen.setInverseVignette(v); gen.setUndistortMap(map)
    cap = gen.generate(raw)     -> CaptureImage.cpp:108-262
    cap.getGrayImage(l), cap.getDerivativeImage(l), cap.getWeightedGradientNorm(l)   (CaptureImage.h:40-104)
No CPU fallback: without libcmlba.so or a CUDA device the constructor raises.
"""
import ctypes as C

import numpy as np

from .binding import CmlbaError, load_library

IMG_SYMBOLS = ["cmlimg_create", "cmlimg_destroy", "cmlimg_last_error", "cmlimg_set_photometric", "cmlimg_set_undistort_map", "cmlimg_input_buffer", "cmlimg_prepare", "cmlimg_prepare_u8",
               "cmlimg_levels", "cmlimg_read", "cmlimg_device_ptr", "cmlimg_bench"]
_bound = False


def _bind(lib):
    global _bound
    if _bound:
        return lib
    vp, fp = C.c_void_p, C.POINTER(C.c_float)
    lib.cmlimg_create.argtypes = [C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.POINTER(vp)]
    lib.cmlimg_destroy.argtypes = [vp]
    lib.cmlimg_last_error.restype = C.c_char_p
    lib.cmlimg_last_error.argtypes = [vp]
    lib.cmlimg_set_photometric.argtypes = [vp, fp, fp]
    lib.cmlimg_set_undistort_map.argtypes = [vp, fp]
    lib.cmlimg_input_buffer.restype = fp
    lib.cmlimg_input_buffer.argtypes = [vp]
    lib.cmlimg_prepare.argtypes = [vp, fp, fp]
    lib.cmlimg_prepare_u8.argtypes = [vp, C.POINTER(C.c_uint8), fp]
    lib.cmlimg_levels.argtypes = [vp, C.POINTER(C.c_int32), C.POINTER(C.c_int32)]
    lib.cmlimg_read.restype = C.c_int64
    lib.cmlimg_read.argtypes = [vp, C.c_char_p, vp, C.c_int64]
    lib.cmlimg_device_ptr.restype = vp
    lib.cmlimg_device_ptr.argtypes = [vp, C.c_char_p]
    lib.cmlimg_bench.argtypes = [vp, C.c_int, C.c_int, fp]
    _bound = True
    return lib


def _fp(a):
    return a.ctypes.data_as(C.POINTER(C.c_float)) if a is not None else None


class CaptureImage:
    """View of the levels the last generate() left on the device (read back on demand)."""

    def __init__(self, gen):
        self.gen = gen

    def getPyramidLevels(self):
        return len(self.gen.sizes)

    def devicePtr(self, name):
        """Device address (int) of "gray<l>" / "texel<l>" for the device-resident entry points of the tracker, the tracer and the BA."""
        p = self.gen.lib.cmlimg_device_ptr(self.gen.h, name.encode())
        if not p:
            raise KeyError(name)
        return int(p)

    def _texel(self, level):
        w, h = self.gen.sizes[level]
        return self.gen._read(f"texel{level}", (h, w, 4))

    def getGrayImage(self, level):
        w, h = self.gen.sizes[level]
        return self.gen._read(f"gray{level}", (h, w))

    def getDerivativeImage(self, level):
        return self._texel(level)[:, :, :3]

    def getWeightedGradientNorm(self, level):
        return self._texel(level)[:, :, 3]


class CaptureImageGenerator:
    def __init__(self, in_width, in_height, out_width=None, out_height=None, levels=0, device=0):
        self.lib = _bind(load_library())
        self.in_size = (int(in_width), int(in_height))
        ow, oh = int(out_width or in_width), int(out_height or in_height)
        self.h = C.c_void_p()
        rc = self.lib.cmlimg_create(device, self.in_size[0], self.in_size[1], ow, oh, int(levels), C.byref(self.h))
        if rc != 0:
            raise CmlbaError(rc, self.lib.cmlimg_last_error(None).decode())
        n = C.c_int32(); wh = (C.c_int32 * 16)()
        self.lib.cmlimg_levels(self.h, C.byref(n), wh)
        self.sizes = [(wh[2 * l], wh[2 * l + 1]) for l in range(n.value)]
        self._lut = None; self._vig = None
        self.last_gpu_ms = 0.0

    def close(self):
        if getattr(self, "h", None):
            self.lib.cmlimg_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _ck(self, rc):
        if rc != 0:
            raise CmlbaError(rc, self.lib.cmlimg_last_error(self.h).decode())

    def _photometric(self):
        self._ck(self.lib.cmlimg_set_photometric(self.h, _fp(self._lut), _fp(self._vig)))

    def setLut(self, lut):
        self._lut = None if lut is None else np.ascontiguousarray(lut, dtype=np.float32).reshape(256)
        self._photometric()

    def setInverseVignette(self, inv_vignette):
        if inv_vignette is not None:
            v = np.ascontiguousarray(inv_vignette, dtype=np.float32)
            if v.shape != (self.in_size[1], self.in_size[0]):
                raise ValueError("inverse vignette must have the input image's shape")
            self._vig = v
        else:
            self._vig = None
        self._photometric()

    def setUndistortMap(self, undistort_map):
        m = None
        if undistort_map is not None:
            m = np.ascontiguousarray(undistort_map, dtype=np.float32)
            if m.shape != (self.sizes[0][1], self.sizes[0][0], 2):
                raise ValueError("undistortion map must be [out_height][out_width][2]")
        self._ck(self.lib.cmlimg_set_undistort_map(self.h, _fp(m)))

    def inputBuffer(self):
        return np.ctypeslib.as_array(self.lib.cmlimg_input_buffer(self.h), shape=(self.in_size[1], self.in_size[0]))

    def generate(self, raw):
        """raw: float32 (0..255) or uint8 sensor image [in_height][in_width]."""
        u8 = np.asarray(raw).dtype == np.uint8
        r = np.ascontiguousarray(raw, dtype=np.uint8 if u8 else np.float32)
        if r.shape != (self.in_size[1], self.in_size[0]):
            raise ValueError("raw image has the wrong shape")
        ms = C.c_float()
        if u8:
            self._ck(self.lib.cmlimg_prepare_u8(self.h, r.ctypes.data_as(C.POINTER(C.c_uint8)), C.byref(ms)))
        else:
            self._ck(self.lib.cmlimg_prepare(self.h, _fp(r), C.byref(ms)))
        self.last_gpu_ms = ms.value
        return CaptureImage(self)

    def bench(self, repeats=20, flush_l2=True):
        ms = C.c_float()
        self._ck(self.lib.cmlimg_bench(self.h, int(repeats), 1 if flush_l2 else 0, C.byref(ms)))
        return ms.value

    def _read(self, name, shape):
        out = np.empty(shape, dtype=np.float32)
        n = self.lib.cmlimg_read(self.h, name.encode(), out.ctypes.data_as(C.c_void_p), out.nbytes)
        if n < 0:
            raise CmlbaError(int(n), self.lib.cmlimg_last_error(self.h).decode())
        return out
