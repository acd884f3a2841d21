"""ctypes mirror of CML::Optimization::DSOTracer over the C ABI of include/cmltrc.h (SURVEY.md 8f NEXT #2).

Method names follow the reference class (optimization/dso/DSOTracer.h:34-206):
    makeNewTracesFrom(frame, corners)        -> DSOTracer.cpp:538-583
    traceNewCoarse(frameToTrace)             -> DSOTracer.cpp:17-60 (trace(): 585-832)
    optimizeImmaturePoint(points, minObs)    -> DSOTracer.cpp:280-411
Frames are registered with addFrame (the reference reads them from the Map's frame group).  No CPU fallback.
"""
import ctypes as C

import numpy as np

from .binding import CmlbaError, load_library

STATUS = ["GOOD", "OOB", "OUTLIER", "SKIPPED", "BADCONDITION", "UNINITIALIZED"]


class TracerConfig(C.Structure):
    _fields_ = [("min_idepth_h_act", C.c_float), ("gn_iterations", C.c_int), ("huber_threshold", C.c_float), ("outlier_th", C.c_float),
                ("outlier_th_sum_component", C.c_float), ("max_pix_search", C.c_float), ("max_slack_interval", C.c_float), ("trace_step_size", C.c_float),
                ("min_improvement_factor", C.c_float), ("min_trace_test_radius", C.c_float), ("extra_slack_on_th", C.c_float)]


POINT = np.dtype([("status", "<i4"), ("host_frame_slot", "<i4"), ("idepth_min", "<f8"), ("idepth_max", "<f8"), ("last_trace_uv", "<f8", 2),
                  ("last_trace_pixel_interval", "<f8"), ("quality", "<f8"), ("grad_h", "<f8", 4), ("energy_th", "<f8")])
ACTIVATION = np.dtype([("rc", "<i4"), ("idepth", "<f4"), ("in_mask", "<u4")])


class ActivateStats(C.Structure):
    _fields_ = [("current_minimum_distance", C.c_double), ("urgently_need_new_points", C.c_int32), ("num_deleted_outlier", C.c_int32), ("num_deleted_oob", C.c_int32),
                ("num_skipped_status", C.c_int32), ("num_skipped_pixel_interval", C.c_int32), ("num_skipped_quality", C.c_int32), ("num_skipped_depth", C.c_int32),
                ("num_to_optimize", C.c_int32), ("num_mapped", C.c_int32), ("num_non_mapped", C.c_int32), ("num_dropped", C.c_int32)]

TRACER_SYMBOLS = ["cmltrc_default_config", "cmltrc_create", "cmltrc_destroy", "cmltrc_last_error", "cmltrc_add_frame", "cmltrc_add_frame_device", "cmltrc_set_frame_pose", "cmltrc_remove_frame",
                  "cmltrc_make_new_traces", "cmltrc_remove_points", "cmltrc_num_points", "cmltrc_trace_new_coarse", "cmltrc_optimize_immature", "cmltrc_get_points", "cmltrc_activate_points", "cmltrc_set_minimum_distance"]

_bound = False


def _bind(lib):
    global _bound
    if _bound:
        return lib
    vp, dp, fp, i64 = C.c_void_p, C.POINTER(C.c_double), C.POINTER(C.c_float), C.c_int64
    lib.cmltrc_default_config.argtypes = [C.POINTER(TracerConfig)]
    lib.cmltrc_create.argtypes = [C.POINTER(TracerConfig), C.c_int, C.c_int, C.c_int, C.c_double, C.c_double, C.c_double, C.c_double, C.POINTER(vp)]
    lib.cmltrc_destroy.argtypes = [vp]
    lib.cmltrc_last_error.restype = C.c_char_p
    lib.cmltrc_last_error.argtypes = [vp]
    lib.cmltrc_add_frame.argtypes = [vp, i64, fp, dp, dp]
    lib.cmltrc_add_frame_device.argtypes = [vp, i64, vp, vp, dp, dp]
    lib.cmltrc_set_frame_pose.argtypes = [vp, i64, dp, dp]
    lib.cmltrc_remove_frame.argtypes = [vp, i64]
    lib.cmltrc_make_new_traces.argtypes = [vp, i64, C.c_int, fp, C.POINTER(i64)]
    lib.cmltrc_remove_points.argtypes = [vp, C.c_int, C.POINTER(i64)]
    lib.cmltrc_num_points.restype = i64
    lib.cmltrc_num_points.argtypes = [vp]
    lib.cmltrc_trace_new_coarse.argtypes = [vp, i64, C.POINTER(C.c_int32), fp]
    lib.cmltrc_optimize_immature.argtypes = [vp, C.c_int, C.POINTER(i64), C.c_int, vp, fp]
    lib.cmltrc_get_points.argtypes = [vp, i64, C.c_int, vp]
    lib.cmltrc_activate_points.argtypes = [vp, i64, C.c_int, dp, C.c_int, C.c_float, C.c_int, C.POINTER(i64), fp, C.c_int, C.POINTER(i64), vp, C.POINTER(C.c_int32),
                                           C.POINTER(i64), C.POINTER(C.c_int32), C.POINTER(ActivateStats)]
    lib.cmltrc_set_minimum_distance.argtypes = [vp, C.c_double]
    _bound = True
    return lib


def _dp(a):
    return a.ctypes.data_as(C.POINTER(C.c_double))


class DSOTracer:
    """One handle = one DSOTracer instance bound to an image size and a pinhole calibration."""

    def __init__(self, width, height, calib, device=0, **params):
        self.lib = _bind(load_library())
        cfg = TracerConfig()
        self.lib.cmltrc_default_config(C.byref(cfg))
        for k, v in params.items():
            if not hasattr(cfg, k):
                raise KeyError(f"unknown tracer parameter {k}")
            setattr(cfg, k, v)
        self.cfg = cfg
        self.width, self.height = int(width), int(height)
        self.h = C.c_void_p()
        fx, fy, cx, cy = [float(v) for v in calib]
        rc = self.lib.cmltrc_create(C.byref(cfg), device, self.width, self.height, fx, fy, cx, cy, C.byref(self.h))
        if rc != 0:
            raise CmlbaError(rc, self.lib.cmltrc_last_error(None).decode())
        self.last_gpu_ms = 0.0

    def close(self):
        if getattr(self, "h", None):
            self.lib.cmltrc_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _ck(self, rc):
        if rc != 0:
            raise CmlbaError(rc, self.lib.cmltrc_last_error(self.h).decode())

    def addFrame(self, frame_id, gray, camera, exposure):
        g = np.ascontiguousarray(gray, dtype=np.float32)
        if g.shape != (self.height, self.width):
            raise ValueError(f"gray image must be [{self.height}][{self.width}]")
        cam = np.ascontiguousarray(camera, dtype=np.float64).reshape(12); ex = np.ascontiguousarray(exposure, dtype=np.float64).reshape(3)
        self._ck(self.lib.cmltrc_add_frame(self.h, int(frame_id), g.ctypes.data_as(C.POINTER(C.c_float)), _dp(cam), _dp(ex)))

    def addFrameDevice(self, frame_id, capture, camera, exposure):
        """addFrame from a device-resident CaptureImage (libcml_b200.imgprep): level-0 gray and texels are copied device to device."""
        cam = np.ascontiguousarray(camera, dtype=np.float64).reshape(12); ex = np.ascontiguousarray(exposure, dtype=np.float64).reshape(3)
        self._ck(self.lib.cmltrc_add_frame_device(self.h, int(frame_id), C.c_void_p(capture.devicePtr("gray0")), C.c_void_p(capture.devicePtr("texel0")), _dp(cam), _dp(ex)))

    def setFramePose(self, frame_id, camera, exposure):
        cam = np.ascontiguousarray(camera, dtype=np.float64).reshape(12); ex = np.ascontiguousarray(exposure, dtype=np.float64).reshape(3)
        self._ck(self.lib.cmltrc_set_frame_pose(self.h, int(frame_id), _dp(cam), _dp(ex)))

    def removeFrame(self, frame_id):
        self._ck(self.lib.cmltrc_remove_frame(self.h, int(frame_id)))

    # ---- DSOTracer::makeNewTracesFrom(frame, group)
    def makeNewTracesFrom(self, frame_id, corners):
        xy = np.ascontiguousarray(corners, dtype=np.float32).reshape(-1, 2)
        first = C.c_int64()
        self._ck(self.lib.cmltrc_make_new_traces(self.h, int(frame_id), xy.shape[0], xy.ctypes.data_as(C.POINTER(C.c_float)), C.byref(first)))
        return np.arange(first.value, first.value + xy.shape[0], dtype=np.int64)

    def removePoints(self, ids):
        a = np.ascontiguousarray(ids, dtype=np.int64)
        self._ck(self.lib.cmltrc_remove_points(self.h, a.size, a.ctypes.data_as(C.POINTER(C.c_int64))))

    def numPoints(self):
        return int(self.lib.cmltrc_num_points(self.h))

    # ---- DSOTracer::traceNewCoarse(frameToTrace, frameGroup)
    def traceNewCoarse(self, frame_id):
        """Returns the status histogram of this pass (good, oob, outlier, skipped, badcondition, uninitialized)."""
        hist = (C.c_int32 * 6)(); ms = C.c_float()
        self._ck(self.lib.cmltrc_trace_new_coarse(self.h, int(frame_id), hist, C.byref(ms)))
        self.last_gpu_ms = ms.value
        return np.array(hist)

    # ---- DSOTracer::optimizeImmaturePoint(point, minObs, frameGroup), for a batch of points
    def optimizeImmaturePoint(self, ids, minObs=1):
        a = np.ascontiguousarray(ids, dtype=np.int64)
        out = np.zeros(a.size, dtype=ACTIVATION); ms = C.c_float()
        self._ck(self.lib.cmltrc_optimize_immature(self.h, a.size, a.ctypes.data_as(C.POINTER(C.c_int64)), int(minObs), out.ctypes.data_as(C.c_void_p), C.byref(ms)))
        self.last_gpu_ms = ms.value
        return out

    # ---- DSOTracer::activatePoints(frameGroup, pointGroup)
    def activatePoints(self, last_frame_id, active_xy, immature_ids, desiredPointDensity=800, types=None, minTraceQuality=3.0):
        """active_xy [A][2]: the active points projected into the last frame; immature_ids: the immature points in the caller's iteration order.
        Returns (activated ids, activation records (rc == 1), removed ids, stats); activated and removed points leave the immature set."""
        axy = np.ascontiguousarray(active_xy, dtype=np.float64).reshape(-1, 2)
        ids = np.ascontiguousarray(immature_ids, dtype=np.int64)
        ty = None if types is None else np.ascontiguousarray(types, dtype=np.float32)
        cap = max(ids.size, 1)
        a_ids = np.zeros(cap, np.int64); act = np.zeros(cap, ACTIVATION); r_ids = np.zeros(cap, np.int64)
        na, nr, st = C.c_int32(), C.c_int32(), ActivateStats()
        self._ck(self.lib.cmltrc_activate_points(self.h, int(last_frame_id), axy.shape[0], _dp(axy), int(desiredPointDensity), float(minTraceQuality), ids.size,
                                                 ids.ctypes.data_as(C.POINTER(C.c_int64)), None if ty is None else ty.ctypes.data_as(C.POINTER(C.c_float)), cap,
                                                 a_ids.ctypes.data_as(C.POINTER(C.c_int64)), act.ctypes.data_as(C.c_void_p), C.byref(na),
                                                 r_ids.ctypes.data_as(C.POINTER(C.c_int64)), C.byref(nr), C.byref(st)))
        return a_ids[:na.value].copy(), act[:na.value].copy(), r_ids[:nr.value].copy(), st

    def setMinimumDistance(self, v):
        self._ck(self.lib.cmltrc_set_minimum_distance(self.h, float(v)))

    def getPoints(self, first=0, count=None):
        n = self.numPoints() - first if count is None else count
        out = np.zeros(n, dtype=POINT)
        self._ck(self.lib.cmltrc_get_points(self.h, int(first), int(n), out.ctypes.data_as(C.c_void_p)))
        return out
