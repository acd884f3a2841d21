"""ctypes mirror of CML::Optimization::DSOTracker over the C ABI of include/cmltrk.h (SURVEY.md 8f NEXT #1).

Method names and argument meaning follow the reference class (optimization/dso/DSOTracker.h:198-470):
    makeCoarseDepthL0(reference, points)                       -> DSOTracker.cpp:494-725
    optimize(numTry, frameToTrack, reference, camera, exposure) -> DSOTracker.cpp:15-246, returns a Residual
Frames and points are passed as plain arrays (the reference passes PFrame / PointSet objects).
There is no CPU fallback: without libcmlba.so or a CUDA device the constructor raises.
"""
import ctypes as C

import numpy as np

from .binding import CmlbaError, load_library

OPT_LEVELS = 5
MAX_CANDIDATES = 32


class TrackerConfig(C.Structure):
    _fields_ = [("huber_threshold", C.c_double), ("cutoff_threshold", C.c_double), ("scale_rotation", C.c_double), ("scale_translation", C.c_double),
                ("scale_light_a", C.c_double), ("scale_light_b", C.c_double), ("optimize_a", C.c_int), ("optimize_b", C.c_int),
                ("saturated_ratio_threshold", C.c_double), ("levels", C.c_int), ("cluster_ctas", C.c_int), ("cta_threads", C.c_int)]


class TrackerResult(C.Structure):
    _fields_ = [("cam", C.c_double * 12), ("affine", C.c_double * 2), ("E", C.c_double * OPT_LEVELS), ("num_terms_in_E", C.c_int32 * OPT_LEVELS),
                ("num_saturated", C.c_int32 * OPT_LEVELS), ("num_robust", C.c_int32 * OPT_LEVELS), ("level_cutoff_repeat", C.c_double * OPT_LEVELS),
                ("flow_vector", C.c_double * 3), ("rel_aff", C.c_double * 2), ("covariance", C.c_double * 6), ("is_correct", C.c_int32),
                ("too_many_saturated", C.c_int32), ("iterations", C.c_int32), ("levels_used", C.c_int32), ("gpu_ms", C.c_float), ("kernel_launches", C.c_int32)]


TRACKER_SYMBOLS = ["cmltrk_default_config", "cmltrk_create", "cmltrk_destroy", "cmltrk_last_error", "cmltrk_make_coarse_depth", "cmltrk_set_frame", "cmltrk_optimize",
                   "cmltrk_track", "cmltrk_read", "cmltrk_bench_optimize", "cmltrk_frame_buffer", "cmltrk_make_coarse_depth_device", "cmltrk_set_frame_device",
                   "cmltrk_track_with_motion_model", "cmltrk_reset_motion_model", "cmltrk_motion_model_state"]

_bound = False


def _bind(lib):
    global _bound
    if _bound:
        return lib
    vp, dp, fp = C.c_void_p, C.POINTER(C.c_double), C.POINTER(C.c_float)
    lib.cmltrk_default_config.argtypes = [C.POINTER(TrackerConfig)]
    lib.cmltrk_create.argtypes = [C.POINTER(TrackerConfig), C.c_int, C.c_int, C.c_int, C.c_double, C.c_double, C.c_double, C.c_double, C.POINTER(vp)]
    lib.cmltrk_destroy.argtypes = [vp]
    lib.cmltrk_last_error.restype = C.c_char_p
    lib.cmltrk_last_error.argtypes = [vp]
    lib.cmltrk_make_coarse_depth.argtypes = [vp, fp, dp, dp, C.c_int, dp, C.c_int, C.POINTER(C.c_int32), fp, dp, dp]
    lib.cmltrk_set_frame.argtypes = [vp, fp, C.c_double]
    lib.cmltrk_make_coarse_depth_device.argtypes = [vp, C.c_int, C.POINTER(vp), dp, dp, C.c_int, dp, C.c_int, C.POINTER(C.c_int32), fp, dp, dp]
    lib.cmltrk_set_frame_device.argtypes = [vp, C.c_int, C.POINTER(vp), C.c_double]
    lib.cmltrk_optimize.argtypes = [vp, C.c_int, dp, dp, dp, C.POINTER(TrackerResult)]
    lib.cmltrk_track.argtypes = [vp, fp, C.c_double, C.c_int, dp, dp, dp, C.POINTER(TrackerResult)]
    lib.cmltrk_read.restype = C.c_int64
    lib.cmltrk_read.argtypes = [vp, C.c_char_p, vp, C.c_int64]
    lib.cmltrk_bench_optimize.argtypes = [vp, C.c_int, fp]
    lib.cmltrk_track_with_motion_model.argtypes = [vp, C.c_int, dp, dp, C.c_int, C.POINTER(C.c_int), C.POINTER(TrackerResult), C.POINTER(C.c_int)]
    lib.cmltrk_reset_motion_model.argtypes = [vp]
    lib.cmltrk_motion_model_state.argtypes = [vp, dp, dp]
    lib.cmltrk_frame_buffer.restype = fp
    lib.cmltrk_frame_buffer.argtypes = [vp]
    _bound = True
    return lib


def _dp(a):
    return a.ctypes.data_as(C.POINTER(C.c_double))


def _fp(a):
    return a.ctypes.data_as(C.POINTER(C.c_float))


class Residual:
    """DSOTracker::Residual (DSOTracker.h:202-236) + the optimised camera / exposure parameters of one start pose."""

    def __init__(self, r):
        n = r.levels_used
        self.camera = np.array(r.cam)                       # world-to-camera [R | t]; the start pose when not isCorrect
        self.exposure = np.array(r.affine)                  # (a, b)
        self.E = np.array(r.E)[:n]
        self.numTermsInE = np.array(r.num_terms_in_E)[:n]
        self.numSaturated = np.array(r.num_saturated)[:n]
        self.numRobust = np.array(r.num_robust)[:n]
        self.levelCutoffRepeat = np.array(r.level_cutoff_repeat)[:n]
        self.flowVector = np.array(r.flow_vector)
        self.relAff = np.array(r.rel_aff)
        self.covariance = np.array(r.covariance)
        self.isCorrect = bool(r.is_correct)
        self.tooManySaturated = bool(r.too_many_saturated)
        self.iterations = r.iterations
        self.gpu_ms = r.gpu_ms
        self.kernel_launches = r.kernel_launches

    def rmse(self, i=0):
        assert self.numTermsInE[i] > 0, "Invalid residual"
        return self.E[i] / self.numTermsInE[i]

    def saturatedRatio(self, i=0):
        assert self.numTermsInE[i] > 0, "Invalid residual"
        return self.numSaturated[i] / self.numTermsInE[i]


class DSOTracker:
    """One handle = one DSOTracker instance bound to an image size and a pinhole calibration."""

    def __init__(self, width, height, calib, device=0, **params):
        self.lib = _bind(load_library())
        cfg = TrackerConfig()
        self.lib.cmltrk_default_config(C.byref(cfg))
        for k, v in params.items():
            if not hasattr(cfg, k):
                raise KeyError(f"unknown tracker parameter {k}")
            setattr(cfg, k, v)
        self.cfg = cfg
        self.width, self.height = int(width), int(height)
        self.h = C.c_void_p()
        fx, fy, cx, cy = [float(v) for v in calib]
        rc = self.lib.cmltrk_create(C.byref(cfg), device, self.width, self.height, fx, fy, cx, cy, C.byref(self.h))
        if rc != 0:
            raise CmlbaError(rc, self.lib.cmltrk_last_error(None).decode())
        self.mLastResidual = None

    def close(self):
        if getattr(self, "h", None):
            self.lib.cmltrk_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _ck(self, rc):
        if rc != 0:
            raise CmlbaError(rc, self.lib.cmltrk_last_error(self.h).decode())

    def _gray(self, gray):
        g = np.ascontiguousarray(gray, dtype=np.float32)
        if g.shape != (self.height, self.width):
            raise ValueError(f"gray image must be [{self.height}][{self.width}]")
        return g

    # ---- DSOTracker::makeCoarseDepthL0(reference, points)
    def makeCoarseDepthL0(self, ref_gray, ref_camera, ref_exposure, frame_cameras, pt_frame, pt_xy, pt_idepth, pt_uncertainty):
        """reference = (ref_gray [H][W], ref_camera [12] world-to-camera, ref_exposure (time, a, b)); points = host frame index into
        frame_cameras [F][12], pixel in the host frame, inverse depth, uncertainty."""
        g = self._gray(ref_gray)
        rc_ = np.ascontiguousarray(ref_camera, dtype=np.float64).reshape(12)
        re_ = np.ascontiguousarray(ref_exposure, dtype=np.float64).reshape(3)
        fc = np.ascontiguousarray(frame_cameras, dtype=np.float64).reshape(-1, 12)
        pf = np.ascontiguousarray(pt_frame, dtype=np.int32)
        xy = np.ascontiguousarray(pt_xy, dtype=np.float32).reshape(-1, 2)
        idp = np.ascontiguousarray(pt_idepth, dtype=np.float64)
        unc = np.ascontiguousarray(pt_uncertainty, dtype=np.float64)
        if not (pf.size == xy.shape[0] == idp.size == unc.size):
            raise ValueError("point arrays differ in length")
        self._ck(self.lib.cmltrk_make_coarse_depth(self.h, _fp(g), _dp(rc_), _dp(re_), fc.shape[0], _dp(fc), pf.size, pf.ctypes.data_as(C.POINTER(C.c_int32)), _fp(xy),
                                                   _dp(idp), _dp(unc)))

    def makeCoarseDepthL0Device(self, ref_capture, ref_camera, ref_exposure, frame_cameras, pt_frame, pt_xy, pt_idepth, pt_uncertainty):
        """makeCoarseDepthL0 with the reference keyframe's gray levels taken from a device-resident CaptureImage (libcml_b200.imgprep)."""
        L = ref_capture.getPyramidLevels()
        lv = (C.c_void_p * L)(*[ref_capture.devicePtr(f"gray{l}") for l in range(L)])
        rc_ = np.ascontiguousarray(ref_camera, dtype=np.float64).reshape(12)
        re_ = np.ascontiguousarray(ref_exposure, dtype=np.float64).reshape(3)
        fc = np.ascontiguousarray(frame_cameras, dtype=np.float64).reshape(-1, 12)
        pf = np.ascontiguousarray(pt_frame, dtype=np.int32)
        xy = np.ascontiguousarray(pt_xy, dtype=np.float32).reshape(-1, 2)
        idp = np.ascontiguousarray(pt_idepth, dtype=np.float64)
        unc = np.ascontiguousarray(pt_uncertainty, dtype=np.float64)
        self._ck(self.lib.cmltrk_make_coarse_depth_device(self.h, L, lv, _dp(rc_), _dp(re_), fc.shape[0], _dp(fc), pf.size, pf.ctypes.data_as(C.POINTER(C.c_int32)), _fp(xy),
                                                          _dp(idp), _dp(unc)))

    def setFrameDevice(self, capture, exposure_time=1.0):
        """Frame to track = the texel levels of a device-resident CaptureImage, sampled in place (valid until the generator's next generate())."""
        L = capture.getPyramidLevels()
        lv = (C.c_void_p * L)(*[capture.devicePtr(f"texel{l}") for l in range(L)])
        self._ck(self.lib.cmltrk_set_frame_device(self.h, L, lv, float(exposure_time)))

    def frameBuffer(self):
        """The handle's page-locked staging image as a numpy view: fill it in place and pass it as `gray` to skip the host-side copy."""
        p = self.lib.cmltrk_frame_buffer(self.h)
        return np.ctypeslib.as_array(p, shape=(self.height, self.width))

    def setFrame(self, gray, exposure_time=1.0):
        """Uploads the frame to track and builds its pyramid (the CaptureImage role)."""
        g = self._gray(gray)
        self._ck(self.lib.cmltrk_set_frame(self.h, _fp(g), float(exposure_time)))

    def _start(self, cameras, exposures):
        cams = np.ascontiguousarray(cameras, dtype=np.float64).reshape(-1, 12)
        aff = np.ascontiguousarray(exposures, dtype=np.float64).reshape(-1, 2)
        if cams.shape[0] != aff.shape[0] or not (1 <= cams.shape[0] <= MAX_CANDIDATES):
            raise ValueError("need 1..32 start poses with one (a, b) each")
        last = None
        if self.mLastResidual is not None and self.mLastResidual.isCorrect:
            last = np.array([self.mLastResidual.rmse(l) if l < len(self.mLastResidual.E) else 0.0 for l in range(OPT_LEVELS)])
        return cams, aff, last

    # ---- DSOTracker::optimize(numTry, frameToTrack, reference, camera&, exposure&)
    def optimize(self, cameras, exposures, gray=None, exposure_time=1.0):
        """cameras [K][12] start poses (world-to-camera), exposures [K][2] start (a, b).  With `gray` the frame is uploaded in the same call
        (cmltrk_track).  Returns a list of K Residual (a single Residual when one start pose was given as a flat array)."""
        single = np.ndim(cameras) == 1
        cams, aff, last = self._start(cameras, exposures)
        res = (TrackerResult * cams.shape[0])()
        lp = _dp(last) if last is not None else None
        if gray is not None:
            g = self._gray(gray)
            self._ck(self.lib.cmltrk_track(self.h, _fp(g), float(exposure_time), cams.shape[0], _dp(cams), _dp(aff), lp, res))
        else:
            self._ck(self.lib.cmltrk_optimize(self.h, cams.shape[0], _dp(cams), _dp(aff), lp, res))
        out = [Residual(r) for r in res]
        return out[0] if single else out

    # ---- DSOTracker::trackWithMotionModel (DSOTracker.h:240-360) through the C ABI (cmltrk_track_with_motion_model)
    def trackWithMotionModel(self, cameras, initial_exposure=(0.0, 0.0), failure_mode=0):
        """cameras [K][12] = the poses of Map::multiConstantVelocityMotionModel, tried in order on the frame given to setFrame / setFrameDevice.
        Returns (ok, camera, exposure, residual) like the reference (frame->setCamera / setExposureParameters / residual)."""
        cams = np.ascontiguousarray(cameras, dtype=np.float64).reshape(-1, 12)
        init = np.ascontiguousarray(initial_exposure, dtype=np.float64).reshape(2)
        ok, tried, res = C.c_int(0), C.c_int(0), TrackerResult()
        self._ck(self.lib.cmltrk_track_with_motion_model(self.h, cams.shape[0], _dp(cams), _dp(init), int(failure_mode), C.byref(ok), C.byref(res), C.byref(tried)))
        self.lastTriedCameras = tried.value
        a, b = C.c_double(), C.c_double()
        self.lib.cmltrk_motion_model_state(self.h, C.byref(a), C.byref(b))
        self.mLastCoarseRMSE, self.mFirstRMSE = a.value, b.value
        r = Residual(res)
        if not ok.value:
            return False, None, None, (r if r.isCorrect else None)
        return True, r.camera, r.exposure, r

    # the same loop in Python on top of cmltrk_optimize (kept as the executable specification the C entry point is tested against)
    def trackWithMotionModelPy(self, cameras, initial_exposure=(0.0, 0.0), failure_mode=0):
        """cameras [K][12] = the poses of Map::multiConstantVelocityMotionModel, tried in order on the frame given to setFrame / setFrameDevice.
        Returns (ok, camera, exposure, residual) like the reference (frame->setCamera / setExposureParameters / residual); the candidates are
        optimised one after the other because each run is gated by the best residual so far (mLastResidual = trackingResult) and the loop
        stops at the first candidate that is good enough (achievedRes < 1.5 * mLastCoarseRMSE)."""
        cams = np.ascontiguousarray(cameras, dtype=np.float64).reshape(-1, 12)
        init = np.asarray(initial_exposure, dtype=np.float64).reshape(2)
        have, best, camera, exposure = False, None, None, None
        achieved = float("inf")
        self.lastTriedCameras = 0
        for i, cam in enumerate(cams):
            self.lastTriedCameras = i + 1
            self.mLastResidual = best
            test = self.optimize(cam, init)
            test_ok = test.isCorrect and test.numTermsInE[0] > 0 and np.isfinite(test.rmse())
            best_sat = True if best is None else best.tooManySaturated              # Residual() default: tooManySaturated = true
            if best_sat and not test.tooManySaturated and test_ok:
                have, camera, exposure, best = True, test.camera, test.exposure, test
            if test_ok and not (test.rmse() >= achieved):
                if (True if best is None else best.tooManySaturated) or not test.tooManySaturated:
                    have, camera, exposure, best = True, test.camera, test.exposure, test
            if have and test.numTermsInE[0] > 0 and test.rmse() < achieved:
                achieved = test.rmse()
            if have and achieved < getattr(self, "mLastCoarseRMSE", float("inf")) * 1.5:
                break
            if have and i >= 50:
                break
        if not have:
            if failure_mode != 1:
                return False, None, None, best
            self.mLastResidual = best
            best = self.optimize(cams[0], init)
            camera, exposure = best.camera, best.exposure
            return True, camera, exposure, best
        self.mLastCoarseRMSE = achieved
        if getattr(self, "mFirstRMSE", -1.0) < 0:
            self.mFirstRMSE = achieved
        return True, camera, exposure, best

    def benchOptimize(self, repeats=20):
        ms = C.c_float()
        self._ck(self.lib.cmltrk_bench_optimize(self.h, int(repeats), C.byref(ms)))
        return ms.value

    def read(self, name, dtype, max_bytes=None):
        cap = max_bytes or (self.width * self.height * 16 + 4096)
        buf = np.empty(cap, dtype=np.uint8)
        n = self.lib.cmltrk_read(self.h, name.encode(), buf.ctypes.data_as(C.c_void_p), cap)
        if n < 0:
            raise CmlbaError(int(n), self.lib.cmltrk_last_error(self.h).decode())
        return buf[:n].view(dtype).copy()
