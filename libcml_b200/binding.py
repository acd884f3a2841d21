"""ctypes binding of libcmlba.so + a host-side mirror of the reference operator interface.

`DSOBundleAdjustment` below has the public methods of CML::Optimization::DSOBundleAdjustment
(reference: src/cml/optimization/dso/DSOBundleAdjustment.h:26-101) with the same names, argument meaning
and error behaviour (run() returns False where the reference returns false).  Frames and points are
identified by integer ids instead of CML's Ptr<Frame>/Ptr<MapPoint>; INTEGRATION.md shows the C++ adapter
that does the same translation inside CML.  Everything here is plumbing: the arithmetic lives in the CUDA
kernels of csrc/kernels.cuh.
"""
import ctypes as C
import os
import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))


def lib_path():
    # CMLBA_LIB: development aid (tuning builds of the same sources with other compile-time constants)
    return os.environ.get("CMLBA_LIB") or os.path.join(_HERE, "libcmlba.so")


class CmlbaError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(f"cmlba error {code}: {msg}")
        self.code = code


class Config(C.Structure):
    _fields_ = [("iterations", C.c_int), ("huber_threshold", C.c_float), ("outlier_th_sum", C.c_float), ("th_opt_iterations", C.c_float),
                ("scale_rotation", C.c_float), ("scale_translation", C.c_float), ("scale_light_a", C.c_float), ("scale_light_b", C.c_float),
                ("scale_f", C.c_float), ("scale_c", C.c_float), ("force_accept", C.c_int), ("fix_lambda", C.c_int), ("fixed_lambda", C.c_float),
                ("idepth_fix_prior", C.c_int), ("solver_mode_delta", C.c_float), ("optimize_light_a", C.c_int), ("optimize_light_b", C.c_int),
                ("disable_marginalization", C.c_int), ("max_frames", C.c_int), ("frame_min_age", C.c_int), ("min_idepth_h_marg", C.c_float), ("async_image_upload", C.c_int)]


class RunResult(C.Structure):
    _fields_ = [("iterations_done", C.c_int), ("num_residuals", C.c_int), ("num_dropped", C.c_int), ("num_outliers", C.c_int),
                ("energy_first", C.c_double), ("energy_last", C.c_double), ("gpu_ms", C.c_double), ("kernel_launches", C.c_int), ("num_rejected", C.c_int)]


class BenchResult(C.Structure):
    _fields_ = [("steps", C.c_int), ("residuals", C.c_int), ("points", C.c_int), ("frames", C.c_int), ("launches_per_pass", C.c_int),
                ("ms_pass", C.c_double), ("ms_linearize", C.c_double), ("ms_accumulate", C.c_double), ("ms_schur", C.c_double), ("ms_stitch", C.c_double),
                ("ms_assemble", C.c_double), ("ms_event_overhead", C.c_double)]


# every symbol include/cmlba.h declares (tests/test_abi.py checks the .so exports all of them)
SYMBOLS = ["cmlba_default_config", "cmlba_create", "cmlba_destroy", "cmlba_last_error", "cmlba_set_calib", "cmlba_add_frame", "cmlba_add_frame_gray", "cmlba_add_frame_device", "cmlba_add_points",
           "cmlba_remove_point", "cmlba_remove_frame", "cmlba_flag_frames_for_marginalization", "cmlba_try_marginalize", "cmlba_marginalize_points",
           "cmlba_marginalize_frames", "cmlba_run", "cmlba_num_frames", "cmlba_num_points", "cmlba_num_residuals", "cmlba_get_frames",
           "cmlba_get_points", "cmlba_get_outliers", "cmlba_get_residuals", "cmlba_get_statistics", "cmlba_statistic_name", "cmlba_prepare", "cmlba_linearize", "cmlba_apply", "cmlba_solve", "cmlba_step",
           "cmlba_read", "cmlba_reset", "cmlba_bench_pass", "cmlba_nccl_unique_id", "cmlba_comm_init", "cmlba_comm_ipc_handle", "cmlba_comm_ipc_open", "cmlba_version"]

_lib = None


def load_library():
    """Loads libcmlba.so (built by `make -C libcml_b200/csrc` or __graft_entry__.build()).  Fails loudly if absent."""
    global _lib
    if _lib is not None:
        return _lib
    p = lib_path()
    if not os.path.exists(p):
        raise ImportError(f"{p} is missing: build it with `make -C libcml_b200/csrc` (there is no CPU fallback)")
    lib = C.CDLL(p, mode=C.RTLD_GLOBAL)
    vp, dp, fp, ip = C.c_void_p, C.POINTER(C.c_double), C.POINTER(C.c_float), C.POINTER(C.c_int)
    lib.cmlba_version.restype = C.c_char_p
    lib.cmlba_last_error.restype = C.c_char_p
    lib.cmlba_last_error.argtypes = [vp]
    lib.cmlba_default_config.argtypes = [C.POINTER(Config)]
    lib.cmlba_create.argtypes = [C.POINTER(Config), C.c_int, C.POINTER(vp)]
    lib.cmlba_destroy.argtypes = [vp]
    lib.cmlba_set_calib.argtypes = [vp, C.c_double, C.c_double, C.c_double, C.c_double, C.c_int, C.c_int]
    lib.cmlba_add_frame.argtypes = [vp, C.c_int64, dp, C.c_double, C.c_double, C.c_double, fp, C.c_int]
    lib.cmlba_add_frame_gray.argtypes = [vp, C.c_int64, dp, C.c_double, C.c_double, C.c_double, fp, C.c_int]
    lib.cmlba_add_frame_device.argtypes = [vp, C.c_int64, dp, C.c_double, C.c_double, C.c_double, vp, C.c_int]
    lib.cmlba_add_points.argtypes = [vp, C.c_int, C.POINTER(C.c_int64), C.POINTER(C.c_int64), fp, dp]
    lib.cmlba_remove_point.argtypes = [vp, C.c_int64]
    lib.cmlba_remove_frame.argtypes = [vp, C.c_int64]
    lib.cmlba_run.argtypes = [vp, dp, C.c_int, C.c_int, C.POINTER(RunResult)]
    lib.cmlba_flag_frames_for_marginalization.argtypes = [vp, dp, ip, C.POINTER(C.c_int64), ip]
    lib.cmlba_try_marginalize.argtypes = [vp, ip, ip]
    lib.cmlba_marginalize_points.argtypes = [vp, C.POINTER(C.c_int64), ip]
    lib.cmlba_marginalize_frames.argtypes = [vp, C.POINTER(C.c_int64), ip]
    for f in ("cmlba_num_frames", "cmlba_num_points", "cmlba_num_residuals"):
        getattr(lib, f).argtypes = [vp]
    lib.cmlba_get_frames.argtypes = [vp, C.POINTER(C.c_int64), dp, dp, dp, dp, dp]
    lib.cmlba_get_points.argtypes = [vp, C.POINTER(C.c_int64), dp, dp, fp, fp, ip, ip]
    lib.cmlba_get_outliers.argtypes = [vp, C.POINTER(C.c_int64), ip]
    lib.cmlba_get_residuals.argtypes = [vp, C.POINTER(C.c_int64), C.POINTER(C.c_int64), ip, dp]
    lib.cmlba_prepare.argtypes = [vp, dp]
    lib.cmlba_linearize.argtypes = [vp, C.c_int, dp]
    lib.cmlba_apply.argtypes = [vp]
    lib.cmlba_solve.argtypes = [vp, C.c_int]
    lib.cmlba_step.argtypes = [vp, C.c_int, ip]
    lib.cmlba_read.argtypes = [vp, C.c_char_p, vp, C.c_size_t, C.POINTER(C.c_size_t)]
    lib.cmlba_reset.argtypes = [vp]
    lib.cmlba_bench_pass.argtypes = [vp, C.c_int, C.c_int, C.c_int, C.POINTER(BenchResult)]
    lib.cmlba_nccl_unique_id.argtypes = [vp]
    lib.cmlba_comm_init.argtypes = [vp, vp, C.c_int, C.c_int]
    lib.cmlba_comm_ipc_handle.argtypes = [vp, vp]
    lib.cmlba_comm_ipc_open.argtypes = [vp, vp]
    _lib = lib
    return lib


def _ptr(a, ctype):
    return a.ctypes.data_as(C.POINTER(ctype))


def default_config():
    cfg = Config()
    load_library().cmlba_default_config(C.byref(cfg))
    return cfg


class DSOBundleAdjustment:
    """Mirror of CML::Optimization::DSOBundleAdjustment over the C ABI (one handle = one BA instance)."""

    def __init__(self, device=0, **params):
        self.lib = load_library()
        cfg = default_config()
        for k, v in params.items():
            if not hasattr(cfg, k):
                raise KeyError(f"unknown BA parameter {k}")
            setattr(cfg, k, v)
        self.cfg = cfg
        self.h = C.c_void_p()
        rc = self.lib.cmlba_create(C.byref(cfg), device, C.byref(self.h))
        if rc != 0:
            raise CmlbaError(rc, self.lib.cmlba_last_error(None).decode())
        self.last_result = None

    def close(self):
        if getattr(self, "h", None):
            self.lib.cmlba_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _ck(self, rc):
        if rc != 0:
            raise CmlbaError(rc, self.lib.cmlba_last_error(self.h).decode())

    # ---- reference surface (DSOBundleAdjustment.h:26-101)
    def setCalibration(self, fx, fy, cx, cy, width, height):
        """mPinhole/mWidth/mHeight cached by addNewFrame (BA:419-425)."""
        self._ck(self.lib.cmlba_set_calib(self.h, fx, fy, cx, cy, int(width), int(height)))

    def addNewFrame(self, frame_id, world_to_cam, aff_a, aff_b, exposure_time, grad_image, is_init_frame=False):
        """addNewFrame(PFrame, immatureGroup) (BA:417-462). grad_image: [H,W,3] fp32 (I,dx,dy)."""
        w2c = np.ascontiguousarray(world_to_cam, dtype=np.float64).reshape(12)
        g = np.ascontiguousarray(grad_image, dtype=np.float32)
        self._ck(self.lib.cmlba_add_frame(self.h, int(frame_id), _ptr(w2c, C.c_double), float(aff_a), float(aff_b), float(exposure_time),
                                          _ptr(g, C.c_float), int(bool(is_init_frame))))

    def addNewFrameGray(self, frame_id, world_to_cam, aff_a, aff_b, exposure_time, gray_image, is_init_frame=False):
        """addNewFrame from the rectified level-0 gray image [H,W] fp32; the derivative image is built on the device (bit-identical)."""
        w2c = np.ascontiguousarray(world_to_cam, dtype=np.float64).reshape(12)
        g = np.ascontiguousarray(gray_image, dtype=np.float32)
        self._ck(self.lib.cmlba_add_frame_gray(self.h, int(frame_id), _ptr(w2c, C.c_double), float(aff_a), float(aff_b), float(exposure_time),
                                               _ptr(g, C.c_float), int(bool(is_init_frame))))

    def addNewFrameDevice(self, frame_id, w2c, aff_a, aff_b, exposure_time, d_texels, is_init_frame=False):
        """addNewFrame from device-resident level-0 texels (an int device pointer, e.g. CaptureImage.devicePtr("texel0"))."""
        w2c = np.ascontiguousarray(w2c, dtype=np.float64).reshape(12)
        self._ck(self.lib.cmlba_add_frame_device(self.h, int(frame_id), _ptr(w2c, C.c_double), float(aff_a), float(aff_b), float(exposure_time),
                                                 C.c_void_p(int(d_texels)), 1 if is_init_frame else 0))

    def addPoints(self, point_ids, host_frame_ids, xy, idepth):
        """addPoints(const PointSet&) (BA:382-415)."""
        pid = np.ascontiguousarray(point_ids, dtype=np.int64); hid = np.ascontiguousarray(host_frame_ids, dtype=np.int64)
        xyf = np.ascontiguousarray(xy, dtype=np.float32).reshape(-1, 2); idp = np.ascontiguousarray(idepth, dtype=np.float64)
        self._ck(self.lib.cmlba_add_points(self.h, int(pid.size), _ptr(pid, C.c_int64), _ptr(hid, C.c_int64), _ptr(xyf, C.c_float), _ptr(idp, C.c_double)))

    def removePoint(self, point_id):
        self._ck(self.lib.cmlba_remove_point(self.h, int(point_id)))

    def removeFrame(self, frame_id):
        self._ck(self.lib.cmlba_remove_frame(self.h, int(frame_id)))

    # ---- window maintenance (Hybrid::directMap, slam/modslam/direct/Mapping.cpp:61-100)
    def flagFramesForMarginalization(self, cams=None, num_immature=None):
        """flagFramesForMarginalization (BA:603-708); the reference calls it at the top of addNewFrame. Returns flagged frame ids."""
        n = self.lib.cmlba_num_frames(self.h)
        cp = None
        if cams is not None:
            cams = np.ascontiguousarray(cams, dtype=np.float64).reshape(-1, 12); cp = _ptr(cams, C.c_double)
        ip = None
        if num_immature is not None:
            num_immature = np.ascontiguousarray(num_immature, dtype=np.int32); ip = _ptr(num_immature, C.c_int)
        ids = np.zeros(max(n, 1), np.int64); cnt = C.c_int(n)
        self._ck(self.lib.cmlba_flag_frames_for_marginalization(self.h, cp, ip, _ptr(ids, C.c_int64), C.byref(cnt)))
        return ids[:cnt.value]

    def tryMarginalize(self):
        """tryMarginalize (BA:2240-2363). Returns (#points dropped -> getOutliers(), #points marked for marginalisation)."""
        a, b = C.c_int(0), C.c_int(0)
        self._ck(self.lib.cmlba_try_marginalize(self.h, C.byref(a), C.byref(b)))
        return a.value, b.value

    def marginalizePointsF(self):
        """marginalizePointsF (BA:2466-2513). Returns the ids of the points that left the window as marginalised."""
        n = self.lib.cmlba_num_points(self.h)
        ids = np.zeros(max(n, 1), np.int64); cnt = C.c_int(n)
        self._ck(self.lib.cmlba_marginalize_points(self.h, _ptr(ids, C.c_int64), C.byref(cnt)))
        return ids[:cnt.value]

    def marginalizeFrames(self):
        """marginalizeFrames (BA:710-742). Returns the ids of the removed frames."""
        n = self.lib.cmlba_num_frames(self.h)
        ids = np.zeros(max(n, 1), np.int64); cnt = C.c_int(n)
        self._ck(self.lib.cmlba_marginalize_frames(self.h, _ptr(ids, C.c_int64), C.byref(cnt)))
        return ids[:cnt.value]

    def frameCounters(self):
        """[N,4] int32: flaggedForMarginalization, numMarginalized, numResidualsOut, residuals targeting the frame."""
        return self.read("frame_counters", np.int32).reshape(-1, 4)

    def run(self, cams=None, updatePointsOnly=False, iterations=0):
        """bool run(bool updatePointsOnly) (BA:744-910). cams: [N,12] current Frame::getCamera() per window frame."""
        res = RunResult()
        cp = None
        if cams is not None:
            cams = np.ascontiguousarray(cams, dtype=np.float64).reshape(-1, 12)
            cp = _ptr(cams, C.c_double)
        rc = self.lib.cmlba_run(self.h, cp, int(iterations), int(bool(updatePointsOnly)), C.byref(res))
        self.last_result = res
        if rc == -4:  # CMLBA_ERR_NUMERIC: the reference returns false
            return False
        self._ck(rc)
        return True

    def numPoints(self):
        return int(self.lib.cmlba_num_points(self.h))

    def getStatistics(self):
        """{name: latest value} of the 19 Statistic series of DSOBundleAdjustment.h:215-233 (the reference's own names)."""
        v = np.zeros(19)
        self._ck(self.lib.cmlba_get_statistics(self.h, _ptr(v, C.c_double)))
        self.lib.cmlba_statistic_name.restype = C.c_char_p
        return {self.lib.cmlba_statistic_name(i).decode(): float(v[i]) for i in range(19)}

    def getOutliers(self):
        n = C.c_int(0)
        self._ck(self.lib.cmlba_get_outliers(self.h, None, C.byref(n)))
        out = np.zeros(max(n.value, 1), dtype=np.int64)
        m = C.c_int(n.value)
        self._ck(self.lib.cmlba_get_outliers(self.h, _ptr(out, C.c_int64), C.byref(m)))
        return out[:n.value]

    def getFrames(self):
        n = self.lib.cmlba_num_frames(self.h)
        ids = np.zeros(n, np.int64); w2c = np.zeros((n, 12)); ab = np.zeros((n, 2)); st = np.zeros((n, 10)); ev = np.zeros((n, 12)); th = np.zeros(n)
        self._ck(self.lib.cmlba_get_frames(self.h, _ptr(ids, C.c_int64), _ptr(w2c, C.c_double), _ptr(ab, C.c_double), _ptr(st, C.c_double), _ptr(ev, C.c_double), _ptr(th, C.c_double)))
        return dict(id=ids, world_to_cam=w2c, affine=ab, state=st, evalpt=ev, energy_th=th)

    def getPoints(self):
        n = self.lib.cmlba_num_points(self.h)
        ids = np.zeros(n, np.int64); idp = np.zeros(n); unc = np.zeros(n); idh = np.zeros(n, np.float32); mrb = np.zeros(n, np.float32)
        ng = np.zeros(n, np.int32); gft = np.zeros(n, np.int32)
        self._ck(self.lib.cmlba_get_points(self.h, _ptr(ids, C.c_int64), _ptr(idp, C.c_double), _ptr(unc, C.c_double), _ptr(idh, C.c_float), _ptr(mrb, C.c_float),
                                           _ptr(ng, C.c_int), _ptr(gft, C.c_int)))
        return dict(id=ids, idepth=idp, uncertainty=unc, idepth_hessian=idh, max_rel_baseline=mrb, num_good_residuals=ng, good_for_tracking=gft)

    def getGoodPointsForTracking(self):
        p = self.getPoints()
        return p["id"][p["good_for_tracking"] != 0]

    def getResiduals(self):
        n = self.lib.cmlba_num_residuals(self.h)
        pid = np.zeros(n, np.int64); tid = np.zeros(n, np.int64); st = np.zeros(n, np.int32); en = np.zeros(n)
        self._ck(self.lib.cmlba_get_residuals(self.h, _ptr(pid, C.c_int64), _ptr(tid, C.c_int64), _ptr(st, C.c_int), _ptr(en, C.c_double)))
        return dict(point_id=pid, target_frame_id=tid, state=st, energy=en)

    # ---- stage entry points (protected members of the reference class)
    def prepare(self, cams=None):
        cp = None
        if cams is not None:
            cams = np.ascontiguousarray(cams, dtype=np.float64).reshape(-1, 12)
            cp = _ptr(cams, C.c_double)
        self._ck(self.lib.cmlba_prepare(self.h, cp))

    def linearizeAll(self, fixLinearization=False):
        e = C.c_double(0)
        self._ck(self.lib.cmlba_linearize(self.h, int(bool(fixLinearization)), C.byref(e)))
        return e.value

    def applyActiveRes(self):
        self._ck(self.lib.cmlba_apply(self.h))

    def solveSystem(self, iteration):
        self._ck(self.lib.cmlba_solve(self.h, int(iteration)))

    def doStepFromBackup(self, updatePointsOnly=False):
        cb = C.c_int(0)
        self._ck(self.lib.cmlba_step(self.h, int(bool(updatePointsOnly)), C.byref(cb)))
        return bool(cb.value)

    def read(self, name, dtype, shape=None):
        nb = C.c_size_t(0)
        self._ck(self.lib.cmlba_read(self.h, name.encode(), None, 0, C.byref(nb)))
        dt = np.dtype(dtype)
        out = np.zeros(nb.value // dt.itemsize, dtype=dt)
        if nb.value:
            self._ck(self.lib.cmlba_read(self.h, name.encode(), out.ctypes.data_as(C.c_void_p), nb.value, C.byref(nb)))
        return out.reshape(shape) if shape is not None else out

    def initCommunicator(self, rank, world, peer_memory=True):
        """One process per GPU: NCCL communicator (unique id broadcast through torch.distributed) and, optionally, the peer-memory
        exchange of the reduced system (cudaIpc handles all-gathered through torch.distributed)."""
        import torch
        import torch.distributed as dist
        uid = torch.zeros(128, dtype=torch.uint8)
        if rank == 0:
            buf = np.zeros(128, dtype=np.uint8)
            self._ck(self.lib.cmlba_nccl_unique_id(buf.ctypes.data))
            uid = torch.from_numpy(buf)
        uid = uid.cuda()
        dist.broadcast(uid, 0)
        ub = uid.cpu().numpy()
        self._ck(self.lib.cmlba_comm_init(self.h, ub.ctypes.data, rank, world))
        if peer_memory and world > 1:
            # every rank must end up in the same mode: a failed cudaIpc mapping anywhere sends all ranks back to NCCL
            ok = 1
            mine = np.zeros(64, dtype=np.uint8)
            if self.lib.cmlba_comm_ipc_handle(self.h, mine.ctypes.data) != 0:
                ok = 0
            allh = [torch.zeros(64, dtype=torch.uint8, device="cuda") for _ in range(world)]
            dist.all_gather(allh, torch.from_numpy(mine).cuda())
            packed = np.ascontiguousarray(torch.stack(allh).cpu().numpy())
            if ok and self.lib.cmlba_comm_ipc_open(self.h, packed.ctypes.data) != 0:
                ok = 0
            flag = torch.tensor([ok], device="cuda")
            dist.all_reduce(flag, op=dist.ReduceOp.MIN)
            self.peer_memory = bool(flag.item())
            if not self.peer_memory:
                self._ck(self.lib.cmlba_comm_ipc_open(self.h, None))
        else:
            self.peer_memory = False

    def reset(self):
        self._ck(self.lib.cmlba_reset(self.h))

    def benchPass(self, steps, warmup=3, flush_l2=True):
        r = BenchResult()
        self._ck(self.lib.cmlba_bench_pass(self.h, int(steps), int(warmup), int(bool(flush_l2)), C.byref(r)))
        return r

    def enableDebugDump(self):
        self._ck(self.lib.cmlba_read(self.h, b"enable_dbg", None, 0, None))

    # ---- convenience: build a window from a synthetic / snapshot dict (libcml_b200.synth / cmlw)
    def loadWindow(self, win):
        W, H = int(win["size"][0]), int(win["size"][1])
        fx, fy, cx, cy = [float(v) for v in win["calib"]]
        self.setCalibration(fx, fy, cx, cy, W, H)
        N = win["frame_evalpt"].shape[0]
        grad = win.get("grad")
        if grad is None:
            from .synth import gradient_image
            grad = [gradient_image(win["gray"][i]) for i in range(N)]
        init = win.get("frame_init", np.zeros(N, np.uint8))
        for i in range(N):
            self.addNewFrame(i, win["frame_evalpt"][i], win["frame_affine"][i, 0], win["frame_affine"][i, 1], win["frame_exposure"][i], grad[i], bool(init[i]))
        P = win["pt_host"].size
        self.addPoints(np.arange(P), win["pt_host"], win["pt_xy"], win["pt_idepth"])
        return win["frame_cam"]
