"""CPU tests: the C-ABI library loads, exports every symbol include/cmlba.h declares, and fails loudly without a GPU."""
import ctypes as C
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    src = open(os.path.join(ROOT, "include", "cmlba.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(cmlba_[a-z_0-9]+)\s*\(", src)))


def test_header_symbols_exported(lib):
    syms = declared_symbols()
    assert len(syms) >= 25
    for s in syms:
        assert hasattr(lib, s), f"libcmlba.so does not export {s}"


def test_binding_lists_every_symbol():
    from libcml_b200 import binding
    assert sorted(binding.SYMBOLS) == declared_symbols()


def test_tracker_header_symbols_exported(lib):
    """include/cmltrk.h (coarse tracker boundary): every declared entry point is exported and listed by the binding."""
    from libcml_b200 import tracker
    src = open(os.path.join(ROOT, "include", "cmltrk.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    syms = sorted(set(re.findall(r"\b(cmltrk_[a-z_0-9]+)\s*\(", src)))
    assert sorted(tracker.TRACKER_SYMBOLS) == syms and len(syms) == 16
    for s in syms:
        assert hasattr(lib, s), f"libcmlba.so does not export {s}"
    cfg = tracker.TrackerConfig()
    tracker._bind(lib).cmltrk_default_config(C.byref(cfg))
    # reference defaults, DSOTracker.h:479-518
    assert cfg.huber_threshold == 9.0 and cfg.cutoff_threshold == 20.0 and cfg.scale_translation == 0.5 and cfg.scale_light_a == 10.0
    assert cfg.scale_light_b == 1000.0 and cfg.optimize_a == 1 and cfg.optimize_b == 1 and cfg.saturated_ratio_threshold == 0.33


def test_tracer_header_symbols_exported(lib):
    """include/cmltrc.h (immature-point tracer boundary)."""
    from libcml_b200 import tracer
    src = open(os.path.join(ROOT, "include", "cmltrc.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    syms = sorted(set(re.findall(r"\b(cmltrc_[a-z_0-9]+)\s*\(", src)))
    assert sorted(tracer.TRACER_SYMBOLS) == syms and len(syms) == 16
    for s in syms:
        assert hasattr(lib, s), f"libcmlba.so does not export {s}"
    cfg = tracer.TracerConfig()
    tracer._bind(lib).cmltrc_default_config(C.byref(cfg))
    # reference defaults, DSOTracer.h:186-203
    assert cfg.min_idepth_h_act == 100.0 and cfg.gn_iterations == 3 and cfg.huber_threshold == 9.0 and cfg.outlier_th == 144.0
    assert cfg.outlier_th_sum_component == 2500.0 and abs(cfg.max_pix_search - 0.027) < 1e-7 and cfg.max_slack_interval == 1.5
    assert ctypes_sizeof_matches(tracer)
    import torch
    if not torch.cuda.is_available():
        h = C.c_void_p()
        assert lib.cmltrc_create(None, 0, 640, 480, 500.0, 500.0, 320.0, 240.0, C.byref(h)) == -2 and b"no CUDA device" in lib.cmltrc_last_error(None)


def test_imgprep_header_symbols_exported(lib):
    """include/cmlimg.h (image preparation boundary)."""
    from libcml_b200 import imgprep
    src = open(os.path.join(ROOT, "include", "cmlimg.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    syms = sorted(set(re.findall(r"\b(cmlimg_[a-z_0-9]+)\s*\(", src)))
    assert sorted(imgprep.IMG_SYMBOLS) == syms and len(syms) == 12
    for s in syms:
        assert hasattr(lib, s), f"libcmlba.so does not export {s}"
    import torch
    if not torch.cuda.is_available():
        h = C.c_void_p()
        assert imgprep._bind(lib).cmlimg_create(0, 640, 480, 640, 480, 0, C.byref(h)) == -2 and b"no CUDA device" in lib.cmlimg_last_error(None)


def test_selector_header_symbols_exported(lib):
    """include/cmlsel.h (pixel selector boundary)."""
    from libcml_b200 import selector
    src = open(os.path.join(ROOT, "include", "cmlsel.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    syms = sorted(set(re.findall(r"\b(cmlsel_[a-z_0-9]+)\s*\(", src)))
    assert sorted(selector.SEL_SYMBOLS) == syms and len(syms) == 7
    for s in syms:
        assert hasattr(lib, s), f"libcmlba.so does not export {s}"
    import torch
    if not torch.cuda.is_available():
        h = C.c_void_p()
        assert selector._bind(lib).cmlsel_create(0, 640, 480, C.byref(h)) == -2 and b"no CUDA device" in lib.cmlsel_last_error(None)


def test_fast_header_symbols_exported(lib):
    """include/cmlfast.h (FAST-9 boundary)."""
    from libcml_b200 import fast
    src = open(os.path.join(ROOT, "include", "cmlfast.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    syms = sorted(set(re.findall(r"\b(cmlfast_[a-z_0-9]+)\s*\(", src)))
    assert sorted(fast.FAST_SYMBOLS) == syms and len(syms) == 4
    for s in syms:
        assert hasattr(lib, s), f"libcmlba.so does not export {s}"
    import torch
    if not torch.cuda.is_available():
        h = C.c_void_p()
        assert fast._bind(lib).cmlfast_create(0, 640, 480, C.byref(h)) == -2 and b"no CUDA device" in lib.cmlfast_last_error(None)


def ctypes_sizeof_matches(tracer):
    # cmltrc_point: 2 x int32 + 11 doubles; cmltrc_activation: int32 + float + uint32
    return tracer.POINT.itemsize == 96 and tracer.ACTIVATION.itemsize == 12


def test_tracker_has_no_cpu_fallback(lib):
    import torch
    from libcml_b200 import tracker
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    h = C.c_void_p()
    rc = tracker._bind(lib).cmltrk_create(None, 0, 640, 480, 500.0, 500.0, 320.0, 240.0, C.byref(h))
    assert rc == -2 and not h.value and b"no CUDA device" in lib.cmltrk_last_error(None)


def test_version_and_default_config(lib):
    from libcml_b200 import binding
    assert lib.cmlba_version().decode().endswith("sm_100a")
    cfg = binding.default_config()
    # reference defaults, DSOBundleAdjustment.h:235-288
    assert cfg.iterations == 4 and cfg.huber_threshold == 9.0 and cfg.outlier_th_sum == 2500.0
    assert cfg.scale_translation == 0.5 and cfg.scale_light_a == 10.0 and cfg.scale_light_b == 1000.0
    assert cfg.force_accept == 1 and cfg.fix_lambda == 1 and abs(cfg.fixed_lambda - 1e-5) < 1e-12
    assert cfg.idepth_fix_prior == 2500 and cfg.disable_marginalization == 1 and cfg.max_frames == 6


def test_no_cpu_fallback(lib):
    """Without a CUDA device creation must fail loudly (CMLBA_ERR_CUDA), never fall back to the CPU."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    h = C.c_void_p()
    rc = lib.cmlba_create(None, 0, C.byref(h))
    assert rc == -2 and not h.value
    assert b"no CUDA device" in lib.cmlba_last_error(None)


def test_null_handle_is_an_error(lib):
    assert lib.cmlba_set_calib(None, 1.0, 1.0, 1.0, 1.0, 64, 64) == -1
    assert lib.cmlba_run(None, None, 1, 0, None) == -1


def test_tools_do_not_touch_oracle():
    """tools/ are measurement aids of the product: the reference CPU timings they report come through bench.py's cpu_baseline callback
    (`python bench.py --component ...`), the one place outside tests/ and smoke() that may execute oracle/."""
    for f in os.listdir(os.path.join(ROOT, "tools")):
        if f.endswith(".py"):
            txt = open(os.path.join(ROOT, "tools", f)).read()
            assert "oracle" not in txt, f


def test_product_does_not_import_oracle():
    """The product path must never route through oracle/ (only tests, smoke() and bench.py's CPU legs may)."""
    pkg = os.path.join(ROOT, "libcml_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".cpp")):
                txt = open(os.path.join(dirpath, f)).read()
                assert "ba_oracle" not in txt and "oracle/" not in txt.replace("oracle/cmlw_io.h", "").replace("oracle/ref_driver.cpp", ""), f


def test_statistic_names_are_the_reference_ones(lib):
    """cmlba_statistic_name(i) = the 19 createStatistic("...") names of DSOBundleAdjustment.h:215-233, in declaration order (the committed list below is
    what the reference header holds; when /root/reference is mounted the header itself is parsed as well)."""
    import ctypes as C
    expected = ["P Energy ( All residuals )", "R Energy", "L Energy ( Linearized )", "M Energy ( Marginalized )", "Total Energy", "X Norm", " Hessian P Norm",
                " Hessian L Norm", " Hessian M Norm", " Hessian SC Norm", "P B Norm", "L B Norm", "M B Norm", "SC B Norm", "OOB", "In", "InIn", "Nores", "Num Linearized"]
    lib.cmlba_statistic_name.restype = C.c_char_p
    lib.cmlba_statistic_name.argtypes = [C.c_int]
    names = [lib.cmlba_statistic_name(i).decode() for i in range(19)]
    assert names == expected and lib.cmlba_statistic_name(19) is None
    hdr = "/root/reference/src/cml/optimization/dso/DSOBundleAdjustment.h"
    if os.path.exists(hdr):
        assert re.findall(r'createStatistic\("([^"]*)"\)', open(hdr).read()) == expected


def test_bench_arms_share_one_config_object():
    """bench.py: the `config` of our arm and of `--impl reference` comes from the same function (the driver compares them), per workload and scaling."""
    import importlib.util
    spec = importlib.util.spec_from_file_location("bench_mod", os.path.join(ROOT, "bench.py"))
    b = importlib.util.module_from_spec(spec); spec.loader.exec_module(b)
    for wl in ("c2", "c4"):
        for sc in ("weak", "strong"):
            c = b.config_dict(wl, sc)
            assert c == b.config_dict(wl, sc) and "workload" in c and not any(k in c for k in ("model", "seq_len", "global_batch"))
    src = open(os.path.join(ROOT, "bench.py")).read()
    assert src.count('"config": config_dict(args.workload, args.scaling)') == 2        # once per arm
