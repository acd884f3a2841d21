"""Measures the image preparation (SURVEY.md 8f NEXT #3): device time of the two kernels with the L2 flushed between frames, achieved HBM bandwidth
against the measured peak (MEASURED_PEAKS.json), end-to-end cmlimg_prepare (host image in).  Run as `python bench.py --component prepare` it also
receives bench.py's cpu_baseline callback: the unmodified reference's CaptureImageGenerator::generate on the host CPU for the same inputs (its
own undistortion map is then used and the outputs are compared).  Prints one JSON line."""
import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from libcml_b200 import CaptureImageGenerator, synth  # noqa: E402


def main(argv=None, reference=None):
    ap = argparse.ArgumentParser()
    ap.add_argument("--width", type=int, default=1920); ap.add_argument("--height", type=int, default=1080); ap.add_argument("--repeats", type=int, default=30)
    a = ap.parse_args(argv)
    Wi, Hi = a.width + a.width // 8, a.height + a.height // 8            # sensor image a little larger than the rectified one
    w = synth.prepare_scenario(Wi, Hi, a.width, a.height)
    out = {"workload": f"{Wi}x{Hi} sensor -> {a.width}x{a.height} rectified, LUT + vignette + radtan undistortion"}
    g = reference("prepare", w, 5) if reference else None
    umap = g["prep_map"] if g is not None else synth.radtan_undistort_map(w)
    gen = CaptureImageGenerator(Wi, Hi, a.width, a.height)
    gen.setLut(w["lut"]); gen.setInverseVignette(w["inv_vignette"]); gen.setUndistortMap(umap)
    cap = gen.generate(w["raw"])
    L = cap.getPyramidLevels()
    out["levels"] = L
    if g is not None:
        out["max_abs_diff_vs_reference"] = max(float(np.abs(cap.getGrayImage(l) - g[f"prep_gray{l}"]).max()) for l in range(L))
    for _ in range(3):
        gen.generate(w["raw"])
    t0 = time.perf_counter()
    for _ in range(a.repeats):
        gen.generate(w["raw"])
    e2e = (time.perf_counter() - t0) / a.repeats * 1e3
    buf = gen.inputBuffer(); buf[...] = w["raw"]
    t0 = time.perf_counter()
    for _ in range(a.repeats):
        gen.generate(buf)
    e2e_pinned = (time.perf_counter() - t0) / a.repeats * 1e3
    raw8 = np.clip(np.rint(w["raw"]), 0, 255).astype(np.uint8)
    gen.generate(raw8)
    t0 = time.perf_counter()
    for _ in range(a.repeats):
        gen.generate(raw8)
    e2e_u8 = (time.perf_counter() - t0) / a.repeats * 1e3
    gen.generate(buf)
    cold = gen.bench(a.repeats, True); warm = gen.bench(a.repeats, False)
    px_out = sum(wl * hl for wl, hl in gen.sizes)
    px0 = a.width * a.height
    fin = float(np.isfinite(umap[..., 0]).mean())
    # algorithmic bytes: map (8 B per rectified pixel) + raw and vignette of the sampled footprint (4 + 4 B per finite pixel: neighbouring
    # rectified pixels share their taps) + gray levels written once (4 B) + read once by the texel kernel (L2 at this size, not counted) + texels (16 B)
    alg = px0 * 8 + fin * px0 * 8 + px_out * 4 + px_out * 16
    peak = None
    try:
        peak = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))).get("hbm_gbs")
    except Exception:
        pass
    out.update({"kernels_device_ms_l2_flushed": round(cold, 4), "kernels_device_ms_l2_warm": round(warm, 4), "algorithmic_bytes": int(alg),
                "achieved_gbs": round(alg / (cold * 1e-3) / 1e9, 1), "hbm_peak_gbs": peak, "roofline_frac": round(alg / (cold * 1e-3) / 1e9 / peak, 3) if peak else None,
                "prepare_e2e_ms": round(e2e, 4), "prepare_e2e_from_input_buffer_ms": round(e2e_pinned, 4), "prepare_e2e_u8_image_ms": round(e2e_u8, 4), "h2d_bytes_per_frame": int(w["raw"].nbytes),
                })
    if g is not None:
        out.update({"reference_cpu_generate_ms": round(float(g["prep_seconds"][0]) * 1e3, 3), "speedup_e2e_vs_reference_cpu": round(float(g["prep_seconds"][0]) * 1e3 / e2e_pinned, 1)})
    print(json.dumps(out))


if __name__ == "__main__":
    main()
