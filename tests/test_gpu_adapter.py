"""The COMPILED adapter class (adapter/DSOBundleAdjustmentB200.h: the reference class's public surface on top of libcmlba's C ABI) against the
unmodified reference class on the same synthetic Map graph: oracle/_ref/cmlba_adapter_check (oracle/adapter_check.cpp, built by
`make -C oracle adapter` against the reference headers, -lcmlba) runs run() -> tryMarginalize -> removePoint(outliers) -> marginalizePointsF ->
marginalizeFrames -> run() through both classes and compares what they leave in the graph (frame cameras, exposure parameters, inverse depths,
surviving / good-for-tracking / outlier point sets, marginalised frames, the 19 Statistic series under the reference's names)."""
import json
import os
import subprocess

import pytest

from parity_util import GOLDEN, ROOT, record

pytestmark = pytest.mark.gpu
BIN = os.path.join(ROOT, "oracle", "_ref", "cmlba_adapter_check")


def _runs():
    if not os.path.exists(BIN):
        return False
    try:
        return subprocess.run([BIN], capture_output=True, timeout=30).returncode == 2      # prints usage
    except Exception:
        return False


@pytest.mark.parametrize("window,maintain", [("tiny_window", 0), ("tiny_affine_window", 0), ("maint_window", 1)])
def test_adapter_class_against_reference_class(window, maintain):
    if not _runs():
        pytest.skip("oracle/_ref/cmlba_adapter_check not built / not runnable on this host (make -C oracle adapter)")
    r = subprocess.run([BIN, "--window", os.path.join(GOLDEN, window + ".cmlw"), "--maintain", str(maintain)], capture_output=True, text=True, timeout=600)
    line = [l for l in r.stdout.splitlines() if l.startswith("{")]
    assert line, r.stdout[-2000:] + r.stderr[-2000:]
    d = json.loads(line[-1])
    flat = {}
    for k, v in d.items():
        if isinstance(v, dict):
            flat.update({f"{k}.{kk}": vv for kk, vv in v.items()})
        elif not isinstance(v, (bool, str)):
            flat[k] = v
    record(f"adapter[{window}]", **flat)
    assert r.returncode == 0 and d["ok"], d
