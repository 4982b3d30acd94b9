"""The B200 backend INSIDE the reference application: SourceManager.set_source constructs, starts and re-tunes the
replacement classes (core/source_manager.py:376-494,554-572).  Needs the reference checkout, so it runs in the build
container only (the GPU box has no /root/reference); hardware libraries are faked."""
import json
import os
import subprocess
import sys

import pytest
import torch

REF = os.environ.get("TDSA_REFERENCE_ROOT", "/root/reference")
REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

pytestmark = pytest.mark.skipif(not os.path.isdir(os.path.join(REF, "core")), reason="reference checkout not present")


def _drive(*flags):
    env = dict(os.environ)
    env.pop("TDSA_BACKEND", None)
    r = subprocess.run([sys.executable, os.path.join(REPO, "tests", "helpers", "ref_app_driver.py"), REF, REPO, *flags],
                       capture_output=True, text=True, timeout=300, env=env)
    lines = [l for l in r.stdout.splitlines() if l.startswith("RESULT ")]
    assert r.returncode == 0 and lines, r.stdout[-2000:] + r.stderr[-4000:]
    return json.loads(lines[-1][7:])


def _check_common(o):
    assert o["in_reference_app"] is True                 # the class really derives from the reference's ABC
    assert o["rtl_class"] == "B200RtlSamples", o["rtl_status"]
    assert o["rtl_running"] and o["rtl_isinstance_ref"]
    assert o["post_start_ran"]                           # _post_start_sample_source pushed 1024 bins to every widget
    # start(): opened the dongle itself and programmed rate, centre (rtl_samples.py:42-46); rate read back
    assert o["rtl_log"][0] == ["open"] and o["rtl_log"][1][0] == "rate" and o["rtl_log"][2][0] == "centre"
    assert abs(o["rtl_rate_readback"] - o["rtl_log"][1][1] * 1.000001) < 1e-3
    # span change: new rate, centre written again after it (rtl_samples.py:125-129), then the new centre + flush
    assert o["retune_log"] == [["rate", 1000000], ["centre", 98000000], ["centre", 99000000]]
    assert o["retune_flush"] == 5                        # max(3, int(0.006 * 1000001 / 1024)) discarded reads (rtl_samples.py:99-101)
    assert o["hackrf_class"] == "B200HackrfSamples", o["hackrf_status"]
    assert o["rtl_paused_kept"]                          # isinstance(RtlSamplesDataSource) kept the pause path
    assert o["hackrf_gains"] == [24, 30]                 # set by _initialise_hackrf_samples before start()
    assert o["hackrf_log"][0] == ["hackrf_open"] and o["hackrf_log"][1][0] == "set_sample_rate"
    assert ["set_lna_gain", 24] in o["hackrf_log"] and ["set_vga_gain", 30] in o["hackrf_log"]
    assert o["hackrf_thread_alive"] and o["hackrf_raw_len"] == 1024
    assert o["rtl_resumed_same_object"] and o["hackrf_closed"] and o["rtl_closed_for_sweep"]
    assert o["uninstalled"] == "RtlSamplesDataSource"
    # the IQ-in sweep source behind "hackrf_sweep": constructed by _initialise_hackrf_sweep(start, stop, bin_size=30000)
    assert o["sweep_class"] == "B200SweepDataSource", o["sweep_status"]
    assert o["sweep_running"] and o["sweep_isinstance_ref"] and o["sweep_gains"] == [24, 30]
    assert o["sweep_grid_len"] > 0 and o["sweep_stopped"]


@pytest.mark.skipif(torch.cuda.is_available(), reason="CPU variant (plan stubbed); the gpu variant runs the real thing")
def test_set_source_constructs_b200_backend_cpu():
    _check_common(_drive("--no-gpu"))


@pytest.mark.gpu
def test_set_source_constructs_b200_backend_gpu():
    o = _drive()
    _check_common(o)
    assert o["frame_shape"] == [1024] and o["frame_finite"] and o["hackrf_frame_dtype"] == "float32"
    assert o["sweep_data_finite"]
