"""Drives the reference's SourceManager with the B200 backend installed (run in a subprocess by
tests/test_reference_app.py with the reference's root first on sys.path).

Hardware libraries are replaced by fakes (pattern: /root/reference/test_smoke.py:26-37); the main window is a stub
(pattern: /root/reference/test_preset_manager.py:33-70).  Prints one JSON line with what happened.
"""
import json
import os
import sys
import tempfile
import types
from unittest.mock import MagicMock

import numpy as np

REF = sys.argv[1]
REPO = sys.argv[2]
sys.path.insert(0, REF)
sys.path.insert(1, REPO)

LOG = []


class FakeRtlSdr:
    """pyrtlsdr's surface; records what the source programs."""
    def __init__(self):
        self._fs, self._fc, self.gain, self.closed = 0.0, 0.0, None, False
        self.rng = np.random.default_rng(0)
        LOG.append(("open",))

    sample_rate = property(lambda s: s._fs, lambda s, v: (LOG.append(("rate", int(v))), setattr(s, "_fs", float(v) * 1.000001)))
    center_freq = property(lambda s: s._fc, lambda s, v: (LOG.append(("centre", int(v))), setattr(s, "_fc", float(v))))

    def get_sample_rate(self):
        return self._fs

    def get_center_freq(self):
        return self._fc

    def read_samples(self, n):
        return (self.rng.standard_normal(n) + 1j * self.rng.standard_normal(n)) / np.sqrt(2)

    def close(self):
        self.closed = True
        LOG.append(("close",))


class FakeHackRF:
    def __init__(self):
        LOG.append(("hackrf_open",))
        self.rng = np.random.default_rng(1)

    def __getattr__(self, name):
        if name.startswith(("set_", "enable_", "disable_")):
            return lambda *a: LOG.append((name,) + tuple(int(x) for x in a))
        raise AttributeError(name)

    def read_samples(self, n):
        import time
        time.sleep(0.002)
        return (self.rng.standard_normal(n) + 1j * self.rng.standard_normal(n)).astype(np.complex64)

    def close(self):
        LOG.append(("hackrf_close",))


rtl_mod = types.ModuleType("rtlsdr"); rtl_mod.RtlSdr = FakeRtlSdr
hk_mod = types.ModuleType("hackrf"); hk_mod.HackRF = FakeHackRF
sys.modules["rtlsdr"], sys.modules["hackrf"], sys.modules["sounddevice"] = rtl_mod, hk_mod, MagicMock()

from utils.frequency_selector import FrequencyRange                      # noqa: E402  (reference)
import core.source_manager as sm_mod                                      # noqa: E402  (reference)

tmp = tempfile.mkdtemp()
sm_mod.config_dir = lambda: __import__("pathlib").Path(tmp)

import topdogspectrumanalyser_b200.datasources as ds                      # noqa: E402
from topdogspectrumanalyser_b200.datasources import b200_samples          # noqa: E402


class _Label:
    def __init__(self): self.text = ""
    def setText(self, t): self.text = t
    def setEnabled(self, v): pass


class _Widget:
    def __init__(self): self.bins = None
    def update_frequency_bins(self, b): self.bins = np.asarray(b)


class _FreqMgr:
    def __init__(self, mw): self.mw = mw
    def set_frequency_range(self, a, b): self.mw.frequency.set_start_stop(a, b)
    def update_frequency_values(self): pass
    def _update_display_bins(self): pass


class _DispMgr:
    resets = 0
    def _reset_dsp_state(self): _DispMgr.resets += 1
    def set_display(self, *a): pass


class FakeMW:
    def __init__(self):
        self.frequency = FrequencyRange(88e6, 108e6)
        self.current_source = None
        self.current_source_id = None
        self.current_stacked_index = 0
        self.hackrf_lna_gain, self.hackrf_vga_gain = 24, 30
        self.status_label, self.output_source = _Label(), _Label()
        self.button_peak_search = self.button_max_hold = self.button_hold = _Label()
        self.two_d_widget, self.three_d_widget, self.waterfall_widget, self.surface_widget = (_Widget() for _ in range(4))
        self.display_manager = _DispMgr()
        self.frequency_manager = _FreqMgr(self)
        self.calibration_manager = None


out = {"in_reference_app": ds.IN_REFERENCE_APP}
no_gpu = "--no-gpu" in sys.argv
if no_gpu:
    # this container has no CUDA device: the plan (the only thing start() needs the GPU for) is stubbed so the
    # registration / construction / start / post-start path itself can be checked on CPU
    b200_samples.B200SampleDataSource._ensure_plan = lambda self: None

assert b200_samples.install_backend() is None, "backend must stay off without TDSA_BACKEND=b200"
os.environ["TDSA_BACKEND"] = "b200"
SM = b200_samples.install_backend()
mw = FakeMW()
mgr = SM(mw)

mgr.set_source("rtl_samples")
src = mw.current_source
out["rtl_class"] = type(src).__name__
out["rtl_status"] = mw.status_label.text
out["rtl_running"] = bool(src is not None and src.running)
out["rtl_isinstance_ref"] = isinstance(src, sm_mod.RtlSamplesDataSource) and isinstance(src, sm_mod.SampleDataSource)
out["post_start_ran"] = getattr(mw, "last_span", None) == mw.frequency.span and mw.two_d_widget.bins is not None \
    and len(mw.two_d_widget.bins) == 1024
out["rtl_log"] = list(LOG)
out["rtl_rate_readback"] = src.sample_rate if src is not None else None
# a span change goes through _perform_full_frequency_update -> update_frequency: rate, then the centre re-tune
del LOG[:]
mw.frequency.set_start_stop(99e6 - 0.5e6, 99e6 + 0.5e6)
mgr.update_source_frequency()
out["retune_log"] = list(LOG)
out["retune_flush"] = src._flush_reads_remaining
if not no_gpu:
    p, bins = src.get_power_levels()
    out["frame_shape"] = list(p.shape)
    out["frame_finite"] = bool(np.isfinite(p).all() and p.any())
# switching to the HackRF source pauses the RTL one (smart RTL handling needs the isinstance to hold)
del LOG[:]
mgr.set_source("hackrf_samples")
hs = mw.current_source
out["hackrf_class"] = type(hs).__name__
out["hackrf_status"] = mw.status_label.text
out["rtl_paused_kept"] = mgr.paused_rtl_source is src and not src.running and not src.sdr.closed
out["hackrf_gains"] = [hs.lna_gain, hs.vga_gain] if hs is not None else None
out["hackrf_log"] = [e for e in LOG if e[0] != "rate"][:8]
out["hackrf_thread_alive"] = bool(hs is not None and hs.thread is not None and hs.thread.is_alive())
if hs is not None:
    x = hs.read_samples_only()
    out["hackrf_raw_len"] = None if x is None else len(x)
    if not no_gpu:
        p, bins = hs.get_power_levels()
        out["hackrf_frame_dtype"] = str(p.dtype)
mgr.set_source("rtl_samples")                      # resumes the paused object, no second open
out["rtl_resumed_same_object"] = mw.current_source is src and src.running
out["hackrf_closed"] = ("hackrf_close",) in LOG
mgr._stop_current_source("rtl_sweep")
out["rtl_closed_for_sweep"] = ("close",) in LOG
# optional third replacement: the HackRF sweep source computed from raw IQ (no external hackrf_sweep binary)
from topdogspectrumanalyser_b200.datasources import b200_sweep
if no_gpu:
    b200_sweep.B200SweepDataSource._ensure_sweep = lambda self: None
    b200_sweep.B200SweepDataSource.process_sweep = lambda self, iq: None
b200_samples.install_backend(sweep=True)
del LOG[:]
mgr.set_source("hackrf_sweep")
sw = mw.current_source
out["sweep_class"] = type(sw).__name__
out["sweep_status"] = mw.status_label.text
out["sweep_running"] = bool(sw is not None and sw.is_running)
out["sweep_gains"] = [sw.lna_gain, sw.vga_gain] if sw is not None else None
out["sweep_grid_len"] = len(sw.frequency_grid) if sw is not None else None
out["sweep_isinstance_ref"] = isinstance(sw, sm_mod.SweepDataSource)
import time as _t
_t.sleep(0.3)
if not no_gpu and sw is not None:
    t0 = _t.time()
    while sw.sweep_rate is None and _t.time() - t0 < 20:
        _t.sleep(0.05)
    d = sw.get_data()
    out["sweep_data_finite"] = bool(len(d) == len(sw.frequency_grid) and np.isfinite(d).all())
mgr._stop_current_source("rtl_samples")
out["sweep_stopped"] = bool(sw is not None and not sw.is_running) and ("hackrf_close",) in LOG
b200_samples.uninstall_backend()
out["uninstalled"] = SM.SOURCE_CLASSES["rtl_samples"].__name__
print("RESULT " + json.dumps(out))
