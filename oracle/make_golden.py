#!/usr/bin/env python
"""Generate tests/golden/*.npz by EXECUTING the reference (test infrastructure).

Run in the build container only (``/root/reference`` does not exist on the GPU
box):  ``python oracle/make_golden.py``.  Hardware modules (``rtlsdr``,
``hackrf``, ``sounddevice``) are mocked; numpy and scipy are the real ones, so
the arrays stored here are what the reference's own classes return for the
given seeded IQ.  Nothing from the reference is copied into the repo: its
modules are imported (or, for the Qt-bound waterfall widget, three helper
methods are compiled from its source file at run time) and only their outputs
are written.

Versions used are recorded in each fixture (``meta``).
"""
from __future__ import annotations

import ast
import json
import os
import sys
import types
from unittest.mock import MagicMock

import numpy as np
import scipy

REF = os.environ.get("TDSA_REFERENCE_ROOT", "/root/reference")
HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
OUT = os.path.join(ROOT, "tests", "golden")

for _m in ("rtlsdr", "hackrf", "sounddevice"):
    sys.modules[_m] = MagicMock()
sys.path.insert(0, REF)
sys.path.insert(0, ROOT)

from datasources.rtl_samples import RtlSamplesDataSource          # noqa: E402
from datasources.hackrf_samples import HackrfSamplesDataSource    # noqa: E402
from datasources.hackrf_sweep import HackRFSweepDataSource        # noqa: E402
from utils.signal_processing import TraceAverager                 # noqa: E402
from core.display_data_processor import DataProcessor             # noqa: E402
from core.tare_state import TareState                             # noqa: E402

from topdogspectrumanalyser_b200 import synth                     # noqa: E402

META = json.dumps({"numpy": np.__version__, "scipy": scipy.__version__,
                   "reference_commit": "1e46762", "generator": "oracle/make_golden.py"})


class FakeSdr:
    """Stands in for pyrtlsdr's RtlSdr: hands out complex128 frames like the real one."""

    def __init__(self, frames, fs, fc):
        self.frames, self.i, self.fs, self.fc = frames, 0, fs, fc

    def get_sample_rate(self):
        return self.fs

    def get_center_freq(self):
        return self.fc

    def read_samples(self, n):
        f = self.frames[self.i]
        self.i += 1
        assert len(f) == n
        return f.astype(np.complex128)


def run_rtl(iq, n, fs, fc, window="hanning", psd=False, avg=None):
    src = RtlSamplesDataSource(int(fs), int(fc))
    src.set_fft_size(n)
    src.set_window_type(window)
    src.set_psd_mode(psd)
    if avg is not None:
        src.set_averaging(*avg)
    src.sdr = FakeSdr(iq, fs, fc)
    src.running = True
    rows, bins = [], None
    for _ in range(len(iq)):
        p, bins = src.get_power_levels()
        rows.append(np.array(p, dtype=np.float64, copy=True))
    return np.stack(rows), bins


def run_hackrf(iq, n, fs, fc, psd=False, avg=None):
    src = HackrfSamplesDataSource(int(fs), int(fc))
    src.num_samples = n
    src._allocate_fft_resources()
    src.set_psd_mode(psd)
    if avg is not None:
        src.set_averaging(*avg)
    src.running = True
    rows = []
    for f in iq:
        chunk = np.zeros(65536, dtype=np.complex64)      # READ_CHUNK; newest tail is consumed
        chunk[-n:] = f
        src._sample_queue.put(chunk)
        p, bins = src.get_power_levels()
        rows.append(np.array(p, copy=True))
    return np.stack(rows), bins, src._window.copy()


def gen_rtl():
    fs, fc = 2.048e6, 98e6
    out = {"meta": META, "fs": fs, "fc": fc}
    # config 1: N=1024, Hann, power mode, tone at +250 kHz
    iq = synth.cfg1_frames(b=8, n=1024, fs=fs, seed=0)
    db, bins = run_rtl(iq, 1024, fs, fc)
    out.update(cfg1_iq=iq, cfg1_db=db, cfg1_bins=bins)
    # every window x {power, psd} at N=4096 (cfg-2 style signal), 2 frames each
    iq = synth.cfg2_frames(b=2, n=4096, seed=1)
    out["w_iq"] = iq
    for w in ("hanning", "hamming", "rectangle"):
        for psd in (False, True):
            db, bins = run_rtl(iq, 4096, fs, fc, window=w, psd=psd)
            out[f"w_{w}_{'psd' if psd else 'power'}"] = db
    out["w_bins"] = bins
    # other sizes (Hann, power)
    for n in (512, 2048, 8192):
        iq = synth.cfg2_frames(b=2, n=n, seed=10 + n)
        db, bins = run_rtl(iq, n, fs, fc)
        out[f"n{n}_iq"], out[f"n{n}_db"] = iq, db
    # known-answer frames at N=1024: impulse, zeros, DC, on-bin tone
    n = 1024
    kat = np.zeros((4, n), dtype=np.complex64)
    kat[0, 0] = 1.0
    kat[2, :] = 1.0
    kat[3, :] = np.exp(2j * np.pi * 100 * np.arange(n) / n).astype(np.complex64)
    db, _ = run_rtl(kat, n, fs, fc)
    out.update(kat_iq=kat, kat_db=db)
    # averaging sequences through the source (exp n=8, lin n=4), 12 frames, N=512
    iq = synth.cfg2_frames(b=12, n=512, seed=21)
    out["avg_iq"] = iq
    out["avg_exp8_power"], _ = run_rtl(iq, 512, fs, fc, avg=("exp", 8))
    out["avg_lin4_power"], _ = run_rtl(iq, 512, fs, fc, avg=("lin", 4))
    out["avg_lin4_psd"], _ = run_rtl(iq, 512, fs, fc, psd=True, avg=("lin", 4))
    np.savez_compressed(os.path.join(OUT, "rtl_chain.npz"), **out)


def gen_hackrf():
    fs, fc = 20e6, 2450e6
    out = {"meta": META, "fs": fs, "fc": fc}
    iq = synth.cfg2_frames(b=6, n=1024, seed=31)
    iq = (iq + np.complex64(0.25 - 0.1j)).astype(np.complex64)     # a DC offset to remove
    iq[3] = 0                                                      # silence -> hold last good
    out["iq"] = iq
    db, bins, win = run_hackrf(iq, 1024, fs, fc)
    out.update(mag20=db, bins=bins, window=win)
    out["psd"], _, _ = run_hackrf(iq, 1024, fs, fc, psd=True)
    out["avg_exp4"], _, _ = run_hackrf(iq, 1024, fs, fc, avg=("exp", 4))
    # consume policy (_consume_samples, hackrf_samples.py:254-305): chunks are index ramps so positions are visible
    src = HackrfSamplesDataSource(int(fs), int(fc))
    src.CONSUME_TIMEOUT = 0.05
    def chunk(k):
        return (np.arange(65536, dtype=np.float32) + 100000.0 * k).astype(np.complex64)
    script, heads = [], []
    def read(n):
        r = src._consume_samples(n)
        script.append(("read", n))
        heads.append([-1.0, -1.0, 0] if r is None else [float(r[0].real), float(r[-1].real), len(r)])
    src._sample_queue.put(chunk(0)); script.append(("put", 0))
    read(1024); read(1024); read(4096)
    for k in (1, 2, 3):
        src._sample_queue.put(chunk(k)); script.append(("put", k))
    read(2048); read(65536 - 2048 - 10); read(1024)          # newest chunk only; then reservoir too short -> timeout
    src._sample_queue.put(chunk(4)); script.append(("put", 4))
    read(512)
    out["consume_script"] = np.array([[0 if a == "put" else 1, b] for a, b in script])
    out["consume_heads"] = np.array(heads)
    np.savez_compressed(os.path.join(OUT, "hackrf_chain.npz"), **out)


def gen_averager():
    rng = np.random.default_rng(41)
    frames = rng.random((10, 64)) * 10.0
    out = {"meta": META, "frames": frames}
    for mode, n in (("off", 8), ("exp", 1), ("exp", 8), ("lin", 4), ("lin", 100)):
        a = TraceAverager()
        a.set_mode(mode, n)
        out[f"{mode}{n}"] = np.stack([np.array(a.process(f), copy=True) for f in frames])
    # float32 input is accumulated in float64 (signal_processing.py:47)
    a = TraceAverager()
    a.set_mode("exp", 4)
    f32 = frames.astype(np.float32)
    out["exp4_f32in"] = np.stack([np.array(a.process(f), copy=True) for f in f32])
    np.savez_compressed(os.path.join(OUT, "trace_averager.npz"), **out)


def gen_holds_tare_sweepavg():
    rng = np.random.default_rng(51)
    frames = rng.normal(-60.0, 8.0, (40, 96))
    frames[0, 5] = np.nan
    frames[3, 7] = np.nan
    frames[4, 5] = np.nan
    mw = types.SimpleNamespace(max_power_levels=None, min_power_levels=None, min_hold_enabled=True,
                               status_label=MagicMock(), tare_active=False, baseline_power_levels=None)
    dm = types.SimpleNamespace(max_peak_search_enabled=True, tare_state=TareState(),
                               _update_tare_button_label=lambda *_: None, _clear_tare=lambda: None)
    dp = DataProcessor(mw, dm)
    mx, mn = [], []
    for f in frames[:8]:
        f = f.copy()
        dp._update_max_hold(f)
        dp._update_min_hold(f)
        mx.append(mw.max_power_levels.copy())
        mn.append(mw.min_power_levels.copy())
    out = {"meta": META, "frames": frames, "max_hold": np.stack(mx), "min_hold": np.stack(mn)}
    # tare: collect 32 frames, then subtract
    clean = np.nan_to_num(frames, nan=-70.0)
    dm.tare_state = TareState(collecting=True)
    tared = [np.array(dp._apply_tare(f.copy()), copy=True) for f in clean]
    out.update(tare_in=clean, tare_out=np.stack(tared), tare_baseline=mw.baseline_power_levels)
    # sweep-domain averaging (display_data_processor.py:214-218) with the reference averager
    dp._sweep_averager.set_mode("exp", 4)
    sw = []
    for f in clean[:10]:
        linear = 10.0 ** (f / 10.0)
        sw.append(10.0 * np.log10(np.maximum(dp._sweep_averager.process(linear), 1e-30)))
    out["sweep_avg_exp4"] = np.stack(sw)
    # top-5 peak finder on planted peaks (display_data_processor.py:432-471)
    p = rng.normal(-90.0, 1.0, 512)
    for i, v in ((40, -20.0), (200, -35.0), (207, -36.0), (333, -30.0), (450, -50.0), (500, -25.0)):
        p[i] = v
    fb = np.linspace(1e6, 2e6, 512)
    out["peaks_power"], out["peaks_bins"] = p, fb
    out["peaks"] = np.array(DataProcessor._find_top_peaks(fb, p), dtype=np.float64)
    np.savez_compressed(os.path.join(OUT, "trace_state.npz"), **out)


def gen_stitch():
    start, stop, bin_size = 2400_000_000, 2500_000_000, 100_000
    src = HackRFSweepDataSource(start, stop, bin_size)
    rng = np.random.default_rng(61)
    # hackrf_sweep emits 5 MHz rows out of order (interleaved quarters); 20 MHz steps
    rows, los, his = [], [], []
    for step in range(5):
        base = start + step * 20_000_000
        for q in (0, 2, 1, 3):
            lo = base + q * 5_000_000
            los.append(lo)
            his.append(lo + 5_000_000)
            rows.append(rng.normal(-70.0, 5.0, 50).astype(np.float32))
    def feed():
        for r, lo, hi in zip(rows, los, his):
            line = ", ".join(["2026-01-01", "00:00:00", str(lo), str(hi), "100000.00", "20"]
                             + [f"{v:.2f}" for v in r])
            src._parse(line)
    feed()
    first = src.get_data()          # still NaN: sweep not wrapped yet
    feed()                          # wrap -> stitch of the first pass
    full = src.get_data()
    parsed = np.stack([np.array([float(f"{v:.2f}") for v in r], dtype=np.float32) for r in rows])
    # the same pass as wire data: CSV text (with a malformed line and a short line the reference skips) and -B records
    lines = []
    for r, lo, hi in zip(rows, los, his):
        lines.append(", ".join(["2026-01-01", "00:00:00", str(lo), str(hi), "100000.00", "20"] + [f"{v:.2f}" for v in r]))
    csv_text = "\n".join(lines[:3] + ["garbage, line", "2026-01-01, 00:00:00, 1, 2, x, 20, notafloat"] + lines[3:]) + "\n"
    import struct
    binary = b"".join(struct.pack("<IQQ", 16 + 4 * len(r), lo, hi) + np.asarray(p, dtype="<f4").tobytes()
                      for r, p, lo, hi in zip(rows, parsed, los, his))
    np.savez_compressed(os.path.join(OUT, "sweep_stitch.npz"), meta=META, start=start, stop=stop,
                        bin_size=bin_size, rows=parsed, lo=np.array(los), hi=np.array(his),
                        grid=src.frequency_grid, before_wrap=first, stitched=full,
                        csv_text=np.frombuffer(csv_text.encode(), dtype=np.uint8),
                        binary=np.frombuffer(binary, dtype=np.uint8))


class _QtStub(types.ModuleType):
    """A module whose every attribute is a do-nothing class: lets Qt-bound reference modules be IMPORTED here (PyQt6,
    pyqtgraph are not installed) so that their numpy arithmetic can be executed; nothing Qt does is relied upon."""

    def __getattr__(self, name):
        if name.startswith("__"):
            raise AttributeError(name)
        cls = type(name, (), {"__init__": lambda self, *a, **k: None,
                              "__getattr__": lambda self, n: (lambda *a, **k: None)})
        setattr(self, name, cls)
        return cls


def _install_qt_stubs():
    for m in ("PyQt6", "PyQt6.QtCore", "PyQt6.QtWidgets", "PyQt6.QtGui", "pyqtgraph", "pyqtgraph.exporters"):
        if m not in sys.modules:
            sys.modules[m] = _QtStub(m)
    sys.modules["PyQt6"].QtCore = sys.modules["PyQt6.QtCore"]
    sys.modules["PyQt6"].QtWidgets = sys.modules["PyQt6.QtWidgets"]
    sys.modules["PyQt6"].QtGui = sys.modules["PyQt6.QtGui"]
    sys.modules["pyqtgraph"].exporters = sys.modules["pyqtgraph.exporters"]


def gen_analytics():
    """Colour map (core/export_manager.py:67-84), density histogram (displays/density_display.py:306-319), band power
    (core/marker_manager.py:308-318) and marker snap (:74-99), all by EXECUTING the reference's own methods: the
    Qt-bound modules are imported against stub Qt modules (see _QtStub) and driven with stand-in widgets."""
    _install_qt_stubs()
    from core.export_manager import ExportManager                  # noqa: E402  (reference)
    from core.marker_manager import MarkerManager                  # noqa: E402  (reference)
    from displays.density_display import DensityDisplay            # noqa: E402  (reference)
    from utils.constants import DisplayMode                        # noqa: E402  (reference)
    rng = np.random.default_rng(81)
    # ---- waterfall export colour map: run export_display('png') and capture the bytes handed to QImage -------------
    arr = rng.normal(-70.0, 25.0, (6, 64)).astype(np.float32)
    arr[0, 0] = np.float32(-100.0); arr[0, 1] = np.float32(-20.0)
    wf_min_db, wf_max_db = -100.0, -20.0
    lut = rng.integers(0, 256, (256, 4), dtype=np.uint8)
    captured = {}

    class _QImage:
        class Format:
            Format_RGBA8888 = 0

        def __init__(self, data, w, h, stride, fmt):
            captured["rgba"] = np.frombuffer(data, dtype=np.uint8).reshape(h, w, 4).copy()

        def isNull(self):
            return False

    class _QPixmap:
        @staticmethod
        def fromImage(qi):
            return _QPixmap()

        def save(self, filename, fmt):
            return True

    class _QFileDialog:
        @staticmethod
        def getSaveFileName(*a, **k):
            return "/tmp/tdsa_golden_export.png", ""

    sys.modules["PyQt6.QtGui"].QImage, sys.modules["PyQt6.QtGui"].QPixmap = _QImage, _QPixmap
    sys.modules["PyQt6.QtWidgets"].QFileDialog = _QFileDialog
    wf = types.SimpleNamespace(waterfall_array=arr, wf_min_db=wf_min_db, wf_max_db=wf_max_db, _lut_rgba=lut)
    mw = types.SimpleNamespace(current_stacked_index=DisplayMode.WATERFALL, paused=False, waterfall_widget=wf,
                               status_label=types.SimpleNamespace(setText=lambda t: captured.setdefault("status", t)))
    ExportManager(mw).export_display("png")
    rgba = captured["rgba"]
    # ---- density histogram: DensityDisplay._update_hist on a stand-in object, three frames, decay "medium" ----------
    frames = rng.normal(-80.0, 30.0, (3, 48)).astype(np.float32)
    frames[1, 3] = np.nan; frames[2, 5] = np.float32(-200.2); frames[2, 6] = np.float32(120.0)
    sink = types.SimpleNamespace(setImage=lambda *a, **k: None, setRect=lambda *a, **k: None)
    dd = types.SimpleNamespace(_hist=None, _hist_n_freq=None, _decay=0.96, _img=sink, _last_freq_rect=None,
                               plot_widget=types.SimpleNamespace(setXRange=lambda *a, **k: None))
    dd._ensure_hist = types.MethodType(DensityDisplay._ensure_hist, dd)
    fb = np.linspace(88e6, 108e6, 48)
    hists = []
    for live in frames:
        DensityDisplay._update_hist(dd, live.astype(np.float64), fb)
        hists.append(dd._hist.copy())
    # ---- band power and marker snap: MarkerManager on a stand-in main window -------------------------------------------
    bins = np.linspace(88e6, 108e6, 2048)
    levels = rng.normal(-90.0, 6.0, 2048).astype(np.float32)
    lo, hi = 95e6, 99.5e6
    mm_mw = types.SimpleNamespace(frequency_bins=bins, live_power_levels=levels,
                                  status_label=types.SimpleNamespace(setText=lambda t: None))
    mm = MarkerManager(mm_mw)
    bp = mm._band_power(lo, hi)
    mm._sync_display = lambda name: None
    mm._refresh_status = lambda: None
    snaps_levels, snaps_thr, snaps_exc, snaps_idx = [], [], [], []
    for case in range(8):
        lv = rng.normal(-90.0, 4.0, 1024).astype(np.float32)
        for _ in range(case % 4):
            k = int(rng.integers(3, 1020))
            lv[k] += np.float32(rng.uniform(8, 40))
        if case == 5:
            lv[300:304] = lv.max() + np.float32(5.0)            # flat-topped peak
        if case == 6:
            lv = np.sort(lv)                                     # no peak: argmax fallback
        thr, exc = (-200.0, 6.0) if case % 2 == 0 else (-80.0, 10.0)
        b = np.linspace(88e6, 108e6, 1024)
        mm_mw.frequency_bins, mm_mw.live_power_levels = b, lv
        mm_mw.peak_threshold, mm_mw.peak_excursion = thr, exc
        mm.active_marker = "F1"
        mm.snap_to_peak()
        snaps_levels.append(lv); snaps_thr.append(thr); snaps_exc.append(exc)
        snaps_idx.append(int(np.argmin(np.abs(b - mm.markers["F1"].position))))
    np.savez_compressed(os.path.join(OUT, "analytics.npz"), meta=META, cm_rows=arr, cm_lo=wf_min_db, cm_hi=wf_max_db,
                        cm_lut=lut, cm_rgba=rgba, dens_frames=frames, dens_hists=np.stack(hists),
                        bp_bins=bins, bp_levels=levels, bp_lo=lo, bp_hi=hi, bp_value=bp,
                        snap_levels=np.stack(snaps_levels), snap_thr=np.array(snaps_thr), snap_exc=np.array(snaps_exc),
                        snap_idx=np.array(snaps_idx), executed_reference=True)


def gen_waterfall():
    """Compile _init_buffer/_add_row/_display_view/_calc_lines from the widget's source."""
    path = os.path.join(REF, "displays", "waterfall.py")
    tree = ast.parse(open(path).read())
    want = {"_calc_lines", "_init_buffer", "_add_row", "_display_view"}
    fns = []
    for node in ast.walk(tree):
        if isinstance(node, ast.ClassDef):
            fns += [n for n in node.body if isinstance(n, ast.FunctionDef) and n.name in want]
    assert {f.name for f in fns} == want
    mod = ast.Module(body=[ast.ClassDef(name="Ring", bases=[], keywords=[], body=fns, decorator_list=[])],
                     type_ignores=[])
    ast.fix_missing_locations(mod)
    ns = {"np": np, "_MAX_HISTORY": 2000, "logger": MagicMock()}
    exec(compile(mod, path, "exec"), ns)
    ring = ns["Ring"]()
    ring.seconds_per_row, ring.wf_time_span, ring.wf_min_db = 0.5, 6.0, -100.0   # H = 12
    ring.frequency_bins = np.arange(32)
    ring._init_buffer()
    rng = np.random.default_rng(71)
    rows = rng.normal(-60, 5, (30, 32)).astype(np.float32)
    views = []
    for r in rows:
        ring._add_row(r)
        views.append(ring._display_view().copy())
    np.savez_compressed(os.path.join(OUT, "waterfall_ring.npz"), meta=META, h=ring.history_lines,
                        rows=rows, views=np.stack(views), fill=-100.0)


if __name__ == "__main__":
    os.makedirs(OUT, exist_ok=True)
    gen_rtl()
    gen_hackrf()
    gen_averager()
    gen_holds_tare_sweepavg()
    gen_stitch()
    gen_analytics()
    gen_waterfall()
    for f in sorted(os.listdir(OUT)):
        print(f, os.path.getsize(os.path.join(OUT, f)))
