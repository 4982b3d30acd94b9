"""CPU oracle for the sample-mode hot path.  TEST INFRASTRUCTURE ONLY.

This module is a float64 numpy/scipy restatement of the arithmetic the
reference (CWNE88/topdogspectrumanalyser @ 1e46762) performs on the path
complex64 IQ -> window -> FFT -> fftshift -> |.|^2 -> averager -> 10*log10,
plus the trace state (averager, max/min hold, tare, sweep averaging), the
hackrf_sweep stitch and the waterfall ring.  Every function cites the
reference file:line it follows (paths relative to the reference root).

Who may import it: ``tests/``, ``__graft_entry__.smoke()`` and the
``cpu_baseline`` / ``--impl reference`` legs of ``bench.py``.  Nothing under
``topdogspectrumanalyser_b200/`` imports it; the product path has no CPU
fallback.

Parity pinning: the reference holds NO golden vectors for this path (its tests
mock ``scipy.fft``).  The oracle is therefore pinned
against outputs of the reference's own classes EXECUTED in the build container
(``oracle/make_golden.py`` imports ``/root/reference`` and writes
``tests/golden/*.npz``); ``tests/test_oracle_golden.py`` replays those
fixtures through this module.  The FFT itself is third-party pocketfft
(``scipy.fft`` pinned scipy==1.16.3, ``numpy.fft`` pinned numpy==1.26.4 in the
reference's requirements.txt:50,99; this image has scipy 1.18.1 / numpy 2.3.5).
Truth is always the float64 chain: complex64 IQ up-cast to complex128,
float64 window, complex128 FFT.
"""
from __future__ import annotations

import numpy as np
from scipy import fft as _sfft

# utils/constants.py:152-155
LOG_FLOOR = 1e-12        # magnitude / PSD domain
POWER_LOG_FLOOR = 1e-10  # power domain
# core/display_data_processor.py:214-218,349-351
SWEEP_LINEAR_FLOOR = 1e-30
# utils/constants.py:141
TARE_NUM_SAMPLES = 32

MODE_POWER = "power"   # 10*log10(|X|^2 + 1e-10)          rtl_samples.py:180-184
MODE_PSD = "psd"       # 10*log10(|X|^2/(fs*N) + 1e-12)   rtl_samples.py:175-179
MODE_MAG20 = "mag20"   # 20*log10(|X| + 1e-12)            hackrf_samples.py:383


# --------------------------------------------------------------------------
# windows
# --------------------------------------------------------------------------
def make_window(name: str, n: int) -> np.ndarray:
    """float64 window as the RTL source builds it.

    datasources/rtl_samples.py:199-206 (``np.hanning`` / ``np.hamming`` /
    ``np.ones``; unknown names fall back to hanning).  ``blackman`` is in the
    reference's WindowType enum (utils/constants.py:68-73) without an
    implementation; BASELINE.json asks for it, so it is defined here as
    ``np.blackman`` (symmetric, like the other two) and marked an extension.
    """
    name = name.lower()
    if name in ("hanning", "hann"):
        return np.hanning(n)
    if name == "hamming":
        return np.hamming(n)
    if name in ("rectangle", "rect", "ones"):
        return np.ones(n)
    if name == "blackman":
        return np.blackman(n)
    return np.hanning(n)


def make_window_hackrf(n: int) -> np.ndarray:
    """float32 RMS-normalised Hann, datasources/hackrf_samples.py:311-316."""
    window = np.hanning(n).astype(np.float32)
    window /= np.sqrt(np.mean(window ** 2))
    return window


def freq_bins(n: int, fs: float, fc: float) -> np.ndarray:
    """datasources/rtl_samples.py:188; hackrf_samples.py:318-323."""
    return _sfft.fftshift(_sfft.fftfreq(n, 1 / fs)) + fc


# --------------------------------------------------------------------------
# TraceAverager, utils/signal_processing.py:5-73
# --------------------------------------------------------------------------
class TraceAverager:
    """Restatement of utils/signal_processing.py:12-65 (same aliasing)."""

    def __init__(self):
        self._mode = "off"
        self._n = 1
        self._buffer = None
        self._count = 0

    def set_mode(self, mode: str, n: int) -> None:      # :19-28
        self._mode = mode
        self._n = max(1, n)
        self.reset()

    def reset(self) -> None:                            # :30-33
        self._buffer = None
        self._count = 0

    def process(self, linear_power: np.ndarray) -> np.ndarray:   # :35-61
        if self._mode == "off" or self._n <= 1:
            return linear_power
        if self._buffer is None or self._buffer.shape != linear_power.shape:
            self._buffer = linear_power.astype(np.float64).copy()
            self._count = 1
            return self._buffer
        if self._mode == "exp":
            alpha = 1.0 / self._n
            self._buffer *= (1.0 - alpha)
            self._buffer += alpha * linear_power
        elif self._mode == "lin":
            if self._count < self._n:
                self._count += 1
            self._buffer += (linear_power - self._buffer) / self._count
        return self._buffer

    @property
    def is_active(self) -> bool:                        # :63-65
        return self._mode != "off" and self._n > 1


# --------------------------------------------------------------------------
# RTL-style canonical chain, datasources/rtl_samples.py:169-184
# --------------------------------------------------------------------------
def linear_power_frame(iq: np.ndarray, window: np.ndarray) -> np.ndarray:
    """|fftshift(fft(x*w))|^2 for one frame, float64 (rtl_samples.py:169-173,181)."""
    samples = iq.astype(np.complex128) * window          # :169 (pyrtlsdr hands complex128)
    spectrum = _sfft.fft(samples, n=len(window))         # :170
    spectrum = _sfft.fftshift(spectrum)                  # :173
    return np.abs(spectrum) ** 2                         # :177,181


def power_db_frame(iq, window, mode=MODE_POWER, fs=1.0, averager=None) -> np.ndarray:
    """One call of RtlSamplesDataSource.get_power_levels' arithmetic (:169-184)."""
    p = linear_power_frame(iq, window)
    n = len(window)
    if mode == MODE_PSD:
        p = p / (fs * n)                                 # :177
        if averager is not None:
            p = averager.process(p)                      # :178
        return 10 * np.log10(p + LOG_FLOOR)              # :179
    if mode == MODE_POWER:
        if averager is not None:
            p = averager.process(p)                      # :183
        return 10 * np.log10(p + POWER_LOG_FLOOR)        # :184
    if mode == MODE_MAG20:                               # hackrf_samples.py:372,383
        return 20 * np.log10(np.sqrt(p) + LOG_FLOOR)
    raise ValueError(mode)


def linear_power_batch(iq: np.ndarray, window: np.ndarray, workers: int = 1) -> np.ndarray:
    """Vectorised ``linear_power_frame`` over rows of ``iq[B, N]`` (same pocketfft)."""
    x = iq.astype(np.complex128) * window[None, :]
    spec = _sfft.fftshift(_sfft.fft(x, n=window.shape[0], axis=-1, workers=workers), axes=-1)
    return np.abs(spec) ** 2


def power_db_batch(iq, window, mode=MODE_POWER, fs=1.0, workers: int = 1) -> np.ndarray:
    """Row-by-row rtl_samples.py:169-184 without averaging, float64 ``[B, N]``."""
    p = linear_power_batch(iq, window, workers)
    n = window.shape[0]
    if mode == MODE_PSD:
        return 10 * np.log10(p / (fs * n) + LOG_FLOOR)
    if mode == MODE_POWER:
        return 10 * np.log10(p + POWER_LOG_FLOOR)
    if mode == MODE_MAG20:
        return 20 * np.log10(np.sqrt(p) + LOG_FLOOR)
    raise ValueError(mode)


# --------------------------------------------------------------------------
# HackRF-style chain, datasources/hackrf_samples.py:351-386 (float64 truth)
# --------------------------------------------------------------------------
def hackrf_power_db_frame(iq, window_f32, use_psd=False, fs=1.0, averager=None,
                          dc_estimate=0.0 + 0.0j, dc_alpha=1.0):
    """Returns ``(power_db or None, new_dc_estimate)``.

    ``None`` is the silence hold of hackrf_samples.py:351-355 (caller keeps the
    last good frame).  DC tracker :359-365, window :368, fft/shift :370,
    branches :374-383.  Arithmetic is carried in float64 from the complex64
    input and the float32 (RMS-normalised) window values.
    """
    x = iq.astype(np.complex128)
    if np.mean(np.abs(x) ** 2) < 1e-20:                          # :351
        return None, dc_estimate
    mean = np.mean(x)                                            # :360
    dc = (1.0 - dc_alpha) * dc_estimate + dc_alpha * mean        # :361-364
    x = (x - dc) * window_f32.astype(np.float64)                 # :365,368
    mag = np.abs(_sfft.fftshift(_sfft.fft(x)))                   # :370,372
    n = x.shape[0]
    if use_psd:
        psd = (mag ** 2) / (fs * n)                              # :375
        if averager is not None:
            psd = averager.process(psd)
        return 10 * np.log10(psd + LOG_FLOOR), dc                # :377
    if averager is not None and averager.is_active:
        p = averager.process(mag ** 2)                           # :379-380
        return 10 * np.log10(p + POWER_LOG_FLOOR), dc            # :381
    return 20 * np.log10(mag + LOG_FLOOR), dc                    # :383


# --------------------------------------------------------------------------
# trace holds, core/display_data_processor.py:371-395,473-480
# --------------------------------------------------------------------------
def nan_safe(arr: np.ndarray, fill: float) -> np.ndarray:
    """:473-480 — identity return for clean arrays is part of the contract."""
    if not np.any(np.isnan(arr)):
        return arr
    out = arr.copy()
    out[np.isnan(out)] = fill
    return out


def max_hold_update(hold, power_db):
    """:371-382 with hold enabled. Returns the (possibly aliased) hold buffer."""
    if hold is None or hold.shape != power_db.shape:
        return nan_safe(power_db, -500.0)
    np.fmax(hold, power_db, out=hold)
    return hold


def min_hold_update(hold, power_db):
    """:384-395 with hold enabled."""
    if hold is None or hold.shape != power_db.shape:
        return nan_safe(power_db, 500.0)
    np.fmin(hold, power_db, out=hold)
    return hold


def sweep_average_db(power_db: np.ndarray, averager: TraceAverager):
    """core/display_data_processor.py:211-218. Returns None for all-NaN frames."""
    if np.all(np.isnan(power_db)):
        return None
    if not averager.is_active:
        return power_db
    linear = 10.0 ** (power_db / 10.0)
    return 10.0 * np.log10(np.maximum(averager.process(linear), SWEEP_LINEAR_FLOOR))


class Tare:
    """core/display_data_processor.py:329-369 without the Qt label updates."""

    def __init__(self):
        self.collecting = False
        self.buffer = None
        self.count = 0
        self.active = False
        self.baseline = None

    def start(self):
        self.collecting, self.buffer, self.count = True, None, 0

    def apply(self, power_db: np.ndarray) -> np.ndarray:
        if self.collecting:
            linear = 10.0 ** (power_db / 10.0)                       # :336
            if self.buffer is None or self.buffer.shape != linear.shape:
                self.buffer = linear.copy()
                self.count = 1
            else:
                self.buffer += linear
                self.count += 1
            if self.count >= TARE_NUM_SAMPLES:                       # :350
                avg = self.buffer / self.count
                self.baseline = 10.0 * np.log10(np.maximum(avg, SWEEP_LINEAR_FLOOR))
                self.active = True
                self.collecting, self.buffer, self.count = False, None, 0
        if self.active and self.baseline is not None:
            if power_db.shape != self.baseline.shape:                # :361-364
                self.active, self.baseline = False, None
            else:
                power_db = power_db - self.baseline                  # :366
        return power_db


# --------------------------------------------------------------------------
# hackrf_sweep stitch, datasources/hackrf_sweep.py:32-40,135-168
# --------------------------------------------------------------------------
def sweep_grid(start_hz: int, stop_hz: int, bin_size: int) -> np.ndarray:
    """:32-40."""
    num_bins = int((stop_hz - start_hz) / bin_size)
    return np.linspace(start_hz, stop_hz, num_bins)


def row_bin_centres(hz_low: float, hz_high: float, k: int) -> np.ndarray:
    """:159-164."""
    bw = (hz_high - hz_low) / k
    return np.arange(hz_low + bw / 2, hz_high, bw)


def stitch_rows(rows_db, rows_lo, rows_hi, grid: np.ndarray) -> np.ndarray:
    """One completed sweep: accumulate (x, y), argsort, np.interp onto the grid (:150-166).

    ``rows_db`` is a sequence of float32 rows in arrival order.
    """
    xs, ys = [], []
    for row, lo, hi in zip(rows_db, rows_lo, rows_hi):
        row = np.asarray(row, dtype=np.float32)
        xs.extend(row_bin_centres(lo, hi, len(row)))
        ys.extend(row)
    order = np.argsort(xs)
    return np.interp(grid, np.array(xs)[order], np.array(ys)[order])


# --------------------------------------------------------------------------
# waterfall ring, displays/waterfall.py:163-180
# --------------------------------------------------------------------------
class WaterfallRing:
    def __init__(self, h: int, w: int, fill: float):
        self.h = h
        self.buf = np.full((2 * h, w), fill, dtype=np.float32)   # :168
        self.ptr = 0

    def add_row(self, row):                                      # :173-177
        self.ptr = (self.ptr - 1) % self.h
        self.buf[self.ptr] = row
        self.buf[self.ptr + self.h] = row

    def view(self):                                              # :179-180
        return self.buf[self.ptr:self.ptr + self.h]


# --------------------------------------------------------------------------
# config 3: Welch average + peak hold built from the reference primitives
# --------------------------------------------------------------------------
def welch_avg_peak_db(stream: np.ndarray, window: np.ndarray, hop: int,
                      floor: float = POWER_LOG_FLOOR):
    """a1 per segment -> TraceAverager('lin', n=nseg) -> dB; peak = fmax of per-segment dB.

    SURVEY.md section 8(d) config 3: the reference has no Welch routine; this
    is its per-frame arithmetic (rtl_samples.py:169-184) applied to overlapping
    segments, combined with signal_processing.py:56-59 and
    display_data_processor.py:371-382.
    """
    n = window.shape[0]
    nseg = (stream.shape[0] - n) // hop + 1
    avg = TraceAverager()
    avg.set_mode("lin", max(nseg, 2))
    peak = None
    out = None
    for s in range(nseg):
        p = linear_power_frame(stream[s * hop:s * hop + n], window)
        db = 10 * np.log10(p + floor)
        peak = max_hold_update(peak, db.copy())
        out = avg.process(p)
    return 10 * np.log10(out + floor), peak
