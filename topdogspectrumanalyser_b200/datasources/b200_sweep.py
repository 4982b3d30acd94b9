"""B200SweepDataSource — an IQ-in sweep source: what ``HackRFSweepDataSource`` gets from the external ``hackrf_sweep``
binary (datasources/hackrf_sweep.py:56-74: per-step FFTs and dB rows on stdout) is computed here from raw IQ on the GPU.

Same constructor, attributes and methods as the reference class (``start_freq``, ``stop_freq``, ``bin_size``,
``frequency_grid``, ``full_power_array``, ``lna_gain`` / ``vga_gain`` / ``amp_enabled``, ``is_running``, ``sweep_rate``,
``lock``, ``thread``; ``start`` / ``stop`` / ``get_data`` / ``get_number_of_points`` / ``set_gains`` / ``set_amplifier``), so
``DataProcessor._process_sweep_data`` (core/display_data_processor.py:185-228) and
``SourceManager._initialise_hackrf_sweep`` (core/source_manager.py:510-523) use it unchanged.

A sweep = the span cut into sub-bands of ``band_hz`` (the tuner's sample rate); for each sub-band the tuner is moved to
its centre and ``frames`` x ``n_fft`` samples are read; when the last sub-band is in, one launch forms every sub-band's
mean power row (kernel 1 per frame, linear mean over the frames, dB: ``tdsa_group_avg_db``) and ``tdsa_stitch`` puts
the rows on the fixed grid with the geometry of ``_parse`` (:150-166: bin centres ``lo + bw/2 + i bw``, ``np.interp`` onto
``linspace(start, stop, int((stop - start) / bin_size))``).  The external binary's own window / bin selection is not
reproducible from the reference repo, so the per-sub-band stage is defined as kernel 1; the stitch
is compared bit for bit.
"""
from __future__ import annotations

import logging
import threading
import time
from typing import Optional

import numpy as np
import torch

from .base import SweepDataSource

logger = logging.getLogger(__name__)


class B200SweepDataSource(SweepDataSource):
    def __init__(self, start_freq: float, stop_freq: float, bin_size: int, feed=None, band_hz: float = 20e6,
                 frames: int = 16, n_fft: Optional[int] = None, precision: str = "f64", device: Optional[str] = None):
        super().__init__()
        self.start_freq = int(start_freq)
        self.stop_freq = int(stop_freq)
        self.bin_size = int(bin_size)
        self.lna_gain = 20                       # hackrf_sweep.py:16-18
        self.vga_gain = 20
        self.amp_enabled = True
        self.is_running = False
        self.sweep_complete = False
        self.full_power_array = np.array([])
        self.lock = threading.Lock()
        self.thread: Optional[threading.Thread] = None
        self.sweep_rate: Optional[float] = None
        self.feed = feed                         # tuner: sample_rate / center_freq setters + read_samples(n)
        self.band_hz, self.frames, self.precision = float(band_hz), int(frames), precision
        self._n_fft_arg, self._device_name = n_fft, device
        self._owns_feed = False
        self._halt = threading.Event()
        self._sweep = None
        self._create_frequency_grid()

    # ---- geometry (hackrf_sweep.py:32-40) --------------------------------------------------------------
    def _create_frequency_grid(self):
        num_bins = int((self.stop_freq - self.start_freq) / self.bin_size)
        self.frequency_grid = np.linspace(self.start_freq, self.stop_freq, num_bins)
        with self.lock:
            self.full_power_array = np.full(num_bins, np.nan)      # NaN = not swept yet
        self._sweep = None

    @property
    def n_fft(self) -> int:
        """FFT size per sub-band: the power of two whose bin width is closest to (not coarser than) ``bin_size``."""
        if self._n_fft_arg:
            return int(self._n_fft_arg)
        n = 1
        while self.band_hz / n > self.bin_size and n < 8192:
            n *= 2
        return max(n, 64)

    @property
    def n_bands(self) -> int:
        return max(1, int(np.ceil((self.stop_freq - self.start_freq) / self.band_hz)))

    def band_centres(self) -> np.ndarray:
        return self.start_freq + self.band_hz * (np.arange(self.n_bands) + 0.5)

    # ---- device side -----------------------------------------------------------------------------------
    def _ensure_sweep(self):
        if self._sweep is None:
            if not torch.cuda.is_available():
                raise RuntimeError("B200 backend needs a CUDA device; there is no CPU fallback")
            from ..engine import SpectrumPlan
            dev = torch.device(self._device_name or f"cuda:{torch.cuda.current_device()}")
            n = self.n_fft
            self._plan = SpectrumPlan(n, "hanning", mode="power", precision=self.precision, fs=self.band_hz, device=dev)
            self._pin = torch.empty((self.n_bands, self.frames, n), dtype=torch.complex64).pin_memory()
            self._dev_iq = torch.empty((self.n_bands, self.frames, n), dtype=torch.complex64, device=dev)
            self._lo = torch.tensor([self.start_freq + b * self.band_hz for b in range(self.n_bands)], dtype=torch.float64,
                                    device=dev)
            self._sweep = dev
        return self._sweep

    def process_sweep(self, iq: np.ndarray) -> np.ndarray:
        """``iq[n_bands, frames, n_fft]`` (one completed sweep) -> the stitched grid; also publishes it to ``get_data``."""
        from ..engine import stitch
        self._ensure_sweep()
        self._pin.numpy()[:] = iq
        self._dev_iq.copy_(self._pin, non_blocking=True)
        rows = self._plan.group_avg_db(self._dev_iq)
        m = len(self.frequency_grid)
        grid = stitch(rows, self._lo, self.band_hz, float(self.start_freq), float(self.stop_freq), m).cpu().numpy()
        with self.lock:
            self.full_power_array = grid
        self.sweep_complete = True
        return grid

    # ---- acquisition ------------------------------------------------------------------------------------
    def _open_own_feed(self):
        from .feeds import HackrfDeviceFeed
        feed = HackrfDeviceFeed(self.band_hz, float(self.band_centres()[0]), self.lna_gain, self.vga_gain, self.amp_enabled)
        feed.open()
        return feed

    def _tune(self, centre: float) -> None:
        if hasattr(self.feed, "retune"):
            self.feed.retune(None, centre)
        else:
            self.feed.center_freq = centre

    def acquire_sweep(self) -> Optional[np.ndarray]:
        """Step the tuner through the sub-bands and read ``frames * n_fft`` samples at each; None if stopped midway."""
        n, out = self.n_fft, np.empty((self.n_bands, self.frames, self.n_fft), dtype=np.complex64)
        for b, fc in enumerate(self.band_centres()):
            if self._halt.is_set():
                return None
            self._tune(float(fc))
            x = self.feed.read_samples(self.frames * n)
            if x is None:
                return None
            out[b] = np.asarray(x, dtype=np.complex64).reshape(self.frames, n)
        return out

    def _sweep_loop(self):
        t_prev = time.monotonic()
        while self.is_running and not self._halt.is_set():
            try:
                iq = self.acquire_sweep()
                if iq is None:
                    continue
                self.process_sweep(iq)
                now = time.monotonic()
                self.sweep_rate = 1.0 / max(now - t_prev, 1e-9)          # what the reference parses from stderr (:113-124)
                t_prev = now
            except Exception as e:                                         # noqa: BLE001
                if self.is_running:
                    logger.error("Sweep error: %s", e)
                    time.sleep(0.05)

    def start(self, frequency=None):
        if frequency:
            self.start_freq, self.stop_freq = int(frequency.start), int(frequency.stop)
            self._create_frequency_grid()
        if self.is_running:
            self.stop()
        try:
            if self.feed is None:
                self.feed, self._owns_feed = self._open_own_feed(), True
            self._ensure_sweep()
            self._halt.clear()
            self.is_running = True
            self.thread = threading.Thread(target=self._sweep_loop, daemon=True, name="B200-Sweep")
            self.thread.start()
        except Exception as e:
            self.is_running = False
            logger.error("Error starting B200 sweep: %s", e)
            raise RuntimeError(f"B200 sweep start failed: {e}") from e

    def stop(self):
        self.is_running = False
        self._halt.set()
        t, self.thread = self.thread, None
        if t is not None and t.is_alive():
            t.join(timeout=2.0)
        if self._owns_feed and self.feed is not None:
            try:
                self.feed.close()
            except Exception as e:                                         # noqa: BLE001
                logger.debug("error closing feed: %s", e)
            self.feed, self._owns_feed = None, False

    # ---- what the GUI tick reads (hackrf_sweep.py:224-233) ---------------------------------------------------
    def get_data(self):
        with self.lock:
            if self.full_power_array.size == 0:
                return np.array([])
            return self.full_power_array.copy()

    def get_number_of_points(self):
        with self.lock:
            return len(self.full_power_array)

    def set_gains(self, lna_gain=None, vga_gain=None):
        if lna_gain is not None:
            self.lna_gain = int(lna_gain)
        if vga_gain is not None:
            self.vga_gain = int(vga_gain)
        if hasattr(self.feed, "set_gains"):
            self.feed.set_gains(lna_gain, vga_gain)

    def set_amplifier(self, enabled: bool):
        self.amp_enabled = bool(enabled)
        if hasattr(self.feed, "set_amplifier"):
            self.feed.set_amplifier(enabled)


class SyntheticTunerFeed:
    """A tuner with pyrtlsdr's surface for tests and demos: AWGN plus one tone per MHz-aligned carrier list, seen through
    whatever centre frequency is currently set."""

    def __init__(self, sample_rate: float, carriers_hz=(), amp: float = 0.5, seed: int = 0):
        self.sample_rate, self.center_freq = float(sample_rate), 0.0
        self.carriers, self.amp = list(carriers_hz), amp
        self.rng = np.random.default_rng(seed)
        self.t0 = 0

    def get_sample_rate(self):
        return self.sample_rate

    def get_center_freq(self):
        return self.center_freq

    def read_samples(self, n: int) -> np.ndarray:
        s = np.float32(np.sqrt(0.5))
        x = (self.rng.standard_normal(n, dtype=np.float32) * s + 1j * self.rng.standard_normal(n, dtype=np.float32) * s)
        t = np.arange(self.t0, self.t0 + n, dtype=np.float64) / self.sample_rate
        for f in self.carriers:
            off = f - self.center_freq
            if abs(off) < self.sample_rate / 2:
                x = x + self.amp * np.exp(2j * np.pi * off * t)
        self.t0 += n
        return x.astype(np.complex64)

    def close(self):
        pass
