"""B200SampleDataSource — drop-in ``SampleDataSource`` whose DSP runs in libtdsa.so on a B200.

It mirrors ``RtlSamplesDataSource`` (datasources/rtl_samples.py:16-255; ``style="rtl"``) and the
DSP of ``HackrfSamplesDataSource.get_power_levels`` (datasources/hackrf_samples.py:339-386;
``style="hackrf"``): same attribute names, setters, return types and error convention, so
``core.display_data_processor.DataProcessor._process_sample_data`` (:153-183) can call it unchanged.
Samples come from any *feed* exposing pyrtlsdr's call surface
(``read_samples(n)``, ``get_sample_rate()``, ``get_center_freq()``, optional ``close()``).
"""
from __future__ import annotations

import logging
from typing import Optional

import numpy as np
import torch

from .. import _lib as L
from ..engine import LOG_FLOOR, POWER_LOG_FLOOR, SpectrumPlan, TraceState
from .base import IN_REFERENCE_APP, AveragerSettings, SampleDataSource

logger = logging.getLogger(__name__)


class SyntheticIQFeed:
    """Seeded complex64 IQ with pyrtlsdr's surface: AWGN plus one tone (for tests and demos)."""

    def __init__(self, sample_rate: float = 2.048e6, centre_freq: float = 98e6, tone_hz: float = 250e3,
                 tone_amp: float = 0.5, seed: int = 0):
        self.sample_rate, self.center_freq = float(sample_rate), float(centre_freq)
        self.tone_hz, self.tone_amp = tone_hz, tone_amp
        self.rng = np.random.default_rng(seed)
        self.t0 = 0
        self.gain = "auto"

    def get_sample_rate(self):
        return self.sample_rate

    def get_center_freq(self):
        return self.center_freq

    def read_samples(self, n: int) -> np.ndarray:
        s = np.float32(np.sqrt(0.5))
        x = np.empty(n, dtype=np.complex64)
        x.real = self.rng.standard_normal(n, dtype=np.float32) * s
        x.imag = self.rng.standard_normal(n, dtype=np.float32) * s
        t = np.arange(self.t0, self.t0 + n, dtype=np.float64)
        x += (self.tone_amp * np.exp(2j * np.pi * self.tone_hz * t / self.sample_rate)).astype(np.complex64)
        self.t0 += n
        return x

    def close(self):
        pass


class ReplayFeed:
    """Feed that hands out pre-recorded frames one per ``read_samples`` call (file replay, tests).

    ``dtype`` is what the device library would return: complex128 for pyrtlsdr, complex64 for pyhackrf."""

    def __init__(self, frames, sample_rate: float, centre_freq: float, dtype=np.complex128):
        self.frames, self.i, self.dtype = frames, 0, dtype
        self.fs, self.fc = float(sample_rate), float(centre_freq)
        self.sample_rate, self.center_freq, self.gain = self.fs, self.fc, "auto"

    def get_sample_rate(self):
        return self.fs

    def get_center_freq(self):
        return self.fc

    def read_samples(self, n: int):
        f = self.frames[self.i]
        self.i += 1
        if len(f) != n:
            raise ValueError(f"replay frame has {len(f)} samples, {n} requested")
        return np.asarray(f).astype(self.dtype)

    def close(self):
        pass


class HackrfChunkFeed:
    """Feed with the reference HackRF source's consume policy (datasources/hackrf_samples.py:28-29,254-305).

    A reader thread ``put``s 65 536-sample chunks into a depth-4 queue (oldest dropped when full, :221-237);
    ``read_samples(n)`` drains the queue keeping only the newest chunk as reservoir and hands out the LAST n
    samples of it, shrinking the reservoir from the end; returns ``None`` after ``timeout`` seconds without data.
    """
    READ_CHUNK, MAX_QUEUE_SIZE, CONSUME_TIMEOUT = 65536, 4, 0.5

    def __init__(self, sample_rate: float, centre_freq: float, timeout: Optional[float] = None):
        import queue
        self._queue_mod = queue
        self.sample_rate, self.center_freq = float(sample_rate), float(centre_freq)
        self._q = queue.Queue(maxsize=self.MAX_QUEUE_SIZE)
        self._reservoir = np.array([], dtype=np.complex64)
        self.timeout = self.CONSUME_TIMEOUT if timeout is None else timeout
        self.stats = {"samples_dropped": 0, "queue_overflows": 0}
        self.gain = None

    def get_sample_rate(self):
        return self.sample_rate

    def get_center_freq(self):
        return self.center_freq

    def put(self, chunk: np.ndarray) -> None:
        """Producer side: non-blocking put, drop the OLDEST chunk on overflow (:221-237)."""
        try:
            self._q.put(chunk, block=False)
        except self._queue_mod.Full:
            try:
                old = self._q.get_nowait()
                self.stats["samples_dropped"] += len(old)
                self.stats["queue_overflows"] += 1
            except self._queue_mod.Empty:
                pass
            self._q.put(chunk, block=False)

    def read_samples(self, count: int):
        import time
        if count <= 0:
            return np.array([], dtype=np.complex64)
        fresh = None
        while True:                                               # :269-276
            try:
                fresh = self._q.get_nowait()
            except self._queue_mod.Empty:
                break
        if fresh is not None:
            self._reservoir = fresh
        if len(self._reservoir) >= count:                         # :281-284
            result = self._reservoir[-count:]
            self._reservoir = self._reservoir[:-count]
            return result
        start = time.time()
        while len(self._reservoir) < count:                       # :287-301
            if time.time() - start > self.timeout:
                return None
            try:
                chunk = self._q.get(timeout=0.01)
                while True:
                    try:
                        chunk = self._q.get_nowait()
                    except self._queue_mod.Empty:
                        break
                self._reservoir = chunk
            except self._queue_mod.Empty:
                continue
        result = self._reservoir[-count:]
        self._reservoir = self._reservoir[:-count]
        return result

    def close(self):
        pass


class B200SampleDataSource(SampleDataSource):
    """Sample-mode source with the window -> FFT -> |.|^2 -> avg -> dB chain on the GPU."""

    _DC_ALPHA = 1.0          # hackrf_samples.py:32

    def __init__(self, sample_rate: int, centre_freq: int, feed=None, style: str = "rtl",
                 precision: str = "f64", device: Optional[str] = None, out_dtype=np.float64):
        super().__init__(sample_rate, centre_freq)
        if style not in ("rtl", "hackrf"):
            raise ValueError("style must be 'rtl' or 'hackrf'")
        self.style = style
        self.precision = precision
        self.out_dtype = np.float32 if style == "hackrf" else out_dtype    # hackrf path returns float32 under numpy>=2
        self.fft_size = 1024                     # rtl_samples.py:20 / hackrf_samples.py:38
        self.window_type = "hanning"
        self.use_psd = False
        self.running = False
        self.last_sample_rate = sample_rate
        self.sdr = feed                          # same attribute name as RtlSamplesDataSource
        self._gain = "auto"
        self._flush_reads_remaining = 0
        self._device_name = device
        self._plan: Optional[SpectrumPlan] = None
        self._state: Optional[TraceState] = None
        self._freq_bins = None
        self._freq_key = None
        self._last_good_power = None
        self._dc_state = None
        # _averager: the reference's TraceAverager inside the app (unused for arithmetic), settings-only otherwise
        self._avg_settings = AveragerSettings(on_change=self._on_averager_reset)

    # ---- properties the managers read (core/source_manager.py:344,361-369,685,794) ---------
    @property
    def num_samples(self) -> int:                # hackrf naming (test_fft_size_detection.py:30-33)
        return self.fft_size

    @property
    def is_running(self) -> bool:
        return self.running

    @property
    def sample_count(self) -> int:
        return self.fft_size

    @sample_count.setter
    def sample_count(self, value: int):
        self.set_fft_size(value)

    # ---- lifecycle -------------------------------------------------------------------------
    def _device(self) -> torch.device:
        if not torch.cuda.is_available():
            raise RuntimeError("B200 backend needs a CUDA device; there is no CPU fallback")
        return torch.device(self._device_name or f"cuda:{torch.cuda.current_device()}")

    def _ensure_plan(self) -> SpectrumPlan:
        if self._plan is None or self._plan.n_fft != self.fft_size:
            if self._plan is not None:
                self._plan.close()
            dev = self._device()
            mode, floor = self._mode_and_floor()
            norm = "rms" if self.style == "hackrf" else "none"
            window = "hanning" if self.style == "hackrf" else self.window_type
            self._plan = SpectrumPlan(self.fft_size, window, norm, mode, floor, float(self._fs()), self.precision, dev)
            old = self._state
            self._state = TraceState(self.fft_size, dev)
            if old is not None:
                self._state.avg_mode, self._state.avg_n = old.avg_mode, old.avg_n
            else:
                self._state.avg_mode, self._state.avg_n = self._avg_settings.mode, self._avg_settings.n
            self._x_dev = torch.empty(self.fft_size, dtype=torch.complex64, device=dev)
            self._pin_in = torch.empty(self.fft_size, dtype=torch.complex64).pin_memory()
            self._dc_state = torch.zeros(2, dtype=torch.float64, device=dev)
        return self._plan

    def _fs(self) -> float:
        return float(self.sample_rate or 1.0)

    def _mode_and_floor(self):
        if self.use_psd:
            return "psd", LOG_FLOOR                                   # rtl_samples.py:179 / hackrf_samples.py:377
        if self.style == "hackrf" and not self._avg_settings.is_active:
            return "mag20", LOG_FLOOR                                 # hackrf_samples.py:383
        return "power", POWER_LOG_FLOOR                               # rtl_samples.py:184 / hackrf_samples.py:381

    def start(self, frequency=None):
        """rtl_samples.py:30-58: span -> sample rate, centre -> centre_freq; RuntimeError on failure."""
        if frequency:
            self.centre_freq = int(frequency.centre)
            self.sample_rate = int(frequency.span)
        if self.running:
            return
        try:
            if self.sdr is None:
                raise RuntimeError("no IQ feed attached (pass feed=... or set .sdr)")
            for attr, val in (("sample_rate", self.sample_rate), ("center_freq", self.centre_freq)):
                try:
                    setattr(self.sdr, attr, val)
                except Exception:
                    pass
            actual = self.sdr.get_sample_rate()
            self.sample_rate = actual
            self.last_sample_rate = actual
            self._ensure_plan()
            self.running = True
        except Exception as e:
            self.running = False
            logger.error("B200 source initialisation failed: %s", e)
            raise RuntimeError(f"B200 source initialisation failed: {e}")

    def pause(self):
        self.running = False

    def resume(self):
        if self.sdr is not None:
            self.running = True

    def stop(self):
        if self.sdr is not None and hasattr(self.sdr, "close"):
            try:
                self.sdr.close()
            except Exception as e:
                logger.error("Error closing feed: %s", e)
        self.running = False

    # ---- configuration (same names as the reference sources) ---------------------------------
    def set_window_type(self, window_type: str):
        """rtl_samples.py:199-206; unknown names fall back to hanning."""
        name = window_type.lower()
        self.window_type = name if name in ("hanning", "hamming", "rectangle", "blackman") else "hanning"
        if self._plan is not None and self.style == "rtl":
            self._plan.set_window(self.window_type)

    def set_fft_size(self, fft_size: int):
        """rtl_samples.py:208-215: window returns to Hann, averager resets."""
        if fft_size == self.fft_size:
            return
        self.fft_size = int(fft_size)
        self.window_type = "hanning"
        self._freq_key = None
        self._last_good_power = None
        self._on_averager_reset()
        if self._plan is not None:
            self._ensure_plan()

    def set_num_samples(self, num_samples: int):
        if num_samples <= 0:
            raise ValueError("num_samples must be positive")
        self.set_fft_size(num_samples)

    def set_psd_mode(self, enabled: bool):
        self.use_psd = bool(enabled)

    def set_gain(self, gain) -> None:
        self._gain = gain
        if self.sdr is not None:
            try:
                self.sdr.gain = gain
            except Exception as e:
                logger.error("Error setting gain: %s", e)

    def set_averaging(self, mode: str, n: int) -> None:           # base.py:158-165
        self._avg_settings.set_mode(mode, n)
        try:
            self._averager.set_mode(mode, n)
        except Exception:
            pass
        if self._state is not None:
            self._state.set_averaging(mode, n)

    def reset_averaging(self) -> None:                            # base.py:167-169
        self._on_averager_reset()
        try:
            self._averager.reset()
        except Exception:
            pass

    def _on_averager_reset(self) -> None:
        if self._state is not None:
            self._state.reset_averaging()

    def update_centre_frequency(self, centre_freq: float):
        """rtl_samples.py:85-105 (flush count included)."""
        if not self.running:
            return
        centre_freq = int(centre_freq)
        if centre_freq == self.centre_freq:
            return
        self.centre_freq = centre_freq
        try:
            self.sdr.center_freq = centre_freq
            self._flush_reads_remaining = max(3, int(0.006 * self.sample_rate / self.fft_size))
        except Exception as e:
            raise RuntimeError(f"Error updating centre frequency: {e}")

    def update_sample_rate(self, sample_rate: float):
        sample_rate = int(sample_rate)
        if sample_rate == self.last_sample_rate:
            return
        if self.running and self.sdr is not None:
            try:
                self.sdr.sample_rate = sample_rate
                actual = self.sdr.get_sample_rate()
                self.sample_rate = actual
                self.last_sample_rate = actual
            except Exception as e:
                raise RuntimeError(f"Error updating sample rate: {e}")
        else:
            self.sample_rate = sample_rate

    def update_frequency(self, sample_rate: float, centre_freq: float):
        sample_rate, centre_freq = int(sample_rate), int(centre_freq)
        if sample_rate != self.last_sample_rate:
            self.update_sample_rate(sample_rate)
        if centre_freq != self.centre_freq:
            self.update_centre_frequency(centre_freq)

    # ---- the hot call -----------------------------------------------------------------------
    def _zeros(self):
        return np.zeros(self.fft_size), np.linspace(self.centre_freq - self.sample_rate / 2,
                                                    self.centre_freq + self.sample_rate / 2, self.fft_size)

    def _bins(self, fs: float, fc: float) -> np.ndarray:
        """rtl_samples.py:188 — cached instead of rebuilt every frame (same values)."""
        key = (self.fft_size, fs, fc)
        if self._freq_key != key:
            self._freq_bins = np.fft.fftshift(np.fft.fftfreq(self.fft_size, 1 / fs)) + fc
            self._freq_key = key
        return self._freq_bins

    def read_samples_only(self):
        if not self.running or self.sdr is None:
            return None
        try:
            samples = np.asarray(self.sdr.read_samples(self.fft_size))
            self._store_raw(samples.copy())
            return self._last_raw_samples
        except Exception as e:
            logger.error("Error reading samples: %s", e)
            return None

    def get_power_levels(self):
        """One frame: returns a FRESH ``(power_db[N], freq_bins[N])`` (never a view of a reused buffer)."""
        if not self.running:
            return self._zeros()
        try:
            row, bins, held = self._device_row()
            if held is not None:
                return held, bins
            power_db = row[0].cpu().numpy().astype(self.out_dtype)
            if self.style == "hackrf":
                self._last_good_power = power_db
            return power_db, bins
        except Exception as e:
            logger.error("Error computing power levels: %s", e)
            return self._zeros()

    def get_power_levels_device(self):
        """Same frame as :meth:`get_power_levels` but left on the GPU: ``(float32 CUDA row [1, N], freq_bins)``.

        Used by ``frame_pipeline.B200FramePipeline`` so cal/tare/holds/peaks run without a host round trip.
        Returns ``(None, bins)`` when the HackRF-style silence hold applies (caller keeps its last row)."""
        row, bins, held = self._device_row()
        return (None if held is not None else row), bins

    def _device_row(self):
        """Read one frame from the feed and run the fused kernel. Returns (device row, bins, held_host_row)."""
        fs = self.sdr.get_sample_rate()
        fc = self.sdr.get_center_freq()
        if self._flush_reads_remaining > 0:
            for _ in range(self._flush_reads_remaining):
                self.sdr.read_samples(self.fft_size)
            self._flush_reads_remaining = 0
        samples = self.sdr.read_samples(self.fft_size)
        if samples is None:                                    # feed timed out (hackrf_samples.py:351)
            if self._last_good_power is not None:
                return None, self._bins(fs, fc), self._last_good_power
            return None, self._bins(fs, fc), np.zeros(self.fft_size)
        samples = np.asarray(samples)
        self._store_raw(samples.copy())
        plan = self._ensure_plan()
        mode, floor = self._mode_and_floor()
        if (plan.mode, plan.log_floor, plan.fs) != (mode, floor, float(fs)):
            plan.set_mode(mode, floor, float(fs))
        self._pin_in.numpy()[:] = samples                      # complex128 feeds narrow to complex64 here
        self._x_dev.copy_(self._pin_in, non_blocking=True)
        x = self._x_dev.view(1, self.fft_size)
        st = self._state
        st.avg_mode, st.avg_n = self._avg_settings.mode, self._avg_settings.n
        if self.style == "hackrf":
            if st.averaging or self.use_psd:
                # |X|^2 [/(fs*N)] -> averager -> 10*log10 (hackrf_samples.py:374-381)
                db, silent = plan.psd_db_avg_hold_dc(x, st, self._dc_state, self._DC_ALPHA, last_only=True)
            else:
                db, silent = plan.psd_db_dc(x, self._dc_state, self._DC_ALPHA)       # :383
            if int(silent.item()):
                if self._last_good_power is not None:                                # :351-355
                    return None, self._bins(fs, fc), self._last_good_power
                return None, self._bins(fs, fc), np.zeros(self.fft_size)
        elif st.averaging:
            db = plan.psd_db_avg_hold(x, st, last_only=True)
        else:
            db = plan.psd_db(x)
        return db, self._bins(fs, fc), None

    def get_power_levels_batch(self, n_frames: int):
        """Extension for streaming use: ``n_frames`` consecutive frames in one launch -> ``[B, N]`` float32."""
        if not self.running:
            raise RuntimeError("source not running")
        fs, fc = self.sdr.get_sample_rate(), self.sdr.get_center_freq()
        plan = self._ensure_plan()
        mode, floor = self._mode_and_floor()
        plan.set_mode(mode, floor, float(fs))
        iq = np.ascontiguousarray(np.asarray(self.sdr.read_samples(self.fft_size * n_frames), dtype=np.complex64)
                                  .reshape(n_frames, self.fft_size))
        self._store_raw(iq[-1].copy())
        return plan.psd_db_host(iq), self._bins(fs, fc)


def register_with_source_manager(source_type: str = "b200_samples", display_name: str = "B200 Samples",
                                 like: str = "rtl_samples"):
    """Add this backend to the reference's registry (core/source_manager.py:24-70) without editing it.

    Only meaningful inside the reference application; returns the SourceManager class.
    """
    from core.source_manager import SourceManager          # type: ignore  (reference module)
    SourceManager.SOURCE_CLASSES[source_type] = B200SampleDataSource
    SourceManager.SOURCE_DISPLAY_NAMES[source_type] = display_name
    SourceManager._SAMPLE_SOURCES = frozenset(set(SourceManager._SAMPLE_SOURCES) | {source_type})
    SourceManager._SOURCE_LIMITS[source_type] = dict(SourceManager._SOURCE_LIMITS[like])
    SourceManager._SOURCE_DEFAULTS[source_type] = dict(SourceManager._SOURCE_DEFAULTS[like])
    return SourceManager
