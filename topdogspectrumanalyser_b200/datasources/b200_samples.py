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
from .feeds import ChunkRingFeed, HackrfDeviceFeed, ReplayFeed, SyntheticIQFeed, open_rtlsdr  # noqa: F401

HackrfChunkFeed = ChunkRingFeed          # round-1 name of the streaming feed

logger = logging.getLogger(__name__)


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
        self._owns_feed = False                  # True when start() opened the device itself
        self._gain = "auto"
        self.lna_gain, self.vga_gain, self.amplifier = 16, 20, True          # hackrf_samples.py:45-47
        self._flush_reads_remaining = 0
        self._h2d_done: Optional[torch.cuda.Event] = None
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

    def _open_own_feed(self):
        """No feed was injected: open the device this style stands for (what the reference's start() does)."""
        if self.style == "rtl":
            return open_rtlsdr(self.sample_rate, self.centre_freq, self._gain)            # rtl_samples.py:42-46
        feed = HackrfDeviceFeed(self.sample_rate, self.centre_freq, self.lna_gain, self.vga_gain, self.amplifier)
        feed.open()                                                                        # hackrf_samples.py:97-104
        return feed

    def start(self, frequency=None):
        """Contract of rtl_samples.py:30-58 / hackrf_samples.py:82-106: the span becomes the sample rate, the hardware's
        actual rate is read back, failures surface as RuntimeError (SourceManager shows them, source_manager.py:490-494)."""
        if frequency:
            self.centre_freq, self.sample_rate = int(frequency.centre), int(frequency.span)
        if self.running:
            return
        try:
            if self.sdr is None:
                self.sdr, self._owns_feed = self._open_own_feed(), True
            else:
                self._push_to_feed(sample_rate=self.sample_rate, center_freq=self.centre_freq)
                if isinstance(self.sdr, HackrfDeviceFeed):
                    self.sdr.open()
            self.sample_rate = self.last_sample_rate = self.sdr.get_sample_rate()
            self._ensure_plan()
        except Exception as e:
            self.running = False
            if self._owns_feed:
                self._drop_feed()
            logger.error("B200 source initialisation failed: %s", e)
            raise RuntimeError(f"B200 source initialisation failed: {e}") from e
        self.running = True

    def _push_to_feed(self, **settings) -> None:
        """Best-effort attribute pushes for feeds that are plain objects (replay, synthetic)."""
        for name, value in settings.items():
            try:
                setattr(self.sdr, name, value)
            except Exception:                                  # noqa: BLE001 - a read-only feed is fine
                pass

    def _drop_feed(self) -> None:
        feed, self.sdr, self._owns_feed = self.sdr, None, False
        closer = getattr(feed, "close", None)
        if closer is not None:
            try:
                closer()
            except Exception as e:                             # noqa: BLE001
                logger.error("Error closing feed: %s", e)

    @property
    def thread(self):
        """Reader thread, if the feed has one (SourceManager joins it, source_manager.py:361-369)."""
        return getattr(self.sdr, "thread", None)

    @thread.setter
    def thread(self, value):
        pass

    def pause(self):
        self.running = False                                   # device stays open (rtl_samples.py:60-63)

    def resume(self):
        self.running = self.sdr is not None or self.running    # nothing to resume without a device (:65-71)

    def stop(self):
        """Close the feed. A feed this object opened itself is forgotten, so the next start() opens a fresh one
        (rtl_samples.py:73-82); an injected feed stays attached for tests that restart the source."""
        self.running = False
        if self.sdr is None:
            return
        if self._owns_feed:
            self._drop_feed()
        else:
            closer = getattr(self.sdr, "close", None)
            if closer is not None:
                try:
                    closer()
                except Exception as e:                         # noqa: BLE001
                    logger.error("Error closing feed: %s", e)

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

    # ---- re-tuning: one path, three public spellings -------------------------------------------------------
    _PLL_SETTLE_S = 0.006        # reads discarded after a centre change cover this much signal (rtl_samples.py:99-101)

    def _retune(self, rate: Optional[int] = None, centre: Optional[int] = None) -> None:
        """Re-program the feed. ``rate`` / ``centre`` are None when that quantity is not being changed.

        rtl style (rtl_samples.py:85-146): the dongle is re-programmed in place; a new rate is read back from the
        hardware and the centre is written again afterwards because the tuner may shift with the rate (:125-129);
        a new centre schedules max(3, 6 ms worth) of discarded reads.
        hackrf style (hackrf_samples.py:460-548): streaming stops, the device is re-programmed, buffered chunks and
        the held frame are dropped, the DC estimate survives."""
        live = self.running and self.sdr is not None
        if self.style == "hackrf":
            if rate is not None:
                self.sample_rate = self.last_sample_rate = rate
            if centre is not None:
                self.centre_freq = centre
            self._last_good_power = None
            if live:
                if hasattr(self.sdr, "retune"):
                    self.sdr.retune(rate, centre)
                else:
                    self._push_to_feed(**({"sample_rate": rate} if rate is not None else {}),
                                       **({"center_freq": centre} if centre is not None else {}))
                    if hasattr(self.sdr, "flush"):
                        self.sdr.flush()
            return
        try:
            if rate is not None:
                if live:
                    self.sdr.sample_rate = rate
                    self.sample_rate = self.last_sample_rate = self.sdr.get_sample_rate()
                    self.sdr.center_freq = self.centre_freq if centre is None else centre
                    self.centre_freq = self.sdr.get_center_freq()
                    if centre is not None:
                        centre = None if int(self.centre_freq) == centre else centre
                else:
                    self.sample_rate = rate
            if centre is not None and live:
                self.centre_freq = centre
                self.sdr.center_freq = centre
                self._flush_reads_remaining = max(3, int(self._PLL_SETTLE_S * self.sample_rate / self.fft_size))
        except Exception as e:
            what = "sample rate" if rate is not None else "centre frequency"
            logger.error("Error updating %s: %s", what, e)
            raise RuntimeError(f"Error updating {what}: {e}") from e

    def update_centre_frequency(self, centre_freq: float):
        centre = int(centre_freq)
        if centre == self.centre_freq or (self.style == "rtl" and not self.running):
            return                                             # rtl ignores a re-tune while stopped (:87-89)
        self._retune(centre=centre)

    def update_sample_rate(self, sample_rate: float):
        rate = int(sample_rate)
        if rate != self.last_sample_rate:
            self._retune(rate=rate)

    def update_frequency(self, sample_rate: float, centre_freq: float):
        rate, centre = int(sample_rate), int(centre_freq)
        if self.style == "hackrf":                             # one stop/re-program/start for both (:520-548)
            rate_new, centre_new = rate != self.last_sample_rate, centre != self.centre_freq
            if rate_new or centre_new:
                self._retune(rate if rate_new else None, centre if centre_new else None)
            return
        self.update_sample_rate(rate)                          # rtl_samples.py:136-146: rate first, then centre
        self.update_centre_frequency(centre)

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
        if self._h2d_done is not None:
            self._h2d_done.synchronize()                       # the previous frame's async copy has left the staging buffer
        self._pin_in.numpy()[:] = samples                      # complex128 feeds narrow to complex64 here
        self._x_dev.copy_(self._pin_in, non_blocking=True)
        if self._h2d_done is None:
            self._h2d_done = torch.cuda.Event()
        self._h2d_done.record()
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


class B200RtlSamples(B200SampleDataSource):
    """What ``SourceManager._initialise_rtl_samples`` constructs when the B200 backend is installed:
    ``cls(sample_rate=span, centre_freq=centre)`` (core/source_manager.py:554-572). Opens the RTL-SDR itself."""

    def __init__(self, sample_rate: int, centre_freq: int, feed=None, **kw):
        super().__init__(sample_rate, centre_freq, feed=feed, style="rtl", **kw)


class B200HackrfSamples(B200SampleDataSource):
    """Replacement for ``HackrfSamplesDataSource`` in ``SourceManager._initialise_hackrf_samples``
    (core/source_manager.py:535-552), which sets ``lna_gain`` / ``vga_gain`` before ``start()``."""

    def __init__(self, sample_rate: int, centre_freq: int, feed=None, **kw):
        super().__init__(sample_rate, centre_freq, feed=feed, style="hackrf", **kw)

    def set_gains(self, lna_gain=None, vga_gain=None):         # hackrf_samples.py:624-649
        if lna_gain is not None:
            self.lna_gain = lna_gain
        if vga_gain is not None:
            self.vga_gain = vga_gain
        if hasattr(self.sdr, "set_gains"):
            self.sdr.set_gains(lna_gain, vga_gain)

    def set_amplifier(self, enabled: bool):                    # hackrf_samples.py:656-666
        self.amplifier = bool(enabled)
        if hasattr(self.sdr, "set_amplifier"):
            self.sdr.set_amplifier(enabled)

    def set_dc_alpha(self, alpha: float) -> None:              # hackrf_samples.py:651-654
        self._DC_ALPHA = max(0.0, min(1.0, float(alpha)))

    def get_stats(self) -> dict:                               # hackrf_samples.py:679-696 (the counters the feed keeps)
        stats = dict(getattr(self.sdr, "stats", {}))
        stats.update(is_running=self.running, num_samples=self.fft_size, sample_rate=self.sample_rate,
                     centre_freq=self.centre_freq, queue_size=getattr(self.sdr, "pending", 0),
                     queue_capacity=getattr(self.sdr, "SLOTS", 0))
        return stats


BACKEND_ENV = "TDSA_BACKEND"


def install_backend(force: bool = False, sweep: Optional[bool] = None):
    """Make the reference application use the B200 backend for its two IQ sample sources.

    ``SourceManager.set_source`` validates the id against ``SOURCE_CLASSES`` and then constructs through a fixed
    ``if/elif`` over the five built-in ids (core/source_manager.py:389-448), so a NEW id is never constructed; the
    backend therefore REPLACES the classes behind ``"rtl_samples"`` and ``"hackrf_samples"``.  The replacements are
    registered as virtual subclasses of the classes they stand in for, because ``SourceManager`` branches on
    ``isinstance(src, RtlSamplesDataSource / HackrfSamplesDataSource)`` (:319, :254-257).

    Active when ``TDSA_BACKEND=b200`` or ``force``; returns the ``SourceManager`` class, or None
    when the switch is off.  Only meaningful with the reference's root on ``sys.path``.
    """
    import os
    if not force and os.environ.get(BACKEND_ENV, "").lower() != "b200":
        return None
    from core.source_manager import SourceManager              # type: ignore  (reference module)
    from datasources.hackrf_samples import HackrfSamplesDataSource   # type: ignore
    from datasources.rtl_samples import RtlSamplesDataSource   # type: ignore
    RtlSamplesDataSource.register(B200RtlSamples)
    HackrfSamplesDataSource.register(B200HackrfSamples)
    if not IN_REFERENCE_APP:                                   # this package was imported before the reference's root was on sys.path
        from datasources.base import SampleDataSource as RefBase   # type: ignore
        RefBase.register(B200SampleDataSource)
    SourceManager.SOURCE_CLASSES["rtl_samples"] = B200RtlSamples
    SourceManager.SOURCE_CLASSES["hackrf_samples"] = B200HackrfSamples
    # optional: the HackRF sweep source computed from raw IQ instead of the external hackrf_sweep binary
    # (TDSA_BACKEND_SWEEP=b200 or sweep=True); _initialise_hackrf_sweep constructs cls(start, stop, bin_size=...)
    if sweep or (sweep is None and os.environ.get("TDSA_BACKEND_SWEEP", "").lower() == "b200"):
        from datasources.hackrf_sweep import HackRFSweepDataSource            # type: ignore
        from .b200_sweep import B200SweepDataSource
        HackRFSweepDataSource.register(B200SweepDataSource) if hasattr(HackRFSweepDataSource, "register") else None
        SourceManager.SOURCE_CLASSES["hackrf_sweep"] = B200SweepDataSource
    return SourceManager


def uninstall_backend():
    """Put the reference's own classes back (tests)."""
    from core.source_manager import SourceManager              # type: ignore
    from datasources.hackrf_samples import HackrfSamplesDataSource   # type: ignore
    from datasources.rtl_samples import RtlSamplesDataSource   # type: ignore
    from datasources.hackrf_sweep import HackRFSweepDataSource   # type: ignore
    SourceManager.SOURCE_CLASSES["rtl_samples"] = RtlSamplesDataSource
    SourceManager.SOURCE_CLASSES["hackrf_samples"] = HackrfSamplesDataSource
    SourceManager.SOURCE_CLASSES["hackrf_sweep"] = HackRFSweepDataSource
    return SourceManager
