"""IQ feeds for :class:`B200SampleDataSource`.

A *feed* is anything with pyrtlsdr's call surface — ``read_samples(n)``, ``get_sample_rate()``,
``get_center_freq()``, optional ``close()`` and settable ``sample_rate`` / ``center_freq`` / ``gain`` — so an
``rtlsdr.RtlSdr`` object is a feed as it is.  This module adds:

* :class:`SyntheticIQFeed`, :class:`ReplayFeed` — seeded / recorded IQ for tests, demos and file replay;
* :class:`ChunkRingFeed` — the streaming hand-over between a producer thread and the GUI tick, with the
  *policy* of the reference HackRF source (datasources/hackrf_samples.py:28-30: 65 536-sample chunks, four
  pending at most, newest data wins, 0.5 s patience) on a fixed slot ring with sequence counters;
* :class:`HackrfDeviceFeed` — a :class:`ChunkRingFeed` whose producer is a reader thread on a pyhackrf device
  (what ``HackrfSamplesDataSource`` does in ``start``/``_reader_loop``, datasources/hackrf_samples.py:82-252);
* :func:`open_rtlsdr` — opens the RTL-SDR dongle the way ``RtlSamplesDataSource.start`` does
  (datasources/rtl_samples.py:42-46).
"""
from __future__ import annotations

import logging
import threading
import time
from typing import Optional

import numpy as np

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


class ChunkRingFeed:
    """Producer/consumer hand-over of IQ chunks: a ring of ``SLOTS`` chunk slots and two sequence counters.

    ``_head`` counts chunks ever offered, ``_tail`` is the sequence number of the oldest chunk still pending;
    slot ``seq % SLOTS`` holds chunk ``seq``.  The consumer keeps a *cursor* into the chunk it is currently
    eating: ``(_cur, _cur_end)`` = the array and how much of its front part is still unread.

    Policy (identical in effect to datasources/hackrf_samples.py:221-237 and :254-305, checked by replaying the
    executed reference in tests/test_host_logic.py):

    * ``put`` never blocks; when all slots are pending the oldest pending chunk is discarded and counted;
    * ``read_samples(n)`` first jumps to the newest pending chunk, if there is one, discarding everything older
      (including what was left of the current chunk); it then hands out the LAST ``n`` unread samples of the
      current chunk (the end of a transfer is the most recent data) and shortens the chunk from the back;
    * when fewer than ``n`` samples are left it waits — in 10 ms slices, ``timeout`` seconds at most — for the
      producer, again jumping to the newest chunk, and returns ``None`` if the time runs out.
    """
    READ_CHUNK = 65536          # samples per chunk the producers use (hackrf_samples.py:28)
    SLOTS = 4                   # pending chunks at most (hackrf_samples.py:29)
    PATIENCE_S = 0.5            # hackrf_samples.py:30
    POLL_S = 0.01

    _EMPTY = np.zeros(0, dtype=np.complex64)

    def __init__(self, sample_rate: float, centre_freq: float, timeout: Optional[float] = None):
        self.sample_rate, self.center_freq = float(sample_rate), float(centre_freq)
        self.gain = None
        self.timeout = self.PATIENCE_S if timeout is None else float(timeout)
        self._slots = [None] * self.SLOTS
        self._head = 0
        self._tail = 0
        self._cur = self._EMPTY
        self._cur_end = 0
        self._cv = threading.Condition()
        self.stats = {"samples_dropped": 0, "queue_overflows": 0}

    # pyrtlsdr surface
    def get_sample_rate(self):
        return self.sample_rate

    def get_center_freq(self):
        return self.center_freq

    def close(self):
        pass

    # ---- producer side -------------------------------------------------------------------------------
    def put(self, chunk: np.ndarray) -> None:
        with self._cv:
            if self._head - self._tail == self.SLOTS:          # every slot pending: the oldest one gives way
                victim = self._slots[self._tail % self.SLOTS]
                self.stats["samples_dropped"] += len(victim)
                self.stats["queue_overflows"] += 1
                self._tail += 1
            self._slots[self._head % self.SLOTS] = chunk
            self._head += 1
            self._cv.notify()

    @property
    def pending(self) -> int:
        with self._cv:
            return self._head - self._tail

    def flush(self) -> None:
        """Forget everything buffered (the reference does this around every re-tune, hackrf_samples.py:441-458)."""
        with self._cv:
            self._tail = self._head
            self._slots = [None] * self.SLOTS
            self._cur, self._cur_end = self._EMPTY, 0

    # ---- consumer side -------------------------------------------------------------------------------
    def _jump_to_newest(self) -> bool:
        """With the lock held: make the newest pending chunk current. False when nothing is pending."""
        if self._head == self._tail:
            return False
        newest = self._head - 1
        self._cur = self._slots[newest % self.SLOTS]
        self._cur_end = len(self._cur)
        for seq in range(self._tail, self._head):
            self._slots[seq % self.SLOTS] = None
        self._tail = self._head
        return True

    def _take_tail(self, n: int) -> np.ndarray:
        out = self._cur[self._cur_end - n:self._cur_end]
        self._cur_end -= n
        return out

    def read_samples(self, n: int):
        if n <= 0:
            return np.zeros(0, dtype=np.complex64)
        with self._cv:
            self._jump_to_newest()
            if self._cur_end >= n:
                return self._take_tail(n)
            deadline = time.monotonic() + self.timeout
            while self._cur_end < n:
                left = deadline - time.monotonic()
                if left <= 0:
                    return None
                if self._head == self._tail:
                    self._cv.wait(min(self.POLL_S, left))
                self._jump_to_newest()
            return self._take_tail(n)


class HackrfDeviceFeed(ChunkRingFeed):
    """:class:`ChunkRingFeed` fed by a reader thread on a HackRF (pyhackrf's ``HackRF`` object).

    Mirrors the device handling of ``HackrfSamplesDataSource``: configuration order of ``_setup_device``
    (hackrf_samples.py:108-135), a daemon reader that gives up after five consecutive read errors (:196,239-250)
    and a ``close`` that force-closes the device when the reader is stuck in a USB transfer (:147-160).
    ``device`` may be injected (tests); otherwise ``hackrf.HackRF()`` is opened.
    """
    MAX_CONSECUTIVE_ERRORS = 5
    JOIN_S = 2.0

    def __init__(self, sample_rate: float, centre_freq: float, lna_gain: int = 16, vga_gain: int = 20,
                 amplifier: bool = True, device=None, timeout: Optional[float] = None):
        super().__init__(sample_rate, centre_freq, timeout)
        self.lna_gain, self.vga_gain, self.amplifier = lna_gain, vga_gain, amplifier
        self.stats.update(read_errors=0, last_read_time=0.0)
        self._dev = device
        self._dev_lock = threading.Lock()
        self._halt = threading.Event()
        self._thread: Optional[threading.Thread] = None

    @property
    def thread(self):
        return self._thread

    def _program(self) -> None:
        d = self._dev
        d.set_sample_rate(int(self.sample_rate))
        d.set_freq(int(self.center_freq))
        d.set_lna_gain(self.lna_gain)
        d.set_vga_gain(self.vga_gain)
        (d.enable_amp if self.amplifier else d.disable_amp)()

    def open(self) -> None:
        if self._thread is not None:
            return
        with self._dev_lock:
            if self._dev is None:
                try:
                    from hackrf import HackRF                  # type: ignore
                except (ImportError, OSError) as e:
                    raise RuntimeError("HackRF library (libhackrf) not available on this system") from e
                self._dev = HackRF()
            try:
                self._program()
            except Exception:
                self._close_device()
                raise
        self.flush()
        self._halt.clear()
        self._thread = threading.Thread(target=self._pump, daemon=True, name="B200-HackRF-Reader")
        self._thread.start()

    def _pump(self) -> None:
        strikes = 0
        while not self._halt.is_set():
            try:
                with self._dev_lock:
                    if self._dev is None:
                        return
                    block = self._dev.read_samples(self.READ_CHUNK)
                if block is None or len(block) == 0:
                    continue
                strikes = 0
                self.stats["last_read_time"] = time.time()
                self.put(block)
            except Exception as e:                              # noqa: BLE001 - device libraries raise anything
                strikes += 1
                self.stats["read_errors"] += 1
                if strikes >= self.MAX_CONSECUTIVE_ERRORS:
                    logger.error("HackRF reader: %d consecutive read errors (%s); stopping", strikes, e)
                    self._halt.set()
                    return
                time.sleep(self.POLL_S)

    @property
    def alive(self) -> bool:
        return self._thread is not None and self._thread.is_alive()

    def _stop_pump(self, join_s: float) -> None:
        self._halt.set()
        t, self._thread = self._thread, None
        if t is not None and t.is_alive():
            t.join(join_s)
            if t.is_alive():                                    # stuck inside a USB transfer: pull the device away
                logger.warning("HackRF reader stuck; force-closing the device")
                self._close_device(locked=False)
                t.join(1.0)

    def _close_device(self, locked: bool = True) -> None:
        d, self._dev = self._dev, None
        if d is not None:
            try:
                d.close()
            except Exception as e:                              # noqa: BLE001
                logger.debug("error closing HackRF: %s", e)

    def retune(self, sample_rate: Optional[float] = None, centre_freq: Optional[float] = None) -> None:
        """Stop streaming, re-program the (kept) device, drop stale chunks, stream again (hackrf_samples.py:560-618)."""
        running = self._thread is not None
        if running:
            self._stop_pump(0.5)
        if sample_rate is not None:
            self.sample_rate = float(sample_rate)
        if centre_freq is not None:
            self.center_freq = float(centre_freq)
        if running:
            self.open()
        else:
            self.flush()

    def set_gains(self, lna_gain: Optional[int] = None, vga_gain: Optional[int] = None) -> None:
        if lna_gain is not None:
            if not 0 <= lna_gain <= 40:
                raise ValueError(f"LNA gain must be between 0 and 40, got {lna_gain}")
            self.lna_gain = lna_gain
        if vga_gain is not None:
            if not 0 <= vga_gain <= 62:
                raise ValueError(f"VGA gain must be between 0 and 62, got {vga_gain}")
            self.vga_gain = vga_gain
        with self._dev_lock:
            if self._dev is not None and self.alive:
                if lna_gain is not None:
                    self._dev.set_lna_gain(self.lna_gain)
                if vga_gain is not None:
                    self._dev.set_vga_gain(self.vga_gain)

    def set_amplifier(self, enabled: bool) -> None:
        self.amplifier = bool(enabled)
        with self._dev_lock:
            if self._dev is not None and self.alive:
                (self._dev.enable_amp if enabled else self._dev.disable_amp)()

    def close(self) -> None:
        self._stop_pump(self.JOIN_S)
        with self._dev_lock:
            self._close_device()
        self.flush()


def open_rtlsdr(sample_rate: float, centre_freq: float, gain="auto"):
    """Open the RTL-SDR dongle and program it in the order rtl_samples.py:42-46 does. The returned ``RtlSdr`` object
    is the feed."""
    try:
        from rtlsdr import RtlSdr                              # type: ignore
    except (ImportError, OSError) as e:
        raise RuntimeError("RTL-SDR library (librtlsdr) not available on this system") from e
    sdr = RtlSdr()
    sdr.sample_rate = sample_rate
    sdr.center_freq = centre_freq
    sdr.gain = gain
    return sdr
