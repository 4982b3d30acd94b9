"""Host-side wrapper over the C ABI: torch tensors are only the device-memory container.

``SpectrumPlan`` owns one ``tdsa_handle_t``; every method hands ``data_ptr()``s and the
current CUDA stream to libtdsa.so.  No arithmetic of the hot path happens in Python or
in torch ops.
"""
from __future__ import annotations

import ctypes as C
from typing import Optional

import numpy as np
import torch

from . import _lib as L

# utils/constants.py:152-155 of the reference
LOG_FLOOR = 1e-12
POWER_LOG_FLOOR = 1e-10


def default_floor(mode: str) -> float:
    return POWER_LOG_FLOOR if mode == "power" else LOG_FLOOR


def numpy_window(name: str, n: int, norm: str = "none") -> np.ndarray:
    """The float64 table the reference would build (rtl_samples.py:199-206, hackrf_samples.py:311-316)."""
    name = name.lower()
    fn = {"hanning": np.hanning, "hann": np.hanning, "hamming": np.hamming, "rectangle": np.ones,
          "rect": np.ones, "blackman": np.blackman}.get(name, np.hanning)
    w = fn(n)
    if norm == "rms":
        w32 = w.astype(np.float32)
        w32 /= np.sqrt(np.mean(w32 ** 2))
        return w32.astype(np.float64)
    return np.asarray(w, dtype=np.float64)


def _require_cuda() -> None:
    if not torch.cuda.is_available():
        raise L.TdsaError("topdogspectrumanalyser_b200 needs a CUDA device (B200); there is no CPU fallback")


def _stream_ptr() -> int:
    return torch.cuda.current_stream().cuda_stream


# words of the device flag block (include/tdsa.h: TDSA_FLAG_*)
FLAG_COUNT, FLAG_MAX_VALID, FLAG_MIN_VALID, FLAG_LIVE, FLAG_TARE_COLLECTING, FLAG_TARE_ACTIVE, FLAG_TARE_COUNT = range(7)
FLAG_WORDS = 8


class TraceState:
    """Device-resident trace state: TraceAverager buffer + max/min hold rows + the scalars that go with them.

    Mirrors ``TraceAverager._buffer/_count`` (utils/signal_processing.py:15-17),
    ``mw.max_power_levels / mw.min_power_levels`` (core/display_data_processor.py:371-395) and ``TareState``
    (core/tare_state.py).  The scalars (count, "hold initialised", tare collecting / active / count) live in
    ``flags``, an int32 block ON THE DEVICE, so no kernel call has to hand anything back to the host; the
    controls below are stream-ordered device writes, the read-outs (``count``, ``valid``, ``tare_active``)
    synchronise because the caller asked for a host value.
    """
    TARE_NUM_SAMPLES = 32                                 # utils/constants.py:141

    def __init__(self, width: int, device, avg_mode: str = "off", avg_n: int = 1,
                 max_hold_enabled: bool = False, min_hold_enabled: bool = False):
        self.width, self.device = int(width), torch.device(device)
        self.avg_mode, self.avg_n = avg_mode, int(avg_n)
        self.max_hold_enabled, self.min_hold_enabled = max_hold_enabled, min_hold_enabled
        self.avg = torch.zeros(self.width, dtype=torch.float64, device=self.device)
        self.max_hold = torch.zeros(self.width, dtype=torch.float32, device=self.device)
        self.min_hold = torch.zeros(self.width, dtype=torch.float32, device=self.device)
        self.flags = torch.zeros(FLAG_WORDS, dtype=torch.int32, device=self.device)
        self.last_row = torch.zeros(self.width, dtype=torch.float32, device=self.device)   # last good dB row
        # tare / normalisation (core/tare_state.py; display_data_processor.py:329-369)
        self.tare_buf = torch.zeros(self.width, dtype=torch.float64, device=self.device)
        self.tare_baseline = torch.zeros(self.width, dtype=torch.float64, device=self.device)

    # ---- controls: device writes, no synchronisation ------------------------------------------------
    def start_tare(self) -> None:
        """Begin collecting a baseline (TareState(collecting=True))."""
        self.flags[FLAG_TARE_COLLECTING:FLAG_TARE_COUNT + 1] = torch.tensor([1, 0, 0], dtype=torch.int32).to(
            self.device, non_blocking=True)

    def clear_tare(self) -> None:
        self.flags[FLAG_TARE_COLLECTING:FLAG_TARE_COUNT + 1].zero_()

    def set_averaging(self, mode: str, n: int) -> None:       # TraceAverager.set_mode, :19-28
        self.avg_mode, self.avg_n = mode, max(1, int(n))
        self.reset_averaging()

    def reset_averaging(self) -> None:                        # TraceAverager.reset, :30-33
        self.flags[FLAG_COUNT:FLAG_COUNT + 1].zero_()

    def clear_holds(self) -> None:
        self.flags[FLAG_MAX_VALID:FLAG_MIN_VALID + 1].zero_()

    def clear_max_hold(self) -> None:
        self.flags[FLAG_MAX_VALID:FLAG_MAX_VALID + 1].zero_()

    def clear_min_hold(self) -> None:
        self.flags[FLAG_MIN_VALID:FLAG_MIN_VALID + 1].zero_()

    # ---- read-outs: host values, synchronise -----------------------------------------------------------
    def host_flags(self) -> list:
        return self.flags.cpu().tolist()

    @property
    def count(self) -> int:
        return int(self.flags[FLAG_COUNT].item())

    @property
    def valid(self) -> tuple:
        f = self.host_flags()
        return (f[FLAG_MAX_VALID], f[FLAG_MIN_VALID])

    @property
    def live_frames(self) -> int:
        """Frames of the last call that were not silent / all-NaN."""
        return int(self.flags[FLAG_LIVE].item())

    @property
    def tare_active(self) -> bool:
        return bool(self.flags[FLAG_TARE_ACTIVE].item())

    @property
    def tare_collecting(self) -> bool:
        return bool(self.flags[FLAG_TARE_COLLECTING].item())

    @property
    def averaging(self) -> bool:                              # TraceAverager.is_active, :63-65
        return self.avg_mode != "off" and self.avg_n > 1


class SpectrumPlan:
    """One FFT size on one GPU. Thread-compatible, like the reference's data sources."""

    def __init__(self, n_fft: int, window: str = "hanning", window_norm: str = "none", mode: str = "power",
                 log_floor: Optional[float] = None, fs: float = 1.0, precision: str = "f64",
                 device: Optional[torch.device] = None):
        _require_cuda()
        self.lib = L.load()
        self.device = torch.device(device if device is not None else f"cuda:{torch.cuda.current_device()}")
        self.n_fft = int(n_fft)
        self.mode = mode
        self.fs = float(fs)
        self.log_floor = default_floor(mode) if log_floor is None else float(log_floor)
        self.precision = precision
        self.window_name, self.window_norm = window, window_norm
        self._h = C.c_void_p()
        with torch.cuda.device(self.device):
            L.check(self.lib.tdsa_create(self.n_fft, L.WINDOW_RECT, L.NORM_NONE, L.MODE_IDS[mode], self.log_floor,
                                         self.fs, L.PREC_IDS[precision], C.byref(self._h)))
        self.set_window(window, window_norm)

    # ---- configuration -----------------------------------------------------------------
    def close(self) -> None:
        if getattr(self, "_h", None) and self._h.value:
            self.lib.tdsa_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def set_window(self, window: str, norm: str = "none") -> None:
        """Install numpy's own table so the window is bit-identical to the reference's."""
        self.window_name, self.window_norm = window, norm
        w = np.ascontiguousarray(numpy_window(window, self.n_fft, norm))
        L.check(self.lib.tdsa_set_window_table_host(self._h, w.ctypes.data))

    def set_window_builtin(self, window: str, norm: str = "none") -> None:
        """Use the library's own C implementation of the window (no numpy on the path)."""
        L.check(self.lib.tdsa_set_window(self._h, L.WINDOW_IDS[window.lower()],
                                         L.NORM_RMS_F32 if norm == "rms" else L.NORM_NONE))

    def window_table(self) -> np.ndarray:
        w = np.empty(self.n_fft, dtype=np.float64)
        L.check(self.lib.tdsa_get_window_table_host(self._h, w.ctypes.data))
        return w

    def set_mode(self, mode: str, log_floor: Optional[float] = None, fs: Optional[float] = None) -> None:
        self.mode = mode
        self.log_floor = default_floor(mode) if log_floor is None else float(log_floor)
        if fs is not None:
            self.fs = float(fs)
        L.check(self.lib.tdsa_set_mode(self._h, L.MODE_IDS[mode], self.log_floor, self.fs))

    def set_precision(self, precision: str) -> None:
        self.precision = precision
        L.check(self.lib.tdsa_set_precision(self._h, L.PREC_IDS[precision]))

    def info(self) -> dict:
        v = [C.c_int32() for _ in range(5)]
        L.check(self.lib.tdsa_plan_info(self._h, *[C.byref(x) for x in v]))
        keys = ("n_fft", "threads_per_cta", "ctas_per_sm", "smem_bytes", "grid")
        return dict(zip(keys, (int(x.value) for x in v)))

    # ---- helpers -----------------------------------------------------------------------
    def _bind(self) -> None:
        L.check(self.lib.tdsa_set_stream(self._h, _stream_ptr()))

    def _frames(self, iq: torch.Tensor, n_frames, frame_stride):
        if iq.dtype != torch.complex64 or not iq.is_cuda or not iq.is_contiguous():
            raise ValueError("iq must be a contiguous complex64 CUDA tensor")
        if n_frames is None:
            if iq.dim() != 2 or iq.shape[1] != self.n_fft:
                raise ValueError(f"iq must be [B, {self.n_fft}] (or pass n_frames/frame_stride for a flat stream)")
            return int(iq.shape[0]), self.n_fft
        stride = self.n_fft if frame_stride is None else int(frame_stride)
        if (n_frames - 1) * stride + self.n_fft > iq.numel():
            raise ValueError("frames run past the end of iq")
        return int(n_frames), stride

    # ---- kernel 1 ----------------------------------------------------------------------
    def psd_db(self, iq: torch.Tensor, out: Optional[torch.Tensor] = None, n_frames: Optional[int] = None,
               frame_stride: Optional[int] = None) -> torch.Tensor:
        """dB rows ``float32 [B, N]`` of fftshift(FFT(iq*window)) — rtl_samples.py:169-184 per row."""
        b, stride = self._frames(iq, n_frames, frame_stride)
        if out is None:
            out = torch.empty((b, self.n_fft), dtype=torch.float32, device=iq.device)
        self._bind()
        L.check(self.lib.tdsa_psd_db_batch(self._h, iq.data_ptr(), b, stride, out.data_ptr()))
        return out

    def power_linear(self, iq: torch.Tensor, out: Optional[torch.Tensor] = None, n_frames: Optional[int] = None,
                     frame_stride: Optional[int] = None) -> torch.Tensor:
        b, stride = self._frames(iq, n_frames, frame_stride)
        if out is None:
            out = torch.empty((b, self.n_fft), dtype=torch.float64, device=iq.device)
        self._bind()
        L.check(self.lib.tdsa_power_linear_batch(self._h, iq.data_ptr(), b, stride, out.data_ptr()))
        return out

    def psd_db_dc(self, iq: torch.Tensor, dc_state: torch.Tensor, dc_alpha: float = 1.0,
                  out: Optional[torch.Tensor] = None):
        """HackRF-style front end (hackrf_samples.py:351-368). Returns ``(db, silent_flags)``."""
        b, stride = self._frames(iq, None, None)
        if out is None:
            out = torch.empty((b, self.n_fft), dtype=torch.float32, device=iq.device)
        silent = torch.empty(b, dtype=torch.int32, device=iq.device)
        self._bind()
        L.check(self.lib.tdsa_psd_db_batch_dc(self._h, iq.data_ptr(), b, stride, float(dc_alpha), dc_state.data_ptr(),
                                              silent.data_ptr(), out.data_ptr()))
        return out, silent

    def psd_db_avg_hold(self, iq: torch.Tensor, state: TraceState, last_only: bool = False,
                        out: Optional[torch.Tensor] = None) -> torch.Tensor:
        """Frames folded into the running trace state (averager + holds) on the device; nothing is read back."""
        return self._avg_hold(iq, state, None, 1.0, last_only, out)[0]

    def psd_db_avg_hold_dc(self, iq: torch.Tensor, state: TraceState, dc_state: torch.Tensor, dc_alpha: float = 1.0,
                           last_only: bool = False, out: Optional[torch.Tensor] = None):
        """HackRF front end + trace state (hackrf_samples.py:351-381). Returns ``(db, silent_flags)``."""
        return self._avg_hold(iq, state, dc_state, dc_alpha, last_only, out)

    def _avg_hold(self, iq, state: TraceState, dc_state, dc_alpha, last_only, out):
        b, stride = self._frames(iq, None, None)
        rows = 1 if last_only else b
        if out is None:
            out = torch.empty((rows, self.n_fft), dtype=torch.float32, device=iq.device)
        silent = torch.empty(b, dtype=torch.int32, device=iq.device) if dc_state is not None else None
        self._bind()
        L.check(self.lib.tdsa_psd_db_avg_hold_dev(
            self._h, iq.data_ptr(), b, stride, int(dc_state is not None), float(dc_alpha),
            dc_state.data_ptr() if dc_state is not None else None, silent.data_ptr() if silent is not None else None,
            L.AVG_IDS[state.avg_mode], state.avg_n, state.avg.data_ptr(),
            state.max_hold.data_ptr() if state.max_hold_enabled else None,
            state.min_hold.data_ptr() if state.min_hold_enabled else None, state.flags.data_ptr(),
            state.last_row.data_ptr(), int(last_only), out.data_ptr()))
        return out, silent

    def group_avg_db(self, iq: torch.Tensor) -> torch.Tensor:
        """``iq[G, F, N]`` -> ``float32 [G, N]``: linear mean over the F frames of each group, then dB (config 4)."""
        if iq.dtype != torch.complex64 or not iq.is_cuda or not iq.is_contiguous() or iq.dim() != 3 \
                or iq.shape[2] != self.n_fft:
            raise ValueError(f"iq must be a contiguous complex64 CUDA tensor [G, F, {self.n_fft}]")
        g, f = int(iq.shape[0]), int(iq.shape[1])
        out = torch.empty((g, self.n_fft), dtype=torch.float32, device=iq.device)
        self._bind()
        L.check(self.lib.tdsa_group_avg_db(self._h, iq.data_ptr(), g, f, out.data_ptr()))
        return out

    def welch(self, stream: torch.Tensor, hop: int):
        """Config 3: ``(avg_db, peak_db)`` over overlapping segments of a flat complex64 stream."""
        if stream.dtype != torch.complex64 or not stream.is_cuda or not stream.is_contiguous():
            raise ValueError("stream must be a contiguous complex64 CUDA tensor")
        avg = torch.empty(self.n_fft, dtype=torch.float32, device=stream.device)
        peak = torch.empty(self.n_fft, dtype=torch.float32, device=stream.device)
        self._bind()
        L.check(self.lib.tdsa_welch(self._h, stream.data_ptr(), stream.numel(), int(hop), avg.data_ptr(), peak.data_ptr()))
        return avg, peak

    # ---- host-buffer entry point (what the data source calls) ----------------------------
    def psd_db_host(self, iq_host: np.ndarray, out_host: Optional[np.ndarray] = None,
                    chunk_frames: int = 0) -> np.ndarray:
        """complex64 host frames -> float32 host dB rows; H2D, kernel and D2H are pipelined in the library."""
        if iq_host.dtype != np.complex64 or not iq_host.flags.c_contiguous or iq_host.ndim != 2 \
                or iq_host.shape[1] != self.n_fft:
            raise ValueError(f"iq_host must be C-contiguous complex64 [B, {self.n_fft}]")
        b = iq_host.shape[0]
        if out_host is None:
            out_host = np.empty((b, self.n_fft), dtype=np.float32)
        self._bind()
        with torch.cuda.device(self.device):
            L.check(self.lib.tdsa_psd_db_batch_host(self._h, iq_host.ctypes.data, b, self.n_fft,
                                                    out_host.ctypes.data, int(chunk_frames)))
        return out_host


# ---- plan-free operators -------------------------------------------------------------------
def trace_update(rows: torch.Tensor, state: TraceState, cal_offset_db: float = 0.0,
                 out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """Device-resident DataProcessor state on dB rows (display_data_processor.py:211-218,317-327,371-395)."""
    _require_cuda()
    lib = L.load()
    if rows.dtype != torch.float32 or not rows.is_cuda or not rows.is_contiguous() or rows.dim() != 2:
        raise ValueError("rows must be contiguous float32 CUDA [R, W]")
    r, w = rows.shape
    if out is None:
        out = torch.empty_like(rows)
    scratch = torch.empty(max(r, 1), dtype=torch.int32, device=rows.device)
    L.check(lib.tdsa_trace_update_dev(rows.data_ptr(), r, w, float(cal_offset_db), L.AVG_IDS[state.avg_mode], state.avg_n,
                                      state.avg.data_ptr(),
                                      state.max_hold.data_ptr() if state.max_hold_enabled else None,
                                      state.min_hold.data_ptr() if state.min_hold_enabled else None,
                                      state.flags.data_ptr(), out.data_ptr(), _stream_ptr(), scratch.data_ptr(),
                                      state.TARE_NUM_SAMPLES, state.tare_buf.data_ptr(), state.tare_baseline.data_ptr()))
    return out


def stitch(rows: torch.Tensor, row_lo_hz: torch.Tensor, row_hz: float, start_hz: float, stop_hz: float,
           m: int, g0: int = 0, count: Optional[int] = None) -> torch.Tensor:
    """hackrf_sweep stitch on the device (datasources/hackrf_sweep.py:150-166).

    ``g0`` / ``count`` select a slice of the m-point grid (sharded stitch across ranks); the slice is bit-identical
    to the same elements of the full grid."""
    _require_cuda()
    lib = L.load()
    if rows.dtype != torch.float32 or row_lo_hz.dtype != torch.float64:
        raise ValueError("rows float32 [R, K], row_lo_hz float64 [R]")
    rows, row_lo_hz = rows.contiguous(), row_lo_hz.contiguous()
    r, k = rows.shape
    count = int(m) - int(g0) if count is None else int(count)
    out = torch.empty(count, dtype=torch.float64, device=rows.device)
    order = torch.empty(r, dtype=torch.int32, device=rows.device)
    L.check(lib.tdsa_stitch_range(rows.data_ptr(), row_lo_hz.data_ptr(), float(row_hz), r, k, float(start_hz),
                                  float(stop_hz), int(m), int(g0), count, out.data_ptr(), _stream_ptr(), order.data_ptr()))
    return out


class WaterfallRing:
    """Device history ring with the widget's semantics (displays/waterfall.py:163-180, 330-336).

    ``dedupe=True`` applies the widget's duplicate filter (a row identical to the previous one is not added); the
    write pointer then lives on the device (``state``), because only the device knows how many rows were new.
    ``image()`` hands out the colour-mapped RGBA picture of the display view (core/export_manager.py:67-84)."""

    def __init__(self, h: int, w: int, fill: float, device, dedupe: bool = False):
        _require_cuda()
        self.lib = L.load()
        self.h, self.w, self.dedupe = int(h), int(w), bool(dedupe)
        self.buf = torch.full((2 * self.h, self.w), float(fill), dtype=torch.float32, device=device)
        self.ptr = C.c_int64(0)                                        # host pointer of the plain push
        self.state = torch.zeros(4, dtype=torch.int64, device=device)  # {ptr, has_last, rows added by the last push, -}
        self.last_row = torch.zeros(self.w, dtype=torch.float32, device=device)
        self._on_device = self.dedupe

    def push(self, rows: torch.Tensor) -> None:
        rows = rows.reshape(-1, self.w)
        if rows.dtype != torch.float32 or not rows.is_contiguous():
            raise ValueError("rows must be contiguous float32")
        if not self._on_device:
            L.check(self.lib.tdsa_ring_push(rows.data_ptr(), rows.shape[0], self.buf.data_ptr(), self.h, self.w,
                                            C.byref(self.ptr), _stream_ptr()))
            return
        r = rows.shape[0]
        slot = torch.empty(max(r, 1), dtype=torch.int64, device=rows.device)
        differs = torch.empty(max(r, 1), dtype=torch.int32, device=rows.device)
        L.check(self.lib.tdsa_ring_push_dev(rows.data_ptr(), r, self.buf.data_ptr(), self.h, self.w, self.state.data_ptr(),
                                            self.last_row.data_ptr(), int(self.dedupe), slot.data_ptr(), differs.data_ptr(),
                                            _stream_ptr()))

    def use_device_pointer(self) -> None:
        """Move the write pointer to the device (needed by ``image()`` and by dedupe); keeps the current position."""
        if not self._on_device:
            self.state[0] = int(self.ptr.value)
            self._on_device = True

    @property
    def rows_added(self) -> int:
        """Rows the last device-side push really added (duplicates excluded). Synchronises."""
        return int(self.state[2].item())

    def view(self) -> torch.Tensor:
        p = int(self.state[0].item()) if self._on_device else int(self.ptr.value)
        return self.buf[p:p + self.h]

    def image(self, lo_db: float, hi_db: float, lut_rgba: torch.Tensor, out: Optional[torch.Tensor] = None) -> torch.Tensor:
        """uint8 [H, W, 4]: the display view through the widget's LUT, newest row first; no host round trip."""
        if lut_rgba.dtype != torch.uint8 or tuple(lut_rgba.shape) != (256, 4) or not lut_rgba.is_cuda:
            raise ValueError("lut_rgba must be a uint8 CUDA tensor [256, 4]")
        self.use_device_pointer()
        if out is None:
            out = torch.empty((self.h, self.w, 4), dtype=torch.uint8, device=self.buf.device)
        L.check(self.lib.tdsa_ring_image_rgba(self.buf.data_ptr(), self.h, self.w, self.state.data_ptr(), float(lo_db),
                                              float(hi_db), lut_rgba.contiguous().data_ptr(), out.data_ptr(), _stream_ptr()))
        return out
