"""Display-side analytics on device-resident traces.

Each operator names the reference code whose result it reproduces; all arithmetic runs in libtdsa.so.
"""
from __future__ import annotations

from typing import List, Optional, Tuple

import torch

from . import _lib as L
from .engine import _require_cuda, _stream_ptr

AMP_BINS, AMP_MIN, AMP_RNG = 512, -200.0, 300.0          # displays/density_display.py:12-14
DECAY_RATES = {"fast": 0.88, "medium": 0.96, "slow": 0.995, "off": 1.0}   # :15


def colormap_rgba(rows: torch.Tensor, lo_db: float, hi_db: float, lut_rgba: torch.Tensor) -> torch.Tensor:
    """dB rows -> RGBA bytes exactly as the waterfall export does (core/export_manager.py:72-79)."""
    _require_cuda()
    if rows.dtype != torch.float32 or not rows.is_cuda or not rows.is_contiguous():
        raise ValueError("rows must be a contiguous float32 CUDA tensor")
    if lut_rgba.dtype != torch.uint8 or tuple(lut_rgba.shape) != (256, 4) or not lut_rgba.is_cuda:
        raise ValueError("lut_rgba must be a uint8 CUDA tensor [256, 4]")
    out = torch.empty(tuple(rows.shape) + (4,), dtype=torch.uint8, device=rows.device)
    L.check(L.load().tdsa_colormap_rgba(rows.data_ptr(), rows.numel(), float(lo_db), float(hi_db),
                                        lut_rgba.contiguous().data_ptr(), out.data_ptr(), _stream_ptr()))
    return out


class DensityHistogram:
    """Persistence histogram with decay (displays/density_display.py:296-319), float32 [W, 512] on the device."""

    def __init__(self, width: int, device, decay: str = "medium"):
        _require_cuda()
        self.hist = torch.zeros((width, AMP_BINS), dtype=torch.float32, device=device)
        self.decay = DECAY_RATES.get(decay, 0.96)

    def update(self, live_db: torch.Tensor) -> None:
        if live_db.dtype != torch.float32 or live_db.numel() != self.hist.shape[0]:
            raise ValueError("live_db must be float32 with one value per frequency bin")
        L.check(L.load().tdsa_density_update(live_db.contiguous().data_ptr(), live_db.numel(), float(self.decay),
                                             self.hist.data_ptr(), _stream_ptr()))


def band_power(freq_bins: torch.Tensor, levels_db: torch.Tensor, f_start: float, f_stop: float) -> Optional[float]:
    """Integrated power between two frequencies (core/marker_manager.py:308-318); None if no bin falls inside."""
    _require_cuda()
    if freq_bins.dtype != torch.float64 or levels_db.dtype != torch.float32:
        raise ValueError("freq_bins float64, levels_db float32")
    out = torch.empty(1, dtype=torch.float64, device=levels_db.device)
    L.check(L.load().tdsa_band_power(freq_bins.contiguous().data_ptr(), levels_db.contiguous().data_ptr(),
                                     levels_db.numel(), float(f_start), float(f_stop), out.data_ptr(), _stream_ptr()))
    v = float(out.item())
    return None if v != v else v


def top_peaks(freq_bins, power_db: torch.Tensor, n: int = 5, min_sep_bins: int = 10,
              min_excursion_db: float = 10.0) -> List[Tuple[float, float]]:
    """Up to n (freq, power) tuples, strongest first (core/display_data_processor.py:432-471)."""
    _require_cuda()
    if power_db.dtype != torch.float32 or not power_db.is_cuda:
        raise ValueError("power_db must be a float32 CUDA tensor")
    dev = power_db.device
    idx = torch.empty(16, dtype=torch.int32, device=dev)
    pwr = torch.empty(16, dtype=torch.float32, device=dev)
    cnt = torch.zeros(1, dtype=torch.int32, device=dev)
    L.check(L.load().tdsa_top_peaks(power_db.contiguous().data_ptr(), power_db.numel(), int(n), int(min_sep_bins),
                                    float(min_excursion_db), idx.data_ptr(), pwr.data_ptr(), cnt.data_ptr(), _stream_ptr()))
    k = int(cnt.item())
    ii, pp = idx[:k].cpu().tolist(), pwr[:k].cpu().tolist()
    return [(float(freq_bins[i]), float(p)) for i, p in zip(ii, pp)]


def snap_to_peak(freq_bins, levels_db: torch.Tensor, threshold_db: float = -200.0, excursion_db: float = 6.0,
                 distance: int = 3) -> Tuple[float, int, bool]:
    """Marker snap (core/marker_manager.py:74-99): ``scipy.signal.find_peaks(levels, height=threshold,
    prominence=excursion, distance=3)``, then the highest surviving peak; ``argmax(levels)`` when there is none.
    Returns ``(frequency, bin index, used_fallback)``."""
    _require_cuda()
    if levels_db.dtype != torch.float32 or not levels_db.is_cuda:
        raise ValueError("levels_db must be a float32 CUDA tensor")
    out = torch.empty(3, dtype=torch.int32, device=levels_db.device)
    L.check(L.load().tdsa_find_peaks_snap(levels_db.contiguous().data_ptr(), levels_db.numel(), float(threshold_db),
                                          float(excursion_db), int(distance), out.data_ptr(), _stream_ptr()))
    idx, _, fb = out.cpu().tolist()
    return float(freq_bins[idx]), int(idx), bool(fb)
