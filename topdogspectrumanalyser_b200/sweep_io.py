"""hackrf_sweep wire formats -> row arrays -> device stitch.

``SweepAssembler`` mirrors the state machine of ``HackRFSweepDataSource._parse``
(datasources/hackrf_sweep.py:135-168): rows accumulate until a row whose ``hz_low`` is within 1 MHz of the
sweep start arrives while data is pending; the pending rows are then stitched onto the fixed grid
(``_create_frequency_grid`` :32-40) — here by ``tdsa_stitch`` on the GPU instead of list.extend + argsort +
np.interp.  Parsing itself is host work and lives in libtdsa.so (``tdsa_parse_sweep_csv_host`` /
``tdsa_parse_sweep_binary_host``); no GPU is needed for it.
"""
from __future__ import annotations

import ctypes as C
from typing import Optional, Tuple

import numpy as np

from . import _lib as L


def _parse(fn, data: bytes, max_rows: int, max_bins: int):
    lo = np.empty(max_rows, dtype=np.float64)
    hi = np.empty(max_rows, dtype=np.float64)
    vals = np.empty((max_rows, max_bins), dtype=np.float32)
    nb = np.empty(max_rows, dtype=np.int32)
    n, used = C.c_int64(0), C.c_int64(0)
    L.check(fn(data, len(data), max_rows, max_bins, lo.ctypes.data, hi.ctypes.data, vals.ctypes.data, nb.ctypes.data,
               C.byref(n), C.byref(used)))
    k = int(n.value)
    return lo[:k], hi[:k], vals[:k], nb[:k], int(used.value)


def parse_csv(text: bytes, max_rows: int = 4096, max_bins: int = 256):
    """CSV lines of ``hackrf_sweep`` -> (lo_hz, hi_hz, values[rows, max_bins], n_bins, bytes_consumed)."""
    return _parse(L.load().tdsa_parse_sweep_csv_host, text, max_rows, max_bins)


def parse_binary(buf: bytes, max_rows: int = 4096, max_bins: int = 256):
    """``hackrf_sweep -B`` records -> same tuple as :func:`parse_csv`."""
    return _parse(L.load().tdsa_parse_sweep_binary_host, buf, max_rows, max_bins)


class SweepAssembler:
    def __init__(self, start_freq: float, stop_freq: float, bin_size: int, device=None, binary: bool = False):
        self.start_freq, self.stop_freq, self.bin_size = int(start_freq), int(stop_freq), int(bin_size)
        self.num_bins = int((self.stop_freq - self.start_freq) / self.bin_size)        # hackrf_sweep.py:35
        self.binary = binary
        self.device = device
        self._tail = b""
        self._rows, self._lo, self._hi = [], [], []
        self.full_power_array: Optional[np.ndarray] = None
        self.full_power_dev = None

    def _stitch(self):
        import torch
        from .engine import stitch
        rows = np.stack(self._rows)
        lo = np.asarray(self._lo, dtype=np.float64)
        width = float(self._hi[0] - self._lo[0])
        dev = self.device or torch.device("cuda", torch.cuda.current_device())
        grid = stitch(torch.from_numpy(rows).to(dev), torch.from_numpy(lo).to(dev), width, float(self.start_freq),
                      float(self.stop_freq), self.num_bins)
        self.full_power_dev = grid
        self.full_power_array = grid.cpu().numpy()
        self._rows, self._lo, self._hi = [], [], []

    def feed(self, data: bytes) -> int:
        """Consume raw stdout bytes of hackrf_sweep; returns how many completed sweeps were stitched."""
        data = self._tail + data
        lo, hi, vals, nb, used = (parse_binary if self.binary else parse_csv)(data)
        self._tail = data[used:]
        done = 0
        for r in range(len(lo)):
            at_start = abs(lo[r] - self.start_freq) < 1e6                               # :148
            if at_start and self._rows:                                                 # :150
                self._stitch()
                done += 1
            self._rows.append(vals[r, :nb[r]].copy())
            self._lo.append(lo[r])
            self._hi.append(hi[r])
        return done

    def get_data(self) -> np.ndarray:
        """Last complete sweep, NaN before the first one (hackrf_sweep.py:39,224-229)."""
        if self.full_power_array is None:
            return np.full(self.num_bins, np.nan)
        return self.full_power_array.copy()
