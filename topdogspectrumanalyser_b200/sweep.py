"""Config 4 — wideband stitch: sub-band frame batches sharded across GPUs, one all-gather of dB rows.

The reference's sweep source spawns the external ``hackrf_sweep`` binary for the per-sub-band
FFTs (datasources/hackrf_sweep.py:58-74) and only parses and stitches its dB rows (:135-168).
Here the per-sub-band stage is kernel 1 applied per sub-band (window -> FFT -> |.|^2 -> linear
mean over F frames -> dB), and the stitch is ``tdsa_stitch`` with the reference's geometry:
bin centres ``arange(lo + bw/2, hi, bw)`` and ``np.interp`` onto ``linspace(start, stop, M)``.

Sharding: contiguous blocks of sub-bands per rank (300 over 8 ranks = 38/38/38/38/37/37/37/37),
padded to the largest block so one fixed-size ``all_gather`` moves every rank's rows.
"""
from __future__ import annotations

from typing import List, Optional, Tuple

import torch
import torch.distributed as dist


def shard_bands(n_bands: int, world: int, rank: int) -> range:
    """Contiguous block of sub-bands owned by ``rank`` (first ``n_bands % world`` ranks get one extra)."""
    base, extra = divmod(n_bands, world)
    lo = rank * base + min(rank, extra)
    return range(lo, lo + base + (1 if rank < extra else 0))


def max_shard(n_bands: int, world: int) -> int:
    return -(-n_bands // world)


def gather_rows(local_rows: torch.Tensor, n_bands: int, group=None) -> torch.Tensor:
    """All-gather the per-rank dB rows ``[n_local, W]`` into ``[n_bands, W]`` (rank order = band order).

    One collective: every rank contributes a block padded to ``max_shard`` rows.  Works with the
    NCCL backend on GPUs (NVLink / NVSwitch) and with gloo on CPU tensors (host-logic tests).
    """
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return local_rows[:n_bands]
    world = dist.get_world_size(group)
    pad = max_shard(n_bands, world)
    w = local_rows.shape[1]
    block = torch.full((pad, w), float("nan"), dtype=local_rows.dtype, device=local_rows.device)
    block[:local_rows.shape[0]] = local_rows
    out = torch.empty((world * pad, w), dtype=local_rows.dtype, device=local_rows.device)
    dist.all_gather_into_tensor(out, block, group=group)
    parts = [out[r * pad:r * pad + len(shard_bands(n_bands, world, r))] for r in range(world)]
    return torch.cat(parts, dim=0)


class WidebandSweep:
    """6 GHz span as ``n_bands`` x ``band_hz`` sub-bands, ``n_fft`` points each."""

    def __init__(self, n_bands: int = 300, band_hz: float = 20e6, n_fft: int = 8192, start_hz: float = 0.0,
                 precision: str = "f64", device: Optional[torch.device] = None):
        from .engine import SpectrumPlan
        self.n_bands, self.band_hz, self.n_fft = n_bands, float(band_hz), n_fft
        self.start_hz = float(start_hz)
        self.stop_hz = self.start_hz + n_bands * self.band_hz
        self.bin_hz = self.band_hz / n_fft
        # hackrf_sweep.py:35: num_bins = int((stop - start) / bin_size)
        self.m = int((self.stop_hz - self.start_hz) / self.bin_hz)
        self.plan = SpectrumPlan(n_fft, "hanning", mode="power", precision=precision, fs=band_hz, device=device)
        self.device = self.plan.device

    def band_lo_hz(self, bands) -> torch.Tensor:
        return torch.tensor([self.start_hz + b * self.band_hz for b in bands], dtype=torch.float64, device=self.device)

    def local_rows(self, iq_local: torch.Tensor) -> torch.Tensor:
        """``iq_local[n_local, F, N]`` -> dB rows ``[n_local, N]`` (kernel 1 + linear mean over F)."""
        return self.plan.group_avg_db(iq_local)

    def stitch(self, rows: torch.Tensor, arrival_order: Optional[List[int]] = None) -> torch.Tensor:
        """Stitch all sub-band rows onto the fixed grid (float64 ``[M]``), reference geometry."""
        from .engine import stitch
        bands = list(range(self.n_bands)) if arrival_order is None else arrival_order
        return stitch(rows, self.band_lo_hz(bands), self.band_hz, self.start_hz, self.stop_hz, self.m)

    def run(self, iq_local: torch.Tensor, group=None) -> Tuple[torch.Tensor, torch.Tensor]:
        """Rank-local IQ -> (all dB rows ``[n_bands, N]``, stitched grid ``[M]``) on every rank."""
        rows = gather_rows(self.local_rows(iq_local), self.n_bands, group)
        return rows, self.stitch(rows)
