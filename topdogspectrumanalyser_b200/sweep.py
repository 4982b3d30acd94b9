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


def shard_grid(m: int, world: int, rank: int) -> Tuple[int, int]:
    """(first grid point, count) of the slice of the stitched grid that ``rank`` interpolates."""
    base, extra = divmod(m, world)
    g0 = rank * base + min(rank, extra)
    return g0, base + (1 if rank < extra else 0)


class WidebandSweep:
    """6 GHz span as ``n_bands`` x ``band_hz`` sub-bands, ``n_fft`` points each.

    Multi-GPU (one process per GPU, ``torch.distributed`` initialised): sub-bands are sharded by rank.
    ``exchange="peer"`` (default when symmetric memory is available) fuses the per-band mean with the gather: the FFT
    kernel's epilogue stores every finished dB row straight into ALL ranks' row tables over NVLink
    (``tdsa_group_avg_db_peers``), followed by one cross-rank barrier; ``exchange="nccl"`` computes local rows and
    all-gathers them.  ``grid="sharded"`` lets each rank interpolate only its slice of the stitched grid
    (``shard_grid``); ``"replicated"`` computes the whole grid on every rank like round 1.
    """

    def __init__(self, n_bands: int = 300, band_hz: float = 20e6, n_fft: int = 8192, start_hz: float = 0.0,
                 precision: str = "f64", device: Optional[torch.device] = None, exchange: str = "auto",
                 grid: str = "replicated", group=None):
        from .engine import SpectrumPlan
        self.n_bands, self.band_hz, self.n_fft = n_bands, float(band_hz), n_fft
        self.start_hz = float(start_hz)
        self.stop_hz = self.start_hz + n_bands * self.band_hz
        self.bin_hz = self.band_hz / n_fft
        # hackrf_sweep.py:35: num_bins = int((stop - start) / bin_size)
        self.m = int((self.stop_hz - self.start_hz) / self.bin_hz)
        self.plan = SpectrumPlan(n_fft, "hanning", mode="power", precision=precision, fs=band_hz, device=device)
        self.device = self.plan.device
        self.group, self.grid_mode = group, grid
        self.world = dist.get_world_size(group) if dist.is_available() and dist.is_initialized() else 1
        self.rank = dist.get_rank(group) if self.world > 1 else 0
        self._lo_all = self.band_lo_hz(range(n_bands))
        self._symm = None
        self.exchange = "none" if self.world == 1 else exchange
        if self.world > 1 and exchange in ("auto", "peer"):
            try:
                self._setup_peer_rows()
                self.exchange = "peer"
            except Exception as e:                         # noqa: BLE001 - no symmetric memory on this system
                if exchange == "peer":
                    raise
                self.exchange, self.peer_error = "nccl", repr(e)

    def _setup_peer_rows(self) -> None:
        import ctypes as C
        import torch.distributed._symmetric_memory as symm_mem
        if self.n_fft not in (4096, 8192) or self.world > 8:
            raise RuntimeError("peer exchange needs N = 4096 / 8192 and at most 8 ranks")
        self.rows_all = symm_mem.empty((self.n_bands, self.n_fft), dtype=torch.float32, device=self.device)
        self.rows_all.fill_(float("nan"))
        self._symm = symm_mem.rendezvous(self.rows_all, self.group if self.group is not None else dist.group.WORLD)
        self._peer_ptrs = (C.c_uint64 * self.world)(*[int(p) for p in self._symm.buffer_ptrs])
        self._symm.barrier()

    def band_lo_hz(self, bands) -> torch.Tensor:
        return torch.tensor([self.start_hz + b * self.band_hz for b in bands], dtype=torch.float64, device=self.device)

    def local_rows(self, iq_local: torch.Tensor) -> torch.Tensor:
        """``iq_local[n_local, F, N]`` -> dB rows ``[n_local, N]`` (kernel 1 + linear mean over F)."""
        return self.plan.group_avg_db(iq_local)

    def all_rows(self, iq_local: torch.Tensor) -> torch.Tensor:
        """Rank-local IQ -> every sub-band's dB row ``[n_bands, N]`` on every rank."""
        if self.exchange == "peer":
            from . import _lib as L
            mine = shard_bands(self.n_bands, self.world, self.rank)
            n_local, frames = int(iq_local.shape[0]), int(iq_local.shape[1])
            self.plan._bind()
            if n_local:
                L.check(self.plan.lib.tdsa_group_avg_db_peers(self.plan._h, iq_local.data_ptr(), n_local, frames, mine.start,
                                                              self._peer_ptrs, self.world))
            self._symm.barrier()                           # every rank's rows have landed in every table
            return self.rows_all
        return gather_rows(self.local_rows(iq_local), self.n_bands, self.group)

    def stitch(self, rows: torch.Tensor, arrival_order: Optional[List[int]] = None, sharded: Optional[bool] = None) -> torch.Tensor:
        """Stitch all sub-band rows onto the fixed grid (float64), reference geometry; ``sharded``: this rank's slice."""
        from .engine import stitch
        lo = self._lo_all if arrival_order is None else self.band_lo_hz(arrival_order)
        sharded = (self.grid_mode == "sharded" and self.world > 1) if sharded is None else sharded
        g0, count = shard_grid(self.m, self.world, self.rank) if sharded else (0, self.m)
        return stitch(rows, lo, self.band_hz, self.start_hz, self.stop_hz, self.m, g0, count)

    def run(self, iq_local: torch.Tensor, group=None) -> Tuple[torch.Tensor, torch.Tensor]:
        """Rank-local IQ -> (all dB rows ``[n_bands, N]``, stitched grid: whole ``[M]`` or this rank's slice)."""
        rows = self.all_rows(iq_local)
        return rows, self.stitch(rows)
