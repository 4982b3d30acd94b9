"""Config 5 — streaming waterfall: continuous IQ in 65 536-sample chunks, H2D overlapped on a side stream.

Shape of the reference's HackRF ingest (datasources/hackrf_samples.py:28-29,191-305): a reader
hands over 64 Ki-sample chunks; the reference then analyses ONE frame per 20 ms tick and drops the
rest.  Here every sample is analysed: each chunk (16 frames at N = 4096) is copied from pinned
host memory with ``tdsa_h2d_async`` on a side stream while the previous chunk is in the fused
kernel; frames are folded into the running average (TraceAverager 'exp', n = 8) and the dB rows
are pushed into a device ring with the waterfall widget's semantics (displays/waterfall.py:163-180).
"""
from __future__ import annotations

import ctypes as C
import time
from typing import Callable, Optional

import numpy as np
import torch

from . import _lib as L
from .engine import SpectrumPlan, TraceState, WaterfallRing


class WaterfallStreamer:
    def __init__(self, n_fft: int = 4096, chunk_samples: int = 65536, history: int = 1024, avg_mode: str = "exp",
                 avg_n: int = 8, precision: str = "f64", fill_db: float = -100.0, depth: int = 4,
                 device: Optional[torch.device] = None):
        if chunk_samples % n_fft:
            raise ValueError("chunk_samples must be a multiple of n_fft")
        self.plan = SpectrumPlan(n_fft, "hanning", mode="power", precision=precision, device=device)
        self.device = self.plan.device
        self.n_fft, self.chunk, self.frames = n_fft, chunk_samples, chunk_samples // n_fft
        self.depth = depth                                  # like MAX_QUEUE_SIZE = 4 (hackrf_samples.py:29)
        self.state = TraceState(n_fft, self.device)
        self.state.set_averaging(avg_mode, avg_n)
        self.ring = WaterfallRing(history, n_fft, fill_db, self.device)
        self.pinned = [torch.empty(chunk_samples, dtype=torch.complex64).pin_memory() for _ in range(depth)]
        self.dev = [torch.empty((self.frames, n_fft), dtype=torch.complex64, device=self.device) for _ in range(depth)]
        self.rows = [torch.empty((self.frames, n_fft), dtype=torch.float32, device=self.device) for _ in range(depth)]
        self.side = torch.cuda.Stream(device=self.device)
        self.h2d_done = [torch.cuda.Event() for _ in range(depth)]
        self.slot_free = [torch.cuda.Event() for _ in range(depth)]
        self.lib = L.load()
        self.chunks_in = 0

    def push_chunk(self, samples: np.ndarray) -> None:
        """Queue one chunk: host copy into the pinned slot, async H2D on the side stream, compute on the main one."""
        i = self.chunks_in % self.depth
        if self.chunks_in >= self.depth:
            self.slot_free[i].synchronize()                 # the kernel that read this slot has finished
        self.pinned[i].numpy()[:] = samples
        L.check(self.lib.tdsa_h2d_async(self.pinned[i].data_ptr(), self.dev[i].data_ptr(), self.chunk * 8,
                                        self.side.cuda_stream, None))
        self.h2d_done[i].record(self.side)
        main = torch.cuda.current_stream(self.device)
        main.wait_event(self.h2d_done[i])
        self.plan.psd_db_avg_hold(self.dev[i], self.state, last_only=False, out=self.rows[i])
        self.ring.push(self.rows[i])
        self.slot_free[i].record(main)
        self.chunks_in += 1

    def run(self, source: Callable[[int], np.ndarray], n_chunks: int, sample_rate: float = 20e6) -> dict:
        """Drive ``n_chunks`` chunks from ``source(index)``; returns throughput and the real-time factor."""
        torch.cuda.synchronize(self.device)
        t0 = time.perf_counter()
        for c in range(n_chunks):
            self.push_chunk(source(c))
        torch.cuda.synchronize(self.device)
        dt = time.perf_counter() - t0
        samples = n_chunks * self.chunk
        return {"samples": samples, "seconds": dt, "samples_per_s": samples / dt,
                "real_time_factor": samples / dt / sample_rate, "frames": n_chunks * self.frames}

    def history(self) -> torch.Tensor:
        return self.ring.view()
