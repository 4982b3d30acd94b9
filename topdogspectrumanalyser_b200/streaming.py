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
    """Pinned chunk slots -> H2D on a side stream -> (per chunk) running average + dB rows + ring push on the main stream.

    ``use_graphs``: the device work of a chunk (FFT kernel, frame-ordered scan, flag update, ring scatter: four launches
    with fixed arguments per slot, all state on the device) is captured once per slot into a CUDA graph and replayed, so
    a chunk costs one graph launch instead of ~10 library / torch calls.  ``timeline()`` returns CUDA-event time stamps
    of the last chunks' copies and kernels: the evidence that copy i+1 runs under compute i.
    """

    def __init__(self, n_fft: int = 4096, chunk_samples: int = 65536, history: int = 1024, avg_mode: str = "exp",
                 avg_n: int = 8, precision: str = "f64", fill_db: float = -100.0, depth: int = 4,
                 device: Optional[torch.device] = None, use_graphs: bool = True, dedupe: bool = False):
        if chunk_samples % n_fft:
            raise ValueError("chunk_samples must be a multiple of n_fft")
        self.plan = SpectrumPlan(n_fft, "hanning", mode="power", precision=precision, device=device)
        self.device = self.plan.device
        self.n_fft, self.chunk, self.frames = n_fft, chunk_samples, chunk_samples // n_fft
        self.depth = depth                                  # like MAX_QUEUE_SIZE = 4 (hackrf_samples.py:29)
        self.state = TraceState(n_fft, self.device)
        self.state.set_averaging(avg_mode, avg_n)
        self.ring = WaterfallRing(history, n_fft, fill_db, self.device, dedupe=dedupe)
        self.ring.use_device_pointer()                      # fixed kernel arguments: the write pointer lives on the device
        self.pinned = [torch.empty(chunk_samples, dtype=torch.complex64).pin_memory() for _ in range(depth)]
        self.dev = [torch.empty((self.frames, n_fft), dtype=torch.complex64, device=self.device) for _ in range(depth)]
        self.rows = [torch.empty((self.frames, n_fft), dtype=torch.float32, device=self.device) for _ in range(depth)]
        self._slot_scratch = [torch.empty(self.frames, dtype=torch.int64, device=self.device) for _ in range(depth)]
        self._diff_scratch = [torch.empty(self.frames, dtype=torch.int32, device=self.device) for _ in range(depth)]
        self.side = torch.cuda.Stream(device=self.device)
        self.main = torch.cuda.Stream(device=self.device)   # graphs cannot be captured on the legacy default stream
        self.h2d_done = [torch.cuda.Event() for _ in range(depth)]
        self.slot_free = [torch.cuda.Event() for _ in range(depth)]
        self.lib = L.load()
        self.chunks_in = 0
        self.use_graphs = use_graphs
        self._graphs = [None] * depth
        self._trace = None                                  # event time stamps when timeline recording is on

    def _device_work(self, i: int) -> None:
        """Everything the device does for the chunk in slot i (current stream = self.main)."""
        self.plan.psd_db_avg_hold(self.dev[i], self.state, last_only=False, out=self.rows[i])
        r = self.ring
        L.check(self.lib.tdsa_ring_push_dev(self.rows[i].data_ptr(), self.frames, r.buf.data_ptr(), r.h, r.w,
                                            r.state.data_ptr(), r.last_row.data_ptr(), int(r.dedupe),
                                            self._slot_scratch[i].data_ptr(), self._diff_scratch[i].data_ptr(),
                                            self.main.cuda_stream))

    def acquire(self) -> np.ndarray:
        """The pinned host buffer of the next slot, for a producer that writes its samples in place (no extra host
        copy); blocks until the copy that last read this slot has finished.  Follow with :meth:`commit`."""
        i = self.chunks_in % self.depth
        if self.chunks_in >= self.depth:
            self.h2d_done[i].synchronize()                  # the previous copy out of this pinned buffer is complete
        return self.pinned[i].numpy()

    def push_chunk(self, samples: np.ndarray) -> None:
        """Queue one chunk: host copy into the pinned slot, async H2D on the side stream, compute on the main one."""
        self.acquire()[:] = samples
        self.commit()

    def commit(self) -> None:
        """Send the slot filled through :meth:`acquire` on its way: async H2D on the side stream, compute on the main one."""
        i = self.chunks_in % self.depth
        if self.chunks_in >= self.depth:
            self.slot_free[i].synchronize()                 # the kernels that read this slot's device buffer have finished
        tr = self._trace
        if tr is not None:
            e = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
            e[0].record(self.side)
        L.check(self.lib.tdsa_h2d_async(self.pinned[i].data_ptr(), self.dev[i].data_ptr(), self.chunk * 8,
                                        self.side.cuda_stream, None))
        self.h2d_done[i].record(self.side)
        if tr is not None:
            e[1].record(self.side)
        self.main.wait_event(self.h2d_done[i])
        with torch.cuda.stream(self.main):
            if tr is not None:
                e[2].record(self.main)
            if not self.use_graphs:
                self._device_work(i)
            elif self._graphs[i] is not None:
                self._graphs[i].replay()
            elif self.chunks_in < self.depth:
                self._device_work(i)                        # first pass over the slots: eager (allocations, tensor maps)
            else:
                g = torch.cuda.CUDAGraph()
                with torch.cuda.graph(g, stream=self.main):
                    self._device_work(i)
                self._graphs[i] = g
                g.replay()                                  # capture does not execute
            if tr is not None:
                e[3].record(self.main)
                tr.append(e)
            self.slot_free[i].record(self.main)
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

    def record_timeline(self, on: bool = True) -> None:
        self._trace = [] if on else None

    def timeline(self) -> list:
        """[{chunk, h2d: [start_us, end_us], compute: [start_us, end_us]}] relative to the first recorded copy."""
        torch.cuda.synchronize(self.device)
        tr = self._trace or []
        if not tr:
            return []
        base = tr[0][0]
        return [{"chunk": k, "h2d_us": [round(base.elapsed_time(e[0]) * 1e3, 1), round(base.elapsed_time(e[1]) * 1e3, 1)],
                 "compute_us": [round(base.elapsed_time(e[2]) * 1e3, 1), round(base.elapsed_time(e[3]) * 1e3, 1)]}
                for k, e in enumerate(tr)]

    def history(self) -> torch.Tensor:
        torch.cuda.current_stream(self.device).wait_stream(self.main)
        return self.ring.view()
