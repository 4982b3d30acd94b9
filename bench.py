#!/usr/bin/env python
"""bench.py — headline metric of BASELINE.json on N GPUs of one node.

metric : complex64 IQ samples/s through the fused window+FFT+PSD+dB path at N=4096
workload: config 2 — 8192 frames x 4096 points per GPU, Hann window, power dB, synthetic IQ
          (AWGN + three tones, seed 1).  One "step" = one pass of the hot path over the batch.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--precision f64|f32] [--impl reference]

N > 1 is launched by torchrun (one rank per GPU); batches are sharded by rank with no
data-path collective ("weak" scaling: every rank owns a full 8192-frame batch).
Prints ONE JSON line on rank 0.
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

N_FFT = 4096
BATCH = 8192
BYTES_PER_SAMPLE = 12          # 8 B complex64 read + 4 B float32 dB written
METRIC = "complex64 IQ samples/s through fused window+FFT+PSD+dB at N=4096"
UNIT = "samples/s"


def workload_name(n_gpus):
    return (f"cfg2: 4096-pt FFT, batch {BATCH} frames per GPU x {n_gpus} GPU(s), Hann window, power dB, "
            "synthetic complex64 IQ (AWGN + 3 tones, seed 1)")


# ------------------------------------------------------------------------------------------
# reference arm / cpu_baseline: the reference's own per-frame numpy/scipy chain on host cores
# ------------------------------------------------------------------------------------------
_W_IQ = None
_W_OUT = None
_W_WIN = None


FS_HZ, FC_HZ = 2.048e6, 98e6       # RTL defaults (core/source_manager.py:67): only the freq-bin rebuild uses them


def _cpu_worker(span):
    """What one RtlSamplesDataSource.get_power_levels call does per frame (datasources/rtl_samples.py:167-188):
    the raw-sample copy (:168), window * FFT * shift * |.|^2 * dB (:169-184) and the frequency-bin rebuild (:188)."""
    from oracle import oracle as O
    lo, hi = span
    for f in range(lo, hi):
        raw = _W_IQ[f].copy()                                              # :168  self._store_raw(samples.copy())
        _W_OUT[f] = O.power_db_frame(raw, _W_WIN, O.MODE_POWER)            # :169-184
        O.freq_bins(N_FFT, FS_HZ, FC_HZ)                                   # :188  rebuilt on every call
    return hi - lo


class CpuReference:
    """All host cores, each running the reference's per-frame chain on a contiguous block of frames."""

    def __init__(self, sample_frames: int, cores: int):
        import multiprocessing as mp
        from oracle import oracle as O
        from topdogspectrumanalyser_b200 import synth
        global _W_IQ, _W_OUT, _W_WIN
        self.frames, self.cores = sample_frames, cores
        _W_IQ = synth.cfg2_frames(b=sample_frames, n=N_FFT, seed=1)
        _W_WIN = O.make_window("hanning", N_FFT)
        raw = mp.RawArray("d", sample_frames * N_FFT)
        _W_OUT = np.frombuffer(raw, dtype=np.float64).reshape(sample_frames, N_FFT)
        self.out = _W_OUT
        edges = np.linspace(0, sample_frames, cores + 1).astype(int)
        self.spans = [(int(a), int(b)) for a, b in zip(edges[:-1], edges[1:]) if b > a]
        self.pool = mp.get_context("fork").Pool(cores) if cores > 1 else None

    def step(self):
        if self.pool is None:
            _cpu_worker((0, self.frames))
        else:
            self.pool.map(_cpu_worker, self.spans)

    def close(self):
        if self.pool is not None:
            self.pool.close()
            self.pool.join()


def host_cores() -> int:
    try:
        return len(os.sched_getaffinity(0))
    except AttributeError:
        return os.cpu_count() or 1


def run_cpu_baseline(budget_s: float, sample_frames: int, cores: int):
    ref = CpuReference(sample_frames, cores)
    ref.step()                                   # warm-up (page faults, pocketfft plan cache)
    t0 = time.perf_counter()
    steps = 0
    while True:
        ref.step()
        steps += 1
        dt = time.perf_counter() - t0
        if dt >= budget_s:
            break
    ref.close()
    value = steps * sample_frames * N_FFT / dt
    return {"value": value, "unit": UNIT, "cores": cores, "kind": "port",
            "sample": f"{steps} x {sample_frames} frames of the cfg2 batch (N=4096), per-frame scipy.fft chain "
                      f"of rtl_samples.py:169-184 fanned over {cores} process(es), {dt:.1f} s",
            "numpy": np.__version__}


def reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cores = host_cores()
    sample_frames = BATCH                       # one step = the whole 8192-frame batch, like the GPU arm
    ref = CpuReference(sample_frames, cores)
    for _ in range(max(args.warmup, 1)):
        ref.step()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        ref.step()
    dt = time.perf_counter() - t0
    ref.close()
    value = args.steps * sample_frames * N_FFT / dt
    import scipy
    try:
        affinity = sorted(os.sched_getaffinity(0))
        aff = f"{affinity[0]}-{affinity[-1]}" if affinity == list(range(affinity[0], affinity[-1] + 1)) else str(affinity)
    except AttributeError:
        aff = "n/a"
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": workload_name(args.gpus), "n_fft": N_FFT, "window": "hanning", "mode": "power",
                   "batch_per_step": sample_frames,
                   "per_frame_work": "raw-sample copy + window*FFT*shift*|.|^2*dB + freq-bin rebuild "
                                     "(rtl_samples.py:168-188), one Python call per frame like the reference's tick",
                   "note": "CPU arm: one host, all cores; the same 8192-frame batch per step whatever --gpus says"},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port", "affinity": aff,
                         "sample": f"{sample_frames} frames/step x {args.steps} steps, reference per-frame chain "
                                   f"(scipy {scipy.__version__}, numpy {np.__version__}) over {cores} processes; "
                                   "the reference is pure Python whose tree is absent on the GPU box, so the arm runs the "
                                   "oracle's line-for-line restatement (oracle/oracle.py, pinned by tests/golden)"},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))


# ------------------------------------------------------------------------------------------
# clocks sampler (NVML polled by a child process so that millisecond-long timed regions still get samples)
# ------------------------------------------------------------------------------------------
def _clock_worker(index, conn):
    """Child process: poll NVML as fast as it answers; send (monotonic time, SM MHz, reason bits) tuples on request."""
    try:
        import pynvml as nv
        nv.nvmlInit()
        h = nv.nvmlDeviceGetHandleByIndex(index)
        conn.send(("ready", nv.nvmlDeviceGetMaxClockInfo(h, nv.NVML_CLOCK_SM)))
    except Exception as e:                       # no NVML: the parent records that there are no samples
        conn.send(("error", repr(e)))
        return
    out, errors = [], 0
    while not conn.poll(0):
        try:
            mhz = nv.nvmlDeviceGetClockInfo(h, nv.NVML_CLOCK_SM)
            try:
                r = nv.nvmlDeviceGetCurrentClocksEventReasons(h)
            except Exception:
                r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(h)
            out.append((time.monotonic(), mhz, int(r)))
            time.sleep(0.00005)                  # ~10 kHz is plenty; do not hammer the driver next to the launches
        except Exception:                        # a transient NVML error must not end the sampling
            errors += 1
            time.sleep(0.0002)
    conn.send((out, errors))


class ClockSampler:
    """SM clock and throttle reasons during the timed region, sampled by a separate PROCESS (the launching thread holds
    the GIL for the whole region, which starves an in-process sampler); only samples inside [t0, t1] are kept."""

    NAMES = {0x8: "hw_slowdown", 0x40: "hw_thermal_slowdown", 0x20: "sw_thermal_slowdown", 0x4: "sw_power_cap",
             0x80: "hw_power_brake_slowdown"}

    def __init__(self, index: int):
        import multiprocessing as mp
        self.max_mhz, self.proc = None, None
        try:
            ctx = mp.get_context("spawn")
            self.conn, child = ctx.Pipe()
            self.proc = ctx.Process(target=_clock_worker, args=(index, child), daemon=True)
            self.proc.start()
            if self.conn.poll(30):
                tag, val = self.conn.recv()
                if tag == "ready":
                    self.max_mhz = val
                else:
                    self.proc = None
            else:
                self.proc = None
        except Exception:
            self.proc = None

    def mark_warmup(self):
        self.tw = time.monotonic()

    def start(self):
        self.t0 = time.monotonic()

    def stop(self):
        t1 = time.monotonic()
        samples, errors = [], None
        if self.proc is not None:
            try:
                self.conn.send("stop")
                if self.conn.poll(10):
                    samples, errors = self.conn.recv()
                self.proc.join(timeout=5)
            except Exception:
                samples = []
        timed = [s for s in samples if self.t0 <= s[0] <= t1]
        # the warm-up steps run the same launches back to back with the timed ones: same load, longer window
        inside = [s for s in samples if getattr(self, "tw", self.t0) <= s[0] <= t1]
        if not inside and samples:               # NVML stalled across the whole window: take the closest sample
            mid = 0.5 * (self.t0 + t1)
            inside = [min(samples, key=lambda s: abs(s[0] - mid))]
        reasons = set()
        for _, _, r in inside:
            for bit, name in self.NAMES.items():
                if r & bit:
                    reasons.add(name)
        med = float(np.median([s[1] for s in inside])) if inside else None
        return {"sm_mhz": med, "sm_max_mhz": self.max_mhz, "reasons": sorted(reasons), "samples": len(inside),
                "samples_in_timed_region": len(timed), "timed_region_ms": (t1 - self.t0) * 1e3,
                "window": "warm-up + timed steps (identical launches, back to back)", "nvml_errors": errors}


# ------------------------------------------------------------------------------------------
# config 4 inside the same line: 300 sub-bands x 16 frames x 8192 points, sharded by rank, rows exchanged, grid stitched
# ------------------------------------------------------------------------------------------
CFG4 = {"bands": 300, "frames": 16, "n_fft": 8192, "band_hz": 20e6}
_CFG4_KEEP = {}


def measure_cfg3(dev, reps=5):
    """Config 3 at full size on this rank's GPU: 2^26 samples, 65536-point Welch, hop 32768 (2047 segments) -> mean and
    peak rows.  Input made on the device (noise + two tones); parity of this path at this size is a -m gpu test."""
    import torch
    from topdogspectrumanalyser_b200.engine import SpectrumPlan
    n, hop, total = 65536, 32768, 1 << 26
    g = torch.Generator(device=dev); g.manual_seed(3)
    x = torch.randn(total, 2, generator=g, device=dev, dtype=torch.float32).mul_(0.05)
    t = torch.arange(total, device=dev, dtype=torch.float64)
    for k, a in ((0.1234, 0.5), (-0.3101, 0.05)):
        ph = (2 * np.pi * k) * t
        x[:, 0] += (a * torch.cos(ph)).float(); x[:, 1] += (a * torch.sin(ph)).float()
    del t
    stream = torch.view_as_complex(x)
    nseg = (total - n) // hop + 1
    out = {"workload": f"cfg3: {total} samples, {n}-point segments, hop {hop}: {nseg} segments -> avg + peak rows", "segments": nseg}
    for prec in ("f64", "f32"):
        plan = SpectrumPlan(n, precision=prec, device=dev)
        for _ in range(2):
            avg, peak = plan.welch(stream, hop)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps):
            avg, peak = plan.welch(stream, hop)
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / reps
        out[prec] = {"ms": ms, "input_samples_per_s": total / (ms * 1e-3), "segment_samples_per_s": nseg * n / (ms * 1e-3),
                     "finite": bool(torch.isfinite(avg).all() and torch.isfinite(peak).all()),
                     "peak_bin": int(torch.argmax(avg))}
        plan.close()
    return out


def measure_cfg4(dev, world, rank, precision, reps=20):
    import torch
    import torch.distributed as dist
    from topdogspectrumanalyser_b200 import synth
    from topdogspectrumanalyser_b200.sweep import WidebandSweep, shard_bands, shard_grid
    nb, fr, n = CFG4["bands"], CFG4["frames"], CFG4["n_fft"]
    mine = shard_bands(nb, world, rank)
    iq = torch.from_numpy(synth.cfg4_subbands(nb, fr, n, seed=3, bands=mine)).to(dev)
    out = {"workload": f"cfg4: {nb} sub-bands x {fr} frames x {n} points, {world} rank(s), "
                       f"{len(shard_bands(nb, world, 0))} sub-bands on the largest shard",
           "samples_per_sweep": nb * fr * n, "grid_bins": None}

    def timed(sw, what):
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
        for _ in range(3):
            sw.stitch(sw.all_rows(iq))
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        t_rows = t_st = 0.0
        trials = []
        for _ in range(3):                                  # the loop is short and launch-bound: best of three trials
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(reps):
                ev[0].record()
                rows = sw.all_rows(iq)
                ev[1].record()
                grid = sw.stitch(rows)
                ev[2].record()
            e1.record()
            torch.cuda.synchronize()
            trials.append(e0.elapsed_time(e1) / reps)
        # phase split from the LAST repetition's events (device time), whole-loop time for the rate
        t_rows, t_st = ev[0].elapsed_time(ev[1]), ev[1].elapsed_time(ev[2])
        ms = torch.tensor([min(trials), t_rows, t_st], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        ms = ms.tolist()
        return rows, grid, {"ms_per_sweep": ms[0], "ms_per_sweep_trials": [round(t, 4) for t in trials],
                            "rows_and_exchange_us": ms[1] * 1e3, "stitch_us": ms[2] * 1e3,
                            "samples_per_s": nb * fr * n / (ms[0] * 1e-3), "exchange": sw.exchange, "grid": what}

    modes = [("auto", "sharded")] if world > 1 else [("auto", "replicated")]
    if world > 1:
        modes.append(("nccl", "replicated"))               # round 1's path, for comparison
    for exchange, grid_mode in modes:
        sw = WidebandSweep(nb, CFG4["band_hz"], n, 0.0, precision=precision, device=dev, exchange=exchange, grid=grid_mode)
        rows, grid, res = timed(sw, grid_mode)
        key = "fused" if exchange == "auto" else "nccl_allgather"
        out[key] = res
        out["grid_bins"] = sw.m
        if exchange == "auto":
            if world > 1:
                # every rank recomputes ALL sub-bands on its own GPU (single-GPU path) and compares bit for bit
                full = torch.from_numpy(synth.cfg4_subbands(nb, fr, n, seed=3)).to(dev)
                solo = WidebandSweep(nb, CFG4["band_hz"], n, 0.0, precision=precision, device=dev, exchange="none")
                solo.world, solo.rank = 1, 0
                want_rows = solo.local_rows(full)
                g0, cnt = shard_grid(sw.m, world, rank)
                want_grid = solo.stitch(want_rows, sharded=False)[g0:g0 + cnt]
                ok = torch.tensor([int(torch.equal(rows, want_rows)), int(torch.equal(grid, want_grid))], device=dev)
                dist.all_reduce(ok, op=dist.ReduceOp.MIN)
                out["rows_equal_single_gpu"], out["grid_equal_single_gpu"] = bool(ok[0].item()), bool(ok[1].item())
            else:
                _CFG4_KEEP.update(rows_host=rows.cpu().numpy(), grid_host=grid.cpu().numpy(), m=sw.m)
    return out


def check_cfg4_against_oracle(keep):
    """rows within 1e-4 dB of the float64 chain, grid bit-equal to the reference's stitch (hackrf_sweep.py:150-166)."""
    from oracle import oracle as O
    from topdogspectrumanalyser_b200 import synth
    nb, fr, n = CFG4["bands"], CFG4["frames"], CFG4["n_fft"]
    w = O.make_window("hanning", n)
    worst = 0.0
    for b0 in range(0, nb, 20):
        blk = synth.cfg4_subbands(nb, fr, n, seed=3, bands=range(b0, min(b0 + 20, nb)))
        for i, band in enumerate(blk):
            want = 10 * np.log10(O.linear_power_batch(band, w).mean(axis=0) + 1e-10)
            worst = max(worst, float(np.abs(keep["rows_host"][b0 + i] - want).max()))
    los = [CFG4["band_hz"] * i for i in range(nb)]
    g = O.stitch_rows(keep["rows_host"], los, [l + CFG4["band_hz"] for l in los],
                      O.sweep_grid(0, int(nb * CFG4["band_hz"]), CFG4["band_hz"] / n))
    return {"rows_max_err_db": worst, "grid_equal": bool(np.array_equal(g, keep["grid_host"]))}


# ------------------------------------------------------------------------------------------
# host placement: each rank on the cores (and therefore the DRAM) of its GPU's NUMA node
# ------------------------------------------------------------------------------------------
def _parse_cpulist(text):
    cpus = []
    for part in text.strip().split(","):
        if not part:
            continue
        a, _, b = part.partition("-")
        cpus.extend(range(int(a), int(b or a) + 1))
    return cpus


def bind_to_gpu_numa(index: int) -> dict:
    """Pin this process to the CPU cores of the NUMA node the GPU hangs off, BEFORE any pinned buffer is allocated, so
    that cudaHostAlloc's pages (first touch) land in that node's DRAM.  Returns what was done, for the JSON line."""
    info = {"gpu": index, "numa_node": None, "cpus": None}
    try:
        import pynvml as nv
        nv.nvmlInit()
        bus = nv.nvmlDeviceGetPciInfo(nv.nvmlDeviceGetHandleByIndex(index)).busId
        bus = (bus.decode() if isinstance(bus, bytes) else bus).lower()
        if len(bus.split(":")[0]) == 8:                      # NVML prints an 8-digit domain, sysfs a 4-digit one
            bus = bus[4:]
        node = int(open(f"/sys/bus/pci/devices/{bus}/numa_node").read())
        info["numa_node"] = node
        if node >= 0:
            cpus = _parse_cpulist(open(f"/sys/devices/system/node/node{node}/cpulist").read())
            allowed = sorted(set(cpus) & set(os.sched_getaffinity(0)))
            if allowed:
                os.sched_setaffinity(0, allowed)
                info["cpus"] = f"{allowed[0]}-{allowed[-1]} ({len(allowed)})"
    except Exception as e:                                   # noqa: BLE001 - placement is best effort
        info["error"] = repr(e)[:120]
    return info


# ------------------------------------------------------------------------------------------
# our arm
# ------------------------------------------------------------------------------------------
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--precision", default="f64", choices=["f64", "f32"],
                    help="f64 = the reference's float64 arithmetic (strict 1e-4 dB parity); f32 = fast path")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--e2e-steps", type=int, default=3)
    ap.add_argument("--min-seconds", type=float, default=0.25,
                    help="also run the headline launch back to back for at least this long (sustained clocks); 0 = skip")
    ap.add_argument("--no-cfg4", action="store_true", help="skip the config-4 (wideband stitch) object")
    ap.add_argument("--no-cfg3", action="store_true", help="skip the config-3 (65536-point Welch) object")
    ap.add_argument("--no-other-sizes", action="store_true")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3)

    if args.impl == "reference":
        reference_arm(args)
        return

    import torch
    import torch.distributed as dist

    from topdogspectrumanalyser_b200 import _lib, synth
    from topdogspectrumanalyser_b200.engine import SpectrumPlan

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device; there is no CPU fallback (use --impl reference for the CPU arm)")
    placement = bind_to_gpu_numa(local_rank)                # before the first pinned allocation
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        # keep stdout to the ONE JSON line: with NCCL_DEBUG set (the GPU boxes export it) NCCL prints its version banner
        # to stdout when the first communicator comes up, so file descriptor 1 points at stderr until that has happened
        sys.stdout.flush()
        saved_fd = os.dup(1)
        os.dup2(2, 1)
        try:
            dist.init_process_group("nccl", device_id=dev)
            dist.barrier()
            torch.cuda.synchronize()
        finally:
            sys.stdout.flush()
            os.dup2(saved_fd, 1)
            os.close(saved_fd)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- inputs: each rank owns its own 8192-frame batch (different seed per rank) ----------
    iq_host_t = torch.empty((BATCH, N_FFT), dtype=torch.complex64).pin_memory()
    iq_host = iq_host_t.numpy()
    base = synth.cfg2_frames(b=1024, n=N_FFT, seed=1 + rank)
    for i in range(BATCH // 1024):                         # 8 phase-rotated copies: distinct frames, cheap to build
        iq_host[i * 1024:(i + 1) * 1024] = base * np.complex64(np.exp(1j * 0.37 * i))
    db_host_t = torch.empty((BATCH, N_FFT), dtype=torch.float32).pin_memory()
    db_host = db_host_t.numpy()
    x = iq_host_t.to(dev)
    out = torch.empty((BATCH, N_FFT), dtype=torch.float32, device=dev)
    samples_per_step = BATCH * N_FFT

    def measure_size(n_fft, batch, precision, steps=10, warmup=3):
        """Device time of kernel 1 at another FFT size (same bytes per launch as the headline batch)."""
        xs = x.view(batch, n_fft)
        os_ = out.view(batch, n_fft)
        plan = SpectrumPlan(n_fft, "hanning", mode="power", precision=precision, device=dev)
        for _ in range(warmup):
            plan.psd_db(xs, out=os_)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            plan.psd_db(xs, out=os_)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / steps
        info = plan.info()
        plan.close()
        return {"kernel_ms": ms, "roofline_frac": BYTES_PER_SAMPLE * samples_per_step / (ms * 1e-3) / 1e9 / peak,
                "threads_per_cta": info["threads_per_cta"], "ctas_per_sm": info["ctas_per_sm"]}

    def measure_sustained(precision, seconds):
        """The headline launch back to back for >= `seconds` (power-capped clocks), CUDA events, max over ranks."""
        plan = SpectrumPlan(N_FFT, "hanning", mode="power", precision=precision, device=dev)
        for _ in range(3):
            plan.psd_db(x, out=out)
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        steps, t0 = 0, time.perf_counter()
        e0.record()
        while True:
            for _ in range(100):
                plan.psd_db(x, out=out)
            steps += 100
            if time.perf_counter() - t0 >= seconds:        # host time only decides when to stop queueing
                break
        e1.record()
        torch.cuda.synchronize()
        ms = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        plan.close()
        total_ms = float(ms.item())
        return {"seconds": total_ms * 1e-3, "steps": steps, "ms_per_step": total_ms / steps,
                "value": world * samples_per_step * steps / (total_ms * 1e-3), "unit": UNIT,
                "roofline_frac": BYTES_PER_SAMPLE * samples_per_step * steps / (total_ms * 1e-3) / 1e9 / peak}

    def measure(precision, steps, warmup, sample_clocks):
        plan = SpectrumPlan(N_FFT, "hanning", mode="power", precision=precision, device=dev)
        sampler = ClockSampler(local_rank) if sample_clocks else None     # child process is polling from here on
        if sampler:
            plan.psd_db(x, out=out)                                       # module load / first launch before the window opens
            torch.cuda.synchronize()
            sampler.mark_warmup()
        for _ in range(warmup):
            plan.psd_db(x, out=out)
        barrier()
        if sampler:
            sampler.start()                                               # only samples from now on are kept
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(steps + 1)]
        n0 = _lib.launch_count()
        ev[0].record()
        for i in range(steps):
            plan.psd_db(x, out=out)
            ev[i + 1].record()
        barrier()
        launches = _lib.launch_count() - n0
        clocks = sampler.stop() if sampler else None
        total_ms = ev[0].elapsed_time(ev[steps])
        per = np.array([ev[i].elapsed_time(ev[i + 1]) for i in range(steps)])
        t = torch.tensor([total_ms], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        info = plan.info()
        plan.close()
        return {"total_ms": float(t.item()), "kernel_ms": float(per.mean()), "kernel_ms_best": float(per.min()),
                "launches": launches, "clocks": clocks, "info": info}

    def measure_e2e(precision, steps):
        plan = SpectrumPlan(N_FFT, "hanning", mode="power", precision=precision, device=dev)
        plan.psd_db_host(iq_host, db_host)                  # warm-up (allocates the staging buffers)
        barrier()
        t0 = time.perf_counter()
        for _ in range(steps):
            plan.psd_db_host(iq_host, db_host)              # blocks until the dB rows are in host memory
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        t = torch.tensor([dt], dtype=torch.float64, device=dev)
        per_rank = [torch.zeros(1, dtype=torch.float64, device=dev) for _ in range(world)]
        if world > 1:
            dist.all_gather(per_rank, t.clone())
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        else:
            per_rank = [t.clone()]
        plan.close()
        secs = [float(v.item()) / steps for v in per_rank]
        return float(t.item()) / steps, secs

    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    peak = float(peaks.get("hbm_gbs", 6650.0))

    def measure_host_link():
        """What the host <-> device links give when every rank copies at once and nothing else runs: plain pinned
        cudaMemcpyAsync of the e2e step's buffers, H2D and D2H concurrently on two streams (the ceiling of `e2e`)."""
        s_in, s_out = torch.cuda.Stream(), torch.cuda.Stream()
        reps = 3
        for timed in (False, True):
            barrier()
            t0 = time.perf_counter()
            for _ in range(reps if timed else 1):
                with torch.cuda.stream(s_in):
                    x.copy_(iq_host_t, non_blocking=True)
                with torch.cuda.stream(s_out):
                    db_host_t.copy_(out, non_blocking=True)
            s_in.synchronize()
            s_out.synchronize()
            dt = (time.perf_counter() - t0) / reps
        t = torch.tensor([dt], dtype=torch.float64, device=dev)
        per_rank = [torch.zeros(1, dtype=torch.float64, device=dev) for _ in range(world)]
        if world > 1:
            dist.all_gather(per_rank, t)
        else:
            per_rank = [t]
        secs = [float(v.item()) for v in per_rank]
        return {"what": "all ranks at once: 268 MB pinned H2D + 134 MB pinned D2H on two streams, no kernel",
                "per_rank_h2d_gbs": [round(samples_per_step * 8 / v / 1e9, 1) for v in secs],
                "aggregate_h2d_plus_d2h_gbs": round(sum(samples_per_step * 12 / v for v in secs) / 1e9, 1),
                "samples_per_s_if_copy_bound": sum(samples_per_step / v for v in secs)}

    main_run = measure(args.precision, args.steps, args.warmup, True)
    other = "f32" if args.precision == "f64" else "f64"
    other_run = measure(other, min(args.steps, 20), 3, False)
    sustained = measure_sustained(args.precision, args.min_seconds) if args.min_seconds > 0 else None
    e2e_s, e2e_per_rank = measure_e2e(args.precision, args.e2e_steps)
    host_link = measure_host_link()
    other_sizes = None
    if not args.no_other_sizes:
        other_sizes = {}
        for n_fft in (1024, 2048, 8192):
            other_sizes[str(n_fft)] = {pr: measure_size(n_fft, BATCH * N_FFT // n_fft, pr) for pr in ("f64", "f32")}
    cfg4 = None
    if not args.no_cfg4:
        try:
            cfg4 = measure_cfg4(dev, world, rank, args.precision)
        except Exception as e:                              # noqa: BLE001 - the headline line must still be printed
            cfg4 = {"error": repr(e)[:300]}
    cfg3 = None
    if not args.no_cfg3 and rank == 0:
        try:
            cfg3 = measure_cfg3(dev)
        except Exception as e:                              # noqa: BLE001
            cfg3 = {"error": repr(e)[:300]}
    peak_src = "MEASURED_PEAKS.json hbm_gbs (burst copy)" if "hbm_gbs" in peaks else "fallback 6650 GB/s (B200_PROFILING.md)"
    traffic = None
    try:
        traffic = json.load(open(os.path.join(ROOT, "profiles", "traffic.json"))).get(args.precision)
    except Exception:
        pass

    def sm_side(precision, kernel_ms):
        """Arithmetic-pipe view of the same launch, from the instruction mix in profiles/ (per thread-frame)."""
        threads = BATCH * (N_FFT // 16)
        if precision == "f64":
            ops = 632                                   # DADD + DMUL + DFMA per thread-frame (ncu source view)
            peak = 148 * 64 * float(peaks.get("sm_max_mhz", 1965.0)) * 1e6      # FP64 lanes/clk/SM x SMs x clock
            pipe = "fp64"
        else:
            ops = 612                                   # FADD + FMUL + FFMA per thread-frame
            peak = 148 * 128 * float(peaks.get("sm_max_mhz", 1965.0)) * 1e6
            pipe = "fp32"
        rate = ops * threads / (kernel_ms * 1e-3)
        return {"pipe": pipe, "thread_instr_per_s": rate, "peak_thread_instr_per_s": peak, "frac": rate / peak}

    def roof(run):
        achieved = BYTES_PER_SAMPLE * samples_per_step / (run["kernel_ms"] * 1e-3) / 1e9
        return achieved

    if rank == 0:
        value = world * samples_per_step * args.steps / (main_run["total_ms"] * 1e-3)
        achieved = roof(main_run)
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": main_run["total_ms"] / args.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": args.precision, "data": "synthetic",
            "config": {"workload": workload_name(world), "n_fft": N_FFT, "batch_per_gpu": BATCH, "window": "hanning",
                       "mode": "power", "precision": args.precision,
                       "l2": "inputs larger than L2: 268 MB read + 134 MB written per step vs 126 MB L2",
                       "kernel": main_run["info"]},
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "traffic": traffic, "peak_source": peak_src,
                         "kernel": (f"fft_wl_kernel<{'double' if args.precision == 'f64' else 'float'},EpiDb,NB=1,ACC=0> "
                                    "(warp-local 16x256 plan, swizzled TMA tensor staging, dynamic frame scheduling)"
                                    if os.environ.get("TDSA_WL", "1") != "0" else
                                    f"fft_fused_kernel<{'double' if args.precision == 'f64' else 'float'},12,EpiDb>"),
                         "algorithmic_bytes_per_launch": BYTES_PER_SAMPLE * samples_per_step,
                         "kernel_ms_avg": main_run["kernel_ms"], "kernel_ms_best": main_run["kernel_ms_best"],
                         "note": ("HBM is the bound the metric names; the float64 kernel is limited by the FP64 pipe (DFMA-class "
                                  "work plus the float<->double conversions that run on it: ~3000 of ~4400 cycles per frame "
                                  "and SM) and by how well FP sections overlap shared-memory phases, not by HBM" if args.precision == "f64" else
                                  "ncu: issue slots and the L1/shared data pipe co-limit with HBM (see DESIGN.md 3.1)"),
                         "sm_side": sm_side(args.precision, main_run["kernel_ms"])},
            "e2e": {"value": world * samples_per_step / e2e_s, "unit": UNIT,
                    "h2d_bytes_per_step": samples_per_step * 8, "d2h_bytes_per_step": samples_per_step * 4,
                    "api": "SpectrumPlan.psd_db_host -> tdsa_psd_db_batch_host (pinned host in, pinned host out; "
                           "H2D, kernel and D2H on three streams)",
                    "ms_per_step": e2e_s * 1e3,
                    "per_rank_h2d_gbs": [round(samples_per_step * 8 / t / 1e9, 1) for t in e2e_per_rank],
                    "per_rank_d2h_gbs": [round(samples_per_step * 4 / t / 1e9, 1) for t in e2e_per_rank],
                    "host_placement": placement, "host_link_ceiling": host_link},
            "sustained": sustained,
            "other_sizes": other_sizes,
            "cfg3": cfg3,
            "cfg4": cfg4,
            "gpu_launches": main_run["launches"],
            "clocks": main_run["clocks"],
            "other_precision": {"precision": other, "value": world * samples_per_step / (other_run["kernel_ms"] * 1e-3),
                                "unit": UNIT, "roofline_frac": roof(other_run) / peak,
                                "kernel_ms_avg": other_run["kernel_ms"],
                                "note": "f32 = fast path (deep-null tail, see DESIGN.md); f64 = strict 1e-4 dB parity"},
        }
        if world == 1 and not args.no_cpu_baseline:
            os.sched_setaffinity(0, range(os.cpu_count() or 1)) if hasattr(os, "sched_setaffinity") else None   # all cores again
            cores = host_cores()
            line["cpu_baseline"] = run_cpu_baseline(12.0, 2048, cores)
            line["cpu_baseline"]["single_core"] = run_cpu_baseline(4.0, 512, 1)["value"]
            if cfg4 is not None and "rows_host" in _CFG4_KEEP:
                cfg4.update(check_cfg4_against_oracle(_CFG4_KEEP))      # the checker, in the CPU leg only
        _CFG4_KEEP.clear()
        print(json.dumps(line))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
