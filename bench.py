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


def _cpu_worker(span):
    """Literal per-frame loop of datasources/rtl_samples.py:169-184 over frames [lo, hi)."""
    from oracle import oracle as O
    lo, hi = span
    for f in range(lo, hi):
        _W_OUT[f] = O.power_db_frame(_W_IQ[f], _W_WIN, O.MODE_POWER)
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
    sample_frames = 2048
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
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": workload_name(args.gpus), "n_fft": N_FFT, "window": "hanning", "mode": "power",
                   "step_sample": f"{sample_frames} frames per step (bounded sample of the 8192-frame batch)"},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port",
                         "sample": f"{sample_frames} frames/step, reference per-frame chain "
                                   f"(scipy {scipy.__version__}, numpy {np.__version__}) over {cores} processes"},
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
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        plan.close()
        return float(t.item()) / steps

    main_run = measure(args.precision, args.steps, args.warmup, True)
    other = "f32" if args.precision == "f64" else "f64"
    other_run = measure(other, min(args.steps, 20), 3, False)
    e2e_s = measure_e2e(args.precision, args.e2e_steps)

    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    peak = float(peaks.get("hbm_gbs", 6650.0))
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
                         "kernel": (f"fft_wl_kernel<{'double' if args.precision == 'f64' else 'float'},EpiDb> "
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
                    "api": "SpectrumPlan.psd_db_host -> tdsa_psd_db_batch_host (pinned host in, pinned host out)",
                    "ms_per_step": e2e_s * 1e3},
            "gpu_launches": main_run["launches"],
            "clocks": main_run["clocks"],
            "other_precision": {"precision": other, "value": world * samples_per_step / (other_run["kernel_ms"] * 1e-3),
                                "unit": UNIT, "roofline_frac": roof(other_run) / peak,
                                "kernel_ms_avg": other_run["kernel_ms"],
                                "note": "f32 = fast path (deep-null tail, see DESIGN.md); f64 = strict 1e-4 dB parity"},
        }
        if world == 1 and not args.no_cpu_baseline:
            cores = host_cores()
            line["cpu_baseline"] = run_cpu_baseline(12.0, 2048, cores)
            line["cpu_baseline"]["single_core"] = run_cpu_baseline(4.0, 512, 1)["value"]
        print(json.dumps(line))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
