#!/usr/bin/env python
"""Full-size runs of BASELINE.json configs 3, 4 (one GPU) and 5; one JSON line each (for profiles/)."""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from topdogspectrumanalyser_b200 import synth
from topdogspectrumanalyser_b200.engine import SpectrumPlan
from topdogspectrumanalyser_b200.streaming import WaterfallStreamer
from topdogspectrumanalyser_b200.sweep import WidebandSweep
dev = torch.device("cuda:0")

def ev_time(fn, reps):
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps * 1e-3

# config 3: 2^26 samples, N = 65536, hop 32768 -> 2047 segments, avg + peak rows
stream = torch.from_numpy(synth.cfg3_stream(1 << 26, seed=2)).to(dev)
for prec in ("f64", "f32"):
    plan = SpectrumPlan(65536, precision=prec, device=dev)
    t = ev_time(lambda: plan.welch(stream, 32768), 5)
    print(json.dumps({"config": "cfg3 welch 65536-pt, 50% overlap, 2^26 samples, 2047 segments", "precision": prec,
                      "ms": t * 1e3, "input_samples_per_s": (1 << 26) / t, "segment_samples_per_s": 2047 * 65536 / t,
                      "hbm_frac_8B_per_input_sample": 8 * (1 << 26) / t / 6534.1e9}))
    plan.close()
del stream
# config 4 on one GPU: 300 sub-bands x 16 frames x 8192
iq = torch.from_numpy(synth.cfg4_subbands(300, 16, 8192, seed=3)).to(dev)
for prec in ("f64", "f32"):
    sw = WidebandSweep(300, 20e6, 8192, 0.0, precision=prec, device=dev)
    t = ev_time(lambda: sw.run(iq), 10)
    print(json.dumps({"config": "cfg4 wideband stitch 300x16x8192 -> 2457600-bin grid (1 GPU)", "precision": prec,
                      "ms_per_sweep": t * 1e3, "samples_per_s": 300 * 16 * 8192 / t}))
del iq
# config 5: 20 Msps for 10 s in 65536-sample chunks (3052 chunks), N=4096, exp avg n=8, ring H=1024
chunks = [synth.cfg5_chunk(c) for c in range(64)]
for prec in ("f64", "f32"):
    st = WaterfallStreamer(4096, 65536, 1024, "exp", 8, precision=prec, device=dev)
    st.run(lambda c: chunks[c % 64], 64)
    stats = st.run(lambda c: chunks[c % 64], 3052)
    # per-chunk latency: push one chunk and wait until its rows are visible in the ring
    torch.cuda.synchronize(); lat = []
    for c in range(50):
        t0 = time.perf_counter(); st.push_chunk(chunks[c]); torch.cuda.synchronize(); lat.append(time.perf_counter() - t0)
    print(json.dumps({"config": "cfg5 streaming waterfall 20 Msps x 10 s, 65536-sample pinned chunks", "precision": prec,
                      "seconds_for_10s_of_signal": stats["seconds"], "real_time_factor": stats["real_time_factor"],
                      "samples_per_s": stats["samples_per_s"], "chunk_latency_ms_median": float(np.median(lat)) * 1e3}))
