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
import os
for path, env in (("head kernel + warp-local tail kernel (default)", {"TDSA_WELCH_SUB": "1", "TDSA_WELCH_FUSED": "0"}),
                  ("head + warp-local tails in one kernel, L2 ring (opt-in)", {"TDSA_WELCH_SUB": "1", "TDSA_WELCH_FUSED": "1"}),
                  ("cluster kernel (round 1)", {"TDSA_WELCH_SUB": "0", "TDSA_WELCH_CLUSTER": "1"}),
                  ("two kernels + linear rows (round 1)", {"TDSA_WELCH_SUB": "0", "TDSA_WELCH_CLUSTER": "0"})):
    os.environ.update(env)
    for prec in ("f64", "f32"):
        plan = SpectrumPlan(65536, precision=prec, device=dev)
        t = ev_time(lambda: plan.welch(stream, 32768), 5)
        print(json.dumps({"config": "cfg3 welch 65536-pt, 50% overlap, 2^26 samples, 2047 segments", "path": path, "precision": prec,
                          "ms": t * 1e3, "input_samples_per_s": (1 << 26) / t, "segment_samples_per_s": 2047 * 65536 / t,
                          "hbm_frac_8B_per_input_sample": 8 * (1 << 26) / t / 6534.1e9}), flush=True)
        plan.close()
for k in ("TDSA_WELCH_SUB", "TDSA_WELCH_CLUSTER", "TDSA_WELCH_FUSED"): os.environ.pop(k, None)
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
    # the same 10 s of signal with the producer writing into the pinned slot in place (acquire / commit: what a device
    # reader thread does), i.e. without the 512 KB host copy of push_chunk
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for c in range(3052):
        st.acquire(); st.commit()
    torch.cuda.synchronize(); inplace_s = time.perf_counter() - t0
    # per-chunk latency: push one chunk and wait until its rows are visible in the ring
    torch.cuda.synchronize(); lat = []
    for c in range(50):
        t0 = time.perf_counter(); st.push_chunk(chunks[c]); torch.cuda.synchronize(); lat.append(time.perf_counter() - t0)
    # timeline of a burst: the producer has filled every pinned slot in place (acquire), then the slots are committed
    # back to back, so the host loop is out of the way and the copy of chunk k+1 can be seen under the compute of chunk k
    torch.cuda.synchronize()
    for i in range(st.depth):
        st.pinned[(st.chunks_in + i) % st.depth].numpy()[:] = chunks[i]
    st.record_timeline(True)
    for _ in range(6):
        for i in range(st.depth):
            st.commit()
    tl = st.timeline()
    st.record_timeline(False)
    # copy of chunk k+1 under the compute of chunk k: overlap of the two intervals, summed over the recorded chunks
    ov = sum(max(0.0, min(tl[k + 1]["h2d_us"][1], tl[k]["compute_us"][1]) - max(tl[k + 1]["h2d_us"][0], tl[k]["compute_us"][0]))
             for k in range(len(tl) - 1))
    h2d = sum(t["h2d_us"][1] - t["h2d_us"][0] for t in tl[1:])
    if prec == "f64":
        os.makedirs("gpurun_out", exist_ok=True)
        json.dump(tl, open("gpurun_out/cfg5_timeline.json", "w"))
    print(json.dumps({"config": "cfg5 streaming waterfall 20 Msps x 10 s, 65536-sample pinned chunks", "precision": prec,
                      "seconds_for_10s_of_signal": stats["seconds"], "real_time_factor": stats["real_time_factor"],
                      "samples_per_s": stats["samples_per_s"], "chunk_latency_ms_median": float(np.median(lat)) * 1e3,
                      "in_place_producer": {"seconds_for_10s_of_signal": inplace_s, "real_time_factor": 3052 * 65536 / inplace_s / 20e6},
                      "graphs": st.use_graphs, "h2d_under_previous_compute_frac": ov / h2d if h2d else None,
                      "timeline_first_chunks": tl[:4]}))
