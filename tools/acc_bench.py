#!/usr/bin/env python
"""Device time of the trace-state paths over the config-2 batch (8192 frames x 4096 points), CUDA events, for profiles/."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from topdogspectrumanalyser_b200 import synth
from topdogspectrumanalyser_b200.engine import SpectrumPlan, TraceState
dev = torch.device("cuda:0")
PEAK = json.load(open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json"))).get("hbm_gbs", 6650.0) \
    if os.path.exists(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")) else 6650.0

def ev_time(fn, reps=10):
    for _ in range(2): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps * 1e-3

base = torch.from_numpy(synth.cfg2_frames(b=1024, n=4096, seed=1)).to(dev)
x = base.repeat(8, 1).contiguous()
out = torch.empty((8192, 4096), dtype=torch.float32, device=dev)
samples = 8192 * 4096
for prec in ("f64", "f32"):
    plan = SpectrumPlan(4096, precision=prec, device=dev)
    st = TraceState(4096, dev); st.set_averaging("exp", 8)
    t = ev_time(lambda: plan.psd_db_avg_hold(x, st, last_only=True))
    print(json.dumps({"path": "fused running average (last row only), 8192 x 4096", "precision": prec, "us": t * 1e6,
                      "algorithmic_bytes": 8 * samples, "hbm_frac": 8 * samples / t / 1e9 / PEAK}))
    st2 = TraceState(4096, dev, max_hold_enabled=True, min_hold_enabled=True)
    t = ev_time(lambda: plan.psd_db_avg_hold(x, st2, out=out))
    print(json.dumps({"path": "fused max/min hold + every dB row, 8192 x 4096", "precision": prec, "us": t * 1e6,
                      "algorithmic_bytes": 12 * samples, "hbm_frac": 12 * samples / t / 1e9 / PEAK}))
    st3 = TraceState(4096, dev, max_hold_enabled=True); st3.set_averaging("exp", 8)
    t = ev_time(lambda: plan.psd_db_avg_hold(x, st3, out=out), reps=3)
    print(json.dumps({"path": "general path: averaging + hold + every row (L2-chunked float64 rows + frame-ordered scan)",
                      "precision": prec, "us": t * 1e6, "algorithmic_bytes": 12 * samples, "hbm_frac": 12 * samples / t / 1e9 / PEAK}))
    t = ev_time(lambda: plan.welch(x.view(-1), 2048))
    nseg = (samples - 4096) // 2048 + 1
    print(json.dumps({"path": f"Welch 4096-pt, 50% overlap, {nseg} segments (sum + max in TMEM)", "precision": prec, "us": t * 1e6,
                      "segment_samples_per_s": nseg * 4096 / t}))
    plan.close()
g = torch.from_numpy(synth.cfg4_subbands(300, 16, 8192, seed=3)).to(dev)
for prec in ("f64", "f32"):
    plan8 = SpectrumPlan(8192, precision=prec, device=dev)
    t = ev_time(lambda: plan8.group_avg_db(g))
    print(json.dumps({"path": "group mean 300 x 16 x 8192 (two-engine kernel)", "precision": prec, "us": t * 1e6,
                      "samples_per_s": g.numel() / t}))
    plan8.close()
