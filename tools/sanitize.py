#!/usr/bin/env python
"""Small workload for compute-sanitizer: every kernel family once, few frames."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from topdogspectrumanalyser_b200 import synth
from topdogspectrumanalyser_b200.engine import SpectrumPlan, TraceState, trace_update, stitch, WaterfallRing
from topdogspectrumanalyser_b200 import analytics as A
dev = torch.device("cuda:0")
for n in (512, 1024, 4096, 8192, 65536):
    for prec in ("f64", "f32"):
        b = 700 if n == 4096 else (3 if n > 8192 else 40)       # 4096: more frames than resident CTAs -> ring refills
        x = torch.from_numpy(synth.cfg2_frames(b=b, n=n, seed=n)).to(dev)
        plan = SpectrumPlan(n, precision=prec, device=dev)
        y = plan.psd_db(x)
        st = TraceState(n, dev); st.set_averaging("exp", 4); st.max_hold_enabled = st.min_hold_enabled = True
        if n <= 8192:
            plan.psd_db_avg_hold(x[:8], st)
            plan.psd_db_dc(x[:4], torch.zeros(2, dtype=torch.float64, device=dev))
        torch.cuda.synchronize(); plan.close()
        print("ok", n, prec, float(y[0, 0]))
# round 2: reducing epilogues in tensor memory, two-engine 8192 kernel, blocked scan, display kernels
for prec in ("f64", "f32"):
    x = torch.from_numpy(synth.cfg2_frames(b=320, n=4096, seed=7)).to(dev)
    plan = SpectrumPlan(4096, precision=prec, device=dev)
    st = TraceState(4096, dev); st.set_averaging("lin", 100)
    plan.psd_db_avg_hold(x, st, last_only=True)                      # fused running average (ACC = sum)
    st = TraceState(4096, dev, max_hold_enabled=True, min_hold_enabled=True)
    plan.psd_db_avg_hold(x, st)                                       # fused holds (ACC = max | min | rows)
    st = TraceState(4096, dev, max_hold_enabled=True); st.set_averaging("exp", 8)
    plan.psd_db_avg_hold(x, st)                                       # general path, blocked scan
    plan.welch(x.view(-1), 2048)                                      # Welch 4096 (ACC = sum | max)
    plan.group_avg_db(x.view(20, 16, 4096))                           # group mean, split into units
    plan.close()
    g = torch.from_numpy(synth.cfg4_subbands(9, 4, 8192, seed=3)).to(dev)
    plan8 = SpectrumPlan(8192, precision=prec, device=dev)
    plan8.group_avg_db(g)                                             # two engines, named barriers, stage hand-back
    plan8.group_avg_db(g[:, :1].contiguous())
    torch.cuda.synchronize(); plan8.close()
    print("ok round-2 kernels", prec)
# config 3: head kernel (cp.async staging) + warp-local tails (direct loads, per-class claims); the fused one-kernel variant
# (group counters in global memory, L2 ring); the round-1 cluster kernel and two-kernel path
s3 = torch.from_numpy(synth.cfg3_stream(65536 + 32768 * 20)).to(dev)
for env in ({}, {"TDSA_WELCH_FUSED": "1"}, {"TDSA_WELCH_SUB": "0"}, {"TDSA_WELCH_SUB": "0", "TDSA_WELCH_CLUSTER": "0"}):
    for k in ("TDSA_WELCH_FUSED", "TDSA_WELCH_SUB", "TDSA_WELCH_CLUSTER"):
        os.environ.pop(k, None)
    os.environ.update(env)
    for prec in ("f64", "f32"):
        plan = SpectrumPlan(65536, precision=prec, device=dev)
        a, pk = plan.welch(s3, 32768)
        torch.cuda.synchronize(); plan.close()
        print("ok cfg3", env, prec, float(a[0]), bool(torch.isfinite(a).all()))
for k in ("TDSA_WELCH_FUSED", "TDSA_WELCH_SUB", "TDSA_WELCH_CLUSTER"):
    os.environ.pop(k, None)
rows = torch.randn(6, 4096, device=dev)
st = TraceState(4096, dev); st.start_tare(); st.max_hold_enabled = True
trace_update(rows, st, -1.0)
A.top_peaks(np.arange(4096.0), rows[0]); A.DensityHistogram(4096, dev).update(rows[0])
ring = WaterfallRing(8, 4096, -100.0, dev); ring.push(rows)
ring2 = WaterfallRing(8, 4096, -100.0, dev, dedupe=True); ring2.push(torch.cat([rows, rows[-1:]]))
ring2.image(-100.0, 0.0, torch.zeros((256, 4), dtype=torch.uint8, device=dev))
A.snap_to_peak(np.arange(4096.0), rows[0])
stitch(rows[:, :50].contiguous(), torch.arange(6, dtype=torch.float64, device=dev) * 5e6, 5e6, 0.0, 30e6, 300)
torch.cuda.synchronize(); print("all ok")
