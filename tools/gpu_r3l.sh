#!/bin/bash
timeout 900 python -m pytest tests/test_gpu_kernel1.py tests/test_gpu_state.py tests/test_gpu_next_rows.py -q -x 2>&1 | tail -2
timeout 300 python /dev/stdin <<'PY'
import os, sys, torch
sys.path.insert(0, os.getcwd())
from topdogspectrumanalyser_b200 import synth
from topdogspectrumanalyser_b200.engine import SpectrumPlan
dev = torch.device("cuda:0")
for n in (256, 512, 1024, 2048):
    for prec in ("f64", "f32"):
        b = 33554432 // n
        x = torch.from_numpy(synth.cfg2_frames(b=1024, n=n, seed=1)).to(dev).repeat(b // 1024, 1).contiguous()
        out = torch.empty((b, n), dtype=torch.float32, device=dev)
        plan = SpectrumPlan(n, precision=prec, device=dev)
        for _ in range(3): plan.psd_db(x, out=out)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(20): plan.psd_db(x, out=out)
        e1.record(); torch.cuda.synchronize()
        t = e0.elapsed_time(e1) / 20 * 1e3
        print(n, prec, "%.1f us  %.3f of HBM" % (t, 12 * 33554432 / t / 1e3 / 6547.8), plan.info(), flush=True)
        plan.close()
PY
