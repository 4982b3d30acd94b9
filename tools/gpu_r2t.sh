#!/bin/bash
# N = 8192 pair mode: parity + timings (rows, group mean), A/B against TDSA_WL_PAIR=0
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_round2.py tests/test_gpu_kernel1.py tests/test_gpu_state.py -q -x -k "8192 or group or size or wideband or kernel1" 2>&1 | tail -5
timeout 600 python -m pytest tests/test_gpu_full_size.py -q -x -k "cfg4" 2>&1 | tail -3
cat > /tmp/t8192.py <<'PY'
import os, sys, json, torch
sys.path.insert(0, os.getcwd())
from topdogspectrumanalyser_b200 import synth
from topdogspectrumanalyser_b200.engine import SpectrumPlan
dev = torch.device("cuda:0")
def ev_time(fn, reps=20):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps * 1e3
base = torch.from_numpy(synth.cfg2_frames(b=512, n=8192, seed=1)).to(dev)
x = base.repeat(8, 1).contiguous()
out = torch.empty((4096, 8192), dtype=torch.float32, device=dev)
g = torch.from_numpy(synth.cfg4_subbands(300, 16, 8192, seed=3)).to(dev)
for pair in ("1", "0"):
    os.environ["TDSA_WL_PAIR"] = pair
    for prec in ("f64", "f32"):
        plan = SpectrumPlan(8192, precision=prec, device=dev)
        t = ev_time(lambda: plan.psd_db(x, out=out))
        t2 = ev_time(lambda: plan.group_avg_db(g))
        print(f"pair={pair} {prec}: rows 4096x8192 {t:.1f} us ({12*4096*8192/t/1e3/6534.1:.3f} of HBM); group mean 300x16x8192 {t2:.1f} us", flush=True)
        plan.close()
PY
timeout 300 python /tmp/t8192.py
