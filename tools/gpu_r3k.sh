#!/bin/bash
# N = 1024 / 2048 float64: more CTAs per SM (launch-bounds variants)
cat > /tmp/ts.py <<'PY'
import os, sys, torch, numpy as np
sys.path.insert(0, os.getcwd())
from topdogspectrumanalyser_b200 import synth
from topdogspectrumanalyser_b200.engine import SpectrumPlan
dev = torch.device("cuda:0")
for n in (512, 1024, 2048):
    b = 33554432 // n
    x = torch.from_numpy(synth.cfg2_frames(b=1024, n=n, seed=1)).to(dev).repeat(b // 1024, 1).contiguous()
    out = torch.empty((b, n), dtype=torch.float32, device=dev)
    plan = SpectrumPlan(n, precision="f64", device=dev)
    for _ in range(3): plan.psd_db(x, out=out)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20): plan.psd_db(x, out=out)
    e1.record(); torch.cuda.synchronize()
    print(os.environ.get("TDSA_LIB", "default")[-16:], n, "f64 %.1f us" % (e0.elapsed_time(e1) / 20 * 1e3), plan.info(), flush=True)
    plan.close()
PY
for lib in "" variants/libtdsa_sc5.so variants/libtdsa_sc6.so variants/libtdsa_sc7.so; do
  if [ -n "$lib" ]; then export TDSA_LIB=$PWD/$lib; fi
  timeout 300 python /tmp/ts.py
done
