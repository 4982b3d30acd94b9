#!/bin/bash
# cfg3 fused kernel: ring depth variants
mkdir -p gpurun_out
cat > /tmp/cfg3_time.py <<'PY'
import os, sys, numpy as np, torch
sys.path.insert(0, os.getcwd())
from topdogspectrumanalyser_b200 import synth
from topdogspectrumanalyser_b200.engine import SpectrumPlan
dev = torch.device("cuda:0")
stream = torch.from_numpy(synth.cfg3_stream(1 << 26, seed=2)).to(dev)
ref = {}
for fused in ("0", "1"):
    os.environ["TDSA_WELCH_FUSED"] = fused
    for prec in ("f64", "f32"):
        plan = SpectrumPlan(65536, precision=prec, device=dev)
        for _ in range(3): out = plan.welch(stream, 32768)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(10): plan.welch(stream, 32768)
        e1.record(); torch.cuda.synchronize()
        if fused == "0": ref[prec] = out
        d = max(float((out[0] - ref[prec][0]).abs().max()), float((out[1] - ref[prec][1]).abs().max()))
        print(os.environ.get("TDSA_LIB", "default"), "cfg3 fused", fused, prec, "ms %.4f" % (e0.elapsed_time(e1) / 10), "vs two-launch max diff dB %.2e" % d, flush=True)
        plan.close()
PY
for lib in "" variants/libtdsa_ring4.so variants/libtdsa_ring6.so; do
  if [ -n "$lib" ]; then export TDSA_LIB=$PWD/$lib; fi
  timeout 300 python /tmp/cfg3_time.py 2>&1 | grep cfg3
done
