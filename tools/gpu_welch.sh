#!/bin/bash
mkdir -p gpurun_out
cat > /tmp/welch_time.py <<'PY'
import sys, os, json, torch
sys.path.insert(0, os.getcwd())
from topdogspectrumanalyser_b200 import synth
from topdogspectrumanalyser_b200.engine import SpectrumPlan
dev = torch.device("cuda:0")
stream = torch.from_numpy(synth.cfg3_stream(1 << 26, seed=2)).to(dev)
for prec in ("f64", "f32"):
    plan = SpectrumPlan(65536, precision=prec, device=dev)
    plan.welch(stream, 32768); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5): a, p = plan.welch(stream, 32768)
    e1.record(); torch.cuda.synchronize()
    print(prec, "cfg3 welch ms:", e0.elapsed_time(e1) / 5)
    plan.close()
PY
timeout -s KILL 300 python -m pytest tests/test_gpu_state.py -m gpu -x -q -k welch 2>&1 | tail -3
echo "== cluster kernel"; TDSA_DEBUG_PRINT=1 timeout -s KILL 300 python /tmp/welch_time.py
for lib in variants/libtdsa_vd*.so; do echo "== $lib"; TDSA_LIB=$PWD/$lib timeout -s KILL 300 python /tmp/welch_time.py; done
