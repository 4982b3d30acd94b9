#!/bin/bash
# N = 8192 pair mode: phase time stamps
mkdir -p gpurun_out
export TDSA_LIB=$PWD/variants/libtdsa_timing.so TDSA_DEBUG_TIMING_OUT=$PWD/gpurun_out/timing_pair
cat > /tmp/tp.py <<'PY'
import os, sys, torch
sys.path.insert(0, os.getcwd())
from topdogspectrumanalyser_b200 import synth
from topdogspectrumanalyser_b200.engine import SpectrumPlan
dev = torch.device("cuda:0")
n = int(sys.argv[2])
base = torch.from_numpy(synth.cfg2_frames(b=512, n=n, seed=1)).to(dev)
x = base.repeat(33554432 // n // 512, 1).contiguous()
plan = SpectrumPlan(n, precision=sys.argv[1], device=dev)
for _ in range(3): plan.psd_db(x)
torch.cuda.synchronize()
PY
for prec in f64 f32; do
  python /tmp/tp.py $prec 8192; mv gpurun_out/timing_pair_wl_${prec}_g296.bin gpurun_out/timing_pair8192_wl_${prec}_g296.bin
  python tools/phase_timing_wl.py gpurun_out/timing_pair8192_wl_${prec}_g296.bin
done
