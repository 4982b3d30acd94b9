#!/bin/bash
# ncu --set full of the four-pass classic kernel at N = 8192 (plain dB rows), float64 and float32
mkdir -p gpurun_out
cat > /tmp/p8192.py <<'PY'
import os, sys, torch
sys.path.insert(0, os.getcwd())
from topdogspectrumanalyser_b200 import synth
from topdogspectrumanalyser_b200.engine import SpectrumPlan
dev = torch.device("cuda:0")
base = torch.from_numpy(synth.cfg2_frames(b=512, n=8192, seed=1)).to(dev)
x = base.repeat(8, 1).contiguous()
out = torch.empty((4096, 8192), dtype=torch.float32, device=dev)
plan = SpectrumPlan(8192, precision=sys.argv[1], device=dev)
for _ in range(3): plan.psd_db(x, out=out)
torch.cuda.synchronize()
PY
for p in f64 f32; do
timeout -s KILL 600 ncu --set full --clock-control none --import-source on -k regex:fft_fused -s 2 -c 1 -f -o gpurun_out/r02_classic8192_$p python /tmp/p8192.py $p > gpurun_out/ncu_classic8192_$p.log 2>&1; echo "ncu $p rc=$?"
done
