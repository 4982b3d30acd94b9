#!/bin/bash
# quick correctness + timing, then one ncu --set full capture per precision of the fused kernel
mkdir -p gpurun_out
timeout 300 python tests/dev/gpu_dev.py quick > gpurun_out/dev.log 2>&1; echo "dev rc=$?"; grep -E "^N=|^time" gpurun_out/dev.log
for p in f64 f32; do
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:fft_fused -s 3 -c 1 -f -o gpurun_out/prof_$p python bench.py --precision $p --steps 2 --warmup 3 --no-cpu-baseline --e2e-steps 1 > gpurun_out/ncu_$p.log 2>&1; echo "ncu $p rc=$?"
done
