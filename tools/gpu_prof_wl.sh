#!/bin/bash
# ncu --set full capture of the warp-local kernel, both precisions
mkdir -p gpurun_out
for p in f64 f32; do
  timeout -s KILL 600 ncu --set full --clock-control none --import-source on -k regex:fft_wl -s 3 -c 1 -f -o gpurun_out/prof_wl_$p python bench.py --precision $p --steps 2 --warmup 3 --no-cpu-baseline --e2e-steps 1 > gpurun_out/ncu_wl_$p.log 2>&1; echo "ncu $p rc=$?"
done
ls -la gpurun_out
