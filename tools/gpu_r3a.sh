#!/bin/bash
# ncu --set full of the two-engine N = 8192 group-mean kernel (config 4's rows), float64 and float32
mkdir -p gpurun_out
for p in f64 f32; do
timeout -s KILL 600 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k regex:"fft_wl_kernel.*\(int\)2, \(int\)9>" -s 1 -c 1 -f -o gpurun_out/r02_group8192_$p python tools/prof_targets.py $p > gpurun_out/ncu_group8192_$p.log 2>&1; echo "ncu $p rc=$?"
done
ls -la gpurun_out/r02_group8192_*.ncu-rep
