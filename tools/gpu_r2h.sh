#!/bin/bash
# round 2, run H: ncu evidence. Launch list of the bench command; --set full captures of the headline kernel (both
# precisions) and of the accumulating-epilogue kernels.
mkdir -p gpurun_out
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/r02_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --e2e-steps 1 --min-seconds 0 --no-other-sizes > gpurun_out/ncu_bench.log 2>&1; echo "launch list rc=$?"
for p in f64 f32; do
  timeout -s KILL 600 ncu --set full --clock-control none --import-source on -k regex:fft_wl -s 3 -c 1 -f -o gpurun_out/r02_wl_$p python bench.py --precision $p --steps 2 --warmup 3 --no-cpu-baseline --e2e-steps 1 --min-seconds 0 --no-cfg4 --no-other-sizes > gpurun_out/ncu_wl_$p.log 2>&1; echo "ncu $p rc=$?"
done
timeout -s KILL 900 ncu --set full --clock-control none --import-source on -k regex:"fft_wl|trace_scan_dev|avg_finish|hold_finish|group_finish|fft_fused" -s 6 -c 7 -f -o gpurun_out/r02_acc_f64 python tools/prof_targets.py f64 > gpurun_out/ncu_acc_f64.log 2>&1; echo "ncu acc rc=$?"
ls -la gpurun_out/*.ncu-rep
