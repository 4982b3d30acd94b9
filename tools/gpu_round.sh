#!/bin/bash
# One GPU-box session: tests, bench, ncu launch list, ncu full capture of the fused kernel (both precisions).
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -15
timeout 600 python bench.py --steps 50 --warmup 5 > gpurun_out/bench.json 2> gpurun_out/bench_err.log; echo "bench rc=$?"; cat gpurun_out/bench.json; tail -5 gpurun_out/bench_err.log
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 80 --csv --log-file gpurun_out/launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --e2e-steps 1 > gpurun_out/ncu_bench.log 2>&1; echo "ncu list rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:fft_fused -s 3 -c 2 -f -o gpurun_out/prof_f64 python bench.py --steps 2 --warmup 3 --no-cpu-baseline --e2e-steps 1 > gpurun_out/ncu_f64.log 2>&1; echo "ncu f64 rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:fft_fused -s 3 -c 2 -f -o gpurun_out/prof_f32 python bench.py --precision f32 --steps 2 --warmup 3 --no-cpu-baseline --e2e-steps 1 > gpurun_out/ncu_f32.log 2>&1; echo "ncu f32 rc=$?"
ls -la gpurun_out
