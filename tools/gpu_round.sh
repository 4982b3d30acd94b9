#!/bin/bash
# One GPU-box session for the record: tests, bench, ncu launch list, ncu full capture of the fused kernel (both precisions).
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -5
TDSA_LOGR_F32=3 TDSA_LOGR_F64=3 timeout 900 python -m pytest tests/test_gpu_kernel1.py -m gpu -x -q 2>&1 | tail -2
timeout 600 python bench.py --steps 50 --warmup 5 > gpurun_out/bench.json 2> gpurun_out/bench_err.log; echo "bench rc=$?"; cat gpurun_out/bench.json; tail -5 gpurun_out/bench_err.log
timeout 600 python bench.py --impl reference --steps 10 --warmup 3 > gpurun_out/bench_ref.json 2>> gpurun_out/bench_err.log; cut -c 1-400 gpurun_out/bench_ref.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 80 --csv --log-file gpurun_out/launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --e2e-steps 1 > gpurun_out/ncu_bench.log 2>&1; echo "ncu list rc=$?"
for p in f64 f32; do
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:fft_ -s 3 -c 1 -f -o gpurun_out/prof_$p python bench.py --precision $p --steps 2 --warmup 3 --no-cpu-baseline --e2e-steps 1 > gpurun_out/ncu_$p.log 2>&1; echo "ncu $p rc=$?"
done
