#!/bin/bash
# after switching the regions-free barrier to one arrival per thread: kernel tests + racecheck over the whole sanitizer workload
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_kernel1.py tests/test_gpu_round2.py tests/test_gpu_state.py -q -x 2>&1 | tail -2
timeout 300 python tests/dev/gpu_dev.py quick 2>&1 | grep -E "time N=4096"
timeout -s KILL 2400 compute-sanitizer --tool racecheck --racecheck-report all python tools/sanitize.py > gpurun_out/r02_san_race.log 2>&1; echo "racecheck rc=$?"; grep -E "RACECHECK SUMMARY|all ok" gpurun_out/r02_san_race.log; grep -E "hazard detected" gpurun_out/r02_san_race.log | sed 's/.*\(Potential [A-Z]* hazard detected[^.]*\).*/\1/' | sed 's/at __shared__.*//' | sort | uniq -c | head; grep -A6 "hazard detected" gpurun_out/r02_san_race.log | grep -E "at .*\(" | sed 's/ in \/.*//' | sed 's/.*at //' | cut -c 1-120 | sort | uniq -c | sort -rn | head -12
