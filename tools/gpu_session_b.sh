#!/bin/bash
# sanitizer on the new kernel paths, configs 3/4/5 at full size, quick look at 2048/8192 timings
mkdir -p gpurun_out
timeout -s KILL 900 compute-sanitizer --tool memcheck python tools/sanitize.py > gpurun_out/san_mem.log 2>&1; echo "memcheck rc=$?"; grep -E "ERROR SUMMARY|all ok" gpurun_out/san_mem.log
timeout -s KILL 1200 compute-sanitizer --tool racecheck --racecheck-report all python tools/sanitize.py > gpurun_out/san_race.log 2>&1; echo "racecheck rc=$?"; grep -E "RACECHECK SUMMARY|all ok" gpurun_out/san_race.log; grep -E "hazard detected" gpurun_out/san_race.log | sed 's/.*\(Potential [A-Z]* hazard detected[^.]*\).*/\1/' | sort | uniq -c | head; grep -A6 "hazard detected" gpurun_out/san_race.log | grep -E "at .*\(" | sed 's/ in \/.*//' | sort | uniq -c | sort -rn | head -12
timeout -s KILL 900 python tools/configs_bench.py > gpurun_out/configs.jsonl 2> gpurun_out/configs.err; echo "configs rc=$?"; cut -c 1-330 gpurun_out/configs.jsonl; tail -3 gpurun_out/configs.err
