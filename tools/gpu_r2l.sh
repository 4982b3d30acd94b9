#!/bin/bash
# sanitizer on every kernel family (memcheck; racecheck on shared memory)
mkdir -p gpurun_out
timeout -s KILL 1500 compute-sanitizer --tool memcheck python tools/sanitize.py > gpurun_out/r02_san_mem.log 2>&1; echo "memcheck rc=$?"; grep -E "ERROR SUMMARY|all ok|ok round-2" gpurun_out/r02_san_mem.log
timeout -s KILL 2400 compute-sanitizer --tool racecheck --racecheck-report all python tools/sanitize.py > gpurun_out/r02_san_race.log 2>&1; echo "racecheck rc=$?"; grep -E "RACECHECK SUMMARY|all ok" gpurun_out/r02_san_race.log; grep -E "hazard detected" gpurun_out/r02_san_race.log | sed 's/.*\(Potential [A-Z]* hazard detected[^.]*\).*/\1/' | sort | uniq -c | head; grep -A6 "hazard detected" gpurun_out/r02_san_race.log | grep -E "at .*\(" | sed 's/ in \/.*//' | sort | uniq -c | sort -rn | head -12
