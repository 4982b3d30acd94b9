#!/bin/bash
# racecheck of the cfg3 kernels with one regions-free arrival per THREAD (variant) -- is the WAR report an artefact of the
# per-warp arrival?  plus the variant's speed on the headline kernel
mkdir -p gpurun_out
export TDSA_LIB=$PWD/variants/libtdsa_splitall.so
timeout -s KILL 1200 compute-sanitizer --tool racecheck --racecheck-report all --racecheck-detect-level error --print-limit 40 python tools/sanitize_cfg3.py f64 > gpurun_out/r02_san_race_cfg3_splitall.log 2>&1; echo "racecheck rc=$?"
grep -E "RACECHECK SUMMARY|ok cfg3" gpurun_out/r02_san_race_cfg3_splitall.log
grep -E "^========= (Error|Warning):" gpurun_out/r02_san_race_cfg3_splitall.log | sed 's/at __shared__ 0x[0-9a-f]* in block ([0-9,]*)//' | sort | uniq -c | sort -rn | head
grep -A3 "^========= Error" gpurun_out/r02_san_race_cfg3_splitall.log | head -24
timeout 300 python tests/dev/gpu_dev.py quick 2>&1 | grep -E "time N=4096"
