#!/bin/bash
# synccheck over every kernel family; cfg5 with an in-place producer
mkdir -p gpurun_out
timeout -s KILL 1500 compute-sanitizer --tool synccheck python tools/sanitize.py > gpurun_out/r02_san_sync.log 2>&1; echo "synccheck rc=$?"; grep -E "ERROR SUMMARY|all ok" gpurun_out/r02_san_sync.log; grep -E "^========= (Error|Warning|Barrier error)" gpurun_out/r02_san_sync.log | sort | uniq -c | head -5
timeout -s KILL 1500 compute-sanitizer --tool initcheck python tools/sanitize_cfg3.py > gpurun_out/r02_san_init_cfg3.log 2>&1; echo "initcheck rc=$?"; grep -E "ERROR SUMMARY|ok cfg3" gpurun_out/r02_san_init_cfg3.log | tail -5
timeout 900 python tools/configs_bench.py 2>/dev/null | grep cfg5 | cut -c 1-420
