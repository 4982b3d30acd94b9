#!/bin/bash
# general trace path: transform of chunk i+1 on the side stream under the scan of chunk i
timeout 900 python -m pytest tests/test_gpu_round2.py tests/test_gpu_state.py tests/test_gpu_next_rows.py -q -x 2>&1 | tail -2
for pl in 1 0; do
  echo "== TDSA_SCAN_PIPELINE=$pl"
  TDSA_SCAN_PIPELINE=$pl timeout 300 python tools/acc_bench.py 2>&1 | grep "general path" | cut -c 1-60,100-175
done
for mb in 32 96; do
  echo "== pipelined, TDSA_SCAN_CHUNK_MB=$mb"
  TDSA_SCAN_CHUNK_MB=$mb timeout 300 python tools/acc_bench.py 2>&1 | grep "general path" | cut -c 1-60,100-175
done
