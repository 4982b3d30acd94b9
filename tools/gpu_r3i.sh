#!/bin/bash
# general trace path (averaging + hold + every row): scan chunk size
for mb in 16 32 64 96 128 256; do
  echo "== TDSA_SCAN_CHUNK_MB=$mb"
  TDSA_SCAN_CHUNK_MB=$mb timeout 300 python tools/acc_bench.py 2>&1 | grep "general path" | cut -c 1-60,100-175
done
