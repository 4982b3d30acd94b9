#!/bin/bash
mkdir -p gpurun_out
echo "== default"; timeout 300 python tests/dev/gpu_dev.py quick 2>&1 | grep -E "^time N=(4096)|FAILED|Error" | cut -c 1-110
for lib in variants/libtdsa_*.so; do
  [ -f "$lib" ] || continue
  echo "== $lib"; TDSA_LIB=$PWD/$lib timeout 300 python tests/dev/gpu_dev.py quick 2>&1 | grep -E "^time N=(4096)|FAILED|Error" | cut -c 1-110
  TDSA_LIB=$PWD/$lib timeout 600 python -m pytest tests/test_gpu_wl_kernel.py tests/test_gpu_kernel1.py -m gpu -q -x 2>&1 | tail -2
done
timeout 1200 python -m pytest tests/test_gpu_state.py tests/test_gpu_full_size.py tests/test_gpu_next_rows.py -m gpu -q -x 2>&1 | tail -4
timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-other-sizes --min-seconds 0 > gpurun_out/bench_k.json 2> gpurun_out/bench_k.err; echo "bench rc=$?"; python - <<'PY'
import json
d = json.loads(open("gpurun_out/bench_k.json").read().strip().splitlines()[-1])
print({k: d[k] for k in ("value", "ms_per_step")}, d["roofline"]["frac"]); print(json.dumps(d["cfg4"])[:900])
PY
