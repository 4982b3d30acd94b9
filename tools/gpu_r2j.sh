#!/bin/bash
mkdir -p gpurun_out
rm -f gpurun_out/parity.jsonl
timeout 1200 python -m pytest tests/test_gpu_round2.py tests/test_gpu_state.py tests/test_gpu_full_size.py -m gpu -q -x 2>&1 | tail -6
timeout 600 python tools/acc_bench.py 2>&1 | tee gpurun_out/acc_bench.jsonl | cut -c 1-250
timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-other-sizes --min-seconds 0 > gpurun_out/bench_j.json 2> gpurun_out/bench_j.err; echo "bench rc=$?"; python - <<'PY'
import json
d = json.loads(open("gpurun_out/bench_j.json").read().strip().splitlines()[-1])
print({k: d[k] for k in ("value", "ms_per_step")}, d["roofline"]["frac"]); print(json.dumps(d["cfg4"])[:900])
PY
