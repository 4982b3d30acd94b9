#!/bin/bash
mkdir -p gpurun_out
rm -f gpurun_out/parity.jsonl
timeout 600 python tools/acc_bench.py 2>&1 | tee gpurun_out/acc_bench.jsonl | cut -c 1-330
timeout 2400 python -m pytest tests -m gpu -q 2>&1 | tail -6
timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/bench_i.json 2> gpurun_out/bench_i.err; echo "bench rc=$?"; python - <<'PY'
import json
d = json.loads(open("gpurun_out/bench_i.json").read().strip().splitlines()[-1])
print({k: d[k] for k in ("value", "ms_per_step")}, d["roofline"]["frac"], d["e2e"]["value"], d["e2e"]["host_link_ceiling"]); print(json.dumps(d["cfg4"])[:900])
PY
