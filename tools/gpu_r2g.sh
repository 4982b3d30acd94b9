#!/bin/bash
# round 2, run G (8 GPUs): bench under torchrun: weak scaling of the headline, e2e with host placement, config 4 with the fused exchange
mkdir -p gpurun_out
N=${1:-8}
nvidia-smi topo -m 2>/dev/null | head -14
lscpu | grep -E "NUMA|Socket|Model name|^CPU\(s\)" | head -8
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus $N --steps 20 --warmup 5 > gpurun_out/bench_g$N.json 2> gpurun_out/bench_g$N.err; echo "bench g$N rc=$?"; python - <<PY
import json
try:
    d = json.loads(open("gpurun_out/bench_g$N.json").read().strip().splitlines()[-1])
    print({k: d[k] for k in ("value", "ms_per_step", "n_gpus")}); print("e2e", d["e2e"]); print("cfg4", json.dumps(d["cfg4"])[:1800])
except Exception as e:
    print("parse failed", e); print(open("gpurun_out/bench_g$N.json").read()[-1500:])
PY
tail -5 gpurun_out/bench_g$N.err
