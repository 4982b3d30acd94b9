#!/bin/bash
# round 2, run F (2 GPUs): the two-rank exchange tests (NCCL all-gather and the fused peer-store path), bench under torchrun
mkdir -p gpurun_out
nvidia-smi -L
timeout 900 python -m pytest tests/test_gpu_full_size.py -m gpu -q -k two_rank 2>&1 | tail -15
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29531 bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/bench_g2.json 2> gpurun_out/bench_g2.err; echo "bench g2 rc=$?"; python - <<'PY'
import json
try:
    d = json.loads(open("gpurun_out/bench_g2.json").read().strip().splitlines()[-1])
    print({k: d[k] for k in ("value", "ms_per_step", "n_gpus")}); print("e2e", d["e2e"]); print("cfg4", json.dumps(d["cfg4"])[:1800])
except Exception as e:
    print("parse failed", e); print(open("gpurun_out/bench_g2.json").read()[-1500:])
PY
tail -5 gpurun_out/bench_g2.err
