#!/bin/bash
# 8-GPU checks: bench under torchrun (weak scaling), cfg-4 sweep sharded 38/37 sub-bands per rank with one NCCL all-gather
mkdir -p gpurun_out
N=${1:-8}
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus $N --steps 30 --warmup 5 > gpurun_out/bench_g$N.json 2> gpurun_out/bench_g$N.err; echo "bench g$N rc=$?"; cut -c 1-420 gpurun_out/bench_g$N.json; tail -2 gpurun_out/bench_g$N.err
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29522 tests/dev/sweep_demo.py --bands 300 --frames 16 --check 2>&1 | grep -E "^\{" | tee gpurun_out/sweep_g${N}_check.json
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29523 tests/dev/sweep_demo.py 2>&1 | grep -E "^\{" | tee gpurun_out/sweep_g$N.json
