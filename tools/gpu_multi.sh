#!/bin/bash
# 2-GPU checks: bench under torchrun, cfg-4 sweep with NCCL all-gather (checked against the oracle)
mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 30 --warmup 5 > gpurun_out/bench_g2.json 2> gpurun_out/bench_g2.err; echo "bench g2 rc=$?"; tail -c 1500 gpurun_out/bench_g2.json; tail -3 gpurun_out/bench_g2.err
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 tests/dev/sweep_demo.py --bands 30 --frames 4 --check 2>&1 | tail -3
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 tests/dev/sweep_demo.py 2>&1 | tail -2
timeout 300 python tests/dev/sweep_demo.py 2>&1 | tail -1
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29514 bench.py --impl reference --gpus 2 --steps 3 --warmup 1 2>&1 | tail -2 | cut -c 1-300
