#!/bin/bash
# A/B: default library and every variants/libtdsa_v*.so: N = 4096 error statistics + timings, then the kernel tests
mkdir -p gpurun_out
echo "== default"; timeout -s KILL 120 python tests/dev/gpu_dev.py quick 2>&1 | grep -E "^N=  4096|^time N=(4096)|FAILED|Error|error" | cut -c 1-200
for lib in variants/libtdsa_v*.so; do
  [ -f "$lib" ] || continue
  echo "== $lib"; TDSA_LIB=$PWD/$lib timeout -s KILL 120 python tests/dev/gpu_dev.py quick 2>&1 | grep -E "^N=  4096|^time N=(4096)|FAILED|Error|error" | cut -c 1-200
  TDSA_LIB=$PWD/$lib timeout -s KILL 600 python -m pytest tests/test_gpu_kernel1.py tests/test_gpu_wl_kernel.py -m gpu -x -q --deselect tests/test_gpu_wl_kernel.py::test_launch_geometry_is_the_warp_local_kernel 2>&1 | tail -2
done
timeout -s KILL 600 python -m pytest tests/test_gpu_kernel1.py tests/test_gpu_wl_kernel.py -m gpu -x -q 2>&1 | tail -2
