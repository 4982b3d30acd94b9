#!/bin/bash
# cfg3 fused kernel: phase time stamps
mkdir -p gpurun_out
export TDSA_LIB=$PWD/variants/libtdsa_timing.so TDSA_DEBUG_TIMING_OUT=$PWD/gpurun_out/timing_fused
for prec in f64 f32; do
WELCH_PROF_PREC=$prec timeout 300 python tools/welch_prof.py
python tools/phase_timing_fused.py gpurun_out/timing_fused_wl_${prec}_g288.bin
done
