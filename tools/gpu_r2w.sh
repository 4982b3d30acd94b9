#!/bin/bash
# round 2, record run 2: whole GPU test-suite, smoke, bench (both arms), configs 3/4/5, trace-path timings, cfg3 per-kernel
# ncu list + full captures of its head and tail kernels, launch list of the bench
mkdir -p gpurun_out
rm -f gpurun_out/parity.jsonl
timeout 2400 python -m pytest tests -m gpu -q 2>&1 | tail -5
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -4
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/bench_w.json 2> gpurun_out/bench_w.err; echo "bench rc=$?"; python - <<'PY'
import json
d = json.loads(open("gpurun_out/bench_w.json").read().strip().splitlines()[-1])
print({k: d[k] for k in ("value", "ms_per_step", "gpu_launches")}, "frac", d["roofline"]["frac"], "e2e", d["e2e"]["value"], "sustained", d["sustained"]["roofline_frac"])
print("other", json.dumps(d["other_sizes"])[:600]); print("cfg3", json.dumps(d["cfg3"])[:700]); print("cfg4", json.dumps(d["cfg4"])[:500]); print("cpu", d.get("cpu_baseline", {}).get("value"), d["clocks"])
PY
timeout 600 python bench.py --impl reference --steps 5 --warmup 2 2>&1 | tail -1 | cut -c 1-300
timeout 900 python tools/configs_bench.py > gpurun_out/configs_w.jsonl 2> gpurun_out/configs_w.err; echo "configs rc=$?"; cut -c 1-300 gpurun_out/configs_w.jsonl
timeout 600 python tools/acc_bench.py 2>&1 | tee gpurun_out/acc_bench_w.jsonl | cut -c 1-200
timeout -s KILL 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,lts__t_bytes.sum,sm__throughput.avg.pct_of_peak_sustained_elapsed,sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active,launch__registers_per_thread,launch__grid_size --clock-control none --csv --log-file gpurun_out/r02_cfg3_launches.csv python tools/welch_prof.py > gpurun_out/welch_prof.log 2>&1; echo "ncu cfg3 list rc=$?"
WELCH_PROF_PREC=f64 timeout -s KILL 600 ncu --set full --clock-control none --import-source on -k regex:big_head_wl -s 1 -c 1 -f -o gpurun_out/r02_cfg3_head_f64 python tools/welch_prof.py > gpurun_out/ncu_cfg3_head.log 2>&1; echo "ncu head rc=$?"
WELCH_PROF_PREC=f64 timeout -s KILL 600 ncu --set full --clock-control none --import-source on -k regex:fft_wl -s 1 -c 1 -f -o gpurun_out/r02_cfg3_tail_f64 python tools/welch_prof.py > gpurun_out/ncu_cfg3_tail.log 2>&1; echo "ncu tail rc=$?"
timeout -s KILL 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02_launches.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --min-seconds 0 > gpurun_out/bench_under_ncu.log 2>&1; echo "ncu launch list rc=$?"
