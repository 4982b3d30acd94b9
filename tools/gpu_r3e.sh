#!/bin/bash
# round 2, record run 4 (final library): GPU suite, smoke, bench both arms, configs, trace paths, launch list, captures of the
# two-engine kernel (group mean and rows) after the tensor-memory window
mkdir -p gpurun_out
rm -f gpurun_out/parity.jsonl
timeout 2400 python -m pytest tests -m gpu -q 2>&1 | tail -3
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/bench_e.json 2> gpurun_out/bench_e.err; echo "bench rc=$?"; python - <<'PY'
import json
d = json.loads(open("gpurun_out/bench_e.json").read().strip().splitlines()[-1])
print({k: d[k] for k in ("value", "ms_per_step", "gpu_launches")}, "frac", d["roofline"]["frac"], "e2e", d["e2e"]["value"], "sustained", d["sustained"]["roofline_frac"])
print("other", json.dumps(d["other_sizes"])[:700]); print("cfg3", json.dumps(d["cfg3"])[:500]); print("cfg4", json.dumps(d["cfg4"])[:500]); print("cpu", d.get("cpu_baseline", {}).get("value"), d["clocks"])
PY
timeout 600 python bench.py --impl reference --steps 5 --warmup 2 > gpurun_out/bench_e_reference.json 2>/dev/null; tail -1 gpurun_out/bench_e_reference.json | cut -c 1-200
timeout 900 python tools/configs_bench.py > gpurun_out/configs_e.jsonl 2> gpurun_out/configs_e.err; echo "configs rc=$?"; grep -E "default|cfg4|cfg5" gpurun_out/configs_e.jsonl | cut -c 1-230
timeout 600 python tools/acc_bench.py 2>&1 | tee gpurun_out/acc_bench_e.jsonl | cut -c 1-160
timeout -s KILL 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02_launches.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --min-seconds 0 > gpurun_out/bench_under_ncu.log 2>&1; echo "ncu launch list rc=$?"
timeout -s KILL 600 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k regex:"fft_wl_kernel.*\(int\)2, \(int\)9>" -s 1 -c 1 -f -o gpurun_out/r02_group8192_f64 python tools/prof_targets.py f64 > gpurun_out/ncu_group8192_f64.log 2>&1; echo "ncu group rc=$?"
