#!/bin/bash
# round 2, record run: whole GPU test-suite, smoke, bench (both arms), configs 3/4/5, trace-path timings, ncu of the headline kernel
mkdir -p gpurun_out
rm -f gpurun_out/parity.jsonl
timeout 2400 python -m pytest tests -m gpu -q 2>&1 | tail -5
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -4
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/bench_o.json 2> gpurun_out/bench_o.err; echo "bench rc=$?"; python - <<'PY'
import json
d = json.loads(open("gpurun_out/bench_o.json").read().strip().splitlines()[-1])
print({k: d[k] for k in ("value", "ms_per_step", "gpu_launches")}, "frac", d["roofline"]["frac"], "e2e", d["e2e"]["value"], "sustained", d["sustained"]["roofline_frac"])
print("other", json.dumps(d["other_sizes"])[:600]); print("cfg4", json.dumps(d["cfg4"])[:700]); print("cpu", d.get("cpu_baseline", {}).get("value"), d["clocks"])
PY
timeout 600 python bench.py --impl reference --steps 5 --warmup 2 2>&1 | tail -1 | cut -c 1-300
timeout 900 python tools/configs_bench.py > gpurun_out/configs_o.jsonl 2> gpurun_out/configs_o.err; echo "configs rc=$?"; cut -c 1-260 gpurun_out/configs_o.jsonl
timeout 600 python tools/acc_bench.py 2>&1 | tee gpurun_out/acc_bench_o.jsonl | cut -c 1-200
timeout -s KILL 600 ncu --set full --clock-control none --import-source on -k regex:fft_wl -s 3 -c 1 -f -o gpurun_out/r02_wl_f64_final python bench.py --precision f64 --steps 2 --warmup 3 --no-cpu-baseline --e2e-steps 1 --min-seconds 0 --no-cfg4 --no-cfg3 --no-other-sizes > gpurun_out/ncu_wl_f64_final.log 2>&1; echo "ncu rc=$?"
