#!/bin/bash
# round 2, record run 3 (after the bulk-copy head): GPU suite, smoke, bench both arms, configs, cfg3 ncu list + head capture,
# sanitizer on the cfg3 kernels
mkdir -p gpurun_out
rm -f gpurun_out/parity.jsonl
timeout 2400 python -m pytest tests -m gpu -q 2>&1 | tail -4
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/bench_z.json 2> gpurun_out/bench_z.err; echo "bench rc=$?"; python - <<'PY'
import json
d = json.loads(open("gpurun_out/bench_z.json").read().strip().splitlines()[-1])
print({k: d[k] for k in ("value", "ms_per_step", "gpu_launches")}, "frac", d["roofline"]["frac"], "e2e", d["e2e"]["value"], "sustained", d["sustained"]["roofline_frac"])
print("cfg3", json.dumps(d["cfg3"])[:700]); print("cfg4", json.dumps(d["cfg4"])[:400]); print("cpu", d.get("cpu_baseline", {}).get("value"), d["clocks"])
PY
timeout 600 python bench.py --impl reference --steps 5 --warmup 2 2>&1 | tail -1 | cut -c 1-200
timeout 900 python tools/configs_bench.py > gpurun_out/configs_z.jsonl 2> gpurun_out/configs_z.err; echo "configs rc=$?"; cut -c 1-250 gpurun_out/configs_z.jsonl | head -4
timeout -s KILL 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,lts__t_bytes.sum,sm__throughput.avg.pct_of_peak_sustained_elapsed,sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active,launch__registers_per_thread,launch__grid_size --clock-control none --csv --log-file gpurun_out/r02_cfg3_launches.csv python tools/welch_prof.py > gpurun_out/welch_prof.log 2>&1; echo "ncu cfg3 list rc=$?"
WELCH_PROF_PREC=f64 timeout -s KILL 600 ncu --set full --clock-control none --import-source on -k regex:big_head_wl -s 1 -c 1 -f -o gpurun_out/r02_cfg3_head_f64 python tools/welch_prof.py > gpurun_out/ncu_cfg3_head.log 2>&1; echo "ncu head rc=$?"
timeout -s KILL 900 compute-sanitizer --tool memcheck python tools/sanitize_cfg3.py > gpurun_out/r02_san_mem_cfg3.log 2>&1; echo "memcheck rc=$?"; grep -E "ERROR SUMMARY|ok cfg3" gpurun_out/r02_san_mem_cfg3.log
timeout -s KILL 1200 compute-sanitizer --tool racecheck --racecheck-report all python tools/sanitize_cfg3.py > gpurun_out/r02_san_race_cfg3.log 2>&1; echo "racecheck rc=$?"; grep -E "RACECHECK SUMMARY" gpurun_out/r02_san_race_cfg3.log
grep -E "^========= (Error|Warning):" gpurun_out/r02_san_race_cfg3.log | sed 's/at __shared__ 0x[0-9a-f]* in block ([0-9,]*)//' | sort | uniq -c | sort -rn | head -5
