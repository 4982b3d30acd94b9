#!/bin/bash
# cfg3: parity (few segments, psd), timing, per-kernel breakdown; head staging by bulk copies against 8-byte cp.async
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_full_size.py -q -k "few_segments or psd_blackman" -x 2>&1 | tail -3
cat > /tmp/cfg3_time.py <<'PY'
import os, sys, numpy as np, torch
sys.path.insert(0, os.getcwd())
from topdogspectrumanalyser_b200 import synth
from topdogspectrumanalyser_b200.engine import SpectrumPlan
dev = torch.device("cuda:0")
stream = torch.from_numpy(synth.cfg3_stream(1 << 26, seed=2)).to(dev)
for prec in ("f64", "f32"):
    plan = SpectrumPlan(65536, precision=prec, device=dev)
    for _ in range(3): out = plan.welch(stream, 32768)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10): plan.welch(stream, 32768)
    e1.record(); torch.cuda.synchronize()
    print("cfg3 TDSA_HEAD_BULK=%s" % os.environ.get("TDSA_HEAD_BULK", "1"), prec, "ms %.4f" % (e0.elapsed_time(e1) / 10), "nan", bool(torch.isnan(out[0]).any()), flush=True)
    plan.close()
PY
timeout 300 python /tmp/cfg3_time.py
TDSA_HEAD_BULK=0 timeout 300 python /tmp/cfg3_time.py
timeout -s KILL 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,lts__t_bytes.sum,sm__throughput.avg.pct_of_peak_sustained_elapsed,sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active,launch__registers_per_thread,launch__grid_size --clock-control none --csv --log-file gpurun_out/r02_cfg3_launches.csv python tools/welch_prof.py > gpurun_out/welch_prof.log 2>&1; echo "ncu rc=$?"
python - <<'PY'
import csv
rows = list(csv.reader(l for l in open("gpurun_out/r02_cfg3_launches.csv") if l.startswith('"')))
h = rows[0]; iK = h.index("Kernel Name"); iM = h.index("Metric Name"); iV = h.index("Metric Value"); iI = h.index("ID")
cur = {}
for r in rows[1:]:
    cur.setdefault((r[iI], r[iK][:40]), {})[r[iM]] = r[iV]
seen = set()
for (i, k), m in cur.items():
    if k in seen: continue
    seen.add(k); print(i, k, {a.split("__")[-1][:14]: b for a, b in m.items()})
PY
