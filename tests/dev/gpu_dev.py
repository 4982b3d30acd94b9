#!/usr/bin/env python
"""Developer diagnostics on a GPU box: per-size error statistics and quick kernel timings."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np
import torch
from oracle import oracle as O
from topdogspectrumanalyser_b200 import synth, _lib
from topdogspectrumanalyser_b200.engine import SpectrumPlan

dev = torch.device("cuda:0")
print(torch.cuda.get_device_name(0), "SMs", torch.cuda.get_device_properties(0).multi_processor_count)

def stats(n, prec, b=4):
    iq = synth.cfg2_frames(b=b, n=n, seed=100 + n)
    want = O.power_db_batch(iq, O.make_window("hanning", n))
    try:
        plan = SpectrumPlan(n, "hanning", precision=prec, device=dev)
        got = plan.psd_db(torch.from_numpy(iq).to(dev)).cpu().numpy().astype(np.float64)
        torch.cuda.synchronize()
        info = plan.info()
        plan.close()
    except Exception as e:
        print(f"N={n:6d} {prec}: FAILED {e}")
        return
    err = np.abs(got - want)
    print(f"N={n:6d} {prec}: max={err.max():.3e} med={np.median(err):.3e} >1e-4: {(err>1e-4).sum()}/{err.size} "
          f"argmax={np.unravel_index(err.argmax(), err.shape)} info={info}")

sizes = [64, 128, 256, 512, 1024, 2048, 4096, 8192, 16384, 32768, 65536]
if len(sys.argv) > 1 and sys.argv[1] == "quick":
    sizes = [1024, 4096, 8192]
for n in sizes:
    for prec in ("f64", "f32"):
        stats(n, prec)

def timeit(n, b, prec, reps=20):
    x = torch.from_numpy(synth.cfg2_frames(b=min(b, 1024), n=n, seed=1)).to(dev)
    if b > x.shape[0]:
        x = x.repeat(b // x.shape[0], 1).contiguous()
    plan = SpectrumPlan(n, "hanning", precision=prec, device=dev)
    out = torch.empty((b, n), dtype=torch.float32, device=dev)
    for _ in range(3):
        plan.psd_db(x, out=out)
    torch.cuda.synchronize()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(reps + 1)]
    ev[0].record()
    for i in range(reps):
        plan.psd_db(x, out=out)
        ev[i + 1].record()
    torch.cuda.synchronize()
    ts = np.array([ev[i].elapsed_time(ev[i + 1]) for i in range(reps)]) * 1e-3
    t = np.median(ts)
    gbs = 12.0 * b * n / t / 1e9
    print(f"time N={n} B={b} {prec}: median {t*1e6:.1f} us  best {ts.min()*1e6:.1f} us  {b*n/t/1e9:.1f} Gsamples/s  "
          f"{gbs:.0f} GB/s ({gbs/6534.1*100:.1f}% of measured HBM) info={plan.info()}")
    plan.close()

for prec in ("f32", "f64"):
    timeit(4096, 8192, prec)
    timeit(1024, 32768, prec)
    timeit(8192, 4096, prec)
timeit(65536, 512, "f32")
timeit(65536, 512, "f64")
print("launches", _lib.launch_count())
