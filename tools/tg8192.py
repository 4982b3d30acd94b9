import os, sys, torch
sys.path.insert(0, os.getcwd())
from topdogspectrumanalyser_b200 import synth
from topdogspectrumanalyser_b200.engine import SpectrumPlan
dev = torch.device("cuda:0")
def ev_time(fn, reps=20):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps * 1e3
g = torch.from_numpy(synth.cfg4_subbands(300, 16, 8192, seed=3)).to(dev)
for prec in ("f64", "f32"):
    plan = SpectrumPlan(8192, precision=prec, device=dev)
    print(prec, "group mean 300x16x8192: %.1f us" % ev_time(lambda: plan.group_avg_db(g)), "| 38x16: %.1f us" % ev_time(lambda: plan.group_avg_db(g[:38])), flush=True)
    plan.close()
