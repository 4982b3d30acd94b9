"""cfg3 (65536-point Welch over 2^26 samples) three times per precision, for ncu launch lists / full captures."""
import sys, os, torch
sys.path.insert(0, os.getcwd())
from topdogspectrumanalyser_b200 import synth
from topdogspectrumanalyser_b200.engine import SpectrumPlan
dev = torch.device("cuda:0")
n_samples = int(os.environ.get("WELCH_PROF_SAMPLES", 1 << 26))
stream = torch.from_numpy(synth.cfg3_stream(n_samples, seed=2)).to(dev)
for prec in os.environ.get("WELCH_PROF_PREC", "f64,f32").split(","):
    plan = SpectrumPlan(65536, precision=prec, device=dev)
    for _ in range(3):
        plan.welch(stream, 32768)
    torch.cuda.synchronize()
    plan.close()
