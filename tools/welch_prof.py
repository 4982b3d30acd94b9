import sys, os, torch
sys.path.insert(0, os.getcwd())
from topdogspectrumanalyser_b200 import synth
from topdogspectrumanalyser_b200.engine import SpectrumPlan
dev = torch.device("cuda:0")
stream = torch.from_numpy(synth.cfg3_stream(1 << 24, seed=2)).to(dev)
for prec in ("f64", "f32"):
    plan = SpectrumPlan(65536, precision=prec, device=dev)
    for _ in range(3):
        plan.welch(stream, 32768)
    torch.cuda.synchronize()
    plan.close()
