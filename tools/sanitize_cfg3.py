#!/usr/bin/env python
"""compute-sanitizer workload for the config-3 kernels only (head + tails, fused variant), few segments."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from topdogspectrumanalyser_b200 import synth
from topdogspectrumanalyser_b200.engine import SpectrumPlan
dev = torch.device("cuda:0")
s3 = torch.from_numpy(synth.cfg3_stream(65536 + 32768 * 20)).to(dev)
for env in ({}, {"TDSA_WELCH_FUSED": "1"}):
    os.environ.pop("TDSA_WELCH_FUSED", None)
    os.environ.update(env)
    for prec in (sys.argv[1:] or ["f64", "f32"]):
        plan = SpectrumPlan(65536, precision=prec, device=dev)
        a, pk = plan.welch(s3, 32768)
        torch.cuda.synchronize(); plan.close()
        print("ok cfg3", env, prec, float(a[0]), bool(torch.isfinite(a).all()))
