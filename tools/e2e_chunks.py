#!/usr/bin/env python
"""e2e (pinned host in -> pinned host out) time of config 2 for several chunk sizes."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from topdogspectrumanalyser_b200 import synth
from topdogspectrumanalyser_b200.engine import SpectrumPlan
dev = torch.device("cuda:0")
iq_t = torch.empty((8192, 4096), dtype=torch.complex64).pin_memory(); iq = iq_t.numpy()
iq[:] = np.tile(synth.cfg2_frames(b=1024, n=4096, seed=1), (8, 1))
out_t = torch.empty((8192, 4096), dtype=torch.float32).pin_memory(); out = out_t.numpy()
plan = SpectrumPlan(4096, device=dev)
for chunk in (0, 2048, 1024, 512, 256, 128):
    plan.psd_db_host(iq, out, chunk_frames=chunk)
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(5): plan.psd_db_host(iq, out, chunk_frames=chunk)
    torch.cuda.synchronize(); dt = (time.perf_counter() - t0) / 5
    print(f"chunk_frames={chunk:5d}: {dt*1e3:.3f} ms/step  {8192*4096/dt/1e9:.2f} Gsamples/s  H2D {8192*4096*8/dt/1e9:.1f} GB/s")
