#!/usr/bin/env python
"""Config 4 end to end on N GPUs (torchrun): shard sub-bands, kernel 1 + group average, NCCL all-gather, stitch.

  python -m torch.distributed.run --nproc-per-node N --master-addr 127.0.0.1 tests/dev/sweep_demo.py [--check]
"""
import argparse, os, sys, time, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np
import torch
import torch.distributed as dist
from topdogspectrumanalyser_b200 import synth
from topdogspectrumanalyser_b200.sweep import WidebandSweep, shard_bands

ap = argparse.ArgumentParser()
ap.add_argument("--bands", type=int, default=300)
ap.add_argument("--frames", type=int, default=16)
ap.add_argument("--n", type=int, default=8192)
ap.add_argument("--reps", type=int, default=10)
ap.add_argument("--precision", default="f64")
ap.add_argument("--check", action="store_true", help="compare rows and grid with the oracle on rank 0 (slow)")
args = ap.parse_args()

world = int(os.environ.get("WORLD_SIZE", "1")); rank = int(os.environ.get("RANK", "0")); lr = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(lr); dev = torch.device("cuda", lr)
if world > 1:
    dist.init_process_group("nccl", device_id=dev)
mine = shard_bands(args.bands, world, rank)
iq = torch.from_numpy(synth.cfg4_subbands(args.bands, args.frames, args.n, seed=3, bands=mine)).to(dev)
sw = WidebandSweep(args.bands, 20e6, args.n, 0.0, precision=args.precision, device=dev)
rows, grid = sw.run(iq)
torch.cuda.synchronize()
if world > 1: dist.barrier()
ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
ev0.record()
for _ in range(args.reps):
    rows, grid = sw.run(iq)
ev1.record(); torch.cuda.synchronize()
ms = torch.tensor([ev0.elapsed_time(ev1) / args.reps], device=dev, dtype=torch.float64)
if world > 1: dist.all_reduce(ms, op=dist.ReduceOp.MAX)
if rank == 0:
    samples = args.bands * args.frames * args.n
    out = {"config": "cfg4 wideband stitch", "n_gpus": world, "bands": args.bands, "frames": args.frames, "n_fft": args.n,
           "ms_per_sweep": float(ms.item()), "samples_per_s": samples / (float(ms.item()) * 1e-3), "grid_bins": int(grid.numel()),
           "rows_nan": int(torch.isnan(rows).sum().item())}
    if args.check:
        from oracle import oracle as O
        full = synth.cfg4_subbands(args.bands, args.frames, args.n, seed=3)
        w = O.make_window("hanning", args.n)
        want = np.stack([10 * np.log10(O.linear_power_batch(b, w).mean(axis=0) + 1e-10) for b in full])
        got = rows.cpu().numpy()
        out["rows_max_err_db"] = float(np.abs(got - want).max())
        los = [20e6 * i for i in range(args.bands)]
        g = O.stitch_rows(got, los, [l + 20e6 for l in los], O.sweep_grid(0, int(args.bands * 20e6), 20e6 / args.n))
        out["grid_equal"] = bool(np.array_equal(g, grid.cpu().numpy()))
    print(json.dumps(out))
if world > 1:
    dist.barrier(); dist.destroy_process_group()
