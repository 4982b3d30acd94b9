#!/usr/bin/env python
"""Read a -DTDSA_DEBUG_TIMING dump and print, for the CTAs sharing one SM, how long each phase of a frame took.

stamps: 0 loop top, 1 stage landed, 2 pass-0 math done, 3 exchange-1 stores issued, 4 barrier passed,
5 pass-1 loads+math done, 6 exchange-2 stores issued, 7 barrier passed, 8 pass-2 loads+math done,
9 dB stores issued, 10 end-of-frame barrier passed."""
import sys
import numpy as np

NAMES = ["wait-stage", "p0 ld+math", "ex1 sts", "bar1", "p1 ld+math", "ex2 sts", "bar2", "p2 ld+math", "emit", "bar3"]


def main(path):
    raw = np.fromfile(path, dtype=np.int64)
    grid = int(path.rsplit("_g", 1)[1].split(".")[0])
    st = raw[: grid * 8 * 32 * 16].reshape(grid, 8, 32, 16)
    smid = raw[grid * 8 * 32 * 16: grid * 8 * 32 * 16 + grid]
    sm0 = smid[0]
    blocks = np.nonzero(smid == sm0)[0]
    print(f"{path}: grid {grid}, SM {sm0} hosts blocks {blocks.tolist()}")
    base = min(st[b, :, 0, 0].min() for b in blocks)
    for b in blocks:
        print(f" block {b}: per-frame period (warp 0, frames 4..20): "
              f"{np.diff(st[b, 0, 4:21, 0]).mean():.0f} cycles")
        d = np.diff(st[b, :, 4:20, 0:11], axis=-1)          # [warp, frame, phase]
        mean = d.mean(axis=(0, 1))
        print("   mean phase cycles: " + ", ".join(f"{n}={v:.0f}" for n, v in zip(NAMES, mean)) + f"  sum={mean.sum():.0f}")
    # timeline of frames 8..10 for warp 0 of each block, relative to base
    for b in blocks:
        for it in (8, 9):
            row = st[b, 0, it, 0:11] - base
            print(f"   block {b} warp0 frame {it}: " + " ".join(str(int(v)) for v in row))
    # spread between warps of one CTA at barrier arrival (stamp 3) and pass-2 end (stamp 8)
    b = blocks[0]
    for i in (2, 3, 5, 8):
        sp = st[b, :, 4:20, i].max(axis=0) - st[b, :, 4:20, i].min(axis=0)
        print(f"   block {b}: spread across warps at stamp {i}: mean {sp.mean():.0f} max {sp.max()}")


if __name__ == "__main__":
    for p in sys.argv[1:]:
        main(p)
