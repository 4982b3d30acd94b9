#!/usr/bin/env python
"""Phase durations of fft_wl_kernel from a -DTDSA_DEBUG_TIMING dump (current stamp layout), even / odd iterations apart.

Stamps per warp and iteration: 0 top, 1 stage landed, 2 pass A done, 3 team transpose stored (+ claim), 4 pass B done,
5 Y stores issued, 6 Y barrier passed, 7 refill issued + last-pass loads done, 8 last pass done, 9 epilogue done."""
import sys
import numpy as np

NAMES = ["wait stage", "pass A", "transpose", "pass B", "Y sts", "Y barrier", "refill+L loads", "last pass", "epilogue", "to next top"]


def main(path, lo=4, hi=20):
    raw = np.fromfile(path, dtype=np.int64)
    grid = int(path.rsplit("_g", 1)[1].split(".")[0])
    st = raw[: grid * 8 * 32 * 16].reshape(grid, 8, 32, 16).astype(np.float64)
    smid = raw[grid * 8 * 32 * 16: grid * 8 * 32 * 16 + grid]
    its = (st[:, 0, :, 0] > 0).sum(axis=1)
    print(f"{path}: grid {grid}; stamped iterations per CTA: min {its.min()} median {np.median(its):.0f} max {its.max()}")
    for sm in (smid[0], smid[grid // 3]):
        blocks = np.nonzero(smid == sm)[0]
        print(f"SM {sm} hosts blocks {blocks.tolist()}")
        for b in blocks:
            n = int(its[b])
            h = min(hi, n - 1)
            if h <= lo + 2:
                print(f" block {b}: only {n} iterations"); continue
            for par, tag in ((0, "even"), (1, "odd ")):
                idx = [i for i in range(lo, h) if i % 2 == par]
                x = st[b][:, idx, :]
                nxt = st[b][:, [i + 1 for i in idx], 0]
                d = [x[..., k + 1] - x[..., k] for k in range(9)] + [nxt - x[..., 9]]
                tot = nxt - x[..., 0]
                print(f" block {b} {tag}: period {tot.mean():.0f}; " + ", ".join(f"{nm}={v.mean():.0f}" for nm, v in zip(NAMES, d)))


if __name__ == "__main__":
    main(sys.argv[1])
