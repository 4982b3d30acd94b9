#!/usr/bin/env python
"""Phase durations of the fused cfg3 kernel (fft_wl_kernel, kAccFused) from a -DTDSA_DEBUG_TIMING dump.

Stamps per warp and iteration: 0 top, 10 fused start, 11 ring slot free + samples landed, 12 staging read, 13 head math
done, 14 head stores issued, 15 before the heads-done wait, 1 wait passed, 2 ring loads + pass A, 3 team transpose,
4 pass B, 5 Y stores, 6 Y barrier, 7 last-pass loads, 8 last pass, 9 accumulate."""
import sys
import numpy as np

ORDER = [0, 10, 11, 12, 13, 14, 15, 1, 2, 3, 4, 5, 6, 7, 8, 9]
NAMES = ["top", "wait slot+samples", "staging read", "head math", "head stores", "(gap)", "wait heads", "ring ld+pass A",
         "transpose", "pass B", "Y sts", "Y barrier", "L loads", "last pass", "accumulate"]


def main(path):
    raw = np.fromfile(path, dtype=np.int64)
    grid = int(path.rsplit("_g", 1)[1].split(".")[0])
    st = raw[: grid * 8 * 32 * 16].reshape(grid, 8, 32, 16).astype(np.float64)
    smid = raw[grid * 8 * 32 * 16: grid * 8 * 32 * 16 + grid]
    for sm in (smid[0], smid[grid // 2 + 3]):
        blocks = np.nonzero(smid == sm)[0]
        print(f"SM {sm} hosts blocks {blocks.tolist()}")
        for b in blocks:
            x = st[b, :, 4:20, :][..., ORDER]
            d = np.diff(x, axis=-1).mean(axis=(0, 1))
            period = np.diff(st[b, :, 4:21, 0], axis=1).mean()
            print(f" block {b}: period {period:.0f} cycles; " + ", ".join(f"{n}={v:.0f}" for n, v in zip(NAMES[1:], d)) + f"  sum={d.sum():.0f}")


if __name__ == "__main__":
    main(sys.argv[1])
