#!/usr/bin/env python
"""Read a -DTDSA_DEBUG_TIMING dump and print, for the CTAs sharing one SM, how long each phase of a frame took.

stamps: 0 loop top, 1 stage landed, 2 pass-0 math done, 3 exchange-1 stores issued, 4 barrier passed,
5 pass-1 loads+math done, 6 exchange-2 stores issued, 7 barrier passed, 8 pass-2 loads+math done,
9 dB stores issued, 10 end-of-frame barrier passed."""
import sys
import numpy as np

NAMES = ["wait-stage", "p0 ld+math", "ex1 sts", "bar1", "p1 ld+math", "ex2 sts", "bar2", "p2 ld+math", "emit", "bar3"]


def main_wl(path):
    """fft_wl_kernel stamps: S(k): 0 start, 1 stage landed, 2 pass A done, 3 re-arm + buffer-free wait + team stores
    done, 4 pass B done, 5 Y stores + arrive; L(k): 7 start, 6 Y-full wait passed, 8 loads + math done, 9 emit done."""
    raw = np.fromfile(path, dtype=np.int64)
    grid = int(path.rsplit("_g", 1)[1].split(".")[0])
    st = raw[: grid * 8 * 32 * 16].reshape(grid, 8, 32, 16)
    smid = raw[grid * 8 * 32 * 16: grid * 8 * 32 * 16 + grid]
    blocks = np.nonzero(smid == smid[0])[0]
    print(f"{path}: grid {grid}, SM {smid[0]} hosts blocks {blocks.tolist()}")
    fr = slice(4, 20)
    for b in blocks:
        x = st[b, :, fr, :].astype(np.float64)
        period = np.diff(st[b, :, 4:21, 0], axis=1).mean()
        seg = {"wait stage": x[..., 1] - x[..., 0], "pass A": x[..., 2] - x[..., 1], "arm+free-wait+team sts": x[..., 3] - x[..., 2],
               "pass B": x[..., 4] - x[..., 3], "Y sts": x[..., 5] - x[..., 4], "wait Y full": x[..., 6] - x[..., 7],
               "L ld+math": x[..., 8] - x[..., 6], "emit": x[..., 9] - x[..., 8]}
        print(f" block {b}: period {period:.0f} cycles/frame; " + ", ".join(f"{k}={v.mean():.0f}" for k, v in seg.items())
              + f"  sum={sum(v.mean() for v in seg.values()):.0f}")
        starts = st[b, :, 10, 0] - st[b, :, 10, 0].min()
        print(f"   warp skew at S(10) start: {starts.tolist()}")


def main_pp(path):
    """fft_wlpp_kernel: 0 top, 1 ready for A, 2 token, 3 A done, 4 ready for B, 5 token, 6 B done, 7 ready for L, 8 token,
    9 L done, 10 emit done.  Entry (block*2+g)."""
    raw = np.fromfile(path, dtype=np.int64)
    grid = int(path.rsplit("_g", 1)[1].split(".")[0])
    st = raw.reshape(grid * 2, 8, 32, 16).astype(np.float64)
    names = ["prep(stage,cvt)", "acq A", "math A", "N1(sts,lds)", "acq B", "math B", "N2(sts,bar,lds,bar)", "acq L", "math L", "emit"]
    for e in (0, 1):
        x = st[e, :, 4:20, :]
        d = np.diff(x[..., 0:11], axis=-1).mean(axis=(0, 1))
        per = np.diff(st[e, :, 4:21, 0], axis=1).mean()
        print(f" block 0 group {e}: period {per:.0f}; " + ", ".join(f"{n}={v:.0f}" for n, v in zip(names, d)) + f" sum={d.sum():.0f}")
    base = st[0, 0, 8, 0]
    for e in (0, 1):
        print(f"   group {e} warp0 frame 8: " + " ".join(str(int(v - base)) for v in st[e, 0, 8, 0:11]))
        print(f"   group {e} warp5 frame 8: " + " ".join(str(int(v - base)) for v in st[e, 5, 8, 0:11]))


def main(path):
    if "_pp_" in path:
        return main_pp(path)
    if "_wl_" in path:
        return main_wl(path)
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
