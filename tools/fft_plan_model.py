#!/usr/bin/env python
"""Numpy model of the kernel's in-place DIF index algebra + shared-memory bank check.

Not product code: a design aid that mirrors csrc/tdsa_fft_kernel.cuh so the
decomposition (radix plan, thread->butterfly mapping, digit reversal of the
final pass, padded exchange layout) can be verified on the CPU, and so the
padding table in the kernel can be brute-force checked for bank conflicts.

Layout: 16 points per thread, T = N/16 threads per frame.  Passes are radix 16
until fewer than 4 bits remain, then one final pass of radix 2/4/8 (or 16).
"""
import itertools
import sys

import numpy as np


LOGR = 4          # digit width: 4 = radix 16 / 16 points per thread, 3 = radix 8 / 8 points per thread


def plan(n):
    lg = n.bit_length() - 1
    radices = [1 << LOGR] * (lg // LOGR)
    if lg % LOGR:
        radices.append(1 << (lg % LOGR))
    return radices


def hexrev(v, digits):
    out = 0
    for _ in range(digits):
        out = (out << LOGR) | (v & ((1 << LOGR) - 1))
        v >>= LOGR
    return out


def phys(p, pads):
    return p + sum(c * (p >> a) for a, c in pads)


def accesses(n):
    """Yield (pass_index, kind, lane_positions[T]) for each register slot j of each pass."""
    radices = plan(n)
    t_count = n >> LOGR
    t = np.arange(t_count)
    length = n
    m = len(radices)
    for i, r in enumerate(radices):
        s_i = length // r
        nb = (1 << LOGR) // r
        last = i == m - 1
        for u in range(nb):
            b = t + t_count * u
            if last and m > 1:
                digits = m - 1
                s = np.array([hexrev(int(v), digits) for v in b])
                base = s * r
                stride = 1
            else:
                c = b % s_i
                s = b // s_i
                base = s * length + c
                stride = s_i
            for j in range(r):
                yield i, u, j, base + j * stride
        length //= r


def model_fft(x, window=None):
    """Run the exact pass structure on one frame; returns fftshifted spectrum order check."""
    n = len(x)
    radices = plan(n)
    buf = x.astype(np.complex128).copy()
    length = n
    m = len(radices)
    t_count = n >> LOGR
    out = np.zeros(n, dtype=np.complex128)
    for i, r in enumerate(radices):
        s_i = length // r
        nb = (1 << LOGR) // r
        last = i == m - 1
        new = buf.copy()
        for u in range(nb):
            for t in range(t_count):
                b = t + t_count * u
                if last and m > 1:
                    s = hexrev(b, m - 1)
                    pos = s * r + np.arange(r)
                    c = 0
                else:
                    c, s = b % s_i, b // s_i
                    pos = s * length + c + np.arange(r) * s_i
                a = buf[pos]
                q = np.arange(r)
                dft = np.exp(-2j * np.pi * np.outer(q, q) / r) @ a
                if not last:
                    dft = dft * np.exp(-2j * np.pi * c * q / length)
                    new[pos] = dft
                else:
                    kl = b
                    out[kl + (n // r) * q] = dft
        buf = new
        length //= r
    return out


def wavefronts(pos, elem_bytes, pads):
    """Total shared-memory wavefronts for one warp-wide access of elem_bytes per lane."""
    total = 0
    ideal = 0
    lanes_per_phase = 128 // elem_bytes
    for w0 in range(0, len(pos), 32):
        warp = pos[w0:w0 + 32]
        for p0 in range(0, len(warp), lanes_per_phase):
            grp = warp[p0:p0 + lanes_per_phase]
            addr = np.array([phys(int(p), pads) for p in grp]) * elem_bytes
            words = set()
            per_bank = {}
            for a in addr:
                for wd in range(a // 4, (a + elem_bytes) // 4):
                    if wd not in words:
                        words.add(wd)
                        per_bank[wd % 32] = per_bank.get(wd % 32, 0) + 1
            total += max(per_bank.values())
            ideal += 1
    return total, ideal


def check(n, elem_bytes, pads):
    tot = ide = 0
    worst = 1.0
    for i, u, j, pos in accesses(n):
        if i == 0 and False:
            continue
        w, d = wavefronts(pos, elem_bytes, pads)
        tot += w
        ide += d
        worst = max(worst, w / d)
    return tot / ide, worst


def search(n, elem_bytes):
    lg = n.bit_length() - 1
    shifts = list(range(1, lg))
    best = None
    cands = [()]
    for k in (1, 2, 3):
        for combo in itertools.combinations(shifts, k):
            for coefs in itertools.product((1, 2, 4, 8), repeat=k):
                cands.append(tuple(zip(combo, coefs)))
    for pads in cands:
        extra = phys(n - 1, pads) + 1 - n
        if extra > n // 8:
            continue
        avg, worst = check(n, elem_bytes, pads)
        key = (round(avg, 4), len(pads), extra)
        if best is None or key < best[0]:
            best = (key, pads, worst)
            if avg == 1.0:
                break
    return best


if __name__ == "__main__":
    if "r8" in sys.argv:
        LOGR = 3
    rng = np.random.default_rng(0)
    for n in (64, 128, 256, 512, 1024, 2048, 4096, 8192):
        x = rng.standard_normal(n) + 1j * rng.standard_normal(n)
        err = np.max(np.abs(model_fft(x) - np.fft.fft(x)))
        print(f"N={n:5d} plan={plan(n)} max|err|={err:.2e}")
    if "banks" in sys.argv:
        for n in (512, 1024, 2048, 4096, 8192, 16384):
            for eb in (8, 16):
                (key, pads, worst) = search(n, eb)
                print(f"N={n:5d} elem={eb:2d}B pads={pads} avg={key[0]} worst={worst:.2f} extra={key[2]}")
