#!/usr/bin/env python
"""Numpy model of fft_wl_kernel's index algebra (csrc/tdsa_fft_wl.cuh) and of its shared-memory access patterns.

Not product code: it mirrors, thread by thread, what the 4096-point warp-local kernel does —
lane -> (sub-transform r, team lane c), the 128-byte-swizzled stage reads, pass A, the team-local 16x16 transpose
(row pitch 17), pass B with W256^(j ka), the Y regions, the last pass with W4096^(r kk) — so that the decomposition
can be checked against numpy.fft on the CPU and every access pattern can be checked for bank conflicts.
"""
import sys

import numpy as np

N, TH = 4096, 256


def thread_identity(tid):
    """tdsa_launch.cuh: wl_thread_identity (the host permutes the window with the same map)."""
    w, l = tid >> 5, tid & 31
    return 2 * w + ((l >> 3) & 1), (l & 7) + 8 * (l >> 4)


def stage_offset(tid, j):
    """Byte offset inside a staged frame of sample n = r + 16 c + 256 j: row m = n >> 4 (128 bytes), column r;
    the TMA swizzle XORs the 16-byte chunk index with (row & 7)."""
    w, l = tid >> 5, tid & 31
    h, c = (l >> 3) & 1, (l & 7) + 8 * (l >> 4)
    return c * 128 + (((w ^ c) & 7) << 4) + h * 8 + j * 2048


def swizzled_address(n):
    """Where cp.async.bulk.tensor with CU_TENSOR_MAP_SWIZZLE_128B puts sample n (8 bytes) of a frame."""
    byte = 8 * n
    row, chunk, rest = byte >> 7, (byte >> 4) & 7, byte & 15
    return (row << 7) | ((chunk ^ (row & 7)) << 4) | rest


def dft16(a):
    k = np.arange(16)
    return np.exp(-2j * np.pi * np.outer(k, k) / 16) @ a


def model_fft(x):
    """One frame through the kernel's three passes; returns the (unshifted) spectrum."""
    region = np.zeros((16, 272), dtype=np.complex128)
    y = np.zeros((16, 272), dtype=np.complex128)
    # pass A + team transpose
    a_out = {}
    for tid in range(TH):
        r, c = thread_identity(tid)
        a_out[(r, c)] = dft16(np.array([x[r + 16 * c + 256 * j] for j in range(16)]))
        for q in range(16):
            region[r, c + 17 * q] = a_out[(r, c)][q]
    # pass B: thread ka = c reads A_j[ka] at 17*ka + j
    for tid in range(TH):
        r, ka = thread_identity(tid)
        inp = np.array([region[r, 17 * ka + j] for j in range(16)])
        tw = np.exp(-2j * np.pi * np.arange(16) * ka / 256)
        out = dft16(inp * tw)
        for kb in range(16):
            y[r, ka + 17 * kb] = out[kb]                       # Y_r[ka + 16 kb] at kk + (kk >> 4)
    # last pass: thread kk reads Y_r[kk], r = 0..15
    spec = np.zeros(N, dtype=np.complex128)
    for kk in range(TH):
        inp = np.array([y[r, kk + (kk >> 4)] for r in range(16)])
        tw = np.exp(-2j * np.pi * np.arange(16) * kk / 4096)
        out = dft16(inp * tw)
        for q in range(16):
            spec[kk + 256 * q] = out[q]
    return spec


def engine_tables(e):
    """Twiddle tables of engine e (tdsa_api.cu: build_wl_tables): pass B [j][ka] and last pass [j][kk].
    Engine 0 is the plain 4096-point plan; engine 1 evaluates every table at the half-integer bin k + 1/2."""
    j = np.arange(16)[:, None]
    if e == 0:
        return np.exp(-2j * np.pi * j * np.arange(16)[None, :] / 256), np.exp(-2j * np.pi * j * np.arange(256)[None, :] / 4096)
    return (np.exp(-2j * np.pi * j * (2 * np.arange(16)[None, :] + 1) / 512),
            np.exp(-2j * np.pi * j * (2 * np.arange(256)[None, :] + 1) / 8192))


def model_fft8192(x):
    """N = 8192 on two engines: radix-2 DIF on the staged read, engine 0 = even bins (sum), engine 1 = odd bins
    (difference, half-bin tables, pass-A inputs pre-twiddled by the constants W32^j)."""
    spec = np.zeros(8192, dtype=np.complex128)
    w32 = np.exp(-2j * np.pi * np.arange(16) / 32)
    for e in range(2):
        s = x[:4096] + (1 - 2 * e) * x[4096:]
        twb, twl = engine_tables(e)
        region = np.zeros((16, 272), dtype=np.complex128)
        y = np.zeros((16, 272), dtype=np.complex128)
        for tid in range(TH):
            r, c = thread_identity(tid)
            inp = np.array([s[r + 16 * c + 256 * j] for j in range(16)])
            out = dft16(inp * (w32 if e == 1 else 1.0))
            for q in range(16):
                region[r, c + 17 * q] = out[q]
        for tid in range(TH):
            r, ka = thread_identity(tid)
            inp = np.array([region[r, 17 * ka + j] for j in range(16)])
            out = dft16(inp * twb[:, ka])
            for kb in range(16):
                y[r, ka + 17 * kb] = out[kb]
        for kk in range(TH):
            inp = np.array([y[r, kk + (kk >> 4)] for r in range(16)])
            out = dft16(inp * twl[:, kk])
            for q in range(16):
                spec[2 * (kk + 256 * q) + e] = out[q]
    return spec


def wavefronts(byte_addrs, elem_bytes):
    """Shared-memory wavefronts of one warp-wide access (32 byte addresses), processed in phases of 128/elem lanes."""
    lanes = 128 // elem_bytes
    total = 0
    for p0 in range(0, 32, lanes):
        per_bank = {}
        seen = set()
        for a in byte_addrs[p0:p0 + lanes]:
            for wd in range(a // 4, (a + elem_bytes) // 4):
                if wd not in seen:
                    seen.add(wd)
                    per_bank[wd % 32] = per_bank.get(wd % 32, 0) + 1
        total += max(per_bank.values())
    return total, 32 // lanes


def bank_report(elem_bytes):
    """(pattern name, worst wavefronts / ideal) for every shared-memory access pattern of the kernel."""
    region = 272 + (8 if elem_bytes == 8 else 0)
    out = []

    def worst(name, fn, eb):
        ratio = 0.0
        for w in range(8):
            for k in range(16):
                addrs = [fn(32 * w + l, k) for l in range(32)]
                got, ideal = wavefronts(addrs, eb)
                ratio = max(ratio, got / ideal)
        out.append((name, ratio))

    ident = thread_identity
    worst("stage read (swizzled, LDS.64)", lambda tid, j: stage_offset(tid, j), 8)
    worst("team store A_c[q] at c + 17 q", lambda tid, q: (ident(tid)[0] * region + ident(tid)[1] + 17 * q) * elem_bytes, elem_bytes)
    worst("team load A_j[ka] at 17 ka + j", lambda tid, j: (ident(tid)[0] * region + 17 * ident(tid)[1] + j) * elem_bytes, elem_bytes)
    worst("pass-B twiddle load [j][ka]", lambda tid, j: (j * 16 + ident(tid)[1]) * elem_bytes, elem_bytes)
    worst("last-pass load Y_r[kk]", lambda tid, r: (r * region + tid + (tid >> 4)) * elem_bytes, elem_bytes)
    return out


def check_swizzle():
    """The kernel's closed-form stage offset equals the TMA swizzle applied to the sample's linear address."""
    for tid in range(TH):
        r, c = thread_identity(tid)
        for j in range(16):
            if stage_offset(tid, j) != swizzled_address(r + 16 * c + 256 * j):
                return False
    return True


if __name__ == "__main__":
    rng = np.random.default_rng(0)
    x = rng.standard_normal(N) + 1j * rng.standard_normal(N)
    print("max |model - numpy.fft| =", np.abs(model_fft(x) - np.fft.fft(x)).max())
    x8 = rng.standard_normal(8192) + 1j * rng.standard_normal(8192)
    print("max |model8192 - numpy.fft| =", np.abs(model_fft8192(x8) - np.fft.fft(x8)).max())
    print("stage offsets match the 128B swizzle:", check_swizzle())
    print("thread identity is a bijection:", len({thread_identity(t) for t in range(TH)}) == TH)
    for eb in (8, 16):
        for name, ratio in bank_report(eb):
            print(f"elem {eb:2d} B  {name:34s} wavefronts / ideal = {ratio:.2f}")
    sys.exit(0)
