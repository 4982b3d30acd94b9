"""Seeded synthetic IQ for the five BASELINE.json configs.

Everything is ``np.random.default_rng(seed)``; IQ is stored complex64,
C-contiguous ``[B, N]`` (or a flat stream).  Used by the tests, ``bench.py``
and the golden-vector generator so that all three see the same samples.
"""
from __future__ import annotations

import numpy as np


def awgn(rng: np.random.Generator, shape, sigma2: float = 1.0) -> np.ndarray:
    """Circular complex Gaussian noise with E|x|^2 = sigma2, complex64."""
    s = np.sqrt(sigma2 / 2.0)
    out = np.empty(shape, dtype=np.complex64)
    out.real = rng.standard_normal(shape, dtype=np.float32) * np.float32(s)
    out.imag = rng.standard_normal(shape, dtype=np.float32) * np.float32(s)
    return out


def tone(n_total: int, cycles_per_n: float, n: int, amp: float, phase: float = 0.0) -> np.ndarray:
    """exp(j*2*pi*k*t/n) over a flat stream of n_total samples (complex128)."""
    t = np.arange(n_total, dtype=np.float64)
    return amp * np.exp(1j * (2.0 * np.pi * cycles_per_n * t / n + phase))


def cfg1_frames(b: int = 64, n: int = 1024, fs: float = 2.048e6, seed: int = 0) -> np.ndarray:
    """Config 1: (g1 + j g2)/sqrt(2) plus a tone A=0.5 at +250 kHz."""
    rng = np.random.default_rng(seed)
    x = awgn(rng, (b, n)).astype(np.complex128)
    t = np.arange(b * n, dtype=np.float64).reshape(b, n)
    x += 0.5 * np.exp(2j * np.pi * 250e3 * t / fs)
    return np.ascontiguousarray(x.astype(np.complex64))


def cfg2_frames(b: int = 8192, n: int = 4096, seed: int = 1, tones: bool = True,
                chunk: int = 1024) -> np.ndarray:
    """Config 2: AWGN sigma^2=1 + on-bin k=512, off-bin k=1800.37, weak -60 dBc k=3000.

    Generated in chunks of frames so the float64 temporaries stay small.
    """
    rng = np.random.default_rng(seed)
    out = np.empty((b, n), dtype=np.complex64)
    t = np.arange(n, dtype=np.float64)
    if tones:
        base = (4.0 * np.exp(2j * np.pi * 512.0 * t / n)
                + 2.0 * np.exp(2j * np.pi * 1800.37 * t / n + 0.3j)
                + 4.0e-3 * np.exp(2j * np.pi * 3000.0 * t / n + 1.1j))
    for lo in range(0, b, chunk):
        hi = min(b, lo + chunk)
        x = awgn(rng, (hi - lo, n))
        if tones:
            # frame-dependent start phase so frames differ deterministically
            ph = np.exp(2j * np.pi * 0.61803398875 * np.arange(lo, hi, dtype=np.float64))[:, None]
            x = (x.astype(np.complex128) + base[None, :] * ph).astype(np.complex64)
        out[lo:hi] = x
    return out


def cfg3_stream(n_samples: int = 1 << 26, seed: int = 2, chunk: int = 1 << 22) -> np.ndarray:
    """Config 3: AWGN + slow chirp, flat complex64 stream."""
    rng = np.random.default_rng(seed)
    out = np.empty(n_samples, dtype=np.complex64)
    for lo in range(0, n_samples, chunk):
        hi = min(n_samples, lo + chunk)
        t = np.arange(lo, hi, dtype=np.float64)
        # instantaneous frequency sweeps 0.05 -> 0.15 cycles/sample over the stream
        phase = 2.0 * np.pi * (0.05 * t + 0.05 * t * t / n_samples)
        x = awgn(rng, hi - lo).astype(np.complex128) + 3.0 * np.exp(1j * phase)
        out[lo:hi] = x.astype(np.complex64)
    return out


def cfg4_subbands(n_bands: int = 300, frames: int = 16, n: int = 8192, seed: int = 3,
                  bands: range | None = None) -> np.ndarray:
    """Config 4: per-band AWGN + band-specific tone -> ``[n_bands, frames, n]`` complex64.

    ``bands`` restricts generation to a contiguous block (what one rank owns);
    each band draws from its own ``default_rng([seed, band])`` so shards agree
    with the full array.
    """
    bands = range(n_bands) if bands is None else bands
    out = np.empty((len(bands), frames, n), dtype=np.complex64)
    t = np.arange(frames * n, dtype=np.float64).reshape(frames, n)
    for i, band in enumerate(bands):
        rng = np.random.default_rng([seed, band])
        k = 64.0 + (band * 97) % (n - 128)           # band-specific bin
        amp = 1.0 + (band % 7)
        x = awgn(rng, (frames, n)).astype(np.complex128)
        x += amp * np.exp(2j * np.pi * k * t / n)
        out[i] = x.astype(np.complex64)
    return out


def cfg5_chunk(index: int, chunk: int = 65536, n: int = 4096, seed: int = 4) -> np.ndarray:
    """Config 5: chunk ``index`` of the continuous 20 Msps stream (AWGN + hopping tone)."""
    rng = np.random.default_rng([seed, index])
    t = np.arange(chunk, dtype=np.float64) + float(index) * chunk
    k = 200.0 + 37.0 * (index % 64)                  # tone hops once per chunk
    x = awgn(rng, chunk).astype(np.complex128) + 2.0 * np.exp(2j * np.pi * k * t / n)
    return x.astype(np.complex64)
