"""CPU check of the warp-local 4096-point plan (tools/wl_plan_model.py mirrors csrc/tdsa_fft_wl.cuh thread by thread):
index algebra against numpy.fft, the closed-form swizzled stage offsets against the TMA 128-byte swizzle, and every
shared-memory access pattern against the bank model."""
import importlib.util
import os

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
spec = importlib.util.spec_from_file_location("wl_plan_model", os.path.join(ROOT, "tools", "wl_plan_model.py"))
M = importlib.util.module_from_spec(spec)
spec.loader.exec_module(M)


def test_three_passes_reproduce_the_dft():
    rng = np.random.default_rng(3)
    x = rng.standard_normal(M.N) + 1j * rng.standard_normal(M.N)
    assert np.abs(M.model_fft(x) - np.fft.fft(x)).max() < 1e-10


def test_two_engine_8192_plan_reproduces_the_dft():
    """Radix-2 DIF on the staged read + half-bin tables on engine 1 (csrc/tdsa_fft_wl.cuh, NB = 2)."""
    rng = np.random.default_rng(4)
    x = rng.standard_normal(8192) + 1j * rng.standard_normal(8192)
    assert np.abs(M.model_fft8192(x) - np.fft.fft(x)).max() < 1e-10


def test_thread_identity_and_window_permutation():
    ids = [M.thread_identity(t) for t in range(M.TH)]
    assert len(set(ids)) == M.TH and all(0 <= r < 16 and 0 <= c < 16 for r, c in ids)
    # every sample of a frame is read by exactly one (thread, j)
    seen = sorted(r + 16 * c + 256 * j for r, c in ids for j in range(16))
    assert seen == list(range(M.N))


def test_stage_offsets_follow_the_tma_swizzle():
    assert M.check_swizzle()
    # and the swizzle is a permutation of the frame's 8-byte slots
    assert sorted(M.swizzled_address(n) for n in range(M.N)) == [8 * n for n in range(M.N)]


def test_every_shared_memory_pattern_is_conflict_free():
    for elem_bytes in (8, 16):
        for name, ratio in M.bank_report(elem_bytes):
            assert ratio == 1.0, (elem_bytes, name, ratio)
