"""CPU check of the classic kernel's pass structure (tools/fft_plan_model.py mirrors csrc/tdsa_fft.cuh): the radix plan,
thread -> butterfly mapping and digit-reversed last pass reproduce numpy.fft for every single-CTA size."""
import importlib.util
import os

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
spec = importlib.util.spec_from_file_location("fft_plan_model", os.path.join(ROOT, "tools", "fft_plan_model.py"))
M = importlib.util.module_from_spec(spec)
spec.loader.exec_module(M)


@pytest.mark.parametrize("n", [64, 128, 256, 512, 1024, 2048, 4096, 8192])
def test_pass_structure_reproduces_the_dft(n):
    rng = np.random.default_rng(n)
    x = rng.standard_normal(n) + 1j * rng.standard_normal(n)
    assert np.abs(M.model_fft(x) - np.fft.fft(x)).max() < 1e-9 * n


def test_plans():
    assert M.plan(4096) == [16, 16, 16] and M.plan(8192) == [16, 16, 16, 2] and M.plan(1024) == [16, 16, 4]
