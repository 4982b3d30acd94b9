import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def golden():
    import numpy as np

    def load(name):
        return np.load(os.path.join(GOLDEN, name), allow_pickle=False)
    return load


@pytest.fixture(scope="session")
def parity_log():
    """record(name, max_err, **extra): keeps the measured error of every parity check and appends it to
    gpurun_out/parity.jsonl (copied into profiles/ per round), so that the bar and the measurement sit side by side."""
    import json
    out_dir = os.path.join(ROOT, "gpurun_out")
    os.makedirs(out_dir, exist_ok=True)
    path = os.path.join(out_dir, "parity.jsonl")

    def record(name, max_err, **extra):
        rec = {"check": name, "max_err": float(max_err)}
        rec.update(extra)
        with open(path, "a") as f:
            f.write(json.dumps(rec) + "\n")
        print(f"[parity] {name}: max err {float(max_err):.3e} {extra if extra else ''}")
        return float(max_err)
    return record
