"""pytest configuration: `gpu` marker, repo on sys.path, shared fixtures.

-m "not gpu": oracle vs golden vectors / cv2, host logic, C-ABI export check, gloo world_size-2 tests.
-m gpu      : the parity tests proper — CUDA path (through the C ABI) vs the CPU oracle and the fixtures.
Nothing here reads /root/reference at run time.
"""
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")
    config.addinivalue_line("markers", "slow: long-running sweep")


@pytest.fixture(scope="session")
def oracle_mod():
    from oracle import oracle
    oracle.lib()
    return oracle


@pytest.fixture(scope="session")
def orbx_mod():
    from dani_slam_b200 import orbx
    L = orbx.lib()
    if L.orbx_device_count() < 1:
        pytest.fail("GPU test selected but no CUDA device is visible (liborbx has no CPU fallback)")
    return orbx


def golden_cases():
    return sorted(f[len("extract_"):-4] for f in os.listdir(GOLDEN) if f.startswith("extract_"))


def load_golden(name):
    return np.load(os.path.join(GOLDEN, f"extract_{name}.npz"))


def golden_frame(g):
    from dani_slam_b200 import synth
    kind, seed, W, H = str(g["kind"]), int(g["seed"]), int(g["W"]), int(g["H"])
    img = synth.parity_frame(seed, W, H) if kind == "parity" else synth.throughput_frame(seed, W, H)
    import hashlib
    assert hashlib.sha256(img.tobytes()).hexdigest() == str(g["frame_sha"]), "synthetic frame generator drifted"
    return img
