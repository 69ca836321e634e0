import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLD = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session", autouse=True)
def _built():
    """Build the checker (oracle/) and the product library once per session."""
    from oracle import cfd_oracle
    cfd_oracle.build()
    from compact_finite_differences_b200 import build as b
    b.build()
    yield


def convergence_ratios(derivative, axis, sizes=(16, 32, 64, 128)):
    """Order-of-convergence experiment of the reference (code/cuda/test/test_convergence.py:25-52): field
    sin x + cos(xy) + zx on [0, 2 pi]^3 (cyclically renamed so that `axis` plays the part of x), errors normalised by
    max |f'|; returns (mean-error ratios, max-error ratios) between successive grid sizes.  `derivative(f, axis, h)`
    is the implementation under test."""
    import numpy as np
    mean_errs, max_errs = [], []
    for n in sizes:
        t = np.linspace(0, 2 * np.pi, n)
        c = np.meshgrid(t, t, t, indexing="ij")           # c[0] = z, c[1] = y, c[2] = x
        u = c[2 - axis]                                    # the differentiated coordinate
        v, w = c[2 - (axis + 1) % 3], c[2 - (axis + 2) % 3]
        f = np.sin(u) + np.cos(v * u) + w * u
        true = np.cos(u) - v * np.sin(v * u) + w
        d = derivative(np.ascontiguousarray(f), axis, t[1] - t[0])
        err = np.abs(d - true) / np.abs(d).max()
        mean_errs.append(err.mean())
        max_errs.append(err.max())
    k = len(sizes) - 1
    return [mean_errs[i] / mean_errs[i + 1] for i in range(k)], [max_errs[i] / max_errs[i + 1] for i in range(k)]
