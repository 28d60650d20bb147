import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a B200 (run with -m gpu on the GPU box)")


def uniform_cloud(n, d=3, dtype=np.float32, seed=1808, massive_ratio=1.0):
    """The reference bench's workload shape (benches/benchmark.rs:22-37): positions
    U[-5e3, 5e3)^D, mu U[1e3, 1e9) for the first round(n * ratio) bodies, 0 for the rest."""
    rng = np.random.default_rng(seed)
    pos = rng.uniform(-5e3, 5e3, (n, d))
    mu = rng.uniform(1e3, 1e9, (n, 1))
    mu[int(round(n * massive_ratio)):] = 0.0
    return np.ascontiguousarray(np.concatenate([pos, mu], axis=1).astype(dtype))


def plummer_cloud(n, d=3, seed=1808, dtype=np.float32):
    """Plummer sphere, scale a = 1, r < 50 a, equal mu = 1/n."""
    rng = np.random.default_rng(seed)
    r = np.empty(0)
    while len(r) < n:
        u = rng.uniform(1e-12, 1.0, n)
        rr = 1.0 / np.sqrt(u ** (-2.0 / 3.0) - 1.0)
        r = np.concatenate([r, rr[rr < 50.0]])
    r = r[:n]
    v = rng.normal(size=(n, d))
    v /= np.linalg.norm(v, axis=1, keepdims=True)
    pos = v * r[:, None]
    mu = np.full((n, 1), 1.0 / n)
    return np.ascontiguousarray(np.concatenate([pos, mu], axis=1).astype(dtype))


def rel_err(a, ref):
    """Per-particle relative error ||a - ref|| / ||ref||."""
    a = np.asarray(a, dtype=np.float64)
    ref = np.asarray(ref, dtype=np.float64)
    den = np.linalg.norm(ref, axis=1)
    num = np.linalg.norm(a - ref, axis=1)
    out = np.zeros_like(num)
    nz = den > 0
    out[nz] = num[nz] / den[nz]
    out[~nz] = num[~nz]
    return out


@pytest.fixture(scope="session")
def ctx():
    import particular_b200 as pb
    c = pb.CudaContext(0)
    yield c
    c.close()
