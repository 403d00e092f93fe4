"""Shared fixtures.  Tests that need a GPU carry @pytest.mark.gpu; everything else runs on CPU.

Only tests (and smoke / bench's cpu_baseline) may import the oracle."""
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def _gpu_count():
    """CUDA devices as libldagpu sees them (0 when the library or the driver is missing)."""
    try:
        import ldagroupedgibbssampler_b200 as L
        return int(L.load().ldagpu_device_count())
    except Exception:
        return 0


def pytest_collection_modifyitems(config, items):
    """A plain `pytest tests` on a box without a GPU skips the gpu-marked tests instead of failing them."""
    if not any("gpu" in it.keywords for it in items):
        return
    if _gpu_count() > 0:
        return
    skip = pytest.mark.skip(reason="no CUDA device: libldagpu has no CPU fallback")
    for it in items:
        if "gpu" in it.keywords:
            it.add_marker(skip)


def make_corpus(D, V, mean_len, seed, sort_docs=True, empty_every=0, zipf=1.1):
    """Small ragged corpus (numpy only): Zipf-ish types, Poisson lengths, optional empty documents."""
    rng = np.random.default_rng(seed)
    lens = rng.poisson(mean_len, D).astype(np.int64)
    if empty_every:
        lens[::empty_every] = 0
    off = np.zeros(D + 1, np.int64)
    np.cumsum(lens, out=off[1:])
    p = 1.0 / np.arange(1, V + 1) ** zipf
    p /= p.sum()
    tokens = rng.choice(V, size=int(off[-1]), p=p).astype(np.int32)
    if sort_docs:
        for d in range(D):
            tokens[off[d]:off[d + 1]].sort()
    return off, tokens


@pytest.fixture(scope="session")
def cats():
    f = np.load(os.path.join(GOLDEN, "cats_corpus.npz"))
    return f["doc_offsets"].astype(np.int64), f["tokens"].astype(np.int32)


@pytest.fixture(scope="session")
def golden():
    return np.load(os.path.join(GOLDEN, "oracle_golden.npz"))


@pytest.fixture(scope="session")
def oracle():
    from oracle import oracle as O
    O.lib()
    return O
