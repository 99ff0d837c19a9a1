import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "ecg-byte_b200")):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def oracle():
    from oracle import oracle as O
    O.build()
    return O


@pytest.fixture(scope="session")
def small_corpus():
    """20 synthetic records, 12 x 1000 samples, with their stats dict."""
    import numpy as np
    from ecgbyte import synth
    x = synth.corpus(7, 20, L=1000, dtype=np.float64)
    return x, synth.percentiles(x, seed=7)


@pytest.fixture(scope="session")
def small_table(oracle, small_corpus):
    """(pairs, merges-in-reference-form) of 600 merges trained by the oracle."""
    x, pct = small_corpus
    sym = oracle.quantize(x, pct["percentile_1"], pct["percentile_99"])
    ids, pairs, counts, ntied = oracle.train_pairs(sym.reshape(-1), 600, fast=True)
    _, vocab, merges = oracle.to_reference_types(ids, pairs)
    return pairs, vocab, merges
