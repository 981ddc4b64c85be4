import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a B200 (run with -m gpu on the GPU box)")


@pytest.fixture(scope="session")
def oracle():
    import oracle_bind
    oracle_bind.lib()
    return oracle_bind


@pytest.fixture(scope="session")
def ctx():
    """GPU context; fails (not skips) when the CUDA library or the device is missing"""
    import stark_pure_rust_b200 as sb
    return sb.default_context()


P = 21888242871839275222246405745257275088548364400416034343698204186575808495617


def random_elems(n, seed):
    """n canonical field elements, seeded, as (n, 4) uint64 Montgomery limbs (values themselves are
    uniform in [0, p): drawing the Montgomery representative uniformly is the same distribution)"""
    rng = np.random.default_rng(seed)
    out = np.empty((n, 4), dtype=np.uint64)
    filled = 0
    p_limbs = [(P >> (64 * i)) & (2**64 - 1) for i in range(4)]
    while filled < n:
        m = n - filled
        cand = rng.integers(0, 2**64, size=(m + m // 4 + 16, 4), dtype=np.uint64)
        cand[:, 3] &= np.uint64((1 << 62) - 1)          # < 2^254
        # keep candidates < p (compare from the top limb down)
        lt = np.zeros(cand.shape[0], dtype=bool)
        eq = np.ones(cand.shape[0], dtype=bool)
        for i in (3, 2, 1, 0):
            pl = np.uint64(p_limbs[i])
            lt |= eq & (cand[:, i] < pl)
            eq &= cand[:, i] == pl
        good = cand[lt][:m]
        out[filled:filled + good.shape[0]] = good
        filled += good.shape[0]
    return out
