"""TEST INFRASTRUCTURE ONLY -- seeded inputs shared by oracle/make_golden.py (which stores the reference outputs)
and the tests (which regenerate the same inputs and compare against tests/golden/*.npz)."""
import os

import numpy as np
import torch

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")
FULL_WINDOWS = [(0, 2048), (65536, 2048), (131072 - 1024, 2048), (262144 - 2048, 2048)]
FULL_STRIDE = 257
BLOCK_LEN, BLOCK_CH_STRIDE = 3000, 16


def make_cond(n, seed, dim=2048):
    g = torch.Generator()
    g.manual_seed(seed)
    return torch.randn(n, dim, generator=g).abs() * 0.5


def fx_input(i, L):
    rng = np.random.RandomState(100 + i)
    x = (rng.randn(L, 2) * 0.1).astype(np.float32)
    x[:, 1] = (0.6 * x[:, 0] + 0.4 * x[:, 1]).astype(np.float32)
    # a few loud bursts so the compressor's attack/release pattern is exercised
    env = (1.0 + 4.0 * (np.sin(np.arange(L) / 700.0 + i) > 0.8)).astype(np.float32)
    return np.clip(x * env[:, None], -1, 1).astype(np.float32)


def load_golden(name):
    return np.load(os.path.join(GOLDEN_DIR, name))
