"""The env-gated TCN kernel variants must stay parity-green: the whole TCN parity file is re-run in
a child process per variant, because the library reads the switches once per process."""
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

VARIANTS = [
    {"MST_TCN_PAIRED": "0"},                                   # f16f8 (default precision), plain 256-row tiles
    {"MST_TCN_PRECISION": "f16f8", "MST_TCN_PIPE": "2"},       # f16f8 through the dual-ring kernel of tcn_f8.cu
    {"MST_TCN_PRECISION": "bf16x3"},                           # three bf16 products, paired sub-tiles
    {"MST_TCN_PRECISION": "bf16x3", "MST_TCN_PAIRED": "0"},
    {"MST_TCN_PRECISION": "bf16x3", "MST_TCN_PIPE": "2"},
    {"MST_TCN_PRECISION": "bf16x3", "MST_TCN_KCHUNK": "32"},
]


@pytest.mark.gpu
@pytest.mark.parametrize("env", VARIANTS, ids=lambda e: ",".join(f"{k}={v}" for k, v in e.items()))
def test_tcn_parity_under_variant(env):
    child_env = dict(os.environ)
    child_env.update(env)
    r = subprocess.run([sys.executable, "-m", "pytest", os.path.join(ROOT, "tests", "test_gpu_tcn.py"), "-m", "gpu", "-x", "-q"],
                       cwd=ROOT, env=child_env, capture_output=True, text=True, timeout=900)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-2000:]
