"""TEST INFRASTRUCTURE ONLY -- pymixconsole.components.allpass.Allpass (pymixconsole==0.0.1, third-party, NOT in the reference
tree): the Schroeder all-pass section of Freeverb, restated from the published recurrence.  PARITY UNPINNED.  Call site:
AlgorithmicReverb.update, common_audioeffects.py:1516-1523 -- Allpass(buffer_size, feedback, block_size)."""
import numpy as np

try:
    from numba import njit
except Exception:  # pragma: no cover
    def njit(*a, **k):
        return (lambda f: f) if not (a and callable(a[0])) else a[0]


@njit(cache=False)
def _allpass(x, buf, idx, feedback):
    out = np.empty_like(x)
    n_buf = buf.shape[0]
    for n in range(x.shape[0]):
        b = buf[idx]
        out[n] = -x[n] + b
        buf[idx] = x[n] + (b * feedback)
        idx += 1
        if idx >= n_buf:
            idx = 0
    return out, idx


class Allpass:
    def __init__(self, buffer_size, feedback, block_size=None):
        self.buffer = np.zeros(int(buffer_size), dtype=np.float64)
        self.feedback, self.idx = float(feedback), 0

    def process(self, x):
        out, self.idx = _allpass(np.asarray(x, dtype=np.float64), self.buffer, self.idx, self.feedback)
        return out
