"""TEST INFRASTRUCTURE ONLY -- pymixconsole.components.comb.Comb (pymixconsole==0.0.1, third-party, NOT in the reference tree):
the damped feedback comb filter of Freeverb, restated from the published recurrence.  PARITY UNPINNED.  Call site:
AlgorithmicReverb.update, common_audioeffects.py:1525-1540 -- Comb(buffer_size, damp, feedback, block_size)."""
import numpy as np

try:
    from numba import njit
except Exception:  # pragma: no cover
    def njit(*a, **k):
        return (lambda f: f) if not (a and callable(a[0])) else a[0]


@njit(cache=False)
def _comb(x, buf, store, idx, damp1, damp2, feedback):
    out = np.empty_like(x)
    n_buf = buf.shape[0]
    for n in range(x.shape[0]):
        o = buf[idx]
        store = (o * damp2) + (store * damp1)
        buf[idx] = x[n] + (store * feedback)
        out[n] = o
        idx += 1
        if idx >= n_buf:
            idx = 0
    return out, store, idx


class Comb:
    def __init__(self, buffer_size, damp, feedback, block_size=None):
        self.buffer = np.zeros(int(buffer_size), dtype=np.float64)
        self.damp1, self.damp2, self.feedback = float(damp), 1.0 - float(damp), float(feedback)
        self.filterstore, self.idx = 0.0, 0

    def process(self, x):
        out, self.filterstore, self.idx = _comb(np.asarray(x, dtype=np.float64), self.buffer, self.filterstore, self.idx,
                                                self.damp1, self.damp2, self.feedback)
        return out
