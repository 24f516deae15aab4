"""TEST INFRASTRUCTURE ONLY -- restatement of pymixconsole.components.iirfilter.IIRfilter.

The real class lives in the un-vendored dependency pymixconsole==0.0.1 (reference requirements.txt:12); its
source is NOT under /root/reference, so this is a restatement of the published algorithm (RBJ Audio-EQ-Cookbook
biquads, the reference's own docstring common_audioeffects.py:375-376) anchored on the reference call sites
common_audioeffects.py:460 (ctor), :471-476 (G/fc/rate/Q attributes), :494,:499,:512 (reset_state),
:513 (apply_filter).  PARITY UNPINNED: no golden vector in the reference pins this boundary.
"""
import numpy as np
import scipy.signal


def rbj_coefficients(G, Q, fc, rate, filter_type):
    """RBJ cookbook low_shelf / peaking / high_shelf, normalised by a0 (float64)."""
    A = 10.0 ** (G / 40.0)
    w0 = 2.0 * np.pi * (fc / rate)
    alpha = np.sin(w0) / (2.0 * Q)
    c = np.cos(w0)
    s = 2.0 * np.sqrt(A) * alpha
    if filter_type == "peaking":
        b = [1.0 + alpha * A, -2.0 * c, 1.0 - alpha * A]
        a = [1.0 + alpha / A, -2.0 * c, 1.0 - alpha / A]
    elif filter_type == "low_shelf":
        b = [A * ((A + 1) - (A - 1) * c + s), 2 * A * ((A - 1) - (A + 1) * c), A * ((A + 1) - (A - 1) * c - s)]
        a = [(A + 1) + (A - 1) * c + s, -2 * ((A - 1) + (A + 1) * c), (A + 1) + (A - 1) * c - s]
    elif filter_type == "high_shelf":
        b = [A * ((A + 1) + (A - 1) * c + s), -2 * A * ((A - 1) + (A + 1) * c), A * ((A + 1) + (A - 1) * c - s)]
        a = [(A + 1) - (A - 1) * c + s, 2 * ((A - 1) - (A + 1) * c), (A + 1) - (A - 1) * c - s]
    else:
        raise ValueError(f"unknown filter_type {filter_type!r}")
    b = np.asarray(b, dtype=np.float64) / a[0]
    a = np.asarray(a, dtype=np.float64) / a[0]
    return b, a


class IIRfilter:
    def __init__(self, G, Q, fc, rate, filter_type, n_channels=2):
        self.G, self.Q, self.fc, self.rate = G, Q, fc, rate
        self.filter_type = filter_type
        self.n_channels = n_channels
        self.reset_state()

    def reset_state(self):
        self.zi = np.zeros((2, self.n_channels), dtype=np.float64)

    def apply_filter(self, x):
        b, a = rbj_coefficients(self.G, self.Q, self.fc, self.rate, self.filter_type)
        y, self.zi = scipy.signal.lfilter(b, a, np.asarray(x, dtype=np.float64), axis=0, zi=self.zi)
        return y
