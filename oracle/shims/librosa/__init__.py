"""TEST INFRASTRUCTURE ONLY -- stand-in so that the reference's mixing_manipulator modules import in this image (librosa is
not installed).  `stft` forwards to the restatement in oracle/norm_oracle.py (center=False only, which is all the pinned
reference code uses: common_miscellaneous.py:72-76); `util.frame` is numpy's sliding window (get_mean_peak)."""
from . import display  # noqa: F401
from . import util  # noqa: F401


def stft(y, n_fft=2048, hop_length=None, window='hann', center=True, **_):
    if center or isinstance(window, str):
        raise NotImplementedError("librosa shim: only center=False with an explicit window array")
    from oracle.norm_oracle import stft as _stft
    return _stft(y, n_fft, hop_length, window)
