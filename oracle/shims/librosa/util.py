"""TEST INFRASTRUCTURE ONLY -- librosa.util.frame (axis=-1): [frame_length, n_frames] view of a 1-D signal."""
import numpy as np


def frame(x, frame_length, hop_length, axis=-1):
    v = np.lib.stride_tricks.sliding_window_view(np.asarray(x), frame_length)[::hop_length]
    return v.T
