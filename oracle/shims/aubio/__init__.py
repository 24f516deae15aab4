"""TEST INFRASTRUCTURE ONLY -- aubio (the onset detector behind the compressor matching, utils_data_normalization.py:302-312)
is a third-party C library that is not installed here.  This stand-in is NOT aubio's 'hfc' detector: it is a small
deterministic rule (a frame is an onset when its energy exceeds twice the running mean of the frames before it) with the same
call protocol -- `o = onset(method, buf_size=, hop_size=, samplerate=)`, `o(frame)` truthy on an onset, `o.get_last()` the onset
position in samples -- so that the reference's own search logic around it (get_mean_peak, get_comp_matching) runs and can be
pinned.  The GPU tests inject the same rule (oracle.norm_oracle.stub_onsets) into the product."""
import numpy as np


class onset:
    def __init__(self, method="default", buf_size=1024, hop_size=512, samplerate=44100):
        self.hop, self.n, self.mean, self.last = hop_size, 0, 0.0, 0

    def __call__(self, frame):
        e = float(np.mean(np.square(np.asarray(frame, dtype=np.float64))))
        hit = self.n > 0 and e > 2.0 * self.mean and e > 1e-8
        if hit:
            self.last = self.n * self.hop
        self.mean = (self.mean * self.n + e) / (self.n + 1)
        self.n += 1
        return np.array([1.0 if hit else 0.0], dtype=np.float32)

    def get_last(self):
        return self.last
