"""TEST INFRASTRUCTURE ONLY -- numpy restatement of the reference's WAV sample handling around the forward
(SURVEY.md 8f-1), the parity oracle for csrc/pcm.cu.  Paths relative to /root/reference/.

  decode()      load_wav_segment, mixing_style_transfer/data_loader/loader_utils.py:54-70 (x / 2^15 or x / 2^31 in
                float64, stereo de-interleave), the stem clamp data_loader/data_loader.py:589-590, `.float()` of the
                inference entry (inference/style_transfer.py data path) and the mono duplication of
                inference/feature_extraction.py:87-89.
  encode_mix()  inference/style_transfer.py:165-177: `sum(inst_outputs)` over float32 arrays in instrument order, then the
                PCM_16 file.  The file is written by `soundfile` (libsndfile), which is NOT in this image and not vendored
                by the reference: the quantisation below (scale 2^15, round half to even, clip) restates what this repo's
                host writer does -- ** parity unpinned against libsndfile **.
Pinned against the reference's own `load_wav_segment` on its sample WAVs by tests/test_oracle_pinned.py (run where
/root/reference exists).
"""
import numpy as np


def decode(pcm):
    """pcm: int16 / int32 [n_frames, n_channels] -> float32 [2, n_frames]."""
    pcm = np.asarray(pcm)
    if pcm.ndim == 1:
        pcm = pcm[:, None]
    if pcm.dtype == np.int16:
        X = pcm / float(2 ** 15)          # :57-58
    elif pcm.dtype == np.int32:
        X = pcm / float(2 ** 31)          # :60-61
    else:
        raise ValueError("ValueError: input audio's bit depth should be 16 or 32-bit")   # :62-63
    X = X.T                                # [channel, frame] (axis=0 de-interleave, :65-69)
    if X.shape[0] == 1:
        X = np.concatenate((X, X), axis=0)  # feature_extraction.py:87-89
    return np.clip(X, -1.0, 1.0).astype(np.float32)   # data_loader.py:589-590, then .float()


def encode_mix(stems, n_frames=None):
    """stems: float32 [n_stems, 2, T] -> int16 [n_frames, 2]."""
    stems = np.asarray(stems, dtype=np.float32)
    mix = sum(stems[i] for i in range(stems.shape[0]))          # style_transfer.py:176, float32 adds in order
    n_frames = mix.shape[-1] if n_frames is None else n_frames
    data = mix[:, :n_frames].transpose(-1, -2)
    return np.clip(np.rint(np.asarray(data, dtype=np.float64) * 32768.0), -32768, 32767).astype('<i2')
