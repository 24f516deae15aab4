"""GPU sample-format kernels (csrc/pcm.cu) vs the numpy oracle: bit-exact (integer / byte work)."""
import os
import wave

import numpy as np
import pytest
import torch

from oracle import io_oracle

pytestmark = pytest.mark.gpu


def _pcm(n, ch, dtype, seed):
    rng = np.random.RandomState(seed)
    info = np.iinfo(dtype)
    x = rng.randint(info.min, info.max + 1, size=(n, ch), dtype=np.int64).astype(dtype)
    if n >= 4:
        x[0, :] = info.min
        x[1, :] = info.max
        x[2, :] = 0
        x[3, :] = -1
    return x


@pytest.mark.parametrize("n", [0, 1, 3, 4, 1023, 4096, 70001, 441000])
@pytest.mark.parametrize("ch", [1, 2])
@pytest.mark.parametrize("dtype", [np.int16, np.int32])
def test_decode_bit_exact(n, ch, dtype):
    from music_mixing_style_transfer_b200 import wav_io
    x = _pcm(n, ch, dtype, seed=n + ch)
    got = wav_io.decode_pcm(x).cpu().numpy()
    ref = io_oracle.decode(x) if n else np.zeros((2, 0), np.float32)
    assert got.shape == (2, n) and got.dtype == np.float32
    assert np.array_equal(got, ref)


def test_decode_unaligned_device_view():
    """A device tensor whose storage offset breaks the 16-byte alignment takes the scalar path."""
    from music_mixing_style_transfer_b200 import wav_io
    x = _pcm(5001, 2, np.int16, seed=5)
    dev = torch.from_numpy(x).cuda()
    got = wav_io.decode_pcm(dev[1:]).cpu().numpy()
    assert np.array_equal(got, io_oracle.decode(x[1:]))


@pytest.mark.parametrize("n,T,stems", [(0, 8, 1), (1, 1, 4), (4099, 4099, 4), (65536, 70000, 4), (262144, 262144, 2), (333, 1000, 3)])
def test_encode_mix_bit_exact(n, T, stems):
    from music_mixing_style_transfer_b200 import wav_io
    rng = np.random.RandomState(n + stems)
    x = (rng.randn(stems, 2, T) * 0.4).astype(np.float32)
    if T >= 8:
        # ties of the half-to-even rounding, the clip on both sides, exact +-1
        x[:, :, 0] = 0.0; x[0, :, 0] = 0.5 / 32768.0
        x[:, :, 1] = 0.0; x[0, :, 1] = 1.5 / 32768.0
        x[:, :, 2] = 0.0; x[0, :, 2] = -0.5 / 32768.0
        x[:, :, 3] = 1.0
        x[:, :, 4] = -1.0
        x[:, :, 5] = 0.0; x[0, :, 5] = 1.0
        x[:, :, 6] = 0.0; x[0, :, 6] = -1.0
        x[:, :, 7] = 0.0; x[0, :, 7] = 32766.5 / 32768.0
    got = wav_io.encode_mix_pcm16(torch.from_numpy(x).cuda(), n).cpu().numpy()
    ref = io_oracle.encode_mix(x, n)
    assert got.shape == (n, 2) and got.dtype == np.int16
    assert np.array_equal(got, ref)


def test_wav_file_round_trip(tmp_path):
    """File -> device -> file: load_wav_to_device equals the oracle on the same bytes, and writing the decoded stereo signal
    back as PCM_16 reproduces the file's samples exactly (x / 2^15 * 2^15 is exact)."""
    from music_mixing_style_transfer_b200 import wav_io
    x = _pcm(50000, 2, np.int16, seed=77)
    src = os.path.join(tmp_path, "a.wav")
    with wave.open(src, "wb") as w:
        w.setnchannels(2); w.setsampwidth(2); w.setframerate(44100); w.writeframes(x.tobytes())
    dev = wav_io.load_wav_to_device(src)
    assert np.array_equal(dev.cpu().numpy(), io_oracle.decode(x))
    dst = os.path.join(tmp_path, "b.wav")
    wav_io.write_wav_pcm16_from_device(dst, dev, 44100)
    with wave.open(dst, "rb") as w:
        assert (w.getnchannels(), w.getsampwidth(), w.getframerate(), w.getnframes()) == (2, 2, 44100, 50000)
        y = np.frombuffer(w.readframes(50000), dtype="<i2").reshape(-1, 2)
    assert np.array_equal(x, y)
    with pytest.raises(ValueError):
        wav_io.read_wav_pcm(src, sample_rate=48000)
