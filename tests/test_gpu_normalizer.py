"""Input FX normaliser (loudness, imager) and the panner / Haas processors on the GPU against
oracle/norm_oracle.py, which tests/test_oracle_pinned.py pins to the reference's own data_normalization / fx_utils /
normalization_imager code (the BS.1770 meter underneath is pyloudnorm's: third-party, restated, unpinned)."""
import numpy as np
import pytest
import torch

from gpu_helpers import err_stats
from oracle import fixtures, norm_oracle as N

pytestmark = pytest.mark.gpu
FEATS = {"loudness": {"drums": np.array([-28.9674596]), "bass": np.array([-24.37411392])},
         "imager": {"drums": np.float32(0.94471526), "bass": np.float32(0.9816045)}}


def _stem(n=60000, wide=True):
    g = fixtures.load_golden("real_audio.npz")
    x = (g["x_drums"].astype(np.float64) / 32768.0).astype(np.float32)[:n]          # [n, 2], real drum transients
    if wide:
        x[:, 1] = 0.3 * x[:, 1] + 0.6 * np.roll(x[:, 0], 4410)
    return np.ascontiguousarray(x)


def test_stereo_building_blocks():
    from music_mixing_style_transfer_b200.mixing_manipulator import data_normalization as dn
    x = _stem(50001)                                                               # odd length: scalar path
    xt = torch.from_numpy(np.ascontiguousarray(x.T))[None].cuda()
    st = dn.stereo_stats(xt)[0]
    x64 = x.astype(np.float64)
    ref = [np.sum(x64[:, 0] ** 2), np.sum(x64[:, 1] ** 2), np.sum(x64[:, 0] * x64[:, 1]), np.abs(x).max()]
    assert np.allclose(st, ref, rtol=1e-6, atol=0) and st[3] == ref[3]
    M = np.array([0.7, -0.2, 0.1, 1.3], np.float32)
    y = dn.stereo_mix(xt, M[None])[0].cpu().numpy()
    assert np.allclose(y[0], M[0] * x[:, 0] + M[1] * x[:, 1], atol=1e-7) and np.allclose(y[1], M[2] * x[:, 0] + M[3] * x[:, 1], atol=1e-7)
    for delay, ch in ((1764, 'left'), (-333, 'right'), (0, 'left')):
        y = dn.haas(xt, [delay], [0.4], [0 if ch == 'left' else 1])[0].cpu().numpy().T
        assert np.abs(y - N.haas_process(x, delay, np.float32(0.4), ch)).max() <= 1e-7, (delay, ch)


@pytest.mark.parametrize("order", [['loudness'], ['imager'], ['loudness', 'imager', 'loudness']])
def test_normalizer_matches_oracle(order):
    from music_mixing_style_transfer_b200.mixing_manipulator import Audio_Effects_Normalizer
    x = _stem()
    norm = Audio_Effects_Normalizer(FEATS, STEMS=['drums', 'bass'], EFFECTS=order)
    ref = N.normalize_audio(x.copy(), order, FEATS, src='drums')
    got_np = norm.normalize_audio(x.copy(), src='drums')                            # the reference's call: numpy [n, 2]
    got_t = norm.normalize_audio(torch.from_numpy(np.ascontiguousarray(x.T)).cuda(), src='drums')   # the engine's call
    assert got_np.shape == ref.shape and np.array_equal(got_np, got_t.cpu().numpy().T)
    e = err_stats(got_np.T, np.asarray(ref, np.float64).T)
    # float32 K-weighting scan vs float64 lfilter, float64 vs float32 energy sums: ~1e-6 relative on the gains
    assert e["rel"] <= 2e-5 and e["max"] <= 2e-5 * max(1.0, np.abs(ref).max()), (order, e)


def test_normalizer_gates_and_haas_branch():
    from music_mixing_style_transfer_b200.mixing_manipulator import Audio_Effects_Normalizer
    norm = Audio_Effects_Normalizer(FEATS, STEMS=['drums', 'bass'], EFFECTS=['imager'])
    quiet = _stem(30000) * 1e-3                                                     # peak below -40 dB: untouched (:102-103)
    assert np.array_equal(norm.normalize_audio(quiet.copy(), src='drums'), quiet)
    mono = _stem(40000, wide=False)
    mono[:, 1] = mono[:, 0]                                                         # exactly mono -> Haas, then the balances
    norm.haas_rng = np.random.RandomState(5)
    y = norm.normalize_audio(mono.copy(), src='drums').astype(np.float64)
    mid, side = y[:, 0] + y[:, 1], y[:, 0] - y[:, 1]
    bal = np.sum(mid ** 2) / (np.sum(mid ** 2) + np.sum(side ** 2))
    assert np.isfinite(y).all() and abs(bal - float(FEATS["imager"]["drums"])) < 2e-2, bal
    with pytest.raises(NotImplementedError):
        Audio_Effects_Normalizer(FEATS, EFFECTS=['loudness', 'eq'])


def test_panner_and_haas_processors_in_a_chain():
    from music_mixing_style_transfer_b200.mixing_manipulator import Haas, Panner, create_effects_augmentation_chain
    x = _stem(20000)
    p = Panner()
    for law in ('-4.5dB', 'linear', 'constant_power'):
        p.parameters.pan.value, p.parameters.pan_law.value = 0.3, law
        p.update()
        assert np.abs(p.process(x) - x * N.pan_gains(0.3, law)).max() <= 1e-7, law
    h = Haas(sample_rate=44100)
    h.parameters.delay.value, h.parameters.feedback.value, h.parameters.wet_channel.value = 900, 0.5, 'right'
    assert np.abs(h.process(x) - N.haas_process(x, 900, np.float32(0.5), 'right')).max() <= 1e-7
    np.random.seed(3)
    chain = create_effects_augmentation_chain([('pan', 1.0), ('imager', 1.0)], sample_rate=44100)
    y = chain([x.copy()])[0]
    assert y.shape == x.shape and np.isfinite(y).all()
    # both effects are RMS re-normalised by the chain (common_audioeffects.py:142-145)
    assert abs(np.sqrt(np.mean(y ** 2)) / np.sqrt(np.mean(x ** 2)) - 1.0) < 1e-4
