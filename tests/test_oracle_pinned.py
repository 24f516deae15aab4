"""Pins the oracle (oracle/networks_oracle.py, oracle/fx_oracle.py) -- CPU only.

 (1) against the committed golden vectors in tests/golden/ (outputs of the UNMODIFIED reference, produced by
     oracle/make_golden.py in the build container), and
 (2) against the reference modules themselves when /root/reference is present (it is not on the GPU box).
"""
import os

import numpy as np
import pytest
import torch

from oracle import fixtures, fx_oracle, networks_oracle as O, ref_import, weights as W

torch.set_num_threads(8)
ENC_SD = None
TCN_SD = None


def sds():
    global ENC_SD, TCN_SD
    if ENC_SD is None:
        ENC_SD, TCN_SD = W.make_encoder_state_dict(0), W.make_tcn_state_dict(0)
    return ENC_SD, TCN_SD


def test_golden_encoder():
    esd, _ = sds()
    x = W.synthetic_audio(2, 32768, seed=11)
    with torch.no_grad():
        emb = O.fxencoder_forward(x, esd, W.ENC_KERNELS, W.ENC_STRIDES).numpy()
    ref = fixtures.load_golden("enc_small.npz")["emb"]
    assert emb.shape == ref.shape == (2, 2048)
    assert np.abs(emb - ref).max() <= 1e-6 * max(1.0, np.abs(ref).max())


def test_golden_tcn_small_and_percond():
    _, tsd = sds()
    with torch.no_grad():
        y = O.tcn_forward(W.synthetic_audio(2, 8191, seed=12), fixtures.make_cond(1, 21), tsd).numpy()
        y2 = O.tcn_forward(W.synthetic_audio(3, 4099, seed=13), fixtures.make_cond(3, 22), tsd).numpy()
    assert np.abs(y - fixtures.load_golden("tcn_small.npz")["y"]).max() <= 2e-6
    assert np.abs(y2 - fixtures.load_golden("tcn_percond.npz")["y"]).max() <= 2e-6


def test_golden_real_audio_windows():
    """The oracle on REAL audio (windows of samples/style_transfer/#0 decoded by the reference's loader) against the
    reference's outputs on the same windows (tests/golden/real_audio.npz, oracle/make_golden_real.py)."""
    esd, tsd = sds()
    g = fixtures.load_golden("real_audio.npz")
    x = torch.from_numpy((g["x_vocals"].astype(np.float64).T / 32768.0).astype(np.float32))[None]
    with torch.no_grad():
        emb = O.fxencoder_forward(x, esd, W.ENC_KERNELS, W.ENC_STRIDES)[0].numpy()
        y = O.tcn_forward(x, torch.from_numpy(g["cond_vocals"])[None], tsd)[0].numpy()
    assert np.abs(emb - g["emb_vocals"]).max() <= 1e-6 * max(1.0, np.abs(g["emb_vocals"]).max())
    assert np.abs(y - g["y_vocals"]).max() <= 2e-6
    P = g["fx_params"]
    for i in range(2):
        xi = (g[f"fx_x{i}"].astype(np.float64) / 32768.0).astype(np.float32)
        yi = fx_oracle.fx_chain(xi, P[i])
        assert np.sqrt(np.mean((yi - g[f"fx_y{i}"]) ** 2)) <= 1e-7, i


def test_golden_tcn_blocks():
    _, tsd = sds()
    g = fixtures.load_golden("tcn_blocks.npz")
    cond = fixtures.make_cond(1, 24)
    for n in (0, 1, 4, 9, 13):
        gen = torch.Generator()
        gen.manual_seed(300 + n)
        xb = torch.randn(1, 2 if n == 0 else 128, fixtures.BLOCK_LEN, generator=gen) * 0.5
        with torch.no_grad():
            y = O.tcn_block(xb, cond, tsd, f"blocks.{n}", 15, 2 ** n)[0, ::fixtures.BLOCK_CH_STRIDE].numpy()
        assert np.abs(y - g[f"b{n}"]).max() <= 2e-5, n


def test_golden_fx_chain():
    g = fixtures.load_golden("fx_chain.npz")
    P = g["params"]
    assert np.array_equal(P, fx_oracle.random_params(3, seed=77))
    for i in range(3):
        y = fx_oracle.fx_chain(fixtures.fx_input(i, 16000), P[i])
        d = y.astype(np.float64) - g[f"y{i}"]
        assert np.sqrt(np.mean(d ** 2)) <= 1e-6, i


def test_receptive_field_and_segmentation_quirks():
    assert O.compute_receptive_field() == 229363                      # inference/configs.yaml:22 (5.2 s)
    song = torch.arange(2 * 1000, dtype=torch.float32).reshape(2, 1000)
    b = O.batchwise_segmentization(song, 250, 3)                      # exact multiple -> one extra all-zero segment (q1)
    assert [t.shape[0] for t in b] == [3, 2] and float(b[-1][-1].abs().sum()) == 0.0
    b = O.batchwise_segmentization(song, 300, 8)
    assert b[0].shape == (4, 2, 300) and float(b[0][3, :, 100:].abs().sum()) == 0.0
    with pytest.raises(AssertionError):
        O.batchwise_segmentization(song, 2000, 1)


needs_ref = pytest.mark.skipif(not ref_import.reference_available(), reason="/root/reference not present (GPU box)")


@needs_ref
def test_oracle_equals_reference_modules():
    esd, tsd = sds()
    enc, tcn = ref_import.build_reference_models(esd, tsd)
    assert len(enc.state_dict()) == 168 and len(tcn.state_dict()) == 128
    assert sum(p.numel() for p in enc.parameters()) == 81392682
    assert sum(p.numel() for p in tcn.parameters()) == 10547970
    assert tcn.compute_receptive_field() == O.compute_receptive_field()
    x = W.synthetic_audio(2, 20011, seed=5)
    with torch.no_grad():
        assert torch.equal(enc(x), O.fxencoder_forward(x, esd, W.ENC_KERNELS, W.ENC_STRIDES))
        cond = fixtures.make_cond(2, 6)
        assert torch.equal(tcn(x, cond), O.tcn_forward(x, cond, tsd))
        assert torch.equal(tcn(x, [cond] * 14), O.tcn_forward(x, [cond] * 14, tsd))


@needs_ref
def test_fx_oracle_equals_reference_code():
    from oracle.make_golden import reference_fx_chain
    ca = ref_import.import_reference_fx()
    P = fx_oracle.random_params(4, seed=3)
    for i in range(4):
        x = fixtures.fx_input(10 + i, 12000)
        y_ref = reference_fx_chain(ca, P[i])([x.copy()])[0]
        d = y_ref.astype(np.float64) - fx_oracle.fx_chain(x, P[i])
        assert np.sqrt(np.mean(d ** 2)) <= 1e-7


@needs_ref
def test_io_oracle_equals_reference_loader():
    """oracle/io_oracle.decode == the reference's load_wav_segment (+ stem clamp, float cast) on the reference's own sample
    stems (stereo int16, read with the stdlib `wave` module exactly as loader_utils.py:47-70 does)."""
    import glob
    import wave
    from oracle import io_oracle
    lu = ref_import.import_reference_loader_utils()
    paths = sorted(glob.glob(os.path.join(ref_import.REFERENCE_ROOT, "samples", "style_transfer", "*", "separated", "*", "input", "*.wav")))[:3]
    assert paths, "reference sample stems not found"
    for p in paths:
        n = 150000
        ref = np.clip(lu.load_wav_segment(p, start_point=1000, duration=n, axis=0), -1.0, 1.0).astype(np.float32)
        with wave.open(p, "r") as w:
            w.setpos(1000)
            raw = np.frombuffer(w.readframes(n), dtype="<i2").reshape(-1, w.getnchannels())
        got = io_oracle.decode(raw)
        assert got.shape == ref.shape and np.array_equal(got, ref), p


def test_io_oracle_encode_known_answers():
    """PCM_16 quantisation of the remix: ties to even, clip on both sides, float32 stem sum in order."""
    from oracle import io_oracle
    s = np.zeros((2, 2, 6), np.float32)
    s[0, :, 0] = 0.5 / 32768; s[0, :, 1] = 1.5 / 32768; s[0, :, 2] = -2.5 / 32768
    s[:, :, 3] = 0.75                      # 1.5 -> clip to 32767
    s[:, :, 4] = -0.75                     # -1.5 -> clip to -32768
    s[0, 0, 5] = 1e-3; s[1, 0, 5] = 2e-3
    out = io_oracle.encode_mix(s)
    assert out.dtype == np.int16 and out.shape == (6, 2)
    assert out[:5, 0].tolist() == [0, 2, -2, 32767, -32768]
    assert out[5, 0] == int(np.rint(float(np.float32(1e-3) + np.float32(2e-3)) * 32768)) and out[5, 1] == 0
    assert io_oracle.encode_mix(s, 2).shape == (2, 2)


@needs_ref
def test_configs_yaml_equals_reference():
    """inference/configs.yaml is written in a different layout but must parse to the reference's dictionaries."""
    import yaml
    ours = yaml.full_load(open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))),
                                            "music_mixing_style_transfer_b200", "inference", "configs.yaml")))
    ref = yaml.full_load(open(os.path.join(ref_import.REFERENCE_ROOT, "inference", "configs.yaml")))
    assert ours == ref


@pytest.mark.skipif(not ref_import.reference_available(), reason="reference tree not present (GPU box)")
def test_normalizer_oracle_equals_reference_code(tmp_path):
    """oracle/norm_oracle.py against the reference's own lufs_normalize, normalize_imager and
    Audio_Effects_Normalizer.normalize_audio (loudness, imager) on a real stem window.  The BS.1770 meter underneath is the
    restated pyloudnorm (third-party, absent) on BOTH sides: that part stays unpinned."""
    from oracle import norm_oracle as N
    dn, fx_utils, nimg = ref_import.import_reference_normalizer()
    g = fixtures.load_golden("real_audio.npz")
    x = (g["x_drums"].astype(np.float64) / 32768.0).astype(np.float32)[:50000]          # [n, 2]
    x[:, 1] = 0.3 * x[:, 1] + 0.6 * np.roll(x[:, 0], 4410)                             # wide (no Haas branch), unbalanced
    ref = fx_utils.lufs_normalize(x.copy(), 44100, np.array([-28.9]), log=False)
    assert np.abs(N.lufs_normalize(x.copy(), 44100, np.array([-28.9])) - ref).max() <= 1e-9
    ref = nimg.normalize_imager(x.copy(), target_side_mid_bal=0.9447, mono_threshold=0.975, sr=44100)
    assert np.abs(N.normalize_imager(x.copy(), 0.9447, 0.975) - ref).max() <= 1e-7
    feats = {"loudness": {"drums": np.array([-28.9674596])}, "imager": {"drums": np.float32(0.94471526)}}
    np.save(tmp_path / "feats.npy", feats, allow_pickle=True)
    order = ['loudness', 'imager', 'loudness']
    norm = dn.Audio_Effects_Normalizer(str(tmp_path / "feats.npy"), STEMS=['drums'], EFFECTS=order)
    ref = norm.normalize_audio(x.copy(), src='drums')
    got = N.normalize_audio(x.copy(), order, feats, src='drums')
    assert ref.shape == got.shape and np.abs(got - ref).max() <= 1e-7


@pytest.mark.skipif(not ref_import.reference_available(), reason="reference tree not present (GPU box)")
def test_eq_matching_oracle_equals_reference_code(tmp_path, monkeypatch):
    """oracle/norm_oracle.get_eq_matching and the 'eq' effect of normalize_audio against the reference's own get_eq_matching /
    Audio_Effects_Normalizer on a real stem window and the reference's own target spectra (weights/*.npy).  librosa.stft and
    the BS.1770 meter are restated third-party code on both sides (unpinned); firwin2 / filtfilt / savgol are the real scipy.
    scipy >= 1.12 no longer accepts the `nyq=None` keyword the reference passes to firwin2: it is dropped by a wrapper."""
    import scipy.signal
    from oracle import norm_oracle as N
    dn, fx_utils, nimg = ref_import.import_reference_normalizer()
    import utils_data_normalization as udn
    real_firwin2 = scipy.signal.firwin2
    monkeypatch.setattr(scipy.signal, "firwin2", lambda *a, nyq=None, **k: real_firwin2(*a, **k))
    feats = np.load(os.path.join(ref_import.REFERENCE_ROOT, "weights", "musdb18_fxfeatures_eqcompimagegain.npy"), allow_pickle=True)[()]
    g = fixtures.load_golden("real_audio.npz")
    x = (g["x_drums"].astype(np.float64) / 32768.0).astype(np.float32)[:40000]
    spec = N.smooth_eq_feature(feats['eq']['drums'], 'drums')
    xs = np.pad(x[:, 0], (65536, 65536))
    ref = udn.get_eq_matching(xs, spec, sr=44100, n_fft=65536, hop_length=16384, min_db=-40, ntaps=1001, lufs=-30)
    assert np.abs(N.get_eq_matching(xs, spec) - ref).max() <= 1e-12
    assert np.array_equal(N.get_eq_matching(xs * 1e-4, spec), xs * 1e-4)                 # below min_db: untouched
    sub = {"eq": {"drums": feats['eq']['drums'].copy()}, "loudness": {"drums": feats['loudness']['drums']}}
    np.save(tmp_path / "feats.npy", sub, allow_pickle=True)
    order = ['loudness', 'eq', 'loudness']
    norm = dn.Audio_Effects_Normalizer(str(tmp_path / "feats.npy"), STEMS=['drums'], EFFECTS=order)
    ref = norm.normalize_audio(x.copy(), src='drums')
    got = N.normalize_audio(x.copy(), order, {"eq": {"drums": spec}, "loudness": sub["loudness"]}, src='drums')
    assert ref.shape == got.shape and np.abs(got - ref).max() <= 1e-7


@pytest.mark.skipif(not ref_import.reference_available(), reason="reference tree not present (GPU box)")
def test_comp_matching_oracle_equals_reference_code(tmp_path):
    """oracle/norm_oracle.get_comp_matching / get_mean_peak against the reference's own functions (the reference's numba
    compressor, its search loops and break conditions) on a real drum stem.  aubio (third-party C library, absent) is the
    stand-in detector of oracle/shims/aubio on both sides, pyloudnorm.normalize.peak the restated one-liner."""
    from oracle import norm_oracle as N
    dn, fx_utils, nimg = ref_import.import_reference_normalizer()
    import utils_data_normalization as udn
    g = fixtures.load_golden("real_audio.npz")
    x = (g["x_drums"].astype(np.float64) / 32768.0).astype(np.float32)[:40000]
    xs = np.pad(x[:, 0], (65536, 65536))
    k = 10 ** (-10.0 / 20) / np.abs(xs).max()
    assert np.allclose(udn.get_mean_peak((k * xs)[:, None], 44100, n_mels=128, true_peak=False, percentile=75),
                       N.get_mean_peak((k * xs)[:, None]), rtol=0, atol=1e-12)
    for ref_peak, ref_std, changed in ((-13.53084647, 1.12951587, True), (-5.0, 1.0, False), (-10.3, 0.5, False)):
        ref = udn.get_comp_matching(xs, ref_peak, ref_std, 6, 10.0, 180.0, sr=44100, min_db=-40, min_th=-40, comp_peak_norm=-10.0,
                                    max_ratio=20, n_mels=128, true_peak=False, percentile=75, expander=False)
        got = N.get_comp_matching(xs, ref_peak, ref_std, 6, 10.0, 180.0)
        assert ref.shape == got.shape and np.abs(ref - got).max() <= 1e-7
        assert (np.abs(ref[:, 0] - k * xs).max() > 1e-3) == changed
    feats = {"compression": {"drums": np.array([-13.53084647, 1.12951587])}}
    np.save(tmp_path / "feats.npy", feats, allow_pickle=True)
    norm = dn.Audio_Effects_Normalizer(str(tmp_path / "feats.npy"), STEMS=['drums'], EFFECTS=['compression'])
    x2 = x.copy()
    x2[:, 1] *= 0.5
    ref = norm.normalize_audio(x2.copy(), src='drums')
    got = N.normalize_audio(x2.copy(), ['compression'], feats, src='drums')
    assert np.abs(got - ref).max() <= 1e-7


@needs_ref
def test_reverb_oracles_and_factory_equal_reference_code():
    """(1) oracle/fx_oracle.algorithmic_reverb (block-wise comb / all-pass, the decomposition of the GPU kernel)
    against the reference's AlgorithmicReverb class -- its network wiring, the comb-5 overwrite, the mix -- running on the
    restated pymixconsole Comb / Allpass sample loops (third-party, unpinned); (2) convolutional_reverb against the reference's
    ConvolutionalReverb.process (scipy oaconvolve); (3) the per-instrument factory's chain structure against the reference's."""
    import sys
    ca = ref_import.import_reference_fx()
    x = fixtures.fx_input(3, 20000)
    r = ca.AlgorithmicReverb(sample_rate=44100)
    for name, v in (("room_size", 0.6), ("damping", 0.3), ("dry_mix", 0.8), ("wet_mix", 0.35), ("width", 0.6)):
        getattr(r.parameters, name).value = v
    r.update(None)
    ref = r.process(x.copy())
    assert np.abs(ref - fx_oracle.algorithmic_reverb(x, 0.6, 0.3, 0.8, 0.35, 0.6)).max() <= 1e-12
    rng = np.random.RandomState(1)
    for m, ch in ((3000, 1), (9000, 2)):
        h = (rng.randn(m, ch) * np.exp(-np.arange(m) / (m / 6.0))[:, None]).astype(np.float32)
        h[37] *= 8.0                                                       # the peak the cut index is taken from
        cr = ca.ConvolutionalReverb([[{'impulse_response': lambda h=h: h}]], 44100)
        cr.parameters.wet.value, cr.parameters.dry.value, cr.parameters.pre_delay.value = 0.7, 0.4, 3
        cr.update()
        ref = cr.process(x.copy())
        assert np.abs(ref - fx_oracle.convolutional_reverb(x, h, 3, 0.7, 0.4)).max() <= 1e-6
    # factory structure: the reference module imports soundfile / librosa (stand-ins) through common_dataprocessing
    import types
    sys.modules.setdefault("soundfile", types.ModuleType("soundfile"))
    import importlib
    ref_chain = importlib.import_module("audio_effects_chain")
    from music_mixing_style_transfer_b200.mixing_manipulator import create_inst_effects_augmentation_chain
    from test_host_logic import _chain_structure
    prob = {"eq": 0.9, "comp": 0.8, "pan": 0.7, "imager": 0.6, "reverb": 0.5, "gain": 1.0}
    for inst in ("drums", "bass"):
        a = _chain_structure(ref_chain.create_inst_effects_augmentation_chain(inst, prob, algorithmic=True))
        b = _chain_structure(create_inst_effects_augmentation_chain(inst, prob, algorithmic=True))
        assert a == b, inst


def test_oracles_match_normalizer_and_reverb_goldens():
    """The CPU oracles against the committed reference outputs (runs everywhere: the goldens travel, /root/reference does not).
    float32 storage of the goldens bounds the agreement at ~1e-7."""
    import golden_checks
    from oracle import norm_oracle as N
    golden_checks.check_normalizer(lambda x, order, feats: N.normalize_audio(x, order, feats, src="drums"), rel_tol=2e-7, smooth=True)
    golden_checks.check_reverbs(fx_oracle.algorithmic_reverb, fx_oracle.convolutional_reverb, tol=2e-7)
