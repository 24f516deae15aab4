"""SURVEY.md 8f-4 on the GPU: algorithmic reverb (csrc/reverb.cu), convolutional reverb (mst_fft_convolve in csrc/spectral.cu),
one-shelf equalisers and the per-instrument chain factory, against oracle/fx_oracle.py -- which tests/test_oracle_pinned.py pins
to the reference's own classes (the comb / all-pass sample loops underneath are restated pymixconsole: unpinned)."""
import numpy as np
import pytest
import torch

from gpu_helpers import err_stats
from oracle import fixtures, fx_oracle

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("n,params", [(30000, (0.6, 0.3, 0.8, 0.35, 0.6)), (50001, (0.85, 0.0, 0.0, 1.0, 1.0)),
                                      (1000, (0.05, 1.0, 0.9, 0.1, 0.0)), (12345, (0.5, 0.1, 0.9, 0.1, 0.7))])
def test_algorithmic_reverb_matches_oracle(n, params):
    from music_mixing_style_transfer_b200.mixing_manipulator import AlgorithmicReverb
    x = fixtures.fx_input(5, n)
    r = AlgorithmicReverb(sample_rate=44100)
    for name, v in zip(("room_size", "damping", "dry_mix", "wet_mix", "width"), params):
        getattr(r.parameters, name).value = v
    got = r.process(x.copy())
    ref = fx_oracle.algorithmic_reverb(x, *params)
    e = err_stats(got.T, ref.T)
    # float32 delay lines with feedback <= 0.85 against the float64 oracle
    assert got.shape == ref.shape and got.dtype == np.float64 and e["rel"] <= 1e-5 and e["max"] <= 1e-5 * max(1.0, np.abs(ref).max()), e
    mono = r.process(x[:, :1].copy())                                 # mono in: both sides read channel 0 (:1448-1456)
    assert np.abs(mono - fx_oracle.algorithmic_reverb(np.repeat(x[:, :1], 2, axis=1), *params)).max() <= 1e-5 * max(1.0, np.abs(ref).max())


def test_algorithmic_reverb_batch_is_per_segment():
    """The C entry takes [B, 2, L] with per-segment parameters: segments must not see each other."""
    from music_mixing_style_transfer_b200 import _cabi
    lib = _cabi.lib()
    B, L = 5, 20000
    x = torch.from_numpy(np.ascontiguousarray(np.stack([fixtures.fx_input(20 + i, L).T for i in range(B)]))).cuda()   # np.stack keeps the F order of the views
    rng = np.random.RandomState(0)
    P = np.stack([rng.uniform(0.05, 0.85, B), rng.rand(B), rng.rand(B), rng.rand(B), rng.rand(B)], 1).astype(np.float32)
    p = torch.from_numpy(P).cuda()
    ws = torch.empty(lib.mst_algo_reverb_workspace_bytes(B, L), dtype=torch.uint8, device="cuda")
    y = torch.empty_like(x)
    _cabi.check(lib.mst_algo_reverb(_cabi.ptr(x), _cabi.ptr(p), _cabi.ptr(y), B, L, _cabi.ptr(ws), ws.numel(), _cabi.current_stream()), "algo_reverb")
    y = y.cpu().numpy()
    for i in (0, B - 1):
        ref = fx_oracle.algorithmic_reverb(np.ascontiguousarray(x[i].cpu().numpy().T), *[float(v) for v in P[i]]).T
        assert np.abs(y[i] - ref).max() <= 1e-5 * max(1.0, np.abs(ref).max()), i


@pytest.mark.parametrize("m,ch,n", [(3000, 1, 20000), (9000, 2, 50001), (70000, 2, 40000), (200, 1, 700)])
def test_convolutional_reverb_matches_oracle(m, ch, n):
    from music_mixing_style_transfer_b200.mixing_manipulator import ConvolutionalReverb
    x = fixtures.fx_input(7, n)
    rng = np.random.RandomState(m)
    h = (rng.randn(m, ch) * np.exp(-np.arange(m) / (m / 6.0))[:, None]).astype(np.float32)
    h[min(37, m - 1)] *= 8.0
    cr = ConvolutionalReverb([[{'impulse_response': lambda: h}]], 44100)
    cr.parameters.wet.value, cr.parameters.dry.value, cr.parameters.pre_delay.value = 0.7, 0.4, 3
    cr.update()
    got = cr.process(x.copy())
    ref = fx_oracle.convolutional_reverb(x, h, 3, 0.7, 0.4)
    e = err_stats(got.T, np.asarray(ref, np.float64).T)
    # float32 FFT of up to 2^16 points on both sides (scipy's oaconvolve transforms float32 input in single precision too)
    assert got.shape == ref.shape and e["rel"] <= 2e-5 and e["max"] <= 2e-5 * np.abs(ref).max(), e
    cr.parameters.decay.value = 0.5                                   # fade-out of the response (:719-733)
    cr.update()
    assert cr.h.shape[0] < m or m <= 200
    cr.parameters.wet.value = 0.0
    assert np.array_equal(cr.process(x.copy()), x)


def test_one_shelf_equaliser_and_instrument_chain():
    import scipy.signal
    from music_mixing_style_transfer_b200.mixing_manipulator import (Equaliser, Parameter, ParameterList,
                                                                      create_inst_effects_augmentation_chain)
    x = fixtures.fx_input(9, 30000)
    for band, kind in (("high_shelf", "high_shelf"), ("low_shelf", "low_shelf")):
        params = ParameterList()
        params.add(Parameter(band + '_gain', -50.0, 'float', minimum=-50.0, maximum=-50.0))
        params.add(Parameter(band + '_freq', 100.0, 'float', minimum=100.0, maximum=100.0))
        eq = Equaliser(n_channels=2, sample_rate=44100, bands=[band], parameters=params)
        got = eq.process(x.copy())
        b, a = fx_oracle.rbj_biquad(-50.0, 0.707, 100.0, 44100, kind)
        ref = scipy.signal.lfilter(b, a, x.astype(np.float64), axis=0).astype(np.float32)
        assert np.abs(got - ref).max() <= 2e-6 * max(1.0, np.abs(ref).max()), band
    prob = {"eq": 1.0, "comp": 1.0, "pan": 1.0, "imager": 1.0, "reverb": 1.0, "gain": 1.0}
    for inst in ("drums", "vocals"):
        np.random.seed(4)
        chain = create_inst_effects_augmentation_chain(inst, prob, algorithmic=True)
        y = chain([x.copy()])[0]
        assert y.shape == x.shape and np.isfinite(y).all() and np.abs(y).max() > 1e-3, inst


def test_reverbs_against_reference_golden():
    """Both reverbs against outputs of the reference's own classes (tests/golden/reverbs.npz)."""
    import golden_checks
    from music_mixing_style_transfer_b200.mixing_manipulator import AlgorithmicReverb, ConvolutionalReverb

    def algo(x, room_size, damping, dry_mix, wet_mix, width):
        r = AlgorithmicReverb(sample_rate=44100)
        for name, v in zip(("room_size", "damping", "dry_mix", "wet_mix", "width"), (room_size, damping, dry_mix, wet_mix, width)):
            getattr(r.parameters, name).value = v
        return r.process(x)

    def conv(x, h, pre_delay_ms, wet, dry):
        cr = ConvolutionalReverb([[{'impulse_response': lambda: h}]], 44100)
        cr.parameters.wet.value, cr.parameters.dry.value, cr.parameters.pre_delay.value = wet, dry, pre_delay_ms
        cr.update()
        return cr.process(x)
    golden_checks.check_reverbs(algo, conv, tol=1e-5)
