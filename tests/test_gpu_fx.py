"""FX chain parity on the GPU (EQ -> compressor -> imager -> gain, RMS re-normalisation) vs the numpy oracle + golden."""
import numpy as np
import pytest
import torch

from gpu_helpers import RMS_TOL, err_stats
from oracle import fixtures, fx_oracle

pytestmark = pytest.mark.gpu


def run_gpu(x_list, P, stages=None):
    from music_mixing_style_transfer_b200.mixing_manipulator import FX_ALL, fx_chain_forward
    x = torch.from_numpy(np.stack([np.ascontiguousarray(a.T) for a in x_list])).cuda()
    y = fx_chain_forward(x, torch.from_numpy(P).cuda(), FX_ALL if stages is None else stages)
    return [np.ascontiguousarray(y[i].cpu().numpy().T) for i in range(len(x_list))]


def test_golden_chain():
    g = fixtures.load_golden("fx_chain.npz")
    xs = [fixtures.fx_input(i, 16000) for i in range(3)]
    ys = run_gpu(xs, g["params"])
    for i in range(3):
        e = err_stats(ys[i].T, g[f"y{i}"].T)
        assert e["rms"] <= RMS_TOL and e["rms"] <= 3e-5 and e["rel"] <= 5e-4, (i, e)


@pytest.mark.parametrize("L", [4096, 16000, 44100, 70001])
def test_chain_vs_oracle_random_params(L):
    B = 6
    P = fx_oracle.random_params(B, seed=L)
    xs = [fixtures.fx_input(20 + i, L) for i in range(B)]
    ys = run_gpu(xs, P)
    for i in range(B):
        ref = fx_oracle.fx_chain(xs[i], P[i])
        e = err_stats(ys[i].T, ref.T)
        # absolute gate = the north_star tolerance; the relative bound documents what the float32-local / float64-carry
        # arithmetic delivers (the imager can amplify the ~1e-5 float32 noise of the EQ by its side gain)
        assert e["rms"] <= RMS_TOL and e["rms"] <= 3e-5 and e["rel"] <= 5e-4, (L, i, e, P[i])


def test_single_stages_vs_oracle():
    from music_mixing_style_transfer_b200.mixing_manipulator import FX_COMP, FX_EQ, FX_GAIN, FX_IMAGER, FX_RMSNORM
    P = fx_oracle.random_params(4, seed=5)
    xs = [fixtures.fx_input(40 + i, 30000) for i in range(4)]
    cases = [(FX_EQ, lambda x, p: fx_oracle.equaliser(x, p)),
             (FX_EQ | FX_RMSNORM, lambda x, p: fx_oracle.rms_normalize(x, fx_oracle.equaliser(x, p))),
             (FX_COMP, lambda x, p: fx_oracle.compressor(x, p)),
             (FX_COMP | FX_RMSNORM, lambda x, p: fx_oracle.rms_normalize(x, fx_oracle.compressor(x, p))),
             (FX_IMAGER, lambda x, p: fx_oracle.imager(x, p)),
             (FX_IMAGER | FX_RMSNORM, lambda x, p: fx_oracle.rms_normalize(x, fx_oracle.imager(x, p))),
             (FX_GAIN, lambda x, p: fx_oracle.gain(x, p))]
    for stages, fn in cases:
        ys = run_gpu(xs, P, stages)
        for i in range(4):
            ref = np.asarray(fn(xs[i], P[i]), dtype=np.float32)
            e = err_stats(ys[i].T, ref.T)
            assert e["rms"] <= RMS_TOL and e["rms"] <= 3e-5 and e["rel"] <= 5e-4, (stages, i, e)


def test_list_api_chain_matches_oracle():
    from music_mixing_style_transfer_b200.mixing_manipulator import create_effects_augmentation_chain
    chain = create_effects_augmentation_chain(["eq", "comp", "imager", "gain"])
    chain.randomize_param_value = False
    p = fx_oracle.random_params(1, seed=9)[0]
    eq, comp, im, gn = [f for f, _, _ in chain.fxs]
    from music_mixing_style_transfer_b200.mixing_manipulator.common_audioeffects import COMP_PARAM_NAMES, EQ_PARAM_NAMES
    for j, n in enumerate(EQ_PARAM_NAMES):
        getattr(eq.parameters, n).value = float(p[j])
    for j, n in enumerate(COMP_PARAM_NAMES):
        getattr(comp.parameters, n).value = float(p[13 + j])
    im.parameters.bal.value = float(p[17])
    gn.parameters.gain.value = float(p[18])
    gn.parameters.invert.value = bool(p[19] >= 0.5)
    x = fixtures.fx_input(60, 20000)
    y = chain([x.copy(), x.copy() * 0.5])
    assert len(y) == 2 and y[0].shape == x.shape and y[0].dtype == np.float32
    e = err_stats(y[0].T, fx_oracle.fx_chain(x, p).T)
    assert e["rms"] <= RMS_TOL, e
    e = err_stats(y[1].T, fx_oracle.fx_chain(x * 0.5, p).T)
    assert e["rms"] <= RMS_TOL, e


def test_edge_cases_silence_and_mono():
    """All-zero input (compressor floor -120 dB, RMS guards 1e-7) and L == R (side energy 0, imager 1e-3 guards)."""
    P = fx_oracle.random_params(2, seed=12)
    z = np.zeros((5000, 2), np.float32)
    m = fixtures.fx_input(70, 5000)
    m[:, 1] = m[:, 0]
    ys = run_gpu([z, m], P)
    assert np.all(np.isfinite(ys[0])) and np.abs(ys[0]).max() == 0.0
    e = err_stats(ys[1].T, fx_oracle.fx_chain(m, P[1]).T)
    assert e["rms"] <= RMS_TOL, e


@pytest.mark.parametrize("L", [8192, 8192 * 3 + 4, 8191, 12289])
def test_tile_boundaries(L):
    """fx2.cu tiles: EQ 8192 frames, compressor 4096 frames; lengths at / around the tile size, with L % 4 != 0 (scalar staging
    path) and L % 4 == 0 with a partial last tile (mixed fast / slow tiles)."""
    B = 3
    P = fx_oracle.random_params(B, seed=100 + L)
    xs = [fixtures.fx_input(80 + i, L) for i in range(B)]
    ys = run_gpu(xs, P)
    for i in range(B):
        e = err_stats(ys[i].T, fx_oracle.fx_chain(xs[i], P[i]).T)
        assert e["rms"] <= 3e-5 and e["rel"] <= 5e-4, (L, i, e)


def test_compressor_attack_slower_than_release_and_ratio_cases():
    """Outside the randomisation ranges: attack slower than release (the smoother's max becomes a min), ratio < 1 and ratio == 1
    (common_audioeffects.py:564-573), compressor switched off by threshold 0 / ratio 1 (:635)."""
    from music_mixing_style_transfer_b200.mixing_manipulator import FX_COMP
    x = fixtures.fx_input(90, 30000)
    cases = [(-30.0, 300.0, 20.0, 8.0), (-30.0, 5.0, 100.0, 0.5), (-20.0, 5.0, 100.0, 1.0), (0.0, 5.0, 100.0, 1.0)]
    for thr, att, rel, ratio in cases:
        p = fx_oracle.random_params(1, seed=3)[0]
        p[13], p[14], p[15], p[16] = thr, att, rel, ratio
        y = run_gpu([x], p[None, :], FX_COMP)[0]
        ref = np.asarray(fx_oracle.compressor(x, p), np.float32)
        e = err_stats(y.T, ref.T)
        assert e["rms"] <= 3e-5 and e["rel"] <= 5e-4, ((thr, att, rel, ratio), e)
