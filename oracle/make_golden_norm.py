"""TEST INFRASTRUCTURE ONLY -- golden vectors of the input FX normaliser and the reverbs: the UNMODIFIED
reference code run on a window of its own sample stems.

    python oracle/make_golden_norm.py        (this container only; /root/reference does not exist on the GPU box)

tests/golden/normalizer.npz
  x                 the 40,000-frame drums window of real_audio.npz made wide (so that the randomised Haas branch stays off)
  eq_drums ...      the reference's own targets for the drums stem (weights/musdb18_fxfeatures_eqcompimagegain.npy)
  y_eq, y_chain     Audio_Effects_Normalizer(EFFECTS=['eq']) / (['loudness', 'eq', 'imager', 'loudness']).normalize_audio(x, 'drums')
  y_comp            EFFECTS=['compression'] with the stand-in onset detector of oracle/shims/aubio (aubio is not installed)
tests/golden/reverbs.npz
  x, algo_params, y_algo        AlgorithmicReverb.process on restated pymixconsole comb / all-pass loops (third-party: UNPINNED)
  h_mono, h_stereo, y_conv_*    ConvolutionalReverb.process (scipy oaconvolve), wet 0.7, dry 0.4, pre_delay 3 ms
Third-party code that is not in the reference tree (pyloudnorm, librosa.stft, aubio, pymixconsole components) runs as the
restatements of oracle/shims on the reference's side, see the module headers there; scipy >= 1.12 rejects the `nyq=None`
keyword the reference passes to firwin2, which is dropped by a wrapper.
"""
import os
import sys

import numpy as np
import scipy.signal

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import fixtures, ref_import  # noqa: E402
from oracle.fixtures import GOLDEN_DIR as OUT  # noqa: E402


def main():
    dn, fx_utils, nimg = ref_import.import_reference_normalizer()
    real_firwin2 = scipy.signal.firwin2
    scipy.signal.firwin2 = lambda *a, nyq=None, **k: real_firwin2(*a, **k)
    feats = np.load(os.path.join(ref_import.REFERENCE_ROOT, "weights", "musdb18_fxfeatures_eqcompimagegain.npy"), allow_pickle=True)[()]
    g = fixtures.load_golden("real_audio.npz")
    x = (g["x_drums"].astype(np.float64) / 32768.0).astype(np.float32)[:40000]
    x[:, 1] = 0.3 * x[:, 1] + 0.6 * np.roll(x[:, 0], 4410)
    out = {"x": x, "eq_drums": feats["eq"]["drums"].astype(np.float32), "loudness_drums": np.asarray(feats["loudness"]["drums"], np.float64),
           "imager_drums": np.float32(feats["imager"]["drums"]), "compression_drums": np.asarray(feats["compression"]["drums"], np.float64)}
    import tempfile
    with tempfile.TemporaryDirectory() as tmp:
        def run(order):
            sub = {k: {"drums": np.copy(v["drums"])} for k, v in feats.items() if k in order}
            np.save(os.path.join(tmp, "f.npy"), sub, allow_pickle=True)
            norm = dn.Audio_Effects_Normalizer(os.path.join(tmp, "f.npy"), STEMS=["drums"], EFFECTS=order)
            return norm.normalize_audio(x.copy(), src="drums").astype(np.float32)
        out["y_eq"] = run(["eq"])
        out["y_chain"] = run(["loudness", "eq", "imager", "loudness"])
        out["y_comp"] = run(["compression"])
    np.savez_compressed(os.path.join(OUT, "normalizer.npz"), **out)

    ca = ref_import.import_reference_fx()
    xr = fixtures.fx_input(3, 20000)
    params = (0.6, 0.3, 0.8, 0.35, 0.6)
    r = ca.AlgorithmicReverb(sample_rate=44100)
    for name, v in zip(("room_size", "damping", "dry_mix", "wet_mix", "width"), params):
        getattr(r.parameters, name).value = v
    r.update(None)
    rv = {"x": xr, "algo_params": np.asarray(params, np.float64), "y_algo": r.process(xr.copy()).astype(np.float32)}
    rng = np.random.RandomState(1)
    for tag, m, ch in (("mono", 3000, 1), ("stereo", 9000, 2)):
        h = (rng.randn(m, ch) * np.exp(-np.arange(m) / (m / 6.0))[:, None]).astype(np.float32)
        h[37] *= 8.0
        cr = ca.ConvolutionalReverb([[{"impulse_response": lambda h=h: h}]], 44100)
        cr.parameters.wet.value, cr.parameters.dry.value, cr.parameters.pre_delay.value = 0.7, 0.4, 3
        cr.update()
        rv[f"h_{tag}"] = h
        rv[f"y_conv_{tag}"] = np.asarray(cr.process(xr.copy()), np.float32)
    np.savez_compressed(os.path.join(OUT, "reverbs.npz"), **rv)
    for f in ("normalizer.npz", "reverbs.npz"):
        print(f, os.path.getsize(os.path.join(OUT, f)), "bytes")


if __name__ == "__main__":
    main()
