"""Comparisons against the reference-generated golden vectors of the input FX normaliser and the reverbs
(tests/golden/normalizer.npz, reverbs.npz; oracle/make_golden_norm.py).  The same checks run with the CPU oracle (`-m "not gpu"`,
tight tolerance: the oracle restates the reference) and with the CUDA path (`-m gpu`, the float32 tolerance of the kernels)."""
import numpy as np

from oracle import fixtures

ORDERS = {"y_eq": ["eq"], "y_chain": ["loudness", "eq", "imager", "loudness"], "y_comp": ["compression"]}


def normalizer_features(g, smooth):
    """features_mean as the normaliser holds it.  `smooth`: the oracle takes the already smoothed EQ target, the product (like the
    reference class) smooths it itself."""
    from oracle import norm_oracle as N
    eq = g["eq_drums"]
    return {"eq": {"drums": N.smooth_eq_feature(eq, "drums") if smooth else eq.copy()},
            "loudness": {"drums": g["loudness_drums"]}, "imager": {"drums": g["imager_drums"]},
            "compression": {"drums": g["compression_drums"]}}


def check_normalizer(normalize, rel_tol, smooth):
    """normalize(x [n, 2] float32, order, features) -> [n, 2]"""
    g = fixtures.load_golden("normalizer.npz")
    feats = normalizer_features(g, smooth)
    for key, order in ORDERS.items():
        ref = g[key].astype(np.float64)
        got = np.asarray(normalize(g["x"].copy(), order, feats), dtype=np.float64)
        assert got.shape == ref.shape, key
        rms = np.sqrt(np.mean((got - ref) ** 2)) / np.sqrt(np.mean(ref ** 2))
        assert rms <= rel_tol and np.abs(got - ref).max() <= rel_tol * max(1.0, np.abs(ref).max()) * 4, (key, rms, np.abs(got - ref).max())


def check_reverbs(algo, conv, tol):
    """algo(x, room_size, damping, dry_mix, wet_mix, width) -> [n, 2];  conv(x, h, pre_delay_ms, wet, dry) -> [n, 2]"""
    g = fixtures.load_golden("reverbs.npz")
    x = g["x"]
    ref = g["y_algo"].astype(np.float64)
    got = np.asarray(algo(x.copy(), *[float(v) for v in g["algo_params"]]), dtype=np.float64)
    assert got.shape == ref.shape and np.abs(got - ref).max() <= tol * max(1.0, np.abs(ref).max()), np.abs(got - ref).max()
    for tag in ("mono", "stereo"):
        ref = g[f"y_conv_{tag}"].astype(np.float64)
        got = np.asarray(conv(x.copy(), g[f"h_{tag}"], 3, 0.7, 0.4), dtype=np.float64)
        assert got.shape == ref.shape and np.abs(got - ref).max() <= 4 * tol * np.abs(ref).max(), (tag, np.abs(got - ref).max())
