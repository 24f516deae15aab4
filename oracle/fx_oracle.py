"""TEST INFRASTRUCTURE ONLY -- numpy/scipy CPU restatement of the reference's FX chain
(EQ -> compressor -> mid/side imager -> gain with RMS re-normalisation), the parity oracle for BASELINE config 3.

Each function cites the reference lines it follows (paths relative to
/root/reference/mixing_style_transfer/mixing_manipulator/).  The compressor / imager / gain / chain logic is pinned
against the UNMODIFIED reference code (common_audioeffects.py imported on the pymixconsole / soxbindings stubs in
oracle/shims) by tests/test_oracle_pinned.py and by the fixtures in tests/golden/.

EQ: the biquad arithmetic lives in the un-vendored PyPI dependency pymixconsole==0.0.1 (requirements.txt:12),
whose source is not under /root/reference.  `equaliser()` restates the published algorithm (RBJ cookbook
biquads, float64 `scipy.signal.lfilter` from zero state; the reference's own docstring common_audioeffects.py:375-376
and comment :518) anchored on the reference call sites :460, :511-519.  ** PARITY UNPINNED for the EQ biquads. **

Array convention = the reference's: float32 [n_samples, n_channels] (time-major, channel-last).
Parameter vector (20 floats per segment; the batched GPU entry point uses the same order):
   0 low_shelf_gain   1 low_shelf_freq
   2 first_band_gain  3 first_band_freq  4 first_band_q
   5 second_band_gain 6 second_band_freq 7 second_band_q
   8 third_band_gain  9 third_band_freq 10 third_band_q
  11 high_shelf_gain 12 high_shelf_freq
  13 threshold(dB)   14 attack_time(ms)  15 release_time(ms) 16 ratio
  17 bal             18 gain(dB)         19 invert(0/1)
"""
import numpy as np
import scipy.signal

N_PARAMS = 20
SAMPLE_RATE = 44100

# (min, max) per parameter: common_audioeffects.py:416-432 (EQ), :615-618 (comp), :963 (imager), :1035-1036 (gain)
PARAM_RANGES = np.array([
    (-15, 15), (30, 200),
    (-15, 15), (200, 1000), (0.1, 2.0),
    (-15, 15), (1000, 3000), (0.1, 2.0),
    (-15, 15), (3000, 8000), (0.1, 2.0),
    (-15, 15), (5000, 10000),
    (-80, -5), (1, 20), (50, 500), (4, 40),
    (0, 2), (-6, 9), (0, 1)], dtype=np.float64)


def random_params(batch: int, seed: int = 1234) -> np.ndarray:
    """Uniform over the documented ranges; `invert` is a fair coin.  float32 [batch, 20]."""
    rng = np.random.RandomState(seed)
    u = rng.rand(batch, N_PARAMS)
    p = PARAM_RANGES[:, 0] + u * (PARAM_RANGES[:, 1] - PARAM_RANGES[:, 0])
    p[:, 19] = (u[:, 19] < 0.5).astype(np.float64)
    return p.astype(np.float32)


def rbj_biquad(G, Q, fc, rate, filter_type):
    """RBJ Audio-EQ-Cookbook coefficients normalised by a0 (float64): SURVEY.md Appendix A.4."""
    A = 10.0 ** (G / 40.0)
    w0 = 2.0 * np.pi * (fc / rate)
    alpha = np.sin(w0) / (2.0 * Q)
    c = np.cos(w0)
    s = 2.0 * np.sqrt(A) * alpha
    if filter_type == "peaking":
        b = [1.0 + alpha * A, -2.0 * c, 1.0 - alpha * A]
        a = [1.0 + alpha / A, -2.0 * c, 1.0 - alpha / A]
    elif filter_type == "low_shelf":
        b = [A * ((A + 1) - (A - 1) * c + s), 2 * A * ((A - 1) - (A + 1) * c), A * ((A + 1) - (A - 1) * c - s)]
        a = [(A + 1) + (A - 1) * c + s, -2 * ((A - 1) + (A + 1) * c), (A + 1) + (A - 1) * c - s]
    elif filter_type == "high_shelf":
        b = [A * ((A + 1) + (A - 1) * c + s), -2 * A * ((A - 1) + (A + 1) * c), A * ((A + 1) + (A - 1) * c - s)]
        a = [(A + 1) - (A - 1) * c + s, 2 * ((A - 1) - (A + 1) * c), (A + 1) - (A - 1) * c - s]
    else:
        raise ValueError(filter_type)
    return np.asarray(b, np.float64) / a[0], np.asarray(a, np.float64) / a[0]


def eq_biquads(p, rate=SAMPLE_RATE):
    """The 5 cascaded sections in dict order low_shelf, first, second, third, high_shelf (:391, :438-462);
    shelves use Q = 0.707 (:454).  Returns [(b, a)] * 5."""
    p = np.asarray(p, np.float64)
    return [rbj_biquad(p[0], 0.707, p[1], rate, "low_shelf"),
            rbj_biquad(p[2], p[4], p[3], rate, "peaking"),
            rbj_biquad(p[5], p[7], p[6], rate, "peaking"),
            rbj_biquad(p[8], p[10], p[9], rate, "peaking"),
            rbj_biquad(p[11], 0.707, p[12], rate, "high_shelf")]


def equaliser(x, p, rate=SAMPLE_RATE):
    """Equaliser.process (:501-525): state reset before every band (:512), float64 cascade, cast to float32 (:519)."""
    y = np.asarray(x, np.float64)
    for b, a in eq_biquads(p, rate):
        y = scipy.signal.lfilter(b, a, y, axis=0)
    return y.astype(np.float32)


def _compressor_channel(x, threshold, attack_time, release_time, ratio, makeup_gain, sample_rate):
    """compressor_process (:529-587), one channel, float64 internals.  `np.log10(np.abs(x[i]))` is evaluated in
    float32 inside numba (x is a float32 array) and promoted by the `20 *`; mirrored here."""
    M = x.shape[0]
    ax = np.abs(x.astype(np.float32))
    x_g = np.where(ax < np.float32(0.000001), -120.0, 20.0 * np.log10(np.maximum(ax, np.float32(1e-30))).astype(np.float64))
    if ratio > 1:
        y_g = np.where(x_g >= threshold, threshold + (x_g - threshold) / ratio, x_g)
    elif ratio < 1:
        y_g = np.where(x_g <= threshold, threshold + (x_g - threshold) / (1 / ratio), x_g)
    else:
        y_g = np.zeros(M)  # ratio == 1 leaves y_g = 0 (:564-573)
    x_l = x_g - y_g
    alpha_attack = np.exp(-1 / (0.001 * sample_rate * attack_time))
    alpha_release = np.exp(-1 / (0.001 * sample_rate * release_time))
    y_l = _smooth(x_l, alpha_attack, alpha_release)
    c = np.power(10.0, (makeup_gain - y_l) / 20.0)
    return x * c   # float32 * float64 -> float64; stored into a float32 array by the caller (:638)


try:  # numba is in the image; fall back to a python loop (slow but identical) if it is not
    from numba import njit

    @njit(cache=False)
    def _smooth(x_l, alpha_attack, alpha_release):
        y = np.zeros(x_l.shape[0])
        prev = 0.0
        for i in range(x_l.shape[0]):
            if x_l[i] > prev:
                prev = alpha_attack * prev + (1 - alpha_attack) * x_l[i]
            else:
                prev = alpha_release * prev + (1 - alpha_release) * x_l[i]
            y[i] = prev
        return y
except Exception:  # pragma: no cover
    def _smooth(x_l, alpha_attack, alpha_release):
        y = np.zeros(x_l.shape[0])
        prev = 0.0
        for i in range(x_l.shape[0]):
            a = alpha_attack if x_l[i] > prev else alpha_release
            prev = a * prev + (1 - a) * x_l[i]
            y[i] = prev
        return y


def compressor(x, p, rate=SAMPLE_RATE):
    """Compressor.process (:624-652): per channel, makeup 0, state reset per call (yL_prev overwritten :553)."""
    thr, att, rel, ratio = float(p[13]), float(p[14]), float(p[15]), float(p[16])
    if thr == 0.0 and ratio == 1.0:  # :635
        return x
    y = np.zeros_like(x)
    for ch in range(x.shape[1]):
        y[:, ch] = _compressor_channel(x[:, ch], thr, att, rel, ratio, 0.0, rate)
    return y


def imager(x, p):
    """MidSideImager.process (:965-992), float32 in / float32 out."""
    left, right = x[:, 0], x[:, 1]
    mid, side = left + right, left - right
    mid_e, side_e = np.sum(mid ** 2), np.sum(side ** 2)
    total_e = mid_e + side_e
    max_side_multiplier = np.sqrt(total_e / (side_e + 1e-3))
    cur_bal = round(float(p[17]), 3)
    side_gain = cur_bal if cur_bal <= 1.0 else max_side_multiplier * (cur_bal - 1)
    new_side = side * side_gain
    new_side_e = side_e * (side_gain ** 2)
    left_mid_e = total_e - new_side_e
    mid_gain = np.sqrt(left_mid_e / (mid_e + 1e-3))
    new_mid = mid * mid_gain
    return np.stack([(new_mid + new_side) / 2, (new_mid - new_side) / 2], 1)


def gain(x, p):
    """Gain.process (:1038-1051)."""
    g = 10 ** (float(p[18]) / 20.0)
    if p[19] >= 0.5:
        g = -g
    return g * x


def rms_normalize(x, y):
    """AugmentationChain.apply_processor (:142-145): y *= sqrt(mean(x^2) / max(1e-7, mean(y^2)))."""
    scale = np.sqrt(np.mean(np.square(x)) / np.maximum(1e-7, np.mean(np.square(y))))
    return y * scale


def fx_chain(x, p, rate=SAMPLE_RATE):
    """create_effects_augmentation_chain(['eq','comp','imager','gain']) applied with every gate on (p=1) and the
    given parameters: audio_effects_chain.py:17-95 (rms_normalize False only for Gain, :92) and
    AugmentationChain.__call__ (common_audioeffects.py:156-192).  x: float32 [n, 2]."""
    x = np.asarray(x, np.float32)
    y = rms_normalize(x, equaliser(x, p, rate)).astype(np.float32)
    y = rms_normalize(y, compressor(y, p, rate)).astype(np.float32)
    y = rms_normalize(y, imager(y, p)).astype(np.float32)
    return np.asarray(gain(y, p), np.float32)


# ---- the reverbs ---------------------------------------------------------------------------------------
COMB_DELAYS = (1116, 1188, 1277, 1356, 1422, 1491, 1557, 1617)      # common_audioeffects.py:1525-1540
ALLPASS_DELAYS = ((556, 441, 341, 225), (556 + 23, 441 + 23, 341 + 23, 255 + 23))   # :1516-1523 (R4 is `255 + ss` in the reference)


def _comb_blocks(x, D, damp, feedback):
    """Freeverb comb (restated pymixconsole component, see oracle/shims/pymixconsole/components/comb.py) block by block: the
    delay line couples sample n to n - D only, the damping one-pole runs along n -- the decomposition the GPU kernel uses."""
    n = len(x)
    out = np.zeros(n)
    buf = np.zeros(D)
    store = 0.0
    for k in range(0, n, D):
        m = min(D, n - k)
        y = buf[:m].copy()
        f, _ = scipy.signal.lfilter([1.0 - damp], [1.0, -damp], y, zi=[damp * store])
        store = f[-1]
        buf[:m] = x[k:k + m] + f * feedback
        out[k:k + m] = y
    return out


def _allpass_blocks(x, D, feedback):
    n = len(x)
    out = np.zeros(n)
    buf = np.zeros(D)
    for k in range(0, n, D):
        m = min(D, n - k)
        b = buf[:m].copy()
        out[k:k + m] = -x[k:k + m] + b
        buf[:m] = x[k:k + m] + b * feedback
    return out


def algorithmic_reverb(x, room_size, damping, dry_mix, wet_mix, width):
    """AlgorithmicReverb.process (common_audioeffects.py:1446-1509) on float [n, 2]; float64 [n, 2] out like the reference.
    Combs 1-4 are computed and then overwritten by comb 5 in the reference (:1478, :1487): only combs 5-8 matter."""
    chans = []
    for c in range(2):
        src = x[:, c].copy() * 0.2                                   # `dataL.copy() * self.scalegain` in the data's dtype
        s = None
        for D in COMB_DELAYS[4:]:
            y = _comb_blocks(np.asarray(src, dtype=np.float64), D + (23 if c else 0), damping, room_size)
            s = y if s is None else s + y
        for D in ALLPASS_DELAYS[c]:
            s = _allpass_blocks(s, D, room_size)
        chans.append(s)
    wet1 = wet_mix * ((width / 2) + 0.5)
    wet2 = wet_mix * ((1 - width) / 2)
    out = np.zeros((x.shape[0], 2))
    out[:, 0] = (wet1 * chans[0]) + (wet2 * chans[1]) + (dry_mix * x[:, 0])
    out[:, 1] = (wet1 * chans[1]) + (wet2 * chans[0]) + (dry_mix * x[:, 1])
    return out


def convolutional_reverb(x, h, pre_delay_ms=0, wet=1.0, dry=0.0, sample_rate=SAMPLE_RATE):
    """ConvolutionalReverb.process (common_audioeffects.py:735-764) with a given impulse response h [m, 1 or 2]."""
    from scipy.signal import oaconvolve
    if h.shape[1] == 1 and x.shape[1] > 1:
        h = np.hstack([h] * x.shape[1])
    if wet == 0.0:
        return x
    y = oaconvolve(x, h, mode='full', axes=0)
    idx = np.argmax(np.max(np.abs(h), axis=1), axis=0)
    idx += int(0.001 * np.abs(pre_delay_ms) * sample_rate)
    idx = np.clip(idx, 0, h.shape[0] - 1)
    y = y[idx:idx + x.shape[0], :]
    return dry * x + wet * y
