"""TEST INFRASTRUCTURE ONLY -- CPU (numpy / scipy float64) restatement of the input FX normaliser's EQ-matching, loudness and
imager effects.  Paths relative to /root/reference/mixing_style_transfer/mixing_manipulator/.

  normalize_audio / normalize_audio_per_effect   data_normalization.py:77-155   (pad by FFT_SIZE, -40 dB gate, crop)
  lufs_normalize                                  fx_utils.py:220-238
  get_eq_matching / compute_stft                  utils_data_normalization.py:65-107, common_miscellaneous.py:50-77
                                                  (librosa.stft underneath is third-party and absent: restated in stft())
  normalize_imager / process_balance              normalization_imager.py:22-118 (Haas branch excluded: it draws random
                                                  parameters through pymixconsole's Processor.randomize)
  Meter / loudness_gain_apply                     pyloudnorm==0.1.0 (requirements.txt:9), third-party, NOT in the reference
                                                  tree: restated from the published BS.1770-4 algorithm -> PARITY UNPINNED

Pinned where the reference's own code can run here: tests/test_oracle_pinned.py imports the reference's fx_utils,
normalization_imager and data_normalization on the shims of oracle/shims/ (whose `pyloudnorm` is THIS file's Meter) and
compares lufs_normalize / normalize_imager / Audio_Effects_Normalizer.normalize_audio_per_effect with the functions below.
Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may import this module.
"""
import numpy as np
import scipy.signal

FFT_SIZE = 2 ** 16          # data_normalization.py:31
MIN_DB = -40                # :37
SR = 44100


def rbj_high_shelf(G, Q, fc, rate):
    A = 10 ** (G / 40.0)
    w0 = 2.0 * np.pi * (fc / rate)
    alpha = np.sin(w0) / (2.0 * Q)
    c, s = np.cos(w0), 2 * np.sqrt(A) * alpha
    b = np.array([A * ((A + 1) + (A - 1) * c + s), -2 * A * ((A - 1) + (A + 1) * c), A * ((A + 1) + (A - 1) * c - s)])
    a = np.array([(A + 1) - (A - 1) * c + s, 2 * ((A - 1) - (A + 1) * c), (A + 1) - (A - 1) * c - s])
    return b / a[0], a / a[0]


def rbj_high_pass(Q, fc, rate):
    w0 = 2.0 * np.pi * (fc / rate)
    alpha = np.sin(w0) / (2.0 * Q)
    c = np.cos(w0)
    b = np.array([(1 + c) / 2, -(1 + c), (1 + c) / 2])
    a = np.array([1 + alpha, -2 * c, 1 - alpha])
    return b / a[0], a / a[0]


def k_weighting(rate=SR):
    """The two K-weighting stages of pyloudnorm's default meter, in application order."""
    return [rbj_high_shelf(4.0, 1 / np.sqrt(2), 1500.0, rate), rbj_high_pass(0.5, 38.0, rate)]


def gating_block_bounds(num_samples, rate=SR, block_size=0.400, overlap=0.75):
    """(lower, upper) sample bounds of every gating block, with pyloudnorm's float arithmetic."""
    T_g, step = block_size, 1.0 - overlap
    T = num_samples / rate
    num_blocks = int(np.round(((T - T_g) / (T_g * step))) + 1)
    j = np.arange(0, max(num_blocks, 0))
    lo = np.array([int(T_g * (jj * step) * rate) for jj in j], dtype=np.int64)
    hi = np.array([int(T_g * (jj * step + 1) * rate) for jj in j], dtype=np.int64)
    return lo, hi


def gated_loudness(z, block_size=0.400, rate=SR):
    """z[channel][block] = SUM of squares of the K-weighted signal per gating block -> integrated loudness (LUFS)."""
    G = [1.0, 1.0, 1.0, 1.41, 1.41]
    z = np.asarray(z, dtype=np.float64) * (1.0 / (block_size * rate))
    n_ch, n_blocks = z.shape
    with np.errstate(divide='ignore', invalid='ignore'):
        l = np.array([-0.691 + 10.0 * np.log10(np.sum([G[i] * z[i, j] for i in range(n_ch)])) for j in range(n_blocks)])
        J_g = [j for j in range(n_blocks) if l[j] >= -70.0]
        z_avg = [np.mean([z[i, j] for j in J_g]) for i in range(n_ch)]
        gamma_r = -0.691 + 10.0 * np.log10(np.sum([G[i] * z_avg[i] for i in range(n_ch)])) - 10.0
        J_g = [j for j in range(n_blocks) if (l[j] > gamma_r and l[j] > -70.0)]
        z_avg = np.nan_to_num(np.array([np.mean([z[i, j] for j in J_g]) for i in range(n_ch)]))
        return float(-0.691 + 10.0 * np.log10(np.sum([G[i] * z_avg[i] for i in range(n_ch)])))


class Meter:
    """pyloudnorm.Meter(rate) with the default K-weighting filters and 400 ms blocks."""

    def __init__(self, rate, filter_class="K-weighting", block_size=0.400):
        self.rate, self.block_size = rate, block_size
        self._filters = k_weighting(rate)

    def integrated_loudness(self, data):
        x = np.array(data, copy=True)
        if x.ndim == 1:
            x = x.reshape(-1, 1)
        for b, a in self._filters:
            for ch in range(x.shape[1]):
                x[:, ch] = scipy.signal.lfilter(b, a, x[:, ch])          # stored back in the data's dtype
        lo, hi = gating_block_bounds(x.shape[0], self.rate, self.block_size)
        z = np.array([[np.sum(np.square(x[l:u, ch])) for l, u in zip(lo, hi)] for ch in range(x.shape[1])], dtype=np.float64)
        return gated_loudness(z.reshape(x.shape[1], len(lo)), self.block_size, self.rate)


def loudness_gain_apply(data, input_loudness, target_loudness):
    """pyloudnorm.normalize.loudness: gain = 10^((target - input) / 20) as a numpy float64 scalar -> float64 output."""
    gain = np.power(10.0, (target_loudness - input_loudness) / 20.0)
    return gain * data


def lufs_normalize(x, sr, lufs):
    """fx_utils.py:220-238 (log=False)."""
    meter = Meter(sr)
    loudness = meter.integrated_loudness(x + 1e-10)
    y = loudness_gain_apply(x, loudness, lufs)
    y = y / np.maximum(1.0, 1e-6 + np.max(np.abs(y)))
    return y


def sqrt_hann(n_fft):
    return np.sqrt(np.hanning(n_fft + 1)[:-1])          # utils_data_normalization.py:77


def stft(y, n_fft, hop_length, window):
    """librosa.stft(y, n_fft=, hop_length=, window=, center=False) -- librosa==0.9.2 is third-party and absent: restated from
    its documented behaviour (frames y[f*hop : f*hop + n_fft] * window, rfft, [1 + n_fft/2, n_frames], complex64 for float32
    input and complex128 for float64)."""
    y = np.asarray(y)
    n_frames = 1 + (len(y) - n_fft) // hop_length
    out = np.empty((n_fft // 2 + 1, n_frames), dtype=np.complex64 if y.dtype == np.float32 else np.complex128)
    for f in range(n_frames):
        out[:, f] = np.fft.rfft(window * y[f * hop_length:f * hop_length + n_fft])
    return out


def stft_mag_mean(x, n_fft=FFT_SIZE, hop=FFT_SIZE // 4):
    """compute_stft (common_miscellaneous.py:50-77: complex64 storage) + np.abs + np.mean over frames
    (utils_data_normalization.py:74-79) of a 1-D signal."""
    n_frames = 1 + int((x.shape[0] - n_fft) / hop)
    D = np.empty((n_frames, n_fft // 2 + 1), dtype=np.complex64)
    D[:, :] = stft(x, n_fft, hop, sqrt_hann(n_fft)).transpose()
    return np.mean(np.abs(D), axis=0)


def get_eq_matching(audio_t, ref_spec, sr=SR, n_fft=FFT_SIZE, hop_length=FFT_SIZE // 4, min_db=MIN_DB, ntaps=1001, lufs=-30):
    """utils_data_normalization.py:65-107 on one channel (1-D)."""
    audio_t = np.copy(audio_t)
    with np.errstate(divide='ignore'):
        max_db = 20 * np.log10(np.max(np.abs(audio_t)) + 1e-30)
    if not max_db > min_db:
        return audio_t
    audio_t = lufs_normalize(audio_t, sr, lufs)
    avg = stft_mag_mean(audio_t, n_fft, hop_length)
    m = ref_spec.shape[0]
    frq = np.arange(m) / (m / sr) / 2
    with np.errstate(divide='ignore'):
        diff_eq = (20 * np.log10(ref_spec + 1e-30)) - (20 * np.log10(avg + 1e-30))
    diff_eq = np.sqrt(10 ** (diff_eq / 20))
    taps = scipy.signal.firwin2(ntaps, frq / np.max(frq), diff_eq, nfreqs=None, window='hamming', antisymmetric=False)
    return scipy.signal.filtfilt(taps, 1, audio_t, axis=-1, padtype='odd', padlen=None, method='pad', irlen=None)


def smooth_eq_feature(ref_spec, src):
    """Audio_Effects_Normalizer.smooth_feature for 'eq' (data_normalization.py:158-170)."""
    return scipy.signal.savgol_filter(ref_spec, 401 if src in ('other', 'vocals') else 151, 1, mode='mirror')


def stub_onsets(x, sr=SR, window=1024):
    """Onset positions (samples) of a 1-D signal under the rule of oracle/shims/aubio (NOT aubio's detector), driven exactly as
    get_mean_peak drives aubio (utils_data_normalization.py:302-312): float32 frames of `window` samples, hop = window."""
    import importlib.util
    import os
    spec = importlib.util.spec_from_file_location("_mst_aubio_stub", os.path.join(os.path.dirname(os.path.abspath(__file__)), "shims", "aubio", "__init__.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    o = mod.onset('hfc', buf_size=window, hop_size=window, samplerate=sr)
    frames = np.float32(np.lib.stride_tricks.sliding_window_view(np.asarray(x), window)[::window])
    return [o.get_last() for fr in frames if o(fr)]


def get_mean_peak(audio, onset_fn=stub_onsets, sr=SR, percentile=75):
    """utils_data_normalization.py:284-337 (true_peak=False): mean / std in dB of the inter-onset peaks above the percentile."""
    peak, std = [], []
    for ch in range(audio.shape[-1]):
        x = np.ascontiguousarray(audio[:, ch])
        onset_times = onset_fn(x, sr, 2 ** 10)
        samples = []
        if onset_times:
            for i in range(len(onset_times) - 1):
                samples.append(onset_times[i] + np.argmax(np.abs(x[onset_times[i]:onset_times[i + 1]])))
            samples.append(onset_times[-1] + np.argmax(np.abs(x[onset_times[-1]:])))
        p_value = [20 * np.log10(np.abs(x[p]) + 1e-30) for p in samples]
        p_value_ = [p for p in p_value if p > np.percentile(p_value, percentile)]
        if p_value_:
            peak.append(np.mean(p_value_)); std.append(np.std(p_value_))
        elif p_value:
            peak.append(np.mean(p_value)); std.append(np.std(p_value))
        else:
            return None
    return [np.mean(peak), np.mean(std)]


def get_comp_matching(audio, ref_peak, ref_std, ratio, attack, release, onset_fn=stub_onsets, sr=SR, min_db=MIN_DB,
                      comp_peak_norm=-10.0, min_th=-40, max_ratio=20, percentile=75, expander=False):
    """utils_data_normalization.py:357-429 on one channel (1-D in, [n, 1] out).  The compressor is oracle/fx_oracle's (pinned to
    the reference's numba code); pyloudnorm.normalize.peak is restated (gain = 10^(target/20) / max|x|)."""
    from oracle import fx_oracle
    x = audio.copy()
    if x.ndim < 2:
        x = np.expand_dims(x, 1)
    with np.errstate(divide='ignore'):
        max_db = 20 * np.log10(np.max(np.abs(x)) + 1e-30)
    if not max_db > min_db:
        return x
    x = (np.power(10.0, comp_peak_norm / 20.0) / np.max(np.abs(x))) * x
    peak, std = get_mean_peak(x, onset_fn, sr, percentile)            # None -> TypeError, like the reference
    if ref_peak - ref_std < peak < ref_peak + ref_std:
        return x
    if peak > (ref_peak - ref_std):
        ratios = np.linspace(ratio, max_ratio, max_ratio - ratio + 1)
        ths = np.linspace(-1 - 9, min_th, 2 * np.abs(min_th) - 1 - 18)
        y = x
        for rt in ratios:
            done = False
            for th in ths:
                p = np.zeros(20)
                p[13], p[14], p[15], p[16] = th, attack, release, rt
                y = fx_oracle.compressor(x, p, sr)
                if np.max(np.abs(y)) >= 1.0:
                    y = np.clip(y, -1.0, 1.0)
                peak, std = get_mean_peak(y, onset_fn, sr, percentile)
                if peak < (ref_peak + ref_std):
                    done = True
                    break
            if done:
                break
        return y
    if expander:
        raise NotImplementedError("COMP_USE_EXPANDER is False in the reference (data_normalization.py:40)")
    return x


COMP_SETTINGS = {'vocals': (7.5, 400.0, 4), 'drums': (10.0, 180.0, 6), 'bass': (10.0, 500.0, 5), 'other': (15.0, 666.0, 4)}   # data_normalization.py:47-71


def process_balance(d1, d2, tgt_e1_bal=0.5, eps=1e-04):
    """normalization_imager.py:84-99."""
    e1, e2 = np.sum(d1 ** 2), np.sum(d2 ** 2)
    total = e1 + e2
    g1 = np.sqrt(tgt_e1_bal * total / (e1 + eps))
    left = total - e1 * (g1 ** 2)
    g2 = np.sqrt(left / (e2 + 1e-3))
    return d1 * g1, d2 * g2


def imager_is_almost_mono(data, mono_threshold):
    mid, side = data[:, 0] + data[:, 1], data[:, 0] - data[:, 1]
    mid_e, side_e = np.sum(mid ** 2), np.sum(side ** 2)
    return bool(mid_e / (mid_e + side_e) > mono_threshold)


def normalize_imager(data, target_side_mid_bal=0.9, mono_threshold=0.95, eps=1e-04):
    """normalization_imager.py:22-81 for inputs that are NOT almost mono (no Haas)."""
    if imager_is_almost_mono(data, mono_threshold):
        raise ValueError("almost-mono input: the reference applies a randomised Haas effect here (no deterministic oracle)")
    mid, side = data[:, 0] + data[:, 1], data[:, 0] - data[:, 1]
    mid, side = process_balance(mid, side, target_side_mid_bal, eps)
    left, right = (mid + side) / 2, (mid - side) / 2
    left, right = process_balance(left, right, 0.5, eps)
    mid, side = left + right, left - right
    mid, side = process_balance(mid, side, target_side_mid_bal, eps)
    return np.stack([(mid + side) / 2, (mid - side) / 2], 1)


def normalize_audio_per_effect(audio, effect, feature, src="drums", onset_fn=stub_onsets):
    """data_normalization.py:88-155 for effect in ('eq', 'loudness', 'imager').  audio: [n, 2]; feature = features_mean[effect][src]
    as the normaliser holds it after smooth_feature."""
    audio = audio.astype(np.float32)
    track = np.pad(audio, ((FFT_SIZE, FFT_SIZE), (0, 0)), mode='constant')
    out = track.copy()
    with np.errstate(divide='ignore'):
        max_db = 20.0 * np.log10(np.max(np.abs(out)) + 1e-30)       # amp_to_db, utils_data_normalization.py:35-36
    if max_db > MIN_DB:
        if effect == 'eq':
            for ch in range(out.shape[1]):        # feature = the SMOOTHED target spectrum (smooth_eq_feature)
                np.copyto(out[:, ch], get_eq_matching(out[:, ch], feature), casting='same_kind')
        elif effect == 'compression':
            att, rel, ratio = COMP_SETTINGS[src]
            for ch in range(out.shape[1]):
                try:
                    y = get_comp_matching(out[:, ch], feature[0], feature[1], ratio, att, rel, onset_fn=onset_fn)
                    np.copyto(out[:, ch], y[:, 0], casting='same_kind')
                except Exception:          # the reference's bare `except: break` (data_normalization.py:137-138)
                    break
        elif effect == 'loudness':
            out = lufs_normalize(out, SR, feature)
        elif effect == 'imager':
            np.copyto(out, normalize_imager(out, target_side_mid_bal=feature,
                                            mono_threshold=0.99 if src == 'bass' else 0.975))
        else:
            raise NotImplementedError(effect)
    return out[FFT_SIZE:FFT_SIZE + audio.shape[0]]


def normalize_audio(audio, effects, features, src="drums", onset_fn=stub_onsets):
    """data_normalization.py:77-85."""
    y = audio
    for e in effects:
        y = normalize_audio_per_effect(y, e, features[e][src], src, onset_fn)
    return y


def haas_process(x, delay, feedback, wet_channel):
    """common_audioeffects.py:767-787.  x: [n, 2]."""
    y = np.copy(x)
    if wet_channel == 'left':
        y[:, 0] += feedback * np.roll(x[:, 0], delay)
    elif wet_channel == 'right':
        y[:, 1] += feedback * np.roll(x[:, 1], delay)
    return y


def pan_gains(pan, pan_law='-4.5dB'):
    """Panner._calculate_pan_coefficents (common_audioeffects.py:877-908), float32 like `self.dtype`."""
    g = np.zeros(2, dtype=np.float32)
    theta = pan * (np.pi / 2)
    if pan_law == 'linear':
        g[0], g[1] = ((np.pi / 2) - theta) * (2 / np.pi), theta * (2 / np.pi)
    elif pan_law == 'constant_power':
        g[0], g[1] = np.cos(theta), np.sin(theta)
    elif pan_law == '-4.5dB':
        g[0] = np.sqrt(((np.pi / 2) - theta) * (2 / np.pi) * np.cos(theta))
        g[1] = np.sqrt(theta * (2 / np.pi) * np.sin(theta))
    else:
        raise ValueError(f'Invalid pan_law {pan_law}.')
    return g
