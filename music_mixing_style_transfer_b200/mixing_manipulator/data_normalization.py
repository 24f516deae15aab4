"""Audio effects chain normalisation on the GPU -- same surface as the reference's
`mixing_manipulator/data_normalization.py` (class Audio_Effects_Normalizer :19-175).

  normalize_audio(audio, src)             :77-85     the effects of `EFFECTS` in order
  normalize_audio_per_effect(...)         :88-155    pad by FFT_SIZE zeros, skip below -40 dB peak, effect, crop
    'loudness'  fx_utils.lufs_normalize (fx_utils.py:220-238): BS.1770 integrated loudness of the padded stem (the meter is
                pyloudnorm's, a third-party package: K-weighting biquads -> 400 ms / 75 % gating blocks -> absolute and
                relative gates), gain to the stem's target LUFS, then division by max(1, peak)
    'imager'    normalize_imager (normalization_imager.py:22-81): mid/side balance to the stem's target, left/right balance
                50-50, mid/side balance again; a Haas effect first when the stem is almost mono
    'eq'        get_eq_matching (utils_data_normalization.py:65-107): per channel, mono loudness normalisation to -30 LUFS,
                65,536-point sqrt-Hann STFT averaged over frames (csrc/spectral.cu), difference to the stem's target spectrum
                -> scipy.signal.firwin2 (1001 taps, host) -> zero-phase filtering (mst_fir_filtfilt, float64)
    'compression' get_comp_matching (:357-429): peak-normalise to -10 dB, mean inter-onset peak (onset detector = aubio, a
                third-party package, on the host), then a (ratio, threshold) grid of compressor runs -- those are the FX chain's
                compressor kernel, several thresholds per launch -- until the peak falls into the stem's target band

Device work per effect: one reduction pass (mst_stereo_stats, and for loudness mst_biquad_cascade + mst_block_energy), a few
dozen float64 operations on the host with the reference's own formulas, one 2x2 mix pass (mst_stereo_mix).  The three
process_balance steps of the imager collapse into ONE 2x2 matrix: every step is linear in (L, R), so the energies each step
needs follow from (sum L^2, sum R^2, sum LR) of the input.  `audio` may be a float32 CUDA tensor [2, T] (the engine's stems)
or a numpy array [n, 2] / [n, 1] (the reference's call, :77); the result has the same kind.  No CPU path.
"""
import ctypes

import numpy as np
import torch

from .. import _cabi

FFT_SIZE = 2 ** 16
_G = np.array([1.0, 1.0, 1.0, 1.41, 1.41])     # BS.1770 channel weights


def _rbj(kind, G, Q, fc, rate):
    A = 10 ** (G / 40.0)
    w0 = 2.0 * np.pi * (fc / rate)
    alpha, c = np.sin(w0) / (2.0 * Q), np.cos(w0)
    if kind == 'high_shelf':
        s = 2 * np.sqrt(A) * alpha
        b = [A * ((A + 1) + (A - 1) * c + s), -2 * A * ((A - 1) + (A + 1) * c), A * ((A + 1) + (A - 1) * c - s)]
        a = [(A + 1) - (A - 1) * c + s, 2 * ((A - 1) - (A + 1) * c), (A + 1) - (A - 1) * c - s]
    else:   # high_pass
        b = [(1 + c) / 2, -(1 + c), (1 + c) / 2]
        a = [1 + alpha, -2 * c, 1 - alpha]
    return [b[0] / a[0], b[1] / a[0], b[2] / a[0], a[1] / a[0], a[2] / a[0]]


def _stream():
    return _cabi.current_stream()


def stereo_stats(x):
    """x float32 CUDA [B, 2, L] -> float64 numpy [B, 4] = (sum L^2, sum R^2, sum LR, max |x|)."""
    B, _, L = x.shape
    out = torch.empty(B, 4, dtype=torch.float64, device=x.device)
    _cabi.check(_cabi.lib().mst_stereo_stats(_cabi.ptr(x), B, L, out.data_ptr(), _stream()), "stereo_stats")
    return out.cpu().numpy()


def stereo_mix(x, matrices, out=None):
    """y = M x per frame; x float32 CUDA [B, 2, L], matrices [B, 4] = (m0, m1, m2, m3) (numpy or tensor)."""
    B, _, L = x.shape
    m = torch.as_tensor(np.asarray(matrices, dtype=np.float32).reshape(B, 4)).to(x.device) \
        if not isinstance(matrices, torch.Tensor) else matrices.to(device=x.device, dtype=torch.float32).contiguous()
    y = torch.empty_like(x) if out is None else out
    _cabi.check(_cabi.lib().mst_stereo_mix(_cabi.ptr(x), _cabi.ptr(m), _cabi.ptr(y), B, L, _stream()), "stereo_mix")
    return y


def haas(x, delay, feedback, channel):
    """haas_process (common_audioeffects.py:767-787) on x float32 CUDA [B, 2, L]; per-segment delay (samples, may be negative:
    np.roll wraps), feedback and wet channel (0 = left, 1 = right)."""
    B, _, L = x.shape
    dev = x.device
    d = torch.as_tensor(np.asarray(delay, dtype=np.int32).reshape(B)).to(dev)
    f = torch.as_tensor(np.asarray(feedback, dtype=np.float32).reshape(B)).to(dev)
    c = torch.as_tensor(np.asarray(channel, dtype=np.int32).reshape(B)).to(dev)
    y = torch.empty_like(x)
    _cabi.check(_cabi.lib().mst_haas(_cabi.ptr(x), _cabi.ptr(y), B, L, d.data_ptr(), _cabi.ptr(f), c.data_ptr(), _stream()), "haas")
    return y


def gating_block_bounds(num_samples, rate=44100, block_size=0.400, overlap=0.75):
    """Sample bounds of the loudness meter's gating blocks, in the float arithmetic of the meter the reference uses."""
    T_g, step = block_size, 1.0 - overlap
    num_blocks = int(np.round(((num_samples / rate - T_g) / (T_g * step))) + 1)
    j = range(max(num_blocks, 0))
    return (np.array([int(T_g * (jj * step) * rate) for jj in j], dtype=np.int64),
            np.array([int(T_g * (jj * step + 1) * rate) for jj in j], dtype=np.int64))


def gated_loudness(z_sums, rate=44100, block_size=0.400):
    """[channels, blocks] block sums of squares of the K-weighted signal -> integrated loudness in LUFS (BS.1770-4: absolute
    gate -70 LUFS, relative gate 10 LU below the mean of the blocks that passed it)."""
    z = np.asarray(z_sums, dtype=np.float64) / (block_size * rate)
    g = _G[:z.shape[0], None]
    with np.errstate(divide='ignore', invalid='ignore'):
        l = -0.691 + 10.0 * np.log10(np.sum(g * z, axis=0))
        keep = l >= -70.0
        gamma_r = -0.691 + 10.0 * np.log10(np.sum(g[:, 0] * z[:, keep].mean(axis=1))) - 10.0 if keep.any() else np.nan
        keep = (l > gamma_r) & (l > -70.0)
        z_avg = np.nan_to_num(z[:, keep].mean(axis=1)) if keep.any() else np.zeros(z.shape[0])
        return float(-0.691 + 10.0 * np.log10(np.sum(g[:, 0] * z_avg)))


def kweighted_block_sums(x, rate=44100):
    """x float32 CUDA [2, T] -> float64 numpy [2, blocks]: sums of squares of the K-weighted (x + 1e-10) per gating block
    (fx_utils.py:224: the meter is fed x + 1e-10), or None when the signal is shorter than one block."""
    lib = _cabi.lib()
    T = x.shape[-1]
    xe = (x + 1e-10).unsqueeze(0).contiguous()
    coef = torch.tensor([[_rbj('high_shelf', 4.0, 1 / np.sqrt(2), 1500.0, rate), _rbj('high_pass', 0.0, 0.5, 38.0, rate)]],
                        dtype=torch.float64, device=x.device)
    ws = torch.empty(max(lib.mst_fx_workspace_bytes(1, T), 4096), dtype=torch.uint8, device=x.device)
    y = torch.empty_like(xe)
    _cabi.check(lib.mst_biquad_cascade(_cabi.ptr(xe), coef.data_ptr(), 2, _cabi.ptr(y), 1, T, _cabi.ptr(ws), ws.numel(),
                                       _stream()), "biquad_cascade")
    lo, hi = gating_block_bounds(T, rate)
    if len(lo) == 0:
        return None
    lo_d, hi_d = torch.from_numpy(lo).to(x.device), torch.from_numpy(hi).to(x.device)
    z = torch.empty(2, len(lo), dtype=torch.float64, device=x.device)
    _cabi.check(lib.mst_block_energy(_cabi.ptr(y), 2, T, lo_d.data_ptr(), hi_d.data_ptr(), len(lo), z.data_ptr(), _stream()),
                "block_energy")
    return z.cpu().numpy()


def integrated_loudness(x, rate=44100):
    """x float32 CUDA [2, T] -> LUFS of (x + 1e-10), as fx_utils.lufs_normalize measures it (fx_utils.py:224)."""
    z = kweighted_block_sums(x, rate)
    return float('-inf') if z is None else gated_loudness(z, rate)


def integrated_loudness_per_channel(x, rate=44100):
    """Each channel of x float32 CUDA [2, T] measured as a MONO signal (get_eq_matching normalises one channel at a time,
    utils_data_normalization.py:72): one K-weighting pass over both channels, the gating per channel."""
    z = kweighted_block_sums(x, rate)
    return [float('-inf')] * 2 if z is None else [gated_loudness(z[c:c + 1], rate) for c in range(2)]


def lufs_normalize(x, sr, lufs):
    """fx_utils.lufs_normalize (fx_utils.py:220-238) on x float32 CUDA [2, T]: gain to `lufs`, then / max(1, 1e-6 + peak)."""
    loudness = integrated_loudness(x, sr)
    gain = float(np.power(10.0, (float(np.asarray(lufs).reshape(-1)[0]) - loudness) / 20.0))
    peak = gain * float(stereo_stats(x.unsqueeze(0))[0, 3])
    k = gain / max(1.0, 1e-6 + peak)
    return stereo_mix(x.unsqueeze(0), [[k, 0.0, 0.0, k]])[0]


def row_absmax(x):
    """x float32 CUDA [R, T] -> float64 numpy [R]: max |x| per row."""
    R, T = x.shape
    out = torch.empty(R, dtype=torch.float64, device=x.device)
    _cabi.check(_cabi.lib().mst_row_absmax(_cabi.ptr(x, True), R, T, max(x.stride(0), T), out.data_ptr(), _stream()), "row_absmax")
    return out.cpu().numpy()


_window_cache = {}


def stft_mag_mean(x, n_fft=FFT_SIZE, hop=FFT_SIZE // 4):
    """x float32 CUDA [R, T] -> float64 numpy [R, n_fft/2 + 1]: the frame-averaged magnitude of the sqrt-Hann STFT
    (compute_stft + np.abs + np.mean, utils_data_normalization.py:74-79)."""
    lib = _cabi.lib()
    R, T = x.shape
    key = (n_fft, x.device)
    if key not in _window_cache:
        _window_cache[key] = torch.from_numpy(np.sqrt(np.hanning(n_fft + 1)[:-1]).astype(np.float32)).to(x.device)
    ws = torch.empty(lib.mst_stft_workspace_bytes(R, n_fft), dtype=torch.uint8, device=x.device)
    out = torch.empty(R, n_fft // 2 + 1, dtype=torch.float64, device=x.device)
    _cabi.check(lib.mst_stft_mag_mean(_cabi.ptr(x, True), R, T, max(x.stride(0), T), n_fft, hop, _cabi.ptr(_window_cache[key]), out.data_ptr(),
                                      _cabi.ptr(ws), ws.numel(), _stream()), "stft_mag_mean")
    return out.cpu().numpy()


def fir_filtfilt(x, taps, scale=None):
    """scipy.signal.filtfilt(taps[r], 1, x[r]) per row (odd padding of 3 * n_taps, float64), times scale[r], as float32.
    x float32 CUDA [R, T]; taps float64 [R, n_taps]."""
    lib = _cabi.lib()
    R, T = x.shape
    taps = np.ascontiguousarray(np.asarray(taps, dtype=np.float64).reshape(R, -1))
    n_taps = taps.shape[1]
    t_d = torch.from_numpy(taps).to(x.device)
    s_d = None if scale is None else torch.from_numpy(np.asarray(scale, dtype=np.float64).reshape(R).copy()).to(x.device)
    ws = torch.empty(lib.mst_fir_filtfilt_workspace_bytes(R, T, n_taps), dtype=torch.uint8, device=x.device)
    y = torch.empty(R, T, dtype=torch.float32, device=x.device)
    _cabi.check(lib.mst_fir_filtfilt(_cabi.ptr(x, True), R, T, max(x.stride(0), T), t_d.data_ptr(), n_taps, None if s_d is None else s_d.data_ptr(),
                                     _cabi.ptr(y), y.stride(0), _cabi.ptr(ws), ws.numel(), _stream()), "fir_filtfilt")
    return y


def _amp_to_db(x):
    return 20 * np.log10(x + 1e-30)


def eq_matching(x, ref_spec, sr=44100, n_fft=FFT_SIZE, hop=FFT_SIZE // 4, min_db=-40, ntaps=1001, lufs=-30):
    """get_eq_matching (utils_data_normalization.py:65-107) on both channels of x float32 CUDA [2, T] at once: each channel is
    loudness-normalised as a mono signal (gain k, folded into the output), its averaged STFT magnitude is compared with the
    target spectrum, the difference becomes a 1001-tap linear-phase FIR (scipy.signal.firwin2 on the host: a filter DESIGN on
    32,769 numbers) and the channel is filtered forwards and backwards.  A channel below min_db passes through unchanged."""
    import scipy.signal
    peaks = row_absmax(x)
    with np.errstate(divide='ignore'):
        active = [_amp_to_db(float(p)) > min_db for p in peaks]
    if not any(active):
        return x
    loud = integrated_loudness_per_channel(x, sr)
    avg = stft_mag_mean(x, n_fft, hop)
    ref_spec = np.asarray(ref_spec)
    m = ref_spec.shape[0]
    frq = np.arange(m) / (m / sr) / 2
    taps = np.zeros((2, ntaps))
    scale = np.ones(2)
    for c in range(2):
        if not active[c]:
            taps[c, 0] = 1.0                         # identity: the channel is returned as it is (:104-105)
            continue
        gain = np.power(10.0, (lufs - loud[c]) / 20.0)
        scale[c] = gain / np.maximum(1.0, 1e-6 + gain * peaks[c])          # fx_utils.py:229-232
        with np.errstate(divide='ignore'):
            diff_eq = np.sqrt(np.power(10.0, (_amp_to_db(ref_spec) - _amp_to_db(scale[c] * avg[c])) / 20))
        taps[c] = scipy.signal.firwin2(ntaps, frq / np.max(frq), diff_eq, nfreqs=None, window='hamming', antisymmetric=False)
    return fir_filtfilt(x, taps, scale)


def aubio_onsets(x, sr=44100, window=1024):
    """Onset positions (samples) of a 1-D float signal from aubio's 'hfc' detector, driven as get_mean_peak drives it
    (utils_data_normalization.py:302-312): float32 frames of `window` samples, hop = window.  aubio is a third-party C library
    the reference depends on (requirements.txt); it is imported here on first use."""
    try:
        import aubio
    except ImportError as e:
        raise NotImplementedError(
            "the 'compression' effect of the input FX normaliser needs the `aubio` package for its onset detection "
            "(utils_data_normalization.py:302), as in the reference; install it, pass onset_detector=..., or leave "
            "'compression' out of normalization_order") from e
    onset_func = aubio.onset('hfc', buf_size=window, hop_size=window, samplerate=sr)
    x = np.ascontiguousarray(x)
    frames = np.float32(np.lib.stride_tricks.sliding_window_view(x, window)[::window])
    return [onset_func.get_last() for frame in frames if onset_func(frame)]


def mean_peak(x, onset_fn, sr=44100, percentile=75):
    """get_mean_peak (utils_data_normalization.py:284-337, true_peak=False) of a 1-D host signal: the largest |x| between
    consecutive onsets, in dB; mean and std of the values above the percentile.  None when there is no onset."""
    onset_times = onset_fn(x, sr, 2 ** 10)
    if not onset_times:
        return None
    a = np.abs(x)
    bounds = list(onset_times) + [len(x)]
    samples = [bounds[i] + int(np.argmax(a[bounds[i]:bounds[i + 1]])) for i in range(len(onset_times))]
    p_value = np.array([_amp_to_db(a[p]) for p in samples])
    top = p_value[p_value > np.percentile(p_value, percentile)]
    sel = top if len(top) else p_value
    return float(np.mean(sel)), float(np.std(sel))


def comp_matching(x, ref_peak, ref_std, ratio, attack, release, onset_fn, sr=44100, min_db=-40, comp_peak_norm=-10.0, min_th=-40,
                  max_ratio=20, percentile=75, chunk=8):
    """get_comp_matching (utils_data_normalization.py:357-429, expander off) for both channels of x float32 CUDA [2, T].
    The reference searches each channel on its own: peak-normalise to -10 dB, measure the mean inter-onset peak, and if it lies
    above the target band walk a (ratio, threshold) grid -- up to 15 x 61 compressor runs over the whole stem, each followed by
    the onset detector -- until the peak drops into the band.  Here the compressor runs are the FX chain's compressor kernel on
    the GPU, `chunk` thresholds per launch with both channels in lockstep (the candidate order is the same for both; each
    channel keeps the output of ITS first accepted candidate); the onset detection and the few dozen scalars per candidate stay
    on the host, as in the reference.  A channel whose measurement fails (no onset) stops the loop over channels like the
    reference's `except: break` (data_normalization.py:137-138): it and the following channel are returned untouched."""
    from .common_audioeffects import FX_COMP, N_PARAMS, fx_chain_forward
    T = x.shape[-1]
    peaks = row_absmax(x)
    out = x.clone()
    xn = torch.empty_like(x)
    todo = []
    for c in range(2):
        with np.errstate(divide='ignore'):
            if not _amp_to_db(float(peaks[c])) > min_db:
                continue                                      # below min_db: the channel is returned as it is (:428-429)
        k = np.power(10.0, comp_peak_norm / 20.0) / peaks[c]   # pyloudnorm.normalize.peak (:374)
        xn[c] = stereo_mix(x[c].reshape(1, 1, T).expand(1, 2, T).contiguous(), [[k, 0.0, 0.0, k]])[0, 0]
        m = mean_peak(xn[c].cpu().numpy(), onset_fn, sr, percentile)
        if m is None:
            break                                             # TypeError in the reference -> `except: break`
        out[c] = xn[c]
        if m[0] >= ref_peak + ref_std:
            todo.append(c)                                    # downward compression needed (:382)
        # inside the band, or below it with the expander off: the peak-normalised channel (:379-380, :425-426)
    if not todo:
        return out
    ratios = np.linspace(ratio, max_ratio, max_ratio - ratio + 1)
    ths = np.linspace(-1 - 9, min_th, 2 * np.abs(min_th) - 1 - 18)
    cands = [(rt, th) for rt in ratios for th in ths]
    src = torch.stack([xn[todo[0]], xn[todo[-1]]])            # one or two channels to search, as the (L, R) of a segment
    pending = set(todo)
    last = None
    for i0 in range(0, len(cands), chunk):
        part = cands[i0:i0 + chunk]
        P = np.zeros((len(part), N_PARAMS), np.float32)
        P[:, 13] = [th for _, th in part]
        P[:, 14], P[:, 15] = attack, release
        P[:, 16] = [rt for rt, _ in part]
        y = fx_chain_forward(src.unsqueeze(0).expand(len(part), 2, T).contiguous(), torch.from_numpy(P).to(x.device), FX_COMP, float(sr))
        # ratio > 1: the gain 10^(-y_l / 20) never exceeds 1 and |xn| <= 10^(-10/20), so the reference's clip to +-1 (:349-350) is idle
        y_host = y.cpu().numpy()
        for j in range(len(part)):
            for c in sorted(pending):
                row = 0 if c == todo[0] else 1
                m = mean_peak(y_host[j, row], onset_fn, sr, percentile)
                if m is None:
                    # the reference raises here and its caller breaks out of the channel loop: this channel keeps what it had
                    # before the effect, and so does every later channel
                    for cc in range(c, 2):
                        out[cc] = x[cc]
                    pending = {p for p in pending if p < c}
                    break
                if m[0] < ref_peak + ref_std:
                    out[c] = y[j, row]
                    pending.discard(c)
            if not pending:
                return out
        last = y[len(part) - 1]
    for c in pending:                                          # grid exhausted: the last candidate's output (:399)
        out[c] = last[0 if c == todo[0] else 1]
    return out


def _balance_gains(e1, e2, tgt_e1_bal, eps):
    """process_balance (normalization_imager.py:84-99) on energies: gains for the two signals."""
    total = e1 + e2
    g1 = np.sqrt(tgt_e1_bal * total / (e1 + eps))
    g2 = np.sqrt((total - e1 * g1 ** 2) / (e2 + 1e-3))
    return g1, g2


def imager_matrix(ll, rr, lr, target_side_mid_bal, eps=1e-4):
    """The three balance steps of normalize_imager (normalization_imager.py:54-76) as ONE 2x2 matrix on (L, R), from the input's
    (sum L^2, sum R^2, sum LR).  Energy of a linear combination a L + b R = a^2 ll + 2 a b lr + b^2 rr."""
    def energy(row):
        return row[0] ** 2 * ll + 2 * row[0] * row[1] * lr + row[1] ** 2 * rr
    to_ms = np.array([[1.0, 1.0], [1.0, -1.0]])
    to_lr = np.array([[0.5, 0.5], [0.5, -0.5]])
    M = np.eye(2)
    for basis, back, bal in ((to_ms, to_lr, float(target_side_mid_bal)), (np.eye(2), np.eye(2), 0.5), (to_ms, to_lr, float(target_side_mid_bal))):
        A = basis @ M                                   # the two signals this step balances, as rows over (L, R)
        g1, g2 = _balance_gains(energy(A[0]), energy(A[1]), bal, eps)
        M = back @ (np.diag([g1, g2]) @ A)
    return M


class Audio_Effects_Normalizer:
    def __init__(self, precomputed_feature_path, STEMS=['drums', 'bass', 'other', 'vocals'],
                 EFFECTS=['eq', 'compression', 'imager', 'loudness'], onset_detector=None):
        self.STEMS = STEMS          # Stems to be normalized
        self.EFFECTS = EFFECTS      # Effects to be normalized, order matters
        unsupported = [e for e in EFFECTS if e not in ('loudness', 'imager', 'eq', 'compression')]
        if unsupported:
            raise NotImplementedError(
                f"Audio_Effects_Normalizer: unknown effects {unsupported} (the reference knows 'eq', 'compression', 'imager', "
                "'loudness'; data_normalization.py:22)")
        # Audio settings
        self.SR = 44100
        self.SUBTYPE = 'PCM_16'
        # General Settings
        self.FFT_SIZE = FFT_SIZE
        self.HOP_LENGTH = self.FFT_SIZE // 4
        # Loudness
        self.NTAPS = 1001
        self.LUFS = -30
        self.MIN_DB = -40           # Min amplitude to apply the effects
        # Compressor (data_normalization.py:39-72): attack ms, release ms, starting ratio per stem
        self.COMP_PEAK_NORM = -10.0
        self.COMP_PERCENTILE = 75
        self.COMP_MIN_TH = -40
        self.COMP_MAX_RATIO = 20
        self.comp_settings = {'vocals': {'attack': 7.5, 'release': 400.0, 'ratio': 4}, 'drums': {'attack': 10.0, 'release': 180.0, 'ratio': 6},
                              'bass': {'attack': 10.0, 'release': 500.0, 'ratio': 5}, 'other': {'attack': 15.0, 'release': 666.0, 'ratio': 4}}
        # onset detector of the compressor matching: callable (x 1-D numpy, sr, window) -> onset sample positions; aubio's 'hfc' by default
        self.onset_detector = onset_detector if onset_detector is not None else aubio_onsets
        # Load Pre-computed Audio Effects Features
        if isinstance(precomputed_feature_path, dict):
            features_mean = {k: dict(v) for k, v in precomputed_feature_path.items()}
        else:
            features_mean = np.load(precomputed_feature_path, allow_pickle='TRUE')[()]
        self.features_mean = self.smooth_feature(features_mean)
        self.haas_rng = np.random.RandomState()

    # normalize current audio input with the order of designed audio FX
    def normalize_audio(self, audio, src):
        assert src in self.STEMS
        as_numpy = not isinstance(audio, torch.Tensor)
        if as_numpy:
            if not torch.cuda.is_available():
                raise RuntimeError("Audio_Effects_Normalizer (B200 engine): no CUDA device; there is no CPU fallback")
            a = np.asarray(audio)
            assert len(a.shape) == 2    # Always expects two dimensions
            if a.shape[1] == 1:         # Converts mono to stereo with repeated channels
                a = np.repeat(a, 2, axis=-1)
            x = torch.from_numpy(np.ascontiguousarray(a.T, dtype=np.float32)).cuda()
        else:
            x = _cabi.require_cuda_f32(audio, "normalizer input")
        for cur_effect in self.EFFECTS:
            x = self.normalize_audio_per_effect(x, src=src, effect=cur_effect)
        return np.ascontiguousarray(x.cpu().numpy().T) if as_numpy else x

    # normalize current audio input with current targeted audio FX; audio: float32 CUDA [2, T]
    def normalize_audio_per_effect(self, audio, src, effect):
        T = audio.shape[-1]
        track = torch.nn.functional.pad(audio, (self.FFT_SIZE, self.FFT_SIZE))
        peak = float(stereo_stats(track.unsqueeze(0))[0, 3])
        with np.errstate(divide='ignore'):
            max_db = 20 * np.log10(peak + 1e-30)
        if max_db > self.MIN_DB:
            if effect == 'eq':
                track = eq_matching(track, self.features_mean[effect][src], sr=self.SR, n_fft=self.FFT_SIZE,
                                    hop=self.HOP_LENGTH, min_db=self.MIN_DB, ntaps=self.NTAPS, lufs=self.LUFS)
            elif effect == 'compression':
                feat = self.features_mean[effect][src]
                assert len(feat) == 2
                cs = self.comp_settings[src]
                track = comp_matching(track, float(feat[0]), float(feat[1]), cs['ratio'], cs['attack'], cs['release'],
                                      self.onset_detector, sr=self.SR, min_db=self.MIN_DB, comp_peak_norm=self.COMP_PEAK_NORM,
                                      min_th=self.COMP_MIN_TH, max_ratio=self.COMP_MAX_RATIO, percentile=self.COMP_PERCENTILE)
            elif effect == 'loudness':
                track = lufs_normalize(track, self.SR, self.features_mean[effect][src])
            elif effect == 'imager':
                # threshold of applying Haas effects
                mono_threshold = 0.99 if src == 'bass' else 0.975
                track = self.normalize_imager(track, self.features_mean[effect][src], mono_threshold)
            else:
                raise NotImplementedError(effect)
        return track[:, self.FFT_SIZE:self.FFT_SIZE + T].contiguous()

    def smooth_feature(self, feature_dict_):
        """data_normalization.py:158-175: Savitzky-Golay smoothing of the target spectra (and panning features)."""
        import scipy.signal
        for effect in self.EFFECTS:
            for key in self.STEMS:
                if effect == 'eq':
                    f = 401 if key in ['other', 'vocals'] else 151
                    feature_dict_[effect][key] = scipy.signal.savgol_filter(feature_dict_[effect][key], f, 1, mode='mirror')
                elif effect == 'panning':
                    feature_dict_[effect][key] = scipy.signal.savgol_filter(feature_dict_[effect][key], 501, 1, mode='mirror')
        return feature_dict_

    def normalize_imager(self, data, target_side_mid_bal, mono_threshold, eps=1e-4):
        """normalization_imager.normalize_imager on data float32 CUDA [2, T]."""
        ll, rr, lr, _ = stereo_stats(data.unsqueeze(0))[0]
        mid_e, side_e = ll + rr + 2 * lr, max(ll + rr - 2 * lr, 0.0)
        # apply haas effect to almost-mono signal (:39-42); the reference draws delay / feedback / channel from its
        # Processor.randomize() (third-party distributions): uniform over the documented ranges here
        if mid_e / (mid_e + side_e) > mono_threshold:
            delay = int(self.haas_rng.randint(int(-0.040 * self.SR), int(0.040 * self.SR) + 1))
            feedback = float(self.haas_rng.uniform(0.33, 0.66))
            n = 2.0 * data.shape[-1]
            data = haas(data.unsqueeze(0), [delay], [feedback], [int(self.haas_rng.randint(2))])[0]
            hl, hr, hlr, _ = stereo_stats(data.unsqueeze(0))[0]
            # the chain applies it with RMS re-normalisation (AugmentationChain fxs=[(Haas, 1, True)], common_audioeffects.py:142-145)
            k = float(np.sqrt(((ll + rr) / n) / np.maximum(1e-7, (hl + hr) / n)))
            data = stereo_mix(data.unsqueeze(0), [[k, 0.0, 0.0, k]])[0]
            ll, rr, lr = hl * k * k, hr * k * k, hlr * k * k
        M = imager_matrix(ll, rr, lr, target_side_mid_bal, eps)
        return stereo_mix(data.unsqueeze(0), [[M[0, 0], M[0, 1], M[1, 0], M[1, 1]]])[0]
