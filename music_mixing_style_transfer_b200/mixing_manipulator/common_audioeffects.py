"""FX processors and the augmentation chain -- same surface as the reference's
`mixing_manipulator/common_audioeffects.py` for the four effects on the hot path (EQ, compressor, mid/side imager, gain).

Mirrors (reference paths relative to /root/reference/mixing_style_transfer/mixing_manipulator/):
  AugmentationChain   common_audioeffects.py:91-201
  Equaliser           :370-525    Compressor :590-661    MidSideImager :956-1007    Gain :1011-1051
  Parameter / ParameterList / Processor: the (un-vendored) pymixconsole classes the reference builds on (:24-27,40-88)

`Processor.process(x)` and `AugmentationChain.__call__(list of float32 [n, 2] arrays)` keep the reference semantics
(Bernoulli gates, optional randomisation, RMS re-normalisation, parallel dry/wet mix) but the DSP runs in the batched
sm_100a kernels of libmst_b200.so (csrc/fx.cu).  The batched tensor entry point `fx_chain_forward(x[B,2,L], params[B,20])`
is an ADDITION for GPU-side data augmentation; the list API stays.  No CPU path.

Parameter vector order (20 floats): EQ 0-12 (low_shelf gain,freq | first/second/third band gain,freq,q | high_shelf
gain,freq), compressor 13-16 (threshold, attack_time, release_time, ratio), imager 17 (bal), gain 18-19 (gain, invert).
"""
import ctypes
from random import shuffle
from typing import List, Optional, Tuple, Union

import numpy as np
import torch

from .. import _cabi
from .._cabi import FX_ALL, FX_COMP, FX_EQ, FX_GAIN, FX_IMAGER, FX_RMSNORM  # noqa: F401

N_PARAMS = _cabi.FX_NPARAMS
EQ_PARAM_NAMES = ['low_shelf_gain', 'low_shelf_freq',
                  'first_band_gain', 'first_band_freq', 'first_band_q',
                  'second_band_gain', 'second_band_freq', 'second_band_q',
                  'third_band_gain', 'third_band_freq', 'third_band_q',
                  'high_shelf_gain', 'high_shelf_freq']
COMP_PARAM_NAMES = ['threshold', 'attack_time', 'release_time', 'ratio']
# neutral parameter vector: every stage enabled by its mask only; values here are never read for disabled stages
_NEUTRAL = np.array([0, 80, 0, 400, .7, 0, 2000, .7, 0, 4000, .7, 0, 8000, -20, 2, 100, 4, 1, 0, 0], dtype=np.float32)


# ---------------------------------------------------------------------------------------------------------------------
# batched tensor entry point
# ---------------------------------------------------------------------------------------------------------------------
_ws_cache = {}


def fx_chain_forward(x: torch.Tensor, params: torch.Tensor, stages: int = FX_ALL, sample_rate: float = 44100.0,
                     out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """EQ -> compressor -> imager -> gain on a batch of stereo segments.
    x: float32 CUDA [B, 2, L]; params: float32 CUDA [B, 20]; stages: FX_* bit mask (FX_RMSNORM = the chain's RMS
    re-normalisation after EQ / comp / imager).  Returns float32 [B, 2, L]."""
    x = _cabi.require_cuda_f32(x, "fx input")
    params = _cabi.require_cuda_f32(params, "fx params")
    if x.dim() != 3 or x.shape[1] != 2:
        raise RuntimeError(f"fx_chain_forward expects [B, 2, L], got {tuple(x.shape)}")
    B, _, L = x.shape
    if tuple(params.shape) != (B, N_PARAMS):
        raise RuntimeError(f"fx params must be [{B}, {N_PARAMS}], got {tuple(params.shape)}")
    lib = _cabi.lib()
    y = torch.empty_like(x) if out is None else _cabi.require_cuda_f32(out, "fx output")
    nbytes = lib.mst_fx_workspace_bytes(B, L)
    key = (x.device, torch.cuda.current_stream().cuda_stream)
    ws = _ws_cache.get(key)
    if ws is None or ws.numel() < nbytes:
        ws = torch.empty(max(nbytes, 4096), dtype=torch.uint8, device=x.device)
        _ws_cache[key] = ws
    _cabi.check(lib.mst_fx_chain_forward(_cabi.ptr(x), _cabi.ptr(params), _cabi.ptr(y), B, L, float(sample_rate),
                                         int(stages), _cabi.ptr(ws), ws.numel(), _cabi.current_stream()),
                "fx_chain_forward")
    return y


def _process_numpy(x: np.ndarray, params: np.ndarray, stages: int, sample_rate) -> np.ndarray:
    """list-API helper: float32 [n, 2] (time-major like the reference) -> device -> kernels -> float32 [n, 2]."""
    if not torch.cuda.is_available():
        raise RuntimeError("mixing_manipulator (B200 engine): no CUDA device; there is no CPU fallback")
    x = np.asarray(x)
    if x.ndim != 2 or x.shape[1] != 2:
        raise ValueError(f"expected audio of shape [n_samples, 2], got {x.shape}")
    xt = torch.from_numpy(np.ascontiguousarray(x.T, dtype=np.float32)).unsqueeze(0).cuda()
    pt = torch.from_numpy(np.asarray(params, dtype=np.float32).reshape(1, N_PARAMS)).cuda()
    y = fx_chain_forward(xt, pt, stages, float(sample_rate if sample_rate else 44100.0))
    return np.ascontiguousarray(y[0].cpu().numpy().T)


# ---------------------------------------------------------------------------------------------------------------------
# pymixconsole-style parameter containers (the reference gets these from the un-vendored pymixconsole package)
# ---------------------------------------------------------------------------------------------------------------------
class Parameter:
    def __init__(self, name, value, kind, processor=None, units="", minimum=None, maximum=None, options=None):
        self.name, self.value, self.kind = name, value, kind
        self.processor, self.units = processor, units
        self.min, self.max, self.options = minimum, maximum, options

    def randomize(self):
        """Uniform over [min, max] (fair coin for bools).  The exact sampling distributions of pymixconsole are
        third-party and unpinned; the batched path takes explicit parameter tensors instead."""
        if self.kind == 'bool':
            self.value = bool(np.random.rand() < 0.5)
        elif self.kind in ('float', 'int') and self.min is not None and self.max is not None:
            v = self.min + np.random.rand() * (self.max - self.min)
            self.value = int(round(v)) if self.kind == 'int' else float(v)
        elif self.kind == 'string' and self.options:
            self.value = self.options[np.random.randint(len(self.options))]

    def __repr__(self):
        return f"Parameter({self.name}={self.value})"


class ParameterList:
    def __init__(self):
        self._names = []

    def add(self, parameter):
        setattr(self, parameter.name, parameter)
        self._names.append(parameter.name)

    def __iter__(self):
        return iter(getattr(self, n) for n in self._names)

    def __repr__(self):
        return "ParameterList(" + ", ".join(repr(p) for p in self) + ")"


class Processor:
    """Base processor with the reference's patched constructor (common_audioeffects.py:40-88)."""

    def __init__(self, name, parameters, block_size, sample_rate, dtype='float32'):
        self.name = name
        self.parameters = parameters
        self.block_size = block_size
        self.sample_rate = sample_rate
        self.dtype = dtype

    def __repr__(self):
        return f'Processor(name={self.name!r}, parameters={self.parameters!r}'

    def update(self, parameter_name):
        pass

    def randomize(self):
        for p in self.parameters:
            p.randomize()
            self.update(p.name)

    # slice of the 20-float parameter vector this processor owns, and its stage bit
    STAGE = 0

    def param_vector(self) -> np.ndarray:
        raise NotImplementedError

    def process(self, x):
        return _process_numpy(x, self.param_vector(), self.STAGE, self.sample_rate)


_ALL_BANDS = ['low_shelf', 'first_band', 'second_band', 'third_band', 'high_shelf']


# %%%%%%%%%%%%%%%%%%%%%%%%%%%%%%%% EQUALISER %%%%%%%%%%%%%%%%%%%%%%%%%%%%%%%%%%%%
class Equaliser(Processor):
    """Five band parametric equaliser (two shelves and three central bands): cascade of five RBJ biquads."""
    STAGE = FX_EQ

    def __init__(self, n_channels, sample_rate, gain_range=(-15.0, 15.0), q_range=(0.1, 2.0),
                 bands=['low_shelf', 'first_band', 'second_band', 'third_band', 'high_shelf'],
                 hard_clip=False, name='Equaliser', parameters=None):
        super().__init__(name, parameters=parameters, block_size=None, sample_rate=sample_rate)
        if n_channels != 2:
            raise NotImplementedError("Equaliser: the B200 path is stereo (n_channels=2) like the chain factory builds it")
        unknown = [b for b in bands if b not in _ALL_BANDS]
        if unknown:
            raise ValueError(f"Equaliser: unknown bands {unknown}")
        if list(bands) != _ALL_BANDS:
            # a subset of the bands (the per-instrument factory builds one-shelf equalisers, audio_effects_chain.py:125-146):
            # explicit RBJ coefficients through mst_biquad_cascade instead of the 13-parameter fused stage
            self.STAGE = 0
        self.n_channels = n_channels
        MIN_GAIN, MAX_GAIN = gain_range
        MIN_Q, MAX_Q = q_range
        if not parameters:
            self.parameters = ParameterList()
            self.parameters.add(Parameter('low_shelf_gain', 0.0, 'float', minimum=MIN_GAIN, maximum=MAX_GAIN))
            self.parameters.add(Parameter('low_shelf_freq', 80.0, 'float', minimum=30.0, maximum=200.0))
            self.parameters.add(Parameter('first_band_gain', 0.0, 'float', minimum=MIN_GAIN, maximum=MAX_GAIN))
            self.parameters.add(Parameter('first_band_freq', 400.0, 'float', minimum=200.0, maximum=1000.0))
            self.parameters.add(Parameter('first_band_q', 0.7, 'float', minimum=MIN_Q, maximum=MAX_Q))
            self.parameters.add(Parameter('second_band_gain', 0.0, 'float', minimum=MIN_GAIN, maximum=MAX_GAIN))
            self.parameters.add(Parameter('second_band_freq', 2000.0, 'float', minimum=1000.0, maximum=3000.0))
            self.parameters.add(Parameter('second_band_q', 0.7, 'float', minimum=MIN_Q, maximum=MAX_Q))
            self.parameters.add(Parameter('third_band_gain', 0.0, 'float', minimum=MIN_GAIN, maximum=MAX_GAIN))
            self.parameters.add(Parameter('third_band_freq', 4000.0, 'float', minimum=3000.0, maximum=8000.0))
            self.parameters.add(Parameter('third_band_q', 0.7, 'float', minimum=MIN_Q, maximum=MAX_Q))
            self.parameters.add(Parameter('high_shelf_gain', 0.0, 'float', minimum=MIN_GAIN, maximum=MAX_GAIN))
            self.parameters.add(Parameter('high_shelf_freq', 8000.0, 'float', minimum=5000.0, maximum=10000.0))
        self.bands = bands
        self.hard_clip = hard_clip

    def param_vector(self):
        v = _NEUTRAL.copy()
        for i, n in enumerate(EQ_PARAM_NAMES):
            v[i] = getattr(self.parameters, n).value
        return v

    def reset_state(self):
        """Filters start from zero state at every `process` call (common_audioeffects.py:512); nothing to reset."""

    def _band_coefficients(self):
        """(b0, b1, b2, a1, a2) per band of `self.bands`, RBJ cookbook, float64 (setup_filters, common_audioeffects.py:438-462)."""
        rows = []
        for band in self.bands:
            G = float(getattr(self.parameters, band + '_gain').value)
            fc = float(getattr(self.parameters, band + '_freq').value)
            shelf = band in ('low_shelf', 'high_shelf')
            Q = 0.707 if shelf else float(getattr(self.parameters, band + '_q').value)
            A = 10.0 ** (G / 40.0)
            w0 = 2.0 * np.pi * (fc / self.sample_rate)
            alpha, c = np.sin(w0) / (2.0 * Q), np.cos(w0)
            sq = 2.0 * np.sqrt(A) * alpha
            if not shelf:
                b = [1.0 + alpha * A, -2.0 * c, 1.0 - alpha * A]
                a_ = [1.0 + alpha / A, -2.0 * c, 1.0 - alpha / A]
            elif band == 'low_shelf':
                b = [A * ((A + 1) - (A - 1) * c + sq), 2 * A * ((A - 1) - (A + 1) * c), A * ((A + 1) - (A - 1) * c - sq)]
                a_ = [(A + 1) + (A - 1) * c + sq, -2 * ((A - 1) + (A + 1) * c), (A + 1) + (A - 1) * c - sq]
            else:
                b = [A * ((A + 1) + (A - 1) * c + sq), -2 * A * ((A - 1) + (A + 1) * c), A * ((A + 1) + (A - 1) * c - sq)]
                a_ = [(A + 1) - (A - 1) * c + sq, 2 * ((A - 1) - (A + 1) * c), (A + 1) - (A - 1) * c - sq]
            rows.append([b[0] / a_[0], b[1] / a_[0], b[2] / a_[0], a_[1] / a_[0], a_[2] / a_[0]])
        return np.asarray(rows, dtype=np.float64)

    def process(self, x):
        if self.STAGE:
            y = super().process(x)
        else:
            xt = _to_device(x)
            lib = _cabi.lib()
            n = xt.shape[-1]
            coef = torch.from_numpy(self._band_coefficients()[None]).to(xt.device)
            ws = torch.empty(max(lib.mst_fx_workspace_bytes(1, n), 4096), dtype=torch.uint8, device=xt.device)
            yt = torch.empty_like(xt)
            _cabi.check(lib.mst_biquad_cascade(_cabi.ptr(xt), coef.data_ptr(), len(self.bands), _cabi.ptr(yt), 1, n, _cabi.ptr(ws),
                                               ws.numel(), _cabi.current_stream()), "biquad_cascade")
            y = np.ascontiguousarray(yt[0].cpu().numpy().T)
        if self.hard_clip:
            y = np.clip(y, -1.0, 1.0)
        return y


# %%%%%%%%%%%%%%%%%%%%%%%%%%%%%%%% COMPRESSOR %%%%%%%%%%%%%%%%%%%%%%%%%%%%%%%%
class Compressor(Processor):
    """Single band stereo dynamic range compressor (threshold, attack_time, release_time, ratio; makeup 0)."""
    STAGE = FX_COMP

    def __init__(self, sample_rate, name='Compressor', parameters=None):
        super().__init__(name=name, parameters=parameters, block_size=None, sample_rate=sample_rate)
        if not parameters:
            self.parameters = ParameterList()
            self.parameters.add(Parameter('threshold', -20.0, 'float', units='dB', minimum=-80.0, maximum=-5.0))
            self.parameters.add(Parameter('attack_time', 2.0, 'float', units='ms', minimum=1., maximum=20.0))
            self.parameters.add(Parameter('release_time', 100.0, 'float', units='ms', minimum=50.0, maximum=500.0))
            self.parameters.add(Parameter('ratio', 4.0, 'float', minimum=4., maximum=40.0))
        self.yL_prev = None  # the reference zeroes the envelope state inside every call (:553)

    def param_vector(self):
        v = _NEUTRAL.copy()
        for i, n in enumerate(COMP_PARAM_NAMES):
            v[13 + i] = getattr(self.parameters, n).value
        return v

    def update(self, parameter_name=None):
        self.yL_prev = None


# %%%%%%%%%%%%%%%%%%%%%%%%%%%%%% STEREO IMAGER %%%%%%%%%%%%%%%%%%%%%%%%%%%%%%%%%%
class MidSideImager(Processor):
    STAGE = FX_IMAGER

    def __init__(self, name='IMAGER', parameters=None):
        super().__init__(name, parameters=parameters, block_size=None, sample_rate=None)
        if not parameters:
            self.parameters = ParameterList()
            # 0.0~1.0 : more centered, 1.0~2.0 : wider
            self.parameters.add(Parameter("bal", 0.0, "float", processor=self, minimum=0.0, maximum=2.0))

    def param_vector(self):
        v = _NEUTRAL.copy()
        v[17] = getattr(self.parameters, "bal").value
        return v

    def update(self, parameter_name=None):
        return parameter_name


# %%%%%%%%%%%%%%%%%%%%%%%%%%%%%%%%% GAIN %%%%%%%%%%%%%%%%%%%%%%%%%%%%%%%%%%
class Gain(Processor):
    """Applies gain in dB and can also invert polarity."""
    STAGE = FX_GAIN

    def __init__(self, name='Gain', parameters=None):
        super().__init__(name, parameters=parameters, block_size=None, sample_rate=None)
        if not parameters:
            self.parameters = ParameterList()
            self.parameters.add(Parameter('gain', 1.0, 'float', units='dB', minimum=-6.0, maximum=9.0))
            self.parameters.add(Parameter('invert', False, 'bool'))

    def param_vector(self):
        v = _NEUTRAL.copy()
        v[18] = self.parameters.gain.value
        v[19] = 1.0 if self.parameters.invert.value else 0.0
        return v


# %%%%%%%%%%%%%%%%%%%%%%%%%%%%% HAAS EFFECT %%%%%%%%%%%%%%%%%%%%%%%%%%%%%%%%%%%%%%%
class Haas(Processor):
    """Haas Effect Processor (common_audioeffects.py:790-850): one channel gets a delayed copy of itself added,
    y[:, ch] += feedback * roll(x[:, ch], delay) (:767-787); kernel mst_haas."""
    STAGE = 0

    def __init__(self, sample_rate, delay_range=(-0.040, 0.040), name='Haas', parameters=None):
        super().__init__(name=name, parameters=parameters, block_size=None, sample_rate=sample_rate)
        if not parameters:
            self.parameters = ParameterList()
            self.parameters.add(Parameter('delay', int(delay_range[1] * sample_rate), 'int', units='samples',
                                          minimum=int(delay_range[0] * sample_rate),
                                          maximum=int(delay_range[1] * sample_rate)))
            self.parameters.add(Parameter('feedback', 0.35, 'float', minimum=0.33, maximum=0.66))
            self.parameters.add(Parameter('wet_channel', 'left', 'string', options=['left', 'right']))

    def process(self, x):
        from .data_normalization import haas
        x = np.asarray(x)
        assert x.shape[1] == 1 or x.shape[1] == 2, 'Haas effect only works with monaural or stereo audio.'
        if x.shape[1] < 2:
            x = np.repeat(x, 2, axis=1)
        xt = _to_device(x)
        ch = {'left': 0, 'right': 1}.get(self.parameters.wet_channel.value, -1)     # any other string: no wet channel (:781-784)
        y = haas(xt, [int(self.parameters.delay.value)], [float(self.parameters.feedback.value)], [ch])
        return np.ascontiguousarray(y[0].cpu().numpy().T)


# %%%%%%%%%%%%%%%%%%%%%%%%%%%%%% PANNER %%%%%%%%%%%%%%%%%%%%%%%%%%%%%%%%%%
class Panner(Processor):
    """Simple stereo panner (common_audioeffects.py:854-952): per-channel gains from the pan law; kernel mst_stereo_mix."""
    STAGE = 0

    def __init__(self, name='Panner', parameters=None):
        super().__init__(name=name, parameters=parameters, block_size=None, sample_rate=None)
        if not parameters:
            self.parameters = ParameterList()
            self.parameters.add(Parameter('pan', 0.5, 'float', minimum=0., maximum=1.))
            self.parameters.add(Parameter('pan_law', '-4.5dB', 'string', options=['-4.5dB', 'linear', 'constant_power']))
        self.update()

    def _calculate_pan_coefficents(self):
        self.gains = np.zeros(2, dtype=self.dtype)
        theta = self.parameters.pan.value * (np.pi / 2)         # [0, 1] -> [0, pi/2]
        law = self.parameters.pan_law.value
        if law == 'linear':
            self.gains[0] = ((np.pi / 2) - theta) * (2 / np.pi)
            self.gains[1] = theta * (2 / np.pi)
        elif law == 'constant_power':
            self.gains[0] = np.cos(theta)
            self.gains[1] = np.sin(theta)
        elif law == '-4.5dB':
            self.gains[0] = np.sqrt(((np.pi / 2) - theta) * (2 / np.pi) * np.cos(theta))
            self.gains[1] = np.sqrt(theta * (2 / np.pi) * np.sin(theta))
        else:
            raise ValueError(f'Invalid pan_law {law}.')

    def update(self, parameter_name=None):
        self._calculate_pan_coefficents()

    def process(self, x):
        from .data_normalization import stereo_mix
        x = np.asarray(x)
        assert x.shape[1] == 1 or x.shape[1] == 2, 'Panner only works with monaural or stereo audio.'
        if x.shape[1] < 2:
            x = np.repeat(x, 2, axis=1)
        y = stereo_mix(_to_device(x), [[float(self.gains[0]), 0.0, 0.0, float(self.gains[1])]])
        return np.ascontiguousarray(y[0].cpu().numpy().T)


# %%%%%%%%%%%%%%%%%%%%%%%%%% CONVOLUTIONAL REVERB %%%%%%%%%%%%%%%%%%%%%%%%%%%%%
class ConvolutionalReverb(Processor):
    """Convolutional reverb (common_audioeffects.py:665-764): the input convolved with a sampled impulse response, cut at the
    response's peak (+ pre-delay) and mixed with the dry signal.  `impulse_responses` is the reference's structure: a list (one
    entry per RT60 class) of lists of dicts whose 'impulse_response' entry is a callable returning float [m, 1 or 2].  The
    convolution is mst_fft_convolve (partitioned overlap-add on the device FFT)."""
    STAGE = 0

    def __init__(self, impulse_responses, sample_rate, name='ConvolutionalReverb', parameters=None):
        super().__init__(name=name, parameters=parameters, block_size=None, sample_rate=sample_rate)
        if impulse_responses is None:
            raise ValueError('List of impulse responses must be provided for ConvolutionalReverb processor.')
        self.impulse_responses = impulse_responses
        if not parameters:
            self.parameters = ParameterList()
            self.max_ir_num = len(max(impulse_responses, key=len))
            self.parameters.add(Parameter('index', 0, 'int', minimum=0, maximum=len(impulse_responses)))
            self.parameters.add(Parameter('index_ir', 0, 'int', minimum=0, maximum=self.max_ir_num))
            self.parameters.add(Parameter('wet', 1.0, 'float', minimum=1.0, maximum=1.0))
            self.parameters.add(Parameter('dry', 0.0, 'float', minimum=0.0, maximum=0.0))
            self.parameters.add(Parameter('decay', 1.0, 'float', minimum=1.0, maximum=1.0))
            self.parameters.add(Parameter('pre_delay', 0, 'int', units='ms', minimum=0, maximum=0))
        self.h = None

    def update(self, parameter_name=None):
        """Pick the impulse response (RT60 class `index`, response `index_ir` modulo the class size) and apply the decay fade
        (:712-733)."""
        ir_class = self.impulse_responses[min(int(self.parameters.index.value), len(self.impulse_responses) - 1)]
        self.h = np.copy(ir_class[int(self.parameters.index_ir.value) % len(ir_class)]['impulse_response']())
        decay = self.parameters.decay.value
        if decay < 1.:
            n_h = self.h.shape[0]
            peak = int(np.argmax(np.max(np.abs(self.h), axis=1), axis=0))
            fstart = min(n_h, peak + int(decay * (n_h - peak)))
            fstop = min(n_h, fstart + int(0.020 * self.sample_rate))          # constant 20 ms fade out
            flen = fstop - fstart
            fade = np.power(0.1, np.arange(1, flen + 1, dtype=self.dtype) / flen * 5)
            self.h[fstart:fstop, :] *= fade[:, np.newaxis]
            self.h = self.h[:fstop]

    def process(self, x):
        x = np.asarray(x)
        n_channels = x.shape[1]
        if self.h is None:
            self.update()
        if self.h.shape[1] > 1 and n_channels == 1:
            self.h = self.h[:, np.random.randint(self.h.shape[1]), np.newaxis]      # randomly choose one IR channel (:740-741)
        if self.parameters.wet.value == 0.0:
            return x
        h = self.h
        idx = int(np.argmax(np.max(np.abs(h), axis=1), axis=0))                     # the response's own delay (:751-757)
        idx += int(0.001 * np.abs(self.parameters.pre_delay.value) * self.sample_rate)
        idx = int(np.clip(idx, 0, h.shape[0] - 1))
        xs = np.repeat(x, 2, axis=1) if n_channels == 1 else x
        xt = _to_device(xs)[0]                                                       # [2, n]
        ht = torch.from_numpy(np.ascontiguousarray(h.T, dtype=np.float32)).cuda()    # [1 or 2, m]
        lib = _cabi.lib()
        T, M = xt.shape[-1], ht.shape[-1]
        ws = torch.empty(lib.mst_fft_convolve_workspace_bytes(T, M), dtype=torch.uint8, device=xt.device)
        y = torch.empty_like(xt)
        _cabi.check(lib.mst_fft_convolve(_cabi.ptr(xt, True), T, max(xt.stride(0), T), _cabi.ptr(ht, True), M, max(ht.stride(0), M), ht.shape[0], idx,
                                         float(self.parameters.dry.value), float(self.parameters.wet.value), _cabi.ptr(y),
                                         y.stride(0), _cabi.ptr(ws), ws.numel(), _cabi.current_stream()), "fft_convolve")
        out = np.ascontiguousarray(y.cpu().numpy().T)
        return out[:, :1] if n_channels == 1 else out


# %%%%%%%%%%%%%%%%%%%%%%%%%%%%%% ALGORITHMIC REVERB %%%%%%%%%%%%%%%%%%%%%%%%%%%%%%%%%%
class AlgorithmicReverb(Processor):
    """Freeverb-style comb / all-pass reverb (common_audioeffects.py:1429-1536); kernel mst_algo_reverb (csrc/reverb.cu).
    Difference to the reference: every `process` call starts from silent delay lines.  The reference rebuilds its filters in
    `update` (i.e. once per chain call) but not between the arrays of one list, so there the tail of one array rings into the
    next one."""
    STAGE = 0

    def __init__(self, name="algoreverb", parameters=None, sample_rate=44100, **kwargs):
        super().__init__(name=name, parameters=parameters, block_size=None, sample_rate=sample_rate)
        if not parameters:
            self.parameters = ParameterList()
            self.parameters.add(Parameter("room_size", 0.5, "float", minimum=0.05, maximum=0.85))
            self.parameters.add(Parameter("damping", 0.1, "float", minimum=0.0, maximum=1.0))
            self.parameters.add(Parameter("dry_mix", 0.9, "float", minimum=0.0, maximum=1.0))
            self.parameters.add(Parameter("wet_mix", 0.1, "float", minimum=0.0, maximum=1.0))
            self.parameters.add(Parameter("width", 0.7, "float", minimum=0.0, maximum=1.0))
        # Tuning
        self.stereospread = 23
        self.scalegain = 0.2

    def process(self, data):
        data = np.asarray(data)
        if data.ndim < 2:
            data = data[:, None]
        xs = data if data.shape[1] == 2 else np.repeat(data[:, :1], 2, axis=1)      # mono: both sides read channel 0 (:1448-1456)
        xt = _to_device(xs)
        pr = self.parameters
        p5 = torch.tensor([[pr.room_size.value, pr.damping.value, pr.dry_mix.value, pr.wet_mix.value, pr.width.value]],
                          dtype=torch.float32, device=xt.device)
        lib = _cabi.lib()
        L = xt.shape[-1]
        ws = torch.empty(lib.mst_algo_reverb_workspace_bytes(1, L), dtype=torch.uint8, device=xt.device)
        y = torch.empty_like(xt)
        _cabi.check(lib.mst_algo_reverb(_cabi.ptr(xt), _cabi.ptr(p5), _cabi.ptr(y), 1, L, _cabi.ptr(ws), ws.numel(),
                                        _cabi.current_stream()), "algo_reverb")
        return np.ascontiguousarray(y[0].cpu().numpy().T).astype(np.float64)          # the reference returns float64 (:1458)


def _to_device(x):
    """float [n, 2] numpy (time-major, as the reference passes audio) -> float32 CUDA [1, 2, n]."""
    if not torch.cuda.is_available():
        raise RuntimeError("mixing_manipulator (B200 engine): no CUDA device; there is no CPU fallback")
    return torch.from_numpy(np.ascontiguousarray(np.asarray(x).T, dtype=np.float32)).unsqueeze(0).cuda()


_STAGE_SLICE = {FX_EQ: slice(0, 13), FX_COMP: slice(13, 17), FX_IMAGER: slice(17, 18), FX_GAIN: slice(18, 20)}


def _top_stage(mask):
    return max(s for s in (FX_EQ, FX_COMP, FX_IMAGER, FX_GAIN) if mask & s) if mask else 0


class _DeviceList:
    """The chain's list of [n, 2] arrays, either on the host (numpy) or on the device (float32 [2, n] tensors)."""

    def __init__(self, x_list):
        self._host, self._dev = list(x_list), None

    def host(self):
        if self._dev is not None:
            self._host = [np.ascontiguousarray(t.cpu().numpy().T) for t in self._dev]
            self._dev = None
        return self._host

    def set_host(self, y_list):
        self._host, self._dev = list(y_list), None

    def run_stages(self, stages, rms_normalize, params, *sample_rates):
        if self._dev is None:
            if not torch.cuda.is_available():
                raise RuntimeError("mixing_manipulator (B200 engine): no CUDA device; there is no CPU fallback")
            for x in self._host:
                if np.asarray(x).ndim != 2 or np.asarray(x).shape[1] != 2:
                    raise ValueError(f"expected audio of shape [n_samples, 2], got {np.asarray(x).shape}")
            self._dev = [torch.from_numpy(np.ascontiguousarray(np.asarray(x).T, dtype=np.float32)).cuda() for x in self._host]
        mask = stages | (FX_RMSNORM if rms_normalize else 0)
        sr = float(sample_rates[0]) if sample_rates else 44100.0
        by_len = {}
        for i, t in enumerate(self._dev):
            by_len.setdefault(t.shape[-1], []).append(i)
        for idx in by_len.values():
            batch = torch.stack([self._dev[i] for i in idx], dim=0)
            p = torch.from_numpy(np.tile(np.asarray(params, np.float32), (len(idx), 1))).to(batch.device)
            y = fx_chain_forward(batch, p, mask, sr)
            for k, i in enumerate(idx):
                self._dev[i] = y[k]


# ---------------------------------------------------------------------------------------------------------------------
class AugmentationChain:
    """Basic audio Fx chain which is used for data augmentation (common_audioeffects.py:91-201)."""

    def __init__(self,
                 fxs: Optional[List[Tuple[Union[Processor, 'AugmentationChain'], float, bool]]] = [],
                 shuffle: Optional[bool] = False,
                 parallel: Optional[bool] = False,
                 parallel_weight_factor=None,
                 randomize_param_value=True):
        self.fxs = fxs
        self.shuffle = shuffle
        self.parallel = parallel
        self.parallel_weight_factor = parallel_weight_factor
        self.randomize_param_value = randomize_param_value

    def apply_processor(self, x, processor: Processor, rms_normalize):
        """One effect (+ RMS re-normalisation, :142-145) on one [n, 2] array; the normalisation is fused in the kernel."""
        unfused = (rms_normalize and processor.STAGE == FX_GAIN) or getattr(processor, "hard_clip", False) or processor.STAGE == 0
        if unfused:
            # Gain is never normalised by the factory (audio_effects_chain.py:92) and hard_clip is off by default:
            # rare combinations take the effect kernel, then the reference's own normalisation arithmetic on the host
            y = processor.process(x)
            if rms_normalize:
                y = y * np.sqrt(np.mean(np.square(x)) / np.maximum(1e-7, np.mean(np.square(y))))
            return y
        stages = processor.STAGE | (FX_RMSNORM if rms_normalize else 0)
        return _process_numpy(x, processor.param_vector(), stages, processor.sample_rate)

    def apply_same_processor(self, x_list, processor: Processor, rms_normalize):
        for i in range(len(x_list)):
            x_list[i] = self.apply_processor(x_list[i], processor, rms_normalize)
        return x_list

    def __call__(self, x_list):
        """Same semantics and the same order of `np.random` draws as the reference (:156-192): per effect a Bernoulli gate,
        then the randomisation of its parameters.  The DSP is batched: the gated EQ / compressor / imager / gain stages that
        follow each other in the kernel's stage order are ONE `fx_chain_forward` launch chain over the whole list (arrays of
        equal length share a batch), and the list crosses PCIe once in each direction instead of once per effect and array."""
        # randomly shuffle effect order if `self.shuffle` is True
        if self.shuffle:
            shuffle(self.fxs)
        state = _DeviceList(x_list)
        run = None                                   # pending fused run: [stage mask, rms flag of its normalised stages, params]
        for fx, p, rms_normalize in self.fxs:
            if np.random.rand() < p:
                if isinstance(fx, Processor):
                    # randomize all effect parameters (also calls `update()` for each processor)
                    if self.randomize_param_value:
                        fx.randomize()
                    else:
                        fx.update(None)
                    fusable = fx.STAGE != 0 and not getattr(fx, "hard_clip", False) and not (fx.STAGE == FX_GAIN and rms_normalize)
                    if fusable:
                        norm = None if fx.STAGE == FX_GAIN else bool(rms_normalize)     # the gain stage is never normalised
                        if run is not None and (fx.STAGE <= _top_stage(run[0]) or (norm is not None and run[1] not in (None, norm))):
                            state.run_stages(*run)
                            run = None
                        if run is None:
                            run = [0, None, _NEUTRAL.copy()]
                        run[0] |= fx.STAGE
                        run[1] = norm if norm is not None else run[1]
                        run[2][_STAGE_SLICE[fx.STAGE]] = fx.param_vector()[_STAGE_SLICE[fx.STAGE]]
                        run.append(fx.sample_rate) if fx.sample_rate else None
                    else:
                        if run is not None:
                            state.run_stages(*run)
                            run = None
                        state.set_host(self.apply_same_processor(state.host(), fx, rms_normalize))
                else:
                    if run is not None:
                        state.run_stages(*run)
                        run = None
                    state.set_host(fx(state.host()))
        if run is not None:
            state.run_stages(*run)
        y_list = state.host()
        if self.parallel:
            # weighting factor of input signal in the range of (0.0 ~ 0.5)
            weight_in = self.parallel_weight_factor if self.parallel_weight_factor else np.random.rand() / 2.
            for i in range(len(y_list)):
                y_list[i] = weight_in * x_list[i] + (1 - weight_in) * y_list[i]
        return y_list

    def param_tensor(self, batch: int = 1) -> torch.Tensor:
        """Current parameter values of the chain's processors as a [batch, 20] float32 tensor (host)."""
        v = _NEUTRAL.copy()
        for fx, _, _ in self.fxs:
            if isinstance(fx, Processor) and fx.STAGE:
                pv = fx.param_vector()
                v[_STAGE_SLICE[fx.STAGE]] = pv[_STAGE_SLICE[fx.STAGE]]
        return torch.from_numpy(np.tile(v, (batch, 1)))

    def __repr__(self):
        return f'AugmentationChain(fxs={self.fxs!r}, shuffle={self.shuffle!r})'
