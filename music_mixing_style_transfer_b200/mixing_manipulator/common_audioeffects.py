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
        if list(bands) != ['low_shelf', 'first_band', 'second_band', 'third_band', 'high_shelf']:
            raise NotImplementedError("Equaliser: only the default five bands have a B200 path")
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

    def process(self, x):
        y = super().process(x)
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
        unfused = (rms_normalize and processor.STAGE == FX_GAIN) or getattr(processor, "hard_clip", False)
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
        # randomly shuffle effect order if `self.shuffle` is True
        if self.shuffle:
            shuffle(self.fxs)
        # apply effects with probabilities given in `self.fxs`
        y_list = x_list.copy()
        for fx, p, rms_normalize in self.fxs:
            if np.random.rand() < p:
                if isinstance(fx, Processor):
                    # randomize all effect parameters (also calls `update()` for each processor)
                    if self.randomize_param_value:
                        fx.randomize()
                    else:
                        fx.update(None)
                    y_list = self.apply_same_processor(y_list, fx, rms_normalize)
                else:
                    y_list = fx(y_list)
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
            if isinstance(fx, Processor):
                pv = fx.param_vector()
                sl = {FX_EQ: slice(0, 13), FX_COMP: slice(13, 17), FX_IMAGER: slice(17, 18), FX_GAIN: slice(18, 20)}[fx.STAGE]
                v[sl] = pv[sl]
        return torch.from_numpy(np.tile(v, (batch, 1)))

    def __repr__(self):
        return f'AugmentationChain(fxs={self.fxs!r}, shuffle={self.shuffle!r})'
