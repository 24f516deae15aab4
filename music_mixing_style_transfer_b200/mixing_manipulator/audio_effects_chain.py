"""Effects-chain factories -- same surface as the reference's `mixing_manipulator/audio_effects_chain.py`:
`create_effects_augmentation_chain` (:17-95) over eq / comp / pan / imager / gain / algorithmic and convolutional reverb, and
the per-instrument `create_inst_effects_augmentation_chain` (:99-164) with its nested and parallel chains.
'expand' raises: the reference's factory names an `Expander` class that does not exist in its common_audioeffects.py."""
import os
from glob import glob

import numpy as np

from .common_audioeffects import (AlgorithmicReverb, AugmentationChain, Compressor, ConvolutionalReverb, Equaliser, Gain,
                                  MidSideImager, Panner, Parameter, ParameterList, Processor)


def load_impulse_responses(ir_dir_path, sample_rate):
    """The impulse-response data set of the convolutional reverb, as the reference's factory assembles it (:58-82):
    `<ir_dir_path>*/RT60_avg/<lo-hi>/<name>/impulse_response.wav`; one list per RT60 class below 3000 ms, all longer classes
    merged into a last one.  Entries are dicts {'impulse_response': callable -> float32 [m, channels]} (create_dataset /
    generate_data, common_dataprocessing.py:110-204, 318-390: integer PCM scaled by 1 / (1 + max))."""
    import functools
    import scipy.io.wavfile as wavfile

    def as_float(samples):
        return samples.astype(dtype=np.float32) * (1. / (1. + np.iinfo(samples.dtype).max))

    classes = {}
    for rt_path in glob(f"{ir_dir_path}*/RT60_avg/[!0-]*"):
        cur_rt = rt_path.split('/')[-1]
        for d in os.listdir(rt_path):
            merged = None
            for f in os.listdir(os.path.join(rt_path, d)):
                if os.path.splitext(f)[0] != 'impulse_response':
                    continue
                fs, samples = wavfile.read(os.path.join(rt_path, d, f))
                samples = samples[:, np.newaxis] if samples.ndim == 1 else samples
                assert samples.dtype == np.int16 or samples.dtype == np.int32
                if fs != sample_rate:
                    raise ValueError(f'File has fs = {fs}Hz but expected {[sample_rate]}Hz.')
                merged = samples if merged is None else np.vstack((samples, merged))
            if merged is not None:
                classes.setdefault(cur_rt, []).append({'impulse_response': functools.partial(as_float, merged)})
    ir_list, long_irs = [], []
    for cur_rt, entries in classes.items():
        if int(cur_rt.split('-')[0]) < 3000:
            ir_list.append(entries)
        else:
            long_irs.extend(entries)
    ir_list.append(long_irs)
    return ir_list


# create augmentation effects chain according to targeted effects with their applying probability
def create_effects_augmentation_chain(effects,
                                      ir_dir_path=None,
                                      sample_rate=44100,
                                      shuffle=False,
                                      parallel=False,
                                      parallel_weight_factor=None):
    '''
        Args:
            effects (list of tuples or string) : First tuple element is string denoting the target effects.
                                                    Second tuple element is probability of applying current effects.
            ir_dir_path (string) : directory path that contains directories of impulse responses organized according to RT60
            sample_rate (int) : using sampling rate
            shuffle (boolean) : shuffle FXs inside current FX chain
            parallel (boolean) : compute parallel FX computation (alpha * input + (1-alpha) * manipulated output)
            parallel_weight_factor : the value of alpha for parallel FX computation. default=None : random value in between (0.0, 0.5)
    '''
    def build(name):
        """One processor from its effect name; the reference matches substrings in this order (:41-84)."""
        key = name.lower()
        if key == 'gain':
            return Gain()
        rules = (('eq', lambda: Equaliser(n_channels=2, sample_rate=sample_rate)),
                 ('comp', lambda: Compressor(sample_rate=sample_rate)),
                 ('expand', None),
                 ('pan', lambda: Panner()),
                 ('image', lambda: MidSideImager()),
                 ('algorithmic', lambda: AlgorithmicReverb(sample_rate=sample_rate)),
                 # convolution reverberation needs impulse responses; without a directory the algorithmic one is used
                 ('reverb', lambda: AlgorithmicReverb(sample_rate=sample_rate) if ir_dir_path is None
                  else ConvolutionalReverb(load_impulse_responses(ir_dir_path, sample_rate), sample_rate)))
        for fragment, make in rules:
            if fragment in key:
                if make is None:
                    raise NotImplementedError(f"effect {name!r}: the reference's factory builds `Expander(...)` here (:49), a class "
                                              "its common_audioeffects.py does not define (NameError there)")
                return make()
        raise ValueError(f"make sure the target effects are in the Augment FX chain : received fx called {name}")

    entries = []
    for item in effects:
        fx, prob = item if isinstance(item, tuple) else (item, 1)        # probability of applying the effect, 100 % by default
        if not isinstance(fx, (AugmentationChain, Processor)):
            fx = build(fx)
        # nested chains and the gain are never RMS re-normalised (:92)
        entries.append((fx, prob, not (isinstance(fx, AugmentationChain) or fx.name == 'Gain')))
    return AugmentationChain(fxs=entries, shuffle=shuffle, parallel=parallel, parallel_weight_factor=parallel_weight_factor)


def _one_shelf_equaliser(band, sample_rate):
    """An Equaliser with a single -50 dB shelf at 100 Hz: `high_shelf` = low pass, `low_shelf` = high pass (:123-146)."""
    params = ParameterList()
    params.add(Parameter(band + '_gain', -50.0, 'float', minimum=-50.0, maximum=-50.0))
    params.add(Parameter(band + '_freq', 100.0, 'float', minimum=100.0, maximum=100.0))
    return Equaliser(n_channels=2, sample_rate=sample_rate, bands=[band], parameters=params)


# create audio FX-chain according to input instrument
def create_inst_effects_augmentation_chain(inst,
                                           apply_prob_dict,
                                           ir_dir_path=None,
                                           algorithmic=False,
                                           sample_rate=44100):
    '''
        Args:
            inst (string) : FXmanipulator for target instrument. Only 'drums' is treated differently (reverberation)
            apply_prob_dict (dictionary of (FX name, probability)) : applying proababilities for each FX
            ir_dir_path (string) : directory path that contains directories of impulse responses organized according to RT60
            algorithmic (boolean) : rather to use algorithmic reverberation (True) or convolution reverberation (False)
            sample_rate (int) : using sampling rate
    '''
    def chain(effects, **kw):
        return create_effects_augmentation_chain(effects, ir_dir_path=ir_dir_path, sample_rate=sample_rate, **kw)

    reverb_type = 'algorithmic' if algorithmic else 'reverb'
    p = apply_prob_dict
    eq_comp_rand = chain([('eq', p['eq']), ('comp', p['comp'])], shuffle=True)
    pan_image_rand = chain([('pan', p['pan']), ('imager', p['imager'])], shuffle=True)
    if inst == 'drums':
        # reverberation mostly on the band above 100 Hz; on the band below it with 1 % of the probability
        reverb_low = chain([_one_shelf_equaliser('high_shelf', sample_rate), (reverb_type, p['reverb'] * 0.01)],
                           parallel=True, parallel_weight_factor=0.8)
        reverb_high = chain([_one_shelf_equaliser('low_shelf', sample_rate), (reverb_type, p['reverb'])],
                            parallel=True, parallel_weight_factor=0.6)
        reverb_parallel = chain([reverb_low, reverb_high])
    else:
        reverb_parallel = chain([(reverb_type, p['reverb'])], parallel=True)
    # full effects chain
    return chain([eq_comp_rand, pan_image_rand, reverb_parallel, ('gain', p['gain'])])
